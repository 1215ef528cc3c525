timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python - <<'PY'
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import probe_perf as p
for B in (64, 148, 296):
    p.run('C2', B, 'outer')
p.run('C2', 296, 'pointwise')
p.run('C2', 1, 'outer')
p.run('C2', 1, 'pointwise')
PY
