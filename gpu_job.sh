timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python - <<'PY'
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import probe_perf as p
for B in (1, 24, 64, 74, 148, 296):
    p.run('C2', B, 'outer')
p.run('C2', 148, 'pointwise')
PY
for cs in 1 2 8; do LCS_OUTER_CLUSTER=$cs timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import probe_perf as p
p.run('C2', 64, 'outer')
PY
done
