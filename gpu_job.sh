timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import probe_perf as p
p.run('C2', 296, 'outer')
p.run('C2', 296, 'pointwise')
p.run('C2', 296, 'pointwise', pair='f32')
p.run('C2', 296, 'outer', pair='f32')
p.run('C2', 1, 'pointwise')
p.run('C3', 1, 'pointwise')
PY
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('outer value %.4e e2e %.4e ms %.2f e2e_ms %.2f frac %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
