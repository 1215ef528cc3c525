timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'scripts'); sys.path.insert(0,'.')
import probe_perf as p
for B in (8, 19, 37, 64, 74, 100, 148, 200, 296, 400):
    p.run('C2', B, 'outer')
PY
for c in 74 148; do python bench.py --no-cpu-baseline --chunk $c --steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $c', 'value %.4e e2e %.4e ms %.2f e2e_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"; done
