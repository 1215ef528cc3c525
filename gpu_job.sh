python -m pytest tests -m gpu -q 2>&1 | tail -30
python bench.py --steps 3 --warmup 2 --xclamp pointwise --no-cpu-baseline 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -3
python scripts/probe_perf.py 2>&1 | grep -E '"B": (1|64), "xmode": "(pointwise|outer)", "order": 3, "pair": "f64", "strict": false' | head -4
python bench.py --steps 2 --warmup 1 --xclamp pointwise --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:advect_fused -s 1 -c 1 -o gpurun_out/prof_fused -f python bench.py --steps 2 --warmup 1 --xclamp pointwise --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
