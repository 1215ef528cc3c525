python bench.py > gpurun_out/bench_outer.json 2> gpurun_out/bench_outer.err
python bench.py --xclamp pointwise --no-cpu-baseline > gpurun_out/bench_pointwise.json 2>&1
