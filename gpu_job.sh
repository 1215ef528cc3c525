timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_outer.json 2> gpurun_out/bench_outer.err
python bench.py --xclamp pointwise > gpurun_out/bench_pointwise.json 2>&1
python bench.py --precision f32 --no-cpu-baseline > gpurun_out/bench_f32.json 2>&1
python bench.py --order 1 --no-cpu-baseline > gpurun_out/bench_p1.json 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:advect_outer_cluster -s 1 -c 1 -o gpurun_out/prof_cluster2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/launches_outer.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -1 gpurun_out/ncu2.log
