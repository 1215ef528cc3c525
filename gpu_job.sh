python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_2gpu.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2>/dev/null
