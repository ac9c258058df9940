mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n2.log 2>&1
tail -2 gpurun_out/bench_n2.log | cut -c1-230
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
tail -1 gpurun_out/bench_n1.log | cut -c1-230
