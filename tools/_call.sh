mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "training_step or shard" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1
tail -1 gpurun_out/bench_n2.log | cut -c1-1300
python examples/ddp_train_step.py --steps 10 > gpurun_out/ddp_n1.log 2>&1; tail -1 gpurun_out/ddp_n1.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 examples/ddp_train_step.py --steps 10 > gpurun_out/ddp_n2.log 2>&1; tail -2 gpurun_out/ddp_n2.log | cut -c1-600
