mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n$n.log 2>&1
tail -1 gpurun_out/bench_n$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['decode_ms'], d['breakdown'], d['clocks'])"
done
