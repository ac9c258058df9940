#!/bin/bash
# usage: tools/run_multi.sh N [dev|full|short]   -- configs 2-5 (short: 2 and 3 only) at N GPUs of one box (one rank per GPU), logs in gpurun_out/
N=${1:-1}; MODE=${2:-full}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
if [ "$MODE" = dev ]; then C3F=$((4096*N)); C5F=2048; ST=5; else C3F=65536; C5F=1048576; ST=20; fi
T=gpurun_out/multi_n${N}_${MODE}
echo "== C2 bench (weak)"; timeout 900 $RUN bench.py --gpus $N --steps $ST --warmup 3 --no-cpu-baseline > ${T}_c2.log 2>&1; echo rc=$?; grep '^{' ${T}_c2.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'h2d/rank',d['breakdown']['h2d_gbs_per_rank_min'],'numa',d['breakdown']['numa_bound_cpus'],'gather_ms',d['breakdown']['final_gather_ms'])"
echo "== C3 job (strong)"; timeout 900 $RUN bench.py --workload c3 --job --job-frames $C3F --gpus $N > ${T}_c3.log 2>&1; echo rc=$?; grep '^{' ${T}_c3.log | tail -1 | cut -c1-400
[ "$MODE" = short ] && exit 0
echo "== C5 sweep (strong)"; timeout 1200 $RUN bench.py --workload c5 --job --job-frames $C5F --gpus $N > ${T}_c5.log 2>&1; echo rc=$?; grep '^{' ${T}_c5.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['views'],d['keypoints'],round(d['frames_per_s']),round(d['decode_tri_gbs_per_gpu']))"
echo "== C4 DDP step"; timeout 900 $RUN examples/ddp_train_step.py --steps $ST --bf16 > ${T}_c4.log 2>&1; echo rc=$?; grep '^{' ${T}_c4.log | tail -1 | cut -c1-1200
