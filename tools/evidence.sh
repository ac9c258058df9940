#!/bin/bash
# Everything profiles/ cites, from ONE box: `gpurun --timeout 1500 -- 'bash tools/evidence.sh'`, results under gpurun_out/ev_*.
# Order: plain runs first (bench values are never taken under a profiler), then the ncu passes of the SAME commands.
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -s > $O/ev_pytest.log 2>&1; echo "pytest rc=$?" >> $O/ev_pytest.log )
grep -E "passed|failed|rc=|FAILED|^E  |regime sweep" $O/ev_pytest.log | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err; echo "bench rc=$?"
rm -f $O/ev_bench_workloads.jsonl
for w in c1 c2 c3 c5; do timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline >> $O/ev_bench_workloads.jsonl 2>> $O/ev_bench_n1.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/ev_bench_reference.json 2>> $O/ev_bench_n1.err
rm -f $O/ev_regimes.jsonl
timeout 400 python tools/decode_regimes.py --tag r02 --shapes c2,native32,native16,c5 --out $O/ev_regimes.jsonl > $O/ev_regimes.log 2>&1
[ -f build/libmvgeo_r01.so ] && timeout 600 python tools/decode_regimes.py --tag r01 --lib build/libmvgeo_r01.so --shapes c2,native32,native16 --iters 3 --out $O/ev_regimes.jsonl >> $O/ev_regimes.log 2>&1
timeout 300 python tools/shim_latency.py > $O/ev_shim_latency.log 2>> $O/ev_bench_n1.err && cp $O/shim_latency.json $O/ev_shim_latency.json
timeout 300 python examples/ddp_train_step.py --steps 10 --bf16 > $O/ev_c4_n1.json 2>> $O/ev_bench_n1.err; echo "c4 rc=$?"
# profiler passes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --quick > $O/ev_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tma -s 3 -c 1 -f -o $O/ev_decode_c2 \
  python tools/decode_regimes.py --shapes c2 --only blob:1.0:100.0 --iters 1 --out $O/ev_scratch.jsonl > $O/ev_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tma -s 3 -c 1 -f -o $O/ev_decode_native16 \
  python tools/decode_regimes.py --shapes native16 --only blob:1.0:100.0 --iters 1 --out $O/ev_scratch.jsonl > $O/ev_ncu_n16.log 2>&1; echo "ncu native16 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/ev_bench_n1.json",):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"], d["e2e"]["value"], d["clocks"])
for l in open("gpurun_out/ev_bench_workloads.jsonl"):
    d = json.loads(l); print(d["config"]["workload"][:3], round(d["value"]), d["roofline"]["achieved"], d["roofline"]["frac"])
rows = {}
for l in open("gpurun_out/ev_regimes.jsonl"):
    d = json.loads(l); rows.setdefault((d["lib"], d["shape"]), []).append(round(d["gbs"]))
for k, x in rows.items():
    print(k, min(x), max(x), sum(x) // len(x))
PY
