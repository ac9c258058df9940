mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest.log; tail -2 gpurun_out/pytest.log
( for g in 1 2 4; do MVGEO_DECODE_GROUPS=$g timeout 200 python tools/perf_decode.py; done
  timeout 200 python tools/perf_decode.py 2048 4 7 240 320 bf16 global
  timeout 200 python tools/perf_decode.py 8 3 8 120 160 f32 global ) 2>&1 | grep -v Warning | cut -c1-150
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-2200
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_tma -s 3 -c 2 -o gpurun_out/prof_decode_tma python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"dlt_kernel|fk_reproj" -s 6 -c 2 -o gpurun_out/prof_geom python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
ls gpurun_out
