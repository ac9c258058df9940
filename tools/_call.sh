mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "decode or pipeline" 2>&1 | tail -3
timeout 200 python tools/perf_decode.py
timeout 200 python tools/perf_decode.py 1024 4 8 240 320 bf16 none
timeout 200 python tools/perf_decode.py 1024 4 8 240 320 bf16 window
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1900
