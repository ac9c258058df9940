mkdir -p gpurun_out
timeout 120 python tools/perf_decode.py 8 2 4 64 96 bf16 global > gpurun_out/tiny.log 2>&1 || { echo TINY_FAILED; tail -5 gpurun_out/tiny.log; exit 1; }
timeout 120 python tools/perf_decode.py 3000 1 1 64 96 f32 window >> gpurun_out/tiny.log 2>&1 || { echo TINY2_FAILED; tail -5 gpurun_out/tiny.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
( MVGEO_DECODE_VARIANT=0 timeout 200 python tools/perf_decode.py
  timeout 200 python tools/perf_decode.py
  for v in u2s4 u2s6 u1s6 u1s8 u4s2 u4s3 u2s2 u1s4; do MVGEO_LIB=$PWD/build/libmvgeo_$v.so timeout 200 python tools/perf_decode.py; done
  timeout 200 python tools/perf_decode.py 1024 4 8 240 320 bf16 none
  timeout 200 python tools/perf_decode.py 1024 4 8 240 320 bf16 window
  MVGEO_DECODE_VARIANT=0 timeout 200 python tools/perf_decode.py 1024 4 8 240 320 bf16 none
  timeout 200 python tools/perf_decode.py 512 4 8 240 320 f32 global
  MVGEO_DECODE_VARIANT=0 timeout 200 python tools/perf_decode.py 512 4 8 240 320 f32 global
  timeout 200 python tools/perf_decode.py 128 8 8 480 640 bf16 global
  MVGEO_DECODE_VARIANT=0 timeout 200 python tools/perf_decode.py 128 8 8 480 640 bf16 global
  timeout 200 python tools/perf_decode.py 8192 3 7 128 128 f32 global
  MVGEO_DECODE_VARIANT=0 timeout 200 python tools/perf_decode.py 8192 3 7 128 128 f32 global
  timeout 200 python tools/perf_decode.py 8 3 8 120 160 f32 global
) 2>&1 | grep -v Warning > gpurun_out/sweep.log
cat gpurun_out/sweep.log | sed -E 's#/[^ ]*/build/##' | cut -c1-170
