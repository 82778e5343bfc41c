#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "init_conv" --maxfail=10 --tb=short --timeout=100 --timeout-method=thread > $OUT/pytest_r5a.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r5a.log | tail -8
for v in default head; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== $v"
  timeout 60 python tools/bench_init.py 64 64 2>&1 | tail -1
  timeout 60 python tools/bench_init.py 32 64 2>&1 | tail -1
  if [ $v = default ]; then
    DIQT_INIT_TY=4 timeout 60 python tools/bench_init.py 64 64 2>&1 | tail -1
    DIQT_INIT_TY=2 timeout 60 python tools/bench_init.py 64 64 2>&1 | tail -1
  fi
  timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v ms/iter %.4f' % d['ms_per_denoise_iteration'])"
done
unset DIQT_LIB_PATH
timeout 400 python -m pytest tests/test_gpu_unet.py tests/test_gpu_sampler.py tests/test_gpu_config_size.py -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread > $OUT/pytest_r5a2.log 2>&1; echo "pytest2 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r5a2.log | tail -8
