#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -q -x --tb=short --timeout=100 --timeout-method=thread > $OUT/pytest_r6d.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r6d.log | tail -8
for v in default head default head; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']; print('$v ms/iter %.4f fused %.2f us plain %.2f us' % (d['ms_per_denoise_iteration'], r['ms_per_launch']*1e3, r['plain_conv']['ms_per_launch']*1e3))"
done
