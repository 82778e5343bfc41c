#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "init_conv" --maxfail=10 --tb=short --timeout=100 --timeout-method=thread > $OUT/pytest_r5p.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r5p.log | tail -8
for v in default xf1 xf1nf default xf1 xf1nf; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  timeout 100 python bench.py --timesteps 20 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']; print('$v fused %.2f us plain %.2f us  ms/iter %.4f' % (r['ms_per_launch']*1e3, r['plain_conv']['ms_per_launch']*1e3, d['ms_per_denoise_iteration']))"
done
