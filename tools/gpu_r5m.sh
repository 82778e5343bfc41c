#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in default rs512 default rs512; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  timeout 100 python tools/bench_residual.py 2>/dev/null | head -4 | cut -c1-230
  timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v ms/iter %.4f residual %.2f us' % (d['ms_per_denoise_iteration'], d['roofline_elementwise']['ms_per_launch']*1e3))"
done
export DIQT_LIB_PATH=$PWD/build/variants/rs512.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "residual or se_" --maxfail=5 --tb=short --timeout=100 --timeout-method=thread 2>&1 | tail -3
