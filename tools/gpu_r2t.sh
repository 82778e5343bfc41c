#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_kernels.py::test_se_scale_residual_ring_kernel" "tests/test_gpu_kernels.py::test_se_scale_residual" tests/test_gpu_unet.py tests/test_gpu_sampler.py -m gpu -q --tb=short --timeout=120 --timeout-method=thread --maxfail=6 > $OUT/pytest_ring_r2t.log 2>&1; rc=$?; echo "pytest(ring) rc=$rc"; grep -E "^(FAILED|ERROR)|passed|failed|Timeout" $OUT/pytest_ring_r2t.log | tail
if [ $rc -ne 0 ]; then tail -40 $OUT/pytest_ring_r2t.log; fi
timeout 200 python tools/bench_residual.py 2>&1 | cut -c1-300
DIQT_DISABLE_RESIDUAL_RING=1 timeout 200 python tools/bench_residual.py 2>&1 | cut -c1-200 | head -2
for cfg in "" "DIQT_DISABLE_RESIDUAL_RING=1"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline 2>$OUT/ab_t.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$cfg] rep$rep ms/iter %.4f' % (d['ms_per_denoise_iteration']))
except Exception as e:
    print('[$cfg] failed', e, open('$OUT/ab_t.err').read()[-800:])
"
  done
done
