#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_config_size.py -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread > $OUT/pytest_r4g.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r4g.log | tail -5
for v in default head; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  timeout 300 python tools/bench_sweep.py conv 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    if d.get('op','').startswith('conv') and d['ours_ms'] < 0.06: print('$v', d.get('channels'), 'ours %.1f us torch %.1f us' % (d['ours_ms']*1e3, min(x for x in d['torch_ms'].values() if x)*1e3))
"
  timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v ms/iter %.4f' % d['ms_per_denoise_iteration'])"
done
