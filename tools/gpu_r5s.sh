#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "split_k or tcgen05" --tb=short --timeout=60 --timeout-method=thread > $OUT/pytest_r5s.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r5s.log | tail -8
tail -5 $OUT/pytest_r5s.log | cut -c1-300
for v in default head; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  timeout 200 python tools/bench_sweep.py conv 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    if d.get('op','').startswith('conv') and d['side'] == 16: print('$v', d.get('channels'), 'ours %.1f us torch %.1f us' % (d['ours_ms']*1e3, min(x for x in d['torch_ms'].values() if x)*1e3))
"
  timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v ms/iter %.4f' % d['ms_per_denoise_iteration'])"
done
