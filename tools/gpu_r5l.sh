#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 700 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_volume.py tests/test_gpu_config_size.py tests/test_gpu_elucidated.py -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread > $OUT/pytest_r5l.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r5l.log | tail -8
for v in 1 0 1 0; do
  DIQT_NOISE_IN_GRAPH=$v timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('noise_in_graph=$v ms/iter %.4f e2e %.4f' % (d['ms_per_denoise_iteration'], d['e2e']['value']))"
done
