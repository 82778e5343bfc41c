#!/bin/bash
TAG="${1:-r2e}"
OUT=gpurun_out
mkdir -p $OUT
timeout 240 python -m pytest "tests/test_gpu_kernels.py::test_conv_zmarch_cta_pair_equals_single_cta" -m gpu -q -x --tb=short --timeout=60 --timeout-method=thread > $OUT/pytest_pair_$TAG.log 2>&1; rc=$?; echo "pytest(pair) rc=$rc"; tail -2 $OUT/pytest_pair_$TAG.log
timeout 200 python tools/bench_conv_gn.py > $OUT/conv_gn_pair_$TAG.jsonl 2>$OUT/conv_gn_pair_$TAG.err; cut -c1-230 $OUT/conv_gn_pair_$TAG.jsonl; tail -3 $OUT/conv_gn_pair_$TAG.err
DIQT_ZM_2CTA=0 timeout 200 python tools/bench_conv_gn.py > $OUT/conv_gn_single_$TAG.jsonl 2>$OUT/conv_gn_single_$TAG.err; cut -c1-230 $OUT/conv_gn_single_$TAG.jsonl
for cfg in "" "DIQT_ZM_2CTA=0"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline 2>$OUT/ab_$TAG.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$cfg] rep$rep ms/iter %.4f zm_us %.2f frac %.3f' % (d['ms_per_denoise_iteration'], d['roofline']['ms_per_launch']*1e3, d['roofline']['frac']))
except Exception as e:
    print('[$cfg] failed', e, open('$OUT/ab_$TAG.err').read()[-500:])
"
  done
done
timeout 900 python tools/bench_sweep.py conv > $OUT/sweep_conv_$TAG.jsonl 2> $OUT/sweep_conv_$TAG.err; cut -c1-420 $OUT/sweep_conv_$TAG.jsonl; tail -3 $OUT/sweep_conv_$TAG.err
