#!/bin/bash
# CTA-pair z-march kernel: bounded parity tests first; then A/B against DIQT_ZM_2CTA=0, full suite, bench, ncu.
TAG="${1:-r2d}"
OUT=gpurun_out
mkdir -p $OUT
timeout 240 python -m pytest "tests/test_gpu_kernels.py::test_conv_zmarch_cta_pair_equals_single_cta" -m gpu -q -x --tb=short --timeout=60 --timeout-method=thread > $OUT/pytest_pair_$TAG.log 2>&1; rc=$?; echo "pytest(pair) rc=$rc"
grep -E "^(FAILED|ERROR)|passed|failed|Error|Timeout" $OUT/pytest_pair_$TAG.log | tail -20
if [ $rc -ne 0 ]; then
  tail -50 $OUT/pytest_pair_$TAG.log
  nvidia-smi > $OUT/smi_after_fail_$TAG.txt 2>&1
  echo "PAIR KERNEL FAILED: continuing with DIQT_ZM_2CTA=0"
  export DIQT_ZM_2CTA=0
fi
timeout 200 python tools/bench_conv_gn.py > $OUT/conv_gn_pair_$TAG.jsonl 2>$OUT/conv_gn_pair_$TAG.err; cat $OUT/conv_gn_pair_$TAG.jsonl; tail -3 $OUT/conv_gn_pair_$TAG.err
DIQT_ZM_2CTA=0 timeout 200 python tools/bench_conv_gn.py > $OUT/conv_gn_single_$TAG.jsonl 2>$OUT/conv_gn_single_$TAG.err; cat $OUT/conv_gn_single_$TAG.jsonl
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
PYTEST_TIMEOUT=900 bash tools/gpu_check.sh $TAG
python tools/ncu_summary.py $OUT/prof_zm_${TAG}_raw.csv $OUT/ncu_zm_${TAG}.csv; cat $OUT/ncu_zm_${TAG}.csv
python tools/launch_summary.py $OUT/launches_$TAG.csv 0 x > $OUT/launch_summary_$TAG.txt 2>&1; head -24 $OUT/launch_summary_$TAG.txt
