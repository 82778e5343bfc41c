#!/bin/bash
# Fused-GroupNorm conv: kernel tests first (bounded), then the whole suite, bench, launch list, ncu of conv_zm.
TAG="${1:-r2b}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_kernels.py::test_conv_zmarch_fused_groupnorm_film_mish" "tests/test_gpu_kernels.py::test_conv_zmarch_bf16" tests/test_gpu_unet.py -m gpu -q -x --tb=short --timeout=120 --timeout-method=thread > $OUT/pytest_k_$TAG.log 2>&1; rc=$?; echo "pytest(kernels) rc=$rc"
grep -E "^(FAILED|ERROR)|passed|failed|Error" $OUT/pytest_k_$TAG.log | tail -20
if [ $rc -ne 0 ]; then tail -60 $OUT/pytest_k_$TAG.log; exit 0; fi
bash tools/gpu_check.sh $TAG
python tools/ncu_summary.py $OUT/prof_zm_${TAG}_raw.csv $OUT/ncu_zm_${TAG}.csv
python tools/launch_summary.py $OUT/launches_$TAG.csv 0 > $OUT/launch_summary_$TAG.txt 2>&1; head -30 $OUT/launch_summary_$TAG.txt
