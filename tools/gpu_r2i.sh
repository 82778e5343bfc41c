#!/bin/bash
TAG="${1:-r2i}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_attn.py" -m gpu -q --tb=short --timeout=90 --timeout-method=thread --maxfail=6 2>&1 | tail -3
timeout 300 python tools/bench_sweep.py attn linattn > $OUT/sweep_attn_$TAG.jsonl 2>$OUT/sweep_attn_$TAG.err; cut -c1-330 $OUT/sweep_attn_$TAG.jsonl; tail -3 $OUT/sweep_attn_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_attn_tc2_kernel.2 -c 1 -f -o $OUT/prof_attn_$TAG python tools/bench_sweep.py attn > /dev/null 2>&1; echo "ncu attn rc=$?"
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page raw --csv > $OUT/prof_attn_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_attn_${TAG}_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_attn_${TAG}_raw.csv $OUT/ncu_attn_${TAG}.csv; cut -c1-600 $OUT/ncu_attn_${TAG}.csv
PYTEST_TIMEOUT=900 SKIP_NCU=1 bash tools/gpu_check.sh $TAG
