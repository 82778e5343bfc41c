#!/bin/bash
TAG="${1:-r2j}"
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc2_kernel -s 2 -c 1 -f -o $OUT/prof_attn_$TAG python tools/bench_sweep.py attn > $OUT/ncu_attn_$TAG.log 2>&1; echo "ncu attn rc=$?"
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page raw --csv > $OUT/prof_attn_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_attn_${TAG}_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_attn_${TAG}_raw.csv $OUT/ncu_attn_${TAG}.csv; cut -c1-700 $OUT/ncu_attn_${TAG}.csv
timeout 900 python tools/bench_configs.py 1 4 32 > $OUT/bench_configs_$TAG.jsonl 2>$OUT/bench_configs_$TAG.err; cat $OUT/bench_configs_$TAG.jsonl; tail -3 $OUT/bench_configs_$TAG.err
