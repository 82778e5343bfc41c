#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:init_conv_tc -s 10 -c 2 -f -o $OUT/prof_init_r5f python tools/bench_init.py 64 64 > $OUT/ncu_init_r5f.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_init_r5f.ncu-rep --page raw --csv > $OUT/prof_init_r5f_raw.csv 2>/dev/null
ncu -i $OUT/prof_init_r5f.ncu-rep --page source --csv --print-source sass > $OUT/prof_init_r5f_source.csv 2>/dev/null
ncu -i $OUT/prof_init_r5f.ncu-rep --page details > $OUT/prof_init_r5f_details.txt 2>/dev/null
rm -f $OUT/*.ncu-rep
ls -la $OUT | tail -5
