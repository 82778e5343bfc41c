#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
# one split-K conv_tc launch of the 16^3 level (grid 128) with source-level sampling: launches 61.. of the step are conv_tc (16^3 level)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^conv_tc_kernel" -s 39 -c 8 -f -o $OUT/prof_tc16_r5v \
  python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step > $OUT/ncu_tc16_r5v.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_tc16_r5v.ncu-rep --page raw --csv > $OUT/prof_tc16_r5v_raw.csv 2>/dev/null
ncu -i $OUT/prof_tc16_r5v.ncu-rep --page source --csv --print-source sass > $OUT/prof_tc16_r5v_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls -la $OUT/prof_tc16_r5v*
