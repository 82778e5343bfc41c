#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
# the convs of the 32^3 and 16^3 levels as the step launches them: conv_zm<1,0> (fused, single CTA) and conv_tc (split-K, 1x1x1, up / down)
timeout 600 ncu --set full --clock-control none -k regex:"conv_zm_kernel<\(bool\)1, \(bool\)0>|conv_tc_kernel" -s 60 -c 26 -f -o $OUT/prof_small_r5u \
  python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step > $OUT/ncu_small_r5u.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_small_r5u.ncu-rep --page raw --csv > $OUT/prof_small_r5u_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls -la $OUT/prof_small_r5u_raw.csv
