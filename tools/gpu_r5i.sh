#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file $OUT/launches_r5i.csv \
  python bench.py --timesteps 3 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step > $OUT/ncu_launch_r5i.log 2>&1; echo "ncu launches rc=$?"
