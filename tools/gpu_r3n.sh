#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
BENCH_TRAIN_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 9000 --csv --log-file $OUT/launches_train_r3n.csv python tools/bench_train.py 64 1 > $OUT/ncu_train_r3n.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py $OUT/launches_train_r3n.csv | head -36
