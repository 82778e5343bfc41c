#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/bench_train.py 32 27 > $OUT/bench_train_b27_r3z.jsonl 2> $OUT/bench_train_b27_r3z.err; echo rc=$?; cat $OUT/bench_train_b27_r3z.jsonl; tail -2 $OUT/bench_train_b27_r3z.err | cut -c1-300
BENCH_TRAIN_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 9000 --csv --log-file $OUT/launches_train_b27_r3z.csv python tools/bench_train.py 32 27 > $OUT/ncu_train_b27_r3z.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py $OUT/launches_train_b27_r3z.csv 2>/dev/null | head -24
