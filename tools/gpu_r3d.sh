#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/bench_train.py 64 1 > $OUT/bench_train_r3d.jsonl 2> $OUT/bench_train_r3d.err; echo rc=$?; cat $OUT/bench_train_r3d.jsonl; tail -3 $OUT/bench_train_r3d.err
timeout 600 python tools/bench_train.py 32 8 >> $OUT/bench_train_r3d.jsonl 2>> $OUT/bench_train_r3d.err; echo rc=$?; tail -1 $OUT/bench_train_r3d.jsonl
