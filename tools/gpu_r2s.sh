#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/bench_residual.py > $OUT/bench_residual_r2s.jsonl 2>$OUT/bench_residual_r2s.err; cat $OUT/bench_residual_r2s.jsonl; tail -3 $OUT/bench_residual_r2s.err
timeout 200 python tools/op_timing.py 64 10 > $OUT/op_timing_r2s.txt 2>&1; head -50 $OUT/op_timing_r2s.txt
