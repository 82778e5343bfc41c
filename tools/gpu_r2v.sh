#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export DIQT_LIB_PATH=$PWD/build/variants/zmtrace.so
for a in "fused pair" "fused single" "plain pair" "plain single"; do timeout 100 python tools/zm_trace.py $a 2>&1 | tail -20; done > $OUT/zm_timeline_r2v.txt
cat $OUT/zm_timeline_r2v.txt
