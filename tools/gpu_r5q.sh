#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in zmtrace zmtrace_xf1 zmtrace_xf2; do
  export DIQT_LIB_PATH=$PWD/build/variants/$v.so
  echo "=== $v" >> $OUT/zm_timeline_r5q.txt
  timeout 100 python tools/zm_trace_graph.py fused >> $OUT/zm_timeline_r5q.txt 2>&1
done
DIQT_LIB_PATH=$PWD/build/variants/zmtrace.so timeout 100 python tools/zm_trace_graph.py plain >> $OUT/zm_timeline_r5q.txt 2>&1
tail -5 $OUT/zm_timeline_r5q.txt
