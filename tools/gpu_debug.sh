#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 120 python tools/volume_diag.py fp32
  timeout 120 python tools/volume_diag.py fp32 sync
  DIQT_DISABLE_CUDA_GRAPH=1 timeout 120 python tools/volume_diag.py fp32
  DIQT_DISABLE_PDL=1 timeout 120 python tools/volume_diag.py fp32
  timeout 120 python tools/volume_diag.py bf16
  DIQT_DISABLE_PDL=1 timeout 120 python tools/volume_diag.py bf16 ) > $OUT/volume_diag.log 2>&1
cat $OUT/volume_diag.log | cut -c1-900
