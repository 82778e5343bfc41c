#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { ( env "$@" timeout 120 python tools/race_diag.py 2>&1 | tail -1 ) >> $OUT/race_diag.log; }
: > $OUT/race_diag.log
run A=1
run NOISE=gpu
run NOISE=pinned
run HOSTLISTS=0
run DIQT_DISABLE_GROUPED=1
run DIQT_DISABLE_PDL=1
run DIQT_DISABLE_PDL=1 DIQT_DISABLE_GROUPED=1
run DIQT_DEBUG_SYNC_REPLAY=before
run DIQT_DEBUG_SYNC_REPLAY=after
run DTYPE=bf16
run DTYPE=bf16 DIQT_DISABLE_GROUPED=1 DIQT_DISABLE_PDL=1
cat $OUT/race_diag.log
