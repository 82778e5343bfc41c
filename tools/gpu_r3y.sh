#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_train.py -m gpu -q --maxfail=20 --tb=short --timeout=300 --timeout-method=thread -k "deconv or train or wgrad or gradients" > $OUT/pytest_r3y.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r3y.log | tail -20
grep -E "^E  " $OUT/pytest_r3y.log | head -20
timeout 600 python tools/bench_train.py 64 1 > $OUT/bench_train_r3y.jsonl 2> $OUT/bench_train_r3y.err; echo rc=$?; cat $OUT/bench_train_r3y.jsonl; tail -3 $OUT/bench_train_r3y.err
