#!/bin/bash
# training step: parity of the backward kernels and of the whole reverse pass against autograd of the oracle
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --maxfail=30 --tb=short --timeout=300 --timeout-method=thread "$@" > $OUT/pytest_r3b.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r3b.log | tail -40
grep -E "^E  " $OUT/pytest_r3b.log | head -60
