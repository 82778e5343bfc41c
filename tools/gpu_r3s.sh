#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_config_size.py -m gpu -q --maxfail=5 --tb=short --timeout=600 --timeout-method=thread -k "training" > $OUT/pytest_r3s.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r3s.log | tail -10
grep -E "^E  " $OUT/pytest_r3s.log | head -20
