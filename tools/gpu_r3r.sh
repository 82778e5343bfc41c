#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_volume.py -m gpu -q --maxfail=5 --tb=short --timeout=600 --timeout-method=thread -k "trained" -s > $OUT/pytest_r3r.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|trained weights|literal" $OUT/pytest_r3r.log | tail -10
grep -E "^E  " $OUT/pytest_r3r.log | head -20
