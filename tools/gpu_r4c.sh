#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_config_size.py -m gpu -q --maxfail=20 --tb=short --timeout=300 --timeout-method=thread -k "train or wgrad or gradients or groupnorm or p_losses or adam or checkpoint" > $OUT/pytest_r4c.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r4c.log | tail -10
grep -E "^E  " $OUT/pytest_r4c.log | head -10
