#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_config_size.py -m gpu -q --maxfail=20 --tb=short --timeout=300 --timeout-method=thread -k "train or wgrad or gradients or groupnorm or p_losses or adam" > $OUT/pytest_r4a.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r4a.log | tail -10
grep -E "^E  " $OUT/pytest_r4a.log | head -10
for shape in "64 1" "32 27"; do timeout 600 python tools/bench_train.py $shape 2>/dev/null | cut -c1-500; done | tee $OUT/bench_train_r4a.jsonl
