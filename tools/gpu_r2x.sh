#!/bin/bash
# linear attention on the tensor cores: parity tests, then the config-5 sweep lines
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_unet.py -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread -k "linear" > $OUT/pytest_r2x.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r2x.log | tail -15
grep -E "^E " $OUT/pytest_r2x.log | head -20
timeout 300 python tools/bench_sweep.py linattn > $OUT/sweep_linattn_r2x.jsonl 2> $OUT/sweep_linattn_r2x.err; echo "sweep rc=$?"
cut -c1-500 $OUT/sweep_linattn_r2x.jsonl; tail -5 $OUT/sweep_linattn_r2x.err
