#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_unet.py -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread -k "attn or attention" > $OUT/pytest_r3m.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r3m.log | tail -15
grep -E "^E " $OUT/pytest_r3m.log | head -20
timeout 300 python tools/bench_sweep.py attn > $OUT/sweep_attn_r3m.jsonl 2> $OUT/sweep_attn_r3m.err; echo "sweep rc=$?"
cut -c1-420 $OUT/sweep_attn_r3m.jsonl; tail -5 $OUT/sweep_attn_r3m.err
