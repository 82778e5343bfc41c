#!/bin/bash
# Round-2 first GPU pass: the new parity tests, smoke, a short bench with the new sub-records, the cfg5 sweep.
TAG="${1:-r2a}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_config_size.py tests/test_gpu_volume.py "tests/test_gpu_kernels.py::test_conv_zmarch_bf16" -m gpu -q -s --maxfail=20 --tb=short --timeout=300 --timeout-method=thread > $OUT/pytest_new_$TAG.log 2>&1; echo "pytest(new) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|rel-L2|literal metric|config-4" $OUT/pytest_new_$TAG.log | tail -40
timeout 200 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/smoke_$TAG.log
timeout 600 python bench.py --steps 2 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
timeout 900 python tools/bench_sweep.py > $OUT/sweep_$TAG.jsonl 2> $OUT/sweep_$TAG.err; echo "sweep rc=$?"; cat $OUT/sweep_$TAG.jsonl; tail -5 $OUT/sweep_$TAG.err
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 --tb=short --timeout=300 --timeout-method=thread --deselect tests/test_gpu_config_size.py --deselect tests/test_gpu_volume.py > $OUT/pytest_all_$TAG.log 2>&1; echo "pytest(rest) rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_all_$TAG.log | tail -20
du -sh $OUT
