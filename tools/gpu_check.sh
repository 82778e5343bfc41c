#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list, ncu full capture of the conv kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag>'
TAG="${1:-run}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q --timeout=90 --timeout-method=thread > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 150 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
timeout 240 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
if [ -z "${SKIP_NCU:-}" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --timesteps 3 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_zm_kernel -s 39 -c 39 -f -o $OUT/prof_zm_$TAG \
  python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_zm_$TAG.log 2>&1; echo "ncu zm rc=$?"
if [ -n "${NCU_TC:-}" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 28 -c 28 -f -o $OUT/prof_tc_$TAG \
  python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_tc_$TAG.log 2>&1; echo "ncu tc rc=$?"
fi
fi
ls -la $OUT
