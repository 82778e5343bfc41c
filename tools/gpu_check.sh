#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list, ncu full capture of the dominant conv kernel.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag>'
# gpurun_out/ must stay below 64 MiB or NOTHING is copied back: .ncu-rep files are exported to CSV and deleted on the box.
TAG="${1:-run}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests -m gpu -q --maxfail=${MAXFAIL:-12} --tb=short --timeout=150 --timeout-method=thread ${PYTEST_ARGS:-} > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | tail -30
timeout 150 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
if [ -z "${SKIP_BENCH:-}" ]; then
timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
fi
if [ -z "${SKIP_NCU:-}" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --timesteps 3 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_zm_kernel -s 40 -c 3 -f -o $OUT/prof_zm_$TAG \
  python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline > $OUT/ncu_zm_$TAG.log 2>&1; echo "ncu zm rc=$?"
ncu -i $OUT/prof_zm_$TAG.ncu-rep --page raw --csv > $OUT/prof_zm_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_zm_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_zm_${TAG}_source.csv 2>/dev/null
if [ -n "${NCU_TC:-}" ]; then
timeout 600 ncu --set full --clock-control none -k regex:"${NCU_TC}" -c ${NCU_TC_COUNT:-6} -f -o $OUT/prof_x_$TAG \
  python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline > $OUT/ncu_x_$TAG.log 2>&1; echo "ncu x rc=$?"
ncu -i $OUT/prof_x_$TAG.ncu-rep --page raw --csv > $OUT/prof_x_${TAG}_raw.csv 2>/dev/null
fi
rm -f $OUT/*.ncu-rep
fi
du -sh $OUT; ls -la $OUT
