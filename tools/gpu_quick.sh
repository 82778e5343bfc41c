#!/bin/bash
# Quick GPU pass: parity tests (optionally a subset), then short A/B benches of build variants (default + $VARIANTS).
# Usage: gpurun -- 'VARIANTS="nc_late" bash tools/gpu_quick.sh <tag> [pytest args]'
TAG="${1:-q}"; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread "$@" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | tail -15
for v in default ${VARIANTS:-}; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  timeout 200 python bench.py --timesteps ${TIMESTEPS:-300} --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    print("$v ms/iter %.4f  zm_us %.2f  whole %.4f  e2e %.4f" % (d["ms_per_denoise_iteration"], d["roofline"]["ms_per_launch"] * 1e3, d["roofline"]["whole_step"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print("$v bench failed", e, open("$OUT/bench_${TAG}_$v.err").read()[-800:])
PY
done
if [ -n "${LAUNCHES:-}" ]; then
unset DIQT_LIB_PATH
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --timesteps 3 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
fi
