#!/bin/bash
# Quick GPU pass: kernel + unet + sampler tests, then short A/B benches.  Usage: gpurun -- 'bash tools/gpu_quick.sh <tag> [pytest args]'
TAG="${1:-q}"; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 420 python -m pytest tests -m gpu -x -q --timeout=90 --timeout-method=thread "$@" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_$TAG.log
for pdl in 0 1; do
  DIQT_DISABLE_PDL=$pdl timeout 150 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_${TAG}_nopdl$pdl.json 2> $OUT/bench_${TAG}_nopdl$pdl.err
  echo "DISABLE_PDL=$pdl rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_nopdl$pdl.json").read().strip().splitlines()[-1])
    print("ms/iter", d["ms_per_denoise_iteration"], "zm", d["roofline"]["ms_per_launch"], d["roofline"]["achieved"], "whole", d["roofline"]["whole_step"]["frac"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_${TAG}_nopdl$pdl.err").read()[-2000:])
PY
done
