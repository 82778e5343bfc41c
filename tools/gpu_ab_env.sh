#!/bin/bash
# A/B of runtime knobs (environment variables) with the short bench.  Usage: gpurun -- 'bash tools/gpu_ab_env.sh "A=1 B=2" "A=3" ...'
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/ab_env.log
for cfg in "" "$@"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ab_env.json 2> $OUT/ab_env.err
    python - >> $OUT/ab_env.log <<PY
import json
try:
    d = json.loads(open("$OUT/ab_env.json").read().strip().splitlines()[-1])
    print("[%s] rep$rep ms/iter %.4f" % ("$cfg" or "default", d["ms_per_denoise_iteration"]))
except Exception as e:
    print("[$cfg] failed", e, open("$OUT/ab_env.err").read()[-500:])
PY
  done
done
cat $OUT/ab_env.log
