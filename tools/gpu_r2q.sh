#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "" "DIQT_GN_FUSION_MIN=0" "DIQT_GN_FUSION_MIN=2097152" "DIQT_GN_FUSION_MIN=8388608" "DIQT_ZM_2CTA=0" "DIQT_ZM_2CTA_MIN_PAIRS=4"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline 2>$OUT/ab_q.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$cfg] rep$rep ms/iter %.4f' % (d['ms_per_denoise_iteration']))
except Exception as e:
    print('[$cfg] failed', e, open('$OUT/ab_q.err').read()[-800:])
"
  done
done
