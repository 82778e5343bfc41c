#!/bin/bash
for v in "A" "B" "C" "A" "B" "C"; do
  case $v in
    A) export DIQT_TC_SMALL=1; unset DIQT_GN_FUSION_MIN;;
    B) export DIQT_TC_SMALL=0; unset DIQT_GN_FUSION_MIN;;
    C) export DIQT_TC_SMALL=0; export DIQT_GN_FUSION_MIN=262144;;
  esac
  timeout 200 python bench.py --timesteps 300 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v (tc_small=$DIQT_TC_SMALL fusion_min=$DIQT_GN_FUSION_MIN) ms/iter %.4f' % d['ms_per_denoise_iteration'])"
done
