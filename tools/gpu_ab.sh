#!/bin/bash
# A/B of the build variants on one GPU box: short bench (200 iterations) + the back-to-back reproducibility tests per variant.
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/ab.log
for v in default ${VARIANTS:-cg_early nc_early nc_late}; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  for rep in 1 2; do
    timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ab_$v.json 2> $OUT/ab_$v.err
    python - >> $OUT/ab.log <<PY
import json
try:
    d = json.loads(open("$OUT/ab_$v.json").read().strip().splitlines()[-1])
    print("$v rep$rep ms/iter %.4f  zm_us %.2f  e2e %.4f" % (d["ms_per_denoise_iteration"], d["roofline"]["ms_per_launch"] * 1e3, d["e2e"]["value"]))
except Exception as e:
    print("$v rep$rep bench failed", e, open("$OUT/ab_$v.err").read()[-600:])
PY
  done
  timeout 300 python -m pytest tests/test_gpu_volume.py tests/test_gpu_sampler.py -q -m gpu --tb=line 2>&1 | tail -2 >> $OUT/ab.log
done
cat $OUT/ab.log
