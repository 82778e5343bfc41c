#!/bin/bash
TAG="${1:-r2c}"
OUT=gpurun_out
mkdir -p $OUT
timeout 120 ./build/umma_probe 4 > $OUT/umma_probe_e4_$TAG.log 2>&1; echo "probe rc=$?"; cat $OUT/umma_probe_e4_$TAG.log
timeout 300 python -m pytest "tests/test_gpu_kernels.py::test_conv_zmarch_fused_groupnorm_film_mish" "tests/test_gpu_kernels.py::test_conv_zmarch_fused_affine_mish_any_batch" "tests/test_gpu_kernels.py::test_conv_zmarch_bf16" "tests/test_gpu_kernels.py::test_se_scale_residual" tests/test_gpu_unet.py tests/test_gpu_config_size.py -m gpu -q -x --tb=short --timeout=200 --timeout-method=thread > $OUT/pytest_k_$TAG.log 2>&1; rc=$?; echo "pytest(kernels) rc=$rc"
grep -E "^(FAILED|ERROR)|passed|failed|Error" $OUT/pytest_k_$TAG.log | tail -20
if [ $rc -ne 0 ]; then tail -60 $OUT/pytest_k_$TAG.log; fi
for v in default mishv2; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== variant $v"
  timeout 200 python tools/bench_conv_gn.py > $OUT/conv_gn_${v}_$TAG.jsonl 2>$OUT/conv_gn_${v}_$TAG.err; cat $OUT/conv_gn_${v}_$TAG.jsonl; tail -3 $OUT/conv_gn_${v}_$TAG.err
  for rep in 1 2; do
    timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline > $OUT/ab_${v}_$TAG.json 2> $OUT/ab_${v}_$TAG.err
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/ab_${v}_$TAG.json").read().strip().splitlines()[-1])
    print("$v rep$rep ms/iter %.4f  zm_us %.2f  e2e %.4f" % (d["ms_per_denoise_iteration"], d["roofline"]["ms_per_launch"] * 1e3, d["e2e"]["value"]))
except Exception as e:
    print("$v rep$rep bench failed", e, open("$OUT/ab_${v}_$TAG.err").read()[-600:])
PY
  done
done
unset DIQT_LIB_PATH
DIQT_DISABLE_GN_FUSION=1 timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unfused ms/iter %.4f' % d['ms_per_denoise_iteration'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_zm_kernel -s 40 -c 2 -f -o $OUT/prof_zm_$TAG python bench.py --timesteps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline > $OUT/ncu_zm_$TAG.log 2>&1; echo "ncu zm rc=$?"
ncu -i $OUT/prof_zm_$TAG.ncu-rep --page raw --csv > $OUT/prof_zm_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_zm_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_zm_${TAG}_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_zm_${TAG}_raw.csv $OUT/ncu_zm_${TAG}.csv; cat $OUT/ncu_zm_${TAG}.csv
