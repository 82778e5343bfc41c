#!/bin/bash
TAG="${1:-r2l}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_kernels.py::test_init_conv_fused_tensor_core_kernel" "tests/test_gpu_kernels.py::test_init_conv_im2col_tensor_core_path" tests/test_gpu_unet.py tests/test_gpu_sampler.py -m gpu -q --tb=short --timeout=120 --timeout-method=thread --maxfail=6 > $OUT/pytest_init_$TAG.log 2>&1; echo "pytest(init) rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_init_$TAG.log | tail; 
for cfg in "" "DIQT_DISABLE_INIT_FUSED=1"; do
  for rep in 1 2; do
    env $cfg timeout 200 python bench.py --timesteps 200 --steps 2 --warmup 1 --no-cpu-baseline --no-volume --no-torch-gpu-baseline 2>$OUT/ab_$TAG.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$cfg] rep$rep ms/iter %.4f' % (d['ms_per_denoise_iteration']))
except Exception as e:
    print('[$cfg] failed', e, open('$OUT/ab_$TAG.err').read()[-800:])
"
  done
done
export DIQT_LIB_PATH=$PWD/build/variants/attn_loads1.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_attn_tc2_kernel -s 1 -c 1 -f -o $OUT/prof_attn_$TAG python tools/attn_once.py > $OUT/ncu_attn_$TAG.log 2>&1; echo "ncu attn rc=$?"
unset DIQT_LIB_PATH
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page raw --csv > $OUT/prof_attn_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_attn_${TAG}_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_attn_${TAG}_raw.csv $OUT/ncu_attn_${TAG}.csv; cut -c1-700 $OUT/ncu_attn_${TAG}.csv
grep -E "tmem|tensor_op|xu|inst_executed_pipe" $OUT/prof_attn_${TAG}_raw.csv | head -5
