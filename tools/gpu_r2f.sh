#!/bin/bash
TAG="${1:-r2f}"
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest "tests/test_gpu_attn.py::test_softmax_attention_tensor_core_kernel" "tests/test_gpu_attn.py::test_softmax_attention_tensor_core_13824_tokens" -m gpu -q --tb=short --timeout=90 --timeout-method=thread --maxfail=6 > $OUT/pytest_attn_$TAG.log 2>&1; rc=$?; echo "pytest(attn) rc=$rc"
grep -E "^(FAILED|ERROR)|passed|failed|Timeout" $OUT/pytest_attn_$TAG.log | tail -20
if [ $rc -ne 0 ]; then tail -40 $OUT/pytest_attn_$TAG.log; fi
timeout 300 python tools/bench_sweep.py attn > $OUT/sweep_attn_v2_$TAG.jsonl 2>$OUT/sweep_attn_$TAG.err; cut -c1-400 $OUT/sweep_attn_v2_$TAG.jsonl; tail -3 $OUT/sweep_attn_$TAG.err
DIQT_ATTN_TC_VERSION=1 timeout 300 python tools/bench_sweep.py attn > $OUT/sweep_attn_v1_$TAG.jsonl 2>/dev/null; cut -c1-300 $OUT/sweep_attn_v1_$TAG.jsonl
for v in default xf1 xf2; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== variant $v"
  timeout 200 python tools/bench_conv_gn.py 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print('%d^3 %d->%d fused %.2f plain %.2f two-kernel %.2f' % (d['side'], d['c_in'], d['c_out'], d['fused_us'], d['plain_conv_us'], d['apply_plus_conv_us']))
"
done
unset DIQT_LIB_PATH
timeout 300 python -m pytest "tests/test_gpu_kernels.py::test_conv_zmarch_cta_pair_equals_single_cta" tests/test_gpu_attn.py tests/test_gpu_unet.py -m gpu -q --tb=short --timeout=120 --timeout-method=thread --maxfail=6 2>&1 | tail -5
