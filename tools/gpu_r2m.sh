#!/bin/bash
TAG="${1:-r2n}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_attn.py" -m gpu -q --tb=short --timeout=90 --timeout-method=thread --maxfail=6 -k "tensor_core" 2>&1 | tail -3
for v in default attn_loads1; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== $v"
  for rep in 1 2; do timeout 300 python tools/bench_sweep.py attn 2>/dev/null | grep softmax | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print('  tokens %5d ours %.4f ms (%.3f of burst) sdpa %.4f ms err %.1e' % (d['tokens'], d['ours_ms'], d['ours_frac_burst'], d['torch_ms']['sdpa_bf16'], d['max_rel_vs_sdpa_fp32']))
"; done
done
unset DIQT_LIB_PATH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_attn_tc2_kernel -s 1 -c 1 -f -o $OUT/prof_attn_$TAG python tools/attn_once.py > $OUT/ncu_attn_$TAG.log 2>&1; echo "ncu attn rc=$?"
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page raw --csv > $OUT/prof_attn_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_attn_${TAG}_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_attn_${TAG}_raw.csv $OUT/ncu_attn_${TAG}.csv; cut -c1-700 $OUT/ncu_attn_${TAG}.csv
