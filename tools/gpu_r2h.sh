#!/bin/bash
TAG="${1:-r2h}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_attn.py" -m gpu -q --tb=short --timeout=90 --timeout-method=thread --maxfail=6 2>&1 | tail -3
timeout 300 python tools/bench_sweep.py attn linattn > $OUT/sweep_attn_$TAG.jsonl 2>$OUT/sweep_attn_$TAG.err; cut -c1-330 $OUT/sweep_attn_$TAG.jsonl; tail -3 $OUT/sweep_attn_$TAG.err
for v in default xf3 w3 ring3w2; do
  if [ $v = default ]; then unset DIQT_LIB_PATH; else export DIQT_LIB_PATH=$PWD/build/variants/$v.so; fi
  echo "== variant $v"
  timeout 200 python tools/bench_conv_gn.py 2>&1 | python -c "
import json,sys
for ln in sys.stdin:
    try:
        d=json.loads(ln); print('%d^3 %d->%d fused %.2f plain %.2f two-kernel %.2f' % (d['side'], d['c_in'], d['c_out'], d['fused_us'], d['plain_conv_us'], d['apply_plus_conv_us']))
    except Exception: print(ln.strip()[:200])
"
done
unset DIQT_LIB_PATH
timeout 600 ncu --set full --clock-control none -k regex:softmax_attn_tc2 -c 4 -f -o $OUT/prof_attn_$TAG python tools/bench_sweep.py attn > /dev/null 2>&1; echo "ncu attn rc=$?"
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page raw --csv > $OUT/prof_attn_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_attn_${TAG}_raw.csv $OUT/ncu_attn_${TAG}.csv; cut -c1-400 $OUT/ncu_attn_${TAG}.csv
