#!/bin/bash
TAG="${1:-r2g}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_attn.py" -m gpu -q --tb=short --timeout=90 --timeout-method=thread --maxfail=6 2>&1 | tail -3
timeout 300 python tools/bench_sweep.py attn linattn > $OUT/sweep_attn_$TAG.jsonl 2>$OUT/sweep_attn_$TAG.err; cut -c1-330 $OUT/sweep_attn_$TAG.jsonl; tail -3 $OUT/sweep_attn_$TAG.err
timeout 400 python bench.py --steps 2 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("ms/iter", d["ms_per_denoise_iteration"], "value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d["roofline"], indent=1)[:1800])
print(d.get("volume"))
PY
tail -3 $OUT/bench_$TAG.err
timeout 600 ncu --set full --clock-control none -k regex:softmax_attn_tc2 -c 2 -f -o $OUT/prof_attn_$TAG python tools/bench_sweep.py attn > /dev/null 2>&1; echo "ncu attn rc=$?"
ncu -i $OUT/prof_attn_$TAG.ncu-rep --page raw --csv > $OUT/prof_attn_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_attn_${TAG}_raw.csv $OUT/ncu_attn_${TAG}.csv; cat $OUT/ncu_attn_${TAG}.csv
