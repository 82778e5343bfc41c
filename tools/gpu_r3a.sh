#!/bin/bash
# attention: single-pass softmax kernel with V as an MN-major operand (no transposed copy), per-kernel times of the linear attention
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_attn.py tests/test_gpu_unet.py -m gpu -q --maxfail=10 --tb=short --timeout=150 --timeout-method=thread -k "linear" > $OUT/pytest_r3a.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r3a.log | tail -15
grep -E "^E " $OUT/pytest_r3a.log | head -20
timeout 300 python tools/bench_sweep.py linattn > $OUT/sweep_attn_r3a.jsonl 2> $OUT/sweep_attn_r3a.err; echo "sweep rc=$?"
cut -c1-420 $OUT/sweep_attn_r3a.jsonl; tail -5 $OUT/sweep_attn_r3a.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:linattn --csv --log-file $OUT/launches_linattn_r3a.csv python tools/bench_sweep.py linattn > $OUT/ncu_linattn_r3a.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
agg = collections.OrderedDict()
for r in csv.reader(open("gpurun_out/launches_linattn_r3a.csv")):
    if len(r) < 15 or not r[0].isdigit():
        continue
    unit, val = r[13], float(r[14].replace(",", ""))
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * (1000.0 if unit in ("ms", "msecond") else 1.0)
    agg.setdefault((r[4].split("(")[0], r[8]), []).append(us)
for k, v in agg.items():
    v.sort()
    print(k, "n=%d median %.2f us" % (len(v), v[len(v) // 2]))
PY
