#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --maxfail=30 --tb=short --timeout=300 --timeout-method=thread > $OUT/pytest_r3h.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_r3h.log | tail -10
timeout 600 python tools/bench_train.py 64 1 > $OUT/bench_train_r3h.jsonl 2> $OUT/bench_train_r3h.err; echo rc=$?; cat $OUT/bench_train_r3h.jsonl; tail -3 $OUT/bench_train_r3h.err
BENCH_TRAIN_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 9000 --csv --log-file $OUT/launches_train_r3h.csv python tools/bench_train.py 64 1 > $OUT/ncu_train_r3h.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = []
for r in csv.reader(open("gpurun_out/launches_train_r3h.csv")):
    if len(r) < 15 or not r[0].isdigit():
        continue
    unit, val = r[13], float(r[14].replace(",", ""))
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * (1000.0 if unit in ("ms", "msecond") else 1.0)
    rows.append((r[4].split("(")[0][:70], us))
step = rows   # BENCH_TRAIN_ONLY runs exactly one step
agg = collections.defaultdict(lambda: [0, 0.0])
for k, us in step:
    agg[k][0] += 1; agg[k][1] += us
tot = sum(v[1] for v in agg.values())
print("one step: %d launches, %.2f ms of kernel time" % (len(step), tot / 1000))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-72s n=%4d  %9.1f us  %5.1f %%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
