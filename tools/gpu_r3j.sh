#!/bin/bash
# bench.py with the training-step record; ncu --set full of the new tcgen05 kernels (weight gradient, linear attention)
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python bench.py --timesteps 200 --steps 2 --warmup 3 --no-cpu-baseline --no-volume > $OUT/bench_r3j.json 2> $OUT/bench_r3j.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r3j.json").read().strip().splitlines()[-1])
print("ms/iter %.4f whole %.4f" % (d["ms_per_denoise_iteration"], d["roofline"]["whole_step"]["frac"]))
print(json.dumps(d.get("train_step")))
PY
tail -3 $OUT/bench_r3j.err
BENCH_TRAIN_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 10 -c 4 -f -o $OUT/prof_wg_r3j python tools/bench_train.py 64 1 > $OUT/ncu_wg_r3j.log 2>&1; echo "ncu wgrad rc=$?"
ncu -i $OUT/prof_wg_r3j.ncu-rep --page raw --csv > $OUT/prof_wg_r3j_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:linattn_ -s 60 -c 8 -f -o $OUT/prof_la_r3j python tools/bench_sweep.py linattn > $OUT/ncu_la_r3j.log 2>&1; echo "ncu linattn rc=$?"
ncu -i $OUT/prof_la_r3j.ncu-rep --page raw --csv > $OUT/prof_la_r3j_raw.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
python tools/ncu_summary.py $OUT/prof_wg_r3j_raw.csv $OUT/r3j_ncu_wgrad_tc.csv
python tools/ncu_summary.py $OUT/prof_la_r3j_raw.csv $OUT/r3j_ncu_linattn.csv
cat $OUT/r3j_ncu_wgrad_tc.csv | cut -c1-400; cat $OUT/r3j_ncu_linattn.csv | cut -c1-400
du -sh $OUT
