#!/bin/bash
TAG="${1:-r2p}"
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_gpu_kernels.py::test_conv_tcgen05_split_k_small_volumes" -m gpu -q --tb=short --timeout=90 --timeout-method=thread --maxfail=8 2>&1 | tail -3
for cfg in "" "DIQT_DISABLE_SPLITK=1"; do
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
timeout 900 python tools/bench_sweep.py conv > $OUT/sweep_conv_$TAG.jsonl 2> $OUT/sweep_conv_$TAG.err; python - <<PY
import json
for ln in open("$OUT/sweep_conv_$TAG.jsonl"):
    d=json.loads(ln)
    if d.get("op")!="conv3x3x3" or d["side"]!=16: continue
    print("%2d^3 x %3d: ours %.4f ms (%s) others %s torch best %.4f speedup %.2f" % (d["side"], d["channels"], d["ours_ms"], d["ours_kernel"], {k[:12]: round(v,4) for k,v in d["ours_other_kernels_ms"].items()}, min(d["torch_ms"].values()), d["speedup"]))
PY
