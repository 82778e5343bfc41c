"""Condense an `ncu -i x.ncu-rep --page raw --csv` export into the handful of columns the profiles/ summaries quote, one row per
captured launch, and (with --traffic) refresh profiles/ncu_traffic.json for bench.py's `roofline.traffic`, stamped with the hash of the
kernel sources so that bench.py can tell a stale capture from a current one.

    python tools/ncu_summary.py gpurun_out/prof_zm_r2a_raw.csv profiles/r2a_ncu_conv_zm.csv [--traffic "conv_zm_kernel 3x3x3 64->64 @64^3"]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum"]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}   # -> MB and us


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    ki = hdr.index("Kernel Name")
    cols = [c for c in COLS if c in hdr]
    out = [["Kernel Name"] + cols]
    for r in data:
        if len(r) <= ki:
            continue
        line = [r[ki]]
        for c in cols:
            i = hdr.index(c)
            try:
                v = float(r[i].replace(",", ""))
                line.append("%.6f" % (v * SCALE.get(units[i], 1.0)))
            except ValueError:
                line.append(r[i])
        out.append(line)
    with open(dst, "w", newline="") as f:
        csv.writer(f).writerows(out)
    print(f"{len(out) - 1} launches -> {dst} (bytes in MB, times in us)")
    if "--traffic" in sys.argv:
        import bench
        label = sys.argv[sys.argv.index("--traffic") + 1]
        kname = label.split()[0]
        sel = [r for r in out[1:] if kname in r[0]]
        ri, wi, ti = out[0].index("dram__bytes_read.sum"), out[0].index("dram__bytes_write.sum"), out[0].index("gpu__time_duration.sum")
        rd = sum(float(r[ri]) for r in sel) / len(sel)
        wr = sum(float(r[wi]) for r in sel) / len(sel)
        path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        t = json.load(open(path)) if os.path.isfile(path) else {}
        t[label] = dict(dram_read_mb=rd, dram_write_mb=wr, bytes_per_launch=(rd + wr) * 1e6, gpu_time_us_under_ncu=sum(float(r[ti]) for r in sel) / len(sel),
                        launches=len(sel), source=os.path.relpath(dst, ROOT) + " (ncu --set full --clock-control none)", source_sha=bench.kernel_source_sha())
        json.dump(t, open(path, "w"), indent=1)
        print(f"{path}: {label}: {(rd + wr):.2f} MB per launch, sha {t[label]['source_sha']}")


if __name__ == "__main__":
    main()
