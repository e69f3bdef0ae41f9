"""Turn the ncu outputs of scripts/gpu_profile.sh into the small text/JSON summaries committed
under profiles/ (the .ncu-rep files themselves stay in gpurun_out/, which is scratch).

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_bench_b8.json
    python scripts/summarize_ncu.py full gpurun_out/xcorr_full.ncu-rep profiles/r01_xcorr_ncu_full.json
    python scripts/summarize_ncu.py families gpurun_out/ncu profiles/r02_ncu_kernels.json   # one row per *.ncu-rep
"""
import csv
import glob
import json
import os
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared_atom.sum",
]


def launches(src, dst):
    per = defaultdict(lambda: [0, 0.0])
    n = 0
    with open(src, newline="") as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * 1e3 if r[iu] in ("ms", "msecond") else v  # -> us
        per[name][0] += 1
        per[name][1] += v
        n += 1
    tot = sum(v[1] for v in per.values())
    out = {"source": src, "launches": n, "total_us": round(tot, 1),
           "note": "ncu --metrics gpu__time_duration.sum --clock-control none: per-launch times are cold-cache and "
                   "serialised; compare SHARES with bench.py's CUDA-event shares, not absolutes",
           "kernels": {k: {"launches": c, "total_us": round(t, 1), "share": round(t / tot, 4)}
                       for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in list(out["kernels"].items())[:12]:
        print(f"{k:32s} {v['launches']:5d} {v['total_us']:12.1f} us {100 * v['share']:6.2f} %")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = {"source": src, "captures": []}
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                rec[h] = {"value": vals[i], "unit": units[i]}
        out["captures"].append(rec)
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["captures"][0], indent=1)[:3000])


FAMILY_KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__inst_executed.sum": "warp_instructions",
}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
    return float(v) * m.get(unit, 1)


def families(src_dir, dst):
    out = {"note": "ncu --set full --clock-control none, one launch per kernel family; dram_gbs = (read+write)/time; "
                   "peak for the fraction = MEASURED_PEAKS.json hbm_gbs 6551.4", "kernels": {}}
    for rep in sorted(glob.glob(os.path.join(src_dir, "*.ncu-rep"))):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        rec = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
        for i, h in enumerate(hdr):
            if h in FAMILY_KEYS:
                k = FAMILY_KEYS[h]
                if k == "time":
                    rec["time_us"] = round(to_us(vals[i], units[i]), 2)
                elif k in ("dram_read", "dram_write"):
                    rec[k + "_mb"] = round(to_bytes(vals[i], units[i]) / 1e6, 2)
                else:
                    rec[k] = round(float(vals[i].replace(",", "")), 2)
        if "time_us" in rec:
            gbs = (rec.get("dram_read_mb", 0) + rec.get("dram_write_mb", 0)) * 1e6 / (rec["time_us"] * 1e-6) / 1e9
            rec["dram_gbs"] = round(gbs, 1)
            rec["dram_frac_of_measured_peak"] = round(gbs / 6551.4, 4)
        out["kernels"][os.path.basename(rep)[:-8]] = rec
    json.dump(out, open(dst, "w"), indent=1)
    for k, r in out["kernels"].items():
        print(f'{k:22s} {r["kernel"][:28]:28s} {r.get("time_us", 0):9.1f} us dram {r.get("dram_gbs", 0):7.1f} GB/s ({100 * r.get("dram_frac_of_measured_peak", 0):5.1f}%) '
              f'L2 {r.get("l2_pct", 0):5.1f}% L1 {r.get("l1_pct", 0):5.1f}% issue {r.get("issue_active_pct", 0):5.1f}% occ {r.get("achieved_occupancy_pct", 0):5.1f}% regs {int(r.get("registers", 0))}')



if __name__ == "__main__":
    {"launches": launches, "full": full, "families": families}[sys.argv[1]](sys.argv[2], sys.argv[3])
