"""profiles/r01_ncu_kernels.json: one row per kernel family from the ncu --set full captures of
scripts/gpu_ncu_all.sh (gpurun_out/ncu/*.ncu-rep)."""
import csv
import glob
import json
import os
import subprocess
import sys

KEYS = {
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


def main(src_dir, dst):
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
            if h in KEYS:
                k = KEYS[h]
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
    main(sys.argv[1], sys.argv[2])
