"""Print the headline numbers and the per-kernel table of one bench.py JSON line (a file)."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads([l for l in open(f) if l.startswith("{")][0])
    e = d.get("e2e") or {}
    print(f"{f}: value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {e.get('value', 0):.1f} ({e.get('ms_per_step', 0):.3f} ms)"
          f"  h2d {e.get('h2d_gbs_this_rank_all_ranks_copying')} GB/s  cpu {d.get('cpu_baseline') and d['cpu_baseline'].get('value')}")
    ks = d.get("kernels_serial") or {}
    if ks:
        print("  serial ms/step", ks["ms_per_step"])
        for k, v in sorted(ks["kernels"].items(), key=lambda kv: -kv[1]["launches_per_step"] * kv[1]["avg_launch_ms"]):
            t = v["launches_per_step"] * v["avg_launch_ms"]
            print(f"  {k:30s} n={v['launches_per_step']:4.1f} avg={v['avg_launch_ms']:.4f} tot={t:.3f} hbm_frac={v.get('frac_of_hbm_peak')}")
