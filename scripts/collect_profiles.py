#!/usr/bin/env python
"""Copy / summarise the outputs of scripts/gpu_profile.sh (gpurun_out/, scratch) into profiles/
(tracked).  Run here after the gpurun call has merged its files back:

    python scripts/collect_profiles.py [round_prefix]        # default r01
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def main(pre="r02"):
    py = sys.executable
    copies = {"bench_default.json": f"{pre}_bench_b16.json", "bench_b8.json": f"{pre}_bench_b8.json",
              "bench_reference.json": f"{pre}_bench_reference_arm.json", "ops_brats_b16.json": f"{pre}_ops_brats_b16.json",
              "ops_512.json": f"{pre}_ops_512cubed.json", "texture_b4.json": f"{pre}_texture_config4_b4.json"}
    for src, dst in copies.items():
        s = os.path.join(G, src)
        if os.path.exists(s) and os.path.getsize(s) > 0:
            shutil.copyfile(s, os.path.join(P, dst))
            print("copied", src, "->", dst)
        else:
            print("MISSING", src)
    if os.path.exists(os.path.join(G, "launches.csv")):
        subprocess.check_call([py, os.path.join(ROOT, "scripts", "summarize_ncu.py"), "launches",
                               os.path.join(G, "launches.csv"), os.path.join(P, f"{pre}_launches_bench_b8.json")])
    if os.path.isdir(os.path.join(G, "ncu")):
        subprocess.check_call([py, os.path.join(ROOT, "scripts", "summarize_ncu.py"), "families", os.path.join(G, "ncu"),
                               os.path.join(P, f"{pre}_ncu_kernels.json")])
    xr = os.path.join(G, "ncu", "xcorr.ncu-rep")
    if os.path.exists(xr):
        subprocess.check_call([py, os.path.join(ROOT, "scripts", "summarize_ncu.py"), "full", xr,
                               os.path.join(P, f"{pre}_xcorr_ncu_full.json")], stdout=subprocess.DEVNULL)
        # DRAM traffic per voxel of the dominant kernel, read by bench.py for roofline.traffic
        k = json.load(open(os.path.join(P, f"{pre}_ncu_kernels.json")))["kernels"].get("xcorr")
        if k:
            nvox = 16 * 155 * 240 * 240
            t = {"_comment": "DRAM traffic per voxel of one launch, from ncu --set full (dram__bytes_read.sum + "
                             "dram__bytes_write.sum) / voxels of that launch; source capture named per entry",
                 "xcorr_kernel": {"dram_bytes_per_voxel": round((k["dram_read_mb"] + k["dram_write_mb"]) * 1e6 / nvox, 3),
                                  "read_mb": k["dram_read_mb"], "write_mb": k["dram_write_mb"], "voxels": nvox,
                                  "source": f"profiles/{pre}_ncu_kernels.json (16 BraTS volumes, rho=5, k=100)"}}
            json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
            print("traffic.json:", t["xcorr_kernel"])


if __name__ == "__main__":
    main(*(sys.argv[1:2]))
