#!/usr/bin/env python
"""bench.py -- BraTS-shape GTV-spec throughput (volumes/s), the metric of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm on the host cores

A "step" is one pass of the whole GTV segmentation spec (Greta.pdf p.48) over one batch
of synthetic BraTS-shaped volumes (240x240x155 f32) per
GPU, driven through the ImgQL driver -> ctypes -> libvoxcuda.so (the C ABI).
  value : volumes/s with the batch already resident in HBM (weak scaling: every rank
          owns its own batch; the path shards by volume, no data-path collective).
  e2e   : the same with the batch in pinned HOST memory: every step's int16 voxels (what a
          NIfTI file holds) go host -> device and its GTV masks come back as bits, inside the
          timed region; one host thread uploads ahead of the analysis threads (--prefetch).
  roofline     : dominant kernel by device time, algorithmic bytes / CUDA-event time
                 against the measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline : the CPU oracle (a port: the reference implementation is not mounted)
                 timed on this box's host cores on one full-size volume.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "brats_gtv_spec_throughput"
UNIT = "volumes/s"
BRATS = (155, 240, 240)
NVOX = BRATS[0] * BRATS[1] * BRATS[2]
WORKLOAD = ("BraTS-shape 240x240x155 GTV segmentation spec (Greta.pdf p.48: threshold, touch/grow via through, "
            "filter via distlt/distgeq, percentiles, crossCorrelation rho=5 k=100)")
# algorithmic HBM bytes per voxel per launch (SURVEY.md section 8d / DESIGN.md kernel table)
ALG_BYTES = {
    # a boolean voxel is one BIT on the device (1/8 byte), an f32 4 bytes, a 16-bit code 2
    "xcorr_kernel": 5.0,             # u8 bins in, f32 coefficient out (13 B/voxel for the whole operator)
    "pointwise_program_kernel": None,  # depends on the program (4.125 for a threshold, 8/12 for arithmetic)
    "boolean_program_kernel": None,    # (leaves + 1) / 8
    "morph_bits_kernel": 0.25, "ball_or_kernel": 0.25,
    "pack_bits_kernel": 0.25,        # bits in, packed rows out (+ emax pre-dilated levels, emax/8 more)
    # union-find kernels as launched by `through`: bits in/out, node ids are 4 bytes per 32-voxel word
    "ccl_nodes_kernel": 0.375, "ccl_merge_kernel": None, "ccl_flatten_kernel": None,
    "ccl_select_kernel": 0.25, "ccl_seed_kernel": 0.25,
    "code_compare_kernel": 2.25,     # 16-bit codes (+ mask bits) in, result bits out
    "xc_bin_kernel": 3.0, "xc_region_hist_kernel": 0.125, "pc_compact_kernel": 4.125,
    # the sort kernels move 8-byte pairs of the MASKED voxels only: no per-voxel figure
    "pc_rank_kernel": None, "rs_hist_kernel": None, "rs_scatter_kernel": None, "rs_rowscan_kernel": None,
    # percentiles of a quantised (16-bit origin) image by counting: codes 2 + mask bits in (f32 out when the array is made)
    "pq_hist_kernel": 2.125, "pq_apply_kernel": 6.125,
    "reduce_f32_kernel": 4.0, "reduce_volume_kernel": 0.125,
    "border_kernel": 0.125, "edt_x_kernel": 4.125, "edt_axis_kernel_y": 8.0, "edt_axis_kernel_z": 8.0,
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling guide)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_batch(batch: int, rank: int, shape=BRATS) -> np.ndarray:
    """`batch` distinct BraTS-like volumes as a file holds them: int16 voxels (scl_slope 1/4095).
    Up to 4 are generated (seeds per SURVEY 8d config 5), the rest are axis flips of them (distinct
    data, identical statistics; generating a full-size volume costs ~1 s of host time)."""
    from voxlogica_b200.synth import brats_like_int16
    uniq = min(batch, 4)
    base = [brats_like_int16(seed=rank * 64 + i, shape=shape)[0] for i in range(uniq)]
    vols = []
    for i in range(batch):
        v = base[i % uniq]
        f = i // uniq
        if f & 1:
            v = v[:, :, ::-1]
        if f & 2:
            v = v[:, ::-1, :]
        if f & 4:
            v = v[::-1, :, :]
        vols.append(np.ascontiguousarray(v))
    return np.stack(vols)


# --------------------------------------------------------------------------------- CPU arm
def cpu_gtv_sample(n_threads: int, shape, seed=1234):
    """One GTV spec pass of the CPU oracle on one volume of `shape`.  Returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["OMP_NUM_THREADS"] = str(n_threads)
    import oracle as orc
    from gtv_ref import gtv_oracle
    from voxlogica_b200.synth import brats_like_int16, int16_to_f32
    raw, slope = brats_like_int16(seed=seed, shape=shape)
    vol = int16_to_f32(raw, slope)      # the f32 image the GPU arm forms on the device from the same int16 voxels
    orc.lib()
    t0 = time.perf_counter()
    gtv_oracle(orc, vol)
    return time.perf_counter() - t0


CPU_SAMPLE_SHAPE = BRATS   # the CPU arm runs the benchmark's own volume size, never a scaled-down one


def _sample_text(nsteps):
    return (f"{nsteps} full synthetic BraTS-shape volume(s) of {BRATS[2]}x{BRATS[1]}x{BRATS[0]} voxels, each through "
            "the whole GTV spec with the OpenMP CPU oracle built -O3 -march=native (a restatement: the "
            "SimpleITK/F# reference is not mounted); volumes/s = volumes / wall seconds, no scaling")


def cpu_baseline(n_threads: int):
    cpu_gtv_sample(n_threads, (20, 32, 32))   # build/load the oracle, warm the thread pool
    sec = cpu_gtv_sample(n_threads, CPU_SAMPLE_SHAPE)
    return {"value": 1.0 / sec, "unit": UNIT, "cores": n_threads, "kind": "port", "seconds": sec,
            "sample": _sample_text(1)}


REF_MAX_TIMED_STEPS = 4   # a full-size CPU step takes several seconds: the arm times at most this many


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    cpu_gtv_sample(n_threads, (20, 32, 32))  # build/load the oracle, warm caches
    nwarm = 1 if args.warmup > 0 else 0
    ntimed = max(1, min(args.steps, REF_MAX_TIMED_STEPS))
    times = []
    for i in range(nwarm + ntimed):
        sec = cpu_gtv_sample(n_threads, CPU_SAMPLE_SHAPE, seed=1234 + i)
        if i >= nwarm:
            times.append(sec)
    total = sum(times)
    val = len(times) / total
    cb = {"value": val, "unit": UNIT, "cores": n_threads, "kind": "port", "sample": _sample_text(len(times))}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": nwarm, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/int64",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": WORKLOAD, "volumes_per_step": 1, "voxels_per_volume": NVOX, "adjacency": 26,
                   "steps_requested": args.steps, "warmup_requested": args.warmup,
                   "note": "CPU arm: one full-size volume per step on all host threads; the SimpleITK/F# "
                           "reference is not mounted and not runnable, so this is the repo's CPU oracle (a port); "
                           f"timed steps are capped at {REF_MAX_TIMED_STEPS} because one step takes seconds"},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(local: int):
    """Run this rank (and allocate its pinned staging buffers) on the CPUs next to its GPU, so that the
    host<->device copies of the e2e path do not cross the socket interconnect.  Source of the
    affinity: sysfs numa_node of the GPU's PCI function; where the (virtualised) host reports -1
    there, the "CPU Affinity" column of `nvidia-smi topo -m`.  Returns a description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpus, how = set(), ""
        try:
            node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        except OSError:
            node = -1
        if node >= 0:
            cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
            how = f"sysfs node {node}"
        else:
            out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            import re
            out = re.sub(r"\x1b\[[0-9;]*m", "", out)
            lines = [ln for ln in out.splitlines() if ln.strip()]
            hdr = [h.strip() for h in lines[0].split("\t")]
            col = hdr.index("CPU Affinity")
            for ln in lines[1:]:
                f = [x.strip() for x in ln.split("\t")]
                if f and f[0] == f"GPU{local}" and len(f) > col:
                    cpus = _cpulist(f[col])
                    how = "nvidia-smi topo CPU affinity"
                    break
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"{dev}: no usable affinity ({how or 'none reported'})"
        os.sched_setaffinity(0, cpus)
        return f"{dev}: {how}, {len(cpus)} cpus ({min(cpus)}-{max(cpus)})"
    except Exception as e:  # pragma: no cover
        return f"unavailable ({type(e).__name__}: {e})"


def xcorr_note():
    """what binds the dominant kernel, with the figures of the committed ncu capture"""
    txt = "xcorr_kernel is bound by the shared-memory LSU pipe, not by HBM (SURVEY.md 8d); the fraction against HBM is " \
          "reported as the contract requires"
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r02_xcorr_ncu_full.json")))["captures"][0]
        f = lambda k: float(cap[k]["value"])
        txt += (" (ncu, profiles/r02_xcorr_ncu_full.json: %.1f%% of peak shared-memory wavefronts, %.0f%% issue active, "
                "DRAM %.1f%%)" % (f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                                  f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                  f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
    except Exception:
        pass
    return txt


# --------------------------------------------------------------------------------- slab leg (N > 1)
def slab_leg(ctx, dist, rank, world, local):
    """BASELINE config 5b: ONE oversized volume slab-decomposed along z over the N ranks, after the
    timed region.  (1) an on-box check: near26 / through26 / distlt(5) on a 640^3 volume in slabs
    against the same operators on the whole volume (every rank runs the whole volume on its own GPU
    and compares its slab, bit for bit); (2) per-operator times on the large volume (1024^3 at N=2,
    1536^3 at N=4, 2048^3 at N=8) with the bytes each exchange moves over NVLink and the phases
    of `through`, so that the exchange that dominates it is named."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from slab_gpu import blob_slab
    from voxlogica_b200.slab import CudaEngine, Slab, SlabComm, SlabOps, slab_bounds
    dev = f"cuda:{local}"
    sp = (1.0, 1.0, 1.0)
    out = {}

    def make(n, ops):
        z0, z1 = slab_bounds(n, world)[rank]
        b_t = blob_slab((n, n, n), z0, z1, dev)
        seed_t = torch.zeros_like(b_t)
        if rank == 0:
            seed_t[0] = b_t[0]                     # everything connected to the first plane
        B = Slab(ctx.from_tensor(b_t[None], spacing=sp), z0, z1, n, sp)
        A = Slab(ctx.from_tensor(seed_t[None], spacing=sp), z0, z1, n, sp)
        return b_t, A, B, z0, z1

    # ---- (1) check against the whole volume
    n = 640
    ops = SlabOps(CudaEngine(ctx), SlabComm())
    ops.enable_peer_halo(n, n)
    b_t, A, B, z0, z1 = make(n, ops)
    parts = [torch.empty((b - a, n, n), dtype=torch.uint8, device=dev) for a, b in slab_bounds(n, world)]
    dist.all_gather(parts, b_t)
    whole = torch.cat(parts)
    del parts
    W = ctx.from_tensor(whole[None], spacing=sp)
    wseed = torch.zeros_like(whole)
    wseed[0] = whole[0]
    WS = ctx.from_tensor(wseed[None], spacing=sp)
    bad = []
    for name, fs, fw in (("near26", lambda: ops.near(B, 26), lambda: ctx.near(W, 26)),
                         ("through26", lambda: ops.through(A, B, 26), lambda: ctx.through(WS, W, 26)),
                         ("distlt5", lambda: ops.distlt(5.0, B), lambda: ctx.distlt(5.0, W))):
        r_slab, r_whole = fs(), fw()             # keep the images alive while their views are compared
        got = ctx.as_tensor(r_slab.data)[0]
        want = ctx.as_tensor(r_whole)[0][z0:z1]
        ctx.sync()
        if not torch.equal(got, want):
            bad.append(name)
        del got, want, r_slab, r_whole
    flags = [None] * world
    dist.all_gather_object(flags, bad)
    out["check"] = "ok" if not any(flags) else f"FAILED {flags}"
    out["check_volume"] = f"{n}^3 in {world} z-slabs vs the whole volume on one GPU: near26, through26, distlt(5), bit for bit"
    del W, WS, whole, wseed, A, B, b_t, ops
    torch.cuda.empty_cache()

    # ---- (2) the large volume
    n = {2: 1024, 4: 1536, 8: 2048}.get(world, 1024)
    ops = SlabOps(CudaEngine(ctx), SlabComm())
    ops.enable_peer_halo(n, n)
    b_t, A, B, z0, z1 = make(n, ops)
    del b_t
    torch.cuda.empty_cache()

    def timed(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            ctx.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            ctx.sync(); torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t[0]))
        return float(np.median(ts))

    plane = n * n
    res = {}
    for name, fn, nvl in (
            ("near26", lambda: ops.near(B, 26),
             {"peer_reads_per_interior_rank": 2 * plane, "what": "two u8 boundary planes read from the neighbours' IPC mailboxes"}),
            ("through26", lambda: ops.through(A, B, 26),
             {"p2p_send_per_interior_rank": 2 * 8 * plane, "what": "boundary label planes (int64 global ids) to both neighbours, "
              "then two all-gathers of the cross-slab pairs (16 B each) and reached ids (8 B each)"}),
            ("distlt5", lambda: ops.distlt(5.0, B),
             {"p2p_send_per_interior_rank": 2 * 5 * plane, "what": "five u8 halo planes to each neighbour"})):
        sec = timed(fn)
        res[name] = {"ms": round(1e3 * sec, 3), "gvox_per_s": round(n ** 3 / sec / 1e9, 2), "nvlink_bytes": nvl}
    # phases of `through` (device-synchronised after every phase, so their sum exceeds the timed call)
    ops.trace = {}
    ops.through(A, B, 26)
    tr, ops.trace = ops.trace, None
    phases = {k: round(1e3 * v, 3) for k, v in tr.items() if isinstance(v, float)}
    tmax = torch.tensor([phases.get(k, 0.0) for k in sorted(phases)], device=dev, dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    phases = {k: round(float(v), 3) for k, v in zip(sorted(phases), tmax.tolist())}
    dom = max(phases, key=phases.get) if phases else None
    res["through26"]["phases_ms_max_over_ranks"] = phases
    res["through26"]["dominant_phase"] = dom
    res["through26"]["pairs_gathered"] = tr.get("pairs_gathered")
    res["through26"]["reached_gathered"] = tr.get("reached_gathered")
    reached = ops.volume(ops.through(A, B, 26))
    out.update({"volume": f"{n}^3 u8 (blobs), {world} z-slabs of {n // world} planes", "voxels": n ** 3, "ops": res,
                "foreground_voxels": ops.volume(B), "reached_from_first_plane": reached,
                "timing": "host clock around device-synchronised calls, max over ranks, median of 3"})
    return out


# --------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from voxlogica_b200 import Context
    from voxlogica_b200.imgql import run_gtv

    numa = bind_to_gpu_numa(local)
    ctx = Context(local)
    if args.independent_uploads:
        from voxlogica_b200 import _abi as A
        ctx.set_option(A.VX_OPT_ORDERED_UPLOADS, 0)
    B = args.batch
    shape4 = (B,) + BRATS
    from voxlogica_b200.synth import INT16_SLOPE
    host = make_batch(B, rank)                      # int16 voxels, as a NIfTI file holds them
    in_bytes, out_bytes = host.nbytes, (B * NVOX + 7) // 8

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        ctx.sync()

    def step_resident(flair):
        out = run_gtv(ctx, flair)
        return out["gtv"], out["_printed"]["gtvVolume"]   # volume(gtv) syncs: the step's result

    # ---- device-resident throughput.  The batch is uploaded once; `args.threads` host threads
    # (= CUDA streams of the library, as the F# scheduler's thread-pool threads would be) take
    # the K steps in turn, each step one full pass of the spec over the resident batch.  Steps of
    # different threads overlap on the device, which fills the bubbles around the host-visible
    # scalars (min/max of the input, volume(gtv)) that every step has to wait for.
    import threading
    flair = ctx.upload_scaled(host, INT16_SLOPE)    # f32 in HBM = fl(raw * slope), formed on the device
    ctx.sync()
    nres = max(1, args.threads)
    last = [None] * nres

    class StepQueue:
        """the K steps of a run, handed to whichever host thread is free next (a fixed split leaves one
        thread with an extra step when K is not a multiple of the thread count)"""
        def __init__(self, n):
            self.n, self.next, self.mu = n, 0, threading.Lock()

        def take(self):
            with self.mu:
                if self.next >= self.n:
                    return False
                self.next += 1
                return True

    def resident_worker(t, queue, errors):
        try:
            while queue.take():
                _, last[t] = step_resident(flair)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    def run_resident(nsteps):
        errors = []
        queue = StepQueue(nsteps)
        ths = [threading.Thread(target=resident_worker, args=(t, queue, errors)) for t in range(nres)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        if errors:
            raise errors[0]
        ctx.sync()

    run_resident(max(args.warmup, 2) * nres)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # per-kernel CUDA events are OFF inside the timed region; the kernel table comes from an
    # untimed single-stream pass afterwards (--prof-timed puts them back, for diagnosis only)
    ctx.prof_reset()
    ctx.prof_enable(bool(args.prof_timed))
    l0 = ctx.launch_count()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)                      # main stream, every stream idle: device-side start mark
    run_resident(args.steps)            # returns after all streams have drained
    ctx.record(e1)                      # device-side end mark
    ms = ctx.elapsed_ms(e0, e1)
    vols = next((v for v in last if v is not None), None)
    barrier()
    ctx.prof_enable(False)
    launches = ctx.launch_count() - l0
    prof = ctx.prof_report() if args.prof_timed else {}
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host -> HBM -> spec -> pinned host.
    # The public API is re-entrant with one CUDA stream per calling thread (the F# scheduler
    # calls from thread-pool threads), so the e2e path uses the same `args.threads` host threads, which
    # take whole steps in turn: thread t runs steps t, t+T, ... each one upload of the full batch
    # -> GTV spec -> download of the masks, with its own pinned buffers.  The copies of one
    # thread's step overlap the kernels of the other's.  Every step moves the whole batch both ways.
    import ctypes as C
    import threading
    del flair
    nthr = max(1, args.threads)
    pins_in = [ctx.host_alloc(host.nbytes, write_combined=args.wc) for _ in range(nthr + max(0, args.prefetch))]
    pins_out = [ctx.host_alloc(out_bytes) for _ in range(nthr)]
    for p_in in pins_in:
        C.memmove(p_in, host.ctypes.data, host.nbytes)

    def e2e_worker(t, queue, errors):
        try:
            while queue.take():
                f = ctx.upload_scaled_ptr(pins_in[t], np.int16, shape4, INT16_SLOPE)
                out = run_gtv(ctx, f)
                ctx.download_bits_ptr(out["gtv"], pins_out[t], out_bytes)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    # With --prefetch D (default 2) the uploads come from their own host thread (its own library stream), as a
    # harness that reads the next file while the current one is being analysed would issue them: it keeps at
    # most D uploaded batches waiting, the `nthr` analysis threads take them in turn (the library orders a
    # consumer's stream behind the upload through the event in the handle).  Every step still copies its whole
    # batch host -> device and its masks device -> host inside the timed region; what changes is that both
    # analysis streams stay busy while the copy engine works (without it each thread's upload sits between
    # its own steps, so for 5.4 of every 14 ms only one step's kernels are on the device).
    import queue as pyqueue

    class Pipeline:
        """the uploading thread and the analysis threads live for the whole end-to-end leg (warm-up and timed
        run): a host thread keeps its library stream and the device blocks cached on it, and the blocks an
        uploader needs (one batch of int16 + one of f32) are not the ones an analysis step recycles"""

        def __init__(self):
            self.jobs = pyqueue.Queue()                        # one token per step, None = shut down
            self.ready = pyqueue.Queue(maxsize=args.prefetch)  # uploaded batches waiting for an analysis thread
            self.done = pyqueue.Queue()                        # one entry per finished step: None or the exception
            self.threads = [threading.Thread(target=self.uploader, daemon=True)] + \
                           [threading.Thread(target=self.analyst, args=(t,), daemon=True) for t in range(nthr)]
            [t.start() for t in self.threads]

        def uploader(self):
            i = 0
            while self.jobs.get() is not None:
                try:
                    f = ctx.upload_scaled_ptr(pins_in[i % len(pins_in)], np.int16, shape4, INT16_SLOPE)
                except Exception as e:  # pragma: no cover
                    self.done.put(e)
                    continue
                self.ready.put(f)
                del f
                i += 1
            for _ in range(nthr):
                self.ready.put(None)

        def analyst(self, t):
            while True:
                f = self.ready.get()
                if f is None:
                    return
                try:
                    out = run_gtv(ctx, f)
                    del f
                    ctx.download_bits_ptr(out["gtv"], pins_out[t], out_bytes)
                    del out
                    self.done.put(None)
                except Exception as e:  # pragma: no cover
                    f = None
                    self.done.put(e)

        def run(self, nsteps):
            for _ in range(nsteps):
                self.jobs.put(1)
            for _ in range(nsteps):
                r = self.done.get()
                if r is not None:
                    raise r

        def close(self):
            self.jobs.put(None)
            [t.join(timeout=30) for t in self.threads]

    pipe = Pipeline() if args.prefetch > 0 else None

    def run_e2e(nsteps):
        if pipe is not None:
            return pipe.run(nsteps)
        errors = []
        queue = StepQueue(nsteps)
        ths = [threading.Thread(target=e2e_worker, args=(t, queue, errors)) for t in range(nthr)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        if errors:
            raise errors[0]

    run_e2e(2 * nthr + (args.prefetch + 2 if pipe is not None else 0))   # every block the pipeline cycles through exists
    barrier()
    # achieved host->device bandwidth of this rank with every rank copying at once (what bounds e2e
    # when the host fabric is shared): five uploads of the batch from pinned memory, untimed otherwise
    t0 = time.perf_counter()
    for _ in range(5):
        f = ctx.upload_scaled_ptr(pins_in[0], np.int16, shape4, INT16_SLOPE)
        del f
    ctx.sync()
    h2d_gbs = 5 * host.nbytes / (time.perf_counter() - t0) / 1e9
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    ctx.sync()
    ms_e2e = (time.perf_counter() - t0) * 1e3      # all streams drained: host clock around device work
    barrier()
    if pipe is not None:
        pipe.close()

    # ---- per-kernel table from an untimed pass with ONE host thread (one stream): launch
    # durations without other streams' kernels sharing the SMs, which is what a roofline
    # fraction of a single kernel means.  (Inside the timed region `args.threads` streams overlap,
    # so the per-launch times there include waiting for SMs held by other steps' kernels.)
    serial = {}
    if rank == 0:
        flair = ctx.upload_scaled(host, INT16_SLOPE)
        ctx.sync()
        step_resident(flair)
        ctx.sync()
        ctx.prof_reset(); ctx.prof_enable(True)
        for _ in range(4):
            step_resident(flair)
        ctx.sync()
        ctx.prof_enable(False)
        serial = ctx.prof_report()
        del flair

    slab = None
    by_rank = None
    if dist is not None and not args.no_slab:
        del pins_in, pins_out
        try:
            slab = slab_leg(ctx, dist, rank, world, local)
        except Exception as e:  # the slab record must never take the bench line down
            slab = {"error": f"{type(e).__name__}: {e}"}
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=f"cuda:{local}", dtype=torch.float64)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        by_rank = [[round(float(x[0]) / args.steps, 3), round(float(x[1]) / args.steps, 3)] for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        n = torch.tensor([launches], device=f"cuda:{local}", dtype=torch.int64)
        dist.all_reduce(n)
        launches = int(n[0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    total_vol = world * B * args.steps
    value = total_vol / (ms / 1e3)
    e2e = total_vol / (ms_e2e / 1e3)
    peak, how = peaks()
    # dominant kernel by device time, from the untimed single-stream pass that follows the timed
    # region (same process, same resident batch): per-launch CUDA-event durations without other
    # streams' kernels sharing the SMs.  The timed region itself carries no per-kernel events.
    tot_ms = sum(v[1] for v in prof.values()) or 1.0
    ser_total = sum(v[1] for v in serial.values())
    name = max(serial.items(), key=lambda kv: kv[1][1])[0]
    cnt, kms = serial[name]
    share_serial = (serial[name][1] / ser_total) if ser_total else None
    per_launch_vox = B * NVOX
    alg = ALG_BYTES.get(name) or 5.0
    achieved = alg * per_launch_vox / ((kms / cnt) * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            per_vox = json.load(open(tp)).get(name, {}).get("dram_bytes_per_voxel")
            traffic = per_vox * per_launch_vox if per_vox is not None else None   # bytes per launch
        except Exception:
            traffic = None
    kernels = {k: {"launches": c, "ms_total": round(m, 4), "share": round(m / tot_ms, 4),
                   "gvox_per_s": round(B * NVOX / ((m / c) * 1e-3) / 1e9, 2)}
               for k, (c, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    peak_v = peak
    kernels_serial = {}
    ser_tot = sum(v[1] for v in serial.values()) or 1.0
    for k, (c, m) in sorted(serial.items(), key=lambda kv: -kv[1][1]):
        row = {"launches_per_step": c / 4.0, "avg_launch_ms": round(m / c, 5), "share": round(m / ser_tot, 4)}
        ab = ALG_BYTES.get(k)
        if ab:
            gbs = ab * per_launch_vox / ((m / c) * 1e-3) / 1e9
            row.update({"alg_bytes_per_voxel": ab, "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak_v, 4)})
        kernels_serial[k] = row
    cb = cpu_baseline(os.cpu_count() or 1) if world == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/int64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "volumes_per_step_per_gpu": B, "voxels_per_volume": NVOX, "adjacency": 26,
                   "sharding": "independent volumes per GPU, no collective", "host_threads": nres,
                   "l2": f"inputs larger than L2 ({B * NVOX * 4 / 1e6:.0f} MB f32 batch per step)",
                   "input": "int16 voxels with scl_slope 1/4095 (a NIfTI file's content); the f32 image is formed on "
                            "the device by vx_upload_scaled, bit-identical to converting on the host",
                   "data_note": "min(batch,4) generated volumes per rank + axis flips"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "ms_per_step": ms_e2e / args.steps, "host_threads": nthr, "prefetch": args.prefetch,
                "upload_thread": "own host thread, at most `prefetch` uploaded batches ahead" if args.prefetch > 0 else "none",
                "numa": numa,
                "h2d_gbs_this_rank_all_ranks_copying": round(h2d_gbs, 2),
                "timing": "host perf_counter around K steps with all streams synchronised on both sides"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": how, "traffic": traffic,
                     "alg_bytes_per_voxel": alg, "avg_launch_ms": kms / cnt,
                     # comparable with the ncu launch list (profiles/), which serialises the launches
                     "share_of_step": share_serial,
                     "how": "CUDA events around every launch of 4 untimed single-stream steps run right after "
                            "the timed region (kernels_serial); the timed region carries no per-kernel events",
                     "traffic_unit": "bytes per launch (dram read+write, ncu --set full, profiles/traffic.json)",
                     "note": xcorr_note() if name == "xcorr_kernel" else ""},
        "kernels": kernels,
        "kernels_serial": {"note": "one host thread / one stream, 4 untimed steps after the timed region: "
                                   "per-launch durations without cross-stream sharing of the SMs; "
                                   "achieved_gbs = algorithmic bytes per launch / duration",
                           "ms_per_step": round(ser_tot / 4.0, 4), "kernels": kernels_serial},
        "clocks": clocks,
        "cpu_baseline": cb,
        "slab": slab,
        "ms_per_step_by_rank": by_rank,   # [resident, end to end] of every rank; value and e2e use the slowest
        "gtv_voxels_last_step": [float(v) for v in (vols if vols is not None else [])][:4],
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="BraTS volumes per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads", type=int, default=2,
                    help="host threads (= CUDA streams of the library) that take the steps in turn: the kernels of one "
                         "step fill the gaps and tails of the other's (three or four change nothing: the device is "
                         "busy).  In the end-to-end leg these are the analysis threads; see --prefetch")
    ap.add_argument("--prefetch", type=int, default=2,
                    help="end-to-end leg: uploaded batches the uploading host thread keeps ahead of the analysis threads "
                         "(0: every analysis thread uploads its own step's batch, the round-1 scheme)")
    ap.add_argument("--wc", action="store_true", help="write-combined pinned staging buffers for the uploads")
    ap.add_argument("--independent-uploads", action="store_true",
                    help="diagnosis: turn VX_OPT_ORDERED_UPLOADS off (uploads of different threads share the link)")
    ap.add_argument("--no-slab", action="store_true", help="skip the slab-decomposed volume record of N > 1 runs")
    ap.add_argument("--prof-timed", action="store_true",
                    help="diagnosis: per-kernel CUDA events inside the timed region (default: untimed pass only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
