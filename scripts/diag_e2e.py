"""Diagnostic: where the end-to-end step of bench.py spends its host time.
(1) one host thread, batch resident: wall per step against the device time of the step;
(2) the prefetch pipeline (one uploading thread + N analysis threads): per-call host times
    (upload call, wait for a batch, run_gtv, download) and the step rate.
usage: python scripts/diag_e2e.py [batch] [steps] [analysis threads] [prefetch]"""
import os, sys, time, threading, queue
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from voxlogica_b200 import Context
from voxlogica_b200.imgql import run_gtv
from voxlogica_b200.synth import INT16_SLOPE
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
nthr = int(sys.argv[3]) if len(sys.argv) > 3 else 2
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ctx = Context(0)
host = bench.make_batch(B, 0)
shape4 = (B,) + bench.BRATS
out_bytes = (B * bench.NVOX + 7) // 8

flair = ctx.upload_scaled(host, INT16_SLOPE); ctx.sync()
for _ in range(2):
    run_gtv(ctx, flair)
ctx.sync()
t0 = time.perf_counter()
e0, e1 = ctx.event(), ctx.event()
ctx.record(e0)
for _ in range(6):
    out = run_gtv(ctx, flair)
ctx.record(e1); ctx.sync()
print(f"one thread, resident: wall {1e3 * (time.perf_counter() - t0) / 6:.2f} ms/step, device {ctx.elapsed_ms(e0, e1) / 6:.2f} ms/step",
      flush=True)
del flair, out

pins_in = [ctx.host_alloc(host.nbytes) for _ in range(nthr + depth)]
pins_out = [ctx.host_alloc(out_bytes) for _ in range(nthr)]
for p in pins_in:
    C.memmove(p, host.ctypes.data, host.nbytes)


def pipeline(nsteps, log):
    ready = queue.Queue(maxsize=depth)
    T0 = time.perf_counter()

    def uploader():
        for i in range(nsteps):
            a = time.perf_counter()
            f = ctx.upload_scaled_ptr(pins_in[i % len(pins_in)], np.int16, shape4, INT16_SLOPE)
            b = time.perf_counter()
            ready.put(f)
            c = time.perf_counter()
            log.append(("up", i, round(1e3 * (a - T0), 2), round(1e3 * (b - a), 2), round(1e3 * (c - b), 2)))
            del f
        for _ in range(nthr):
            ready.put(None)

    def analyst(t):
        while True:
            a = time.perf_counter()
            f = ready.get()
            if f is None:
                return
            b = time.perf_counter()
            out = run_gtv(ctx, f)
            c = time.perf_counter()
            del f
            ctx.download_bits_ptr(out["gtv"], pins_out[t], out_bytes)
            d = time.perf_counter()
            del out
            e = time.perf_counter()
            log.append((f"an{t}", round(1e3 * (a - T0), 2), "wait", round(1e3 * (b - a), 2), "gtv", round(1e3 * (c - b), 2),
                        "down", round(1e3 * (d - c), 2), "release", round(1e3 * (e - d), 2)))

    ths = [threading.Thread(target=uploader)] + [threading.Thread(target=analyst, args=(t,)) for t in range(nthr)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    ctx.sync()
    return time.perf_counter() - T0


pipeline(2 * nthr, [])
log = []
sec = pipeline(steps, log)
print(f"prefetch pipeline: {nthr} analysis threads, depth {depth}: {1e3 * sec / steps:.2f} ms/step", flush=True)
for row in sorted(log, key=lambda r: r[2] if r[0] == "up" else r[1]):
    print(row)
print("mem", ctx.mem_info())
