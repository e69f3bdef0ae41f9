"""Diagnostic: wall-clock per GTV step and host time spent inside each operator call."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from voxlogica_b200 import Context
from voxlogica_b200.imgql import run_gtv
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = Context(0)
host = bench.make_batch(B, 0)
flair = ctx.upload(host); ctx.sync()
acc = collections.defaultdict(float)
for name in dir(Context):
    if name.startswith("_") or name in ("event", "record", "elapsed_ms", "sync", "close", "prof_report", "mem_info"):
        continue
    fn = getattr(ctx, name)
    if not callable(fn):
        continue
    def wrap(fn=fn, name=name):
        def w(*a, **k):
            t = time.perf_counter(); r = fn(*a, **k); acc[name] += time.perf_counter() - t; return r
        return w
    setattr(ctx, name, wrap())
for s in range(steps):
    acc.clear()
    e0, e1 = ctx.event(), ctx.event()
    t0 = time.perf_counter()
    ctx.record(e0)
    out = run_gtv(ctx, flair)
    vol = out["_printed"]["gtvVolume"]
    ctx.record(e1)
    t1 = time.perf_counter()
    del out
    t2 = time.perf_counter()
    gpu = ctx.elapsed_ms(e0, e1)
    top = sorted(acc.items(), key=lambda kv: -kv[1])[:6]
    print(f"step {s}: wall {1e3*(t1-t0):.1f} ms, release {1e3*(t2-t1):.1f} ms, gpu {gpu:.1f} ms, mem {ctx.mem_info()}",
          [(k, round(1e3*v, 2)) for k, v in top], flush=True)
