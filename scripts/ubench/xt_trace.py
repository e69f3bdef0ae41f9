"""Timeline of one block of the crossCorrelation kernel (debug build with -DXT_TRACE):
copy warp issue times, flag warp full/ready times, consumer warp 0 step times."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from voxlogica_b200 import _abi  # noqa: E402

_abi._lib = _abi.load_library(os.path.join(ROOT, "scripts", "ubench", "libvoxcuda_trace.so"))
from voxlogica_b200 import Context, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "zeros"
ctx = Context(0)
B = 16
shape = (155, 240, 240)
if mode == "zeros":
    a = ctx.upload(np.zeros((B,) + shape, np.float32))
    m, M = 0.0, 1.0
else:
    h = np.stack([np.roll(synth.brats_like(1234, shape), i * 3, axis=2) for i in range(B)])
    a = ctx.upload(h)
    m, M = ctx.min(a), ctx.max(a)
reg = ctx.cmp_scalar(a, ">", 0.75) if mode != "zeros" else ctx.cmp_scalar(a, ">", -1.0)
for _ in range(2):
    r = ctx.cross_correlation(5.0, a, a, reg, m, M, 100)
ctx.sync()
lib = _abi.lib()
buf = (C.c_longlong * 1024)()
lib.vx_debug_xt_trace.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.vx_debug_xt_trace(buf))
t = np.frombuffer(buf, dtype=np.int64).copy()
t0 = t[0]
rel = lambda i: int(t[i] - t0) if t[i] else None
print("block start (before set-up)", rel(998), "block end (consumer warp 0)", rel(999))
print("consumer: enter", rel(300), "first chunks ready", rel(301), "init done", rel(302))
print("copy warp issue times  :", [rel(1 + j) for j in range(34)])
print("flag warp full times   :", [rel(100 + j) for j in range(34)])
print("flag warp ready times  :", [rel(200 + j) for j in range(34)])
print("consumer step times    :", [rel(400 + j) for j in range(0, 120)])
print("fast-path ensure passed:", [rel(600 + j) for j in range(0, 120, 4)])
