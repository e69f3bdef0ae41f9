"""Minimal ImgQL driver: tokenizer, parser, memoising evaluator over the C-ABI operators.

SURVEY.md section 8(f) rank 1.  VoxLogicA's real front end (F# parser + reducer, Greta.pdf
p.7 "syntactic manipulation, memoization") is neither mounted nor runnable here, so this
small driver exists to run specs such as the GTV one on Greta.pdf p.48 *unchanged* on top
of libvoxcuda.  It is a harness, not a re-implementation of VoxLogicA: it supports

    import "stdlib.imgql"          load x = "file"         let x = e
    let f(a, b) = e                save "file" e           print "label" e

with function application, numbers, prefix ``!`` and the dotted infix operators
(``<. >. <=. >=. =.`` image-vs-number, ``.< .> ...`` number-vs-image, ``+ - * /`` on
images, ``+. .+ .+.`` mixed / numeric, ``& |`` boolean).  Every operator evaluates to a
call into ``ops.Context``; identical sub-terms are evaluated once (hash-consing), which is
the memoisation the reference's evaluator performs.

The derived operators are restated from the slide semantics in
STDLIB below; the real stdlib.imgql is not in the mount.
"""
from __future__ import annotations

import re
import weakref
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .ops import Context, Image

STDLIB = """
// derived operators, restated (Greta.pdf p.35-38, p.48, p.51-53)
let complement(a) = !a
let not(a) = !a
let and(a,b) = a & b
let or(a,b) = a | b
let N(a) = near(a)
let I(a) = interior(a)
let touch(a,b) = a & through(near(b), a)
let grow(a,b) = a | touch(b,a)
let flt(r,a) = distlt(r, distgeq(r, !a))
let filter(r,a) = flt(r,a)
let reach(t,p) = near(t) | near(through(near(t), p))
let surrounded(a,b) = a & !reach(!(a | b), !b)
"""

GTV_SPEC = """
import "stdlib.imgql"
load flair = "flair"
let similarTo(a) = crossCorrelation(5, intensity(flair), intensity(flair), a, min(intensity(flair)), max(intensity(flair)), 100)
// Greta.pdf p.48 (Image290), constants per SURVEY.md Q5
let background   = touch(intensity(flair) <. 0.1, border)
let brain        = complement(background)
let normFlair    = percentiles(intensity(flair), brain)
let hyperIntense = filter(5.0, normFlair >. 0.95)
let veryIntense  = filter(2.0, normFlair >. 0.86)
let growTum      = grow(hyperIntense, veryIntense)
let tumSim       = similarTo(growTum)
let tumStatCC    = filter(2.0, tumSim >. 0.6)
let gtv          = grow(growTum, tumStatCC)
save "gtv" gtv
print "gtvVolume" volume(gtv)
"""

# BASELINE config 4: texture similarity on co-registered modalities.  The tumour seed region
# (growTum of the GTV spec, from FLAIR) is compared with every voxel's neighbourhood in each
# modality (Greta.pdf p.43 "Texture analysis", p.53); the mean coefficient is thresholded.
TEXTURE_SPEC = """
import "stdlib.imgql"
load flair = "flair"
load t1 = "t1"
load t1c = "t1c"
load t2 = "t2"
let background   = touch(intensity(flair) <. 0.1, border)
let brain        = complement(background)
let normFlair    = percentiles(intensity(flair), brain)
let hyperIntense = filter(5.0, normFlair >. 0.95)
let veryIntense  = filter(2.0, normFlair >. 0.86)
let growTum      = grow(hyperIntense, veryIntense)
let simFlair = crossCorrelation(5, intensity(flair), intensity(flair), growTum, min(intensity(flair)), max(intensity(flair)), 100)
let simT1    = crossCorrelation(5, intensity(t1), intensity(t1), growTum, min(intensity(t1)), max(intensity(t1)), 100)
let simT1c   = crossCorrelation(5, intensity(t1c), intensity(t1c), growTum, min(intensity(t1c)), max(intensity(t1c)), 100)
let simT2    = crossCorrelation(5, intensity(t2), intensity(t2), growTum, min(intensity(t2)), max(intensity(t2)), 100)
let texture  = (simFlair +. simT1 +. simT1c +. simT2) /. 4
let similar  = filter(2.0, (texture >. 0.6) & brain)
let tumour   = grow(growTum, similar)
save "texture" texture
save "tumour" tumour
print "tumourVolume" volume(tumour)
"""

# ---------------------------------------------------------------------------- lexer
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<num>\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<str>"[^"]*")
  | (?P<punct>[(),])
  | (?P<op>[!<>=.&|+\-*/^~%#@$]+)
""", re.X)


def tokenize(src: str) -> List[Tuple[str, str]]:
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"ImgQL: unexpected character {src[pos]!r} at offset {pos}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        text = m.group()
        if kind == "op" and text == "=":
            kind = "eq"
        out.append((kind, text))
    out.append(("eof", ""))
    return out


# ---------------------------------------------------------------------------- AST
@dataclass(frozen=True)
class Num:
    v: float


@dataclass(frozen=True)
class Var:
    name: str


@dataclass(frozen=True)
class Call:
    fn: str
    args: tuple


_PREC = {"|": 1, "&": 2, "<": 3, "<=": 3, ">": 3, ">=": 3, "=": 3, "==": 3, "!=": 3,
         "+": 4, "-": 4, "*": 5, "/": 5}


class Parser:
    def __init__(self, src: str):
        self.t = tokenize(src)
        self.i = 0

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expect(self, kind, text=None):
        k, s = self.next()
        if k != kind or (text is not None and s != text):
            raise SyntaxError(f"ImgQL: expected {text or kind}, found {s!r}")
        return s

    def program(self):
        stmts = []
        while self.peek()[0] != "eof":
            k, s = self.next()
            if k != "id":
                raise SyntaxError(f"ImgQL: expected a statement, found {s!r}")
            if s == "import":
                stmts.append(("import", self.expect("str").strip('"')))
            elif s == "load":
                name = self.expect("id")
                self.expect("eq")
                stmts.append(("load", name, self.expect("str").strip('"')))
            elif s == "let":
                k2, name = self.next()
                if k2 not in ("id", "op"):
                    raise SyntaxError(f"ImgQL: bad name {name!r} in let")
                params = None
                if self.peek() == ("punct", "("):
                    self.next()
                    params = []
                    while self.peek() != ("punct", ")"):
                        params.append(self.expect("id"))
                        if self.peek() == ("punct", ","):
                            self.next()
                    self.next()
                self.expect("eq")
                stmts.append(("let", name, params, self.expr(0)))
            elif s == "save":
                path = self.expect("str").strip('"')
                stmts.append(("save", path, self.expr(0)))
            elif s == "print":
                label = self.expect("str").strip('"')
                stmts.append(("print", label, self.expr(0)))
            else:
                raise SyntaxError(f"ImgQL: unknown statement {s!r}")
        return stmts

    def expr(self, minprec):
        lhs = self.unary()
        while True:
            k, s = self.peek()
            if k == "eq":
                k, s = "op", "="
            if k != "op":
                break
            core = s.strip(".")
            if core not in _PREC or _PREC[core] < minprec:
                break
            self.next()
            rhs = self.expr(_PREC[core] + 1)
            lhs = Call("infix:" + s, (lhs, rhs))
        return lhs

    def unary(self):
        k, s = self.peek()
        if k == "op" and s == "!":
            self.next()
            return Call("prefix:!", (self.unary(),))
        if k == "op" and s == "-":
            self.next()
            return Call("neg", (self.unary(),))
        return self.atom()

    def atom(self):
        k, s = self.next()
        if k == "num":
            return Num(float(s))
        if k == "punct" and s == "(":
            e = self.expr(0)
            self.expect("punct", ")")
            return e
        if k == "id":
            if self.peek() == ("punct", "("):
                self.next()
                args = []
                while self.peek() != ("punct", ")"):
                    args.append(self.expr(0))
                    if self.peek() == ("punct", ","):
                        self.next()
                self.next()
                return Call(s, tuple(args))
            return Var(s)
        raise SyntaxError(f"ImgQL: unexpected token {s!r}")


# ---------------------------------------------------------------------------- pointwise fusion
# "boolean and intensity pointwise chains ... fused along the DAG" (BASELINE.json north_star):
# pointwise operators do not run when they are applied; they build a Lazy node.  When a
# non-pointwise operator (or save / print / numpy()) needs the value, the whole pending sub-DAG
# is compiled into ONE vx_fused_pointwise program: leaf images are read once, the root is
# written once, intermediates stay in registers / shared memory.
_B, _F = "b", "f"


class Lazy:
    """A pending pointwise expression over HBM-resident leaves."""

    __slots__ = ("ctx", "op", "args", "aux", "imm", "kind", "_img", "__weakref__")

    def __init__(self, ctx, op, args, kind, aux=0, imm=0.0):
        self.ctx, self.op, self.args, self.kind, self.aux, self.imm = ctx, op, tuple(args), kind, aux, imm
        self._img = None

    # -- the Image protocol used by the evaluator, the tests and the bench
    def force(self) -> Image:
        if self._img is None:
            try:
                self._img = _compile_and_run(self)
            except _FusionLimit:
                for a in self.args:       # too wide for one program: materialise the operands
                    force(a)
                self._img = _compile_and_run(self)
            self.args = ()            # drop the sub-DAG: its leaves may be freed now
        return self._img

    def numpy(self):
        return self.force().numpy()

    def info(self):
        return self.force().info()

    @property
    def handle(self):
        return self.force().handle

    @property
    def nb(self):
        return _first_leaf(self).nb

    @property
    def shape(self):
        return _first_leaf(self).shape

    @property
    def spacing(self):
        return _first_leaf(self).spacing


class _FusionLimit(RuntimeError):
    pass


def _first_leaf(v):
    while isinstance(v, Lazy):
        if v._img is not None:
            return v._img
        v = next(a for a in v.args if isinstance(a, (Lazy, Image)))
    return v


def force(v):
    """Lazy -> Image (anything else is returned unchanged)"""
    return v.force() if isinstance(v, Lazy) else v


def _count_ops(v, seen) -> int:
    if not isinstance(v, Lazy) or v._img is not None or id(v) in seen:
        return 0
    seen.add(id(v))
    return 1 + sum(_count_ops(a, seen) for a in v.args)


def _compile_and_run(root: "Lazy") -> Image:
    from . import _abi as A
    ctx = root.ctx
    # keep programs inside the limits of the register machine: materialise big operands first
    while True:
        leaves, seen = {}, set()

        def scan(v):
            if isinstance(v, Lazy) and v._img is None:
                if id(v) not in seen:
                    seen.add(id(v))
                    for a in v.args:
                        scan(a)
            elif isinstance(v, (Lazy, Image)):
                leaves[force(v).handle] = force(v)
        scan(root)
        if len(leaves) <= A.VXP_MAX_LEAVES and len(seen) + len(leaves) <= A.VXP_MAX_OPS - 4 and len(seen) <= 40:
            break
        big = max((a for a in root.args if isinstance(a, Lazy) and a._img is None),
                  key=lambda a: _count_ops(a, set()), default=None)
        if big is None:
            raise _FusionLimit("pointwise expression does not fit the fusion limits")
        big.force()
    leaf_idx = {h: i for i, h in enumerate(leaves)}
    leaf_imgs = list(leaves.values())
    # post-order emission with use counts for register reuse
    uses = {}

    def count(v):
        if isinstance(v, Lazy) and v._img is None:
            uses[id(v)] = uses.get(id(v), 0) + 1
            if uses[id(v)] == 1:
                for a in v.args:
                    count(a)
        elif isinstance(v, (Lazy, Image)):
            h = force(v).handle
            uses[("leaf", h)] = uses.get(("leaf", h), 0) + 1
    count(root)
    prog, reg_of, free = [], {}, list(range(A.VXP_MAX_REGS - 1, -1, -1))
    scalars = []

    def release(key):
        uses[key] -= 1
        if uses[key] == 0:
            free.append(reg_of.pop(key))

    def emit(v):
        if isinstance(v, Lazy) and v._img is None:
            key = id(v)
            if key in reg_of:
                return key
            keys = [emit(a) for a in v.args if isinstance(a, (Lazy, Image))]
            regs = [reg_of[k] for k in keys]
            for k in keys:
                release(k)
            if not free:
                raise _FusionLimit("out of fusion registers")
            dst = free.pop()
            a = regs[0] if regs else 0
            b = regs[1] if len(regs) > 1 else 0
            opcode, aux, imm = v.op, v.aux, v.imm
            if opcode in (A.VXP_CMPSB, A.VXP_ARITHSB):
                b = len(scalars)
                scalars.append(np.asarray(v.imm, dtype=np.float32))
                imm = 0.0
            prog.append((opcode, dst, a, b, int(aux), float(imm)))
            reg_of[key] = dst
            return key
        h = force(v).handle
        key = ("leaf", h)
        if key not in reg_of:
            if not free:
                raise _FusionLimit("out of fusion registers")
            dst = free.pop()
            prog.append((A.VXP_LOAD, dst, leaf_idx[h], 0, 0, 0.0))
            reg_of[key] = dst
        return key
    try:
        emit(root)
    finally:
        # the recursive closures above refer to themselves through their own cells: clear them,
        # otherwise the leaf images stay alive until Python's cyclic collector happens to run
        # (and their HBM blocks miss the allocator's free lists for the next batch)
        scan = count = emit = release = None
        leaves.clear(); uses.clear(); reg_of.clear()
    bs = np.stack(scalars) if scalars else None
    return ctx.fused(prog, leaf_imgs, bs)


# ---------------------------------------------------------------------------- evaluator
class Number:
    """An ImgQL number: one value per volume of the batch."""

    def __init__(self, v):
        self.v = np.atleast_1d(np.asarray(v, dtype=np.float64))

    def __repr__(self):
        return f"Number({self.v.tolist()})"


class Loaded:
    """Result of ``load``: the channels of one image file, each an HBM-resident f32 Image."""

    def __init__(self, channels: Dict[str, Image]):
        self.channels = channels

    def channel(self, name: str) -> Image:
        if name in self.channels:
            return self.channels[name]
        raise KeyError(f"ImgQL: image has no {name!r} channel (has {sorted(self.channels)})")


def _is_img(v):
    # `is_image`: image-like values of other back ends of the driver (a z-slab of a distributed
    # volume, voxlogica_b200.slab.Slab)
    return isinstance(v, (Image, Lazy)) or getattr(v, "is_image", False)


class Interpreter:
    def __init__(self, ctx: Context, loader: Optional[Callable[[str], Image]] = None, conn: int = 26,
                 fuse: bool = True, gc: bool = False):
        self.ctx = ctx
        self.fuse = fuse                     # build Lazy nodes for pointwise operators
        # Device-memory GC policy (Greta.pdf p.78-79: the reference's GPU branch runs out of memory at
        # 4099 tasks unless intermediate results are collected).  With gc=True an intermediate lives
        # exactly as long as something can still consume it: the memo table holds image values
        # weakly, and a `let`-bound name is dropped after the last statement that mentions it
        # (directly or through a function body).  The handle's reference count then reaches zero and
        # the block goes back to the stream's free list for the next task.  A value that is asked
        # for again after it died is simply recomputed (operators have no side effects, p.7).
        self.gc = gc
        self._uid = weakref.WeakKeyDictionary()   # value -> serial number (memo keys must outlive id())
        self._next_uid = 0
        self._depth = 0
        self.loader = loader
        self.saver = None                    # optional callable(path, Image) run by `save`
        self.conn = conn                     # adjacency of near/through
        self.globals: Dict[str, object] = {}
        self.functions: Dict[str, Tuple[List[str], object]] = {}
        self.base: Optional[Image] = None    # shape provider for border / tt / ff
        self.memo: Dict[tuple, object] = {}
        self.saved: Dict[str, Image] = {}
        self.printed: Dict[str, object] = {}
        self.n_tasks = 0                     # operator calls actually executed

    # -- program level
    _SCALAR_FNS = ("min", "max", "volume")

    def _hoist_scalars(self, stmts):
        """The reference's scheduler runs independent tasks in parallel, so a global reduction of
        a LOADED image (min/max of the input, needed later as histogram range) is computed at
        the start, not where the spec text happens to mention it.  This pre-pass finds such
        calls (arguments built only from loaded images and `intensity`/channel selectors) and
        evaluates them right away; the memo then serves the later uses.  Evaluating them late
        would drain the CUDA stream in the middle of the pipeline just to read one number."""
        loaded = {k for k, v in self.globals.items() if isinstance(v, Loaded) or _is_img(v)}
        found = []
        for st in stmts:
            if st[0] == "let":
                _find_hoistable(st[3], loaded, self._SCALAR_FNS, found)
            elif st[0] in ("save", "print"):
                _find_hoistable(st[2], loaded, self._SCALAR_FNS, found)
        # min(e) and max(e) of the same image: ONE pass (vx_minmax) serves both tasks
        for e in found:
            twin = next((f for f in found if f is not e and {e.fn, f.fn} == {"min", "max"} and f.args == e.args), None)
            if e.fn == "min" and twin is not None:
                arg = self.eval(e.args[0], {})
                klo, khi = self._key("min", [arg]), self._key("max", [arg])
                if klo not in self.memo and khi not in self.memo:
                    lo, hi = self.ctx.minmax(force(arg))
                    self.memo[klo] = (Number(lo), [arg])
                    self.memo[khi] = (Number(hi), [arg])
                    self.n_tasks += 2
        for e in found:
            self.eval(e, {})

    def _last_uses(self, stmts):
        """name -> index of the last statement that can read it (through function bodies too)"""
        def names(e, acc):
            if isinstance(e, Var):
                acc.add(e.name)
            elif isinstance(e, Call):
                acc.add(e.fn)
                for a in e.args:
                    names(a, acc)
            return acc
        fn_reads = {n: names(b, set()) for n, (_, b) in self.functions.items()}
        for st in stmts:
            if st[0] == "let" and st[2] is not None:
                fn_reads[st[1]] = names(st[3], set()) - set(st[2])
        last = {}
        for i, st in enumerate(stmts):
            body = st[3] if st[0] == "let" else st[2] if st[0] in ("save", "print") else None
            if body is None:
                continue
            reads, todo = set(), list(names(body, set()))
            while todo:
                n = todo.pop()
                if n not in reads:
                    reads.add(n)
                    todo.extend(fn_reads.get(n, ()))
            for n in reads:
                last[n] = i
        return last

    def run(self, src: str):
        stmts = Parser(src).program()
        hoisted = False
        # imported files only define things for the importing program: liveness is decided at the top level
        last_use = self._last_uses(stmts) if self.gc and not self._depth else None
        self._depth += 1
        for i, st in enumerate(stmts):
            kind = st[0]
            if not hoisted and kind not in ("import", "load"):
                self._hoist_scalars(stmts[i:])
                hoisted = True
            if kind == "import":
                if st[1].endswith("stdlib.imgql"):
                    self.run(STDLIB)
                else:
                    with open(st[1]) as f:
                        self.run(f.read())
            elif kind == "load":
                if self.loader is None:
                    raise RuntimeError("ImgQL: no loader configured for `load`")
                img = self.loader(st[2])
                self.globals[st[1]] = img
                if self.base is None:
                    self.base = next(iter(img.channels.values())) if isinstance(img, Loaded) else img
            elif kind == "let":
                _, name, params, body = st
                if params is None:
                    self.globals[name] = self.eval(body, {})
                else:
                    self.functions[name] = (params, body)
            elif kind == "save":
                self.saved[st[1]] = force(self.eval(st[2], {}))
                if self.saver is not None:
                    self.saver(st[1], self.saved[st[1]])
            elif kind == "print":
                v = self.eval(st[2], {})
                self.printed[st[1]] = v.v.copy() if isinstance(v, Number) else v
            if last_use is not None:
                # liveness from the program text: a name nobody reads any more gives up its value
                for name in [n for n, v in self.globals.items() if last_use.get(n, -1) <= i and _is_img(v)]:
                    del self.globals[name]
        self._depth -= 1
        return self

    # -- expression level.  Values are identified by id() for hash-consing.
    def eval(self, e, env):
        if isinstance(e, Num):
            return Number(e.v)
        if isinstance(e, Var):
            if e.name in env:
                return env[e.name]
            if e.name in self.globals:
                return self.globals[e.name]
            if e.name in ("border", "tt", "ff"):
                return self.apply(e.name, [])
            raise NameError(f"ImgQL: unbound identifier {e.name!r}")
        if isinstance(e, Call):
            args = [self.eval(a, env) for a in e.args]
            if e.fn in self.functions:
                params, body = self.functions[e.fn]
                if len(params) != len(args):
                    raise TypeError(f"ImgQL: {e.fn} takes {len(params)} arguments")
                return self.eval(body, dict(zip(params, args)))
            return self.apply(e.fn, args)
        raise TypeError(e)

    def _key(self, fn, args):
        ks = []
        for a in args:
            if isinstance(a, Number):
                ks.append(("n", a.v.tobytes()))
            elif self.gc:
                # serial numbers instead of id(): the memo does not keep operands alive in this
                # mode, and a dead operand's id() may be handed to a new object
                u = self._uid.get(a)
                if u is None:
                    u = self._uid[a] = self._next_uid
                    self._next_uid += 1
                ks.append(("u", u))
            else:
                ks.append(("i", id(a)))
        return (fn, tuple(ks))

    def apply(self, fn, args):
        key = self._key(fn, args)
        hit = self.memo.get(key)
        if hit is not None:
            val = hit[0]() if isinstance(hit[0], weakref.ref) else hit[0]
            if val is not None:
                return val
        val = self._apply(fn, args)
        if self.gc and _is_img(val):
            memo = self.memo
            self.memo[key] = (weakref.ref(val, lambda _r, k=key: memo.pop(k, None)), None)
        else:
            self.memo[key] = (val, args)   # keep args alive so id() stays unique
        self.n_tasks += 1
        return val

    def _like(self) -> Image:
        if self.base is None:
            raise RuntimeError("ImgQL: border/tt/ff need a loaded image for their shape")
        return self.base

    def _apply(self, fn, a):
        c = self.ctx
        if fn.startswith("infix:"):
            return self._infix(fn[6:], a[0], a[1])
        if fn == "prefix:!":
            return self._pw_not(a[0])
        if fn == "neg":
            return Number(-a[0].v) if isinstance(a[0], Number) else self._pw_arith_scalar(a[0], "r-", 0.0)
        if fn == "mask":
            return self._pw_mask(a[0], a[1])
        a = [force(x) for x in a]     # every other operator needs materialised operands
        table = {
            "intensity": self._intensity,
            "red": lambda x: x.channel("red"),
            "green": lambda x: x.channel("green"),
            "blue": lambda x: x.channel("blue"),
            "alpha": lambda x: x.channel("alpha"),
            "border": lambda: c.border(self._like()),
            "tt": lambda: c.tt(self._like()),
            "ff": lambda: c.ff(self._like()),
            "near": lambda x: c.near(x, self.conn),
            "interior": lambda x: c.interior(x, self.conn),
            "through": lambda x, y: c.through(x, y, self.conn),
            "maxvol": lambda x: c.maxvol(x, self.conn),
            "distlt": lambda r, x: c.distlt(r.v[0], x),
            "distleq": lambda r, x: c.distleq(r.v[0], x),
            "distgeq": lambda r, x: c.distgeq(r.v[0], x),
            "distgt": lambda r, x: c.distgt(r.v[0], x),
            "dt": lambda x: c.edt_signed(x, self.conn),
            "edt": lambda x: c.edt(x),
            "volume": lambda x: Number(c.volume(x)),
            "min": lambda x: Number(c.min(x)),
            "max": lambda x: Number(c.max(x)),
            "constant": lambda v: c.constant(self._like(), v.v[0]),
            "percentiles": lambda x, m, corr=None: c.percentiles(x, m, 0.0 if corr is None else corr.v[0]),
            "crossCorrelation": lambda rho, x, y, reg, m, M, k: c.cross_correlation(
                rho.v[0], x, y, reg, m.v, M.v, int(k.v[0])),
        }
        if fn not in table:
            raise NameError(f"ImgQL: unknown operator {fn!r}")
        return table[fn](*a)

    def _intensity(self, x):
        """single-channel: the identity; RGB: Rec.709 luma 0.2126 r + 0.7152 g + 0.0722 b [RECALL]"""
        if not isinstance(x, Loaded):
            return x
        if "intensity" in x.channels:
            return x.channels["intensity"]
        r, g, b = (x.channel(n) for n in ("red", "green", "blue"))
        return self._pw_arith(self._pw_arith(self._pw_arith_scalar(r, "*", 0.2126), "+",
                                             self._pw_arith_scalar(g, "*", 0.7152)), "+",
                              self._pw_arith_scalar(b, "*", 0.0722))

    # -- pointwise constructors: Lazy nodes when fusing, direct calls otherwise
    def _pw_not(self, x):
        from . import _abi as A
        if not self.fuse:
            return self.ctx.not_(force(x))
        return Lazy(self.ctx, A.VXP_NOT, [x], _B)

    def _pw_bool(self, core, l, r):
        from . import _abi as A
        if not self.fuse:
            return self.ctx.and_(force(l), force(r)) if core == "&" else self.ctx.or_(force(l), force(r))
        return Lazy(self.ctx, A.VXP_AND if core == "&" else A.VXP_OR, [l, r], _B)

    def _pw_cmp_scalar(self, x, k, s):
        from . import _abi as A
        from .ops import _CMP
        if not self.fuse:
            return self.ctx.cmp_scalar(force(x), k, s)
        if np.ndim(s) == 0:
            return Lazy(self.ctx, A.VXP_CMPS, [x], _B, _CMP[k], float(s))
        return Lazy(self.ctx, A.VXP_CMPSB, [x], _B, _CMP[k], np.broadcast_to(np.asarray(s, np.float32), (x.nb,)))

    def _pw_cmp(self, l, k, r):
        from . import _abi as A
        from .ops import _CMP
        if not self.fuse:
            return self.ctx.cmp(force(l), k, force(r))
        return Lazy(self.ctx, A.VXP_CMP, [l, r], _B, _CMP[k])

    def _pw_arith_scalar(self, x, op, s):
        from . import _abi as A
        from .ops import _ARITH
        if not self.fuse:
            return self.ctx.arith_scalar(force(x), op, s)
        if np.ndim(s) == 0:
            return Lazy(self.ctx, A.VXP_ARITHS, [x], _F, _ARITH[op], float(s))
        return Lazy(self.ctx, A.VXP_ARITHSB, [x], _F, _ARITH[op], np.broadcast_to(np.asarray(s, np.float32), (x.nb,)))

    def _pw_arith(self, l, op, r):
        from . import _abi as A
        from .ops import _ARITH
        if not self.fuse:
            return self.ctx.arith(force(l), op, force(r))
        return Lazy(self.ctx, A.VXP_ARITH, [l, r], _F, _ARITH[op])

    def _pw_mask(self, x, m):
        from . import _abi as A
        if not self.fuse:
            return self.ctx.mask(force(x), force(m), 0.0)
        return Lazy(self.ctx, A.VXP_SELECT, [m, x], _F, 0, 0.0)

    def _infix(self, op, l, r):
        core = op.strip(".")
        cmp_map = {"<": "<", "<=": "<=", ">": ">", ">=": ">=", "=": "==", "==": "==", "!=": "!="}
        swap = {"<": ">", "<=": ">=", ">": "<", ">=": "<=", "==": "==", "!=": "!="}
        if core in ("&", "|"):
            return self._pw_bool(core, l, r)
        li, ri = _is_img(l), _is_img(r)
        scal = lambda n: n.v[0] if n.v.size == 1 else n.v
        if core in cmp_map:
            k = cmp_map[core]
            if li and ri:
                return self._pw_cmp(l, k, r)
            if li:
                return self._pw_cmp_scalar(l, k, scal(r))
            if ri:
                return self._pw_cmp_scalar(r, swap[k], scal(l))
            return Number({"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal,
                           "==": np.equal, "!=": np.not_equal}[k](l.v, r.v).astype(np.float64))
        if core in ("+", "-", "*", "/"):
            if li and ri:
                return self._pw_arith(l, core, r)
            if li:
                return self._pw_arith_scalar(l, core, scal(r))
            if ri:
                rop = {"+": "+", "*": "*", "-": "r-", "/": "r/"}[core]
                return self._pw_arith_scalar(r, rop, scal(l))
            return Number({"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide}[core](l.v, r.v))
        raise NameError(f"ImgQL: unknown infix operator {op!r}")


def _closed_over_loads(e, loaded) -> bool:
    if isinstance(e, Var):
        return e.name in loaded
    if isinstance(e, Call):
        return e.fn in ("intensity", "red", "green", "blue", "alpha") and all(_closed_over_loads(a, loaded) for a in e.args)
    return False


def _find_hoistable(e, loaded, fns, found) -> None:
    if isinstance(e, Call):
        if e.fn in fns and len(e.args) == 1 and _closed_over_loads(e.args[0], loaded) and e not in found:
            found.append(e)
        for a in e.args:
            _find_hoistable(a, loaded, fns, found)


def run_spec(ctx: Context, src: str, images: Dict[str, Image], conn: int = 26, fuse: bool = True,
             gc: bool = False) -> Interpreter:
    """Run an ImgQL spec whose `load` statements name entries of ``images``."""
    it = Interpreter(ctx, loader=lambda name: images[name], conn=conn, fuse=fuse, gc=gc)
    return it.run(src)


def run_file(ctx: Context, path: str, conn: int = 26) -> Interpreter:
    """Run an .imgql file: `load` reads NIfTI / PNG files (relative to the spec's directory) into
    HBM, `save` writes the result next to it (boolean -> u8 mask, numbers -> f32)."""
    import os

    from . import io as vio

    base = os.path.dirname(os.path.abspath(path))
    spacing = {}

    def loader(name):
        full = os.path.join(base, name)
        if full.lower().endswith((".nii", ".nii.gz")):
            # a 16-bit file travels as 16 bits: half the PCIe bytes, the same f32 image, and the operators
            # that rank or bin its values take their counting paths (vx_upload_scaled keeps the codes)
            raw, slope, inter, sp = vio.read_nifti_raw(full)
            if raw.dtype in (np.int16, np.uint16):
                spacing["sp"] = sp
                return Loaded({"intensity": ctx.upload_scaled(raw, slope, inter, spacing=sp)})
        chans = vio.load_image(full)
        out = {}
        for k, (arr, sp) in chans.items():
            out[k] = ctx.upload(arr, spacing=sp)
            spacing["sp"] = sp
        return Loaded(out)

    def saver(name, val):
        if isinstance(val, Image):
            vio.save_image(os.path.join(base, name), val.numpy()[0], spacing.get("sp", (1.0, 1.0, 1.0)))

    it = Interpreter(ctx, loader=loader, conn=conn)
    it.saver = saver
    with open(path) as f:
        return it.run(f.read())


def run_texture(ctx: Context, images: Dict[str, Image], conn: int = 26) -> Dict[str, object]:
    """The multi-modal texture-similarity spec (BASELINE config 4) on (batched) flair/t1/t1c/t2."""
    it = run_spec(ctx, TEXTURE_SPEC, images, conn)
    out = dict(it.globals)
    out["_tasks"] = it.n_tasks
    out["_printed"] = it.printed
    return out


def run_gtv(ctx: Context, flair: Image, conn: int = 26) -> Dict[str, object]:
    """The BraTS GTV segmentation spec of Greta.pdf p.48 on a (batched) FLAIR image."""
    it = run_spec(ctx, GTV_SPEC, {"flair": flair}, conn)
    out = dict(it.globals)
    out["_tasks"] = it.n_tasks
    out["_printed"] = it.printed
    return out
