"""Slab decomposition of ONE oversized volume across ranks.

A volume [nz][ny][nx] is split along z into contiguous slabs, one per rank (one process per
GPU, ``torch.distributed``: NCCL over NVLink on the GPUs, gloo in the CPU tests).  Every
operator has at most one exchange step:

    pointwise ............ none
    volume / min / max ... one scalar all-reduce
    near / interior ...... 1-plane halo from each z-neighbour: either send/recv + an extended copy
                           of the slab, or (``SlabOps.enable_peer_halo``) FUSED: every rank
                           publishes its two boundary planes in a CUDA-IPC mailbox and the
                           dilation kernel reads the neighbours' planes straight out of peer
                           memory over NVLink -- no extended slab, no crop, one barrier
    distlt / distgeq ..... ceil(r / sz)-plane halo (the interval test only looks r far)
    crossCorrelation ..... ceil(rho / sz)-plane halo of `a` + all-reduce of the region histogram
    through .............. local union-find per slab, exchange of the boundary label planes,
                           all-gather of the cross-slab label equivalences and of the labels
                           already reached by a seed, union-find over that (small) boundary graph
                           replicated on every rank, then one more local pass seeded with the
                           labels that became reachable through a neighbour

The local work is done by an *engine* (``CudaEngine`` = libvoxcuda through the C ABI; the CPU
tests plug in an oracle-backed engine to exercise exactly this host logic with gloo).  Planes
travel as torch tensors: zero-copy views of the HBM-resident images on the GPU path.

    edt (f32 map) ........ x and y passes local, then an all-to-all re-partition of the columns
                           (each rank gets a range of rows of EVERY plane), z pass + sqrt on
                           complete columns, all-to-all back

``percentiles`` (distributed sample sort of the masked voxels) and ``maxvol`` (boundary classes
summed over the ranks) are decomposed as well, see their docstrings.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def slab_bounds(nz: int, world: int) -> List[Tuple[int, int]]:
    """[z0, z1) of every rank: as even as possible, earlier ranks get the extra planes."""
    if world < 1 or nz < world:
        raise ValueError(f"cannot split {nz} planes over {world} ranks")
    base, extra = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


@dataclass
class Slab:
    """One rank's part of a distributed volume."""
    data: object                 # engine array: planes [z0, z1) as [1][nz_local][ny][nx]
    z0: int
    z1: int
    nz_global: int
    spacing: Tuple[float, float, float] = (1.0, 1.0, 1.0)

    is_image = True              # the ImgQL driver treats a slab as an image value
    nb = 1

    def like(self, data) -> "Slab":
        return Slab(data, self.z0, self.z1, self.nz_global, self.spacing)


class CudaEngine:
    """Local operations of one rank on libvoxcuda images (voxlogica_b200.ops.Context)."""

    def __init__(self, ctx):
        self.ctx = ctx

    # geometry / movement
    def nz(self, x):
        return x.shape[1]

    def crop_z(self, x, z0, n):
        return self.ctx.crop_z(x, z0, n)

    def concat_z(self, parts):
        return self.ctx.concat_z(list(parts))

    def to_tensor(self, x):
        t = self.ctx.as_tensor(x)
        self.ctx.sync()                      # producer finished: safe on torch's streams
        return t[0]

    def from_tensor(self, t, like):
        return self.ctx.from_tensor(t[None], spacing=like.spacing)

    def zeros_u8(self, like):
        return self.ctx.ff(like)

    # Operators whose engine signature is the Context's own are handed through by name: the local
    # half of a slab operator is the whole-volume operator on the slab.
    _SAME_AS_CONTEXT = frozenset((
        "not_", "and_", "or_", "cmp_scalar", "cmp", "arith_scalar", "arith", "mask", "to_f32",
        "near", "interior", "through", "components", "maxvol", "border", "percentiles",
        # the stages of `through` on a kept union-find state: one local labelling instead of three
        "ccl_create", "ccl_seed", "ccl_flag", "ccl_select", "ccl_destroy", "ccl_count", "ccl_sizes", "ccl_flag_size",
        # building blocks of the distributed sample sort and of the slab EDT / crossCorrelation
        "pc_pairs", "pc_sort", "pc_rank", "pc_scatter", "edt_sq_xy", "edt_z_from_sq", "cross_correlation_h2"))

    def __getattr__(self, name):
        if name in CudaEngine._SAME_AS_CONTEXT:
            return getattr(self.ctx, name)
        raise AttributeError(name)

    def ccl_plane(self, st, z):
        lab, hit = self.ctx.ccl_plane(st, z)
        tl, th = self.ctx.as_tensor(lab), self.ctx.as_tensor(hit)
        self.ctx.sync()
        return tl[0, 0], th[0, 0]

    def dist(self, op, r, a):
        return {"<": self.ctx.distlt, "<=": self.ctx.distleq, ">=": self.ctx.distgeq, ">": self.ctx.distgt}[op](r, a)

    def region_histogram(self, b, region, m, M, k): return self.ctx.region_histogram(b, region, m, M, k)[0]
    def volume(self, a): return float(self.ctx.volume(a)[0])
    def min(self, a): return float(self.ctx.min(a)[0])
    def max(self, a): return float(self.ctx.max(a)[0])


class SlabComm:
    """The collective side: a torch.distributed process group (NCCL or gloo)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_planes(self, to_lower, to_upper):
        """send `to_lower` to rank-1 and `to_upper` to rank+1 (tensors or None at the ends);
        returns (from_lower, from_upper), None where there is no neighbour."""
        import torch
        dist = self.dist
        ops, from_lower, from_upper = [], None, None
        if self.rank > 0:
            from_lower = torch.empty_like(to_lower)
            ops.append(dist.P2POp(dist.isend, to_lower.contiguous(), self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, from_lower, self.rank - 1, self.group))
        if self.rank + 1 < self.world:
            from_upper = torch.empty_like(to_upper)
            ops.append(dist.P2POp(dist.isend, to_upper.contiguous(), self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, from_upper, self.rank + 1, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            if to_lower.is_cuda:
                torch.cuda.current_stream(to_lower.device).synchronize()
        return from_lower, from_upper

    def barrier(self):
        self.dist.barrier(group=self.group)

    def exchange(self, send, recv_shapes):
        """personalised all-to-all built from point-to-point operations (works with NCCL and gloo)"""
        import torch
        dist = self.dist
        recv = [torch.empty(shape, dtype=send[0].dtype, device=send[0].device) for shape in recv_shapes]
        ops = []
        for r in range(self.world):
            if r == self.rank:
                recv[r].copy_(send[r])
                continue
            # empty messages are skipped on both sides (sender and receiver know the sizes)
            if send[r].numel():
                ops.append(dist.P2POp(dist.isend, send[r].contiguous(), r, self.group))
            if recv[r].numel():
                ops.append(dist.P2POp(dist.irecv, recv[r], r, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            if send[0].is_cuda:
                torch.cuda.current_stream(send[0].device).synchronize()
        return recv

    def share_mailbox(self, ctx, ptr: int, handle: bytes):
        """exchange CUDA IPC handles; returns the device pointers of the (lower, upper) neighbours'
        mailboxes mapped into this process (None at the ends)"""
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle, group=self.group)
        lower = ctx.ipc_open(handles[self.rank - 1]) if self.rank > 0 else None
        upper = ctx.ipc_open(handles[self.rank + 1]) if self.rank + 1 < self.world else None
        return lower, upper

    def all_gather_rows(self, t):
        """all-gather of [n_i, c] int64 tensors with different n_i -> one [sum n_i, c] tensor"""
        import torch
        dist = self.dist
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = torch.zeros(self.world, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(sizes, n, group=self.group)
        sizes = sizes.tolist()                       # one host sync
        cap = max(max(sizes), 1)
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        buf = torch.empty((self.world * cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=self.group)
        return torch.cat([buf[r * cap: r * cap + s] for r, s in enumerate(sizes)], dim=0)

    def all_reduce(self, values: Sequence[float], op: str, device) -> np.ndarray:
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device=device)
        red = {"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX}[op]
        self.dist.all_reduce(t, op=red, group=self.group)
        return t.cpu().numpy()

    def all_reduce_i64(self, arr: np.ndarray, device) -> np.ndarray:
        import torch
        t = torch.as_tensor(np.asarray(arr, dtype=np.int64), device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class PeerHalo:
    """Double-buffered boundary-plane mailboxes in peer-visible device memory.

    Layout of one rank's mailbox: [parity 0|1][side 0 = my bottom plane, 1 = my top plane][ny*nx].
    Protocol per operator: write own planes into mailbox[parity] on the library stream,
    synchronise the stream, barrier, launch the fused kernel with pointers INTO THE NEIGHBOURS'
    mailboxes, flip parity.  The barrier of operator k+1 is passed only after every rank has
    drained its stream, i.e. after all kernels of operator k have finished reading, so
    overwriting parity k's slots in operator k+2 is safe without a second barrier."""

    def __init__(self, ctx, comm, plane_bytes: int):
        self.ctx, self.comm, self.plane = ctx, comm, int(plane_bytes)
        self.ptr, handle = ctx.ipc_alloc(4 * self.plane)
        self.lower, self.upper = comm.share_mailbox(ctx, self.ptr, handle)
        self.parity = 0

    def slot(self, base: int, parity: int, side: int) -> int:
        return base + (2 * parity + side) * self.plane


class SlabOps:
    """ImgQL operators on Slab values.  Same names and argument order as ops.Context."""

    def __init__(self, engine, comm: SlabComm):
        self.e = engine
        self.c = comm
        self.peer: Optional[PeerHalo] = None
        self.use_ccl_state = True     # `through` on a kept union-find state when the engine has one
        self.trace: Optional[dict] = None   # when a dict: seconds per phase of `through` (device-synchronised)

    def _tick(self, name: str, t0: float) -> float:
        """phase timing for the bench: synchronise the device, add the elapsed seconds to trace[name]"""
        if self.trace is None:
            return t0
        import time
        if hasattr(self.e, "ctx"):
            self.e.ctx.sync()
        t1 = time.perf_counter()
        self.trace[name] = self.trace.get(name, 0.0) + (t1 - t0)
        return t1

    def enable_peer_halo(self, ny: int, nx: int) -> None:
        """collective: allocate and exchange the IPC mailboxes for [ny][nx] u8 boundary planes"""
        self.peer = PeerHalo(self.e.ctx, self.c, ny * nx)

    def _near_peer(self, a: Slab, conn: int, erode: bool) -> Slab:
        ph, ctx = self.peer, self.e.ctx
        nzl = self.e.nz(a.data)
        par = ph.parity
        ctx.copy_planes_to(a.data, 0, 1, ph.slot(ph.ptr, par, 0))
        ctx.copy_planes_to(a.data, nzl - 1, 1, ph.slot(ph.ptr, par, 1))
        ctx.sync()
        self.c.barrier()
        lo = ph.slot(ph.lower, par, 1) if ph.lower else 0     # the lower neighbour's TOP plane
        hi = ph.slot(ph.upper, par, 0) if ph.upper else 0     # the upper neighbour's BOTTOM plane
        out = ctx.near_halo(a.data, conn, lo, hi, erode)
        ph.parity ^= 1
        return a.like(out)

    # ------------------------------------------------------------------ distribution
    def scatter_from_global(self, upload, vol: np.ndarray, spacing=(1.0, 1.0, 1.0)) -> Slab:
        """test/harness helper: every rank holds the global array and keeps its slab"""
        z0, z1 = slab_bounds(vol.shape[0], self.c.world)[self.c.rank]
        return Slab(upload(np.ascontiguousarray(vol[z0:z1]), spacing), z0, z1, vol.shape[0], tuple(spacing))

    def _device_of(self, x: Slab):
        return self.e.to_tensor(self.e.crop_z(x.data, 0, 1)).device

    # ------------------------------------------------------------------ halo machinery
    def _with_halo(self, x: Slab, h: int):
        """-> (extended array, planes added below, planes added above)"""
        if h <= 0 or self.c.world == 1:
            return x.data, 0, 0
        nzl = self.e.nz(x.data)
        # a halo comes from the immediate neighbour only; decide from the SMALLEST slab so that every
        # rank raises together (deciding from the local slab left the other ranks inside the exchange)
        smallest = min(z1 - z0 for z0, z1 in slab_bounds(x.nz_global, self.c.world))
        if h > smallest:
            raise ValueError(f"halo of {h} planes exceeds the smallest slab ({smallest} planes): use fewer ranks")
        lo = self.e.to_tensor(self.e.crop_z(x.data, 0, h))
        hi = self.e.to_tensor(self.e.crop_z(x.data, nzl - h, h))
        from_lower, from_upper = self.c.exchange_planes(lo, hi)
        parts, below, above = [], 0, 0
        if from_lower is not None:
            parts.append(self.e.from_tensor(from_lower, x))
            below = h
        parts.append(x.data)
        if from_upper is not None:
            parts.append(self.e.from_tensor(from_upper, x))
            above = h
        return (self.e.concat_z(parts) if len(parts) > 1 else x.data), below, above

    def _halo_op(self, x: Slab, h: int, fn) -> Slab:
        ext, below, above = self._with_halo(x, h)
        out = fn(ext)
        if below or above:
            out = self.e.crop_z(out, below, self.e.nz(x.data))
        return x.like(out)

    # ------------------------------------------------------------------ local operators
    def not_(self, a: Slab) -> Slab: return a.like(self.e.not_(a.data))
    def and_(self, a: Slab, b: Slab) -> Slab: return a.like(self.e.and_(a.data, b.data))
    def or_(self, a: Slab, b: Slab) -> Slab: return a.like(self.e.or_(a.data, b.data))
    def cmp_scalar(self, a: Slab, op: str, s: float) -> Slab: return a.like(self.e.cmp_scalar(a.data, op, s))
    def cmp(self, a: Slab, op: str, b: Slab) -> Slab: return a.like(self.e.cmp(a.data, op, b.data))
    def arith_scalar(self, a: Slab, op: str, s: float) -> Slab: return a.like(self.e.arith_scalar(a.data, op, s))
    def arith(self, a: Slab, op: str, b: Slab) -> Slab: return a.like(self.e.arith(a.data, op, b.data))
    def mask(self, a: Slab, m: Slab, fill: float = 0.0) -> Slab: return a.like(self.e.mask(a.data, m.data, fill))
    def ff(self, like: Slab) -> Slab: return like.like(self.e.zeros_u8(like.data))
    def tt(self, like: Slab) -> Slab: return like.like(self.e.not_(self.e.zeros_u8(like.data)))

    # ------------------------------------------------------------------ reductions
    def volume(self, a: Slab) -> float:
        return float(self.c.all_reduce([self.e.volume(a.data)], "sum", self._device_of(a))[0])

    def min(self, a: Slab) -> float:
        return float(self.c.all_reduce([self.e.min(a.data)], "min", self._device_of(a))[0])

    def max(self, a: Slab) -> float:
        return float(self.c.all_reduce([self.e.max(a.data)], "max", self._device_of(a))[0])

    # ------------------------------------------------------------------ halo operators
    def near(self, a: Slab, conn: int = 26) -> Slab:
        if self.peer is not None and self.c.world > 1:
            return self._near_peer(a, conn, False)
        return self._halo_op(a, 1, lambda x: self.e.near(x, conn))

    def interior(self, a: Slab, conn: int = 26) -> Slab:
        if self.peer is not None and self.c.world > 1:
            return self._near_peer(a, conn, True)
        return self._halo_op(a, 1, lambda x: self.e.interior(x, conn))

    def _dist(self, op: str, r: float, a: Slab) -> Slab:
        h = 0 if not (r > 0) else int(math.ceil(float(r) / float(a.spacing[2])))
        h = min(h, a.nz_global)
        return self._halo_op(a, h, lambda x: self.e.dist(op, r, x))

    def distlt(self, r, a): return self._dist("<", r, a)
    def distleq(self, r, a): return self._dist("<=", r, a)
    def distgeq(self, r, a): return self._dist(">=", r, a)
    def distgt(self, r, a): return self._dist(">", r, a)

    def flt(self, r, a: Slab) -> Slab:
        return self.distlt(r, self.distgeq(r, self.not_(a)))

    def cross_correlation(self, rho: float, a: Slab, b: Slab, region: Slab, m: float, M: float, k: int) -> Slab:
        h2_local = np.asarray(self.e.region_histogram(b.data, region.data, m, M, k), dtype=np.int64)
        h2 = self.c.all_reduce_i64(h2_local, self._device_of(a)).astype(np.uint64)
        h = 0 if not (rho > 0) else min(int(math.ceil(float(rho) / float(a.spacing[2]))), a.nz_global)
        return self._halo_op(a, h, lambda x: self.e.cross_correlation_h2(rho, x, h2, m, M, k))

    # ------------------------------------------------------------------ reachability
    def through(self, a: Slab, b: Slab, conn: int = 26) -> Slab:
        """union of the connected components of b (global, across slabs) that contain a voxel of a"""
        import torch
        e, c = self.e, self.c
        if c.world == 1:
            return b.like(e.through(a.data, b.data, conn))
        nzl = e.nz(b.data)
        import time
        tk = time.perf_counter() if self.trace is not None else 0.0
        stateful = self.use_ccl_state and hasattr(e, "ccl_create")
        if stateful:
            st = e.ccl_create(b.data, conn)
            e.ccl_seed(st, a.data)
            lab_bot, hit_bot = e.ccl_plane(st, 0)
            lab_top, hit_top = e.ccl_plane(st, nzl - 1)
            lab_bot, lab_top = lab_bot.to(torch.int64) & 0xffffffff, lab_top.to(torch.int64) & 0xffffffff
            hit_bot, hit_top = hit_bot != 0, hit_top != 0
        else:
            r0 = e.through(a.data, b.data, conn)
            labels = e.components(b.data, conn)
            lab_bot = e.to_tensor(e.crop_z(labels, 0, 1))[0].to(torch.int64) & 0xffffffff
            lab_top = e.to_tensor(e.crop_z(labels, nzl - 1, 1))[0].to(torch.int64) & 0xffffffff
            hit_bot = e.to_tensor(e.crop_z(r0, 0, 1))[0] != 0
            hit_top = e.to_tensor(e.crop_z(r0, nzl - 1, 1))[0] != 0
        tag = (c.rank + 1) << 32                       # global id of a local label: (rank+1, label)
        gid_bot = torch.where(lab_bot > 0, lab_bot + tag, torch.zeros_like(lab_bot))
        gid_top = torch.where(lab_top > 0, lab_top + tag, torch.zeros_like(lab_top))
        tk = self._tick("local_ccl_and_planes", tk)
        # the upper neighbour needs my top plane, I need the lower neighbour's top plane
        below_top, _ = c.exchange_planes(gid_bot, gid_top)   # from_lower = what rank-1 sent "to_upper"
        tk = self._tick("exchange_label_planes", tk)
        pairs = torch.zeros((0, 2), dtype=torch.int64, device=gid_bot.device)
        if below_top is not None:
            pairs = boundary_pairs(below_top, gid_bot, conn)
        reached = torch.unique(torch.cat([gid_bot[hit_bot], gid_top[hit_top]]))
        tk = self._tick("boundary_pairs", tk)
        all_pairs = c.all_gather_rows(pairs).cpu().numpy()
        all_reached = c.all_gather_rows(reached[:, None]).cpu().numpy()[:, 0]
        tk = self._tick("all_gather_pairs_and_reached", tk)
        if self.trace is not None:
            self.trace["pairs_gathered"] = int(all_pairs.shape[0])
            self.trace["reached_gathered"] = int(all_reached.shape[0])
        newly = resolve_boundary_graph(all_pairs, all_reached)
        mine = newly[(newly >> 32) == c.rank + 1]
        tk = self._tick("resolve_boundary_graph_host", tk)
        if stateful:
            if mine.size:
                e.ccl_flag(st, (mine & 0xffffffff).astype(np.uint32))
            out = e.ccl_select(st)
            e.ccl_destroy(st)
            self._tick("flag_and_select", tk)
            return b.like(out)
        if mine.size == 0:
            return b.like(r0)
        mine_t = torch.as_tensor(mine, device=gid_bot.device)
        seed_bot = torch.isin(gid_bot, mine_t).to(torch.uint8)
        seed_top = torch.isin(gid_top, mine_t).to(torch.uint8)
        # extra seeds live on the two boundary planes only
        parts = []
        if nzl == 1:
            parts.append(e.from_tensor((seed_bot | seed_top)[None], b))
        else:
            parts.append(e.from_tensor(seed_bot[None], b))
            if nzl > 2:
                parts.append(e.crop_z(e.zeros_u8(b.data), 0, nzl - 2))
            parts.append(e.from_tensor(seed_top[None], b))
        extra = e.concat_z(parts) if len(parts) > 1 else parts[0]
        return b.like(e.through(e.or_(a.data, extra), b.data, conn))

    def touch(self, a: Slab, b: Slab, conn: int = 26) -> Slab:
        return self.and_(a, self.through(self.near(b, conn), a, conn))

    def grow(self, a: Slab, b: Slab, conn: int = 26) -> Slab:
        return self.or_(a, self.touch(b, a, conn))

    # ------------------------------------------------------------------ distance map
    def edt(self, a: Slab) -> Slab:
        """exact Euclidean distance map (f32): x and y passes on the local slab, then the columns
        are re-partitioned (rank r gets rows [y0_r, y1_r) of EVERY plane), the z pass and the
        square root run on complete columns, and the result travels back to the z-slabs."""
        import torch
        e, c = self.e, self.c
        if c.world == 1:
            sq = e.edt_sq_xy(a.data)
            t = e.to_tensor(sq)
            nz, ny, nx = t.shape
            bound = (a.spacing[0] * (nx - 1)) ** 2 + (a.spacing[1] * (ny - 1)) ** 2
            return a.like(e.edt_z_from_sq(sq, bound))
        sq = e.to_tensor(e.edt_sq_xy(a.data))                    # [nzl][ny][nx]
        nzl, ny, nx = sq.shape
        zb = slab_bounds(a.nz_global, c.world)
        if ny < c.world:
            raise ValueError(f"cannot re-partition {ny} rows over {c.world} ranks")
        yb = slab_bounds(ny, c.world)
        my0, my1 = yb[c.rank]
        # forward: my planes, every rank's rows  ->  every rank's planes, my rows
        send = [sq[:, y0:y1, :].contiguous() for (y0, y1) in yb]
        recv = c.exchange(send, [(z1 - z0, my1 - my0, nx) for (z0, z1) in zb])
        cols = torch.cat(recv, dim=0)                            # [nz][my rows][nx]
        bound = (a.spacing[0] * (nx - 1)) ** 2 + (a.spacing[1] * (ny - 1)) ** 2
        like = Slab(None, 0, a.nz_global, a.nz_global, a.spacing)
        d = e.to_tensor(e.edt_z_from_sq(e.from_tensor(cols, like), bound))
        # back: my rows of every plane -> my planes, all rows
        send = [d[z0:z1].contiguous() for (z0, z1) in zb]
        recv = c.exchange(send, [(nzl, y1 - y0, nx) for (y0, y1) in yb])
        return a.like(e.from_tensor(torch.cat(recv, dim=1), a))

    # ------------------------------------------------------------------ border and the GTV spec
    def border(self, like: Slab) -> Slab:
        """voxels on the outer faces of the GLOBAL volume: x / y faces on every plane, full planes
        only where the slab really ends the volume (a cut between two slabs is not a face)"""
        e, c = self.e, self.c
        if c.world == 1:
            return like.like(e.border(like.data))
        nzl = e.nz(like.data)
        one = e.crop_z(like.data, 0, 1)
        ring = e.border(one)                       # an axis of extent 1 has no faces: x / y faces only
        full = e.not_(e.zeros_u8(one))             # a real z face
        zfaces = like.nz_global > 1
        bottom = full if (zfaces and like.z0 == 0) else ring
        top = full if (zfaces and like.z1 == like.nz_global) else ring
        if nzl == 1:
            return like.like(bottom if (zfaces and like.z0 == 0) else top)
        parts = [bottom]
        if nzl > 2:
            parts.append(e.crop_z(e.border(like.data), 1, nzl - 2))   # interior planes of a local border are rings
        parts.append(top)
        return like.like(e.concat_z(parts))

    def gtv(self, flair: Slab, conn: int = 26) -> dict:
        """the GTV segmentation spec of Greta.pdf p.48 (imgql.GTV_SPEC) on a slab-decomposed volume:
        every operator of the spec is a distributed one here"""
        r = {}
        r["background"] = self.touch(self.cmp_scalar(flair, "<", 0.1), self.border(flair), conn)
        r["brain"] = self.not_(r["background"])
        r["normFlair"] = self.percentiles(flair, r["brain"])
        r["hyperIntense"] = self.flt(5.0, self.cmp_scalar(r["normFlair"], ">", 0.95))
        r["veryIntense"] = self.flt(2.0, self.cmp_scalar(r["normFlair"], ">", 0.86))
        r["growTum"] = self.grow(r["hyperIntense"], r["veryIntense"], conn)
        m, M = self.min(flair), self.max(flair)
        r["tumSim"] = self.cross_correlation(5.0, flair, flair, r["growTum"], m, M, 100)
        r["tumStatCC"] = self.flt(2.0, self.cmp_scalar(r["tumSim"], ">", 0.6))
        r["gtv"] = self.grow(r["growTum"], r["tumStatCC"], conn)
        r["gtvVolume"] = self.volume(r["gtv"])
        return r

    # ------------------------------------------------------------------ rank normalisation
    def percentiles(self, a: Slab, mask: Slab, correction: float = 0.0) -> Slab:
        """rank of every masked voxel among ALL masked voxels of the distributed volume -- an exact
        distributed sort by sampling: every rank compacts and sorts its (key, voxel) pairs, the
        key space is cut at W-1 splitters taken from a sample of every rank's sorted keys, each
        rank receives the pairs of its key range (a personalised all-to-all of contiguous slices),
        sorts them and ranks them against the global count (equal keys always meet on one rank,
        so ties are resolved there), and the ranks travel back the same way."""
        import torch
        e, c = self.e, self.c
        if c.world == 1:
            return a.like(e.percentiles(a.data, mask.data, correction))
        W = c.world
        pairs = e.pc_sort(e.pc_pairs(a.data, mask.data))              # int64: key << 32 | voxel index
        dev = pairs.device
        keys = (pairs >> 32) & 0xffffffff                              # unsigned key as a non-negative int64
        n = int(pairs.numel())
        # splitters: up to 1024 evenly spaced sorted keys per rank
        if n:
            idx = torch.linspace(0, n - 1, min(n, 1024), device=dev).round().to(torch.int64)
            sample = keys[idx]
        else:
            sample = torch.zeros(0, dtype=torch.int64, device=dev)
        allsamp = torch.sort(c.all_gather_rows(sample[:, None])[:, 0]).values
        if allsamp.numel():
            pos = (torch.arange(1, W, device=dev, dtype=torch.float64) * allsamp.numel() / W).floor().to(torch.int64)
            splitters = allsamp[pos.clamp(max=allsamp.numel() - 1)]
        else:
            splitters = torch.zeros(W - 1, dtype=torch.int64, device=dev)
        cuts = torch.searchsorted(keys, splitters, right=False).tolist() if n else [0] * (W - 1)
        bounds = [0] + cuts + [n]                                      # rank r gets keys in [splitter r-1, splitter r)
        send = [pairs[bounds[r]:bounds[r + 1]] for r in range(W)]
        sizes = c.all_gather_rows(torch.tensor([[bounds[r + 1] - bounds[r] for r in range(W)]],
                                               dtype=torch.int64, device=dev)).cpu().numpy()   # [source][destination]
        recv = c.exchange(send, [(int(sizes[s, c.rank]),) for s in range(W)])
        got = torch.cat(recv) if recv else torch.zeros(0, dtype=torch.int64, device=dev)
        m = int(got.numel())
        per_dest = sizes.sum(axis=0)
        less_before, n_total = int(per_dest[:c.rank].sum()), int(per_dest.sum())
        # payload := position in the receive order, so that the ranks can be sent back slice by slice
        tagged = (got & ~0xffffffff) | torch.arange(m, dtype=torch.int64, device=dev)
        if m and n_total:
            ranks_recv = e.pc_rank(e.pc_sort(tagged), less_before, n_total, correction)
        else:
            ranks_recv = torch.zeros(m, dtype=torch.float32, device=dev)
        offs = np.concatenate([[0], np.cumsum(sizes[:, c.rank])]).astype(np.int64)
        back = c.exchange([ranks_recv[int(offs[s]):int(offs[s + 1])] for s in range(W)],
                          [(bounds[r + 1] - bounds[r],) for r in range(W)])
        ranks_local = torch.cat(back) if back else torch.zeros(0, dtype=torch.float32, device=dev)
        return a.like(e.pc_scatter(pairs, ranks_local, a.data))

    # ------------------------------------------------------------------ largest component
    def maxvol(self, a: Slab, conn: int = 26) -> Slab:
        """the connected component(s) of a (global, across slabs) with the largest voxel count.
        Every rank labels and counts its slab once (kept union-find state); the labels of the two
        boundary planes with their local counts are all-gathered together with the cross-slab
        label equivalences; the classes of the (small) boundary graph are resolved on every rank
        and their counts summed.  The global maximum is the larger of the largest local count
        (an all-reduce) and the largest class sum; winners are the local components of exactly
        that size plus the members of the winning classes."""
        import torch
        e, c = self.e, self.c
        if c.world == 1:
            return a.like(e.maxvol(a.data, conn))
        if not hasattr(e, "ccl_count"):
            raise NotImplementedError("maxvol on slabs needs an engine with a kept union-find state")
        nzl = e.nz(a.data)
        st = e.ccl_create(a.data, conn)
        local_max = int(e.ccl_count(st))
        lab_bot, _ = e.ccl_plane(st, 0)
        lab_top, _ = e.ccl_plane(st, nzl - 1)
        lab_bot, lab_top = lab_bot.to(torch.int64) & 0xffffffff, lab_top.to(torch.int64) & 0xffffffff
        tag = (c.rank + 1) << 32
        gid_bot = torch.where(lab_bot > 0, lab_bot + tag, torch.zeros_like(lab_bot))
        gid_top = torch.where(lab_top > 0, lab_top + tag, torch.zeros_like(lab_top))
        below_top, _ = c.exchange_planes(gid_bot, gid_top)
        pairs = torch.zeros((0, 2), dtype=torch.int64, device=gid_bot.device)
        if below_top is not None:
            pairs = boundary_pairs(below_top, gid_bot, conn)
        mine = torch.unique(torch.cat([lab_bot[lab_bot > 0], lab_top[lab_top > 0]]))
        sizes = np.asarray(e.ccl_sizes(st, mine.cpu().numpy().astype(np.uint32)), dtype=np.int64)
        rows = torch.stack([mine + tag, torch.as_tensor(sizes, device=mine.device)], dim=1) if mine.numel() else \
            torch.zeros((0, 2), dtype=torch.int64, device=gid_bot.device)
        all_pairs = c.all_gather_rows(pairs).cpu().numpy()
        all_sizes = c.all_gather_rows(rows).cpu().numpy()
        ids, totals = boundary_class_sizes(all_pairs, all_sizes)
        gmax = int(c.all_reduce([float(local_max)], "max", gid_bot.device)[0])
        if totals.size:
            gmax = max(gmax, int(totals.max()))
        if gmax > 0:
            if gmax <= 0xffffffff:
                e.ccl_flag_size(st, gmax)        # components that are the largest on their own
            win = ids[(totals == gmax) & ((ids >> 32) == c.rank + 1)]
            if win.size:
                e.ccl_flag(st, (win & 0xffffffff).astype(np.uint32))
        out = e.ccl_select(st)
        e.ccl_destroy(st)
        return a.like(out)



def boundary_pairs(lower, upper, conn: int):
    """unique (id in the plane below, id in the plane above) pairs of adjacent foreground voxels.
    lower/upper: [ny][nx] int64 ids (tag << 32 | label), 0 = background.  6-adjacency: same
    (y, x); 26: 3x3.  All ids of one plane carry the same tag, so a pair is packed into ONE int64
    (label_below << 32 | label_above) and de-duplicated with a single 1-D unique."""
    import torch
    ny, nx = lower.shape
    mask32 = 0xffffffff
    tag_l = int(lower.max().item()) >> 32 << 32 if lower.numel() else 0
    tag_u = int(upper.max().item()) >> 32 << 32 if upper.numel() else 0
    lo32, up32 = lower & mask32, upper & mask32
    out = []
    shifts = [(0, 0)] if conn == 6 else [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    for dy, dx in shifts:
        ys, ye = max(0, -dy), min(ny, ny - dy)
        xs, xe = max(0, -dx), min(nx, nx - dx)
        if ys >= ye or xs >= xe:
            continue
        u = up32[ys:ye, xs:xe]
        l = lo32[ys + dy:ye + dy, xs + dx:xe + dx]
        m = (u > 0) & (l > 0)
        out.append(((l << 32) | u)[m])
    packed = torch.unique(torch.cat(out)) if out else torch.zeros(0, dtype=torch.int64, device=lower.device)
    if packed.numel() == 0:
        return torch.zeros((0, 2), dtype=torch.int64, device=lower.device)
    # (packed >> 32) of a value with bit 63 set would sign-extend: labels are < 2^31, so it cannot
    return torch.stack([(packed >> 32) + tag_l, (packed & mask32) + tag_u], dim=1)


def boundary_class_sizes(pairs: np.ndarray, sizes: np.ndarray):
    """pairs: [m, 2] ids joined across a slab boundary; sizes: [n, 2] rows (id, local voxel count)
    of EVERY boundary component.  Returns (ids, total): for each id the voxel count of its class
    (the global component it belongs to) -- int64, a component can exceed 2^32 voxels."""
    if sizes.size == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    ids = np.unique(np.concatenate([sizes[:, 0], pairs.ravel()]).astype(np.int64))
    n = ids.size
    own = np.zeros(n, dtype=np.int64)
    own[np.searchsorted(ids, sizes[:, 0])] = sizes[:, 1]
    if pairs.size:
        e0, e1 = np.searchsorted(ids, pairs[:, 0]), np.searchsorted(ids, pairs[:, 1])
        graph = coo_matrix((np.ones(e0.size, dtype=np.int8), (e0, e1)), shape=(n, n))
        ncomp, comp = connected_components(graph, directed=False)
    else:
        ncomp, comp = n, np.arange(n)
    total = np.bincount(comp, weights=own.astype(np.float64), minlength=ncomp)   # exact below 2^53
    return ids, np.rint(total[comp]).astype(np.int64)


def resolve_boundary_graph(pairs: np.ndarray, reached: np.ndarray) -> np.ndarray:
    """Connected components of the boundary-label graph (replicated on every rank).  Returns the
    ids that are NOT in `reached` but share a class with an id that is: those components become
    reachable through a neighbouring slab."""
    if pairs.size == 0:
        return np.zeros(0, dtype=np.int64)
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    ids, inv = np.unique(pairs.ravel(), return_inverse=True)
    edges = inv.reshape(-1, 2)
    n = ids.size
    graph = coo_matrix((np.ones(edges.shape[0], dtype=np.int8), (edges[:, 0], edges[:, 1])), shape=(n, n))
    ncomp, comp = connected_components(graph, directed=False)
    is_reached = np.isin(ids, reached)
    comp_reached = np.zeros(ncomp, dtype=bool)
    comp_reached[comp[is_reached]] = True
    return ids[comp_reached[comp] & ~is_reached].astype(np.int64)


class SlabContext:
    """The operator table of ops.Context on Slab values, for the ImgQL driver: with it an .imgql
    spec runs unchanged on a volume that is spread over the ranks of a process group
    (`run_spec_on_slabs`).  Scalars are per volume as in ops.Context: arrays of one element."""

    def __init__(self, ops: "SlabOps"):
        self.o = ops

    def __getattr__(self, name):           # not_ and_ or_ cmp cmp_scalar arith arith_scalar mask near interior
        return getattr(self.o, name)       # through maxvol dist* edt percentiles border tt ff: same signatures

    def volume(self, a): return np.array([self.o.volume(a)])
    def min(self, a, mask=None): return np.array([self.o.min(a)])
    def max(self, a, mask=None): return np.array([self.o.max(a)])
    def minmax(self, a, mask=None): return np.array([self.o.min(a)]), np.array([self.o.max(a)])

    def constant(self, like, value):
        zero = like.like(self.o.e.to_f32(self.o.e.zeros_u8(like.data)))
        return self.o.arith_scalar(zero, "+", float(value))

    def cross_correlation(self, rho, a, b, region, m, M, k):
        return self.o.cross_correlation(rho, a, b, region, float(np.ravel(m)[0]), float(np.ravel(M)[0]), int(k))


def run_spec_on_slabs(ops: "SlabOps", src: str, images: dict, conn: int = 26):
    """Run an ImgQL spec whose `load` statements name Slab values of `images` (every rank calls
    this with its own slabs).  Pointwise operators are applied one by one (no fusion across the
    adapter); returns the interpreter (globals / saved / printed as in imgql.run_spec)."""
    from .imgql import Interpreter
    it = Interpreter(SlabContext(ops), loader=lambda name: images[name], conn=conn, fuse=False)
    return it.run(src)
