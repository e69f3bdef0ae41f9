"""Oracle-backed local engine for the CPU (gloo) tests of voxlogica_b200/slab.py.  TEST
INFRASTRUCTURE: it lets world_size-2 gloo runs exercise the slab host logic (partition, halo
exchange, boundary-label merge) without a GPU.  Arrays are numpy [1][nz][ny][nx]."""
import numpy as np
import torch

import oracle as orc


class OracleEngine:
    def __init__(self, spacing=(1.0, 1.0, 1.0)):
        self.sp = tuple(spacing)

    def nz(self, x): return x.shape[1]
    def crop_z(self, x, z0, n): return np.ascontiguousarray(x[:, z0:z0 + n])
    def concat_z(self, parts): return np.ascontiguousarray(np.concatenate(list(parts), axis=1))

    def to_tensor(self, x):
        a = x[0]
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        return torch.from_numpy(np.ascontiguousarray(a))

    def from_tensor(self, t, like):
        a = t.numpy()
        if a.dtype == np.int32:
            a = a.view(np.uint32)
        return np.ascontiguousarray(a[None])

    def zeros_u8(self, like): return np.zeros(like.shape, np.uint8)
    def not_(self, a): return orc.not_(a[0])[None]
    def and_(self, a, b): return orc.and_(a[0], b[0])[None]
    def or_(self, a, b): return orc.or_(a[0], b[0])[None]
    def cmp_scalar(self, a, op, s): return orc.cmp_scalar(a[0], op, s)[None]
    def cmp(self, a, op, b): return orc.cmp(a[0], op, b[0])[None]
    def arith_scalar(self, a, op, s): return orc.arith_scalar(a[0], op, s)[None]
    def arith(self, a, op, b): return orc.arith(a[0], op, b[0])[None]
    def mask(self, a, m, fill=0.0): return orc.mask(a[0], m[0], fill)[None]
    def to_f32(self, a): return orc.to_f32(a[0])[None]
    def near(self, a, conn): return orc.near(a[0], conn)[None]
    def interior(self, a, conn): return orc.interior(a[0], conn)[None]
    def through(self, a, b, conn): return orc.through(a[0], b[0], conn)[None]
    def components(self, a, conn): return orc.components(a[0], conn).astype(np.uint32)[None]
    def dist(self, op, r, a): return orc.dist_cmp(a[0], op, r, self.sp)[None]
    def maxvol(self, a, conn): return orc.maxvol(a[0], conn)[None]
    def border(self, like): return orc.border(like[0])[None]
    def percentiles(self, a, mask, correction): return orc.percentiles(a[0], mask[0], correction)[None]

    # building blocks of the distributed percentiles (vx_pc_* of the C ABI) in numpy
    @staticmethod
    def _sortable(v):
        u = (v.astype(np.float32) + np.float32(0.0)).view(np.uint32).astype(np.uint64)
        return np.where(u & 0x80000000, (~u) & 0xffffffff, u | 0x80000000).astype(np.uint64)

    def pc_pairs(self, a, mask):
        idx = np.flatnonzero(mask[0].ravel() != 0).astype(np.uint64)
        keys = self._sortable(a[0].ravel()[idx.astype(np.int64)])
        return torch.from_numpy(((keys << np.uint64(32)) | idx).view(np.int64))

    def pc_sort(self, pairs):
        p = pairs.numpy().view(np.uint64)
        order = np.argsort(p >> np.uint64(32), kind="stable")
        return torch.from_numpy(np.ascontiguousarray(p[order]).view(np.int64))

    def pc_rank(self, sorted_pairs, less_before, n_total, correction):
        p = sorted_pairs.numpy().view(np.uint64)
        keys, payload = p >> np.uint64(32), (p & np.uint64(0xffffffff)).astype(np.int64)
        less = np.searchsorted(keys, keys, side="left").astype(np.float64)
        equal = np.searchsorted(keys, keys, side="right").astype(np.float64) - less
        out = np.zeros(p.size, np.float32)
        out[payload] = ((float(less_before) + less + correction * equal) / float(n_total)).astype(np.float32)
        return torch.from_numpy(out)

    def pc_scatter(self, pairs, values, like):
        out = np.zeros(like[0].size, np.float32)
        out[(pairs.numpy().view(np.uint64) & np.uint64(0xffffffff)).astype(np.int64)] = values.numpy()
        return out.reshape(like.shape)

    # kept union-find state (vx_ccl_* of the C ABI), restated on oracle labels
    class _State:
        def __init__(self, lab):
            self.lab = lab                                    # uint32 [nz][ny][nx], 0 = background
            self.flag = np.zeros(int(lab.max()) + 1, bool)    # per label: selected / reached
            self.cnt = None

    def ccl_create(self, b, conn): return self._State(orc.components(b[0], conn).astype(np.uint32))

    def ccl_seed(self, st, a):
        st.flag[np.unique(st.lab[a[0] != 0])] = True
        st.flag[0] = False

    def ccl_plane(self, st, z):
        lab = st.lab[z]
        return (torch.from_numpy(np.ascontiguousarray(lab).view(np.int32)),
                torch.from_numpy(st.flag[lab].astype(np.uint8)))

    def ccl_flag(self, st, labels): st.flag[np.asarray(labels, dtype=np.int64)] = True

    def ccl_count(self, st):
        st.cnt = np.bincount(st.lab.ravel(), minlength=st.flag.size)
        st.cnt[0] = 0
        return int(st.cnt.max())

    def ccl_sizes(self, st, labels): return st.cnt[np.asarray(labels, dtype=np.int64)].astype(np.uint32)

    def ccl_flag_size(self, st, size):
        st.flag |= st.cnt == size
        st.flag[0] = False

    def ccl_select(self, st): return st.flag[st.lab].astype(np.uint8)[None]
    def ccl_destroy(self, st): pass

    # the two halves of the distance transform, restated with the arithmetic contract of
    # oracle/vx_oracle.c (binary32, one rounding per step, minimum over explicit offsets)
    def edt_sq_xy(self, a):
        v = a[0] != 0
        nz, ny, nx = v.shape
        sx, sy = np.float32(self.sp[0]), np.float32(self.sp[1])
        out = np.full(v.shape, np.inf, np.float32)
        dx2 = ((np.arange(nx, dtype=np.float32)[:, None] - np.arange(nx, dtype=np.float32)[None, :]) * sx) ** 2
        for z in range(nz):
            g1 = np.where(v[z][:, None, :], dx2[None].astype(np.float32), np.float32(np.inf)).min(axis=2)  # [ny][nx]
            for y in range(ny):
                dy2 = ((np.arange(ny, dtype=np.float32) - np.float32(y)) * sy) ** 2
                out[z, y] = (g1 + dy2[:, None].astype(np.float32)).astype(np.float32).min(axis=0)
        return out[None]

    def edt_z_from_sq(self, sq, max_in):
        g = sq[0]
        nz = g.shape[0]
        sz = np.float32(self.sp[2])
        out = np.empty_like(g)
        for z in range(nz):
            dz2 = ((np.arange(nz, dtype=np.float32) - np.float32(z)) * sz) ** 2
            out[z] = np.sqrt((g + dz2[:, None, None].astype(np.float32)).astype(np.float32).min(axis=0))
        return out[None]

    def region_histogram(self, b, region, m, M, k):
        v = b[0][region[0] != 0].astype(np.float64)
        v = v[(v >= m) & (v < M)]
        idx = np.minimum(((v - m) * k / (M - m)).astype(np.int64), k - 1)
        return np.bincount(idx, minlength=k).astype(np.uint64)

    def cross_correlation_h2(self, rho, a, h2, m, M, k):
        return orc.cross_correlation_h2(rho, a[0], np.asarray(h2, dtype=np.uint64), m, M, k, self.sp)[None]

    def volume(self, a): return float(orc.volume(a[0]))
    def min(self, a): return float(orc.min_(a[0]))
    def max(self, a): return float(orc.max_(a[0]))
