/*
 * vx_oracle.c -- CPU restatement of VoxLogicA's ImgQL operator semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / the timed CPU baseline.  The product
 * (libvoxcuda.so) never calls into it and has no CPU fallback.
 *
 * PARITY UNPINNED.  The mounted reference (/root/reference) is the repository's
 * `landing` branch: README.md and Greta.pdf, no F# sources, no tests, no golden
 * vectors; SimpleITK/ITK (where the reference's arithmetic lives) is neither mounted
 * nor installed (SURVEY.md sections 0 and 8c).  This file therefore restates
 *   - the formal semantics in the slide deck (cited per function as Greta.pdf p.N),
 *   - the published algorithms ITK implements (Maurer/Qi/Raghavan 2003 exact EDT
 *     semantics; standard connected components; flat structuring elements),
 *   - the defaults SURVEY.md section 7 lists for every choice the deck leaves open,
 * and is cross-checked against scipy.ndimage / OpenCV in tests/test_oracle.py.  It is
 * NOT verified against SimpleITK.
 *
 * Style: every function is written for obviousness, not speed, and deliberately uses
 * different algorithms from the CUDA library (BFS flood fill instead of union-find,
 * per-offset scans instead of bit-parallel dilation, full-ball histograms instead of
 * sliding ones, qsort instead of radix sort).  OpenMP is used over independent
 * planes/volumes only, so results do not depend on the thread count.
 *
 * Layout: one volume = [nz][ny][nx], x fastest; u8 booleans are 0/1.
 * Build:  see oracle/Makefile  (-ffp-contract=off: one rounding per operation).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OR_API __attribute__((visibility("default")))

/* comparison / arithmetic codes mirror include/voxcuda.h */
enum { OR_LT = 0, OR_LE, OR_GT, OR_GE, OR_EQ, OR_NE };
enum { OR_ADD = 0, OR_SUB, OR_MUL, OR_DIV, OR_RSUB, OR_RDIV, OR_MIN, OR_MAX };

static size_t nvox(int nx, int ny, int nz) { return (size_t)nx * ny * nz; }

/* ----------------------------------------------------------------------------
 * A1 boolean operators: Greta.pdf p.30 (not, and, true), p.35 (or, false). */
OR_API void or_not(const uint8_t* a, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = a[i] ? 0 : 1;
}
OR_API void or_and(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (a[i] && b[i]) ? 1 : 0;
}
OR_API void or_or(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (a[i] || b[i]) ? 1 : 0;
}
OR_API void or_xor(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = ((a[i] != 0) != (b[i] != 0)) ? 1 : 0;
}

/* A2 thresholds and arithmetic: Greta.pdf p.43 "Colour & thresholds"; p.48 uses
 * `<.` and `>.`.  IEEE binary32, one rounding per operator. */
static int cmp1(float x, float y, int op) {
    switch (op) {
        case OR_LT: return x < y;
        case OR_LE: return x <= y;
        case OR_GT: return x > y;
        case OR_GE: return x >= y;
        case OR_EQ: return x == y;
        default: return x != y;
    }
}
static float arith1(float x, float y, int op) {
    switch (op) {
        case OR_ADD: return x + y;
        case OR_SUB: return x - y;
        case OR_MUL: return x * y;
        case OR_DIV: return x / y;
        case OR_RSUB: return y - x;
        case OR_RDIV: return y / x;
        case OR_MIN: return fminf(x, y);
        default: return fmaxf(x, y);
    }
}
OR_API void or_cmp_scalar(const float* a, int op, float s, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)cmp1(a[i], s, op);
}
OR_API void or_cmp(const float* a, const float* b, int op, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)cmp1(a[i], b[i], op);
}
OR_API void or_arith_scalar(const float* a, int op, float s, float* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = arith1(a[i], s, op);
}
OR_API void or_arith(const float* a, const float* b, int op, float* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = arith1(a[i], b[i], op);
}
OR_API void or_to_f32(const uint8_t* a, float* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = a[i] ? 1.0f : 0.0f;
}
OR_API void or_to_bool(const float* a, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = a[i] != 0.0f;
}
OR_API void or_mask(const float* a, const uint8_t* m, float fill, float* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = m[i] ? a[i] : fill;
}

/* A10 border: the outer faces of the volume (Greta.pdf p.51 touch(..., border)).
 * An axis of extent 1 has no faces, so a 2D image (nz == 1) is not all border. */
OR_API void or_border(uint8_t* out, int nx, int ny, int nz) {
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                int b = (nx > 1 && (x == 0 || x == nx - 1)) || (ny > 1 && (y == 0 || y == ny - 1)) ||
                        (nz > 1 && (z == 0 || z == nz - 1));
                out[((size_t)z * ny + y) * nx + x] = (uint8_t)b;
            }
}

/* adjacency R of the closure space (Greta.pdf p.28 "regular grids with chosen
 * adjacency"): conn 6 = faces, 26 = faces+edges+corners; never the voxel itself. */
static int adjacent(int dx, int dy, int dz, int conn) {
    int nzero = (dx != 0) + (dy != 0) + (dz != 0);
    if (nzero == 0) return 0;
    return conn == 26 ? 1 : nzero == 1;
}

/* A4 near: closure C_R(A) = A u {x | exists a in A. (a,x) in R}  (Greta.pdf p.24, p.30).
 * Only voxels inside the image exist. */
OR_API void or_near(const uint8_t* a, uint8_t* out, int nx, int ny, int nz, int conn) {
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                int v = a[((size_t)z * ny + y) * nx + x] != 0;
                for (int dz = -1; dz <= 1 && !v; dz++)
                    for (int dy = -1; dy <= 1 && !v; dy++)
                        for (int dx = -1; dx <= 1 && !v; dx++) {
                            if (!adjacent(dx, dy, dz, conn)) continue;
                            int xx = x + dx, yy = y + dy, zz = z + dz;
                            if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                            if (a[((size_t)zz * ny + yy) * nx + xx]) v = 1;
                        }
                out[((size_t)z * ny + y) * nx + x] = (uint8_t)v;
            }
}

/* A4 interior: I(A) = !N(!A)  (Greta.pdf p.35), literally. */
OR_API void or_interior(const uint8_t* a, uint8_t* out, int nx, int ny, int nz, int conn) {
    size_t n = nvox(nx, ny, nz);
    uint8_t* t = (uint8_t*)malloc(n);
    or_not(a, t, n);
    or_near(t, out, nx, ny, nz, conn);
    or_not(out, out, n);
    free(t);
}

/* Connected components by breadth-first flood fill in raster order.
 * label = 1 + linear index of the component's first voxel (canonical), 0 = background.
 * Greta.pdf p.78 "Reachability via Connected Components". */
OR_API void or_components(const uint8_t* a, uint32_t* lab, int nx, int ny, int nz, int conn) {
    size_t n = nvox(nx, ny, nz);
    memset(lab, 0, n * sizeof(uint32_t));
    uint32_t* queue = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (size_t s = 0; s < n; s++) {
        if (!a[s] || lab[s]) continue;
        size_t head = 0, tail = 0;
        queue[tail++] = (uint32_t)s;
        lab[s] = (uint32_t)s + 1;
        while (head < tail) {
            size_t c = queue[head++];
            int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / ((size_t)nx * ny));
            for (int dz = -1; dz <= 1; dz++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!adjacent(dx, dy, dz, conn)) continue;
                        int xx = x + dx, yy = y + dy, zz = z + dz;
                        if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                        size_t q = ((size_t)zz * ny + yy) * nx + xx;
                        if (a[q] && !lab[q]) {
                            lab[q] = (uint32_t)s + 1;
                            queue[tail++] = (uint32_t)q;
                        }
                    }
        }
    }
    free(queue);
}

/* A3 through(a, b): the union of the connected components of b that contain a voxel of
 * a -- every voxel from which a voxel of a can be reached by a path that stays in b
 * (Greta.pdf p.33, p.36, p.38; p.78). */
OR_API void or_through(const uint8_t* a, const uint8_t* b, uint8_t* out, int nx, int ny, int nz, int conn) {
    size_t n = nvox(nx, ny, nz);
    memset(out, 0, n);
    uint32_t* queue = (uint32_t*)malloc(n * sizeof(uint32_t));
    size_t head = 0, tail = 0;
    for (size_t s = 0; s < n; s++)
        if (a[s] && b[s]) {
            out[s] = 1;
            queue[tail++] = (uint32_t)s;
        }
    while (head < tail) {
        size_t c = queue[head++];
        int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / ((size_t)nx * ny));
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (!adjacent(dx, dy, dz, conn)) continue;
                    int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                    size_t q = ((size_t)zz * ny + yy) * nx + xx;
                    if (b[q] && !out[q]) {
                        out[q] = 1;
                        queue[tail++] = (uint32_t)q;
                    }
                }
    }
    free(queue);
}

/* A8 maxvol: the component(s) of a with the largest voxel count (all of them on a
 * tie).  Greta.pdf p.43 "Volume, max, min (global)". */
OR_API void or_maxvol(const uint8_t* a, uint8_t* out, int nx, int ny, int nz, int conn) {
    size_t n = nvox(nx, ny, nz);
    uint32_t* lab = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t* cnt = (uint32_t*)calloc(n + 1, sizeof(uint32_t));
    or_components(a, lab, nx, ny, nz, conn);
    uint32_t best = 0;
    for (size_t i = 0; i < n; i++)
        if (lab[i]) {
            cnt[lab[i]]++;
            if (cnt[lab[i]] > best) best = cnt[lab[i]];
        }
    for (size_t i = 0; i < n; i++) out[i] = (lab[i] && cnt[lab[i]] == best) ? 1 : 0;
    free(lab);
    free(cnt);
}

/* ----------------------------------------------------------------------------
 * A5 exact Euclidean distance transform in physical units (Greta.pdf p.37 "distance
 * modality", p.43 D^I; the reference computes it with ITK's Maurer filter, whose
 * result is the exact distance to the nearest feature voxel).  Separable squared
 * distance in binary32, one rounding per operation:
 *     g1 = fl(fl(dx*sx)^2)            nearest feature in the row
 *     g2 = min_j fl(g1[j] + fl(fl((i-j)*sy)^2))
 *     g3 = min_j fl(g2[j] + fl(fl((i-j)*sz)^2)),   d = sqrtf(g3)
 * exact for unit spacing (integers < 2^24).  0 inside a; +inf when a is empty. */
static float sqstep(int d, float s) {
    float t = (float)d * s;
    return t * t;
}
static void edt_squared(const uint8_t* a, float* out, int nx, int ny, int nz, const float sp[3]) {
    size_t n = nvox(nx, ny, nz);
    float* g1 = (float*)malloc(n * sizeof(float));
    float* g2 = (float*)malloc(n * sizeof(float));
    /* x: nearest feature in the row, scanning every candidate */
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)nz * ny; r++) {
        const uint8_t* row = a + (size_t)r * nx;
        float* g = g1 + (size_t)r * nx;
        int any = 0;
        for (int x = 0; x < nx; x++) any |= row[x];
        for (int x = 0; x < nx; x++) {
            int best = -1;
            for (int k = 0; any && k < nx && best < 0; k++) {
                if (x - k >= 0 && row[x - k]) best = k;
                else if (x + k < nx && row[x + k]) best = k;
            }
            g[x] = best < 0 ? INFINITY : sqstep(best, sp[0]);
        }
    }
    /* y */
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; z++)
        for (int x = 0; x < nx; x++)
            for (int i = 0; i < ny; i++) {
                float best = INFINITY;
                for (int j = 0; j < ny; j++) {
                    float t = sqstep(i - j, sp[1]);
                    if (!(t < best) && j > i) break; /* (i-j)^2 grows with j > i */
                    float v = g1[((size_t)z * ny + j) * nx + x] + t;
                    if (v < best) best = v;
                }
                g2[((size_t)z * ny + i) * nx + x] = best;
            }
    /* z */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++)
            for (int i = 0; i < nz; i++) {
                float best = INFINITY;
                for (int j = 0; j < nz; j++) {
                    float t = sqstep(i - j, sp[2]);
                    if (!(t < best) && j > i) break;
                    float v = g2[((size_t)j * ny + y) * nx + x] + t;
                    if (v < best) best = v;
                }
                out[((size_t)i * ny + y) * nx + x] = best;
            }
    free(g1);
    free(g2);
}
OR_API void or_edt(const uint8_t* a, float* out, int nx, int ny, int nz, const float sp[3]) {
    size_t n = nvox(nx, ny, nz);
    edt_squared(a, out, nx, ny, nz, sp);
    for (size_t i = 0; i < n; i++) out[i] = sqrtf(out[i]);
}
/* distlt / distleq / distgeq / distgt: the interval test on the distance map. */
OR_API void or_dist_cmp(const uint8_t* a, int op, float r, uint8_t* out, int nx, int ny, int nz,
                        const float sp[3]) {
    size_t n = nvox(nx, ny, nz);
    float* d = (float*)malloc(n * sizeof(float));
    or_edt(a, d, nx, ny, nz, sp);
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)cmp1(d[i], r, op);
    free(d);
}
/* Signed variant in the Maurer convention (negative inside): outside = d(x, a);
 * inside = -d(x, boundary(a)), boundary = voxels of a adjacent to a voxel outside a. */
OR_API void or_edt_signed(const uint8_t* a, float* out, int nx, int ny, int nz, const float sp[3], int conn) {
    size_t n = nvox(nx, ny, nz);
    uint8_t* t = (uint8_t*)malloc(n);
    uint8_t* bnd = (uint8_t*)malloc(n);
    float* din = (float*)malloc(n * sizeof(float));
    or_not(a, t, n);
    or_near(t, bnd, nx, ny, nz, conn);
    or_and(a, bnd, bnd, n);
    or_edt(a, out, nx, ny, nz, sp);
    or_edt(bnd, din, nx, ny, nz, sp);
    for (size_t i = 0; i < n; i++)
        if (a[i]) out[i] = 0.0f - din[i];
    free(t);
    free(bnd);
    free(din);
}

/* ----------------------------------------------------------------------------
 * A7 / A11 global reductions (Greta.pdf p.43 "Volume, max, min (global)"). */
OR_API double or_volume(const uint8_t* a, size_t n) {
    double c = 0;
    for (size_t i = 0; i < n; i++) c += a[i] != 0;
    return c;
}
OR_API double or_min(const float* a, const uint8_t* mask, size_t n) {
    double m = INFINITY;
    for (size_t i = 0; i < n; i++)
        if (!mask || mask[i]) m = fmin(m, (double)a[i]);
    return m;
}
OR_API double or_max(const float* a, const uint8_t* mask, size_t n) {
    double m = -INFINITY;
    for (size_t i = 0; i < n; i++)
        if (!mask || mask[i]) m = fmax(m, (double)a[i]);
    return m;
}
OR_API double or_sum(const float* a, size_t n) {
    double s = 0;
    for (size_t i = 0; i < n; i++) s += (double)a[i];
    return s;
}

/* A9 percentiles(img, mask, correction): rank of a voxel's value among the masked
 * values, normalised (Greta.pdf p.43 "Normalization", p.52).  SURVEY.md Q6:
 *   (#less + correction * #equal) / |mask|   in mask,   0 outside. */
static int cmp_float(const void* p, const void* q) {
    float a = *(const float*)p, b = *(const float*)q;
    return (a > b) - (a < b);
}
OR_API void or_percentiles(const float* a, const uint8_t* mask, double correction, float* out, size_t n) {
    size_t cnt = 0;
    for (size_t i = 0; i < n; i++) cnt += mask[i] != 0;
    float* v = (float*)malloc((cnt + 1) * sizeof(float));
    size_t k = 0;
    for (size_t i = 0; i < n; i++)
        if (mask[i]) v[k++] = a[i];
    qsort(v, cnt, sizeof(float), cmp_float);
    for (size_t i = 0; i < n; i++) {
        if (!mask[i]) {
            out[i] = 0.0f;
            continue;
        }
        size_t lo = 0, hi = cnt; /* first index with v >= a[i] */
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (v[mid] < a[i]) lo = mid + 1; else hi = mid;
        }
        size_t less = lo;
        hi = cnt; /* first index with v > a[i] */
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (v[mid] <= a[i]) lo = mid + 1; else hi = mid;
        }
        size_t eq = lo - less;
        double num = (double)less;
        if (correction != 0.0) num = num + correction * (double)eq;
        out[i] = (float)(num / (double)cnt);
    }
    free(v);
}

/* ----------------------------------------------------------------------------
 * A6 crossCorrelation(rho, a, b, region, m, M, k)  (Greta.pdf p.43 "Texture
 * analysis", p.53).  Histogram bins [m + i*D, m + (i+1)*D), D=(M-m)/k;
 * values outside [m, M) are not counted.  The neighbourhood is the Euclidean ball
 * {y : |y-x| <= rho} in physical units, clipped to the image. */
static int bin_of(float v, double m, double M, int k) {
    double dv = (double)v;
    if (!(dv >= m) || !(dv < M)) return -1;
    double t = ((dv - m) * (double)k) / (M - m);
    int i = (int)t;
    if (i >= k) i = k - 1;
    return i;
}
static int in_ball(int dx, int dy, int dz, const float sp[3], double r2) {
    double ax = dx * (double)sp[0], ay = dy * (double)sp[1], az = dz * (double)sp[2];
    double s = ax * ax;
    s = s + ay * ay;
    s = s + az * az;
    return s <= r2;
}
/* (double) of the exact integer k*S22 - n2^2.  The integer can exceed 63 bits (one bin of a
 * region histogram holding > 4.3e8 voxels with k = 100), so it is formed in 128 bits; the
 * int128 -> binary64 conversion is correctly rounded (libgcc __floattidf).  *positive = d2 > 0. */
static double region_d2(const long long* h2, int k, long long* n2_out, int* positive) {
    unsigned __int128 s22 = 0;
    long long n2 = 0;
    for (int i = 0; i < k; i++) {
        n2 += h2[i];
        s22 += (unsigned __int128)(unsigned long long)h2[i] * (unsigned long long)h2[i];
    }
    __int128 d2 = (__int128)((unsigned __int128)k * s22) - (__int128)n2 * (__int128)n2;
    *n2_out = n2;
    *positive = d2 > 0;
    return (double)d2;
}

/* textbook = 0: exact-integer form  (k*P - n1*n2) / sqrt((k*S11 - n1^2)(k*S22 - n2^2))
 * textbook = 1: sum((h1-mean1)(h2-mean2)) / sqrt(sum((h1-mean1)^2) sum((h2-mean2)^2))
 * The two agree to rounding; the first is the arithmetic contract of the CUDA path.
 * Every voxel's bin is computed once up front (bin_of is a pure function of the voxel);
 * the neighbourhood histogram is still rebuilt from the whole ball for every voxel. */
static void xcorr_core(const float* a, const long long* h2, float rho, double m, double M, int k,
                       float* out, int nx, int ny, int nz, const float sp[3], int textbook) {
    long long n2 = 0;
    int d2pos = 0;
    const double d2d = region_d2(h2, k, &n2, &d2pos);
    double r2 = (double)rho * (double)rho;
    int wx = 0, wy = 0, wz = 0;
    if (rho >= 0.0f) {
        while (in_ball(wx + 1, 0, 0, sp, r2)) wx++;
        while (in_ball(0, wy + 1, 0, sp, r2)) wy++;
        while (in_ball(0, 0, wz + 1, sp, r2)) wz++;
    }
    /* the ball as an explicit offset list */
    int cap = (2 * wx + 1) * (2 * wy + 1) * (2 * wz + 1), noff = 0;
    int* off = (int*)malloc((size_t)(cap > 0 ? cap : 1) * 3 * sizeof(int));
    if (rho >= 0.0f)
        for (int dz = -wz; dz <= wz; dz++)
            for (int dy = -wy; dy <= wy; dy++)
                for (int dx = -wx; dx <= wx; dx++)
                    if (in_ball(dx, dy, dz, sp, r2)) {
                        off[3 * noff] = dx; off[3 * noff + 1] = dy; off[3 * noff + 2] = dz;
                        noff++;
                    }
    const size_t n = nvox(nx, ny, nz);
    int16_t* bins = (int16_t*)malloc((n ? n : 1) * sizeof(int16_t));
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) bins[i] = (int16_t)bin_of(a[i], m, M, k);
#pragma omp parallel
    {
        int* h1 = (int*)malloc((size_t)k * sizeof(int));
#pragma omp for schedule(dynamic, 1)
        for (int z = 0; z < nz; z++)
            for (int y = 0; y < ny; y++)
                for (int x = 0; x < nx; x++) {
                    memset(h1, 0, (size_t)k * sizeof(int));
                    const int inner = x >= wx && x + wx < nx && y >= wy && y + wy < ny && z >= wz && z + wz < nz;
                    for (int o = 0; o < noff; o++) {
                        int xx = x + off[3 * o], yy = y + off[3 * o + 1], zz = z + off[3 * o + 2];
                        if (!inner && (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz)) continue;
                        int bi = bins[((size_t)zz * ny + yy) * nx + xx];
                        if (bi >= 0) h1[bi]++;
                    }
                    float res = 0.0f;
                    if (!textbook) {
                        long long n1 = 0, s11 = 0, p = 0;
                        for (int i = 0; i < k; i++) {
                            n1 += h1[i];
                            s11 += (long long)h1[i] * h1[i];
                            p += (long long)h1[i] * h2[i];
                        }
                        long long num = (long long)k * p - n1 * n2;
                        long long d1 = (long long)k * s11 - n1 * n1;
                        if (d1 > 0 && d2pos) res = (float)((double)num / sqrt((double)d1 * d2d));
                    } else {
                        double m1 = 0, m2 = 0;
                        for (int i = 0; i < k; i++) {
                            m1 += (double)h1[i];
                            m2 += (double)h2[i];
                        }
                        m1 /= k;
                        m2 /= k;
                        double sxy = 0, sxx = 0, syy = 0;
                        for (int i = 0; i < k; i++) {
                            double u = (double)h1[i] - m1, w = (double)h2[i] - m2;
                            sxy += u * w;
                            sxx += u * u;
                            syy += w * w;
                        }
                        if (sxx > 0 && syy > 0) res = (float)(sxy / sqrt(sxx * syy));
                    }
                    out[((size_t)z * ny + y) * nx + x] = res;
                }
        free(h1);
    }
    free(bins);
    free(off);
}

OR_API void or_cross_correlation(const float* a, const float* b, const uint8_t* region, float rho,
                                 double m, double M, int k, float* out, int nx, int ny, int nz,
                                 const float sp[3], int textbook) {
    size_t n = nvox(nx, ny, nz);
    long long* h2 = (long long*)calloc(k, sizeof(long long));
    for (size_t i = 0; i < n; i++)
        if (region[i]) {
            int bi = bin_of(b[i], m, M, k);
            if (bi >= 0) h2[bi]++;
        }
    xcorr_core(a, h2, rho, m, M, k, out, nx, ny, nz, sp, textbook);
    free(h2);
}

/* The same with the region histogram h2 given by the caller (a slab of a larger volume whose
 * region histogram was accumulated over all slabs). */
OR_API void or_cross_correlation_h2(const float* a, const unsigned long long* h2u, float rho, double m,
                                    double M, int k, float* out, int nx, int ny, int nz, const float sp[3]) {
    long long* h2 = (long long*)calloc(k, sizeof(long long));
    for (int i = 0; i < k; i++) h2[i] = (long long)h2u[i];
    xcorr_core(a, h2, rho, m, M, k, out, nx, ny, nz, sp, 0);
    free(h2);
}

/* Host-side check value for the device's 128-bit statistic: (double)(k*S22 - n2^2) and its sign. */
OR_API double or_region_d2(const unsigned long long* h2u, int k, int* positive) {
    long long* h2 = (long long*)calloc(k, sizeof(long long));
    for (int i = 0; i < k; i++) h2[i] = (long long)h2u[i];
    long long n2;
    double d = region_d2(h2, k, &n2, positive);
    free(h2);
    return d;
}

OR_API int or_version(void) { return 101; }
