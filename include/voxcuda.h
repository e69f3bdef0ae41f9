/*
 * voxcuda.h — C ABI of libvoxcuda.so, the B200 (sm_100a) execution backend for
 * VoxLogicA's ImgQL spatial-logic operators.
 *
 * What this replaces.  In VoxLogicA the ImgQL evaluator hands every task of the
 * operator DAG to a "model layer" whose operators are SimpleITK calls
 * (reference: README.md:8 "describe your tasks using short text files";
 * Greta.pdf p.7 "global model checking -> parallel execution, GPU computation,
 * caching"; Greta.pdf p.43 operator families).  The F# sources of that layer are
 * NOT in the mounted reference, so every entry point below
 * cites the slide that defines the operator it implements, not an F# line.
 *
 * Calling convention.  extern "C", cdecl, blittable types only, so that the F#
 * side is a table of [<DllImport("voxcuda")>] one-liners.
 *   - every function returns int32_t: VX_OK (0) or a negative VX_E_* code;
 *     vx_last_error() returns a thread-local, library-owned message.
 *   - images are opaque 64-bit handles (0 = invalid).  The caller owns ONE
 *     reference per handle returned through an out-parameter and drops it with
 *     vx_release(); vx_retain()/vx_release() are atomic and may be called from any
 *     thread, including a .NET finaliser thread.
 *   - all calls are asynchronous with respect to the device: they enqueue work on
 *     the calling thread's CUDA stream of the context and return.  Dependencies
 *     between streams are tracked with events stored in the handles.  The host
 *     blocks only in vx_download, vx_sync and the scalar-returning reductions.
 *   - images are resident in HBM: dtype u8 (boolean, values 0/1), f32 (numeric)
 *     or u32 (component labels); layout [nb][nz][ny][nx], x fastest, dense.
 *     nb is a leading batch of independent volumes processed by one launch
 *     (BraTS-sized volumes are far too small to fill a B200 one at a time).
 *   - there is no CPU fallback anywhere: without a CUDA device vx_init fails.
 */
#ifndef VOXCUDA_H
#define VOXCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VX_API __declspec(dllexport)
#else
#define VX_API __attribute__((visibility("default")))
#endif

typedef uint64_t vx_ctx; /* opaque context: one per (process, device) */
typedef uint64_t vx_img; /* opaque image handle; 0 is never valid       */
typedef uint64_t vx_evt; /* opaque timing event                         */

/* status codes */
#define VX_OK 0
#define VX_E_INVALID_ARG (-1)
#define VX_E_SHAPE_MISMATCH (-2)
#define VX_E_DTYPE (-3)
#define VX_E_OOM (-4)
#define VX_E_CUDA (-5)
#define VX_E_NCCL (-6)
#define VX_E_UNSUPPORTED (-7)

/* element types */
#define VX_U8 1  /* boolean valuation, 0 or 1           */
#define VX_F32 2 /* numeric valuation                    */
#define VX_U32 3 /* component labels (vx_components)     */
/* host-side voxel types accepted by vx_upload_scaled only (a file's native 16-bit integers) */
#define VX_I16 4
#define VX_U16 5

/* comparison operators for vx_cmp / vx_cmp_scalar (ImgQL  <.  <=.  >.  >=.  =.) */
#define VX_LT 0
#define VX_LE 1
#define VX_GT 2
#define VX_GE 3
#define VX_EQ 4
#define VX_NE 5

/* arithmetic operators for vx_arith / vx_arith_scalar */
#define VX_ADD 0
#define VX_SUB 1
#define VX_MUL 2
#define VX_DIV 3
#define VX_RSUB 4 /* scalar - image  */
#define VX_RDIV 5 /* scalar / image  */
#define VX_MIN 6
#define VX_MAX 7

/* adjacency ("regular grids with chosen adjacency", Greta.pdf p.28):
 * 6 = face neighbours (4 in 2D), 26 = face+edge+corner (8 in 2D). */
#define VX_CONN_FACE 6
#define VX_CONN_FULL 26

/* ------------------------------------------------------------------ lifecycle */

/* Library version: major*10000 + minor*100 + patch. */
VX_API int32_t vx_version(void);
/* Thread-local message describing the last failure on the calling thread. */
VX_API const char* vx_last_error(void);

/* Create a context on CUDA device `device`.  Fails with VX_E_CUDA when no CUDA
 * device is usable (there is no CPU fallback). */
VX_API int32_t vx_init(int32_t device, vx_ctx* out);
VX_API int32_t vx_shutdown(vx_ctx ctx);
/* Block until every stream of the context is idle. */
VX_API int32_t vx_sync(vx_ctx ctx);

/* Copy a host array into a new HBM-resident image.  `host` holds
 * nb*nz*ny*nx elements of `dtype`; spacing may be NULL (= 1,1,1).  The copy is
 * enqueued with cudaMemcpyAsync; `host` may be reused once the call returns only
 * if it is pageable memory (CUDA stages it), for pinned memory call vx_sync or
 * use the image first. */
VX_API int32_t vx_upload(vx_ctx ctx, const void* host, int32_t dtype, int32_t nx, int32_t ny,
                         int32_t nz, const float* spacing, vx_img* out);
VX_API int32_t vx_upload_batch(vx_ctx ctx, const void* host, int32_t dtype, int32_t nx, int32_t ny,
                               int32_t nz, int32_t nb, const float* spacing, vx_img* out);
/* Upload of a file's native 16-bit voxels (NIfTI datatype 4 INT16 / 512 UINT16, Greta.pdf p.48
 * `load flair = "flair.nii.gz"`) together with the header's scl_slope / scl_inter: the numeric image
 *     fl( fl( (float)raw * slope ) + inter )
 * is formed on the device -- two binary32 roundings, no fused multiply-add, bit-identical to
 * converting on the host and uploading f32 -- so the host->device copy moves 2 bytes per voxel
 * instead of 4.  host_dtype: VX_I16 or VX_U16; slope 1, inter 0 give the plain conversion.
 * The 16-bit codes stay on the device with the image (2 more bytes per voxel) whenever the scaling is
 * strictly monotone over the type's whole range: such an image has at most 65536 distinct values in
 * the order of their codes, and vx_percentiles ranks it by counting (histogram, scan, table look-up)
 * instead of sorting -- same bits, about a sixth of the time (VX_OPT_QUANTISED_PATHS turns this off). */
VX_API int32_t vx_upload_scaled(vx_ctx ctx, const void* host, int32_t host_dtype, float slope, float inter,
                                int32_t nx, int32_t ny, int32_t nz, int32_t nb, const float* spacing,
                                vx_img* out);
/* Blocking copy of the whole image to host memory (`bytes` must equal its size). */
VX_API int32_t vx_download(vx_ctx ctx, vx_img img, void* host, size_t bytes);
/* Blocking copy of a boolean image as a bit stream: bit (i & 7) of byte (i >> 3) = voxel i of the
 * batch (x fastest, then y, z, volume), unused bits of the last byte 0; `bytes` must equal
 * ceil(voxels / 8).  The device->host copy moves 1 bit per voxel instead of 1 byte. */
VX_API int32_t vx_download_bits(vx_ctx ctx, vx_img img, void* host, size_t bytes);
/* Any out pointer may be NULL. */
VX_API int32_t vx_info(vx_img img, int32_t* dtype, int32_t* nx, int32_t* ny, int32_t* nz,
                       int32_t* nb, float* spacing3);
VX_API int32_t vx_retain(vx_img img);
VX_API int32_t vx_release(vx_img img);
/* View of volumes [first, first+count) of a batch as a new batch (copies). */
VX_API int32_t vx_batch_slice(vx_ctx ctx, vx_img img, int32_t first, int32_t count, vx_img* out);

/* Slab helpers (one oversized volume split along z across GPUs):
 * planes [z0, z0+nz) of every volume as a new image; the parts stacked along z. */
VX_API int32_t vx_crop_z(vx_ctx ctx, vx_img img, int32_t z0, int32_t nz, vx_img* out);
VX_API int32_t vx_concat_z(vx_ctx ctx, const vx_img* parts, int32_t n_parts, vx_img* out);
/* Copy an image to another context of the same process (another GPU; NVLink peer copy).
 * Keep the source referenced until vx_sync(dst_ctx). */
VX_API int32_t vx_transfer(vx_ctx dst_ctx, vx_img img, vx_img* out);
/* Interop with other CUDA code of the process (e.g. an NCCL halo exchange): the raw device
 * pointer of an image, valid while a reference is held, and the calling thread's CUDA stream
 * (as a cudaStream_t).  vx_upload / vx_download also accept DEVICE pointers as `host`. */
VX_API int32_t vx_device_ptr(vx_ctx ctx, vx_img img, void** ptr);
VX_API int32_t vx_stream_handle(vx_ctx ctx, void** stream);

/* Peer-visible device memory ("mailboxes") for the slab path: vx_ipc_alloc returns device memory
 * and a 64-byte CUDA IPC handle another PROCESS on the same node passes to vx_ipc_open to map
 * it (threads of the same process use the raw pointer).  vx_copy_planes_to packs planes
 * [z0, z0+nz) of every volume of an image into raw device memory on the caller's stream. */
VX_API int32_t vx_ipc_alloc(vx_ctx ctx, size_t bytes, void** ptr, void* handle64);
VX_API int32_t vx_ipc_open(vx_ctx ctx, const void* handle64, void** ptr);
VX_API int32_t vx_ipc_close(vx_ctx ctx, void* ptr);
VX_API int32_t vx_ipc_free(vx_ctx ctx, void* ptr);
VX_API int32_t vx_copy_planes_to(vx_ctx ctx, vx_img img, int32_t z0, int32_t nz, void* dst);

/* Pinned host staging memory for the end-to-end path. */
VX_API int32_t vx_host_alloc(size_t bytes, void** out);
/* Write-combined pinned memory: for upload staging buffers the CPU only writes (sequentially)
 * and the GPU reads -- not snooped during the DMA; very slow for CPU reads. */
VX_API int32_t vx_host_alloc_wc(size_t bytes, void** out);
VX_API int32_t vx_host_free(void* p);

/* Device-memory accounting ("garbage collection", Greta.pdf p.78-79): live image
 * bytes, number of live handles, bytes reserved by the stream-ordered pool. */
VX_API int32_t vx_mem_info(vx_ctx ctx, uint64_t* live_bytes, uint64_t* live_handles,
                           uint64_t* pool_reserved);
/* Return cached pool memory above `keep_bytes` to the driver. */
VX_API int32_t vx_mem_trim(vx_ctx ctx, uint64_t keep_bytes);
/* Device-memory budget of the context (the GPU may be shared with the host application's own
 * work): a soft cap, in bytes, on the memory held by live images, scratch and the free lists
 * together; 0 = no cap.  An allocation that would exceed it first hands every cached block back
 * to the pool -- the path a real out-of-memory takes -- and fails with VX_E_OOM only if live data
 * alone still does not leave room (Greta.pdf p.79: the reference's GPU branch died of exactly
 * this at 4099 tasks for want of any policy). */
VX_API int32_t vx_mem_set_budget(vx_ctx ctx, uint64_t bytes);
/* Allocator counters: bytes in blocks currently handed out, bytes parked in the free lists, the
 * high-water mark of their sum since the last vx_mem_set_budget call, and the number of
 * allocations that succeeded only after the free lists were released.  Any out pointer may be
 * NULL. */
VX_API int32_t vx_mem_stats(vx_ctx ctx, uint64_t* in_use_bytes, uint64_t* cached_bytes, uint64_t* peak_bytes,
                            uint64_t* oom_retries);

/* ------------------------------------------------- A1/A2 pointwise operators
 * Greta.pdf p.30 (not, and, true), p.35 (or, false), p.43 ("Colour & thresholds"). */
VX_API int32_t vx_const_bool(vx_ctx ctx, vx_img like, int32_t value, vx_img* out); /* tt / ff  */
VX_API int32_t vx_const_f32(vx_ctx ctx, vx_img like, float value, vx_img* out);
VX_API int32_t vx_not(vx_ctx ctx, vx_img a, vx_img* out);
VX_API int32_t vx_and(vx_ctx ctx, vx_img a, vx_img b, vx_img* out);
VX_API int32_t vx_or(vx_ctx ctx, vx_img a, vx_img b, vx_img* out);
VX_API int32_t vx_xor(vx_ctx ctx, vx_img a, vx_img b, vx_img* out);
/* out = (a OP scalar), f32 -> u8; the comparison is exact IEEE-754 binary32. */
VX_API int32_t vx_cmp_scalar(vx_ctx ctx, vx_img a, int32_t op, float scalar, vx_img* out);
/* per-volume scalars (one per batch entry) */
VX_API int32_t vx_cmp_scalar_batch(vx_ctx ctx, vx_img a, int32_t op, const float* scalars,
                                   vx_img* out);
VX_API int32_t vx_cmp(vx_ctx ctx, vx_img a, vx_img b, int32_t op, vx_img* out);
VX_API int32_t vx_arith_scalar(vx_ctx ctx, vx_img a, int32_t op, float scalar, vx_img* out);
VX_API int32_t vx_arith_scalar_batch(vx_ctx ctx, vx_img a, int32_t op, const float* scalars,
                                     vx_img* out);
VX_API int32_t vx_arith(vx_ctx ctx, vx_img a, vx_img b, int32_t op, vx_img* out);
/* u8 -> f32 (0.0/1.0) and f32 -> u8 (x != 0) conversions. */
VX_API int32_t vx_to_f32(vx_ctx ctx, vx_img a, vx_img* out);
VX_API int32_t vx_to_bool(vx_ctx ctx, vx_img a, vx_img* out);
/* mask(a, m): f32 image a where m is true, `fill` elsewhere. */
VX_API int32_t vx_mask(vx_ctx ctx, vx_img a, vx_img m, float fill, vx_img* out);

/* Fusion hook: a whole boolean/arithmetic sub-DAG in one launch, so that a chain
 * costs its leaf reads plus one root write (north_star: "fused along the DAG").
 * The program is a straight-line register machine.  Register r (0 <= r < 24) holds
 * either a boolean or a numeric value per voxel; instruction i writes register
 * `dst`.  Opcodes: */
#define VXP_LOAD 0   /* dst = leaves[a]                        */
#define VXP_NOT 1    /* dst = !r[a]                            */
#define VXP_AND 2    /* dst = r[a] & r[b]                      */
#define VXP_OR 3     /* dst = r[a] | r[b]                      */
#define VXP_XOR 4    /* dst = r[a] ^ r[b]                      */
#define VXP_CMPS 5   /* dst = r[a] <aux> imm          (f32->u8) */
#define VXP_CMP 6    /* dst = r[a] <aux> r[b]         (f32->u8) */
#define VXP_ARITHS 7 /* dst = r[a] <aux> imm          (f32)     */
#define VXP_ARITH 8  /* dst = r[a] <aux> r[b]         (f32)     */
#define VXP_TOF32 9  /* dst = (float) r[a]            (u8->f32) */
#define VXP_TOBOOL 10 /* dst = r[a] != 0              (f32->u8) */
#define VXP_CONSTB 11 /* dst = (aux != 0)             (u8)      */
#define VXP_CONSTF 12 /* dst = imm                    (f32)     */
#define VXP_SELECT 13 /* dst = r[a] ? r[b] : imm      (f32; a is u8) */
#define VXP_CMPSB 14  /* dst = r[a] <aux> per-volume scalar number b (f32->u8)  */
#define VXP_ARITHSB 15 /* dst = r[a] <aux> per-volume scalar number b (f32)     */
typedef struct vx_op {
    int32_t opcode; /* VXP_*                                     */
    int32_t dst;    /* destination register                      */
    int32_t a, b;   /* source registers / leaf index             */
    int32_t aux;    /* VX_LT.. / VX_ADD.. / constant             */
    float imm;      /* scalar operand                            */
} vx_op;
#define VXP_MAX_OPS 64
#define VXP_MAX_REGS 24
#define VXP_MAX_LEAVES 8
/* The result is the register written by the last instruction.  `batch_scalars`
 * (may be NULL) is an array [n_scalars][nb] of per-volume numbers used by
 * VXP_CMPSB / VXP_ARITHSB. */
VX_API int32_t vx_fused_pointwise(vx_ctx ctx, const vx_op* program, int32_t n_ops,
                                  const vx_img* leaves, int32_t n_leaves,
                                  const float* batch_scalars, int32_t n_scalars, vx_img* out);

/* ------------------------------------------------------ A10 border, A4 near/interior
 * border: true on the outer faces of the volume (Greta.pdf p.51 touch(.., border)).
 * near:     closure  N(A) = A u {x | exists a in A adjacent to x}  (p.24, p.30)
 * interior: I(A) = !N(!A)                                           (p.35)
 * Voxels outside the image do not exist: they neither dilate in nor erode. */
VX_API int32_t vx_border(vx_ctx ctx, vx_img like, vx_img* out);
VX_API int32_t vx_near(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out);
VX_API int32_t vx_interior(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out);
/* The same for one z-slab of a larger volume, fused with the halo exchange: the planes z = -1 and
 * z = nz are read by the kernel from lo_plane / hi_plane (nb*ny*nx bytes of device-accessible
 * memory each -- typically a neighbour GPU's mailbox mapped with vx_ipc_open, pulled over NVLink
 * while the kernel runs; NULL where the volume ends).  The caller orders their producers first. */
VX_API int32_t vx_near_halo(vx_ctx ctx, vx_img a, int32_t conn, const void* lo_plane,
                            const void* hi_plane, vx_img* out);
VX_API int32_t vx_interior_halo(vx_ctx ctx, vx_img a, int32_t conn, const void* lo_plane,
                                const void* hi_plane, vx_img* out);

/* --------------------------------------------- A3 reachability (Greta.pdf p.32-38, p.78)
 * through(a, b): union of the connected components of b (adjacency `conn`) that
 * contain at least one voxel of a.  "Reachability via Connected Components". */
VX_API int32_t vx_through(vx_ctx ctx, vx_img a, vx_img b, int32_t conn, vx_img* out);
/* Connected-component labels of a (u32 image): 0 on background, otherwise
 * 1 + linear index (z*ny*nx + y*nx + x, within its volume) of the component's
 * first voxel in raster order -- a canonical labelling, so it is comparable
 * bit-for-bit with any other implementation after the same canonicalisation. */
VX_API int32_t vx_components(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out);
/* The stages of vx_through as separate calls on a kept union-find state, for a caller that owns
 * only one z-slab of a larger volume: label the local slab once, seed it,
 * read the labels and "reached" flags of its boundary planes (a label is 1 + the linear index of
 * the component's first voxel inside the slab), flag further components by label after the
 * boundary equivalences have been resolved with the neighbours, then select.  Calls on one
 * state must be issued in program order (any thread); the state owns device scratch until
 * vx_ccl_destroy. */
typedef uint64_t vx_ccl;
VX_API int32_t vx_ccl_create(vx_ctx ctx, vx_img b, int32_t conn, vx_ccl* out);
VX_API int32_t vx_ccl_seed(vx_ctx ctx, vx_ccl state, vx_img a);
VX_API int32_t vx_ccl_plane(vx_ctx ctx, vx_ccl state, int32_t z, vx_img* labels_u32, vx_img* reached_u8);
/* labels: n values, host or device memory; single-volume states only */
VX_API int32_t vx_ccl_flag(vx_ctx ctx, vx_ccl state, const uint32_t* labels, uint64_t n);
VX_API int32_t vx_ccl_select(vx_ctx ctx, vx_ccl state, vx_img* out);
/* Component sizes on a kept state -- the local part of a slab-decomposed maxvol (single-volume
 * states).  vx_ccl_count counts the voxels of every local component and returns the largest
 * count (host blocks); vx_ccl_sizes looks up the counts of the components with the given labels
 * (labels / sizes: n values, host or device memory; host blocks); vx_ccl_flag_size flags every
 * component of exactly `size` voxels, like vx_ccl_flag does by label, for vx_ccl_select. */
VX_API int32_t vx_ccl_count(vx_ctx ctx, vx_ccl state, uint64_t* local_max);
VX_API int32_t vx_ccl_sizes(vx_ctx ctx, vx_ccl state, const uint32_t* labels, uint64_t n, uint32_t* sizes);
VX_API int32_t vx_ccl_flag_size(vx_ctx ctx, vx_ccl state, uint32_t size);
VX_API int32_t vx_ccl_destroy(vx_ctx ctx, vx_ccl state);

/* A8 maxvol: the component(s) of a with the largest voxel count (all of them on ties). */
VX_API int32_t vx_maxvol(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out);

/* --------------------------------------------- A5 distance (Greta.pdf p.37, p.43)
 * edt: exact Euclidean distance (physical units, image spacing) from every voxel
 * to the nearest voxel of a; 0 inside a; +inf when a is empty in that volume. */
VX_API int32_t vx_edt(vx_ctx ctx, vx_img a, vx_img* out);
/* The two halves of vx_edt for a caller that owns one z-slab of a larger volume: the squared
 * distance after the x and y passes (in-plane, +inf where a plane holds no feature), and the z
 * pass + square root on such a map once its columns have been re-partitioned so that every
 * column is complete.  max_in: an upper bound of the finite values of
 * `sq` (selects the exact-integer lower-envelope pass when everything stays below 2^24). */
VX_API int32_t vx_edt_sq_xy(vx_ctx ctx, vx_img a, vx_img* out);
VX_API int32_t vx_edt_z_from_sq(vx_ctx ctx, vx_img sq, double max_in, vx_img* out);
/* Signed variant following the Maurer signed-distance convention: outside as
 * vx_edt; inside, minus the distance to the nearest voxel of a that has a
 * background neighbour (`conn`); 0 on those boundary voxels. */
VX_API int32_t vx_edt_signed(vx_ctx ctx, vx_img a, int32_t conn, vx_img* out);
/* Fused distance + interval test, u8 -> u8, never materialising the f32 map:
 * dist_lt: d(x,a) < r    dist_leq: <= r    dist_geq: >= r    dist_gt: > r  */
VX_API int32_t vx_dist_lt(vx_ctx ctx, vx_img a, float r, vx_img* out);
VX_API int32_t vx_dist_leq(vx_ctx ctx, vx_img a, float r, vx_img* out);
VX_API int32_t vx_dist_geq(vx_ctx ctx, vx_img a, float r, vx_img* out);
VX_API int32_t vx_dist_gt(vx_ctx ctx, vx_img a, float r, vx_img* out);

/* --------------------------------------------- A6 texture similarity (p.43, p.53)
 * For every voxel x of every volume: h1 = k-bin histogram of a over the voxels y
 * with |y-x| <= rho (physical units), h2 = k-bin histogram of b over `region`;
 * bins are [m + i*(M-m)/k, m + (i+1)*(M-m)/k), values outside [m,M) are not
 * counted; result = Pearson correlation coefficient of (h1,h2) in [-1,1]
 * (0 when either histogram has zero variance).  m, M: one value per volume. */
VX_API int32_t vx_cross_correlation(vx_ctx ctx, vx_img a, vx_img b, vx_img region, float rho,
                                    const double* m, const double* M, int32_t k, vx_img* out);

/* The two halves of vx_cross_correlation, for callers that own only a slab of the volume:
 * the k-bin histogram of b over region (h2: nb*k counters, accumulated on the host across
 * slabs by the caller) and the correlation map against a given h2. */
VX_API int32_t vx_region_histogram(vx_ctx ctx, vx_img b, vx_img region, const double* m,
                                   const double* M, int32_t k, uint64_t* h2);
VX_API int32_t vx_cross_correlation_h2(vx_ctx ctx, vx_img a, const uint64_t* h2, float rho,
                                       const double* m, const double* M, int32_t k, vx_img* out);

/* Host-only: the k+1 break points of the binning (edges[i] = smallest binary32 whose bin is >= i;
 * edges[0]: smallest value >= m, edges[k]: smallest value >= M).  Needs no device. */
VX_API int32_t vx_xcorr_bin_edges(double m, double M, int32_t k, float* edges);
/* Host-only: the region statistic of the correlation's denominator for one k-bin histogram h2
 * (counters < 2^32): *d2 = (double)(k*sum(h2^2) - sum(h2)^2), the exact integer formed in 128 bits
 * and rounded once; *positive = the integer is > 0.  Needs no device. */
VX_API int32_t vx_xcorr_region_stat(const uint64_t* h2, int32_t k, double* d2, int32_t* positive);

/* --------------------------------------------- A7/A9/A11 reductions and normalisation
 * Scalar results: `out` points to nb doubles (one per volume of the batch). */
VX_API int32_t vx_volume(vx_ctx ctx, vx_img a, double* out);
VX_API int32_t vx_min(vx_ctx ctx, vx_img a, double* out);
VX_API int32_t vx_max(vx_ctx ctx, vx_img a, double* out);
/* min/max restricted to mask (+inf/-inf when the mask is empty). */
VX_API int32_t vx_min_masked(vx_ctx ctx, vx_img a, vx_img mask, double* out);
VX_API int32_t vx_max_masked(vx_ctx ctx, vx_img a, vx_img mask, double* out);
VX_API int32_t vx_sum(vx_ctx ctx, vx_img a, double* out);
/* min and max of every volume from ONE pass over the image (the histogram range of
 * crossCorrelation, Greta.pdf p.48 similarTo: min(intensity(flair)), max(intensity(flair))).
 * mask = 0: whole image.  Same values as vx_min / vx_max. */
VX_API int32_t vx_minmax(vx_ctx ctx, vx_img a, vx_img mask, double* out_min, double* out_max);
/* percentiles(img, mask, correction): for a voxel x in mask,
 *   ( #{y in mask : img(y) < img(x)} + correction * #{y in mask : img(y) = img(x)} ) / |mask|
 * and 0 outside the mask (Greta.pdf p.52 percentiles(intensity, brain)). */
VX_API int32_t vx_percentiles(vx_ctx ctx, vx_img a, vx_img mask, double correction, vx_img* out);

/* Building blocks of a slab-decomposed percentiles (one rank of a z-slab decomposition,
 * DESIGN.md section 6): compact the masked voxels of the local slab into (sortable key << 32 |
 * voxel index) pairs, sort pairs by key, rank sorted pairs against a global count (the ranks are
 * written at dst[payload]), scatter values back into an image by payload.  All buffers are
 * caller-owned DEVICE memory; single-volume images; the calls block until the result is there. */
VX_API int32_t vx_pc_pairs(vx_ctx ctx, vx_img a, vx_img mask, uint64_t* pairs_dev, uint64_t capacity, uint64_t* n);
VX_API int32_t vx_pc_sort(vx_ctx ctx, uint64_t* pairs_dev, uint64_t n, uint64_t* scratch_dev);
VX_API int32_t vx_pc_rank(vx_ctx ctx, const uint64_t* sorted_dev, uint64_t n, uint64_t less_before,
                          uint64_t n_total, double correction, float* dst_dev);
VX_API int32_t vx_pc_scatter(vx_ctx ctx, const uint64_t* pairs_dev, const float* values_dev, uint64_t n,
                             vx_img like, vx_img* out);

/* --------------------------------------------- measurement helpers
 * Events are recorded on the calling thread's stream of the context. */
VX_API int32_t vx_event_create(vx_ctx ctx, vx_evt* out);
VX_API int32_t vx_event_record(vx_ctx ctx, vx_evt ev);
VX_API int32_t vx_event_sync(vx_evt ev);
VX_API int32_t vx_event_elapsed_ms(vx_evt start, vx_evt stop, float* ms);
VX_API int32_t vx_event_destroy(vx_evt ev);
/* Number of kernels this library has launched in the context so far. */
VX_API int32_t vx_launch_count(vx_ctx ctx, uint64_t* out);
/* Per-kernel device timing: when enabled, every kernel launch is bracketed by
 * CUDA events on its own stream.  vx_prof_report drains the events and writes
 * one line per kernel name: "name count total_ms\n". */
VX_API int32_t vx_prof_enable(vx_ctx ctx, int32_t on);
VX_API int32_t vx_prof_reset(vx_ctx ctx);
VX_API int32_t vx_prof_report(vx_ctx ctx, char* buf, size_t buf_bytes);
/* Context options.  VX_OPT_QUANTISED_PATHS (default 1): operators may use the 16-bit codes kept
 * with images made by vx_upload_scaled (see there); 0 forces the general algorithms -- the results
 * are identical, the switch exists for tests and measurements. */
#define VX_OPT_QUANTISED_PATHS 1
/* VX_OPT_ORDERED_UPLOADS (default 1): uploads of 1 MiB and more issued by different caller threads are
 * chained on the device (each copy waits for the one issued before it), so the host->device link
 * serves them one at a time in arrival order at its full rate instead of sharing it between them:
 * the caller that asked first has its data first.  0 = independent copies. */
#define VX_OPT_ORDERED_UPLOADS 2
VX_API int32_t vx_set_option(vx_ctx ctx, int32_t option, int64_t value);
/* Write a buffer larger than L2 so the next kernel starts cold. */
VX_API int32_t vx_flush_l2(vx_ctx ctx);

#ifdef __cplusplus
}
#endif
#endif /* VOXCUDA_H */
