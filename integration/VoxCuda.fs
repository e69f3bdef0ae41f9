// VoxCuda.fs -- P/Invoke binding of libvoxcuda.so (include/voxcuda.h) for VoxLogicA's model layer.
//
// STATUS: source only.  No .NET toolchain exists in the build image, so this file has never been
// compiled; it is kept mechanical (one DllImport per export of include/voxcuda.h, blittable types
// only) so that a maintainer can drop it next to the SimpleITK-backed model and register the same
// operator names.  The Python table voxlogica_b200/_abi.py::SIGNATURES is the tested twin of the
// declarations below (tests/test_imgql_host.py checks that every symbol of the header is exported by
// the library AND declared here with the same number of parameters).
//
// Mapping to the reference: VoxLogicA's evaluator hands each task of the operator DAG to the model
// layer (Greta.pdf p.7, p.43); the F# sources are not in the mounted reference (SURVEY.md section 0),
// so operator names below follow the ImgQL primitives used on Greta.pdf p.48 and p.51-53.
namespace VoxLogicA.Cuda

open System
open System.Runtime.InteropServices


/// one instruction of a fused pointwise program (vx_op of include/voxcuda.h)
[<Struct; StructLayout(LayoutKind.Sequential)>]
type VxOp =
    val mutable opcode: int
    val mutable dst: int
    val mutable a: int
    val mutable b: int
    val mutable aux: int
    val mutable imm: float32

module Native =
    [<Literal>]
    let Lib = "voxcuda" // resolves to libvoxcuda.so on Linux

    // ---- lifecycle
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_version()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern nativeint vx_last_error()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_init(int device, uint64& ctx)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_shutdown(uint64 ctx)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_sync(uint64 ctx)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_upload_batch(uint64 ctx, nativeint host, int dtype, int nx, int ny, int nz, int nb, float32[] spacing, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_download(uint64 ctx, uint64 img, nativeint host, unativeint bytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_upload_scaled(uint64 ctx, nativeint host, int hostDtype, float32 slope, float32 inter, int nx, int ny, int nz, int nb, float32[] spacing, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_download_bits(uint64 ctx, uint64 img, nativeint host, unativeint bytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_info(uint64 img, int& dtype, int& nx, int& ny, int& nz, int& nb, float32[] spacing3)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_retain(uint64 img)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_release(uint64 img)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_mem_info(uint64 ctx, uint64& liveBytes, uint64& liveHandles, uint64& poolReserved)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_mem_trim(uint64 ctx, uint64 keepBytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_set_option(uint64 ctx, int option, int64 value)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_mem_set_budget(uint64 ctx, uint64 bytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_mem_stats(uint64 ctx, uint64& inUseBytes, uint64& cachedBytes, uint64& peakBytes, uint64& oomRetries)

    // ---- A1/A2 pointwise
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_const_bool(uint64 ctx, uint64 like, int value, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_const_f32(uint64 ctx, uint64 like, float32 value, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_not(uint64 ctx, uint64 a, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_and(uint64 ctx, uint64 a, uint64 b, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_or(uint64 ctx, uint64 a, uint64 b, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_xor(uint64 ctx, uint64 a, uint64 b, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_cmp_scalar(uint64 ctx, uint64 a, int op, float32 s, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_cmp(uint64 ctx, uint64 a, uint64 b, int op, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_arith_scalar(uint64 ctx, uint64 a, int op, float32 s, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_arith(uint64 ctx, uint64 a, uint64 b, int op, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_to_f32(uint64 ctx, uint64 a, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_to_bool(uint64 ctx, uint64 a, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_mask(uint64 ctx, uint64 a, uint64 m, float32 fill, uint64& out)

    // ---- A10 / A4 / A3 / A8
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_border(uint64 ctx, uint64 like, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_near(uint64 ctx, uint64 a, int conn, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_interior(uint64 ctx, uint64 a, int conn, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_through(uint64 ctx, uint64 a, uint64 b, int conn, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_components(uint64 ctx, uint64 a, int conn, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_maxvol(uint64 ctx, uint64 a, int conn, uint64& out)

    // ---- A5 distance
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_edt(uint64 ctx, uint64 a, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_edt_signed(uint64 ctx, uint64 a, int conn, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_dist_lt(uint64 ctx, uint64 a, float32 r, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_dist_leq(uint64 ctx, uint64 a, float32 r, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_dist_geq(uint64 ctx, uint64 a, float32 r, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_dist_gt(uint64 ctx, uint64 a, float32 r, uint64& out)

    // ---- A6 / A7 / A9 / A11
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_cross_correlation(uint64 ctx, uint64 a, uint64 b, uint64 region, float32 rho, double[] m, double[] M, int k, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_volume(uint64 ctx, uint64 a, double[] out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_min(uint64 ctx, uint64 a, double[] out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_max(uint64 ctx, uint64 a, double[] out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_minmax(uint64 ctx, uint64 a, uint64 mask, double[] outMin, double[] outMax)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_percentiles(uint64 ctx, uint64 a, uint64 mask, double correction, uint64& out)


    // ---- the remaining exports of include/voxcuda.h (generated from the header, one per VX_API
    // function: single-volume upload, batch slicing, the fusion hook, per-volume scalar operands,
    // masked reductions, the slab building blocks, events / profiling).  Device or host-or-device
    // buffers are nativeint; scalar outputs are byrefs.
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_upload(uint64 ctx, nativeint host, int dtype, int nx, int ny, int nz, float32[] spacing, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_batch_slice(uint64 ctx, uint64 img, int first, int count, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_crop_z(uint64 ctx, uint64 img, int z0, int nz, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_concat_z(uint64 ctx, uint64[] parts, int n_parts, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_transfer(uint64 dst_ctx, uint64 img, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_device_ptr(uint64 ctx, uint64 img, nativeint& ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_stream_handle(uint64 ctx, nativeint& stream)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ipc_alloc(uint64 ctx, unativeint bytes, nativeint& ptr, nativeint handle64)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ipc_open(uint64 ctx, nativeint handle64, nativeint& ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ipc_close(uint64 ctx, nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ipc_free(uint64 ctx, nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_copy_planes_to(uint64 ctx, uint64 img, int z0, int nz, nativeint dst)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_host_alloc(unativeint bytes, nativeint& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_host_alloc_wc(unativeint bytes, nativeint& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_host_free(nativeint p)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_cmp_scalar_batch(uint64 ctx, uint64 a, int op, float32[] scalars, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_arith_scalar_batch(uint64 ctx, uint64 a, int op, float32[] scalars, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_fused_pointwise(uint64 ctx, VxOp[] program, int n_ops, uint64[] leaves, int n_leaves, float32[] batch_scalars, int n_scalars, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_near_halo(uint64 ctx, uint64 a, int conn, nativeint lo_plane, nativeint hi_plane, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_interior_halo(uint64 ctx, uint64 a, int conn, nativeint lo_plane, nativeint hi_plane, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_create(uint64 ctx, uint64 b, int conn, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_seed(uint64 ctx, uint64 state, uint64 a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_plane(uint64 ctx, uint64 state, int z, uint64& labels_u32, uint64& reached_u8)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_flag(uint64 ctx, uint64 state, nativeint labels, uint64 n)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_select(uint64 ctx, uint64 state, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_count(uint64 ctx, uint64 state, uint64& local_max)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_sizes(uint64 ctx, uint64 state, nativeint labels, uint64 n, nativeint sizes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_flag_size(uint64 ctx, uint64 state, uint32 size)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_ccl_destroy(uint64 ctx, uint64 state)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_edt_sq_xy(uint64 ctx, uint64 a, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_edt_z_from_sq(uint64 ctx, uint64 sq, double max_in, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_region_histogram(uint64 ctx, uint64 b, uint64 region, double[] m, double[] M, int k, uint64[] h2)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_cross_correlation_h2(uint64 ctx, uint64 a, uint64[] h2, float32 rho, double[] m, double[] M, int k, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_xcorr_bin_edges(double m, double M, int k, float32[] edges)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_xcorr_region_stat(uint64[] h2, int k, double& d2, int& positive)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_min_masked(uint64 ctx, uint64 a, uint64 mask, double[] out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_max_masked(uint64 ctx, uint64 a, uint64 mask, double[] out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_sum(uint64 ctx, uint64 a, double[] out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_pc_pairs(uint64 ctx, uint64 a, uint64 mask, nativeint pairs_dev, uint64 capacity, uint64& n)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_pc_sort(uint64 ctx, nativeint pairs_dev, uint64 n, nativeint scratch_dev)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_pc_rank(uint64 ctx, nativeint sorted_dev, uint64 n, uint64 less_before, uint64 n_total, double correction, nativeint dst_dev)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_pc_scatter(uint64 ctx, nativeint pairs_dev, nativeint values_dev, uint64 n, uint64 like, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_event_create(uint64 ctx, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_event_record(uint64 ctx, uint64 ev)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_event_sync(uint64 ev)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_event_elapsed_ms(uint64 start, uint64 stop, float32& ms)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_event_destroy(uint64 ev)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_launch_count(uint64 ctx, uint64& out)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_prof_enable(uint64 ctx, int on)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_prof_reset(uint64 ctx)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_prof_report(uint64 ctx, nativeint buf, unativeint buf_bytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int vx_flush_l2(uint64 ctx)

exception VoxCudaException of code: int * message: string

/// Owns one reference of a device-resident image.  Finalisers may run on any thread:
/// vx_release is atomic and only enqueues a stream-ordered free.
type ImageHandle(h: uint64) =
    inherit SafeHandle(IntPtr.Zero, true)
    do base.SetHandle(nativeint (int64 h))
    member _.Value = h
    override this.IsInvalid = (this.handle = IntPtr.Zero)
    override this.ReleaseHandle() = Native.vx_release (uint64 (int64 this.handle)) = 0

/// The operator table a VoxLogicA model registers (names = ImgQL primitives).
type CudaModel(device: int) =
    let mutable ctx = 0UL
    let check rc =
        if rc <> 0 then
            raise (VoxCudaException(rc, Marshal.PtrToStringAnsi(Native.vx_last_error ())))
    /// the status check of a call that took image handles: the handles are kept reachable until the
    /// native call has returned (a raw uint64 handed to P/Invoke does not keep its SafeHandle alive, so
    /// without this the finaliser thread could release an operand while the kernel launch is still
    /// being set up; the library additionally pins every operand for the length of the call)
    let checkKeeping (operands: ImageHandle list) rc =
        for h in operands do GC.KeepAlive h
        check rc
    do check (Native.vx_init (device, &ctx))
    let conn = 26 // SURVEY.md Q1: adjacency of near/through; 6 selects face adjacency
    let unary (f: uint64 * uint64 * byref<uint64> -> int) (a: ImageHandle) =
        let mutable o = 0UL
        checkKeeping [ a ] (f (ctx, a.Value, &o))
        new ImageHandle(o)

    member _.Not(a: ImageHandle) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_not (ctx, a.Value, &o)); new ImageHandle(o)
    member _.And(a: ImageHandle, b: ImageHandle) = let mutable o = 0UL in checkKeeping [ a; b ] (Native.vx_and (ctx, a.Value, b.Value, &o)); new ImageHandle(o)
    member _.Or(a: ImageHandle, b: ImageHandle) = let mutable o = 0UL in checkKeeping [ a; b ] (Native.vx_or (ctx, a.Value, b.Value, &o)); new ImageHandle(o)
    member _.Border(like: ImageHandle) = let mutable o = 0UL in checkKeeping [ like ] (Native.vx_border (ctx, like.Value, &o)); new ImageHandle(o)
    member _.Near(a: ImageHandle) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_near (ctx, a.Value, conn, &o)); new ImageHandle(o)
    member _.Interior(a: ImageHandle) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_interior (ctx, a.Value, conn, &o)); new ImageHandle(o)
    member _.Through(a: ImageHandle, b: ImageHandle) = let mutable o = 0UL in checkKeeping [ a; b ] (Native.vx_through (ctx, a.Value, b.Value, conn, &o)); new ImageHandle(o)
    member _.MaxVol(a: ImageHandle) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_maxvol (ctx, a.Value, conn, &o)); new ImageHandle(o)
    member _.Lt(a: ImageHandle, s: float) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_cmp_scalar (ctx, a.Value, 0, float32 s, &o)); new ImageHandle(o)
    member _.Leq(a: ImageHandle, s: float) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_cmp_scalar (ctx, a.Value, 1, float32 s, &o)); new ImageHandle(o)
    member _.Gt(a: ImageHandle, s: float) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_cmp_scalar (ctx, a.Value, 2, float32 s, &o)); new ImageHandle(o)
    member _.Geq(a: ImageHandle, s: float) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_cmp_scalar (ctx, a.Value, 3, float32 s, &o)); new ImageHandle(o)
    member _.DistLt(r: float, a: ImageHandle) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_dist_lt (ctx, a.Value, float32 r, &o)); new ImageHandle(o)
    member _.DistGeq(r: float, a: ImageHandle) = let mutable o = 0UL in checkKeeping [ a ] (Native.vx_dist_geq (ctx, a.Value, float32 r, &o)); new ImageHandle(o)
    member _.Percentiles(a: ImageHandle, mask: ImageHandle, correction: float) =
        let mutable o = 0UL in checkKeeping [ a; mask ] (Native.vx_percentiles (ctx, a.Value, mask.Value, correction, &o)); new ImageHandle(o)
    member _.CrossCorrelation(rho: float, a: ImageHandle, b: ImageHandle, region: ImageHandle, m: float, M: float, k: int) =
        let mutable o = 0UL
        checkKeeping [ a; b; region ] (Native.vx_cross_correlation (ctx, a.Value, b.Value, region.Value, float32 rho, [| m |], [| M |], k, &o))
        new ImageHandle(o)
    /// a whole boolean / arithmetic sub-DAG in ONE launch: the evaluator may hand over a maximal
    /// pointwise sub-term of the operator DAG instead of one task per operator (include/voxcuda.h: vx_op)
    member _.Fused(program: VxOp[], leaves: ImageHandle[]) =
        let mutable o = 0UL
        let hs = leaves |> Array.map (fun l -> l.Value)
        checkKeeping (List.ofArray leaves) (Native.vx_fused_pointwise (ctx, program, program.Length, hs, hs.Length, null, 0, &o))
        new ImageHandle(o)
    member _.Volume(a: ImageHandle) = let r = [| 0.0 |] in checkKeeping [ a ] (Native.vx_volume (ctx, a.Value, r)); r.[0]
    member _.Min(a: ImageHandle) = let r = [| 0.0 |] in checkKeeping [ a ] (Native.vx_min (ctx, a.Value, r)); r.[0]
    member _.Max(a: ImageHandle) = let r = [| 0.0 |] in checkKeeping [ a ] (Native.vx_max (ctx, a.Value, r)); r.[0]
    /// min and max from one pass (the scheduler can serve both `min(img)` and `max(img)` tasks from it)
    member _.MinMax(a: ImageHandle) =
        let lo, hi = [| 0.0 |], [| 0.0 |]
        checkKeeping [ a ] (Native.vx_minmax (ctx, a.Value, 0UL, lo, hi))
        lo.[0], hi.[0]

    /// pinned managed array -> HBM (dtype 1 = u8 boolean, 2 = f32 number)
    member _.Upload(data: Array, dtype: int, nx: int, ny: int, nz: int, spacing: float32[]) =
        let pin = GCHandle.Alloc(data, GCHandleType.Pinned)
        try
            let mutable o = 0UL
            check (Native.vx_upload_batch (ctx, pin.AddrOfPinnedObject(), dtype, nx, ny, nz, 1, spacing, &o))
            check (Native.vx_sync ctx) // the array is unpinned on return
            new ImageHandle(o)
        finally
            pin.Free()
    member _.Download(img: ImageHandle, dst: Array, bytes: int64) =
        let pin = GCHandle.Alloc(dst, GCHandleType.Pinned)
        try checkKeeping [ img ] (Native.vx_download (ctx, img.Value, pin.AddrOfPinnedObject(), unativeint bytes))
        finally pin.Free()
    interface IDisposable with
        member _.Dispose() = if ctx <> 0UL then (Native.vx_shutdown ctx |> ignore; ctx <- 0UL)
