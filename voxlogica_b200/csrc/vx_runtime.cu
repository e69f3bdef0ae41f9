// Runtime of libvoxcuda.so: contexts, image handles (T1 in SURVEY.md section 8a),
// stream-ordered device memory, per-thread streams, events, profiling.
//
// Design notes (B200-first, not a port of anything):
//  * images never leave HBM between operators; the only host traffic is
//    vx_upload / vx_download and the nb doubles of a scalar reduction;
//  * allocation is cudaMallocAsync on the device's default pool with the release
//    threshold raised to "never", so steady-state alloc/free of intermediates is a
//    pointer bump ("garbage collection", Greta.pdf p.78-79: the OpenCL backend
//    either ran out of memory or paid up to 64% for its GC);
//  * each host thread gets its own CUDA stream per context; handles carry a
//    `ready` event so a consumer on another stream waits on the device, never on
//    the host (the F# scheduler calls concurrently from thread-pool threads).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <type_traits>

#include "vx_internal.h"
#include "vx_tma.h"

namespace vx {

static thread_local char g_err[512] = "";
static thread_local Ctx* g_cur_ctx = nullptr;   // context of the call running on this thread (fault latch)

bool sticky_cuda_error(cudaError_t e) {
    switch (e) {
        case cudaErrorIllegalAddress: case cudaErrorLaunchFailure: case cudaErrorIllegalInstruction:
        case cudaErrorMisalignedAddress: case cudaErrorHardwareStackError: case cudaErrorInvalidPc:
        case cudaErrorInvalidAddressSpace: case cudaErrorECCUncorrectable: case cudaErrorAssert:
        case cudaErrorLaunchTimeout: case cudaErrorCudartUnloading: case cudaErrorDevicesUnavailable:
            return true;
        default:
            return false;
    }
}
void latch_fault(cudaError_t e, const char* what) {
    Ctx* c = g_cur_ctx;
    if (!c || !sticky_cuda_error(e)) return;
    int expected = 0;
    if (c->fault.compare_exchange_strong(expected, (int)e))
        snprintf(c->fault_msg, sizeof(c->fault_msg), "%s (first seen in %s)", cudaGetErrorString(e), what);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---- handle tables ----------------------------------------------------------
static std::mutex g_tab_mu;
static std::vector<Image*> g_imgs(1, nullptr);   // slot 0 unused
static std::vector<uint32_t> g_gen(1, 0);
static std::vector<uint32_t> g_free_slots;
static std::vector<Ctx*> g_ctxs(1, nullptr);
static std::atomic<uint64_t> g_ctx_serial{1};

static uint64_t register_image(Image* im) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    uint32_t slot;
    if (!g_free_slots.empty()) {
        slot = g_free_slots.back();
        g_free_slots.pop_back();
    } else {
        slot = (uint32_t)g_imgs.size();
        g_imgs.push_back(nullptr);
        g_gen.push_back(0);
    }
    g_gen[slot]++;
    g_imgs[slot] = im;
    im->handle = ((uint64_t)g_gen[slot] << 32) | slot;
    return im->handle;
}

Image* get_img(vx_img h) {
    uint32_t slot = (uint32_t)(h & 0xffffffffu), gen = (uint32_t)(h >> 32);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (slot == 0 || slot >= g_imgs.size() || g_gen[slot] != gen) return nullptr;
    return g_imgs[slot];
}

Image* pin_img(vx_img h) {
    uint32_t slot = (uint32_t)(h & 0xffffffffu), gen = (uint32_t)(h >> 32);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (slot == 0 || slot >= g_imgs.size() || g_gen[slot] != gen || !g_imgs[slot]) return nullptr;
    Image* im = g_imgs[slot];
    // the last reference may be going away right now on another thread (rc already 0, the image
    // still registered until drop() takes this lock): such a handle is dead
    int rc = im->rc.load();
    while (rc > 0 && !im->rc.compare_exchange_weak(rc, rc + 1)) {
    }
    return rc > 0 ? im : nullptr;
}

static void unregister_image(Image* im) {
    uint32_t slot = (uint32_t)(im->handle & 0xffffffffu);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    g_imgs[slot] = nullptr;
    g_free_slots.push_back(slot);
}

Ctx* get_ctx(vx_ctx h) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (h == 0 || h >= g_ctxs.size()) return nullptr;
    return g_ctxs[h];
}

// ---- streams ------------------------------------------------------------------
// A host thread keeps one stream index per context for its whole life.  When the thread
// exits (worker threads of a scheduler come and go) its indices return to the context's
// free list, so the next new thread inherits the stream AND the blocks cached on it;
// without this every short-lived thread would start on a cold allocator.
struct ThreadStreams {
    std::vector<std::pair<Ctx*, int>> mine;     // indices this thread owns alone
    std::vector<std::pair<Ctx*, int>> shared;   // indices shared with other threads (never recycled)
    ~ThreadStreams() {
        for (auto& p : mine) {
            std::lock_guard<std::mutex> lk(p.first->mu);
            p.first->free_streams.push_back(p.second);
        }
    }
};

int my_stream_index(Ctx* c) {
    static thread_local ThreadStreams ts;
    for (auto& p : ts.mine)
        if (p.first == c) return p.second;
    for (auto& p : ts.shared)
        if (p.first == c) return p.second;
    int idx;
    bool shared = false;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->free_streams.empty()) {
            idx = c->free_streams.back();
            c->free_streams.pop_back();
        } else {
            const int n = c->next_stream.fetch_add(1);
            shared = n >= kMaxStreams;     // more live threads than streams: share, round robin
            idx = n % kMaxStreams;
        }
        if (!c->streams[idx]) {
            cudaStream_t st = nullptr;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
                (void)cudaGetLastError();
                if (!shared) c->free_streams.push_back(idx);
                return -1;          // reported by the caller; never fall back to stream 0 silently
            }
            c->streams[idx] = st;
        }
    }
    // an index handed out beyond kMaxStreams is shared between threads: it must not be returned
    // to the free list when one of them exits
    if (!shared) ts.mine.emplace_back(c, idx);
    else ts.shared.emplace_back(c, idx);
    return idx;
}

static cudaEvent_t get_event(Ctx* c) {
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->event_pool.empty()) {
            cudaEvent_t e = c->event_pool.back();
            c->event_pool.pop_back();
            return e;
        }
    }
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}
static void put_event(Ctx* c, cudaEvent_t e) {
    if (!e) return;
    std::lock_guard<std::mutex> lk(c->mu);
    c->event_pool.push_back(e);
}

// ---- device memory: caching sub-allocator over the stream-ordered pool ---------------
static size_t round_block(size_t bytes) {
    const size_t g = bytes < (1u << 20) ? 512 : (64u << 10);
    return (bytes + g - 1) / g * g;
}

// Gives every cached block back to the pool and waits for the frees to take effect, so that the
// retry that follows can actually reuse the memory.
static void release_cache_and_wait(Ctx* c) {
    dev_cache_release(c, 0);
    for (int i = 0; i < kMaxStreams; i++)
        if (c->streams[i]) cudaStreamSynchronize(c->streams[i]);
    (void)cudaGetLastError();
}

int dev_alloc(Ctx* c, int s, size_t bytes, void** out) {
    const size_t sz = round_block(bytes ? bytes : 1);
    bool over_budget = false;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        auto it = c->cache[s].find(sz);
        if (it != c->cache[s].end() && !it->second.empty()) {
            *out = it->second.back();
            it->second.pop_back();
            c->cached_bytes -= sz;
            c->alloc_bytes += sz;
            return VX_OK;   // cache -> in use: the sum (and its high-water mark) is unchanged
        }
        over_budget = c->budget && c->alloc_bytes + c->cached_bytes + sz > c->budget;
    }
    bool retried = false;
    if (over_budget) {
        // the budget counts cached blocks: hand back as many as it takes to make room (stream-ordered
        // frees; a block freed on the allocating stream is reusable at once, so no host wait here)
        uint64_t keep = 0;
        {
            std::lock_guard<std::mutex> lk(c->mu);
            if (c->alloc_bytes + sz > c->budget)
                return fail(VX_E_OOM, "allocation of %zu bytes exceeds the memory budget (%llu bytes live, budget %llu)",
                            sz, (unsigned long long)c->alloc_bytes, (unsigned long long)c->budget);
            keep = c->budget - c->alloc_bytes - sz;
        }
        dev_cache_release(c, keep);
        retried = true;
    }
    cudaError_t e = cudaMallocAsync(out, sz, c->streams[s]);
    if (e == cudaErrorMemoryAllocation) {  // give every cached block back, wait for the frees, retry once
        (void)cudaGetLastError();
        release_cache_and_wait(c);
        retried = true;
        e = cudaMallocAsync(out, sz, c->streams[s]);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? VX_E_OOM : VX_E_CUDA,
                    "cudaMallocAsync(%zu bytes) failed: %s", sz, cudaGetErrorString(e));
    }
    if (retried) c->oom_retries.fetch_add(1);
    std::lock_guard<std::mutex> lk(c->mu);
    c->block_size[*out] = sz;
    c->alloc_bytes += sz;
    if (c->alloc_bytes + c->cached_bytes > c->peak_bytes) c->peak_bytes = c->alloc_bytes + c->cached_bytes;
    return VX_OK;
}

void dev_free(Ctx* c, int s, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->block_size.find(p);
    if (it == c->block_size.end()) return;
    const size_t sz = it->second;
    c->alloc_bytes -= sz;
    if (c->cached_bytes + sz > c->cache_limit) {  // over budget: hand it back to the pool
        c->block_size.erase(it);
        cudaFreeAsync(p, c->streams[s]);
        return;
    }
    c->cache[s][sz].push_back(p);
    c->cached_bytes += sz;
}

void dev_cache_release(Ctx* c, uint64_t keep_bytes) {
    std::lock_guard<std::mutex> lk(c->mu);
    for (int s = 0; s < kMaxStreams; s++)
        for (auto& kv : c->cache[s])
            while (!kv.second.empty() && c->cached_bytes > keep_bytes) {
                void* p = kv.second.back();
                kv.second.pop_back();
                c->cached_bytes -= kv.first;
                c->block_size.erase(p);
                cudaFreeAsync(p, c->streams[s]);
            }
}

// ---- images -------------------------------------------------------------------
static size_t dtype_size(int dtype) {
    switch (dtype) {
        case VX_U8: return 1;
        case VX_F32: return 4;
        case VX_U32: return 4;
    }
    return 0;
}

static int new_image_impl(Ctx* c, int dtype, int nx, int ny, int nz, int nb, const float* sp, bool as_bits, Image** out,
                          bool deferred = false) {
    size_t es = dtype_size(dtype);
    VX_REQUIRE(es != 0, VX_E_DTYPE, "unknown dtype %d", dtype);
    VX_REQUIRE(nx > 0 && ny > 0 && nz > 0 && nb > 0, VX_E_INVALID_ARG,
               "bad shape nx=%d ny=%d nz=%d nb=%d", nx, ny, nz, nb);
    VX_REQUIRE((size_t)nx * ny * nz < ((size_t)1 << 31), VX_E_UNSUPPORTED,
               "a single volume must have fewer than 2^31 voxels (slab-decompose it)");
    int s = my_stream_index(c);
    VX_REQUIRE(s >= 0, VX_E_CUDA, "could not create a CUDA stream for the calling thread");
    Image* im = new Image();
    im->dtype = dtype;
    im->nx = nx; im->ny = ny; im->nz = nz; im->nb = nb;
    if (sp) { im->sp[0] = sp[0]; im->sp[1] = sp[1]; im->sp[2] = sp[2]; }
    im->bytes = es * im->count();
    im->ctx = c;
    im->home = s;
    // kernels read and write whole 16-element groups: pad the allocation so that the
    // last group of the batch stays inside it (the padding is never data)
    size_t alloc = ((im->count() + 15) / 16 * 16) * es + 64;
    int arc = deferred ? VX_OK
              : as_bits ? dev_alloc(c, s, bits_alloc_bytes(im->count()), (void**)&im->bits) : dev_alloc(c, s, alloc, &im->d);
    if (arc != VX_OK) {
        delete im;
        return arc;
    }
    im->ready = get_event(c);
    c->live_bytes += deferred ? 0 : as_bits ? im->bit_words() * 4 : im->bytes;
    c->live_handles += 1;
    register_image(im);
    *out = im;
    return VX_OK;
}
int new_image(Ctx* c, int dtype, int nx, int ny, int nz, int nb, const float* sp, Image** out) {
    return new_image_impl(c, dtype, nx, ny, nz, nb, sp, false, out);
}
int new_bits_image(Ctx* c, int nx, int ny, int nz, int nb, const float* sp, Image** out) {
    return new_image_impl(c, VX_U8, nx, ny, nz, nb, sp, true, out);
}
int new_deferred_image(Ctx* c, int nx, int ny, int nz, int nb, const float* sp, Image** out) {
    return new_image_impl(c, VX_F32, nx, ny, nz, nb, sp, false, out, true);
}

int use_input(Ctx* c, Image* img, int s) {
    if (img->home != s) VX_CUDA(cudaStreamWaitEvent(c->streams[s], img->ready, 0));
    if (img->dtype == VX_U8 || img->lazy) {   // a representation added later by another stream
        std::lock_guard<std::mutex> lk(img->mu);
        if (img->alt_ready && img->alt_stream != s) VX_CUDA(cudaStreamWaitEvent(c->streams[s], img->alt_ready, 0));
    }
    return VX_OK;
}

int done_input(Ctx* c, Image* img, int s) {
    if (img->home == s) return VX_OK;
    // a foreign stream read the image: remember when it finishes so that the free
    // (ordered on the home stream) cannot overtake it
    std::lock_guard<std::mutex> lk(img->mu);
    for (auto& p : img->foreign)
        if (p.first == s) {
            VX_CUDA(cudaEventRecord(p.second, c->streams[s]));
            return VX_OK;
        }
    cudaEvent_t e = get_event(c);
    VX_CUDA(cudaEventRecord(e, c->streams[s]));
    img->foreign.emplace_back(s, e);
    return VX_OK;
}

int publish(Ctx* c, Image* img, int s) {
    VX_CUDA(cudaEventRecord(img->ready, c->streams[s]));
    return VX_OK;
}

void drop(Image* im) {
    if (im->rc.fetch_sub(1) != 1) return;
    Ctx* c = im->ctx;
    unregister_image(im);
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
    cudaStream_t st = c->streams[im->home];
    for (auto& p : im->foreign) {
        cudaStreamWaitEvent(st, p.second, 0);
        put_event(c, p.second);
    }
    uint64_t held = 0;
    if (im->d) { dev_free(c, im->home, im->d); held += im->bytes; }
    if (im->bits) { dev_free(c, im->home, im->bits); held += im->bit_words() * 4; }
    if (im->q16) dev_free(c, im->home, im->q16);
    if (im->lazy) {
        if (im->lazy->table) dev_free(c, im->home, im->lazy->table);
        if (im->lazy->src) { im->lazy->src->part_users.fetch_sub(1); drop(im->lazy->src); }
        if (im->lazy->mask) { im->lazy->mask->part_users.fetch_sub(1); drop(im->lazy->mask); }
        delete im->lazy;
    }
    if (im->alt_ready) put_event(c, im->alt_ready);
    put_event(c, im->ready);
    c->live_bytes -= held;
    c->live_handles -= 1;
    if (prev != c->device && prev >= 0) cudaSetDevice(prev);
    delete im;
}

int scratch_alloc(Ctx* c, int s, size_t bytes, void** out) { return dev_alloc(c, s, bytes, out); }
int scratch_free(Ctx* c, int s, void* p) {
    dev_free(c, s, p);
    return VX_OK;
}

// ---- TMA tensor maps ------------------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query: the
// library links the static CUDA runtime only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static std::mutex mu;
    static EncodeTiledFn fn = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

int tma_encode_u8_4d(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides[3],
                     const uint32_t box[4]) {
    EncodeTiledFn fn = encode_tiled_fn();
    VX_REQUIRE(fn != nullptr, VX_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VX_REQUIRE(r == CUDA_SUCCESS, VX_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu x %llu x %llu)",
               (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
               (unsigned long long)dims[3]);
    return VX_OK;
}

int same_shape(const Image* a, const Image* b) {
    if (a->nx != b->nx || a->ny != b->ny || a->nz != b->nz || a->nb != b->nb)
        return fail(VX_E_SHAPE_MISMATCH, "shape mismatch: %dx%dx%dx%d vs %dx%dx%dx%d", a->nx,
                    a->ny, a->nz, a->nb, b->nx, b->ny, b->nz, b->nb);
    return VX_OK;
}

// ---- launch accounting ---------------------------------------------------------
static cudaEvent_t get_timing_event(Ctx* c) {
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->timing_pool.empty()) {
            cudaEvent_t e = c->timing_pool.back();
            c->timing_pool.pop_back();
            return e;
        }
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

LaunchScope::LaunchScope(Ctx* c_, int s_, const char* name_) : c(c_), s(s_), name(name_) {
    c->launches.fetch_add(1, std::memory_order_relaxed);
    if (c->prof_on.load(std::memory_order_relaxed)) {
        e0 = get_timing_event(c);
        cudaEventRecord(e0, c->streams[s]);
    }
}
LaunchScope::~LaunchScope() {
    if (e0) {
        cudaEvent_t e1 = get_timing_event(c);
        cudaEventRecord(e1, c->streams[s]);
        std::lock_guard<std::mutex> lk(c->mu);
        c->prof.push_back({name, e0, e1});
    }
}

int check_launch(const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(VX_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
    return VX_OK;
}

// ---- OpGuard ----------------------------------------------------------------
int OpGuard::begin(vx_ctx ctx) {
    c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    g_cur_ctx = c;
    if (c->fault.load())
        return fail(VX_E_CUDA, "the context is dead after an earlier CUDA fault: %s", c->fault_msg);
    VX_CUDA(cudaSetDevice(c->device));
    s = my_stream_index(c);
    VX_REQUIRE(s >= 0, VX_E_CUDA, "could not create a CUDA stream for the calling thread");
    st = c->streams[s];
    return VX_OK;
}
OpGuard::~OpGuard() {
    for (void* p : temps) dev_free(c, s, p);
    for (Image* im : ins) drop(im);   // the pins taken by input()
    for (Image* im : metas) drop(im);
}
int OpGuard::input(vx_img h, Image** out, int want_dtype) {
    Image* im = pin_img(h);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle %llu",
               (unsigned long long)h);
    ins.push_back(im);   // pinned from here on, released by the destructor on every path
    VX_REQUIRE(im->ctx == c, VX_E_INVALID_ARG, "image belongs to another context");
    VX_REQUIRE(want_dtype == 0 || im->dtype == want_dtype, VX_E_DTYPE,
               "operand has dtype %d, expected %d", im->dtype, want_dtype);
    VX_TRY(use_input(c, im, s));
    if (im->dtype == VX_U8) VX_TRY(ensure_u8(c, im, s));
    if (im->lazy) VX_TRY(ensure_f32(c, im, s));
    *out = im;
    return VX_OK;
}
__global__ void __launch_bounds__(256) pack_stream_kernel(const uint8_t* __restrict__ a, unsigned* __restrict__ bits, size_t n);
__global__ void __launch_bounds__(256) unpack_stream_kernel(const unsigned* __restrict__ bits, uint8_t* __restrict__ a, size_t n);

// The second representation of a boolean image, made on the calling stream after the image's
// producer (use_input has already ordered this stream behind it).  Other streams wait for
// alt_ready; the image's free waits for this stream through done_input like for any reader.
// (called with im->mu held) the representation was there already: if another stream made it, this
// stream has to be behind that -- use_input's wait may have run before it was recorded
static int alt_wait(Ctx* c, Image* im, int s) {
    if (im->alt_ready && im->alt_stream != s) VX_CUDA(cudaStreamWaitEvent(c->streams[s], im->alt_ready, 0));
    return VX_OK;
}
static int alt_done(Ctx* c, Image* im, int s) {
    if (!im->alt_ready) im->alt_ready = get_event(c);
    im->alt_stream = s;
    VX_CUDA(cudaEventRecord(im->alt_ready, c->streams[s]));
    return VX_OK;
}
int OpGuard::input_f32_deferred(vx_img h, Image** out, const void** arr) {
    Image* im = pin_img(h);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle %llu", (unsigned long long)h);
    ins.push_back(im);
    VX_REQUIRE(im->ctx == c, VX_E_INVALID_ARG, "image belongs to another context");
    VX_REQUIRE(im->dtype == VX_F32, VX_E_DTYPE, "operand has dtype %d, expected %d", im->dtype, VX_F32);
    VX_TRY(use_input(c, im, s));
    *arr = im->d;
    if (im->lazy) {
        // the array may be appearing on another stream right now: look under the image's lock
        std::lock_guard<std::mutex> lk(im->mu);
        *arr = im->d;
        if (im->d) VX_TRY(alt_wait(c, im, s));
    }
    if (*arr == nullptr) {
        // The image's ready event already lies behind the producers of src and mask.  They are read on
        // this stream now: pin them like operands so that finish() records the read before they can go.
        Image* parts[2] = {im->lazy->src, im->lazy->mask};
        for (Image* p : parts) {
            p->rc.fetch_add(1);
            ins.push_back(p);
        }
    }
    *out = im;
    return VX_OK;
}
// The f32 array of a deferred image, made on the calling stream (use_input has ordered it behind the
// image's producer, which ran behind the producers of src and mask).
int ensure_f32(Ctx* c, Image* im, int s) {
    if (!im->lazy) return VX_OK;
    std::lock_guard<std::mutex> lk(im->mu);
    if (im->d) return alt_wait(c, im, s);
    void* p = nullptr;
    VX_TRY(dev_alloc(c, s, ((im->count() + 15) / 16 * 16) * 4 + 64, &p));
    int rc = pq_apply_launch(c, s, im->lazy->src, im->lazy->mask->bits, im->lazy->table, (float*)p);
    if (rc == VX_OK) rc = done_input(c, im->lazy->src, s);
    if (rc == VX_OK) rc = done_input(c, im->lazy->mask, s);
    if (rc != VX_OK) { dev_free(c, s, p); return rc; }
    im->d = p;
    c->live_bytes += im->bytes;
    return alt_done(c, im, s);
}
int ensure_u8(Ctx* c, Image* im, int s) {
    if (im->dtype != VX_U8) return VX_OK;
    std::lock_guard<std::mutex> lk(im->mu);
    if (im->d) return alt_wait(c, im, s);
    const size_t n = im->count();
    void* p = nullptr;
    VX_TRY(dev_alloc(c, s, ((n + 15) / 16 * 16) + 64, &p));
    {
        VX_LAUNCH(c, s, "unpack_stream_kernel");
        unpack_stream_kernel<<<grid_for((n + 31) / 32, 256, 8), 256, 0, c->streams[s]>>>(im->bits, (uint8_t*)p, n);
    }
    int rc = check_launch("unpack_stream_kernel");
    if (rc != VX_OK) { dev_free(c, s, p); return rc; }
    im->d = p;
    c->live_bytes += im->bytes;
    return alt_done(c, im, s);
}
static int pack_bits_of(Ctx* c, int s, const Image* im, unsigned* dst) {
    const size_t n = im->count();
    {
        VX_LAUNCH(c, s, "pack_stream_kernel");
        pack_stream_kernel<<<grid_for((n + 31) / 32, 256, 8), 256, 0, c->streams[s]>>>((const uint8_t*)im->d, dst, n);
    }
    return check_launch("pack_stream_kernel");
}
int ensure_bits(Ctx* c, Image* im, int s) {
    if (im->dtype != VX_U8) return fail(VX_E_DTYPE, "only boolean images have a bit form");
    std::lock_guard<std::mutex> lk(im->mu);
    if (im->bits) return alt_wait(c, im, s);
    void* p = nullptr;
    VX_TRY(dev_alloc(c, s, bits_alloc_bytes(im->count()), &p));
    int rc = pack_bits_of(c, s, im, (unsigned*)p);
    if (rc != VX_OK) { dev_free(c, s, p); return rc; }
    im->bits = (unsigned*)p;
    c->live_bytes += im->bit_words() * 4;
    return alt_done(c, im, s);
}
int OpGuard::input_meta(vx_img h, Image** out) {
    Image* im = pin_img(h);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle %llu", (unsigned long long)h);
    metas.push_back(im);
    VX_REQUIRE(im->ctx == c, VX_E_INVALID_ARG, "image belongs to another context");
    *out = im;
    return VX_OK;
}
int OpGuard::input_bits(vx_img h, Image** out, const unsigned** bits) {
    Image* im = pin_img(h);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle %llu", (unsigned long long)h);
    ins.push_back(im);
    VX_REQUIRE(im->ctx == c, VX_E_INVALID_ARG, "image belongs to another context");
    VX_REQUIRE(im->dtype == VX_U8, VX_E_DTYPE, "operand has dtype %d, expected %d", im->dtype, VX_U8);
    VX_TRY(use_input(c, im, s));
    if (im->bits_frozen) {
        // the byte form may be rewritten through a raw pointer the caller holds: pack a private copy now
        void* p = nullptr;
        VX_TRY(dev_alloc(c, s, bits_alloc_bytes(im->count()), &p));
        temps.push_back(p);
        VX_TRY(pack_bits_of(c, s, im, (unsigned*)p));
        *bits = (const unsigned*)p;
    } else {
        VX_TRY(ensure_bits(c, im, s));
        *bits = im->bits;
    }
    *out = im;
    return VX_OK;
}
int OpGuard::finish(Image* out_img, vx_img* out) {
    for (Image* im : ins) VX_TRY(done_input(c, im, s));
    VX_TRY(publish(c, out_img, s));
    *out = out_img->handle;
    return VX_OK;
}
int OpGuard::finish_scalar() {
    for (Image* im : ins) VX_TRY(done_input(c, im, s));
    return VX_OK;
}

// ---- transfer kernels ---------------------------------------------------------
// raw 16-bit voxels -> f32 with the file's slope/intercept, one rounding per operation (the host
// loader computes raw.astype(f32) * slope + inter the same way); 8 voxels per thread
// qrange[2 v] / qrange[2 v + 1] receive the smallest / largest order-preserving code of volume v
// (u = q + 32768 for int16, q for uint16; initialised to 65535 / 0 by the caller).  grid (blocks, nb).
template <typename T>
__global__ void __launch_bounds__(256)
convert16_kernel(const T* __restrict__ raw_all, float* __restrict__ out_all, size_t nvox, float slope, float inter,
                 int* __restrict__ qrange) {
    constexpr int bias = std::is_signed<T>::value ? 32768 : 0;
    const size_t vol = blockIdx.y;
    const T* raw = raw_all + vol * nvox;
    float* out = out_all + vol * nvox;
    int lo_u = 65535, hi_u = 0;
    if ((nvox & 7) == 0) {   // every volume starts 16-byte aligned: 8 voxels per thread and trip
        const size_t ng = nvox >> 3;
        for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ng; g += (size_t)gridDim.x * 256) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(raw) + g);
            const unsigned ww[4] = {w.x, w.y, w.z, w.w};
            float f[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const T lo = (T)(ww[i] & 0xffffu), hi = (T)(ww[i] >> 16);
                f[2 * i] = __fadd_rn(__fmul_rn((float)lo, slope), inter);
                f[2 * i + 1] = __fadd_rn(__fmul_rn((float)hi, slope), inter);
                lo_u = min(lo_u, min((int)lo, (int)hi) + bias);
                hi_u = max(hi_u, max((int)lo, (int)hi) + bias);
            }
            float4* o = reinterpret_cast<float4*>(out) + 2 * g;
            o[0] = make_float4(f[0], f[1], f[2], f[3]);
            o[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (size_t)gridDim.x * 256) {
            const T q = raw[i];
            out[i] = __fadd_rn(__fmul_rn((float)q, slope), inter);
            lo_u = min(lo_u, (int)q + bias);
            hi_u = max(hi_u, (int)q + bias);
        }
    }
    lo_u = __reduce_min_sync(0xffffffffu, lo_u);
    hi_u = __reduce_max_sync(0xffffffffu, hi_u);
    if ((threadIdx.x & 31) == 0 && lo_u <= hi_u) {
        atomicMin(qrange + 2 * vol, lo_u);
        atomicMax(qrange + 2 * vol + 1, hi_u);
    }
}

// Is q -> fl(fl((float)q * slope) + inter) strictly monotone over every code of the type (as the
// percentile ranks order values: -0.0 = +0.0, everything finite)?  65536 evaluations; the answer
// for the last (dtype, slope, inter) is remembered, a loader sends the same scaling for every file.
static int quant_monotone(int host_dtype, float slope, float inter) {   // +1 increasing, -1 decreasing, 0 neither
    static std::mutex mu;
    static int c_dtype = 0, c_ans = 0;
    static float c_slope = 0.f, c_inter = 0.f;
    std::lock_guard<std::mutex> lk(mu);
    if (c_dtype == host_dtype && memcmp(&c_slope, &slope, 4) == 0 && memcmp(&c_inter, &inter, 4) == 0) return c_ans;
    const int q0 = host_dtype == VX_I16 ? -32768 : 0;
    auto val = [&](int q) {
        volatile float p = (float)q * slope;   // two roundings, as the device computes it (no FMA)
        volatile float v = p + inter;
        return (float)v + 0.0f;
    };
    int ans = 0;
    float prev = val(q0);
    bool inc = true, dec = true, finite = std::isfinite(prev);
    for (int i = 1; i < 65536 && finite && (inc || dec); i++) {
        const float v = val(q0 + i);
        finite = std::isfinite(v);
        inc = inc && v > prev;
        dec = dec && v < prev;
        prev = v;
    }
    if (finite) ans = inc ? 1 : dec ? -1 : 0;
    c_dtype = host_dtype; c_slope = slope; c_inter = inter; c_ans = ans;
    return ans;
}

// bit stream -> boolean bytes, 32 voxels per thread (one word in, two 16-byte stores; the byte
// buffer is padded, so the last group is written whole)
__global__ void __launch_bounds__(256)
unpack_stream_kernel(const unsigned* __restrict__ bits, uint8_t* __restrict__ a, size_t n) {
    const size_t nw = (n + 31) >> 5;
    for (size_t w = (size_t)blockIdx.x * 256 + threadIdx.x; w < nw; w += (size_t)gridDim.x * 256) {
        const unsigned v = __ldg(bits + w);
        unsigned o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = (((v >> (4 * i)) & 0xfu) * 0x00204081u) & 0x01010101u;
        uint4* dst = reinterpret_cast<uint4*>(a) + 2 * w;
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// boolean image -> bit stream, 32 voxels per thread (two 16-byte loads, one word out)
__global__ void __launch_bounds__(256)
pack_stream_kernel(const uint8_t* __restrict__ a, unsigned* __restrict__ bits, size_t n) {
    const size_t nw = (n + 31) >> 5;
    for (size_t w = (size_t)blockIdx.x * 256 + threadIdx.x; w < nw; w += (size_t)gridDim.x * 256) {
        unsigned out = 0;
        if ((w << 5) + 32 <= n) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(a) + 2 * w);
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(a) + 2 * w + 1);
            const unsigned ww[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                // bytes are 0/1 (any non-zero counts): gather the four low bits of the word
                unsigned t = ww[i];
                t = (t | (t >> 1) | (t >> 2) | (t >> 3) | (t >> 4) | (t >> 5) | (t >> 6) | (t >> 7)) & 0x01010101u;
                out |= ((t * 0x01020408u) >> 24 & 0xfu) << (4 * i);
            }
        } else {
            for (int i = 0; i < 32; i++)
                if ((w << 5) + i < n && a[(w << 5) + i]) out |= 1u << i;
        }
        bits[w] = out;
    }
}

cudaError_t h2d_copy(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes < (1u << 20) || !c->ordered_uploads.load(std::memory_order_relaxed))
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st);
    std::lock_guard<std::mutex> lk(c->h2d_mu);
    cudaError_t e = cudaSuccess;
    if (c->h2d_tail >= 0) e = cudaStreamWaitEvent(st, c->h2d_ev[c->h2d_tail], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st);
    if (e != cudaSuccess) return e;
    const int i = c->h2d_next;
    if (!c->h2d_ev[i]) e = cudaEventCreateWithFlags(&c->h2d_ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(c->h2d_ev[i], st);
    if (e != cudaSuccess) return e;
    c->h2d_tail = i;
    c->h2d_next = (i + 1) % 8;
    return cudaSuccess;
}

}  // namespace vx

using namespace vx;

// =============================================================== C ABI: lifecycle
extern "C" {

int32_t vx_version(void) { return 100; }
const char* vx_last_error(void) { return g_err; }

int32_t vx_init(int32_t device, vx_ctx* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return fail(VX_E_CUDA, "no CUDA device available (%s); libvoxcuda has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    VX_REQUIRE(device >= 0 && device < n, VX_E_INVALID_ARG, "device %d out of range [0,%d)",
               device, n);
    VX_CUDA(cudaSetDevice(device));
    Ctx* c = new Ctx();
    c->device = device;
    VX_CUDA(cudaDeviceGetDefaultMemPool(&c->pool, device));
    uint64_t never = UINT64_MAX;
    VX_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &never));
    size_t free_b = 0, total_b = 0;
    VX_CUDA(cudaMemGetInfo(&free_b, &total_b));
    c->cache_limit = total_b / 2;
    {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        c->id = g_ctxs.size();
        g_ctxs.push_back(c);
    }
    *out = c->id;
    return VX_OK;
}

int32_t vx_shutdown(vx_ctx ctx) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    VX_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < kMaxStreams; i++)
        if (c->streams[i]) cudaStreamSynchronize(c->streams[i]);
    // images still alive keep pointing at the context; the context object itself is
    // retained (leaked) so late vx_release calls from finalisers stay safe.
    if (c->flush_buf) { cudaFree(c->flush_buf); c->flush_buf = nullptr; }
    return VX_OK;
}

int32_t vx_sync(vx_ctx ctx) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    VX_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < kMaxStreams; i++)
        if (c->streams[i]) VX_CUDA(cudaStreamSynchronize(c->streams[i]));
    return VX_OK;
}

int32_t vx_upload_batch(vx_ctx ctx, const void* host, int32_t dtype, int32_t nx, int32_t ny,
                        int32_t nz, int32_t nb, const float* spacing, vx_img* out) {
    VX_REQUIRE(host != nullptr && out != nullptr, VX_E_INVALID_ARG, "NULL argument");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(new_image(g.c, dtype, nx, ny, nz, nb, spacing, &im));
    cudaError_t e = h2d_copy(g.c, im->d, host, im->bytes, g.st);
    if (e != cudaSuccess) {
        drop(im);
        return fail(VX_E_CUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    return g.finish(im, out);
}

int32_t vx_upload(vx_ctx ctx, const void* host, int32_t dtype, int32_t nx, int32_t ny, int32_t nz,
                  const float* spacing, vx_img* out) {
    return vx_upload_batch(ctx, host, dtype, nx, ny, nz, 1, spacing, out);
}

int32_t vx_upload_scaled(vx_ctx ctx, const void* host, int32_t host_dtype, float slope, float inter,
                         int32_t nx, int32_t ny, int32_t nz, int32_t nb, const float* spacing, vx_img* out) {
    VX_REQUIRE(host != nullptr && out != nullptr, VX_E_INVALID_ARG, "NULL argument");
    VX_REQUIRE(host_dtype == VX_I16 || host_dtype == VX_U16, VX_E_DTYPE,
               "vx_upload_scaled takes VX_I16 or VX_U16 voxels, not dtype %d", host_dtype);
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(new_image(g.c, VX_F32, nx, ny, nz, nb, spacing, &im));
    const size_t n = im->count();
    // the 16-bit codes stay on the device with the image (the quantisation tag of vx_internal.h):
    // 2 more bytes per voxel buy operators that rank values a counting pass instead of a sort
    void* raw = nullptr;
    const size_t code_bytes = ((n + 7) / 8 * 8) * 2;
    int rc = dev_alloc(g.c, im->home, code_bytes + 8 * (size_t)nb, &raw);
    int* qrange = rc == VX_OK ? reinterpret_cast<int*>(static_cast<char*>(raw) + code_bytes) : nullptr;
    if (rc == VX_OK) {
        std::vector<int> init(2 * (size_t)nb);
        for (int v = 0; v < nb; v++) { init[2 * v] = 65535; init[2 * v + 1] = 0; }
        cudaError_t e = h2d_copy(g.c, raw, host, n * 2, g.st);
        // (pageable source: the runtime stages it before returning)
        if (e == cudaSuccess) e = cudaMemcpyAsync(qrange, init.data(), 8 * (size_t)nb, cudaMemcpyHostToDevice, g.st);
        if (e != cudaSuccess) rc = fail(VX_E_CUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    if (rc == VX_OK) {
        const size_t nvox = im->nvox();
        const int per_vol = std::max(1, std::min((int)((nvox / 8 + 255) / 256), (kNumSMs * 8 + nb - 1) / nb));
        const dim3 grid(per_vol, nb);
        {
            VX_LAUNCH(g.c, g.s, "convert16_kernel");
            if (host_dtype == VX_I16)
                convert16_kernel<short><<<grid, 256, 0, g.st>>>((const short*)raw, (float*)im->d, nvox, slope, inter, qrange);
            else
                convert16_kernel<unsigned short><<<grid, 256, 0, g.st>>>((const unsigned short*)raw, (float*)im->d, nvox, slope, inter, qrange);
        }
        rc = check_launch("convert16_kernel");
    }
    if (rc != VX_OK) {
        if (raw) dev_free(g.c, im->home, raw);
        drop(im);
        return rc;
    }
    const int mono = quant_monotone(host_dtype, slope, inter);
    if (mono != 0) {
        im->q16 = raw;
        im->q_dtype = host_dtype;
        im->q_increasing = mono > 0;
        im->q_slope = slope;
        im->q_inter = inter;
    } else {
        dev_free(g.c, im->home, raw);   // stream-ordered: recycled after the conversion kernel
    }
    return g.finish(im, out);
}

int32_t vx_download_bits(vx_ctx ctx, vx_img img, void* host, size_t bytes) {
    VX_REQUIRE(host != nullptr, VX_E_INVALID_ARG, "host is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    const unsigned* bits = nullptr;
    VX_TRY(g.input_bits(img, &im, &bits));   // the image's own bit form: results of the bit-space operators have it already
    const size_t n = im->count();
    VX_REQUIRE(bytes == (n + 7) / 8, VX_E_INVALID_ARG, "download size %zu != ceil(%zu voxels / 8) = %zu", bytes, n,
               (n + 7) / 8);
    VX_CUDA(cudaMemcpyAsync(host, bits, bytes, cudaMemcpyDefault, g.st));
    VX_CUDA(cudaStreamSynchronize(g.st));
    return g.finish_scalar();
}

int32_t vx_download(vx_ctx ctx, vx_img img, void* host, size_t bytes) {
    VX_REQUIRE(host != nullptr, VX_E_INVALID_ARG, "host is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(g.input(img, &im, 0));
    VX_REQUIRE(bytes == im->bytes, VX_E_INVALID_ARG, "download size %zu != image size %zu", bytes,
               im->bytes);
    VX_CUDA(cudaMemcpyAsync(host, im->d, bytes, cudaMemcpyDefault, g.st));
    VX_CUDA(cudaStreamSynchronize(g.st));
    return g.finish_scalar();
}

int32_t vx_info(vx_img img, int32_t* dtype, int32_t* nx, int32_t* ny, int32_t* nz, int32_t* nb,
                float* spacing3) {
    Image* im = get_img(img);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle");
    if (dtype) *dtype = im->dtype;
    if (nx) *nx = im->nx;
    if (ny) *ny = im->ny;
    if (nz) *nz = im->nz;
    if (nb) *nb = im->nb;
    if (spacing3) { spacing3[0] = im->sp[0]; spacing3[1] = im->sp[1]; spacing3[2] = im->sp[2]; }
    return VX_OK;
}

int32_t vx_retain(vx_img img) {
    // look-up and increment under the table lock: a racing last vx_release cannot free the image
    // between the two
    Image* im = pin_img(img);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle");
    return VX_OK;
}

int32_t vx_release(vx_img img) {
    Image* im = get_img(img);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle");
    drop(im);
    return VX_OK;
}

int32_t vx_batch_slice(vx_ctx ctx, vx_img img, int32_t first, int32_t count, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(g.input(img, &im, 0));
    VX_REQUIRE(first >= 0 && count > 0 && first + count <= im->nb, VX_E_INVALID_ARG,
               "slice [%d,%d) outside batch of %d", first, first + count, im->nb);
    Image* o = nullptr;
    VX_TRY(new_image(g.c, im->dtype, im->nx, im->ny, im->nz, count, im->sp, &o));
    size_t per = im->bytes / im->nb;
    cudaError_t e = cudaMemcpyAsync(o->d, (const char*)im->d + per * first, per * count,
                                    cudaMemcpyDeviceToDevice, g.st);
    if (e != cudaSuccess) {
        drop(o);
        return fail(VX_E_CUDA, "slice copy failed: %s", cudaGetErrorString(e));
    }
    return g.finish(o, out);
}

// ---- slab helpers: z ranges of every volume of a batch -------------------------------------
int32_t vx_crop_z(vx_ctx ctx, vx_img img, int32_t z0, int32_t nz, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(g.input(img, &im, 0));
    VX_REQUIRE(z0 >= 0 && nz > 0 && z0 + nz <= im->nz, VX_E_INVALID_ARG,
               "planes [%d,%d) outside a volume of %d planes", z0, z0 + nz, im->nz);
    Image* o = nullptr;
    VX_TRY(new_image(g.c, im->dtype, im->nx, im->ny, nz, im->nb, im->sp, &o));
    const size_t es = im->bytes / im->count();
    const size_t plane = (size_t)im->nx * im->ny * es;
    cudaError_t e = cudaMemcpy2DAsync(o->d, plane * nz, (const char*)im->d + plane * z0, plane * im->nz,
                                      plane * nz, im->nb, cudaMemcpyDeviceToDevice, g.st);
    if (e != cudaSuccess) {
        drop(o);
        return fail(VX_E_CUDA, "crop copy failed: %s", cudaGetErrorString(e));
    }
    return g.finish(o, out);
}

int32_t vx_concat_z(vx_ctx ctx, const vx_img* parts, int32_t n_parts, vx_img* out) {
    VX_REQUIRE(out != nullptr && parts != nullptr && n_parts >= 1 && n_parts <= 64, VX_E_INVALID_ARG,
               "bad argument (1..64 parts)");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    std::vector<Image*> ims(n_parts, nullptr);
    int nz = 0;
    for (int i = 0; i < n_parts; i++) {
        VX_TRY(g.input(parts[i], &ims[i], 0));
        VX_REQUIRE(ims[i]->dtype == ims[0]->dtype && ims[i]->nx == ims[0]->nx && ims[i]->ny == ims[0]->ny &&
                       ims[i]->nb == ims[0]->nb,
                   VX_E_SHAPE_MISMATCH, "part %d differs in dtype, nx, ny or nb", i);
        nz += ims[i]->nz;
    }
    Image* o = nullptr;
    VX_TRY(new_image(g.c, ims[0]->dtype, ims[0]->nx, ims[0]->ny, nz, ims[0]->nb, ims[0]->sp, &o));
    const size_t es = o->bytes / o->count();
    const size_t plane = (size_t)o->nx * o->ny * es;
    int z = 0;
    for (int i = 0; i < n_parts; i++) {
        cudaError_t e = cudaMemcpy2DAsync((char*)o->d + plane * z, plane * nz, ims[i]->d, plane * ims[i]->nz,
                                          plane * ims[i]->nz, o->nb, cudaMemcpyDeviceToDevice, g.st);
        if (e != cudaSuccess) {
            drop(o);
            return fail(VX_E_CUDA, "concat copy failed: %s", cudaGetErrorString(e));
        }
        z += ims[i]->nz;
    }
    return g.finish(o, out);
}

// Copy an image into another context (another GPU of the same process): peer-to-peer over
// NVLink when the devices can reach each other, staged by the driver otherwise.
int32_t vx_transfer(vx_ctx dst_ctx, vx_img img, vx_img* out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    Image* im = get_img(img);
    VX_REQUIRE(im != nullptr, VX_E_INVALID_ARG, "invalid image handle");
    OpGuard g;
    VX_TRY(g.begin(dst_ctx));
    Ctx* src = im->ctx;
    Image* o = nullptr;
    if (im->lazy && !im->d) {   // a deferred f32 image travels as its array
        const int ss = my_stream_index(src);
        VX_REQUIRE(ss >= 0, VX_E_CUDA, "could not create a CUDA stream for the calling thread");
        VX_CUDA(cudaSetDevice(src->device));
        VX_TRY(use_input(src, im, ss));
        VX_TRY(ensure_f32(src, im, ss));
        VX_CUDA(cudaSetDevice(g.c->device));
    }
    const bool as_bits = im->d == nullptr;   // a boolean image that only exists as bits travels as bits
    if (as_bits) VX_TRY(new_bits_image(g.c, im->nx, im->ny, im->nz, im->nb, im->sp, &o));
    else VX_TRY(new_image(g.c, im->dtype, im->nx, im->ny, im->nz, im->nb, im->sp, &o));
    // order the copy after the producer of `img` (an event of another device may be waited on)
    cudaError_t e = cudaStreamWaitEvent(g.st, im->ready, 0);
    if (e == cudaSuccess && im->alt_ready) e = cudaStreamWaitEvent(g.st, im->alt_ready, 0);
    if (e == cudaSuccess)
        e = as_bits ? cudaMemcpyPeerAsync(o->bits, g.c->device, im->bits, src->device, im->bit_words() * 4, g.st)
                    : cudaMemcpyPeerAsync(o->d, g.c->device, im->d, src->device, im->bytes, g.st);
    if (e != cudaSuccess) {
        drop(o);
        return fail(VX_E_CUDA, "transfer failed: %s", cudaGetErrorString(e));
    }
    // the source must outlive the copy: the caller keeps its reference until vx_sync(dst_ctx)
    VX_TRY(publish(g.c, o, g.s));
    *out = o->handle;
    return VX_OK;
}

// Interop with other CUDA users of the process (torch.distributed / NCCL in the slab path):
// the raw device pointer of an image and the CUDA stream of the calling thread.  Work
// enqueued by the library on that stream before the call is ordered before anything the
// caller enqueues on it afterwards.
int32_t vx_device_ptr(vx_ctx ctx, vx_img img, void** ptr) {
    VX_REQUIRE(ptr != nullptr, VX_E_INVALID_ARG, "ptr is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(g.input(img, &im, 0));   // makes the producer's work visible to this thread's stream
    if (im->dtype == VX_U8) {
        // the caller may write through the pointer (the slab path receives planes into images): the
        // bytes are the truth from now on, a cached bit form would go stale
        std::lock_guard<std::mutex> lk(im->mu);
        im->bits_frozen = true;
        // (a deferred image that reads these bits keeps them: it holds the mask's value as of its own call)
        if (im->bits && im->part_users.load() == 0) {
            dev_free(g.c, g.s, im->bits);   // stream-ordered behind every reader enqueued so far on this stream
            g.c->live_bytes -= im->bit_words() * 4;
            im->bits = nullptr;
        }
    } else if (im->q16) {
        // same for the 16-bit codes of a quantised image: they stop describing the array
        std::lock_guard<std::mutex> lk(im->mu);
        im->q_frozen = true;
    }
    *ptr = im->d;
    return g.finish_scalar();
}
int32_t vx_stream_handle(vx_ctx ctx, void** stream) {
    VX_REQUIRE(stream != nullptr, VX_E_INVALID_ARG, "stream is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    *stream = (void*)g.st;
    return VX_OK;
}

// ---- peer-visible mailboxes (CUDA IPC) for the slab path ------------------------------------
int32_t vx_ipc_alloc(vx_ctx ctx, size_t bytes, void** ptr, void* handle64) {
    VX_REQUIRE(ptr != nullptr && handle64 != nullptr && bytes > 0, VX_E_INVALID_ARG, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    VX_CUDA(cudaMalloc(ptr, bytes));
    VX_CUDA(cudaMemset(*ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        memset(handle64, 0, 64);   // same-process use (raw pointer) still works
    } else {
        memcpy(handle64, &h, 64);
    }
    return VX_OK;
}
int32_t vx_ipc_open(vx_ctx ctx, const void* handle64, void** ptr) {
    VX_REQUIRE(ptr != nullptr && handle64 != nullptr, VX_E_INVALID_ARG, "bad argument");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    VX_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VX_OK;
}
int32_t vx_ipc_close(vx_ctx ctx, void* ptr) {
    OpGuard g;
    VX_TRY(g.begin(ctx));
    if (ptr) VX_CUDA(cudaIpcCloseMemHandle(ptr));
    return VX_OK;
}
int32_t vx_ipc_free(vx_ctx ctx, void* ptr) {
    OpGuard g;
    VX_TRY(g.begin(ctx));
    if (ptr) VX_CUDA(cudaFree(ptr));
    return VX_OK;
}
// planes [z0, z0+nz) of every volume of `img`, densely packed ([nb][nz][ny][nx]), into raw device
// memory `dst` (e.g. this rank's mailbox), on the calling thread's stream
int32_t vx_copy_planes_to(vx_ctx ctx, vx_img img, int32_t z0, int32_t nz, void* dst) {
    VX_REQUIRE(dst != nullptr, VX_E_INVALID_ARG, "dst is NULL");
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Image* im = nullptr;
    VX_TRY(g.input(img, &im, 0));
    VX_REQUIRE(z0 >= 0 && nz > 0 && z0 + nz <= im->nz, VX_E_INVALID_ARG,
               "planes [%d,%d) outside a volume of %d planes", z0, z0 + nz, im->nz);
    const size_t es = im->bytes / im->count();
    const size_t plane = (size_t)im->nx * im->ny * es;
    VX_CUDA(cudaMemcpy2DAsync(dst, plane * nz, (const char*)im->d + plane * z0, plane * im->nz, plane * nz,
                              im->nb, cudaMemcpyDeviceToDevice, g.st));
    return g.finish_scalar();
}

int32_t vx_host_alloc(size_t bytes, void** out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return VX_OK;
}
int32_t vx_host_alloc_wc(size_t bytes, void** out) {
    VX_REQUIRE(out != nullptr, VX_E_INVALID_ARG, "out is NULL");
    VX_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocWriteCombined));
    return VX_OK;
}
int32_t vx_host_free(void* p) {
    if (p) VX_CUDA(cudaFreeHost(p));
    return VX_OK;
}

int32_t vx_mem_info(vx_ctx ctx, uint64_t* live_bytes, uint64_t* live_handles,
                    uint64_t* pool_reserved) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    if (live_bytes) *live_bytes = c->live_bytes.load();
    if (live_handles) *live_handles = c->live_handles.load();
    if (pool_reserved) {
        uint64_t v = 0;
        VX_CUDA(cudaMemPoolGetAttribute(c->pool, cudaMemPoolAttrReservedMemCurrent, &v));
        *pool_reserved = v;
    }
    return VX_OK;
}

int32_t vx_mem_set_budget(vx_ctx ctx, uint64_t bytes) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    std::lock_guard<std::mutex> lk(c->mu);
    c->budget = bytes;
    c->peak_bytes = c->alloc_bytes + c->cached_bytes;
    return VX_OK;
}

int32_t vx_mem_stats(vx_ctx ctx, uint64_t* in_use_bytes, uint64_t* cached_bytes, uint64_t* peak_bytes,
                     uint64_t* oom_retries) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    std::lock_guard<std::mutex> lk(c->mu);
    if (in_use_bytes) *in_use_bytes = c->alloc_bytes;
    if (cached_bytes) *cached_bytes = c->cached_bytes;
    if (peak_bytes) *peak_bytes = c->peak_bytes;
    if (oom_retries) *oom_retries = c->oom_retries.load();
    return VX_OK;
}

int32_t vx_set_option(vx_ctx ctx, int32_t option, int64_t value) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    switch (option) {
        case VX_OPT_QUANTISED_PATHS: c->quantised_paths.store(value != 0); return VX_OK;
        case VX_OPT_ORDERED_UPLOADS: c->ordered_uploads.store(value != 0); return VX_OK;
    }
    return fail(VX_E_INVALID_ARG, "unknown option %d", option);
}

int32_t vx_mem_trim(vx_ctx ctx, uint64_t keep_bytes) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    VX_CUDA(cudaSetDevice(c->device));
    dev_cache_release(c, keep_bytes);
    VX_TRY(vx_sync(ctx));
    VX_CUDA(cudaMemPoolTrimTo(c->pool, (size_t)keep_bytes));
    return VX_OK;
}

// ------------------------------------------------------------------ measurement
int32_t vx_event_create(vx_ctx ctx, vx_evt* out) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr && out != nullptr, VX_E_INVALID_ARG, "invalid argument");
    VX_CUDA(cudaSetDevice(c->device));
    cudaEvent_t e;
    VX_CUDA(cudaEventCreate(&e));
    *out = (vx_evt)(uintptr_t)e;
    return VX_OK;
}
int32_t vx_event_record(vx_ctx ctx, vx_evt ev) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr && ev != 0, VX_E_INVALID_ARG, "invalid argument");
    VX_CUDA(cudaSetDevice(c->device));
    int s = my_stream_index(c);
    VX_CUDA(cudaEventRecord((cudaEvent_t)(uintptr_t)ev, c->streams[s]));
    return VX_OK;
}
int32_t vx_event_sync(vx_evt ev) {
    VX_REQUIRE(ev != 0, VX_E_INVALID_ARG, "invalid event");
    VX_CUDA(cudaEventSynchronize((cudaEvent_t)(uintptr_t)ev));
    return VX_OK;
}
int32_t vx_event_elapsed_ms(vx_evt start, vx_evt stop, float* ms) {
    VX_REQUIRE(start != 0 && stop != 0 && ms != nullptr, VX_E_INVALID_ARG, "invalid argument");
    VX_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)(uintptr_t)start, (cudaEvent_t)(uintptr_t)stop));
    return VX_OK;
}
int32_t vx_event_destroy(vx_evt ev) {
    if (ev) VX_CUDA(cudaEventDestroy((cudaEvent_t)(uintptr_t)ev));
    return VX_OK;
}

int32_t vx_launch_count(vx_ctx ctx, uint64_t* out) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr && out != nullptr, VX_E_INVALID_ARG, "invalid argument");
    *out = c->launches.load();
    return VX_OK;
}

int32_t vx_prof_enable(vx_ctx ctx, int32_t on) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    c->prof_on.store(on ? 1 : 0);
    return VX_OK;
}

static int prof_drain(Ctx* c) {
    std::vector<ProfRec> recs;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        recs.swap(c->prof);
    }
    for (auto& r : recs) {
        float ms = 0.f;
        cudaEventSynchronize(r.e1);
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            std::lock_guard<std::mutex> lk(c->mu);
            auto& acc = c->prof_acc[r.name];
            acc.first += 1;
            acc.second += ms;
        }
        std::lock_guard<std::mutex> lk(c->mu);
        c->timing_pool.push_back(r.e0);
        c->timing_pool.push_back(r.e1);
    }
    return VX_OK;
}

int32_t vx_prof_reset(vx_ctx ctx) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr, VX_E_INVALID_ARG, "invalid context handle");
    VX_CUDA(cudaSetDevice(c->device));
    prof_drain(c);
    std::lock_guard<std::mutex> lk(c->mu);
    c->prof_acc.clear();
    return VX_OK;
}

int32_t vx_prof_report(vx_ctx ctx, char* buf, size_t buf_bytes) {
    Ctx* c = get_ctx(ctx);
    VX_REQUIRE(c != nullptr && buf != nullptr && buf_bytes > 0, VX_E_INVALID_ARG,
               "invalid argument");
    VX_CUDA(cudaSetDevice(c->device));
    prof_drain(c);
    std::lock_guard<std::mutex> lk(c->mu);
    size_t off = 0;
    buf[0] = 0;
    for (auto& kv : c->prof_acc) {
        int n = snprintf(buf + off, buf_bytes - off, "%s %llu %.6f\n", kv.first.c_str(),
                         (unsigned long long)kv.second.first, kv.second.second);
        if (n < 0 || (size_t)n >= buf_bytes - off) break;
        off += n;
    }
    return VX_OK;
}

int32_t vx_flush_l2(vx_ctx ctx) {
    OpGuard g;
    VX_TRY(g.begin(ctx));
    Ctx* c = g.c;
    if (!c->flush_buf) {
        c->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
        VX_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    VX_CUDA(cudaMemsetAsync(c->flush_buf, 0, c->flush_bytes, g.st));
    return VX_OK;
}

}  // extern "C"
