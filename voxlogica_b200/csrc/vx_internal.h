// Internal runtime of libvoxcuda.so: contexts, HBM-resident image handles, the
// stream-ordered memory pool, per-thread streams and the error convention.
// Nothing in this header crosses the C ABI (include/voxcuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/voxcuda.h"

namespace vx {

constexpr int kMaxStreams = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

struct Ctx;

// T1: a device-resident batch of volumes.
struct Image {
    // Boolean images (dtype VX_U8) have two interchangeable device representations: `d`, one byte
    // per voxel (0/1), and `bits`, one BIT per voxel -- a flat little-endian bit stream over the
    // whole batch, bit (i & 31) of word (i >> 5) = voxel i, x fastest (the format of
    // vx_download_bits).  An operator produces whichever suits it and consumers ask for what they
    // read (OpGuard::input / input_bits); the missing one is made on demand, once, and kept.
    // Operators that work in bit space (connected components, distance balls, near/interior, the
    // boolean pointwise programs, volume) therefore never touch the byte form: 1/8 of the traffic.
    void* d = nullptr;
    unsigned* bits = nullptr;
    cudaEvent_t alt_ready = nullptr;   // completion of the representation that was added after publication ...
    int alt_stream = -1;               // ... and the stream index it was made on
    bool bits_frozen = false;          // the byte form was handed out as a raw pointer (vx_device_ptr): never cache bits
    size_t bytes = 0;
    int dtype = 0;
    int nx = 0, ny = 0, nz = 0, nb = 0;
    float sp[3] = {1.f, 1.f, 1.f};
    std::atomic<int> rc{1};
    Ctx* ctx = nullptr;
    int home = 0;                 // stream index the image was produced on
    cudaEvent_t ready = nullptr;  // recorded on `home` after the producer
    std::mutex mu;                // guards `foreign`
    std::vector<std::pair<int, cudaEvent_t>> foreign;  // last use on other streams
    uint64_t handle = 0;
    // Quantisation tag (set by vx_upload_scaled, never inherited): the voxels are
    //     v = fl(fl((float)q * q_slope) + q_inter)
    // for the 16-bit codes q kept in q16 (same voxel order, owned by the image), and the map q -> v
    // was verified on the host to be STRICTLY monotone over the whole code range -- so equal
    // voxels have equal codes and the order of the codes is the order of the voxels.  Operators
    // that only need the ORDER of the values (percentiles) may work on the codes: 65536 possible
    // values make ranking a counting problem instead of a sort.  q16 also holds, after the codes,
    // two ints per volume: its smallest and largest order-preserving code u (q + 32768 for int16).
    void* q16 = nullptr;
    int q_dtype = 0;              // VX_I16 / VX_U16, 0 = no tag
    bool q_increasing = true;
    float q_slope = 1.f, q_inter = 0.f;
    // Deferred f32 form (set by vx_percentiles on a quantised operand): the voxels are
    //     v = mask ? table[volume][u(q)] : 0
    // for the 16-bit codes q of `src` (u = the order-preserving code), the bit form of `mask` and a
    // per-volume table owned by this image.  `d` stays nullptr until somebody needs the f32 array
    // (ensure_f32, through OpGuard::input); the pointwise programs read the three parts directly,
    // 2.125 bytes per voxel instead of writing 4 and reading them back.  src and mask are held by a
    // reference each until this image dies.
    struct Lazy {
        Image* src = nullptr;
        Image* mask = nullptr;
        float* table = nullptr;   // [nb][65536]
        // the table is non-decreasing over the codes of every volume's range taken in value order
        // (0 <= correction <= 1, entries of absent codes filled in): `table[code] OP c` is a comparison
        // of the code's position with a per-volume break point (pointwise programs, vx_pointwise.cu)
        bool monotone = false;
    };
    Lazy* lazy = nullptr;
    std::atomic<int> part_users{0};   // deferred images that read this image's q16 / bits: those stay allocated
    bool q_frozen = false;            // the f32 array was handed out as a raw pointer: the codes no longer describe it
    bool has_codes() const { return q16 != nullptr && !q_frozen; }
    const int* q_range() const { return reinterpret_cast<const int*>(static_cast<const char*>(q16) + ((count() + 7) / 8 * 8) * 2); }
    size_t bit_words() const { return (count() + 31) / 32; }
    size_t nvox() const { return (size_t)nx * ny * nz; }       // per volume
    size_t count() const { return (size_t)nx * ny * nz * nb; }  // whole batch
};

struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
};

struct Ctx {
    int device = 0;
    uint64_t id = 0;
    cudaStream_t streams[kMaxStreams] = {};
    std::atomic<int> next_stream{0};
    std::vector<int> free_streams;  // indices handed back by exited threads (guarded by mu)
    std::mutex mu;  // guards pools and accounting below
    std::vector<cudaEvent_t> event_pool;
    std::vector<cudaEvent_t> timing_pool;  // timing-enabled events for the profiler
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> live_bytes{0};
    std::atomic<uint64_t> live_handles{0};
    std::atomic<int> prof_on{0};
    std::vector<ProfRec> prof;
    std::map<std::string, std::pair<uint64_t, double>> prof_acc;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    cudaMemPool_t pool = nullptr;
    // caching sub-allocator: freed blocks are kept per (stream, rounded size) and handed
    // out again without touching the driver (the operator DAG of a spec repeats sizes)
    std::map<size_t, std::vector<void*>> cache[kMaxStreams];
    std::unordered_map<void*, size_t> block_size;  // every block handed out by dev_alloc
    uint64_t cached_bytes = 0;
    uint64_t cache_limit = 0;  // bytes kept in the free lists before blocks go back to the pool
    uint64_t alloc_bytes = 0;  // bytes of every block handed out by dev_alloc and not yet given back
    uint64_t budget = 0;       // vx_mem_set_budget: soft cap on alloc_bytes + cached_bytes (0 = none)
    std::atomic<int> quantised_paths{1};   // VX_OPT_QUANTISED_PATHS
    uint64_t peak_bytes = 0;   // high-water mark of alloc_bytes + cached_bytes since the last vx_mem_set_budget
    std::atomic<uint64_t> oom_retries{0};   // allocations that succeeded only after the cache was released
    // First sticky CUDA fault seen on this context (illegal address, launch failure, ...): the CUDA
    // context is unusable from then on, and every later call reports the fault instead of
    // whatever error the dead context produces next.
    std::atomic<int> fault{0};
    char fault_msg[256] = "";
    // Host->device link: large uploads issued by different caller threads (streams) are chained, each
    // copy waiting for the one issued before it, so the PCIe link serves them one at a time in
    // arrival order instead of time-slicing -- the step whose upload was asked for first gets
    // its data first (VX_OPT_ORDERED_UPLOADS, on by default).
    std::mutex h2d_mu;
    cudaEvent_t h2d_ev[8] = {};
    int h2d_next = 0, h2d_tail = -1;
    std::atomic<int> ordered_uploads{1};
};
// cudaMemcpyAsync host->device on stream st, chained behind the previous large upload of the context
cudaError_t h2d_copy(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st);
// Classify a CUDA error; sticky ones are latched on the context of the calling thread's current call.
bool sticky_cuda_error(cudaError_t e);
void latch_fault(cudaError_t e, const char* what);

// Device memory for images and scratch, ordered on stream index s.
int dev_alloc(Ctx* c, int s, size_t bytes, void** out);
void dev_free(Ctx* c, int s, void* p);
void dev_cache_release(Ctx* c, uint64_t keep_bytes);

// ---- error convention ------------------------------------------------------
int fail(int code, const char* fmt, ...);
#define VX_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            vx::latch_fault(_e, #call);                                                 \
            return vx::fail(_e == cudaErrorMemoryAllocation ? VX_E_OOM : VX_E_CUDA,     \
                            "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),     \
                            __FILE__, __LINE__);                                        \
        }                                                                               \
    } while (0)
#define VX_TRY(expr)               \
    do {                           \
        int _r = (expr);           \
        if (_r != VX_OK) return _r; \
    } while (0)
#define VX_REQUIRE(cond, code, ...)                    \
    do {                                               \
        if (!(cond)) return vx::fail(code, __VA_ARGS__); \
    } while (0)

// ---- handles ---------------------------------------------------------------
Ctx* get_ctx(vx_ctx h);
Image* get_img(vx_img h);  // nullptr when invalid; the pointer is only safe while a reference is held
// Look the handle up and take a reference in one step (under the table lock): the image cannot be
// freed by a concurrent vx_release between the look-up and the increment.  Pair with drop().
Image* pin_img(vx_img h);
// Stream of the calling thread in this context (created lazily, round-robin).
int my_stream_index(Ctx* c);   // -1 when the stream could not be created
inline cudaStream_t stream_of(Ctx* c, int idx) { return c->streams[idx]; }

// Allocate a new image on the calling thread's stream.  On success *out holds a
// handle with one reference owned by the caller.
int new_image(Ctx* c, int dtype, int nx, int ny, int nz, int nb, const float* sp, Image** out);
// A boolean image that starts life as bits only (im->d == nullptr until somebody needs bytes).
int new_bits_image(Ctx* c, int nx, int ny, int nz, int nb, const float* sp, Image** out);
// Make the missing representation of a boolean image on stream index s (no-op when present).
int ensure_u8(Ctx* c, Image* im, int s);
// Make the f32 array of an image that only has its deferred form (no-op otherwise).
int ensure_f32(Ctx* c, Image* im, int s);
// An f32 image without its array (Image::lazy is filled in by the caller).
int new_deferred_image(Ctx* c, int nx, int ny, int nz, int nb, const float* sp, Image** out);
// vx_percentiles.cu: out = mask ? table[volume][u(q)] : 0 over the whole batch, on stream index s
int pq_apply_launch(Ctx* c, int s, const Image* src, const unsigned* maskbits, const float* table, float* out);
int ensure_bits(Ctx* c, Image* im, int s);
inline size_t bits_alloc_bytes(size_t voxels) { return ((voxels + 127) / 128) * 16 + 64; }   // whole uint4 groups + slack
// Make `img` safe to read on stream index `s` (waits for its producer if needed).
int use_input(Ctx* c, Image* img, int s);
// Called after the consuming kernels were enqueued on stream `s`.
int done_input(Ctx* c, Image* img, int s);
// Called after the producing kernels were enqueued: records the ready event.
int publish(Ctx* c, Image* img, int s);
void drop(Image* img);  // release one reference

// Scratch memory from the stream-ordered pool.
int scratch_alloc(Ctx* c, int s, size_t bytes, void** out);
int scratch_free(Ctx* c, int s, void* p);

// RAII device scratch bound to (context, stream index): every block goes back to the stream's
// free list when the holder leaves scope -- on success and on every early return alike.  Safe
// right after the kernels were enqueued: blocks are recycled on the same stream, in order.
struct Scratch {
    Ctx* c;
    int s;
    std::vector<void*> blocks;
    Scratch(Ctx* c_, int s_) : c(c_), s(s_) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() {
        for (void* p : blocks) dev_free(c, s, p);
    }
    template <typename T>
    int alloc(size_t bytes, T** out) {
        void* p = nullptr;
        int rc = dev_alloc(c, s, bytes, &p);
        if (rc != VX_OK) return rc;
        blocks.push_back(p);
        *out = static_cast<T*>(p);
        return VX_OK;
    }
};

// ---- launch accounting / profiling ----------------------------------------
struct LaunchScope {
    Ctx* c;
    int s;
    const char* name;
    cudaEvent_t e0 = nullptr;
    LaunchScope(Ctx* c, int s, const char* name);
    ~LaunchScope();
};
// Use as: { VX_LAUNCH(c, s, "kernel_name"); kernel<<<...>>>(...); }
#define VX_LAUNCH(c, s, name) vx::LaunchScope _ls_##__LINE__(c, s, name)
int check_launch(const char* name);

// Grid sizing helper: a multiple of the SM count, capped by the work available.
inline int grid_for(size_t work_items, int per_block, int ctas_per_sm) {
    size_t need = (work_items + per_block - 1) / per_block;
    size_t cap = (size_t)kNumSMs * ctas_per_sm;
    if (need == 0) need = 1;
    return (int)(need < cap ? need : cap);
}

// RAII guard that validates common arguments and pins the inputs for one op: every input holds
// one extra reference from input() until the guard dies, so a vx_release from another thread (a
// .NET finaliser) cannot free an image -- or recycle its device block -- under a running call.
struct OpGuard {
    Ctx* c = nullptr;
    int s = 0;
    cudaStream_t st = nullptr;
    std::vector<Image*> ins;
    std::vector<Image*> metas;  // pinned for metadata only
    std::vector<void*> temps;   // scratch made for this call (bit copies of frozen images), freed with the guard
    OpGuard() = default;
    OpGuard(const OpGuard&) = delete;
    OpGuard& operator=(const OpGuard&) = delete;
    ~OpGuard();
    int begin(vx_ctx ctx);
    // A boolean input comes with its byte form (`d`) materialised: kernels written for bytes need
    // nothing else.  input_bits() is for kernels that read the bit form: no bytes are made.
    int input(vx_img h, Image** out, int want_dtype /*0 = any*/);
    // an f32 operand whose deferred form (Image::lazy) is acceptable: no array is made; the parts
    // the deferred form refers to are pinned and ordered like operands of their own
    // (*arr = the f32 array when the image has one, nullptr = read the parts)
    int input_f32_deferred(vx_img h, Image** out, const void** arr);
    int input_bits(vx_img h, Image** out, const unsigned** bits);
    int input_meta(vx_img h, Image** out);   // pinned for its metadata only: no stream ordering, no conversion
    int finish(Image* out_img, vx_img* out);  // publish output + mark inputs used
    int finish_scalar();                       // ops without an image result
};

int same_shape(const Image* a, const Image* b);

}  // namespace vx
