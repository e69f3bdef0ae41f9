/* Plain-C caller of libvoxcuda.so: what a P/Invoke (or any FFI) host does, without Python.
 *   c_smoke cpu : no GPU needed -- the library must load, report its version and refuse to
 *                 create a context (there is no CPU fallback), with a message.
 *   c_smoke gpu : a tiny pipeline on cuda:0 (threshold -> near -> through -> volume, edt),
 *                 checked against values worked out by hand. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../include/voxcuda.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int32_t rc_ = (call);                                                            \
        if (rc_ != VX_OK) {                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, vx_last_error());         \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char** argv) {
    const int gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    if (vx_version() < 100) { fprintf(stderr, "bad version\n"); return 1; }
    vx_ctx ctx = 0;
    if (!gpu) {
        int32_t rc = vx_init(0, &ctx);
        if (rc == VX_OK) { printf("a CUDA device is present; cpu mode has nothing to check\n"); return 0; }
        if (rc != VX_E_CUDA || strlen(vx_last_error()) == 0) { fprintf(stderr, "expected VX_E_CUDA\n"); return 1; }
        if (vx_release(12345) != VX_E_INVALID_ARG) { fprintf(stderr, "expected VX_E_INVALID_ARG\n"); return 1; }
        printf("c_smoke cpu ok: %s\n", vx_last_error());
        return 0;
    }
    CHECK(vx_init(0, &ctx));
    /* 1 x 4 x 16 x 32 ramp: value = x */
    enum { NX = 32, NY = 16, NZ = 4, N = NX * NY * NZ };
    static float img[N];
    for (int i = 0; i < N; i++) img[i] = (float)(i % NX);
    vx_img f = 0, lo = 0, nr = 0, seed = 0, thr = 0, d = 0;
    CHECK(vx_upload(ctx, img, VX_F32, NX, NY, NZ, NULL, &f));
    CHECK(vx_cmp_scalar(ctx, f, VX_LT, 8.0f, &lo));          /* x < 8        */
    CHECK(vx_near(ctx, lo, VX_CONN_FULL, &nr));              /* x < 9        */
    CHECK(vx_cmp_scalar(ctx, f, VX_EQ, 0.0f, &seed));        /* x == 0       */
    CHECK(vx_through(ctx, seed, nr, VX_CONN_FACE, &thr));    /* all of x < 9 */
    double vol = 0, dmax = 0;
    CHECK(vx_volume(ctx, thr, &vol));
    CHECK(vx_edt(ctx, lo, &d));                              /* distance to {x < 8}: max(0, x - 7) */
    CHECK(vx_max(ctx, d, &dmax));
    static float dist[N];
    CHECK(vx_download(ctx, d, dist, sizeof(dist)));
    int bad = 0;
    for (int i = 0; i < N; i++) {
        float want = (i % NX) > 7 ? (float)((i % NX) - 7) : 0.0f;
        if (dist[i] != want) bad++;
    }
    vx_img all[6] = {f, lo, nr, seed, thr, d};
    for (int i = 0; i < 6; i++) CHECK(vx_release(all[i]));
    uint64_t live = 1, handles = 1;
    CHECK(vx_mem_info(ctx, &live, &handles, NULL));
    CHECK(vx_shutdown(ctx));
    if (vol != 9.0 * NY * NZ || dmax != 24.0 || bad || handles != 0 || live != 0) {
        fprintf(stderr, "c_smoke gpu FAILED: volume %g max %g bad %d live %llu handles %llu\n", vol, dmax, bad,
                (unsigned long long)live, (unsigned long long)handles);
        return 1;
    }
    printf("c_smoke gpu ok: volume %g, max distance %g\n", vol, dmax);
    return 0;
}
