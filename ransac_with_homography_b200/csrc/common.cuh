// common.cuh -- context, scratch memory and error plumbing shared by the kernels' host wrappers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200geo.h"

static_assert(sizeof(b200geo_warp_job) == 184, "b200geo_warp_job layout is part of the ABI");
static_assert(sizeof(b200geo_compose) == 80, "b200geo_compose layout is part of the ABI");
static_assert(sizeof(b200geo_src_layout) == 32, "b200geo_src_layout layout is part of the ABI");

namespace b200geo {

static thread_local char g_err[512] = "";

inline int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return code;
}

#define B2G_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) return ::b200geo::fail(B200GEO_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

#define B2G_CHECK_LAUNCH(ctx)                                                                   \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ == cudaSuccess && getenv("B200GEO_SYNC_LAUNCHES")) e__ = cudaDeviceSynchronize();  /* debugging: fault attribution */ \
        if (e__ != cudaSuccess) {                                                               \
            if (getenv("B200GEO_SYNC_LAUNCHES")) fprintf(stderr, "[b200geo] %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return ::b200geo::fail(B200GEO_ERR_CUDA, "kernel launch: %s%s", cudaGetErrorString(e__)); \
        }                                                                                       \
    } while (0)

#define B2G_ARG(cond)                                                                           \
    do {                                                                                        \
        if (!(cond)) return ::b200geo::fail(B200GEO_ERR_ARG, "invalid argument: %s%s", #cond);    \
    } while (0)

enum ScratchSlot {
    S_PTS = 0, S_IDX, S_H, S_HINV, S_FLAGS, S_COUNTS, S_KEYS, S_RESULT, S_MASK, S_JOBS, S_SRC, S_DST, S_MISC, S_CMAX, S_STATS, S_WORKLIST, S_TILES, S_PIPE_SRC, S_PIPE_DST, S_PIPE_JOBS, S_WARPFIX, S_STRIPS, S_FINISH, S_WLCOUNT, S_ORB_PYR, S_ORB_SCORE, S_ORB_BLUR, S_ORB_CAND, S_ORB_PTS,
    S_NUM
};

}  // namespace b200geo

struct b200geo_ctx {
    int device;
    int sm_count;
    int cc;
    int clock_khz;
    size_t mem_bytes;
    int64_t launches;
    int force_exact;       // B200GEO_OPT_FORCE_EXACT: score with the reference-order chain only
    int warp_exact;        // B200GEO_OPT_WARP_EXACT: warp uint8 images with the float64 kernel only
    int warp_variant;      // B200GEO_OPT_WARP_VARIANT: 0 = strips per warp + persistent kernel for the rest, 1 = one CTA per tile, 2 = persistent kernel only
    int warp_pipe_ready;   // dynamic shared memory attribute of warp_pipe_kernel set
    cudaStream_t pipe_stream[3];   // b200geo_warp_batch_host: h2d, compute, d2h
    cudaEvent_t pipe_event[3][2];  // [h2d done, compute done, d2h done][slot]
    int pipe_ready;
    void *tmap_encode;     // cuTensorMapEncodeTiled (driver entry point), or NULL
    int tmap_probed;
    CUtensorMap tmap;      // cached 2-D / 3-D tensor map of the last source layout (box of warp_pipe_kernel)
    CUtensorMap tmap_strip; // same array, box of warp_strip_kernel (3-D layouts only)
    int tmap_strip_ok;
    const void *tmap_src;
    long long tmap_pitch, tmap_rows, tmap_image_rows, tmap_row_bytes;
    void *scratch[b200geo::S_NUM];
    size_t scratch_cap[b200geo::S_NUM];
    const void *cmax_pts;  // S_CMAX holds the largest |coordinate| of these cmax_B sets of cmax_M correspondences
    int cmax_M, cmax_B;
    cudaStream_t order_stream;  // stream of the latest _dev call (the scratch buffers are shared: see stream_order)
    int order_valid;
    cudaEvent_t order_event;
    void *finish_done_zeroed;   // finish_kernel's per-pair arrival counters: where they are, and how many are known to be zero
    int finish_done_n;
    void *pinned;          // small pinned staging block for scalar results
    size_t pinned_cap;
    void *peer_box;        // this rank's inbox of the peer-memory record exchange (cudaMalloc, exported by CUDA IPC), or NULL
    void *peer_map[64];    // every rank's inbox as mapped into this process (own rank: peer_box)
    int peer_world, peer_rank;
    unsigned long long peer_epoch;
    void *jpeg;            // b200geo::JpegState (nvJPEG handle and codec states), created by the first JPEG call
};

namespace b200geo {

// grow-on-demand device scratch; contents are NOT preserved across a growth
inline int scratch(b200geo_ctx *ctx, int slot, size_t bytes, void **out)
{
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_cap[slot] < bytes) {
        if (ctx->scratch[slot]) {
            B2G_CUDA(cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_cap[slot] = 0;
        }
        size_t cap = bytes + bytes / 4;
        cap = (cap + 255) & ~(size_t)255;
        B2G_CUDA(cudaMalloc(&ctx->scratch[slot], cap));
        ctx->scratch_cap[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return B200GEO_OK;
}

inline int pinned(b200geo_ctx *ctx, size_t bytes, void **out)
{
    if (ctx->pinned_cap < bytes) {
        if (ctx->pinned) B2G_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
        B2G_CUDA(cudaMallocHost(&ctx->pinned, bytes));
        ctx->pinned_cap = bytes;
    }
    *out = ctx->pinned;
    return B200GEO_OK;
}

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// A context's scratch buffers, counters and cached tensor maps are shared by all its calls.  Calls issued on ONE stream are
// ordered by the stream; when a _dev call arrives on a different stream than the previous one, it is made to wait for
// everything queued on that previous stream so far (an event recorded there now), so that two streams never run two calls
// of one context concurrently.  No cost while the stream does not change.
inline void stream_order(b200geo_ctx *ctx, cudaStream_t st)
{
    if (ctx->order_valid && ctx->order_stream != st) {
        if (!ctx->order_event) cudaEventCreateWithFlags(&ctx->order_event, cudaEventDisableTiming);
        if (ctx->order_event && cudaEventRecord(ctx->order_event, ctx->order_stream) == cudaSuccess)
            cudaStreamWaitEvent(st, ctx->order_event, 0);
        cudaGetLastError();          // (the previous stream may have been destroyed by its owner: nothing left to wait for)
    }
    ctx->order_stream = st;
    ctx->order_valid = 1;
}

#define B2G_ENTER(ctx)                                                                          \
    B2G_ARG((ctx) != nullptr);                                                                  \
    ::b200geo::DeviceGuard guard__((ctx)->device);                                              \
    if (!guard__.ok) return ::b200geo::fail(B200GEO_ERR_CUDA, "cudaSetDevice failed%s%s")

}  // namespace b200geo
