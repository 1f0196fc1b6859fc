// b200geo.cu -- the C ABI of libb200geo.so (declared in include/b200geo.h).
// Host-side launch logic only; the kernels live in solve.cuh / score.cuh / select_refit.cuh / warp.cuh.
#include "common.cuh"
#include "score.cuh"
#include "select_refit.cuh"
#include "solve.cuh"
#include "warp.cuh"
#include "warp_fast.cuh"
#include "warp_strip.cuh"
#include "compose.cuh"
#include "match.cuh"
#include "remap.cuh"
#include "jpeg_io.cuh"
#include "orb.cuh"

#include <new>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <cmath>
#define CV_PI_F64 3.1415926535897932384626433832795

using namespace b200geo;

#ifndef B200GEO_SOLVE_T
#define B200GEO_SOLVE_T double
#endif
typedef B200GEO_SOLVE_T solve_t;

extern "C" {

int b200geo_abi_version(void) { return B200GEO_ABI_VERSION; }
const char *b200geo_last_error(void) { return g_err; }

int b200geo_create(int device, b200geo_ctx **out)
{
    if (!out) return fail(B200GEO_ERR_ARG, "out is NULL%s%s");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(B200GEO_ERR_NODEVICE, "no CUDA device: %s%s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(B200GEO_ERR_ARG, "device index out of range%s%s");
    cudaDeviceProp p;
    B2G_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10)
        return fail(B200GEO_ERR_NODEVICE, "device %s is not sm_100 class; this library is built for sm_100a only%s", p.name);
    b200geo_ctx *ctx = new (std::nothrow) b200geo_ctx();
    if (!ctx) return fail(B200GEO_ERR_ARG, "out of host memory%s%s");
    memset(ctx, 0, sizeof *ctx);
    ctx->device = device;
    ctx->sm_count = p.multiProcessorCount;
    ctx->cc = p.major * 10 + p.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    ctx->clock_khz = khz;
    ctx->mem_bytes = p.totalGlobalMem;
    *out = ctx;
    return B200GEO_OK;
}

int b200geo_destroy(b200geo_ctx *ctx)
{
    if (!ctx) return B200GEO_OK;
    {
        DeviceGuard g(ctx->device);
        for (int i = 0; i < S_NUM; ++i)
            if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        if (ctx->order_event) cudaEventDestroy(ctx->order_event);
        jpeg_release(ctx);
        if (ctx->peer_box) {
            for (int r = 0; r < ctx->peer_world; ++r)
                if (ctx->peer_map[r] && ctx->peer_map[r] != ctx->peer_box) cudaIpcCloseMemHandle(ctx->peer_map[r]);
            cudaFree(ctx->peer_box);
        }
        if (ctx->pipe_ready) {
            for (int i = 0; i < 3; ++i) cudaStreamDestroy(ctx->pipe_stream[i]);
            for (int i = 0; i < 3; ++i)
                for (int b = 0; b < 2; ++b) cudaEventDestroy(ctx->pipe_event[i][b]);
        }
    }
    delete ctx;
    return B200GEO_OK;
}

int b200geo_device_info(b200geo_ctx *ctx, int *sm_count, int *cc, int *clock_khz, size_t *mem_bytes)
{
    B2G_ARG(ctx != nullptr);
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc) *cc = ctx->cc;
    if (clock_khz) *clock_khz = ctx->clock_khz;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    return B200GEO_OK;
}

int64_t b200geo_launch_count(b200geo_ctx *ctx) { return ctx ? ctx->launches : -1; }

int b200geo_stats(b200geo_ctx *ctx, unsigned long long *exact_path_pairs, int reset)
{
    B2G_ENTER(ctx);
    unsigned long long v = 0;
    if (ctx->scratch_cap[S_STATS]) {
        B2G_CUDA(cudaDeviceSynchronize());
        B2G_CUDA(cudaMemcpy(&v, ctx->scratch[S_STATS], 8, cudaMemcpyDeviceToHost));
        if (reset) B2G_CUDA(cudaMemset(ctx->scratch[S_STATS], 0, 8));
    }
    if (exact_path_pairs) *exact_path_pairs = v;
    return B200GEO_OK;
}

int b200geo_warp_stats(b200geo_ctx *ctx, unsigned long long *exact_pixels, int reset)
{
    B2G_ENTER(ctx);
    unsigned long long v = 0;
    if (ctx->scratch_cap[S_STATS]) {
        B2G_CUDA(cudaDeviceSynchronize());
        B2G_CUDA(cudaMemcpy(&v, (char *)ctx->scratch[S_STATS] + 16, 8, cudaMemcpyDeviceToHost));
        if (reset) B2G_CUDA(cudaMemset((char *)ctx->scratch[S_STATS] + 16, 0, 8));
    }
    if (exact_pixels) *exact_pixels = v;
    return B200GEO_OK;
}

int b200geo_set_option(b200geo_ctx *ctx, int option, int value)
{
    B2G_ARG(ctx != nullptr);
    if (option == B200GEO_OPT_FORCE_EXACT) { ctx->force_exact = value; return B200GEO_OK; }
    if (option == B200GEO_OPT_WARP_EXACT) { ctx->warp_exact = value; return B200GEO_OK; }
    if (option == B200GEO_OPT_WARP_VARIANT) { ctx->warp_variant = value; return B200GEO_OK; }
    return fail(B200GEO_ERR_ARG, "unknown option%s%s");
}

int b200geo_sync(b200geo_ctx *ctx, void *stream)
{
    B2G_ENTER(ctx);
    B2G_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B200GEO_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// internal launch helpers (device pointers, stream ordered)
// ------------------------------------------------------------------------------------------------
namespace {

size_t type_size(int t) { return t == B200GEO_U8 ? 1 : (t == B200GEO_F32 ? 4 : 8); }

// 128-byte statistics / counter block (8-byte words): [0] score pairs taken by the exact chain, [1] deferred-work counter,
// [2] warp pixels evaluated in float64 by the fast warp kernel, [3] tile tickets of warp_pipe_kernel,
// [4] routed launch: number of strip tiles, number of old-kernel tiles (2 x u32), [5] tile tickets of warp_strip_kernel
int ensure_stats(b200geo_ctx *ctx, cudaStream_t st, unsigned long long **out)
{
    if (ctx->scratch_cap[S_STATS] == 0) {
        void *p;
        int rc;
        if ((rc = scratch(ctx, S_STATS, 128, &p))) return rc;
        B2G_CUDA(cudaMemsetAsync(p, 0, 128, st));
    }
    *out = (unsigned long long *)ctx->scratch[S_STATS];
    return B200GEO_OK;
}

int launch_dlt4(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int k_begin, int k_end, int K_stride,
                int B, float *H, float *Hinv, uint8_t *flags, cudaStream_t st)
{
    int nk = k_end - k_begin;
    if (nk <= 0 || B <= 0) return B200GEO_OK;
    dim3 grid((nk + 127) / 128, B);
    dlt4_kernel<solve_t><<<grid, 128, 0, st>>>((const float4 *)pts, M, (const int4 *)idx, k_begin, k_end, K_stride, H, Hinv, flags);
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

// exact kernel: any method, any threshold
int launch_score_exact(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv, const uint8_t *flags,
                       int k_begin, int k_end, int K_stride, int B, float th, int method, int32_t *counts, cudaStream_t st)
{
    int nk = k_end - k_begin;
    int gx = (nk + SCORE_THREADS - 1) / SCORE_THREADS;
    // split the correspondence range when there are too few hypothesis blocks to fill the GPU
    int splits = 1;
    long long blocks = (long long)gx * B;
    long long want = 2LL * ctx->sm_count;
    if (blocks < want && M > SCORE_TILE) {
        splits = (int)((want + blocks - 1) / blocks);
        int max_splits = (M + SCORE_TILE - 1) / SCORE_TILE;
        if (splits > max_splits) splits = max_splits;
    }
    int per = (M + splits - 1) / splits;
    per = ((per + SCORE_TILE - 1) / SCORE_TILE) * SCORE_TILE;
    splits = M > 0 ? (M + per - 1) / per : 1;
    if (splits < 1) splits = 1;
    if (splits > 1) {
        // (one 2-D memset for all pairs: a memset per pair was 0.3 ms of a 128-pair step)
        B2G_CUDA(cudaMemset2DAsync(counts + k_begin, sizeof(int32_t) * (size_t)K_stride, 0, sizeof(int32_t) * (size_t)nk, (size_t)B, st));
    }
    dim3 grid(gx, splits, B);
    const float4 *p4 = (const float4 *)pts;
    switch (method) {
    case B200GEO_FWD:
        score_exact_kernel<B200GEO_FWD><<<grid, SCORE_THREADS, 0, st>>>(p4, M, H, Hinv, flags, k_begin, k_end, K_stride, th, per, counts);
        break;
    case B200GEO_BACKWARD:
        score_exact_kernel<B200GEO_BACKWARD><<<grid, SCORE_THREADS, 0, st>>>(p4, M, H, Hinv, flags, k_begin, k_end, K_stride, th, per, counts);
        break;
    default:
        score_exact_kernel<B200GEO_REPROJ><<<grid, SCORE_THREADS, 0, st>>>(p4, M, H, Hinv, flags, k_begin, k_end, K_stride, th, per, counts);
        break;
    }
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

int launch_cmax(b200geo_ctx *ctx, const float *pts, int M, int B, float **cmax_out, cudaStream_t st)
{
    float *cmax;
    int rc;
    if ((rc = scratch(ctx, S_CMAX, sizeof(float) * (size_t)B, (void **)&cmax))) return rc;
    B2G_CUDA(cudaMemsetAsync(cmax, 0, sizeof(float) * (size_t)B, st));
    cmax_kernel<<<dim3((M + 2047) / 2048, B), 256, 0, st>>>((const float4 *)pts, M, cmax);
    B2G_CHECK_LAUNCH(ctx);
    *cmax_out = cmax;
    ctx->cmax_pts = (const void *)pts; ctx->cmax_M = M; ctx->cmax_B = B;
    return B200GEO_OK;
}

// Point-range splits of a scoring launch: `blocks` CTAs per split, `slots` CTAs resident on the GPU.  A CTA's time grows with the
// correspondences it streams, the launch's with the number of waves: minimise waves x points-per-split (+ a CTA's prologue) over
// the split count.  (Round 1 took "at least 16 waves" and whole 1024-point tiles: 8.3 waves -> 9 at K/8 = 125k sample sets, one
// tenth of the kernel's time lost to the last wave; the same search also serves the unsharded launch.)
int choose_splits(long long blocks, int M, int slots, int granule, int min_pts, int *per_out)
{
    int best_s = 1, best_per = (M + granule - 1) / granule * granule;
    double best_cost = 1e300;
    int max_s = M / min_pts;
    if (max_s < 1) max_s = 1;
    if (max_s > 512) max_s = 512;
    for (int s = 1; s <= max_s; ++s) {
        int per = ((M + s - 1) / s + granule - 1) / granule * granule;
        if (per < 1) per = granule;
        const int s_eff = (M + per - 1) / per;
        const long long ctas = blocks * s_eff;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (per + 40.0);
        if (cost < best_cost * 0.9999) { best_cost = cost; best_s = s_eff; best_per = per; }
    }
    *per_out = best_per;
    return best_s;
}

// The deferred exact-work list of a scoring launch: one segment of seg_cap entries per CTA (counts[cta] entries used, written
// by the CTA itself -- no initialisation needed).  `want` = entries the whole launch may need (1M .. 64M in total).
int worklist(b200geo_ctx *ctx, unsigned long long want, unsigned long long n_ctas, FastDeferred **wl, unsigned **counts, unsigned *n_seg,
             unsigned *seg_cap)
{
    if (want < (1ull << 20)) want = 1ull << 20;
    if (want > (1ull << 26)) want = 1ull << 26;
    if (n_ctas < 1) n_ctas = 1;
    unsigned long long per = want / n_ctas;
    if (per < 64) per = 64;                          // (a full segment falls back to settling the chunk inside the kernel)
    if (per * n_ctas > (1ull << 27)) per = (1ull << 27) / n_ctas;
    if (per < 1) per = 1;
    int rc;
    if ((rc = scratch(ctx, S_WORKLIST, per * n_ctas * sizeof(FastDeferred), (void **)wl))) return rc;
    if ((rc = scratch(ctx, S_WLCOUNT, n_ctas * sizeof(unsigned), (void **)counts))) return rc;
    *n_seg = (unsigned)n_ctas;
    *seg_cap = (unsigned)per;
    return B200GEO_OK;
}

// K2.  fwd / backward with a positive finite threshold take the division-free FFMA2 kernel (exact
// chain only inside the ambiguity shell); everything else takes the exact kernel.
int launch_score(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv, const uint8_t *flags,
                 int k_begin, int k_end, int K_stride, int B, float th, int method, int32_t *counts, cudaStream_t st)
{
    int nk = k_end - k_begin;
    if (nk <= 0 || B <= 0) return B200GEO_OK;
    const bool fast_ok = th > 0.f && th < 1.0e15f && M > 0 && !ctx->force_exact && (method != B200GEO_REPROJ || Hinv != nullptr);
    if (!fast_ok) return launch_score_exact(ctx, pts, M, H, Hinv, flags, k_begin, k_end, K_stride, B, th, method, counts, st);

    float *cmax;
    unsigned long long *stats;
    int rc;
    if ((rc = ensure_stats(ctx, st, &stats))) return rc;
    if ((rc = launch_cmax(ctx, pts, M, B, &cmax, st))) return rc;

    if (method == B200GEO_REPROJ) {
        // two-sided bracket of the sum of the two errors (score_fast2_kernel), same deferred exact list
        const int per_cta2 = FAST2_THREADS * FAST2_T;
        int gx2 = (nk + per_cta2 - 1) / per_cta2;
        int per2 = M;
        int splits2 = choose_splits((long long)gx2 * B, M, FAST2_CTAS * ctx->sm_count, FAST_CHUNK, 256, &per2);
        if (splits2 > 1)
            B2G_CUDA(cudaMemset2DAsync(counts + k_begin, sizeof(int32_t) * (size_t)K_stride, 0, sizeof(int32_t) * (size_t)nk, (size_t)B, st));
        unsigned long long cap2 = (unsigned long long)nk * B * (unsigned long long)((M + FAST_CHUNK - 1) / FAST_CHUNK) / 32;
        FastDeferred *wl2;
        unsigned *wl2_count, n_seg2, seg_cap2;
        if ((rc = worklist(ctx, cap2, (unsigned long long)gx2 * splits2 * B, &wl2, &wl2_count, &n_seg2, &seg_cap2))) return rc;
        score_fast2_kernel<<<dim3(gx2, splits2, B), FAST2_THREADS, 0, st>>>((const float4 *)pts, M, H, Hinv, flags, k_begin, k_end, K_stride, th,
                                                                            cmax, per2, counts, wl2, wl2_count, seg_cap2);
        B2G_CHECK_LAUNCH(ctx);
        score_fixup2_kernel<<<8 * ctx->sm_count, 256, 0, st>>>((const float4 *)pts, M, H, Hinv, K_stride, th, cmax, wl2, wl2_count, n_seg2, seg_cap2,
                                                               counts, stats);
        B2G_CHECK_LAUNCH(ctx);
        return B200GEO_OK;
    }

    const int per_cta = FAST_THREADS * FAST_T;
    int gx = (nk + per_cta - 1) / per_cta;
    int per = M;
    int splits = choose_splits((long long)gx * B, M, 2 * ctx->sm_count, FAST_CHUNK, 256, &per);
    if (splits > 1) {
        // (one 2-D memset for all pairs: a memset per pair was 0.3 ms of a 128-pair step)
        B2G_CUDA(cudaMemset2DAsync(counts + k_begin, sizeof(int32_t) * (size_t)K_stride, 0, sizeof(int32_t) * (size_t)nk, (size_t)B, st));
    }
    // deferred exact work: capacity for a 1/64 trigger rate per (hypothesis, chunk), 1M .. 64M entries
    unsigned long long want_cap = (unsigned long long)nk * B * (unsigned long long)((M + FAST_CHUNK - 1) / FAST_CHUNK) / 64;
    FastDeferred *wl;
    unsigned *wl_count, n_seg, wl_cap;
    if ((rc = worklist(ctx, want_cap, (unsigned long long)gx * splits * B, &wl, &wl_count, &n_seg, &wl_cap))) return rc;
    dim3 grid(gx, splits, B);
    const float4 *p4 = (const float4 *)pts;
    int fix_blocks = 8 * ctx->sm_count;
    if (method == B200GEO_FWD) {
        score_fast_kernel<B200GEO_FWD><<<grid, FAST_THREADS, 0, st>>>(p4, M, H, flags, k_begin, k_end, K_stride, th, cmax, per, counts, stats, wl, wl_count, wl_cap);
        B2G_CHECK_LAUNCH(ctx);
        score_fixup_kernel<B200GEO_FWD><<<fix_blocks, 256, 0, st>>>(p4, M, H, K_stride, th, cmax, wl, wl_count, n_seg, wl_cap, counts, stats);
    } else {
        score_fast_kernel<B200GEO_BACKWARD><<<grid, FAST_THREADS, 0, st>>>(p4, M, Hinv, flags, k_begin, k_end, K_stride, th, cmax, per, counts, stats, wl, wl_count, wl_cap);
        B2G_CHECK_LAUNCH(ctx);
        score_fixup_kernel<B200GEO_BACKWARD><<<fix_blocks, 256, 0, st>>>(p4, M, Hinv, K_stride, th, cmax, wl, wl_count, n_seg, wl_cap, counts, stats);
    }
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

int launch_select(b200geo_ctx *ctx, const int32_t *counts, int k_begin, int k_end, int K_stride, int B, double thr,
                  unsigned long long *keys, cudaStream_t st)
{
    int nk = k_end - k_begin;
    if (nk <= 0 || B <= 0) return B200GEO_OK;
    int gx = (nk + 256 * 8 - 1) / (256 * 8);
    if (gx > 4 * ctx->sm_count) gx = 4 * ctx->sm_count;
    if (gx < 1) gx = 1;
    select_kernel<<<dim3(gx, B), 256, 0, st>>>(counts, k_begin, k_end, K_stride, thr, keys);
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

// solve + score + select over idx[k_begin, k_end) for B independent pairs
int search(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int k_begin, int k_end, int K_stride, int B,
           float th, double thr, int method, unsigned long long *keys, int32_t *counts, cudaStream_t st)
{
    size_t nH = (size_t)B * K_stride;
    float *H = nullptr, *Hinv = nullptr;
    uint8_t *flags = nullptr;
    int rc;
    if ((rc = scratch(ctx, S_H, nH * 9 * sizeof(float), (void **)&H))) return rc;
    if (method != B200GEO_FWD && (rc = scratch(ctx, S_HINV, nH * 9 * sizeof(float), (void **)&Hinv))) return rc;
    if ((rc = scratch(ctx, S_FLAGS, nH, (void **)&flags))) return rc;
    if (!counts && (rc = scratch(ctx, S_COUNTS, nH * sizeof(int32_t), (void **)&counts))) return rc;
    if ((rc = launch_dlt4(ctx, pts, M, idx, k_begin, k_end, K_stride, B, H, Hinv, flags, st))) return rc;
    if ((rc = launch_score(ctx, pts, M, H, Hinv, flags, k_begin, k_end, K_stride, B, th, method, counts, st))) return rc;
    return launch_select(ctx, counts, k_begin, k_end, K_stride, B, thr, keys, st);
}

// K3 over B pairs: G CTAs per pair (mask slice + Gram partials), the last one solves.  S_CMAX is reused when it holds the largest
// coordinate magnitudes of exactly these B sets (left there by the search that produced `keys`).
int launch_finish(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K_stride, int B,
                  const unsigned long long *keys, float th, int method, double *result, uint8_t *mask, cudaStream_t st,
                  const WinRec *recs = nullptr, int n_recs = 0)
{
    float *cmax = (float *)ctx->scratch[S_CMAX];
    int rc;
    const bool cmax_valid = cmax && ctx->cmax_pts == (const void *)pts && ctx->cmax_M == M && ctx->cmax_B == B;
    if (!cmax_valid) {
        if ((rc = launch_cmax(ctx, pts, M, B, &cmax, st))) return rc;
    }
    int G = (M + 1023) / 1024;                     // ~4 correspondences per thread
    if (G > 64) G = 64;
    if ((long long)G * B > 8LL * ctx->sm_count && B > 1) { G = (int)(8LL * ctx->sm_count / B); if (G < 1) G = 1; }
    int per = (M + G - 1) / G;
    G = (M + per - 1) / per;
    double *partial;
    unsigned *done;
    if ((rc = scratch(ctx, S_FINISH, sizeof(double) * FINISH_NSUM * (size_t)G * B + sizeof(unsigned) * (size_t)B + 64, (void **)&partial))) return rc;
    done = (unsigned *)(partial + (size_t)FINISH_NSUM * G * B);
    if (ctx->finish_done_zeroed != (void *)done || ctx->finish_done_n < B) {      // the kernel leaves the counters at zero
        B2G_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned) * (size_t)B, st));
        ctx->finish_done_zeroed = (void *)done;
        ctx->finish_done_n = B;
    }
    finish_kernel<solve_t><<<dim3(G, B), FINISH_THREADS, 0, st>>>((const float4 *)pts, M, (const int4 *)idx, K_stride, keys, recs, n_recs, th, method,
                                                                 cmax, per, partial, done, result, mask);
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

template <typename SrcT, int C>
int launch_warp_t(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs_dev, int n_jobs, int max_ow,
                  int max_oh, int dst_type, int sampler, cudaStream_t st)
{
    dim3 block(64, 4);
    dim3 grid((max_ow + 63) / 64, (max_oh + 3) / 4, n_jobs);
    const uint8_t *s = (const uint8_t *)src;
    uint8_t *d = (uint8_t *)dst;
#define B2G_WARP_CASE(DT, BL) warp_generic_kernel<SrcT, DT, C, BL><<<grid, block, 0, st>>>(s, d, jobs_dev)
    if (sampler == B200GEO_BILINEAR) {
        if (dst_type == B200GEO_U8) B2G_WARP_CASE(uint8_t, 1);
        else if (dst_type == B200GEO_F32) B2G_WARP_CASE(float, 1);
        else B2G_WARP_CASE(double, 1);
    } else {
        if (dst_type == B200GEO_U8) B2G_WARP_CASE(uint8_t, 0);
        else if (dst_type == B200GEO_F32) B2G_WARP_CASE(float, 0);
        else B2G_WARP_CASE(double, 0);
    }
#undef B2G_WARP_CASE
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

// 2-D tensor map over the source byte array (uint8, rows x pitch), box WF_TMA_BOX_W x WF_TMA_BOX_H.
// Returns 2 (2-D) or 3 (3-D: per-image zero fill) and fills ctx->tmap when the layout is usable, 0 otherwise (never an error: the
// kernels fall back to per-row bulk copies).
int prepare_tmap(b200geo_ctx *ctx, const void *src, const b200geo_src_layout *layout)
{
    if (!layout || layout->pitch <= 0 || layout->rows <= 0) return 0;
    if ((layout->pitch & 15) || ((size_t)src & 15) || layout->pitch < WF_TMA_BOX_W || layout->rows < 1) return 0;
    if (!ctx->tmap_probed) {
        ctx->tmap_probed = 1;
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            ctx->tmap_encode = fn;
        cudaGetLastError();
    }
    if (!ctx->tmap_encode) return 0;
    const bool d3 = layout->image_rows >= WF_TMA_BOX_H && layout->rows % layout->image_rows == 0 && layout->row_bytes >= WF_TMA_BOX_W &&
                    layout->row_bytes <= layout->pitch && layout->image_rows < (1 << 24) && layout->rows / layout->image_rows < (1 << 24);
    const int mode = d3 ? 3 : 2;
    if (ctx->tmap_src == src && ctx->tmap_pitch == layout->pitch && ctx->tmap_rows == layout->rows &&
        ctx->tmap_image_rows == (d3 ? layout->image_rows : 0) && ctx->tmap_row_bytes == (d3 ? layout->row_bytes : 0)) return mode;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t gdim[3] = {(cuuint64_t)(d3 ? layout->row_bytes : layout->pitch), (cuuint64_t)(d3 ? layout->image_rows : layout->rows),
                                (cuuint64_t)(d3 ? layout->rows / layout->image_rows : 1)};
    const cuuint64_t gstride[2] = {(cuuint64_t)layout->pitch, (cuuint64_t)layout->pitch * (cuuint64_t)(d3 ? layout->image_rows : 1)};
    const cuuint32_t box[3] = {(cuuint32_t)WF_TMA_BOX_W, (cuuint32_t)WF_TMA_BOX_H, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = ((encode_fn)ctx->tmap_encode)(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, d3 ? 3 : 2, const_cast<void *>(src), gdim, gstride, box,
                                               estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->tmap_src = nullptr; return 0; }
    ctx->tmap_strip_ok = 0;
    if (d3) {       // the per-warp strips of warp_strip_kernel fetch a 160-byte x 26-row box of the same array
        const cuuint32_t sbox[3] = {(cuuint32_t)WS_RAW_W, (cuuint32_t)WS_ROWS, 1};
        r = ((encode_fn)ctx->tmap_encode)(&ctx->tmap_strip, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(src), gdim, gstride, sbox,
                                          estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ctx->tmap_strip_ok = r == CUDA_SUCCESS;
    }
    ctx->tmap_src = src; ctx->tmap_pitch = layout->pitch; ctx->tmap_rows = layout->rows;
    ctx->tmap_image_rows = d3 ? layout->image_rows : 0; ctx->tmap_row_bytes = d3 ? layout->row_bytes : 0;
    return mode;
}

int launch_warp(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs_dev, int n_jobs, int max_ow,
                int max_oh, int channels, int src_type, int dst_type, int sampler, cudaStream_t st,
                const b200geo_src_layout *layout = nullptr)
{
    if (n_jobs <= 0 || max_ow <= 0 || max_oh <= 0) return B200GEO_OK;
    if (n_jobs > 65535) return fail(B200GEO_ERR_ARG, "at most 65535 warp jobs per call%s%s");
    if (src_type == B200GEO_U8 && dst_type == B200GEO_U8 && channels == 3 && !ctx->warp_exact) {
        // fast path: tile-staged, float32 tile-relative coordinates, integer blend (warp_fast.cuh)
        unsigned long long *stats;
        int rc;
        if ((rc = ensure_stats(ctx, st, &stats))) return rc;
        const int tiles_x = (max_ow + WF_TW - 1) / WF_TW, tiles_y = (max_oh + WF_TH - 1) / WF_TH;
        if (tiles_y > 65535) return fail(B200GEO_ERR_ARG, "output too tall for one launch%s%s");
        // tile records: 64-byte descriptors for the one-CTA-per-tile kernel, 192-byte pipeline records for the persistent one
        const size_t tile_bytes = ctx->warp_variant == 1 ? sizeof(WfDesc) : sizeof(WpTile);
        void *tile_mem;
        if ((rc = scratch(ctx, S_TILES, tile_bytes * (size_t)tiles_x * tiles_y * n_jobs, &tile_mem))) return rc;
        WfDesc *descs = (WfDesc *)tile_mem;
        WpTile *tiles = (WpTile *)tile_mem;
        dim3 sgrid((tiles_x + 31) / 32, (tiles_y + 7) / 8, n_jobs), sblock(32, 8);
        const uint8_t *s8 = (const uint8_t *)src;
        uint8_t *d8 = (uint8_t *)dst;
        // deferred float64 pixels: list sized for ~0.4 % of the output, settled by warp_fixup_kernel.  The
        // scratch block starts with the per-segment counters (4 KB), the entries follow.
        unsigned long long fix_want = (unsigned long long)max_ow * max_oh * n_jobs / 256 + 65536;
        if (fix_want > (1ull << 27)) fix_want = 1ull << 27;
        constexpr size_t FIX_HDR = 4096;
        char *fix_block;
        if ((rc = scratch(ctx, S_WARPFIX, FIX_HDR + fix_want * sizeof(WfFixEntry), (void **)&fix_block))) return rc;
        unsigned *fix_count = (unsigned *)fix_block;
        WfFixEntry *fix_list = (WfFixEntry *)(fix_block + FIX_HDR);
        const unsigned fix_cap = (unsigned)((ctx->scratch_cap[S_WARPFIX] - FIX_HDR) / sizeof(WfFixEntry));
        unsigned *tile_counter = (unsigned *)(stats + 3);       // dynamic tile tickets of the persistent kernel
        B2G_CUDA(cudaMemsetAsync(tile_counter, 0, sizeof(unsigned), st));
        B2G_CUDA(cudaMemsetAsync(fix_count, 0, sizeof(unsigned), st));
        unsigned n_seg = 1, seg_cap = fix_cap;
        const int use_tmap = ctx->warp_variant != 1 ? prepare_tmap(ctx, src, layout) : 0;
        if (getenv("B200GEO_DEBUG"))
            fprintf(stderr, "[b200geo] warp: jobs %d tiles %dx%d sampler %d layout %p use_tmap %d (encode %p)\n", n_jobs, tiles_x, tiles_y,
                    sampler, (const void *)layout, use_tmap, ctx->tmap_encode);
        // use_tmap: 0 no tensor map, 2 the source as a 2-D array, 3 as (row_bytes x image_rows x images)
        const WfTma tma = {use_tmap ? (long long)layout->pitch : 0LL, use_tmap == 3 ? (int)layout->image_rows : 0, use_tmap == 3 ? (int)layout->row_bytes : 0};
        const long long n_tiles = (long long)tiles_x * tiles_y * n_jobs;
        if (n_tiles > 0x3fffffffLL) return fail(B200GEO_ERR_ARG, "too many warp tiles for one launch%s%s");
        if (ctx->warp_variant != 1 && !ctx->warp_pipe_ready) {
            B2G_CUDA(cudaFuncSetAttribute(warp_pipe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WpSmem)));
            B2G_CUDA(cudaFuncSetAttribute(warp_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WpSmem)));
            B2G_CUDA(cudaFuncSetAttribute(warp_strip_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
            B2G_CUDA(cudaFuncSetAttribute(warp_strip_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
            ctx->warp_pipe_ready = 1;
        }
        if (ctx->warp_variant == 1) {
            // one CTA per tile (kept for comparison)
            if (sampler == B200GEO_BILINEAR) warp_setup_kernel<1><<<sgrid, sblock, 0, st>>>(jobs_dev, s8, tma, tiles_x, tiles_y, descs);
            else warp_setup_kernel<0><<<sgrid, sblock, 0, st>>>(jobs_dev, s8, tma, tiles_x, tiles_y, descs);
            B2G_CHECK_LAUNCH(ctx);
            dim3 grid(tiles_x, tiles_y, n_jobs);
            if (sampler == B200GEO_BILINEAR) warp_fast_kernel<1><<<grid, WF_THREADS, 0, st>>>(s8, d8, jobs_dev, descs, stats + 2, fix_list, fix_count, fix_cap);
            else warp_fast_kernel<0><<<grid, WF_THREADS, 0, st>>>(s8, d8, jobs_dev, descs, stats + 2, fix_list, fix_count, fix_cap);
        } else if (use_tmap == 3 && ctx->tmap_strip_ok && ctx->warp_variant == 0) {
            // routed: every 64 x 40 tile goes either to warp_strip_kernel (four 32 x 20 strips, one warp each) or, if one of its
            // strips does not qualify, to warp_pipe_kernel; both lists are compact, their lengths stay on the device
            WsStrip *strips;
            cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
            const bool timing = getenv("B200GEO_DEBUG_TIMING") != nullptr;      // debugging: per-kernel times of this call (synchronises)
            if (timing) for (int i = 0; i < 5; ++i) cudaEventCreate(&tev[i]);
            if (timing) cudaEventRecord(tev[0], st);
            if ((rc = scratch(ctx, S_STRIPS, WS_PER_TILE * sizeof(WsStrip) * (size_t)n_tiles, (void **)&strips))) return rc;
            unsigned *route = (unsigned *)(stats + 4), *strip_ticket = (unsigned *)(stats + 5);
            B2G_CUDA(cudaMemsetAsync(route, 0, 16, st));        // both counts and the strip ticket
            if (sampler == B200GEO_BILINEAR) warp_route_setup_kernel<1><<<sgrid, sblock, 0, st>>>(jobs_dev, s8, d8, tma, use_tmap, tiles_x, tiles_y, strips, tiles, route);
            else warp_route_setup_kernel<0><<<sgrid, sblock, 0, st>>>(jobs_dev, s8, d8, tma, use_tmap, tiles_x, tiles_y, strips, tiles, route);
            B2G_CHECK_LAUNCH(ctx);
            if (timing) cudaEventRecord(tev[1], st);
            long long sctas = (n_tiles + WS_WARPS - 1) / WS_WARPS;
            if (sctas > ctx->sm_count) sctas = ctx->sm_count;
            long long pctas = (long long)ctx->sm_count * WP_OCC;
            if (pctas > n_tiles) pctas = n_tiles;
            n_seg = (unsigned)(sctas + pctas); seg_cap = fix_cap / n_seg;
            if (n_seg * sizeof(unsigned) > FIX_HDR || seg_cap == 0) return fail(B200GEO_ERR_ARG, "fix-up list segments%s%s");
            if (sampler == B200GEO_BILINEAR)
                warp_strip_kernel<1><<<(unsigned)sctas, WS_THREADS, WS_SMEM_BYTES, st>>>(s8, jobs_dev, strips, route, ctx->tmap_strip, fix_list, fix_count, seg_cap, strip_ticket, stats + 2);
            else
                warp_strip_kernel<0><<<(unsigned)sctas, WS_THREADS, WS_SMEM_BYTES, st>>>(s8, jobs_dev, strips, route, ctx->tmap_strip, fix_list, fix_count, seg_cap, strip_ticket, stats + 2);
            B2G_CHECK_LAUNCH(ctx);
            if (timing) cudaEventRecord(tev[2], st);
            WfFixEntry *fl = fix_list + (size_t)sctas * seg_cap;
            unsigned *fc = fix_count + sctas;
            if (sampler == B200GEO_BILINEAR)
                warp_pipe_kernel<1><<<(unsigned)pctas, WP_THREADS, sizeof(WpSmem), st>>>(s8, d8, jobs_dev, tiles, 0, route + 1, tiles_x, tiles_y, stats + 2, ctx->tmap, use_tmap, tma.image_rows, fl, fc, seg_cap, tile_counter);
            else
                warp_pipe_kernel<0><<<(unsigned)pctas, WP_THREADS, sizeof(WpSmem), st>>>(s8, d8, jobs_dev, tiles, 0, route + 1, tiles_x, tiles_y, stats + 2, ctx->tmap, use_tmap, tma.image_rows, fl, fc, seg_cap, tile_counter);
            if (timing) {
                cudaEventRecord(tev[3], st);
                cudaEventSynchronize(tev[3]);
                float a = 0, b = 0, c = 0;
                cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[1], tev[2]); cudaEventElapsedTime(&c, tev[2], tev[3]);
                unsigned rc2[2] = {0, 0};
                cudaMemcpy(rc2, route, 8, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[b200geo] warp timing: set-up %.3f ms, strip kernel %.3f ms (%u tiles), tile kernel %.3f ms (%u tiles)\n", a, b, rc2[0], c, rc2[1]);
                for (int i = 0; i < 5; ++i) cudaEventDestroy(tev[i]);
            }
        } else {
            // persistent CTAs, TMA bulk copies one tile ahead
            if (sampler == B200GEO_BILINEAR) warp_tile_setup_kernel<1><<<sgrid, sblock, 0, st>>>(jobs_dev, s8, d8, tma, use_tmap, tiles_x, tiles_y, tiles);
            else warp_tile_setup_kernel<0><<<sgrid, sblock, 0, st>>>(jobs_dev, s8, d8, tma, use_tmap, tiles_x, tiles_y, tiles);
            B2G_CHECK_LAUNCH(ctx);
            long long ctas = (long long)ctx->sm_count * WP_OCC;
            if (ctas > n_tiles) ctas = n_tiles;
            if (ctas > (long long)(FIX_HDR / sizeof(unsigned))) ctas = FIX_HDR / sizeof(unsigned);
            n_seg = (unsigned)ctas; seg_cap = fix_cap / n_seg;
            if (sampler == B200GEO_BILINEAR)
                warp_pipe_kernel<1><<<(unsigned)ctas, WP_THREADS, sizeof(WpSmem), st>>>(s8, d8, jobs_dev, tiles, (int)n_tiles, nullptr, tiles_x, tiles_y, stats + 2, ctx->tmap, use_tmap, tma.image_rows, fix_list, fix_count, seg_cap, tile_counter);
            else
                warp_pipe_kernel<0><<<(unsigned)ctas, WP_THREADS, sizeof(WpSmem), st>>>(s8, d8, jobs_dev, tiles, (int)n_tiles, nullptr, tiles_x, tiles_y, stats + 2, ctx->tmap, use_tmap, tma.image_rows, fix_list, fix_count, seg_cap, tile_counter);
        }
        B2G_CHECK_LAUNCH(ctx);
        const dim3 fgrid(n_seg == 1 ? 2 * ctx->sm_count : 2, n_seg);
        if (sampler == B200GEO_BILINEAR) warp_fixup_kernel<1><<<fgrid, 128, 0, st>>>(s8, d8, jobs_dev, fix_list, fix_count, n_seg, seg_cap, stats + 2);
        else warp_fixup_kernel<0><<<fgrid, 128, 0, st>>>(s8, d8, jobs_dev, fix_list, fix_count, n_seg, seg_cap, stats + 2);
        B2G_CHECK_LAUNCH(ctx);
        return B200GEO_OK;
    }
    if (src_type == B200GEO_U8 && channels == 3) return launch_warp_t<uint8_t, 3>(ctx, src, dst, jobs_dev, n_jobs, max_ow, max_oh, dst_type, sampler, st);
    if (src_type == B200GEO_U8 && channels == 4) return launch_warp_t<uint8_t, 4>(ctx, src, dst, jobs_dev, n_jobs, max_ow, max_oh, dst_type, sampler, st);
    if (src_type == B200GEO_F32 && channels == 3) return launch_warp_t<float, 3>(ctx, src, dst, jobs_dev, n_jobs, max_ow, max_oh, dst_type, sampler, st);
    if (src_type == B200GEO_F32 && channels == 4) return launch_warp_t<float, 4>(ctx, src, dst, jobs_dev, n_jobs, max_ow, max_oh, dst_type, sampler, st);
    return fail(B200GEO_ERR_ARG, "unsupported source type / channel count%s%s");
}


template <typename SrcT, int C>
int launch_sample_t(b200geo_ctx *ctx, const void *src, int ih, int iw, int bh, int bw, const double *zx, const double *zy,
                    long long N, int sampler, int dst_type, void *dst, cudaStream_t st)
{
    long long want = (N + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 16LL * ctx->sm_count ? 16LL * ctx->sm_count : want));
    const SrcT *s = (const SrcT *)src;
    size_t pitch = (size_t)iw * C;
#define B2G_SAMPLE_CASE(DT, BL) sample_kernel<SrcT, DT, C, BL><<<blocks, 256, 0, st>>>(s, pitch, iw, ih, bw, bh, zx, zy, N, (DT *)dst)
    if (sampler == B200GEO_BILINEAR) {
        if (dst_type == B200GEO_U8) B2G_SAMPLE_CASE(uint8_t, 1);
        else if (dst_type == B200GEO_F32) B2G_SAMPLE_CASE(float, 1);
        else B2G_SAMPLE_CASE(double, 1);
    } else {
        if (dst_type == B200GEO_U8) B2G_SAMPLE_CASE(uint8_t, 0);
        else if (dst_type == B200GEO_F32) B2G_SAMPLE_CASE(float, 0);
        else B2G_SAMPLE_CASE(double, 0);
    }
#undef B2G_SAMPLE_CASE
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// exported entry points
// ------------------------------------------------------------------------------------------------
extern "C" {

int b200geo_dlt4_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K, float *H, float *Hinv,
                     uint8_t *flags, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && idx && H && M > 0 && K >= 0);
    return launch_dlt4(ctx, pts, M, idx, 0, K, K, 1, H, Hinv, flags, (cudaStream_t)stream);
}

int b200geo_dlt4_host(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K, float *H, float *Hinv,
                      uint8_t *flags)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(pts && idx && H && M > 0 && K >= 0);
    if (K == 0) return B200GEO_OK;
    float *dp, *dH, *dHi = nullptr;
    int32_t *di;
    uint8_t *df;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 16, (void **)&dp))) return rc;
    if ((rc = scratch(ctx, S_IDX, (size_t)K * 16, (void **)&di))) return rc;
    if ((rc = scratch(ctx, S_H, (size_t)K * 36, (void **)&dH))) return rc;
    if (Hinv && (rc = scratch(ctx, S_HINV, (size_t)K * 36, (void **)&dHi))) return rc;
    if ((rc = scratch(ctx, S_FLAGS, (size_t)K, (void **)&df))) return rc;
    B2G_CUDA(cudaMemcpyAsync(dp, pts, (size_t)M * 16, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(di, idx, (size_t)K * 16, cudaMemcpyHostToDevice, 0));
    if ((rc = launch_dlt4(ctx, dp, M, di, 0, K, K, 1, dH, dHi, df, 0))) return rc;
    B2G_CUDA(cudaMemcpyAsync(H, dH, (size_t)K * 36, cudaMemcpyDeviceToHost, 0));
    if (Hinv) B2G_CUDA(cudaMemcpyAsync(Hinv, dHi, (size_t)K * 36, cudaMemcpyDeviceToHost, 0));
    if (flags) B2G_CUDA(cudaMemcpyAsync(flags, df, (size_t)K, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_homography4_host(b200geo_ctx *ctx, const double *u, const double *v, double *H)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(u && v && H);
    double *d;
    int rc;
    if ((rc = scratch(ctx, S_MISC, 16 * 8 + 9 * 8 + 8, (void **)&d))) return rc;
    B2G_CUDA(cudaMemcpyAsync(d, u, 64, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(d + 8, v, 64, cudaMemcpyHostToDevice, 0));
    homography4_kernel<<<1, 32, 0, 0>>>(d, d + 16, (int *)(d + 25));
    B2G_CHECK_LAUNCH(ctx);
    double host[10];
    B2G_CUDA(cudaMemcpy(host, d + 16, sizeof host, cudaMemcpyDeviceToHost));
    int status;
    memcpy(&status, &host[9], sizeof status);
    memcpy(H, host, 72);
    if (status) return fail(B200GEO_ERR_SINGULAR, "singular 4-point system (the reference raises LinAlgError)%s%s");
    return B200GEO_OK;
}

int b200geo_score_dev(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv, const uint8_t *flags,
                      int K, float th, int method, int32_t *counts, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && H && counts && M >= 0 && K >= 0);
    B2G_ARG(method == B200GEO_FWD || method == B200GEO_BACKWARD || method == B200GEO_REPROJ);
    B2G_ARG(method == B200GEO_FWD || Hinv != nullptr);
    return launch_score(ctx, pts, M, H, Hinv, flags, 0, K, K, 1, th, method, counts, (cudaStream_t)stream);
}

int b200geo_score_host(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv, int K, float th,
                       int method, int32_t *counts)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(pts && H && counts && M >= 0 && K >= 0);
    B2G_ARG(method == B200GEO_FWD || method == B200GEO_BACKWARD || method == B200GEO_REPROJ);
    B2G_ARG(method == B200GEO_FWD || Hinv != nullptr);
    if (K == 0) return B200GEO_OK;
    float *dp, *dH, *dHi = nullptr;
    int32_t *dc;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 16, (void **)&dp))) return rc;
    if ((rc = scratch(ctx, S_H, (size_t)K * 36, (void **)&dH))) return rc;
    if (Hinv && (rc = scratch(ctx, S_HINV, (size_t)K * 36, (void **)&dHi))) return rc;
    if ((rc = scratch(ctx, S_COUNTS, (size_t)K * 4, (void **)&dc))) return rc;
    if (M) B2G_CUDA(cudaMemcpyAsync(dp, pts, (size_t)M * 16, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dH, H, (size_t)K * 36, cudaMemcpyHostToDevice, 0));
    if (Hinv) B2G_CUDA(cudaMemcpyAsync(dHi, Hinv, (size_t)K * 36, cudaMemcpyHostToDevice, 0));
    if ((rc = launch_score(ctx, dp, M, dH, dHi, nullptr, 0, K, K, 1, th, method, dc, 0))) return rc;
    B2G_CUDA(cudaMemcpyAsync(counts, dc, (size_t)K * 4, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_loss_host(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv, int method, float *err)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(pts && H && err && M >= 0);
    B2G_ARG(method == B200GEO_FWD || Hinv != nullptr);
    if (M == 0) return B200GEO_OK;
    float *dp, *dH, *de;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 16, (void **)&dp))) return rc;
    if ((rc = scratch(ctx, S_H, 72, (void **)&dH))) return rc;
    if ((rc = scratch(ctx, S_MISC, (size_t)M * 4, (void **)&de))) return rc;
    B2G_CUDA(cudaMemcpyAsync(dp, pts, (size_t)M * 16, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dH, H, 36, cudaMemcpyHostToDevice, 0));
    if (Hinv) B2G_CUDA(cudaMemcpyAsync(dH + 9, Hinv, 36, cudaMemcpyHostToDevice, 0));
    int blocks = (M + 255) / 256;
    if (blocks > 8 * ctx->sm_count) blocks = 8 * ctx->sm_count;
    loss_kernel<<<blocks, 256, 0, 0>>>((const float4 *)dp, M, dH, Hinv ? dH + 9 : nullptr, method, de);
    B2G_CHECK_LAUNCH(ctx);
    B2G_CUDA(cudaMemcpy(err, de, (size_t)M * 4, cudaMemcpyDeviceToHost));
    return B200GEO_OK;
}

int b200geo_project_host(b200geo_ctx *ctx, const float *Hm, const float *P, int M, float *out)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(Hm && P && out && M >= 0);
    if (M == 0) return B200GEO_OK;
    float *dp, *dh, *dout;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 12, (void **)&dp))) return rc;
    if ((rc = scratch(ctx, S_H, 36, (void **)&dh))) return rc;
    if ((rc = scratch(ctx, S_MISC, (size_t)M * 12, (void **)&dout))) return rc;
    B2G_CUDA(cudaMemcpyAsync(dp, P, (size_t)M * 12, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dh, Hm, 36, cudaMemcpyHostToDevice, 0));
    int blocks = (M + 255) / 256;
    if (blocks > 8 * ctx->sm_count) blocks = 8 * ctx->sm_count;
    project_kernel<<<blocks, 256, 0, 0>>>(dp, M, dh, dout);
    B2G_CHECK_LAUNCH(ctx);
    B2G_CUDA(cudaMemcpy(out, dout, (size_t)M * 12, cudaMemcpyDeviceToHost));
    return B200GEO_OK;
}

int b200geo_dist_host(b200geo_ctx *ctx, const float *a, const float *b, int rows, int M, float *out)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(a && b && out && rows > 0 && M >= 0);
    if (M == 0) return B200GEO_OK;
    float *da, *dout;
    int rc;
    size_t n = (size_t)rows * M;
    if ((rc = scratch(ctx, S_PTS, n * 8, (void **)&da))) return rc;
    if ((rc = scratch(ctx, S_MISC, (size_t)M * 4, (void **)&dout))) return rc;
    B2G_CUDA(cudaMemcpyAsync(da, a, n * 4, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(da + n, b, n * 4, cudaMemcpyHostToDevice, 0));
    int blocks = (M + 255) / 256;
    if (blocks > 8 * ctx->sm_count) blocks = 8 * ctx->sm_count;
    dist_kernel<<<blocks, 256, 0, 0>>>(da, da + n, rows, M, dout);
    B2G_CHECK_LAUNCH(ctx);
    B2G_CUDA(cudaMemcpy(out, dout, (size_t)M * 4, cudaMemcpyDeviceToHost));
    return B200GEO_OK;
}

int b200geo_ransac_search_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int k_begin, int k_end,
                              float th, double thr, int method, unsigned long long *keys, int zero_keys, int32_t *counts,
                              void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && idx && keys && M > 0 && k_begin >= 0 && k_end >= k_begin);
    B2G_ARG(method == B200GEO_FWD || method == B200GEO_BACKWARD || method == B200GEO_REPROJ);
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_keys) B2G_CUDA(cudaMemsetAsync(keys, 0, 16, st));
    // scratch H / flags are indexed by absolute k: stride = k_end
    return search(ctx, pts, M, idx, k_begin, k_end, k_end, 1, th, thr, method, keys, counts, st);
}

int b200geo_ransac_finish_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, const unsigned long long *keys,
                              float th, int method, double *result, uint8_t *mask, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && idx && keys && result && mask && M > 0);
    return launch_finish(ctx, pts, M, idx, 0, 1, keys, th, method, result, mask, (cudaStream_t)stream);
}

int b200geo_ransac_search_shard_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx_shard, int k_offset, int n_local,
                                    float th, double thr, int method, void *rec, int32_t *counts_shard, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && idx_shard && rec && M > 0 && k_offset >= 0 && n_local >= 0 && (long long)k_offset + n_local <= 0x7fffffffLL);
    B2G_ARG(method == B200GEO_FWD || method == B200GEO_BACKWARD || method == B200GEO_REPROJ);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *dk;
    int rc;
    if ((rc = scratch(ctx, S_KEYS, 16, (void **)&dk))) return rc;
    B2G_CUDA(cudaMemsetAsync(dk, 0, 16, st));
    // the kernels address sample set k as idx[k]: a base pointer k_offset rows before the shard makes the shard rows
    // k_offset .. k_offset + n_local - 1 (nothing outside them is dereferenced)
    const int32_t *idx_virtual = idx_shard - 4 * (ptrdiff_t)k_offset;
    int32_t *counts_virtual = counts_shard ? counts_shard - (ptrdiff_t)k_offset : nullptr;
    if ((rc = search(ctx, pts, M, idx_virtual, k_offset, k_offset + n_local, k_offset + n_local, 1, th, thr, method, dk, counts_virtual, st))) return rc;
    pack_winner_kernel<<<1, 32, 0, st>>>(dk, (const int4 *)idx_virtual, (WinRec *)rec);
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

int b200geo_ransac_finish_recs_dev(b200geo_ctx *ctx, const float *pts, int M, const void *recs, int n_recs, float th, int method,
                                   double *result, uint8_t *mask, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && recs && n_recs > 0 && n_recs <= 4096 && result && mask && M > 0);
    return launch_finish(ctx, pts, M, nullptr, 0, 1, nullptr, th, method, result, mask, (cudaStream_t)stream, (const WinRec *)recs, n_recs);
}

namespace {
// the three non-blocking streams (h2d, compute, d2h) and their events, shared by the host-buffer pipelines
int ensure_pipe(b200geo_ctx *ctx)
{
    if (!ctx->pipe_ready) {
        for (int i = 0; i < 3; ++i) B2G_CUDA(cudaStreamCreateWithFlags(&ctx->pipe_stream[i], cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i)
            for (int b = 0; b < 2; ++b) B2G_CUDA(cudaEventCreateWithFlags(&ctx->pipe_event[i][b], cudaEventDisableTiming));
        ctx->pipe_ready = 1;
    }
    return B200GEO_OK;
}
}  // namespace

int b200geo_ransac_host(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K, float th, double thr,
                        int method, float *out_H_best, double *out_H, uint8_t *out_mask, int32_t *out_winner,
                        int32_t *out_count, int32_t *out_counts, int32_t *out_last_count)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(pts && idx && M > 0 && K > 0);
    B2G_ARG(method == B200GEO_FWD || method == B200GEO_BACKWARD || method == B200GEO_REPROJ);
    float *dp;
    int32_t *di, *dc;
    unsigned long long *dk;
    double *dr;
    uint8_t *dm;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 16, (void **)&dp))) return rc;
    if ((rc = scratch(ctx, S_IDX, (size_t)K * 16, (void **)&di))) return rc;
    if ((rc = scratch(ctx, S_COUNTS, (size_t)K * 4, (void **)&dc))) return rc;
    if ((rc = scratch(ctx, S_KEYS, 16, (void **)&dk))) return rc;
    if ((rc = scratch(ctx, S_RESULT, 32 * 8, (void **)&dr))) return rc;
    if ((rc = scratch(ctx, S_MASK, (size_t)M, (void **)&dm))) return rc;
    if ((rc = ensure_pipe(ctx))) return rc;
    // earlier work of this context may still be running on the legacy stream the scratch blocks are shared with
    B2G_CUDA(cudaStreamSynchronize(0));
    cudaStream_t s_in = ctx->pipe_stream[0], s_k = ctx->pipe_stream[1];
    B2G_CUDA(cudaMemcpyAsync(dp, pts, (size_t)M * 16, cudaMemcpyHostToDevice, s_k));
    B2G_CUDA(cudaMemsetAsync(dk, 0, 16, s_k));
    // Sample sets in two chunks: the first is one full wave of the scoring kernel (2 CTAs per SM x FAST_T x FAST_THREADS
    // hypotheses), the rest crosses the link while it is solved and scored (the keys accumulate).  Equal or smaller
    // chunks were measured slower: each one pays its own partial wave and fix-up launches.
    const long long wave = 2LL * ctx->sm_count * FAST_T * FAST_THREADS;
    int cut[3] = {0, K, K};
    int n_chunks = 1;
    if ((long long)K >= 2 * wave) { cut[1] = (int)wave; n_chunks = 2; }
    for (int c = 0; c < n_chunks; ++c) {
        const int kb = cut[c], ke = cut[c + 1];
        B2G_CUDA(cudaMemcpyAsync(di + (size_t)kb * 4, idx + (size_t)kb * 4, (size_t)(ke - kb) * 16, cudaMemcpyHostToDevice, s_in));
        B2G_CUDA(cudaEventRecord(ctx->pipe_event[0][c & 1], s_in));
        B2G_CUDA(cudaStreamWaitEvent(s_k, ctx->pipe_event[0][c & 1], 0));
        if ((rc = search(ctx, dp, M, di, kb, ke, K, 1, th, thr, method, dk, dc, s_k))) return rc;
    }
    if ((rc = launch_finish(ctx, dp, M, di, K, 1, dk, th, method, dr, dm, s_k))) return rc;
    double res[32];
    B2G_CUDA(cudaMemcpyAsync(res, dr, sizeof res, cudaMemcpyDeviceToHost, s_k));
    if (out_mask) B2G_CUDA(cudaMemcpyAsync(out_mask, dm, (size_t)M, cudaMemcpyDeviceToHost, s_k));
    if (out_counts) B2G_CUDA(cudaMemcpyAsync(out_counts, dc, (size_t)K * 4, cudaMemcpyDeviceToHost, s_k));
    if (out_last_count) B2G_CUDA(cudaMemcpyAsync(out_last_count, dc + (K - 1), 4, cudaMemcpyDeviceToHost, s_k));
    B2G_CUDA(cudaStreamSynchronize(s_k));
    if (out_winner) *out_winner = (int32_t)res[0];
    if (out_count) *out_count = (int32_t)res[1];
    if (out_H_best) for (int j = 0; j < 9; ++j) out_H_best[j] = (float)res[3 + j];
    if (out_H) for (int j = 0; j < 9; ++j) out_H[j] = res[12 + j];
    int status = (int)res[2];
    if (status == B200GEO_ERR_NOMODEL) return fail(status, "no hypothesis has any inlier (the reference fails in np.where(None), ransac.py:206)%s%s");
    if (status == B200GEO_ERR_SINGULAR) return fail(status, "singular refit system (the reference raises LinAlgError)%s%s");
    return status;
}

int b200geo_ransac_batched_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K, int B, float th,
                               double thr, int method, double *results, uint8_t *masks, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(pts && idx && results && masks && M > 0 && K > 0 && B > 0 && B <= 65535);
    B2G_ARG(method == B200GEO_FWD || method == B200GEO_BACKWARD || method == B200GEO_REPROJ);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *dk;
    int rc;
    if ((rc = scratch(ctx, S_KEYS, (size_t)B * 16, (void **)&dk))) return rc;
    B2G_CUDA(cudaMemsetAsync(dk, 0, (size_t)B * 16, st));
    if ((rc = search(ctx, pts, M, idx, 0, K, K, B, th, thr, method, dk, nullptr, st))) return rc;
    return launch_finish(ctx, pts, M, idx, K, B, dk, th, method, results, masks, st);
}

int b200geo_refit_host(b200geo_ctx *ctx, const float *pts, int M, const uint8_t *mask, double *H)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(pts && H && M > 0);
    float *dp;
    uint8_t *dm = nullptr;
    double *dr;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 16, (void **)&dp))) return rc;
    if (mask && (rc = scratch(ctx, S_MASK, (size_t)M, (void **)&dm))) return rc;
    if ((rc = scratch(ctx, S_RESULT, 32 * 8, (void **)&dr))) return rc;
    B2G_CUDA(cudaMemcpyAsync(dp, pts, (size_t)M * 16, cudaMemcpyHostToDevice, 0));
    if (mask) B2G_CUDA(cudaMemcpyAsync(dm, mask, (size_t)M, cudaMemcpyHostToDevice, 0));
    refit_kernel<<<1, REFIT_THREADS, 0, 0>>>((const float4 *)dp, M, dm, dr);
    B2G_CHECK_LAUNCH(ctx);
    double res[10];
    B2G_CUDA(cudaMemcpy(res, dr, sizeof res, cudaMemcpyDeviceToHost));
    for (int j = 0; j < 9; ++j) H[j] = res[1 + j];
    if ((int)res[0] != 0) return fail(B200GEO_ERR_SINGULAR, "singular least-squares system (the reference raises LinAlgError)%s%s");
    return B200GEO_OK;
}

int b200geo_homography_svd_host(b200geo_ctx *ctx, const float *pts, int M, double *H)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(pts && H && M > 0);
    float *dp;
    double *dr;
    int rc;
    if ((rc = scratch(ctx, S_PTS, (size_t)M * 16, (void **)&dp))) return rc;
    if ((rc = scratch(ctx, S_RESULT, 32 * 8, (void **)&dr))) return rc;
    B2G_CUDA(cudaMemcpyAsync(dp, pts, (size_t)M * 16, cudaMemcpyHostToDevice, 0));
    homography_svd_kernel<<<1, REFIT_THREADS, 0, 0>>>((const float4 *)dp, M, dr);
    B2G_CHECK_LAUNCH(ctx);
    double res[10];
    B2G_CUDA(cudaMemcpy(res, dr, sizeof res, cudaMemcpyDeviceToHost));
    for (int j = 0; j < 9; ++j) H[j] = res[1 + j];
    if ((int)res[0] != 0) return fail(B200GEO_ERR_SINGULAR, "degenerate homogeneous system%s%s");
    return B200GEO_OK;
}

int b200geo_warp_jobs_dev(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs_dev, int n_jobs,
                          int max_ow, int max_oh, int channels, int src_type, int dst_type, int sampler, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(src && dst && jobs_dev && n_jobs >= 0);
    B2G_ARG(sampler == B200GEO_NN || sampler == B200GEO_BILINEAR);
    B2G_ARG(dst_type == B200GEO_U8 || dst_type == B200GEO_F32 || dst_type == B200GEO_F64);
    return launch_warp(ctx, src, dst, jobs_dev, n_jobs, max_ow, max_oh, channels, src_type, dst_type, sampler, (cudaStream_t)stream);
}

int b200geo_warp_batch_dev(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs_dev, int n_jobs,
                           int max_ow, int max_oh, int channels, int src_type, int dst_type, int sampler,
                           const b200geo_src_layout *layout, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(src && dst && jobs_dev && n_jobs >= 0);
    B2G_ARG(sampler == B200GEO_NN || sampler == B200GEO_BILINEAR);
    B2G_ARG(dst_type == B200GEO_U8 || dst_type == B200GEO_F32 || dst_type == B200GEO_F64);
    return launch_warp(ctx, src, dst, jobs_dev, n_jobs, max_ow, max_oh, channels, src_type, dst_type, sampler, (cudaStream_t)stream, layout);
}

int b200geo_warp_dev(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs, int n_jobs, int channels,
                     int src_type, int dst_type, int sampler, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(src && dst && jobs && n_jobs >= 0);
    B2G_ARG(sampler == B200GEO_NN || sampler == B200GEO_BILINEAR);
    B2G_ARG(dst_type == B200GEO_U8 || dst_type == B200GEO_F32 || dst_type == B200GEO_F64);
    if (n_jobs == 0) return B200GEO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    b200geo_warp_job *dj;
    void *hp;
    int rc;
    size_t bytes = sizeof(b200geo_warp_job) * (size_t)n_jobs;
    if ((rc = scratch(ctx, S_JOBS, bytes, (void **)&dj))) return rc;
    if ((rc = pinned(ctx, bytes, &hp))) return rc;
    // the pinned block is reused by the next call: wait for earlier copies out of it
    B2G_CUDA(cudaStreamSynchronize(st));
    memcpy(hp, jobs, bytes);
    int max_ow = 0, max_oh = 0;
    // the jobs are visible here: describe the source as one 2-D byte array when they share a pitch
    b200geo_src_layout lay = {jobs[0].src_pitch, 0, 0, 0};
    bool uniform = true, same = channels == 3 && src_type == B200GEO_U8;     // same: equal images at multiples of their size -> 3-D tensor map
    for (int i = 0; i < n_jobs; ++i) {
        B2G_ARG(jobs[i].iw > 0 && jobs[i].ih > 0 && jobs[i].ow >= 0 && jobs[i].oh >= 0);
        B2G_ARG((size_t)jobs[i].src_pitch >= (size_t)jobs[i].iw * channels * type_size(src_type));
        B2G_ARG((size_t)jobs[i].dst_pitch >= (size_t)jobs[i].ow * channels * type_size(dst_type));
        if (jobs[i].ow > max_ow) max_ow = jobs[i].ow;
        if (jobs[i].oh > max_oh) max_oh = jobs[i].oh;
        if (jobs[i].src_pitch != lay.pitch || lay.pitch <= 0 || jobs[i].src_offset < 0 || jobs[i].src_offset % lay.pitch) uniform = false;
        else if (jobs[i].src_offset / lay.pitch + jobs[i].ih > lay.rows) lay.rows = jobs[i].src_offset / lay.pitch + jobs[i].ih;
        if (jobs[i].ih != jobs[0].ih || jobs[i].iw != jobs[0].iw || jobs[i].ih <= 0 || jobs[i].src_offset < 0 ||
            (lay.pitch > 0 && jobs[i].src_offset % (lay.pitch * jobs[0].ih))) same = false;
    }
    if (uniform && same && lay.rows % jobs[0].ih == 0 && 3LL * jobs[0].iw <= lay.pitch) {
        lay.image_rows = jobs[0].ih;
        lay.row_bytes = 3LL * jobs[0].iw;
    }
    B2G_CUDA(cudaMemcpyAsync(dj, hp, bytes, cudaMemcpyHostToDevice, st));
    return launch_warp(ctx, src, dst, dj, n_jobs, max_ow, max_oh, channels, src_type, dst_type, sampler, st, uniform ? &lay : nullptr);
}

int b200geo_warp_batch_host(b200geo_ctx *ctx, const void *src_host, void *dst_host, const b200geo_warp_job *jobs, int n_jobs,
                            int channels, int src_type, int dst_type, int sampler, int chunk_jobs)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(src_host && dst_host && jobs && n_jobs >= 0);
    B2G_ARG(sampler == B200GEO_NN || sampler == B200GEO_BILINEAR);
    B2G_ARG(dst_type == B200GEO_U8 || dst_type == B200GEO_F32 || dst_type == B200GEO_F64);
    if (n_jobs == 0) return B200GEO_OK;
    if (chunk_jobs <= 0) chunk_jobs = n_jobs >= 128 ? 16 : 8;     // measured on 1080p frames: 8 is best for 64 jobs, 16 for 256
    if (chunk_jobs > n_jobs) chunk_jobs = n_jobs;
    int prc;
    if ((prc = ensure_pipe(ctx))) return prc;
    B2G_CUDA(cudaStreamSynchronize(0));      // earlier work of this context (ordered onto the legacy stream above) shares the scratch blocks
    cudaStream_t s_in = ctx->pipe_stream[0], s_k = ctx->pipe_stream[1], s_out = ctx->pipe_stream[2];
    // device images of one chunk are packed back to back (256-byte aligned); jobs are rebased onto them
    const int n_chunks = (n_jobs + chunk_jobs - 1) / chunk_jobs;
    std::vector<b200geo_warp_job> dj(jobs, jobs + n_jobs);
    size_t max_src = 0, max_dst = 0;
    bool uniform = true, same = channels == 3 && src_type == B200GEO_U8;
    b200geo_src_layout lay = {jobs[0].src_pitch, 0, 0, 0};
    for (int j = 0; j < n_jobs; ++j)
        if (jobs[j].ih != jobs[0].ih || jobs[j].iw != jobs[0].iw || jobs[j].src_pitch != jobs[0].src_pitch) same = false;
    if (same && (((size_t)jobs[0].src_pitch * jobs[0].ih) & 255) == 0 && 3LL * jobs[0].iw <= jobs[0].src_pitch) {
        // equal images whose size is a multiple of the 256-byte packing granule: each chunk is (row_bytes x image_rows x images)
        lay.image_rows = jobs[0].ih;
        lay.row_bytes = 3LL * jobs[0].iw;
    }
    for (int c = 0; c < n_chunks; ++c) {
        size_t so = 0, dofs = 0;
        for (int j = c * chunk_jobs; j < n_jobs && j < (c + 1) * chunk_jobs; ++j) {
            B2G_ARG(jobs[j].ow > 0 && jobs[j].oh > 0 && jobs[j].iw > 0 && jobs[j].ih > 0 && jobs[j].src_pitch > 0 && jobs[j].dst_pitch > 0);
            B2G_ARG((size_t)jobs[j].src_pitch >= (size_t)jobs[j].iw * channels * type_size(src_type));
            B2G_ARG((size_t)jobs[j].dst_pitch >= (size_t)jobs[j].ow * channels * type_size(dst_type));
            dj[j].src_offset = (int64_t)so;
            dj[j].dst_offset = (int64_t)dofs;
            if (jobs[j].src_pitch != lay.pitch || (so % (size_t)lay.pitch)) uniform = false;
            so += ((size_t)jobs[j].src_pitch * jobs[j].ih + 255) & ~(size_t)255;
            dofs += ((size_t)jobs[j].dst_pitch * jobs[j].oh + 255) & ~(size_t)255;
        }
        if (so > max_src) max_src = so;
        if (dofs > max_dst) max_dst = dofs;
    }
    // (the 256-byte rounding keeps every image 16-byte aligned; a uniform pitch that divides the
    // rounded size makes the chunk one 2-D array for the TMA path)
    uint8_t *dsrc, *ddst;
    b200geo_warp_job *djobs;
    void *hp;
    int rc;
    if ((rc = scratch(ctx, S_PIPE_SRC, 2 * max_src, (void **)&dsrc))) return rc;
    if ((rc = scratch(ctx, S_PIPE_DST, 2 * max_dst, (void **)&ddst))) return rc;
    if ((rc = scratch(ctx, S_PIPE_JOBS, sizeof(b200geo_warp_job) * (size_t)n_jobs, (void **)&djobs))) return rc;
    if ((rc = pinned(ctx, sizeof(b200geo_warp_job) * (size_t)n_jobs, &hp))) return rc;
    B2G_CUDA(cudaStreamSynchronize(s_k));
    memcpy(hp, dj.data(), sizeof(b200geo_warp_job) * (size_t)n_jobs);
    B2G_CUDA(cudaMemcpyAsync(djobs, hp, sizeof(b200geo_warp_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, s_k));
    for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1, j0 = c * chunk_jobs, j1 = (j0 + chunk_jobs < n_jobs) ? j0 + chunk_jobs : n_jobs;
        uint8_t *cs = dsrc + (size_t)b * max_src, *cd = ddst + (size_t)b * max_dst;
        // stage in: the slot's previous warp (chunk c-2) must have finished reading it
        if (c >= 2) B2G_CUDA(cudaStreamWaitEvent(s_in, ctx->pipe_event[1][b], 0));
        int max_ow = 0, max_oh = 0;
        long long rows = 0;
        for (int j = j0; j < j1; ++j) {
            B2G_CUDA(cudaMemcpyAsync(cs + dj[j].src_offset, (const uint8_t *)src_host + jobs[j].src_offset,
                                     (size_t)jobs[j].src_pitch * jobs[j].ih, cudaMemcpyHostToDevice, s_in));
            if (jobs[j].ow > max_ow) max_ow = jobs[j].ow;
            if (jobs[j].oh > max_oh) max_oh = jobs[j].oh;
            if (uniform) rows = dj[j].src_offset / lay.pitch + jobs[j].ih;
        }
        B2G_CUDA(cudaEventRecord(ctx->pipe_event[0][b], s_in));
        // warp: needs the inputs, and the slot's previous outputs (chunk c-2) copied out
        B2G_CUDA(cudaStreamWaitEvent(s_k, ctx->pipe_event[0][b], 0));
        if (c >= 2) B2G_CUDA(cudaStreamWaitEvent(s_k, ctx->pipe_event[2][b], 0));
        lay.rows = rows;
        if ((rc = launch_warp(ctx, cs, cd, djobs + j0, j1 - j0, max_ow, max_oh, channels, src_type, dst_type, sampler, s_k,
                              uniform ? &lay : nullptr)))
            return rc;
        B2G_CUDA(cudaEventRecord(ctx->pipe_event[1][b], s_k));
        // stage out
        B2G_CUDA(cudaStreamWaitEvent(s_out, ctx->pipe_event[1][b], 0));
        for (int j = j0; j < j1; ++j)
            B2G_CUDA(cudaMemcpyAsync((uint8_t *)dst_host + jobs[j].dst_offset, cd + dj[j].dst_offset,
                                     (size_t)jobs[j].dst_pitch * jobs[j].oh, cudaMemcpyDeviceToHost, s_out));
        B2G_CUDA(cudaEventRecord(ctx->pipe_event[2][b], s_out));
    }
    B2G_CUDA(cudaStreamSynchronize(s_out));
    B2G_CUDA(cudaStreamSynchronize(s_k));
    return B200GEO_OK;
}

int b200geo_warp_host(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *job, int channels,
                      int src_type, int dst_type, int sampler)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(src && dst && job);
    B2G_ARG(job->ow > 0 && job->oh > 0 && job->iw > 0 && job->ih > 0);
    size_t sbytes = (size_t)job->src_pitch * (size_t)job->ih;
    size_t dbytes = (size_t)job->dst_pitch * (size_t)job->oh;
    B2G_ARG((size_t)job->src_pitch >= (size_t)job->iw * channels * type_size(src_type));
    B2G_ARG((size_t)job->dst_pitch >= (size_t)job->ow * channels * type_size(dst_type));
    void *ds, *dd;
    int rc;
    if ((rc = scratch(ctx, S_SRC, sbytes, &ds))) return rc;
    if ((rc = scratch(ctx, S_DST, dbytes, &dd))) return rc;
    b200geo_warp_job j = *job;
    j.src_offset = 0;
    j.dst_offset = 0;
    B2G_CUDA(cudaMemcpyAsync(ds, (const uint8_t *)src + job->src_offset, sbytes, cudaMemcpyHostToDevice, 0));
    if ((rc = b200geo_warp_dev(ctx, ds, dd, &j, 1, channels, src_type, dst_type, sampler, nullptr))) return rc;
    B2G_CUDA(cudaMemcpyAsync((uint8_t *)dst + job->dst_offset, dd, dbytes, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_match_hamming_host(b200geo_ctx *ctx, const uint8_t *descA, int nA, const uint8_t *descB, int nB, int desc_bytes,
                               int cross_check, int32_t *train_idx, int32_t *dist)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(nA >= 0 && nB >= 0 && desc_bytes > 0 && desc_bytes % 4 == 0 && desc_bytes <= 4 * HM_MAX_WORDS);
    B2G_ARG((nA == 0 || (descA && train_idx && dist)) && (nB == 0 || descB));
    if (nA == 0) return B200GEO_OK;
    if (nB == 0) {
        for (int i = 0; i < nA; ++i) { train_idx[i] = -1; dist[i] = 0; }
        return B200GEO_OK;
    }
    const int words = desc_bytes / 4;
    const size_t abytes = (size_t)nA * desc_bytes, bbytes = (size_t)nB * desc_bytes;
    uint8_t *dd;
    int32_t *dr;
    int rc;
    if ((rc = scratch(ctx, S_SRC, ((abytes + 15) & ~(size_t)15) + bbytes, (void **)&dd))) return rc;
    if ((rc = scratch(ctx, S_MISC, sizeof(int32_t) * (2 * (size_t)nA + (size_t)nB), (void **)&dr))) return rc;
    uint8_t *dA = dd, *dB = dd + ((abytes + 15) & ~(size_t)15);
    int32_t *bestA = dr, *distA = dr + nA, *bestB = dr + 2 * (size_t)nA;
    B2G_CUDA(cudaMemcpyAsync(dA, descA, abytes, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dB, descB, bbytes, cudaMemcpyHostToDevice, 0));
    const int nmax = nA > nB ? nA : nB;
    const dim3 grid((nmax + HM_THREADS - 1) / HM_THREADS, cross_check ? 2 : 1);
    const size_t smem = (size_t)HM_TILE * words * sizeof(uint32_t);
    if (words == 8)
        hamming_best_kernel<8><<<grid, HM_THREADS, smem, 0>>>((const uint32_t *)dA, nA, (const uint32_t *)dB, nB, words, bestA, distA, bestB);
    else
        hamming_best_kernel<0><<<grid, HM_THREADS, smem, 0>>>((const uint32_t *)dA, nA, (const uint32_t *)dB, nB, words, bestA, distA, bestB);
    B2G_CHECK_LAUNCH(ctx);
    if (cross_check) {
        hamming_cross_kernel<<<(nA + 255) / 256, 256, 0, 0>>>(bestA, bestB, nA);
        B2G_CHECK_LAUNCH(ctx);
    }
    B2G_CUDA(cudaMemcpyAsync(train_idx, bestA, sizeof(int32_t) * (size_t)nA, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaMemcpyAsync(dist, distA, sizeof(int32_t) * (size_t)nA, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_stitch_compose_dev(b200geo_ctx *ctx, const void *imgT, const void *imgQ, void *canvas, const b200geo_warp_job *job,
                               const b200geo_compose *p, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(imgT && imgQ && canvas && job && p);
    B2G_ARG(p->fw > 0 && p->fh > 0 && p->qw > 0 && p->qh > 0 && job->iw > 0 && job->ih > 0 && job->ow >= 0 && job->oh >= 0);
    B2G_ARG(p->blend == 0 || p->blend == 1);
    B2G_ARG((size_t)p->canvas_pitch >= (size_t)p->fw * 3 && (size_t)p->q_pitch >= (size_t)p->qw * 3 && (size_t)job->src_pitch >= (size_t)job->iw * 3);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((p->fw + 31) / 32, (p->fh + 7) / 8), block(32, 8);
    B2G_ARG(grid.y <= 65535);
    if (p->blend) stitch_compose_kernel<1><<<grid, block, 0, st>>>((const uint8_t *)imgT, (const uint8_t *)imgQ, (uint8_t *)canvas, *job, *p);
    else stitch_compose_kernel<0><<<grid, block, 0, st>>>((const uint8_t *)imgT, (const uint8_t *)imgQ, (uint8_t *)canvas, *job, *p);
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

int b200geo_stitch_jobs_dev(b200geo_ctx *ctx, const double *results, int B, int ih, int iw, int canvas_h, int canvas_w,
                            b200geo_warp_job *jobs_dev, int32_t *boxes_dev, void *stream)
{
    B2G_ENTER(ctx);
    stream_order(ctx, (cudaStream_t)stream);
    B2G_ARG(results && jobs_dev && boxes_dev && B > 0 && ih > 0 && iw > 0 && canvas_h > 0 && canvas_w > 0);
    const long long sp = 3LL * iw, dp = 3LL * canvas_w;
    stitch_jobs_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(results, B, ih, iw, canvas_h, canvas_w, sp, sp * ih, dp, dp * canvas_h,
                                                                      jobs_dev, boxes_dev);
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

int b200geo_stitch_compose_host(b200geo_ctx *ctx, const void *imgT, const void *imgQ, void *canvas, const b200geo_warp_job *job,
                                const b200geo_compose *p)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(imgT && imgQ && canvas && job && p && p->fh > 0 && p->qh > 0 && job->ih > 0);
    const size_t tbytes = (size_t)job->src_pitch * (size_t)job->ih, qbytes = (size_t)p->q_pitch * (size_t)p->qh;
    const size_t cbytes = (size_t)p->canvas_pitch * (size_t)p->fh;
    void *dt, *dq, *dc;
    int rc;
    if ((rc = scratch(ctx, S_SRC, tbytes, &dt))) return rc;
    if ((rc = scratch(ctx, S_MISC, qbytes, &dq))) return rc;
    if ((rc = scratch(ctx, S_DST, cbytes, &dc))) return rc;
    b200geo_warp_job j = *job;
    j.src_offset = 0;
    B2G_CUDA(cudaMemcpyAsync(dt, (const uint8_t *)imgT + job->src_offset, tbytes, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dq, imgQ, qbytes, cudaMemcpyHostToDevice, 0));
    if ((rc = b200geo_stitch_compose_dev(ctx, dt, dq, dc, &j, p, nullptr))) return rc;
    B2G_CUDA(cudaMemcpyAsync(canvas, dc, cbytes, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_sample_host(b200geo_ctx *ctx, const void *src, int ih, int iw, int channels, int src_type, const double *zx,
                        const double *zy, long long N, int bh, int bw, int sampler, int dst_type, void *dst)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(src && zx && zy && dst && ih > 0 && iw > 0 && N >= 0);
    B2G_ARG(sampler == B200GEO_NN || sampler == B200GEO_BILINEAR);
    B2G_ARG(dst_type == B200GEO_U8 || dst_type == B200GEO_F32 || dst_type == B200GEO_F64);
    B2G_ARG((channels == 3 || channels == 4) && (src_type == B200GEO_U8 || src_type == B200GEO_F32));
    if (N == 0) return B200GEO_OK;
    size_t sbytes = (size_t)ih * iw * channels * type_size(src_type);
    size_t dbytes = (size_t)N * channels * type_size(dst_type);
    void *ds, *dd;
    double *dz;
    int rc;
    if ((rc = scratch(ctx, S_SRC, sbytes, &ds))) return rc;
    if ((rc = scratch(ctx, S_DST, dbytes, &dd))) return rc;
    if ((rc = scratch(ctx, S_MISC, (size_t)N * 16, (void **)&dz))) return rc;
    B2G_CUDA(cudaMemcpyAsync(ds, src, sbytes, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dz, zx, (size_t)N * 8, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dz + N, zy, (size_t)N * 8, cudaMemcpyHostToDevice, 0));
    if (src_type == B200GEO_U8 && channels == 3) rc = launch_sample_t<uint8_t, 3>(ctx, ds, ih, iw, bh, bw, dz, dz + N, N, sampler, dst_type, dd, 0);
    else if (src_type == B200GEO_U8) rc = launch_sample_t<uint8_t, 4>(ctx, ds, ih, iw, bh, bw, dz, dz + N, N, sampler, dst_type, dd, 0);
    else if (channels == 3) rc = launch_sample_t<float, 3>(ctx, ds, ih, iw, bh, bw, dz, dz + N, N, sampler, dst_type, dd, 0);
    else rc = launch_sample_t<float, 4>(ctx, ds, ih, iw, bh, bw, dz, dz + N, N, sampler, dst_type, dd, 0);
    if (rc) return rc;
    B2G_CUDA(cudaMemcpyAsync(dst, dd, dbytes, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_remap_linear_host(b200geo_ctx *ctx, const uint8_t *src, int ih, int iw, int channels, const float *map_x, const float *map_y,
                              int oh, int ow, uint8_t *dst)
{
    B2G_ENTER(ctx);
    stream_order(ctx, 0);
    B2G_ARG(ih > 0 && iw > 0 && oh >= 0 && ow >= 0 && (channels == 1 || channels == 2 || channels == 4));
    if (oh == 0 || ow == 0) return B200GEO_OK;
    B2G_ARG(src && map_x && map_y && dst);
    const size_t sbytes = (size_t)ih * iw * channels, dbytes = (size_t)oh * ow * channels, mbytes = (size_t)oh * ow * sizeof(float);
    uint8_t *ds, *dd;
    float *dm;
    int rc;
    if ((rc = scratch(ctx, S_SRC, sbytes, (void **)&ds))) return rc;
    if ((rc = scratch(ctx, S_DST, dbytes, (void **)&dd))) return rc;
    if ((rc = scratch(ctx, S_MISC, 2 * mbytes, (void **)&dm))) return rc;
    B2G_CUDA(cudaMemcpyAsync(ds, src, sbytes, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dd, dst, dbytes, cudaMemcpyHostToDevice, 0));          // transparent pixels keep the caller's bytes
    B2G_CUDA(cudaMemcpyAsync(dm, map_x, mbytes, cudaMemcpyHostToDevice, 0));
    B2G_CUDA(cudaMemcpyAsync(dm + (size_t)oh * ow, map_y, mbytes, cudaMemcpyHostToDevice, 0));
    const dim3 grid((ow + 255) / 256, oh);
    B2G_ARG(oh <= 65535);
    const long long sp = (long long)iw * channels, dp = (long long)ow * channels;
    if (channels == 4) remap_linear_kernel<4><<<grid, 256, 0, 0>>>(ds, ih, iw, sp, dm, dm + (size_t)oh * ow, oh, ow, dd, dp);
    else if (channels == 2) remap_linear_kernel<2><<<grid, 256, 0, 0>>>(ds, ih, iw, sp, dm, dm + (size_t)oh * ow, oh, ow, dd, dp);
    else remap_linear_kernel<1><<<grid, 256, 0, 0>>>(ds, ih, iw, sp, dm, dm + (size_t)oh * ow, oh, ow, dd, dp);
    B2G_CHECK_LAUNCH(ctx);
    B2G_CUDA(cudaMemcpyAsync(dst, dd, dbytes, cudaMemcpyDeviceToHost, 0));
    B2G_CUDA(cudaStreamSynchronize(0));
    return B200GEO_OK;
}

int b200geo_jpeg_info(b200geo_ctx *ctx, const uint8_t *data, size_t nbytes, int *height, int *width, int *components)
{
    B2G_ENTER(ctx);
    B2G_ARG(data && nbytes > 0);
    JpegState *st;
    int rc;
    if ((rc = jpeg_state(ctx, &st))) return rc;
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    B2G_JPEG(st, GetImageInfo, st->handle, data, nbytes, &nc, &sub, ws, hs);
    if (height) *height = hs[0];
    if (width) *width = ws[0];
    if (components) *components = nc;
    return B200GEO_OK;
}

int b200geo_jpeg_decode(b200geo_ctx *ctx, const uint8_t *data, size_t nbytes, int order, void *dst, long long dst_pitch, int dst_on_device,
                        void *stream)
{
    B2G_ENTER(ctx);
    cudaStream_t cs = (cudaStream_t)stream;
    stream_order(ctx, cs);
    B2G_ARG(data && nbytes > 0 && dst && (order == B200GEO_BGR || order == B200GEO_RGB));
    JpegState *st;
    int rc;
    if ((rc = jpeg_state(ctx, &st))) return rc;
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    B2G_JPEG(st, GetImageInfo, st->handle, data, nbytes, &nc, &sub, ws, hs);
    const int h = hs[0], w = ws[0];
    B2G_ARG(h > 0 && w > 0 && dst_pitch >= 3LL * w);
    nvjpegImage_t im;
    memset(&im, 0, sizeof im);
    if (dst_on_device) {
        im.channel[0] = static_cast<unsigned char *>(dst);
        im.pitch[0] = (size_t)dst_pitch;
    } else {
        void *d;
        if ((rc = scratch(ctx, S_DST, (size_t)h * w * 3, &d))) return rc;
        im.channel[0] = static_cast<unsigned char *>(d);
        im.pitch[0] = (size_t)w * 3;
    }
    B2G_JPEG(st, Decode, st->handle, st->dec, data, nbytes, order == B200GEO_BGR ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI, &im, cs);
    if (!dst_on_device) {
        B2G_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, im.channel[0], im.pitch[0], (size_t)w * 3, (size_t)h, cudaMemcpyDeviceToHost, cs));
        B2G_CUDA(cudaStreamSynchronize(cs));
    }
    return B200GEO_OK;
}

int b200geo_jpeg_encode(b200geo_ctx *ctx, const void *src, int height, int width, long long src_pitch, int order, int src_on_device, int quality,
                        uint8_t *out, size_t capacity, size_t *nbytes, void *stream)
{
    B2G_ENTER(ctx);
    cudaStream_t cs = (cudaStream_t)stream;
    stream_order(ctx, cs);
    B2G_ARG(src && nbytes && height > 0 && width > 0 && src_pitch >= 3LL * width && quality >= 1 && quality <= 100);
    B2G_ARG(order == B200GEO_BGR || order == B200GEO_RGB);
    JpegState *st;
    int rc;
    if ((rc = jpeg_state(ctx, &st))) return rc;
    nvjpegImage_t im;
    memset(&im, 0, sizeof im);
    if (src_on_device) {
        im.channel[0] = const_cast<unsigned char *>(static_cast<const unsigned char *>(src));
        im.pitch[0] = (size_t)src_pitch;
    } else {
        void *d;
        if ((rc = scratch(ctx, S_SRC, (size_t)height * width * 3, &d))) return rc;
        B2G_CUDA(cudaMemcpy2DAsync(d, (size_t)width * 3, src, (size_t)src_pitch, (size_t)width * 3, (size_t)height, cudaMemcpyHostToDevice, cs));
        im.channel[0] = static_cast<unsigned char *>(d);
        im.pitch[0] = (size_t)width * 3;
    }
    // cv2.imwrite's defaults: quality 95 unless given, 4:2:0 chroma, baseline Huffman tables
    B2G_JPEG(st, EncoderParamsSetQuality, st->params, quality, cs);
    B2G_JPEG(st, EncoderParamsSetOptimizedHuffman, st->params, 0, cs);
    B2G_JPEG(st, EncoderParamsSetSamplingFactors, st->params, NVJPEG_CSS_420, cs);
    B2G_JPEG(st, EncodeImage, st->handle, st->enc, st->params, &im, order == B200GEO_BGR ? NVJPEG_INPUT_BGRI : NVJPEG_INPUT_RGBI, width, height, cs);
    size_t len = 0;
    B2G_JPEG(st, EncodeRetrieveBitstream, st->handle, st->enc, nullptr, &len, cs);
    *nbytes = len;
    if (!out) return B200GEO_OK;                       // size query
    if (capacity < len) return fail(B200GEO_ERR_ARG, "jpeg_encode: output buffer too small%s%s");
    B2G_JPEG(st, EncodeRetrieveBitstream, st->handle, st->enc, out, &len, cs);
    B2G_CUDA(cudaStreamSynchronize(cs));
    *nbytes = len;
    return B200GEO_OK;
}

// KeyPointsFilter::retainBest (features2d/src/keypoint.cpp): the n best by response plus every point tied with the n-th,
// in the element order std::nth_element and std::partition of libstdc++ leave behind (OpenCV's output order inherits it).
namespace {
struct OrbKp { int x, y; float response; float harris; };
void orb_retain_best(std::vector<OrbKp> &kps, int n)
{
    if (n >= 0 && kps.size() > (size_t)n) {
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const OrbKp &a, const OrbKp &b) { return a.response > b.response; });
        const float ambiguous = kps[n - 1].response;
        auto new_end = std::partition(kps.begin() + n, kps.end(), [ambiguous](const OrbKp &k) { return k.response >= ambiguous; });
        kps.resize(new_end - kps.begin());
    }
}
}  // namespace

int b200geo_orb_host(b200geo_ctx *ctx, const uint8_t *gray, int height, int width, long long pitch, int channels, int nfeatures, float *keypoints,
                     uint8_t *descriptors, int capacity, int *n_out)
{
    try {
        B2G_ENTER(ctx);
        stream_order(ctx, 0);
        B2G_ARG(gray && n_out && height > 0 && width > 0 && height < 32768 && width < 32768 && (channels == 1 || channels == 3) && pitch >= (long long)width * channels &&
                nfeatures >= 0 && capacity >= 0);
        B2G_ARG(capacity == 0 || (keypoints && descriptors));
        *n_out = 0;
        // ---- level geometry: ORB_Impl::detectAndCompute / computeKeyPoints (orb.cpp), default parameters ----
        const double scale_factor = (double)1.2f;                      // ORB::create(float scaleFactor = 1.2f) stored as double
        float scales[ORB_LEVELS];
        OrbLevels L;
        memset(&L, 0, sizeof L);
        L.n = ORB_LEVELS;
        long long bytes = 0;
        unsigned cand_total = 0;
        int max_w = 0, max_h = 0;
        for (int lv = 0; lv < ORB_LEVELS; ++lv) {
            scales[lv] = (float)std::pow(scale_factor, (double)lv);
            const float inv = 1.0f / scales[lv];
            OrbLevel &o = L.lv[lv];
            o.w = (int)lrintf((float)width * inv);
            o.h = (int)lrintf((float)height * inv);
            B2G_ARG(o.w >= 1 && o.h >= 1);
            o.pitch = (o.w + 2 * ORB_BORDER + 15) & ~15;
            o.pw = lv ? L.lv[lv - 1].w : width;
            o.ph = lv ? L.lv[lv - 1].h : height;
            o.ofs = bytes;
            bytes += (long long)o.pitch * (o.h + 2 * ORB_BORDER);
            o.cand_ofs = cand_total;
            o.cand_cap = (unsigned)(((o.w + 1) / 2) * ((o.h + 1) / 2));    // strict 3 x 3 maxima: at most one per 2 x 2 block
            cand_total += o.cand_cap;
            max_w = std::max(max_w, o.w);
            max_h = std::max(max_h, o.h);
        }
        OrbRows R;
        memset(&R, 0, sizeof R);
        unsigned rows_total = 0, words_total = 0;
        for (int lv = 0; lv < ORB_LEVELS; ++lv) {
            const int rows = std::max(L.lv[lv].h - 2 * ORB_EDGE, 0), cols = std::max(L.lv[lv].w - 2 * ORB_EDGE, 0);
            R.row_ofs[lv] = rows_total;
            R.word_ofs[lv] = words_total;
            R.words_per_row[lv] = (cols + 31) / 32;
            rows_total += (unsigned)rows;
            words_total += (unsigned)rows * (unsigned)R.words_per_row[lv];
        }
        int per_level[ORB_LEVELS];
        {
            const float factor = (float)(1.0 / scale_factor);
            float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)ORB_LEVELS));
            int sum = 0;
            for (int lv = 0; lv < ORB_LEVELS - 1; ++lv) {
                per_level[lv] = (int)lrintf(nd);
                sum += per_level[lv];
                nd *= factor;
            }
            per_level[ORB_LEVELS - 1] = std::max(nfeatures - sum, 0);
        }
        OrbUmax um;
        {
            // the end of each row of the circular patch (computeKeyPoints)
            const int hp = ORB_HALF_PATCH;
            const int vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
            for (int v = 0; v <= hp + 1; ++v) um.u[v] = 0;
            for (int v = 0; v <= vmax; ++v) um.u[v] = (int)lrint(std::sqrt((double)hp * hp - v * v));
            for (int v = hp, v0 = 0; v >= vmin; --v) {
                while (um.u[v0] == um.u[v0 + 1]) ++v0;
                um.u[v] = v0;
                ++v0;
            }
        }
        // ---- device buffers ----
        uint8_t *d_in, *d_pyr, *d_score, *d_blur;
        OrbCandidate *d_cand;
        unsigned *d_counts;
        int rc;
        if ((rc = scratch(ctx, S_SRC, (size_t)height * width * channels, (void **)&d_in))) return rc;
        if ((rc = scratch(ctx, S_ORB_PYR, (size_t)bytes, (void **)&d_pyr))) return rc;
        if ((rc = scratch(ctx, S_ORB_SCORE, (size_t)bytes, (void **)&d_score))) return rc;
        if ((rc = scratch(ctx, S_ORB_BLUR, (size_t)bytes, (void **)&d_blur))) return rc;
        if ((rc = scratch(ctx, S_ORB_CAND, (size_t)cand_total * sizeof(OrbCandidate), (void **)&d_cand))) return rc;
        unsigned *d_rows, *d_masks;
        if ((rc = scratch(ctx, S_MISC, 64 + sizeof(unsigned) * ((size_t)rows_total + words_total + 2), (void **)&d_counts))) return rc;
        d_rows = d_counts + 16;
        d_masks = d_rows + rows_total + 1;
        B2G_CUDA(cudaMemcpy2DAsync(d_in, (size_t)width * channels, gray, (size_t)pitch, (size_t)width * channels, (size_t)height, cudaMemcpyHostToDevice, 0));
        B2G_CUDA(cudaMemsetAsync(d_score, 0, (size_t)bytes, 0));
        B2G_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned) * ORB_LEVELS, 0));
        for (int lv = 0; lv < ORB_LEVELS; ++lv) {
            const OrbLevel &o = L.lv[lv];
            const dim3 grid((o.w + 2 * ORB_BORDER + 255) / 256, o.h + 2 * ORB_BORDER);
            if (lv == 0) orb_level_kernel<<<grid, 256, 0, 0>>>(d_in, (long long)width * channels, width, height, d_pyr + o.ofs, o.pitch, o.w, o.h, channels);
            else {
                const OrbLevel &q = L.lv[lv - 1];
                orb_level_kernel<<<grid, 256, 0, 0>>>(d_pyr + q.ofs + (long long)ORB_BORDER * q.pitch + ORB_BORDER, (long long)q.pitch, q.w, q.h, d_pyr + o.ofs, o.pitch,
                                                      o.w, o.h, 0);
            }
            B2G_CHECK_LAUNCH(ctx);
        }
        if (max_w > 6 && max_h > 6) {
            orb_fast_kernel<<<dim3((max_w - 6 + 255) / 256, max_h - 6, ORB_LEVELS), 256, 0, 0>>>(d_pyr, d_score, L);
            B2G_CHECK_LAUNCH(ctx);
        }
        if (max_w > 2 * ORB_EDGE && max_h > 2 * ORB_EDGE) {
            const int max_rows = max_h - 2 * ORB_EDGE;
            orb_rowflags_kernel<<<dim3((max_rows + 3) / 4, ORB_LEVELS), 128, 0, 0>>>(d_score, L, R, d_masks, d_rows);
            B2G_CHECK_LAUNCH(ctx);
            orb_rowscan_kernel<<<ORB_LEVELS, 1024, 0, 0>>>(L, R, d_rows, d_counts);
            B2G_CHECK_LAUNCH(ctx);
            orb_rowwrite_kernel<<<dim3((max_rows + 3) / 4, ORB_LEVELS), 128, 0, 0>>>(d_score, L, R, d_masks, d_rows, d_cand);
            B2G_CHECK_LAUNCH(ctx);
            orb_harris_kernel<<<dim3((L.lv[0].cand_cap + 255) / 256, ORB_LEVELS), 256, 0, 0>>>(d_pyr, L, d_counts, d_cand);
            B2G_CHECK_LAUNCH(ctx);
        }
        unsigned counts[ORB_LEVELS];
        B2G_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof counts, cudaMemcpyDeviceToHost, 0));
        orb_blur_kernel<<<dim3((max_w + 255) / 256, max_h, ORB_LEVELS), 256, 0, 0>>>(d_pyr, d_blur, L);      // runs under the host's selection below
        B2G_CHECK_LAUNCH(ctx);
        B2G_CUDA(cudaStreamSynchronize(0));
        if (const char *dump = getenv("B200GEO_ORB_DUMP")) {           // debugging aid: pyramid, scores and level table to a file
            std::vector<uint8_t> hp((size_t)bytes), hs((size_t)bytes);
            B2G_CUDA(cudaMemcpy(hp.data(), d_pyr, (size_t)bytes, cudaMemcpyDeviceToHost));
            B2G_CUDA(cudaMemcpy(hs.data(), d_score, (size_t)bytes, cudaMemcpyDeviceToHost));
            if (FILE *f = fopen(dump, "wb")) {
                fwrite(&bytes, sizeof bytes, 1, f);
                fwrite(&L, sizeof L, 1, f);
                fwrite(hp.data(), 1, hp.size(), f);
                fwrite(hs.data(), 1, hs.size(), f);
                fclose(f);
            }
        }
        // ---- selection on the host, level by level, exactly as computeKeyPoints orders it ----
        std::vector<OrbCandidate> recs;
        std::vector<OrbKp> level_kps[ORB_LEVELS];
        for (int lv = 0; lv < ORB_LEVELS; ++lv) {
            const OrbLevel &o = L.lv[lv];
            const unsigned n = std::min(counts[lv], o.cand_cap);
            if (!n) continue;
            recs.resize(n);
            B2G_CUDA(cudaMemcpy(recs.data(), d_cand + o.cand_ofs, sizeof(OrbCandidate) * n, cudaMemcpyDeviceToHost));
            // (the records arrive in FAST's row-major order)
            std::vector<OrbKp> &kps = level_kps[lv];
            kps.resize(n);
            for (unsigned i = 0; i < n; ++i) kps[i] = OrbKp{(int)(recs[i].pos & 0xFFFFu), (int)(recs[i].pos >> 16), (float)recs[i].score, recs[i].harris};
            orb_retain_best(kps, 2 * per_level[lv]);             // by FAST score (HARRIS_SCORE keeps twice the level's share first)
            for (OrbKp &k : kps) k.response = k.harris;
            orb_retain_best(kps, per_level[lv]);                 // by Harris response
        }
        std::vector<OrbPoint> pts;
        for (int lv = 0; lv < ORB_LEVELS; ++lv)
            for (const OrbKp &k : level_kps[lv]) pts.push_back(OrbPoint{lv, k.x, k.y, 0.f, 0.f});
        const int n = (int)pts.size();
        *n_out = n;
        if (n == 0) return B200GEO_OK;
        if (n > capacity) return fail(B200GEO_ERR_ARG, "orb: more keypoints than the output buffers hold%s%s");
        OrbPoint *d_pts;
        float *d_angle;
        uint8_t *d_desc;
        if ((rc = scratch(ctx, S_ORB_PTS, (size_t)n * (sizeof(OrbPoint) + sizeof(float) + 32), (void **)&d_pts))) return rc;
        d_angle = reinterpret_cast<float *>(d_pts + n);
        d_desc = reinterpret_cast<uint8_t *>(d_angle + n);
        B2G_CUDA(cudaMemcpyAsync(d_pts, pts.data(), sizeof(OrbPoint) * n, cudaMemcpyHostToDevice, 0));
        orb_angle_kernel<<<(n * 32 + 127) / 128, 128, 0, 0>>>(d_pyr, L, d_pts, n, um, d_angle);
        B2G_CHECK_LAUNCH(ctx);
        std::vector<float> angle(n);
        B2G_CUDA(cudaMemcpyAsync(angle.data(), d_angle, sizeof(float) * n, cudaMemcpyDeviceToHost, 0));
        B2G_CUDA(cudaStreamSynchronize(0));
        {
            int i = 0;
            for (int lv = 0; lv < ORB_LEVELS; ++lv)
                for (const OrbKp &k : level_kps[lv]) {
                    const float sf = scales[lv];
                    float *o = keypoints + 6 * (size_t)i;
                    o[0] = (float)k.x * sf; o[1] = (float)k.y * sf;          // pt *= scale
                    o[2] = ORB_PATCH * sf;                                   // size
                    o[3] = angle[i];
                    o[4] = k.response;                                       // Harris response
                    o[5] = (float)lv;                                        // octave
                    // computeOrbDescriptors: position back on the level, cos / sin of the angle in float (libm's cosf / sinf)
                    const float inv = 1.f / sf;
                    float ang = angle[i];
                    ang *= (float)(CV_PI_F64 / 180.f);
                    pts[i].x = (int)lrintf(o[0] * inv);
                    pts[i].y = (int)lrintf(o[1] * inv);
                    pts[i].a = cosf(ang);
                    pts[i].b = sinf(ang);
                    ++i;
                }
        }
        B2G_CUDA(cudaMemcpyAsync(d_pts, pts.data(), sizeof(OrbPoint) * n, cudaMemcpyHostToDevice, 0));
        orb_describe_kernel<<<(n * 32 + 127) / 128, 128, 0, 0>>>(d_blur, L, d_pts, n, d_desc);
        B2G_CHECK_LAUNCH(ctx);
        B2G_CUDA(cudaMemcpyAsync(descriptors, d_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, 0));
        B2G_CUDA(cudaStreamSynchronize(0));
        return B200GEO_OK;
    
    } catch (const std::exception &e) {
        return fail(B200GEO_ERR_ARG, "orb: host allocation failed: %s%s", e.what());     // no C++ exception crosses the C ABI
    }
}

int b200geo_peer_export(b200geo_ctx *ctx, int world, void *handle_out)
{
    B2G_ENTER(ctx);
    B2G_ARG(world >= 1 && world <= PEER_MAX && handle_out);
    if (ctx->peer_box) return fail(B200GEO_ERR_ARG, "peer exchange already set up on this context (b200geo_peer_close first)%s%s");
    const size_t bytes = (size_t)2 * world * 64;
    B2G_CUDA(cudaMalloc(&ctx->peer_box, bytes));
    B2G_CUDA(cudaMemset(ctx->peer_box, 0, bytes));
    B2G_CUDA(cudaDeviceSynchronize());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ctx->peer_box);
    if (e != cudaSuccess) {
        cudaFree(ctx->peer_box);
        ctx->peer_box = nullptr;
        return fail(B200GEO_ERR_CUDA, "cudaIpcGetMemHandle: %s%s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    ctx->peer_world = world;
    return B200GEO_OK;
}

int b200geo_peer_open(b200geo_ctx *ctx, int rank, int world, const void *handles)
{
    B2G_ENTER(ctx);
    B2G_ARG(ctx->peer_box && world == ctx->peer_world && rank >= 0 && rank < world && handles);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { ctx->peer_map[r] = ctx->peer_box; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + 64 * (size_t)r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ctx->peer_map[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q)
                if (q != rank && ctx->peer_map[q]) { cudaIpcCloseMemHandle(ctx->peer_map[q]); ctx->peer_map[q] = nullptr; }
            return fail(B200GEO_ERR_CUDA, "cudaIpcOpenMemHandle: %s%s", cudaGetErrorString(e));
        }
    }
    ctx->peer_rank = rank;
    ctx->peer_epoch = 0ull;
    return B200GEO_OK;
}

int b200geo_peer_close(b200geo_ctx *ctx)
{
    B2G_ENTER(ctx);
    if (!ctx->peer_box) return B200GEO_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < ctx->peer_world; ++r)
        if (ctx->peer_map[r] && ctx->peer_map[r] != ctx->peer_box) cudaIpcCloseMemHandle(ctx->peer_map[r]);
    memset(ctx->peer_map, 0, sizeof ctx->peer_map);
    cudaFree(ctx->peer_box);
    ctx->peer_box = nullptr;
    ctx->peer_world = 0;
    return B200GEO_OK;
}

int b200geo_ransac_exchange_peer_dev(b200geo_ctx *ctx, const void *rec, void *recs, void *stream)
{
    B2G_ENTER(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    stream_order(ctx, st);
    B2G_ARG(rec && recs);
    if (!ctx->peer_box || !ctx->peer_map[ctx->peer_rank]) return fail(B200GEO_ERR_ARG, "peer exchange is not set up (b200geo_peer_export / _open)%s%s");
    PeerBoxes P;
    memset(&P, 0, sizeof P);
    for (int r = 0; r < ctx->peer_world; ++r) P.box[r] = static_cast<unsigned long long *>(ctx->peer_map[r]);
    const unsigned long long epoch = ++ctx->peer_epoch;
    peer_publish_kernel<<<1, PEER_MAX, 0, st>>>(static_cast<const WinRec *>(rec), P, ctx->peer_world, ctx->peer_rank, epoch);
    B2G_CHECK_LAUNCH(ctx);
    peer_gather_kernel<<<1, PEER_MAX, 0, st>>>(static_cast<unsigned long long *>(ctx->peer_box), ctx->peer_world, epoch, static_cast<WinRec *>(recs),
                                               4000000ull);           // a few seconds of polling (64 ns sleep + the load per iteration)
    B2G_CHECK_LAUNCH(ctx);
    return B200GEO_OK;
}

}  // extern "C"
