// score.cuh -- K2: score hypotheses against the correspondence set.
//
// Replaces HomoModel.computeLoss + threshold + count (ransac.py:55-98, 182-184) for K hypotheses.
// Layout: one thread owns one hypothesis (its 3x3 in registers); the CTA streams the
// correspondence set through shared memory in tiles of 16-byte rows (x, y, x', y') which every
// thread reads as a broadcast LDS.128 -- each correspondence is fetched from L2/HBM once per CTA and
// reused by all of the CTA's hypotheses.  Counts accumulate in a register; no cross-lane traffic.
//
// exact_err() is the reference's float32 chain, operation for operation:
//   r_i = fma(h_i2, 1, fma(h_i1, y, h_i0 * x))      (OpenBLAS sgemm order, k ascending)
//   w = r_2 + 1e-10f ; p = r_0 / w, r_1 / w         (IEEE division)
//   e = sqrt(dx*dx + dy*dy)                          (two rounded products, rounded sum, IEEE sqrt)
// written with the _rn intrinsics so the compiler can neither contract nor reorder it.
#pragma once

#include "common.cuh"

namespace b200geo {

__device__ __forceinline__ float row3(float h0, float h1, float h2, float x, float y)
{
    return __fadd_rn(__fmaf_rn(h1, y, __fmul_rn(h0, x)), h2);   // fma(h2, 1, t) == fl(t + h2)
}

__device__ __forceinline__ float exact_err(const float (&h)[9], float x, float y, float X, float Y)
{
    float r0 = row3(h[0], h[1], h[2], x, y);
    float r1 = row3(h[3], h[4], h[5], x, y);
    float r2 = row3(h[6], h[7], h[8], x, y);
    float w = __fadd_rn(r2, 1e-10f);
    float dx = __fsub_rn(__fdiv_rn(r0, w), X);
    float dy = __fsub_rn(__fdiv_rn(r1, w), Y);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

template <int METHOD>
__device__ __forceinline__ float exact_loss(const float (&h)[9], const float (&hi)[9], float4 p)
{
    if (METHOD == B200GEO_FWD) return exact_err(h, p.x, p.y, p.z, p.w);
    if (METHOD == B200GEO_BACKWARD) return exact_err(hi, p.z, p.w, p.x, p.y);
    return __fadd_rn(exact_err(h, p.x, p.y, p.z, p.w), exact_err(hi, p.z, p.w, p.x, p.y));
}

constexpr int SCORE_THREADS = 256;
constexpr int SCORE_TILE = 2048;   // correspondences per shared-memory tile (32 KB)

// Grid: (ceil(nk / SCORE_THREADS), point_splits, B).  With point_splits > 1 the counts are
// accumulated with atomicAdd and must be zeroed beforehand.
template <int METHOD>
__global__ void __launch_bounds__(SCORE_THREADS)
score_exact_kernel(const float4 *__restrict__ pts, int M, const float *__restrict__ H,
                   const float *__restrict__ Hinv, const uint8_t *__restrict__ flags, int k_begin, int k_end,
                   int K_stride, float th, int pts_per_split, int32_t *__restrict__ counts)
{
    __shared__ float4 tile[SCORE_TILE];
    const int b = blockIdx.z;
    const int k = k_begin + blockIdx.x * SCORE_THREADS + threadIdx.x;
    const bool live = k < k_end;
    pts += (size_t)b * M;
    const size_t o = (size_t)b * K_stride + (live ? k : k_begin);

    float h[9], hi[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        h[j] = __ldg(H + o * 9 + j);
        hi[j] = (METHOD != B200GEO_FWD) ? __ldg(Hinv + o * 9 + j) : 0.f;
    }
    const bool dead = flags ? (flags[o] != 0) : false;

    const int j_begin = blockIdx.y * pts_per_split;
    const int j_end = min(M, j_begin + pts_per_split);
    int cnt = 0;
    for (int t0 = j_begin; t0 < j_end; t0 += SCORE_TILE) {
        const int n = min(SCORE_TILE, j_end - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SCORE_THREADS) tile[i] = __ldg(pts + t0 + i);
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            float e = exact_loss<METHOD>(h, hi, tile[i]);
            cnt += (e < th) ? 1 : 0;
        }
    }
    if (live) {
        if (dead) cnt = 0;
        if (gridDim.y > 1) atomicAdd(counts + o, cnt);
        else counts[o] = cnt;
    }
}

// ------------------------------------------------------------------------------------------------
// Fast scoring kernel (fwd / backward loss).
//
// The reference's chain costs ~60 SASS instructions per (hypothesis, correspondence) because of
// the two IEEE divisions and the IEEE square root.  Whether  e < th  holds is, for all but a thin
// shell of correspondences around the threshold circle, decided by a division-free test:
//     e < th   <=>   a^2 + b^2 - (th*w)^2 < 0 ,   a = r0 - x'*w ,  b = r1 - y'*w
// evaluated as 11 FMA-pipe operations:
//     w' = T*w           = fma(T*h6, x, fma(T*h7, y, T*h8))
//     a  = r0 - x'*w     = fma(h0, x, fma(h1, y, fma(-x'/T, w', h2)))        (b likewise)
//     q  = fma(a, a, fma(b, b, -(w'*w')))
// All rounding errors of this form AND of the reference's chain are absolute in the "times w"
// domain, so one per-hypothesis constant `band` (derived in DESIGN.md, from the magnitudes of the
// rows of H and the largest coordinate) bounds them: |q| > band  =>  sign(q) IS the reference's
// decision.  Correspondences with |q| <= band (a ~1e-2 px shell) are re-evaluated with the exact
// chain, so the counts are bit-identical to the reference's for the same H.
//
// Blackwell specifics: the 11 operations are issued as packed FFMA2/FMUL2 (fma.rn.f32x2), one
// instruction for TWO hypotheses, which halves the issue slots and leaves the kernel bound by the
// FMA pipe itself; the sign bit of q is accumulated with one LEA.HI, min |q| with one FMNMX.
// Correspondences are staged in shared memory already duplicated as (x,x) (y,y) (-x'/T,-x'/T)
// (-y'/T,-y'/T) so that every operand of an FFMA2 is a ready 64-bit register pair.
// ------------------------------------------------------------------------------------------------
#ifndef B2G_SCORE_PAIRED
#define B2G_SCORE_PAIRED 1
#endif
#ifndef B2G_SCORE_PACKCNT
#define B2G_SCORE_PACKCNT 1
#endif
constexpr int FAST_THREADS = 256;
constexpr int FAST_T = 4;          // hypotheses per thread (two FFMA2 lanes-pairs)
constexpr int FAST_TILE = 1024;    // correspondences per tile: 32 bytes each -> 32 KB
constexpr int FAST_CHUNK = 8;      // correspondences between two ambiguity checks

__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

// (q.x < 0 ? 0x0000FFFF : 0) | (q.y < 0 ? 0xFFFF0000 : 0): one PRMT (selector nibbles with the sign-replication bit)
__device__ __forceinline__ unsigned sign_mask2(float2 q)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(r) : "r"(__float_as_uint(q.x)), "r"(__float_as_uint(q.y)));
    return r;
}
// A sum of n such masks is (Nx - Ny) * 65536 - Nx modulo 2^32 (Nx, Ny = how many q.x, q.y were negative; both < 65536).
__device__ __forceinline__ void sign_mask2_decode(unsigned acc, unsigned &nx, unsigned &ny)
{
    nx = (0u - acc) & 0xFFFFu;
    ny = (nx - ((acc + nx) >> 16)) & 0xFFFFu;
}

// the division-free quantity for one hypothesis pair and one correspondence
__device__ __forceinline__ float2 fast_q2(const float2 (&hp)[9], float4 pa, float4 pb)
{
    const float2 xx = make_float2(pa.x, pa.y), yy = make_float2(pa.z, pa.w);
    const float2 nX = make_float2(pb.x, pb.y), nY = make_float2(pb.z, pb.w);
    const float2 w = __ffma2_rn(hp[6], xx, __ffma2_rn(hp[7], yy, hp[8]));
    const float2 a = __ffma2_rn(hp[0], xx, __ffma2_rn(hp[1], yy, __ffma2_rn(nX, w, hp[2])));
    const float2 b = __ffma2_rn(hp[3], xx, __ffma2_rn(hp[4], yy, __ffma2_rn(nY, w, hp[5])));
    const float2 t = __fmul2_rn(w, w);
    return __ffma2_rn(a, a, __ffma2_rn(b, b, neg2(t)));
}

// scalar twin of fast_q2 (bit-identical per lane: both are IEEE fma / mul, round to nearest)
__device__ __forceinline__ float fast_q1(const float (&h)[9], float x, float y, float nX, float nY)
{
    const float w = __fmaf_rn(h[6], x, __fmaf_rn(h[7], y, h[8]));
    const float a = __fmaf_rn(h[0], x, __fmaf_rn(h[1], y, __fmaf_rn(nX, w, h[2])));
    const float b = __fmaf_rn(h[3], x, __fmaf_rn(h[4], y, __fmaf_rn(nY, w, h[5])));
    const float t = __fmul_rn(w, w);
    return __fmaf_rn(a, a, __fmaf_rn(b, b, -t));
}

// Per-hypothesis half-width of the ambiguous shell in q (derivation: DESIGN.md "scoring: error
// bound").  E bounds, in the "times |w|" domain, the sum of every rounding error of the
// reference's chain and of fast_q for a correspondence whose true error is within 2*th:
//   reference: r_i (3 roundings each), w (+1e-10), two divisions, two subtractions, the sqrt chain
//   fast path: T*h6.. coefficient roundings, x'/T scaling, the 3+3+2 fused operations
// with |x|,|y|,|x'|,|y'| <= C and |w| <= S2.  Then  | |q| | > band  =>  sign(q) is the
// reference's decision, because  band >= E*(2*T*S2 + E) + (rounding of q itself).
__device__ __forceinline__ float fast_band(const float (&h)[9], float C, float T)
{
    const float u = 5.9604645e-8f;   // 2^-24
    const float S2 = (fabsf(h[6]) + fabsf(h[7])) * C + fabsf(h[8]);
    const float P = C + 2.f * T;
    const float E = u * (2.f * (fabsf(h[0]) + fabsf(h[3])) * C + 2.f * (fabsf(h[2]) + fabsf(h[5])) +
                         P * (12.f * S2 + 2.f * fabsf(h[6]) * C + 2.f * fabsf(h[8])) + 12.f * C * S2 + 18.f * T * S2) +
                    2.f * P * 1e-10f;
    return 1.25f * (E * (2.0001f * T * S2 + E) + 3.f * u * T * T * S2 * S2);
}

// One deferred unit of exact work: hypothesis slot (b*K_stride + k) and the first correspondence of
// the FAST_CHUNK-sized chunk in which some |q| fell inside the shell.
struct FastDeferred { unsigned slot, j0; };

// largest |coordinate| of each correspondence set (NaN -> inf); grid (ceil(M / 2048), B), block 256;
// cmax[] must be zero beforehand (non-negative floats order like their bit patterns: atomicMax on int)
__global__ void __launch_bounds__(256) cmax_kernel(const float4 *__restrict__ pts, int M, float *__restrict__ cmax)
{
    pts += (size_t)blockIdx.y * M;
    float m = 0.f;
    const int end = min(M, (int)(blockIdx.x + 1) * 2048);
    for (int i = blockIdx.x * 2048 + threadIdx.x; i < end; i += 256) {
        float4 p = __ldg(pts + i);
        float v = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fmaxf(fabsf(p.z), fabsf(p.w)));
        if (p.x != p.x || p.y != p.y || p.z != p.z || p.w != p.w) v = __int_as_float(0x7f800000);
        m = fmaxf(m, v);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(cmax) + blockIdx.y, __float_as_int(m));
}

// Grid: (ceil(nk / (FAST_THREADS*FAST_T)), point_splits, B).  METHOD is FWD or BACKWARD; for
// BACKWARD `H` must point at the inverse matrices and the tile is filled with the roles of the
// two point sets exchanged.  T = (float)th > 0 and finite.
template <int METHOD>
__global__ void __launch_bounds__(FAST_THREADS, 2)
score_fast_kernel(const float4 *__restrict__ pts, int M, const float *__restrict__ H, const uint8_t *__restrict__ flags,
                  int k_begin, int k_end, int K_stride, float T, const float *__restrict__ cmax, int pts_per_split,
                  int32_t *__restrict__ counts, unsigned long long *__restrict__ stat_ambiguous,
                  FastDeferred *__restrict__ wl, unsigned *__restrict__ wl_count, unsigned wl_cap)
{
    __shared__ float4 tile[2 * FAST_TILE];
    const int b = blockIdx.z;
    // deferred exact work goes to THIS CTA's segment of the list (wl_cap entries, counted in shared memory): one global
    // counter for all CTAs was an atomic per deferred chunk on a single address -- about one per nanosecond on the batched
    // configuration, i.e. the L2's serialisation rate (score_fast_kernel ran at 79 % of its single-pair rate there)
    __shared__ unsigned s_wl;
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    FastDeferred *const wl_seg = wl + (size_t)cta * wl_cap;
    if (threadIdx.x == 0) s_wl = 0u;
    __syncthreads();
    pts += (size_t)b * M;
    const float C = cmax[b];
    const float invT = 1.0f / T;       // only to scale x', y' in the tile; its rounding is inside `band`

    // this thread's FAST_T hypotheses: k = base + t*FAST_THREADS + tid  (t = 0..3), paired (0,1) (2,3)
    const int base = k_begin + blockIdx.x * (FAST_THREADS * FAST_T) + threadIdx.x;
    float hs[FAST_T][9];
    float band[FAST_T];
    bool live[FAST_T], dead[FAST_T];
#pragma unroll
    for (int t = 0; t < FAST_T; ++t) {
        const int k = base + t * FAST_THREADS;
        live[t] = k < k_end;
        const size_t o = (size_t)b * K_stride + (live[t] ? k : k_begin);
        bool bad = flags ? (flags[o] != 0) : false;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            hs[t][j] = __ldg(H + o * 9 + j);
            bad = bad || !(fabsf(hs[t][j]) <= 3.0e38f);                 // NaN / inf
        }
        dead[t] = bad || !live[t];
        if (dead[t]) {
            // scored as the identity (cheap, never ambiguous in a way that matters) and discarded
#pragma unroll
            for (int j = 0; j < 9; ++j) hs[t][j] = (j == 0 || j == 4 || j == 8) ? 1.f : 0.f;
        }
        band[t] = fast_band(hs[t], C, T);
        hs[t][6] = __fmul_rn(T, hs[t][6]);
        hs[t][7] = __fmul_rn(T, hs[t][7]);
        hs[t][8] = __fmul_rn(T, hs[t][8]);
    }
    float2 hp[FAST_T / 2][9];
#pragma unroll
    for (int g = 0; g < FAST_T / 2; ++g)
#pragma unroll
        for (int j = 0; j < 9; ++j) hp[g][j] = make_float2(hs[2 * g][j], hs[2 * g + 1][j]);

    unsigned cnt[FAST_T];
    float mq[FAST_T];
#pragma unroll
    for (int t = 0; t < FAST_T; ++t) { cnt[t] = 0u; mq[t] = 3.0e38f; }
    unsigned n_amb = 0;
#if B2G_SCORE_PAIRED && B2G_SCORE_PACKCNT
    unsigned pk[FAST_T / 2];
#pragma unroll
    for (int g = 0; g < FAST_T / 2; ++g) pk[g] = 0u;
    static_assert(FAST_TILE < 65536, "packed sign counts are decoded once per tile");
#endif

    const int j_begin = blockIdx.y * pts_per_split;
    const int j_end = min(M, j_begin + pts_per_split);
    for (int t0 = j_begin; t0 < j_end; t0 += FAST_TILE) {
        const int n = min(FAST_TILE, j_end - t0);
        const int n_pad = (n + FAST_CHUNK - 1) / FAST_CHUNK * FAST_CHUNK;
        __syncthreads();
        for (int i = threadIdx.x; i < n_pad; i += FAST_THREADS) {
            float4 p;
            if (i < n) {
                p = __ldg(pts + t0 + i);
                if (METHOD == B200GEO_BACKWARD) p = make_float4(p.z, p.w, p.x, p.y);
            } else {
                // padding: a correspondence no hypothesis can count (q = +huge), never ambiguous
                p = make_float4(0.f, 0.f, 1.0e18f, 1.0e18f);
            }
            const float nX = -__fmul_rn(p.z, invT), nY = -__fmul_rn(p.w, invT);
            tile[2 * i] = make_float4(p.x, p.x, p.y, p.y);
            tile[2 * i + 1] = make_float4(nX, nX, nY, nY);
        }
        __syncthreads();
        for (int i = 0; i < n_pad; i += FAST_CHUNK) {
#if B2G_SCORE_PAIRED
            // two correspondences per step: the running minimum takes both |q| in one 3-input FMNMX3 (sm_100), i.e. half an
            // ALU-pipe instruction per pair less (2159 vs 2107 Gpairs/s on C3)
#pragma unroll
            for (int u = 0; u < FAST_CHUNK; u += 2) {
                const float4 pa0 = tile[2 * (i + u)], pb0 = tile[2 * (i + u) + 1];
                const float4 pa1 = tile[2 * (i + u) + 2], pb1 = tile[2 * (i + u) + 3];
#pragma unroll
                for (int g = 0; g < FAST_T / 2; ++g) {
                    const float2 q0 = fast_q2(hp[g], pa0, pb0), q1 = fast_q2(hp[g], pa1, pb1);
#if B2G_SCORE_PACKCNT
                    // sign masks of the two hypotheses of a pair in the halves of one word (PRMT with sign replication:
                    // 0xFFFF where q < 0), two correspondences added at once (IADD3): 3 ALU instructions per 4 pairs instead of 4
                    pk[g] += sign_mask2(q0) + sign_mask2(q1);
#else
                    cnt[2 * g] += __float_as_uint(q0.x) >> 31;
                    cnt[2 * g + 1] += __float_as_uint(q0.y) >> 31;
                    cnt[2 * g] += __float_as_uint(q1.x) >> 31;
                    cnt[2 * g + 1] += __float_as_uint(q1.y) >> 31;
#endif
                    mq[2 * g] = fminf(fminf(mq[2 * g], fabsf(q0.x)), fabsf(q1.x));
                    mq[2 * g + 1] = fminf(fminf(mq[2 * g + 1], fabsf(q0.y)), fabsf(q1.y));
                }
            }
#else
#pragma unroll
            for (int u = 0; u < FAST_CHUNK; ++u) {
                const float4 pa = tile[2 * (i + u)], pb = tile[2 * (i + u) + 1];
#pragma unroll
                for (int g = 0; g < FAST_T / 2; ++g) {
                    const float2 q = fast_q2(hp[g], pa, pb);
                    cnt[2 * g] += __float_as_uint(q.x) >> 31;
                    cnt[2 * g + 1] += __float_as_uint(q.y) >> 31;
                    mq[2 * g] = fminf(mq[2 * g], fabsf(q.x));
                    mq[2 * g + 1] = fminf(mq[2 * g + 1], fabsf(q.y));
                }
            }
#endif
            bool amb = false;
#pragma unroll
            for (int t = 0; t < FAST_T; ++t) amb = amb || !(mq[t] > band[t]);
            if (amb) {
                // Some |q| of this chunk is inside the shell (or NaN).  The exact re-evaluation is
                // DEFERRED to score_fixup_kernel through a global work list, so that one lane's
                // rare event does not stall the other 31; only if the list is full is the chunk
                // redone here.
#pragma unroll
                for (int t = 0; t < FAST_T; ++t) {
                    if (mq[t] > band[t]) continue;
                    mq[t] = 3.0e38f;
                    if (dead[t]) continue;
                    const size_t o = (size_t)b * K_stride + base + t * FAST_THREADS;
                    const unsigned w = atomicAdd(&s_wl, 1u);
                    if (w < wl_cap) {
                        wl_seg[w].slot = (unsigned)o;
                        wl_seg[w].j0 = (unsigned)(t0 + i);
                        continue;
                    }
                    float hx[9], hq[9];
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        hx[j] = __ldg(H + o * 9 + j);
                        hq[j] = (t & 1) ? hp[t / 2][j].y : hp[t / 2][j].x;
                    }
                    for (int u = 0; u < FAST_CHUNK; ++u) {
                        const int jj = t0 + i + u;
                        if (jj >= j_end) break;
                        const float4 pa = tile[2 * (i + u)], pb = tile[2 * (i + u) + 1];
                        const float q = fast_q1(hq, pa.x, pa.z, pb.x, pb.z);
                        if (fabsf(q) > band[t]) continue;
                        float4 p = __ldg(pts + jj);
                        if (METHOD == B200GEO_BACKWARD) p = make_float4(p.z, p.w, p.x, p.y);
                        const int exact_in = exact_err(hx, p.x, p.y, p.z, p.w) < T ? 1 : 0;
                        cnt[t] += (unsigned)(exact_in - (int)(__float_as_uint(q) >> 31));
                        ++n_amb;
                    }
                }
            }
        }
#if B2G_SCORE_PAIRED && B2G_SCORE_PACKCNT
#pragma unroll
        for (int g = 0; g < FAST_T / 2; ++g) {          // this tile's packed sign counts (at most FAST_TILE each) -> the counters
            unsigned nx, ny;
            sign_mask2_decode(pk[g], nx, ny);
            cnt[2 * g] += nx;
            cnt[2 * g + 1] += ny;
            pk[g] = 0u;
        }
#endif
    }
#pragma unroll
    for (int t = 0; t < FAST_T; ++t) {
        if (!live[t]) continue;
        const size_t o = (size_t)b * K_stride + base + t * FAST_THREADS;
        const int c = dead[t] ? 0 : (int)cnt[t];
        if (gridDim.y > 1) { if (c) atomicAdd(counts + o, c); }
        else counts[o] = c;
    }
    if (stat_ambiguous && n_amb) atomicAdd(stat_ambiguous, (unsigned long long)n_amb);
    __syncthreads();
    if (threadIdx.x == 0) wl_count[cta] = min(s_wl, wl_cap);
}

// Exact re-evaluation of the deferred chunks: one thread per deferred chunk (its hypothesis, band and scaled rows are set up
// once and its FAST_CHUNK correspondences walked in a loop -- one thread per (chunk, correspondence) repeated that set-up eight
// times and was 7 % of the batched step).  Recomputes q with the scalar twin of fast_q2 (bit-identical), and for |q| <= band
// replaces the fast decision by the reference-order chain's:  counts[slot] += exact - fast.
template <int METHOD>
__global__ void __launch_bounds__(256)
score_fixup_kernel(const float4 *__restrict__ pts, int M, const float *__restrict__ H, int K_stride, float T,
                   const float *__restrict__ cmax, const FastDeferred *__restrict__ wl, const unsigned *__restrict__ wl_count,
                   unsigned n_seg, unsigned wl_cap, int32_t *__restrict__ counts, unsigned long long *__restrict__ stat_ambiguous)
{
    const float invT = 1.0f / T;
    unsigned local = 0;
    // segment `sgm` of the list holds wl_count[sgm] entries of the CTA that filled it; a block takes whole segments
    for (unsigned sgm = blockIdx.x; sgm < n_seg; sgm += gridDim.x)
    for (unsigned i = threadIdx.x, total = min(wl_count[sgm], wl_cap); i < total; i += blockDim.x) {
        const FastDeferred e = wl[(size_t)sgm * wl_cap + i];
        const unsigned b = e.slot / (unsigned)K_stride;
        float hx[9], hq[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) hx[c] = hq[c] = __ldg(H + (size_t)e.slot * 9 + c);
        const float band = fast_band(hx, cmax[b], T);
        hq[6] = __fmul_rn(T, hq[6]);
        hq[7] = __fmul_rn(T, hq[7]);
        hq[8] = __fmul_rn(T, hq[8]);
        const float4 *pp = pts + (size_t)b * M + e.j0;
        const int n = min(FAST_CHUNK, M - (int)e.j0);
        int d = 0;
        for (int u = 0; u < n; ++u) {
            float4 p = __ldg(pp + u);
            if (METHOD == B200GEO_BACKWARD) p = make_float4(p.z, p.w, p.x, p.y);
            const float q = fast_q1(hq, p.x, p.y, -__fmul_rn(p.z, invT), -__fmul_rn(p.w, invT));
            if (fabsf(q) > band) continue;
            const int exact_in = exact_err(hx, p.x, p.y, p.z, p.w) < T ? 1 : 0;
            d += exact_in - (int)(__float_as_uint(q) >> 31);
            ++local;
        }
        if (d) atomicAdd(counts + e.slot, d);
    }
    // one atomic per block: 300k threads adding to one word serialise at the L2 (measured: 188 us for a 125k-hypothesis shard,
    // more than the work of the kernel)
    local = __reduce_add_sync(0xffffffffu, local);
    __shared__ unsigned s_local;
    if (threadIdx.x == 0) s_local = 0u;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_local, local);
    __syncthreads();
    if (threadIdx.x == 0 && stat_ambiguous && s_local) atomicAdd(stat_ambiguous, (unsigned long long)s_local);
}

// ------------------------------------------------------------------------------------------------
// Fast scoring kernel for the reproj loss (RANSAC.run's default method, ransac.py:91-95):
//     loss = e_fwd + e_bwd ,  e_fwd = |H x - x'| ,  e_bwd = |H^-1 x' - x|     (each a float32 sqrt)
// The sum of two square roots has no polynomial form, but both sides of the threshold are
// bracketed by two division-free tests per direction, evaluated from the same a, b, w:
//     accept:  e_fwd < T/2  and  e_bwd < T/2      =>  fl(e_fwd + e_bwd) <= 2 pred(T/2) = pred(T) < T
//     reject:  e_fwd >= T   or   e_bwd >= T       =>  fl(e_fwd + e_bwd) >= T   (the other term is >= 0)
// With the rows scaled by Th = T/2:  q = a^2 + b^2 - (Th w)^2 < 0  <=>  e < T/2 ,  and
// r = q - 3 (Th w)^2 = a^2 + b^2 - (T w)^2 >= 0  <=>  e >= T.  12 FMA-pipe operations per direction,
// packed two hypotheses per instruction (FFMA2).  Inliers of a good hypothesis (noise << T/2) are
// accepted, outliers rejected; what falls between -- or within the rounding band of either test --
// goes to the same deferred exact work list as in score_fast_kernel, so the counts are those of
// the reference-order chain.
// ------------------------------------------------------------------------------------------------
#ifndef B2G_FAST2_G
#define B2G_FAST2_G 1        // measured on C3 (reproj, Gpairs/s): 1 pair x 256 threads x 2 CTAs 453; 2 pairs: 348 (256 x 1), 368 (256 x 2, spills), 424 (192 x 2), 449 (128 x 3)
#endif
#ifndef B2G_FAST2_THREADS
#define B2G_FAST2_THREADS 256
#endif
#ifndef B2G_FAST2_CTAS
#define B2G_FAST2_CTAS 2
#endif
constexpr int FAST2_THREADS = B2G_FAST2_THREADS;
constexpr int FAST2_G = B2G_FAST2_G;          // packed hypothesis pairs per thread
constexpr int FAST2_T = 2 * FAST2_G;          // hypotheses per thread
constexpr int FAST2_CTAS = B2G_FAST2_CTAS;    // CTAs per SM the register budget is planned for
constexpr int FAST2_TILE = 512;    // correspondences per tile: 64 bytes each -> 32 KB

struct Fast2Q { float2 q1, r1, q2, r2; };

__device__ __forceinline__ Fast2Q fast2_q(const float2 (&hf)[9], const float2 (&hb)[9], float4 a0, float4 a1, float4 b0, float4 b1)
{
    Fast2Q o;
    {
        const float2 xx = make_float2(a0.x, a0.y), yy = make_float2(a0.z, a0.w), nX = make_float2(a1.x, a1.y), nY = make_float2(a1.z, a1.w);
        const float2 w = __ffma2_rn(hf[6], xx, __ffma2_rn(hf[7], yy, hf[8]));
        const float2 a = __ffma2_rn(hf[0], xx, __ffma2_rn(hf[1], yy, __ffma2_rn(nX, w, hf[2])));
        const float2 b = __ffma2_rn(hf[3], xx, __ffma2_rn(hf[4], yy, __ffma2_rn(nY, w, hf[5])));
        const float2 t = __fmul2_rn(w, w);
        o.q1 = __ffma2_rn(a, a, __ffma2_rn(b, b, neg2(t)));
        o.r1 = __ffma2_rn(make_float2(-3.f, -3.f), t, o.q1);
    }
    {
        const float2 xx = make_float2(b0.x, b0.y), yy = make_float2(b0.z, b0.w), nX = make_float2(b1.x, b1.y), nY = make_float2(b1.z, b1.w);
        const float2 w = __ffma2_rn(hb[6], xx, __ffma2_rn(hb[7], yy, hb[8]));
        const float2 a = __ffma2_rn(hb[0], xx, __ffma2_rn(hb[1], yy, __ffma2_rn(nX, w, hb[2])));
        const float2 b = __ffma2_rn(hb[3], xx, __ffma2_rn(hb[4], yy, __ffma2_rn(nY, w, hb[5])));
        const float2 t = __fmul2_rn(w, w);
        o.q2 = __ffma2_rn(a, a, __ffma2_rn(b, b, neg2(t)));
        o.r2 = __ffma2_rn(make_float2(-3.f, -3.f), t, o.q2);
    }
    return o;
}

// scalar twin for one hypothesis and one direction (bit-identical per lane): returns q, r
__device__ __forceinline__ void fast2_q1(const float (&h)[9], float x, float y, float nX, float nY, float &q, float &r)
{
    const float w = __fmaf_rn(h[6], x, __fmaf_rn(h[7], y, h[8]));
    const float a = __fmaf_rn(h[0], x, __fmaf_rn(h[1], y, __fmaf_rn(nX, w, h[2])));
    const float b = __fmaf_rn(h[3], x, __fmaf_rn(h[4], y, __fmaf_rn(nY, w, h[5])));
    const float t = __fmul_rn(w, w);
    q = __fmaf_rn(a, a, __fmaf_rn(b, b, -t));
    r = __fmaf_rn(-3.f, t, q);
}

// bands of the two tests of one direction: accept at Th = T/2, reject at T (the latter also covers
// forming r from q and the rounded square)
__device__ __forceinline__ void fast2_bands(const float (&h)[9], float C, float T, float &accept_band, float &reject_band)
{
    const float u = 5.9604645e-8f;
    const float S2 = (fabsf(h[6]) + fabsf(h[7])) * C + fabsf(h[8]);
    accept_band = fast_band(h, C, 0.5f * T);
    reject_band = fast_band(h, C, T) + 8.f * u * T * T * S2 * S2;
}

// 0 = rejected, 1 = accepted, 2 = undecided (exact chain needed); NaN anywhere -> undecided
__device__ __forceinline__ int fast2_decide(float q1, float r1, float q2, float r2, float a1, float rj1, float a2, float rj2)
{
    const bool acc = (q1 < -a1) && (q2 < -a2);
    const bool rej = (r1 > rj1) || (r2 > rj2);
    return acc ? 1 : (rej ? 0 : 2);
}

// Grid: (ceil(nk / (FAST2_T*FAST2_THREADS)), point_splits, B).  T > 0 and finite; H and Hinv both required.
// A thread owns FAST2_T = 2 FAST2_G hypotheses (FAST2_G packed pairs).  One pair per thread is the measured optimum: the kernel
// is bound by fixed-latency dependencies of its chains and by the deferral traffic (ncu: issue slots 56 % busy, top stalls wait
// / math throttle / barrier; 1.8 % of the chunks deferred), not by the four broadcast LDS.128 per correspondence, and two pairs per
// thread cost occupancy (112 -> 168+ registers).
__global__ void __launch_bounds__(FAST2_THREADS, FAST2_CTAS)
score_fast2_kernel(const float4 *__restrict__ pts, int M, const float *__restrict__ H, const float *__restrict__ Hinv,
                   const uint8_t *__restrict__ flags, int k_begin, int k_end, int K_stride, float T, const float *__restrict__ cmax,
                   int pts_per_split, int32_t *__restrict__ counts, FastDeferred *__restrict__ wl, unsigned *__restrict__ wl_count,
                   unsigned wl_cap)
{
    __shared__ float4 tile[4 * FAST2_TILE];
    const int b = blockIdx.z;
    // deferred exact work goes to THIS CTA's segment of the list (wl_cap entries, counted in shared memory): one global
    // counter for all CTAs was an atomic per deferred chunk on a single address -- about one per nanosecond on the batched
    // configuration, i.e. the L2's serialisation rate (score_fast_kernel ran at 79 % of its single-pair rate there)
    __shared__ unsigned s_wl;
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    FastDeferred *const wl_seg = wl + (size_t)cta * wl_cap;
    if (threadIdx.x == 0) s_wl = 0u;
    __syncthreads();
    pts += (size_t)b * M;
    const float C = cmax[b];
    const float Th = 0.5f * T, invTh = 1.0f / Th;

    // this thread's hypotheses: k = base + t * FAST2_THREADS, t = 0 .. FAST2_T - 1, packed (0,1) (2,3) ...
    const int base = k_begin + blockIdx.x * (FAST2_THREADS * FAST2_T) + threadIdx.x;
    float2 hf[FAST2_G][9], hb[FAST2_G][9];
    float ba1[FAST2_T], br1[FAST2_T], ba2[FAST2_T], br2[FAST2_T];
    bool live[FAST2_T], dead[FAST2_T];
#pragma unroll
    for (int t = 0; t < FAST2_T; ++t) {
        const int k = base + t * FAST2_THREADS;
        live[t] = k < k_end;
        const size_t o = (size_t)b * K_stride + (live[t] ? k : k_begin);
        bool bad = flags ? (flags[o] != 0) : false;
        float f[9], g[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            f[j] = __ldg(H + o * 9 + j);
            g[j] = __ldg(Hinv + o * 9 + j);
            bad = bad || !(fabsf(f[j]) <= 3.0e38f) || !(fabsf(g[j]) <= 3.0e38f);
        }
        dead[t] = bad || !live[t];
        if (dead[t]) {
#pragma unroll
            for (int j = 0; j < 9; ++j) f[j] = g[j] = (j == 0 || j == 4 || j == 8) ? 1.f : 0.f;
        }
        fast2_bands(f, C, T, ba1[t], br1[t]);
        fast2_bands(g, C, T, ba2[t], br2[t]);
#pragma unroll
        for (int j = 6; j < 9; ++j) { f[j] = __fmul_rn(Th, f[j]); g[j] = __fmul_rn(Th, g[j]); }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if ((t & 1) == 0) { hf[t / 2][j].x = f[j]; hb[t / 2][j].x = g[j]; }
            else { hf[t / 2][j].y = f[j]; hb[t / 2][j].y = g[j]; }
        }
    }
    unsigned cnt[FAST2_T];
#pragma unroll
    for (int t = 0; t < FAST2_T; ++t) cnt[t] = 0u;
    const int j_begin = blockIdx.y * pts_per_split;
    const int j_end = min(M, j_begin + pts_per_split);
    for (int t0 = j_begin; t0 < j_end; t0 += FAST2_TILE) {
        const int n = min(FAST2_TILE, j_end - t0);
        const int n_pad = (n + FAST_CHUNK - 1) / FAST_CHUNK * FAST_CHUNK;
        __syncthreads();
        for (int i = threadIdx.x; i < n_pad; i += FAST2_THREADS) {
            // padding: a correspondence every hypothesis rejects outright (huge error both ways)
            float4 p = i < n ? __ldg(pts + t0 + i) : make_float4(0.f, 0.f, 1.0e18f, 1.0e18f);
            tile[4 * i + 0] = make_float4(p.x, p.x, p.y, p.y);
            const float nX = -__fmul_rn(p.z, invTh), nY = -__fmul_rn(p.w, invTh);
            tile[4 * i + 1] = make_float4(nX, nX, nY, nY);
            tile[4 * i + 2] = make_float4(p.z, p.z, p.w, p.w);
            const float nx = -__fmul_rn(p.x, invTh), ny = -__fmul_rn(p.y, invTh);
            tile[4 * i + 3] = make_float4(nx, nx, ny, ny);
        }
        __syncthreads();
        for (int i = 0; i < n_pad; i += FAST_CHUNK) {
            bool und[FAST2_T];
#pragma unroll
            for (int t = 0; t < FAST2_T; ++t) und[t] = false;
#pragma unroll
            for (int u = 0; u < FAST_CHUNK; ++u) {
                const float4 p0 = tile[4 * (i + u)], p1 = tile[4 * (i + u) + 1], p2 = tile[4 * (i + u) + 2], p3 = tile[4 * (i + u) + 3];
#pragma unroll
                for (int g = 0; g < FAST2_G; ++g) {
                    const Fast2Q q = fast2_q(hf[g], hb[g], p0, p1, p2, p3);
                    const int d0 = fast2_decide(q.q1.x, q.r1.x, q.q2.x, q.r2.x, ba1[2 * g], br1[2 * g], ba2[2 * g], br2[2 * g]);
                    const int d1 = fast2_decide(q.q1.y, q.r1.y, q.q2.y, q.r2.y, ba1[2 * g + 1], br1[2 * g + 1], ba2[2 * g + 1], br2[2 * g + 1]);
                    cnt[2 * g] += (d0 == 1); und[2 * g] = und[2 * g] || (d0 == 2);
                    cnt[2 * g + 1] += (d1 == 1); und[2 * g + 1] = und[2 * g + 1] || (d1 == 2);
                }
            }
#pragma unroll
            for (int t = 0; t < FAST2_T; ++t) {
                if (!und[t] || dead[t]) continue;
                // some correspondence of this chunk is undecided: defer the chunk to score_fixup2_kernel
                const size_t o = (size_t)b * K_stride + base + t * FAST2_THREADS;
                const unsigned w = atomicAdd(&s_wl, 1u);
                if (w < wl_cap) {
                    wl_seg[w].slot = (unsigned)o;
                    wl_seg[w].j0 = (unsigned)(t0 + i);
                    continue;
                }
                // list full: settle the chunk here
                float hx[9], hix[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) { hx[j] = __ldg(H + o * 9 + j); hix[j] = __ldg(Hinv + o * 9 + j); }
                for (int u = 0; u < FAST_CHUNK; ++u) {
                    const int jj = t0 + i + u;
                    if (jj >= j_end) break;
                    const Fast2Q q = fast2_q(hf[t / 2], hb[t / 2], tile[4 * (i + u)], tile[4 * (i + u) + 1], tile[4 * (i + u) + 2], tile[4 * (i + u) + 3]);
                    const int d = (t & 1) == 0 ? fast2_decide(q.q1.x, q.r1.x, q.q2.x, q.r2.x, ba1[t], br1[t], ba2[t], br2[t])
                                               : fast2_decide(q.q1.y, q.r1.y, q.q2.y, q.r2.y, ba1[t], br1[t], ba2[t], br2[t]);
                    if (d != 2) continue;
                    cnt[t] += exact_loss<B200GEO_REPROJ>(hx, hix, __ldg(pts + jj)) < T ? 1u : 0u;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < FAST2_T; ++t) {
        if (!live[t]) continue;
        const size_t o = (size_t)b * K_stride + base + t * FAST2_THREADS;
        const int c = dead[t] ? 0 : (int)cnt[t];
        if (gridDim.y > 1) { if (c) atomicAdd(counts + o, c); }
        else counts[o] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) wl_count[cta] = min(s_wl, wl_cap);
}

// Exact settlement of the deferred reproj chunks: one thread per deferred chunk (see score_fixup_kernel).
__global__ void __launch_bounds__(256)
score_fixup2_kernel(const float4 *__restrict__ pts, int M, const float *__restrict__ H, const float *__restrict__ Hinv, int K_stride,
                    float T, const float *__restrict__ cmax, const FastDeferred *__restrict__ wl, const unsigned *__restrict__ wl_count,
                    unsigned n_seg, unsigned wl_cap, int32_t *__restrict__ counts, unsigned long long *__restrict__ stat_ambiguous)
{
    const float Th = 0.5f * T, invTh = 1.0f / Th;
    unsigned local = 0;
    for (unsigned sgm = blockIdx.x; sgm < n_seg; sgm += gridDim.x)
    for (unsigned i = threadIdx.x, total = min(wl_count[sgm], wl_cap); i < total; i += blockDim.x) {
        const FastDeferred e = wl[(size_t)sgm * wl_cap + i];
        const unsigned b = e.slot / (unsigned)K_stride;
        float hx[9], hix[9], f[9], g[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) { hx[c] = f[c] = __ldg(H + (size_t)e.slot * 9 + c); hix[c] = g[c] = __ldg(Hinv + (size_t)e.slot * 9 + c); }
        float a1, r1b, a2, r2b;
        fast2_bands(hx, cmax[b], T, a1, r1b);
        fast2_bands(hix, cmax[b], T, a2, r2b);
#pragma unroll
        for (int c = 6; c < 9; ++c) { f[c] = __fmul_rn(Th, f[c]); g[c] = __fmul_rn(Th, g[c]); }
        const float4 *pp = pts + (size_t)b * M + e.j0;
        const int n = min(FAST_CHUNK, M - (int)e.j0);
        int d = 0;
        for (int u = 0; u < n; ++u) {
            const float4 p = __ldg(pp + u);
            float q1, r1, q2, r2;
            fast2_q1(f, p.x, p.y, -__fmul_rn(p.z, invTh), -__fmul_rn(p.w, invTh), q1, r1);
            fast2_q1(g, p.z, p.w, -__fmul_rn(p.x, invTh), -__fmul_rn(p.y, invTh), q2, r2);
            if (fast2_decide(q1, r1, q2, r2, a1, r1b, a2, r2b) != 2) continue;
            if (exact_loss<B200GEO_REPROJ>(hx, hix, p) < T) ++d;
            ++local;
        }
        if (d) atomicAdd(counts + e.slot, d);
    }
    // one atomic per block: 300k threads adding to one word serialise at the L2 (measured: 188 us for a 125k-hypothesis shard,
    // more than the work of the kernel)
    local = __reduce_add_sync(0xffffffffu, local);
    __shared__ unsigned s_local;
    if (threadIdx.x == 0) s_local = 0u;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_local, local);
    __syncthreads();
    if (threadIdx.x == 0 && stat_ambiguous && s_local) atomicAdd(stat_ambiguous, (unsigned long long)s_local);
}

// per-point loss of one hypothesis (HomoModel.computeLoss, ransac.py:84-98)
__global__ void loss_kernel(const float4 *__restrict__ pts, int M, const float *__restrict__ H,
                            const float *__restrict__ Hinv, int method, float *__restrict__ err)
{
    float h[9], hi[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { h[j] = H[j]; hi[j] = Hinv ? Hinv[j] : 0.f; }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        float4 p = __ldg(pts + i);
        float e;
        if (method == B200GEO_FWD) e = exact_loss<B200GEO_FWD>(h, hi, p);
        else if (method == B200GEO_BACKWARD) e = exact_loss<B200GEO_BACKWARD>(h, hi, p);
        else e = exact_loss<B200GEO_REPROJ>(h, hi, p);
        err[i] = e;
    }
}

// HomoModel.fwd / reproj as standalone calls (ransac.py:55-76): P is 3 x M row-major float32
// (third row = 1 when the caller passed 2 rows); out = (Hm @ P) / (row2 + 1e-10), all three rows.
__global__ void project_kernel(const float *__restrict__ P, int M, const float *__restrict__ Hm, float *__restrict__ out)
{
    float h[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) h[j] = Hm[j];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        float x = P[i], y = P[M + i], z = P[2 * M + i];
        float r0 = __fmaf_rn(h[2], z, __fmaf_rn(h[1], y, __fmul_rn(h[0], x)));
        float r1 = __fmaf_rn(h[5], z, __fmaf_rn(h[4], y, __fmul_rn(h[3], x)));
        float r2 = __fmaf_rn(h[8], z, __fmaf_rn(h[7], y, __fmul_rn(h[6], x)));
        float w = __fadd_rn(r2, 1e-10f);
        out[i] = __fdiv_rn(r0, w);
        out[M + i] = __fdiv_rn(r1, w);
        out[2 * M + i] = __fdiv_rn(r2, w);
    }
}

// HomoModel.dist (ransac.py:78-82): sqrt(sum over rows of (a - b)^2), rows added in order.
__global__ void dist_kernel(const float *__restrict__ a, const float *__restrict__ b, int rows, int M, float *__restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) {
            float d = __fsub_rn(a[(size_t)r * M + i], b[(size_t)r * M + i]);
            float q = __fmul_rn(d, d);
            s = (r == 0) ? q : __fadd_rn(s, q);
        }
        out[i] = __fsqrt_rn(s);
    }
}

}  // namespace b200geo
