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
