// select_refit.cuh -- K2b best-hypothesis selection and K3 inlier mask + least-squares refit.
//
// Selection restates the sequential rule of RANSAC.run (ransac.py:169-202) as two reductions:
//   winner = min{ k : count_k >= thr }  if it exists (the reference's early `break`),
//            else the lowest k attaining max count (strict `>` keeps the first).
// Both are encoded so that ONE unsigned 64-bit MAX reduces them, within a warp (shuffles), across
// CTAs (atomicMax) and across GPUs (a two-element NCCL max all-reduce):
//   keys[0] = (count << 32) | (0xFFFFFFFF - k)
//   keys[1] = count >= thr ? 0x100000000 - k : 0
//
// The refit replaces calcHomographyLinear(..., collective=True) (homography.py:48-69, 90-105):
// the minimiser of |A h - b|^2 over the inlier rows.  The reference evaluates inv(A^T A) A^T b in
// float32 (cond(A^T A) ~ 1e15 on its own fixture); here the same least-squares problem is solved
// exactly: the 9x9 Gram matrix of the homogeneous rows is accumulated in float64 in centred,
// scaled coordinates (where it is well conditioned), the constraint h22 = 1 is carried through the
// change of coordinates, and the resulting 8x8 SPD system is solved in float64.
#pragma once

#include "common.cuh"
#include "score.cuh"
#include "solve.cuh"

namespace b200geo {

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// Grid: (blocks, B).  keys: B x 2, zero-initialised.
__global__ void __launch_bounds__(256)
select_kernel(const int32_t *__restrict__ counts, int k_begin, int k_end, int K_stride, double thr,
              unsigned long long *__restrict__ keys)
{
    const int b = blockIdx.y;
    counts += (size_t)b * K_stride;
    unsigned long long best = 0ull, early = 0ull;
    for (int k = k_begin + blockIdx.x * blockDim.x + threadIdx.x; k < k_end; k += gridDim.x * blockDim.x) {
        unsigned c = (unsigned)counts[k];
        unsigned long long kb = ((unsigned long long)c << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)k);
        best = kb > best ? kb : best;
        if ((double)c >= thr) {
            unsigned long long ke = 0x100000000ull - (unsigned long long)k;
            early = ke > early ? ke : early;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long ob = shfl_xor_u64(best, m), oe = shfl_xor_u64(early, m);
        best = ob > best ? ob : best;
        early = oe > early ? oe : early;
    }
    __shared__ unsigned long long sb[8], se[8];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sb[warp] = best; se[warp] = early; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            best = sb[w] > best ? sb[w] : best;
            early = se[w] > early ? se[w] : early;
        }
        if (best) atomicMax(keys + 2 * b, best);
        if (early) atomicMax(keys + 2 * b + 1, early);
    }
}

// winner index from the two keys; -1 when every hypothesis scored 0 (ransac.py:206 would fail)
__device__ __forceinline__ int decode_winner(unsigned long long kbest, unsigned long long kearly)
{
    if (kearly != 0ull) return (int)(0x100000000ull - kearly);
    if ((kbest >> 32) == 0ull) return -1;
    return (int)(0xFFFFFFFFu - (unsigned)(kbest & 0xFFFFFFFFull));
}

constexpr int REFIT_THREADS = 1024;
constexpr int REFIT_NSUM = 24;

// Gram sums over the masked rows, then the constrained solve.  Shared by finish_kernel and
// refit_kernel.  `inl` tells whether this thread's row j participates.  All threads must call.
struct RefitShared {
    double red[32][REFIT_NSUM];
    double bc[12];
};

__device__ __forceinline__ void block_reduce_sums(double (&v)[REFIT_NSUM], int n, RefitShared &sh)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < n; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) sh.red[warp][i] = s;
    }
    __syncthreads();
    if (warp == 0) {
        for (int i = 0; i < n; ++i) {
            double s = lane < nw ? sh.red[lane][i] : 0.0;
            s = warp_sum(s);
            if (lane == 0) sh.red[0][i] = s;
        }
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) v[i] = sh.red[0][i];
    __syncthreads();
}

// mask-driven least squares; returns status in out[0..9] = {status, H[9]}; all threads call.
__device__ void refit_block(const float4 *__restrict__ pts, int M, const uint8_t *__restrict__ mask,
                            RefitShared &sh, double *out_H, int *out_status, int *out_count)
{
    // pass 1: count, centroid and extent of the inlier rows (both point sets)
    double v[REFIT_NSUM];
#pragma unroll
    for (int i = 0; i < REFIT_NSUM; ++i) v[i] = 0.0;
    float lo[4] = {3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f}, hi[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        if (mask && !mask[j]) continue;
        float4 p = __ldg(pts + j);
        v[0] += 1.0; v[1] += p.x; v[2] += p.y; v[3] += p.z; v[4] += p.w;
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        lo[3] = fminf(lo[3], p.w); hi[3] = fmaxf(hi[3], p.w);
    }
    // extents ride along as sums of (per-thread max - min) upper bounds: only a scale is needed
    v[5] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        lo[i] = -lo[i];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            lo[i] = fmaxf(lo[i], __shfl_xor_sync(0xffffffffu, lo[i], m));
            hi[i] = fmaxf(hi[i], __shfl_xor_sync(0xffffffffu, hi[i], m));
        }
    }
    __shared__ float ext[32][8];
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { ext[warp][i] = lo[i]; ext[warp][4 + i] = hi[i]; }
        }
    }
    block_reduce_sums(v, 5, sh);     // contains the __syncthreads that publish ext[][]
    const int nw = blockDim.x >> 5;
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        e[i] = ext[0][i];
        for (int w = 1; w < nw; ++w) e[i] = fmaxf(e[i], ext[w][i]);
    }
    const double n = v[0];
    *out_count = (int)n;
    if (n < 4.0) {
        if (threadIdx.x == 0) *out_status = B200GEO_ERR_SINGULAR;
        return;
    }
    const double cx = v[1] / n, cy = v[2] / n, cX = v[3] / n, cY = v[4] / n;
    // extents: max + (-min)
    double es = fmax((double)e[4] + (double)e[0], (double)e[5] + (double)e[1]);
    double ed = fmax((double)e[6] + (double)e[2], (double)e[7] + (double)e[3]);
    const double ss = es > 0.0 ? 2.0 / es : 1.0, sd = ed > 0.0 ? 2.0 / ed : 1.0;

    // pass 2: P = sum p p^T, PX = sum X p p^T, PY = sum Y p p^T, PR = sum (X^2+Y^2) p p^T,  p = (x, y, 1)
#pragma unroll
    for (int i = 0; i < REFIT_NSUM; ++i) v[i] = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        if (mask && !mask[j]) continue;
        float4 q = __ldg(pts + j);
        double x = ((double)q.x - cx) * ss, y = ((double)q.y - cy) * ss;
        double X = ((double)q.z - cX) * sd, Y = ((double)q.w - cY) * sd;
        double pp[6] = {x * x, x * y, x, y * y, y, 1.0};
        double R = X * X + Y * Y;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            v[i] += pp[i];
            v[6 + i] += X * pp[i];
            v[12 + i] += Y * pp[i];
            v[18 + i] += R * pp[i];
        }
    }
    block_reduce_sums(v, REFIT_NSUM, sh);

    if (threadIdx.x == 0) {
        // symmetric 3x3 blocks from the 6 packed sums (xx, xy, x, yy, y, 1)
        auto blk = [&](int base, int r, int c) -> double {
            const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
            return v[base + map[r][c]];
        };
        // G (9x9) = [[P,0,-PX],[0,P,-PY],[-PX,-PY,PR]]
        double G[9][9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double P = blk(0, r, c), PX = blk(6, r, c), PY = blk(12, r, c), PR = blk(18, r, c);
                G[r][c] = P;          G[r][3 + c] = 0.0;      G[r][6 + c] = -PX;
                G[3 + r][c] = 0.0;    G[3 + r][3 + c] = P;    G[3 + r][6 + c] = -PY;
                G[6 + r][c] = -PX;    G[6 + r][3 + c] = -PY;  G[6 + r][6 + c] = PR;
            }
        // constraint h22 = 1 in normalised coordinates:  g8 = 1 + a*g6 + b*g7,  a = ss*cx, b = ss*cy
        // g = e8 + E y,  E = [I8 ; (0..0, a, b)]  ->  (E^T G E) y = -E^T G e8
        const double ea[8] = {0, 0, 0, 0, 0, 0, ss * cx, ss * cy};
        double A[8][9];
        for (int r = 0; r < 8; ++r) {
            for (int c = 0; c < 8; ++c)
                A[r][c] = G[r][c] + ea[r] * G[8][c] + G[r][8] * ea[c] + ea[r] * G[8][8] * ea[c];
            A[r][8] = -(G[r][8] + ea[r] * G[8][8]);
        }
        double yv[8];
        bool ok = gauss_solve<double, 8>(A, yv);
        double g[9];
        for (int i = 0; i < 8; ++i) g[i] = yv[i];
        g[8] = 1.0 + ea[6] * yv[6] + ea[7] * yv[7];
        // H = Td^-1 * G * Ts
        const double dd = 1.0 / sd;
        double m[9];
        for (int r = 0; r < 3; ++r) {
            m[3 * r + 0] = g[3 * r + 0] * ss;
            m[3 * r + 1] = g[3 * r + 1] * ss;
            m[3 * r + 2] = g[3 * r + 2] - (g[3 * r + 0] * cx + g[3 * r + 1] * cy) * ss;
        }
        double h[9];
        for (int c = 0; c < 3; ++c) {
            h[c] = m[c] * dd + cX * m[6 + c];
            h[3 + c] = m[3 + c] * dd + cY * m[6 + c];
            h[6 + c] = m[6 + c];
        }
        const double inv = 1.0 / h[8];
        for (int i = 0; i < 8; ++i) { h[i] *= inv; ok = ok && (fabs(h[i]) <= 1.0e300); }
        h[8] = 1.0;
        for (int i = 0; i < 9; ++i) out_H[i] = h[i];
        *out_status = ok ? 0 : B200GEO_ERR_SINGULAR;
    }
}

// K3.  Grid (B); block REFIT_THREADS.  result: B x 32 doubles (layout in b200geo.h).
template <typename T>
__global__ void __launch_bounds__(REFIT_THREADS)
finish_kernel(const float4 *__restrict__ pts, int M, const int4 *__restrict__ idx, int K_stride,
              const unsigned long long *__restrict__ keys, float th, int method, double *__restrict__ result,
              uint8_t *__restrict__ mask)
{
    __shared__ RefitShared sh;
    __shared__ float sH[9], sHi[9];
    __shared__ int sWin, sStatus, sCount;
    const int b = blockIdx.x;
    pts += (size_t)b * M;
    mask += (size_t)b * M;
    result += (size_t)b * 32;

    if (threadIdx.x == 0) {
        int w = decode_winner(keys[2 * b], keys[2 * b + 1]);
        sWin = w;
        sStatus = 0;
        if (w >= 0) {
            float Hf[9], Hif[9];
            int4 id = __ldg(idx + (size_t)b * K_stride + w);
            bool deg = solve_sample<T>(pts, id, Hf, Hif, true);
            for (int j = 0; j < 9; ++j) { sH[j] = Hf[j]; sHi[j] = Hif[j]; }
            if (deg) sStatus = B200GEO_ERR_SINGULAR;
        } else {
            sStatus = B200GEO_ERR_NOMODEL;
        }
    }
    __syncthreads();
    const int win = sWin;
    if (win < 0 || sStatus != 0) {
        for (int j = threadIdx.x; j < M; j += blockDim.x) mask[j] = 0;
        if (threadIdx.x == 0) {
            result[0] = (double)win; result[1] = 0.0; result[2] = (double)sStatus;
            for (int j = 3; j < 32; ++j) result[j] = 0.0;
        }
        return;
    }
    float h[9], hi[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { h[j] = sH[j]; hi[j] = sHi[j]; }
    // inliers_pos = err_total < th   (ransac.py:183)
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        float4 p = __ldg(pts + j);
        float e;
        if (method == B200GEO_FWD) e = exact_loss<B200GEO_FWD>(h, hi, p);
        else if (method == B200GEO_BACKWARD) e = exact_loss<B200GEO_BACKWARD>(h, hi, p);
        else e = exact_loss<B200GEO_REPROJ>(h, hi, p);
        mask[j] = (uint8_t)(e < th);
    }
    __syncthreads();   // mask[] written by this block is read below (global, same block)
    __threadfence_block();
    refit_block(pts, M, mask, sh, result + 12, &sStatus, &sCount);
    __syncthreads();
    if (threadIdx.x == 0) {
        result[0] = (double)win;
        result[1] = (double)sCount;
        result[2] = (double)sStatus;
        for (int j = 0; j < 9; ++j) result[3 + j] = (double)sH[j];
        for (int j = 21; j < 32; ++j) result[j] = 0.0;
    }
}

__global__ void __launch_bounds__(REFIT_THREADS)
refit_kernel(const float4 *__restrict__ pts, int M, const uint8_t *__restrict__ mask, double *__restrict__ result)
{
    __shared__ RefitShared sh;
    __shared__ int sStatus, sCount;
    if (threadIdx.x == 0) sStatus = 0;
    __syncthreads();
    refit_block(pts, M, mask, sh, result + 1, &sStatus, &sCount);
    __syncthreads();
    if (threadIdx.x == 0) result[0] = (double)sStatus;
}

}  // namespace b200geo
