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

// The constrained 8x8 solve from the 24 Gram sums v[] (P, PX, PY, PR blocks of 6: xx, xy, x, yy, y, 1) accumulated in the
// normalised coordinates x = (x - cx) ss, X = (X - cX) sd ...  One thread; the 9x9 / 8x9 matrices live in shared memory
// (in registers they spilled 748 bytes of finish_kernel's frame).  Returns false for a singular system.
struct RefitSolveMem { double G[9][9]; double A[8][9]; };
__device__ __noinline__ bool refit_solve(const double *v, double cx, double cy, double cX, double cY, double ss, double sd,
                                         RefitSolveMem &m_, double *out_H)
{
    double (&G)[9][9] = m_.G;
    double (&A)[8][9] = m_.A;
    auto blk = [&](int base, int r, int c) -> double {
        const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        return v[base + map[r][c]];
    };
    // G (9x9) = [[P,0,-PX],[0,P,-PY],[-PX,-PY,PR]]
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double P = blk(0, r, c), PX = blk(6, r, c), PY = blk(12, r, c), PR = blk(18, r, c);
            G[r][c] = P;          G[r][3 + c] = 0.0;      G[r][6 + c] = -PX;
            G[3 + r][c] = 0.0;    G[3 + r][3 + c] = P;    G[3 + r][6 + c] = -PY;
            G[6 + r][c] = -PX;    G[6 + r][3 + c] = -PY;  G[6 + r][6 + c] = PR;
        }
    // constraint h22 = 1 in normalised coordinates:  g8 = 1 + a*g6 + b*g7,  a = ss*cx, b = ss*cy
    // g = e8 + E y,  E = [I8 ; (0..0, a, b)]  ->  (E^T G E) y = -E^T G e8
    const double ea6 = ss * cx, ea7 = ss * cy;
    for (int r = 0; r < 8; ++r) {
        const double er = r == 6 ? ea6 : (r == 7 ? ea7 : 0.0);
        for (int c = 0; c < 8; ++c) {
            const double ec = c == 6 ? ea6 : (c == 7 ? ea7 : 0.0);
            A[r][c] = G[r][c] + er * G[8][c] + G[r][8] * ec + er * G[8][8] * ec;
        }
        A[r][8] = -(G[r][8] + er * G[8][8]);
    }
    double yv[8];
    bool ok = gauss_solve<double, 8>(A, yv);
    double g[9];
    for (int i = 0; i < 8; ++i) g[i] = yv[i];
    g[8] = 1.0 + ea6 * yv[6] + ea7 * yv[7];
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
    return ok;
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
        __shared__ RefitSolveMem sm_solve;
        const bool ok = refit_solve(v, cx, cy, cX, cY, ss, sd, sm_solve, out_H);
        *out_status = ok ? 0 : B200GEO_ERR_SINGULAR;
    }
}

// K3.  Grid (G, B), block FINISH_THREADS: CTA (g, b) takes the correspondences [g * per, (g + 1) * per) of pair b.
// Every CTA decodes the winner and re-solves its four points (one thread, ~200 float64 operations), writes its slice of the
// inlier mask with the reference-order chain and accumulates the Gram sums of its inliers; the last CTA of a pair to finish
// (a counter in global memory, reset for the next launch) adds the G partial sums in index order -- the result does not depend
// on the arrival order -- and solves.  The normalisation (centre and scale of the coordinates the sums are taken in) must be
// known before the inliers are: the centroid of the winning sample's four points and the largest coordinate magnitude of
// the set (`cmax`, already computed for the scoring band); the least-squares minimiser does not depend on it, only the
// conditioning of its float64 evaluation does.  (Round 1: one 1024-thread CTA, two passes, 67 us at M = 20000 and 748 bytes
// of spills -- the non-sharded part of the multi-GPU step.)
// result: B x 32 doubles (layout in b200geo.h); entries 3..20 are NaN whenever the status is not 0.
// What a rank contributes to the choice of the winner when the sample sets are sharded AND each rank holds only its own shard of
// them: its two keys and the four correspondence indices of the sample sets they name (the other ranks cannot look those up).
struct __align__(16) WinRec {
    unsigned long long best, early;      // keys[0], keys[1] of this rank's hypothesis block
    int4 ibest, iearly;                  // idx[k] of the sample set each key names (unused when the key is 0)
};
static_assert(sizeof(WinRec) == 48, "WinRec is exchanged as 48 bytes per rank");

// Grid (1), block (32): keys (2 x u64, this rank's block reduced) + idx -> rec.  idx is indexed by absolute k.
__global__ void pack_winner_kernel(const unsigned long long *__restrict__ keys, const int4 *__restrict__ idx, WinRec *__restrict__ rec)
{
    if (threadIdx.x != 0) return;
    WinRec r;
    r.best = keys[0]; r.early = keys[1];
    r.ibest = make_int4(0, 0, 0, 0); r.iearly = make_int4(0, 0, 0, 0);
    if ((r.best >> 32) != 0ull) r.ibest = __ldg(idx + (0xFFFFFFFFu - (unsigned)(r.best & 0xFFFFFFFFull)));
    if (r.early != 0ull) r.iearly = __ldg(idx + (unsigned)(0x100000000ull - r.early));
    *rec = r;
}

// ---- the 48-byte winner records of a sharded search exchanged over peer memory (NVLink / NVSwitch) instead of a collective ----
// Every rank owns an inbox in its own HBM: 2 parities x world slots of 64 bytes (record + epoch flag), mapped into every other
// rank's address space (CUDA IPC).  peer_publish_kernel: thread r writes this rank's record into rank r's inbox (six 8-byte
// peer stores, a system-scope fence, then the epoch as the flag).  peer_gather_kernel: thread r waits for rank r's flag in the
// local inbox and copies the record into the contiguous array finish_kernel reads.  The parity alternates per exchange: a
// rank can be at most one exchange ahead of the slowest one (it needs that rank's next record to get further), so a slot is
// never overwritten before its reader has consumed it.
constexpr int PEER_MAX = 64;
struct PeerBoxes { unsigned long long *box[PEER_MAX]; };

__global__ void peer_publish_kernel(const WinRec *__restrict__ rec, const PeerBoxes P, int world, int rank, unsigned long long epoch)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    unsigned long long *slot = P.box[r] + ((size_t)(epoch & 1ull) * world + rank) * 8;
    const unsigned long long *s = reinterpret_cast<const unsigned long long *>(rec);
#pragma unroll
    for (int i = 0; i < 6; ++i) slot[i] = s[i];
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot + 7), "l"(epoch) : "memory");
}

// budget: spin iterations before giving up (a rank that never publishes must not hang this GPU): the record is then zero,
// which finish_kernel reports as "no model"
__global__ void peer_gather_kernel(unsigned long long *__restrict__ inbox, int world, unsigned long long epoch, WinRec *__restrict__ recs,
                                   unsigned long long budget)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    unsigned long long *slot = inbox + ((size_t)(epoch & 1ull) * world + r) * 8;
    unsigned long long f = 0ull, spins = 0ull;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(slot + 7) : "memory");
        if (f == epoch || ++spins > budget) break;
        __nanosleep(64);
    }
    unsigned long long *d = reinterpret_cast<unsigned long long *>(recs) + (size_t)r * 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = f == epoch ? slot[i] : 0ull;
}

constexpr int FINISH_THREADS = 256;
constexpr int FINISH_NSUM = 25;      // inlier count + the 24 Gram sums

template <typename T>
__global__ void __launch_bounds__(FINISH_THREADS)
finish_kernel(const float4 *__restrict__ pts, int M, const int4 *__restrict__ idx, int K_stride,
              const unsigned long long *__restrict__ keys, const WinRec *__restrict__ recs, int n_recs, float th, int method,
              const float *__restrict__ cmax, int per, double *__restrict__ partial, unsigned *__restrict__ done,
              double *__restrict__ result, uint8_t *__restrict__ mask)
{
    __shared__ double red[FINISH_THREADS / 32][FINISH_NSUM];
    __shared__ float sH[9], sHi[9];
    __shared__ double sC[4];
    __shared__ int sWin, sStatus, sLast;
    __shared__ RefitSolveMem sm_solve;
    const int b = blockIdx.y, G = gridDim.x;
    pts += (size_t)b * M;
    mask += (size_t)b * M;
    result += (size_t)b * 32;
    partial += ((size_t)b * G + blockIdx.x) * FINISH_NSUM;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    if (threadIdx.x == 0) {
        int w;
        int4 id = make_int4(0, 0, 0, 0);
        if (recs) {          // (B = 1) the per-rank records of a sharded search: the max of the keys, and the indices that came with it
            unsigned long long kb = 0ull, ke = 0ull;
            int4 ib = id, ie = id;
            for (int r = 0; r < n_recs; ++r) {
                const WinRec q = recs[r];
                if (q.best > kb) { kb = q.best; ib = q.ibest; }
                if (q.early > ke) { ke = q.early; ie = q.iearly; }
            }
            w = decode_winner(kb, ke);
            id = ke != 0ull ? ie : ib;
        } else {
            w = decode_winner(keys[2 * b], keys[2 * b + 1]);
            if (w >= 0) id = __ldg(idx + (size_t)b * K_stride + w);
        }
        sWin = w;
        sStatus = 0;
        if (w >= 0) {
            float Hf[9], Hif[9];
            if (((unsigned)id.x >= (unsigned)M) | ((unsigned)id.y >= (unsigned)M) | ((unsigned)id.z >= (unsigned)M) | ((unsigned)id.w >= (unsigned)M))
                id = make_int4(0, 0, 0, 0);        // (cannot win: such a sample set is flagged and scores 0)
            bool deg = solve_sample<T>(pts, id, Hf, Hif, true);
            for (int j = 0; j < 9; ++j) { sH[j] = Hf[j]; sHi[j] = Hif[j]; }
            if (deg) sStatus = B200GEO_ERR_SINGULAR;
            const float4 p0 = __ldg(pts + id.x), p1 = __ldg(pts + id.y), p2 = __ldg(pts + id.z), p3 = __ldg(pts + id.w);
            sC[0] = 0.25 * ((double)p0.x + p1.x + p2.x + p3.x); sC[1] = 0.25 * ((double)p0.y + p1.y + p2.y + p3.y);
            sC[2] = 0.25 * ((double)p0.z + p1.z + p2.z + p3.z); sC[3] = 0.25 * ((double)p0.w + p1.w + p2.w + p3.w);
        } else {
            sStatus = B200GEO_ERR_NOMODEL;
        }
    }
    __syncthreads();
    const int win = sWin;
    const int j0 = blockIdx.x * per, j1 = min(M, j0 + per);
    if (win < 0 || sStatus != 0) {
        for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) mask[j] = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            result[0] = (double)win; result[1] = 0.0; result[2] = (double)sStatus;
            for (int j = 3; j < 21; ++j) result[j] = qnan;
            for (int j = 21; j < 32; ++j) result[j] = 0.0;
        }
        return;
    }
    float h[9], hi[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { h[j] = sH[j]; hi[j] = sHi[j]; }
    const double cx = sC[0], cy = sC[1], cX = sC[2], cY = sC[3];
    const double C = (double)cmax[b];
    const double ss = (C > 0.0 && C < 1.0e300) ? 1.0 / C : 1.0, sd = ss;
    double v[FINISH_NSUM];
#pragma unroll
    for (int i = 0; i < FINISH_NSUM; ++i) v[i] = 0.0;
    // inliers_pos = err_total < th (ransac.py:183); P = sum p p^T, PX = sum X p p^T, PY = sum Y p p^T, PR = sum (X^2+Y^2) p p^T, p = (x, y, 1)
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const float4 q = __ldg(pts + j);
        float e;
        if (method == B200GEO_FWD) e = exact_loss<B200GEO_FWD>(h, hi, q);
        else if (method == B200GEO_BACKWARD) e = exact_loss<B200GEO_BACKWARD>(h, hi, q);
        else e = exact_loss<B200GEO_REPROJ>(h, hi, q);
        const bool in = e < th;
        mask[j] = (uint8_t)in;
        if (in) {
            const double x = ((double)q.x - cx) * ss, y = ((double)q.y - cy) * ss;
            const double X = ((double)q.z - cX) * sd, Y = ((double)q.w - cY) * sd;
            const double pp[6] = {x * x, x * y, x, y * y, y, 1.0};
            const double R = X * X + Y * Y;
            v[24] += 1.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                v[i] += pp[i];
                v[6 + i] += X * pp[i];
                v[12 + i] += Y * pp[i];
                v[18 + i] += R * pp[i];
            }
        }
    }
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < FINISH_NSUM; ++i) {
            const double sacc = warp_sum(v[i]);
            if (lane == 0) red[warp][i] = sacc;
        }
        __syncthreads();
        if (threadIdx.x < FINISH_NSUM) {
            double sacc = 0.0;
            for (int w = 0; w < FINISH_THREADS / 32; ++w) sacc += red[w][threadIdx.x];
            partial[threadIdx.x] = sacc;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(done + b, 1u);
        sLast = (t == (unsigned)G - 1u);
        if (sLast) done[b] = 0u;               // ready for the next launch (stream order)
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    // last CTA of the pair: add the partial sums in index order, solve
    if (threadIdx.x < FINISH_NSUM) {
        const double *p = partial - (size_t)blockIdx.x * FINISH_NSUM + threadIdx.x;
        double sacc = 0.0;
        for (int g = 0; g < G; ++g) sacc += __ldcg(p + (size_t)g * FINISH_NSUM);
        red[0][threadIdx.x] = sacc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double n = red[0][24];
        int status = 0;
        double Hout[9];
        for (int j = 0; j < 9; ++j) Hout[j] = qnan;
        // the reference's fit(collective=True) needs more than n = 4 columns (ransac.py:38); 4 exact inliers still have a
        // unique solution here, fewer do not
        if (n < 4.0) status = B200GEO_ERR_SINGULAR;
        else if (!refit_solve(red[0], cx, cy, cX, cY, ss, sd, sm_solve, Hout)) status = B200GEO_ERR_SINGULAR;
        result[0] = (double)win;
        result[1] = n;
        result[2] = (double)status;
        for (int j = 0; j < 9; ++j) result[3 + j] = (double)sH[j];
        for (int j = 0; j < 9; ++j) result[12 + j] = status == 0 ? Hout[j] : qnan;
        for (int j = 21; j < 32; ++j) result[j] = 0.0;
    }
}

// S6: calcHomography(u, v, collective=True) (homography.py:30-46, 71-88): the unit-norm minimiser of |A h| over the 2M x 9
// matrix of calc_correspCollective -- the last right singular vector of A, scaled to h22 = 1.  The reference takes a float32 SVD
// of A; here the 9 x 9 Gram matrix A^T A is accumulated in float64 from the same float32 rows (products x x', ... rounded to
// float32 as the reference's matrix holds them) and its smallest eigenvector is found by cyclic Jacobi rotations, which
// resolve small eigenvalues of a graded matrix to high relative accuracy (measured against a float64 SVD of A: 5e-11 on the
// fixture, where LAPACK's tridiagonal eigensolver on the same Gram matrix loses 3e-4).  One CTA; result: [status, H[9]].
__global__ void __launch_bounds__(REFIT_THREADS)
homography_svd_kernel(const float4 *__restrict__ pts, int M, double *__restrict__ result)
{
    __shared__ RefitShared sh;
    __shared__ double G[9][9], V[9][9];
    // 30 sums: a a^T (6: xx xy x yy y 1), a (x) cx (9), a (x) cy (9), cx cx^T + cy cy^T (6), with a = (x, y, 1),
    // cx = (x X, y X, X), cy = (x Y, y Y, Y)
    double v[REFIT_NSUM], v2[6];
#pragma unroll
    for (int i = 0; i < REFIT_NSUM; ++i) v[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) v2[i] = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const float4 q = __ldg(pts + j);
        const double a[3] = {(double)q.x, (double)q.y, 1.0};
        const double cx[3] = {(double)__fmul_rn(q.x, q.z), (double)__fmul_rn(q.y, q.z), (double)q.z};
        const double cy[3] = {(double)__fmul_rn(q.x, q.w), (double)__fmul_rn(q.y, q.w), (double)q.w};
        v[0] += a[0] * a[0]; v[1] += a[0] * a[1]; v[2] += a[0]; v[3] += a[1] * a[1]; v[4] += a[1]; v[5] += 1.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { v[6 + 3 * r + c] += a[r] * cx[c]; v[15 + 3 * r + c] += a[r] * cy[c]; }
        v2[0] += cx[0] * cx[0] + cy[0] * cy[0]; v2[1] += cx[0] * cx[1] + cy[0] * cy[1]; v2[2] += cx[0] * cx[2] + cy[0] * cy[2];
        v2[3] += cx[1] * cx[1] + cy[1] * cy[1]; v2[4] += cx[1] * cx[2] + cy[1] * cy[2]; v2[5] += cx[2] * cx[2] + cy[2] * cy[2];
    }
    block_reduce_sums(v, REFIT_NSUM, sh);
    double w6[REFIT_NSUM];
#pragma unroll
    for (int i = 0; i < REFIT_NSUM; ++i) w6[i] = i < 6 ? v2[i] : 0.0;
    block_reduce_sums(w6, 6, sh);
    if (threadIdx.x != 0) return;
    const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            // rows of A: (-a, 0, cx) and (0, -a, cy)
            G[r][c] = v[map[r][c]];       G[r][3 + c] = 0.0;                G[r][6 + c] = -v[6 + 3 * r + c];
            G[3 + r][c] = 0.0;            G[3 + r][3 + c] = v[map[r][c]];   G[3 + r][6 + c] = -v[15 + 3 * r + c];
            G[6 + r][c] = -v[6 + 3 * c + r]; G[6 + r][3 + c] = -v[15 + 3 * c + r]; G[6 + r][6 + c] = w6[map[r][c]];
        }
    for (int r = 0; r < 9; ++r)
        for (int c = 0; c < 9; ++c) V[r][c] = r == c ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 9; ++p) {
            diag += G[p][p] * G[p][p];
            for (int q = p + 1; q < 9; ++q) off += G[p][q] * G[p][q];
        }
        if (!(off > 1e-60 * diag)) break;
        for (int p = 0; p < 8; ++p)
            for (int q = p + 1; q < 9; ++q) {
                const double apq = G[p][q];
                if (fabs(apq) < 1e-300 || fabs(apq) <= 1e-19 * sqrt(fabs(G[p][p] * G[q][q]))) continue;
                const double theta = (G[q][q] - G[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 9; ++k) {       // columns p, q
                    const double gkp = G[k][p], gkq = G[k][q];
                    G[k][p] = c * gkp - sn * gkq; G[k][q] = sn * gkp + c * gkq;
                }
                for (int k = 0; k < 9; ++k) {       // rows p, q
                    const double gpk = G[p][k], gqk = G[q][k];
                    G[p][k] = c * gpk - sn * gqk; G[q][k] = sn * gpk + c * gqk;
                }
                for (int k = 0; k < 9; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int kmin = 0;
    for (int k = 1; k < 9; ++k)
        if (G[k][k] < G[kmin][kmin]) kmin = k;
    const double h8 = V[8][kmin];
    bool ok = fabs(h8) > 1e-300 && M >= 4;
    for (int i = 0; i < 9; ++i) {
        const double hv = V[i][kmin] / h8;         // h / h[2,2] (homography.py:83-84)
        result[1 + i] = hv;
        ok = ok && fabs(hv) <= 1.0e300;
    }
    result[0] = ok ? 0.0 : (double)B200GEO_ERR_SINGULAR;
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
