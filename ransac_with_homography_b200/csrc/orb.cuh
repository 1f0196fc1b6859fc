// orb.cuh -- K7: cv2.ORB_create().detectAndCompute(gray, None) on the GPU (reference ransac.py:255-257: the feature
// front-end of stitching()).  The algorithm lives in OpenCV (third-party, unpinned by the
// reference; this image has 4.13.0: features2d/src/orb.cpp, fast.cpp, fast_score.cpp, keypoint.cpp, imgproc/src/resize.cpp,
// filter.simd.hpp, core/src/mathfuncs_core.simd.hpp).  Default parameters as the reference uses them: 8 levels, scale 1.2f,
// edge threshold 31, patch 31, FAST threshold 20, Harris score, WTA_K 2.  Every stage reproduces OpenCV's integer / float32
// arithmetic operation for operation, so keypoints (position, size, angle, response, octave, ORDER) and descriptors are
// bit-identical to cv2's (the test suite holds a CPU restatement pinned to cv2 stage by stage).
//
//   orb_level_kernel      level L (with its 32 px BORDER_REFLECT_101 frame) from level L-1: INTER_LINEAR_EXACT resize,
//                         8.8 fixed-point coefficients from float64 coordinates, (sum + 2^15) >> 16
//   orb_fast_kernel       FAST-9/16 corner score of every pixel of every level (0 = no corner)
//   orb_rowflags / rowscan / rowwrite / harris kernels: 3 x 3 non-maximum suppression + border filter -> candidate records per
//                         level in FAST's row-major order (ballot words + row scan, no atomics, no sort) + Harris response;
//                         the host runs retainBest = std::nth_element + std::partition twice on them, as KeyPointsFilter does
//   orb_angle_kernel      intensity-centroid orientation, one warp per keypoint, cv::fastAtan2 in float32
//   orb_blur_kernel       7 x 7 Gaussian, sigma 2: sepFilter2D's float32 path in the operation order of OpenCV's AVX2 build
//   orb_describe_kernel   256 rotated pair tests per keypoint, one warp per keypoint (cos / sin come from the host's libm,
//                         the one OpenCV calls)
#pragma once

#include "common.cuh"

namespace b200geo {

constexpr int ORB_LEVELS = 8, ORB_BORDER = 32, ORB_EDGE = 31, ORB_HALF_PATCH = 15, ORB_PATCH = 31, ORB_FAST_T = 20;

struct OrbLevel {
    int w, h;            // level size
    int pitch;           // bytes per row of the framed level (w + 64, rounded up to 16)
    int pw, ph;          // size of the previous level (the resize source); level 0: the input
    long long ofs;       // byte offset of the framed level's first byte in the pyramid buffer
    unsigned cand_ofs;   // first candidate record of this level
    unsigned cand_cap;
};
struct OrbLevels { OrbLevel lv[ORB_LEVELS]; int n; };

struct OrbCandidate { unsigned pos; unsigned score; float harris; };      // pos = y << 16 | x
struct OrbPoint { int level, x, y; float a, b; };                        // keypoint on its level; a = cosf, b = sinf of its angle

__device__ __forceinline__ int orb_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// interpolationLinear<uint8_t>::getCoeffs (resize.cpp): source offset and the 8.8 weight of the second tap
__device__ __forceinline__ void orb_coeff(int d, int src, int dst, int &ofs, int &c1)
{
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
    const double f = __dsub_rn(__dmul_rn(scale, (double)d + 0.5), 0.5);
    const int i = (int)floor(f);
    ofs = 0; c1 = 0;
    if (i >= 0 && src > 1) {
        if (i < src - 1) { ofs = i; c1 = __double2int_rn(__dmul_rn(__dsub_rn(f, (double)i), 256.0)); }
        else ofs = src - 1;
    }
}

// One framed level.  level0 = 1: copy of the grey input; 3: cv2.cvtColor(COLOR_RGB2GRAY) of the interleaved RGB input
// ((9798 R + 19235 G + 3735 B + 2^14) >> 15, OpenCV's 8-bit path); 0: INTER_LINEAR_EXACT resize of the previous level's interior.
__global__ void __launch_bounds__(256)
orb_level_kernel(const uint8_t *__restrict__ src, long long src_pitch, int sw, int sh, uint8_t *__restrict__ dst, int pitch, int w, int h, int level0)
{
    const int ex = blockIdx.x * blockDim.x + threadIdx.x, ey = blockIdx.y;
    if (ex >= w + 2 * ORB_BORDER || ey >= h + 2 * ORB_BORDER) return;
    const int x = orb_reflect101(ex - ORB_BORDER, w), y = orb_reflect101(ey - ORB_BORDER, h);
    uint8_t v;
    if (level0 == 1) v = src[(long long)y * src_pitch + x];
    else if (level0 == 3) {
        const uint8_t *q = src + (long long)y * src_pitch + 3 * x;
        v = (uint8_t)((9798u * q[0] + 19235u * q[1] + 3735u * q[2] + (1u << 14)) >> 15);
    }
    else {
        int xo, cx, yo, cy;
        orb_coeff(x, sw, w, xo, cx);
        orb_coeff(y, sh, h, yo, cy);
        const int xn = min(xo + 1, sw - 1), yn = min(yo + 1, sh - 1);
        const uint8_t *r0 = src + (long long)yo * src_pitch, *r1 = src + (long long)yn * src_pitch;
        const unsigned h0 = r0[xo] * (256 - cx) + r0[xn] * cx, h1 = r1[xo] * (256 - cx) + r1[xn] * cx;
        v = (uint8_t)((h0 * (unsigned)(256 - cy) + h1 * (unsigned)cy + (1u << 15)) >> 16);
    }
    dst[(long long)ey * pitch + ex] = v;
}

// FAST-9/16 (fast.cpp, fast_score.cpp): score = max over the 16 arcs of 9 consecutive ring pixels of the smallest
// difference to the centre (brighter or darker), minus 1; stored where it exceeds the threshold, else 0.
// Grid: (ceil(max_w / 256), max_h, levels).  Scores live in a buffer of the pyramid's geometry (frame = 0).
__global__ void __launch_bounds__(256)
orb_fast_kernel(const uint8_t *__restrict__ pyr, uint8_t *__restrict__ score, const OrbLevels L)
{
    const OrbLevel &lv = L.lv[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x + 3, y = blockIdx.y + 3;
    if (x >= lv.w - 3 || y >= lv.h - 3) return;
    const uint8_t *c = pyr + lv.ofs + (long long)(y + ORB_BORDER) * lv.pitch + (x + ORB_BORDER);
    const int p = lv.pitch, v = c[0];
    int d[16];
    d[0] = c[3 * p] - v;       d[1] = c[3 * p + 1] - v;   d[2] = c[2 * p + 2] - v;   d[3] = c[p + 3] - v;
    d[4] = c[3] - v;           d[5] = c[-p + 3] - v;      d[6] = c[-2 * p + 2] - v;  d[7] = c[-3 * p + 1] - v;
    d[8] = c[-3 * p] - v;      d[9] = c[-3 * p - 1] - v;  d[10] = c[-2 * p - 2] - v; d[11] = c[-p - 3] - v;
    d[12] = c[-3] - v;         d[13] = c[p - 3] - v;      d[14] = c[2 * p - 2] - v;  d[15] = c[3 * p - 1] - v;
    // min / max over the 16 windows of 9 consecutive ring pixels by doubling (windows of 2, 4, 8, then one more element).
    // (The direct form -- for every start k a running min and max over d[(k + j) & 15] -- is miscompiled by nvcc 12.9 for
    // sm_100a: the min / max chains are fused into 3-input VIMNMX3 with the two roles exchanged, and the kernel returned the
    // largest ring difference instead; isolated reproducer and this form checked on the device against the host on 2e5 rings.)
    int m2[16], M2[16], m4[16], M4[16], m8[16], M8[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { m2[i] = min(d[i], d[(i + 1) & 15]); M2[i] = max(d[i], d[(i + 1) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { m4[i] = min(m2[i], m2[(i + 2) & 15]); M4[i] = max(M2[i], M2[(i + 2) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { m8[i] = min(m4[i], m4[(i + 4) & 15]); M8[i] = max(M4[i], M4[(i + 4) & 15]); }
    int bright = -256, dark = 256;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        bright = max(bright, min(m8[i], d[(i + 8) & 15]));
        dark = min(dark, max(M8[i], d[(i + 8) & 15]));
    }
    const int best = max(bright, -dark);
    score[lv.ofs + (long long)(y + ORB_BORDER) * lv.pitch + (x + ORB_BORDER)] = best > ORB_FAST_T ? (uint8_t)(best - 1) : (uint8_t)0;
}

// Candidates in OpenCV's order without a sort.  FAST emits its keypoints row by row, left to right, and the two retainBest
// selections that follow inherit that order (std::nth_element's result depends on the input sequence), so the candidate list
// must be row-major.  Three small kernels instead of an atomic append + a host sort:
//   orb_rowflags_kernel   one warp per row of the level's interior (31 px margin: KeyPointsFilter::runByImageBorder): strict
//                         3 x 3 non-maximum suppression of the FAST scores, one ballot word per 32 columns, the row's count
//   orb_rowscan_kernel    exclusive scan of the row counts (one block per level) -> first record of every row, level totals
//   orb_rowwrite_kernel   the same walk: position and FAST score of every flagged pixel to its slot (row offset + rank in the row)
//   orb_harris_kernel     HarrisResponses (orb.cpp: 7 x 7 block, Sobel-like integer gradients, float32 score in OpenCV's
//                         operation order), one thread per candidate
struct OrbRows { unsigned row_ofs[ORB_LEVELS]; unsigned word_ofs[ORB_LEVELS]; int words_per_row[ORB_LEVELS]; };

__global__ void __launch_bounds__(128)
orb_rowflags_kernel(const uint8_t *__restrict__ score, const OrbLevels L, const OrbRows R, unsigned *__restrict__ masks, unsigned *__restrict__ row_count)
{
    const OrbLevel &lv = L.lv[blockIdx.y];
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int rows = lv.h - 2 * ORB_EDGE;
    if (row >= rows) return;
    const int y = row + ORB_EDGE, p = lv.pitch;
    const uint8_t *srow = score + lv.ofs + (long long)(y + ORB_BORDER) * p + ORB_BORDER;
    unsigned *mrow = masks + R.word_ofs[blockIdx.y] + (size_t)row * R.words_per_row[blockIdx.y];
    unsigned n = 0;
    for (int c = 0; c < R.words_per_row[blockIdx.y]; ++c) {
        const int x = ORB_EDGE + 32 * c + lane;
        bool f = false;
        if (x < lv.w - ORB_EDGE) {
            const uint8_t *s = srow + x;
            const int v = s[0];
            f = v != 0 && v > s[-1] && v > s[1] && v > s[-p - 1] && v > s[-p] && v > s[-p + 1] && v > s[p - 1] && v > s[p] && v > s[p + 1];
        }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) mrow[c] = m;
        n += __popc(m);
    }
    if (lane == 0) row_count[R.row_ofs[blockIdx.y] + row] = n;
}

// one block of 1024 threads per level; row_count -> exclusive offsets in place, counts[level] = total (clipped to the capacity)
__global__ void __launch_bounds__(1024)
orb_rowscan_kernel(const OrbLevels L, const OrbRows R, unsigned *__restrict__ row_count, unsigned *__restrict__ counts)
{
    __shared__ unsigned part[1024];
    const OrbLevel &lv = L.lv[blockIdx.x];
    const int rows = max(lv.h - 2 * ORB_EDGE, 0);
    unsigned *rc = row_count + R.row_ofs[blockIdx.x];
    const int per = (rows + 1023) / 1024, r0 = threadIdx.x * per, r1 = min(rows, r0 + per);
    unsigned sum = 0;
    for (int r = r0; r < r1; ++r) sum += rc[r];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                       // inclusive scan of the per-thread sums
        const unsigned v = threadIdx.x >= d ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - sum;
    for (int r = r0; r < r1; ++r) { const unsigned c = rc[r]; rc[r] = run; run += c; }
    if (threadIdx.x == 1023) counts[blockIdx.x] = min(part[1023], lv.cand_cap);
}

__global__ void __launch_bounds__(128)
orb_rowwrite_kernel(const uint8_t *__restrict__ score, const OrbLevels L, const OrbRows R, const unsigned *__restrict__ masks,
                    const unsigned *__restrict__ row_first, OrbCandidate *__restrict__ cand)
{
    const OrbLevel &lv = L.lv[blockIdx.y];
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= lv.h - 2 * ORB_EDGE) return;
    const int y = row + ORB_EDGE;
    const uint8_t *srow = score + lv.ofs + (long long)(y + ORB_BORDER) * lv.pitch + ORB_BORDER;
    const unsigned *mrow = masks + R.word_ofs[blockIdx.y] + (size_t)row * R.words_per_row[blockIdx.y];
    unsigned at = row_first[R.row_ofs[blockIdx.y] + row];
    for (int c = 0; c < R.words_per_row[blockIdx.y]; ++c) {
        const unsigned m = mrow[c];
        if (m >> lane & 1u) {
            const unsigned k = at + __popc(m & ((1u << lane) - 1u));
            if (k < lv.cand_cap) {
                const int x = ORB_EDGE + 32 * c + lane;
                OrbCandidate rec;
                rec.pos = ((unsigned)y << 16) | (unsigned)x; rec.score = srow[x]; rec.harris = 0.f;
                cand[lv.cand_ofs + k] = rec;
            }
        }
        at += __popc(m);
    }
}

// Grid: (ceil(max candidates of a level / 256), levels)
__global__ void __launch_bounds__(256)
orb_harris_kernel(const uint8_t *__restrict__ pyr, const OrbLevels L, const unsigned *__restrict__ counts, OrbCandidate *__restrict__ cand)
{
    const OrbLevel &lv = L.lv[blockIdx.y];
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= counts[blockIdx.y]) return;
    OrbCandidate &rec = cand[lv.cand_ofs + k];
    const int x = (int)(rec.pos & 0xFFFFu), y = (int)(rec.pos >> 16), p = lv.pitch;
    const uint8_t *c = pyr + lv.ofs + (long long)(y + ORB_BORDER) * p + (x + ORB_BORDER);
    int a = 0, b = 0, cc = 0;
    for (int i = -3; i <= 3; ++i)
        for (int j = -3; j <= 3; ++j) {
            const uint8_t *q = c + i * p + j;
            const int Ix = (q[1] - q[-1]) * 2 + (q[-p + 1] - q[-p - 1]) + (q[p + 1] - q[p - 1]);
            const int Iy = (q[p] - q[-p]) * 2 + (q[p - 1] - q[-p - 1]) + (q[p + 1] - q[-p + 1]);
            a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
        }
    const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));                       // 1.f / ((1 << 2) * blockSize * 255.f)
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float af = (float)a, bf = (float)b, cf = (float)cc, sum = __fadd_rn(af, bf);
    rec.harris = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(af, bf), __fmul_rn(cf, cf)), __fmul_rn(__fmul_rn(0.04f, sum), sum)), s4);
}

struct OrbUmax { int u[ORB_HALF_PATCH + 2]; };

// cv::fastAtan2 (mathfuncs_core.simd.hpp, scalar form), degrees
__device__ __forceinline__ float orb_fast_atan2(float y, float x)
{
    const float p1 = __uint_as_float(0x4265226fu), p3 = __uint_as_float(0xc19556eeu), p5 = __uint_as_float(0x410e9fbfu), p7 = __uint_as_float(0xc0228ad9u);
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

// ICAngles (orb.cpp): first-order moments over the radius-15 disc; one warp per keypoint, lane = column u + 15.
__global__ void __launch_bounds__(128)
orb_angle_kernel(const uint8_t *__restrict__ pyr, const OrbLevels L, const OrbPoint *__restrict__ pts, int n, const OrbUmax um, float *__restrict__ angle)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n) return;
    const OrbPoint pt = pts[k];
    const OrbLevel &lv = L.lv[pt.level];
    const int p = lv.pitch;
    const uint8_t *c = pyr + lv.ofs + (long long)(pt.y + ORB_BORDER) * p + (pt.x + ORB_BORDER);
    int m01 = 0, m10 = 0;
    const int u = lane - ORB_HALF_PATCH;
    if (lane <= 2 * ORB_HALF_PATCH) {
        m10 = u * c[u];
        for (int v = 1; v <= ORB_HALF_PATCH; ++v) {
            if (abs(u) <= um.u[v]) {
                const int plus = c[u + v * p], minus = c[u - v * p];
                m01 += v * (plus - minus);
                m10 += u * (plus + minus);
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, s);
        m10 += __shfl_xor_sync(0xffffffffu, m10, s);
    }
    if (lane == 0) angle[k] = orb_fast_atan2((float)m01, (float)m10);
}

// GaussianBlur(level, level, Size(7, 7), 2, 2, BORDER_REFLECT_101) as ORB applies it inside its pyramid buffer: a sub-matrix
// without BORDER_ISOLATED misses cv::GaussianBlur's bit-exact fixed-point branch and goes through sepFilter2D with the
// float32 kernel.  Operation order of OpenCV 4.13's AVX2 build (pinned against cv2.sepFilter2D on 48 M random pixels in the tests):
// rows: s = k0*p0; s += k_i*p_i for i = 1..6, fused in the 32-column blocks of the vector loop, multiply-then-add in the last
// (w mod 32) columns; columns: s = k3*r0; s = fma(k_{3+i}, r_{+i} + r_{-i}, s); result = saturate(round half even).
__global__ void __launch_bounds__(256)
orb_blur_kernel(const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blurred, const OrbLevels L)
{
    const OrbLevel &lv = L.lv[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= lv.w || y >= lv.h) return;
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu), k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    const float kk[7] = {k0, k1, k2, k3, k2, k1, k0};
    const int p = lv.pitch;
    const long long o = lv.ofs + (long long)(y + ORB_BORDER) * p + (x + ORB_BORDER);
    const uint8_t *c = pyr + o;
    const bool fused = x < (lv.w / 32) * 32;
    float r[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint8_t *q = c + (i - 3) * p - 3;
        float s = __fmul_rn(kk[0], (float)q[0]);
#pragma unroll
        for (int j = 1; j < 7; ++j) s = fused ? __fmaf_rn(kk[j], (float)q[j], s) : __fadd_rn(s, __fmul_rn(kk[j], (float)q[j]));
        r[i] = s;
    }
    float s = __fmul_rn(k3, r[3]);
    s = __fmaf_rn(k2, __fadd_rn(r[4], r[2]), s);
    s = __fmaf_rn(k1, __fadd_rn(r[5], r[1]), s);
    s = __fmaf_rn(k0, __fadd_rn(r[6], r[0]), s);
    blurred[o] = (uint8_t)min(255, max(0, __float2int_rn(s)));
}

__device__ const signed char orb_pattern[512][2] = {
#include "orb_pattern.inc"
};

// computeOrbDescriptors (orb.cpp, WTA_K = 2): one warp per keypoint, lane = descriptor byte.
__global__ void __launch_bounds__(128)
orb_describe_kernel(const uint8_t *__restrict__ blurred, const OrbLevels L, const OrbPoint *__restrict__ pts, int n, uint8_t *__restrict__ desc)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n) return;
    const OrbPoint pt = pts[k];
    const OrbLevel &lv = L.lv[pt.level];
    const int p = lv.pitch;
    const uint8_t *c = blurred + lv.ofs + (long long)(pt.y + ORB_BORDER) * p + (pt.x + ORB_BORDER);
    const float a = pt.a, b = pt.b;
    unsigned val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int t[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = (lane * 8 + j) * 2 + e;
            const float px = (float)orb_pattern[idx][0], py = (float)orb_pattern[idx][1];
            const float xr = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
            const float yr = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
            t[e] = c[__float2int_rn(yr) * p + __float2int_rn(xr)];
        }
        val |= (unsigned)(t[0] < t[1]) << j;
    }
    desc[(size_t)k * 32 + lane] = (uint8_t)val;
}

}  // namespace b200geo
