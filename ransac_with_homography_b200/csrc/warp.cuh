// warp.cuh -- K4: backward perspective warp.
//
// Replaces wrapPerspective / wrapPerspectiveScan and the samplers nearestNeighbor / bilinear
// (homography.py:108-209) plus the astype(uint8) of transformImage[H] (homography.py:225, 241).
// Element semantics:
//   (zx, zy, zw) = invH . (gx[c], gy[r], 1)  in float64 ;  x = zx/zw, y = zy/zw
//   nn:       xi = trunc(x + 0.5), yi = trunc(y + 0.5); outside [0,bw-1]x[0,bh-1] -> pixel (0,0)
//   bilinear: x > bw-1 | x < 0 | y > bh-1 | y < 0 -> (0,0); else 4 taps, float64 lerp
//   source pixel (0,0) reads as 0 (the reference zeroes it in place before sampling)
//   uint8 output = truncation toward zero
// This file holds the generic kernel: any channel count (3, 4), uint8 or float32 source,
// uint8 / float32 / float64 destination, arithmetic in float64 like the reference.
#pragma once

#include "common.cuh"

namespace b200geo {

template <typename SrcT> __device__ __forceinline__ double load_px(const SrcT *p) { return (double)__ldg(p); }

template <typename DstT> __device__ __forceinline__ DstT store_cvt(double v);
template <> __device__ __forceinline__ uint8_t store_cvt<uint8_t>(double v) { return (uint8_t)(int)v; }   // astype(np.uint8)
template <> __device__ __forceinline__ float store_cvt<float>(double v) { return (float)v; }
template <> __device__ __forceinline__ double store_cvt<double>(double v) { return v; }

__device__ __forceinline__ double grid_coord(double c0, double dc, double c_last, int i, int n)
{
    // np.linspace: arange(n) * step + start, last sample forced to `stop`
    return (i == n - 1) ? c_last : __dadd_rn(__dmul_rn((double)i, dc), c0);
}

// Sample the source at the real-valued position (x, y) and store one output pixel.
// nearestNeighbor (homography.py:108-121) / bilinear (homography.py:123-138).
template <typename SrcT, typename DstT, int C, int BILINEAR>
__device__ __forceinline__ void sample_store(const SrcT *__restrict__ src, size_t spitch, int iw, int ih, int bw, int bh,
                                             double x, double y, DstT *__restrict__ dst)
{
    auto tap = [&](int yy, int xx, int ch) -> double {
        if (yy < 0 || xx < 0 || yy >= ih || xx >= iw || (yy == 0 && xx == 0)) return 0.0;
        return load_px<SrcT>(src + (size_t)yy * spitch + (size_t)xx * C + ch);
    };
    if (!BILINEAR) {
        double xr = x + 0.5, yr = y + 0.5;
        bool bad = !(fabs(xr) < 2147483648.0 && fabs(yr) < 2147483648.0);   // also catches NaN
        int xi = bad ? 0 : (int)xr, yi = bad ? 0 : (int)yr;                  // truncation toward zero
        if (bad || xi > bw - 1 || xi < 0 || yi > bh - 1 || yi < 0) { xi = 0; yi = 0; }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) dst[ch] = store_cvt<DstT>(tap(yi, xi, ch));
    } else {
        bool bad = !(fabs(x) < 2147483648.0 && fabs(y) < 2147483648.0);
        if (bad || x > (double)(bw - 1) || x < 0.0 || y > (double)(bh - 1) || y < 0.0) { x = 0.0; y = 0.0; }
        const int xi = (int)x, yi = (int)y;
        const double fx = x - (double)xi, fy = y - (double)yi;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            double a = tap(yi, xi, ch), b = tap(yi, xi + 1, ch);
            double cc = tap(yi + 1, xi, ch), d = tap(yi + 1, xi + 1, ch);
            double top = __dadd_rn(__dmul_rn(a, 1.0 - fx), __dmul_rn(b, fx));
            double bot = __dadd_rn(__dmul_rn(cc, 1.0 - fx), __dmul_rn(d, fx));
            dst[ch] = store_cvt<DstT>(__dadd_rn(__dmul_rn(top, 1.0 - fy), __dmul_rn(bot, fy)));
        }
    }
}

// One thread per output pixel; blockIdx.z = job.
template <typename SrcT, typename DstT, int C, int BILINEAR>
__global__ void __launch_bounds__(256)
warp_generic_kernel(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
                    const b200geo_warp_job *__restrict__ jobs)
{
    const b200geo_warp_job &J = jobs[blockIdx.z];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= J.ow || r >= J.oh) return;
    // a job whose rows overlap (pitch smaller than a row) is not executed: the in-bounds arguments below assume iw / ow columns per row
    if (J.src_pitch < (long long)J.iw * C * (long long)sizeof(SrcT) || J.dst_pitch < (long long)J.ow * C * (long long)sizeof(DstT)) return;
    const SrcT *src = reinterpret_cast<const SrcT *>(src_base + J.src_offset);
    const size_t spitch = (size_t)J.src_pitch / sizeof(SrcT);
    DstT *dst = reinterpret_cast<DstT *>(dst_base + J.dst_offset + (size_t)r * (size_t)J.dst_pitch) + (size_t)c * C;

    const double Xo = grid_coord(J.x0, J.dx, J.x_last, c, J.ow);
    const double Yo = grid_coord(J.y0, J.dy, J.y_last, r, J.oh);
    const double zx = fma(J.invH[2], 1.0, fma(J.invH[1], Yo, J.invH[0] * Xo));
    const double zy = fma(J.invH[5], 1.0, fma(J.invH[4], Yo, J.invH[3] * Xo));
    const double zw = fma(J.invH[8], 1.0, fma(J.invH[7], Yo, J.invH[6] * Xo));
    sample_store<SrcT, DstT, C, BILINEAR>(src, spitch, J.iw, J.ih, J.bw, J.bh, zx / zw, zy / zw, dst);
}

// Samplers called on explicit coordinates (nearestNeighbor / bilinear as standalone functions,
// homography.py:108-138): z is the 3 x N float64 array z_t (rows x, y, 1), already divided.
template <typename SrcT, typename DstT, int C, int BILINEAR>
__global__ void __launch_bounds__(256)
sample_kernel(const SrcT *__restrict__ src, size_t spitch, int iw, int ih, int bw, int bh,
              const double *__restrict__ zx, const double *__restrict__ zy, long long N, DstT *__restrict__ dst)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        sample_store<SrcT, DstT, C, BILINEAR>(src, spitch, iw, ih, bw, bh, zx[i], zy[i], dst + i * C);
}

}  // namespace b200geo
