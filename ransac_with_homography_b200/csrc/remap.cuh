// remap.cuh -- cv2.remap(img, map_x, map_y, INTER_LINEAR, borderMode=BORDER_TRANSPARENT) for uint8 images: the resampling
// step of the reference's cylindricalWarp (homography.py:391-410; it passes INTER_AREA, which cv::remap treats as INTER_LINEAR).
//
// The algorithm lives in OpenCV (third-party; this image has 4.13.0, modules/imgproc/src/imgwarp.cpp: remap -> convertMaps
// to fixed point -> remapBilinear<FixedPtCast<int, uchar, 15>>):
//     sx = cvRound(map_x * 32), sy = cvRound(map_y * 32)          (float32 product, round half to even)
//     ix = sx >> 5, iy = sy >> 5 (saturated to int16), ax = sx & 31, ay = sy & 31
//     integer position inside the image (0 <= ix < w and 0 <= iy < h; taps beyond the last column / row repeat it):
//         out = (w00*S[iy][ix] + w01*S[iy][ix+1] + w10*S[iy+1][ix] + w11*S[iy+1][ix+1] + 2^14) >> 15
//         with the table weights w00 = (32-ax)(32-ay)*32, w01 = ax(32-ay)*32, w10 = (32-ax)ay*32, w11 = ax*ay*32 (sum 2^15;
//         the table's saturation of 32768 to 32767 at ax = ay = 0 cannot change an 8-bit result)
//     else, BORDER_TRANSPARENT: the destination pixel is left as it is.
// (Behaviour pinned against cv2.remap of OpenCV 4.13 on random maps: tests/golden/ext.npz.)
// One thread per output pixel; the 2 x 2 taps of a 4-channel image are four 32-bit loads.
#pragma once

#include "common.cuh"

namespace b200geo {

__device__ __forceinline__ int remap_fix(float v)
{
    const float s = __fmul_rn(v, 32.f);
    if (!(s > -2147483520.f)) return (int)0x80000000;          // cvRound of NaN / out-of-range: INT_MIN (cvtss2si)
    if (!(s < 2147483520.f)) return (int)0x80000000;
    return __float2int_rn(s);
}

template <int CH>
__global__ void __launch_bounds__(256)
remap_linear_kernel(const uint8_t *__restrict__ src, int ih, int iw, long long src_pitch, const float *__restrict__ mx,
                    const float *__restrict__ my, int oh, int ow, uint8_t *__restrict__ dst, long long dst_pitch)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= ow || r >= oh) return;
    const size_t o = (size_t)r * ow + c;
    const int sx = remap_fix(mx[o]), sy = remap_fix(my[o]);
    int ix = sx >> 5, iy = sy >> 5;
    ix = max(-32768, min(32767, ix));                          // saturate_cast<short>
    iy = max(-32768, min(32767, iy));
    if (!((unsigned)ix < (unsigned)iw && (unsigned)iy < (unsigned)ih)) return;      // transparent: untouched
    const int ax = sx & 31, ay = sy & 31;
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    const uint8_t *p0 = src + (long long)iy * src_pitch + (long long)ix * CH, *p1 = iy + 1 < ih ? p0 + src_pitch : p0;
    const int nx = ix + 1 < iw ? CH : 0;
    uint8_t *d = dst + (long long)r * dst_pitch + (long long)c * CH;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int t = w00 * p0[k] + w01 * p0[nx + k] + w10 * p1[k] + w11 * p1[nx + k];
        d[k] = (uint8_t)min(255, max(0, (t + (1 << 14)) >> 15));
    }
}

}  // namespace b200geo
