// compose.cuh -- K5: stitch compose, fused with the warp.
//
// Replaces the tail of stitchPanorama (reference homography.py:288-338): addAlpha('Rate'),
// transformImageH (bbox warp, bilinear, float64), canvas placement, paste / alpha blend and the final
// astype(uint8).  The reference materialises imgT as float32 RGBA, warps it to a float64 RGBA image,
// builds a float32 RGBA canvas and blends with array expressions; here one thread owns one canvas
// pixel, reads the uint8 sources directly and writes the uint8 canvas -- 3 bytes out per pixel, no
// intermediate image.  Element semantics (each rounding as numpy performs it):
//   paste (blending falsy):   canvas = 0 ; canvas[T rect] = uint8(warp(imgT)) ; canvas[Q rect] = imgQ      (:335-338)
//   blend (blending 'Rate'):  Q side  q = float32(imgQ) inside the Q rect else 0, alpha aq = float32(1 + 1e-10 - rate)
//                                     inside, float32(1e-10) outside                                         (:322-329)
//                             T side  (t, at) = float64 bilinear sample of the float32 RGBA image whose alpha plane is
//                                     float32(rate + 1e-10) everywhere; source pixel (0,0) reads 0 in all four channels
//                             base = aq + at ; v = (aq / base) * q + (at / base) * t   (float64, separate roundings)
//                             canvas = uint8(float32(v)) inside the T rect, uint8(q) elsewhere               (:330-334)
#pragma once

#include "warp.cuh"

namespace b200geo {

// Grid: (ceil(fw / 32), ceil(fh / 8)), block (32, 8).
template <int BLEND>
__global__ void __launch_bounds__(256)
stitch_compose_kernel(const uint8_t *__restrict__ imgT, const uint8_t *__restrict__ imgQ, uint8_t *__restrict__ canvas,
                      const b200geo_warp_job J, const b200geo_compose P)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= P.fw || y >= P.fh) return;
    const int tx = x - P.tsx, ty = y - P.tsy, qx = x - P.qsx, qy = y - P.qsy;
    const bool in_t = tx >= 0 && ty >= 0 && tx < J.ow && ty < J.oh;
    const bool in_q = qx >= 0 && qy >= 0 && qx < P.qw && qy < P.qh;
    uint8_t *o = canvas + (size_t)y * (size_t)P.canvas_pitch + (size_t)x * 3;
    uint8_t q8[3] = {0, 0, 0};
    if (in_q) {
        const uint8_t *q = imgQ + (size_t)qy * (size_t)P.q_pitch + (size_t)qx * 3;
        q8[0] = __ldg(q); q8[1] = __ldg(q + 1); q8[2] = __ldg(q + 2);
    }
    if (!BLEND && in_q) { o[0] = q8[0]; o[1] = q8[1]; o[2] = q8[2]; return; }
    if (!in_t) { o[0] = q8[0]; o[1] = q8[1]; o[2] = q8[2]; return; }      // blend: uint8(float32(q)) = q ; paste: 0

    const double Xo = grid_coord(J.x0, J.dx, J.x_last, tx, J.ow);
    const double Yo = grid_coord(J.y0, J.dy, J.y_last, ty, J.oh);
    const double zx = fma(J.invH[2], 1.0, fma(J.invH[1], Yo, J.invH[0] * Xo));
    const double zy = fma(J.invH[5], 1.0, fma(J.invH[4], Yo, J.invH[3] * Xo));
    const double zw = fma(J.invH[8], 1.0, fma(J.invH[7], Yo, J.invH[6] * Xo));
    const double sx = zx / zw, sy = zy / zw;
    const uint8_t *src = imgT + J.src_offset;
    if (!BLEND) {
        sample_store<uint8_t, uint8_t, 3, 1>(src, (size_t)J.src_pitch, J.iw, J.ih, J.bw, J.bh, sx, sy, o);
        return;
    }
    double t[3];
    sample_store<uint8_t, double, 3, 1>(src, (size_t)J.src_pitch, J.iw, J.ih, J.bw, J.bh, sx, sy, t);
    // the alpha plane through the same sampler (bilinear, homography.py:123-138): a constant, except that
    // pixel (0,0) and everything outside the image read 0
    double at;
    {
        double xs = sx, ys = sy;
        const bool bad = !(fabs(xs) < 2147483648.0 && fabs(ys) < 2147483648.0);
        if (bad || xs > (double)(J.bw - 1) || xs < 0.0 || ys > (double)(J.bh - 1) || ys < 0.0) { xs = 0.0; ys = 0.0; }
        const int xi = (int)xs, yi = (int)ys;
        const double fx = xs - (double)xi, fy = ys - (double)yi;
        auto tap = [&](int yy, int xx) -> double {
            return (yy < 0 || xx < 0 || yy >= J.ih || xx >= J.iw || (yy == 0 && xx == 0)) ? 0.0 : P.alpha_t;
        };
        const double top = __dadd_rn(__dmul_rn(tap(yi, xi), 1.0 - fx), __dmul_rn(tap(yi, xi + 1), fx));
        const double bot = __dadd_rn(__dmul_rn(tap(yi + 1, xi), 1.0 - fx), __dmul_rn(tap(yi + 1, xi + 1), fx));
        at = __dadd_rn(__dmul_rn(top, 1.0 - fy), __dmul_rn(bot, fy));
    }
    const double aq = in_q ? P.alpha_q : P.alpha_bg;
    const double base = __dadd_rn(aq, at);
    const double wq = __ddiv_rn(aq, base), wt = __ddiv_rn(at, base);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double v = __dadd_rn(__dmul_rn(wq, (double)q8[ch]), __dmul_rn(wt, t[ch]));
        o[ch] = (uint8_t)(int)__double2float_rn(v);                         // float32 canvas, then astype(np.uint8)
    }
}

}  // namespace b200geo
