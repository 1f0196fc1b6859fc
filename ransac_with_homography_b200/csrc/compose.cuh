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

#include "solve.cuh"
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

// Batched stitch pipeline (config 5): the warp jobs of B pairs built ON THE DEVICE from the RANSAC results, so that no host
// round trip separates the refit from the warp.  Per pair: transformImageH(frame_b, H_b) -> wrapPerspective's output rectangle
// (homography.py:143-163): the four corners (0,0), (w-1,0), (w-1,h-1), (0,h-1) through H in float64, int() truncation of the
// extremes, and inv(H) (homography.py:172) by the adjugate.  The output goes to canvas slot b, clipped to the slot.
// Pairs whose search failed (status != 0), whose H or projected corners are not finite, or whose rectangle is absurd get the
// identity warp and ok = 0.  results: B x 32 doubles (b200geo_ransac_finish_dev layout); boxes: B x 5 int32
// (min_x, min_y, out_w, out_h, ok).  One thread per pair.
__global__ void __launch_bounds__(128)
stitch_jobs_kernel(const double *__restrict__ results, int B, int ih, int iw, int canvas_h, int canvas_w, long long src_pitch,
                   long long src_stride, long long dst_pitch, long long dst_stride, b200geo_warp_job *__restrict__ jobs,
                   int *__restrict__ boxes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double *r = results + (size_t)b * 32;
    double H[9];
    bool ok = r[2] == 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) { H[j] = r[12 + j]; ok = ok && fabs(H[j]) < 1.0e300; }
    double mnx = 0, mxx = 0, mny = 0, mxy = 0;
    if (ok) {
        const double cxs[4] = {0.0, (double)(iw - 1), (double)(iw - 1), 0.0}, cys[4] = {0.0, 0.0, (double)(ih - 1), (double)(ih - 1)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // (numpy: H @ bnd with float64 dgemm, then the division by row 2)
            const double zx = fma(H[2], 1.0, fma(H[1], cys[k], H[0] * cxs[k]));
            const double zy = fma(H[5], 1.0, fma(H[4], cys[k], H[3] * cxs[k]));
            const double zw = fma(H[8], 1.0, fma(H[7], cys[k], H[6] * cxs[k]));
            const double px = zx / zw, py = zy / zw;
            ok = ok && fabs(px) < 1.0e9 && fabs(py) < 1.0e9;        // also NaN
            if (k == 0) { mnx = mxx = px; mny = mxy = py; }
            else { mnx = fmin(mnx, px); mxx = fmax(mxx, px); mny = fmin(mny, py); mxy = fmax(mxy, py); }
        }
    }
    double Hi[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (ok) {
        invert3<double>(H, Hi);
#pragma unroll
        for (int j = 0; j < 9; ++j) ok = ok && fabs(Hi[j]) < 1.0e300;
    }
    long long min_x = 0, min_y = 0, ow = iw, oh = ih;
    if (ok) {
        min_x = (long long)mnx; min_y = (long long)mny;             // int(): toward zero
        ow = (long long)mxx - min_x + 1; oh = (long long)mxy - min_y + 1;
    } else {
        Hi[0] = Hi[4] = Hi[8] = 1.0; Hi[1] = Hi[2] = Hi[3] = Hi[5] = Hi[6] = Hi[7] = 0.0;
    }
    if (ow > canvas_w) ow = canvas_w;
    if (oh > canvas_h) oh = canvas_h;
    if (ow < 0) ow = 0;
    if (oh < 0) oh = 0;
    b200geo_warp_job J;
#pragma unroll
    for (int j = 0; j < 9; ++j) J.invH[j] = Hi[j];
    J.x0 = (double)min_x; J.dx = 1.0; J.x_last = (double)(min_x + ow - 1);
    J.y0 = (double)min_y; J.dy = 1.0; J.y_last = (double)(min_y + oh - 1);
    J.ow = (int)ow; J.oh = (int)oh; J.bw = iw; J.bh = ih; J.iw = iw; J.ih = ih;
    J.reserved0 = 0; J.reserved1 = 0;
    J.src_offset = (long long)b * src_stride; J.src_pitch = src_pitch;
    J.dst_offset = (long long)b * dst_stride; J.dst_pitch = dst_pitch;
    jobs[b] = J;
    int *bx = boxes + (size_t)b * 5;
    bx[0] = (int)min_x; bx[1] = (int)min_y; bx[2] = (int)ow; bx[3] = (int)oh; bx[4] = ok ? 1 : 0;
}

}  // namespace b200geo
