// warp_fast.cuh -- K4 fast path: backward warp of uint8 RGB images to uint8 RGB, bilinear or nn.
//
// Same element semantics as warp.cuh (reference homography.py:108-209 + astype(uint8) of :225/:241),
// organised for throughput.  The output is cut into 64 x 40 tiles (WF_TW x WF_TH).
//
// Set-up (one THREAD per tile, float64; wf_setup): re-expresses the projective map of the tile around
//   the tile's centre pixel and an integer anchor (xc, yc) next to the centre's source position,
//       rel_x = x - xc = (nx0 + ax*di + bx*dj) / (1 + gw*di + hw*dj)        (di, dj = pixel - centre)
//   rounds the nine coefficients to float32 once, bounds the float32 evaluation error of rel
//   (`band`), maps the four tile corners and classifies the tile: ZERO (wholly masked), INTERIOR
//   (wholly unmasked), BORDER (per-pixel mask), SPLIT (footprint too large: redone per quadrant) or
//   SLOW (the bound does not hold: w changes too much, irregular grid, NaN -> float64 per pixel).
//   warp_tile_setup_kernel writes a complete 192-byte record per tile for the persistent kernel,
//   warp_setup_kernel the 64-byte descriptor alone for the one-CTA-per-tile comparison kernel.
//
// Per tile (warp_pipe_kernel: persistent, warp-specialised, see the comment above it; warp_fast_kernel:
// one CTA per tile, kept for comparison):
//   1. staging: the source footprint (bounding box of the four mapped corners -- the image of a
//      convex quad under a projective map with w > 0 is convex) arrives in shared memory by one TMA
//      tensor copy (or per-row bulk copies / plain loads) and is re-laid-out as one paired entry per
//      source pixel x: plane 0 (R[x] R[x+1] G[x] G[x+1]), plane 1 (B[x] B[x+1] . .), i.e. horizontally
//      adjacent taps already paired for the dot-product instruction.  Pixels outside the image and
//      pixel (0,0) are staged as 0, which is exactly how the reference's samplers see them.
//   2. sampling: every float32 term is tile-local (|di| <= 32, |rel| <= a few hundred), so rel has
//      an absolute error of ~1e-5..1e-4 px although the image is thousands of pixels wide.  Per
//      pixel: FFMA + FFMA2 + MUFU.RCP + FMUL2 + 2 F2I give 16.16 fixed-point coordinates, four LDS.32
//      fetch the taps, and the blend is integer arithmetic with 16-bit weights that sum to 65535
//      plus a bias of 255 (IDP.2A, 2 per channel; see wf_blend): flat regions and integer / dyadic
//      translations are reproduced exactly and byte 2 of the accumulator is the reference's
//      truncating astype(uint8).
//   3. store: full tiles write 96 contiguous bytes per warp and store instruction straight to global
//      memory (SHFL + PRMT pack); other tiles go through a swizzled shared-memory output stage.
//
// Parity: mask decisions (bilinear: x > bw-1 | x < 0 ...; nn: the rounded index) are taken on the
// float32 coordinates only when they are further than `band` from the decision boundary; pixels
// inside the band, and every pixel of a SLOW tile, are evaluated by sample_store() in float64
// exactly as warp_generic_kernel does (deferred to warp_fixup_kernel).  nn output is therefore
// bit-identical to the reference; bilinear output differs from the reference's float64 blend by
// < 0.05 grey levels before the truncation, i.e. within +-1 after it (north_star tolerance),
// identical on flat regions and for integer / dyadic translations.
#pragma once

#include "warp.cuh"

namespace b200geo {

#ifndef B2G_WP_OCC
#define B2G_WP_OCC 2
#endif
#ifndef B2G_UNROLL_P
#define B2G_UNROLL_P 4
#endif
constexpr int WF_UNROLL_P = B2G_UNROLL_P;   // rows of a thread's pixels unrolled together in the full-tile loop
constexpr int WF_TW = 64;             // output tile width  (pixels)
#ifndef B2G_TH
#define B2G_TH 40       // 1080 = 27 x 40: no partial tile row at the benchmark's frame height (32 and 48 measured slower there)
#endif
#ifndef B2G_ENTRIES
#define B2G_ENTRIES 4096
#endif
#ifndef B2G_BOX_H
#define B2G_BOX_H 48
#endif
constexpr int WF_TH = B2G_TH;         // output tile height (multiple of 8)
static_assert(B2G_TH % 8 == 0 && B2G_TH >= 16 && B2G_TH <= 96, "tile height");
constexpr int WF_THREADS = 256;
constexpr int WF_SRC_ENTRIES = B2G_ENTRIES;  // staged source tile capacity: two planes of 16 KB (one 32-bit word per entry and plane)
constexpr uint32_t WF_PLANE = WF_SRC_ENTRIES * 4u;   // byte offset of plane 1 (B pairs) from plane 0 (R/G pairs)
constexpr float WF_W0 = 1.0f / 65536.f;   // constant term of the scaled denominator
constexpr float WF_MAX_REL = 1024.f;  // |rel| beyond this: SLOW (keeps the 16.16 conversion and the bound sane)
constexpr float WF_MAX_BAND = 1.0f / 64.f;

enum { WF_EMPTY = 0, WF_ZERO = 1, WF_INTERIOR = 2, WF_BORDER = 3, WF_SLOW = 4, WF_SPLIT = 5, WF_DONE = 6 /* pipeline sentinel, never in a descriptor */ };
constexpr int WF_SUB_W = 32, WF_SUB_H = 16;   // SPLIT: the tile is redone as quadrants, each with its own (smaller) footprint
constexpr int WF_TMA_BOX_W = 256;          // bytes per row of the 2-D TMA box (the hardware maximum for one dimension)
constexpr int WF_TMA_BOX_H = B2G_BOX_H;    // rows of the box: 12 KB per tile at 48
constexpr int WF_FLAG_TMA = 0x400;         // footprint fits the TMA box and the job lies in the described 2-D source array
constexpr int WF_FLAG_UNCHECKED = 0x100;   // footprint (+ read slack) lies inside the image and excludes pixel (0,0)

struct __align__(16) WfDesc {
    float nx0, ax, bx, ny0, ay, by, gw, hw;       // gw, hw pre-scaled by 2^-16 (w0 = WF_W0): nx / w = rel * 65536
    int srow;                                     // TMA tiles: row of the footprint's first row in the 2-D source array
    float band;                                   // ambiguity half-width, in rel * 65536 units
    int xc, yc;                                   // anchor (absolute source pixel)
    short ox, oy;                                 // staged tile origin, relative to the anchor
    unsigned short P, rows;                       // staged tile pitch (entries, multiple of 4) and rows
    int mode;                                     // WF_* | flags
    unsigned magic;                               // ceil(2^32 / (P / 4)): division by the chunk count
};
static_assert(sizeof(WfDesc) == 64, "WfDesc is read as four 16-byte words");

// Special registers read through volatile asm: the persistent loop is register-bound, and values
// derived from %tid / %nctaid that the compiler hoists out of it end up spilled to local memory
// (an LDL per tile on the critical path); re-deriving them per tile is a few ALU instructions.
__device__ __forceinline__ int wf_tid()
{
    int v;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(v));
    return v;
}
__device__ __forceinline__ int wf_nctaid()
{
    int v;
    asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(v));
    return v;
}

__device__ __forceinline__ float wf_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// rel * 65536 of output pixel (centre + di, centre + dj)
__device__ __forceinline__ void wf_rel16(const WfDesc &T, float di, float dj, float &rx, float &ry)
{
    const float nx = __fmaf_rn(T.ax, di, __fmaf_rn(T.bx, dj, T.nx0));
    const float ny = __fmaf_rn(T.ay, di, __fmaf_rn(T.by, dj, T.ny0));
    const float w = __fmaf_rn(T.gw, di, __fmaf_rn(T.hw, dj, WF_W0));
    const float rw = wf_rcp(w);
    rx = __fmul_rn(nx, rw);
    ry = __fmul_rn(ny, rw);
}

// unmasked range of rel * 65536 for a tile anchored at (xc, yc)
template <int BILINEAR>
__device__ __forceinline__ void wf_bounds(int xc, int yc, int bw, int bh, float &lox, float &hix, float &loy, float &hiy)
{
    if (BILINEAR) {   // x > bw-1 | x < 0 | y > bh-1 | y < 0 -> masked (homography.py:126-130)
        lox = (float)(0 - xc) * 65536.f; hix = (float)(bw - 1 - xc) * 65536.f;
        loy = (float)(0 - yc) * 65536.f; hiy = (float)(bh - 1 - yc) * 65536.f;
    } else {          // trunc(x + 0.5) in [0, bw-1]  <=>  x in (-1.5, bw - 0.5)   (homography.py:109-116)
        lox = ((float)(0 - xc) - 1.5f) * 65536.f; hix = ((float)(bw - xc) - 0.5f) * 65536.f;
        loy = ((float)(0 - yc) - 1.5f) * 65536.f; hiy = ((float)(bh - yc) - 0.5f) * 65536.f;
    }
}

// How the source array is described to the TMA unit (b200geo_src_layout): pitch = 0: not at all; image_rows = 0:
// one 2-D byte array (a box must lie inside the image: rows of the neighbouring image are not zeros);
// image_rows > 0: a 3-D array (row_bytes x image_rows x images) -- a box may hang over any image edge, the
// hardware fills the out-of-bounds part with zeros, which is how the samplers see it.
struct WfTma {
    long long pitch;
    int image_rows, row_bytes;
};

// jobs that arrive in device memory cannot be checked by the host: a job whose pitches are smaller than its rows is skipped
__device__ __forceinline__ bool wf_job_ok(const b200geo_warp_job &J)
{
    return J.src_pitch >= 3LL * J.iw && J.dst_pitch >= 3LL * J.ow && J.iw > 0 && J.ih > 0;
}

template <int BILINEAR>
__device__ void wf_setup(const b200geo_warp_job &J, int c0, int r0, int tw, int th, const WfTma &tma, WfDesc &T)
{
    T.mode = WF_SLOW;
    T.P = 0; T.rows = 0; T.ox = 0; T.oy = 0; T.xc = 0; T.yc = 0; T.band = 0.f; T.magic = 0u;
    T.nx0 = T.ax = T.bx = T.ny0 = T.ay = T.by = T.gw = T.hw = 0.f; T.srow = 0;
    const int cm = c0 + WF_TW / 2, rm = r0 + WF_TH / 2;
    const double Xm = (double)cm * J.dx + J.x0, Ym = (double)rm * J.dy + J.y0;
    // the grid must be the arithmetic progression the tile-affine form assumes (np.linspace's
    // forced last sample differs from it by an ulp; anything else goes the exact way)
    const double exl = (double)(J.ow - 1) * J.dx + J.x0, eyl = (double)(J.oh - 1) * J.dy + J.y0;
    bool ok = fabs(exl - J.x_last) <= 1e-9 * fmax(1.0, fabs(J.x_last)) && fabs(eyl - J.y_last) <= 1e-9 * fmax(1.0, fabs(J.y_last));
    ok = ok && J.bw >= 2 && J.bh >= 2 && J.bw < (1 << 22) && J.bh < (1 << 22) && J.iw >= 1 && J.ih >= 1;
    const double *h = J.invH;
    const double zx = h[0] * Xm + h[1] * Ym + h[2];
    const double zy = h[3] * Xm + h[4] * Ym + h[5];
    const double zw = h[6] * Xm + h[7] * Ym + h[8];
    ok = ok && fabs(zw) > 1e-200 && fabs(zw) < 1e200;
    if (!ok) return;
    const double s = 1.0 / zw;
    const double xs = zx * s, ys = zy * s;
    if (!(fabs(xs) < 4.0e6 && fabs(ys) < 4.0e6)) return;          // also NaN
    const double xc = rint(xs), yc = rint(ys);
    const double ax = (h[0] - xc * h[6]) * s * J.dx, bx = (h[1] - xc * h[7]) * s * J.dy;
    const double ay = (h[3] - yc * h[6]) * s * J.dx, by = (h[4] - yc * h[7]) * s * J.dy;
    const double gw = h[6] * s * J.dx, hw = h[7] * s * J.dy;
    const double q = fabs(gw) * (WF_TW / 2) + fabs(hw) * (WF_TH / 2);   // |w - 1| over the tile
    if (!(q <= 0.25)) return;
    const double Sx = fabs(xs - xc) + fabs(ax) * (WF_TW / 2) + fabs(bx) * (WF_TH / 2);
    const double Sy = fabs(ys - yc) + fabs(ay) * (WF_TW / 2) + fabs(by) * (WF_TH / 2);
    if (!(Sx < 1.0e6 && Sy < 1.0e6)) return;
    T.nx0 = (float)(xs - xc); T.ax = (float)ax; T.bx = (float)bx;
    T.ny0 = (float)(ys - yc); T.ay = (float)ay; T.by = (float)by;
    const double sc = 1.0 / 65536.0;
    T.gw = (float)(gw * sc); T.hw = (float)(hw * sc);
    T.xc = (int)xc; T.yc = (int)yc;

    // four corners of the tile through the float32 formula the pixels use
    float rxmin = 3.0e38f, rxmax = -3.0e38f, rymin = 3.0e38f, rymax = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float di = (float)(((k & 1) ? tw - 1 : 0) - WF_TW / 2), dj = (float)(((k & 2) ? th - 1 : 0) - WF_TH / 2);
        float rx, ry;
        wf_rel16(T, di, dj, rx, ry);
        rxmin = fminf(rxmin, rx); rxmax = fmaxf(rxmax, rx);
        rymin = fminf(rymin, ry); rymax = fmaxf(rymax, ry);
    }
    const float R = fmaxf(fmaxf(fabsf(rxmin), fabsf(rxmax)), fmaxf(fabsf(rymin), fabsf(rymax))) * (1.0f / 65536.f);
    if (!(R < WF_MAX_REL)) return;
    // float32 error of rel (DESIGN.md "warp: coordinate error bound") with a 3.5x margin, plus the float64-side slack
    const float band = 2.4e-7f * (6.f * (float)fmax(Sx, Sy) + 10.f * R) + 4.0e-6f;
    if (!(band <= WF_MAX_BAND)) return;
    const float B = band * 65536.f;
    T.band = B;
    float lox, hix, loy, hiy;
    wf_bounds<BILINEAR>(T.xc, T.yc, J.bw, J.bh, lox, hix, loy, hiy);
    // nn: an INTERIOR tile must also keep the rounded index non-negative (x > -0.5), so that the
    // fast loop needs no trunc-toward-zero clamp
    const float nn_lo = BILINEAR ? 0.f : 65536.f;
    const bool all_in = rxmin >= lox + nn_lo + B && rxmax <= hix - B && rymin >= loy + nn_lo + B && rymax <= hiy - B;
    const bool all_out = rxmax < lox - B || rxmin > hix + B || rymax < loy - B || rymin > hiy + B;
    if (all_out) { T.mode = WF_ZERO; return; }
    // index range of the entries the pixels can address (relative to the anchor); an entry holds x and x+1
    int x_lo, x_hi, y_lo, y_hi;
    if (BILINEAR) {
        x_lo = __float2int_rd((rxmin - B) * (1.f / 65536.f)); x_hi = __float2int_rd((rxmax + B) * (1.f / 65536.f));
        y_lo = __float2int_rd((rymin - B) * (1.f / 65536.f)); y_hi = __float2int_rd((rymax + B) * (1.f / 65536.f)) + 1;
    } else {
        x_lo = __float2int_rd((rxmin - B) * (1.f / 65536.f) + 0.5f) - 1; x_hi = __float2int_rd((rxmax + B) * (1.f / 65536.f) + 0.5f) + 1;
        y_lo = __float2int_rd((rymin - B) * (1.f / 65536.f) + 0.5f) - 1; y_hi = __float2int_rd((rymax + B) * (1.f / 65536.f) + 0.5f) + 1;
    }
    if (!all_in) {
        // Clip to what an unmasked pixel can touch.  The unrolled BORDER loops clamp every coordinate into the image
        // before addressing, so the staged range must contain the clamped position of EVERY pixel of the tile: an axis
        // whose range misses the image altogether (the tile is not ZERO only because of `band`) collapses onto the
        // image edge the clamp produces (+ the second tap row for bilinear) instead of onto an arbitrary entry.
        const int ix_lo = 0 - T.xc, ix_hi = J.bw - 1 - T.xc, iy_lo = 0 - T.yc, iy_hi = J.bh - 1 - T.yc, ytap = BILINEAR ? 1 : 0;
        if (x_hi < ix_lo) x_lo = x_hi = ix_lo;
        else if (x_lo > ix_hi) x_lo = x_hi = ix_hi;
        else { x_lo = max(x_lo, ix_lo); x_hi = min(x_hi, ix_hi); }
        if (y_hi < iy_lo) { y_lo = iy_lo; y_hi = iy_lo + ytap; }
        else if (y_lo > iy_hi) { y_lo = iy_hi; y_hi = iy_hi + ytap; }
        else { y_lo = max(y_lo, iy_lo); y_hi = min(y_hi, iy_hi + ytap); }
    }
    if (x_hi < x_lo || y_hi < y_lo) { x_hi = x_lo; y_hi = y_lo; }
    // absolute origin column: a multiple of 4 pixels = 12 bytes, so that every 4-entry chunk of a
    // row starts on a 32-bit boundary relative to the row start
    const long long tma_pitch = tma.pitch;
    const bool tma_job = tma_pitch > 0 && J.src_pitch == tma_pitch && J.src_offset >= 0 && J.src_offset % tma_pitch == 0;
    const int ax0 = (T.xc + x_lo) & ~3;
    const int ox = ax0 - T.xc;
    const long long width = (long long)x_hi - ox + 1, rows = (long long)y_hi - y_lo + 1;
    const long long P = max(8LL, (width + 3) & ~3LL);       // >= 2 chunks: the conversion's division by multiplication needs chunks > 1
    if (ox < -32000 || y_lo < -32000) return;
    if (P * rows > WF_SRC_ENTRIES) { T.mode = WF_SPLIT; return; }        // footprint does not fit: the kernel retries per quadrant
    T.ox = (short)ox;
    T.oy = (short)y_lo;
    T.P = (unsigned short)P;
    T.rows = (unsigned short)rows;
    const unsigned chunks = (unsigned)(P >> 2);
    T.magic = chunks > 1 ? (unsigned)((0x100000000ULL + chunks - 1) / chunks) : 0u;
    T.mode = all_in ? WF_INTERIOR : WF_BORDER;
    // unchecked staging: every 16-byte read of the staging loop stays inside the image rows
    const int ay0 = T.yc + y_lo;
    const bool inside = ax0 >= 0 && (long long)ax0 + P + 3 <= J.iw && ay0 >= 0 && (long long)ay0 + rows <= J.ih && !(ax0 == 0 && ay0 == 0);
    if (inside) T.mode |= WF_FLAG_UNCHECKED;
    // a TMA box must start on a 16-byte boundary of the global array (measured: an unaligned innermost
    // coordinate faults): the box starts at (3*ax0) & ~15 and the rows carry a phase of 0/4/8/12 bytes
    const bool box_ok = tma_job && ((3 * ax0) & 15) + 3 * (P + 1) <= WF_TMA_BOX_W && rows <= WF_TMA_BOX_H;
    if (tma.image_rows <= 0) {
        if (inside && box_ok) {
            T.mode |= WF_FLAG_TMA;
            T.srow = (int)(J.src_offset / tma_pitch + ay0);
        }
    } else {
        // 3-D description: edge footprints too, except the one holding source pixel (0,0) (it must read as 0)
        const bool has00 = ax0 <= 0 && (long long)ax0 + P + 1 > 0 && ay0 <= 0 && (long long)ay0 + rows > 0;
        const bool img_ok = J.ih == tma.image_rows && 3LL * J.iw == tma.row_bytes && J.src_offset % (tma_pitch * tma.image_rows) == 0;
        if (box_ok && img_ok && !has00 && ax0 > -(1 << 24) && ay0 > -(1 << 24)) {
            T.mode |= WF_FLAG_TMA;
            T.srow = ay0;                                 // row inside its image; the producer adds the image index
        }
    }
}

// Grid: (ceil(tiles_x / 32), ceil(tiles_y / 8), n_jobs), block (32, 8): one thread per tile.
template <int BILINEAR>
__global__ void __launch_bounds__(256)
warp_setup_kernel(const b200geo_warp_job *__restrict__ jobs, const uint8_t *__restrict__ src_base, const WfTma tma,
                  int tiles_x, int tiles_y, WfDesc *__restrict__ descs)
{
    const int tx = blockIdx.x * 32 + threadIdx.x, ty = blockIdx.y * 8 + threadIdx.y;
    if (tx >= tiles_x || ty >= tiles_y) return;
    const b200geo_warp_job &J = jobs[blockIdx.z];
    WfDesc T;
    const int c0 = tx * WF_TW, r0 = ty * WF_TH;
    if (c0 >= J.ow || r0 >= J.oh || !wf_job_ok(J)) {
        T.mode = WF_EMPTY;
        T.P = 0; T.rows = 0; T.ox = 0; T.oy = 0; T.xc = 0; T.yc = 0; T.band = 0.f; T.magic = 0u;
        T.nx0 = T.ax = T.bx = T.ny0 = T.ay = T.by = T.gw = T.hw = 0.f; T.srow = 0;
    } else {
        wf_setup<BILINEAR>(J, c0, r0, min(WF_TW, J.ow - c0), min(WF_TH, J.oh - r0), tma, T);
    }
    const uint4 *s = reinterpret_cast<const uint4 *>(&T);
    uint4 *d = reinterpret_cast<uint4 *>(descs + ((size_t)blockIdx.z * tiles_y + ty) * tiles_x + tx);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

// ------------------------------------------------------------------------------------------------
// shared-memory helpers (explicit .shared addressing keeps the address arithmetic 32-bit)
// ------------------------------------------------------------------------------------------------
// entry e of a staged tile: plane 0 word e = (R[x] R[x+1] G[x] G[x+1]), plane 1 word e = (B[x] B[x+1] . .).
// Two planes (instead of one 8-byte record) make a lane's four entries one contiguous 16-byte store
// per plane -- conflict-free -- at the price of two 32-bit loads per tap pair instead of one 64-bit.
__device__ __forceinline__ uint2 wf_lds_entry(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v.x) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v.y) : "r"(a), "n"(WF_PLANE));
    return v;
}
__device__ __forceinline__ void wf_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void wf_sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// word index of output pixel (row, col) in the swizzled output stage (copy-out reads it with
// conflict-free 128-bit loads)
__device__ __forceinline__ int wf_out_index(int row, int col)
{
    const int u = col >> 2;
    return row * WF_TW + (((u ^ ((u >> 3) & 1) ^ ((row & 1) << 1)) << 2) | (col & 3));
}

// 16 raw bytes w0..w3 = pixels x..x+4 (15 bytes used) -> the chunk's four entries, one 16-byte store per plane
__device__ __forceinline__ void wf_store_chunk(uint32_t d, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    uint4 rg, bb;
    rg.x = __byte_perm(w0, w1, 0x4130); bb.x = __byte_perm(w0, w1, 0x5252);
    rg.y = __byte_perm(w0, w1, 0x7463); bb.y = __byte_perm(w1, w2, 0x4141);
    rg.z = __byte_perm(w1, w2, 0x6352); bb.z = __byte_perm(w2, w2, 0x3030);
    rg.w = __byte_perm(w2, w3, 0x5241); bb.w = __byte_perm(w2, w3, 0x6363);
    wf_sts128(d, rg);
    wf_sts128(d + WF_PLANE, bb);
}

// one source pixel as the reference's samplers see it: 0 outside the image and at (0,0)
__device__ __forceinline__ uint32_t wf_src_px(const uint8_t *__restrict__ src, size_t spitch, int iw, int ih, int y, int x)
{
    if (y < 0 || y >= ih || x < 0 || x >= iw || (x == 0 && y == 0)) return 0u;
    const uint8_t *p = src + (size_t)y * spitch + 3 * (size_t)x;
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
}

// Copy the source footprint into shared memory as paired entries.
__device__ __forceinline__ void wf_stage(const WfDesc &T, const uint8_t *__restrict__ src, size_t spitch, int iw, int ih,
                                         uint32_t s_src)
{
    const int P = T.P, chunks = P >> 2, total = (int)T.rows * chunks;
    const int ay0 = T.yc + T.oy, ax0 = T.xc + T.ox;
    if (T.mode & WF_FLAG_UNCHECKED) {
        const uint8_t *base = src + (size_t)ay0 * spitch + 3 * (size_t)ax0;
        const bool aligned = ((((size_t)base) | spitch) & 3) == 0;
        const uint32_t pitch32 = (uint32_t)spitch;
        for (int t = threadIdx.x; t < total; t += WF_THREADS) {
            const int i = chunks > 1 ? (int)__umulhi((unsigned)t, T.magic) : t;
            const int k = t - i * chunks;
            const uint8_t *a = base + ((uint32_t)i * pitch32 + 12u * (uint32_t)k);
            uint32_t w0, w1, w2, w3;
            if (aligned) {
                const uint32_t *a4 = reinterpret_cast<const uint32_t *>(a);
                w0 = __ldg(a4); w1 = __ldg(a4 + 1); w2 = __ldg(a4 + 2); w3 = __ldg(a4 + 3);
            } else {
                const uint32_t sh = ((uint32_t)(size_t)a & 3u) * 8u;
                const uint32_t *a4 = reinterpret_cast<const uint32_t *>((size_t)a & ~(size_t)3);
                const uint32_t v0 = __ldg(a4), v1 = __ldg(a4 + 1), v2 = __ldg(a4 + 2), v3 = __ldg(a4 + 3);
                const uint32_t v4 = sh ? __ldg(a4 + 4) : 0u;
                w0 = __funnelshift_r(v0, v1, sh); w1 = __funnelshift_r(v1, v2, sh);
                w2 = __funnelshift_r(v2, v3, sh); w3 = __funnelshift_r(v3, v4, sh);
            }
            wf_store_chunk(s_src + (uint32_t)(i * P + 4 * k) * 4u, w0, w1, w2, w3);
        }
    } else {
        for (int t = threadIdx.x; t < total; t += WF_THREADS) {
            const int i = chunks > 1 ? (int)__umulhi((unsigned)t, T.magic) : t;
            const int k = t - i * chunks;
            const int sy = ay0 + i, sx = ax0 + 4 * k;
            uint32_t px[5];
#pragma unroll
            for (int e = 0; e < 5; ++e) px[e] = wf_src_px(src, spitch, iw, ih, sy, sx + e);
            uint4 rg, bb;
            rg.x = __byte_perm(px[0], px[1], 0x5140); bb.x = __byte_perm(px[0], px[1], 0x7362);
            rg.y = __byte_perm(px[1], px[2], 0x5140); bb.y = __byte_perm(px[1], px[2], 0x7362);
            rg.z = __byte_perm(px[2], px[3], 0x5140); bb.z = __byte_perm(px[2], px[3], 0x7362);
            rg.w = __byte_perm(px[3], px[4], 0x5140); bb.w = __byte_perm(px[3], px[4], 0x7362);
            const uint32_t d = s_src + (uint32_t)(i * P + 4 * k) * 4u;
            wf_sts128(d, rg);
            wf_sts128(d + WF_PLANE, bb);
        }
    }
}

// one pixel of a SLOW tile / inside the ambiguity band: float64, exactly as warp_generic_kernel
template <int BILINEAR>
__device__ __noinline__ uint32_t wf_exact_px(const b200geo_warp_job &J, const uint8_t *__restrict__ src, int c, int r)
{
    const double Xo = grid_coord(J.x0, J.dx, J.x_last, c, J.ow);
    const double Yo = grid_coord(J.y0, J.dy, J.y_last, r, J.oh);
    const double zx = fma(J.invH[2], 1.0, fma(J.invH[1], Yo, J.invH[0] * Xo));
    const double zy = fma(J.invH[5], 1.0, fma(J.invH[4], Yo, J.invH[3] * Xo));
    const double zw = fma(J.invH[8], 1.0, fma(J.invH[7], Yo, J.invH[6] * Xo));
    uint8_t o[3];
    sample_store<uint8_t, uint8_t, 3, BILINEAR>(src, (size_t)J.src_pitch, J.iw, J.ih, J.bw, J.bh, zx / zw, zy / zw, o);
    return (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16);
}

// Pixels that need the float64 path are rare (within `band` of a mask / rounding boundary) but slow
// (float64 divisions, dependent global loads): settling one inline stalls its warp -- and at the next
// tile barrier its whole CTA -- for microseconds.  They are pushed to a global list instead and
// settled by warp_fixup_kernel after the main kernel (stream order); the main kernel stores its
// speculative value, the fix-up overwrites the three bytes.
struct WfFixEntry { int job, c, r; };
struct WfFix {
    WfFixEntry *list;
    unsigned *count;
    unsigned cap;
    int job;
};
__device__ __forceinline__ bool wf_defer(const WfFix &f, int c, int r)
{
    if (!f.list) return false;
    const unsigned i = atomicAdd(f.count, 1u);
    if (i >= f.cap) return false;          // list full: the caller settles the pixel inline
    f.list[i].job = f.job; f.list[i].c = c; f.list[i].r = r;
    return true;
}

// The list is cut into n_seg segments of seg_cap entries, segment s holding min(counts[s], seg_cap) of
// them: the persistent kernel gives every CTA its own segment and counts in shared memory (one global
// counter made every deferred pixel wait for a returning atomic on one address -- measured: nn 25 %
// slower than bilinear); the one-CTA-per-tile kernel uses a single segment with a global counter.
// Grid: (x: any, y: <= n_seg); one thread per deferred pixel.  stats[0] += pixels settled.
template <int BILINEAR>
__global__ void __launch_bounds__(128)
warp_fixup_kernel(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base, const b200geo_warp_job *__restrict__ jobs,
                  const WfFixEntry *__restrict__ list, const unsigned *__restrict__ counts, unsigned n_seg, unsigned seg_cap,
                  unsigned long long *__restrict__ stats)
{
    unsigned long long settled = 0;
    for (unsigned sgm = blockIdx.y; sgm < n_seg; sgm += gridDim.y) {
        const unsigned n = min(counts[sgm], seg_cap);
        const WfFixEntry *seg = list + (size_t)sgm * seg_cap;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const WfFixEntry e = seg[i];
            const b200geo_warp_job &J = jobs[e.job];
            const uint32_t v = wf_exact_px<BILINEAR>(J, src_base + J.src_offset, e.c, e.r);
            uint8_t *q = dst_base + J.dst_offset + (size_t)e.r * (size_t)J.dst_pitch + (size_t)e.c * 3;
            q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) settled += n;
    }
    if (stats && settled) atomicAdd(stats, settled);
}

// bilinear blend of two staged entry pairs with 16-bit weights.  The four weights sum to 65535 (not
// 65536) so that every one of them fits its 16-bit field for every (fx, fy) -- no special case for
// fx = fy = 0 -- and the accumulators start at 255:  v = 65535 * value + 255 = 65536 * value + (255 - value),
// whose byte 2 is floor(value) whenever `value` is an integer (flat regions, integer and dyadic
// translations stay exact) and floor(value + eps), 0 <= eps < 0.004, otherwise.
//   w11 = ceil(fx * fy / 65536) keeps w00 = 65535 - fx - fy + w11 >= 0:  (N - fx)(N - fy) >= 0 with N = 65535
//   gives fx * fy / N >= fx + fy - N, and fx * fy / 65536 > fx * fy / N - 1.
__device__ __forceinline__ uint32_t wf_blend(uint2 top, uint2 bot, uint32_t fx, uint32_t fy)
{
    const uint32_t w11 = (fx * fy + 65535u) >> 16;
    const uint32_t Wbot = w11 * 65535u + fy;                       // (fy - w11) | w11 << 16
    const uint32_t Wtop = fx * 65535u + 65535u - Wbot;             // (65535 - fx - fy + w11) | (fx - w11) << 16
    const uint32_t vr = __dp2a_lo(Wtop, top.x, __dp2a_lo(Wbot, bot.x, 255u));
    const uint32_t vg = __dp2a_hi(Wtop, top.x, __dp2a_hi(Wbot, bot.x, 255u));
    const uint32_t vb = __dp2a_lo(Wtop, top.y, __dp2a_lo(Wbot, bot.y, 255u));
    return __byte_perm(__byte_perm(vr, vg, 0x4462), vb, 0x4610);
}

// Full INTERIOR tile: no bounds, no mask; everything unrolled (8 pixels per thread).
// DIRECT = 1: results go straight to global memory.  Four neighbouring lanes hold four RGBX pixels =
// 12 output bytes = three 32-bit words: each lane fetches its right neighbour's pixel (one SHFL),
// builds one word with a lane-dependent PRMT, and three of every four lanes store it -- 96
// contiguous bytes per warp and store instruction.  Requires dst_tile and dst_pitch 4-byte aligned.
// DIRECT = 0: results go to the swizzled shared-memory stage `s_out` (copied out by wf_copy_out).
template <int BILINEAR, int DIRECT>
__device__ __forceinline__ void wf_compute_full(const WfDesc &T, const b200geo_warp_job &J, const uint8_t *__restrict__ src,
                                                uint32_t s_src, uint32_t s_out, int c0, int r0, unsigned &n_exact, const WfFix &fix,
                                                uint8_t *__restrict__ dst_tile = nullptr, size_t dst_pitch = 0)
{
    const int tid_ = wf_tid();
    const int lane = tid_ & 31, wrp = tid_ >> 5;
    const float di0 = (float)(lane - WF_TW / 2), dj0 = (float)(wrp - WF_TH / 2);
    // coefficients live in registers; the two quotients of a pixel are one FMUL2
    const float ax = T.ax, ay = T.ay, bx = T.bx, by = T.by, gw = T.gw, hw = T.hw;
    // (nx, ny) travel as one register pair: each step is one packed FFMA2 (same roundings as two FFMAs)
    const float2 A2 = make_float2(ax, ay), B2 = make_float2(bx, by);
    const float2 nb = __ffma2_rn(A2, make_float2(di0, di0), __ffma2_rn(B2, make_float2(dj0, dj0), make_float2(T.nx0, T.ny0)));
    const float wb = __fmaf_rn(gw, di0, __fmaf_rn(hw, dj0, WF_W0));
    const uint32_t P4 = (uint32_t)T.P * 4u;
    const uint32_t base = s_src - (uint32_t)(((int)T.oy * (int)T.P + (int)T.ox) * 4);
    uint32_t oaddr[WF_TW / 32];
#pragma unroll
    for (int hh = 0; hh < WF_TW / 32; ++hh) {
        oaddr[hh] = s_out + 4u * (uint32_t)wf_out_index(wrp, lane + 32 * hh);
        asm volatile("" : "+r"(oaddr[hh]));      // opaque: keep it in a register instead of re-deriving the swizzle per pixel
    }
    // DIRECT: word (lane & 3) of the 12 bytes of pixels 4*(lane>>2) .. +3:  RGBR | GBRG | BRGB
    const uint32_t psel = (lane & 3) == 0 ? 0x4210u : ((lane & 3) == 1 ? 0x5421u : 0x6542u);
    uint8_t *gptr = dst_tile + (size_t)wrp * dst_pitch + (size_t)(12 * (lane >> 2) + 4 * (lane & 3));
    const int B16 = (int)T.band + 2;
    constexpr int NPX = (WF_TH / 8) * (WF_TW / 32);
    uint32_t outv[NPX];          // nn only
    uint32_t amb = 0u;           // nn only: pixels to settle in float64
#pragma unroll WF_UNROLL_P
    for (int p = 0; p < WF_TH / 8; ++p) {
        const float2 np_ = p ? __ffma2_rn(B2, make_float2((float)(8 * p), (float)(8 * p)), nb) : nb;
        const float wp = p ? __fmaf_rn(hw, (float)(8 * p), wb) : wb;
#pragma unroll
        for (int hh = 0; hh < WF_TW / 32; ++hh) {
            const float2 n2 = hh ? __ffma2_rn(A2, make_float2((float)(32 * hh), (float)(32 * hh)), np_) : np_;
            const float w = hh ? __fmaf_rn(gw, (float)(32 * hh), wp) : wp;
            const float rw = wf_rcp(w);
            const float2 r = __fmul2_rn(n2, make_float2(rw, rw));
            const int fxd = __float2int_rd(r.x), fyd = __float2int_rd(r.y);
            uint32_t out;
            if (BILINEAR) {
                // integer parts by arithmetic shift (a PRMT with byte 3 replicated is only a sign extension for |rel| < 256 px:
                // found by tools/fuzz_warp.py on coordinates clamped far outside a tile's own range), then IMAD + LEA
                const int xi = fxd >> 16, yi = fyd >> 16;
                const uint32_t a = (uint32_t)yi * P4 + base + ((uint32_t)xi << 2);
                const uint2 top = wf_lds_entry(a), bot = wf_lds_entry(a + P4);
                out = wf_blend(top, bot, (uint32_t)fxd & 0xFFFFu, (uint32_t)fyd & 0xFFFFu);
            } else {
                // nn: fetch speculatively (the footprint has a one-entry margin either way), remember
                // which pixels lie within `band` of a rounding boundary; they are settled in float64
                // after the branch-free pass so that the eight fetches of a thread stay interleaved
                const int hx = fxd + 32768, hy = fyd + 32768;
                const uint32_t fx = (uint32_t)hx & 0xFFFFu, fy = (uint32_t)hy & 0xFFFFu;
                if (fx - (uint32_t)B16 > (uint32_t)(65535 - 2 * B16) || fy - (uint32_t)B16 > (uint32_t)(65535 - 2 * B16))
                    amb |= 1u << (p * (WF_TW / 32) + hh);
                const int xi = hx >> 16, yi = hy >> 16;
                const uint2 e = wf_lds_entry((uint32_t)yi * P4 + base + ((uint32_t)xi << 2));
                out = __byte_perm(e.x, e.y, 0x4420);
            }
            if (!BILINEAR) {
                outv[p * (WF_TW / 32) + hh] = out;
            } else if (DIRECT) {
                const uint32_t nb = __shfl_down_sync(0xffffffffu, out, 1);
                const uint32_t word = __byte_perm(out, nb, psel);
                if ((lane & 3) != 3) *reinterpret_cast<uint32_t *>(gptr + 96 * hh) = word;
                if (hh == WF_TW / 32 - 1) gptr += 8 * dst_pitch;
            } else {
                // row = wrp + 8p keeps its parity, so the swizzle of oaddr[] holds for every p
                asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(oaddr[hh]), "n"(0), "r"(out));
                if (hh == WF_TW / 32 - 1) {
#pragma unroll
                    for (int k = 0; k < WF_TW / 32; ++k) oaddr[k] += (uint32_t)(8 * WF_TW * 4);
                }
            }
        }
    }
    if (!BILINEAR) {
        if (amb) {
#pragma unroll
            for (int k = 0; k < NPX; ++k)
                if (amb & (1u << k)) {
                    const int c = c0 + lane + 32 * (k % (WF_TW / 32)), r = r0 + wrp + 8 * (k / (WF_TW / 32));
                    if (!wf_defer(fix, c, r)) { outv[k] = wf_exact_px<BILINEAR>(J, src, c, r); ++n_exact; }
                }
        }
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
            const int hh = k % (WF_TW / 32), p = k / (WF_TW / 32);
            if (DIRECT) {
                const uint32_t nb = __shfl_down_sync(0xffffffffu, outv[k], 1);
                const uint32_t word = __byte_perm(outv[k], nb, psel);
                if ((lane & 3) != 3) *reinterpret_cast<uint32_t *>(gptr + (size_t)p * 8 * dst_pitch + 96 * hh) = word;
            } else {
                asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(oaddr[hh] + (uint32_t)(p * 8 * WF_TW * 4)), "n"(0), "r"(outv[k]));
            }
        }
    }
}

// Bilinear, full INTERIOR tile, direct stores: the same arithmetic as wf_compute_full<1, 1>, software-
// pipelined by hand.  The kernel is latency-bound, not issue-bound (4 consumer warps per scheduler,
// ~30-cycle shared-memory loads feeding a dependent integer chain): the eight tap loads of row p + 1
// are issued (coordinates, addresses, LDS) before the blend of row p, so that a warp always has the
// taps of two to four pixels in flight.
#ifndef B2G_SWP
#define B2G_SWP 1
#endif
#ifndef B2G_DIAG
#define B2G_DIAG 0      // diagnostics only: 1 = no sampling arithmetic, 2 = no conversion, 4 = no global stores
#endif
// BORDER = 1: the tile crosses the mask boundary of the sampler (x > bw-1 | x < 0 | y > bh-1 | y < 0 -> 0,
// homography.py:126-130).  The decision is taken on the float32 coordinates where they are further than
// `band` from the boundary; coordinates are clamped into the unmasked range before addressing (the staged
// footprint of a BORDER tile is clipped to it), masked pixels are stored as 0, and pixels inside the band go
// to the fix-up list (float64) -- or are settled here if the list is full.
// address of plane-0 entry (0, 0) relative to the anchor, for a tile staged at shared address s_src
__device__ __forceinline__ uint32_t wf_entry_base(const WfDesc &T, uint32_t s_src)
{
    return s_src - (uint32_t)(((int)T.oy * (int)T.P + (int)T.ox) * 4);
}

// DIRECT = 0: the tile's destination is not 4-byte aligned; same arithmetic, results go to the shared output
// stage `s_out` (so that a tile's bytes never depend on buffer alignment).
template <int BORDER, int DIRECT>
__device__ __forceinline__ void wf_full_bilinear(const WfDesc &T, const b200geo_warp_job &J, const uint8_t *__restrict__ src,
                                                 uint32_t base, uint32_t P4, int c0, int r0, unsigned &n_exact, const WfFix &fix,
                                                 uint8_t *__restrict__ dst_tile, size_t dst_pitch, uint32_t s_out = 0u)
{
    const int tid_ = wf_tid();
    const int lane = tid_ & 31, wrp = tid_ >> 5;
    const float di0 = (float)(lane - WF_TW / 2), dj0 = (float)(wrp - WF_TH / 2);
    const float gw = T.gw, hw = T.hw;
    const float2 A2 = make_float2(T.ax, T.ay), B2 = make_float2(T.bx, T.by);
    const float2 nb = __ffma2_rn(A2, make_float2(di0, di0), __ffma2_rn(B2, make_float2(dj0, dj0), make_float2(T.nx0, T.ny0)));
    const float wb = __fmaf_rn(gw, di0, __fmaf_rn(hw, dj0, WF_W0));
    const uint32_t psel = (lane & 3) == 0 ? 0x4210u : ((lane & 3) == 1 ? 0x5421u : 0x6542u);
    uint8_t *gptr = DIRECT ? dst_tile + (size_t)wrp * dst_pitch + (size_t)(12 * (lane >> 2) + 4 * (lane & 3)) : nullptr;
    constexpr int G = WF_TW / 32, R = WF_TH / 8;
    float lox = 0.f, hix = 0.f, loy = 0.f, hiy = 0.f;
    if (BORDER) wf_bounds<1>(T.xc, T.yc, J.bw, J.bh, lox, hix, loy, hiy);
    const float band = T.band;
    uint32_t inmask = 0u, amb = 0u;
    uint32_t fx[2][G], fy[2][G];
    uint2 top[2][G], bot[2][G];
    auto fetch = [&](int p, int b) {
        const float2 np_ = p ? __ffma2_rn(B2, make_float2((float)(8 * p), (float)(8 * p)), nb) : nb;
        const float wp = p ? __fmaf_rn(hw, (float)(8 * p), wb) : wb;
#pragma unroll
        for (int hh = 0; hh < G; ++hh) {
            const float2 n2 = hh ? __ffma2_rn(A2, make_float2((float)(32 * hh), (float)(32 * hh)), np_) : np_;
            const float w = hh ? __fmaf_rn(gw, (float)(32 * hh), wp) : wp;
            const float rw = wf_rcp(w);
            float2 r = __fmul2_rn(n2, make_float2(rw, rw));
            if (BORDER) {
                const bool in = r.x >= lox + band && r.x <= hix - band && r.y >= loy + band && r.y <= hiy - band;
                const bool outp = r.x < lox - band || r.x > hix + band || r.y < loy - band || r.y > hiy + band;
                if (in) inmask |= 1u << (p * G + hh);
                else if (!outp) amb |= 1u << (p * G + hh);          // also NaN coordinates
                r.x = fminf(fmaxf(r.x, lox), hix);                  // (NaN -> lox)
                r.y = fminf(fmaxf(r.y, loy), hiy);
            }
            const int fxd = __float2int_rd(r.x), fyd = __float2int_rd(r.y);
            const int xi = fxd >> 16, yi = fyd >> 16;
#ifdef B2G_DEBUG_FOOT
            if (xi < T.ox || xi >= T.ox + (int)T.P || yi < T.oy || yi + 1 >= T.oy + (int)T.rows)
                printf("foot: border %d tile c0 %d r0 %d mode %x xi %d yi %d ox %d oy %d P %d rows %d xc %d yc %d r %g %g lo %g %g hi %g %g band %g\n", BORDER, c0, r0, T.mode,
                       xi, yi, (int)T.ox, (int)T.oy, (int)T.P, (int)T.rows, T.xc, T.yc, r.x, r.y, lox, loy, hix, hiy, band);
#endif
            const uint32_t a = (uint32_t)yi * P4 + base + ((uint32_t)xi << 2);
            top[b][hh] = wf_lds_entry(a);
            bot[b][hh] = wf_lds_entry(a + P4);
            fx[b][hh] = (uint32_t)fxd & 0xFFFFu;
            fy[b][hh] = (uint32_t)fyd & 0xFFFFu;
        }
    };
    if (!(B2G_DIAG & 1)) fetch(0, 0);
#pragma unroll
    for (int p = 0; p < R; ++p) {
        if (p + 1 < R && !(B2G_DIAG & 1)) fetch(p + 1, (p + 1) & 1);
#pragma unroll
        for (int hh = 0; hh < G; ++hh) {
            uint32_t out = (B2G_DIAG & 1) ? (uint32_t)(lane + p) : wf_blend(top[p & 1][hh], bot[p & 1][hh], fx[p & 1][hh], fy[p & 1][hh]);
            if (BORDER && !(inmask & (1u << (p * G + hh)))) out = 0u;
            if (DIRECT) {
                const uint32_t nbr = __shfl_down_sync(0xffffffffu, out, 1);
                const uint32_t word = __byte_perm(out, nbr, psel);
                if ((B2G_DIAG & 4) ? word == 0x12345678u : (lane & 3) != 3) *reinterpret_cast<uint32_t *>(gptr + 96 * hh) = word;
            } else {
                wf_sts32(s_out + 4u * (uint32_t)wf_out_index(wrp + 8 * p, lane + 32 * hh), out);
            }
        }
        if (DIRECT) gptr += 8 * dst_pitch;
    }
    if (BORDER && __any_sync(0xffffffffu, amb != 0u)) {
        __syncwarp();                                   // the word stores of every lane precede the byte stores below
        if (amb) {
#pragma unroll 1
            for (uint32_t rest = amb; rest; rest &= rest - 1u) {
                const int k = __ffs((int)rest) - 1;
                {
                    const int col = lane + 32 * (k % G), row = wrp + 8 * (k / G);
                    if (!wf_defer(fix, c0 + col, r0 + row)) {
                        const uint32_t v = wf_exact_px<1>(J, src, c0 + col, r0 + row);
                        ++n_exact;
                        if (DIRECT) {
                            uint8_t *q = dst_tile + (size_t)row * dst_pitch + (size_t)col * 3;
                            q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16);
                        } else {
                            wf_sts32(s_out + 4u * (uint32_t)wf_out_index(row, col), v);
                        }
                    }
                }
            }
        }
    }
}

// Nearest neighbour, full INTERIOR tile, direct stores: same decisions as wf_compute_full<0, 1>, laid out
// like the bilinear loop (the fetches of row p + 1 precede the stores of row p).  Every pixel is stored
// speculatively; the ones within `band` of a rounding boundary are pushed to the fix-up list
// (warp_fixup_kernel overwrites them in float64 after this kernel) or, if the list is full, settled here.
// BORDER = 1: the tile crosses the mask boundary (rounded index outside [0, bw-1] x [0, bh-1] -> 0, homography.py:109-116):
// decided on the float32 coordinates where they are further than `band` from the boundary, indices clamped into the image
// (which also performs the truncation toward zero of (-1, 0) -> 0), masked pixels stored as 0, the rest deferred.
template <int BORDER>
__device__ __forceinline__ void wf_full_nn_direct(const WfDesc &T, const b200geo_warp_job &J, const uint8_t *__restrict__ src, uint32_t base,
                                                  uint32_t P4, int c0, int r0, unsigned &n_exact, const WfFix &fix, uint8_t *__restrict__ dst_tile,
                                                  size_t dst_pitch)
{
    const int tid_ = wf_tid();
    const int lane = tid_ & 31, wrp = tid_ >> 5;
    const float di0 = (float)(lane - WF_TW / 2), dj0 = (float)(wrp - WF_TH / 2);
    const float gw = T.gw, hw = T.hw;
    const float2 A2 = make_float2(T.ax, T.ay), B2 = make_float2(T.bx, T.by);
    const float2 nb = __ffma2_rn(A2, make_float2(di0, di0), __ffma2_rn(B2, make_float2(dj0, dj0), make_float2(T.nx0, T.ny0)));
    const float wb = __fmaf_rn(gw, di0, __fmaf_rn(hw, dj0, WF_W0));
    const uint32_t psel = (lane & 3) == 0 ? 0x4210u : ((lane & 3) == 1 ? 0x5421u : 0x6542u);
    uint8_t *gptr = dst_tile + (size_t)wrp * dst_pitch + (size_t)(12 * (lane >> 2) + 4 * (lane & 3));
    const uint32_t B16 = (uint32_t)((int)T.band + 2);
    constexpr int G = WF_TW / 32, R = WF_TH / 8;
    float lox = 0.f, hix = 0.f, loy = 0.f, hiy = 0.f;
    if (BORDER) wf_bounds<0>(T.xc, T.yc, J.bw, J.bh, lox, hix, loy, hiy);
    const float band = T.band;
    const int xlo = -T.xc, xhi = J.bw - 1 - T.xc, ylo = -T.yc, yhi = J.bh - 1 - T.yc;
    uint2 e[2][G];
    uint32_t amb = 0u, inmask = 0u;
    auto fetch = [&](int p, int b) {
        const float2 np_ = p ? __ffma2_rn(B2, make_float2((float)(8 * p), (float)(8 * p)), nb) : nb;
        const float wp = p ? __fmaf_rn(hw, (float)(8 * p), wb) : wb;
#pragma unroll
        for (int hh = 0; hh < G; ++hh) {
            const float2 n2 = hh ? __ffma2_rn(A2, make_float2((float)(32 * hh), (float)(32 * hh)), np_) : np_;
            const float w = hh ? __fmaf_rn(gw, (float)(32 * hh), wp) : wp;
            const float rw = wf_rcp(w);
            float2 r = __fmul2_rn(n2, make_float2(rw, rw));
            bool in = true;
            if (BORDER) {
                in = r.x >= lox + band && r.x <= hix - band && r.y >= loy + band && r.y <= hiy - band;
                const bool outp = r.x < lox - band || r.x > hix + band || r.y < loy - band || r.y > hiy + band;
                if (in) inmask |= 1u << (p * G + hh);
                else if (!outp) amb |= 1u << (p * G + hh);          // also NaN coordinates
                r.x = fminf(fmaxf(r.x, lox), hix);                  // keeps the conversion below in range (NaN -> lox)
                r.y = fminf(fmaxf(r.y, loy), hiy);
            }
            const int hx = __float2int_rd(r.x) + 32768, hy = __float2int_rd(r.y) + 32768;
            const uint32_t fx = (uint32_t)hx & 0xFFFFu, fy = (uint32_t)hy & 0xFFFFu;
            if (in && (fx - B16 > 65535u - 2u * B16 || fy - B16 > 65535u - 2u * B16)) amb |= 1u << (p * G + hh);
            int xi = hx >> 16, yi = hy >> 16;
            if (BORDER) { xi = min(max(xi, xlo), xhi); yi = min(max(yi, ylo), yhi); }
            e[b][hh] = wf_lds_entry((uint32_t)yi * P4 + base + ((uint32_t)xi << 2));
        }
    };
    fetch(0, 0);
#pragma unroll
    for (int p = 0; p < R; ++p) {
        if (p + 1 < R) fetch(p + 1, (p + 1) & 1);
#pragma unroll
        for (int hh = 0; hh < G; ++hh) {
            uint32_t out = __byte_perm(e[p & 1][hh].x, e[p & 1][hh].y, 0x4420);
            if (BORDER && !(inmask & (1u << (p * G + hh)))) out = 0u;
            const uint32_t nbr = __shfl_down_sync(0xffffffffu, out, 1);
            const uint32_t word = __byte_perm(out, nbr, psel);
            if ((lane & 3) != 3) *reinterpret_cast<uint32_t *>(gptr + 96 * hh) = word;
        }
        gptr += 8 * dst_pitch;
    }
    if (__any_sync(0xffffffffu, amb != 0u)) {
        __syncwarp();                                   // the speculative word stores of every lane precede the byte stores below
        if (amb) {
#pragma unroll 1
            for (uint32_t rest = amb; rest; rest &= rest - 1u) {
                const int k = __ffs((int)rest) - 1;
                const int col = lane + 32 * (k % G), row = wrp + 8 * (k / G);
                if (!wf_defer(fix, c0 + col, r0 + row)) {
                    const uint32_t v = wf_exact_px<0>(J, src, c0 + col, r0 + row);
                    ++n_exact;
                    uint8_t *q = dst_tile + (size_t)row * dst_pitch + (size_t)col * 3;
                    q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16);
                }
            }
        }
    }
}

// Any other staged tile: partial tiles and BORDER tiles (per-pixel mask with the ambiguity band).
template <int BILINEAR, int BORDER>
__device__ __forceinline__ void wf_compute_general(const WfDesc &T, const b200geo_warp_job &J, const uint8_t *__restrict__ src,
                                                uint32_t s_src, uint32_t s_out, int c0, int r0, int tw, int th, unsigned &n_exact,
                                                const WfFix &fix, int col_off = 0, int row_off = 0)
{
    // (c0, r0) is the output pixel of local (0, 0); (col_off, row_off) its place inside the staged output tile
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int P4 = (int)T.P * 4;
    const uint32_t base = s_src - (uint32_t)(((int)T.oy * (int)T.P + (int)T.ox) * 4);
    const int B16 = (int)T.band + 2;
    float lox = 0.f, hix = 0.f, loy = 0.f, hiy = 0.f;
    if (BORDER) wf_bounds<BILINEAR>(T.xc, T.yc, J.bw, J.bh, lox, hix, loy, hiy);
    for (int row = wrp; row < th; row += 8) {
        const float dj = (float)(row - WF_TH / 2);
        for (int col = lane; col < tw; col += 32) {
            float rx, ry;
            wf_rel16(T, (float)(col - WF_TW / 2), dj, rx, ry);
            uint32_t out = 0u;
            bool fast = true;
            if (BORDER) {
                const bool in = rx >= lox + T.band && rx <= hix - T.band && ry >= loy + T.band && ry <= hiy - T.band;
                const bool outp = rx < lox - T.band || rx > hix + T.band || ry < loy - T.band || ry > hiy + T.band;
                if (!in) {
                    fast = false;
                    if (!outp && !wf_defer(fix, c0 + col, r0 + row)) { out = wf_exact_px<BILINEAR>(J, src, c0 + col, r0 + row); ++n_exact; }
                }
            }
            if (fast) {
                const int fxd = __float2int_rd(rx), fyd = __float2int_rd(ry);
                if (BILINEAR) {
                    const uint32_t a = base + (uint32_t)((fyd >> 16) * P4 + (fxd >> 16) * 4);
                    const uint2 top = wf_lds_entry(a), bot = wf_lds_entry(a + (uint32_t)P4);
                    out = wf_blend(top, bot, (uint32_t)fxd & 0xFFFFu, (uint32_t)fyd & 0xFFFFu);
                } else {
                    const int hx = fxd + 32768, hy = fyd + 32768;
                    const uint32_t fx = (uint32_t)hx & 0xFFFFu, fy = (uint32_t)hy & 0xFFFFu;
                    if (fx - (uint32_t)B16 > (uint32_t)(65535 - 2 * B16) || fy - (uint32_t)B16 > (uint32_t)(65535 - 2 * B16)) {
                        if (!wf_defer(fix, c0 + col, r0 + row)) { out = wf_exact_px<BILINEAR>(J, src, c0 + col, r0 + row); ++n_exact; }
                    } else {
                        const int xi = max(hx >> 16, -T.xc), yi = max(hy >> 16, -T.yc);   // trunc toward zero: (-1, 0) -> 0
                        const uint2 e = wf_lds_entry(base + (uint32_t)(yi * P4 + xi * 4));
                        out = __byte_perm(e.x, e.y, 0x4420);
                    }
                }
            }
            wf_sts32(s_out + 4u * (uint32_t)wf_out_index(row + row_off, col + col_off), out);
        }
    }
}

// SPLIT tile: four quadrants, each set up here (warp 0, float64, like warp_setup_kernel does per
// tile), staged synchronously and sampled by the general loop.  `sync` is the barrier of the
// calling kernel's sampling threads.  A quadrant whose footprint still does not fit is exact.
template <int BILINEAR, typename Sync>
__device__ __noinline__ void wf_split_tile(const b200geo_warp_job &J, const uint8_t *__restrict__ src, WfDesc *s_sub, uint32_t s_src,
                                           uint32_t s_out, uint32_t *s_out_mem, int c0, int r0, int tw, int th, unsigned &n_exact,
                                           const WfFix &fix, Sync sync)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    for (int q = 0; q < (WF_TW / WF_SUB_W) * ((WF_TH + WF_SUB_H - 1) / WF_SUB_H); ++q) {
        const int co = (q % (WF_TW / WF_SUB_W)) * WF_SUB_W, ro = (q / (WF_TW / WF_SUB_W)) * WF_SUB_H;
        const int stw = min(WF_SUB_W, tw - co), sth = min(WF_SUB_H, th - ro);
        if (stw <= 0 || sth <= 0) continue;                       // uniform
        sync();                                                   // previous quadrant done with s_sub / the staged entries
        if (wrp == 0) {
            WfDesc t;
            wf_setup<BILINEAR>(J, c0 + co, r0 + ro, stw, sth, WfTma{0, 0, 0}, t);
            if ((t.mode & 0xFF) == WF_SPLIT) t.mode = WF_SLOW;
            if (lane == 0) *s_sub = t;
        }
        sync();
        const int m = s_sub->mode & 0xFF;
        if (m == WF_INTERIOR || m == WF_BORDER) {
            wf_stage(*s_sub, src, (size_t)J.src_pitch, J.iw, J.ih, s_src);
            sync();
            if (m == WF_INTERIOR) wf_compute_general<BILINEAR, 0>(*s_sub, J, src, s_src, s_out, c0 + co, r0 + ro, stw, sth, n_exact, fix, co, ro);
            else wf_compute_general<BILINEAR, 1>(*s_sub, J, src, s_src, s_out, c0 + co, r0 + ro, stw, sth, n_exact, fix, co, ro);
        } else {
            for (int row = wrp; row < sth; row += 8)
                for (int col = lane; col < stw; col += 32) {
                    uint32_t v = 0u;
                    if (m != WF_ZERO) { v = wf_exact_px<BILINEAR>(J, src, c0 + co + col, r0 + ro + row); ++n_exact; }
                    s_out_mem[wf_out_index(row + ro, col + co)] = v;
                }
        }
    }
}

// write the tile: RGBX words in shared memory -> packed RGB bytes in global memory
__device__ __forceinline__ void wf_copy_out(uint8_t *__restrict__ dst_job, size_t dst_pitch, const uint32_t *s_out_mem, int c0, int r0,
                                            int tw, int th)
{
    uint8_t *dst = dst_job + (size_t)r0 * dst_pitch + (size_t)c0 * 3;
    const bool vec = (((size_t)dst | dst_pitch) & 15) == 0 && tw == WF_TW;
    if (vec) {
        // task = 8 pixels of one row = 24 bytes = three 8-byte stores; 256 tasks for a full tile
#pragma unroll 1
        for (int t = wf_tid(); t < th * (WF_TW / 8); t += WF_THREADS) {
            const int row = t >> 3, h = t & 7;
            const int sw = (h >> 2) ^ ((row & 1) << 1);           // (u >> 3) & 1 with u = 2h + j
            const uint4 *s = reinterpret_cast<const uint4 *>(s_out_mem + row * WF_TW);
            const uint4 a = s[(2 * h + 0) ^ sw], b = s[(2 * h + 1) ^ sw];
            // 4 pixels (x,y,z,w = R G B .) -> 3 words: RGBR GBRG BRGB
            uint2 o0, o1, o2;
            o0.x = __byte_perm(a.x, a.y, 0x4210); o0.y = __byte_perm(a.y, a.z, 0x5421); o1.x = __byte_perm(a.z, a.w, 0x6542);
            o1.y = __byte_perm(b.x, b.y, 0x4210); o2.x = __byte_perm(b.y, b.z, 0x5421); o2.y = __byte_perm(b.z, b.w, 0x6542);
            uint2 *q = reinterpret_cast<uint2 *>(dst + (size_t)row * dst_pitch + h * 24);
            q[0] = o0; q[1] = o1; q[2] = o2;
        }
    } else {
        for (int i = threadIdx.x; i < th * WF_TW; i += WF_THREADS) {
            const int row = i >> 6, col = i & (WF_TW - 1);
            if (col >= tw) continue;
            const uint32_t v = s_out_mem[wf_out_index(row, col)];
            uint8_t *q = dst + (size_t)row * dst_pitch + col * 3;
            q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16);
        }
    }
}

// Grid: (tiles_x, tiles_y, n_jobs).  stats[0] += pixels evaluated in float64.
#ifndef B2G_WF_OCC
#define B2G_WF_OCC 4
#endif
template <int BILINEAR>
__global__ void __launch_bounds__(WF_THREADS, B2G_WF_OCC)
warp_fast_kernel(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
                 const b200geo_warp_job *__restrict__ jobs, const WfDesc *__restrict__ descs,
                 unsigned long long *__restrict__ stats, WfFixEntry *__restrict__ fix_list, unsigned *__restrict__ fix_count, unsigned fix_cap)
{
    // (static shared memory: the comparison kernel exists only for tile shapes whose buffers fit 48 KB)
    constexpr bool FITS = (2 * WF_SRC_ENTRIES + WF_TW * WF_TH) * 4 <= 47 * 1024;
    __shared__ __align__(16) uint32_t s_src_mem[FITS ? 2 * WF_SRC_ENTRIES : 4];
    __shared__ __align__(16) uint32_t s_out_mem[FITS ? WF_TW * WF_TH : 4];
    if (!FITS) { __trap(); }

    const WfDesc &T = descs[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x];
    const int mode = T.mode & 0xFF;
    if (mode == WF_EMPTY) return;
    const b200geo_warp_job &J = jobs[blockIdx.z];
    const int c0 = blockIdx.x * WF_TW, r0 = blockIdx.y * WF_TH;
    const int tw = min(WF_TW, J.ow - c0), th = min(WF_TH, J.oh - r0);
    const uint8_t *src = src_base + J.src_offset;
    const uint32_t s_src = (uint32_t)__cvta_generic_to_shared(s_src_mem);
    const uint32_t s_out = (uint32_t)__cvta_generic_to_shared(s_out_mem);
    unsigned n_exact = 0;
    const WfFix fix = {fix_list, fix_count, fix_cap, (int)blockIdx.z};

    if (mode == WF_INTERIOR || mode == WF_BORDER) {
        wf_stage(T, src, (size_t)J.src_pitch, J.iw, J.ih, s_src);
        __syncthreads();
        if (mode == WF_INTERIOR && tw == WF_TW && th == WF_TH) wf_compute_full<BILINEAR, 0>(T, J, src, s_src, s_out, c0, r0, n_exact, fix);
        else if (BILINEAR && B2G_SWP && mode == WF_BORDER && tw == WF_TW && th == WF_TH)
            wf_full_bilinear<1, 0>(T, J, src, wf_entry_base(T, s_src), (uint32_t)T.P * 4u, c0, r0, n_exact, fix, nullptr, 0, s_out);
        else if (mode == WF_INTERIOR) wf_compute_general<BILINEAR, 0>(T, J, src, s_src, s_out, c0, r0, tw, th, n_exact, fix);
        else wf_compute_general<BILINEAR, 1>(T, J, src, s_src, s_out, c0, r0, tw, th, n_exact, fix);
    } else if (mode == WF_ZERO) {
        for (int i = threadIdx.x; i < WF_TW * WF_TH; i += WF_THREADS) s_out_mem[i] = 0u;
    } else if (mode == WF_SPLIT) {
        __shared__ WfDesc s_sub;
        wf_split_tile<BILINEAR>(J, src, &s_sub, s_src, s_out, s_out_mem, c0, r0, tw, th, n_exact, fix, [] { __syncthreads(); });
    } else {
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        for (int row = wrp; row < th; row += 8)
            for (int col = lane; col < tw; col += 32) {
                s_out_mem[wf_out_index(row, col)] = wf_exact_px<BILINEAR>(J, src, c0 + col, r0 + row);
                ++n_exact;
            }
    }
    __syncthreads();

    wf_copy_out(dst_base + J.dst_offset, (size_t)J.dst_pitch, s_out_mem, c0, r0, tw, th);
    if (stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
}


// ------------------------------------------------------------------------------------------------
// Persistent, warp-specialised, software-pipelined form of warp_fast_kernel.
//
// warp_fast_kernel exposes two serial memory latencies per tile (descriptor, then the source rows
// the descriptor points at) and hides them only with the 4 CTAs an SM holds.  Here each CTA walks
// tiles  blockIdx.x, blockIdx.x + gridDim.x, ...  with 8 consumer warps and 1 producer warp:
//
//   producer: reads descriptor i (prefetched two tiles ahead), the job's geometry, decides how the
//     tile is staged, publishes a WpTile record in shared memory, arms the stage's `full` mbarrier
//     with the byte count and issues one TMA bulk copy (cp.async.bulk, SASS UBLKCP) per source
//     row -- 16-byte aligned raw RGB bytes -- into the stage's raw buffer.  It runs up to two tiles
//     ahead of the consumers, gated by the stage's `empty` mbarrier.
//   consumers: wait on `full` (normally already complete), convert the raw rows to paired entries
//     shared-to-shared, release the stage (`empty`), sample, and write the tile out.  They
//     synchronise among themselves with a named barrier; the producer never joins it.
//
// Tiles whose rows cannot be copied in bulk (image edges, footprints larger than the raw buffer)
// are staged synchronously by the consumers as in warp_fast_kernel.
// ------------------------------------------------------------------------------------------------
#ifndef B2G_RAW_SLACK
#define B2G_RAW_SLACK 2048
#endif
constexpr int WP_RAW_BYTES = WF_TMA_BOX_W * WF_TMA_BOX_H + B2G_RAW_SLACK;   // one raw-row buffer (x2)
constexpr int WP_OCC = B2G_WP_OCC;    // CTAs per SM (shared memory: 2 x raw 14 KB + 2 x entries 32 KB + output 8 KB)
constexpr int WP_THREADS = WF_THREADS + 32;

// One tile of the persistent kernel: written to global memory by warp_tile_setup_kernel (192 bytes per tile), copied
// into a shared-memory ring slot by the producer, which only adds the addresses that depend on the pipeline stage.
struct __align__(16) WpTile {
    WfDesc T;
    const uint8_t *src;      // image base (job offset applied)
    const uint8_t *a0;       // address of source pixel (ay0, ax0)
    uint8_t *dst;            // output base of the job
    uint8_t *dst_tile;       // address of the tile's first output byte
    long long src_pitch, dst_pitch;
    int iw, ih, c0, r0;
    int tw, th, jz, img;     // img: index of the job's image in the 3-D tensor map
    uint32_t RB;             // bytes per raw row in shared memory (multiple of 16)
    int staged, bulk;
    int tma;                 // raw rows arrive as one tensor copy: row stride WF_TMA_BOX_W, one phase
    int direct;              // full tile on the unrolled loop, stored straight to global memory
    int tma_x;               // tensor copy: first byte column of the box (multiple of 16, may be negative)
    // entry (x, y) relative to the anchor lives at sm_base + y * sm_P4 + 4 x (plane 0); set-up leaves the offset
    // relative to the entries buffer in sm_base, the producer adds the buffer's address
    uint32_t sm_base, sm_P4;
    // tensor-copy tiles: chunk t = i * chunks + k is read at cv_a0 + 12 t + i * cv_D, written at cv_dst + 16 t
    // (set-up leaves the row phase in cv_a0; the producer adds the raw stage's address and sets cv_dst)
    uint32_t cv_a0, cv_D, cv_dst;
    int cv_total;            // chunks to convert (0 unless tma)
};
static_assert(sizeof(WpTile) == 192, "WpTile is copied as twelve 16-byte words");
constexpr int WP_TILE_WORDS = sizeof(WpTile) / 16;

#ifndef B2G_STAGES
#define B2G_STAGES 2
#endif
constexpr int WP_STAGES = B2G_STAGES; // raw-row buffers in flight
constexpr int WP_TILE_RING = 8;       // tile records: a record outlives its raw buffer (read until the copy-out); >= WP_STAGES + 2

#ifndef B2G_PAIRS
#define B2G_PAIRS 2     // entry buffers: 2 = conversion of tile i+1 overlaps sampling of tile i inside a CTA; 1 = serialised (less shared memory)
#endif
#ifndef B2G_PROF
#define B2G_PROF 0      // diagnostics only: per-warp cycle accounting printed by a few CTAs
#endif
#if B2G_PROF
#define WP_T(var) const long long var = clock64()
#define WP_ACC(acc, a, b) acc += (b) - (a)
#else
#define WP_T(var)
#define WP_ACC(acc, a, b)
#endif
struct WpSmem {
    uint8_t raw[WP_STAGES][WP_RAW_BYTES];
    uint32_t pairs[B2G_PAIRS][2 * WF_SRC_ENTRIES];   // double-buffered; each: plane 0, plane 1
    uint32_t out[WF_TW * WF_TH];             // output stage of the tiles that are not written directly
    WpTile tile[WP_TILE_RING];
    WfDesc sub;              // SPLIT tiles: the quadrant being processed
    unsigned long long full[WP_STAGES], empty[WP_STAGES];
    unsigned long long conv[2];              // consumer-side: entries of buffer s complete
    unsigned fix_n;                          // deferred float64 pixels of this CTA (its segment of the fix-up list)
};

__device__ __forceinline__ void wp_mbar_init(uint32_t a, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void wp_mbar_expect_tx(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wp_mbar_arrive(uint32_t a)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
#ifndef B2G_WAIT
#define B2G_WAIT 1      // 0: try_wait with a 10 ms suspend hint + nanosleep back-off; 1: plain try_wait loop; 2: test_wait spin
#endif
__device__ __forceinline__ bool wp_mbar_try(uint32_t a, uint32_t parity)
{
    uint32_t ok;
#if B2G_WAIT == 0
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity), "r"(0x989680u) : "memory");
#elif B2G_WAIT == 1
    // no suspend-time hint: the hardware's default (short) suspend window; measured on B200, the hinted
    // form (SASS NANOSLEEP.SYNCS <hint>) wakes late enough to cost ~0.5 us per tile of pure protocol time
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#endif
    return ok != 0;
}
__device__ __noinline__ void wp_mbar_wait_slow(uint32_t a, uint32_t parity)
{
    // a lost copy / protocol error must not hang the GPU: give up after 4 s of wall time (a legitimate
    // wait lasts microseconds; the limit is on time, not on tries -- a try may return early)
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 64; ++i) {
            if (wp_mbar_try(a, parity)) return;
#if B2G_WAIT == 0
            __nanosleep(64);
#elif B2G_WAIT == 2
            __nanosleep(20);
#endif
        }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void wp_mbar_wait(uint32_t a, uint32_t parity)
{
    if (wp_mbar_try(a, parity)) return;
    if (wp_mbar_try(a, parity)) return;
    wp_mbar_wait_slow(a, parity);
}
__device__ __forceinline__ void wp_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ uint32_t wp_lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void wp_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WF_THREADS) : "memory"); }

// Raw RGB rows in shared memory -> entries, for rows that arrived by per-row bulk copies: row i of the footprint starts
// `(phase0 + i * pitch32) & 15` bytes into its raw row (each copy starts at the row address rounded down to 16).
// ALIGNED: all phases are multiples of 4.  (Tensor-copy tiles -- one phase for all rows -- are converted inline in
// the consumer loop from constants the producer leaves in the tile record.)
template <int ALIGNED>
__device__ __forceinline__ void wp_convert(const WfDesc &T, uint32_t magic, uint32_t phase0, uint32_t pitch32, uint32_t RB, uint32_t s_raw,
                                           uint32_t s_pairs)
{
    const int P = T.P, chunks = P >> 2, total = (B2G_DIAG & 2) ? 0 : (int)T.rows * chunks;
#pragma unroll 2
    for (int t = wf_tid(); t < total; t += WF_THREADS) {
        const int i = chunks > 1 ? (int)__umulhi((unsigned)t, magic) : t;
        const int k = t - i * chunks;
        const uint32_t o = ((phase0 + (uint32_t)i * pitch32) & 15u) + 12u * (uint32_t)k;   // byte offset in the raw row
        const uint32_t a = s_raw + (uint32_t)i * RB + (o & ~3u);
        uint32_t w0 = wp_lds32(a), w1 = wp_lds32(a + 4), w2 = wp_lds32(a + 8), w3 = wp_lds32(a + 12);
        if (!ALIGNED) {
            const uint32_t sh = (o & 3u) * 8u, w4 = wp_lds32(a + 16);
            w0 = __funnelshift_r(w0, w1, sh); w1 = __funnelshift_r(w1, w2, sh);
            w2 = __funnelshift_r(w2, w3, sh); w3 = __funnelshift_r(w3, w4, sh);
        }
        wf_store_chunk(s_pairs + (uint32_t)(i * P + 4 * k) * 4u, w0, w1, w2, w3);
    }
}

// the rare ways of filling the entries (per-row bulk copies; synchronous staging at image edges without a 3-D tensor
// map): out of line, so that they do not weigh on the register allocation of the consumer loop
__device__ __noinline__ void wp_convert_cold(const WpTile &W, uint32_t s_raw, uint32_t s_pairs)
{
    if (W.bulk) wp_convert<0>(W.T, W.T.magic, (uint32_t)(size_t)W.a0, (uint32_t)W.src_pitch, W.RB, s_raw, s_pairs);
    else if (W.staged) wf_stage(W.T, W.src, (size_t)W.src_pitch, W.iw, W.ih, s_pairs);
}

// Grid: min(n_tiles, SMs * WP_OCC) CTAs of WP_THREADS; dynamic shared memory sizeof(WpSmem).
// Grid: (ceil(tiles_x / 32), ceil(tiles_y / 8), n_jobs), block (32, 8): one thread per tile.  The whole tile record
// (everything that does not depend on the pipeline stage) is prepared here, so that the producer warp of the
// persistent kernel does no address arithmetic, divisions or job look-ups: measured, it had become the slowest warp.
// The complete record of one tile (everything that does not depend on the pipeline stage).
template <int BILINEAR>
__device__ __forceinline__ void wp_build_tile(const b200geo_warp_job &J, int jz, int c0, int r0, const uint8_t *__restrict__ src_base,
                                              uint8_t *__restrict__ dst_base, const WfTma &tma, int use_tmap, WpTile &W)
{
    {
        uint4 *z = reinterpret_cast<uint4 *>(&W);
#pragma unroll
        for (int i = 0; i < WP_TILE_WORDS; ++i) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    wf_setup<BILINEAR>(J, c0, r0, min(WF_TW, J.ow - c0), min(WF_TH, J.oh - r0), tma, W.T);
    const int mode = W.T.mode & 0xFF;
    const int P = W.T.P, rows = W.T.rows;
    const uint8_t *src = src_base + J.src_offset;
    uint8_t *dstj = dst_base + J.dst_offset;
    const int ay0 = W.T.yc + W.T.oy, ax0 = W.T.xc + W.T.ox;
    const uint8_t *a0 = src + (long long)ay0 * J.src_pitch + 3LL * ax0;
    const uint32_t RB = (uint32_t)(15 + 3 * (P + 1) + 15) & ~15u;
    const bool staged = mode == WF_INTERIOR || mode == WF_BORDER;
    const bool tm = use_tmap && staged && (W.T.mode & WF_FLAG_TMA);
    bool bulk = tm;
    if (!tm && staged && !(W.T.mode & WF_FLAG_TMA) && (W.T.mode & WF_FLAG_UNCHECKED) && (uint32_t)rows * RB <= (uint32_t)WP_RAW_BYTES) {
        const size_t pitch = (size_t)J.src_pitch;
        const size_t lo = (size_t)a0 & ~(size_t)15;
        const size_t hi = ((size_t)a0 + (size_t)(rows - 1) * pitch) & ~(size_t)15;
        const size_t img_end = (size_t)src + (size_t)(J.ih - 1) * pitch + 3 * (size_t)J.iw;
        bulk = lo >= (size_t)src && hi + RB <= img_end;
    }
    W.src = src; W.a0 = a0; W.dst = dstj;
    W.src_pitch = J.src_pitch; W.dst_pitch = J.dst_pitch;
    W.iw = J.iw; W.ih = J.ih; W.c0 = c0; W.r0 = r0;
    W.tw = min(WF_TW, J.ow - c0); W.th = min(WF_TH, J.oh - r0);
    W.jz = jz;
    W.img = (tma.image_rows > 0 && J.src_pitch > 0) ? (int)(J.src_offset / (J.src_pitch * tma.image_rows)) : 0;
    W.RB = tm ? (uint32_t)WF_TMA_BOX_W : RB; W.staged = staged; W.bulk = bulk; W.tma = tm;
    uint8_t *const dt = dstj + (long long)r0 * J.dst_pitch + 3LL * c0;
    W.dst_tile = dt;
    W.direct = (mode == WF_INTERIOR || (B2G_SWP && mode == WF_BORDER)) && W.tw == WF_TW && W.th == WF_TH &&
               ((((size_t)dt) | (size_t)J.dst_pitch) & 3) == 0;
    W.tma_x = (3 * ax0) & ~15;
    W.sm_P4 = (uint32_t)P * 4u;
    W.sm_base = wf_entry_base(W.T, 0u);
    W.cv_total = tm ? rows * (P >> 2) : 0;
    W.cv_D = (uint32_t)WF_TMA_BOX_W - 12u * (uint32_t)(P >> 2);
    W.cv_a0 = (uint32_t)(size_t)a0 & 15u;
}

template <int BILINEAR>
__global__ void __launch_bounds__(256)
warp_tile_setup_kernel(const b200geo_warp_job *__restrict__ jobs, const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
                       const WfTma tma, int use_tmap, int tiles_x, int tiles_y, WpTile *__restrict__ tiles)
{
    const int tx = blockIdx.x * 32 + threadIdx.x, ty = blockIdx.y * 8 + threadIdx.y;
    if (tx >= tiles_x || ty >= tiles_y) return;
    const b200geo_warp_job &J = jobs[blockIdx.z];
    WpTile W;
    const int c0 = tx * WF_TW, r0 = ty * WF_TH;
    if (c0 >= J.ow || r0 >= J.oh || !wf_job_ok(J)) {
        uint4 *z = reinterpret_cast<uint4 *>(&W);
#pragma unroll
        for (int i = 0; i < WP_TILE_WORDS; ++i) z[i] = make_uint4(0u, 0u, 0u, 0u);
        W.T.mode = WF_EMPTY;
    } else {
        wp_build_tile<BILINEAR>(J, (int)blockIdx.z, c0, r0, src_base, dst_base, tma, use_tmap, W);
    }
    const uint4 *sv = reinterpret_cast<const uint4 *>(&W);
    uint4 *d = reinterpret_cast<uint4 *>(tiles + ((size_t)blockIdx.z * tiles_y + ty) * tiles_x + tx);
#pragma unroll
    for (int i = 0; i < WP_TILE_WORDS; ++i) d[i] = sv[i];
}

// Registers: 9 warps are allocated as 10 (pairs), so two CTAs per SM leave 65536 / (2 x 320) = 102 -> 96 per
// thread, which is what __launch_bounds__ makes ptxas use (__maxnreg__(112) was measured: one CTA per SM).
template <int BILINEAR>
__global__ void __launch_bounds__(WP_THREADS, WP_OCC)
warp_pipe_kernel(const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
                 const b200geo_warp_job *__restrict__ jobs, const WpTile *__restrict__ tiles,
                 int n_tiles_host, const unsigned *__restrict__ n_tiles_dev, int tiles_x, int tiles_y, unsigned long long *__restrict__ stats,
                 const __grid_constant__ CUtensorMap tmap, int use_tmap, int tma_image_rows,
                 WfFixEntry *__restrict__ fix_list, unsigned *__restrict__ fix_counts, unsigned seg_cap, unsigned *__restrict__ tile_counter)
{
    extern __shared__ __align__(16) unsigned char wp_smem[];
    WpSmem &S = *reinterpret_cast<WpSmem *>(wp_smem);
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int first = blockIdx.x;
    // routed launches (warp_strip.cuh) keep the tile list compact and its length on the device
    const int n_tiles = n_tiles_dev ? (int)*n_tiles_dev : n_tiles_host;
    if (first >= n_tiles) {
        if (tid == 0) fix_counts[blockIdx.x] = 0u;
        return;
    }
    // shared-space addresses are computed, never held in (locally-indexed) arrays
    const uint32_t s_raw0 = (uint32_t)__cvta_generic_to_shared(S.raw[0]);
    const uint32_t s_full0 = (uint32_t)__cvta_generic_to_shared(&S.full[0]);
    const uint32_t s_empty0 = (uint32_t)__cvta_generic_to_shared(&S.empty[0]);
    const uint32_t s_pairs0 = (uint32_t)__cvta_generic_to_shared(S.pairs[0]);
    constexpr uint32_t PAIRS_BYTES = 2u * WF_SRC_ENTRIES * 4u;

    if (tid == 0) {
        S.fix_n = 0u;
        for (int i = 0; i < WP_STAGES; ++i) {
            wp_mbar_init(s_full0 + 8u * i, 1);
            wp_mbar_init(s_empty0 + 8u * i, WF_THREADS / 32);
        }
        for (int i = 0; i < 2; ++i) {
            wp_mbar_init((uint32_t)__cvta_generic_to_shared(&S.conv[i]), WF_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (wrp == WF_THREADS / 32) {
        // ---------------------------------------------------------------- producer warp
        // Tiles are claimed dynamically: the first one is blockIdx.x, every further one gridDim.x + a ticket
        // from a global counter, drawn three tiles ahead (the record prefetch distance is two).  Static striding
        // left the CTAs 10 % apart at the end (BORDER / SPLIT / SLOW tiles cost several times a direct one).
        auto claim = [&]() {
            int v = 0;
            if (lane == 0) v = (int)atomicAdd(tile_counter, 1u);
            return __shfl_sync(0xffffffffu, v, 0) + (int)gridDim.x;
        };
        const uint4 *gt = reinterpret_cast<const uint4 *>(tiles);
        int t_cur = first, t_n1 = claim(), t_n2 = claim();
        uint4 nd0 = make_uint4(0u, 0u, 0u, 0u), nd1 = nd0;      // lanes 0..11: records of the next two tiles
        if (lane < WP_TILE_WORDS) {
            nd0 = __ldg(gt + (size_t)t_cur * WP_TILE_WORDS + lane);
            if (t_n1 < n_tiles) nd1 = __ldg(gt + (size_t)t_n1 * WP_TILE_WORDS + lane);
        }
        int it = 0;
#if B2G_PROF
        long long prof_wait = 0, prof_claim = 0, prof_build = 0;
        const long long prof_t0 = clock64();
#endif
        for (;; ++it) {
            // the ticket of the tile three iterations ahead: drawn here, read at the end of the iteration, so that
            // the atomic's round trip is hidden
            int ticket = 0;
            if (lane == 0) ticket = (int)atomicAdd(tile_counter, 1u);
            const int b = it % WP_STAGES, use = it / WP_STAGES;
            const uint32_t s_full = s_full0 + 8u * b, s_raw = s_raw0 + (uint32_t)WP_RAW_BYTES * b;
            WP_T(tp0);
            if (use >= 1) wp_mbar_wait(s_empty0 + 8u * b, (uint32_t)(use - 1) & 1u);
            WP_T(tp1);
            WP_ACC(prof_wait, tp0, tp1);
            WpTile &W = S.tile[it % WP_TILE_RING];
            if (t_cur >= n_tiles) {
                // out of tiles: publish the sentinel the consumers stop at
                if (lane == 0) {
                    W.T.mode = WF_DONE; W.staged = 0; W.bulk = 0; W.tma = 0;
                    wp_mbar_arrive(s_full);
                }
                break;
            }
            if (lane < WP_TILE_WORDS) reinterpret_cast<uint4 *>(&W)[lane] = nd0;
            nd0 = nd1;
            if (lane < WP_TILE_WORDS && t_n2 < n_tiles) nd1 = __ldg(gt + (size_t)t_n2 * WP_TILE_WORDS + lane);
            __syncwarp();
            const bool tma = W.tma != 0, bulk = W.bulk != 0;
            const int rows = W.T.rows;
            const uint32_t RB = W.RB;
            const uint8_t *a0 = W.a0;
            const size_t pitch = (size_t)W.src_pitch;
            if (lane == 0) {
                const uint32_t pairs_it = s_pairs0 + PAIRS_BYTES * (uint32_t)(it & (B2G_PAIRS - 1));
                W.sm_base += pairs_it;
                W.cv_a0 += s_raw;
                W.cv_dst = pairs_it;
                if (tma && (B2G_DIAG & 8)) wp_mbar_arrive(s_full);
                else if (tma) {
                    wp_mbar_expect_tx(s_full, (uint32_t)(WF_TMA_BOX_W * WF_TMA_BOX_H));
                    if (tma_image_rows > 0)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                     ::"r"(s_raw), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(W.tma_x), "r"(W.T.srow), "r"(W.img), "r"(s_full) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                     ::"r"(s_raw), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(W.tma_x), "r"(W.T.srow), "r"(s_full) : "memory");
                } else if (bulk) wp_mbar_expect_tx(s_full, (uint32_t)rows * RB);
                else wp_mbar_arrive(s_full);
            }
            WP_T(tp2);
            // next tile of this CTA
            t_cur = t_n1; t_n1 = t_n2; t_n2 = __shfl_sync(0xffffffffu, ticket, 0) + (int)gridDim.x;
            __syncwarp();
            WP_T(tp3);
            WP_ACC(prof_claim, tp2, tp3);
            WP_ACC(prof_build, tp1, tp2);
            if (bulk && !tma)
                for (int r = lane; r < rows; r += 32) {
                    const size_t ar = ((size_t)a0 + (size_t)r * pitch) & ~(size_t)15;
                    wp_bulk_g2s(s_raw + (uint32_t)r * RB, reinterpret_cast<const void *>(ar), RB, s_full);
                }
        }
#if B2G_PROF
        if (lane == 0 && (blockIdx.x % 97) == 0)
            printf("cta %3d producer: tiles %d total %lld wait_empty %lld work/tile %lld build/tile %lld claim/tile %lld\n", (int)blockIdx.x, it,
                   clock64() - prof_t0, prof_wait, (clock64() - prof_t0 - prof_wait) / (it ? it : 1), prof_build / (it ? it : 1), prof_claim / (it ? it : 1));
#endif
        return;
    }

    // -------------------------------------------------------------------- consumer warps
    // Per warp and tile `it` (buffer s = it & 1):
    //     wait conv[s] -> sample tile it from entries[s], storing straight to global memory
    //                  -> convert tile it+1 (raw -> entries[s^1]) -> release the raw stage -> arrive conv[s^1]
    // The warps meet only at that one wait, a whole phase after the arrive: a warp that finishes
    // sampling early goes straight on to the (shared-memory-bound) conversion of the next tile
    // while slower warps still sample (issue-bound).  entries[s^1] is rewritten only after conv[s]
    // completed, i.e. after every warp converted tile it, which each does after sampling tile it-1
    // (the last reader of entries[s^1]).  Tiles that are not full INTERIOR tiles (9 % on the
    // benchmark) go through the shared output stage between two consumer barriers.
    const uint32_t s_out = (uint32_t)__cvta_generic_to_shared(S.out);
    const uint32_t s_conv0 = (uint32_t)__cvta_generic_to_shared(&S.conv[0]);
    unsigned n_exact = 0;
    // the tile index itself travels in the WpTile record: the consumers only count iterations, up to the sentinel record

#if B2G_PROF
    long long prof_conv_wait = 0, prof_full_wait = 0, prof_sample = 0, prof_convert = 0, prof_staged = 0;
    const long long prof_t0 = clock64();
#endif
    auto convert = [&](int t) {        // raw stage t % WP_STAGES -> entries[t & 1]
        const int b = t % WP_STAGES, use = t / WP_STAGES;
        WP_T(tc0);
        wp_mbar_wait(s_full0 + 8u * b, (uint32_t)use & 1u);
        WP_T(tc1);
        WP_ACC(prof_full_wait, tc0, tc1);
        // the record's ring slot is rewritten WP_TILE_RING tiles later; the producer is at most
        // WP_STAGES tiles ahead of this conversion, so the record stays valid through the copy-out
        const WpTile &W = S.tile[t % WP_TILE_RING];
        const uint32_t s_pairs = s_pairs0 + PAIRS_BYTES * (uint32_t)(t & (B2G_PAIRS - 1)), s_raw = s_raw0 + (uint32_t)WP_RAW_BYTES * b;
        if (B2G_PAIRS == 1 && t > 0) wp_consumer_sync();      // every warp is done sampling the previous tile from the one buffer
        if (W.tma) {
            // (chunk, row) from t by multiplication (chunks >= 2 by construction); all constants come from the tile record
            const int total = (B2G_DIAG & 2) ? 0 : W.cv_total;
            const uint32_t a0 = W.cv_a0, D = W.cv_D, dst = W.cv_dst, magic = W.T.magic;
#pragma unroll 2
            for (int tt = tid; tt < total; tt += WF_THREADS) {
                const uint32_t a = a0 + 12u * (uint32_t)tt + __umulhi((unsigned)tt, magic) * D;
                const uint32_t w0 = wp_lds32(a), w1 = wp_lds32(a + 4), w2 = wp_lds32(a + 8), w3 = wp_lds32(a + 12);
                wf_store_chunk(dst + 16u * (uint32_t)tt, w0, w1, w2, w3);
            }
        } else wp_convert_cold(W, s_raw, s_pairs);
        __syncwarp();
        if (lane == 0) {
            wp_mbar_arrive(s_empty0 + 8u * b);                 // raw[b] may be refilled
            wp_mbar_arrive(s_conv0 + 8u * (uint32_t)(t & 1));
        }
        WP_T(tc2);
        WP_ACC(prof_convert, tc1, tc2);
    };

    convert(0);
    int it = 0;
    for (;; ++it) {
        const int s = it & 1;
        const uint32_t s_pairs = s_pairs0 + PAIRS_BYTES * (uint32_t)(s & (B2G_PAIRS - 1));
        const WpTile &W = S.tile[it % WP_TILE_RING];
        const int mode = W.T.mode & 0xFF;
        WP_T(tw0);
        wp_mbar_wait(s_conv0 + 8u * (uint32_t)s, (uint32_t)(it >> 1) & 1u);
        if (mode == WF_DONE) break;
        WP_T(tw1);
        WP_ACC(prof_conv_wait, tw0, tw1);
        uint8_t *const dst_tile = W.dst_tile;
        const WfFix fix = {fix_list + (size_t)blockIdx.x * seg_cap, &S.fix_n, seg_cap, W.jz};
        if (W.direct) {
            if (BILINEAR && B2G_SWP) {
                if (mode == WF_INTERIOR) wf_full_bilinear<0, 1>(W.T, jobs[W.jz], W.src, W.sm_base, W.sm_P4, W.c0, W.r0, n_exact, fix, dst_tile, (size_t)W.dst_pitch);
                else wf_full_bilinear<1, 1>(W.T, jobs[W.jz], W.src, W.sm_base, W.sm_P4, W.c0, W.r0, n_exact, fix, dst_tile, (size_t)W.dst_pitch);
            } else if (!BILINEAR && B2G_SWP) {
                if (mode == WF_INTERIOR) wf_full_nn_direct<0>(W.T, jobs[W.jz], W.src, W.sm_base, W.sm_P4, W.c0, W.r0, n_exact, fix, dst_tile, (size_t)W.dst_pitch);
                else wf_full_nn_direct<1>(W.T, jobs[W.jz], W.src, W.sm_base, W.sm_P4, W.c0, W.r0, n_exact, fix, dst_tile, (size_t)W.dst_pitch);
            }
            else wf_compute_full<BILINEAR, 1>(W.T, jobs[W.jz], W.src, s_pairs, 0u, W.c0, W.r0, n_exact, fix, dst_tile, (size_t)W.dst_pitch);
            WP_T(tw2);
            WP_ACC(prof_sample, tw1, tw2);
            convert(it + 1);
            continue;
        }
        // staged tiles: every consumer warp takes this branch for the same tiles (the record is shared)
        if (mode != WF_EMPTY && !(B2G_DIAG & 16)) {
            const b200geo_warp_job &J = jobs[W.jz];
            wp_consumer_sync();          // the previous staged tile's copy-out has finished reading S.out
            if (W.staged) {
                // (a full tile whose destination is not 4-byte aligned takes the same arithmetic as the direct path)
                if (mode == WF_INTERIOR && W.tw == WF_TW && W.th == WF_TH) wf_compute_full<BILINEAR, 0>(W.T, J, W.src, s_pairs, s_out, W.c0, W.r0, n_exact, fix);
                else if (BILINEAR && B2G_SWP && mode == WF_BORDER && W.tw == WF_TW && W.th == WF_TH)
                    wf_full_bilinear<1, 0>(W.T, J, W.src, W.sm_base, W.sm_P4, W.c0, W.r0, n_exact, fix, nullptr, 0, s_out);
                else if (mode == WF_INTERIOR) wf_compute_general<BILINEAR, 0>(W.T, J, W.src, s_pairs, s_out, W.c0, W.r0, W.tw, W.th, n_exact, fix);
                else wf_compute_general<BILINEAR, 1>(W.T, J, W.src, s_pairs, s_out, W.c0, W.r0, W.tw, W.th, n_exact, fix);
            } else if (mode == WF_ZERO) {
                for (int i = tid; i < WF_TW * WF_TH; i += WF_THREADS) S.out[i] = 0u;
            } else if (mode == WF_SPLIT) {
                wf_split_tile<BILINEAR>(J, W.src, &S.sub, s_pairs, s_out, S.out, W.c0, W.r0, W.tw, W.th, n_exact, fix, [] { wp_consumer_sync(); });
            } else {
                for (int row = wrp; row < W.th; row += 8)
                    for (int col = lane; col < W.tw; col += 32) {
                        S.out[wf_out_index(row, col)] = wf_exact_px<BILINEAR>(J, W.src, W.c0 + col, W.r0 + row);
                        ++n_exact;
                    }
            }
            wp_consumer_sync();
        }
        WP_T(tw3);
        WP_ACC(prof_staged, tw1, tw3);
        convert(it + 1);
        if (mode != WF_EMPTY && !(B2G_DIAG & 16)) wf_copy_out(W.dst, (size_t)W.dst_pitch, S.out, W.c0, W.r0, W.tw, W.th);
    }
#if B2G_PROF
    if (lane == 0 && (blockIdx.x % 97) == 0 && (wrp == 0 || wrp == 7))
        printf("cta %3d warp %d: tiles %d total %lld conv_wait %lld full_wait %lld sample %lld convert %lld staged %lld\n", (int)blockIdx.x, wrp,
               it, clock64() - prof_t0, prof_conv_wait, prof_full_wait, prof_sample, prof_convert, prof_staged);
#endif
    if (stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
    wp_consumer_sync();                                  // every consumer warp has pushed its last deferred pixel
    if (tid == 0) fix_counts[blockIdx.x] = S.fix_n;
}

}  // namespace b200geo
