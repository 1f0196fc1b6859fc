// warp_strip.cuh -- K4 fast path, second generation: one WARP per 32 x 20 output strip, no CTA-wide lockstep.
//
// Same element semantics and the same parity rules as warp_fast.cuh (reference homography.py:108-209 + the
// astype(uint8) of :225/:241); what changes is the organisation.  warp_pipe_kernel (warp_fast.cuh) walks 64 x 40
// tiles with 8 sampling warps per CTA that meet at one barrier per tile, and spends 57 thread-instructions per
// pixel, a quarter of them on the per-tile protocol (profiles/r01_warp_pipe_decomposition.md).  Here
//
//   * every warp owns a strip of 32 x 20 output pixels (lane = column, 20 pixels per lane): its own TMA tensor
//     copy of the strip's source footprint (160 B x 26 rows, issued one strip ahead by the warp itself), its own
//     staged entries (48 entries x 26 rows), its own mbarrier.  Warps never wait for each other: a 16-warp CTA
//     is 16 independent pipelines that share only the tile-ticket counter and the fix-up list segment.
//   * a 64 x 40 tile is four strips; a warp draws one tile ticket (global atomic) per four strips.  The strip
//     records (80 B: coefficients, tensor-copy coordinates, destination, footprint shape) are written by the
//     set-up kernel and fetched into shared memory one strip ahead with cp.async.
//   * coordinates: the float32 tile-relative quotient is rounded to 16.16 fixed point by the FFMA that forms it
//     (n/w + 1.5 * 2^23: the integer lands in the mantissa), so the two F2I, the integer-part shift of a signed
//     value and the base subtraction of the old loop disappear: FFMA2 + FFMA + MUFU.RCP + FFMA2 per pixel.
//     (The coefficients carry a bias of -1/4 unit so that the rounding is a floor for every exactly representable
//     coordinate: integer and dyadic translations stay bit-exact; elsewhere the 16.16 value differs from the floor
//     by at most one unit = 1.5e-5 px, which the ambiguity band covers.)
//   * entries: per staged row 48 (R[x] R[x+1] G[x] G[x+1]) words followed by 48 (B[x] B[x+1] . .) words, row pitch
//     384 B = 3 x 128: lanes that sit on different source rows never collide on a bank, the conversion's 16-byte
//     stores are conflict-free, and the second tap row is the immediate offset +384.
//
// Strips that do not fit this scheme (partial strips, footprints wider than 44 entries or taller than 26 rows,
// SLOW / SPLIT geometry, the footprint holding source pixel (0,0), unaligned destinations, sources without a 3-D
// tensor map) are left to warp_pipe_kernel, tile by tile: the set-up kernel routes every 64 x 40 tile to exactly one
// of the two kernels (compact lists, counts on the device).
#pragma once

#include "warp_fast.cuh"

namespace b200geo {

#ifndef B2G_WS_H
#define B2G_WS_H 40
#endif
constexpr int WS_W = 32, WS_H = B2G_WS_H;    // strip size: a 64 x 40 tile = two strips side by side (WS_H = 40) or 2 x 2 strips (20)
static_assert(WF_TW == 2 * WS_W && (WF_TH == WS_H || WF_TH == 2 * WS_H), "a tile is two or four strips");
constexpr int WS_PER_TILE = 2 * (WF_TH / WS_H);
constexpr int WS_HALF = 20;                  // rows sampled per pass of the unrolled loop
constexpr int WS_EPR = 48;                   // entries per staged row
constexpr int WS_PLANE = WS_EPR * 4;         // byte offset of the B-pair words inside a staged row
constexpr int WS_ROWB = WS_EPR * 8;          // staged row pitch: 384 B
constexpr int WS_ROWS = WS_H == 40 ? 50 : 26; // rows of the tensor-copy box
constexpr int WS_RAW_W = 160;                // tensor-copy box: bytes per row (multiple of 16) ...
constexpr int WS_MAXP = 44;                  // ... which holds 4 + 12 + 3 * (44 + 1) <= 160 bytes
#ifndef B2G_WS_DIRECT
#define B2G_WS_DIRECT 1     // 1: sample straight from the raw rows of the tensor copy (no staged entries, no conversion); 0: staged entries
#endif
#ifndef B2G_WS_WARPS
#define B2G_WS_WARPS (B2G_WS_DIRECT ? (B2G_WS_H == 40 ? 14 : 20) : 16)
#endif
#ifndef B2G_WS_NRAW
#define B2G_WS_NRAW (B2G_WS_DIRECT ? 2 : 1)
#endif
static_assert(!B2G_WS_DIRECT || B2G_WS_NRAW >= 2, "direct sampling holds a raw buffer until the strip is sampled: a second one keeps a copy in flight");
constexpr int WS_WARPS = B2G_WS_WARPS;       // warps per CTA (one CTA per SM; shared memory is what limits them)
constexpr int WS_NRAW = B2G_WS_NRAW;         // raw-row buffers per warp = strips a tensor copy is issued ahead of its use
constexpr int WS_NREC = WS_NRAW + 1;         // record slots per warp
constexpr int WS_THREADS = WS_WARPS * 32;
constexpr int WS_RAW_BYTES = ((WS_RAW_W * WS_ROWS + 127) / 128) * 128;
constexpr int WS_ENT_BYTES = B2G_WS_DIRECT ? 0 : WS_ROWB * WS_ROWS;
constexpr int WS_WARP_BYTES = WS_NRAW * WS_RAW_BYTES + WS_ENT_BYTES;
static_assert(WS_WARP_BYTES % 128 == 0, "tensor-copy destinations are 128-byte aligned");
constexpr int WS_REC_BYTES = 80;
constexpr int WS_SMEM_BYTES = WS_WARPS * (WS_WARP_BYTES + WS_NREC * WS_REC_BYTES + 8 * WS_NRAW) + 16;
static_assert(WS_SMEM_BYTES <= 227 * 1024, "one CTA per SM");
constexpr float WS_MAGIC = 12582912.f;       // 1.5 * 2^23: float32(MAGIC + v) has the integer v in its low 23 bits for |v| < 2^22
constexpr int WS_MAGIC_HI = 0x4B40;          // bits 16.. of MAGIC's pattern: the offset of the 16.16 integer part
constexpr float WS_MAX_REL = 62.f;           // |rel| beyond this does not fit the 23-bit fixed point: old kernel

struct __align__(16) WsStrip {
    float nx0, ny0, ax, ay, bx, by, gw, hw;  // as WfDesc (x / y pairs adjacent: loaded as float2), n.. / a.. / b.. carrying the rounding bias
    int tma_x, tma_y;                        // tensor-copy coordinates: byte column (multiple of 16), row inside the image
    uint8_t *dst;                            // first output byte of the strip
    int dst_pitch;
    uint16_t img, job;                       // image index in the 3-D tensor map; job index
    int8_t ox, oy;                           // staged origin relative to the anchor
    uint8_t rows, chunks;                    // footprint rows, 4-entry chunks per row
    uint8_t mode, phase;                     // WF_ZERO | WF_INTERIOR | WF_BORDER; byte offset of entry column 0 inside a raw row (0, 4, 8, 12)
    uint16_t b16;                            // ambiguity half-width in units of 2^-16 px, rounded up (+2)
    uint16_t c0, r0;                         // first output pixel of the strip
    int xc, yc;                              // anchor (absolute source pixel)
    uint32_t magic;                          // tw | th << 8: the strip's extent (32, 40 unless cut by the output's edge)
};
static_assert(sizeof(WsStrip) == WS_REC_BYTES, "strip records are fetched as five 16-byte words");

// One strip: the analogue of wf_setup for a 32 x 20 strip.  Returns false if the strip must go to the old kernel.
template <int BILINEAR>
__device__ bool ws_setup(const b200geo_warp_job &J, int job, int c0, int r0, const WfTma &tma, uint8_t *__restrict__ dst_base, WsStrip &S)
{
    if (tma.image_rows <= 0) return false;                                    // needs the 3-D tensor map
    const int tw = min(WS_W, J.ow - c0), th = min(WS_H, J.oh - r0);           // < 32 / < 40: cut by the output's right / bottom edge
    if (tw <= 0 || th <= 0) {                                                 // the tile's second strip lies outside the output
        memset(&S, 0, sizeof S);
        S.mode = (uint8_t)WF_EMPTY;
        return true;
    }
    const int cm = c0 + WS_W / 2, rm = r0 + WS_H / 2;
    const double Xm = (double)cm * J.dx + J.x0, Ym = (double)rm * J.dy + J.y0;
    const double exl = (double)(J.ow - 1) * J.dx + J.x0, eyl = (double)(J.oh - 1) * J.dy + J.y0;
    bool ok = fabs(exl - J.x_last) <= 1e-9 * fmax(1.0, fabs(J.x_last)) && fabs(eyl - J.y_last) <= 1e-9 * fmax(1.0, fabs(J.y_last));
    ok = ok && J.bw >= 2 && J.bh >= 2 && J.bw < (1 << 22) && J.bh < (1 << 22) && J.iw >= 1 && J.ih >= 1;
    const double *h = J.invH;
    const double zx = h[0] * Xm + h[1] * Ym + h[2];
    const double zy = h[3] * Xm + h[4] * Ym + h[5];
    const double zw = h[6] * Xm + h[7] * Ym + h[8];
    ok = ok && fabs(zw) > 1e-200 && fabs(zw) < 1e200;
    if (!ok) return false;
    const double s = 1.0 / zw;
    const double xs = zx * s, ys = zy * s;
    if (!(fabs(xs) < 4.0e6 && fabs(ys) < 4.0e6)) return false;
    const double xc = rint(xs), yc = rint(ys);
    const double ax = (h[0] - xc * h[6]) * s * J.dx, bx = (h[1] - xc * h[7]) * s * J.dy;
    const double ay = (h[3] - yc * h[6]) * s * J.dx, by = (h[4] - yc * h[7]) * s * J.dy;
    const double gw = h[6] * s * J.dx, hw = h[7] * s * J.dy;
    const double q = fabs(gw) * (WS_W / 2) + fabs(hw) * (WS_H / 2);
    if (!(q <= 0.25)) return false;
    const double Sx = fabs(xs - xc) + fabs(ax) * (WS_W / 2) + fabs(bx) * (WS_H / 2);
    const double Sy = fabs(ys - yc) + fabs(ay) * (WS_W / 2) + fabs(by) * (WS_H / 2);
    if (!(Sx < 1.0e6 && Sy < 1.0e6)) return false;
    const double sc = 1.0 / 65536.0;
    WfDesc T;                                  // unbiased float32 coefficients: the corners through the pixels' formula
    T.nx0 = (float)(xs - xc); T.ax = (float)ax; T.bx = (float)bx;
    T.ny0 = (float)(ys - yc); T.ay = (float)ay; T.by = (float)by;
    T.gw = (float)(gw * sc); T.hw = (float)(hw * sc);
    T.xc = (int)xc; T.yc = (int)yc;
    float rxmin = 3.0e38f, rxmax = -3.0e38f, rymin = 3.0e38f, rymax = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float di = (float)(((k & 1) ? tw - 1 : 0) - WS_W / 2), dj = (float)(((k & 2) ? th - 1 : 0) - WS_H / 2);
        float rx, ry;
        wf_rel16(T, di, dj, rx, ry);
        rxmin = fminf(rxmin, rx); rxmax = fmaxf(rxmax, rx);
        rymin = fminf(rymin, ry); rymax = fmaxf(rymax, ry);
    }
    const float R = fmaxf(fmaxf(fabsf(rxmin), fabsf(rxmax)), fmaxf(fabsf(rymin), fabsf(rymax))) * (1.0f / 65536.f);
    if (!(R < WS_MAX_REL)) return false;
    // float32 error of rel as in wf_setup, plus two units of 2^-16 px for the biased round-to-nearest of the fixed point
    const float band = 2.4e-7f * (6.f * (float)fmax(Sx, Sy) + 10.f * R) + 4.0e-6f + 3.1e-5f;
    if (!(band <= WF_MAX_BAND)) return false;
    const float B = band * 65536.f;
    float lox, hix, loy, hiy;
    wf_bounds<BILINEAR>(T.xc, T.yc, J.bw, J.bh, lox, hix, loy, hiy);
    const float nn_lo = BILINEAR ? 0.f : 65536.f;
    const bool all_in = rxmin >= lox + nn_lo + B && rxmax <= hix - B && rymin >= loy + nn_lo + B && rymax <= hiy - B;
    const bool all_out = rxmax < lox - B || rxmin > hix + B || rymax < loy - B || rymin > hiy + B;
    // destination: 32-bit stores
    if (J.dst_pitch <= 0 || J.dst_pitch >= (1LL << 31)) return false;
    uint8_t *const dt = dst_base + J.dst_offset + (long long)r0 * J.dst_pitch + 3LL * c0;
    if ((((size_t)dt) | (size_t)J.dst_pitch) & 3) return false;
    if ((((size_t)dt) >> 32) != (((size_t)dt + (size_t)(th - 1) * (size_t)J.dst_pitch + 3 * tw) >> 32)) return false;   // row steps add to the low word only
    if (c0 > 65535 || r0 > 65535 || job > 65535) return false;
    S.dst = dt; S.dst_pitch = (int)J.dst_pitch; S.job = (uint16_t)job; S.c0 = (uint16_t)c0; S.r0 = (uint16_t)r0;
    S.xc = T.xc; S.yc = T.yc; S.b16 = (uint16_t)((int)B + 2);
    S.tma_x = S.tma_y = 0; S.img = 0; S.phase = 0; S.ox = S.oy = 0; S.rows = 0; S.chunks = 2; S.magic = (uint32_t)(tw | (th << 8));
    S.nx0 = S.ax = S.bx = S.ny0 = S.ay = S.by = S.gw = S.hw = 0.f;
    if (all_out) { S.mode = (uint8_t)WF_ZERO; return true; }
    int x_lo, x_hi, y_lo, y_hi;
    if (BILINEAR) {
        x_lo = __float2int_rd((rxmin - B) * (1.f / 65536.f)); x_hi = __float2int_rd((rxmax + B) * (1.f / 65536.f));
        y_lo = __float2int_rd((rymin - B) * (1.f / 65536.f)); y_hi = __float2int_rd((rymax + B) * (1.f / 65536.f)) + 1;
    } else {
        x_lo = __float2int_rd((rxmin - B) * (1.f / 65536.f) + 0.5f) - 1; x_hi = __float2int_rd((rxmax + B) * (1.f / 65536.f) + 0.5f) + 1;
        y_lo = __float2int_rd((rymin - B) * (1.f / 65536.f) + 0.5f) - 1; y_hi = __float2int_rd((rymax + B) * (1.f / 65536.f) + 0.5f) + 1;
    }
    if (!all_in) {      // clip to what a clamped coordinate can touch (see wf_setup)
        const int ix_lo = 0 - T.xc, ix_hi = J.bw - 1 - T.xc, iy_lo = 0 - T.yc, iy_hi = J.bh - 1 - T.yc, ytap = BILINEAR ? 1 : 0;
        if (x_hi < ix_lo) x_lo = x_hi = ix_lo;
        else if (x_lo > ix_hi) x_lo = x_hi = ix_hi;
        else { x_lo = max(x_lo, ix_lo); x_hi = min(x_hi, ix_hi); }
        if (y_hi < iy_lo) { y_lo = iy_lo; y_hi = iy_lo + ytap; }
        else if (y_lo > iy_hi) { y_lo = iy_hi; y_hi = iy_hi + ytap; }
        else { y_lo = max(y_lo, iy_lo); y_hi = min(y_hi, iy_hi + ytap); }
    }
    if (x_hi < x_lo || y_hi < y_lo) { x_hi = x_lo; y_hi = y_lo; }
    // staged origin: sampling straight from the raw rows works at any byte phase, so the footprint starts at its first pixel
    // (the staged-entries form converts in chunks of four pixels and needs a 4-pixel-aligned origin)
    const int ax0 = B2G_WS_DIRECT ? T.xc + x_lo : (T.xc + x_lo) & ~3;
    const int ox = ax0 - T.xc;
    const long long width = (long long)x_hi - ox + 1, rows = (long long)y_hi - y_lo + 1;
    const long long P = B2G_WS_DIRECT ? max(2LL, width) : max(8LL, (width + 3) & ~3LL);
    if (rows > WS_ROWS) return false;
    if (B2G_WS_DIRECT) {
        // the bytes of pixels ax0 .. ax0 + P (the right tap of the last pixel included) must lie inside the 160-byte box,
        // whose first byte is the 16-byte-aligned one at or before pixel ax0 (reads past them return bytes nobody uses)
        const long long ph = (3LL * ax0) & 15;
        if (ph + 3 * (P + 1) > WS_RAW_W) return false;
    } else if (P > WS_MAXP) return false;
    const int ay0 = T.yc + y_lo;
    // the one footprint per image that holds source pixel (0,0) (it must read as 0) stays with the old kernel
    const bool has00 = ax0 <= 0 && (long long)ax0 + P + 1 > 0 && ay0 <= 0 && (long long)ay0 + rows > 0;
    const bool img_ok = J.ih == tma.image_rows && 3LL * J.iw == tma.row_bytes && J.src_pitch == tma.pitch && J.src_offset >= 0 &&
                        J.src_offset % (tma.pitch * tma.image_rows) == 0;
    if (has00 || !img_ok || !(ax0 > -(1 << 24) && ay0 > -(1 << 24))) return false;
    const long long img = J.src_offset / (tma.pitch * tma.image_rows);
    if (ox < -128 || y_lo < -128 || ox > 127 || y_lo > 127 || img > 65535) return false;
    S.mode = all_in ? WF_INTERIOR : WF_BORDER;
    S.ox = (int8_t)ox; S.oy = (int8_t)y_lo; S.rows = (uint8_t)rows; S.chunks = (uint8_t)((P + 3) >> 2);
    S.tma_x = (3 * ax0) & ~15; S.phase = (uint8_t)((3 * ax0) & 15); S.tma_y = ay0;
    S.img = (uint16_t)img;
    // biased coefficients: n' = n + bias * w, so that  round(n' / w) = floor(n / w [+ 1/2 px for nn])  up to a quarter unit
    const double bias = BILINEAR ? -0.25 : 32768.0 - 0.25;
    S.nx0 = (float)((xs - xc) + bias * sc); S.ax = (float)(ax + bias * gw * sc); S.bx = (float)(bx + bias * hw * sc);
    S.ny0 = (float)((ys - yc) + bias * sc); S.ay = (float)(ay + bias * gw * sc); S.by = (float)(by + bias * hw * sc);
    S.gw = T.gw; S.hw = T.hw;
    return true;
}

// Set-up of the routed launch.  One thread per 64 x 40 tile: the four strip records if all four strips qualify (appended to
// `strips`, count in counts[0]), else the tile's WpTile record (appended to `tiles`, count in counts[1]).
// Appends are aggregated per warp (one atomic per warp and list).
template <int BILINEAR>
__global__ void __launch_bounds__(256)
warp_route_setup_kernel(const b200geo_warp_job *__restrict__ jobs, const uint8_t *__restrict__ src_base, uint8_t *__restrict__ dst_base,
                        const WfTma tma, int use_tmap, int tiles_x, int tiles_y, WsStrip *__restrict__ strips, WpTile *__restrict__ tiles,
                        unsigned *__restrict__ counts)
{
    const int tx = blockIdx.x * 32 + threadIdx.x, ty = blockIdx.y * 8 + threadIdx.y;
    const int lane = threadIdx.x;
    bool live = tx < tiles_x && ty < tiles_y;
    const b200geo_warp_job &J = jobs[blockIdx.z];
    const int c0 = tx * WF_TW, r0 = ty * WF_TH;
    if (live && (c0 >= J.ow || r0 >= J.oh || !wf_job_ok(J))) live = false;
    WsStrip S[WS_PER_TILE];
    bool fast = live;
    if (fast) {
#pragma unroll
        for (int k = 0; k < WS_PER_TILE; ++k)
            if (fast) fast = ws_setup<BILINEAR>(J, (int)blockIdx.z, c0 + (k & 1) * WS_W, r0 + (k >> 1) * WS_H, tma, dst_base, S[k]);
    }
    const unsigned m_fast = __ballot_sync(0xffffffffu, fast), m_old = __ballot_sync(0xffffffffu, live && !fast);
    unsigned b_fast = 0, b_old = 0;
    if (lane == 0) {
        if (m_fast) b_fast = atomicAdd(counts + 0, (unsigned)__popc(m_fast));
        if (m_old) b_old = atomicAdd(counts + 1, (unsigned)__popc(m_old));
    }
    b_fast = __shfl_sync(0xffffffffu, b_fast, 0);
    b_old = __shfl_sync(0xffffffffu, b_old, 0);
    const unsigned below = (1u << lane) - 1u;
    if (fast) {
        uint4 *d = reinterpret_cast<uint4 *>(strips + WS_PER_TILE * (size_t)(b_fast + __popc(m_fast & below)));
        const uint4 *sv = reinterpret_cast<const uint4 *>(S);
#pragma unroll
        for (int i = 0; i < WS_PER_TILE * WS_REC_BYTES / 16; ++i) d[i] = sv[i];
    } else if (live) {
        WpTile W;
        wp_build_tile<BILINEAR>(J, (int)blockIdx.z, c0, r0, src_base, dst_base, tma, use_tmap, W);
        uint4 *d = reinterpret_cast<uint4 *>(tiles + (size_t)(b_old + __popc(m_old & below)));
        const uint4 *sv = reinterpret_cast<const uint4 *>(&W);
#pragma unroll
        for (int i = 0; i < WP_TILE_WORDS; ++i) d[i] = sv[i];
    }
}

__device__ __forceinline__ uint32_t ws_lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float ws_ldsf(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t ws_ldsu8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int ws_ldss8(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t ws_ldsu16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t ws_lds32o(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ void ws_stg32(uint8_t *p, uint32_t v)
{
    asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#define WS_REC(field) ((uint32_t)offsetof(WsStrip, field))

// The 80-byte record as the kernel holds it in registers (four / five 16-byte shared-memory loads per strip: the data pipe
// of the SM is the kernel's limiter, and twenty scalar loads per strip were 8 % of its wavefronts).
struct WsRegs {
    uint4 w0, w1, w2, w3, w4;  // (nx0 ny0 ax ay) (bx by gw hw) (tma_x tma_y dst) (dst_pitch img|job ox|oy|rows|chunks mode|phase|b16) (c0|r0 xc yc magic)
    __device__ __forceinline__ int rows() const { return (int)((w3.z >> 16) & 0xFFu); }
    __device__ __forceinline__ int chunks() const { return (int)(w3.z >> 24); }
    __device__ __forceinline__ int ox() const { return (int)(int8_t)(w3.z & 0xFFu); }
    __device__ __forceinline__ int oy() const { return (int)(int8_t)((w3.z >> 8) & 0xFFu); }
    __device__ __forceinline__ int mode() const { return (int)(w3.w & 0xFFu); }
    __device__ __forceinline__ uint32_t phase() const { return (w3.w >> 8) & 0xFFu; }
    __device__ __forceinline__ uint32_t b16() const { return w3.w >> 16; }
    __device__ __forceinline__ uint32_t job() const { return w3.y >> 16; }
    __device__ __forceinline__ uint32_t img() const { return w3.y & 0xFFFFu; }
};
static_assert(offsetof(WsStrip, bx) == 16 && offsetof(WsStrip, tma_x) == 32 && offsetof(WsStrip, dst) == 40 && offsetof(WsStrip, dst_pitch) == 48 &&
              offsetof(WsStrip, img) == 52 && offsetof(WsStrip, job) == 54 && offsetof(WsStrip, ox) == 56 && offsetof(WsStrip, oy) == 57 &&
              offsetof(WsStrip, rows) == 58 && offsetof(WsStrip, chunks) == 59 && offsetof(WsStrip, mode) == 60 && offsetof(WsStrip, phase) == 61 &&
              offsetof(WsStrip, b16) == 62 && offsetof(WsStrip, c0) == 64 && offsetof(WsStrip, xc) == 68 && offsetof(WsStrip, magic) == 76,
              "WsRegs decodes the record by position");
__device__ __forceinline__ uint4 ws_lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

__device__ __forceinline__ void ws_convert_chunk(uint32_t a, uint32_t d)
{
    const uint32_t w0 = ws_lds32(a), w1 = ws_lds32o<4>(a), w2 = ws_lds32o<8>(a), w3 = ws_lds32o<12>(a);
    uint4 rg, bb;
    rg.x = __byte_perm(w0, w1, 0x4130); bb.x = __byte_perm(w0, w1, 0x5252);
    rg.y = __byte_perm(w0, w1, 0x7463); bb.y = __byte_perm(w1, w2, 0x4141);
    rg.z = __byte_perm(w1, w2, 0x6352); bb.z = __byte_perm(w2, w2, 0x3030);
    rg.w = __byte_perm(w2, w3, 0x5241); bb.w = __byte_perm(w2, w3, 0x6363);
    wf_sts128(d, rg);
    wf_sts128(d + WS_PLANE, bb);
}

// raw RGB rows (tensor copy: row pitch WS_RAW_W, entry column 0 at byte `phase`) -> staged entries of this warp.
// A chunk = four entries = 12 raw bytes -> 16 bytes per plane.  Phase A: chunks 0..7 of every row, eight lanes per row (each
// quarter-warp writes 128 contiguous bytes of one row: no bank conflict; the flattened row-major order, whose quarter-warps
// straddle rows 384 = 3 x 128 bytes apart, measured 2-way conflicts on every store).  Phase B: the one to three chunks
// beyond the eighth, flattened over the rows.
__device__ __forceinline__ void ws_convert(const WsRegs &R, uint32_t s_raw, uint32_t s_ent, int lane)
{
    const int rows = R.rows(), chunks = R.chunks();
    const uint32_t phase = R.phase();
    {
        const int k = lane & 7;
        uint32_t ra = s_raw + phase + (uint32_t)(lane >> 3) * WS_RAW_W + 12u * (uint32_t)k;
        uint32_t ea = s_ent + (uint32_t)(lane >> 3) * WS_ROWB + 16u * (uint32_t)k;
        if (k < chunks) {
#pragma unroll 2
            for (int row = lane >> 3; row < rows; row += 4) {
                ws_convert_chunk(ra, ea);
                ra += 4u * WS_RAW_W; ea += 4u * WS_ROWB;
            }
        }
    }
    const int cb = chunks - 8;
    if (cb > 0) {          // (warp-uniform)
        const int total = rows * cb;
#pragma unroll 1
        for (int t = lane; t < total; t += 32) {
            const int i = cb == 1 ? t : (cb == 2 ? t >> 1 : (int)(((unsigned)t * 43691u) >> 17));     // t / cb, cb <= 3
            const int k = 8 + t - i * cb;
            ws_convert_chunk(s_raw + phase + (uint32_t)i * WS_RAW_W + 12u * (uint32_t)k, s_ent + (uint32_t)i * WS_ROWB + 16u * (uint32_t)k);
        }
    }
}

#ifndef B2G_WS_GROUP
#define B2G_WS_GROUP 2      // rows of a lane whose taps are in flight together (2 measured best: 576 vs 569 (4), 556 (5) Gpix/s)
#endif
#ifndef B2G_WS_UNROLL
#define B2G_WS_UNROLL 1     // 1: the 20 rows of an INTERIOR strip fully unrolled (row index = immediate operand)
#endif

__device__ __forceinline__ float2 ws_ldsf2(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
// opaque to the optimiser's algebra: keeps "integer part, then scaled add" as SHF + IMAD / LEA
__device__ __forceinline__ uint32_t ws_hi16(uint32_t v)
{
    uint32_t r;
    asm("shr.u32 %0, %1, 16;" : "=r"(r) : "r"(v));
    return r;
}
// 32-bit store through a pointer given as (low word, high word): the row step touches the low word only (the set-up
// kernel routes a strip here only if its 20 rows do not cross a 4 GiB boundary)
__device__ __forceinline__ void ws_stg32_lh(uint32_t lo, uint32_t hi, uint32_t v)
{
    asm volatile("{\n\t.reg .b64 a;\n\tmov.b64 a, {%0, %1};\n\tst.global.u32 [a], %2;\n\t}" ::"r"(lo), "r"(hi), "r"(v) : "memory");
}

// One strip: 40 pixels per lane (column c0 + lane, rows r0 .. r0+39), in two passes of 20 rows.
// BORDER = 0: wholly unmasked full strip (rows unrolled); 1: per-pixel mask (the strip crosses the sampler's bounds);
// 2: as 1 for a strip cut by the output's right / bottom edge: lanes >= tw and rows >= th repeat the last valid position
// (every address stays inside the footprint) and store nothing; the valid pixels are stored byte by byte.
template <int BILINEAR, int BORDER>
__device__ __forceinline__ void ws_sample(const WsRegs &R, uint32_t s_ent, int lane, const b200geo_warp_job *__restrict__ jobs,
                                          const uint8_t *__restrict__ src_base, const WfFix &fix0, unsigned &n_exact)
{
    const float2 A2 = make_float2(__uint_as_float(R.w0.z), __uint_as_float(R.w0.w));
    const float2 B2 = make_float2(__uint_as_float(R.w1.x), __uint_as_float(R.w1.y));
    const float gw = __uint_as_float(R.w1.z), hw = __uint_as_float(R.w1.w);
    const int tw = BORDER == 2 ? (int)(R.w4.w & 0xFFu) : WS_W, th = BORDER == 2 ? (int)((R.w4.w >> 8) & 0xFFu) : WS_H;
    const float di0 = (float)(min(lane, tw - 1) - WS_W / 2);
    const float2 nb = __ffma2_rn(A2, make_float2(di0, di0), make_float2(__uint_as_float(R.w0.x), __uint_as_float(R.w0.y)));
    const float wb = __fmaf_rn(gw, di0, WF_W0);
    const int ox = R.ox(), oy = R.oy();
    // DIRECT: byte address of R[xi] in raw row yi = yi * 160 + 3 xi + base; else plane-0 word of entry (xi, yi)
    // [xi, yi: bit patterns of the magic-number fixed point, integer parts offset by WS_MAGIC_HI]
    const uint32_t base = B2G_WS_DIRECT ? s_ent + R.phase() - (uint32_t)((WS_MAGIC_HI + oy) * WS_RAW_W + (WS_MAGIC_HI + ox) * 3)
                                        : s_ent - (uint32_t)((WS_MAGIC_HI + oy) * WS_ROWB + (WS_MAGIC_HI + ox) * 4);
    const unsigned long long dstrip = ((unsigned long long)R.w2.w << 32) | R.w2.z;      // first output byte of the strip
    // word (lane & 3) of the 12 bytes of pixels 4 * (lane >> 2) .. + 3:  RGBR | GBRG | BRGB  (three of four lanes store)
    uint32_t g_lo = (uint32_t)dstrip + (uint32_t)(12 * (lane >> 2) + 4 * (lane & 3));
    const uint32_t g_hi = (uint32_t)(dstrip >> 32);
    const uint32_t pitch = R.w3.x;
    const uint32_t psel = (lane & 3) == 0 ? 0x4210u : ((lane & 3) == 1 ? 0x5421u : 0x6542u);
    const bool st_lane = (lane & 3) != 3;
    float lox = 0.f, hix = 0.f, loy = 0.f, hiy = 0.f, band = 0.f;
    int xlo = 0, xhi = 0, ylo = 0, yhi = 0;
    uint32_t B16 = 0u;
    if (BORDER || !BILINEAR) {
        B16 = R.b16();
        band = (float)B16;
    }
    if (BORDER) {
        const b200geo_warp_job &J = jobs[R.job()];
        const int xc = (int)R.w4.y, yc = (int)R.w4.z;
        wf_bounds<BILINEAR>(xc, yc, J.bw, J.bh, lox, hix, loy, hiy);
        xlo = -xc; xhi = J.bw - 1 - xc; ylo = -yc; yhi = J.bh - 1 - yc;
        if (!BILINEAR) {     // the biased coordinate carries + 1/2 px
            lox += 32768.f; hix += 32768.f; loy += 32768.f; hiy += 32768.f;
        }
    }
    constexpr int G = B2G_WS_GROUP;
    static_assert(WS_HALF % G == 0 && WS_H % WS_HALF == 0, "row groups");
#pragma unroll 1
    for (int hf = 0; hf < WS_H / WS_HALF; ++hf) {             // (the 'g0' below count rows inside this pass)
    if (BORDER == 2 && hf * WS_HALF >= th) break;
    const float djh = (float)(hf * WS_HALF - WS_H / 2);
    const float2 nbh = __ffma2_rn(B2, make_float2(djh, djh), nb);
    const float wbh = __fmaf_rn(hw, djh, wb);
    uint32_t amb = 0u, inmask = 0u;
    auto group = [&](const int g0, const float2 n2g, const float wg) {
        // n2g, wg: numerators / denominator of row g0 (for the looped form: of row g0, the rows of the group add constants)
        uint32_t t0[G], t1[G], t2[G], t3[G], bx_[G], by_[G];
        uint32_t u0[G], u1[G], u2[G], v0[G], v1[G], v2[G], sh_[G];     // DIRECT: three raw words per tap row, byte phase * 8
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float dj = (float)g;
            float2 n2 = g ? __ffma2_rn(B2, make_float2(dj, dj), n2g) : n2g;
            float w = g ? __fmaf_rn(hw, dj, wg) : wg;
            bool valid = true;
            if (BORDER == 2) {          // rows past the strip's last one repeat it
                const int row = hf * WS_HALF + g0 + g;
                valid = row < th && lane < tw;
                const float rj = (float)(min(row, th - 1) - hf * WS_HALF);
                n2 = __ffma2_rn(B2, make_float2(rj, rj), nbh);
                w = __fmaf_rn(hw, rj, wbh);
            }
            const float rw = wf_rcp(w);
            uint32_t bxv, byv;
            if (BORDER) {
                float2 r = __fmul2_rn(n2, make_float2(rw, rw));
                const bool in = r.x >= lox + band && r.x <= hix - band && r.y >= loy + band && r.y <= hiy - band;
                const bool outp = r.x < lox - band || r.x > hix + band || r.y < loy - band || r.y > hiy + band;
                if (in && valid) inmask |= 1u << (g0 + g);
                else if (!outp && valid) amb |= 1u << (g0 + g);        // also NaN coordinates
                r.x = fminf(fmaxf(r.x, lox), hix);                     // (NaN -> lox)
                r.y = fminf(fmaxf(r.y, loy), hiy);
                const float2 f = __fadd2_rn(r, make_float2(WS_MAGIC, WS_MAGIC));
                bxv = __float_as_uint(f.x); byv = __float_as_uint(f.y);
            } else {
                const float2 f = __ffma2_rn(n2, make_float2(rw, rw), make_float2(WS_MAGIC, WS_MAGIC));
                bxv = __float_as_uint(f.x); byv = __float_as_uint(f.y);
            }
            uint32_t xi = 0u, yi = 0u;
            if (!(B2G_WS_DIRECT && !BORDER)) { xi = ws_hi16(bxv); yi = ws_hi16(byv); }
            if (!BILINEAR) {
                const uint32_t fx = bxv & 0xFFFFu, fy = byv & 0xFFFFu;
                const bool near = fx - B16 > 65535u - 2u * B16 || fy - B16 > 65535u - 2u * B16;
                if (BORDER) {
                    if (near && (inmask >> (g0 + g) & 1u)) amb |= 1u << (g0 + g);      // (inmask is only set for valid pixels)
                    xi = (uint32_t)(min(max((int)xi - WS_MAGIC_HI, xlo), xhi) + WS_MAGIC_HI);
                    yi = (uint32_t)(min(max((int)yi - WS_MAGIC_HI, ylo), yhi) + WS_MAGIC_HI);
                } else if (near) amb |= 1u << (g0 + g);
            }
            if (B2G_WS_DIRECT) {
                // the six bytes of pixels xi, xi+1 start at any byte phase: three aligned words cover them; the funnel shifts
                // below (shift count taken modulo 32: 8 x the byte phase) bring them to (R0 G0 B0 R1) (G1 B1 . .)
                // byte address of R[xi] in row yi: yi * 160 + xi * 3 + base -- one two-way dot product on the packed integer
                // parts (IDP.2A: u16 x u8) when they come straight from the fixed-point patterns
                uint32_t a;
                if (!BORDER) a = __dp2a_lo(__byte_perm(bxv, byv, 0x7632), (uint32_t)(3 | (WS_RAW_W << 8)), base);
                else a = yi * (uint32_t)WS_RAW_W + (xi * 3u + base);
                const uint32_t W = a & ~3u;
                sh_[g] = a << 3;
                u0[g] = ws_lds32(W); u1[g] = ws_lds32o<4>(W);
                if (BILINEAR) {
                    u2[g] = ws_lds32o<8>(W);
                    v0[g] = ws_lds32o<WS_RAW_W>(W); v1[g] = ws_lds32o<WS_RAW_W + 4>(W); v2[g] = ws_lds32o<WS_RAW_W + 8>(W);
                }
            } else {
                const uint32_t a = yi * (uint32_t)WS_ROWB + (xi * 4u + base);
                t0[g] = ws_lds32(a);
                t1[g] = ws_lds32o<WS_PLANE>(a);
                if (BILINEAR) {
                    t2[g] = ws_lds32o<WS_ROWB>(a);
                    t3[g] = ws_lds32o<WS_ROWB + WS_PLANE>(a);
                }
            }
            bx_[g] = bxv; by_[g] = byv;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            uint32_t out;
            if (B2G_WS_DIRECT) {
                const uint32_t p0 = __funnelshift_r(u0[g], u1[g], sh_[g]);
                if (BILINEAR) {
                    const uint32_t p1 = __funnelshift_r(u1[g], u2[g], sh_[g]);
                    const uint32_t q0 = __funnelshift_r(v0[g], v1[g], sh_[g]), q1 = __funnelshift_r(v1[g], v2[g], sh_[g]);
                    t0[g] = __byte_perm(p0, p1, 0x4130); t1[g] = __byte_perm(p0, p1, 0x5252);
                    t2[g] = __byte_perm(q0, q1, 0x4130); t3[g] = __byte_perm(q0, q1, 0x5252);
                } else out = p0;                                      // (R G B .): byte 3 is never stored
            }
            if (BILINEAR) out = wf_blend(make_uint2(t0[g], t1[g]), make_uint2(t2[g], t3[g]), bx_[g] & 0xFFFFu, by_[g] & 0xFFFFu);
            else if (!B2G_WS_DIRECT) out = __byte_perm(t0[g], t1[g], 0x4420);
            if (BORDER && !(inmask >> (g0 + g) & 1u)) out = 0u;
            if (BORDER == 2) {
                if (lane < tw && hf * WS_HALF + g0 + g < th) {
                    uint8_t *p = reinterpret_cast<uint8_t *>(dstrip) + (long long)(hf * WS_HALF + g0 + g) * (int)pitch + 3 * lane;
                    p[0] = (uint8_t)out; p[1] = (uint8_t)(out >> 8); p[2] = (uint8_t)(out >> 16);
                }
            } else {
                const uint32_t nbr = __shfl_down_sync(0xffffffffu, out, 1);
                const uint32_t word = __byte_perm(out, nbr, psel);
                if (st_lane) ws_stg32_lh(g_lo, g_hi, word);
                g_lo += pitch;
            }
        }
    };
    if (B2G_WS_UNROLL && !BORDER) {
#pragma unroll
        for (int g0 = 0; g0 < WS_HALF; g0 += G) {
            const float dj = (float)g0;
            group(g0, g0 ? __ffma2_rn(B2, make_float2(dj, dj), nbh) : nbh, g0 ? __fmaf_rn(hw, dj, wbh) : wbh);
        }
    } else {
        float dj = 0.f;
#pragma unroll 1
        for (int g0 = 0; g0 < WS_HALF; g0 += G) {
            group(g0, __ffma2_rn(B2, make_float2(dj, dj), nbh), __fmaf_rn(hw, dj, wbh));
            dj += (float)G;
        }
    }
    if (BORDER || !BILINEAR) {
        if (__any_sync(0xffffffffu, amb != 0u)) {
            __syncwarp();                                   // every lane's word stores precede the byte stores below
            if (amb) {
                const int job = (int)R.job();
                const int c = (int)(R.w4.x & 0xFFFFu) + lane, r0 = (int)(R.w4.x >> 16);
                WfFix fix = fix0;
                fix.job = job;
#pragma unroll 1
                for (uint32_t rest = amb; rest; rest &= rest - 1u) {
                    const int k = __ffs((int)rest) - 1 + hf * WS_HALF;
                    if (!wf_defer(fix, c, r0 + k)) {
                        const b200geo_warp_job &J = jobs[job];
                        const uint32_t v = wf_exact_px<BILINEAR>(J, src_base + J.src_offset, c, r0 + k);
                        ++n_exact;
                        uint8_t *p = reinterpret_cast<uint8_t *>(dstrip) + (long long)k * (int)pitch + 3 * lane;
                        p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
                    }
                }
            }
        }
    }
    }   // hf
}

// Grid: min(ceil(n_tiles / WS_WARPS), SMs) CTAs of WS_WARPS warps, dynamic shared memory WS_SMEM_BYTES (one CTA per SM).
// counts[0] = number of routed tiles (four strip records each); ticket: global tile-ticket counter (zeroed by the host).
//
// Per warp, strip n of its sequence (raw buffer n % WS_NRAW, record slot n % WS_NREC):
//     record -> registers; wait for the tensor copy of strip n; convert raw rows -> entries; the raw buffer is free:
//     issue the tensor copy of strip n + WS_NRAW into it; fetch the record of strip n + WS_NRAW + 1 (cp.async) into the
//     slot of strip n; sample strip n from the entries.
// A tensor copy therefore has WS_NRAW - 1 whole strip periods plus one sampling phase to arrive (measured with one buffer
// per warp: a strip period of ~6000 cycles of which ~2500 were the wait for the copy -- nn ran no faster than bilinear).
template <int BILINEAR>
__global__ void __launch_bounds__(WS_THREADS, 1)
warp_strip_kernel(const uint8_t *__restrict__ src_base, const b200geo_warp_job *__restrict__ jobs, const WsStrip *__restrict__ strips,
                  const unsigned *__restrict__ counts, const __grid_constant__ CUtensorMap tmap, WfFixEntry *__restrict__ fix_list,
                  unsigned *__restrict__ fix_counts, unsigned seg_cap, unsigned *__restrict__ ticket, unsigned long long *__restrict__ stats)
{
    extern __shared__ __align__(128) unsigned char ws_smem[];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(ws_smem);
    const uint32_t s_raw0 = s0 + (uint32_t)(wrp * WS_WARP_BYTES), s_ent = s_raw0 + WS_NRAW * WS_RAW_BYTES;
    const uint32_t s_recs = s0 + (uint32_t)(WS_WARPS * WS_WARP_BYTES) + (uint32_t)(wrp * WS_NREC * WS_REC_BYTES);
    const uint32_t s_bar0 = s0 + (uint32_t)(WS_WARPS * (WS_WARP_BYTES + WS_NREC * WS_REC_BYTES)) + 8u * (uint32_t)(WS_NRAW * wrp);
    unsigned *const s_fix_n = reinterpret_cast<unsigned *>(ws_smem + WS_WARPS * (WS_WARP_BYTES + WS_NREC * WS_REC_BYTES + 8 * WS_NRAW));
    if (threadIdx.x == 0) *s_fix_n = 0u;
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < WS_NRAW; ++b) wp_mbar_init(s_bar0 + 8u * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const unsigned n_tiles = counts[0];
    const unsigned n_warps = gridDim.x * WS_WARPS;
    unsigned n_exact = 0;
    const WfFix fix = {fix_list + (size_t)blockIdx.x * seg_cap, s_fix_n, seg_cap, 0};

    // the warp's strip sequence: four strips per tile, tiles by ticket (the first one is the warp's global index)
    unsigned g_tile = blockIdx.x * WS_WARPS + wrp;
    int g_j = -1;
    unsigned g_tk = 0;                                            // lane 0: ticket of the tile after g_tile
    constexpr unsigned NONE = 0xFFFFFFFFu;
    auto next_strip = [&]() -> unsigned {                         // warp-uniform
        if (g_j < 0) {
            if (g_tile >= n_tiles) return NONE;
            g_j = 0;
            if (lane == 0) g_tk = atomicAdd(ticket, 1u);
            return g_tile * (unsigned)WS_PER_TILE;
        }
        if (g_j < WS_PER_TILE - 1) return g_tile * (unsigned)WS_PER_TILE + (unsigned)(++g_j);
        if (g_tile == NONE) return NONE;
        g_tile = __shfl_sync(0xffffffffu, g_tk, 0) + n_warps;
        if (g_tile >= n_tiles) { g_tile = NONE; return NONE; }
        g_j = 0;
        if (lane == 0) g_tk = atomicAdd(ticket, 1u);
        return g_tile * (unsigned)WS_PER_TILE;
    };
    auto fetch_record = [&](uint32_t slot, unsigned strip) {      // asynchronous: lanes 0..4 copy 16 bytes each
        if (strip != NONE && lane < WS_REC_BYTES / 16) {
            const char *g = reinterpret_cast<const char *>(strips + strip) + 16 * lane;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + 16u * (uint32_t)lane), "l"(g) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto issue_copy = [&](uint32_t slot, uint32_t raw, uint32_t bar) {   // lane 0: tensor copy of the strip's footprint into a raw buffer of this warp
        const int m_ = (int)ws_ldsu8(slot + WS_REC(mode));
        if (lane == 0 && (m_ == WF_INTERIOR || m_ == WF_BORDER)) {
            uint32_t x, y;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(slot + WS_REC(tma_x)));
            const int z = (int)ws_ldsu16(slot + WS_REC(img));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // this warp's generic reads of the raw rows precede the overwrite
            wp_mbar_expect_tx(bar, (uint32_t)(WS_RAW_W * WS_ROWS));
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(raw), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((int)x), "r"((int)y), "r"(z), "r"(bar) : "memory");
        }
    };

    // prologue: records of strips 0 .. WS_NRAW, tensor copies of strips 0 .. WS_NRAW - 1
    unsigned ids[WS_NREC];                                        // ids[k]: strip n + k (constant indices only: registers)
#pragma unroll
    for (int k = 0; k < WS_NREC; ++k) {
        ids[k] = next_strip();
        fetch_record(s_recs + (uint32_t)(k * WS_REC_BYTES), ids[k]);
    }
    if (ids[0] != NONE) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < WS_NRAW; ++k)
            if (ids[k] != NONE) issue_copy(s_recs + (uint32_t)(k * WS_REC_BYTES), s_raw0 + (uint32_t)(k * WS_RAW_BYTES), s_bar0 + 8u * k);
        uint32_t ph = 0u;                                         // bit b: parity of raw buffer b's mbarrier
        uint32_t slot = s_recs;                                   // record slot of strip n
        int buf = 0;                                              // raw buffer of strip n
        for (;;) {
            WsRegs R;
            R.w0 = ws_lds128(slot); R.w1 = ws_lds128(slot + 16u); R.w2 = ws_lds128(slot + 32u); R.w3 = ws_lds128(slot + 48u);
            R.w4 = ws_lds128(slot + 64u);
            const int mode = R.mode();
            const uint32_t raw = s_raw0 + (uint32_t)(buf * WS_RAW_BYTES), bar = s_bar0 + 8u * (uint32_t)buf;
            if (mode == WF_INTERIOR || mode == WF_BORDER) {
                wp_mbar_wait(bar, (ph >> buf) & 1u);
                ph ^= 1u << buf;
                if (!B2G_WS_DIRECT) ws_convert(R, raw, s_ent, lane);
            }
            auto refill = [&]() {
                // strip n + WS_NRAW: its record arrived during the previous iteration; its copy goes into the buffer just freed;
                // the record of strip n + WS_NRAW + 1 goes into the slot of strip n (held in registers by now)
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                if (ids[WS_NRAW] != NONE) {
                    uint32_t sl = slot + (uint32_t)(WS_NRAW * WS_REC_BYTES);
                    if (sl >= s_recs + (uint32_t)(WS_NREC * WS_REC_BYTES)) sl -= (uint32_t)(WS_NREC * WS_REC_BYTES);
                    issue_copy(sl, raw, bar);
                }
            };
            if (!B2G_WS_DIRECT) {
                __syncwarp();                                     // entries written, raw rows read by every lane
                refill();
            }
            const uint32_t s_src = B2G_WS_DIRECT ? raw : s_ent;
            // (votes: the branch conditions are warp-uniform by construction, so the shuffles inside need no divergence check)
            const bool partial = (R.w4.w & 0xFFFFu) != (uint32_t)(WS_W | (WS_H << 8));
            if (__all_sync(0xffffffffu, (mode == WF_INTERIOR || mode == WF_BORDER) && partial)) ws_sample<BILINEAR, 2>(R, s_src, lane, jobs, src_base, fix, n_exact);
            else if (__all_sync(0xffffffffu, mode == WF_INTERIOR)) ws_sample<BILINEAR, 0>(R, s_src, lane, jobs, src_base, fix, n_exact);
            else if (__all_sync(0xffffffffu, mode == WF_BORDER)) ws_sample<BILINEAR, 1>(R, s_src, lane, jobs, src_base, fix, n_exact);
            else if (mode == WF_ZERO) {
                uint8_t *gp = reinterpret_cast<uint8_t *>(((unsigned long long)R.w2.w << 32) | R.w2.z);
                const int pitch = (int)R.w3.x, tw = (int)(R.w4.w & 0xFFu), th = (int)((R.w4.w >> 8) & 0xFFu);
                if (!partial) {
                    if (lane < 24) {
#pragma unroll 4
                        for (int r = 0; r < WS_H; ++r) ws_stg32(gp + 4 * lane + (long long)r * pitch, 0u);
                    }
                } else {
                    for (int r = 0; r < th; ++r)
                        for (int i = lane; i < 3 * tw; i += 32) gp[(long long)r * pitch + i] = 0;
                }
            }
            if (B2G_WS_DIRECT) {
                __syncwarp();                                     // raw rows read by every lane
                refill();
            }
            const unsigned id_new = ids[WS_NRAW] != NONE ? next_strip() : NONE;      // strip n + WS_NRAW + 1 -> the slot of strip n
            fetch_record(slot, id_new);
#pragma unroll
            for (int k = 0; k < WS_NRAW; ++k) ids[k] = ids[k + 1];
            ids[WS_NRAW] = id_new;
            if (ids[0] == NONE) break;
            __syncwarp();                                         // every lane is done with the entries
            slot += WS_REC_BYTES;
            if (slot >= s_recs + (uint32_t)(WS_NREC * WS_REC_BYTES)) slot = s_recs;
            buf = buf + 1 == WS_NRAW ? 0 : buf + 1;
        }
    }
    if (stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
    __syncthreads();                                              // every warp has pushed its last deferred pixel
    if (threadIdx.x == 0) fix_counts[blockIdx.x] = *s_fix_n;
}

}  // namespace b200geo
