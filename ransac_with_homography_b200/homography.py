"""Drop-in for the reference's ``homography`` module: same names, positional order, defaults and
return types (reference: homography.py), with the arithmetic done by the sm_100a kernels of
libb200geo.so through its C ABI (include/b200geo.h).  No CPU fallback: every solve / warp below
runs on the GPU or raises.

Host code here only does what the reference's callers observe around the kernels: argument
checks, the O(1) output-grid set-up (bounding box of four warped corners, ``np.linspace``
parameters), the in-place zeroing of ``img[0,0,:]`` the reference's samplers perform on the
caller's array, and the final crop.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi

__all__ = [
    "calc_corresp", "calc_correspLinear", "calc_correspCollective", "calc_correspLinearCollective",
    "calcHomography", "calcHomographyLinear", "nearestNeighbor", "bilinear", "convertfunc",
    "wrapPerspective", "wrapPerspectiveScan", "transformImage", "transformImageH",
    "BLENDDIR", "addAlpha", "stitchPanorama", "cylindericalTransform", "cylindricalWarp", "cylindericlMap",
]

_F32 = np.float32


# ------------------------------------------------------------------------------------------------
# DLT design matrices (reference homography.py:4-69).  On the GPU path these matrices are never
# materialised (the kernels build them in registers); the builders are kept because they are
# importable names of the reference module.  Pure layout code.
# ------------------------------------------------------------------------------------------------
def _null_rows(u, v):
    u, v = np.asarray(u), np.asarray(v)
    n = u.shape[0]
    m = np.zeros((n, 2, 9), dtype=_F32)
    m[:, 0, 0], m[:, 0, 1], m[:, 0, 2] = -u[:, 0], -u[:, 1], -1
    m[:, 0, 6], m[:, 0, 7], m[:, 0, 8] = u[:, 0] * v[:, 0], u[:, 1] * v[:, 0], v[:, 0]
    m[:, 1, 3], m[:, 1, 4], m[:, 1, 5] = -u[:, 0], -u[:, 1], -1
    m[:, 1, 6], m[:, 1, 7], m[:, 1, 8] = u[:, 0] * v[:, 1], u[:, 1] * v[:, 1], v[:, 1]
    return m.reshape(2 * n, 9)


def _inhom_rows(u, v):
    u, v = np.asarray(u), np.asarray(v)
    n = u.shape[0]
    a = np.zeros((n, 2, 8), dtype=_F32)
    a[:, 0, 0], a[:, 0, 1], a[:, 0, 2] = u[:, 0], u[:, 1], 1
    a[:, 0, 6], a[:, 0, 7] = -u[:, 0] * v[:, 0], -u[:, 1] * v[:, 0]
    a[:, 1, 3], a[:, 1, 4], a[:, 1, 5] = u[:, 0], u[:, 1], 1
    a[:, 1, 6], a[:, 1, 7] = -u[:, 0] * v[:, 1], -u[:, 1] * v[:, 1]
    b = np.asarray(v[:, :2], dtype=_F32).reshape(2 * n, 1)
    return a.reshape(2 * n, 8), b


def calc_corresp(u, v):
    """8x9 float32 DLT matrix of 4 correspondences (reference homography.py:4-14)."""
    return _null_rows(np.asarray(u)[:4], np.asarray(v)[:4])


def calc_correspLinear(u, v):
    """(A 8x8, b 8x1) float32 of the inhomogeneous system (reference homography.py:16-28)."""
    return _inhom_rows(np.asarray(u)[:4], np.asarray(v)[:4])


def calc_correspCollective(u, v):
    """2M x 9 float32 (reference homography.py:30-46)."""
    return _null_rows(u, v)


def calc_correspLinearCollective(u, v):
    """(A 2M x 8, b 2M x 1) float32 (reference homography.py:48-69)."""
    return _inhom_rows(u, v)


# ------------------------------------------------------------------------------------------------
# solvers
# ------------------------------------------------------------------------------------------------
def _pack_uv(u, v):
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    assert u.ndim == 2 and v.ndim == 2 and u.shape[0] == v.shape[0] and u.shape[1] >= 2 and v.shape[1] >= 2
    return np.ascontiguousarray(u[:, :2]), np.ascontiguousarray(v[:, :2])


def _rows_f32(u, v):
    """(M,4) float32 rows (x, y, x', y') -- the correspondence layout of the C ABI."""
    u, v = np.asarray(u), np.asarray(v)
    return np.ascontiguousarray(np.concatenate([u[:, :2], v[:, :2]], axis=1), dtype=_F32)


def calcHomography(u, v, collective=False):
    """DLT homography u -> v as a (3,3) float32 with h22 = 1 (reference homography.py:71-88).

    4 points: the kernel solves the same system the reference's SVD null vector satisfies
    (b200geo_dlt4).  collective=True (M points; unused by RANSAC.run): the unit-norm minimiser of
    |A h| over the 2M x 9 matrix of calc_correspCollective, i.e. the reference's last right singular
    vector, from the float64 Gram matrix by Jacobi rotations (b200geo_homography_svd_host)."""
    ctx = _cabi.context()
    if collective:
        H = np.empty(9, np.float64)
        rows = _rows_f32(u, v)
        _cabi.check(ctx.lib.b200geo_homography_svd_host(ctx.handle, _cabi.ptr(rows), rows.shape[0], _cabi.ptr(H)))
        return H.reshape(3, 3).astype(_F32)
    uu, vv = _pack_uv(u, v)
    assert uu.shape[0] == 4, "calcHomography needs 4 correspondences"
    H = np.empty(9, np.float64)
    _cabi.check(ctx.lib.b200geo_homography4_host(ctx.handle, _cabi.ptr(uu), _cabi.ptr(vv), _cabi.ptr(H)))
    return H.reshape(3, 3).astype(_F32)


def calcHomographyLinear(u, v, collective=False):
    """Homography u -> v as a (3,3) float64, h22 = 1 (reference homography.py:90-105).

    The reference evaluates inv(A^T A) A^T b in float32; the kernels solve the same system
    (4 points: exactly; M points: its least-squares minimiser) in float64."""
    ctx = _cabi.context()
    H = np.empty(9, np.float64)
    if collective:
        rows = _rows_f32(u, v)
        _cabi.check(ctx.lib.b200geo_refit_host(ctx.handle, _cabi.ptr(rows), rows.shape[0], None, _cabi.ptr(H)))
    else:
        uu, vv = _pack_uv(u, v)
        assert uu.shape[0] == 4, "calcHomographyLinear needs 4 correspondences (or collective=True)"
        _cabi.check(ctx.lib.b200geo_homography4_host(ctx.handle, _cabi.ptr(uu), _cabi.ptr(vv), _cabi.ptr(H)))
    return H.reshape(3, 3)


# ------------------------------------------------------------------------------------------------
# samplers on explicit coordinates (reference homography.py:108-140)
# ------------------------------------------------------------------------------------------------
def _zero_origin(img):
    """The reference's samplers zero img[0,0,:] of the caller's array (homography.py:112-116)."""
    chn = img.shape[2]
    img[0, 0, 0] = 0
    img[0, 0, 1] = 0
    img[0, 0, 2] = 0
    if chn == 4:
        img[0, 0, 3] = 0


def _src_type(img):
    if img.dtype == np.uint8:
        return _cabi.U8
    if img.dtype == np.float32:
        return _cabi.F32
    raise TypeError(f"image dtype {img.dtype} not supported (uint8 or float32)")


def _sample(z_t, img, h, w, mh, mw, sampler):
    assert img.ndim == 3 and img.shape[2] in (3, 4)
    _zero_origin(img)
    src = np.ascontiguousarray(img)
    st = _src_type(src)
    z = np.asarray(z_t, dtype=np.float64)
    zx, zy = np.ascontiguousarray(z[0]), np.ascontiguousarray(z[1])
    n = zx.shape[0]
    chn = src.shape[2]
    if sampler == _cabi.NN:
        out = np.empty((n, chn), src.dtype)
        dt = st
    else:
        out = np.empty((n, chn), np.float64)
        dt = _cabi.F64
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_sample_host(ctx.handle, _cabi.ptr(src), src.shape[0], src.shape[1], chn, st,
                                            _cabi.ptr(zx), _cabi.ptr(zy), n, int(h), int(w), sampler, dt,
                                            _cabi.ptr(out)))
    return out.reshape(mh, mw, chn)


def nearestNeighbor(z_t, img, h, w, mh, mw):
    """Reference homography.py:108-121: round-half-up by truncation, out-of-bounds -> pixel (0,0)."""
    return _sample(z_t, img, h, w, mh, mw, _cabi.NN)


def bilinear(z_t, img, h, w, mh, mw):
    """Reference homography.py:123-138: 4-tap float64 interpolation."""
    return _sample(z_t, img, h, w, mh, mw, _cabi.BILINEAR)


convertfunc = {'nn': nearestNeighbor, 'bilinear': bilinear}


# ------------------------------------------------------------------------------------------------
# backward warps
# ------------------------------------------------------------------------------------------------
def _linspace_params(start, stop, num):
    """(x0, step, last) such that np.linspace(start, stop, num)[i] == (last if i == num-1 else i*step + x0)."""
    start, stop = float(start), float(stop)
    if num <= 1:
        return start, 0.0, start
    step = (stop - start) / (num - 1)
    return start, step, stop


def _warp(img, H, gx, gy, bounds_hw, convert, out_kind):
    """Run K4 on one host image.  gx, gy = (start, stop, num) of the output grid; out_kind in
    {'native', 'u8'}: 'native' is what the reference's sampler returns (nn: img dtype; bilinear:
    float64), 'u8' fuses the astype(np.uint8) of transformImage[H]."""
    sampler = _cabi.SAMPLERS[convert]          # KeyError for unknown names, like convertfunc[convert]
    assert img.ndim == 3 and img.shape[2] in (3, 4)
    ih, iw, chn = img.shape
    _zero_origin(img)
    src = np.ascontiguousarray(img)
    st = _src_type(src)
    ow, oh = int(gx[2]), int(gy[2])
    if out_kind == 'u8':
        out = np.empty((oh, ow, chn), np.uint8)
        dt = _cabi.U8
    elif sampler == _cabi.NN:
        out = np.empty((oh, ow, chn), src.dtype)
        dt = st
    else:
        out = np.empty((oh, ow, chn), np.float64)
        dt = _cabi.F64
    if ow <= 0 or oh <= 0:
        return out
    invH = np.linalg.inv(np.asarray(H, dtype=np.float64))      # reference homography.py:172, 203
    job = _cabi.WarpJob()
    job.invH[:] = invH.reshape(9).tolist()
    job.x0, job.dx, job.x_last = _linspace_params(*gx)
    job.y0, job.dy, job.y_last = _linspace_params(*gy)
    job.ow, job.oh = ow, oh
    job.bh, job.bw = int(bounds_hw[0]), int(bounds_hw[1])
    job.ih, job.iw = ih, iw
    job.src_offset, job.src_pitch = 0, src.strides[0]
    job.dst_offset, job.dst_pitch = 0, out.strides[0]
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_warp_host(ctx.handle, _cabi.ptr(src), _cabi.ptr(out), C.byref(job), chn, st, dt, sampler))
    return out


def _bbox(img, H, boundary):
    """Output rectangle of wrapPerspective (reference homography.py:143-163)."""
    h, w, _ = img.shape
    bnd = np.array([[0, w - 1, w - 1, 0],
                    [0, 0, h - 1, h - 1],
                    [1., 1, 1, 1]])
    bnd_n = H @ bnd
    bnd_n /= bnd_n[-1, :]
    max_x = int(np.max(bnd_n[0, :]))
    min_x = int(np.min(bnd_n[0, :]))
    max_y = int(np.max(bnd_n[1, :]))
    min_y = int(np.min(bnd_n[1, :]))
    if boundary:
        min_x = max(min_x, 0)
        min_y = max(min_y, 0)
    return min_x, max_x, min_y, max_y


def wrapPerspective(img, H, convert='nn', boundary=0, crop=True, _out='native'):
    """Backward-warp img by H into the bounding box of its warped corners
    (reference homography.py:142-184).  Returns (img_n, min_x, min_y)."""
    H = np.asarray(H, dtype=np.float64)
    h, w, c = img.shape
    min_x, max_x, min_y, max_y = _bbox(img, H, boundary)
    max_w = max_x - min_x + 1
    max_h = max_y - min_y + 1
    img_t = _warp(img, H, (min_x, max_x, max_w), (min_y, max_y, max_h), (h, w), convert, _out)
    if crop:
        return img_t, min_x, min_y
    img_n = np.zeros((max_h, max_w, 3), dtype=np.float32)          # reference homography.py:181-183
    img_n[min_y:min_y + max_h, min_x:min_x + max_w, :] = img_t
    return img_n, min_x, min_y


def wrapPerspectiveScan(img, H, res, convert='nn', _out='native'):
    """Backward-warp img by H onto a fixed res = (h, w) grid, np.linspace(0, w, w) x
    np.linspace(0, h, h) (reference homography.py:186-209).  As in the reference, the sampler's
    bounds are ``res`` (not the image size)."""
    H = np.asarray(H, dtype=np.float64)
    h, w = res
    h, w = int(h), int(w)
    img_n = _warp(img, H, (0, w, w), (0, h, h), (h, w), convert, _out)
    return img_n, 0, 0


def transformImage(img, u, v, box=None, method='bilinear'):
    """Scanner entry point (reference homography.py:211-228; called from app.py:356).
    4-point homography u -> v, backward warp, uint8, crop to the destination rectangle."""
    H = calcHomographyLinear(u.T[:, :2], v.T[:, :2])
    if box is None:
        imgn, mx, my = wrapPerspective(img, H, convert=method, _out='u8')
    else:
        imgn, mx, my = wrapPerspectiveScan(img, H, box, convert=method, _out='u8')
    sx = int(v[0, 0] - mx); sy = int(v[1, 0] - my)
    ex = int(v[0, 2] - mx); ey = int(v[1, 2] - my)
    return imgn[sy:ey + 1, sx:ex + 1, :]


def transformImageH(img, H, method='bilinear'):
    """Stitch entry point (reference homography.py:230-242): warp by H; 3-channel images come
    back as uint8, 4-channel (RGBA float32 from addAlpha) as float64.  Returns (img, mx, my)."""
    if img.shape[2] == 3:
        return wrapPerspective(img, H, convert=method, boundary=0, crop=True, _out='u8')
    return wrapPerspective(img, H, convert=method, boundary=0, crop=True)


# ------------------------------------------------------------------------------------------------
# compose (reference homography.py:244-338).  For uint8 RGB images with
# paste or 'Rate' blending, stitchPanorama is ONE fused kernel (K5, csrc/compose.cuh): warp of imgT,
# canvas placement, paste / alpha blend and the uint8 conversion, with no intermediate RGBA image.
# Other inputs (float images, 'Gradient' alpha) take the reference's own sequence of steps around
# the GPU warp.
# ------------------------------------------------------------------------------------------------
class BLENDDIR(object):
    LEFT = 0
    RIGHT = 1
    TOP = 2
    DOWN = 3


def _ramp(h, w, reverse=False):
    x = np.linspace(w - 1, 0, w) if reverse else np.linspace(0, w - 1, w)
    y = np.linspace(h - 1, 0, h) if reverse else np.linspace(0, h - 1, h)
    xx, yy = np.meshgrid(x, y)
    return (xx + yy) / (w + h) * 0.5


def addAlpha(img, method="Rate", rate=0.2, direction=BLENDDIR.LEFT, alphaOnly=False):
    """Append (or return) an alpha plane (reference homography.py:250-286); float32."""
    h, w, c = img.shape
    rate += 1e-10
    if not alphaOnly:
        imgn = np.zeros((h, w, c + 1), dtype=np.float32)
        imgn[:, :, :c] = img
        if method == 'Rate':
            print(rate)
            imgn[:, :, c] = rate
        elif method == 'Gradient':
            imgn[:, :, c] = _ramp(h, w)       # the reference only defines the LEFT ramp here
            print('not implement yet')
        return imgn
    alpha = np.zeros((h, w, 1), dtype=np.float32)
    if method == 'Rate':
        print(rate)
        alpha[:, :, 0] = rate
    elif method == 'Gradient':
        alpha[:, :, 0] = _ramp(h, w, reverse=(direction == BLENDDIR.RIGHT))
        print('not implement yet')
    return alpha


def _placement(mx, my, wt, ht, wq, hq):
    """Canvas rectangles of the warped image T and the query image Q from the warp offset
    (reference homography.py:303-321)."""
    tsx, tsy, tex, tey = 0, 0, wt - 1, ht - 1
    qsx, qsy, qex, qey = 0, 0, wq - 1, hq - 1
    if mx < 0 and my < 0:
        qsx, qsy, qex, qey = -mx, -my, -mx + wq - 1, -my + hq - 1
    elif mx < 0:
        tsy, tey = my, my + ht - 1
        qsx, qex = -mx, -mx + wq - 1
    elif my < 0:
        tsx, tex = mx, mx + wt - 1
        qsy, qey = -my, -my + hq - 1
    else:
        tsx, tsy, tex, tey = mx, my, mx + wt - 1, my + ht - 1
    return (tsx, tsy, tex, tey), (qsx, qsy, qex, qey), max(tex + 1, qex + 1), max(tey + 1, qey + 1)


def _fusable(imgQ, imgT, blending):
    return (isinstance(imgQ, np.ndarray) and isinstance(imgT, np.ndarray) and imgQ.dtype == np.uint8 and imgT.dtype == np.uint8
            and imgQ.ndim == 3 and imgT.ndim == 3 and imgQ.shape[2] == 3 and imgT.shape[2] == 3
            and (not blending or (isinstance(blending, str) and blending == 'Rate')))


def _stitch_fused(imgQ, imgT, H, blending, blendrate):
    """stitchPanorama for uint8 RGB inputs as one kernel (b200geo_stitch_compose_host)."""
    H = np.asarray(H, dtype=np.float64)
    if blending:
        print(blendrate + 1e-10)                 # addAlpha prints the rate it stores (reference homography.py:257)
    else:
        _zero_origin(imgT)                       # transformImageH(imgT) zeroes the caller's pixel; with blending it is addAlpha's copy
    h, w, _ = imgT.shape
    hq, wq, _ = imgQ.shape
    min_x, max_x, min_y, max_y = _bbox(imgT, H, 0)
    wt, ht = max_x - min_x + 1, max_y - min_y + 1
    (tsx, tsy, tex, tey), (qsx, qsy, qex, qey), fw, fh = _placement(min_x, min_y, wt, ht, wq, hq)
    canvas = np.empty((fh, fw, 3), np.uint8)
    if wt <= 0 or ht <= 0 or fw <= 0 or fh <= 0 or min(tsx, tsy, qsx, qsy) < 0:
        raise ValueError("stitchPanorama: empty or negative canvas placement")   # the reference's slice assignment fails here
    src, q = np.ascontiguousarray(imgT), np.ascontiguousarray(imgQ)
    job = _cabi.WarpJob()
    job.invH[:] = np.linalg.inv(H).reshape(9).tolist()
    job.x0, job.dx, job.x_last = _linspace_params(min_x, max_x, wt)
    job.y0, job.dy, job.y_last = _linspace_params(min_y, max_y, ht)
    job.ow, job.oh, job.bw, job.bh, job.iw, job.ih = wt, ht, w, h, w, h
    job.src_offset, job.src_pitch, job.dst_offset, job.dst_pitch = 0, src.strides[0], 0, 0
    p = _cabi.Compose()
    p.fw, p.fh, p.tsx, p.tsy, p.qsx, p.qsy, p.qw, p.qh = fw, fh, tsx, tsy, qsx, qsy, wq, hq
    p.q_pitch, p.canvas_pitch = q.strides[0], canvas.strides[0]
    p.alpha_t = float(np.float32(blendrate + 1e-10))
    p.alpha_q = float(np.float32(1 + 1e-10 - blendrate))
    p.alpha_bg = float(np.float32(1e-10))
    p.blend = 1 if blending else 0
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_stitch_compose_host(ctx.handle, _cabi.ptr(src), _cabi.ptr(q), _cabi.ptr(canvas), C.byref(job), C.byref(p)))
    return canvas


def stitchPanorama(imgQ, imgT, H, method='bilinear', blending=False, blendrate=0.2):
    """Warp imgT by H (GPU) and compose it with imgQ on a common canvas
    (reference homography.py:288-338).  Returns uint8."""
    if _fusable(imgQ, imgT, blending):
        return _stitch_fused(imgQ, imgT, H, blending, blendrate)
    if blending:
        imgT = addAlpha(imgT, method=blending, rate=blendrate)
    img_t, mx, my = transformImageH(imgT, H)
    ht, wt, ct = img_t.shape
    hq, wq, cq = imgQ.shape
    (tsx, tsy, tex, tey), (qsx, qsy, qex, qey), fw, fh = _placement(mx, my, wt, ht, wq, hq)
    if blending:
        imgn = np.zeros((fh, fw, ct), dtype=np.float32)
        imgn[qsy:qey + 1, qsx:qex + 1, :3] = imgQ[:, :, :3].astype(np.float32)
        imgn[:, :, 3] += 1e-10
        imgn[qsy:qey + 1, qsx:qex + 1, 3] = (1 + 1e-10 - blendrate) if blending == 'Rate' else 1
        t = (slice(tsy, tey + 1), slice(tsx, tex + 1))
        base = imgn[t][:, :, 3:4] + img_t[:, :, 3:4]
        imgn[t[0], t[1], :3] = (imgn[t][:, :, 3:4] / base) * imgn[t][:, :, :3] + (img_t[:, :, 3:4] / base) * img_t[:, :, :3]
        return imgn[:, :, :3].astype(np.uint8)
    imgn = np.zeros((fh, fw, ct), dtype=np.uint8)
    imgn[tsy:tey + 1, tsx:tex + 1, :] = img_t
    imgn[qsy:qey + 1, qsx:qex + 1, :] = imgQ
    return imgn


# ------------------------------------------------------------------------------------------------
# cylindrical projections (reference homography.py:340-416).  Dead code for the app
# (ransac.py:249-251 is commented out), kept because `from homography import ... cylindericlMap` (ransac.py:3)
# and homography.example1 name them.
# ------------------------------------------------------------------------------------------------
def cylindericalTransform(img, f, xc=0.5, yc=0.5, method='bilinear', wrap="fwd"):
    """Reference homography.py:340-389, statement for statement around the GPU sampler.

    What the reference computes here is narrower than its docstring: `z` is (3, N) and the reference indexes its
    COLUMNS (`z[:,0]`, `_z[:,1]`), so only output pixels 0 and 1 receive cylindrical coordinates and every other
    pixel is sampled at its own position; the bounding box (min_x, min_y) likewise comes from the first two
    columns of `bnd`.  That is reproduced as is (the same numpy expressions build `_z`; the sampler runs on the
    GPU).  Two branches of the reference do not produce a defined result and raise here instead:
      * method='bilinear' (the default): the grid holds x = w-1 exactly, which the sampler's mask lets through,
        so the reference dies with IndexError at homography.py:135 for every image -- same exception here;
      * wrap='backward': `_z = np.empty(z.shape)` is sampled with all but two columns uninitialised
        (homography.py:373-378) -- the output is whatever the allocator left: ValueError here."""
    h, w, c = img.shape
    _xc = int(xc * w) - 1
    _yc = int(yc * h) - 1
    __xc = int((1 - xc) * w) - 1
    __yc = int((1 - yc) * h) - 1
    bnd = np.array([[0, 0, _xc, w - 1, w - 1, __xc, w - 1, 0],
                    [0, _yc, 0, __yc, 0, h - 1, h - 1, h - 1],
                    [1., 1, 1, 1, 1, 1, 1, 1]])
    bnd_n = bnd.copy()
    tmp = (bnd[:, 0] - _xc) / f
    bnd_n[:, 0] = f * np.tan(tmp) + _xc
    bnd_n[:, 1] = (bnd[:, 1] - _yc) / np.cos(tmp) + _yc
    max_x = int(np.max(bnd_n[0, :]))
    min_x = int(np.min(bnd_n[0, :]))
    max_y = int(np.max(bnd_n[1, :]))
    min_y = int(np.min(bnd_n[1, :]))
    if wrap == "backward":
        raise ValueError("cylindericalTransform(wrap='backward') samples an uninitialised coordinate array in the reference "
                         "(homography.py:373-378): there is no defined output to reproduce")
    sampler = convertfunc[method]
    if method == 'bilinear':
        raise IndexError(f"index {w} is out of bounds for axis 1 with size {w}")      # homography.py:135, always (x = w-1 is sampled)
    x = np.linspace(0, w - 1, w)
    y = np.linspace(0, h - 1, h)
    xv, yv = np.meshgrid(x, y)
    z = np.dstack([xv, yv, np.ones((h, w))]).reshape([w * h, 3]).T
    _z = z.copy()
    tmp = (z[:, 0] - _xc) / f
    _z[:, 0] = f * np.tan(tmp) + _xc
    _z[:, 1] = (z[:, 1] - _yc) / np.cos(tmp) + _yc
    img_n = sampler(_z, img, h, w, h, w).astype(np.uint8)
    return img_n, min_x, min_y


def cylindricalWarp(img, K):
    """Reference homography.py:391-410: cylindrical warp of a BGR uint8 image for intrinsics K; returns BGRA with
    transparent (alpha 0) pixels where the cylinder leaves the image.

    The coordinate maps are the reference's own numpy expressions (float64 sin / cos, then float32); the resampling
    -- cv2.cvtColor(BGR2BGRA) + cv2.remap(..., INTER_AREA, BORDER_TRANSPARENT) in the reference -- runs on the GPU
    (b200geo_remap_linear_host: OpenCV's fixed-point bilinear remap, bit-exact).  cv2.remap leaves the pixels it
    does not write (taps outside the image) as its freshly allocated output had them, which is zero-filled memory in
    practice; here they are zero by construction."""
    h_, w_ = img.shape[:2]
    y_i, x_i = np.indices((h_, w_))
    X = np.stack([x_i, y_i, np.ones_like(x_i)], axis=-1).reshape(h_ * w_, 3)
    Kinv = np.linalg.inv(K)
    X = Kinv.dot(X.T).T
    A = np.stack([np.sin(X[:, 0]), X[:, 1], np.cos(X[:, 0])], axis=-1).reshape(w_ * h_, 3)
    B = K.dot(A.T).T
    B = B[:, :-1] / B[:, [-1]]
    B[(B[:, 0] < 0) | (B[:, 0] >= w_) | (B[:, 1] < 0) | (B[:, 1] >= h_)] = -1
    B = B.reshape(h_, w_, -1)
    assert img.ndim == 3 and img.shape[2] == 3 and img.dtype == np.uint8       # cv2.cvtColor(BGR2BGRA) accepts nothing else here
    img_rgba = np.empty((h_, w_, 4), np.uint8)
    img_rgba[:, :, :3] = img
    img_rgba[:, :, 3] = 255
    mx = np.ascontiguousarray(B[:, :, 0].astype(np.float32))
    my = np.ascontiguousarray(B[:, :, 1].astype(np.float32))
    out = np.zeros((h_, w_, 4), np.uint8)
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_remap_linear_host(ctx.handle, _cabi.ptr(img_rgba), h_, w_, 4, _cabi.ptr(mx), _cabi.ptr(my), h_, w_,
                                                  _cabi.ptr(out)))
    return out


def cylindericlMap(img, f=1600):
    """Reference homography.py:412-416."""
    h, w, c = img.shape
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]])
    img_cyl = cylindricalWarp(img, K)
    return img_cyl
