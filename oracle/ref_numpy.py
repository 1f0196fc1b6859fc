"""numpy restatement ("port") of the reference's hot path -- TEST INFRASTRUCTURE ONLY.

Same numpy / LAPACK / BLAS primitives as the reference, so it has the reference's numerical
behaviour and its performance character; used (a) to cross-check ``oracle.c`` against the live
reference when the golden vectors are generated, and (b) as the timed CPU arm of ``bench.py``
(``cpu_baseline.kind = "port"``; the Python reference itself cannot travel to the GPU box).

Not imported by the product.  Citations are file:line in /root/reference.
"""
from __future__ import annotations

import numpy as np

_F = np.float32


# ------------------------------------------------------------------------------------------------
# 4-point / M-point solves
# ------------------------------------------------------------------------------------------------
def dlt_rows_null(u, v):
    """2n x 9 float32 matrix whose null vector is the homography (homography.py:4-14, 30-46).
    Products are formed in the input dtype (float32 in RANSAC) before the final cast."""
    u = np.asarray(u)
    v = np.asarray(v)
    n = u.shape[0]
    m = np.zeros((n, 2, 9), dtype=_F)
    m[:, 0, 0], m[:, 0, 1], m[:, 0, 2] = -u[:, 0], -u[:, 1], -1
    m[:, 0, 6], m[:, 0, 7], m[:, 0, 8] = u[:, 0] * v[:, 0], u[:, 1] * v[:, 0], v[:, 0]
    m[:, 1, 3], m[:, 1, 4], m[:, 1, 5] = -u[:, 0], -u[:, 1], -1
    m[:, 1, 6], m[:, 1, 7], m[:, 1, 8] = u[:, 0] * v[:, 1], u[:, 1] * v[:, 1], v[:, 1]
    return m.reshape(2 * n, 9)


def homography_null(u, v):
    """calcHomography (homography.py:71-88): last right-singular vector, divided by h[8]; float32."""
    _, _, vt = np.linalg.svd(dlt_rows_null(u, v))
    h = vt[-1].reshape(3, 3)
    return h / h.item(8)


def dlt_rows_inhom(u, v):
    """(A, b) of the inhomogeneous system, float32 (homography.py:16-28, 48-69)."""
    u = np.asarray(u)
    v = np.asarray(v)
    n = u.shape[0]
    a = np.zeros((n, 2, 8), dtype=_F)
    a[:, 0, 0], a[:, 0, 1], a[:, 0, 2] = u[:, 0], u[:, 1], 1
    a[:, 0, 6], a[:, 0, 7] = -u[:, 0] * v[:, 0], -u[:, 1] * v[:, 0]
    a[:, 1, 3], a[:, 1, 4], a[:, 1, 5] = u[:, 0], u[:, 1], 1
    a[:, 1, 6], a[:, 1, 7] = -u[:, 0] * v[:, 1], -u[:, 1] * v[:, 1]
    b = np.asarray(v[:, :2], dtype=_F).reshape(2 * n, 1)
    return a.reshape(2 * n, 8), b


def homography_normal_eq(u, v):
    """calcHomographyLinear (homography.py:90-105): inv(A^T A) @ (A^T b) in float32 -> float64 3x3."""
    A, b = dlt_rows_inhom(u, v)
    h = np.linalg.inv(A.T @ A) @ (A.T @ b)
    return np.append(h.ravel().astype(np.float64), 1.0).reshape(3, 3)


# ------------------------------------------------------------------------------------------------
# scoring + RANSAC loop
# ------------------------------------------------------------------------------------------------
def _project(Hm, P):
    """HomoModel.fwd / reproj (ransac.py:55-76): float32 homogeneous product and divide."""
    x = np.ones((3, P.shape[1]), dtype=_F)
    x[:2, :] = P[:2]
    y = Hm @ x
    return y / (y[-1, :] + 1e-10)


def _l2(a, b):
    d = a - b
    return np.sqrt(np.sum(d * d, axis=0))   # ransac.py:78-82


def loss(Hm, X, Y, method="reproj"):
    """HomoModel.computeLoss (ransac.py:84-98)."""
    if method == "fwd":
        return _l2(_project(Hm, X)[:2], Y)
    if method == "backward":
        return _l2(_project(np.linalg.inv(Hm), Y)[:2], X)
    if method == "reproj":
        e = _l2(_project(Hm, X)[:2], Y)
        e += _l2(_project(np.linalg.inv(Hm), Y)[:2], X)
        return e
    raise SystemExit("Invalid method!")      # ransac.py:97


def ransac_run(X, Y, th=5, d=50, n=4, k=1000, method="reproj", idx=None, refit=True):
    """RANSAC.run (ransac.py:159-213).  ``idx`` (k,4) replaces the per-iteration
    np.random.randint(0, M, n) draw (ransac.py:177) with explicit sample sets.
    Returns (H, (inlier indices,), count, per-iteration counts evaluated, winner index)."""
    M = X.shape[1]
    assert M == Y.shape[1], "data observation not consistent!"
    need = M * d / 100 + n
    best, best_mask, win = 0, None, -1
    counts = []
    c = 0
    for i in range(k):
        s = np.random.randint(0, M, n) if idx is None else idx[i]
        Hm = homography_null(X[:, s].T[:, :2], Y[:, s].T[:, :2])
        mask = loss(Hm, X, Y, method) < th
        c = int(np.sum(mask))
        counts.append(c)
        if c >= need:
            best, best_mask, win = c, mask, i
            break
        if c > best:
            best, best_mask, win = c, mask, i
    inl = np.where(best_mask)
    Hf = homography_normal_eq(X[:, inl[0]].T[:, :2], Y[:, inl[0]].T[:, :2]) if refit else None
    return Hf, inl, best, np.asarray(counts, np.int32), win


def score_block(X, Y, idx, th, method="fwd"):
    """Counts for a block of explicit sample sets (no early exit) -- the unit the CPU baseline
    times for the synthetic single-pair config."""
    out = np.empty(len(idx), np.int32)
    for i, s in enumerate(idx):
        Hm = homography_null(X[:, s].T[:, :2], Y[:, s].T[:, :2])
        out[i] = np.sum(loss(Hm, X, Y, method) < th)
    return out


# ------------------------------------------------------------------------------------------------
# backward warp (homography.py:108-209)
# ------------------------------------------------------------------------------------------------
def _sample(zt, img, bh, bw, oh, ow, bilinear):
    """Samplers nearestNeighbor / bilinear (homography.py:108-138).  Zeroes img[0,0,:] of the
    array it is given, as the reference does."""
    img[0, 0, :] = 0
    if not bilinear:
        q = (zt + 0.5).astype(np.int32).T
        bad = (q[:, 0] > bw - 1) | (q[:, 0] < 0) | (q[:, 1] > bh - 1) | (q[:, 1] < 0)
        q[bad, 0:2] = 0
        return img[q[:, 1], q[:, 0], :].reshape(oh, ow, -1)
    q = zt.T
    bad = (q[:, 0] > bw - 1) | (q[:, 0] < 0) | (q[:, 1] > bh - 1) | (q[:, 1] < 0)
    q[bad, 0:2] = 0
    qi = q.astype(np.int32)
    f = q - qi
    fx, fy = f[:, 0:1], f[:, 1:2]
    top = img[qi[:, 1], qi[:, 0], :] * (1 - fx) + img[qi[:, 1], qi[:, 0] + 1, :] * fx
    bot = img[qi[:, 1] + 1, qi[:, 0], :] * (1 - fx) + img[qi[:, 1] + 1, qi[:, 0] + 1, :] * fx
    return (top * (1 - fy) + bot * fy).reshape(oh, ow, -1)


def _backmap(H, gx, gy):
    xv, yv = np.meshgrid(gx, gy)
    z = np.dstack([xv, yv, np.ones(xv.shape)]).reshape(-1, 3).T
    zt = np.linalg.inv(H) @ z
    zt /= zt[-1, :]
    return zt


def warp_bbox(img, H, bilinear=True, boundary=0):
    """wrapPerspective, crop=True (homography.py:142-184)."""
    h, w, _ = img.shape
    q = H @ np.array([[0, w - 1, w - 1, 0], [0, 0, h - 1, h - 1], [1.0, 1, 1, 1]])
    q /= q[-1, :]
    x1, x0 = int(np.max(q[0])), int(np.min(q[0]))
    y1, y0 = int(np.max(q[1])), int(np.min(q[1]))
    if boundary:
        x0, y0 = max(x0, 0), max(y0, 0)
    ow, oh = x1 - x0 + 1, y1 - y0 + 1
    zt = _backmap(H, np.linspace(x0, x1, ow), np.linspace(y0, y1, oh))
    return _sample(zt, img, h, w, oh, ow, bilinear), x0, y0


def warp_scan(img, H, res, bilinear=True):
    """wrapPerspectiveScan (homography.py:186-209); the sampler's bounds are ``res``."""
    h, w = res
    zt = _backmap(H, np.linspace(0, w, w), np.linspace(0, h, h))
    return _sample(zt, img, h, w, h, w, bilinear), 0, 0


# ------------------------------------------------------------------------------------------------
# stitch compose (homography.py:250-338) -- the step after the warp
# ------------------------------------------------------------------------------------------------
def stitch_panorama(imgQ, imgT, H, blending=False, blendrate=0.2):
    """stitchPanorama (homography.py:288-338) with addAlpha(method='Rate') (:250-259) and
    transformImageH (:230-242) restated step for step: same dtypes, same order of the array
    expressions.  `blending` is False or 'Rate'.  Returns the uint8 canvas.  imgT[0,0,:] of the
    caller's array is zeroed in the paste path, as the reference does."""
    if blending:
        rgba = np.zeros(imgT.shape[:2] + (4,), dtype=np.float32)               # addAlpha :254-259
        rgba[:, :, :3] = imgT
        rgba[:, :, 3] = blendrate + 1e-10
        img_t, mx, my = warp_bbox(rgba, H, True)                                # float64, 4 channels (:241-242)
    else:
        img_t, mx, my = warp_bbox(imgT, H, True)
        img_t = img_t.astype(np.uint8)                                          # :240
    ht, wt, ct = img_t.shape
    hq, wq, _ = imgQ.shape
    tsx, tsy, tex, tey = 0, 0, wt - 1, ht - 1
    qsx, qsy, qex, qey = 0, 0, wq - 1, hq - 1
    if mx < 0 and my < 0:                                                       # :309-321
        qsx, qsy, qex, qey = -mx, -my, -mx + wq - 1, -my + hq - 1
    elif mx < 0:
        tsy, tey = my, my + ht - 1
        qsx, qex = -mx, -mx + wq - 1
    elif my < 0:
        tsx, tex = mx, mx + wt - 1
        qsy, qey = -my, -my + hq - 1
    else:
        tsx, tsy, tex, tey = mx, my, mx + wt - 1, my + ht - 1
    fw, fh = max(tex + 1, qex + 1), max(tey + 1, qey + 1)
    if blending:
        imgn = np.zeros((fh, fw, ct), dtype=np.float32)                         # :323-334
        imgn[qsy:qey + 1, qsx:qex + 1, :3] = imgQ[:, :, :3].astype(np.float32)
        imgn[:, :, 3] += 1e-10
        imgn[qsy:qey + 1, qsx:qex + 1, 3] = 1 + 1e-10 - blendrate
        base = imgn[tsy:tey + 1, tsx:tex + 1, 3:4] + img_t[:, :, 3:4]
        imgn[tsy:tey + 1, tsx:tex + 1, :3] = ((imgn[tsy:tey + 1, tsx:tex + 1, 3:4] / base) * imgn[tsy:tey + 1, tsx:tex + 1, :3]
                                              + (img_t[:, :, 3:4] / base) * img_t[:, :, :3])
        return imgn[:, :, :3].astype(np.uint8)
    imgn = np.zeros((fh, fw, ct), dtype=np.uint8)                               # :335-338
    imgn[tsy:tey + 1, tsx:tex + 1, :] = img_t
    imgn[qsy:qey + 1, qsx:qex + 1, :] = imgQ
    return imgn


# ------------------------------------------------------------------------------------------------
# brute-force Hamming matching (ransac.py:258-259: cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match)
# ------------------------------------------------------------------------------------------------
def bf_match_hamming(A, B, cross_check=True):
    """OpenCV 4.x BFMatcher.match semantics for NORM_HAMMING (third-party: modules/core/src/batch_distance.cpp,
    K = 1): nearest train descriptor per query descriptor, the first index among equal distances; with
    cross-check the pass is repeated with the roles swapped and (i, j) is kept iff best_of_train[j] == i.
    Returns (queryIdx, trainIdx, distance) arrays in query order.  Pinned against cv2 itself by
    tests/test_oracle_golden.py (cv2 is a library of the image, not part of the reference)."""
    A, B = np.asarray(A, np.uint8), np.asarray(B, np.uint8)
    if A.shape[0] == 0 or B.shape[0] == 0:
        z = np.zeros(0, np.int64)
        return z, z, z
    bits = np.unpackbits(A, axis=1).astype(np.int32)[:, None, :] != np.unpackbits(B, axis=1).astype(np.int32)[None, :, :]
    D = bits.sum(axis=2)
    ja = D.argmin(axis=1)                     # numpy argmin = first index of the minimum
    q = np.arange(A.shape[0])
    keep = np.ones(A.shape[0], bool)
    if cross_check:
        ib = D.argmin(axis=0)
        keep = ib[ja] == q
    return q[keep], ja[keep], D[q, ja][keep]


def remap_linear_transparent(src, map_x, map_y, dst):
    """cv2.remap(src, map_x, map_y, cv2.INTER_LINEAR, dst=dst, borderMode=cv2.BORDER_TRANSPARENT) for uint8 images with
    1, 2 or 4 channels (OpenCV 4.13 imgwarp.cpp: float32 maps -> 1/32 px fixed point with round-half-even, 15-bit weights,
    (sum + 2^14) >> 15; pixels whose integer source position lies outside the image keep dst, taps
    beyond the last column / row repeat it).  The resampling step of the
    reference's cylindricalWarp (homography.py:410).  Pinned to cv2.remap in tests/test_oracle_golden.py."""
    src = np.asarray(src)
    ih, iw = src.shape[:2]
    s = src.reshape(ih, iw, -1).astype(np.int64)
    out = np.array(dst, copy=True)
    o = out.reshape(out.shape[0], out.shape[1], -1)
    fx = np.rint(np.asarray(map_x, np.float32) * np.float32(32)).astype(np.int64)       # rint: half to even, like cvRound
    fy = np.rint(np.asarray(map_y, np.float32) * np.float32(32)).astype(np.int64)
    ix, iy, ax, ay = fx >> 5, fy >> 5, fx & 31, fy & 31
    ok = (ix >= 0) & (ix < iw) & (iy >= 0) & (iy < ih)          # (observed on 4.13: written iff the integer position is inside; ...
    r, c = np.nonzero(ok)
    x, y, a, b = ix[r, c], iy[r, c], ax[r, c][:, None], ay[r, c][:, None]
    x1, y1 = np.minimum(x + 1, iw - 1), np.minimum(y + 1, ih - 1)  # ... the taps beyond the last column / row repeat it)
    t = (32 - a) * (32 - b) * 32 * s[y, x] + a * (32 - b) * 32 * s[y, x1] + (32 - a) * b * 32 * s[y1, x] + a * b * 32 * s[y1, x1]
    o[r, c] = np.clip((t + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return out


def cylindrical_maps(h, w, K):
    """The float32 coordinate maps of the reference's cylindricalWarp (homography.py:393-406), same numpy expressions."""
    y_i, x_i = np.indices((h, w))
    X = np.stack([x_i, y_i, np.ones_like(x_i)], axis=-1).reshape(h * w, 3)
    X = np.linalg.inv(K).dot(X.T).T
    A = np.stack([np.sin(X[:, 0]), X[:, 1], np.cos(X[:, 0])], axis=-1).reshape(w * h, 3)
    B = K.dot(A.T).T
    B = B[:, :-1] / B[:, [-1]]
    B[(B[:, 0] < 0) | (B[:, 0] >= w) | (B[:, 1] < 0) | (B[:, 1] >= h)] = -1
    B = B.reshape(h, w, -1)
    return B[:, :, 0].astype(np.float32), B[:, :, 1].astype(np.float32)


def cylindrical_warp(img, K):
    """cylindricalWarp (homography.py:391-410) with zero-initialised transparent pixels."""
    h, w = img.shape[:2]
    mx, my = cylindrical_maps(h, w, K)
    rgba = np.dstack([img, np.full((h, w), 255, np.uint8)])
    return remap_linear_transparent(rgba, mx, my, np.zeros((h, w, 4), np.uint8))
