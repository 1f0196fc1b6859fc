"""CPU restatement of cv2.ORB_create().detectAndCompute(gray, None) -- TEST INFRASTRUCTURE ONLY (never imported by the product).

The reference calls it at ransac.py:255-257.  The algorithm lives in OpenCV (third-party, unpinned
by the reference; this image has 4.13.0: modules/features2d/src/orb.cpp, fast.cpp, fast_score.cpp, keypoint.cpp;
modules/imgproc/src/resize.cpp, filter.dispatch.cpp, mathfuncs_core.simd.hpp).  Default parameters only, as the reference
uses them: nfeatures 500, scaleFactor 1.2f, nlevels 8, edgeThreshold 31, firstLevel 0, WTA_K 2, HARRIS_SCORE, patchSize 31,
fastThreshold 20, no mask.  Pinned against OpenCV itself, stage by stage and end to end (tests/test_oracle_orb.py and the
recorded outputs in tests/golden/orb.npz).

    pyramid      level L = resize(level L-1, round(size / scale_L), INTER_LINEAR_EXACT): 8.8 fixed-point coefficients,
                 (sum + 2^15) >> 16; every level extended by 32 px of BORDER_REFLECT_101
    FAST-9/16    score = max over the 16 arcs of 9 ring pixels of min |difference| (either sign) - 1, corner iff > 20,
                 3 x 3 non-maximum suppression (strict), row-major order
    selection    drop points within 31 px of the level's border; retainBest(2 n_L) by FAST score, Harris response on a
                 7 x 7 block (integer gradients, float32 score), retainBest(n_L) by Harris -- std::nth_element + std::partition
                 of libstdc++, whose element order the output inherits
    orientation  intensity centroid in the radius-15 disc, cv::fastAtan2
    descriptor   7 x 7 Gaussian (sigma 2) through sepFilter2D's float32 path, 256 pair tests on the rotated pattern
"""
from __future__ import annotations

import math

import numpy as np

NFEATURES, NLEVELS, EDGE, PATCH, FAST_T, HALF_PATCH, BORDER = 500, 8, 31, 31, 20, 15, 32
SCALE_FACTOR = float(np.float32(1.2))        # ORB::create takes a float, ORB_Impl stores it as double
HARRIS_K = np.float32(0.04)


# ------------------------------------------------------------------------------------------------ pyramid
def rgb2gray(rgb):
    """cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY) for uint8 (ransac.py:252-253): 15-bit fixed point, round half up."""
    a = rgb.astype(np.int64)
    return ((a[..., 0] * 9798 + a[..., 1] * 19235 + a[..., 2] * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def level_scales():
    return [np.float32(SCALE_FACTOR ** lv) for lv in range(NLEVELS)]          # getScale: (float)pow(scaleFactor, level)


def cv_round(v):
    """cvRound: round half to even (cvtsd2si / lrint)."""
    return int(np.rint(v))


def level_sizes(h, w):
    out = []
    for s in level_scales():
        inv = np.float32(1.0) / s
        out.append((cv_round(np.float32(h) * inv), cv_round(np.float32(w) * inv)))
    return out


def _linear_coeffs(src, dst):
    """interpolationLinear<uint8_t>::getCoeffs for every destination index: offsets and ufixedpoint16 (8.8) coefficient pairs."""
    scale = 1.0 / (float(dst) / float(src))                 # softdouble::one() / softdouble(inv_scale); inv_scale = dsize / ssize
    ofs = np.zeros(dst, np.int64)
    c0 = np.zeros(dst, np.int64)
    c1 = np.zeros(dst, np.int64)
    for d in range(dst):
        f = scale * (d + 0.5) - 0.5
        i = math.floor(f)
        if i >= 0 and src > 1:
            if i < src - 1:
                ofs[d] = i
                c1[d] = cv_round((f - i) * 256.0)
                c0[d] = 256 - c1[d]
            else:
                ofs[d], c0[d], c1[d] = src - 1, 256, 0
        else:
            ofs[d], c0[d], c1[d] = 0, 256, 0
    return ofs, c0, c1


def resize_linear_exact(img, dh, dw):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for one-channel uint8."""
    sh, sw = img.shape
    xo, x0, x1 = _linear_coeffs(sw, dw)
    yo, y0, y1 = _linear_coeffs(sh, dh)
    s = img.astype(np.int64)
    xn = np.minimum(xo + 1, sw - 1)
    hl = s[:, xo] * x0[None, :] + s[:, xn] * x1[None, :]                       # ufixedpoint16 rows (8.8)
    yn = np.minimum(yo + 1, sh - 1)
    v = hl[yo, :] * y0[:, None] + hl[yn, :] * y1[:, None]                      # ufixedpoint32 (16.16)
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def pyramid(gray):
    """[(level image, level image with its 32-pixel BORDER_REFLECT_101 frame)] for the 8 levels."""
    h, w = gray.shape
    out = []
    prev = gray
    for lv, (lh, lw) in enumerate(level_sizes(h, w)):
        cur = gray if lv == 0 else resize_linear_exact(prev, lh, lw)
        out.append((cur, np.pad(cur, BORDER, mode="reflect")))
        prev = cur
    return out


# ------------------------------------------------------------------------------------------------ FAST
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast_scores(img, threshold=FAST_T):
    """Score map of FastFeatureDetector(threshold, nonmax=True): cornerScore<16> where the pixel is a FAST-9 corner, else 0.
    Rows 3 .. h-4 and columns 3 .. w-4 are examined (fast.cpp)."""
    h, w = img.shape
    sc = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return sc
    v = img[3:h - 3, 3:w - 3].astype(np.int16)
    d = np.stack([img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int16) - v for dx, dy in RING])       # ring - centre
    best = np.full(v.shape, -1000, np.int16)
    for k in range(16):
        idx = [(k + j) % 16 for j in range(9)]
        arc = d[idx]
        best = np.maximum(best, np.maximum(arc.min(0), (-arc).min(0)))
    s = np.where(best > threshold, best - 1, 0)
    sc[3:h - 3, 3:w - 3] = s
    return sc


def fast_detect(img, threshold=FAST_T):
    """Keypoints (x, y, response) in OpenCV's order (row by row)."""
    sc = fast_scores(img, threshold)
    h, w = sc.shape
    c = sc[1:-1, 1:-1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= c > sc[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    ys, xs = np.nonzero(keep)
    return np.stack([xs + 1, ys + 1, c[ys, xs]], 1).astype(np.int64)


# ------------------------------------------------------------------------------------------------ selection
def _nth_element(a, first, nth, last, comp):
    """libstdc++ std::nth_element (introselect) on the Python list a[first:last]."""
    def move_median_to_first(res, x, y, z):
        if comp(a[x], a[y]):
            if comp(a[y], a[z]): a[res], a[y] = a[y], a[res]
            elif comp(a[x], a[z]): a[res], a[z] = a[z], a[res]
            else: a[res], a[x] = a[x], a[res]
        elif comp(a[x], a[z]): a[res], a[x] = a[x], a[res]
        elif comp(a[y], a[z]): a[res], a[z] = a[z], a[res]
        else: a[res], a[y] = a[y], a[res]

    def unguarded_partition(f, l, pivot):
        while True:
            while comp(a[f], a[pivot]): f += 1
            l -= 1
            while comp(a[pivot], a[l]): l -= 1
            if not f < l: return f
            a[f], a[l] = a[l], a[f]
            f += 1

    def heap_select(f, m, l):
        # std::__heap_select: make_heap on [f, m), then for i in [m, l): if comp(a[i], a[f]) pop_heap(f, m, i)
        def adjust(f, hole, n, value):
            top = hole
            child = hole
            while child < (n - 1) // 2:
                child = 2 * (child + 1)
                if comp(a[f + child], a[f + child - 1]): child -= 1
                a[f + hole] = a[f + child]
                hole = child
            if (n & 1) == 0 and child == (n - 2) // 2:
                child = 2 * (child + 1)
                a[f + hole] = a[f + child - 1]
                hole = child - 1
            parent = (hole - 1) // 2
            while hole > top and comp(a[f + parent], value):
                a[f + hole] = a[f + parent]
                hole = parent
                parent = (hole - 1) // 2
            a[f + hole] = value
        n = m - f
        if n >= 2:
            parent = (n - 2) // 2
            while True:
                adjust(f, parent, n, a[f + parent])
                if parent == 0: break
                parent -= 1
        for i in range(m, l):
            if comp(a[i], a[f]):
                value = a[i]
                a[i] = a[f]
                adjust(f, 0, n, value)

    if first == last or nth == last:
        return
    depth = 2 * ((last - first).bit_length() - 1)
    while last - first > 3:
        if depth == 0:
            heap_select(first, nth + 1, last)
            a[first], a[nth] = a[nth], a[first]
            return
        depth -= 1
        mid = first + (last - first) // 2
        move_median_to_first(first, first + 1, mid, last - 1)
        cut = unguarded_partition(first + 1, last, first)
        if cut <= nth: first = cut
        else: last = cut
    for i in range(first + 1, last):                       # std::__insertion_sort
        val = a[i]
        if comp(val, a[first]):
            a[first + 1:i + 1] = a[first:i]
            a[first] = val
        else:
            j = i
            while comp(val, a[j - 1]):
                a[j] = a[j - 1]
                j -= 1
            a[j] = val


def _partition(a, first, last, pred):
    """libstdc++ std::partition for bidirectional iterators."""
    while True:
        while True:
            if first == last: return first
            if pred(a[first]): first += 1
            else: break
        last -= 1
        while True:
            if first == last: return first
            if not pred(a[last]): last -= 1
            else: break
        a[first], a[last] = a[last], a[first]
        first += 1


def retain_best(kps, n, key):
    """KeyPointsFilter::retainBest (keypoint.cpp): the n best by response plus everything tied with the n-th, in the order
    std::nth_element and std::partition leave them.  kps: list of records; key(record) -> response."""
    kps = list(kps)
    if n >= 0 and len(kps) > n:
        if n == 0:
            return []
        _nth_element(kps, 0, n - 1, len(kps), lambda p, q: key(p) > key(q))
        amb = key(kps[n - 1])
        end = _partition(kps, n, len(kps), lambda p: key(p) >= amb)
        kps = kps[:end]
    return kps


def features_per_level(nfeatures=NFEATURES):
    factor = np.float32(1.0 / SCALE_FACTOR)
    nd = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(float(factor) ** NLEVELS))
    out, total = [], 0
    for _ in range(NLEVELS - 1):
        out.append(cv_round(nd))
        total += out[-1]
        nd = np.float32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def harris_response(ext, x, y):
    """HarrisResponses (orb.cpp) for integer positions (x, y) of a level, `ext` = the level with its 32 px frame."""
    e = ext.astype(np.int64)
    x = np.asarray(x) + BORDER
    y = np.asarray(y) + BORDER
    a = np.zeros(len(x), np.int64)
    b = np.zeros(len(x), np.int64)
    c = np.zeros(len(x), np.int64)
    for i in range(7):
        for j in range(7):
            yy, xx = y - 3 + i, x - 3 + j
            ix = (e[yy, xx + 1] - e[yy, xx - 1]) * 2 + (e[yy - 1, xx + 1] - e[yy - 1, xx - 1]) + (e[yy + 1, xx + 1] - e[yy + 1, xx - 1])
            iy = (e[yy + 1, xx] - e[yy - 1, xx]) * 2 + (e[yy + 1, xx - 1] - e[yy - 1, xx - 1]) + (e[yy + 1, xx + 1] - e[yy - 1, xx + 1])
            a += ix * ix
            b += iy * iy
            c += ix * iy
    f = np.float32
    scale = f(1.0) / (f(4 * 7) * f(255.0))
    s4 = scale * scale * scale * scale
    af, bf, cf = a.astype(f), b.astype(f), c.astype(f)
    s = af + bf
    return ((af * bf - cf * cf) - (HARRIS_K * s) * s) * s4


def umax_table():
    umax = [0] * (HALF_PATCH + 2)
    vmax = math.floor(np.float32(HALF_PATCH) * np.sqrt(np.float32(2.0)) / np.float32(2) + np.float32(1))
    vmin = math.ceil(np.float32(HALF_PATCH) * np.sqrt(np.float32(2.0)) / np.float32(2))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(float(HALF_PATCH * HALF_PATCH - v * v)))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees), float32 arithmetic without contraction."""
    f = np.float32
    p1, p3 = f(0.9997878412794807) * f(180 / np.pi), f(-0.3258083974640975) * f(180 / np.pi)
    p5, p7 = f(0.1555786518463281) * f(180 / np.pi), f(-0.04432655554792128) * f(180 / np.pi)
    eps = f(2.220446049250313e-16)
    y, x = f(y), f(x)
    ax, ay = abs(x), abs(y)
    if ax >= ay:
        c = ay / (ax + eps)
        c2 = c * c
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    else:
        c = ax / (ay + eps)
        c2 = c * c
        a = f(90.0) - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    if x < 0:
        a = f(180.0) - a
    if y < 0:
        a = f(360.0) - a
    return f(a)


def ic_angle(ext, x, y, umax):
    e = ext.astype(np.int64)
    cx, cy = x + BORDER, y + BORDER
    m01 = m10 = 0
    u = np.arange(-HALF_PATCH, HALF_PATCH + 1)
    m10 += int((u * e[cy, cx - HALF_PATCH:cx + HALF_PATCH + 1]).sum())
    for v in range(1, HALF_PATCH + 1):
        d = umax[v]
        u = np.arange(-d, d + 1)
        plus, minus = e[cy + v, cx - d:cx + d + 1], e[cy - v, cx - d:cx + d + 1]
        m01 += v * int((plus - minus).sum())
        m10 += int((u * (plus + minus)).sum())
    return fast_atan2(np.float32(m01), np.float32(m10))


# ------------------------------------------------------------------------------------------------ descriptors
def gauss_kernel():
    """getGaussianKernel(7, 2, CV_32F): float32 taps."""
    x = np.arange(7) - 3
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    return (k / k.sum()).astype(np.float32)


def _fma(a, b, c):
    """float32 fused multiply-add (the product of two float32 values is exact in float64)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def blur_level(ext):
    """GaussianBlur(level, level, (7, 7), 2, 2, BORDER_REFLECT_101) as ORB applies it to a level inside the pyramid buffer.
    The level is a sub-matrix without BORDER_ISOLATED, so cv::GaussianBlur skips its bit-exact fixed-point branch and runs
    sepFilter2D with the float32 kernel: uint8 -> float32 rows, float32 columns, cvRound.  The float32 operation order is
    that of OpenCV 4.13's AVX2 build (pinned against cv2.sepFilter2D on 48 M random pixels, no difference):
        row    s = k0*p0, then s = s + k_i*p_i for taps 1..6 in order -- fused (FMA) in the blocks of 32 columns the
               vector loop covers, multiply-then-add in the remaining (width mod 32) columns
        column s = k3*r0, then s = fma(k_{3+i}, r_{+i} + r_{-i}, s) for i = 1..3
        out    = saturate(round-half-even(s))
    Returns the blurred level WITH its (unblurred) 32 px frame, which is what the descriptor's tests read near the edge."""
    k = gauss_kernel()
    f = np.float32
    e = ext.astype(f)
    H, W = e.shape
    w = W - 2 * BORDER
    x0 = BORDER - 3                                         # first column the row filter reads for output column 0
    src = e[BORDER - 3:H - BORDER + 3, :]                   # the rows the column filter needs
    rf = k[0] * src[:, x0:x0 + w]
    rn = rf.copy()
    for i in range(1, 7):
        t = src[:, x0 + i:x0 + i + w]
        rf = _fma(np.full_like(t, k[i]), t, rf)
        rn = (rn + (k[i] * t).astype(f)).astype(f)
    body = (w // 32) * 32
    r = rf
    r[:, body:] = rn[:, body:]
    h = r.shape[0] - 6
    s = (k[3] * r[3:3 + h]).astype(f)
    for i in range(1, 4):
        t = (r[3 + i:3 + i + h] + r[3 - i:3 - i + h]).astype(f)
        s = _fma(np.full_like(t, k[3 + i]), t, s)
    out = ext.copy()
    out[BORDER:H - BORDER, BORDER:W - BORDER] = np.clip(np.rint(s), 0, 255).astype(np.uint8)
    return out


def describe(blurred_ext, x, y, angle_deg, pattern):
    """computeOrbDescriptors for one keypoint (WTA_K = 2): 32 bytes."""
    f = np.float32
    ang = f(angle_deg) * f(np.pi / 180.0)
    a, b = f(math.cos(float(ang))), f(math.sin(float(ang)))          # (see orb_sincos: OpenCV calls cosf / sinf)
    return describe_ab(blurred_ext, x, y, a, b, pattern)


def describe_ab(blurred_ext, x, y, a, b, pattern):
    f = np.float32
    px, py = pattern[:, 0].astype(f), pattern[:, 1].astype(f)
    xr = px * a - py * b
    yr = px * b + py * a
    ix = np.rint(xr).astype(np.int64)
    iy = np.rint(yr).astype(np.int64)
    vals = blurred_ext[y + BORDER + iy, x + BORDER + ix].astype(np.int64)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)              # 256 tests: point 2k against point 2k+1
    return np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1).reshape(32)


def sincosf(angle_deg):
    """(cosf, sinf) of the keypoint angle as computeOrbDescriptors forms them (glibc's float routines through ctypes: the same
    libm OpenCV's wheel resolves at run time)."""
    import ctypes
    import ctypes.util
    m = ctypes.CDLL(ctypes.util.find_library("m"))
    m.cosf.restype = m.sinf.restype = ctypes.c_float
    m.cosf.argtypes = m.sinf.argtypes = [ctypes.c_float]
    ang = np.float32(angle_deg) * np.float32(np.pi / 180.0)
    return np.float32(m.cosf(float(ang))), np.float32(m.sinf(float(ang)))


# ------------------------------------------------------------------------------------------------ whole pipeline
def detect_and_compute(gray, pattern, nfeatures=NFEATURES):
    """Returns (keypoints, descriptors): keypoints = float32 array (n, 6): x, y, size, angle, response, octave."""
    pyr = pyramid(gray)
    nper = features_per_level(nfeatures)
    scales = level_scales()
    umax = umax_table()
    kept = []
    for lv, (img, ext) in enumerate(pyr):
        kp = fast_detect(img)
        h, w = img.shape
        kp = [tuple(int(t) for t in r) for r in kp if EDGE <= r[0] < w - EDGE and EDGE <= r[1] < h - EDGE]
        kp = retain_best(kp, 2 * nper[lv], key=lambda r: r[2])
        if kp:
            hr = harris_response(ext, [r[0] for r in kp], [r[1] for r in kp])
            kp = [(r[0], r[1], float(v)) for r, v in zip(kp, hr)]
            kp = retain_best(kp, nper[lv], key=lambda r: np.float32(r[2]))
        kept.append(kp)
    kps, descs = [], []
    for lv, (img, ext) in enumerate(pyr):
        if not kept[lv]:
            continue
        blurred = blur_level(ext)
        sf = scales[lv]
        for (x, y, resp) in kept[lv]:
            ang = ic_angle(ext, x, y, umax)
            fx, fy = np.float32(x) * sf, np.float32(y) * sf
            inv = np.float32(1.0) / sf
            xi, yi = cv_round(fx * inv), cv_round(fy * inv)
            a, b = sincosf(ang)
            descs.append(describe_ab(blurred, xi, yi, a, b, pattern))
            kps.append((fx, fy, np.float32(PATCH) * sf, ang, np.float32(resp), lv))
    if not kps:
        return np.zeros((0, 6), np.float32), np.zeros((0, 32), np.uint8)
    return np.array(kps, np.float32), np.stack(descs)
