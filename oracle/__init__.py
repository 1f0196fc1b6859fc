"""CPU oracle for the geometric hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``ransac_with_homography_b200`` imports this package.  It is used by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs, as the
checker and as the timed CPU arm -- never as part of the product path.

Two layers:

* ``oracle.c`` (ctypes, built by ``make -C oracle``): scalar restatement of the reference's
  element-level arithmetic -- scoring chain, selection, backward warp -- plus a float64 4-point
  solve used as arbiter.  Bit-level semantics; deterministic on any x86-64 host.
* ``ref_numpy.py``: numpy restatement of the reference's numpy pipeline (same LAPACK/BLAS calls),
  used as the CPU baseline ("port") and to cross-check the C layer against the live reference.

Pinned against outputs of the live reference: ``tests/golden/*.npz`` (``make_golden.py``).
Reference citations are file:line in /root/reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

FWD, BACKWARD, REPROJ = 0, 1, 2
METHODS = {"fwd": FWD, "backward": BACKWARD, "reproj": REPROJ}


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (a few hundred ms).  Returns the .so path."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        fp, ip, u8p, dp, vp = (c.POINTER(c.c_float), c.POINTER(c.c_int32), c.POINTER(c.c_uint8),
                               c.POINTER(c.c_double), c.c_void_p)
        L.oracle_errors.argtypes = [fp, c.c_int, fp, fp, c.c_int, fp]
        L.oracle_errors.restype = None
        L.oracle_score_counts.argtypes = [fp, c.c_int, fp, fp, c.c_int, c.c_float, c.c_int, ip]
        L.oracle_score_counts.restype = None
        L.oracle_inlier_mask.argtypes = [fp, c.c_int, fp, fp, c.c_float, c.c_int, u8p]
        L.oracle_inlier_mask.restype = None
        L.oracle_select.argtypes = [ip, c.c_int, c.c_double, ip, ip]
        L.oracle_select.restype = c.c_int
        L.oracle_dlt4_f64.argtypes = [fp, ip, c.c_int, dp, u8p]
        L.oracle_dlt4_f64.restype = None
        L.oracle_warp.argtypes = [vp, c.c_int, c.c_int, c.c_int, c.c_int, dp, dp, c.c_int, dp, c.c_int,
                                  c.c_int, c.c_int, c.c_int, dp]
        L.oracle_warp.restype = None
        L.oracle_to_u8.argtypes = [dp, c.c_size_t, u8p]
        L.oracle_to_u8.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def pack_points(X, Y) -> np.ndarray:
    """[X (2|3,M), Y (2|3,M)] as the reference passes them (ransac.py:165) -> (M,4) float32 rows
    (x, y, x', y')."""
    X = np.asarray(X)
    Y = np.asarray(Y)
    return _f32(np.concatenate([X[:2].T, Y[:2].T], axis=1))


def errors(pts, H, Hinv=None, method="fwd") -> np.ndarray:
    """HomoModel.computeLoss (ransac.py:84-98) for one hypothesis; float32 (M,)."""
    pts = _f32(pts)
    H = _f32(H).reshape(9)
    Hi = _f32(Hinv).reshape(9) if Hinv is not None else None
    m = METHODS[method]
    assert m == FWD or Hi is not None
    err = np.empty(pts.shape[0], np.float32)
    lib().oracle_errors(_p(pts, ctypes.c_float), pts.shape[0], _p(H, ctypes.c_float),
                        _p(Hi, ctypes.c_float) if Hi is not None else None, m, _p(err, ctypes.c_float))
    return err


def score_counts(pts, H, th, Hinv=None, method="fwd") -> np.ndarray:
    """Inlier count per hypothesis (ransac.py:182-184); H (K,3,3) float32."""
    pts = _f32(pts)
    H = _f32(H).reshape(-1, 9)
    Hi = _f32(Hinv).reshape(-1, 9) if Hinv is not None else None
    m = METHODS[method]
    assert m == FWD or Hi is not None
    counts = np.empty(H.shape[0], np.int32)
    lib().oracle_score_counts(_p(pts, ctypes.c_float), pts.shape[0], _p(H, ctypes.c_float),
                              _p(Hi, ctypes.c_float) if Hi is not None else None, H.shape[0],
                              np.float32(th), m, _p(counts, ctypes.c_int32))
    return counts


def inlier_mask(pts, H, th, Hinv=None, method="fwd") -> np.ndarray:
    pts = _f32(pts)
    H = _f32(H).reshape(9)
    Hi = _f32(Hinv).reshape(9) if Hinv is not None else None
    mask = np.empty(pts.shape[0], np.uint8)
    lib().oracle_inlier_mask(_p(pts, ctypes.c_float), pts.shape[0], _p(H, ctypes.c_float),
                             _p(Hi, ctypes.c_float) if Hi is not None else None, np.float32(th),
                             METHODS[method], _p(mask, ctypes.c_uint8))
    return mask.astype(bool)


def select(counts, M, d, n=4):
    """Sequential winner rule of RANSAC.run (ransac.py:169-202).
    Returns (winner index or -1, best count, iterations the reference would evaluate)."""
    counts = np.ascontiguousarray(counts, np.int32)
    thr = M * d / 100 + n
    best = ctypes.c_int32(0)
    nev = ctypes.c_int32(0)
    win = lib().oracle_select(_p(counts, ctypes.c_int32), counts.shape[0], float(thr),
                              ctypes.byref(best), ctypes.byref(nev))
    return int(win), int(best.value), int(nev.value)


def dlt4_f64(pts, idx):
    """float64 arbiter for the per-hypothesis 4-point solve (homography.py:4-14,71-88).
    Returns (H (K,3,3) float64 with h22 = 1, degenerate flags (K,) bool)."""
    pts = _f32(pts)
    idx = np.ascontiguousarray(idx, np.int32).reshape(-1, 4)
    K = idx.shape[0]
    H = np.empty((K, 9), np.float64)
    flags = np.empty(K, np.uint8)
    lib().oracle_dlt4_f64(_p(pts, ctypes.c_float), _p(idx, ctypes.c_int32), K, _p(H, ctypes.c_double),
                          _p(flags, ctypes.c_uint8))
    return H.reshape(K, 3, 3), flags.astype(bool)


def refit_f64(pts, mask) -> np.ndarray:
    """float64 least-squares arbiter for the final refit (homography.py:48-69,90-105): the
    minimiser of |A h - b| over the inlier rows, by numpy's SVD-based lstsq in float64 (the
    reference evaluates inv(A^T A) A^T b in float32)."""
    p = np.asarray(pts, np.float64)[np.asarray(mask, bool)]
    x, y, X, Y = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    n = p.shape[0]
    A = np.zeros((2 * n, 8))
    b = np.zeros(2 * n)
    A[0::2, 0], A[0::2, 1], A[0::2, 2] = x, y, 1
    A[0::2, 6], A[0::2, 7] = -x * X, -y * X
    A[1::2, 3], A[1::2, 4], A[1::2, 5] = x, y, 1
    A[1::2, 6], A[1::2, 7] = -x * Y, -y * Y
    b[0::2], b[1::2] = X, Y
    # column equilibration keeps lstsq's rcond from truncating the small singular values
    s = np.linalg.norm(A, axis=0)
    s[s == 0] = 1.0
    h = np.linalg.lstsq(A / s, b, rcond=None)[0] / s
    return np.append(h, 1.0).reshape(3, 3)


def bbox_grid(H, h, w, boundary=0):
    """Output grid of wrapPerspective (homography.py:143-167): returns (gx, gy, min_x, min_y)."""
    H = np.asarray(H, np.float64)
    bnd = np.array([[0, w - 1, w - 1, 0], [0, 0, h - 1, h - 1], [1.0, 1, 1, 1]])
    q = H @ bnd
    q = q / q[-1, :]
    max_x, min_x = int(np.max(q[0])), int(np.min(q[0]))
    max_y, min_y = int(np.max(q[1])), int(np.min(q[1]))
    if boundary:
        min_x, min_y = max(min_x, 0), max(min_y, 0)
    gx = np.linspace(min_x, max_x, max_x - min_x + 1)
    gy = np.linspace(min_y, max_y, max_y - min_y + 1)
    return gx, gy, min_x, min_y


def scan_grid(res):
    """Output grid of wrapPerspectiveScan (homography.py:188-198)."""
    h, w = res
    return np.linspace(0, w, w), np.linspace(0, h, h)


def warp(img, invH, gx, gy, bounds_hw, sampler="bilinear") -> np.ndarray:
    """Backward warp over an explicit grid; float64 (oh, ow, C).  ``img`` is NOT modified;
    pixel (0,0) is read as 0 (homography.py:112-116,126-130)."""
    img = np.ascontiguousarray(img)
    assert img.dtype in (np.uint8, np.float32) and img.ndim == 3
    ih, iw, C = img.shape
    invH = np.ascontiguousarray(invH, np.float64).reshape(9)
    gx = np.ascontiguousarray(gx, np.float64)
    gy = np.ascontiguousarray(gy, np.float64)
    out = np.empty((gy.shape[0], gx.shape[0], C), np.float64)
    lib().oracle_warp(img.ctypes.data, int(img.dtype == np.float32), ih, iw, C,
                      _p(invH, ctypes.c_double), _p(gx, ctypes.c_double), gx.shape[0],
                      _p(gy, ctypes.c_double), gy.shape[0], int(bounds_hw[0]), int(bounds_hw[1]),
                      int(sampler == "bilinear"), _p(out, ctypes.c_double))
    return out


def to_u8(v) -> np.ndarray:
    v = np.ascontiguousarray(v, np.float64)
    out = np.empty(v.shape, np.uint8)
    lib().oracle_to_u8(_p(v, ctypes.c_double), v.size, _p(out, ctypes.c_uint8))
    return out


def wrapPerspective(img, H, convert="nn", boundary=0):
    """homography.py:142-184 (crop=True branch): returns (float64 image, min_x, min_y)."""
    h, w, _ = img.shape
    gx, gy, mx, my = bbox_grid(H, h, w, boundary)
    out = warp(img, np.linalg.inv(np.asarray(H, np.float64)), gx, gy, (h, w), convert)
    return out, mx, my


def wrapPerspectiveScan(img, H, res, convert="nn"):
    """homography.py:186-209: the sampler is given ``res`` as its bounds."""
    gx, gy = scan_grid(res)
    out = warp(img, np.linalg.inv(np.asarray(H, np.float64)), gx, gy, res, convert)
    return out, 0, 0


_ref_mods = None


def reference_modules():
    """(ransac, homography): the modules of the UNMODIFIED reference as copied to oracle/_ref/ by oracle/make_ref.py
    (digests checked against the manifest), or None when the copy is absent.  They are imported under the private names
    ``_ref_ransac`` / ``_ref_homography`` so that they never shadow the product's modules.  bench.py's reference arm and
    cpu_baseline time these; nothing in the product imports them."""
    global _ref_mods
    if _ref_mods is not None:
        return _ref_mods or None
    import hashlib
    import importlib
    import json
    import sys
    d = os.path.join(_HERE, "_ref")
    try:
        with open(os.path.join(d, "MANIFEST.json")) as f:
            man = json.load(f)["sha256"]
        for name in ("ransac.py", "homography.py"):
            with open(os.path.join(d, name), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != man[name]:
                    raise ValueError(name + " differs from the manifest")
    except (OSError, ValueError, KeyError):
        _ref_mods = ()
        return None
    saved = {k: sys.modules.pop(k, None) for k in ("ransac", "homography")}
    sys.path.insert(0, d)
    dwb, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        hg = importlib.import_module("homography")
        rr = importlib.import_module("ransac")           # does `from homography import ...`
    finally:
        sys.dont_write_bytecode = dwb
        sys.path.remove(d)
        for k in ("ransac", "homography"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    sys.modules["_ref_ransac"], sys.modules["_ref_homography"] = rr, hg
    _ref_mods = (rr, hg)
    return _ref_mods

