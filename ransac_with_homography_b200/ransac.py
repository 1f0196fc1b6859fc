"""Drop-in for the reference's ``ransac`` module (reference: ransac.py): ``Model``, ``HomoModel``,
``RANSAC`` and ``stitching`` with the reference's signatures, defaults, return types and side
effects; hypothesis generation, scoring, selection, inlier mask and refit run on the GPU through
the C ABI (include/b200geo.h).  No CPU fallback.

Parallel form of the loop (reference ransac.py:176-202): the k sample sets are drawn up front from
the same global ``np.random`` stream the reference consumes (``randint(0, M, (k, n))`` is
stream-identical to k sequential ``randint(0, M, n)`` calls; after an early exit the stream is rewound
to where the reference's ``break`` leaves it), all k hypotheses are solved and
scored at once, and the winner is the first iteration meeting ``M*d/100 + n`` if any (the
reference's ``break``), else the lowest-index maximum (the reference's strict ``>``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .homography import calcHomographyLinear, calcHomography, stitchPanorama, cylindericlMap  # noqa: F401

LVL = 0


def DEBUG(*args):
    if LVL >= 1:
        print("[DEBUG]", *args)


def _rows(X, Y):
    """[X (2|3, M), Y (2|3, M)] -> (M, 4) float32 rows (x, y, x', y') for the C ABI."""
    X, Y = np.asarray(X), np.asarray(Y)
    out = np.empty((X.shape[1], 4), dtype=np.float32)
    out[:, 0:2] = X[:2].T
    out[:, 2:4] = Y[:2].T
    return out


class Model(object):
    __slots__ = ('val', 'th', 'd', 'n')

    def fit(self, X, Y):
        raise NotImplementedError

    def fwd(self, X):
        raise NotImplementedError

    def dist(self, predY, trueY):
        raise NotImplementedError


class HomoModel(Model):
    """Homography model: threshold th (px), required inlier percentage d, sample size n
    (reference ransac.py:22-98)."""

    def __init__(self, th=5, d=50, n=4):
        self.th = th
        self.d = d
        self.n = n
        self.val = np.empty((3, 3), dtype=np.float32)

    def fit(self, X, Y, collective=False):
        """X, Y: (2|3, 4) -> 4-point solve (float32); (2|3, m>4) with collective=True -> least
        squares over all columns (float64).  Reference ransac.py:30-53."""
        nx, mx = X.shape
        ny, my = Y.shape
        assert ((mx == my) and (mx == self.n)) or ((mx == my) and (mx > self.n) and collective), "invalid data size should be %d" % self.n
        assert (nx == ny) and nx in [2, 3], "invalid input dimension for row numbers"
        if collective:
            self.val = calcHomographyLinear(X.T[:, :2], Y.T[:, :2], True)
        else:
            self.val = calcHomography(X.T[:, :2], Y.T[:, :2], False)
        return self.val

    def fwd(self, X):
        """val @ [X;1] / (w + 1e-10), float32 (3, M) (reference ransac.py:55-64)."""
        nx, mx = X.shape
        assert nx in [2, 3], "invalid input dimension for row numbers"
        return _project(self.val, X)

    def reproj(self, Y):
        """inv(val) @ [Y;1] / (w + 1e-10) (reference ransac.py:66-76)."""
        ny, my = Y.shape
        assert ny in [2, 3], "invalid input dimension for row numbers"
        return _project(np.linalg.inv(self.val), Y)

    def dist(self, predY, trueY):
        """sqrt(sum((pred - true)^2, axis=0)), float32 (reference ransac.py:78-82)."""
        a = np.ascontiguousarray(predY, dtype=np.float32)
        b = np.ascontiguousarray(trueY, dtype=np.float32)
        assert a.shape == b.shape and a.ndim == 2
        out = np.empty(a.shape[1], dtype=np.float32)
        ctx = _cabi.context()
        _cabi.check(ctx.lib.b200geo_dist_host(ctx.handle, _cabi.ptr(a), _cabi.ptr(b), a.shape[0], a.shape[1], _cabi.ptr(out)))
        return out

    def computeLoss(self, X, Y, method="reproj"):
        """Per-point loss of the current model, float32 (M,), on the GPU (reference ransac.py:84-98)."""
        if method not in _cabi.METHODS:
            exit("Invalid method!")
        ctx = _cabi.context()
        rows = _rows(X, Y)
        H = np.ascontiguousarray(self.val, dtype=np.float32)
        Hinv = None
        if method != "fwd":
            Hinv = np.ascontiguousarray(np.linalg.inv(np.asarray(self.val, dtype=np.float32)), dtype=np.float32)
        err = np.empty(rows.shape[0], dtype=np.float32)
        _cabi.check(ctx.lib.b200geo_loss_host(ctx.handle, _cabi.ptr(rows), rows.shape[0], _cabi.ptr(H),
                                              _cabi.ptr(Hinv), _cabi.METHODS[method], _cabi.ptr(err)))
        return err


def _project(Hm, P):
    """(Hm @ [P;1]) / (row 2 + 1e-10) on the GPU: the homogeneous product and divide of
    HomoModel.fwd / reproj (reference ransac.py:58-64), float32 (3, M)."""
    P = np.asarray(P)
    m = P.shape[1]
    x = np.ones((3, m), dtype=np.float32)
    x[:P.shape[0], :] = P
    Hm = np.ascontiguousarray(Hm, dtype=np.float32)
    out = np.empty((3, m), dtype=np.float32)
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_project_host(ctx.handle, _cabi.ptr(Hm), _cabi.ptr(x), m, _cabi.ptr(out)))
    return out


class RANSAC(object):
    """reference ransac.py:137-213"""

    __slots__ = ('model', 'th', 'd', 'n', 'k')

    def __init__(self, model, k=1000):
        self.model = model
        self.th = model.th
        self.d = model.d
        self.n = model.n
        self.k = k

    def computeLoss(self, X, Y, method="reproj"):
        return self.model.computeLoss(X, Y, method)

    def run(self, data, method="reproj", *, idx=None, return_details=False):
        """RANSAC over data = [X (2|3, M), Y (2|3, M)].

        Returns ``(H, inliers, count)`` like the reference: H (3,3) float64 refit on the inliers of
        the winning hypothesis (also stored in ``model.val``), ``inliers`` the tuple
        ``(int64 indices,)`` that ``np.where`` returns, ``count`` an ``np.int64``.

        Keyword-only extensions: ``idx`` (k, 4) explicit sample sets instead of drawing from
        ``np.random``; ``return_details`` appends a dict (winner, per-hypothesis counts, H of the
        winning sample)."""
        X, Y = data
        nx, mx = X.shape
        ny, my = Y.shape
        assert mx == my, "data observation not consistent!"
        if method not in _cabi.METHODS:
            exit("Invalid method!")
        assert self.n == 4, "the homography model samples 4 correspondences"
        d = mx * self.d / 100
        k = int(self.k)
        rng_state = None
        if idx is None:
            rng_state = np.random.get_state()
            idx = np.random.randint(0, mx, (k, self.n))      # same stream as k draws of n (ransac.py:177)
        idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1, 4)
        k = idx.shape[0]
        rows = _rows(X, Y)
        ctx = _cabi.context()
        H_best = np.empty(9, np.float32)
        H = np.empty(9, np.float64)
        mask = np.empty(mx, np.uint8)
        winner = np.zeros(1, np.int32)
        count = np.zeros(1, np.int32)
        counts = np.empty(k, np.int32) if return_details else None       # K values back only when asked for
        last_count = np.zeros(1, np.int32)
        rc = ctx.lib.b200geo_ransac_host(ctx.handle, _cabi.ptr(rows), mx, _cabi.ptr(idx), k, float(np.float32(self.th)),
                                         float(d + self.n), _cabi.METHODS[method], _cabi.ptr(H_best), _cabi.ptr(H),
                                         _cabi.ptr(mask), _cabi.ptr(winner), _cabi.ptr(count),
                                         _cabi.ptr(counts) if counts is not None else None, _cabi.ptr(last_count))
        if rc == _cabi.ERR_NOMODEL:
            # the reference reaches np.where(None) here (ransac.py:206)
            raise ValueError("RANSAC: no hypothesis has any inlier (" + _cabi.last_error() + ")")
        _cabi.check(rc)
        w = int(winner[0])
        totalfitLen = np.int64(count[0])
        if rng_state is not None and count[0] >= (d + self.n) and w + 1 < k:
            # the reference stops drawing at its `break` (ransac.py:177,186-190): leave the global stream where it
            # leaves it -- after (w + 1) draws of n -- so that seeded sequences of calls stay in step
            np.random.set_state(rng_state)
            np.random.randint(0, mx, (w + 1, self.n))
        # the reference's warning tests the LAST evaluated iteration's count (ransac.py:203-204)
        # (count[0] is the winner's count; if it met the threshold the loop broke there, else it ran to sample set k-1)
        last_cnt = int(count[0]) if count[0] >= (d + self.n) else int(last_count[0])
        if last_cnt < (d + self.n):
            print("Warning:: fitting model does not exceed required threshold %d vs %d" % (totalfitLen, d + self.n))
        inliers = np.where(mask.astype(bool))
        DEBUG("Fitting final model using all inliers")
        finalModel = H.reshape(3, 3)
        self.model.val = finalModel
        if return_details:
            return finalModel, inliers, totalfitLen, dict(winner=w, counts=counts, H_sample=H_best.reshape(3, 3))
        return finalModel, inliers, totalfitLen


def example0(FILE="matchespoints.npy"):
    """reference ransac.py:220-232 (the fixture is a pickled dict: allow_pickle is required on numpy >= 1.16.3)"""
    a = np.load(FILE, allow_pickle=True).tolist()
    ptsA, ptsB = a['ptsA'].T, a['ptsB'].T
    model = HomoModel(th=5, d=70, n=4)
    ransac = RANSAC(model, k=1000)
    H, inliers, _len = ransac.run([ptsA, ptsB], method="fwd")
    print(H, _len, '\n')
    print(0)


class Match(object):
    """One match, with cv2.DMatch's attribute names (the reference reads .queryIdx / .trainIdx / .distance,
    ransac.py:260-266)."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx, trainIdx, distance):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = int(queryIdx), int(trainIdx), 0, float(distance)

    def __repr__(self):
        return f"Match(queryIdx={self.queryIdx}, trainIdx={self.trainIdx}, distance={self.distance})"


def match_hamming(featuresA, featuresB, crossCheck=True):
    """cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck).match(featuresA, featuresB) (reference ransac.py:258-259) on the
    GPU (K6, b200geo_match_hamming_host): for every descriptor of A the nearest descriptor of B by Hamming distance,
    first index on ties, kept only if mutual when crossCheck.  Returns a list of Match in query order, like OpenCV.
    None inputs (ORB found nothing) give an empty list."""
    if featuresA is None or featuresB is None:
        return []
    A = np.ascontiguousarray(featuresA, dtype=np.uint8)
    B = np.ascontiguousarray(featuresB, dtype=np.uint8)
    assert A.ndim == 2 and B.ndim == 2 and (A.shape[1] == B.shape[1] or A.shape[0] == 0 or B.shape[0] == 0)
    nA, nB = A.shape[0], B.shape[0]
    if nA == 0 or nB == 0:
        return []
    tr, ds = np.empty(nA, np.int32), np.empty(nA, np.int32)
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_match_hamming_host(ctx.handle, _cabi.ptr(A), nA, _cabi.ptr(B), nB, A.shape[1], 1 if crossCheck else 0,
                                                   _cabi.ptr(tr), _cabi.ptr(ds)))
    return [Match(i, tr[i], ds[i]) for i in np.nonzero(tr >= 0)[0]]


class KeyPoint:
    """The fields of cv2.KeyPoint the reference and OpenCV's ORB fill (ransac.py:261-262 reads .pt)."""
    __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

    def __init__(self, x, y, size, angle, response, octave):
        self.pt = (float(x), float(y))
        self.size, self.angle, self.response, self.octave, self.class_id = float(size), float(angle), float(response), int(octave), -1

    def __repr__(self):
        return f"KeyPoint(pt={self.pt}, size={self.size}, angle={self.angle}, response={self.response}, octave={self.octave})"


def orb_detect_and_compute(img, nfeatures=500, as_arrays=False):
    """cv2.ORB_create(nfeatures).detectAndCompute(gray, None) (reference ransac.py:255-257) on the GPU (K7,
    b200geo_orb_host): same keypoints in the same order, same 32-byte descriptors as OpenCV 4.13.

    img: (h, w) uint8 grey image, or (h, w, 3) uint8 RGB -- the cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) of
    ransac.py:252-253 is then done on the device (bit-exact).  Returns (keypoints, descriptors) like OpenCV: a list of
    KeyPoint and an (n, 32) uint8 array, or ((), None) when nothing is found; as_arrays=True returns the keypoints as an
    (n, 6) float32 array (x, y, size, angle, response, octave) instead."""
    a = np.ascontiguousarray(img)
    assert a.dtype == np.uint8 and (a.ndim == 2 or (a.ndim == 3 and a.shape[2] == 3))
    h, w = a.shape[:2]
    ch = 1 if a.ndim == 2 else 3
    cap = 2 * int(nfeatures) + 256
    kp = np.empty((cap, 6), np.float32)
    ds = np.empty((cap, 32), np.uint8)
    n = C.c_int()
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_orb_host(ctx.handle, _cabi.ptr(a), h, w, w * ch, ch, int(nfeatures), _cabi.ptr(kp), _cabi.ptr(ds), cap, C.byref(n)))
    kp, ds = kp[:n.value].copy(), ds[:n.value].copy()
    if as_arrays:
        return kp, (ds if n.value else None)
    if n.value == 0:
        return (), None
    return [KeyPoint(*r) for r in kp], ds


def stitching(trainImg, queryImg,
              ransacMet="fwd",
              th=5,
              d=70,
              n=4,
              k=1000,
              blending=False,
              blendrate=0.2,
              mode=None,
              override=0,
              cylinderT=1):
    """Stitch trainImg into queryImg's frame (reference ransac.py:235-283; called from app.py:364).
    Grey conversion + ORB detection and description (K7), the brute-force Hamming matcher with cross-check (K6), RANSAC
    (K1-K3), the backward warp and the compose (K4 / K5) all run on the GPU; only the alternates the reference itself
    delegates to OpenCV (mode != None: cv2.findHomography; override: cv2.Stitcher) stay there."""
    if override == 0:
        (kpsA, featuresA) = orb_detect_and_compute(trainImg)            # cvtColor(RGB2GRAY) + ORB_create().detectAndCompute
        (kpsB, featuresB) = orb_detect_and_compute(queryImg)

        matches = sorted(match_hamming(featuresA, featuresB, crossCheck=True), key=lambda m: m.distance)

        kpsA = np.float32([kp.pt for kp in kpsA])
        kpsB = np.float32([kp.pt for kp in kpsB])
        ptsA = np.float32([kpsA[m.queryIdx] for m in matches])
        ptsB = np.float32([kpsB[m.trainIdx] for m in matches])

        if mode is None:
            model = HomoModel(th=th, d=d, n=4)
            ransac = RANSAC(model, k=k)
            H, inliers, _len = ransac.run([ptsA.T, ptsB.T], method=ransacMet)
        else:
            import cv2
            reprojThresh = 4
            (H, status) = cv2.findHomography(ptsA, ptsB, cv2.RANSAC, reprojThresh)

        imgn = stitchPanorama(queryImg, trainImg, H=H, blending=blending, blendrate=blendrate)
    else:
        import cv2
        stitcher = cv2.Stitcher_create()
        (status, imgn) = stitcher.stitch([trainImg, queryImg])
    return imgn
