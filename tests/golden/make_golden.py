#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ from the LIVE reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):   PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every array written here is an output of the unmodified reference (ransac.py / homography.py
imported from /root/reference) on the inputs stored beside it.  The generating environment
(numpy / OpenBLAS / CPU) is recorded in meta.json: the float32 LAPACK results (per-hypothesis
SVD, float32 normal equations) are not reproducible across BLAS builds.
"""
import io
import json
import os
import platform
import sys
import warnings
from contextlib import redirect_stdout

import numpy as np

REF = os.environ.get("REFERENCE_DIR", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

import cv2  # noqa: E402
import homography as RH  # noqa: E402  (the reference)
import ransac as RR  # noqa: E402      (the reference)

OUT = os.path.dirname(os.path.abspath(__file__))


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def per_hypothesis(X, Y, idx, th, method):
    """Reference fit + computeLoss for every explicit sample set (no early exit)."""
    model = RR.HomoModel(th=th, d=50, n=4)
    K = len(idx)
    H = np.zeros((K, 3, 3), np.float32)
    Hinv = np.zeros((K, 3, 3), np.float32)
    counts = np.zeros(K, np.int32)
    finite = np.zeros(K, bool)
    for i in range(K):
        Hi = model.fit(X[:, idx[i]], Y[:, idx[i]])
        H[i] = Hi
        finite[i] = bool(np.all(np.isfinite(Hi)))
        try:
            if method != "fwd":
                Hinv[i] = np.linalg.inv(Hi)
            counts[i] = np.sum(model.computeLoss(X, Y, method) < th)
        except np.linalg.LinAlgError:
            finite[i] = False
    return H, Hinv, counts, finite


def ransac_case(X, Y, seed, th, d, k, method):
    M = X.shape[1]
    np.random.seed(seed)
    idx = np.random.randint(0, M, (k, 4))          # stream-identical to k sequential draws of 4
    np.random.seed(seed)
    model = RR.HomoModel(th=th, d=d, n=4)
    Hf, inl, cnt = quiet(RR.RANSAC(model, k=k).run, [X, Y], method=method)
    H, Hinv, counts, finite = per_hypothesis(X, Y, idx, th, method)
    return dict(idx=idx.astype(np.int32), H=H, Hinv=Hinv, counts=counts, finite=finite,
                run_H=np.asarray(Hf, np.float64), run_inliers=inl[0].astype(np.int64),
                run_count=np.int64(cnt), th=np.float64(th), d=np.float64(d), k=np.int64(k))


def synth_pair(seed, M, outlier_frac=0.5, noise=0.5):
    """The C3 generator of SURVEY.md 8(d)."""
    rs = np.random.RandomState(seed)
    X = (rs.rand(2, M) * np.array([[1920.0], [1080.0]])).astype(np.float32)
    Ht = np.array([[0.9, 0.05, 30], [-0.04, 1.1, -20], [1e-5, -2e-5, 1]])
    q = Ht @ np.vstack([X.astype(np.float64), np.ones(M)])
    Y = q[:2] / q[2] + rs.randn(2, M) * noise
    out = rs.rand(M) < outlier_frac
    Y[:, out] = rs.rand(2, int(out.sum())) * np.array([[1920.0], [1080.0]])
    return X, Y.astype(np.float32)


def main():
    meta = dict(numpy=np.__version__, cv2=cv2.__version__, python=platform.python_version(),
                machine=platform.machine(), processor=platform.processor(),
                blas=str(np.show_config(mode="dicts").get("Build Dependencies", {}).get("blas", {}).get("name")),
                reference="choice17/ransac_with_homography @ /root/reference (unmodified)")

    # ---- fixture points (matchespoints.npy is a pickled dict; homography.ipynb:463-467) ----
    a = np.load(os.path.join(REF, "matchespoints.npy"), allow_pickle=True).tolist()
    ptsA, ptsB = a["ptsA"].astype(np.float32), a["ptsB"].astype(np.float32)
    np.savez(os.path.join(OUT, "fixture_points.npz"), ptsA=ptsA, ptsB=ptsB)

    # ---- RANSAC on the fixture (config 1): example0 settings and app settings ----
    cases = {}
    X, Y = ptsA.T, ptsB.T
    for seed in (0, 1, 2):
        cases[f"A_s{seed}"] = ransac_case(X, Y, seed, 5, 70, 1000, "fwd")     # ransac.py:228-230
        cases[f"B_s{seed}"] = ransac_case(X, Y, seed, 4, 95, 1500, "fwd")     # app.py:364-367
    cases["R_s0"] = ransac_case(X, Y, 0, 5, 60, 300, "reproj")                # run()'s default method
    cases["K_s0"] = ransac_case(X, Y, 0, 5, 60, 300, "backward")              # regression.ipynb:561-568
    cases["E_s1"] = ransac_case(X, Y, 1, 5, 60, 1000, "fwd")                  # early exit: 115 needed
    flat = {f"{c}__{k}": v for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "ransac_fixture.npz"), **flat)

    # ---- synthetic pairs: per-hypothesis H / counts by the reference ----
    syn = {}
    for name, (M, K, th) in {"m2000": (2000, 256, 4.0), "m20000": (20000, 48, 3.0), "m37": (37, 64, 3.0)}.items():
        Xs, Ys = synth_pair(1, M)
        idx = np.random.RandomState(0).randint(0, M, (K, 4)).astype(np.int32)
        for method in ("fwd", "reproj"):
            H, Hinv, counts, finite = per_hypothesis(Xs, Ys, idx, th, method)
            syn.update({f"{name}_{method}__H": H, f"{name}_{method}__Hinv": Hinv,
                        f"{name}_{method}__counts": counts, f"{name}_{method}__finite": finite})
        syn.update({f"{name}__X": Xs, f"{name}__Y": Ys, f"{name}__idx": idx, f"{name}__th": np.float64(th)})
    np.savez_compressed(os.path.join(OUT, "ransac_synth.npz"), **syn)

    # ---- 4-point / M-point solves (S2, S5) ----
    u = np.array([[400, 100], [600, 200], [450, 620], [620, 700.0]])          # regression.ipynb:22848-22853
    v = np.array([[0, 0], [400, 0], [0, 400], [400, 400.0]])
    u2 = np.array([[50, 50], [470, 40], [600, 300], [90, 360.0]])             # homography.py:421-426
    v2 = np.array([[0, 0], [400, 0], [400, 780], [0, 780.0]])
    solves = dict(u=u, v=v, H_linear=RH.calcHomographyLinear(u, v),
                  H_svd=RH.calcHomography(u.astype(np.float32), v.astype(np.float32)),
                  u2=u2, v2=v2, H2_linear=RH.calcHomographyLinear(u2, v2))
    inl = cases["A_s0"]["run_inliers"]
    solves["refit_inliers"] = inl
    solves["refit_H"] = RH.calcHomographyLinear(ptsA[inl], ptsB[inl], True)
    np.savez(os.path.join(OUT, "solves.npz"), **solves)

    # ---- backward warps (configs 1, 2 at reduced size) ----
    img = cv2.imread(os.path.join(REF, "foto1A.jpg"))[:, :, ::-1]
    img = np.ascontiguousarray(img[::6, ::6])                                 # 114 x 171 x 3
    img2 = np.ascontiguousarray(cv2.imread(os.path.join(REF, "foto1B.jpg"))[:, :, ::-1][::6, ::6])
    Hs = {
        "stitch": np.array([[0.79, 0.055, 73.0], [-0.127, 0.93, 11.8], [-1.9e-4 * 6, -1.9e-5 * 6, 1]]),
        "rot": np.array([[0.8, -0.6, 40.3], [0.6, 0.8, -25.7], [2e-4, 1e-4, 1]]),
        "shrink": np.array([[0.41, 0.02, -3.5], [0.03, 0.37, 2.25], [1e-4, -3e-4, 1]]),
        "shift": np.array([[1.0, 0, 7.25], [0, 1.0, -5.5], [0, 0, 1]]),   # integer shifts hit x == w-1: IndexError in the reference
    }
    W = dict(img=img, img2=img2)
    for name, H in Hs.items():
        W[f"H_{name}"] = H
        for conv in ("nn", "bilinear"):
            o, mx, my = RH.wrapPerspective(img.copy(), H, conv)
            W[f"bbox_{name}_{conv}"] = o.astype(np.uint8)
            W[f"bbox_{name}_{conv}_off"] = np.array([mx, my], np.int64)
        o, mx, my = RH.transformImageH(img.copy(), H)
        assert o.dtype == np.uint8
        W[f"tih_{name}"] = o
    rgba = quiet(RH.addAlpha, img.copy(), "Rate", 0.2)
    o, mx, my = RH.wrapPerspective(rgba.copy(), Hs["stitch"], "bilinear")
    W["bbox_stitch_rgba"] = o                                                 # float64, 4 channels
    # Scanner: 4 clicked corners -> rectangle (app.py:350-359); box = [h, w]
    uu = np.array([[12.0, 150, 160, 8], [10, 6, 100, 106], [1, 1, 1, 1]])
    for tag, (hh, ww) in {"s1": (90, 64), "s2": (100, 110)}.items():
        vv = np.array([[0.0, ww, ww, 0], [0, 0, hh, hh], [1, 1, 1, 1]])
        W[f"scan_{tag}_u"], W[f"scan_{tag}_v"] = uu, vv
        Hsc = RH.calcHomographyLinear(uu.T[:, :2], vv.T[:, :2])
        W[f"scan_{tag}_H"] = Hsc
        for conv in ("nn", "bilinear"):
            o, _, _ = RH.wrapPerspectiveScan(img.copy(), Hsc, [hh, ww], conv)
            W[f"scan_{tag}_{conv}"] = o.astype(np.uint8)
            W[f"scan_{tag}_{conv}_ti"] = RH.transformImage(img.copy(), uu, vv, box=[hh, ww], method=conv)
        W[f"scan_{tag}_ti_nobox"] = RH.transformImage(img.copy(), uu, vv)
    # compose (homography.py:288-338): all four sign cases of (mx, my), paste and "Rate" blend
    for name, H in {"pp": Hs["stitch"], "nn_": np.array([[0.9, 0.05, -20.0], [0.02, 0.95, -9.0], [1e-4, 0, 1]]),
                    "np_": np.array([[0.9, 0.05, -20.0], [0.02, 0.95, 14.0], [1e-4, 0, 1]]),
                    "pn_": np.array([[0.9, 0.05, 20.0], [0.02, 0.95, -14.0], [1e-4, 0, 1]])}.items():
        W[f"pano_{name}_H"] = H
        W[f"pano_{name}_paste"] = quiet(RH.stitchPanorama, img2.copy(), img.copy(), H)
        W[f"pano_{name}_blend"] = quiet(RH.stitchPanorama, img2.copy(), img.copy(), H, blending="Rate", blendrate=0.2)
    np.savez_compressed(os.path.join(OUT, "warps.npz"), **W)

    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
