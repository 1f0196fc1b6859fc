#!/usr/bin/env python
"""Full-size golden records from the LIVE reference (build container only; imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_full.py

The reference's own fixture images (foto1A.jpg / foto1B.jpg: ransac.py:288-294; cv.jpg: the Scanner
photo of the README; notebook.jpg: homography.py:418-426) are copied next to this script, and the
unmodified reference is run on them at the sizes of BASELINE configs 1 and 2.  Full-size outputs
(up to 3508 x 2480 x 3) are too large to commit, so each is pinned by its shape, its SHA-256 and a
sparse sample (every 9973rd byte); tests/test_oracle_golden.py requires the CPU oracle to reproduce
the digests, and the GPU tests compare the kernels with the oracle's full arrays on the same inputs.
Also written: the reference's design matrices (calc_corresp*), its boundary=1 / crop=False outputs,
calcHomography(collective=True), HomoModel.fit / reproj, and the global-RNG position after a run.
"""
import hashlib
import io
import json
import os
import shutil
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
IMAGES = ["foto1A.jpg", "foto1B.jpg", "cv.jpg", "notebook.jpg"]
STEP = 9973


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def record(a):
    a = np.ascontiguousarray(a)
    return {"shape": list(a.shape), "dtype": str(a.dtype), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "sample": a.reshape(-1).view(np.uint8)[::STEP].tolist()}


def main():
    for f in IMAGES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
    G = {"images": {}, "cases": {}}
    img = {}
    for f in IMAGES:
        img[f] = np.ascontiguousarray(cv2.imread(os.path.join(OUT, f))[:, :, ::-1])      # RGB as ransac.example1 / app.py
        G["images"][f] = record(img[f])                                                    # decoder check on the GPU box
    small = np.load(os.path.join(OUT, "ransac_fixture.npz"))
    H_run = small["A_s0__run_H"]                         # the reference's own RANSAC result on the fixture (maps A -> B)
    G["H_stitch"] = H_run.tolist()
    A, B = img["foto1A.jpg"], img["foto1B.jpg"]
    c = G["cases"]
    # ---- config 1: stitch warp + compose (ransac.py:279: stitchPanorama(queryImg = B, trainImg = A, H)) ----
    o, mx, my = RH.transformImageH(A.copy(), H_run)
    c["c1_tih_bilinear"] = dict(record(o), off=[int(mx), int(my)])
    o, mx, my = RH.wrapPerspective(A.copy(), H_run, "nn")
    c["c1_bbox_nn"] = dict(record(o.astype(np.uint8)), off=[int(mx), int(my)])
    c["c1_pano_paste"] = record(quiet(RH.stitchPanorama, B.copy(), A.copy(), H_run))
    c["c1_pano_blend"] = record(quiet(RH.stitchPanorama, B.copy(), A.copy(), H_run, blending="Rate", blendrate=0.2))
    # the app's order (README): foto1B warped into foto1A's frame
    Hi = np.linalg.inv(H_run)
    Hi = Hi / Hi[2, 2]
    G["H_stitch_inv"] = Hi.tolist()
    c["c1_pano_paste_swapped"] = record(quiet(RH.stitchPanorama, A.copy(), B.copy(), Hi))
    # ---- config 2: Scanner, cv.jpg quad -> A4 at 300 dpi and the app's native A4 (app.py:21) ----
    cv = img["cv.jpg"]
    u = np.array([[100.0, 560, 740, 80], [160, 125, 490, 555], [1, 1, 1, 1]])
    G["scan_u"] = u.tolist()
    for tag, (hh, ww) in {"a4_300dpi": (3508, 2480), "a4_app": (1188, 840)}.items():
        v = np.array([[0.0, ww, ww, 0], [0, 0, hh, hh], [1, 1, 1, 1]])
        Hs = RH.calcHomographyLinear(u.T[:, :2], v.T[:, :2])                   # the reference's float32 normal-equation solve
        for conv in ("bilinear", "nn"):
            o = RH.transformImage(cv.copy(), u, v, box=[hh, ww], method=conv)
            c[f"c2_{tag}_{conv}"] = dict(record(o), box=[hh, ww], H=np.asarray(Hs, np.float64).tolist())
    v = np.array([[0.0, 840, 840, 0], [0, 0, 1188, 1188], [1, 1, 1, 1]])
    c["c2_nobox_bilinear"] = dict(record(RH.transformImage(cv.copy(), u, v)),
                                  H=np.asarray(RH.calcHomographyLinear(u.T[:, :2], v.T[:, :2]), np.float64).tolist())
    # ---- homography.example0: notebook.jpg ----
    nb = img["notebook.jpg"]
    u0 = np.array([[50, 470, 600.0, 90], [50, 40, 300.0, 360], [1, 1, 1, 1.0]])
    v0 = np.array([[0, 400, 400.0, 0], [0, 0, 780.0, 780], [1, 1, 1, 1]])
    H0 = RH.calcHomographyLinear(u0.T[:, :2], v0.T[:, :2])
    G["H_notebook"] = np.asarray(H0, np.float64).tolist()
    o, mx, my = RH.wrapPerspective(nb.copy(), H0, convert="bilinear")
    c["notebook_bilinear"] = dict(record(o.astype(np.uint8)), off=[int(mx), int(my)])
    with open(os.path.join(OUT, "full_size.json"), "w") as f:
        json.dump(G, f)

    # ---- small API vectors ----
    a = np.load(os.path.join(REF, "matchespoints.npy"), allow_pickle=True).tolist()
    ptsA, ptsB = a["ptsA"].astype(np.float32), a["ptsB"].astype(np.float32)
    Z = {}
    u4, v4 = ptsA[[3, 50, 99, 140]], ptsB[[3, 50, 99, 140]]
    Z["u4"], Z["v4"] = u4, v4
    Z["corresp"] = RH.calc_corresp(u4, v4)
    la, lb = RH.calc_correspLinear(u4, v4)
    Z["corresp_linear_A"], Z["corresp_linear_b"] = la, lb
    inl = small["A_s0__run_inliers"]
    Z["inl"] = inl
    Z["corresp_collective"] = RH.calc_correspCollective(ptsA[inl], ptsB[inl])
    ca, cb = RH.calc_correspLinearCollective(ptsA[inl], ptsB[inl])
    Z["corresp_linear_collective_A"], Z["corresp_linear_collective_b"] = ca, cb
    Z["H_collective_svd"] = RH.calcHomography(ptsA[inl], ptsB[inl], True)                 # S6: float32 SVD minimiser
    # the same minimiser in float64 (the arbiter for S6: the reference's float32 SVD of the raw matrix is noisy)
    M64 = RH.calc_correspCollective(ptsA[inl].astype(np.float64), ptsB[inl].astype(np.float64)).astype(np.float64)
    vt = np.linalg.svd(M64)[2][-1].reshape(3, 3)
    Z["H_collective_svd_f64"] = vt / vt[2, 2]
    # HomoModel.fit (4 points and collective) / fwd / reproj
    m = RR.HomoModel(th=5, d=70, n=4)
    Z["fit4"] = m.fit(u4.T, v4.T)
    Z["fit4_fwd"] = m.fwd(ptsA.T)
    Z["fit4_reproj"] = m.reproj(ptsB.T)
    Z["fit_collective"] = m.fit(ptsA[inl].T, ptsB[inl].T, collective=True)
    # wrapPerspective boundary=1 and crop=False on the decimated fixture image of warps.npz
    W = np.load(os.path.join(OUT, "warps.npz"))
    simg = W["img"]
    for name in ("stitch", "shrink"):
        H = W[f"H_{name}"]
        for conv in ("nn", "bilinear"):
            o, mx, my = RH.wrapPerspective(simg.copy(), H, conv, boundary=1)
            Z[f"b1_{name}_{conv}"] = o.astype(np.uint8)
            Z[f"b1_{name}_{conv}_off"] = np.array([mx, my], np.int64)
    Hn = np.array([[0.9, 0.05, -20.0], [0.02, 0.95, -9.0], [1e-4, 0, 1]])           # negative offsets: boundary=1 clamps them
    Z["H_neg"] = Hn
    o, mx, my = RH.wrapPerspective(simg.copy(), Hn, "bilinear", boundary=1)
    Z["b1_neg_bilinear"], Z["b1_neg_bilinear_off"] = o.astype(np.uint8), np.array([mx, my], np.int64)
    Hp = np.array([[0.97, 0.0, 0.0], [0.0, 0.96, 0.0], [1e-5, 0, 1]])                 # crop=False only works for min_x = min_y = 0
    Z["H_nocrop"] = Hp
    o, mx, my = RH.wrapPerspective(simg.copy(), Hp, "nn", crop=False)
    Z["nocrop_nn"], Z["nocrop_nn_off"] = o, np.array([mx, my], np.int64)
    # position of the global np.random stream after a run with an early exit (E_s1) and without (A_s0)
    for case, (seed, th, d, k) in {"E_s1": (1, 5, 60, 1000), "A_s0": (0, 5, 70, 1000)}.items():
        np.random.seed(seed)
        quiet(RR.RANSAC(RR.HomoModel(th=th, d=d, n=4), k=k).run, [ptsA.T, ptsB.T], method="fwd")
        Z[f"rng_after_{case}"] = np.random.randint(0, 1 << 30, 8)
    np.savez_compressed(os.path.join(OUT, "api.npz"), **Z)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
