#!/usr/bin/env python
"""Golden records of the components either side of the hot path, from the LIVE
reference and from OpenCV itself (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ext.py

Writes tests/golden/orb.npz (cv2.ORB_create().detectAndCompute on the reference's images, a noise image and a small
image, and the matched point lists stitching() feeds to RANSAC) and tests/golden/ext.npz: cylindericlMap / cylindricalWarp and cylindericalTransform('nn', 'fwd') of the
unmodified reference (homography.py:340-416) on the 114 x 171 fixture crop, and cv2.remap on random maps.
"""
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("REFERENCE_DIR", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

import cv2  # noqa: E402
import homography as RH  # noqa: E402  (the reference)

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    Z = {}
    img = np.load(os.path.join(OUT, "warps.npz"))["img"]
    Z["cyl_img"] = img
    for f in (100, 300, 1600):
        Z[f"cyl_map_f{f}"] = RH.cylindericlMap(img.copy(), f=f)          # BGRA; pixels cv2.remap does not write are arbitrary
    K = np.array([[120.0, 3.0, 80.5], [0, 95.0, 60.25], [0, 0, 1]])
    Z["cyl_K"] = K
    Z["cyl_warp_K"] = RH.cylindricalWarp(img.copy(), K)
    o, mx, my = RH.cylindericalTransform(img.copy(), 100, method="nn")
    Z["cyl_transform_nn"], Z["cyl_transform_nn_off"] = o, np.array([mx, my])
    o, mx, my = RH.cylindericalTransform(img.copy(), 37.5, xc=0.3, yc=0.8, method="nn")
    Z["cyl_transform_nn2"], Z["cyl_transform_nn2_off"] = o, np.array([mx, my])
    # cv2.remap itself on random maps (all channel counts the transparent branch covers), into a non-zero destination
    rs = np.random.RandomState(5)
    for ch in (1, 2, 4):
        src = rs.randint(0, 256, (37, 53, ch)).astype(np.uint8)
        mx_ = (rs.rand(41, 47) * 60 - 4).astype(np.float32)
        my_ = (rs.rand(41, 47) * 44 - 4).astype(np.float32)
        mx_[::7, ::5] = np.round(mx_[::7, ::5])                          # exact integers and halves
        my_[::3, ::4] = np.round(my_[::3, ::4] * 2) / 2
        mx_[0, :4] = [np.nan, np.inf, -np.inf, 1e30]
        dst = rs.randint(0, 256, (41, 47, ch)).astype(np.uint8)
        ref = cv2.remap(np.squeeze(src) if ch == 1 else src, mx_, my_, cv2.INTER_AREA, dst=(np.squeeze(dst.copy()) if ch == 1 else dst.copy()),
                        borderMode=cv2.BORDER_TRANSPARENT)
        Z[f"remap_src{ch}"], Z[f"remap_mx{ch}"], Z[f"remap_my{ch}"], Z[f"remap_dst{ch}"] = src, mx_, my_, dst
        Z[f"remap_ref{ch}"] = ref.reshape(41, 47, ch)
    np.savez_compressed(os.path.join(OUT, "ext.npz"), **Z)
    print("wrote ext.npz:", sorted(Z))

    # ---- ORB: cv2.ORB_create().detectAndCompute exactly as stitching() calls it (ransac.py:252-257), and the matcher after it ----
    Q = {}

    def kparr(kps):
        return np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)

    rs = np.random.RandomState(9)
    noise = rs.randint(0, 256, (211, 317, 3)).astype(np.uint8)
    small = cv2.resize(cv2.imread(os.path.join(OUT, "cv.jpg")), (163, 121))[:, :, ::-1].copy()
    cases = {f.split(".")[0]: np.ascontiguousarray(cv2.imread(os.path.join(OUT, f))[:, :, ::-1]) for f in ("foto1A.jpg", "foto1B.jpg", "cv.jpg", "notebook.jpg")}
    cases["noise"], cases["small"] = noise, small
    Q["noise_rgb"], Q["small_rgb"] = noise, small
    for name, rgb in cases.items():
        gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
        for nf in ((500, 60) if name in ("foto1A", "noise") else (500,)):
            kps, des = cv2.ORB_create(nf).detectAndCompute(gray, None) if nf != 500 else cv2.ORB_create().detectAndCompute(gray, None)
            tag = name if nf == 500 else f"{name}_n{nf}"
            Q[f"{tag}__kp"] = kparr(kps)
            Q[f"{tag}__desc"] = des if des is not None else np.zeros((0, 32), np.uint8)
            print("orb", tag, gray.shape, len(kps))
    # the correspondences stitching() hands to RANSAC for the reference's own pair (ransac.example1)
    ga, gb = (cv2.cvtColor(cases[n], cv2.COLOR_RGB2GRAY) for n in ("foto1A", "foto1B"))
    ka, fa = cv2.ORB_create().detectAndCompute(ga, None)
    kb, fb = cv2.ORB_create().detectAndCompute(gb, None)
    m = sorted(cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(fa, fb), key=lambda x: x.distance)
    pa, pb = np.float32([k.pt for k in ka]), np.float32([k.pt for k in kb])
    Q["pair_ptsA"] = np.float32([pa[x.queryIdx] for x in m])
    Q["pair_ptsB"] = np.float32([pb[x.trainIdx] for x in m])
    np.savez_compressed(os.path.join(OUT, "orb.npz"), **Q)
    print("wrote orb.npz:", len(Q), "arrays;", len(m), "matches on the reference's pair")


if __name__ == "__main__":
    main()
