"""Stage-by-stage comparison of the GPU ORB (b200geo_orb_host) with the CPU oracle on one image.   python tools/debug_orb.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orb_numpy as O
from ransac_with_homography_b200.ransac import orb_detect_and_compute
pat = np.load("tests/golden/orb_pattern.npy").astype(np.int64)
z = np.load("tests/golden/orb.npz")
for name in ("small", "noise"):
    rgb = z[name + "_rgb"]
    kp, ds = orb_detect_and_compute(rgb, 500, as_arrays=True)
    ref, rds = z[name + "__kp"], z[name + "__desc"]
    print(name, "n", len(kp), len(ref))
    print(" per level got", np.bincount(kp[:, 5].astype(int), minlength=8), "ref", np.bincount(ref[:, 5].astype(int), minlength=8))
    sg = {tuple(r) for r in kp[:, [0, 1, 5]]}; sr = {tuple(r) for r in ref[:, [0, 1, 5]]}
    print(" position sets: common", len(sg & sr), "only gpu", len(sg - sr), "only ref", len(sr - sg))
    m = min(len(kp), len(ref))
    for j, nm in enumerate(["x", "y", "size", "angle", "response", "octave"]):
        print("  ", nm, "mismatch (same index)", int((kp[:m, j] != ref[:m, j]).sum()))
    print(" first rows gpu", kp[:3], "ref", ref[:3])
    if ds is not None and len(ds) == len(rds):
        print(" desc rows differing", int((ds != rds).any(1).sum()))
    # by position: compare response/angle/desc for common keypoints
    idx = {tuple(r): i for i, r in enumerate(ref[:, [0, 1, 5]])}
    ra = rr = rd = 0; c = 0
    for i, r in enumerate(kp[:, [0, 1, 5]]):
        j = idx.get(tuple(r))
        if j is None: continue
        c += 1; ra += kp[i, 3] != ref[j, 3]; rr += kp[i, 4] != ref[j, 4]; rd += (ds[i] != rds[j]).any()
    print(" common", c, "angle diff", int(ra), "response diff", int(rr), "desc diff", int(rd))
