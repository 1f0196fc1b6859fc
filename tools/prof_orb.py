"""A few ORB calls on a 1080p frame (for ncu / timing).   python tools/prof_orb.py [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ransac_with_homography_b200.ransac import orb_detect_and_compute
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rs = np.random.RandomState(2)
h, w = 1080, 1920
base = rs.randint(0, 256, (h // 8 + 1, w // 8 + 1)).astype(np.float32)
gray = np.clip(np.kron(base, np.ones((8, 8), np.float32))[:h, :w] + rs.randn(h, w).astype(np.float32) * 4, 0, 255).astype(np.uint8)
import cv2
photo = cv2.cvtColor(cv2.resize(cv2.imread("tests/golden/foto1A.jpg"), (w, h)), cv2.COLOR_BGR2GRAY)
for name, img in (("blocky", gray), ("photo", photo)):
    for _ in range(2):
        orb_detect_and_compute(img, 500, as_arrays=True)
    t0 = time.perf_counter()
    for _ in range(reps):
        kp, ds = orb_detect_and_compute(img, 500, as_arrays=True)
    print(name, "ms per call", (time.perf_counter() - t0) / reps * 1e3, "keypoints", len(kp), flush=True)
