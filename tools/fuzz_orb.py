"""Randomised parity of the GPU ORB (K7, b200geo_orb_host) against OpenCV itself on the machine it runs on: random sizes, content
(noise, blocks, blurred noise, gradients, crops and rescalings of the reference's photographs), feature budgets, grey and RGB input.
Keypoints must be equal as float32 bit patterns in OpenCV's order, descriptors byte for byte.   python tools/fuzz_orb.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from ransac_with_homography_b200.ransac import orb_detect_and_compute

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    photos = [cv2.imread(os.path.join(ROOT, "tests", "golden", f))[:, :, ::-1] for f in ("foto1A.jpg", "foto1B.jpg", "cv.jpg", "notebook.jpg")]
    bad = empty = 0
    worst = None
    for c in range(cases):
        h, w = int(rs.randint(40, 900)), int(rs.randint(40, 1300))
        kind = rs.randint(0, 6)
        if kind == 0:
            img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        elif kind == 1:
            b = int(rs.choice([3, 5, 8, 16]))
            base = rs.randint(0, 256, (h // b + 1, w // b + 1, 3)).astype(np.float32)
            img = np.clip(np.kron(base, np.ones((b, b, 1), np.float32))[:h, :w] + rs.randn(h, w, 3) * rs.uniform(0, 12), 0, 255).astype(np.uint8)
        elif kind == 2:
            img = cv2.GaussianBlur(rs.randint(0, 256, (h, w, 3)).astype(np.uint8), (0, 0), rs.uniform(0.6, 2.5))
        elif kind == 3:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([(xx * rs.uniform(0.05, 1) + yy * rs.uniform(0.05, 1)) % 256] * 3, -1)
            img = np.clip(img + rs.randn(h, w, 3) * 8, 0, 255).astype(np.uint8)
        else:
            p = photos[rs.randint(0, len(photos))]
            y0, x0 = rs.randint(0, p.shape[0] // 2), rs.randint(0, p.shape[1] // 2)
            crop = p[y0:y0 + rs.randint(60, p.shape[0] - y0 + 1), x0:x0 + rs.randint(60, p.shape[1] - x0 + 1)]
            img = cv2.resize(crop, (w, h), interpolation=cv2.INTER_AREA if rs.rand() < 0.5 else cv2.INTER_CUBIC)
        img = np.ascontiguousarray(img)
        nf = int(rs.choice([500, 500, 100, 1000, 37, 2000]))
        use_gray = rs.rand() < 0.4
        gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
        kps, des = cv2.ORB_create(nf).detectAndCompute(gray, None)
        ref = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)
        kp, ds = orb_detect_and_compute(gray if use_gray else img, nf, as_arrays=True)
        if len(ref) == 0:
            empty += 1
        ok = kp.shape == ref.shape and np.array_equal(kp.view(np.uint32), ref.view(np.uint32)) and \
            (len(ref) == 0 or np.array_equal(ds, des))
        if not ok:
            bad += 1
            info = {"case": c, "shape": (h, w), "kind": int(kind), "nf": nf, "n": (len(kp), len(ref))}
            if kp.shape == ref.shape and len(ref):
                info["kp_fields"] = int((kp.view(np.uint32) != ref.view(np.uint32)).sum())
                info["desc_bits"] = int(np.unpackbits(ds ^ des).sum())
            if worst is None:
                worst = info
            print("MISMATCH", info, flush=True)
    print(f"{cases} cases, {bad} failures, {empty} without keypoints" + (f"; first: {worst}" if worst else ""))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
