"""Randomised check of the fused stitch compose (K5) and the Hamming matcher (K6) against the numpy restatements
(pinned to the live reference / to OpenCV): identical bytes / identical match lists.
    python tools/fuzz_compose.py [cases] [seed]"""
import io, os, sys
from contextlib import redirect_stdout
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_numpy as port
from ransac_with_homography_b200 import homography as hg
from ransac_with_homography_b200.ransac import match_hamming


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for c in range(cases):
        h, w = int(rs.randint(8, 300)), int(rs.randint(8, 400))
        hq, wq = int(rs.randint(8, 300)), int(rs.randint(8, 400))
        imgT = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        imgQ = rs.randint(0, 256, (hq, wq, 3)).astype(np.uint8)
        a = rs.uniform(-0.5, 0.5)
        H = np.array([[np.cos(a) * rs.uniform(0.5, 1.6), -np.sin(a), rs.uniform(-150, 150)],
                      [np.sin(a), np.cos(a) * rs.uniform(0.5, 1.6), rs.uniform(-150, 150)],
                      [rs.uniform(-8e-4, 8e-4), rs.uniform(-8e-4, 8e-4), 1.0]])
        blending, rate = [(False, 0.2), ("Rate", 0.2), ("Rate", float(rs.uniform(0, 1)))][rs.randint(0, 3)]
        try:
            ref = port.stitch_panorama(imgQ.copy(), imgT.copy(), H, blending, rate)
        except Exception:
            continue              # placements the reference itself cannot compose
        with redirect_stdout(io.StringIO()):
            out = hg.stitchPanorama(imgQ.copy(), imgT.copy(), H, blending=blending, blendrate=rate)
        if out.shape != ref.shape or not np.array_equal(out, ref):
            bad += 1
            print(f"FAIL compose case {c}: T {h}x{w} Q {hq}x{wq} blending {blending} {rate}: shapes {out.shape} {ref.shape}, "
                  f"{int((out != ref).sum()) if out.shape == ref.shape else -1} bytes differ")
        nA, nB, nb = int(rs.randint(1, 700)), int(rs.randint(1, 700)), int(rs.choice([4, 32, 32, 64]))
        A = rs.randint(0, 256, (nA, nb)).astype(np.uint8)
        B = rs.randint(0, 256, (nB, nb)).astype(np.uint8)
        k = int(rs.randint(0, min(nA, nB) + 1))
        if k:
            B[rs.randint(0, nB, k)] = A[rs.randint(0, nA, k)]        # duplicates: ties
        cross = bool(rs.randint(0, 2))
        m = match_hamming(A, B, crossCheck=cross)
        q, t, d = port.bf_match_hamming(A, B, cross)
        if [x.queryIdx for x in m] != q.tolist() or [x.trainIdx for x in m] != t.tolist() or [int(x.distance) for x in m] != d.tolist():
            bad += 1
            print(f"FAIL match case {c}: {nA}x{nB}x{nb} cross {cross}")
    print(f"{cases} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
