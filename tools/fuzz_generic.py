"""Randomised check of the drop-in warp API (wrapPerspective / wrapPerspectiveScan, nn and bilinear, uint8 RGB and
float32 RGBA sources -- the float64 kernel K4-generic) against the C oracle (pinned to the live reference):
bit-identical values, identical offsets.      python tools/fuzz_generic.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle
from ransac_with_homography_b200 import homography as hg
from fuzz_warp import rand_h


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    skipped = 0
    for c in range(cases):
        h, w = int(rs.randint(2, 160)), int(rs.randint(2, 200))
        rgba = rs.rand() < 0.3
        img = (rs.rand(h, w, 4) * 255).astype(np.float32) if rgba else rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        H = rand_h(rs, max(h, 8), max(w, 8))
        conv = "nn" if rs.rand() < 0.5 else "bilinear"
        try:
            if rs.rand() < 0.5:
                ref, rmx, rmy = oracle.wrapPerspective(img.copy(), H, conv)
                if ref.size > 4e6:
                    skipped += 1
                    continue
                out, mx, my = hg.wrapPerspective(img.copy(), H, convert=conv)
            else:
                res = [int(rs.randint(2, 200)), int(rs.randint(2, 260))]
                ref, rmx, rmy = oracle.wrapPerspectiveScan(img.copy(), H, res, conv)
                out, mx, my = hg.wrapPerspectiveScan(img.copy(), H, res, convert=conv)
        except (ValueError, MemoryError, np.linalg.LinAlgError):
            skipped += 1
            continue
        ok = out.shape == ref.shape and (mx, my) == (rmx, rmy) and np.array_equal(np.asarray(out, np.float64), ref)
        if not ok:
            bad += 1
            n = int((np.asarray(out, np.float64) != ref).sum()) if out.shape == ref.shape else -1
            print(f"FAIL case {c}: {h}x{w} rgba={rgba} {conv}: shape {out.shape} vs {ref.shape}, offsets {(mx, my)} vs {(rmx, rmy)}, {n} values differ")
    print(f"{cases} cases ({skipped} skipped: too large or raised like numpy), {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
