"""Randomised parity check of the fast scoring kernels (K2: division-free test + proven ambiguity band + exact fix-up)
against the reference-order exact kernel (itself bit-identical to the oracle): random set sizes, coordinate scales,
thresholds, loss modes, hypotheses from random 4-point samples (K1) and from perturbed / wild homographies.
Counts must be identical.      python tools/fuzz_score.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ransac_with_homography_b200 import _cabi, device


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = _cabi.context(0)
    bad = 0
    for c in range(cases):
        M = int(rs.choice([1, 3, 17, 100, 1023, 1024, 1025, 2000, 5000, int(rs.randint(1, 9000))]))
        K = int(rs.choice([1, 5, 255, 256, 257, 1024, 3000, int(rs.randint(1, 5000))]))
        scale = float(rs.choice([1.0, 30.0, 1000.0, 4000.0, 2.0e4]))
        th = float(rs.choice([0.05, 0.5, 3.0, 5.0, 40.0, scale * 0.3]))
        method = ["fwd", "backward", "reproj"][rs.randint(0, 3)]
        X = rs.uniform(0, scale, (M, 2))
        Ht = np.array([[rs.uniform(0.7, 1.3), rs.uniform(-0.2, 0.2), rs.uniform(-0.1, 0.1) * scale],
                       [rs.uniform(-0.2, 0.2), rs.uniform(0.7, 1.3), rs.uniform(-0.1, 0.1) * scale],
                       [rs.uniform(-1, 1) * 1e-1 / scale, rs.uniform(-1, 1) * 1e-1 / scale, 1.0]])
        Yh = np.c_[X, np.ones(M)] @ Ht.T
        Y = Yh[:, :2] / Yh[:, 2:3] + rs.normal(0, th * 0.5, (M, 2))
        out = rs.rand(M) < 0.4
        Y[out] = rs.uniform(0, scale, (int(out.sum()), 2))
        pts = torch.from_numpy(np.ascontiguousarray(np.c_[X, Y], np.float32)).cuda()
        kind = rs.randint(0, 3)
        if kind == 0 and M >= 4:       # hypotheses from 4-point samples (degenerate ones included)
            idx = torch.from_numpy(rs.randint(0, M, (K, 4)).astype(np.int32)).cuda()
            H, Hinv, flags = device.dlt4(pts, idx, want_inv=True)
        else:                          # the true map, perturbed mildly or wildly
            amp = 1e-3 if kind == 1 else 0.3
            Hn = (Ht[None] * (1 + rs.normal(0, amp, (K, 3, 3)))).astype(np.float32)
            H = torch.from_numpy(np.ascontiguousarray(Hn)).cuda()
            Hinv = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(Hn.astype(np.float64)).astype(np.float32))).cuda()
            flags = None
        res = {}
        for m in ("fast", "exact"):
            ctx.force_exact(m == "exact")
            res[m] = device.score(pts, H, th, method, Hinv=Hinv, flags=flags).cpu().numpy()
        ctx.force_exact(False)
        if not np.array_equal(res["fast"], res["exact"]):
            bad += 1
            d = res["fast"] != res["exact"]
            print(f"FAIL case {c}: M={M} K={K} scale={scale} th={th} {method} kind={kind}: {int(d.sum())} counts differ, max |d| "
                  f"{int(np.abs(res['fast'] - res['exact']).max())}")
    print(f"{cases} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
