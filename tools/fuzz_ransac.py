"""Randomised end-to-end check of RANSAC.run (solve, score, select / early exit, mask, refit) against the oracle's
protocol on random problems: counts from OUR per-hypothesis H through the oracle's reference-order chain, the
sequential winner rule incl. early exit, the inlier mask of the winner and the float64 least-squares refit.
    python tools/fuzz_ransac.py [cases] [seed]"""
import io, os, sys
from contextlib import redirect_stdout
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from ransac_with_homography_b200 import device
from ransac_with_homography_b200.ransac import HomoModel, RANSAC


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    skipped = 0
    for c in range(cases):
        M = int(rs.choice([8, 40, 185, 1000, 2049, int(rs.randint(8, 4000))]))
        K = int(rs.choice([1, 7, 300, 1000, int(rs.randint(1, 3000))]))
        scale = float(rs.choice([100.0, 1000.0, 3000.0]))
        th = float(rs.choice([0.5, 3.0, 5.0]))
        d = float(rs.choice([10, 45, 70, 95, 99]))
        method = ["fwd", "backward", "reproj"][rs.randint(0, 3)]
        X = rs.uniform(0, scale, (M, 2))
        Ht = np.array([[rs.uniform(0.8, 1.2), rs.uniform(-0.1, 0.1), rs.uniform(-0.05, 0.05) * scale],
                       [rs.uniform(-0.1, 0.1), rs.uniform(0.8, 1.2), rs.uniform(-0.05, 0.05) * scale],
                       [rs.uniform(-1, 1) * 3e-2 / scale, rs.uniform(-1, 1) * 3e-2 / scale, 1.0]])
        Yh = np.c_[X, np.ones(M)] @ Ht.T
        Y = Yh[:, :2] / Yh[:, 2:3] + rs.normal(0, 0.4, (M, 2))
        out = rs.rand(M) < rs.choice([0.0, 0.3, 0.6])
        Y[out] = rs.uniform(0, scale, (int(out.sum()), 2))
        Xf, Yf = X.astype(np.float32), Y.astype(np.float32)
        idx = rs.randint(0, M, (K, 4)).astype(np.int32)
        model = HomoModel(th=th, d=d, n=4)
        try:
            with redirect_stdout(io.StringIO()):
                H, inl, cnt, det = RANSAC(model, k=K).run([Xf.T, Yf.T], method=method, idx=idx, return_details=True)
        except (ValueError, np.linalg.LinAlgError):
            skipped += 1          # no model / singular refit: the reference raises too
            continue
        pts = oracle.pack_points(Xf.T, Yf.T)
        Hs, Hi, flags = device.dlt4(torch.from_numpy(pts).cuda(), torch.from_numpy(idx).cuda(), want_inv=True)
        Hs, Hi, live = Hs.cpu().numpy(), Hi.cpu().numpy(), ~flags.cpu().numpy().astype(bool)
        want = np.zeros(K, np.int32)
        if live.any():
            want[live] = oracle.score_counts(pts, Hs[live], th, Hi[live] if method != "fwd" else None, method)
        msg = None
        if not np.array_equal(det["counts"], want):
            msg = f"{int((det['counts'] != want).sum())} counts differ"
        else:
            win, best, _ = oracle.select(want, M, d)
            if det["winner"] != win or int(cnt) != best:
                msg = f"winner {det['winner']} / {win}, count {int(cnt)} / {best}"
            else:
                mask = oracle.inlier_mask(pts, det["H_sample"], th, Hi[win] if method != "fwd" else None, method)
                if not np.array_equal(inl[0], np.where(mask)[0]):
                    msg = "inlier sets differ"
                elif mask.sum() >= 8:
                    T = oracle.refit_f64(pts, mask)
                    err = np.linalg.norm(H - T) / np.linalg.norm(T)
                    if not err < 1e-4:
                        msg = f"refit rel err {err:.2e} ({int(mask.sum())} inliers)"
        if msg:
            bad += 1
            print(f"FAIL case {c}: M={M} K={K} scale={scale} th={th} d={d} {method}: {msg}")
    print(f"{cases} cases, {skipped} raised like the reference, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
