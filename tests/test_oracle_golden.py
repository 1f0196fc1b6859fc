"""CPU: the oracle (oracle/oracle.c + oracle/ref_numpy.py) against the golden vectors produced by
the live reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import oracle
from oracle import ref_numpy as port
from conftest import FIXTURE_CASES, case_method, descriptors as _descriptors


@pytest.mark.parametrize("name", FIXTURE_CASES)
def test_scoring_chain_bit_exact_on_fixture(golden, name):
    A, B = golden.points()
    c = golden.case("ransac_fixture", name)
    pts = oracle.pack_points(A.T, B.T)
    m = case_method(name)
    ok = c["finite"]
    counts = oracle.score_counts(pts, c["H"], float(c["th"]), c["Hinv"] if m != "fwd" else None, m)
    assert np.array_equal(counts[ok], c["counts"][ok])


@pytest.mark.parametrize("name", ["m37", "m2000", "m20000"])
@pytest.mark.parametrize("method", ["fwd", "reproj"])
def test_scoring_chain_bit_exact_on_synthetic(golden, name, method):
    z = golden.file("ransac_synth")
    pts = oracle.pack_points(z[f"{name}__X"], z[f"{name}__Y"])
    ok = z[f"{name}_{method}__finite"]
    counts = oracle.score_counts(pts, z[f"{name}_{method}__H"], float(z[f"{name}__th"]),
                                 z[f"{name}_{method}__Hinv"] if method != "fwd" else None, method)
    assert np.array_equal(counts[ok], z[f"{name}_{method}__counts"][ok])


@pytest.mark.parametrize("name", FIXTURE_CASES)
def test_selection_and_inlier_set_match_reference_run(golden, name):
    """Per-hypothesis reference counts -> oracle selection rule -> the reference run's own result."""
    A, B = golden.points()
    c = golden.case("ransac_fixture", name)
    m = case_method(name)
    win, best, nev = oracle.select(c["counts"], A.shape[0], float(c["d"]))
    assert win >= 0 and best == int(c["run_count"])
    pts = oracle.pack_points(A.T, B.T)
    mask = oracle.inlier_mask(pts, c["H"][win], float(c["th"]), c["Hinv"][win] if m != "fwd" else None, m)
    assert np.array_equal(np.where(mask)[0], c["run_inliers"])
    thr = A.shape[0] * float(c["d"]) / 100 + 4
    if c["counts"][win] >= thr:      # early exit: the reference stopped before k iterations (E_s1, K_s0)
        assert nev == win + 1 and name in ("E_s1", "K_s0")
    else:
        assert nev == int(c["k"])


def test_port_reproduces_reference_run(golden):
    """ref_numpy (same LAPACK calls as the reference) reproduces the reference run on this host."""
    A, B = golden.points()
    for name in ["A_s0", "B_s1", "E_s1"]:
        c = golden.case("ransac_fixture", name)
        H, inl, cnt, counts, win = port.ransac_run(A.T, B.T, float(c["th"]), float(c["d"]), 4, int(c["k"]),
                                                   case_method(name), idx=c["idx"])
        assert cnt == int(c["run_count"])
        assert np.array_equal(inl[0], c["run_inliers"])
        np.testing.assert_allclose(H, c["run_H"], rtol=1e-4, atol=1e-6)


def test_dlt4_truth_vs_reference_svd(golden):
    """The float64 arbiter agrees with the reference's float32 SVD solve to the reference's own
    accuracy (p99 of the elementwise relative error ~ 6.5e-3 on the fixture)."""
    A, B = golden.points()
    c = golden.case("ransac_fixture", "A_s0")
    pts = oracle.pack_points(A.T, B.T)
    H64, flags = oracle.dlt4_f64(pts, c["idx"])
    dup = np.array([len(set(r)) < 4 for r in c["idx"].tolist()])
    assert np.array_equal(flags, dup)
    good = ~flags & c["finite"]
    rel = np.linalg.norm((H64[good] - c["H"][good]).reshape(-1, 9), axis=1) / np.linalg.norm(H64[good].reshape(-1, 9), axis=1)
    assert np.median(rel) < 1e-4 and np.quantile(rel, 0.9) < 5e-3


def test_refit_truth_vs_reference(golden):
    A, B = golden.points()
    s = golden.file("solves")
    pts = oracle.pack_points(A.T, B.T)
    mask = np.zeros(A.shape[0], bool)
    mask[s["refit_inliers"]] = True
    H = oracle.refit_f64(pts, mask)
    # the reference's float32 normal equations are only good to 2-3 digits (cond ~ 1e15)
    corners = np.array([[0, 1023, 1023, 0], [0, 0, 682, 682], [1, 1, 1, 1.0]])
    a, b = H @ corners, s["refit_H"] @ corners
    d = np.abs(a[:2] / a[2] - b[:2] / b[2]).max()
    assert d < 40.0
    # and it must be a better least-squares solution than the reference's
    def resid(Hm):
        p = pts[mask].astype(np.float64)
        w = Hm[2, 0] * p[:, 0] + Hm[2, 1] * p[:, 1] + 1
        rx = Hm[0, 0] * p[:, 0] + Hm[0, 1] * p[:, 1] + Hm[0, 2] - p[:, 2] * w
        ry = Hm[1, 0] * p[:, 0] + Hm[1, 1] * p[:, 1] + Hm[1, 2] - p[:, 3] * w
        return float(np.sum(rx * rx + ry * ry))
    assert resid(H) <= resid(s["refit_H"]) * (1 + 1e-9)


def test_4pt_known_answers(golden):
    """regression.ipynb:22848-22853 / 22895-22897 / 22938-22952: transfer error of the 4 corners."""
    s = golden.file("solves")
    u, v = s["u"], s["v"]
    pts = np.concatenate([u, v], 1).astype(np.float32)
    H, flags = oracle.dlt4_f64(pts, np.array([[0, 1, 2, 3]], np.int32))
    assert not flags[0]
    q = H[0] @ np.vstack([u.T, np.ones(4)])
    assert np.abs(q[:2] / q[2] - v.T).max() < 1e-9
    recorded = np.array([[2.0881958, -0.20079422, -815.1875], [-0.345853806, 0.691696167, 69.1699219],
                         [8.86917114e-05, -2.96704471e-04, 1]])
    np.testing.assert_allclose(H[0], recorded, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(H[0], s["H_linear"], rtol=2e-3, atol=2e-3)


WARP_H = ["stitch", "rot", "shrink", "shift"]


@pytest.mark.parametrize("name", WARP_H)
@pytest.mark.parametrize("conv", ["nn", "bilinear"])
def test_warp_bbox_matches_reference(golden, name, conv):
    w = golden.file("warps")
    out, mx, my = oracle.wrapPerspective(w["img"], w[f"H_{name}"], conv)
    assert [mx, my] == w[f"bbox_{name}_{conv}_off"].tolist()
    assert np.array_equal(oracle.to_u8(out), w[f"bbox_{name}_{conv}"])
    if conv == "bilinear":
        assert np.array_equal(oracle.to_u8(out), w[f"tih_{name}"])


def test_warp_rgba_matches_reference(golden):
    w = golden.file("warps")
    rgba = np.zeros(w["img"].shape[:2] + (4,), np.float32)
    rgba[:, :, :3] = w["img"]
    rgba[:, :, 3] = 0.2 + 1e-10
    out, _, _ = oracle.wrapPerspective(rgba, w["H_stitch"], "bilinear")
    assert np.array_equal(out, w["bbox_stitch_rgba"])


@pytest.mark.parametrize("tag", ["s1", "s2"])
@pytest.mark.parametrize("conv", ["nn", "bilinear"])
def test_warp_scan_matches_reference(golden, tag, conv):
    w = golden.file("warps")
    v = w[f"scan_{tag}_v"]
    res = [int(v[1, 2]), int(v[0, 2])]
    out, _, _ = oracle.wrapPerspectiveScan(w["img"], w[f"scan_{tag}_H"], res, conv)
    assert np.array_equal(oracle.to_u8(out), w[f"scan_{tag}_{conv}"])


def test_port_warp_matches_reference(golden):
    w = golden.file("warps")
    out, mx, my = port.warp_bbox(w["img"].copy(), w["H_rot"], True)
    assert np.array_equal(out.astype(np.uint8), w["bbox_rot_bilinear"])
    out, _, _ = port.warp_scan(w["img"].copy(), w["scan_s1_H"], [90, 64], False)
    assert np.array_equal(out, w["scan_s1_nn"])


@pytest.mark.parametrize("name", ["pp", "nn_", "np_", "pn_"])
def test_port_stitch_panorama_matches_reference(golden, name):
    """The numpy restatement of stitchPanorama (paste and 'Rate' blend) against the live reference's canvases,
    in the four sign cases of the warp offset (reference homography.py:309-321)."""
    w = golden.file("warps")
    H = w[f"pano_{name}_H"]
    assert np.array_equal(port.stitch_panorama(w["img2"].copy(), w["img"].copy(), H), w[f"pano_{name}_paste"])
    assert np.array_equal(port.stitch_panorama(w["img2"].copy(), w["img"].copy(), H, "Rate", 0.2), w[f"pano_{name}_blend"])


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("cross", [True, False])
def test_port_hamming_matcher_equals_opencv(seed, cross):
    """The numpy restatement of BFMatcher(NORM_HAMMING, crossCheck).match against OpenCV itself (ties included)."""
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(seed)
    A, B = _descriptors(rs, 157 + 40 * seed), _descriptors(rs, 203 - 30 * seed)
    B[::7] = A[rs.randint(0, A.shape[0], B[::7].shape[0])]        # exact duplicates across the sets
    m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=cross).match(A, B)
    q, t, d = port.bf_match_hamming(A, B, cross)
    assert [x.queryIdx for x in m] == q.tolist() and [x.trainIdx for x in m] == t.tolist()
    assert [int(x.distance) for x in m] == d.tolist()
