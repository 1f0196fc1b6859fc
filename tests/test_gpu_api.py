"""GPU: the remaining names of the reference's API surface against outputs of the LIVE reference (tests/golden/api.npz):
calcHomography(collective=True) (S6), HomoModel.fit / fwd / reproj (R2-R4), wrapPerspective(boundary=1) and
wrapPerspective(crop=False) (W3)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def test_calc_homography_collective_is_the_svd_minimiser(golden):
    """S6 (homography.py:30-46, 71-88): the last right singular vector of the 2M x 9 matrix, scaled to h22 = 1."""
    from ransac_with_homography_b200 import homography as hg
    z = golden.file("api")
    A, B = golden.points()
    inl = z["inl"]
    H = hg.calcHomography(A[inl], B[inl], True)
    assert H.shape == (3, 3) and H.dtype == np.float32 and H[2, 2] == 1.0
    ref = z["H_collective_svd"]                                   # the reference's float32 SVD
    assert np.linalg.norm(H - ref) / np.linalg.norm(ref) < 1e-4   # north_star: homography within 1e-4 relative
    # float64 truth on the reference's own float32 matrix
    M = z["corresp_collective"].astype(np.float64)
    vt = np.linalg.svd(M)[2][-1].reshape(3, 3)
    t = vt / vt[2, 2]
    assert np.linalg.norm(H - t) / np.linalg.norm(t) < 2e-7       # float32 rounding of the result only
    # and it is NOT the inhomogeneous least-squares solution calcHomographyLinear(collective=True) returns
    L = hg.calcHomographyLinear(A[inl], B[inl], True)
    assert np.linalg.norm(L - t) / np.linalg.norm(t) > 1e-3
    with pytest.raises(np.linalg.LinAlgError):
        hg.calcHomography(np.zeros((3, 2), np.float32), np.ones((3, 2), np.float32), True)


def test_homomodel_fit_fwd_reproj(golden):
    from ransac_with_homography_b200.ransac import HomoModel
    z = golden.file("api")
    A, B = golden.points()
    inl = z["inl"]
    m = HomoModel(th=5, d=70, n=4)
    H = m.fit(z["u4"].T, z["v4"].T)
    assert H is m.val and H.dtype == np.float32 and H.shape == (3, 3)
    assert np.linalg.norm(H - z["fit4"]) / np.linalg.norm(z["fit4"]) < 1e-4
    # fwd / reproj of the reference's own model: bit-identical (OpenBLAS sgemm order + IEEE division)
    m.val = z["fit4"]
    f = m.fwd(A.T)
    assert f.dtype == np.float32 and np.array_equal(f.view(np.uint32), z["fit4_fwd"].view(np.uint32))
    r = m.reproj(B.T)
    assert r.dtype == np.float32 and np.array_equal(r.view(np.uint32), z["fit4_reproj"].view(np.uint32))
    # collective fit: float64 least squares; the reference's float32 normal equations carry 2-3 digits
    Hc = m.fit(A[inl].T, B[inl].T, collective=True)
    assert Hc.dtype == np.float64 and m.val is Hc
    T = oracle.refit_f64(oracle.pack_points(A.T, B.T), np.isin(np.arange(185), inl))
    assert np.linalg.norm(Hc - T) / np.linalg.norm(T) < 1e-4
    assert np.linalg.norm(Hc - z["fit_collective"]) / np.linalg.norm(z["fit_collective"]) < 2e-2
    with pytest.raises(AssertionError):
        m.fit(A[:5].T, B[:5].T)                     # 5 columns without collective (ransac.py:38)
    with pytest.raises(AssertionError):
        m.fit(A[:4], B[:4])                         # wrong orientation: 4 rows (ransac.py:39)


def test_wrap_perspective_boundary_and_nocrop(golden):
    """wrapPerspective(boundary=1): min_x / min_y clamped to >= 0 (homography.py:150-156); crop=False: the warped image
    placed into a float32 canvas (homography.py:180-183, only valid for non-negative offsets)."""
    from ransac_with_homography_b200 import homography as hg
    z = golden.file("api")
    w = golden.file("warps")
    img = w["img"]
    for name, H in (("stitch", w["H_stitch"]), ("shrink", w["H_shrink"]), ("neg", z["H_neg"])):
        for conv in (("nn", "bilinear") if name != "neg" else ("bilinear",)):
            out, mx, my = hg.wrapPerspective(img.copy(), H, conv, boundary=1)
            assert [mx, my] == z[f"b1_{name}_{conv}_off"].tolist() and min(mx, my) >= 0
            assert np.array_equal(out.astype(np.uint8), z[f"b1_{name}_{conv}"])
    out, mx, my = hg.wrapPerspective(img.copy(), z["H_nocrop"], "nn", crop=False)
    assert [mx, my] == z["nocrop_nn_off"].tolist()
    assert out.dtype == z["nocrop_nn"].dtype and np.array_equal(out, z["nocrop_nn"])
