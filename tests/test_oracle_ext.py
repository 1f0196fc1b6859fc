"""CPU: the oracle's restatements of the components either side of the hot path against
tests/golden/ext.npz -- outputs of the live reference (cylindrical warps, homography.py:340-416) and of OpenCV 4.13
itself (cv2.remap) recorded by tests/golden/make_golden_ext.py."""
import numpy as np
import pytest

from oracle import ref_numpy as R


@pytest.mark.parametrize("ch", [1, 2, 4])
def test_remap_restatement_equals_cv2_remap(golden, ch):
    z = golden.file("ext")
    with np.errstate(invalid="ignore"):
        o = R.remap_linear_transparent(z[f"remap_src{ch}"], z[f"remap_mx{ch}"], z[f"remap_my{ch}"], z[f"remap_dst{ch}"])
    ref = z[f"remap_ref{ch}"]
    assert np.array_equal(o, ref)
    assert 0.2 < float((ref != z[f"remap_dst{ch}"]).any(-1).mean()) < 0.9          # both branches (written / transparent) are exercised


def test_remap_restatement_equals_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(11)
    src = rs.randint(0, 256, (64, 80, 4)).astype(np.uint8)
    mx = (rs.rand(50, 70) * 90 - 5).astype(np.float32)
    my = (rs.rand(50, 70) * 74 - 5).astype(np.float32)
    dst = rs.randint(0, 256, (50, 70, 4)).astype(np.uint8)
    ref = cv2.remap(src, mx, my, cv2.INTER_AREA, dst=dst.copy(), borderMode=cv2.BORDER_TRANSPARENT)
    assert np.array_equal(R.remap_linear_transparent(src, mx, my, dst), ref)


@pytest.mark.parametrize("f", [100, 300, 1600])
def test_cylindrical_warp_restatement_equals_reference(golden, f):
    """cylindericlMap (homography.py:412-416).  The reference's output holds arbitrary bytes where cv2.remap wrote nothing
    (a fresh, uninitialised array); the comparison is on the written pixels, the rest must be transparent zeros here."""
    z = golden.file("ext")
    img = z["cyl_img"]
    h, w = img.shape[:2]
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]])
    o = R.cylindrical_warp(img, K)
    ref = z[f"cyl_map_f{f}"]
    t = o[..., 3] == 255
    assert o.shape == ref.shape and t.mean() > 0.7
    assert np.array_equal(o[t], ref[t]) and not o[~t].any()
