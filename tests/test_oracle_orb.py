"""CPU: the oracle's restatement of cv2.ORB_create().detectAndCompute (oracle/orb_numpy.py; the reference calls it at
ransac.py:255-257) against OpenCV 4.13's recorded outputs (tests/golden/orb.npz, tests/golden/make_golden_ext.py) and,
where cv2 is importable, against OpenCV live, stage by stage."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import orb_numpy as O

PATTERN = np.load(os.path.join(GOLDEN, "orb_pattern.npy")).astype(np.int64)


def _rgb(golden, name):
    z = golden.file("orb")
    if name + "_rgb" in z.files:
        return z[name + "_rgb"]
    cv2 = pytest.importorskip("cv2")
    return np.ascontiguousarray(cv2.imread(os.path.join(GOLDEN, name + ".jpg"))[:, :, ::-1])


@pytest.mark.parametrize("name,nf", [("noise", 500), ("noise", 60), ("small", 500), ("foto1A", 500), ("notebook", 500)])
def test_orb_restatement_equals_opencv(golden, name, nf):
    """Keypoints (x, y, size, angle, response, octave -- as float32 bit patterns, in OpenCV's order) and descriptors."""
    z = golden.file("orb")
    tag = name if nf == 500 else f"{name}_n{nf}"
    kp, ds = O.detect_and_compute(O.rgb2gray(_rgb(golden, name)), PATTERN, nf)
    assert kp.shape == z[tag + "__kp"].shape
    assert np.array_equal(kp.view(np.uint32), z[tag + "__kp"].view(np.uint32))
    assert np.array_equal(ds, z[tag + "__desc"])


def test_orb_stages_against_live_opencv():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(4)
    rgb = rs.randint(0, 256, (97, 131, 3)).astype(np.uint8)
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
    assert np.array_equal(O.rgb2gray(rgb), gray)
    # INTER_LINEAR_EXACT at the pyramid's ratios
    prev = gray
    for (h, w) in O.level_sizes(*gray.shape)[1:4]:
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(O.resize_linear_exact(prev, h, w), ref)
        prev = ref
    # FAST-9/16 with non-maximum suppression: positions, order and scores
    smooth = cv2.GaussianBlur(rs.randint(0, 256, (150, 200)).astype(np.uint8), (5, 5), 1.2)
    ref = np.array([[k.pt[0], k.pt[1], k.response] for k in cv2.FastFeatureDetector_create(20, True).detect(smooth, None)]).astype(np.int64).reshape(-1, 3)
    assert len(ref) > 20 and np.array_equal(O.fast_detect(smooth), ref)
    # the descriptor's blur: sepFilter2D with the float32 Gaussian (widths with and without a scalar tail)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    assert np.array_equal(k.ravel(), O.gauss_kernel())
    for w in (96, 131, 200):
        img = rs.randint(0, 256, (60, w)).astype(np.uint8)
        ref = cv2.sepFilter2D(img, -1, k, k, borderType=cv2.BORDER_REFLECT_101)
        got = O.blur_level(np.pad(img, O.BORDER, mode="reflect"))[O.BORDER:-O.BORDER, O.BORDER:-O.BORDER]
        assert np.array_equal(got, ref)
    # cv::fastAtan2
    for y, x in rs.randint(-5000, 5000, (200, 2)):
        assert O.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_retain_best_matches_libstdcxx_order():
    """retainBest keeps ties with the n-th response and leaves std::nth_element's order: a fixed small case worked by hand
    from libstdc++'s introselect (median-of-three to the front, unguarded partition, insertion sort of the last <= 3)."""
    pts = [(i, v) for i, v in enumerate([5, 9, 1, 9, 7, 3, 9, 2, 8, 6])]
    out = O.retain_best(pts, 3, key=lambda r: r[1])
    assert sorted(v for _, v in out) == [9, 9, 9]
    out = O.retain_best(pts, 4, key=lambda r: r[1])
    assert sorted(v for _, v in out) == [8, 9, 9, 9]
    out = O.retain_best([(i, 4) for i in range(7)], 2, key=lambda r: r[1])
    assert len(out) == 7                                           # every point ties with the boundary response
    assert O.retain_best(pts, 0, key=lambda r: r[1]) == [] and O.retain_best(pts, 50, key=lambda r: r[1]) == pts


def test_product_pattern_table_is_the_extracted_one():
    inc = open(os.path.join(os.path.dirname(GOLDEN), "..", "ransac_with_homography_b200", "csrc", "orb_pattern.inc")).read()
    import re
    vals = np.array([int(v) for v in re.findall(r"-?\d+", inc.split("\n", 1)[1])]).reshape(512, 2)
    assert np.array_equal(vals, PATTERN)
