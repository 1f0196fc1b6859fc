"""GPU parity tests for the Hamming matcher (K6) through the C ABI and the drop-in helper of stitching().
Integer work: indices and distances must equal the oracle's (itself pinned to OpenCV) exactly."""
import numpy as np
import pytest

from oracle import ref_numpy as port
from conftest import descriptors as _descriptors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nA,nB,nbytes", [(500, 437, 32), (129, 1, 32), (1, 300, 32), (128, 128, 32), (257, 511, 64), (77, 90, 4)])
@pytest.mark.parametrize("cross", [True, False])
def test_match_hamming_equals_oracle(nA, nB, nbytes, cross):
    from ransac_with_homography_b200 import ransac as rs_mod
    rs = np.random.RandomState(nA * 1000 + nB + nbytes)
    A, B = _descriptors(rs, nA, nbytes), _descriptors(rs, nB, nbytes)
    B[::5] = A[rs.randint(0, nA, B[::5].shape[0])]                # exact duplicates across the sets: ties
    m = rs_mod.match_hamming(A, B, crossCheck=cross)
    q, t, d = port.bf_match_hamming(A, B, cross)
    assert [x.queryIdx for x in m] == q.tolist()
    assert [x.trainIdx for x in m] == t.tolist()
    assert [int(x.distance) for x in m] == d.tolist()
    if nbytes == 32:
        cv2 = pytest.importorskip("cv2")
        ref = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=cross).match(A, B)
        assert [(x.queryIdx, x.trainIdx, int(x.distance)) for x in ref] == [(x.queryIdx, x.trainIdx, int(x.distance)) for x in m]


def test_match_hamming_empty_inputs():
    from ransac_with_homography_b200 import ransac as rs_mod
    A = np.zeros((5, 32), np.uint8)
    assert rs_mod.match_hamming(None, A) == [] and rs_mod.match_hamming(A, None) == []
    assert rs_mod.match_hamming(np.zeros((0, 32), np.uint8), A) == [] and rs_mod.match_hamming(A, np.zeros((0, 32), np.uint8)) == []
    m = rs_mod.match_hamming(A, A)                                # all equal: every query's first minimum is train 0; only 0 is mutual
    assert [(x.queryIdx, x.trainIdx, x.distance) for x in m] == [(0, 0, 0.0)]


def test_stitching_end_to_end_on_synthetic_pair():
    """stitching() (reference ransac.py:235-283) with ORB on OpenCV and matcher / RANSAC / compose on the GPU: a
    textured image and a translated crop of it must be registered to the known shift."""
    cv2 = pytest.importorskip("cv2")
    from ransac_with_homography_b200 import ransac as rs_mod
    rs = np.random.RandomState(3)
    base = np.zeros((420, 640, 3), np.uint8)
    for _ in range(400):                                          # corners for ORB: random filled rectangles
        x, y = int(rs.randint(0, 600)), int(rs.randint(0, 380))
        cv2.rectangle(base, (x, y), (x + int(rs.randint(8, 40)), y + int(rs.randint(8, 40))), [int(v) for v in rs.randint(30, 255, 3)], -1)
    train, query = np.ascontiguousarray(base[:, 120:520]), np.ascontiguousarray(base[:, 40:440])   # train is 80 px to the right of query
    np.random.seed(0)
    out = rs_mod.stitching(train.copy(), query.copy(), ransacMet="fwd", th=3, d=60, k=500)
    assert out.dtype == np.uint8 and out.ndim == 3 and out.shape[2] == 3
    assert abs(out.shape[1] - 480) <= 8 and abs(out.shape[0] - 420) <= 8          # 400 + 80 columns, up to the estimate's error
    # paste: the query image sits on top, unmodified, at the placement offset (reference homography.py:337)
    found = [(dy, dx) for dy in range(0, 9) for dx in range(0, 9)
             if out.shape[0] >= dy + 420 and out.shape[1] >= dx + 400 and np.array_equal(out[dy:dy + 420, dx:dx + 400], query)]
    assert len(found) >= 1
    # and the part only the warped train image covers reproduces the scene (columns 400.. of the union crop)
    dy, dx = found[0]
    ref = base[:, 440:520]
    got = out[dy:dy + 420, dx + 400:dx + 480]
    w = min(got.shape[1], ref.shape[1]) - 6
    diff = np.abs(got[8:412, :w].astype(np.int16) - ref[8:412, :w].astype(np.int16))
    assert np.mean(diff > 40) < 0.1
