"""GPU: K7, the ORB front-end of stitching() (reference ransac.py:252-257), through the C ABI (b200geo_orb_host) against
OpenCV 4.13's recorded outputs (tests/golden/orb.npz) and the CPU oracle (oracle/orb_numpy.py): keypoints as float32 bit
patterns in OpenCV's order, descriptors byte for byte."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import orb_numpy as O

pytestmark = pytest.mark.gpu
REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    """Census of what these tests measured -> gpurun_out/PARITY_r02_orb.json (copied to profiles/ by hand)."""
    yield
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "PARITY_r02_orb.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)
PATTERN = np.load(os.path.join(GOLDEN, "orb_pattern.npy")).astype(np.int64)


def _rgb(golden, name):
    z = golden.file("orb")
    if name + "_rgb" in z.files:
        return z[name + "_rgb"]
    from ransac_with_homography_b200 import imgio
    return imgio.imread(os.path.join(GOLDEN, name + ".jpg"), order="rgb")


def _jpeg_rgb(name):
    cv2 = pytest.importorskip("cv2")          # decoder of record for the fixtures' pixels (nvJPEG's differ by a few grey levels)
    return np.ascontiguousarray(cv2.imread(os.path.join(GOLDEN, name + ".jpg"))[:, :, ::-1])


@pytest.mark.parametrize("name,nf", [("noise", 500), ("noise", 60), ("small", 500)])
def test_orb_equals_opencv_on_recorded_images(golden, name, nf):
    from ransac_with_homography_b200.ransac import orb_detect_and_compute
    z = golden.file("orb")
    tag = name if nf == 500 else f"{name}_n{nf}"
    rgb = z[name + "_rgb"]
    kp, ds = orb_detect_and_compute(rgb, nf, as_arrays=True)
    assert kp.shape == z[tag + "__kp"].shape
    assert np.array_equal(kp.view(np.uint32), z[tag + "__kp"].view(np.uint32))
    assert np.array_equal(ds, z[tag + "__desc"])
    kp2, ds2 = orb_detect_and_compute(O.rgb2gray(rgb), nf, as_arrays=True)          # grey input: same result
    assert np.array_equal(kp2, kp) and np.array_equal(ds2, ds)


@pytest.mark.parametrize("name,nf", [("foto1A", 500), ("foto1A", 60), ("foto1B", 500), ("cv", 500), ("notebook", 500)])
def test_orb_equals_opencv_on_the_reference_images(golden, name, nf):
    from ransac_with_homography_b200.ransac import orb_detect_and_compute
    z = golden.file("orb")
    tag = name if nf == 500 else f"{name}_n{nf}"
    kps, ds = orb_detect_and_compute(_jpeg_rgb(name), nf)
    ref = z[tag + "__kp"]
    got = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32)
    REPORT[tag] = {"keypoints": int(len(got)), "opencv_keypoints": int(len(ref)),
                   "keypoint_fields_differing": int((got.view(np.uint32) != ref.view(np.uint32)).sum()) if got.shape == ref.shape else -1,
                   "descriptor_bits_differing": int(np.unpackbits(ds ^ z[tag + "__desc"]).sum()) if ds.shape == z[tag + "__desc"].shape else -1}
    assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(ds, z[tag + "__desc"])
    assert kps[0].class_id == -1 and isinstance(kps[0].pt, tuple)


def test_orb_random_images_against_oracle():
    from ransac_with_homography_b200.ransac import orb_detect_and_compute
    rs = np.random.RandomState(21)
    for (h, w) in ((140, 190), (97, 300), (64, 64), (70, 75), (300, 101)):
        base = rs.randint(0, 256, (h // 4 + 2, w // 4 + 2)).astype(np.float64)
        img = np.kron(base, np.ones((4, 4)))[:h, :w] + rs.randn(h, w) * 6                 # blocky corners + noise
        gray = np.clip(img, 0, 255).astype(np.uint8)
        kp, ds = orb_detect_and_compute(gray, 120, as_arrays=True)
        okp, ods = O.detect_and_compute(gray, PATTERN, 120)
        assert kp.shape == okp.shape, (h, w)
        if len(kp):
            assert np.array_equal(kp.view(np.uint32), okp.view(np.uint32)) and np.array_equal(ds, ods)
    kps, ds = orb_detect_and_compute(np.full((90, 120), 7, np.uint8))
    assert len(kps) == 0 and ds is None                                                     # as OpenCV: no keypoints, no descriptors


def test_stitching_front_end_feeds_ransac_the_reference_points(golden):
    """cvtColor + ORB + BFMatcher(crossCheck) + sort of stitching() (ransac.py:252-267) on the reference's own pair."""
    from ransac_with_homography_b200.ransac import match_hamming, orb_detect_and_compute
    z = golden.file("orb")
    ka, fa = orb_detect_and_compute(_jpeg_rgb("foto1A"))
    kb, fb = orb_detect_and_compute(_jpeg_rgb("foto1B"))
    m = sorted(match_hamming(fa, fb, crossCheck=True), key=lambda x: x.distance)
    pa, pb = np.float32([k.pt for k in ka]), np.float32([k.pt for k in kb])
    assert np.array_equal(np.float32([pa[x.queryIdx] for x in m]), z["pair_ptsA"])
    assert np.array_equal(np.float32([pb[x.trainIdx] for x in m]), z["pair_ptsB"])
    REPORT["stitching_front_end_foto1A_foto1B"] = {"matches": len(m), "point_lists_equal_to_opencv_pipeline": True}


def test_stitching_end_to_end_on_the_reference_pair():
    """ransac.example1 (ransac.py:286-301) without OpenCV in the loop: the panorama of foto1A / foto1B."""
    from ransac_with_homography_b200.ransac import stitching
    A, B = _jpeg_rgb("foto1A"), _jpeg_rgb("foto1B")
    np.random.seed(0)
    with redirect_stdout(io.StringIO()):
        out = stitching(A, B, "fwd", blending="Rate", blendrate=0.2)
    assert out.dtype == np.uint8 and out.ndim == 3 and out.shape[2] == 3
    assert out.shape[0] >= B.shape[0] and out.shape[1] > B.shape[1] and out.shape[1] < 3 * B.shape[1]
    assert (out.reshape(-1, 3).any(1)).mean() > 0.5


def test_orb_short_fuzz_against_live_opencv():
    """tools/fuzz_orb.py (random sizes, contents, feature budgets, grey / RGB input against OpenCV on this machine): a short run."""
    pytest.importorskip("cv2")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_orb.py"), "40", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    REPORT["fuzz_orb_short"] = r.stdout.strip().splitlines()[-1]


def test_orb_edge_cases():
    """Sizes with empty pyramid levels, tiny feature budgets, a constant image, and the capacity of the output buffers."""
    cv2 = pytest.importorskip("cv2")
    from ransac_with_homography_b200.ransac import orb_detect_and_compute
    rs = np.random.RandomState(8)
    for (h, w, nf) in ((30, 30, 500), (63, 200, 500), (200, 63, 500), (64, 65, 500), (120, 160, 1), (120, 160, 2), (120, 160, 0), (97, 1300, 37)):
        img = cv2.GaussianBlur(rs.randint(0, 256, (h, w)).astype(np.uint8), (0, 0), 1.0)
        kps, des = cv2.ORB_create(nf).detectAndCompute(img, None)
        ref = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)
        kp, ds = orb_detect_and_compute(img, nf, as_arrays=True)
        assert kp.shape == ref.shape and np.array_equal(kp.view(np.uint32), ref.view(np.uint32)), (h, w, nf)
        if len(ref):
            assert np.array_equal(ds, des), (h, w, nf)
        else:
            assert ds is None
    with pytest.raises(AssertionError):
        orb_detect_and_compute(np.zeros((40, 40, 4), np.uint8))
    with pytest.raises(AssertionError):
        orb_detect_and_compute(np.zeros((40, 40), np.float32))
