"""GPU: components either side of the hot path: the cylindrical warps (homography.py:340-416) and the
cv2.remap kernel behind them, against the live reference's outputs (tests/golden/ext.npz) and the oracle."""
import numpy as np
import pytest

from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    """Census of what these tests measured -> gpurun_out/PARITY_r02_ext.json (copied to profiles/ by hand)."""
    yield
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "PARITY_r02_ext.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("ch", [1, 2, 4])
def test_remap_kernel_is_bit_exact(golden, ch):
    from ransac_with_homography_b200 import _cabi
    z = golden.file("ext")
    src, mx, my, dst = z[f"remap_src{ch}"], z[f"remap_mx{ch}"], z[f"remap_my{ch}"], z[f"remap_dst{ch}"]
    out = np.ascontiguousarray(dst.copy())
    ctx = _cabi.context()
    _cabi.check(ctx.lib.b200geo_remap_linear_host(ctx.handle, _cabi.ptr(np.ascontiguousarray(src)), src.shape[0], src.shape[1], ch,
                                                  _cabi.ptr(np.ascontiguousarray(mx)), _cabi.ptr(np.ascontiguousarray(my)), mx.shape[0], mx.shape[1],
                                                  _cabi.ptr(out)))
    assert np.array_equal(out, z[f"remap_ref{ch}"])              # cv2.remap of OpenCV 4.13, NaN / inf / huge coordinates included


def test_remap_kernel_random_against_oracle():
    from ransac_with_homography_b200 import _cabi
    rs = np.random.RandomState(3)
    ctx = _cabi.context()
    for _ in range(6):
        ih, iw, oh, ow = rs.randint(1, 90), rs.randint(1, 120), rs.randint(1, 70), rs.randint(1, 300)
        src = rs.randint(0, 256, (ih, iw, 4)).astype(np.uint8)
        mx = (rs.rand(oh, ow) * (iw + 8) - 4).astype(np.float32)
        my = (rs.rand(oh, ow) * (ih + 8) - 4).astype(np.float32)
        dst = rs.randint(0, 256, (oh, ow, 4)).astype(np.uint8)
        out = dst.copy()
        _cabi.check(ctx.lib.b200geo_remap_linear_host(ctx.handle, _cabi.ptr(src), ih, iw, 4, _cabi.ptr(mx), _cabi.ptr(my), oh, ow, _cabi.ptr(out)))
        assert np.array_equal(out, R.remap_linear_transparent(src, mx, my, dst))


@pytest.mark.parametrize("f", [100, 300, 1600])
def test_cylindericl_map_equals_reference(golden, f):
    from ransac_with_homography_b200 import homography as hg
    z = golden.file("ext")
    img = z["cyl_img"]
    o = hg.cylindericlMap(img.copy(), f=f)
    ref = z[f"cyl_map_f{f}"]
    h, w = img.shape[:2]
    assert o.shape == ref.shape and o.dtype == np.uint8
    assert np.array_equal(o, R.cylindrical_warp(img, np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]])))
    t = o[..., 3] == 255
    assert np.array_equal(o[t], ref[t]) and not o[~t].any()       # the reference leaves the unwritten pixels uninitialised
    REPORT[f"cylindericlMap_f{f}"] = {"written_pixel_share": float(t.mean()), "written_pixels_differing_from_reference": int((o[t] != ref[t]).any(-1).sum())}


def test_cylindrical_warp_general_intrinsics(golden):
    from ransac_with_homography_b200 import homography as hg
    z = golden.file("ext")
    o = hg.cylindricalWarp(z["cyl_img"].copy(), z["cyl_K"])
    t = o[..., 3] == 255
    assert t.mean() > 0.5 and np.array_equal(o[t], z["cyl_warp_K"][t]) and not o[~t].any()


def test_cylinderical_transform(golden):
    from ransac_with_homography_b200 import homography as hg
    z = golden.file("ext")
    img = z["cyl_img"]
    o, mx, my = hg.cylindericalTransform(img.copy(), 100, method="nn")
    assert np.array_equal(o, z["cyl_transform_nn"]) and [mx, my] == z["cyl_transform_nn_off"].tolist()
    o, mx, my = hg.cylindericalTransform(img.copy(), 37.5, xc=0.3, yc=0.8, method="nn")
    assert np.array_equal(o, z["cyl_transform_nn2"]) and [mx, my] == z["cyl_transform_nn2_off"].tolist()
    with pytest.raises(IndexError):                               # the reference's default method dies at homography.py:135
        hg.cylindericalTransform(img.copy(), 100)
    with pytest.raises(ValueError):                               # samples np.empty in the reference
        hg.cylindericalTransform(img.copy(), 100, method="nn", wrap="backward")


# ---- device-side JPEG I/O (b200geo_jpeg_*, nvJPEG): cv2.imread / cv2.imwrite of the reference's own fixtures ----
def _psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return 10 * np.log10(255.0 ** 2 / max(float((d * d).mean()), 1e-12))


@pytest.mark.parametrize("name", ["foto1A.jpg", "cv.jpg"])
def test_jpeg_decode_against_cv2(name):
    """cv2.imread decodes with libjpeg-turbo (integer IDCT, 'fancy' chroma up-sampling); nvJPEG's IDCT and up-sampling
    differ, so the comparison is a census: same shape, and the decoded pixels within a few grey levels of cv2's."""
    import os
    cv2 = pytest.importorskip("cv2")
    import torch
    from ransac_with_homography_b200 import imgio
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, name)
    ref = cv2.imread(path)
    got = imgio.imread(path)
    assert got.shape == ref.shape and got.dtype == np.uint8
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    census = {"max": int(d.max()), "mean": float(d.mean()), "p99": float(np.percentile(d, 99)), "equal": float((d == 0).mean())}
    print("jpeg decode census", name, census)
    REPORT["jpeg_decode_vs_cv2_" + name] = dict(census, psnr_db=float(_psnr(got, ref)))
    assert census["mean"] < 1.5 and census["p99"] <= 12 and _psnr(got, ref) > 35, census
    # device-resident form, RGB order (what ransac.example1 makes of cv2.imread with [:, :, ::-1])
    with open(path, "rb") as f:
        data = f.read()
    t = imgio.decode(data, order="rgb")
    torch.cuda.synchronize()
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), got[:, :, ::-1])
    assert imgio.info(data)[:2] == ref.shape[:2]


def test_jpeg_encode_round_trip():
    import os
    cv2 = pytest.importorskip("cv2")
    import torch
    from ransac_with_homography_b200 import imgio
    from conftest import GOLDEN
    img = cv2.imread(os.path.join(GOLDEN, "foto1B.jpg"))
    data = imgio.imencode(img)                                    # quality 95, 4:2:0: cv2.imwrite's defaults
    assert data[:2] == b"\xff\xd8" and data[-2:] == b"\xff\xd9"
    back = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)        # an independent decoder reads our file
    ok, ref_bytes = cv2.imencode(".jpg", img)
    ref_back = cv2.imdecode(ref_bytes, cv2.IMREAD_COLOR)
    assert back.shape == img.shape
    p_ours, p_cv = _psnr(back, img), _psnr(ref_back, img)
    print("jpeg encode: psnr ours", p_ours, "cv2", p_cv, "bytes ours", len(data), "cv2", len(ref_bytes))
    REPORT["jpeg_encode_q95_foto1B"] = {"psnr_db": float(p_ours), "cv2_psnr_db": float(p_cv), "bytes": len(data), "cv2_bytes": len(ref_bytes)}
    # nvJPEG's quantisation / colour conversion give 43.7 dB where libjpeg-turbo gives 49.0 dB at the same size on this image (B200, r02)
    assert p_ours > 40.0 and p_ours > p_cv - 7.0 and 0.5 < len(data) / len(ref_bytes) < 2.0
    t = torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1])).cuda()
    data2 = imgio.encode(t, order="rgb")
    assert data2 == data                                          # same pixels, device-resident source
    with pytest.raises(ValueError):
        imgio.imwrite("/tmp/x.png", img)


def test_jpeg_gray_file_and_bad_input():
    """A single-component JPEG decodes to three equal channels (cv2.IMREAD_COLOR); garbage is an error, not a crash."""
    cv2 = pytest.importorskip("cv2")
    from ransac_with_homography_b200 import _cabi, imgio
    rs = np.random.RandomState(1)
    gray = cv2.GaussianBlur(rs.randint(0, 256, (75, 101)).astype(np.uint8), (0, 0), 2.0)
    ok, data = cv2.imencode(".jpg", gray)
    ref = cv2.imdecode(data, cv2.IMREAD_COLOR)
    got = imgio.imdecode(data.tobytes())
    assert got.shape == ref.shape == (75, 101, 3)
    assert np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 1], got[..., 2])
    assert np.abs(got.astype(np.int16) - ref.astype(np.int16)).max() <= 2
    assert imgio.info(data.tobytes()) == (75, 101, 1)
    with pytest.raises(_cabi.B200GeoError):
        imgio.imdecode(b"not a jpeg at all" * 10)
    assert imgio.imread("/nonexistent/file.jpg") is None
