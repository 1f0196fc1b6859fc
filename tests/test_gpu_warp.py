"""GPU parity tests for the backward warp (K4) through the C ABI and the drop-in Python API.
Golden vectors = outputs of the live reference; tolerance per BASELINE.json north_star:
warped pixels within +-1 on uint8 (the float64 kernel is expected to be exact)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

WARP_H = ["stitch", "rot", "shrink", "shift"]
U8_TOL = 1     # north_star: "warped pixels match within +-1 on uint8"
# uint8 RGB -> uint8 RGB warps run the fast kernel (warp_fast.cuh) unless the context is switched to
# the float64 kernel.  "exact": bit-identical to the reference (census < 1e-4 allows for the
# reference's own last-ulp noise); "fast": nn still bit-identical, bilinear within +-1 with a
# bounded share of off-by-one pixels (float32 tile-relative coordinates, 16-bit blend weights).
MODES = ["fast", "exact"]
CENSUS = {"fast": 0.02, "exact": 1e-4}


@pytest.fixture(params=MODES)
def mode(request):
    from ransac_with_homography_b200 import _cabi
    ctx = _cabi.context()
    ctx.warp_exact(request.param == "exact")
    yield request.param
    ctx.warp_exact(False)


def _maxdiff(a, b):
    return int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) if a.size else 0


@pytest.mark.parametrize("name", WARP_H)
@pytest.mark.parametrize("conv", ["nn", "bilinear"])
def test_wrap_perspective_matches_reference(golden, name, conv):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    img = w["img"].copy()
    out, mx, my = hg.wrapPerspective(img, w[f"H_{name}"], conv)
    assert [mx, my] == w[f"bbox_{name}_{conv}_off"].tolist()
    assert (img[0, 0] == 0).all()                         # side effect of the reference's samplers
    assert out.dtype == (np.uint8 if conv == "nn" else np.float64)
    got = out.astype(np.uint8)
    ref = w[f"bbox_{name}_{conv}"]
    assert got.shape == ref.shape
    assert _maxdiff(got, ref) <= (0 if conv == "nn" else U8_TOL)
    assert np.mean(got != ref) < 1e-4
    if conv == "bilinear":
        o64, _, _ = oracle.wrapPerspective(w["img"], w[f"H_{name}"], conv)
        np.testing.assert_allclose(out, o64, rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", WARP_H)
def test_transform_image_h_matches_reference(golden, name, mode):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    out, mx, my = hg.transformImageH(w["img"].copy(), w[f"H_{name}"])
    assert out.dtype == np.uint8 and [mx, my] == w[f"bbox_{name}_bilinear_off"].tolist()
    assert _maxdiff(out, w[f"tih_{name}"]) <= U8_TOL and np.mean(out != w[f"tih_{name}"]) < CENSUS[mode]


@pytest.mark.parametrize("name", WARP_H)
def test_nn_u8_fast_kernel_is_bit_exact(golden, name):
    """uint8 nn goes through the fast kernel by default and must equal the reference everywhere."""
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    out, mx, my = hg.wrapPerspective(w["img"].copy(), w[f"H_{name}"], "nn", _out="u8")
    assert np.array_equal(out, w[f"bbox_{name}_nn"])


def test_rgba_float_warp_matches_reference(golden):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    rgba = np.zeros(w["img"].shape[:2] + (4,), np.float32)
    rgba[:, :, :3] = w["img"]
    rgba[:, :, 3] = 0.2 + 1e-10
    out, mx, my = hg.transformImageH(rgba, w["H_stitch"])
    assert out.dtype == np.float64 and out.shape == w["bbox_stitch_rgba"].shape
    np.testing.assert_allclose(out, w["bbox_stitch_rgba"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("tag", ["s1", "s2"])
@pytest.mark.parametrize("conv", ["nn", "bilinear"])
def test_scanner_matches_reference(golden, tag, conv, mode):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    u, v = w[f"scan_{tag}_u"], w[f"scan_{tag}_v"]
    res = [int(v[1, 2]), int(v[0, 2])]
    # same H to both sides: the Scan-form warp
    out, mx, my = hg.wrapPerspectiveScan(w["img"].copy(), w[f"scan_{tag}_H"], res, conv)
    assert (mx, my) == (0, 0)
    got = out.astype(np.uint8)
    assert _maxdiff(got, w[f"scan_{tag}_{conv}"]) <= (0 if conv == "nn" else U8_TOL)
    assert np.mean(got != w[f"scan_{tag}_{conv}"]) < 1e-4
    # whole Scanner call (our own float64 4-point solve; the reference's is float32 normal equations,
    # ~8e-3 px of source displacement): +-1 except at nn rounding boundaries / strong edges
    # fused uint8 output with the same H: the kernel under test for this mode
    out8, _, _ = hg.wrapPerspectiveScan(w["img"].copy(), w[f"scan_{tag}_H"], res, conv, _out="u8")
    assert _maxdiff(out8, w[f"scan_{tag}_{conv}"]) <= (0 if conv == "nn" else U8_TOL)
    assert np.mean(out8 != w[f"scan_{tag}_{conv}"]) < (1e-9 if conv == "nn" else CENSUS[mode])
    ti = hg.transformImage(w["img"].copy(), u, v, box=res, method=conv)
    ref = w[f"scan_{tag}_{conv}_ti"]
    assert ti.dtype == np.uint8 and ti.shape == ref.shape
    frac = np.mean(np.abs(ti.astype(np.int16) - ref.astype(np.int16)) > U8_TOL)
    assert frac < (0.02 if conv == "nn" else 0.002)


def test_transform_image_without_box(golden):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    ti = hg.transformImage(w["img"].copy(), w["scan_s1_u"], w["scan_s1_v"])
    ref = w["scan_s1_ti_nobox"]
    assert ti.shape == ref.shape and ti.dtype == np.uint8
    assert np.mean(np.abs(ti.astype(np.int16) - ref.astype(np.int16)) > U8_TOL) < 0.002


@pytest.mark.parametrize("name", ["pp", "nn_", "np_", "pn_"])
def test_stitch_panorama_matches_reference(golden, name, capsys, mode):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    H = w[f"pano_{name}_H"]
    imgT = w["img"].copy()
    paste = hg.stitchPanorama(w["img2"].copy(), imgT, H)
    assert paste.dtype == np.uint8 and paste.shape == w[f"pano_{name}_paste"].shape
    assert (imgT[0, 0] == 0).all()                          # the reference's sampler zeroes the caller's pixel (homography.py:112-116)
    # the fused compose kernel evaluates the warp in float64 like the reference: identical canvases
    assert np.array_equal(paste, w[f"pano_{name}_paste"])
    imgT = w["img"].copy()
    blend = hg.stitchPanorama(w["img2"].copy(), imgT, H, blending="Rate", blendrate=0.2)
    assert blend.shape == w[f"pano_{name}_blend"].shape and blend.dtype == np.uint8
    assert np.array_equal(imgT, w["img"])                   # with blending the zeroed pixel is addAlpha's copy
    assert np.array_equal(blend, w[f"pano_{name}_blend"])
    assert "0.2" in capsys.readouterr().out                 # addAlpha prints the rate (homography.py:257)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_stitch_compose_random_shapes_vs_port(seed, capsys, mode):
    """K5 against the numpy restatement (pinned to the live reference by tests/test_oracle_golden.py) on random
    image sizes, homographies in all offset sign cases, paste and 'Rate' blend with several rates."""
    from oracle import ref_numpy as port
    from ransac_with_homography_b200 import homography as hg
    rs = np.random.RandomState(100 + seed)
    h, w = int(rs.randint(37, 200)), int(rs.randint(41, 260))
    hq, wq = int(rs.randint(30, 220)), int(rs.randint(30, 280))
    imgT = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    imgQ = rs.randint(0, 256, (hq, wq, 3)).astype(np.uint8)
    a = rs.uniform(-0.3, 0.3)
    H = np.array([[np.cos(a) * rs.uniform(0.7, 1.3), -np.sin(a), rs.uniform(-60, 60)],
                  [np.sin(a), np.cos(a) * rs.uniform(0.7, 1.3), rs.uniform(-60, 60)],
                  [rs.uniform(-4e-4, 4e-4), rs.uniform(-4e-4, 4e-4), 1.0]])
    for blending, rate in [(False, 0.2), ("Rate", 0.2), ("Rate", 0.65), ("Rate", 0.0)]:
        ref = port.stitch_panorama(imgQ.copy(), imgT.copy(), H, blending, rate)
        out = hg.stitchPanorama(imgQ.copy(), imgT.copy(), H, blending=blending, blendrate=rate)
        assert out.shape == ref.shape and out.dtype == np.uint8
        assert np.array_equal(out, ref), (blending, rate, int((out != ref).sum()))
    capsys.readouterr()


def test_samplers_on_explicit_coordinates(golden):
    from ransac_with_homography_b200 import homography as hg
    w = golden.file("warps")
    img = w["img"]
    h, wd = img.shape[:2]
    rs = np.random.RandomState(2)
    n = 37 * 41
    z = np.vstack([rs.rand(n) * (wd + 20) - 10, rs.rand(n) * (h + 20) - 10, np.ones(n)])
    z[0, :5] = [0.0, wd - 1.0, -0.4, wd - 0.6, -1.49]       # edge rows of SURVEY A.5
    z[1, :5] = [0.0, 3.0, 2.0, 2.0, 7.0]
    for name, fn in (("nn", hg.nearestNeighbor), ("bilinear", hg.bilinear)):
        got = fn(z.copy(), img.copy(), h, wd, 37, 41)
        inv = np.eye(3)
        want = oracle.warp(img, inv, z[0], np.array([0.0]), (h, wd), name)   # placeholder shape, replaced below
        # oracle over explicit coordinates: one row per point via a 1 x n grid is not expressible, so
        # evaluate point by point through a degenerate grid
        want = np.stack([oracle.warp(img, inv, z[0, i:i + 1], z[1, i:i + 1], (h, wd), name)[0, 0] for i in range(n)])
        want = want.reshape(37, 41, 3)
        if name == "nn":
            assert got.dtype == np.uint8 and np.array_equal(got, want.astype(np.uint8))
        else:
            assert got.dtype == np.float64
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


def test_edge_semantics():
    """A.5/A.6 quirks: pixel (0,0) reads 0; nn truncation x in (-1.5,-0.5] -> column 0 only from -0.5;
    taps beyond the image read 0 where the reference would raise; Scan bounds are `res`."""
    from ransac_with_homography_b200 import homography as hg
    img = (np.arange(6 * 8 * 3).reshape(6, 8, 3) % 251 + 3).astype(np.uint8)
    T = np.array([[1.0, 0, 0], [0, 1, 0], [0, 0, 1]])
    out, mx, my = hg.wrapPerspective(img.copy(), T, "nn")
    want = img.copy()
    want[0, 0] = 0
    assert (mx, my) == (0, 0) and np.array_equal(out, want)
    # identity + bilinear: x == w-1 exactly -> the reference raises IndexError; here the zero-weight tap reads 0
    outb, _, _ = hg.wrapPerspective(img.copy(), T, "bilinear")
    assert np.array_equal(outb.astype(np.uint8), want)
    o, _, _ = oracle.wrapPerspective(img, T, "bilinear")
    assert np.array_equal(outb, o)
    # Scan form with res larger than the image: coordinates between image and res read 0
    S = np.array([[1.0, 0, 0], [0, 1, 0], [0, 0, 1]])
    outs, _, _ = hg.wrapPerspectiveScan(img.copy(), S, [9, 12], "bilinear")
    os_, _, _ = oracle.wrapPerspectiveScan(img, S, [9, 12], "bilinear")
    assert outs.shape == (9, 12, 3) and np.array_equal(outs, os_)
    # singular-ish map: w crosses zero inside the grid -> inf / NaN coordinates are masked
    P = np.array([[1.0, 0, 0], [0, 1, 0], [-0.3, 0, 1]])
    outp, _, _ = hg.wrapPerspectiveScan(img.copy(), P, [6, 8], "nn")
    op, _, _ = oracle.wrapPerspectiveScan(img, P, [6, 8], "nn")
    assert np.array_equal(outp, op.astype(np.uint8))


def test_batched_device_warp_and_round_trip(mode):
    """Batched, device-resident K4 (config 4 shape at reduced size) vs the oracle, plus a
    size-independent property at full 1080p: an integer nn shift followed by its inverse
    reproduces the interior of the frame exactly."""
    import torch
    from ransac_with_homography_b200 import device
    rs = np.random.RandomState(4)
    B, h, w = 5, 90, 160
    frames = rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
    Hs = []
    for b in range(B):
        src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
        dst = src + rs.uniform(-0.05, 0.05, (4, 2)) * [w, h]
        pts = np.concatenate([src, dst], 1).astype(np.float32)
        Hs.append(oracle.dlt4_f64(pts, np.array([[0, 1, 2, 3]], np.int32))[0][0])
    Hs = np.stack(Hs)
    src_t = torch.from_numpy(frames).cuda()
    for sampler in ("nn", "bilinear"):
        jobs = device.scan_jobs(Hs, (h, w), (h, w))
        dst_t = torch.zeros((B, h, w, 3), dtype=torch.uint8, device="cuda")
        device.warp(src_t, dst_t, jobs, 3, sampler)
        got = dst_t.cpu().numpy()
        for b in range(B):
            want = oracle.to_u8(oracle.wrapPerspectiveScan(frames[b], Hs[b], [h, w], sampler)[0])
            assert _maxdiff(got[b], want) <= (0 if sampler == "nn" else U8_TOL)
            assert np.mean(got[b] != want) < (1e-9 if sampler == "nn" else CENSUS[mode])
        # bbox form, packed outputs
        jb, boxes = device.bbox_jobs(Hs, (h, w))
        total = int((jb["dst_pitch"] * jb["oh"]).sum())
        out_t = torch.zeros(total, dtype=torch.uint8, device="cuda")
        device.warp(src_t, out_t, jb, 3, sampler)
        flat = out_t.cpu().numpy()
        for b in range(B):
            o, mx, my = oracle.wrapPerspective(frames[b], Hs[b], sampler)
            assert boxes[b].tolist() == [mx, my, o.shape[1], o.shape[0]]
            g = flat[jb["dst_offset"][b]: jb["dst_offset"][b] + o.size].reshape(o.shape)
            assert _maxdiff(g, oracle.to_u8(o)) <= (0 if sampler == "nn" else U8_TOL)
    # full-size property: 1080p frame, integer shift by (+37, -21) and back, nn
    big = torch.randint(0, 256, (1, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
    S = np.array([[[1.0, 0, 37], [0, 1, -21], [0, 0, 1]]])
    j1 = device.scan_jobs(S, (1080, 1920), (1080, 1920))
    # Scan grids sample at c*w/(w-1); use the bbox form (integer grid) for an exact round trip
    j1, box1 = device.bbox_jobs(S, (1080, 1920))
    mid = torch.zeros((int(box1[0][3]), int(box1[0][2]), 3), dtype=torch.uint8, device="cuda")
    device.warp(big, mid, j1, 3, "nn")
    j2, box2 = device.bbox_jobs(np.linalg.inv(S), tuple(mid.shape[:2]))
    back = torch.zeros((int(box2[0][3]), int(box2[0][2]), 3), dtype=torch.uint8, device="cuda")
    device.warp(mid, back, j2, 3, "nn")
    a = big[0].cpu().numpy()
    bk = back.cpu().numpy()
    assert box1[0][:2].tolist() == [37, -21] and box2[0][:2].tolist() == [-37, 21]
    # mid[r, c] = big[r, c] (offset folded into min_x/min_y); interior must survive the round trip
    assert np.array_equal(bk[2:1000, 2:1800], a[2:1000, 2:1800])


def _rand_h(rs, h, w, kind):
    """Homographies that exercise every tile class of the fast kernel."""
    c, s_ = np.cos, np.sin
    if kind == "jitter":
        src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
        dst = src + rs.uniform(-0.08, 0.08, (4, 2)) * [w, h]
        pts = np.concatenate([src, dst], 1).astype(np.float32)
        return oracle.dlt4_f64(pts, np.array([[0, 1, 2, 3]], np.int32))[0][0]
    if kind == "rot":
        a = rs.uniform(0, 2 * np.pi)
        R = np.array([[c(a), -s_(a), 0], [s_(a), c(a), 0], [0, 0, 1]])
        T = np.array([[1, 0, w / 2], [0, 1, h / 2], [0, 0, 1.0]])
        return T @ R @ np.linalg.inv(T)
    if kind == "zoom":
        z = rs.uniform(0.3, 3.0)
        return np.array([[z, 0, rs.uniform(-20, 20)], [0, z * rs.uniform(0.8, 1.25), rs.uniform(-20, 20)], [0, 0, 1.0]])
    if kind == "persp":   # strong perspective: w changes a lot, some tiles must take the float64 path
        return np.array([[1.0, 0.1, 5], [0.05, 1.0, -7], [rs.uniform(-4e-3, 4e-3), rs.uniform(-4e-3, 4e-3), 1.0]])
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["jitter", "rot", "zoom", "persp"])
@pytest.mark.parametrize("sampler", ["nn", "bilinear"])
def test_fast_kernel_vs_float64_kernel_random(kind, sampler):
    """Fast uint8 kernel against the float64 kernel (itself bit-identical to the oracle) on
    homographies / sizes that hit interior, border, zero and slow tiles, aligned and unaligned
    pitches; Scan and bbox grids.  nn: identical.  bilinear: +-1, bounded off-by-one share."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    ctx = _cabi.context()
    rs = np.random.RandomState(hash((kind, sampler)) % 2**31)
    for (h, w) in [(240, 320), (203, 331), (97, 64), (33, 1000)]:
        B = 3
        frames = rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
        frames[1, :, : w // 2] = 200                      # a flat region: must be reproduced exactly
        Hs = np.stack([_rand_h(rs, h, w, kind) for _ in range(B)])
        src_t = torch.from_numpy(frames).cuda()
        for form in ("scan", "bbox"):
            if form == "scan":
                jobs = device.scan_jobs(Hs, (h, w), (h, w))
            else:
                jobs, _ = device.bbox_jobs(Hs, (h, w))
                if int((jobs["dst_pitch"] * jobs["oh"]).sum()) > 200e6:
                    continue
            total = int((jobs["dst_offset"] + jobs["dst_pitch"] * jobs["oh"]).max())
            out = {}
            for m in ("fast", "exact"):
                ctx.warp_exact(m == "exact")
                dst_t = torch.full((total,), 7, dtype=torch.uint8, device="cuda")
                device.warp(src_t, dst_t, jobs, 3, sampler)
                out[m] = dst_t.cpu().numpy()
            ctx.warp_exact(False)
            d = np.abs(out["fast"].astype(np.int16) - out["exact"].astype(np.int16))
            if sampler == "nn":
                assert d.max() == 0, (kind, form, h, w, int(d.max()), float((d > 0).mean()))
            else:
                assert d.max() <= U8_TOL, (kind, form, h, w, int(d.max()), float((d > 1).mean()))
                assert (d > 0).mean() < 0.02, (kind, form, h, w, float((d > 0).mean()))
        # flat half of frame 1 under the identity-like Scan grid stays exactly 200 where unmasked
    # the float64 kernel itself against the oracle on one small case
    h, w = 60, 90
    f = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    H = _rand_h(rs, h, w, kind)
    ctx.warp_exact(False)
    got = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    device.warp(torch.from_numpy(f).cuda(), got, device.scan_jobs(H[None], (h, w), (h, w)), 3, sampler)
    want = oracle.to_u8(oracle.wrapPerspectiveScan(f, H, [h, w], sampler)[0])
    assert _maxdiff(got.cpu().numpy(), want) <= (0 if sampler == "nn" else U8_TOL)


def test_fast_kernel_band_statistics():
    """The float64 fall-back must stay rare on the benchmark's workload shape (1080p, +-5% corner
    jitter): it only serves pixels within the coordinate-error band of a mask / rounding boundary."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    ctx = _cabi.context()
    rs = np.random.RandomState(11)
    h, w, B = 1080, 1920, 4
    Hs = np.stack([_rand_h(rs, h, w, "jitter") for _ in range(B)])
    src = torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    jobs = device.scan_jobs(Hs, (h, w), (h, w))
    for sampler, limit in (("bilinear", 1e-4), ("nn", 3e-3)):
        ctx.warp_exact_pixels(reset=True)
        device.warp(src, dst, jobs, 3, sampler)
        frac = ctx.warp_exact_pixels(reset=True) / (B * h * w)
        assert frac < limit, (sampler, frac)


@pytest.mark.parametrize("sampler", ["nn", "bilinear"])
def test_host_batch_pipeline_equals_device_path(sampler):
    """b200geo_warp_batch_host (chunked three-stream pipeline over host buffers) must give exactly
    what the device-resident launch gives; ragged batch (not a multiple of the chunk), bbox outputs
    of different sizes, pinned and pageable host memory."""
    import torch
    from ransac_with_homography_b200 import device
    rs = np.random.RandomState(5)
    B, h, w = 11, 120, 208
    frames = rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
    Hs = np.stack([_rand_h(rs, h, w, "jitter") for _ in range(B)])
    for form in ("scan", "bbox"):
        jobs = device.scan_jobs(Hs, (h, w), (h, w)) if form == "scan" else device.bbox_jobs(Hs, (h, w))[0]
        total = int((jobs["dst_offset"] + jobs["dst_pitch"] * jobs["oh"]).max())
        want = torch.zeros(total, dtype=torch.uint8, device="cuda")
        device.warp(torch.from_numpy(frames).cuda(), want, jobs, 3, sampler)
        want = want.cpu().numpy()
        for pinned in (True, False):
            src_h = torch.from_numpy(frames.copy())
            out_h = torch.zeros(total, dtype=torch.uint8)
            if pinned:
                src_h, out_h = src_h.pin_memory(), out_h.pin_memory()
            device.warp_host(src_h, out_h, jobs, 3, sampler, chunk_jobs=4)
            # (the pipeline re-bases the frames into its staging buffers, so a tile may be routed to the other fast kernel:
            # nn is bit-exact either way, bilinear stays within the +-1 both keep against the float64 kernel)
            d = np.abs(out_h.numpy().astype(np.int16) - want.astype(np.int16))
            assert d.max() <= (0 if sampler == "nn" else U8_TOL), (form, pinned, int(d.max()))


def test_device_stitch_compose_matches_host_entry(golden, capsys):
    """device.stitch_compose (device-resident images, b200geo_stitch_compose_dev) against the golden canvases."""
    import torch
    from ransac_with_homography_b200 import device
    w = golden.file("warps")
    for name in ["pp", "nn_", "np_", "pn_"]:
        H = w[f"pano_{name}_H"]
        q, t = torch.from_numpy(w["img2"].copy()).cuda(), torch.from_numpy(w["img"].copy()).cuda()
        assert np.array_equal(device.stitch_compose(q, t, H).cpu().numpy(), w[f"pano_{name}_paste"])
        assert np.array_equal(device.stitch_compose(q, t, H, "Rate", 0.2).cpu().numpy(), w[f"pano_{name}_blend"])
        assert np.array_equal(t.cpu().numpy(), w["img"])            # the device entry point leaves imgT untouched


@pytest.mark.parametrize("kind", ["jitter", "rot", "zoom", "persp"])
@pytest.mark.parametrize("sampler", ["nn", "bilinear"])
def test_layout_hints_do_not_change_results(kind, sampler):
    """b200geo_warp_batch_dev with no layout, the 2-D layout and the 3-D layout (per-image zero fill: edge
    footprints arrive by tensor copy too; full tiles go to the per-warp strip kernel) against the float64 kernel."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    ctx = _cabi.context()
    rs = np.random.RandomState(hash((kind, sampler, "layout")) % 2**31)
    for (h, w) in [(240, 320), (203, 332), (97, 128)]:
        B = 4
        frames = rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
        Hs = np.stack([_rand_h(rs, h, w, kind) for _ in range(B)])
        src_t = torch.from_numpy(frames).cuda()
        jobs = device.scan_jobs(Hs, (h, w), (h, w))
        jd = device.upload_jobs(jobs, "cuda")
        outs = []
        for layout in (None, (w * 3, B * h), (w * 3, B * h, h, w * 3)):
            dst_t = torch.full((B, h, w, 3), 9, dtype=torch.uint8, device="cuda")
            device.warp_resident(src_t, dst_t, jd, B, w, h, 3, sampler, layout=layout)
            outs.append(dst_t.cpu().numpy())
        # no layout and the 2-D layout run the same kernel (only the staging differs): identical bytes.  The 3-D layout
        # routes full tiles to the per-warp strip kernel, whose float32 coordinates are relative to other anchors:
        # nn stays identical (it is bit-exact by construction), bilinear stays within the same +-1 of the float64 kernel.
        assert np.array_equal(outs[0], outs[1]), (kind, h, w)
        if sampler == "nn":
            assert np.array_equal(outs[0], outs[2]), (kind, h, w)
        ctx.warp_exact(True)
        dst_t = torch.full((B, h, w, 3), 9, dtype=torch.uint8, device="cuda")
        device.warp_resident(src_t, dst_t, jd, B, w, h, 3, sampler)
        ctx.warp_exact(False)
        for o in outs:
            d = np.abs(o.astype(np.int16) - dst_t.cpu().numpy().astype(np.int16))
            assert d.max() <= (0 if sampler == "nn" else U8_TOL), (kind, h, w, int(d.max()))


def test_border_tile_far_from_its_anchor_regression():
    """A strong-perspective frame whose BORDER tiles clamp coordinates more than 256 px away from the tile anchor:
    the integer part of the 16.16 coordinate was once taken with a byte permute that only sign-extends 24-bit values
    (illegal shared-memory address; found by tools/fuzz_warp.py, seed 1 case 261)."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    ctx = _cabi.context()
    h, w = 114, 744
    H = np.array([[1.0, 0.1, 5.0], [0.05, 1.0, -7.0], [1.9822643559266172e-03, -7.9026834333958615e-04, 1.0]])
    frames = np.random.RandomState(5).randint(0, 256, (1, h, w, 3)).astype(np.uint8)
    jobs = device.scan_jobs(H[None], (h, w), (h, w))
    src_t = torch.from_numpy(frames).cuda()
    out = {}
    for sampler in ("bilinear", "nn"):
        for m in ("fast", "exact"):
            ctx.warp_exact(m == "exact")
            dst_t = torch.full((h * w * 3,), 7, dtype=torch.uint8, device="cuda")
            device.warp(src_t, dst_t, jobs, 3, sampler)
            torch.cuda.synchronize()
            out[m] = dst_t.cpu().numpy().astype(np.int16)
        ctx.warp_exact(False)
        assert np.abs(out["fast"] - out["exact"]).max() <= (0 if sampler == "nn" else U8_TOL)


def test_fuzz_fast_kernel_short():
    """A short run of tools/fuzz_warp.py (random sizes, batches, homography families, samplers, grids, layout hints)."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_warp", os.path.join(root, "tools", "fuzz_warp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv = sys.argv
    try:
        sys.argv = ["fuzz_warp.py", "80", "7"]
        assert mod.main() == 0
    finally:
        sys.argv = argv


def test_calls_on_two_streams_do_not_interleave():
    """One context shares its scratch buffers between calls; a call that arrives on another stream than the previous one
    waits for it: alternating two torch streams gives the results of the serial order."""
    import torch
    from ransac_with_homography_b200 import device
    rs = np.random.RandomState(3)
    h, w, B = 240, 320, 6
    frames = torch.from_numpy(rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)).cuda()
    Hs = np.stack([_rand_h(rs, h, w, "jitter") for _ in range(B)])
    jobs = device.scan_jobs(Hs, (h, w), (h, w))
    want = torch.empty_like(frames)
    device.warp(frames, want, jobs, 3, "bilinear")
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [torch.empty_like(frames) for _ in range(8)]
    for i, o in enumerate(outs):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            device.warp(frames, o, jobs, 3, "bilinear")
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, want)


def test_malformed_jobs_are_rejected_or_skipped():
    """Pitches smaller than a row: host-visible job arrays are refused with an argument error; jobs that are
    already on the device are skipped by the kernels (nothing is read or written out of bounds)."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    h, w = 96, 128
    frames = torch.randint(0, 256, (2, h, w, 3), dtype=torch.uint8, device="cuda")
    jobs = device.scan_jobs(np.stack([np.eye(3), np.eye(3)]), (h, w), (h, w))
    bad = jobs.copy()
    bad["dst_pitch"][1] = w * 3 - 4
    out = torch.full((2, h, w, 3), 9, dtype=torch.uint8, device="cuda")
    with pytest.raises(_cabi.B200GeoError):
        device.warp(frames, out, bad, 3, "bilinear")
    bad2 = jobs.copy()
    bad2["src_pitch"][0] = w * 3 - 1
    with pytest.raises(_cabi.B200GeoError):
        device.warp_host(frames.cpu().numpy(), np.empty((2, h, w, 3), np.uint8), bad2, 3, "nn")
    for layout in (None, (w * 3, 2 * h, h, w * 3)):
        out.fill_(9)
        device.warp_resident(frames, out, device.upload_jobs(bad, "cuda"), 2, w, h, 3, "bilinear", layout=layout)
        torch.cuda.synchronize()
        assert bool((out[1] == 9).all()) and not bool((out[0] == 9).all())      # job 1 skipped, job 0 done
