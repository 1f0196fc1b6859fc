"""GPU parity at the BENCHMARKED configurations: the fast uint8 warp through the exact launch paths
bench.py times -- C4: 1080p frames, 3-D tensor-map layout, Scan grid; C2: 755 x 1226 -> 3508 x 2480 Scanner; C5:
device.stitch_batched with 1080p frames into 2160 x 3840 canvas slots -- and configs 1 / 2 on the reference's own images,
against the CPU oracle (itself pinned to the live reference at these sizes by tests/test_oracle_full_size.py).
The census of every tolerated difference is written to gpurun_out/PARITY_r02.json (copied to profiles/ by hand).

Tolerances (BASELINE.json north_star): nn bit-exact; bilinear within +-1 on uint8 (stated as U8_TOL)."""
import io
import json
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

import fullsize
import oracle
from conftest import FIXTURE_CASES, case_method, synth_pair

pytestmark = pytest.mark.gpu

U8_TOL = 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "PARITY_r02.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _diff_stats(got, want):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    return {"max_abs": int(d.max()), "off_by_one_share": float((d == 1).mean()), "beyond_tol_share": float((d > U8_TOL).mean())}


def _frame_h(rs, h, w, jitter=0.05):
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
    dst = src + rs.uniform(-jitter, jitter, (4, 2)) * [w, h]
    A = np.zeros((8, 8)); b = np.zeros(8)
    for j, ((x, y), (X, Y)) in enumerate(zip(src, dst)):
        A[2 * j] = [x, y, 1, 0, 0, 0, -x * X, -y * X]; A[2 * j + 1] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
        b[2 * j], b[2 * j + 1] = X, Y
    return np.append(np.linalg.solve(A, b), 1).reshape(3, 3)


# ------------------------------------------------------------------------------------------------ configs 1 and 2
def test_config1_stitch_on_the_reference_images():
    """foto1A warped by the reference's own RANSAC result and composed with foto1B (ransac.py:279), full size."""
    from ransac_with_homography_b200 import homography as hg
    m = fullsize.meta()
    H = np.array(m["H_stitch"])
    A, B = fullsize.image("foto1A.jpg"), fullsize.image("foto1B.jpg")
    want, off = fullsize.oracle_case("c1_tih_bilinear")
    got, mx, my = hg.transformImageH(A.copy(), H)
    assert [mx, my] == off == m["cases"]["c1_tih_bilinear"]["off"] and got.shape == want.shape
    st = _diff_stats(got, want)
    assert st["max_abs"] <= U8_TOL
    REPORT["c1_transformImageH_683x1024"] = st
    want, off = fullsize.oracle_case("c1_bbox_nn")
    got, mx, my = hg.wrapPerspective(A.copy(), H, "nn")
    assert [mx, my] == off and np.array_equal(got, want)
    if fullsize.decoder_matches():
        assert fullsize.same(got, m["cases"]["c1_bbox_nn"])                       # == the live reference's bytes
    # compose (K5): float64 like the reference -> identical canvases
    with redirect_stdout(io.StringIO()):
        paste = hg.stitchPanorama(B.copy(), A.copy(), H)
        blend = hg.stitchPanorama(B.copy(), A.copy(), H, blending="Rate", blendrate=0.2)
        swapped = hg.stitchPanorama(A.copy(), B.copy(), np.array(m["H_stitch_inv"]))
    assert np.array_equal(paste, fullsize.oracle_case("c1_pano_paste")[0])
    assert np.array_equal(blend, fullsize.oracle_case("c1_pano_blend")[0])
    assert np.array_equal(swapped, fullsize.oracle_case("c1_pano_paste_swapped")[0])
    if fullsize.decoder_matches():
        assert fullsize.same(paste, m["cases"]["c1_pano_paste"]) and fullsize.same(blend, m["cases"]["c1_pano_blend"])
    REPORT["c1_stitchPanorama"] = {"paste_identical": True, "blend_identical": True, "canvas": list(paste.shape)}


@pytest.mark.parametrize("tag", ["a4_300dpi", "a4_app"])
def test_config2_scanner_full_size(tag):
    """cv.jpg (755 x 1226) rectified to A4 at 300 dpi (3508 x 2480) and to the app's A4 (1188 x 840) through
    transformImage(box=...) (homography.py:211-228): same H on both sides, then the end-to-end call with our own solve."""
    from ransac_with_homography_b200 import _cabi, homography as hg
    ctx = _cabi.context()
    m = fullsize.meta()
    hh, ww = fullsize.SCAN_TAGS[tag]
    u, v = fullsize.scan_uv(tag)
    cv = fullsize.image("cv.jpg")
    for conv in ("bilinear", "nn"):
        c = m["cases"][f"c2_{tag}_{conv}"]
        want, _ = fullsize.oracle_case(f"c2_{tag}_{conv}")
        H = np.array(c["H"])
        got, _, _ = hg.wrapPerspectiveScan(cv.copy(), H, [hh, ww], conv, _out="u8")
        got = fullsize.crop_transform_image(got, v, 0, 0)
        assert got.shape == want.shape == tuple(c["shape"])
        st = _diff_stats(got, want)
        assert st["max_abs"] <= (0 if conv == "nn" else U8_TOL), st
        REPORT[f"c2_{tag}_{conv}"] = st
        if conv == "nn" and fullsize.decoder_matches():
            assert fullsize.same(got, c)
        if conv == "bilinear":      # the float64 kernel reproduces the reference's bytes
            ctx.warp_exact(True)
            try:
                ex, _, _ = hg.wrapPerspectiveScan(cv.copy(), H, [hh, ww], conv, _out="u8")
            finally:
                ctx.warp_exact(False)
            assert np.array_equal(fullsize.crop_transform_image(ex, v, 0, 0), want)
    # end to end: transformImage solves the 4-point system itself (float64; the reference: float32 normal equations)
    H_ours = hg.calcHomographyLinear(u.T[:, :2], v.T[:, :2])
    H_ref = np.array(m["cases"][f"c2_{tag}_bilinear"]["H"])
    REPORT[f"c2_{tag}_H_rel_to_reference_f32_solve"] = float(np.linalg.norm(H_ours - H_ref) / np.linalg.norm(H_ref))
    got = hg.transformImage(cv.copy(), u, v, box=[hh, ww], method="bilinear")
    o, _, _ = oracle.wrapPerspectiveScan(cv, H_ours, [hh, ww], "bilinear")
    want = fullsize.crop_transform_image(oracle.to_u8(o), v, 0, 0)
    assert got.shape == want.shape and _diff_stats(got, want)["max_abs"] <= U8_TOL


def test_notebook_example_full_size():
    from ransac_with_homography_b200 import homography as hg
    m = fullsize.meta()
    want, off = fullsize.oracle_case("notebook_bilinear")
    got, mx, my = hg.wrapPerspective(fullsize.image("notebook.jpg"), np.array(m["H_notebook"]), convert="bilinear", _out="u8")
    assert [mx, my] == off and got.shape == want.shape and _diff_stats(got, want)["max_abs"] <= U8_TOL


# ------------------------------------------------------------------------------------------------ config 4: the timed path
@pytest.mark.parametrize("sampler", ["bilinear", "nn"])
def test_config4_benchmark_path_at_1080p(sampler):
    """bench.py's C4 launch: device.warp_resident(..., layout=(w*3, n*h, h, w*3)) on 1080 x 1920 frames, Scan grid,
    +-5 % corner jitter -- against oracle.wrapPerspectiveScan on every frame and against the float64 kernel."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    ctx = _cabi.context()
    h, w, n = 1080, 1920, 8
    rs = np.random.RandomState(2024)
    frames = rs.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    frames[1] = np.stack([(xx // 3) % 256, (yy // 2) % 256, ((xx + yy) // 5) % 256], -1).astype(np.uint8)   # smooth ramps
    frames[2, :, : w // 2] = 200                                                                        # a flat half
    frames[3] = np.tile(fullsize.image("cv.jpg"), (2, 2, 1))[:h, :w]                                    # a photograph
    Hs = np.stack([_frame_h(rs, h, w) for _ in range(n)])
    Hs[4] = np.array([[1.0, 0, 7.0], [0, 1.0, -5.0], [0, 0, 1.0]])                                       # integer translation
    Hs[5] = np.array([[1.0, 0, 3.5], [0, 1.0, 0.25], [0, 0, 1.0]])                                       # dyadic translation
    jobs = device.scan_jobs(Hs, (h, w), (h, w))
    jd = device.upload_jobs(jobs, "cuda")
    src = torch.from_numpy(frames).cuda()
    dst = torch.full((n, h, w, 3), 9, dtype=torch.uint8, device="cuda")
    ctx.warp_exact_pixels(reset=True)
    device.warp_resident(src, dst, jd, n, w, h, 3, sampler, layout=(w * 3, n * h, h, w * 3))
    fallback = ctx.warp_exact_pixels(reset=True) / (n * h * w)
    got = dst.cpu().numpy()
    ctx.warp_exact(True)
    try:
        ex = torch.empty_like(dst)
        device.warp_resident(src, ex, jd, n, w, h, 3, sampler)
    finally:
        ctx.warp_exact(False)
    ex = ex.cpu().numpy()
    worst = {"max_abs": 0, "off_by_one_share": 0.0, "beyond_tol_share": 0.0}
    for i in range(n):
        want = oracle.to_u8(oracle.wrapPerspectiveScan(frames[i], Hs[i], [h, w], sampler)[0])
        assert np.array_equal(ex[i], want), i                         # float64 kernel == oracle, every byte
        st = _diff_stats(got[i], want)
        assert st["max_abs"] <= (0 if sampler == "nn" else U8_TOL), (i, st)
        worst = {k: max(worst[k], st[k]) for k in worst}
    REPORT[f"c4_1080p_3d_tmap_{sampler}"] = dict(worst, float64_fallback_pixel_share=float(fallback), frames=n)
    # the other launch forms of the same frames give the same answer within the same tolerance
    for variant, layout in ((2, (w * 3, n * h, h, w * 3)), (0, None), (0, (w * 3, n * h))):
        ctx.warp_variant(variant)
        try:
            d2 = torch.full((n, h, w, 3), 9, dtype=torch.uint8, device="cuda")
            device.warp_resident(src, d2, jd, n, w, h, 3, sampler, layout=layout)
        finally:
            ctx.warp_variant(0)
        d = np.abs(d2.cpu().numpy().astype(np.int16) - ex.astype(np.int16))
        assert d.max() <= (0 if sampler == "nn" else U8_TOL), (variant, layout)


def test_config4_bbox_semantics_subset():
    """C4's parity subset with wrapPerspective (bbox) semantics at 1080p, 3-D layout."""
    import torch
    from ransac_with_homography_b200 import device
    h, w, n = 1080, 1920, 3
    rs = np.random.RandomState(77)
    frames = rs.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    Hs = np.stack([_frame_h(rs, h, w) for _ in range(n)])
    jobs, boxes = device.bbox_jobs(Hs, (h, w))
    total = int((jobs["dst_offset"] + jobs["dst_pitch"] * jobs["oh"]).max())
    dst = torch.full((total,), 9, dtype=torch.uint8, device="cuda")
    device.warp_resident(torch.from_numpy(frames).cuda(), dst, device.upload_jobs(jobs, "cuda"), n, int(jobs["ow"].max()), int(jobs["oh"].max()),
                         3, "bilinear", layout=(w * 3, n * h, h, w * 3))
    got = dst.cpu().numpy()
    for i in range(n):
        o, mx, my = oracle.wrapPerspective(frames[i], Hs[i], "bilinear")
        want = oracle.to_u8(o)
        assert [mx, my] == boxes[i, :2].tolist() and list(want.shape[:2]) == [boxes[i, 3], boxes[i, 2]]
        g = got[jobs["dst_offset"][i]: jobs["dst_offset"][i] + want.size].reshape(want.shape)
        assert _diff_stats(g, want)["max_abs"] <= U8_TOL


# ------------------------------------------------------------------------------------------------ config 5
def test_config5_stitch_batched_full_size():
    """device.stitch_batched at C5's shapes: M = 2000 correspondences x K = 8192 sample sets per pair, 1080p frames warped
    into 2160 x 3840 canvas slots -- against RANSAC.run + the oracle's wrapPerspective per pair."""
    import torch
    from bench import C5, stitch_pairs
    from ransac_with_homography_b200 import device
    from ransac_with_homography_b200.ransac import HomoModel, RANSAC
    B, h, w, canvas = 4, C5["h"], C5["w"], C5["canvas"]
    pts, idx = stitch_pairs(0, B)
    rs = np.random.RandomState(3)
    frames = rs.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
    canv, boxes, results, masks = device.stitch_batched(torch.from_numpy(pts).cuda(), torch.from_numpy(idx).cuda(),
                                                        torch.from_numpy(frames).cuda(), C5["th"], C5["d"], canvas)
    canv, res, masks = canv.cpu().numpy(), results.cpu().numpy(), masks.cpu().numpy()
    boxes = boxes.cpu().numpy()
    assert boxes[:, 4].all()
    worst = 0
    for b in range(B):
        with redirect_stdout(io.StringIO()):
            H, inl, cnt = RANSAC(HomoModel(th=C5["th"], d=C5["d"], n=4), k=C5["K"]).run([pts[b, :, :2].T, pts[b, :, 2:].T], method="fwd", idx=idx[b])
        assert int(res[b, 1]) == int(cnt) and np.array_equal(np.where(masks[b])[0], inl[0])
        np.testing.assert_allclose(res[b, 12:21].reshape(3, 3), H, rtol=1e-12, atol=1e-12)
        o, mx, my = oracle.wrapPerspective(frames[b], H, "bilinear")
        want = oracle.to_u8(o)
        assert boxes[b, :2].tolist() == [mx, my]
        oh, ow = min(want.shape[0], canvas[0]), min(want.shape[1], canvas[1])
        assert boxes[b, 2:4].tolist() == [ow, oh]
        st = _diff_stats(canv[b, :oh, :ow], want[:oh, :ow])
        assert st["max_abs"] <= U8_TOL, (b, st)
        worst = max(worst, st["off_by_one_share"])
        assert not canv[b, oh:].any() and not canv[b, :, ow:].any()       # the rest of the slot stays zero
    REPORT["c5_stitch_batched_1080p_into_2160x3840"] = {"pairs": B, "max_abs": 1 if worst else 0, "off_by_one_share": float(worst)}


# ------------------------------------------------------------------------------------------------ RANSAC census
def test_ransac_census_against_the_live_reference(golden):
    """Whole-run parity on the 9 golden runs of the live reference (north_star: chosen hypothesis and inlier set exact; the
    differences that the reference's own float32 noise causes are REPORTED): per case the share of hypotheses whose count
    differs from the reference's (its per-hypothesis H is a float32 SVD), winner / inlier-set agreement, the distance of the
    refit H from the reference's float32 normal-equation refit, and the duplicate-index samples (scored 0 here)."""
    from ransac_with_homography_b200.ransac import HomoModel, RANSAC
    A, B = golden.points()
    pts = oracle.pack_points(A.T, B.T)
    out = {}
    for name in FIXTURE_CASES:
        c = golden.case("ransac_fixture", name)
        m = case_method(name)
        th, d, k = float(c["th"]), float(c["d"]), int(c["k"])
        with redirect_stdout(io.StringIO()):
            H, inl, cnt, det = RANSAC(HomoModel(th=th, d=d, n=4), k=k).run([A.T, B.T], method=m, idx=c["idx"], return_details=True)
        dup = np.array([len(set(r.tolist())) < 4 for r in c["idx"]])
        ok = c["finite"] & ~dup
        ref_win = oracle.select(c["counts"], 185, d)[0]
        Hr = c["run_H"]
        corners = np.array([[0, 1023, 1023, 0], [0, 0, 682, 682], [1.0, 1, 1, 1]])
        p, q = H @ corners, Hr @ corners
        out[name] = {
            "hypotheses": k, "count_mismatch_share": float((det["counts"][ok] != c["counts"][ok]).mean()),
            "count_max_abs_diff": int(np.abs(det["counts"][ok].astype(int) - c["counts"][ok]).max()),
            "winner": int(det["winner"]), "reference_winner": int(ref_win), "winner_equal": bool(det["winner"] == ref_win),
            "inlier_set_equal": bool(np.array_equal(inl[0], c["run_inliers"])), "count": int(cnt), "reference_count": int(c["run_count"]),
            "H_rel_to_reference_f32_refit": float(np.linalg.norm(H - Hr) / np.linalg.norm(Hr)),
            "H_rel_to_float64_lstsq": float(np.linalg.norm(H - oracle.refit_f64(pts, np.isin(np.arange(185), inl[0]))) /
                                            np.linalg.norm(oracle.refit_f64(pts, np.isin(np.arange(185), inl[0])))),
            "corner_px_vs_reference": float(np.abs(p[:2] / p[2] - q[:2] / q[2]).max()),
            "duplicate_index_samples": int(dup.sum()), "max_reference_count_on_duplicates": int(c["counts"][dup].max()) if dup.any() else 0,
        }
        # the chosen hypothesis and its inlier set equal the live reference's on every golden run
        assert out[name]["winner_equal"] and out[name]["inlier_set_equal"], (name, out[name])
        assert out[name]["H_rel_to_float64_lstsq"] < 1e-4
        assert out[name]["max_reference_count_on_duplicates"] < int(c["run_count"])      # a degenerate sample never was the reference's winner
    REPORT["ransac_fixture_runs"] = out


def test_c3_band_hit_census():
    """The pairs of C3 (20k correspondences x 1M sample sets) whose decision fell inside the rounding shell of the
    division-free test and were re-evaluated by the reference-order chain (north_star: 'reported')."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    ctx = _cabi.context()
    M, K = 20000, 1_000_000
    X, Y = synth_pair(1, M)
    pts = torch.from_numpy(oracle.pack_points(X, Y)).cuda()
    idx = torch.from_numpy(np.random.RandomState(0).randint(0, M, (K, 4)).astype(np.int32)).cuda()
    ctx.exact_path_pairs(reset=True)
    res, mask = device.ransac(pts, idx, 3.0, 99)
    hits = ctx.exact_path_pairs(reset=True)
    r = res.cpu().numpy()
    REPORT["c3_20k_x_1M"] = {"pairs": K * M, "exact_chain_pairs": int(hits), "share": hits / (K * M), "winner": int(r[0]), "inliers": int(r[1])}
    assert hits / (K * M) < 1e-3


def test_device_built_stitch_jobs_equal_the_numpy_rectangles():
    """b200geo_stitch_jobs_dev (the output rectangle of wrapPerspective and inv(H), computed by a kernel from the refit
    results) against device.bbox_jobs (numpy, the reference's expressions): census over 4096 random homographies of the C5
    family -- rectangles equal, inverse homographies equal to 1e-12 relative; failed / degenerate pairs flagged."""
    import torch
    from ransac_with_homography_b200 import _cabi, device
    rs = np.random.RandomState(17)
    B, h, w, canvas = 4096, 1080, 1920, (2160, 3840)
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
    Hs = np.empty((B, 3, 3))
    for i in range(B):
        dst = src + rs.uniform(-0.1, 0.1, (4, 2)) * [w, h] + rs.uniform(-1, 1, 2) * [900, 500]
        A = np.zeros((8, 8)); b = np.zeros(8)
        for j, ((x, y), (X, Y)) in enumerate(zip(src, dst)):
            A[2 * j] = [x, y, 1, 0, 0, 0, -x * X, -y * X]; A[2 * j + 1] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
            b[2 * j], b[2 * j + 1] = X, Y
        Hs[i] = np.append(np.linalg.solve(A, b), 1).reshape(3, 3)
    res = np.zeros((B, 32))
    res[:, 12:21] = Hs.reshape(B, 9)
    res[5, 2] = -4                                   # a failed search
    res[6, 12:21] = np.nan                           # a non-finite H
    res[7, 12:21] = [1, 0, 0, 0, 1, 0, -1.0 / (w - 1), 0, 1]   # a corner at w = 0
    jobs_t, boxes_t = device.stitch_jobs(torch.from_numpy(res).cuda(), (h, w), canvas)
    boxes = boxes_t.cpu().numpy()
    jobs = jobs_t.cpu().numpy().view(_cabi.WARP_JOB_DTYPE)
    good = np.ones(B, bool); good[[5, 6, 7]] = False
    assert not boxes[~good, 4].any() and boxes[good, 4].all()
    want_jobs, want_boxes = device.bbox_jobs(Hs[good], (h, w), canvas_hw=canvas)
    diff = (boxes[good, :4] != want_boxes).any(axis=1)
    REPORT["c5_device_built_rectangles"] = {"pairs": int(good.sum()), "rectangles_differing_from_numpy": int(diff.sum())}
    assert not diff.any()
    rel = np.abs(jobs["invH"][good] - want_jobs["invH"]).max(axis=1) / np.abs(want_jobs["invH"]).max(axis=1)
    assert rel.max() < 1e-12
    for f in ("x0", "dx", "x_last", "y0", "dy", "y_last", "ow", "oh", "bw", "bh", "iw", "ih", "src_pitch", "dst_pitch"):
        assert np.array_equal(jobs[f][good], want_jobs[f]), f
    assert np.array_equal(jobs["src_offset"], np.arange(B, dtype=np.int64) * (w * 3 * h))
    assert np.array_equal(jobs["dst_offset"], np.arange(B, dtype=np.int64) * (canvas[1] * 3 * canvas[0]))
    assert np.array_equal(jobs["invH"][~good], np.tile(np.eye(3).reshape(9), (3, 1)))
