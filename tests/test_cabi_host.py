"""CPU: the C-ABI library loads and exports every symbol include/b200geo.h declares; host-side
logic (grid parameters, job builders, sharding, key protocol, canvas placement)."""
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT
from ransac_with_homography_b200 import _cabi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200geo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200geo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200geo.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names
    assert lib.b200geo_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.B200GeoError) as e:
        _cabi.Context(0)
    assert e.value.code == _cabi.ERR_NODEVICE
    from ransac_with_homography_b200 import homography
    with pytest.raises(_cabi.B200GeoError):
        homography.calcHomographyLinear(np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]]), np.array([[0, 0], [2, 0], [2, 2], [0, 2.0]]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ransac_with_homography_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "liboracle" not in src, f


def test_warp_job_struct_layout():
    assert _cabi.WARP_JOB_DTYPE.itemsize == 184
    for name, _ in _cabi.WarpJob._fields_:
        assert getattr(_cabi.WarpJob, name).offset == _cabi.WARP_JOB_DTYPE.fields[name][1]


def test_abi_struct_sizes_match_the_header():
    """ctypes mirrors of the structs include/b200geo.h declares; the C side static_asserts the same sizes (csrc/common.cuh),
    and the header must list the fields in the order the mirrors use."""
    import ctypes as C
    assert C.sizeof(_cabi.WarpJob) == 184 and C.sizeof(_cabi.Compose) == 80 and C.sizeof(_cabi.SrcLayout) == 32
    hdr = open(os.path.join(ROOT, "include", "b200geo.h")).read()
    for struct, cls in (("b200geo_compose", _cabi.Compose), ("b200geo_src_layout", _cabi.SrcLayout), ("b200geo_warp_job", _cabi.WarpJob)):
        body = re.search(r"typedef struct " + struct + r" \{(.*?)\} " + struct + ";", hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [n.split("[")[0] for decl in body.split(";") if decl.strip() for n in re.sub(r"^\s*\w+\s+", "", decl.strip()).replace(" ", "").split(",")]
        mine = [n for n, _ in cls._fields_]
        assert [n.replace("reserved0", "r0") for n in names] == [n.replace("reserved0", "r0") for n in mine], (struct, names, mine)


@pytest.mark.parametrize("start,stop,num", [(0, 2480, 2480), (0, 1080, 1080), (-27, 170, 198), (5, 5, 1), (0, 7, 7), (-3, 69, 73)])
def test_linspace_params_reproduce_numpy(start, stop, num):
    from ransac_with_homography_b200.homography import _linspace_params
    x0, dx, last = _linspace_params(start, stop, num)
    i = np.arange(num, dtype=np.float64)
    mine = i * dx + x0
    mine[-1] = last
    assert np.array_equal(mine, np.linspace(start, stop, num))


def test_job_builders_match_oracle_grids():
    from ransac_with_homography_b200 import device
    rs = np.random.RandomState(0)
    H = np.stack([np.array([[1 + 0.05 * rs.randn(), 0.03 * rs.randn(), 30 * rs.randn()],
                            [0.03 * rs.randn(), 1 + 0.05 * rs.randn(), 20 * rs.randn()],
                            [1e-5 * rs.randn(), 1e-5 * rs.randn(), 1]]) for _ in range(5)])
    jobs, boxes = device.bbox_jobs(H, (108, 192))
    for b in range(5):
        gx, gy, mx, my = oracle.bbox_grid(H[b], 108, 192)
        assert boxes[b].tolist() == [mx, my, len(gx), len(gy)]
        assert jobs["x0"][b] == gx[0] and jobs["x_last"][b] == gx[-1] and jobs["dx"][b] == 1.0
        assert np.array_equal(jobs["invH"][b].reshape(3, 3), np.linalg.inv(H[b]))
    sj = device.scan_jobs(H, (108, 192), (100, 120))
    gx, gy = oracle.scan_grid((100, 120))
    i = np.arange(120.0)
    assert np.array_equal((i * sj["dx"][0] + sj["x0"][0])[:-1], gx[:-1]) and sj["x_last"][0] == gx[-1]
    assert sj["bw"][0] == 120 and sj["bh"][0] == 100 and sj["iw"][0] == 192
    assert sj["src_offset"][2] == 2 * 108 * 192 * 3 and sj["dst_offset"][3] == 3 * 100 * 120 * 3


def test_shard_range_partitions_exactly():
    from ransac_with_homography_b200.device import shard_range
    for K in (0, 1, 7, 1000, 1_000_000, 8193):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(K, i, world) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == K
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _pack_keys(counts, k0, thr):
    """python statement of the key protocol of select_kernel (include/b200geo.h, (2b))"""
    best, early = 0, 0
    for i, c in enumerate(counts.tolist()):
        k = k0 + i
        best = max(best, (c << 32) | (0xFFFFFFFF - k))
        if c >= thr:
            early = max(early, 0x100000000 - k)
    return best, early


def _decode(best, early):
    if early:
        return 0x100000000 - early
    if (best >> 32) == 0:
        return -1
    return 0xFFFFFFFF - (best & 0xFFFFFFFF)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_key_protocol_equals_sequential_rule(golden, world):
    """Sharded max-reduction of the two keys == the reference's sequential winner rule."""
    from ransac_with_homography_b200.device import shard_range
    rs = np.random.RandomState(3)
    cases = [golden.case("ransac_fixture", n)["counts"] for n in ("A_s0", "E_s1", "B_s2")]
    cases += [rs.randint(0, 50, 1000).astype(np.int32), np.zeros(64, np.int32), rs.randint(0, 3, 17).astype(np.int32)]
    for counts in cases:
        for d in (10, 60, 95):
            M = 185
            thr = M * d / 100 + 4
            keys = [_pack_keys(counts[a:b], a, thr) for a, b in (shard_range(len(counts), r, world) for r in range(world))]
            best, early = max(k[0] for k in keys), max(k[1] for k in keys)
            win, cnt, _ = oracle.select(counts, M, d)
            assert _decode(best, early) == win
            if win >= 0 and not early:
                assert (best >> 32) == cnt


def _gloo_worker(rank, world, port, counts, thr, out):
    import torch
    import torch.distributed as dist
    from ransac_with_homography_b200.device import shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a, b = shard_range(len(counts), rank, world)
    best, early = _pack_keys(counts[a:b], a, thr)
    keys = torch.tensor([best, early], dtype=torch.int64)
    dist.all_reduce(keys, op=dist.ReduceOp.MAX)
    out[rank] = _decode(int(keys[0]), int(keys[1]))
    dist.destroy_process_group()


def test_two_rank_gloo_all_reduce_picks_reference_winner(golden):
    """world_size-2 gloo run of the N>1 selection path (what device.ransac does over NCCL)."""
    import torch.multiprocessing as mp
    counts = golden.case("ransac_fixture", "A_s1")["counts"]
    thr = 185 * 70 / 100 + 4
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        procs = [mp.get_context("spawn").Process(target=_gloo_worker, args=(r, 2, port, counts, thr, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        win, _, _ = oracle.select(counts, 185, 70)
        assert out[0] == win and out[1] == win


def test_canvas_placement_all_sign_cases(golden):
    from ransac_with_homography_b200.homography import _placement
    w = golden.file("warps")
    hq, wq = w["img2"].shape[:2]
    for name in ("pp", "nn_", "np_", "pn_"):
        H = w[f"pano_{name}_H"]
        t, mx, my = oracle.wrapPerspective(w["img"], H, "bilinear")
        (tsx, tsy, tex, tey), (qsx, qsy, qex, qey), fw, fh = _placement(mx, my, t.shape[1], t.shape[0], wq, hq)
        assert (fh, fw) == w[f"pano_{name}_paste"].shape[:2]
        canvas = np.zeros((fh, fw, 3), np.uint8)
        canvas[tsy:tey + 1, tsx:tex + 1] = oracle.to_u8(t)
        canvas[qsy:qey + 1, qsx:qex + 1] = w["img2"]
        assert np.array_equal(canvas, w[f"pano_{name}_paste"])


def test_dropin_modules_resolve_under_the_reference_names():
    """`from homography import transformImage` / `from ransac import stitching` (reference app.py:9-10)
    must resolve to this package when dropin/ is on the path; importing needs no GPU."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "dropin"))
    saved = {k: sys.modules.pop(k, None) for k in ("homography", "ransac")}
    try:
        hg, rs = importlib.import_module("homography"), importlib.import_module("ransac")
        import ransac_with_homography_b200.homography as ours_h
        import ransac_with_homography_b200.ransac as ours_r
        assert hg.transformImage is ours_h.transformImage and hg.transformImageH is ours_h.transformImageH
        assert rs.stitching is ours_r.stitching and rs.RANSAC is ours_r.RANSAC and rs.HomoModel is ours_r.HomoModel
        for name in ("calcHomography", "calcHomographyLinear", "wrapPerspective", "wrapPerspectiveScan", "nearestNeighbor",
                     "bilinear", "stitchPanorama", "addAlpha", "BLENDDIR", "cylindericlMap"):
            assert hasattr(hg, name), name
    finally:
        sys.path.pop(0)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
