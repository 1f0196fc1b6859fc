"""Device-resident / batched / multi-GPU interface to the hot path.

PyTorch is used here for plumbing only: device memory (tensors), streams and
``torch.distributed`` (NCCL).  All arithmetic goes through the ``_dev`` entry points of the C ABI
(include/b200geo.h) on the current torch CUDA stream; nothing here synchronises with the host
unless it says so.

Multi-GPU (one process per GPU):
  * single-pair RANSAC shards the K sample sets by hypothesis block; every rank reduces its block
    to two packed 64-bit keys and ONE ``all_reduce(MAX)`` of those 16 bytes picks the global winner
    (``ransac``).  The winner's H, inlier mask and refit are then recomputed on every rank from
    the shared index set -- no broadcast.
  * batched warps and batched pairs are sharded by frame / pair with no communication.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi


_peer_ready = {}      # device index -> world size the peer-memory inboxes were opened for (peer_setup)


def _ctx(t: torch.Tensor) -> _cabi.Context:
    assert t.is_cuda, "device tensors required (no CPU fallback)"
    return _cabi.context(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _stream(t: torch.Tensor):
    return _cabi._vp(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: torch.Tensor | None):
    return None if t is None else _cabi._vp(t.data_ptr())


def _chk_pts(pts: torch.Tensor):
    assert pts.dtype == torch.float32 and pts.is_contiguous() and pts.shape[-1] == 4, "pts: (..., M, 4) float32 rows (x, y, x', y')"


def dlt4(pts: torch.Tensor, idx: torch.Tensor, want_inv: bool = False):
    """K1.  pts (M,4) float32, idx (K,4) int32 -> H (K,3,3) float32, Hinv or None, flags (K,) uint8."""
    _chk_pts(pts)
    assert idx.dtype == torch.int32 and idx.is_contiguous() and idx.shape[-1] == 4
    ctx = _ctx(pts)
    K = idx.shape[0]
    H = torch.empty((K, 3, 3), dtype=torch.float32, device=pts.device)
    Hinv = torch.empty_like(H) if want_inv else None
    flags = torch.empty((K,), dtype=torch.uint8, device=pts.device)
    _cabi.check(ctx.lib.b200geo_dlt4_dev(ctx.handle, _p(pts), pts.shape[0], _p(idx), K, _p(H), _p(Hinv), _p(flags), _stream(pts)))
    return H, Hinv, flags


def score(pts: torch.Tensor, H: torch.Tensor, th: float, method: str = "fwd", Hinv: torch.Tensor | None = None,
          flags: torch.Tensor | None = None, counts: torch.Tensor | None = None) -> torch.Tensor:
    """K2.  Inlier count of every hypothesis H[k] over pts; int32 (K,)."""
    _chk_pts(pts)
    assert H.dtype == torch.float32 and H.is_contiguous()
    ctx = _ctx(pts)
    K = H.numel() // 9
    if counts is None:
        counts = torch.empty((K,), dtype=torch.int32, device=pts.device)
    _cabi.check(ctx.lib.b200geo_score_dev(ctx.handle, _p(pts), pts.shape[0], _p(H), _p(Hinv), _p(flags), K,
                                          float(np.float32(th)), _cabi.METHODS[method], _p(counts), _stream(pts)))
    return counts


def ransac_search(pts, idx, k_begin, k_end, th, thr, method="fwd", keys=None, counts=None, zero_keys=True):
    """Solve + score + select over idx[k_begin:k_end]; returns the (2,) int64 key tensor."""
    _chk_pts(pts)
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    ctx = _ctx(pts)
    if keys is None:
        keys = torch.empty((2,), dtype=torch.int64, device=pts.device)
    _cabi.check(ctx.lib.b200geo_ransac_search_dev(ctx.handle, _p(pts), pts.shape[0], _p(idx), int(k_begin), int(k_end),
                                                  float(np.float32(th)), float(thr), _cabi.METHODS[method], _p(keys),
                                                  int(bool(zero_keys)), _p(counts), _stream(pts)))
    return keys


def ransac_finish(pts, idx, keys, th, method="fwd", result=None, mask=None):
    """K3 from reduced keys: result (32,) float64 [winner, count, status, H_sample(9), H_refit(9)], mask (M,) uint8."""
    ctx = _ctx(pts)
    if result is None:
        result = torch.empty((32,), dtype=torch.float64, device=pts.device)
    if mask is None:
        mask = torch.empty((pts.shape[0],), dtype=torch.uint8, device=pts.device)
    _cabi.check(ctx.lib.b200geo_ransac_finish_dev(ctx.handle, _p(pts), pts.shape[0], _p(idx), _p(keys),
                                                  float(np.float32(th)), _cabi.METHODS[method], _p(result), _p(mask),
                                                  _stream(pts)))
    return result, mask


def shard_range(K: int, rank: int, world: int):
    """Contiguous hypothesis block of `rank` (blocks differ by at most one sample set)."""
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ransac(pts, idx, th, d, n=4, method="fwd", group=None, keys=None, result=None, mask=None, counts=None):
    """RANSAC.run over K explicit sample sets, hypothesis-block sharded over the ranks of `group`
    (all ranks hold the same pts and idx).  Returns (result, mask) on every rank; see ransac_finish.
    With torch.distributed uninitialised (or world size 1) no collective is issued."""
    import torch.distributed as dist
    M, K = pts.shape[0], idx.shape[0]
    thr = M * d / 100 + n
    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(K, rank, world)
    if world > 1 and _peer_ready.get(pts.device.index) == world and counts is None:
        # peer-memory exchange set up (peer_setup): the winner records travel by peer stores, no collective at all
        return ransac_sharded(pts, idx[lo:hi], lo, th, d, n, method, group=group, result=result, mask=mask)
    keys = ransac_search(pts, idx, lo, hi, th, thr, method, keys=keys, counts=counts)
    if world > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)     # 16 bytes over NVLink
    return ransac_finish(pts, idx, keys, th, method, result=result, mask=mask)


WINREC_BYTES = 48      # struct WinRec (csrc/select_refit.cuh): keys[2] + the sample indices each key names


def peer_setup(group=None, device=None) -> bool:
    """Open the peer-memory record exchange among the ranks of `group` (one process per GPU of ONE node): every rank allocates
    an inbox in its HBM, the 64-byte CUDA IPC handles are all-gathered once, and every rank maps the others' inboxes
    (b200geo_peer_export / _open).  `ransac_sharded` then exchanges its 48-byte winner records by peer stores over NVLink and a
    flag wait instead of an NCCL all-gather.  Returns False (and leaves the collective path in place) if any rank cannot
    export or map -- e.g. GPUs without peer access."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.cuda.current_device() if device is None else device
    if world < 2 or world > 64:
        return False
    ctx = _cabi.context(dev)
    h = np.zeros(64, np.uint8)
    ok = ctx.lib.b200geo_peer_export(ctx.handle, world, _cabi.ptr(h)) == _cabi.OK
    mine = torch.from_numpy(np.concatenate([h, np.array([1 if ok else 0], np.uint8)])).to(f"cuda:{dev}")
    allh = torch.empty((world, 65), dtype=torch.uint8, device=f"cuda:{dev}")
    dist.all_gather_into_tensor(allh, mine.view(1, 65), group=group)
    allh = allh.cpu().numpy()
    ok = bool(allh[:, 64].all())
    if ok:
        ok = ctx.lib.b200geo_peer_open(ctx.handle, rank, world, _cabi.ptr(np.ascontiguousarray(allh[:, :64]))) == _cabi.OK
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{dev}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)          # also the barrier before the first peer store
    ok = bool(flag.item())
    if not ok:
        ctx.lib.b200geo_peer_close(ctx.handle)
        _peer_ready.pop(dev, None)
    else:
        _peer_ready[dev] = world
    return ok


def peer_close(device=None):
    dev = torch.cuda.current_device() if device is None else device
    if _peer_ready.pop(dev, None) is not None:
        ctx = _cabi.context(dev)
        ctx.lib.b200geo_peer_close(ctx.handle)


def ransac_sharded(pts, idx_shard, k_offset, th, d, n=4, method="fwd", group=None, rec=None, recs=None, result=None, mask=None,
                   counts=None):
    """RANSAC.run over K sample sets of which THIS rank holds only rows k_offset .. k_offset + len(idx_shard) - 1 (its own
    hypothesis block; `pts` is replicated).  Each rank reduces its block to a 48-byte record -- its two keys and the four
    correspondence indices of the sample sets they name --, ONE exchange of those records (48 bytes per rank over NVLink: peer
    stores + flag wait after `peer_setup`, else an NCCL all-gather) replaces the max all-reduce of `ransac`, and every rank finishes from the gathered records (no broadcast: the winner's
    indices travel in the record).  Same result as `ransac` on the full index array.  Returns (result, mask)."""
    import torch.distributed as dist
    _chk_pts(pts)
    assert idx_shard.dtype == torch.int32 and idx_shard.is_contiguous() and idx_shard.shape[-1] == 4
    ctx = _ctx(pts)
    M = pts.shape[0]
    thr = M * d / 100 + n
    world = 1
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
    if rec is None:
        rec = torch.empty((WINREC_BYTES,), dtype=torch.uint8, device=pts.device)
    _cabi.check(ctx.lib.b200geo_ransac_search_shard_dev(ctx.handle, _p(pts), M, _p(idx_shard), int(k_offset), int(idx_shard.shape[0]),
                                                        float(np.float32(th)), float(thr), _cabi.METHODS[method], _p(rec), _p(counts),
                                                        _stream(pts)))
    if world > 1:
        if recs is None:
            recs = torch.empty((world * WINREC_BYTES,), dtype=torch.uint8, device=pts.device)
        if _peer_ready.get(pts.device.index) == world:
            # records by peer stores into every rank's inbox + a flag wait (two single-block kernels), no collective
            _cabi.check(ctx.lib.b200geo_ransac_exchange_peer_dev(ctx.handle, _p(rec), _p(recs), _stream(pts)))
        else:
            dist.all_gather_into_tensor(recs, rec, group=group)
    else:
        recs = rec
    if result is None:
        result = torch.empty((32,), dtype=torch.float64, device=pts.device)
    if mask is None:
        mask = torch.empty((M,), dtype=torch.uint8, device=pts.device)
    _cabi.check(ctx.lib.b200geo_ransac_finish_recs_dev(ctx.handle, _p(pts), M, _p(recs), world, float(np.float32(th)),
                                                       _cabi.METHODS[method], _p(result), _p(mask), _stream(pts)))
    return result, mask


def ransac_batched(pts, idx, th, d, n=4, method="fwd", results=None, masks=None):
    """B independent pairs: pts (B,M,4), idx (B,K,4) -> results (B,32) float64, masks (B,M) uint8."""
    _chk_pts(pts)
    assert pts.dim() == 3 and idx.dim() == 3 and idx.dtype == torch.int32 and idx.is_contiguous()
    B, M, _ = pts.shape
    K = idx.shape[1]
    ctx = _ctx(pts)
    if results is None:
        results = torch.empty((B, 32), dtype=torch.float64, device=pts.device)
    if masks is None:
        masks = torch.empty((B, M), dtype=torch.uint8, device=pts.device)
    _cabi.check(ctx.lib.b200geo_ransac_batched_dev(ctx.handle, _p(pts), M, _p(idx), K, B, float(np.float32(th)),
                                                   float(M * d / 100 + n), _cabi.METHODS[method], _p(results), _p(masks),
                                                   _stream(pts)))
    return results, masks


# ------------------------------------------------------------------------------------------------
# batched warps
# ------------------------------------------------------------------------------------------------
def scan_jobs(H: np.ndarray, src_hw, res_hw, channels=3, src_itemsize=1, dst_itemsize=1) -> np.ndarray:
    """Jobs for wrapPerspectiveScan semantics (reference homography.py:186-209) on a batch of
    equally sized frames stored back to back: H (B,3,3) float64; output res = (h, w) per frame."""
    H = np.asarray(H, np.float64).reshape(-1, 3, 3)
    B = H.shape[0]
    ih, iw = src_hw
    oh, ow = res_hw
    jobs = np.zeros(B, _cabi.WARP_JOB_DTYPE)
    jobs["invH"] = np.linalg.inv(H).reshape(B, 9)
    jobs["x0"], jobs["dx"], jobs["x_last"] = 0.0, (ow / (ow - 1) if ow > 1 else 0.0), (float(ow) if ow > 1 else 0.0)
    jobs["y0"], jobs["dy"], jobs["y_last"] = 0.0, (oh / (oh - 1) if oh > 1 else 0.0), (float(oh) if oh > 1 else 0.0)
    jobs["ow"], jobs["oh"] = ow, oh
    jobs["bw"], jobs["bh"] = ow, oh            # the reference passes `res` as the sampler's bounds
    jobs["iw"], jobs["ih"] = iw, ih
    sp, dp = iw * channels * src_itemsize, ow * channels * dst_itemsize
    jobs["src_pitch"], jobs["dst_pitch"] = sp, dp
    jobs["src_offset"] = np.arange(B, dtype=np.int64) * (sp * ih)
    jobs["dst_offset"] = np.arange(B, dtype=np.int64) * (dp * oh)
    return jobs


def bbox_jobs(H: np.ndarray, src_hw, channels=3, src_itemsize=1, dst_itemsize=1, canvas_hw=None):
    """Jobs for wrapPerspective semantics (reference homography.py:142-179): per-frame bounding box
    of the warped corners.  Outputs are packed back to back unless canvas_hw is given (each output
    then occupies a canvas_hw slot and is clipped to it).  Returns (jobs, (min_x, min_y, w, h) per frame)."""
    H = np.asarray(H, np.float64).reshape(-1, 3, 3)
    B = H.shape[0]
    ih, iw = src_hw
    bnd = np.array([[0, iw - 1, iw - 1, 0], [0, 0, ih - 1, ih - 1], [1.0, 1, 1, 1]])
    q = H @ bnd
    q = q / q[:, 2:3, :]
    bad = ~np.isfinite(q[:, :2]).all(axis=(1, 2)) | (np.abs(q[:, :2]) > 1e9).any(axis=(1, 2))
    if bad.any():                                     # H finite but a corner at w ~ 0: the reference's int() raises here
        raise ValueError(f"bbox_jobs: projected corners are not finite for frames {np.nonzero(bad)[0].tolist()}")
    trunc = lambda a: np.trunc(a).astype(np.int64)   # int(): toward zero
    min_x, max_x = trunc(q[:, 0].min(1)), trunc(q[:, 0].max(1))
    min_y, max_y = trunc(q[:, 1].min(1)), trunc(q[:, 1].max(1))
    ow, oh = max_x - min_x + 1, max_y - min_y + 1
    if canvas_hw is not None:
        ow, oh = np.minimum(ow, canvas_hw[1]), np.minimum(oh, canvas_hw[0])
    jobs = np.zeros(B, _cabi.WARP_JOB_DTYPE)
    jobs["invH"] = np.linalg.inv(H).reshape(B, 9)
    jobs["x0"], jobs["dx"], jobs["x_last"] = min_x, 1.0, min_x + ow - 1
    jobs["y0"], jobs["dy"], jobs["y_last"] = min_y, 1.0, min_y + oh - 1
    jobs["ow"], jobs["oh"] = ow, oh
    jobs["bw"], jobs["bh"] = iw, ih
    jobs["iw"], jobs["ih"] = iw, ih
    sp = iw * channels * src_itemsize
    jobs["src_pitch"] = sp
    jobs["src_offset"] = np.arange(B, dtype=np.int64) * (sp * ih)
    if canvas_hw is not None:
        dp = canvas_hw[1] * channels * dst_itemsize
        jobs["dst_pitch"] = dp
        jobs["dst_offset"] = np.arange(B, dtype=np.int64) * (dp * canvas_hw[0])
    else:
        dpitch = ow * channels * dst_itemsize
        jobs["dst_pitch"] = dpitch
        sizes = dpitch * oh
        jobs["dst_offset"] = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    return jobs, np.stack([min_x, min_y, ow, oh], axis=1)


_TORCH_TYPE = {torch.uint8: _cabi.U8, torch.float32: _cabi.F32, torch.float64: _cabi.F64}


def warp(src: torch.Tensor, dst: torch.Tensor, jobs: np.ndarray, channels: int = 3, sampler: str = "bilinear"):
    """K4 over a batch: src / dst are device buffers the jobs' offsets index into."""
    assert src.is_cuda and dst.is_cuda and jobs.dtype == _cabi.WARP_JOB_DTYPE
    ctx = _ctx(src)
    jobs = np.ascontiguousarray(jobs)
    _cabi.check(ctx.lib.b200geo_warp_dev(ctx.handle, _p(src), _p(dst), _cabi._vp(jobs.ctypes.data), jobs.shape[0],
                                         channels, _TORCH_TYPE[src.dtype], _TORCH_TYPE[dst.dtype],
                                         _cabi.SAMPLERS[sampler], _stream(src)))
    return dst


def warp_host(src, dst, jobs: np.ndarray, channels: int = 3, sampler: str = "bilinear", chunk_jobs: int = 0, device: int | None = None):
    """K4 over a batch of images in HOST memory (numpy arrays or CPU tensors; pinned memory gives
    full overlap): chunks of jobs flow through a host->device / warp / device->host pipeline on
    three streams (b200geo_warp_batch_host).  jobs' offsets index the host buffers.  Synchronous."""
    assert jobs.dtype == _cabi.WARP_JOB_DTYPE
    ctx = _cabi.context(device)
    jobs = np.ascontiguousarray(jobs)

    def _hp(a):
        if isinstance(a, torch.Tensor):
            assert not a.is_cuda and a.is_contiguous()
            return _cabi._vp(a.data_ptr()), _TORCH_TYPE[a.dtype]
        assert a.flags.c_contiguous
        return _cabi._vp(a.ctypes.data), {np.dtype(np.uint8): _cabi.U8, np.dtype(np.float32): _cabi.F32, np.dtype(np.float64): _cabi.F64}[a.dtype]
    sp, st = _hp(src)
    dp, dt = _hp(dst)
    _cabi.check(ctx.lib.b200geo_warp_batch_host(ctx.handle, sp, dp, _cabi._vp(jobs.ctypes.data), jobs.shape[0], channels, st, dt,
                                                _cabi.SAMPLERS[sampler], int(chunk_jobs)))
    return dst


def stitch_jobs(results: torch.Tensor, src_hw, canvas_hw, jobs: torch.Tensor | None = None, boxes: torch.Tensor | None = None):
    """Warp jobs of B pairs from their RANSAC results, built on the device (b200geo_stitch_jobs_dev): wrapPerspective's
    output rectangle (reference homography.py:143-163) and inv(H) per pair, outputs clipped to canvas slots.
    Returns (jobs (B*184,) uint8 device, boxes (B,5) int32 device = min_x, min_y, out_w, out_h, ok)."""
    B = results.shape[0]
    ctx = _ctx(results)
    assert results.dtype == torch.float64 and results.is_contiguous() and results.shape[1] == 32
    if jobs is None:
        jobs = torch.empty((B * _cabi.WARP_JOB_DTYPE.itemsize,), dtype=torch.uint8, device=results.device)
    if boxes is None:
        boxes = torch.empty((B, 5), dtype=torch.int32, device=results.device)
    _cabi.check(ctx.lib.b200geo_stitch_jobs_dev(ctx.handle, _p(results), B, int(src_hw[0]), int(src_hw[1]), int(canvas_hw[0]), int(canvas_hw[1]),
                                                _p(jobs), _p(boxes), _stream(results)))
    return jobs, boxes


def stitch_batched(pts, idx, frames, th, d, canvas_hw, n=4, method="fwd", sampler="bilinear", canvases=None, jobs=None, boxes=None):
    """Batched stitch pipeline (config "1024 pairs x 2k correspondences x 8192 hypotheses + warp to a
    4K canvas"): for every pair b, RANSAC.run over pts[b] with the sample sets idx[b] (reference
    ransac.py:159-213), then transformImageH(frames[b], H_b) (reference homography.py:230-242) into
    canvas slot b, clipped to canvas_hw.  pts (B,M,4) f32, idx (B,K,4) i32, frames (B,h,w,3) u8, all
    on the device.  Nothing returns to the host in between: the output rectangles (bounding box of the
    warped corners, int() truncation) and inv(H) are computed by a kernel from the refit results.
    Returns (canvases (B,H,W,3) u8, boxes (B,5) int32 device = min_x, min_y, out_w, out_h, ok,
    results (B,32) f64, masks (B,M) u8).  Pairs with ok = 0 (failed search / degenerate H) get the identity warp."""
    B, h, w, c = frames.shape
    assert c == 3 and frames.dtype == torch.uint8 and frames.is_contiguous()
    results, masks = ransac_batched(pts, idx, th, d, n, method)
    jobs, boxes = stitch_jobs(results, (h, w), canvas_hw, jobs, boxes)
    if canvases is None:
        canvases = torch.zeros((B, canvas_hw[0], canvas_hw[1], 3), dtype=torch.uint8, device=frames.device)
    # the frames are one 3-D byte array (row bytes x rows x frames): each strip's / tile's footprint arrives by one tensor copy.
    # The launch is sized for the largest rectangle a job can have: the canvas slot.
    warp_resident(frames, canvases, jobs, B, min(canvas_hw[1], 1 << 16), min(canvas_hw[0], 1 << 16), 3, sampler, layout=(w * 3, B * h, h, w * 3))
    return canvases, boxes, results, masks


def stitch_compose(imgQ: torch.Tensor, imgT: torch.Tensor, H, blending=False, blendrate=0.2, canvas: torch.Tensor | None = None):
    """stitchPanorama (reference homography.py:288-338) on device-resident uint8 RGB images: warp of imgT by H,
    canvas placement, paste (blending False) or 'Rate' alpha blend, uint8 canvas -- one kernel (K5), asynchronous on
    the current stream.  Unlike the host entry point it does not write to imgT: the kernel itself reads source
    pixel (0,0) as 0.  Returns the canvas tensor (fh, fw, 3)."""
    from . import homography as hg
    assert imgQ.dtype == torch.uint8 and imgT.dtype == torch.uint8 and imgQ.is_contiguous() and imgT.is_contiguous()
    assert imgQ.ndim == 3 and imgT.ndim == 3 and imgQ.shape[2] == 3 and imgT.shape[2] == 3
    assert not blending or blending == 'Rate'
    H = np.asarray(H, dtype=np.float64)
    h, w, _ = imgT.shape
    hq, wq, _ = imgQ.shape
    min_x, max_x, min_y, max_y = hg._bbox(np.empty((h, w, 0)), H, 0)
    wt, ht = max_x - min_x + 1, max_y - min_y + 1
    (tsx, tsy, _, _), (qsx, qsy, _, _), fw, fh = hg._placement(min_x, min_y, wt, ht, wq, hq)
    if canvas is None:
        canvas = torch.empty((fh, fw, 3), dtype=torch.uint8, device=imgT.device)
    assert canvas.shape == (fh, fw, 3) and canvas.dtype == torch.uint8 and canvas.is_contiguous()
    job = _cabi.WarpJob()
    job.invH[:] = np.linalg.inv(H).reshape(9).tolist()
    job.x0, job.dx, job.x_last = hg._linspace_params(min_x, max_x, wt)
    job.y0, job.dy, job.y_last = hg._linspace_params(min_y, max_y, ht)
    job.ow, job.oh, job.bw, job.bh, job.iw, job.ih = wt, ht, w, h, w, h
    job.src_offset, job.src_pitch, job.dst_offset, job.dst_pitch = 0, w * 3, 0, 0
    p = _cabi.Compose()
    p.fw, p.fh, p.tsx, p.tsy, p.qsx, p.qsy, p.qw, p.qh = fw, fh, tsx, tsy, qsx, qsy, wq, hq
    p.q_pitch, p.canvas_pitch = wq * 3, fw * 3
    p.alpha_t, p.alpha_q = float(np.float32(blendrate + 1e-10)), float(np.float32(1 + 1e-10 - blendrate))
    p.alpha_bg = float(np.float32(1e-10))
    p.blend = 1 if blending else 0
    import ctypes as C
    ctx = _ctx(imgT)
    _cabi.check(ctx.lib.b200geo_stitch_compose_dev(ctx.handle, _p(imgT), _p(imgQ), _p(canvas), C.byref(job), C.byref(p), _stream(imgT)))
    return canvas


def upload_jobs(jobs: np.ndarray, device) -> torch.Tensor:
    """Copy a job array to the device once (for repeated launches with b200geo_warp_jobs_dev)."""
    return torch.from_numpy(np.ascontiguousarray(jobs).view(np.uint8).copy()).to(device)


def warp_resident(src, dst, jobs_dev: torch.Tensor, n_jobs: int, max_ow: int, max_oh: int, channels=3, sampler="bilinear",
                  layout=None):
    """K4 with the job array already on the device: no host work per launch.  layout = (pitch_bytes,
    rows[, image_rows, row_bytes]) describes `src` as one 2-D byte array (frames stored back to back:
    rows = n_frames * height) and lets the kernel fetch each tile's source footprint with one TMA tensor
    copy; with image_rows / row_bytes (equal frames of image_rows rows and row_bytes = 3 * width bytes)
    footprints over a frame edge are fetched the same way, zero-filled by the hardware."""
    ctx = _ctx(src)
    lay = None
    if layout is not None:
        lay = _cabi.SrcLayout(int(layout[0]), int(layout[1]), int(layout[2]) if len(layout) > 2 else 0, int(layout[3]) if len(layout) > 3 else 0)
    import ctypes as C
    _cabi.check(ctx.lib.b200geo_warp_batch_dev(ctx.handle, _p(src), _p(dst), _p(jobs_dev), int(n_jobs), int(max_ow),
                                               int(max_oh), channels, _TORCH_TYPE[src.dtype], _TORCH_TYPE[dst.dtype],
                                               _cabi.SAMPLERS[sampler], C.byref(lay) if lay is not None else None,
                                               _stream(src)))
    return dst
