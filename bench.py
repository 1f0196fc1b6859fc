#!/usr/bin/env python
"""bench.py -- headline benchmark of the geometric hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|ransac|warp]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Prints ONE JSON line (rank 0).  BASELINE.json's metric has two halves, so the line has two records:

* top level  -- RANSAC throughput in hypothesis x correspondences per second on BASELINE config 3
  ("synthetic single-pair RANSAC: 20k correspondences, 50% outliers, 1M hypotheses", hypothesis-block
  sharded over the ranks, one 16-byte NCCL max all-reduce per step);
* "warp"     -- backward-warp throughput in Gpix/s on BASELINE config 4 ("4096 synthetic 1920x1080
  RGB frames with random homographies", frame-sharded, no communication).

A step = one pass of the hot path over the whole workload.  `value` is timed with CUDA events with
the inputs resident in HBM; `e2e` is the same metric through the host-buffer entry points (pinned
host memory -> device -> kernels -> host) with the copies inside the timed region.
`--impl reference` times the UNMODIFIED reference (oracle/_ref: ransac.py / homography.py copied by
oracle/make_ref.py; falls back to the numpy port oracle/ref_numpy.py when the copy is absent) on all
host cores, on a bounded sample of the same workloads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 25.0          # algorithmic FLOP per hypothesis x correspondence, fwd loss
# DRAM traffic of the dominant kernels from the committed ncu --set full captures (dram__bytes_read + write):
#   warp_strip_kernel, profiles/r02_strip_v6.ncu-rep: 442.2 MB + 361.0 MB for 64 frames 1080p = 6.05 B per output pixel
#   score_fast_kernel, profiles/r02_score_fast.ncu-rep: 41.5 MB + 3.9 MB for 1M x 20k pairs (FP32-bound: traffic is incidental)
# (constants of those captures scaled by the launch's units, not per-run measurements: re-captured whenever a kernel changes)
WARP_TRAFFIC_B_PER_PIXEL = (442.207744e6 + 361.041408e6) / (64 * 1080 * 1920)
SCORE_TRAFFIC_B_PER_PAIR = (41.497088e6 + 3.899392e6) / (1000000 * 20000)
C3 = dict(M=20000, K=1_000_000, th=3.0, d=99.0, seed=1, idx_seed=0)
C4 = dict(frames=4096, h=1080, w=1920, jitter=0.05)


# ------------------------------------------------------------------------------------------------
# workload generators
# ------------------------------------------------------------------------------------------------
def synth_pair(seed, M, outlier_frac=0.5, noise=0.5):
    rs = np.random.RandomState(seed)
    X = (rs.rand(2, M) * np.array([[1920.0], [1080.0]])).astype(np.float32)
    Ht = np.array([[0.9, 0.05, 30], [-0.04, 1.1, -20], [1e-5, -2e-5, 1]])
    q = Ht @ np.vstack([X.astype(np.float64), np.ones(M)])
    Y = q[:2] / q[2] + rs.randn(2, M) * noise
    out = rs.rand(M) < outlier_frac
    Y[:, out] = rs.rand(2, int(out.sum())) * np.array([[1920.0], [1080.0]])
    return X, Y.astype(np.float32)


def frame_homographies(n, h, w, jitter, seed=0):
    """Per-frame H from 4-corner jitter U(+-jitter of W/H), solved in float64 (plain 8x8 solve)."""
    rs = np.random.RandomState(seed)
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
    Hs = np.empty((n, 3, 3))
    for i in range(n):
        dst = src + rs.uniform(-jitter, jitter, (4, 2)) * [w, h]
        A = np.zeros((8, 8))
        b = np.zeros(8)
        for j, ((x, y), (X, Y)) in enumerate(zip(src, dst)):
            A[2 * j] = [x, y, 1, 0, 0, 0, -x * X, -y * X]
            A[2 * j + 1] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
            b[2 * j], b[2 * j + 1] = X, Y
        Hs[i] = np.append(np.linalg.solve(A, b), 1).reshape(3, 3)
    return Hs


C5 = dict(pairs=1024, M=2000, K=8192, th=4.0, d=95.0, h=1080, w=1920, canvas=(2160, 3840))


def stitch_pairs(p0, p1):
    """C5 pairs p0..p1-1: per pair a near-identity homography (corner jitter +-10%, translation within
    +-900 / +-500 px), M correspondences with 50% outliers and 0.5 px noise, K explicit sample sets."""
    M, K, h, w = C5["M"], C5["K"], C5["h"], C5["w"]
    n = p1 - p0
    pts = np.empty((n, M, 4), np.float32)
    idx = np.empty((n, K, 4), np.int32)
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
    for i, p in enumerate(range(p0, p1)):
        rs = np.random.RandomState(1000 + p)
        dst = src + rs.uniform(-0.1, 0.1, (4, 2)) * [w, h] + rs.uniform(-1, 1, 2) * [900, 500]
        A = np.zeros((8, 8)); b = np.zeros(8)
        for j, ((x, y), (X, Y)) in enumerate(zip(src, dst)):
            A[2 * j] = [x, y, 1, 0, 0, 0, -x * X, -y * X]; A[2 * j + 1] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
            b[2 * j], b[2 * j + 1] = X, Y
        Ht = np.append(np.linalg.solve(A, b), 1).reshape(3, 3)
        X = rs.rand(2, M) * np.array([[w], [h]], np.float64)
        q = Ht @ np.vstack([X, np.ones(M)])
        Y = q[:2] / q[2] + rs.randn(2, M) * 0.5
        out = rs.rand(M) < 0.5
        Y[:, out] = rs.rand(2, int(out.sum())) * np.array([[w], [h]])
        pts[i, :, :2], pts[i, :, 2:] = X.T, Y.T
        idx[i] = np.random.RandomState(p).randint(0, M, (K, 4))
    return pts, idx


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "display_clocks": 0x100}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join(1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def merge_clocks(a, b):
    out = dict(a)
    out["reasons"] = sorted(set(a.get("reasons", [])) | set(b.get("reasons", [])))
    if a.get("sm_mhz") and b.get("sm_mhz"):
        out["sm_mhz"] = min(a["sm_mhz"], b["sm_mhz"])
    return out


def numa_bind(local_gpu):
    """Run this rank on the CPUs NVML names as local to its GPU, so that pinned host buffers allocated afterwards are
    NUMA-local to the GPU's PCIe root.  Returns a short description for the record."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {local_gpu} (nvml)"
    except Exception as e:          # noqa: BLE001 -- a missing NVML call must not stop the benchmark
        return f"not bound ({type(e).__name__})"
    return "not bound"


def measure_link(torch, dist, world, mb=256, reps=3):
    """Host link of this box as the e2e paths use it: pinned host <-> device copies of `mb` MB, each direction alone and
    both at once (two streams), on all ranks at the same time.  GB/s per rank: the MIN over the ranks (what the slowest rank
    sees while the others copy too); `sum_*` adds the ranks up."""
    n = mb << 20
    hin, hout = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    din, dout = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t0) / 1e9
    run(True, True)
    vals = [run(True, False), run(False, True), run(True, True)]
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    tmin, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    a, b = tmin.cpu().tolist(), tsum.cpu().tolist()
    return {"h2d_alone_gbs": a[0], "d2h_alone_gbs": a[1], "each_direction_concurrent_gbs": a[2], "sum_ranks_each_direction_concurrent_gbs": b[2],
            "how": f"{mb} MB pinned <-> device copies, all {world} rank(s) at once; per-rank figures are the minimum over ranks"}


# ------------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------------
def _reference_modules():
    """(ransac, homography) of the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py) or None."""
    import oracle
    return oracle.reference_modules()


def _cpu_ransac_block(args):
    """One worker: hypotheses of config 3 on the host.  kind 'reference': RANSAC.run of the unmodified reference
    (reference ransac.py:159-213: it draws its own sample sets from the seeded global RNG; d = 99 -> no early exit, all k
    hypotheses are solved and scored, then the refit).  kind 'port': the numpy restatement on explicit sample sets."""
    kind, X, Y, idx_or_k, th, seed = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import io
    import warnings
    from contextlib import redirect_stdout
    warnings.filterwarnings("ignore")
    if kind == "reference":
        RR, _ = _reference_modules()
        np.random.seed(seed)
        model = RR.HomoModel(th=th, d=C3["d"], n=4)
        t = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            RR.RANSAC(model, k=int(idx_or_k)).run([X, Y], method="fwd")
        return time.perf_counter() - t
    from oracle import ref_numpy as port
    t = time.perf_counter()
    port.score_block(X, Y, idx_or_k, th, "fwd")
    return time.perf_counter() - t


def _cpu_warp_block(args):
    kind, frames, Hs, h, w = args
    t = time.perf_counter()
    if kind == "reference":
        _, RH = _reference_modules()
        for f, H in zip(frames, Hs):
            RH.wrapPerspectiveScan(f, H, [h, w], "bilinear")[0].astype(np.uint8)      # reference homography.py:186-209 + :225
    else:
        from oracle import ref_numpy as port
        for f, H in zip(frames, Hs):
            port.warp_scan(f, H, [h, w], True)[0].astype(np.uint8)
    return time.perf_counter() - t


def cpu_kind():
    return "reference" if _reference_modules() is not None else "port"


def cpu_ransac(n_hyp, cores, pool=None, kind=None):
    """RANSAC.run's loop on a K = n_hyp sample of config 3 on `cores` host processes; returns hyp*pts/s."""
    import warnings
    warnings.filterwarnings("ignore")
    kind = kind or cpu_kind()
    X, Y = synth_pair(C3["seed"], C3["M"])
    idx = np.random.RandomState(C3["idx_seed"]).randint(0, C3["M"], (n_hyp, 4))
    t0 = time.perf_counter()
    if cores <= 1:
        _cpu_ransac_block((kind, X, Y, n_hyp if kind == "reference" else idx, C3["th"], 0))
    else:
        blocks = np.array_split(idx, cores)
        pool.map(_cpu_ransac_block, [(kind, X, Y, len(b) if kind == "reference" else b, C3["th"], i) for i, b in enumerate(blocks)])
    dt = time.perf_counter() - t0
    return n_hyp * C3["M"] / dt, dt


def cpu_warp(n_frames, cores, kind=None):
    """wrapPerspectiveScan + bilinear + astype(uint8) on n_frames 1080p frames on `cores` host processes; Gpix/s."""
    kind = kind or cpu_kind()
    h, w = C4["h"], C4["w"]
    rs = np.random.RandomState(0)
    frames = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for _ in range(n_frames)]
    Hs = frame_homographies(n_frames, h, w, C4["jitter"], seed=0)
    t0 = time.perf_counter()
    if cores <= 1:
        _cpu_warp_block((kind, frames, Hs, h, w))
    else:
        import multiprocessing as mp
        parts = [(kind, frames[i::cores], Hs[i::cores], h, w) for i in range(cores) if frames[i::cores]]
        with mp.get_context("fork").Pool(len(parts)) as pool:
            pool.map(_cpu_warp_block, parts)
    dt = time.perf_counter() - t0
    return n_frames * h * w / dt / 1e9, dt


C3_WORKLOAD = ("C3 synthetic single-pair RANSAC: M=20000 correspondences (50% outliers), K=1000000 sample sets, "
               "fwd loss, th=3, d=99 (no early exit); solve + score + select + mask + refit per step")
C4_WORKLOAD = ("C4 batched backward warp: 4096 synthetic 1920x1080 RGB uint8 frames, per-frame random homography "
               "(corner jitter +-5%), bilinear, fixed 1080x1920 output (wrapPerspectiveScan grid)")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, bounded samples of the
    same workloads.  With oracle/_ref present (the unmodified reference, copied by oracle/make_ref.py) it is the reference
    itself -- RANSAC.run / wrapPerspectiveScan -- one process per core; otherwise the numpy port (oracle/ref_numpy.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    steps = max(1, args.steps)
    n_hyp = cores * 2500                          # ~0.2 ms per hypothesis per core at M = 20k -> ~0.5 s per step
    import multiprocessing as mp
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
    for _ in range(min(args.warmup, 2)):
        cpu_ransac(cores * 100, cores, pool, kind)
    vals, times = [], []
    for _ in range(steps):
        v, dt = cpu_ransac(n_hyp, cores, pool, kind)
        vals.append(v)
        times.append(dt)
    if pool is not None:
        pool.close()
    value = float(np.mean(vals))
    what = ("RANSAC.run of the unmodified reference (oracle/_ref/ransac.py:159-213), sample sets drawn by the reference from the seeded global RNG"
            if kind == "reference" else "numpy port of ransac.py:176-184 (4-point SVD solve + scoring)")
    line = {
        "impl": "reference", "metric": "ransac_hyp_pts_per_s", "value": value, "unit": "hyp*pts/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C3_WORKLOAD, "sample": f"{n_hyp} of 1000000 sample sets per step", "sharding": f"one process per host core ({cores})"},
        "cpu_baseline": {"value": value, "unit": "hyp*pts/s", "cores": cores, "kind": kind,
                         "sample": f"{n_hyp} of 1000000 sample sets x 20000 correspondences per step, {what}, one process per core"},
        "e2e": {"value": value, "unit": "hyp*pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload in ("all", "warp"):
        n_frames = max(cores, 8)
        wv, wdt = cpu_warp(n_frames, cores, kind)
        wwhat = ("wrapPerspectiveScan(..., 'bilinear').astype(uint8) of the unmodified reference (oracle/_ref/homography.py:186-209,123-138)"
                 if kind == "reference" else "numpy port of homography.py:186-209,123-138")
        line["warp"] = {"metric": "warp_gpix_per_s", "value": wv, "unit": "Gpix/s", "ms_per_step": wdt * 1e3,
                        "config": {"workload": C4_WORKLOAD, "sample": f"{n_frames} of 4096 frames"},
                        "cpu_baseline": {"value": wv, "unit": "Gpix/s", "cores": cores, "kind": kind, "sample": f"{n_frames} frames, {wwhat}"},
                        "e2e": {"value": wv, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def timed_steps(torch, dist, world, fn, steps, warmup, flush=None, per_step=None):
    """W warm-up steps, then K timed steps (CUDA events on the current stream; optional L2 flush
    between steps, outside the events); barrier + synchronize on both sides; max over ranks.
    Returns (total ms over the K steps as max over ranks)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    if per_step is not None:
        per_step.extend(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_ransac(torch, dist, rank, world, args, ctx, peaks, local_gpu):
    from ransac_with_homography_b200 import device
    from ransac_with_homography_b200.ransac import HomoModel, RANSAC
    M, K, th, d = C3["M"], C3["K"], C3["th"], C3["d"]
    X, Y = synth_pair(C3["seed"], M)
    pts_h = np.empty((M, 4), np.float32)
    pts_h[:, :2], pts_h[:, 2:] = X.T, Y.T
    idx_h = np.random.RandomState(C3["idx_seed"]).randint(0, M, (K, 4)).astype(np.int32)
    pts, idx = torch.from_numpy(pts_h).cuda(), torch.from_numpy(idx_h).cuda()
    keys = torch.zeros(2, dtype=torch.int64, device="cuda")
    result = torch.zeros(32, dtype=torch.float64, device="cuda")
    mask = torch.zeros(M, dtype=torch.uint8, device="cuda")
    l2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
    flush = lambda: l2.fill_(1)
    group = None
    # N > 1: the per-rank winner records are exchanged by peer stores over NVLink / NVSwitch + a flag wait inside two single-block
    # kernels (device.peer_setup: CUDA IPC inboxes); if the inboxes cannot be mapped the NCCL collectives stay in place
    peer = device.peer_setup(group) if world > 1 and not os.environ.get("B200GEO_NO_PEER") else False

    def step():
        device.ransac(pts, idx, th, d, 4, "fwd", group=group, keys=keys, result=result, mask=mask)

    l0 = ctx.launch_count()
    with ClockSampler(local_gpu) as cs:
        ms = timed_steps(torch, dist, world, step, args.steps, args.warmup, flush)
    launches = (ctx.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    value = K * M * args.steps / (ms / 1e3)
    res = result.cpu().numpy()

    # dominant kernel alone (K2 scoring of this rank's shard), CUDA events on the launching stream
    lo, hi = device.shard_range(K, rank, world)
    H, _, flags = device.dlt4(pts, idx[lo:hi].contiguous())
    counts = torch.zeros(hi - lo, dtype=torch.int32, device="cuda")
    kms = timed_steps(torch, dist, 1, lambda: device.score(pts, H, th, "fwd", flags=flags, counts=counts), args.steps, args.warmup, flush)
    k_avg_s = kms / args.steps / 1e3
    flops = FLOP_PER_PAIR * (hi - lo) * M
    peak_tflops = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    clocks = cs.summary()
    roof = {"bound": "fp32", "kernel": "score (K2)", "achieved": flops / k_avg_s / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": flops / k_avg_s / 1e12 / peak_tflops, "traffic": SCORE_TRAFFIC_B_PER_PAIR * (hi - lo) * M,
            "traffic_source": "ncu dram bytes per pair (profiles/r02_score_fast.ncu-rep) x pairs of this launch",
            "peak_source": f"148 SM x 128 FP32 lanes x 2 x sm_max_mhz ({peaks['sm_max_mhz']:.0f} MHz, MEASURED_PEAKS.json: {peaks['source']})",
            "algorithmic_flop_per_hyp_pt": FLOP_PER_PAIR, "kernel_ms": k_avg_s * 1e3,
            "kernel_share_of_step": (k_avg_s * 1e3) / (ms / args.steps)}

    # end to end: host buffers in, host results out, copies inside the timed region
    h2d = pts_h.nbytes + idx_h.nbytes
    d2h = 32 * 8 + M
    if world == 1:
        model = HomoModel(th=th, d=d, n=4)
        runner = RANSAC(model, k=K)
        import io
        from contextlib import redirect_stdout

        def e2e_step():
            with redirect_stdout(io.StringIO()):
                runner.run([X, Y], method="fwd", idx=idx_h)
        d2h += 4                           # the last sample set's count, for the reference's closing warning (ransac.py:203-204)
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        e2e_how = "ransac.RANSAC.run([X, Y], 'fwd', idx=...) on numpy arrays (b200geo_ransac_host: the sample sets past the first scoring wave are uploaded under that wave's scoring)"
    else:
        # every rank uploads the correspondences and ONLY ITS OWN shard of the sample sets; the winner's four indices reach the
        # other ranks inside the 48-byte record that is all-gathered instead of the 16-byte max all-reduce (device.ransac_sharded)
        pin_p, pin_i = torch.from_numpy(pts_h).pin_memory(), torch.from_numpy(np.ascontiguousarray(idx_h[lo:hi])).pin_memory()
        out_h = torch.empty(32, dtype=torch.float64).pin_memory()
        mask_h = torch.empty(M, dtype=torch.uint8).pin_memory()
        rec = torch.empty(device.WINREC_BYTES, dtype=torch.uint8, device="cuda")
        recs = torch.empty(world * device.WINREC_BYTES, dtype=torch.uint8, device="cuda")

        def e2e_step():
            p = pin_p.cuda(non_blocking=True)
            i = pin_i.cuda(non_blocking=True)
            r, m = device.ransac_sharded(p, i, lo, th, d, 4, "fwd", group=group, rec=rec, recs=recs, result=result, mask=mask)
            out_h.copy_(r, non_blocking=True)
            mask_h.copy_(m, non_blocking=True)
            torch.cuda.synchronize()
        n_e2e = max(2, min(args.steps, 5))
        ems = timed_steps(torch, dist, world, e2e_step, n_e2e, 1)
        e2e_s = ems / n_e2e / 1e3
        assert int(out_h[0]) == int(res[0]) and int(out_h[1]) == int(res[1]), "sharded e2e run disagrees with the device-resident run"
        h2d = world * pts_h.nbytes + idx_h.nbytes
        d2h = world * (32 * 8 + M)
        e2e_how = ("per rank: pinned host correspondences + this rank's shard of the sample sets -> device.ransac_sharded (48-byte winner records "
                   + ("over peer memory" if peer else "by one NCCL all-gather") + ") -> pinned host; bytes are the whole job's")
    e2e = {"value": K * M / e2e_s, "unit": "hyp*pts/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_s * 1e3, "how": e2e_how}
    rec = {"metric": "ransac_hyp_pts_per_s", "value": value, "unit": "hyp*pts/s", "ms_per_step": ms / args.steps,
           "config": {"workload": C3_WORKLOAD,
                      "sharding": f"hypothesis blocks over {world} rank(s), " + ("none" if world == 1 else ("48-byte winner records exchanged over peer memory (NVLink stores + flag wait), no collective"
                                                                                         if peer else "one 16-byte all-reduce(MAX) per step (NCCL)")),
                      "l2": "256 MB buffer written between timed steps (inputs are 16.3 MB and L2-resident by nature)"},
           "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
           "result": {"winner": int(res[0]), "inliers": int(res[1])}}
    return rec


def bench_warp(torch, dist, rank, world, args, ctx, peaks, local_gpu):
    from ransac_with_homography_b200 import device
    h, w = C4["h"], C4["w"]
    total = C4["frames"]
    lo, hi = device.shard_range(total, rank, world)
    n = hi - lo
    free, _ = torch.cuda.mem_get_info()
    per_frame = h * w * 3
    if 2 * n * per_frame > 0.85 * free:
        raise RuntimeError("not enough device memory for the frame shard")
    src = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda")
    for i in range(0, n, 64):                                   # generated on the device, seed = first frame id
        g.manual_seed(lo + i)
        m = min(64, n - i)
        src[i:i + m] = torch.randint(0, 256, (m, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    dst = torch.empty_like(src)
    Hs = frame_homographies(total, h, w, C4["jitter"], seed=0)[lo:hi]
    jobs = device.scan_jobs(Hs, (h, w), (h, w))
    jobs_dev = device.upload_jobs(jobs, "cuda")

    layout = (w * 3, n * h, h, w * 3)          # the shard's frames are one 2-D byte array: rows = frames * height

    def step():
        device.warp_resident(src, dst, jobs_dev, n, w, h, 3, "bilinear", layout=layout)

    # one cold call after a pause, timed alone: the rate before the power cap pulls the SM clock down (burst); the timed steps
    # below run back to back at the clock the 1000 W cap allows (sustained: tools/loop_strip.py shows 1965 -> ~1670 MHz from the
    # second call on), and `value` is the sustained figure
    step(); torch.cuda.synchronize(); time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    burst_ms = e0.elapsed_time(e1)
    l0 = ctx.launch_count()
    steps_ms = []
    with ClockSampler(local_gpu) as cs:
        ms = timed_steps(torch, dist, world, step, args.steps, args.warmup, None, per_step=steps_ms)
    launches = (ctx.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    pix = total * h * w
    value = pix * args.steps / (ms / 1e3) / 1e9
    bytes_per_launch = 2.0 * n * per_frame                      # each source byte once + each output byte once
    k_avg_s = ms / args.steps / 1e3
    roof = {"bound": "hbm", "kernel": "warp_strip_kernel (K4, bilinear u8 RGB; + warp_route_setup_kernel 2.7 %, warp_pipe_kernel for the tiles the strips leave, warp_fixup_kernel)", "achieved": bytes_per_launch / k_avg_s / 1e9, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": bytes_per_launch / k_avg_s / 1e9 / peaks["hbm_gbs"], "traffic": WARP_TRAFFIC_B_PER_PIXEL * n * h * w,
            "traffic_source": "ncu dram bytes per output pixel (profiles/r02_strip_v6.ncu-rep, 64 frames) x pixels of this launch",
            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})", "algorithmic_bytes_per_pixel": 6.0,
            "kernel_ms": k_avg_s * 1e3, "kernel_share_of_step": 1.0,
            "burst": {"ms": burst_ms, "achieved": bytes_per_launch / burst_ms / 1e6, "frac": bytes_per_launch / burst_ms / 1e6 / peaks["hbm_gbs"],
                      "what": "one launch sequence after a 1 s pause (SM clock at its maximum); the timed steps run under sw_power_cap"},
            "step_ms": [round(v, 3) for v in steps_ms]}

    # end to end on a bounded batch: pinned host frames in, pinned host frames out, through the host-buffer entry
    # point b200geo_warp_batch_host (chunked host->device / warp / device->host pipeline inside the library)
    nb = min(256, n)
    pin_in = torch.empty((nb, h, w, 3), dtype=torch.uint8).pin_memory()
    pin_in.copy_(src[:nb])
    pin_out = torch.empty((nb, h, w, 3), dtype=torch.uint8).pin_memory()
    jb = device.scan_jobs(Hs[:nb], (h, w), (h, w))

    def e2e_step():
        device.warp_host(pin_in, pin_out, jb, 3, "bilinear", device=local_gpu)
    n_e2e = max(2, min(args.steps, 5))
    ems = timed_steps(torch, dist, world, e2e_step, n_e2e, 1)
    e2e_v = world * nb * h * w * n_e2e / (ems / 1e3) / 1e9
    chk = torch.empty_like(src[:nb])
    device.warp(src[:nb], chk, jb, 3, "bilinear")
    # both are fast paths (the chunked pipeline re-bases the frames, so a tile may be routed to the other kernel): within the
    # +-1 of the parity contract of each other; parity proper is tests/test_gpu_full_size.py
    assert int((chk.cpu().to(torch.int16) - pin_out.to(torch.int16)).abs().max()) <= 1, "host-batch pipeline and device-resident warp disagree"
    rec = {"metric": "warp_gpix_per_s", "value": value, "unit": "Gpix/s", "ms_per_step": ms / args.steps,
           "dtype": "u8 pixels; f64 per-tile set-up, f32 tile-relative coordinates, 16.16 integer blend",
           "config": {"workload": C4_WORKLOAD, "launches": "per step: tile routing set-up, per-warp strip kernel, tile kernel for the rest, float64 fix-up",
                      "sharding": f"frames over {world} rank(s), no communication",
                      "l2": "inputs (25.5 GB per 4096 frames) are larger than L2"},
           "roofline": roof,
           "e2e": {"value": e2e_v, "unit": "Gpix/s", "h2d_bytes_per_step": int(world * nb * per_frame), "d2h_bytes_per_step": int(world * nb * per_frame),
                   "ms_per_step": ems / n_e2e, "how": f"{nb} frames per rank per step: pinned host -> device.warp_host (b200geo_warp_batch_host: 16-frame chunks, "
                                                        "copies and kernels overlapped on three streams) -> pinned host"},
           "gpu_launches": int(launches), "clocks": cs.summary()}
    del src, dst
    torch.cuda.empty_cache()
    return rec


def bench_stitch(torch, dist, rank, world, args, ctx, peaks, local_gpu):
    """C5: batched stitch pipeline, pairs sharded over the ranks, no communication."""
    from ransac_with_homography_b200 import device
    lo, hi = device.shard_range(C5["pairs"], rank, world)
    n = hi - lo
    h, w, (ch, cw) = C5["h"], C5["w"], C5["canvas"]
    pts_h, idx_h = stitch_pairs(lo, hi)
    pts, idx = torch.from_numpy(pts_h).cuda(), torch.from_numpy(idx_h).cuda()
    frames = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda")
    for i in range(0, n, 64):
        g.manual_seed(7000 + lo + i)
        m = min(64, n - i)
        frames[i:i + m] = torch.randint(0, 256, (m, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    canv = torch.zeros((n, ch, cw, 3), dtype=torch.uint8, device="cuda")
    from ransac_with_homography_b200 import _cabi as _c
    jobs_buf = torch.empty((n * _c.WARP_JOB_DTYPE.itemsize,), dtype=torch.uint8, device="cuda")
    boxes_buf = torch.empty((n, 5), dtype=torch.int32, device="cuda")
    out = {}

    def step():
        out["r"] = device.stitch_batched(pts, idx, frames, C5["th"], C5["d"], (ch, cw), canvases=canv, jobs=jobs_buf, boxes=boxes_buf)
    l0 = ctx.launch_count()
    ms = timed_steps(torch, dist, world, step, args.steps, args.warmup, None)
    launches = (ctx.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    _, boxes, results, _ = out["r"]
    res = results.cpu().numpy()
    boxes = boxes.cpu().numpy()
    pix = int((boxes[:, 2].astype(np.int64) * boxes[:, 3]).sum())
    if world > 1:
        t = torch.tensor([pix], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        pix = int(t.item())
    s = ms / args.steps / 1e3
    return {"metric": "stitch_pairs_per_s", "value": C5["pairs"] / s, "unit": "pairs/s", "ms_per_step": ms / args.steps,
            "ransac_hyp_pts_per_s": C5["pairs"] * C5["M"] * C5["K"] / s, "warp_gpix_per_s": pix / s / 1e9,
            "config": {"workload": f"C5 batched stitch: {C5['pairs']} pairs x {C5['M']} correspondences x {C5['K']} sample sets (fwd, th=4, d=95) "
                                   "+ refit + bilinear bbox warp of a 1080p frame into a 2160x3840 canvas slot; RANSAC and warp timed together "
                                   "(the two rates above both divide by the whole step)",
                       "sharding": f"pairs over {world} rank(s), no communication"},
            "gpu_launches": int(launches), "pairs_ok": int((res[:, 2] == 0).sum()), "median_inliers": float(np.median(res[:, 1]))}


def bench_aux(torch, ctx):
    """The two kernels beside the hot path that the driver's line did not show so far: K5 stitch
    compose (stitchPanorama as one kernel, device-resident 1080p pair) and K6 Hamming matcher (500 x 500 ORB descriptors with
    cross-check, host API: copies inside).  Latency-bound single calls: milliseconds, not rooflines."""
    from ransac_with_homography_b200 import device
    from ransac_with_homography_b200.ransac import match_hamming
    h, w = C4["h"], C4["w"]
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    q = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    t = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    H = frame_homographies(1, h, w, 0.1, seed=3)[0]
    H[0, 2] += 400.0
    out = {}
    for name, blend in (("paste", False), ("blend_rate", "Rate")):
        canvas = device.stitch_compose(q, t, H, blend, 0.2)
        for _ in range(3):
            device.stitch_compose(q, t, H, blend, 0.2, canvas=canvas)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            device.stitch_compose(q, t, H, blend, 0.2, canvas=canvas)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        out["compose_" + name] = {"ms_per_call": ms, "canvas": list(canvas.shape[:2]), "gpix_per_s": canvas.shape[0] * canvas.shape[1] / ms / 1e6,
                                  "what": "K5 stitch_compose_kernel: warp of a 1080p image + placement + " + name + " + uint8, device-resident"}
    rs = np.random.RandomState(2)
    A = rs.randint(0, 256, (500, 32)).astype(np.uint8)
    Bd = rs.randint(0, 256, (500, 32)).astype(np.uint8)
    for _ in range(3):
        m = match_hamming(A, Bd, True)
    t0 = time.perf_counter()
    for _ in range(20):
        m = match_hamming(A, Bd, True)
    out["match_hamming"] = {"ms_per_call": (time.perf_counter() - t0) / 20 * 1e3, "matches": len(m),
                            "what": "K6 hamming_best_kernel + cross-check, 500 x 500 descriptors of 256 bits, host API (copies and Match list inside)"}
    # K7: ORB detect + describe through the host API (upload, grey conversion, 8 pyramid levels, FAST, Harris, the host's two
    # retainBest selections, orientation, blur, 256 pair tests, download) on the reference's own C1 image (tests/golden/foto1A.jpg,
    # 683 x 1024) and on a synthetic 1080p frame of 8 x 8 blocks + noise (~20x the FAST candidates of a photograph: the host's
    # ordering of the candidates dominates there); OpenCV on one host core beside it when importable
    from ransac_with_homography_b200 import imgio
    from ransac_with_homography_b200.ransac import orb_detect_and_compute
    base = rs.randint(0, 256, (h // 8 + 1, w // 8 + 1)).astype(np.float32)
    blocky = np.clip(np.kron(base, np.ones((8, 8), np.float32))[:h, :w] + rs.randn(h, w).astype(np.float32) * 4, 0, 255).astype(np.uint8)
    try:
        import cv2
        cv2.setNumThreads(1)
    except ImportError:
        cv2 = None
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "foto1A.jpg")
    photo = np.ascontiguousarray(cv2.imread(path)[:, :, ::-1]) if cv2 is not None else imgio.imread(path, order="rgb")
    for key, img, what in (("orb", photo, "C1 image foto1A.jpg 683 x 1024 RGB (grey conversion on the device)"), ("orb_1080p_blocks", blocky, "synthetic 1080p grey frame, 8 x 8 blocks + noise")):
        for _ in range(3):
            kp, ds = orb_detect_and_compute(img, 500, as_arrays=True)
        t0 = time.perf_counter()
        for _ in range(10):
            kp, ds = orb_detect_and_compute(img, 500, as_arrays=True)
        rec = {"ms_per_call": (time.perf_counter() - t0) / 10 * 1e3, "keypoints": int(len(kp)), "frame": list(img.shape[:2]),
               "what": "K7 ORB (cv2.cvtColor + cv2.ORB_create().detectAndCompute), host API: " + what}
        if cv2 is not None:
            orb = cv2.ORB_create()
            g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) if img.ndim == 3 else img
            orb.detectAndCompute(g, None)
            t0 = time.perf_counter()
            for _ in range(3):
                g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) if img.ndim == 3 else img
                k2, d2 = orb.detectAndCompute(g, None)
            rec["opencv_cpu_ms_per_call"] = (time.perf_counter() - t0) / 3 * 1e3
            rec["opencv_threads"] = 1
            rec["identical_to_opencv"] = bool(len(k2) == len(kp) and np.array_equal(d2, ds) and
                                              np.array_equal(np.float32([[k.pt[0], k.pt[1], k.angle, k.response] for k in k2]), kp[:, [0, 1, 3, 4]]))
        out[key] = rec
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa = numa_bind(local)
    from ransac_with_homography_b200 import _cabi
    ctx = _cabi.context(local)
    peaks = measured_peaks()
    link = measure_link(torch, dist, world)
    link["numa"] = numa
    line = None
    warp_rec = None
    if args.workload in ("all", "ransac"):
        line = bench_ransac(torch, dist, rank, world, args, ctx, peaks, local)
    if args.workload in ("all", "warp"):
        warp_rec = bench_warp(torch, dist, rank, world, args, ctx, peaks, local)
    stitch_rec = None
    if args.workload in ("all", "stitch"):
        torch.cuda.empty_cache()
        stitch_rec = bench_stitch(torch, dist, rank, world, args, ctx, peaks, local)
    if line is None and warp_rec is None:
        line, stitch_rec = stitch_rec, None
        line.setdefault("clocks", {"sm_mhz": None, "sm_max_mhz": None, "reasons": []})
    if line is None:
        line, warp_rec = warp_rec, None
    # e2e roofline: the copies of a step against the measured host link (both directions at once)
    for rec in (line, warp_rec):
        if rec is not None and "e2e" in rec:
            e = rec["e2e"]
            e["link"] = link
            ms = e.get("ms_per_step")
            if ms:
                gbs = (e["h2d_bytes_per_step"] + e["d2h_bytes_per_step"]) / (ms / 1e3) / 1e9
                e["copy_gbs"] = gbs              # whole job, both directions
                e["link_frac"] = gbs / (2.0 * link["sum_ranks_each_direction_concurrent_gbs"])
    line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
                 "vs_baseline": None, "data": "synthetic"})
    line.setdefault("dtype", "f32")
    if stitch_rec is not None:
        line["stitch"] = stitch_rec
        line["gpu_launches"] = line["gpu_launches"] + stitch_rec["gpu_launches"]
    if warp_rec is not None:
        line["warp"] = warp_rec
        line["clocks"] = merge_clocks(line["clocks"], warp_rec["clocks"])
        line["gpu_launches"] = line["gpu_launches"] + warp_rec["gpu_launches"]
    # CPU baseline: rank 0, N = 1 only, bounded sample, single process (the reference is single-threaded as shipped);
    # the unmodified reference when oracle/_ref is present (kind "reference"), the numpy port beside it as a second figure
    if rank == 0 and world == 1 and not args.no_cpu:
        kind = cpu_kind()
        if line["metric"].startswith("ransac"):
            v, dt = cpu_ransac(50000, 1, None, kind)
            line["cpu_baseline"] = {"value": v, "unit": "hyp*pts/s", "cores": 1, "kind": kind, "seconds": dt,
                                    "sample": "50000 of 1000000 sample sets x 20000 correspondences, " +
                                              ("RANSAC.run of the unmodified reference (oracle/_ref)" if kind == "reference" else "numpy port of ransac.py:176-184 (oracle/ref_numpy.py)") +
                                              ", single process as the reference runs"}
            if kind == "reference":
                v2, dt2 = cpu_ransac(20000, 1, None, "port")
                line["cpu_baseline_port"] = {"value": v2, "unit": "hyp*pts/s", "cores": 1, "kind": "port", "seconds": dt2,
                                             "sample": "20000 sample sets, numpy port (oracle/ref_numpy.py)"}
        tgt = line.get("warp") if "warp" in line else (line if line["metric"].startswith("warp") else None)
        if tgt is not None:
            v, dt = cpu_warp(16, 1, kind)
            tgt["cpu_baseline"] = {"value": v, "unit": "Gpix/s", "cores": 1, "kind": kind, "seconds": dt,
                                   "sample": "16 of 4096 frames 1080p bilinear, " +
                                             ("wrapPerspectiveScan of the unmodified reference (oracle/_ref)" if kind == "reference" else "numpy port of homography.py:186-209,123-138")}
    if rank == 0 and world == 1 and args.workload == "all":
        line["aux"] = bench_aux(torch, ctx)
    if rank == 0:
        order = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"]
        out = {k: line[k] for k in order if k in line}
        out.update({k: v for k, v in line.items() if k not in out})
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "ransac", "warp", "stitch"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
