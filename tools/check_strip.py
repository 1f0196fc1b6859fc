"""Quick check + timing of the routed strip kernel (variant 0) against the tile kernel (variant 2) and the float64 kernel.
    python tools/check_strip.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import frame_homographies
from ransac_with_homography_b200 import _cabi, device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h, w = 1080, 1920
ctx = _cabi.context(0)
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
jobs = device.upload_jobs(device.scan_jobs(frame_homographies(n, h, w, 0.05), (h, w), (h, w)), "cuda")
layout = (w * 3, n * h, h, w * 3)
for sampler in ("bilinear", "nn"):
    m = min(n, 8)
    ctx.warp_exact(True)
    ex = torch.empty((m, h, w, 3), dtype=torch.uint8, device="cuda")
    device.warp_resident(src[:m], ex, jobs, m, w, h, 3, sampler)
    ctx.warp_exact(False)
    for variant in (0, 2):
        ctx.warp_variant(variant)
        dst = torch.full((n, h, w, 3), 9, dtype=torch.uint8, device="cuda")
        ctx.warp_exact_pixels(reset=True)
        device.warp_resident(src, dst, jobs, n, w, h, 3, sampler, layout=layout)
        torch.cuda.synchronize()
        fb = ctx.warp_exact_pixels(reset=True)
        d = (dst[:m].to(torch.int16) - ex.to(torch.int16)).abs()
        for _ in range(3):
            device.warp_resident(src, dst, jobs, n, w, h, 3, sampler, layout=layout)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 10
        ev[0].record()
        for _ in range(reps):
            device.warp_resident(src, dst, jobs, n, w, h, 3, sampler, layout=layout)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        print(f"{sampler} variant {variant} frames {n}: {ms:.3f} ms {n*h*w/ms/1e6:.1f} Gpix/s {2*n*h*w*3/ms/1e6:.0f} GB/s | vs float64 kernel: max {int(d.max())} "
              f"off-by-one {float((d > 0).float().mean()):.5f} | float64 pixels {fb/(n*h*w):.2e}", flush=True)
    ctx.warp_variant(0)
