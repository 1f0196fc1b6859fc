"""Microbenchmark of K4 on the C4 shape (1080p RGB frames, +-5% corner jitter, Scan grid).
    python tools/bench_warp.py [frames] [sampler] [exact]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import frame_homographies
from ransac_with_homography_b200 import _cabi, device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sampler = sys.argv[2] if len(sys.argv) > 2 else "bilinear"
exact = len(sys.argv) > 3 and sys.argv[3] == "exact"
h, w = 1080, 1920
ctx = _cabi.context(0)
ctx.warp_exact(exact)
ctx.warp_variant(int(os.environ.get("WARP_VARIANT", "0")))
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
jobs = device.upload_jobs(device.scan_jobs(frame_homographies(n, h, w, float(os.environ.get("JITTER", "0.05"))), (h, w), (h, w)), "cuda")
LAYOUT = None if os.environ.get('NO_TMAP') else ((w * 3, n * h, h, w * 3) if not os.environ.get('TMAP2D') else (w * 3, n * h))
for _ in range(3):
    device.warp_resident(src, dst, jobs, n, w, h, 3, sampler, layout=LAYOUT)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 5
ev[0].record()
for _ in range(reps):
    device.warp_resident(src, dst, jobs, n, w, h, 3, sampler, layout=LAYOUT)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
print(f"{sampler} exact={exact} frames={n}: {ms:.3f} ms  {n*h*w/ms/1e6:.1f} Gpix/s  {2*n*h*w*3/ms/1e6:.0f} GB/s  "
      f"exact-path pixels {ctx.warp_exact_pixels()/(reps+3)/(n*h*w):.2e} of all")
