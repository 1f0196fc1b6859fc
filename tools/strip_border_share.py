"""How much of the C4 time goes to strips that cross the sampler's bounds: the same launch with jitter 0 (a pure grid rescale, every
strip interior except the frame's last row / column) against the benchmark's 5 % jitter.   python tools/strip_border_share.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import frame_homographies
from ransac_with_homography_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h, w = 1080, 1920
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
lay = (w * 3, n * h, h, w * 3)
for jit in (0.0, 0.01, 0.05, 0.10):
    jobs = device.upload_jobs(device.scan_jobs(frame_homographies(n, h, w, jit), (h, w), (h, w)), "cuda")
    for _ in range(3):
        device.warp_resident(src, dst, jobs, n, w, h, 3, "bilinear", layout=lay)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        device.warp_resident(src, dst, jobs, n, w, h, 3, "bilinear", layout=lay)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"jitter {jit:.2f}: {ms:.3f} ms  {n*h*w/ms/1e6:.1f} Gpix/s", flush=True)
    import time; time.sleep(1.0)
