"""A few launches of the routed fast warp (bilinear, C4 shape) for ncu.   python tools/prof_strip.py [frames] [sampler]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import frame_homographies
from ransac_with_homography_b200 import _cabi, device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sampler = sys.argv[2] if len(sys.argv) > 2 else "bilinear"
h, w = 1080, 1920
ctx = _cabi.context(0)
ctx.warp_variant(int(os.environ.get("WARP_VARIANT", "0")))
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
jobs = device.upload_jobs(device.scan_jobs(frame_homographies(n, h, w, 0.05), (h, w), (h, w)), "cuda")
for _ in range(5):
    device.warp_resident(src, dst, jobs, n, w, h, 3, sampler, layout=(w * 3, n * h, h, w * 3))
torch.cuda.synchronize()
print("ok")
