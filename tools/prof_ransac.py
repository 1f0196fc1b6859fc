"""Short driver for ncu: one warm-up and two timed passes of the C3 RANSAC search (K = 1M, M = 20k)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import C3, synth_pair  # noqa: E402
from ransac_with_homography_b200 import _cabi, device  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else C3["K"]
M = C3["M"]
X, Y = synth_pair(C3["seed"], M)
pts_h = np.empty((M, 4), np.float32)
pts_h[:, :2], pts_h[:, 2:] = X.T, Y.T
idx_h = np.random.RandomState(0).randint(0, M, (K, 4)).astype(np.int32)
pts, idx = torch.from_numpy(pts_h).cuda(), torch.from_numpy(idx_h).cuda()
ctx = _cabi.context(0)
for i in range(3):
    ctx.exact_path_pairs(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res, mask = device.ransac(pts, idx, C3["th"], C3["d"])
    b.record()
    torch.cuda.synchronize()
    r = res.cpu().numpy()
    print(f"pass {i}: {a.elapsed_time(b):.3f} ms  winner {int(r[0])} inliers {int(r[1])}  exact-path pairs {ctx.exact_path_pairs()}"
          f" ({ctx.exact_path_pairs() / (K * M):.2e} of all)")
