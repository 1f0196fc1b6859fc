"""Scoring throughput per loss method on the C3 correspondence set (M = 20000), K hypotheses.
    python tools/bench_score.py [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_pair, C3
from ransac_with_homography_b200 import _cabi, device
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
M = C3["M"]
X, Y = synth_pair(C3["seed"], M)
pts = torch.from_numpy(np.ascontiguousarray(np.concatenate([X.T, Y.T], 1), np.float32)).cuda()
idx = torch.from_numpy(np.random.RandomState(0).randint(0, M, (K, 4)).astype(np.int32)).cuda()
H, Hinv, flags = device.dlt4(pts, idx, want_inv=True)
ctx = _cabi.context(0)
for method in ("fwd", "backward", "reproj"):
    for exact in (False, True):
        ctx.force_exact(exact)
        counts = torch.zeros(K, dtype=torch.int32, device="cuda")
        for _ in range(2):
            device.score(pts, H, C3["th"], method, Hinv=Hinv, flags=flags, counts=counts)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.exact_path_pairs(reset=True)
        a.record()
        for _ in range(3):
            device.score(pts, H, C3["th"], method, Hinv=Hinv, flags=flags, counts=counts)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{method:9s} exact={exact}: {ms:8.3f} ms  {K*M/ms/1e6:9.1f} Gpairs/s  best {int(counts.max())}  exact-path {ctx.exact_path_pairs()/3/(K*M):.2e}")
ctx.force_exact(False)
