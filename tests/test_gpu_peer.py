"""GPU, two or more devices: the winner records of a sharded search exchanged over peer memory (b200geo_peer_*,
b200geo_ransac_exchange_peer_dev) give the same result as the NCCL all-gather and as the unsharded search.  Skipped on a
single-GPU box."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from conftest import synth_pair
    from ransac_with_homography_b200 import device
    M, K = 3000, 40000
    X, Y = synth_pair(3, M)
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate([X.T, Y.T], 1), np.float32)).cuda()
    idx = np.random.RandomState(1).randint(0, M, (K, 4)).astype(np.int32)
    lo, hi = device.shard_range(K, rank, world)
    shard = torch.from_numpy(idx[lo:hi]).cuda()
    res = {}
    r0, m0 = device.ransac_sharded(pts, shard, lo, 3.0, 99.0)                 # NCCL all-gather
    res["nccl"] = (r0.cpu().numpy().copy(), m0.cpu().numpy().copy())
    ok = device.peer_setup()
    res["peer_ok"] = ok
    if ok:
        for it in range(5):                                                   # several epochs: both parities, reuse of the slots
            r1, m1 = device.ransac_sharded(pts, shard, lo, 3.0 + 0.25 * (it % 2), 99.0)
            if it % 2 == 0:
                res["peer"] = (r1.cpu().numpy().copy(), m1.cpu().numpy().copy())
        device.peer_close()
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_equals_collective():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = min(n, 4)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        res = out[r]
        assert res["peer_ok"], "peer memory could not be mapped on a multi-GPU B200 node"
        a, b = res["nccl"], res["peer"]
        assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[0], out[0]["nccl"][0], equal_nan=True)          # every rank finishes with the same result
