"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding helpers and the
variable-count feature all-gather (uneven N_g including 0) used before the global team fit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, q):
    sys.path.insert(0, os.path.join(ROOT, "hockey-vision-analytics_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hvb.dist import all_gather_features, shard_range
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(counts[rank], 625, generator=g, dtype=torch.float64)
    out = all_gather_features(local)
    lo, hi = shard_range(21, rank, world)
    q.put((rank, out.numpy(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(5, 9), (0, 7), (4, 0), (0, 0)])
def test_all_gather_features_world2(counts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.concatenate([torch.randn(counts[r], 625, generator=torch.Generator().manual_seed(100 + r), dtype=torch.float64).numpy()
                             for r in range(2)])
    for rank, out, rng_ in res:
        assert out.shape == expect.shape and np.array_equal(out, expect)       # bit-exact, rank order
    assert res[0][2] == (0, 11) and res[1][2] == (11, 21)


def test_shard_helpers():
    from hvb.dist import shard_clips, shard_range
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard_clips(8, 3, 8) == [3]
