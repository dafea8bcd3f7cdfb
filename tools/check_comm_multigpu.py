"""Multi-rank check of the C ABI's feature exchange against torch.distributed (run under torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_comm_multigpu.py
Every rank holds a different number of float64[*, 625] rows (one rank holds none); both backends must return the same
matrix bit for bit on every rank.  Prints one JSON line from rank 0 with the two exchange times."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hockey-vision-analytics_b200")]
import numpy as np
import torch
import torch.distributed as dist
from hvb.dist import all_gather_features

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = {}
for rows_of in (lambda r: 0 if r == 1 else 150 + 37 * r, lambda r: 250):
    n = rows_of(rank)
    x = torch.from_numpy(np.random.default_rng(rank).standard_normal((n, 625))).cuda()
    a = all_gather_features(x, backend="torch")
    b = all_gather_features(x, backend="hvb")
    assert a.shape == b.shape == (sum(rows_of(r) for r in range(world)), 625), (a.shape, b.shape)
    assert torch.equal(a, b), "hvb exchange differs from torch.distributed's"
    lo = sum(rows_of(r) for r in range(rank))
    assert torch.equal(b[lo:lo + n], x)
    for tag in ("torch", "hvb"):
        for _ in range(3):
            all_gather_features(x, backend=tag)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            all_gather_features(x, backend=tag)
        torch.cuda.synchronize()
        res["%s_ms_rows%d" % (tag, a.shape[0])] = 1e3 * (time.perf_counter() - t0) / 20
if rank == 0:
    print(json.dumps({"world": world, "ok": True, **res}))
dist.destroy_process_group()
