"""GPU: the C ABI's feature exchange (hvb_comm_* / hvb_allgather_counts / hvb_allgather_features) on a
world of one rank — communicator life cycle, counts, rows, the empty case.  The multi-rank comparison against
torch.distributed's all-gather runs under torchrun (tools/check_comm_multigpu.py, 2+ GPUs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_world_of_one_roundtrip():
    from hvb.runtime import get_context
    ctx = get_context(0)
    uid = ctx.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    comm = ctx.comm_create(uid, 1, 0)
    try:
        rng = np.random.default_rng(0)
        x = torch.from_numpy(rng.standard_normal((37, 625))).cuda()
        out, counts = ctx.allgather_features(comm, x, 1)
        assert counts.tolist() == [37]
        assert torch.equal(out, x) and out.data_ptr() != x.data_ptr()
        out, counts = ctx.allgather_features(comm, x[:0], 1)
        assert counts.tolist() == [0] and out.shape == (0, 625)
    finally:
        ctx.comm_destroy(comm)


def test_bad_arguments_are_loud():
    from hvb import _ffi
    from hvb.runtime import get_context
    ctx = get_context(0)
    with pytest.raises(_ffi.HvbError):
        ctx.comm_create(ctx.comm_unique_id(), 2, 5)            # rank outside the world
    with pytest.raises(ValueError):
        ctx.allgather_features(None, torch.zeros((3, 4), device="cuda"), 1)   # float32
