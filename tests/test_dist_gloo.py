"""(e) multi-GPU host logic on CPU, world_size 2, gloo: the generated batch is sharded by contiguous image ranges
with NO data-path collective; each rank evaluates the Philox values of its GLOBAL element indices, so the union of
the shards reproduces the single-process stream.  Only the final gather of results uses a collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox


def shard_range(n_global: int, rank: int, world: int):
    """Contiguous image range of a rank (what bench.py and AbsorbingDiffusion.sample(n_global, shard_base) use)."""
    per = (n_global + world - 1) // world
    lo = min(rank * per, n_global)
    return lo, min(lo + per, n_global)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_global, K, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_global, rank, world)
    hw = 49
    # the two draws of one sampling step for this rank's tokens only (global indices!)
    u = philox.uniform(7, 0, n_global * hw, 148, 2048, index_base=lo * hw, numel=(hi - lo) * hw)
    e = philox.exponential(7, 4, n_global * hw * K, 148, 2048, index_base=lo * hw * K, numel=(hi - lo) * hw * K)
    # a stand-in "sampling step": argmax of uniform probs / exponential, masked by the unmask decision
    tok = torch.from_numpy(e).reshape(-1, K).reciprocal().argmax(-1)
    tok[torch.from_numpy(u) >= 0.5] = K
    per = (n_global + world - 1) // world
    buf = torch.full((per * hw,), -1, dtype=torch.int64)
    buf[: tok.numel()] = tok
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)            # result gather only: outside the timed sampling path
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.cat(out)
        ret["tokens"] = full[full >= 0].numpy()
        ret["tmax"] = float(t)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stream_equals_single_process():
    n_global, K, world = 5, 128, 2        # odd batch: ragged last shard
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_range(4096, r, 8)[1] - shard_range(4096, r, 8)[0] for r in range(8)] == [512] * 8
    assert shard_range(1, 1, 2) == (1, 1)  # empty shard
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_global, K, ret), nprocs=world, join=True)
    hw = 49
    u = philox.uniform(7, 0, n_global * hw, 148, 2048)
    e = philox.exponential(7, 4, n_global * hw * K, 148, 2048)
    tok = torch.from_numpy(e).reshape(-1, K).reciprocal().argmax(-1)
    tok[torch.from_numpy(u) >= 0.5] = K
    assert np.array_equal(ret["tokens"], tok.numpy())
    assert ret["tmax"] == 2.0
