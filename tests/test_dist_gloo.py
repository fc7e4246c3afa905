"""CPU, world_size 2 over gloo: scene sharding covers every scene exactly once, and the globally
normalised masked cost equals the single-process value."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from desire_b200.dist import global_masked_cost, shard_scenes


def _worker(rank, world, port, rows, mask, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_scenes(rows.shape[0], rank, world)
    s = torch.tensor(float((rows[mine] * mask[mine]).sum()))
    c = torch.tensor(float(mask[mine].sum()))
    q.put((rank, mine, float(global_masked_cost(s, c))))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_scenes_partition():
    for n in (0, 1, 5, 32):
        for w in (1, 2, 3, 8):
            parts = [shard_scenes(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_global_cost_matches_single_process():
    rng = np.random.default_rng(0)
    rows = rng.random((7, 5))                     # 7 scenes x 5 agents of per-agent loss
    mask = (rng.random((7, 5)) > 0.3).astype(np.float64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, rows, mask, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    ref = (rows * mask).sum() / mask.sum()
    seen = []
    for rank, mine, cost in res:
        assert abs(cost - ref) < 1e-6
        seen += mine
    assert sorted(seen) == list(range(7))
