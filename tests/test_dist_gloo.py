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


def _grad_worker(rank, world, port, q):
    """Data-parallel train-step arithmetic on CPU: each rank differentiates the cost of ITS scenes (the autograd
    oracle twin stands in for the CUDA backward), normalised by the GLOBAL agent count, then the flat gradients
    are summed with the same helpers the GPU engine uses."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import np_batch, np_params, small_cfg
    from oracle import desire_oracle_torch as OT
    from desire_b200.dist import all_reduce_gradients_, global_count_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = small_cfg(d_dim=16, max_num_obj=4, num_samples=2)
    B = 4
    inp, tgt, eps, _ = np_batch(cfg, B, n_missing=1, dtype=np.float64)
    N = cfg.max_num_obj
    mine = shard_scenes(B, rank, world)
    P = OT.to_torch(np_params(cfg, dtype=np.float64))
    eps_b = eps.reshape(B, N, cfg.K, cfg.Z)
    out = OT.generate_forward(P, dict(K=cfg.K, Z=cfg.Z), inp[mine], tgt[mine], eps_b[mine].reshape(-1, cfg.K, cfg.Z))
    mask = torch.as_tensor(inp[mine][:, :, 0, 0] != 0).reshape(-1)
    count = global_count_(mask.sum().double().reshape(1))
    local = ((out["recon_rows"] + out["kld_rows"]) * mask).sum() / count[0]
    local.backward()
    names = [k for k, v in P.items() if v.grad is not None]
    flat = torch.cat([P[k].grad.reshape(-1) for k in names])
    all_reduce_gradients_(flat)
    q.put((rank, names, flat.numpy(), float(count[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_the_single_process_gradient():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import np_batch, np_params, small_cfg
    from oracle import desire_oracle_torch as OT
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=300) for _ in ps]
    [p.join(timeout=60) for p in ps]
    cfg = small_cfg(d_dim=16, max_num_obj=4, num_samples=2)
    inp, tgt, eps, _ = np_batch(cfg, 4, n_missing=1, dtype=np.float64)
    P = OT.to_torch(np_params(cfg, dtype=np.float64))
    OT.generate_forward(P, dict(K=cfg.K, Z=cfg.Z), inp, tgt, eps)["cost"].backward()
    for rank, names, flat, count in res:
        ref = torch.cat([P[k].grad.reshape(-1) for k in names]).numpy()
        assert count == float((inp[:, :, 0, 0] != 0).sum())
        assert np.allclose(flat, ref, rtol=1e-9, atol=1e-12), rank
