"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): SURVEY §4 plan (iv) / §8e on hardware.
Two NCCL ranks each differentiate `cost + ioc_cost` of THEIR scenes (normalised by the all-reduced agent count), the
flat gradients are all-reduced — and must equal the gradient one GPU computes for the union minibatch.  Then one
clip + Adam step on every rank must leave bit-identical replicas that match the single-GPU update."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_l2, small_cfg

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from desire_b200.config import init_params
    from desire_b200.dist import all_reduce_gradients_, shard_scenes
    from desire_b200.engine import TrainPath, existing_agents, flatten_params
    from desire_b200.synthetic import make_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    cfg = small_cfg(d_dim=128, max_num_obj=12, num_samples=4, n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3)
    B = 4
    host = make_batch(cfg, B, 0, 2)
    mine = shard_scenes(B, rank, world)
    N, K = cfg.max_num_obj, cfg.K
    obs, tgt, scene = [host[i][mine].contiguous().to(dev) for i in (0, 1, 3)]
    eps = host[2].reshape(B, N, K, cfg.Z)[mine].reshape(-1, K, cfg.Z).contiguous().to(dev)
    flat, views, offs = flatten_params(init_params(cfg, 1), dev)
    tp = TrainPath(cfg, flat, views, offs, len(mine), dev, train_ioc=True)
    tp.set_count(obs, tgt)                       # all-reduced: the global normaliser
    tp.run(obs, tgt, eps, scene, stages=("generate",))
    tp.backward(obs, tgt, eps, scene)
    all_reduce_gradients_(tp.grad_flat)
    torch.cuda.synchronize()
    g_dp = tp.grad_flat.cpu().numpy().copy()
    n_global = float(tp.count[0])
    # one optimiser step (apply() all-reduces again, so feed it the local gradient: redo the backward)
    tp.backward(obs, tgt, eps, scene)
    tp.apply(1e-3, 10.0)
    torch.cuda.synchronize()
    w_dp = tp.flat.cpu().numpy().copy()
    out = {"rank": rank, "g_dp": g_dp, "w_dp": w_dp, "count": n_global}
    if rank == 0:
        # the single-GPU answer for the union minibatch (no collective: the count is set by hand)
        full = [t.to(dev) for t in host]
        flat1, views1, offs1 = flatten_params(init_params(cfg, 1), dev)
        t1 = TrainPath(cfg, flat1, views1, offs1, B, dev, train_ioc=True)
        t1.count.copy_(existing_agents(full[0], full[1], cfg.exist_mode).sum().float().reshape(1))
        t1.run(*full, stages=("generate",))
        t1.backward(*full)
        torch.cuda.synchronize()
        out["g_1"] = t1.grad_flat.cpu().numpy().copy()
        out["count_1"] = float(t1.count[0])
        lib = t1.lib
        import ctypes as C
        from desire_b200 import _lib
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.desire_sumsq_fwd(C.c_void_p(t1.grad_flat.data_ptr()), t1.flat.numel(), C.c_void_p(t1.sumsq.data_ptr()), 0, st), "sumsq")
        _lib.check(lib.desire_adam_step(C.c_void_p(t1.flat.data_ptr()), C.c_void_p(t1.grad_flat.data_ptr()),
                                        C.c_void_p(t1.adam_m.data_ptr()), C.c_void_p(t1.adam_v.data_ptr()), t1.flat.numel(),
                                        C.c_void_p(t1.sumsq.data_ptr()), 1e-3, 0.9, 0.999, 1e-8, 1, 10.0, 1.0, st), "adam")
        torch.cuda.synchronize()
        out["w_1"] = t1.flat.cpu().numpy().copy()
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_allreduced_gradient_equals_single_gpu_gradient_of_the_union_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=600) for _ in ps], key=lambda d: d["rank"])
    [p.join(timeout=120) for p in ps]
    r0, r1 = res
    assert r0["count"] == r1["count"] == r0["count_1"]
    assert np.array_equal(r0["g_dp"], r1["g_dp"])                 # the all-reduce leaves identical buffers
    e = rel_l2(r0["g_dp"], r0["g_1"])
    print("all-reduced 2-rank gradient vs 1-GPU union-batch gradient: rel-L2 %.3e (|g| %.3e)" % (e, np.linalg.norm(r0["g_1"])))
    assert e <= 2e-5
    assert np.array_equal(r0["w_dp"], r1["w_dp"])                 # replicas stay bit-identical after clip + Adam
    dw = float(np.linalg.norm(r0["w_dp"].astype(np.float64) - r0["w_1"]) / np.linalg.norm(r0["w_1"].astype(np.float64)))
    print("weights after one step, 2 ranks vs 1 GPU: |w_dp - w_1| / |w_1| = %.3e" % dw)
    assert dw <= 1e-5          # lr 1e-3: an Adam step moves a weight by at most ~1e-3; sign flips of ~zero gradients stay far below
