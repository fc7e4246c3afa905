"""Train-script entry point with the reference's flags and loop shape (train.py:24-207).

    python -m desire_b200.train --d_dim 128 --batch_size 32 --num_samples 20 ...

Same 19 argparse flags with the same names and defaults (train.py:28-88) plus the knobs the build adds
(DESIGN.md D1/D2/D11).  The reference's loop only ever evaluates `model.cost` (train.py:181; the optimiser op of
model/model.py:394 is created and never run).  Here every minibatch is one optimiser step (D9):
forward + backward of `cost` + clip_by_global_norm(grad_clip) + Adam(learning_rate * decay_rate**epoch) on the GPU
(`--optimize 0` restores the reference's evaluate-only loop, which then also runs IOC ranking/refinement), the
reference's log line, and a checkpoint every `save_every` steps that `--resume` restores (weights + Adam moments +
step; the restore path the reference lacks, train.py:197-207).  Where the reference ran one sess.run per SEQUENCE
(train.py:146-181), a whole minibatch of scenes is one pass here.
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import time

import numpy as np


def build_parser():
    p = argparse.ArgumentParser()
    # ---- the reference's flags, verbatim (train.py:30-88)
    p.add_argument('--rnn_size', type=int, default=512, help='size of RNN hidden state')
    p.add_argument('--num_layers', type=int, default=1, help='number of layers in the RNN')
    p.add_argument('--model', type=str, default='gru', help='rnn, gru, or lstm')
    p.add_argument('--batch_size', type=int, default=10, help='minibatch size')
    p.add_argument('--seq_length', type=int, default=8, help='RNN sequence length')
    p.add_argument('--num_epochs', type=int, default=100, help='number of epochs')
    p.add_argument('--save_every', type=int, default=400, help='save frequency')
    p.add_argument('--grad_clip', type=float, default=10., help='clip gradients at this value')
    p.add_argument('--learning_rate', type=float, default=0.005, help='learning rate')
    p.add_argument('--decay_rate', type=float, default=0.95, help='decay rate for rmsprop')
    p.add_argument('--keep_prob', type=float, default=0.8, help='dropout keep probability')
    p.add_argument('--embedding_size', type=int, default=64, help='Embedding dimension for the spatial coordinates')
    p.add_argument('--neighborhood_size', type=int, default=32, help='Neighborhood size to be considered for social grid')
    p.add_argument('--grid_size', type=int, default=4, help='Grid size of the social grid')
    p.add_argument('--max_num_obj', type=int, default=60, help='Maximum Number of Moving objects')
    p.add_argument('--leave_dataset', type=int, default=5, help='The dataset index to be left out in training')
    p.add_argument('--latent_size', type=int, default=128, help='The latent size for CVAE')
    p.add_argument('--e_dim', type=int, default=256, help="The encoder's output dimension")
    p.add_argument('--d_dim', type=int, default=16, help="The decoder's output dimension")
    p.add_argument('--stride', type=int, default=1, help='Stride size for the Temporal Convolution')
    # ---- added
    p.add_argument('--pred_length', type=int, default=12, help='T_f, future frames to predict (D2)')
    p.add_argument('--num_samples', type=int, default=20, help='K, CVAE samples per agent (D1)')
    p.add_argument('--ioc_iters', type=int, default=2, help='IOC ranking/refinement iterations (D11)')
    p.add_argument('--scene_size', type=int, default=256, help='scene image side fed to the scene CNN (D11)')
    p.add_argument('--data_dir', type=str, default='data/', help='directory holding <scene>/<video>/annotations_processed.csv')
    p.add_argument('--save_dir', type=str, default='save', help='checkpoint / config directory (reference: save/)')
    p.add_argument('--clip_objects', action='store_true', help='keep the first max_num_obj objects of crowded frames instead of raising')
    p.add_argument('--norm_w', type=float, default=0.0, help='divide x by this (0 with --norm_h 0: the extent of the loaded data, so positions land in [0,1])')
    p.add_argument('--norm_h', type=float, default=0.0, help='divide y by this')
    p.add_argument('--raw_pixels', action='store_true', help="keep the reference's raw pixel coordinates (the IOC stage's scene gather and log-polar radii assume the unit square: only sensible with --ioc_iters 0)")
    p.add_argument('--fix_id0', action='store_true', help='shift track ids by +1 so SDD track 0 is not taken for the "no object" sentinel (reference quirk, utils/data_loader.py:221-222)')
    p.add_argument('--prefetch', type=int, default=2, help='minibatches built ahead by the DataLoader thread into pinned buffers (0 = build each batch inside the step loop)')
    p.add_argument('--max_batches', type=int, default=0, help='stop an epoch after this many batches (0 = all)')
    p.add_argument('--optimize', type=int, default=1, help='1: run the Adam step per minibatch (D9); 0: evaluate cost only, as the reference loop does')
    p.add_argument('--resume', type=str, default='', help='checkpoint file written by a previous run to continue from')
    p.add_argument('--seed', type=int, default=1)
    p.add_argument('--device', type=str, default='cuda:0')
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    train(args)


def save_checkpoint(model, path, next_step, loader_state=None):
    """Weights by name (the reference's tf.train.Saver keeps variables by name, train.py:114,200-205) plus the
    optimiser state (Adam moments + step, kept on the model next to the flat weight buffer and shared by every batch
    size) and the DataLoader's pointers / RNG state, so a resumed run continues the same trajectory on the same data."""
    import torch
    state = {"weights": {k: v.detach().cpu() for k, v in model.weights.items()}, "next_step": int(next_step),
             "adam_m": model.adam_m.cpu(), "adam_v": model.adam_v.cpu(), "adam_t": int(model.adam_step[0]),
             "loader": loader_state}
    torch.save(state, path)


def load_checkpoint(model, path, data_loader=None):
    import torch
    state = torch.load(path, map_location="cpu", weights_only=False)
    for k, v in state["weights"].items():
        model.weights[k].copy_(v)
    if "adam_m" in state:
        model.adam_m.copy_(state["adam_m"])
        model.adam_v.copy_(state["adam_v"])
        model.adam_step[0] = int(state["adam_t"])
    if data_loader is not None and state.get("loader") is not None:
        data_loader.set_state(state["loader"])
    return int(state.get("next_step", 0))


def _init_distributed(args):
    """Under torchrun (WORLD_SIZE > 1): one process per GPU, NCCL; every minibatch of `batch_size` scenes is sharded
    round-robin over the ranks, gradients are all-reduced inside DESIREModel.train_step.
    Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    args.device = "cuda:%d" % local
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.batch_size % world:
        raise ValueError("batch_size %d must be divisible by the number of ranks %d" % (args.batch_size, world))
    return rank, world


def _check_replicas_in_sync(model, world):
    """Data-parallel invariant: every rank applied the identical all-reduced gradient, so the flat weight buffers
    are bit-identical.  One 16-byte all-reduce per epoch; raises if a replica drifted."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    cs = model.flat_weights.double().sum().reshape(1)
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if float(lo) != float(hi):
        raise RuntimeError("replicas out of sync: weight checksum min %.17g max %.17g" % (float(lo), float(hi)))


def train(args):
    from desire_b200.dist import global_masked_cost
    from desire_b200.model.model import DESIREModel
    from desire_b200.utils.data_loader import DataLoader
    import torch
    import torch.distributed as dist

    rank, world = _init_distributed(args)
    if args.raw_pixels:
        norm = None
        if args.ioc_iters > 0 and rank == 0:
            print("warning: --raw_pixels with ioc_iters > 0: the scene gather clamps every sample to the map border and no "
                  "neighbour falls inside the log-polar radii [r_min, r_max) = unit-square fractions")
    else:
        norm = (args.norm_w, args.norm_h) if args.norm_w > 0 and args.norm_h > 0 else "auto"
    lkw = dict(data_dir=args.data_dir, pred_length=args.pred_length, clip=args.clip_objects, normalize=norm,
               seed=args.seed, fix_id0=args.fix_id0)     # the SAME seed on every rank: all ranks cut the same minibatches
    if world > 1:
        # rank 0 alone (re)builds the preprocessed cache; the others read it after the barrier
        if rank == 0:
            DataLoader(args.batch_size, args.seq_length, args.max_num_obj, args.leave_dataset, preprocess=False, **lkw)
        dist.barrier()
    data_loader = DataLoader(args.batch_size, args.seq_length, args.max_num_obj, args.leave_dataset, preprocess=False, **lkw)
    if rank == 0 and norm == "auto":
        print("coordinates divided by the data extent {} x {} (pass --norm_w/--norm_h to fix it)".format(*data_loader.normalize))
    os.makedirs(args.save_dir, exist_ok=True)
    if rank == 0:
        with open(os.path.join(args.save_dir, 'config.pkl'), 'wb') as fh:     # train.py:102-103
            pickle.dump(args, fh)

    model = DESIREModel(args, device=args.device, seed=args.seed)      # same seed => identical weights on every rank
    model.batch_size = args.batch_size // world
    start_step = 0
    if args.resume:
        start_step = load_checkpoint(model, args.resume, data_loader)
        print("resumed from {} at step {}".format(args.resume, start_step))
    losses = []
    shard = (rank, world) if world > 1 else None
    for epoch in range(args.num_epochs):
        model.learning_rate = args.learning_rate * (args.decay_rate ** epoch)   # train.py:122-126
        nb = data_loader.num_batches if not args.max_batches else min(args.max_batches, data_loader.num_batches)
        first = 0
        if start_step > epoch * data_loader.num_batches:
            # resuming inside this epoch: the checkpoint restored the loader's pointers / RNG at `start_step`
            first = min(nb, start_step - epoch * data_loader.num_batches)
            if first >= nb:
                continue
        else:
            data_loader.reset_batch_pointer()
        if args.prefetch > 0:
            batches = data_loader.prefetch_epoch(nb - first, args.scene_size, depth=args.prefetch, shard=shard)
        else:
            batches = _inline_batches(data_loader, nb - first, args.scene_size, shard)
        start = time.time()
        for batch, (x, y, scene, _dval, lstate) in enumerate(batches, start=first):
            step = epoch * data_loader.num_batches + batch
            if args.optimize:
                c = model.train_step(x, y, eps=None, scene=scene, seed=args.seed + epoch * 100003 + batch + 7919 * rank)
                # cost over the whole minibatch: sum_ranks(cost_r * n_r) / sum_ranks(n_r)   (model/model.py:376)
                loss_batch = float(global_masked_cost(c[0] * c[1], c[1]))
            else:
                out = model.forward(x, y, eps=None, scene=scene, seed=args.seed + epoch * 100003 + batch)
                loss_batch = float(out["cost"])          # masked mean over existing agents (model.py:351-376)
            torch.cuda.synchronize()
            end = time.time()
            losses.append(loss_batch)
            if rank == 0:
                print("{}/{} (epoch {}), train_loss = {:.3f}, time/batch = {:.3f}"
                      .format(step, args.num_epochs * data_loader.num_batches, epoch, loss_batch, end - start))
                sys.stdout.flush()
            if rank == 0 and step % args.save_every == 0 and step > 0:    # train.py:197-207
                checkpoint_path = os.path.join(args.save_dir, 'social_model.ckpt')
                save_checkpoint(model, "%s-%d" % (checkpoint_path, step), step + 1, lstate)
                print("model saved to {}".format(checkpoint_path))
                sys.stdout.flush()
            start = time.time()
        _check_replicas_in_sync(model, world)
    return losses


def _inline_batches(data_loader, n, scene_size, shard):
    """--prefetch 0: each minibatch is built inside the step loop (the reference's order of work, train.py:140)."""
    from desire_b200.utils.data_loader import DataLoader
    for _ in range(n):
        xval, yval, dval = data_loader.next_batch()
        x = DataLoader.to_model_layout(xval)         # [B,N,Tp,3] agent-major (the transpose train.py:158-173 forgot)
        y = DataLoader.to_model_layout(yval)
        scene = data_loader.scene_images(dval, scene_size)   # reference.jpg next to the CSV, blank if absent
        if shard is not None:
            mine = list(range(shard[0], x.shape[0], shard[1]))
            x, y, scene = np.ascontiguousarray(x[mine]), np.ascontiguousarray(y[mine]), np.ascontiguousarray(scene[mine])
            dval = [dval[j] for j in mine]
        yield x, y, scene, dval, data_loader.state()


if __name__ == '__main__':
    main()
