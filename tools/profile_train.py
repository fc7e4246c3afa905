"""Run the train step un-graphed at the bench workload so ncu can list / capture single launches.

    python tools/profile_train.py [--steps 1] [--scenes 32] [--no-ioc]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from desire_b200.config import DesireConfig, init_params  # noqa: E402
from desire_b200.engine import TrainPath, flatten_params  # noqa: E402
from desire_b200.synthetic import make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--scenes", type=int, default=32)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--agents", type=int, default=60)
    ap.add_argument("--samples", type=int, default=20)
    ap.add_argument("--ioc-iters", type=int, default=2)
    ap.add_argument("--no-ioc", action="store_true")
    a = ap.parse_args()
    cfg = DesireConfig(d_dim=a.hidden, max_num_obj=a.agents, num_samples=a.samples, ioc_iters=a.ioc_iters)
    flat, views, offs = flatten_params(init_params(cfg, 1), "cuda:0")
    tp = TrainPath(cfg, flat, views, offs, a.scenes, train_ioc=not a.no_ioc)
    inp = [t.cuda() for t in make_batch(cfg, a.scenes, 100)]
    for _ in range(a.steps):
        tp.train_step(*inp, lr=1e-4, use_graph=False)
    torch.cuda.synchronize()
    print("done", float(tp.buf["cost"][0]), float(tp.buf["ioc_cost"][0]))


if __name__ == "__main__":
    main()
