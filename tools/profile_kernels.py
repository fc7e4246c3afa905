"""Run the hot path un-graphed at the bench workload so that `ncu -k regex:<kernel>` can capture
single launches (see profiles/README.md for the exact commands).

    python tools/profile_kernels.py [--passes 2] [--scenes 32]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from desire_b200.config import DesireConfig, init_params  # noqa: E402
from desire_b200.engine import HotPath  # noqa: E402
from desire_b200.synthetic import make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--scenes", type=int, default=32)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--agents", type=int, default=60)
    ap.add_argument("--samples", type=int, default=20)
    ap.add_argument("--ioc-iters", type=int, default=1)
    ap.add_argument("--serial", action="store_true", help="no parallel branches / IOC chains (full-size launches)")
    ap.add_argument("--pred-length", type=int, default=12)
    ap.add_argument("--scene-size", type=int, default=256)
    a = ap.parse_args()
    cfg = DesireConfig(d_dim=a.hidden, max_num_obj=a.agents, num_samples=a.samples, ioc_iters=a.ioc_iters,
                       pred_length=a.pred_length, scene_size=a.scene_size)
    hp = HotPath(cfg, init_params(cfg, 1), a.scenes)
    hp.serial = a.serial
    inp = [t.cuda() for t in make_batch(cfg, a.scenes, 100)]
    for _ in range(a.passes):
        hp.run(*inp)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
