"""Cut a small real-data fixture out of the reference's SDD annotations (data, not source):
the first FRAMES frames of data/bookstore/video0/annotations_processed.csv, same 4-row CSV format
(scripts/preprocess.py:30-34).  Written to tests/golden/sdd/bookstore/video0/annotations_processed.csv.

    python tools/make_sdd_fixture.py            # needs /root/reference (not available on the GPU box)
"""
import os
import sys

import numpy as np

SRC = "/root/reference/data/bookstore/video0/annotations_processed.csv"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "tests", "golden", "sdd", "bookstore", "video0", "annotations_processed.csv")
FRAMES = 40

data = np.loadtxt(SRC, delimiter=",", ndmin=2)
frames = np.unique(data[0])[:FRAMES]
keep = np.isin(data[0], frames)
sub = data[:, keep]
os.makedirs(os.path.dirname(DST), exist_ok=True)
with open(DST, "w") as fh:
    for r, fmt in zip(sub, ("%d", "%d", "%.1f", "%.1f")):
        fh.write(",".join(fmt % v for v in r) + "\n")
print(DST, sub.shape, "frames", frames[0], "..", frames[-1], "objects/frame", [int((sub[0] == f).sum()) for f in frames[:5]])
