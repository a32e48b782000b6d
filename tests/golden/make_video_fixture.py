"""Builds tests/golden/video_interactions.npz from the reference's only complete real dataset
(/root/reference/data/Video.{train,test}.rating: 256 094 train rows, 31 013 test rows).

Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_video_fixture.py
Stored as compact int32 arrays (the TSV's rating and timestamp columns are constant 1 and are dropped) plus the
Epoch-0 lines the reference logged for this dataset (out/janEval/Video_{apr,bpr}_*.out:3), which the statistical
known-answer test compares against.
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(path):
    u, i, r = [], [], []
    for line in open(path):
        a = line.split("\t")
        if len(a) < 3:
            continue
        u.append(int(a[0])); i.append(int(a[1])); r.append(float(a[2]))
    return np.asarray(u, np.int32), np.asarray(i, np.int32), np.asarray(r, np.float32)


def main():
    tu, ti, tr = read(os.path.join(REF, "data/Video.train.rating"))
    eu, ei, er = read(os.path.join(REF, "data/Video.test.rating"))
    assert (tr == 1).all() and (er == 1).all()
    logs = []
    d = os.path.join(REF, "out/janEval")
    for fn in sorted(os.listdir(d)):
        if re.match(r"Video_(apr|bpr)_.*\.out$", fn):
            for line in open(os.path.join(d, fn)):
                if line.startswith("Epoch 0 "):
                    m = re.search(r"HR = ([\d.]+), NDCG = ([\d.]+) ACC = ([\d.]+) ACC_adv = ([\d.]+) \[[\d.]+s\], "
                                  r"\|P\|=([\d.]+), \|Q\|=([\d.]+)", line)
                    logs.append([float(x) for x in m.groups()])
    logs = np.asarray(logs, np.float64)  # columns: HR@100, NDCG@100, ACC before, ACC after, |P|, |Q|
    out = os.path.join(HERE, "video_interactions.npz")
    np.savez_compressed(out, train_u=tu, train_i=ti, test_u=eu, test_i=ei, epoch0_logged=logs)
    print(out, os.path.getsize(out), "bytes; epoch-0 log rows:", logs.tolist())


if __name__ == "__main__":
    sys.exit(main())
