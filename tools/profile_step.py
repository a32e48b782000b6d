#!/usr/bin/env python
"""Profiling driver for the training step (run under ncu by tools/gpu_profile.sh): BASELINE.json configs[3] tables
(10M x 2M, d from argv), two calls of 20 steps of 65 536 triples each -- the first builds the executable graphs and warms
up, the second is the one to read in the launch list."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apr_b200 import engine  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 20
U, I, B = 10_000_000, 2_000_000, 65536
dev = engine.require_cuda()
P = torch.empty((U, d), device=dev)
Q = torch.empty((I, d), device=dev)
engine.init_truncated_normal(P, 0.01, 2019, 0)
engine.init_truncated_normal(Q, 0.01, 2019, 1)
aP, aQ = torch.full_like(P, 0.1), torch.full_like(Q, 0.1)
ws = engine.TrainWorkspace(S, B, d, dev)
rng = np.random.default_rng(2019)
for rep in range(2):
    u, i, j = [torch.from_numpy(rng.integers(0, n, size=(S, B), dtype=np.int32)).to(dev) for n in (U, I, I)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.train_steps(P, Q, aP, aQ, u, i, j, 0.05, 0.0, 1.0, 0.5, 1, ws, mode=0)
    e1.record()
    torch.cuda.synchronize()
    print("call %d: %.1f us/step" % (rep, 1e3 * e0.elapsed_time(e1) / S))
cnt = ws.unique_counts(S).astype(np.int64)
print("bytes_per_step_model %.0f" % (float(16 * d * cnt.sum() + 12 * B * S) / S))
