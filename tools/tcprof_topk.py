#!/usr/bin/env python
"""Profiling driver for the tensor-core evaluation with top-100 (run under ncu by tools/gpu_profile.sh):
16 384 users x 524 288 items, d = 128 (the shape of profiles/r1k_*), counting pass with group maxima, threshold, candidate
pass, exact re-scoring, selection."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apr_b200 import engine  # noqa: E402

dev = engine.require_cuda()
U, I, d, K = 16384, 524288, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 100
g = torch.Generator(device=dev)
g.manual_seed(2019)
P = torch.randn((U, d), device=dev, generator=g) / d ** 0.5
Q = torch.randn((I, d), device=dev, generator=g) / d ** 0.5
test = torch.randint(0, I, (U,), device=dev, dtype=torch.int32, generator=g)
users = torch.arange(U, device=dev, dtype=torch.int32)
ptr = torch.arange(0, U + 1, device=dev, dtype=torch.int64)
for rep in range(2):
    pos, ids, sc, info = engine.eval_fullrank_tc(P, Q, users, test, 0, I, ptr, test.clone(), k_top=K)
    torch.cuda.synchronize()
print(info, "hr10 %.5f" % float((pos < 10).float().mean()))
