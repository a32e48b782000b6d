import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from apr_b200 import engine
dev = torch.device('cuda')
U, I, B, S = 10_000_000, 2_000_000, 65536, 64
for d in (64, 128):
    P = torch.randn(U, d, device=dev) * 0.01; Q = torch.randn(I, d, device=dev) * 0.01
    aP = torch.full_like(P, 0.1); aQ = torch.full_like(Q, 0.1)
    g = torch.Generator(device=dev); g.manual_seed(1)
    u = torch.randint(0, U, (S, B), device=dev, dtype=torch.int32, generator=g)
    i = torch.randint(0, I, (S, B), device=dev, dtype=torch.int32, generator=g)
    j = torch.randint(0, I, (S, B), device=dev, dtype=torch.int32, generator=g)
    ws = engine.TrainWorkspace(S, B, d, dev)
    for rep in range(3):
        engine.train_prepare(P, Q, u, i, j, ws); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        engine.train_run(P, Q, aP, aQ, u, i, j, 0.05, 0.0, 1.0, 0.5, 1, ws, mode=0)
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print("d=%d: host enqueue %.1f us/step, device %.1f us/step, wall %.1f us/step" % (d, (t1 - t0) / S * 1e6, e0.elapsed_time(e1) / S * 1e3, (t2 - t0) / S * 1e6))
    del P, Q, aP, aQ, ws
