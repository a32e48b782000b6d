import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from apr_b200 import engine
dev = torch.device('cuda')
def run(U, I, d, reps=3, exact_too=True):
    g = torch.Generator(device=dev); g.manual_seed(2019)
    P = torch.randn((U, d), device=dev, generator=g) / d ** 0.5
    Q = torch.randn((I + 1, d), device=dev, generator=g) / d ** 0.5
    test = torch.randint(0, I, (U,), device=dev, dtype=torch.int32, generator=g)
    ptr = torch.arange(0, U + 1, device=dev, dtype=torch.int64)          # exclusion = the held-out item only
    idx = test.clone()
    users = torch.arange(U, device=dev, dtype=torch.int32)
    a = [P, Q, users, test, 0, I, ptr, idx]
    res = {}
    if exact_too:
        engine.eval_fullrank(*a, 0, exact=True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): pe, _, _ = engine.eval_fullrank(*a, 0, exact=True)
        e1.record(); torch.cuda.synchronize()
        res['exact_ms'] = e0.elapsed_time(e1) / reps
    pt, namb = engine.eval_fullrank_tc(*a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): pt, _ = engine.eval_fullrank_tc(*a, check=False)
    e1.record(); torch.cuda.synchronize()
    res['tc_ms'] = e0.elapsed_time(e1) / reps
    if exact_too: res['equal'] = bool(torch.equal(pt, pe))
    flops_issued = 2.0 * U * I * 3 * d
    res.update(U=U, I=I, d=d, n_amb=namb, amb_per_user=namb / U, tc_users_per_s=U / res['tc_ms'] * 1e3,
               tc_issued_tflops=flops_issued / res['tc_ms'] / 1e9, useful_tflops=flops_issued / 3 / res['tc_ms'] / 1e9)
    if exact_too: res['exact_users_per_s'] = U / res['exact_ms'] * 1e3
    print(res, flush=True)
run(25677, 25815, 128)
run(16384, 2_000_000, 128, reps=1, exact_too=False)
run(4096, 10_000_000, 256, reps=1, exact_too=False)
