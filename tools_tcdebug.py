import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from apr_b200 import engine, _lib
dev = torch.device('cuda')
U, I, d = 25677, 25815, 128
g = torch.Generator(device=dev); g.manual_seed(2019)
P = torch.randn((U, d), device=dev, generator=g) / d ** 0.5
Q = torch.randn((I + 1, d), device=dev, generator=g) / d ** 0.5
test = torch.randint(0, I, (U,), device=dev, dtype=torch.int32, generator=g)
ptr = torch.arange(0, U + 1, device=dev, dtype=torch.int64); idx = test.clone()
users = torch.arange(U, device=dev, dtype=torch.int32)
a = [P, Q, users, test, 0, I, ptr, idx]
def tc(tag):
    try:
        pt, namb = engine.eval_fullrank_tc(*a)
        print(tag, "ambiguous", namb)
    except Exception as e:
        print(tag, "EXC", e)
tc("tc before exact")
pe, _, _ = engine.eval_fullrank(*a, 0, exact=True); torch.cuda.synchronize()
tc("tc after 1 exact")
junk = torch.full((200_000_000,), 3.0e38, device=dev); del junk     # poison the allocator's free blocks
tc("tc after poison")
