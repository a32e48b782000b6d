import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from apr_b200 import engine
dev = torch.device('cuda')
U, I, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device=dev); g.manual_seed(2019)
P = torch.randn((U, d), device=dev, generator=g) / d ** 0.5
Q = torch.randn((I + 1, d), device=dev, generator=g) / d ** 0.5
test = torch.randint(0, I, (U,), device=dev, dtype=torch.int32, generator=g)
ptr = torch.arange(0, U + 1, device=dev, dtype=torch.int64); idx = test.clone()
users = torch.arange(U, device=dev, dtype=torch.int32)
for _ in range(2):
    pt, namb = engine.eval_fullrank_tc(P, Q, users, test, 0, I, ptr, idx)
torch.cuda.synchronize()
print("ok", namb)
