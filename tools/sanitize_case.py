#!/usr/bin/env python
"""The small parity shapes run under compute-sanitizer (one --tool per gpurun call; tools/gpu_sanitize.sh).

Covers the kernels that rely on CAS hash tables, red.global.add.v4, hand-rolled mbarriers and grid / cluster barriers:
index preparation + step kernels in modes 0 (with and without graph replay), 1 and 2, BPR / APR / --adv random / adver 3,
the exact and the tcgen05 evaluation incl. the top-k passes, the merge kernel and the sampler.  Every result is also
checked against the oracle, so a clean sanitizer log belongs to a run that computed the right thing.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apr_b200 import engine  # noqa: E402
from apr_b200.Dataset import build_sorted_csr  # noqa: E402
from oracle import apr_oracle as O  # noqa: E402

which = set(sys.argv[1:]) or {"train", "eval", "sampler"}
dev = engine.require_cuda()
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
rng = np.random.RandomState(0)


def close(got, ref, tag):
    err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
    assert err <= 1e-5, (tag, err)


if "train" in which:
    for (U, I, d, S, B) in ((300, 200, 64, 3, 512), (60, 50, 24, 2, 100), (4000, 3000, 128, 2, 2048)):
        P = (rng.randn(U, d) * 0.1).astype(np.float32)
        Q = (rng.randn(I, d) * 0.1).astype(np.float32)
        u, i, j = [rng.randint(0, n, (S, B)).astype(np.int32) for n in (U, I, I)]
        for adver in (0, 1, 3):
            rP, rQ = P.copy(), Q.copy()
            raP, raQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
            for s in range(S):
                O.apr_step(rP, rQ, raP, raQ, u[s], i[s], j[s], 0.05, 0.01, 1.0, 0.5, adver)
            for mode in (0, 1, 2):
                tP, tQ = t(P, torch.float32), t(Q, torch.float32)
                aP, aQ = torch.full_like(tP, 0.1), torch.full_like(tQ, 0.1)
                ws = engine.TrainWorkspace(S, B, d, dev)
                st = torch.zeros((S, 2), dtype=torch.float32, device=dev)
                engine.train_steps(tP, tQ, aP, aQ, t(u, torch.int32), t(i, torch.int32), t(j, torch.int32), 0.05, 0.01, 1.0, 0.5,
                                   adver, ws, mode=mode, stats=st)
                torch.cuda.synchronize()
                close(tP.cpu().numpy(), rP, ("P", U, d, adver, mode))
                close(tQ.cpu().numpy(), rQ, ("Q", U, d, adver, mode))
        # --adv random
        rP, rQ = P.copy(), Q.copy()
        raP, raQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
        for s in range(S):
            O.apr_step_random(rP, rQ, raP, raQ, u[s], i[s], j[s], 0.05, 0.01, 1.0, 0.5, 2019, 3 + s)
        tP, tQ = t(P, torch.float32), t(Q, torch.float32)
        aP, aQ = torch.full_like(tP, 0.1), torch.full_like(tQ, 0.1)
        ws = engine.TrainWorkspace(S, B, d, dev)
        engine.train_steps_random(tP, tQ, aP, aQ, t(u, torch.int32), t(i, torch.int32), t(j, torch.int32), 0.05, 0.01, 1.0, 0.5, ws,
                                  2019, 3)
        torch.cuda.synchronize()
        close(tP.cpu().numpy(), rP, ("P random", U, d))
        close(tQ.cpu().numpy(), rQ, ("Q random", U, d))
    print("train: ok (modes 0/1/2 x BPR/APR/adver3, --adv random; graph replay %s)" % os.environ.get("APR_GRAPH", "1"))

if "eval" in which:
    for (U, I, d, K) in ((130, 1500, 64, 10), (90, 2600, 128, 100), (40, 1200, 40, 5)):
        P = rng.randn(U + 1, d).astype(np.float32)
        Q = rng.randn(I + 1, d).astype(np.float32)
        Q[I // 2] = Q[1]
        train = [sorted(set(rng.randint(0, I, rng.randint(0, 30)).tolist())) for _ in range(U)]
        train[5] = sorted(set(rng.randint(0, I, I // 2).tolist()))        # takes the exact per-user kernel
        test = rng.randint(0, I, U).astype(np.int32)
        ptr, idx = build_sorted_csr([train[k] + [int(test[k])] for k in range(U)])
        a = [t(P, torch.float32), t(Q, torch.float32), t(np.arange(U, dtype=np.int32), torch.int32), t(test, torch.int32), 0, I,
             t(ptr, torch.int64), t(idx, torch.int32)]
        pe, ie, se = engine.eval_fullrank(*a, K, exact=True)
        pt, it, stc, info = engine.eval_fullrank_tc(*a, k_top=K)
        assert torch.equal(pe, pt) and torch.equal(ie, it) and torch.equal(se.view(torch.int32), stc.view(torch.int32)), (U, I, d)
        p0, _ = engine.eval_fullrank_tc(*a)
        assert torch.equal(p0, pe)
        mi, ms = engine.topk_merge(torch.cat([it, it], dim=1).contiguous(), torch.cat([stc, stc - 1.0], dim=1).contiguous(), K)
        for k in range(0, U, 17):
            p, _, _, _ = O.eval_fullrank_user(P, Q, k, int(test[k]), train[k], I, 1)
            assert p == int(pe[k])
    cand = rng.randint(0, 1200, (40, 20)).astype(np.int32)
    pos, _ = engine.eval_candidates(a[0], a[1], a[2], t(np.arange(0, 41 * 20, 20), torch.int64), t(cand.reshape(-1), torch.int32))
    torch.cuda.synchronize()
    print("eval: ok (exact, tcgen05 count, tcgen05 top-k, merge, candidates)")

if "sampler" in which:
    U, I = 300, 200
    lists = [sorted(set(rng.randint(0, I, rng.randint(1, 30)).tolist())) for _ in range(U)]
    pu = np.concatenate([[k] * len(l) for k, l in enumerate(lists)]).astype(np.int32)
    pi = np.concatenate(lists).astype(np.int32)
    cptr, cidx = O.build_csr(lists)
    for dns, fw in ((1, 0), (3, 0), (1, 4)):
        got = engine.sample_epoch(t(pu, torch.int32), t(pi, torch.int32), 64, I, t(cptr, torch.int64), t(cidx, torch.int32), 2019, 3,
                                  dns, fork_workers=fw)
        want = O.sample_epoch(pu, pi, 64, I, cptr, cidx, 2019, 3, dns, fork_workers=fw)
        assert np.array_equal(got[3].cpu().numpy(), want[3])
    W = torch.empty((50, 16), device=dev)
    engine.init_truncated_normal(W, 0.01, 2019, 0)
    la = engine.loss_acc(a[0], a[1], t(rng.randint(0, 40, (2, 64)).astype(np.int32), torch.int32),
                         t(rng.randint(0, 1200, (2, 64)).astype(np.int32), torch.int32),
                         t(rng.randint(0, 1200, (2, 64)).astype(np.int32), torch.int32)) if "eval" in which else None
    torch.cuda.synchronize()
    print("sampler: ok")
print("sanitize_case: all ok")
