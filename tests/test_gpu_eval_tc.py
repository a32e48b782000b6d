"""tcgen05 (tensor-core) full-rank evaluation: positions must equal the exact fp32 kernel bit for bit, which in turn
equals the oracle (tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

from oracle import apr_oracle as O

pytestmark = pytest.mark.gpu


def _case(rng, U, I, d, scale):
    P = (rng.randn(U + 1, d) * scale).astype(np.float32)
    Q = (rng.randn(I + 1, d) * scale).astype(np.float32)
    Q[I // 2] = Q[1]
    Q[I // 3] = Q[1]
    train = [sorted(set(rng.randint(0, I, rng.randint(0, 30)).tolist())) for _ in range(U)]
    test = rng.randint(0, I + 1, U).astype(np.int32)
    test[0], test[1] = 1, I // 2
    return P, Q, train, test


@pytest.mark.timeout(300)
@pytest.mark.parametrize("U,I,d,scale", [(130, 1000, 64, 1.0), (300, 5000, 128, 1.0), (257, 3333, 128, 0.01),
                                         (100, 2000, 256, 0.3), (64, 700, 8, 1.0), (50, 300, 200, 1.0)])
def test_tc_positions_equal_exact(cuda_device, U, I, d, scale):
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    rng = np.random.RandomState(U + I + d)
    P, Q, train, test = _case(rng, U, I, d, scale)
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    dev = cuda_device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    args = [t(P, torch.float32), t(Q, torch.float32), t(np.arange(U, dtype=np.int32), torch.int32), t(test, torch.int32), 0, I,
            t(ptr, torch.int64), t(idx, torch.int32)]
    exact, _, _ = engine.eval_fullrank(*args, 0, exact=True)
    got, n_amb = engine.eval_fullrank_tc(*args)
    assert torch.equal(got, exact), (got - exact).abs().max().item()
    assert 0 <= n_amb <= U * max(256, I // 256)
    # and a few users straight against the oracle
    for u in range(0, U, max(1, U // 7)):
        p, _, _, _ = O.eval_fullrank_user(P, Q, u, int(test[u]), train[u], I, 1)
        assert p == int(got[u])
    # item-sharded use: two ranges accumulate to the same counts
    acc = torch.zeros(U, dtype=torch.int32, device=dev)
    mid = (I // 2) // 128 * 128 + 5
    a2 = list(args)
    a2[4], a2[5] = 0, mid
    engine.eval_fullrank_tc(*a2, position=acc)
    a2[4], a2[5] = mid, I
    engine.eval_fullrank_tc(*a2, position=acc)
    assert torch.equal(acc, exact)


def _topk_merged(neg_ids, p, held_out, K):
    """library contract -> reference list: the top-k of the NEGATIVES, the held-out item entering at rank `position`
    (it loses every tie: utils.py:215, evaluation.py:59,73)."""
    neg = [t for t in neg_ids if t >= 0]
    return (neg[:p] + [held_out] + neg[p:] if p < K else neg)[:K]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("U,I,d,K,scale,fat", [(130, 1500, 64, 10, 1.0, 0), (300, 5000, 128, 100, 1.0, 3),
                                               (257, 20000, 128, 100, 0.01, 0), (100, 9000, 256, 128, 0.3, 2),
                                               (64, 1100, 8, 5, 1.0, 1), (50, 2500, 200, 100, 1.0, 0),
                                               (140, 40000, 64, 10, 1.0, 0)])
def test_tc_topk_equals_exact(cuda_device, U, I, d, K, scale, fat):
    """Top-k ids AND scores of the tensor-core path (group-maxima threshold -> candidate pass -> exact re-scoring ->
    selection, with the exact per-user kernel for the users the filter cannot serve) are bit-identical to the exact
    CUDA-core kernel for every user, and to the oracle on a sample.  `fat` users carry train lists of ~I/2 items: they
    exceed the number of group maxima and must take the exact per-user kernel."""
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    rng = np.random.RandomState(U + I + d + K)
    P, Q, train, test = _case(rng, U, I, d, scale)
    for f in range(fat):
        train[5 + 7 * f] = sorted(set(rng.randint(0, I, I // 2).tolist()))
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    dev = cuda_device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    args = [t(P, torch.float32), t(Q, torch.float32), t(np.arange(U, dtype=np.int32), torch.int32), t(test, torch.int32), 0, I,
            t(ptr, torch.int64), t(idx, torch.int32)]
    pos_e, ids_e, sc_e = engine.eval_fullrank(*args, K, exact=True)
    pos_t, ids_t, sc_t, info = engine.eval_fullrank_tc(*args, k_top=K)
    assert torch.equal(pos_t, pos_e)
    assert torch.equal(ids_t, ids_e), (ids_t != ids_e).nonzero()[:5]
    assert torch.equal(sc_t.view(torch.int32), sc_e.view(torch.int32))
    assert info["exact_fallback_users"] >= fat
    assert info["exact_fallback_users"] < max(fat + 1, U // 2), info    # the filter serves the ordinary users
    ids = ids_t.cpu().numpy()
    for u in list(range(0, U, max(1, U // 9))) + [5]:
        p, _, tids, _ = O.eval_fullrank_user(P, Q, u, int(test[u]), train[u], I, K)
        assert _topk_merged(ids[u].tolist(), p, int(test[u]), K) == tids.tolist()[:K], u
    # the same through engine.eval_fullrank's default route (exact=False -> tensor cores)
    pos_r, ids_r, sc_r = engine.eval_fullrank(*args, K)
    assert torch.equal(pos_r, pos_e) and torch.equal(ids_r, ids_e)
    # item-sharded: two ranges (local top-k each) + the K10 merge kernel == the single range
    mid = (I // 2) // 128 * 128
    a2 = list(args)
    a2[4], a2[5] = 0, mid
    _, i0, s0, _ = engine.eval_fullrank_tc(*a2, k_top=K)
    a2[4], a2[5] = mid, I
    _, i1, s1, _ = engine.eval_fullrank_tc(*a2, k_top=K)
    mi, ms = engine.topk_merge(torch.cat([i0, i1], dim=1).contiguous(), torch.cat([s0, s1], dim=1).contiguous(), K)
    assert torch.equal(mi, ids_e) and torch.equal(ms.view(torch.int32), sc_e.view(torch.int32))


def test_topk_merge_kernel_matches_host_merge(cuda_device):
    from apr_b200 import engine
    from apr_b200.distributed import merge_topk
    rng = np.random.RandomState(5)
    n, m, k = 700, 8 * 100, 100
    ids = rng.permutation(n * m).reshape(n, m).astype(np.int32) % 50000
    sc = rng.randint(0, 40, (n, m)).astype(np.float32) / 8        # many ties: the id order decides
    pad = rng.rand(n, m) < 0.3
    pad[3] = True                                                  # a user with nothing at all
    pad[4, 5:] = True                                              # fewer than k entries
    ids[pad], sc[pad] = -1, -np.inf
    # duplicates of one id inside a row cannot occur across item shards: make the ids of a row unique
    for u in range(n):
        ok = ~pad[u]
        ids[u, ok] = rng.choice(50000, int(ok.sum()), replace=False)
    wi, wsc = merge_topk(ids, sc, k)
    gi, gs = engine.topk_merge(torch.from_numpy(ids).to(cuda_device), torch.from_numpy(sc).to(cuda_device), k)
    assert np.array_equal(gi.cpu().numpy(), wi)
    assert np.array_equal(gs.cpu().numpy().view(np.uint32), wsc.view(np.uint32))


def test_tc_q_image_cache_and_sharded_table(cuda_device):
    """(a) cache_q: the item-operand image is built once and found again by the next user tile; a table write
    (engine.touch / torch in-place op) invalidates it.  (b) a rank that holds ONLY its shard of Q (rows [lo, hi)) plus
    externally supplied held-out scores produces the same counts / top-k as the whole-table call on that range."""
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    U, I, d, K = 260, 6000, 64, 10
    rng = np.random.RandomState(21)
    P, Q, train, test = _case(rng, U, I, d, 1.0)
    test = np.minimum(test, I - 1).astype(np.int32)
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    dev = cuda_device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    tP, tQ = t(P, torch.float32), t(Q, torch.float32)
    users, ttest, tptr, tidx = t(np.arange(U, dtype=np.int32), torch.int32), t(test, torch.int32), t(ptr, torch.int64), t(idx, torch.int32)
    want, wi, _ = engine.eval_fullrank(tP, tQ, users, ttest, 0, I, tptr, tidx, K, exact=True)
    for rep in range(2):                                   # second call hits the cached image
        got, gi, _, _ = engine.eval_fullrank_tc(tP, tQ, users, ttest, 0, I, tptr, tidx, k_top=K, cache_q=True)
        assert torch.equal(got, want) and torch.equal(gi, wi)
    tQ[7] += 1.0                                            # torch in-place op: version tag changes, image is rebuilt
    want2, wi2, _ = engine.eval_fullrank(tP, tQ, users, ttest, 0, I, tptr, tidx, K, exact=True)
    got2, gi2, _, _ = engine.eval_fullrank_tc(tP, tQ, users, ttest, 0, I, tptr, tidx, k_top=K, cache_q=True)
    assert torch.equal(got2, want2) and torch.equal(gi2, wi2)
    engine.release_eval_workspace()
    # (b) shard [lo, hi) only
    lo, hi = 2048, 4480
    spos = engine.score_pairs(tP, tQ, users, ttest)
    q_shard = tQ[lo:hi].clone()
    part_w, pi_w, _ = engine.eval_fullrank(tP, tQ, users, ttest, lo, hi, tptr, tidx, K, exact=True)
    part, pi, _, _ = engine.eval_fullrank_tc(tP, q_shard, users, None, lo, hi, tptr, tidx, k_top=K, spos=spos, q_row_offset=lo)
    assert torch.equal(part, part_w) and torch.equal(pi, pi_w)


def test_tc_eval_tiles_large_user_sets(cuda_device, monkeypatch):
    """More users than one library call takes: engine.eval_fullrank_tc tiles them (cached item image) and returns the same
    positions / top-k as the exact kernel on the whole set."""
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    monkeypatch.setattr(engine, "EVAL_USER_TILE", 96)
    U, I, d, K = 300, 3000, 64, 10
    rng = np.random.RandomState(31)
    P, Q, train, test = _case(rng, U, I, d, 1.0)
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    dev = cuda_device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    args = [t(P, torch.float32), t(Q, torch.float32), t(np.arange(U, dtype=np.int32), torch.int32), t(test, torch.int32), 0, I,
            t(ptr, torch.int64), t(idx, torch.int32)]
    pos_e, ids_e, sc_e = engine.eval_fullrank(*args, K, exact=True)
    pos_t, ids_t, sc_t, info = engine.eval_fullrank_tc(*args, k_top=K)
    assert torch.equal(pos_t, pos_e) and torch.equal(ids_t, ids_e) and torch.equal(sc_t.view(torch.int32), sc_e.view(torch.int32))
    pos_0, n_amb = engine.eval_fullrank_tc(*args)
    assert torch.equal(pos_0, pos_e) and n_amb >= 0
    engine.release_eval_workspace()
