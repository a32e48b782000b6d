"""Parity at BASELINE.json's full evaluation shapes (configs[1], [2] and a tile of [4]); the oracle is consulted on a
sample of users, the rest is covered by cross-checks between independent CUDA paths and by size-independent properties.
(configs[3], the full-size training step, lives in test_gpu_parity.py.)"""
import numpy as np
import pytest
import torch

from oracle import apr_oracle as O

pytestmark = pytest.mark.gpu


def _t(a, dt, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)


def test_config2_pinterest_shape_sampled_eval(cuda_device):
    """configs[1]: 55 187 users, 9 916 items, d=64, 99 sampled negatives + the held-out item per user (He protocol,
    utils.py:244-254).  Oracle on 300 users; all users: position in [0, 99] and equal to a torch recount of the scores
    the kernel itself returns (bit-exact scores are checked against the oracle on the sample)."""
    from apr_b200 import engine
    U, I, d, C = 55187, 9916, 64, 100
    rng = np.random.RandomState(2019)
    P = (rng.randn(U, d) * 0.1).astype(np.float32)
    Q = (rng.randn(I, d) * 0.1).astype(np.float32)
    cand = rng.randint(0, I, (U, C)).astype(np.int32)       # last column = held-out item
    ptr = np.arange(0, (U + 1) * C, C, dtype=np.int64)
    users = np.arange(U, dtype=np.int32)
    pos, sc = engine.eval_candidates(_t(P, torch.float32, cuda_device), _t(Q, torch.float32, cuda_device),
                                     _t(users, torch.int32, cuda_device), _t(ptr, torch.int64, cuda_device),
                                     _t(cand.reshape(-1), torch.int32, cuda_device), want_scores=True)
    sc2 = sc.reshape(U, C)
    recount = (sc2[:, :-1] >= sc2[:, -1:]).sum(dim=1).to(torch.int32)
    assert torch.equal(pos, recount)
    pos = pos.cpu().numpy()
    assert pos.min() >= 0 and pos.max() <= C - 1
    sample = rng.choice(U, 300, replace=False)
    sch = sc2.cpu().numpy()
    for u in sample:
        assert pos[u] == O.eval_candidates_position(P, Q, int(u), cand[u].tolist())
        want = O.score_pairs(P, Q, np.full(C, u), cand[u])
        assert np.array_equal(sch[u].view(np.uint32), want.view(np.uint32))


def test_config3_yelp_shape_fullrank_k100(cuda_device):
    """configs[2]: 25 677 users x 25 815 items, d=128, full rank.  Exact kernel == tensor-core path for every user;
    oracle positions and top-100 ids for 40 users."""
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    U, I, d, K = 25677, 25815, 128, 100
    rng = np.random.RandomState(7)
    P = (rng.randn(U, d) / np.sqrt(d)).astype(np.float32)
    Q = (rng.randn(I, d) / np.sqrt(d)).astype(np.float32)
    test = rng.randint(0, I, U).astype(np.int32)
    cnt = rng.randint(5, 60, U)
    flat = rng.randint(0, I, int(cnt.sum())).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(cnt)])
    train = [sorted(set(flat[off[u]:off[u + 1]].tolist())) for u in range(U)]
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    a = [_t(P, torch.float32, cuda_device), _t(Q, torch.float32, cuda_device),
         _t(np.arange(U, dtype=np.int32), torch.int32, cuda_device), _t(test, torch.int32, cuda_device), 0, I,
         _t(ptr, torch.int64, cuda_device), _t(idx, torch.int32, cuda_device)]
    pos_e, ids, _ = engine.eval_fullrank(*a, K, exact=True)
    pos_t, n_amb = engine.eval_fullrank_tc(*a)
    assert torch.equal(pos_e, pos_t)
    assert 0 < n_amb < U * 256
    pos = pos_e.cpu().numpy()
    ids = ids.cpu().numpy()
    assert pos.min() >= 0 and pos.max() < I
    for u in rng.choice(U, 40, replace=False):
        p, _, tids, _ = O.eval_fullrank_user(P, Q, int(u), int(test[u]), train[u], I, K)
        assert pos[u] == p
        neg = [t for t in ids[u].tolist() if t >= 0]
        merged = neg[:p] + [int(test[u])] + neg[p:] if p < K else neg
        assert merged[:K] == tids.tolist()[:K]


def test_config5_tile_tc_equals_exact(cuda_device):
    """A tile of configs[4]: 512 users x 10 000 000 items, d=256 -- the tensor-core count must equal the exact fp32
    kernel for every user at the full item count (10 M-wide rows of the ambiguous list, 78 125 item tiles per CTA row)."""
    from apr_b200 import engine
    U, I, d = 512, 10_000_000, 256
    g = torch.Generator(device=cuda_device)
    g.manual_seed(2019)
    P = torch.randn((U, d), device=cuda_device, generator=g) / d ** 0.5
    Q = torch.randn((I, d), device=cuda_device, generator=g) / d ** 0.5
    test = torch.randint(0, I, (U,), device=cuda_device, dtype=torch.int32, generator=g)
    ptr = torch.arange(0, U + 1, device=cuda_device, dtype=torch.int64)
    a = [P, Q, torch.arange(U, device=cuda_device, dtype=torch.int32), test, 0, I, ptr, test.clone()]
    pos_t, n_amb = engine.eval_fullrank_tc(*a)
    pos_e, _, _ = engine.eval_fullrank(*a, 0, exact=True)
    assert torch.equal(pos_t, pos_e)
    assert n_amb > 0
    # oracle on one user against a 1/64 slice of the items, no exclusions (the held-out item counts itself if inside)
    u = 3
    Ph, sl = P[u:u + 1].cpu().numpy(), slice(0, I // 64)
    Qh = Q[sl].cpu().numpy()
    s = O.score_pairs(Ph, Qh, np.zeros(Qh.shape[0], np.int64), np.arange(Qh.shape[0]))
    sp = O.score_pairs(Ph, Q[int(test[u])][None].cpu().numpy(), np.zeros(1, np.int64), np.zeros(1, np.int64))[0]
    part, _, _ = engine.eval_fullrank(P[u:u + 1].contiguous(), Q, torch.zeros(1, dtype=torch.int32, device=cuda_device),
                                      test[u:u + 1].contiguous(), 0, I // 64, torch.zeros(2, dtype=torch.int64, device=cuda_device),
                                      torch.zeros(1, dtype=torch.int32, device=cuda_device), 0, exact=True)
    assert int(part.item()) == int((s >= sp).sum())


def test_tc_path_with_non_finite_rows(cuda_device):
    """Inf / NaN rows (a diverged run).  Non-finite ITEM rows: the tensor-core path must agree with the exact kernel
    without overflowing its ambiguous list (the poisoned block falls back to per-column bounds, so only the bad column is
    re-scored).  A non-finite HELD-OUT score makes every pair of that user undecidable: the library either still agrees or
    reports the list overflow that utils.eval_positions answers with the exact kernel -- never a wrong position."""
    from apr_b200 import engine
    U, I, d = 300, 4096, 64
    rng = np.random.RandomState(11)
    P = (rng.randn(U, d) / 8).astype(np.float32)
    Q = (rng.randn(I, d) / 8).astype(np.float32)
    Q[100, 3] = np.inf
    Q[777, :] = np.nan
    Q[2048, 0] = -np.inf
    Q[4000, 5] = 3e38          # finite, but the squared norm overflows
    test = rng.randint(0, I, U).astype(np.int32)
    test[np.isin(test, [100, 777, 2048, 4000])] = 5
    ptr = np.arange(U + 1, dtype=np.int64)

    def args(test):
        return [_t(P, torch.float32, cuda_device), _t(Q, torch.float32, cuda_device),
                _t(np.arange(U, dtype=np.int32), torch.int32, cuda_device), _t(test, torch.int32, cuda_device), 0, I,
                _t(ptr, torch.int64, cuda_device), _t(test, torch.int32, cuda_device)]

    a = args(test)
    pos_e, _, _ = engine.eval_fullrank(*a, 0, exact=True)
    pos_t, n_amb = engine.eval_fullrank_tc(*a)
    assert torch.equal(pos_e, pos_t)
    assert n_amb < 40 * U                      # ~4 bad columns per user + the usual band, not whole blocks

    test2 = test.copy()
    test2[0], test2[1], test2[150] = 100, 777, 2048   # held-out item itself non-finite
    a = args(test2)
    pos_e, _, _ = engine.eval_fullrank(*a, 0, exact=True)
    try:
        pos_t, _ = engine.eval_fullrank_tc(*a)
    except RuntimeError as e:
        assert "overflow" in str(e)
    else:
        assert torch.equal(pos_e, pos_t)
