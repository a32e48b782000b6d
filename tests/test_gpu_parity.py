"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): sampled indices, rank positions and top-K ids bit-exact; losses, gradients and
embeddings within 1e-5 relative in fp32.  "Relative" for a tensor means |a-b| <= RTOL * max|ref| element-wise
(sums of signed terms cancel, so a per-element relative bound is not meaningful for fp32 atomics).
"""
import os

import numpy as np
import pytest
import torch

from oracle import apr_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(got, ref, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = float(np.abs(got - ref).max()) / scale
    assert err <= rtol, "max error %.3g of scale %.3g" % (err * scale, scale)


def _close_update(got, ref, before, rtol=RTOL):
    """The north_star bar on what a step CHANGES: |(got - before) - (ref - before)| <= rtol * max|ref - before|.
    (`_close` on the tables themselves is ~10x looser, because an update is ~1e-2 of a 0.1-scale table; the Adagrad
    kernels use MUFU.RSQ and __fdividef, so this is the check that bounds them.)"""
    got, ref, before = [np.asarray(a, dtype=np.float64) for a in (got, ref, before)]
    dg, dr = got - before, ref - before
    scale = max(float(np.abs(dr).max()), 1e-30)
    err = float(np.abs(dg - dr).max()) / scale
    assert err <= rtol, "update error %.3g of update scale %.3g" % (err * scale, scale)


def _dev(a, dtype, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


def _problem(rng, U, I, d, S, B, scale=0.1, zipf=False):
    P = (rng.randn(U, d) * scale).astype(np.float32)
    Q = (rng.randn(I, d) * scale).astype(np.float32)
    if zipf:
        w = 1.0 / np.arange(1, I + 1) ** 1.05
        w /= w.sum()
        i = rng.choice(I, size=(S, B), p=w)
        j = rng.choice(I, size=(S, B), p=w)
        wu = 1.0 / np.arange(1, U + 1)
        u = rng.choice(U, size=(S, B), p=wu / wu.sum())
    else:
        u = rng.randint(0, U, (S, B))
        i = rng.randint(0, I, (S, B))
        j = rng.randint(0, I, (S, B))
    return P, Q, u.astype(np.int32), i.astype(np.int32), j.astype(np.int32)


def _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, adver):
    P, Q = P.copy(), Q.copy()
    aP = np.full_like(P, 0.1)
    aQ = np.full_like(Q, 0.1)
    stats = []
    for s in range(u.shape[0]):
        stats.append(O.loss_acc(P, Q, u[s], i[s], j[s]))
        O.apr_step(P, Q, aP, aQ, u[s], i[s], j[s], lr, reg, reg_adv, eps, adver)
    return P, Q, aP, aQ, np.asarray(stats)


def _run_cuda_steps(dev, P, Q, u, i, j, lr, reg, reg_adv, eps, adver, mode):
    from apr_b200 import engine
    tP, tQ = _dev(P, torch.float32, dev), _dev(Q, torch.float32, dev)
    aP, aQ = torch.full_like(tP, 0.1), torch.full_like(tQ, 0.1)
    S, B = u.shape
    ws = engine.TrainWorkspace(S, B, P.shape[1], dev)
    stats = torch.zeros((S, 2), dtype=torch.float32, device=dev)
    engine.train_steps(tP, tQ, aP, aQ, _dev(u, torch.int32, dev), _dev(i, torch.int32, dev), _dev(j, torch.int32, dev), lr,
                       reg, reg_adv, eps, adver, ws, mode=mode, stats=stats)
    torch.cuda.synchronize()
    counts = ws.unique_counts(S)
    L = engine.train_layout(S, B, P.shape[1])
    _run_cuda_steps.last_npair = ws.buf[L["npair"]:L["npair"] + 4 * S].view(torch.int32).cpu().numpy().copy()
    # the workspace must be back to its all-zero invariant (shared-item G_Q / H_Q slots re-zeroed by phase C)
    rows_bytes = (B * P.shape[1] * 4 + 255) // 256 * 256
    assert int(ws.buf[256:256 + 2 * rows_bytes].count_nonzero().item()) == 0
    return tP.cpu().numpy(), tQ.cpu().numpy(), aP.cpu().numpy(), aQ.cpu().numpy(), stats.cpu().numpy(), counts


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("adver", [0, 1])
@pytest.mark.parametrize("d,U,I,S,B,zipf", [
    (64, 300, 200, 6, 512, False),     # reference batch size, heavy duplication
    (128, 5000, 3000, 3, 2048, False),
    (8, 50, 40, 4, 96, False),         # lane groups of 4 with predication
    (20, 64, 64, 3, 100, True),        # d not a power of two, Zipf users and items
    (256, 700, 900, 2, 777, True),     # two float4 per lane
    (384, 100, 100, 2, 130, False),
    (512, 64, 80, 2, 64, False),
    (64, 100000, 50000, 3, 1024, False),    # sparse batch: almost every segment takes the in-register fast path
    (128, 200000, 100000, 2, 4096, False),
])
def test_train_steps_match_oracle(cuda_device, mode, adver, d, U, I, S, B, zipf):
    rng = np.random.RandomState(d + B)
    P, Q, u, i, j = _problem(rng, U, I, d, S, B, zipf=zipf)
    lr, reg, reg_adv, eps = 0.05, 0.01, 1.0, 0.5
    rP, rQ, raP, raQ, rstats = _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, adver)
    gP, gQ, gaP, gaQ, gstats, counts = _run_cuda_steps(cuda_device, P, Q, u, i, j, lr, reg, reg_adv, eps, adver, mode)
    _close(gP, rP)
    _close(gQ, rQ)
    _close(gaP, raP)
    _close(gaQ, raQ)
    _close(gstats[:, 0], rstats[:, 0])
    assert np.array_equal(gstats[:, 1].astype(np.int64), rstats[:, 1].astype(np.int64))
    for s in range(S):  # exact unique-row counts feed the roofline byte model
        assert counts[s, 0] == np.unique(u[s]).size
        assert counts[s, 1] == np.unique(np.concatenate([i[s], j[s]])).size
    # one step, measured on the UPDATE: embeddings' and accumulators' changes within 1e-5 of the update's own scale
    r1 = _run_oracle_steps(P, Q, u[:1], i[:1], j[:1], lr, reg, reg_adv, eps, adver)
    g1 = _run_cuda_steps(cuda_device, P, Q, u[:1], i[:1], j[:1], lr, reg, reg_adv, eps, adver, mode)
    _close_update(g1[0], r1[0], P)
    _close_update(g1[1], r1[1], Q)
    _close_update(g1[2], r1[2], np.full_like(P, 0.1))
    _close_update(g1[3], r1[3], np.full_like(Q, 0.1))


@pytest.mark.parametrize("d,U,I,S,B,zipf", [(64, 300, 200, 4, 512, False), (128, 5000, 3000, 2, 2048, False),
                                            (24, 64, 64, 3, 100, True), (256, 700, 900, 2, 300, True),
                                            (64, 100000, 50000, 2, 1024, False)])
def test_adv_random_matches_oracle(cuda_device, d, U, I, S, B, zipf):
    """`--adv random` (A5', APR.py:170-177): Delta = eps * l2_normalize(truncated_normal) drawn afresh per step for every
    row, generated where a row is touched (Philox keyed by (seed, global step, table, row, column)) -- against the oracle
    that materialises the two noise tables.  Duplicated users / items, pair work units and fast segments all occur."""
    from apr_b200 import engine
    rng = np.random.RandomState(d + B + 1)
    P, Q, u, i, j = _problem(rng, U, I, d, S, B, zipf=zipf)
    lr, reg, reg_adv, eps, seed, step0 = 0.05, 0.01, 1.0, 0.5, 2019, 7
    rP, rQ = P.copy(), Q.copy()
    raP, raQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
    rstats = []
    for s in range(S):
        rstats.append(O.loss_acc(rP, rQ, u[s], i[s], j[s]))
        O.apr_step_random(rP, rQ, raP, raQ, u[s], i[s], j[s], lr, reg, reg_adv, eps, seed, step0 + s)
        if s == 0:
            first = [a.copy() for a in (rP, rQ, raP, raQ)]
    dev = cuda_device
    for n_steps in (S, 1):
        tP, tQ = _dev(P, torch.float32, dev), _dev(Q, torch.float32, dev)
        aP, aQ = torch.full_like(tP, 0.1), torch.full_like(tQ, 0.1)
        ws = engine.TrainWorkspace(n_steps, B, d, dev)
        stats = torch.zeros((n_steps, 2), dtype=torch.float32, device=dev)
        engine.train_steps_random(tP, tQ, aP, aQ, _dev(u[:n_steps], torch.int32, dev), _dev(i[:n_steps], torch.int32, dev),
                                  _dev(j[:n_steps], torch.int32, dev), lr, reg, reg_adv, eps, ws, seed, step0, stats=stats)
        got = [t.cpu().numpy() for t in (tP, tQ, aP, aQ)]
        if n_steps == S:
            for g, r in zip(got, (rP, rQ, raP, raQ)):
                _close(g, r)
            st = stats.cpu().numpy()
            _close(st[:, 0], np.asarray(rstats)[:, 0])
            assert np.array_equal(st[:, 1].astype(np.int64), np.asarray(rstats)[:, 1].astype(np.int64))
        else:
            for g, r, b in zip(got, first, (P, Q, np.full_like(P, 0.1), np.full_like(Q, 0.1))):
                _close_update(g, r, b)


def test_dns_branch_on_adversarial_graph(cuda_device):
    """utils.py:121-139 with args.adver = 1 and dns > 1: no update_P/update_Q, but the optimizer is the adversarial
    graph's (opt_loss with Delta == 0): data term x (1 + reg_adv), regulariser counted twice (engine adver mode 3)."""
    rng = np.random.RandomState(33)
    P, Q, u, i, j = _problem(rng, 300, 200, 64, 3, 512)
    lr, reg, reg_adv, eps = 0.05, 0.01, 0.7, 0.5
    for mode in (0, 1, 2):
        r = _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, 3)
        g = _run_cuda_steps(cuda_device, P, Q, u, i, j, lr, reg, reg_adv, eps, 3, mode)
        for a, b in zip(g[:4], r[:4]):
            _close(a, b)
    r1 = _run_oracle_steps(P, Q, u[:1], i[:1], j[:1], lr, reg, reg_adv, eps, 3)
    g1 = _run_cuda_steps(cuda_device, P, Q, u[:1], i[:1], j[:1], lr, reg, reg_adv, eps, 3, 0)
    _close_update(g1[0], r1[0], P)
    _close_update(g1[1], r1[1], Q)
    # and it differs from the plain BPR step (what round 1 ran in this branch)
    b1 = _run_oracle_steps(P, Q, u[:1], i[:1], j[:1], lr, reg, reg_adv, eps, 0)
    assert np.abs(b1[0] - r1[0]).max() > 1e-4


def test_out_of_range_ids_raise(cuda_device):
    """An id outside its table must not train silently (the reference's embedding_lookup raises InvalidArgument): the
    Session surfaces the index preparation's flag at the end of training_batch."""
    import types

    from apr_b200.APR import MF, Session
    args = types.SimpleNamespace(embed_size=16, lr=0.05, reg=0.0, dns=1, adv="grad", eps=0.5, adver=1, reg_adv=1.0, epochs=1)
    model = MF(50, 40, args)
    model.build_graph()
    rng = np.random.RandomState(1)
    u = rng.randint(0, 50, (2, 64)).astype(np.int32)
    i = rng.randint(0, 40, (2, 64)).astype(np.int32)
    j = rng.randint(0, 40, (2, 64)).astype(np.int32)
    sess = Session()
    sess.train_steps(model, *[_dev(x, torch.int32, cuda_device) for x in (u, i, j)], adver=True, check=True)   # fine
    i[1, 5] = 41 + model.extra_row      # one past the last row
    with pytest.raises(IndexError):
        sess.train_steps(model, *[_dev(x, torch.int32, cuda_device) for x in (u, i, j)], adver=True, check=True)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("adver", [0, 1])
@pytest.mark.parametrize("d", [64, 128, 24])
def test_pair_work_units_match_oracle(cuda_device, mode, adver, d):
    """Hand-built batch that exercises the PAIR path (csrc/train.cu: prep_pair_kernel / pair_unit) in every role --
    the twice-occurring item positive in both triples, negative in both, positive in one and negative in the other --
    next to everything that must NOT become a pair: an item occurring three times, a segment with two shared items, a
    user with two triples, i == j inside one triple, and plain singleton (fast-path) triples.  Two steps, so the second
    one reads what the first one wrote."""
    U, I = 64, 200
    rng = np.random.RandomState(d)
    P = (rng.randn(U, d) * 0.1).astype(np.float32)
    Q = (rng.randn(I, d) * 0.1).astype(np.float32)
    step = [
        (0, 10, 11), (1, 10, 12),        # item 10: pos / pos            -> pair
        (2, 13, 14), (3, 15, 14),        # item 14: neg / neg            -> pair
        (4, 16, 17), (5, 18, 16),        # item 16: pos / neg            -> pair
        (6, 20, 21), (7, 20, 22), (8, 23, 20),   # item 20 three times   -> general
        (9, 30, 31), (10, 30, 32), (11, 33, 31),  # segment 9 has two shared items -> general (30 and 31 occur twice)
        (12, 40, 41), (12, 42, 40),      # user 12 twice, item 40 twice  -> general
        (13, 50, 50),                    # i == j                        -> general (one segment registers twice)
        (14, 60, 61), (15, 62, 63), (16, 64, 65),   # singletons          -> fast path
        (17, 70, 71), (18, 72, 70),      # item 70: pos / neg            -> pair
    ]
    B = len(step)
    u = np.asarray([[t[0] for t in step], [t[0] for t in reversed(step)]], np.int32)
    i = np.asarray([[t[1] for t in step], [t[1] for t in reversed(step)]], np.int32)
    j = np.asarray([[t[2] for t in step], [t[2] for t in reversed(step)]], np.int32)
    assert u.shape == (2, B)
    lr, reg, reg_adv, eps = 0.05, 0.01, 1.0, 0.5
    rP, rQ, raP, raQ, rstats = _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, adver)
    gP, gQ, gaP, gaQ, gstats, counts = _run_cuda_steps(cuda_device, P, Q, u, i, j, lr, reg, reg_adv, eps, adver, mode)
    # items 10, 14, 16 and 70 make pairs; 20, 30/31, 40, 50 must not
    want_pairs = 0 if os.environ.get("APR_PAIRS", "1") == "0" else 4
    assert _run_cuda_steps.last_npair.tolist() == [want_pairs, want_pairs]
    _close(gP, rP)
    _close(gQ, rQ)
    _close(gaP, raP)
    _close(gaQ, raQ)
    _close(gstats[:, 0], rstats[:, 0])
    assert np.array_equal(gstats[:, 1].astype(np.int64), rstats[:, 1].astype(np.int64))


def test_train_step_single_triple_and_self_pair(cuda_device):
    # B = 1, and a triple whose positive and negative item coincide (x = 0, gradients cancel on the item row)
    rng = np.random.RandomState(5)
    P, Q, u, i, j = _problem(rng, 4, 4, 16, 2, 1)
    j[1] = i[1]
    for adver in (0, 1):
        rP, rQ, raP, raQ, _ = _run_oracle_steps(P, Q, u, i, j, 0.05, 0.0, 1.0, 0.5, adver)
        gP, gQ, gaP, gaQ, _, _ = _run_cuda_steps(cuda_device, P, Q, u, i, j, 0.05, 0.0, 1.0, 0.5, adver, 0)
        _close(gP, rP)
        _close(gQ, rQ)
        _close(gaQ, raQ)


def test_train_rejects_bad_arguments(cuda_device):
    from apr_b200 import _lib, engine
    dev = cuda_device
    P = torch.zeros((10, 8), device=dev)
    Q = torch.zeros((10, 8), device=dev)
    u = torch.zeros((1, 4), dtype=torch.int32, device=dev)
    ws = engine.TrainWorkspace(1, 4, 8, dev)
    with pytest.raises(ValueError):
        engine.train_steps(P, Q, P.clone(), Q.clone(), u, u, u[:, :2], 0.05, 0, 1, 0.5, 1, ws)
    with pytest.raises(ValueError):  # d not a multiple of 4
        engine.TrainWorkspace(1, 4, 6, dev)
    bad = torch.full((1, 4), 10, dtype=torch.int32, device=dev)  # id == rows -> flagged, not a fault
    engine.train_steps(P, Q, torch.full_like(P, 0.1), torch.full_like(Q, 0.1), bad, u, u, 0.05, 0, 1, 0.5, 1, ws)
    with pytest.raises(_lib.AprError):
        ws.unique_counts(1)


def test_loss_acc_matches_oracle(cuda_device):
    from apr_b200 import engine
    rng = np.random.RandomState(11)
    for d in (64, 128, 24):
        P, Q, u, i, j = _problem(rng, 400, 300, d, 5, 700, scale=0.8)
        out = engine.loss_acc(_dev(P, torch.float32, cuda_device), _dev(Q, torch.float32, cuda_device),
                              _dev(u, torch.int32, cuda_device), _dev(i, torch.int32, cuda_device),
                              _dev(j, torch.int32, cuda_device)).cpu().numpy()
        for s in range(5):
            l, c = O.loss_acc(P, Q, u[s], i[s], j[s])
            assert abs(out[s, 0] - l) <= RTOL * abs(l)
            # x > 0 can flip only where |x| is at rounding level
            x, _, _ = O.forward(P, Q, u[s], i[s], j[s])
            assert abs(int(out[s, 1]) - c) <= int((np.abs(x) < 1e-6).sum())


def test_sampler_bit_exact(cuda_device):
    from apr_b200 import engine
    rng = np.random.RandomState(2)
    U, I = 700, 450
    lists = [sorted(set(rng.randint(0, I, rng.randint(1, 60)).tolist())) for _ in range(U)]
    lists[3] = list(range(I - 1))  # a user with a single admissible negative
    pu = np.concatenate([[u] * len(l) for u, l in enumerate(lists)]).astype(np.int32)
    pi = np.concatenate(lists).astype(np.int32)
    ptr, idx = O.build_csr(lists[:-5])  # trainList shorter than num_users: the tail users reject nothing
    for B, dns, epoch in ((512, 1, 0), (100, 3, 7), (1, 1, 2)):
        want = O.sample_epoch(pu, pi, B, I, ptr, idx, 2019, epoch, dns)
        got = engine.sample_epoch(_dev(pu, torch.int32, cuda_device), _dev(pi, torch.int32, cuda_device), B, I,
                                  _dev(ptr, torch.int64, cuda_device), _dev(idx, torch.int32, cuda_device), 2019, epoch, dns)
        assert int(got[4].item()) == 0
        for g, w in zip(got[:4], want):
            assert np.array_equal(g.cpu().numpy(), w)
    # fork-emulating mode: the W chunks of a pool.map round share one counter range
    want = O.sample_epoch(pu, pi, 64, I, ptr, idx, 2019, 1, 2, fork_workers=5)
    got = engine.sample_epoch(_dev(pu, torch.int32, cuda_device), _dev(pi, torch.int32, cuda_device), 64, I,
                              _dev(ptr, torch.int64, cuda_device), _dev(idx, torch.int32, cuda_device), 2019, 1, 2, fork_workers=5)
    for g, w in zip(got[:4], want):
        assert np.array_equal(g.cpu().numpy(), w)


def test_sampler_rank_shards_are_slices_of_the_epoch(cuda_device):
    """SURVEY 8(e) sampler row: rank r of G draws columns [r*B/G, (r+1)*B/G) of every batch; because every random number
    is keyed by the triple's index in the whole epoch the shards are bit-identical slices of the oracle's epoch."""
    from apr_b200 import engine
    rng = np.random.RandomState(12)
    U, I = 500, 300
    lists = [sorted(set(rng.randint(0, I, rng.randint(1, 40)).tolist())) for _ in range(U)]
    pu = np.concatenate([[u] * len(l) for u, l in enumerate(lists)]).astype(np.int32)
    pi = np.concatenate(lists).astype(np.int32)
    ptr, idx = O.build_csr(lists)
    a = [_dev(pu, torch.int32, cuda_device), _dev(pi, torch.int32, cuda_device)]
    c = [_dev(ptr, torch.int64, cuda_device), _dev(idx, torch.int32, cuda_device)]
    for B, dns, G in ((512, 1, 4), (96, 3, 8), (64, 1, 2)):
        want = O.sample_epoch(pu, pi, B, I, ptr, idx, 2019, 5, dns)
        bl = B // G
        for r in range(G):
            got = engine.sample_epoch(*a, B, I, *c, 2019, 5, dns, rank=r, world=G)
            assert int(got[4].item()) == 0
            assert np.array_equal(got[0].cpu().numpy(), want[0][:, r * bl:(r + 1) * bl])
            assert np.array_equal(got[1].cpu().numpy(), want[1][:, r * bl:(r + 1) * bl])
            assert np.array_equal(got[2].cpu().numpy(), want[2][:, r * bl * dns:(r + 1) * bl * dns])
            assert np.array_equal(got[3].cpu().numpy(), want[3][:, r * bl * dns:(r + 1) * bl * dns])


def test_select_dns_bit_exact(cuda_device):
    from apr_b200 import engine
    rng = np.random.RandomState(4)
    P, Q, _, _, _ = _problem(rng, 100, 90, 64, 1, 1, scale=1.0)
    ud = np.repeat(rng.randint(0, 100, 500), 4).astype(np.int32)
    jd = rng.randint(0, 90, 2000).astype(np.int32)
    jd[4:8] = jd[4]  # all-equal scores: first wins
    got = engine.select_dns(_dev(P, torch.float32, cuda_device), _dev(Q, torch.float32, cuda_device),
                            _dev(ud, torch.int32, cuda_device), _dev(jd, torch.int32, cuda_device), 4).cpu().numpy()
    assert np.array_equal(got, O.select_dns(P, Q, ud, jd, 4))


def test_truncated_normal_matches_oracle(cuda_device):
    from apr_b200 import engine
    W = torch.empty((1000, 64), dtype=torch.float32, device=cuda_device)
    engine.init_truncated_normal(W, 0.01, 2019, 1)
    ref = O.truncated_normal(1000, 64, 0.01, 2019, 1)
    got = W.cpu().numpy()
    # transcendental functions differ in the last ulp between libm and CUDA: tolerance, not bit-exactness
    assert np.abs(got - ref).max() <= 1e-5 * 0.02
    assert np.abs(got).max() <= 0.02


@pytest.mark.parametrize("d", [8, 64, 100, 128, 256])
def test_scores_bit_exact(cuda_device, d):
    from apr_b200 import engine
    rng = np.random.RandomState(d)
    P, Q, _, _, _ = _problem(rng, 200, 300, d, 1, 1, scale=1.0)
    us = rng.randint(0, 200, 5000).astype(np.int32)
    it = rng.randint(0, 300, 5000).astype(np.int32)
    got = engine.score_pairs(_dev(P, torch.float32, cuda_device), _dev(Q, torch.float32, cuda_device),
                             _dev(us, torch.int32, cuda_device), _dev(it, torch.int32, cuda_device)).cpu().numpy()
    want = O.score_pairs(P, Q, us, it)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_eval_candidates_bit_exact(cuda_device):
    from apr_b200 import engine
    rng = np.random.RandomState(9)
    P, Q, _, _, _ = _problem(rng, 120, 500, 64, 1, 1, scale=1.0)
    Q[17] = Q[400]  # exact ties, including with held-out items
    users = np.arange(120, dtype=np.int32)
    rows, ptr = [], [0]
    for u in users:
        n = [0, 1, 100, 101, 257][u % 5]
        c = rng.randint(0, 500, n).tolist() + [17 if u % 3 == 0 else int(rng.randint(0, 500))]
        if n:
            c[0] = 400
        rows.append(c)
        ptr.append(ptr[-1] + len(c))
    cand = np.concatenate(rows).astype(np.int32)
    pos, sc = engine.eval_candidates(_dev(P, torch.float32, cuda_device), _dev(Q, torch.float32, cuda_device),
                                     _dev(users, torch.int32, cuda_device), _dev(np.asarray(ptr), torch.int64, cuda_device),
                                     _dev(cand, torch.int32, cuda_device), want_scores=True)
    pos = pos.cpu().numpy()
    for u in users:
        assert pos[u] == O.eval_candidates_position(P, Q, int(u), rows[u])
    want = O.score_pairs(P, Q, np.repeat(users, np.diff(ptr)), cand)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), want.view(np.uint32))


def _fullrank_case(rng, U, I, d, rows_extra=1, scale=1.0, ties=True):
    P = (rng.randn(U + rows_extra, d) * scale).astype(np.float32)
    Q = (rng.randn(I + rows_extra, d) * scale).astype(np.float32)
    if ties:
        Q[I // 2] = Q[1]
        Q[I // 3] = Q[1]
    train = [sorted(set(rng.randint(0, I, rng.randint(0, 30)).tolist())) for _ in range(U)]
    test = rng.randint(0, I + rows_extra, U).astype(np.int32)  # may be == num_items (Video: test id 23714)
    test[0] = 1
    test[1] = I // 2
    if train[2]:
        test[2] = train[2][0]  # held-out item that is also a train item
    return P, Q, train, test


@pytest.mark.parametrize("U,I,d,k_top", [(150, 333, 64, 0), (150, 333, 64, 10), (70, 1000, 128, 100), (33, 2100, 32, 128),
                                         (64, 64, 256, 5), (200, 130, 20, 100)])
def test_eval_fullrank_bit_exact(cuda_device, U, I, d, k_top):
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    rng = np.random.RandomState(U + I)
    P, Q, train, test = _fullrank_case(rng, U, I, d)
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    users = np.arange(U, dtype=np.int32)
    pos, ids, sc = engine.eval_fullrank(_dev(P, torch.float32, cuda_device), _dev(Q, torch.float32, cuda_device),
                                        _dev(users, torch.int32, cuda_device), _dev(test, torch.int32, cuda_device), 0, I,
                                        _dev(ptr, torch.int64, cuda_device), _dev(idx, torch.int32, cuda_device), k_top,
                                        exact=True)
    pos = pos.cpu().numpy()
    ids = ids.cpu().numpy() if k_top else None
    for u in range(U):
        p, nneg, tids, tsc = O.eval_fullrank_user(P, Q, u, int(test[u]), train[u], I, max(k_top, 1))
        assert pos[u] == p, (u, pos[u], p)
        if k_top:
            # library returns the top-k of the NEGATIVES; the held-out item enters at rank `position` (loses ties)
            neg_ids = [t for t in ids[u].tolist() if t >= 0]
            merged = neg_ids[:p] + [int(test[u])] + neg_ids[p:] if p < k_top else neg_ids
            assert merged[:k_top] == tids.tolist()[:k_top], u


def test_eval_fullrank_item_sharded_sums(cuda_device):
    """Item-sharded evaluation: per-shard counts add up to the single-range count."""
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    rng = np.random.RandomState(77)
    U, I, d = 90, 1500, 64
    P, Q, train, test = _fullrank_case(rng, U, I, d)
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    a = [_dev(P, torch.float32, cuda_device), _dev(Q, torch.float32, cuda_device),
         _dev(np.arange(U, dtype=np.int32), torch.int32, cuda_device), _dev(test, torch.int32, cuda_device)]
    e = [_dev(ptr, torch.int64, cuda_device), _dev(idx, torch.int32, cuda_device)]
    whole, _, _ = engine.eval_fullrank(*a, 0, I, *e, 0, exact=True)
    acc = torch.zeros(U, dtype=torch.int32, device=cuda_device)
    for lo, hi in ((0, 401), (401, 1000), (1000, I)):
        engine.eval_fullrank(*a, lo, hi, *e, 0, exact=True, position=acc)
    assert torch.equal(whole, acc)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("adver", [0, 1])
def test_row_sharded_step_matches_oracle_emulated_ranks(cuda_device, world, adver):
    """Row-sharded tables + every-G-th-segment work split (the multi-GPU training path) checked on ONE GPU: the G ranks'
    stage launches run back to back (no launch waits on another), shards live in this device's memory."""
    from apr_b200 import engine
    from apr_b200.distributed import ShardedTables, train_steps_sharded
    rng = np.random.RandomState(world * 10 + adver)
    U, I, d, S, B = 3001, 1501, 64, 3, 1536
    P, Q, u, i, j = _problem(rng, U, I, d, S, B)
    lr, reg, reg_adv, eps = 0.05, 0.01, 1.0, 0.5
    rP, rQ, raP, raQ, _ = _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, adver)
    dev = cuda_device
    t = ShardedTables(U, I, d, B, dev, world=world, rank=0, symmetric=False)
    t.load_full("P", _dev(P, torch.float32, dev))
    t.load_full("Q", _dev(Q, torch.float32, dev))
    t.load_full("accP", torch.full((U, d), 0.1, device=dev))
    t.load_full("accQ", torch.full((I, d), 0.1, device=dev))
    ws = engine.TrainWorkspace(S, B, d, dev)
    train_steps_sharded(t, _dev(u, torch.int32, dev), _dev(i, torch.int32, dev), _dev(j, torch.int32, dev), lr, reg, reg_adv,
                        eps, adver, ws, ranks=list(range(world)))
    torch.cuda.synchronize()
    _close(t.gather_full("P", U).cpu().numpy(), rP)
    _close(t.gather_full("Q", I).cpu().numpy(), rQ)
    _close(t.gather_full("accP", U).cpu().numpy(), raP)
    _close(t.gather_full("accQ", I).cpu().numpy(), raQ)
    for r in range(world):  # shared-item workspace shards back to zero
        assert int(t.local("GQ", r).count_nonzero().item()) == 0 and int(t.local("HQ", r).count_nonzero().item()) == 0


def test_graph_cache_never_reinstantiates_after_first_sight(cuda_device):
    """Round 1 re-instantiated its step graph whenever a call's step count differed from the previous one (a 20-step timed
    call after a 5-step warm-up paid two cudaGraphInstantiate inside the timed region).  Now the first call of a
    configuration builds the executable graphs of EVERY group size; any later step count replays them: the library's
    instantiation counter must not move, and the results stay the oracle's."""
    from apr_b200 import engine
    rng = np.random.RandomState(77)
    U, I, d, B = 5000, 3000, 64, 2048
    S_all = 3 + 1 + 20 + 7 + 33
    P, Q, u, i, j = _problem(rng, U, I, d, S_all, B)
    lr, reg, reg_adv, eps = 0.05, 0.01, 1.0, 0.5
    dev = cuda_device
    tP, tQ = _dev(P, torch.float32, dev), _dev(Q, torch.float32, dev)
    aP, aQ = torch.full_like(tP, 0.1), torch.full_like(tQ, 0.1)
    ws = engine.TrainWorkspace(40, B, d, dev)
    tu, ti, tj = _dev(u, torch.int32, dev), _dev(i, torch.int32, dev), _dev(j, torch.int32, dev)
    s0, counts = 0, []
    for n in (3, 1, 20, 7, 33):                      # odd step counts: every group size 32/16/8/4/2/1 gets replayed
        engine.train_steps(tP, tQ, aP, aQ, tu[s0:s0 + n], ti[s0:s0 + n], tj[s0:s0 + n], lr, reg, reg_adv, eps, 1, ws, mode=0)
        torch.cuda.synchronize()
        counts.append(engine.context_stats())
        s0 += n
    assert counts[0]["graph_launches"] > 0, "graph replay is off for this batch size"
    for c in counts[1:]:
        assert c["graph_instantiations"] == counts[0]["graph_instantiations"], counts
        assert c["graph_updates"] == counts[0]["graph_updates"], counts
    assert counts[-1]["graph_launches"] > counts[0]["graph_launches"]
    # BPR on the same tables / workspace is a different configuration: it builds its own graphs once, then replays
    engine.train_steps(tP, tQ, aP, aQ, tu[:5], ti[:5], tj[:5], lr, reg, reg_adv, eps, 0, ws, mode=0)
    c_bpr = engine.context_stats()
    engine.train_steps(tP, tQ, aP, aQ, tu[:9], ti[:9], tj[:9], lr, reg, reg_adv, eps, 0, ws, mode=0)
    engine.train_steps(tP, tQ, aP, aQ, tu[:11], ti[:11], tj[:11], lr, reg, reg_adv, eps, 1, ws, mode=0)
    assert engine.context_stats()["graph_instantiations"] == c_bpr["graph_instantiations"]
    # and the arithmetic of the replayed steps is the oracle's (the APR part, before the extra calls: re-run it)
    tP2, tQ2 = _dev(P, torch.float32, dev), _dev(Q, torch.float32, dev)
    aP2, aQ2 = torch.full_like(tP2, 0.1), torch.full_like(tQ2, 0.1)
    s0 = 0
    for n in (3, 1, 20, 7, 33):
        engine.train_steps(tP2, tQ2, aP2, aQ2, tu[s0:s0 + n], ti[s0:s0 + n], tj[s0:s0 + n], lr, reg, reg_adv, eps, 1, ws, mode=0)
        s0 += n
    rP, rQ, raP, raQ, _ = _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, 1)
    _close(tP2.cpu().numpy(), rP, 1e-4)              # 64 sequential APR steps (drift bound as in tests/test_gpu_e2e.py;
    _close(tQ2.cpu().numpy(), rQ, 1e-4)              # a wrong step index would be off by orders of magnitude more)


@pytest.mark.parametrize("adver", [0, 1])
def test_sharded_trainer_pipeline_single_rank(cuda_device, adver):
    """ShardedTrainer's pipeline (side-stream preparation per sub-chunk, ready events, two alternating workspaces, calls
    shorter than steps_per_call) on one rank (G = 1: no collective): five calls against the oracle."""
    from apr_b200.distributed import ShardedTables, ShardedTrainer
    rng = np.random.RandomState(40 + adver)
    U, I, d, S, B = 2000, 1200, 64, 11, 1024
    P, Q, u, i, j = _problem(rng, U, I, d, S, B)
    lr, reg, reg_adv, eps = 0.05, 0.01, 1.0, 0.5
    rP, rQ, raP, raQ, rstats = _run_oracle_steps(P, Q, u, i, j, lr, reg, reg_adv, eps, adver)
    dev = cuda_device
    t = ShardedTables(U, I, d, B, dev, world=1, rank=0, symmetric=False)
    t.load_full("P", _dev(P, torch.float32, dev))
    t.load_full("Q", _dev(Q, torch.float32, dev))
    t.load_full("accP", torch.full((U, d), 0.1, device=dev))
    t.load_full("accQ", torch.full((I, d), 0.1, device=dev))
    tr = ShardedTrainer(t, 3, B)
    for s0 in (0, 3, 6, 9):                                # 3 + 3 + 3 + 2 steps
        s1 = min(S, s0 + 3)
        host = s0 == 3                                      # one call fed from pinned host memory
        loc = [torch.from_numpy(x[s0:s1]).pin_memory() if host else _dev(x[s0:s1], torch.int32, dev) for x in (u, i, j)]
        tr.train_steps(*loc, lr, reg, reg_adv, eps, adver)
    tr.synchronize()
    tr.check()
    _close(t.gather_full("P", U).cpu().numpy(), rP)
    _close(t.gather_full("Q", I).cpu().numpy(), rQ)
    _close(t.gather_full("accP", U).cpu().numpy(), raP)
    _close(t.gather_full("accQ", I).cpu().numpy(), raQ)


@pytest.mark.timeout(900)
def test_full_size_config4_step_matches_oracle_on_touched_rows(cuda_device):
    """BASELINE.json configs[3] at FULL size (10M users x 2M items, d=128, 65 536 triples per step): two APR steps on the
    GPU; every row they touch is gathered and compared with the oracle run on those rows; untouched rows must be
    bit-identical to what they were (checksum of a strided sample)."""
    from apr_b200 import engine
    dev = cuda_device
    if torch.cuda.get_device_properties(dev).total_memory < 40e9:
        pytest.skip("needs ~13 GB of device memory")
    U, I, d, S, B = 10_000_000, 2_000_000, 128, 2, 65536
    P = torch.empty((U, d), device=dev)
    Q = torch.empty((I, d), device=dev)
    engine.init_truncated_normal(P, 0.01, 2019, 0)
    engine.init_truncated_normal(Q, 0.01, 2019, 1)
    aP, aQ = torch.full_like(P, 0.1), torch.full_like(Q, 0.1)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    u = torch.randint(0, U, (S, B), device=dev, dtype=torch.int32, generator=g)
    i = torch.randint(0, I, (S, B), device=dev, dtype=torch.int32, generator=g)
    j = torch.randint(0, I, (S, B), device=dev, dtype=torch.int32, generator=g)
    uu = torch.unique(u.flatten().long())
    ii = torch.unique(torch.cat([i.flatten(), j.flatten()]).long())
    P0, Q0 = P[uu].cpu().numpy(), Q[ii].cpu().numpy()
    sample_p = torch.arange(0, U, 997, device=dev)
    mask = ~torch.isin(sample_p, uu)
    before = P[sample_p[mask]].clone()
    ws = engine.TrainWorkspace(S, B, d, dev)
    engine.train_steps(P, Q, aP, aQ, u, i, j, 0.05, 0.0, 1.0, 0.5, 1, ws)
    torch.cuda.synchronize()
    # oracle on the compact tables of touched rows
    un, inn, jn = u.cpu().numpy(), i.cpu().numpy(), j.cpu().numpy()
    uun, iin = uu.cpu().numpy(), ii.cpu().numpy()
    raP, raQ = np.full_like(P0, 0.1), np.full_like(Q0, 0.1)
    for s in range(S):
        O.apr_step(P0, Q0, raP, raQ, np.searchsorted(uun, un[s]), np.searchsorted(iin, inn[s]), np.searchsorted(iin, jn[s]),
                   0.05, 0.0, 1.0, 0.5, 1)
    _close(P[uu].cpu().numpy(), P0)
    _close(Q[ii].cpu().numpy(), Q0)
    _close(aP[uu].cpu().numpy(), raP)
    _close(aQ[ii].cpu().numpy(), raQ)
    assert torch.equal(P[sample_p[mask]], before)                  # untouched rows untouched
    cnt = ws.unique_counts(S)
    assert cnt[:, 0].sum() == sum(np.unique(un[s]).size for s in range(S))
    assert cnt[:, 1].sum() == sum(np.unique(np.concatenate([inn[s], jn[s]])).size for s in range(S))
