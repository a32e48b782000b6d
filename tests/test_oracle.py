"""CPU tests that pin the oracle (oracle/apr_oracle.py) -- see its header for what each one anchors."""
import math

import numpy as np
import pytest
import torch

from oracle import apr_oracle as O


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = [int(x) for x in O.philox4x32_10(*[np.uint32(v) for v in c], k[0], k[1])]
        assert got == want


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000, 4097, 65536, 100003])
def test_feistel_is_a_bijection(n):
    p = O.feistel_permutation(np.arange(n), n, 2019, 5)
    assert np.array_equal(np.sort(p), np.arange(n))
    if n > 1000:
        q = O.feistel_permutation(np.arange(n), n, 2019, 6)
        assert (p != q).mean() > 0.9  # different epochs -> different permutations


def test_fma_emulation_is_exact():
    rng = np.random.RandomState(1)
    a = rng.randn(50000).astype(np.float32)
    b = rng.randn(50000).astype(np.float32)
    c = (rng.randn(50000) * rng.choice([1e-7, 1.0, 1e4], 50000)).astype(np.float32)
    got = O._fma32(a, b, c)
    want = np.array([np.float32(math.fma(float(x), float(y), float(z))) if hasattr(math, "fma") else 0
                     for x, y, z in zip(a[:0], b[:0], c[:0])], dtype=np.float32)
    # exact reference through Python integers (fractions): a*b+c rounded once to binary32
    from fractions import Fraction
    idx = rng.choice(50000, 3000, replace=False)
    for k in idx:
        exact = Fraction(float(a[k])) * Fraction(float(b[k])) + Fraction(float(c[k]))
        lo = np.nextafter(got[k], np.float32(-np.inf))
        hi = np.nextafter(got[k], np.float32(np.inf))
        err = abs(Fraction(float(got[k])) - exact)
        assert err <= abs(Fraction(float(lo)) - exact) and err <= abs(Fraction(float(hi)) - exact)
    assert want.size == 0


def _torch_loss(P, Q, dP, dQ, u, i, j, reg, reg_adv, adver):
    """opt_loss exactly as written in APR.py:143-165 (float64 torch)."""
    p, q, n = P[u], Q[i], Q[j]
    out = (p * q).sum(1, keepdim=True)
    out_neg = (p * n).sum(1, keepdim=True)
    result = torch.clamp(out - out_neg, -80.0, 1e8)
    loss = torch.nn.functional.softplus(-result).sum()
    opt = loss + reg * torch.mean(p ** 2 + q ** 2 + n ** 2)
    if adver:
        pd, qd, nd = p + dP[u], q + dQ[i], n + dQ[j]
        result_adv = torch.clamp((pd * qd).sum(1, keepdim=True) - (pd * nd).sum(1, keepdim=True), -80.0, 1e8)
        loss_adv = torch.nn.functional.softplus(-result_adv).sum()
        opt = opt + reg_adv * loss_adv + reg * torch.mean(p ** 2 + q ** 2 + n ** 2)
    return loss, opt


@pytest.mark.parametrize("adver,reg", [(0, 0.0), (0, 0.3), (1, 0.0), (1, 0.3)])
def test_step_matches_torch_autograd_and_adagrad(adver, reg):
    rng = np.random.RandomState(3)
    U, I, d, B = 13, 9, 8, 40  # heavy duplication of users and items, items appear as pos and neg
    P = rng.randn(U, d) * 0.5
    Q = rng.randn(I, d) * 0.5
    u = rng.randint(0, U, B)
    i = rng.randint(0, I, B)
    j = rng.randint(0, I, B)
    lr, reg_adv, eps = 0.05, 0.7, 0.5
    accP = np.full_like(P, 0.1)
    accQ = np.full_like(Q, 0.1)
    P0, Q0 = P.copy(), Q.copy()
    info = O.apr_step(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, adver)

    tP = torch.tensor(P0, requires_grad=True)
    tQ = torch.tensor(Q0, requires_grad=True)
    tu, ti, tj = [torch.tensor(x, dtype=torch.long) for x in (u, i, j)]
    # update_P/update_Q: gradient of the PLAIN loss, row-normalised (APR.py:183-191)
    loss, _ = _torch_loss(tP, tQ, None, None, tu, ti, tj, reg, reg_adv, 0)
    gP, gQ = torch.autograd.grad(loss, [tP, tQ])
    dP = torch.nn.functional.normalize(gP, dim=1, eps=1e-6) * eps  # x / max(||x||, 1e-6) == x*rsqrt(max(ss,1e-12))
    dQ = torch.nn.functional.normalize(gQ, dim=1, eps=1e-6) * eps
    if adver:
        assert np.allclose(info["dP"], dP.numpy(), rtol=1e-9, atol=1e-12)
        assert np.allclose(info["dQ"], dQ.numpy(), rtol=1e-9, atol=1e-12)
    opt = torch.optim.Adagrad([tP, tQ], lr=lr, initial_accumulator_value=0.1, eps=0.0)
    _, opt_loss = _torch_loss(tP, tQ, dP.detach(), dQ.detach(), tu, ti, tj, reg, reg_adv, adver)
    opt.zero_grad()
    opt_loss.backward()
    assert np.allclose(info["gP"], tP.grad.numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(info["gQ"], tQ.grad.numpy(), rtol=1e-9, atol=1e-12)
    opt.step()
    assert np.allclose(P, tP.detach().numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(Q, tQ.detach().numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("variant", ["random", "delta_zero"])
def test_given_delta_steps_match_torch_autograd(variant):
    """`--adv random` (APR.py:170-177: Delta = eps * l2_normalize(noise), a constant of the step) and the dns > 1 branch
    on an adversarial graph (utils.py:121-139: Delta == 0) against torch autograd + Adagrad of opt_loss as written in
    APR.py:153-165, fp64."""
    rng = np.random.RandomState(5)
    U, I, d, B = 13, 9, 8, 40
    P = rng.randn(U, d) * 0.5
    Q = rng.randn(I, d) * 0.5
    u, i, j = rng.randint(0, U, B), rng.randint(0, I, B), rng.randint(0, I, B)
    lr, reg, reg_adv, eps, seed, step = 0.05, 0.01, 0.7, 0.5, 2019, 11
    accP, accQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
    P0, Q0 = P.copy(), Q.copy()
    if variant == "random":
        info = O.apr_step_random(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, seed, step)
        dP, dQ = info["dP"], info["dQ"]
        assert np.allclose(np.linalg.norm(dP, axis=1), eps, rtol=1e-5)       # every row: length eps
        assert np.array_equal(dP, O.random_delta(U, d, eps, seed, step, 0))  # counter-based: reproducible
        assert not np.array_equal(dP, O.random_delta(U, d, eps, seed, step + 1, 0))   # fresh noise at every step
        assert not np.array_equal(dP[:I], dQ)                                # P and Q tables differ
    else:
        info = O.apr_step(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, 3)
        dP, dQ = np.zeros_like(P0), np.zeros_like(Q0)
    tP = torch.tensor(P0, requires_grad=True)
    tQ = torch.tensor(Q0, requires_grad=True)
    tu, ti, tj = [torch.tensor(x, dtype=torch.long) for x in (u, i, j)]
    opt = torch.optim.Adagrad([tP, tQ], lr=lr, initial_accumulator_value=0.1, eps=0.0)
    _, opt_loss = _torch_loss(tP, tQ, torch.tensor(dP.astype(np.float64)), torch.tensor(dQ.astype(np.float64)), tu, ti, tj,
                              reg, reg_adv, 1)
    opt.zero_grad()
    opt_loss.backward()
    assert np.allclose(info["gP"], tP.grad.numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(info["gQ"], tQ.grad.numpy(), rtol=1e-9, atol=1e-12)
    opt.step()
    assert np.allclose(P, tP.detach().numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(Q, tQ.detach().numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("kind", ["mf", "bpr"])
def test_keras_step_gradients_match_torch_autograd(kind):
    """N3: oracle.keras_step (MF.py:21-24 BCE on the clipped raw dot product / BPR.py:11-21 `1 - log sigmoid`): the loss
    and, through the first Adam moments (m_1 = (1 - b1) g, v_1 = (1 - b2) g^2), the dense gradients equal torch autograd
    of the loss as written; the update follows Keras 2.2's Adam formula (epsilon outside the bias correction)."""
    rng = np.random.RandomState(0)
    U, I, d, n = 30, 20, 8, 64
    P = rng.uniform(-0.4, 0.4, (U, d))
    Q = rng.uniform(-0.4, 0.4, (I, d))
    u, i, j = rng.randint(0, U, n), rng.randint(0, I, n), rng.randint(0, I, n)
    y = rng.randint(0, 2, n).astype(np.float64)
    tP, tQ = torch.tensor(P, requires_grad=True), torch.tensor(Q, requires_grad=True)
    tu, ti, tj = [torch.tensor(x, dtype=torch.long) for x in (u, i, j)]
    if kind == "mf":
        ph = (tP[tu] * tQ[ti]).sum(1).clamp(1e-7, 1 - 1e-7)
        loss = -(torch.tensor(y) * ph.log() + (1 - torch.tensor(y)) * (1 - ph).log()).mean()
    else:
        x = (tP[tu] * tQ[ti]).sum(1) - (tP[tu] * tQ[tj]).sum(1)
        loss = (1 - torch.nn.functional.logsigmoid(x)).mean()
    gP, gQ = torch.autograd.grad(loss, [tP, tQ])
    mP, vP, mQ, vQ = [np.zeros_like(a) for a in (P, P, Q, Q)]
    P2, Q2 = P.copy(), Q.copy()
    total = O.keras_step(P2, Q2, mP, vP, mQ, vQ, u, i, j=None if kind == "mf" else j, y=y if kind == "mf" else None, t=1)
    assert abs(total / n - loss.item()) < 1e-12
    assert np.allclose(mP, 0.1 * gP.numpy(), rtol=1e-9, atol=1e-15) and np.allclose(mQ, 0.1 * gQ.numpy(), rtol=1e-9, atol=1e-15)
    assert np.allclose(vP, 0.001 * gP.numpy() ** 2, rtol=1e-9, atol=1e-18)
    lr_t = 0.001 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(P2, P - lr_t * mP / (np.sqrt(vP) + 1e-7), rtol=1e-12, atol=0)
    assert np.all(P2[np.setdiff1d(np.arange(U), u)] == P[np.setdiff1d(np.arange(U), u)])   # first step: untouched rows have g = 0


def test_clip_blocks_gradient_outside_interval():
    P = np.array([[40.0, 0.0]], dtype=np.float64)
    Q = np.array([[0.0, 0.0], [3.0, 0.0]], dtype=np.float64)
    # x = <p,q0> - <p,q1> = -120 < -80  -> gradient masked (tf.clip_by_value)
    GP, GQ, c, x = O.plain_row_gradients(P, Q, np.array([0]), np.array([0]), np.array([1]))
    assert x[0] == -120.0 and c[0] == 0.0 and not GP.any() and not GQ.any()


def test_metrics_from_position_matches_reference_formulas():
    pos = np.array([0, 3, 9, 10, 250])
    nneg = np.array([100, 100, 50, 1000, 300])
    res = O.metrics_from_position(pos, nneg, 10)
    for r, (p, n) in enumerate(zip(pos, nneg)):
        for k in range(1, 11):  # utils.py:257-261
            assert res[r, 0, k - 1] == (p < k)
            assert res[r, 1, k - 1] == (math.log(2) / math.log(p + 2) if p < k else 0)
            assert res[r, 2, k - 1] == 1 - (p / n)


def test_fullrank_user_semantics():
    rng = np.random.RandomState(0)
    P = rng.randn(3, 8).astype(np.float32)
    Q = rng.randn(12, 8).astype(np.float32)
    Q[5] = Q[7]  # a tie between two negatives and ...
    Q[9] = Q[7]  # ... with the held-out item
    pos, nneg, ids, sc = O.eval_fullrank_user(P, Q, 1, 9, [0, 3], 11, 5)
    s = O.score_pairs(P, Q, np.full(12, 1), np.arange(12))
    cands = [c for c in range(11) if c not in (0, 3, 9)]
    assert nneg == len(cands)
    assert pos == sum(s[c] >= s[9] for c in cands)  # held-out loses ties
    order = sorted(cands + [9], key=lambda c: (-float(s[c]), c if c != 9 else 10 ** 9))
    assert ids.tolist() == order[:5]


def test_sampler_properties():
    rng = np.random.RandomState(0)
    U, I = 50, 40
    lists = [sorted(set(rng.randint(0, I, rng.randint(1, 12)).tolist())) for _ in range(U)]
    pu = np.concatenate([[u] * len(l) for u, l in enumerate(lists)]).astype(np.int32)
    pi = np.concatenate(lists).astype(np.int32)
    ptr, idx = O.build_csr(lists)
    u, i, ud, j = O.sample_epoch(pu, pi, 16, I, ptr, idx, 2019, 0, dns=2)
    S = len(pu) // 16
    assert u.shape == (S, 16) and j.shape == (S, 32)
    # positives are a prefix of a permutation of the pairs; negatives are never train items
    seen = set(zip(u.ravel().tolist(), i.ravel().tolist()))
    assert len(seen) == S * 16 and seen <= set(zip(pu.tolist(), pi.tolist()))
    for uu, jj in zip(ud.ravel(), j.ravel()):
        assert 0 <= jj < I and jj not in lists[uu]
    u2, i2, _, j2 = O.sample_epoch(pu, pi, 16, I, ptr, idx, 2019, 0, dns=2)
    assert np.array_equal(u, u2) and np.array_equal(j, j2)
    u3, _, _, j3 = O.sample_epoch(pu, pi, 16, I, ptr, idx, 2019, 1, dns=2)
    assert not np.array_equal(u, u3)


def test_truncated_normal_statistics():
    w = O.truncated_normal(2000, 16, 0.01, 2019, 0)
    assert np.abs(w).max() <= 0.02 + 1e-9
    assert abs(w.mean()) < 2e-4
    assert abs(w.std() - 0.01 * 0.8796) < 2e-4  # std of N(0,1) truncated at 2 sigma


def test_trainlist_quirk():
    u = np.array([0, 0, 2, 2, 3])  # user 1 missing
    i = np.array([5, 6, 7, 8, 9])
    q = O.train_list_reference(u, i, quirk=True)   # Dataset.py:316-320: cursor advances one per line
    assert q == [[5, 6], [7], [8], [9]]
    t = O.train_list_reference(u, i, quirk=False)
    assert t == [[5, 6], [], [7, 8], [9]]
