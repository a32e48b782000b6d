"""CPU oracle for the APR / BPR-MF hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a plain NumPy restatement of the reference's arithmetic for the path named in
BASELINE.json:north_star.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package never does;
it fails loudly when the CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):
  * The reference (TensorFlow-1.x graph code) cannot be imported in this image and ships no
    tests / golden tensors, so parity at the TensorFlow boundary is UNPINNED ("parity unpinned").
  * What pins this oracle instead:
      - the closed-form gradients below are checked against torch autograd of the loss exactly as
        written in APR.py:143-165 (tests/test_oracle.py);
      - Adagrad is checked against torch.optim.Adagrad(initial_accumulator_value=0.1, eps=0);
      - Philox4x32-10 is checked against the Random123 known-answer vectors;
      - a statistical known-answer test reproduces the reference's logged Epoch-0 line on the Video
        dataset (out/janEval/Video_{apr,bpr}_*.out:3) with the forked-RNG sampler emulated
        (tests/test_oracle_video_kat.py).

Every function cites the reference file:line it follows (paths relative to /root/reference).
All embedding math is float32 unless ``dtype=np.float64`` is passed.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

SEED = 2019  # the reference's only seed constant: utils.py:203, Dataset.py:40,88, run_adv_ori.py:189

# ----------------------------------------------------------------------------------------------
# Counter-based RNG: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3").
# The reference has NO seed on the training path (np.random unseeded, APR.py:51,76) so "bit-exact
# sampled indices" is defined against this specification, shared with csrc/philox.cuh.
# ----------------------------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

# stream tags (key word 1); key word 0 is the user seed
STREAM_PERM = 0x50455231  # "PER1"  epoch permutation round keys
STREAM_NEG = 0x4E454731   # "NEG1"  negative item draws
STREAM_INIT = 0x494E4931  # "INI1"  truncated-normal table init
STREAM_ADV = 0x41445631   # "ADV1"  --adv random perturbations


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32)
    c3 = np.asarray(c3, dtype=np.uint32)
    shape = np.broadcast(c0, c1, c2, c3).shape
    c0, c1, c2, c3 = [np.broadcast_to(c, shape).copy() for c in (c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + PHILOX_W0)
            k1 = np.uint32(k1 + PHILOX_W1)
    return c0, c1, c2, c3


def _fmix32(x):
    """murmur3 finaliser on uint32 arrays (round function of the Feistel permutation)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x85EBCA6B)
        x ^= x >> np.uint32(13)
        x *= np.uint32(0xC2B2AE35)
        x ^= x >> np.uint32(16)
    return x


PERM_ROUNDS = 8


def perm_round_keys(seed: int, epoch: int) -> np.ndarray:
    """8 uint32 Feistel round keys for (seed, epoch): two Philox blocks."""
    a = philox4x32_10(0, 0, 0, np.uint32(epoch), np.uint32(seed), np.uint32(STREAM_PERM))
    b = philox4x32_10(1, 0, 0, np.uint32(epoch), np.uint32(seed), np.uint32(STREAM_PERM))
    return np.array([int(w) for w in a] + [int(w) for w in b], dtype=np.uint32)


def perm_bits(n: int) -> int:
    b = max(2, int(n - 1).bit_length())
    return b + (b & 1)


def feistel_permutation(idx, n: int, seed: int, epoch: int) -> np.ndarray:
    """Pseudo-random bijection of [0,n): replaces ``np.random.shuffle(_index)`` (APR.py:51).

    Balanced Feistel network over 2*half bits, 8 rounds, F(R) = fmix32(R ^ key_r) & mask,
    cycle-walking until the value falls in [0,n).  Position t of the epoch takes pair perm(t).
    """
    idx = np.asarray(idx, dtype=np.uint32)
    keys = perm_round_keys(seed, epoch)
    bits = perm_bits(n)
    half = np.uint32(bits // 2)
    mask = np.uint32((1 << (bits // 2)) - 1)
    out = idx.copy()
    todo = np.ones(out.shape, dtype=bool)
    while todo.any():
        x = out[todo]
        L = x >> half
        R = x & mask
        for r in range(PERM_ROUNDS):
            F = _fmix32(R ^ keys[r]) & mask
            L, R = R, L ^ F
        y = (L << half) | R
        out[todo] = y
        todo_idx = np.flatnonzero(todo)
        todo[todo_idx[y < n]] = False
    return out.astype(np.int64)


# ----------------------------------------------------------------------------------------------
# Sampler  (APR.py:30-81)
# ----------------------------------------------------------------------------------------------
def sampling_pairs(train_u: np.ndarray, train_i: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """APR.py:30-36 ``sampling``: enumerate (u,i) of the dok matrix in insertion (file) order,
    duplicates collapsed (a dok key exists once; first insertion fixes its position)."""
    key = train_u.astype(np.int64) * (int(train_i.max()) + 2) + train_i.astype(np.int64)
    _, first = np.unique(key, return_index=True)
    first.sort()
    return train_u[first].astype(np.int32), train_i[first].astype(np.int32)


def build_csr(lists: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """Sorted, de-duplicated CSR of per-user item lists (membership structure for
    ``j in trainList[u]`` APR.py:77 and ``set(trainList[user])`` utils.py:211)."""
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    rows = []
    for r, items in enumerate(lists):
        a = np.unique(np.asarray(items, dtype=np.int32)) if len(items) else np.zeros(0, np.int32)
        rows.append(a)
        ptr[r + 1] = ptr[r] + a.size
    idx = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
    return ptr, idx


def csr_gkey(ptr, idx) -> np.ndarray:
    """Globally sorted int64 keys row * 2^32 + item of a sorted CSR (rows ascending, items ascending)."""
    row_of = np.repeat(np.arange(ptr.size - 1, dtype=np.int64), np.diff(ptr))
    return (row_of << 32) + idx.astype(np.int64)


def _in_gkey(gkey, n_rows, users, items) -> np.ndarray:
    """Vectorised membership test items[k] in row users[k]; users >= n_rows have empty rows."""
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    out = np.zeros(users.shape, dtype=bool)
    if gkey.size == 0:
        return out
    key = (users << 32) + items
    pos = np.searchsorted(gkey, key)
    ok = pos < gkey.size
    out[ok] = gkey[pos[ok]] == key[ok]
    return out & (users < n_rows)


def _in_csr(ptr, idx, users, items) -> np.ndarray:
    return _in_gkey(csr_gkey(ptr, idx), ptr.size - 1, users, items)


MAX_NEG_ATTEMPTS = 1 << 16


def fork_counter(f: np.ndarray, num_batch: int, batch_size: int, dns: int, workers: int) -> np.ndarray:
    """Counter remap that emulates the reference's forked sampler (APR.py:51-56): Pool(workers) forked
    AFTER np.random.shuffle, so every worker starts from the parent's RNG state and pool.map hands out chunks of
    ceil(num_batch / (4*workers)) batches -- the W chunks of one round draw from the SAME random stream.  A draw with
    epoch index f in chunk c = f // chunk_draws therefore takes the counter of round c // W at its offset in the chunk."""
    chunk_draws = np.uint64(max(1, -(-num_batch // (4 * workers))) * batch_size * dns)
    return ((f // chunk_draws) // np.uint64(workers)) * chunk_draws + (f % chunk_draws)


def sample_epoch(pairs_u: np.ndarray, pairs_i: np.ndarray, batch_size: int, num_items: int,
                 csr_ptr: np.ndarray, csr_idx: np.ndarray, seed: int, epoch: int, dns: int = 1, fork_workers: int = 0):
    """Counter-based restatement of ``shuffle`` + ``_get_train_batch`` (APR.py:39-81).

    Returns (u[S,B], i[S,B], u_dns[S,B*dns], j[S,B*dns]) int32.  Tail batch dropped (APR.py:52).
    Negative f = t*dns + k draws j = mulhi32(word, num_items) from Philox counter
    (f_lo, f_hi, attempt>>2, epoch), key (seed, STREAM_NEG), word index attempt&3, until
    j not in trainList[u]  (APR.py:76-78: range [0,num_items), held-out item MAY be drawn).
    ``fork_workers`` > 0: the negatives' counters are remapped by ``fork_counter`` (the reference's fork-duplicated
    streams, statistically); 0 = independent draws.
    """
    n = pairs_u.shape[0]
    S = n // batch_size
    T = S * batch_size
    t = np.arange(T, dtype=np.uint32)
    pair = feistel_permutation(t, n, seed, epoch)
    u = pairs_u[pair].astype(np.int32)
    i = pairs_i[pair].astype(np.int32)
    f = (np.arange(T, dtype=np.uint64)[:, None] * np.uint64(dns) + np.arange(dns, dtype=np.uint64)[None, :]).ravel()
    if fork_workers > 0:
        f = fork_counter(f, S, batch_size, dns, fork_workers)
    uu = np.repeat(u, dns)
    j = np.full(f.shape, -1, dtype=np.int64)
    gkey = csr_gkey(csr_ptr, csr_idx)
    todo = np.arange(f.size)
    attempt = 0
    while todo.size:
        if attempt >= MAX_NEG_ATTEMPTS:
            raise RuntimeError("negative sampling did not terminate (user has every item?)")
        ff = f[todo]
        w = philox4x32_10((ff & np.uint64(0xFFFFFFFF)).astype(np.uint32), (ff >> np.uint64(32)).astype(np.uint32),
                          np.uint32(attempt >> 2), np.uint32(epoch), np.uint32(seed), np.uint32(STREAM_NEG))
        word = w[attempt & 3]
        cand = ((word.astype(np.uint64) * np.uint64(num_items)) >> np.uint64(32)).astype(np.int64)
        rej = _in_gkey(gkey, csr_ptr.size - 1, uu[todo], cand)
        acc = ~rej
        j[todo[acc]] = cand[acc]
        todo = todo[rej]
        attempt += 1
    return (u.reshape(S, batch_size), i.reshape(S, batch_size),
            uu.reshape(S, batch_size * dns).astype(np.int32), j.reshape(S, batch_size * dns).astype(np.int32))


def select_dns(P, Q, u_dns, j_dns, dns: int) -> np.ndarray:
    """utils.py:121-139: for each positive pick the argmax-scored of its ``dns`` sampled negatives
    (first maximum wins, np.argmax)."""
    s = score_pairs(P, Q, u_dns, j_dns).reshape(-1, dns)
    k = np.argmax(s, axis=1)
    return np.asarray(j_dns).reshape(-1, dns)[np.arange(s.shape[0]), k].astype(np.int32)


# ----------------------------------------------------------------------------------------------
# Table init  (APR.py:105-119)
# ----------------------------------------------------------------------------------------------
def truncated_normal(n_rows: int, d: int, stddev: float, seed: int, table_id: int,
                     stream: int = STREAM_INIT) -> np.ndarray:
    """tf.truncated_normal(mean=0, stddev) restated on Philox (APR.py:107-112): element e, attempt a
    -> Philox counter (e_lo, e_hi, a, table_id), key (seed, stream) -> 4 Box-Muller normals
    (z0,z1 from words 0,1; z2,z3 from words 2,3); first with |z| <= 2 wins; else next attempt."""
    n = n_rows * d
    e = np.arange(n, dtype=np.uint64)
    out = np.zeros(n, dtype=np.float32)
    todo = np.arange(n)
    a = 0
    two_pi = np.float32(6.283185307179586)
    while todo.size:
        ee = e[todo]
        w = philox4x32_10((ee & np.uint64(0xFFFFFFFF)).astype(np.uint32), (ee >> np.uint64(32)).astype(np.uint32),
                          np.uint32(a), np.uint32(table_id), np.uint32(seed), np.uint32(stream))
        # uniform in (0,1]: (w + 1) * 2^-32 in float32 arithmetic restated via float64 then cast
        u = [((x.astype(np.float64) + 1.0) * (1.0 / 4294967296.0)).astype(np.float32) for x in w]
        u = [np.maximum(x, np.float32(1e-38)) for x in u]
        z = []
        for a_, b_ in ((u[0], u[1]), (u[2], u[3])):
            r = np.sqrt(np.float32(-2.0) * np.log(a_)).astype(np.float32)
            z.append((r * np.cos(two_pi * b_)).astype(np.float32))
            z.append((r * np.sin(two_pi * b_)).astype(np.float32))
        zz = np.stack(z, axis=1)
        ok = np.abs(zz) <= np.float32(2.0)
        first = np.argmax(ok, axis=1)
        has = ok.any(axis=1)
        sel = zz[np.arange(zz.shape[0]), first]
        out[todo[has]] = sel[has] * np.float32(stddev)
        todo = todo[~has]
        a += 1
    return out.reshape(n_rows, d)


ADAGRAD_INIT = 0.1  # tf.train.AdagradOptimizer default initial_accumulator_value (APR.py:195)


# ----------------------------------------------------------------------------------------------
# The step  (APR.py:121-195, utils.py:106-119)
# ----------------------------------------------------------------------------------------------
def _softplus(x):
    # tf.nn.softplus(x) = log(1+exp(x)), evaluated stably
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def forward(P, Q, u, i, j, dP=None, dQ=None):
    """APR.py:121-150 (plain) / APR.py:130-141,158-162 (with deltas).  Returns x, r, loss_terms."""
    dt = P.dtype
    p, q, n = P[u], Q[i], Q[j]
    if dP is not None:
        p = p + dP[u]
        q = q + dQ[i]
        n = n + dQ[j]
    y_pos = (p * q).sum(axis=1, dtype=dt)
    y_neg = (p * n).sum(axis=1, dtype=dt)
    x = y_pos - y_neg
    r = np.clip(x, dt.type(-80.0), dt.type(1e8))
    return x, r, _softplus(-r).astype(dt)


def plain_row_gradients(P, Q, u, i, j):
    """Dense G_P, G_Q = d(sum softplus(-clip(x)))/dP, dQ with duplicates summed
    (APR.py:183-187: tf.gradients -> IndexedSlices -> stop_gradient densifies)."""
    dt = P.dtype
    x, r, _ = forward(P, Q, u, i, j)
    m = ((x >= dt.type(-80.0)) & (x <= dt.type(1e8))).astype(dt)
    c = (-m / (dt.type(1.0) + np.exp(r))).astype(dt)
    p, q, n = P[u], Q[i], Q[j]
    GP = np.zeros_like(P)
    GQ = np.zeros_like(Q)
    np.add.at(GP, u, c[:, None] * (q - n))
    np.add.at(GQ, i, c[:, None] * p)
    np.add.at(GQ, j, -c[:, None] * p)
    return GP, GQ, c, x


def l2_normalize_eps(G, eps):
    """tf.nn.l2_normalize(G, 1) * eps with epsilon=1e-12 (APR.py:190-191)."""
    dt = G.dtype
    ss = (G * G).sum(axis=1, keepdims=True, dtype=dt)
    return (G / np.sqrt(np.maximum(ss, dt.type(1e-12))) * dt.type(eps)).astype(dt)


def step_gradients(P, Q, u, i, j, reg, reg_adv, eps, adver, dP=None, dQ=None):
    """Per-row total gradients of opt_loss (APR.py:153-165) with Delta held constant.

    If ``adver`` and dP/dQ are None, Delta is computed from the plain gradient at the CURRENT
    parameters (utils.py:117-119: update_P/update_Q run immediately before the optimizer).
    Returns gP[U,d], gQ[I,d], dP, dQ, info dict.
    """
    dt = P.dtype
    u = np.asarray(u).reshape(-1)
    i = np.asarray(i).reshape(-1)
    j = np.asarray(j).reshape(-1)
    B, d = u.shape[0], P.shape[1]
    GP, GQ, c, x = plain_row_gradients(P, Q, u, i, j)
    p, q, n = P[u], Q[i], Q[j]
    k = dt.type(2.0 * reg * (2 if adver else 1) / (B * d))
    gP = GP.copy()
    gQ = GQ.copy()
    info = {"x": x, "c": c}
    if adver:
        if dP is None:
            dP = l2_normalize_eps(GP, eps)
            dQ = l2_normalize_eps(GQ, eps)
        xa, ra, _ = forward(P, Q, u, i, j, dP, dQ)
        ma = ((xa >= dt.type(-80.0)) & (xa <= dt.type(1e8))).astype(dt)
        ca = (-ma / (dt.type(1.0) + np.exp(ra))).astype(dt) * dt.type(reg_adv)
        pd, qd, nd = p + dP[u], q + dQ[i], n + dQ[j]
        np.add.at(gP, u, ca[:, None] * (qd - nd))
        np.add.at(gQ, i, ca[:, None] * pd)
        np.add.at(gQ, j, -ca[:, None] * pd)
        info.update({"x_adv": xa, "c_adv": ca})
    if reg != 0:
        np.add.at(gP, u, k * p)
        np.add.at(gQ, i, k * q)
        np.add.at(gQ, j, k * n)
    return gP, gQ, dP, dQ, info


def adagrad_apply(W, A, g, lr):
    """TF1 AdagradOptimizer sparse apply after duplicate summation (APR.py:195):
    A += g^2 ; W -= lr * g / sqrt(A); rows with g == 0 are unchanged."""
    dt = W.dtype
    touched = np.flatnonzero(np.any(g != 0, axis=1))
    A[touched] += g[touched] * g[touched]
    W[touched] -= dt.type(lr) * g[touched] / np.sqrt(A[touched])


STREAM_ADV = 0x41445631


def random_delta(n_rows: int, d: int, eps: float, seed: int, step: int, table: int) -> np.ndarray:
    """``--adv random`` (APR.py:170-177; shape-consistent form evaluation_adv.py:182-189):
    Delta = eps * l2_normalize(truncated_normal([rows, d], 0, 0.01), 1), drawn afresh at every step.  The draw of
    global step ``step`` for table ``table`` (0: P, 1: Q) is the counter-based table (2*step + table) of stream ADV."""
    z = truncated_normal(n_rows, d, 0.01, seed, ((step << 1) | table) & 0xFFFFFFFF, STREAM_ADV)
    return l2_normalize_eps(z, eps)


def apr_step_random(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, seed, step):
    """One batch of ``training_batch`` with ``--adv random``: update_P/update_Q assign the noise Delta
    (APR.py:176-177), then the optimizer runs on L + reg_adv L_adv + reg-terms (APR.py:158-165,195)."""
    dP = random_delta(P.shape[0], P.shape[1], eps, seed, step, 0)
    dQ = random_delta(Q.shape[0], Q.shape[1], eps, seed, step, 1)
    gP, gQ, dP, dQ, info = step_gradients(P, Q, u, i, j, reg, reg_adv, eps, 1, dP, dQ)
    adagrad_apply(P, accP, gP, lr)
    adagrad_apply(Q, accQ, gQ, lr)
    info.update({"gP": gP, "gQ": gQ, "dP": dP, "dQ": dQ})
    return info


def apr_step(P, Q, accP, accQ, u, i, j, lr, reg=0.0, reg_adv=1.0, eps=0.5, adver=1):
    """One batch of ``training_batch`` (utils.py:113-119), in place.  Returns info dict.
    ``adver`` = 3: the dns > 1 branch on an adversarial graph (utils.py:121-139): update_P/update_Q never run there, so
    the optimizer sees opt_loss = L + reg_adv L_adv + 2 reg-terms with Delta == 0."""
    if adver == 3:
        gP, gQ, dP, dQ, info = step_gradients(P, Q, u, i, j, reg, reg_adv, eps, 1, np.zeros_like(P), np.zeros_like(Q))
    else:
        gP, gQ, dP, dQ, info = step_gradients(P, Q, u, i, j, reg, reg_adv, eps, adver)
    adagrad_apply(P, accP, gP, lr)
    adagrad_apply(Q, accQ, gQ, lr)
    info.update({"gP": gP, "gQ": gQ, "dP": dP, "dQ": dQ})
    return info


def loss_acc(P, Q, u, i, j):
    """One batch of ``training_loss_acc`` (utils.py:159-175, output_adv=0):
    returns (sum softplus(-r), count(x > 0))."""
    x, r, sp = forward(P, Q, np.asarray(u).reshape(-1), np.asarray(i).reshape(-1), np.asarray(j).reshape(-1))
    return float(sp.sum(dtype=np.float64)), int((x > 0).sum())


def training_loss_acc(P, Q, U, I, J):
    """utils.py:159-175 over all batches: (sum_b L_b / num_batch, mean_b(mean(x>0)))."""
    S = len(U)
    tl, ac = 0.0, 0.0
    for s in range(S):
        l, nc = loss_acc(P, Q, U[s], I[s], J[s])
        tl += l
        ac += nc / len(np.asarray(U[s]).reshape(-1))
    return tl / S, ac / S


# ----------------------------------------------------------------------------------------------
# Scores with a FIXED summation order (defines "bit-exact positions / top-K ids")
# ----------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    """Exact IEEE-754 binary32 fma(a,b,c) on float32 arrays, emulated through float64 with
    round-to-odd so the final rounding to binary32 is a single correct rounding."""
    a64 = a.astype(np.float64)
    b64 = b.astype(np.float64)
    c64 = c.astype(np.float64)
    prod = a64 * b64                      # exact: 24+24 significant bits
    s = prod + c64                        # RN to 53 bits
    bb = s - prod                         # TwoSum error term
    err = (prod - (s - bb)) + (c64 - bb)
    si = s.view(np.int64).copy()
    inexact = (err != 0) & np.isfinite(s)
    even = (si & 1) == 0
    # move to the neighbouring double in the direction of the true value, making the LSB odd
    toward_larger_mag = ((err > 0) == (s > 0))
    adj = np.where(toward_larger_mag, 1, -1).astype(np.int64)
    fix = inexact & even
    # s == 0 with err != 0 cannot happen (then s would be err); guard anyway
    si = np.where(fix & (s != 0), si + adj, si)
    return si.view(np.float64).astype(np.float32)


def score_pairs(P, Q, users, items) -> np.ndarray:
    """score(u,c) = <P[u],Q[c]> in float32 with the order  acc = fma(P[u][k], Q[c][k], acc),
    k = 0..d-1, acc0 = 0  (the arithmetic of utils.py:246-251 / APR.py:121-128 with a pinned
    summation order; csrc uses the same chain so comparisons are bit-exact)."""
    users = np.asarray(users).reshape(-1)
    items = np.asarray(items).reshape(-1)
    p = P[users].astype(np.float32)
    q = Q[items].astype(np.float32)
    acc = np.zeros(users.shape[0], dtype=np.float32)
    for k in range(P.shape[1]):
        acc = _fma32(p[:, k], q[:, k], acc)
    return acc


def score_user_all(P, Q, user: int, num_items: int) -> np.ndarray:
    p = P[user].astype(np.float32)
    q = Q[:num_items].astype(np.float32)
    acc = np.zeros(num_items, dtype=np.float32)
    for k in range(P.shape[1]):
        acc = _fma32(np.broadcast_to(p[k], acc.shape), q[:, k], acc)
    return acc


# ----------------------------------------------------------------------------------------------
# Evaluation  (utils.py:178-267, evaluation.py:23-135)
# ----------------------------------------------------------------------------------------------
def metrics_from_position(position: np.ndarray, n_neg: np.ndarray, K: int) -> np.ndarray:
    """utils.py:253-261: res[user, {hr,ndcg,auc}, k-1] for k = 1..K."""
    position = np.asarray(position, dtype=np.int64)
    n_neg = np.asarray(n_neg, dtype=np.float64)
    ks = np.arange(1, K + 1)[None, :]
    hit = position[:, None] < ks
    nd = np.log(2.0) / np.log(position.astype(np.float64) + 2.0)
    res = np.zeros((position.shape[0], 3, K), dtype=np.float64)
    res[:, 0, :] = hit
    res[:, 1, :] = np.where(hit, nd[:, None], 0.0)
    res[:, 2, :] = (1.0 - position / n_neg)[:, None]
    return res


def eval_candidates_position(P, Q, user: int, cands: Sequence[int]) -> int:
    """utils.py:244-254: candidates list with the held-out item LAST; position = #(neg >= pos)."""
    s = score_pairs(P, Q, np.full(len(cands), user), np.asarray(cands))
    return int((s[:-1] >= s[-1]).sum())


def fullrank_candidates(num_items: int, train_row: Sequence[int], test_item: int) -> List[int]:
    """utils.py:210-215 (eval_mode == "all"): ascending range(I) minus train minus test, + [test]."""
    s = set(range(num_items)) - set(int(x) for x in train_row)
    s.discard(int(test_item))
    return sorted(s) + [int(test_item)]


def eval_fullrank_user(P, Q, user: int, test_item: int, train_row: Sequence[int], num_items: int, K: int):
    """Full-rank evaluation of one user: (position, n_neg, topk_ids, topk_scores).

    top-K over the negatives+held-out candidate list ordered by (score desc, item id asc) except
    that the held-out item loses ties against every negative (it is appended last:
    utils.py:215 and heapq.nlargest keeps first-inserted on ties, evaluation.py:59,73)."""
    cands = fullrank_candidates(num_items, train_row, test_item)
    s = score_pairs(P, Q, np.full(len(cands), user), np.asarray(cands))
    pos = int((s[:-1] >= s[-1]).sum())
    order_key = np.arange(len(cands))  # ascending ids for negatives, test last
    idx = np.lexsort((order_key, -s.astype(np.float64)))[:K]
    ids = np.asarray(cands)[idx].astype(np.int32)
    return pos, len(cands) - 1, ids, s[idx]


def evaluate_fullrank(P, Q, test_items: Sequence[int], train_lists: Sequence[Sequence[int]], num_items: int,
                      K: int = 100, users: Optional[Sequence[int]] = None):
    """utils.evaluate with eval_mode == "all" (utils.py:221-241): ((hr,ndcg,auc)[K], res[U,3,K])."""
    users = range(len(test_items)) if users is None else users
    pos, nneg = [], []
    for uidx in users:
        p, n, _, _ = eval_fullrank_user(P, Q, uidx, test_items[uidx], train_lists[uidx] if uidx < len(train_lists) else [],
                                        num_items, 1)
        pos.append(p)
        nneg.append(n)
    res = metrics_from_position(np.array(pos), np.array(nneg), K)
    return tuple(res.mean(axis=0).tolist()), res, np.array(pos)


def sampled_candidates_reference(user_test_item: int, train_row: Sequence[int], iid_column: Sequence[int],
                                 n: int = 100) -> List[int]:
    """utils.py:201-209 (eval_mode == "sample"): random.seed(2019) PER USER then n x random.choice
    over the train-file iid column with rejection of train items and the test item."""
    import random
    random.seed(SEED)
    tr = set(int(x) for x in train_row)
    out = []
    for _ in range(n):
        r = random.choice(iid_column)
        while r in tr or user_test_item == r:
            r = random.choice(iid_column)
        out.append(int(r))
    return out


def evaluate_model_topk(P, Q, test_items: Dict[int, int], test_negatives: Dict[int, Sequence[int]], K: int):
    """evaluation.py:23-91 ``evaluate_model``: per user idx (iterated from 1, evaluation.py:40,47)
    candidates = negatives + [gt]; dict de-duplicates ids keeping first position; top-K by
    heapq.nlargest (ties -> earlier inserted); HR = gt in topK; NDCG = ln2/ln(rank+2)."""
    hits, ndcgs = [], []
    for idx in sorted(test_items.keys()):
        gt = int(test_items[idx])
        items = [int(x) for x in test_negatives[idx]] + [gt]
        s = score_pairs(P, Q, np.full(len(items), idx), np.asarray(items))
        seen, uniq_items, uniq_scores = {}, [], []
        for it, sc in zip(items, s):
            if it in seen:
                uniq_scores[seen[it]] = sc
            else:
                seen[it] = len(uniq_items)
                uniq_items.append(it)
                uniq_scores.append(sc)
        us = np.asarray(uniq_scores, dtype=np.float64)
        order = np.lexsort((np.arange(len(uniq_items)), -us))[:K]
        rank = [uniq_items[o] for o in order]
        if gt in rank:
            hits.append(1)
            ndcgs.append(math.log(2) / math.log(rank.index(gt) + 2))
        else:
            hits.append(0)
            ndcgs.append(0)
    return hits, ndcgs


# ----------------------------------------------------------------------------------------------
# Data loading  (Dataset.py:226-327 OriginalDataset, Dataset.py:112-223 HeDataset)
# ----------------------------------------------------------------------------------------------
def read_rating_file(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """TSV ``uid \t iid \t rating \t timestamp``."""
    us, is_, rs = [], [], []
    with open(path, "r") as f:
        for line in f:
            if not line.strip():
                continue
            a = line.split("\t")
            us.append(int(a[0]))
            is_.append(int(a[1]))
            rs.append(float(a[2]) if len(a) > 2 else 1.0)
    return np.asarray(us, np.int32), np.asarray(is_, np.int32), np.asarray(rs, np.float32)


def train_list_reference(train_u: np.ndarray, train_i: np.ndarray, quirk: bool = True) -> List[List[int]]:
    """Dataset.py:306-325 ``load_training_file_as_list``.  With ``quirk`` the user cursor advances by
    at most one per line: when a user id is skipped, the first item of the next user is
    filed under the missing user.  Without it, rows are grouped by the true uid."""
    if not quirk:
        n = int(train_u.max()) + 1 if train_u.size else 0
        lists: List[List[int]] = [[] for _ in range(n)]
        for a, b in zip(train_u.tolist(), train_i.tolist()):
            lists[a].append(b)
        return lists
    u_ = 0
    lists, items = [], []
    for a, b in zip(train_u.tolist(), train_i.tolist()):
        if u_ < a:
            lists.append(items)
            items = []
            u_ += 1
        items.append(b)
    lists.append(items)
    return lists


class OracleDataset:
    """OriginalDataset restated (Dataset.py:235-252): num_users/num_items = dok shape =
    (max train uid + 1, max train iid + 1); rating > 0 kept; testRatings as [user,item] rows."""

    def __init__(self, train_u, train_i, test_u, test_i, train_r=None, quirk=True):
        if train_r is not None:
            keep = train_r > 0
        else:
            keep = np.ones(train_u.shape, bool)
        self.num_users = int(train_u.max()) + 1
        self.num_items = int(train_i.max()) + 1
        self.pairs_u, self.pairs_i = sampling_pairs(train_u[keep], train_i[keep])
        self.trainList = train_list_reference(train_u, train_i, quirk)
        self.testRatings = np.stack([test_u, test_i], axis=1).astype(np.int32)
        self.iid_column = train_i.tolist()
        self.csr_ptr, self.csr_idx = build_csr(self.trainList)

    @classmethod
    def from_files(cls, prefix: str, quirk=True):
        tu, ti, tr = read_rating_file(prefix + ".train.rating")
        eu, ei, _ = read_rating_file(prefix + ".test.rating")
        return cls(tu, ti, eu, ei, tr, quirk)


def read_negative_file(path: str) -> List[List[int]]:
    """Dataset.py:161-172: ``(u,i) \t n1 \t ... \t n99`` -> list of negative lists."""
    out = []
    with open(path, "r") as f:
        for line in f:
            if not line.strip():
                continue
            a = line.rstrip("\n").split("\t")
            out.append([int(x) for x in a[1:]])
    return out


# ----------------------------------------------------------------------------------------------
# Legacy (reference-like) sampler, to explain the logged numbers
# ----------------------------------------------------------------------------------------------
def legacy_fork_epoch(pairs_u, pairs_i, batch_size, num_items, train_sets, rng: np.random.RandomState, workers: int):
    """Emulates APR.py:51-56: ``np.random.shuffle`` in the parent, then a forked Pool(workers) whose
    workers all start from the parent's RNG state; pool.map hands out chunks of
    ceil(num_batch / (4*workers)) batches; chunk c is taken by worker c % workers."""
    n = len(pairs_u)
    index = np.arange(n)
    rng.shuffle(index)
    S = n // batch_size
    chunk = max(1, -(-S // (4 * workers)))
    state = rng.get_state()
    wr = []
    for _ in range(workers):
        r = np.random.RandomState()
        r.set_state(state)
        wr.append(r)
    U = np.zeros((S, batch_size), np.int32)
    I = np.zeros((S, batch_size), np.int32)
    J = np.zeros((S, batch_size), np.int32)
    for s in range(S):
        r = wr[(s // chunk) % workers]
        sel = index[s * batch_size:(s + 1) * batch_size]
        U[s] = pairs_u[sel]
        I[s] = pairs_i[sel]
        for b in range(batch_size):
            tr = train_sets[U[s, b]]
            jj = r.randint(num_items)
            while jj in tr:
                jj = r.randint(num_items)
            J[s, b] = jj
    return U, I, J


# ----------------------------------------------------------------------------------------------
# N3: the Keras Recommender models (MF.py:7-59, BPR.py:23-102) -- parity unpinned (Keras / TF not installable)
# ----------------------------------------------------------------------------------------------
KERAS_EPS = 1e-7   # K.epsilon(): clip of binary_crossentropy and Adam's epsilon (Keras 2.2 defaults)


def keras_step(P, Q, mP, vP, mQ, vQ, u, i, j=None, y=None, lr=0.001, b1=0.9, b2=0.999, t=1):
    """One batch of ``model.fit`` for MF.py (``y`` given: binary_crossentropy(y, <u,i>) with the raw dot product clipped
    to [1e-7, 1-1e-7], MF.py:21-24) or BPR.py (``j`` given: mean(1 - log sigmoid(<u,i> - <u,j>)), BPR.py:11-21), then
    Keras 2.2's Adam on DENSIFIED Embedding gradients: lr_t = lr sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g;
    v = b2 v + (1-b2) g^2; w -= lr_t m / (sqrt(v) + 1e-7) for EVERY row.  In place; returns the batch's summed loss."""
    dt = P.dtype
    u, i = np.asarray(u).reshape(-1), np.asarray(i).reshape(-1)
    n = u.size
    p, q = P[u], Q[i]
    gP, gQ = np.zeros_like(P), np.zeros_like(Q)
    if y is not None:
        y = np.asarray(y, dtype=dt).reshape(-1)
        pred = (p * q).sum(axis=1, dtype=dt)
        ph = np.clip(pred, dt.type(KERAS_EPS), dt.type(1.0 - KERAS_EPS))
        loss = -(y * np.log(ph) + (1 - y) * np.log(1 - ph))
        inside = ((pred >= dt.type(KERAS_EPS)) & (pred <= dt.type(1.0 - KERAS_EPS))).astype(dt)
        c = (inside * (-y / ph + (1 - y) / (1 - ph)) / dt.type(n)).astype(dt)
        np.add.at(gP, u, c[:, None] * q)
        np.add.at(gQ, i, c[:, None] * p)
    else:
        j = np.asarray(j).reshape(-1)
        r = Q[j]
        x = (p * q).sum(axis=1, dtype=dt) - (p * r).sum(axis=1, dtype=dt)
        loss = 1.0 + _softplus(-x)
        c = (-1.0 / (1.0 + np.exp(x)) / dt.type(n)).astype(dt)
        np.add.at(gP, u, c[:, None] * (q - r))
        np.add.at(gQ, i, c[:, None] * p)
        np.add.at(gQ, j, -c[:, None] * p)
    lr_t = dt.type(lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    for W, M, V, g in ((P, mP, vP, gP), (Q, mQ, vQ, gQ)):
        M[:] = dt.type(b1) * M + dt.type(1.0 - b1) * g
        V[:] = dt.type(b2) * V + dt.type(1.0 - b2) * g * g
        W -= lr_t * M / (np.sqrt(V) + dt.type(KERAS_EPS))
    return float(loss.sum(dtype=np.float64))
