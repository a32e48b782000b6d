"""Statistical known-answer test: the oracle reproduces the reference's logged Epoch-0 line on the Video dataset.

The reference is unseeded and its forked sampler workers share one RNG state, so the logs can only be
matched statistically.  One BPR epoch of oracle.apr_step (adver=0, d=64, lr=0.05, batch 512) with the fork-emulating
sampler must land inside the spread of the two logged runs (out/janEval/Video_{apr,bpr}_*.out:3, stored in the fixture):
    ACC_adv 0.9519-0.9524, |P| 13.34-13.48, |Q| 11.96-12.13, HR@100 0.0363-0.0369, NDCG@100 0.0093-0.0104
This is the only anchor to the reference's OWN outputs that exists (no TensorFlow here, no golden tensors upstream).
"""
import os

import numpy as np
import pytest

from oracle import apr_oracle as O

FIX = os.path.join(os.path.dirname(__file__), "golden", "video_interactions.npz")


def _fullrank_positions(P, Q, test_i, train_lists, num_items):
    """position per user with plain float32 GEMM scores (statistical use only, not the pinned-order path)."""
    U = len(test_i)
    pos = np.zeros(U, np.int64)
    nneg = np.zeros(U, np.int64)
    Qc = Q[:num_items]
    for a in range(0, U, 2048):
        b = min(U, a + 2048)
        S = P[a:b] @ Qc.T
        s_t = np.einsum("ud,ud->u", P[a:b], Q[test_i[a:b]])
        cnt = (S >= s_t[:, None]).sum(1)
        for k, u in enumerate(range(a, b)):
            ex = set(train_lists[u]) if u < len(train_lists) else set()
            ex.add(int(test_i[u]))
            ex = np.fromiter((e for e in ex if e < num_items), dtype=np.int64)
            cnt[k] -= int((S[k, ex] >= s_t[k]).sum())
            nneg[u] = num_items - ex.size
        pos[a:b] = cnt
    return pos, nneg


@pytest.mark.timeout(600)
def test_video_epoch0_matches_reference_logs():
    z = np.load(FIX)
    ds = O.OracleDataset(z["train_u"], z["train_i"], z["test_u"], z["test_i"], quirk=True)
    assert ds.num_users == 31013 and ds.num_items == 23714 and ds.pairs_u.size == 256094
    logged = z["epoch0_logged"]
    d, B, lr = 64, 512, 0.05
    P = O.truncated_normal(ds.num_users + 1, d, 0.01, 2019, 0)
    Q = O.truncated_normal(ds.num_items + 1, d, 0.01, 2019, 1)
    aP, aQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
    train_sets = [set(l) for l in ds.trainList]
    rng = np.random.RandomState(2019)
    Ub, Ib, Jb = O.legacy_fork_epoch(ds.pairs_u, ds.pairs_i, B, ds.num_items, train_sets, rng, workers=26)
    assert Ub.shape == (500, 512)
    _, prev_acc = O.training_loss_acc(P, Q, Ub, Ib, Jb)
    for s in range(Ub.shape[0]):
        # sparse form of oracle.apr_step (identical arithmetic on the touched rows)
        uu, inv_u = np.unique(Ub[s], return_inverse=True)
        ii, inv = np.unique(np.concatenate([Ib[s], Jb[s]]), return_inverse=True)
        Pc, Qc, aPc, aQc = P[uu], Q[ii], aP[uu], aQ[ii]
        O.apr_step(Pc, Qc, aPc, aQc, inv_u, inv[:B], inv[B:], lr, 0.0, 1.0, 0.5, 0)
        P[uu], Q[ii], aP[uu], aQ[ii] = Pc, Qc, aPc, aQc
    _, post_acc = O.training_loss_acc(P, Q, Ub, Ib, Jb)
    nP, nQ = float(np.linalg.norm(P)), float(np.linalg.norm(Q))
    pos, nneg = _fullrank_positions(P, Q, ds.testRatings[:, 1], ds.trainList, ds.num_items)
    res = O.metrics_from_position(pos, nneg, 100).mean(axis=0)
    hr100, ndcg100 = res[0, -1], res[1, -1]
    print("oracle epoch 0: HR@100 %.4f NDCG@100 %.4f ACC %.4f ACC_adv %.4f |P| %.2f |Q| %.2f" %
          (hr100, ndcg100, prev_acc, post_acc, nP, nQ), "logged:", logged.tolist())
    lo, hi = logged.min(axis=0), logged.max(axis=0)
    assert abs(prev_acc - 0.5) < 0.01
    assert lo[3] - 0.003 <= post_acc <= hi[3] + 0.003          # ACC after the epoch
    assert lo[4] - 0.25 <= nP <= hi[4] + 0.25                  # |P|_F
    assert lo[5] - 0.25 <= nQ <= hi[5] + 0.25                  # |Q|_F
    assert lo[0] - 0.006 <= hr100 <= hi[0] + 0.006             # HR@100
    assert lo[1] - 0.003 <= ndcg100 <= hi[1] + 0.003           # NDCG@100
