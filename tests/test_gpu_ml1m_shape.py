"""BASELINE.json configs[0] / north_star "HR@10/NDCG@10 identical to the reference on ml-1m".

The reference ships only the ml-1m TEST file (data/ml-1m-sort.train.rating is a missing blob), so -- as
SURVEY 8(c) prescribes -- the real held-out items (tests/golden/ml1m_test_items.npy, built by
tests/golden/make_ml1m_fixture.py from data/ml-1m-sort.test.rating) are joined with a synthetic ml-1m-shaped train set
(6 040 users, 3 706 items, ~1 M pairs, >= 20 per user, a user's test item never in their train set), batch 512, d = 64,
eps 0.5, reg_adv 1, lr 0.05 (the reference's defaults).

What is bounded: one BPR epoch and one APR epoch (1 940+ steps each).  The sampled triples are bit-identical to the
oracle's sampler.  The tables are compared TEACHER-FORCED: every 97 steps the oracle restarts from the GPU's tables, so
each comparison bounds the drift over 97 consecutive steps (APR at eps = 0.5 is a chaotic map -- DESIGN.md section 4 --
so a free-running epoch is not comparable in fp32).  After each epoch the GPU's rank positions of ALL users on the GPU's
own tables equal the C oracle's positions on those same tables exactly, hence HR@k / NDCG@k are identical for every k.
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import apr_oracle as O
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu

FIX = os.path.join(os.path.dirname(__file__), "golden", "ml1m_test_items.npy")
DRIFT_RTOL = 3e-4      # per 97-step segment, relative to max|table| (see tests/test_gpu_e2e.py)


def _ml1m_shaped_train(test_items, num_items, rng):
    U = test_items.size
    counts = np.clip(np.round(np.exp(rng.normal(4.66, 0.95, U))), 19, 2300).astype(np.int64)   # ml-1m: >= 20 ratings per user
    pop = 1.0 / np.arange(1, num_items + 1) ** 0.8
    perm = rng.permutation(num_items)                  # popularity is not ordered by item id
    tu, ti = [], []
    for u in range(U):
        w = pop.copy()
        w[np.where(perm == test_items[u])[0]] = 0.0    # leave-one-out: the held-out item is not a train item
        items = perm[rng.choice(num_items, size=min(int(counts[u]), num_items - 1), replace=False, p=w / w.sum())]
        tu.append(np.full(items.size, u, np.int64))
        ti.append(np.sort(items))
    return np.concatenate(tu), np.concatenate(ti)


@pytest.mark.timeout(1500)
def test_ml1m_shape_epochs_and_metrics_match_oracle(cuda_device):
    from apr_b200 import engine
    from apr_b200.APR import MF, Session, sampling, shuffle
    from apr_b200.Dataset import ArrayDataset
    from apr_b200.utils import eval_positions, evaluate, init_eval_model, metrics_from_position
    test_items = np.load(FIX)
    U, I, d, B = 6040, 3706, 64, 512
    assert test_items.shape == (U,) and test_items.max() < I
    rng = np.random.RandomState(2019)
    tu, ti = _ml1m_shaped_train(test_items, I, rng)
    assert 950_000 < tu.size < 1_050_000      # ml-1m-sort: 994 169 train pairs
    ds = ArrayDataset(tu, ti, np.arange(U), test_items, num_users=U, num_items=I)
    ods = O.OracleDataset(tu.astype(np.int32), ti.astype(np.int32), np.arange(U, dtype=np.int32), test_items.astype(np.int32))
    args = types.SimpleNamespace(embed_size=d, lr=0.05, reg=0.0, dns=1, adv="grad", eps=0.5, adver=0, reg_adv=1.0, epochs=2,
                                 seed=2019, batch_size=B, eval_mode="all")
    model = MF(U, I, args)
    model.build_graph()
    feed = init_eval_model(ds, args)
    samples = sampling(ds)
    ex_ptr, ex_idx = feed.excl_ptr_h, feed.excl_idx_h
    seg = 97
    with Session() as sess:
        for epoch, adver in ((0, 0), (1, 1)):
            if adver:                                  # BPR -> APR switch: Adagrad slots restart (APR.py:222-232)
                model.adver = 1
                model.reset_optimizer()
            batches = shuffle(samples, B, ds, model, epoch=epoch)
            ou, oi, _, oj = O.sample_epoch(ods.pairs_u, ods.pairs_i, B, I, ods.csr_ptr, ods.csr_idx, 2019, epoch, 1)
            Ud, Id, Jd = batches[0].t, batches[1].t, batches[3].t
            assert np.array_equal(Ud.cpu().numpy(), ou) and np.array_equal(Id.cpu().numpy(), oi)      # bit-exact triples
            assert np.array_equal(Jd.cpu().numpy(), oj)
            S = ou.shape[0]
            assert S == tu.size // B
            worst = 0.0
            for s0 in range(0, S, seg):
                s1 = min(S, s0 + seg)
                P, Q = model.embedding_P.cpu().numpy().copy(), model.embedding_Q.cpu().numpy().copy()
                aP, aQ = model.acc_P.cpu().numpy().copy(), model.acc_Q.cpu().numpy().copy()
                sess.train_steps(model, Ud[s0:s1], Id[s0:s1], Jd[s0:s1], adver=bool(adver))
                for s in range(s0, s1):
                    C.step(P, Q, aP, aQ, ou[s], oi[s], oj[s], 0.05, 0.0, 1.0, 0.5, adver)
                for got, ref in ((model.embedding_P, P), (model.embedding_Q, Q), (model.acc_P, aP), (model.acc_Q, aQ)):
                    err = float(np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max())
                    worst = max(worst, err)
                    assert err <= DRIFT_RTOL, (epoch, s0, err)
            sess.check(model)
            print("epoch %d (%s): worst 97-step drift %.2e of max|table|" % (epoch, "APR" if adver else "BPR", worst))
            # evaluation on the GPU's own tables: every user's position == the C oracle's, so HR@k / NDCG@k are identical
            gP, gQ = model.embedding_P.cpu().numpy(), model.embedding_Q.cpu().numpy()
            pos = eval_positions(model, feed).cpu().numpy()
            pos_exact = eval_positions(model, feed, exact=True).cpu().numpy()
            want = C.positions(gP, gQ, np.arange(U, dtype=np.int32), test_items, I, ex_ptr, ex_idx)
            assert np.array_equal(pos, want) and np.array_equal(pos_exact, want)
            (hr, ndcg, auc), res = evaluate(model, sess, ds, feed, 0, args)
            ores = metrics_from_position(want, feed.n_neg(), 100)
            assert np.array_equal(res, ores)
            ohr, ondcg, _ = ores.mean(axis=0).tolist()
            assert hr[9] == ohr[9] and ndcg[9] == ondcg[9]          # HR@10 / NDCG@10 identical
            print("epoch %d: HR@10 %.4f NDCG@10 %.4f (identical to the oracle on the same tables)" % (epoch, hr[9], ndcg[9]))
        # top-10 ids through the tensor-core path == the oracle's ranking for a sample of users
        dfeed = feed.to_device(model.device)
        _, ids, _, info = engine.eval_fullrank_tc(model.embedding_P, model.embedding_Q, dfeed["users"], dfeed["test"], 0, I,
                                                  dfeed["excl_ptr"], dfeed["excl_idx"], k_top=10)
        ids = ids.cpu().numpy()
        for u in range(0, U, 151):
            p, _, tids, _ = O.eval_fullrank_user(gP, gQ, u, int(test_items[u]), ods.trainList[u], I, 10)
            neg = [t for t in ids[u].tolist() if t >= 0]
            merged = (neg[:p] + [int(test_items[u])] + neg[p:] if p < 10 else neg)[:10]
            assert merged == tids.tolist()[:10], u
