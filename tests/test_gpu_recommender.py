"""The Recommender plug-in surface (Recommender.py:3-27) on the CUDA engine."""
import numpy as np
import pytest
import torch

from oracle import apr_oracle as O

pytestmark = pytest.mark.gpu


def _toy(rng, U=80, I=60):
    from apr_b200.Dataset import ArrayDataset
    tu = np.repeat(np.arange(U), 8)
    ti = np.concatenate([rng.choice(I - 1, 8, replace=False) + 1 for _ in range(U)])
    return ArrayDataset(tu, ti, np.arange(U), rng.randint(1, I, U))


@pytest.mark.parametrize("cls_name", ["BPRRecommender", "FastAdversarialMF"])
def test_recommender_train_rank_save(cuda_device, tmp_path, cls_name):
    from apr_b200.FastAdversarialMF import FastAdversarialMF
    from apr_b200.MFRecommender import BPRRecommender
    from apr_b200.Recommender import Recommender
    rng = np.random.RandomState(0)
    ds = _toy(rng)
    r = BPRRecommender(ds.num_users, ds.num_items, 16) if cls_name == "BPRRecommender" else \
        FastAdversarialMF(ds.num_users, ds.num_items, 16, weight=1.0, pop_percent=0.2)
    assert isinstance(r, Recommender)
    x, y = r.get_train_instances(ds.trainMatrix)          # ([users, pos, neg], labels)  BPR.py:83-99
    assert len(x) == 3 and x[0].shape == x[1].shape == x[2].shape == y.shape
    assert all((int(u), int(j)) not in ds.trainMatrix for u, j in zip(x[0], x[2])) and x[2].min() >= 1
    P0 = r.mf.embedding_P.cpu().numpy().copy()
    Q0 = r.mf.embedding_Q.cpu().numpy().copy()
    loss = r.train(x, y, 64)
    # oracle on the same instances, batches in order, tail dropped
    n = (x[0].shape[0] // 64) * 64
    aP, aQ = np.full_like(P0, 0.1), np.full_like(Q0, 0.1)
    adver = 1 if cls_name == "FastAdversarialMF" else 0
    tot = 0.0
    for s in range(0, n, 64):
        l, _ = O.loss_acc(P0, Q0, x[0][s:s + 64], x[1][s:s + 64], x[2][s:s + 64])
        tot += l
        O.apr_step(P0, Q0, aP, aQ, x[0][s:s + 64], x[1][s:s + 64], x[2][s:s + 64], 0.05, 0.0, 1.0, 0.5, adver)
    assert abs(loss - tot / n) <= 1e-4 * abs(tot / n)
    assert np.abs(r.mf.embedding_P.cpu().numpy() - P0).max() <= 2e-4 * np.abs(P0).max()
    # rank == pinned-order scores of the current tables
    users = np.array([3, 3, 7, 9], dtype=np.int32)
    items = np.array([1, 5, 2, 59], dtype=np.int32)
    got = r.rank(users, items)
    want = O.score_pairs(r.mf.embedding_P.cpu().numpy(), r.mf.embedding_Q.cpu().numpy(), users, items)
    assert np.array_equal(np.asarray(got).view(np.uint32), want.view(np.uint32))
    path = str(tmp_path / "w.npz")
    r.save(path)
    r2 = BPRRecommender(ds.num_users, ds.num_items, 16)
    r2.load_pre_train(path)
    assert np.array_equal(r2.rank(users, items), got)
    assert isinstance(r.get_params(), str)
    if cls_name == "FastAdversarialMF":
        r.init(x[0], x[1])                                  # FastAdversarialMF.py:83-85,129-145
        assert len(r.popular_item_x) == int(len(np.unique(x[1])) * 0.2)


def test_evaluation_module_batched_on_gpu(cuda_device):
    from apr_b200 import evaluation
    from apr_b200.MFRecommender import BPRRecommender
    rng = np.random.RandomState(1)
    ds = _toy(rng)
    r = BPRRecommender(ds.num_users, ds.num_items, 16)
    P, Q = r.mf.embedding_P.cpu().numpy(), r.mf.embedding_Q.cpu().numpy()
    test = [ds.testRatings[u][1] for u in range(ds.num_users)]
    negs = [rng.randint(0, ds.num_items, 99).tolist() for _ in range(ds.num_users)]
    hits, ndcgs = evaluation.evaluate_model(r, test, negs, 10, 1)      # one launch for all users (rank_batched)
    ohits, ondcgs = O.evaluate_model_topk(P, Q, {u: test[u] for u in range(1, ds.num_users)}, negs, 10)
    assert hits == ohits and np.allclose(ndcgs, ondcgs)


def test_all_item_scorer_and_recommend(cuda_device):
    """N4: `all_rating = u . Q^T` (IRGAN.py:36-39, APL.py:205-211) bit-exact against the oracle's pinned score
    order, and top-k recommendation over the whole catalogue (tensor-core path) == argsort of those scores with ties to
    the smaller id and the excluded items left out."""
    from apr_b200.MFRecommender import BPRRecommender
    rng = np.random.RandomState(8)
    U, I, d, K = 60, 2500, 64, 20
    rec = BPRRecommender(U, I, d)
    P = rec.mf.embedding_P.cpu().numpy() * 50
    Q = rec.mf.embedding_Q.cpu().numpy() * 50
    Q[77] = Q[5]                                               # a tie
    rec.mf.embedding_P.copy_(torch.from_numpy(P)); rec.mf.embedding_Q.copy_(torch.from_numpy(Q))
    users = np.asarray([3, 17, 59, 0], dtype=np.int32)
    got = rec.rank_all(users)
    want = np.stack([O.score_pairs(P, Q, np.full(I, u), np.arange(I)) for u in users])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    part = rec.rank_all(users, 100, 1500)
    assert np.array_equal(part, got[:, 100:1500])
    excl = [sorted(set(rng.randint(0, I, 15).tolist())) for _ in users]
    ptr = np.concatenate([[0], np.cumsum([len(e) for e in excl])]).astype(np.int64)
    ids, sc = rec.recommend(users, K, ptr, np.concatenate(excl).astype(np.int32))
    for k, u in enumerate(users):
        cand = np.setdiff1d(np.arange(I), excl[k])
        order = cand[np.lexsort((cand, -want[k][cand].astype(np.float64)))][:K]
        assert ids[k].tolist() == order.tolist()
        assert np.array_equal(sc[k].view(np.uint32), want[k][order].view(np.uint32))


@pytest.mark.parametrize("kind", ["mf", "bpr"])
def test_keras_models_match_oracle(cuda_device, kind):
    """N3: MF.py `MatrixFactorization` (BCE on the raw dot product) and BPR.py `BPR` (1 - log sigmoid) with
    Keras' dense Adam: five batches through apr_keras_step against oracle.keras_step (every row moves at every batch)."""
    from apr_b200 import engine
    rng = np.random.RandomState(2)
    U, I, d, n = 90, 70, 32, 200
    P = rng.uniform(-0.05, 0.05, (U, d)).astype(np.float32)
    Q = rng.uniform(-0.05, 0.05, (I, d)).astype(np.float32)
    if kind == "mf":
        P *= 8                                      # dot products inside (1e-7, 1): the clip passes gradient for some, not all
        Q *= 8
    dev = cuda_device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    tP, tQ = t(P, torch.float32), t(Q, torch.float32)
    st = [torch.zeros_like(x) for x in (tP, tP, tQ, tQ, tP, tQ)]
    rP, rQ = P.copy(), Q.copy()
    rs = [np.zeros_like(x) for x in (P, P, Q, Q)]
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    want = 0.0
    for step in range(1, 6):
        u, i, j = rng.randint(0, U, n).astype(np.int32), rng.randint(0, I, n).astype(np.int32), rng.randint(0, I, n).astype(np.int32)
        y = rng.randint(0, 2, n).astype(np.float32)
        if kind == "mf":
            want += O.keras_step(rP, rQ, *rs, u, i, y=y, t=step)
            engine.keras_step(tP, tQ, *st, t(u, torch.int32), t(i, torch.int32), y=t(y, torch.float32), t=step, loss_sum=total)
        else:
            want += O.keras_step(rP, rQ, *rs, u, i, j=j, t=step)
            engine.keras_step(tP, tQ, *st, t(u, torch.int32), t(i, torch.int32), j=t(j, torch.int32), t=step, loss_sum=total)
    assert abs(float(total.item()) - want) <= 1e-5 * abs(want)
    for got, ref in ((tP, rP), (tQ, rQ), (st[0], rs[0]), (st[1], rs[1]), (st[2], rs[2]), (st[3], rs[3])):
        g = got.cpu().numpy()
        assert np.abs(g - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-30), kind
    assert int(st[4].count_nonzero().item()) == 0 and int(st[5].count_nonzero().item()) == 0     # gradient scratch left zero
    assert np.abs(rP - P).min() > 0                                                               # dense Adam: every row moved


def test_keras_recommender_surface(cuda_device):
    from apr_b200.BPR import BPR
    from apr_b200.MF import MatrixFactorization
    rng = np.random.RandomState(4)
    ds = _toy(rng)
    for cls in (MatrixFactorization, BPR):
        m = cls(ds.num_users, ds.num_items, 16)
        x, y = m.get_train_instances(ds.trainMatrix)
        assert len(x) == (2 if cls is MatrixFactorization else 3) and len(y) == len(x[0])
        l0 = m.train(x, y, 64)
        for _ in range(30):
            l1 = m.train(x, y, 64)
        assert np.isfinite(l0) and np.isfinite(l1) and l1 < l0            # the loss goes down
        s = m.rank(np.asarray([1, 2, 3]), np.asarray([4, 5, 6]))
        assert s.shape == (3,) and np.all(np.isfinite(s))
