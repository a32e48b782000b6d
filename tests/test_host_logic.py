"""CPU tests of the host-side mirror of the reference interface: loaders, candidate construction, metric and top-K
semantics, CLI defaults.  No CUDA needed."""
import heapq
import math
import random

import numpy as np
import pytest

from apr_b200 import evaluation, run_adv, run_adv_ori
from apr_b200.Dataset import ArrayDataset, HeDataset, OriginalDataset, build_sorted_csr
from apr_b200.utils import init_eval_model, metrics_from_position
from oracle import apr_oracle as O

TRAIN = "0\t5\t1\t10\n0\t7\t1\t11\n0\t5\t1\t12\n2\t3\t1\t5\n2\t9\t0\t6\n2\t1\t1\t7\n3\t2\t1\t1\n"
TEST = "0\t8\t1\t20\n1\t4\t1\t20\n2\t0\t1\t20\n3\t9\t1\t20\n"


def _write(tmp_path, negatives=False):
    p = tmp_path / "toy"
    (tmp_path / "toy.train.rating").write_text(TRAIN)
    (tmp_path / "toy.test.rating").write_text(TEST)
    if negatives:
        (tmp_path / "toy.test.negative").write_text("(0,8)\t1\t2\t3\n(1,4)\t5\t6\t7\n(2,0)\t4\t5\t6\n(3,9)\t1\t3\t5\n")
    return str(p)


def test_original_dataset_semantics(tmp_path):
    ds = OriginalDataset(_write(tmp_path))
    # dok shape = (max uid + 1, max iid + 1) of the TRAIN file (Dataset.py:278-304)
    assert (ds.num_users, ds.num_items) == (4, 10)
    # rating > 0 only, duplicates collapse, insertion order kept (dok keys)
    assert ds.trainMatrix.keys() == [(0, 5), (0, 7), (2, 3), (2, 1), (3, 2)]
    assert (2, 9) not in ds.trainMatrix and (0, 7) in ds.trainMatrix
    # trainList keeps EVERY line (also rating 0, also duplicates) and has the cursor quirk: user 1 is missing from the
    # file, so user 2's first item is filed under user 1 (Dataset.py:316-320)
    assert ds.trainList == [[5, 7, 5], [3], [9, 1], [2]]
    assert OriginalDataset(_write(tmp_path), reproduce_quirk=False).trainList == [[5, 7, 5], [], [3, 9, 1], [2]]
    assert ds.testRatings == [[0, 8], [1, 4], [2, 0], [3, 9]]
    assert ds.df.iid.tolist() == [5, 7, 5, 3, 9, 1, 2]
    ptr, idx = ds.train_csr()
    assert ptr.tolist() == [0, 2, 3, 5, 6] and idx.tolist() == [5, 7, 3, 1, 9, 2]
    # the oracle's loader agrees
    o = O.OracleDataset.from_files(_write(tmp_path))
    assert o.trainList == ds.trainList and o.num_items == ds.num_items
    assert list(zip(o.pairs_u.tolist(), o.pairs_i.tolist())) == ds.trainMatrix.keys()


def test_he_dataset_negatives(tmp_path):
    ds = HeDataset(_write(tmp_path, negatives=True))
    assert ds.testNegatives == [[1, 2, 3], [5, 6, 7], [4, 5, 6], [1, 3, 5]]
    assert len(ds.testRatings) == len(ds.testNegatives)


def test_eval_inputs_all_mode_match_reference_candidates(tmp_path):
    ds = OriginalDataset(_write(tmp_path))
    import types
    feed = init_eval_model(ds, types.SimpleNamespace(eval_mode="all"))
    for u in range(ds.num_users):
        user_input, item_input = feed[u]
        want = O.fullrank_candidates(ds.num_items, ds.trainList[u], ds.testRatings[u][1])  # utils.py:210-215
        assert item_input[:, 0].tolist() == want and (user_input == u).all()
    assert feed.n_neg().tolist() == [len(O.fullrank_candidates(ds.num_items, ds.trainList[u], ds.testRatings[u][1])) - 1
                                     for u in range(ds.num_users)]


def test_eval_inputs_sample_mode_replays_python_random():
    rng = np.random.RandomState(0)
    U, I = 30, 400
    tu = np.repeat(np.arange(U), 12)
    ti = rng.randint(0, I, tu.size)
    ds = ArrayDataset(tu, ti, np.arange(U), rng.randint(0, I, U))
    import types
    feed = init_eval_model(ds, types.SimpleNamespace(eval_mode="sample"))
    iid = ds.iid_column.tolist()
    for u in (0, 7, 29):
        # utils.py:201-209 verbatim: random.seed(2019) per user, random.choice over the train iid column
        random.seed(2019)
        want = []
        for _ in range(100):
            r = random.choice(iid)
            while r in ds.trainList[u] or ds.testRatings[u][1] == r:
                r = random.choice(iid)
            want.append(r)
        got = feed[u][1][:, 0].tolist()
        assert got == want + [ds.testRatings[u][1]]


def test_metrics_match_reference_loop():
    pos = np.array([0, 1, 9, 10, 99, 100, 5000])
    n = np.array([50, 60, 70, 80, 200, 300, 6000])
    res = metrics_from_position(pos, n, 100)
    for r, (p, nn) in enumerate(zip(pos, n)):
        for k in (1, 10, 100):  # utils.py:257-261
            assert res[r, 0, k - 1] == (p < k)
            assert res[r, 1, k - 1] == (math.log(2) / math.log(p + 2) if p < k else 0)
            assert res[r, 2, k - 1] == 1 - (p / nn)
    assert np.array_equal(res, O.metrics_from_position(pos, n, 100))


def test_topk_first_inserted_equals_heapq_on_dict():
    rng = np.random.RandomState(5)
    for _ in range(50):
        items = rng.randint(0, 30, 40).tolist()        # repeated ids collapse in the dict (evaluation.py:67-69)
        scores = np.round(rng.randn(40), 1).tolist()   # many ties
        m = {}
        for it, sc in zip(items, scores):
            m[it] = sc
        want = heapq.nlargest(10, m, key=m.get)        # evaluation.py:73
        assert evaluation._topk_first_inserted(items, scores, 10) == want


class _TableRanker(object):
    """a Recommender-shaped ranker with fixed scores (rank() per user, like any non-CUDA model)"""

    def __init__(self, S):
        self.S = S

    def rank(self, users, items):
        return self.S[users, items]


def test_evaluation_module_against_oracle():
    rng = np.random.RandomState(2)
    U, I, d = 12, 60, 8
    P, Q = rng.randn(U, d).astype(np.float32), rng.randn(I, d).astype(np.float32)
    S = np.stack([O.score_pairs(P, Q, np.full(I, u), np.arange(I)) for u in range(U)])
    testRatings = {u: int(rng.randint(0, I)) for u in range(U)}
    negs = {u: rng.randint(0, I, 20).tolist() for u in range(U)}
    ranker = _TableRanker(S)
    hits, ndcgs = evaluation.evaluate_model(ranker, [testRatings[u] for u in range(U)], [negs[u] for u in range(U)], 5, 1)
    ohits, ondcgs = O.evaluate_model_topk(P, Q, {u: testRatings[u] for u in range(1, U)}, negs, 5)
    assert hits == ohits and np.allclose(ndcgs, ondcgs)
    before = [list(negs[u]) for u in range(U)]
    evaluation.evaluate_model(ranker, [testRatings[u] for u in range(U)], [negs[u] for u in range(U)], 5, 1)
    assert [negs[u] for u in range(U)] == before  # the reference's list-mutation bug (evaluation.py:59) is not reproduced
    h2, n2 = evaluation.evaluate_apr_mode(ranker, [[u, testRatings[u]] for u in range(U)], [negs[u] for u in range(U)])
    for u in range(U):
        p = O.eval_candidates_position(P, Q, u, negs[u][:100] + [testRatings[u]])
        assert h2[u] == [p < k for k in range(1, 101)]


def test_cli_defaults_match_reference():
    a = run_adv.parse_args([])        # run_adv.py:15-54
    assert (a.model, a.dataset, a.batch_size, a.epochs, a.adv_epoch, a.embed_size, a.dns, a.reg, a.lr, a.reg_adv, a.ckpt,
            a.adv, a.eps, a.opath, a.verbose, a.restore) == ("apr", "ml-1m", 512, 2, 1, 64, 1, 0, 0.05, 1, 1, "grad", 0.5,
                                                              "aaa/", 1, None)
    b = run_adv_ori.parse_args([])    # run_adv_ori.py:17-64
    assert (b.model, b.dataset, b.epochs, b.adv_epoch, b.ckpt, b.eval_mode, b.eps_dense, b.eps_conv, b.eps_pos) == \
        ("pop", "fsq11-sort", 10, 0, 10, "sample", 0.5, 0.5, 0.5)


def test_build_sorted_csr():
    ptr, idx = build_sorted_csr([[3, 1, 3], [], [2]])
    assert ptr.tolist() == [0, 2, 2, 3] and idx.tolist() == [1, 3, 2]


def test_trainlist_cursor_closed_form_equals_the_sequential_cursor():
    """csrc/loader.cu computes the reference's trainList row of line k (Dataset.py:316-320: a cursor that advances by at
    most one user per line and never passes the line's uid) as min(k + 1, k + prefix_min(uid_j - j)) -- valid for
    uid-sorted files.  Checked against the sequential cursor on random sorted sequences with gaps and repeats."""
    rng = np.random.RandomState(0)

    def sequential(a):
        c, out = 0, []
        for x in a:
            if c < x:
                c += 1
            out.append(c)
        return np.asarray(out)

    for _ in range(5000):
        n = rng.randint(1, 60)
        a = np.sort(rng.randint(0, rng.randint(1, 40), n))
        k = np.arange(n)
        closed = np.minimum(k + 1, k + np.minimum.accumulate(a - k))
        assert np.array_equal(closed, sequential(a)), a


def test_trainer_schedule_covers_every_step_once():
    """ShardedTrainer's sub-chunk schedule (host arithmetic): contiguous cover of [0, n), at most Sc steps per sub-chunk,
    every variant; the staging slots are sized by the longest schedule over all call lengths."""
    from apr_b200.distributed import trainer_schedule
    for Sc in (1, 2, 3, 8, 32):
        for G in (1, 2, 4, 8):
            for ramp, split in ((False, False), (True, False), (False, True)):
                longest = max(len(trainer_schedule(n, Sc, G, ramp, split)) for n in range(1, 70))
                for n in range(1, 70):
                    subs = trainer_schedule(n, Sc, G, ramp, split)
                    assert [c for c, _, _ in subs] == list(range(len(subs))) and len(subs) <= longest
                    pos = 0
                    for _, s0, ns in subs:
                        assert s0 == pos and 1 <= ns <= Sc
                        pos += ns
                    assert pos == n
                if not ramp and not split:
                    assert len(trainer_schedule(64, Sc, G)) == -(-64 // Sc)
