"""Leave-one-out evaluators over a ``Recommender`` with the reference's names (evaluation.py:23-135).

``evaluate_model`` (top-K by heapq.nlargest semantics) and ``evaluate_apr_mode`` (position metric on the first 100
negatives).  When the ranker is backed by the CUDA tables (has ``rank_batched``) every user is scored in ONE kernel
launch; any other ``Recommender`` is driven through ``rank(users, items)`` per user like the reference does.
Two reference bugs are NOT reproduced: ``testNegatives[idx]`` is not mutated (evaluation.py:59), and
``num_thread`` is accepted but ignored (fork pools cannot share a CUDA context).
"""
import math

import numpy as np


def getHitRatio(ranklist, gtItem):
    for item in ranklist:
        if item == gtItem:
            return 1
    return 0


def getNDCG(ranklist, gtItem):
    for i in range(len(ranklist)):
        if ranklist[i] == gtItem:
            return math.log(2) / math.log(i + 2)
    return 0


def _gt_of(testRatings, idx):
    r = testRatings[idx]
    if isinstance(r, (list, tuple, np.ndarray)):
        return int(r[0]), int(r[1])
    return int(idx), int(r)


def _scores_for(model, users_items):
    """users_items: list of (user, items list).  Returns a list of score arrays."""
    if hasattr(model, "rank_batched"):
        return model.rank_batched(users_items)
    out = []
    for u, items in users_items:
        users = np.full(len(items), u, dtype='int32')
        out.append(np.asarray(model.rank(users, np.array(items))).reshape(-1))
    return out


def _topk_first_inserted(items, scores, K):
    """heapq.nlargest(K, map_item_score, key=map_item_score.get) of evaluation.py:73: the dict collapses repeated ids
    (keeping the first position, last score) and ties keep the earlier-inserted id."""
    seen, uniq, sc = {}, [], []
    for it, s in zip(items, scores):
        if it in seen:
            sc[seen[it]] = s
        else:
            seen[it] = len(uniq)
            uniq.append(it)
            sc.append(s)
    order = np.lexsort((np.arange(len(uniq)), -np.asarray(sc, dtype=np.float64)))[:K]
    return [uniq[o] for o in order]


def evaluate_model(model, testRatings, testNegatives, K, num_thread):
    """evaluation.py:23-51.  Iterates idx from 1 like the reference (evaluation.py:40,47); user id == idx when
    testRatings holds bare items (the 1-based loaders), else the [user, item] pair's user."""
    work, gts = [], []
    for idx in range(1, len(testRatings)):
        u, gt = _gt_of(testRatings, idx)
        work.append((u, list(testNegatives[idx]) + [gt]))
        gts.append(gt)
    scores = _scores_for(model, work)
    hits, ndcgs = [], []
    for (u, items), gt, s in zip(work, gts, scores):
        ranklist = _topk_first_inserted(items, s, K)
        hits.append(getHitRatio(ranklist, gt))
        ndcgs.append(getNDCG(ranklist, gt))
    return (hits, ndcgs)


def evaluate_apr_mode(model, testRatings, testNegatives):
    """evaluation.py:93-135: position = #(neg >= pos) over the first 100 negatives; HR@k / NDCG@k for k = 1..100."""
    K = 100
    work = []
    for idx in range(len(testRatings)):
        u, gt = _gt_of(testRatings, idx)
        work.append((u, list(testNegatives[idx][:100]) + [gt]))
    scores = _scores_for(model, work)
    hits, ndcgs = [], []
    for s in scores:
        position = int((s[:-1] >= s[-1]).sum())
        hits.append([position < k for k in range(1, K + 1)])
        ndcgs.append([math.log(2) / math.log(position + 2) if position < k else 0 for k in range(1, K + 1)])
    return (hits, ndcgs)
