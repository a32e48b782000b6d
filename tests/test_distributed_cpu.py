"""world_size-2 gloo tests of the multi-GPU host logic (shard bounds, all_reduce of positions, top-K merge order).
The compute callable is the ORACLE here (tests may stand it in for the CUDA library); the GPU run of the same
function is tests/test_gpu_distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apr_b200.distributed import evaluate_item_sharded, merge_topk, shard_bounds
from oracle import apr_oracle as O


def test_shard_bounds_cover_exactly():
    for n in (1, 63, 64, 65, 1000, 25815):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                lo, hi = shard_bounds(n, w, r, align=64)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
                if r < w - 1 and hi < n:
                    assert hi % 64 == 0
            assert got == list(range(n))


def test_merge_topk_order():
    ids = np.array([[5, 9, -1, 2, 7, -1]], dtype=np.int32)
    sc = np.array([[3.0, 1.0, -np.inf, 3.0, 1.0, -np.inf]], dtype=np.float32)
    mi, ms = merge_topk(ids, sc, 3)
    assert mi.tolist() == [[2, 5, 7]]  # (score desc, id asc)
    assert ms.tolist() == [[3.0, 3.0, 1.0]]


def _case():
    rng = np.random.RandomState(3)
    U, I, d = 40, 300, 16
    P = rng.randn(U, d).astype(np.float32)
    Q = rng.randn(I + 1, d).astype(np.float32)
    Q[17] = Q[200]
    train = [sorted(set(rng.randint(0, I, 6).tolist())) for _ in range(U)]
    test = rng.randint(0, I, U).astype(np.int32)
    return P, Q, train, test, I


def _oracle_range(P, Q, train, test, lo, hi, k):
    """what engine.eval_fullrank returns for the item range [lo, hi): counts + top-k of the NEGATIVES in range"""
    U = P.shape[0]
    pos = np.zeros(U, np.int32)
    ids = np.full((U, k), -1, np.int32)
    sc = np.full((U, k), -np.inf, np.float32)
    for u in range(U):
        s_t = O.score_pairs(P, Q, [u], [test[u]])[0]
        cands = [c for c in range(lo, hi) if c not in train[u] and c != test[u]]
        if cands:
            s = O.score_pairs(P, Q, np.full(len(cands), u), np.asarray(cands))
            pos[u] = int((s >= s_t).sum())
            order = np.lexsort((np.asarray(cands), -s.astype(np.float64)))[:k]
            ids[u, :order.size] = np.asarray(cands)[order]
            sc[u, :order.size] = s[order]
    return torch.from_numpy(pos), torch.from_numpy(ids), torch.from_numpy(sc)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, Q, train, test, I = _case()
    pos, ids, sc = evaluate_item_sharded(lambda lo, hi: _oracle_range(P, Q, train, test, lo, hi, 5), I, k_top=5)
    if rank == 0:
        torch.save((pos, ids, sc), out)
    dist.destroy_process_group()


def _worker_sharded_table(rank, world, port, out):
    """Every rank holds ONLY its rows of the item table: the held-out scores come from the owning rank through the
    all_reduce of evaluate_item_sharded (x + 0 == x), exactly as on the GPU path (evaluate_item_sharded_cuda)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, Q, train, test, I = _case()
    lo_r, hi_r = shard_bounds(I, world, rank, 128)
    Q_local = Q[lo_r:hi_r].copy()                      # the only item rows this rank may touch

    def held_out(lo, hi):
        s = np.zeros(P.shape[0], np.float32)
        for u in range(P.shape[0]):
            if lo <= test[u] < hi:
                s[u] = O.score_pairs(P, Q_local, [u], [test[u] - lo])[0]
        return torch.from_numpy(s)

    def eval_range(lo, hi, spos):
        U = P.shape[0]
        pos = np.zeros(U, np.int32)
        ids = np.full((U, 5), -1, np.int32)
        sc = np.full((U, 5), -np.inf, np.float32)
        for u in range(U):
            cands = [c for c in range(lo, hi) if c not in train[u] and c != test[u]]
            if cands:
                s = O.score_pairs(P, Q_local, np.full(len(cands), u), np.asarray(cands) - lo)
                pos[u] = int((s >= spos[u].item()).sum())
                order = np.lexsort((np.asarray(cands), -s.astype(np.float64)))[:5]
                ids[u, :order.size] = np.asarray(cands)[order]
                sc[u, :order.size] = s[order]
        return torch.from_numpy(pos), torch.from_numpy(ids), torch.from_numpy(sc)

    pos, ids, sc = evaluate_item_sharded(eval_range, I, k_top=5, held_out_scores=held_out)
    if rank == 0:
        torch.save((pos, ids, sc), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_item_sharded_eval_with_sharded_item_table(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker_sharded_table, args=(2, port, out), nprocs=2, join=True)
    pos, ids, sc = torch.load(out)
    P, Q, train, test, I = _case()
    wpos, wids, wsc = _oracle_range(P, Q, train, test, 0, I, 5)
    assert torch.equal(pos, wpos) and torch.equal(ids, wids) and torch.equal(sc, wsc)


@pytest.mark.timeout(300)
def test_item_sharded_eval_equals_single_range(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    pos, ids, sc = torch.load(out)
    P, Q, train, test, I = _case()
    wpos, wids, wsc = _oracle_range(P, Q, train, test, 0, I, 5)
    assert torch.equal(pos, wpos)
    assert torch.equal(ids, wids)
    assert torch.equal(sc, wsc)
    for u in range(P.shape[0]):  # and the positions are the reference metric (utils.py:253-254)
        p, _, _, _ = O.eval_fullrank_user(P, Q, u, int(test[u]), train[u], I, 1)
        assert p == int(pos[u])
