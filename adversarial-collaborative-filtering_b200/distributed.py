"""Multi-GPU paths: one process per GPU, torch.distributed for the plumbing.

Evaluation is item-sharded: rank r scores its item range with ``apr_eval_fullrank(item_lo, item_hi)`` and the per-user
counts are summed with ONE all_reduce; the per-shard top-K lists are all_gathered and merged by (score desc, id asc).
No other collective touches the data path.

Training is row-sharded over peer-mapped (NVLink) tables: row r of a table lives on rank ``r % G`` at local row
``r // G``; every rank's kernels address any row through the G base pointers (``apr_train_steps_sharded``), so the
"exchange" is the kernels' own 128-bit loads / stores / vector REDs over NVLink, fused into the step.

The host-side partitioning logic (shard bounds, merge order) is backend-agnostic and is covered on CPU with gloo; the
compute callables are injected so the tests can stand the oracle in for the CUDA library.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank's shard of n items; interior boundaries are multiples of ``align``."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def merge_topk(ids: np.ndarray, scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard top-k lists [n_users, world*k] (id -1 = padding) into the global top-k by
    (score desc, item id asc) -- the order the single-range kernel produces."""
    n = ids.shape[0]
    out_i = np.full((n, k), -1, dtype=np.int32)
    out_s = np.full((n, k), -np.inf, dtype=np.float32)
    for u in range(n):
        ok = ids[u] >= 0
        ii, ss = ids[u][ok], scores[u][ok]
        order = np.lexsort((ii, -ss.astype(np.float64)))[:k]
        out_i[u, :order.size] = ii[order]
        out_s[u, :order.size] = ss[order]
    return out_i, out_s


def evaluate_item_sharded(eval_range: Callable[[int, int], Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]],
                          num_items: int, k_top: int = 0, group=None):
    """Item-sharded full-rank evaluation.

    ``eval_range(lo, hi)`` -> (position[int32 n], topk_ids[n,k]|None, topk_scores[n,k]|None) for the item range
    [lo, hi) (``engine.eval_fullrank`` on the GPU path).  Returns the global (position, topk_ids, topk_scores) on every
    rank; positions are exact sums, the top-k merge reproduces the single-range order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(num_items, world, rank, align=64)
    if hi > lo:
        pos, ids, sc = eval_range(lo, hi)
    else:  # more ranks than 64-item blocks: this rank has nothing to score
        pos, ids, sc = eval_range(0, 0)
    if world == 1:
        return pos, ids, sc
    dist.all_reduce(pos, op=dist.ReduceOp.SUM, group=group)
    if not k_top:
        return pos, None, None
    gi = [torch.empty_like(ids) for _ in range(world)]
    gs = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(gi, ids, group=group)
    dist.all_gather(gs, sc, group=group)
    mi, ms = merge_topk(torch.cat(gi, dim=1).cpu().numpy(), torch.cat(gs, dim=1).cpu().numpy(), k_top)
    return pos, torch.from_numpy(mi).to(pos.device), torch.from_numpy(ms).to(pos.device)


def evaluate_item_sharded_cuda(P, Q, users, test_item, num_items, excl_ptr, excl_idx, k_top=0, group=None):
    """CUDA front end.  Every rank holds the user rows it evaluates and a replica of Q (10 GB even for config 5:
    10M x 256 fp32) but SCORES only its item range, so the GEMM work and the top-K selection are sharded G ways."""
    from . import engine

    def eval_range(lo, hi):
        n = users.numel()
        if hi <= lo:
            z = torch.zeros(n, dtype=torch.int32, device=P.device)
            if k_top:
                return (z, torch.full((n, k_top), -1, dtype=torch.int32, device=P.device),
                        torch.full((n, k_top), float("-inf"), dtype=torch.float32, device=P.device))
            return z, None, None
        return engine.eval_fullrank(P, Q, users, test_item, lo, hi, excl_ptr, excl_idx, k_top)

    return evaluate_item_sharded(eval_range, num_items, k_top, group)


# ---------------------------------------------------------------------------------------------------------
# row-sharded training
# ---------------------------------------------------------------------------------------------------------
class ShardedTables(object):
    """P, Q, their Adagrad accumulators and the shared-item workspace, row-sharded over the ranks of ``group``.

    Row r lives on rank ``r % G`` at local row ``r // G``.  With ``symmetric=True`` the shards are carved from one
    torch symmetric-memory buffer per rank, so every rank holds the peer (NVLink) address of every shard;
    ``symmetric=False`` builds all G shards in this process's own memory (single-GPU emulation used by the tests)."""

    NAMES = ("P", "accP", "Q", "accQ", "GQ", "HQ", "sig")

    def __init__(self, rows_p: int, rows_q: int, d: int, batch_global: int, device, world: int = 1, rank: int = 0,
                 group=None, symmetric: bool = False):
        assert world in (1, 2, 4, 8), "row sharding needs a power-of-two number of ranks <= 8"
        self.rows_p, self.rows_q, self.d, self.world, self.rank, self.device = rows_p, rows_q, d, world, rank, device
        lp, lq, ls = -(-rows_p // world), -(-rows_q // world), batch_global // world + 2
        self.local_rows = {"P": lp, "accP": lp, "Q": lq, "accQ": lq, "GQ": ls, "HQ": ls, "sig": max(1, -(-64 // d))}
        sizes = [self.local_rows[n] * d for n in self.NAMES]  # "sig": >= 64 words reinterpreted as int32 signal slots
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(offs[-1])
        self.handle = None
        if symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(total, dtype=torch.float32, device=device)
            self.handle = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            bases = [int(p) for p in self.handle.buffer_ptrs]
            self.bufs = {rank: buf}
        else:
            self.bufs = {r: torch.empty(total, dtype=torch.float32, device=device) for r in range(world)}
            bases = [self.bufs[r].data_ptr() for r in range(world)]
        self.ptrs = {n: [bases[r] + int(offs[k]) * 4 for r in range(world)] for k, n in enumerate(self.NAMES)}
        self.views = {r: {n: b[int(offs[k]):int(offs[k + 1])].view(self.local_rows[n], d) for k, n in enumerate(self.NAMES)}
                      for r, b in self.bufs.items()}
        for v in self.views.values():
            v["GQ"].zero_()
            v["HQ"].zero_()
            v["sig"].zero_()
        self.err = torch.zeros(1, dtype=torch.int32, device=device)

    def local(self, name: str, rank: Optional[int] = None) -> torch.Tensor:
        return self.views[self.rank if rank is None else rank][name]

    def load_full(self, name: str, full: torch.Tensor) -> None:
        """Scatter the rows of a full (unsharded) table into the shards this process holds."""
        for r, v in self.views.items():
            rows = full[r::self.world]
            v[name][:rows.shape[0]].copy_(rows)

    def gather_full(self, name: str, rows: int) -> torch.Tensor:
        """Inverse of load_full for the shards this process holds (emulation: all of them)."""
        out = torch.empty((rows, self.d), dtype=torch.float32, device=self.device)
        for r, v in self.views.items():
            n = out[r::self.world].shape[0]
            out[r::self.world] = v[name][:n]
        return out


def stream_barrier(group=None, token: Optional[torch.Tensor] = None) -> None:
    """Cross-rank, stream-ordered barrier: a one-element all_reduce on the current stream (no host synchronisation)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(token, group=group)


def train_steps_sharded(tables: ShardedTables, U: torch.Tensor, I: torch.Tensor, J: torch.Tensor, lr, reg, reg_adv, eps,
                        adver, ws, aux_stream=None, group=None, ranks: Optional[Sequence[int]] = None,
                        stats: Optional[torch.Tensor] = None) -> None:
    """``training_batch`` (utils.py:113-119) for GLOBAL batches U,I,J [S, Bg] on row-sharded tables.

    Index preparation of sub-chunk c runs on rank c % G and its arrays are broadcast; every rank then processes every
    G-th segment of each step, with a cross-rank barrier after the plain stage, after the adversarial stage and at the end
    of the step (the Delta of a shared item needs every rank's contribution; the next step must see every update).
    ``ranks`` = the ranks THIS process executes: [rank] normally, all ranks in the single-GPU emulation."""
    from . import engine
    S, Bg = U.shape
    d, G = tables.d, tables.world
    ranks = [tables.rank] if ranks is None else list(ranks)
    multi = dist.is_initialized() and dist.get_world_size(group) > 1 and len(ranks) == 1
    L = engine.train_layout(S, Bg, d)
    token = torch.zeros(1, device=U.device) if multi else None
    wsb = ws.buf
    views = {
        "ucnt": (L["ucnt"], 4, 1), "icnt": (L["icnt"], 4, 1), "iall": (L["iall"], 4, 1), "nslow": (L["nslow"], 4, 1),
        "seg_hdr": (L["seg_hdr"], 32 * Bg, Bg * 8), "rec": (L["rec"], 16 * Bg, Bg * 4), "iu_item": (L["iu_item"], 4 * Bg, Bg),
        # pair work units (two single-triple segments coupled by one twice-occurring item; csrc/train.cu pair_unit)
        "npair": (L["npair"], 4, 1), "pairs": (L["pairs"], 48 * (Bg // 2 + 1), 12 * (Bg // 2 + 1)),
    }
    main = torch.cuda.current_stream()
    for c, s0 in enumerate(range(0, S, L["Sc"])):
        ns = min(L["Sc"], S - s0)
        src = c % G
        if src in ranks or not multi:
            engine.train_prepare_range(tables.rows_p, tables.rows_q, U, I, J, ws, s0, ns, clear=True)
        if multi:
            for off, step_bytes, _ in views.values():
                dist.broadcast(wsb[off + s0 * step_bytes: off + (s0 + ns) * step_bytes], src=dist.get_global_rank(group, src)
                               if group is not None else src, group=group)
        if len(ranks) == 1:
            # real run: the library issues every launch and cross-rank barrier of these steps itself
            engine.train_steps_sharded(tables.ptrs, G, ranks[0], d, S, Bg, lr, reg, reg_adv, eps, adver, ws, s0, ns,
                                       tables.err, stats)
            continue
        for s in range(s0, s0 + ns):
            launch = lambda r, stage: engine.train_stage_sharded(tables.ptrs, G, r, d, S, Bg, lr, reg, reg_adv, eps, adver,
                                                                 ws, s, stage, stats)
            if aux_stream is not None:
                aux_stream.wait_stream(main)
                with torch.cuda.stream(aux_stream):
                    for r in ranks:
                        launch(r, 3)
                        launch(r, 4)
            else:
                for r in ranks:
                    launch(r, 3)
                    launch(r, 4)
            if adver:
                for r in ranks:
                    launch(r, 0)
                stream_barrier(group, token)
            for r in ranks:
                launch(r, 1)
            stream_barrier(group, token)
            for r in ranks:
                launch(r, 2)
            if aux_stream is not None:
                main.wait_stream(aux_stream)
            stream_barrier(group, token)
