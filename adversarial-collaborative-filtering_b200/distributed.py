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
    """Host statement of the K10 merge (the GPU path uses the ``apr_topk_merge`` kernel): per-shard top-k lists
    [n_users, world*k] (id -1 = padding) -> the global top-k by (score desc, item id asc), the order the single-range
    kernel produces.  Vectorised over users."""
    n = ids.shape[0]
    pad = ids < 0
    key_s = np.where(pad, np.inf, -scores.astype(np.float64))           # padding sorts last
    order = np.lexsort((ids, key_s), axis=-1)[:, :k]
    out_i = np.take_along_axis(ids, order, axis=1).astype(np.int32)
    out_s = np.take_along_axis(scores, order, axis=1).astype(np.float32)
    gone = np.take_along_axis(pad, order, axis=1)
    out_i[gone] = -1
    out_s[gone] = -np.inf
    if out_i.shape[1] < k:                                             # fewer than k entries per user in total
        out_i = np.concatenate([out_i, np.full((n, k - out_i.shape[1]), -1, np.int32)], axis=1)
        out_s = np.concatenate([out_s, np.full((n, k - out_s.shape[1]), -np.inf, np.float32)], axis=1)
    return out_i, out_s


def _merge_any(ids: torch.Tensor, scores: torch.Tensor, k: int):
    if ids.is_cuda:
        from . import engine
        return engine.topk_merge(ids.contiguous(), scores.contiguous(), k)        # K10 on the GPU
    mi, ms = merge_topk(ids.numpy(), scores.numpy(), k)
    return torch.from_numpy(mi), torch.from_numpy(ms)


def evaluate_item_sharded(eval_range: Callable, num_items: int, k_top: int = 0, group=None,
                          held_out_scores: Optional[Callable[[int, int], torch.Tensor]] = None):
    """Item-sharded full-rank evaluation (utils.py:221-267 over a partitioned catalogue).

    ``eval_range(lo, hi)`` -> (position[int32 n], topk_ids[n,k]|None, topk_scores[n,k]|None) for the item range
    [lo, hi).  When a rank holds only its shard of the item table it cannot score a held-out item that lives elsewhere:
    ``held_out_scores(lo, hi)`` then returns, per user, score(u, test_item[u]) if lo <= test_item[u] < hi else 0, ONE
    all_reduce(sum) makes the exact scores known everywhere (x + 0 == x) and ``eval_range(lo, hi, spos)`` receives
    them.  Returns the global (position, topk_ids, topk_scores) on every rank: positions are exact sums (second
    all_reduce), the per-shard top-k lists are all_gathered and merged by (score desc, id asc) -- on the GPU by the
    ``apr_topk_merge`` kernel."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(num_items, world, rank, align=128)
    if hi <= lo:  # more ranks than 128-item blocks: this rank has nothing to score
        lo = hi = 0
    if held_out_scores is not None:
        spos = held_out_scores(lo, hi)
        if world > 1:
            dist.all_reduce(spos, op=dist.ReduceOp.SUM, group=group)
        pos, ids, sc = eval_range(lo, hi, spos)
    else:
        pos, ids, sc = eval_range(lo, hi)
    if world == 1:
        return pos, ids, sc
    dist.all_reduce(pos, op=dist.ReduceOp.SUM, group=group)
    if not k_top:
        return pos, None, None
    n = ids.shape[0]
    gi = torch.empty((world * n, k_top), dtype=ids.dtype, device=ids.device)
    gs = torch.empty((world * n, k_top), dtype=sc.dtype, device=sc.device)
    dist.all_gather_into_tensor(gi, ids.contiguous(), group=group)
    dist.all_gather_into_tensor(gs, sc.contiguous(), group=group)
    gi, gs = gi.view(world, n, k_top), gs.view(world, n, k_top)
    mi, ms = _merge_any(gi.permute(1, 0, 2).reshape(n, world * k_top), gs.permute(1, 0, 2).reshape(n, world * k_top), k_top)
    return pos, mi, ms


def evaluate_item_sharded_cuda(P, Q, users, test_item, num_items, excl_ptr, excl_idx, k_top=0, group=None,
                               q_row_offset: Optional[int] = None, exact: bool = False, cache_q: bool = False):
    """CUDA front end.  ``Q`` is either the whole item table (``q_row_offset`` None) or ONLY this rank's shard: the rows
    [q_row_offset, q_row_offset + Q.shape[0]) with q_row_offset == shard_bounds(num_items, world, rank, 128)[0], so that
    a 10M x 256 catalogue costs 10.2 GB / world per GPU.  Every rank holds the user rows it evaluates (``P``; callers with
    row-sharded user tables all_gather one user tile at a time) and scores them against its item range on the tensor cores
    (``apr_eval_fullrank_tc_topk``; ``exact`` forces the fp32 CUDA-core kernel, which needs the whole table)."""
    from . import engine
    n = users.numel()
    dev = P.device
    off = 0 if q_row_offset is None else int(q_row_offset)
    use_tc = (not exact) and engine.tc_supported(P.shape[1])
    if q_row_offset is not None and not use_tc:
        raise ValueError("a sharded item table needs the tensor-core path (d % 8 == 0, d <= 256, exact=False)")

    def empty():
        z = torch.zeros(n, dtype=torch.int32, device=dev)
        if k_top:
            return (z, torch.full((n, k_top), -1, dtype=torch.int32, device=dev),
                    torch.full((n, k_top), float("-inf"), dtype=torch.float32, device=dev))
        return z, None, None

    def held_out_scores(lo, hi):
        # score(u, test_item[u]) where this rank owns the row, 0 elsewhere -- no host synchronisation (masked, not compacted)
        if hi <= lo:
            return torch.zeros(n, dtype=torch.float32, device=dev)
        mine = (test_item >= lo) & (test_item < hi)
        items = torch.where(mine, test_item, torch.full_like(test_item, lo))
        s = engine.score_pairs(P, Q, users, items, q_row_offset=off)
        return torch.where(mine, s, torch.zeros_like(s))

    def eval_range(lo, hi, spos=None):
        if hi <= lo:
            return empty()
        if use_tc:
            r = engine.eval_fullrank_tc(P, Q, users, test_item, lo, hi, excl_ptr, excl_idx, check=False, k_top=k_top,
                                        cache_q=cache_q, spos=spos, q_row_offset=off)
            return (r[0], r[1], r[2]) if k_top else (r[0], None, None)
        return engine.eval_fullrank(P, Q, users, test_item, lo, hi, excl_ptr, excl_idx, k_top, exact=True)

    return evaluate_item_sharded(eval_range, num_items, k_top, group,
                                 held_out_scores if (q_row_offset is not None) else None)


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


def trainer_schedule(n: int, Sc: int, G: int, ramp: bool = False, split: bool = False):
    """Sub-chunks (index, first step, steps) of a ShardedTrainer call of n steps, at most Sc steps each (the index
    preparation's scratch covers Sc steps).  Default: Sc steps each.  split: ceil(n / G) steps each, so that every rank
    prepares an equal part of the call at the same time; ramp: 2, 4, ... Sc steps.  Sub-chunk c of call number k belongs to
    rank (c + k) % G.  (Both variants were measured and bring nothing: the head of a cold call is bound by the host
    issuing the side-stream work)"""
    if split and G > 1:
        size = max(1, min(Sc, -(-n // G)))
        return [(c, s0, min(size, n - s0)) for c, s0 in enumerate(range(0, n, size))]
    subs, s0, size = [], 0, (min(2, Sc) if ramp else Sc)
    while s0 < n:
        ns = min(size, n - s0)
        subs.append((len(subs), s0, ns))
        s0, size = s0 + ns, min(Sc, 2 * size)
    return subs


class ShardedTrainer(object):
    """``training_batch`` on row-sharded tables with the index preparation and its exchange taken off the critical path.

    Round 1 ran, per sub-chunk of steps, [index preparation on ONE rank] -> [9 broadcasts of its arrays] -> [the steps],
    all on one stream: every rank idled while one prepared G times its own batch and nine NCCL launches shipped the
    result (the serial fraction grew with G).  Here

      * the ranks' local triples are exchanged (every rank pushes its block into every peer's symmetric-memory staging
        area); a call is cut into sub-chunks of ceil(n / G) steps (at most Sc) owned round-robin: every rank first prepares
        ITS sub-chunks and PACKS their nine regions into one block each (the G ranks work at the same time), then pushes
        the block into every peer's staging slot with plain device-to-device copies over NVLink and announces it through
        the symmetric-memory signal pad; receivers wait for the signal and unpack -- all on a side stream, no collective
        kernel (``APR_TRAINER_EXCHANGE=nccl`` selects all_gather + one broadcast per sub-chunk instead);
      * the step kernels of sub-chunk c wait only for that sub-chunk's "ready" event, so preparation and exchange of
        sub-chunk c+1 (and of the next call: two workspaces alternate) run under the steps of c;
      * the library issues every launch and cross-rank barrier of the steps (``apr_train_steps_sharded``); the barrier
        error flag is checked by ``check()`` / at the end of ``train_steps(check=True)`` and raises.
    """

    REGIONS = ("ucnt", "icnt", "iall", "nslow", "seg_hdr", "rec", "iu_item", "npair", "pairs")

    def __init__(self, tables: ShardedTables, steps_per_call: int, batch_local: int, group=None):
        from . import engine
        self.t, self.group = tables, group
        self.G, self.rank, self.d, self.dev = tables.world, tables.rank, tables.d, tables.device
        self.multi = dist.is_initialized() and dist.get_world_size(group) > 1
        self.S, self.Bl, self.Bg = steps_per_call, batch_local, batch_local * tables.world
        self.ws = [engine.TrainWorkspace(self.S, self.Bg, self.d, self.dev) for _ in range(2)]
        L = engine.train_layout(self.S, self.Bg, self.d)
        self.Sc = L["Sc"]
        Bg = self.Bg
        step_bytes = {"ucnt": 4, "icnt": 4, "iall": 4, "nslow": 4, "seg_hdr": 32 * Bg, "rec": 16 * Bg, "iu_item": 4 * Bg,
                      "npair": 4, "pairs": 48 * (Bg // 2 + 1)}
        self.regions = [(L[n], step_bytes[n]) for n in self.REGIONS]
        self.bytes_per_step = sum(-(-b // 16) * 16 for _, b in self.regions)
        self.stage_bytes = self.bytes_per_step * self.Sc + 256        # one packed sub-chunk
        self.stage, self.stage_own = None, []
        self.side = torch.cuda.Stream(device=self.dev)
        # Exchange transport.  "symm" (default with peer-mapped tables): the packed block of a sub-chunk and the ranks'
        # local triples are PUSHED into the peers' symmetric-memory staging slots by plain device-to-device copies and
        # announced with the symmetric-memory signal pad (put_signal / wait_signal, stream ordered): copy engines and two
        # one-thread kernels, no collective kernel next to the step kernels, no NCCL call in the loop.  "nccl": all_gather +
        # one broadcast per sub-chunk (2 GPUs, steady state: 417 M triples/s against 430 M).
        import os
        want = os.environ.get("APR_TRAINER_EXCHANGE", "symm")
        # schedule knobs (measured in DESIGN.md section 6): ramped sub-chunks at the head of a call, owners prepare all
        # their sub-chunks before the exchange starts, one sub-chunk of look-ahead on the main stream
        self.ramp = os.environ.get("APR_TRAINER_RAMP", "0") != "0"
        self.split = os.environ.get("APR_TRAINER_SPLIT", "0") != "0"
        self.own_first = os.environ.get("APR_TRAINER_ORDER", "own_first") == "own_first"
        self.lookahead = os.environ.get("APR_TRAINER_LOOKAHEAD", "all")      # "0", "1" (one sub-chunk) or "all"
        self.exchange = "symm" if (self.multi and want == "symm" and tables.handle is not None) else "nccl"
        if self.exchange == "symm":
            import torch.distributed._symmetric_memory as symm_mem
            self.max_subs = max(len(self._schedule(n)) for n in range(1, self.S + 1))
            self.slot_bytes = -(-self.stage_bytes // 1024) * 1024
            self.gather_bytes = -(-(3 * self.S * self.Bl * 4) // 1024) * 1024
            per_parity = self.max_subs * self.slot_bytes + self.G * self.gather_bytes
            self.xbuf = symm_mem.empty(2 * per_parity, dtype=torch.uint8, device=self.dev)
            self.xh = symm_mem.rendezvous(self.xbuf, group if group is not None else dist.group.WORLD)
            self.per_parity = per_parity
        self.free = [None, None]          # recorded on the main stream after the steps of the call that used workspace k
        # Nothing is allocated after construction.  With per-call tensors the third call of a trainer's life needed a
        # third set of global batches while the first call's was still held; whether the caching allocator then had to
        # cudaMalloc (slow and erratic with peer-mapped memory in the process) decided whether a cold 20-step call took
        # 7.5 or 360 ms.  Global batches of a call, per workspace parity:
        self.glob = [[torch.empty((self.S, self.Bg), dtype=torch.int32, device=self.dev) for _ in range(3)] for _ in range(2)]
        if self.multi and self.exchange == "nccl":
            self.stage = torch.empty(self.stage_bytes, dtype=torch.uint8, device=self.dev)        # receive side
            self.loc = torch.empty((self.S, self.Bl), dtype=torch.int32, device=self.dev)
            self.gbuf = torch.empty((self.G * self.S, self.Bl), dtype=torch.int32, device=self.dev)
        if self.multi:
            n_own = max(sum(1 for c, _, _ in self._schedule(n) if (c + k) % self.G == self.rank)
                        for n in range(1, self.S + 1) for k in range(self.G))
            # packed blocks of this rank's own sub-chunks of one call
            self.stage_own = [torch.empty(self.stage_bytes, dtype=torch.uint8, device=self.dev)
                              for _ in range(max(1, n_own if self.own_first else 1))]
        self.calls = 0

    # symmetric staging layout per call parity k: [max_subs sub-chunk slots][G gather blocks (one per source rank)]
    def _slot(self, rank, k, c, nbytes):
        off = k * self.per_parity + c * self.slot_bytes
        return self.xh.get_buffer(rank, (nbytes,), torch.uint8, off)

    def _gather_block(self, rank, k, src, n):
        off = k * self.per_parity + self.max_subs * self.slot_bytes + src * self.gather_bytes
        return self.xh.get_buffer(rank, (3, n, self.Bl), torch.int32, off // 4)

    def _schedule(self, n):
        return trainer_schedule(n, self.Sc, self.G if self.multi else 1, self.ramp, self.split)

    SIG_TIMEOUT_MS = 30000   # a peer that never arrives traps after 30 s instead of hanging the GPU

    def _pack(self, ws, s0, ns, unpack: bool, stage=None):
        stage = self.stage if stage is None else stage
        o = 0
        for off, sb in self.regions:
            n = sb * ns
            a = ws.buf[off + s0 * sb: off + s0 * sb + n]
            b = stage[o: o + n]
            if unpack:
                a.copy_(b, non_blocking=True)
            else:
                b.copy_(a, non_blocking=True)
            o += -(-n // 16) * 16
        return o

    def train_steps(self, u_loc: torch.Tensor, i_loc: torch.Tensor, j_loc: torch.Tensor, lr, reg, reg_adv, eps, adver,
                    stats: Optional[torch.Tensor] = None, check: bool = False, inputs_ready: Optional[torch.cuda.Event] = None):
        """n <= steps_per_call steps; ``*_loc`` = THIS rank's triples [n, batch_local], int32, on the device or in
        (pinned) host memory.  Device inputs must be complete when the call is made, or ``inputs_ready`` names the event
        that says so (the side stream waits for it -- not the main stream, which is busy with the previous call's steps)."""
        from . import engine
        n = u_loc.shape[0]
        assert n <= self.S and u_loc.shape[1] == self.Bl
        k = self.calls & 1
        self.calls += 1
        ws = self.ws[k]
        main = torch.cuda.current_stream(self.dev)
        side = self.side
        if self.free[k] is not None:
            side.wait_event(self.free[k])                 # the steps that read this workspace two calls ago are done
        if inputs_ready is not None:
            side.wait_event(inputs_ready)
        ready = []
        with torch.cuda.stream(side):
            glob = [g[:n] for g in self.glob[k]]
            if self.exchange == "symm":
                # every rank pushes its [3, n, Bl] block into block `rank` of every peer's gather area, then signals
                mine3 = self._gather_block(self.rank, k, self.rank, n)
                for q, x in enumerate((u_loc, i_loc, j_loc)):
                    mine3[q].copy_(x, non_blocking=True)
                for r in range(self.G):
                    if r != self.rank:
                        self._gather_block(r, k, self.rank, n).copy_(mine3, non_blocking=True)
                        self.xh.put_signal(r, 1, self.SIG_TIMEOUT_MS)
                for r in range(self.G):
                    if r != self.rank:
                        self.xh.wait_signal(r, 1, self.SIG_TIMEOUT_MS)
                # the G blocks [3, n, Bl] of this rank's gather area (gather_bytes apart) -> global batches [n, G * Bl]
                off = (k * self.per_parity + self.max_subs * self.slot_bytes) // 4
                area = self.xh.get_buffer(self.rank, (self.G, self.gather_bytes // 4), torch.int32, off)
                allb = area[:, :3 * n * self.Bl].view(self.G, 3, n, self.Bl)
                for q in range(3):
                    glob[q].view(n, self.G, self.Bl).copy_(allb[:, q].permute(1, 0, 2))
            else:
                for q, x in enumerate((u_loc, i_loc, j_loc)):
                    if self.multi:
                        self.loc[:n].copy_(x, non_blocking=True)
                        dist.all_gather_into_tensor(self.gbuf[:self.G * n], self.loc[:n], group=self.group)
                        glob[q].view(n, self.G, self.Bl).copy_(self.gbuf[:self.G * n].view(self.G, n, self.Bl).permute(1, 0, 2))
                    else:
                        glob[q].copy_(x, non_blocking=True)
            U, I, J = glob
            subs = self._schedule(n)
            mine = {}
            owner = lambda c: (c + self.calls) % self.G      # self.calls is the same number on every rank

            def prepare_own(c, s0, ns):
                self._prepare(U, I, J, ws, n, s0, ns)
                if self.multi:
                    slot = len(mine) if self.own_first else 0
                    self._pack(ws, s0, ns, unpack=False, stage=self.stage_own[slot])
                    mine[c] = self.stage_own[slot]

            # own_first: every rank prepares and packs ITS sub-chunks before the exchange, so the G ranks work at the same
            # time; else the owner prepares sub-chunk c when its turn comes (the others wait for its block)
            if self.own_first or not self.multi:
                for c, s0, ns in subs:
                    if not self.multi or owner(c) == self.rank:
                        prepare_own(c, s0, ns)
            # ONE block per sub-chunk, in order; receivers unpack
            for c, s0, ns in subs:
                if self.multi and not self.own_first and owner(c) == self.rank:
                    prepare_own(c, s0, ns)
                if self.multi and self.exchange == "symm":
                    src = owner(c)
                    nbytes = self.bytes_per_step * ns
                    if src == self.rank:
                        for r in range(self.G):
                            if r != self.rank:
                                self._slot(r, k, c, nbytes).copy_(mine[c][:nbytes], non_blocking=True)
                                self.xh.put_signal(r, 2, self.SIG_TIMEOUT_MS)
                    else:
                        self.xh.wait_signal(src, 2, self.SIG_TIMEOUT_MS)
                        self._pack(ws, s0, ns, unpack=True, stage=self._slot(self.rank, k, c, nbytes))
                elif self.multi:
                    src = owner(c)
                    nbytes = self.bytes_per_step * ns
                    buf = mine[c] if src == self.rank else self.stage
                    gsrc = dist.get_global_rank(self.group, src) if self.group is not None else src
                    dist.broadcast(buf[:nbytes], src=gsrc, group=self.group)
                    if src != self.rank:
                        self._pack(ws, s0, ns, unpack=True, stage=self.stage)
                ev = torch.cuda.Event()
                ev.record(side)
                ready.append((s0, ns, ev))
        for c, (s0, ns, ev) in enumerate(ready):
            # "all" (default): the steps of a call start when the WHOLE call is prepared and exchanged on this rank.  In
            # steady state that costs nothing (the side stream works one call ahead, under the previous call's steps); on a
            # cold pipeline it keeps the cross-rank signalling of the exchange and the cross-rank barriers of the steps
            # from running against each other, which measured as sporadic 10-180 ms stalls.
            ahead = {"0": c, "1": min(c + 1, len(ready) - 1)}.get(self.lookahead, len(ready) - 1)
            main.wait_event(ready[ahead][2])
            engine.train_steps_sharded(self.t.ptrs, self.G, self.rank, self.d, self.S, self.Bg, lr, reg, reg_adv, eps, adver, ws,
                                       s0, ns, self.t.err, None if stats is None else stats)
        e = torch.cuda.Event()
        e.record(main)
        self.free[k] = e
        if check:
            self.check()

    def _prepare(self, U, I, J, ws, n, s0, ns):
        from . import engine
        # the workspace layout is the one of steps_per_call steps whatever this call's length is
        engine.train_prepare_range(self.t.rows_p, self.t.rows_q, U, I, J, ws, s0, ns, clear=True, n_steps=self.S)

    def check(self) -> None:
        """Raise if a cross-rank barrier timed out (a rank was ~2 s late: lazy module load, a hung peer).  The device side
        has already stopped touching the tables (StepCtx::abort); synchronises."""
        if int(self.t.err.item()) != 0:
            raise RuntimeError("row-sharded training: a cross-rank barrier timed out on rank %d; the step was aborted"
                               % self.rank)

    def synchronize(self) -> None:
        self.side.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
