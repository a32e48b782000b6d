"""Thin tensor-level wrappers over the C ABI: torch is used for device memory and streams only.

Every function takes CUDA torch tensors (contiguous, float32 / int32 / int64 as documented in include/apr_b200.h),
enqueues on the current torch stream and returns torch tensors.  No function here computes on the CPU.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

STREAM_INIT = 0x494E4931
STREAM_ADV = 0x41445631


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("apr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None) -> int:
    if t is None:
        return 0
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise ValueError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


def device_info() -> Tuple[int, int, int]:
    a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.lib().apr_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def context_stats() -> dict:
    """Graph-cache counters of the current device's library context (see include/apr_b200.h)."""
    out = (ctypes.c_int64 * 4)()
    _lib.check(_lib.lib().apr_context_stats(out))
    return {"graph_instantiations": int(out[0]), "graph_updates": int(out[1]), "graph_launches": int(out[2]),
            "graphs_cached": int(out[3])}


def init_truncated_normal(W: torch.Tensor, stddev: float, seed: int, table_id: int, tag: int = STREAM_INIT) -> None:
    rows, d = W.shape
    _lib.check(_lib.lib().apr_init_truncated_normal(_ptr(W, torch.float32), rows, d, float(stddev), seed & 0xFFFFFFFF,
                                                    table_id & 0xFFFFFFFF, tag, _stream()))
    touch(W)


def fill(x: torch.Tensor, value: float) -> None:
    _lib.check(_lib.lib().apr_fill_f32(_ptr(x, torch.float32), x.numel(), float(value), _stream()))
    touch(x)


def sample_epoch(pairs_u: torch.Tensor, pairs_i: torch.Tensor, batch: int, num_items: int, csr_ptr: torch.Tensor,
                 csr_idx: torch.Tensor, seed: int, epoch: int, dns: int = 1, rank: int = 0, world: int = 1,
                 fork_workers: int = 0):
    """-> (u[S,Bl], i[S,Bl], u_dns[S,Bl*dns], j[S,Bl*dns], err) int32 device tensors (APR.py:39-81).
    ``batch`` is the GLOBAL batch; with world > 1 this rank draws its slice [rank*Bl, (rank+1)*Bl) of every batch,
    Bl = batch // world -- bit-identical to the same columns of the world == 1 result (no collective)."""
    n = pairs_u.numel()
    S = n // batch
    if S < 1:
        raise ValueError("fewer training pairs than one batch")
    if batch % world:
        raise ValueError("the global batch must be a multiple of the number of ranks")
    bl = batch // world
    dev = pairs_u.device
    u = torch.empty((S, bl), dtype=torch.int32, device=dev)
    i = torch.empty((S, bl), dtype=torch.int32, device=dev)
    ud = torch.empty((S, bl * dns), dtype=torch.int32, device=dev)
    j = torch.empty((S, bl * dns), dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().apr_sample_epoch_shard(_ptr(pairs_u, torch.int32), _ptr(pairs_i, torch.int32), n, batch, num_items,
                                                 _ptr(csr_ptr, torch.int64), _ptr(csr_idx, torch.int32) if csr_idx.numel() else 0,
                                                 csr_ptr.numel() - 1, seed & 0xFFFFFFFF, epoch & 0xFFFFFFFF, dns, rank * bl, bl,
                                                 int(fork_workers), _ptr(u), _ptr(i), _ptr(ud), _ptr(j), _ptr(err), _stream()))
    return u, i, ud, j, err


def select_dns(P: torch.Tensor, Q: torch.Tensor, u_dns: torch.Tensor, j_dns: torch.Tensor, dns: int) -> torch.Tensor:
    n_pos = u_dns.numel() // dns
    out = torch.empty(n_pos, dtype=torch.int32, device=P.device)
    _lib.check(_lib.lib().apr_select_dns(_ptr(P, torch.float32), _ptr(Q, torch.float32), P.shape[1], _ptr(u_dns, torch.int32),
                                         _ptr(j_dns, torch.int32), n_pos, dns, _ptr(out), _stream()))
    return out


class TrainWorkspace:
    """Caller-owned scratch of apr_train_steps for up to (n_steps, batch, d)."""

    def __init__(self, n_steps: int, batch: int, d: int, device):
        nbytes = _lib.lib().apr_train_workspace_bytes(n_steps, batch, d)
        if nbytes < 0:
            raise ValueError("bad workspace shape (n_steps=%d, batch=%d, d=%d)" % (n_steps, batch, d))
        self.n_steps, self.batch, self.d, self.nbytes = n_steps, batch, d, nbytes
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(_lib.lib().apr_train_workspace_init(self.buf.data_ptr(), nbytes, _stream()))

    def fits(self, n_steps: int, batch: int, d: int) -> bool:
        return batch == self.batch and d == self.d and n_steps <= self.n_steps

    def unique_counts(self, n_steps: int) -> np.ndarray:
        """[n_steps, 2] {unique users, unique items} of the last prepared chunk (synchronises).
        Raises if an out-of-range id was seen."""
        out = np.zeros((n_steps, 2), dtype=np.int32)
        _lib.check(_lib.lib().apr_train_unique_counts(self.buf.data_ptr(), self.nbytes, n_steps, self.batch, self.d,
                                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _stream()))
        return out


def _train_args(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, adver, mode, ws: TrainWorkspace, stats):
    S, B = u.shape
    d = P.shape[1]
    if i.shape != u.shape or j.shape != u.shape:
        raise ValueError("u, i, j must have the same [n_steps, batch] shape")
    if not ws.fits(S, B, d):
        raise ValueError("workspace was sized for (%d,%d,%d), got (%d,%d,%d)" % (ws.n_steps, ws.batch, ws.d, S, B, d))
    if accP.shape != P.shape or accQ.shape != Q.shape:
        raise ValueError("accumulator shape mismatch")
    return (_ptr(P, torch.float32), _ptr(Q, torch.float32), _ptr(accP, torch.float32), _ptr(accQ, torch.float32),
            P.shape[0], Q.shape[0], d, _ptr(u, torch.int32), _ptr(i, torch.int32), _ptr(j, torch.int32), S, B, float(lr),
            float(reg), float(reg_adv), float(eps), int(adver), int(mode), ws.buf.data_ptr(), ws.nbytes,
            _ptr(stats, torch.float32), _stream())


def train_steps(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, adver, ws: TrainWorkspace, mode: int = 0,
                stats: Optional[torch.Tensor] = None) -> None:
    """training_batch over u.shape[0] batches (utils.py:106-119), in place on P, Q, accP, accQ."""
    _lib.check(_lib.lib().apr_train_steps(*_train_args(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, adver, mode, ws, stats)))
    touch(P)
    touch(Q)


def train_steps_random(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, ws: TrainWorkspace, noise_seed: int,
                       noise_first_step: int, stats: Optional[torch.Tensor] = None) -> None:
    """training_batch with ``--adv random`` (APR.py:170-177): step s uses the noise of global step noise_first_step + s."""
    a = _train_args(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, 1, 0, ws, stats)
    _lib.check(_lib.lib().apr_train_steps_random(*a[:16], noise_seed & 0xFFFFFFFF, noise_first_step & 0xFFFFFFFF, *a[18:]))
    touch(P)
    touch(Q)


def train_status(ws: TrainWorkspace) -> int:
    """Sticky status word of the workspace (synchronises): bit 0 = an out-of-range id reached the index preparation."""
    f = ctypes.c_int32(0)
    _lib.check(_lib.lib().apr_train_status(ws.buf.data_ptr(), ctypes.byref(f), _stream()))
    return int(f.value)


def train_prepare(P, Q, u, i, j, ws: TrainWorkspace) -> None:
    S, B = u.shape
    _lib.check(_lib.lib().apr_train_prepare(_ptr(u, torch.int32), _ptr(i, torch.int32), _ptr(j, torch.int32), S, B, P.shape[1],
                                            P.shape[0], Q.shape[0], ws.buf.data_ptr(), ws.nbytes, _stream()))


def train_run(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, adver, ws: TrainWorkspace, mode: int = 0,
              stats: Optional[torch.Tensor] = None) -> None:
    _lib.check(_lib.lib().apr_train_run(*_train_args(P, Q, accP, accQ, u, i, j, lr, reg, reg_adv, eps, adver, mode, ws, stats)))
    touch(P)
    touch(Q)


def train_layout(n_steps: int, batch: int, d: int) -> dict:
    """Byte offsets of the index arrays inside a TrainWorkspace (for the sharded driver's broadcasts)."""
    out = (ctypes.c_int64 * 13)()
    _lib.check(_lib.lib().apr_train_layout(n_steps, batch, d, out))
    keys = ["total", "Sc", "ucnt", "icnt", "iall", "nslow", "seg_hdr", "rec", "iu_item", "hdr", "npair", "nfast", "pairs"]
    return dict(zip(keys, [int(v) for v in out]))


def train_prepare_range(rows_p: int, rows_q: int, u, i, j, ws: "TrainWorkspace", s0: int, ns: int, clear: bool = True,
                        n_steps: Optional[int] = None) -> None:
    """Index preparation of steps [s0, s0+ns).  ``n_steps`` = the step count the workspace LAYOUT is addressed with
    (default: u.shape[0]); the batches only have to cover the prepared steps."""
    S, B = u.shape
    if n_steps is not None:
        if s0 + ns > S:
            raise ValueError("batches do not cover steps [%d, %d)" % (s0, s0 + ns))
        S = n_steps
    _lib.check(_lib.lib().apr_train_prepare_range(_ptr(u, torch.int32), _ptr(i, torch.int32), _ptr(j, torch.int32), S, B,
                                                  ws.d, rows_p, rows_q, ws.buf.data_ptr(), ws.nbytes, s0, ns, int(clear),
                                                  _stream()))


def _ptr_array(ptrs):
    arr = (ctypes.c_void_p * len(ptrs))()
    for k, p in enumerate(ptrs):
        arr[k] = int(p)
    return arr


def train_stage_sharded(ptrs: dict, nranks: int, rank: int, d: int, n_steps: int, batch: int, lr, reg, reg_adv, eps, adver,
                        ws: "TrainWorkspace", step: int, stage: int, stats: Optional[torch.Tensor] = None) -> None:
    """One stage of one step on row-sharded tables.  ``ptrs`` maps P,Q,accP,accQ,GQ,HQ to lists of nranks shard base
    pointers (ints).  stage 0,1,2 = general path, 3 = fast kernel."""
    a = ptrs.get("_c")
    if a is None:
        a = ptrs["_c"] = [_ptr_array(ptrs[k]) for k in ("P", "Q", "accP", "accQ", "GQ", "HQ")]
    _lib.check(_lib.lib().apr_train_stage_sharded(*a, nranks, rank, d, n_steps, batch, float(lr), float(reg), float(reg_adv),
                                                  float(eps), int(bool(adver)), ws.buf.data_ptr(), ws.nbytes,
                                                  _ptr(stats, torch.float32), step, stage, _stream()))


def train_steps_sharded(ptrs: dict, nranks: int, rank: int, d: int, n_steps: int, batch: int, lr, reg, reg_adv, eps, adver,
                        ws: "TrainWorkspace", first_step: int, count: int, err: torch.Tensor,
                        stats: Optional[torch.Tensor] = None) -> None:
    """Steps [first_step, first_step+count) of the sharded path, all launches and cross-rank barriers issued by the
    library.  ``ptrs`` additionally maps "sig" to the ranks' signal-word pointers."""
    a = ptrs.get("_c7")
    if a is None:
        a = ptrs["_c7"] = [_ptr_array(ptrs[k]) for k in ("P", "Q", "accP", "accQ", "GQ", "HQ", "sig")]
    _lib.check(_lib.lib().apr_train_steps_sharded(*a, nranks, rank, d, n_steps, batch, float(lr), float(reg), float(reg_adv),
                                                  float(eps), int(bool(adver)), ws.buf.data_ptr(), ws.nbytes,
                                                  _ptr(stats, torch.float32), first_step, count, _ptr(err, torch.int32),
                                                  _stream()))


def loss_acc(P, Q, u, i, j) -> torch.Tensor:
    """-> float64 [S,2]: per batch {sum softplus(-r), count(x>0)} (utils.py:159-175)."""
    S, B = u.shape
    out = torch.empty((S, 2), dtype=torch.float64, device=P.device)
    _lib.check(_lib.lib().apr_loss_acc(_ptr(P, torch.float32), _ptr(Q, torch.float32), P.shape[1], _ptr(u, torch.int32),
                                       _ptr(i, torch.int32), _ptr(j, torch.int32), S, B, _ptr(out), _stream()))
    return out


def score_pairs(P, Q, users, items, q_row_offset: int = 0) -> torch.Tensor:
    """``q_row_offset``: Q holds rows [q_row_offset, ...) of the item table (item-sharded callers)."""
    n = users.numel()
    out = torch.empty(n, dtype=torch.float32, device=P.device)
    if n:
        _lib.check(_lib.lib().apr_score_pairs(_ptr(P, torch.float32), _ptr(Q, torch.float32) - q_row_offset * P.shape[1] * 4, P.shape[1],
                                              _ptr(users, torch.int32), _ptr(items, torch.int32), n, _ptr(out), _stream()))
    return out


def score_all_items(P, Q, users, item_lo: int = 0, item_hi: Optional[int] = None) -> torch.Tensor:
    """N4: dense [n_users, item_hi - item_lo] scores u . Q^T in the pinned order (IRGAN.py:36-39 `all_rating`)."""
    item_hi = Q.shape[0] if item_hi is None else item_hi
    n = users.numel()
    out = torch.empty((n, item_hi - item_lo), dtype=torch.float32, device=P.device)
    if n:
        _lib.check(_lib.lib().apr_score_all_items(_ptr(P, torch.float32), _ptr(Q, torch.float32), P.shape[1], _ptr(users, torch.int32),
                                                  n, item_lo, item_hi, _ptr(out), _stream()))
    return out


def eval_candidates(P, Q, users, cand_ptr, cand_idx, want_scores: bool = False):
    n = users.numel()
    pos = torch.empty(n, dtype=torch.int32, device=P.device)
    scores = torch.empty(cand_idx.numel(), dtype=torch.float32, device=P.device) if want_scores else None
    _lib.check(_lib.lib().apr_eval_candidates(_ptr(P, torch.float32), _ptr(Q, torch.float32), P.shape[1],
                                              _ptr(users, torch.int32), _ptr(cand_ptr, torch.int64),
                                              _ptr(cand_idx, torch.int32), n, _ptr(pos), _ptr(scores), _stream()))
    return pos, scores


def eval_fullrank(P, Q, users, test_item, item_lo: int, item_hi: int, excl_ptr, excl_idx, k_top: int = 0,
                  exact: bool = False, position: Optional[torch.Tensor] = None):
    """-> (position[int32 n], topk_ids[n,k] | None, topk_scores[n,k] | None)  (utils.py:210-215,244-261)."""
    n = users.numel()
    d = P.shape[1]
    dev = P.device
    if not exact and tc_supported(d) and n >= 1 and item_hi - item_lo >= 1024:
        # tensor-core route of the same call (apr_eval_fullrank_tc_topk): identical positions / ids
        r = eval_fullrank_tc(P, Q, users, test_item, item_lo, item_hi, excl_ptr, excl_idx, position=position, check=False,
                             k_top=k_top)
        return (r[0], r[1], r[2]) if k_top else (r[0], None, None)
    if position is None:
        position = torch.zeros(n, dtype=torch.int32, device=dev)
    nbytes = _lib.lib().apr_eval_workspace_bytes(n, k_top, d)
    if nbytes < 0:
        raise ValueError("bad eval shape (n_users=%d, k_top=%d, d=%d); k_top must be <= 128" % (n, k_top, d))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    ids = torch.empty((n, k_top), dtype=torch.int32, device=dev) if k_top else None
    sc = torch.empty((n, k_top), dtype=torch.float32, device=dev) if k_top else None
    if excl_idx.numel() == 0:
        excl_idx = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().apr_eval_fullrank(_ptr(P, torch.float32), _ptr(Q, torch.float32), d, _ptr(users, torch.int32),
                                            _ptr(test_item, torch.int32), n, item_lo, item_hi, _ptr(excl_ptr, torch.int64),
                                            _ptr(excl_idx, torch.int32), k_top, _ptr(position, torch.int32), _ptr(ids),
                                            _ptr(sc), int(bool(exact)), ws.data_ptr(), ws.numel(), _stream()))
    return position, ids, sc


def tc_supported(d: int) -> bool:
    return d % 8 == 0 and d <= 256


def eval_tc_timing(enable: bool) -> float:
    """Switch the GEMM-kernel event timing of eval_fullrank_tc on/off; returns the last timed kernel's ms (-1: none)."""
    ms = ctypes.c_float(-1.0)
    _lib.check(_lib.lib().apr_eval_tc_timing(1 if enable else 0, ctypes.byref(ms)))
    return float(ms.value)


EVAL_USER_TILE = 16384   # users per tensor-core evaluation call (larger user sets are tiled by eval_fullrank_tc)

# grow-only evaluation workspace per device: keeping its address stable is what lets the library find the cached
# item-operand image again on the next user tile (apr_eval_fullrank_tc_topk, q_version)
_EVAL_WS = {}
_TABLE_VERSION = {}


def _eval_workspace(nbytes: int, dev) -> torch.Tensor:
    key = str(dev)
    ws = _EVAL_WS.get(key)
    if ws is None or ws.numel() < nbytes + 1024:
        _EVAL_WS.pop(key, None)
        ws = None
        torch.cuda.empty_cache()
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        _EVAL_WS[key] = ws
    return ws


def release_eval_workspace() -> None:
    _EVAL_WS.clear()


def touch(t: torch.Tensor) -> None:
    """Note that the CONTENT of table ``t`` changed through the raw-pointer kernels (torch's own version counter only sees
    torch ops): invalidates the cached evaluation image of that table."""
    _TABLE_VERSION[t.data_ptr()] = _TABLE_VERSION.get(t.data_ptr(), 0) + 1


def table_version(t: torch.Tensor) -> int:
    """Non-zero content tag of a table: torch's in-place version counter combined with the library-side write counter."""
    v = ((int(t._version) + 1) << 24) ^ (_TABLE_VERSION.get(t.data_ptr(), 0) + 1) ^ ((t.data_ptr() >> 8) << 40)
    return (v & ((1 << 63) - 1)) | 1


def eval_fullrank_tc(P, Q, users, test_item, item_lo: int, item_hi: int, excl_ptr, excl_idx,
                     position: Optional[torch.Tensor] = None, check: bool = True, k_top: int = 0, cache_q: bool = False,
                     spos: Optional[torch.Tensor] = None, q_row_offset: int = 0):
    """Positions (and, with ``k_top``, the top-k ids / scores of the non-excluded items) through the tcgen05 bf16x3
    filter + exact re-scoring: same results as eval_fullrank(exact=True).
    -> (position, n_ambiguous) or, with k_top, (position, ids, scores, info).  ``check`` synchronises to verify the
    pipeline flag and the list capacity.  ``cache_q`` keeps the item-operand image of Q in a persistent workspace
    across calls (user tiles) until the table changes (``touch`` / any torch in-place op on it).
    Item-sharded callers hold only the rows [q_row_offset, q_row_offset + Q.shape[0]) of the item table in ``Q`` and
    pass ``spos`` (score of each user's held-out item, from the rank that owns that row); then ``test_item`` may be None."""
    n = users.numel()
    d = P.shape[1]
    dev = P.device
    if position is None:
        position = torch.zeros(n, dtype=torch.int32, device=dev)
    if n > EVAL_USER_TILE:
        # one library call takes a tile of the users (its per-CTA counter block and its (user, item) list are sized per
        # call); the item-operand image is built once and reused by every tile
        outs = []
        for a0 in range(0, n, EVAL_USER_TILE):
            sl = slice(a0, min(n, a0 + EVAL_USER_TILE))
            sub_ptr = (excl_ptr[sl.start:sl.stop + 1] - excl_ptr[sl.start]).contiguous()
            e0, e1 = int(excl_ptr[sl.start].item()), int(excl_ptr[sl.stop].item())
            outs.append(eval_fullrank_tc(P, Q, users[sl].contiguous(), None if test_item is None else test_item[sl].contiguous(),
                                         item_lo, item_hi, sub_ptr, excl_idx[e0:e1].contiguous(), position=position[sl],
                                         check=check, k_top=k_top, cache_q=True,
                                         spos=None if spos is None else spos[sl].contiguous(), q_row_offset=q_row_offset))
        if k_top:
            info = {k: sum(o[3][k] for o in outs) for k in outs[0][3]} if check else {"ambiguous": -1}
            return position, torch.cat([o[1] for o in outs]), torch.cat([o[2] for o in outs]), info
        return position, (sum(o[1] for o in outs) if check else -1)
    n_items = item_hi - item_lo
    if q_row_offset and (item_lo < q_row_offset or item_hi > q_row_offset + Q.shape[0]):
        raise ValueError("item range [%d, %d) is outside the local shard of Q" % (item_lo, item_hi))
    nbytes = _lib.lib().apr_eval_tc_topk_workspace_bytes(n, n_items, d, k_top)
    if nbytes < 0:
        raise ValueError("tensor-core evaluation needs d % 8 == 0, d <= 256 and k_top <= 128")
    ws = _eval_workspace(nbytes, dev) if cache_q else torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    if excl_idx.numel() == 0:
        excl_idx = torch.zeros(1, dtype=torch.int32, device=dev)
    ids = torch.empty((n, k_top), dtype=torch.int32, device=dev) if k_top else None
    sc = torch.empty((n, k_top), dtype=torch.float32, device=dev) if k_top else None
    _lib.check(_lib.lib().apr_eval_fullrank_tc_topk(
        _ptr(P, torch.float32), _ptr(Q, torch.float32) - q_row_offset * d * 4, d, _ptr(users, torch.int32),
        _ptr(test_item, torch.int32), n, item_lo, item_hi, _ptr(excl_ptr, torch.int64), _ptr(excl_idx, torch.int32), k_top,
        _ptr(position, torch.int32), _ptr(ids), _ptr(sc), table_version(Q) if cache_q else 0, _ptr(spos, torch.float32),
        ws.data_ptr() + off, nbytes, _ptr(err), _stream()))
    info = {"ambiguous": -1}
    if check:
        c = (ctypes.c_int32 * 4)()
        _lib.check(_lib.lib().apr_eval_tc_ambiguous(ws.data_ptr() + off, n, n_items, d, k_top, c, _stream()))
        info = {"ambiguous": int(c[0]), "capacity": int(c[1]), "topk_candidates": int(c[2]), "exact_fallback_users": int(c[3])}
        if int(err.item()) != 0:
            raise RuntimeError("tcgen05 evaluation pipeline timed out (error flag set)")
        if int(c[1]) == 0:
            raise RuntimeError("ambiguous list overflow (%d pairs): use the exact path" % int(c[0]))
    if k_top:
        return position, ids, sc, info
    return position, info["ambiguous"]


def topk_merge(ids: torch.Tensor, scores: torch.Tensor, k: int):
    """K10: [n_users, m] per-shard lists (id < 0 = padding) -> the k best by (score desc, id asc), on the GPU."""
    n, m = ids.shape
    out_i = torch.empty((n, k), dtype=torch.int32, device=ids.device)
    out_s = torch.empty((n, k), dtype=torch.float32, device=ids.device)
    if n:
        _lib.check(_lib.lib().apr_topk_merge(_ptr(ids, torch.int32), _ptr(scores, torch.float32), n, m, k, _ptr(out_i),
                                             _ptr(out_s), _stream()))
    return out_i, out_s


def sum_squares(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().apr_sum_squares(_ptr(x, torch.float32), x.numel(), _ptr(out), _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------
# N1: device-side data loader (csrc/loader.cu)
# ---------------------------------------------------------------------------------------------------------
def _file_to_device(filename: str, device) -> torch.Tensor:
    """One read() of the file into pinned memory, one H2D copy."""
    size = __import__("os").path.getsize(filename)
    host = torch.empty(max(size, 1), dtype=torch.uint8).pin_memory()
    with open(filename, "rb") as f:
        f.readinto(host.numpy())
    return host[:size].to(device, non_blocking=True) if size else torch.zeros(0, dtype=torch.uint8, device=device)


def _loader_ws(n_bytes: int, max_lines: int, device) -> torch.Tensor:
    return torch.empty(_lib.lib().apr_loader_workspace_bytes(n_bytes, max_lines), dtype=torch.uint8, device=device)


def parse_rating_text(text: torch.Tensor):
    """uint8 device tensor holding a `*.rating` TSV -> (uid int32[n], iid int32[n], rating float32[n]) on the device."""
    dev, nb = text.device, text.numel()
    if nb == 0:
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return z, z.clone(), torch.zeros(0, dtype=torch.float32, device=dev)
    ws = _loader_ws(nb, 0, dev)
    n = ctypes.c_int64(0)
    _lib.check(_lib.lib().apr_tsv_count_lines(text.data_ptr(), nb, ws.data_ptr(), ws.numel(), ctypes.byref(n), _stream()))
    n = int(n.value)
    u = torch.empty(n, dtype=torch.int32, device=dev)
    i = torch.empty(n, dtype=torch.int32, device=dev)
    r = torch.empty(n, dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    if n:
        _lib.check(_lib.lib().apr_tsv_parse(text.data_ptr(), nb, ws.data_ptr(), n, 0, _ptr(u), _ptr(i), _ptr(r), 0, 0, 0, _ptr(err),
                                            _stream()))
        if int(err.item()):
            raise ValueError("malformed rating file (uid / iid must be non-negative integers, rating a decimal number)")
    return u, i, r


def parse_negative_text(text: torch.Tensor):
    """He-format `.test.negative` text -> CSR (ptr int64[n+1], idx int32) of the ids after the first tab of every line."""
    dev, nb = text.device, text.numel()
    ws = _loader_ws(nb, 0, dev)
    n = ctypes.c_int64(0)
    _lib.check(_lib.lib().apr_tsv_count_lines(text.data_ptr(), nb, ws.data_ptr(), ws.numel(), ctypes.byref(n), _stream()))
    n = int(n.value)
    cnt = torch.zeros(n, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().apr_tsv_parse(text.data_ptr(), nb, ws.data_ptr(), n, 1, 0, 0, 0, _ptr(cnt), 0, 0, _ptr(err), _stream()))
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(cnt, 0)
    idx = torch.empty(int(ptr[-1].item()), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().apr_tsv_parse(text.data_ptr(), nb, ws.data_ptr(), n, 2, 0, 0, 0, 0, _ptr(ptr), _ptr(idx) if idx.numel() else 0,
                                        _ptr(err), _stream()))
    if int(err.item()):
        raise ValueError("malformed negatives file")
    return ptr, idx


def train_rows(u: torch.Tensor, quirk: bool = True):
    """trainList row of every train line (Dataset.py:306-325) -> (row int32[n], file_is_uid_sorted)."""
    n = u.numel()
    row = torch.empty(n, dtype=torch.int32, device=u.device)
    flag = torch.zeros(1, dtype=torch.int32, device=u.device)
    if n:
        ws = _loader_ws(0, n, u.device)
        _lib.check(_lib.lib().apr_loader_train_rows(_ptr(u, torch.int32), n, int(bool(quirk)), _ptr(row), _ptr(flag), ws.data_ptr(),
                                                    ws.numel(), _stream()))
    return row, int(flag.item()) == 0


def build_csr_device(row: torch.Tensor, item: torch.Tensor, n_rows: int):
    """Sorted, de-duplicated CSR of (row, item) on the device -> (ptr int64[n_rows+1], idx int32[nnz])."""
    n = row.numel()
    dev = row.device
    ptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    if n == 0:
        return ptr, torch.zeros(0, dtype=torch.int32, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    nu = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _loader_ws(0, n, dev)
    _lib.check(_lib.lib().apr_loader_csr(_ptr(row, torch.int32), _ptr(item, torch.int32), n, n_rows, _ptr(ptr), _ptr(idx), _ptr(nu),
                                         ws.data_ptr(), ws.numel(), _stream()))
    return ptr, idx[:int(nu.item())].clone()


def unique_pairs_device(u: torch.Tensor, i: torch.Tensor, rating: torch.Tensor):
    """(u, i) of the lines with rating > 0, duplicates collapsed onto the first occurrence, in file order (APR.py:30-36)."""
    n = u.numel()
    dev = u.device
    if n == 0:
        return u.clone(), i.clone()
    ou = torch.empty(n, dtype=torch.int32, device=dev)
    oi = torch.empty(n, dtype=torch.int32, device=dev)
    m = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _loader_ws(0, n, dev)
    _lib.check(_lib.lib().apr_loader_unique_pairs(_ptr(u, torch.int32), _ptr(i, torch.int32), _ptr(rating, torch.float32), n, _ptr(ou),
                                                  _ptr(oi), _ptr(m), ws.data_ptr(), ws.numel(), _stream()))
    k = int(m.item())
    return ou[:k].clone(), oi[:k].clone()


# ---------------------------------------------------------------------------------------------------------
# N3: Keras MatrixFactorization / BPR Recommender models (csrc/keras_mf.cu)
# ---------------------------------------------------------------------------------------------------------
def keras_step(P, Q, mP, vP, mQ, vQ, gP, gQ, u, i, j=None, y=None, lr=0.001, beta1=0.9, beta2=0.999, t=1,
               loss_sum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One ``model.fit`` batch of MF.py (``y``: labels) or BPR.py (``j``: negative items) with Keras' dense Adam.
    Returns the device double that accumulates the summed loss."""
    if loss_sum is None:
        loss_sum = torch.zeros(1, dtype=torch.float64, device=P.device)
    kind = 0 if y is not None else 1
    _lib.check(_lib.lib().apr_keras_step(_ptr(P, torch.float32), _ptr(Q, torch.float32), _ptr(mP, torch.float32),
                                         _ptr(vP, torch.float32), _ptr(mQ, torch.float32), _ptr(vQ, torch.float32),
                                         _ptr(gP, torch.float32), _ptr(gQ, torch.float32), P.shape[0], Q.shape[0], P.shape[1],
                                         _ptr(u, torch.int32), _ptr(i, torch.int32), _ptr(j, torch.int32), _ptr(y, torch.float32),
                                         u.numel(), kind, float(lr), float(beta1), float(beta2), int(t), _ptr(loss_sum), _stream()))
    touch(P)
    touch(Q)
    return loss_sum
