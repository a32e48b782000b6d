"""Step driver, loss/accuracy, evaluation and logging helpers with the reference's names and argument meaning
(utils.py:18-32, 81-277), running on the CUDA library.  ``sess`` is an ``apr_b200.APR.Session``: an opaque engine
handle that replaces ``tf.Session`` (it owns the stream-ordered workspaces; tables live on the model).
"""
from __future__ import annotations

import logging
import math
import os
import random
from time import localtime, strftime, time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine


# ---------------------------------------------------------------------------------------------------------
# logging / files (utils.py:18-32, 270-277) -- byte-compatible formats
# ---------------------------------------------------------------------------------------------------------
def write2file(path, name, output):
    print(output)
    if not os.path.exists(path):
        os.makedirs(path)
    with open(path + name, 'a') as thefile:
        thefile.write("%s\n" % output)


def prediction2file(path, name, pred):
    if not os.path.exists(path):
        os.makedirs(path)
    with open(path + name, 'w') as thefile:
        for item in pred:
            thefile.write("%f\n" % item)


def set_seed(seed, cuda=False):
    np.random.seed(seed)
    random.seed(seed)
    if cuda:
        torch.cuda.manual_seed(seed)
    else:
        torch.manual_seed(seed)


def init_logging(args, time_stamp):
    path = "Log/%s_%s/" % (strftime('%Y-%m-%d_%H', localtime()), args.task)
    if not os.path.exists(path):
        os.makedirs(path)
    logging.basicConfig(filename=path + "%s_log_embed_size%d_%s" % (args.dataset, args.embed_size, time_stamp),
                        level=logging.INFO)
    logging.info(args)
    print(args)


# ---------------------------------------------------------------------------------------------------------
# batches
# ---------------------------------------------------------------------------------------------------------
class DeviceBatches(object):
    """List-like view of an int32 device tensor [num_batch, B]: ``len()`` = num_batch, ``[s]`` = the [B,1] batch the
    reference's lists hold (APR.py:80-81).  Produced by the GPU sampler; accepted by every driver function."""

    def __init__(self, t: torch.Tensor):
        assert t.dim() == 2 and t.dtype == torch.int32
        self.t = t

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, s):
        return self.t[s].unsqueeze(1)

    def __iter__(self):
        for s in range(len(self)):
            yield self[s]

    def numpy(self) -> np.ndarray:
        return self.t.cpu().numpy()


def as_device_batches(x, device) -> torch.Tensor:
    """DeviceBatches | list of [B,1] arrays | [S,B] array  ->  contiguous int32 CUDA tensor [S,B].
    Host inputs go through pinned memory (this is the feed_dict copy of the reference's sess.run)."""
    if isinstance(x, DeviceBatches):
        return x.t
    if isinstance(x, torch.Tensor):
        t = x.reshape(x.shape[0], -1) if x.dim() != 2 else x
        return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    if isinstance(x, (list, tuple)):
        if len(x) and isinstance(x[0], torch.Tensor):
            return torch.stack([b.reshape(-1) for b in x]).to(device=device, dtype=torch.int32).contiguous()
        arr = np.stack([np.asarray(b).reshape(-1) for b in x]).astype(np.int32, copy=False)
    else:
        arr = np.asarray(x)
        arr = arr.reshape(arr.shape[0], -1).astype(np.int32, copy=False)
    host = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return host.to(device, non_blocking=True)


# ---------------------------------------------------------------------------------------------------------
# the step driver (utils.py:106-140)
# ---------------------------------------------------------------------------------------------------------
def training_batch(model, sess, batches, adver=False):
    """For every batch: [if adver: update_P/update_Q] then the optimizer (dns == 1); with dns > 1 the best-scored of
    the dns sampled negatives is chosen per positive and there is no adversarial update (utils.py:121-139) -- but the
    optimizer is still the one of the graph the model was built with: on an adversarial model (args.adver) it minimises
    L + reg_adv L_adv + 2 reg-terms with Delta == 0, i.e. a plain step scaled by (1 + reg_adv) (engine adver mode 3).
    Ids outside the tables raise at the end of the call (the reference's embedding_lookup raises InvalidArgument)."""
    user_input, item_input_pos, user_dns_list, item_dns_list = batches
    dev = model.device
    U = as_device_batches(user_input, dev)
    I = as_device_batches(item_input_pos, dev)
    if model.dns == 1:
        J = as_device_batches(item_dns_list, dev)
        sess.train_steps(model, U, I, J, adver=bool(adver), check=True)
        item_input_neg = item_dns_list
    else:
        UD = as_device_batches(user_dns_list, dev)
        JD = as_device_batches(item_dns_list, dev)
        S, B = U.shape
        J = torch.empty((S, B), dtype=torch.int32, device=dev)
        for s in range(S):
            # the choice must see the parameters as updated by the previous batches
            J[s] = engine.select_dns(model.embedding_P, model.embedding_Q, UD[s].contiguous(), JD[s].contiguous(), model.dns)
            sess.train_steps(model, U[s:s + 1], I[s:s + 1], J[s:s + 1], adver=3 if model.adver else 0)
        sess.check(model)
        item_input_neg = DeviceBatches(J)
    return user_input, item_input_pos, item_input_neg


def training_loss_acc(model, sess, train_batches, output_adv):
    """utils.py:159-175: (sum_b loss_b / num_batch, mean_b mean(x > 0))."""
    if output_adv:
        raise NotImplementedError("output_adv=1 needs persistent Delta tables; the reference drivers always pass 0")
    user_input, item_input_pos, item_input_neg = train_batches
    dev = model.device
    U = as_device_batches(user_input, dev)
    I = as_device_batches(item_input_pos, dev)
    J = as_device_batches(item_input_neg, dev)
    out = engine.loss_acc(model.embedding_P, model.embedding_Q, U, I, J).cpu().numpy()
    num_batch, B = U.shape
    return float(out[:, 0].sum() / num_batch), float((out[:, 1] / B).sum() / num_batch)


# ---------------------------------------------------------------------------------------------------------
# evaluation (utils.py:178-267)
# ---------------------------------------------------------------------------------------------------------
class EvalInputs(object):
    """What ``init_eval_model`` returns instead of per-user numpy feed dicts: device-resident candidate structure.

    eval_mode == "all":    excl CSR = sorted(trainList[u] united with {test item}), users, test items
    eval_mode == "sample": explicit candidate CSR (100 sampled negatives + the test item last)
    ``feed[user]`` still materialises the reference's (user_input[n,1], item_input[n,1]) pair on demand."""

    def __init__(self, mode, users, test_items, num_items, excl_ptr=None, excl_idx=None, cand_ptr=None, cand_idx=None,
                 device=None):
        self.mode = mode
        self.num_items = num_items
        self.users_h = np.asarray(users, dtype=np.int32)
        self.test_h = np.asarray(test_items, dtype=np.int32)
        self.excl_ptr_h, self.excl_idx_h = excl_ptr, excl_idx
        self.cand_ptr_h, self.cand_idx_h = cand_ptr, cand_idx
        self.device = device
        self._dev = None

    def __len__(self):
        return self.users_h.shape[0]

    def n_neg(self) -> np.ndarray:
        if self.mode == "all":
            # negatives = range(num_items) minus (trainList[u] united with {test}); ids >= num_items never were candidates
            row_of = np.repeat(np.arange(len(self), dtype=np.int64), np.diff(self.excl_ptr_h))
            inrange = np.bincount(row_of[self.excl_idx_h < self.num_items], minlength=len(self))
            return self.num_items - inrange
        return np.diff(self.cand_ptr_h) - 1

    def to_device(self, device):
        if self._dev is None or self.device != device:
            self.device = device
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
            d = {"users": t(self.users_h, torch.int32), "test": t(self.test_h, torch.int32)}
            if self.mode == "all":
                d["excl_ptr"] = t(self.excl_ptr_h, torch.int64)
                d["excl_idx"] = t(self.excl_idx_h, torch.int32)
            else:
                d["cand_ptr"] = t(self.cand_ptr_h, torch.int64)
                d["cand_idx"] = t(self.cand_idx_h, torch.int32)
            self._dev = d
        return self._dev

    def __getitem__(self, k):
        user = int(self.users_h[k])
        if self.mode == "all":
            ex = self.excl_idx_h[self.excl_ptr_h[k]:self.excl_ptr_h[k + 1]]
            items = np.setdiff1d(np.arange(self.num_items, dtype=np.int64), ex).tolist() + [int(self.test_h[k])]
        else:
            items = self.cand_idx_h[self.cand_ptr_h[k]:self.cand_ptr_h[k + 1]].tolist()
        return np.full(len(items), user, dtype='int32')[:, None], np.array(items)[:, None]


_PY_CHOICE_CACHE = {}


def _choice_index_stream(n: int, count: int) -> np.ndarray:
    """The index sequence ``random.seed(2019); random.choice(seq)`` walks for a sequence of length n
    (utils.py:203-205): identical for every user because the seed is reset per user."""
    key = n
    have = _PY_CHOICE_CACHE.get(key)
    if have is None or have.shape[0] < count:
        rng = random.Random(2019)
        k = n.bit_length()
        out = []
        total = max(count, 1024)
        while len(out) < total:
            r = rng.getrandbits(k)
            while r >= n:
                r = rng.getrandbits(k)
            out.append(r)
        have = np.asarray(out, dtype=np.int64)
        _PY_CHOICE_CACHE[key] = have
    return have


def init_eval_model(dataset, args):
    """utils.py:178-218.  eval_mode "all": candidates = range(num_items) - trainList[u] - {test} + [test];
    "sample": 100 x random.choice(train iid column) with random.seed(2019) per user, rejecting train/test items."""
    num_users = dataset.num_users
    mode = getattr(args, "eval_mode", "all")
    if mode == "all" and hasattr(dataset, "device_eval_exclusion"):
        # DeviceDataset (N1): the exclusion CSR is built on the GPU (sort + unique), no per-user Python
        ptr, idx = dataset.device_eval_exclusion()
        return EvalInputs("all", np.arange(num_users, dtype=np.int32), dataset.device_test_items().cpu().numpy(),
                          dataset.num_items, excl_ptr=ptr.cpu().numpy(), excl_idx=idx.cpu().numpy())
    test = np.asarray([dataset.testRatings[u][1] for u in range(num_users)], dtype=np.int32)
    users = np.arange(num_users, dtype=np.int32)
    train_list = dataset.trainList
    if mode == "sample":
        iid = np.asarray(dataset.iid_column)
        need = 400
        stream = iid[_choice_index_stream(len(iid), need)]
        ptr = np.zeros(num_users + 1, dtype=np.int64)
        rows = []
        for u in range(num_users):
            tr = np.asarray(train_list[u] if u < len(train_list) else [], dtype=np.int64)
            while True:
                ok = ~np.isin(stream[:need], tr) & (stream[:need] != test[u])
                if ok.sum() >= 100:
                    break
                need *= 2
                stream = iid[_choice_index_stream(len(iid), need)]
            sel = stream[:need][ok][:100]
            rows.append(np.concatenate([sel, [test[u]]]).astype(np.int32))
            ptr[u + 1] = ptr[u] + 101
        return EvalInputs("sample", users, test, dataset.num_items, cand_ptr=ptr, cand_idx=np.concatenate(rows))
    from .Dataset import build_sorted_csr
    lists = [(list(train_list[u]) if u < len(train_list) else []) + [int(test[u])] for u in range(num_users)]
    ptr, idx = build_sorted_csr(lists)
    return EvalInputs("all", users, test, dataset.num_items, excl_ptr=ptr, excl_idx=idx)


def metrics_from_position(position: np.ndarray, n_neg: np.ndarray, K: int) -> np.ndarray:
    """utils.py:253-261 vectorised: res[user, (hr, ndcg, auc), k-1]."""
    position = np.asarray(position, dtype=np.int64)
    ks = np.arange(1, K + 1)[None, :]
    hit = position[:, None] < ks
    nd = math.log(2) / np.log(position.astype(np.float64) + 2.0)
    res = np.empty((position.shape[0], 3, K), dtype=np.float64)
    res[:, 0, :] = hit
    res[:, 1, :] = np.where(hit, nd[:, None], 0.0)
    res[:, 2, :] = (1.0 - position / np.asarray(n_neg, dtype=np.float64))[:, None]
    return res


def eval_positions(model, feed: EvalInputs, exact: bool = False) -> torch.Tensor:
    """Device int32 rank position (#negatives scoring >= the held-out item) of every evaluation user."""
    d = feed.to_device(model.device)
    if feed.mode == "all":
        if not exact and engine.tc_supported(model.embedding_P.shape[1]) and feed.num_items >= 1024:
            # tcgen05 bf16x3 filter + exact re-scoring: same positions as the fp32 kernel (tests/test_gpu_eval_tc.py)
            try:
                pos, _ = engine.eval_fullrank_tc(model.embedding_P, model.embedding_Q, d["users"], d["test"], 0,
                                                 feed.num_items, d["excl_ptr"], d["excl_idx"])
                return pos
            except RuntimeError as e:  # ambiguous-list overflow (degenerate score distribution): exact kernel
                if "overflow" not in str(e):
                    raise
        pos, _, _ = engine.eval_fullrank(model.embedding_P, model.embedding_Q, d["users"], d["test"], 0, feed.num_items,
                                         d["excl_ptr"], d["excl_idx"], 0, exact=True)
        return pos
    pos, _ = engine.eval_candidates(model.embedding_P, model.embedding_Q, d["users"], d["cand_ptr"], d["cand_idx"])
    return pos


def evaluate(model, sess, dataset, feed_dicts, output_adv, args):
    """utils.py:221-241 -> ((hr[K], ndcg[K], auc[K]), res[U,3,K]); K = 100 if eval_mode == "all" else 10.
    Scores use the unperturbed P, Q (output_adv is always 0 in the drivers, APR.py:268)."""
    if output_adv:
        raise NotImplementedError("output_adv=1 needs persistent Delta tables; the reference drivers always pass 0")
    K = 100 if getattr(args, "eval_mode", "all") == "all" else 10
    pos = eval_positions(model, feed_dicts).cpu().numpy()
    res = metrics_from_position(pos, feed_dicts.n_neg(), K)
    hr, ndcg, auc = (res.mean(axis=0)).tolist()
    return (hr, ndcg, auc), res


def output_evaluate(model, sess, dataset, train_batches, eval_feed_dicts, epoch_count, batch_time, train_time, prev_acc,
                    runName, args, output_adv):
    """utils.py:81-101: post-epoch loss/acc pass, evaluation, Frobenius norms, one `.out` line."""
    train_loss, post_acc = training_loss_acc(model, sess, train_batches, output_adv)
    eval_begin = time()
    result, raw_result = evaluate(model, sess, dataset, eval_feed_dicts, output_adv, args)
    eval_time = time() - eval_begin
    norm_p = math.sqrt(float(engine.sum_squares(model.embedding_P).item()))
    norm_q = math.sqrt(float(engine.sum_squares(model.embedding_Q).item()))
    hr, ndcg, auc = np.swapaxes(result, 0, 1)[-1]
    res = "Epoch %d [%.1fs + %.1fs]: HR = %.4f, NDCG = %.4f ACC = %.4f ACC_adv = %.4f [%.1fs], |P|=%.2f, |Q|=%.2f" % \
          (epoch_count, batch_time, train_time, hr, ndcg, prev_acc, post_acc, eval_time, norm_p, norm_q)
    write2file(args.path + "out/" + args.opath, runName + ".out", res)
    return post_acc, ndcg, result, raw_result
