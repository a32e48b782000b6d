"""The 1-based variant of the path (evaluation_adv.py:32-486), driven by run_adv.py.

Differences from APR.py + utils.py that are kept:
  * tables have exactly num_users / num_items rows (evaluation_adv.py:119-129);
  * ``dataset.testRatings`` maps user -> held-out item (evaluation_adv.py:87,427);
  * evaluation covers users 1 .. num_users-1 and candidates never contain item 0 (evaluation_adv.py:414,428-432);
  * K is fixed at 100 (evaluation_adv.py:450); ``init_eval_model(model, dataset)``; checkpoint paths are prefixed with
    ``args.path`` (evaluation_adv.py:222-228).
The arithmetic is the same CUDA path as apr_b200.APR.
"""
import numpy as np

from . import APR as _apr
from .APR import Session, sampling, shuffle  # noqa: F401  (same surface as the reference module)
from .Dataset import build_sorted_csr
from .utils import (EvalInputs, eval_positions, metrics_from_position, prediction2file, training_batch,  # noqa: F401
                    training_loss_acc, write2file)


class MF(_apr.MF):
    extra_row = 0


def _test_item(dataset, user):
    r = dataset.testRatings[user]
    return int(r[1]) if isinstance(r, (list, tuple, np.ndarray)) else int(r)


def init_eval_model(model, dataset):
    """evaluation_adv.py:406-437: users 1..U-1; candidates = range(1, num_items) - trainList[u] - {test} + [test]."""
    users = np.arange(1, dataset.num_users, dtype=np.int32)
    test = np.asarray([_test_item(dataset, int(u)) for u in users], dtype=np.int32)
    tl = dataset.trainList
    # item 0 is always excluded: model it as a member of every user's exclusion set
    lists = [(list(tl[u]) if u < len(tl) else []) + [int(t), 0] for u, t in zip(users.tolist(), test.tolist())]
    ptr, idx = build_sorted_csr(lists)
    return EvalInputs("all", users, test, dataset.num_items, excl_ptr=ptr, excl_idx=idx)


def evaluate(model, sess, dataset, feed_dicts, output_adv):
    """evaluation_adv.py:440-461 -> ((hr[100], ndcg[100], auc[100]), res[U-1, 3, 100])."""
    if output_adv:
        raise NotImplementedError("output_adv=1 needs persistent Delta tables; the reference drivers always pass 0")
    pos = eval_positions(model, feed_dicts).cpu().numpy()
    res = metrics_from_position(pos, feed_dicts.n_neg(), 100)
    hr, ndcg, auc = (res.mean(axis=0)).tolist()
    return (hr, ndcg, auc), res


class _Args(object):
    """args proxy: evaluation_adv always evaluates every item."""

    def __init__(self, args):
        self.__dict__["_a"] = args

    def __getattr__(self, k):
        if k == "eval_mode":
            return "all"
        return getattr(self._a, k)

    def __setattr__(self, k, v):
        setattr(self._a, k, v)


def training(model, dataset, args, runName, epoch_start, epoch_end, time_stamp):
    """evaluation_adv.py:218-306: same loop as APR.training with this module's evaluator and path prefix."""
    saved = (_apr.init_eval_model, _apr.output_evaluate)

    def _init(ds, a):
        return init_eval_model(model, ds)

    try:
        _apr.init_eval_model = _init
        return _apr.training(model, dataset, _Args(args), runName, epoch_start, epoch_end, time_stamp,
                             ckpt_prefix=args.path)
    finally:
        _apr.init_eval_model, _apr.output_evaluate = saved
