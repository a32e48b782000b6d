"""BPR-MF / APR model and training loop with the reference's surface (APR.py:30-292), on hand-written sm_100a CUDA.

    MF(num_users, num_items, args)   reads args.{embed_size, lr, reg, dns, adv, eps, adver, reg_adv, epochs}
    MF.build_graph()                 allocates the device tables (instead of building a TF graph)
    sampling(dataset)                -> (user_input, item_input_pos)
    shuffle(samples, batch_size, dataset, model) -> (user_list, item_pos_list, user_dns_list, item_dns_list)
    training(model, dataset, args, runName, epoch_start, epoch_end, time_stamp)

Differences that are deliberate: there are no delta_P/delta_Q tables (Delta lives in the step
workspace); the sampler is a seeded counter-based GPU sampler (the reference is unseeded and its forked workers
share one RNG state); checkpoints are .npz files in the reference's Pretrain/ tree.
"""
from __future__ import annotations

import logging
import os
from time import time
from typing import Optional

import numpy as np
import torch

from . import engine
from .utils import (DeviceBatches, as_device_batches, init_eval_model, output_evaluate, prediction2file, training_batch,
                    training_loss_acc, write2file)

SEED = 2019          # the reference's only seed constant (utils.py:203, Dataset.py:40,88)
ADAGRAD_INIT = 0.1   # tf.train.AdagradOptimizer initial_accumulator_value (APR.py:195)
MAX_CHUNK_TRIPLES = 1 << 22


class Session(object):
    """Opaque engine handle standing where ``tf.Session`` stood: owns the caller-provided workspaces of the C ABI."""

    def __init__(self, mode: Optional[int] = None):
        self.device = engine.require_cuda()
        # -1 = automatic: the one-cluster persistent kernel (mode 2) for small batches whose steps are a few microseconds
        # long (the reference's default B = 512: 14 us/step against 23 with per-phase launches), per-phase launches
        # with the fast kernel on a second stream (mode 0) otherwise
        self.mode = int(os.environ.get("APR_B200_STEP_MODE", "-1")) if mode is None else mode
        self.cluster_max_batch = int(os.environ.get("APR_B200_CLUSTER_MAX_BATCH", "768"))
        self._ws = None
        self._host_pipe = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._ws = None
        return False

    def workspace(self, n_steps: int, batch: int, d: int) -> engine.TrainWorkspace:
        if self._ws is None or not self._ws.fits(n_steps, batch, d):
            self._ws = engine.TrainWorkspace(n_steps, batch, d, self.device)
        return self._ws

    def step_mode(self, batch: int) -> int:
        if self.mode >= 0:
            return self.mode
        return 2 if batch <= self.cluster_max_batch else 0

    def train_steps(self, model, U, I, J, adver, stats=None, check: bool = False) -> None:
        """S steps over the batches U/I/J [S, B].  Device tensors run as they are; HOST batches (CPU tensors or ndarrays,
        pinned or not -- the reference's feed_dict, utils.py:117-119) are streamed: chunk k+1 travels host->device on a
        copy stream while chunk k trains, so the copies hide behind the step kernels."""
        if not (isinstance(U, torch.Tensor) and U.is_cuda):
            self._train_steps_host(model, U, I, J, adver, stats)
            if check:
                self.check(model)
            return
        S, B = U.shape
        chunk = max(1, min(S, MAX_CHUNK_TRIPLES // max(B, 1)))
        ws = self.workspace(chunk, B, model.embedding_size)
        for s0 in range(0, S, chunk):
            s1 = min(S, s0 + chunk)
            self._steps(model, U[s0:s1], I[s0:s1], J[s0:s1], adver, ws, None if stats is None else stats[s0:s1])
        if check:
            self.check(model)

    def _steps(self, model, U, I, J, adver, ws, stats) -> None:
        """One library call.  ``adver``: False/0 = BPR, True/1 = the model's adversarial mode (--adv grad | random),
        3 = optimizer of the adversarial loss with Delta == 0 (dns > 1 branch, utils.py:121-139)."""
        a = (model.embedding_P, model.embedding_Q, model.acc_P, model.acc_Q, U, I, J, model.learning_rate, model.reg,
             model.reg_adv, model.eps)
        if int(adver) == 1 and getattr(model, "adv", "grad") == "random":
            # APR.py:170-177: update_P/update_Q draw a fresh Delta for every batch
            engine.train_steps_random(*a, ws, model.seed, model.adv_step, stats=stats)
            model.adv_step += U.shape[0]
        else:
            engine.train_steps(*a, int(adver), ws, mode=self.step_mode(U.shape[1]), stats=stats)

    def check(self, model=None) -> None:
        """Raise if an id outside its table reached the step (the reference's embedding_lookup raises InvalidArgument);
        synchronises."""
        if self._ws is not None and engine.train_status(self._ws) & 1:
            raise IndexError("a user / item id outside the embedding tables was fed to training_batch "
                             "(1-based ids, or a model built without the extra row?)")

    def _train_steps_host(self, model, U, I, J, adver: bool, stats=None) -> None:
        host = []
        for x in (U, I, J):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
            t = t.reshape(t.shape[0], -1)
            host.append(t if t.dtype == torch.int32 else t.to(torch.int32))
        S, B = host[0].shape
        chunk = max(1, min(S, MAX_CHUNK_TRIPLES // max(B, 1)))
        ws = self.workspace(chunk, B, model.embedding_size)
        st = self._host_pipe
        if st is None or st["dev"][0].shape != (3, chunk, B):
            st = {"dev": [torch.empty((3, chunk, B), dtype=torch.int32, device=self.device) for _ in range(2)],
                  "pin": [None, None], "copy": torch.cuda.Stream(device=self.device),
                  "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)]}
            self._host_pipe = st
        cur = torch.cuda.current_stream(self.device)
        pinned = all(t.is_pinned() and t.is_contiguous() for t in host)
        for k, s0 in enumerate(range(0, S, chunk)):
            s1 = min(S, s0 + chunk)
            n, b = s1 - s0, k & 1
            dev = st["dev"][b]
            if k >= 2:
                st["copy"].wait_event(st["free"][b])     # the steps that read this buffer two chunks ago are done
            if not pinned:                                 # pageable input: stage through a pinned buffer of our own
                if st["pin"][b] is None or st["pin"][b].shape != (3, chunk, B):
                    st["pin"][b] = torch.empty((3, chunk, B), dtype=torch.int32).pin_memory()
                elif k >= 2:
                    st["ready"][b].synchronize()           # its previous host->device copy has left the staging buffer
                for q in range(3):
                    st["pin"][b][q, :n].copy_(host[q][s0:s1])
            with torch.cuda.stream(st["copy"]):
                for q in range(3):
                    src = host[q][s0:s1] if pinned else st["pin"][b][q, :n]
                    dev[q, :n].copy_(src, non_blocking=True)
                st["ready"][b].record(st["copy"])
            cur.wait_event(st["ready"][b])
            self._steps(model, dev[0, :n], dev[1, :n], dev[2, :n], adver, ws, None if stats is None else stats[s0:s1])
            st["free"][b].record(cur)


class MF:
    """APR.py:85-202.  ``extra_row`` = 1 gives the APR.py table shapes [num_users+1, d] / [num_items+1, d]
    (APR.py:107-117); evaluation_adv.MF uses exactly num_users / num_items rows (evaluation_adv.py:119-129)."""

    extra_row = 1

    def __init__(self, num_users, num_items, args):
        self.num_items = num_items
        self.num_users = num_users
        self.embedding_size = args.embed_size
        self.learning_rate = args.lr
        self.reg = args.reg
        self.dns = args.dns
        self.adv = args.adv
        self.eps = args.eps
        self.adver = args.adver
        self.reg_adv = args.reg_adv
        self.epochs = args.epochs
        self.seed = getattr(args, "seed", SEED)
        self.shuffle_count = 0
        self.adv_step = 0        # global step counter of the `--adv random` noise (one fresh Delta per update_P/update_Q run)
        self.embedding_P = None
        self.embedding_Q = None

    def build_graph(self):
        if self.embedding_size % 4 != 0 or not (4 <= self.embedding_size <= 512):
            raise ValueError("embed_size must be a multiple of 4 in [4, 512]")
        if self.adver and self.adv not in ("grad", "random"):
            raise ValueError("--adv must be 'grad' or 'random' (APR.py:170-191)")
        self.device = engine.require_cuda()
        d = self.embedding_size
        kw = dict(dtype=torch.float32, device=self.device)
        self.embedding_P = torch.empty((self.num_users + self.extra_row, d), **kw)
        self.embedding_Q = torch.empty((self.num_items + self.extra_row, d), **kw)
        self.acc_P = torch.empty_like(self.embedding_P)
        self.acc_Q = torch.empty_like(self.embedding_Q)
        self.initialize()

    def initialize(self):
        """tf.global_variables_initializer: truncated_normal(0, 0.01) tables (APR.py:107-112), Adagrad slots 0.1."""
        engine.init_truncated_normal(self.embedding_P, 0.01, self.seed, 0)
        engine.init_truncated_normal(self.embedding_Q, 0.01, self.seed, 1)
        self.reset_optimizer()

    def reset_optimizer(self):
        engine.fill(self.acc_P, ADAGRAD_INIT)
        engine.fill(self.acc_Q, ADAGRAD_INIT)

    # Recommender-style scoring with the unperturbed tables (== sess.run(model.output), utils.py:250-251)
    def predict(self, users, items) -> np.ndarray:
        u = torch.as_tensor(np.asarray(users).reshape(-1), dtype=torch.int32).to(self.device)
        i = torch.as_tensor(np.asarray(items).reshape(-1), dtype=torch.int32).to(self.device)
        return engine.score_pairs(self.embedding_P, self.embedding_Q, u, i).cpu().numpy()

    def save_weights(self, path_prefix: str, global_step: int) -> str:
        """Stand-in for tf.train.Saver({'embedding_P','embedding_Q'}).save (APR.py:222,290-292): Adagrad slots are
        NOT saved, exactly like the reference."""
        d = os.path.dirname(path_prefix)
        os.makedirs(d, exist_ok=True)
        fn = "%s-%d.npz" % (path_prefix, global_step)
        np.savez(fn, embedding_P=self.embedding_P.cpu().numpy(), embedding_Q=self.embedding_Q.cpu().numpy())
        with open(os.path.join(d, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(fn))
        return fn

    def restore_weights(self, ckpt_dir: str) -> bool:
        """tf.train.get_checkpoint_state + saver.restore (APR.py:226-230).  Returns False -- the caller then keeps the
        fresh initialisation, as the reference does -- when the directory holds no checkpoint of THIS format (e.g. it was
        written by the reference's TF Saver: index/data files, not .npz).  A checkpoint whose tables do
        not have this model's shapes (APR.MF has num_users+1 / num_items+1 rows, evaluation_adv.MF exactly num_users /
        num_items) is an error, not a silent partial restore."""
        state = os.path.join(ckpt_dir, "checkpoint")
        if not os.path.exists(state):
            return False
        with open(state) as f:
            parts = f.readline().split('"')
        if len(parts) < 2:
            return False
        fn = os.path.join(ckpt_dir, parts[1])
        if not fn.endswith(".npz") or not os.path.exists(fn):
            logging.warning("checkpoint %s is not an apr_b200 .npz checkpoint (TensorFlow Saver files cannot be read here): "
                            "continuing from the initialisation", fn)
            return False
        z = np.load(fn)
        for key, table in (("embedding_P", self.embedding_P), ("embedding_Q", self.embedding_Q)):
            if key not in z.files:
                raise ValueError("checkpoint %s has no array %r" % (fn, key))
            if tuple(z[key].shape) != tuple(table.shape):
                raise ValueError("checkpoint %s: %s has shape %s, the model expects %s (APR.MF tables carry one extra row, "
                                 "evaluation_adv.MF tables do not)" % (fn, key, tuple(z[key].shape), tuple(table.shape)))
        self.embedding_P.copy_(torch.from_numpy(z["embedding_P"]))
        self.embedding_Q.copy_(torch.from_numpy(z["embedding_Q"]))
        engine.touch(self.embedding_P)
        engine.touch(self.embedding_Q)
        return True


# ---------------------------------------------------------------------------------------------------------
# sampling / shuffle (APR.py:30-81)
# ---------------------------------------------------------------------------------------------------------
def sampling(dataset):
    """APR.py:30-36: the (u, i) keys of trainMatrix in insertion order."""
    if hasattr(dataset, "device_pairs"):       # DeviceDataset (N1): the pair list already lives on the GPU
        return dataset.device_pairs()
    if hasattr(dataset.trainMatrix, "pairs"):
        u, i = dataset.trainMatrix.pairs()
    else:
        keys = list(dataset.trainMatrix.keys())
        u = np.asarray([k[0] for k in keys], dtype=np.int32)
        i = np.asarray([k[1] for k in keys], dtype=np.int32)
    return u, i


_DEVICE_CACHE = {}


def _cached_device(key, arr: np.ndarray, dtype, device) -> torch.Tensor:
    k = (key, id(arr), str(device))
    t = _DEVICE_CACHE.get(k)
    if t is None:
        if len(_DEVICE_CACHE) > 64:
            _DEVICE_CACHE.clear()
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device=device, dtype=dtype)
        _DEVICE_CACHE[k] = (t, arr)  # keep arr alive so id() stays unique
        return t
    return t[0]


def shuffle(samples, batch_size, dataset, model, epoch: Optional[int] = None, rank: int = 0, world: int = 1,
            fork_workers: Optional[int] = None):
    """APR.py:39-61 on the GPU: epoch permutation + ``model.dns`` uniform negatives per positive, rejected against
    trainList[u]; tail batch dropped.  Counter-based: the result depends only on (model.seed, epoch).  ``epoch``
    defaults to the number of previous shuffle calls on this model.  Returns four DeviceBatches.
    Data-parallel runs: rank r of ``world`` gets columns [r*B/world, (r+1)*B/world) of every batch of the SAME epoch
    (bit-identical slices, no collective: SURVEY 8e).  ``fork_workers`` (default: model.fork_workers, 0) > 0 emulates
    the reference's fork-duplicated negative streams -- for comparing against its logs only."""
    if epoch is None:
        epoch = model.shuffle_count
    model.shuffle_count += 1
    dev = model.device
    u_h, i_h = samples
    if isinstance(u_h, torch.Tensor) and u_h.is_cuda:          # DeviceDataset: nothing to upload
        pu, pi = u_h, i_h
    else:
        u_h = np.asarray(u_h, dtype=np.int32) if not isinstance(u_h, np.ndarray) else u_h
        i_h = np.asarray(i_h, dtype=np.int32) if not isinstance(i_h, np.ndarray) else i_h
        pu = _cached_device("pairs_u", u_h, torch.int32, dev)
        pi = _cached_device("pairs_i", i_h, torch.int32, dev)
    if hasattr(dataset, "device_train_csr"):
        ptr, idx = dataset.device_train_csr()
    else:
        ptr_h, idx_h = dataset.train_csr()
        ptr = _cached_device("csr_ptr", ptr_h, torch.int64, dev)
        idx = _cached_device("csr_idx", idx_h, torch.int32, dev)
    u, i, ud, j, err = engine.sample_epoch(pu, pi, batch_size, dataset.num_items, ptr, idx, model.seed, epoch, model.dns,
                                           rank=rank, world=world,
                                           fork_workers=getattr(model, "fork_workers", 0) if fork_workers is None else fork_workers)
    model._sampler_err = err
    return DeviceBatches(u), DeviceBatches(i), DeviceBatches(ud), DeviceBatches(j)


# ---------------------------------------------------------------------------------------------------------
# training (APR.py:206-292)
# ---------------------------------------------------------------------------------------------------------
def _ckpt_paths(args, time_stamp, prefix=""):
    if args.adver:
        save = "%sPretrain/%s/APR/embed_%d/%s/" % (prefix, args.dataset, args.embed_size, time_stamp)
        restore = "%sPretrain/%s/MF_BPR/embed_%d/%s/" % (prefix, args.dataset, args.embed_size, time_stamp)
    else:
        save = "%sPretrain/%s/MF_BPR/embed_%d/%s/" % (prefix, args.dataset, args.embed_size, time_stamp)
        restore = 0 if args.restore is None else "%sPretrain/%s/MF_BPR/embed_%d/%s/" % (
            prefix, args.dataset, args.embed_size, args.restore)
    return save, restore


def training(model, dataset, args, runName, epoch_start, epoch_end, time_stamp, ckpt_prefix=""):
    """Epoch loop of APR.py:206-292 (same log lines, same best-NDCG bookkeeping, same checkpoint tree).

    As in the reference, entering ``training`` re-initialises every variable (``global_variables_initializer``,
    APR.py:225) and then restores P, Q only, so the Adagrad accumulators restart at 0.1 at the BPR->APR switch."""
    with Session() as sess:
        ckpt_save_path, ckpt_restore_path = _ckpt_paths(args, time_stamp, ckpt_prefix)
        os.makedirs(ckpt_save_path, exist_ok=True)
        if ckpt_restore_path:
            os.makedirs(ckpt_restore_path, exist_ok=True)

        model.initialize()
        if args.restore is not None or epoch_start:
            if ckpt_restore_path and model.restore_weights(ckpt_restore_path):
                print("restored")
        else:
            logging.info("Initialized from scratch")

        eval_feed_dicts = init_eval_model(dataset, args)
        samples = sampling(dataset)

        max_ndcg = 0
        best_res = {}
        ndcg = None
        epoch_count = epoch_start - 1
        for epoch_count in range(epoch_start, epoch_end + 1):
            batch_begin = time()
            batches = shuffle(samples, args.batch_size, dataset, model, epoch=epoch_count)
            torch.cuda.synchronize()
            batch_time = time() - batch_begin

            prev_batch = batches[0], batches[1], batches[3]
            _, prev_acc = training_loss_acc(model, sess, prev_batch, output_adv=0)

            train_begin = time()
            train_batches = training_batch(model, sess, batches, args.adver)
            torch.cuda.synchronize()
            train_time = time() - train_begin

            if epoch_count % args.verbose == 0:
                _, ndcg, cur_res, raw_result = output_evaluate(model, sess, dataset, train_batches, eval_feed_dicts,
                                                               epoch_count, batch_time, train_time, prev_acc, runName,
                                                               args, output_adv=0)
                if max_ndcg < ndcg:
                    max_ndcg = ndcg
                    best_res['result'] = cur_res
                    best_res['epoch'] = epoch_count
                    prediction2file(args.path + "out/" + args.opath, runName + ".hr", raw_result[:, 0, -1])
                    prediction2file(args.path + "out/" + args.opath, runName + ".ndcg", raw_result[:, 1, -1])

            if model.epochs == epoch_count and best_res:
                write2file(args.path + "out/" + args.opath, runName + ".out",
                           "Epoch %d is the best epoch" % best_res['epoch'])
                for idx, (hr_k, ndcg_k, auc_k) in enumerate(np.swapaxes(best_res['result'], 0, 1)):
                    res = "K = %d: HR = %.4f, NDCG = %.4f AUC = %.4f" % (idx + 1, hr_k, ndcg_k, auc_k)
                    write2file(args.path + "out/" + args.opath, runName + ".out", res)

            if args.ckpt > 0 and epoch_count % args.ckpt == 0:
                model.save_weights(ckpt_save_path + 'weights', epoch_count)

        if epoch_count >= epoch_start:
            model.save_weights(ckpt_save_path + 'weights', epoch_count)
        if getattr(model, "_sampler_err", None) is not None and int(model._sampler_err.item()) != 0:
            raise RuntimeError("negative sampling did not terminate for some user (every item is a train item?)")
        return best_res
