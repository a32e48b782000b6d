"""``Recommender`` plug-ins backed by the CUDA BPR / APR engine (the reference's L1'' seam: Recommender.py:3-27,
implemented there by MF.py:7-59, BPR.py:23-102, FastAdversarialMF.py:13-145).

    get_train_instances(train) -> ([users, pos_items, neg_items], labels)     (BPR.py:83-99 layout)
    train(x_train, y_train, batch_size) -> mean BPR loss of the pass          (one epoch over x_train, in order)
    rank(users, items) -> scores                                              (MF.py:37-39)
"""
import types

import numpy as np
import torch

from . import engine
from .APR import MF, Session
from .Recommender import Recommender
from .utils import as_device_batches


class BPRRecommender(Recommender):
    """BPR-MF (adver=0) or APR (adver=1) behind the six Recommender methods."""

    def __init__(self, uNum, iNum, dim, lr=0.05, reg=0.0, adver=0, eps=0.5, reg_adv=1.0, seed=2019):
        self.uNum, self.iNum, self.dim = uNum, iNum, dim
        self.dns = 1
        args = types.SimpleNamespace(embed_size=dim, lr=lr, reg=reg, dns=1, adv="grad", eps=eps, adver=adver,
                                     reg_adv=reg_adv, epochs=0, seed=seed)
        # Keras Embedding(input_dim=uNum) has exactly uNum rows (MF.py:17-18)
        self.mf = MF(uNum, iNum, args)
        self.mf.extra_row = 0
        self.mf.build_graph()
        self.sess = Session()
        self._rng = np.random.RandomState(seed)

    def get_params(self):
        return "_d%d_lr%g_adv%d_eps%g" % (self.dim, self.mf.learning_rate, self.mf.adver, self.mf.eps)

    def load_pre_train(self, pre):
        z = np.load(pre)
        self.mf.embedding_P.copy_(torch.from_numpy(z["embedding_P"]))
        self.mf.embedding_Q.copy_(torch.from_numpy(z["embedding_Q"]))
        self.mf.reset_optimizer()

    def save(self, path):
        np.savez(path, embedding_P=self.mf.embedding_P.cpu().numpy(), embedding_Q=self.mf.embedding_Q.cpu().numpy())

    def get_train_instances(self, train):
        """One negative per positive, j ~ randint(1, iNum) rejected against ``train`` (BPR.py:83-99)."""
        if hasattr(train, "pairs"):
            u, i = train.pairs()
            keyset = set((u.astype(np.int64) * (self.iNum + 1) + i).tolist())
            contains = lambda a, b: a * (self.iNum + 1) + b in keyset
        else:
            keys = list(train.keys())
            u = np.asarray([k[0] for k in keys], dtype=np.int32)
            i = np.asarray([k[1] for k in keys], dtype=np.int32)
            contains = lambda a, b: (a, b) in train
        j = self._rng.randint(1, self.iNum, size=u.shape[0])
        for k in range(u.shape[0]):
            while contains(int(u[k]), int(j[k])):
                j[k] = self._rng.randint(1, self.iNum)
        return [np.asarray(u), np.asarray(i), j.astype(np.int32)], np.ones(u.shape[0])

    def train(self, x_train, y_train, batch_size):
        users, pos, neg = [np.asarray(a).reshape(-1) for a in x_train[:3]]
        n = (users.shape[0] // batch_size) * batch_size
        if n == 0:
            raise ValueError("fewer instances than one batch")
        dev = self.mf.device
        U = as_device_batches(users[:n].reshape(-1, batch_size), dev)
        I = as_device_batches(pos[:n].reshape(-1, batch_size), dev)
        J = as_device_batches(neg[:n].reshape(-1, batch_size), dev)
        stats = torch.zeros((U.shape[0], 2), dtype=torch.float32, device=dev)
        self.sess.train_steps(self.mf, U, I, J, adver=bool(self.mf.adver), stats=stats)
        return float(stats[:, 0].sum().item() / n)

    def rank(self, users, items):
        return self.mf.predict(users, items)

    def rank_all(self, users, item_lo=0, item_hi=None):
        """N4: `all_rating` of IRGAN.py:36-39 / APL.py:205-211 -- every item's score for each listed user, [n, items]."""
        u = torch.as_tensor(np.asarray(users).reshape(-1), dtype=torch.int32).to(self.mf.device)
        return engine.score_all_items(self.mf.embedding_P, self.mf.embedding_Q, u, item_lo, item_hi).cpu().numpy()

    def recommend(self, users, k, exclude_ptr=None, exclude_idx=None):
        """Top-k item ids per user over the whole catalogue on the tensor cores (score desc, ties to the smaller id);
        ``exclude_*`` = sorted CSR of items to leave out (e.g. the train items)."""
        dev = self.mf.device
        u = torch.as_tensor(np.asarray(users).reshape(-1), dtype=torch.int32).to(dev)
        n = u.numel()
        if exclude_ptr is None:
            exclude_ptr, exclude_idx = np.zeros(n + 1, np.int64), np.zeros(0, np.int32)
        ptr = torch.as_tensor(np.asarray(exclude_ptr), dtype=torch.int64).to(dev)
        idx = torch.as_tensor(np.asarray(exclude_idx), dtype=torch.int32).to(dev)
        spos = torch.full((n,), 3.0e38, dtype=torch.float32, device=dev)   # no held-out item here: a finite score nothing reaches
        _, ids, sc = engine.eval_fullrank(self.mf.embedding_P, self.mf.embedding_Q, u, torch.zeros(n, dtype=torch.int32, device=dev),
                                          0, self.iNum, ptr, idx, k, exact=True) if not engine.tc_supported(self.dim) or self.iNum < 1024 \
            else engine.eval_fullrank_tc(self.mf.embedding_P, self.mf.embedding_Q, u, None, 0, self.iNum, ptr, idx, check=False,
                                         k_top=k, spos=spos)[:3]
        return ids.cpu().numpy(), sc.cpu().numpy()

    def rank_batched(self, users_items):
        """Scores for many (user, candidate list) pairs in one launch (used by apr_b200.evaluation)."""
        lens = [len(it) for _, it in users_items]
        uu = np.repeat(np.asarray([u for u, _ in users_items], dtype=np.int32), lens)
        ii = np.concatenate([np.asarray(it, dtype=np.int32) for _, it in users_items])
        s = self.mf.predict(uu, ii)
        return np.split(s, np.cumsum(lens)[:-1])
