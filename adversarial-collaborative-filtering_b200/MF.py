"""``MatrixFactorization`` with the reference's surface (MF.py:7-59) on the CUDA engine (N3).

    MatrixFactorization(uNum, iNum, dim)
    get_train_instances(train) -> ([users, items], labels)     one sampled negative (label 0) after every positive
    train(x_train, y_train, batch_size) -> mean loss           model.fit(..., epochs=1, shuffle=True): Keras' dense Adam
    rank(users, items) -> <U[u], V[i]>;  save / load_pre_train / get_params

Arithmetic: binary_crossentropy on the raw dot product (clipped to [1e-7, 1-1e-7]) + Adam(lr 0.001, 0.9, 0.999, 1e-7) over
densified Embedding gradients, restated in oracle.keras_step (Keras / TF cannot run here: parity unpinned).  Keras'
Embedding initialiser is uniform(-0.05, 0.05); ``fit`` shuffles the instances with an unseeded NumPy permutation -- here a
permutation seeded by (seed, number of train() calls)."""
import numpy as np
import torch

from . import engine
from .Recommender import Recommender


class _KerasMFBase(Recommender):
    loss_kind = 0

    def __init__(self, uNum, iNum, dim, lr=0.001, seed=2019):
        self.uNum, self.iNum, self.dim = uNum, iNum, dim
        self.lr, self.seed = lr, seed
        self.device = engine.require_cuda()
        g = torch.Generator(device=self.device)
        g.manual_seed(seed)
        kw = dict(device=self.device, dtype=torch.float32)
        self.U = (torch.rand((uNum, dim), generator=g, **kw) * 0.1 - 0.05).contiguous()     # Keras 'uniform' initialiser
        self.V = (torch.rand((iNum, dim), generator=g, **kw) * 0.1 - 0.05).contiguous()
        self.state = [torch.zeros_like(t) for t in (self.U, self.U, self.V, self.V, self.U, self.V)]   # mU vU mV vV gU gV
        self.iterations = 0
        self.epochs_run = 0
        self._rng = np.random.RandomState(seed)

    def get_params(self):
        return ""

    def load_pre_train(self, pre):
        pass

    def save(self, path):
        pass

    def rank(self, users, items):
        u = torch.as_tensor(np.asarray(users).reshape(-1), dtype=torch.int32).to(self.device)
        i = torch.as_tensor(np.asarray(items).reshape(-1), dtype=torch.int32).to(self.device)
        return engine.score_pairs(self.U, self.V, u, i).cpu().numpy()

    def _negative(self, train, u):
        j = int(self._rng.randint(1, self.iNum))
        while (u, j) in train:
            j = int(self._rng.randint(1, self.iNum))
        return j

    def _fit(self, cols, y, batch_size):
        """model.fit(x, y, batch_size, epochs=1, shuffle=True): every batch (the last one may be short) is one Adam step;
        history['loss'] is the sample-weighted mean of the batch losses = total loss / number of instances."""
        n = cols[0].shape[0]
        order = np.random.RandomState((self.seed * 1000003 + self.epochs_run) & 0x7FFFFFFF).permutation(n)
        self.epochs_run += 1
        dev = self.device
        tcols = [torch.from_numpy(np.ascontiguousarray(np.asarray(c).reshape(-1)[order]).astype(np.int32)).to(dev) for c in cols]
        ty = None if y is None else torch.from_numpy(np.asarray(y, dtype=np.float32).reshape(-1)[order]).to(dev)
        mU, vU, mV, vV, gU, gV = self.state
        total = torch.zeros(1, dtype=torch.float64, device=dev)
        for a in range(0, n, batch_size):
            b = min(n, a + batch_size)
            self.iterations += 1
            engine.keras_step(self.U, self.V, mU, vU, mV, vV, gU, gV, tcols[0][a:b], tcols[1][a:b],
                              j=tcols[2][a:b] if len(tcols) > 2 else None, y=None if ty is None else ty[a:b], lr=self.lr,
                              t=self.iterations, loss_sum=total)
        return float(total.item()) / n


class MatrixFactorization(_KerasMFBase):
    """MF.py:7-59."""

    def get_train_instances(self, train):
        user_input, item_input, labels = [], [], []
        for (u, i) in train.keys():
            user_input += [u, u]
            item_input += [i, self._negative(train, u)]
            labels += [1, 0]
        return [np.array(user_input), np.array(item_input)], np.array(labels)

    def train(self, x_train, y_train, batch_size):
        return self._fit([np.asarray(x_train[0]), np.asarray(x_train[1])], y_train, batch_size)
