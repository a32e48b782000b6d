"""``BPR`` with the reference's surface (BPR.py:23-102) on the CUDA engine (N3): the Keras pairwise model --
loss mean(1 - log sigmoid(<u,i> - <u,j>)) (BPR.py:11-21), Adam -- not to be confused with the TensorFlow BPR-MF phase
of APR.py (softplus loss, Adagrad), which is ``apr_b200.APR.MF`` with adver = 0.  See apr_b200/MF.py for the restated
Keras semantics and the parity status."""
import numpy as np

from .MF import _KerasMFBase


class BPR(_KerasMFBase):
    loss_kind = 1

    def __init__(self, uNum, iNum, dim, lr=0.001, seed=2019):
        super().__init__(uNum, iNum, dim, lr, seed)
        self.dns = 1

    def get_train_instances(self, train):
        user_input, pos_item_input, neg_item_input, labels = [], [], [], []
        for (u, i) in train.keys():
            user_input.append(u)
            pos_item_input.append(i)
            neg_item_input.append(self._negative(train, u))
            labels.append(1)
        return [np.array(user_input), np.array(pos_item_input), np.array(neg_item_input)], np.array(labels)

    def train(self, x_train, y_train, batch_size):
        return self._fit([np.asarray(x_train[0]), np.asarray(x_train[1]), np.asarray(x_train[2])], None, batch_size)
