"""Data loaders with the reference's attribute surface (Dataset.py:112-327).

``OriginalDataset(path)`` / ``HeDataset(path)`` expose ``trainMatrix, trainList, testRatings, testNegatives,
num_users, num_items, df, trainSeq`` exactly as the reference's loaders do, but are built from one vectorised pass
over the TSV instead of per-line Python loops and scipy dok insertions.  Device-side views used by the CUDA path
(``pairs``, sorted CSR of trainList) are created lazily.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def _read_rating_tsv(filename: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``uid \\t iid \\t rating \\t timestamp`` -> (uid, iid, rating); the timestamp column may hold dates."""
    us, is_, rs = [], [], []
    with open(filename, "r") as f:
        for line in f:
            if line == "" or line == "\n":
                continue
            arr = line.split("\t")
            us.append(int(arr[0]))
            is_.append(int(arr[1]))
            rs.append(float(arr[2]) if len(arr) > 2 else 1.0)
    return np.asarray(us, dtype=np.int64), np.asarray(is_, dtype=np.int64), np.asarray(rs, dtype=np.float32)


class InteractionMatrix:
    """Stand-in for the scipy dok ``trainMatrix`` (Dataset.py:278-304): ``shape``, ``keys()`` in insertion order with
    duplicates collapsed, ``(u, i) in m``, ``nnz``; ``todok()`` materialises the real thing on demand."""

    def __init__(self, u: np.ndarray, i: np.ndarray, shape: Tuple[int, int]):
        self.shape = shape
        key = u.astype(np.int64) * (shape[1] + 1) + i.astype(np.int64)
        _, first = np.unique(key, return_index=True)
        first.sort()
        self._u = u[first].astype(np.int32)
        self._i = i[first].astype(np.int32)
        self._keyset = None
        self._mult = shape[1] + 1

    @property
    def nnz(self) -> int:
        return int(self._u.shape[0])

    def __len__(self) -> int:
        return self.nnz

    def keys(self):
        return list(zip(self._u.tolist(), self._i.tolist()))

    def pairs(self) -> Tuple[np.ndarray, np.ndarray]:
        return self._u, self._i

    def __contains__(self, ui) -> bool:
        if self._keyset is None:
            self._keyset = set((self._u.astype(np.int64) * self._mult + self._i).tolist())
        return int(ui[0]) * self._mult + int(ui[1]) in self._keyset

    def todok(self):
        import scipy.sparse as sp
        m = sp.coo_matrix((np.ones(self.nnz, np.float32), (self._u, self._i)), shape=self.shape)
        return m.todok()


def load_training_file_as_list(u: np.ndarray, i: np.ndarray, reproduce_quirk: bool = True) -> List[List[int]]:
    """Dataset.py:306-325.  The reference advances its user cursor by at most one per line, so when a user id is
    missing from the (user-sorted) train file the first item of the next user is filed under the missing id
    (SURVEY B.4).  ``reproduce_quirk=False`` groups by the true uid instead."""
    if not reproduce_quirk:
        n = int(u.max()) + 1 if u.size else 0
        lists: List[List[int]] = [[] for _ in range(n)]
        for a, b in zip(u.tolist(), i.tolist()):
            lists[a].append(b)
        return lists
    u_ = 0
    lists, items = [], []
    for a, b in zip(u.tolist(), i.tolist()):
        if u_ < a:
            lists.append(items)
            items = []
            u_ += 1
        items.append(b)
    lists.append(items)
    return lists


def build_sorted_csr(lists: Sequence[Iterable[int]], n_rows: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Sorted, de-duplicated CSR (int64 ptr, int32 idx) of per-row id lists."""
    n = len(lists) if n_rows is None else n_rows
    ptr = np.zeros(n + 1, dtype=np.int64)
    rows = []
    for r in range(n):
        a = np.unique(np.asarray(list(lists[r]), dtype=np.int32)) if r < len(lists) and len(lists[r]) else np.zeros(0, np.int32)
        rows.append(a)
        ptr[r + 1] = ptr[r] + a.size
    idx = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
    return ptr, idx


class _DatasetBase(object):
    reproduce_quirk = True

    def _load_train(self, filename: str):
        u, i, r = _read_rating_tsv(filename)
        self._train_u, self._train_i, self._train_r = u, i, r
        num_users = int(u.max()) + 1 if u.size else 1
        num_items = int(i.max()) + 1 if i.size else 1
        keep = r > 0
        self.trainMatrix = InteractionMatrix(u[keep], i[keep], (num_users, num_items))
        self.trainList = load_training_file_as_list(u, i, self.reproduce_quirk)
        self.num_users, self.num_items = self.trainMatrix.shape
        self._csr = None

    def load_rating_file_as_list(self, filename: str) -> List[List[int]]:
        u, i, _ = _read_rating_tsv(filename)
        return [[int(a), int(b)] for a, b in zip(u.tolist(), i.tolist())]

    def load_negative_file(self, filename: str) -> List[List[int]]:
        """Dataset.py:161-172: ``(u,i) \\t n1 ... \\t n99``."""
        out = []
        with open(filename, "r") as f:
            for line in f:
                if line == "" or line == "\n":
                    continue
                arr = line.rstrip("\n").split("\t")
                out.append([int(x) for x in arr[1:]])
        return out

    @property
    def df(self):
        """pandas frame of the train file (Dataset.py:246-247); only built when asked for."""
        if getattr(self, "_df", None) is None:
            import pandas as pd
            self._df = pd.DataFrame({"uid": self._train_u, "iid": self._train_i, "rating": self._train_r})
        return self._df

    @property
    def iid_column(self) -> np.ndarray:
        """``dataset.df.iid.tolist()`` of utils.py:186 without pandas."""
        return self._train_i

    @property
    def trainSeq(self) -> Dict[int, List[int]]:
        if getattr(self, "_trainSeq", None) is None:
            seq = defaultdict(list)
            for a, b in zip(self._train_u.tolist(), self._train_i.tolist()):
                seq[a].append(b)
            self._trainSeq = seq
        return self._trainSeq

    def train_csr(self) -> Tuple[np.ndarray, np.ndarray]:
        """Sorted CSR of ``trainList`` (rows = len(trainList)): the membership structure for negative rejection
        (APR.py:77) and for the evaluation candidate sets (utils.py:211)."""
        if self._csr is None:
            self._csr = build_sorted_csr(self.trainList)
        return self._csr


class OriginalDataset(_DatasetBase):
    """Dataset.py:226-252: ``<path>.train.rating`` + ``<path>.test.rating`` (no negatives file)."""

    def __init__(self, path: str, reproduce_quirk: bool = True):
        self.reproduce_quirk = reproduce_quirk
        self._load_train(path + ".train.rating")
        self.testRatings = self.load_rating_file_as_list(path + ".test.rating")
        self.testNegatives = None


class HeDataset(_DatasetBase):
    """Dataset.py:112-147: He et al. format with ``.test.negative`` (mode 0) or ``<path>Train/Test/TestNegative``
    (mode 1)."""

    def __init__(self, path: str, mode: int = 0, reproduce_quirk: bool = True):
        self.reproduce_quirk = reproduce_quirk
        if mode == 0:
            self._load_train(path + ".train.rating")
            self.testRatings = self.load_rating_file_as_list(path + ".test.rating")
            self.testNegatives = self.load_negative_file(path + ".test.negative")
        else:
            self._load_train(path + "Train")
            self.testRatings = self.load_rating_file_as_list(path + "Test")
            self.testNegatives = self.load_negative_file(path + "TestNegative")
        assert len(self.testRatings) == len(self.testNegatives)


class ArrayDataset(_DatasetBase):
    """Same surface built from in-memory arrays (synthetic benchmarks, tests)."""

    def __init__(self, train_u, train_i, test_u, test_i, test_negatives=None, reproduce_quirk: bool = True,
                 num_users: Optional[int] = None, num_items: Optional[int] = None):
        self.reproduce_quirk = reproduce_quirk
        u = np.asarray(train_u, dtype=np.int64)
        i = np.asarray(train_i, dtype=np.int64)
        self._train_u, self._train_i, self._train_r = u, i, np.ones(u.shape, np.float32)
        nu = int(u.max()) + 1 if num_users is None else num_users
        ni = int(i.max()) + 1 if num_items is None else num_items
        self.trainMatrix = InteractionMatrix(u, i, (nu, ni))
        self.trainList = load_training_file_as_list(u, i, reproduce_quirk)
        self.num_users, self.num_items = nu, ni
        self.testRatings = [[int(a), int(b)] for a, b in zip(np.asarray(test_u).tolist(), np.asarray(test_i).tolist())]
        self.testNegatives = test_negatives
        self._csr = None


class DeviceDataset(_DatasetBase):
    """N1: ``OriginalDataset`` / ``HeDataset`` semantics with the parsing, de-duplication and CSR
    construction done on the GPU (csrc/loader.cu): the file is read once into pinned memory, parsed by one thread per
    line, the ``sampling`` pair list, the trainList rows (incl. the reference's cursor quirk) and the sorted
    CSR come out of scans / radix sorts on the device and STAY there -- ``APR.sampling`` / ``APR.shuffle`` /
    ``utils.init_eval_model`` pick the device tensors up directly, so no per-line Python runs between the file and the
    first training step.  The reference's host attributes (``trainMatrix``, ``trainList``, ``testRatings``,
    ``testNegatives``) are materialised lazily, only if somebody asks for them.

    A train file that is not sorted by uid falls back to the host cursor for the trainList rows (the two-scan form of the
    cursor holds for sorted files, which is what process_data.py:30-51 writes)."""

    def __init__(self, path: str, reproduce_quirk: bool = True, negatives: bool = False, device=None):
        import torch

        from . import engine
        self.reproduce_quirk = reproduce_quirk
        self.device = engine.require_cuda() if device is None else device
        u, i, r = engine.parse_rating_text(engine._file_to_device(path + ".train.rating", self.device))
        self._du, self._di, self._dr = u, i, r
        self.num_users = int(u.max().item()) + 1 if u.numel() else 1
        self.num_items = int(i.max().item()) + 1 if i.numel() else 1
        self._pairs = engine.unique_pairs_device(u, i, r)
        row, is_sorted = engine.train_rows(u, reproduce_quirk)
        if reproduce_quirk and not is_sorted:
            lists = load_training_file_as_list(u.cpu().numpy().astype(np.int64), i.cpu().numpy().astype(np.int64), True)
            row = torch.from_numpy(np.repeat(np.arange(len(lists), dtype=np.int32), [len(l) for l in lists])).to(self.device)
        self._row = row
        self._n_lists = int(row.max().item()) + 1 if row.numel() else 1
        self._csr_dev = engine.build_csr_device(row, i, self._n_lists)
        tu, ti, _ = engine.parse_rating_text(engine._file_to_device(path + ".test.rating", self.device))
        self._test_u, self._test_i = tu, ti
        self._neg_dev = engine.parse_negative_text(engine._file_to_device(path + ".test.negative", self.device)) if negatives else None
        self._csr = None
        self._host = {}

    # ---- device views (what the CUDA path consumes) ----
    def device_pairs(self):
        return self._pairs

    def device_train_csr(self):
        return self._csr_dev

    def device_eval_exclusion(self):
        """Sorted CSR of trainList[u] united with {test item} for u in range(num_users) (utils.py:200-215)."""
        import torch

        from . import engine
        keep = self._row < self.num_users
        rows = torch.cat([self._row[keep], torch.arange(self.num_users, dtype=torch.int32, device=self.device)])
        items = torch.cat([self._di[keep], self.device_test_items()])
        return engine.build_csr_device(rows.contiguous(), items.contiguous(), self.num_users)

    def device_test_items(self):
        """test item of user u = testRatings[u][1] (the file holds one row per user in uid order, App. C)."""
        if self._test_i.numel() < self.num_users:
            raise ValueError("the test file has fewer rows than there are users")
        return self._test_i[:self.num_users].contiguous()

    # ---- the reference's host attributes, on demand ----
    def _h(self, name, fn):
        if name not in self._host:
            self._host[name] = fn()
        return self._host[name]

    @property
    def _train_u(self):
        return self._h("u", lambda: self._du.cpu().numpy().astype(np.int64))

    @property
    def _train_i(self):
        return self._h("i", lambda: self._di.cpu().numpy().astype(np.int64))

    @property
    def _train_r(self):
        return self._h("r", lambda: self._dr.cpu().numpy())

    @property
    def trainMatrix(self):
        keep = self._train_r > 0
        return self._h("m", lambda: InteractionMatrix(self._train_u[keep], self._train_i[keep], (self.num_users, self.num_items)))

    @property
    def trainList(self):
        def build():
            row = self._row.cpu().numpy()
            cut = np.flatnonzero(np.diff(row)) + 1
            parts = np.split(self._train_i, cut)
            lists = [[] for _ in range(self._n_lists)]
            for r, p in zip(row[np.concatenate([[0], cut])].tolist() if row.size else [], parts):
                lists[r] = lists[r] + p.tolist()
            return lists
        return self._h("l", build)

    @property
    def testRatings(self):
        return self._h("t", lambda: [[int(a), int(b)] for a, b in zip(self._test_u.cpu().tolist(), self._test_i.cpu().tolist())])

    @property
    def testNegatives(self):
        if self._neg_dev is None:
            return None
        def build():
            ptr, idx = self._neg_dev[0].cpu().numpy(), self._neg_dev[1].cpu().numpy()
            return [idx[ptr[k]:ptr[k + 1]].tolist() for k in range(ptr.size - 1)]
        return self._h("n", build)

    def train_csr(self):
        if self._csr is None:
            self._csr = (self._csr_dev[0].cpu().numpy(), self._csr_dev[1].cpu().numpy())
        return self._csr
