"""Builds tests/golden/ml1m_test_items.npy from the reference's data/ml-1m-sort.test.rating (run in the build container,
where /root/reference exists; the GPU box only sees the committed fixture).

The reference ships the ml-1m TEST file only (data/ml-1m-sort.train.rating is a missing blob), so the
configs[0] parity test (tests/test_gpu_ml1m_shape.py) joins these real held-out items -- one (user, item) row per user,
users 0..6039 in order -- with a synthetic ml-1m-shaped train set.

    python tests/golden/make_ml1m_fixture.py
"""
import os

import numpy as np

SRC = "/root/reference/data/ml-1m-sort.test.rating"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ml1m_test_items.npy")

if __name__ == "__main__":
    rows = [l.split("\t") for l in open(SRC).read().splitlines() if l.strip()]
    users = np.asarray([int(r[0]) for r in rows], dtype=np.int32)
    items = np.asarray([int(r[1]) for r in rows], dtype=np.int32)
    assert np.array_equal(users, np.arange(users.size)), "one row per user, in uid order"
    np.save(DST, items)
    print("wrote", DST, items.shape, "max item id", int(items.max()))
