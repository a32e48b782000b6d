"""N1: the device-side loader against the host loader that mirrors Dataset.py:112-327 -- same pair list (file
order, duplicates collapsed, rating > 0 only), same trainList rows incl. the cursor quirk, same sorted CSR,
same test vector, same He-format negatives -- and an epoch sampled / trained from it equals one from the host dataset."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FIX = os.path.join(os.path.dirname(__file__), "golden", "video_interactions.npz")


def _write_files(tmp_path, tu, ti, eu, ei, rng, negatives=False):
    """TSV files in the formats of SURVEY App. C: integer / date / constant timestamps, a few rating-0 lines, duplicate
    pairs, blank lines, a missing final newline."""
    base = str(tmp_path / "ds")
    lines = []
    for k, (a, b) in enumerate(zip(tu.tolist(), ti.tolist())):
        rating = "0" if k % 97 == 5 else ("4.5" if k % 11 == 0 else "%d" % (1 + k % 5))
        ts = "2014-07-%02d 10:%02d:00" % (1 + k % 28, k % 60) if k % 3 == 0 else ("1" if k % 3 == 1 else str(978300000 + k))
        lines.append("%d\t%d\t%s\t%s" % (a, b, rating, ts))
        if k % 1000 == 7:
            lines.append(lines[-1])            # duplicate (u, i) pair
        if k % 5000 == 11:
            lines.append("")                   # blank line
    open(base + ".train.rating", "w").write("\n".join(lines))          # no trailing newline
    open(base + ".test.rating", "w").write("".join("%d\t%d\t1\t1\n" % (a, b) for a, b in zip(eu.tolist(), ei.tolist())))
    if negatives:
        with open(base + ".test.negative", "w") as f:
            for a, b in zip(eu.tolist(), ei.tolist()):
                f.write("(%d,%d)\t%s\n" % (a, b, "\t".join(str(x) for x in rng.randint(0, 1000, rng.randint(1, 12)).tolist())))
    return base


def test_device_loader_matches_host_loader_on_video(cuda_device, tmp_path):
    from apr_b200.Dataset import DeviceDataset, OriginalDataset
    z = np.load(FIX)
    rng = np.random.RandomState(0)
    n = 60000                                                       # a prefix: the 30 skipped users of SURVEY B.4 start early
    base = _write_files(tmp_path, z["train_u"][:n], z["train_i"][:n], z["test_u"], z["test_i"], rng)
    host = OriginalDataset(base)
    devd = DeviceDataset(base)
    assert (devd.num_users, devd.num_items) == (host.num_users, host.num_items)
    hu, hi = host.trainMatrix.pairs()
    du, di = devd.device_pairs()
    assert np.array_equal(du.cpu().numpy(), hu) and np.array_equal(di.cpu().numpy(), hi)
    hptr, hidx = host.train_csr()
    dptr, didx = devd.device_train_csr()
    assert np.array_equal(dptr.cpu().numpy(), hptr) and np.array_equal(didx.cpu().numpy(), hidx)
    assert devd.trainList == host.trainList                        # the quirk is reproduced (Video has skipped uids)
    assert any(len(l) and k < len(host.trainList) for k, l in enumerate(host.trainList))
    assert devd.testRatings == host.testRatings
    # quirk off: rows are the true uids
    h2, d2 = OriginalDataset(base, reproduce_quirk=False), DeviceDataset(base, reproduce_quirk=False)
    assert np.array_equal(d2.device_train_csr()[0].cpu().numpy(), h2.train_csr()[0])
    assert np.array_equal(d2.device_train_csr()[1].cpu().numpy(), h2.train_csr()[1])


def test_device_loader_feeds_the_same_epoch_and_evaluation(cuda_device, tmp_path):
    from apr_b200.APR import MF, Session, sampling, shuffle
    from apr_b200.Dataset import DeviceDataset, HeDataset
    from apr_b200.utils import eval_positions, init_eval_model, training_batch
    rng = np.random.RandomState(3)
    U, I = 400, 300
    tu, ti, eu, ei = [], [], [], []
    for u in range(U):
        if u in (17, 230):
            eu.append(u); ei.append(int(rng.randint(I)))              # users missing from the train file (cursor quirk)
            continue
        items = rng.choice(I, size=rng.randint(3, 25), replace=False)
        eu.append(u); ei.append(int(items[0]))
        tu += [u] * (items.size - 1); ti += items[1:].tolist()
    base = _write_files(tmp_path, np.asarray(tu), np.asarray(ti), np.asarray(eu), np.asarray(ei), rng, negatives=True)
    host, devd = HeDataset(base), DeviceDataset(base, negatives=True)
    assert devd.testNegatives == host.testNegatives
    args = types.SimpleNamespace(embed_size=32, lr=0.05, reg=0.0, dns=1, adv="grad", eps=0.5, adver=1, reg_adv=1.0, epochs=1,
                                 seed=2019, batch_size=128, eval_mode="all")
    out, models, feeds = [], [], []
    for ds in (host, devd):
        model = MF(ds.num_users, ds.num_items, args)
        model.build_graph()
        feed = init_eval_model(ds, args)
        batches = shuffle(sampling(ds), 128, ds, model, epoch=0)
        with Session() as sess:
            training_batch(model, sess, batches, 1)
        models.append(model)
        feeds.append(feed)
        out.append((batches[0].numpy(), batches[3].numpy(), feed.excl_ptr_h, feed.excl_idx_h))
    for a, b in zip(*out):                   # the same epoch (bit-identical triples) and the same evaluation feed
        assert np.array_equal(a, b)
    # ONE model through both feeds: identical positions.  (The two trained models are not compared bit for bit: the step
    # sums duplicate rows with floating-point REDs in arrival order, so two runs may differ in the last bit and a
    # position at a near-tie may move by one.)
    p0 = eval_positions(models[0], feeds[0], exact=True).cpu().numpy()
    p1 = eval_positions(models[0], feeds[1], exact=True).cpu().numpy()
    assert np.array_equal(p0, p1)
    p2 = eval_positions(models[1], feeds[1], exact=True).cpu().numpy()
    assert (np.abs(p2.astype(np.int64) - p0) <= 1).mean() >= 0.99


def test_malformed_rating_file_raises(cuda_device, tmp_path):
    from apr_b200 import engine
    p = tmp_path / "bad.train.rating"
    p.write_text("0\t1\t5\t1\nx\t2\t5\t1\n")
    with pytest.raises(ValueError):
        engine.parse_rating_text(engine._file_to_device(str(p), cuda_device))
