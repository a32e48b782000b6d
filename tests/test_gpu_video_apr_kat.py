"""Statistical known-answer test of the ADVERSARIAL phase against the reference's own log.

The only outputs of the reference that exist are its logs (no TensorFlow here, no golden tensors upstream).  Its Video run
(out/janEval/Video_apr_d64_e0.500000_l1.000000_2020_01_24_12_07_42.out) trains 1000 BPR epochs, then switches to APR:

    line 52  Epoch 980  : HR = 0.2361, NDCG = 0.0641 ... |P|=607.17, |Q|=542.44          (last logged BPR epoch)
    line 53  Initialize APR
    line 54  Epoch 1000 : HR = 0.2318, NDCG = 0.0622 ACC = 0.9996 ACC_adv = 1.0000 ... |P|=619.33, |Q|=562.31
    line 55  Epoch 1020 : HR = 0.2431, NDCG = 0.0654 ... |P|=676.89, |Q|=627.78

i.e. the phase-switch signature: the norms, which crept up by ~1.4 / 1.1 per 20 BPR epochs, jump by ~11 / ~19 in the FIRST
adversarial epoch and by another ~58 / ~65 in the next 20, the training accuracy on the epoch's own batches reaches
1.0000, and HR@100 dips and then rises above the BPR plateau.  None of that can come out of the BPR arithmetic: it is the
Delta = eps * G/||G|| term, the second forward and the restarted Adagrad accumulators (APR.py:130-141,158-165,180-191,
222-232) -- so this run is where the adversarial step kernels touch a reference-held number.

The run here: the real Video data (tests/golden/video_interactions.npz), d = 64, batch 512, lr 0.05, eps 0.5, reg_adv 1,
the trainList quirk reproduced, the GPU sampler in its fork-emulating mode (26 workers), 1000 BPR epochs +
21 APR epochs through the same Session / shuffle / training_batch path the drivers use (about 15 s on a B200).
The reference is unseeded, so the comparison is statistical: bands of a few percent around the logged values.
"""
import math
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FIX = os.path.join(os.path.dirname(__file__), "golden", "video_interactions.npz")
# (HR@100, NDCG@100, |P|, |Q|) as logged
LOG_980 = (0.2361, 0.0641, 607.17, 542.44)
LOG_1000 = (0.2318, 0.0622, 619.33, 562.31)
LOG_1020 = (0.2431, 0.0654, 676.89, 627.78)


@pytest.mark.timeout(1500)
def test_video_bpr_to_apr_phase_switch_matches_reference_log(cuda_device):
    from apr_b200 import engine
    from apr_b200.APR import MF, Session, sampling, shuffle
    from apr_b200.Dataset import ArrayDataset
    from apr_b200.utils import evaluate, init_eval_model, training_batch, training_loss_acc
    z = np.load(FIX)
    ds = ArrayDataset(z["train_u"], z["train_i"], z["test_u"], z["test_i"], reproduce_quirk=True)
    args = types.SimpleNamespace(embed_size=64, lr=0.05, reg=0.0, dns=1, adv="grad", eps=0.5, adver=0, reg_adv=1.0, epochs=1021,
                                 seed=2019, batch_size=512, eval_mode="all")
    model = MF(ds.num_users, ds.num_items, args)
    model.fork_workers = 26
    model.build_graph()
    feed = init_eval_model(ds, args)
    samples = sampling(ds)

    def norms():
        return (math.sqrt(float(engine.sum_squares(model.embedding_P).item())),
                math.sqrt(float(engine.sum_squares(model.embedding_Q).item())))

    def snapshot(sess, tag):
        (hr, ndcg, _), _ = evaluate(model, sess, ds, feed, 0, args)
        nP, nQ = norms()
        print("%s: HR@100 %.4f NDCG@100 %.4f |P| %.2f |Q| %.2f" % (tag, hr[-1], ndcg[-1], nP, nQ))
        return hr[-1], ndcg[-1], nP, nQ

    with Session() as sess:
        for epoch in range(981):                               # BPR epochs 0 .. 980
            training_batch(model, sess, shuffle(samples, 512, ds, model, epoch=epoch), 0)
        hr980, nd980, p980, q980 = snapshot(sess, "epoch 980 (BPR)")
        for epoch in range(981, 1000):
            training_batch(model, sess, shuffle(samples, 512, ds, model, epoch=epoch), 0)
        p999, q999 = norms()
        # phase switch: global_variables_initializer + restore P, Q only -> Adagrad slots back to 0.1 (APR.py:222-232)
        model.adver = 1
        model.reset_optimizer()
        batches = shuffle(samples, 512, ds, model, epoch=1000)
        _, prev_acc = training_loss_acc(model, sess, (batches[0], batches[1], batches[3]), 0)
        training_batch(model, sess, batches, 1)
        _, post_acc = training_loss_acc(model, sess, (batches[0], batches[1], batches[3]), 0)
        hr1000, nd1000, p1000, q1000 = snapshot(sess, "epoch 1000 (first APR epoch)")
        print("ACC %.4f ACC_adv %.4f  d|P| %.2f d|Q| %.2f (log: 0.9996, 1.0000, ~+10.8, ~+18.8)" %
              (prev_acc, post_acc, p1000 - p999, q1000 - q999))
        for epoch in range(1001, 1021):
            training_batch(model, sess, shuffle(samples, 512, ds, model, epoch=epoch), 1)
        hr1020, nd1020, p1020, q1020 = snapshot(sess, "epoch 1020 (APR)")

    rel = lambda got, want: abs(got - want) / want
    # the BPR plateau the adversarial phase starts from
    # (measured on a B200, round 2: 605.50 / 543.62, HR 0.2372, NDCG 0.0661 -- within 0.3 % of the logged norms)
    assert rel(p980, LOG_980[2]) < 0.02 and rel(q980, LOG_980[3]) < 0.02
    assert abs(hr980 - LOG_980[0]) < 0.01 and abs(nd980 - LOG_980[1]) < 0.005
    # first adversarial epoch: accuracy on its own batches saturates, the norms jump by an order of magnitude more than
    # a BPR epoch moves them (logged +10.8 / +18.8 against +0.07 / +0.055 per BPR epoch)
    assert prev_acc > 0.999 and post_acc > 0.9995
    # (measured: +10.96 / +18.22, |P| 617.71, |Q| 562.84)
    assert 0.7 * 10.8 < p1000 - p999 < 1.3 * 10.8
    assert 0.7 * 18.8 < q1000 - q999 < 1.3 * 18.8
    assert rel(p1000, LOG_1000[2]) < 0.02 and rel(q1000, LOG_1000[3]) < 0.02
    assert abs(hr1000 - LOG_1000[0]) < 0.01
    # twenty epochs later
    # (measured: 675.40 / 627.15, HR 0.2446, NDCG 0.0673)
    assert rel(p1020, LOG_1020[2]) < 0.02 and rel(q1020, LOG_1020[3]) < 0.02
    assert abs(hr1020 - LOG_1020[0]) < 0.01 and abs(nd1020 - LOG_1020[1]) < 0.005
    assert hr1020 > hr980                                      # the adversarial phase lifts HR@100 above the BPR plateau
