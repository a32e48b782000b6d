"""Wall-clock of one epoch on the ml-1m-sort SHAPE (BASELINE.json configs[0]: U=6040, I=3706, ~994k pairs, batch 512,
d=64): sampler, BPR epoch, APR epoch, loss/acc pass, full-rank evaluation -- the numbers the reference logs per epoch."""
import os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apr_b200.APR import MF, Session, sampling, shuffle
from apr_b200.Dataset import ArrayDataset
from apr_b200.utils import evaluate, init_eval_model, training_batch, training_loss_acc
rng = np.random.default_rng(0)
U, I = 6040, 3706
cnt = np.clip(rng.lognormal(4.6, 0.9, U).astype(int), 20, 2300)
cnt = (cnt * (994169 / cnt.sum())).astype(int).clip(19, I - 10)
w = 1.0 / np.arange(1, I + 1) ** 0.9; w /= w.sum()
tu, ti, eu, ei = [], [], [], []
for u in range(U):
    it = rng.choice(I, size=cnt[u] + 1, replace=False, p=w)
    eu.append(u); ei.append(int(it[0])); tu += [u] * cnt[u]; ti += it[1:].tolist()
ds = ArrayDataset(np.asarray(tu), np.asarray(ti), np.asarray(eu), np.asarray(ei))
print("pairs", len(tu), "users", ds.num_users, "items", ds.num_items)
for mode in (0, 2):
    args = types.SimpleNamespace(embed_size=64, lr=0.05, reg=0.0, dns=1, adv="grad", eps=0.5, adver=0, reg_adv=1.0, epochs=0, seed=2019, batch_size=512, eval_mode="all")
    model = MF(ds.num_users, ds.num_items, args); model.build_graph()
    sess = Session(mode=mode); samples = sampling(ds); feed = init_eval_model(ds, args)
    def timed(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r
    for epoch, adver in ((0, 0), (1, 0), (2, 1), (3, 1)):
        model.adver = adver
        ts, batches = timed(lambda: shuffle(samples, 512, ds, model, epoch=epoch))
        tt, _ = timed(lambda: training_batch(model, sess, batches, adver))
        tl, _ = timed(lambda: training_loss_acc(model, sess, (batches[0], batches[1], batches[3]), 0))
        te, (res, _) = timed(lambda: evaluate(model, sess, ds, feed, 0, args))
        n = len(batches[0]) * 512
        print("mode %d epoch %d adver %d: sampler %.1f ms | steps %.1f ms (%.1f M triples/s, %.1f us/step) | loss/acc %.1f ms | eval %.1f ms (%.0f users/s) HR@10 %.4f" %
              (mode, epoch, adver, ts * 1e3, tt * 1e3, n / tt / 1e6, tt / len(batches[0]) * 1e6, tl * 1e3, te * 1e3, U / te, res[0][9]))
