"""Out-of-bounds guard of our own (compute-sanitizer is closed on this GPU pool: profiles/r2s_compute_sanitizer_closed.log).

Every buffer the library writes -- tables, Adagrad accumulators, the training workspace, stats, the evaluation workspace,
positions, top-k outputs -- is carved out of one arena with 64 KB canary zones on both sides; after training steps in all
modes (BPR / APR / --adv random / adver 3), the exact and the tensor-core evaluation with top-k and the sampler, every
canary byte must be untouched and the results must still equal the oracle's.  Catches writes outside a buffer (index
arithmetic, ragged last tiles, hash-table probes, list overflows); reads outside are caught where they matter -- they
would change results that are compared with the oracle in the same run."""
import numpy as np
import pytest
import torch

from oracle import apr_oracle as O

pytestmark = pytest.mark.gpu

CANARY = 64 * 1024
PATTERN = 0xA5


class Arena(object):
    def __init__(self, dev, nbytes):
        self.buf = torch.full((nbytes,), PATTERN, dtype=torch.uint8, device=dev)
        self.off = 0
        self.zones = []

    def take(self, shape, dtype, fill=None):
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        self.off = (self.off + 1023) // 1024 * 1024
        self.zones.append((self.off, self.off + CANARY))
        self.off += CANARY
        t = self.buf[self.off:self.off + n].view(dtype).view(*shape)
        self.off += (n + 1023) // 1024 * 1024
        self.zones.append((self.off, self.off + CANARY))
        self.off += CANARY
        assert self.off <= self.buf.numel()
        if fill is not None:
            t.copy_(fill) if isinstance(fill, torch.Tensor) else t.fill_(fill)
        return t

    def check(self, what):
        torch.cuda.synchronize()
        for a, b in self.zones:
            z = self.buf[a:b]
            assert bool((z == PATTERN).all().item()), "canary zone [%d, %d) overwritten after %s" % (a, b, what)


def test_training_writes_stay_inside_their_buffers(cuda_device):
    from apr_b200 import _lib, engine
    dev = cuda_device
    rng = np.random.RandomState(9)
    for (U, I, d, S, B) in ((301, 203, 64, 3, 512), (67, 53, 24, 2, 100), (4001, 3001, 128, 2, 2048), (5000, 3000, 128, 2, 3000)):
        P = (rng.randn(U, d) * 0.1).astype(np.float32)
        Q = (rng.randn(I, d) * 0.1).astype(np.float32)
        u, i, j = [rng.randint(0, n, (S, B)).astype(np.int32) for n in (U, I, I)]
        ws_bytes = _lib.lib().apr_train_workspace_bytes(S, B, d)
        ar = Arena(dev, 4 * (U + I) * d * 4 + ws_bytes + 3 * S * B * 4 + 40 * CANARY + (1 << 20))
        tP, tQ = ar.take((U, d), torch.float32), ar.take((I, d), torch.float32)
        aP, aQ = ar.take((U, d), torch.float32), ar.take((I, d), torch.float32)
        stats = ar.take((S, 2), torch.float32)
        tu, ti, tj = [ar.take((S, B), torch.int32, torch.from_numpy(x).to(dev)) for x in (u, i, j)]
        wsbuf = ar.take((ws_bytes,), torch.uint8)
        ws = engine.TrainWorkspace.__new__(engine.TrainWorkspace)
        ws.n_steps, ws.batch, ws.d, ws.nbytes, ws.buf = S, B, d, ws_bytes, wsbuf
        for adver in (0, 1, 2, 3):
            rP, rQ = P.copy(), Q.copy()
            raP, raQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
            for s in range(S):
                if adver == 2:
                    O.apr_step_random(rP, rQ, raP, raQ, u[s], i[s], j[s], 0.05, 0.01, 1.0, 0.5, 2019, s)
                else:
                    O.apr_step(rP, rQ, raP, raQ, u[s], i[s], j[s], 0.05, 0.01, 1.0, 0.5, adver)
            for mode in ((0,) if adver == 2 else (0, 1, 2)):
                tP.copy_(torch.from_numpy(P)); tQ.copy_(torch.from_numpy(Q)); aP.fill_(0.1); aQ.fill_(0.1)
                _lib.check(_lib.lib().apr_train_workspace_init(wsbuf.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream))
                if adver == 2:
                    engine.train_steps_random(tP, tQ, aP, aQ, tu, ti, tj, 0.05, 0.01, 1.0, 0.5, ws, 2019, 0, stats=stats)
                else:
                    engine.train_steps(tP, tQ, aP, aQ, tu, ti, tj, 0.05, 0.01, 1.0, 0.5, adver, ws, mode=mode, stats=stats)
                ar.check("train (U=%d d=%d B=%d adver=%d mode=%d)" % (U, d, B, adver, mode))
                assert np.abs(tP.cpu().numpy() - rP).max() <= 1e-5 * np.abs(rP).max()
                assert np.abs(tQ.cpu().numpy() - rQ).max() <= 1e-5 * np.abs(rQ).max()


def test_evaluation_writes_stay_inside_their_buffers(cuda_device):
    from apr_b200 import _lib, engine
    from apr_b200.Dataset import build_sorted_csr
    dev = cuda_device
    rng = np.random.RandomState(10)
    L = _lib.lib()
    for (U, I, d, K) in ((130, 1500, 64, 10), (257, 5003, 128, 100), (33, 1100, 40, 128), (200, 20001, 256, 100)):
        P = rng.randn(U + 1, d).astype(np.float32)
        Q = rng.randn(I + 1, d).astype(np.float32)
        train = [sorted(set(rng.randint(0, I, rng.randint(0, 30)).tolist())) for _ in range(U)]
        train[3] = sorted(set(rng.randint(0, I, I // 2).tolist()))           # exact per-user kernel
        test = rng.randint(0, I, U).astype(np.int32)
        ptr, idx = build_sorted_csr([train[k] + [int(test[k])] for k in range(U)])
        nb_tc = L.apr_eval_tc_topk_workspace_bytes(U, I, d, K)
        nb_ex = L.apr_eval_workspace_bytes(U, K, d)
        ar = Arena(dev, (U + I + 2) * d * 4 + nb_tc + nb_ex + U * K * 16 + idx.size * 4 + 60 * CANARY + (1 << 20))
        tP, tQ = ar.take((U + 1, d), torch.float32, torch.from_numpy(P).to(dev)), ar.take((I + 1, d), torch.float32, torch.from_numpy(Q).to(dev))
        users = ar.take((U,), torch.int32, torch.arange(U, dtype=torch.int32, device=dev))
        ttest = ar.take((U,), torch.int32, torch.from_numpy(test).to(dev))
        tptr = ar.take((U + 1,), torch.int64, torch.from_numpy(ptr).to(dev))
        tidx = ar.take((idx.size,), torch.int32, torch.from_numpy(idx).to(dev))
        pos_e, pos_t = ar.take((U,), torch.int32, 0), ar.take((U,), torch.int32, 0)
        ids_e, ids_t = ar.take((U, K), torch.int32), ar.take((U, K), torch.int32)
        sc_e, sc_t = ar.take((U, K), torch.float32), ar.take((U, K), torch.float32)
        err = ar.take((1,), torch.int32, 0)
        ws_e = ar.take((nb_ex,), torch.uint8)
        ws_t = ar.take((nb_tc + 1024,), torch.uint8)
        off = (-ws_t.data_ptr()) % 1024
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.apr_eval_fullrank(tP.data_ptr(), tQ.data_ptr(), d, users.data_ptr(), ttest.data_ptr(), U, 0, I, tptr.data_ptr(),
                                       tidx.data_ptr(), K, pos_e.data_ptr(), ids_e.data_ptr(), sc_e.data_ptr(), 1, ws_e.data_ptr(),
                                       nb_ex, st))
        ar.check("exact evaluation (U=%d I=%d d=%d K=%d)" % (U, I, d, K))
        _lib.check(L.apr_eval_fullrank_tc_topk(tP.data_ptr(), tQ.data_ptr(), d, users.data_ptr(), ttest.data_ptr(), U, 0, I,
                                               tptr.data_ptr(), tidx.data_ptr(), K, pos_t.data_ptr(), ids_t.data_ptr(),
                                               sc_t.data_ptr(), 0, 0, ws_t.data_ptr() + off, nb_tc, err.data_ptr(), st))
        ar.check("tensor-core evaluation (U=%d I=%d d=%d K=%d)" % (U, I, d, K))
        assert int(err.item()) == 0
        assert torch.equal(pos_e, pos_t) and torch.equal(ids_e, ids_t) and torch.equal(sc_e.view(torch.int32), sc_t.view(torch.int32))
        for k in (0, 3, U - 1):
            p, _, _, _ = O.eval_fullrank_user(P, Q, k, int(test[k]), train[k], I, 1)
            assert p == int(pos_t[k])
