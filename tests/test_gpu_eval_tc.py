"""tcgen05 (tensor-core) full-rank evaluation: positions must equal the exact fp32 kernel bit for bit, which in turn
equals the oracle (tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

from oracle import apr_oracle as O

pytestmark = pytest.mark.gpu


def _case(rng, U, I, d, scale):
    P = (rng.randn(U + 1, d) * scale).astype(np.float32)
    Q = (rng.randn(I + 1, d) * scale).astype(np.float32)
    Q[I // 2] = Q[1]
    Q[I // 3] = Q[1]
    train = [sorted(set(rng.randint(0, I, rng.randint(0, 30)).tolist())) for _ in range(U)]
    test = rng.randint(0, I + 1, U).astype(np.int32)
    test[0], test[1] = 1, I // 2
    return P, Q, train, test


@pytest.mark.timeout(300)
@pytest.mark.parametrize("U,I,d,scale", [(130, 1000, 64, 1.0), (300, 5000, 128, 1.0), (257, 3333, 128, 0.01),
                                         (100, 2000, 256, 0.3), (64, 700, 8, 1.0), (50, 300, 200, 1.0)])
def test_tc_positions_equal_exact(cuda_device, U, I, d, scale):
    from apr_b200 import engine
    from apr_b200.Dataset import build_sorted_csr
    rng = np.random.RandomState(U + I + d)
    P, Q, train, test = _case(rng, U, I, d, scale)
    ptr, idx = build_sorted_csr([train[u] + [int(test[u])] for u in range(U)])
    dev = cuda_device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    args = [t(P, torch.float32), t(Q, torch.float32), t(np.arange(U, dtype=np.int32), torch.int32), t(test, torch.int32), 0, I,
            t(ptr, torch.int64), t(idx, torch.int32)]
    exact, _, _ = engine.eval_fullrank(*args, 0, exact=True)
    got, n_amb = engine.eval_fullrank_tc(*args)
    assert torch.equal(got, exact), (got - exact).abs().max().item()
    assert 0 <= n_amb <= U * max(256, I // 256)
    # and a few users straight against the oracle
    for u in range(0, U, max(1, U // 7)):
        p, _, _, _ = O.eval_fullrank_user(P, Q, u, int(test[u]), train[u], I, 1)
        assert p == int(got[u])
    # item-sharded use: two ranges accumulate to the same counts
    acc = torch.zeros(U, dtype=torch.int32, device=dev)
    mid = (I // 2) // 128 * 128 + 5
    a2 = list(args)
    a2[4], a2[5] = 0, mid
    engine.eval_fullrank_tc(*a2, position=acc)
    a2[4], a2[5] = mid, I
    engine.eval_fullrank_tc(*a2, position=acc)
    assert torch.equal(acc, exact)
