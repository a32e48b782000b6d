"""The C/OpenMP restatement (oracle/apr_oracle_c.c) agrees with the NumPy oracle."""
import numpy as np
import pytest

from oracle import apr_oracle as O
from oracle import c_oracle as C


@pytest.mark.parametrize("adver,reg", [(0, 0.0), (1, 0.0), (1, 0.05)])
def test_c_step_equals_numpy_oracle(adver, reg):
    rng = np.random.RandomState(4)
    U, I, d, B = 90, 70, 24, 300
    P = (rng.randn(U, d) * 0.3).astype(np.float32)
    Q = (rng.randn(I, d) * 0.3).astype(np.float32)
    u, i, j = rng.randint(0, U, B), rng.randint(0, I, B), rng.randint(0, I, B)
    aP, aQ = np.full_like(P, 0.1), np.full_like(Q, 0.1)
    P2, Q2, aP2, aQ2 = P.copy(), Q.copy(), aP.copy(), aQ.copy()
    for _ in range(3):
        O.apr_step(P, Q, aP, aQ, u, i, j, 0.05, reg, 1.0, 0.5, adver)
        C.step(P2, Q2, aP2, aQ2, u, i, j, 0.05, reg, 1.0, 0.5, adver)
    for a, b in ((P, P2), (Q, Q2), (aP, aP2), (aQ, aQ2)):
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()


def test_c_positions_equal_numpy_oracle():
    rng = np.random.RandomState(6)
    U, I, d = 25, 300, 16
    P, Q = rng.randn(U, d).astype(np.float32), rng.randn(I + 1, d).astype(np.float32)
    Q[40] = Q[7]
    train = [sorted(set(rng.randint(0, I, 9).tolist())) for _ in range(U)]
    test = rng.randint(0, I + 1, U)
    test[0] = 7
    ptr, idx = O.build_csr([train[k] + [int(test[k])] for k in range(U)])
    got = C.positions(P, Q, np.arange(U), test, I, ptr, idx)
    for k in range(U):
        p, _, _, _ = O.eval_fullrank_user(P, Q, k, int(test[k]), train[k], I, 1)
        assert p == got[k]
    assert C.threads() >= 1
