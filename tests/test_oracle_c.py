"""The C restatement (used for CPU-baseline timing) must agree with the pinned Python oracle."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import chain as ochain


@pytest.mark.parametrize("cplx,inverse", [(False, "IpA"), (True, "IpUV"), (True, "IpA"), (False, "IpUV")])
def test_c_oracle_matches_python_oracle(sla, cplx, inverse):
    syn = sla.synthetic
    cfg = syn.HubbardConfig("t", 6, 60, 10, 3, dtype=np.complex128 if cplx else np.float64, phase=0.3)
    expK, fields = cfg.expK(), cfg.fields(range(3))
    G, ld, sg = c_oracle.eval_batch(expK, fields, cfg.lam, cfg.n_stab, inverse, nthreads=2)
    for c in range(3):
        Gr, lr, sr, _ = ochain.eval_greens(expK, fields[c], cfg.lam, cfg.n_stab, inverse)
        assert np.max(np.abs(G[c] - Gr)) / np.max(np.abs(Gr)) < 1e-11
        assert abs(ld[c] - lr) < 1e-10 * max(1, abs(lr))
        assert abs(sg[c] - sr) < 1e-10


def test_fields_are_reproducible(sla):
    syn = sla.synthetic
    a = syn.hs_fields([0, 1, 2, 3], 20, 16)
    b = syn.hs_fields([2, 3], 20, 16)
    assert np.array_equal(a[2:], b) and set(np.unique(a)) == {-1, 1}
    assert abs(a.mean()) < 0.1
