"""Golden fixtures (tests/golden/golden_small.npz, made by tests/golden/make_golden.py): the oracle must
reproduce them on any box (CPU test) and the CUDA path must match them through the C ABI (GPU test)."""
import os

import numpy as np
import pytest

from oracle import chain as ochain

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz"))
CASES = (("hub_real", np.float64, 0.0, "IpA"), ("hub_cplx", np.complex128, 0.3, "IpUV"))


def relerr(x, ref):
    return float(np.max(np.abs(x - ref)) / np.max(np.abs(ref)))


def test_golden_scalars_closed_form():
    assert abs(float(GOLD["analytic_real_logdet0"]) - (-484.158883083360)) < 1e-9
    assert abs(GOLD["analytic_real_G0"][0, 0] - 0.5) < 1e-12 and abs(GOLD["analytic_real_G0"][0, 1] + 0.1875) < 1e-12


@pytest.mark.parametrize("tag,dtype,phase,inv", CASES)
def test_oracle_reproduces_golden(sla, tag, dtype, phase, inv):
    cfg = sla.synthetic.HubbardConfig("g", 4, 40, 10, 3, dtype=dtype, phase=phase)
    expK, fields = cfg.expK(), cfg.fields(range(3))
    for c in range(3):
        G, ld, sg, _ = ochain.eval_greens(expK, fields[c], cfg.lam, cfg.n_stab, inv)
        assert relerr(G, GOLD[f"{tag}_G"][c]) < 1e-11
        assert abs(ld - GOLD[f"{tag}_logdet"][c]) < 1e-10 and abs(sg - GOLD[f"{tag}_sign"][c]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dtype,phase,inv", CASES)
def test_cuda_matches_golden(sla, tag, dtype, phase, inv):
    import torch
    cfg = sla.synthetic.HubbardConfig("g", 4, 40, 10, 3, dtype=dtype, phase=phase)
    G, ld, sg, _ = sla.greens_chain(sla.to_device(cfg.expK()), torch.from_numpy(cfg.fields(range(3))).cuda(), cfg.lam,
                                    cfg.n_stab, inv)
    assert relerr(sla.to_host(G), GOLD[f"{tag}_G"]) < 1e-10
    assert np.allclose(ld.cpu().numpy(), GOLD[f"{tag}_logdet"], atol=1e-10)
    assert np.allclose(sg.cpu().numpy(), GOLD[f"{tag}_sign"], atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["real", "cplx"])
def test_cuda_matches_closed_form_golden(sla, tag):
    import torch
    Bbar = GOLD[f"analytic_{tag}_Bbar"]
    dB = sla.to_device(Bbar)
    ws = sla.ldr_workspace(dB)
    F = sla.ldr(dB)
    for _ in range(40):
        sla.lmul_(dB, F, ws)
    G = torch.empty_like(dB)
    ld, sg = sla.inv_IpA_(G, F, ws)
    assert relerr(sla.to_host(G)[0], GOLD[f"analytic_{tag}_G0"]) < 1e-9
    assert abs(float(ld[0]) - float(GOLD[f"analytic_{tag}_logdet0"])) < 1e-8 and abs(complex(sg[0]) - 1.0) < 1e-9
