"""Pin the CPU oracle against the reference's own test-suite (test/runtests.jl:57-284): every
executed @test is replayed with Julia's default isapprox (norm-wise rtol = sqrt(eps))."""
import numpy as np
import pytest

from oracle import analytic as P
from oracle import sla_oracle as O

RTOL = np.sqrt(np.finfo(np.float64).eps)


def approx(x, y):
    x, y = np.asarray(x), np.asarray(y)
    return np.linalg.norm(x - y) <= RTOL * max(np.linalg.norm(x), np.linalg.norm(y))


class Setup:
    def __init__(self, t):
        # runtests.jl:59-100
        self.Lx, self.t, self.mu, self.beta, self.dtau = 4, t, 0.0, 40.0, 0.1
        self.Ltau = round(self.beta / self.dtau)
        self.ns = 10
        self.Ns = self.Ltau // self.ns
        self.eps, self.U, self.H = P.hamiltonian(self.Lx, t, self.mu)
        self.N = self.H.shape[0]
        self.B, _, _ = P.propagator(self.dtau, self.eps, self.U)
        self.Bi, _, _ = P.propagator(-self.dtau, self.eps, self.U)
        self.G0, self.sgn0, self.logdet0 = P.greens(0.0, self.beta, self.eps, self.U)
        self.Gh, self.sgnh, self.logdeth = P.greens(self.beta / 2, self.beta, self.eps, self.U)
        Bbar = np.eye(self.N, dtype=self.B.dtype)
        Bbari = np.eye(self.N, dtype=self.B.dtype)
        for _ in range(self.ns):
            Bbar = self.B @ Bbar
            Bbari = Bbari @ self.Bi
        self.Bbar, self.Bbari = np.asfortranarray(Bbar), np.asfortranarray(Bbari)
        self.ws = O.ldr_workspace(self.Bbar)


@pytest.fixture(scope="module", params=["real", "complex"])
def S(request):
    # the reference only tests t = 1.0 (runtests.jl:61); the complex-hopping variant pins ComplexF64
    return Setup(1.0 if request.param == "real" else np.exp(0.3j))


def test_golden_scalars():
    # SURVEY.md section 4 / BASELINE.md section 1: closed-form scalars of the real test setup
    s = Setup(1.0)
    assert abs(s.logdet0 - (-484.158883083360)) < 1e-9
    assert abs(s.logdeth - (-484.158883083360)) < 1e-9
    assert abs(s.G0[0, 0] - 0.5) < 1e-12
    assert abs(s.G0[0, 1] - (-0.1875)) < 1e-12
    assert abs(np.max(np.abs(s.Gh)) - 0.1875) < 1e-12
    assert sorted(np.round(s.eps).astype(int).tolist()) == [-4] + [-2] * 4 + [0] * 6 + [2] * 4 + [4]


def test_ldr_logabsdet(S):
    # runtests.jl:105-112 (upstream forgot the @test; asserted here)
    rng = np.random.default_rng(1)
    A = np.eye(S.N) + 0.1 * rng.random((S.N, S.N))
    A = A.astype(S.B.dtype)
    F = O.ldr(A, S.ws)
    logdetF, sgnF = O.logabsdet(F, S.ws)
    sgnA, logdetA = np.linalg.slogdet(A)
    assert abs(logdetF - logdetA) < 1e-10
    assert abs(sgnF - sgnA) < 1e-10


def test_ldr_roundtrips(S):
    A = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)                       # :115-117
    O.copyto_(A, F, S.ws)
    assert approx(A, np.eye(S.N))
    F = O.ldr(S.Bbar, S.ws)                 # :120-122
    O.copyto_(A, F, S.ws)
    assert approx(A, S.Bbar)
    O.ldr_(F, S.Bbar, S.ws)                 # :125-127
    O.copyto_(A, F, S.ws)
    assert approx(A, S.Bbar)
    for Fi in O.ldrs(S.Bbar, 3):            # :130-135
        O.copyto_(A, Fi, S.ws)
        assert approx(A, np.eye(S.N))
    for Fi in O.ldrs(S.Bbar, 3, S.ws):      # :138-143
        O.copyto_(A, Fi, S.ws)
        assert approx(A, S.Bbar)
    # 3-D ldrs / ldrs! (LDR.jl:232-257; never exercised upstream)
    A3 = np.stack([S.Bbar, S.Bbari, S.B], axis=2)
    Fs = O.ldrs(A3, S.ws)
    O.ldrs_(Fs, A3, S.ws)
    for i, Fi in enumerate(Fs):
        O.copyto_(A, Fi, S.ws)
        assert approx(A, A3[:, :, i])


def test_adjoint(S):
    rng = np.random.default_rng(2)          # :146-149
    A = rng.standard_normal((S.N, S.N)).astype(S.B.dtype)
    if S.B.dtype.kind == "c":
        A = A + 1j * rng.standard_normal((S.N, S.N))
    F = O.ldr(A, S.ws)
    G = np.zeros_like(A)
    O.adjoint_(G, F, S.ws)
    assert approx(G, A.conj().T)


def _check_G0(S, G, logdet, sgn):
    assert approx(G, S.G0)
    assert approx(sgn, S.sgn0)
    assert approx(logdet, S.logdet0)


def test_chain_lmul_matrix(S):              # :152-160
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    for _ in range(S.Ns):
        O.lmul_(S.Bbar, F, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))
    assert F.d.min() < 1e-60 and F.d.max() > 1e60   # the product really is ill-conditioned


def test_chain_lmul_ldr(S):                 # :163-171
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    FB = O.ldr(S.Bbar, S.ws)
    for _ in range(S.Ns):
        O.lmul_(FB, F, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_chain_rmul_matrix(S):              # :174-181
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    for _ in range(S.Ns):
        O.rmul_(F, S.Bbar, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_chain_rmul_ldr(S):                 # :184-192
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    FB = O.ldr(S.Bbar, S.ws)
    for _ in range(S.Ns):
        O.rmul_(F, FB, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_chain_ldiv_ldr(S):                 # :195-203
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    FBi = O.ldr(S.Bbari, S.ws)
    for _ in range(S.Ns):
        O.ldiv_(FBi, F, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_chain_ldiv_matrix(S):              # :206-213
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    for _ in range(S.Ns):
        O.ldiv_(S.Bbari, F, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_ldiv_identity(S):                  # :216-219
    FB = O.ldr(S.Bbar, S.ws)
    A = S.Bbar.copy(order="F")
    O.ldiv_(FB, A, S.ws)
    assert approx(A, np.eye(S.N))


def test_chain_rdiv_ldr(S):                 # :222-230
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    FBi = O.ldr(S.Bbari, S.ws)
    for _ in range(S.Ns):
        O.rdiv_(F, FBi, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_chain_rdiv_matrix(S):              # :233-240
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    for _ in range(S.Ns):
        O.rdiv_(F, S.Bbari, S.ws)
    _check_G0(S, G, *O.inv_IpA_(G, F, S.ws))


def test_rdiv_identity(S):                  # :243-246
    FB = O.ldr(S.Bbar, S.ws)
    A = S.Bbar.copy(order="F")
    O.rdiv_(A, FB, S.ws)
    assert approx(A, np.eye(S.N))


def test_inv_IpUV(S):                       # :249-261
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    Fp = O.ldr(S.Bbar)
    for _ in range(S.Ns // 2):
        O.rmul_(F, S.Bbar, S.ws)
    for _ in range(S.Ns // 2, S.Ns):
        O.rmul_(Fp, S.Bbar, S.ws)
    _check_G0(S, G, *O.inv_IpUV_(G, F, Fp, S.ws))


def test_inv_UpV(S):                        # :264-273
    G = np.zeros_like(S.Bbar)
    Fp = O.ldr(S.Bbari)
    F = O.ldr(S.Bbar)
    for _ in range(S.Ns // 2):
        O.lmul_(S.Bbar, F, S.ws)
        O.rmul_(Fp, S.Bbari, S.ws)
    logdet, sgn = O.inv_UpV_(G, Fp, F, S.ws)
    assert approx(G, S.Gh) and approx(sgn, S.sgnh) and approx(logdet, S.logdeth)


def test_inv_invUpV(S):                     # :276-283 (same object passed as U and V)
    G = np.zeros_like(S.Bbar)
    F = O.ldr(S.Bbar)
    for _ in range(S.Ns // 2):
        O.lmul_(S.Bbar, F, S.ws)
    logdet, sgn = O.inv_invUpV_(G, F, F, S.ws)
    assert approx(G, S.Gh) and approx(sgn, S.sgnh) and approx(logdet, S.logdeth)


def test_mul_wrappers(S):
    # overloaded_functions.jl:314-376 and :97,:205 -- never exercised upstream
    rng = np.random.default_rng(3)
    X = rng.standard_normal((S.N, S.N)).astype(S.B.dtype)
    FB = O.ldr(S.Bbar, S.ws)
    H = np.zeros_like(X)
    O.mul_(H, FB, X, S.ws)
    assert approx(H, S.Bbar @ X)
    O.mul_(H, X, FB, S.ws)
    assert approx(H, X @ S.Bbar)
    A = np.zeros_like(X)
    for (U, V, ref) in ((X, FB, X @ S.Bbar), (FB, X, S.Bbar @ X), (FB, FB, S.Bbar @ S.Bbar)):
        HF = O.ldr(S.Bbar)
        O.mul_(HF, U, V, S.ws)
        O.copyto_(A, HF, S.ws)
        assert approx(A, ref)


def test_sweep_oracle_self_consistency():
    """oracle/sweep.py: without field updates G(beta) = G(0), every re-stabilised G(tau_g) equals the
    direct evaluation (I + B(tau_g,0) B(beta,tau_g))^-1, log|det| and sign do not depend on tau, wrap errors are noise."""
    import sla_b200 as sla
    from oracle import chain as ochain
    from oracle import sweep as osweep
    cfg = sla.synthetic.HubbardConfig("t", 4, 40, 5, 1)
    expK = cfg.expK()
    expKinv = np.linalg.inv(expK)
    f = cfg.fields([3])[0]
    r = osweep.sweep(expK, expKinv, f, cfg.lam, cfg.n_stab)
    G_direct, ld, sg, _ = ochain.eval_greens(expK, f, cfg.lam, cfg.n_stab, "IpA")
    assert np.max(np.abs(r["G0"] - G_direct)) < 1e-12
    assert np.max(np.abs(r["G"] - r["G0"])) < 1e-11
    assert np.allclose(r["logdet"], ld, rtol=0, atol=1e-9) and abs(r["logdet0"] - ld) < 1e-10
    assert np.all(r["sign"] == sg) and r["sign0"] == sg
    assert np.all(r["dG"] < 1e-10)
    # G(tau_2) directly: cyclic shift of the slices by two groups
    shift = 2 * cfg.n_stab
    G2, _, _, _ = ochain.eval_greens(expK, np.concatenate([f[shift:], f[:shift]]), cfg.lam, cfg.n_stab, "IpA")
    r2 = osweep.sweep(expK, expKinv, np.concatenate([f[shift:], f[:shift]]), cfg.lam, cfg.n_stab)
    assert np.max(np.abs(r2["G0"] - G2)) < 1e-12
