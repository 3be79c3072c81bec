"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): adaptive DOPRI5/DOP853 -- per-trajectory final states within 1e-9 relative and
accepted/rejected counts equal on >= 99.9% of trajectories.  Lorenz on t in [0,100] is chaotic, so that bar is only
reachable bit-exactly: these tests assert BITWISE equality of every output (states, t_eval rows, counters, status).
Fixed-step RK4 and SDE paths: 1e-12 relative (asserted bitwise where the arithmetic is IEEE-exact on both sides).
"""
import importlib

import math

import numpy as np
import pytest

import oracle_binding as ob

deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta
pytestmark = pytest.mark.gpu


def bits(a):
    """Bit patterns, with every NaN mapped to one pattern: which NaN an invalid operation produces is a property of the
    processor (x86 SSE: the negative 'indefinite' 0xFFF8..., the GPU: a positive canonical NaN), not of the algorithm."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = a.view(np.uint64).copy()
    b[np.isnan(a)] = np.uint64(0x7FF8000000000000)
    return b


def assert_same_solution(gpu, cpu, exact=True, rtol=0.0):
    assert np.array_equal(gpu.status, cpu.status), "status differs"
    assert np.array_equal(gpu.accepted, cpu.accepted), "accepted step counts differ"
    assert np.array_equal(gpu.rejected, cpu.rejected), "rejected step counts differ"
    assert np.array_equal(gpu.evals, cpu.evals), "function evaluation counts differ"
    assert np.array_equal(gpu.n_emitted, cpu.n_emitted), "number of emitted t_eval rows differs"
    assert np.array_equal(gpu.t_rows, cpu.t_rows)
    mask = np.arange(gpu.y_eval.shape[1])[None, :] < gpu.n_emitted[:, None]
    assert (gpu.t_out is None) == (cpu.t_out is None)
    if gpu.t_out is not None:
        assert np.array_equal(bits(gpu.t_out)[mask], bits(cpu.t_out)[mask]), "row times differ bitwise"
    if exact:
        assert np.array_equal(bits(gpu.t_final), bits(cpu.t_final)), "t_final differs bitwise"
        assert np.array_equal(bits(gpu.y_final), bits(cpu.y_final)), "y_final differs bitwise"
        assert np.array_equal(bits(gpu.y_eval)[mask], bits(cpu.y_eval)[mask]), "t_eval rows differ bitwise"
    else:
        np.testing.assert_allclose(gpu.t_final, cpu.t_final, rtol=rtol, atol=0)
        np.testing.assert_allclose(gpu.y_final, cpu.y_final, rtol=rtol, atol=1e-300)
        np.testing.assert_allclose(gpu.y_eval[mask], cpu.y_eval[mask], rtol=rtol, atol=1e-300)


def lorenz():
    return deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0)


# ------------------------------------------------------------------------------------------ the controller's pow
@pytest.mark.parametrize("y", [-0.2, -0.125, 0.2, 0.125])
def test_device_pow_is_libm_pow(y):
    """The device port of glibc pow returns libm's bits (the reference's powf, ordinary.rs:154 / h_init.rs:124)."""
    lib = deb.load_library()
    rng = np.random.default_rng(7)
    n = 1 << 22
    x = np.concatenate([
        np.exp(rng.uniform(-27.6, 27.6, n)),                               # [1e-12, 1e12] log-uniform
        rng.integers(1, 0x7FEFFFFFFFFFFFFF, n, dtype=np.int64).view(np.float64),  # every positive finite exponent
        1.0 + rng.uniform(-0.5, 0.5, n // 4) * 2.0 ** -rng.integers(0, 60, n // 4),  # near 1
        rng.integers(1, 1 << 52, 4096, dtype=np.int64).view(np.float64),  # subnormals
        np.array([0.0, np.inf, np.nan, 1.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308]),
    ])
    out = np.empty_like(x)
    rc = lib.deb_pow_device(x.ctypes.data_as(deb._dp), y, x.size, out.ctypes.data_as(deb._dp), 0)
    assert rc == 0, lib.deb_last_error()
    ref = ob.libm_pow(x, y)
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(out), nan)
    assert np.array_equal(bits(out)[~nan], bits(ref)[~nan]), f"{np.count_nonzero(bits(out)[~nan] != bits(ref)[~nan])} mismatches"


# ------------------------------------------------------------------------------------------ config C1
def test_c1_lorenz_dopri5_1024_bit_exact():
    """Config C1: Lorenz DOPRI5 rtol=1e-8 (atol default 1e-6), t in [0,100], 1024 perturbed initial conditions, plus
    the t_eval grid of C2 (100 points, the last equal to tf: exact-hit branch of t_eval.rs:113)."""
    y0 = ob.lorenz_ensemble_y0(1024)
    def prob():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0).t_eval(np.arange(1.0, 101.0)).method(E.dopri5().rtol(1e-8))
    gpu, cpu = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(gpu, cpu)
    assert (gpu.status == deb.DEB_STATUS_COMPLETE).sum() > 900
    # the size-independent book-keeping identity: evals = 3 + 6*attempts + accepted
    assert np.array_equal(gpu.evals, 3 + 6 * (gpu.accepted + gpu.rejected) + gpu.accepted)


def test_lorenz_dopri5_queue_refill_more_trajectories_than_threads():
    """More trajectories than resident threads with short, very unequal lifetimes: exercises the warp-aggregated
    work-queue refill; results must not depend on which lane ran a trajectory."""
    n = 200_000
    y0 = ob.lorenz_ensemble_y0(n, seed=99)
    tf = 0.6
    def prob(sub=slice(None)):
        return deb.EnsembleIVP.ode(lorenz(), 0.0, tf, y0[sub]).t_eval([0.1, 0.3, 0.6]).method(E.dopri5().rtol(1e-8))
    gpu = prob().solve()
    sub = slice(0, n, 37)
    cpu = ob.oracle_solve(prob(sub))
    for name in ("status", "accepted", "rejected", "evals", "n_emitted"):
        assert np.array_equal(getattr(gpu, name)[sub], getattr(cpu, name)), name
    assert np.array_equal(bits(gpu.y_final[sub]), bits(cpu.y_final))
    assert np.array_equal(bits(gpu.y_eval[sub]), bits(cpu.y_eval))
    assert (gpu.status == 0).all()


def test_dopri5_all_builtin_systems_bit_exact():
    rng = np.random.default_rng(3)
    cases = [
        (deb.ExponentialGrowth(1.0), 0.0, 10.0, 1.0 + rng.uniform(-0.1, 0.1, (64, 1))),
        (deb.LinearEquation(1.0, 1.0), 0.0, 10.0, 1.0 + rng.uniform(-0.1, 0.1, (64, 1))),
        (deb.HarmonicOscillator(1.0), 0.0, 10.0, np.array([1.0, 0.0]) + rng.uniform(-0.1, 0.1, (64, 2))),
        (deb.LogisticEquation(1.0, 10.0), 0.0, 10.0, 0.1 + rng.uniform(0, 0.1, (64, 1))),
        (deb.VanDerPolOscillator(rng.uniform(0.1, 5.0, 64)), 0.0, 10.0, np.tile([2.0, 0.0], (64, 1))),
        (deb.BrusselatorSystem(1.0, 3.0), 0.0, 10.0, np.array([1.5, 3.0]) + rng.uniform(-0.1, 0.1, (64, 2))),
    ]
    for sysm, t0, tf, y0 in cases:
        for m in (E.dopri5, E.dop853):
            def prob():
                return deb.EnsembleIVP.ode(sysm, t0, tf, y0).t_eval(np.linspace(t0, tf, 23)).method(m().rtol(1e-7).atol(1e-9))
            assert_same_solution(prob().solve(), ob.oracle_solve(prob()))


# ------------------------------------------------------------------------------------------ config C3 (reduced)
def test_c3_van_der_pol_dop853_sweep_bit_exact():
    """Config C3 at oracle-sized N: mu sweep in [0.1, 50], DOP853 rtol=atol=1e-8, y0=(2,0), t in [0,100]."""
    n = 1536
    mu = 0.1 + 49.9 * np.arange(n) / (n - 1)
    y0 = np.tile([2.0, 0.0], (n, 1))
    def prob():
        return (deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 0.0, 100.0, y0).t_eval(np.arange(5.0, 101.0, 5.0))
                .method(E.dop853().rtol(1e-8).atol(1e-8)))
    gpu, cpu = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(gpu, cpu)
    assert (gpu.status == 0).all()
    # evals = 3 + 11*attempts + 4*accepted for DOP853 (dense stages counted on every accepted step)
    assert np.array_equal(gpu.evals, 3 + 11 * (gpu.accepted + gpu.rejected) + 4 * gpu.accepted)


# ------------------------------------------------------------------------------------------ errors / edge cases
def test_error_statuses_match_reference_semantics():
    y0 = ob.lorenz_ensemble_y0(64)
    # MaxSteps: atol=1e-8 needs > 10000 attempts
    def p1():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0).method(E.dopri5().rtol(1e-8).atol(1e-8))
    g, c = p1().solve(), ob.oracle_solve(p1())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_MAX_STEPS).any()
    with pytest.raises(deb.MaxSteps):
        g[int(np.argmax(g.status == deb.DEB_STATUS_MAX_STEPS))]
    # BadInput: tf == t0 (tests/ode/errors.rs:87) and h0 larger than the interval (errors.rs:110)
    def p2():
        return deb.EnsembleIVP.ode(lorenz(), 1.0, 1.0, y0).method(E.dopri5())
    g, c = p2().solve(), ob.oracle_solve(p2())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_BAD_INPUT).all()
    def p3():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0).method(E.dopri5().h0(2.0))
    g, c = p3().solve(), ob.oracle_solve(p3())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_BAD_INPUT).all()
    with pytest.raises(deb.BadInput):
        g[0]
    # StepSize: a blow-up (y' = y^2-like growth through the logistic equation with negative capacity)
    def p4():
        return deb.EnsembleIVP.ode(deb.LogisticEquation(1.0, -1.0), 0.0, 10.0, np.full((8, 1), 1.0)).method(E.dopri5())
    g, c = p4().solve(), ob.oracle_solve(p4())
    assert_same_solution(g, c)
    assert set(g.status.tolist()) <= {deb.DEB_STATUS_STEP_SIZE, deb.DEB_STATUS_MAX_STEPS}


def test_t_eval_ordering_filtering_and_backward_integration():
    y0 = np.array([1.0, 0.0]) + np.linspace(0, 0.1, 32)[:, None]
    pts = [3.0, 0.0, 11.0, -1.0, 0.5, 10.0, 0.5, 7.25]  # unsorted, duplicates, outside the interval, t0 and tf themselves
    for m in (E.dopri5, E.dop853, lambda: E.rk4(0.01)):
        def fwd():
            return deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 10.0, y0).t_eval(pts).method(m())
        g, c = fwd().solve(), ob.oracle_solve(fwd())
        assert_same_solution(g, c)
        assert g.t_rows.tolist() == [0.5, 0.5, 3.0, 7.25, 10.0, 11.0]  # -1 and 0.0 (not first) are consumed at t0
        def bwd():
            return deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 10.0, 0.0, y0).t_eval(pts).method(m())
        if m in (E.dopri5, E.dop853):  # (backward integration with a fixed step needs a negative h: below)
            g, c = bwd().solve(), ob.oracle_solve(bwd())
            assert_same_solution(g, c)
            assert g.t_rows.tolist() == [7.25, 3.0, 0.5, 0.5, 0.0, -1.0]
    def bwd_rk4():
        return deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 10.0, 0.0, y0).t_eval(pts).method(E.rk4(-0.01))
    assert_same_solution(bwd_rk4().solve(), ob.oracle_solve(bwd_rk4()))
    # first point equal to t0 is emitted by the solout call before the loop
    def t0_first():
        return deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 2.0, y0).t_eval([0.0, 1.0, 2.0]).method(E.dopri5())
    g, c = t0_first().solve(), ob.oracle_solve(t0_first())
    assert_same_solution(g, c)
    assert np.array_equal(g.y_eval[:, 0, :], y0) and (g.n_emitted >= 2).all()


def test_vector_tolerances_and_options_cross_the_abi():
    y0 = ob.lorenz_ensemble_y0(64, seed=5)
    def prob():
        m = (E.dopri5().rtol([1e-7, 1e-8, 1e-6]).atol([1e-9, 1e-6, 1e-7]).h_max(0.05).h_min(1e-9).safety_factor(0.8)
             .min_scale(0.3).max_scale(5.0).max_steps(5000).h0(1e-3))
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 20.0, y0).t_eval(np.linspace(0.5, 20, 40)).method(m)
    assert_same_solution(prob().solve(), ob.oracle_solve(prob()))


# ------------------------------------------------------------------------------------------ fixed step
@pytest.mark.parametrize("ctor", ["euler", "midpoint", "heun", "ralston", "ssp_rk3", "rk4", "three_eighths"])
def test_fixed_step_methods_bit_exact(ctor):
    """Fixed-step ERK (bar: 1e-12; asserted bitwise -- both sides execute the same IEEE operations)."""
    y0 = ob.lorenz_ensemble_y0(300, seed=11)
    def prob():
        return (deb.EnsembleIVP.ode(lorenz(), 0.0, 2.0, y0).t_eval([0.0, 0.333, 1.0, 1.5, 2.0])
                .method(getattr(E, ctor)(1e-3)))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c)
    # t accumulates h (fixed/ordinary.rs:123): 2000 steps, or 2001 with a last sliver step, the same for all lanes
    assert (g.status == 0).all() and g.accepted[0] in (2000, 2001) and (g.accepted == g.accepted[0]).all()


def test_fixed_step_default_h_and_max_steps():
    y0 = np.full((16, 1), 1.0)
    def p1():  # h0 == 0 -> |tf-t0|/100 (fixed/ordinary.rs:23-28)
        return deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, y0).method(E.rk4(0.0))
    g, c = p1().solve(), ob.oracle_solve(p1())
    assert_same_solution(g, c)
    assert (g.accepted == 100).all()
    def p2():  # 20001 steps needed, default max_steps 10000 -> MaxSteps
        return deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, y0).method(E.euler(5e-5))
    g, c = p2().solve(), ob.oracle_solve(p2())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_MAX_STEPS).all() and (g.accepted == 10000).all()


def test_reference_from_fn_euler_kat():
    """tests/ode/from_fn.rs:4-18: Euler h=0.1 on y'=y over [0,1] gives 2.5937 +- 1e-3."""
    s = deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.euler(0.1)).solve()[0]
    assert abs(s.y_final[0] - 2.5937) < 1e-3


# ------------------------------------------------------------------------------------------ SDE (config C4, reduced)
@pytest.mark.parametrize("which", ["ou", "gbm"])
def test_c4_euler_maruyama_matches_host_regenerated_philox(which):
    n = 4096
    if which == "ou":   # theta=0.5, mu=1.0, sigma=0.3, y0=5, t in [0,10], euler(0.01): 1001 steps
        sysm, t0, tf, h, y0 = deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), 0.0, 10.0, 0.01, 5.0
        nsteps = 1001
    else:               # mu=0.1, sigma=0.2, y0=100, t in [0,1], euler(1e-3): 1000 steps
        sysm, t0, tf, h, y0 = deb.GeometricBrownianMotion(0.1, 0.2), 0.0, 1.0, 1e-3, 100.0
        nsteps = 1000
    def prob():
        return (deb.EnsembleIVP.sde(sysm, t0, tf, np.full(n, y0), seed=2026, path_offset=123456789)
                .t_eval([tf / 3, tf / 2, tf]).method(E.euler(h)))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c, exact=False, rtol=1e-12)
    assert (g.accepted == nsteps).all() and (g.status == 0).all()
    # paths differ from each other (the noise is per path) and have the right spread
    assert np.unique(g.y_final).size == n


def test_sde_rk4_drift_stages():
    """examples/sde/03_ornstein_uhlenbeck/main.rs:66 drives the SDE with rk4(dt): RK drift stages + EM diffusion."""
    n = 512
    def prob():
        return (deb.EnsembleIVP.sde(deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), 0.0, 10.0, np.full(n, 5.0), seed=42)
                .t_eval([2.0, 5.0, 8.0]).method(E.rk4(0.01)))
    assert_same_solution(prob().solve(), ob.oracle_solve(prob()), exact=False, rtol=1e-12)


def test_sde_step_schedule_edge_cases():
    """The SDE step schedule is planned on the host (it does not depend on the path): MaxSteps part-way (rows up to the
    failure), BadInput (h0 larger than the interval / wrong sign), the default h0 = |tf - t0| / 100, a clipped last step,
    backward time, t_eval points on and between step ends, duplicated and out-of-range points."""
    n = 256
    ou = deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3)
    cases = [
        (0.0, 1.0, E.euler(0.01).max_steps(37), [0.1, 0.2, 0.5, 0.9]),        # MaxSteps at t = 0.37
        (0.0, 1.0, E.euler(2.0), [0.5]),                                       # BadInput: h0 > interval
        (0.0, 1.0, E.euler(-0.01), [0.5]),                                     # BadInput: wrong direction
        (0.0, 1.0, E.euler(0.0), [0.005, 0.01, 0.5, 1.0]),                     # default h0
        (0.0, 1.0, E.rk4(0.03), [0.0, 0.03, 0.5, 0.99, 1.0, 1.5, -1.0, 0.5]),  # clipped last step, duplicates, out of range
        (2.0, -1.0, E.heun(-0.07), [1.93, 0.0, -1.0, 1.0]),                    # backward
        (0.0, 1.0, deb.Milstein.new(0.013), [0.4, 1.0]),
        (0.0, 1.0, E.euler(0.25), []),                                         # no rows
    ]
    for t0, tf, meth, te in cases:
        def prob():
            return deb.EnsembleIVP.sde(ou, t0, tf, np.linspace(0.5, 5.0, n), seed=11, path_offset=5).t_eval(te).method(meth)
        g, c = prob().solve(), ob.oracle_solve(prob())
        assert_same_solution(g, c, exact=False, rtol=1e-12)


def test_heston_vector_sde_matches_host_regenerated_philox():
    """examples/sde/02_heston_model: two-dimensional state, diagonal noise with correlated increments (rho), three_eighths(dt)
    drift stages, Euler-Maruyama and Milstein; a sweep over rho; sample moments of the price."""
    n = 8192
    y0 = np.tile([100.0, 0.04], (n, 1))
    for meth in (lambda: E.three_eighths(0.01), lambda: E.euler(0.002), lambda: deb.Milstein.new(0.005)):
        def prob():
            return (deb.EnsembleIVP.sde(deb.HestonModel(0.1, 2.0, 0.04, 0.3, -0.7), 0.0, 1.0, y0, seed=42, path_offset=77)
                    .t_eval([0.25, 0.5, 1.0]).method(meth()))
        g, c = prob().solve(), ob.oracle_solve(prob())
        assert g.y_eval.shape == (n, 3, 2) and (g.status == 0).all()
        assert_same_solution(g, c, exact=False, rtol=1e-12)
        ok = np.isfinite(g.y_final).all(axis=1)   # a variance that dips below zero gives sqrt(v) = NaN, as in the reference
        assert ok.mean() > 0.95
        assert abs(g.y_final[ok, 0].mean() / (100.0 * math.exp(0.1)) - 1.0) < 0.02   # E[S_T] = S_0 e^{mu T}
        assert abs(g.y_final[ok, 1].mean() - 0.04) < 0.004                            # v starts at its long-run mean theta
    rho = np.linspace(-0.9, 0.9, n)
    prm = np.stack([np.full(n, 0.1), np.full(n, 2.0), np.full(n, 0.04), np.full(n, 0.3), rho], axis=1)
    def p2():
        return deb.EnsembleIVP.sde(deb.SdeSystem(deb.DEB_SDE_HESTON, prm, 2), 0.0, 0.5, y0, seed=3).method(E.three_eighths(0.01))
    assert_same_solution(p2().solve(), ob.oracle_solve(p2()), exact=False, rtol=1e-12)


# ------------------------------------------------------------------------------------------ heat (config C5, reduced)
@pytest.mark.parametrize("n,lo,hi", [(4097, 0.0, 4096.0), (4096, 0.0, 1.0), (41, 0.0, 1.0), (2, 0.0, 1.0), (65537, 0.0, 65536.0)])
@pytest.mark.parametrize("bc", ["dirichlet", "neumann", "mixed"])
def test_c5_heat_method_of_lines_rk4(n, lo, hi, bc):
    """dx = 1 (power of two: multiplication path) and dx = 1/4095, 1/40 (IEEE division path); odd and even N."""
    x = np.linspace(lo, hi, n)
    u0 = np.sin(np.pi * (x - lo) / (hi - lo)) + 0.25 * np.cos(3 * np.pi * (x - lo) / (hi - lo))
    dx = (hi - lo) / (n - 1)
    alpha = 0.1
    h = 0.2 * dx * dx / alpha
    bcl, bcu = {"dirichlet": (("dirichlet", 0.0), ("dirichlet", 0.0)), "neumann": (("neumann", 0.3), ("neumann", -0.2)),
                "mixed": (("dirichlet", 0.0), ("neumann", 0.1))}[bc]
    for meth in (E.rk4(h), E.three_eighths(h), E.euler(h / 4)):
        tf = 37.5 * h
        g = deb.solve_heat_mol(u0, lo, hi, alpha, meth, 0.0, tf, bcl, bcu)
        c = ob.oracle_heat(u0, lo, hi, alpha, meth, 0.0, tf, bcl, bcu)
        assert g.status == c.status == "Complete" and g.steps == c.steps and g.t == c.t
        assert np.array_equal(bits(g.u), bits(c.u)), f"max diff {np.abs(g.u - c.u).max()}"


def test_heat_rhs_reference_kats():
    """tests/pde/method_of_lines.rs:73-91 (Dirichlet rows exactly 0) and :143-157 (Neumann KAT)."""
    du = deb.heat_rhs([2.0, 1.0, 0.0, -0.5, -1.0], 0.0, 1.0, 1.0, ("dirichlet", 2.0), ("dirichlet", -1.0))
    assert du[0] == 0.0 and du[-1] == 0.0
    du = deb.heat_rhs([0.0, 1.0, 0.0, -1.0, 0.0], 0.0, 1.0, 1.0)
    assert du[0] == 0.0 and du[4] == 0.0 and abs(du[2]) < 1e-12
    dx = 0.25
    du = deb.heat_rhs([1.0, 2.0, 2.0, 2.0, 2.0], 0.0, 1.0, 1.0, ("neumann", 0.0), ("neumann", 0.0))
    assert abs(du[0] - (1.0 / dx) / dx) < 1e-12
    rng = np.random.default_rng(0)
    for n in (2, 3, 64, 65, 1000, 4099):
        u = rng.normal(size=n)
        for bcl, bcu in ((("dirichlet", 0.0), ("neumann", 0.5)), (("neumann", -1.0), ("dirichlet", 0.0))):
            assert np.array_equal(bits(deb.heat_rhs(u, -1.0, 2.0, 0.7, bcl, bcu)), bits(ob.oracle_heat_rhs(u, -1.0, 2.0, 0.7, bcl, bcu)))


def test_heat_matches_analytic_mode():
    """tests/pde/method_of_lines.rs:37-70: N=41, alpha=0.1, rk4(1e-4), t=0.02, tolerance 5e-4 against the analytic mode."""
    x = np.linspace(0.0, 1.0, 41)
    g = deb.solve_heat_mol(np.sin(np.pi * x), 0.0, 1.0, 0.1, E.rk4(1.0e-4), 0.0, 0.02)
    expected = np.exp(-0.1 * np.pi ** 2 * 0.02) * np.sin(np.pi * x)
    assert np.abs(g.u - expected).max() < 5.0e-4


# ------------------------------------------------------------------------------------------ ensemble statistics
def test_ensemble_stats_match_numpy():
    import ctypes as C
    lib = deb.load_library()
    y0 = ob.lorenz_ensemble_y0(5000, seed=1)
    c = ob.oracle_solve(deb.EnsembleIVP.ode(lorenz(), 0.0, 5.0, y0[:256]).method(E.dopri5().rtol(1e-8)))
    g_max_steps = int(np.median(c.accepted + c.rejected))  # about half of the trajectories run out of steps
    g = deb.EnsembleIVP.ode(lorenz(), 0.0, 5.0, y0).t_eval(np.linspace(0.5, 5.0, 10)).method(E.dopri5().rtol(1e-8).max_steps(g_max_steps)).solve()
    assert (g.status != 0).any() and (g.status == 0).any()  # some trajectories stop early: ragged n_emitted
    sums = np.zeros((10, 3, 2))
    counts = np.zeros(10, np.int64)
    rc = lib.deb_ensemble_stats(g.y_eval.ctypes.data, g.n_emitted.ctypes.data, 5000, 10, 3, sums.ctypes.data, counts.ctypes.data, 0,
                                deb.DEB_MEM_HOST, None)
    assert rc == 0, lib.deb_last_error()
    mask = np.arange(10)[None, :] < g.n_emitted[:, None]
    assert np.array_equal(counts, mask.sum(0))
    ye = np.where(mask[:, :, None], g.y_eval, 0.0)
    np.testing.assert_allclose(sums[:, :, 0], ye.sum(0), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sums[:, :, 1], (ye * ye).sum(0), rtol=1e-12)


# ------------------------------------------------------------------------------------------ properties at scale
def test_large_ensemble_properties_and_subset_parity():
    """1M Lorenz trajectories (config C2 at 1/10 size): size-independent properties + a strided subset against the oracle."""
    n = 1_000_000
    y0 = ob.lorenz_ensemble_y0(n)
    te = np.arange(1.0, 101.0)
    g = deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0).t_eval(te).method(E.dopri5().rtol(1e-8)).solve()
    ok = g.status == 0
    assert ok.mean() > 0.9
    assert np.array_equal(g.evals, 3 + 6 * (g.accepted + g.rejected) + g.accepted)
    assert (g.accepted + g.rejected <= 10000).all()
    assert ((g.status == deb.DEB_STATUS_MAX_STEPS) == (g.accepted + g.rejected == 10000) | (g.status == deb.DEB_STATUS_MAX_STEPS)).all()
    assert (g.n_emitted[ok] == 100).all() and np.array_equal(bits(g.y_eval[ok, 99]), bits(g.y_final[ok]))  # exact hit at tf
    assert np.isfinite(g.y_eval[ok]).all() and np.abs(g.y_eval[ok]).max() < 100.0  # on the attractor
    # the first 1024 trajectories are config C1: identical results whatever the ensemble size (no cross-talk)
    g1 = deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0[:1024]).t_eval(te).method(E.dopri5().rtol(1e-8)).solve()
    assert np.array_equal(bits(g.y_eval[:1024]), bits(g1.y_eval)) and np.array_equal(g.accepted[:1024], g1.accepted)
    sub = slice(0, n, 4001)
    c = ob.oracle_solve(deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0[sub]).t_eval(te).method(E.dopri5().rtol(1e-8)))
    assert np.array_equal(g.status[sub], c.status) and np.array_equal(g.accepted[sub], c.accepted)
    assert np.array_equal(g.rejected[sub], c.rejected)
    assert np.array_equal(bits(g.y_final[sub]), bits(c.y_final))
    m = np.arange(100)[None, :] < c.n_emitted[:, None]
    assert np.array_equal(bits(g.y_eval[sub])[m], bits(c.y_eval)[m])


def test_host_path_chunk_pipeline_is_transparent(monkeypatch):
    """The HOST-memspace call pipelines the ensemble in chunks over two streams; chunking must not change a bit,
    including with per-trajectory parameters, a ragged last chunk and rows that stop early."""
    n = 5003
    mu = np.linspace(0.1, 8.0, n)
    y0 = np.tile([2.0, 0.0], (n, 1)) + ob.splitmix64_uniform(3, 2 * n).reshape(n, 2) * 0.1
    def prob():
        return (deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 0.0, 12.0, y0).t_eval(np.linspace(0.0, 12.0, 7))
                .method(E.dopri5().rtol(1e-7).max_steps(158)))
    monkeypatch.setenv("DEB_HOST_CHUNK", "100000000")
    one = prob().solve()
    monkeypatch.setenv("DEB_HOST_CHUNK", "700")
    monkeypatch.setenv("DEB_WM_SHIFT", "6")  # chunks are whole watermark blocks: 64-trajectory blocks -> chunks of 704
    many = prob().solve()
    assert (one.status != 0).any() and (one.status == 0).any()
    for name in ("status", "accepted", "rejected", "evals", "n_emitted"):
        assert np.array_equal(getattr(one, name), getattr(many, name)), name
    assert np.array_equal(bits(one.y_final), bits(many.y_final)) and np.array_equal(bits(one.t_final), bits(many.t_final))
    m = np.arange(7)[None, :] < one.n_emitted[:, None]
    assert np.array_equal(bits(one.y_eval)[m], bits(many.y_eval)[m])
    assert_same_solution(many, ob.oracle_solve(prob()))


# ------------------------------------------------------------------------------------------ user-defined RHS (NVRTC)
LORENZ_SRC = """
    const double x = y[0], yv = y[1], z = y[2];
    dydt[0] = p[0] * (yv - x);
    dydt[1] = x * (p[1] - z) - yv;
    dydt[2] = x * yv - p[2] * z;
"""
# tests/ode/systems.rs:122-158 (powi(2) = x*x, powi(3) = x*x*x)
CR3BP_SRC = """
    const double mu = p[0];
    const double rx = y[0], ry = y[1], rz = y[2], vx = y[3], vy = y[4], vz = y[5];
    const double a1 = rx + mu, a2 = rx - 1.0 + mu;
    const double r13 = sqrt(a1 * a1 + ry * ry + rz * rz);
    const double r23 = sqrt(a2 * a2 + ry * ry + rz * rz);
    const double r13c = r13 * r13 * r13, r23c = r23 * r23 * r23;
    dydt[0] = vx;
    dydt[1] = vy;
    dydt[2] = vz;
    dydt[3] = rx + 2.0 * vy - (1.0 - mu) * (rx + mu) / r13c - mu * (rx - 1.0 + mu) / r23c;
    dydt[4] = ry - 2.0 * vx - (1.0 - mu) * ry / r13c - mu * ry / r23c;
    dydt[5] = -(1.0 - mu) * rz / r13c - mu * rz / r23c;
"""


def cr3bp_py(mu):
    import math
    def f(t, y):
        rx, ry, rz, vx, vy, vz = y
        a1, a2 = rx + mu, rx - 1.0 + mu
        r13 = math.sqrt(a1 * a1 + ry * ry + rz * rz)
        r23 = math.sqrt(a2 * a2 + ry * ry + rz * rz)
        r13c, r23c = r13 * r13 * r13, r23 * r23 * r23
        return [vx, vy, vz,
                rx + 2.0 * vy - (1.0 - mu) * (rx + mu) / r13c - mu * (rx - 1.0 + mu) / r23c,
                ry - 2.0 * vx - (1.0 - mu) * ry / r13c - mu * ry / r23c,
                -(1.0 - mu) * rz / r13c - mu * rz / r23c]
    return f


def test_user_defined_rhs_equals_builtin_bitwise():
    """deb_define_ode: the Lorenz system written as source text goes through the same kernel template (compiled by NVRTC)
    and must reproduce the built-in system -- hence the oracle -- bit for bit, for adaptive and fixed-step methods,
    shared and per-trajectory parameters."""
    y0 = ob.lorenz_ensemble_y0(3000, seed=21)
    te = np.linspace(0.0, 8.0, 17)
    user = deb.ode_from_source(3, LORENZ_SRC, [10.0, 28.0, 8.0 / 3.0])
    for m in (lambda: E.dopri5().rtol(1e-8), lambda: E.dop853().rtol(1e-9).atol(1e-9), lambda: E.rk4(0.005)):
        u = deb.EnsembleIVP.ode(user, 0.0, 8.0, y0).t_eval(te).method(m()).solve()
        b = deb.EnsembleIVP.ode(lorenz(), 0.0, 8.0, y0).t_eval(te).method(m()).solve()
        assert_same_solution(u, b)
    rho = np.linspace(20.0, 35.0, 3000)
    prm = np.stack([np.full(3000, 10.0), rho, np.full(3000, 8.0 / 3.0)], axis=1)
    u = deb.EnsembleIVP.ode(deb.ode_from_source(3, LORENZ_SRC, prm), 0.0, 5.0, y0).t_eval(te).method(E.dopri5().rtol(1e-8)).solve()
    c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, rho, 8.0 / 3.0), 0.0, 5.0, y0).t_eval(te).method(E.dopri5().rtol(1e-8)))
    assert_same_solution(u, c)


def test_user_defined_six_dimensional_system_vs_python_restatement():
    """A system that is NOT built in (CR3BP, dim 6, tests/ode/systems.rs:122-158), against the independent pure-Python
    restatement of the DOPRI5 / DOP853 loops: bitwise states, dense rows and counters."""
    import py_restatement as pr
    mu = 0.012150585609624
    y0 = np.array([1.021881345465263, 0.0, -0.182000000000000, 0.0, -0.102950816739606, 0.0])
    y0s = y0[None, :] + ob.splitmix64_uniform(8, 6 * 40).reshape(40, 6) * 1e-3
    te = [0.5, 1.0, 1.5110806241094467]
    user = deb.ode_from_source(6, CR3BP_SRC, [mu])
    for meth in ("dopri5", "dop853"):
        g = deb.EnsembleIVP.ode(user, 0.0, 1.5110806241094467, y0s).t_eval(te).method(getattr(E, meth)().rtol(1e-9).atol(1e-9)).solve()
        assert (g.status == 0).all()
        for i in (0, 7, 39):
            p = pr.solve_dp(cr3bp_py(mu), meth, 0.0, 1.5110806241094467, list(y0s[i]), rtol=1e-9, atol=1e-9, t_eval=te)
            assert (p["accepted"], p["rejected"], p["evals"]) == (int(g.accepted[i]), int(g.rejected[i]), int(g.evals[i]))
            assert np.array_equal(bits(p["y"]), bits(g.y_final[i]))
            assert np.array_equal(bits([r[1] for r in p["rows"]]), bits(g.y_eval[i, :len(p["rows"])]))


def test_user_defined_wide_system_forward_sensitivities():
    """A 12-dimensional user system: Lorenz with its forward sensitivities dy/dp (the augmented system of
    src/ode/sensitivity/forward.rs: S' = J_y S + J_p, row-major S after y).  Wide systems do not park their t_eval rows
    (the stash would not fit shared memory), so this covers the immediate-emission path of DOPRI5 and of the cubic-Hermite
    family; bitwise against the independent Python restatement, and dy/drho against a finite difference."""
    import py_restatement as pr
    src = """
const double s = p[0], r = p[1], b = p[2];
const double x = y[0], v = y[1], z = y[2];
dydt[0] = s * (v - x); dydt[1] = x * (r - z) - v; dydt[2] = x * v - b * z;
const double J[3][3] = {{-s, s, 0.0}, {r - z, -1.0, -x}, {v, x, -b}};
const double Jp[3][3] = {{v - x, 0.0, 0.0}, {0.0, x, 0.0}, {0.0, 0.0, -z}};
for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double a = Jp[i][j];
    for (int k = 0; k < 3; k++) a = a + J[i][k] * y[3 + 3 * k + j];
    dydt[3 + 3 * i + j] = a;
}"""
    def f(t, y, s=10.0, r=28.0, b=8.0 / 3.0):
        x, v, z = y[0], y[1], y[2]
        J = [[-s, s, 0.0], [r - z, -1.0, -x], [v, x, -b]]
        Jp = [[v - x, 0.0, 0.0], [0.0, x, 0.0], [0.0, 0.0, -z]]
        out = [s * (v - x), x * (r - z) - v, x * v - b * z] + [0.0] * 9
        for i in range(3):
            for j in range(3):
                a = Jp[i][j]
                for k in range(3):
                    a = a + J[i][k] * y[3 + 3 * k + j]
                out[3 + 3 * i + j] = a
        return out
    sysm = deb.ode_from_source(12, src, params=[10.0, 28.0, 8.0 / 3.0])
    y0 = np.zeros((40, 12))
    y0[:, :3] = ob.lorenz_ensemble_y0(40, seed=81)
    te = [0.0, 0.4, 1.1, 2.0]
    for meth in ("dopri5", "rkf45", "dop853"):
        g = deb.EnsembleIVP.ode(sysm, 0.0, 2.0, y0).t_eval(te).method(getattr(E, meth)().rtol(1e-8).atol(1e-8)).solve()
        assert (g.status == 0).all() and (g.n_emitted == 4).all()
        for i in (0, 17, 39):
            if meth == "rkf45":
                p = pr.solve_adaptive(f, meth, 0.0, 2.0, y0[i].tolist(), rtol=1e-8, atol=1e-8, t_eval=te)
            else:
                p = pr.solve_dp(f, meth, 0.0, 2.0, y0[i].tolist(), rtol=1e-8, atol=1e-8, t_eval=te)
            assert (int(g.accepted[i]), int(g.rejected[i]), int(g.evals[i])) == (p["accepted"], p["rejected"], p["evals"])
            assert np.array_equal(bits(g.y_final[i]), bits(p["y"])) and np.array_equal(bits(g.y_eval[i]), bits([r[1] for r in p["rows"]]))
    # dy/drho from the sensitivities against a central finite difference of two plain Lorenz runs
    d = 1e-6
    hi = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0 + d, 8.0 / 3.0), 0.0, 2.0, y0[:, :3].copy()).method(E.dop853().rtol(1e-12).atol(1e-12)).solve()
    lo = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0 - d, 8.0 / 3.0), 0.0, 2.0, y0[:, :3].copy()).method(E.dop853().rtol(1e-12).atol(1e-12)).solve()
    fd = (hi.y_final - lo.y_final) / (2.0 * d)
    sens = deb.EnsembleIVP.ode(sysm, 0.0, 2.0, y0).method(E.dop853().rtol(1e-12).atol(1e-12)).solve().y_final[:, 3:].reshape(-1, 3, 3)[:, :, 1]
    np.testing.assert_allclose(sens, fd, rtol=2e-4, atol=1e-6)


def test_user_defined_rhs_compile_error_is_reported():
    bad = deb.ode_from_source(1, "dydt[0] = undefined_symbol * y[0];", [1.0])
    with pytest.raises(ValueError, match="did not compile"):
        deb.EnsembleIVP.ode(bad, 0.0, 1.0, [[1.0]]).method(E.dopri5()).solve()


@pytest.mark.parametrize("meth", ["dopri5", "dop853", "rkf45", "rkv655e", "rk4"])
def test_non_finite_and_degenerate_initial_states(meth):
    """NaN, +-inf, all-zero, -0.0 and 1e300 initial states next to ordinary ones: every family reproduces what the reference
    arithmetic does with them (Dormand-Prince: the 2-norm keeps NaN -> rejected until MaxSteps; adaptive family: the
    infinity norm drops NaN terms -> "Complete" with a NaN state; h_init overflow -> BadInput), for three recorders."""
    y0 = ob.lorenz_ensemble_y0(8)
    y0[1, 0] = np.nan; y0[2, 1] = np.inf; y0[3, 2] = -np.inf; y0[4, :] = 0.0; y0[5, :] = 1e300; y0[6, :] = -0.0
    for setter in (lambda i: i.t_eval([0.5, 1.0]), lambda i: i.even(0.25), lambda i: i.every_step(50)):
        def prob():
            m = E.rk4(0.01) if meth == "rk4" else getattr(E, meth)().rtol(1e-8)
            return setter(deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0)).method(m)
        g, c = prob().solve(), ob.oracle_solve(prob())
        for name in ("status", "accepted", "rejected", "evals", "n_emitted"):
            assert np.array_equal(getattr(g, name), getattr(c, name)), name
        assert np.array_equal(bits(g.y_final), bits(c.y_final)) and np.array_equal(bits(g.t_final), bits(c.t_final))
        m = np.arange(g.y_eval.shape[1])[None, :] < np.minimum(g.n_emitted, g.y_eval.shape[1])[:, None]
        assert np.array_equal(bits(g.y_eval)[m], bits(c.y_eval)[m])
        assert g.status[0] == 0 and g.status[7] == 0 and np.isfinite(g.y_final[[0, 4, 6, 7]]).all()


def test_degenerate_options_bitwise():
    """Options at and beyond their sensible range -- zero tolerances (division by a zero scale), max_steps 0 / 1, h_min
    above the natural step, h_min/h_max pinching, h0 too large or of the wrong sign, a safety factor above 1, min_scale
    above max_scale, max_rejects 0 / 1 -- with duplicated / out-of-range t_eval points: whatever the reference arithmetic
    does with them, the GPU does the same, bit for bit."""
    y0 = ob.lorenz_ensemble_y0(40)
    cases = [lambda m: m.rtol(0.0).atol(0.0), lambda m: m.rtol(0.0).atol(1e-9), lambda m: m.max_steps(0), lambda m: m.max_steps(1),
             lambda m: m.h_min(1e-3), lambda m: m.h_min(1e-3).h_max(2e-3), lambda m: m.h0(0.9), lambda m: m.h0(-0.1),
             lambda m: m.safety_factor(1.5).min_scale(0.9).max_scale(1.1), lambda m: m.min_scale(5.0).max_scale(0.1),
             lambda m: m.max_rejects(1), lambda m: m.max_rejects(0)]
    for meth in ("dopri5", "dop853", "cash_karp", "rkv766e"):
        for opt in cases:
            for te in ([0.3, 0.3, 1.0, -1.0, 2.0], []):
                def prob():
                    return deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0).t_eval(te).method(opt(getattr(E, meth)()))
                assert_same_solution(prob().solve(), ob.oracle_solve(prob()))


# ------------------------------------------------------------------------------------------ adaptive family
@pytest.mark.parametrize("ctor", ["rkf45", "cash_karp"])
def test_adaptive_family_bit_exact(ctor):
    """RKF45 / Cash-Karp (src/methods/erk/adaptive/ordinary.rs): y_high - y_low error in the infinity norm,
    max_rejects stiffness rule, cubic-Hermite dense output, the doubled h_init evaluation count."""
    y0 = ob.lorenz_ensemble_y0(2000, seed=31)
    te = np.linspace(0.0, 10.0, 21)
    def prob():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 10.0, y0).t_eval(te).method(getattr(E, ctor)().rtol(1e-7).atol(1e-8))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c)
    assert (g.status == 0).all()
    assert np.array_equal(g.evals, 5 + 5 * (g.accepted + g.rejected) + g.accepted)
    # parameter sweep, backward time, vector tolerances, explicit h0 (no h_init: evals base 1)
    mu = np.linspace(0.1, 6.0, 500)
    def p2():
        return (deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 10.0, 0.0, np.tile([2.0, 0.0], (500, 1))).t_eval([7.5, 2.5, 0.0])
                .method(getattr(E, ctor)().rtol([1e-6, 1e-7]).atol([1e-8, 1e-9]).h0(-1e-3)))
    g, c = p2().solve(), ob.oracle_solve(p2())
    assert_same_solution(g, c)
    assert np.array_equal(g.evals, 1 + 5 * (g.accepted + g.rejected) + g.accepted)
    # the max_rejects rule: a stiff problem with a tiny max_rejects ends in Err(Stiffness), the rejected attempt uncounted
    def p3():
        return deb.EnsembleIVP.ode(deb.RobertsonProblem(), 0.0, 40.0, np.tile([1.0, 0.0, 0.0], (16, 1))).method(getattr(E, ctor)().h0(0.5).max_rejects(3))
    g, c = p3().solve(), ob.oracle_solve(p3())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_STIFFNESS).all()
    with pytest.raises(deb.Stiffness):
        g[0]


VERNER = {"rkv655e": (9, 10, True), "rkv656e": (9, 12, True), "rkv766e": (10, 13, False), "rkv767e": (10, 16, False),
          "rkv877e": (13, 17, False), "rkv878e": (13, 21, False), "rkv988e": (16, 21, False), "rkv989e": (16, 26, False)}


@pytest.mark.parametrize("ctor", sorted(VERNER))
def test_verner_family_bit_exact(ctor):
    """Verner pairs (adaptive/mod.rs:59-122, tableau/verner.rs): the adaptive-family stepper plus I-S dense-output stages,
    the Horner dense-output polynomial (adaptive/ordinary.rs:246-277), FSAL for the 6(5) pairs, and the evaluation count
    of the reference (dense stages on every accepted step)."""
    S, I, fsal = VERNER[ctor]
    per_acc = (I - S) + (0 if fsal else 1)
    y0 = ob.lorenz_ensemble_y0(1200, seed=51)
    te = np.concatenate([np.linspace(0.0, 6.0, 25), [6.0, 3.3333]])
    def prob():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 6.0, y0).t_eval(te).method(getattr(E, ctor)().rtol(1e-8).atol(1e-9))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c)
    assert (g.status == 0).all() and (g.n_emitted == len(te)).all()
    assert np.array_equal(g.evals, 5 + (S - 1) * (g.accepted + g.rejected) + per_acc * g.accepted)
    # per-trajectory parameters, backward time, vector tolerances, explicit h0, bounded h
    mu = np.linspace(0.1, 4.0, 300)
    def p2():
        return (deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 5.0, 0.0, np.tile([2.0, 0.0], (300, 1))).t_eval([3.75, 1.25, 0.0, 4.999])
                .method(getattr(E, ctor)().rtol([1e-6, 1e-7]).atol([1e-8, 1e-9]).h0(-1e-3).h_max(0.25)))
    g, c = p2().solve(), ob.oracle_solve(p2())
    assert_same_solution(g, c)
    assert np.array_equal(g.evals, 1 + (S - 1) * (g.accepted + g.rejected) + per_acc * g.accepted)
    # even(dt): every row through the polynomial; and the interpolant is accurate (harmonic oscillator, closed form)
    def p3():
        return (deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 10.0, np.tile([1.0, 0.0], (96, 1)) * np.linspace(0.5, 2.0, 96)[:, None])
                .even(0.37).method(getattr(E, ctor)().rtol(1e-10).atol(1e-10)))
    g, c = p3().solve(), ob.oracle_solve(p3())
    assert_same_solution(g, c)
    s = g[95]
    np.testing.assert_allclose(s.y[:, 0], 2.0 * np.cos(s.t), atol=5e-7)
    # failures: MaxSteps mid-way (rows up to the failure), Stiffness through max_rejects
    def p4():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 20.0, y0[:64]).t_eval(np.arange(0.5, 20.0, 0.5)).method(getattr(E, ctor)().rtol(1e-9).max_steps(40))
    g, c = p4().solve(), ob.oracle_solve(p4())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_MAX_STEPS).all()
    def p5():
        return deb.EnsembleIVP.ode(deb.RobertsonProblem(), 0.0, 40.0, np.tile([1.0, 0.0, 0.0], (16, 1))).method(getattr(E, ctor)().h0(0.5).max_rejects(3))
    g, c = p5().solve(), ob.oracle_solve(p5())
    assert_same_solution(g, c)


def test_verner_user_defined_rhs_bitwise():
    """The NVRTC path instantiates the same kernel for a user system: Lorenz written as source == the built-in, for an
    FSAL pair and for the widest tableau."""
    y0 = ob.lorenz_ensemble_y0(200, seed=52)
    src = "dydt[0] = p[0] * (y[1] - y[0]); dydt[1] = y[0] * (p[1] - y[2]) - y[1]; dydt[2] = y[0] * y[1] - p[2] * y[2];"
    for ctor in ("rkv655e", "rkv989e"):
        usr = deb.ode_from_source(3, src, params=[10.0, 28.0, 8.0 / 3.0])
        def prob(sysm):
            return deb.EnsembleIVP.ode(sysm, 0.0, 3.0, y0).t_eval([0.7, 1.9, 3.0]).method(getattr(E, ctor)().rtol(1e-8))
        g, b = prob(usr).solve(), prob(lorenz()).solve()
        assert_same_solution(g, b)


# ------------------------------------------------------------------------------------------ per-step recorders
RECORDERS = {"every_step": lambda ivp, cap: ivp.every_step(cap), "dense3": lambda ivp, cap: ivp.dense(3, 3 * cap),
             "dense1": lambda ivp, cap: ivp.dense(1, cap), "cross_x_both": lambda ivp, cap: ivp.crossing(0, 0.5, deb.CROSSING_BOTH, 64),
             "cross_z_up": lambda ivp, cap: ivp.crossing(2, 25.0, deb.CROSSING_POSITIVE, 64),
             "cross_y_down": lambda ivp, cap: ivp.crossing(1, -2.0, deb.CROSSING_NEGATIVE, 64),
             "plane_xy": lambda ivp, cap: ivp.hyperplane_crossing([0.0, 0.0], [1.0, -1.0], [0, 1], deb.CROSSING_BOTH, 64),
             "plane_xyz_up": lambda ivp, cap: ivp.hyperplane_crossing([1.0, 2.0, 20.0], [0.3, -0.2, 1.0], [0, 1, 2], deb.CROSSING_POSITIVE, 64)}


@pytest.mark.parametrize("rec", sorted(RECORDERS))
@pytest.mark.parametrize("meth", ["dopri5", "dop853", "cash_karp", "rkv655e", "rkv878e", "rk4"])
def test_per_step_recorders_bit_exact(meth, rec):
    """DefaultSolout / DenseSolout / CrossingSolout / HyperplaneCrossingSolout (src/solout/default.rs, dense.rs, crossing.rs,
    hyperplane.rs): rows with their own times, Newton-refined crossings on each method's own dense output; kernels compiled
    at first use."""
    y0 = ob.lorenz_ensemble_y0(300, seed=61)
    def prob():
        m = E.rk4(0.01) if meth == "rk4" else getattr(E, meth)().rtol(1e-7).atol(1e-8)
        return RECORDERS[rec](deb.EnsembleIVP.ode(lorenz(), 0.0, 6.0, y0), 700).method(m)
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c)
    assert (g.status == 0).all()
    s = g[17]
    assert len(s.t) == g.n_emitted[17] and (np.diff(s.t) > 0).all()
    if rec == "every_step":
        assert np.array_equal(g.n_emitted, g.accepted + 1) and s.t[0] == 0.0 and s.t[-1] == 6.0
    if rec == "dense3":
        assert np.array_equal(g.n_emitted, 3 * g.accepted + 1)
    if rec == "cross_z_up":
        # Newton-refined rows sit on the threshold; rows where the reference's iteration gives up use its linear fallback
        assert len(s.t) >= 2
        if meth == "dopri5":
            assert np.median(np.abs(s.y[:, 2] - 25.0)) < 1e-9


def test_per_step_recorders_capacity_backward_and_failures():
    """Row capacity smaller than the number of rows (counted, not stored), backward time, trajectories that fail mid-way."""
    y0 = np.tile([1.0, 0.0], (64, 1)) * np.linspace(0.5, 2.0, 64)[:, None]
    def p1():  # capacity overflow
        return deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 20.0, y0).every_step(10).method(E.dopri5().rtol(1e-8))
    g, c = p1().solve(), ob.oracle_solve(p1())
    assert np.array_equal(g.n_emitted, c.n_emitted) and (g.n_emitted > 10).all()
    assert np.array_equal(bits(g.y_eval), bits(c.y_eval)) and np.array_equal(bits(g.t_out), bits(c.t_out))
    with pytest.raises(ValueError, match="row capacity"):
        g[0]
    for setter in (lambda ivp: ivp.every_step(400), lambda ivp: ivp.dense(4, 1600), lambda ivp: ivp.crossing(0, 0.1, deb.CROSSING_BOTH, 16)):
        for m in (lambda: E.dopri5().rtol(1e-9), lambda: E.dop853().rtol(1e-9), lambda: E.rkf45().rtol(1e-8), lambda: E.rkv766e().rtol(1e-9),
                  lambda: E.rk4(-0.05)):
            def p2():  # backward
                return setter(deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 10.0, 0.0, y0)).method(m())
            g, c = p2().solve(), ob.oracle_solve(p2())
            assert_same_solution(g, c)
            assert (g.status == 0).all()
    def p3():  # MaxSteps: rows up to the failure
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 50.0, ob.lorenz_ensemble_y0(96, seed=62)).dense(2, 400).method(E.dopri5().rtol(1e-9).max_steps(150))
    g, c = p3().solve(), ob.oracle_solve(p3())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_MAX_STEPS).all()


def test_crossing_example_08_damped_oscillator_user_rhs():
    """examples/ode/08_damped_oscillator: x'' = -b x' - k x as a user-defined right-hand side, zero crossings of x with
    dopri5().rtol(1e-8).atol(1e-8): against the independent Python restatement (bitwise) and the closed form."""
    import py_restatement as pr
    osc = deb.ode_from_source(2, "dydt[0] = y[1]; dydt[1] = -p[0] * y[1] - p[1] * y[0];", params=[0.5, 1.0])
    g = (deb.EnsembleIVP.ode(osc, 0.0, 20.0, [[1.0, 0.0]]).crossing(0, 0.0, deb.CROSSING_BOTH, 32)
         .method(E.dopri5().rtol(1e-8).atol(1e-8)).solve())
    f = lambda t, y: [y[1], -0.5 * y[1] - 1.0 * y[0]]
    p = pr.solve_dp(f, "dopri5", 0.0, 20.0, [1.0, 0.0], rtol=1e-8, atol=1e-8, recorder=("crossing", 0, 0.0, 0))
    s = g[0]
    assert (int(g.accepted[0]), int(g.rejected[0]), int(g.evals[0])) == (p["accepted"], p["rejected"], p["evals"])
    assert np.array_equal(bits(s.t), bits([r[0] for r in p["rows"]])) and np.array_equal(bits(s.y), bits([r[1] for r in p["rows"]]))
    # x(t) = e^{-t/4} (cos wt + sin wt / (4w)), w = sqrt(15)/4: zeros at wt = atan(-4w) + k pi
    w = math.sqrt(15.0) / 4.0
    zeros = [(math.atan(-4.0 * w) + k * math.pi) / w for k in range(1, 8)]
    zeros = [z for z in zeros if z < 20.0]
    assert len(s.t) == len(zeros)
    np.testing.assert_allclose(s.t, zeros, atol=1e-5)
    assert np.abs(s.y[:, 0]).max() < 1e-9


# ------------------------------------------------------------------------------------------ events
@pytest.mark.parametrize("meth", ["dopri5", "dop853", "rkf45", "rkv655e", "rk4"])
def test_event_detection_bit_exact(meth):
    """IVP::event(&e) (EventWrappedSolout + Brent-Dekker, src/solout/event.rs:300-470) around every recorder: rows, event
    rows, counters and the Interrupted status, bitwise against the oracle."""
    n = 200
    def m():
        return E.rk4(0.01) if meth == "rk4" else getattr(E, meth)().rtol(1e-9).atol(1e-9)
    # examples/ode/03_logistic_growth as a sweep over the carrying capacity: even(2.0) + terminal event y - 9
    cap = np.linspace(9.5, 30.0, n)
    def p1():
        return (deb.EnsembleIVP.ode(deb.LogisticEquation(1.0, cap), 0.0, 10.0, np.ones((n, 1))).even(2.0)
                .event(deb.LinearEvent(-9.0, 0.0, [1.0]), terminate=1).method(m()))
    g, c = p1().solve(), ob.oracle_solve(p1())
    assert_same_solution(g, c)
    assert (g.status == deb.DEB_STATUS_INTERRUPTED).all()
    s = g[0]
    assert s.status == "Interrupted" and abs(s.y[-1, 0] - 9.0) < 1e-6 and (np.diff(s.t) > 0).all()
    # zero crossings of x on t_eval rows, every direction, terminate_after(k) and never
    y0 = np.tile([1.0, 0.0], (n, 1)) * np.linspace(0.5, 2.0, n)[:, None]
    kk = np.linspace(0.5, 4.0, n)
    for direction, term in ((deb.CROSSING_BOTH, None), (deb.CROSSING_POSITIVE, None), (deb.CROSSING_NEGATIVE, 2), (deb.CROSSING_BOTH, 3)):
        def p2():
            return (deb.EnsembleIVP.ode(deb.HarmonicOscillator(kk), 0.0, 10.0, y0).t_eval([1.0, 2.0, 4.0, 7.5, 9.0, 10.0])
                    .event(deb.LinearEvent(0.0, 0.0, [1.0, 0.0]), direction, term).method(m()))
        g, c = p2().solve(), ob.oracle_solve(p2())
        assert_same_solution(g, c)
        assert set(np.unique(g.status)) <= {0, deb.DEB_STATUS_INTERRUPTED}
    # plain solve() + event (every step + event rows), time-dependent event, capacity too small for some trajectories
    def p3():
        return (deb.EnsembleIVP.ode(lorenz(), 0.0, 4.0, ob.lorenz_ensemble_y0(n, seed=71))
                .event(deb.LinearEvent(-25.0, 1.0, [0.0, 0.0, 1.0]), deb.CROSSING_BOTH, None, max_event_rows=300).method(m()))
    g, c = p3().solve(), ob.oracle_solve(p3())
    assert_same_solution(g, c)
    # dense(2) and crossing bases; backward time (no events by the reference's bounds test)
    for setter in (lambda ivp: ivp.dense(2, 900), lambda ivp: ivp.crossing(1, 0.0, deb.CROSSING_BOTH, 64)):
        def p4():
            return setter(deb.EnsembleIVP.ode(deb.HarmonicOscillator(kk), 0.0, 6.0, y0)).event(deb.LinearEvent(-0.25, 0.0, [1.0, 0.0])).method(m())
        g, c = p4().solve(), ob.oracle_solve(p4())
        assert_same_solution(g, c)
    def p5():
        mm = E.rk4(-0.01) if meth == "rk4" else m()
        return deb.EnsembleIVP.ode(deb.HarmonicOscillator(kk), 6.0, 0.0, y0).t_eval([3.0, 0.0]).event(deb.LinearEvent(0.0, 0.0, [1.0, 0.0])).method(mm)
    g, c = p5().solve(), ob.oracle_solve(p5())
    assert_same_solution(g, c)


def test_events_in_a_few_trajectories_only():
    """Lanes whose step holds an event candidate wait for each other (erk_ensemble.cuh: rec_go).  An ensemble in which one
    trajectory in eight ever has events (Lorenz rho = 28 among rho = 10: z settles at 9 and never reaches 25) never fills the
    warp's threshold: the waiting lanes are released by their idle budget.  Results as ever: bitwise the oracle's."""
    n = 512
    rho = np.where(np.arange(n) % 8 == 3, 28.0, 10.0)
    y0 = ob.lorenz_ensemble_y0(n, seed=77)
    ev = deb.LinearEvent(-25.0, 0.0, [0.0, 0.0, 1.0])
    for meth in ("dopri5", "rkv656e"):
        def prob():
            return (deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, rho, 8.0 / 3.0), 0.0, 8.0, y0).t_eval(np.linspace(0.0, 8.0, 33))
                    .event(ev, max_event_rows=64).method(getattr(E, meth)().rtol(1e-8)))
        g, c = prob().solve(), ob.oracle_solve(prob())
        assert_same_solution(g, c)
        assert (g.n_emitted[rho == 10.0] <= 34).all() and (g.n_emitted[rho == 28.0] > 36).all()


def test_user_defined_event_function_vs_python_restatement():
    """examples/ode/05_damped_pendulum pattern (linearised so that the right-hand side is IEEE-exact on both sides): a
    user-defined right-hand side AND a user-defined terminal event g = max(|theta|, |omega|) - 0.01 on t_eval rows, rkf45():
    bitwise against the independent Python restatement."""
    import py_restatement as pr
    pend = deb.ode_from_source(2, "dydt[0] = y[1]; dydt[1] = -(p[0] / p[1]) * y[1] - (p[2] / p[3]) * y[0];", params=[0.2, 1.0, 9.81, 1.0])
    near_rest = deb.event_from_source(2, "return fmax(fabs(y[0]), fabs(y[1])) - 0.01;")
    t_out = [0.0, 1.0, 3.0, 4.5, 6.9, 10.0]
    g = (deb.EnsembleIVP.ode(pend, 0.0, 100.0, [[1.0, 0.0]]).t_eval(t_out).event(near_rest, terminate=1).method(E.rkf45()).solve())
    f = lambda t, y: [y[1], -(0.2 / 1.0) * y[1] - (9.81 / 1.0) * y[0]]
    p = pr.solve_adaptive(f, "rkf45", 0.0, 100.0, [1.0, 0.0], t_eval=t_out, event=(lambda t, y: pr.rmax(abs(y[0]), abs(y[1])) - 0.01, 0, 1))
    s = g[0]
    assert s.status == p["status"] == "Interrupted" and len(s.t) == 7 and s.t[:6].tolist() == t_out
    assert (int(g.accepted[0]), int(g.rejected[0]), int(g.evals[0])) == (p["accepted"], p["rejected"], p["evals"])
    assert np.array_equal(bits(s.t), bits([r[0] for r in p["rows"]])) and np.array_equal(bits(s.y), bits([r[1] for r in p["rows"]]))
    assert abs(max(abs(s.y[-1, 0]), abs(s.y[-1, 1])) - 0.01) < 1e-6 and 40.0 < s.t[-1] < 60.0


# ------------------------------------------------------------------------------------------ EvenSolout
@pytest.mark.parametrize("meth", ["dopri5", "dop853", "rkf45", "rk4"])
def test_even_solout_bit_exact(meth):
    """IVP::even(dt) (src/solout/even.rs): t0, t0+dt, ... accumulated, every point interpolated (no exact-hit shortcut),
    and the final-point rule at tf (replace a near-duplicate / append)."""
    y0 = ob.lorenz_ensemble_y0(1500, seed=41)
    for dt, tf in ((0.3, 2.0), (0.5, 2.0), (0.25, 6.0), (7.0, 3.0)):  # append tf / replace at tf / many rows / dt > interval
        def prob():
            m = getattr(E, meth)(0.01) if meth == "rk4" else getattr(E, meth)().rtol(1e-8).atol(1e-8)
            return deb.EnsembleIVP.ode(lorenz(), 0.0, tf, y0).even(dt).method(m)
        g, c = prob().solve(), ob.oracle_solve(prob())
        assert_same_solution(g, c)
        assert (g.status == 0).all()
        s = g[7]
        assert s.t[0] == 0.0 and s.t[-1] == tf and len(s.t) == len(s.y)
    # backward integration and a trajectory that stops early (MaxSteps): rows up to the failure only
    def back():
        m = E.dopri5().max_steps(60) if meth != "rk4" else E.rk4(-0.01).max_steps(150)
        return deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 3.0, 0.0, np.tile([1.0, 0.0], (64, 1)) + np.linspace(0, 0.2, 64)[:, None]).even(0.4).method(m)
    g, c = back().solve(), ob.oracle_solve(back())
    assert_same_solution(g, c)


def test_milstein_matches_host_regenerated_philox():
    """Milstein::new(h) (src/methods/milstein.rs:107-180), derivative-free correction, on GBM (multiplicative noise, where
    the correction is non-zero) and OU (additive noise: the correction vanishes identically)."""
    n = 4096
    for sysm, y0, tf, h in ((deb.GeometricBrownianMotion(0.1, 0.2), 100.0, 1.0, 1e-3), (deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), 5.0, 2.0, 0.01)):
        def prob(meth):
            return deb.EnsembleIVP.sde(sysm, 0.0, tf, np.full(n, y0), seed=7, path_offset=10 ** 10).t_eval([tf / 3, tf]).method(meth)
        g, c = prob(deb.Milstein.new(h)).solve(), ob.oracle_solve(prob(deb.Milstein.new(h)))
        assert_same_solution(g, c, exact=False, rtol=1e-12)
        assert (g.status == 0).all() and np.array_equal(g.evals, 1 + 3 * g.accepted)
        em = prob(E.euler(h)).solve()
        if sysm.system_id == deb.DEB_SDE_OU:
            np.testing.assert_allclose(g.y_final, em.y_final, rtol=1e-12)  # additive noise: Milstein == Euler-Maruyama
        else:
            assert np.abs(g.y_final - em.y_final).max() > 1e-6               # multiplicative noise: the correction acts


def test_plain_c_caller_reproduces_the_known_answers(tmp_path):
    """examples/c_caller/lorenz_ensemble.c through the C ABI with host buffers, no Python in the loop: the L100 / L10 known
    answers of SURVEY.md appendix A (step counts, final state and t_eval rows of the (1,1,1) Lorenz trajectory), bit for bit."""
    import subprocess
    from test_abi_cpu import build_c_caller
    out = subprocess.run([build_c_caller(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "known answers reproduced" in out.stdout, out.stdout + out.stderr


def test_cpp_host_mirror_example_on_the_gpu(tmp_path):
    """examples/cpp_caller/builder_api.cpp: the crate's builder calls in C++ (include/deb_ensemble.hpp) on the GPU -- known answers
    bit for bit, a swept terminal event, a user-defined right-hand side with crossings, the Error variants."""
    import subprocess
    from test_abi_cpu import build_cpp_caller
    out = subprocess.run([build_cpp_caller(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
