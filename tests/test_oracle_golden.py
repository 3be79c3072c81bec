"""CPU tests (no GPU): pin the oracle (oracle/oracle.cpp) against every golden vector the reference's own tests hold
for the explicit-RK path, against the survey's probe values, and against a second independent restatement.

The reference is a Rust crate and cannot be run here (no toolchain): at the bit level parity is "unpinned" (see
oracle/oracle.cpp header); these tests are everything that CAN be pinned."""
import importlib
import json
import math
import os

import numpy as np
import pytest

import oracle_binding as ob
import py_restatement as pr

deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SYSTEMS = {"exponential": lambda p: deb.ExponentialGrowth(*p), "linear": lambda p: deb.LinearEquation(*p),
           "harmonic": lambda p: deb.HarmonicOscillator(*p), "logistic": lambda p: deb.LogisticEquation(*p),
           "robertson": lambda p: deb.RobertsonProblem()}


def make_method(c):
    if c["solver"] in ("dopri5", "dop853", "rkf45", "cash_karp") or c["solver"].startswith("rkv"):
        m = getattr(E, c["solver"])()
    else:
        m = getattr(E, c["solver"])(c["h"])
    if c["rtol"] is not None:
        m.rtol(c["rtol"])
    if c["atol"] is not None:
        m.atol(c["atol"])
    return m


def test_reference_accuracy_goldens():
    """tests/ode/accuracy.rs: final state vs SciPy DOP853 constants with the reference's own tolerances."""
    cases = json.load(open(os.path.join(GOLDEN, "reference_accuracy.json")))["accuracy"]
    assert len(cases) == 87
    for c in cases:
        ivp = deb.EnsembleIVP.ode(SYSTEMS[c["system"]](c["params"]), c["t0"], c["tf"], [c["y0"]]).method(make_method(c))
        s = ob.oracle_solve(ivp)[0]  # the reference test unwrap()s: every case must solve
        err = np.abs(s.y_final - np.array(c["expected"]))
        assert (err < c["tolerance"]).all(), (c["case"], c["solver"], s.y_final, c["expected"])


def test_reference_accuracy_tight():
    """The same SciPy constants are good to ~1e-8 relative: DOP853 at rtol=atol=1e-12 must reproduce them that well
    (much tighter than the reference's 1e-3), which pins tableau constants and controller far better."""
    cases = [c for c in json.load(open(os.path.join(GOLDEN, "reference_accuracy.json")))["accuracy"]
             if c["solver"] == "dop853" and c["system"] != "robertson"]
    for c in cases:
        ivp = deb.EnsembleIVP.ode(SYSTEMS[c["system"]](c["params"]), c["t0"], c["tf"], [c["y0"]]).method(make_method(c))
        s = ob.oracle_solve(ivp)[0]
        np.testing.assert_allclose(s.y_final, c["expected"], rtol=2e-8, atol=1e-8)


def test_reference_interpolation_kat():
    """tests/ode/interpolation.rs:35-79: y' = y, t_eval([0.5, 1.0, 1.69]); every solver within 1e-3 of e^t.  Only the
    listed points come back, in order."""
    for m in (E.dop853(), E.dopri5(), E.rkv655e(), E.rkv877e(), E.rkv988e(), E.rkf45(), E.rk4(0.01)):  # the explicit-RK solvers of that test
        s = ob.oracle_solve(deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 2.0, [[1.0]]).t_eval([0.5, 1.0, 1.69]).method(m))[0]
        assert s.t.tolist() == [0.5, 1.0, 1.69]
        np.testing.assert_allclose(s.y[:, 0], np.exp([0.5, 1.0, 1.69]), atol=1e-3)


def test_reference_from_fn_euler_kat():
    """tests/ode/from_fn.rs:4-18."""
    s = ob.oracle_solve(deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.euler(0.1)))[0]
    assert abs(s.y_final[0] - 2.5937) < 1e-3


def test_reference_error_kats():
    """tests/ode/errors.rs:87-130: tf == t0 and h0 > interval are BadInput."""
    for ivp in (deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 0.0, [[1.0]]).method(E.dopri5()),
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.dopri5().h0(10.0)),
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.rk4(10.0)),
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 0.0, [[1.0]]).method(E.rkv655e().h0(0.1)),   # errors.rs:102-103
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 0.0, [[1.0]]).method(E.rkv988e().h0(0.1)),
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.rkv655e().h0(10.0)),  # errors.rs:125-126
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.rkv988e().h0(10.0)),
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.rk4(-0.1)),
                deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 1.0, [[1.0]]).method(E.dopri5().h_min(1.0).h_max(0.5))):
        with pytest.raises(deb.BadInput):
            ob.oracle_solve(ivp)[0]


def test_reference_pde_kats():
    """tests/pde/method_of_lines.rs:37-157."""
    x = np.linspace(0.0, 1.0, 41)
    s = ob.oracle_heat(np.sin(np.pi * x), 0.0, 1.0, 0.1, E.rk4(1.0e-4), 0.0, 0.02)
    assert np.abs(s.u - math.exp(-0.1 * math.pi ** 2 * 0.02) * np.sin(np.pi * x)).max() < 5.0e-4
    du = ob.oracle_heat_rhs([2.0, 1.0, 0.0, -0.5, -1.0], 0.0, 1.0, 1.0, ("dirichlet", 2.0), ("dirichlet", -1.0))
    assert du[0] == 0.0 and du[-1] == 0.0
    du = ob.oracle_heat_rhs([0.0, 1.0, 0.0, -1.0, 0.0], 0.0, 1.0, 1.0)
    assert du[0] == 0.0 and du[4] == 0.0 and abs(du[2]) < 1e-12
    du = ob.oracle_heat_rhs([1.0, 2.0, 2.0, 2.0, 2.0], 0.0, 1.0, 1.0, ("neumann", 0.0), ("neumann", 0.0))
    assert abs(du[0] - (1.0 / 0.25) / 0.25) < 1e-12


def test_survey_probe_known_answers():
    """SURVEY.md Appendix A (an independent Python restatement run by the surveyor; tentative known-answers)."""
    lz = deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0)
    s = ob.oracle_solve(deb.EnsembleIVP.ode(lz, 0.0, 100.0, [[1.0, 1.0, 1.0]]).method(E.dopri5().rtol(1e-8)))
    assert (int(s.accepted[0]), int(s.rejected[0]), int(s.evals[0])) == (6008, 411, 44525)
    assert [v.hex() for v in s.y_final[0]] == ["-0x1.8c65ec78fb24fp+1", "-0x1.4d4949b09d281p+1", "0x1.5bf4de76ed043p+4"]
    s = ob.oracle_solve(deb.EnsembleIVP.ode(lz, 0.0, 10.0, [[1.0, 1.0, 1.0]]).t_eval([1.0, 2.5, 10.0]).method(E.dopri5().rtol(1e-8)))
    assert (int(s.accepted[0]), int(s.rejected[0]), int(s.evals[0])) == (504, 8, 3579)
    assert [v.hex() for v in s.y_final[0]] == ["-0x1.39c5be73410dbp+2", "-0x1.df3a8c489ba9ep+1", "0x1.8b0d329428ea8p+4"]
    assert s.y_eval[0, 0].tolist() == [-9.378567031548476, -8.357039113339846, 29.36231537793399]
    assert s.y_eval[0, 1].tolist() == [-6.959579728354202, -7.272475545395783, 24.70312694098664]
    assert np.array_equal(s.y_eval[0, 2], s.y_final[0])  # exact-hit branch
    s = ob.oracle_solve(deb.EnsembleIVP.ode(lz, 0.0, 100.0, [[1.0, 1.0, 1.0]]).method(E.dopri5().rtol(1e-8).atol(1e-8)))
    assert s.status[0] == deb.DEB_STATUS_MAX_STEPS  # 10343 attempts needed
    s = ob.oracle_solve(deb.EnsembleIVP.ode(lz, 0.0, 100.0, [[1.0, 1.0, 1.0]]).method(E.dopri5().rtol(1e-8).atol(1e-8).max_steps(20000)))
    assert (int(s.accepted[0]), int(s.rejected[0]), int(s.evals[0])) == (10126, 217, 72187)
    s = ob.oracle_solve(deb.EnsembleIVP.ode(deb.VanDerPolOscillator(5.0), 0.0, 10.0, [[2.0, 0.0]]).method(E.dop853().rtol(1e-8).atol(1e-8)))
    assert (int(s.accepted[0]), int(s.rejected[0]), int(s.evals[0])) == (83, 12, 1380)
    assert s.y_final[0].tolist() == [-1.1587012635243283, 0.43046981176622706]
    mu = np.array([0.1, 1.0, 5.0, 10.0, 20.0, 50.0])
    s = ob.oracle_solve(deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 0.0, 100.0, np.tile([2.0, 0.0], (6, 1))).method(E.dop853().rtol(1e-8).atol(1e-8)))
    assert (s.accepted + s.rejected).tolist() == [287, 767, 1078, 1069, 1073, 1813] and (s.status == 0).all()
    # fixed-step schedules of the solve loop
    for (tf, h, steps) in ((1.0, 1e-3, 1000), (10.0, 0.01, 1001), (100.0, 1.0, 100), (1.0, 0.01, 100)):
        s = ob.oracle_solve(deb.EnsembleIVP.ode(deb.ExponentialGrowth(0.0), 0.0, tf, [[1.0]]).method(E.euler(h).max_steps(100000)))
        assert int(s.accepted[0]) == steps
    # dense-output quirks kept as written: DOPRI5 cont[4] == 0, DOP853 s/s1 order
    e5 = ob.oracle_solve(deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 2.0, [[1.0]]).t_eval([0.5]).method(E.dopri5()))
    e8 = ob.oracle_solve(deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 2.0, [[1.0]]).t_eval([0.5]).method(E.dop853()))
    assert int(e5.accepted[0]) == 9 and int(e8.accepted[0]) == 3
    assert abs((e5.y_eval[0, 0, 0] - math.exp(0.5)) - (-1.89e-5)) < 1e-7
    assert abs((e8.y_eval[0, 0, 0] - math.exp(0.5)) - 3.76e-4) < 1e-6


def _same_bits(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float64).view(np.uint64), np.asarray(b, dtype=np.float64).view(np.uint64))


def test_oracle_equals_independent_python_restatement_bitwise():
    """Two independently written restatements (C++ oracle, pure-Python tests/py_restatement.py) agree bit for bit on
    states, dense output rows and counters."""
    cases = [
        ("dopri5", pr.lorenz(10.0, 28.0, 8.0 / 3.0), deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), [1.0, 1.0, 1.0], 0.0, 10.0, dict(rtol=1e-8, atol=1e-6)),
        ("dopri5", pr.lorenz(10.0, 28.0, 8.0 / 3.0), deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), [0.7, 1.3, 1.1], 0.0, 12.0, dict(rtol=1e-6, atol=1e-9)),
        ("dop853", pr.van_der_pol(5.0), deb.VanDerPolOscillator(5.0), [2.0, 0.0], 0.0, 10.0, dict(rtol=1e-8, atol=1e-8)),
        ("dop853", pr.lorenz(10.0, 28.0, 8.0 / 3.0), deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), [1.0, 1.0, 1.0], 0.0, 8.0, dict(rtol=1e-9, atol=1e-9)),
        ("dopri5", pr.harmonic(1.0), deb.HarmonicOscillator(1.0), [1.0, 0.0], 10.0, 0.0, dict(rtol=1e-7, atol=1e-7)),  # backward
        ("dop853", pr.exponential(1.0), deb.ExponentialGrowth(1.0), [1.0], 0.0, 2.0, dict(rtol=1e-6, atol=1e-6)),
    ]
    for meth, f, sysm, y0, t0, tf, tol in cases:
        te = list(np.linspace(t0, tf, 9)) + [0.5 * (t0 + tf) + 0.123]
        p = pr.solve_dp(f, meth, t0, tf, y0, t_eval=te, **tol)
        c = ob.oracle_solve(deb.EnsembleIVP.ode(sysm, t0, tf, [y0]).t_eval(te).method(getattr(E, meth)().rtol(tol["rtol"]).atol(tol["atol"])))
        assert p["status"] == "Complete" and c.status[0] == 0
        assert (p["accepted"], p["rejected"], p["evals"]) == (int(c.accepted[0]), int(c.rejected[0]), int(c.evals[0]))
        assert _same_bits(p["y"], c.y_final[0]) and _same_bits([p["t"]], [c.t_final[0]])
        assert len(p["rows"]) == int(c.n_emitted[0])
        assert [r[0] for r in p["rows"]] == c.t_rows[:len(p["rows"])].tolist()
        assert _same_bits([r[1] for r in p["rows"]], c.y_eval[0, :len(p["rows"])])
    for meth in ("euler", "midpoint", "heun", "ralston", "ssp_rk3", "rk4", "three_eighths"):
        p = pr.solve_fixed(pr.lorenz(10.0, 28.0, 8.0 / 3.0), meth, 0.01, 0.0, 1.0, [1.0, 1.0, 1.0], t_eval=[0.0, 0.255, 0.5, 1.0])
        c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, [[1.0, 1.0, 1.0]]).t_eval([0.0, 0.255, 0.5, 1.0]).method(getattr(E, meth)(0.01)))
        assert (p["accepted"], p["evals"]) == (int(c.accepted[0]), int(c.evals[0]))
        assert _same_bits(p["y"], c.y_final[0]) and _same_bits([r[1] for r in p["rows"]], c.y_eval[0, :len(p["rows"])])


def test_oracle_properties():
    """Book-keeping identities and ensemble invariances of the oracle itself."""
    y0 = ob.lorenz_ensemble_y0(64)
    lz = deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0)
    s = ob.oracle_solve(deb.EnsembleIVP.ode(lz, 0.0, 20.0, y0).t_eval(np.arange(1.0, 21.0)).method(E.dopri5().rtol(1e-8)))
    assert np.array_equal(s.evals, 3 + 6 * (s.accepted + s.rejected) + s.accepted)
    assert (s.n_emitted == 20).all() and _same_bits(s.y_eval[:, 19], s.y_final)
    # thread count and ensemble position do not change a trajectory
    s1 = ob.oracle_solve(deb.EnsembleIVP.ode(lz, 0.0, 20.0, y0[::-1].copy()).t_eval(np.arange(1.0, 21.0)).method(E.dopri5().rtol(1e-8)), n_threads=1)
    assert _same_bits(s1.y_eval[::-1], s.y_eval) and np.array_equal(s1.accepted[::-1], s.accepted)
    s8 = ob.oracle_solve(deb.EnsembleIVP.ode(deb.VanDerPolOscillator(1.0), 0.0, 10.0, np.tile([2.0, 0.0], (8, 1))).method(E.dop853()))
    assert np.array_equal(s8.evals, 3 + 11 * (s8.accepted + s8.rejected) + 4 * s8.accepted)


def test_adaptive_family_restatements_agree_bitwise():
    """RKF45 / Cash-Karp (adaptive/ordinary.rs): the C++ oracle and the independent pure-Python restatement agree bit
    for bit, including the doubled h_init evaluation count and the max_rejects rule."""
    lz = pr.lorenz(10.0, 28.0, 8.0 / 3.0)
    for meth in ("rkf45", "cash_karp"):
        cases = ((lz, deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 10.0, [1.0, 1.0, 1.0], dict(rtol=1e-7, atol=1e-8)),
                 (lz, deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 6.0, [0.3, -1.0, 20.0], dict(rtol=1e-6, atol=1e-6)),
                 (pr.harmonic(1.0), deb.HarmonicOscillator(1.0), 10.0, 0.0, [1.0, 0.0], dict(rtol=1e-8, atol=1e-8)))  # backward
        for (f, sysm, t0, tf, y0, tol) in cases:
            te = list(np.linspace(t0, tf, 7)) + [0.5 * (t0 + tf) + 0.0123]
            p = pr.solve_adaptive(f, meth, t0, tf, y0, t_eval=te, **tol)
            c = ob.oracle_solve(deb.EnsembleIVP.ode(sysm, t0, tf, [y0]).t_eval(te)
                                .method(getattr(E, meth)().rtol(tol["rtol"]).atol(tol["atol"])))
            assert p["status"] == "Complete" and c.status[0] == 0
            assert (p["accepted"], p["rejected"], p["evals"]) == (int(c.accepted[0]), int(c.rejected[0]), int(c.evals[0]))
            assert p["evals"] == 5 + 5 * (p["accepted"] + p["rejected"]) + p["accepted"]
            assert _same_bits(p["y"], c.y_final[0]) and _same_bits([r[1] for r in p["rows"]], c.y_eval[0, :len(p["rows"])])
        rob = lambda t, y: [-0.04 * y[0] + 1.0e4 * y[1] * y[2], 0.04 * y[0] - 1.0e4 * y[1] * y[2] - 3.0e7 * y[1] * y[1], 3.0e7 * y[1] * y[1]]
        p = pr.solve_adaptive(rob, meth, 0.0, 40.0, [1.0, 0.0, 0.0], h0=0.5, max_rejects=3)
        c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.RobertsonProblem(), 0.0, 40.0, [[1.0, 0.0, 0.0]]).method(getattr(E, meth)().h0(0.5).max_rejects(3)))
        assert p["status"] == "Stiffness" and c.status[0] == deb.DEB_STATUS_STIFFNESS
        assert (p["accepted"], p["rejected"], p["evals"]) == (int(c.accepted[0]), int(c.rejected[0]), int(c.evals[0]))
        assert _same_bits(p["y"], c.y_final[0]) and p["t"] == c.t_final[0]


VERNER = (("rkv655e", 6, 9, 10, True), ("rkv656e", 6, 9, 12, True), ("rkv766e", 7, 10, 13, False), ("rkv767e", 7, 10, 16, False),
          ("rkv877e", 8, 13, 17, False), ("rkv878e", 8, 13, 21, False), ("rkv988e", 9, 16, 21, False), ("rkv989e", 9, 16, 26, False))


def test_verner_restatements_agree_bitwise():
    """Verner pairs (adaptive/ordinary.rs with bi = Some, adaptive/mod.rs:59-122): C++ oracle vs the independent pure-Python
    restatement, bit for bit -- dense-output stages on every accepted step, Horner polynomial in s, FSAL for the 6(5)
    pairs -- plus the evaluation-count identity and the accuracy of the interpolant at its design order."""
    lz = pr.lorenz(10.0, 28.0, 8.0 / 3.0)
    for (meth, O, S, I, fsal) in VERNER:
        cases = ((lz, deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 4.0, [1.0, 1.0, 1.0], dict(rtol=1e-7, atol=1e-8)),
                 (pr.van_der_pol(1.5), deb.VanDerPolOscillator(1.5), 0.0, 6.0, [2.0, 0.0], dict(rtol=1e-5, atol=1e-6)),
                 (pr.harmonic(1.0), deb.HarmonicOscillator(1.0), 5.0, 0.0, [1.0, 0.0], dict(rtol=1e-9, atol=1e-9)))  # backward
        for (f, sysm, t0, tf, y0, tol) in cases:
            te = list(np.linspace(t0, tf, 7)) + [0.5 * (t0 + tf) + 0.0123]
            p = pr.solve_adaptive(f, meth, t0, tf, y0, t_eval=te, **tol)
            c = ob.oracle_solve(deb.EnsembleIVP.ode(sysm, t0, tf, [y0]).t_eval(te)
                                .method(getattr(E, meth)().rtol(tol["rtol"]).atol(tol["atol"])))
            assert p["status"] == "Complete" and c.status[0] == 0
            assert (p["accepted"], p["rejected"], p["evals"]) == (int(c.accepted[0]), int(c.rejected[0]), int(c.evals[0]))
            assert p["evals"] == 5 + (S - 1) * (p["accepted"] + p["rejected"]) + p["accepted"] * ((I - S) + (0 if fsal else 1))
            assert len(p["rows"]) == 8 == c.n_emitted[0]
            assert _same_bits(p["y"], c.y_final[0]) and _same_bits([r[1] for r in p["rows"]], c.y_eval[0])
        # the interpolant is a real dense output: harmonic oscillator rows against the closed form
        sol = ob.oracle_solve(deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 10.0, [[1.0, 0.0]]).t_eval(np.linspace(0.05, 9.95, 37))
                              .method(getattr(E, meth)().rtol(1e-10).atol(1e-10)))[0]
        np.testing.assert_allclose(sol.y[:, 0], np.cos(sol.t), atol=2e-7)
        np.testing.assert_allclose(sol.y[:, 1], -np.sin(sol.t), atol=2e-7)
        # even(dt) goes through the same polynomial
        ev = ob.oracle_solve(deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 3.0, [[1.0, 0.0]]).even(0.4)
                             .method(getattr(E, meth)().rtol(1e-9).atol(1e-9)))[0]
        assert ev.t[0] == 0.0 and ev.t[-1] == 3.0
        np.testing.assert_allclose(ev.y[:, 0], np.cos(ev.t), atol=1e-6)


def test_per_step_recorder_restatements_agree_bitwise():
    """DefaultSolout / DenseSolout / CrossingSolout (src/solout/default.rs, dense.rs, crossing.rs): C++ oracle vs the
    independent Python restatement, bit for bit (row times and states), for every kernel family; and the documented
    shapes (every step; n rows per step; rows on the threshold)."""
    lz = pr.lorenz(10.0, 28.0, 8.0 / 3.0)
    recs = ((("default",), lambda ivp: ivp.every_step(600)), (("dense", 3), lambda ivp: ivp.dense(3, 1800)),
            (("dense", 1), lambda ivp: ivp.dense(1, 600)), (("crossing", 0, 0.5, 0), lambda ivp: ivp.crossing(0, 0.5, 0, 64)),
            (("crossing", 2, 20.0, 1), lambda ivp: ivp.crossing(2, 20.0, 1, 64)), (("crossing", 1, -1.0, -1), lambda ivp: ivp.crossing(1, -1.0, -1, 64)),
            (("hyperplane", [0.0, 0.0], [1.0, -1.0], [0, 1], 0), lambda ivp: ivp.hyperplane_crossing([0.0, 0.0], [1.0, -1.0], [0, 1], 0, 64)),
            (("hyperplane", [1.0, 2.0, 20.0], [0.3, -0.2, 1.0], [0, 1, 2], 1), lambda ivp: ivp.hyperplane_crossing([1.0, 2.0, 20.0], [0.3, -0.2, 1.0], [0, 1, 2], 1, 64)),
            (("hyperplane", [25.0], [-2.0], [2], -1), lambda ivp: ivp.hyperplane_crossing([25.0], [-2.0], [2], -1, 64)))
    for rec, setter in recs:
        for meth in ("dopri5", "dop853", "rkf45", "rkv655e", "rkv878e", "rk4"):
            if meth in ("dopri5", "dop853"):
                p, m = pr.solve_dp(lz, meth, 0.0, 8.0, [1.0, 1.0, 1.0], rtol=1e-7, atol=1e-8, recorder=rec), getattr(E, meth)().rtol(1e-7).atol(1e-8)
            elif meth == "rk4":
                p, m = pr.solve_fixed(lz, "rk4", 0.02, 0.0, 8.0, [1.0, 1.0, 1.0], recorder=rec), E.rk4(0.02)
            else:
                p, m = pr.solve_adaptive(lz, meth, 0.0, 8.0, [1.0, 1.0, 1.0], rtol=1e-7, atol=1e-8, recorder=rec), getattr(E, meth)().rtol(1e-7).atol(1e-8)
            c = ob.oracle_solve(setter(deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 8.0, [[1.0, 1.0, 1.0]])).method(m))
            sol = c[0]
            assert len(p["rows"]) == len(sol.t) and p["evals"] == c.evals[0]
            assert _same_bits([r[0] for r in p["rows"]], sol.t) and _same_bits([r[1] for r in p["rows"]], sol.y)
            if rec[0] == "default":
                assert len(sol.t) == c.accepted[0] + 1 and sol.t[0] == 0.0 and sol.t[-1] == 8.0
            if rec == ("dense", 3):
                assert len(sol.t) == 3 * c.accepted[0] + 1
    # crossings of cos(t) through 0 (examples/ode/08 pattern): pi/2 + k pi; direction filter
    for d, want in ((0, [0.5, 1.5, 2.5]), (1, [1.5]), (-1, [0.5, 2.5])):
        sol = ob.oracle_solve(deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 10.0, [[1.0, 0.0]]).crossing(0, 0.0, d, 16)
                              .method(E.dopri5().rtol(1e-10).atol(1e-10)))[0]
        np.testing.assert_allclose(sol.t / np.pi, want, atol=1e-8)
        assert np.abs(sol.y[:, 0]).max() < 1e-12


def _solve_py(f, meth, t0, tf, y0, tol, **kw):
    if meth in ("dopri5", "dop853"):
        return pr.solve_dp(f, meth, t0, tf, y0, rtol=tol, atol=tol, **kw)
    if meth == "rk4":
        kw.pop("even", None)
        return pr.solve_fixed(f, "rk4", math.copysign(0.01, tf - t0), t0, tf, y0, **kw)
    return pr.solve_adaptive(f, meth, t0, tf, y0, rtol=tol, atol=tol, **kw)


def _method(meth, tol, t0=0.0, tf=1.0):
    return E.rk4(math.copysign(0.01, tf - t0)) if meth == "rk4" else getattr(E, meth)().rtol(tol).atol(tol)


def test_event_restatements_agree_bitwise():
    """IVP::event (EventWrappedSolout + Brent-Dekker, src/solout/event.rs:300-470): C++ oracle vs the independent Python
    restatement, bit for bit, on the patterns of the reference's own examples and tests; event times against closed forms."""
    # examples/ode/03_logistic_growth: even(2.0).event(y - 0.9 m).terminal(), dop853 rtol = atol = 1e-12
    k, mm = 1.0, 10.0
    f = lambda t, y: [k * y[0] * (1.0 - y[0] / mm)]
    g = lambda t, y: y[0] - 0.9 * mm
    for meth in ("dop853", "dopri5", "rkv655e", "cash_karp"):
        p = _solve_py(f, meth, 0.0, 10.0, [1.0], 1e-12, even=2.0, event=(g, 0, 1)) if meth in ("dop853", "dopri5") else None
        c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.LogisticEquation(k, mm), 0.0, 10.0, [[1.0]]).even(2.0)
                            .event(deb.LinearEvent(-0.9 * mm, 0.0, [1.0]), terminate=1).method(_method(meth, 1e-12)))
        sol = c[0]
        assert sol.status == "Interrupted" and sol.t[:3].tolist() == [0.0, 2.0, 4.0] and len(sol.t) == 4
        assert abs(sol.t[3] - math.log(81.0)) < 1e-8 and abs(sol.y[3, 0] - 9.0) < 1e-9  # y = m / (1 + 9 e^-t) = 0.9 m  at  t = ln 81
        assert 4.39 < c.t_final[0] < 6.0  # the step that contains the event is kept (solver state), then Interrupted
        if p is not None:
            assert p["status"] == "Interrupted" and p["evals"] == c.evals[0] and p["t"] == c.t_final[0]
            assert _same_bits([r[0] for r in p["rows"]], sol.t) and _same_bits([r[1] for r in p["rows"]], sol.y)
    # tests/ode/errors.rs:12-28: y' = y with the terminal event t - 10 on a plain solve() (every step + the event row)
    for meth in ("dopri5", "dop853", "rkf45", "rkv878e", "rk4"):
        p = _solve_py(lambda t, y: [y[0]], meth, 0.0, 20.0, [1.0], 1e-6, recorder=("default",), event=(lambda t, y: t - 10.0, 0, 1))
        c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.ExponentialGrowth(1.0), 0.0, 20.0, [[1.0]])
                            .event(deb.LinearEvent(-10.0, 1.0, [0.0]), terminate=1, max_event_rows=2500).method(_method(meth, 1e-6, 0.0, 20.0)))
        sol = c[0]
        assert sol.status == p["status"] == "Interrupted" and abs(sol.t[-1] - 10.0) <= 1e-11 and c.t_final[0] >= 10.0
        assert len(sol.t) == c.accepted[0] + 2 or (len(sol.t) == c.accepted[0] + 1 and sol.t[-2] != 10.0)  # t0 + steps (+ event row unless it coincides)
        assert _same_bits([r[0] for r in p["rows"]], sol.t) and _same_bits([r[1] for r in p["rows"]], sol.y)
    # non-terminal and counted events on t_eval: zero crossings of cos t, direction filter, terminate_after(3)
    hf = pr.harmonic(1.0)
    te = [1.0, 2.0, 4.0, 7.5, 9.0]
    for meth in ("dopri5", "dop853", "rkf45", "rkv766e", "rk4"):
        for direction, term, want in ((0, 0, [0.5, 1.5, 2.5]), (1, 0, [1.5]), (-1, 0, [0.5, 2.5]), (0, 3, [0.5, 1.5, 2.5]), (0, 2, [0.5, 1.5])):
            p = _solve_py(hf, meth, 0.0, 10.0, [1.0, 0.0], 1e-10, t_eval=te, event=(lambda t, y: y[0], direction, term))
            c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 10.0, [[1.0, 0.0]]).t_eval(te)
                                .event(deb.LinearEvent(0.0, 0.0, [1.0, 0.0]), direction, term or None).method(_method(meth, 1e-10, 0.0, 10.0)))
            sol = c[0]
            assert sol.status == p["status"] == ("Interrupted" if term else "Complete")
            assert _same_bits([r[0] for r in p["rows"]], sol.t) and _same_bits([r[1] for r in p["rows"]], sol.y)
            ev_t = np.array([t for t in sol.t if t not in te])
            np.testing.assert_allclose(ev_t / np.pi, want, atol=2e-6)
    # backward time: the reference's interpolate() bounds test rejects every interior point, so Brent-Dekker gives up
    # (`.ok()?`) unless an end point is an exact zero -- as written
    p = pr.solve_dp(hf, "dopri5", 10.0, 0.0, [1.0, 0.0], rtol=1e-8, atol=1e-8, t_eval=[5.0], event=(lambda t, y: y[0], 0, 0))
    c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 10.0, 0.0, [[1.0, 0.0]]).t_eval([5.0])
                        .event(deb.LinearEvent(0.0, 0.0, [1.0, 0.0])).method(E.dopri5().rtol(1e-8).atol(1e-8)))
    assert [r[0] for r in p["rows"]] == c[0].t.tolist() == [5.0]


def test_even_solout_restatements_agree_bitwise():
    """EvenSolout (src/solout/even.rs:69-199): C++ oracle vs the independent Python restatement, bit for bit; and the
    documented output shape (t0 first, tf last, spacing dt)."""
    for meth, dt, tf in (("dopri5", 0.3, 2.0), ("dop853", 0.5, 2.0), ("dopri5", 0.25, 10.0), ("dop853", 0.7, 5.0)):
        p = pr.solve_dp(pr.lorenz(10.0, 28.0, 8.0 / 3.0), meth, 0.0, tf, [1.0, 1.0, 1.0], rtol=1e-8, atol=1e-8, even=dt)
        c = ob.oracle_solve(deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, tf, [[1.0, 1.0, 1.0]]).even(dt)
                            .method(getattr(E, meth)().rtol(1e-8).atol(1e-8)))
        sol = c[0]
        assert [r[0] for r in p["rows"]] == sol.t.tolist()
        assert _same_bits([r[1] for r in p["rows"]], sol.y)
        assert sol.t[0] == 0.0 and sol.t[-1] == tf and np.allclose(np.diff(sol.t[:-1]), dt)


# ------------------------------------------------------------------------------------------ bit-level pin against the crate
REFERENCE_BITS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_bits.json")


def _unhex(v):
    if isinstance(v, list):
        return [_unhex(x) for x in v]
    return np.array([int(v, 16)], dtype=np.uint64).view(np.float64)[0]


def _replay_case(case):
    """Run one case of tests/golden/reference_bits.json (written by the Rust program
    oracle/crate_pin/src/main.rs from the real crate) through the oracle."""
    y0 = np.array(_unhex(case["y0"]), dtype=np.float64).reshape(len(case["results"]), -1)
    if "params_per_traj" in case:
        prm = np.array(_unhex(case["params_per_traj"]))
    else:
        prm = _unhex(case["params"])
    sysm = {"lorenz": lambda p: deb.LorenzSystem(*p), "van_der_pol": lambda p: deb.VanDerPolOscillator(p),
            "exponential": lambda p: deb.ExponentialGrowth(*p)}[case["system"]](prm)
    m = getattr(E, case["method"])
    meth = m(_unhex(case["h0"])) if "h0" in case else m()
    for opt in ("rtol", "atol", "h_max"):
        if opt in case:
            getattr(meth, opt)(_unhex(case[opt]))
    if "max_steps" in case:
        meth.max_steps(case["max_steps"])
    ivp = deb.EnsembleIVP.ode(sysm, _unhex(case["t0"]), _unhex(case["tf"]), y0).t_eval(_unhex(case["t_eval"])).method(meth)
    return ob.oracle_solve(ivp)


def _assert_case_matches(case, got):
    for i, want in enumerate(case["results"]):
        status = deb._STATUS_NAME[int(got.status[i])]
        assert status == want["status"], (case["name"], i, status, want["status"])
        if want["status"] in ("Complete", "Interrupted"):
            assert (int(got.accepted[i]), int(got.rejected[i]), int(got.evals[i])) == (want["accepted"], want["rejected"], want["evals"]), (case["name"], i)
            t = np.array(_unhex(want["t"]))
            y = np.array(_unhex(want["y"])).reshape(len(t), -1)
            sol = got[i]
            assert _same_bits(sol.t, t) and _same_bits(sol.y, y), (case["name"], i)
        elif "t_final" in want:
            assert _same_bits([got.t_final[i]], [_unhex(want["t_final"])]) and _same_bits(got.y_final[i], _unhex(want["y_final"])), (case["name"], i)


def test_oracle_matches_the_crate_bit_for_bit():
    """Consumes tests/golden/reference_bits.json when it exists (tools/pin_oracle_against_crate.sh, needs cargo).  Without the
    file the oracle stays 'parity unpinned' at the bit level: two independent restatements and the reference's own goldens pin
    it, the crate's bits do not."""
    if not os.path.exists(REFERENCE_BITS):
        pytest.skip("tests/golden/reference_bits.json not generated: no Rust toolchain in this image (tools/pin_oracle_against_crate.sh)")
    doc = json.load(open(REFERENCE_BITS))
    assert doc["crate"] == "differential-equations"
    for case in doc["cases"]:
        _assert_case_matches(case, _replay_case(case))


def test_reference_bits_replay_machinery_round_trips():
    """The replay code above must not rot while no toolchain is around: build a document of the same shape from the oracle
    itself (hex bit patterns, Error variants included) and replay it."""
    def hx(v):
        if isinstance(v, (list, tuple, np.ndarray)):
            return [hx(x) for x in v]
        return "%016x" % np.array([v], dtype=np.float64).view(np.uint64)[0]
    te = [float(i) for i in range(1, 6)]
    y0 = ob.lorenz_ensemble_y0(3)
    cases = [dict(name="lorenz_dopri5", system="lorenz", params=hx([10.0, 28.0, 8.0 / 3.0]), method="dopri5", rtol=hx(1e-8), t0=hx(0.0), tf=hx(5.0),
                  t_eval=hx(te), y0=hx(y0.tolist())),
             dict(name="stiffness", system="exponential", params=hx([1.0000001]), method="dopri5", h_max=hx(0.002), max_steps=100000, t0=hx(0.0),
                  tf=hx(10.0), t_eval=hx([1.0, 2.0]), y0=hx([[1.0]])),
             dict(name="rk4", system="lorenz", params=hx([10.0, 28.0, 8.0 / 3.0]), method="rk4", h0=hx(0.01), t0=hx(0.0), tf=hx(1.0), t_eval=hx([0.5, 1.0]),
                  y0=hx([[1.0, 1.0, 1.0]]))]
    for case in cases:
        case["results"] = [None] * len(case["y0"])
        got = _replay_case(case)
        res = []
        for i in range(len(case["y0"])):
            st = deb._STATUS_NAME[int(got.status[i])]
            if st == "Complete":
                sol = got[i]
                res.append(dict(status=st, accepted=int(got.accepted[i]), rejected=int(got.rejected[i]), evals=int(got.evals[i]), t=hx(sol.t), y=hx(sol.y.tolist())))
            else:
                res.append(dict(status=st, t_final=hx(got.t_final[i]), y_final=hx(got.y_final[i].tolist())))
        case["results"] = res
        _assert_case_matches(json.loads(json.dumps(case)), _replay_case(case))
    assert cases[1]["results"][0]["status"] == "Stiffness"


# ------------------------------------------------------------------------------------------ an output of the real crate
def _docs_example(doc):
    ev = deb.LinearEvent(*doc["event"]["linear"][:2], doc["event"]["linear"][2:])
    return (deb.EnsembleIVP.ode(deb.LogisticEquation(*doc["params"]), doc["t0"], doc["tf"], [[doc["y0"]]]).even(doc["even_dt"])
            .event(ev, doc["event"]["direction"], terminate=doc["event"]["terminate"])
            .method(getattr(E, doc["method"])().rtol(doc["rtol"]).atol(doc["atol"])))


def check_docs_example(sol_set, doc):
    want = doc["expected"]
    sol = sol_set[0]
    assert sol.status == want["status"]
    assert (sol.evals.function, sol.steps.accepted, sol.steps.rejected) == (want["function_evaluations"], want["accepted"], want["rejected"])
    assert sol.steps.accepted + sol.steps.rejected == want["steps_total"]
    got = [[round(float(t), 4), round(float(y[0]), 4)] for t, y in zip(sol.t, sol.y)]
    assert got == want["rows_4_decimals"]


def test_oracle_reproduces_the_output_printed_in_the_crates_documentation():
    """docs/ode.md prints what the crate itself produced for its logistic-growth example: 325 function evaluations, 20 accepted
    and 2 rejected steps, six rows.  Step counts are integers that depend on every detail of h_init, the DOP853 stages, the
    error norm, the controller, EvenSolout and the Brent-Dekker event location: the C++ oracle must hit them exactly."""
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_docs_output.json")))
    check_docs_example(ob.oracle_solve(_docs_example(doc)), doc)
