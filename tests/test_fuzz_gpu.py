"""Randomised parity sweep (GPU): many random problem descriptions -- system, method, direction, tolerances (scalar and
vector), h0 / h_min / h_max / max_steps / safety / scale bounds / max_rejects, t_eval sets or even(dt) -- each solved on
the GPU and by the CPU oracle; every output must agree bit for bit.  Seeds are fixed: failures are reproducible."""
import importlib

import numpy as np
import pytest

import oracle_binding as ob

deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta
pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


SYSTEMS = [
    ("exponential", lambda r, n: deb.ExponentialGrowth(r.uniform(-1.5, 1.0, n)), 1, lambda r, n: r.uniform(0.5, 2.0, (n, 1))),
    ("linear", lambda r, n: deb.LinearEquation(r.uniform(-1, 1, n), r.uniform(-1.0, 0.5, n)), 1, lambda r, n: r.uniform(-1, 1, (n, 1))),
    ("harmonic", lambda r, n: deb.HarmonicOscillator(r.uniform(0.5, 4.0, n)), 2, lambda r, n: r.uniform(-1, 1, (n, 2))),
    ("logistic", lambda r, n: deb.LogisticEquation(r.uniform(0.5, 2.0, n), r.uniform(5.0, 20.0, n)), 1, lambda r, n: r.uniform(0.1, 1.0, (n, 1))),
    ("vdp", lambda r, n: deb.VanDerPolOscillator(r.uniform(0.1, 6.0, n)), 2, lambda r, n: np.array([2.0, 0.0]) + r.uniform(-0.3, 0.3, (n, 2))),
    ("lorenz", lambda r, n: deb.LorenzSystem(10.0, r.uniform(20.0, 35.0, n), 8.0 / 3.0), 3, lambda r, n: 1.0 + r.uniform(-0.5, 0.5, (n, 3))),
    ("brusselator", lambda r, n: deb.BrusselatorSystem(1.0, r.uniform(1.5, 3.0, n)), 2, lambda r, n: np.array([1.5, 3.0]) + r.uniform(-0.2, 0.2, (n, 2))),
    ("lorenz_shared", lambda r, n: deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 3, lambda r, n: 1.0 + r.uniform(-0.5, 0.5, (n, 3))),
    # finite-time blow-up (y' = y + y^2): StepSize / MaxSteps errors and non-finite intermediate values
    ("blowup", lambda r, n: deb.LogisticEquation(1.0, -1.0), 1, lambda r, n: r.uniform(0.3, 3.0, (n, 1))),
    # stiff: the explicit methods crawl, reject, and (adaptive family with a small max_rejects) report Stiffness
    ("robertson", lambda r, n: deb.RobertsonProblem(), 3, lambda r, n: np.array([1.0, 0.0, 0.0]) + r.uniform(0.0, 1e-3, (n, 3))),
]
ADAPTIVE = ["dopri5", "dop853", "rkf45", "cash_karp", "rkv655e", "rkv656e", "rkv766e", "rkv767e", "rkv877e", "rkv878e", "rkv988e", "rkv989e"]
FIXED = ["euler", "midpoint", "heun", "ralston", "ssp_rk3", "rk4", "three_eighths"]


def random_case(seed):
    r = np.random.default_rng(seed)
    n = int(r.integers(33, 200))
    name, mk_sys, dim, mk_y0 = SYSTEMS[int(r.integers(len(SYSTEMS)))]
    sysm, y0 = mk_sys(r, n), mk_y0(r, n)
    span = float(r.uniform(0.5, 6.0))
    t0 = float(r.uniform(-2.0, 2.0))
    backward = bool(r.random() < 0.3) and name not in ("lorenz", "lorenz_shared", "vdp", "brusselator", "robertson")  # backward blows up for dissipative systems
    tf = t0 - span if backward else t0 + span
    d = -1.0 if backward else 1.0
    adaptive = bool(r.random() < 0.7)
    u = r.random()  # recorder kind; per-step recorders are compiled at first use, so they get a small method set
    per_step = 0.25 <= u < 0.5
    has_event = bool(r.random() < 0.15)  # IVP::event(..) around whichever recorder; also compiled at first use
    per_step = per_step or has_event     # (only restricts the method pool below)
    if adaptive:
        pool = ["dopri5", "rkf45"] if per_step else ADAPTIVE
        ctor = pool[int(r.integers(len(pool)))]
        m = getattr(E, ctor)()
        if r.random() < 0.5:
            m.rtol(float(10.0 ** r.uniform(-10, -4))).atol(float(10.0 ** r.uniform(-11, -5)))
        elif r.random() < 0.5:
            m.rtol((10.0 ** r.uniform(-9, -4, dim)).tolist()).atol((10.0 ** r.uniform(-10, -5, dim)).tolist())
        if r.random() < 0.3:
            m.h0(d * span * float(10.0 ** r.uniform(-4, -1)))
        if r.random() < 0.3:
            m.h_max(span * float(r.uniform(0.005, 0.2)))
        if r.random() < 0.2:
            m.h_min(span * 1e-7)
        if r.random() < 0.3:
            m.max_steps(int(r.integers(20, 400)))
        if r.random() < 0.3:
            m.safety_factor(float(r.uniform(0.7, 0.95))).min_scale(float(r.uniform(0.1, 0.5))).max_scale(float(r.uniform(2.0, 10.0)))
        if r.random() < 0.2:
            m.max_rejects(int(r.integers(1, 6)))
    else:
        pool = ["rk4"] if per_step else FIXED
        ctor = pool[int(r.integers(len(pool)))]
        m = getattr(E, ctor)(d * span / float(r.integers(20, 400)) if r.random() < 0.85 else 0.0)
        if r.random() < 0.2:
            m.max_steps(int(r.integers(10, 100)))
    ivp = deb.EnsembleIVP.ode(sysm, t0, tf, y0)
    if u < 0.25:
        ivp.even(span / float(r.uniform(2.5, 40.0)))
    elif u < 0.5:  # per-step recorders (capacity sometimes too small on purpose)
        cap = int(r.integers(5, 600))
        which = int(r.integers(4))
        if which == 0:
            ivp.every_step(cap)
        elif which == 1:
            ivp.dense(int(r.integers(0, 5)), cap)
        elif which == 2:
            ivp.crossing(int(r.integers(dim)), float(np.median(y0[:, 0]) + r.uniform(-0.5, 0.5)), int(r.integers(-1, 2)), int(r.integers(1, 40)))
        else:
            comps = sorted(r.choice(dim, size=int(r.integers(1, dim + 1)), replace=False).tolist())
            ivp.hyperplane_crossing((np.median(y0[:, comps], axis=0) + r.uniform(-0.3, 0.3, len(comps))).tolist(),
                                    r.uniform(-1.0, 1.0, len(comps)).tolist(), comps, int(r.integers(-1, 2)), int(r.integers(1, 40)))
    else:
        k = int(r.integers(0, 12))
        pts = r.uniform(min(t0, tf) - 0.2 * span, max(t0, tf) + 0.2 * span, k).tolist()
        if r.random() < 0.5:
            pts += [t0, tf, 0.5 * (t0 + tf), tf]
        r.shuffle(pts)
        ivp.t_eval(pts)
    if has_event:
        coef = r.uniform(-1.0, 1.0, dim)
        coef[int(r.integers(dim))] = 1.0
        c0 = -float(np.median(y0 @ coef)) + float(r.uniform(-0.3, 0.3))
        ct = float(r.uniform(-0.2, 0.2)) if r.random() < 0.3 else 0.0
        ivp.event(deb.LinearEvent(c0, ct, coef.tolist()), int(r.integers(-1, 2)), int(r.integers(1, 4)) if r.random() < 0.5 else None,
                  max_event_rows=int(r.integers(2, 300)))
    return ivp.method(m), f"seed {seed}: {name} n={n} {ctor} t0={t0:.3g} tf={tf:.3g}"


@pytest.mark.parametrize("block", range(8))
def test_random_problem_descriptions_bitwise(block):
    for seed in range(1000 + 40 * block, 1000 + 40 * (block + 1)):
        ivp, label = random_case(seed)
        ivp2, _ = random_case(seed)
        g, c = ivp.solve(), ob.oracle_solve(ivp2)
        for name in ("status", "accepted", "rejected", "evals", "n_emitted"):
            assert np.array_equal(getattr(g, name), getattr(c, name)), (label, name)
        assert np.array_equal(g.t_rows, c.t_rows), label
        fin = np.isfinite(c.y_final).all(axis=1)  # NaN payloads may differ; every finite result must match exactly
        assert np.array_equal(np.isfinite(g.y_final).all(axis=1), fin), label
        assert np.array_equal(bits(g.t_final), bits(c.t_final)), label
        assert np.array_equal(bits(g.y_final[fin]), bits(c.y_final[fin])), label
        m = (np.arange(g.y_eval.shape[1])[None, :] < g.n_emitted[:, None]) & fin[:, None]
        assert np.array_equal(bits(g.y_eval)[m], bits(c.y_eval)[m]), label
        assert (g.t_out is None) == (c.t_out is None), label
        if g.t_out is not None:
            assert np.array_equal(bits(g.t_out)[m], bits(c.t_out)[m]), label
