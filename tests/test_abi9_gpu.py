"""GPU tests of what ABI 9 added: the streamed HOST pipeline (completion watermark), the device list, the row-major
layout, fused ensemble statistics, the step-size filter, user-defined SDEs, the forward-sensitivity generator, the fixed-step
schedule with a double clip at tf, struct_size compatibility.  Everything is compared bitwise with the CPU oracle or with
an equivalent call of the library itself.
"""
import ctypes as C
import importlib

import numpy as np
import pytest

import oracle_binding as ob
from test_parity_gpu import assert_same_solution, bits, lorenz

deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta
pytestmark = pytest.mark.gpu


def n_gpus():
    return deb.load_library().deb_device_count()


# ------------------------------------------------------------------------------------------ fixed-step schedule
DOUBLE_CLIP = (-22246.572886668695, 976.800691167858, 7199.779502015325)  # the clip at tf fires twice: 5 steps, the last one -1.1e-13... wide


@pytest.mark.parametrize("ctor", ["euler", "rk4", "ssp_rk3"])
def test_fixed_step_schedule_with_two_clipped_steps(ctor):
    """t + (tf - t) can miss tf by one ulp when |tf| is large: the reference then takes one more tiny step.  The host-planned
    schedule must carry both clipped steps (round 1 carried only the last one and ended 5575 time units past tf)."""
    t0, tf, h = DOUBLE_CLIP
    y0 = np.linspace(0.5, 1.5, 70).reshape(-1, 1)
    te = [t0, -20000.0, -5000.0, 0.0, 900.0, tf]
    def prob():
        return deb.EnsembleIVP.ode(deb.LinearEquation(1e-5, -1e-4), t0, tf, y0).t_eval(te).method(getattr(E, ctor)(h))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c)
    assert (g.status == 0).all() and (g.accepted == 5).all() and np.all(np.abs(g.t_final - tf) < 1e-9)


def test_fixed_step_schedule_random_sweep_matches_oracle():
    """Random (t0, tf, h) with t0 << 0 < tf and coarse h: the class of cases where the second clip shows up."""
    rng = np.random.default_rng(11)
    y0 = np.array([[1.0], [2.0]])
    seen_extra = 0
    for _ in range(60):
        t0 = -rng.uniform(1e3, 1e5)
        tf = rng.uniform(1e2, 1e4)
        h = (tf - t0) / rng.uniform(2.2, 9.7)
        def prob():
            return deb.EnsembleIVP.ode(deb.ExponentialGrowth(-1e-5), t0, tf, y0).t_eval([t0, 0.0, tf]).method(E.rk4(h))
        g, c = prob().solve(), ob.oracle_solve(prob())
        assert_same_solution(g, c)
        seen_extra += int(g.accepted[0] != int(np.ceil((tf - t0) / h)))
    assert seen_extra > 0, "the sweep never hit a doubly clipped schedule"


def test_sde_schedule_with_two_clipped_steps():
    # the second clipped step is +4.5e-13 here (with DOUBLE_CLIP it is negative and sqrt(h) is NaN, in the reference too)
    t0, tf, h = -94916.29526658714, 3187.131374903806, 18251.97563541899
    y0 = np.linspace(0.5, 1.5, 64)
    def prob():
        return deb.EnsembleIVP.sde(deb.OrnsteinUhlenbeck(1e-4, 1.0, 1e-3), t0, tf, y0, seed=5).t_eval([t0, 0.0, tf]).method(E.euler(h))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert np.array_equal(g.accepted, c.accepted) and (g.accepted == 7).all() and np.array_equal(g.status, c.status)
    assert np.array_equal(g.n_emitted, c.n_emitted) and np.isfinite(g.y_final).all()
    np.testing.assert_allclose(g.y_final, c.y_final, rtol=1e-12)
    np.testing.assert_allclose(g.y_eval, c.y_eval, rtol=1e-12)
    assert np.array_equal(bits(g.t_final), bits(c.t_final))


# ------------------------------------------------------------------------------------------ streamed HOST pipeline
def vdp_problem(n=5003):
    mu = np.linspace(0.1, 8.0, n)
    y0 = np.tile([2.0, 0.0], (n, 1)) + ob.splitmix64_uniform(3, 2 * n).reshape(n, 2) * 0.1
    def prob():
        return (deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 0.0, 12.0, y0).t_eval(np.linspace(0.0, 12.0, 7))
                .method(E.dopri5().rtol(1e-7).max_steps(158)))
    return prob


@pytest.mark.parametrize("shift,chunk", [(4, 0), (6, 700), (12, 0), (5, 64)])
def test_watermark_streaming_is_transparent(monkeypatch, shift, chunk):
    """The HOST call launches one persistent kernel per chunk and copies finished 2^shift-trajectory blocks out while it runs.
    Block size and chunking must not change a bit -- including a ragged last block, rows that stop early (MaxSteps) and
    per-trajectory parameters."""
    prob = vdp_problem()
    ref = ob.oracle_solve(prob())
    monkeypatch.setenv("DEB_WM_SHIFT", str(shift))
    if chunk:
        monkeypatch.setenv("DEB_HOST_CHUNK", str(chunk))
    got = prob().solve()
    assert (got.status != 0).any() and (got.status == 0).any()
    assert_same_solution(got, ref)
    assert got.gpu_launches >= 1


@pytest.mark.parametrize("ctor", ["rk4", "dop853", "rkv766e"])
def test_watermark_streaming_other_kernels(monkeypatch, ctor):
    """The fixed-step kernel and the immediate-emission (non-parked) adaptive kernels publish the watermark too."""
    monkeypatch.setenv("DEB_WM_SHIFT", "5")
    y0 = ob.lorenz_ensemble_y0(1500, seed=17)
    def prob():
        m = E.rk4(0.01) if ctor == "rk4" else getattr(E, ctor)().rtol(1e-8)
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 2.0, y0).t_eval(np.linspace(0.0, 2.0, 9)).method(m)
    assert_same_solution(prob().solve(), ob.oracle_solve(prob()))


def test_row_staging_all_dims_and_row_counts():
    """RowStage groups 4/gcd(dim,4) rows per store: dim 1, 2, 3 with row counts that are / are not multiples of the group
    (aligned vector stores or the scalar fall-back), t0 emitted or not, trajectories that stop mid-group."""
    rng = np.random.default_rng(5)
    for sysm, y0 in ((deb.ExponentialGrowth(-0.7), 1.0 + rng.uniform(0, 1, (300, 1))),
                     (deb.HarmonicOscillator(2.0), rng.uniform(-1, 1, (300, 2))),
                     (lorenz(), ob.lorenz_ensemble_y0(300, seed=8))):
        for n_rows in (1, 2, 3, 4, 5, 7, 8, 100):
            for first in (0.0, 0.01):
                te = np.linspace(first, 3.0, n_rows)
                for meth in (E.dopri5().rtol(1e-7), E.dopri5().rtol(1e-7).max_steps(25), E.rk4(0.05), E.dop853()):
                    def prob():
                        return deb.EnsembleIVP.ode(sysm, 0.0, 3.0, y0).t_eval(te).method(meth)
                    assert_same_solution(prob().solve(), ob.oracle_solve(prob()))


def test_row_major_layout_is_the_transpose():
    y0 = ob.lorenz_ensemble_y0(3001, seed=23)
    te = np.linspace(0.0, 2.0, 11)
    def prob():
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 2.0, y0).t_eval(te).method(E.dopri5().rtol(1e-8))
    a = prob().solve()
    b = prob().layout(deb.DEB_LAYOUT_ROW_MAJOR).solve()
    assert b.y_eval_row_major.shape == (11, 3, 3001)
    assert np.array_equal(bits(a.y_eval), bits(b.y_eval))
    assert np.array_equal(bits(a.y_final), bits(b.y_final)) and np.array_equal(a.accepted, b.accepted)


def test_fused_statistics_match_the_rows(monkeypatch):
    """deb_result.stats_sums / stats_counts: reduced on the device while the rows are resident; equal to the sums over the
    returned rows (to rounding: the summation order differs) and to deb_ensemble_stats on the same rows; chunking adds the
    chunks' sums in a fixed order."""
    prob = vdp_problem(4100)
    g = prob().with_stats().solve()
    m = (np.arange(7)[None, :] < g.n_emitted[:, None])
    y = np.where(m[:, :, None], g.y_eval, 0.0)
    np.testing.assert_allclose(g.stats_sums[:, :, 0], y.sum(axis=0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g.stats_sums[:, :, 1], (y * y).sum(axis=0), rtol=1e-12, atol=1e-12)
    assert np.array_equal(g.stats_counts, m.sum(axis=0))
    lib = deb.load_library()
    sums = np.zeros((7, 2, 2)); counts = np.zeros(7, np.int64)
    rc = lib.deb_ensemble_stats(g.y_eval.ctypes.data, g.n_emitted.ctypes.data, 4100, 7, 2, sums.ctypes.data, counts.ctypes.data, 0, deb.DEB_MEM_HOST, None)
    assert rc == 0, lib.deb_last_error()
    assert np.array_equal(bits(sums), bits(g.stats_sums)) and np.array_equal(counts, g.stats_counts)
    monkeypatch.setenv("DEB_WM_SHIFT", "6")
    monkeypatch.setenv("DEB_HOST_CHUNK", "1000")
    g2 = prob().with_stats().solve()
    np.testing.assert_allclose(g2.stats_sums, g.stats_sums, rtol=1e-12, atol=1e-12)
    assert np.array_equal(g2.stats_counts, g.stats_counts)


# ------------------------------------------------------------------------------------------ device list
@pytest.mark.parametrize("shift", [5, 12])
def test_device_list_equals_one_device(monkeypatch, shift):
    """deb_ode_problem.devices: the ensemble split over every visible GPU inside one call (block-cyclic), results and statistics
    exactly as on one device.  (On a one-GPU box the list has one entry: the same code path with G = 1.)"""
    monkeypatch.setenv("DEB_WM_SHIFT", str(shift))
    devs = list(range(max(1, min(n_gpus(), 8))))
    prob = vdp_problem(20011 if shift == 12 else 5003)
    one = prob().with_stats().solve()
    many = prob().devices(devs).with_stats().solve()
    assert_same_solution(many, one)
    np.testing.assert_allclose(many.stats_sums, one.stats_sums, rtol=1e-12, atol=1e-12)
    assert np.array_equal(many.stats_counts, one.stats_counts)
    assert_same_solution(many, ob.oracle_solve(prob()))
    if len(devs) > 1:
        rm = prob().devices(devs).layout(deb.DEB_LAYOUT_ROW_MAJOR).solve()
        m = np.arange(one.y_eval.shape[1])[None, :] < one.n_emitted[:, None]  # rows beyond n_emitted are unspecified
        assert np.array_equal(bits(rm.y_eval)[m], bits(one.y_eval)[m])


def test_device_list_statistics_without_nccl():
    """When no libnccl can be loaded the per-device sums are added on the host (DEB_NO_NCCL=1 takes that path): the same
    statistics to rounding.  A fresh process, because the library looks NCCL up once."""
    if n_gpus() < 2:
        pytest.skip("needs two devices")
    import os, subprocess, sys
    code = (
        "import importlib, sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "deb = importlib.import_module('differential-equations_b200'); E = deb.ExplicitRungeKutta\n"
        "mu = np.linspace(0.5, 3.0, 9001)\n"
        "def prob():\n"
        "    return (deb.EnsembleIVP.ode(deb.VanDerPolOscillator(mu), 0.0, 3.0, np.tile([2.0, 0.0], (9001, 1))).t_eval(np.linspace(0.0, 3.0, 7))\n"
        "            .method(E.dopri5().rtol(1e-8)))\n"
        "one = prob().with_stats().solve(); two = prob().devices([0, 1]).with_stats().solve()\n"
        "np.testing.assert_allclose(two.stats_sums, one.stats_sums, rtol=1e-12, atol=1e-12)\n"
        "assert np.array_equal(two.stats_counts, one.stats_counts)\n"
        "assert np.array_equal(two.y_eval.view(np.uint64), one.y_eval.view(np.uint64))\n"
        "print('ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DEB_NO_NCCL="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_device_list_is_validated():
    prob = vdp_problem(100)
    with pytest.raises(ValueError, match="duplicate"):
        prob().devices([0, 0]).solve()
    with pytest.raises(ValueError, match="out of range"):
        prob().devices([0, 99]).solve()


# ------------------------------------------------------------------------------------------ step-size filter
@pytest.mark.parametrize("ctor,bits_kept", [("dopri5", 20), ("dopri5", 32), ("dop853", 24), ("rkf45", 16), ("rkv766e", 40)])
def test_mantissa_truncating_filter_bit_exact(ctor, bits_kept):
    """`.filter(|h| from_bits(h.to_bits() & MASK))`: applied at init, after every step-size update and in set_h (the clip at tf),
    as dormandprince/ordinary.rs:33,267,288 and adaptive/ordinary.rs:34,207,228 do."""
    y0 = ob.lorenz_ensemble_y0(500, seed=3)
    def prob(f=True):
        m = getattr(E, ctor)().rtol(1e-8)
        if f:
            m = m.filter_truncate_mantissa(bits_kept)
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 5.0, y0).t_eval(np.linspace(0.0, 5.0, 13)).method(m)
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert_same_solution(g, c)
    plain = prob(False).solve()
    assert not np.array_equal(bits(g.y_final), bits(plain.y_final)), "the filter changed nothing"
    # a fixed-step method never calls the hook
    def pf(f):
        m = E.rk4(0.013)
        if f:
            m = m.filter_truncate_mantissa(8)
        return deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0[:64]).method(m)
    assert_same_solution(pf(True).solve(), pf(False).solve())


# ------------------------------------------------------------------------------------------ user-defined SDEs
def test_user_defined_sde_equals_builtins_bitwise():
    """deb_define_sde: drift / diffusion / noise as text, compiled into the same kernel template -> the built-in OU, GBM and
    Heston results bit for bit, for Euler-Maruyama, RK drift stages and Milstein."""
    n = 4096
    te = np.linspace(0.0, 1.0, 6)
    ou = deb.sde_from_source(1, "dydt[0] = p[0] * (p[1] - y[0]);", "g[0] = p[2];", params=[0.7, 1.2, 0.3])
    gbm = deb.sde_from_source(1, "dydt[0] = p[0] * y[0];", "g[0] = p[1] * y[0];", params=[0.05, 0.2])
    heston = deb.sde_from_source(2, "dydt[0] = p[0] * y[0]; dydt[1] = p[1] * (p[2] - y[1]);",
                                 "g[0] = y[0] * sqrt(y[1]); g[1] = p[3] * sqrt(y[1]);", params=[0.1, 2.0, 0.04, 0.3, -0.7],
                                 noise_body="dw[1] = p[4] * dw[0] + sqrt(1.0 - p[4] * p[4]) * dw[1];")
    cases = [(ou, deb.OrnsteinUhlenbeck(0.7, 1.2, 0.3), np.full(n, 0.5)), (gbm, deb.GeometricBrownianMotion(0.05, 0.2), np.full(n, 100.0)),
             (heston, deb.HestonModel(0.1, 2.0, 0.04, 0.3, -0.7), np.tile([100.0, 0.04], (n, 1)))]
    for usr, builtin, y0 in cases:
        for m in (E.euler(0.01), E.rk4(0.01), deb.Milstein.new(0.01)):
            def prob(s):
                return deb.EnsembleIVP.sde(s, 0.0, 1.0, y0, seed=99, path_offset=7).t_eval(te).method(m)
            a, b = prob(usr).solve(), prob(builtin).solve()
            for name in ("status", "accepted", "evals", "n_emitted"):
                assert np.array_equal(getattr(a, name), getattr(b, name)), name
            assert np.array_equal(bits(a.y_final), bits(b.y_final)) and np.array_equal(bits(a.y_eval), bits(b.y_eval))


def test_user_defined_sde_compile_error_is_reported():
    bad = deb.sde_from_source(1, "dydt[0] = nonsense;", "g[0] = 1.0;", params=[1.0])
    with pytest.raises(ValueError, match="did not compile"):
        deb.EnsembleIVP.sde(bad, 0.0, 1.0, np.ones(8), seed=1).method(E.euler(0.1)).solve()


# ------------------------------------------------------------------------------------------ forward sensitivities
def test_generated_forward_sensitivity_system_equals_hand_written():
    """deb_define_ode_sensitivity builds S' = J_y S + J_p from the bodies of diff / jacobian / jacobian_p (forward.rs:82-115);
    Lorenz with all three parameters = the 12-dimensional system written out by hand in test_parity_gpu, bit for bit."""
    diff = "dydt[0] = p[0] * (y[1] - y[0]); dydt[1] = y[0] * (p[1] - y[2]) - y[1]; dydt[2] = y[0] * y[1] - p[2] * y[2];"
    jac = "J[0] = -p[0]; J[1] = p[0]; J[3] = p[1] - y[2]; J[4] = -1.0; J[5] = -y[0]; J[6] = y[1]; J[7] = y[0]; J[8] = -p[2];"
    jacp = "Jp[0] = y[1] - y[0]; Jp[4] = y[0]; Jp[8] = -y[2];"
    hand = """
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
    prm = [10.0, 28.0, 8.0 / 3.0]
    gen = deb.ode_sensitivity_from_source(3, diff, jac, jacp, prm)
    ref = deb.ode_from_source(12, hand, params=prm)
    assert gen.dim == 12
    y0 = np.zeros((64, 12))
    y0[:, :3] = ob.lorenz_ensemble_y0(64, seed=81)
    for meth in ("dopri5", "dop853", "rk4"):
        def prob(s):
            m = E.rk4(0.01) if meth == "rk4" else getattr(E, meth)().rtol(1e-8).atol(1e-8)
            return deb.EnsembleIVP.ode(s, 0.0, 2.0, y0).t_eval([0.0, 0.4, 1.1, 2.0]).method(m)
        assert_same_solution(prob(gen).solve(), prob(ref).solve())
    # a parameter sweep through the generated system: one parameter row per trajectory
    sweep = np.tile(prm, (64, 1))
    sweep[:, 1] = np.linspace(20.0, 30.0, 64)
    gen2 = deb.ode_sensitivity_from_source(3, diff, jac, jacp, sweep)
    ref2 = deb.OdeSystem(ref.system_id, 12, sweep)
    def prob2(s):
        return deb.EnsembleIVP.ode(s, 0.0, 1.0, y0).method(E.dopri5().rtol(1e-8))
    assert_same_solution(prob2(gen2).solve(), prob2(ref2).solve())


# ------------------------------------------------------------------------------------------ struct_size compatibility
def test_older_struct_size_is_accepted_and_newer_rejected():
    """Fields are only appended: an ABI-8 caller (struct ends at plane_normal / t_out) still works; a larger struct is refused."""
    lib = deb.load_library()
    y0 = ob.lorenz_ensemble_y0(64)
    ivp = deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0).t_eval([0.5, 1.0]).method(E.dopri5().rtol(1e-8))
    full = ivp.solve()
    P, res, arrs, t_sorted, keep = ivp.build_problem()
    P.struct_size = deb.OdeProblem.filter.offset       # sizeof(deb_ode_problem) of ABI 8
    res.struct_size = deb.Result.stats_sums.offset     # sizeof(deb_result) of ABI 8
    P.filter, P.filter_bits, P.n_devices = 1, 7, 5     # garbage beyond the declared size must not be read
    assert lib.deb_solve_ode(C.byref(P), C.byref(res)) == 0, lib.deb_last_error()
    assert np.array_equal(bits(arrs["y_final"]), bits(full.y_final)) and np.array_equal(bits(arrs["y_eval"]), bits(full.y_eval))
    P.struct_size = C.sizeof(deb.OdeProblem) + 8
    assert lib.deb_solve_ode(C.byref(P), C.byref(res)) == deb.DEB_ERR_BAD_ARG and b"struct_size" in lib.deb_last_error()


def test_heat_max_steps_zero_means_the_reference_default():
    lib = deb.load_library()
    u0 = np.sin(np.linspace(0.0, np.pi, 257))
    P, out, (tfin, steps, status), keep = deb.build_heat_problem(u0, 0.0, 1.0, 1e-4, E.rk4(0.01), 0.0, 1.0, ("dirichlet", 0.0), ("dirichlet", 0.0))
    P.max_steps = 0
    assert lib.deb_solve_heat_mol(C.byref(P)) == 0, lib.deb_last_error()
    assert status.value == deb.DEB_STATUS_COMPLETE and steps.value == 100
    ref = ob.oracle_heat(u0, 0.0, 1.0, 1e-4, E.rk4(0.01), 0.0, 1.0)
    assert np.array_equal(bits(out), bits(ref.u))


def test_launch_counter_counts_kernels():
    lib = deb.load_library()
    y0 = ob.lorenz_ensemble_y0(1000)
    before = lib.deb_launch_count()
    s = deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0).t_eval([1.0]).method(E.dopri5()).solve()
    assert lib.deb_launch_count() - before == 1 and s.gpu_launches == 1
    s = deb.EnsembleIVP.ode(lorenz(), 0.0, 1.0, y0).t_eval([1.0]).method(E.dopri5()).with_stats().solve()
    assert s.gpu_launches == 3  # integration + the two statistics kernels


# ------------------------------------------------------------------------------------------ an output of the real crate
def test_gpu_reproduces_the_output_printed_in_the_crates_documentation():
    """docs/ode.md:108-123 (tests/golden/reference_docs_output.json): the kernels -- DOP853, EvenSolout, event detection with a terminal
    event -- give the 325 evaluations, 20 + 2 steps and six rows the real crate printed, for every trajectory of an ensemble of copies."""
    import json
    import os
    from test_oracle_golden import _docs_example, check_docs_example
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_docs_output.json")))
    ivp = _docs_example(doc)
    ivp.y0s = np.full((64, 1), doc["y0"])
    sol = ivp.solve()
    check_docs_example(sol, doc)
    assert (sol.evals == 325).all() and (sol.accepted == 20).all() and (sol.rejected == 2).all() and (sol.status == deb.DEB_STATUS_INTERRUPTED).all()
