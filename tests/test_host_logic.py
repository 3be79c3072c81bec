"""CPU tests (no GPU) of host-side logic: tableau data, the Philox stream definition, the t_eval row plan, the
splitmix ensemble generator, and the builder mirror."""
import ctypes as C
import importlib
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_binding as ob
import py_restatement as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta


def test_tableau_data_is_what_the_generator_extracts_from_the_reference():
    if not os.path.isdir("/root/reference/src/tableau"):
        pytest.skip("reference tree not present on this box")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tableau.py")], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(ROOT, "differential-equations_b200", "csrc", "erk_tableau_data.h")).read()
    assert out == open(os.path.join(ROOT, "oracle", "erk_tableau_data.h")).read()


def test_two_independent_tableau_extractions_agree():
    """The Python restatement reads its tableaux from tests/golden/reference_tableaux.json (tests/support/reference_tableaux.py:
    tokenizer + recursive-descent evaluator over the Rust sources); the kernels and the C++ oracle are built from
    erk_tableau_data.h (tools/gen_tableau.py: regular expressions + Python eval).  Two readers, no shared code, same bits."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
    import reference_tableaux as rt
    fixture = rt.load_fixture()
    header = pr.load_header_tableaux()
    assert sorted(fixture) == sorted(header) and len(fixture) == 19
    for name, t in header.items():
        for kind, vals in t.items():
            a = np.array(vals, dtype=np.float64)
            b = np.array(fixture[name][kind], dtype=np.float64)
            if kind == "BI" and a.shape != b.shape:  # the crate declares the Verner dense-output array I x I; columns >= order are zero
                assert (b[:, a.shape[1]:] == 0.0).all()
                b = b[:, :a.shape[1]]
            assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64)), (name, kind)
    if os.path.isdir("/root/reference/src/tableau"):  # the committed fixture is what the reader extracts from the crate today
        fresh = rt.read_all()
        for name, t in fresh.items():
            for kind, vals in t.items():
                assert np.array_equal(np.array(vals, dtype=np.float64).view(np.uint64), np.array(fixture[name][kind], dtype=np.float64).view(np.uint64)), (name, kind)


def test_tableau_consistency_conditions():
    """Row sums c_i = sum_j a_ij and sum b_i = 1 (to the precision of the reference's literals), DOPRI5 FSAL row,
    the misplaced DOPRI5 dense row, literal spot checks."""
    T = pr.TAB
    for name, stages in (("DOPRI5", 7), ("DOP853", 12), ("RK4", 4), ("THREE_EIGHTHS", 4), ("MIDPOINT", 2), ("HEUN", 2), ("RALSTON", 2), ("SSP_RK3", 3), ("EULER", 1)):
        t = T[name]
        assert abs(sum(t["B"]) - 1.0) < 5e-15, name
        for i in range(stages):
            assert abs(sum(t["A"][i]) - t["C"][i]) < 2e-15, (name, i)
    d5 = T["DOPRI5"]
    assert d5["A"][6][:6] == d5["B"][:6] and d5["B"][6] == 0.0           # FSAL row
    assert d5["C"][4] == 8.0 / 9.0 and d5["A"][3][1] == -56.0 / 15.0    # quotients evaluated in f64
    assert any(v != 0.0 for v in d5["BI"][0]) and all(v == 0.0 for r in d5["BI"][4:] for v in r)  # cont[4] == 0 quirk
    d8 = T["DOP853"]
    assert d8["C"][1] == 5.260015195876773e-2 and d8["B"][0] == 5.4293734116568765e-2  # the reference's truncated literals
    assert d8["C"][12] == 0.0 and d8["C"][13] == 0.1 and d8["BH"][8] == 7.338466882816118e-1
    assert all(v == 0.0 for r in d8["BI"][:4] for v in r) and d8["BI"][7][15] == -1.4972683625798564e2
    # extra dense stages: c_i = sum_j a_ij too (stage 12 is the derivative at the new point, c = 1 implied)
    for i in (13, 14, 15):
        assert abs(sum(d8["A"][i]) - d8["C"][i]) < 2e-15


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    lib = ob.load_oracle()
    def run(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        lib.orc_philox4x32_10(c, k, o)
        return [int(x) for x in o]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_wiener_increment_definition_and_statistics():
    """dW = sqrt(h) * z with z the documented Box-Muller mapping of the Philox words; regenerate in numpy."""
    lib = ob.load_oracle()
    def philox_np(pair, path, seed):
        M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
        c = [np.uint64(pair & 0xffffffff), np.uint64(pair >> 32), np.uint64(path & 0xffffffff), np.uint64(path >> 32)]
        k0, k1 = seed & 0xffffffff, seed >> 32
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & np.uint64(0xffffffff), (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & np.uint64(0xffffffff)]
            k0, k1 = (k0 + 0x9E3779B9) & 0xffffffff, (k1 + 0xBB67AE85) & 0xffffffff
        return [int(x) for x in c]
    seed, h = 2026, 0.01
    for path, step in ((0, 0), (0, 1), (123456789, 6), (2 ** 33 + 5, 7), (99, 1000)):
        w = philox_np(step >> 1, path, seed)
        a = ((w[0] << 32) | w[1]) >> 11
        b = ((w[2] << 32) | w[3]) >> 11
        u1, u2 = (a + 1) * 2.0 ** -53, b * 2.0 ** -53
        r, th = math.sqrt(-2.0 * math.log(u1)), 6.283185307179586 * u2
        z = r * math.sin(th) if step & 1 else r * math.cos(th)
        got = lib.orc_wiener_increment(seed, path, step, 0, 1, h)
        assert got == math.sqrt(h) * z
    zs = np.array([lib.orc_wiener_increment(7, p, s, 0, 1, 1.0) for p in range(200) for s in range(100)])
    assert abs(zs.mean()) < 0.03 and abs(zs.std() - 1.0) < 0.03 and abs((zs ** 3).mean()) < 0.08


def test_row_plan_equals_what_the_oracle_emits():
    rng = np.random.default_rng(5)
    for t0, tf in ((0.0, 10.0), (10.0, 0.0), (-3.0, 2.0)):
        for _ in range(20):
            pts = rng.choice([t0, tf, 0.5 * (t0 + tf), t0 - 1.0, tf + (tf - t0), t0 + 0.3 * (tf - t0)], size=rng.integers(1, 7)).tolist() + rng.uniform(min(t0, tf) - 2, max(t0, tf) + 2, 3).tolist()
            ivp = deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), t0, tf, [[1.0, 0.0]]).t_eval(pts).method(E.dopri5())
            s = ob.oracle_solve(ivp)
            m = int(s.n_emitted[0])
            rows = deb._plan_rows(np.asarray(pts), t0, tf)
            inside = [r for r in rows if (r <= tf if tf > t0 else r >= tf)]
            assert m == len(inside) and np.isfinite(s.y_eval[0, :m]).all()
            # emitted states agree with the closed form cos/sin to the solver tolerance
            np.testing.assert_allclose(s.y_eval[0, :m, 0], np.cos(np.array(inside) - t0), atol=2e-4)


def test_splitmix_generator_known_values():
    u = ob.splitmix64_uniform(2026, 6)
    assert ((u >= -0.5) & (u < 0.5)).all()
    # reference implementation of splitmix64 in Python ints
    def sm(seed, k):
        z = (seed + k * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        return z ^ (z >> 31)
    assert u.tolist() == [(sm(2026, k) >> 11) * 2.0 ** -53 - 0.5 for k in range(1, 7)]
    y0 = ob.lorenz_ensemble_y0(10)
    assert y0.shape == (10, 3) and np.array_equal(y0[3], 1.0 + ob.splitmix64_uniform(2026, 12)[9:12])


def test_builder_mirror_semantics():
    ivp = deb.EnsembleIVP.ode(deb.VanDerPolOscillator([0.5, 1.5]), 0.0, 1.0, [[2.0, 0.0], [2.0, 0.0]]).method(E.dop853()).rtol(1e-9).atol([1e-9, 1e-10])
    P, R, arrs, ts, keep = ivp.build_problem()
    assert P.params_shared == 0 and P.n_params == 1 and P.dim == 2 and P.n_traj == 2 and P.method == deb.DEB_DOP853
    assert P.opt.rtol == 1e-9 and P.opt.atol_vec[1] == 1e-10 and not P.opt.rtol_vec
    with pytest.raises(ValueError):
        deb.EnsembleIVP.ode(deb.VanDerPolOscillator(1.0), 0.0, 1.0, [[2.0, 0.0]]).solve()  # no method
    with pytest.raises(ValueError):
        deb.EnsembleIVP.ode(deb.VanDerPolOscillator(1.0), 0.0, 1.0, [[2.0, 0.0]]).method(E.dopri5().rtol([1e-3])).build_problem()


def test_heston_vector_sde_oracle_against_a_python_restatement():
    """examples/sde/02_heston_model: 2-D state, diagonal noise, correlated increments, three_eighths(0.01) drift stages
    (fixed/stochastic.rs:67-146).  The oracle against a direct Python restatement that takes its Wiener increments from the
    same stream definition (component c of step s = normal number 2s + c); plus Milstein (milstein.rs:107-180)."""
    import math
    lib = ob.load_oracle()
    mu, kappa, theta, sigma, rho = 0.1, 2.0, 0.04, 0.3, -0.7
    seed, off, h = 42, 1000, 0.01
    drift = lambda y: [mu * y[0], kappa * (theta - y[1])]
    diff = lambda y: [y[0] * math.sqrt(y[1]), sigma * math.sqrt(y[1])]
    T = pr.TAB["THREE_EIGHTHS"]
    def path_py(i, milstein):
        t, y, steps = 0.0, [100.0, 0.04], 0
        dydt = drift(y)
        hh = h
        while True:
            if t + hh - 1.0 > 0.0:
                if abs(1.0 - t) < pr.EPS10:
                    break
                hh = 1.0 - t
            steps += 1
            dw = [lib.orc_wiener_increment(seed, off + i, steps - 1, c, 2, hh) for c in range(2)]
            dw[1] = rho * dw[0] + math.sqrt(1.0 - rho * rho) * dw[1]
            g = diff(y)
            if milstein:
                sq = math.sqrt(hh)
                ga = diff([y[c] + sq * g[c] for c in range(2)])
                factor = 1.0 / (2.0 * sq)
                yn = [((y[c] + 1.0 * (dydt[c] * hh)) + 1.0 * (g[c] * dw[c])) + 1.0 * ((ga[c] - g[c]) * (dw[c] * dw[c] - hh) * factor) for c in range(2)]
            else:
                k = [dydt]
                for s_ in range(1, 4):
                    ys = list(y)
                    for j in range(s_):
                        ah = T["A"][s_][j] * hh
                        ys = [ys[c] + ah * k[j][c] for c in range(2)]
                    k.append(drift(ys))
                inc = [0.0, 0.0]
                for s_ in range(4):
                    w = T["B"][s_] * hh
                    inc = [inc[c] + w * k[s_][c] for c in range(2)]
                yn = [(y[c] + 1.0 * inc[c]) + 1.0 * (g[c] * dw[c]) for c in range(2)]
            t += hh
            y = yn
            dydt = drift(y)
            if abs(1.0 - t) <= pr.EPS10:
                break
        return y, steps
    for meth, mil in ((E.three_eighths(h), False), (deb.Milstein.new(h), True)):
        sol = ob.oracle_solve(deb.EnsembleIVP.sde(deb.HestonModel(mu, kappa, theta, sigma, rho), 0.0, 1.0, np.tile([100.0, 0.04], (6, 1)),
                                                  seed=seed, path_offset=off).t_eval([0.5, 1.0]).method(meth))
        assert sol.y_eval.shape == (6, 2, 2) and (sol.status == 0).all()
        for i in range(6):
            y, steps = path_py(i, mil)
            assert steps == sol.accepted[i] and np.array_equal(np.array(y).view(np.uint64), sol.y_final[i].view(np.uint64))
            assert np.array_equal(sol.y_eval[i, 1].view(np.uint64), sol.y_final[i].view(np.uint64))


def test_fixed_step_schedule_planner_follows_the_reference_loop():
    """deb_plan_fixed_steps (host only): the schedule the fixed-step / SDE kernels follow = the reference's loop
    (solve_ivp.rs:193-209, :263), including the cases where the clip at tf fires twice (t + (tf - t) misses tf by an ulp when
    t and tf have opposite signs: one more tiny step).  Random sweep of the class the round-1 review found (t0 << 0 < tf, coarse
    h) plus ordinary cases: replaying (n_steps, h0, tail) must visit exactly the reference loop's times."""
    import ctypes as C
    import importlib
    deb = importlib.import_module("differential-equations_b200")
    lib = deb.load_library()
    eps10 = 10.0 * np.finfo(float).eps

    def reference_loop(t0, tf, h, max_steps):
        d = math.copysign(1.0, tf - t0)
        if tf == t0 or math.copysign(1.0, h) != d or abs(h) > abs(tf - t0) or h == 0.0:  # validate_step_size_parameters, utils.rs:60-157
            return [], t0, deb.DEB_STATUS_BAD_INPUT
        t, hs = t0, []
        while True:
            if (t + h - tf) * d > 0.0:
                h_new = tf - t
                if abs(h_new) < eps10:
                    return hs, t, deb.DEB_STATUS_COMPLETE
                h = h_new
            if len(hs) >= max_steps:
                return hs, t, deb.DEB_STATUS_MAX_STEPS
            hs.append(h)
            t = t + h
            if abs(tf - t) <= eps10:
                return hs, t, deb.DEB_STATUS_COMPLETE

    rng = np.random.default_rng(11)
    cases = [(-22246.572886668695, 976.800691167858, 7199.779502015325)]  # the reviewer's example: 5 steps, the last two clipped
    for _ in range(3000):
        t0 = -10.0 ** rng.uniform(0, 5)
        tf = 10.0 ** rng.uniform(0, 4)
        cases.append((t0, tf, (tf - t0) / rng.uniform(1.5, 12.0)))
    for _ in range(1000):
        t0, span = rng.uniform(-5, 5), 10.0 ** rng.uniform(-2, 2)
        sgn = 1.0 if rng.uniform() < 0.5 else -1.0
        cases.append((t0, t0 + sgn * span, sgn * span / rng.integers(1, 400)))
    double_clips = 0
    for t0, tf, h in cases:
        n_steps, n_tail, status = C.c_int64(0), C.c_int32(0), C.c_int32(-1)
        tail = (C.c_double * 4)()
        rc = lib.deb_plan_fixed_steps(t0, tf, h, 0.0, float("inf"), 10000, C.byref(n_steps), C.byref(n_tail), tail, C.byref(status))
        hs, t_end, st = reference_loop(t0, tf, h, 10000)
        assert rc == 0 and status.value == st and n_steps.value == len(hs), (t0, tf, h)
        k = n_steps.value - n_tail.value
        replay = [h] * k + [tail[q] for q in range(n_tail.value)]
        assert replay == hs, (t0, tf, h, replay[-3:], hs[-3:])
        double_clips += n_tail.value >= 2
    assert double_clips > 100  # the sweep does reach the double clip
    # BadInput (utils.rs:60-157) and MaxSteps
    for t0, tf, h in ((0.0, 0.0, 0.1), (0.0, 1.0, -0.1), (0.0, 1.0, 2.0)):
        n_steps, n_tail, status = C.c_int64(0), C.c_int32(0), C.c_int32(-1)
        tail = (C.c_double * 4)()
        assert lib.deb_plan_fixed_steps(t0, tf, h, 0.0, float("inf"), 10000, C.byref(n_steps), C.byref(n_tail), tail, C.byref(status)) == 0
        assert status.value == deb.DEB_STATUS_BAD_INPUT and n_steps.value == 0
    n_steps, n_tail, status = C.c_int64(0), C.c_int32(0), C.c_int32(-1)
    tail = (C.c_double * 4)()
    assert lib.deb_plan_fixed_steps(0.0, 1.0, 1e-3, 0.0, float("inf"), 100, C.byref(n_steps), C.byref(n_tail), tail, C.byref(status)) == 0
    assert status.value == deb.DEB_STATUS_MAX_STEPS and n_steps.value == 100


def test_device_list_split_is_a_partition():
    """deb_shard_layout (host only): the block-cyclic split behind deb_ode_problem.devices -- every block of 2^shift trajectories on
    exactly one device, block b on device b mod G, the last block partial; counts add up for every ensemble size around the block
    boundaries."""
    import ctypes as C
    import importlib
    deb = importlib.import_module("differential-equations_b200")
    lib = deb.load_library()
    for shift in (2, 5, 12):
        B = 1 << shift
        for G in (1, 2, 3, 8, 16):
            for n in sorted({0, 1, B - 1, B, B + 1, G * B - 1, G * B, G * B + 1, 7 * B + 3, 10_000_000 if shift == 12 else 1000}):
                nb = (n + B - 1) // B
                seen, total = set(), 0
                for g in range(G):
                    nlb, cnt, gb = C.c_int64(-1), C.c_int64(-1), C.c_int64(-2)
                    assert lib.deb_shard_layout(n, G, g, shift, 0, C.byref(nlb), C.byref(cnt), C.byref(gb)) == 0
                    mine = [b for b in range(nb) if b % G == g]
                    assert nlb.value == len(mine) and cnt.value == sum(min(B, n - b * B) for b in mine), (shift, G, n, g)
                    assert gb.value == (mine[0] if mine else -1)
                    probe = [0, len(mine) - 1] if len(mine) > 1 else list(range(len(mine)))
                    for lb in probe + ([len(mine) // 2] if len(mine) > 2 else []):
                        assert lib.deb_shard_layout(n, G, g, shift, lb, C.byref(nlb), C.byref(cnt), C.byref(gb)) == 0
                        assert gb.value == mine[lb]
                    assert lib.deb_shard_layout(n, G, g, shift, len(mine), C.byref(nlb), C.byref(cnt), C.byref(gb)) == 0 and gb.value == -1
                    seen.update(mine)
                    total += cnt.value
                assert total == n and seen == set(range(nb))
    nlb, cnt, gb = C.c_int64(), C.c_int64(), C.c_int64()
    assert lib.deb_shard_layout(100, 2, 2, 12, 0, C.byref(nlb), C.byref(cnt), C.byref(gb)) == deb.DEB_ERR_BAD_ARG
