#!/usr/bin/env python3
"""Small end-to-end invocations of every kernel path that changed in round 2, sized for compute-sanitizer (memcheck / racecheck run the
kernels 10-100x slower): row staging (dims 1-3, aligned and unaligned row counts), the streamed HOST pipeline with several chunks and
tiny watermark blocks, parked emission of the methods with extra dense stages, the fixed-step and SDE kernels with rows, the row-major
transpose, fused statistics, a user-defined SDE and a generated sensitivity system.
    compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DEB_WM_SHIFT", "4")
os.environ.setdefault("DEB_HOST_CHUNK", "200")
deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta


def main():
    n = 700
    y3 = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n))
    lor = deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0)
    done = 0
    for meth in (E.dopri5().rtol(1e-7), E.dop853().rtol(1e-7), E.rkv989e().rtol(1e-7), E.rkv655e().rtol(1e-7), E.rkf45().rtol(1e-7), E.rk4(0.01), E.dopri5().max_steps(30)):
        for te in (np.linspace(0.0, 1.0, 9), np.linspace(0.1, 1.0, 7), [1.0]):
            s = deb.EnsembleIVP.ode(lor, 0.0, 1.0, y3).t_eval(te).method(meth).with_stats().solve()
            assert s.n_emitted.max() <= len(te)
            done += 1
    s = deb.EnsembleIVP.ode(lor, 0.0, 1.0, y3).t_eval(np.linspace(0.0, 1.0, 5)).method(E.dop853()).layout(deb.DEB_LAYOUT_ROW_MAJOR).solve()
    s = deb.EnsembleIVP.ode(deb.HarmonicOscillator(2.0), 0.0, 2.0, np.tile([1.0, 0.0], (n, 1))).even(0.25).method(E.rkv767e()).solve()
    s = deb.EnsembleIVP.ode(deb.ExponentialGrowth(-0.5), 0.0, 2.0, np.ones((n, 1))).t_eval(np.linspace(0.0, 2.0, 6)).method(E.dop853()).solve()
    done += 3
    for sde, y0 in ((deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), np.full(n, 2.0)), (deb.HestonModel(0.1, 2.0, 0.04, 0.3, -0.7), np.tile([100.0, 0.04], (n, 1)))):
        for m in (E.euler(0.01), deb.Milstein.new(0.01)):
            deb.EnsembleIVP.sde(sde, 0.0, 0.5, y0, seed=3).t_eval(np.linspace(0.0, 0.5, 6)).method(m).solve()
            done += 1
    ou = deb.sde_from_source(1, "dydt[0] = p[0] * (p[1] - y[0]);", "g[0] = p[2];", params=[0.5, 1.0, 0.3])
    deb.EnsembleIVP.sde(ou, 0.0, 0.5, np.full(n, 2.0), seed=3).t_eval([0.25, 0.5]).method(E.euler(0.01)).solve()
    sens = deb.ode_sensitivity_from_source(1, "dydt[0] = p[0] * y[0] * (1.0 - y[0] / p[1]);", "J[0] = p[0] * (1.0 - 2.0 * y[0] / p[1]);",
                                           "Jp[0] = y[0] * (1.0 - y[0] / p[1]); Jp[1] = p[0] * y[0] * y[0] / (p[1] * p[1]);", [1.0, 10.0])
    z0 = np.zeros((n, 3)); z0[:, 0] = np.linspace(0.5, 2.0, n)
    deb.EnsembleIVP.ode(sens, 0.0, 2.0, z0).t_eval([1.0, 2.0]).method(E.dop853()).solve()
    deb.EnsembleIVP.ode(lor, 0.0, 1.0, y3).every_step(400).event(deb.LinearEvent(-25.0, 0.0, [0.0, 0.0, 1.0]), max_event_rows=400).method(E.dopri5()).solve()
    done += 3
    print("sanitizer smoke: %d solves done" % done)


if __name__ == "__main__":
    main()
