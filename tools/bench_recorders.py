#!/usr/bin/env python3
"""Kernel times of the recorder / event variants on one shape (Lorenz DOPRI5 rtol 1e-8, t in [0,20], 200 k trajectories): the t_eval kernel
(ahead of time), and the run-time compiled recorder kernels -- every step, dense(2), crossing, t_eval + event, plus the same with DOP853.
Prints one JSON line per configuration (kernel_ms as reported by the library for the HOST call: device time of the kernel)."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    y0 = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n))
    te = np.linspace(0.2, 20.0, 100)
    ev = deb.LinearEvent(-25.0, 0.0, [0.0, 0.0, 1.0])  # g = z - 25
    for meth in ("dopri5", "dop853"):
        def base():
            return deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 20.0, y0)
        m = lambda: getattr(E, meth)().rtol(1e-8)
        cases = [("t_eval 100 rows (ahead-of-time kernel)", lambda: base().t_eval(te).method(m())),
                 ("every_step, 1500-row capacity", lambda: base().every_step(1500).method(m())),
                 ("dense(2), 3000-row capacity", lambda: base().dense(2, 3000).method(m())),
                 ("crossing z=25 both", lambda: base().crossing(2, 25.0, deb.CROSSING_BOTH, 128).method(m())),
                 ("t_eval 100 rows + non-terminal linear event z=25", lambda: base().t_eval(te).event(ev, max_event_rows=128).method(m())),
                 ("every_step + terminal event z=25 (terminate after 3)", lambda: base().every_step(400).event(ev, terminate=3, max_event_rows=400).method(m()))]
        # events in one trajectory of eight only (rho = 10: z settles at 9 and never reaches 25): the warp's threshold is never filled
        rho = np.where(np.arange(n) % 8 == 3, 28.0, 10.0)
        cases.append(("t_eval 100 rows + event z=25, events in 1 trajectory of 8 only",
                      lambda: deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, rho, 8.0 / 3.0), 0.0, 20.0, y0).t_eval(te).event(ev, max_event_rows=128).method(m())))
        for label, mk in cases:
            mk().solve()  # compile / warm up
            best = None
            for _ in range(3):
                s = mk().solve()
                best = s.kernel_ms if best is None else min(best, s.kernel_ms)
            acc = int(s.accepted.sum())
            print(json.dumps({"config": f"Lorenz {meth.upper()} rtol 1e-8 t in [0,20], {label}", "n_traj": n, "kernel_ms": round(best, 2), "accepted": acc,
                              "rows": int(np.minimum(s.n_emitted, s.y_eval.shape[1]).sum()), "accepted_steps_per_s": acc / (best * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
