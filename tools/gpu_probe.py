"""Quick GPU probe: FP64 issue peak and the C2 kernel at a few ensemble sizes (not a benchmark)."""
import ctypes as C, importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_binding as ob
deb = ob.deb
lib = deb.load_library()
for fma in (0, 1):
    v, ms = C.c_double(0), C.c_float(0)
    assert lib.deb_fp64_issue_peak(0, fma, C.byref(v), C.byref(ms)) == 0
    print(f"fp64 issue peak ({'DFMA' if fma else 'DADD/DMUL'}): {v.value/1e12:.3f} T DP inst/s  ({ms.value:.2f} ms)", flush=True)
E = deb.ExplicitRungeKutta
for n in (100_000, 1_000_000):
    y0 = ob.lorenz_ensemble_y0(n)
    for rep in range(2):
        g = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 100.0, y0).t_eval(np.arange(1.0, 101.0)).method(E.dopri5().rtol(1e-8)).solve()
        acc = int(g.accepted.sum()); att = acc + int(g.rejected.sum())
        print(f"n={n} kernel {g.kernel_ms:.1f} ms total {g.total_ms:.1f} ms  acc={acc} att={att}  {acc/g.kernel_ms/1e6:.2f} G acc-steps/s", flush=True)
