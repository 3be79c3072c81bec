set -x
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_t6.log 2>&1; tail -5 gpurun_out/r2_t6.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitizer_smoke.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
DEB_WM_SHIFT=4 timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -c "
import importlib, numpy as np, sys
sys.path.insert(0, '.')
deb = importlib.import_module('differential-equations_b200'); E = deb.ExplicitRungeKutta
y3 = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(300)); lor = deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0)
for m in (E.dopri5(), E.dop853(), E.rkv878e(), E.rk4(0.01)):
    deb.EnsembleIVP.ode(lor, 0.0, 0.5, y3).t_eval(np.linspace(0.0, 0.5, 7)).method(m).with_stats().solve()
deb.EnsembleIVP.sde(deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), 0.0, 0.2, np.full(300, 2.0), seed=3).t_eval(np.linspace(0.0, 0.2, 6)).method(E.euler(0.01)).solve()
print('racecheck smoke done')
" > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck.txt
