"""Parity at BASELINE.json's full sizes (the other parity tests use sizes the oracle finishes in a second or two).

  C2  10 M Lorenz / DOPRI5 trajectories, t_eval at 100 points: ONE device-resident launch of the real size; every 100th
      trajectory (100 000 of them, spread over the whole index range and over every watermark block class) bitwise against the
      oracle -- rows, final states, counters; ensemble mean / variance per t_eval point over all 10 M
  C2' the same ensemble shape through the HOST path (streamed copies, every visible device), 2 M trajectories, same check
  C4  Euler-Maruyama OU and GBM, 1 M paths x 1000 steps: final states within 1e-12 of the oracle on the host-regenerated stream
  C5  heat equation on 2^24 nodes, RK4, 100 steps: bitwise against the oracle
plus the size-independent identities that hold for all 10 M trajectories (evals = 3 + 6*attempts + accepted, rows complete).
"""
import ctypes as C
import importlib

import numpy as np
import pytest

import oracle_binding as ob
from test_parity_gpu import bits

deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta
pytestmark = pytest.mark.gpu

N_EVAL = 100


def lorenz():
    return deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0)


def check_subset(sub_idx, got, n_eval=N_EVAL):
    """got: dict of numpy arrays for the trajectories sub_idx (global numbers of the bench ensemble)."""
    y0 = deb.perturbed_ensemble([1.0, 1.0, 1.0], sub_idx, seed=2026)
    cpu = ob.oracle_solve(deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0).t_eval(np.arange(1.0, n_eval + 1.0)).method(E.dopri5().rtol(1e-8)))
    for name in ("status", "accepted", "rejected", "evals", "n_emitted"):
        assert np.array_equal(got[name], getattr(cpu, name)), name
    assert np.array_equal(bits(got["y_final"]), bits(cpu.y_final)), "final states differ bitwise"
    assert np.array_equal(bits(got["t_final"]), bits(cpu.t_final))
    assert np.array_equal(bits(got["y_eval"]), bits(cpu.y_eval)), "t_eval rows differ bitwise"


def test_c2_full_10m_launch_subset_bitwise():
    torch = pytest.importorskip("torch")
    lib = deb.load_library()
    n = 10_000_000
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info(0)
    if free < 30e9:
        pytest.skip("needs 30 GB of device memory")
    y0 = torch.from_numpy(deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n), seed=2026)).to(dev)
    out = dict(y_eval=torch.empty((n, N_EVAL, 3), dtype=torch.float64, device=dev), n_emitted=torch.empty(n, dtype=torch.int32, device=dev),
               t_final=torch.empty(n, dtype=torch.float64, device=dev), y_final=torch.empty((n, 3), dtype=torch.float64, device=dev),
               status=torch.empty(n, dtype=torch.int32, device=dev), accepted=torch.empty(n, dtype=torch.int32, device=dev),
               rejected=torch.empty(n, dtype=torch.int32, device=dev), evals=torch.empty(n, dtype=torch.int32, device=dev))
    params = np.array([10.0, 28.0, 8.0 / 3.0])
    t_eval = np.arange(1.0, N_EVAL + 1.0)
    P = deb.OdeProblem()
    P.struct_size = C.sizeof(deb.OdeProblem)
    P.system, P.method, P.dim, P.n_params = deb.DEB_SYS_LORENZ, deb.DEB_DOPRI5, 3, 3
    P.n_traj, P.y0, P.params, P.params_shared = n, y0.data_ptr(), params.ctypes.data, 1
    P.n_eval, P.t_eval, P.t0, P.tf = N_EVAL, t_eval.ctypes.data_as(deb._dp), 0.0, 100.0
    lib.deb_erk_options_default(C.byref(P.opt))
    P.opt.rtol = 1e-8
    P.device, P.memspace, P.stream = 0, deb.DEB_MEM_DEVICE, torch.cuda.current_stream(dev).cuda_stream
    R = deb.Result()
    R.struct_size = C.sizeof(deb.Result)
    for k, v in out.items():
        setattr(R, k, v.data_ptr())
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == 0, lib.deb_last_error()
    torch.cuda.synchronize()
    # identities over ALL trajectories
    acc, rej = out["accepted"].long(), out["rejected"].long()
    assert bool((out["status"] == 0).all()) and bool((out["n_emitted"] == N_EVAL).all())
    assert bool((out["evals"].long() == 3 + 6 * (acc + rej) + acc).all())
    assert int(acc.sum()) == 59955285202 and int(rej.sum()) == 4043829491  # the totals every bench line of this ensemble reports
    assert bool(torch.isfinite(out["y_eval"]).all())
    # ensemble mean / variance per t_eval point over all 10 M trajectories: the library's sums against torch's
    # (different summation order: to rounding)
    sums = torch.zeros((N_EVAL, 3, 2), dtype=torch.float64, device=dev)
    counts = torch.zeros(N_EVAL, dtype=torch.int64, device=dev)
    assert lib.deb_ensemble_stats(out["y_eval"].data_ptr(), out["n_emitted"].data_ptr(), n, N_EVAL, 3, sums.data_ptr(), counts.data_ptr(), 0,
                                  deb.DEB_MEM_DEVICE, torch.cuda.current_stream(dev).cuda_stream) == 0, lib.deb_last_error()
    torch.cuda.synchronize()
    assert bool((counts == n).all())
    s1, s2 = out["y_eval"].sum(dim=0), torch.zeros((N_EVAL, 3), dtype=torch.float64, device=dev)
    for lo in range(0, n, 1_000_000):  # squares in slices: no second 24 GB tensor
        s2 += (out["y_eval"][lo:lo + 1_000_000] ** 2).sum(dim=0)
    mean_ref, var_ref = s1 / n, s2 / n - (s1 / n) ** 2
    mean, var = sums[..., 0] / n, sums[..., 1] / n - (sums[..., 0] / n) ** 2
    assert torch.allclose(mean, mean_ref, rtol=0.0, atol=1e-9) and torch.allclose(var, var_ref, rtol=1e-9, atol=1e-9)
    assert 100.0 < float(var[-1].sum()) < 300.0  # the attractor's spread: the ensemble has decorrelated by t = 100
    # every 100th trajectory -- 100 000 of them, over the whole index range -- bitwise against the oracle (all host cores)
    sub = np.arange(0, n, 100)
    sel = torch.from_numpy(sub).to(dev)
    check_subset(sub, {k: v[sel].cpu().numpy() for k, v in out.items()})


def test_c2_host_path_2m_all_devices_subset_bitwise():
    lib = deb.load_library()
    n = 2_000_000
    devs = list(range(max(1, min(lib.deb_device_count(), 8))))
    y0 = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n), seed=2026)
    ivp = (deb.EnsembleIVP.ode(lorenz(), 0.0, 100.0, y0).t_eval(np.arange(1.0, N_EVAL + 1.0)).method(E.dopri5().rtol(1e-8))
           .devices(devs).with_stats())
    g = ivp.solve()
    assert (g.status == 0).all() and (g.n_emitted == N_EVAL).all()
    assert np.array_equal(g.evals, 3 + 6 * (g.accepted + g.rejected) + g.accepted)
    sub = np.arange(0, n, 997)
    check_subset(sub, {k: getattr(g, k)[sub] for k in ("status", "accepted", "rejected", "evals", "n_emitted", "y_final", "t_final", "y_eval")})
    np.testing.assert_allclose(g.stats_sums[:, :, 0], g.y_eval.sum(axis=0), rtol=1e-11)
    assert (g.stats_counts == n).all()


@pytest.mark.parametrize("which", ["ou", "gbm"])
def test_c4_one_million_paths_against_the_host_stream(which):
    n = 1_000_000
    if which == "ou":
        sde, y0, t0, tf, h = deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), np.full(n, 5.0), 0.0, 10.0, 0.01
    else:
        sde, y0, t0, tf, h = deb.GeometricBrownianMotion(0.1, 0.2), np.full(n, 100.0), 0.0, 1.0, 1e-3
    def prob():
        return deb.EnsembleIVP.sde(sde, t0, tf, y0, seed=2026).t_eval([tf]).method(E.euler(h))
    g, c = prob().solve(), ob.oracle_solve(prob())
    assert np.array_equal(g.status, c.status) and np.array_equal(g.accepted, c.accepted) and np.array_equal(g.evals, c.evals)
    # north_star: SDE paths within 1e-12 on identical Philox streams.  Relative to the scale of the process: an OU path that happens to end
    # within 1e-5 of zero (2 in 10^6 do) carries the same ~1e-16 absolute error as its neighbours, which is not 1e-12 of ITSELF
    scale = float(np.abs(c.y_final).mean())
    np.testing.assert_allclose(g.y_final, c.y_final, rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(g.y_eval, c.y_eval, rtol=1e-12, atol=1e-12 * scale)
    assert g.accepted[0] in (1000, 1001)


def test_c5_heat_2_24_nodes_bitwise():
    n = 1 << 24
    x = np.arange(n, dtype=np.float64)
    u0 = np.sin(np.pi * x / (n - 1))
    m = E.rk4(1.0)
    g = deb.solve_heat_mol(u0, 0.0, float(n - 1), 0.1, m, 0.0, 100.0)
    c = ob.oracle_heat(u0, 0.0, float(n - 1), 0.1, m, 0.0, 100.0)
    assert g.steps == c.steps == 100 and g.status == c.status == "Complete" and g.t == c.t
    assert np.array_equal(bits(g.u), bits(c.u))
