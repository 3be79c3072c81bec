#!/usr/bin/env python3
"""Timings of the other BASELINE.json configurations (C3 Van der Pol/DOP853 sweep, C4 Euler-Maruyama, C5 heat MoL) on
one B200, device-resident, with the roofline each kernel is bounded by.  Not the headline bench (that is bench.py /
config C2); the numbers go to profiles/ and DESIGN.md.

    python tools/bench_configs.py [--c3 N] [--c4 N] [--c5 LOG2N] [--reps R]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
deb = importlib.import_module("differential-equations_b200")
lib = dev = stream = None
DEVICE = 0


def init(device=0):
    """Bind the helpers of this module to one GPU (bench.py calls this with its local rank)."""
    global lib, dev, stream, DEVICE
    DEVICE = device
    lib = deb.load_library()
    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream(dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def result_buffers(n, n_eval, dim):
    R = deb.Result()
    R.struct_size = C.sizeof(deb.Result)
    bufs = dict(y_eval=torch.empty((n, max(n_eval, 1), dim), dtype=torch.float64, device=dev), n_emitted=torch.empty(n, dtype=torch.int32, device=dev),
                t_final=torch.empty(n, dtype=torch.float64, device=dev), y_final=torch.empty((n, dim), dtype=torch.float64, device=dev),
                status=torch.empty(n, dtype=torch.int32, device=dev), accepted=torch.empty(n, dtype=torch.int32, device=dev),
                rejected=torch.empty(n, dtype=torch.int32, device=dev), evals=torch.empty(n, dtype=torch.int32, device=dev))
    for k, v in bufs.items():
        setattr(R, k, v.data_ptr())
    return R, bufs


def fp64_peak():
    v = C.c_double(0)
    lib.deb_fp64_issue_peak(DEVICE, 0, C.byref(v), None)
    return v.value


def c3(n, reps):
    """Van der Pol mu in [0.1, 50] sweep, DOP853 rtol=atol=1e-8, y0=(2,0), t in [0,100], final state + counters."""
    mu = torch.from_numpy(0.1 + 49.9 * np.arange(n) / max(n - 1, 1)).to(dev)
    y0 = torch.from_numpy(np.tile([2.0, 0.0], (n, 1))).to(dev)
    te = np.array([100.0])
    P = deb.OdeProblem()
    P.struct_size = C.sizeof(deb.OdeProblem)
    P.system, P.method, P.dim, P.n_params = deb.DEB_SYS_VAN_DER_POL, deb.DEB_DOP853, 2, 1
    P.n_traj, P.y0, P.params, P.params_shared = n, y0.data_ptr(), mu.data_ptr(), 0
    P.n_eval, P.t_eval, P.t0, P.tf = 1, te.ctypes.data_as(deb._dp), 0.0, 100.0
    lib.deb_erk_options_default(C.byref(P.opt))
    P.opt.rtol = P.opt.atol = 1e-8
    P.device, P.memspace, P.stream = DEVICE, deb.DEB_MEM_DEVICE, stream.cuda_stream
    R, bufs = result_buffers(n, 1, 2)
    def run():
        assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == 0, lib.deb_last_error()
    best, avg = timed(run, reps)
    acc = int(bufs["accepted"].sum(dtype=torch.int64)); rej = int(bufs["rejected"].sum(dtype=torch.int64))
    ok = int((bufs["status"] == 0).sum())
    # DP operations the kernel has to execute: 474 per step attempt + 7 per accepted step; the reference also
    # evaluates the three dense-output stages and cont[4..7] on EVERY accepted step (+367), the kernel only for a step that
    # emits a row -- the fraction below counts the work that is needed, not the work that was avoided
    ops = 474 * (acc + rej) + 7 * acc
    pk = fp64_peak()
    return {"config": "C3 Van der Pol mu-sweep DOP853", "n_traj": n, "ms": best, "ms_avg": avg, "accepted": acc, "rejected": rej, "complete": ok,
            "accepted_steps_per_s": acc / (best * 1e-3), "algorithmic_ops": ops,
            "roofline": {"bound": "fp64", "achieved_TFLOPs": ops / (best * 1e-3) / 1e12, "peak": pk / 1e12, "frac": ops / (best * 1e-3) / pk,
                         "frac_reference_count": (474 * (acc + rej) + 374 * acc) / (best * 1e-3) / pk,
                         "note": "474 DP ops per attempt + 7 per accepted step; frac_reference_count uses 474 + 374 (dense stages on every accepted step, as the reference evaluates them)"}}


def c4(n, which, reps):
    dim = 1
    if which == "ou":
        sysid, params, t0, tf, h, y0v, nsteps = deb.DEB_SDE_OU, np.array([0.5, 1.0, 0.3]), 0.0, 10.0, 0.01, 5.0, 1001
    elif which == "heston":  # examples/sde/02_heston_model (2-D state), Euler-Maruyama h = 1e-3
        sysid, params, t0, tf, h, y0v, nsteps, dim = deb.DEB_SDE_HESTON, np.array([0.1, 2.0, 0.04, 0.3, -0.7]), 0.0, 1.0, 1e-3, [100.0, 0.04], 1000, 2
    else:
        sysid, params, t0, tf, h, y0v, nsteps = deb.DEB_SDE_GBM, np.array([0.1, 0.2]), 0.0, 1.0, 1e-3, 100.0, 1000
    y0 = np.array(y0v, dtype=np.float64).reshape(-1)
    te = np.array([tf])
    P = deb.SdeProblem()
    P.struct_size = C.sizeof(deb.SdeProblem)
    P.system, P.method, P.dim, P.n_params = sysid, deb.DEB_EULER, dim, params.size
    P.n_traj, P.y0, P.y0_shared, P.params, P.params_shared = n, y0.ctypes.data, 1, params.ctypes.data, 1
    P.n_eval, P.t_eval, P.t0, P.tf = 1, te.ctypes.data_as(deb._dp), t0, tf
    lib.deb_erk_options_default(C.byref(P.opt))
    P.opt.h0 = h
    P.seed, P.path_offset = 2026, 0
    P.device, P.memspace, P.stream = DEVICE, deb.DEB_MEM_DEVICE, stream.cuda_stream
    R, bufs = result_buffers(n, 1, dim)
    def run():
        assert lib.deb_solve_sde(C.byref(P), C.byref(R)) == 0, lib.deb_last_error()
    best, avg = timed(run, reps)
    steps = int(bufs["accepted"].sum(dtype=torch.int64))
    yf = bufs["y_final"].reshape(-1, dim)[:, 0]
    yf = yf[torch.isfinite(yf)]
    mean, var = float(yf.mean()), float(yf.var())
    if which == "heston":
        exp_mean, exp_var = 100.0 * np.exp(0.1), float("nan")
    elif which == "ou":  # exact moments of the EM chain are close to the SDE's: mean 1 + 4 e^{-5}, var sigma^2/(2 theta)(1 - e^{-10})
        exp_mean, exp_var = 1.0 + 4.0 * np.exp(-5.0), 0.09 / 1.0 * (1 - np.exp(-10.0))
    else:
        exp_mean, exp_var = 100.0 * np.exp(0.1), 100.0 ** 2 * np.exp(0.2) * (np.exp(0.04) - 1)
    return {"config": f"C4 Euler-Maruyama {which.upper()}", "n_paths": n, "steps_per_path": nsteps, "ms": best, "ms_avg": avg,
            "path_steps_per_s": steps / (best * 1e-3), "mean": mean, "var": var, "sde_mean": exp_mean, "sde_var": exp_var}


def c5(log2n, reps):
    n = 1 << log2n
    x = np.arange(n, dtype=np.float64)
    u0 = torch.from_numpy(np.sin(np.pi * x / (n - 1))).to(dev)
    out = torch.empty_like(u0)
    P = deb.HeatProblem()
    P.struct_size = C.sizeof(deb.HeatProblem)
    P.n_nodes, P.lo, P.hi, P.alpha = n, 0.0, float(n - 1), 0.1
    P.bc_lower_kind = P.bc_upper_kind = 0
    P.method, P.h, P.t0, P.tf, P.max_steps = deb.DEB_RK4, 1.0, 0.0, 100.0, 10000
    P.u0, P.u_final = u0.data_ptr(), out.data_ptr()
    steps, status, tfin = C.c_int64(0), C.c_int32(-1), C.c_double(0)
    P.steps, P.status, P.t_final = C.pointer(steps), C.pointer(status), C.pointer(tfin)
    P.device, P.memspace, P.stream = DEVICE, deb.DEB_MEM_DEVICE, stream.cuda_stream
    def run():
        assert lib.deb_solve_heat_mol(C.byref(P)) == 0, lib.deb_last_error()
    best, avg = timed(run, reps)
    bytes_alg = 16.0 * n * steps.value + 4 * 8.0 * n  # whole-step kernel: read y, write y' per step (+ copy in/out)
    ops = 52.0 * n * steps.value  # DP operations per node and RK4 step (dx a power of two): 4 x 9 stencil + 3 x 2 stage + 2 x 4 + 2 solution
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    return {"config": "C5 heat MoL RK4", "n_nodes": n, "steps": steps.value, "status": status.value, "ms": best, "ms_avg": avg, "ms_per_step": best / max(steps.value, 1),
            "node_steps_per_s": n * steps.value / (best * 1e-3),
            "roofline": {"bound": "hbm", "achieved_GBs": bytes_alg / (best * 1e-3) / 1e9, "peak_GBs": hbm, "frac": bytes_alg / (best * 1e-3) / 1e9 / hbm,
                         "fp64_achieved_Tops": ops / (best * 1e-3) / 1e12, "fp64_peak_Tops": fp64_peak() / 1e12,
                         "sweep_equivalent_GBs": 128.0 * n * steps.value / (best * 1e-3) / 1e9,
                         "note": "whole deb_solve_heat_mol call incl. buffer allocation, D2D copy in/out and one launch per step; "
                                 "sweep_equivalent = what a stage-by-stage sweep (16 arrays per RK4 step) would have to move in the same time"}}


def widened(n, reps):
    """The widened rows: RKF45 / Cash-Karp, the Verner pairs, EvenSolout, a user-defined (NVRTC) right-hand side, on the C2 shape."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    y0h = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n))
    y0 = torch.from_numpy(y0h).to(dev)
    prm = np.array([10.0, 28.0, 8.0 / 3.0])
    te = np.arange(1.0, 101.0)
    user = deb.ode_from_source(3, "const double x = y[0], yv = y[1], z = y[2]; dydt[0] = p[0]*(yv-x); dydt[1] = x*(p[1]-z)-yv; dydt[2] = x*yv-p[2]*z;", prm)
    out = []
    for label, system, method, solout in (("DOPRI5 t_eval (C2 shape)", deb.DEB_SYS_LORENZ, deb.DEB_DOPRI5, 0), ("DOPRI5 even(1.0)", deb.DEB_SYS_LORENZ, deb.DEB_DOPRI5, 1),
                                          ("RKF45 t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKF45, 0), ("Cash-Karp t_eval", deb.DEB_SYS_LORENZ, deb.DEB_CASH_KARP, 0),
                                          ("DOP853 t_eval", deb.DEB_SYS_LORENZ, deb.DEB_DOP853, 0), ("user-defined Lorenz (NVRTC) DOPRI5 t_eval", user.system_id, deb.DEB_DOPRI5, 0),
                                          ("RKV655e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV655E, 0), ("RKV656e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV656E, 0),
                                          ("RKV766e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV766E, 0), ("RKV767e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV767E, 0),
                                          ("RKV877e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV877E, 0), ("RKV878e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV878E, 0),
                                          ("RKV988e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV988E, 0), ("RKV989e t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RKV989E, 0),
                                          ("RKV655e even(1.0)", deb.DEB_SYS_LORENZ, deb.DEB_RKV655E, 1),
                                          ("RK4 h=0.01 (10^4 fixed steps) t_eval", deb.DEB_SYS_LORENZ, deb.DEB_RK4, 0),
                                          ("Euler h=0.01 (10^4 fixed steps) t_eval", deb.DEB_SYS_LORENZ, deb.DEB_EULER, 0)):
        P = deb.OdeProblem()
        P.struct_size = C.sizeof(deb.OdeProblem)
        P.system, P.method, P.dim, P.n_params = system, method, 3, 3
        P.n_traj, P.y0, P.params, P.params_shared = n, y0.data_ptr(), prm.ctypes.data, 1
        P.n_eval, P.t_eval, P.t0, P.tf = (102 if solout else 100), te.ctypes.data_as(deb._dp), 0.0, 100.0
        lib.deb_erk_options_default(C.byref(P.opt))
        P.opt.rtol = 1e-8
        if method < deb.DEB_DOPRI5:  # fixed-step constructors take h
            P.opt.h0 = 0.01
            P.opt.max_steps = 20000
        P.solout, P.even_dt = solout, 1.0
        P.device, P.memspace, P.stream = DEVICE, deb.DEB_MEM_DEVICE, stream.cuda_stream
        R, bufs = result_buffers(n, 102, 3)
        def run():
            assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == 0, lib.deb_last_error()
        best, avg = timed(run, reps)
        acc = int(bufs["accepted"].sum(dtype=torch.int64)); rej = int(bufs["rejected"].sum(dtype=torch.int64))
        out.append({"config": label, "n_traj": n, "ms": best, "accepted": acc, "rejected": rej, "complete": int((bufs["status"] == 0).sum()),
                    "accepted_steps_per_s": acc / (best * 1e-3)})
    # Milstein vs Euler-Maruyama, GBM
    for label, meth in (("GBM Euler-Maruyama", deb.DEB_EULER), ("GBM Milstein", deb.DEB_MILSTEIN)):
        params, y0v = np.array([0.1, 0.2]), np.array([100.0])
        tev = np.array([1.0])
        P = deb.SdeProblem()
        P.struct_size = C.sizeof(deb.SdeProblem)
        P.system, P.method, P.dim, P.n_params = deb.DEB_SDE_GBM, meth, 1, 2
        P.n_traj, P.y0, P.y0_shared, P.params, P.params_shared = 10 * n, y0v.ctypes.data, 1, params.ctypes.data, 1
        P.n_eval, P.t_eval, P.t0, P.tf = 1, tev.ctypes.data_as(deb._dp), 0.0, 1.0
        lib.deb_erk_options_default(C.byref(P.opt))
        P.opt.h0 = 1e-3
        P.seed = 2026
        P.device, P.memspace, P.stream = DEVICE, deb.DEB_MEM_DEVICE, stream.cuda_stream
        R, bufs = result_buffers(10 * n, 1, 1)
        def run():
            assert lib.deb_solve_sde(C.byref(P), C.byref(R)) == 0, lib.deb_last_error()
        best, avg = timed(run, reps)
        out.append({"config": label, "n_paths": 10 * n, "ms": best, "path_steps_per_s": int(bufs["accepted"].sum(dtype=torch.int64)) / (best * 1e-3)})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3", type=int, default=4_000_000)
    ap.add_argument("--c4", type=int, default=100_000_000)
    ap.add_argument("--c5", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--widened", type=int, default=0, help="also time the widened rows on this many trajectories")
    a = ap.parse_args()
    init(0)
    if a.c3:
        print(json.dumps(c3(a.c3, a.reps)), flush=True)
    if a.c4:
        for w in ("ou", "gbm", "heston"):
            print(json.dumps(c4(a.c4, w, a.reps)), flush=True)
    if a.c5:
        print(json.dumps(c5(a.c5, a.reps)), flush=True)
    if a.widened:
        for r in widened(a.widened, a.reps):
            print(json.dumps(r), flush=True)
