"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

deb = importlib.import_module("differential-equations_b200")

_oracle = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def load_oracle():
    global _oracle
    if _oracle is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.cpp", "philox_ref.h", "erk_tableau_data.h")]
        if not os.path.exists(ORACLE_LIB) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in srcs):
            build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        lib.orc_solve_ode.argtypes = [C.POINTER(deb.OdeProblem), C.POINTER(deb.Result), C.c_int]
        lib.orc_solve_sde.argtypes = [C.POINTER(deb.SdeProblem), C.POINTER(deb.Result), C.c_int]
        lib.orc_solve_heat_mol.argtypes = [C.POINTER(deb.HeatProblem), C.c_int]
        lib.orc_wiener_increment.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_double]
        lib.orc_wiener_increment.restype = C.c_double
        lib.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.orc_philox4x32_10.restype = None
        lib.orc_hardware_threads.restype = C.c_int
        lib.orc_pow_array.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_void_p]
        lib.orc_pow_array.restype = None
        lib.orc_heat_rhs.argtypes = [C.POINTER(deb.HeatProblem), C.c_void_p, C.c_void_p]
        _oracle = lib
    return _oracle


def oracle_solve(ivp, n_threads=0):
    """Run the problem described by an EnsembleIVP builder through the CPU oracle; same result container."""
    lib = load_oracle()
    # the oracle is one trajectory after the other on the host: no device list, trajectory-major rows, no fused statistics
    saved = (ivp._devices, ivp._layout, ivp._stats)
    ivp._devices, ivp._layout, ivp._stats = [], deb.DEB_LAYOUT_TRAJ_MAJOR, False
    try:
        P, res, arrs, t_sorted, keep = ivp.build_problem()
    finally:
        ivp._devices, ivp._layout, ivp._stats = saved
    fn = lib.orc_solve_ode if ivp.kind == "ode" else lib.orc_solve_sde
    rc = fn(C.byref(P), C.byref(res), int(n_threads))
    if rc != 0:
        raise ValueError(f"oracle rejected the problem (rc={rc})")
    rows = deb._plan_rows(ivp._t_eval, ivp.t0, ivp.tf) if ivp._even_dt <= 0.0 else deb._even_rows(ivp.t0, ivp.tf, ivp._even_dt)
    return ivp.wrap_result(arrs, rows, res)


def oracle_heat(u0, lo, hi, alpha, method, t0, tf, bc_lower=("dirichlet", 0.0), bc_upper=("dirichlet", 0.0), n_threads=0):
    lib = load_oracle()
    P, out, (tfin, steps, status), keep = deb.build_heat_problem(u0, lo, hi, alpha, method, t0, tf, bc_lower, bc_upper)
    rc = lib.orc_solve_heat_mol(C.byref(P), int(n_threads))
    if rc != 0:
        raise ValueError(f"oracle rejected the problem (rc={rc})")
    return deb.HeatSolution(out, tfin.value, steps.value, deb._STATUS_NAME.get(status.value, "?"))


def splitmix64_uniform(seed: int, count: int) -> np.ndarray:
    """u_k for k = 1..count (the package's counter-based generator)."""
    return deb.splitmix64_uniform(seed, np.arange(1, count + 1, dtype=np.uint64))


def lorenz_ensemble_y0(n: int, seed: int = 2026) -> np.ndarray:
    """y0_i = (1,1,1) + (u_{3i}, u_{3i+1}, u_{3i+2})  (config C1/C2)."""
    return deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n), seed)


def oracle_heat_rhs(u, lo, hi, alpha, bc_lower=("dirichlet", 0.0), bc_upper=("dirichlet", 0.0)):
    lib = load_oracle()
    u = np.ascontiguousarray(u, dtype=np.float64)
    P, _o, _s, keep = deb.build_heat_problem(u, lo, hi, alpha, deb.ExplicitRungeKutta.euler(0.0), 0.0, 1.0, bc_lower, bc_upper)
    du = np.empty_like(u)
    assert lib.orc_heat_rhs(C.byref(P), u.ctypes.data, du.ctypes.data) == 0
    return du


def libm_pow(x: np.ndarray, y: float) -> np.ndarray:
    """glibc pow() elementwise (NOT numpy's own vectorised pow): what f64::powf evaluates to on this box."""
    lib = load_oracle()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.orc_pow_array(x.ctypes.data, float(y), x.size, out.ctypes.data)
    return out
