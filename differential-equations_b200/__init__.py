"""differential-equations_b200 -- host-side mirror of the reference's operator API over the B200 C ABI.

The reference (Ryan-D-Gast/differential-equations v0.6.1) is a Rust crate; this image has no Rust toolchain, so
the host layer above the C ABI (include/deb_ensemble.h, libdeb200.so) is written here in Python with the same
names, argument meaning and error behaviour as the crate's builder API, so the parity tests read like the
reference's own tests:

    reference (Rust)                                            this module
    ---------------------------------------------------------   -----------------------------------------------
    ExplicitRungeKutta::dopri5().rtol(1e-8)                      ExplicitRungeKutta.dopri5().rtol(1e-8)
      src/methods/erk/dormandprince/mod.rs:45-58, erk/mod.rs:164-228
    ExplicitRungeKutta::rk4(h)  / euler(h) ...                   ExplicitRungeKutta.rk4(h) / .euler(h) ...
      src/methods/erk/fixed/mod.rs:41-89
    IVP::ode(&sys, t0, tf, y0).t_eval(pts).method(m).solve()     EnsembleIVP.ode(sys, t0, tf, y0s).t_eval(pts).method(m).solve()
      src/ivp.rs:279,656,632,781
    IVP::sde(&mut sde, t0, tf, y0)...                            EnsembleIVP.sde(sde, t0, tf, y0s, seed=...)...
      src/ivp.rs:504,857
    Solution{t, y, status, evals, steps}                         EnsembleSolution[i] -> Solution  (or raises the Error)
      src/solution.rs:30-53, src/error.rs:13-41

The one difference is the ensemble axis: `y0s` is an (N, dim) array of initial states and a system carries either
one parameter set or an (N, n_params) array (a parameter sweep); `solve()` returns all N results at once.

PyTorch is not needed by this module (ctypes + numpy); bench.py uses torch only for device buffers and streams.
The CUDA extension is mandatory: importing works without it (so CPU-only tooling can read the constants), but
any solve raises if libdeb200.so is missing or no GPU is present -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeb200.so")

# ---------------------------------------------------------------------------------------------- enums (deb_ensemble.h)
DEB_EULER, DEB_MIDPOINT, DEB_HEUN, DEB_RALSTON, DEB_SSP_RK3, DEB_RK4, DEB_THREE_EIGHTHS = range(7)
DEB_DOPRI5, DEB_DOP853, DEB_RKF45, DEB_CASH_KARP = 16, 17, 18, 19
(DEB_RKV655E, DEB_RKV656E, DEB_RKV766E, DEB_RKV767E, DEB_RKV877E, DEB_RKV878E, DEB_RKV988E, DEB_RKV989E) = range(20, 28)
DEB_MILSTEIN = 32
(DEB_SYS_EXPONENTIAL, DEB_SYS_LINEAR, DEB_SYS_HARMONIC, DEB_SYS_LOGISTIC, DEB_SYS_VAN_DER_POL, DEB_SYS_LORENZ,
 DEB_SYS_BRUSSELATOR, DEB_SYS_ROBERTSON) = range(8)
DEB_SDE_OU, DEB_SDE_GBM, DEB_SDE_HESTON = 0, 1, 2
DEB_STATUS_COMPLETE, DEB_STATUS_MAX_STEPS, DEB_STATUS_STEP_SIZE, DEB_STATUS_STIFFNESS, DEB_STATUS_BAD_INPUT = range(5)
DEB_MEM_HOST, DEB_MEM_DEVICE = 0, 1
DEB_SOLOUT_T_EVAL, DEB_SOLOUT_EVEN, DEB_SOLOUT_DEFAULT, DEB_SOLOUT_DENSE, DEB_SOLOUT_CROSSING, DEB_SOLOUT_HYPERPLANE = 0, 1, 2, 3, 4, 5
CROSSING_BOTH, CROSSING_POSITIVE, CROSSING_NEGATIVE = 0, 1, -1  # CrossingDirection, src/solout/mod.rs
DEB_EVENT_NONE, DEB_EVENT_LINEAR = 0, 1
DEB_MAX_DIM = 16
DEB_STATUS_INTERRUPTED = 5
DEB_OK, DEB_ERR_BAD_ARG, DEB_ERR_NO_DEVICE, DEB_ERR_CUDA, DEB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
DEB_ABI_VERSION = 9
DEB_MAX_DEVICES = 16
DEB_LAYOUT_TRAJ_MAJOR, DEB_LAYOUT_ROW_MAJOR = 0, 1
DEB_FILTER_IDENTITY, DEB_FILTER_TRUNCATE_MANTISSA = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class ErkOptions(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("rtol_vec", _dp), ("atol_vec", _dp), ("h0", C.c_double),
                ("h_min", C.c_double), ("h_max", C.c_double), ("max_steps", C.c_int64), ("safety_factor", C.c_double),
                ("min_scale", C.c_double), ("max_scale", C.c_double), ("max_rejects", C.c_int64)]


class OdeProblem(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("system", C.c_int32), ("method", C.c_int32), ("dim", C.c_int32),
                ("n_params", C.c_int32), ("n_traj", C.c_int64), ("y0", C.c_void_p), ("params", C.c_void_p),
                ("params_shared", C.c_int32), ("n_eval", C.c_int32), ("t_eval", _dp), ("t0", C.c_double), ("tf", C.c_double),
                ("opt", ErkOptions), ("device", C.c_int32), ("memspace", C.c_int32), ("stream", C.c_void_p),
                ("solout", C.c_int32), ("dense_n", C.c_int32), ("even_dt", C.c_double),
                ("cross_component", C.c_int32), ("cross_direction", C.c_int32), ("cross_threshold", C.c_double),
                ("event", C.c_int32), ("event_direction", C.c_int32), ("event_terminate", C.c_int32), ("row_capacity", C.c_int32),
                ("event_coef", C.c_double * (DEB_MAX_DIM + 2)),
                ("plane_dim", C.c_int32), ("plane_index", C.c_int32 * DEB_MAX_DIM), ("plane_point", C.c_double * DEB_MAX_DIM),
                ("plane_normal", C.c_double * DEB_MAX_DIM),
                # appended in ABI 9
                ("filter", C.c_int32), ("filter_bits", C.c_int32), ("layout", C.c_int32), ("n_devices", C.c_int32),
                ("devices", C.c_int32 * DEB_MAX_DEVICES)]


class SdeProblem(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("system", C.c_int32), ("method", C.c_int32), ("dim", C.c_int32),
                ("n_params", C.c_int32), ("n_traj", C.c_int64), ("y0", C.c_void_p), ("y0_shared", C.c_int32),
                ("params_shared", C.c_int32), ("params", C.c_void_p), ("n_eval", C.c_int32), ("t_eval", _dp),
                ("t0", C.c_double), ("tf", C.c_double), ("opt", ErkOptions), ("seed", C.c_uint64), ("path_offset", C.c_int64),
                ("device", C.c_int32), ("memspace", C.c_int32), ("stream", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("y_eval", C.c_void_p), ("n_emitted", C.c_void_p), ("t_final", C.c_void_p),
                ("y_final", C.c_void_p), ("status", C.c_void_p), ("accepted", C.c_void_p), ("rejected", C.c_void_p),
                ("evals", C.c_void_p), ("t_rows", _dp), ("n_rows", C.c_int32), ("kernel_ms", C.c_float),
                ("total_ms", C.c_float), ("t_out", C.c_void_p),
                # appended in ABI 9
                ("stats_sums", C.c_void_p), ("stats_counts", C.c_void_p), ("gpu_launches", C.c_int32), ("reserved0", C.c_int32)]


class HeatProblem(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("n_nodes", C.c_int64), ("lo", C.c_double), ("hi", C.c_double),
                ("alpha", C.c_double), ("bc_lower_kind", C.c_int32), ("bc_upper_kind", C.c_int32),
                ("bc_lower_value", C.c_double), ("bc_upper_value", C.c_double), ("method", C.c_int32), ("h", C.c_double),
                ("t0", C.c_double), ("tf", C.c_double), ("max_steps", C.c_int64), ("u0", C.c_void_p), ("u_final", C.c_void_p),
                ("t_final", _dp), ("steps", C.POINTER(C.c_int64)), ("status", _ip), ("device", C.c_int32),
                ("memspace", C.c_int32), ("stream", C.c_void_p)]


# every symbol include/deb_ensemble.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = ["deb_abi_version", "deb_last_error", "deb_device_count", "deb_erk_options_default", "deb_define_ode", "deb_define_event",
               "deb_define_sde", "deb_check_sde", "deb_define_ode_sensitivity", "deb_launch_count", "deb_check_ode", "deb_trim_memory", "deb_solve_ode",
               "deb_solve_sde", "deb_solve_heat_mol", "deb_heat_rhs", "deb_ensemble_stats", "deb_malloc", "deb_free", "deb_memcpy_h2d",
               "deb_memcpy_d2h", "deb_synchronize", "deb_pow_device", "deb_fp64_issue_peak", "deb_plan_fixed_steps", "deb_shard_layout"]

_lib = None


class ExtensionMissing(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen libdeb200.so (built in-tree by __graft_entry__.build()).  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DEB200_LIB", LIB_PATH)  # (override used to A/B alternative builds of the same library)
    if not os.path.exists(path):
        raise ExtensionMissing(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the ensemble kernels)")
    lib = C.CDLL(path)
    lib.deb_last_error.restype = C.c_char_p
    lib.deb_solve_ode.argtypes = [C.POINTER(OdeProblem), C.POINTER(Result)]
    lib.deb_solve_sde.argtypes = [C.POINTER(SdeProblem), C.POINTER(Result)]
    lib.deb_solve_heat_mol.argtypes = [C.POINTER(HeatProblem)]
    lib.deb_define_ode.argtypes = [C.c_int32, C.c_int32, C.c_char_p, _ip]
    lib.deb_check_ode.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.deb_define_event.argtypes = [C.c_int32, C.c_char_p, _ip]
    lib.deb_define_sde.argtypes = [C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_char_p, _ip]
    lib.deb_check_sde.argtypes = [C.c_int32, C.c_int32]
    lib.deb_define_ode_sensitivity.argtypes = [C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_char_p, _ip]
    lib.deb_launch_count.restype = C.c_int64
    lib.deb_heat_rhs.argtypes = [C.POINTER(HeatProblem), C.c_void_p, C.c_void_p]
    lib.deb_erk_options_default.argtypes = [C.POINTER(ErkOptions)]
    lib.deb_erk_options_default.restype = None
    lib.deb_ensemble_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p]
    lib.deb_pow_device.argtypes = [_dp, C.c_double, C.c_int64, _dp, C.c_int32]
    lib.deb_fp64_issue_peak.argtypes = [C.c_int32, C.c_int32, _dp, C.POINTER(C.c_float)]
    lib.deb_shard_layout.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.deb_plan_fixed_steps.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int32), _dp, C.POINTER(C.c_int32)]
    lib.deb_malloc.argtypes = [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.deb_free.argtypes = [C.c_int32, C.c_void_p]
    lib.deb_memcpy_h2d.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.deb_memcpy_d2h.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.deb_synchronize.argtypes = [C.c_int32]
    _lib = lib
    return lib


def _check(lib, rc: int, what: str):
    if rc != DEB_OK:
        msg = lib.deb_last_error().decode("utf-8", "replace")
        if rc == DEB_ERR_NO_DEVICE:
            raise RuntimeError(f"{what}: no CUDA device ({msg}); the product path has no CPU fallback")
        if rc == DEB_ERR_BAD_ARG:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: error {rc}: {msg}")


# ---------------------------------------------------------------------------------------------- errors / status
class Error(Exception):
    """Mirror of `Error<T, Y>` (src/error.rs:13-41)."""


class BadInput(Error):
    pass


class MaxSteps(Error):
    def __init__(self, t, y):
        super().__init__(f"MaxSteps {{ t: {t}, y: {y} }}")
        self.t, self.y = t, y


class StepSize(Error):
    def __init__(self, t, y):
        super().__init__(f"StepSize {{ t: {t}, y: {y} }}")
        self.t, self.y = t, y


class Stiffness(Error):
    def __init__(self, t, y):
        super().__init__(f"Stiffness {{ t: {t}, y: {y} }}")
        self.t, self.y = t, y


@dataclass
class Evals:  # src/stats.rs:16
    function: int = 0


@dataclass
class Steps:  # src/stats.rs:69
    accepted: int = 0
    rejected: int = 0


@dataclass
class Solution:
    """Mirror of `Solution<T, Y>` (src/solution.rs:30-53) for one trajectory."""
    t: np.ndarray
    y: np.ndarray
    status: str
    evals: Evals
    steps: Steps
    t_final: float = math.nan
    y_final: Optional[np.ndarray] = None


# ---------------------------------------------------------------------------------------------- systems
@dataclass
class OdeSystem:
    """A built-in `ODE` (src/ode/ode.rs:20-44): id + parameters (one set, or one row per trajectory)."""
    system_id: int
    dim: int
    params: np.ndarray  # (n_params,) or (N, n_params)


def ExponentialGrowth(k): return OdeSystem(DEB_SYS_EXPONENTIAL, 1, _params(k))            # tests/ode/systems.rs:8-16
def LinearEquation(a, b): return OdeSystem(DEB_SYS_LINEAR, 1, _params(a, b))               # :20-29
def HarmonicOscillator(k): return OdeSystem(DEB_SYS_HARMONIC, 2, _params(k))               # :33-43
def LogisticEquation(k, m): return OdeSystem(DEB_SYS_LOGISTIC, 1, _params(k, m))           # :48-59
def VanDerPolOscillator(mu): return OdeSystem(DEB_SYS_VAN_DER_POL, 2, _params(mu))         # :66-78
def LorenzSystem(sigma, rho, beta): return OdeSystem(DEB_SYS_LORENZ, 3, _params(sigma, rho, beta))  # :85-101
def BrusselatorSystem(a, b): return OdeSystem(DEB_SYS_BRUSSELATOR, 2, _params(a, b))       # :106-120
def RobertsonProblem(): return OdeSystem(DEB_SYS_ROBERTSON, 3, np.zeros(0))                  # :161-173


def ode_from_source(dim: int, diff_body: str, params=(), lib=None) -> OdeSystem:
    """`IVP::ode_from_fn(|t, y, dydt| ..)` (src/ivp.rs:291-317) for the device: the body of
    `void diff(double t, const double* y, double* dydt, const double* p)` as CUDA C++ text, compiled at first use
    (NVRTC).  `params`: one parameter set, or an (N, n_params) array for a sweep."""
    lib = lib or load_library()
    prm = np.ascontiguousarray(params, dtype=np.float64)
    if prm.ndim == 0:
        prm = prm.reshape(1)
    sid = C.c_int32(-1)
    _check(lib, lib.deb_define_ode(int(dim), int(prm.shape[-1]), diff_body.encode(), C.byref(sid)), "deb_define_ode")
    return OdeSystem(sid.value, int(dim), prm)


def check_ode(system: OdeSystem, method, solout: int = DEB_SOLOUT_T_EVAL, event=None, lib=None) -> None:
    """Compile the kernel for (system, method, recorder, event) now (NVRTC, no GPU needed); raises ValueError with the
    compiler log when a user-defined right-hand side or event function does not compile."""
    lib = lib or load_library()
    mid = method.method_id if hasattr(method, "method_id") else int(method)
    eid = 0 if event is None else event.event_id
    _check(lib, lib.deb_check_ode(int(system.system_id), mid, int(solout), int(eid)), "deb_check_ode")


def ode_sensitivity_from_source(dim: int, diff_body: str, jac_y_body: str, jac_p_body: str, params, lib=None) -> OdeSystem:
    """`ForwardSensitivityOde::new(ode, y_proto)` (src/ode/sensitivity/forward.rs:43-116) for the device: the augmented system
    [y, S] with S' = J_y S + J_p (S row-major after y, S[r][c] = dy_r/dp_c) is generated from the bodies of
    `diff(t, y, dydt, p)`, `jacobian(t, y, J, p)` (J[r*dim + k], zeroed) and `jacobian_p(t, y, Jp, p)` (Jp[r*m + c], zeroed).
    The returned system has dimension dim*(1 + len(params)); start it from [y0, 0...]."""
    lib = lib or load_library()
    prm = np.ascontiguousarray(params, dtype=np.float64)
    if prm.ndim == 0:
        prm = prm.reshape(1)
    m = int(prm.shape[-1])
    sid = C.c_int32(-1)
    _check(lib, lib.deb_define_ode_sensitivity(int(dim), m, diff_body.encode(), jac_y_body.encode(), jac_p_body.encode(), C.byref(sid)),
           "deb_define_ode_sensitivity")
    return OdeSystem(sid.value, int(dim) * (1 + m), prm)


@dataclass
class EventSpec:
    """Mirror of an `impl Event` (src/solout/event.rs:60-70): the function g(t, y) whose zero crossings are located."""
    event_id: int
    coef: tuple = ()


def LinearEvent(c0: float, ct: float, cy: Sequence[float]) -> EventSpec:
    """g(t, y) = c0 + ct*t + sum_i cy[i]*y[i], accumulated in that order (`y[0] - 0.9*m` is LinearEvent(-0.9*m, 0, [1]))."""
    return EventSpec(DEB_EVENT_LINEAR, (float(c0), float(ct)) + tuple(float(v) for v in cy))


def event_from_source(dim: int, event_body: str, lib=None) -> EventSpec:
    """`impl Event { fn event(&self, t, y) -> T }` for the device: the body of
    `double event(double t, const double* y, const double* p)` as CUDA C++ text (p = the trajectory's ODE parameters)."""
    lib = lib or load_library()
    eid = C.c_int32(-1)
    _check(lib, lib.deb_define_event(int(dim), event_body.encode(), C.byref(eid)), "deb_define_event")
    return EventSpec(eid.value)


@dataclass
class SdeSystem:
    system_id: int
    params: np.ndarray
    dim: int = 1


def OrnsteinUhlenbeck(theta, mu, sigma): return SdeSystem(DEB_SDE_OU, _params(theta, mu, sigma))
def GeometricBrownianMotion(mu, sigma): return SdeSystem(DEB_SDE_GBM, _params(mu, sigma))
def HestonModel(mu, kappa, theta, sigma, rho):  # examples/sde/02_heston_model/main.rs: y = (price, variance)
    return SdeSystem(DEB_SDE_HESTON, _params(mu, kappa, theta, sigma, rho), 2)


def sde_from_source(dim: int, drift_body: str, diffusion_body: str, params=(), noise_body: Optional[str] = None, lib=None) -> SdeSystem:
    """`impl SDE for S { fn drift; fn diffusion; fn noise }` (src/sde/sde.rs:16-67) for the device: the bodies of
    `drift(t, y, dydt, p)`, `diffusion(t, y, g, p)` (diagonal noise: dY_c += g[c]*dW_c) and, optionally, `noise(dw, p)` which
    mixes the independent N(0, h) increments of the library's Philox stream in place (e.g. Heston's correlation), as CUDA C++
    text.  Compiled at first use (NVRTC)."""
    lib = lib or load_library()
    prm = np.ascontiguousarray(params, dtype=np.float64)
    if prm.ndim == 0:
        prm = prm.reshape(1)
    sid = C.c_int32(-1)
    _check(lib, lib.deb_define_sde(int(dim), int(prm.shape[-1]), drift_body.encode(), diffusion_body.encode(),
                                   noise_body.encode() if noise_body else None, C.byref(sid)), "deb_define_sde")
    return SdeSystem(sid.value, prm, int(dim))


def _params(*cols) -> np.ndarray:
    arrs = [np.asarray(c, dtype=np.float64) for c in cols]
    if all(a.ndim == 0 for a in arrs):
        return np.array([float(a) for a in arrs], dtype=np.float64)
    n = max(a.shape[0] for a in arrs if a.ndim == 1)
    out = np.empty((n, len(arrs)), dtype=np.float64)
    for j, a in enumerate(arrs):
        out[:, j] = a
    return out


# ---------------------------------------------------------------------------------------------- method builder
class ExplicitRungeKutta:
    """Mirror of `ExplicitRungeKutta` constructors and setters (src/methods/erk/mod.rs:135-228)."""

    def __init__(self, method_id: int, h0: float = 0.0):
        self.method_id = method_id
        self._rtol = 1.0e-6
        self._atol = 1.0e-6
        self._h0 = float(h0)
        self._h_min = 0.0
        self._h_max = math.inf
        self._max_steps = 10_000
        self._max_rejects = 100  # read by the adaptive family (rkf45, cash_karp, rkv*) only
        self._safety_factor = 0.9
        self._min_scale = 0.2
        self._max_scale = 10.0
        self._filter = (DEB_FILTER_IDENTITY, 0)

    # constructors: dormandprince/mod.rs:45-58, fixed/mod.rs:41-89
    @classmethod
    def dopri5(cls): return cls(DEB_DOPRI5)
    @classmethod
    def dop853(cls): return cls(DEB_DOP853)
    @classmethod
    def rkf45(cls): return cls(DEB_RKF45)          # adaptive/mod.rs:47-53
    @classmethod
    def cash_karp(cls): return cls(DEB_CASH_KARP)  # adaptive/mod.rs:54-60
    # Verner pairs with a dense-output polynomial, adaptive/mod.rs:59-122
    @classmethod
    def rkv655e(cls): return cls(DEB_RKV655E)
    @classmethod
    def rkv656e(cls): return cls(DEB_RKV656E)
    @classmethod
    def rkv766e(cls): return cls(DEB_RKV766E)
    @classmethod
    def rkv767e(cls): return cls(DEB_RKV767E)
    @classmethod
    def rkv877e(cls): return cls(DEB_RKV877E)
    @classmethod
    def rkv878e(cls): return cls(DEB_RKV878E)
    @classmethod
    def rkv988e(cls): return cls(DEB_RKV988E)
    @classmethod
    def rkv989e(cls): return cls(DEB_RKV989E)
    @classmethod
    def euler(cls, h0): return cls(DEB_EULER, h0)
    @classmethod
    def midpoint(cls, h0): return cls(DEB_MIDPOINT, h0)
    @classmethod
    def heun(cls, h0): return cls(DEB_HEUN, h0)
    @classmethod
    def ralston(cls, h0): return cls(DEB_RALSTON, h0)
    @classmethod
    def ssp_rk3(cls, h0): return cls(DEB_SSP_RK3, h0)
    @classmethod
    def rk4(cls, h0): return cls(DEB_RK4, h0)
    @classmethod
    def three_eighths(cls, h0): return cls(DEB_THREE_EIGHTHS, h0)

    def rtol(self, v): self._rtol = v; return self
    def atol(self, v): self._atol = v; return self
    def h0(self, v): self._h0 = float(v); return self
    def h_min(self, v): self._h_min = float(v); return self
    def h_max(self, v): self._h_max = float(v); return self
    def max_steps(self, v): self._max_steps = int(v); return self
    def max_rejects(self, v): self._max_rejects = int(v); return self
    def safety_factor(self, v): self._safety_factor = float(v); return self
    def min_scale(self, v): self._min_scale = float(v); return self
    def max_scale(self, v): self._max_scale = float(v); return self

    def filter_truncate_mantissa(self, bits: int):
        """`.filter(|h| f64::from_bits(h.to_bits() & MASK))` (erk/mod.rs:225) with MASK keeping the leading `bits` mantissa
        bits: the built-in step-size filter that can cross the C ABI (a Rust fn pointer cannot)."""
        self._filter = (DEB_FILTER_TRUNCATE_MANTISSA, int(bits))
        return self

    def fill_options(self, opt: ErkOptions, dim: int, keep: list):
        def tol(v, name):
            if np.ndim(v) == 0:
                return float(v), None
            arr = np.ascontiguousarray(v, dtype=np.float64)
            if arr.shape != (dim,):
                raise ValueError(f"{name} vector must have {dim} entries")
            keep.append(arr)
            return float(arr[0]), arr.ctypes.data_as(_dp)
        opt.rtol, opt.rtol_vec = tol(self._rtol, "rtol")
        opt.atol, opt.atol_vec = tol(self._atol, "atol")
        opt.h0, opt.h_min, opt.h_max = self._h0, self._h_min, self._h_max
        opt.max_steps = self._max_steps
        opt.safety_factor, opt.min_scale, opt.max_scale = self._safety_factor, self._min_scale, self._max_scale
        opt.max_rejects = self._max_rejects


class Milstein(ExplicitRungeKutta):
    """Mirror of `Milstein::new(h0)` (src/methods/milstein.rs:37-68): derivative-free Milstein for SDE ensembles;
    settings h_min, h_max, max_steps."""

    def __init__(self, h0: float):
        super().__init__(DEB_MILSTEIN, h0)

    @classmethod
    def new(cls, h0):
        return cls(h0)


# ---------------------------------------------------------------------------------------------- results
_STATUS_NAME = {DEB_STATUS_INTERRUPTED: "Interrupted", DEB_STATUS_COMPLETE: "Complete", DEB_STATUS_MAX_STEPS: "MaxSteps", DEB_STATUS_STEP_SIZE: "StepSize",
                DEB_STATUS_STIFFNESS: "Stiffness", DEB_STATUS_BAD_INPUT: "BadInput"}


class EnsembleSolution:
    """All N results of one ensemble solve, as flat arrays plus a per-trajectory `Solution` view."""

    def __init__(self, n, dim, t_rows, y_eval, n_emitted, t_final, y_final, status, accepted, rejected, evals, kernel_ms,
                 total_ms):
        self.n, self.dim = n, dim
        self.t_rows = t_rows            # times of the rows a trajectory can emit, in integration order
        self.y_eval = y_eval            # (N, n_eval, dim); rows >= n_emitted[i] are unspecified
        self.n_emitted = n_emitted
        self.t_final, self.y_final = t_final, y_final
        self.status, self.accepted, self.rejected, self.evals = status, accepted, rejected, evals
        self.kernel_ms, self.total_ms = kernel_ms, total_ms
        self.even_tf = None
        self.t_out = None               # (N, n_eval) per-trajectory row times (per-step recorders), else None

    def __len__(self):
        return self.n

    def __getitem__(self, i) -> Solution:
        """`Result<Solution, Error>` of trajectory i: returns the Solution or raises the reference's Error variant."""
        st = int(self.status[i])
        yf = self.y_final[i].copy()
        tfin = float(self.t_final[i])
        if st == DEB_STATUS_BAD_INPUT:
            raise BadInput("Invalid input")
        if st == DEB_STATUS_MAX_STEPS:
            raise MaxSteps(tfin, yf)
        if st == DEB_STATUS_STEP_SIZE:
            raise StepSize(tfin, yf)
        if st == DEB_STATUS_STIFFNESS:
            raise Stiffness(tfin, yf)
        m = int(self.n_emitted[i])
        if self.t_out is not None and m > self.y_eval.shape[1]:
            raise ValueError(f"trajectory {i} produced {m} rows but the row capacity is {self.y_eval.shape[1]}: raise max_rows")
        ts = self.row_times(i)
        return Solution(t=ts, y=self.y_eval[i, :m].copy(), status=_STATUS_NAME[st],
                        evals=Evals(int(self.evals[i])), steps=Steps(int(self.accepted[i]), int(self.rejected[i])),
                        t_final=tfin, y_final=yf)

    def row_times(self, i) -> np.ndarray:
        """Solution.t of trajectory i."""
        m = int(self.n_emitted[i])
        if self.t_out is not None:
            return self.t_out[i, :min(m, self.t_out.shape[1])].copy()
        ts = np.empty(m)
        k = min(m, self.t_rows.size)
        ts[:k] = self.t_rows[:k]
        if self.even_tf is not None and m > 0 and self.t_final[i] == self.even_tf:
            ts[m - 1] = self.even_tf  # even.rs:166-188: the last point is replaced by / completed with (tf, y(tf))
        elif m > k:
            ts[k:] = np.nan
        return ts

    def status_names(self) -> List[str]:
        return [_STATUS_NAME[int(s)] for s in self.status]


def alloc_result_arrays(n, n_eval, dim, with_times=False):
    extra = dict(t_out=np.full((n, max(n_eval, 0)), np.nan)) if with_times else {}
    return dict(**extra, y_eval=np.full((n, max(n_eval, 0), dim), np.nan), n_emitted=np.zeros(n, np.int32), t_final=np.zeros(n),
                y_final=np.zeros((n, dim)), status=np.full(n, -1, np.int32), accepted=np.zeros(n, np.int32),
                rejected=np.zeros(n, np.int32), evals=np.zeros(n, np.int32))


def bind_result(res: Result, arrs: dict, t_sorted: np.ndarray):
    res.struct_size = C.sizeof(Result)
    for k in ("y_eval", "n_emitted", "t_final", "y_final", "status", "accepted", "rejected", "evals", "t_out", "stats_sums", "stats_counts"):
        a = arrs.get(k)
        setattr(res, k, a.ctypes.data if a is not None and a.size > 0 else None)
    res.t_rows = t_sorted.ctypes.data_as(_dp) if t_sorted.size else None


# ---------------------------------------------------------------------------------------------- the builder
class EnsembleIVP:
    """Mirror of the `IVP` builder (src/ivp.rs) for an ensemble of N problems sharing (t0, tf, method)."""

    def __init__(self, kind, system, t0, tf, y0s, seed=0, path_offset=0):
        self.kind, self.system = kind, system
        self.t0, self.tf = float(t0), float(tf)
        y0s = np.ascontiguousarray(y0s, dtype=np.float64)
        if kind == "ode":
            if y0s.ndim == 1:
                y0s = y0s.reshape(1, -1) if system.dim > 1 or y0s.shape[0] == 1 else y0s.reshape(-1, 1)
            if y0s.ndim != 2 or y0s.shape[1] != system.dim:
                raise ValueError(f"y0 must have shape (N, {system.dim})")
        else:
            y0s = y0s.reshape(-1, system.dim) if system.dim > 1 else y0s.reshape(-1)
        self.y0s = y0s
        self._t_eval = np.zeros(0)
        self._even_dt = 0.0
        self._recorder = None  # (deb_solout, row capacity, dense n, component, threshold, direction)
        self._event = None     # (EventSpec, direction, terminate count or 0, extra rows)
        self._method: Optional[ExplicitRungeKutta] = None
        self._device = 0
        self._devices: List[int] = []
        self._layout = DEB_LAYOUT_TRAJ_MAJOR
        self._stats = False
        self.seed, self.path_offset = int(seed), int(path_offset)

    @classmethod
    def ode(cls, system: OdeSystem, t0, tf, y0s):  # IVP::ode, ivp.rs:279
        return cls("ode", system, t0, tf, y0s)

    @classmethod
    def sde(cls, system: SdeSystem, t0, tf, y0s, seed=0, path_offset=0):  # IVP::sde, ivp.rs:504
        return cls("sde", system, t0, tf, y0s, seed, path_offset)

    def t_eval(self, pts: Sequence[float]):  # ivp.rs:656
        self._t_eval = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1)
        return self

    def even(self, dt: float):  # ivp.rs:643
        """Evenly spaced output t0, t0+dt, ... (EvenSolout, src/solout/even.rs), plus the final state at tf."""
        if self.kind != "ode":
            raise ValueError("even(dt) is implemented for ODE ensembles")
        if not (dt > 0.0):
            raise ValueError("dt must be positive")
        self._even_dt = float(dt)
        self._t_eval = np.zeros(0)
        return self

    # Per-step recorders: the rows of a trajectory depend on its own steps, so each trajectory gets `max_rows` row slots
    # (rows beyond that are counted in n_emitted but not stored) and rows carry their own time (EnsembleSolution.t_out).
    def _per_step(self, mode, max_rows, n=0, comp=0, thr=0.0, direction=0):
        if self.kind != "ode":
            raise ValueError("per-step recorders are implemented for ODE ensembles")
        if int(max_rows) < 1:
            raise ValueError("max_rows must be >= 1")
        self._recorder = (mode, int(max_rows), int(n), int(comp), float(thr), int(direction))
        self._even_dt, self._t_eval = 0.0, np.zeros(0)
        return self

    def every_step(self, max_rows: int):
        """The recorder of a plain `IVP::solve()` (DefaultSolout, src/solout/default.rs): (t0, y0) and every accepted step."""
        return self._per_step(DEB_SOLOUT_DEFAULT, max_rows)

    def dense(self, n: int, max_rows: int):  # ivp.rs:650
        """`IVP::dense(n)` (DenseSolout, src/solout/dense.rs): n-1 interpolated points inside every step + the step end."""
        return self._per_step(DEB_SOLOUT_DENSE, max_rows, n=n)

    def crossing(self, component_idx: int, threshold: float, direction: int = CROSSING_BOTH, max_rows: int = 64):  # ivp.rs:682
        """`IVP::crossing(component, threshold, direction)` (CrossingSolout, src/solout/crossing.rs): only the points where
        the component crosses the threshold, located by the reference's Newton iteration on the dense output."""
        return self._per_step(DEB_SOLOUT_CROSSING, max_rows, comp=component_idx, thr=threshold, direction=direction)

    def hyperplane_crossing(self, point: Sequence[float], normal: Sequence[float], components: Sequence[int],
                            direction: int = CROSSING_BOTH, max_rows: int = 64):  # ivp.rs:695
        """`IVP::hyperplane_crossing(point, normal, extractor, direction)` (HyperplaneCrossingSolout, src/solout/hyperplane.rs):
        rows where the selected state components (the extractor) cross the hyperplane through `point` with normal `normal`."""
        if not (len(point) == len(normal) == len(components)) or len(components) < 1:
            raise ValueError("point, normal and components must have the same, non-zero length")
        self._per_step(DEB_SOLOUT_HYPERPLANE, max_rows, direction=direction)
        self._plane = ([float(v) for v in point], [float(v) for v in normal], [int(v) for v in components])
        return self

    def event(self, ev: EventSpec, direction: int = CROSSING_BOTH, terminate: Optional[int] = None, max_event_rows: int = 16):  # ivp.rs:662
        """`IVP::event(&e)`: wrap the current recorder with event detection (EventWrappedSolout, src/solout/event.rs).
        `direction`/`terminate` are the event's `EventConfig` (terminate=1 is `.terminal()`).  Without a recorder the base is
        the reference's default (every accepted step) and `max_event_rows` is the whole row capacity; with t_eval / even(dt)
        it is the number of extra row slots for event rows."""
        if self.kind != "ode":
            raise ValueError("events are implemented for ODE ensembles")
        self._event = (ev, int(direction), int(terminate or 0), int(max_event_rows))
        return self

    def method(self, m: ExplicitRungeKutta):  # ivp.rs:632
        self._method = m
        return self

    def rtol(self, v):  # ivp.rs:764 (ToleranceConfig forwarding)
        self._method.rtol(v); return self

    def atol(self, v):  # ivp.rs:771
        self._method.atol(v); return self

    def device(self, ordinal: int):
        self._device = int(ordinal)
        return self

    def devices(self, ordinals: Sequence[int]):
        """Split the ensemble over several GPUs inside one solve() (deb_ode_problem.devices): blocks of 4096 trajectories
        dealt round-robin, one host thread and one persistent kernel per device, results exactly as for one device."""
        self._devices = [int(d) for d in ordinals]
        return self

    def layout(self, layout: int):
        """DEB_LAYOUT_TRAJ_MAJOR (default, y_eval[i][r][c]) or DEB_LAYOUT_ROW_MAJOR (y_eval[r][c][i])."""
        self._layout = int(layout)
        return self

    def with_stats(self, on: bool = True):
        """Also reduce the per-t_eval ensemble sums {sum y, sum y^2} and counts on the device(s) (all-reduced across the
        devices of the call): EnsembleSolution.stats_sums (n_eval, dim, 2) and .stats_counts (n_eval)."""
        self._stats = bool(on)
        return self

    def build_problem(self):
        """Assemble the C-ABI problem/result structs for this builder (host buffers).  Returns
        (problem, result, result_arrays, t_sorted, keepalive)."""
        if self._method is None:
            raise ValueError("method(...) must be set before solve()")
        n = int(self.y0s.shape[0])
        keep: list = []
        n_eval = int(self._t_eval.size)
        if self._even_dt > 0.0:
            n_eval = int(math.floor(abs(self.tf - self.t0) / self._even_dt)) + 3  # row capacity per trajectory
        if self._recorder is not None:
            n_eval = self._recorder[1]
        rows_cap = n_eval
        if self.kind == "ode" and self._event is not None:
            if self._recorder is None and self._even_dt <= 0.0 and self._t_eval.size == 0:
                self._recorder = (DEB_SOLOUT_DEFAULT, self._event[3], 0, 0, 0.0, 0)  # plain solve().event(): every step + events
                n_eval = rows_cap = self._event[3]
            elif self._recorder is None:
                rows_cap = n_eval + self._event[3]
        res = Result()
        t_sorted = np.zeros(max(n_eval, 1))
        if self.kind == "ode":
            sysm = self.system
            dim = sysm.dim
            P = OdeProblem()
            P.struct_size = C.sizeof(OdeProblem)
            P.system, P.method, P.dim = sysm.system_id, self._method.method_id, dim
            params = np.ascontiguousarray(sysm.params, dtype=np.float64)
        else:
            dim = self.system.dim
            P = SdeProblem()
            P.struct_size = C.sizeof(SdeProblem)
            P.system, P.method, P.dim = self.system.system_id, self._method.method_id, dim
            params = np.ascontiguousarray(self.system.params, dtype=np.float64)
            P.y0_shared = 0
            P.seed, P.path_offset = self.seed, self.path_offset
        if params.ndim == 2 and params.shape[0] != n:
            raise ValueError("params must have one row per trajectory")
        P.params_shared = 1 if params.ndim == 1 else 0
        P.n_params = params.shape[-1]
        P.y0 = self.y0s.ctypes.data
        P.n_traj = n
        P.params = params.ctypes.data
        P.n_eval = n_eval
        P.t_eval = self._t_eval.ctypes.data_as(_dp) if self._t_eval.size else None
        P.t0, P.tf = self.t0, self.tf
        if self.kind == "ode" and self._even_dt > 0.0:
            P.solout, P.even_dt = DEB_SOLOUT_EVEN, self._even_dt
        if self.kind == "ode" and self._recorder is not None:
            P.solout, _, P.dense_n, P.cross_component, P.cross_threshold, P.cross_direction = self._recorder
            if P.solout == DEB_SOLOUT_HYPERPLANE:
                pt, nrm, comps = self._plane
                P.plane_dim = len(comps)
                for q in range(len(comps)):
                    P.plane_index[q], P.plane_point[q], P.plane_normal[q] = comps[q], pt[q], nrm[q]
        if self.kind == "ode" and self._event is not None:
            ev, P.event_direction, P.event_terminate, _ = self._event
            P.event, P.row_capacity = ev.event_id, rows_cap
            for q, v in enumerate(ev.coef):
                P.event_coef[q] = v
        self._method.fill_options(P.opt, dim, keep)
        P.device, P.memspace, P.stream = self._device, DEB_MEM_HOST, None
        arrs = alloc_result_arrays(n, rows_cap, dim, with_times=(self._recorder is not None or (self.kind == "ode" and self._event is not None)))
        if self.kind == "ode":
            P.filter, P.filter_bits = self._method._filter
            P.layout = self._layout
            if self._layout == DEB_LAYOUT_ROW_MAJOR:
                arrs["y_eval"] = np.full((max(rows_cap, 0), dim, n), np.nan)
            P.n_devices = len(self._devices)
            for q, d in enumerate(self._devices):
                P.devices[q] = d
            if self._stats:
                arrs["stats_sums"] = np.zeros((max(rows_cap, 0), dim, 2))
                arrs["stats_counts"] = np.zeros(max(rows_cap, 0), np.int64)
        bind_result(res, arrs, t_sorted)
        keep += [params, self.y0s, self._t_eval]
        return P, res, arrs, t_sorted, keep

    def solve(self, lib=None) -> EnsembleSolution:  # ivp.rs:781 / :857
        """Run the ensemble on the GPU through the C ABI (deb_solve_ode / deb_solve_sde)."""
        lib = lib or load_library()
        P, res, arrs, t_sorted, keep = self.build_problem()
        if self.kind == "ode":
            _check(lib, lib.deb_solve_ode(C.byref(P), C.byref(res)), "deb_solve_ode")
        else:
            _check(lib, lib.deb_solve_sde(C.byref(P), C.byref(res)), "deb_solve_sde")
        rows = t_sorted[:res.n_rows].copy()
        return self.wrap_result(arrs, rows, res)


    def wrap_result(self, arrs, rows, res) -> EnsembleSolution:
        n = int(self.y0s.shape[0])
        dim = self.system.dim
        sol = EnsembleSolution(n, dim, rows, arrs["y_eval"], arrs["n_emitted"], arrs["t_final"], arrs["y_final"],
                               arrs["status"], arrs["accepted"], arrs["rejected"], arrs["evals"], res.kernel_ms, res.total_ms)
        if self._even_dt > 0.0:
            sol.even_tf = self.tf  # EvenSolout: a trajectory that lands exactly on tf has its last row at tf
        sol.t_out = arrs.get("t_out")
        sol.stats_sums, sol.stats_counts = arrs.get("stats_sums"), arrs.get("stats_counts")
        sol.gpu_launches = int(res.gpu_launches)
        if self.kind == "ode" and self._layout == DEB_LAYOUT_ROW_MAJOR:
            sol.y_eval_row_major = sol.y_eval                       # (n_eval, dim, N) as the library wrote it
            sol.y_eval = np.ascontiguousarray(np.transpose(sol.y_eval, (2, 0, 1)))  # per-trajectory view for Solution
        return sol


def _even_rows(t0: float, tf: float, dt: float) -> np.ndarray:
    """The points EvenSolout visits: t0, then repeatedly + dt*direction (accumulated), while not past tf."""
    d = math.copysign(1.0, tf - t0)
    rows, ti = [], t0
    while (ti <= tf) if d > 0 else (ti >= tf):
        rows.append(ti)
        ti += dt * d
    return np.asarray(rows, dtype=np.float64)


def _plan_rows(t_eval: np.ndarray, t0: float, tf: float) -> np.ndarray:
    """Times of the rows TEvalSolout can emit (t_eval.rs:87-171): sorted by direction; points not after t0 are
    consumed by the first solout call, which emits only a leading point equal to t0."""
    if t_eval.size == 0:
        return np.zeros(0)
    fwd = tf >= t0
    pts = np.sort(t_eval, kind="stable") if fwd else np.sort(t_eval, kind="stable")[::-1]
    rows = []
    for i, v in enumerate(pts):
        if i == 0 and v == t0:
            rows.append(v)
        elif (v > t0) if fwd else (v < t0):
            rows.append(v)
    return np.asarray(rows, dtype=np.float64)


# ---------------------------------------------------------------------------------------------- method of lines (heat)
@dataclass
class HeatSolution:
    u: np.ndarray
    t: float
    steps: int
    status: str


def build_heat_problem(u0, lo, hi, alpha, method: ExplicitRungeKutta, t0, tf, bc_lower, bc_upper, device=0):
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    out = np.empty_like(u0)
    P = HeatProblem()
    P.struct_size = C.sizeof(HeatProblem)
    P.n_nodes, P.lo, P.hi, P.alpha = u0.size, lo, hi, alpha
    kinds = {"dirichlet": 0, "neumann": 1}
    P.bc_lower_kind, P.bc_lower_value = kinds[bc_lower[0]], bc_lower[1]
    P.bc_upper_kind, P.bc_upper_value = kinds[bc_upper[0]], bc_upper[1]
    P.method, P.h, P.t0, P.tf, P.max_steps = method.method_id, method._h0, t0, tf, method._max_steps
    P.u0, P.u_final = u0.ctypes.data, out.ctypes.data
    tfin, steps, status = C.c_double(0), C.c_int64(0), C.c_int32(-1)
    P.t_final, P.steps, P.status = C.pointer(tfin), C.pointer(steps), C.pointer(status)
    P.device, P.memspace, P.stream = device, DEB_MEM_HOST, None
    return P, out, (tfin, steps, status), [u0]


def solve_heat_mol(u0, lo, hi, alpha, method: ExplicitRungeKutta, t0, tf, bc_lower=("dirichlet", 0.0),
                   bc_upper=("dirichlet", 0.0), device=0, lib=None) -> HeatSolution:
    """IVP::pde(&heat, t0, tf, u0).space(MethodOfLines::finite_difference(grid).boundary(bc)).method(rk4(h)).solve()
    (src/ivp.rs:419,713; tests/pde/method_of_lines.rs:37-70) on one GPU."""
    lib = lib or load_library()
    P, out, (tfin, steps, status), keep = build_heat_problem(u0, lo, hi, alpha, method, t0, tf, bc_lower, bc_upper, device)
    _check(lib, lib.deb_solve_heat_mol(C.byref(P)), "deb_solve_heat_mol")
    return HeatSolution(out, tfin.value, steps.value, _STATUS_NAME.get(status.value, "?"))


def heat_rhs(u, lo, hi, alpha, bc_lower=("dirichlet", 0.0), bc_upper=("dirichlet", 0.0), device=0, lib=None) -> np.ndarray:
    """`MethodOfLines::finite_difference(grid).boundary(bc).discretize(&heat).diff(t, &y, &mut dydt)`
    (tests/pde/method_of_lines.rs:73-157) evaluated on the GPU."""
    lib = lib or load_library()
    u = np.ascontiguousarray(u, dtype=np.float64)
    P, _out, _s, keep = build_heat_problem(u, lo, hi, alpha, ExplicitRungeKutta.euler(0.0), 0.0, 1.0, bc_lower, bc_upper, device)
    du = np.empty_like(u)
    _check(lib, lib.deb_heat_rhs(C.byref(P), u.ctypes.data, du.ctypes.data), "deb_heat_rhs")
    return du


# ---------------------------------------------------------------------------------------------- ensemble front end
def splitmix64_uniform(seed: int, index: np.ndarray) -> np.ndarray:
    """u_k = (splitmix64(seed + k*golden) >> 11) * 2^-53 - 0.5 for the (1-based) stream positions in `index`:
    counter-based, so any shard of an ensemble can be generated without the rest."""
    k = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 - 0.5


def perturbed_ensemble(center, traj_index: np.ndarray, seed: int = 2026) -> np.ndarray:
    """Initial states y0_i = center + (u_{d*i}, ..., u_{d*i+d-1}) for the global trajectory numbers `traj_index`
    (perturbation uniform in [-0.5, 0.5)^d): the ensemble of configs C1/C2."""
    center = np.asarray(center, dtype=np.float64)
    d = center.size
    idx = np.asarray(traj_index, dtype=np.uint64)
    k = (idx[:, None] * np.uint64(d) + np.arange(d, dtype=np.uint64)[None, :]).reshape(-1) + np.uint64(1)
    return center[None, :] + splitmix64_uniform(seed, k).reshape(idx.size, d)


def shard_indices(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global trajectory numbers owned by `rank`: interleaved (i mod world == rank), so that a sorted parameter sweep
    (config C3) is balanced across GPUs; trajectories are independent, there is no exchange during integration."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def allreduce_ensemble_stats(sums, counts, dist=None):
    """The only cross-GPU step: sum the per-t_eval {sum y, sum y^2} and counts over ranks (NCCL on GPUs, gloo in the
    CPU tests).  `sums`/`counts` are torch tensors, reduced in place; a no-op for a single process."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums)
        dist.all_reduce(counts)
    return sums, counts


def stats_to_mean_var(sums, counts):
    """mean and (population) variance per t_eval row and component from the reduced sums."""
    s = np.asarray(sums, dtype=np.float64)
    c = np.maximum(np.asarray(counts, dtype=np.float64), 1.0)[:, None]
    mean = s[:, :, 0] / c
    return mean, s[:, :, 1] / c - mean * mean
