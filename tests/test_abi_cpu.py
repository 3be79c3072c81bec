"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/deb_ensemble.h declares, agrees with the
ctypes mirrors on struct layout, validates arguments, and -- without a CUDA device -- refuses to compute instead of
falling back to a CPU path."""
import ctypes as C
import importlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
deb = importlib.import_module("differential-equations_b200")
E = deb.ExplicitRungeKutta


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(deb.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    return deb.load_library()


def header_functions():
    txt = open(os.path.join(ROOT, "include", "deb_ensemble.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(deb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 15
    assert sorted(names) == sorted(deb.ABI_SYMBOLS), "ABI_SYMBOLS must list exactly what the header declares"
    for n in names:
        getattr(lib, n)  # AttributeError if the symbol is not exported
    assert lib.deb_abi_version() == deb.DEB_ABI_VERSION
    # the symbols are plain C (no mangling) and nothing else leaks a torch/C++ type in its name
    out = subprocess.run(["nm", "-D", "--defined-only", deb.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_struct_layout_matches_header(lib):
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "deb_ensemble.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(deb_erk_options), sizeof(deb_ode_problem), sizeof(deb_sde_problem), sizeof(deb_result), sizeof(deb_heat_problem));
  printf("%zu %zu %zu %zu\n", offsetof(deb_ode_problem, opt), offsetof(deb_ode_problem, device), offsetof(deb_result, n_rows), offsetof(deb_heat_problem, status));
  printf("%zu\n", offsetof(deb_ode_problem, even_dt));
  printf("%zu %zu\n", offsetof(deb_sde_problem, seed), offsetof(deb_sde_problem, device));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(deb_ode_problem, dense_n), offsetof(deb_ode_problem, cross_threshold), offsetof(deb_ode_problem, event),
         offsetof(deb_ode_problem, row_capacity), offsetof(deb_ode_problem, event_coef), offsetof(deb_ode_problem, plane_dim), offsetof(deb_ode_problem, plane_normal));
  printf("%zu\n", offsetof(deb_result, t_out));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [C.sizeof(deb.ErkOptions), C.sizeof(deb.OdeProblem), C.sizeof(deb.SdeProblem), C.sizeof(deb.Result), C.sizeof(deb.HeatProblem),
            deb.OdeProblem.opt.offset, deb.OdeProblem.device.offset, deb.Result.n_rows.offset, deb.HeatProblem.status.offset,
            deb.SdeProblem.seed.offset, deb.SdeProblem.device.offset]
    want.insert(9, deb.OdeProblem.even_dt.offset)
    want += [deb.OdeProblem.dense_n.offset, deb.OdeProblem.cross_threshold.offset, deb.OdeProblem.event.offset, deb.OdeProblem.row_capacity.offset,
             deb.OdeProblem.event_coef.offset, deb.OdeProblem.plane_dim.offset, deb.OdeProblem.plane_normal.offset, deb.Result.t_out.offset]
    assert got == want


def test_defaults_match_reference(lib):
    """erk/mod.rs:135-144."""
    o = deb.ErkOptions()
    lib.deb_erk_options_default(C.byref(o))
    assert (o.rtol, o.atol, o.h0, o.h_min, o.max_steps, o.safety_factor, o.min_scale, o.max_scale) == (1e-6, 1e-6, 0.0, 0.0, 10000, 0.9, 0.2, 10.0)
    assert o.h_max == float("inf") and not o.rtol_vec and not o.atol_vec and o.max_rejects == 100
    m = E.dopri5()
    assert (m._rtol, m._atol, m._max_steps, m._max_rejects, m._safety_factor, m._min_scale, m._max_scale) == (1e-6, 1e-6, 10000, 100, 0.9, 0.2, 10.0)


def test_argument_validation_needs_no_device(lib):
    ivp = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, np.ones((4, 3))).t_eval([0.5]).method(E.dopri5())
    P, R, arrs, ts, keep = ivp.build_problem()
    P.struct_size = 8
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG and b"struct_size" in lib.deb_last_error()
    P, R, arrs, ts, keep = ivp.build_problem()
    P.system = 99
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG
    P, R, arrs, ts, keep = ivp.build_problem()
    P.method = 9
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_UNSUPPORTED
    P, R, arrs, ts, keep = ivp.build_problem()
    P.dim = 2
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG
    P, R, arrs, ts, keep = ivp.build_problem()
    P.y0 = None
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG
    bad = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, np.ones((4, 3))).t_eval([0.5, float("nan")]).method(E.dopri5())
    P, R, arrs, ts, keep = bad.build_problem()
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG and b"NaN" in lib.deb_last_error()
    assert lib.deb_solve_ode(None, None) == deb.DEB_ERR_BAD_ARG
    with pytest.raises(ValueError):
        deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, np.ones((4, 2)))
    with pytest.raises(ValueError):
        deb.EnsembleIVP.ode(deb.VanDerPolOscillator(np.ones(3)), 0.0, 1.0, np.ones((4, 2))).method(E.dopri5()).build_problem()


def test_empty_ensemble_and_t_eval_plan_without_device(lib):
    """n_traj == 0 returns OK with the row plan filled in (no kernel, no device)."""
    ivp = deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 0.0, 10.0, np.zeros((0, 2))).t_eval([3.0, 0.0, 11.0, -1.0, 0.5, 10.0, 0.5]).method(E.dopri5())
    s = ivp.solve(lib)
    assert len(s) == 0 and s.t_rows.tolist() == [0.5, 0.5, 3.0, 10.0, 11.0]
    assert s.t_rows.tolist() == deb._plan_rows(ivp._t_eval, 0.0, 10.0).tolist()
    back = deb.EnsembleIVP.ode(deb.HarmonicOscillator(1.0), 10.0, 0.0, np.zeros((0, 2))).t_eval([3.0, 10.0, 11.0, -1.0, 0.5]).method(E.dopri5())
    # sorted descending: 11 lies before t0 and is skipped; 10.0 == t0 is then NOT the first point any more, so the
    # reference never emits it (t_eval.rs:100-129)
    assert back.solve(lib).t_rows.tolist() == [3.0, 0.5, -1.0]


def test_user_rhs_compiles_for_every_method_without_a_device(lib):
    """deb_check_ode: the run-time (NVRTC) instantiation of the kernels for a user-defined right-hand side, for one method
    of every kernel family -- this is a compile, no device and no compute.  A broken body reports the compiler log."""
    E = deb.ExplicitRungeKutta
    duffing = deb.ode_from_source(2, "dydt[0] = y[1]; dydt[1] = -p[0]*y[1] - y[0]*y[0]*y[0] + p[1]*cos(t);", params=[0.2, 0.3])
    for m in (E.dopri5(), E.dop853(), E.rkf45(), E.rkv655e(), E.rkv989e(), E.rk4(0.01)):
        deb.check_ode(duffing, m)
    bad = deb.ode_from_source(1, "dydt[0] = undefined_symbol * y[0];", params=[1.0])
    with pytest.raises(ValueError, match="did not compile"):
        deb.check_ode(bad, E.dopri5())
    deb.check_ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), E.dopri5())  # compiled ahead of time
    # per-step recorders are compiled at first use, for built-in and user-defined systems alike
    for m in (E.dopri5(), E.dop853(), E.cash_karp(), E.rkv767e(), E.heun(0.01)):
        deb.check_ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), m, deb.DEB_SOLOUT_DENSE)
    deb.check_ode(duffing, E.dopri5(), deb.DEB_SOLOUT_CROSSING)
    # event detection wraps any recorder: linear event and an `impl Event` body
    near_rest = deb.event_from_source(2, "return fmax(fabs(y[0]), fabs(y[1])) - 0.01;")
    deb.check_ode(deb.HarmonicOscillator(1.0), E.rkf45(), deb.DEB_SOLOUT_T_EVAL, deb.LinearEvent(0.0, 0.0, [1.0, 0.0]))
    deb.check_ode(duffing, E.dop853(), deb.DEB_SOLOUT_EVEN, near_rest)
    deb.check_ode(deb.HarmonicOscillator(1.0), E.rk4(0.01), deb.DEB_SOLOUT_DEFAULT, near_rest)
    with pytest.raises(ValueError, match="did not compile"):
        deb.check_ode(deb.HarmonicOscillator(1.0), E.dopri5(), deb.DEB_SOLOUT_T_EVAL, deb.event_from_source(2, "return nonsense(y[0]);"))


def test_no_cpu_fallback_without_a_device(lib):
    """On a box without a GPU every compute entry point must fail loudly with DEB_ERR_NO_DEVICE."""
    if lib.deb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ivp = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, np.ones((4, 3))).method(E.dopri5())
    P, R, arrs, ts, keep = ivp.build_problem()
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.deb_last_error()
    assert (arrs["status"] == -1).all()  # nothing was computed
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ivp.solve(lib)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        deb.EnsembleIVP.sde(deb.OrnsteinUhlenbeck(0.5, 1.0, 0.3), 0.0, 1.0, np.ones(4)).method(E.euler(0.01)).solve(lib)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        deb.solve_heat_mol(np.ones(8), 0.0, 1.0, 0.1, E.rk4(1e-3), 0.0, 0.01, lib=lib)
    v = C.c_double(0)
    assert lib.deb_fp64_issue_peak(0, 0, C.byref(v), None) == deb.DEB_ERR_NO_DEVICE


def test_product_never_references_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "differential-equations_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs", ".toml")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle_binding" not in txt and "orc_" not in txt and "oracle/" not in txt.replace("the oracle keeps its own copy: oracle/", ""), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", deb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


# ------------------------------------------------------------------------------------------ ABI 9
def test_abi9_struct_tail_matches_header(lib):
    """Offsets of every field appended in ABI 9 (device list, layout, filter; fused statistics, launch count)."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "deb_ensemble.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %d\n", offsetof(deb_ode_problem, filter), offsetof(deb_ode_problem, filter_bits), offsetof(deb_ode_problem, layout),
         offsetof(deb_ode_problem, n_devices), offsetof(deb_ode_problem, devices), DEB_MAX_DEVICES);
  printf("%zu %zu %zu %d\n", offsetof(deb_result, stats_sums), offsetof(deb_result, stats_counts), offsetof(deb_result, gpu_launches), DEB_ABI_VERSION);
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        got = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()]
    O, R = deb.OdeProblem, deb.Result
    assert got == [O.filter.offset, O.filter_bits.offset, O.layout.offset, O.n_devices.offset, O.devices.offset, deb.DEB_MAX_DEVICES,
                   R.stats_sums.offset, R.stats_counts.offset, R.gpu_launches.offset, deb.DEB_ABI_VERSION]


def test_struct_size_rules_need_no_device(lib):
    """A smaller struct_size (older header) is accepted, a larger one (newer header) and a truncated one are refused."""
    ivp = deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), 0.0, 1.0, np.zeros((0, 3))).t_eval([0.5]).method(E.dopri5())
    P, R, arrs, ts, keep = ivp.build_problem()
    P.struct_size = deb.OdeProblem.filter.offset   # ABI 8 size
    R.struct_size = deb.Result.stats_sums.offset
    P.n_devices = 77                                # beyond the declared size: must be ignored
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_OK, lib.deb_last_error()
    assert R.struct_size == deb.Result.stats_sums.offset and R.n_rows == 1
    P.struct_size = C.sizeof(deb.OdeProblem) + 16
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG and b"struct_size" in lib.deb_last_error()
    P.struct_size = deb.OdeProblem.opt.offset       # cut inside the mandatory part
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG
    # the ABI 9 fields are validated
    for field, value, msg in (("filter", 9, b"filter"), ("layout", 3, b"layout"), ("n_devices", 99, b"n_devices")):
        P, R, arrs, ts, keep = ivp.build_problem()
        setattr(P, field, value)
        assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG and msg in lib.deb_last_error(), field
    P, R, arrs, ts, keep = ivp.build_problem()
    P.filter, P.filter_bits = deb.DEB_FILTER_TRUNCATE_MANTISSA, 0
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG and b"filter_bits" in lib.deb_last_error()
    P, R, arrs, ts, keep = ivp.build_problem()
    P.memspace, P.n_devices = deb.DEB_MEM_DEVICE, 2
    P.devices[0], P.devices[1] = 0, 1
    assert lib.deb_solve_ode(C.byref(P), C.byref(R)) == deb.DEB_ERR_BAD_ARG and b"DEB_MEM_HOST" in lib.deb_last_error()


def test_user_sde_and_sensitivity_system_compile_without_a_device(lib, tmp_path, monkeypatch):
    """deb_check_sde / deb_check_ode compile the run-time kernels of a user-defined SDE (drift, diffusion, noise mixing) and
    of a generated forward-sensitivity system; the cubins land in the disk cache and the second compile is a cache hit."""
    monkeypatch.setenv("DEB_CACHE_DIR", str(tmp_path))
    heston = deb.sde_from_source(2, "dydt[0] = p[0] * y[0]; dydt[1] = p[1] * (p[2] - y[1]);", "g[0] = y[0] * sqrt(y[1]); g[1] = p[3] * sqrt(y[1]);",
                                 params=[0.1, 2.0, 0.04, 0.3, -0.7], noise_body="dw[1] = p[4] * dw[0] + sqrt(1.0 - p[4] * p[4]) * dw[1];")
    import time
    t0 = time.perf_counter()
    for mid in (deb.DEB_EULER, deb.DEB_RK4, deb.DEB_MILSTEIN):
        assert lib.deb_check_sde(heston.system_id, mid) == deb.DEB_OK, lib.deb_last_error()
    cold = time.perf_counter() - t0
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 3 and all(f.startswith("deb200-") and f.endswith(".cubin") for f in files)
    t0 = time.perf_counter()
    for mid in (deb.DEB_EULER, deb.DEB_RK4, deb.DEB_MILSTEIN):
        assert lib.deb_check_sde(heston.system_id, mid) == deb.DEB_OK
    assert time.perf_counter() - t0 < cold / 5, "the second compile did not come from the disk cache"
    assert lib.deb_check_sde(heston.system_id, deb.DEB_DOPRI5) == deb.DEB_ERR_UNSUPPORTED
    assert lib.deb_check_sde(deb.DEB_SDE_OU, deb.DEB_EULER) == deb.DEB_OK
    bad = deb.sde_from_source(1, "dydt[0] = nonsense;", "g[0] = 1.0;", params=[1.0])
    assert lib.deb_check_sde(bad.system_id, deb.DEB_EULER) == deb.DEB_ERR_BAD_ARG and b"did not compile" in lib.deb_last_error()
    # forward sensitivities of the logistic equation: z = [y, dy/dk, dy/dm]
    sens = deb.ode_sensitivity_from_source(1, "dydt[0] = p[0] * y[0] * (1.0 - y[0] / p[1]);", "J[0] = p[0] * (1.0 - 2.0 * y[0] / p[1]);",
                                           "Jp[0] = y[0] * (1.0 - y[0] / p[1]); Jp[1] = p[0] * y[0] * y[0] / (p[1] * p[1]);", [1.0, 10.0])
    assert sens.dim == 3
    for m in (E.dopri5(), E.dop853(), E.rk4(0.01)):
        deb.check_ode(sens, m)
    with pytest.raises(ValueError, match="DEB_MAX_DIM"):
        deb.ode_sensitivity_from_source(6, "dydt[0] = 0.0;", "", "", [1.0, 2.0, 3.0])


def test_launch_counter_is_exported(lib):
    assert lib.deb_launch_count() >= 0


def test_rust_shim_structs_follow_the_header_field_for_field():
    """The Rust shim cannot be compiled here (no toolchain): at least keep its #[repr(C)] transcriptions in lock step with
    include/deb_ensemble.h -- same structs, same fields, same order, same array lengths -- and its ABI constant current."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "deb_ensemble.h")).read(), flags=re.S)
    rs = re.sub(r"//[^\n]*", "", open(os.path.join(ROOT, "differential-equations_b200", "rust", "src", "lib.rs")).read())
    consts = {"DEB_MAX_DIM": 16, "DEB_MAX_DEVICES": 16}
    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            rest = re.match(r"^(?:const )?\w+ ?\** ?(.*)$", decl).group(1)  # drop the type (and its pointer stars)
            for part in rest.split(","):
                m = re.match(r"^\*? ?(\w+)(?:\[(.+)\])?$", part.strip())
                out.append((m.group(1), eval(m.group(2), {}, consts) if m.group(2) else None))
        return out
    def rs_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, re.S).group(1)
        out = []
        for m in re.finditer(r"pub (\w+): ([^,\n]+),", body):
            arr = re.match(r"\[\w+; (.+)\]", m.group(2).strip())
            out.append((m.group(1), eval(arr.group(1), {}, consts) if arr else None))
        return out
    for s in ("deb_erk_options", "deb_ode_problem", "deb_sde_problem", "deb_result", "deb_heat_problem"):
        assert rs_fields(s) == c_fields(s), s
    assert "pub const DEB_ABI_VERSION: i32 = %d;" % deb.DEB_ABI_VERSION in rs
    for fn in ("deb_solve_ode", "deb_solve_sde", "deb_solve_heat_mol", "deb_define_ode", "deb_define_event", "deb_define_sde", "deb_define_ode_sensitivity"):
        assert "pub fn %s(" % fn in rs, fn


def build_c_caller(tmpdir):
    """examples/c_caller/lorenz_ensemble.c: a plain-C program against include/deb_ensemble.h and libdeb200.so (gcc only)."""
    exe = os.path.join(str(tmpdir), "lorenz_ensemble")
    pkg = os.path.join(ROOT, "differential-equations_b200")
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_caller", "lorenz_ensemble.c"), "-L", pkg, "-ldeb200", "-Wl,-rpath," + pkg, "-o", exe], check=True)
    return exe


def test_plain_c_caller_links_and_fails_loudly_without_a_device(lib, tmp_path):
    """The boundary is usable from C with nothing but the header and the shared library; without a GPU the call reports
    DEB_ERR_NO_DEVICE (exit code 3 of the example) instead of computing anything on the CPU."""
    exe = build_c_caller(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    import torch
    if torch.cuda.is_available():
        assert out.returncode == 0 and "known answers reproduced" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "no CPU fallback" in out.stderr, out.stdout + out.stderr


def build_cpp_caller(tmpdir):
    """examples/cpp_caller/builder_api.cpp: the C++ host mirror (include/deb_ensemble.hpp) against libdeb200.so."""
    exe = os.path.join(str(tmpdir), "builder_api")
    pkg = os.path.join(ROOT, "differential-equations_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "cpp_caller", "builder_api.cpp"), "-L", pkg, "-ldeb200", "-Wl,-rpath," + pkg, "-o", exe], check=True)
    return exe


def test_cpp_host_mirror_example_links_and_fails_loudly_without_a_device(lib, tmp_path):
    exe = build_cpp_caller(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    import torch
    if torch.cuda.is_available():
        assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "no CPU fallback" in out.stderr, out.stdout + out.stderr


def test_cpp_host_mirror_marshalling_against_the_oracle(tmp_path):
    """include/deb_ensemble.hpp (EnsembleIVP / ExplicitRungeKutta / System / Event / Solution / Error in C++) with the CPU oracle
    behind the C ABI (tests/support/abi_on_oracle.cpp, test infrastructure): recorder and event marshalling, row capacities,
    the per-trajectory Solution view and the Error variants, the reference's known answers."""
    import oracle_binding as ob
    ob.load_oracle()  # builds oracle/liboracle.so when missing
    exe = os.path.join(str(tmp_path), "hpp_check")
    sup = os.path.join(ROOT, "tests", "support")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(sup, "hpp_host_mirror_check.cpp"), os.path.join(sup, "abi_on_oracle.cpp"), "-L", os.path.join(ROOT, "oracle"),
                    "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
