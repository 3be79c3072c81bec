#!/usr/bin/env python3
"""bench.py -- accepted trajectory-steps/s of the Lorenz DOPRI5 f64 ensemble (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W           (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                    (the reference algorithm's CPU path, all host threads)
    python bench.py --gpus N --single-process               (ONE process, the C ABI's device list: deb_ode_problem.devices)

A "step" is one pass of the hot path over the whole ensemble: integrate every trajectory from t0 to tf with DOPRI5
(t_eval at 100 points recorded to HBM), then reduce the per-t_eval ensemble statistics and all-reduce them.
Workload: Lorenz sigma=10 rho=28 beta=8/3, y0 = (1,1,1)+U[-0.5,0.5)^3 from splitmix64(2026),
t in [0,100], dopri5().rtol(1e-8) (atol 1e-6, max_steps 10000, automatic h0), t_eval = 1..100, 10 M trajectories in
total, split evenly across the ranks ("strong" scaling: the metric names a 10 M ensemble at 1/2/4/8 GPUs).

Printed JSON (one line, rank 0): see the contract in the task statement.
  value        inputs resident in HBM, CUDA events on the launch stream, max over ranks
  e2e          the public C-ABI call with pinned HOST buffers: H2D of y0/params and D2H of every result inside the timed region
  roofline     FP64 instruction-issue roofline of the integration kernel against the DADD/DMUL issue peak measured in this run
  gpu_launches kernels the library launched inside the timed region (deb_launch_count, counted by the library itself)
  extra_configs  the other BASELINE.json configurations (C3 DOP853 sweep, C4 Euler-Maruyama, C5 heat) timed in the same run (N = 1)
  cpu_baseline the CPU oracle (a C++ port of the reference algorithm; the Rust crate cannot be built in this image)
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "accepted traj-steps/sec, Lorenz DOPRI5 f64 10M ensemble"
UNIT = "accepted traj-steps/s"
N_EVAL = 100
T0, TF = 0.0, 100.0
# algorithmic DP operations (SURVEY.md 8d / DESIGN.md): 309 per step attempt + 38 per accepted step
OPS_PER_ATTEMPT, OPS_PER_ACCEPT = 309, 38
# constants that can only come from a profiler run are read from this file (written from the ncu captures of the same
# round by tools/ncu_constants.py); a missing file or key reports null, never a literal
PROFILE_CONSTANTS = os.path.join(ROOT, "profiles", "r02_measured_constants.json")
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md: the stated fallback when MEASURED_PEAKS.json is absent


def lorenz_problem(deb, y0, device=0):
    return (deb.EnsembleIVP.ode(deb.LorenzSystem(10.0, 28.0, 8.0 / 3.0), T0, TF, y0)
            .t_eval(np.arange(1.0, N_EVAL + 1.0)).method(deb.ExplicitRungeKutta.dopri5().rtol(1e-8)).device(device))


def profile_constant(key):
    try:
        return json.load(open(PROFILE_CONSTANTS)).get(key)
    except (OSError, ValueError):
        return None


def hbm_peak():
    """(GB/s, source) -- the driver-measured copy bandwidth, else the recipe's stated fallback, labelled as such."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, ValueError, KeyError):
        return HBM_FALLBACK_GBS, "of fallback (B200_PROFILING.md: 6.65 TB/s; MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ NUMA
def bind_near_gpu(torch, local_rank):
    """Pin this process (CPU affinity and the preferred memory node, which pinned allocations follow) to the NUMA node the
    GPU hangs off, when the box has more than one.  Returns a description for the JSON line."""
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        nodes = []
    if len(nodes) < 2:
        return f"single NUMA node ({len(nodes)} listed): nothing to bind"
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    except (OSError, ValueError, AttributeError) as e:
        return f"GPU NUMA node unknown ({e}); not bound"
    if node < 0:
        return "GPU NUMA node reported as -1; not bound"
    try:
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = [c for c in cpus if c in allowed]
        if use:
            os.sched_setaffinity(0, use)
        libc = C.CDLL(None, use_errno=True)
        mask = C.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238  # x86-64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, C.byref(mask), C.c_ulong(64))
        return f"bound to NUMA node {node} of GPU {bdf}: {len(use)} CPUs, set_mempolicy(PREFERRED) rc={rc}"
    except (OSError, ValueError) as e:
        return f"binding to NUMA node {node} failed: {e}"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_rate(ob, deb, n_sample, threads):
    y0 = ob.lorenz_ensemble_y0(n_sample)
    t = time.perf_counter()
    s = ob.oracle_solve(lorenz_problem(deb, y0), threads)
    dt = time.perf_counter() - t
    return float(s.accepted.sum()) / dt, dt, int(s.accepted.sum())


def cpu_baseline(ob, deb, target_s=15.0):
    threads = ob.load_oracle().orc_hardware_threads()
    pilot = max(256, 32 * threads)
    rate, dt, _ = cpu_rate(ob, deb, pilot, threads)
    n = int(min(200_000, max(pilot, pilot * target_s / max(dt, 1e-3))))
    rate, dt, acc = cpu_rate(ob, deb, n, threads)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} trajectories of the same ensemble (same generator, same t_eval), {acc} accepted steps in {dt:.1f} s; "
                      "C++ port of the reference algorithm (oracle/), g++ -O2 -ffp-contract=off, one std::thread per core"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (no Rust toolchain here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    deb = importlib.import_module("differential-equations_b200")
    import oracle_binding as ob
    threads = ob.load_oracle().orc_hardware_threads()
    pilot = max(256, 32 * threads)
    _, dt, _ = cpu_rate(ob, deb, pilot, threads)
    total_steps = args.steps + args.warmup
    n = int(min(200_000, max(pilot, pilot * (120.0 / total_steps) / max(dt, 1e-3))))  # whole run ~2 minutes
    for _ in range(args.warmup):
        cpu_rate(ob, deb, n, threads)
    t = time.perf_counter()
    acc = 0
    for _ in range(args.steps):
        _, _, a = cpu_rate(ob, deb, n, threads)
        acc += a
    dt = time.perf_counter() - t
    value = acc / dt
    sample = (f"each step = first {n} trajectories of the 10M ensemble (same generator, t_eval, options) through the C++ port of the "
              f"reference algorithm (oracle/; the Rust crate cannot be built here), {threads} host threads")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"Lorenz DOPRI5 f64 rtol=1e-8, t in [0,100], t_eval at 100 points; bounded sample of {n} trajectories per step"},
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ------------------------------------------------------------------------------------------------ the other configurations
def extra_configs(local_rank, quick=False):
    """C3, C4 (OU and GBM), C5 at BASELINE.json's sizes, device-resident, best of 2 after one warm-up; each with the roofline that
    bounds its kernel.  ~10 s in total."""
    import bench_configs as bc
    bc.init(local_rank)
    out = {}
    try:
        r = bc.c3(400_000 if quick else 4_000_000, 2)
        out["C3"] = {"workload": "Van der Pol mu in [0.1,50] sweep, DOP853 rtol=atol=1e-8, t in [0,100], %d trajectories" % r["n_traj"],
                     "value": r["accepted_steps_per_s"], "unit": UNIT, "kernel_ms": r["ms"], "accepted": r["accepted"], "rejected": r["rejected"],
                     "complete": r["complete"], "algorithmic_ops": r["algorithmic_ops"], "roofline": r["roofline"]}
    except Exception as e:  # noqa: BLE001 -- an extra line must not take the headline down
        out["C3"] = {"error": repr(e)}
    inst = profile_constant("c4_warp_inst_per_path_step")  # ncu: smsp__inst_executed.sum / (paths * steps / 32)
    for which in ("ou", "gbm"):
        try:
            r = bc.c4(10_000_000 if quick else 100_000_000, which, 2)
            sm_clock = 1.965e9
            issue_peak = 148 * 4 * sm_clock  # warp instructions per second: one per scheduler and cycle
            roof = None
            if inst:
                ach = r["path_steps_per_s"] / 32.0 * inst[which]
                roof = {"bound": "issue", "achieved": ach / 1e12, "peak": issue_peak / 1e12, "unit": "T warp-inst/s", "frac": ach / issue_peak,
                        "warp_inst_per_path_step": inst[which],
                        "note": "no HBM traffic per step and no FP64-only stream (Philox is integer work, Box-Muller mixes FP64/MUFU): "
                                "the bound is instruction issue; instructions per path-step from the ncu capture of this round"}
            out["C4_" + which] = {"workload": "Euler-Maruyama %s, %d paths x %d steps, Philox4x32-10" % (which.upper(), r["n_paths"], r["steps_per_path"]),
                                  "value": r["path_steps_per_s"], "unit": "path-steps/s", "kernel_ms": r["ms"],
                                  "sample_mean": r["mean"], "sde_mean": r["sde_mean"], "sample_var": r["var"], "sde_var": r["sde_var"], "roofline": roof}
        except Exception as e:  # noqa: BLE001
            out["C4_" + which] = {"error": repr(e)}
    try:
        r = bc.c5(20 if quick else 24, 2)
        out["C5"] = {"workload": "heat equation u_t = 0.1 u_xx, N = %d nodes, RK4 h = 1, %d steps (whole deb_solve_heat_mol call)" % (r["n_nodes"], r["steps"]),
                     "value": r["node_steps_per_s"], "unit": "node-steps/s", "kernel_ms": r["ms"], "ms_per_step": r["ms_per_step"],
                     "algorithmic_bytes": 16.0 * r["n_nodes"] * r["steps"], "algorithmic_ops": 52.0 * r["n_nodes"] * r["steps"],
                     "roofline": {"bound": "fp64", "achieved": r["roofline"]["fp64_achieved_Tops"], "peak": r["roofline"]["fp64_peak_Tops"], "unit": "TFLOP/s",
                                  "frac": r["roofline"]["fp64_achieved_Tops"] / r["roofline"]["fp64_peak_Tops"],
                                  "hbm": {"achieved": r["roofline"]["achieved_GBs"], "peak": r["roofline"]["peak_GBs"], "frac": r["roofline"]["frac"]},
                                  "note": "all RK stages of a node on chip: 16 B and 52 DP ops per node and step -> FP64-issue bound, HBM second"}}
    except Exception as e:  # noqa: BLE001
        out["C5"] = {"error": repr(e)}
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-traj", type=int, default=10_000_000, help="total ensemble size (BASELINE config: 10M)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3/C4/C5 lines")
    ap.add_argument("--quick-extra", action="store_true", help="C3/C4/C5 at a tenth of their size (smoke runs)")
    ap.add_argument("--e2e-only", action="store_true", help="diagnostic: only the end-to-end leg (one warm-up device-resident step for the counters)")
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives --gpus devices through the C ABI's device list (e2e only); not the driver's launch mode")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.e2e_only:
        args.steps, args.warmup, args.no_cpu, args.no_extra = 1, 0, True, True

    import torch
    deb = importlib.import_module("differential-equations_b200")
    import oracle_binding as ob  # cpu_baseline leg + input generator only
    lib = deb.load_library()  # raises if the extension is missing: no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ensemble kernels have no CPU fallback")
    if args.single_process:
        return run_single_process(args, torch, deb, lib)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(torch, local_rank)  # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # ---- shard: trajectories split evenly, rank r takes i = r, r+world, ... (interleaved)
    n_total = args.n_traj
    idx = deb.shard_indices(n_total, rank, world)
    n = idx.size
    y0_host = deb.perturbed_ensemble([1.0, 1.0, 1.0], idx, seed=2026)  # only this shard's initial conditions
    params_host = np.array([10.0, 28.0, 8.0 / 3.0])
    t_eval = np.arange(1.0, N_EVAL + 1.0)

    # ---- device-resident buffers (torch = allocator + stream plumbing)
    d_y0 = torch.from_numpy(y0_host).to(dev)
    d_y_eval = torch.empty((n, N_EVAL, 3), dtype=torch.float64, device=dev)
    d_n_emitted = torch.empty(n, dtype=torch.int32, device=dev)
    d_t_final = torch.empty(n, dtype=torch.float64, device=dev)
    d_y_final = torch.empty((n, 3), dtype=torch.float64, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    d_acc = torch.empty(n, dtype=torch.int32, device=dev)
    d_rej = torch.empty(n, dtype=torch.int32, device=dev)
    d_evals = torch.empty(n, dtype=torch.int32, device=dev)
    d_sums = torch.zeros((N_EVAL, 3, 2), dtype=torch.float64, device=dev)
    d_counts = torch.zeros(N_EVAL, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)

    P = deb.OdeProblem()
    P.struct_size = C.sizeof(deb.OdeProblem)
    P.system, P.method, P.dim, P.n_params = deb.DEB_SYS_LORENZ, deb.DEB_DOPRI5, 3, 3
    P.n_traj, P.y0, P.params, P.params_shared = n, d_y0.data_ptr(), params_host.ctypes.data, 1  # shared set: host pointer
    P.n_eval, P.t_eval, P.t0, P.tf = N_EVAL, t_eval.ctypes.data_as(deb._dp), T0, TF
    lib.deb_erk_options_default(C.byref(P.opt))
    P.opt.rtol = 1e-8
    P.device, P.memspace, P.stream = local_rank, deb.DEB_MEM_DEVICE, stream.cuda_stream
    R = deb.Result()
    R.struct_size = C.sizeof(deb.Result)
    R.y_eval, R.n_emitted, R.t_final, R.y_final = d_y_eval.data_ptr(), d_n_emitted.data_ptr(), d_t_final.data_ptr(), d_y_final.data_ptr()
    R.status, R.accepted, R.rejected, R.evals = d_status.data_ptr(), d_acc.data_ptr(), d_rej.data_ptr(), d_evals.data_ptr()

    kernel_events = []

    def step(timed):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = lib.deb_solve_ode(C.byref(P), C.byref(R))           # dp_ensemble_kernel<Lorenz, DOPRI5>
        if rc != 0:
            raise RuntimeError(lib.deb_last_error().decode())
        e1.record(stream)
        rc = lib.deb_ensemble_stats(d_y_eval.data_ptr(), d_n_emitted.data_ptr(), n, N_EVAL, 3, d_sums.data_ptr(), d_counts.data_ptr(),
                                    local_rank, deb.DEB_MEM_DEVICE, stream.cuda_stream)  # stats_partial_kernel + stats_final_kernel
        if rc != 0:
            raise RuntimeError(lib.deb_last_error().decode())
        deb.allreduce_ensemble_stats(d_sums, d_counts, dist)  # the only cross-GPU traffic: 4.8 KB of sums + counts
        if timed:
            kernel_events.append((e0, e1))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.deb_launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        step(True)
    t1.record(stream)
    barrier()
    launches = int(lib.deb_launch_count() - launches0)
    clocks = sampler.stop()  # every rank samples its own GPU; rank 0 reports the worst (lowest median clock, union of reasons)
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, clocks)
        meds = [g["sm_mhz"] for g in gathered if g.get("sm_mhz") is not None]
        clocks = {"sm_mhz": min(meds) if meds else None, "sm_mhz_per_rank": [g.get("sm_mhz") for g in gathered],
                  "sm_max_mhz": max((g["sm_max_mhz"] for g in gathered if g.get("sm_max_mhz")), default=None),
                  "reasons": sorted(set(r for g in gathered for r in g.get("reasons", []))), "samples": sum(g.get("samples", 0) for g in gathered)}
    elapsed_ms = t0.elapsed_time(t1)
    kernel_ms_steps = [a.elapsed_time(b) for a, b in kernel_events]
    kernel_ms = sum(kernel_ms_steps) / len(kernel_ms_steps)

    acc_local = int(d_acc.sum(dtype=torch.int64).item())
    rej_local = int(d_rej.sum(dtype=torch.int64).item())
    n_complete = int((d_status == 0).sum().item())
    red = torch.tensor([acc_local, rej_local, n_complete, launches], dtype=torch.int64, device=dev)
    tmax = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(red)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    acc_total, rej_total, complete_total, launches_total = (int(x) for x in red.tolist())
    elapsed_ms, kernel_ms_max = tmax.tolist()
    ms_per_step = elapsed_ms / args.steps
    value = acc_total / (ms_per_step * 1e-3)

    # ---- roofline of the integration kernel: algorithmic DP ops / kernel time vs the measured DP issue peak
    peak = C.c_double(0)
    pk_ms = C.c_float(0)
    lib.deb_fp64_issue_peak(local_rank, 0, C.byref(peak), C.byref(pk_ms))
    ops_local = OPS_PER_ATTEMPT * (acc_local + rej_local) + OPS_PER_ACCEPT * acc_local
    achieved = ops_local / (kernel_ms * 1e-3) / 1e12
    # the same launch against the HBM roofline (why the bound is not "hbm"): algorithmic bytes = y0 in + rows and finals out
    hbm_gbs, hbm_src = hbm_peak()
    alg_bytes = n * (3 * 8 + N_EVAL * 3 * 8 + 3 * 8 + 8 + 5 * 4)
    hbm_achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    dram_per_traj = profile_constant("c2_dram_bytes_per_trajectory")  # ncu: (dram__bytes_read.sum + dram__bytes_write.sum) / n_traj of the 10 M launch
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak.value / 1e12, "unit": "TFLOP/s",
                "frac": achieved / (peak.value / 1e12),
                "hbm": {"achieved": hbm_achieved, "peak": hbm_gbs, "peak_source": hbm_src, "unit": "GB/s", "frac": hbm_achieved / hbm_gbs,
                        "algorithmic_bytes": alg_bytes,
                        "note": "algorithmic bytes per launch / kernel time: two orders of magnitude below the HBM roofline"},
                "traffic": (dram_per_traj["total"] * n) if dram_per_traj else None,
                "traffic_source": (f"profiles/r02_measured_constants.json: ncu dram__bytes_read.sum + dram__bytes_write.sum per trajectory "
                                   f"({dram_per_traj['read']:.1f} B read + {dram_per_traj['write']:.1f} B written; algorithmic 24 B in + 2452 B out) x {n} trajectories"
                                   if dram_per_traj else "no ncu capture of this round under profiles/"),
                "kernel": "deb::dp_ensemble_kernel<SysLorenz, TabDopri5, 128, 5, shared-params>", "kernel_ms": kernel_ms,
                "kernel_ms_steps_rank0": kernel_ms_steps[:32],
                "peak_source": "measured in this run: register-only DADD/DMUL stream on all SMs (deb_fp64_issue_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry. The reference arithmetic forbids FMA fusion, so the bound is DP "
                               "instruction issue (1 op per instruction), not the 2x DFMA figure",
                "algorithmic_ops": f"{OPS_PER_ATTEMPT}*(accepted+rejected) + {OPS_PER_ACCEPT}*accepted DP ops per launch (rank 0: {ops_local})"}

    # ---- e2e: the public C-ABI call with pinned HOST buffers; H2D + kernel + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        del d_y_eval  # make room: the HOST path keeps its own device mirror of y_eval
        torch.cuda.empty_cache()
        e2e = run_e2e(args, torch, deb, lib, dist, dev, [local_rank], y0_host, params_host, t_eval, n, barrier)
        e2e["numa"] = numa

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        lib.deb_trim_memory(local_rank)
        extra = extra_configs(local_rank, quick=args.quick_extra)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(ob, deb)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": "Lorenz (sigma=10,rho=28,beta=8/3) DOPRI5 f64 rtol=1e-8 atol=1e-6, t in [0,100], t_eval at 100 points, "
                                      f"{n_total} trajectories total split evenly over {world} GPU(s) (BASELINE.json configs[1])",
                          "n_traj_total": n_total, "n_traj_per_gpu": n, "t_eval_points": N_EVAL,
                          "l2": "inputs and outputs larger than L2 (y0 240 MB, y_eval 24 GB per 10M); no reuse between steps",
                          "accepted_steps": acc_total, "rejected_steps": rej_total, "complete_trajectories": complete_total},
               "gpu_launches": launches_total,
               "gpu_launches_how": "deb_launch_count() after - before the timed region, summed over ranks: per step 1 integration kernel + 2 statistics kernels",
               "clocks": clocks, "roofline": roofline}
        if e2e is not None:
            out["e2e"] = e2e
        if extra is not None:
            out["extra_configs"] = extra
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_e2e(args, torch, deb, lib, dist, dev, devices, y0_host, params_host, t_eval, n, barrier):
    """deb_solve_ode(memspace = HOST) on pinned host buffers, statistics included (deb_result.stats_sums)."""
    pin = dict(pin_memory=True)
    h_y0 = torch.from_numpy(y0_host).pin_memory()
    h_y_eval = torch.empty((n, N_EVAL, 3), dtype=torch.float64, **pin)
    h_n_emitted = torch.empty(n, dtype=torch.int32, **pin)
    h_t_final = torch.empty(n, dtype=torch.float64, **pin)
    h_y_final = torch.empty((n, 3), dtype=torch.float64, **pin)
    h_i32 = [torch.empty(n, dtype=torch.int32, **pin) for _ in range(4)]
    h_sums = np.zeros((N_EVAL, 3, 2))
    h_counts = np.zeros(N_EVAL, np.int64)
    PH = deb.OdeProblem()
    PH.struct_size = C.sizeof(deb.OdeProblem)
    PH.system, PH.method, PH.dim, PH.n_params = deb.DEB_SYS_LORENZ, deb.DEB_DOPRI5, 3, 3
    PH.n_traj, PH.y0, PH.params, PH.params_shared = n, h_y0.data_ptr(), params_host.ctypes.data, 1
    PH.n_eval, PH.t_eval, PH.t0, PH.tf = N_EVAL, t_eval.ctypes.data_as(deb._dp), T0, TF
    lib.deb_erk_options_default(C.byref(PH.opt))
    PH.opt.rtol = 1e-8
    PH.device, PH.memspace, PH.stream = devices[0], deb.DEB_MEM_HOST, None
    if len(devices) > 1:
        PH.n_devices = len(devices)
        for q, d in enumerate(devices):
            PH.devices[q] = d
    RH = deb.Result()
    RH.struct_size = C.sizeof(deb.Result)
    RH.y_eval, RH.n_emitted, RH.t_final, RH.y_final = h_y_eval.data_ptr(), h_n_emitted.data_ptr(), h_t_final.data_ptr(), h_y_final.data_ptr()
    RH.status, RH.accepted, RH.rejected, RH.evals = (t.data_ptr() for t in h_i32)
    RH.stats_sums, RH.stats_counts = h_sums.ctypes.data, h_counts.ctypes.data
    h2d = h_y0.numel() * 8 + params_host.size * 8
    d2h = h_y_eval.numel() * 8 + h_t_final.numel() * 8 + h_y_final.numel() * 8 + 5 * n * 4 + h_sums.nbytes + h_counts.nbytes

    def e2e_step():
        rc = lib.deb_solve_ode(C.byref(PH), C.byref(RH))
        if rc != 0:
            raise RuntimeError(lib.deb_last_error().decode())
        if dist is not None:  # ranks: the per-rank statistics are summed like in the device-resident step
            s = torch.from_numpy(h_sums).to(dev); c = torch.from_numpy(h_counts).to(dev)
            deb.allreduce_ensemble_stats(s, c, dist)
    e2e_step()  # warm-up (page-touch of the pinned buffers, allocator, streams)
    e2e_step()
    barrier()
    e_steps = max(5, args.steps)
    launches0 = lib.deb_launch_count()
    w0 = time.perf_counter()
    kernel_ms, lib_ms = [], []
    for _ in range(e_steps):
        e2e_step()
        kernel_ms.append(RH.kernel_ms)
        lib_ms.append(RH.total_ms)
    barrier()
    e_ms = (time.perf_counter() - w0) * 1e3 / e_steps
    launches = int(lib.deb_launch_count() - launches0)
    acc_e = torch.tensor([int(h_i32[1].sum(dtype=torch.int64).item()), launches], dtype=torch.int64, device=dev)
    e_t = torch.tensor([e_ms, sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(acc_e)
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    return {"value": int(acc_e[0].item()) / (e_t[0].item() * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": e_t[0].item(), "kernel_ms": e_t[1].item(), "library_call_ms_rank0": sum(lib_ms) / len(lib_ms), "steps": e_steps,
            "gpu_launches": int(acc_e[1].item()),
            "devices_per_process": len(devices),
            "how": "deb_solve_ode(memspace=HOST) on pinned host buffers: H2D y0+params, ONE persistent kernel per device whose finished "
                   "4096-trajectory blocks are copied out while it runs (completion watermark), statistics reduced on the device; "
                   "wall clock, max over ranks"}


def run_single_process(args, torch, deb, lib):
    """One host process, --gpus devices behind one deb_solve_ode call (deb_ode_problem.devices); the library splits the ensemble
    block-cyclically, runs one thread + persistent kernel per device and all-reduces the statistics with NCCL."""
    n_dev = args.gpus
    if torch.cuda.device_count() < n_dev:
        raise SystemExit(f"--gpus {n_dev} but {torch.cuda.device_count()} device(s) visible")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    n = args.n_traj
    y0_host = deb.perturbed_ensemble([1.0, 1.0, 1.0], np.arange(n), seed=2026)
    params_host = np.array([10.0, 28.0, 8.0 / 3.0])
    t_eval = np.arange(1.0, N_EVAL + 1.0)

    def barrier():
        for d in range(n_dev):
            torch.cuda.synchronize(d)
    sampler = ClockSampler(0)
    sampler.start()
    e2e = run_e2e(args, torch, deb, lib, None, dev, list(range(n_dev)), y0_host, params_host, t_eval, n, barrier)
    clocks = sampler.stop()
    out = {"metric": METRIC, "value": e2e["value"], "unit": UNIT, "n_gpus": n_dev, "steps": e2e["steps"], "warmup": 2,
           "ms_per_step": e2e["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "mode": "single-process: one deb_solve_ode call with a device list (value = the end-to-end rate; no device-resident leg in this mode)",
           "config": {"workload": "Lorenz (sigma=10,rho=28,beta=8/3) DOPRI5 f64 rtol=1e-8 atol=1e-6, t in [0,100], t_eval at 100 points, "
                                  f"{n} trajectories, {n_dev} device(s) behind one C-ABI call (BASELINE.json configs[1])",
                      "n_traj_total": n, "t_eval_points": N_EVAL},
           "gpu_launches": e2e["gpu_launches"], "clocks": clocks, "e2e": e2e}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
