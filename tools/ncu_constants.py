#!/usr/bin/env python3
"""Turn the ncu CSV logs of this round into profiles/r02_measured_constants.json, the file bench.py reads for the numbers that
only a profiler can give (DRAM bytes of the flagship launch, warp instructions per path-step of the Euler-Maruyama kernel).

    tools/ncu_constants.py --dram profiles/r02_dram_traffic_10M.csv --n-traj 10000000 \
                           --sde profiles/r02_ncu_sde_inst.csv --paths 10000000 --steps-ou 1001 --steps-gbm 1000
"""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    for r in rd:
        yield dict(zip(hdr, r))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dram")
    ap.add_argument("--n-traj", type=int, default=10_000_000)
    ap.add_argument("--sde")
    ap.add_argument("--paths", type=int, default=10_000_000)
    ap.add_argument("--steps-ou", type=int, default=1001)
    ap.add_argument("--steps-gbm", type=int, default=1000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_measured_constants.json"))
    a = ap.parse_args()
    out = json.load(open(a.out)) if os.path.exists(a.out) else {}
    if a.dram:
        rd = wr = None
        for r in rows(a.dram):
            if "dp_ensemble_kernel" not in r["Kernel Name"]:
                continue
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"].lower()
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
            if r["Metric Name"] == "dram__bytes_read.sum" and rd is None:
                rd = v
            if r["Metric Name"] == "dram__bytes_write.sum" and wr is None:
                wr = v
        out["c2_dram_bytes_per_trajectory"] = {"read": rd / a.n_traj, "write": wr / a.n_traj, "total": (rd + wr) / a.n_traj,
                                               "launch_read_bytes": rd, "launch_write_bytes": wr, "n_traj": a.n_traj,
                                               "source": os.path.relpath(a.dram, ROOT)}
    if a.sde:
        inst = {}
        for r in rows(a.sde):
            if r["Metric Name"] != "smsp__inst_executed.sum":
                continue
            name = r["Kernel Name"]
            which = "ou" if "SdeOU" in name else "gbm" if "SdeGBM" in name else None
            if which and which not in inst:
                steps = a.steps_ou if which == "ou" else a.steps_gbm
                inst[which] = float(r["Metric Value"].replace(",", "")) / (a.paths / 32.0 * steps)
        out["c4_warp_inst_per_path_step"] = dict(inst, source=os.path.relpath(a.sde, ROOT), paths=a.paths)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
