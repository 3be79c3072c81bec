#!/usr/bin/env python3
"""Extract the golden vectors the reference's own tests hold for the explicit-RK path into tests/golden/*.json.

The reference's tests pin this path with SciPy-DOP853-derived constants (tests/ode/accuracy.rs, header comment line 1:
"results of SciPy using DOP853 & Tolerences = 1e-12") and a few known-answer checks.  /root/reference does not exist
on the GPU box, so the numbers are copied out ONCE, here, into a small fixture that travels with the repo:

  tests/ode/accuracy.rs:68-740      final states of 7 systems x explicit RK solvers, with the test's own tolerances
  tests/ode/interpolation.rs:35-79  t_eval([0.5, 1.0, 1.69]) on y' = y, y(t) = e^t, tolerance 1e-3
  tests/ode/from_fn.rs:4-18         Euler h = 0.1 on y' = y over [0,1]: 2.5937 +- 1e-3
  tests/ode/errors.rs:87-130        BadInput for tf == t0 and for h0 > |tf - t0|
  tests/pde/method_of_lines.rs      heat equation KATs (coded directly in the tests; they are structural)

Only cases whose system AND solver the ensemble path implements are kept (explicit RK: DOP853, DOPRI5, RK4,
RKF45, CashKarp, ThreeEighths, Euler, Midpoint, Heun, Ralston, the Verner pairs); implicit, Adams and BDF solvers are out of scope.

Usage: python tools/extract_reference_goldens.py   (writes tests/golden/reference_accuracy.json)
"""
import json
import os
import re

REF = "/root/reference/tests/ode/"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_accuracy.json")

SYSTEMS = {"ExponentialGrowth": "exponential", "LinearEquation": "linear", "HarmonicOscillator": "harmonic",
           "LogisticEquation": "logistic", "RobertsonProblem": "robertson"}
SOLVERS = {"DOP853": "dop853", "DOPRI5": "dopri5", "RKF45": "rkf45", "CashKarp": "cash_karp", "RK4": "rk4", "ThreeEighths": "three_eighths", "Euler": "euler",
           "Midpoint": "midpoint", "Heun": "heun", "Ralston": "ralston",
           "RKV65": "rkv655e", "RKV766e": "rkv766e", "RKV767e": "rkv767e", "RKV877e": "rkv877e", "RKV878e": "rkv878e",
           "RKV988e": "rkv988e", "RKV989e": "rkv989e"}


def vec(s):
    s = s.strip()
    m = re.match(r"vector!\[(.*)\]", s)
    return [float(x) for x in m.group(1).split(",")] if m else [float(s)]


def main():
    src = open(REF + "accuracy.rs").read()
    cases = []
    for b in re.findall(r"test_ode!\s*\{(.*?)\n    \}", src, re.S):
        name = re.search(r"system_name:\s*(\w+)", b).group(1)
        ode = re.search(r"ode:\s*(.*?),\n", b).group(1).strip()
        sysname = re.match(r"(\w+)", ode).group(1)
        if sysname not in SYSTEMS:
            continue
        params = [float(x) for x in re.findall(r":\s*([-0-9.eE]+)", ode)]
        t0 = float(re.search(r"t0:\s*(.*?),", b).group(1))
        tf = float(re.search(r"tf:\s*(.*?),", b).group(1))
        y0 = vec(re.search(r"y0:\s*(.*?),\n", b).group(1))
        expected = vec(re.search(r"expected_result:\s*(.*?),\n", b).group(1))
        for sname, spec, tol in re.findall(r"solver_name:\s*(\w+),\s*solver:\s*(.*?),\s*tolerance:\s*([^,\n]+)", b, re.S):
            if sname not in SOLVERS:
                continue
            spec = re.sub(r"\s+", "", spec)
            m = re.match(r"ExplicitRungeKutta::(\w+)\(([-0-9.eE]*)\)(.*)", spec)
            assert m and m.group(1) == SOLVERS[sname], spec
            opts = dict(re.findall(r"\.(rtol|atol)\(([-0-9.eE]+)\)", m.group(3)))
            cases.append({"case": name, "system": SYSTEMS[sysname], "params": params, "t0": t0, "tf": tf, "y0": y0,
                          "expected": expected, "solver": SOLVERS[sname], "h": float(m.group(2)) if m.group(2) else None,
                          "rtol": float(opts["rtol"]) if "rtol" in opts else None,
                          "atol": float(opts["atol"]) if "atol" in opts else None, "tolerance": float(tol),
                          "source": "tests/ode/accuracy.rs"})
    out = {"_generated_by": "tools/extract_reference_goldens.py from /root/reference/tests/ode/accuracy.rs "
                            "(SciPy DOP853 rtol=atol=1e-12 reference values, the reference's own tolerances)",
           "accuracy": cases}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(f"{len(cases)} cases -> {OUT}")


if __name__ == "__main__":
    main()
