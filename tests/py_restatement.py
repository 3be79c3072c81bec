"""A second, independent restatement of the reference's Dormand-Prince / fixed-step / Euler-Maruyama loops in
plain Python floats (IEEE doubles, math.pow == libm pow), written from the Rust sources, NOT from oracle/oracle.cpp.

TEST INFRASTRUCTURE.  Pure-Python loops: only for small cases.  Its purpose is to pin the C++ oracle: two
independently written restatements must agree bit for bit (tests/test_oracle_golden.py), and both must reproduce the
survey's probe values.

Follows /root/reference/src/methods/erk/dormandprince/ordinary.rs:16-337, src/methods/h_init.rs:45-135,
src/ode/solve_ivp.rs:139-277, src/solout/t_eval.rs:87-171, src/methods/erk/fixed/ordinary.rs:16-221,
src/interpolate.rs:40-74.
"""
import math
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_header_tableaux():
    """The coefficient header the kernels and the C++ oracle are built from (tools/gen_tableau.py): only used to cross-check
    the two extractions against each other (tests/test_host_logic.py), NOT by the restatement itself."""
    txt = open(os.path.join(ROOT, "oracle", "erk_tableau_data.h")).read().replace("\\\n", " ")
    out = {}
    for m in re.finditer(r"#define DEB_(\w+?)_(C|A|B|BH|ER|BI) (.*)", txt):
        name, kind, body = m.group(1), m.group(2), m.group(3)
        rows = re.findall(r"\{([^{}]*)\}", body)
        vals = [[float.fromhex(x.strip()) for x in r.split(",")] for r in rows]
        out.setdefault(name, {})[kind] = vals[0] if kind in ("C", "B", "BH", "ER") else vals
    return out


def load_tableaux():
    """The restatement's own tableaux: read from the reference's Rust sources by tests/support/reference_tableaux.py, a reader
    that shares nothing with the generator of the header above, and committed as tests/golden/reference_tableaux.json."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
    import reference_tableaux
    return reference_tableaux.load_fixture()


TAB = load_tableaux()
EPS10 = 2.220446049250313e-16 * 10.0


def signum(x):
    return math.copysign(1.0, x)


def rmax(a, b):  # f64::max ignores NaN
    if a != a:
        return b
    if b != b:
        return a
    return max(a, b)


def rmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return min(a, b)


def powf(x, y):
    if x == 0.0:
        return math.inf if y < 0 else 0.0
    try:
        return math.pow(x, y)
    except (OverflowError, ZeroDivisionError):
        return math.inf
    except ValueError:
        return math.nan


def h_init(f, t0, tf, y0, order, rtol, atol, h_min, h_max):
    posneg = signum(tf - t0)
    n = len(y0)
    f0 = f(t0, y0)
    dnf = dny = 0.0
    sk = []
    for i in range(n):
        s = atol + rtol * abs(y0[i])
        sk.append(s)
        a = f0[i] / s
        dnf += a * a
        b = y0[i] / s
        dny += b * b
    h = 1.0e-6 if (dnf <= 1.0e-10 or dny <= 1.0e-10) else math.sqrt(dny / dnf) * 0.01
    h = rmin(h, h_max)
    h *= posneg
    y1 = [y0[i] + h * f0[i] for i in range(n)]
    f1 = f(t0 + h, y1)
    der2 = 0.0
    for i in range(n):
        d = (f1[i] - f0[i]) / sk[i]
        der2 += d * d
    der2 = math.sqrt(der2) / abs(h)
    der12 = rmax(math.sqrt(dnf), der2)
    if der12 <= 1.0e-15:
        h1 = abs(h) * rmax(1.0e-3, 1.0e-6)
    else:
        h1 = powf(0.01 / der12, 1.0 / float(order))
    interval = abs(tf - t0)
    h = rmin(rmax(rmin(rmin(abs(h) * 100.0, h1), h_max), h_min), interval)
    return h * posneg


def make_even_solout(dt, t0, tf, rows, interpolate):
    """EvenSolout::solout, /root/reference/src/solout/even.rs:69-199.  `interpolate(ti)` is the method's dense output."""
    dirn = signum(tf - t0)
    st = dict(last=None)
    tol = abs(dt) * 1e-12 + 2.220446049250313e-16 * 10.0

    def solout(t_curr, t_prev, y_curr, y_prev):
        offset = math.fmod(t0, dt)
        if st["last"] is not None:
            start_t = st["last"] + dt * dirn
        else:
            if abs(t_prev - t0) < 2.220446049250313e-16:
                rows.append((t0, list(y_prev)))
                st["last"] = t0
                start_t = t0 + dt * dirn
            else:
                rem = math.fmod(t_prev - offset, dt)
                if dirn > 0:
                    start_t = t_prev if abs(rem) < 2.220446049250313e-16 else t_prev + (dt - rem)
                else:
                    start_t = t_prev if abs(rem) < 2.220446049250313e-16 else t_prev - rem
        ti = start_t
        while (dirn > 0 and ti <= t_curr) or (dirn < 0 and ti >= t_curr):
            if (dirn > 0 and t_prev <= ti <= t_curr) or (dirn < 0 and t_prev >= ti >= t_curr):
                if st["last"] is not None and abs(ti - st["last"]) <= tol:
                    pass
                else:
                    rows.append((ti, interpolate(ti)))
                    st["last"] = ti
            ti += dt * dirn
        if t_curr == tf:
            if st["last"] is not None:
                if abs(st["last"] - tf) <= tol:
                    rows.pop()
                    rows.append((tf, list(y_curr)))
                    st["last"] = tf
                elif st["last"] != tf:
                    rows.append((tf, list(y_curr)))
                    st["last"] = tf
            else:
                rows.append((tf, list(y_curr)))
                st["last"] = tf
    return solout


def make_per_step_solout(recorder, rows, interpolate):
    """DefaultSolout (src/solout/default.rs:54-75), DenseSolout (dense.rs:74-108), CrossingSolout (crossing.rs:115-263).
    recorder = ("default",) | ("dense", n) | ("crossing", component, threshold, direction in {0, +1, -1})
             | ("hyperplane", point, normal, component indices, direction)  (src/solout/hyperplane.rs)."""
    kind = recorder[0]
    st = dict(last=None)
    eps = 2.220446049250313e-16

    def newton(comp, thr, t_lower, t_upper, off_lower, off_upper):
        t = t_lower - off_lower * (t_upper - t_lower) / (off_upper - off_lower)
        tolerance = eps * 100.0
        for _ in range(10):
            off = interpolate(t)[comp] - thr
            if abs(off) < tolerance:
                return t
            delta_t = (t_upper - t_lower) * 1e-6
            off_plus = interpolate(t + delta_t)[comp] - thr
            derivative = (off_plus - off) / delta_t
            if abs(derivative) < eps * 10.0:
                break
            t_new = t - off / derivative
            if t_new < t_lower or t_new > t_upper:
                t = (t_lower + t_upper) / 2.0
            else:
                change = abs(t_new - t)
                t = t_new
                if change < tolerance * 0.1:
                    break
        off = interpolate(t)[comp] - thr
        return t if abs(off) < tolerance * 10.0 else None

    if kind == "hyperplane":  # HyperplaneCrossingSolout::new, src/solout/hyperplane.rs:124-139
        _, pl_point, pl_normal, pl_comps, pl_dir = recorder
        nsq = 0.0
        for v in pl_normal:
            nsq = nsq + v * v
        norm = math.sqrt(nsq)
        pl_n = [v * (1.0 / norm) for v in pl_normal] if norm > eps else list(pl_normal)

        def distance(y):
            sm = 0.0
            for i, cidx in enumerate(pl_comps):
                sm = sm + (y[cidx] + (-1.0) * pl_point[i]) * pl_n[i]
            return sm

        def plane_newton(t_lower, t_upper, d_lower, d_upper):
            t = t_lower - d_lower * (t_upper - t_lower) / (d_upper - d_lower)
            tolerance = eps * 100.0
            for _ in range(10):
                dist = distance(interpolate(t))
                if abs(dist) < tolerance:
                    return t
                delta_t = (t_upper - t_lower) * 1e-6
                derivative = (distance(interpolate(t + delta_t)) - dist) / delta_t
                if abs(derivative) < eps:
                    break
                t_new = t - dist / derivative
                t = (t_lower + t_upper) / 2.0 if (t_new < t_lower or t_new > t_upper) else t_new
            return t if abs(distance(interpolate(t))) < tolerance * 10.0 else None

    def solout(t_curr, t_prev, y_curr):
        if kind == "hyperplane":  # hyperplane.rs:182-236
            dist = distance(y_curr)
            last = st["last"]
            if last is not None and (signum(last) != signum(dist) or (last == 0.0 and dist != 0.0) or (last != 0.0 and dist == 0.0)):
                record = (last < 0.0 and dist >= 0.0) if pl_dir > 0 else (last > 0.0 and dist <= 0.0) if pl_dir < 0 else True
                if record:
                    tc = plane_newton(t_prev, t_curr, last, dist)
                    if tc is None:
                        tc = t_prev + (-last / (dist - last)) * (t_curr - t_prev)
                    rows.append((tc, interpolate(tc)))
            st["last"] = dist
        elif kind == "default":
            rows.append((t_curr, list(y_curr)))
        elif kind == "dense":
            n = recorder[1]
            if t_prev != t_curr:
                for i in range(1, n):
                    h_old = t_curr - t_prev
                    ti = t_prev + float(i) * h_old / float(n)
                    rows.append((ti, interpolate(ti)))
            rows.append((t_curr, list(y_curr)))
        else:
            _, comp, thr, direction = recorder
            off = y_curr[comp] - thr
            last = st["last"]
            if last is not None and signum(last) != signum(off):
                record = (last < 0.0 and off >= 0.0) if direction > 0 else (last > 0.0 and off <= 0.0) if direction < 0 else True
                if record:
                    tc = newton(comp, thr, t_prev, t_curr, last, off)
                    if tc is None:
                        frac = -last / (off - last)
                        tc = t_prev + frac * (t_curr - t_prev)
                    rows.append((tc, interpolate(tc)))
            st["last"] = off
    return solout


def make_event_solout(base, event, rows, interpolate, dirn, bounds):
    """EventWrappedSolout (src/solout/event.rs:300-470).  event = (g(t, y), direction in {0, +1, -1}, terminate count or 0);
    bounds() -> (t_prev, t_curr) of the step for the reference's interpolate() bounds test.  Returns True to terminate."""
    g, direction, terminate = event
    st = dict(last=None, count=0)
    rel_tol, abs_tol = 1e-12, 1e-14

    def brent_dekker(a, b, fa, fb):
        if abs(fa) < abs(fb):
            a, b, fa, fb = b, a, fb, fa
        c, fc = a, fa
        d = b - a
        e = d
        for _ in range(50):
            if fb == 0.0:
                return b
            if signum(fa) == signum(fb):
                a, fa = c, fc
                c, fc = b, fb
                d = b - a
                e = d
            if abs(fa) < abs(fb):
                c = b
                b = a
                a = c
                fc = fb
                fb = fa
                fa = fc
            tol = rmax(abs_tol, rel_tol * abs(b))
            m = 0.5 * (a - b)
            if abs(m) <= tol or fb == 0.0:
                return b
            use_bis = True
            if abs(e) > tol and abs(fa) > abs(fb):
                s_ = fb / fa
                if a == c:
                    p_ = 2.0 * m * s_
                    q_ = 1.0 - s_
                else:
                    q1 = fa / fc
                    r = fb / fc
                    p_ = s_ * (2.0 * m * q1 * (q1 - r) - (b - a) * (r - 1.0))
                    q_ = (q1 - 1.0) * (r - 1.0) * (s_ - 1.0)
                if q_ > 0.0:
                    p_ = -p_
                else:
                    q_ = -q_
                if abs(2.0 * p_) < (3.0 * m * q_ - abs(tol * q_)) and p_ < abs(e * 0.5 * q_):
                    e = d
                    d = p_ / q_
                    use_bis = False
            if use_bis:
                d = m
                e = m
            a, fa = b, fb
            b = b + d if abs(d) > tol else b + (tol if m > 0.0 else -tol)
            tp, tc = bounds()
            if b < tp or b > tc:  # interpolate(b) -> Err(OutOfBounds) -> .ok()? -> None
                return None
            fb = g(b, interpolate(b))
            c, fc = a, fa
        return None

    def solout(t_curr, t_prev, y_curr):
        base(t_curr, t_prev, y_curr)
        g_curr = g(t_curr, y_curr)
        if st["last"] is None:
            st["last"] = g_curr
            return False
        g_prev = st["last"]
        sign_change = signum(g_prev) != signum(g_curr)
        if direction > 0:
            ok = sign_change and g_prev < 0.0 and g_curr >= 0.0
        elif direction < 0:
            ok = sign_change and g_prev > 0.0 and g_curr <= 0.0
        else:
            ok = sign_change
        if ok:
            a, b, fa, fb = t_prev, t_curr, g_prev, g_curr
            if (dirn > 0 and a > b) or (dirn < 0 and a < b):
                a, b, fa, fb = b, a, fb, fa
            if fa * fb <= 0.0:
                te = brent_dekker(a, b, fa, fb)
                if te is not None:
                    ye = interpolate(te)
                    if not rows or abs(te - rows[-1][0]) > abs_tol:
                        rows.append((te, ye))
                    st["count"] += 1
                    if terminate and st["count"] >= terminate:
                        st["last"] = g_curr
                        return True
        st["last"] = g_curr
        return False
    return solout


def solve_dp(f, method, t0, tf, y0, rtol=1e-6, atol=1e-6, t_eval=(), even=None, recorder=None, event=None, h0=0.0, h_min=0.0, h_max=math.inf, max_steps=10000,
             safety=0.9, min_scale=0.2, max_scale=10.0):
    """Returns dict(status, t, y, accepted, rejected, evals, rows=[(t, y)])."""
    T = TAB["DOPRI5" if method == "dopri5" else "DOP853"]
    O, S, I = (5, 7, 7) if method == "dopri5" else (8, 12, 16)
    c, A, b, er, bi = T["C"], T["A"], T["B"], T["ER"], T["BI"]
    bh = T.get("BH")
    n = len(y0)
    evals = 0
    if tf == t0:
        return dict(status="BadInput")
    dirn = signum(tf - t0)
    if h0 == 0.0:
        h0 = h_init(f, t0, tf, y0, O, rtol, atol, h_min, h_max)
        evals += 2
    if (signum(h0) != dirn or h_min < 0 or h_max < 0 or h_min > h_max or abs(h0) < h_min or abs(h0) > h_max
            or abs(h0) > abs(tf - t0) or h0 == 0.0):
        return dict(status="BadInput")
    h, t, y = h0, t0, list(y0)
    k = [[0.0] * n for _ in range(I)]
    cont = [[0.0] * n for _ in range(O)]
    k[0] = f(t, y)
    evals += 1
    t_prev, h_prev = t, 0.0
    steps = stiff = nonstiff = acc = rej = 0
    rejected = False
    pts = sorted(t_eval) if dirn > 0 else sorted(t_eval, reverse=True)
    rows = []
    state = dict(idx=0)

    def interpolate(ti):
        s = (ti - t_prev) / h_prev
        s1 = 1.0 - s
        ilast = O - 1
        accv = list(cont[ilast])
        for i in range(ilast - 1, 0, -1):
            if i >= 4:
                factor = s1 if (ilast - i) % 2 == 1 else s
            else:
                factor = s1 if i % 2 == 1 else s
            accv = [v * factor for v in accv]
            accv = [accv[j] + 1.0 * cont[i][j] for j in range(n)]
        return [cont[0][j] + s * accv[j] for j in range(n)]

    def solout(t_curr, tp, y_curr):
        idx = state["idx"]
        while idx < len(pts):
            te = pts[idx]
            if dirn > 0:
                in_range = (te == tp and idx == 0) or (te > tp and te <= t_curr)
            else:
                in_range = (te == tp and idx == 0) or (te < tp and te >= t_curr)
            if in_range:
                rows.append((te, list(y_curr) if te == t_curr else interpolate(te)))
                idx += 1
            else:
                if (dirn > 0 and te > t_curr) or (dirn < 0 and te < t_curr):
                    break
                idx += 1
        state["idx"] = idx

    y_before = list(y)
    if even is not None:
        even_solout = make_even_solout(even, t0, tf, rows, interpolate)
        solout = lambda tc, tp, yc: even_solout(tc, tp, yc, y_before)
    if recorder is not None:
        solout = make_per_step_solout(recorder, rows, interpolate)
    if event is not None:
        solout = make_event_solout(solout, event, rows, interpolate, dirn, lambda: (t_prev, t))
    solout(t, t_prev, y)
    status = "Complete"
    while True:
        if (t + h - tf) * dirn > 0.0:
            h_new = tf - t
            if abs(h_new) < EPS10:
                break
            h = h_new
        if abs(h) < abs(h_prev) * 1e-14:
            status = "StepSize"
            break
        if steps >= max_steps:
            status = "MaxSteps"
            break
        steps += 1
        ys = [0.0] * n
        for i in range(1, S):
            ys = list(y)
            for j in range(i):
                ah = A[i][j] * h
                ys = [ys[q] + ah * k[j][q] for q in range(n)]
            k[i] = f(t + c[i] * h, ys)
        ysti = list(ys)
        yseg = [0.0] * n
        for i in range(S):
            yseg = [yseg[q] + b[i] * k[i][q] for q in range(n)]
        y_new = [y[q] + h * yseg[q] for q in range(n)]
        t_new = t + h
        step_evals = S - 1
        es = [0.0] * n
        for j in range(S):
            es = [es[q] + er[j] * k[j][q] for q in range(n)]

        def enorm(e):
            tot = 0.0
            for q in range(n):
                sk = atol + rtol * rmax(abs(y[q]), abs(y_new[q]))
                v = e[q] / sk
                tot += v * v
            return tot
        err = enorm(es)
        err2 = 0.0
        if bh is not None:
            e2 = list(yseg)
            for j in range(S):
                e2 = [e2[q] + (-bh[j]) * k[j][q] for q in range(n)]
            err2 = enorm(e2)
        deno = err + 0.01 * err2
        if deno <= 0.0:
            deno = 1.0
        err = abs(h) * err * math.sqrt(1.0 / (deno * float(n)))
        scale = safety * powf(err, -(1.0 / float(O)))
        scale = rmin(rmax(scale, min_scale), max_scale)
        if err <= 1.0:
            dydt = f(t_new, y_new)
            step_evals += 1
            if steps % 100 == 0:
                stdnum = 0.0
                stden = 0.0
                for q in range(n):
                    d = yseg[q] - k[S - 1][q]
                    stdnum += d * d
                for q in range(n):
                    d = dydt[q] - ysti[q]
                    stden += d * d
                if stden > 0.0:
                    if h * math.sqrt(stdnum / stden) > 6.1:
                        nonstiff = 0
                        stiff += 1
                        if stiff == 15:
                            status = "Stiffness"
                            break
                else:
                    nonstiff += 1
                    if nonstiff == 6:
                        stiff = 0
            cont[0] = list(y)
            ydiff = [y_new[q] + (-1.0) * y[q] for q in range(n)]
            cont[1] = list(ydiff)
            bspl = [0.0 + h * k[0][q] for q in range(n)]
            bspl = [bspl[q] + (-1.0) * ydiff[q] for q in range(n)]
            cont[2] = list(bspl)
            c3 = [ydiff[q] + (-h) * dydt[q] for q in range(n)]
            c3 = [c3[q] + (-1.0) * bspl[q] for q in range(n)]
            cont[3] = c3
            if I > S:
                k[S] = list(dydt)
                for i in range(S + 1, I):
                    ys2 = list(y)
                    for j in range(i):
                        ah = A[i][j] * h
                        ys2 = [ys2[q] + ah * k[j][q] for q in range(n)]
                    k[i] = f(t + c[i] * h, ys2)
                    step_evals += 1
            for i in range(4, O):
                ci = [0.0] * n
                for j in range(I):
                    ci = [ci[q] + bi[i][j] * k[j][q] for q in range(n)]
                cont[i] = [v * h for v in ci]
            t_prev, h_prev = t, h
            t, y = t_new, y_new
            k[0] = list(dydt)
            if rejected:
                rejected = False
                scale = rmin(scale, 1.0)
            accepted = True
        else:
            rejected = True
            accepted = False
        h *= scale
        sg = signum(h)
        if abs(h) < h_min:
            h = sg * h_min
        elif abs(h) > h_max:
            h = sg * h_max
        evals += step_evals
        if not accepted:
            rej += 1
            continue
        acc += 1
        if solout(t, t_prev, y):  # ControlFlag::Terminate (an event), solve_ivp.rs:255-260
            status = "Interrupted"
            break
        if abs(tf - t) <= EPS10:
            break
    return dict(status=status, t=t, y=y, accepted=acc, rejected=rej, evals=evals, rows=rows)


def solve_fixed(f, method, h0, t0, tf, y0, t_eval=(), max_steps=10000, recorder=None, event=None):
    T = TAB[method.upper()]
    c, A, b = T["C"], T["A"], T["B"]
    S = len(b)
    n = len(y0)
    if tf == t0:
        return dict(status="BadInput")
    dirn = signum(tf - t0)
    if h0 == 0.0:
        h0 = abs(tf - t0) / 100.0
    if signum(h0) != dirn or abs(h0) > abs(tf - t0):
        return dict(status="BadInput")
    h, t, y = h0, t0, list(y0)
    dydt = f(t, y)
    evals = 1
    t_prev, y_prev, d_prev = t, list(y), list(dydt)
    pts = sorted(t_eval) if dirn > 0 else sorted(t_eval, reverse=True)
    rows = []
    state = dict(idx=0)

    def interpolate(ti):
        hh = t - t_prev
        s = (ti - t_prev) / hh
        s2 = s * s
        s3 = s2 * s
        h00 = 2.0 * s3 - 3.0 * s2 + 1.0
        h10 = s3 - 2.0 * s2 + s
        h01 = -2.0 * s3 + 3.0 * s2
        h11 = s3 - s2
        out = [0.0] * n
        out = [out[q] + h00 * y_prev[q] for q in range(n)]
        out = [out[q] + (h10 * hh) * d_prev[q] for q in range(n)]
        out = [out[q] + h01 * y[q] for q in range(n)]
        out = [out[q] + (h11 * hh) * dydt[q] for q in range(n)]
        return out

    def solout(t_curr, tp, y_curr):
        idx = state["idx"]
        while idx < len(pts):
            te = pts[idx]
            if dirn > 0:
                in_range = (te == tp and idx == 0) or (te > tp and te <= t_curr)
            else:
                in_range = (te == tp and idx == 0) or (te < tp and te >= t_curr)
            if in_range:
                rows.append((te, list(y_curr) if te == t_curr else interpolate(te)))
                idx += 1
            else:
                if (dirn > 0 and te > t_curr) or (dirn < 0 and te < t_curr):
                    break
                idx += 1
        state["idx"] = idx

    if recorder is not None:
        solout = make_per_step_solout(recorder, rows, interpolate)
    if event is not None:
        solout = make_event_solout(solout, event, rows, interpolate, dirn, lambda: (t_prev, t))
    solout(t, t_prev, y)
    steps = 0
    status = "Complete"
    while True:
        if (t + h - tf) * dirn > 0.0:
            h_new = tf - t
            if abs(h_new) < EPS10:
                break
            h = h_new
        if steps >= max_steps:
            status = "MaxSteps"
            break
        steps += 1
        k = [list(dydt)] + [None] * (S - 1)
        for i in range(1, S):
            ys = list(y)
            for j in range(i):
                ah = A[i][j] * h
                ys = [ys[q] + ah * k[j][q] for q in range(n)]
            k[i] = f(t + c[i] * h, ys)
        t_prev, y_prev, d_prev = t, list(y), list(k[0])
        yn = list(y)
        for i in range(S):
            bh_ = b[i] * h
            yn = [yn[q] + bh_ * k[i][q] for q in range(n)]
        t += h
        y = yn
        dydt = f(t, y)
        evals += S
        if solout(t, t_prev, y):  # ControlFlag::Terminate (an event), solve_ivp.rs:255-260
            status = "Interrupted"
            break
        if abs(tf - t) <= EPS10:
            break
    return dict(status=status, t=t, y=y, accepted=steps, rejected=0, evals=evals, rows=rows)


# right-hand sides, written from /root/reference/tests/ode/systems.rs
def lorenz(sigma, rho, beta):
    def f(t, y):
        x, yv, z = y
        return [sigma * (yv - x), x * (rho - z) - yv, x * yv - beta * z]
    return f


def van_der_pol(mu):
    def f(t, y):
        y1, y2 = y
        return [y2, mu * (1.0 - y1 * y1) * y2 - y1]
    return f


def exponential(k):
    return lambda t, y: [k * y[0]]


def harmonic(k):
    return lambda t, y: [y[1], -k * y[0]]


# (order, fsal): adaptive/mod.rs:43-122
ADAPTIVE_ORDER_FSAL = dict(rkf45=(5, False), cash_karp=(5, False), rkv655e=(6, True), rkv656e=(6, True), rkv766e=(7, False),
                           rkv767e=(7, False), rkv877e=(8, False), rkv878e=(8, False), rkv988e=(9, False), rkv989e=(9, False))


def solve_adaptive(f, method, t0, tf, y0, rtol=1e-6, atol=1e-6, t_eval=(), recorder=None, event=None, h0=0.0, h_min=0.0, h_max=math.inf, max_steps=10000,
                   max_rejects=100, safety=0.9, min_scale=0.2, max_scale=10.0):
    """Generic adaptive family (RKF45, Cash-Karp, Verner pairs): /root/reference/src/methods/erk/adaptive/ordinary.rs:16-211,
    dense output = the method's polynomial when it has one (:246-277) else cubic Hermite (:282-295), driven by
    src/ode/solve_ivp.rs:139-277."""
    T = TAB[method.upper()]
    c, A, b, bh = T["C"], T["A"], T["B"], T["BH"]
    BI = T.get("BI")
    S, I = len(b), len(c)
    O, fsal = ADAPTIVE_ORDER_FSAL[method.lower()]
    n = len(y0)
    evals = 0
    if tf == t0:
        return dict(status="BadInput")
    dirn = signum(tf - t0)
    if h0 == 0.0:
        h0 = h_init(f, t0, tf, y0, O, rtol, atol, h_min, h_max)
        evals += 4  # (sic) 2 inside compute() + 2 again in init()
    if (signum(h0) != dirn or h_min < 0 or h_max < 0 or h_min > h_max or abs(h0) < h_min or abs(h0) > h_max
            or abs(h0) > abs(tf - t0) or h0 == 0.0):
        return dict(status="BadInput")
    h, t, y = h0, t0, list(y0)
    dydt = f(t, y)
    evals += 1
    t_prev, y_prev, d_prev, h_prev = t, list(y), list(dydt), 0.0
    steps = stiff = acc = rej = 0
    rejected = False
    pts = sorted(t_eval) if dirn > 0 else sorted(t_eval, reverse=True)
    rows = []
    state = dict(idx=0)

    kk = [None]  # the stage vectors of the last step (self.k)

    def interpolate(ti):
        if BI is not None:  # :246-277
            s = (ti - t_prev) / h_prev
            out = list(y_prev)
            cont = []
            for i in range(I):
                ci = BI[i][O - 1]
                for j in range(O - 2, -1, -1):
                    ci = ci * s + BI[i][j]
                cont.append(ci * s)
            for i in range(I):
                w = cont[i] * h_prev
                out = [out[q] + w * kk[0][i][q] for q in range(n)]
            return out
        hh = t - t_prev
        s = (ti - t_prev) / hh
        s2 = s * s
        s3 = s2 * s
        h00 = 2.0 * s3 - 3.0 * s2 + 1.0
        h10 = s3 - 2.0 * s2 + s
        h01 = -2.0 * s3 + 3.0 * s2
        h11 = s3 - s2
        out = [0.0] * n
        out = [out[q] + h00 * y_prev[q] for q in range(n)]
        out = [out[q] + (h10 * hh) * d_prev[q] for q in range(n)]
        out = [out[q] + h01 * y[q] for q in range(n)]
        out = [out[q] + (h11 * hh) * dydt[q] for q in range(n)]
        return out

    def solout(t_curr, tp, y_curr):
        idx = state["idx"]
        while idx < len(pts):
            te = pts[idx]
            if dirn > 0:
                in_range = (te == tp and idx == 0) or (te > tp and te <= t_curr)
            else:
                in_range = (te == tp and idx == 0) or (te < tp and te >= t_curr)
            if in_range:
                rows.append((te, list(y_curr) if te == t_curr else interpolate(te)))
                idx += 1
            else:
                if (dirn > 0 and te > t_curr) or (dirn < 0 and te < t_curr):
                    break
                idx += 1
        state["idx"] = idx

    if recorder is not None:
        solout = make_per_step_solout(recorder, rows, interpolate)
    if event is not None:
        solout = make_event_solout(solout, event, rows, interpolate, dirn, lambda: (t_prev, t))
    solout(t, t_prev, y)
    status = "Complete"
    while True:
        if (t + h - tf) * dirn > 0.0:
            h_new = tf - t
            if abs(h_new) < EPS10:
                break
            h = h_new
        if abs(h) < abs(h_prev) * 1e-14:
            status = "StepSize"
            break
        if steps >= max_steps:
            status = "MaxSteps"
            break
        steps += 1
        k = [list(dydt)] + [None] * (S - 1)
        for i in range(1, S):
            ys = list(y)
            for j in range(i):
                ah = A[i][j] * h
                ys = [ys[q] + ah * k[j][q] for q in range(n)]
            k[i] = f(t + c[i] * h, ys)
        step_evals = S - 1
        y_high = list(y)
        for i in range(S):
            w = b[i] * h
            y_high = [y_high[q] + w * k[i][q] for q in range(n)]
        y_low = list(y)
        for i in range(S):
            w = bh[i] * h
            y_low = [y_low[q] + w * k[i][q] for q in range(n)]
        err = [y_high[q] + (-1.0) * y_low[q] for q in range(n)]
        err_norm = 0.0
        for q in range(n):
            sk = atol + rtol * rmax(abs(y[q]), abs(y_high[q]))
            err_norm = rmax(err_norm, abs(err[q] / sk))
        scale = safety * powf(err_norm, -(1.0 / float(O)))
        scale = rmin(rmax(scale, min_scale), max_scale)
        if err_norm <= 1.0:
            t_prev, y_prev, d_prev, h_prev = t, list(y), list(k[0]), h
            if rejected:
                stiff = 0
                rejected = False
                scale = rmin(scale, 1.0)
            if BI is not None:  # extra stages for the dense-output polynomial, :145-160
                for i in range(S, I):
                    ys = list(y)
                    for j in range(i):
                        ah = A[i][j] * h
                        ys = [ys[q] + ah * k[j][q] for q in range(n)]
                    k.append(f(t + c[i] * h, ys))
                step_evals += I - S
                kk[0] = k
            t += h
            y = y_high
            if fsal:
                dydt = list(k[S - 1])
            else:
                dydt = f(t, y)
                step_evals += 1
            accepted = True
        else:
            rejected = True
            stiff += 1
            if stiff >= max_rejects:
                status = "Stiffness"
                break
            accepted = False
        h *= scale
        sg = signum(h)
        if abs(h) < h_min:
            h = sg * h_min
        elif abs(h) > h_max:
            h = sg * h_max
        evals += step_evals
        if not accepted:
            rej += 1
            continue
        acc += 1
        if solout(t, t_prev, y):  # ControlFlag::Terminate (an event), solve_ivp.rs:255-260
            status = "Interrupted"
            break
        if abs(tf - t) <= EPS10:
            break
    return dict(status=status, t=t, y=y, accepted=acc, rejected=rej, evals=evals, rows=rows)
