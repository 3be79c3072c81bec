// step_recorder.cuh -- the per-step output recorders of the reference, on the device:
//     DefaultSolout    /root/reference/src/solout/default.rs:54-75     every accepted step (t, y)
//     DenseSolout      /root/reference/src/solout/dense.rs:74-108      n-1 interpolated points per step + the step end
//     CrossingSolout   /root/reference/src/solout/crossing.rs:115-263  component crossing a threshold, refined by the
//                                                                      reference's Newton iteration on the dense output
//     HyperplaneCrossingSolout  src/solout/hyperplane.rs:170-330       selected components crossing a hyperplane
// Unlike t_eval / even(dt), the row times depend on the trajectory, so rows carry their own time (t_out) and the
// number of rows is not known in advance: each trajectory owns `row_stride` row slots; pushes beyond that are counted
// (n_emitted) but not stored.
//
// DenseStep is the dense output of one accepted step, as the reference's `Interpolation::interpolate` computes it:
//     Dormand-Prince family   dormandprince/ordinary.rs:196-234 (cont) + :301-337 (nested polynomial)
//     Verner pairs            adaptive/ordinary.rs:145-160 (extra stages) + :246-277 (Horner in s)
//     everything else         cubic Hermite, src/interpolate.rs:40-60
// The reference's bounds check (`t < t_prev || t > t_curr` => Err(OutOfBounds), unwrapped by the recorders) is written
// for forward time; where it would panic (backward time, or the Newton probe t + 1e-6*(step) just past the step end)
// the polynomial is simply evaluated.
// (included by erk_ensemble.cuh after OdeKernelArgs and d_signum)
#pragma once

namespace deb {

enum { REC_T_EVAL = 0, REC_EVEN = 1, REC_DEFAULT = 2, REC_DENSE = 3, REC_CROSSING = 4, REC_HYPERPLANE = 5 };  // = deb_solout values

// The recorder kernels are compiled at first use, one per recorder: the translation unit fixes the mode (DEB_JIT_REC_MODE), so
// the code of the other recorders -- dense-output preparation, Newton, the row plan -- is not part of the kernel at all.
#ifdef DEB_JIT_REC_MODE
__device__ __forceinline__ constexpr int rec_mode_of(const OdeKernelArgs&) { return DEB_JIT_REC_MODE; }
// rows are staged in shared memory (StepRecorder::push) by the recorders that emit in every step; for the others -- a row
// every few dozen steps -- the staging code only costs registers (measured: t_eval + event 19.5 -> 22.6 ms with it)
constexpr bool REC_ROWS_STAGED = (DEB_JIT_REC_MODE == 2 /* REC_DEFAULT */ || DEB_JIT_REC_MODE == 3 /* REC_DENSE */);
#else
__device__ __forceinline__ int rec_mode_of(const OdeKernelArgs& a) { return a.rec_mode; }
constexpr bool REC_ROWS_STAGED = true;
#endif

// Event functions g(t, y) (the `Event` trait, src/solout/event.rs:60-70).  EvtNone: no event detection.
struct EvtNone {
    static constexpr bool ENABLED = false;
    __device__ __forceinline__ static double g(const OdeKernelArgs&, double, const double*, const double*) { return 0.0; }
};
// g(t, y) = c0 + c1*t + sum_c c[2+c]*y[c], accumulated in that order (e.g. `y[0] - 0.9*m` is {-0.9*m, 0, 1}: same bits)
template <int N>
struct EvtLinear {
    static constexpr bool ENABLED = true;
    __device__ __forceinline__ static double g(const OdeKernelArgs& a, double t, const double* y, const double*) {
        double v = a.event_coef[0] + a.event_coef[1] * t;
#pragma unroll
        for (int c = 0; c < N; c++) v = v + a.event_coef[2 + c] * y[c];
        return v;
    }
};

template <class Sys, class Tab>
struct DenseStep {
    static constexpr int N = Sys::DIM, S = Tab::S, I = Tab::I, O = Tab::O;
    static constexpr bool DP = Tab::ADAPTIVE && Tab::DP, BI = Tab::ADAPTIVE && Tab::BI_POLY;
    static constexpr int NK = BI ? I : 1;
    double t, h, t_new;
    double y[N], yn[N];
    double c1[N], c2[N], c3[N], ch[(O > 4) ? (O - 4) : 1][N];  // DP family
    double kk[NK][N];                                          // Verner pairs: all I stage vectors
    double d0[N], d1[N];                                       // Hermite: derivatives at both ends

    // state before the step (t, y, k[0] = f(t,y)), the step (h, k[1..S-1], y_new) and dydt = f(t+h, y_new)
    __device__ __forceinline__ void prepare(double t_, double h_, const double (&y_)[N], const double (&yn_)[N],
                                            const double (&k)[S][N], const double (&dydt)[N], const double* p) {
        t = t_; h = h_; t_new = t_ + h_;
#pragma unroll
        for (int c = 0; c < N; c++) { y[c] = y_[c]; yn[c] = yn_[c]; d0[c] = k[0][c]; d1[c] = dydt[c]; }
        if constexpr (DP) {
#pragma unroll
            for (int c = 0; c < N; c++) {
                c1[c] = yn[c] - y[c];
                c2[c] = __dadd_rn(0.0, h * k[0][c]) - c1[c];
                c3[c] = (c1[c] + (-h) * dydt[c]) - c2[c];
            }
            double kx[(I > S) ? (I - S) : 1][N];
#pragma unroll
            for (int c = 0; c < N; c++) kx[0][c] = dydt[c];
#pragma unroll
            for (int i = S + 1; i < I; i++) {
                double ys[N];
#pragma unroll
                for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
                for (int j = 0; j < i; j++) {
                    if (Tab::a(i, j) != 0.0) {
                        const double ah = Tab::av(i, j) * h;
#pragma unroll
                        for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * ((j < S) ? k[(j < S) ? j : 0][c] : kx[(j >= S) ? (j - S) : 0][c]);
                    }
                }
                Sys::rhs(t + Tab::cv(i) * h, ys, kx[i - S], p);
            }
#pragma unroll
            for (int i = 4; i < O; i++) {
#pragma unroll
                for (int c = 0; c < N; c++) ch[i - 4][c] = 0.0;
#pragma unroll
                for (int j = 0; j < I; j++) {
                    if (Tab::bi(i, j) != 0.0) {
#pragma unroll
                        for (int c = 0; c < N; c++)
                            ch[i - 4][c] = __dadd_rn(ch[i - 4][c], Tab::biv(i, j) * ((j < S) ? k[(j < S) ? j : 0][c] : kx[(j >= S) ? (j - S) : 0][c]));
                    }
                }
#pragma unroll
                for (int c = 0; c < N; c++) ch[i - 4][c] = ch[i - 4][c] * h;
            }
        } else if constexpr (BI) {
#pragma unroll
            for (int i = 0; i < S; i++) {
#pragma unroll
                for (int c = 0; c < N; c++) kk[i][c] = k[i][c];
            }
#pragma unroll
            for (int i = S; i < I; i++) {
                double ys[N];
#pragma unroll
                for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
                for (int j = 0; j < i; j++) {
                    if (Tab::a(i, j) != 0.0) {
                        const double ah = Tab::av(i, j) * h;
#pragma unroll
                        for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * kk[j][c];
                    }
                }
                Sys::rhs(t + Tab::cv(i) * h, ys, kk[i], p);
            }
        }
    }

    __device__ __forceinline__ void eval(double te, double (&row)[N]) const {
        if constexpr (DP) {
            const double sx = (te - t) / h;
            const double s1 = 1.0 - sx;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double accp = (O > 4) ? ch[(O > 4) ? (O - 5) : 0][c] : c3[c];
#pragma unroll
                for (int i = O - 2; i >= 1; i--) {
                    double factor;
                    if (i >= 4) factor = (((O - 1) - i) % 2 == 1) ? s1 : sx;
                    else factor = (i % 2 == 1) ? s1 : sx;
                    const double ci = (i >= 4) ? ch[(i >= 4) ? (i - 4) : 0][c] : (i == 3 ? c3[c] : (i == 2 ? c2[c] : c1[c]));
                    accp = accp * factor + ci;
                }
                row[c] = y[c] + sx * accp;
            }
        } else if constexpr (BI) {
            const double sx = (te - t) / h;
#pragma unroll
            for (int c = 0; c < N; c++) row[c] = y[c];
#pragma unroll
            for (int i = 0; i < I; i++) {
                if (Tab::bi_row(i)) {
                    double ci = Tab::biv(i, O - 1);
#pragma unroll
                    for (int j = O - 2; j >= 0; j--) ci = ci * sx + Tab::biv(i, j);
                    ci = ci * sx;
                    const double w = ci * h;
#pragma unroll
                    for (int c = 0; c < N; c++) row[c] = row[c] + w * kk[i][c];
                }
            }
        } else {
            const double hh = t_new - t;
            const double sx = (te - t) / hh;
            const double s2 = sx * sx, s3 = s2 * sx;
            const double h00 = 2.0 * s3 - 3.0 * s2 + 1.0;
            const double h10 = s3 - 2.0 * s2 + sx;
            const double h01 = -2.0 * s3 + 3.0 * s2;
            const double h11 = s3 - s2;
            const double w10 = h10 * hh, w11 = h11 * hh;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double v = __dadd_rn(0.0, h00 * y[c]);
                v = v + w10 * d0[c];
                v = v + h01 * yn[c];
                v = v + w11 * d1[c];
                row[c] = v;
            }
        }
    }
};

// Base recorder + optional event detection (EventWrappedSolout, src/solout/event.rs:300-470): the base recorder pushes its
// rows first, then a sign change of g over the step is located with Brent-Dekker on the dense output and pushed; after
// `event_terminate` events the integration stops with Status::Interrupted.
template <class Sys, class Tab, class Evt = EvtNone, int BLOCK = 128>
struct StepRecorder {
    static constexpr int N = Sys::DIM, S = Tab::S;
    // Rows are staged per lane in warp-private shared memory (RowStage, erk_ensemble.cuh) and leave as whole 32-byte
    // sectors: four (t) and 4/gcd(N,4) (y) rows per group.  A lane appends one row at a time to its own trajectory, so
    // without staging every 8-byte store of a warp is a separate partial-sector write at L2.
    using RowsY = RowStage<N, BLOCK, REC_ROWS_STAGED>;
    using RowsT = RowStage<1, BLOCK, REC_ROWS_STAGED>;
    static constexpr int STAGE_SLOTS = RowsY::SLOTS + RowsT::SLOTS;
    int rows = 0;             // pushes so far (Solution.t.len())
    double last_t = 0.0;      // time of the last pushed row (solution.t.last())
    bool have_last = false;   // CrossingSolout::last_offset_value
    double last_off = 0.0;
    int idx = 0;              // t_eval / even bases: next entry of the row plan
    double last_g = 0.0;      // EventWrappedSolout::last_g (always Some after the first call)
    int event_count = 0;

    __device__ __forceinline__ void reset() { rows = 0; have_last = false; last_off = 0.0; idx = 0; last_g = 0.0; event_count = 0; last_t = 0.0; }

    // `rows` is both the row being written and the number of rows before it, also when EvenSolout has just popped its
    // last point (rows -= 1; push): the staged group still holds the rows around the popped one and is written again.
    __device__ __forceinline__ void push(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, double t, const double (&y)[N]) {
        if (rows < a.row_stride) {
            if (a.y_eval) RowsY::put(RowSink{a.y_eval, a.row_stride, a.rows_vec}, buf, lane, traj, rows, rows, y);
            if (a.t_out) {
                const double tr[1] = {t};
                RowsT::put(RowSink{a.t_out, a.row_stride, a.tout_vec}, buf + RowsY::SLOTS, lane, traj, rows, rows, tr);
            }
        }
        rows += 1;
        last_t = t;
    }

    // the trajectory has ended: the rows of the incomplete last groups
    __device__ __forceinline__ void finish(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj) const {
        const int stored = (rows < a.row_stride) ? rows : a.row_stride;
        if (a.y_eval) RowsY::finish(RowSink{a.y_eval, a.row_stride, a.rows_vec}, buf, lane, traj, stored);
        if (a.t_out) RowsT::finish(RowSink{a.t_out, a.row_stride, a.tout_vec}, buf + RowsY::SLOTS, lane, traj, stored);
    }

    __device__ __forceinline__ static double component(const double (&y)[N], int idx_) {
        double v = y[0];
#pragma unroll
        for (int c = 1; c < N; c++) v = (c == idx_) ? y[c] : v;
        return v;
    }

    // signed_distance((extractor)(y)), hyperplane.rs:160-162: (pos - point) . normal, summed from zero in index order
    __device__ __forceinline__ static double plane_distance(const OdeKernelArgs& a, const double (&y)[N]) {
        double sum = 0.0;
        for (int i = 0; i < a.plane_dim; i++) {
            const double d = component(y, a.plane_index[i]) + (-1.0) * a.plane_point[i];
            sum = sum + d * a.plane_normal[i];
        }
        return sum;
    }

    // find_crossing_newton of the hyperplane recorder, hyperplane.rs:262-330 (no step-size exit, derivative floor eps)
    __device__ __forceinline__ bool newton_plane(const OdeKernelArgs& a, const DenseStep<Sys, Tab>& ds, double t_lower, double t_upper,
                                                 double dist_lower, double dist_upper, double* t_out) const {
        double t = t_lower - dist_lower * (t_upper - t_lower) / (dist_upper - dist_lower);
        const double tolerance = DBL_EPSILON * 100.0;
        double row[N];
        double dist;
        for (int it = 0; it < 10; it++) {
            ds.eval(t, row);
            dist = plane_distance(a, row);
            if (fabs(dist) < tolerance) { *t_out = t; return true; }
            const double delta_t = (t_upper - t_lower) * 1e-6;
            ds.eval(t + delta_t, row);
            const double dist_plus = plane_distance(a, row);
            const double derivative = (dist_plus - dist) / delta_t;
            if (fabs(derivative) < DBL_EPSILON) break;
            const double t_next = t - dist / derivative;
            t = (t_next < t_lower || t_next > t_upper) ? (t_lower + t_upper) / 2.0 : t_next;
        }
        ds.eval(t, row);
        dist = plane_distance(a, row);
        *t_out = t;
        return fabs(dist) < tolerance * 10.0;
    }

    // the solout call that precedes the loop (solve_ivp.rs:160): t_prev == t_curr == t0
    __device__ __forceinline__ void first(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, double t0, const double (&y0)[N],
                                          const double* p) {
        if (rec_mode_of(a) == REC_CROSSING) {
            last_off = component(y0, a.cross_component) - a.cross_threshold;
            have_last = true;
        } else if (rec_mode_of(a) == REC_HYPERPLANE) {
            last_off = plane_distance(a, y0);
            have_last = true;
        } else if (rec_mode_of(a) == REC_T_EVAL || rec_mode_of(a) == REC_EVEN) {
            if (a.emit_t0) { push(a, buf, lane, traj, t0, y0); idx = 1; }  // the row plan starts with t0 (t_eval.rs:113-114 / even.rs:100-104)
        } else {
            push(a, buf, lane, traj, t0, y0);  // Default: always; Dense: t_prev == t_curr, only the point itself
        }
        if (Evt::ENABLED) last_g = Evt::g(a, t0, y0, p);  // event.rs:373-380: first call only stores g
    }

    // find_crossing_newton, crossing.rs:185-263
    __device__ __forceinline__ bool newton(const OdeKernelArgs& a, const DenseStep<Sys, Tab>& ds, double t_lower, double t_upper,
                                           double off_lower, double off_upper, double* t_out) const {
        double t = t_lower - off_lower * (t_upper - t_lower) / (off_upper - off_lower);
        const double tolerance = DBL_EPSILON * 100.0;
        double row[N];
        double off;
        for (int it = 0; it < 10; it++) {
            ds.eval(t, row);
            off = component(row, a.cross_component) - a.cross_threshold;
            if (fabs(off) < tolerance) { *t_out = t; return true; }
            const double delta_t = (t_upper - t_lower) * 1e-6;
            const double t_plus = t + delta_t;
            ds.eval(t_plus, row);
            const double off_plus = component(row, a.cross_component) - a.cross_threshold;
            const double derivative = (off_plus - off) / delta_t;
            if (fabs(derivative) < DBL_EPSILON * 10.0) break;
            const double t_next = t - off / derivative;
            if (t_next < t_lower || t_next > t_upper) {
                t = (t_lower + t_upper) / 2.0;
            } else {
                const double change = fabs(t_next - t);
                if (change < tolerance * 0.1) { t = t_next; break; }
                t = t_next;
            }
        }
        ds.eval(t, row);
        off = component(row, a.cross_component) - a.cross_threshold;
        *t_out = t;
        return fabs(off) < tolerance * 10.0;
    }

    // brent_dekker, event.rs:386-470.  `interpolate(b).ok()?`: a point outside [t_prev, t_curr] AS THE REFERENCE TESTS IT
    // (t < t_prev || t > t_curr, written for forward time) ends the search without an event.
    __device__ __forceinline__ bool brent_dekker(const OdeKernelArgs& ka, const DenseStep<Sys, Tab>& ds, const double* p, double a, double b,
                                                 double fa, double fb, double* t_event) const {
        const double rel_tol = 1e-12, abs_tol = 1e-14;
        if (fabs(fa) < fabs(fb)) { double x = a; a = b; b = x; x = fa; fa = fb; fb = x; }
        double c = a, fc = fa, d = b - a, e = d;
        for (int it = 0; it < 50; it++) {
            if (fb == 0.0) { *t_event = b; return true; }
            if (d_signum(fa) == d_signum(fb)) { a = c; fa = fc; c = b; fc = fb; d = b - a; e = d; }
            if (fabs(fa) < fabs(fb)) { c = b; b = a; a = c; fc = fb; fb = fa; fa = fc; }
            const double tol = fmax(abs_tol, rel_tol * fabs(b));
            const double m = 0.5 * (a - b);
            if (fabs(m) <= tol || fb == 0.0) { *t_event = b; return true; }
            bool use_bis = true;
            if (fabs(e) > tol && fabs(fa) > fabs(fb)) {
                const double s = fb / fa;
                double pp, q;
                if (a == c) {
                    pp = 2.0 * m * s;
                    q = 1.0 - s;
                } else {
                    const double q1 = fa / fc;
                    const double r = fb / fc;
                    pp = s * (2.0 * m * q1 * (q1 - r) - (b - a) * (r - 1.0));
                    q = (q1 - 1.0) * (r - 1.0) * (s - 1.0);
                }
                double q_mod = q, p_mod = pp;
                if (q_mod > 0.0) p_mod = -p_mod; else q_mod = -q_mod;
                if (fabs(2.0 * p_mod) < (3.0 * m * q_mod - fabs(tol * q_mod)) && p_mod < fabs(e * 0.5 * q_mod)) {
                    e = d;
                    d = p_mod / q_mod;
                    use_bis = false;
                }
            }
            if (use_bis) { d = m; e = m; }
            a = b;
            fa = fb;
            b = (fabs(d) > tol) ? (b + d) : (b + ((m > 0.0) ? tol : -tol));
            if (b < ds.t || b > ds.t_new) return false;  // interpolate(b) is Err(OutOfBounds)
            double yb[N];
            ds.eval(b, yb);
            fb = Evt::g(ka, b, yb, p);
            c = a;
            fc = fa;
        }
        return false;
    }

    // Does the solout call for the accepted step ending at (t_new, yn) need a refinement -- a crossing / event root search on
    // the dense output, or interpolated rows?  (Same tests as in step(); a superset is harmless: the caller only uses the
    // answer to let such lanes of a warp refine together.)  The dense(n) recorder interpolates in every step of every lane.
    __device__ __forceinline__ bool slow_work(const OdeKernelArgs& a, double t_new, const double (&yn)[N], const double* p) const {
        bool slow = false;
        if (rec_mode_of(a) == REC_CROSSING) {
            const double off = component(yn, a.cross_component) - a.cross_threshold;
            slow = have_last && d_signum(last_off) != d_signum(off);
        } else if (rec_mode_of(a) == REC_HYPERPLANE) {
            const double dist = plane_distance(a, yn);
            slow = have_last && (d_signum(last_off) != d_signum(dist) || (last_off == 0.0) != (dist == 0.0));
        } else if ((rec_mode_of(a) == REC_T_EVAL || rec_mode_of(a) == REC_EVEN) && a.rec_park_rows) {
            const double dir = d_signum(a.tf - a.t0);
            slow = idx < a.n_rows && (a.t_rows[idx] - t_new) * dir <= 0.0;
        }
        if (Evt::ENABLED) slow = slow || d_signum(last_g) != d_signum(Evt::g(a, t_new, yn, p));
        return slow;
    }

    // solout after an accepted step from (t, y) to (t + h, yn); k[0] = f(t, y), dydt = f(t + h, yn).  Returns true when
    // an event asks to terminate (ControlFlag::Terminate).
    __device__ __forceinline__ bool step(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, double t, double h, const double (&y)[N],
                                         const double (&yn)[N], const double (&k)[S][N], const double (&dydt)[N], const double* p) {
        const double t_new = t + h;
        const double dir = d_signum(a.tf - a.t0);
        DenseStep<Sys, Tab> ds;
        bool prepared = false;
        // ---- base recorder
        if (rec_mode_of(a) == REC_DEFAULT) {
            push(a, buf, lane, traj, t_new, yn);
        } else if (rec_mode_of(a) == REC_DENSE) {
            if (t != t_new && a.dense_n > 1) {
                ds.prepare(t, h, y, yn, k, dydt, p);
                prepared = true;
                for (int i = 1; i < a.dense_n; i++) {
                    const double h_old = t_new - t;
                    const double ti = t + (double)i * h_old / (double)a.dense_n;
                    double row[N];
                    ds.eval(ti, row);
                    push(a, buf, lane, traj, ti, row);
                }
            }
            push(a, buf, lane, traj, t_new, yn);
        } else if (rec_mode_of(a) == REC_CROSSING) {
            const double off = component(yn, a.cross_component) - a.cross_threshold;
            if (have_last) {
                const bool is_crossing = d_signum(last_off) != d_signum(off);  // NaN != anything
                if (is_crossing) {
                    const bool record = (a.cross_direction > 0) ? (last_off < 0.0 && off >= 0.0)
                                      : (a.cross_direction < 0) ? (last_off > 0.0 && off <= 0.0) : true;
                    if (record) {
                        ds.prepare(t, h, y, yn, k, dydt, p);
                        prepared = true;
                        double t_cross;
                        if (!newton(a, ds, t, t_new, last_off, off, &t_cross)) {
                            const double frac = -last_off / (off - last_off);
                            t_cross = t + frac * (t_new - t);
                        }
                        double row[N];
                        ds.eval(t_cross, row);
                        push(a, buf, lane, traj, t_cross, row);
                    }
                }
            }
            last_off = off;
            have_last = true;
        } else if (rec_mode_of(a) == REC_HYPERPLANE) {
            const double dist = plane_distance(a, yn);
            if (have_last) {
                const double last = last_off;
                const bool is_crossing = d_signum(last) != d_signum(dist) || (last == 0.0 && dist != 0.0) || (last != 0.0 && dist == 0.0);
                if (is_crossing) {
                    const bool record = (a.cross_direction > 0) ? (last < 0.0 && dist >= 0.0)
                                      : (a.cross_direction < 0) ? (last > 0.0 && dist <= 0.0) : true;
                    if (record) {
                        ds.prepare(t, h, y, yn, k, dydt, p);
                        prepared = true;
                        double t_cross;
                        if (!newton_plane(a, ds, t, t_new, last, dist, &t_cross)) {
                            const double frac = -last / (dist - last);
                            t_cross = t + frac * (t_new - t);
                        }
                        double row[N];
                        ds.eval(t_cross, row);
                        push(a, buf, lane, traj, t_cross, row);
                    }
                }
            }
            last_off = dist;
            have_last = true;
        } else {
            // t_eval / even(dt) through the host's row plan (see erk_ensemble.cuh: every planned point a step passes is emitted;
            // even: always interpolated, and the last plan entry is the tf sentinel of the final-point rule, even.rs:166-188)
            while (idx < a.n_rows && (a.t_rows[idx] - t_new) * dir <= 0.0) {
                const double te = a.t_rows[idx];
                if (rec_mode_of(a) == REC_EVEN && idx == a.n_rows - 1) {
                    if (t_new == a.tf) {
                        const double t_prev_row = a.t_rows[idx - 1];
                        if (fabs(t_prev_row - a.tf) <= a.even_tol) rows -= 1;  // solution.pop(): replace the near-duplicate
                        push(a, buf, lane, traj, a.tf, yn);
                    }
                    idx = a.n_rows;
                    break;
                }
                if (te == t_new && rec_mode_of(a) == REC_T_EVAL) {
                    push(a, buf, lane, traj, te, yn);
                } else {
                    if (!prepared) { ds.prepare(t, h, y, yn, k, dydt, p); prepared = true; }
                    double row[N];
                    ds.eval(te, row);
                    push(a, buf, lane, traj, te, row);
                }
                idx += 1;
            }
        }
        // ---- event detection, event.rs:361-384 (detect_event)
        bool terminate = false;
        if (Evt::ENABLED) {
            const double g_curr = Evt::g(a, t_new, yn, p);
            const double g_prev = last_g;
            const bool sign_change = d_signum(g_prev) != d_signum(g_curr);
            const bool direction_ok = (a.event_direction > 0) ? (sign_change && g_prev < 0.0 && g_curr >= 0.0)
                                    : (a.event_direction < 0) ? (sign_change && g_prev > 0.0 && g_curr <= 0.0) : sign_change;
            if (direction_ok) {
                double ea = t, eb = t_new, fa = g_prev, fb = g_curr;
                if ((dir > 0.0 && ea > eb) || (dir < 0.0 && ea < eb)) { double x = ea; ea = eb; eb = x; x = fa; fa = fb; fb = x; }
                if (fa * fb <= 0.0) {
                    if (!prepared) { ds.prepare(t, h, y, yn, k, dydt, p); prepared = true; }
                    double t_event;
                    if (brent_dekker(a, ds, p, ea, eb, fa, fb, &t_event)) {
                        double row[N];
                        ds.eval(t_event, row);
                        const bool push_point = (rows > 0) ? (fabs(t_event - last_t) > 1e-14) : true;
                        if (push_point) push(a, buf, lane, traj, t_event, row);
                        event_count += 1;
                        if (a.event_terminate > 0 && event_count >= a.event_terminate) terminate = true;
                    }
                }
            }
            last_g = g_curr;
        }
        return terminate;
    }
};

}  // namespace deb
