// erk_fixed.cuh -- fixed-step explicit Runge-Kutta ensembles (Euler, midpoint, Heun, Ralston, SSP-RK3, RK4, 3/8).
//
// Fuses, per trajectory (one per thread, state in registers):
//     solve_ode loop               /root/reference/src/ode/solve_ivp.rs:139-277
//     Fixed init/step              /root/reference/src/methods/erk/fixed/ordinary.rs:16-139
//     cubic-Hermite dense output   /root/reference/src/interpolate.rs:40-60 (via fixed/ordinary.rs:206-216)
//     TEvalSolout                  /root/reference/src/solout/t_eval.rs:87-137
// All trajectories of an ensemble share t0, tf and h, so every lane takes the same number of steps: no queue and
// no divergence; a plain grid-stride loop over trajectories is enough.
// Note the association of the solution update here is y + sum_i (b_i*h)*k_i (fixed/ordinary.rs:98-102), which is
// NOT the Dormand-Prince form y + h*(sum_i b_i*k_i).
#pragma once
#include "erk_ensemble.cuh"

namespace deb {

template <class Sys, class Tab, int BLOCK, bool REC = false, class Evt = EvtNone>
__global__ void __launch_bounds__(BLOCK) fixed_ensemble_kernel(const OdeKernelArgs a) {
    constexpr int N = Sys::DIM, NP = Sys::NP, S = Tab::S;
    const double t0 = a.t0, tf = a.tf;
    const double dir = d_signum(tf - t0);
    const double te_none = (dir > 0.0) ? (1.0 / 0.0) : -(1.0 / 0.0);
    const long long stride = (long long)gridDim.x * BLOCK;
    using Rows = RowStage<N, BLOCK, !REC>;  // whole-sector row groups, see erk_ensemble.cuh
    using Recorder = StepRecorder<Sys, Tab, Evt, BLOCK>;
    __shared__ double s_lane[BLOCK / 32][REC ? Recorder::STAGE_SLOTS : Rows::SLOTS][32];
    double (*s_rows)[32] = s_lane[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;

    for (long long traj = (long long)blockIdx.x * BLOCK + threadIdx.x; traj < a.n_traj; traj += stride) {
        double y[N], dydt[N], k[S][N], p[NP > 0 ? NP : 1];
#pragma unroll
        for (int c = 0; c < N; c++) y[c] = a.y0[traj * N + c];
#pragma unroll
        for (int q = 0; q < NP; q++) p[q] = a.params ? a.params[traj * a.params_stride + q] : a.pc[q];
        int steps = 0, evals = 0, n_emit = 0, idx = 0;
        int fin = -1;
        Recorder recd;
        double t = t0;
        // ---- init, fixed/ordinary.rs:16-56 (BadInput was decided on the host: utils.rs:60-157 does not look at the state)
        const double h_full = (a.h0 == 0.0) ? fabs(tf - t0) / 100.0 : a.h0;
        double h = h_full;
        if (a.fx_status == DEB_STATUS_BAD_INPUT) {
            fin = DEB_STATUS_BAD_INPUT;
        } else {
            Sys::rhs(t, y, dydt, p);
            evals = 1;
            if (!REC && a.emit_t0) {
                if (a.y_eval) Rows::put(a, s_rows, lane, traj, 0, 0, y);
                n_emit = 1;
                idx = 1;
            }
            if constexpr (REC) recd.first(a, s_rows, lane, traj, t0, y, p);  // the solout call that precedes the loop
        }
        double te = (idx < a.n_rows) ? a.t_rows[idx] : te_none;
        // the loop head (clip at tf, solve_ivp.rs:193-209), the max_steps test and the end test (:263) ran on the host
        for (int step = 0; fin < 0 && step < a.fx_n_steps; step++) {
            const int tail = step - (a.fx_n_steps - a.fx_n_tail);  // >= 0: one of the clipped steps at the end
            h = (tail >= 0) ? a.fx_h_tail[tail] : h_full;
            steps += 1;
            // ---- step, fixed/ordinary.rs:58-139
#pragma unroll
            for (int c = 0; c < N; c++) k[0][c] = dydt[c];
#pragma unroll
            for (int i = 1; i < S; i++) {
                double ys[N];
#pragma unroll
                for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
                for (int j = 0; j < i; j++) {
                    const double aij = Tab::a(i, j);
                    if (aij != 0.0) {
                        const double ah = Tab::av(i, j) * h;
#pragma unroll
                        for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * k[j][c];
                    }
                }
                Sys::rhs(t + Tab::cv(i) * h, ys, k[i], p);
            }
            double ynew[N], dnew[N];
#pragma unroll
            for (int c = 0; c < N; c++) ynew[c] = y[c];
#pragma unroll
            for (int i = 0; i < S; i++) {
                if (Tab::b(i) != 0.0) {
                    const double bh = Tab::bv(i) * h;
#pragma unroll
                    for (int c = 0; c < N; c++) ynew[c] = ynew[c] + bh * k[i][c];
                }
            }
            const double t_new = t + h;
            Sys::rhs(t_new, ynew, dnew, p);
            evals += S;  // S-1 stages + the new derivative (fsal = false)
            bool interrupt = false;
            if constexpr (REC) interrupt = recd.step(a, s_rows, lane, traj, t, h, y, ynew, k, dnew, p);
            // ---- TEvalSolout with cubic Hermite interpolation
            while (!REC && ((dir > 0.0) ? (te <= t_new) : (te >= t_new))) {
                if (a.even && idx == a.n_rows - 1) {  // the tf sentinel: EvenSolout final-point rule, even.rs:166-188
                    int w = -1;
                    if (t_new == tf) {
                        const double t_last = a.t_rows[idx - 1];
                        w = (fabs(t_last - tf) <= a.even_tol) ? idx - 1 : idx;
                    }
                    if (w >= 0 && a.y_eval) Rows::put(a, s_rows, lane, traj, w, n_emit, ynew);
                    if (w == idx) { idx += 1; n_emit += 1; }
                    te = te_none;
                    break;
                }
                double row[N];
                if (te == t_new && !a.even) {
#pragma unroll
                    for (int c = 0; c < N; c++) row[c] = ynew[c];
                } else {
                    const double hh = t_new - t;
                    const double s = (te - t) / hh;
                    const double s2 = s * s, s3 = s2 * s;
                    const double h00 = 2.0 * s3 - 3.0 * s2 + 1.0;
                    const double h10 = s3 - 2.0 * s2 + s;
                    const double h01 = -2.0 * s3 + 3.0 * s2;
                    const double h11 = s3 - s2;
                    const double w10 = h10 * hh, w11 = h11 * hh;
#pragma unroll
                    for (int c = 0; c < N; c++) {  // linear_combination: 0, then += term by term (traits.rs:357-370)
                        double v = __dadd_rn(0.0, h00 * y[c]);
                        v = v + w10 * dydt[c];
                        v = v + h01 * ynew[c];
                        v = v + w11 * dnew[c];
                        row[c] = v;
                    }
                }
                if (a.y_eval) Rows::put(a, s_rows, lane, traj, n_emit, n_emit, row);
                n_emit += 1;
                idx += 1;
                te = (idx < a.n_rows) ? a.t_rows[idx] : te_none;
            }
            t = t_new;
#pragma unroll
            for (int c = 0; c < N; c++) { y[c] = ynew[c]; dydt[c] = dnew[c]; }
            if (REC && interrupt) fin = DEB_STATUS_INTERRUPTED;    // solve_ivp.rs:255-260 (an event asked to terminate)
        }
        if (fin < 0) fin = a.fx_status;  // Complete, or MaxSteps after fx_n_steps steps
        if (!REC && a.y_eval) Rows::finish(a, s_rows, lane, traj, n_emit);
        if (a.status) a.status[traj] = fin;
        if (a.t_final) a.t_final[traj] = t;
        if (a.y_final) {
#pragma unroll
            for (int c = 0; c < N; c++) a.y_final[traj * N + c] = y[c];
        }
        if (a.accepted) a.accepted[traj] = steps;
        if (a.rejected) a.rejected[traj] = 0;
        if (a.evals) a.evals[traj] = evals;
        if constexpr (REC) recd.finish(a, s_rows, lane, traj);
        if (a.n_emitted) a.n_emitted[traj] = REC ? recd.rows : n_emit;
        wm_publish(a, traj);
    }
}

}  // namespace deb
