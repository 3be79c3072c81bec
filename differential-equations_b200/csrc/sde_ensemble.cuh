// sde_ensemble.cuh -- fixed-step stochastic ERK ensembles (Euler-Maruyama = euler(h) on an SDE), scalar state.
//
// Fuses, per path (one per thread, everything in registers, zero HBM traffic per step):
//     solve_sde loop             /root/reference/src/sde/solve_ivp.rs:158-287
//     Stochastic Fixed init/step /root/reference/src/methods/erk/fixed/stochastic.rs:18-146
//     linear dense output        /root/reference/src/interpolate.rs:71-74 (via stochastic.rs:177-190)
//     TEvalSolout                /root/reference/src/solout/t_eval.rs:87-137
// `SDE::noise` is the counter-based Philox stream of philox.h (two normals per Philox call, kept in registers).
#pragma once
#include "erk_ensemble.cuh"
#include "philox.h"

namespace deb {

struct SdeKernelArgs {
    const double* y0;
    int y0_stride;       // 1, or 0 when one y0 is shared by all paths
    const double* params;
    int params_stride;   // NP, or 0 when shared
    double pc[8];        // the shared parameter set by value (params == nullptr)
    long long n_traj;
    long long path_offset;
    unsigned long long seed;
    double t0, tf, h0, h_min, h_max;
    int max_steps;
    const double* t_rows;
    int n_rows, row_stride, emit_t0;
    double* y_eval;
    int* n_emitted;
    double* t_final;
    double* y_final;
    int* status;
    int* accepted;
    int* rejected;
    int* evals;
};

// MILSTEIN: derivative-free Milstein step (/root/reference/src/methods/milstein.rs:107-180) instead of the stochastic ERK step
// (Tab is then unused).
template <class Sde, class Tab, int BLOCK, bool MILSTEIN = false>
__global__ void __launch_bounds__(BLOCK) sde_ensemble_kernel(const SdeKernelArgs a) {
    constexpr int NP = Sde::NP, S = Tab::S;
    const double t0 = a.t0, tf = a.tf;
    const double dir = d_signum(tf - t0);
    const double eps10 = DBL_EPSILON * 10.0;
    const double te_none = (dir > 0.0) ? (1.0 / 0.0) : -(1.0 / 0.0);
    const long long stride = (long long)gridDim.x * BLOCK;

    for (long long traj = (long long)blockIdx.x * BLOCK + threadIdx.x; traj < a.n_traj; traj += stride) {
        double p[NP > 0 ? NP : 1];
#pragma unroll
        for (int q = 0; q < NP; q++) p[q] = a.params ? a.params[traj * a.params_stride + q] : a.pc[q];
        double y = a.y0[traj * a.y0_stride];
        const unsigned long long path = (unsigned long long)(a.path_offset + traj);
        int steps = 0, evals = 0, n_emit = 0, idx = 0, fin = -1;
        double t = t0, dydt = 0.0;
        // ---- init, stochastic.rs:18-65
        double h = a.h0;
        if (h == 0.0) h = fabs(tf - t0) / 100.0;
        if (!validate_step_size_parameters(h, a.h_min, a.h_max, t0, tf)) {
            fin = DEB_STATUS_BAD_INPUT;
        } else {
            dydt = Sde::drift(t, y, p);
            evals = MILSTEIN ? 1 : 2;  // ERK: drift + diffusion (the initial diffusion value is not used); Milstein: drift only
            if (a.emit_t0) {
                if (a.y_eval) a.y_eval[traj * a.row_stride] = y;
                n_emit = 1;
                idx = 1;
            }
        }
        double te = (idx < a.n_rows) ? a.t_rows[idx] : te_none;
        double h_cached = h, sqrt_h = sqrt(h);
        double z_odd = 0.0;
        while (fin < 0) {
            if ((t + h - tf) * dir > 0.0) {  // solve_ivp.rs:211-227
                const double h_new = tf - t;
                if (fabs(h_new) < eps10) { fin = DEB_STATUS_COMPLETE; break; }
                h = h_new;
            }
            if (steps >= a.max_steps) { fin = DEB_STATUS_MAX_STEPS; break; }  // stochastic.rs:74-83
            const unsigned long long q = (unsigned long long)steps;       // normal index of this step (dim = 1)
            steps += 1;
            if (h != h_cached) { h_cached = h; sqrt_h = sqrt(h); }
            double z;
            if ((q & 1ull) == 0) normal_pair(a.seed, path, q >> 1, &z, &z_odd);
            else z = z_odd;
            const double dw = sqrt_h * z;  // noise(h, dw)
            double y_next;
            if (MILSTEIN) {
                const double g = Sde::diffusion(t, y, p);
                const double g_aux = Sde::diffusion(t, y + sqrt_h * g, p);  // b(t_n, y_n + b sqrt(h))
                const double factor = 1.0 / (2.0 * sqrt_h);
                const double milstein_term = ((g_aux - g) * (dw * dw - h)) * factor;  // milstein.rs:148-156
                y_next = ((y + dydt * h) + g * dw) + milstein_term;                   // milstein.rs:158-170
                evals += 3;  // diffusion, auxiliary diffusion, new drift
            } else {
            double k[S];
            k[0] = dydt;
#pragma unroll
            for (int i = 1; i < S; i++) {  // drift stages, stochastic.rs:96-104
                double ys = y;
#pragma unroll
                for (int j = 0; j < i; j++) {
                    if (Tab::a(i, j) != 0.0) ys = ys + (Tab::av(i, j) * h) * k[j];
                }
                k[i] = Sde::drift(t + Tab::cv(i) * h, ys, p);
            }
            double drift_inc = 0.0;  // stochastic.rs:107-110
#pragma unroll
            for (int i = 0; i < S; i++) {
                if (Tab::b(i) != 0.0) drift_inc = __dadd_rn(drift_inc, (Tab::bv(i) * h) * k[i]);
            }
            const double g = Sde::diffusion(t, y, p);  // stochastic.rs:113-115
            y_next = (y + drift_inc) + g * dw;         // stochastic.rs:122-128 (coefficients 1.0); dw = noise(h), :118-119
            evals += S + 1;  // S-1 drift stages + diffusion + new drift
            }
            const double t_new = t + h;
            const double d_new = Sde::drift(t_new, y_next, p);
            while ((dir > 0.0) ? (te <= t_new) : (te >= t_new)) {
                double row;
                if (te == t_new) row = y_next;
                else {
                    const double s = (te - t) / (t_new - t);
                    row = __dadd_rn(0.0, (1.0 - s) * y) + s * y_next;
                }
                if (a.y_eval) a.y_eval[traj * a.row_stride + n_emit] = row;
                n_emit += 1;
                idx += 1;
                te = (idx < a.n_rows) ? a.t_rows[idx] : te_none;
            }
            t = t_new;
            y = y_next;
            dydt = d_new;
            if (fabs(tf - t) <= eps10) fin = DEB_STATUS_COMPLETE;
        }
        if (a.status) a.status[traj] = fin;
        if (a.t_final) a.t_final[traj] = t;
        if (a.y_final) a.y_final[traj] = y;
        if (a.accepted) a.accepted[traj] = steps;
        if (a.rejected) a.rejected[traj] = 0;
        if (a.evals) a.evals[traj] = evals;
        if (a.n_emitted) a.n_emitted[traj] = n_emit;
    }
}

}  // namespace deb
