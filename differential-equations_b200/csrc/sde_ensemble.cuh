// sde_ensemble.cuh -- fixed-step stochastic ERK ensembles (Euler-Maruyama = euler(h) on an SDE), diagonal noise.
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
    int y0_stride;       // DIM, or 0 when one y0 is shared by all paths
    const double* params;
    int params_stride;   // NP, or 0 when shared
    double pc[8];        // the shared parameter set by value (params == nullptr)
    long long n_traj;
    long long path_offset;
    unsigned long long seed;
    unsigned int round_keys[20];  // Philox round keys of `seed`: (lo32 + r*0x9E3779B9, hi32 + r*0xBB67AE85), r = 0..9
    double t0, tf, h0, h_min, h_max;
    int max_steps;
    const double* t_rows;
    int n_rows, row_stride, emit_t0;
    // The step schedule does not depend on the path (same t0, tf, h0, t_eval for every path), so the host runs the
    // solve_sde bookkeeping once (solve_ivp.rs:211-227, :263; stochastic.rs:74-83) and the kernel's step loop carries no
    // end-of-interval tests: n_steps steps of size h0 except the last n_tail (>= 1) ones, of sizes h_tail[0..n_tail) (the
    // clip at tf can fire twice, see OdeKernelArgs); row r is emitted in step row_step[r] with the interpolation weight
    // row_s[r] (negative: the row is the end of the step itself).
    int n_steps;
    int n_tail;
    double h_tail[DEB_FX_MAX_TAIL];
    int final_status;       // DEB_STATUS_COMPLETE / MAX_STEPS / BAD_INPUT (then n_steps = 0)
    const int* row_step;    // [n_rows]
    const double* row_s;    // [n_rows]
    double* y_eval;
    int rows_vec;           // y_eval row groups start 32-byte aligned (RowStage, erk_ensemble.cuh)
    int* n_emitted;
    double* t_final;
    double* y_final;
    int* status;
    int* accepted;
    int* rejected;
    int* evals;
};

// MILSTEIN: derivative-free Milstein step (/root/reference/src/methods/milstein.rs:107-180) instead of the stochastic ERK step
// (Tab is then unused).  Vector states with diagonal noise: component c of step s uses normal number s*DIM + c.
template <class Sde, class Tab, int BLOCK, bool MILSTEIN = false>
__global__ void __launch_bounds__(BLOCK) sde_ensemble_kernel(const SdeKernelArgs a) {
    constexpr int N = Sde::DIM, NP = Sde::NP, S = Tab::S;
    const double t0 = a.t0, tf = a.tf;
    const long long stride = (long long)gridDim.x * BLOCK;
    using Rows = RowStage<N, BLOCK>;  // rows leave the chip as whole 32-byte sectors (erk_ensemble.cuh)
    __shared__ double s_lane[BLOCK / 32][Rows::SLOTS][32];
    double (*s_rows)[32] = s_lane[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const RowSink sink{a.y_eval, a.row_stride, a.rows_vec};

    for (long long traj = (long long)blockIdx.x * BLOCK + threadIdx.x; traj < a.n_traj; traj += stride) {
        double p[(NP + Sde::NPX) > 0 ? (NP + Sde::NPX) : 1];
#pragma unroll
        for (int q = 0; q < NP; q++) p[q] = a.params ? a.params[traj * a.params_stride + q] : a.pc[q];
        Sde::prepare(p);
        double y[N], dydt[N];
#pragma unroll
        for (int c = 0; c < N; c++) { y[c] = a.y0[traj * a.y0_stride + c]; dydt[c] = 0.0; }
        const unsigned long long path = (unsigned long long)(a.path_offset + traj);
        int evals = 0, n_emit = 0, idx = 0;
        double t = t0;
        const double h0 = (a.h0 == 0.0) ? fabs(tf - t0) / 100.0 : a.h0;  // stochastic.rs:23-28
        // ---- init, stochastic.rs:18-65 (BadInput was decided on the host: utils.rs:60-157 does not look at the state)
        if (a.final_status != DEB_STATUS_BAD_INPUT) {
            Sde::drift(t, y, dydt, p);
            evals = MILSTEIN ? 1 : 2;  // ERK: drift + diffusion (the initial diffusion value is not used); Milstein: drift only
            if (a.emit_t0) {
                if (a.y_eval) Rows::put(sink, s_rows, lane, traj, 0, 0, y);
                n_emit = 1;
                idx = 1;
            }
        }
        int next_step = (idx < a.n_rows) ? a.row_step[idx] : -1;
        const double sqrt_h0 = sqrt(h0);
        const int n_main = a.n_steps - a.n_tail;  // steps of size h0
        // one step from (t, y, dydt) with the Wiener increments dw
        auto do_step = [&](const int step, const double h, const double sqrt_h, double (&dw)[N]) {
            Sde::mix(dw, p);
            double y_next[N], g[N];
            Sde::diffusion(t, y, g, p);  // stochastic.rs:113-115 / milstein.rs:127-130
            if (MILSTEIN) {
                double y_aux[N], g_aux[N];
#pragma unroll
                for (int c = 0; c < N; c++) y_aux[c] = y[c] + sqrt_h * g[c];  // b(t_n, y_n + b sqrt(h))
                Sde::diffusion(t, y_aux, g_aux, p);
                const double factor = 1.0 / (2.0 * sqrt_h);
#pragma unroll
                for (int c = 0; c < N; c++) {
                    const double milstein_term = ((g_aux[c] - g[c]) * (dw[c] * dw[c] - h)) * factor;  // milstein.rs:148-156
                    y_next[c] = ((y[c] + dydt[c] * h) + g[c] * dw[c]) + milstein_term;                // milstein.rs:158-170
                }
                evals += 3;  // diffusion, auxiliary diffusion, new drift
            } else {
                double k[S][N];
#pragma unroll
                for (int c = 0; c < N; c++) k[0][c] = dydt[c];
#pragma unroll
                for (int i = 1; i < S; i++) {  // drift stages, stochastic.rs:96-104
                    double ys[N];
#pragma unroll
                    for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
                    for (int j = 0; j < i; j++) {
                        if (Tab::a(i, j) != 0.0) {
                            const double ah = Tab::av(i, j) * h;
#pragma unroll
                            for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * k[j][c];
                        }
                    }
                    Sde::drift(t + Tab::cv(i) * h, ys, k[i], p);
                }
#pragma unroll
                for (int c = 0; c < N; c++) {
                    double drift_inc = 0.0;  // stochastic.rs:107-110
#pragma unroll
                    for (int i = 0; i < S; i++) {
                        if (Tab::b(i) != 0.0) drift_inc = __dadd_rn(drift_inc, (Tab::bv(i) * h) * k[i][c]);
                    }
                    y_next[c] = (y[c] + drift_inc) + g[c] * dw[c];  // stochastic.rs:122-128 (coefficients 1.0)
                }
                evals += S + 1;  // S-1 drift stages + diffusion + new drift
            }
            const double t_new = t + h;
            double d_new[N];
            Sde::drift(t_new, y_next, d_new, p);
            while (next_step == step) {  // TEvalSolout rows of this step (t_eval.rs:100-130), linear dense output
                const double sw = a.row_s[idx];
                double row[N];
                if (sw < 0.0) {
#pragma unroll
                    for (int c = 0; c < N; c++) row[c] = y_next[c];
                } else {
#pragma unroll
                    for (int c = 0; c < N; c++) row[c] = __dadd_rn(0.0, (1.0 - sw) * y[c]) + sw * y_next[c];
                }
                if (a.y_eval) Rows::put(sink, s_rows, lane, traj, n_emit, n_emit, row);
                n_emit += 1;
                idx += 1;
                next_step = (idx < a.n_rows) ? a.row_step[idx] : -1;
            }
            t = t_new;
#pragma unroll
            for (int c = 0; c < N; c++) { y[c] = y_next[c]; dydt[c] = d_new[c]; }
        };
        // noise(h, dw): component c of step s is normal number s*N + c of the path's stream, times sqrt(h)
        int step = 0;
        if constexpr (N == 1) {
            // one normal per step: a Philox call feeds two consecutive steps; all but the last step have size h0
            for (; step + 2 <= n_main; step += 2) {
                double ze, zo;
                normal_pair(a.round_keys, path, (unsigned long long)(step >> 1), &ze, &zo);
                double dw0[1] = {sqrt_h0 * ze}, dw1[1] = {sqrt_h0 * zo};
                do_step(step, h0, sqrt_h0, dw0);
                do_step(step + 1, h0, sqrt_h0, dw1);
            }
        }
        double z_odd = 0.0;                          // the odd normal of the last Philox call ...
        unsigned long long odd_pair = ~0ull;         // ... and its pair index
        for (; step < a.n_steps; step++) {
            const int tail = step - n_main;
            const double h = (tail >= 0) ? a.h_tail[tail] : h0;
            const double sqrt_h = (tail >= 0) ? sqrt(h) : sqrt_h0;
            const unsigned long long q0 = (unsigned long long)step * N;  // first normal index of this step
            double dw[N];
            if constexpr (N % 2 == 0) {
                // even dimension: the step's normals are whole Philox pairs
#pragma unroll
                for (int c = 0; c < N; c += 2) {
                    double ze, zo;
                    normal_pair(a.round_keys, path, (q0 + c) >> 1, &ze, &zo);
                    dw[c] = sqrt_h * ze;
                    dw[c + 1] = sqrt_h * zo;
                }
            } else {
#pragma unroll
                for (int c = 0; c < N; c++) {
                    const unsigned long long q = q0 + c;
                    double z;
                    if ((q & 1ull) == 0) { normal_pair(a.round_keys, path, q >> 1, &z, &z_odd); odd_pair = q >> 1; }
                    else if (odd_pair == (q >> 1)) z = z_odd;
                    else { double ze; normal_pair(a.round_keys, path, q >> 1, &ze, &z); }
                    dw[c] = sqrt_h * z;
                }
            }
            do_step(step, h, sqrt_h, dw);
        }
        const int fin = a.final_status;
        const int steps = a.n_steps;
        if (a.y_eval) Rows::finish(sink, s_rows, lane, traj, n_emit);
        if (a.status) a.status[traj] = fin;
        if (a.t_final) a.t_final[traj] = t;
        if (a.y_final) {
#pragma unroll
            for (int c = 0; c < N; c++) a.y_final[traj * N + c] = y[c];
        }
        if (a.accepted) a.accepted[traj] = steps;
        if (a.rejected) a.rejected[traj] = 0;
        if (a.evals) a.evals[traj] = evals;
        if (a.n_emitted) a.n_emitted[traj] = n_emit;
    }
}

}  // namespace deb
