/* A plain-C caller of the drop-in boundary (include/deb_ensemble.h): no CUDA toolkit, no Python, host buffers only.
 *
 * The same problem as `IVP::ode(&lorenz, 0.0, 100.0, y0).t_eval([1.0, 2.5, 100.0]).method(ExplicitRungeKutta::dopri5().rtol(1e-8)).solve()`
 * (src/ivp.rs:279,656,632,781), for 4096 initial states at once.  Trajectory 0 starts at (1, 1, 1): its step counts and final state are
 * the known answers of SURVEY.md appendix A, case L100 (6008 accepted / 411 rejected / 44525 evaluations; Lorenz is chaotic over
 * t in [0, 100], so the final state only matches when every operation of every step does), checked here bit for bit.
 *
 *   gcc -O2 -I../../include lorenz_ensemble.c -L../../differential-equations_b200 -ldeb200 -Wl,-rpath,'$ORIGIN/../../differential-equations_b200' -o lorenz_ensemble
 *
 * Exit code: 0 = results as expected, 3 = no CUDA device (the library has no CPU fallback), 1 = anything else. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "deb_ensemble.h"

int main(void) {
    enum { N = 4096, DIM = 3, N_EVAL = 3 };
    if (deb_abi_version() != DEB_ABI_VERSION) {
        fprintf(stderr, "header is ABI %d, library is ABI %d\n", DEB_ABI_VERSION, deb_abi_version());
        return 1;
    }
    double* y0 = malloc(sizeof(double) * N * DIM);
    double* y_eval = malloc(sizeof(double) * N * N_EVAL * DIM);
    double* y_final = malloc(sizeof(double) * N * DIM);
    double* t_final = malloc(sizeof(double) * N);
    int32_t* status = malloc(sizeof(int32_t) * N);
    int32_t* accepted = malloc(sizeof(int32_t) * N);
    int32_t* rejected = malloc(sizeof(int32_t) * N);
    int32_t* evals = malloc(sizeof(int32_t) * N);
    int32_t* n_emitted = malloc(sizeof(int32_t) * N);
    for (int i = 0; i < N; i++)
        for (int c = 0; c < DIM; c++) y0[i * DIM + c] = 1.0 + 1e-3 * i * (c + 1);  /* trajectory 0: (1, 1, 1) */
    const double params[3] = {10.0, 28.0, 8.0 / 3.0};
    const double t_eval[N_EVAL] = {1.0, 2.5, 100.0};
    double t_rows[N_EVAL];

    deb_ode_problem P;
    memset(&P, 0, sizeof P);
    P.struct_size = sizeof P;
    P.system = DEB_SYS_LORENZ;
    P.method = DEB_DOPRI5;
    P.dim = DIM;
    P.n_params = 3;
    P.n_traj = N;
    P.y0 = y0;
    P.params = params;
    P.params_shared = 1;
    P.n_eval = N_EVAL;
    P.t_eval = t_eval;
    P.t0 = 0.0;
    P.tf = 100.0;
    deb_erk_options_default(&P.opt);
    P.opt.rtol = 1e-8;
    P.device = 0;
    P.memspace = DEB_MEM_HOST;

    deb_result R;
    memset(&R, 0, sizeof R);
    R.struct_size = sizeof R;
    R.y_eval = y_eval;
    R.n_emitted = n_emitted;
    R.t_final = t_final;
    R.y_final = y_final;
    R.status = status;
    R.accepted = accepted;
    R.rejected = rejected;
    R.evals = evals;
    R.t_rows = t_rows;

    const int rc = deb_solve_ode(&P, &R);
    if (rc != DEB_OK) {
        fprintf(stderr, "deb_solve_ode: %d (%s)\n", rc, deb_last_error());
        return rc == DEB_ERR_NO_DEVICE ? 3 : 1;
    }
    long long acc = 0, rej = 0;
    int complete = 0;
    for (int i = 0; i < N; i++) {
        acc += accepted[i];
        rej += rejected[i];
        complete += (status[i] == DEB_STATUS_COMPLETE && n_emitted[i] == N_EVAL);
    }
    printf("%d of %d trajectories complete; %lld accepted / %lld rejected steps; kernel %.2f ms, call %.2f ms, %d launches\n", complete, N, acc, rej,
           R.kernel_ms, R.total_ms, R.gpu_launches);
    printf("trajectory 0: %d accepted, %d rejected, %d evaluations, y(100) = (%a, %a, %a)\n", accepted[0], rejected[0], evals[0], y_final[0], y_final[1],
           y_final[2]);
    /* SURVEY.md appendix A, L100 and L10 t_eval */
    int ok = complete == N && accepted[0] == 6008 && rejected[0] == 411 && evals[0] == 44525;
    ok = ok && y_final[0] == -0x1.8c65ec78fb24fp+1 && y_final[1] == -0x1.4d4949b09d281p+1 && y_final[2] == 0x1.5bf4de76ed043p+4;
    ok = ok && y_eval[0] == -9.378567031548476 && y_eval[1] == -8.357039113339846 && y_eval[2] == 29.36231537793399;          /* t = 1.0 */
    ok = ok && y_eval[3] == -6.959579728354202 && y_eval[4] == -7.272475545395783 && y_eval[5] == 24.70312694098664;          /* t = 2.5 */
    ok = ok && y_eval[6] == y_final[0] && y_eval[7] == y_final[1] && y_eval[8] == y_final[2] && t_rows[2] == 100.0;           /* exact hit at tf */
    printf(ok ? "known answers reproduced\n" : "MISMATCH against the known answers\n");
    return ok ? 0 : 1;
}
