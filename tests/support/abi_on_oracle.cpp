// TEST INFRASTRUCTURE ONLY.  Resolves the C-ABI entry points that include/deb_ensemble.hpp uses against the CPU oracle
// (oracle/liboracle.so), so that the C++ host mirror's marshalling -- row capacities, recorder and event fields, the
// per-trajectory Solution / Error view -- is exercised in the CPU test suite.  Never part of the product.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "deb_ensemble.h"

extern "C" int orc_solve_ode(const deb_ode_problem* P, deb_result* R, int n_threads);
extern "C" int orc_solve_sde(const deb_sde_problem* P, deb_result* R, int n_threads);
extern "C" int orc_solve_heat_mol(const deb_heat_problem* P, int n_threads);

static thread_local std::string g_err;

extern "C" {
int deb_abi_version(void) { return DEB_ABI_VERSION; }
const char* deb_last_error(void) { return g_err.c_str(); }
void deb_erk_options_default(deb_erk_options* o) {  // erk/mod.rs:135-144
    o->rtol = 1.0e-6; o->atol = 1.0e-6; o->rtol_vec = nullptr; o->atol_vec = nullptr;
    o->h0 = 0.0; o->h_min = 0.0; o->h_max = INFINITY; o->max_steps = 10000;
    o->safety_factor = 0.9; o->min_scale = 0.2; o->max_scale = 10.0; o->max_rejects = 100;
}
int deb_define_ode(int32_t, int32_t, const char*, int32_t*) { g_err = "user-defined systems need the GPU library"; return DEB_ERR_UNSUPPORTED; }
int deb_define_sde(int32_t, int32_t, const char*, const char*, const char*, int32_t*) { g_err = "user-defined SDEs need the GPU library"; return DEB_ERR_UNSUPPORTED; }
int deb_solve_sde(const deb_sde_problem* P, deb_result* R) {
    if (orc_solve_sde(P, R, 0) != 0) { g_err = "the oracle rejected the problem"; return DEB_ERR_BAD_ARG; }
    int nr = 0;
    for (int i = 0; i < P->n_eval; i++)
        if (P->t_eval[i] > P->t0 || (i == 0 && P->t_eval[i] == P->t0)) { if (R->t_rows) R->t_rows[nr] = P->t_eval[i]; nr++; }
    R->n_rows = nr;
    return DEB_OK;
}
int deb_solve_heat_mol(const deb_heat_problem* P) {
    if (orc_solve_heat_mol(P, 0) != 0) { g_err = "the oracle rejected the problem"; return DEB_ERR_BAD_ARG; }
    return DEB_OK;
}
int deb_define_ode_sensitivity(int32_t, int32_t, const char*, const char*, const char*, int32_t*) { g_err = "user-defined systems need the GPU library"; return DEB_ERR_UNSUPPORTED; }
int deb_define_event(int32_t, const char*, int32_t*) { g_err = "user-defined events need the GPU library"; return DEB_ERR_UNSUPPORTED; }

int deb_solve_ode(const deb_ode_problem* P, deb_result* R) {
    deb_ode_problem Q = *P;  // the oracle is one trajectory after the other on the host: no device list, no fused statistics
    Q.n_devices = 0;
    Q.layout = DEB_LAYOUT_TRAJ_MAJOR;
    deb_result S = *R;
    S.stats_sums = nullptr;
    S.stats_counts = nullptr;
    if (orc_solve_ode(&Q, &S, 0) != 0) { g_err = "the oracle rejected the problem"; return DEB_ERR_BAD_ARG; }
    // the row plan the library publishes (forward time, sorted input -- what the test programs use)
    std::vector<double> rows;
    if (P->solout == DEB_SOLOUT_EVEN) {
        for (double t = P->t0; t <= P->tf; t += P->even_dt) rows.push_back(t);
    } else if (P->solout == DEB_SOLOUT_T_EVAL) {
        for (int i = 0; i < P->n_eval; i++)
            if (P->t_eval[i] > P->t0 || (i == 0 && P->t_eval[i] == P->t0)) rows.push_back(P->t_eval[i]);
    }
    R->n_rows = (int32_t)rows.size();
    if (R->t_rows) std::memcpy(R->t_rows, rows.data(), sizeof(double) * rows.size());
    R->kernel_ms = 0.f;
    R->total_ms = 0.f;
    R->gpu_launches = 0;
    return DEB_OK;
}
}
