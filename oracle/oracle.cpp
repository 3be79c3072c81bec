// oracle.cpp -- CPU restatement of the reference's explicit-RK hot path.  TEST INFRASTRUCTURE ONLY.
//
// Nothing under differential-equations_b200/ may include, link or call this file.  It exists so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can check and time
// the CUDA path against the reference's algorithm.
//
// PARITY STATUS: "parity unpinned" at the bit level.  The reference is a Rust crate; this image has no
// Rust toolchain, so the crate itself cannot be run here.  The reference's own tests pin this path only
// coarsely (tests/ode/accuracy.rs 1e-3..1e3, tests/ode/interpolation.rs 1e-3, tests/ode/from_fn.rs 1e-3,
// tests/pde/method_of_lines.rs 5e-4 / 1e-12 KATs); tests/test_oracle_golden.py checks this file against
// every one of those.  One output of the real crate exists in the tree -- the run printed in docs/ode.md:108-123
// (DOP853, even(1.0), terminal event): 325 function evaluations, 20 accepted + 2 rejected steps, six rows -- and this
// file reproduces it exactly (tests/golden/reference_docs_output.json).  oracle/crate_pin/ + tools/pin_oracle_against_crate.sh
// produce the bit-level pin on any machine with cargo.  Everything beyond that is a line-by-line restatement, with the
// operation order of the Rust source kept exactly (no FMA contraction: build with -ffp-contract=off; libm pow/sqrt).
//
// Each function cites the reference file:line it follows (paths relative to /root/reference/).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/deb_ensemble.h"
#include "erk_tableau_data.h"
#include "philox_ref.h"

namespace {

// State vector of one trajectory: inline storage like the reference's `[f64; N]` / SVector states (no heap traffic).
struct Vec {
    double v[DEB_MAX_DIM];
    int n = 0;
    Vec() {}
    Vec(const Vec& o) : n(o.n) { for (int i = 0; i < n; i++) v[i] = o.v[i]; }
    Vec& operator=(const Vec& o) { n = o.n; for (int i = 0; i < n; i++) v[i] = o.v[i]; return *this; }
    Vec(int n_, double fill) : n(n_) { for (int i = 0; i < n; i++) v[i] = fill; }
    Vec(const double* b, const double* e) : n((int)(e - b)) { for (int i = 0; i < n; i++) v[i] = b[i]; }
    size_t size() const { return (size_t)n; }
    double& operator[](size_t i) { return v[i]; }
    const double& operator[](size_t i) const { return v[i]; }
    double* data() { return v; }
    const double* data() const { return v; }
    double* begin() { return v; }
    double* end() { return v + n; }
    void assign(int n_, double fill) { n = n_; for (int i = 0; i < n; i++) v[i] = fill; }
};
typedef std::vector<double> BigVec;  // method-of-lines state (2^24 nodes)
typedef void (*rhs_fn)(double t, const double* y, double* dydt, const double* p);

// ---------------------------------------------------------------- systems (tests/ode/systems.rs)
void sys_exponential(double, const double* y, double* d, const double* p) { d[0] = p[0] * y[0]; }             // :12-16
void sys_linear(double, const double* y, double* d, const double* p) { d[0] = p[0] + p[1] * y[0]; }           // :25-29
void sys_harmonic(double, const double* y, double* d, const double* p) { d[0] = y[1]; d[1] = -p[0] * y[0]; }  // :38-43
void sys_logistic(double, const double* y, double* d, const double* p) { d[0] = p[0] * y[0] * (1.0 - y[0] / p[1]); }  // :55-59
void sys_vdp(double, const double* y, double* d, const double* p) {                                            // :70-78
    double y1 = y[0], y2 = y[1];
    d[0] = y2;
    d[1] = p[0] * (1.0 - y1 * y1) * y2 - y1;
}
void sys_lorenz(double, const double* y, double* d, const double* p) {  // :91-101
    double x = y[0], yv = y[1], z = y[2];
    d[0] = p[0] * (yv - x);
    d[1] = x * (p[1] - z) - yv;
    d[2] = x * yv - p[2] * z;
}
void sys_brusselator(double, const double* y, double* d, const double* p) {  // :112-120
    double y1 = y[0], y2 = y[1];
    d[0] = p[0] + y1 * y1 * y2 - (p[1] + 1.0) * y1;
    d[1] = p[1] * y1 - y1 * y1 * y2;
}
void sys_robertson(double, const double* y, double* d, const double*) {  // :161-173
    double y1 = y[0], y2 = y[1], y3 = y[2];
    d[0] = -0.04 * y1 + 1.0e4 * y2 * y3;
    d[1] = 0.04 * y1 - 1.0e4 * y2 * y3 - 3.0e7 * y2 * y2;
    d[2] = 3.0e7 * y2 * y2;
}
struct SysInfo { rhs_fn f; int dim, np; };
bool get_system(int id, SysInfo* s) {
    switch (id) {
        case DEB_SYS_EXPONENTIAL: *s = {sys_exponential, 1, 1}; return true;
        case DEB_SYS_LINEAR: *s = {sys_linear, 1, 2}; return true;
        case DEB_SYS_HARMONIC: *s = {sys_harmonic, 2, 1}; return true;
        case DEB_SYS_LOGISTIC: *s = {sys_logistic, 1, 2}; return true;
        case DEB_SYS_VAN_DER_POL: *s = {sys_vdp, 2, 1}; return true;
        case DEB_SYS_LORENZ: *s = {sys_lorenz, 3, 3}; return true;
        case DEB_SYS_BRUSSELATOR: *s = {sys_brusselator, 2, 2}; return true;
        case DEB_SYS_ROBERTSON: *s = {sys_robertson, 3, 0}; return true;
    }
    return false;
}

// ---------------------------------------------------------------- tableau (src/tableau/mod.rs:39-46)
struct Tableau {
    int order, stages, dense;  // O, S, I
    bool adaptive;     // has an error estimate + step-size controller
    bool dp = false;   // Dormand-Prince family (dormandprince/ordinary.rs) vs generic adaptive family (adaptive/ordinary.rs)
    const double* c;   // [I]
    const double* a;   // [I][I]
    const double* b;   // [S]
    const double* bh;  // [S] or null
    const double* er;  // [S] or null
    const double* bi;  // DP family: [I][I]; adaptive family with a dense-output polynomial: [I][O]; or null
    bool fsal = false; // adaptive family only (adaptive/mod.rs:59-122)
};
const double D5_C[7] = DEB_DOPRI5_C;  const double D5_A[7][7] = DEB_DOPRI5_A;  const double D5_B[7] = DEB_DOPRI5_B;
const double D5_ER[7] = DEB_DOPRI5_ER; const double D5_BI[7][7] = DEB_DOPRI5_BI;
const double D8_C[16] = DEB_DOP853_C; const double D8_A[16][16] = DEB_DOP853_A; const double D8_B[12] = DEB_DOP853_B;
const double D8_BH[12] = DEB_DOP853_BH; const double D8_ER[12] = DEB_DOP853_ER; const double D8_BI[16][16] = DEB_DOP853_BI;
const double RK4_C[4] = DEB_RK4_C; const double RK4_A[4][4] = DEB_RK4_A; const double RK4_B[4] = DEB_RK4_B;
const double T38_C[4] = DEB_THREE_EIGHTHS_C; const double T38_A[4][4] = DEB_THREE_EIGHTHS_A; const double T38_B[4] = DEB_THREE_EIGHTHS_B;
const double MID_C[2] = DEB_MIDPOINT_C; const double MID_A[2][2] = DEB_MIDPOINT_A; const double MID_B[2] = DEB_MIDPOINT_B;
const double HEUN_C[2] = DEB_HEUN_C; const double HEUN_A[2][2] = DEB_HEUN_A; const double HEUN_B[2] = DEB_HEUN_B;
const double RAL_C[2] = DEB_RALSTON_C; const double RAL_A[2][2] = DEB_RALSTON_A; const double RAL_B[2] = DEB_RALSTON_B;
const double SSP_C[3] = DEB_SSP_RK3_C; const double SSP_A[3][3] = DEB_SSP_RK3_A; const double SSP_B[3] = DEB_SSP_RK3_B;
const double F45_C[6] = DEB_RKF45_C; const double F45_A[6][6] = DEB_RKF45_A; const double F45_B[6] = DEB_RKF45_B; const double F45_BH[6] = DEB_RKF45_BH;
const double CK_C[6] = DEB_CASH_KARP_C; const double CK_A[6][6] = DEB_CASH_KARP_A; const double CK_B[6] = DEB_CASH_KARP_B; const double CK_BH[6] = DEB_CASH_KARP_BH;
// Verner pairs, tableau/verner.rs; (O, S, I, fsal) from adaptive/mod.rs:59-122
#define ORC_VERNER(NM, PFX, O, S, I) \
    const double NM##_C[I] = DEB_##PFX##_C; const double NM##_A[I][I] = DEB_##PFX##_A; const double NM##_B[S] = DEB_##PFX##_B; \
    const double NM##_BH[S] = DEB_##PFX##_BH; const double NM##_BI[I][O] = DEB_##PFX##_BI;
ORC_VERNER(V655, RKV655E, 6, 9, 10)
ORC_VERNER(V656, RKV656E, 6, 9, 12)
ORC_VERNER(V766, RKV766E, 7, 10, 13)
ORC_VERNER(V767, RKV767E, 7, 10, 16)
ORC_VERNER(V877, RKV877E, 8, 13, 17)
ORC_VERNER(V878, RKV878E, 8, 13, 21)
ORC_VERNER(V988, RKV988E, 9, 16, 21)
ORC_VERNER(V989, RKV989E, 9, 16, 26)
const double EU_C[1] = DEB_EULER_C; const double EU_A[1][1] = DEB_EULER_A; const double EU_B[1] = DEB_EULER_B;

bool get_tableau(int m, Tableau* t) {
    switch (m) {  // (order, S, I): dormandprince/mod.rs:45-58, fixed/mod.rs:41-89
        case DEB_DOPRI5: *t = {5, 7, 7, true, true, D5_C, &D5_A[0][0], D5_B, nullptr, D5_ER, &D5_BI[0][0]}; return true;
        case DEB_DOP853: *t = {8, 12, 16, true, true, D8_C, &D8_A[0][0], D8_B, D8_BH, D8_ER, &D8_BI[0][0]}; return true;
        // adaptive family, adaptive/mod.rs:47-60: (order, S, I) = (5, 6, 6), fsal = false, bi = None
        case DEB_RKF45: *t = {5, 6, 6, true, false, F45_C, &F45_A[0][0], F45_B, F45_BH, nullptr, nullptr}; return true;
        case DEB_CASH_KARP: *t = {5, 6, 6, true, false, CK_C, &CK_A[0][0], CK_B, CK_BH, nullptr, nullptr}; return true;
#define ORC_VERNER_CASE(ID, NM, O, S, I, FSAL) \
        case ID: *t = {O, S, I, true, false, NM##_C, &NM##_A[0][0], NM##_B, NM##_BH, nullptr, &NM##_BI[0][0], FSAL}; return true;
        ORC_VERNER_CASE(DEB_RKV655E, V655, 6, 9, 10, true)
        ORC_VERNER_CASE(DEB_RKV656E, V656, 6, 9, 12, true)
        ORC_VERNER_CASE(DEB_RKV766E, V766, 7, 10, 13, false)
        ORC_VERNER_CASE(DEB_RKV767E, V767, 7, 10, 16, false)
        ORC_VERNER_CASE(DEB_RKV877E, V877, 8, 13, 17, false)
        ORC_VERNER_CASE(DEB_RKV878E, V878, 8, 13, 21, false)
        ORC_VERNER_CASE(DEB_RKV988E, V988, 9, 16, 21, false)
        ORC_VERNER_CASE(DEB_RKV989E, V989, 9, 16, 26, false)
        case DEB_RK4: *t = {4, 4, 4, false, false, RK4_C, &RK4_A[0][0], RK4_B, nullptr, nullptr, nullptr}; return true;
        case DEB_THREE_EIGHTHS: *t = {4, 4, 4, false, false, T38_C, &T38_A[0][0], T38_B, nullptr, nullptr, nullptr}; return true;
        case DEB_MIDPOINT: *t = {2, 2, 2, false, false, MID_C, &MID_A[0][0], MID_B, nullptr, nullptr, nullptr}; return true;
        case DEB_HEUN: *t = {2, 2, 2, false, false, HEUN_C, &HEUN_A[0][0], HEUN_B, nullptr, nullptr, nullptr}; return true;
        case DEB_RALSTON: *t = {2, 2, 2, false, false, RAL_C, &RAL_A[0][0], RAL_B, nullptr, nullptr, nullptr}; return true;
        case DEB_SSP_RK3: *t = {3, 3, 3, false, false, SSP_C, &SSP_A[0][0], SSP_B, nullptr, nullptr, nullptr}; return true;
        case DEB_EULER: *t = {1, 1, 1, false, false, EU_C, &EU_A[0][0], EU_B, nullptr, nullptr, nullptr}; return true;
    }
    return false;
}

// ---------------------------------------------------------------- State ops (src/traits.rs)
inline void add_scaled(Vec& s, double alpha, const Vec& o) {  // traits.rs:316-319  *val += alpha * other[i]
    for (size_t i = 0; i < s.size(); i++) s[i] += alpha * o[i];
}
inline Vec plus_scaled(const Vec& s, double alpha, const Vec& o) {  // traits.rs:336-340
    Vec out = s;
    add_scaled(out, alpha, o);
    return out;
}
inline Vec minus(const Vec& s, const Vec& o) { return plus_scaled(s, -1.0, o); }  // traits.rs:352-354
inline void scale_by(Vec& s, double alpha) { for (double& v : s) v *= alpha; }     // traits.rs:323-326
inline double diff_norm_squared(const Vec& a, const Vec& b) {                      // traits.rs:383-391
    double sum = 0.0;
    for (size_t i = 0; i < a.size(); i++) { double d = a[i] - b[i]; sum += d * d; }
    return sum;
}
// Rust f64::max/min ignore a NaN operand; std::fmax/fmin have the same rule.
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }
inline double signum(double x) { return std::isnan(x) ? x : std::copysign(1.0, x); }  // f64::signum
struct Tol { double s; const double* v; double operator[](size_t i) const { return v ? v[i] : s; } };  // tolerance.rs:32-41
inline double error_norm(const Vec& y, const Vec& y_new, const Vec& err, const Tol& atol, const Tol& rtol) {  // traits.rs:394-410
    double sum = 0.0;
    for (size_t i = 0; i < y.size(); i++) {
        double sk = atol[i] + rtol[i] * rmax(std::fabs(y[i]), std::fabs(y_new[i]));
        double e = err[i] / sk;
        sum += e * e;
    }
    return sum;
}

inline double error_norm_inf(const Vec& y, const Vec& y_new, const Vec& err, const Tol& atol, const Tol& rtol) {  // traits.rs:413-434
    double mx = 0.0;
    for (size_t i = 0; i < y.size(); i++) {
        double sk = atol[i] + rtol[i] * rmax(std::fabs(y[i]), std::fabs(y_new[i]));
        mx = rmax(mx, std::fabs(err[i] / sk));
    }
    return mx;
}

// ---------------------------------------------------------------- step-size utilities (src/utils.rs)
inline double constrain_step_size(double h, double h_min, double h_max) {  // utils.rs:22-33
    double sign = signum(h);
    if (std::fabs(h) < h_min) return sign * h_min;
    if (std::fabs(h) > h_max) return sign * h_max;
    return h;
}
inline bool validate_step_size_parameters(double h0, double h_min, double h_max, double t0, double tf) {  // utils.rs:60-157
    if (tf == t0) return false;
    double sign = signum(tf - t0);
    if (signum(h0) != sign) return false;
    if (h_min < 0.0) return false;
    if (h_max < 0.0) return false;
    if (h_min > h_max) return false;
    if (std::fabs(h0) < h_min) return false;
    if (std::fabs(h0) > h_max) return false;
    if (std::fabs(h0) > std::fabs(tf - t0)) return false;
    if (h0 == 0.0) return false;
    return true;
}

struct Problem {
    rhs_fn f; const double* p; int n;
    void diff(double t, const Vec& y, Vec& d) const { f(t, y.data(), d.data(), p); }
};

// InitialStepSize::<Ordinary>::compute, src/methods/h_init.rs:45-135
double h_init(const Problem& ode, double t0, double tf, const Vec& y0, int order, const Tol& rtol, const Tol& atol,
              double h_min, double h_max, int* evals) {
    double posneg = signum(tf - t0);
    int n = ode.n;
    Vec f0(n, 0.0), f1(n, 0.0);
    ode.diff(t0, y0, f0);
    *evals += 1;
    double dnf = 0.0, dny = 0.0;
    Vec sk_vec(n, 0.0);
    for (int i = 0; i < n; i++) {
        double sk = atol[i] + rtol[i] * std::fabs(y0[i]);
        sk_vec[i] = sk;
        double a = f0[i] / sk; dnf += a * a;   // powi(2)
        double b = y0[i] / sk; dny += b * b;
    }
    double h;
    if (dnf <= 1.0e-10 || dny <= 1.0e-10) h = 1.0e-6;
    else h = std::sqrt(dny / dnf) * 0.01;
    h = rmin(h, h_max);
    h *= posneg;
    Vec y1 = plus_scaled(y0, h, f0);
    ode.diff(t0 + h, y1, f1);
    *evals += 1;
    double der2 = 0.0;
    for (int i = 0; i < n; i++) { double d = (f1[i] - f0[i]) / sk_vec[i]; der2 += d * d; }
    der2 = std::sqrt(der2) / std::fabs(h);
    double der12 = rmax(std::sqrt(dnf), der2);
    double h1;
    if (der12 <= 1.0e-15) h1 = std::fabs(h) * rmax(1.0e-3, 1.0e-6);  // precedence as written, h_init.rs:116-120
    else h1 = std::pow(0.01 / der12, 1.0 / (double)order);
    double interval = std::fabs(tf - t0);
    h = rmin(rmax(rmin(rmin(std::fabs(h) * 100.0, h1), h_max), h_min), interval);
    return h * posneg;
}

enum StepOutcome { STEP_OK = 0, STEP_ERR_MAX_STEPS, STEP_ERR_STEP_SIZE, STEP_ERR_STIFFNESS };

// ExplicitRungeKutta<Ordinary, DormandPrince|Fixed> (src/methods/erk/mod.rs:32-110)
struct Erk {
    Tableau tb;
    Tol rtol, atol;
    double h0, h_min, h_max, safety, min_scale, max_scale;
    int64_t max_steps;
    // state
    double t = 0, h = 0, t_prev = 0, h_prev = 0;
    Vec y, dydt, y_prev, dydt_prev;
    std::vector<Vec> k, cont;
    int64_t steps = 0;
    int stiffness_counter = 0, non_stiffness_counter = 0;
    bool rejected = false;  // Status::RejectedStep
    // the `filter: fn(T) -> T` hook (erk/mod.rs:85,225): identity, or keep the leading mantissa bits (deb_filter)
    uint64_t filter_mask = 0;
    double filter(double v) const {
        if (!filter_mask) return v;
        uint64_t b;
        std::memcpy(&b, &v, 8);
        b &= filter_mask;
        std::memcpy(&v, &b, 8);
        return v;
    }

    // ---- Dormand-Prince: init, dormandprince/ordinary.rs:16-61
    bool dp_init(const Problem& ode, double t0, double tf, const Vec& y0, int* evals) {
        if (h0 == 0.0) h0 = h_init(ode, t0, tf, y0, tb.order, rtol, atol, h_min, h_max, evals);
        if (!validate_step_size_parameters(h0, h_min, h_max, t0, tf)) return false;
        h = filter(h0);  // ordinary.rs:33 / adaptive/ordinary.rs:34
        stiffness_counter = 0;
        t = t0; y = y0;
        int n = ode.n;
        dydt.assign(n, 0.0); y_prev = y0; dydt_prev.assign(n, 0.0);
        k.assign(tb.dense, Vec(n, 0.0));
        cont.assign(tb.order, Vec(n, 0.0));
        ode.diff(t, y, k[0]);
        dydt = k[0];
        *evals += 1;
        t_prev = t; y_prev = y; dydt_prev = dydt;
        rejected = false;
        return true;
    }

    // ---- Dormand-Prince: step, dormandprince/ordinary.rs:63-270
    StepOutcome dp_step(const Problem& ode, int* evals_out) {
        int evals = 0;
        const int S = tb.stages, I = tb.dense, n = ode.n;
        if (std::fabs(h) < std::fabs(h_prev) * 1e-14) return STEP_ERR_STEP_SIZE;  // :70
        if (steps >= max_steps) return STEP_ERR_MAX_STEPS;                        // :82
        steps += 1;
        Vec y_stage(n, 0.0);
        for (int i = 1; i < S; i++) {  // :95-104
            y_stage = y;
            for (int j = 0; j < i; j++) add_scaled(y_stage, tb.a[i * I + j] * h, k[j]);
            ode.diff(t + tb.c[i] * h, y_stage, k[i]);
        }
        Vec ysti = y_stage;
        Vec yseg(n, 0.0);
        for (int i = 0; i < S; i++) add_scaled(yseg, tb.b[i], k[i]);  // :110-113
        Vec y_new = plus_scaled(y, h, yseg);                          // :116
        double t_new = t + h;
        evals += S - 1;
        double err2 = 0.0;
        Vec err_state(n, 0.0);
        for (int j = 0; j < S; j++) add_scaled(err_state, tb.er[j], k[j]);  // :128-130
        double err = error_norm(y, y_new, err_state, atol, rtol);
        if (tb.bh) {  // :135-143
            Vec err2_state = yseg;
            for (int j = 0; j < S; j++) add_scaled(err2_state, -tb.bh[j], k[j]);
            err2 = error_norm(y, y_new, err2_state, atol, rtol);
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = std::fabs(h) * err * std::sqrt(1.0 / (deno * (double)n));  // :148
        double order = (double)tb.order;
        double error_exponent = 1.0 / order;
        double scale = safety * std::pow(err, -error_exponent);  // :154
        scale = rmin(rmax(scale, min_scale), max_scale);          // :157
        if (err <= 1.0) {
            ode.diff(t_new, y_new, dydt);
            evals += 1;
            if (steps % 100 == 0) {  // :165-194
                double stdnum = diff_norm_squared(yseg, k[S - 1]);
                double stden = diff_norm_squared(dydt, ysti);
                if (stden > 0.0) {
                    double h_lamb = h * std::sqrt(stdnum / stden);
                    if (h_lamb > 6.1) {
                        non_stiffness_counter = 0;
                        stiffness_counter += 1;
                        if (stiffness_counter == 15) return STEP_ERR_STIFFNESS;
                    }
                } else {
                    non_stiffness_counter += 1;
                    if (non_stiffness_counter == 6) stiffness_counter = 0;
                }
            }
            cont[0] = y;  // :196-207
            Vec ydiff = minus(y_new, y);
            cont[1] = ydiff;
            Vec bspl(n, 0.0);
            add_scaled(bspl, h, k[0]);
            add_scaled(bspl, -1.0, ydiff);
            cont[2] = bspl;
            Vec cont3 = ydiff;
            add_scaled(cont3, -h, dydt);
            add_scaled(cont3, -1.0, bspl);
            cont[3] = cont3;
            if (tb.bi) {  // :210-235
                if (I > S) {
                    k[S] = dydt;
                    for (int i = S + 1; i < I; i++) {
                        Vec ys = y;
                        for (int j = 0; j < i; j++) add_scaled(ys, tb.a[i * I + j] * h, k[j]);
                        ode.diff(t + tb.c[i] * h, ys, k[i]);
                        evals += 1;
                    }
                }
                for (int i = 4; i < tb.order; i++) {
                    std::fill(cont[i].begin(), cont[i].end(), 0.0);
                    for (int j = 0; j < I; j++) add_scaled(cont[i], tb.bi[i * I + j], k[j]);
                    scale_by(cont[i], h);
                }
            }
            t_prev = t; y_prev = y; dydt_prev = k[0]; h_prev = h;  // :238-241
            t = t_new; y = y_new; k[0] = dydt;
            if (rejected) { rejected = false; scale = rmin(scale, 1.0); }  // :249-254
        } else {
            rejected = true;
        }
        h *= scale;                                   // :261
        h = constrain_step_size(h, h_min, h_max);     // :264
        h = filter(h);                                // :267
        *evals_out += evals;
        return STEP_OK;
    }

    // ---- Dormand-Prince dense output, dormandprince/ordinary.rs:301-337 (factor order AS WRITTEN)
    Vec dp_interpolate(double ti) const {
        double s = (ti - t_prev) / h_prev;
        double s1 = 1.0 - s;
        int ilast = (int)cont.size() - 1;
        Vec acc = cont[ilast];
        for (int i = ilast - 1; i >= 1; i--) {
            double factor;
            if (i >= 4) factor = ((ilast - i) % 2 == 1) ? s1 : s;
            else factor = (i % 2 == 1) ? s1 : s;
            scale_by(acc, factor);
            add_scaled(acc, 1.0, cont[i]);
        }
        return plus_scaled(cont[0], s, acc);
    }

    // ---- adaptive family (RKF45, Cash-Karp): init, adaptive/ordinary.rs:16-61
    int64_t max_rejects = 100;
    bool ad_init(const Problem& ode, double t0, double tf, const Vec& y0, int* evals) {
        if (h0 == 0.0) {
            h0 = h_init(ode, t0, tf, y0, tb.order, rtol, atol, h_min, h_max, evals);
            *evals += 2;  // (sic) counted twice: compute() already added its two evaluations (:24-29)
        }
        if (!validate_step_size_parameters(h0, h_min, h_max, t0, tf)) return false;
        h = filter(h0);  // ordinary.rs:33 / adaptive/ordinary.rs:34
        stiffness_counter = 0;
        t = t0; y = y0;
        int n = ode.n;
        dydt.assign(n, 0.0); y_prev = y0; dydt_prev.assign(n, 0.0);
        k.assign(tb.dense, Vec(n, 0.0));
        ode.diff(t, y, dydt);
        *evals += 1;
        t_prev = t; y_prev = y; dydt_prev = dydt;
        rejected = false;
        return true;
    }
    // ---- adaptive family: step, adaptive/ordinary.rs:63-211 (bi = None, fsal = false for RKF45 / Cash-Karp; Verner pairs have bi)
    StepOutcome ad_step(const Problem& ode, int* evals_out) {
        int evals = 0;
        const int S = tb.stages, I = tb.dense;
        if (std::fabs(h) < std::fabs(h_prev) * 1e-14) return STEP_ERR_STEP_SIZE;
        if (steps >= max_steps) return STEP_ERR_MAX_STEPS;
        steps += 1;
        k[0] = dydt;
        for (int i = 1; i < S; i++) {
            Vec ys = y;
            for (int j = 0; j < i; j++) add_scaled(ys, tb.a[i * I + j] * h, k[j]);
            ode.diff(t + tb.c[i] * h, ys, k[i]);
        }
        evals += S - 1;
        Vec y_high = y;
        for (int i = 0; i < S; i++) add_scaled(y_high, tb.b[i] * h, k[i]);
        Vec y_low = y;
        for (int i = 0; i < S; i++) add_scaled(y_low, tb.bh[i] * h, k[i]);
        Vec err = minus(y_high, y_low);
        double err_norm = error_norm_inf(y, y_high, err, atol, rtol);
        double scale = safety * std::pow(err_norm, -(1.0 / (double)tb.order));
        scale = rmin(rmax(scale, min_scale), max_scale);
        if (err_norm <= 1.0) {
            t_prev = t; y_prev = y; dydt_prev = k[0]; h_prev = h;
            if (rejected) { stiffness_counter = 0; rejected = false; scale = rmin(scale, 1.0); }
            if (tb.bi) {  // extra stages for the dense-output polynomial, :145-160 (on every accepted step)
                for (int i = 0; i < I - S; i++) {
                    Vec ys = y;
                    for (int j = 0; j < S + i; j++) add_scaled(ys, tb.a[(S + i) * I + j] * h, k[j]);
                    ode.diff(t + tb.c[S + i] * h, ys, k[S + i]);
                }
                evals += I - S;
            }
            t += h;
            y = y_high;
            if (tb.fsal) {  // :166-169
                dydt = k[S - 1];
            } else {
                ode.diff(t, y, dydt);
                evals += 1;
            }
        } else {
            rejected = true;
            stiffness_counter += 1;
            if (stiffness_counter >= max_rejects) return STEP_ERR_STIFFNESS;  // Err: this attempt's evals are dropped
        }
        h *= scale;
        h = constrain_step_size(h, h_min, h_max);
        h = filter(h);  // adaptive/ordinary.rs:207
        *evals_out += evals;
        return STEP_OK;
    }

    // dense-output polynomial of the adaptive family, adaptive/ordinary.rs:246-277: Horner in s over bi[i][0..O-1], times s
    Vec ad_interpolate(double ti) const {
        const int I = tb.dense, O = tb.order;
        double s = (ti - t_prev) / h_prev;
        Vec out = y_prev;
        double cont[32];
        for (int i = 0; i < I; i++) {
            cont[i] = tb.bi[i * O + (O - 1)];
            for (int j = O - 2; j >= 0; j--) cont[i] = cont[i] * s + tb.bi[i * O + j];
            cont[i] *= s;
        }
        for (int i = 0; i < I; i++) add_scaled(out, cont[i] * h_prev, k[i]);
        return out;
    }

    // ---- fixed step: init, fixed/ordinary.rs:16-56
    bool fx_init(const Problem& ode, double t0, double tf, const Vec& y0, int* evals) {
        if (h0 == 0.0) h0 = std::fabs(tf - t0) / 100.0;
        if (!validate_step_size_parameters(h0, h_min, h_max, t0, tf)) return false;
        h = h0;
        t = t0; y = y0;
        int n = ode.n;
        dydt.assign(n, 0.0); y_prev = y0; dydt_prev.assign(n, 0.0);
        k.assign(tb.dense, Vec(n, 0.0));
        ode.diff(t, y, dydt);
        *evals += 1;
        t_prev = t; y_prev = y; dydt_prev = dydt;
        return true;
    }
    // ---- fixed step: step, fixed/ordinary.rs:58-139 (fsal = false for every constructor, fixed/mod.rs:41-89)
    StepOutcome fx_step(const Problem& ode, int* evals_out) {
        const int S = tb.stages, I = tb.dense;
        if (steps >= max_steps) return STEP_ERR_MAX_STEPS;
        steps += 1;
        k[0] = dydt;
        for (int i = 1; i < S; i++) {
            Vec ys = y;
            for (int j = 0; j < i; j++) add_scaled(ys, tb.a[i * I + j] * h, k[j]);
            ode.diff(t + tb.c[i] * h, ys, k[i]);
        }
        *evals_out += S - 1;
        t_prev = t; y_prev = y; dydt_prev = k[0]; h_prev = h;
        Vec y_next = y;
        for (int i = 0; i < S; i++) add_scaled(y_next, tb.b[i] * h, k[i]);  // :98-102  (b_i*h)*k_i
        t += h;
        y = y_next;
        ode.diff(t, y, dydt);
        *evals_out += 1;
        return STEP_OK;
    }
    // cubic Hermite, src/interpolate.rs:40-60 via fixed/ordinary.rs:206-216
    Vec fx_interpolate(double ti) const {
        double hh = t - t_prev;
        double s = (ti - t_prev) / hh;
        double s2 = s * s, s3 = s2 * s;
        double h00 = 2.0 * s3 - 3.0 * s2 + 1.0;
        double h10 = s3 - 2.0 * s2 + s;
        double h01 = -2.0 * s3 + 3.0 * s2;
        double h11 = s3 - s2;
        Vec out(y.size(), 0.0);  // linear_combination: fill(0) then add_scaled each term, traits.rs:357-370
        add_scaled(out, h00, y_prev);
        add_scaled(out, h10 * hh, dydt_prev);
        add_scaled(out, h01, y);
        add_scaled(out, h11 * hh, dydt);
        return out;
    }
};

// TEvalSolout, src/solout/t_eval.rs:87-171
struct TEval {
    BigVec pts; size_t idx = 0; double dir;
    TEval(const double* p, int n, double t0, double tf) : pts(p, p + n), dir(signum(tf - t0)) {
        if (dir > 0.0) std::stable_sort(pts.begin(), pts.end(), [](double a, double b) { return a < b; });
        else std::stable_sort(pts.begin(), pts.end(), [](double a, double b) { return a > b; });
    }
};

struct Out {  // per-trajectory output slots
    double* y_eval; int32_t* n_emitted; double* t_final; double* y_final;
    int32_t* status; int32_t* accepted; int32_t* rejected; int32_t* evals;
    double* t_out = nullptr;
};

// Row sink = Solution.t / Solution.y of one trajectory: `cap` row slots; pushes beyond that are counted only.
struct Rows {
    double* y_eval; double* t_out; int n; int cap;
    int n_emit = 0;
    double last_t = 0.0;  // solution.t.last()
    void push(double t, const Vec& v) {
        if (n_emit < cap) {
            if (y_eval) std::memcpy(y_eval + (size_t)n_emit * n, v.data(), sizeof(double) * n);
            if (t_out) t_out[n_emit] = t;
        }
        n_emit += 1;
        last_t = t;
    }
    void pop() { n_emit -= 1; }
};

// Per-step recorders: DefaultSolout (src/solout/default.rs:54-75), DenseSolout (src/solout/dense.rs:74-108),
// CrossingSolout (src/solout/crossing.rs:115-263).  Rows carry their own time; `cap` row slots per trajectory, pushes
// beyond that are counted only.  Where the reference's interpolate() would return Err(OutOfBounds) and the recorder
// would panic on unwrap (backward time; the Newton probe just past the step end) the polynomial is evaluated.
struct PerStep {
    int mode = 0, dense_n = 0, comp = 0, direction = 0, cap = 0;
    double threshold = 0.0;
    bool have_last = false;
    double last_off = 0.0;
    // HyperplaneCrossingSolout (src/solout/hyperplane.rs): extractor = component selection, normal normalised in new()
    int plane_dim = 0;
    int plane_index[DEB_MAX_DIM];
    double plane_point[DEB_MAX_DIM], plane_normal[DEB_MAX_DIM];
    double distance(const Vec& y) const {  // signed_distance, hyperplane.rs:160-162: pos.minus(point).dot(normal)
        double sum = 0.0;
        for (int i = 0; i < plane_dim; i++) {
            const double d = y[plane_index[i]] + (-1.0) * plane_point[i];
            sum = sum + d * plane_normal[i];
        }
        return sum;
    }
};
// find_crossing_newton of the hyperplane recorder, hyperplane.rs:262-330
template <class Interp>
bool plane_newton(const PerStep& ps, Interp&& interp, double t_lower, double t_upper, double dist_lower, double dist_upper, double* t_found) {
    double t = t_lower - dist_lower * (t_upper - t_lower) / (dist_upper - dist_lower);
    const double tolerance = DBL_EPSILON * 100.0;
    double dist;
    for (int it = 0; it < 10; it++) {
        dist = ps.distance(interp(t));
        if (std::fabs(dist) < tolerance) { *t_found = t; return true; }
        const double delta_t = (t_upper - t_lower) * 1e-6;
        const double dist_plus = ps.distance(interp(t + delta_t));
        const double derivative = (dist_plus - dist) / delta_t;
        if (std::fabs(derivative) < DBL_EPSILON) break;
        const double t_new = t - dist / derivative;
        if (t_new < t_lower || t_new > t_upper) t = (t_lower + t_upper) / 2.0;
        else t = t_new;
    }
    dist = ps.distance(interp(t));
    *t_found = t;
    return std::fabs(dist) < tolerance * 10.0;
}
template <class Interp>
bool crossing_newton(const PerStep& ps, Interp&& interp, double t_lower, double t_upper, double off_lower, double off_upper, double* t_found) {
    double t = t_lower - off_lower * (t_upper - t_lower) / (off_upper - off_lower);
    const double tolerance = DBL_EPSILON * 100.0;
    double off;
    for (int it = 0; it < 10; it++) {
        Vec y_t = interp(t);
        off = y_t[ps.comp] - ps.threshold;
        if (std::fabs(off) < tolerance) { *t_found = t; return true; }
        const double delta_t = (t_upper - t_lower) * 1e-6;
        const double t_plus = t + delta_t;
        Vec y_plus = interp(t_plus);
        const double off_plus = y_plus[ps.comp] - ps.threshold;
        const double derivative = (off_plus - off) / delta_t;
        if (std::fabs(derivative) < DBL_EPSILON * 10.0) break;
        const double t_new = t - off / derivative;
        if (t_new < t_lower || t_new > t_upper) {
            t = (t_lower + t_upper) / 2.0;
        } else {
            const double change = std::fabs(t_new - t);
            if (change < tolerance * 0.1) { t = t_new; break; }
            t = t_new;
        }
    }
    Vec y_t = interp(t);
    off = y_t[ps.comp] - ps.threshold;
    *t_found = t;
    return std::fabs(off) < tolerance * 10.0;
}
template <class Interp>
void solout_per_step(PerStep& ps, double t_curr, double t_prev, const Vec& y_curr, Interp&& interp, Rows& rows) {
    auto push = [&](double t, const Vec& v) { rows.push(t, v); };
    if (ps.mode == DEB_SOLOUT_DEFAULT) {
        push(t_curr, y_curr);
    } else if (ps.mode == DEB_SOLOUT_DENSE) {
        if (t_prev != t_curr) {
            for (int i = 1; i < ps.dense_n; i++) {
                const double h_old = t_curr - t_prev;
                const double ti = t_prev + (double)i * h_old / (double)ps.dense_n;
                push(ti, interp(ti));
            }
        }
        push(t_curr, y_curr);
    } else if (ps.mode == DEB_SOLOUT_HYPERPLANE) {  // hyperplane.rs:182-236
        const double dist = ps.distance(y_curr);
        if (ps.have_last) {
            const double last = ps.last_off;
            const bool is_crossing = signum(last) != signum(dist) || (last == 0.0 && dist != 0.0) || (last != 0.0 && dist == 0.0);
            if (is_crossing) {
                const bool record = ps.direction > 0 ? (last < 0.0 && dist >= 0.0) : ps.direction < 0 ? (last > 0.0 && dist <= 0.0) : true;
                if (record) {
                    double t_cross;
                    if (!plane_newton(ps, interp, t_prev, t_curr, last, dist, &t_cross)) {
                        const double frac = -last / (dist - last);
                        t_cross = t_prev + frac * (t_curr - t_prev);
                    }
                    push(t_cross, interp(t_cross));
                }
            }
        }
        ps.last_off = dist;
        ps.have_last = true;
    } else {  // crossing
        const double off = y_curr[ps.comp] - ps.threshold;
        if (ps.have_last) {
            const double last = ps.last_off;
            const bool is_crossing = signum(last) != signum(off);
            if (is_crossing) {
                const bool record = ps.direction > 0 ? (last < 0.0 && off >= 0.0) : ps.direction < 0 ? (last > 0.0 && off <= 0.0) : true;
                if (record) {
                    double t_cross;
                    if (!crossing_newton(ps, interp, t_prev, t_curr, last, off, &t_cross)) {
                        const double frac = -last / (off - last);
                        t_cross = t_prev + frac * (t_curr - t_prev);
                    }
                    push(t_cross, interp(t_cross));
                }
            }
        }
        ps.last_off = off;
        ps.have_last = true;
    }
}

template <class Interp>
void solout_teval(TEval& te, double t_curr, double t_prev, const Vec& y_curr, Interp&& interp, Rows& rows) {
    size_t idx = te.idx;
    while (idx < te.pts.size()) {
        double tv = te.pts[idx];
        bool in_range = (te.dir > 0.0) ? ((tv == t_prev && idx == 0) || (tv > t_prev && tv <= t_curr))
                                       : ((tv == t_prev && idx == 0) || (tv < t_prev && tv >= t_curr));
        if (in_range) {
            Vec yv = (tv == t_curr) ? y_curr : interp(tv);
            rows.push(tv, yv);
            idx += 1;
        } else {
            if ((te.dir > 0.0 && tv > t_curr) || (te.dir < 0.0 && tv < t_curr)) break;
            idx += 1;
        }
    }
    te.idx = idx;
}

// EvenSolout, src/solout/even.rs:60-217 (rows are appended to y_eval; *n_emit counts them, the last one may be replaced)
struct Even {
    double dt, t0, tf, dir;
    bool has_last = false;
    double last = 0.0;
    Even(double dt_, double t0_, double tf_) : dt(dt_), t0(t0_), tf(tf_), dir(signum(tf_ - t0_)) {}
};
template <class Interp>
void solout_even(Even& ev, double t_curr, double t_prev, const Vec& y_curr, const Vec& y_prev, Interp&& interp, Rows& rows) {
    auto push = [&](double t, const Vec& v) { rows.push(t, v); };
    const double offset = std::fmod(ev.t0, ev.dt);
    const double tol = std::fabs(ev.dt) * 1e-12 + DBL_EPSILON * 10.0;
    double start_t;
    if (ev.has_last) {
        start_t = ev.last + ev.dt * ev.dir;
    } else {
        if (std::fabs(t_prev - ev.t0) < DBL_EPSILON) {
            push(ev.t0, y_prev);
            ev.last = ev.t0; ev.has_last = true;
            start_t = ev.t0 + ev.dt * ev.dir;
        } else {
            double rem = std::fmod(t_prev - offset, ev.dt);
            if (ev.dir > 0.0) start_t = (std::fabs(rem) < DBL_EPSILON) ? t_prev : t_prev + (ev.dt - rem);
            else start_t = (std::fabs(rem) < DBL_EPSILON) ? t_prev : t_prev - rem;
        }
    }
    double ti = start_t;
    while ((ev.dir > 0.0 && ti <= t_curr) || (ev.dir < 0.0 && ti >= t_curr)) {
        if ((ev.dir > 0.0 && ti >= t_prev && ti <= t_curr) || (ev.dir < 0.0 && ti <= t_prev && ti >= t_curr)) {
            if (ev.has_last && std::fabs(ti - ev.last) <= tol) {
                // near-duplicate: skip
            } else {
                push(ti, interp(ti));
                ev.last = ti; ev.has_last = true;
            }
        }
        ti += ev.dt * ev.dir;
    }
    if (t_curr == ev.tf) {  // even.rs:166-188
        if (ev.has_last) {
            if (std::fabs(ev.last - ev.tf) <= tol) {
                rows.pop();  // solution.pop()
                push(ev.tf, y_curr);
                ev.last = ev.tf;
            } else if (ev.last != ev.tf) {
                push(ev.tf, y_curr);
                ev.last = ev.tf;
            }
        } else {
            push(ev.tf, y_curr);
            ev.last = ev.tf; ev.has_last = true;
        }
    }
}

// Event detection wrapped around a recorder: EventWrappedSolout, src/solout/event.rs:300-470.  The event function is the
// ABI's linear form g(t, y) = c0 + c1*t + sum c[2+i]*y[i] (arbitrary `impl Event` bodies are covered by tests/py_restatement.py).
struct EventState {
    int direction = 0, terminate = 0, n = 0;
    const double* coef = nullptr;
    double dir = 1.0;
    bool has_last = false;
    double last_g = 0.0;
    int count = 0;
    double g(double t, const Vec& y) const {
        double v = coef[0] + coef[1] * t;
        for (int c = 0; c < n; c++) v = v + coef[2 + c] * y[c];
        return v;
    }
};
// brent_dekker, event.rs:386-450.  interpolate(b).ok()?: the reference's bounds test (written for forward time) ends the search
template <class Interp>
bool brent_dekker(const EventState& es, Interp&& interp, double t_prev, double t_curr, double a, double b, double fa, double fb, double* t_event) {
    const double rel_tol = 1e-12, abs_tol = 1e-14;
    if (std::fabs(fa) < std::fabs(fb)) { std::swap(a, b); std::swap(fa, fb); }
    double c = a, fc = fa, d = b - a, e = d;
    for (int it = 0; it < 50; it++) {
        if (fb == 0.0) { *t_event = b; return true; }
        if (signum(fa) == signum(fb)) { a = c; fa = fc; c = b; fc = fb; d = b - a; e = d; }
        if (std::fabs(fa) < std::fabs(fb)) { c = b; b = a; a = c; fc = fb; fb = fa; fa = fc; }
        const double tol = rmax(abs_tol, rel_tol * std::fabs(b));
        const double m = 0.5 * (a - b);
        if (std::fabs(m) <= tol || fb == 0.0) { *t_event = b; return true; }
        bool use_bis = true;
        if (std::fabs(e) > tol && std::fabs(fa) > std::fabs(fb)) {
            const double s = fb / fa;
            double p, q;
            if (a == c) {
                p = 2.0 * m * s;
                q = 1.0 - s;
            } else {
                const double q1 = fa / fc;
                const double r = fb / fc;
                p = s * (2.0 * m * q1 * (q1 - r) - (b - a) * (r - 1.0));
                q = (q1 - 1.0) * (r - 1.0) * (s - 1.0);
            }
            double q_mod = q, p_mod = p;
            if (q_mod > 0.0) p_mod = -p_mod; else q_mod = -q_mod;
            if (std::fabs(2.0 * p_mod) < (3.0 * m * q_mod - std::fabs(tol * q_mod)) && p_mod < std::fabs(e * 0.5 * q_mod)) {
                e = d;
                d = p_mod / q_mod;
                use_bis = false;
            }
        }
        if (use_bis) { d = m; e = m; }
        a = b;
        fa = fb;
        b = (std::fabs(d) > tol) ? (b + d) : (b + ((m > 0.0) ? tol : -tol));
        if (b < t_prev || b > t_curr) return false;  // Err(OutOfBounds).ok()? -> None
        Vec yb = interp(b);
        fb = es.g(b, yb);
        c = a;
        fc = fa;
    }
    return false;
}
// detect_event, event.rs:318-384; returns true for ControlFlag::Terminate
template <class Interp>
bool detect_event(EventState& es, double t_curr, double t_prev, const Vec& y_curr, const Vec& y_prev, Interp&& interp, Rows& rows) {
    const double g_curr = es.g(t_curr, y_curr);
    if (!es.has_last) {
        (void)es.g(t_prev, y_prev);
        es.last_g = g_curr; es.has_last = true;
        return false;
    }
    const double g_prev = es.last_g;
    const bool sign_change = signum(g_prev) != signum(g_curr);
    const bool direction_ok = es.direction > 0 ? (sign_change && g_prev < 0.0 && g_curr >= 0.0)
                            : es.direction < 0 ? (sign_change && g_prev > 0.0 && g_curr <= 0.0) : sign_change;
    if (direction_ok) {
        double a = t_prev, b = t_curr, fa = g_prev, fb = g_curr;
        if ((es.dir > 0.0 && a > b) || (es.dir < 0.0 && a < b)) { std::swap(a, b); std::swap(fa, fb); }
        double t_event;
        if (fa * fb <= 0.0 && brent_dekker(es, interp, t_prev, t_curr, a, b, fa, fb, &t_event)) {
            Vec y_event = interp(t_event);
            const bool push_point = rows.n_emit > 0 ? (std::fabs(t_event - rows.last_t) > 1e-14) : true;
            if (push_point) rows.push(t_event, y_event);
            es.count += 1;
            if (es.terminate > 0 && es.count >= es.terminate) { es.last_g = g_curr; return true; }
        }
    }
    es.last_g = g_curr;
    return false;
}

// solve_ode, src/ode/solve_ivp.rs:116-277, for one trajectory
void solve_one(const deb_ode_problem* P, const SysInfo& si, const Tableau& tb, int64_t i, const Out& o) {
    const int n = si.dim;
    Problem ode{si.f, P->params_shared ? P->params : P->params + (size_t)i * si.np, n};
    Vec y0(P->y0 + (size_t)i * n, P->y0 + (size_t)(i + 1) * n);
    const double t0 = P->t0, tf = P->tf;
    Erk m;
    m.tb = tb;
    m.rtol = {P->opt.rtol, P->opt.rtol_vec}; m.atol = {P->opt.atol, P->opt.atol_vec};
    m.h0 = P->opt.h0; m.h_min = P->opt.h_min; m.h_max = P->opt.h_max; m.max_steps = P->opt.max_steps;
    m.safety = P->opt.safety_factor; m.min_scale = P->opt.min_scale; m.max_scale = P->opt.max_scale;
    m.max_rejects = P->opt.max_rejects;
    if (P->filter == DEB_FILTER_TRUNCATE_MANTISSA) m.filter_mask = ~((1ull << (52 - P->filter_bits)) - 1ull);
    int evals = 0, acc = 0, rej = 0;
    int status = DEB_STATUS_COMPLETE;
    const bool has_event = (P->event != DEB_EVENT_NONE);
    const int row_cap = (has_event && P->row_capacity > 0) ? P->row_capacity : P->n_eval;
    const bool per_step = (P->solout == DEB_SOLOUT_DEFAULT || P->solout == DEB_SOLOUT_DENSE || P->solout == DEB_SOLOUT_CROSSING ||
                           P->solout == DEB_SOLOUT_HYPERPLANE);
    Rows rows{o.y_eval ? o.y_eval + (size_t)i * row_cap * n : nullptr, o.t_out ? o.t_out + (size_t)i * row_cap : nullptr, n,
              (per_step || has_event) ? row_cap : 0x7fffffff};
    int& n_emit = rows.n_emit;
    auto finish = [&](int st, double t, const Vec& y) {
        if (o.status) o.status[i] = st;
        if (o.t_final) o.t_final[i] = t;
        if (o.y_final) std::memcpy(o.y_final + (size_t)i * n, y.data(), sizeof(double) * n);
        if (o.accepted) o.accepted[i] = acc;
        if (o.rejected) o.rejected[i] = rej;
        if (o.evals) o.evals[i] = evals;
        if (o.n_emitted) o.n_emitted[i] = n_emit;
    };
    double dir = signum(tf - t0);
    if (!(dir == 1.0 || dir == -1.0)) { finish(DEB_STATUS_BAD_INPUT, t0, y0); return; }  // :139-147
    bool ok = tb.dp ? m.dp_init(ode, t0, tf, y0, &evals) : tb.adaptive ? m.ad_init(ode, t0, tf, y0, &evals) : m.fx_init(ode, t0, tf, y0, &evals);
    if (!ok) { evals = 0; finish(DEB_STATUS_BAD_INPUT, t0, y0); return; }
    const bool even = (P->solout == DEB_SOLOUT_EVEN);
    const bool no_points = even || per_step;
    TEval te(no_points ? nullptr : P->t_eval, no_points ? 0 : P->n_eval, t0, tf);
    Even ev(P->even_dt, t0, tf);
    // adaptive family without bi: cubic Hermite on (t_prev, t, y_prev, y, dydt_prev, dydt), adaptive/ordinary.rs:282-295
    auto interp = [&](double tv) { return tb.dp ? m.dp_interpolate(tv) : (tb.adaptive && tb.bi) ? m.ad_interpolate(tv) : m.fx_interpolate(tv); };
    PerStep ps;
    ps.mode = P->solout; ps.dense_n = P->dense_n; ps.comp = P->cross_component; ps.direction = P->cross_direction;
    ps.threshold = P->cross_threshold;
    if (P->solout == DEB_SOLOUT_HYPERPLANE) {  // HyperplaneCrossingSolout::new, hyperplane.rs:124-139
        ps.plane_dim = P->plane_dim;
        double nsq = 0.0;
        for (int q = 0; q < P->plane_dim; q++) nsq = nsq + P->plane_normal[q] * P->plane_normal[q];
        const double norm = std::sqrt(nsq);
        for (int q = 0; q < P->plane_dim; q++) {
            ps.plane_index[q] = P->plane_index[q];
            ps.plane_point[q] = P->plane_point[q];
            ps.plane_normal[q] = (norm > DBL_EPSILON) ? P->plane_normal[q] * (1.0 / norm) : P->plane_normal[q];
        }
    }
    EventState es;
    es.direction = P->event_direction; es.terminate = P->event_terminate; es.coef = P->event_coef; es.n = n; es.dir = dir;
    // EventWrappedSolout::solout, event.rs:452-470: the base recorder first, then event detection; true = Terminate
    auto record = [&]() -> bool {
        if (per_step) solout_per_step(ps, m.t, m.t_prev, m.y, interp, rows);
        else if (even) solout_even(ev, m.t, m.t_prev, m.y, m.y_prev, interp, rows);
        else solout_teval(te, m.t, m.t_prev, m.y, interp, rows);
        if (has_event) return detect_event(es, m.t, m.t_prev, m.y, m.y_prev, interp, rows);
        return false;
    };
    record();  // :160
    const double eps10 = DBL_EPSILON * 10.0;
    for (;;) {
        if ((m.t + m.h - tf) * dir > 0.0) {  // :193-209
            double h_new = tf - m.t;
            if (std::fabs(h_new) < eps10) { status = DEB_STATUS_COMPLETE; break; }
            m.h = tb.adaptive ? m.filter(h_new) : h_new;  // set_h: the adaptive steppers filter (ordinary.rs:288), the fixed one does not
        }
        StepOutcome so = tb.dp ? m.dp_step(ode, &evals) : tb.adaptive ? m.ad_step(ode, &evals) : m.fx_step(ode, &evals);
        if (so != STEP_OK) {
            status = so == STEP_ERR_MAX_STEPS ? DEB_STATUS_MAX_STEPS : so == STEP_ERR_STEP_SIZE ? DEB_STATUS_STEP_SIZE : DEB_STATUS_STIFFNESS;
            break;
        }
        if (m.rejected) { rej += 1; continue; }  // :218-221
        acc += 1;
        if (record()) { status = DEB_STATUS_INTERRUPTED; break; }  // ControlFlag::Terminate, :255-260
        if (std::fabs(tf - m.t) <= eps10) break;  // :263
    }
    finish(status, m.t, m.y);
}

// ---------------------------------------------------------------- SDE (fixed step), built-ins
// drift / diffusion: SDE::drift, SDE::diffusion (src/sde/sde.rs:16-52).  mix: what the system's SDE::noise does with the
// independent increments of the library's Philox stream (identity, or the Heston correlation).
struct SdeSys {
    int dim, np;
    void (*drift)(double, const double*, double*, const double*);
    void (*diffusion)(double, const double*, double*, const double*);
    void (*mix)(double*, const double*);
};
void ou_drift(double, const double* y, double* d, const double* p) { d[0] = p[0] * (p[1] - y[0]); }   // examples/sde/03_ornstein_uhlenbeck/main.rs:43-45
void ou_diff(double, const double*, double* g, const double* p) { g[0] = p[2]; }                      // :47-49
void gbm_drift(double, const double* y, double* d, const double* p) { d[0] = p[0] * y[0]; }           // src/sde/solve_ivp.rs doc example
void gbm_diff(double, const double* y, double* g, const double* p) { g[0] = p[1] * y[0]; }
void no_mix(double*, const double*) {}
// examples/sde/02_heston_model/main.rs:53-72: p = {mu, kappa, theta, sigma, rho}; y = {price, variance}
void heston_drift(double, const double* y, double* d, const double* p) { d[0] = p[0] * y[0]; d[1] = p[1] * (p[2] - y[1]); }
void heston_diff(double, const double* y, double* g, const double* p) { g[0] = y[0] * std::sqrt(y[1]); g[1] = p[3] * std::sqrt(y[1]); }
void heston_mix(double* dw, const double* p) { dw[1] = p[4] * dw[0] + std::sqrt(1.0 - p[4] * p[4]) * dw[1]; }
bool get_sde(int id, SdeSys* s) {
    if (id == DEB_SDE_OU) { *s = {1, 3, ou_drift, ou_diff, no_mix}; return true; }
    if (id == DEB_SDE_GBM) { *s = {1, 2, gbm_drift, gbm_diff, no_mix}; return true; }
    if (id == DEB_SDE_HESTON) { *s = {2, 5, heston_drift, heston_diff, heston_mix}; return true; }
    return false;
}

// solve_sde (src/sde/solve_ivp.rs:135-287) + ExplicitRungeKutta<Stochastic, Fixed> (fixed/stochastic.rs:18-146) / Milstein
// (milstein.rs:107-180), diagonal noise.  SDE::noise is the library's Philox Wiener increment (see philox_ref.h):
// component c of step s is normal number s*dim + c of the path's stream, times sqrt(h); then the system's mix.
void solve_one_sde(const deb_sde_problem* P, const SdeSys& ss, const Tableau& tb, int64_t i, const Out& o) {
    const int n = ss.dim;
    const double* p = P->params_shared ? P->params : P->params + (size_t)i * ss.np;
    Vec y0(n, 0.0);
    for (int c = 0; c < n; c++) y0[c] = P->y0_shared ? P->y0[c] : P->y0[(size_t)i * n + c];
    const double t0 = P->t0, tf = P->tf;
    int evals = 0, acc = 0;
    Rows rows{o.y_eval ? o.y_eval + (size_t)i * P->n_eval * n : nullptr, nullptr, n, 0x7fffffff};
    auto finish = [&](int st, double t, const Vec& y) {
        if (o.status) o.status[i] = st;
        if (o.t_final) o.t_final[i] = t;
        if (o.y_final) std::memcpy(o.y_final + (size_t)i * n, y.data(), sizeof(double) * n);
        if (o.accepted) o.accepted[i] = acc;
        if (o.rejected) o.rejected[i] = 0;
        if (o.evals) o.evals[i] = evals;
        if (o.n_emitted) o.n_emitted[i] = rows.n_emit;
    };
    double dir = signum(tf - t0);
    if (!(dir == 1.0 || dir == -1.0)) { finish(DEB_STATUS_BAD_INPUT, t0, y0); return; }
    // init, stochastic.rs:18-65
    double h0 = P->opt.h0;
    if (h0 == 0.0) h0 = std::fabs(tf - t0) / 100.0;
    if (!validate_step_size_parameters(h0, P->opt.h_min, P->opt.h_max, t0, tf)) { finish(DEB_STATUS_BAD_INPUT, t0, y0); return; }
    double h = h0, t = t0;
    Vec y = y0, dydt(n, 0.0), g(n, 0.0);
    int64_t steps = 0;
    const int S = tb.stages, I = tb.dense;
    const bool milstein = (P->method == DEB_MILSTEIN);
    std::vector<Vec> k(8, Vec(n, 0.0));
    ss.drift(t, y.data(), dydt.data(), p);
    if (milstein) {
        evals += 1;  // Milstein::init evaluates the drift only, milstein.rs:88-100
    } else {
        ss.diffusion(t, y.data(), g.data(), p);
        evals += 2;
    }
    double t_prev = t;
    Vec y_prev = y;
    TEval te(P->t_eval, P->n_eval, t0, tf);
    auto emit = [&](double t_curr, double tp, const Vec& y_curr) {
        // linear interpolation, src/interpolate.rs:71-74 via stochastic.rs:177-190
        auto interp = [&](double tv) {
            double s = (tv - tp) / (t_curr - tp);
            Vec out(n, 0.0);
            add_scaled(out, 1.0 - s, y_prev);
            add_scaled(out, s, y_curr);
            return out;
        };
        solout_teval(te, t_curr, tp, y_curr, interp, rows);
    };
    emit(t, t_prev, y);
    const double eps10 = DBL_EPSILON * 10.0;
    const uint64_t path = (uint64_t)(P->path_offset + i);
    int status = DEB_STATUS_COMPLETE;
    auto noise = [&](Vec& dw) {
        for (int c = 0; c < n; c++) dw[c] = deb_ref::wiener_increment(P->seed, path, (uint64_t)(steps - 1), c, n, h);
        ss.mix(dw.data(), p);
    };
    for (;;) {
        if ((t + h - tf) * dir > 0.0) {
            double h_new = tf - t;
            if (std::fabs(h_new) < eps10) break;
            h = h_new;
        }
        // step, stochastic.rs:67-146
        if (steps >= P->opt.max_steps) { status = DEB_STATUS_MAX_STEPS; break; }
        steps += 1;
        t_prev = t; y_prev = y;
        Vec dw(n, 0.0), y_next = y;
        if (milstein) {  // Milstein::step, milstein.rs:107-180
            ss.diffusion(t, y.data(), g.data(), p);
            evals += 1;
            noise(dw);
            double sqrt_h = std::sqrt(h);
            Vec y_aux = y;
            add_scaled(y_aux, sqrt_h, g);
            Vec g_aux(n, 0.0);
            ss.diffusion(t, y_aux.data(), g_aux.data(), p);
            evals += 1;
            double factor = 1.0 / (2.0 * sqrt_h);
            Vec milstein_term(n, 0.0), drift_inc = dydt, diff_inc(n, 0.0);
            for (int c = 0; c < n; c++) {
                double diff = g_aux[c] - g[c];
                double dws_minus_h = dw[c] * dw[c] - h;
                milstein_term[c] = diff * dws_minus_h * factor;
                drift_inc[c] *= h;
                diff_inc[c] = g[c] * dw[c];
            }
            add_scaled(y_next, 1.0, drift_inc);
            add_scaled(y_next, 1.0, diff_inc);
            add_scaled(y_next, 1.0, milstein_term);
        } else {
            k[0] = dydt;
            for (int s = 1; s < S; s++) {
                Vec ys = y;
                for (int j = 0; j < s; j++) add_scaled(ys, tb.a[s * I + j] * h, k[j]);
                ss.drift(t + tb.c[s] * h, ys.data(), k[s].data(), p);
            }
            evals += S - 1;
            Vec drift_inc(n, 0.0);
            for (int s = 0; s < S; s++) add_scaled(drift_inc, tb.b[s] * h, k[s]);
            ss.diffusion(t, y.data(), g.data(), p);
            evals += 1;
            noise(dw);
            Vec diff_inc(n, 0.0);
            for (int c = 0; c < n; c++) diff_inc[c] = g[c] * dw[c];  // component_multiply, linalg/util.rs:21
            add_scaled(y_next, 1.0, drift_inc);                      // plus_linear_combination, traits.rs:343-349
            add_scaled(y_next, 1.0, diff_inc);
        }
        t += h;
        y = y_next;
        ss.drift(t, y.data(), dydt.data(), p);
        evals += 1;
        acc += 1;
        emit(t, t_prev, y);
        if (std::fabs(tf - t) <= eps10) break;
    }
    finish(status, t, y);
}

template <class F>
void parallel_for(int64_t n, int n_threads, F&& body) {
    if (n_threads <= 1 || n < 2) { for (int64_t i = 0; i < n; i++) body(i); return; }
    std::atomic<int64_t> next{0};
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(256, n / (n_threads * 8) + 1));
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&] {
            for (;;) {
                int64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                int64_t e = std::min(n, b + chunk);
                for (int64_t i = b; i < e; i++) body(i);
            }
        });
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

// Same problem/result structs as the product ABI (host pointers only); n_threads <= 0 means all cores.
int orc_solve_ode(const deb_ode_problem* P, deb_result* R, int n_threads) {
    SysInfo si; Tableau tb;
    if (!P || !R || !get_system(P->system, &si) || !get_tableau(P->method, &tb)) return DEB_ERR_BAD_ARG;
    if (si.dim != P->dim || si.np != P->n_params) return DEB_ERR_BAD_ARG;
    if (n_threads <= 0) n_threads = orc_hardware_threads();
    Out o{R->y_eval, R->n_emitted, R->t_final, R->y_final, R->status, R->accepted, R->rejected, R->evals, R->t_out};
    parallel_for(P->n_traj, n_threads, [&](int64_t i) { solve_one(P, si, tb, i, o); });
    // sorted t_eval bookkeeping for the caller
    return DEB_OK;
}

int orc_solve_sde(const deb_sde_problem* P, deb_result* R, int n_threads) {
    SdeSys ss; Tableau tb;
    const bool milstein = P && P->method == DEB_MILSTEIN;
    if (milstein) get_tableau(DEB_EULER, &tb);
    if (!P || !R || !get_sde(P->system, &ss) || (!milstein && (!get_tableau(P->method, &tb) || tb.adaptive))) return DEB_ERR_BAD_ARG;
    if (P->dim != ss.dim || ss.np != P->n_params) return DEB_ERR_BAD_ARG;
    if (n_threads <= 0) n_threads = orc_hardware_threads();
    Out o{R->y_eval, R->n_emitted, R->t_final, R->y_final, R->status, R->accepted, R->rejected, R->evals};
    parallel_for(P->n_traj, n_threads, [&](int64_t i) { solve_one_sde(P, ss, tb, i, o); });
    return DEB_OK;
}

// Wiener increment / normal exposed for the host-side stream regeneration tests.
double orc_wiener_increment(uint64_t seed, uint64_t path, uint64_t step, int comp, int dim, double h) {
    return deb_ref::wiener_increment(seed, path, step, comp, dim, h);
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { deb_ref::philox4x32_10(ctr, key, out); }

// libm pow over an array (what Rust's f64::powf calls): the truth the device pow port is compared with.
void orc_pow_array(const double* x, double y, int64_t n, double* out) { for (int64_t i = 0; i < n; i++) out[i] = std::pow(x[i], y); }

// Method-of-lines heat equation: SemiDiscretePde::diff (src/pde/semi_discrete.rs:250-287) with the FD flux of
// :181-190 / :150-172 for a scalar field on a 1-D uniform grid, zero source, flux = alpha*grad_u, integrated by
// the fixed-step ERK loop above (solve_ode + fixed/ordinary.rs).  Plain stencil loop, not the reference's generic
// per-node machinery; arithmetic per node is the same expression sequence.
// One evaluation of SemiDiscretePde::diff (semi_discrete.rs:250-287) for the heat problem P on state u.
int orc_heat_rhs(const deb_heat_problem* P, const double* u, double* du) {
    if (!P || P->n_nodes < 2) return DEB_ERR_BAD_ARG;
    const int64_t N = P->n_nodes;
    const double dx = (P->hi - P->lo) / (double)(N - 1);
    for (int64_t i = 0; i < N; i++) {
        bool lower = (i == 0), upper = (i == N - 1);
        if ((lower && P->bc_lower_kind == 0) || (upper && P->bc_upper_kind == 0)) { du[i] = 0.0; continue; }
        double gl = !lower ? (u[i] - u[i - 1]) / dx : P->bc_lower_value;
        double gu = !upper ? (u[i + 1] - u[i]) / dx : P->bc_upper_value;
        double fl = P->alpha * gl, fu = P->alpha * gu;
        du[i] = 0.0 + (fu - fl) / dx;
    }
    return DEB_OK;
}

int orc_solve_heat_mol(const deb_heat_problem* P, int n_threads) {
    Tableau tb;
    if (!P || !get_tableau(P->method, &tb) || tb.adaptive || P->n_nodes < 2) return DEB_ERR_BAD_ARG;
    if (n_threads <= 0) n_threads = orc_hardware_threads();
    const int64_t N = P->n_nodes;
    const double dx = (P->hi - P->lo) / (double)(N - 1);  // grid.rs:23-36
    const double alpha = P->alpha;
    auto rhs = [&](const double* u, double* du) {
        parallel_for((N + 65535) / 65536, n_threads, [&](int64_t blk) {
            int64_t b = blk * 65536, e = std::min(N, b + 65536);
            for (int64_t i = b; i < e; i++) {
                bool lower = (i == 0), upper = (i == N - 1);
                if ((lower && P->bc_lower_kind == 0) || (upper && P->bc_upper_kind == 0)) { du[i] = 0.0; continue; }  // :264-267
                double gl, gu;
                if (!lower) gl = (u[i] - u[i - 1]) / dx; else gl = P->bc_lower_value;   // :150-172 (Neumann face: prescribed gradient)
                if (!upper) gu = (u[i + 1] - u[i]) / dx; else gu = P->bc_upper_value;
                double fl = alpha * gl, fu = alpha * gu;                                  // PDE::flux
                du[i] = 0.0 + (fu - fl) / dx;                                             // add_scaled_difference :132-139
            }
        });
    };
    const double t0 = P->t0, tf = P->tf;
    double dir = signum(tf - t0);
    auto set = [&](int st, double t, int64_t steps) { if (P->status) *P->status = st; if (P->t_final) *P->t_final = t; if (P->steps) *P->steps = steps; };
    if (!(dir == 1.0 || dir == -1.0)) { set(DEB_STATUS_BAD_INPUT, t0, 0); return DEB_OK; }
    double h0 = P->h;
    if (h0 == 0.0) h0 = std::fabs(tf - t0) / 100.0;
    if (!validate_step_size_parameters(h0, 0.0, INFINITY, t0, tf)) { set(DEB_STATUS_BAD_INPUT, t0, 0); return DEB_OK; }
    const int S = tb.stages, I = tb.dense;
    BigVec y(P->u0, P->u0 + N), dydt(N), ys(N), ynext(N);
    std::vector<BigVec> k(S, BigVec(N));
    rhs(y.data(), dydt.data());
    double t = t0, h = h0;
    int64_t steps = 0;
    int status = DEB_STATUS_COMPLETE;
    const double eps10 = DBL_EPSILON * 10.0;
    for (;;) {
        if ((t + h - tf) * dir > 0.0) {
            double h_new = tf - t;
            if (std::fabs(h_new) < eps10) break;
            h = h_new;
        }
        if (steps >= P->max_steps) { status = DEB_STATUS_MAX_STEPS; break; }
        steps += 1;
        k[0] = dydt;
        for (int s = 1; s < S; s++) {
            ys = y;
            for (int j = 0; j < s; j++) {
                double ah = tb.a[s * I + j] * h;
                parallel_for((N + 65535) / 65536, n_threads, [&](int64_t blk) {
                    int64_t b = blk * 65536, e = std::min(N, b + 65536);
                    for (int64_t i = b; i < e; i++) ys[i] += ah * k[j][i];
                });
            }
            rhs(ys.data(), k[s].data());
        }
        ynext = y;
        for (int s = 0; s < S; s++) {
            double bh = tb.b[s] * h;
            parallel_for((N + 65535) / 65536, n_threads, [&](int64_t blk) {
                int64_t b = blk * 65536, e = std::min(N, b + 65536);
                for (int64_t i = b; i < e; i++) ynext[i] += bh * k[s][i];
            });
        }
        t += h;
        y.swap(ynext);
        rhs(y.data(), dydt.data());
        if (std::fabs(tf - t) <= eps10) break;
    }
    if (P->u_final) std::memcpy(P->u_final, y.data(), sizeof(double) * N);
    set(status, t, steps);
    return DEB_OK;
}

}  // extern "C"
