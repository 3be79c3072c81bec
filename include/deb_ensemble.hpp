// deb_ensemble.hpp -- C++17 host mirror of the crate's builder API for the accelerated path, header-only, on top of the C ABI
// (deb_ensemble.h / libdeb200.so).  The reference is a Rust crate without an FFI; this is the compiled-language host layer a
// caller writes against when no Rust toolchain is at hand, with the crate's names, defaults and error behaviour:
//
//     reference (one problem)                                              here (N problems per call)
//     IVP::ode(&sys, t0, tf, y0)                        src/ivp.rs:279     deb::EnsembleIVP::ode(sys, t0, tf, y0s)
//        .t_eval(points) / .even(dt) / .dense(n)        :656 / :643 / :650    .t_eval(points) / .even(dt) / .dense(n, max_rows)
//        .crossing(c, thr, dir)                         :682                  .crossing(c, thr, dir, max_rows)
//        .hyperplane_crossing(point, normal, ex, dir)   :695                  .hyperplane_crossing(point, normal, components, dir, max_rows)
//        .event(&e)                                     :662                  .event(deb::Event::linear(..).terminal(), max_event_rows)
//        .method(ExplicitRungeKutta::dopri5().rtol(..)) :632                  .method(deb::ExplicitRungeKutta::dopri5().rtol(..))
//        .solve() -> Result<Solution, Error>            :781                  .solve() -> EnsembleSolution; .at(i) returns the Solution of
//                                                                             trajectory i or throws the reference's Error variant
//     IVP::sde(&mut sde, t0, tf, y0).method(ExplicitRungeKutta::euler(h)).solve()   :504,857   deb::EnsembleSDE::sde(sde, t0, tf, y0s, seed)...solve()
//     IVP::pde(&heat, ..).space(MethodOfLines::finite_difference(grid).boundary(bc)).method(rk4(h)).solve()   :419,713   deb::solve_heat_mol(..)
//
// The ODE right-hand side is a built-in system of the crate's tests (deb::System::lorenz(..), ...) or the body of `ODE::diff` as
// CUDA C++ text (deb::System::from_source), compiled into the same kernels at first use.  Per-trajectory parameter sets
// (a sweep) are passed with System::sweep.  There is NO CPU fallback: without a CUDA device solve() throws deb::CallError with
// code DEB_ERR_NO_DEVICE.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "deb_ensemble.h"

namespace deb {

// a failed library CALL (bad arguments, no device, CUDA error): not a property of a trajectory
struct CallError : std::runtime_error {
    int code;
    CallError(int code_, const std::string& what) : std::runtime_error(what), code(code_) {}
};

inline void check(int rc, const char* what) {
    if (rc != DEB_OK) throw CallError(rc, std::string(what) + ": " + deb_last_error());
}

// Error<T, Y> of the crate (src/error.rs:13-41): what `solve()` returns as Err for one trajectory
struct Error : std::runtime_error {
    enum Kind { BadInput, MaxSteps, StepSize, Stiffness } kind;
    double t;
    std::vector<double> y;
    Error(Kind k, double t_, std::vector<double> y_)
        : std::runtime_error(k == BadInput ? "BadInput" : k == MaxSteps ? "MaxSteps" : k == StepSize ? "StepSize" : "Stiffness"), kind(k), t(t_),
          y(std::move(y_)) {}
};

enum class Status { Complete, Interrupted };  // src/status.rs:26 (the variants an Ok(Solution) can carry here)
struct Evals { int function; };               // src/stats.rs:16
struct Steps { int accepted, rejected; };     // src/stats.rs:69

// Solution<T, Y> (src/solution.rs:30-53)
struct Solution {
    std::vector<double> t;
    std::vector<std::vector<double>> y;
    Status status;
    Evals evals;
    Steps steps;
};

enum class CrossingDirection { Both = 0, Positive = 1, Negative = -1 };  // src/solout/crossing.rs

// ExplicitRungeKutta (src/methods/erk/mod.rs:32-228): constructors and option setters with the crate's defaults
class ExplicitRungeKutta {
public:
    static ExplicitRungeKutta dopri5() { return ExplicitRungeKutta(DEB_DOPRI5); }
    static ExplicitRungeKutta dop853() { return ExplicitRungeKutta(DEB_DOP853); }
    static ExplicitRungeKutta rkf45() { return ExplicitRungeKutta(DEB_RKF45); }
    static ExplicitRungeKutta cash_karp() { return ExplicitRungeKutta(DEB_CASH_KARP); }
    static ExplicitRungeKutta rkv655e() { return ExplicitRungeKutta(DEB_RKV655E); }
    static ExplicitRungeKutta rkv656e() { return ExplicitRungeKutta(DEB_RKV656E); }
    static ExplicitRungeKutta rkv766e() { return ExplicitRungeKutta(DEB_RKV766E); }
    static ExplicitRungeKutta rkv767e() { return ExplicitRungeKutta(DEB_RKV767E); }
    static ExplicitRungeKutta rkv877e() { return ExplicitRungeKutta(DEB_RKV877E); }
    static ExplicitRungeKutta rkv878e() { return ExplicitRungeKutta(DEB_RKV878E); }
    static ExplicitRungeKutta rkv988e() { return ExplicitRungeKutta(DEB_RKV988E); }
    static ExplicitRungeKutta rkv989e() { return ExplicitRungeKutta(DEB_RKV989E); }
    static ExplicitRungeKutta euler(double h) { return ExplicitRungeKutta(DEB_EULER).h0(h); }
    static ExplicitRungeKutta midpoint(double h) { return ExplicitRungeKutta(DEB_MIDPOINT).h0(h); }
    static ExplicitRungeKutta heun(double h) { return ExplicitRungeKutta(DEB_HEUN).h0(h); }
    static ExplicitRungeKutta ralston(double h) { return ExplicitRungeKutta(DEB_RALSTON).h0(h); }
    static ExplicitRungeKutta rk4(double h) { return ExplicitRungeKutta(DEB_RK4).h0(h); }
    static ExplicitRungeKutta three_eighths(double h) { return ExplicitRungeKutta(DEB_THREE_EIGHTHS).h0(h); }
    // Milstein::new(h) (src/methods/milstein.rs:37-68): derivative-free Milstein, SDE ensembles only
    static ExplicitRungeKutta milstein(double h) { return ExplicitRungeKutta(DEB_MILSTEIN).h0(h); }

    ExplicitRungeKutta& rtol(double v) { opt_.rtol = v; rtol_vec_.clear(); return *this; }
    ExplicitRungeKutta& atol(double v) { opt_.atol = v; atol_vec_.clear(); return *this; }
    ExplicitRungeKutta& rtol(std::vector<double> v) { rtol_vec_ = std::move(v); return *this; }  // Tolerance::Vector
    ExplicitRungeKutta& atol(std::vector<double> v) { atol_vec_ = std::move(v); return *this; }
    ExplicitRungeKutta& h0(double v) { opt_.h0 = v; return *this; }
    ExplicitRungeKutta& h_min(double v) { opt_.h_min = v; return *this; }
    ExplicitRungeKutta& h_max(double v) { opt_.h_max = v; return *this; }
    ExplicitRungeKutta& max_steps(long long v) { opt_.max_steps = v; return *this; }
    ExplicitRungeKutta& safety_factor(double v) { opt_.safety_factor = v; return *this; }
    ExplicitRungeKutta& min_scale(double v) { opt_.min_scale = v; return *this; }
    ExplicitRungeKutta& max_scale(double v) { opt_.max_scale = v; return *this; }
    ExplicitRungeKutta& max_rejects(long long v) { opt_.max_rejects = v; return *this; }
    // `.filter(|h| f64::from_bits(h.to_bits() & MASK))` keeping the leading `bits` mantissa bits (erk/mod.rs:225)
    ExplicitRungeKutta& filter_truncate_mantissa(int bits) { filter_ = DEB_FILTER_TRUNCATE_MANTISSA; filter_bits_ = bits; return *this; }

    int method_id() const { return method_; }
    // the C struct for a system of dimension `dim` (the vectors stay owned by *this)
    deb_erk_options options(int dim) const {
        deb_erk_options o = opt_;
        if (!rtol_vec_.empty()) {
            if ((int)rtol_vec_.size() != dim) throw std::invalid_argument("rtol vector: one entry per state component");
            o.rtol_vec = rtol_vec_.data();
        }
        if (!atol_vec_.empty()) {
            if ((int)atol_vec_.size() != dim) throw std::invalid_argument("atol vector: one entry per state component");
            o.atol_vec = atol_vec_.data();
        }
        return o;
    }
    int filter() const { return filter_; }
    int filter_bits() const { return filter_bits_; }

private:
    explicit ExplicitRungeKutta(int method) : method_(method) { deb_erk_options_default(&opt_); }
    int method_;
    deb_erk_options opt_;
    std::vector<double> rtol_vec_, atol_vec_;
    int filter_ = DEB_FILTER_IDENTITY, filter_bits_ = 0;
};

// The systems of the crate's tests (tests/ode/systems.rs), or the body of `ODE::diff(t, y, dydt)` as CUDA C++ text
class System {
public:
    static System exponential_growth(double k) { return System(DEB_SYS_EXPONENTIAL, 1, {k}); }
    static System linear_equation(double a, double b) { return System(DEB_SYS_LINEAR, 1, {a, b}); }
    static System harmonic_oscillator(double k) { return System(DEB_SYS_HARMONIC, 2, {k}); }
    static System logistic_equation(double k, double m) { return System(DEB_SYS_LOGISTIC, 1, {k, m}); }
    static System van_der_pol(double mu) { return System(DEB_SYS_VAN_DER_POL, 2, {mu}); }
    static System lorenz(double sigma, double rho, double beta) { return System(DEB_SYS_LORENZ, 3, {sigma, rho, beta}); }
    static System brusselator(double a, double b) { return System(DEB_SYS_BRUSSELATOR, 2, {a, b}); }
    static System robertson() { return System(DEB_SYS_ROBERTSON, 3, {}); }
    // `impl ODE for S { fn diff(&self, t, y, dydt) { <body> } }` with the struct's fields as p[0..n_params)
    static System from_source(int dim, const std::string& diff_body, std::vector<double> params) {
        int32_t id = 0;
        check(deb_define_ode(dim, (int)params.size(), diff_body.c_str(), &id), "deb_define_ode");
        return System(id, dim, std::move(params));
    }
    // ForwardSensitivityOde::new (src/ode/sensitivity/forward.rs:43-116): the augmented system z = [y, S], S' = J_y S + J_p, generated from the
    // bodies of `diff`, `jacobian` (J[r*dim + c] = df_r/dy_c) and `jacobian_p` (Jp[r*n_params + c] = df_r/dp_c); its dimension is dim * (1 + n_params)
    static System sensitivity_from_source(int dim, const std::string& diff_body, const std::string& jacobian_body, const std::string& jacobian_p_body,
                                          std::vector<double> params) {
        int32_t id = 0;
        check(deb_define_ode_sensitivity(dim, (int)params.size(), diff_body.c_str(), jacobian_body.c_str(), jacobian_p_body.c_str(), &id),
              "deb_define_ode_sensitivity");
        const int np = (int)params.size();
        return System(id, dim * (1 + np), std::move(params));
    }
    // one parameter set per trajectory, [n_traj][n_params] flat (a parameter sweep)
    System& sweep(std::vector<double> per_trajectory) { sweep_ = std::move(per_trajectory); return *this; }

    int id() const { return id_; }
    int dim() const { return dim_; }
    int n_params() const { return (int)params_.size(); }
    const std::vector<double>& params() const { return params_; }
    const std::vector<double>& sweep_params() const { return sweep_; }

private:
    System(int id, int dim, std::vector<double> p) : id_(id), dim_(dim), params_(std::move(p)) {}
    int id_, dim_;
    std::vector<double> params_, sweep_;
};

// `impl Event` (src/solout/event.rs:60-70) + its EventConfig
class Event {
public:
    // g(t, y) = c0 + c1*t + sum_i coef[i]*y[i]
    static Event linear(double c0, double c1, const std::vector<double>& coef) {
        Event e;
        e.id_ = DEB_EVENT_LINEAR;
        e.coef_.assign(DEB_MAX_DIM + 2, 0.0);
        e.coef_[0] = c0;
        e.coef_[1] = c1;
        if (coef.size() > (size_t)DEB_MAX_DIM) throw std::invalid_argument("too many event coefficients");
        for (size_t i = 0; i < coef.size(); i++) e.coef_[2 + i] = coef[i];
        return e;
    }
    // the body of `fn event(&self, t, y) -> T` as CUDA C++ text: `double event(double t, const double* y, const double* p)`
    static Event from_source(int dim, const std::string& body) {
        Event e;
        int32_t id = 0;
        check(deb_define_event(dim, body.c_str(), &id), "deb_define_event");
        e.id_ = id;
        e.coef_.assign(DEB_MAX_DIM + 2, 0.0);
        return e;
    }
    Event& direction(CrossingDirection d) { direction_ = (int)d; return *this; }
    Event& terminal() { terminate_ = 1; return *this; }
    Event& terminate_after(int count) { terminate_ = count; return *this; }

    int id() const { return id_; }
    int direction_code() const { return direction_; }
    int terminate_count() const { return terminate_; }
    const std::vector<double>& coef() const { return coef_; }

private:
    int id_ = DEB_EVENT_NONE, direction_ = 0, terminate_ = 0;
    std::vector<double> coef_;
};

// All N results of one ensemble solve: flat arrays plus the per-trajectory view of the crate
class EnsembleSolution {
public:
    long long n = 0;
    int dim = 0, row_capacity = 0;
    std::vector<double> t_rows;                  // times of the rows of a t_eval / even(dt) recorder, in integration order
    std::vector<double> y_eval;                  // [n][row_capacity][dim]
    std::vector<double> t_out;                   // [n][row_capacity] (recorders whose row times depend on the trajectory), else empty
    std::vector<int32_t> n_emitted, status, accepted, rejected, evals;
    std::vector<double> t_final, y_final;
    std::vector<double> stats_sums;              // [row][dim][2] sums of {y, y^2} over the ensemble (with_stats()), else empty
    std::vector<int64_t> stats_counts;
    float kernel_ms = 0.f, total_ms = 0.f;
    int gpu_launches = 0;
    double even_tf = std::numeric_limits<double>::quiet_NaN();

    // Result<Solution, Error> of trajectory i: the Solution, or the reference's Error variant as an exception
    Solution at(long long i) const {
        std::vector<double> yf(y_final.begin() + i * dim, y_final.begin() + (i + 1) * dim);
        switch (status[i]) {
            case DEB_STATUS_BAD_INPUT: throw Error(Error::BadInput, t_final[i], yf);
            case DEB_STATUS_MAX_STEPS: throw Error(Error::MaxSteps, t_final[i], yf);
            case DEB_STATUS_STEP_SIZE: throw Error(Error::StepSize, t_final[i], yf);
            case DEB_STATUS_STIFFNESS: throw Error(Error::Stiffness, t_final[i], yf);
            default: break;
        }
        const int m = n_emitted[i];
        if (m > row_capacity) throw std::length_error("trajectory produced more rows than the row capacity: raise max_rows");
        Solution s;
        s.status = status[i] == DEB_STATUS_INTERRUPTED ? Status::Interrupted : Status::Complete;
        s.evals = Evals{evals[i]};
        s.steps = Steps{accepted[i], rejected[i]};
        for (int r = 0; r < m; r++) {
            double tr;
            if (!t_out.empty()) tr = t_out[(size_t)i * row_capacity + r];
            else if (r == m - 1 && even_tf == even_tf && t_final[i] == even_tf) tr = even_tf;  // EvenSolout: the final point (even.rs:166-188)
            else tr = (size_t)r < t_rows.size() ? t_rows[(size_t)r] : std::numeric_limits<double>::quiet_NaN();
            s.t.push_back(tr);
            const double* row = &y_eval[((size_t)i * row_capacity + r) * dim];
            s.y.emplace_back(row, row + dim);
        }
        return s;
    }
    bool ok(long long i) const { return status[i] == DEB_STATUS_COMPLETE || status[i] == DEB_STATUS_INTERRUPTED; }
};

// Mirror of the `IVP` builder (src/ivp.rs) for an ensemble of N problems sharing (t0, tf, method, recorder)
class EnsembleIVP {
public:
    // y0s: [n][dim] flat -- the memory of a Vec<[f64; N]>
    static EnsembleIVP ode(System system, double t0, double tf, std::vector<double> y0s) {
        if (system.dim() <= 0 || y0s.size() % (size_t)system.dim() != 0) throw std::invalid_argument("y0s: n * dim values expected");
        return EnsembleIVP(std::move(system), t0, tf, std::move(y0s));
    }
    EnsembleIVP& t_eval(std::vector<double> points) { t_eval_ = std::move(points); solout_ = DEB_SOLOUT_T_EVAL; even_dt_ = 0.0; return *this; }
    EnsembleIVP& even(double dt) { even_dt_ = dt; solout_ = DEB_SOLOUT_EVEN; t_eval_.clear(); return *this; }
    // what a plain IVP::solve() records: (t0, y0) and every accepted step; max_rows row slots per trajectory
    EnsembleIVP& every_step(int max_rows) { return per_step(DEB_SOLOUT_DEFAULT, max_rows); }
    EnsembleIVP& dense(int n, int max_rows) { dense_n_ = n; return per_step(DEB_SOLOUT_DENSE, max_rows); }
    EnsembleIVP& crossing(int component, double threshold, CrossingDirection d = CrossingDirection::Both, int max_rows = 64) {
        cross_component_ = component; cross_threshold_ = threshold; cross_direction_ = (int)d;
        return per_step(DEB_SOLOUT_CROSSING, max_rows);
    }
    EnsembleIVP& hyperplane_crossing(std::vector<double> point, std::vector<double> normal, std::vector<int> components,
                                     CrossingDirection d = CrossingDirection::Both, int max_rows = 64) {
        if (point.size() != normal.size() || point.size() != components.size() || point.empty() || point.size() > (size_t)DEB_MAX_DIM)
            throw std::invalid_argument("point, normal and components must have the same, non-zero length");
        plane_point_ = std::move(point); plane_normal_ = std::move(normal); plane_index_ = std::move(components); cross_direction_ = (int)d;
        return per_step(DEB_SOLOUT_HYPERPLANE, max_rows);
    }
    // wrap the current recorder with event detection; max_event_rows = extra row slots for event rows (with no recorder set: the
    // whole row capacity of the default every-step recorder)
    EnsembleIVP& event(Event e, int max_event_rows = 16) { event_ = std::move(e); has_event_ = true; max_event_rows_ = max_event_rows; return *this; }
    EnsembleIVP& method(ExplicitRungeKutta m) { method_.assign(1, std::move(m)); return *this; }
    EnsembleIVP& device(int d) { device_ = d; devices_.clear(); return *this; }
    // several GPUs behind one call (blocks of 4096 trajectories dealt round-robin)
    EnsembleIVP& devices(std::vector<int> d) { devices_ = std::move(d); return *this; }
    // per-row ensemble sums of {y, y^2}, reduced on the device(s)
    EnsembleIVP& with_stats() { stats_ = true; return *this; }

    EnsembleSolution solve() const {
        if (method_.empty()) throw std::invalid_argument("method(...) must be set before solve()");
        const ExplicitRungeKutta& m = method_[0];
        const int dim = system_.dim();
        const long long n = (long long)(y0s_.size() / (size_t)dim);
        int solout = solout_;
        bool per_step_rec = solout == DEB_SOLOUT_DEFAULT || solout == DEB_SOLOUT_DENSE || solout == DEB_SOLOUT_CROSSING || solout == DEB_SOLOUT_HYPERPLANE;
        int n_eval = (int)t_eval_.size();
        if (solout == DEB_SOLOUT_EVEN) n_eval = (int)std::floor(std::fabs(tf_ - t0_) / even_dt_) + 3;
        if (per_step_rec) n_eval = max_rows_;
        int rows_cap = n_eval;
        if (has_event_) {
            if (!per_step_rec && solout == DEB_SOLOUT_T_EVAL && t_eval_.empty()) {  // plain solve().event(): every step + events
                solout = DEB_SOLOUT_DEFAULT;
                per_step_rec = true;
                n_eval = rows_cap = max_event_rows_;
            } else if (!per_step_rec) {
                rows_cap = n_eval + max_event_rows_;
            }
        }
        const bool with_times = per_step_rec || has_event_;

        EnsembleSolution out;
        out.n = n;
        out.dim = dim;
        out.row_capacity = rows_cap;
        const double nan = std::numeric_limits<double>::quiet_NaN();
        out.y_eval.assign((size_t)n * rows_cap * dim, nan);
        if (with_times) out.t_out.assign((size_t)n * rows_cap, nan);
        out.n_emitted.assign(n, 0);
        out.status.assign(n, -1);
        out.accepted.assign(n, 0);
        out.rejected.assign(n, 0);
        out.evals.assign(n, 0);
        out.t_final.assign(n, 0.0);
        out.y_final.assign((size_t)n * dim, 0.0);
        std::vector<double> t_rows((size_t)(n_eval > 0 ? n_eval : 1), 0.0);
        if (stats_) {
            out.stats_sums.assign((size_t)rows_cap * dim * 2, 0.0);
            out.stats_counts.assign((size_t)rows_cap, 0);
        }

        deb_ode_problem P;
        std::memset(&P, 0, sizeof P);
        P.struct_size = sizeof P;
        P.system = system_.id();
        P.method = m.method_id();
        P.dim = dim;
        P.n_params = system_.n_params();
        P.n_traj = n;
        P.y0 = y0s_.data();
        if (!system_.sweep_params().empty()) {
            if (system_.sweep_params().size() != (size_t)n * system_.n_params()) throw std::invalid_argument("sweep: one parameter set per trajectory");
            P.params = system_.sweep_params().data();
            P.params_shared = 0;
        } else {
            P.params = system_.params().data();
            P.params_shared = 1;
        }
        P.n_eval = n_eval;
        P.t_eval = t_eval_.empty() ? nullptr : t_eval_.data();
        P.t0 = t0_;
        P.tf = tf_;
        P.opt = m.options(dim);
        P.device = device_;
        P.memspace = DEB_MEM_HOST;
        P.solout = solout;
        P.dense_n = dense_n_;
        P.even_dt = even_dt_;
        P.cross_component = cross_component_;
        P.cross_direction = cross_direction_;
        P.cross_threshold = cross_threshold_;
        if (solout == DEB_SOLOUT_HYPERPLANE) {
            P.plane_dim = (int)plane_index_.size();
            for (size_t q = 0; q < plane_index_.size(); q++) {
                P.plane_index[q] = plane_index_[q];
                P.plane_point[q] = plane_point_[q];
                P.plane_normal[q] = plane_normal_[q];
            }
        }
        if (has_event_) {
            P.event = event_.id();
            P.event_direction = event_.direction_code();
            P.event_terminate = event_.terminate_count();
            P.row_capacity = rows_cap;
            for (size_t q = 0; q < event_.coef().size(); q++) P.event_coef[q] = event_.coef()[q];
        }
        P.filter = m.filter();
        P.filter_bits = m.filter_bits();
        P.layout = DEB_LAYOUT_TRAJ_MAJOR;
        if (devices_.size() > (size_t)DEB_MAX_DEVICES) throw std::invalid_argument("too many devices");
        P.n_devices = (int)devices_.size();
        for (size_t q = 0; q < devices_.size(); q++) P.devices[q] = devices_[q];

        deb_result R;
        std::memset(&R, 0, sizeof R);
        R.struct_size = sizeof R;
        R.y_eval = out.y_eval.empty() ? nullptr : out.y_eval.data();
        R.n_emitted = out.n_emitted.data();
        R.t_final = out.t_final.data();
        R.y_final = out.y_final.data();
        R.status = out.status.data();
        R.accepted = out.accepted.data();
        R.rejected = out.rejected.data();
        R.evals = out.evals.data();
        R.t_rows = t_rows.data();
        R.t_out = out.t_out.empty() ? nullptr : out.t_out.data();
        R.stats_sums = out.stats_sums.empty() ? nullptr : out.stats_sums.data();
        R.stats_counts = out.stats_counts.empty() ? nullptr : out.stats_counts.data();
        check(deb_solve_ode(&P, &R), "deb_solve_ode");
        out.t_rows.assign(t_rows.begin(), t_rows.begin() + R.n_rows);
        out.kernel_ms = R.kernel_ms;
        out.total_ms = R.total_ms;
        out.gpu_launches = R.gpu_launches;
        if (solout == DEB_SOLOUT_EVEN && !has_event_) out.even_tf = tf_;
        return out;
    }

private:
    EnsembleIVP(System s, double t0, double tf, std::vector<double> y0s) : system_(std::move(s)), t0_(t0), tf_(tf), y0s_(std::move(y0s)) {}
    EnsembleIVP& per_step(int solout, int max_rows) {
        if (max_rows < 1) throw std::invalid_argument("max_rows must be >= 1");
        solout_ = solout; max_rows_ = max_rows; t_eval_.clear(); even_dt_ = 0.0;
        return *this;
    }
    System system_;
    double t0_, tf_;
    std::vector<double> y0s_, t_eval_;
    int solout_ = DEB_SOLOUT_T_EVAL, max_rows_ = 0, dense_n_ = 0;
    double even_dt_ = 0.0;
    int cross_component_ = 0, cross_direction_ = 0;
    double cross_threshold_ = 0.0;
    std::vector<double> plane_point_, plane_normal_;
    std::vector<int> plane_index_;
    Event event_;
    bool has_event_ = false;
    int max_event_rows_ = 16;
    std::vector<ExplicitRungeKutta> method_;  // 0 or 1 entries (ExplicitRungeKutta has no default constructor, like the crate's builder)
    int device_ = 0;
    std::vector<int> devices_;
    bool stats_ = false;
};

// `impl SDE` (src/sde/sde.rs:16-67): the systems of the crate's SDE examples, or drift / diffusion / noise bodies as CUDA C++ text.
// The Wiener increments come from a counter-based Philox stream keyed by (seed, path index), see deb_sde_problem.
class SdeSystem {
public:
    static SdeSystem ornstein_uhlenbeck(double theta, double mu, double sigma) { return SdeSystem(DEB_SDE_OU, 1, {theta, mu, sigma}); }
    static SdeSystem geometric_brownian_motion(double mu, double sigma) { return SdeSystem(DEB_SDE_GBM, 1, {mu, sigma}); }
    // examples/sde/02_heston_model: y = (price, variance)
    static SdeSystem heston(double mu, double kappa, double theta, double sigma, double rho) { return SdeSystem(DEB_SDE_HESTON, 2, {mu, kappa, theta, sigma, rho}); }
    // drift(t, y, dydt, p), diffusion(t, y, g, p) (diagonal noise) and, optionally, noise(dw, p) mixing the independent increments in place
    static SdeSystem from_source(int dim, const std::string& drift_body, const std::string& diffusion_body, std::vector<double> params,
                                 const std::string& noise_body = std::string()) {
        int32_t id = 0;
        check(deb_define_sde(dim, (int)params.size(), drift_body.c_str(), diffusion_body.c_str(), noise_body.empty() ? nullptr : noise_body.c_str(), &id),
              "deb_define_sde");
        return SdeSystem(id, dim, std::move(params));
    }
    int id() const { return id_; }
    int dim() const { return dim_; }
    const std::vector<double>& params() const { return params_; }

private:
    SdeSystem(int id, int dim, std::vector<double> p) : id_(id), dim_(dim), params_(std::move(p)) {}
    int id_, dim_;
    std::vector<double> params_;
};

// `IVP::sde(&mut sde, t0, tf, y0)` (src/ivp.rs:504) for N paths
class EnsembleSDE {
public:
    static EnsembleSDE sde(SdeSystem system, double t0, double tf, std::vector<double> y0s, uint64_t seed) {
        if (y0s.size() % (size_t)system.dim() != 0) throw std::invalid_argument("y0s: n * dim values expected");
        return EnsembleSDE(std::move(system), t0, tf, std::move(y0s), seed);
    }
    EnsembleSDE& t_eval(std::vector<double> points) { t_eval_ = std::move(points); return *this; }
    EnsembleSDE& method(ExplicitRungeKutta m) { method_.assign(1, std::move(m)); return *this; }
    EnsembleSDE& device(int d) { device_ = d; return *this; }
    // global index of path 0 (an ensemble split over several calls keeps its noise)
    EnsembleSDE& path_offset(long long o) { path_offset_ = o; return *this; }

    EnsembleSolution solve() const {
        if (method_.empty()) throw std::invalid_argument("method(...) must be set before solve()");
        const int dim = system_.dim();
        const long long n = (long long)(y0s_.size() / (size_t)dim);
        const int n_eval = (int)t_eval_.size();
        EnsembleSolution out;
        out.n = n;
        out.dim = dim;
        out.row_capacity = n_eval;
        out.y_eval.assign((size_t)n * n_eval * dim, std::numeric_limits<double>::quiet_NaN());
        out.n_emitted.assign(n, 0);
        out.status.assign(n, -1);
        out.accepted.assign(n, 0);
        out.rejected.assign(n, 0);
        out.evals.assign(n, 0);
        out.t_final.assign(n, 0.0);
        out.y_final.assign((size_t)n * dim, 0.0);
        std::vector<double> t_rows((size_t)(n_eval > 0 ? n_eval : 1), 0.0);
        deb_sde_problem P;
        std::memset(&P, 0, sizeof P);
        P.struct_size = sizeof P;
        P.system = system_.id();
        P.method = method_[0].method_id();
        P.dim = dim;
        P.n_params = (int)system_.params().size();
        P.n_traj = n;
        P.y0 = y0s_.data();
        P.params_shared = 1;
        P.params = system_.params().data();
        P.n_eval = n_eval;
        P.t_eval = t_eval_.empty() ? nullptr : t_eval_.data();
        P.t0 = t0_;
        P.tf = tf_;
        P.opt = method_[0].options(dim);
        P.seed = seed_;
        P.path_offset = path_offset_;
        P.device = device_;
        P.memspace = DEB_MEM_HOST;
        deb_result R;
        std::memset(&R, 0, sizeof R);
        R.struct_size = sizeof R;
        R.y_eval = out.y_eval.empty() ? nullptr : out.y_eval.data();
        R.n_emitted = out.n_emitted.data();
        R.t_final = out.t_final.data();
        R.y_final = out.y_final.data();
        R.status = out.status.data();
        R.accepted = out.accepted.data();
        R.rejected = out.rejected.data();
        R.evals = out.evals.data();
        R.t_rows = t_rows.data();
        check(deb_solve_sde(&P, &R), "deb_solve_sde");
        out.t_rows.assign(t_rows.begin(), t_rows.begin() + R.n_rows);
        out.kernel_ms = R.kernel_ms;
        out.total_ms = R.total_ms;
        out.gpu_launches = R.gpu_launches;
        return out;
    }

private:
    EnsembleSDE(SdeSystem s, double t0, double tf, std::vector<double> y0s, uint64_t seed)
        : system_(std::move(s)), t0_(t0), tf_(tf), y0s_(std::move(y0s)), seed_(seed) {}
    SdeSystem system_;
    double t0_, tf_;
    std::vector<double> y0s_, t_eval_;
    uint64_t seed_;
    long long path_offset_ = 0;
    std::vector<ExplicitRungeKutta> method_;
    int device_ = 0;
};

// BoundaryCondition of the method-of-lines grid (src/pde/boundary.rs)
struct Boundary {
    int kind;  // 0 = Dirichlet(value), 1 = Neumann(gradient)
    double value;
    static Boundary dirichlet(double v) { return Boundary{0, v}; }
    static Boundary neumann(double g) { return Boundary{1, g}; }
};
struct HeatSolution {
    std::vector<double> u;
    double t;
    long long steps;
    int status;  // deb_status
};
// IVP::pde(&HeatEquation{alpha}, t0, tf, u0).space(MethodOfLines::finite_difference(StructuredGrid::uniform([lo], [hi], [n])).boundary(..))
//     .method(ExplicitRungeKutta::rk4(h)).solve()   (src/ivp.rs:419,713; tests/pde/method_of_lines.rs:37-70): one large state on one GPU
inline HeatSolution solve_heat_mol(const std::vector<double>& u0, double lo, double hi, double alpha, const ExplicitRungeKutta& method, double t0,
                                   double tf, Boundary lower = Boundary::dirichlet(0.0), Boundary upper = Boundary::dirichlet(0.0), int device = 0) {
    HeatSolution out;
    out.u.assign(u0.size(), 0.0);
    out.t = 0.0;
    int64_t steps = 0;
    int32_t status = -1;
    const deb_erk_options o = method.options(1);
    deb_heat_problem P;
    std::memset(&P, 0, sizeof P);
    P.struct_size = sizeof P;
    P.n_nodes = (int64_t)u0.size();
    P.lo = lo;
    P.hi = hi;
    P.alpha = alpha;
    P.bc_lower_kind = lower.kind;
    P.bc_lower_value = lower.value;
    P.bc_upper_kind = upper.kind;
    P.bc_upper_value = upper.value;
    P.method = method.method_id();
    P.h = o.h0;
    P.t0 = t0;
    P.tf = tf;
    P.max_steps = o.max_steps;
    P.u0 = u0.data();
    P.u_final = out.u.data();
    P.t_final = &out.t;
    P.steps = &steps;
    P.status = &status;
    P.device = device;
    P.memspace = DEB_MEM_HOST;
    check(deb_solve_heat_mol(&P), "deb_solve_heat_mol");
    out.steps = steps;
    out.status = status;
    return out;
}

}  // namespace deb
