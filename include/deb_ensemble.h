/* deb_ensemble.h -- C ABI of the B200 ensemble integrator (libdeb200.so).
 *
 * The reference crate (Ryan-D-Gast/differential-equations, v0.6.1) has no FFI: its operator API is the Rust
 * trait surface.  This header is the boundary a Rust shim (differential-equations_b200/rust/)
 * binds with `extern "C"`; each entry point replaces the per-trajectory call chain named beside it for an
 * ENSEMBLE of independent initial-value problems:
 *
 *   deb_solve_ode     <->  IVP::ode(&sys,t0,tf,y0).t_eval(pts).method(ExplicitRungeKutta::<m>()...).solve()
 *                          src/ivp.rs:279,656,632,781 -> solve_ode src/ode/solve_ivp.rs:116-277
 *                          -> init/step/interpolate src/methods/erk/dormandprince/ordinary.rs:16,63,301,
 *                             src/methods/erk/adaptive/ordinary.rs:16,63,246 or src/methods/erk/fixed/ordinary.rs:16,58,169
 *                          -> the recorder: TEvalSolout src/solout/t_eval.rs:87-137, or .even(dt) / plain solve() /
 *                             .dense(n) / .crossing(..) / .hyperplane_crossing(..) (src/solout/{even,default,dense,
 *                             crossing,hyperplane}.rs), optionally wrapped by .event(&e) (src/solout/event.rs:300-470)
 *   deb_define_ode    <->  impl ODE for S { fn diff(&self, t, y, dydt) }  (src/ode/ode.rs:20-44), as CUDA C++ text
 *   deb_define_event  <->  impl Event for S { fn event(&self, t, y) -> T }  (src/solout/event.rs:60-70), as CUDA C++ text
 *   deb_solve_sde     <->  IVP::sde(&mut sde,t0,tf,y0).t_eval(pts).method(ExplicitRungeKutta::euler(h)).solve()
 *                          src/ivp.rs:504,857 -> solve_sde src/sde/solve_ivp.rs:135-287
 *                          -> src/methods/erk/fixed/stochastic.rs:18,67,177 ; SDE::noise (src/sde/sde.rs:67)
 *                             is replaced by counter-based Philox4x32-10 Wiener increments (documented below)
 *                          scalar OU / GBM and the two-dimensional Heston model (diagonal noise); Milstein::new(h)
 *   deb_solve_heat_mol <-> IVP::pde(..).space(MethodOfLines::finite_difference(grid).boundary(bc))
 *                             .method(ExplicitRungeKutta::rk4(h)).solve()
 *                          src/ivp.rs:419,713 ; RHS = SemiDiscretePde::diff src/pde/semi_discrete.rs:250-287
 *   deb_ensemble_stats <-> (new) per-t_eval ensemble mean / variance sums, the only quantity that is
 *                          all-reduced across GPUs.
 *
 * Conventions
 *  - Plain C: pointers + sizes, no C++/torch types.  All structs start with `struct_size` (set it to
 *    sizeof(the struct)).  Fields are only ever appended: a caller built against an OLDER header passes a smaller
 *    struct_size and the library reads the missing tail as zeros (= the defaults); a struct_size larger than the
 *    library's own (a NEWER header) is rejected with DEB_ERR_BAD_ARG.
 *  - Every function returns 0 on success, a negative deb_error otherwise; deb_last_error() gives the text.
 *    Numerical failures of a trajectory are NOT call failures: they are reported per trajectory in
 *    `status[]` with the same meaning as the reference's `Error` variants (src/error.rs:13-41), and the
 *    trajectory's (t,y) at failure is returned in t_final/y_final like `Error::MaxSteps{t,y}`.
 *  - Buffers are caller-owned.  `memspace` says where ALL data pointers of a call live: DEB_MEM_HOST
 *    (the library stages through pinned memory and copies both ways, synchronous call) or DEB_MEM_DEVICE
 *    (pointers are device pointers on `device`; the call enqueues on `stream` and returns without
 *    synchronising).
 *  - There is NO CPU fallback: without a CUDA device every solve returns DEB_ERR_NO_DEVICE.
 *  - Layout: trajectory i, component c of y0 / y_final lives at  ptr[i*dim + c]  ("array of states",
 *    i.e. a Rust Vec<[f64; N]> reinterpreted), y_eval at ptr[(i*n_eval + r)*dim + c] for emitted row r.
 */
#ifndef DEB_ENSEMBLE_H
#define DEB_ENSEMBLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEB_ABI_VERSION 9
#define DEB_MAX_DIM 16 /* widest state the register-resident kernels are instantiated for */
#define DEB_MAX_DEVICES 16 /* longest device list of one call */

typedef enum deb_error {
    DEB_OK = 0,
    DEB_ERR_BAD_ARG = -1,    /* malformed problem description (NULL pointer, dim mismatch, ...) */
    DEB_ERR_NO_DEVICE = -2,  /* no CUDA device / extension not usable: the product path never falls back to a CPU */
    DEB_ERR_CUDA = -3,       /* CUDA runtime error, text in deb_last_error() */
    DEB_ERR_UNSUPPORTED = -4 /* valid request that this build has no kernel for (system x method x dim) */
} deb_error;

/* Explicit Runge-Kutta constructors of the reference that the kernels implement
 * (src/methods/erk/fixed/mod.rs:41-89, src/methods/erk/dormandprince/mod.rs:45-58, src/methods/erk/adaptive/mod.rs:47-122). */
typedef enum deb_method {
    DEB_EULER = 0,
    DEB_MIDPOINT = 1,
    DEB_HEUN = 2,
    DEB_RALSTON = 3,
    DEB_SSP_RK3 = 4,
    DEB_RK4 = 5,
    DEB_THREE_EIGHTHS = 6,
    DEB_DOPRI5 = 16,
    DEB_DOP853 = 17,
    /* adaptive family with y_high - y_low error estimate (src/methods/erk/adaptive/mod.rs:47-60) */
    DEB_RKF45 = 18,
    DEB_CASH_KARP = 19,
    /* Verner pairs with their own dense-output polynomial (src/methods/erk/adaptive/mod.rs:59-122, src/tableau/verner.rs):
     * order(embedded order, interpolant order); the two 6(5) pairs are FSAL */
    DEB_RKV655E = 20,
    DEB_RKV656E = 21,
    DEB_RKV766E = 22,
    DEB_RKV767E = 23,
    DEB_RKV877E = 24,
    DEB_RKV878E = 25,
    DEB_RKV988E = 26,
    DEB_RKV989E = 27,
    /* SDE only: derivative-free Milstein, Milstein::new(h) (src/methods/milstein.rs:37-180) */
    DEB_MILSTEIN = 32
} deb_method;

/* Built-in right-hand sides (`ODE::diff`, src/ode/ode.rs:44).  A Rust closure cannot cross to the device,
 * so the systems of the reference's own tests/benches are compiled in, with the reference's expression
 * order (tests/ode/systems.rs).  params are per trajectory (or shared, stride 0). */
typedef enum deb_system {
    DEB_SYS_EXPONENTIAL = 0, /* dim 1, params {k}:         y' = k*y                  systems.rs:12-16  */
    DEB_SYS_LINEAR = 1,      /* dim 1, params {a,b}:       y' = a + b*y              systems.rs:25-29  */
    DEB_SYS_HARMONIC = 2,    /* dim 2, params {k}:         y0' = y1; y1' = -k*y0     systems.rs:38-43  */
    DEB_SYS_LOGISTIC = 3,    /* dim 1, params {k,m}:       y' = k*y*(1 - y/m)        systems.rs:55-59  */
    DEB_SYS_VAN_DER_POL = 4, /* dim 2, params {mu}:        systems.rs:70-78                             */
    DEB_SYS_LORENZ = 5,      /* dim 3, params {sigma,rho,beta}: systems.rs:91-101                       */
    DEB_SYS_BRUSSELATOR = 6, /* dim 2, params {a,b}:       systems.rs:112-120                           */
    DEB_SYS_ROBERTSON = 7    /* dim 3, no params (stiff):  systems.rs:161-173                           */
} deb_system;

/* Built-in SDEs (`SDE::drift/diffusion`, src/sde/sde.rs:16-52), scalar state. */
typedef enum deb_sde_system {
    DEB_SDE_OU = 0, /* params {theta,mu,sigma}: drift theta*(mu-y), diffusion sigma   examples/sde/03_ornstein_uhlenbeck/main.rs:42-49 */
    DEB_SDE_GBM = 1, /* params {mu,sigma}:       drift mu*y,         diffusion sigma*y src/sde/solve_ivp.rs doc example              */
    DEB_SDE_HESTON = 2 /* dim 2 {price, variance}, params {mu,kappa,theta,sigma,rho}: dS = mu S dt + sqrt(v) S dW1,
                          dv = kappa (theta - v) dt + sigma sqrt(v) dW2, dW2 = rho dW1 + sqrt(1-rho^2) dW2'
                          (examples/sde/02_heston_model/main.rs:53-72) */
} deb_sde_system;

/* Per-trajectory outcome, mirrors Status/Error of the reference (src/status.rs:26, src/error.rs:13-41). */
typedef enum deb_status {
    DEB_STATUS_COMPLETE = 0,
    DEB_STATUS_MAX_STEPS = 1, /* Error::MaxSteps{t,y}   dormandprince/ordinary.rs:82-91 */
    DEB_STATUS_STEP_SIZE = 2, /* Error::StepSize{t,y}   dormandprince/ordinary.rs:70-79 */
    DEB_STATUS_STIFFNESS = 3, /* Error::Stiffness{t,y}  dormandprince/ordinary.rs:165-194 */
    DEB_STATUS_BAD_INPUT = 4, /* Error::BadInput{..}    utils.rs:60-157, solve_ivp.rs:139-147 */
    DEB_STATUS_INTERRUPTED = 5 /* Ok(Solution) with Status::Interrupted: an event asked to terminate (solve_ivp.rs:255-260) */
} deb_status;

/* Event functions g(t, y) (the `Event` trait, src/solout/event.rs:60-70) for deb_ode_problem.event */
enum {
    DEB_EVENT_NONE = 0,
    DEB_EVENT_LINEAR = 1 /* g(t, y) = c[0] + c[1]*t + sum_i c[2+i]*y[i] (event_coef), accumulated in that order */
    /* >= 1000: ids returned by deb_define_event */
};

typedef enum deb_memspace { DEB_MEM_HOST = 0, DEB_MEM_DEVICE = 1 } deb_memspace;

/* Memory layout of y_eval (t_eval / even(dt) recorders). */
typedef enum deb_layout {
    DEB_LAYOUT_TRAJ_MAJOR = 0, /* y_eval[i][r][c]: each trajectory's Solution.y is one contiguous block (Vec<[f64; N]>) */
    DEB_LAYOUT_ROW_MAJOR = 1   /* y_eval[r][c][i]: structure of arrays, one contiguous vector of n_traj values per (row, component):
                                  what a per-t_eval reduction or a columnar sink (the crate's polars feature) reads */
} deb_layout;

/* Step-size filter: the `filter: fn(T) -> T` hook of ExplicitRungeKutta (src/methods/erk/mod.rs:85,225), applied to h at init
 * (dormandprince/ordinary.rs:33), after every step-size update (:267) and in set_h, i.e. to the clip at tf (:288).  A Rust
 * function pointer cannot cross the boundary; the built-ins below can. */
typedef enum deb_filter {
    DEB_FILTER_IDENTITY = 0,          /* |h| h   (the default, erk/mod.rs:144) */
    DEB_FILTER_TRUNCATE_MANTISSA = 1  /* |h| f64::from_bits(h.to_bits() & !((1u64 << (52 - filter_bits)) - 1)): keep the leading
                                         filter_bits (1..52) bits of the mantissa; quantises h so that results do not depend on
                                         the last bits of the controller's pow ("Plan B") */
} deb_filter;

/* Output recorder (`Solout`, src/solout/): which rows go to y_eval. */
typedef enum deb_solout {
    DEB_SOLOUT_T_EVAL = 0, /* IVP::t_eval(points): TEvalSolout, src/solout/t_eval.rs:87-171 */
    DEB_SOLOUT_EVEN = 1,   /* IVP::even(dt): EvenSolout, src/solout/even.rs:69-199 -- t0, t0+dt, t0+2dt, ... (accumulated),
                              every point interpolated, plus the exact final state when a step lands on tf */
    /* per-step recorders: rows carry their own times (deb_result.t_out), n_eval is the row capacity per trajectory */
    DEB_SOLOUT_DEFAULT = 2,  /* the recorder of a plain IVP::solve(): every accepted step, src/solout/default.rs:54-75 */
    DEB_SOLOUT_DENSE = 3,    /* IVP::dense(n): DenseSolout, src/solout/dense.rs:74-108 */
    DEB_SOLOUT_CROSSING = 4, /* IVP::crossing(component, threshold, direction): CrossingSolout, src/solout/crossing.rs:115-263 */
    DEB_SOLOUT_HYPERPLANE = 5 /* IVP::hyperplane_crossing(point, normal, extractor, direction): HyperplaneCrossingSolout,
                                 src/solout/hyperplane.rs:170-330; the extractor is a selection of state components */
} deb_solout;

/* Options of ExplicitRungeKutta (src/methods/erk/mod.rs:135-144 defaults, :164-228 setters). */
typedef struct deb_erk_options {
    double rtol;            /* scalar Tolerance; default 1e-6 */
    double atol;            /* default 1e-6 */
    const double* rtol_vec; /* Tolerance::Vector(dim) or NULL -- always HOST memory */
    const double* atol_vec; /* idem */
    double h0;              /* 0 = automatic (h_init.rs:45-135 for DP; |tf-t0|/100 for fixed step) */
    double h_min;           /* default 0 */
    double h_max;           /* default +inf */
    int64_t max_steps;      /* default 10000; counts rejected attempts too */
    double safety_factor;   /* default 0.9 */
    double min_scale;       /* default 0.2 */
    double max_scale;       /* default 10 */
    int64_t max_rejects;    /* default 100: consecutive rejections before Error::Stiffness in the adaptive family
                               (adaptive/ordinary.rs:185-197); the Dormand-Prince stepper never reads it */
} deb_erk_options;

typedef struct deb_ode_problem {
    size_t struct_size;
    int32_t system;       /* deb_system */
    int32_t method;       /* deb_method */
    int32_t dim;          /* must equal the system's dimension */
    int32_t n_params;     /* must equal the system's parameter count */
    int64_t n_traj;
    const double* y0;     /* [n_traj][dim] */
    const double* params; /* [n_traj][n_params] in `memspace`; or, when params_shared != 0, ONE set [n_params] in HOST memory */
    int32_t params_shared;
    int32_t n_eval;       /* number of t_eval points (0 = none) */
    const double* t_eval; /* HOST memory always.  Sorted like TEvalSolout::new (t_eval.rs:154-165) */
    double t0, tf;
    deb_erk_options opt;
    int32_t device;       /* CUDA ordinal */
    int32_t memspace;     /* deb_memspace for y0, params and every result pointer */
    void* stream;         /* cudaStream_t when memspace == DEB_MEM_DEVICE (NULL = default stream) */
    int32_t solout;       /* deb_solout; 0 = t_eval */
    int32_t dense_n;      /* DEB_SOLOUT_DENSE: DenseSolout::new(n), n-1 interpolated points per step + the step end */
    double even_dt;       /* DEB_SOLOUT_EVEN: the spacing dt > 0.  t_eval is ignored; n_eval is the row capacity of y_eval per
                             trajectory and must be at least floor(|tf-t0|/dt) + 2.  t_rows receives t0 + k*dt; a trajectory
                             whose t_final == tf has its last row at tf (even.rs:166-188) */
    /* DEB_SOLOUT_CROSSING: CrossingSolout::new(component, threshold).with_direction(d) (src/solout/crossing.rs) */
    int32_t cross_component;
    int32_t cross_direction; /* 0 = Both, +1 = Positive (below -> above), -1 = Negative */
    double cross_threshold;
    /* IVP::event(&e): event detection wrapped around the recorder above (EventWrappedSolout, src/solout/event.rs:300-470).
     * After the recorder has pushed the rows of a step, a sign change of g over the step is located by Brent-Dekker on the
     * dense output and pushed as a row; after `event_terminate` events the trajectory stops with DEB_STATUS_INTERRUPTED.
     * With an event every recorder produces per-trajectory rows: times in t_out, `row_capacity` rows per trajectory. */
    int32_t event;           /* DEB_EVENT_NONE, DEB_EVENT_LINEAR or an id from deb_define_event */
    int32_t event_direction; /* EventConfig.direction: 0 = Both, +1 = Positive, -1 = Negative */
    int32_t event_terminate; /* EventConfig.terminate: stop after this many events; 0 = None (never) */
    int32_t row_capacity;    /* rows per trajectory in y_eval / t_out when an event is set (0 = n_eval); for t_eval / even(dt)
                                recorders n_eval keeps its meaning (number of points / row-plan capacity) */
    double event_coef[DEB_MAX_DIM + 2];
    /* DEB_SOLOUT_HYPERPLANE: the extractor picks components plane_index[0..plane_dim) of the state; rows are pushed where
     * the signed distance sum_i (y[plane_index[i]] - plane_point[i]) * n_i changes sign, n = plane_normal normalised as in
     * HyperplaneCrossingSolout::new; cross_direction filters the direction. */
    int32_t plane_dim;
    int32_t plane_index[DEB_MAX_DIM];
    double plane_point[DEB_MAX_DIM];
    double plane_normal[DEB_MAX_DIM];
    /* ---- appended in ABI 9 (zero = previous behaviour) */
    int32_t filter;       /* deb_filter; adaptive methods only (the fixed-step stepper never calls the hook) */
    int32_t filter_bits;  /* DEB_FILTER_TRUNCATE_MANTISSA: mantissa bits kept, 1..52 */
    int32_t layout;       /* deb_layout of y_eval; DEB_LAYOUT_ROW_MAJOR needs a t_eval / even(dt) recorder without event */
    /* Device list: with n_devices >= 2 the ensemble is split over devices[0..n_devices) inside ONE call (memspace must be
     * DEB_MEM_HOST): blocks of 4096 consecutive trajectories are dealt round-robin (block b -> devices[b mod n_devices],
     * which balances a sorted parameter sweep), one host thread and one persistent kernel per device, no exchange during
     * the integration; results land in the caller's arrays exactly as for one device.  n_devices <= 1: `device` is used. */
    int32_t n_devices;
    int32_t devices[DEB_MAX_DEVICES];
} deb_ode_problem;

typedef struct deb_sde_problem {
    size_t struct_size;
    int32_t system;       /* deb_sde_system */
    int32_t method;       /* fixed-step deb_method; DEB_EULER = Euler-Maruyama */
    int32_t dim;          /* must equal the system's dimension (1 for OU / GBM, 2 for Heston) */
    int32_t n_params;
    int64_t n_traj;
    const double* y0;     /* [n_traj][dim] in `memspace`; or, when y0_shared != 0, ONE state [dim] in HOST memory */
    int32_t y0_shared;
    int32_t params_shared;
    const double* params; /* [n_traj][n_params] in `memspace`; or ONE set in HOST memory when params_shared != 0 */
    int32_t n_eval;
    const double* t_eval; /* HOST */
    double t0, tf;
    deb_erk_options opt;  /* h0, h_min, h_max, max_steps are used */
    /* Wiener increments: dW(path p, step s, component c) = sqrt(h_s) * z, z = Box-Muller normal number
     * (s*dim + c) of the Philox4x32-10 stream with key (seed lo32, seed hi32) and counter
     * (pair index lo32, pair index hi32, path lo32, path hi32); see differential-equations_b200/csrc/philox.h. */
    uint64_t seed;
    int64_t path_offset;  /* global index of path 0 of this call (for sharding across devices) */
    int32_t device;
    int32_t memspace;
    void* stream;
} deb_sde_problem;

/* Result buffers; any pointer may be NULL to skip that output. */
typedef struct deb_result {
    size_t struct_size;
    double* y_eval;      /* [n_traj][n_eval][dim]: rows 0..n_emitted[i]-1 are valid */
    int32_t* n_emitted;  /* [n_traj] number of t_eval rows recorded (Solution.t.len()) */
    double* t_final;     /* [n_traj] t at completion / at the error */
    double* y_final;     /* [n_traj][dim] */
    int32_t* status;     /* [n_traj] deb_status */
    int32_t* accepted;   /* [n_traj] Steps.accepted  (src/stats.rs:69) */
    int32_t* rejected;   /* [n_traj] Steps.rejected */
    int32_t* evals;      /* [n_traj] Evals.function  (src/stats.rs:16) */
    /* filled by the library (HOST memory, always):
     * TEvalSolout sorts the points by direction (t_eval.rs:154-171); the solout call before the loop consumes
     * every point that is not after t0 (a leading point equal to t0 is emitted, the rest are skipped forever,
     * t_eval.rs:121-129).  t_rows[0..n_rows) are the points that remain emittable, in integration order: row r of
     * every trajectory is the state at t_rows[r] (Solution.t = t_rows[0..n_emitted[i])). */
    double* t_rows;        /* [n_eval] caller-provided, may be NULL */
    int32_t n_rows;
    float kernel_ms;       /* device time of the integration kernel(s), DEB_MEM_HOST calls only */
    float total_ms;        /* H2D + kernel + D2H, DEB_MEM_HOST calls only */
    /* Per-step recorders (DEB_SOLOUT_DEFAULT / DENSE / CROSSING): the row times depend on the trajectory.
     * t_out[i][r] is the time of row r of trajectory i (Solution.t), in `memspace` like y_eval; n_eval is the row
     * capacity per trajectory; n_emitted[i] counts every row the reference would have pushed -- rows beyond the
     * capacity are counted but not stored.  n_rows is 0 for these recorders.  May be NULL. */
    double* t_out;         /* [n_traj][n_eval] */
    /* ---- appended in ABI 9 */
    /* Ensemble statistics per t_eval row, reduced on the device(s) while the rows are still resident, then summed across the
     * devices of the call (NCCL all-reduce over NVLink when n_devices >= 2) -- the only cross-device traffic of a solve.
     * HOST memory always, DEB_MEM_HOST calls with a t_eval / even(dt) recorder only; may be NULL.
     * stats_sums[(r*dim + c)*2 + {0,1}] = sum over trajectories with n_emitted > r of {y, y^2}; stats_counts[r] = how many. */
    double* stats_sums;    /* [n_eval][dim][2] */
    int64_t* stats_counts; /* [n_eval] */
    int32_t gpu_launches;  /* filled by the library: kernels this call launched (all devices) */
    int32_t reserved0;
} deb_result;

/* Method-of-lines heat equation u_t = (alpha u_x)_x on a uniform 1-D grid, second-order finite differences,
 * integrated with a fixed-step ERK method; one large state instead of an ensemble. */
typedef struct deb_heat_problem {
    size_t struct_size;
    int64_t n_nodes;     /* grid nodes, StructuredGrid::uniform (src/pde/grid.rs:23-36): dx = (hi-lo)/(n-1) */
    double lo, hi;
    double alpha;        /* flux = alpha * grad u (examples/pde/01_heat_equation/main.rs:18-22) */
    int32_t bc_lower_kind, bc_upper_kind; /* 0 = Dirichlet(value): node derivative is 0 (semi_discrete.rs:264-267); 1 = Neumann(gradient) */
    double bc_lower_value, bc_upper_value;
    int32_t method;      /* fixed-step deb_method (DEB_RK4) */
    double h;            /* step size (rk4(h)) */
    double t0, tf;
    int64_t max_steps;   /* <= 0: the reference default, 10000 (erk/mod.rs:139) */
    const double* u0;    /* [n_nodes] */
    double* u_final;     /* [n_nodes] */
    double* t_final;     /* HOST, 1 value (may be NULL) */
    int64_t* steps;      /* HOST, 1 value: accepted steps taken (may be NULL) */
    int32_t* status;     /* HOST, 1 value (may be NULL) */
    int32_t device;
    int32_t memspace;
    void* stream;
} deb_heat_problem;

int deb_abi_version(void);
const char* deb_last_error(void);
int deb_device_count(void);

/* Fill with the reference defaults (erk/mod.rs:135-144). */
void deb_erk_options_default(deb_erk_options* opt);

/* User-defined right-hand side: the device-side equivalent of `impl ODE for S { fn diff(&self, t, y, dydt) }`
 * (src/ode/ode.rs:20-44).  `diff_body` is the BODY of
 *     void diff(double t, const double* y, double* dydt, const double* p)
 * as CUDA C++ text (p = the trajectory's parameters; write every dydt[0..dim)).  It is compiled at first use with
 * NVRTC for sm_100a with --fmad=false (a*b+c stays two roundings, like Rust) into the same kernel templates that
 * serve the built-in systems, and cached.  Returns a system id (>= 1000) to put in deb_ode_problem.system.
 * A body that does not compile makes the first deb_solve_ode return DEB_ERR_BAD_ARG with the compiler log. */
int deb_define_ode(int32_t dim, int32_t n_params, const char* diff_body, int32_t* system_id);
/* User-defined event function: the BODY of
 *     double event(double t, const double* y, const double* p)
 * as CUDA C++ text (p = the trajectory's ODE parameters), the device-side `impl Event for S { fn event(&self, t, y) }`.
 * Returns an id (>= 1000) for deb_ode_problem.event. */
int deb_define_event(int32_t dim, const char* event_body, int32_t* event_id);
/* User-defined SDE with diagonal noise: the device-side `impl SDE for S { fn drift; fn diffusion; fn noise }`
 * (src/sde/sde.rs:16-67).  The bodies are CUDA C++ text:
 *     void drift(double t, const double* y, double* dydt, const double* p)        -- write dydt[0..dim)
 *     void diffusion(double t, const double* y, double* g, const double* p)       -- write g[0..dim): dY_c += g[c] * dW_c
 *     void noise(double* dw, const double* p)                                      -- optional (NULL = independent increments):
 *         dw[0..dim) arrive as independent N(0, h) increments from the library's Philox stream; mix them in place, as the
 *         reference's `noise` implementations do after drawing (examples/sde/02_heston_model/main.rs:62-72)
 * Returns a system id (>= 1000) for deb_sde_problem.system (n_params <= 8).  Compiled at first use (NVRTC, sm_100a, --fmad=false). */
int deb_define_sde(int32_t dim, int32_t n_params, const char* drift_body, const char* diffusion_body, const char* noise_body,
                   int32_t* system_id);
/* Compile the SDE kernel for (system, method) now, without a device (like deb_check_ode). */
int deb_check_sde(int32_t system_id, int32_t method);
/* Forward sensitivity analysis, ForwardSensitivityOde (src/ode/sensitivity/forward.rs:43-116): builds the augmented system
 *     z = [y (n = dim), S (n x m, row-major: S[r][c] = dy_r/dp_c at z[n + r*m + c])],  m = n_params,
 *     y' = f(t, y, p),   S' = J_y S + J_p
 * from the text of `ODE::diff`, `ODE::jacobian` and `ParametrizedODE::jacobian_p` and registers it like deb_define_ode
 * (state dimension n*(1+m) <= DEB_MAX_DIM):
 *     diff_body:  body of void diff(double t, const double* y, double* dydt, const double* p)
 *     jac_y_body: body of void jacobian(double t, const double* y, double* J, const double* p)     -- J[r*n + k] = df_r/dy_k
 *     jac_p_body: body of void jacobian_p(double t, const double* y, double* Jp, const double* p)  -- Jp[r*m + c] = df_r/dp_c
 * J and Jp arrive zeroed (the reference's Matrix::full).  Each entry is accumulated as the reference does (forward.rs:103-111):
 * dS_rc = Jp[r][c]; for k in 0..n: dS_rc += J[r][k] * S[k][c].  Initial state: [y0, S(t0)] (S(t0) = 0 for parameters that do
 * not enter y0). */
int deb_define_ode_sensitivity(int32_t dim, int32_t n_params, const char* diff_body, const char* jac_y_body,
                               const char* jac_p_body, int32_t* system_id);
/* Compile the kernel for (system, method, recorder, event) now, without a device and without running anything: DEB_OK, or
 * an error with the compiler log in deb_last_error() (DEB_ERR_BAD_ARG for a user-defined right-hand side that does not
 * compile: what a Rust caller gets from `cargo check`).  Kernels that were compiled ahead of time (built-in system with
 * a t_eval / even(dt) recorder) return DEB_OK at once. */
int deb_check_ode(int32_t system_id, int32_t method, int32_t solout, int32_t event);

/* Device memory the library caches between calls (staging buffers of DEB_MEM_HOST calls, work buffers) lives in a
 * library-owned stream-ordered pool; this releases it back to the driver.  Synchronises the device. */
int deb_trim_memory(int32_t device);

/* Number of CUDA kernels the library has launched in this process so far (all entry points, all devices). */
int64_t deb_launch_count(void);

int deb_solve_ode(const deb_ode_problem* problem, deb_result* result);
int deb_solve_sde(const deb_sde_problem* problem, deb_result* result);
int deb_solve_heat_mol(const deb_heat_problem* problem);
/* One evaluation of the semi-discrete right-hand side (SemiDiscretePde::diff, src/pde/semi_discrete.rs:250-287)
 * for the grid / boundary description in `problem` (its u0, u_final, method, h, t0, tf are ignored):
 * du[i] = f(u)[i].  u/du live in problem->memspace.  Mirrors `system.diff(t, &y, &mut dydt)` of
 * tests/pde/method_of_lines.rs:73-157. */
int deb_heat_rhs(const deb_heat_problem* problem, const double* u, double* du);

/* sums[(r*dim + c)*2 + {0,1}] = sum over trajectories with n_emitted > r of {y, y^2} at row r;
 * counts[r] = number of contributing trajectories.  y_eval/n_emitted in `memspace`; sums/counts likewise. */
int deb_ensemble_stats(const double* y_eval, const int32_t* n_emitted, int64_t n_traj, int32_t n_eval, int32_t dim,
                       double* sums, int64_t* counts, int32_t device, int32_t memspace, void* stream);

/* Device-memory helpers so that a caller without the CUDA runtime can keep an ensemble resident. */
int deb_malloc(int32_t device, size_t bytes, void** ptr);
int deb_free(int32_t device, void* ptr);
int deb_memcpy_h2d(int32_t device, void* dst_dev, const void* src_host, size_t bytes);
int deb_memcpy_d2h(int32_t device, void* dst_host, const void* src_dev, size_t bytes);
int deb_synchronize(int32_t device);

/* Diagnostics.
 * deb_pow_device: out[i] = the controller's pow(x[i], y) evaluated on the device (HOST pointers); used to
 *   prove the device port of glibc pow returns libm's bits.
 * deb_fp64_issue_peak: runs a register-only stream of independent DADD/DMUL (no FMA fusion, like the
 *   reference arithmetic) on every SM and reports DP instructions per second -- the roofline denominator
 *   for the compute-bound ensemble kernels (MEASURED_PEAKS.json has no FP64 entry). */
int deb_pow_device(const double* x, double y, int64_t n, double* out, int32_t device);
int deb_fp64_issue_peak(int32_t device, int32_t use_fma, double* dp_inst_per_s, float* ms);
/* deb_plan_fixed_steps: the step schedule the fixed-step and SDE kernels follow, planned on the host exactly as the reference's loop
 *   produces it (solve_ivp.rs:193-209, :263; fixed/ordinary.rs:16-56): n_steps steps, all of size h0 (|tf - t0| / 100 when h0 == 0)
 *   except the last n_tail (1..4), whose sizes are h_tail[0..n_tail) -- the clip at tf can fire more than once.  status is
 *   DEB_STATUS_COMPLETE, DEB_STATUS_MAX_STEPS or DEB_STATUS_BAD_INPUT (then n_steps = 0).  Host only: no device is touched. */
/* deb_shard_layout: how a call with a device list splits an ensemble (blocks of 2^shift consecutive trajectories, block b on
 *   devices[b mod n_devices]; shift = 12 unless DEB_WM_SHIFT says otherwise): the number of blocks and of trajectories that device
 *   number `index` of the list integrates, and the global number of its local block `local_block` (-1 when it has no such block).
 *   Host only. */
int deb_shard_layout(int64_t n_traj, int32_t n_devices, int32_t index, int32_t shift, int64_t local_block, int64_t* n_local_blocks,
                     int64_t* n_local_traj, int64_t* global_block);
int deb_plan_fixed_steps(double t0, double tf, double h0, double h_min, double h_max, int64_t max_steps, int64_t* n_steps, int32_t* n_tail,
                         double* h_tail, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* DEB_ENSEMBLE_H */
