//! Ensemble front end: `EnsembleIVP::ode(system, t0, tf, y0s).t_eval(pts).method(m).solve()` mirrors
//! `IVP::ode(&sys, t0, tf, y0).t_eval(pts).method(m).solve()` (differential-equations `src/ivp.rs:279,656,632,781`)
//! for N independent problems, executed by the sm_100a kernels behind `include/deb_ensemble.h`.
//!
//! SOURCE ONLY: never compiled (no Rust toolchain in the build image).  The `#[repr(C)]` structs below are a
//! field-for-field transcription of `include/deb_ensemble.h` (ABI version 9); `tests/test_abi_cpu.py` checks the
//! same layout for the Python mirror against the compiled header.
//!
//! What is wrapped (reference call -> method here):
//!   IVP::ode(..).t_eval / .even / plain solve() / .dense / .crossing / .hyperplane_crossing   src/ivp.rs:637-705
//!       -> EnsembleIVP::{t_eval, even, every_step, dense, crossing, hyperplane_crossing}
//!   .event(&e)   src/ivp.rs:662       -> EnsembleIVP::event(Event::Linear{..} | Event::Source(id), direction, terminate)
//!   IVP::ode_from_fn(..)  src/ivp.rs:305  -> System::from_source(dim, n_params, "CUDA C++ body of diff")
//!   ForwardSensitivityOde::new  src/ode/sensitivity/forward.rs:58  -> System::sensitivity_from_source(..)
//!   ExplicitRungeKutta::{..}.filter(f)  src/methods/erk/mod.rs:225 -> Method::filter_truncate_mantissa(bits)
//!   IVP::sde(..)  src/ivp.rs:504,857  -> EnsembleSDE::sde(SdeSystem, ..).method(Method::euler(h) | Method::milstein(h)).solve()
//!   IVP::pde(..).space(MethodOfLines::finite_difference(grid).boundary(bc))  src/ivp.rs:419,713 -> solve_heat_mol(..)
//!   (new) several GPUs behind one call: EnsembleIVP::devices(&[0, 1, ..]); fused ensemble statistics: .with_stats()
#![allow(non_camel_case_types)]

use std::ffi::{c_void, CStr, CString};
use std::os::raw::c_char;

use differential_equations::{
    error::Error,
    solution::Solution,
    stats::{Evals, Steps},
    status::Status,
};

// ------------------------------------------------------------------------------------------------ raw ABI
pub const DEB_ABI_VERSION: i32 = 9;
pub const DEB_MAX_DIM: usize = 16;
pub const DEB_MAX_DEVICES: usize = 16;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct deb_erk_options {
    pub rtol: f64,
    pub atol: f64,
    pub rtol_vec: *const f64,
    pub atol_vec: *const f64,
    pub h0: f64,
    pub h_min: f64,
    pub h_max: f64,
    pub max_steps: i64,
    pub safety_factor: f64,
    pub min_scale: f64,
    pub max_scale: f64,
    pub max_rejects: i64,
}

#[repr(C)]
pub struct deb_ode_problem {
    pub struct_size: usize,
    pub system: i32,
    pub method: i32,
    pub dim: i32,
    pub n_params: i32,
    pub n_traj: i64,
    pub y0: *const f64,
    pub params: *const f64,
    pub params_shared: i32,
    pub n_eval: i32,
    pub t_eval: *const f64,
    pub t0: f64,
    pub tf: f64,
    pub opt: deb_erk_options,
    pub device: i32,
    pub memspace: i32,
    pub stream: *mut c_void,
    pub solout: i32, // 0 = t_eval, 1 = even(dt), 2 = every step (DefaultSolout), 3 = dense(n), 4 = crossing, 5 = hyperplane_crossing
    pub dense_n: i32,
    pub even_dt: f64,
    pub cross_component: i32,
    pub cross_direction: i32, // 0 Both, +1 Positive, -1 Negative
    pub cross_threshold: f64,
    pub event: i32,           // 0 none, 1 linear (event_coef), >= 1000 from deb_define_event
    pub event_direction: i32, // EventConfig.direction
    pub event_terminate: i32, // EventConfig.terminate (0 = None)
    pub row_capacity: i32,    // rows per trajectory when an event is set
    pub event_coef: [f64; DEB_MAX_DIM + 2],
    pub plane_dim: i32, // hyperplane_crossing: number of extracted components
    pub plane_index: [i32; DEB_MAX_DIM],
    pub plane_point: [f64; DEB_MAX_DIM],
    pub plane_normal: [f64; DEB_MAX_DIM],
    // ---- ABI 9
    pub filter: i32,      // 0 identity, 1 truncate mantissa
    pub filter_bits: i32, // mantissa bits kept (1..52)
    pub layout: i32,      // 0 y_eval[i][r][c], 1 y_eval[r][c][i]
    pub n_devices: i32,   // >= 2: split over devices[..] inside one call (memspace HOST)
    pub devices: [i32; DEB_MAX_DEVICES],
}

#[repr(C)]
pub struct deb_sde_problem {
    pub struct_size: usize,
    pub system: i32, // 0 OU, 1 GBM, 2 Heston, >= 1000 from deb_define_sde
    pub method: i32,
    pub dim: i32,
    pub n_params: i32,
    pub n_traj: i64,
    pub y0: *const f64,
    pub y0_shared: i32,
    pub params_shared: i32,
    pub params: *const f64,
    pub n_eval: i32,
    pub t_eval: *const f64,
    pub t0: f64,
    pub tf: f64,
    pub opt: deb_erk_options,
    pub seed: u64,
    pub path_offset: i64,
    pub device: i32,
    pub memspace: i32,
    pub stream: *mut c_void,
}

#[repr(C)]
pub struct deb_result {
    pub struct_size: usize,
    pub y_eval: *mut f64,
    pub n_emitted: *mut i32,
    pub t_final: *mut f64,
    pub y_final: *mut f64,
    pub status: *mut i32,
    pub accepted: *mut i32,
    pub rejected: *mut i32,
    pub evals: *mut i32,
    pub t_rows: *mut f64,
    pub n_rows: i32,
    pub kernel_ms: f32,
    pub total_ms: f32,
    pub t_out: *mut f64, // [n_traj][n_eval] row times of the per-step recorders, or null
    // ---- ABI 9
    pub stats_sums: *mut f64,   // [n_eval][dim][2] HOST, or null
    pub stats_counts: *mut i64, // [n_eval] HOST, or null
    pub gpu_launches: i32,
    pub reserved0: i32,
}

#[repr(C)]
pub struct deb_heat_problem {
    pub struct_size: usize,
    pub n_nodes: i64,
    pub lo: f64,
    pub hi: f64,
    pub alpha: f64,
    pub bc_lower_kind: i32, // 0 Dirichlet, 1 Neumann
    pub bc_upper_kind: i32,
    pub bc_lower_value: f64,
    pub bc_upper_value: f64,
    pub method: i32,
    pub h: f64,
    pub t0: f64,
    pub tf: f64,
    pub max_steps: i64, // <= 0: 10000
    pub u0: *const f64,
    pub u_final: *mut f64,
    pub t_final: *mut f64,
    pub steps: *mut i64,
    pub status: *mut i32,
    pub device: i32,
    pub memspace: i32,
    pub stream: *mut c_void,
}

extern "C" {
    pub fn deb_abi_version() -> i32;
    pub fn deb_last_error() -> *const c_char;
    pub fn deb_device_count() -> i32;
    pub fn deb_launch_count() -> i64;
    /// host only: how a device list splits an ensemble (block b of 2^shift trajectories on devices[b mod n_devices])
    pub fn deb_shard_layout(n_traj: i64, n_devices: i32, index: i32, shift: i32, local_block: i64, n_local_blocks: *mut i64,
                            n_local_traj: *mut i64, global_block: *mut i64) -> i32;
    /// host only: the step schedule of the fixed-step / SDE kernels (n_steps steps of h0, the last n_tail with sizes h_tail[..n_tail])
    pub fn deb_plan_fixed_steps(t0: f64, tf: f64, h0: f64, h_min: f64, h_max: f64, max_steps: i64, n_steps: *mut i64, n_tail: *mut i32,
                                h_tail: *mut f64, status: *mut i32) -> i32;
    pub fn deb_erk_options_default(opt: *mut deb_erk_options);
    pub fn deb_solve_ode(problem: *const deb_ode_problem, result: *mut deb_result) -> i32;
    pub fn deb_solve_sde(problem: *const deb_sde_problem, result: *mut deb_result) -> i32;
    pub fn deb_solve_heat_mol(problem: *const deb_heat_problem) -> i32;
    pub fn deb_heat_rhs(problem: *const deb_heat_problem, u: *const f64, du: *mut f64) -> i32;
    /// release the device memory the library caches between calls
    pub fn deb_trim_memory(device: i32) -> i32;
    /// user-defined right-hand side as CUDA C++ text (the device-side `impl ODE`); returns a system id >= 1000
    pub fn deb_define_ode(dim: i32, n_params: i32, diff_body: *const c_char, system_id: *mut i32) -> i32;
    /// forward-sensitivity system [y, S], S' = J_y S + J_p, generated from diff / jacobian / jacobian_p bodies
    pub fn deb_define_ode_sensitivity(dim: i32, n_params: i32, diff_body: *const c_char, jac_y_body: *const c_char, jac_p_body: *const c_char,
                                      system_id: *mut i32) -> i32;
    /// compile a kernel now (no device needed); the compiler log is in deb_last_error()
    pub fn deb_check_ode(system_id: i32, method: i32, solout: i32, event: i32) -> i32;
    /// user-defined event function g(t, y) as CUDA C++ text (the device-side `impl Event`); returns an event id >= 1000
    pub fn deb_define_event(dim: i32, event_body: *const c_char, event_id: *mut i32) -> i32;
    /// user-defined SDE: drift / diffusion / noise-mixing bodies (the device-side `impl SDE`); returns a system id >= 1000
    pub fn deb_define_sde(dim: i32, n_params: i32, drift_body: *const c_char, diffusion_body: *const c_char, noise_body: *const c_char,
                          system_id: *mut i32) -> i32;
    pub fn deb_check_sde(system_id: i32, method: i32) -> i32;
    pub fn deb_ensemble_stats(y_eval: *const f64, n_emitted: *const i32, n_traj: i64, n_eval: i32, dim: i32, sums: *mut f64, counts: *mut i64,
                              device: i32, memspace: i32, stream: *mut c_void) -> i32;
}

fn last_error(what: &str, rc: i32) -> String {
    let msg = unsafe { CStr::from_ptr(deb_last_error()) }.to_string_lossy().into_owned();
    format!("{what} failed ({rc}): {msg}") // includes DEB_ERR_NO_DEVICE: there is no CPU fallback
}

// ------------------------------------------------------------------------------------------------ systems
/// A right-hand side the device can run: one of the crate's test systems (compiled in), or CUDA C++ text.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct System {
    pub id: i32,
    pub dim: usize,
    pub n_params: usize,
}

impl System {
    pub const EXPONENTIAL: System = System { id: 0, dim: 1, n_params: 1 }; // y' = k y
    pub const LINEAR: System = System { id: 1, dim: 1, n_params: 2 };
    pub const HARMONIC: System = System { id: 2, dim: 2, n_params: 1 };
    pub const LOGISTIC: System = System { id: 3, dim: 1, n_params: 2 };
    pub const VAN_DER_POL: System = System { id: 4, dim: 2, n_params: 1 };
    pub const LORENZ: System = System { id: 5, dim: 3, n_params: 3 };
    pub const BRUSSELATOR: System = System { id: 6, dim: 2, n_params: 2 };
    pub const ROBERTSON: System = System { id: 7, dim: 3, n_params: 0 };

    /// `IVP::ode_from_fn(|t, y, dydt| ..)` for the device: the body of `void diff(double t, const double* y, double* dydt, const double* p)`.
    pub fn from_source(dim: usize, n_params: usize, diff_body: &str) -> Result<System, String> {
        let body = CString::new(diff_body).map_err(|e| e.to_string())?;
        let mut id = -1;
        let rc = unsafe { deb_define_ode(dim as i32, n_params as i32, body.as_ptr(), &mut id) };
        if rc != 0 { return Err(last_error("deb_define_ode", rc)); }
        Ok(System { id, dim, n_params })
    }

    /// `ForwardSensitivityOde::new(ode, y_proto)`: z = [y, S (row-major, dim x n_params)], S' = J_y S + J_p.
    pub fn sensitivity_from_source(dim: usize, n_params: usize, diff_body: &str, jacobian_body: &str, jacobian_p_body: &str) -> Result<System, String> {
        let (d, j, jp) = (CString::new(diff_body).unwrap(), CString::new(jacobian_body).unwrap(), CString::new(jacobian_p_body).unwrap());
        let mut id = -1;
        let rc = unsafe { deb_define_ode_sensitivity(dim as i32, n_params as i32, d.as_ptr(), j.as_ptr(), jp.as_ptr(), &mut id) };
        if rc != 0 { return Err(last_error("deb_define_ode_sensitivity", rc)); }
        Ok(System { id, dim: dim * (1 + n_params), n_params })
    }
}

/// `impl Event for S { fn event(&self, t, y) -> T }` (src/solout/event.rs:60-70) in a form that can cross the boundary.
#[derive(Clone, Debug)]
pub enum Event {
    /// g(t, y) = c0 + ct * t + sum_i cy[i] * y[i]
    Linear { c0: f64, ct: f64, cy: Vec<f64> },
    /// an id returned by `define_event`
    Source(i32),
}

/// The body of `double event(double t, const double* y, const double* p)` as CUDA C++ text.
pub fn define_event(dim: usize, body: &str) -> Result<Event, String> {
    let b = CString::new(body).map_err(|e| e.to_string())?;
    let mut id = -1;
    let rc = unsafe { deb_define_event(dim as i32, b.as_ptr(), &mut id) };
    if rc != 0 { return Err(last_error("deb_define_event", rc)); }
    Ok(Event::Source(id))
}

/// `CrossingDirection` (src/solout/mod.rs) / `EventConfig.direction`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Direction { Both = 0, Positive = 1, Negative = -1 }

// ------------------------------------------------------------------------------------------------ methods
/// `ExplicitRungeKutta::dopri5()/dop853()/rk4(h)/...` with the crate's setters (`src/methods/erk/mod.rs:164-228`).
#[derive(Clone)]
pub struct Method {
    id: i32,
    opt: deb_erk_options,
    rtol_vec: Option<Vec<f64>>,
    atol_vec: Option<Vec<f64>>,
    filter_bits: i32, // 0 = identity
}

impl Method {
    fn new(id: i32, h0: f64) -> Self {
        let mut opt = unsafe { std::mem::zeroed::<deb_erk_options>() };
        unsafe { deb_erk_options_default(&mut opt) }; // rtol = atol = 1e-6, max_steps = 10_000, safety 0.9, scale in [0.2, 10]
        opt.h0 = h0;
        Method { id, opt, rtol_vec: None, atol_vec: None, filter_bits: 0 }
    }
    pub fn euler(h: f64) -> Self { Self::new(0, h) }
    pub fn midpoint(h: f64) -> Self { Self::new(1, h) }
    pub fn heun(h: f64) -> Self { Self::new(2, h) }
    pub fn ralston(h: f64) -> Self { Self::new(3, h) }
    pub fn ssp_rk3(h: f64) -> Self { Self::new(4, h) }
    pub fn rk4(h: f64) -> Self { Self::new(5, h) }
    pub fn three_eighths(h: f64) -> Self { Self::new(6, h) }
    pub fn dopri5() -> Self { Self::new(16, 0.0) }
    pub fn dop853() -> Self { Self::new(17, 0.0) }
    pub fn rkf45() -> Self { Self::new(18, 0.0) }
    pub fn cash_karp() -> Self { Self::new(19, 0.0) }
    pub fn rkv655e() -> Self { Self::new(20, 0.0) }
    pub fn rkv656e() -> Self { Self::new(21, 0.0) }
    pub fn rkv766e() -> Self { Self::new(22, 0.0) }
    pub fn rkv767e() -> Self { Self::new(23, 0.0) }
    pub fn rkv877e() -> Self { Self::new(24, 0.0) }
    pub fn rkv878e() -> Self { Self::new(25, 0.0) }
    pub fn rkv988e() -> Self { Self::new(26, 0.0) }
    pub fn rkv989e() -> Self { Self::new(27, 0.0) }
    /// `Milstein::new(h)` (src/methods/milstein.rs:37-68), SDE ensembles only
    pub fn milstein(h: f64) -> Self { Self::new(32, h) }
    pub fn rtol(mut self, v: f64) -> Self { self.opt.rtol = v; self.rtol_vec = None; self }
    pub fn atol(mut self, v: f64) -> Self { self.opt.atol = v; self.atol_vec = None; self }
    pub fn rtol_vec(mut self, v: Vec<f64>) -> Self { self.rtol_vec = Some(v); self }
    pub fn atol_vec(mut self, v: Vec<f64>) -> Self { self.atol_vec = Some(v); self }
    pub fn h0(mut self, v: f64) -> Self { self.opt.h0 = v; self }
    pub fn h_min(mut self, v: f64) -> Self { self.opt.h_min = v; self }
    pub fn h_max(mut self, v: f64) -> Self { self.opt.h_max = v; self }
    pub fn max_steps(mut self, v: usize) -> Self { self.opt.max_steps = v as i64; self }
    pub fn safety_factor(mut self, v: f64) -> Self { self.opt.safety_factor = v; self }
    pub fn min_scale(mut self, v: f64) -> Self { self.opt.min_scale = v; self }
    pub fn max_scale(mut self, v: f64) -> Self { self.opt.max_scale = v; self }
    pub fn max_rejects(mut self, v: usize) -> Self { self.opt.max_rejects = v as i64; self }
    /// `.filter(|h| f64::from_bits(h.to_bits() & MASK))` with MASK keeping the leading `bits` mantissa bits: the one
    /// `filter: fn(T) -> T` (src/methods/erk/mod.rs:225) that can cross the C ABI.
    pub fn filter_truncate_mantissa(mut self, bits: u32) -> Self { self.filter_bits = bits as i32; self }
    fn options(&self) -> deb_erk_options {
        let mut opt = self.opt;
        opt.rtol_vec = self.rtol_vec.as_ref().map_or(std::ptr::null(), |v| v.as_ptr());
        opt.atol_vec = self.atol_vec.as_ref().map_or(std::ptr::null(), |v| v.as_ptr());
        opt
    }
}

// ------------------------------------------------------------------------------------------------ ODE ensembles
#[derive(Clone, Debug)]
enum Recorder {
    TEval(Vec<f64>),
    Even(f64),
    EveryStep { max_rows: usize },
    Dense { n: usize, max_rows: usize },
    Crossing { component: usize, threshold: f64, direction: Direction, max_rows: usize },
    Hyperplane { point: Vec<f64>, normal: Vec<f64>, components: Vec<usize>, direction: Direction, max_rows: usize },
}

/// Per-t_eval ensemble sums reduced on the device(s): `sums[(r*dim + c)*2 + {0,1}]` = sum of {y, y^2}, `counts[r]`.
pub struct EnsembleStats { pub sums: Vec<f64>, pub counts: Vec<i64> }

pub struct EnsembleOutput<const N: usize> {
    /// exactly what N calls of `IVP::solve()` would have returned (timer excluded)
    pub solutions: Vec<Result<Solution<f64, [f64; N]>, Error<f64, [f64; N]>>>,
    pub stats: Option<EnsembleStats>,
    pub kernel_ms: f32,
    pub gpu_launches: i32,
}

/// N independent IVPs sharing (t0, tf, method): the ensemble analogue of `IVP` (`src/ivp.rs`).
pub struct EnsembleIVP<const N: usize> {
    system: System,
    params: Vec<f64>, // one set (shared) or N_traj sets, flattened
    params_shared: bool,
    t0: f64,
    tf: f64,
    y0: Vec<[f64; N]>,
    recorder: Recorder,
    event: Option<(Event, Direction, u32, usize)>, // (event, direction, terminate after, extra row slots)
    method: Option<Method>,
    devices: Vec<i32>,
    with_stats: bool,
}

impl<const N: usize> EnsembleIVP<N> {
    /// `IVP::ode(&sys, t0, tf, y0)`; `params` holds one parameter set for all trajectories or one per trajectory.
    pub fn ode(system: System, params: Vec<f64>, t0: f64, tf: f64, y0: Vec<[f64; N]>) -> Self {
        assert_eq!(system.dim, N, "state dimension must match the system");
        let params_shared = params.len() == system.n_params;
        assert!(params_shared || params.len() == system.n_params * y0.len(), "params: one set, or one per trajectory");
        EnsembleIVP { system, params, params_shared, t0, tf, y0, recorder: Recorder::TEval(vec![]), event: None, method: None,
                      devices: vec![0], with_stats: false }
    }
    pub fn t_eval(mut self, pts: impl AsRef<[f64]>) -> Self { self.recorder = Recorder::TEval(pts.as_ref().to_vec()); self } // ivp.rs:656
    pub fn even(mut self, dt: f64) -> Self { self.recorder = Recorder::Even(dt); self }                                    // ivp.rs:643
    /// the recorder of a plain `solve()` (DefaultSolout): (t0, y0) and every accepted step; rows beyond `max_rows` are counted, not stored
    pub fn every_step(mut self, max_rows: usize) -> Self { self.recorder = Recorder::EveryStep { max_rows }; self }
    pub fn dense(mut self, n: usize, max_rows: usize) -> Self { self.recorder = Recorder::Dense { n, max_rows }; self }     // ivp.rs:650
    pub fn crossing(mut self, component: usize, threshold: f64, direction: Direction, max_rows: usize) -> Self {             // ivp.rs:682
        self.recorder = Recorder::Crossing { component, threshold, direction, max_rows }; self
    }
    /// `hyperplane_crossing(point, normal, extractor, direction)` with the extractor = a selection of state components (ivp.rs:695)
    pub fn hyperplane_crossing(mut self, point: Vec<f64>, normal: Vec<f64>, components: Vec<usize>, direction: Direction, max_rows: usize) -> Self {
        assert!(point.len() == normal.len() && normal.len() == components.len() && !components.is_empty());
        self.recorder = Recorder::Hyperplane { point, normal, components, direction, max_rows }; self
    }
    /// `.event(&e)` with `EventConfig { direction, terminate }` (ivp.rs:662); `terminate = 1` is `.terminal()`, 0 = never
    pub fn event(mut self, e: Event, direction: Direction, terminate: u32, max_event_rows: usize) -> Self {
        self.event = Some((e, direction, terminate, max_event_rows)); self
    }
    pub fn method(mut self, m: Method) -> Self { self.method = Some(m); self }
    pub fn device(mut self, ordinal: i32) -> Self { self.devices = vec![ordinal]; self }
    /// several GPUs behind this one call: blocks of 4096 trajectories dealt round-robin, statistics all-reduced with NCCL
    pub fn devices(mut self, ordinals: &[i32]) -> Self { self.devices = ordinals.to_vec(); self }
    pub fn with_stats(mut self) -> Self { self.with_stats = true; self }

    pub fn solve(self) -> Result<EnsembleOutput<N>, String> {
        let m = self.method.clone().expect("method(..) must be set");
        let n = self.y0.len();
        // rows per trajectory and the recorder fields of the problem
        let (solout, mut rows_cap, t_eval): (i32, usize, Vec<f64>) = match &self.recorder {
            Recorder::TEval(p) => (0, p.len(), p.clone()),
            Recorder::Even(dt) => (1, ((self.tf - self.t0).abs() / dt).floor() as usize + 3, vec![]),
            Recorder::EveryStep { max_rows } => (2, *max_rows, vec![]),
            Recorder::Dense { max_rows, .. } => (3, *max_rows, vec![]),
            Recorder::Crossing { max_rows, .. } => (4, *max_rows, vec![]),
            Recorder::Hyperplane { max_rows, .. } => (5, *max_rows, vec![]),
        };
        let n_eval = rows_cap;
        let per_traj_times = solout >= 2 || self.event.is_some();
        if let Some((_, _, _, extra)) = &self.event { if solout < 2 { rows_cap += *extra; } }
        let mut y_eval = vec![0.0f64; n * rows_cap * N];
        let mut t_out = vec![0.0f64; if per_traj_times { n * rows_cap } else { 0 }];
        let mut n_emitted = vec![0i32; n];
        let mut t_final = vec![0.0f64; n];
        let mut y_final = vec![[0.0f64; N]; n];
        let (mut status, mut acc, mut rej, mut evals) = (vec![0i32; n], vec![0i32; n], vec![0i32; n], vec![0i32; n]);
        let mut t_rows = vec![0.0f64; n_eval.max(1)];
        let mut stats = if self.with_stats { Some(EnsembleStats { sums: vec![0.0; n_eval * N * 2], counts: vec![0; n_eval] }) } else { None };
        let mut problem: deb_ode_problem = unsafe { std::mem::zeroed() };
        problem.struct_size = std::mem::size_of::<deb_ode_problem>();
        problem.system = self.system.id;
        problem.method = m.id;
        problem.dim = N as i32;
        problem.n_params = self.system.n_params as i32;
        problem.n_traj = n as i64;
        problem.y0 = self.y0.as_ptr() as *const f64; // Vec<[f64; N]> is the ABI's "array of states" layout
        problem.params = self.params.as_ptr();
        problem.params_shared = self.params_shared as i32;
        problem.n_eval = n_eval as i32;
        problem.t_eval = if t_eval.is_empty() { std::ptr::null() } else { t_eval.as_ptr() };
        problem.t0 = self.t0;
        problem.tf = self.tf;
        problem.opt = m.options();
        problem.device = self.devices[0];
        problem.memspace = 0; // DEB_MEM_HOST
        problem.solout = solout;
        match &self.recorder {
            Recorder::Even(dt) => problem.even_dt = *dt,
            Recorder::Dense { n, .. } => problem.dense_n = *n as i32,
            Recorder::Crossing { component, threshold, direction, .. } => {
                problem.cross_component = *component as i32; problem.cross_threshold = *threshold; problem.cross_direction = *direction as i32;
            }
            Recorder::Hyperplane { point, normal, components, direction, .. } => {
                problem.plane_dim = components.len() as i32; problem.cross_direction = *direction as i32;
                for q in 0..components.len() { problem.plane_index[q] = components[q] as i32; problem.plane_point[q] = point[q]; problem.plane_normal[q] = normal[q]; }
            }
            _ => {}
        }
        if let Some((e, direction, terminate, _)) = &self.event {
            problem.event_direction = *direction as i32;
            problem.event_terminate = *terminate as i32;
            problem.row_capacity = rows_cap as i32;
            match e {
                Event::Linear { c0, ct, cy } => {
                    problem.event = 1; problem.event_coef[0] = *c0; problem.event_coef[1] = *ct;
                    for (q, v) in cy.iter().enumerate() { problem.event_coef[2 + q] = *v; }
                }
                Event::Source(id) => problem.event = *id,
            }
        }
        if m.filter_bits > 0 { problem.filter = 1; problem.filter_bits = m.filter_bits; }
        if self.devices.len() > 1 {
            problem.n_devices = self.devices.len() as i32;
            for (q, d) in self.devices.iter().enumerate() { problem.devices[q] = *d; }
        }
        let mut result: deb_result = unsafe { std::mem::zeroed() };
        result.struct_size = std::mem::size_of::<deb_result>();
        result.y_eval = y_eval.as_mut_ptr();
        result.n_emitted = n_emitted.as_mut_ptr();
        result.t_final = t_final.as_mut_ptr();
        result.y_final = y_final.as_mut_ptr() as *mut f64;
        result.status = status.as_mut_ptr();
        result.accepted = acc.as_mut_ptr();
        result.rejected = rej.as_mut_ptr();
        result.evals = evals.as_mut_ptr();
        result.t_rows = t_rows.as_mut_ptr();
        result.t_out = if per_traj_times { t_out.as_mut_ptr() } else { std::ptr::null_mut() };
        if let Some(s) = stats.as_mut() { result.stats_sums = s.sums.as_mut_ptr(); result.stats_counts = s.counts.as_mut_ptr(); }
        let rc = unsafe { deb_solve_ode(&problem, &mut result) };
        if rc != 0 { return Err(last_error("deb_solve_ode", rc)); }
        let even_tf = matches!(self.recorder, Recorder::Even(_));
        let mut out = Vec::with_capacity(n);
        for i in 0..n {
            let (t, y) = (t_final[i], y_final[i]);
            out.push(match status[i] {
                0 | 5 => {
                    let mut s = Solution::new();
                    let rows = (n_emitted[i] as usize).min(rows_cap);
                    for r in 0..rows {
                        let mut row = [0.0; N];
                        row.copy_from_slice(&y_eval[(i * rows_cap + r) * N..(i * rows_cap + r + 1) * N]);
                        // EvenSolout: a trajectory that lands exactly on tf has its last row at tf (even.rs:166-188)
                        let tr = if per_traj_times { t_out[i * rows_cap + r] } else if even_tf && r + 1 == rows && t == self.tf { self.tf } else { t_rows[r] };
                        s.push(tr, row);
                    }
                    s.status = if status[i] == 5 { Status::Interrupted } else { Status::Complete };
                    s.evals = Evals { function: evals[i] as usize, ..Evals::new() };
                    s.steps = Steps { accepted: acc[i] as usize, rejected: rej[i] as usize };
                    Ok(s)
                }
                1 => Err(Error::MaxSteps { t, y }),
                2 => Err(Error::StepSize { t, y }),
                3 => Err(Error::Stiffness { t, y }),
                _ => Err(Error::BadInput { msg: "Invalid input".to_string() }),
            });
        }
        Ok(EnsembleOutput { solutions: out, stats, kernel_ms: result.kernel_ms, gpu_launches: result.gpu_launches })
    }
}

// ------------------------------------------------------------------------------------------------ SDE ensembles
/// `impl SDE` the device can run: OU / GBM / Heston compiled in, or drift / diffusion / noise bodies as CUDA C++ text.
#[derive(Clone, Copy, Debug)]
pub struct SdeSystem { pub id: i32, pub dim: usize, pub n_params: usize }

impl SdeSystem {
    pub const ORNSTEIN_UHLENBECK: SdeSystem = SdeSystem { id: 0, dim: 1, n_params: 3 }; // {theta, mu, sigma}
    pub const GEOMETRIC_BROWNIAN_MOTION: SdeSystem = SdeSystem { id: 1, dim: 1, n_params: 2 }; // {mu, sigma}
    pub const HESTON: SdeSystem = SdeSystem { id: 2, dim: 2, n_params: 5 }; // {mu, kappa, theta, sigma, rho}
    pub fn from_source(dim: usize, n_params: usize, drift: &str, diffusion: &str, noise: Option<&str>) -> Result<SdeSystem, String> {
        let (d, g) = (CString::new(drift).unwrap(), CString::new(diffusion).unwrap());
        let nz = noise.map(|s| CString::new(s).unwrap());
        let mut id = -1;
        let rc = unsafe { deb_define_sde(dim as i32, n_params as i32, d.as_ptr(), g.as_ptr(), nz.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()), &mut id) };
        if rc != 0 { return Err(last_error("deb_define_sde", rc)); }
        Ok(SdeSystem { id, dim, n_params })
    }
}

/// `IVP::sde(&mut sde, t0, tf, y0).t_eval(pts).method(ExplicitRungeKutta::euler(h)).solve()` for N paths.  `SDE::noise` is the
/// library's counter-based Philox4x32-10 stream: path p, step s, component c is reproducible from (seed, p, s, c) alone.
pub struct EnsembleSDE<const N: usize> {
    system: SdeSystem, params: Vec<f64>, t0: f64, tf: f64, y0: Vec<[f64; N]>, t_eval: Vec<f64>, method: Option<Method>, seed: u64, path_offset: i64, device: i32,
}

impl<const N: usize> EnsembleSDE<N> {
    pub fn sde(system: SdeSystem, params: Vec<f64>, t0: f64, tf: f64, y0: Vec<[f64; N]>, seed: u64) -> Self {
        assert_eq!(system.dim, N);
        assert!(params.len() == system.n_params || params.len() == system.n_params * y0.len());
        EnsembleSDE { system, params, t0, tf, y0, t_eval: vec![], method: None, seed, path_offset: 0, device: 0 }
    }
    pub fn t_eval(mut self, pts: impl AsRef<[f64]>) -> Self { self.t_eval = pts.as_ref().to_vec(); self }
    pub fn method(mut self, m: Method) -> Self { self.method = Some(m); self }
    pub fn path_offset(mut self, first_global_path: i64) -> Self { self.path_offset = first_global_path; self }
    pub fn device(mut self, ordinal: i32) -> Self { self.device = ordinal; self }
    pub fn solve(self) -> Result<Vec<Result<Solution<f64, [f64; N]>, Error<f64, [f64; N]>>>, String> {
        let m = self.method.clone().expect("method(..) must be set");
        let (n, ne) = (self.y0.len(), self.t_eval.len());
        let mut y_eval = vec![0.0f64; n * ne * N];
        let mut n_emitted = vec![0i32; n];
        let mut t_final = vec![0.0f64; n];
        let mut y_final = vec![[0.0f64; N]; n];
        let (mut status, mut acc, mut evals) = (vec![0i32; n], vec![0i32; n], vec![0i32; n]);
        let mut t_rows = vec![0.0f64; ne.max(1)];
        let mut p: deb_sde_problem = unsafe { std::mem::zeroed() };
        p.struct_size = std::mem::size_of::<deb_sde_problem>();
        p.system = self.system.id; p.method = m.id; p.dim = N as i32; p.n_params = self.system.n_params as i32;
        p.n_traj = n as i64; p.y0 = self.y0.as_ptr() as *const f64; p.params = self.params.as_ptr();
        p.params_shared = (self.params.len() == self.system.n_params) as i32;
        p.n_eval = ne as i32; p.t_eval = if ne == 0 { std::ptr::null() } else { self.t_eval.as_ptr() };
        p.t0 = self.t0; p.tf = self.tf; p.opt = m.options(); p.seed = self.seed; p.path_offset = self.path_offset; p.device = self.device;
        let mut r: deb_result = unsafe { std::mem::zeroed() };
        r.struct_size = std::mem::size_of::<deb_result>();
        r.y_eval = y_eval.as_mut_ptr(); r.n_emitted = n_emitted.as_mut_ptr(); r.t_final = t_final.as_mut_ptr(); r.y_final = y_final.as_mut_ptr() as *mut f64;
        r.status = status.as_mut_ptr(); r.accepted = acc.as_mut_ptr(); r.evals = evals.as_mut_ptr(); r.t_rows = t_rows.as_mut_ptr();
        let rc = unsafe { deb_solve_sde(&p, &mut r) };
        if rc != 0 { return Err(last_error("deb_solve_sde", rc)); }
        Ok((0..n).map(|i| match status[i] {
            0 => {
                let mut s = Solution::new();
                for q in 0..n_emitted[i] as usize {
                    let mut row = [0.0; N];
                    row.copy_from_slice(&y_eval[(i * ne + q) * N..(i * ne + q + 1) * N]);
                    s.push(t_rows[q], row);
                }
                s.status = Status::Complete;
                s.evals = Evals { function: evals[i] as usize, ..Evals::new() };
                s.steps = Steps { accepted: acc[i] as usize, rejected: 0 };
                Ok(s)
            }
            1 => Err(Error::MaxSteps { t: t_final[i], y: y_final[i] }),
            _ => Err(Error::BadInput { msg: "Invalid input".to_string() }),
        }).collect())
    }
}

// ------------------------------------------------------------------------------------------------ method of lines (heat)
/// Boundary condition of `MethodOfLines::finite_difference(grid).boundary(..)` (src/pde/boundary.rs).
#[derive(Clone, Copy, Debug)]
pub enum Boundary { Dirichlet(f64), Neumann(f64) }

/// `IVP::pde(&heat, t0, tf, u0).space(MethodOfLines::finite_difference(StructuredGrid::uniform(lo, hi, n)).boundary(..)).method(rk4(h)).solve()`
/// for u_t = (alpha u_x)_x on one GPU; returns (u(tf), t reached, steps, status).
pub fn solve_heat_mol(u0: &[f64], lo: f64, hi: f64, alpha: f64, method: &Method, t0: f64, tf: f64, lower: Boundary, upper: Boundary,
                      device: i32) -> Result<(Vec<f64>, f64, i64, Status<f64, Vec<f64>>), String> {
    let mut out = vec![0.0f64; u0.len()];
    let (mut t_final, mut steps, mut status) = (0.0f64, 0i64, -1i32);
    let kind = |b: Boundary| match b { Boundary::Dirichlet(v) => (0, v), Boundary::Neumann(v) => (1, v) };
    let mut p: deb_heat_problem = unsafe { std::mem::zeroed() };
    p.struct_size = std::mem::size_of::<deb_heat_problem>();
    p.n_nodes = u0.len() as i64; p.lo = lo; p.hi = hi; p.alpha = alpha;
    (p.bc_lower_kind, p.bc_lower_value) = kind(lower);
    (p.bc_upper_kind, p.bc_upper_value) = kind(upper);
    p.method = method.id; p.h = method.opt.h0; p.t0 = t0; p.tf = tf; p.max_steps = method.opt.max_steps;
    p.u0 = u0.as_ptr(); p.u_final = out.as_mut_ptr(); p.t_final = &mut t_final; p.steps = &mut steps; p.status = &mut status; p.device = device;
    let rc = unsafe { deb_solve_heat_mol(&p) };
    if rc != 0 { return Err(last_error("deb_solve_heat_mol", rc)); }
    Ok((out, t_final, steps, if status == 0 { Status::Complete } else { Status::Error(Error::MaxSteps { t: t_final, y: vec![] }) }))
}
