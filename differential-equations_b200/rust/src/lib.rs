//! Ensemble front end: `EnsembleIVP::ode(system, t0, tf, y0s).t_eval(pts).method(m).solve()` mirrors
//! `IVP::ode(&sys, t0, tf, y0).t_eval(pts).method(m).solve()` (differential-equations `src/ivp.rs:279,656,632,781`)
//! for N independent problems, executed by the sm_100a kernels behind `include/deb_ensemble.h`.
//!
//! SOURCE ONLY: never compiled (no Rust toolchain in the build image).  The `#[repr(C)]` structs below are a
//! field-for-field transcription of `include/deb_ensemble.h` (ABI version 3); `tests/test_abi_cpu.py` checks the
//! same layout for the Python mirror against the compiled header.
#![allow(non_camel_case_types)]

use std::ffi::{c_void, CStr};
use std::os::raw::c_char;

use differential_equations::{
    error::Error,
    solution::Solution,
    stats::{Evals, Steps},
    status::Status,
};

// ------------------------------------------------------------------------------------------------ raw ABI
pub const DEB_ABI_VERSION: i32 = 8;

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
    pub event_coef: [f64; 18],
    pub plane_dim: i32,          // hyperplane_crossing: number of extracted components
    pub plane_index: [i32; 16],
    pub plane_point: [f64; 16],
    pub plane_normal: [f64; 16],
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
}

extern "C" {
    pub fn deb_abi_version() -> i32;
    pub fn deb_last_error() -> *const c_char;
    pub fn deb_device_count() -> i32;
    pub fn deb_erk_options_default(opt: *mut deb_erk_options);
    pub fn deb_solve_ode(problem: *const deb_ode_problem, result: *mut deb_result) -> i32;
    /// user-defined right-hand side as CUDA C++ text (the device-side `impl ODE`); returns a system id >= 1000
    /// release the device memory the library caches between calls
    pub fn deb_trim_memory(device: i32) -> i32;
    pub fn deb_define_ode(dim: i32, n_params: i32, diff_body: *const c_char, system_id: *mut i32) -> i32;
    /// compile it for a method now (no device needed); the compiler log is in deb_last_error()
    pub fn deb_check_ode(system_id: i32, method: i32, solout: i32, event: i32) -> i32;
    /// user-defined event function g(t, y) as CUDA C++ text (the device-side `impl Event`); returns an event id >= 1000
    pub fn deb_define_event(dim: i32, event_body: *const c_char, event_id: *mut i32) -> i32;
    // deb_solve_sde, deb_solve_heat_mol, deb_heat_rhs, deb_ensemble_stats, deb_malloc, ... : see deb_ensemble.h
}

// ------------------------------------------------------------------------------------------------ safe layer
/// Built-in right-hand sides (`deb_system`); a closure cannot cross to the device.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum System {
    Exponential = 0,
    Linear = 1,
    Harmonic = 2,
    Logistic = 3,
    VanDerPol = 4,
    Lorenz = 5,
    Brusselator = 6,
    Robertson = 7,
}

impl System {
    pub fn dim(self) -> usize {
        match self {
            System::Exponential | System::Linear | System::Logistic => 1,
            System::Harmonic | System::VanDerPol | System::Brusselator => 2,
            System::Lorenz | System::Robertson => 3,
        }
    }
}

/// `ExplicitRungeKutta::dopri5()/dop853()/rk4(h)/...` with the crate's setters (`src/methods/erk/mod.rs:164-228`).
#[derive(Clone)]
pub struct Method {
    id: i32,
    opt: deb_erk_options,
    rtol_vec: Option<Vec<f64>>,
    atol_vec: Option<Vec<f64>>,
}

impl Method {
    fn new(id: i32, h0: f64) -> Self {
        let mut opt = unsafe { std::mem::zeroed::<deb_erk_options>() };
        unsafe { deb_erk_options_default(&mut opt) }; // rtol = atol = 1e-6, max_steps = 10_000, safety 0.9, scale in [0.2, 10]
        opt.h0 = h0;
        Method { id, opt, rtol_vec: None, atol_vec: None }
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
}

/// N independent IVPs sharing (t0, tf, method): the ensemble analogue of `IVP` (`src/ivp.rs`).
pub struct EnsembleIVP<const N: usize> {
    system: System,
    params: Vec<f64>, // one set (shared) or N_traj sets, flattened
    params_shared: bool,
    t0: f64,
    tf: f64,
    y0: Vec<[f64; N]>,
    t_eval: Vec<f64>,
    method: Option<Method>,
    device: i32,
}

impl<const N: usize> EnsembleIVP<N> {
    /// `IVP::ode(&sys, t0, tf, y0)`; `params` holds one parameter set for all trajectories or one per trajectory.
    pub fn ode(system: System, params: Vec<f64>, t0: f64, tf: f64, y0: Vec<[f64; N]>) -> Self {
        assert_eq!(system.dim(), N, "state dimension must match the system");
        let n_params = n_params_of(system);
        let params_shared = params.len() == n_params;
        assert!(params_shared || params.len() == n_params * y0.len(), "params: one set, or one per trajectory");
        EnsembleIVP { system, params, params_shared, t0, tf, y0, t_eval: vec![], method: None, device: 0 }
    }
    pub fn t_eval(mut self, pts: impl AsRef<[f64]>) -> Self { self.t_eval = pts.as_ref().to_vec(); self }
    pub fn method(mut self, m: Method) -> Self { self.method = Some(m); self }
    pub fn device(mut self, ordinal: i32) -> Self { self.device = ordinal; self }

    /// One `Result<Solution, Error>` per trajectory, exactly what N calls of `IVP::solve()` would have returned
    /// (timer excluded).  Host buffers; the library copies both ways (memspace HOST).
    pub fn solve(self) -> Result<Vec<Result<Solution<f64, [f64; N]>, Error<f64, [f64; N]>>>, String> {
        let m = self.method.expect("method(..) must be set");
        let n = self.y0.len();
        let ne = self.t_eval.len();
        let mut y_eval = vec![0.0f64; n * ne * N];
        let mut n_emitted = vec![0i32; n];
        let mut t_final = vec![0.0f64; n];
        let mut y_final = vec![[0.0f64; N]; n];
        let (mut status, mut acc, mut rej, mut evals) = (vec![0i32; n], vec![0i32; n], vec![0i32; n], vec![0i32; n]);
        let mut t_rows = vec![0.0f64; ne.max(1)];
        let mut opt = m.opt;
        opt.rtol_vec = m.rtol_vec.as_ref().map_or(std::ptr::null(), |v| v.as_ptr());
        opt.atol_vec = m.atol_vec.as_ref().map_or(std::ptr::null(), |v| v.as_ptr());
        let problem = deb_ode_problem {
            struct_size: std::mem::size_of::<deb_ode_problem>(),
            system: self.system as i32,
            method: m.id,
            dim: N as i32,
            n_params: n_params_of(self.system) as i32,
            n_traj: n as i64,
            y0: self.y0.as_ptr() as *const f64, // Vec<[f64; N]> is the ABI's "array of states" layout
            params: self.params.as_ptr(),
            params_shared: self.params_shared as i32,
            n_eval: ne as i32,
            t_eval: self.t_eval.as_ptr(),
            t0: self.t0,
            tf: self.tf,
            opt,
            device: self.device,
            memspace: 0, // DEB_MEM_HOST
            stream: std::ptr::null_mut(),
            solout: 0,
            dense_n: 0,
            even_dt: 0.0,
            cross_component: 0,
            cross_direction: 0,
            cross_threshold: 0.0,
            event: 0,
            event_direction: 0,
            event_terminate: 0,
            row_capacity: 0,
            event_coef: [0.0; 18],
            plane_dim: 0,
            plane_index: [0; 16],
            plane_point: [0.0; 16],
            plane_normal: [0.0; 16],
        };
        let mut result = deb_result {
            struct_size: std::mem::size_of::<deb_result>(),
            y_eval: y_eval.as_mut_ptr(),
            n_emitted: n_emitted.as_mut_ptr(),
            t_final: t_final.as_mut_ptr(),
            y_final: y_final.as_mut_ptr() as *mut f64,
            status: status.as_mut_ptr(),
            accepted: acc.as_mut_ptr(),
            rejected: rej.as_mut_ptr(),
            evals: evals.as_mut_ptr(),
            t_rows: t_rows.as_mut_ptr(),
            n_rows: 0,
            kernel_ms: 0.0,
            total_ms: 0.0,
            t_out: std::ptr::null_mut(),
        };
        let rc = unsafe { deb_solve_ode(&problem, &mut result) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(deb_last_error()) }.to_string_lossy().into_owned();
            return Err(format!("deb_solve_ode failed ({rc}): {msg}")); // includes DEB_ERR_NO_DEVICE: there is no CPU fallback
        }
        let mut out = Vec::with_capacity(n);
        for i in 0..n {
            let (t, y) = (t_final[i], y_final[i]);
            out.push(match status[i] {
                0 => {
                    let mut s = Solution::new();
                    for r in 0..n_emitted[i] as usize {
                        let mut row = [0.0; N];
                        row.copy_from_slice(&y_eval[(i * ne + r) * N..(i * ne + r + 1) * N]);
                        s.push(t_rows[r], row);
                    }
                    s.status = Status::Complete;
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
        Ok(out)
    }
}

fn n_params_of(s: System) -> usize {
    match s {
        System::Exponential | System::Harmonic | System::VanDerPol => 1,
        System::Linear | System::Logistic | System::Brusselator => 2,
        System::Lorenz => 3,
        System::Robertson => 0,
    }
}
