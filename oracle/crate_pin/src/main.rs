//! Pins the CPU oracle (oracle/oracle.cpp) against the REAL crate, bit for bit.
//!
//! The build image of this repository has no Rust toolchain, so the oracle is a C++ restatement whose bits have never been
//! compared with the crate itself ("parity unpinned" at the bit level).  On any machine with cargo:
//!
//!     tools/pin_oracle_against_crate.sh          # (cd oracle/crate_pin && cargo run --release) > tests/golden/reference_bits.json
//!     python -m pytest tests/test_oracle_golden.py -k reference_bits
//!
//! The program runs the reference's own `IVP::ode(..).method(..).solve()` (src/ivp.rs:279,632,656,781) on a fixed list of cases
//! -- the first 64 trajectories of config C1/C2, a DOP853 parameter sweep, every fixed-step constructor, backward time, the
//! Dormand-Prince stiffness exit (1499 accepted steps), MaxSteps -- and prints every input and output as hexadecimal bit
//! patterns (f64::to_bits), so that nothing is lost in a decimal round trip.  tests/test_oracle_golden.py replays the same
//! cases through the oracle and compares the bits.
//!
//! Uses the crate's `[f64; N]` State (src/traits.rs:511): the same component-wise arithmetic as every other State impl.
use differential_equations::prelude::*;

struct Lorenz { sigma: f64, rho: f64, beta: f64 }
impl ODE<f64, [f64; 3]> for Lorenz {
    fn diff(&self, _t: f64, y: &[f64; 3], dydt: &mut [f64; 3]) {  // expression order of tests/ode/systems.rs:91-101
        let x = y[0];
        let y_val = y[1];
        let z = y[2];
        dydt[0] = self.sigma * (y_val - x);
        dydt[1] = x * (self.rho - z) - y_val;
        dydt[2] = x * y_val - self.beta * z;
    }
}
struct VanDerPol { mu: f64 }
impl ODE<f64, [f64; 2]> for VanDerPol {
    fn diff(&self, _t: f64, y: &[f64; 2], dydt: &mut [f64; 2]) {  // tests/ode/systems.rs:70-78
        let y1 = y[0];
        let y2 = y[1];
        dydt[0] = y2;
        dydt[1] = self.mu * (1.0 - y1 * y1) * y2 - y1;
    }
}
struct Exponential { k: f64 }
impl ODE<f64, [f64; 1]> for Exponential {
    fn diff(&self, _t: f64, y: &[f64; 1], dydt: &mut [f64; 1]) { dydt[0] = self.k * y[0]; }  // tests/ode/systems.rs:12-16
}

/// u_k = (splitmix64(seed + k*golden) >> 11) * 2^-53 - 0.5, k >= 1: the ensemble generator of the package
/// (differential-equations_b200/__init__.py: splitmix64_uniform / perturbed_ensemble)
fn splitmix_uniform(seed: u64, k: u64) -> f64 {
    let mut z = seed.wrapping_add(k.wrapping_mul(0x9E3779B97F4A7C15));
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^= z >> 31;
    ((z >> 11) as f64) * (2.0f64).powi(-53) - 0.5
}

fn hex(v: f64) -> String { format!("\"{:016x}\"", v.to_bits()) }
fn hex_vec(v: &[f64]) -> String { format!("[{}]", v.iter().map(|x| hex(*x)).collect::<Vec<_>>().join(",")) }

fn report<const N: usize>(r: Result<Solution<f64, [f64; N]>, Error<f64, [f64; N]>>) -> String {
    match r {
        Ok(s) => format!(
            "{{\"status\":\"{:?}\",\"accepted\":{},\"rejected\":{},\"evals\":{},\"t\":{},\"y\":[{}]}}",
            s.status, s.steps.accepted, s.steps.rejected, s.evals.function, hex_vec(&s.t),
            s.y.iter().map(|y| hex_vec(&y[..])).collect::<Vec<_>>().join(",")),
        Err(Error::MaxSteps { t, y }) => format!("{{\"status\":\"MaxSteps\",\"t_final\":{},\"y_final\":{}}}", hex(t), hex_vec(&y[..])),
        Err(Error::StepSize { t, y }) => format!("{{\"status\":\"StepSize\",\"t_final\":{},\"y_final\":{}}}", hex(t), hex_vec(&y[..])),
        Err(Error::Stiffness { t, y }) => format!("{{\"status\":\"Stiffness\",\"t_final\":{},\"y_final\":{}}}", hex(t), hex_vec(&y[..])),
        Err(Error::BadInput { .. }) => "{\"status\":\"BadInput\"}".to_string(),
        Err(e) => format!("{{\"status\":\"{:?}\"}}", e),
    }
}

fn main() {
    let mut cases: Vec<String> = Vec::new();
    let t_eval: Vec<f64> = (1..=100).map(|i| i as f64).collect();

    // ---- C1 / C2: Lorenz, dopri5().rtol(1e-8), t in [0,100], t_eval = 1..100, first 64 trajectories of the bench ensemble
    let lorenz = Lorenz { sigma: 10.0, rho: 28.0, beta: 8.0 / 3.0 };
    for (method, n_traj) in [("dopri5", 64usize), ("dop853", 16), ("rkf45", 8), ("cash_karp", 8), ("rkv655e", 4), ("rkv989e", 4)] {
        let mut y0s = Vec::new();
        let mut res = Vec::new();
        for i in 0..n_traj as u64 {
            let y0 = [1.0 + splitmix_uniform(2026, 3 * i + 1), 1.0 + splitmix_uniform(2026, 3 * i + 2), 1.0 + splitmix_uniform(2026, 3 * i + 3)];
            y0s.push(hex_vec(&y0));
            let ivp = IVP::ode(&lorenz, 0.0, 100.0, y0).t_eval(t_eval.clone());
            res.push(match method {
                "dopri5" => report(ivp.method(ExplicitRungeKutta::dopri5().rtol(1e-8)).solve()),
                "dop853" => report(ivp.method(ExplicitRungeKutta::dop853().rtol(1e-8)).solve()),
                "rkf45" => report(ivp.method(ExplicitRungeKutta::rkf45().rtol(1e-8)).solve()),
                "cash_karp" => report(ivp.method(ExplicitRungeKutta::cash_karp().rtol(1e-8)).solve()),
                "rkv655e" => report(ivp.method(ExplicitRungeKutta::rkv655e().rtol(1e-8)).solve()),
                _ => report(ivp.method(ExplicitRungeKutta::rkv989e().rtol(1e-8)).solve()),
            });
        }
        cases.push(format!(
            "{{\"name\":\"lorenz_{m}\",\"system\":\"lorenz\",\"params\":{p},\"method\":\"{m}\",\"rtol\":{rt},\"t0\":{t0},\"tf\":{tf},\"t_eval\":{te},\"y0\":[{y0}],\"results\":[{r}]}}",
            m = method, p = hex_vec(&[10.0, 28.0, 8.0 / 3.0]), rt = hex(1e-8), t0 = hex(0.0), tf = hex(100.0), te = hex_vec(&t_eval), y0 = y0s.join(","), r = res.join(",")));
    }

    // ---- C3 shape: Van der Pol mu sweep, dop853 rtol = atol = 1e-8, final state only
    {
        let mut y0s = Vec::new();
        let mut res = Vec::new();
        let mut mus = Vec::new();
        for i in 0..16 {
            let mu = 0.1 + 49.9 * (i as f64) / 15.0;
            mus.push(mu);
            y0s.push(hex_vec(&[2.0, 0.0]));
            let sys = VanDerPol { mu };
            res.push(report(IVP::ode(&sys, 0.0, 100.0, [2.0, 0.0]).t_eval(vec![100.0]).method(ExplicitRungeKutta::dop853().rtol(1e-8).atol(1e-8)).solve()));
        }
        cases.push(format!(
            "{{\"name\":\"vdp_dop853\",\"system\":\"van_der_pol\",\"params_per_traj\":{p},\"method\":\"dop853\",\"rtol\":{rt},\"atol\":{rt},\"t0\":{t0},\"tf\":{tf},\"t_eval\":{te},\"y0\":[{y0}],\"results\":[{r}]}}",
            p = hex_vec(&mus), rt = hex(1e-8), t0 = hex(0.0), tf = hex(100.0), te = hex_vec(&[100.0]), y0 = y0s.join(","), r = res.join(",")));
    }

    // ---- fixed-step constructors (Lorenz, h = 0.01, t in [0,2]) and backward time
    for method in ["euler", "midpoint", "heun", "ralston", "ssp_rk3", "rk4", "three_eighths"] {
        let y0 = [1.0, 1.0, 1.0];
        let te = vec![0.5, 1.0, 1.5, 2.0];
        let ivp = IVP::ode(&lorenz, 0.0, 2.0, y0).t_eval(te.clone());
        let r = match method {
            "euler" => report(ivp.method(ExplicitRungeKutta::euler(0.01)).solve()),
            "midpoint" => report(ivp.method(ExplicitRungeKutta::midpoint(0.01)).solve()),
            "heun" => report(ivp.method(ExplicitRungeKutta::heun(0.01)).solve()),
            "ralston" => report(ivp.method(ExplicitRungeKutta::ralston(0.01)).solve()),
            "ssp_rk3" => report(ivp.method(ExplicitRungeKutta::ssp_rk3(0.01)).solve()),
            "rk4" => report(ivp.method(ExplicitRungeKutta::rk4(0.01)).solve()),
            _ => report(ivp.method(ExplicitRungeKutta::three_eighths(0.01)).solve()),
        };
        cases.push(format!(
            "{{\"name\":\"lorenz_{m}\",\"system\":\"lorenz\",\"params\":{p},\"method\":\"{m}\",\"h0\":{h},\"t0\":{t0},\"tf\":{tf},\"t_eval\":{te},\"y0\":[{y0}],\"results\":[{r}]}}",
            m = method, p = hex_vec(&[10.0, 28.0, 8.0 / 3.0]), h = hex(0.01), t0 = hex(0.0), tf = hex(2.0), te = hex_vec(&te), y0 = hex_vec(&y0), r = r));
    }

    // ---- the Dormand-Prince stiffness exit: y' = k y, h_max = 0.002 -> Err(Stiffness) after 1499 accepted steps
    for (k, method) in [(1.0000001, "dopri5"), (1.0, "dopri5"), (1.0000001, "dop853")] {
        let sys = Exponential { k };
        let te: Vec<f64> = (0..50).map(|i| 10.0 * (i as f64) / 49.0).collect();
        let ivp = IVP::ode(&sys, 0.0, 10.0, [1.0]).t_eval(te.clone());
        let r = if method == "dopri5" { report(ivp.method(ExplicitRungeKutta::dopri5().h_max(0.002).max_steps(100000)).solve()) }
                else { report(ivp.method(ExplicitRungeKutta::dop853().h_max(0.002).max_steps(100000)).solve()) };
        cases.push(format!(
            "{{\"name\":\"stiffness_{m}_{kb}\",\"system\":\"exponential\",\"params\":{p},\"method\":\"{m}\",\"h_max\":{hm},\"max_steps\":100000,\"t0\":{t0},\"tf\":{tf},\"t_eval\":{te},\"y0\":[{y0}],\"results\":[{r}]}}",
            m = method, kb = hex(k).trim_matches('"'), p = hex_vec(&[k]), hm = hex(0.002), t0 = hex(0.0), tf = hex(10.0), te = hex_vec(&te), y0 = hex_vec(&[1.0]), r = r));
    }
    println!("{{\"crate\":\"differential-equations\",\"crate_version\":\"{}\",\"cases\":[\n{}\n]}}", env!("CARGO_PKG_VERSION"), cases.join(",\n"));
}
