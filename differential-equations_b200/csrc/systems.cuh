// systems.cuh -- built-in right-hand sides (`ODE::diff`, /root/reference/src/ode/ode.rs:44) as device functors.
//
// A Rust closure cannot cross to the device, so the systems of the reference's own tests and benches are compiled
// in.  Expression order follows /root/reference/tests/ode/systems.rs exactly (this TU is compiled with -fmad=false,
// so `a*b+c` stays two separately rounded operations like in Rust).
#pragma once

namespace deb {

struct SysExponential {  // systems.rs:12-16
    static constexpr int DIM = 1, NP = 1;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) { d[0] = p[0] * y[0]; }
};
struct SysLinear {  // systems.rs:25-29
    static constexpr int DIM = 1, NP = 2;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) { d[0] = p[0] + p[1] * y[0]; }
};
struct SysHarmonic {  // systems.rs:38-43
    static constexpr int DIM = 2, NP = 1;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) {
        d[0] = y[1];
        d[1] = -p[0] * y[0];
    }
};
struct SysLogistic {  // systems.rs:55-59
    static constexpr int DIM = 1, NP = 2;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) {
        d[0] = p[0] * y[0] * (1.0 - y[0] / p[1]);
    }
};
struct SysVanDerPol {  // systems.rs:70-78
    static constexpr int DIM = 2, NP = 1;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) {
        const double y1 = y[0], y2 = y[1];
        d[0] = y2;
        d[1] = p[0] * (1.0 - y1 * y1) * y2 - y1;
    }
};
struct SysLorenz {  // systems.rs:91-101
    static constexpr int DIM = 3, NP = 3;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) {
        const double x = y[0], yv = y[1], z = y[2];
        d[0] = p[0] * (yv - x);
        d[1] = x * (p[1] - z) - yv;
        d[2] = x * yv - p[2] * z;
    }
};
struct SysBrusselator {  // systems.rs:112-120
    static constexpr int DIM = 2, NP = 2;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double* p) {
        const double y1 = y[0], y2 = y[1];
        d[0] = p[0] + y1 * y1 * y2 - (p[1] + 1.0) * y1;
        d[1] = p[1] * y1 - y1 * y1 * y2;
    }
};

struct SysRobertson {  // systems.rs:161-173 (stiff; the reference tests DOP853/DOPRI5 on it with loose tolerances)
    static constexpr int DIM = 3, NP = 0;
    __device__ __forceinline__ static void rhs(double, const double* y, double* d, const double*) {
        const double y1 = y[0], y2 = y[1], y3 = y[2];
        d[0] = -0.04 * y1 + 1.0e4 * y2 * y3;
        d[1] = 0.04 * y1 - 1.0e4 * y2 * y3 - 3.0e7 * y2 * y2;
        d[2] = 3.0e7 * y2 * y2;
    }
};

// SDEs with diagonal noise (`SDE::drift` / `SDE::diffusion`, /root/reference/src/sde/sde.rs:16-52).  `mix` is what the
// system's `SDE::noise` does with the independent Wiener increments of the library's Philox stream.
// NPX extra parameter slots hold per-path constants that `prepare` derives once (loop-invariant parts of `mix`).
struct SdeOU {  // examples/sde/03_ornstein_uhlenbeck/main.rs:42-49
    static constexpr int DIM = 1, NP = 3, NPX = 0;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static void drift(double, const double* y, double* d, const double* p) { d[0] = p[0] * (p[1] - y[0]); }
    __device__ __forceinline__ static void diffusion(double, const double*, double* g, const double* p) { g[0] = p[2]; }
    __device__ __forceinline__ static void mix(double*, const double*) {}
};
struct SdeGBM {  // src/sde/solve_ivp.rs doc example: drift mu*y, diffusion sigma*y
    static constexpr int DIM = 1, NP = 2, NPX = 0;
    __device__ __forceinline__ static void prepare(double*) {}
    __device__ __forceinline__ static void drift(double, const double* y, double* d, const double* p) { d[0] = p[0] * y[0]; }
    __device__ __forceinline__ static void diffusion(double, const double* y, double* g, const double* p) { g[0] = p[1] * y[0]; }
    __device__ __forceinline__ static void mix(double*, const double*) {}
};
struct SdeHeston {  // examples/sde/02_heston_model/main.rs:53-72; p = {mu, kappa, theta, sigma, rho}, y = {price, variance}
    static constexpr int DIM = 2, NP = 5, NPX = 1;
    __device__ __forceinline__ static void prepare(double* p) { p[5] = sqrt(1.0 - p[4] * p[4]); }  // the same value every step
    __device__ __forceinline__ static void drift(double, const double* y, double* d, const double* p) {
        d[0] = p[0] * y[0];
        d[1] = p[1] * (p[2] - y[1]);
    }
    __device__ __forceinline__ static void diffusion(double, const double* y, double* g, const double* p) {
        g[0] = y[0] * sqrt(y[1]);
        g[1] = p[3] * sqrt(y[1]);
    }
    __device__ __forceinline__ static void mix(double* dw, const double* p) { dw[1] = p[4] * dw[0] + p[5] * dw[1]; }
};

}  // namespace deb
