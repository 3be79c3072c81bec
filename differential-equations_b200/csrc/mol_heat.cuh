// mol_heat.cuh -- method-of-lines heat equation, fixed-step ERK, one large state on one GPU (HBM-bound).
//
// RHS = SemiDiscretePde::diff (/root/reference/src/pde/semi_discrete.rs:250-287) for a scalar field on a uniform
// 1-D grid with the finite-difference flux (:181-190), directional gradients (:150-172), zero source and
// flux = alpha*grad_u (examples/pde/01_heat_equation/main.rs:18-22):
//     Dirichlet boundary node:   du = 0                                   (:264-267, mask built at :47-59)
//     gL = (u_i - u_{i-1})/dx,   gU = (u_{i+1} - u_i)/dx                  (Neumann face: the prescribed gradient)
//     du_i = 0 + (alpha*gU - alpha*gL)/dx                                 (add_scaled_difference :132-139)
// Time stepping = Fixed::step (/root/reference/src/methods/erk/fixed/ordinary.rs:58-139): stage i evaluates the RHS at
// y + sum_j (a_ij*h)*k_j, the solution is y + sum_i (b_i*h)*k_i, and the new derivative is evaluated at once.
//
// Kernel design, whole-step kernel (heat_step_kernel, what deb_solve_heat_mol runs): ONE launch per time step; all S
// stages of a node run on chip, in registers, with a 4-node halo per warp that is recomputed by the neighbouring warp.
// Per node and step the kernel reads 8 B and writes 8 B -- instead of the 128 B of a stage-by-stage sweep (16 arrays for
// RK4) -- and every node sees exactly the operations of the reference (k_0 = f(t_n, y_n) is recomputed from y_n; same
// bits as the value the reference carries over from the previous step).
// Per-stage kernels (heat_stage_kernel; STAGE 0 is what deb_heat_rhs runs): each fuses "stage combine + stencil": a stage
// reads y and the k_j it needs and writes k_i (3 N doubles for RK4); the final kernel fuses "solution combine + next k_1"
// (reads y,k_1..k_S, writes y' and k_1: 7 N doubles for RK4).  16 N doubles = 128 B per node per RK4 step.
// Every thread owns two adjacent nodes (16-byte loads/stores).  The combined state of the neighbouring nodes
// comes from the adjacent lanes by warp shuffle (the two warp-edge lanes fetch their halo through L1/L2), and the
// face gradient g = (w_i - w_{i-1})/dx is computed once per face and shared the same way, so a node costs two IEEE
// divisions instead of three.  When dx is a power of two the divisions become multiplications by the exactly
// representable 1/dx (bit-identical: both are the correctly rounded value of the same real number), which makes
// the N = 2^24, dx = 1 configuration purely HBM-bound.
#pragma once
#include <stdint.h>
#include "erk_tableau.cuh"

namespace deb {

struct HeatArgs {
    long long n;          // nodes
    double dx, inv_dx, alpha;
    int bc_lo_kind, bc_hi_kind;  // 0 Dirichlet, 1 Neumann
    double bc_lo_val, bc_hi_val;
    double h;             // step size of this step
    const double* y;      // state at the start of the step
    const double* k[8];   // stage derivatives (k[0] = derivative at the start of the step)
    double* out_k;        // stage kernel: k_i ; final kernel: the new k_1 ; STAGE 0: f(y)
    double* out_y;        // final kernel: the new state (must not alias y: neighbours read y)
};

template <bool POW2>
__device__ __forceinline__ double div_dx(double x, const HeatArgs& a) {
    return POW2 ? x * a.inv_dx : x / a.dx;
}

// Combined state of node i: y (STAGE 0), y + sum_j (a_{STAGE,j} h) k_j (stage), y + sum_i (b_i h) k_i (STAGE == S).
template <class Tab, int STAGE>
__device__ __forceinline__ double combine1(const HeatArgs& a, long long i) {
    constexpr int ST = (STAGE < Tab::S) ? STAGE : 0;
    double w = a.y[i];
    if (STAGE < Tab::S) {
#pragma unroll
        for (int j = 0; j < ST; j++)
            if (Tab::a(ST, j) != 0.0) w = w + (Tab::av(ST, j) * a.h) * a.k[j][i];
    } else {
#pragma unroll
        for (int j = 0; j < Tab::S; j++)
            if (Tab::b(j) != 0.0) w = w + (Tab::bv(j) * a.h) * a.k[j][i];
    }
    return w;
}
template <class Tab, int STAGE>
__device__ __forceinline__ double2 combine2(const HeatArgs& a, long long i0) {
    constexpr int ST = (STAGE < Tab::S) ? STAGE : 0;
    double2 w = *reinterpret_cast<const double2*>(a.y + i0);
    if (STAGE < Tab::S) {
#pragma unroll
        for (int j = 0; j < ST; j++)
            if (Tab::a(ST, j) != 0.0) {
                const double2 kk = *reinterpret_cast<const double2*>(a.k[j] + i0);
                const double ah = Tab::av(ST, j) * a.h;
                w.x = w.x + ah * kk.x;
                w.y = w.y + ah * kk.y;
            }
    } else {
#pragma unroll
        for (int j = 0; j < Tab::S; j++)
            if (Tab::b(j) != 0.0) {
                const double2 kk = *reinterpret_cast<const double2*>(a.k[j] + i0);
                const double bh = Tab::bv(j) * a.h;
                w.x = w.x + bh * kk.x;
                w.y = w.y + bh * kk.y;
            }
    }
    return w;
}

template <class Tab, int STAGE, bool POW2>
__global__ void __launch_bounds__(256) heat_stage_kernel(const HeatArgs a) {
    const long long i0 = 2 * ((long long)blockIdx.x * 256 + threadIdx.x);  // this thread owns nodes i0, i0+1
    const unsigned lane = threadIdx.x & 31u;
    const long long n = a.n;
    const bool in0 = i0 < n, in1 = i0 + 1 < n;
    double w0 = 0.0, w1 = 0.0;
    if (in1) {
        const double2 w = combine2<Tab, STAGE>(a, i0);
        w0 = w.x; w1 = w.y;
    } else if (in0) {
        w0 = combine1<Tab, STAGE>(a, i0);
    }
    // neighbours: w_{i0-1} is the lane below's second node, w_{i0+2} the lane above's first node
    double wl = __shfl_up_sync(0xffffffffu, w1, 1);
    double wr = __shfl_down_sync(0xffffffffu, w0, 1);
    if (lane == 0 && in0 && i0 > 0) wl = combine1<Tab, STAGE>(a, i0 - 1);
    if (lane == 31 && i0 + 2 < n) wr = combine1<Tab, STAGE>(a, i0 + 2);
    // g0: lower face of node i0 (prescribed gradient on a Neumann boundary face)
    const double g0 = (in0 && i0 > 0) ? div_dx<POW2>(w0 - wl, a) : a.bc_lo_val;
    // upper face of node i0+1 == lower face of the next lane's first node
    double g2 = __shfl_down_sync(0xffffffffu, g0, 1);
    if (lane == 31 && i0 + 2 < n) g2 = div_dx<POW2>(wr - w1, a);
    if (!in0) return;
    const double g1 = in1 ? div_dx<POW2>(w1 - w0, a) : a.bc_hi_val;  // face between the two nodes (or the upper boundary face)
    if (!(i0 + 2 < n)) g2 = a.bc_hi_val;
    const double f0 = a.alpha * g0, f1 = a.alpha * g1, f2 = a.alpha * g2;
    double d0 = __dadd_rn(0.0, div_dx<POW2>(f1 - f0, a));
    double d1 = __dadd_rn(0.0, div_dx<POW2>(f2 - f1, a));
    if (a.bc_lo_kind == 0 && i0 == 0) d0 = 0.0;  // Dirichlet boundary nodes keep their value
    if (a.bc_hi_kind == 0) {
        if (i0 == n - 1) d0 = 0.0;
        if (i0 + 1 == n - 1) d1 = 0.0;
    }
    if (in1) {
        if (STAGE == Tab::S) *reinterpret_cast<double2*>(a.out_y + i0) = make_double2(w0, w1);
        *reinterpret_cast<double2*>(a.out_k + i0) = make_double2(d0, d1);
    } else {
        if (STAGE == Tab::S) a.out_y[i0] = w0;
        a.out_k[i0] = d0;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Whole-step kernel: y_{n+1} = y_n + sum_i (b_i h) k_i with all stages evaluated in registers, one launch per step.
// Every WARP is independent: it owns HEAT_OUT = 120 output nodes and works on 128 consecutive nodes (4 halo nodes per
// side, 4 per lane), starting at a multiple of 4 so that each lane's state is two aligned 16-byte loads.  A stage needs
// the stage state of the neighbouring nodes: w[3] of the lane below and w[0] of the lane above, by warp shuffle.  Lanes 0
// and 31 have no valid neighbour on one side; the error this puts into their edge node moves inwards by one node per
// stage and, with at most 4 stages, never leaves those two halo lanes, whose results are not stored.  No shared memory,
// no block-level synchronisation, 6.7 % redundant stencil work.
constexpr int HEAT_OUT = 120;

// EDGE = false: the warp's 128 nodes and their neighbours are all interior nodes (no boundary tests at all).
template <class Tab, bool POW2, bool EDGE>
__device__ __forceinline__ void heat_warp_step(const HeatArgs& a, const long long i0, const unsigned lane) {
    constexpr int S = Tab::S;
    const long long n = a.n;
    double u[4];
    if (!EDGE || (i0 >= 0 && i0 + 3 < n)) {
        const double2 lo = *reinterpret_cast<const double2*>(a.y + i0);
        const double2 hi = *reinterpret_cast<const double2*>(a.y + i0 + 2);
        u[0] = lo.x; u[1] = lo.y; u[2] = hi.x; u[3] = hi.y;
    } else {
#pragma unroll
        for (int m = 0; m < 4; m++) u[m] = (i0 + m >= 0 && i0 + m < n) ? a.y[i0 + m] : 0.0;
    }
    const bool dir_lo = (a.bc_lo_kind == 0), dir_hi = (a.bc_hi_kind == 0);
    double k[S][4];
#pragma unroll
    for (int s = 0; s < S; s++) {
        double w[4];
#pragma unroll
        for (int m = 0; m < 4; m++) w[m] = u[m];
#pragma unroll
        for (int j = 0; j < s; j++) {  // stage state y + sum_j (a_sj h) k_j, fixed/ordinary.rs:80-88
            if (Tab::a(s, j) != 0.0) {
                const double ah = Tab::av(s, j) * a.h;
#pragma unroll
                for (int m = 0; m < 4; m++) w[m] = w[m] + ah * k[j][m];
            }
        }
        const double wl = __shfl_up_sync(0xffffffffu, w[3], 1);
        const double wr = __shfl_down_sync(0xffffffffu, w[0], 1);
        // face gradients (semi_discrete.rs:150-172): g[m] is the lower face of node m, g[4] the upper face of node 3;
        // a face on the domain boundary carries the prescribed (Neumann) gradient
        double g[5];
        if (EDGE) {
            g[0] = (i0 > 0) ? div_dx<POW2>(w[0] - wl, a) : a.bc_lo_val;
#pragma unroll
            for (int m = 1; m < 4; m++)
                g[m] = (i0 + m > 0) ? ((i0 + m < n) ? div_dx<POW2>(w[m] - w[m - 1], a) : a.bc_hi_val) : a.bc_lo_val;
            g[4] = (i0 + 4 < n) ? div_dx<POW2>(wr - w[3], a) : a.bc_hi_val;
        } else {
            g[0] = div_dx<POW2>(w[0] - wl, a);
#pragma unroll
            for (int m = 1; m < 4; m++) g[m] = div_dx<POW2>(w[m] - w[m - 1], a);
            g[4] = div_dx<POW2>(wr - w[3], a);
        }
        double f[5];
#pragma unroll
        for (int m = 0; m < 5; m++) f[m] = a.alpha * g[m];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            double d = __dadd_rn(0.0, div_dx<POW2>(f[m + 1] - f[m], a));                              // :132-139, :181-190
            if (EDGE && ((dir_lo && i0 + m == 0) || (dir_hi && i0 + m == n - 1))) d = 0.0;            // Dirichlet nodes, :264-267
            k[s][m] = d;
        }
    }
    if (lane == 0 || lane == 31) return;  // halo lanes
    double v[4];
#pragma unroll
    for (int m = 0; m < 4; m++) v[m] = u[m];
#pragma unroll
    for (int j = 0; j < S; j++) {  // solution, fixed/ordinary.rs:98-102
        if (Tab::b(j) != 0.0) {
            const double bh = Tab::bv(j) * a.h;
#pragma unroll
            for (int m = 0; m < 4; m++) v[m] = v[m] + bh * k[j][m];
        }
    }
    if (!EDGE || i0 + 3 < n) {
        *reinterpret_cast<double2*>(a.out_y + i0) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(a.out_y + i0 + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int m = 0; m < 4; m++)
            if (i0 + m < n) a.out_y[i0 + m] = v[m];
    }
}

template <class Tab, bool POW2>
__global__ void __launch_bounds__(256) heat_step_kernel(const HeatArgs a) {
    static_assert(Tab::S <= 4, "the 4-node halo covers at most 4 stages");
    const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const long long first = warp * HEAT_OUT - 4;     // first of the warp's 128 nodes
    if (warp * HEAT_OUT >= a.n) return;              // whole warp out of range
    const long long i0 = first + 4 * (long long)lane;
    // warp-uniform: do the 128 nodes or their outer neighbours touch the ends of the grid?
    if (first > 1 && first + 128 < a.n - 1) heat_warp_step<Tab, POW2, false>(a, i0, lane);
    else heat_warp_step<Tab, POW2, true>(a, i0, lane);
}

}  // namespace deb
