// erk_ensemble.cuh -- persistent ensemble kernels for the explicit Runge-Kutta hot path (sm_100a).
//
// One trajectory per thread.  The whole per-trajectory call chain of the reference,
//     solve_ode loop            /root/reference/src/ode/solve_ivp.rs:139-277
//     DormandPrince init/step   /root/reference/src/methods/erk/dormandprince/ordinary.rs:16-270
//     initial step size         /root/reference/src/methods/h_init.rs:45-135
//     step-size utilities       /root/reference/src/utils.rs:22-157
//     State algebra             /root/reference/src/traits.rs:316-410
//     TEvalSolout + dense output  src/solout/t_eval.rs:87-137, dormandprince/ordinary.rs:301-337
// is fused into one kernel: state, the S stage vectors and the controller live in registers, the tableau is
// immediates (erk_tableau.cuh), there is no per-step memory traffic.  A warp executes "one step attempt per lane
// per iteration"; lanes that finished (or failed) are refilled from a global trajectory queue with one
// warp-aggregated atomicAdd (__ballot_sync + __popc), so warps stay full although adaptive step counts diverge.
//
// Arithmetic contract: every + and * below is one separately rounded IEEE f64 operation in the association order
// of the Rust source (this TU is compiled with -fmad=false; Rust never contracts), sqrt and / are IEEE-exact in
// CUDA, max/min ignore NaN like f64::max/min, and powf is the bit-exact glibc port of glibc_pow.h.  Terms whose
// tableau coefficient is zero are skipped (identical for finite stage values; see DESIGN.md for the NaN/-0 note).
#pragma once
#include <float.h>
#include <stdint.h>

#include "../../include/deb_ensemble.h"
#include "erk_tableau.cuh"
#include "glibc_pow.h"

namespace deb {

constexpr int DEB_FX_MAX_TAIL = 4;

struct OdeKernelArgs {
    const double* y0;       // [n_traj][DIM]
    const double* params;   // [n_traj][NP] or [NP]
    int params_stride;      // NP, or 0 when shared
    double pc[8];           // shared parameter set by value: constant-bank operands, no registers (SHARED_P kernels)
    long long n_traj;
    double t0, tf;
    double rtol[DEB_MAX_DIM], atol[DEB_MAX_DIM];  // Tolerance indexed per component (tolerance.rs:32-41)
    double h0, h_min, h_max, safety, min_scale, max_scale;
    int max_steps;
    int max_rejects;        // adaptive family only (adaptive/ordinary.rs:185-197)
    // t_eval: `rows` = the points that can ever be emitted, in integration order (host-filtered, see deb_api.cu)
    const double* t_rows;   // device [n_rows]
    int n_rows;
    int row_stride;         // rows per trajectory in y_eval (= problem n_eval)
    int emit_t0;            // rows[0] == t0: emitted by the solout call that precedes the loop (solve_ivp.rs:160)
    int even;               // EvenSolout (src/solout/even.rs): rows are t0 + k*dt, always interpolated; the LAST entry of
                            // t_rows is a sentinel equal to tf that triggers the final-point rule (even.rs:166-188)
    double even_tol;        // |dt|*1e-12 + 10 eps (even.rs:92-93)
    // per-step recorders (step_recorder.cuh; kernels instantiated with REC = true): rows carry their own time
    int rec_mode;           // 0, or DEB_SOLOUT_DEFAULT / DENSE / CROSSING
    int dense_n;
    int cross_component, cross_direction;
    double cross_threshold;
    double* t_out;          // [n_traj][row_stride] or null
    // event detection wrapped around the recorder (EventWrappedSolout, src/solout/event.rs); REC kernels with an Evt functor
    int event_direction;    // 0 both, +1 positive, -1 negative
    int event_terminate;    // stop after this many events (0 = never)
    double event_coef[DEB_MAX_DIM + 2];  // EvtLinear: g = c0 + c1*t + sum c[2+i]*y[i]
    // fixed-step kernels: the step schedule does not depend on the trajectory (same t0, tf, h for all), so the host runs the
    // solve_ode bookkeeping (solve_ivp.rs:193-209, :263; fixed/ordinary.rs:16-56, :66-75) once: fx_n_steps steps, all of
    // size h0 (|tf-t0|/100 when h0 = 0) except the last fx_n_tail ones, whose sizes are fx_h_tail[0..fx_n_tail) -- the clip
    // at tf can fire more than once (t + (tf - t) may miss tf by an ulp, which exceeds 10 eps for |tf| > 10: one more tiny
    // step).  fx_n_tail >= 1 whenever fx_n_steps >= 1.  Ends with status fx_status (BAD_INPUT: no steps)
    int fx_n_steps;
    int fx_n_tail;
    double fx_h_tail[DEB_FX_MAX_TAIL];
    int fx_status;
    // HyperplaneCrossingSolout: signed distance of the extracted components to the plane (normal already normalised)
    int plane_dim;
    int plane_index[DEB_MAX_DIM];
    double plane_point[DEB_MAX_DIM], plane_normal[DEB_MAX_DIM];
    double* y_eval;
    int* n_emitted;
    double* t_final;
    double* y_final;
    int* status;
    int* accepted;
    int* rejected;
    int* evals;
    unsigned long long* queue;  // next unclaimed trajectory index
    // Completion watermark for the streamed result copy of DEB_MEM_HOST calls (deb_api.cu): wm_done[b] counts the
    // written-out trajectories of block b = traj >> wm_shift; the lane that completes a block raises wm_ready[b] (pinned
    // host memory the host polls) behind a system-scope fence, and the host copies finished blocks out while the kernel
    // is still integrating the rest.  Null: no watermark.
    int* wm_done;
    int* wm_ready;
    int wm_shift;
    int rows_vec;  // y_eval row groups (RowStage) start 32-byte aligned: whole-sector 16-byte vector stores are safe
    // step-size filter (erk/mod.rs:225): 0 = identity, else h = from_bits(to_bits(h) & filter_mask) (mantissa truncation)
    unsigned long long filter_mask;
    // per-step recorders: a lane whose accepted step needs a refinement (crossing / event root search, interpolated rows)
    // waits until this many lanes of its warp need one, then they all refine in the same iteration (0: refine on the spot)
    int tout_vec;  // per-step recorders: groups of four t_out entries start 32-byte aligned
    int rec_park;
    int rec_park_rows;  // EXPERIMENT (DEB_REC_PARK_ROWS): interpolated t_eval / even rows of a recorder kernel gather too
};

// see OdeKernelArgs::wm_done.  Called by the lane that has just written every output of trajectory `traj`.
__device__ __forceinline__ void wm_publish(const OdeKernelArgs& a, long long traj) {
    if (a.wm_done) {
        __threadfence();  // this trajectory's rows and final fields are visible device-wide before it is counted
        const long long b = traj >> a.wm_shift;
        const long long left = a.n_traj - (b << a.wm_shift);
        const long long full = 1ll << a.wm_shift;
        const int expected = (int)(left < full ? left : full);
        if (atomicAdd(a.wm_done + b, 1) == expected - 1) {
            __threadfence_system();  // cumulative: everything the other lanes fenced before their count is ordered before the flag
            *(volatile int*)(a.wm_ready + b) = 1;
        }
    }
}

// FILTER kernels only (compiled at first use: the ahead-of-time instantiations carry no filter code at all)
__device__ __forceinline__ double apply_filter(const OdeKernelArgs& a, double h) {
    return __longlong_as_double((long long)((unsigned long long)__double_as_longlong(h) & a.filter_mask));
}

// Row staging for the t_eval / even(dt) recorders ("output rows written to HBM as coalesced, vectorised stores").
// A trajectory's rows form one contiguous block of y_eval (row_stride * N doubles) and arrive one at a time, each an
// 8N-byte piece: written directly, a 24-byte Lorenz row covers parts of two 32-byte sectors, and because the block of a
// trajectory stays open for its whole lifetime (~100 rows over ~1500 steps, 95 k trajectories in flight: more than L2)
// the partial sectors reach DRAM as read-modify-write (measured in round 1: 1.59x the algorithmic traffic).
// Instead each lane collects ROWG = 4/gcd(N,4) consecutive rows -- a multiple of 32 bytes -- in shared memory and
// writes the group at once: whole sectors, 16-byte vector stores, no fill reads.  Groups are aligned to the start of
// the trajectory's block; `rows_vec` says whether that is 32-byte aligned (row_stride*N % 4 == 0), else scalar stores.
// Writes one complete row group from a lane's column of the staging buffer (element e at src[e * stride]).  Deliberately
// NOT inlined: the kernels that call it keep ~90 registers of trajectory state live across the call site, and inlining the
// group copy there made ptxas spill loop-carried state of the hot loop (measured: 8 local-memory instructions per step
// attempt); as a call, only the call site pays.
// (A 256-bit `st.global.v4.f64` per sector was tried in round 2: ptxas 12.9 turned it into a 64-bit store in some of the per-kernel
// clones of this function -- tools/microbench/st256_check.cu shows the instruction itself works -- so the stores stay 128-bit.)
template <int CNT>
__device__ __noinline__ void store_row_group(double* dst, const double* src, int stride, int vec) {
    if (vec) {
#pragma unroll
        for (int e = 0; e < CNT; e += 2) {
#ifdef DEB_VAR_STCS  // EXPERIMENT: evict-first stores, so that the row stream does not push the finals' partial sectors out of L2
            __stcs(reinterpret_cast<double2*>(dst + e), make_double2(src[e * stride], src[(e + 1) * stride]));
#else
            *reinterpret_cast<double2*>(dst + e) = make_double2(src[e * stride], src[(e + 1) * stride]);
#endif
        }
    } else {
#pragma unroll
        for (int e = 0; e < CNT; e++) dst[e] = src[e * stride];
    }
}

// where the rows of an ensemble go: base of y_eval, rows per trajectory, and whether row groups start 32-byte aligned
struct RowSink {
    double* y_eval;
    int row_stride;
    int rows_vec;
};
__device__ __forceinline__ RowSink row_sink(const OdeKernelArgs& a) { return RowSink{a.y_eval, a.row_stride, a.rows_vec}; }

template <int N, int BLOCK, bool ON = true>
struct RowStage {
    static constexpr int G4 = (N % 4 == 0) ? 4 : (N % 2 == 0) ? 2 : 1;  // gcd(N, 4)
    static constexpr int ROWG = 4 / G4;
    static constexpr bool ENABLED = ON && (ROWG > 1) && (ROWG * N * BLOCK * 8 <= 16 * 1024);
    static constexpr int SLOTS = ENABLED ? ROWG * N : 1;

    // row w of trajectory `traj`; `emitted` = rows emitted before this call (w == emitted for an append; w == emitted - 1
    // when EvenSolout replaces its last point, even.rs:166-188)
    // buf: the lane's warp-private staging rows, buf[slot][lane] (same base register as the parked-step stash)
    __device__ __forceinline__ static void put(const RowSink& a, double (*buf)[32], unsigned lane, long long traj, int w, int emitted,
                                               const double (&row)[N]) {
        if constexpr (ENABLED) {
            if (w / ROWG == emitted / ROWG) {  // the group being collected
                const int s0 = (w % ROWG) * N;
#pragma unroll
                for (int c = 0; c < N; c++) buf[s0 + c][lane] = row[c];
                if (w % ROWG == ROWG - 1)  // group complete
                    store_row_group<ROWG * N>(a.y_eval + ((size_t)traj * a.row_stride + (size_t)(w - (ROWG - 1))) * N, &buf[0][lane], 32, a.rows_vec);
                return;
            }
        }
        double* dst = a.y_eval + ((size_t)traj * a.row_stride + w) * N;
        if ((N % 2 == 0) && a.rows_vec) {
#pragma unroll
            for (int c = 0; c + 1 < N; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(row[c], row[c + 1]);
        } else {
#pragma unroll
            for (int c = 0; c < N; c++) dst[c] = row[c];
        }
    }

    // component-wise variant of put: set() every component of row w, then done()
    __device__ __forceinline__ static void set(const RowSink& a, double (*buf)[32], unsigned lane, long long traj, int w, int emitted,
                                               int c, double v) {
        if (ENABLED && w / ROWG == emitted / ROWG) buf[(w % ROWG) * N + c][lane] = v;
        else a.y_eval[((size_t)traj * a.row_stride + w) * N + c] = v;
    }
    __device__ __forceinline__ static void done(const RowSink& a, double (*buf)[32], unsigned lane, long long traj, int w, int emitted) {
        if constexpr (ENABLED) {
            if (w / ROWG == emitted / ROWG && w % ROWG == ROWG - 1)
                store_row_group<ROWG * N>(a.y_eval + ((size_t)traj * a.row_stride + (size_t)(w - (ROWG - 1))) * N, &buf[0][lane], 32, a.rows_vec);
        }
    }

    // the trajectory has ended with `emitted` rows: write the rows of the incomplete last group
    __device__ __forceinline__ static void finish(const RowSink& a, double (*buf)[32], unsigned lane, long long traj, int emitted) {
        if constexpr (ENABLED) {
            const int first = (emitted / ROWG) * ROWG;
            double* dst = a.y_eval + ((size_t)traj * a.row_stride + first) * N;
            const int n = (emitted - first) * N;
            for (int e = 0; e < n; e++) dst[e] = buf[e][lane];
        }
    }
    // the same, from the kernel arguments of the ODE kernels
    __device__ __forceinline__ static void put(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, int w, int emitted,
                                               const double (&row)[N]) { put(row_sink(a), buf, lane, traj, w, emitted, row); }
    __device__ __forceinline__ static void set(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, int w, int emitted,
                                               int c, double v) { set(row_sink(a), buf, lane, traj, w, emitted, c, v); }
    __device__ __forceinline__ static void done(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, int w, int emitted) {
        done(row_sink(a), buf, lane, traj, w, emitted);
    }
    __device__ __forceinline__ static void finish(const OdeKernelArgs& a, double (*buf)[32], unsigned lane, long long traj, int emitted) {
        finish(row_sink(a), buf, lane, traj, emitted);
    }
};

__device__ __forceinline__ double d_signum(double x) { return (x != x) ? x : copysign(1.0, x); }  // f64::signum

// utils.rs:22-33
__device__ __forceinline__ double constrain_step_size(double h, double h_min, double h_max) {
    const double sign = d_signum(h);
    if (fabs(h) < h_min) return sign * h_min;
    if (fabs(h) > h_max) return sign * h_max;
    return h;
}

// utils.rs:60-157 -- true when every check passes
__device__ __forceinline__ bool validate_step_size_parameters(double h0, double h_min, double h_max, double t0, double tf) {
    if (tf == t0) return false;
    const double sign = d_signum(tf - t0);
    if (d_signum(h0) != sign) return false;
    if (h_min < 0.0) return false;
    if (h_max < 0.0) return false;
    if (h_min > h_max) return false;
    if (fabs(h0) < h_min) return false;
    if (fabs(h0) > h_max) return false;
    if (fabs(h0) > fabs(tf - t0)) return false;
    if (h0 == 0.0) return false;
    return true;
}

// InitialStepSize::<Ordinary>::compute, h_init.rs:45-135.  f0 = f(t0,y0) is returned in f0 (2 RHS evaluations).
template <class Sys>
__device__ __forceinline__ double initial_step_size(double t0, double tf, const double* y0, const double* p, int order,
                                                 const double* rtol, const double* atol, double h_min, double h_max,
                                                 const deb_pow_tables tb) {
    constexpr int N = Sys::DIM;
    const double posneg = d_signum(tf - t0);
    double f0[N], f1[N], sk[N], y1[N];
    Sys::rhs(t0, y0, f0, p);
    double dnf = 0.0, dny = 0.0;
#pragma unroll
    for (int c = 0; c < N; c++) {
        sk[c] = atol[c] + rtol[c] * fabs(y0[c]);
        const double a = f0[c] / sk[c];
        dnf = dnf + a * a;
        const double b = y0[c] / sk[c];
        dny = dny + b * b;
    }
    double h;
    if (dnf <= 1.0e-10 || dny <= 1.0e-10) h = 1.0e-6;
    else h = sqrt(dny / dnf) * 0.01;
    h = fmin(h, h_max);
    h = h * posneg;
#pragma unroll
    for (int c = 0; c < N; c++) y1[c] = y0[c] + h * f0[c];
    Sys::rhs(t0 + h, y1, f1, p);
    double der2 = 0.0;
#pragma unroll
    for (int c = 0; c < N; c++) {
        const double d = (f1[c] - f0[c]) / sk[c];
        der2 = der2 + d * d;
    }
    der2 = sqrt(der2) / fabs(h);
    const double der12 = fmax(sqrt(dnf), der2);
    double h1;
    if (der12 <= 1.0e-15) h1 = fabs(h) * fmax(1.0e-3, 1.0e-6);  // precedence as written, h_init.rs:116-120
    else h1 = deb_pow_pos(0.01 / der12, 1.0 / (double)order, tb);
    const double interval = fabs(tf - t0);
    h = fmin(fmax(fmin(fmin(fabs(h) * 100.0, h1), h_max), h_min), interval);
    return h * posneg;
}

__device__ __forceinline__ void load_pow_tables(double* s_powlog, unsigned long long* s_exp) {
    for (int i = threadIdx.x; i < 384; i += blockDim.x) s_powlog[i] = deb_c_powlog_tab[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exp[i] = deb_c_exp_tab[i];
    __syncthreads();
}

}  // namespace deb

#include "step_recorder.cuh"

namespace deb {

// ------------------------------------------------------------------------------------------------------------
// Dormand-Prince family (DOPRI5, DOP853): adaptive step, embedded error norm, I-controller, dense output.
// ------------------------------------------------------------------------------------------------------------
// Control-flow design.  The kernel alternates between a HOT LOOP and a SERVICE SECTION.
//   Hot loop: one step attempt per lane per iteration.  Every lane of the warp executes the same instruction stream
//   -- stages, error norm, controller, the derivative at the new point -- whether it holds a live trajectory, is about
//   to reject, or is idle (idle lanes carry a benign dummy state; nothing they compute is committed).  Accept/reject,
//   the state shift and the step-size update are predicated.  The loop runs until some lane needs service (one vote
//   per iteration).  Only two short blocks diverge inside it: the stiffness test (every 100th step of a lane) and
//   the parking of a t_eval hit (below).
//   Service section (every ~50 iterations): flush parked t_eval rows, write out finished trajectories, refill idle
//   lanes from the global queue with ONE warp-aggregated atomicAdd (__ballot_sync + __popc + __shfl_sync), leave
//   when the queue is drained and no lane is live.  Keeping these rare blocks out of the hot loop also keeps their
//   register shuffling (phi moves) out of it.
// Parked emission (methods whose dense output needs no extra stages, i.e. DOPRI5): with 32 lanes, some lane has a
//   t_eval point inside its step in ~40% of the iterations; interpolating on the spot ran ~170 instructions with one
//   or two lanes active (measured: 10% of all issued instructions at 4% lane utilisation).  Instead the lane parks the
//   vectors the interpolant needs (t, h, y, y_new, k0, f(y_new)) in shared memory and keeps stepping; the rows of all
//   parked lanes are computed together in the service section.  A lane that hits again while parked does not commit
//   its step, asks for service and redoes the attempt afterwards -- same inputs, same bits.  The arithmetic is
//   unchanged: same operands, same operations, later.  Methods whose dense output needs extra stages (DOP853, the Verner pairs)
//   park the WHOLE step -- the S stage vectors too -- in dynamic shared memory; the extra stages and the rows of all parked lanes
//   are evaluated together by flush_dense_parked in the service section.
// SHARED_P: every trajectory uses the same parameter set (passed by value in a.pc: no registers).
// One attempt of the adaptive family with EVERY tableau term, zero coefficients included, exactly as the reference loops
// (adaptive/ordinary.rs:95-126): 0 * inf = NaN and all.  Out of line and through memory on purpose: it runs only for an
// attempt whose error terms are not finite.  k[0] = f(t, y) on entry; k[1..S-1], y_new and the error norm on exit.
template <class Sys, class Tab>
__device__ __noinline__ void all_terms_attempt(const double* y, double* k, double t, double h, const double* p, const double* rtol,
                                               const double* atol, double* ynew, double* err_out) {
    constexpr int N = Sys::DIM, S = Tab::S;
    for (int i = 1; i < S; i++) {
        double ys[N];
        for (int c = 0; c < N; c++) ys[c] = y[c];
        for (int j = 0; j < i; j++) {
            const double ah = Tab::av(i, j) * h;
            for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * k[j * N + c];
        }
        Sys::rhs(t + Tab::cv(i) * h, ys, k + i * N, p);
    }
    double ylow[N];
    for (int c = 0; c < N; c++) { ynew[c] = y[c]; ylow[c] = y[c]; }
    for (int i = 0; i < S; i++) {
        const double bw = Tab::bv(i) * h;
        for (int c = 0; c < N; c++) ynew[c] = ynew[c] + bw * k[i * N + c];
    }
    for (int i = 0; i < S; i++) {
        const double bw = Tab::bhv(i) * h;
        for (int c = 0; c < N; c++) ylow[c] = ylow[c] + bw * k[i * N + c];
    }
    double err = 0.0;
    for (int c = 0; c < N; c++) {
        const double sk = atol[c] + rtol[c] * fmax(fabs(y[c]), fabs(ynew[c]));
        err = fmax(err, fabs((ynew[c] - ylow[c]) / sk));
    }
    *err_out = err;
}

// Rows [r0, r1) of the t_eval / even(dt) plan lie in the accepted step (t, y) -> (t + h, ynew); k[0] = f(t, y), k[1..S-1] are the
// stages of the step, dydt = f(t + h, ynew).  Evaluates the method's dense output exactly as its `interpolate` does --
//     Dormand-Prince family   cont[0..3], the extra stages S+1..I-1 and cont[4..O-1] (ordinary.rs:196-234), nested polynomial (:301-337)
//     Verner pairs            the I - S extra stages (adaptive/ordinary.rs:145-160; the reference evaluates and counts them on EVERY
//                             accepted step, only a step that emits a row reads them), Horner in s (:246-277)
//     otherwise               cubic Hermite (interpolate.rs:40-60)
// -- and hands the rows to the row staging.  Returns the number of rows emitted so far: r1, unless the EvenSolout tf sentinel
// (the last plan entry) is among them, whose final-point rule (even.rs:166-188) decides.
template <class Sys, class Tab, class RowsT>
__device__ __forceinline__ int emit_dense_rows(const OdeKernelArgs& a, double (*s_rows)[32], unsigned lane, long long traj, double t, double h,
                                               const double (&y)[Sys::DIM], const double (&ynew)[Sys::DIM], const double (&k)[Tab::S][Sys::DIM],
                                               const double (&dydt)[Sys::DIM], const double* p, int r0, int r1) {
    constexpr int N = Sys::DIM, S = Tab::S, I = Tab::I, O = Tab::O;
    const double t_new = t + h;
    double c1[N], c2[N], c3[N];
    double ch[(O > 4) ? (O - 4) : 1][N];  // DP family: cont[4..O-1]
    double kx[(I > S) ? (I - S) : 1][N];  // stage vectors k[S..I-1]
    if constexpr (Tab::BI_POLY) {
#pragma unroll
        for (int i = S; i < I; i++) {
            double ys[N];
#pragma unroll
            for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
            for (int j = 0; j < i; j++) {
                if (Tab::a(i, j) != 0.0) {
                    const double ah = Tab::av(i, j) * h;
#pragma unroll
                    for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * ((j < S) ? k[(j < S) ? j : 0][c] : kx[(j >= S) ? (j - S) : 0][c]);
                }
            }
            Sys::rhs(t + Tab::cv(i) * h, ys, kx[i - S], p);
        }
    } else if constexpr (Tab::DP) {
#pragma unroll
        for (int c = 0; c < N; c++) {  // ordinary.rs:196-207
            c1[c] = ynew[c] - y[c];
            c2[c] = __dadd_rn(0.0, h * k[0][c]) - c1[c];
            c3[c] = (c1[c] + (-h) * dydt[c]) - c2[c];
        }
        // extra dense stages, ordinary.rs:210-225: k[S] = dydt, stages S+1..I-1
#pragma unroll
        for (int c = 0; c < N; c++) kx[0][c] = dydt[c];
#pragma unroll
        for (int i = S + 1; i < I; i++) {
            double ys[N];
#pragma unroll
            for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
            for (int j = 0; j < i; j++) {
                if (Tab::a(i, j) != 0.0) {
                    const double ah = Tab::av(i, j) * h;
#pragma unroll
                    for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * ((j < S) ? k[(j < S) ? j : 0][c] : kx[(j >= S) ? (j - S) : 0][c]);
                }
            }
            Sys::rhs(t + Tab::cv(i) * h, ys, kx[i - S], p);
        }
#pragma unroll
        for (int i = 4; i < O; i++) {  // ordinary.rs:228-234
#pragma unroll
            for (int c = 0; c < N; c++) ch[i - 4][c] = 0.0;
#pragma unroll
            for (int j = 0; j < I; j++) {
                if (Tab::bi(i, j) != 0.0) {
#pragma unroll
                    for (int c = 0; c < N; c++)
                        ch[i - 4][c] = __dadd_rn(ch[i - 4][c], Tab::biv(i, j) * ((j < S) ? k[(j < S) ? j : 0][c] : kx[(j >= S) ? (j - S) : 0][c]));
                }
            }
#pragma unroll
            for (int c = 0; c < N; c++) ch[i - 4][c] = ch[i - 4][c] * h;
        }
    }
    for (int r = r0; r < r1; r++) {
        if (a.even && r == a.n_rows - 1) {  // the tf sentinel: final-point rule, even.rs:166-188
            int w = -1;
            if (t_new == a.tf) {
                const double t_last = a.t_rows[r - 1];  // r >= 1: row 0 (t0) was emitted at init
                w = (fabs(t_last - a.tf) <= a.even_tol) ? r - 1 : r;  // pop + push(tf, y): replace the near-duplicate; or push(tf, y)
            }
            if (w >= 0) RowsT::put(a, s_rows, lane, traj, w, r, ynew);
            return (w == r) ? r + 1 : r;  // the sentinel slot counts only if it was written
        }
        const double te = a.t_rows[r];
        double row[N];
        if (te == t_new && !a.even) {  // exact hit: the solver state itself (t_eval.rs:113-114); EvenSolout always interpolates
#pragma unroll
            for (int c = 0; c < N; c++) row[c] = ynew[c];
        } else if constexpr (Tab::BI_POLY) {  // adaptive/ordinary.rs:246-277: Horner in s over bi[i][0..O-1], times s
            const double sx = (te - t) / h;
#pragma unroll
            for (int c = 0; c < N; c++) row[c] = y[c];
#pragma unroll
            for (int i = 0; i < I; i++) {
                if (Tab::bi_row(i)) {  // an all-zero row adds (+0 * h) * k[i]
                    double ci = Tab::biv(i, O - 1);
#pragma unroll
                    for (int j = O - 2; j >= 0; j--) ci = ci * sx + Tab::biv(i, j);
                    ci = ci * sx;
                    const double w = ci * h;
#pragma unroll
                    for (int c = 0; c < N; c++) row[c] = row[c] + w * ((i < S) ? k[(i < S) ? i : 0][c] : kx[(i >= S) ? (i - S) : 0][c]);
                }
            }
        } else if constexpr (!Tab::DP) {  // cubic Hermite (adaptive family without bi; only wide systems get here)
            const double hh = t_new - t;
            const double sx = (te - t) / hh;
            const double s2 = sx * sx, s3 = s2 * sx;
            const double h00 = 2.0 * s3 - 3.0 * s2 + 1.0;
            const double h10 = s3 - 2.0 * s2 + sx;
            const double h01 = -2.0 * s3 + 3.0 * s2;
            const double h11 = s3 - s2;
            const double w10 = h10 * hh, w11 = h11 * hh;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double v = __dadd_rn(0.0, h00 * y[c]);
                v = v + w10 * k[0][c];
                v = v + h01 * ynew[c];
                v = v + w11 * dydt[c];
                row[c] = v;
            }
        } else {  // interpolate, ordinary.rs:301-337, factor order as written
            const double sx = (te - t) / h;
            const double s1 = 1.0 - sx;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double accp = (O > 4) ? ch[O - 5][c] : c3[c];
#pragma unroll
                for (int i = O - 2; i >= 1; i--) {
                    double factor;
                    if (i >= 4) factor = (((O - 1) - i) % 2 == 1) ? s1 : sx;
                    else factor = (i % 2 == 1) ? s1 : sx;
                    const double ci = (i >= 4) ? ch[(i >= 4) ? (i - 4) : 0][c] : (i == 3 ? c3[c] : (i == 2 ? c2[c] : c1[c]));
                    accp = accp * factor + ci;
                }
                row[c] = y[c] + sx * accp;
            }
        }
        RowsT::put(a, s_rows, lane, traj, r, r, row);
    }
    return r1;
}

// Parked emission for the methods whose dense output needs extra stages (DOP853, the Verner pairs): a lane whose step contains
// t_eval points stores the whole step -- t, h, y, y_new, f(y_new) and the S stage vectors -- in dynamic shared memory and keeps
// stepping; in the service section every parked lane of the warp runs this function TOGETHER (extra stages, cont / Horner rows), instead
// of one or two lanes doing it on the spot while the other thirty wait (round 1: rkv989e 600 ms against rkv988e 338 ms per 1 M
// trajectories).  Slot layout: [t][h][y: N][y_new: N][dydt: N][k: S x N], each slot 32 lanes wide.  Out of line on purpose: its ~200
// registers of working set must not shape the register allocation of the hot loop.
template <class Sys, class Tab>
__host__ __device__ constexpr int dense_park_slots() { return 2 + Sys::DIM * (Tab::S + 3); }

template <class Sys, class Tab, class RowsT>
__device__ __noinline__ int flush_dense_parked(const OdeKernelArgs& a, const double* st, double (*s_rows)[32], unsigned lane, long long traj,
                                               int r0, int r1) {
    constexpr int N = Sys::DIM, S = Tab::S, NP = Sys::NP;
    double y[N], ynew[N], dydt[N], k[S][N], p[NP > 0 ? NP : 1];
    const double t = st[0 * 32], h = st[1 * 32];
#pragma unroll
    for (int c = 0; c < N; c++) {
        y[c] = st[(2 + c) * 32];
        ynew[c] = st[(2 + N + c) * 32];
        dydt[c] = st[(2 + 2 * N + c) * 32];
    }
#pragma unroll
    for (int i = 0; i < S; i++) {
#pragma unroll
        for (int c = 0; c < N; c++) k[i][c] = st[(2 + 3 * N + i * N + c) * 32];
    }
#pragma unroll
    for (int q = 0; q < NP; q++) p[q] = a.params ? a.params[traj * a.params_stride + q] : a.pc[q];
    return emit_dense_rows<Sys, Tab, RowsT>(a, s_rows, lane, traj, t, h, y, ynew, k, dydt, p, r0, r1);
}

// dynamic shared memory of dp_ensemble_kernel<Sys, Tab, BLOCK, .., REC>: the parked steps of the methods with extra dense stages
template <class Sys, class Tab, int BLOCK, bool REC>
__host__ __device__ constexpr bool dense_park_enabled() {
    return (Tab::I > Tab::S) && !REC && ((long long)dense_park_slots<Sys, Tab>() * BLOCK * 8 <= 64 * 1024);
}
template <class Sys, class Tab, int BLOCK, bool REC>
__host__ __device__ constexpr unsigned dp_dynamic_smem_bytes() {
    return dense_park_enabled<Sys, Tab, BLOCK, REC>() ? (unsigned)(dense_park_slots<Sys, Tab>() * BLOCK * 8) : 0u;
}

// REC: the output goes through a per-step recorder (step_recorder.cuh) instead of the t_eval / even(dt) row plan.
// FILTER: the step-size filter hook is not the identity (a.filter_mask).
template <class Sys, class Tab, int BLOCK, int MIN_BLOCKS, bool SHARED_P, bool REC = false, class Evt = EvtNone, bool FILTER = false>
#ifdef DEB_VAR_MAXNREG  // EXPERIMENT: an explicit register cap instead of the one derived from MIN_BLOCKS
__global__ void __maxnreg__(DEB_VAR_MAXNREG) dp_ensemble_kernel(const __grid_constant__ OdeKernelArgs a) {
#else
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) dp_ensemble_kernel(const __grid_constant__ OdeKernelArgs a) {
#endif
    constexpr int N = Sys::DIM, NP = Sys::NP, S = Tab::S, I = Tab::I, O = Tab::O;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double s_powlog[384];
    __shared__ unsigned long long s_exp[256];
    load_pow_tables(s_powlog, s_exp);
    deb_pow_tables tb;
    tb.powlog = s_powlog;
    tb.exptab = s_exp;

    const unsigned lane = threadIdx.x & 31u;
    const double t0 = a.t0, tf = a.tf;
    const double dir = d_signum(tf - t0);
    const double eps10 = DBL_EPSILON * 10.0;
    const double neg_err_exp = -(1.0 / (double)O);  // -error_exponent, ordinary.rs:152-154
    const double te_none = (dir > 0.0) ? (1.0 / 0.0) : -(1.0 / 0.0);
    // init: f(t0,y0) (+2 for the automatic initial step; the adaptive family counts those two twice, adaptive/ordinary.rs:24-29)
    const int evals_base = (a.h0 == 0.0) ? (Tab::DP ? 3 : 5) : 1;
    const bool bounded_h = (a.h_min > 0.0) || (a.h_max < 1.0 / 0.0);  // constrain_step_size can change h at all

    constexpr int NSTASH = 2 + 4 * N;  // t, h, y, y_new, k0, f(y_new)
    // dense output needs no extra stages: emission can be parked -- as long as the stash fits the static shared-memory
    // budget next to the pow tables (wide systems interpolate on the spot)
    constexpr bool DEFER = (I == S) && !REC && (NSTASH * BLOCK * 8 <= 40 * 1024);
    using Recorder = StepRecorder<Sys, Tab, Evt, BLOCK>;
    Recorder recd;
    using Rows = RowStage<N, BLOCK, !REC>;  // (per-step recorders write their rows themselves, step_recorder.cuh)
    // per warp: [parked step (DEFER)] [row group being collected]; one array so that both are addressed from one base register
    constexpr int NPARK = DEFER ? NSTASH : 0;
    constexpr int ROW_SLOTS = REC ? Recorder::STAGE_SLOTS : Rows::SLOTS;  // recorder kernels stage (y, t) rows, see StepRecorder::push
    __shared__ double s_lane[BLOCK / 32][NPARK + ROW_SLOTS][32];
    double (*stash)[32] = s_lane[threadIdx.x >> 5];
    double (*s_rows)[32] = stash + NPARK;
    // methods with extra dense stages park the whole step in dynamic shared memory (flush_dense_parked)
    constexpr bool DEFER2 = dense_park_enabled<Sys, Tab, BLOCK, REC>();
    constexpr int NPARK2 = dense_park_slots<Sys, Tab>();
    extern __shared__ double s_dyn[];
    double (*dstash)[32] = reinterpret_cast<double (*)[32]>(s_dyn) + (DEFER2 ? (threadIdx.x >> 5) * NPARK2 : 0);
    const bool want_rows = (a.y_eval != nullptr);
    bool pending = false;  // this lane has a parked step
    // per-step recorders: lanes whose step needs a refinement gather (OdeKernelArgs::rec_park); warp-uniform
    bool rec_go = !REC || a.rec_park <= 0;
    int rec_idle = 0;  // lane-iterations spent waiting since the last refinement (warp-uniform)
    int pend_idx = 0;      // first row of the parked step

    // per-lane trajectory state (registers); starts as the idle dummy
    bool active = false, exhausted = false;
    long long traj = 0;
    double t = t0, h = 0.0, h_prev = 0.0, te = te_none;
    double y[N], k[S][N], p_reg[(NP > 0 && !SHARED_P) ? NP : 1];
    const double* p = SHARED_P ? a.pc : p_reg;
    int acc = 0, rej = 0, stiff = 0, nonstiff = 0;
    int m100 = 0;  // (acc + rej) % 100, kept incrementally
    int idx = 0;   // next t_eval row == rows emitted so far (rows are pre-filtered: every consumed point is emitted)
    int fin = -1;  // >= 0: the trajectory has ended with this status and waits for the service section
    bool rejected_prev = false;
#pragma unroll
    for (int c = 0; c < N; c++) { y[c] = 1.0; k[0][c] = 0.0; }
    if (!SHARED_P) {
#pragma unroll
        for (int q = 0; q < NP; q++) p_reg[q] = 1.0;
    }

    for (;;) {
        // =====================================================================================================
        // SERVICE SECTION
        // =====================================================================================================
        if (DEFER) {
            // ---- flush parked rows.  Everything is re-read from the stash per row and per component (a parked step nearly
            //      always covers ONE row): a few live registers instead of five N-vectors next to the whole trajectory state.
            if (pending) {
                const double ts = stash[0][lane], hs = stash[1][lane];
                const double tn = ts + hs;
                // the parked step covers rows [pend_idx, idx): idx was advanced past the step when the lane parked
                for (int r = pend_idx; r < idx; r++) {
                    int w = r;           // row slot written
                    int mode;            // 0: the state at the end of the step; 1: dense output of the method
                    double sx = 0.0;
                    if (a.even && r == a.n_rows - 1) {  // the tf sentinel: even.rs:166-188, only when the step landed exactly on tf
                        w = -1;
                        if (tn == a.tf) {
                            const double t_last = a.t_rows[r - 1];  // r >= 1: row 0 (t0) was emitted at init
                            if (fabs(t_last - a.tf) <= a.even_tol) w = r - 1;  // pop + push(tf, y): replace the near-duplicate
                            else w = r;                                       // push(tf, y)
                        }
                        mode = 0;
                        idx = w + 1 > r ? r + 1 : r;  // rows emitted so far: the sentinel slot counts only if it was written
                    } else {
                        const double ter = a.t_rows[r];
                        mode = (ter == tn && !a.even) ? 0 : 1;  // exact hit: the solver state itself (t_eval.rs:113-114); EvenSolout always interpolates
                        // DP: (t - t_prev) / h_prev, ordinary.rs:312; Hermite: (t - t0) / (t1 - t0) with t1 - t0 = tn - ts, interpolate.rs:53
                        sx = (ter - ts) / (Tab::DP ? hs : (tn - ts));
                    }
                    if (w >= 0) {
#pragma unroll
                        for (int c = 0; c < N; c++) {
                            const double ysc = stash[2 + c][lane], ync = stash[2 + N + c][lane];
                            double v = ync;
                            if (mode != 0) {
                                const double k0s = stash[2 + 2 * N + c][lane], dys = stash[2 + 3 * N + c][lane];
                                if (Tab::DP) {  // cont, ordinary.rs:196-207; interpolate, ordinary.rs:301-337 with O = 5
                                    const double c1 = ync - ysc;
                                    const double c2 = __dadd_rn(0.0, hs * k0s) - c1;
                                    const double c3 = (c1 + (-hs) * dys) - c2;
                                    const double c4 = __dmul_rn(0.0, hs);  // bi rows 4.. are all zero => cont[4] = (+0) * h
                                    const double s1 = 1.0 - sx;
                                    double accp = c4 * s1 + c3;
                                    accp = accp * sx + c2;
                                    accp = accp * s1 + c1;
                                    v = ysc + sx * accp;
                                } else {  // cubic Hermite on (t_prev, t, y_prev, y, dydt_prev, dydt): interpolate.rs:40-60 via adaptive/ordinary.rs:282-295
                                    const double hh = tn - ts;
                                    const double s2 = sx * sx, s3 = s2 * sx;
                                    const double h00 = 2.0 * s3 - 3.0 * s2 + 1.0;
                                    const double h10 = s3 - 2.0 * s2 + sx;
                                    const double h01 = -2.0 * s3 + 3.0 * s2;
                                    const double h11 = s3 - s2;
                                    const double w10 = h10 * hh, w11 = h11 * hh;
                                    v = __dadd_rn(0.0, h00 * ysc);
                                    v = v + w10 * k0s;
                                    v = v + h01 * ync;
                                    v = v + w11 * dys;
                                }
                            }
                            Rows::set(a, s_rows, lane, traj, w, r, c, v);
                        }
                        Rows::done(a, s_rows, lane, traj, w, r);
                    }
                }
                pending = false;
            }
            __syncwarp();
        }
        if (DEFER2) {
            if (pending) {  // all parked lanes of the warp together
                idx = flush_dense_parked<Sys, Tab, Rows>(a, &dstash[0][lane], s_rows, lane, traj, pend_idx, idx);
                pending = false;
            }
            __syncwarp();
        }
        // ---- finished trajectories: Solution / Error fields
        if (active && fin >= 0) {
            if (!REC && want_rows) Rows::finish(a, s_rows, lane, traj, idx);
            if (a.status) a.status[traj] = fin;
            if (a.t_final) a.t_final[traj] = t;
            if (a.y_final) {
#pragma unroll
                for (int c = 0; c < N; c++) a.y_final[traj * N + c] = y[c];
            }
            if (a.accepted) a.accepted[traj] = acc;
            if (a.rejected) a.rejected[traj] = rej;
            // Evals.function: 1 (+2 for the automatic h0) + (S-1) per attempt + per accepted step: DP family 1 (+ I-S-1 dense
            // stages); dense-polynomial pairs I-S dense stages (+1 unless FSAL), adaptive/ordinary.rs:145-174
            constexpr int PER_ACC = Tab::BI_POLY ? ((I - S) + (Tab::FSAL ? 0 : 1)) : (1 + ((I > S) ? (I - S - 1) : 0));
            if (a.evals) a.evals[traj] = evals_base + (S - 1) * (acc + rej) + acc * PER_ACC;
            if constexpr (REC) recd.finish(a, s_rows, lane, traj);
            if (a.n_emitted) a.n_emitted[traj] = REC ? recd.rows : idx;
            wm_publish(a, traj);
            active = false;
            // idle dummy state: finite, never committed
            t = t0; h = 0.0; h_prev = 0.0; te = te_none;
#pragma unroll
            for (int c = 0; c < N; c++) { y[c] = 1.0; k[0][c] = 0.0; }
        }
        fin = -1;
        __syncwarp();
        // ---- refill idle lanes from the global queue (one atomic per warp); BadInput trajectories are written
        //      out at once and the lane asks again
        for (;;) {
            const unsigned need = __ballot_sync(FULL, !active && !exhausted);
            if (need == 0u) break;
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!active && !exhausted) {
                traj = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (traj >= a.n_traj) {
                    exhausted = true;
                } else {
                    // ---- init, dormandprince/ordinary.rs:16-61
                    double y0v[N];
#pragma unroll
                    for (int c = 0; c < N; c++) y0v[c] = a.y0[traj * N + c];
                    if (!SHARED_P) {
#pragma unroll
                        for (int q = 0; q < NP; q++) p_reg[q] = a.params[traj * a.params_stride + q];
                    }
                    double h0 = a.h0;
                    if (h0 == 0.0) h0 = initial_step_size<Sys>(t0, tf, y0v, p, O, a.rtol, a.atol, a.h_min, a.h_max, tb);
                    if (!validate_step_size_parameters(h0, a.h_min, a.h_max, t0, tf)) {
                        // Err(BadInput): no solution; report (t0, y0)
                        if (a.status) a.status[traj] = DEB_STATUS_BAD_INPUT;
                        if (a.t_final) a.t_final[traj] = t0;
                        if (a.y_final) {
#pragma unroll
                            for (int c = 0; c < N; c++) a.y_final[traj * N + c] = y0v[c];
                        }
                        if (a.accepted) a.accepted[traj] = 0;
                        if (a.rejected) a.rejected[traj] = 0;
                        if (a.evals) a.evals[traj] = 0;
                        if (a.n_emitted) a.n_emitted[traj] = 0;
                        wm_publish(a, traj);
                    } else {
#pragma unroll
                        for (int c = 0; c < N; c++) y[c] = y0v[c];
                        acc = 0; rej = 0; idx = 0; m100 = 0;
                        t = t0;
                        h = FILTER ? apply_filter(a, h0) : h0;  // ordinary.rs:33
                        h_prev = 0.0;
                        stiff = 0; nonstiff = 0;
                        rejected_prev = false;
                        Sys::rhs(t, y, k[0], p);
                        // solout before the loop (solve_ivp.rs:160): emits rows[0] iff it equals t0
                        if (!REC && a.emit_t0) {
                            if (want_rows) Rows::put(a, s_rows, lane, traj, 0, 0, y);
                            idx = 1;
                        }
                        te = (!REC && idx < a.n_rows) ? a.t_rows[idx] : te_none;
                        if constexpr (REC) {  // the solout call that precedes the loop
                            recd.reset();
                            recd.first(a, s_rows, lane, traj, t0, y, p);
                        }
                        active = true;
                    }
                }
            }
            __syncwarp();
        }
        // ---- leave when the queue is drained and no lane holds a live trajectory
        if (!__any_sync(FULL, active)) break;

        // =====================================================================================================
        // HOT LOOP: one step attempt per lane per iteration, until some lane needs service
        // =====================================================================================================
        bool service;
        do {
            // ---- solve_ode loop head, solve_ivp.rs:193-209; step guards, ordinary.rs:70-92
            {
                const double h_new = tf - t;
                if ((t + h - tf) * dir > 0.0) {
                    if (fabs(h_new) < eps10) fin = DEB_STATUS_COMPLETE;
                    else h = FILTER ? apply_filter(a, h_new) : h_new;  // solver.set_h(h_new) filters too (ordinary.rs:288)
                }
                if (fin < 0) {
                    if (fabs(h) < fabs(h_prev) * 1e-14) fin = DEB_STATUS_STEP_SIZE;
                    else if (acc + rej >= a.max_steps) fin = DEB_STATUS_MAX_STEPS;  // steps counts rejected attempts too
                }
            }
            const bool stepping = active && fin < 0;

            // ---- stages, solution and error estimate.  Terms whose tableau coefficient is zero are dropped
            // (x + (0*h)*k == x for finite k).
            double yseg[N], ynew[N];
            const double t_new = t + h;
            double err = 0.0;
            {
                // stages, ordinary.rs:95-104 (all lanes)
#pragma unroll
                for (int i = 1; i < S; i++) {
                    double ys[N];
#pragma unroll
                    for (int c = 0; c < N; c++) ys[c] = y[c];
#pragma unroll
                    for (int j = 0; j < i; j++) {
                        if (Tab::a(i, j) != 0.0) {
                            const double ah = Tab::av(i, j) * h;
#pragma unroll
                            for (int c = 0; c < N; c++) ys[c] = ys[c] + ah * k[j][c];
                        }
                    }
                    Sys::rhs(t + Tab::cv(i) * h, ys, k[i], p);
                }
                if (Tab::DP) {  // dormandprince/ordinary.rs:106-149
                    double es[N];
#pragma unroll
                    for (int c = 0; c < N; c++) { yseg[c] = 0.0; es[c] = 0.0; }
#pragma unroll
                    for (int i = 0; i < S; i++) {
                        if (Tab::b(i) != 0.0) {
#pragma unroll
                            for (int c = 0; c < N; c++) yseg[c] = __dadd_rn(yseg[c], Tab::bv(i) * k[i][c]);
                        }
                        if (Tab::er(i) != 0.0) {
#pragma unroll
                            for (int c = 0; c < N; c++) es[c] = __dadd_rn(es[c], Tab::erv(i) * k[i][c]);
                        }
                    }
                    double err2 = 0.0;
                    double sk[N];
                    err = 0.0;
#pragma unroll
                    for (int c = 0; c < N; c++) {
                        ynew[c] = y[c] + h * yseg[c];
                        // |y|.max(|y_new|), traits.rs:587-595.  Written as a compare + select on the two non-negative values:
                        // same result as fmax whenever y is not NaN (a NaN y_new is dropped, as f64::max does; a NaN y makes
                        // every stage, hence the error norm, NaN whatever sk is)
                        const double ay = fabs(y[c]), an = fabs(ynew[c]);
                        sk[c] = a.atol[c] + a.rtol[c] * ((an > ay) ? an : ay);
                        const double e = es[c] / sk[c];
                        err = err + e * e;
                    }
                    if (Tab::HAS_BH) {  // DOP853 second estimator, ordinary.rs:135-143
                        double e2[N];
#pragma unroll
                        for (int c = 0; c < N; c++) e2[c] = yseg[c];
#pragma unroll
                        for (int i = 0; i < S; i++) {
                            if (Tab::bh(i) != 0.0) {
#pragma unroll
                                for (int c = 0; c < N; c++) e2[c] = e2[c] + (-Tab::bhv(i)) * k[i][c];
                            }
                        }
#pragma unroll
                        for (int c = 0; c < N; c++) {
                            const double e = e2[c] / sk[c];
                            err2 = err2 + e * e;
                        }
                    }
                    double deno = err + 0.01 * err2;
                    if (deno <= 0.0) deno = 1.0;
                    err = fabs(h) * err * sqrt(1.0 / (deno * (double)N));  // ordinary.rs:148
                } else {  // adaptive/ordinary.rs:107-126: y_high, y_low, infinity norm of (y_high - y_low)/sk
                    double ylow[N];
#pragma unroll
                    for (int c = 0; c < N; c++) { ynew[c] = y[c]; ylow[c] = y[c]; yseg[c] = 0.0; }
#pragma unroll
                    for (int i = 0; i < S; i++) {
                        if (Tab::b(i) != 0.0) {
                            const double bw = Tab::bv(i) * h;
#pragma unroll
                            for (int c = 0; c < N; c++) ynew[c] = ynew[c] + bw * k[i][c];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < S; i++) {
                        if (Tab::bh(i) != 0.0) {
                            const double bw = Tab::bhv(i) * h;
#pragma unroll
                            for (int c = 0; c < N; c++) ylow[c] = ylow[c] + bw * k[i][c];
                        }
                    }
                    double esum = 0.0;  // sum of the (non-negative) error terms: inf or NaN iff some term is not finite, or absurdly large
#pragma unroll
                    for (int c = 0; c < N; c++) {  // error_norm_inf, traits.rs:413-434 (f64::max ignores a NaN term, as fmax does)
                        const double sk = a.atol[c] + a.rtol[c] * fmax(fabs(y[c]), fabs(ynew[c]));
                        const double e = fabs((ynew[c] - ylow[c]) / sk);
                        esum = esum + e;
                        err = fmax(err, e);
                    }
                    // A non-finite error term means some stage value overflowed.  The reference then also produces NaN from
                    // 0 * inf in the zero-coefficient terms dropped above, and its infinity norm drops NaN terms -- such a
                    // step can be ACCEPTED (with a NaN state).  Redo the attempt with every term, out of line.  (Dormand-
                    // Prince family: the 2-norm keeps NaN; the step is rejected and the controller sees min_scale either way.)
                    if (!(esum <= DBL_MAX)) {
                        // copies, so that the live register arrays never have their address taken
                        double kk[S][N], yy[N], yn[N], pp[NP > 0 ? NP : 1], rt[N], at[N], ee;
#pragma unroll
                        for (int c = 0; c < N; c++) { yy[c] = y[c]; kk[0][c] = k[0][c]; rt[c] = a.rtol[c]; at[c] = a.atol[c]; }
#pragma unroll
                        for (int q = 0; q < NP; q++) pp[q] = p[q];
                        all_terms_attempt<Sys, Tab>(yy, &kk[0][0], t, h, pp, rt, at, yn, &ee);
#pragma unroll
                        for (int i = 1; i < S; i++) {
#pragma unroll
                            for (int c = 0; c < N; c++) k[i][c] = kk[i][c];
                        }
#pragma unroll
                        for (int c = 0; c < N; c++) ynew[c] = yn[c];
                        err = ee;
                    }
                    __syncwarp();
                }
            }
            // ---- controller, ordinary.rs:151-157
            double scale = a.safety * deb_pow_pos(err, neg_err_exp, tb);
            // scale.max(min_scale).min(max_scale) (ordinary.rs:157) with f64::max/min NaN rules: a NaN scale becomes
            // min_scale.  Written as ordered compares + selects (the options are never NaN): 6 instructions, not 16.
            scale = (scale >= a.min_scale) ? scale : a.min_scale;
            scale = (scale <= a.max_scale) ? scale : a.max_scale;

            // ---- derivative at the new point (ordinary.rs:162; evaluated by every lane, committed only on accept)
            const bool accept = stepping && (err <= 1.0);
            double dydt[N];
            if constexpr (Tab::FSAL) {  // adaptive/ordinary.rs:166-169: the last stage is the derivative at the new point
#pragma unroll
                for (int c = 0; c < N; c++) dydt[c] = k[S - 1][c];
            } else {
                Sys::rhs(t_new, ynew, dydt, p);
            }

            // steps % 100 == 0 with steps = acc + rej + 1: stiffness test, ordinary.rs:165-194.  With parked emission, a lane whose
            // step contains a t_eval point while its row slot is still occupied does not commit this attempt: it is redone, bit
            // for bit, after the service section.  The counters below must not advance for the attempt that is thrown away
            // (the reference advances them once per 100th step).
            // Per-step recorders: an accepted step that needs a root search or interpolated rows is thrown away in the same manner
            // until enough lanes of the warp wait for one (rec_go); the waiting lanes repeat the attempt, bit for bit, beside the
            // lanes that keep stepping, and then refine together instead of one or two lanes at a time.
            bool rec_blocked = false;
            if constexpr (REC) rec_blocked = !rec_go && accept && fin < 0 && recd.slow_work(a, t_new, ynew, p);
            if (Tab::DP && accept && m100 == 99 && !rec_blocked && !((DEFER || DEFER2) && pending && ((te - t_new) * dir <= 0.0))) {
                // ysti = the argument of the last stage; rebuilt here (same operations, same bits) instead of being
                // kept alive in registers through 99 steps out of 100
                double ysti[N];
#pragma unroll
                for (int c = 0; c < N; c++) ysti[c] = y[c];
#pragma unroll
                for (int j = 0; j < S - 1; j++) {
                    if (Tab::a(S - 1, j) != 0.0) {
                        const double ah = Tab::av(S - 1, j) * h;
#pragma unroll
                        for (int c = 0; c < N; c++) ysti[c] = ysti[c] + ah * k[j][c];
                    }
                }
                double stdnum = 0.0, stden = 0.0;
#pragma unroll
                for (int c = 0; c < N; c++) {
                    const double d1 = yseg[c] - k[S - 1][c];
                    stdnum = stdnum + d1 * d1;
                    const double d2 = dydt[c] - ysti[c];  // (sic) derivative minus stage state, as in the reference
                    stden = stden + d2 * d2;
                }
                if (stden > 0.0) {
                    const double h_lamb = h * sqrt(stdnum / stden);
                    if (h_lamb > 6.1) {
                        nonstiff = 0;
                        stiff += 1;
                        if (stiff == 15) fin = DEB_STATUS_STIFFNESS;  // Err before any state update
                    }
                } else {
                    nonstiff += 1;
                    if (nonstiff == 6) stiff = 0;
                }
            }
            __syncwarp();

            // ---- TEvalSolout (t_eval.rs:100-130): does a t_eval point lie in this step?  (te - t_new == 0 iff te == t_new)
            const bool hit = accept && fin < 0 && ((te - t_new) * dir <= 0.0);
            bool blocked = rec_blocked;
            if (DEFER || DEFER2) {
                blocked = hit && pending;  // slot occupied: do not commit, get the slot flushed, redo this attempt
                if (hit && !blocked) {
                    if (want_rows) {  // park the step
                        if constexpr (DEFER2) {
                            dstash[0][lane] = t;
                            dstash[1][lane] = h;
#pragma unroll
                            for (int c = 0; c < N; c++) {
                                dstash[2 + c][lane] = y[c];
                                dstash[2 + N + c][lane] = ynew[c];
                                dstash[2 + 2 * N + c][lane] = dydt[c];
                            }
#pragma unroll
                            for (int i = 0; i < S; i++) {
#pragma unroll
                                for (int c = 0; c < N; c++) dstash[2 + 3 * N + i * N + c][lane] = k[i][c];
                            }
                        } else {
                            stash[0][lane] = t;
                            stash[1][lane] = h;
#pragma unroll
                            for (int c = 0; c < N; c++) {
                                stash[2 + c][lane] = y[c];
                                stash[2 + N + c][lane] = ynew[c];
                                stash[2 + 2 * N + c][lane] = k[0][c];
                                stash[2 + 3 * N + c][lane] = dydt[c];
                            }
                        }
                        pend_idx = idx;
                        pending = true;
                    }
                    while ((te - t_new) * dir <= 0.0) {
                        idx += 1;
                        te = (idx < a.n_rows) ? a.t_rows[idx] : te_none;
                    }
                }
                __syncwarp();
            } else {
                if (hit) {  // interpolate on the spot (per-step-recorder kernels never get here with rows; wide systems do)
                    const int r0 = idx;
                    while ((te - t_new) * dir <= 0.0) {
                        idx += 1;
                        te = (idx < a.n_rows) ? a.t_rows[idx] : te_none;
                    }
                    if (want_rows) idx = emit_dense_rows<Sys, Tab, Rows>(a, s_rows, lane, traj, t, h, y, ynew, k, dydt, p, r0, idx);
                }
                __syncwarp();
            }

            bool interrupt = false;  // ControlFlag::Terminate from an event: the step is kept, then Status::Interrupted
            if constexpr (REC) {  // solout after an accepted step (solve_ivp.rs:239-246)
                if (accept && fin < 0 && !blocked) interrupt = recd.step(a, s_rows, lane, traj, t, h, y, ynew, k, dydt, p);
                __syncwarp();
            }

            // ---- accept: shift (ordinary.rs:237-254) / reject (ordinary.rs:255-258), as predicated updates
            const bool commit = accept && fin < 0 && !blocked;
            const bool reject = stepping && fin < 0 && !accept;
            if (commit) {
                h_prev = h;
                t = t_new;
#pragma unroll
                for (int c = 0; c < N; c++) { y[c] = ynew[c]; k[0][c] = dydt[c]; }
                if (rejected_prev) {
                    scale = (scale <= 1.0) ? scale : 1.0;  // scale.min(1), scale is not NaN here
                    if (!Tab::DP) stiff = 0;               // adaptive/ordinary.rs:137-143
                }
                rejected_prev = false;
                acc += 1;
                m100 = (m100 == 99) ? 0 : m100 + 1;
            }
            if (reject) {
                rejected_prev = true;  // Status::RejectedStep
                if (!Tab::DP) {        // adaptive/ordinary.rs:182-197: max_rejects consecutive rejections => Err(Stiffness)
                    stiff += 1;
                    if (stiff >= a.max_rejects) fin = DEB_STATUS_STIFFNESS;  // Err: the attempt is not counted
                }
                if (fin < 0) { rej += 1; m100 = (m100 == 99) ? 0 : m100 + 1; }
            }
            if ((commit || reject) && fin < 0) {
                // ---- step-size update, ordinary.rs:261-267 (filter = identity)
                h = h * scale;
                if (bounded_h) h = constrain_step_size(h, a.h_min, a.h_max);  // identity for h_min = 0, h_max = inf
                if constexpr (FILTER) h = apply_filter(a, h);                 // ordinary.rs:267
                // accepted: end-of-interval test, solve_ivp.rs:263
                if (commit && fabs(tf - t) <= eps10) fin = DEB_STATUS_COMPLETE;
                if (REC && interrupt) fin = DEB_STATUS_INTERRUPTED;  // solve_ivp.rs:255-260, before the end-of-interval test
            }
            if constexpr (REC) {
                // refine in the next iteration when enough lanes wait, or when no lane of the warp can advance without it
                // (or when the lanes that wait have idled for 8 iterations per lane of the threshold: an ensemble in which few trajectories
                // have events never fills the threshold, and its waiting lanes must not idle until their neighbours finish)
                const unsigned waiting = __ballot_sync(FULL, blocked);
                const unsigned advancing = __ballot_sync(FULL, stepping && fin < 0 && !blocked);
                const int n_wait = __popc(waiting);
                rec_idle = n_wait ? rec_idle + n_wait : 0;
                rec_go = a.rec_park <= 0 || n_wait >= a.rec_park || rec_idle >= 8 * a.rec_park || (waiting != 0u && advancing == 0u);
                service = active && fin >= 0;
            } else {
                service = (active && fin >= 0) || blocked;
            }
        } while (!__any_sync(FULL, service));
    }
}

}  // namespace deb
