// ens_stats.cuh -- per-t_eval ensemble sums {sum y, sum y^2} and counts: the only quantity that crosses GPUs
// (all-reduced by the caller).  HBM-bound single pass over y_eval[n_traj][n_eval][dim]; deterministic
// (per-CTA partials accumulated in trajectory order, then one ordered reduction), so repeated runs give identical bits.
#pragma once
#include <stdint.h>

namespace deb {

// Each CTA owns a contiguous slab of trajectories; thread e owns element e = r*dim + c of a trajectory's block
// (consecutive threads read consecutive doubles: one trajectory = one coalesced 8*ne-byte read).  Eight trajectories
// are loaded before they are accumulated (memory-level parallelism); the accumulation order stays i, i+1, i+2, ...
// partial[(cta*ne + e)*2 + {0,1}], pcount[cta*n_eval + r].
__global__ void __launch_bounds__(512) stats_partial_kernel(const double* __restrict__ y_eval, const int* __restrict__ n_emitted,
                                                           long long n_traj, int n_eval, int dim, double* partial,
                                                           long long* pcount) {
    const int ne = n_eval * dim;
    const long long per = (n_traj + gridDim.x - 1) / gridDim.x;
    const long long b = (long long)blockIdx.x * per;
    const long long e_end = (b + per < n_traj) ? (b + per) : n_traj;
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const int r = e / dim;
        double s = 0.0, s2 = 0.0;
        long long cnt = 0;
        long long i = b;
        for (; i + 8 <= e_end; i += 8) {
            double v[8];
            bool ok[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                ok[u] = n_emitted[i + u] > r;
                v[u] = ok[u] ? y_eval[(i + u) * ne + e] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (ok[u]) { s += v[u]; s2 += v[u] * v[u]; cnt += 1; }
            }
        }
        for (; i < e_end; i++) {
            if (n_emitted[i] > r) {
                const double v = y_eval[i * ne + e];
                s += v;
                s2 += v * v;
                cnt += 1;
            }
        }
        partial[((long long)blockIdx.x * ne + e) * 2 + 0] = s;
        partial[((long long)blockIdx.x * ne + e) * 2 + 1] = s2;
        if (e % dim == 0) pcount[(long long)blockIdx.x * n_eval + r] = cnt;
    }
}

// One warp per element: lanes stride over the CTAs' partials, then a fixed-order shuffle tree.
__global__ void __launch_bounds__(128) stats_final_kernel(const double* partial, const long long* pcount, int n_cta, int n_eval, int dim,
                                                         double* sums, long long* counts, int accumulate) {
    const int ne = n_eval * dim;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= ne) return;
    double s = 0.0, s2 = 0.0;
    long long cnt = 0;
    for (int c = lane; c < n_cta; c += 32) {
        s += partial[((long long)c * ne + e) * 2 + 0];
        s2 += partial[((long long)c * ne + e) * 2 + 1];
        if (e % dim == 0) cnt += pcount[(long long)c * n_eval + e / dim];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {  // accumulate: add to what the previous chunks of the ensemble left (chunk order is fixed: deterministic)
        sums[e * 2 + 0] = accumulate ? sums[e * 2 + 0] + s : s;
        sums[e * 2 + 1] = accumulate ? sums[e * 2 + 1] + s2 : s2;
        if (e % dim == 0) counts[e / dim] = accumulate ? counts[e / dim] + cnt : cnt;
    }
}

}  // namespace deb
