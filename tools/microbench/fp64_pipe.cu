// Microbenchmark of the B200 FP64 pipe: dependent-issue latency, throughput vs (warps per scheduler, ILP), and the
// effect of interleaved non-FP64 instructions.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false
// Run under gpurun; prints a table.  Used to size the ensemble kernel, not part of the product.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MIX>
__global__ void __launch_bounds__(1024) chain(double* out, int iters, double m, double c, int dummy) {
    double v[ILP];
    int w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { v[i] = 1.0 + 1e-3 * (threadIdx.x + i); w[i] = threadIdx.x + i + dummy; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                v[i] = __dmul_rn(v[i], m);
                if (MIX >= 1) w[i] = w[i] * 3 + dummy;          // IMAD
                v[i] = __dadd_rn(v[i], c);
                if (MIX >= 2) w[i] = (w[i] ^ (w[i] >> 3)) + it; // LOP3/SHF/IADD
            }
        }
    }
    long long t1 = clock64();
    double s = 0.0; int q = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) { s += v[i]; q += w[i]; }
    if (s == 123.456 || q == 0x7fffffff) out[0] = s + q;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0);
}

template <int ILP, int MIX>
void run(int warps_per_sm, double* d_out, int sms) {
    const int iters = 4000;
    int threads = warps_per_sm * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<ILP, MIX><<<sms, threads>>>(d_out, 10, 0.9999999, 1e-7, 1);
    cudaEventRecord(e0);
    chain<ILP, MIX><<<sms, threads>>>(d_out, iters, 0.9999999, 1e-7, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    double fp64_inst = (double)iters * 8 * ILP * 2;           // per thread
    double cyc_per_fp64 = h[1] / fp64_inst;                    // per warp, cycles per FP64 instruction issued by that warp
    double tput = fp64_inst * threads * sms / (ms * 1e-3) / 1e12;
    printf("ILP %d MIX %d warps/SM %2d (%.1f/sched): %6.2f cyc per FP64 inst per warp, %6.2f T DP-inst/s (%.0f%% of 18.58)\n",
           ILP, MIX, warps_per_sm, warps_per_sm / 4.0, cyc_per_fp64, tput, tput / 18.58 * 100);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 64);
    printf("== latency: 1 warp per SM, dependent chain\n");
    run<1, 0>(1, d, sms);
    run<2, 0>(1, d, sms);
    run<4, 0>(1, d, sms);
    printf("== throughput vs warps, ILP\n");
    for (int w : {4, 8, 12, 16, 20, 24, 32}) { run<1, 0>(w, d, sms); }
    for (int w : {4, 8, 12, 16, 20, 24, 32}) { run<3, 0>(w, d, sms); }
    printf("== mixed 1 int per FP64 pair / 2 int per pair\n");
    for (int w : {8, 16, 20, 24, 32}) { run<3, 1>(w, d, sms); }
    for (int w : {8, 16, 20, 24, 32}) { run<3, 2>(w, d, sms); }
    for (int w : {16, 20}) { run<1, 2>(w, d, sms); }
    return 0;
}
