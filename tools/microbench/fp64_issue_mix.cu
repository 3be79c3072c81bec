// Microbenchmark: how much of the FP64 pipe's throughput survives when other instructions share the scheduler?
// Model under test: a DADD/DMUL warp instruction keeps the 16-lane FP64 pipe of a scheduler busy for 2 cycles; can the
// scheduler issue an integer / move / select instruction of ANOTHER (or the same) warp in the second cycle, i.e. is the
// cost of F FP64 and O other instructions max(2F, F + O) (perfect overlap) or 2F + O (no overlap)?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_issue_mix fp64_issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

// per unrolled group: 6 FP64 instructions (3 independent chains x (DMUL, DADD)) and NOTHER other instructions of kind KIND
// KIND 0: LOP3 (ALU)   1: IMAD (FMA-lite / int)   2: IADD (ALU)   3: MOV between registers (predicated)
template <int NOTHER, int KIND>
__global__ void __launch_bounds__(1024) mix(double* out, int iters, double m, double c, int dummy) {
    double v0 = 1.0 + 1e-3 * threadIdx.x, v1 = 1.5 + 1e-3 * threadIdx.x, v2 = 2.0 + 1e-3 * threadIdx.x;
    int w[12];
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = threadIdx.x + i + dummy;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            v0 = __dmul_rn(v0, m); v1 = __dmul_rn(v1, m); v2 = __dmul_rn(v2, m);
#pragma unroll
            for (int i = 0; i < NOTHER; i++) {
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i % 12]) : "r"(dummy), "r"(it));
                if (KIND == 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w[i % 12]) : "r"(dummy), "r"(it));
                if (KIND == 2) asm volatile("add.s32 %0, %0, %1;" : "+r"(w[i % 12]) : "r"(dummy));
                if (KIND == 3) asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0; @p mov.b32 %0, %2; }" : "+r"(w[i % 12]) : "r"(dummy), "r"(it));
            }
            v0 = __dadd_rn(v0, c); v1 = __dadd_rn(v1, c); v2 = __dadd_rn(v2, c);
        }
    }
    double s = v0 + v1 + v2;
    int q = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) q += w[i];
    if (s == 123.456 || q == 0x7fffffff) out[0] = s + q;
}

template <int NOTHER, int KIND>
void run(int warps_per_sm, double* d_out, int sms, double peak) {
    const int iters = 3000;
    const int threads = warps_per_sm * 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NOTHER, KIND><<<sms, threads>>>(d_out, 10, 0.9999999, 1e-7, 1);
    cudaEventRecord(e0);
    mix<NOTHER, KIND><<<sms, threads>>>(d_out, iters, 0.9999999, 1e-7, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fp64 = (double)iters * 8 * 6 * threads * sms;
    const double tput = fp64 / (ms * 1e-3);
    const double ratio = NOTHER / 6.0;
    const char* kinds[] = {"LOP3", "IMAD", "IADD", "@P MOV (SETP+MOV: 2 per unit)"};
    printf("{\"other_per_fp64\": %.3f, \"kind\": \"%s\", \"warps_per_sm\": %d, \"dp_inst_per_s\": %.4e, \"frac_of_peak\": %.3f, "
           "\"model_no_overlap\": %.3f, \"model_full_overlap\": %.3f}\n",
           ratio * (KIND == 3 ? 2 : 1), kinds[KIND], warps_per_sm, tput, tput / peak, 2.0 / (2.0 + ratio * (KIND == 3 ? 2 : 1)),
           1.0 / (ratio * (KIND == 3 ? 2 : 1) > 1.0 ? (1.0 + ratio * (KIND == 3 ? 2 : 1)) / 2.0 : 1.0));
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double peak = (double)sms * 64.0 * khz * 1e3 / 32.0 * 32.0;  // 64 DP lanes per SM and clock
    double* d;
    cudaMalloc(&d, 64);
    printf("# peak %.4e DP inst/s (%d SMs x 64 lanes x %.3f GHz)\n", peak, sms, khz * 1e-6);
    for (int w : {20, 32}) {
        run<0, 0>(w, d, sms, peak);
        run<1, 0>(w, d, sms, peak); run<2, 0>(w, d, sms, peak); run<3, 0>(w, d, sms, peak); run<4, 0>(w, d, sms, peak); run<6, 0>(w, d, sms, peak); run<12, 0>(w, d, sms, peak);
        run<2, 1>(w, d, sms, peak); run<4, 1>(w, d, sms, peak); run<6, 1>(w, d, sms, peak);
        run<2, 2>(w, d, sms, peak); run<4, 2>(w, d, sms, peak); run<6, 2>(w, d, sms, peak);
        run<2, 3>(w, d, sms, peak); run<3, 3>(w, d, sms, peak);
    }
    return 0;
}
