// Does st.global.v4.f64 (STG.256 on sm_100) store what it is given?  Direct and through a noinline device function.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ void store4(double* dst, const double* src, int stride, int vec) {
    if (vec) {
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(src[0]), "d"(src[stride]), "d"(src[2 * stride]), "d"(src[3 * stride]) : "memory");
    } else {
        for (int e = 0; e < 4; e++) dst[e] = src[e * stride];
    }
}
__global__ void k(double* out, int vec, int mode) {
    __shared__ double buf[4][32];
    const int lane = threadIdx.x;
    for (int e = 0; e < 4; e++) buf[e][lane] = 100.0 * lane + e + 1;
    __syncwarp();
    double* dst = out + 4 * lane;
    if (mode == 0) store4(dst, &buf[0][lane], 32, vec);
    else asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(buf[0][lane]), "d"(buf[1][lane]), "d"(buf[2][lane]), "d"(buf[3][lane]) : "memory");
}
int main() {
    double* d; cudaMalloc(&d, 128 * 8);
    for (int mode = 0; mode < 2; mode++) for (int vec = 0; vec < 2; vec++) {
        cudaMemset(d, 0, 128 * 8);
        k<<<1, 32>>>(d, vec, mode);
        double h[128]; cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < 128; i++) if (h[i] != 100.0 * (i / 4) + (i % 4) + 1) bad++;
        printf("mode %d vec %d: %s, %d wrong of 128; first: %g %g %g %g | %g %g %g %g\n", mode, vec, cudaGetErrorString(e), bad, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
    return 0;
}
