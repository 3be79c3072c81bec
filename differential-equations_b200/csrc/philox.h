// philox.h -- counter-based Wiener increments for SDE ensembles, generated in registers.
//
// The reference leaves noise to user code (`SDE::noise(&mut self, dt, dw)`, /root/reference/src/sde/sde.rs:67; the
// examples draw Normal(0, sqrt(dt)) from a sequential StdRng, examples/sde/03_ornstein_uhlenbeck/main.rs:51-54).
// A sequential generator cannot serve 1e8 independent paths, so the ensemble front end defines the stream ONCE,
// here, and the host can regenerate any increment from (seed, path, step, component) alone:
//
//   normal number q = step*dim + comp of path p: Philox4x32-10 (Salmon, Moraes, Dror, Shaw; SC'11) with
//       key     = (seed lo32, seed hi32)
//       counter = ((q>>1) lo32, (q>>1) hi32, p lo32, p hi32)            -> words w0..w3
//       a = (w0<<32 | w1) >> 11,  b = (w2<<32 | w3) >> 11               (53 bits each)
//       u1 = (a+1) * 2^-53 in (0,1],   u2 = b * 2^-53 in [0,1)
//       r = sqrt(-2*log(u1)),  theta = (2*pi)*u2
//       z = r*cos(theta) for even q,  r*sin(theta) for odd q            (Box-Muller; one Philox call feeds two normals)
//   dW = sqrt(h) * z
#pragma once
#include <stdint.h>

namespace deb {

#if defined(__CUDACC__)
#define DEB_PHILOX_HD __host__ __device__ __forceinline__
#else
#define DEB_PHILOX_HD inline
#endif

struct Philox4 { uint32_t w[4]; };

DEB_PHILOX_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#if defined(__CUDA_ARCH__)
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

#if defined(__CUDACC__)
// Both Box-Muller normals of pair index `pair` for path `path`.
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t path, uint64_t pair, double* z_even, double* z_odd) {
    const Philox4 o = philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)path, (uint32_t)(path >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = (((uint64_t)o.w[0] << 32) | o.w[1]) >> 11;
    const uint64_t b = (((uint64_t)o.w[2] << 32) | o.w[3]) >> 11;
    const double u1 = (double)(a + 1) * 0x1p-53;
    const double u2 = (double)b * 0x1p-53;
    const double r = sqrt(-2.0 * log(u1));
    const double theta = 6.283185307179586 * u2;
    double sn, cs;
    sincos(theta, &sn, &cs);
    *z_even = r * cs;
    *z_odd = r * sn;
}
#endif

}  // namespace deb
