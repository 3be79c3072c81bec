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
// Philox4x32-10 with the ten round keys (key + r * Weyl constants) given: they depend on the seed only, so the host puts
// them into the kernel arguments and every XOR takes its key from the constant bank (no per-round key arithmetic).
__device__ __forceinline__ Philox4 philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* rk) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// log(u) for u in [2^-53, 1] and sin / cos of 2*pi*u for u in [0, 1), for the Box-Muller transform.  Written out (instead of
// the CUDA math library's log / sincos) so that every polynomial coefficient is a constant-bank operand of its DFMA: the
// library versions load their ~30 double constants as 64-bit immediates, two UMOV issue slots each, which was 13 % of all
// instructions the Euler-Maruyama kernel issued.  Accuracy (against libm, 3e5 samples): log 3.2e-16 relative, sin / cos
// 7e-16 absolute -- far inside the 1e-12 bar of the SDE path; this is an evaluation detail, not part of the stream definition.
static __constant__ double deb_c_bm[32] = {
    // [0..6]  sin:  (-1)^n / (2n+1)!,  n = 1..7
    -0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19, -0x1.ae64567f544e4p-26,
    0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41,
    // [7..14] cos:  (-1)^n / (2n)!,  n = 1..8
    -0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-16, -0x1.27e4fb7789f5cp-22,
    0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45,
    // [15..24] log(m) = 2 atanh(s):  2 / (2n+1),  n = 1..10
    0x1.5555555555555p-1, 0x1.999999999999ap-2, 0x1.2492492492492p-2, 0x1.c71c71c71c71cp-3, 0x1.745d1745d1746p-3,
    0x1.3b13b13b13b14p-3, 0x1.1111111111111p-3, 0x1.e1e1e1e1e1e1ep-4, 0x1.af286bca1af28p-4, 0x1.8618618618618p-4,
    // [25..28] pi = hi + lo, ln 2 = hi (42 bits) + lo
    0x1.921fb54442d18p+1, 0x1.1a62633145c07p-53, 0x1.62e42fefa3800p-1, 0x1.ef35793c76730p-45,
    // [29] sqrt 2
    0x1.6a09e667f3bcdp+0, 0.0, 0.0};

__device__ __forceinline__ double deb_log_unit(double u) {
    const int hi = __double2hiint(u);
    int e = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u));  // [1, 2)
    const bool up = m > deb_c_bm[29];
    m = up ? m * 0.5 : m;  // [sqrt(1/2), sqrt 2)
    e = up ? e + 1 : e;
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    double p = deb_c_bm[24];
#pragma unroll
    for (int i = 23; i >= 15; i--) p = __fma_rn(p, z, deb_c_bm[i]);
    const double lm = __fma_rn(s * z, p, 2.0 * s);
    const double ed = (double)e;
    return __fma_rn(ed, deb_c_bm[27], __fma_rn(ed, deb_c_bm[28], lm));
}

__device__ __forceinline__ void deb_sincos_2pi(double u, double* sn, double* cs) {
    const double x = 2.0 * u;                    // [0, 2): half-turns
    const int q = __double2int_rn(2.0 * x);      // nearest quarter-turn, 0..4
    const double r = __fma_rn(-0.5, (double)q, x);  // exact, [-1/4, 1/4]
    const double t = __fma_rn(r, deb_c_bm[25], r * deb_c_bm[26]);
    const double t2 = t * t;
    double ps = deb_c_bm[6];
#pragma unroll
    for (int i = 5; i >= 0; i--) ps = __fma_rn(ps, t2, deb_c_bm[i]);
    double pc = deb_c_bm[14];
#pragma unroll
    for (int i = 13; i >= 7; i--) pc = __fma_rn(pc, t2, deb_c_bm[i]);
    const double s = __fma_rn(t * t2, ps, t);
    const double c = __fma_rn(t2, pc, 1.0);
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;   // odd quarter-turns swap the roles
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

// Both Box-Muller normals of pair index `pair` for path `path`; rk = the ten round keys of the seed.
__device__ __forceinline__ void normal_pair(const uint32_t* rk, uint64_t path, uint64_t pair, double* z_even, double* z_odd) {
    const Philox4 o = philox4x32_10_keys((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)path, (uint32_t)(path >> 32), rk);
    const uint64_t a = (((uint64_t)o.w[0] << 32) | o.w[1]) >> 11;
    const uint64_t b = (((uint64_t)o.w[2] << 32) | o.w[3]) >> 11;
    const double u1 = (double)(a + 1) * 0x1p-53;
    const double u2 = (double)b * 0x1p-53;
    const double r = sqrt(-2.0 * deb_log_unit(u1));
    double sn, cs;
    deb_sincos_2pi(u2, &sn, &cs);
    *z_even = r * cs;
    *z_odd = r * sn;
}

#endif

}  // namespace deb
