// philox_ref.h -- host regeneration of the ensemble's Wiener increments.  TEST INFRASTRUCTURE (oracle).
//
// The reference leaves the noise to the user (`SDE::noise(&mut self, dt, dw)`, src/sde/sde.rs:67; its examples
// draw Normal(0, sqrt(dt)) from rand's StdRng, examples/sde/03_ornstein_uhlenbeck/main.rs:51-54).  A sequential
// user RNG cannot serve 1e8 paths, so the ensemble front end fixes the stream once, counter-based:
//
//   normal number q = step*dim + comp of path p comes from Philox4x32-10 (Salmon et al., SC'11) with
//       key     = (seed & 0xffffffff, seed >> 32)
//       counter = ((q>>1) & 0xffffffff, (q>>1) >> 32, p & 0xffffffff, p >> 32)
//   giving 4 words w0..w3;  a = (w0<<32 | w1) >> 11, b = (w2<<32 | w3) >> 11  (53 bits each)
//       u1 = (a + 1) * 2^-53  in (0,1],   u2 = b * 2^-53  in [0,1)
//       r  = sqrt(-2 * log(u1)),  theta = (2*pi) * u2
//       z  = r * cos(theta) if q is even, r * sin(theta) if q is odd        (Box-Muller)
//   dW = sqrt(h) * z
//
// The device implementation is differential-equations_b200/csrc/philox.h; this is the independent host copy.
#pragma once
#include <cmath>
#include <cstdint>

namespace deb_ref {

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline double standard_normal(uint64_t seed, uint64_t path, uint64_t q) {
    uint64_t pair = q >> 1;
    uint32_t ctr[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)path, (uint32_t)(path >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    philox4x32_10(ctr, key, w);
    uint64_t a = (((uint64_t)w[0] << 32) | w[1]) >> 11;
    uint64_t b = (((uint64_t)w[2] << 32) | w[3]) >> 11;
    double u1 = (double)(a + 1) * 0x1p-53;
    double u2 = (double)b * 0x1p-53;
    double r = std::sqrt(-2.0 * std::log(u1));
    double theta = 6.283185307179586 * u2;
    return (q & 1) ? r * std::sin(theta) : r * std::cos(theta);
}

inline double wiener_increment(uint64_t seed, uint64_t path, uint64_t step, int comp, int dim, double h) {
    return std::sqrt(h) * standard_normal(seed, path, step * (uint64_t)dim + (uint64_t)comp);
}

}  // namespace deb_ref
