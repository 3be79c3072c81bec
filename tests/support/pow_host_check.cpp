// Host build of the product's glibc pow port (differential-equations_b200/csrc/glibc_pow.h) compared with libm pow.
// Built and run by tests/test_pow_port_cpu.py: g++ -O2 -mfma -ffp-contract=off.  Prints "<comparisons> <mismatches>".
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "glibc_pow.h"
static inline uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000000;
    const double ys[] = {-0.2, -0.125, 0.2, 0.125, -1.0 / 3, -0.25, -1.0 / 7, 1.0 / 6, 0.5, -0.5, -1.0, 1.0, 1.0 / 9};
    const int ny = sizeof(ys) / sizeof(ys[0]);
    int nt = (int)std::thread::hardware_concurrency(); if (nt < 1) nt = 1;
    std::atomic<long> bad{0}, total{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] {
        deb_pow_tables tb = deb_host_pow_tables();
        uint64_t s = 1234567 + t * 7919; long lb = 0, lt = 0;
        for (long i = 0; i < n; i++) {
            double xs[4];
            uint64_t u = sm64(s);
            uint64_t bits = u & 0x7fffffffffffffffULL; if ((bits >> 52) == 0x7ff) bits &= ~(1ULL << 62);
            memcpy(&xs[0], &bits, 8);                                           // any positive finite double
            xs[1] = std::exp((((sm64(s) >> 11) * 0x1p-53) * 2 - 1) * 27.6);     // [1e-12, 1e12], the controller's range
            xs[2] = 1.0 + (((sm64(s) >> 11) * 0x1p-53) - 0.5) * std::ldexp(1.0, -(int)(sm64(s) % 60));  // near 1
            bits = sm64(s) & 0x000fffffffffffffULL; memcpy(&xs[3], &bits, 8);  // subnormal
            for (double x : xs) for (int j = 0; j < ny; j++) {
                double a = std::pow(x, ys[j]), b = deb_pow_pos(x, ys[j], tb);
                uint64_t ua, ub; memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
                lt++;
                if (ua != ub && !(a != a && b != b)) { lb++; if (lb < 3) fprintf(stderr, "MISMATCH x=%a y=%a libm=%a port=%a\n", x, ys[j], a, b); }
            }
        }
        bad += lb; total += lt;
    });
    for (auto& x : th) x.join();
    deb_pow_tables tb = deb_host_pow_tables();
    const double sp[] = {0.0, INFINITY, NAN, -1.0, -0.0, 1.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308};
    long sb = 0;
    for (double x : sp) for (double y : {-0.2, 0.2, -0.125, 0.125}) {
        double a = std::pow(x, y), b = deb_pow_pos(x, y, tb); uint64_t ua, ub; memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
        if (ua != ub && !(a != a && b != b)) { sb++; fprintf(stderr, "SPECIAL MISMATCH x=%a y=%a %a %a\n", x, y, a, b); }
    }
    printf("%ld %ld\n", (long)total, (long)bad + sb);
    return 0;
}
