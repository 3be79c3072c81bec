// Bit-exact port of glibc's x86-64 (FMA ifunc variant) double-precision pow() for positive bases.
//
// Why this exists: the reference's step-size controller is `scale = safety * err.powf(-1/order)`
// (/root/reference/src/methods/erk/dormandprince/ordinary.rs:151-154) and its initial-step picker uses
// `(0.01/der12).powf(1/order)` (/root/reference/src/methods/h_init.rs:124).  `f64::powf` is libm `pow`,
// which is < 1 ulp but not correctly rounded, and a 1-ulp difference in h changes a chaotic trajectory
// completely.  The ensemble kernels therefore evaluate pow with exactly the operation sequence glibc
// 2.39 executes on every FMA-capable x86-64 host (the Arm "optimized routines" log/exp split with
// 128-entry tables; GCC contracted a*b+c into FMA inside libm, so the sequence below follows the
// machine code of that variant operation by operation: every DEB_FMA is one fused operation, every
// DEB_MUL/DEB_ADD/DEB_SUB one separately rounded operation).
//
// Only what the path needs is ported: x is a non-negative double, NaN or +inf (an error norm or a ratio of
// norms), y is a finite exponent with 2^-65 <= |y| < 2^63 (here +-1/order).  Negative x returns NaN
// like libm for a non-integer exponent.  The main path covers every positive finite x including subnormals.
//
// The same header compiles for the host (tests/ check it against libm on 1e8+ samples) and for the
// device (where it is the product path).
#pragma once
#include <stdint.h>
#include "glibc_pow_tables.h"

#if defined(__CUDA_ARCH__)
#define DEB_HD __device__ __forceinline__
#define DEB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define DEB_MUL(a, b) __dmul_rn((a), (b))
#define DEB_ADD(a, b) __dadd_rn((a), (b))
#define DEB_SUB(a, b) __dsub_rn((a), (b))
#define DEB_AS_U64(x) ((uint64_t)__double_as_longlong(x))
#define DEB_AS_F64(u) __longlong_as_double((long long)(u))
#else
#include <string.h>
#define DEB_HD static inline
// host build must use -ffp-contract=off so that only DEB_FMA fuses
#define DEB_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define DEB_MUL(a, b) ((a) * (b))
#define DEB_ADD(a, b) ((a) + (b))
#define DEB_SUB(a, b) ((a) - (b))
static inline uint64_t deb_as_u64_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double deb_as_f64_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define DEB_AS_U64(x) deb_as_u64_(x)
#define DEB_AS_F64(u) deb_as_f64_(u)
#endif

// Table storage.  Host: plain static arrays.  Device: the kernels copy these __constant__ images into
// shared memory once per CTA (lanes index them divergently, which constant memory serialises) and pass
// the shared pointers in.
#if defined(__CUDACC__)
// scalar coefficients: constant-bank operands on the device (a 64-bit immediate costs two UMOV issue slots per use)
static __constant__ double deb_c_powk[17] = {DEB_POWLOG_LN2HI, DEB_POWLOG_LN2LO, DEB_POWLOG_A0, DEB_POWLOG_A1, DEB_POWLOG_A2, DEB_POWLOG_A3, DEB_POWLOG_A4, DEB_POWLOG_A5, DEB_POWLOG_A6, DEB_EXP_INVLN2N, DEB_EXP_SHIFT, DEB_EXP_NEGLN2HIN, DEB_EXP_NEGLN2LON, DEB_EXP_C2, DEB_EXP_C3, DEB_EXP_C4, DEB_EXP_C5};
static __constant__ double deb_c_powlog_tab[128 * 3] = DEB_POWLOG_TAB_INIT;
static __constant__ unsigned long long deb_c_exp_tab[128 * 2] = DEB_EXP_TAB_INIT;
#endif
#if !defined(__CUDA_ARCH__)
static const double deb_h_powlog_tab[128 * 3] = DEB_POWLOG_TAB_INIT;
static const unsigned long long deb_h_exp_tab[128 * 2] = DEB_EXP_TAB_INIT;
#endif

struct deb_pow_tables {
    const double* powlog;               // [128][3] = invc, logc, logctail
    const unsigned long long* exptab;   // [128][2] = tail bits, scale bits
};

#if !defined(__CUDA_ARCH__)
static inline deb_pow_tables deb_host_pow_tables() {
    deb_pow_tables t; t.powlog = deb_h_powlog_tab; t.exptab = deb_h_exp_tab; return t;
}
#endif

// pow(x, y) for x >= 0 (or NaN/inf), y "ordinary" (see header comment).
#if defined(__CUDA_ARCH__)
#define DEBK_POWLOG_LN2HI deb_c_powk[0]
#define DEBK_POWLOG_LN2LO deb_c_powk[1]
#define DEBK_POWLOG_A0 deb_c_powk[2]
#define DEBK_POWLOG_A1 deb_c_powk[3]
#define DEBK_POWLOG_A2 deb_c_powk[4]
#define DEBK_POWLOG_A3 deb_c_powk[5]
#define DEBK_POWLOG_A4 deb_c_powk[6]
#define DEBK_POWLOG_A5 deb_c_powk[7]
#define DEBK_POWLOG_A6 deb_c_powk[8]
#define DEBK_EXP_INVLN2N deb_c_powk[9]
#define DEBK_EXP_SHIFT deb_c_powk[10]
#define DEBK_EXP_NEGLN2HIN deb_c_powk[11]
#define DEBK_EXP_NEGLN2LON deb_c_powk[12]
#define DEBK_EXP_C2 deb_c_powk[13]
#define DEBK_EXP_C3 deb_c_powk[14]
#define DEBK_EXP_C4 deb_c_powk[15]
#define DEBK_EXP_C5 deb_c_powk[16]
#else
#define DEBK_POWLOG_LN2HI DEB_POWLOG_LN2HI
#define DEBK_POWLOG_LN2LO DEB_POWLOG_LN2LO
#define DEBK_POWLOG_A0 DEB_POWLOG_A0
#define DEBK_POWLOG_A1 DEB_POWLOG_A1
#define DEBK_POWLOG_A2 DEB_POWLOG_A2
#define DEBK_POWLOG_A3 DEB_POWLOG_A3
#define DEBK_POWLOG_A4 DEB_POWLOG_A4
#define DEBK_POWLOG_A5 DEB_POWLOG_A5
#define DEBK_POWLOG_A6 DEB_POWLOG_A6
#define DEBK_EXP_INVLN2N DEB_EXP_INVLN2N
#define DEBK_EXP_SHIFT DEB_EXP_SHIFT
#define DEBK_EXP_NEGLN2HIN DEB_EXP_NEGLN2HIN
#define DEBK_EXP_NEGLN2LON DEB_EXP_NEGLN2LON
#define DEBK_EXP_C2 DEB_EXP_C2
#define DEBK_EXP_C3 DEB_EXP_C3
#define DEBK_EXP_C4 DEB_EXP_C4
#define DEBK_EXP_C5 DEB_EXP_C5
#endif

DEB_HD double deb_pow_pos(double x, double y, const deb_pow_tables tb) {
    uint64_t ix = DEB_AS_U64(x);
    uint32_t topx = (uint32_t)(ix >> 52);
    if (topx - 1u >= 0x7feu) {  // zero, subnormal, inf, nan or negative
        if (x != x) return x + y;                                // NaN in, NaN out
        if ((ix << 1) == 0) return (y < 0.0) ? (1.0 / 0.0) : 0.0; // +-0 (sign only matters for odd integer y)
        if (ix >> 63) return (x - x) / (x - x);                   // negative base, non-integer exponent
        if (ix == 0x7ff0000000000000ULL) return (y < 0.0) ? 0.0 : x;  // +inf
        // subnormal: scale by 2^52 and correct the exponent field
        ix = DEB_AS_U64(DEB_MUL(x, 0x1p52));
        ix &= 0x7fffffffffffffffULL;
        ix -= 52ULL << 52;
    }
    // ---- log_inline: hi + lo ~= log(x), relative error ~2^-68
    const uint64_t OFF = 0x3fe6955500000000ULL;
    uint64_t tmp = ix - OFF;
    int i = (int)((tmp >> 45) & 127);
    int k = (int)((int64_t)tmp >> 52);
    uint64_t iz = ix - (tmp & (0xfffULL << 52));
    double z = DEB_AS_F64(iz);
    double kd = (double)k;
    double invc = tb.powlog[3 * i + 0];
    double logc = tb.powlog[3 * i + 1];
    double logctail = tb.powlog[3 * i + 2];

    double t1 = DEB_FMA(kd, DEBK_POWLOG_LN2HI, logc);
    double lo1 = DEB_FMA(kd, DEBK_POWLOG_LN2LO, logctail);
    double r = DEB_FMA(z, invc, -1.0);
    double ar = DEB_MUL(r, DEBK_POWLOG_A0);
    double p12 = DEB_FMA(r, DEBK_POWLOG_A2, DEBK_POWLOG_A1);
    double p34 = DEB_FMA(r, DEBK_POWLOG_A4, DEBK_POWLOG_A3);
    double t2 = DEB_ADD(r, t1);
    double lo2 = DEB_ADD(DEB_SUB(t1, t2), r);
    double ar2 = DEB_MUL(r, ar);
    double ar3 = DEB_MUL(r, ar2);
    double lo3 = DEB_FMA(ar, r, -ar2);
    double hi = DEB_ADD(t2, ar2);
    double p56 = DEB_FMA(r, DEBK_POWLOG_A6, DEBK_POWLOG_A5);
    double lo4 = DEB_ADD(DEB_SUB(t2, hi), ar2);
    double q = DEB_FMA(p56, ar2, p34);
    double pp = DEB_FMA(ar2, q, p12);
    double lsum = DEB_ADD(DEB_ADD(DEB_ADD(lo1, lo2), lo3), lo4);
    double lo = DEB_FMA(ar3, pp, lsum);
    double lhi = DEB_ADD(hi, lo);
    double llo = DEB_ADD(DEB_SUB(hi, lhi), lo);

    // ---- y * log(x) in double-double
    double ehi = DEB_MUL(y, lhi);
    double elo = DEB_FMA(y, llo, DEB_FMA(lhi, y, -ehi));

    // ---- exp_inline(ehi, elo)
    uint32_t abstop = (uint32_t)(DEB_AS_U64(ehi) >> 52) & 0x7ffu;
    bool big = false;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop - 0x3c9u >= 0x80000000u) return DEB_ADD(1.0, ehi);  // |y log x| < 2^-54
        if (abstop >= 0x409u) {                                          // |y log x| >= 1024
            return (DEB_AS_U64(ehi) >> 63) ? 0.0 : (1.0 / 0.0);
        }
        big = true;  // 512 <= |y log x| < 1024: result may over/underflow, handled below
    }
    double zs = DEB_FMA(ehi, DEBK_EXP_INVLN2N, DEBK_EXP_SHIFT);
    uint64_t ki = DEB_AS_U64(zs);
    double kd2 = DEB_SUB(zs, DEBK_EXP_SHIFT);
    double r0 = DEB_FMA(kd2, DEBK_EXP_NEGLN2HIN, ehi);
    double r1 = DEB_FMA(kd2, DEBK_EXP_NEGLN2LON, r0);
    double rr = DEB_ADD(elo, r1);
    uint32_t idx = 2u * (uint32_t)(ki & 127);
    uint64_t top = ki << 45;
    double tail = DEB_AS_F64(tb.exptab[idx]);
    uint64_t sbits = tb.exptab[idx + 1] + top;
    double c23 = DEB_FMA(rr, DEBK_EXP_C3, DEBK_EXP_C2);
    double tr = DEB_ADD(rr, tail);
    double r2 = DEB_MUL(rr, rr);
    double c45 = DEB_FMA(rr, DEBK_EXP_C5, DEBK_EXP_C4);
    double s1 = DEB_FMA(c23, r2, tr);
    double r4 = DEB_MUL(r2, r2);
    double tm = DEB_FMA(c45, r4, s1);
    if (big) {
        // glibc specialcase(): scale is split so the intermediate neither overflows nor underflows.
        if ((ki & 0x80000000ULL) == 0) {  // k > 0: result may overflow
            sbits -= 1009ULL << 52;
            double sc = DEB_AS_F64(sbits);
            return DEB_MUL(0x1p1009, DEB_FMA(sc, tm, sc));
        }
        sbits += 1022ULL << 52;  // k < 0: result may be subnormal
        double sc = DEB_AS_F64(sbits);
        double yv = DEB_ADD(sc, DEB_MUL(sc, tm));
        if (yv < 1.0 && yv > -1.0) {
            double lo5 = DEB_ADD(DEB_SUB(sc, yv), DEB_MUL(sc, tm));
            double hi5 = DEB_ADD(1.0, yv);
            lo5 = DEB_ADD(DEB_ADD(DEB_SUB(1.0, hi5), yv), lo5);
            yv = DEB_SUB(DEB_ADD(hi5, lo5), 1.0);
            if (yv == 0.0) yv = 0.0;
        }
        return DEB_MUL(0x1p-1022, yv);
    }
    double sc = DEB_AS_F64(sbits);
    return DEB_FMA(tm, sc, sc);
}
