// erk_tableau.cuh -- Butcher tableaux as compile-time accessors.
//
// The reference copies a ButcherTableau{c,a,b,bh,bi,er} (/root/reference/src/tableau/mod.rs:39-46) into every
// method instance and loops over it at run time, multiplying by the many zero entries.  Here every coefficient
// is a constexpr function of compile-time indices: after full unrolling the compiler sees each coefficient as an
// immediate, zero terms vanish (x + (0*h)*k == x for finite k), and nothing is staged in memory at all.
// Values come from erk_tableau_data.h, generated from the reference's own literals (tools/gen_tableau.py).
#pragma once
#include "erk_tableau_data.h"

namespace deb {

// name(i[,j]): constexpr value, used ONLY for compile-time decisions (is this coefficient zero?).
// name##v(i[,j]): the run-time operand, read from a __constant__ image of the same table.  With compile-time
// indices it becomes a constant-bank operand of the DMUL/DADD itself (c[3][off]); a 64-bit immediate would instead
// cost two UMOV issue slots per use (measured: 10% of all issued instructions in the first version of the kernel).
#define DEB_TAB_FN1(name, len, init, cname)                                                                      \
    __host__ __device__ static constexpr double name(int i) { constexpr double v[len] = init; return v[i]; }     \
    __device__ __forceinline__ static double name##v(int i) { return cmem::cname[i]; }
#define DEB_TAB_FN2(name, len, init, cname)                                                                                \
    __host__ __device__ static constexpr double name(int i, int j) { constexpr double v[len][len] = init; return v[i][j]; } \
    __device__ __forceinline__ static double name##v(int i, int j) { return cmem::cname[i][j]; }

namespace cmem {
static __constant__ double d5_c[7] = DEB_DOPRI5_C;
static __constant__ double d5_a[7][7] = DEB_DOPRI5_A;
static __constant__ double d5_b[7] = DEB_DOPRI5_B;
static __constant__ double d5_er[7] = DEB_DOPRI5_ER;
static __constant__ double d5_bi[7][7] = DEB_DOPRI5_BI;
static __constant__ double d8_c[16] = DEB_DOP853_C;
static __constant__ double d8_a[16][16] = DEB_DOP853_A;
static __constant__ double d8_b[12] = DEB_DOP853_B;
static __constant__ double d8_bh[12] = DEB_DOP853_BH;
static __constant__ double d8_er[12] = DEB_DOP853_ER;
static __constant__ double d8_bi[16][16] = DEB_DOP853_BI;
static __constant__ double rkf45_c[6] = DEB_RKF45_C;
static __constant__ double rkf45_a[6][6] = DEB_RKF45_A;
static __constant__ double rkf45_b[6] = DEB_RKF45_B;
static __constant__ double rkf45_bh[6] = DEB_RKF45_BH;
static __constant__ double ck_c[6] = DEB_CASH_KARP_C;
static __constant__ double ck_a[6][6] = DEB_CASH_KARP_A;
static __constant__ double ck_b[6] = DEB_CASH_KARP_B;
static __constant__ double ck_bh[6] = DEB_CASH_KARP_BH;
#define DEB_CMEM_FIXED(pfx, PFX, n)                           \
    static __constant__ double pfx##_c[n] = DEB_##PFX##_C;    \
    static __constant__ double pfx##_a[n][n] = DEB_##PFX##_A; \
    static __constant__ double pfx##_b[n] = DEB_##PFX##_B;
DEB_CMEM_FIXED(euler, EULER, 1)
DEB_CMEM_FIXED(midpoint, MIDPOINT, 2)
DEB_CMEM_FIXED(heun, HEUN, 2)
DEB_CMEM_FIXED(ralston, RALSTON, 2)
DEB_CMEM_FIXED(ssp_rk3, SSP_RK3, 3)
DEB_CMEM_FIXED(rk4, RK4, 4)
DEB_CMEM_FIXED(three_eighths, THREE_EIGHTHS, 4)
#define DEB_CMEM_VERNER(pfx, PFX, o, s, i)                    \
    static __constant__ double pfx##_c[i] = DEB_##PFX##_C;    \
    static __constant__ double pfx##_a[i][i] = DEB_##PFX##_A; \
    static __constant__ double pfx##_b[s] = DEB_##PFX##_B;    \
    static __constant__ double pfx##_bh[s] = DEB_##PFX##_BH;  \
    static __constant__ double pfx##_bi[i][o] = DEB_##PFX##_BI;
DEB_CMEM_VERNER(v655, RKV655E, 6, 9, 10)
DEB_CMEM_VERNER(v656, RKV656E, 6, 9, 12)
DEB_CMEM_VERNER(v766, RKV766E, 7, 10, 13)
DEB_CMEM_VERNER(v767, RKV767E, 7, 10, 16)
DEB_CMEM_VERNER(v877, RKV877E, 8, 13, 17)
DEB_CMEM_VERNER(v878, RKV878E, 8, 13, 21)
DEB_CMEM_VERNER(v988, RKV988E, 9, 16, 21)
DEB_CMEM_VERNER(v989, RKV989E, 9, 16, 26)
}  // namespace cmem

// DOPRI5: /root/reference/src/tableau/dorman_prince.rs:37-117; (O,S,I) = (5,7,7) dormandprince/mod.rs:52-58
struct TabDopri5 {
    static constexpr int O = 5, S = 7, I = 7;
    static constexpr bool ADAPTIVE = true, HAS_BH = false, DP = true, FSAL = false, BI_POLY = false;
    DEB_TAB_FN1(c, 7, DEB_DOPRI5_C, d5_c)
    DEB_TAB_FN2(a, 7, DEB_DOPRI5_A, d5_a)
    DEB_TAB_FN1(b, 7, DEB_DOPRI5_B, d5_b)
    DEB_TAB_FN1(er, 7, DEB_DOPRI5_ER, d5_er)
    __host__ __device__ static constexpr double bh(int) { return 0.0; }
    __device__ __forceinline__ static double bhv(int) { return 0.0; }
    // The stepper reads bi[4][j] (dormandprince/ordinary.rs:229-233) but dopri5() filled row 0
    // (dorman_prince.rs:95-101): rows 4.. are zero, so cont[4] == 0.  Kept as in the reference.
    DEB_TAB_FN2(bi, 7, DEB_DOPRI5_BI, d5_bi)
};

// DOP853: /root/reference/src/tableau/dorman_prince.rs:155-381; (O,S,I) = (8,12,16) dormandprince/mod.rs:45-51
struct TabDop853 {
    static constexpr int O = 8, S = 12, I = 16;
    static constexpr bool ADAPTIVE = true, HAS_BH = true, DP = true, FSAL = false, BI_POLY = false;
    DEB_TAB_FN1(c, 16, DEB_DOP853_C, d8_c)
    DEB_TAB_FN2(a, 16, DEB_DOP853_A, d8_a)
    DEB_TAB_FN1(b, 12, DEB_DOP853_B, d8_b)
    DEB_TAB_FN1(bh, 12, DEB_DOP853_BH, d8_bh)
    DEB_TAB_FN1(er, 12, DEB_DOP853_ER, d8_er)
    DEB_TAB_FN2(bi, 16, DEB_DOP853_BI, d8_bi)
};

// Generic adaptive family (/root/reference/src/methods/erk/adaptive/mod.rs:47-60): error = y_high - y_low with the
// embedded weights bh, infinity-norm error, `max_rejects` stiffness rule, cubic-Hermite dense output (bi = None), no FSAL.
#define DEB_ADAPTIVE_TAB(Name, PFX, pfx)                                               \
    struct Name {                                                                      \
        static constexpr int O = 5, S = 6, I = 6;                                      \
        static constexpr bool ADAPTIVE = true, HAS_BH = true, DP = false;              \
        static constexpr bool FSAL = false, BI_POLY = false;                           \
        DEB_TAB_FN1(c, 6, DEB_##PFX##_C, pfx##_c)                                      \
        DEB_TAB_FN2(a, 6, DEB_##PFX##_A, pfx##_a)                                      \
        DEB_TAB_FN1(b, 6, DEB_##PFX##_B, pfx##_b)                                      \
        DEB_TAB_FN1(bh, 6, DEB_##PFX##_BH, pfx##_bh)                                   \
        __host__ __device__ static constexpr double er(int) { return 0.0; }            \
        __device__ __forceinline__ static double erv(int) { return 0.0; }              \
        __host__ __device__ static constexpr double bi(int, int) { return 0.0; }       \
        __device__ __forceinline__ static double biv(int, int) { return 0.0; }         \
    };
DEB_ADAPTIVE_TAB(TabRkf45, RKF45, rkf45)        // runge_kutta.rs:335
DEB_ADAPTIVE_TAB(TabCashKarp, CASH_KARP, ck)    // runge_kutta.rs:420

// Verner pairs (/root/reference/src/tableau/verner.rs; (O,S,I) and fsal: methods/erk/adaptive/mod.rs:59-122): same stepper as
// RKF45, plus I-S extra stages and a dense-output polynomial bi[i][0..O-1] in s (adaptive/ordinary.rs:145-160, :246-277).
#define DEB_VERNER_TAB(Name, PFX, pfx, order, stages, dense, fsal)                                                  \
    struct Name {                                                                                                   \
        static constexpr int O = order, S = stages, I = dense;                                                      \
        static constexpr bool ADAPTIVE = true, HAS_BH = true, DP = false, FSAL = fsal, BI_POLY = true;              \
        DEB_TAB_FN1(c, dense, DEB_##PFX##_C, pfx##_c)                                                               \
        DEB_TAB_FN2(a, dense, DEB_##PFX##_A, pfx##_a)                                                               \
        DEB_TAB_FN1(b, stages, DEB_##PFX##_B, pfx##_b)                                                              \
        DEB_TAB_FN1(bh, stages, DEB_##PFX##_BH, pfx##_bh)                                                           \
        __host__ __device__ static constexpr double er(int) { return 0.0; }                                         \
        __device__ __forceinline__ static double erv(int) { return 0.0; }                                           \
        __host__ __device__ static constexpr double bi(int i, int j) { constexpr double v[dense][order] = DEB_##PFX##_BI; return v[i][j]; } \
        __device__ __forceinline__ static double biv(int i, int j) { return cmem::pfx##_bi[i][j]; }                 \
        /* does dense stage i enter the polynomial at all? */                                                       \
        __host__ __device__ static constexpr bool bi_row(int i) {                                                   \
            for (int j = 0; j < order; j++) if (bi(i, j) != 0.0) return true;                                       \
            return false;                                                                                           \
        }                                                                                                           \
    };
DEB_VERNER_TAB(TabRkv655e, RKV655E, v655, 6, 9, 10, true)
DEB_VERNER_TAB(TabRkv656e, RKV656E, v656, 6, 9, 12, true)
DEB_VERNER_TAB(TabRkv766e, RKV766E, v766, 7, 10, 13, false)
DEB_VERNER_TAB(TabRkv767e, RKV767E, v767, 7, 10, 16, false)
DEB_VERNER_TAB(TabRkv877e, RKV877E, v877, 8, 13, 17, false)
DEB_VERNER_TAB(TabRkv878e, RKV878E, v878, 8, 13, 21, false)
DEB_VERNER_TAB(TabRkv988e, RKV988E, v988, 9, 16, 21, false)
DEB_VERNER_TAB(TabRkv989e, RKV989E, v989, 9, 16, 26, false)

#define DEB_FIXED_TAB(Name, PFX, pfx, order, stages)                 \
    struct Name {                                                    \
        static constexpr int O = order, S = stages, I = stages;      \
        static constexpr bool ADAPTIVE = false, HAS_BH = false;      \
        static constexpr bool DP = false, FSAL = false, BI_POLY = false; \
        DEB_TAB_FN1(c, stages, DEB_##PFX##_C, pfx##_c)               \
        DEB_TAB_FN2(a, stages, DEB_##PFX##_A, pfx##_a)               \
        DEB_TAB_FN1(b, stages, DEB_##PFX##_B, pfx##_b)               \
    };

// fixed-step constructors, /root/reference/src/methods/erk/fixed/mod.rs:41-89 (order, stages); fsal = false for all
DEB_FIXED_TAB(TabEuler, EULER, euler, 1, 1)
DEB_FIXED_TAB(TabMidpoint, MIDPOINT, midpoint, 2, 2)
DEB_FIXED_TAB(TabHeun, HEUN, heun, 2, 2)
DEB_FIXED_TAB(TabRalston, RALSTON, ralston, 2, 2)
DEB_FIXED_TAB(TabSspRk3, SSP_RK3, ssp_rk3, 3, 3)
DEB_FIXED_TAB(TabRk4, RK4, rk4, 4, 4)
DEB_FIXED_TAB(TabThreeEighths, THREE_EIGHTHS, three_eighths, 4, 4)

}  // namespace deb
