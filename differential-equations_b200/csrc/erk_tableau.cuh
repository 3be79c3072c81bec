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

#define DEB_TAB_FN1(name, len, init) \
    __host__ __device__ static constexpr double name(int i) { constexpr double v[len] = init; return v[i]; }
#define DEB_TAB_FN2(name, len, init) \
    __host__ __device__ static constexpr double name(int i, int j) { constexpr double v[len][len] = init; return v[i][j]; }

// DOPRI5: /root/reference/src/tableau/dorman_prince.rs:37-117; (O,S,I) = (5,7,7) dormandprince/mod.rs:52-58
struct TabDopri5 {
    static constexpr int O = 5, S = 7, I = 7;
    static constexpr bool ADAPTIVE = true, HAS_BH = false;
    DEB_TAB_FN1(c, 7, DEB_DOPRI5_C)
    DEB_TAB_FN2(a, 7, DEB_DOPRI5_A)
    DEB_TAB_FN1(b, 7, DEB_DOPRI5_B)
    DEB_TAB_FN1(er, 7, DEB_DOPRI5_ER)
    __host__ __device__ static constexpr double bh(int) { return 0.0; }
    // The stepper reads bi[4][j] (dormandprince/ordinary.rs:229-233) but dopri5() filled row 0
    // (dorman_prince.rs:95-101): rows 4.. are zero, so cont[4] == 0.  Kept as in the reference.
    DEB_TAB_FN2(bi, 7, DEB_DOPRI5_BI)
};

// DOP853: /root/reference/src/tableau/dorman_prince.rs:155-381; (O,S,I) = (8,12,16) dormandprince/mod.rs:45-51
struct TabDop853 {
    static constexpr int O = 8, S = 12, I = 16;
    static constexpr bool ADAPTIVE = true, HAS_BH = true;
    DEB_TAB_FN1(c, 16, DEB_DOP853_C)
    DEB_TAB_FN2(a, 16, DEB_DOP853_A)
    DEB_TAB_FN1(b, 12, DEB_DOP853_B)
    DEB_TAB_FN1(bh, 12, DEB_DOP853_BH)
    DEB_TAB_FN1(er, 12, DEB_DOP853_ER)
    DEB_TAB_FN2(bi, 16, DEB_DOP853_BI)
};

#define DEB_FIXED_TAB(Name, PFX, order, stages)                      \
    struct Name {                                                    \
        static constexpr int O = order, S = stages, I = stages;      \
        static constexpr bool ADAPTIVE = false, HAS_BH = false;      \
        DEB_TAB_FN1(c, stages, DEB_##PFX##_C)                        \
        DEB_TAB_FN2(a, stages, DEB_##PFX##_A)                        \
        DEB_TAB_FN1(b, stages, DEB_##PFX##_B)                        \
    };

// fixed-step constructors, /root/reference/src/methods/erk/fixed/mod.rs:41-89 (order, stages); fsal = false for all
DEB_FIXED_TAB(TabEuler, EULER, 1, 1)
DEB_FIXED_TAB(TabMidpoint, MIDPOINT, 2, 2)
DEB_FIXED_TAB(TabHeun, HEUN, 2, 2)
DEB_FIXED_TAB(TabRalston, RALSTON, 2, 2)
DEB_FIXED_TAB(TabSspRk3, SSP_RK3, 3, 3)
DEB_FIXED_TAB(TabRk4, RK4, 4, 4)
DEB_FIXED_TAB(TabThreeEighths, THREE_EIGHTHS, 4, 4)

}  // namespace deb
