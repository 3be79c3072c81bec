// ode_dopri.cu -- dp_ensemble_kernel instantiations for the Dormand-Prince family (DOPRI5, DOP853) over the built-in systems.
#include "ode_dispatch.cuh"

using namespace deb_dispatch;

template <class Sys>
static ode_launch_fn pick_method(int method) {
    switch (method) {
#ifdef DEB_VAR_BLOCK  // EXPERIMENT: launch shape of the DOPRI5 kernels
        case DEB_DOPRI5: return launch_dp<Sys, deb::TabDopri5, DEB_VAR_BLOCK, DEB_VAR_MINB>;
#else
        case DEB_DOPRI5: return launch_dp<Sys, deb::TabDopri5, 128, 5>;  // 96 regs, no spills, 20 warps/SM (sweep: profiles/)
#endif
        // DOP853: 12 stage vectors; 4 CTAs/SM (<= 128 regs) for dim <= 2, 3 CTAs/SM (<= 168 regs) for dim 3 (sweep: DESIGN.md 7)
        case DEB_DOP853: return (Sys::DIM <= 2) ? launch_dp<Sys, deb::TabDop853, 128, 4> : launch_dp<Sys, deb::TabDop853, 128, 3>;
    }
    return nullptr;
}

ode_launch_fn deb_pick_dopri(int system, int method) {
    switch (system) {
        case DEB_SYS_EXPONENTIAL: return pick_method<deb::SysExponential>(method);
        case DEB_SYS_LINEAR: return pick_method<deb::SysLinear>(method);
        case DEB_SYS_HARMONIC: return pick_method<deb::SysHarmonic>(method);
        case DEB_SYS_LOGISTIC: return pick_method<deb::SysLogistic>(method);
        case DEB_SYS_VAN_DER_POL: return pick_method<deb::SysVanDerPol>(method);
        case DEB_SYS_LORENZ: return pick_method<deb::SysLorenz>(method);
        case DEB_SYS_BRUSSELATOR: return pick_method<deb::SysBrusselator>(method);
        case DEB_SYS_ROBERTSON: return pick_method<deb::SysRobertson>(method);
    }
    return nullptr;
}
