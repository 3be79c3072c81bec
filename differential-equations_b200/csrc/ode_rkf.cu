// ode_rkf.cu -- dp_ensemble_kernel instantiations for RKF45 and Cash-Karp (adaptive family, cubic-Hermite dense output).
#include "ode_dispatch.cuh"

using namespace deb_dispatch;

template <class Tab>
static ode_launch_fn pick4(int system) {
    switch (system) {
        case DEB_SYS_EXPONENTIAL: return launch_dp<deb::SysExponential, Tab, 128, 4>;
        case DEB_SYS_LINEAR: return launch_dp<deb::SysLinear, Tab, 128, 4>;
        case DEB_SYS_HARMONIC: return launch_dp<deb::SysHarmonic, Tab, 128, 4>;
        case DEB_SYS_LOGISTIC: return launch_dp<deb::SysLogistic, Tab, 128, 4>;
        case DEB_SYS_VAN_DER_POL: return launch_dp<deb::SysVanDerPol, Tab, 128, 4>;
        case DEB_SYS_LORENZ: return launch_dp<deb::SysLorenz, Tab, 128, 4>;
        case DEB_SYS_BRUSSELATOR: return launch_dp<deb::SysBrusselator, Tab, 128, 4>;
        case DEB_SYS_ROBERTSON: return launch_dp<deb::SysRobertson, Tab, 128, 4>;
    }
    return nullptr;
}

ode_launch_fn deb_pick_rkf(int system, int method) {
    switch (method) {
        case DEB_RKF45: return pick4<deb::TabRkf45>(system);
        case DEB_CASH_KARP: return pick4<deb::TabCashKarp>(system);
    }
    return nullptr;
}
