// ode_rkv7.cu -- dp_ensemble_kernel instantiations for Verner's 7(6) pairs (adaptive family with a dense-output polynomial).
#include "ode_dispatch.cuh"

ode_launch_fn deb_pick_rkv7(int system, int method) {
    switch (method) {
        case DEB_RKV766E: return deb_dispatch::pick_system<deb::TabRkv766e>(system);
        case DEB_RKV767E: return deb_dispatch::pick_system<deb::TabRkv767e>(system);
    }
    return nullptr;
}
