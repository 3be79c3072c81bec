// ode_rkv8.cu -- dp_ensemble_kernel instantiations for Verner's 8(7) pairs (adaptive family with a dense-output polynomial).
#include "ode_dispatch.cuh"

ode_launch_fn deb_pick_rkv8(int system, int method) {
    switch (method) {
        case DEB_RKV877E: return deb_dispatch::pick_system<deb::TabRkv877e>(system);
        case DEB_RKV878E: return deb_dispatch::pick_system<deb::TabRkv878e>(system);
    }
    return nullptr;
}
