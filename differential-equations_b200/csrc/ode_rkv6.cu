// ode_rkv6.cu -- dp_ensemble_kernel instantiations for Verner's 6(5) pairs (adaptive family with a dense-output polynomial).
#include "ode_dispatch.cuh"

ode_launch_fn deb_pick_rkv6(int system, int method) {
    switch (method) {
        case DEB_RKV655E: return deb_dispatch::pick_system<deb::TabRkv655e>(system);
        case DEB_RKV656E: return deb_dispatch::pick_system<deb::TabRkv656e>(system);
    }
    return nullptr;
}
