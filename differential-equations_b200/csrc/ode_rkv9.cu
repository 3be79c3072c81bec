// ode_rkv9.cu -- dp_ensemble_kernel instantiations for Verner's 9(8) pairs (adaptive family with a dense-output polynomial).
#include "ode_dispatch.cuh"

ode_launch_fn deb_pick_rkv9(int system, int method) {
    switch (method) {
        case DEB_RKV988E: return deb_dispatch::pick_system<deb::TabRkv988e>(system);
        case DEB_RKV989E: return deb_dispatch::pick_system<deb::TabRkv989e>(system);
    }
    return nullptr;
}
