// ode_dispatch.cuh -- launch wrappers of the adaptive ensemble kernel, shared by the per-family translation units.
//
// The kernel is a template over (system, tableau): every pair is fully unrolled device code, so the instantiations are
// spread over several .cu files (ode_dopri.cu, ode_rkf.cu, ode_rkv6.cu ... ode_rkv9.cu) that compile in parallel; each
// exports one C++ lookup function `deb_pick_<family>(system, method)` that deb_api.cu consults.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/deb_ensemble.h"
#include "erk_ensemble.cuh"
#include "systems.cuh"

// defined in deb_api.cu: records the text for deb_last_error() and returns `code`
int deb_fail(int code, const char* msg);
// defined in deb_api.cu: the library counts its kernel launches (deb_launch_count)
void deb_count_launch(int n);

typedef int (*ode_launch_fn)(const deb::OdeKernelArgs&, int sms, cudaStream_t);

// lookups implemented by the family translation units; nullptr = not mine
ode_launch_fn deb_pick_dopri(int system, int method);
ode_launch_fn deb_pick_rkf(int system, int method);
ode_launch_fn deb_pick_rkv6(int system, int method);
ode_launch_fn deb_pick_rkv7(int system, int method);
ode_launch_fn deb_pick_rkv8(int system, int method);
ode_launch_fn deb_pick_rkv9(int system, int method);

namespace deb_dispatch {

inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return deb_fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? DEB_ERR_NO_DEVICE : DEB_ERR_CUDA, buf);
}
#define DEB_DISPATCH_CUDA(call)                                                                \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return deb_dispatch::cuda_fail(e_, #call, __FILE__, __LINE__);  \
    } while (0)

template <class Sys, class Tab, int BLOCK, int MIN_BLOCKS, bool SHARED_P>
int launch_dp_impl(const deb::OdeKernelArgs& a, int sms, cudaStream_t st) {
    auto kern = deb::dp_ensemble_kernel<Sys, Tab, BLOCK, MIN_BLOCKS, SHARED_P>;
    // methods with extra dense stages park whole steps in dynamic shared memory (erk_ensemble.cuh: flush_dense_parked)
    // (only launches that can emit rows touch it: a final-state-only launch keeps the shared memory for the L1 cache)
    constexpr unsigned dyn_max = deb::dp_dynamic_smem_bytes<Sys, Tab, BLOCK, false>();
    const unsigned dyn = (a.n_rows > 0 && a.y_eval != nullptr) ? dyn_max : 0u;
    if (dyn > 0) DEB_DISPATCH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int per_sm = 0;
    DEB_DISPATCH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, dyn));
    if (per_sm < 1) per_sm = 1;
    // persistent grid: every resident CTA slot of every SM, but never more threads than trajectories
    long long blocks = (long long)sms * per_sm;
    const long long need = (a.n_traj + BLOCK - 1) / BLOCK;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (getenv("DEB_DEBUG_LAUNCH")) fprintf(stderr, "[deb] built-in kernel: grid %lld x %d\n", blocks, BLOCK);
    kern<<<(unsigned)blocks, BLOCK, dyn, st>>>(a);
    DEB_DISPATCH_CUDA(cudaGetLastError());
    deb_count_launch(1);
    return DEB_OK;
}

template <class Sys, class Tab, int BLOCK, int MIN_BLOCKS>
int launch_dp(const deb::OdeKernelArgs& a, int sms, cudaStream_t st) {
    // a.pc is filled by the caller iff the parameter set is shared (params_stride == 0) and lives on the host
    if (a.params_stride == 0 && a.params == nullptr) return launch_dp_impl<Sys, Tab, BLOCK, MIN_BLOCKS, true>(a, sms, st);
    return launch_dp_impl<Sys, Tab, BLOCK, MIN_BLOCKS, false>(a, sms, st);
}

// CTAs per SM for a tableau with S stage vectors of a DIM-dimensional system held in registers (128-thread CTAs):
// 5 (<= 96 regs), 4 (<= 128), 3 (<= 168), 2 (<= 255)
constexpr int min_blocks_for(int stages, int dim) {
    const int need = 80 + (stages + 3) * dim * 2;  // ~80 for the controller and loop state + k[S], y, y_new, a stage argument
    return need <= 96 ? 5 : need <= 128 ? 4 : need <= 168 ? 3 : 2;
}

// one tableau over the eight built-in systems
template <class Tab>
ode_launch_fn pick_system(int system) {
#define DEB_SYS_CASE(ID, T) \
    case ID: return launch_dp<deb::T, Tab, 128, min_blocks_for(Tab::S, deb::T::DIM)>;
    switch (system) {
        DEB_SYS_CASE(DEB_SYS_EXPONENTIAL, SysExponential)
        DEB_SYS_CASE(DEB_SYS_LINEAR, SysLinear)
        DEB_SYS_CASE(DEB_SYS_HARMONIC, SysHarmonic)
        DEB_SYS_CASE(DEB_SYS_LOGISTIC, SysLogistic)
        DEB_SYS_CASE(DEB_SYS_VAN_DER_POL, SysVanDerPol)
        DEB_SYS_CASE(DEB_SYS_LORENZ, SysLorenz)
        DEB_SYS_CASE(DEB_SYS_BRUSSELATOR, SysBrusselator)
        DEB_SYS_CASE(DEB_SYS_ROBERTSON, SysRobertson)
    }
#undef DEB_SYS_CASE
    return nullptr;
}

}  // namespace deb_dispatch
