// deb_api.cu -- implementation of the C ABI declared in include/deb_ensemble.h (libdeb200.so).
//
// Host side of the drop-in boundary: validates the problem like the reference's builder does, sorts/filters
// t_eval like TEvalSolout::new (/root/reference/src/solout/t_eval.rs:154-171), stages buffers, picks the kernel
// instantiation for (system, method) and launches it.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <chrono>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/deb_ensemble.h"
#include "ens_stats.cuh"
#include "erk_ensemble.cuh"
#include "erk_fixed.cuh"
#include "ode_dispatch.cuh"
#include "mol_heat.cuh"
#include "sde_ensemble.cuh"
#include "systems.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

}  // namespace

int deb_fail(int code, const char* msg) { return fail(code, msg); }

namespace {

#define DEB_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char buf_[512];                                                                             \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? DEB_ERR_NO_DEVICE : DEB_ERR_CUDA, buf_); \
        }                                                                                               \
    } while (0)

thread_local int g_device = 0;  // device selected by the current call

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(DEB_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                           "); the ensemble integrator has no CPU fallback");
    if (device < 0 || device >= n) return fail(DEB_ERR_BAD_ARG, "device ordinal out of range");
    DEB_CUDA(cudaSetDevice(device));
    g_device = device;
    return DEB_OK;
}

// The library's own stream-ordered memory pool per device: staging buffers of HOST-memspace calls and work buffers are
// cached across calls (release threshold = never) instead of paying cudaMalloc/cudaFree (tens to hundreds of ms for
// multi-GB buffers, with a device-wide synchronisation) in every call.  deb_trim_memory() gives the memory back.
struct DeviceInfo { int sms = 0; cudaMemPool_t pool = nullptr; };
int device_info(int device, DeviceInfo* di) {
    static std::mutex mu;
    static std::vector<DeviceInfo> cache;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)cache.size() <= device) cache.resize(device + 1);
    if (cache[device].sms == 0) {
        int sms = 0;
        DEB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        DEB_CUDA(cudaMemPoolCreate(&pool, &props));
        unsigned long long keep = ~0ull;
        DEB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        cache[device].pool = pool;
        cache[device].sms = sms;
    }
    *di = cache[device];
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ kernel table
// (ode_launch_fn and the adaptive-family lookups: ode_dispatch.cuh)

template <class Sys, class Tab>
int launch_fixed(const deb::OdeKernelArgs& a, int sms, cudaStream_t st) {
    constexpr int BLOCK = 128;
    auto kern = deb::fixed_ensemble_kernel<Sys, Tab, BLOCK>;
    long long blocks = (a.n_traj + BLOCK - 1) / BLOCK;
    const long long cap = (long long)sms * 16 * 8;  // grid-stride beyond a few waves
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, BLOCK, 0, st>>>(a);
    DEB_CUDA(cudaGetLastError());
    return DEB_OK;
}

template <class Sys>
ode_launch_fn pick_fixed_method(int method) {
    switch (method) {
        case DEB_EULER: return launch_fixed<Sys, deb::TabEuler>;
        case DEB_MIDPOINT: return launch_fixed<Sys, deb::TabMidpoint>;
        case DEB_HEUN: return launch_fixed<Sys, deb::TabHeun>;
        case DEB_RALSTON: return launch_fixed<Sys, deb::TabRalston>;
        case DEB_SSP_RK3: return launch_fixed<Sys, deb::TabSspRk3>;
        case DEB_RK4: return launch_fixed<Sys, deb::TabRk4>;
        case DEB_THREE_EIGHTHS: return launch_fixed<Sys, deb::TabThreeEighths>;
    }
    return nullptr;
}

// (system, method) -> launcher.  The adaptive families live in their own translation units (ode_*.cu).
ode_launch_fn pick_ode(int system, int method, int* dim, int* np) {
    ode_launch_fn fixed = nullptr;
#define DEB_SYS_CASE(ID, T) case ID: *dim = deb::T::DIM; *np = deb::T::NP; fixed = pick_fixed_method<deb::T>(method); break;
    switch (system) {
        DEB_SYS_CASE(DEB_SYS_EXPONENTIAL, SysExponential)
        DEB_SYS_CASE(DEB_SYS_LINEAR, SysLinear)
        DEB_SYS_CASE(DEB_SYS_HARMONIC, SysHarmonic)
        DEB_SYS_CASE(DEB_SYS_LOGISTIC, SysLogistic)
        DEB_SYS_CASE(DEB_SYS_VAN_DER_POL, SysVanDerPol)
        DEB_SYS_CASE(DEB_SYS_LORENZ, SysLorenz)
        DEB_SYS_CASE(DEB_SYS_BRUSSELATOR, SysBrusselator)
        DEB_SYS_CASE(DEB_SYS_ROBERTSON, SysRobertson)
        default: *dim = -1; return nullptr;
    }
#undef DEB_SYS_CASE
    if (fixed) return fixed;
    switch (method) {
        case DEB_DOPRI5: case DEB_DOP853: return deb_pick_dopri(system, method);
        case DEB_RKF45: case DEB_CASH_KARP: return deb_pick_rkf(system, method);
        case DEB_RKV655E: case DEB_RKV656E: return deb_pick_rkv6(system, method);
        case DEB_RKV766E: case DEB_RKV767E: return deb_pick_rkv7(system, method);
        case DEB_RKV877E: case DEB_RKV878E: return deb_pick_rkv8(system, method);
        case DEB_RKV988E: case DEB_RKV989E: return deb_pick_rkv9(system, method);
    }
    return nullptr;
}

typedef int (*sde_launch_fn)(const deb::SdeKernelArgs&, int sms, cudaStream_t);
template <class Sde, class Tab, bool MILSTEIN = false>
int launch_sde(const deb::SdeKernelArgs& a, int sms, cudaStream_t st) {
    constexpr int BLOCK = 256;
    auto kern = deb::sde_ensemble_kernel<Sde, Tab, BLOCK, MILSTEIN>;
    long long blocks = (a.n_traj + BLOCK - 1) / BLOCK;
    const long long cap = (long long)sms * 8 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, BLOCK, 0, st>>>(a);
    DEB_CUDA(cudaGetLastError());
    return DEB_OK;
}
template <class Sde>
sde_launch_fn pick_sde_method(int method) {
    switch (method) {
        case DEB_EULER: return launch_sde<Sde, deb::TabEuler>;
        case DEB_MIDPOINT: return launch_sde<Sde, deb::TabMidpoint>;
        case DEB_HEUN: return launch_sde<Sde, deb::TabHeun>;
        case DEB_RALSTON: return launch_sde<Sde, deb::TabRalston>;
        case DEB_SSP_RK3: return launch_sde<Sde, deb::TabSspRk3>;
        case DEB_RK4: return launch_sde<Sde, deb::TabRk4>;
        case DEB_THREE_EIGHTHS: return launch_sde<Sde, deb::TabThreeEighths>;
        case DEB_MILSTEIN: return launch_sde<Sde, deb::TabEuler, true>;
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------------ user systems (NVRTC)
// A Rust `impl ODE for MySystem { fn diff(&self, t, y, dydt) }` (/root/reference/src/ode/ode.rs:20-44) cannot cross to
// the device as a closure.  The device-side equivalent: the caller hands over the BODY of diff as CUDA C++ text
// (deb_define_ode); the same kernel templates that serve the built-in systems are instantiated for it with NVRTC
// (sm_100a, --fmad=false) the first time a method is used, and cached.
#include "embedded_sources.inc"

struct NvrtcApi {
    void* handle = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
};

const NvrtcApi* nvrtc_api() {
    static NvrtcApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define DEB_SYM(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym)
        DEB_SYM(CreateProgram, "nvrtcCreateProgram");
        DEB_SYM(DestroyProgram, "nvrtcDestroyProgram");
        DEB_SYM(CompileProgram, "nvrtcCompileProgram");
        DEB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        DEB_SYM(GetProgramLog, "nvrtcGetProgramLog");
        DEB_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        DEB_SYM(GetCUBIN, "nvrtcGetCUBIN");
        DEB_SYM(AddNameExpression, "nvrtcAddNameExpression");
        DEB_SYM(GetLoweredName, "nvrtcGetLoweredName");
        DEB_SYM(GetErrorString, "nvrtcGetErrorString");
#undef DEB_SYM
        if (!api.CreateProgram || !api.CompileProgram || !api.GetCUBIN || !api.GetLoweredName) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api.handle ? &api : nullptr;
}

struct UserKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    bool adaptive = false;
    int block = 128;
};
struct UserSystem {
    int dim = 0, np = 0;
    std::string body;
};
std::mutex g_user_mu;
std::vector<std::unique_ptr<UserSystem>> g_user_systems;
const int USER_SYSTEM_BASE = 1000;
struct UserEvent { int dim = 0; std::string body; };
std::vector<std::unique_ptr<UserEvent>> g_user_events;  // ids USER_SYSTEM_BASE + index; guarded by g_user_mu
// run-time compiled kernels: (device, system id, method, per-step recorder, event id) -> loaded kernel.  Besides user systems,
// the per-step-recorder variants of the built-in systems are compiled on first use too (the ahead-of-time
// instantiations cover the t_eval / even(dt) recorders).  Guarded by g_user_mu.
std::map<std::tuple<int, int, int, int, int>, UserKernel> g_jit_kernels;

// built-in system id -> (struct name, dim, n_params)
const char* builtin_system_name(int system, int* dim, int* np) {
#define DEB_SYS_CASE(ID, T) case ID: *dim = deb::T::DIM; *np = deb::T::NP; return "deb::" #T;
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

const char* method_tab_name(int method, bool* adaptive) {
    *adaptive = false;
    switch (method) {
        case DEB_DOPRI5: *adaptive = true; return "deb::TabDopri5";
        case DEB_DOP853: *adaptive = true; return "deb::TabDop853";
        case DEB_RKF45: *adaptive = true; return "deb::TabRkf45";
        case DEB_CASH_KARP: *adaptive = true; return "deb::TabCashKarp";
        case DEB_RKV655E: *adaptive = true; return "deb::TabRkv655e";
        case DEB_RKV656E: *adaptive = true; return "deb::TabRkv656e";
        case DEB_RKV766E: *adaptive = true; return "deb::TabRkv766e";
        case DEB_RKV767E: *adaptive = true; return "deb::TabRkv767e";
        case DEB_RKV877E: *adaptive = true; return "deb::TabRkv877e";
        case DEB_RKV878E: *adaptive = true; return "deb::TabRkv878e";
        case DEB_RKV988E: *adaptive = true; return "deb::TabRkv988e";
        case DEB_RKV989E: *adaptive = true; return "deb::TabRkv989e";
        case DEB_EULER: return "deb::TabEuler";
        case DEB_MIDPOINT: return "deb::TabMidpoint";
        case DEB_HEUN: return "deb::TabHeun";
        case DEB_RALSTON: return "deb::TabRalston";
        case DEB_SSP_RK3: return "deb::TabSspRk3";
        case DEB_RK4: return "deb::TabRk4";
        case DEB_THREE_EIGHTHS: return "deb::TabThreeEighths";
    }
    return nullptr;
}

// Compile the ensemble kernel for (system, method, recorder kind) to a cubin (no device needed).  `us` = the user
// system, or null for a built-in one.
int compile_kernel_cubin(const UserSystem* us, int system, int method, bool rec, int event, std::vector<char>* cubin,
                         std::string* kernel_name, bool* is_adaptive) {
    int sdim = 0, snp = 0;
    const char* sys_name = us ? "deb::UserSys" : builtin_system_name(system, &sdim, &snp);
    if (!sys_name) return fail(DEB_ERR_BAD_ARG, "unknown system id");
    if (us) { sdim = us->dim; snp = us->np; }
    bool adaptive = false;
    const char* tab = method_tab_name(method, &adaptive);
    if (!tab) return fail(DEB_ERR_UNSUPPORTED, "unknown or unsupported method id");
    const NvrtcApi* rt = nvrtc_api();
    if (!rt) return fail(DEB_ERR_UNSUPPORTED, "libnvrtc not found: user-defined systems need the NVRTC runtime compiler");
    // occupancy hint: stage vectors live in registers, wider systems get the whole register file of fewer CTAs
    int min_blocks = 1;
    if (adaptive) {
        if (method == DEB_DOP853) min_blocks = sdim <= 2 ? 4 : sdim == 3 ? 3 : sdim <= 6 ? 2 : 1;
        else if (method >= DEB_RKV655E && method <= DEB_RKV989E) min_blocks = sdim <= 2 ? 4 : sdim == 3 ? 3 : sdim <= 5 ? 2 : 1;
        else min_blocks = sdim <= 3 ? 5 : sdim <= 6 ? 3 : sdim <= 10 ? 2 : 1;
        if (rec && min_blocks > 1) min_blocks -= 1;  // the recorder keeps the dense output of a step live
    }
    // event functor
    const UserEvent* ue = nullptr;
    std::string evt = "deb::EvtNone";
    if (event == DEB_EVENT_LINEAR) {
        evt = "deb::EvtLinear<" + std::to_string(sdim) + ">";
    } else if (event >= USER_SYSTEM_BASE) {
        const size_t k = (size_t)(event - USER_SYSTEM_BASE);
        if (k >= g_user_events.size()) return fail(DEB_ERR_BAD_ARG, "unknown event id");
        ue = g_user_events[k].get();
        if (ue->dim != sdim) return fail(DEB_ERR_BAD_ARG, "the event was defined for a different state dimension");
        evt = "deb::UserEvt";
    } else if (event != DEB_EVENT_NONE) {
        return fail(DEB_ERR_BAD_ARG, "unknown event id");
    }
    if (event != DEB_EVENT_NONE && !rec) return fail(DEB_ERR_BAD_ARG, "internal: events need a recorder kernel");
    char expr[384];
    if (adaptive) snprintf(expr, sizeof expr, "deb::dp_ensemble_kernel<%s, %s, 128, %d, false, %s, %s>", sys_name, tab, min_blocks, rec ? "true" : "false", evt.c_str());
    else snprintf(expr, sizeof expr, "deb::fixed_ensemble_kernel<%s, %s, 128, %s, %s>", sys_name, tab, rec ? "true" : "false", evt.c_str());
    std::string src;
    src += "#include \"erk_fixed.cuh\"\n#include \"systems.cuh\"\n";
    if (us) {
        src += "namespace deb {\nstruct UserSys {\n";
        src += "    static constexpr int DIM = " + std::to_string(us->dim) + ", NP = " + std::to_string(us->np) + ";\n";
        src += "    __device__ __forceinline__ static void rhs(double t, const double* y, double* dydt, const double* p) {\n";
        src += "        (void)t; (void)y; (void)p;\n";
        src += us->body;
        src += "\n    }\n};\n}  // namespace deb\n";
    }
    if (ue) {
        src += "namespace deb {\nstruct UserEvt {\n    static constexpr bool ENABLED = true;\n";
        src += "    __device__ __forceinline__ static double g(const OdeKernelArgs&, double t, const double* y, const double* p) {\n";
        src += "        (void)t; (void)y; (void)p;\n";
        src += ue->body;
        src += "\n    }\n};\n}  // namespace deb\n";
    }
    // headers: the embedded kernel sources + minimal stand-ins for the C headers NVRTC does not ship
    std::vector<const char*> hdr_names, hdr_text;
    for (const auto& e : deb_embedded_sources) { hdr_names.push_back(e.name); hdr_text.push_back(e.text); }
    static const char* k_stdint =
        "#pragma once\ntypedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
        "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";
    static const char* k_float = "#pragma once\n#define DBL_EPSILON 2.2204460492503131e-16\n#define DBL_MAX 1.7976931348623157e+308\n";
    static const char* k_abi =
        "#pragma once\n#define DEB_MAX_DIM 16\n"
        "enum { DEB_STATUS_COMPLETE = 0, DEB_STATUS_MAX_STEPS = 1, DEB_STATUS_STEP_SIZE = 2, DEB_STATUS_STIFFNESS = 3, DEB_STATUS_BAD_INPUT = 4, DEB_STATUS_INTERRUPTED = 5 };\n";
    hdr_names.push_back("stdint.h"); hdr_text.push_back(k_stdint);
    hdr_names.push_back("float.h"); hdr_text.push_back(k_float);
    hdr_names.push_back("../../include/deb_ensemble.h"); hdr_text.push_back(k_abi);
    nvrtcProgram prog = nullptr;
    nvrtcResult r = rt->CreateProgram(&prog, src.c_str(), "deb_user_system.cu", (int)hdr_names.size(), hdr_text.data(), hdr_names.data());
    if (r != NVRTC_SUCCESS) return fail(DEB_ERR_CUDA, std::string("nvrtcCreateProgram: ") + rt->GetErrorString(r));
    struct ProgGuard { const NvrtcApi* rt; nvrtcProgram* p; ~ProgGuard() { if (*p) rt->DestroyProgram(p); } } guard{rt, &prog};
    rt->AddNameExpression(prog, expr);
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo"};
    r = rt->CompileProgram(prog, 4, opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        rt->GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) rt->GetProgramLog(prog, &log[0]);
        return fail((us || ue) ? DEB_ERR_BAD_ARG : DEB_ERR_CUDA, ((us || ue) ? "the right-hand side / event function did not compile (NVRTC):\n" : "run-time kernel compilation failed (NVRTC):\n") + log);
    }
    const char* lowered = nullptr;
    r = rt->GetLoweredName(prog, expr, &lowered);
    if (r != NVRTC_SUCCESS || !lowered) return fail(DEB_ERR_CUDA, "nvrtcGetLoweredName failed");
    *kernel_name = lowered;
    size_t nbin = 0;
    rt->GetCUBINSize(prog, &nbin);
    cubin->resize(nbin);
    if (rt->GetCUBIN(prog, cubin->data()) != NVRTC_SUCCESS) return fail(DEB_ERR_CUDA, "nvrtcGetCUBIN failed");
    *is_adaptive = adaptive;
    if (const char* dump = getenv("DEB_DUMP_CUBIN")) {  // debugging aid: keep the last run-time compiled cubin for cuobjdump
        if (FILE* f = fopen(dump, "wb")) { fwrite(cubin->data(), 1, cubin->size(), f); fclose(f); }
    }
    return DEB_OK;
}

// Compile and load (once per device / system / method / recorder kind) a run-time kernel.  Caller holds g_user_mu.
int jit_kernel(const UserSystem* us, int device, int system, int method, bool rec, int event, UserKernel** out) {
    const auto key = std::make_tuple(device, system, method, rec ? 1 : 0, event);
    auto it = g_jit_kernels.find(key);
    if (it != g_jit_kernels.end()) { *out = &it->second; return DEB_OK; }
    std::vector<char> cubin;
    std::string lowered;
    UserKernel uk;
    if (int rc = compile_kernel_cubin(us, system, method, rec, event, &cubin, &lowered, &uk.adaptive)) return rc;
    DEB_CUDA(cudaLibraryLoadData(&uk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    DEB_CUDA(cudaLibraryGetKernel(&uk.kernel, uk.lib, lowered.c_str()));
    auto ins = g_jit_kernels.emplace(key, uk);
    *out = &ins.first->second;
    return DEB_OK;
}

int launch_user(const UserKernel& uk, const deb::OdeKernelArgs& a, int sms, cudaStream_t st) {
    long long blocks;
    if (uk.adaptive) {
        int per_sm = 0;
        DEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)uk.kernel, uk.block, 0));
        if (per_sm < 1) per_sm = 1;
        blocks = (long long)sms * per_sm;
    } else {
        blocks = (long long)sms * 16 * 8;
    }
    const long long need = (a.n_traj + uk.block - 1) / uk.block;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (getenv("DEB_DEBUG_LAUNCH")) fprintf(stderr, "[deb] run-time kernel: grid %lld x %d\n", blocks, uk.block);
    void* args[] = {(void*)&a};
    DEB_CUDA(cudaLaunchKernel((const void*)uk.kernel, dim3((unsigned)blocks), dim3(uk.block), args, 0, st));
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ t_eval plan
// TEvalSolout::new sorts the points by direction (stable; t_eval.rs:154-171).  The solout call that precedes the
// loop (solve_ivp.rs:160) consumes every point that is not after t0: the first one is emitted iff it equals t0,
// the others are skipped forever (t_eval.rs:121-129).  `rows` = the points a trajectory can still emit, in order.
struct TEvalPlan {
    std::vector<double> rows;
    bool emit_t0 = false;
};
int plan_t_eval(const double* t_eval, int n_eval, double t0, double tf, TEvalPlan* plan) {
    plan->rows.clear();
    plan->emit_t0 = false;
    if (n_eval <= 0) return DEB_OK;
    if (!t_eval) return fail(DEB_ERR_BAD_ARG, "n_eval > 0 but t_eval is NULL");
    std::vector<double> pts(t_eval, t_eval + n_eval);
    for (double v : pts)
        if (v != v) return fail(DEB_ERR_BAD_ARG, "t_eval contains NaN");
    const bool fwd = (tf - t0) > 0.0 || !((tf - t0) < 0.0);
    if (fwd) std::stable_sort(pts.begin(), pts.end(), [](double a, double b) { return a < b; });
    else std::stable_sort(pts.begin(), pts.end(), [](double a, double b) { return a > b; });
    for (int i = 0; i < n_eval; i++) {
        const double v = pts[i];
        if (i == 0 && v == t0) { plan->rows.push_back(v); plan->emit_t0 = true; continue; }
        const bool after = fwd ? (v > t0) : (v < t0);
        if (after) plan->rows.push_back(v);
    }
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ staging
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
    template <class T> T* as() { return (T*)p; }
};

// buffer from the stream-ordered memory pool: allocation and release are ordered on the stream, no device-wide sync
struct PoolBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~PoolBuf() { release(); }
    void release() { if (p) { cudaFreeAsync(p, st); p = nullptr; } }
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        st = s;
        DeviceInfo di;
        if (device_info(g_device, &di) != DEB_OK || !di.pool) return cudaErrorMemoryAllocation;
        return cudaMallocFromPoolAsync(&p, bytes ? bytes : 8, di.pool, s);
    }
    template <class T> T* as() { return (T*)p; }
};

struct ResultPtrs {
    double* y_eval; int* n_emitted; double* t_final; double* y_final; int* status; int* accepted; int* rejected; int* evals;
    double* t_out;
};

// Allocates device mirrors for the requested outputs (HOST memspace) or passes device pointers through.
struct ResultStage {
    DevBuf y_eval, n_emitted, t_final, y_final, status, accepted, rejected, evals, t_out;
    ResultPtrs dev{};
    int setup(const deb_result* R, bool host, long long n, int n_eval, int dim) {
        if (!host) {
            dev = {R->y_eval, R->n_emitted, R->t_final, R->y_final, R->status, R->accepted, R->rejected, R->evals, R->t_out};
            return DEB_OK;
        }
#define DEB_STAGE(field, T, count)                                         \
    if (R->field) {                                                        \
        DEB_CUDA(field.alloc(sizeof(T) * (size_t)(count)));                \
        dev.field = field.as<T>();                                         \
    }
        DEB_STAGE(y_eval, double, (size_t)n * n_eval * dim)
        DEB_STAGE(n_emitted, int, n)
        DEB_STAGE(t_final, double, n)
        DEB_STAGE(y_final, double, (size_t)n * dim)
        DEB_STAGE(status, int, n)
        DEB_STAGE(accepted, int, n)
        DEB_STAGE(rejected, int, n)
        DEB_STAGE(evals, int, n)
        DEB_STAGE(t_out, double, (size_t)n * n_eval)
#undef DEB_STAGE
        return DEB_OK;
    }
    int copy_back(const deb_result* R, long long n, int n_eval, int dim, cudaStream_t st) {
#define DEB_BACK(field, T, count) \
    if (R->field) DEB_CUDA(cudaMemcpyAsync(R->field, dev.field, sizeof(T) * (size_t)(count), cudaMemcpyDeviceToHost, st));
        DEB_BACK(y_eval, double, (size_t)n * n_eval * dim)
        DEB_BACK(n_emitted, int, n)
        DEB_BACK(t_final, double, n)
        DEB_BACK(y_final, double, (size_t)n * dim)
        DEB_BACK(status, int, n)
        DEB_BACK(accepted, int, n)
        DEB_BACK(rejected, int, n)
        DEB_BACK(evals, int, n)
        DEB_BACK(t_out, double, (size_t)n * n_eval)
#undef DEB_BACK
        return DEB_OK;
    }
};

int check_options(const deb_erk_options& o) {
    if (o.max_steps < 0) return fail(DEB_ERR_BAD_ARG, "max_steps must be >= 0");
    return DEB_OK;
}

void publish_rows(deb_result* R, const TEvalPlan& plan, bool even = false) {
    const size_t n = plan.rows.size() - (even ? 1 : 0);  // the EvenSolout tf sentinel is not a row time
    R->n_rows = (int32_t)n;
    if (R->t_rows)
        for (size_t i = 0; i < n; i++) R->t_rows[i] = plan.rows[i];
}

}  // namespace

extern "C" int deb_abi_version(void) { return DEB_ABI_VERSION; }

extern "C" int deb_trim_memory(int32_t device) {
    if (int rc = select_device(device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(device, &di)) return rc;
    DEB_CUDA(cudaDeviceSynchronize());
    DEB_CUDA(cudaMemPoolTrimTo(di.pool, 0));
    return DEB_OK;
}
extern "C" const char* deb_last_error(void) { return g_err.c_str(); }

extern "C" int deb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" void deb_erk_options_default(deb_erk_options* o) {  // erk/mod.rs:135-144
    if (!o) return;
    o->rtol = 1.0e-6;
    o->atol = 1.0e-6;
    o->rtol_vec = nullptr;
    o->atol_vec = nullptr;
    o->h0 = 0.0;
    o->h_min = 0.0;
    o->h_max = INFINITY;
    o->max_steps = 10000;
    o->safety_factor = 0.9;
    o->min_scale = 0.2;
    o->max_scale = 10.0;
    o->max_rejects = 100;
}

extern "C" int deb_solve_ode(const deb_ode_problem* P, deb_result* R) {
    if (!P || !R) return fail(DEB_ERR_BAD_ARG, "NULL problem/result");
    if (P->struct_size != sizeof(deb_ode_problem) || R->struct_size != sizeof(deb_result))
        return fail(DEB_ERR_BAD_ARG, "struct_size mismatch (ABI version skew)");
    int dim = 0, np = 0;
    std::function<int(const deb::OdeKernelArgs&, int, cudaStream_t)> launch;
    UserSystem* user = nullptr;
    if (P->system >= USER_SYSTEM_BASE) {
        std::lock_guard<std::mutex> lk(g_user_mu);
        const size_t k = (size_t)(P->system - USER_SYSTEM_BASE);
        if (k >= g_user_systems.size()) return fail(DEB_ERR_BAD_ARG, "unknown system id");
        user = g_user_systems[k].get();
        dim = user->dim;
        np = user->np;
        bool adaptive;
        if (!method_tab_name(P->method, &adaptive)) return fail(DEB_ERR_UNSUPPORTED, "unknown or unsupported method id");
    } else {
        ode_launch_fn builtin = pick_ode(P->system, P->method, &dim, &np);
        if (dim < 0) return fail(DEB_ERR_BAD_ARG, "unknown system id");
        if (!builtin) return fail(DEB_ERR_UNSUPPORTED, "unknown or unsupported method id");
        launch = builtin;
    }
    if (P->dim != dim || P->n_params != np) return fail(DEB_ERR_BAD_ARG, "dim / n_params do not match the system");
    if (P->n_traj < 0) return fail(DEB_ERR_BAD_ARG, "n_traj < 0");
    if (P->n_traj > 0 && (!P->y0 || (np > 0 && !P->params))) return fail(DEB_ERR_BAD_ARG, "NULL y0/params");
    if (P->n_eval < 0) return fail(DEB_ERR_BAD_ARG, "n_eval < 0");
    if (int rc = check_options(P->opt)) return rc;
    TEvalPlan plan;
    const bool even = (P->solout == DEB_SOLOUT_EVEN);
    const bool per_step = (P->solout == DEB_SOLOUT_DEFAULT || P->solout == DEB_SOLOUT_DENSE || P->solout == DEB_SOLOUT_CROSSING ||
                           P->solout == DEB_SOLOUT_HYPERPLANE);
    if (P->solout != DEB_SOLOUT_T_EVAL && !even && !per_step) return fail(DEB_ERR_BAD_ARG, "unknown solout mode");
    const bool has_event = (P->event != DEB_EVENT_NONE);
    const bool rec = per_step || has_event;  // rows with their own times: recorder kernels (compiled at first use)
    if (has_event) {
        if (P->event_direction < -1 || P->event_direction > 1) return fail(DEB_ERR_BAD_ARG, "event: direction must be -1, 0 or +1");
        if (P->event_terminate < 0) return fail(DEB_ERR_BAD_ARG, "event: terminate count < 0");
        if (P->row_capacity < 0) return fail(DEB_ERR_BAD_ARG, "row_capacity < 0");
        if (!R->y_eval) return fail(DEB_ERR_BAD_ARG, "event detection needs a y_eval buffer");
    }
    const int row_cap = (has_event && P->row_capacity > 0) ? P->row_capacity : P->n_eval;
    if (per_step) {
        // per-step recorders: no row plan; n_eval is the row capacity per trajectory
        if (!R->y_eval) return fail(DEB_ERR_BAD_ARG, "a per-step recorder needs a y_eval buffer");
        if (P->solout == DEB_SOLOUT_DENSE && P->dense_n < 0) return fail(DEB_ERR_BAD_ARG, "dense(n): n < 0");
        if (P->solout == DEB_SOLOUT_CROSSING) {
            if (P->cross_component < 0 || P->cross_component >= dim) return fail(DEB_ERR_BAD_ARG, "crossing: component index out of range");
            if (P->cross_direction < -1 || P->cross_direction > 1) return fail(DEB_ERR_BAD_ARG, "crossing: direction must be -1, 0 or +1");
        }
        if (P->solout == DEB_SOLOUT_HYPERPLANE) {
            if (P->plane_dim < 1 || P->plane_dim > dim) return fail(DEB_ERR_BAD_ARG, "hyperplane_crossing: plane_dim must be in 1..dim");
            for (int q = 0; q < P->plane_dim; q++)
                if (P->plane_index[q] < 0 || P->plane_index[q] >= dim) return fail(DEB_ERR_BAD_ARG, "hyperplane_crossing: component index out of range");
            if (P->cross_direction < -1 || P->cross_direction > 1) return fail(DEB_ERR_BAD_ARG, "hyperplane_crossing: direction must be -1, 0 or +1");
        }
    } else if (even) {
        // EvenSolout::solout (even.rs:69-199): t0 is emitted by the call before the loop, then last + dt*direction,
        // accumulated, while the point is not past tf
        if (!(P->even_dt > 0.0) || !(P->tf != P->t0)) return fail(DEB_ERR_BAD_ARG, "even(dt): dt must be > 0 and tf != t0");
        const double d = (P->tf > P->t0) ? 1.0 : -1.0;
        if (fabs(P->tf - P->t0) / P->even_dt > 1.0e8) return fail(DEB_ERR_BAD_ARG, "even(dt): more than 1e8 output points");
        for (double ti = P->t0; (d > 0.0) ? (ti <= P->tf) : (ti >= P->tf); ti += P->even_dt * d) plan.rows.push_back(ti);
        if (!R->y_eval) return fail(DEB_ERR_BAD_ARG, "even(dt) needs a y_eval buffer");
        plan.emit_t0 = true;
        plan.rows.push_back(P->tf);  // sentinel: reaching it triggers the final-point rule in the kernels (not a row time)
        if ((size_t)P->n_eval < plan.rows.size())
            return fail(DEB_ERR_BAD_ARG, "even(dt): n_eval (row capacity) must be at least floor(|tf-t0|/dt) + 2");
    } else if (int rc = plan_t_eval(P->t_eval, P->n_eval, P->t0, P->tf, &plan)) {
        return rc;
    }
    publish_rows(R, plan, even);
    R->kernel_ms = 0.f;
    R->total_ms = 0.f;
    if (P->n_traj == 0) return DEB_OK;  // empty ensemble: nothing to do, and no device needed
    if (int rc = select_device(P->device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(P->device, &di)) return rc;
    const bool jit = user || rec;
    if (jit) {  // user-defined right-hand side, or a per-step recorder: compile (first use) and bind the run-time kernel
        std::lock_guard<std::mutex> lk(g_user_mu);
        UserKernel* uk = nullptr;
        if (int rc = jit_kernel(user, P->device, P->system, P->method, rec, P->event, &uk)) return rc;
        const UserKernel ukc = *uk;
        launch = [ukc](const deb::OdeKernelArgs& ka, int sms, cudaStream_t s2) { return launch_user(ukc, ka, sms, s2); };
    }
    const bool host = (P->memspace == DEB_MEM_HOST);
    const long long n = P->n_traj;

    // ---- kernel arguments common to every launch of this call
    deb::OdeKernelArgs a;
    memset(&a, 0, sizeof a);
    // run-time kernels are built with SHARED_P = false: a shared set is read through the pointer (stride 0).  DEVICE calls
    // take the few bytes from the stream-ordered pool on the caller's stream (no device-wide synchronisation: the call
    // stays asynchronous); HOST calls are synchronous anyway
    DevBuf d_shared_params;
    PoolBuf p_shared_params;
    if (np > 0 && P->params_shared) {
        // one parameter set for the whole ensemble: HOST memory by contract, passed by value (constant bank)
        for (int q = 0; q < np && q < 8; q++) a.pc[q] = P->params[q];
        if (jit || np > 8) {
            if (P->memspace == DEB_MEM_HOST) {
                DEB_CUDA(d_shared_params.alloc(sizeof(double) * np));
                DEB_CUDA(cudaMemcpy(d_shared_params.p, P->params, sizeof(double) * np, cudaMemcpyHostToDevice));
            } else {
                DEB_CUDA(p_shared_params.alloc(sizeof(double) * np, (cudaStream_t)P->stream));
                DEB_CUDA(cudaMemcpyAsync(p_shared_params.p, P->params, sizeof(double) * np, cudaMemcpyHostToDevice, (cudaStream_t)P->stream));
            }
        }
    }
    const double* shared_params_dev = d_shared_params.p ? d_shared_params.as<double>() : p_shared_params.as<double>();
    a.params_stride = P->params_shared ? 0 : np;
    a.t0 = P->t0;
    a.tf = P->tf;
    for (int c = 0; c < DEB_MAX_DIM; c++) {
        a.rtol[c] = (P->opt.rtol_vec && c < dim) ? P->opt.rtol_vec[c] : P->opt.rtol;
        a.atol[c] = (P->opt.atol_vec && c < dim) ? P->opt.atol_vec[c] : P->opt.atol;
    }
    a.h0 = P->opt.h0;
    a.h_min = P->opt.h_min;
    a.h_max = P->opt.h_max;
    a.safety = P->opt.safety_factor;
    a.min_scale = P->opt.min_scale;
    a.max_scale = P->opt.max_scale;
    a.max_steps = (int)std::min<int64_t>(P->opt.max_steps, 0x7fffffff / 16);
    a.max_rejects = (int)std::min<int64_t>(std::max<int64_t>(P->opt.max_rejects, 0), 0x7fffffff);
    {   // fixed-step methods: the step schedule is the same for every trajectory; run the solve_ode bookkeeping here
        // (validate_step_size_parameters utils.rs:60-157 with the method's h_min / h_max; loop solve_ivp.rs:193-209, :263)
        bool adaptive_m = false;
        method_tab_name(P->method, &adaptive_m);
        a.fx_n_steps = 0;
        a.fx_h_last = 0.0;
        a.fx_status = DEB_STATUS_COMPLETE;
        if (!adaptive_m) {
            const double t0 = P->t0, tf = P->tf;
            const double dd = tf - t0;
            const double dir = (dd != dd) ? dd : copysign(1.0, dd);
            double h = P->opt.h0;
            if (h == 0.0) h = fabs(tf - t0) / 100.0;
            const double sgh = (h != h) ? h : copysign(1.0, h);
            const bool ok = (tf != t0) && (dir == 1.0 || dir == -1.0) && sgh == dir && !(P->opt.h_min < 0.0) && !(P->opt.h_max < 0.0) &&
                            !(P->opt.h_min > P->opt.h_max) && !(fabs(h) < P->opt.h_min) && !(fabs(h) > P->opt.h_max) &&
                            !(fabs(h) > fabs(tf - t0)) && h != 0.0;
            a.fx_h_last = h;
            if (!ok) {
                a.fx_status = DEB_STATUS_BAD_INPUT;
            } else {
                const double eps10 = 2.220446049250313e-16 * 10.0;
                double t = t0;
                long long steps = 0;
                for (;;) {
                    if ((t + h - tf) * dir > 0.0) {
                        const double h_new = tf - t;
                        if (fabs(h_new) < eps10) break;
                        h = h_new;
                    }
                    if (steps >= a.max_steps) { a.fx_status = DEB_STATUS_MAX_STEPS; break; }
                    steps += 1;
                    a.fx_h_last = h;  // only the final step can differ from h0 (the clip at tf)
                    t += h;
                    if (fabs(tf - t) <= eps10) break;
                }
                a.fx_n_steps = (int)steps;
            }
        }
    }
    a.n_rows = (int)plan.rows.size();
    a.row_stride = row_cap;
    a.emit_t0 = plan.emit_t0 ? 1 : 0;
    a.even = even ? 1 : 0;
    a.even_tol = fabs(P->even_dt) * 1e-12 + 2.220446049250313e-16 * 10.0;
    a.rec_mode = P->solout;
    a.event_direction = P->event_direction;
    a.event_terminate = P->event_terminate;
    for (int c = 0; c < DEB_MAX_DIM + 2; c++) a.event_coef[c] = P->event_coef[c];
    if (P->solout == DEB_SOLOUT_HYPERPLANE) {
        // HyperplaneCrossingSolout::new (hyperplane.rs:124-139): normal /= ||normal|| unless the norm is below epsilon
        a.plane_dim = P->plane_dim;
        double nsq = 0.0;
        for (int q = 0; q < P->plane_dim; q++) nsq = nsq + P->plane_normal[q] * P->plane_normal[q];
        const double norm = sqrt(nsq);
        const double inv = 1.0 / norm;
        for (int q = 0; q < P->plane_dim; q++) {
            a.plane_index[q] = P->plane_index[q];
            a.plane_point[q] = P->plane_point[q];
            a.plane_normal[q] = (norm > 2.220446049250313e-16) ? P->plane_normal[q] * inv : P->plane_normal[q];
        }
    }
    a.dense_n = P->dense_n;
    a.cross_component = P->cross_component;
    a.cross_direction = P->cross_direction;
    a.cross_threshold = P->cross_threshold;
    const bool per_traj_params = (np > 0 && !P->params_shared);
    const size_t rows_bytes = sizeof(double) * plan.rows.size();

    if (!host) {
        // ---- DEVICE memspace: pointers are device pointers; enqueue on the caller's stream and return
        cudaStream_t st = (cudaStream_t)P->stream;
        void* d_small = nullptr;  // [queue counter (8 B)] [rows]
        DEB_CUDA(cudaMallocAsync(&d_small, 8 + rows_bytes, st));
        struct SmallFree { void* p; cudaStream_t st; ~SmallFree() { if (p) cudaFreeAsync(p, st); } } small_free{d_small, st};
        DEB_CUDA(cudaMemsetAsync(d_small, 0, 8, st));
        if (rows_bytes) DEB_CUDA(cudaMemcpyAsync((char*)d_small + 8, plan.rows.data(), rows_bytes, cudaMemcpyHostToDevice, st));
        a.queue = (unsigned long long*)d_small;
        a.t_rows = (const double*)((char*)d_small + 8);
        a.y0 = P->y0;
        a.params = per_traj_params ? P->params : shared_params_dev;
        a.n_traj = n;
        a.y_eval = R->y_eval; a.n_emitted = R->n_emitted; a.t_final = R->t_final; a.y_final = R->y_final;
        a.status = R->status; a.accepted = R->accepted; a.rejected = R->rejected; a.evals = R->evals;
        a.t_out = R->t_out;
        return launch(a, di.sms, st);
    }

    // ---- HOST memspace: pipelined chunks.  Two slots, each with its own stream and device buffers, run
    //      H2D(y0) -> kernel -> D2H(results) for alternating chunks, so the result copy of one chunk (the bulk of the
    //      PCIe traffic: n_eval*dim*8 B per trajectory) overlaps the integration of the next, and the tail of one
    //      persistent kernel overlaps the start of the following one.  Device memory is 2 chunks, not the ensemble.
    const auto wall0 = std::chrono::steady_clock::now();
    // 2 Mi trajectories: measured best for C2 (tail loss vs exposed last copy), profiles/; smaller ensembles are still cut
    // into about four chunks (not below 256 Ki) so that their result copy overlaps the integration too
    long long CHUNK = std::min<long long>(1ll << 21, std::max<long long>(n / 4, 1ll << 18));
    if (const char* e = getenv("DEB_HOST_CHUNK")) {  // tuning / test knob
        const long long v = atoll(e);
        if (v > 0) CHUNK = v;
    }
    // Chunks of at most CHUNK trajectories; the last ones shrink geometrically (each takes half of what is left, down to
    // CHUNK/16), because the result copy of the LAST chunk is the one transfer nothing overlaps with (measured on C2:
    // 1845 ms per pass instead of 1930 ms with equal chunks).
    std::vector<long long> chunk_sizes;
    {
        long long min_chunk = std::max<long long>(CHUNK / 16, 1);
        if (const char* e = getenv("DEB_HOST_TAIL")) {  // tuning knob: 0 = equal chunks, k = smallest chunk CHUNK/k
            const long long v = atoll(e);
            min_chunk = (v <= 0) ? CHUNK : std::max<long long>(CHUNK / v, 1);
        }
        long long left = n;
        while (left > 0) {
            long long c = std::min(CHUNK, left);
            if (left > CHUNK) c = std::min(CHUNK, std::max(min_chunk, left / 2));
            else if (left > 2 * min_chunk && n > CHUNK) c = std::max(min_chunk, left / 2);
            chunk_sizes.push_back(c);
            left -= c;
        }
    }
    const long long chunk = *std::max_element(chunk_sizes.begin(), chunk_sizes.end());  // buffer size of a slot
    const int n_slots = (chunk_sizes.size() > 1) ? 2 : 1;
    const int n_eval = row_cap;  // rows per trajectory in y_eval / t_out
    struct Slot {
        cudaStream_t st = nullptr;
        cudaEvent_t k0 = nullptr, k1 = nullptr;
        PoolBuf y0, params, small, y_eval, n_emitted, t_final, y_final, status, accepted, rejected, evals, t_out;
        bool used = false;
        ~Slot() {
            // stream-ordered frees first: they use the stream that is destroyed below
            for (PoolBuf* b : {&y0, &params, &small, &y_eval, &n_emitted, &t_final, &y_final, &status, &accepted, &rejected, &evals, &t_out}) b->release();
            if (k0) cudaEventDestroy(k0);
            if (k1) cudaEventDestroy(k1);
            if (st) cudaStreamDestroy(st);
        }
    } slot[2];
    for (int s = 0; s < n_slots; s++) {
        Slot& S = slot[s];
        DEB_CUDA(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
        DEB_CUDA(cudaEventCreate(&S.k0));
        DEB_CUDA(cudaEventCreate(&S.k1));
        DEB_CUDA(S.y0.alloc(sizeof(double) * (size_t)chunk * dim, S.st));
        if (per_traj_params) DEB_CUDA(S.params.alloc(sizeof(double) * (size_t)chunk * np, S.st));
        DEB_CUDA(S.small.alloc(8 + rows_bytes, S.st));
        if (rows_bytes) DEB_CUDA(cudaMemcpyAsync((char*)S.small.p + 8, plan.rows.data(), rows_bytes, cudaMemcpyHostToDevice, S.st));
        if (R->y_eval) DEB_CUDA(S.y_eval.alloc(sizeof(double) * (size_t)chunk * n_eval * dim, S.st));
        if (R->n_emitted) DEB_CUDA(S.n_emitted.alloc(sizeof(int) * (size_t)chunk, S.st));
        if (R->t_final) DEB_CUDA(S.t_final.alloc(sizeof(double) * (size_t)chunk, S.st));
        if (R->y_final) DEB_CUDA(S.y_final.alloc(sizeof(double) * (size_t)chunk * dim, S.st));
        if (R->status) DEB_CUDA(S.status.alloc(sizeof(int) * (size_t)chunk, S.st));
        if (R->accepted) DEB_CUDA(S.accepted.alloc(sizeof(int) * (size_t)chunk, S.st));
        if (R->rejected) DEB_CUDA(S.rejected.alloc(sizeof(int) * (size_t)chunk, S.st));
        if (R->evals) DEB_CUDA(S.evals.alloc(sizeof(int) * (size_t)chunk, S.st));
        if (R->t_out) DEB_CUDA(S.t_out.alloc(sizeof(double) * (size_t)chunk * n_eval, S.st));
    }
    float kernel_ms = 0.f;
    int ci = 0;
    for (long long off = 0; off < n; off += chunk_sizes[ci], ci++) {
        Slot& S = slot[ci % n_slots];
        const long long cnt = chunk_sizes[ci];
        if (S.used) {  // collect the kernel time of the chunk that used this slot before (its stream has passed k1)
            DEB_CUDA(cudaEventSynchronize(S.k1));
            float ms = 0.f;
            DEB_CUDA(cudaEventElapsedTime(&ms, S.k0, S.k1));
            kernel_ms += ms;
        }
        DEB_CUDA(cudaMemcpyAsync(S.y0.p, P->y0 + (size_t)off * dim, sizeof(double) * (size_t)cnt * dim, cudaMemcpyHostToDevice, S.st));
        if (per_traj_params)
            DEB_CUDA(cudaMemcpyAsync(S.params.p, P->params + (size_t)off * np, sizeof(double) * (size_t)cnt * np, cudaMemcpyHostToDevice, S.st));
        DEB_CUDA(cudaMemsetAsync(S.small.p, 0, 8, S.st));
        deb::OdeKernelArgs ac = a;
        ac.queue = (unsigned long long*)S.small.p;
        ac.t_rows = (const double*)((char*)S.small.p + 8);
        ac.y0 = S.y0.as<double>();
        ac.params = per_traj_params ? S.params.as<double>() : shared_params_dev;
        ac.n_traj = cnt;
        ac.y_eval = S.y_eval.as<double>(); ac.n_emitted = S.n_emitted.as<int>(); ac.t_final = S.t_final.as<double>();
        ac.y_final = S.y_final.as<double>(); ac.status = S.status.as<int>(); ac.accepted = S.accepted.as<int>();
        ac.rejected = S.rejected.as<int>(); ac.evals = S.evals.as<int>();
        ac.t_out = S.t_out.as<double>();
        DEB_CUDA(cudaEventRecord(S.k0, S.st));
        if (int rc = launch(ac, di.sms, S.st)) return rc;
        DEB_CUDA(cudaEventRecord(S.k1, S.st));
        S.used = true;
#define DEB_BACK(field, T, per)                                                                                   \
    if (R->field)                                                                                                 \
        DEB_CUDA(cudaMemcpyAsync(R->field + (size_t)off * (per), S.field.p, sizeof(T) * (size_t)cnt * (per), cudaMemcpyDeviceToHost, S.st));
        DEB_BACK(y_eval, double, (size_t)n_eval * dim)
        DEB_BACK(n_emitted, int, 1)
        DEB_BACK(t_final, double, 1)
        DEB_BACK(y_final, double, dim)
        DEB_BACK(status, int, 1)
        DEB_BACK(accepted, int, 1)
        DEB_BACK(rejected, int, 1)
        DEB_BACK(evals, int, 1)
        DEB_BACK(t_out, double, (size_t)n_eval)
#undef DEB_BACK
    }
    for (int s = 0; s < n_slots; s++) {
        DEB_CUDA(cudaStreamSynchronize(slot[s].st));
        if (slot[s].used) {
            float ms = 0.f;
            DEB_CUDA(cudaEventElapsedTime(&ms, slot[s].k0, slot[s].k1));
            kernel_ms += ms;
        }
    }
    R->kernel_ms = kernel_ms;  // sum over chunks (chunks on the two streams overlap: can exceed the wall time)
    R->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return DEB_OK;
}

extern "C" int deb_define_ode(int32_t dim, int32_t n_params, const char* diff_body, int32_t* system_id) {
    if (!diff_body || !system_id) return fail(DEB_ERR_BAD_ARG, "NULL argument");
    if (dim < 1 || dim > DEB_MAX_DIM) return fail(DEB_ERR_BAD_ARG, "dim must be in 1..DEB_MAX_DIM");
    if (n_params < 0 || n_params > 64) return fail(DEB_ERR_BAD_ARG, "n_params must be in 0..64");
    std::lock_guard<std::mutex> lk(g_user_mu);
    std::unique_ptr<UserSystem> us(new UserSystem);
    us->dim = dim;
    us->np = n_params;
    us->body = diff_body;
    g_user_systems.push_back(std::move(us));
    *system_id = USER_SYSTEM_BASE + (int32_t)g_user_systems.size() - 1;
    return DEB_OK;
}

extern "C" int deb_define_event(int32_t dim, const char* event_body, int32_t* event_id) {
    if (!event_body || !event_id) return fail(DEB_ERR_BAD_ARG, "NULL argument");
    if (dim < 1 || dim > DEB_MAX_DIM) return fail(DEB_ERR_BAD_ARG, "dim must be in 1..DEB_MAX_DIM");
    std::lock_guard<std::mutex> lk(g_user_mu);
    std::unique_ptr<UserEvent> ue(new UserEvent);
    ue->dim = dim;
    ue->body = event_body;
    g_user_events.push_back(std::move(ue));
    *event_id = USER_SYSTEM_BASE + (int32_t)g_user_events.size() - 1;
    return DEB_OK;
}

extern "C" int deb_check_ode(int32_t system_id, int32_t method, int32_t solout, int32_t event) {
    std::lock_guard<std::mutex> lk(g_user_mu);
    const UserSystem* us = nullptr;
    if (system_id >= USER_SYSTEM_BASE) {
        const int u = system_id - USER_SYSTEM_BASE;
        if (u >= (int)g_user_systems.size()) return fail(DEB_ERR_BAD_ARG, "unknown system id");
        us = g_user_systems[u].get();
    }
    const bool rec = (solout == DEB_SOLOUT_DEFAULT || solout == DEB_SOLOUT_DENSE || solout == DEB_SOLOUT_CROSSING ||
                      solout == DEB_SOLOUT_HYPERPLANE || event != DEB_EVENT_NONE);
    if (!us && !rec) {  // built-in system with a row-plan recorder: compiled ahead of time
        int dim = 0, np = 0;
        if (!pick_ode(system_id, method, &dim, &np)) return fail(dim < 0 ? DEB_ERR_BAD_ARG : DEB_ERR_UNSUPPORTED, "unknown system or method id");
        return DEB_OK;
    }
    std::vector<char> cubin;
    std::string name;
    bool adaptive = false;
    return compile_kernel_cubin(us, system_id, method, rec, event, &cubin, &name, &adaptive);
}

extern "C" int deb_solve_sde(const deb_sde_problem* P, deb_result* R) {
    if (!P || !R) return fail(DEB_ERR_BAD_ARG, "NULL problem/result");
    if (P->struct_size != sizeof(deb_sde_problem) || R->struct_size != sizeof(deb_result))
        return fail(DEB_ERR_BAD_ARG, "struct_size mismatch (ABI version skew)");
    sde_launch_fn launch = nullptr;
    int np = 0, dim = 1;
    if (P->system == DEB_SDE_OU) { launch = pick_sde_method<deb::SdeOU>(P->method); np = deb::SdeOU::NP; }
    else if (P->system == DEB_SDE_GBM) { launch = pick_sde_method<deb::SdeGBM>(P->method); np = deb::SdeGBM::NP; }
    else if (P->system == DEB_SDE_HESTON) { launch = pick_sde_method<deb::SdeHeston>(P->method); np = deb::SdeHeston::NP; dim = deb::SdeHeston::DIM; }
    else return fail(DEB_ERR_BAD_ARG, "unknown SDE system id");
    if (!launch) return fail(DEB_ERR_UNSUPPORTED, "SDE ensembles take a fixed-step method id or DEB_MILSTEIN");
    if (P->dim != dim || P->n_params != np) return fail(DEB_ERR_BAD_ARG, "dim / n_params do not match the SDE system");
    if (P->n_traj < 0 || P->n_eval < 0) return fail(DEB_ERR_BAD_ARG, "negative size");
    if (P->n_traj > 0 && (!P->y0 || !P->params)) return fail(DEB_ERR_BAD_ARG, "NULL y0/params");
    if (int rc = check_options(P->opt)) return rc;
    TEvalPlan plan;
    if (int rc = plan_t_eval(P->t_eval, P->n_eval, P->t0, P->tf, &plan)) return rc;
    publish_rows(R, plan);
    R->kernel_ms = 0.f;
    R->total_ms = 0.f;
    if (P->n_traj == 0) return DEB_OK;
    if (int rc = select_device(P->device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(P->device, &di)) return rc;
    const bool host = (P->memspace == DEB_MEM_HOST);
    cudaStream_t st = host ? (cudaStream_t)0 : (cudaStream_t)P->stream;
    const long long n = P->n_traj;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (host) {
        for (auto& e : ev) DEB_CUDA(cudaEventCreate(&e));
        DEB_CUDA(cudaEventRecord(ev[0], st));
    }
    deb::SdeKernelArgs a;
    memset(&a, 0, sizeof a);
    DevBuf d_y0, d_params;
    // y0: [n_traj] in `memspace`, or ONE value in HOST memory when y0_shared; params likewise (see the header)
    double y0_one[DEB_MAX_DIM];
    if (P->y0_shared) {
        for (int c = 0; c < dim; c++) y0_one[c] = P->y0[c];
        DEB_CUDA(d_y0.alloc(sizeof(double) * dim));
        DEB_CUDA(cudaMemcpyAsync(d_y0.p, y0_one, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
        a.y0 = d_y0.as<double>();
    } else if (host) {
        DEB_CUDA(d_y0.alloc(sizeof(double) * (size_t)n * dim));
        DEB_CUDA(cudaMemcpyAsync(d_y0.p, P->y0, sizeof(double) * (size_t)n * dim, cudaMemcpyHostToDevice, st));
        a.y0 = d_y0.as<double>();
    } else {
        a.y0 = P->y0;
    }
    if (P->params_shared) {
        for (int q = 0; q < np && q < 8; q++) a.pc[q] = P->params[q];
        a.params = nullptr;
    } else if (host) {
        DEB_CUDA(d_params.alloc(sizeof(double) * (size_t)n * np));
        DEB_CUDA(cudaMemcpyAsync(d_params.p, P->params, sizeof(double) * (size_t)n * np, cudaMemcpyHostToDevice, st));
        a.params = d_params.as<double>();
    } else {
        a.params = P->params;
    }
    a.y0_stride = P->y0_shared ? 0 : dim;
    a.params_stride = P->params_shared ? 0 : np;
    a.n_traj = n;
    a.path_offset = P->path_offset;
    a.seed = P->seed;
    for (int r = 0; r < 10; r++) {
        a.round_keys[2 * r] = (unsigned int)(P->seed & 0xffffffffu) + (unsigned int)r * 0x9E3779B9u;
        a.round_keys[2 * r + 1] = (unsigned int)(P->seed >> 32) + (unsigned int)r * 0xBB67AE85u;
    }
    a.t0 = P->t0;
    a.tf = P->tf;
    a.h0 = P->opt.h0;
    a.h_min = P->opt.h_min;
    a.h_max = P->opt.h_max;
    a.max_steps = (int)std::min<int64_t>(P->opt.max_steps, 0x7fffffff / 16);
    // ---- the step schedule, identical for every path: solve_sde bookkeeping (solve_ivp.rs:211-227, :263),
    //      Fixed/Milstein init and max_steps (stochastic.rs:18-65, :74-83), TEvalSolout row placement (t_eval.rs:100-130)
    std::vector<int> row_step(plan.rows.size(), -1);
    std::vector<double> row_s(plan.rows.size(), -1.0);
    {
        const double t0 = P->t0, tf = P->tf;
        const double dd = tf - t0;
        const double dir = (dd != dd) ? dd : copysign(1.0, dd);
        double h = P->opt.h0;
        if (h == 0.0) h = fabs(tf - t0) / 100.0;
        const double sgh = (h != h) ? h : copysign(1.0, h);
        const bool ok = (tf != t0) && (dir == 1.0 || dir == -1.0) && sgh == dir && !(P->opt.h_min < 0.0) && !(P->opt.h_max < 0.0) &&
                        !(P->opt.h_min > P->opt.h_max) && !(fabs(h) < P->opt.h_min) && !(fabs(h) > P->opt.h_max) &&
                        !(fabs(h) > fabs(tf - t0)) && h != 0.0;  // validate_step_size_parameters, utils.rs:60-157
        a.n_steps = 0;
        a.h_last = h;
        a.final_status = ok ? DEB_STATUS_COMPLETE : DEB_STATUS_BAD_INPUT;
        if (ok) {
            const double eps10 = 2.220446049250313e-16 * 10.0;
            double t = t0;
            long long steps = 0;
            size_t idx = plan.emit_t0 ? 1 : 0;
            for (;;) {
                if ((t + h - tf) * dir > 0.0) {
                    const double h_new = tf - t;
                    if (fabs(h_new) < eps10) break;
                    h = h_new;
                }
                if (steps >= a.max_steps) { a.final_status = DEB_STATUS_MAX_STEPS; break; }
                const double t_new = t + h;
                while (idx < plan.rows.size() && ((dir > 0.0) ? (plan.rows[idx] <= t_new) : (plan.rows[idx] >= t_new))) {
                    row_step[idx] = (int)steps;
                    row_s[idx] = (plan.rows[idx] == t_new) ? -1.0 : (plan.rows[idx] - t) / (t_new - t);
                    idx++;
                }
                steps += 1;
                a.h_last = h;  // only the final step can differ from h0 (the clip at tf)
                t = t_new;
                if (fabs(tf - t) <= eps10) break;
            }
            a.n_steps = (int)steps;
        }
    }
    const size_t nr = plan.rows.size();
    const size_t rows_bytes = 8 + sizeof(double) * nr * 2 + sizeof(int) * nr;
    void* d_rows = nullptr;
    DEB_CUDA(cudaMallocAsync(&d_rows, rows_bytes, st));
    struct SmallFree { void* p; cudaStream_t st; ~SmallFree() { if (p) cudaFreeAsync(p, st); } } small_free{d_rows, st};
    if (nr) {
        DEB_CUDA(cudaMemcpyAsync(d_rows, plan.rows.data(), sizeof(double) * nr, cudaMemcpyHostToDevice, st));
        DEB_CUDA(cudaMemcpyAsync((char*)d_rows + sizeof(double) * nr, row_s.data(), sizeof(double) * nr, cudaMemcpyHostToDevice, st));
        DEB_CUDA(cudaMemcpyAsync((char*)d_rows + sizeof(double) * nr * 2, row_step.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, st));
    }
    a.t_rows = (const double*)d_rows;
    a.row_s = (const double*)((char*)d_rows + sizeof(double) * nr);
    a.row_step = (const int*)((char*)d_rows + sizeof(double) * nr * 2);
    a.n_rows = (int)nr;
    a.row_stride = P->n_eval;
    a.emit_t0 = plan.emit_t0 ? 1 : 0;
    ResultStage rs;
    if (int rc = rs.setup(R, host, n, P->n_eval, dim)) return rc;
    a.y_eval = rs.dev.y_eval;
    a.n_emitted = rs.dev.n_emitted;
    a.t_final = rs.dev.t_final;
    a.y_final = rs.dev.y_final;
    a.status = rs.dev.status;
    a.accepted = rs.dev.accepted;
    a.rejected = rs.dev.rejected;
    a.evals = rs.dev.evals;
    if (host) DEB_CUDA(cudaEventRecord(ev[1], st));
    if (int rc = launch(a, di.sms, st)) return rc;
    if (host) {
        DEB_CUDA(cudaEventRecord(ev[2], st));
        if (int rc = rs.copy_back(R, n, P->n_eval, dim, st)) return rc;
        DEB_CUDA(cudaEventRecord(ev[3], st));
        DEB_CUDA(cudaStreamSynchronize(st));
        DEB_CUDA(cudaEventElapsedTime(&R->kernel_ms, ev[1], ev[2]));
        DEB_CUDA(cudaEventElapsedTime(&R->total_ms, ev[0], ev[3]));
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ heat MoL
namespace {

template <class Tab, int STAGE>
int heat_launch_stage(const deb::HeatArgs& a, bool pow2, cudaStream_t st) {
    const long long pairs = (a.n + 1) / 2;
    const unsigned blocks = (unsigned)((pairs + 255) / 256);
    if (pow2) deb::heat_stage_kernel<Tab, STAGE, true><<<blocks, 256, 0, st>>>(a);
    else deb::heat_stage_kernel<Tab, STAGE, false><<<blocks, 256, 0, st>>>(a);
    DEB_CUDA(cudaGetLastError());
    return DEB_OK;
}

// launches stages STAGE..S-1 (each writes ks[STAGE])
// One whole time step: a single launch of the all-stages-in-registers kernel (mol_heat.cuh); 8 independent warps per CTA.
template <class Tab>
int heat_launch_step(const deb::HeatArgs& a, bool pow2, cudaStream_t st) {
    const long long warps = (a.n + deb::HEAT_OUT - 1) / deb::HEAT_OUT;
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    if (pow2) deb::heat_step_kernel<Tab, true><<<blocks, 256, 0, st>>>(a);
    else deb::heat_step_kernel<Tab, false><<<blocks, 256, 0, st>>>(a);
    DEB_CUDA(cudaGetLastError());
    return DEB_OK;
}

// solve_ode loop (solve_ivp.rs:193-263) + Fixed::init/step (fixed/ordinary.rs:16-139) around the launches
template <class Tab>
int heat_solve(const deb_heat_problem* P, double* y_a, double* y_b, bool pow2, cudaStream_t st, double* t_out, long long* steps_out,
               int* status_out, double** y_out) {
    const double t0 = P->t0, tf = P->tf;
    const double d = tf - t0;
    const double dir = (d != d) ? d : copysign(1.0, d);
    *t_out = t0; *steps_out = 0; *y_out = y_a;
    if (!(dir == 1.0 || dir == -1.0) || tf == t0) { *status_out = DEB_STATUS_BAD_INPUT; return DEB_OK; }
    double h = P->h;
    if (h == 0.0) h = fabs(tf - t0) / 100.0;  // fixed/ordinary.rs:23-28
    // validate_step_size_parameters (utils.rs:60-157) with h_min = 0, h_max = inf
    const double sg = (h != h) ? h : copysign(1.0, h);
    if (sg != dir || fabs(h) > fabs(tf - t0) || h == 0.0) { *status_out = DEB_STATUS_BAD_INPUT; return DEB_OK; }
    deb::HeatArgs a;
    memset(&a, 0, sizeof a);
    a.n = P->n_nodes;
    a.dx = (P->hi - P->lo) / (double)(P->n_nodes - 1);  // grid.rs:23-36
    a.inv_dx = 1.0 / a.dx;
    a.alpha = P->alpha;
    a.bc_lo_kind = P->bc_lower_kind; a.bc_hi_kind = P->bc_upper_kind;
    a.bc_lo_val = P->bc_lower_value; a.bc_hi_val = P->bc_upper_value;
    double* y = y_a;
    double* y_next = y_b;
    double t = t0;
    long long steps = 0;
    int status = DEB_STATUS_COMPLETE;
    const double eps10 = 2.220446049250313e-16 * 10.0;
    for (;;) {
        if ((t + h - tf) * dir > 0.0) {
            const double h_new = tf - t;
            if (fabs(h_new) < eps10) break;
            h = h_new;
        }
        if (steps >= P->max_steps) { status = DEB_STATUS_MAX_STEPS; break; }
        steps += 1;
        a.y = y; a.h = h; a.out_y = y_next;
        if (int rc = heat_launch_step<Tab>(a, pow2, st)) return rc;
        t += h;
        std::swap(y, y_next);
        if (fabs(tf - t) <= eps10) break;
    }
    *t_out = t; *steps_out = steps; *status_out = status; *y_out = y;
    return DEB_OK;
}

bool is_pow2(double x) {
    if (!(x > 0.0) || isinf(x)) return false;
    int e;
    return frexp(x, &e) == 0.5 && e > -1000 && e < 1000;
}

}  // namespace

extern "C" int deb_solve_heat_mol(const deb_heat_problem* P) {
    if (!P) return fail(DEB_ERR_BAD_ARG, "NULL problem");
    if (P->struct_size != sizeof(deb_heat_problem)) return fail(DEB_ERR_BAD_ARG, "struct_size mismatch (ABI version skew)");
    if (P->n_nodes < 2) return fail(DEB_ERR_BAD_ARG, "StructuredGrid requires at least two nodes per axis");
    if (P->lo == P->hi) return fail(DEB_ERR_BAD_ARG, "StructuredGrid endpoints must be distinct");
    if (!P->u0 || !P->u_final) return fail(DEB_ERR_BAD_ARG, "NULL u0/u_final");
    if ((P->bc_lower_kind | P->bc_upper_kind) & ~1) return fail(DEB_ERR_BAD_ARG, "boundary kind must be 0 (Dirichlet) or 1 (Neumann)");
    if (int rc = select_device(P->device)) return rc;
    const bool host = (P->memspace == DEB_MEM_HOST);
    cudaStream_t st = host ? (cudaStream_t)0 : (cudaStream_t)P->stream;
    const size_t bytes = sizeof(double) * (size_t)P->n_nodes;
    switch (P->method) {
        case DEB_EULER: case DEB_MIDPOINT: case DEB_HEUN: case DEB_RALSTON: case DEB_SSP_RK3: case DEB_RK4: case DEB_THREE_EIGHTHS: break;
        default: return fail(DEB_ERR_UNSUPPORTED, "method of lines takes a fixed-step method id");
    }
    // work buffers: two state buffers (ping-pong) from the stream-ordered pool (no device-wide synchronisation on
    // allocation or release); the stage derivatives never leave the chip
    PoolBuf ya, yb;
    DEB_CUDA(ya.alloc(bytes, st));
    DEB_CUDA(yb.alloc(bytes, st));
    DEB_CUDA(cudaMemcpyAsync(ya.p, P->u0, bytes, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    const double dx = (P->hi - P->lo) / (double)(P->n_nodes - 1);
    const bool pow2 = is_pow2(dx);
    double t = P->t0;
    long long steps = 0;
    int status = 0;
    double* yout = nullptr;
    int rc = DEB_OK;
#define DEB_HEAT_CASE(ID, T) case ID: rc = heat_solve<deb::T>(P, ya.as<double>(), yb.as<double>(), pow2, st, &t, &steps, &status, &yout); break;
    switch (P->method) {
        DEB_HEAT_CASE(DEB_EULER, TabEuler)
        DEB_HEAT_CASE(DEB_MIDPOINT, TabMidpoint)
        DEB_HEAT_CASE(DEB_HEUN, TabHeun)
        DEB_HEAT_CASE(DEB_RALSTON, TabRalston)
        DEB_HEAT_CASE(DEB_SSP_RK3, TabSspRk3)
        DEB_HEAT_CASE(DEB_RK4, TabRk4)
        DEB_HEAT_CASE(DEB_THREE_EIGHTHS, TabThreeEighths)
    }
#undef DEB_HEAT_CASE
    if (rc) return rc;
    DEB_CUDA(cudaMemcpyAsync(P->u_final, yout, bytes, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    DEB_CUDA(cudaStreamSynchronize(st));  // work buffers are freed on return
    if (P->t_final) *P->t_final = t;
    if (P->steps) *P->steps = steps;
    if (P->status) *P->status = status;
    return DEB_OK;
}

extern "C" int deb_heat_rhs(const deb_heat_problem* P, const double* u, double* du) {
    if (!P || !u || !du) return fail(DEB_ERR_BAD_ARG, "NULL pointer");
    if (P->struct_size != sizeof(deb_heat_problem)) return fail(DEB_ERR_BAD_ARG, "struct_size mismatch (ABI version skew)");
    if (P->n_nodes < 2) return fail(DEB_ERR_BAD_ARG, "StructuredGrid requires at least two nodes per axis");
    if (P->lo == P->hi) return fail(DEB_ERR_BAD_ARG, "StructuredGrid endpoints must be distinct");
    if ((P->bc_lower_kind | P->bc_upper_kind) & ~1) return fail(DEB_ERR_BAD_ARG, "boundary kind must be 0 (Dirichlet) or 1 (Neumann)");
    if (int rc = select_device(P->device)) return rc;
    const bool host = (P->memspace == DEB_MEM_HOST);
    cudaStream_t st = host ? (cudaStream_t)0 : (cudaStream_t)P->stream;
    const size_t bytes = sizeof(double) * (size_t)P->n_nodes;
    DevBuf d_u, d_du;
    deb::HeatArgs a;
    memset(&a, 0, sizeof a);
    a.n = P->n_nodes;
    a.dx = (P->hi - P->lo) / (double)(P->n_nodes - 1);
    a.inv_dx = 1.0 / a.dx;
    a.alpha = P->alpha;
    a.bc_lo_kind = P->bc_lower_kind; a.bc_hi_kind = P->bc_upper_kind;
    a.bc_lo_val = P->bc_lower_value; a.bc_hi_val = P->bc_upper_value;
    if (host) {
        DEB_CUDA(d_u.alloc(bytes));
        DEB_CUDA(d_du.alloc(bytes));
        DEB_CUDA(cudaMemcpyAsync(d_u.p, u, bytes, cudaMemcpyHostToDevice, st));
        a.y = d_u.as<double>(); a.out_k = d_du.as<double>();
    } else {
        a.y = u; a.out_k = du;
    }
    const bool pow2 = is_pow2(a.dx);
    if (int rc = heat_launch_stage<deb::TabEuler, 0>(a, pow2, st)) return rc;
    if (host) {
        DEB_CUDA(cudaMemcpyAsync(du, d_du.p, bytes, cudaMemcpyDeviceToHost, st));
        DEB_CUDA(cudaStreamSynchronize(st));
    }
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ statistics
extern "C" int deb_ensemble_stats(const double* y_eval, const int32_t* n_emitted, int64_t n_traj, int32_t n_eval, int32_t dim, double* sums,
                       int64_t* counts, int32_t device, int32_t memspace, void* stream) {
    if (!y_eval || !n_emitted || !sums || !counts) return fail(DEB_ERR_BAD_ARG, "NULL pointer");
    if (n_traj < 0 || n_eval <= 0 || dim <= 0) return fail(DEB_ERR_BAD_ARG, "bad sizes");
    if (int rc = select_device(device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(device, &di)) return rc;
    const bool host = (memspace == DEB_MEM_HOST);
    cudaStream_t st = host ? (cudaStream_t)0 : (cudaStream_t)stream;
    const int ne = n_eval * dim;
    int n_cta = di.sms * 8;
    if ((long long)n_cta > n_traj) n_cta = (int)std::max<int64_t>(1, n_traj);
    DevBuf d_y, d_ne, d_sums, d_counts;
    const double* dy = y_eval;
    const int* dn = n_emitted;
    double* ds = sums;
    long long* dc = (long long*)counts;
    if (host) {
        DEB_CUDA(d_y.alloc(sizeof(double) * (size_t)n_traj * ne));
        DEB_CUDA(d_ne.alloc(sizeof(int) * (size_t)n_traj));
        DEB_CUDA(d_sums.alloc(sizeof(double) * 2 * ne));
        DEB_CUDA(d_counts.alloc(sizeof(long long) * n_eval));
        DEB_CUDA(cudaMemcpyAsync(d_y.p, y_eval, sizeof(double) * (size_t)n_traj * ne, cudaMemcpyHostToDevice, st));
        DEB_CUDA(cudaMemcpyAsync(d_ne.p, n_emitted, sizeof(int) * (size_t)n_traj, cudaMemcpyHostToDevice, st));
        dy = d_y.as<double>(); dn = d_ne.as<int>(); ds = d_sums.as<double>(); dc = d_counts.as<long long>();
    }
    void* scratch = nullptr;
    const size_t pbytes = sizeof(double) * 2 * (size_t)n_cta * ne;
    const size_t cbytes = sizeof(long long) * (size_t)n_cta * n_eval;
    DEB_CUDA(cudaMallocAsync(&scratch, pbytes + cbytes, st));
    struct SmallFree { void* p; cudaStream_t st; ~SmallFree() { if (p) cudaFreeAsync(p, st); } } sf{scratch, st};
    double* partial = (double*)scratch;
    long long* pcount = (long long*)((char*)scratch + pbytes);
    const int pthreads = std::min(512, ((ne + 31) / 32) * 32);
    deb::stats_partial_kernel<<<n_cta, pthreads, 0, st>>>(dy, dn, n_traj, n_eval, dim, partial, pcount);
    DEB_CUDA(cudaGetLastError());
    deb::stats_final_kernel<<<(ne * 32 + 127) / 128, 128, 0, st>>>(partial, pcount, n_cta, n_eval, dim, ds, dc);
    DEB_CUDA(cudaGetLastError());
    if (host) {
        DEB_CUDA(cudaMemcpyAsync(sums, ds, sizeof(double) * 2 * ne, cudaMemcpyDeviceToHost, st));
        DEB_CUDA(cudaMemcpyAsync(counts, dc, sizeof(long long) * n_eval, cudaMemcpyDeviceToHost, st));
        DEB_CUDA(cudaStreamSynchronize(st));
    }
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ memory helpers
extern "C" int deb_malloc(int32_t device, size_t bytes, void** ptr) {
    if (!ptr) return fail(DEB_ERR_BAD_ARG, "NULL ptr");
    if (int rc = select_device(device)) return rc;
    DEB_CUDA(cudaMalloc(ptr, bytes ? bytes : 8));
    return DEB_OK;
}
extern "C" int deb_free(int32_t device, void* ptr) {
    if (int rc = select_device(device)) return rc;
    DEB_CUDA(cudaFree(ptr));
    return DEB_OK;
}
extern "C" int deb_memcpy_h2d(int32_t device, void* dst, const void* src, size_t bytes) {
    if (int rc = select_device(device)) return rc;
    DEB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return DEB_OK;
}
extern "C" int deb_memcpy_d2h(int32_t device, void* dst, const void* src, size_t bytes) {
    if (int rc = select_device(device)) return rc;
    DEB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return DEB_OK;
}
extern "C" int deb_synchronize(int32_t device) {
    if (int rc = select_device(device)) return rc;
    DEB_CUDA(cudaDeviceSynchronize());
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ diagnostics
namespace {
__global__ void pow_kernel(const double* x, double y, long long n, double* out) {
    __shared__ double s_powlog[384];
    __shared__ unsigned long long s_exp[256];
    deb::load_pow_tables(s_powlog, s_exp);
    deb_pow_tables tb;
    tb.powlog = s_powlog;
    tb.exptab = s_exp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = deb_pow_pos(x[i], y, tb);
}

// Register-only stream of independent DP operations: ILP chains per thread, `iters` rounds.
template <bool FMA>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double m, double c) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (FMA) { v[i] = __fma_rn(v[i], m, c); v[i] = __fma_rn(v[i], m, c); }
            else { v[i] = __dmul_rn(v[i], m); v[i] = __dadd_rn(v[i], c); }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 123.456) sink[0] = s;
}
}  // namespace

extern "C" int deb_pow_device(const double* x, double y, int64_t n, double* out, int32_t device) {
    if (!x || !out || n < 0) return fail(DEB_ERR_BAD_ARG, "bad arguments");
    if (n == 0) return DEB_OK;
    if (int rc = select_device(device)) return rc;
    DevBuf dx, dout;
    DEB_CUDA(dx.alloc(sizeof(double) * (size_t)n));
    DEB_CUDA(dout.alloc(sizeof(double) * (size_t)n));
    DEB_CUDA(cudaMemcpy(dx.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    pow_kernel<<<1184, 256>>>(dx.as<double>(), y, n, dout.as<double>());
    DEB_CUDA(cudaGetLastError());
    DEB_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return DEB_OK;
}

extern "C" int deb_fp64_issue_peak(int32_t device, int32_t use_fma, double* dp_inst_per_s, float* ms_out) {
    if (!dp_inst_per_s) return fail(DEB_ERR_BAD_ARG, "NULL output");
    if (int rc = select_device(device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(device, &di)) return rc;
    DevBuf sink;
    DEB_CUDA(sink.alloc(8));
    const int iters = 20000, blocks = di.sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    DEB_CUDA(cudaEventCreate(&e0));
    DEB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        DEB_CUDA(cudaEventRecord(e0));
        if (use_fma) fp64_peak_kernel<true><<<blocks, threads>>>(sink.as<double>(), iters, 0.9999999, 1e-7);
        else fp64_peak_kernel<false><<<blocks, threads>>>(sink.as<double>(), iters, 0.9999999, 1e-7);
        DEB_CUDA(cudaEventRecord(e1));
        DEB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        DEB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double insts = (double)blocks * threads * (double)iters * 16.0;  // 8 chains x 2 DP instructions per round
    *dp_inst_per_s = insts / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return DEB_OK;
}

