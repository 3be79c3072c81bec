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
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
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

// every kernel launch of the library is counted (deb_launch_count, deb_result.gpu_launches)
static std::atomic<long long> g_launches{0};
void deb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t deb_launch_count(void) { return (int64_t)g_launches.load(); }

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
struct DeviceInfo { int sms = 0; cudaMemPool_t pool = nullptr; size_t total_mem = 0; };
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
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) cache[device].total_mem = prop.totalGlobalMem;
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
    deb_count_launch(1);
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
    deb_count_launch(1);
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
    nvrtcResult (*Version)(int*, int*) = nullptr;
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
        DEB_SYM(Version, "nvrtcVersion");
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
    unsigned dyn_smem = 0;  // dynamic shared memory of the instantiation (parked steps of the methods with extra dense stages)
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
std::map<std::tuple<int, int, int, int, int, int>, UserKernel> g_jit_kernels;  // (.., step-size filter)

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

// ---- run-time compilation: NVRTC -> cubin, with a disk cache.
// The cache key is a hash of everything that determines the machine code: the generated translation unit, the kernel
// name expression, the embedded kernel headers, the compiler options and the NVRTC version.  Directory: $DEB_CACHE_DIR
// ("" disables the cache), else <library dir>/jit_cache when writable (so a warmed cache travels with the in-tree build),
// else ~/.cache/deb200.  File = "<lowered kernel name>\n" + cubin bytes.
unsigned long long fnv1a64(const void* data, size_t n, unsigned long long h) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
std::string jit_cache_dir() {
    if (const char* e = getenv("DEB_CACHE_DIR")) return std::string(e);
    Dl_info info;
    if (dladdr((const void*)&deb_fail, &info) && info.dli_fname) {
        std::string lib(info.dli_fname);
        const size_t slash = lib.rfind('/');
        if (slash != std::string::npos) {
            const std::string dir = lib.substr(0, slash) + "/jit_cache";
            mkdir(dir.c_str(), 0755);
            if (access(dir.c_str(), W_OK) == 0) return dir;
        }
    }
    if (const char* h = getenv("HOME")) {
        const std::string base = std::string(h) + "/.cache";
        mkdir(base.c_str(), 0755);
        const std::string dir = base + "/deb200";
        mkdir(dir.c_str(), 0755);
        if (access(dir.c_str(), W_OK) == 0) return dir;
    }
    return std::string();
}

// `user_code`: the translation unit contains caller-supplied text (a compile error is the caller's: DEB_ERR_BAD_ARG)
int nvrtc_compile(const std::string& src, const char* expr, bool user_code, std::vector<char>* cubin, std::string* kernel_name) {
    const NvrtcApi* rt = nvrtc_api();
    if (!rt) return fail(DEB_ERR_UNSUPPORTED, "libnvrtc not found: user-defined systems and per-step recorders need the NVRTC runtime compiler");
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo"};
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
    // ---- disk cache lookup
    std::string cache_file;
    {
        unsigned long long h = 14695981039346656037ull;
        h = fnv1a64(src.data(), src.size(), h);
        h = fnv1a64(expr, strlen(expr), h);
        for (size_t i = 0; i < hdr_text.size(); i++) h = fnv1a64(hdr_text[i], strlen(hdr_text[i]), h);
        for (const char* o : opts) h = fnv1a64(o, strlen(o), h);
        int vmaj = 0, vmin = 0;
        if (rt->Version) rt->Version(&vmaj, &vmin);
        h = fnv1a64(&vmaj, sizeof vmaj, h);
        h = fnv1a64(&vmin, sizeof vmin, h);
        const std::string dir = jit_cache_dir();
        if (!dir.empty()) {
            char name[64];
            snprintf(name, sizeof name, "/deb200-%016llx.cubin", h);
            cache_file = dir + name;
            if (FILE* f = fopen(cache_file.c_str(), "rb")) {
                std::string lowered;
                int ch;
                while ((ch = fgetc(f)) != EOF && ch != '\n') lowered.push_back((char)ch);
                std::vector<char> bin;
                char buf[65536];
                size_t got;
                while ((got = fread(buf, 1, sizeof buf, f)) > 0) bin.insert(bin.end(), buf, buf + got);
                fclose(f);
                if (!lowered.empty() && bin.size() > 64 && memcmp(bin.data(), "\x7f" "ELF", 4) == 0) {
                    *kernel_name = lowered;
                    cubin->swap(bin);
                    return DEB_OK;
                }
            }
        }
    }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = rt->CreateProgram(&prog, src.c_str(), "deb_user_system.cu", (int)hdr_names.size(), hdr_text.data(), hdr_names.data());
    if (r != NVRTC_SUCCESS) return fail(DEB_ERR_CUDA, std::string("nvrtcCreateProgram: ") + rt->GetErrorString(r));
    struct ProgGuard { const NvrtcApi* rt; nvrtcProgram* p; ~ProgGuard() { if (*p) rt->DestroyProgram(p); } } guard{rt, &prog};
    rt->AddNameExpression(prog, expr);
    r = rt->CompileProgram(prog, 4, opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        rt->GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) rt->GetProgramLog(prog, &log[0]);
        return fail(user_code ? DEB_ERR_BAD_ARG : DEB_ERR_CUDA,
                    (user_code ? "the user-defined function did not compile (NVRTC):\n" : "run-time kernel compilation failed (NVRTC):\n") + log);
    }
    const char* lowered = nullptr;
    r = rt->GetLoweredName(prog, expr, &lowered);
    if (r != NVRTC_SUCCESS || !lowered) return fail(DEB_ERR_CUDA, "nvrtcGetLoweredName failed");
    *kernel_name = lowered;
    size_t nbin = 0;
    rt->GetCUBINSize(prog, &nbin);
    cubin->resize(nbin);
    if (rt->GetCUBIN(prog, cubin->data()) != NVRTC_SUCCESS) return fail(DEB_ERR_CUDA, "nvrtcGetCUBIN failed");
    if (const char* dump = getenv("DEB_DUMP_CUBIN")) {  // debugging aid: keep the last run-time compiled cubin for cuobjdump
        if (FILE* f = fopen(dump, "wb")) { fwrite(cubin->data(), 1, cubin->size(), f); fclose(f); }
    }
    if (!cache_file.empty()) {  // write to a temporary name, then rename: concurrent processes never see a partial file
        const std::string tmp = cache_file + ".tmp" + std::to_string((long long)getpid());
        if (FILE* f = fopen(tmp.c_str(), "wb")) {
            fputs(kernel_name->c_str(), f);
            fputc('\n', f);
            fwrite(cubin->data(), 1, cubin->size(), f);
            fclose(f);
            if (rename(tmp.c_str(), cache_file.c_str()) != 0) remove(tmp.c_str());
        }
    }
    return DEB_OK;
}

// (S, I) of the adaptive constructors: dormandprince/mod.rs:45-58, adaptive/mod.rs:47-122
bool method_stage_counts(int method, int* S, int* I) {
    switch (method) {
        case DEB_DOPRI5: *S = 7; *I = 7; return true;
        case DEB_DOP853: *S = 12; *I = 16; return true;
        case DEB_RKF45: case DEB_CASH_KARP: *S = 6; *I = 6; return true;
        case DEB_RKV655E: *S = 9; *I = 10; return true;
        case DEB_RKV656E: *S = 9; *I = 12; return true;
        case DEB_RKV766E: *S = 10; *I = 13; return true;
        case DEB_RKV767E: *S = 10; *I = 16; return true;
        case DEB_RKV877E: *S = 13; *I = 17; return true;
        case DEB_RKV878E: *S = 13; *I = 21; return true;
        case DEB_RKV988E: *S = 16; *I = 21; return true;
        case DEB_RKV989E: *S = 16; *I = 26; return true;
    }
    return false;
}
// dynamic shared memory of a run-time instantiation dp_ensemble_kernel<.., 128, .., REC>: the formula of dp_dynamic_smem_bytes
unsigned jit_dynamic_smem(int method, int dim, bool rec) {
    int S = 0, I = 0;
    if (!method_stage_counts(method, &S, &I) || rec || I <= S) return 0;
    const long long bytes = (2ll + (long long)dim * (S + 3)) * 128 * 8;
    return bytes <= 64 * 1024 ? (unsigned)bytes : 0u;
}

// Compile the ensemble kernel for (system, method, recorder kind) to a cubin (no device needed).  `us` = the user
// system, or null for a built-in one.
int compile_kernel_cubin(const UserSystem* us, int system, int method, int rec_mode, int event, bool filter, std::vector<char>* cubin,
                         std::string* kernel_name, bool* is_adaptive) {
    const bool rec = rec_mode >= 0;  // -1: the row-plan kernels (t_eval / even(dt) without event)
    int sdim = 0, snp = 0;
    const char* sys_name = us ? "deb::UserSys" : builtin_system_name(system, &sdim, &snp);
    if (!sys_name) return fail(DEB_ERR_BAD_ARG, "unknown system id");
    if (us) { sdim = us->dim; snp = us->np; }
    bool adaptive = false;
    const char* tab = method_tab_name(method, &adaptive);
    if (!tab) return fail(DEB_ERR_UNSUPPORTED, "unknown or unsupported method id");
    // occupancy hint: stage vectors live in registers, wider systems get the whole register file of fewer CTAs
    int min_blocks = 1;
    if (adaptive) {
        if (method == DEB_DOP853) min_blocks = sdim <= 2 ? 4 : sdim == 3 ? 3 : sdim <= 6 ? 2 : 1;
        else if (method >= DEB_RKV655E && method <= DEB_RKV989E) min_blocks = sdim <= 2 ? 4 : sdim == 3 ? 3 : sdim <= 5 ? 2 : 1;
        else min_blocks = sdim <= 3 ? 5 : sdim <= 6 ? 3 : sdim <= 10 ? 2 : 1;
        // a recorder that interpolates keeps the dense output of a step live
        if (rec && min_blocks > 1 && !(rec_mode == DEB_SOLOUT_DEFAULT && event == DEB_EVENT_NONE)) min_blocks -= 1;
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
    if (adaptive) snprintf(expr, sizeof expr, "deb::dp_ensemble_kernel<%s, %s, 128, %d, false, %s, %s, %s>", sys_name, tab, min_blocks, rec ? "true" : "false", evt.c_str(), filter ? "true" : "false");
    else snprintf(expr, sizeof expr, "deb::fixed_ensemble_kernel<%s, %s, 128, %s, %s>", sys_name, tab, rec ? "true" : "false", evt.c_str());
    std::string src;
    if (rec) src += "#define DEB_JIT_REC_MODE " + std::to_string(rec_mode) + "\n";  // step_recorder.cuh: one kernel per recorder
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
    *is_adaptive = adaptive;
    return nvrtc_compile(src, expr, us != nullptr || ue != nullptr, cubin, kernel_name);
}

// Compile and load (once per device / system / method / recorder kind) a run-time kernel.  Caller holds g_user_mu.
int jit_kernel(const UserSystem* us, int device, int system, int method, int rec_mode, int event, bool filter, UserKernel** out) {
    const bool rec = rec_mode >= 0;
    const auto key = std::make_tuple(device, system, method, rec_mode, event, filter ? 1 : 0);
    auto it = g_jit_kernels.find(key);
    if (it != g_jit_kernels.end()) { *out = &it->second; return DEB_OK; }
    std::vector<char> cubin;
    std::string lowered;
    UserKernel uk;
    if (int rc = compile_kernel_cubin(us, system, method, rec_mode, event, filter, &cubin, &lowered, &uk.adaptive)) return rc;
    DEB_CUDA(cudaLibraryLoadData(&uk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    DEB_CUDA(cudaLibraryGetKernel(&uk.kernel, uk.lib, lowered.c_str()));
    if (uk.adaptive) {
        int sdim = 0, snp = 0;
        if (us) sdim = us->dim; else builtin_system_name(system, &sdim, &snp);
        uk.dyn_smem = jit_dynamic_smem(method, sdim, rec);
        if (uk.dyn_smem > 0) DEB_CUDA(cudaFuncSetAttribute((const void*)uk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)uk.dyn_smem));
    }
    auto ins = g_jit_kernels.emplace(key, uk);
    *out = &ins.first->second;
    return DEB_OK;
}

int launch_user(const UserKernel& uk, const deb::OdeKernelArgs& a, int sms, cudaStream_t st) {
    long long blocks;
    // parked steps live in dynamic shared memory; a launch that emits no rows does not touch it (ode_dispatch.cuh)
    const unsigned dyn = (a.n_rows > 0 && a.y_eval != nullptr) ? uk.dyn_smem : 0u;
    if (uk.adaptive) {
        int per_sm = 0;
        DEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)uk.kernel, uk.block, dyn));
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
    DEB_CUDA(cudaLaunchKernel((const void*)uk.kernel, dim3((unsigned)blocks), dim3(uk.block), args, dyn, st));
    deb_count_launch(1);
    return DEB_OK;
}

// ------------------------------------------------------------------------------------------------ user SDEs (NVRTC)
// `impl SDE for S { fn drift; fn diffusion; fn noise }` (/root/reference/src/sde/sde.rs:16-67) as CUDA C++ text, compiled into
// the same sde_ensemble_kernel template as the built-in SDEs.
struct UserSde {
    int dim = 0, np = 0;
    std::string drift, diffusion, noise;
};
std::vector<std::unique_ptr<UserSde>> g_user_sdes;  // ids USER_SYSTEM_BASE + index; guarded by g_user_mu
std::map<std::tuple<int, int, int>, UserKernel> g_jit_sde_kernels;  // (device, sde id, method)

const char* sde_tab_name(int method, bool* milstein) {
    *milstein = false;
    switch (method) {
        case DEB_EULER: return "deb::TabEuler";
        case DEB_MIDPOINT: return "deb::TabMidpoint";
        case DEB_HEUN: return "deb::TabHeun";
        case DEB_RALSTON: return "deb::TabRalston";
        case DEB_SSP_RK3: return "deb::TabSspRk3";
        case DEB_RK4: return "deb::TabRk4";
        case DEB_THREE_EIGHTHS: return "deb::TabThreeEighths";
        case DEB_MILSTEIN: *milstein = true; return "deb::TabEuler";
    }
    return nullptr;
}

int compile_sde_cubin(const UserSde& us, int method, std::vector<char>* cubin, std::string* kernel_name) {
    bool milstein = false;
    const char* tab = sde_tab_name(method, &milstein);
    if (!tab) return fail(DEB_ERR_UNSUPPORTED, "SDE ensembles take a fixed-step method id or DEB_MILSTEIN");
    char expr[256];
    snprintf(expr, sizeof expr, "deb::sde_ensemble_kernel<deb::UserSdeSys, %s, 256, %s>", tab, milstein ? "true" : "false");
    std::string src = "#include \"sde_ensemble.cuh\"\nnamespace deb {\nstruct UserSdeSys {\n";
    src += "    static constexpr int DIM = " + std::to_string(us.dim) + ", NP = " + std::to_string(us.np) + ", NPX = 0;\n";
    src += "    __device__ __forceinline__ static void prepare(double*) {}\n";
    src += "    __device__ __forceinline__ static void drift(double t, const double* y, double* dydt, const double* p) {\n        (void)t; (void)y; (void)p;\n";
    src += us.drift;
    src += "\n    }\n    __device__ __forceinline__ static void diffusion(double t, const double* y, double* g, const double* p) {\n        (void)t; (void)y; (void)p;\n";
    src += us.diffusion;
    src += "\n    }\n    __device__ __forceinline__ static void mix(double* dw, const double* p) {\n        (void)dw; (void)p;\n";
    src += us.noise;
    src += "\n    }\n};\n}  // namespace deb\n";
    return nvrtc_compile(src, expr, true, cubin, kernel_name);
}

int jit_sde_kernel(const UserSde& us, int device, int system, int method, UserKernel** out) {
    const auto key = std::make_tuple(device, system, method);
    auto it = g_jit_sde_kernels.find(key);
    if (it != g_jit_sde_kernels.end()) { *out = &it->second; return DEB_OK; }
    std::vector<char> cubin;
    std::string lowered;
    UserKernel uk;
    uk.block = 256;
    if (int rc = compile_sde_cubin(us, method, &cubin, &lowered)) return rc;
    DEB_CUDA(cudaLibraryLoadData(&uk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    DEB_CUDA(cudaLibraryGetKernel(&uk.kernel, uk.lib, lowered.c_str()));
    auto ins = g_jit_sde_kernels.emplace(key, uk);
    *out = &ins.first->second;
    return DEB_OK;
}

int launch_user_sde(const UserKernel& uk, const deb::SdeKernelArgs& a, int sms, cudaStream_t st) {
    long long blocks = (a.n_traj + uk.block - 1) / uk.block;
    const long long cap = (long long)sms * 8 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    void* args[] = {(void*)&a};
    DEB_CUDA(cudaLaunchKernel((const void*)uk.kernel, dim3((unsigned)blocks), dim3(uk.block), args, 0, st));
    deb_count_launch(1);
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

// ------------------------------------------------------------------------------------------------ fixed-step schedule
// For a fixed-step method the loop control of solve_ode / solve_sde (validate_step_size_parameters utils.rs:60-157 with the
// method's h_min / h_max; the clip at tf and the end test, ode/solve_ivp.rs:193-209, :263, sde/solve_ivp.rs:211-227, :263;
// the max_steps test, fixed/ordinary.rs:66-75) does not depend on the trajectory: it runs here, once.  Every step has size
// h0 until the clip `h = tf - t` fires; it can fire again (t + (tf - t) may miss tf by one ulp, more than 10 eps when
// |tf| > 10: the reference then takes one more, tiny, step), so the schedule is "h0, ..., h0, tail[0..n_tail)".
struct FixedSchedule {
    long long n_steps = 0;
    int n_tail = 0;
    double h_tail[deb::DEB_FX_MAX_TAIL] = {0, 0, 0, 0};
    int status = DEB_STATUS_COMPLETE;
};
template <class OnStep>
int plan_fixed_schedule(double t0, double tf, double h0_opt, double h_min, double h_max, long long max_steps, FixedSchedule* fs,
                        OnStep on_step) {
    *fs = FixedSchedule();
    const double dd = tf - t0;
    const double dir = (dd != dd) ? dd : copysign(1.0, dd);
    double h = h0_opt;
    if (h == 0.0) h = fabs(tf - t0) / 100.0;  // fixed/ordinary.rs:23-28
    const double h0 = h;
    const double sgh = (h != h) ? h : copysign(1.0, h);
    const bool ok = (tf != t0) && (dir == 1.0 || dir == -1.0) && sgh == dir && !(h_min < 0.0) && !(h_max < 0.0) && !(h_min > h_max) &&
                    !(fabs(h) < h_min) && !(fabs(h) > h_max) && !(fabs(h) > fabs(tf - t0)) && h != 0.0;
    fs->h_tail[0] = h;
    if (!ok) {
        fs->status = DEB_STATUS_BAD_INPUT;
        return DEB_OK;
    }
    const double eps10 = 2.220446049250313e-16 * 10.0;
    double t = t0;
    long long steps = 0, first_clipped = -1;
    double recent[deb::DEB_FX_MAX_TAIL] = {0, 0, 0, 0};  // sizes of the last DEB_FX_MAX_TAIL steps (ring)
    for (;;) {
        if ((t + h - tf) * dir > 0.0) {
            const double h_new = tf - t;
            if (fabs(h_new) < eps10) break;
            h = h_new;
        }
        if (steps >= max_steps) { fs->status = DEB_STATUS_MAX_STEPS; break; }
        const double t_new = t + h;
        on_step(steps, t, t_new);
        if (first_clipped < 0 && h != h0) first_clipped = steps;
        recent[steps % deb::DEB_FX_MAX_TAIL] = h;
        steps += 1;
        t = t_new;
        if (fabs(tf - t) <= eps10) break;
    }
    fs->n_steps = steps;
    if (steps > 0) {
        long long n_tail = (first_clipped < 0) ? 1 : steps - first_clipped;  // the last step is always carried in the tail
        if (n_tail > deb::DEB_FX_MAX_TAIL)
            return fail(DEB_ERR_UNSUPPORTED, "fixed-step schedule: the clip at tf fired more than DEB_FX_MAX_TAIL times");
        fs->n_tail = (int)n_tail;
        for (long long q = 0; q < n_tail; q++) fs->h_tail[q] = recent[(steps - n_tail + q) % deb::DEB_FX_MAX_TAIL];
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

// ================================================================================================ deb_solve_ode
namespace {

// ---- struct_size handling: fields are only ever appended.  A caller built against an older header passes a smaller
//      struct_size; the missing tail reads as zeros (= the defaults).  A larger struct_size (newer header) is rejected.
template <class T>
int import_struct(const T* user, T* local, size_t min_size, const char* what) {
    if (!user) return fail(DEB_ERR_BAD_ARG, std::string("NULL ") + what);
    const size_t sz = user->struct_size;
    if (sz < min_size || sz > sizeof(T))
        return fail(DEB_ERR_BAD_ARG, std::string("struct_size mismatch (ABI version skew): ") + what + " has struct_size " + std::to_string(sz) +
                                         ", this library accepts " + std::to_string(min_size) + ".." + std::to_string(sizeof(T)));
    memset((void*)local, 0, sizeof(T));
    memcpy((void*)local, (const void*)user, sz);
    local->struct_size = sizeof(T);
    return DEB_OK;
}
// deb_result is in/out: work on a full-size local copy, write the caller's prefix back on every exit
struct ResultIO {
    deb_result* user = nullptr;
    deb_result local;
    size_t sz = 0;
    int open(deb_result* u) {
        if (int rc = import_struct(u, &local, offsetof(deb_result, t_out), "deb_result")) return rc;
        user = u;
        sz = u->struct_size;
        return DEB_OK;
    }
    ~ResultIO() {
        if (user) {
            local.struct_size = sz;
            memcpy((void*)user, (const void*)&local, sz);
        }
    }
};

// ---- block-cyclic distribution of an ensemble over the devices of one call.  Blocks of B = 2^shift consecutive
//      trajectories; global block b belongs to device b mod G and is that device's local block b div G.  Only the last
//      global block can be partial, and it is the last local block of its device.  G = 1: the identity.
struct ShardMap {
    long long n_total = 0;
    int shift = 12;
    int G = 1, g = 0;
    long long B() const { return 1ll << shift; }
    long long blocks_total() const { return (n_total + B() - 1) >> shift; }
    long long local_blocks() const { const long long nb = blocks_total(); return nb > g ? (nb - g + G - 1) / G : 0; }
    long long global_block(long long lb) const { return lb * G + g; }
    long long block_size(long long lb) const { return std::min(B(), n_total - (global_block(lb) << shift)); }
    long long local_count() const {
        const long long nlb = local_blocks();
        return nlb == 0 ? 0 : (nlb - 1) * B() + block_size(nlb - 1);
    }
    // trajectories in local blocks [lb0, lb1)
    long long count(long long lb0, long long lb1) const { return lb1 <= lb0 ? 0 : (lb1 - 1 - lb0) * B() + block_size(lb1 - 1); }
};

// Copy local blocks [lb0, lb1) of a per-trajectory array (bpt bytes per trajectory) between the caller's HOST array
// (indexed by global trajectory) and a device chunk buffer (indexed by local trajectory - lbc*B).
int copy_blocks(const ShardMap& M, char* host_base, char* dev_base, size_t bpt, long long lbc, long long lb0, long long lb1, bool to_host,
                cudaStream_t st) {
    if (lb1 <= lb0 || bpt == 0) return DEB_OK;
    const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    const size_t width = (size_t)M.B() * bpt;
    auto hptr = [&](long long lb) { return host_base + (size_t)(M.global_block(lb) << M.shift) * bpt; };
    auto dptr = [&](long long lb) { return dev_base + (size_t)((lb - lbc) << M.shift) * bpt; };
    long long full_end = lb1;
    if (M.block_size(lb1 - 1) != M.B()) full_end = lb1 - 1;  // the partial block goes separately
    if (full_end > lb0) {
        const long long nf = full_end - lb0;
        if (M.G == 1) {
            if (to_host) DEB_CUDA(cudaMemcpyAsync(hptr(lb0), dptr(lb0), width * (size_t)nf, kind, st));
            else DEB_CUDA(cudaMemcpyAsync(dptr(lb0), hptr(lb0), width * (size_t)nf, kind, st));
        } else if (width * (size_t)M.G <= 0x7fffffffull && nf > 1) {
            if (to_host) DEB_CUDA(cudaMemcpy2DAsync(hptr(lb0), width * M.G, dptr(lb0), width, width, (size_t)nf, kind, st));
            else DEB_CUDA(cudaMemcpy2DAsync(dptr(lb0), width, hptr(lb0), width * M.G, width, (size_t)nf, kind, st));
        } else {
            for (long long lb = lb0; lb < full_end; lb++) {
                if (to_host) DEB_CUDA(cudaMemcpyAsync(hptr(lb), dptr(lb), width, kind, st));
                else DEB_CUDA(cudaMemcpyAsync(dptr(lb), hptr(lb), width, kind, st));
            }
        }
    }
    if (full_end < lb1) {
        const size_t bytes = (size_t)M.block_size(lb1 - 1) * bpt;
        if (to_host) DEB_CUDA(cudaMemcpyAsync(hptr(lb1 - 1), dptr(lb1 - 1), bytes, kind, st));
        else DEB_CUDA(cudaMemcpyAsync(dptr(lb1 - 1), hptr(lb1 - 1), bytes, kind, st));
    }
    return DEB_OK;
}

// ---- per-device resources of the HOST pipeline, kept between calls: streams, events, the pinned flag array
struct SlotRes {
    cudaStream_t st = nullptr;  // H2D + kernel
    cudaStream_t cp = nullptr;  // streamed D2H
    cudaStream_t cpx[3] = {nullptr, nullptr, nullptr};  // optional extra D2H streams (DEB_COPY_STREAMS > 1): more copy engines in flight
    cudaEvent_t cpx_done[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t k0 = nullptr, k1 = nullptr, cp_done = nullptr;
    int* flags = nullptr;       // pinned, mapped: completion flags of the watermark blocks
    size_t n_flags = 0;
};
std::mutex g_slot_mu;
std::map<int, std::vector<SlotRes*>> g_slot_free;  // device -> idle resources

int acquire_slot(int device, size_t n_flags, SlotRes** out) {
    SlotRes* r = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        auto& v = g_slot_free[device];
        if (!v.empty()) { r = v.back(); v.pop_back(); }
    }
    if (!r) {
        r = new SlotRes;
        DEB_CUDA(cudaStreamCreateWithFlags(&r->st, cudaStreamNonBlocking));
        DEB_CUDA(cudaStreamCreateWithFlags(&r->cp, cudaStreamNonBlocking));
        DEB_CUDA(cudaEventCreate(&r->k0));
        DEB_CUDA(cudaEventCreate(&r->k1));
        DEB_CUDA(cudaEventCreateWithFlags(&r->cp_done, cudaEventDisableTiming));
        for (int q = 0; q < 3; q++) {
            DEB_CUDA(cudaStreamCreateWithFlags(&r->cpx[q], cudaStreamNonBlocking));
            DEB_CUDA(cudaEventCreateWithFlags(&r->cpx_done[q], cudaEventDisableTiming));
        }
    }
    if (r->n_flags < n_flags) {
        if (r->flags) cudaFreeHost(r->flags);
        r->flags = nullptr;
        r->n_flags = 0;
        const size_t want = std::max<size_t>(n_flags, 4096);
        DEB_CUDA(cudaHostAlloc((void**)&r->flags, want * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        r->n_flags = want;
    }
    *out = r;
    return DEB_OK;
}
void release_slot(int device, SlotRes* r) {
    if (!r) return;
    std::lock_guard<std::mutex> lk(g_slot_mu);
    g_slot_free[device].push_back(r);
}

// ---- y_eval[i][e] -> out[e][i_off + i] (DEB_LAYOUT_ROW_MAJOR), 32x32 tiles through shared memory; HBM bound
__global__ void __launch_bounds__(256) transpose_rows_kernel(const double* __restrict__ src, long long n, int ne, double* __restrict__ dst,
                                                             long long dst_pitch, long long dst_off) {
    __shared__ double tile[32][33];
    const long long i0 = (long long)blockIdx.x * 32;
    const int e0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r;
        const int e = e0 + tx;
        if (i < n && e < ne) tile[r][tx] = src[i * ne + e];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int e = e0 + r;
        const long long i = i0 + tx;
        if (i < n && e < ne) dst[(long long)e * dst_pitch + dst_off + i] = tile[tx][r];
    }
}
int launch_transpose(const double* src, long long n, int ne, double* dst, long long dst_pitch, long long dst_off, cudaStream_t st) {
    if (n <= 0 || ne <= 0) return DEB_OK;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((ne + 31) / 32));
    transpose_rows_kernel<<<grid, 256, 0, st>>>(src, n, ne, dst, dst_pitch, dst_off);
    DEB_CUDA(cudaGetLastError());
    deb_count_launch(1);
    return DEB_OK;
}

// ---- NCCL, loaded at first use (only calls with n_devices >= 2 and statistics need it)
struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int dtype, int op, void* comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
const NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("DEB_NO_NCCL")) return;  // test knob: take the host-sum path of deb_solve_ode
        // a copy that is already in the process (e.g. the one PyTorch bundles) wins over the system library
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
        for (const char* n : names) {
            if (api.handle) break;
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        }
        if (!api.handle) return;
        *(void**)(&api.CommInitAll) = dlsym(api.handle, "ncclCommInitAll");
        *(void**)(&api.GroupStart) = dlsym(api.handle, "ncclGroupStart");
        *(void**)(&api.GroupEnd) = dlsym(api.handle, "ncclGroupEnd");
        *(void**)(&api.AllReduce) = dlsym(api.handle, "ncclAllReduce");
        *(void**)(&api.GetErrorString) = dlsym(api.handle, "ncclGetErrorString");
        if (!api.CommInitAll || !api.GroupStart || !api.GroupEnd || !api.AllReduce) api.handle = nullptr;
    });
    return api.handle ? &api : nullptr;
}
std::mutex g_nccl_mu;
std::map<std::vector<int>, std::vector<void*>> g_nccl_comms;  // device list -> communicators (kept for the process lifetime)

// Sum `count` doubles (bufs[g] on devices[g], in place) across the devices of a call.  ncclDouble = 8, ncclInt64 = 4, ncclSum = 0.
int allreduce_across_devices(const std::vector<int>& devices, const std::vector<void*>& bufs, size_t count, int nccl_dtype,
                             const std::vector<cudaStream_t>& streams) {
    const NcclApi* nc = nccl_api();
    if (!nc) return fail(DEB_ERR_UNSUPPORTED, "libnccl not found: ensemble statistics across several devices need NCCL");
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    auto it = g_nccl_comms.find(devices);
    if (it == g_nccl_comms.end()) {
        std::vector<void*> comms(devices.size(), nullptr);
        const int rc = nc->CommInitAll(comms.data(), (int)devices.size(), devices.data());
        if (rc != 0) return fail(DEB_ERR_CUDA, std::string("ncclCommInitAll: ") + (nc->GetErrorString ? nc->GetErrorString(rc) : "error"));
        it = g_nccl_comms.emplace(devices, comms).first;
    }
    int rc = nc->GroupStart();
    for (size_t g = 0; g < devices.size() && rc == 0; g++)
        rc = nc->AllReduce(bufs[g], bufs[g], count, nccl_dtype, 0, it->second[g], streams[g]);
    const int rc2 = nc->GroupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) return fail(DEB_ERR_CUDA, std::string("ncclAllReduce: ") + (nc->GetErrorString ? nc->GetErrorString(rc) : "error"));
    return DEB_OK;
}

int launch_stats(const double* dy, const int* dn, long long n_traj, int n_eval, int dim, double* ds, long long* dc, bool accumulate, int sms,
                 cudaStream_t st) {
    const int ne = n_eval * dim;
    int n_cta = sms * 8;
    if ((long long)n_cta > n_traj) n_cta = (int)std::max<long long>(1, n_traj);
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
    deb::stats_final_kernel<<<(ne * 32 + 127) / 128, 128, 0, st>>>(partial, pcount, n_cta, n_eval, dim, ds, dc, accumulate ? 1 : 0);
    DEB_CUDA(cudaGetLastError());
    deb_count_launch(2);
    return DEB_OK;
}

// ---- one validated deb_solve_ode call
struct OdeCall {
    deb_ode_problem P;      // full-size local copy
    deb_result* R = nullptr;
    int dim = 0, np = 0, row_cap = 0;
    bool even = false, per_step = false, has_event = false, rec = false, jit = false, per_traj_params = false;
    UserSystem* user = nullptr;
    ode_launch_fn builtin = nullptr;
    TEvalPlan plan;
    deb::OdeKernelArgs a;   // everything that does not depend on the device / chunk
    std::vector<int> devices;
    bool want_stats = false;
    std::chrono::steady_clock::time_point wall0;  // start of the deb_solve_ode call (DEB_DEBUG_TIMING)
};
typedef std::function<int(const deb::OdeKernelArgs&, int, cudaStream_t)> LaunchFn;

// kernel launcher for `device` (compiles and loads a run-time kernel on first use)
int bind_launcher(const OdeCall& C, int device, LaunchFn* launch) {
    if (!C.jit) {
        *launch = C.builtin;
        return DEB_OK;
    }
    std::lock_guard<std::mutex> lk(g_user_mu);
    UserKernel* uk = nullptr;
    if (int rc = jit_kernel(C.user, device, C.P.system, C.P.method, C.rec ? C.P.solout : -1, C.P.event, C.a.filter_mask != 0ull, &uk)) return rc;
    const UserKernel ukc = *uk;
    *launch = [ukc](const deb::OdeKernelArgs& ka, int sms, cudaStream_t s2) { return launch_user(ukc, ka, sms, s2); };
    return DEB_OK;
}

struct ShardOut {
    int rc = DEB_OK;
    std::string err;
    float kernel_ms = 0.f;
    PoolBuf stats;          // [2*ne doubles][n_eval int64]: this device's sums, kept until the reduction across devices
    cudaStream_t stats_stream = nullptr;
    SlotRes* res[2] = {nullptr, nullptr};
    int device = 0;
};

// The HOST-memspace pipeline of one device.  The device's share of the ensemble (ShardMap) is cut into chunks that fit
// the device-memory budget (normally ONE chunk: 10 M Lorenz trajectories with 100 rows are 24.5 GB); every chunk is ONE
// launch of the persistent kernel over device-resident buffers.  While it runs, the kernel publishes a completion
// watermark per block of 4096 trajectories (wm_publish); this thread polls the flags and copies finished blocks to the
// caller's arrays on a second stream, so that when the kernel ends only the last few blocks are still to be copied.
// Two slots alternate when there are several chunks: the next chunk's kernel is queued before the current one is drained,
// so its CTAs fill the SMs as the current kernel's tail retires.
int run_shard(const OdeCall& C, const ShardMap& M, int device, ShardOut* out) {
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    out->device = device;
    if (getenv("DEB_DEBUG_TIMING"))
        fprintf(stderr, "[deb timing] device %d: shard starts %.2f ms into the call\n", device,
                std::chrono::duration<double, std::milli>(t_start - C.wall0).count());
    if (int rc = select_device(device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(device, &di)) return rc;
    LaunchFn launch;
    if (int rc = bind_launcher(C, device, &launch)) return rc;
    const deb_ode_problem* P = &C.P;
    const deb_result* R = C.R;
    const int dim = C.dim, np = C.np, n_eval = C.row_cap;
    const long long nlb = M.local_blocks();
    const bool row_major = (P->layout == DEB_LAYOUT_ROW_MAJOR) && R->y_eval;
    const size_t ne = (size_t)n_eval * dim;

    // ---- chunking by device-memory budget
    size_t bpt = sizeof(double) * dim + 8 + sizeof(double) * dim + 5 * sizeof(int);  // y0 + finals
    if (C.per_traj_params) bpt += sizeof(double) * np;
    if (R->y_eval || C.want_stats) bpt += sizeof(double) * ne * (row_major ? 2 : 1);
    if (R->t_out) bpt += sizeof(double) * n_eval;
    // device-memory budget of the resident buffers: 40 % of what is free (plus what the library's pool already holds).  The query takes
    // a process-wide driver lock, so a share that is small against the device (every multi-GPU split of C2) does not ask.
    size_t budget = (size_t)(di.total_mem * 0.40);
    if ((double)M.local_count() * (double)bpt > (double)di.total_mem * 0.15) {
        size_t free_b = 0, total_b = 0;
        DEB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        unsigned long long pooled = 0;
        cudaMemPoolGetAttribute(di.pool, cudaMemPoolAttrReservedMemCurrent, &pooled);  // cached staging buffers are reusable
        budget = (size_t)((free_b + pooled) * 0.40);
    }
    if (const char* e = getenv("DEB_HOST_CHUNK_BYTES")) {
        const long long v = atoll(e);
        if (v > 0) budget = (size_t)v;
    }
    long long chunk_blocks = std::max<long long>(1, (long long)(budget / bpt) >> M.shift);
    if (const char* e = getenv("DEB_HOST_CHUNK")) {  // test knob: chunk size in trajectories (rounded up to whole blocks)
        const long long v = atoll(e);
        if (v > 0) chunk_blocks = std::max<long long>(1, (v + M.B() - 1) >> M.shift);
    }
    if (chunk_blocks >= nlb) chunk_blocks = std::max<long long>(nlb, 1);
    else chunk_blocks = std::max<long long>(1, std::min(chunk_blocks, (long long)(budget / 2 / bpt) >> M.shift));  // two slots resident
    const long long n_chunks = nlb == 0 ? 0 : (nlb + chunk_blocks - 1) / chunk_blocks;
    const int n_slots = n_chunks > 1 ? 2 : 1;
    const long long chunk_traj = chunk_blocks << M.shift;
    const size_t rows_bytes = sizeof(double) * C.plan.rows.size();

    struct Slot {
        SlotRes* r = nullptr;
        PoolBuf y0, params, small, y_eval, y_eval_t, n_emitted, t_final, y_final, status, accepted, rejected, evals, t_out;
        long long lb0 = 0, lb1 = 0;  // local blocks of the chunk it currently holds
        bool busy = false;
        void release_buffers() {
            for (PoolBuf* b : {&y0, &params, &small, &y_eval, &y_eval_t, &n_emitted, &t_final, &y_final, &status, &accepted, &rejected, &evals, &t_out}) b->release();
        }
    } slot[2];
    struct SlotGuard {
        Slot* s; int n; int device; ShardOut* out;
        ~SlotGuard() {
            for (int i = 0; i < n; i++) {
                if (s[i].r) {
                    cudaStreamSynchronize(s[i].r->st);
                    cudaStreamSynchronize(s[i].r->cp);
                    for (int q = 0; q < 3; q++) cudaStreamSynchronize(s[i].r->cpx[q]);
                }
                s[i].release_buffers();
                if (s[i].r) release_slot(device, s[i].r);
            }
        }
    } guard{slot, n_slots, device, out};
    const size_t n_wm = (size_t)chunk_blocks;
    for (int s = 0; s < n_slots && n_chunks > 0; s++) {
        Slot& S = slot[s];
        if (int rc = acquire_slot(device, n_wm, &S.r)) return rc;
        cudaStream_t st = S.r->st;
        DEB_CUDA(S.y0.alloc(sizeof(double) * (size_t)chunk_traj * dim, st));
        if (C.per_traj_params) DEB_CUDA(S.params.alloc(sizeof(double) * (size_t)chunk_traj * np, st));
        // [queue counter 8 B][rows][watermark counters]
        DEB_CUDA(S.small.alloc(8 + rows_bytes + sizeof(int) * n_wm, st));
        if (rows_bytes) DEB_CUDA(cudaMemcpyAsync((char*)S.small.p + 8, C.plan.rows.data(), rows_bytes, cudaMemcpyHostToDevice, st));
        if (R->y_eval || C.want_stats) DEB_CUDA(S.y_eval.alloc(sizeof(double) * (size_t)chunk_traj * ne, st));
        if (row_major) DEB_CUDA(S.y_eval_t.alloc(sizeof(double) * (size_t)chunk_traj * ne, st));
        if (R->n_emitted || C.want_stats) DEB_CUDA(S.n_emitted.alloc(sizeof(int) * (size_t)chunk_traj, st));
        if (R->t_final) DEB_CUDA(S.t_final.alloc(sizeof(double) * (size_t)chunk_traj, st));
        if (R->y_final) DEB_CUDA(S.y_final.alloc(sizeof(double) * (size_t)chunk_traj * dim, st));
        if (R->status) DEB_CUDA(S.status.alloc(sizeof(int) * (size_t)chunk_traj, st));
        if (R->accepted) DEB_CUDA(S.accepted.alloc(sizeof(int) * (size_t)chunk_traj, st));
        if (R->rejected) DEB_CUDA(S.rejected.alloc(sizeof(int) * (size_t)chunk_traj, st));
        if (R->evals) DEB_CUDA(S.evals.alloc(sizeof(int) * (size_t)chunk_traj, st));
        if (R->t_out) DEB_CUDA(S.t_out.alloc(sizeof(double) * (size_t)chunk_traj * n_eval, st));
    }
    // shared parameter set for run-time kernels (SHARED_P = false: read through a pointer with stride 0)
    PoolBuf shared_params;
    if (np > 0 && P->params_shared && (C.jit || np > 8) && n_chunks > 0) {
        DEB_CUDA(shared_params.alloc(sizeof(double) * np, slot[0].r->st));
        DEB_CUDA(cudaMemcpyAsync(shared_params.p, P->params, sizeof(double) * np, cudaMemcpyHostToDevice, slot[0].r->st));
        DEB_CUDA(cudaStreamSynchronize(slot[0].r->st));  // the other slot's stream reads it too
    }
    if (C.want_stats && n_chunks > 0) {
        out->stats_stream = slot[0].r->st;
        DEB_CUDA(out->stats.alloc(sizeof(double) * 2 * ne + sizeof(long long) * n_eval, out->stats_stream));
        DEB_CUDA(cudaMemsetAsync(out->stats.p, 0, sizeof(double) * 2 * ne + sizeof(long long) * n_eval, out->stats_stream));
    }

    // D2H of local blocks [b0, b1) of the chunk in slot S, on its copy stream
    int n_copy_streams = 1;  // D2H streams per device; the row copies of a batch are dealt over them
    if (const char* e = getenv("DEB_COPY_STREAMS")) n_copy_streams = std::max(1, std::min(4, atoi(e)));
    auto copy_out = [&](Slot& S, long long b0, long long b1, bool with_rows) -> int {
        cudaStream_t cp = S.r->cp;
#define DEB_OUT(field, T, per) \
    if (R->field) { if (int rc = copy_blocks(M, (char*)R->field, (char*)S.field.p, sizeof(T) * (size_t)(per), S.lb0, b0, b1, true, cp)) return rc; }
        if (with_rows && !row_major && R->y_eval && n_copy_streams > 1 && b1 - b0 >= n_copy_streams) {
            const long long per_stream = (b1 - b0 + n_copy_streams - 1) / n_copy_streams;
            for (int q = 0; q < n_copy_streams; q++) {
                const long long q0 = b0 + q * per_stream, q1 = std::min(b1, q0 + per_stream);
                if (q1 <= q0) break;
                if (int rc = copy_blocks(M, (char*)R->y_eval, (char*)S.y_eval.p, sizeof(double) * ne, S.lb0, q0, q1, true, q == 0 ? cp : S.r->cpx[q - 1])) return rc;
            }
        } else if (with_rows && !row_major) DEB_OUT(y_eval, double, ne)
        DEB_OUT(n_emitted, int, 1)
        DEB_OUT(t_final, double, 1)
        DEB_OUT(y_final, double, dim)
        DEB_OUT(status, int, 1)
        DEB_OUT(accepted, int, 1)
        DEB_OUT(rejected, int, 1)
        DEB_OUT(evals, int, 1)
        DEB_OUT(t_out, double, n_eval)
#undef DEB_OUT
        return DEB_OK;
    };

    auto enqueue = [&](Slot& S, long long lb0, long long lb1) -> int {
        cudaStream_t st = S.r->st;
        S.lb0 = lb0; S.lb1 = lb1;
        const long long cnt = M.count(lb0, lb1);
        const long long nb = lb1 - lb0;
        // slot reuse: its previous chunk's copies have been issued on S.r->cp and recorded in cp_done
        if (S.busy) DEB_CUDA(cudaStreamWaitEvent(st, S.r->cp_done, 0));
        if (int rc = copy_blocks(M, (char*)P->y0, (char*)S.y0.p, sizeof(double) * dim, lb0, lb0, lb1, false, st)) return rc;
        if (C.per_traj_params)
            if (int rc = copy_blocks(M, (char*)P->params, (char*)S.params.p, sizeof(double) * np, lb0, lb0, lb1, false, st)) return rc;
        DEB_CUDA(cudaMemsetAsync(S.small.p, 0, 8, st));
        int* wm_done = (int*)((char*)S.small.p + 8 + rows_bytes);
        DEB_CUDA(cudaMemsetAsync(wm_done, 0, sizeof(int) * (size_t)nb, st));
        for (long long b = 0; b < nb; b++) ((volatile int*)S.r->flags)[b] = 0;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        deb::OdeKernelArgs ac = C.a;
        ac.queue = (unsigned long long*)S.small.p;
        ac.t_rows = (const double*)((char*)S.small.p + 8);
        ac.y0 = S.y0.as<double>();
        ac.params = C.per_traj_params ? S.params.as<double>() : shared_params.as<double>();
        ac.n_traj = cnt;
        ac.y_eval = S.y_eval.as<double>(); ac.n_emitted = S.n_emitted.as<int>(); ac.t_final = S.t_final.as<double>();
        ac.y_final = S.y_final.as<double>(); ac.status = S.status.as<int>(); ac.accepted = S.accepted.as<int>();
        ac.rejected = S.rejected.as<int>(); ac.evals = S.evals.as<int>();
        ac.t_out = S.t_out.as<double>();
        ac.wm_done = wm_done;
        ac.wm_ready = S.r->flags;
        ac.wm_shift = M.shift;
        ac.rows_vec = (ac.y_eval && ((size_t)C.row_cap * dim) % 4 == 0 && ((uintptr_t)ac.y_eval % 32) == 0) ? 1 : 0;
        ac.tout_vec = (ac.t_out && C.row_cap % 4 == 0 && ((uintptr_t)ac.t_out % 32) == 0) ? 1 : 0;
        DEB_CUDA(cudaEventRecord(S.r->k0, st));
        if (int rc = launch(ac, di.sms, st)) return rc;
        DEB_CUDA(cudaEventRecord(S.r->k1, st));
        S.busy = true;
        return DEB_OK;
    };

    const long long batch = 8;  // copy out when at least this many blocks are ready (or the kernel has ended)
    // DEB_DEBUG_TIMING=1: where the wall time of a HOST call goes (stderr)
    const bool dbg_t = getenv("DEB_DEBUG_TIMING") != nullptr;
    long long dbg_blocks_before_end = 0, dbg_copy_calls = 0;
    const double dbg_t_setup = since();
    auto drain = [&](Slot& S) -> int {
        const long long nb = S.lb1 - S.lb0;
        const double dbg_t0 = since();
        double dbg_t_kernel_done = 0.0, dbg_t_first = 0.0;
        volatile int* flags = S.r->flags;
        long long next = 0;
        bool kernel_done = false;
        int idle_spins = 0;
        while (next < nb) {
            long long hi = next;
            while (hi < nb && flags[hi] != 0) hi++;
            std::atomic_thread_fence(std::memory_order_acquire);
            if (dbg_t && dbg_t_first == 0.0 && hi > 0) dbg_t_first = since();
            if (hi > next && (hi - next >= batch || hi == nb || kernel_done)) {
                if (int rc = copy_out(S, S.lb0 + next, S.lb0 + hi, true)) return rc;
                dbg_copy_calls += 1;
                if (!kernel_done) dbg_blocks_before_end += hi - next;
                next = hi;
                idle_spins = 0;
                continue;
            }
            if (kernel_done) {
                // every output is final; anything not flagged (a kernel without watermark support) is copied now
                if (int rc = copy_out(S, S.lb0 + next, S.lb0 + nb, true)) return rc;
                next = nb;
                break;
            }
            // Nothing new: wait on the flags alone (plain host memory).  The kernel-end query is a driver call behind a process-wide
            // lock -- with one polling thread per device it would slow the other threads' copies down -- and every kernel publishes its
            // last block anyway, so it is asked only every few milliseconds, as a safety net.
            idle_spins += 1;
            if (idle_spins % 128 == 0) {
                const cudaError_t q = cudaEventQuery(S.r->k1);
                if (q == cudaSuccess) { kernel_done = true; dbg_t_kernel_done = since(); continue; }
                if (q != cudaErrorNotReady) DEB_CUDA(q);
            }
            std::this_thread::sleep_for(std::chrono::microseconds(idle_spins < 16 ? 5 : 40));
        }
        DEB_CUDA(cudaEventSynchronize(S.r->k1));
        float ms = 0.f;
        DEB_CUDA(cudaEventElapsedTime(&ms, S.r->k0, S.r->k1));
        out->kernel_ms += ms;
        const long long cnt = M.count(S.lb0, S.lb1);
        if (C.want_stats) {
            // the rows are still resident: reduce them here (kernel stream), accumulate over the chunks of this device
            if (int rc = launch_stats(S.y_eval.as<double>(), S.n_emitted.as<int>(), cnt, n_eval, dim, (double*)out->stats.p,
                                      (long long*)((char*)out->stats.p + sizeof(double) * 2 * ne), true, di.sms, S.r->st)) return rc;
        }
        if (row_major) {
            // transposed on the device, then one strided copy per contiguous run of trajectories
            if (int rc = launch_transpose(S.y_eval.as<double>(), cnt, (int)ne, S.y_eval_t.as<double>(), cnt, 0, S.r->st)) return rc;
            DEB_CUDA(cudaStreamSynchronize(S.r->st));
            const size_t total_pitch = sizeof(double) * (size_t)M.n_total;
            long long done_local = 0;
            for (long long lb = S.lb0; lb < S.lb1;) {
                long long run = (M.G == 1) ? (S.lb1 - lb) : 1;  // G > 1: consecutive local blocks are not adjacent in the caller's array
                const long long c = M.count(lb, lb + run);
                DEB_CUDA(cudaMemcpy2DAsync((char*)R->y_eval + sizeof(double) * (size_t)(M.global_block(lb) << M.shift), total_pitch,
                                           (char*)S.y_eval_t.p + sizeof(double) * (size_t)done_local, sizeof(double) * (size_t)cnt,
                                           sizeof(double) * (size_t)c, ne, cudaMemcpyDeviceToHost, S.r->cp));
                done_local += c;
                lb += run;
            }
        }
        if (C.want_stats || row_major) DEB_CUDA(cudaStreamSynchronize(S.r->st));
        for (int q = 0; q + 1 < n_copy_streams; q++) {  // the extra copy streams join the main one
            DEB_CUDA(cudaEventRecord(S.r->cpx_done[q], S.r->cpx[q]));
            DEB_CUDA(cudaStreamWaitEvent(S.r->cp, S.r->cpx_done[q], 0));
        }
        DEB_CUDA(cudaEventRecord(S.r->cp_done, S.r->cp));
        if (dbg_t) {
            const double t_issued = since();
            cudaStreamSynchronize(S.r->cp);
            fprintf(stderr, "[deb timing] device %d chunk of %lld blocks: buffers ready %.1f ms, kernel enqueued %.1f ms, first block done %.1f ms, "
                            "kernel end seen %.1f ms, last copy issued %.1f ms, copies complete %.1f ms; kernel %.1f ms; %lld of %lld blocks copied "
                            "before the kernel ended, %lld copy batches\n",
                    device, nb, dbg_t_setup, dbg_t0, dbg_t_first, dbg_t_kernel_done, t_issued, since(), ms, dbg_blocks_before_end, nb, dbg_copy_calls);
        }
        return DEB_OK;
    };

    for (long long ci = 0; ci < n_chunks; ci++) {
        Slot& S = slot[ci % n_slots];
        const long long lb0 = ci * chunk_blocks, lb1 = std::min(nlb, lb0 + chunk_blocks);
        if (int rc = enqueue(S, lb0, lb1)) return rc;
        if (ci >= 1) { if (int rc = drain(slot[(ci - 1) % n_slots])) return rc; }
    }
    if (n_chunks > 0) { if (int rc = drain(slot[(n_chunks - 1) % n_slots])) return rc; }
    for (int s = 0; s < n_slots && n_chunks > 0; s++) DEB_CUDA(cudaStreamSynchronize(slot[s].r->cp));
    if (dbg_t) fprintf(stderr, "[deb timing] device %d: shard finished %.1f ms (before buffers are released)\n", device, since());
    return DEB_OK;
}

}  // namespace

extern "C" int deb_solve_ode(const deb_ode_problem* P_user, deb_result* R_user) {
    OdeCall C;
    if (int rc = import_struct(P_user, &C.P, offsetof(deb_ode_problem, solout), "deb_ode_problem")) return rc;
    ResultIO rio;
    if (int rc = rio.open(R_user)) return rc;
    const deb_ode_problem* P = &C.P;
    deb_result* R = &rio.local;
    C.R = R;
    int dim = 0, np = 0;
    if (P->system >= USER_SYSTEM_BASE) {
        std::lock_guard<std::mutex> lk(g_user_mu);
        const size_t k = (size_t)(P->system - USER_SYSTEM_BASE);
        if (k >= g_user_systems.size()) return fail(DEB_ERR_BAD_ARG, "unknown system id");
        C.user = g_user_systems[k].get();
        dim = C.user->dim;
        np = C.user->np;
        bool adaptive;
        if (!method_tab_name(P->method, &adaptive)) return fail(DEB_ERR_UNSUPPORTED, "unknown or unsupported method id");
    } else {
        C.builtin = pick_ode(P->system, P->method, &dim, &np);
        if (dim < 0) return fail(DEB_ERR_BAD_ARG, "unknown system id");
        if (!C.builtin) return fail(DEB_ERR_UNSUPPORTED, "unknown or unsupported method id");
    }
    C.dim = dim;
    C.np = np;
    if (P->dim != dim || P->n_params != np) return fail(DEB_ERR_BAD_ARG, "dim / n_params do not match the system");
    if (P->n_traj < 0) return fail(DEB_ERR_BAD_ARG, "n_traj < 0");
    if (P->n_traj > 0 && (!P->y0 || (np > 0 && !P->params))) return fail(DEB_ERR_BAD_ARG, "NULL y0/params");
    if (P->n_eval < 0) return fail(DEB_ERR_BAD_ARG, "n_eval < 0");
    if (int rc = check_options(P->opt)) return rc;
    TEvalPlan& plan = C.plan;
    const bool even = C.even = (P->solout == DEB_SOLOUT_EVEN);
    const bool per_step = C.per_step = (P->solout == DEB_SOLOUT_DEFAULT || P->solout == DEB_SOLOUT_DENSE || P->solout == DEB_SOLOUT_CROSSING ||
                                        P->solout == DEB_SOLOUT_HYPERPLANE);
    if (P->solout != DEB_SOLOUT_T_EVAL && !even && !per_step) return fail(DEB_ERR_BAD_ARG, "unknown solout mode");
    const bool has_event = C.has_event = (P->event != DEB_EVENT_NONE);
    const bool rec = C.rec = per_step || has_event;  // rows with their own times: recorder kernels (compiled at first use)
    if (has_event) {
        if (P->event_direction < -1 || P->event_direction > 1) return fail(DEB_ERR_BAD_ARG, "event: direction must be -1, 0 or +1");
        if (P->event_terminate < 0) return fail(DEB_ERR_BAD_ARG, "event: terminate count < 0");
        if (P->row_capacity < 0) return fail(DEB_ERR_BAD_ARG, "row_capacity < 0");
        if (!R->y_eval) return fail(DEB_ERR_BAD_ARG, "event detection needs a y_eval buffer");
    }
    const int row_cap = C.row_cap = (has_event && P->row_capacity > 0) ? P->row_capacity : P->n_eval;
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
    // ---- ABI 9 fields
    if (P->filter != DEB_FILTER_IDENTITY && P->filter != DEB_FILTER_TRUNCATE_MANTISSA) return fail(DEB_ERR_BAD_ARG, "unknown step-size filter");
    if (P->filter == DEB_FILTER_TRUNCATE_MANTISSA && (P->filter_bits < 1 || P->filter_bits > 52))
        return fail(DEB_ERR_BAD_ARG, "filter_bits must be in 1..52");
    if (P->layout != DEB_LAYOUT_TRAJ_MAJOR && P->layout != DEB_LAYOUT_ROW_MAJOR) return fail(DEB_ERR_BAD_ARG, "unknown y_eval layout");
    if (P->layout == DEB_LAYOUT_ROW_MAJOR && rec) return fail(DEB_ERR_BAD_ARG, "DEB_LAYOUT_ROW_MAJOR needs a t_eval / even(dt) recorder without event");
    C.want_stats = (R->stats_sums != nullptr || R->stats_counts != nullptr);
    if (C.want_stats) {
        if (!R->stats_sums || !R->stats_counts) return fail(DEB_ERR_BAD_ARG, "stats_sums and stats_counts go together");
        if (rec) return fail(DEB_ERR_BAD_ARG, "ensemble statistics need a t_eval / even(dt) recorder without event");
        if (P->memspace != DEB_MEM_HOST) return fail(DEB_ERR_BAD_ARG, "stats_sums / stats_counts are filled by DEB_MEM_HOST calls (use deb_ensemble_stats on device-resident rows)");
        if (P->n_eval <= 0) return fail(DEB_ERR_BAD_ARG, "ensemble statistics need n_eval > 0");
    }
    if (P->n_devices < 0 || P->n_devices > DEB_MAX_DEVICES) return fail(DEB_ERR_BAD_ARG, "n_devices must be in 0..DEB_MAX_DEVICES");
    if (P->n_devices >= 2) {
        if (P->memspace != DEB_MEM_HOST) return fail(DEB_ERR_BAD_ARG, "a device list needs memspace = DEB_MEM_HOST");
        for (int g = 0; g < P->n_devices; g++) {
            for (int q = 0; q < g; q++)
                if (P->devices[q] == P->devices[g]) return fail(DEB_ERR_BAD_ARG, "duplicate device in the device list");
            C.devices.push_back(P->devices[g]);
        }
    } else {
        C.devices.push_back(P->n_devices == 1 ? P->devices[0] : P->device);
    }
    publish_rows(R, plan, even);
    R->kernel_ms = 0.f;
    R->total_ms = 0.f;
    R->gpu_launches = 0;
    if (P->n_traj == 0) {  // empty ensemble: nothing to do, and no device needed
        if (C.want_stats) {
            memset(R->stats_sums, 0, sizeof(double) * 2 * (size_t)P->n_eval * dim);
            memset(R->stats_counts, 0, sizeof(int64_t) * (size_t)P->n_eval);
        }
        return DEB_OK;
    }
    const long long launches0 = g_launches.load();
    bool adaptive_method = false;
    method_tab_name(P->method, &adaptive_method);
    const bool filtered = (P->filter != DEB_FILTER_IDENTITY) && adaptive_method;  // the fixed-step stepper never calls the hook
    C.jit = C.user || rec || filtered;  // kernels with a non-identity filter are compiled at first use
    const bool host = (P->memspace == DEB_MEM_HOST);
    const long long n = P->n_traj;

    // ---- kernel arguments common to every launch of this call
    deb::OdeKernelArgs& a = C.a;
    memset(&a, 0, sizeof a);
    if (np > 0 && P->params_shared) {
        // one parameter set for the whole ensemble: HOST memory by contract, passed by value (constant bank)
        for (int q = 0; q < np && q < 8; q++) a.pc[q] = P->params[q];
    }
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
    // the kernels count steps in 32-bit registers: a larger max_steps cannot be reached before the counters wrap
    if (P->opt.max_steps > 0x7fffffff / 16) return fail(DEB_ERR_UNSUPPORTED, "max_steps above 2^27 - 1 is not supported (32-bit step counters)");
    a.max_steps = (int)P->opt.max_steps;
    a.max_rejects = (int)std::min<int64_t>(std::max<int64_t>(P->opt.max_rejects, 0), 0x7fffffff);
    a.filter_mask = filtered ? ~((1ull << (52 - P->filter_bits)) - 1ull) : 0ull;
    {   // fixed-step methods: the step schedule is the same for every trajectory (plan_fixed_schedule)
        bool adaptive_m = false;
        method_tab_name(P->method, &adaptive_m);
        a.fx_status = DEB_STATUS_COMPLETE;
        if (!adaptive_m) {
            FixedSchedule fs;
            if (int rc = plan_fixed_schedule(P->t0, P->tf, P->opt.h0, P->opt.h_min, P->opt.h_max, a.max_steps, &fs, [](long long, double, double) {}))
                return rc;
            a.fx_n_steps = (int)fs.n_steps;
            a.fx_n_tail = fs.n_tail;
            for (int q = 0; q < deb::DEB_FX_MAX_TAIL; q++) a.fx_h_tail[q] = fs.h_tail[q];
            a.fx_status = fs.status;
        }
    }
    a.n_rows = (int)plan.rows.size();
    a.row_stride = row_cap;
    a.emit_t0 = plan.emit_t0 ? 1 : 0;
    a.even = even ? 1 : 0;
    a.even_tol = fabs(P->even_dt) * 1e-12 + 2.220446049250313e-16 * 10.0;
    a.rec_mode = P->solout;
    {   // lanes that gather before a warp refines crossings / events together (erk_ensemble.cuh: rec_park); measured optimum
        static const int park = [] { const char* e = getenv("DEB_REC_PARK"); return e ? atoi(e) : 16; }();
        a.rec_park = (park < 0) ? 0 : (park > 32 ? 32 : park);
        static const int park_rows = [] { const char* e = getenv("DEB_REC_PARK_ROWS"); return e ? atoi(e) : 0; }();
        a.rec_park_rows = park_rows;
    }
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
    C.per_traj_params = (np > 0 && !P->params_shared);
    const size_t rows_bytes = sizeof(double) * plan.rows.size();
    const size_t ne = (size_t)row_cap * dim;

    if (!host) {
        // ---- DEVICE memspace: pointers are device pointers; enqueue on the caller's stream and return
        if (int rc = select_device(C.devices[0])) return rc;
        DeviceInfo di;
        if (int rc = device_info(C.devices[0], &di)) return rc;
        LaunchFn launch;
        if (int rc = bind_launcher(C, C.devices[0], &launch)) return rc;
        cudaStream_t st = (cudaStream_t)P->stream;
        // run-time kernels (SHARED_P = false) read a shared parameter set through a pointer with stride 0: a few bytes from
        // the stream-ordered pool on the caller's stream (no device-wide synchronisation: the call stays asynchronous)
        PoolBuf p_shared_params;
        if (np > 0 && P->params_shared && (C.jit || np > 8)) {
            DEB_CUDA(p_shared_params.alloc(sizeof(double) * np, st));
            DEB_CUDA(cudaMemcpyAsync(p_shared_params.p, P->params, sizeof(double) * np, cudaMemcpyHostToDevice, st));
        }
        void* d_small = nullptr;  // [queue counter (8 B)] [rows]
        DEB_CUDA(cudaMallocAsync(&d_small, 8 + rows_bytes, st));
        struct SmallFree { void* p; cudaStream_t st; ~SmallFree() { if (p) cudaFreeAsync(p, st); } } small_free{d_small, st};
        DEB_CUDA(cudaMemsetAsync(d_small, 0, 8, st));
        if (rows_bytes) DEB_CUDA(cudaMemcpyAsync((char*)d_small + 8, plan.rows.data(), rows_bytes, cudaMemcpyHostToDevice, st));
        const bool row_major = (P->layout == DEB_LAYOUT_ROW_MAJOR) && R->y_eval;
        PoolBuf tmp_rows;  // ROW_MAJOR: the kernel writes trajectory-major rows here, a transpose pass fills the caller's buffer
        if (row_major) DEB_CUDA(tmp_rows.alloc(sizeof(double) * (size_t)n * ne, st));
        a.queue = (unsigned long long*)d_small;
        a.t_rows = (const double*)((char*)d_small + 8);
        a.y0 = P->y0;
        a.params = C.per_traj_params ? P->params : p_shared_params.as<double>();
        a.n_traj = n;
        a.y_eval = row_major ? tmp_rows.as<double>() : R->y_eval;
        a.n_emitted = R->n_emitted; a.t_final = R->t_final; a.y_final = R->y_final;
        a.status = R->status; a.accepted = R->accepted; a.rejected = R->rejected; a.evals = R->evals;
        a.t_out = R->t_out;
        a.rows_vec = (a.y_eval && (ne % 4) == 0 && ((uintptr_t)a.y_eval % 32) == 0) ? 1 : 0;
        a.tout_vec = (a.t_out && row_cap % 4 == 0 && ((uintptr_t)a.t_out % 32) == 0) ? 1 : 0;
        if (int rc = launch(a, di.sms, st)) return rc;
        if (row_major)
            if (int rc = launch_transpose(tmp_rows.as<double>(), n, (int)ne, R->y_eval, n, 0, st)) return rc;
        R->gpu_launches = (int32_t)(g_launches.load() - launches0);
        return DEB_OK;
    }

    // ---- HOST memspace: one persistent launch per device (and chunk) with watermark-streamed result copies (run_shard)
    const auto wall0 = std::chrono::steady_clock::now();
    C.wall0 = wall0;
    const int G = (int)C.devices.size();
    int shift = 12;
    if (const char* e = getenv("DEB_WM_SHIFT")) {  // test knob: log2 of the watermark / distribution block
        const int v = atoi(e);
        if (v >= 0 && v <= 24) shift = v;
    }
    {
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0)
            return fail(DEB_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); the ensemble integrator has no CPU fallback");
        for (int d : C.devices)
            if (d < 0 || d >= n_dev) return fail(DEB_ERR_BAD_ARG, "device ordinal out of range");
    }
    std::vector<ShardOut> outs(G);
    std::vector<ShardMap> maps(G);
    for (int g = 0; g < G; g++) {
        maps[g].n_total = n;
        maps[g].shift = shift;
        maps[g].G = G;
        maps[g].g = g;
    }
    if (G == 1) {
        outs[0].rc = run_shard(C, maps[0], C.devices[0], &outs[0]);
        if (outs[0].rc) return outs[0].rc;
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g] {
                outs[g].rc = run_shard(C, maps[g], C.devices[g], &outs[g]);
                if (outs[g].rc) outs[g].err = g_err;
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < G; g++)
            if (outs[g].rc) return fail(outs[g].rc, "device " + std::to_string(C.devices[g]) + ": " + outs[g].err);
    }
    float kernel_ms = 0.f;
    for (int g = 0; g < G; g++) kernel_ms = std::max(kernel_ms, outs[g].kernel_ms);
    const bool dbg_t = getenv("DEB_DEBUG_TIMING") != nullptr;
    const double t_joined = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (C.want_stats) {
        // every device holds the sums over its own trajectories; the all-reduce of these 2*ne doubles + n_eval counts is the
        // only cross-device traffic of the call
        const size_t nd = 2 * ne, nc = (size_t)P->n_eval;
        int first = -1;
        for (int g = 0; g < G; g++)
            if (outs[g].stats.p) { first = g; break; }
        if (first < 0) return fail(DEB_ERR_CUDA, "internal: no device produced statistics");
        if (G > 1) {
            std::vector<int> devs;
            std::vector<void*> bufs_d, bufs_c;
            std::vector<cudaStream_t> sts;
            for (int g = 0; g < G; g++) {
                if (!outs[g].stats.p) continue;  // a device without trajectories (fewer blocks than devices)
                devs.push_back(C.devices[g]);
                bufs_d.push_back(outs[g].stats.p);
                bufs_c.push_back((char*)outs[g].stats.p + sizeof(double) * nd);
                sts.push_back(outs[g].stats_stream);
            }
            if (devs.size() > 1 && nccl_api()) {
                if (int rc = allreduce_across_devices(devs, bufs_d, nd, /*ncclDouble*/ 8, sts)) return rc;
                if (int rc = allreduce_across_devices(devs, bufs_c, nc, /*ncclInt64*/ 4, sts)) return rc;
            } else if (devs.size() > 1) {
                // no libnccl in the process or on the loader path: the per-device sums (a few KB each) are added on the host, in device order
                std::vector<double> acc_d(nd, 0.0), tmp_d(nd);
                std::vector<long long> acc_c(nc, 0), tmp_c(nc);
                for (size_t q = 0; q < devs.size(); q++) {
                    if (int rc = select_device(devs[q])) return rc;
                    DEB_CUDA(cudaMemcpyAsync(tmp_d.data(), bufs_d[q], sizeof(double) * nd, cudaMemcpyDeviceToHost, sts[q]));
                    DEB_CUDA(cudaMemcpyAsync(tmp_c.data(), bufs_c[q], sizeof(long long) * nc, cudaMemcpyDeviceToHost, sts[q]));
                    DEB_CUDA(cudaStreamSynchronize(sts[q]));
                    for (size_t i = 0; i < nd; i++) acc_d[i] += tmp_d[i];
                    for (size_t i = 0; i < nc; i++) acc_c[i] += tmp_c[i];
                }
                if (int rc = select_device(C.devices[first])) return rc;
                DEB_CUDA(cudaMemcpyAsync(outs[first].stats.p, acc_d.data(), sizeof(double) * nd, cudaMemcpyHostToDevice, outs[first].stats_stream));
                DEB_CUDA(cudaMemcpyAsync((char*)outs[first].stats.p + sizeof(double) * nd, acc_c.data(), sizeof(long long) * nc, cudaMemcpyHostToDevice,
                                         outs[first].stats_stream));
                DEB_CUDA(cudaStreamSynchronize(outs[first].stats_stream));
            }
        }
        if (int rc = select_device(C.devices[first])) return rc;
        DEB_CUDA(cudaMemcpyAsync(R->stats_sums, outs[first].stats.p, sizeof(double) * nd, cudaMemcpyDeviceToHost, outs[first].stats_stream));
        DEB_CUDA(cudaMemcpyAsync(R->stats_counts, (char*)outs[first].stats.p + sizeof(double) * nd, sizeof(long long) * nc, cudaMemcpyDeviceToHost,
                                 outs[first].stats_stream));
        for (int g = 0; g < G; g++) {
            if (!outs[g].stats.p) continue;
            if (int rc = select_device(C.devices[g])) return rc;
            DEB_CUDA(cudaStreamSynchronize(outs[g].stats_stream));
            outs[g].stats.release();
        }
    }
    R->kernel_ms = kernel_ms;  // per device: sum over its chunks; the call reports the slowest device
    R->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (dbg_t) fprintf(stderr, "[deb timing] call: all %d device(s) done %.1f ms, statistics reduced and copied %.1f ms\n", G, t_joined, (double)R->total_ms);
    R->gpu_launches = (int32_t)(g_launches.load() - launches0);
    return DEB_OK;
}

extern "C" int deb_shard_layout(int64_t n_traj, int32_t n_devices, int32_t index, int32_t shift, int64_t local_block, int64_t* n_local_blocks,
                                int64_t* n_local_traj, int64_t* global_block) {
    if (!n_local_blocks || !n_local_traj || !global_block) return fail(DEB_ERR_BAD_ARG, "NULL argument");
    if (n_traj < 0 || n_devices < 1 || index < 0 || index >= n_devices || shift < 0 || shift > 30) return fail(DEB_ERR_BAD_ARG, "bad shard description");
    ShardMap M;
    M.n_total = n_traj;
    M.shift = shift;
    M.G = n_devices;
    M.g = index;
    *n_local_blocks = M.local_blocks();
    *n_local_traj = M.local_count();
    *global_block = (local_block >= 0 && local_block < M.local_blocks()) ? M.global_block(local_block) : -1;
    return DEB_OK;
}

extern "C" int deb_plan_fixed_steps(double t0, double tf, double h0, double h_min, double h_max, int64_t max_steps, int64_t* n_steps, int32_t* n_tail,
                                    double* h_tail, int32_t* status) {
    if (!n_steps || !n_tail || !h_tail || !status) return fail(DEB_ERR_BAD_ARG, "NULL argument");
    FixedSchedule fs;
    if (int rc = plan_fixed_schedule(t0, tf, h0, h_min, h_max, (long long)max_steps, &fs, [](long long, double, double) {})) return rc;
    *n_steps = fs.n_steps;
    *n_tail = fs.n_tail;
    for (int q = 0; q < deb::DEB_FX_MAX_TAIL; q++) h_tail[q] = fs.h_tail[q];
    *status = fs.status;
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

extern "C" int deb_define_sde(int32_t dim, int32_t n_params, const char* drift_body, const char* diffusion_body, const char* noise_body,
                              int32_t* system_id) {
    if (!drift_body || !diffusion_body || !system_id) return fail(DEB_ERR_BAD_ARG, "NULL argument");
    if (dim < 1 || dim > DEB_MAX_DIM) return fail(DEB_ERR_BAD_ARG, "dim must be in 1..DEB_MAX_DIM");
    if (n_params < 0 || n_params > 8) return fail(DEB_ERR_BAD_ARG, "n_params must be in 0..8");
    std::lock_guard<std::mutex> lk(g_user_mu);
    std::unique_ptr<UserSde> us(new UserSde);
    us->dim = dim;
    us->np = n_params;
    us->drift = drift_body;
    us->diffusion = diffusion_body;
    us->noise = noise_body ? noise_body : "";
    g_user_sdes.push_back(std::move(us));
    *system_id = USER_SYSTEM_BASE + (int32_t)g_user_sdes.size() - 1;
    return DEB_OK;
}

extern "C" int deb_check_sde(int32_t system_id, int32_t method) {
    std::lock_guard<std::mutex> lk(g_user_mu);
    if (system_id < USER_SYSTEM_BASE) {
        bool m = false;
        if (system_id < DEB_SDE_OU || system_id > DEB_SDE_HESTON) return fail(DEB_ERR_BAD_ARG, "unknown SDE system id");
        return sde_tab_name(method, &m) ? DEB_OK : fail(DEB_ERR_UNSUPPORTED, "SDE ensembles take a fixed-step method id or DEB_MILSTEIN");
    }
    const size_t k = (size_t)(system_id - USER_SYSTEM_BASE);
    if (k >= g_user_sdes.size()) return fail(DEB_ERR_BAD_ARG, "unknown SDE system id");
    std::vector<char> cubin;
    std::string name;
    return compile_sde_cubin(*g_user_sdes[k], method, &cubin, &name);
}

// ForwardSensitivityOde::diff (/root/reference/src/ode/sensitivity/forward.rs:82-115) generated as the body of a user system.
extern "C" int deb_define_ode_sensitivity(int32_t dim, int32_t n_params, const char* diff_body, const char* jac_y_body, const char* jac_p_body,
                                          int32_t* system_id) {
    if (!diff_body || !jac_y_body || !jac_p_body || !system_id) return fail(DEB_ERR_BAD_ARG, "NULL argument");
    if (dim < 1 || n_params < 1) return fail(DEB_ERR_BAD_ARG, "dim and n_params must be >= 1");
    const long long aug = (long long)dim * (1 + n_params);
    if (aug > DEB_MAX_DIM) return fail(DEB_ERR_BAD_ARG, "the augmented state dim*(1+n_params) exceeds DEB_MAX_DIM");
    const std::string n = std::to_string(dim), m = std::to_string(n_params);
    std::string b;
    b += "        const double* const yaug_ = y;\n        double* const daug_ = dydt;\n";
    b += "        double yb_[" + n + "], fb_[" + n + "], J_[" + n + " * " + n + "], Jp_[" + n + " * " + m + "];\n";
    b += "#pragma unroll\n        for (int i_ = 0; i_ < " + n + "; i_++) yb_[i_] = yaug_[i_];\n";              // y_cache, forward.rs:84-87
    b += "#pragma unroll\n        for (int i_ = 0; i_ < " + n + " * " + n + "; i_++) J_[i_] = 0.0;\n";          // Matrix::full: zeros
    b += "#pragma unroll\n        for (int i_ = 0; i_ < " + n + " * " + m + "; i_++) Jp_[i_] = 0.0;\n";
    b += "        {\n            const double* y = yb_;\n            double* dydt = fb_;\n            (void)y; (void)dydt;\n";
    b += diff_body;                                                                                            // self.ode.diff, :89-90
    b += "\n        }\n#pragma unroll\n        for (int i_ = 0; i_ < " + n + "; i_++) daug_[i_] = fb_[i_];\n";          // :91-93
    b += "        {\n            const double* y = yb_;\n            double* J = J_;\n            (void)y; (void)J;\n";
    b += jac_y_body;                                                                                           // self.ode.jacobian, :95-96
    b += "\n        }\n        {\n            const double* y = yb_;\n            double* Jp = Jp_;\n            (void)y; (void)Jp;\n";
    b += jac_p_body;                                                                                           // self.ode.jacobian_p, :98-99
    b += "\n        }\n";
    // dS/dt = J_y * S + J_p with S row-major after y, accumulated as written (forward.rs:103-113)
    b += "#pragma unroll\n        for (int r_ = 0; r_ < " + n + "; r_++) {\n#pragma unroll\n            for (int c_ = 0; c_ < " + m + "; c_++) {\n";
    b += "                double ds_ = Jp_[r_ * " + m + " + c_];\n#pragma unroll\n                for (int k_ = 0; k_ < " + n + "; k_++) ds_ = ds_ + J_[r_ * " + n + " + k_] * yaug_[" + n + " + k_ * " + m + " + c_];\n";
    b += "                daug_[" + n + " + r_ * " + m + " + c_] = ds_;\n            }\n        }\n";
    return deb_define_ode((int32_t)aug, n_params, b.c_str(), system_id);
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
    return compile_kernel_cubin(us, system_id, method, rec ? solout : -1, event, false, &cubin, &name, &adaptive);
}

extern "C" int deb_solve_sde(const deb_sde_problem* P_user, deb_result* R_user) {
    deb_sde_problem P_local;
    if (int rc = import_struct(P_user, &P_local, sizeof(deb_sde_problem), "deb_sde_problem")) return rc;
    ResultIO rio;
    if (int rc = rio.open(R_user)) return rc;
    const deb_sde_problem* P = &P_local;
    deb_result* R = &rio.local;
    sde_launch_fn launch = nullptr;
    const UserSde* user = nullptr;
    int np = 0, dim = 1;
    if (P->system == DEB_SDE_OU) { launch = pick_sde_method<deb::SdeOU>(P->method); np = deb::SdeOU::NP; }
    else if (P->system == DEB_SDE_GBM) { launch = pick_sde_method<deb::SdeGBM>(P->method); np = deb::SdeGBM::NP; }
    else if (P->system == DEB_SDE_HESTON) { launch = pick_sde_method<deb::SdeHeston>(P->method); np = deb::SdeHeston::NP; dim = deb::SdeHeston::DIM; }
    else if (P->system >= USER_SYSTEM_BASE) {
        std::lock_guard<std::mutex> lk(g_user_mu);
        const size_t k = (size_t)(P->system - USER_SYSTEM_BASE);
        if (k >= g_user_sdes.size()) return fail(DEB_ERR_BAD_ARG, "unknown SDE system id");
        user = g_user_sdes[k].get();
        np = user->np;
        dim = user->dim;
        bool m = false;
        if (!sde_tab_name(P->method, &m)) return fail(DEB_ERR_UNSUPPORTED, "SDE ensembles take a fixed-step method id or DEB_MILSTEIN");
    }
    else return fail(DEB_ERR_BAD_ARG, "unknown SDE system id");
    if (!launch && !user) return fail(DEB_ERR_UNSUPPORTED, "SDE ensembles take a fixed-step method id or DEB_MILSTEIN");
    if (P->dim != dim || P->n_params != np) return fail(DEB_ERR_BAD_ARG, "dim / n_params do not match the SDE system");
    if (P->n_traj < 0 || P->n_eval < 0) return fail(DEB_ERR_BAD_ARG, "negative size");
    if (P->n_traj > 0 && (!P->y0 || (np > 0 && !P->params))) return fail(DEB_ERR_BAD_ARG, "NULL y0/params");
    if (int rc = check_options(P->opt)) return rc;
    if (P->opt.max_steps > 0x7fffffff / 16) return fail(DEB_ERR_UNSUPPORTED, "max_steps above 2^27 - 1 is not supported (32-bit step counters)");
    TEvalPlan plan;
    if (int rc = plan_t_eval(P->t_eval, P->n_eval, P->t0, P->tf, &plan)) return rc;
    publish_rows(R, plan);
    R->kernel_ms = 0.f;
    R->total_ms = 0.f;
    R->gpu_launches = 0;
    if (P->n_traj == 0) return DEB_OK;
    if (int rc = select_device(P->device)) return rc;
    DeviceInfo di;
    if (int rc = device_info(P->device, &di)) return rc;
    const long long launches0 = g_launches.load();
    UserKernel user_kernel;
    if (user) {
        std::lock_guard<std::mutex> lk(g_user_mu);
        UserKernel* uk = nullptr;
        if (int rc = jit_sde_kernel(*user, P->device, P->system, P->method, &uk)) return rc;
        user_kernel = *uk;
    }
    const bool host = (P->memspace == DEB_MEM_HOST);
    // HOST calls run on a cached stream of the library (no per-call stream / event creation), staging buffers come from
    // the library's stream-ordered pool (kept between calls)
    SlotRes* res = nullptr;
    if (host) { if (int rc = acquire_slot(P->device, 1, &res)) return rc; }
    struct ResGuard { int device; SlotRes* r; ~ResGuard() { if (r) { cudaStreamSynchronize(r->st); release_slot(device, r); } } } res_guard{P->device, res};
    cudaStream_t st = host ? res->st : (cudaStream_t)P->stream;
    const long long n = P->n_traj;
    const auto wall0 = std::chrono::steady_clock::now();
    deb::SdeKernelArgs a;
    memset(&a, 0, sizeof a);
    PoolBuf d_y0, d_params;
    // y0: [n_traj] in `memspace`, or ONE value in HOST memory when y0_shared; params likewise (see the header)
    double y0_one[DEB_MAX_DIM];
    if (P->y0_shared) {
        for (int c = 0; c < dim; c++) y0_one[c] = P->y0[c];
        DEB_CUDA(d_y0.alloc(sizeof(double) * dim, st));
        DEB_CUDA(cudaMemcpyAsync(d_y0.p, y0_one, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
        DEB_CUDA(cudaStreamSynchronize(st));  // y0_one is a local
        a.y0 = d_y0.as<double>();
    } else if (host) {
        DEB_CUDA(d_y0.alloc(sizeof(double) * (size_t)n * dim, st));
        DEB_CUDA(cudaMemcpyAsync(d_y0.p, P->y0, sizeof(double) * (size_t)n * dim, cudaMemcpyHostToDevice, st));
        a.y0 = d_y0.as<double>();
    } else {
        a.y0 = P->y0;
    }
    if (P->params_shared || np == 0) {
        for (int q = 0; q < np && q < 8; q++) a.pc[q] = P->params[q];
        a.params = nullptr;
    } else if (host) {
        DEB_CUDA(d_params.alloc(sizeof(double) * (size_t)n * np, st));
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
        const bool fwd = (P->tf - P->t0) > 0.0;
        size_t idx = plan.emit_t0 ? 1 : 0;
        FixedSchedule fs;
        if (int rc = plan_fixed_schedule(P->t0, P->tf, P->opt.h0, P->opt.h_min, P->opt.h_max, a.max_steps, &fs,
                                         [&](long long step, double t, double t_new) {
                                             while (idx < plan.rows.size() && (fwd ? (plan.rows[idx] <= t_new) : (plan.rows[idx] >= t_new))) {
                                                 row_step[idx] = (int)step;
                                                 row_s[idx] = (plan.rows[idx] == t_new) ? -1.0 : (plan.rows[idx] - t) / (t_new - t);
                                                 idx++;
                                             }
                                         }))
            return rc;
        a.n_steps = (int)fs.n_steps;
        a.n_tail = fs.n_tail;
        for (int q = 0; q < deb::DEB_FX_MAX_TAIL; q++) a.h_tail[q] = fs.h_tail[q];
        a.final_status = fs.status;
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
    // result mirrors (HOST) from the pool, or the caller's device pointers
    PoolBuf m_y_eval, m_n_emitted, m_t_final, m_y_final, m_status, m_accepted, m_rejected, m_evals;
    if (host) {
#define DEB_STAGE(field, T, count) \
    if (R->field) { DEB_CUDA(m_##field.alloc(sizeof(T) * (size_t)(count), st)); a.field = m_##field.as<T>(); }
        DEB_STAGE(y_eval, double, (size_t)n * P->n_eval * dim)
        DEB_STAGE(n_emitted, int, n)
        DEB_STAGE(t_final, double, n)
        DEB_STAGE(y_final, double, (size_t)n * dim)
        DEB_STAGE(status, int, n)
        DEB_STAGE(accepted, int, n)
        DEB_STAGE(rejected, int, n)
        DEB_STAGE(evals, int, n)
#undef DEB_STAGE
    } else {
        a.y_eval = R->y_eval; a.n_emitted = R->n_emitted; a.t_final = R->t_final; a.y_final = R->y_final;
        a.status = R->status; a.accepted = R->accepted; a.rejected = R->rejected; a.evals = R->evals;
    }
    a.rows_vec = (a.y_eval && ((size_t)P->n_eval * dim) % 4 == 0 && ((uintptr_t)a.y_eval % 32) == 0) ? 1 : 0;
    if (host) DEB_CUDA(cudaEventRecord(res->k0, st));
    if (user) { if (int rc = launch_user_sde(user_kernel, a, di.sms, st)) return rc; }
    else if (int rc = launch(a, di.sms, st)) return rc;
    if (host) {
        DEB_CUDA(cudaEventRecord(res->k1, st));
#define DEB_BACK(field, T, count) \
    if (R->field) DEB_CUDA(cudaMemcpyAsync(R->field, a.field, sizeof(T) * (size_t)(count), cudaMemcpyDeviceToHost, st));
        DEB_BACK(y_eval, double, (size_t)n * P->n_eval * dim)
        DEB_BACK(n_emitted, int, n)
        DEB_BACK(t_final, double, n)
        DEB_BACK(y_final, double, (size_t)n * dim)
        DEB_BACK(status, int, n)
        DEB_BACK(accepted, int, n)
        DEB_BACK(rejected, int, n)
        DEB_BACK(evals, int, n)
#undef DEB_BACK
        DEB_CUDA(cudaStreamSynchronize(st));
        DEB_CUDA(cudaEventElapsedTime(&R->kernel_ms, res->k0, res->k1));
        R->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    }
    R->gpu_launches = (int32_t)(g_launches.load() - launches0);
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
    deb_count_launch(1);
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
    deb_count_launch(1);
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

extern "C" int deb_solve_heat_mol(const deb_heat_problem* P_user) {
    deb_heat_problem P_local;
    if (int rc = import_struct(P_user, &P_local, sizeof(deb_heat_problem), "deb_heat_problem")) return rc;
    if (P_local.max_steps <= 0) P_local.max_steps = 10000;  // the reference default (erk/mod.rs:139)
    const deb_heat_problem* P = &P_local;
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

extern "C" int deb_heat_rhs(const deb_heat_problem* P_user, const double* u, double* du) {
    if (!u || !du) return fail(DEB_ERR_BAD_ARG, "NULL pointer");
    deb_heat_problem P_local;
    if (int rc = import_struct(P_user, &P_local, sizeof(deb_heat_problem), "deb_heat_problem")) return rc;
    const deb_heat_problem* P = &P_local;
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
    const int ne = n_eval * dim;
    if (!host) return launch_stats(y_eval, n_emitted, n_traj, n_eval, dim, sums, (long long*)counts, false, di.sms, (cudaStream_t)stream);
    // HOST rows: staged through the library's pool in chunks of at most 1 GiB on a cached stream, accumulated chunk by chunk
    // (rows that are still on the device are better reduced there: deb_result.stats_sums of deb_solve_ode)
    SlotRes* res = nullptr;
    if (int rc = acquire_slot(device, 1, &res)) return rc;
    struct ResGuard { int device; SlotRes* r; ~ResGuard() { cudaStreamSynchronize(r->st); release_slot(device, r); } } res_guard{device, res};
    cudaStream_t st = res->st;
    const long long chunk = std::max<long long>(1, (1ll << 30) / (long long)(sizeof(double) * ne));
    PoolBuf d_y, d_ne, d_out;
    DEB_CUDA(d_y.alloc(sizeof(double) * (size_t)std::min<long long>(chunk, std::max<long long>(n_traj, 1)) * ne, st));
    DEB_CUDA(d_ne.alloc(sizeof(int) * (size_t)std::min<long long>(chunk, std::max<long long>(n_traj, 1)), st));
    DEB_CUDA(d_out.alloc(sizeof(double) * 2 * ne + sizeof(long long) * n_eval, st));
    DEB_CUDA(cudaMemsetAsync(d_out.p, 0, sizeof(double) * 2 * ne + sizeof(long long) * n_eval, st));
    for (long long off = 0; off < n_traj; off += chunk) {
        const long long cnt = std::min(chunk, n_traj - off);
        DEB_CUDA(cudaMemcpyAsync(d_y.p, y_eval + (size_t)off * ne, sizeof(double) * (size_t)cnt * ne, cudaMemcpyHostToDevice, st));
        DEB_CUDA(cudaMemcpyAsync(d_ne.p, n_emitted + off, sizeof(int) * (size_t)cnt, cudaMemcpyHostToDevice, st));
        if (int rc = launch_stats(d_y.as<double>(), d_ne.as<int>(), cnt, n_eval, dim, d_out.as<double>(),
                                  (long long*)((char*)d_out.p + sizeof(double) * 2 * ne), true, di.sms, st)) return rc;
    }
    DEB_CUDA(cudaMemcpyAsync(sums, d_out.p, sizeof(double) * 2 * ne, cudaMemcpyDeviceToHost, st));
    DEB_CUDA(cudaMemcpyAsync(counts, (char*)d_out.p + sizeof(double) * 2 * ne, sizeof(long long) * n_eval, cudaMemcpyDeviceToHost, st));
    DEB_CUDA(cudaStreamSynchronize(st));
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
    deb_count_launch(1);
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
        deb_count_launch(1);
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

