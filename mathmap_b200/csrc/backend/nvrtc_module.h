// NVRTC compilation (sm_100a) and module loading for generated filter kernels.
// Takes the place of the reference's "write C, run gcc, g_module_open" step
// (backends/cc.c:634-758).
#pragma once
#include <cuda_runtime_api.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "cuda_emit.h"

namespace mmbackend {

void set_error(const std::string &msg);

struct KernelConfig {
    int aa = 0, supersampling = 0, edge_x = 0, edge_y = 0, precise = 1, warp_w = 32;
    int rows = 0;  // 32x8 tiles a block renders in sequence (a launch parameter, not part of the key); 0 = chosen per launch
    int specialize = 1;  // compile the main filter's kernel for the frame-constant branch conditions of the frame (not part of the key)
    std::string spec;    // "#define MM_SPEC_<filter>_<i> 0|1" lines (FilterKernel::spec_conds), part of the key
    int fast_compile = 0;  // the complex elementary functions as real calls (MM_COMPLEX_CALLS, mm_runtime.cuh): shorter compile, slower kernel; part of the key
    std::string key() const;
};

// Minimal slice of the CUDA driver API, resolved through the runtime
// (cudaGetDriverEntryPoint) so the library has no link-time dependency on libcuda.
struct DriverApi {
    typedef int (*ModuleLoadData)(void **module, const void *image);
    typedef int (*ModuleGetFunction)(void **func, void *module, const char *name);
    typedef int (*ModuleUnload)(void *module);
    typedef int (*LaunchKernel)(void *f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned smem, void *stream,
                                void **params, void **extra);
    typedef int (*GetErrorString)(int err, const char **str);
    typedef int (*ModuleGetGlobal)(unsigned long long *dptr, size_t *bytes, void *module, const char *name);
    ModuleGetGlobal module_get_global = nullptr;
    ModuleLoadData module_load_data = nullptr;
    ModuleGetFunction module_get_function = nullptr;
    ModuleUnload module_unload = nullptr;
    LaunchKernel launch_kernel = nullptr;
    GetErrorString get_error_string = nullptr;
    bool load(std::string &err);
    std::string error_string(int code) const;
};
DriverApi *driver_api(std::string &err);

struct LoadedModule {
    void *module = nullptr;
    std::map<const mm::Filter *, void *> functions;
    std::map<const mm::Filter *, void *> row_functions;  // row pre-kernels (absent when a filter has none)
    void *call_overflow = nullptr;                       // device address of mm_call_overflow (modules with filter calls)
    ~LoadedModule();
};

// Everything the backend keeps per mmb_module: the generated source, the cubins
// per kernel configuration and the loaded modules per (configuration, device).
struct ModuleBackend {
    CudaModuleSource source;
    std::mutex mu;
    std::map<std::string, std::string> cubins;                         // config key -> cubin bytes
    std::map<std::string, std::shared_ptr<LoadedModule>> loaded;       // config key + device -> module
    std::string full_source(const KernelConfig &cfg) const;            // what NVRTC sees (for inspection)
    std::shared_ptr<LoadedModule> get(const KernelConfig &cfg, int device, std::string &err);
    double last_compile_ms = 0.0;
};

void set_cubin_cache_dir(const char *dir);  // NULL or "": no persistent cache (the default)

std::shared_ptr<ModuleBackend> get_module_backend(mmb_module *m);  // creates on first use; throws mm::CompileError

}  // namespace mmbackend
