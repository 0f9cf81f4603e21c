#include "driver.h"

namespace hnb_rt {

namespace {
template <typename Fn> bool resolve(const char* name, Fn& out, std::string& err) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        err = std::string("cannot resolve CUDA driver entry point ") + name + ": " +
              (e != cudaSuccess ? cudaGetErrorString(e) : "symbol not found");
        (void)cudaGetLastError();
        return false;
    }
    out = reinterpret_cast<Fn>(p);
    return true;
}
}  // namespace

bool load_driver_api(DriverApi& api, std::string& err) {
    if (api.loaded) return true;
    if (!resolve("cuModuleLoadData", api.ModuleLoadData, err)) return false;
    if (!resolve("cuModuleUnload", api.ModuleUnload, err)) return false;
    if (!resolve("cuModuleGetFunction", api.ModuleGetFunction, err)) return false;
    if (!resolve("cuLaunchKernel", api.LaunchKernel, err)) return false;
    if (!resolve("cuLaunchKernelEx", api.LaunchKernelEx, err)) return false;
    if (!resolve("cuFuncSetAttribute", api.FuncSetAttribute, err)) return false;
    if (!resolve("cuFuncGetAttribute", api.FuncGetAttribute, err)) return false;
    if (!resolve("cuOccupancyMaxActiveBlocksPerMultiprocessor", api.OccupancyMaxActiveBlocksPerMultiprocessor, err)) return false;
    if (!resolve("cuGetErrorString", api.GetErrorString, err)) return false;
    api.loaded = true;
    return true;
}

std::string cu_error_string(const DriverApi& api, CUresult r) {
    const char* s = nullptr;
    if (api.GetErrorString && api.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string((int)r);
}

}  // namespace hnb_rt
