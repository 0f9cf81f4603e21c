// driver.h — the few CUDA driver-API entry points needed to load and launch NVRTC-compiled kernels,
// resolved at run time through cudaGetDriverEntryPoint so that the library links only against the
// (static) CUDA runtime and still loads on machines without libcuda (CPU-only test boxes).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

namespace hnb_rt {

struct DriverApi {
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
    decltype(&cuLaunchKernelEx) LaunchKernelEx = nullptr;  // launch attributes (programmatic dependent launch)
    decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
    decltype(&cuFuncGetAttribute) FuncGetAttribute = nullptr;
    decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) OccupancyMaxActiveBlocksPerMultiprocessor = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    bool loaded = false;
};

// Resolve the entry points (idempotent). Returns false and fills `err` when no driver is available.
bool load_driver_api(DriverApi& api, std::string& err);
std::string cu_error_string(const DriverApi& api, CUresult r);

}  // namespace hnb_rt
