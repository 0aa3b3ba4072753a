// kw_search_kernel<REGSCORE = true> — the opt-in register-resident scoring variant (TSGPU_REG_SCORE=1) — in a translation
// unit of its own, with its own copies of the device functions (the namespaces are renamed for this file), so that the
// code generated for the default kernel in tsgpu.cu stays byte-for-byte what was profiled and validated on the GPU
// (tools/sass_funcs.py shows the hash). Compiled into the same libtsgpu.so; reached only through the launcher below.
#define TSGPU_KW_SEARCH_ONLY
#define tsdev tsdev_rs
#define tsk tsk_rs
#include <algorithm>
#include <cstring>
#include "kw_kernels.cuh"

extern "C" __attribute__((visibility("hidden")))
cudaError_t tsgpu_launch_kw_search_regscore(const void* index_dev, const void* kw_params, unsigned n_units, size_t smem, cudaStream_t st) {
    tsk::IndexDev ix;                     // same layouts as tsgpu.cu's tsk::IndexDev / tsk::KwParams (same header)
    tsk::KwParams P;
    std::memcpy(&ix, index_dev, sizeof ix);
    std::memcpy(&P, kw_params, sizeof P);
    auto kern = P.F == 1 ? tsk::kw_search_kernel<true, true> : tsk::kw_search_kernel<true, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024));
    if(e != cudaSuccess) return e;
    kern<<<n_units, tsk::kThreads, smem, st>>>(ix, P);
    return cudaGetLastError();
}
