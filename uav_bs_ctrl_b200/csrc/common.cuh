// Shared helpers for the ubs_gnn kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math_constants.h>
#include <atomic>

namespace ubs {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return 1;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Kernel launch with (optionally) programmatic stream serialization: the grid may be scheduled while the previous kernel
// of the stream is still draining.  Such a kernel executes `griddepcontrol.wait` (pdl_wait()) before it touches anything
// the previous kernel may have written and `griddepcontrol.launch_dependents` (pdl_trigger()) as early as it likes;
// both are no-ops in a normal launch.
template <class Arg>
inline cudaError_t launch_pdl(void (*kernel)(Arg), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl,
                              const Arg& arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// reductions inside aligned groups of G lanes (G power of two <= 32)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace ubs

// Kernels that need more than the 48 KB default of dynamic shared memory opt in to the device maximum ONCE per process.
// The function-local static is initialised thread-safely (C++11), so concurrent first calls are fine and later calls
// cost a load: no unsynchronised "configured so far" counters.
#define UBS_OPT_IN_SMEM(kernel, what)                                                                              \
    do {                                                                                                           \
        static const cudaError_t ubs_rc_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                                227 * 1024);                                      \
        if (ubs_rc_ != cudaSuccess) {                                                                              \
            ubs::set_error("%s: cannot opt in to 227 KB of shared memory: %s", what, cudaGetErrorString(ubs_rc_)); \
            return 1;                                                                                              \
        }                                                                                                          \
    } while (0)

#define UBS_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            ubs::set_error(__VA_ARGS__);  \
            return 2;                     \
        }                                 \
    } while (0)
