// Shared helpers for the sm_100a kernels of the CLIP-PPO observation path.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/clipppo_b200.h"
#include "prof.cuh"

namespace clipppo {

// cudaError_t of the most recent failing runtime/driver call on this host thread.
int& last_cuda_error_ref();

inline int record_cuda(cudaError_t e) {
    if (e == cudaSuccess) return CLIPPPO_OK;
    last_cuda_error_ref() = static_cast<int>(e);
    return CLIPPPO_ERR_CUDA;
}

#define CLIPPPO_CUDA_TRY(expr)                                         \
    do {                                                               \
        cudaError_t e__ = (expr);                                      \
        if (e__ != cudaSuccess) return ::clipppo::record_cuda(e__);    \
    } while (0)

// Launch-error check that never synchronises.
#define CLIPPPO_CHECK_LAUNCH()                        \
    do {                                              \
        ::clipppo::prof_count_launch();               \
        CLIPPPO_CUDA_TRY(cudaGetLastError());         \
    } while (0)

inline cudaStream_t as_stream(clipppo_stream_t s) { return static_cast<cudaStream_t>(s); }

// Launch with the programmatic-stream-serialization attribute (PDL) and an optional cluster size.
// Only for kernels that call pdl_wait() before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// cudaFuncSetAttribute is per DEVICE: a kernel instantiation opts in to large dynamic shared memory / non-portable
// clusters once on every device it is launched on (one process may hold towers on several GPUs).
struct DeviceOnce {
    unsigned long long done = 0;
    bool first_use() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
        const unsigned long long bit = 1ull << dev;
        if (done & bit) return false;
        done |= bit;
        return true;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; `scratch` needs >= 32 floats of shared memory.  Every thread gets the total.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                       // scratch may still be in use by a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    float t = (lane < nwarps) ? scratch[lane] : 0.0f;
    return warp_sum(t);
}

// Programmatic dependent launch (PDL).  Every kernel of the tower calls pdl_launch_dependents() first
// (the next kernel of the stream may begin its prologue - barrier init, TMEM alloc, descriptor
// prefetch - as soon as all CTAs of this one are running) and pdl_wait() before its first access to
// global memory (blocks until the previous grid has completed and its writes are visible).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t ptx_smem(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Streaming 128-bit accesses: read-once / write-once data should not pollute L1.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ld_stream_u32(const void* p) {
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// Four uint8 pixels -> what `obs.float() / 255` gives ON THE DEVICE: ATen divides a CUDA tensor by a host scalar
// as a multiplication by fl(1 / 255) (BinaryDivTrueKernel.cu), which differs from the IEEE quotient in the last bit for
// 126 of the 256 byte values.  No int->float conversion instruction: 0x4B0000vv is the float 2^23 + v.
__device__ __forceinline__ float4 u8x4_over_255(unsigned w) {
    const float r = 1.0f / 255.0f;
    float4 o;
    float* op = &o.x;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        op[i] = __fmul_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 | i)) - 8388608.0f, r);
    return o;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace clipppo
