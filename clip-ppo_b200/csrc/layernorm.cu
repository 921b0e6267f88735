// V2 - LayerNorm over the fp32 residual stream, bf16 output for the next GEMM's A operand.
//
// Replaces [clip] LayerNorm (fp32 upcast -> F.layer_norm(eps=1e-5) -> downcast) as used for
// ln_pre / ln_1 / ln_2 / ln_post in VisionTransformer.forward.  One warp per row, the row is held
// in registers (two-pass mean / biased variance like ATen), 128-bit loads, 64-bit bf16 stores.
#include "common.cuh"
#include "gemm.cuh"

namespace clipppo {

namespace {

// width = NV * 128; lane owns float4 chunks lane, lane+32, ...
// OutT = __nv_bfloat16 (dense rows of `width`) or float (written back in place over x: ln_pre).
template <int NV, typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int rows, long long row_stride, OutT* y) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int width = NV * 128;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * row_stride);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / width);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / width) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + 32 * i;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        if constexpr (sizeof(OutT) == 4) {
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
            reinterpret_cast<float4*>(y + static_cast<size_t>(row) * row_stride)[c4] = o;
            continue;
        }
        __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(y) + static_cast<size_t>(row) * width;
        __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(yr + 4 * c4) = pk;
    }
}

}  // namespace

template <typename OutT>
static int layernorm_dispatch(const float* x, const float* gamma, const float* beta, int rows, int width,
                              long long row_stride, OutT* y, cudaStream_t stream) {
    const int warps_per_block = 8;
    const unsigned grid = (rows + warps_per_block - 1) / warps_per_block;
    switch (width / 128) {
        case 4: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<4, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y)); break;
        case 6: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<6, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y)); break;
        case 8: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<8, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y)); break;
        case 10: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<10, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y)); break;
        default: return CLIPPPO_ERR_UNSUPPORTED;
    }
    prof_count_launch();
    return CLIPPPO_OK;
}

int layernorm_inplace_f32_launch(float* x, const float* gamma, const float* beta, int rows, int width,
                                 long long row_stride, cudaStream_t stream) {
    if (!x || !gamma || !beta) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || width <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (width % 128 || (row_stride % 4) || (reinterpret_cast<uintptr_t>(x) % 16)) return CLIPPPO_ERR_ALIGN;
    return layernorm_dispatch<float>(x, gamma, beta, rows, width, row_stride, x, stream);
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, int rows, int width,
                     long long row_stride, void* y_bf16, cudaStream_t stream) {
    if (!x || !gamma || !beta || !y_bf16) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || width <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (width % 128 || (row_stride % 4) || (reinterpret_cast<uintptr_t>(x) % 16) ||
        (reinterpret_cast<uintptr_t>(y_bf16) % 8))
        return CLIPPPO_ERR_ALIGN;
    return layernorm_dispatch<__nv_bfloat16>(x, gamma, beta, rows, width, row_stride,
                                             static_cast<__nv_bfloat16*>(y_bf16), stream);
}

}  // namespace clipppo

extern "C" int clipppo_layernorm_bf16(const float* x, const float* gamma, const float* beta, int rows,
                                      int width, int64_t row_stride, void* y_bf16, clipppo_stream_t stream) {
    return clipppo::layernorm_launch(x, gamma, beta, rows, width, row_stride, y_bf16, clipppo::as_stream(stream));
}
