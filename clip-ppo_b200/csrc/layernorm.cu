// V2 - LayerNorm of the tower.
//
// Replaces [clip] LayerNorm (fp32 upcast -> F.layer_norm(eps=1e-5) -> downcast) as used for
// ln_pre / ln_1 / ln_2 / ln_post in VisionTransformer.forward.
//   layernorm_kernel  ln_pre (fp32 patch-embed rows -> bf16 residual stream) and ln_post (bf16 CLS rows):
//                     one warp per row, the row is held in registers (two-pass mean / biased variance
//                     like ATen), 128-bit loads, 64-bit bf16 stores.
//   rowstats_kernel   ln_1 / ln_2 are folded through the GEMM that consumes them (gemm_tcgen05.cu,
//                     ROWAFFINE epilogue); all that is left of them is the per-row (mean, rstd) of the
//                     bf16 residual stream: 1.5 KB read and 8 B written per row.
#include "common.cuh"
#include "gemm.cuh"

namespace clipppo {

namespace {

// width = NV * 128; lane owns float4 chunks lane, lane+32, ...
// OutT = __nv_bfloat16 (dense rows of `width`) or float (written back in place over x: ln_pre).
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// Per-row (mean, 1/sqrt(biased var + 1e-5)) of bf16 rows of width NV*256; lane owns 8-element pieces
// lane, lane+32, ...  Two passes over registers: exact mean first, then centred squares.
template <int NV>
__global__ void __launch_bounds__(256)
rowstats_kernel(const __nv_bfloat16* __restrict__ x, int rows, long long row_stride, float2* __restrict__ stats, int nparts) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int width = NV * 256;
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * row_stride);
    float2 v[NV][4];                                   // packed fp32 pairs: FADD2 / FFMA2 halve the arithmetic
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const uint4 u = xr[lane + 32 * i];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[i][t] = make_float2(__uint_as_float(w[t] << 16), __uint_as_float(w[t] & 0xffff0000u));     // bf16 -> fp32, exact
            s2 = __fadd2_rn(s2, v[i][t]);
        }
    }
    const float mean = warp_sum(s2.x + s2.y) * (1.0f / width);
    const float2 nm = make_float2(-mean, -mean);
    float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int t = 0; t < 4; ++t) { const float2 d = __fadd2_rn(v[i][t], nm); q2 = __ffma2_rn(d, d, q2); }
    }
    const float m2 = warp_sum(q2.x + q2.y);
    if (nparts > 0) {
        // partial-sum layout of the RESID_STATS epilogue: (sum, sum of squares) in slice 0, zeros in the others
        const float sm = mean * width;
        if (lane < nparts) stats[static_cast<size_t>(row) * nparts + lane] = lane == 0 ? make_float2(sm, fmaf(mean, sm, m2)) : make_float2(0.f, 0.f);
        return;
    }
    const float rstd = rsqrtf(m2 * (1.0f / width) + 1e-5f);
    if (lane == 0) stats[row] = make_float2(mean, rstd);
}

template <int NV, typename OutT, typename InT = float>
__global__ void __launch_bounds__(256)
layernorm_kernel(const InT* x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int rows, long long row_stride, OutT* y, float2* __restrict__ parts = nullptr, int nparts = 0) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int width = NV * 128;
    const InT* xr = x + static_cast<size_t>(row) * row_stride;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = load4(xr + 4 * (lane + 32 * i));
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
    float so = 0.f, qo = 0.f;                          // (sum, sum of squares) of the ROUNDED output row
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
        const float2 a = __bfloat1622float2(lo), c = __bfloat1622float2(hi);
        so += (a.x + a.y) + (c.x + c.y);
        qo += (a.x * a.x + a.y * a.y) + (c.x * c.x + c.y * c.y);
    }
    if constexpr (sizeof(OutT) == 2) {
        // the row statistics the first block's ln_1 needs, in the partial-sum layout of the RESID_STATS epilogue
        // (gemm_tcgen05.cu): everything in slice 0, zeros in the others
        if (parts != nullptr) {
            so = warp_sum(so); qo = warp_sum(qo);
            if (lane < nparts) parts[static_cast<size_t>(row) * nparts + lane] = lane == 0 ? make_float2(so, qo) : make_float2(0.f, 0.f);
        }
    }
}

}  // namespace

template <typename OutT>
static int layernorm_dispatch(const float* x, const float* gamma, const float* beta, int rows, int width,
                              long long row_stride, OutT* y, cudaStream_t stream, float2* parts = nullptr, int nparts = 0) {
    const int warps_per_block = 8;
    const unsigned grid = (rows + warps_per_block - 1) / warps_per_block;
    switch (width / 128) {
        case 4: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<4, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, parts, nparts)); break;
        case 6: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<6, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, parts, nparts)); break;
        case 8: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<8, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, parts, nparts)); break;
        case 10: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<10, OutT>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, parts, nparts)); break;
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
                     long long row_stride, void* y_bf16, cudaStream_t stream, float* row_parts, int n_parts) {
    if (!x || !gamma || !beta || !y_bf16) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || width <= 0 || n_parts < 0 || n_parts > 32) return CLIPPPO_ERR_BAD_SHAPE;
    if (width % 128 || (row_stride % 4) || (reinterpret_cast<uintptr_t>(x) % 16) ||
        (reinterpret_cast<uintptr_t>(y_bf16) % 8) || (reinterpret_cast<uintptr_t>(row_parts) % 8))
        return CLIPPPO_ERR_ALIGN;
    return layernorm_dispatch<__nv_bfloat16>(x, gamma, beta, rows, width, row_stride,
                                             static_cast<__nv_bfloat16*>(y_bf16), stream,
                                             reinterpret_cast<float2*>(row_parts), row_parts ? n_parts : 0);
}

// ln_post: bf16 rows (the CLS token of every image, row_stride = T*D) -> dense bf16 rows
int layernorm_bf16in_launch(const void* x_bf16, const float* gamma, const float* beta, int rows, int width,
                            long long row_stride, void* y_bf16, cudaStream_t stream) {
    if (!x_bf16 || !gamma || !beta || !y_bf16) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || width <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (width % 128 || (row_stride % 4) || (reinterpret_cast<uintptr_t>(x_bf16) % 8) || (reinterpret_cast<uintptr_t>(y_bf16) % 8))
        return CLIPPPO_ERR_ALIGN;
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_bf16);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
    const unsigned grid = (rows + 7) / 8;
    switch (width / 128) {
        case 4: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<4, __nv_bfloat16, __nv_bfloat16>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, static_cast<float2*>(nullptr), 0)); break;
        case 6: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<6, __nv_bfloat16, __nv_bfloat16>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, static_cast<float2*>(nullptr), 0)); break;
        case 8: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<8, __nv_bfloat16, __nv_bfloat16>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, static_cast<float2*>(nullptr), 0)); break;
        case 10: CLIPPPO_CUDA_TRY(launch_pdl(layernorm_kernel<10, __nv_bfloat16, __nv_bfloat16>, grid, 256, 0, stream, 1, x, gamma, beta, rows, row_stride, y, static_cast<float2*>(nullptr), 0)); break;
        default: return CLIPPPO_ERR_UNSUPPORTED;
    }
    prof_count_launch();
    return CLIPPPO_OK;
}

int rowstats_launch(const void* x_bf16, int rows, int width, long long row_stride, float* stats, cudaStream_t stream, int n_parts) {
    if (!x_bf16 || !stats) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || width <= 0 || n_parts < 0 || n_parts > 32) return CLIPPPO_ERR_BAD_SHAPE;
    if (width % 256 || (row_stride % 8) || (reinterpret_cast<uintptr_t>(x_bf16) % 16) || (reinterpret_cast<uintptr_t>(stats) % 8))
        return CLIPPPO_ERR_ALIGN;
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_bf16);
    float2* st = reinterpret_cast<float2*>(stats);
    const unsigned grid = (rows + 7) / 8;
    switch (width / 256) {
        case 2: CLIPPPO_CUDA_TRY(launch_pdl(rowstats_kernel<2>, grid, 256, 0, stream, 1, x, rows, row_stride, st, n_parts)); break;
        case 3: CLIPPPO_CUDA_TRY(launch_pdl(rowstats_kernel<3>, grid, 256, 0, stream, 1, x, rows, row_stride, st, n_parts)); break;
        case 4: CLIPPPO_CUDA_TRY(launch_pdl(rowstats_kernel<4>, grid, 256, 0, stream, 1, x, rows, row_stride, st, n_parts)); break;
        case 5: CLIPPPO_CUDA_TRY(launch_pdl(rowstats_kernel<5>, grid, 256, 0, stream, 1, x, rows, row_stride, st, n_parts)); break;
        default: return CLIPPPO_ERR_UNSUPPORTED;
    }
    prof_count_launch();
    return CLIPPPO_OK;
}

}  // namespace clipppo

extern "C" int clipppo_rowstats_bf16(const void* x_bf16, int rows, int width, int64_t row_stride, float* stats,
                                     clipppo_stream_t stream) {
    return clipppo::rowstats_launch(x_bf16, rows, width, row_stride, stats, clipppo::as_stream(stream), 0);
}

extern "C" int clipppo_layernorm_bf16(const float* x, const float* gamma, const float* beta, int rows,
                                      int width, int64_t row_stride, void* y_bf16, clipppo_stream_t stream) {
    return clipppo::layernorm_launch(x, gamma, beta, rows, width, row_stride, y_bf16, clipppo::as_stream(stream));
}
