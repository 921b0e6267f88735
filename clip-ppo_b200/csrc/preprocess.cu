// V0 - CLIP image preprocessing fused with im2col:  raw frames -> bf16 patch matrix.
//
// Replaces, in one pass, what reference shared/clip_ppo_utils.py:146-159 (and :201-210) does in
// four eager ops plus the fp16 cast and the unfold inside conv1:
//     u = raw * pre_scale ; bilinear resize to image x image (align_corners=False; antialias is
//     a no-op when up-sampling) ; (u - mean_c) / std_c ; cast ;
//     A[b*G*G + gy*G + gx, c*P*P + ky*P + kx] = pixel(b, c, gy*P+ky, gx*P+kx)
// so the 224x224 fp32 intermediate (79 GB at the Atari config) is never materialised.
// C == 1 inputs (Atari gray frames) are broadcast to the three channels in registers.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "gemm.cuh"

namespace clipppo {

namespace {

struct PreParams {
    const void* img;
    long long s[4];
    int N, C, h, w;
    int dtype;
    float pre_scale;
    int normalize;
    int P, G, image, kpad;
    float scale_h, scale_w;
    __nv_bfloat16* out;
};

__constant__ float kClipMean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
// 1 / (0.26862954, 0.26130258, 0.27577711): the reference divides by std in fp32; a reciprocal multiply differs by
// <= 1 ulp of fp32, far below the bf16 rounding of the patch matrix (8 fp32 divisions per thread were ~80 instructions)
__constant__ float kClipInvStd[3] = {1.0f / 0.26862954f, 1.0f / 0.26130258f, 1.0f / 0.27577711f};

__device__ __forceinline__ float load_px(const PreParams& p, long long off) {
    const float raw = (p.dtype == CLIPPPO_IMG_U8) ? static_cast<float>(static_cast<const uint8_t*>(p.img)[off])
                                                  : static_cast<const float*>(p.img)[off];
    return raw * p.pre_scale;
}

// ATen area_pixel_compute_source_index + guard_index_and_lambda (align_corners = false)
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
    float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    s = fmaxf(s, 0.0f);
    i0 = min(static_cast<int>(s), in_size - 1);
    i1 = i0 + ((i0 + 1 < in_size) ? 1 : 0);
    l1 = fminf(fmaxf(s - static_cast<float>(i0), 0.0f), 1.0f);
}

// Up-sampling path (h, w < image), patch % 8 == 0.  One CTA per (image, patch row gy), two passes:
//   1. horizontal: the ~P*h/image + 2 source rows this patch row touches are interpolated along x
//      into shared memory  tmp[cin][row][ox]  (x indices / weights are per-thread constants);
//   2. vertical + normalise + bf16 + im2col: each thread blends two smem rows for 8 consecutive ox
//      (4 x LDS.128) and writes one 16-byte chunk of the patch matrix.
// Same association as ATen's bilinear kernel: horizontal blends first, then the vertical blend.
__global__ void __launch_bounds__(256) preprocess_resize_kernel(const PreParams p, int max_rows) {
    extern __shared__ __align__(16) float tmp[];
    const int b = blockIdx.x / p.G, gy = blockIdx.x - b * p.G;
    const int P = p.P, G = p.G, IMG = p.image;
    const int Cin = p.C;                                       // 1 (gray, broadcast) or 3
    const long long img_off = static_cast<long long>(b) * p.s[0];
    int y_lo, y_hi, dummy_i; float dummy_f;
    src_index(p.scale_h, gy * P, p.h, y_lo, dummy_i, dummy_f);
    src_index(p.scale_h, gy * P + P - 1, p.h, dummy_i, y_hi, dummy_f);
    const int nrows = min(y_hi - y_lo + 1, max_rows);
    // ---- pass 1: thread = output column ----
    for (int ox = threadIdx.x; ox < IMG; ox += blockDim.x) {
        int x0, x1; float lx;
        src_index(p.scale_w, ox, p.w, x0, x1, lx);
        const float w0 = 1.0f - lx;
        for (int c = 0; c < Cin; ++c) {
            const long long cb = img_off + c * p.s[1];
            for (int r = 0; r < nrows; ++r) {
                const long long rb = cb + static_cast<long long>(y_lo + r) * p.s[2];
                const float a = load_px(p, rb + x0 * p.s[3]), d = load_px(p, rb + x1 * p.s[3]);
                tmp[(c * max_rows + r) * IMG + ox] = w0 * a + lx * d;
            }
        }
    }
    __syncthreads();
    // ---- pass 2: thread = (line (c, ky), 8 consecutive ox) ----
    const int groups = IMG >> 3;                               // 16-byte output chunks per image row
    const int total = 3 * P * groups;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int grp = i % groups, line = i / groups;
        const int c = line / P, ky = line - c * P;
        const int oy = gy * P + ky;
        int y0, y1; float ly;
        src_index(p.scale_h, oy, p.h, y0, y1, ly);
        const int cin = (Cin == 1) ? 0 : c;
        const float* r0p = tmp + (cin * max_rows + (y0 - y_lo)) * IMG + grp * 8;
        const float* r1p = tmp + (cin * max_rows + (y1 - y_lo)) * IMG + grp * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(r0p), a1 = *reinterpret_cast<const float4*>(r0p + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(r1p), b1 = *reinterpret_cast<const float4*>(r1p + 4);
        const float h0 = 1.0f - ly;
        const float mean = p.normalize ? kClipMean[c] : 0.0f, istd = p.normalize ? kClipInvStd[c] : 1.0f;
        float o[8] = {h0 * a0.x + ly * b0.x, h0 * a0.y + ly * b0.y, h0 * a0.z + ly * b0.z, h0 * a0.w + ly * b0.w,
                      h0 * a1.x + ly * b1.x, h0 * a1.y + ly * b1.y, h0 * a1.z + ly * b1.z, h0 * a1.w + ly * b1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = (o[t] - mean) * istd;
        const int ox0 = grp * 8, gx = ox0 / P, kx0 = ox0 - gx * P;
        __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * G * G + gy * G + gx) * p.kpad + (c * P + ky) * P + kx0;
        uint4 pk;
        __nv_bfloat162 q0 = __floats2bfloat162_rn(o[0], o[1]), q1 = __floats2bfloat162_rn(o[2], o[3]);
        __nv_bfloat162 q2 = __floats2bfloat162_rn(o[4], o[5]), q3 = __floats2bfloat162_rn(o[6], o[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
        pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
        *reinterpret_cast<uint4*>(dst) = pk;
    }
}

// 84 x 84 -> 224 x 224 with patch 32: the one up-sampling case of the reference (MiniGrid and Atari frames).
// The ratio is 3 / 8, so source index and weight of an output pixel depend only on (pixel mod 8):
//     x0 = 3g + {-1, 0, 0, 0, 1, 1, 1, 2}[j],  lambda = {11, 1, 7, 13, 3, 9, 15, 5}[j] / 16      (ox = 8g + j),
// the same numbers ATen's area_pixel_compute_source_index gives in fp32 (3/8 and the sixteenths are exact), and the
// same vertically.  Thread = (channel, 8-row group q of the patch row, 8-column group g): 5 source rows x 5 source
// columns from global memory (L1 serves the overlap with the neighbours), 5 x 8 horizontal blends, 8 x 8 vertical
// blends, normalise, eight 16-byte stores.  No shared memory, no barrier, ~7 instructions per output value
// (the generic two-pass kernel: 35, bound by its load/store pipe).
template <int DT>
__global__ void __launch_bounds__(352, 3) preprocess_up84_kernel(const PreParams p) {
    constexpr int W = 84, P = 32, G = 7;
    constexpr float kL[8] = {0.6875f, 0.0625f, 0.4375f, 0.8125f, 0.1875f, 0.5625f, 0.9375f, 0.3125f};
    constexpr int kI[8] = {0, 1, 1, 1, 2, 2, 2, 3};          // first of the two source lines, relative to (3 * group - 1)
    const int t = threadIdx.x;
    if (t >= 3 * 4 * 28) return;
    const int b = blockIdx.x / G, gy = blockIdx.x - b * G;
    const int g = t % 28, cq = t / 28, q = cq & 3, c = cq >> 2;
    using elem_t = typename std::conditional<DT == CLIPPPO_IMG_U8, uint8_t, float>::type;
    const elem_t* img = static_cast<const elem_t*>(p.img) + static_cast<long long>(b) * p.s[0] + ((p.C == 1) ? 0 : c) * p.s[1];
    const float ps = p.pre_scale;
    // source columns 3g-1 .. 3g+3 and source rows 12gy+3q-1 .. +3, clamped to the frame
    long long xo[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) xo[i] = static_cast<long long>(min(max(3 * g - 1 + i, 0), W - 1)) * p.s[3];
    const int y_first = 12 * gy + 3 * q - 1;
    float H[5][8];                                             // horizontally interpolated source rows
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const elem_t* row = img + static_cast<long long>(min(max(y_first + r, 0), W - 1)) * p.s[2];
        float v[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = static_cast<float>(__ldg(row + xo[i])) * ps;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float lx = (g == 0 && j == 0) ? 0.0f : kL[j];   // the left edge clamps the source coordinate to 0
            H[r][j] = (1.0f - lx) * v[kI[j]] + lx * v[kI[j] + 1];
        }
    }
    const float mean = p.normalize ? kClipMean[c] : 0.0f, istd = p.normalize ? kClipInvStd[c] : 1.0f;
    const int gx = g >> 2, kx0 = (g & 3) * 8;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * G * G + gy * G + gx) * p.kpad + (c * P + 8 * q) * P + kx0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {                              // output row ky = 8q + j
        const float ly = (gy == 0 && q == 0 && j == 0) ? 0.0f : kL[j];
        const float h0 = 1.0f - ly;
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float o0 = (h0 * H[kI[j]][2 * e] + ly * H[kI[j] + 1][2 * e] - mean) * istd;
            const float o1 = (h0 * H[kI[j]][2 * e + 1] + ly * H[kI[j] + 1][2 * e + 1] - mean) * istd;
            __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst + j * P) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// grid: (N * G) CTAs, one per (image, patch row gy); each thread produces VEC consecutive kx.
template <int VEC, bool IDENT>
__global__ void __launch_bounds__(256) preprocess_kernel(const PreParams p) {
    const int b = blockIdx.x / p.G, gy = blockIdx.x - b * p.G;
    const int P = p.P, G = p.G;
    const int vec_per_row = P / VEC;                       // vectors per (c, ky) line of one patch
    const int per_patch = 3 * P * vec_per_row;
    const int total = G * per_patch;
    const long long img_off = static_cast<long long>(b) * p.s[0];
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int gx = i / per_patch;
        int rem = i - gx * per_patch;
        const int c = rem / (P * vec_per_row);
        rem -= c * (P * vec_per_row);
        const int ky = rem / vec_per_row;
        const int kx0 = (rem - ky * vec_per_row) * VEC;
        const int oy = gy * P + ky;
        if constexpr (IDENT) {
            // h == w == image: the resize is the identity; 8 contiguous source pixels per thread
            const int cin = (p.C == 1) ? 0 : c;
            const long long off = img_off + cin * p.s[1] + static_cast<long long>(oy) * p.s[2] + (gx * P + kx0);
            const float mean = p.normalize ? kClipMean[c] : 0.0f, istd = p.normalize ? kClipInvStd[c] : 1.0f;
            float o[8];
            if (p.dtype == CLIPPPO_IMG_U8) {
                const uint2 raw = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(p.img) + off);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    o[t] = static_cast<float>((raw.x >> (8 * t)) & 0xffu);
                    o[4 + t] = static_cast<float>((raw.y >> (8 * t)) & 0xffu);
                }
            } else {
                const float4 a = ld_stream_f4(static_cast<const float*>(p.img) + off);
                const float4 b4 = ld_stream_f4(static_cast<const float*>(p.img) + off + 4);
                o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b4.x; o[5] = b4.y; o[6] = b4.z; o[7] = b4.w;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] = (o[t] * p.pre_scale - mean) * istd;
            __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * G * G + gy * G + gx) * p.kpad + (c * P + ky) * P + kx0;
            uint4 pk;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst) = pk;
            continue;
        }
        int y0, y1; float ly;
        src_index(p.scale_h, oy, p.h, y0, y1, ly);
        const int cin = (p.C == 1) ? 0 : c;
        const long long base0 = img_off + cin * p.s[1] + static_cast<long long>(y0) * p.s[2];
        const long long base1 = img_off + cin * p.s[1] + static_cast<long long>(y1) * p.s[2];
        const float mean = p.normalize ? kClipMean[c] : 0.0f, istd = p.normalize ? kClipInvStd[c] : 1.0f;
        float o[VEC];
#pragma unroll
        for (int t = 0; t < VEC; ++t) {
            const int ox = gx * P + kx0 + t;
            int x0, x1; float lx;
            src_index(p.scale_w, ox, p.w, x0, x1, lx);
            const float p00 = load_px(p, base0 + x0 * p.s[3]), p01 = load_px(p, base0 + x1 * p.s[3]);
            const float p10 = load_px(p, base1 + x0 * p.s[3]), p11 = load_px(p, base1 + x1 * p.s[3]);
            const float v = (1.0f - ly) * ((1.0f - lx) * p00 + lx * p01) + ly * ((1.0f - lx) * p10 + lx * p11);
            o[t] = (v - mean) * istd;
        }
        __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * G * G + gy * G + gx) * p.kpad + (c * P + ky) * P + kx0;
        if constexpr (VEC == 8) {
            uint4 pk;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst) = pk;
        } else {
#pragma unroll
            for (int t = 0; t < VEC; t += 2) *reinterpret_cast<__nv_bfloat162*>(dst + t) = __floats2bfloat162_rn(o[t], o[t + 1]);
        }
    }
    // zero the K padding (only when 3*P*P is not a multiple of 64, e.g. patch 14)
    const int kreal = 3 * P * P;
    if (p.kpad > kreal) {
        const int padn = p.kpad - kreal;
        for (int i = threadIdx.x; i < G * padn; i += blockDim.x) {
            const int gx = i / padn, t = i - gx * padn;
            p.out[(static_cast<size_t>(b) * G * G + gy * G + gx) * p.kpad + kreal + t] = __float2bfloat16(0.0f);
        }
    }
}

}  // namespace

int preprocess_launch(const void* images, int img_dtype, const long long strides[4], int N, int C, int h, int w,
                      float pre_scale, int normalize, int patch, int image, int kpad, void* patches_bf16,
                      cudaStream_t stream) {
    if (!images || !patches_bf16) return CLIPPPO_ERR_NULL;
    if (N <= 0 || h <= 0 || w <= 0 || patch <= 0 || image <= 0 || image % patch) return CLIPPPO_ERR_BAD_SHAPE;
    if (C != 1 && C != 3) return CLIPPPO_ERR_BAD_CHANNELS;
    if (img_dtype != CLIPPPO_IMG_F32 && img_dtype != CLIPPPO_IMG_U8) return CLIPPPO_ERR_UNSUPPORTED;
    if (h > image || w > image) return CLIPPPO_ERR_UNSUPPORTED;   // down-sampling needs the antialias filter
    if (patch % 2 || kpad < 3 * patch * patch || (kpad % 8)) return CLIPPPO_ERR_UNSUPPORTED;
    PreParams p;
    p.img = images;
    for (int i = 0; i < 4; ++i) p.s[i] = strides[i];
    p.N = N; p.C = C; p.h = h; p.w = w; p.dtype = img_dtype; p.pre_scale = pre_scale; p.normalize = normalize;
    p.P = patch; p.G = image / patch; p.image = image; p.kpad = kpad;
    p.scale_h = static_cast<float>(h) / static_cast<float>(image);
    p.scale_w = static_cast<float>(w) / static_cast<float>(image);
    p.out = static_cast<__nv_bfloat16*>(patches_bf16);
    const unsigned grid = static_cast<unsigned>(N) * p.G;
    const bool vec8 = patch % 8 == 0 && (reinterpret_cast<uintptr_t>(patches_bf16) % 16 == 0);
    const long long al = (img_dtype == CLIPPPO_IMG_U8) ? 8 : 4;      // elements per 8 / 16 source bytes
    const bool ident = vec8 && h == image && w == image && strides[3] == 1 && strides[0] % al == 0 &&
                       strides[1] % al == 0 && strides[2] % al == 0 &&
                       reinterpret_cast<uintptr_t>(images) % (img_dtype == CLIPPPO_IMG_U8 ? 8 : 16) == 0;
    const int max_rows = (patch * h + image - 1) / image + 2;
    const size_t rs_smem = static_cast<size_t>(C) * max_rows * image * sizeof(float);
    if (ident) {
        preprocess_kernel<8, true><<<grid, 256, 0, stream>>>(p);
    } else if (vec8 && h == 84 && w == 84 && image == 224 && patch == 32 && !getenv("CLIPPPO_PREPROCESS_GENERIC")) {
        if (img_dtype == CLIPPPO_IMG_U8) preprocess_up84_kernel<CLIPPPO_IMG_U8><<<grid, 352, 0, stream>>>(p);
        else preprocess_up84_kernel<CLIPPPO_IMG_F32><<<grid, 352, 0, stream>>>(p);
    } else if (vec8 && image % 8 == 0 && rs_smem <= 200 * 1024) {
        static DeviceOnce configured;
        if (rs_smem > 48 * 1024 && configured.first_use())
            CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(preprocess_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        preprocess_resize_kernel<<<grid, 256, rs_smem, stream>>>(p, max_rows);
    } else if (vec8)
        preprocess_kernel<8, false><<<grid, 256, 0, stream>>>(p);
    else
        preprocess_kernel<2, false><<<grid, 256, 0, stream>>>(p);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

}  // namespace clipppo

using namespace clipppo;

extern "C" int clipppo_preprocess_bf16(const void* images, int img_dtype, const int64_t img_strides_host[4],
                                       int N, int C, int h, int w, float pre_scale, int normalize,
                                       int patch, int image, void* patches_bf16, clipppo_stream_t stream) {
    if (N <= 0 || C <= 0 || h <= 0 || w <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    long long s[4] = {static_cast<long long>(C) * h * w, static_cast<long long>(h) * w, w, 1};
    if (img_strides_host) for (int i = 0; i < 4; ++i) s[i] = img_strides_host[i];
    const int kreal = 3 * patch * patch;
    const int kpad = (kreal + 63) / 64 * 64;
    return preprocess_launch(images, img_dtype, s, N, C, h, w, pre_scale, normalize, patch, image, kpad, patches_bf16,
                             as_stream(stream));
}
