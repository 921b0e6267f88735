// Internal interface of the tcgen05 GEMM (gemm_tcgen05.cu) used by the tower driver (vit.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

namespace clipppo {

// TMA descriptor of a K-major bf16 matrix [rows, K] with leading dimension ld_elems; the box is
// {64 (K), box_rows} under the 128-byte swizzle, i.e. exactly one pipeline stage of the GEMM.
int make_bf16_kmajor_tmap(CUtensorMap* map, const void* ptr, int rows, int K, long long ld_elems, int box_rows);
int gemm_a_box_rows();
int gemm_b_box_rows();

int gemm_bf16_launch(const CUtensorMap& tmap_a, const CUtensorMap& tmap_w, int M, int N, int K, int epilogue,
                     const float* bias, const float* pos, int tokens, void* out, long long ldo, cudaStream_t stream,
                     const float* row_stats = nullptr, const float* colsum = nullptr, int stat_parts = 0,
                     float* stats_out = nullptr);

// tower building blocks (preprocess.cu / layernorm.cu / attention.cu)
int preprocess_launch(const void* images, int img_dtype, const long long strides[4], int N, int C, int h, int w,
                      float pre_scale, int normalize, int patch, int image, int kpad, void* patches_bf16, cudaStream_t stream);
// row_parts != null: also the (sum, sum of squares) of the rounded output rows, fp32 [rows, n_parts, 2] with everything in part 0
int layernorm_launch(const float* x, const float* gamma, const float* beta, int rows, int width,
                     long long row_stride, void* y_bf16, cudaStream_t stream, float* row_parts = nullptr, int n_parts = 0);
int layernorm_inplace_f32_launch(float* x, const float* gamma, const float* beta, int rows, int width,
                                 long long row_stride, cudaStream_t stream);
int layernorm_bf16in_launch(const void* x_bf16, const float* gamma, const float* beta, int rows, int width,
                            long long row_stride, void* y_bf16, cudaStream_t stream);
// n_parts == 0: stats = fp32 [rows, 2] (mean, rstd);  n_parts > 0: fp32 [rows, n_parts, 2] partial (sum, sum of squares), all in part 0
int rowstats_launch(const void* x_bf16, int rows, int width, long long row_stride, float* stats, cudaStream_t stream, int n_parts = 0);
int attention_launch(const void* qkv_bf16, int n_images, int tokens, int heads, int head_dim, void* out_bf16,
                     cudaStream_t stream, bool causal = false);

// tcgen05 / TMEM / TMA attention for 64 < T <= 257, no mask (attention_tc.cu)
bool attention_tc_supported(int tokens, bool causal);
int attention_tc_launch(const void* qkv_bf16, int n_images, int tokens, int heads, void* out_bf16, cudaStream_t stream);
// the same machinery for T <= 64, two images per 128-row tile (ViT-B/32: T = 50), and for causal T <= 128 (text tower: T = 77)
bool attention_tc_pair_supported(int tokens, bool causal);
int attention_tc_pair_launch(const void* qkv_bf16, int n_images, int tokens, int heads, void* out_bf16, cudaStream_t stream,
                             bool causal = false);

}  // namespace clipppo
