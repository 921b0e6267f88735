// Frozen CLIP image tower driver: owns the weight TMA descriptors and sequences the kernels
//   V0 preprocess+im2col -> V1 patch-embed GEMM (+pos) -> ln_pre (+ row statistics) ->
//   L x [ QKV GEMM (ln_1 folded) -> attention -> out-proj GEMM (+= residual, + row statistics) ->
//         c_fc GEMM (ln_2 folded, +QuickGELU) -> c_proj GEMM (+= residual, + row statistics) ]
//   -> ln_post(CLS) -> head GEMM -> optional L2 normalise
// Replaces [clip] VisionTransformer.forward / CLIP.encode_image as called by reference
// shared/clip_ppo_utils.py:163-164 and :212-217 (spec: SURVEY.md Appendix B).
//
// The residual stream X is bf16 [n*T, D] and is the A operand of the QKV / c_fc GEMMs as it stands:
// ln_1 / ln_2 are folded through those GEMMs (gamma in the weights, beta in the bias, mean / rstd
// applied per row in the epilogue), and the row statistics they need are a by-product of the epilogue
// that WROTE the rows (RESID_STATS: x_old by TMA load, fp32 add, one rounding, TMA store, per-row partial
// sums from the registers), so a block is five launches:  QKV -> attention -> out_proj(+=) -> c_fc(+GELU) ->
// c_proj(+=).  The round-1 schedule (residual updates as TMA reduce-adds in L2 + a rowstats pass in front
// of each folded GEMM) is kept for the opt-in split-K latency schedule and for A/B measurements
// (CLIPPPO_GEMM_RESID=reduce).
// Accumulation is fp32 everywhere; patch embedding + ln_pre run in fp32.  Images are processed in
// chunks so the activation workspace stays bounded whatever N is; buffers that are never live
// together (im2col patches / QKV / MLP hidden) share one allocation.
#include <limits.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"

// one residual block's repacked weights ([clip] ResidualAttentionBlock; the same in both towers)
struct clipppo_tower_layer {
    const __nv_bfloat16 *w_qkv, *w_out, *w_fc, *w_proj;     // w_qkv / w_fc carry ln_1 / ln_2 gamma
    const float *b_qkv, *s_qkv, *b_out, *b_fc, *s_fc, *b_proj;
    CUtensorMap tm_qkv, tm_out, tm_fc, tm_proj;
};

struct clipppo_vit_s {
    clipppo_vit_config cfg;
    int tokens, grid, kpatch;
    // handle-owned device arena with the repacked frozen weights
    void* arena = nullptr;
    typedef clipppo_tower_layer Layer;
    std::vector<Layer> layers;
    const __nv_bfloat16 *w_patch, *w_head;
    const float *cls_pos0, *pos, *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
    CUtensorMap tm_patch, tm_head;
};

// Frozen CLIP text tower ([clip] CLIP.encode_text): token + positional embedding, the same residual blocks
// with a causal attention mask, ln_final on the EOT row, text_projection.
struct clipppo_text_s {
    clipppo_text_config cfg;
    void* arena = nullptr;
    std::vector<clipppo_tower_layer> layers;
    const float *tok_emb, *pos, *ln_final_g, *ln_final_b;
    const __nv_bfloat16* w_head;
    CUtensorMap tm_head;
};

namespace clipppo {

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- one-off weight repacking (clipppo_vit_create) ------------------------------------------
// fp32 [rows, k_src] -> bf16 [rows, k_dst], zero-padded columns (conv1 of ViT-L/14: 588 -> 640)
__global__ void convert_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int k_src, int k_dst) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(rows) * k_dst) return;
    const int r = static_cast<int>(i / k_dst), k = static_cast<int>(i - static_cast<size_t>(r) * k_dst);
    dst[i] = __float2bfloat16_rn(k < k_src ? src[static_cast<size_t>(r) * k_src + k] : 0.0f);
}
// proj fp32 [D, O] -> bf16 [O, D] (K-major B operand of the head GEMM)
__global__ void transpose_convert_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int D, int O) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * O) return;
    const int o = i / D, d = i - o * D;
    dst[i] = __float2bfloat16_rn(src[static_cast<size_t>(d) * O + o]);
}
__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
// LayerNorm folded into the linear layer that consumes it ([clip] ln_1 -> attn.in_proj, ln_2 -> mlp.c_fc):
//   LN(x) W^T + b = rstd * (x W'^T - mean * s) + b'   with  W' = W diag(gamma),  s = W' 1,  b' = b + W beta.
// s is summed over the bf16-ROUNDED W' - the values the tensor core multiplies - so the mean term
// cancels exactly what the GEMM accumulated.  One block per output row.
__global__ void __launch_bounds__(128)
fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ bias, int K, __nv_bfloat16* __restrict__ Wf, float* __restrict__ colsum,
               float* __restrict__ bias2) {
    __shared__ float red[64];
    const int n = blockIdx.x;
    const float* w = W + static_cast<size_t>(n) * K;
    float s = 0.f, bb = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float wv = w[k];
        const __nv_bfloat16 h = __float2bfloat16_rn(wv * gamma[k]);
        Wf[static_cast<size_t>(n) * K + k] = h;
        s += __bfloat162float(h);
        bb = fmaf(wv, beta[k], bb);
    }
    s = block_sum(s, red);
    bb = block_sum(bb, red + 32);
    if (threadIdx.x == 0) { colsum[n] = s; bias2[n] = bias[n] + bb; }
}

struct Workspace {
    float* X0;                // [n*T, D] fp32: patch embedding + positional embedding, input of ln_pre
    __nv_bfloat16* X;         // [n*T, D] bf16 residual stream (A operand of the QKV / c_fc GEMMs as is)
    float* stats;             // [n*T, D/128, 2] partial (sum, sum of squares) of the rows of X  (legacy schedule: [n*T, 2] (mean, rstd))
    __nv_bfloat16* Y;         // [n*T, D] attention output
    __nv_bfloat16* H;         // max(patches [n*G*G, kpatch], QKV [n*T, 3D], hidden [n*T, 4D])
    __nv_bfloat16* Ycls;      // [n, D]
    size_t bytes;
};

Workspace carve(const clipppo_vit_s* h, int n, void* base) {
    const size_t T = h->tokens, D = h->cfg.width;
    const size_t rows = static_cast<size_t>(n) * T;
    size_t hcols = 4 * D;
    const size_t patch_elems = static_cast<size_t>(n) * h->grid * h->grid * h->kpatch;
    size_t h_elems = rows * hcols;
    if (patch_elems > h_elems) h_elems = patch_elems;
    Workspace ws;
    size_t off = 0;
    uint8_t* b = static_cast<uint8_t*>(base);
    ws.X0 = reinterpret_cast<float*>(b + off);           off += align_up(rows * D * 4, 1024);
    ws.X = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(rows * D * 2, 1024);
    ws.stats = reinterpret_cast<float*>(b + off);        off += align_up(rows * (D / 128) * 2 * 4, 1024);
    ws.Y = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(rows * D * 2, 1024);
    ws.H = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(h_elems * 2, 1024);
    ws.Ycls = reinterpret_cast<__nv_bfloat16*>(b + off); off += align_up(static_cast<size_t>(n) * D * 2, 1024);
    ws.bytes = off;
    return ws;
}

// X0[img*T + 0, :] = class_embedding + positional_embedding[0]
__global__ void cls_init_kernel(float* __restrict__ X, const float* __restrict__ cls_pos0, int n, int T, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
    const int d4 = D >> 2;
    if (i >= n * d4) return;
    const int img = i / d4, c = i - img * d4;
    reinterpret_cast<float4*>(X + static_cast<size_t>(img) * T * D)[c] = __ldg(reinterpret_cast<const float4*>(cls_pos0) + c);
}

// rows of `out` scaled to unit L2 norm (F.normalize, eps 1e-12); one warp per row
__global__ void l2norm_rows_kernel(float* __restrict__ out, int rows, int dim) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* r = out + static_cast<size_t>(row) * dim;
    float s = 0.f;
    for (int i = lane; i < dim; i += 32) s += r[i] * r[i];
    s = warp_sum(s);
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    for (int i = lane; i < dim; i += 32) r[i] *= inv;
}

#define VIT_TRY(expr) do { int st__ = (expr); if (st__ != CLIPPPO_OK) return st__; } while (0)

// Residual schedule: fused row statistics (default) unless the split-K latency schedule or the round-1 reduce-add
// path is asked for by environment.
bool fused_stats() {                                   // read per tower pass (two getenv calls), so tests can switch it
    const char* r = getenv("CLIPPPO_GEMM_RESID");
    const char* k = getenv("CLIPPPO_GEMM_KSPLIT");
    return !((r && r[0] == 'r') || (k && k[0]));
}

// L x [ x += out_proj(attn(ln_1(x)));  x += c_proj(QuickGELU(c_fc(ln_2(x)))) ] on the bf16 residual stream X [rows, D]
// fused = true: `stats` arrives holding the partial sums of the rows of X (written by ln_pre / rowstats) and every
// residual GEMM refreshes them.
// cls_last (image tower, fused schedule only): the last block's out_proj / c_fc / c_proj run on the n class-token rows alone -
// row 0 of every image, addressed through tensor maps with a row stride of T * D - because ln_post reads nothing else.
int run_blocks(const std::vector<clipppo_tower_layer>& layers, int n, int T, int D, int heads, bool causal,
               __nv_bfloat16* X, float* stats, __nv_bfloat16* Y, __nv_bfloat16* H, cudaStream_t stream, bool fused,
               bool cls_last = false) {
    const int rows = n * T;
    CUtensorMap tmX, tmY, tmH;
    VIT_TRY(make_bf16_kmajor_tmap(&tmX, X, rows, D, D, gemm_a_box_rows()));
    VIT_TRY(make_bf16_kmajor_tmap(&tmY, Y, rows, D, D, gemm_a_box_rows()));
    VIT_TRY(make_bf16_kmajor_tmap(&tmH, H, rows, 4 * D, 4 * D, gemm_a_box_rows()));
    if (fused) {
        const int P = D / 128;
        for (size_t li = 0; li < layers.size(); ++li) {
            const clipppo_tower_layer& w = layers[li];
            VIT_TRY(gemm_bf16_launch(tmX, w.tm_qkv, rows, 3 * D, D, CLIPPPO_EPI_ROWAFFINE_BF16, w.b_qkv, nullptr, 0,
                                     H, 3 * D, stream, stats, w.s_qkv, P));
            VIT_TRY(attention_launch(H, n, T, heads, D / heads, Y, stream, causal));
            if (cls_last && li + 1 == layers.size()) {
                // every key / value of the block was needed (attention above), but only the class token's query row is read
                // from here on: M = n rows with a row stride of T * D; H (the QKV matrix, consumed) takes the [n, 4 D] hidden rows;
                // the statistics of the n rows use the first n entries of `stats`
                const long long ldx = static_cast<long long>(T) * D;
                CUtensorMap tmYc, tmXc, tmHc;
                VIT_TRY(make_bf16_kmajor_tmap(&tmYc, Y, n, D, ldx, gemm_a_box_rows()));
                VIT_TRY(make_bf16_kmajor_tmap(&tmXc, X, n, D, ldx, gemm_a_box_rows()));
                VIT_TRY(make_bf16_kmajor_tmap(&tmHc, H, n, 4 * D, 4 * D, gemm_a_box_rows()));
                VIT_TRY(gemm_bf16_launch(tmYc, w.tm_out, n, D, D, CLIPPPO_EPI_RESID_STATS_BF16, w.b_out, nullptr, 0, X, ldx, stream,
                                         nullptr, nullptr, 0, stats));
                VIT_TRY(gemm_bf16_launch(tmXc, w.tm_fc, n, 4 * D, D, CLIPPPO_EPI_ROWAFFINE_GELU_BF16, w.b_fc, nullptr, 0,
                                         H, 4 * D, stream, stats, w.s_fc, P));
                VIT_TRY(gemm_bf16_launch(tmHc, w.tm_proj, n, D, 4 * D, CLIPPPO_EPI_RESID_STATS_BF16, w.b_proj, nullptr, 0, X, ldx, stream,
                                         nullptr, nullptr, 0, stats));
                break;
            }
            VIT_TRY(gemm_bf16_launch(tmY, w.tm_out, rows, D, D, CLIPPPO_EPI_RESID_STATS_BF16, w.b_out, nullptr, 0, X, D, stream,
                                     nullptr, nullptr, 0, stats));
            VIT_TRY(gemm_bf16_launch(tmX, w.tm_fc, rows, 4 * D, D, CLIPPPO_EPI_ROWAFFINE_GELU_BF16, w.b_fc, nullptr, 0,
                                     H, 4 * D, stream, stats, w.s_fc, P));
            VIT_TRY(gemm_bf16_launch(tmH, w.tm_proj, rows, D, 4 * D, CLIPPPO_EPI_RESID_STATS_BF16, w.b_proj, nullptr, 0, X, D, stream,
                                     nullptr, nullptr, 0, stats));
        }
        return CLIPPPO_OK;
    }
    for (const clipppo_tower_layer& w : layers) {
        // ln_1 lives in the QKV GEMM's epilogue
        VIT_TRY(rowstats_launch(X, rows, D, D, stats, stream));
        VIT_TRY(gemm_bf16_launch(tmX, w.tm_qkv, rows, 3 * D, D, CLIPPPO_EPI_ROWAFFINE_BF16, w.b_qkv, nullptr, 0,
                                 H, 3 * D, stream, stats, w.s_qkv));
        VIT_TRY(attention_launch(H, n, T, heads, D / heads, Y, stream, causal));
        VIT_TRY(gemm_bf16_launch(tmY, w.tm_out, rows, D, D, CLIPPPO_EPI_RESID_BF16, w.b_out, nullptr, 0, X, D, stream));
        VIT_TRY(rowstats_launch(X, rows, D, D, stats, stream));
        VIT_TRY(gemm_bf16_launch(tmX, w.tm_fc, rows, 4 * D, D, CLIPPPO_EPI_ROWAFFINE_GELU_BF16, w.b_fc, nullptr, 0,
                                 H, 4 * D, stream, stats, w.s_fc));
        VIT_TRY(gemm_bf16_launch(tmH, w.tm_proj, rows, D, 4 * D, CLIPPPO_EPI_RESID_BF16, w.b_proj, nullptr, 0, X, D, stream));
    }
    return CLIPPPO_OK;
}

// ---- text tower front / back end -------------------------------------------------------------
// X[n*T + t, :] = bf16(token_embedding[tokens[n, t]] + positional_embedding[t]); one warp per row, ids clamped to the table
__global__ void __launch_bounds__(256)
text_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ emb, const float* __restrict__ pos,
                  __nv_bfloat16* __restrict__ X, int rows, int T, int D, int vocab) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int t = row % T;
    int id = __ldg(tokens + row);
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4* e = reinterpret_cast<const float4*>(emb + static_cast<size_t>(id) * D);
    const float4* p = reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * D);
    uint2* x = reinterpret_cast<uint2*>(X + static_cast<size_t>(row) * D);
    for (int i = lane; i < D / 4; i += 32) {
        const float4 a = __ldg(e + i), b = __ldg(p + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(a.x + b.x, a.y + b.y), hi = __floats2bfloat162_rn(a.z + b.z, a.w + b.w);
        x[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
}
// Xe[n, :] = X[n*T + argmax_t tokens[n, t], :]  (first maximum, like torch.argmax; the EOT token has the largest id)
__global__ void __launch_bounds__(256)
text_gather_eot_kernel(const int* __restrict__ tokens, const __nv_bfloat16* __restrict__ X, __nv_bfloat16* __restrict__ Xe,
                       int n, int T, int D) {
    const int seq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (seq >= n) return;
    int best = INT_MIN, at = 0;
    for (int t = lane; t < T; t += 32) {
        const int v = __ldg(tokens + static_cast<size_t>(seq) * T + t);
        if (v > best) { best = v; at = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), oa = __shfl_xor_sync(0xffffffffu, at, o);
        if (ob > best || (ob == best && oa < at)) { best = ob; at = oa; }
    }
    const uint4* src = reinterpret_cast<const uint4*>(X + (static_cast<size_t>(seq) * T + at) * D);
    uint4* dst = reinterpret_cast<uint4*>(Xe + static_cast<size_t>(seq) * D);
    for (int i = lane; i < D / 8; i += 32) dst[i] = src[i];
}

struct TextWorkspace {
    __nv_bfloat16 *X, *Y, *H, *Xe, *Ycls;
    float* stats;
    size_t bytes;
};

TextWorkspace carve_text(const clipppo_text_s* h, int n, void* base) {
    const size_t T = h->cfg.context, D = h->cfg.width, rows = static_cast<size_t>(n) * T;
    TextWorkspace ws;
    size_t off = 0;
    uint8_t* b = static_cast<uint8_t*>(base);
    ws.X = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(rows * D * 2, 1024);
    ws.stats = reinterpret_cast<float*>(b + off);        off += align_up(rows * (D / 128) * 2 * 4, 1024);
    ws.Y = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(rows * D * 2, 1024);
    ws.H = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(rows * 4 * D * 2, 1024);
    ws.Xe = reinterpret_cast<__nv_bfloat16*>(b + off);   off += align_up(static_cast<size_t>(n) * D * 2, 1024);
    ws.Ycls = reinterpret_cast<__nv_bfloat16*>(b + off); off += align_up(static_cast<size_t>(n) * D * 2, 1024);
    ws.bytes = off;
    return ws;
}

int encode_text_chunk(const clipppo_text_s* h, const int* tokens, int n, int flags, float* out, const TextWorkspace& ws,
                      cudaStream_t stream) {
    const int D = h->cfg.width, T = h->cfg.context, O = h->cfg.out_dim, rows = n * T;
    text_embed_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(tokens, h->tok_emb, h->pos, ws.X, rows, T, D, h->cfg.vocab);
    CLIPPPO_CHECK_LAUNCH();
    const bool fused = fused_stats();
    if (fused) VIT_TRY(rowstats_launch(ws.X, rows, D, D, ws.stats, stream, D / 128));
    VIT_TRY(run_blocks(h->layers, n, T, D, h->cfg.heads, true, ws.X, ws.stats, ws.Y, ws.H, stream, fused));
    // ln_final is row-wise, so normalising only the EOT rows equals normalising everything and selecting them
    text_gather_eot_kernel<<<(n + 7) / 8, 256, 0, stream>>>(tokens, ws.X, ws.Xe, n, T, D);
    CLIPPPO_CHECK_LAUNCH();
    VIT_TRY(layernorm_bf16in_launch(ws.Xe, h->ln_final_g, h->ln_final_b, n, D, D, ws.Ycls, stream));
    CUtensorMap tmC;
    VIT_TRY(make_bf16_kmajor_tmap(&tmC, ws.Ycls, n, D, D, gemm_a_box_rows()));
    VIT_TRY(gemm_bf16_launch(tmC, h->tm_head, n, O, D, CLIPPPO_EPI_F32, nullptr, nullptr, 0, out, O, stream));
    if (flags & CLIPPPO_VIT_L2NORM) {
        l2norm_rows_kernel<<<(n + 7) / 8, 256, 0, stream>>>(out, n, O);
        CLIPPPO_CHECK_LAUNCH();
    }
    return CLIPPPO_OK;
}

int encode_chunk(const clipppo_vit_s* h, const void* images, int img_dtype, const long long strides[4], int n, int C,
                 int ih, int iw, float pre_scale, int flags, float* out, const Workspace& ws, cudaStream_t stream) {
    const int D = h->cfg.width, T = h->tokens, G = h->grid, O = h->cfg.out_dim;
    const int rows = n * T, prow = n * G * G;
    CUtensorMap tmP, tmC;
    VIT_TRY(make_bf16_kmajor_tmap(&tmP, ws.H, prow, h->kpatch, h->kpatch, gemm_a_box_rows()));
    VIT_TRY(make_bf16_kmajor_tmap(&tmC, ws.Ycls, n, D, D, gemm_a_box_rows()));

    VIT_TRY(preprocess_launch(images, img_dtype, strides, n, C, ih, iw, pre_scale,
                              (flags & CLIPPPO_VIT_PRENORMALIZED) ? 0 : 1, h->cfg.patch, h->cfg.image,
                              h->kpatch, ws.H, stream));
    cls_init_kernel<<<(n * (D / 4) + 255) / 256, 256, 0, stream>>>(ws.X0, h->cls_pos0, n, T, D);
    CLIPPPO_CHECK_LAUNCH();
    VIT_TRY(gemm_bf16_launch(tmP, h->tm_patch, prow, D, h->kpatch, CLIPPPO_EPI_PATCH_F32, nullptr, h->pos, T,
                             ws.X0, D, stream));
    const bool fused = fused_stats();
    VIT_TRY(layernorm_launch(ws.X0, h->ln_pre_g, h->ln_pre_b, rows, D, D, ws.X, stream,                // ln_pre -> bf16 residual
                             fused ? ws.stats : nullptr, D / 128));                                   //  (+ the statistics block 0's ln_1 needs)
    VIT_TRY(run_blocks(h->layers, n, T, D, h->cfg.heads, false, ws.X, ws.stats, ws.Y, ws.H, stream, fused,
                       fused && (flags & CLIPPPO_VIT_CLS_LAST_BLOCK) != 0));
    VIT_TRY(layernorm_bf16in_launch(ws.X, h->ln_post_g, h->ln_post_b, n, D, static_cast<long long>(T) * D, ws.Ycls, stream));
    VIT_TRY(gemm_bf16_launch(tmC, h->tm_head, n, O, D, CLIPPPO_EPI_F32, nullptr, nullptr, 0, out, O, stream));
    if (flags & CLIPPPO_VIT_L2NORM) {
        l2norm_rows_kernel<<<(n + 7) / 8, 256, 0, stream>>>(out, n, O);
        CLIPPPO_CHECK_LAUNCH();
    }
    return CLIPPPO_OK;
}

// ---- residual-block weights: arena planning, repacking, TMA descriptors (both towers) --------
struct LayerOff { size_t qkv, out, fc, proj, vec; };

template <class Take>
void plan_layers(std::vector<LayerOff>& lo, int D, Take&& take) {
    const size_t vec_floats = 3 * D + 3 * D + D + 4 * D + 4 * D + D;      // b_qkv s_qkv b_out b_fc s_fc b_proj
    for (LayerOff& o : lo) {
        o.qkv = take(static_cast<size_t>(3) * D * D * 2);
        o.out = take(static_cast<size_t>(D) * D * 2);
        o.fc = take(static_cast<size_t>(4) * D * D * 2);
        o.proj = take(static_cast<size_t>(4) * D * D * 2);
        o.vec = take(vec_floats * 4);
    }
}

bool layers_complete(const clipppo_vit_layer* src, int L) {
    for (int l = 0; l < L; ++l) {
        const clipppo_vit_layer& lw = src[l];
        if (!lw.w_qkv || !lw.b_qkv || !lw.w_out || !lw.b_out || !lw.w_fc || !lw.b_fc || !lw.w_proj || !lw.b_proj ||
            !lw.ln1_g || !lw.ln1_b || !lw.ln2_g || !lw.ln2_b)
            return false;
    }
    return true;
}

void repack_layers(const clipppo_vit_layer* src, std::vector<clipppo_tower_layer>& dst, const std::vector<LayerOff>& lo,
                   uint8_t* A, int D, cudaStream_t s0) {
    auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(A + o); };
    auto blocks = [](size_t n) { return static_cast<unsigned>((n + 255) / 256); };
    for (size_t l = 0; l < dst.size(); ++l) {
        const clipppo_vit_layer& lw = src[l];
        clipppo_tower_layer& hl = dst[l];
        float* v = reinterpret_cast<float*>(A + lo[l].vec);
        float *b_qkv = v, *s_qkv = v + 3 * D, *b_out = v + 6 * D, *b_fc = v + 7 * D, *s_fc = v + 11 * D, *b_proj = v + 15 * D;
        fold_ln_kernel<<<3 * D, 128, 0, s0>>>(lw.w_qkv, lw.ln1_g, lw.ln1_b, lw.b_qkv, D, bf(lo[l].qkv), s_qkv, b_qkv);
        fold_ln_kernel<<<4 * D, 128, 0, s0>>>(lw.w_fc, lw.ln2_g, lw.ln2_b, lw.b_fc, D, bf(lo[l].fc), s_fc, b_fc);
        convert_pad_kernel<<<blocks(static_cast<size_t>(D) * D), 256, 0, s0>>>(lw.w_out, bf(lo[l].out), D, D, D);
        convert_pad_kernel<<<blocks(static_cast<size_t>(4) * D * D), 256, 0, s0>>>(lw.w_proj, bf(lo[l].proj), D, 4 * D, 4 * D);
        cudaMemcpyAsync(b_out, lw.b_out, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s0);
        cudaMemcpyAsync(b_proj, lw.b_proj, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s0);
        hl.w_qkv = bf(lo[l].qkv); hl.w_out = bf(lo[l].out); hl.w_fc = bf(lo[l].fc); hl.w_proj = bf(lo[l].proj);
        hl.b_qkv = b_qkv; hl.s_qkv = s_qkv; hl.b_out = b_out; hl.b_fc = b_fc; hl.s_fc = s_fc; hl.b_proj = b_proj;
    }
}

int map_layers(std::vector<clipppo_tower_layer>& layers, int D) {
    const int nb = gemm_b_box_rows();
    for (clipppo_tower_layer& hl : layers) {
        VIT_TRY(make_bf16_kmajor_tmap(&hl.tm_qkv, hl.w_qkv, 3 * D, D, D, nb));
        VIT_TRY(make_bf16_kmajor_tmap(&hl.tm_out, hl.w_out, D, D, D, nb));
        VIT_TRY(make_bf16_kmajor_tmap(&hl.tm_fc, hl.w_fc, 4 * D, D, D, nb));
        VIT_TRY(make_bf16_kmajor_tmap(&hl.tm_proj, hl.w_proj, D, 4 * D, 4 * D, nb));
    }
    return CLIPPPO_OK;
}

}  // namespace
}  // namespace clipppo

using namespace clipppo;

// Repack the caller's fp32 openai/CLIP weights into the handle-owned arena (bf16 K-major GEMM
// operands, LayerNorm-folded in_proj / c_fc, fp32 vectors) and build the weight TMA descriptors.
extern "C" int clipppo_vit_create(clipppo_vit_t* handle, const clipppo_vit_config* cfg,
                                  const clipppo_vit_weights* weights_host) {
    if (!handle || !cfg || !weights_host || !weights_host->layers_host) return CLIPPPO_ERR_NULL;
    const int D = cfg->width, P = cfg->patch, O = cfg->out_dim, L = cfg->layers;
    if (D <= 0 || L <= 0 || cfg->heads <= 0 || P <= 0 || cfg->image <= 0 || O <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (cfg->image % P || D % cfg->heads) return CLIPPPO_ERR_BAD_SHAPE;
    if (D / cfg->heads != 64 || D % 256 || O % 32) return CLIPPPO_ERR_UNSUPPORTED;
    const clipppo_vit_weights& w = *weights_host;
    if (!w.conv1 || !w.class_embedding || !w.positional_embedding || !w.ln_pre_g || !w.ln_pre_b || !w.ln_post_g ||
        !w.ln_post_b || !w.proj)
        return CLIPPPO_ERR_NULL;
    if (!layers_complete(w.layers_host, L)) return CLIPPPO_ERR_NULL;
    clipppo_vit_s* h = new (std::nothrow) clipppo_vit_s();
    if (!h) return CLIPPPO_ERR_WORKSPACE;
    h->cfg = *cfg;
    h->grid = cfg->image / P;
    h->tokens = h->grid * h->grid + 1;
    const int T = h->tokens, kreal = 3 * P * P;
    h->kpatch = (kreal + 63) / 64 * 64;
    h->layers.resize(L);

    // ---- arena layout ----
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_patch = take(static_cast<size_t>(D) * h->kpatch * 2), o_head = take(static_cast<size_t>(O) * D * 2);
    const size_t o_cls = take(D * 4), o_pos = take(static_cast<size_t>(T) * D * 4);
    const size_t o_lnv = take(4 * static_cast<size_t>(D) * 4);
    std::vector<LayerOff> lo(L);
    plan_layers(lo, D, take);
    int st = record_cuda(cudaMalloc(&h->arena, off));
    if (st) { delete h; return st; }
    uint8_t* A = static_cast<uint8_t*>(h->arena);
    auto bf = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(A + o); };
    auto fp = [&](size_t o) { return reinterpret_cast<float*>(A + o); };
    cudaStream_t s0 = nullptr;
    auto blocks = [](size_t n) { return static_cast<unsigned>((n + 255) / 256); };
    auto copyf = [&](float* dst, const float* src, size_t n) { return cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s0); };

    convert_pad_kernel<<<blocks(static_cast<size_t>(D) * h->kpatch), 256, 0, s0>>>(w.conv1, bf(o_patch), D, kreal, h->kpatch);
    transpose_convert_kernel<<<blocks(static_cast<size_t>(D) * O), 256, 0, s0>>>(w.proj, bf(o_head), D, O);
    add_vec_kernel<<<blocks(D), 256, 0, s0>>>(w.class_embedding, w.positional_embedding, fp(o_cls), D);
    copyf(fp(o_pos), w.positional_embedding, static_cast<size_t>(T) * D);
    copyf(fp(o_lnv), w.ln_pre_g, D); copyf(fp(o_lnv) + D, w.ln_pre_b, D);
    copyf(fp(o_lnv) + 2 * D, w.ln_post_g, D); copyf(fp(o_lnv) + 3 * D, w.ln_post_b, D);
    h->w_patch = bf(o_patch); h->w_head = bf(o_head);
    h->cls_pos0 = fp(o_cls); h->pos = fp(o_pos);
    h->ln_pre_g = fp(o_lnv); h->ln_pre_b = fp(o_lnv) + D; h->ln_post_g = fp(o_lnv) + 2 * D; h->ln_post_b = fp(o_lnv) + 3 * D;
    repack_layers(w.layers_host, h->layers, lo, A, D, s0);
    st = record_cuda(cudaGetLastError());
    if (!st) st = record_cuda(cudaStreamSynchronize(s0));      // the caller may free its fp32 weights on return

    const int nb = gemm_b_box_rows();
    if (!st) st = make_bf16_kmajor_tmap(&h->tm_patch, h->w_patch, D, h->kpatch, h->kpatch, nb);
    if (!st) st = make_bf16_kmajor_tmap(&h->tm_head, h->w_head, O, D, D, nb);
    if (!st) st = map_layers(h->layers, D);
    if (st) { cudaFree(h->arena); delete h; return st; }
    *handle = h;
    return CLIPPPO_OK;
}

extern "C" int clipppo_vit_destroy(clipppo_vit_t handle) {
    if (handle) {
        if (handle->arena) cudaFree(handle->arena);
        delete handle;
    }
    return CLIPPPO_OK;
}

// Images per tower pass for a batch of N.  The batch is cut into near-equal chunks (no tiny tail
// chunk whose 90 launches are pure latency), and among the few chunk counts that keep a chunk
// below kMaxChunkImages the one is taken whose GEMMs waste the least of their last wave of 74
// CTA pairs (weighted by the FLOPs of the N = 768 / 2304 / 3072 GEMMs of a block).
static int plan_chunk(const clipppo_vit_s* h, int N) {
    const char* e = getenv("CLIPPPO_VIT_CHUNK");
    if (e && atoi(e) > 0) return N < atoi(e) ? N : atoi(e);
    constexpr int kMaxChunkImages = 4096;       // bigger is faster up to here (profiles/r01_chunk_sweep_v5.txt): 2.5 GB of workspace for ViT-B/32
    if (N <= kMaxChunkImages) return N;
    const int T = h->tokens;
    const int kmin = (N + kMaxChunkImages - 1) / kMaxChunkImages;
    int best = (N + kmin - 1) / kmin;
    double best_score = -1.0;
    for (int k = kmin; k < kmin + 4; ++k) {
        const int c = (N + k - 1) / k;
        double score = 0.0;
        for (int n0 = 0; n0 < N; n0 += c) {
            const int n = (N - n0 < c) ? (N - n0) : c;
            const int pair_rows = ((n * T + 127) / 128 + 1) / 2;
            auto eff = [&](int n_tiles) {
                const double waves = pair_rows * n_tiles / 74.0;
                return waves / static_cast<double>(static_cast<long long>(waves + 0.999999));
            };
            score += n * (0.41 * eff(3) + 0.25 * eff(9) + 0.34 * eff(12));
        }
        score /= N;
        score -= 0.004 * (k - kmin);                 // more chunks = more launches
        if (score > best_score) { best_score = score; best = c; }
    }
    return best;
}

extern "C" int clipppo_vit_workspace_bytes(clipppo_vit_t handle, int n_images, size_t* bytes) {
    if (!handle || !bytes) return CLIPPPO_ERR_NULL;
    if (n_images <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    const int n = plan_chunk(handle, n_images);
    *bytes = carve(handle, n, nullptr).bytes;
    return CLIPPPO_OK;
}

extern "C" int clipppo_vit_encode(clipppo_vit_t handle, const void* images, int img_dtype,
                                  const int64_t img_strides_host[4], int N, int C, int h, int w,
                                  float pre_scale, int flags, float* out,
                                  void* workspace, size_t workspace_bytes, clipppo_stream_t stream) {
    if (!handle || !images || !out || !workspace) return CLIPPPO_ERR_NULL;
    if (N <= 0 || h <= 0 || w <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (C != 1 && C != 3) return CLIPPPO_ERR_BAD_CHANNELS;
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return CLIPPPO_ERR_ALIGN;
    long long s[4] = {static_cast<long long>(C) * h * w, static_cast<long long>(h) * w, w, 1};
    if (img_strides_host) for (int i = 0; i < 4; ++i) s[i] = img_strides_host[i];
    const int chunk = plan_chunk(handle, N);
    const Workspace ws = carve(handle, chunk, workspace);
    if (ws.bytes > workspace_bytes) return CLIPPPO_ERR_WORKSPACE;
    const size_t esz = (img_dtype == CLIPPPO_IMG_U8) ? 1 : 4;
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int n = (N - n0 < chunk) ? (N - n0) : chunk;
        const uint8_t* img = static_cast<const uint8_t*>(images) + static_cast<size_t>(n0) * s[0] * esz;
        int st = encode_chunk(handle, img, img_dtype, s, n, C, h, w, pre_scale, flags,
                              out + static_cast<size_t>(n0) * handle->cfg.out_dim, ws, as_stream(stream));
        if (st) return st;
    }
    return CLIPPPO_OK;
}

// ---- text tower -------------------------------------------------------------------------------
extern "C" int clipppo_text_create(clipppo_text_t* handle, const clipppo_text_config* cfg,
                                   const clipppo_text_weights* weights_host) {
    if (!handle || !cfg || !weights_host || !weights_host->layers_host) return CLIPPPO_ERR_NULL;
    const int D = cfg->width, O = cfg->out_dim, L = cfg->layers, T = cfg->context, V = cfg->vocab;
    if (D <= 0 || L <= 0 || cfg->heads <= 0 || T <= 0 || V <= 0 || O <= 0 || D % cfg->heads) return CLIPPPO_ERR_BAD_SHAPE;
    if (D / cfg->heads != 64 || D % 256 || O % 32) return CLIPPPO_ERR_UNSUPPORTED;
    const clipppo_text_weights& w = *weights_host;
    if (!w.token_embedding || !w.positional_embedding || !w.ln_final_g || !w.ln_final_b || !w.text_projection) return CLIPPPO_ERR_NULL;
    if (!layers_complete(w.layers_host, L)) return CLIPPPO_ERR_NULL;
    clipppo_text_s* h = new (std::nothrow) clipppo_text_s();
    if (!h) return CLIPPPO_ERR_WORKSPACE;
    h->cfg = *cfg;
    h->layers.resize(L);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_emb = take(static_cast<size_t>(V) * D * 4), o_pos = take(static_cast<size_t>(T) * D * 4);
    const size_t o_lnv = take(2 * static_cast<size_t>(D) * 4), o_head = take(static_cast<size_t>(O) * D * 2);
    std::vector<LayerOff> lo(L);
    plan_layers(lo, D, take);
    int st = record_cuda(cudaMalloc(&h->arena, off));
    if (st) { delete h; return st; }
    uint8_t* A = static_cast<uint8_t*>(h->arena);
    auto fp = [&](size_t o) { return reinterpret_cast<float*>(A + o); };
    cudaStream_t s0 = nullptr;
    cudaMemcpyAsync(fp(o_emb), w.token_embedding, static_cast<size_t>(V) * D * 4, cudaMemcpyDeviceToDevice, s0);
    cudaMemcpyAsync(fp(o_pos), w.positional_embedding, static_cast<size_t>(T) * D * 4, cudaMemcpyDeviceToDevice, s0);
    cudaMemcpyAsync(fp(o_lnv), w.ln_final_g, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s0);
    cudaMemcpyAsync(fp(o_lnv) + D, w.ln_final_b, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s0);
    transpose_convert_kernel<<<static_cast<unsigned>((static_cast<size_t>(D) * O + 255) / 256), 256, 0, s0>>>(
        w.text_projection, reinterpret_cast<__nv_bfloat16*>(A + o_head), D, O);
    h->tok_emb = fp(o_emb); h->pos = fp(o_pos); h->ln_final_g = fp(o_lnv); h->ln_final_b = fp(o_lnv) + D;
    h->w_head = reinterpret_cast<const __nv_bfloat16*>(A + o_head);
    repack_layers(w.layers_host, h->layers, lo, A, D, s0);
    st = record_cuda(cudaGetLastError());
    if (!st) st = record_cuda(cudaStreamSynchronize(s0));      // the caller may free its fp32 weights on return
    if (!st) st = make_bf16_kmajor_tmap(&h->tm_head, h->w_head, O, D, D, gemm_b_box_rows());
    if (!st) st = map_layers(h->layers, D);
    if (st) { cudaFree(h->arena); delete h; return st; }
    *handle = h;
    return CLIPPPO_OK;
}

extern "C" int clipppo_text_destroy(clipppo_text_t handle) {
    if (handle) {
        if (handle->arena) cudaFree(handle->arena);
        delete handle;
    }
    return CLIPPPO_OK;
}

static int text_chunk(int N) {
    constexpr int kMaxChunkTexts = 4096;        // 315 k token rows: 2.3 GB of workspace at width 512
    if (N <= kMaxChunkTexts) return N;
    const int k = (N + kMaxChunkTexts - 1) / kMaxChunkTexts;
    return (N + k - 1) / k;
}

extern "C" int clipppo_text_workspace_bytes(clipppo_text_t handle, int n_texts, size_t* bytes) {
    if (!handle || !bytes) return CLIPPPO_ERR_NULL;
    if (n_texts <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    *bytes = carve_text(handle, text_chunk(n_texts), nullptr).bytes;
    return CLIPPPO_OK;
}

extern "C" int clipppo_text_encode(clipppo_text_t handle, const int32_t* tokens, int N, int flags, float* out,
                                   void* workspace, size_t workspace_bytes, clipppo_stream_t stream) {
    if (!handle || !tokens || !out || !workspace) return CLIPPPO_ERR_NULL;
    if (N <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return CLIPPPO_ERR_ALIGN;
    const int chunk = text_chunk(N);
    const TextWorkspace ws = carve_text(handle, chunk, workspace);
    if (ws.bytes > workspace_bytes) return CLIPPPO_ERR_WORKSPACE;
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int n = (N - n0 < chunk) ? (N - n0) : chunk;
        int st = encode_text_chunk(handle, tokens + static_cast<size_t>(n0) * handle->cfg.context, n, flags,
                                   out + static_cast<size_t>(n0) * handle->cfg.out_dim, ws, as_stream(stream));
        if (st) return st;
    }
    return CLIPPPO_OK;
}
