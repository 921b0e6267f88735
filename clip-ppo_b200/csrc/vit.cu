// Frozen CLIP image tower driver: owns the weight TMA descriptors and sequences the kernels
//   V0 preprocess+im2col -> V1 patch-embed GEMM (+pos) -> ln_pre ->
//   L x [ ln_1 -> QKV GEMM -> attention -> out-proj GEMM (+residual) -> ln_2 -> c_fc GEMM (+QuickGELU)
//         -> c_proj GEMM (+residual) ] -> ln_post(CLS) -> head GEMM -> optional L2 normalise
// Replaces [clip] VisionTransformer.forward / CLIP.encode_image as called by reference
// shared/clip_ppo_utils.py:163-164 and :212-217 (spec: SURVEY.md Appendix B).
//
// Residual stream X is fp32 [n*T, D]; GEMM operands are bf16.  Images are processed in chunks so
// the activation workspace stays bounded (and mostly L2-resident) whatever N is; buffers that are
// never live together (im2col patches / QKV / MLP hidden) share one allocation.
#include <stdlib.h>

#include <new>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"

struct clipppo_vit_s {
    clipppo_vit_config cfg;
    clipppo_vit_weights w;
    std::vector<clipppo_vit_layer> layers;
    int tokens, grid, kpatch;
    CUtensorMap tm_patch, tm_head;
    std::vector<CUtensorMap> tm_qkv, tm_out, tm_fc, tm_proj;
};

namespace clipppo {

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
    float* X;                 // [n*T, D] fp32 residual stream
    __nv_bfloat16* Y;         // [n*T, D] LN output / attention output
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
    ws.X = reinterpret_cast<float*>(b + off);            off += align_up(rows * D * 4, 1024);
    ws.Y = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(rows * D * 2, 1024);
    ws.H = reinterpret_cast<__nv_bfloat16*>(b + off);    off += align_up(h_elems * 2, 1024);
    ws.Ycls = reinterpret_cast<__nv_bfloat16*>(b + off); off += align_up(static_cast<size_t>(n) * D * 2, 1024);
    ws.bytes = off;
    return ws;
}

// X[img*T + 0, :] = class_embedding + positional_embedding[0]
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

int encode_chunk(const clipppo_vit_s* h, const void* images, int img_dtype, const long long strides[4], int n, int C,
                 int ih, int iw, float pre_scale, int flags, float* out, const Workspace& ws, cudaStream_t stream) {
    const int D = h->cfg.width, T = h->tokens, G = h->grid, O = h->cfg.out_dim, L = h->cfg.layers;
    const int rows = n * T, prow = n * G * G;
    CUtensorMap tmY, tmH, tmP, tmC;
    VIT_TRY(make_bf16_kmajor_tmap(&tmP, ws.H, prow, h->kpatch, h->kpatch, gemm_a_box_rows()));
    VIT_TRY(make_bf16_kmajor_tmap(&tmY, ws.Y, rows, D, D, gemm_a_box_rows()));
    VIT_TRY(make_bf16_kmajor_tmap(&tmH, ws.H, rows, 4 * D, 4 * D, gemm_a_box_rows()));
    VIT_TRY(make_bf16_kmajor_tmap(&tmC, ws.Ycls, n, D, D, gemm_a_box_rows()));

    VIT_TRY(preprocess_launch(images, img_dtype, strides, n, C, ih, iw, pre_scale,
                              (flags & CLIPPPO_VIT_PRENORMALIZED) ? 0 : 1, h->cfg.patch, h->cfg.image,
                              h->kpatch, ws.H, stream));
    cls_init_kernel<<<(n * (D / 4) + 255) / 256, 256, 0, stream>>>(ws.X, h->w.cls_pos0, n, T, D);
    CLIPPPO_CHECK_LAUNCH();
    VIT_TRY(gemm_bf16_launch(tmP, h->tm_patch, prow, D, h->kpatch, CLIPPPO_EPI_PATCH_F32, nullptr, h->w.pos, T,
                             ws.X, D, stream));
    VIT_TRY(layernorm_inplace_f32_launch(ws.X, h->w.ln_pre_g, h->w.ln_pre_b, rows, D, D, stream));
    for (int l = 0; l < L; ++l) {
        const clipppo_vit_layer& w = h->layers[l];
        VIT_TRY(layernorm_launch(ws.X, w.ln1_g, w.ln1_b, rows, D, D, ws.Y, stream));
        VIT_TRY(gemm_bf16_launch(tmY, h->tm_qkv[l], rows, 3 * D, D, CLIPPPO_EPI_BIAS_BF16, w.b_qkv, nullptr, 0,
                                 ws.H, 3 * D, stream));
        VIT_TRY(attention_launch(ws.H, n, T, h->cfg.heads, D / h->cfg.heads, ws.Y, stream));
        VIT_TRY(gemm_bf16_launch(tmY, h->tm_out[l], rows, D, D, CLIPPPO_EPI_BIAS_RESID_F32, w.b_out, nullptr, 0,
                                 ws.X, D, stream));
        VIT_TRY(layernorm_launch(ws.X, w.ln2_g, w.ln2_b, rows, D, D, ws.Y, stream));
        VIT_TRY(gemm_bf16_launch(tmY, h->tm_fc[l], rows, 4 * D, D, CLIPPPO_EPI_BIAS_GELU_BF16, w.b_fc, nullptr, 0,
                                 ws.H, 4 * D, stream));
        VIT_TRY(gemm_bf16_launch(tmH, h->tm_proj[l], rows, D, 4 * D, CLIPPPO_EPI_BIAS_RESID_F32, w.b_proj, nullptr, 0,
                                 ws.X, D, stream));
    }
    VIT_TRY(layernorm_launch(ws.X, h->w.ln_post_g, h->w.ln_post_b, n, D, static_cast<long long>(T) * D, ws.Ycls, stream));
    VIT_TRY(gemm_bf16_launch(tmC, h->tm_head, n, O, D, CLIPPPO_EPI_F32, nullptr, nullptr, 0, out, O, stream));
    if (flags & CLIPPPO_VIT_L2NORM) {
        l2norm_rows_kernel<<<(n + 7) / 8, 256, 0, stream>>>(out, n, O);
        CLIPPPO_CHECK_LAUNCH();
    }
    return CLIPPPO_OK;
}

}  // namespace
}  // namespace clipppo

using namespace clipppo;

extern "C" int clipppo_vit_create(clipppo_vit_t* handle, const clipppo_vit_config* cfg,
                                  const clipppo_vit_weights* weights_host) {
    if (!handle || !cfg || !weights_host || !weights_host->layers_host) return CLIPPPO_ERR_NULL;
    const int D = cfg->width, P = cfg->patch;
    if (D <= 0 || cfg->layers <= 0 || cfg->heads <= 0 || P <= 0 || cfg->image <= 0 || cfg->out_dim <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (cfg->image % P || D % cfg->heads) return CLIPPPO_ERR_BAD_SHAPE;
    if (D / cfg->heads != 64 || D % 128 || D % 64 || cfg->out_dim % 32) return CLIPPPO_ERR_UNSUPPORTED;
    const clipppo_vit_weights& w = *weights_host;
    if (!w.w_patch || !w.cls_pos0 || !w.pos || !w.ln_pre_g || !w.ln_pre_b || !w.ln_post_g || !w.ln_post_b || !w.w_head)
        return CLIPPPO_ERR_NULL;
    clipppo_vit_s* h = new (std::nothrow) clipppo_vit_s();
    if (!h) return CLIPPPO_ERR_WORKSPACE;
    h->cfg = *cfg;
    h->w = w;
    h->layers.assign(w.layers_host, w.layers_host + cfg->layers);
    h->w.layers_host = nullptr;
    h->grid = cfg->image / P;
    h->tokens = h->grid * h->grid + 1;
    h->kpatch = (3 * P * P + 63) / 64 * 64;
    const int nb = gemm_b_box_rows();
    int st = make_bf16_kmajor_tmap(&h->tm_patch, w.w_patch, D, h->kpatch, h->kpatch, nb);
    if (!st) st = make_bf16_kmajor_tmap(&h->tm_head, w.w_head, cfg->out_dim, D, D, nb);
    h->tm_qkv.resize(cfg->layers); h->tm_out.resize(cfg->layers); h->tm_fc.resize(cfg->layers); h->tm_proj.resize(cfg->layers);
    for (int l = 0; l < cfg->layers && !st; ++l) {
        const clipppo_vit_layer& lw = h->layers[l];
        if (!lw.w_qkv || !lw.b_qkv || !lw.w_out || !lw.b_out || !lw.w_fc || !lw.b_fc || !lw.w_proj || !lw.b_proj ||
            !lw.ln1_g || !lw.ln1_b || !lw.ln2_g || !lw.ln2_b) { st = CLIPPPO_ERR_NULL; break; }
        st = make_bf16_kmajor_tmap(&h->tm_qkv[l], lw.w_qkv, 3 * D, D, D, nb);
        if (!st) st = make_bf16_kmajor_tmap(&h->tm_out[l], lw.w_out, D, D, D, nb);
        if (!st) st = make_bf16_kmajor_tmap(&h->tm_fc[l], lw.w_fc, 4 * D, D, D, nb);
        if (!st) st = make_bf16_kmajor_tmap(&h->tm_proj[l], lw.w_proj, D, 4 * D, 4 * D, nb);
    }
    if (st) { delete h; return st; }
    *handle = h;
    return CLIPPPO_OK;
}

extern "C" int clipppo_vit_destroy(clipppo_vit_t handle) {
    delete handle;
    return CLIPPPO_OK;
}

// Images per tower pass for a batch of N.  The batch is cut into near-equal chunks (no tiny tail
// chunk whose 90 launches are pure latency), and among the few chunk counts that keep a chunk
// below kMaxChunkImages the one is taken whose GEMMs waste the least of their last wave of 74
// CTA pairs (weighted by the FLOPs of the N = 768 / 2304 / 3072 GEMMs of a block).
static int plan_chunk(const clipppo_vit_s* h, int N) {
    const char* e = getenv("CLIPPPO_VIT_CHUNK");
    if (e && atoi(e) > 0) return N < atoi(e) ? N : atoi(e);
    constexpr int kMaxChunkImages = 1400;
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
