/*
 * clipppo_b200.h - C ABI of the B200-native CLIP-PPO observation path.
 *
 * One shared library (libclipppo_b200.so, sm_100a only).  Plain pointers and sizes, no
 * torch types.  Every per-call entry point is asynchronous on the caller's stream, never allocates
 * or frees device memory (the caller owns all buffers including workspaces) and returns
 * 0 on success or a negative clipppo_status; only clipppo_vit_create / clipppo_text_create (and
 * their _destroy) allocate - the handle's repacked weights - and synchronise once.  Pointers are DEVICE pointers unless the
 * parameter name ends in _host.
 *
 * Each function cites the reference interface it replaces
 * (paths relative to AlexanderBurkhart/CLIP-PPO).
 */
#ifndef CLIPPPO_B200_H_
#define CLIPPPO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPPPO_ABI_VERSION 4

typedef enum clipppo_status {
    CLIPPPO_OK = 0,
    CLIPPPO_ERR_BAD_SHAPE = -1,      /* non-positive dims, k even, window outside image ...     */
    CLIPPPO_ERR_BAD_CHANNELS = -2,   /* contrast needs C in {1,3} (torchvision raises TypeError) */
    CLIPPPO_ERR_BAD_PAD = -3,        /* reflect pad needs k/2 < min(H,W)                         */
    CLIPPPO_ERR_NULL = -4,           /* required pointer is NULL                                 */
    CLIPPPO_ERR_WORKSPACE = -5,      /* workspace too small                                      */
    CLIPPPO_ERR_UNSUPPORTED = -6,    /* config outside what the kernels were built for           */
    CLIPPPO_ERR_ALIGN = -7,          /* pointer / stride alignment requirement not met           */
    CLIPPPO_ERR_CUDA = -8,           /* a CUDA runtime / driver call failed (see last_cuda_error) */
    CLIPPPO_ERR_DIM_MISMATCH = -9    /* latent vs embedding width (reference raises ValueError)  */
} clipppo_status;

typedef void* clipppo_stream_t;      /* a cudaStream_t */

int         clipppo_abi_version(void);
const char* clipppo_strerror(int status);
/* cudaError_t of the most recent CLIPPPO_ERR_CUDA on this thread (0 if none). */
int         clipppo_last_cuda_error(void);

/* Instrumentation (bench.py): launches of this library's kernels since prof_begin; with
 * time_gemms != 0 every tcgen05 GEMM launch is bracketed by CUDA events on its stream and
 * prof_end returns their summed duration (ms) and algorithmic FLOPs.  Off by default. */
int clipppo_prof_begin(int time_gemms);
int clipppo_prof_end(long long* launches, double* gemm_ms, double* gemm_flops, long long* gemm_launches);
/* Per-shape totals of the last prof_end: tag = epilogue << 40 | N << 20 | K; returns non-zero past the end. */
int clipppo_prof_bucket(int index, long long* tag, double* ms, double* flops, long long* launches);

/* ------------------------------------------------------------------------------------------
 * D1  fused visual disturbance: noise -> contrast -> blur -> cutout in ONE launch.
 * Replaces DisturbanceWrapperGPU.apply_disturbances and its four stage methods
 * (shared/disturbances_gpu.py:66-73, 97-99, 117-119, 137-139, 157-172).
 *
 *   x, noise : fp32 [B,C,H,W] with arbitrary element strides (the MiniGrid call site passes an
 *              NHWC-strided view, minigrid_experiments/clip_ppo/clip_ppo_minigrid.py:385).
 *              noise may be NULL iff CLIPPPO_STAGE_NOISE is not in `stages`.
 *   out      : fp32 [B,C,H,W] contiguous.
 *   stages   : bit mask of CLIPPPO_STAGE_*; stages always run in the reference's fixed order.
 *   contrast : the per-call factor c (one for the batch); the gray mean is per image.
 *   k1d_host : HOST pointer to the k normalised 1-D Gaussian taps (k odd, <= 15).
 *   sh,sw,ph,pw : cutout window (same for every image and channel).
 * ---------------------------------------------------------------------------------------- */
#define CLIPPPO_STAGE_NOISE    1
#define CLIPPPO_STAGE_CONTRAST 2
#define CLIPPPO_STAGE_BLUR     4
#define CLIPPPO_STAGE_CUTOUT   8
#define CLIPPPO_STAGE_ALL      15
#define CLIPPPO_MAX_BLUR_TAPS  15

int clipppo_disturb_f32(const float* x, const int64_t x_strides_host[4],
                        const float* noise, const int64_t noise_strides_host[4],
                        float* out, int B, int C, int H, int W, int stages,
                        float noise_sigma, float contrast,
                        const float* k1d_host, int k,
                        int sh, int sw, int ph, int pw,
                        clipppo_stream_t stream);

/* The same chain on uint8 frames [B,C,H,W] (contiguous): every pixel is read as float(v) * fl(1/255), which is what
 * the reference's `.float() / 255` evaluates to on a CUDA tensor (benchmark_disturbances.py:59, the call sites'
 * `/255`; ATen divides by a host scalar as a multiplication by its reciprocal), so the fp32 copy of the batch is
 * never materialised.  noise is contiguous fp32 [B,C,H,W]; out is contiguous fp32.  Served by the fast
 * kernel only (W % 4 == 0, k <= 7): anything else returns CLIPPPO_ERR_UNSUPPORTED and the caller converts first. */
int clipppo_disturb_u8_f32(const uint8_t* x, const float* noise, float* out, int B, int C, int H, int W, int stages,
                           float noise_sigma, float contrast, const float* k1d_host, int k,
                           int sh, int sw, int ph, int pw, clipppo_stream_t stream);

/* Same chain for the MiniGrid env-step call site (clip_ppo_minigrid.py:381-388) and the *_numpy
 * shims (shared/disturbances_gpu.py:75-95): NHWC frames in (uint8, or fp32 holding 0..255),
 * `/255`, disturb, `*255`, truncate to uint8 NHWC.  noise is fp32, logical [B,C,H,W] with the
 * given element strides (randn_like of an NHWC view keeps NHWC strides). */
int clipppo_disturb_nhwc_u8(const void* obs, int obs_is_f32,
                            const float* noise, const int64_t noise_strides_host[4],
                            uint8_t* out_nhwc, int B, int H, int W, int C, int stages,
                            float noise_sigma, float contrast, const float* k1d_host, int k,
                            int sh, int sw, int ph, int pw, clipppo_stream_t stream);

/* The extended form of clipppo_disturb_f32 / clipppo_disturb_u8_f32 (additive; same chain, same kernels):
 *   out_scale  : the result is multiplied by this on the way out (0 or 1: off) - the `* 255` that follows
 *                apply_disturbances at the call sites (clip_ppo_atari.py:584; the rollout buffers hold 0..255),
 *                one fp32 multiply per element inside the kernel, bit-identical to the separate pass;
 *   CLIPPPO_DISTURB_PHILOX ("noise nullable => in-kernel Philox(seed, offset)"): noise must be NULL; the
 *                N(0,1) draw of element e is Box-Muller over Philox4x32-10(key = philox_seed, counter = (quad index
 *                of e in the logical [first_image + B, C, H, W] tensor, philox_offset)) - a function of the GLOBAL
 *                element index only, so shards of a batch draw the noise of the whole batch.  This is NOT the
 *                stream torch.randn_like draws (the default path keeps that contract by reading the tensor);
 *                oracle/philox.py restates it.  8 B (fp32 frames) / 5 B (uint8) per element instead of 12 / 9.
 * Both are served by the fast kernel only (contiguous [C,H,W] images, W % 4 == 0, k <= 7; PHILOX: k >= 3, NCHW):
 * anything else returns CLIPPPO_ERR_UNSUPPORTED and the caller falls back (torch noise / a separate multiply). */
#define CLIPPPO_DISTURB_PHILOX 1
typedef struct clipppo_disturb_desc {
    const void* x;                          /* frames: fp32 in [0,1] or uint8 0..255 (x_dtype = CLIPPPO_IMG_*) */
    int x_dtype;
    const int64_t* x_strides_host;          /* element strides of the logical [B,C,H,W] view, or NULL = contiguous */
    const float* noise;                     /* fp32, or NULL (no noise stage, or CLIPPPO_DISTURB_PHILOX)       */
    const int64_t* noise_strides_host;
    float* out;                             /* contiguous fp32 [B,C,H,W]                                        */
    int B, C, H, W, stages;
    float noise_sigma, contrast;
    const float* k1d_host;
    int k;
    int sh, sw, ph, pw;
    float out_scale;
    int flags;
    uint64_t philox_seed, philox_offset;
    int64_t first_image;
} clipppo_disturb_desc;
int clipppo_disturb_ex(const clipppo_disturb_desc* desc, clipppo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * L1  cosine alignment loss, forward and backward.
 * Replaces compute_cosine_embedding_loss (shared/clip_ppo_utils.py:48-76) and its autograd.
 *   z, c      : fp32 [rows, dim] contiguous.   loss: fp32 scalar.
 *   row_stats : fp32 [rows,3] scratch written by fwd (|z|, |c|, cos) and read by bwd.
 *   grad_z / grad_c may be NULL to skip that side.  grad_loss is a DEVICE scalar.
 * ---------------------------------------------------------------------------------------- */
int clipppo_cosine_loss_fwd(const float* z, const float* c, int rows, int dim,
                            float* loss, float* row_stats, clipppo_stream_t stream);
int clipppo_cosine_loss_bwd(const float* z, const float* c, const float* row_stats,
                            const float* grad_loss, int rows, int dim,
                            float* grad_z, float* grad_c, clipppo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * G1  generalised advantage estimation, one launch for the whole [T,E] rollout.
 * Replaces the 128-step Python loop at clip_ppo_minigrid.py:437-450 / clip_ppo_atari.py:619-632.
 * gamma / gae_lambda are doubles because the reference rounds (gamma*lambda) to fp32 AFTER the
 * double-precision product; all [T,E] arrays are fp32 contiguous, next_value / next_done are [E].
 * ---------------------------------------------------------------------------------------- */
int clipppo_gae_f32(const float* rewards, const float* values, const float* dones,
                    const float* next_value, const float* next_done, int T, int E,
                    double gamma, double gae_lambda, float* advantages, float* returns,
                    clipppo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * P1  PPO minibatch loss, forward (+ gradients wrt newlogprob / entropy / newvalue).
 * Replaces clip_ppo_minigrid.py:498-531,559 (= clip_ppo_atari.py:691-720,747).
 *   stats_out[8] = { loss, pg_loss, v_loss, entropy, old_approx_kl, approx_kl, clipfrac, adv_std }
 *   clip_loss    : DEVICE scalar (may be NULL => 0), weighted by clip_lambda.
 *   g_* may all be NULL for a forward-only call; otherwise they receive d loss / d input.
 * ---------------------------------------------------------------------------------------- */
int clipppo_ppo_loss_f32(const float* newlogprob, const float* entropy, const float* newvalue,
                         const float* old_logprob, const float* advantages, const float* returns,
                         const float* old_values, const float* clip_loss, int n,
                         float clip_coef, float ent_coef, float vf_coef, float clip_lambda,
                         int norm_adv, int clip_vloss, float* stats_out,
                         float* g_newlogprob, float* g_entropy, float* g_newvalue,
                         clipppo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * V*  frozen CLIP image tower (ViT-B/32, ...): bf16 tcgen05 GEMMs with fp32 accumulation, bf16
 * residual stream, ln_1 / ln_2 folded through the GEMMs that consume them.
 * Replaces clip_model.encode_image / clip.model.VisionTransformer.forward as called from
 * generate_clip_embeddings (shared/clip_ppo_utils.py:141-164) and get_frozen_clip_features
 * (:185-217), including their resize / normalise preprocessing and the final L2 normalise.
 *
 * Weights are handed over as they sit in the checkpoint: fp32 DEVICE pointers in openai/CLIP
 * state-dict layout (`visual.*`).  clipppo_vit_create repacks them once
 * into device memory owned by the handle; the caller may free its copies when create returns.
 * ---------------------------------------------------------------------------------------- */
typedef struct clipppo_vit_config {
    int width, layers, heads, patch, image, out_dim;
} clipppo_vit_config;

typedef struct clipppo_vit_layer {          /* visual.transformer.resblocks.{i}.*            */
    const float *ln1_g, *ln1_b;             /* ln_1.weight / bias            [D]             */
    const float *w_qkv, *b_qkv;             /* attn.in_proj_weight / bias    [3D, D], [3D]   */
    const float *w_out, *b_out;             /* attn.out_proj.weight / bias   [D, D],  [D]    */
    const float *ln2_g, *ln2_b;             /* ln_2.weight / bias            [D]             */
    const float *w_fc, *b_fc;               /* mlp.c_fc.weight / bias        [4D, D], [4D]   */
    const float *w_proj, *b_proj;           /* mlp.c_proj.weight / bias      [D, 4D], [D]    */
} clipppo_vit_layer;

typedef struct clipppo_vit_weights {
    const float* conv1;                     /* visual.conv1.weight           [D, 3, P, P]    */
    const float* class_embedding;           /* visual.class_embedding        [D]             */
    const float* positional_embedding;      /* visual.positional_embedding   [T, D]          */
    const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;   /*           [D]             */
    const float* proj;                      /* visual.proj                   [D, out_dim]    */
    const clipppo_vit_layer* layers_host;   /* HOST array of `layers` entries                */
} clipppo_vit_weights;

typedef struct clipppo_vit_s* clipppo_vit_t;

/* Allocates the handle's weight arena (one cudaMalloc, ~2 bytes per parameter), repacks on the
 * default stream and synchronises it once; the only entry point that allocates or synchronises. */
int clipppo_vit_create(clipppo_vit_t* handle, const clipppo_vit_config* cfg,
                       const clipppo_vit_weights* weights_host);
int clipppo_vit_destroy(clipppo_vit_t handle);
/* Bytes of scratch clipppo_vit_encode needs for n_images (<= the value used at encode). */
int clipppo_vit_workspace_bytes(clipppo_vit_t handle, int n_images, size_t* bytes);

#define CLIPPPO_IMG_F32 0
#define CLIPPPO_IMG_U8  1
#define CLIPPPO_VIT_L2NORM        1   /* L2-normalise the output rows (generate_clip_embeddings)      */
#define CLIPPPO_VIT_PRENORMALIZED 2   /* input already (u - mean)/std: skip the normalise step        */
#define CLIPPPO_VIT_CLS_LAST_BLOCK 4  /* opt-in, exact: after the last block's attention only the class-token rows go through
                                        out_proj / c_fc / c_proj - VisionTransformer.forward reads x[:, 0, :] only, the other
                                        tokens' outputs of that block are never used (embeddings bitwise equal; 6 % fewer FLOPs
                                        on ViT-B/32).  Off by default: the default pass runs every token through every block.  */
/*   images : [N,C,h,w] with element strides; C == 3, or C == 1 (gray broadcast to RGB, the
 *            Atari path clip_ppo_atari.py:249-269).  pre_scale multiplies the raw pixel before
 *            resize (1/255 for generate_clip_embeddings, 1/255^2 for the Atari double divide,
 *            1 for get_frozen_clip_features).
 *   flags  : CLIPPPO_VIT_* bits.
 *   out    : fp32 [N, out_dim].                                                            */
int clipppo_vit_encode(clipppo_vit_t handle, const void* images, int img_dtype,
                       const int64_t img_strides_host[4], int N, int C, int h, int w,
                       float pre_scale, int flags, float* out,
                       void* workspace, size_t workspace_bytes, clipppo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * T*  frozen CLIP text tower (ViT-B/32's: width 512, 12 blocks, 8 heads, context 77, vocab 49408).
 * Replaces clip_model.encode_text as called from generate_clip_embeddings(modality="text")
 * (shared/clip_ppo_utils.py:132-139), the reference's DEFAULT MiniGrid modality: token + positional
 * embedding, the same residual blocks as the image tower under a causal mask, ln_final on the EOT
 * row (argmax of the token ids), text_projection, optional L2 normalise.  Token ids come from the
 * caller (clip.tokenize on the host); weights are fp32 DEVICE pointers in openai/CLIP state-dict
 * layout (token_embedding.weight, positional_embedding, transformer.resblocks.{i}.*, ln_final.*,
 * text_projection) and are repacked once into handle-owned memory, as for the image tower.
 * ---------------------------------------------------------------------------------------- */
typedef struct clipppo_text_config {
    int width, layers, heads, context, vocab, out_dim;
} clipppo_text_config;

typedef struct clipppo_text_weights {
    const float* token_embedding;           /* token_embedding.weight        [vocab, D]      */
    const float* positional_embedding;      /* positional_embedding          [context, D]    */
    const float *ln_final_g, *ln_final_b;   /* ln_final.weight / bias        [D]             */
    const float* text_projection;           /* text_projection               [D, out_dim]    */
    const clipppo_vit_layer* layers_host;   /* HOST array: transformer.resblocks.{i}.*       */
} clipppo_text_weights;

typedef struct clipppo_text_s* clipppo_text_t;

int clipppo_text_create(clipppo_text_t* handle, const clipppo_text_config* cfg,
                        const clipppo_text_weights* weights_host);
int clipppo_text_destroy(clipppo_text_t handle);
int clipppo_text_workspace_bytes(clipppo_text_t handle, int n_texts, size_t* bytes);
/*   tokens : DEVICE int32 [N, context] (clip.tokenize output; ids are clamped to the table).
 *   flags  : CLIPPPO_VIT_L2NORM or 0.      out : fp32 [N, out_dim].                         */
int clipppo_text_encode(clipppo_text_t handle, const int32_t* tokens, int N, int flags, float* out,
                        void* workspace, size_t workspace_bytes, clipppo_stream_t stream);

/* Building blocks of the tower, exported for parity tests and micro-benchmarks. */
int clipppo_preprocess_bf16(const void* images, int img_dtype, const int64_t img_strides_host[4],
                            int N, int C, int h, int w, float pre_scale, int normalize,
                            int patch, int image, void* patches_bf16 /* [N*G*G, 3*P*P] */, clipppo_stream_t stream);
int clipppo_layernorm_bf16(const float* x, const float* gamma, const float* beta, int rows,
                           int width, int64_t row_stride, void* y_bf16, clipppo_stream_t stream);
/* fp32 [rows,2] (mean, 1/sqrt(biased var + 1e-5)) of bf16 rows: what is left of ln_1 / ln_2 once their
 * affine part is folded into the next GEMM (CLIPPPO_EPI_ROWAFFINE_*).  width % 256 == 0. */
int clipppo_rowstats_bf16(const void* x_bf16, int rows, int width, int64_t row_stride, float* stats,
                          clipppo_stream_t stream);
#define CLIPPPO_EPI_BIAS_BF16       0   /* out bf16 = acc + bias                       */
#define CLIPPPO_EPI_BIAS_GELU_BF16  1   /* out bf16 = quickgelu(acc + bias)            */
#define CLIPPPO_EPI_BIAS_RESID_F32  2   /* out fp32 += acc + bias (in place residual)  */
#define CLIPPPO_EPI_PATCH_F32       3   /* token rows of X = acc + pos                 */
#define CLIPPPO_EPI_F32             4   /* out fp32 = acc                              */
#define CLIPPPO_EPI_ROWAFFINE_BF16      6   /* out bf16 = rstd[m]*(acc - mean[m]*colsum[n]) + bias[n]          */
#define CLIPPPO_EPI_ROWAFFINE_GELU_BF16 7   /* ... followed by QuickGELU                                        */
#define CLIPPPO_EPI_RESID_BF16          8   /* out bf16 += bf16(acc + bias)  (in-place bf16 residual stream)    */
#define CLIPPPO_EPI_RESID_STATS_BF16    9   /* out bf16 = bf16(out + acc + bias), + per-row partial statistics   */
/* out[M,N] = epilogue(A[M,K] @ W[N,K]^T); A, W bf16 K-major, 16-byte aligned rows. */
int clipppo_gemm_bf16(const void* a_bf16, const void* w_bf16, int M, int N, int K, int epilogue,
                      const float* bias, const float* pos, int tokens, void* out, int64_t ldo,
                      clipppo_stream_t stream);
/* The three epilogues the tower runs on (6..8).  ROWAFFINE is LayerNorm folded through the GEMM
 * ([clip] ln_1 -> in_proj, ln_2 -> c_fc): A holds the UN-normalised residual rows, W the weight with
 * gamma multiplied into its columns, bias = b + W beta, colsum[n] = sum_k W'[n,k], and row_stats the
 * fp32 [M,2] (mean, 1/sqrt(var+eps)) of the A rows; row_stats = colsum = NULL gives the plain bias
 * epilogue.  N must be a multiple of 64.  Outputs leave through TMA stores / TMA reduce-adds. */
int clipppo_gemm_bf16_fused(const void* a_bf16, const void* w_bf16, int M, int N, int K, int epilogue,
                            const float* bias, const float* row_stats, const float* colsum,
                            void* out_bf16, int64_t ldo, clipppo_stream_t stream);
/* The residual update with the NEXT LayerNorm's row statistics fused in ([clip] ResidualAttentionBlock:
 * x = x + attn(ln_1(x)); x = x + mlp(ln_2(x)) - the statistics ln_2 / the next block's ln_1 need are those of
 * the rows this epilogue has just written): x[M,N] (bf16, in place) = bf16(x + A W^T + bias) with ONE rounding,
 * and row_parts_out fp32 [M, ceil(N/128), 2] = (sum, sum of squares) of the updated row over each 128-column slice.
 * N % 64 == 0. */
int clipppo_gemm_bf16_resid_stats(const void* a_bf16, const void* w_bf16, int M, int N, int K, const float* bias,
                                  void* x_bf16, int64_t ldo, float* row_parts_out, clipppo_stream_t stream);
/* clipppo_gemm_bf16_fused's ROWAFFINE epilogues (6, 7) fed with those partial sums instead of (mean, rstd):
 * row_parts fp32 [M, n_parts, 2]; mean = sum/K, rstd = 1/sqrt(max(sumsq/K - mean^2, 0) + 1e-5), parts added in index order. */
int clipppo_gemm_bf16_fused_parts(const void* a_bf16, const void* w_bf16, int M, int N, int K, int epilogue,
                                  const float* bias, const float* row_parts, int n_parts, const float* colsum,
                                  void* out_bf16, int64_t ldo, clipppo_stream_t stream);
#ifdef CLIPPPO_BUILD_PROBES
/* Measurement probe - compiled only with -DCLIPPPO_BUILD_PROBES (CLIPPPO_BUILD_PROBES=1 python clip-ppo_b200/build.py --force),
 * absent from the product library and its ABI (profiles/ only, never on the product path): the GEMM above with parts switched
 * off - dbg 1: the epilogue drains TMEM but does not compute or store; 2: no TMA operand loads;
 * 3: both; 4: boxes staged but never sent (no output traffic); 8: epilogue arithmetic only.
 * Output is garbage by construction; only the duration means anything. */
int clipppo_gemm_bf16_probe(const void* a_bf16, const void* w_bf16, int M, int N, int K, int epilogue,
                            const float* bias, void* out, int64_t ldo, int dbg, clipppo_stream_t stream);
#endif
int clipppo_attention_bf16(const void* qkv_bf16, int n_images, int tokens, int heads, int head_dim,
                           void* out_bf16, clipppo_stream_t stream);
/* the same under the text tower's causal mask ([clip] build_attention_mask): query t attends to keys 0 .. t */
int clipppo_attention_causal_bf16(const void* qkv_bf16, int n_seqs, int tokens, int heads, int head_dim,
                                  void* out_bf16, clipppo_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * §8f-1  the PPO encoder (NatureCNN) of the reference's Agent, forward and backward, fp32.
 * Replaces `self.network(x)` and its autograd (minigrid_experiments/clip_ppo/clip_ppo_minigrid.py:229-242, :245-262;
 * atari_experiments/clip_ppo/clip_ppo_atari.py:196-209, :229): conv 8x8 s4 (C -> 32), conv 4x4 s2 (32 -> 64),
 * conv 3x3 s1 (64 -> 64), flatten, Linear 3136 -> 512, ReLU after each.
 *   obs        fp32, logical [mb, C, 84, 84] addressed through element strides (NCHW stacks or the NHWC view of
 *              MiniGrid frames); multiplied by in_scale on the way in (1/255 = the scripts' `/ 255.0`)
 *   weights    PyTorch layouts: conv [OC, C, KH, KW], fc [512, 3136] (columns in (c, h, w) order), biases
 *   hidden     [mb, 512] = network(x)
 *   workspace  clipppo_nature_workspace_bytes(mb, C) bytes, 256-byte aligned; holds the three activations for backward
 * No im2col matrix is ever materialised: every convolution operand is gathered inside the GEMM that consumes it.
 * Observations whose strides do not keep every 4-element group of a filter row contiguous and 16-byte aligned return
 * CLIPPPO_ERR_UNSUPPORTED (the caller copies them to NCHW first).
 * backward: gradients of all eight parameters for grad_hidden = dL/d hidden, given the forward's observations and
 * workspace and the current conv2 / conv3 weights; deterministic (no atomics, fixed-order split reductions).
 * ---------------------------------------------------------------------------------------- */
int clipppo_nature_workspace_bytes(int mb, int channels, size_t* bytes);
int clipppo_nature_forward(const float* obs, const int64_t obs_strides_host[4], float in_scale, int mb, int channels,
                           const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                           const float* b3, const float* wfc, const float* bfc, float* hidden, void* workspace,
                           size_t workspace_bytes, clipppo_stream_t stream);
int clipppo_nature_backward(const float* grad_hidden, const float* hidden, const float* obs,
                            const int64_t obs_strides_host[4], float in_scale, int mb, int channels, const float* w2,
                            const float* w3, float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3,
                            float* gwfc, float* gbfc, void* workspace, size_t workspace_bytes, clipppo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPPPO_B200_H_ */
