// L1 cosine alignment loss (fwd + bwd), G1 GAE reverse scan, P1 PPO minibatch loss.
//
// All three are tiny, latency-dominated reductions; the point of the kernels is to replace
// ~8 / ~1000 / ~40 eager launches by one or two, with warp-shuffle
// reductions and deterministic summation order.
#include <math.h>

#include "common.cuh"

namespace clipppo {

constexpr float kNormEps = 1e-12f;   // F.normalize default eps

// ---- L1 forward: one warp per row ------------------------------------------------------------
// reference shared/clip_ppo_utils.py:66-74.  row_stats[row] = (|z|, |c|, cos).
__global__ void __launch_bounds__(256)
cosine_rows_kernel(const float* __restrict__ z, const float* __restrict__ c, int rows, int dim,
                   float* __restrict__ row_stats) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float* zr = z + (size_t)warp * dim;
    const float* cr = c + (size_t)warp * dim;
    float zz = 0.f, cc = 0.f, zc = 0.f;
    if ((dim & 3) == 0) {
        const float4* z4 = reinterpret_cast<const float4*>(zr);
        const float4* c4 = reinterpret_cast<const float4*>(cr);
        for (int i = lane; i < (dim >> 2); i += 32) {
            const float4 a = __ldg(z4 + i), b = __ldg(c4 + i);
            zz += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            cc += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
            zc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
    } else {
        for (int i = lane; i < dim; i += 32) {
            const float a = zr[i], b = cr[i];
            zz += a * a; cc += b * b; zc += a * b;
        }
    }
    zz = warp_sum(zz); cc = warp_sum(cc); zc = warp_sum(zc);
    if (lane == 0) {
        const float nz = sqrtf(zz), nc = sqrtf(cc);
        const float cosv = zc / (fmaxf(nz, kNormEps) * fmaxf(nc, kNormEps));
        row_stats[3 * warp + 0] = nz;
        row_stats[3 * warp + 1] = nc;
        row_stats[3 * warp + 2] = cosv;
    }
}

// loss = mean(1 - cos); one CTA, fixed summation order.
__global__ void __launch_bounds__(1024)
cosine_mean_kernel(const float* __restrict__ row_stats, int rows, float* __restrict__ loss) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) s += 1.0f - row_stats[3 * i + 2];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) *loss = s / static_cast<float>(rows);
}

// ---- L1 backward: one warp per row -----------------------------------------------------------
// d/dc_b = -(g/B) * (zhat - cos*chat) / |c|   (and symmetrically for z); below the eps clamp
// F.normalize degenerates to x/eps, whose Jacobian is I/eps.
__global__ void __launch_bounds__(256)
cosine_bwd_kernel(const float* __restrict__ z, const float* __restrict__ c,
                  const float* __restrict__ row_stats, const float* __restrict__ grad_loss,
                  int rows, int dim, float* __restrict__ gz, float* __restrict__ gc) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float nz = row_stats[3 * warp + 0], nc = row_stats[3 * warp + 1], cosv = row_stats[3 * warp + 2];
    const float dz = fmaxf(nz, kNormEps), dc = fmaxf(nc, kNormEps);
    const float coef = -(*grad_loss) / static_cast<float>(rows);
    const bool z_ok = nz >= kNormEps, c_ok = nc >= kNormEps;
    const float* zr = z + (size_t)warp * dim;
    const float* cr = c + (size_t)warp * dim;
    for (int i = lane; i < dim; i += 32) {
        const float zh = zr[i] / dz, ch = cr[i] / dc;
        if (gz) gz[(size_t)warp * dim + i] = coef * (z_ok ? (ch - cosv * zh) : ch) / dz;
        if (gc) gc[(size_t)warp * dim + i] = coef * (c_ok ? (zh - cosv * ch) : zh) / dc;
    }
}

// ---- G1: GAE, one thread per environment, reverse scan over T -------------------------------
// reference clip_ppo_minigrid.py:441-450; every product/sum is rounded like the eager ops.
__global__ void __launch_bounds__(128)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
           const float* __restrict__ dones, const float* __restrict__ next_value,
           const float* __restrict__ next_done, int T, int E, float gamma, float gamma_lambda,
           float* __restrict__ adv, float* __restrict__ ret) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float last = 0.f;
    float nnt = __fsub_rn(1.0f, next_done[e]);
    float nv = next_value[e];
    for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * E + e;
        const float v = values[i], r = rewards[i], d = dones[i];
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, nv), nnt)), v);
        last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, nnt), last));
        adv[i] = last;
        ret[i] = __fadd_rn(last, v);
        nnt = __fsub_rn(1.0f, d);
        nv = v;
    }
}

// ---- P1: PPO minibatch loss, single CTA, deterministic ---------------------------------------
struct PpoParams {
    const float *nlp, *ent, *nv, *olp, *adv, *ret, *ov, *clip_loss;
    int n;
    float clip_coef, ent_coef, vf_coef, clip_lambda;
    int norm_adv, clip_vloss;
    float *stats, *g_nlp, *g_ent, *g_nv;
};

__global__ void __launch_bounds__(1024) ppo_loss_kernel(const PpoParams p) {
    __shared__ float scratch[32];
    const int n = p.n, tid = threadIdx.x, nth = blockDim.x;
    const float inv_n = 1.0f / static_cast<float>(n);
    float mean = 0.f, stdv = 1.f;
    if (p.norm_adv) {
        float s = 0.f;
        for (int i = tid; i < n; i += nth) s += p.adv[i];
        mean = block_sum(s, scratch) * inv_n;
        float q = 0.f;
        for (int i = tid; i < n; i += nth) { const float d = p.adv[i] - mean; q += d * d; }
        q = block_sum(q, scratch);
        stdv = sqrtf(q / static_cast<float>(n - 1));          // unbiased, like Tensor.std()
    }
    const float eps = p.clip_coef;
    float s_pg = 0.f, s_v = 0.f, s_ent = 0.f, s_okl = 0.f, s_kl = 0.f, s_cf = 0.f;
    for (int i = tid; i < n; i += nth) {
        const float lr = p.nlp[i] - p.olp[i];
        const float ratio = expf(lr);
        s_okl += -lr;
        s_kl += (ratio - 1.0f) - lr;
        s_cf += (fabsf(ratio - 1.0f) > eps) ? 1.0f : 0.0f;
        float a = p.adv[i];
        if (p.norm_adv) a = (a - mean) / (stdv + 1e-8f);
        const float rc = fminf(fmaxf(ratio, 1.0f - eps), 1.0f + eps);
        const float pg1 = -a * ratio, pg2 = -a * rc;
        s_pg += fmaxf(pg1, pg2);
        const float inr = (ratio >= 1.0f - eps && ratio <= 1.0f + eps) ? 1.0f : 0.0f;
        float g_ratio;                                         // d max(pg1,pg2) / d ratio
        if (pg1 > pg2) g_ratio = -a;
        else if (pg1 < pg2) g_ratio = -a * inr;
        else g_ratio = 0.5f * (-a) + 0.5f * (-a * inr);        // torch.max splits ties evenly
        const float v = p.nv[i], R = p.ret[i];
        float vterm, g_v;
        const float du = v - R;
        if (p.clip_vloss) {
            const float ov = p.ov[i];
            const float dvo = v - ov;
            const float vc = ov + fminf(fmaxf(dvo, -eps), eps);
            const float dc = vc - R;
            const float vu = du * du, vcl = dc * dc;
            vterm = fmaxf(vu, vcl);
            const float inv = (dvo >= -eps && dvo <= eps) ? 1.0f : 0.0f;
            if (vu > vcl) g_v = 2.0f * du;
            else if (vu < vcl) g_v = 2.0f * dc * inv;
            else g_v = 0.5f * (2.0f * du) + 0.5f * (2.0f * dc * inv);
        } else {
            vterm = du * du;
            g_v = 2.0f * du;
        }
        s_v += vterm;
        s_ent += p.ent[i];
        if (p.g_nlp) p.g_nlp[i] = g_ratio * ratio * inv_n;
        if (p.g_nv) p.g_nv[i] = 0.5f * p.vf_coef * g_v * inv_n;
        if (p.g_ent) p.g_ent[i] = -p.ent_coef * inv_n;
    }
    s_pg = block_sum(s_pg, scratch);
    s_v = block_sum(s_v, scratch);
    s_ent = block_sum(s_ent, scratch);
    s_okl = block_sum(s_okl, scratch);
    s_kl = block_sum(s_kl, scratch);
    s_cf = block_sum(s_cf, scratch);
    if (tid == 0) {
        const float pg = s_pg * inv_n, vl = 0.5f * (s_v * inv_n), en = s_ent * inv_n;
        const float cl = p.clip_loss ? *p.clip_loss : 0.0f;
        p.stats[0] = pg - p.ent_coef * en + vl * p.vf_coef + p.clip_lambda * cl;
        p.stats[1] = pg;
        p.stats[2] = vl;
        p.stats[3] = en;
        p.stats[4] = s_okl * inv_n;
        p.stats[5] = s_kl * inv_n;
        p.stats[6] = s_cf * inv_n;
        p.stats[7] = stdv;
    }
}

}  // namespace clipppo

using namespace clipppo;

extern "C" int clipppo_cosine_loss_fwd(const float* z, const float* c, int rows, int dim,
                                       float* loss, float* row_stats, clipppo_stream_t stream) {
    if (!z || !c || !loss || !row_stats) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || dim <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    const int warps_per_block = 8;
    const int blocks = (rows + warps_per_block - 1) / warps_per_block;
    cosine_rows_kernel<<<blocks, warps_per_block * 32, 0, as_stream(stream)>>>(z, c, rows, dim, row_stats);
    CLIPPPO_CHECK_LAUNCH();
    cosine_mean_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_stats, rows, loss);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

extern "C" int clipppo_cosine_loss_bwd(const float* z, const float* c, const float* row_stats,
                                       const float* grad_loss, int rows, int dim,
                                       float* grad_z, float* grad_c, clipppo_stream_t stream) {
    if (!z || !c || !row_stats || !grad_loss) return CLIPPPO_ERR_NULL;
    if (rows <= 0 || dim <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (!grad_z && !grad_c) return CLIPPPO_OK;
    const int warps_per_block = 8;
    const int blocks = (rows + warps_per_block - 1) / warps_per_block;
    cosine_bwd_kernel<<<blocks, warps_per_block * 32, 0, as_stream(stream)>>>(z, c, row_stats, grad_loss, rows, dim,
                                                                             grad_z, grad_c);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

extern "C" int clipppo_gae_f32(const float* rewards, const float* values, const float* dones,
                               const float* next_value, const float* next_done, int T, int E,
                               double gamma, double gae_lambda, float* advantages, float* returns,
                               clipppo_stream_t stream) {
    if (!rewards || !values || !dones || !next_value || !next_done || !advantages || !returns) return CLIPPPO_ERR_NULL;
    if (T <= 0 || E <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    const int threads = 128;
    gae_kernel<<<(E + threads - 1) / threads, threads, 0, as_stream(stream)>>>(
        rewards, values, dones, next_value, next_done, T, E, static_cast<float>(gamma),
        static_cast<float>(gamma * gae_lambda), advantages, returns);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

extern "C" int clipppo_ppo_loss_f32(const float* newlogprob, const float* entropy, const float* newvalue,
                                    const float* old_logprob, const float* advantages, const float* returns,
                                    const float* old_values, const float* clip_loss, int n,
                                    float clip_coef, float ent_coef, float vf_coef, float clip_lambda,
                                    int norm_adv, int clip_vloss, float* stats_out,
                                    float* g_newlogprob, float* g_entropy, float* g_newvalue,
                                    clipppo_stream_t stream) {
    if (!newlogprob || !entropy || !newvalue || !old_logprob || !advantages || !returns || !stats_out) return CLIPPPO_ERR_NULL;
    if (clip_vloss && !old_values) return CLIPPPO_ERR_NULL;
    if (n <= 0 || (norm_adv && n < 2)) return CLIPPPO_ERR_BAD_SHAPE;
    PpoParams p;
    p.nlp = newlogprob; p.ent = entropy; p.nv = newvalue; p.olp = old_logprob; p.adv = advantages;
    p.ret = returns; p.ov = old_values; p.clip_loss = clip_loss; p.n = n;
    p.clip_coef = clip_coef; p.ent_coef = ent_coef; p.vf_coef = vf_coef; p.clip_lambda = clip_lambda;
    p.norm_adv = norm_adv; p.clip_vloss = clip_vloss;
    p.stats = stats_out; p.g_nlp = g_newlogprob; p.g_ent = g_entropy; p.g_nv = g_newvalue;
    ppo_loss_kernel<<<1, 1024, 0, as_stream(stream)>>>(p);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}
