// V4 - multi-head self-attention core of one ResidualAttentionBlock: softmax(Q K^T / sqrt(d)) V,
// no mask, eval mode.  Replaces the SDPA call inside [clip] nn.MultiheadAttention.
//
// ViT-B/32 has T = 50 tokens and d_head = 64: ~1 % of the tower's FLOPs, HBM/L2-bound (it reads
// the packed QKV rows once and writes the head outputs once), so one CTA owns one (image, head)
// and keeps Q, K, V (padded to 64 rows) in shared memory.  The two tiny matmuls run on the
// warp-level tensor-core path (mma.sync m16n8k16 bf16, fp32 accumulate); the softmax lives in
// the accumulator registers and P never leaves them.  Keys >= T are masked; padded V rows are 0.
#include "common.cuh"
#include "gemm.cuh"

namespace clipppo {

namespace {

constexpr int TP = 64;          // padded tokens
constexpr int DH = 64;          // head dim
constexpr int PITCH = DH + 8;   // bf16 elements per smem row (144 B: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// grid = n_images * heads, 128 threads (4 warps x 16 query rows)
__global__ void __launch_bounds__(128)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, int T, int heads, __nv_bfloat16* __restrict__ out) {
    __shared__ __align__(16) __nv_bfloat16 sQ[TP * PITCH];
    __shared__ __align__(16) __nv_bfloat16 sK[TP * PITCH];
    __shared__ __align__(16) __nv_bfloat16 sV[TP * PITCH];
    const int img = blockIdx.x / heads, head = blockIdx.x - img * heads;
    const int D = heads * DH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- load Q, K, V head slices (rows >= T zero-filled) ----
    const __nv_bfloat16* src = qkv + static_cast<size_t>(img) * T * (3 * D) + head * DH;
    for (int i = tid; i < 3 * TP * 8; i += 128) {
        const int m = i / (TP * 8);                  // 0 Q, 1 K, 2 V
        const int rem = i - m * (TP * 8);
        const int r = rem >> 3, c8 = rem & 7;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (r < T) val = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * (3 * D) + m * D + c8 * 8);
        __nv_bfloat16* dst = (m == 0 ? sQ : (m == 1 ? sK : sV)) + r * PITCH + c8 * 8;
        *reinterpret_cast<uint4*>(dst) = val;
    }
    __syncthreads();

    const int row0 = warp * 16;
    if (row0 < T) {
        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
        const uint32_t q_addr = ptx_smem(sQ) + ((row0 + (lane & 15)) * PITCH + (lane >> 4) * 8) * 2;
        const uint32_t k_addr = ptx_smem(sK) + (((lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 8) * 2;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
            uint32_t a[4];
            ldsm_x4(q_addr + ks * 32, a[0], a[1], a[2], a[3]);
#pragma unroll
            for (int np = 0; np < 4; ++np) {          // pairs of key tiles
                uint32_t b0, b1, b2, b3;
                ldsm_x4(k_addr + (np * 16 * PITCH) * 2 + ks * 32, b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], a, b0, b1);
                mma_bf16_16816(s[2 * np + 1], a, b2, b3);
            }
        }
        // ---- softmax over keys (rows lane/4 and lane/4 + 8), scale 1/sqrt(64) folded into exp2 ----
        const float sl2 = 0.125f * 1.4426950408889634f;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int n = nt * 8 + 2 * (lane & 3);
            if (n >= T)     { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (n + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f((s[nt][0] - mx0) * sl2);
            s[nt][1] = exp2f((s[nt][1] - mx0) * sl2);
            s[nt][2] = exp2f((s[nt][2] - mx1) * sl2);
            s[nt][3] = exp2f((s[nt][3] - mx1) * sl2);
            sum0 += s[nt][0] + s[nt][1];
            sum1 += s[nt][2] + s[nt][3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

        // ---- O = P V ----
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
        const uint32_t v_addr = ptx_smem(sV) + (((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2;
#pragma unroll
        for (int kk = 0; kk < TP / 16; ++kk) {        // 16 keys per step
            uint32_t a[4];
            a[0] = pack2(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
            a[1] = pack2(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
            a[2] = pack2(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
            a[3] = pack2(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {          // pairs of d tiles
                uint32_t b0, b1, b2, b3;
                ldsm_x4_trans(v_addr + (kk * 16 * PITCH + dp * 16) * 2, b0, b1, b2, b3);
                mma_bf16_16816(o[2 * dp], a, b0, b1);
                mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
            }
        }
        // ---- stage the 16 x 64 output in this warp's own Q rows, then 16-byte coalesced stores ----
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * (lane & 3);
            *reinterpret_cast<uint32_t*>(sQ + (row0 + (lane >> 2)) * PITCH + c) = pack2(o[nt][0], o[nt][1]);
            *reinterpret_cast<uint32_t*>(sQ + (row0 + (lane >> 2) + 8) * PITCH + c) = pack2(o[nt][2], o[nt][3]);
        }
        __syncwarp();
        __nv_bfloat16* dst = out + static_cast<size_t>(img) * T * D + head * DH;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int r = row0 + it * 4 + (lane >> 3), c8 = lane & 7;
            if (r < T)
                *reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * D + c8 * 8) =
                    *reinterpret_cast<const uint4*>(sQ + r * PITCH + c8 * 8);
        }
    }
}

}  // namespace

int attention_launch(const void* qkv_bf16, int n_images, int tokens, int heads, int head_dim, void* out_bf16,
                     cudaStream_t stream) {
    if (!qkv_bf16 || !out_bf16) return CLIPPPO_ERR_NULL;
    if (n_images <= 0 || tokens <= 0 || heads <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (head_dim != DH || tokens > TP) return CLIPPPO_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(qkv_bf16) % 16) || (reinterpret_cast<uintptr_t>(out_bf16) % 16)) return CLIPPPO_ERR_ALIGN;
    attention_kernel<<<static_cast<unsigned>(n_images) * heads, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv_bf16), tokens, heads, static_cast<__nv_bfloat16*>(out_bf16));
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

}  // namespace clipppo

extern "C" int clipppo_attention_bf16(const void* qkv_bf16, int n_images, int tokens, int heads, int head_dim,
                                      void* out_bf16, clipppo_stream_t stream) {
    return clipppo::attention_launch(qkv_bf16, n_images, tokens, heads, head_dim, out_bf16, clipppo::as_stream(stream));
}
