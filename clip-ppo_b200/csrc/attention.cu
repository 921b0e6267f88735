// V4 - multi-head self-attention core of one ResidualAttentionBlock: softmax(Q K^T / sqrt(d)) V,
// no mask, eval mode.  Replaces the SDPA call inside [clip] nn.MultiheadAttention.
//
// Since round 2 every shape the towers produce runs on the tcgen05 / TMEM / TMA kernels of attention_tc.cu
// (T <= 64: two images per tile; causal T <= 128: one sequence per tile; 64 < T <= 257 without a mask); the
// mma.sync kernels of this file remain as the fallback for other shapes (T > 257, T < 16, unmasked 64 < T with
// CLIPPPO_ATT_TC=0) and as the A/B reference (CLIPPPO_ATT_TC50=0).  attention_launch() below is the dispatcher.
//
// ViT-B/32 has T = 50 tokens and d_head = 64: ~1 % of the tower's FLOPs, bound by memory latency
// and instruction issue rather than math.  Persistent CTAs (4 per SM) walk over (image, head)
// items; the Q/K/V head slices of item i+1 stream into the second shared-memory buffer with
// cp.async (zero-filled beyond T) while item i is computed, so no warp ever waits on a global
// load.  The two tiny matmuls run on the warp-level tensor-core path (mma.sync m16n8k16 bf16, fp32
// accumulate); the softmax lives in the accumulator registers and P never leaves them.
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

namespace clipppo {

namespace {

constexpr int TP = 64;          // padded tokens
constexpr int DH = 64;          // head dim
constexpr int PITCH = DH + 8;   // bf16 elements per smem row (144 B: conflict-free ldmatrix)
constexpr int MAT_ELEMS = TP * PITCH;
constexpr int BUF_ELEMS = 3 * MAT_ELEMS;                    // Q, K, V of one head
constexpr int ATT_SMEM_BYTES = 2 * BUF_ELEMS * 2;           // double buffered: 55296 B
constexpr int ATT_THREADS = 128;
constexpr int ATT_CTAS_PER_SM = 4;

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
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination instead
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Q, K, V slices of one (image, head): 3 x 64 rows x 8 chunks of 16 B = 12 chunks per thread.  With the
// head count a template constant every one of the 12 addresses is the thread's item pointer plus an
// immediate, and the row-validity predicates are computed once per thread, not per item (address
// arithmetic was 36 % of the kernel's instructions before).
template <int HEADS>
__device__ __forceinline__ void prefetch_item(const __nv_bfloat16* __restrict__ item_ptr, const bool (&row_ok)[4],
                                              const __nv_bfloat16* __restrict__ any_valid, uint32_t dst0) {
    constexpr int D = HEADS * DH;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
            const __nv_bfloat16* g = row_ok[pass] ? item_ptr + (pass * 16 * 3 * D + m * D) : any_valid;
            cp_async16(dst0 + (m * MAT_ELEMS + pass * 16 * PITCH) * 2, g, row_ok[pass] ? 16 : 0);
        }
    }
}

template <int HEADS>
__global__ void __launch_bounds__(ATT_THREADS, ATT_CTAS_PER_SM)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, int n_items, int T, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) __nv_bfloat16 att_smem[];
    constexpr int D = HEADS * DH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = ptx_smem(att_smem);
    const int row0 = warp * 16;
    // per-thread constants of the prefetch: my 16-byte column piece, my first row, which of my 4 rows exist
    const int c8 = tid & 7, pr0 = tid >> 3;
    const bool row_ok[4] = {pr0 < T, pr0 + 16 < T, pr0 + 32 < T, pr0 + 48 < T};
    const size_t thread_off = static_cast<size_t>(pr0) * (3 * D) + c8 * 8;
    const uint32_t dst_off = (pr0 * PITCH + c8 * 8) * 2;
    const size_t img_stride = static_cast<size_t>(T) * (3 * D);
    auto item_ptr = [&](int item) {
        const int img = item / HEADS, head = item - img * HEADS;      // compile-time divisor
        return qkv + img * img_stride + head * DH + thread_off;
    };

    pdl_launch_dependents();
    pdl_wait();
    int item = blockIdx.x;
    if (item < n_items) prefetch_item<HEADS>(item_ptr(item), row_ok, qkv, smem0 + dst_off);
    cp_async_commit();
    int b = 0;
    for (; item < n_items; item += gridDim.x, b ^= 1) {
        const int next = item + gridDim.x;
        if (next < n_items) prefetch_item<HEADS>(item_ptr(next), row_ok, qkv, smem0 + (b ^ 1) * BUF_ELEMS * 2 + dst_off);
        cp_async_commit();
        cp_async_wait<1>();                                   // this item's slices have landed
        __syncthreads();

        __nv_bfloat16* sQ = att_smem + b * BUF_ELEMS;
        const uint32_t q_base = smem0 + b * BUF_ELEMS * 2;
        const uint32_t k_base = q_base + MAT_ELEMS * 2, v_base = k_base + MAT_ELEMS * 2;
        const int img = item / HEADS, head = item - img * HEADS;
        if (row0 < T) {
            // ---- S = Q K^T (16 x 64 per warp) ----
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
            const uint32_t q_addr = q_base + ((row0 + (lane & 15)) * PITCH + (lane >> 4) * 8) * 2;
            const uint32_t k_addr = k_base + (((lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 8) * 2;
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
                if (nt * 8 + 8 > T) {                     // only the tiles that straddle T need masking
                    const int n = nt * 8 + 2 * (lane & 3);
                    if (n >= T)     { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                    if (n + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
                }
                mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float m0 = mx0 * sl2, m1 = mx1 * sl2;
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = ex2(fmaf(s[nt][0], sl2, -m0));
                s[nt][1] = ex2(fmaf(s[nt][1], sl2, -m0));
                s[nt][2] = ex2(fmaf(s[nt][2], sl2, -m1));
                s[nt][3] = ex2(fmaf(s[nt][3], sl2, -m1));
                sum0 += s[nt][0] + s[nt][1];
                sum1 += s[nt][2] + s[nt][3];
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            const float inv0 = __frcp_rn(sum0), inv1 = __frcp_rn(sum1);

            // ---- O = P V ----
            float o[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
            const uint32_t v_addr = v_base + (((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2;
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
        __syncthreads();                                      // buffer b is overwritten by the prefetch of item+2
    }
    cp_async_wait<0>();
}

// ---- any T (ViT-L/14: T = 257) ---------------------------------------------------------------
// One CTA = one (image, head): K and V of all T tokens are loaded into shared memory ONCE (cp.async,
// zero-filled up to a multiple of 16 rows), then the 6 warps walk the ceil(T/16) query tiles of 16 rows;
// a warp stages its query tile through a private 16-row buffer, runs the keys in blocks of 64 with an
// online softmax (running maximum / sum in registers, output accumulator rescaled when the maximum
// moves) and writes its 16 x 64 output tile back through the same buffer.  A last key block of <= 8 keys
// (T = 257: exactly the one key left after four blocks) costs one 8-key MMA tile instead of a padded block.
// Long sequences (T = 257): 6 warps, 2 CTAs / SM.  Short ones (the text tower, T = 77: five query tiles, 20 KB of
// K / V): 3 warps and 4 CTAs / SM, so one CTA's load phase hides under its neighbours' arithmetic.
constexpr int ATTG_QROWS = 16;
constexpr int ATTG_SHORT_T = 128;

__host__ __device__ constexpr int attg_kv_rows(int T) { return (T + 15) / 16 * 16; }
__host__ __device__ constexpr size_t attg_smem_bytes(int T, int warps) {
    return (static_cast<size_t>(2) * attg_kv_rows(T) + warps * ATTG_QROWS) * PITCH * 2;
}

// CAUSAL (the text tower, [clip] build_attention_mask): query t sees keys 0 .. t.  Key blocks that start
// beyond a query tile's last row are skipped; the block on the diagonal is masked element-wise (its first
// key is <= every query row of the tile, so no row of a block is ever fully masked).
template <int HEADS, bool CAUSAL, int ATTG_WARPS>
__global__ void __launch_bounds__(ATTG_WARPS * 32, ATTG_WARPS == 6 ? 2 : 4)
attention_general_kernel(const __nv_bfloat16* __restrict__ qkv, int T, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) __nv_bfloat16 att_smem[];
    constexpr int D = HEADS * DH;
    constexpr int ATTG_THREADS = ATTG_WARPS * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_kv = attg_kv_rows(T);
    const uint32_t k_base = ptx_smem(att_smem);
    const uint32_t v_base = k_base + rows_kv * PITCH * 2;
    __nv_bfloat16* sQ = att_smem + 2 * rows_kv * PITCH + warp * ATTG_QROWS * PITCH;
    const uint32_t q_base = ptx_smem(sQ);
    const int item = blockIdx.x;
    const int img = item / HEADS, head = item - img * HEADS;
    const __nv_bfloat16* base = qkv + static_cast<size_t>(img) * T * (3 * D) + head * DH;
    __nv_bfloat16* obase = out + static_cast<size_t>(img) * T * D + head * DH;

    pdl_launch_dependents();
    pdl_wait();
    // K, V -> smem: 8 x 16-byte pieces per row
    for (int i = tid; i < rows_kv * 8; i += ATTG_THREADS) {
        const int r = i >> 3, c8 = i & 7;
        const bool ok = r < T;
        const __nv_bfloat16* g = ok ? base + static_cast<size_t>(r) * (3 * D) + c8 * 8 : qkv;
        cp_async16(k_base + (r * PITCH + c8 * 8) * 2, g + D, ok ? 16 : 0);
        cp_async16(v_base + (r * PITCH + c8 * 8) * 2, g + 2 * D, ok ? 16 : 0);
    }
    cp_async_commit();

    const float sl2 = 0.125f * 1.4426950408889634f;
    const int n_qt = (T + ATTG_QROWS - 1) / ATTG_QROWS;
    const int n_full = T / TP, k_tail = T - n_full * TP;          // key blocks of 64, keys left over
    bool kv_ready = false;
    for (int qt = warp; qt < n_qt; qt += ATTG_WARPS) {
        // my query tile -> my staging rows (2 pieces per lane), zero-filled beyond T
        {
            const int r = lane >> 1, c0 = (lane & 1) * 4;
            const bool ok = qt * ATTG_QROWS + r < T;
            const __nv_bfloat16* g = ok ? base + static_cast<size_t>(qt * ATTG_QROWS + r) * (3 * D) + c0 * 8 : qkv;
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async16(q_base + (r * PITCH + (c0 + c) * 8) * 2, g + c * 8, ok ? 16 : 0);
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
        }
        if (!kv_ready) { __syncthreads(); kv_ready = true; }     // everybody's K / V pieces have landed (first trip only)
        uint32_t aq[DH / 16][4];
        {
            const uint32_t q_addr = q_base + ((lane & 15) * PITCH + (lane >> 4) * 8) * 2;
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) ldsm_x4(q_addr + ks * 32, aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
        }
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
        float mrun0 = -INFINITY, mrun1 = -INFINITY, lrun0 = 0.f, lrun1 = 0.f;      // rows lane/4 and lane/4 + 8
        int nkb = n_full + (k_tail > 8 ? 1 : 0);                  // blocks run at full width (a wide tail is masked)
        const int q_last = qt * ATTG_QROWS + ATTG_QROWS - 1;      // last query row of my tile
        if (CAUSAL && nkb > q_last / TP + 1) nkb = q_last / TP + 1;
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t kb_k = k_base + kb * TP * PITCH * 2, kb_v = v_base + kb * TP * PITCH * 2;
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
            const uint32_t k_addr = kb_k + (((lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 8) * 2;
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(k_addr + (np * 16 * PITCH) * 2 + ks * 32, b0, b1, b2, b3);
                    mma_bf16_16816(s[2 * np], aq[ks], b0, b1);
                    mma_bf16_16816(s[2 * np + 1], aq[ks], b2, b3);
                }
            }
            const int kleft = T - kb * TP;                        // valid keys in this block (>= 64 unless it is the tail)
            // causal: keys of this block visible to rows lane/4 and lane/4 + 8 of my tile (block-relative count)
            const int vis0 = CAUSAL ? qt * ATTG_QROWS + (lane >> 2) - kb * TP + 1 : TP;
            const bool diag = CAUSAL && kb * TP + TP - 1 > qt * ATTG_QROWS;
            float mx0 = mrun0, mx1 = mrun1;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (nt * 8 + 8 > kleft) {
                    const int n = nt * 8 + 2 * (lane & 3);
                    if (n >= kleft)     { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                    if (n + 1 >= kleft) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
                }
                if (diag) {
                    const int n = nt * 8 + 2 * (lane & 3);
                    if (n >= vis0)         s[nt][0] = -INFINITY;
                    if (n + 1 >= vis0)     s[nt][1] = -INFINITY;
                    if (n >= vis0 + 8)     s[nt][2] = -INFINITY;
                    if (n + 1 >= vis0 + 8) s[nt][3] = -INFINITY;
                }
                mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float c0 = ex2((mrun0 - mx0) * sl2), c1 = ex2((mrun1 - mx1) * sl2);      // exp2(-inf) = 0 on the first block
            mrun0 = mx0; mrun1 = mx1;
            const float m0 = mx0 * sl2, m1 = mx1 * sl2;
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = ex2(fmaf(s[nt][0], sl2, -m0));
                s[nt][1] = ex2(fmaf(s[nt][1], sl2, -m0));
                s[nt][2] = ex2(fmaf(s[nt][2], sl2, -m1));
                s[nt][3] = ex2(fmaf(s[nt][3], sl2, -m1));
                sum0 += s[nt][0] + s[nt][1];
                sum1 += s[nt][2] + s[nt][3];
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            lrun0 = lrun0 * c0 + sum0;
            lrun1 = lrun1 * c1 + sum1;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
            const uint32_t v_addr = kb_v + (((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2;
            const int ksteps = kleft >= TP ? TP / 16 : (kleft + 15) / 16;      // zero-filled rows beyond T contribute nothing
            for (int kk = 0; kk < TP / 16; ++kk) {
                if (kk < ksteps) {
                    uint32_t a[4];
                    a[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
                    a[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
                    a[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                    a[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4_trans(v_addr + (kk * 16 * PITCH + dp * 16) * 2, b0, b1, b2, b3);
                        mma_bf16_16816(o[2 * dp], a, b0, b1);
                        mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
                    }
                }
            }
        }
        if (k_tail > 0 && k_tail <= 8 && (!CAUSAL || n_full * TP <= qt * ATTG_QROWS)) {
            // narrow tail: one 8-key tile (keys n_full*64 .. +7, those beyond T masked), one 16-key PV step
            const uint32_t kb_k = k_base + n_full * TP * PITCH * 2, kb_v = v_base + n_full * TP * PITCH * 2;
            float s0[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t k_addr = kb_k + ((lane & 7) * PITCH + ((lane >> 3) & 1) * 8) * 2;     // lanes 16-31 repeat 0-15 (their matrices are unused)
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(k_addr + ks * 32, b0, b1, b2, b3);
                mma_bf16_16816(s0, aq[ks], b0, b1);
            }
            const int n = 2 * (lane & 3);
            if (n >= k_tail)     { s0[0] = -INFINITY; s0[2] = -INFINITY; }
            if (n + 1 >= k_tail) { s0[1] = -INFINITY; s0[3] = -INFINITY; }
            if (CAUSAL) {
                const int vis0 = qt * ATTG_QROWS + (lane >> 2) - n_full * TP + 1;
                if (n >= vis0)         s0[0] = -INFINITY;
                if (n + 1 >= vis0)     s0[1] = -INFINITY;
                if (n >= vis0 + 8)     s0[2] = -INFINITY;
                if (n + 1 >= vis0 + 8) s0[3] = -INFINITY;
            }
            float mx0 = fmaxf(mrun0, fmaxf(s0[0], s0[1])), mx1 = fmaxf(mrun1, fmaxf(s0[2], s0[3]));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float c0 = ex2((mrun0 - mx0) * sl2), c1 = ex2((mrun1 - mx1) * sl2);
            const float m0 = mx0 * sl2, m1 = mx1 * sl2;
            s0[0] = ex2(fmaf(s0[0], sl2, -m0)); s0[1] = ex2(fmaf(s0[1], sl2, -m0));
            s0[2] = ex2(fmaf(s0[2], sl2, -m1)); s0[3] = ex2(fmaf(s0[3], sl2, -m1));
            float sum0 = s0[0] + s0[1], sum1 = s0[2] + s0[3];
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            lrun0 = lrun0 * c0 + sum0;
            lrun1 = lrun1 * c1 + sum1;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
            uint32_t a[4] = {pack2(s0[0], s0[1]), pack2(s0[2], s0[3]), 0u, 0u};            // keys 8..15 of the step: zero weight
            const uint32_t v_addr = kb_v + (((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2;
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_trans(v_addr + (dp * 16) * 2, b0, b1, b2, b3);
                mma_bf16_16816(o[2 * dp], a, b0, b1);
                mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
            }
        }
        // ---- normalise, stage the 16 x 64 tile in my query rows, 16-byte coalesced stores ----
        const float inv0 = __frcp_rn(lrun0), inv1 = __frcp_rn(lrun1);
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * (lane & 3);
            *reinterpret_cast<uint32_t*>(sQ + (lane >> 2) * PITCH + c) = pack2(o[nt][0] * inv0, o[nt][1] * inv0);
            *reinterpret_cast<uint32_t*>(sQ + ((lane >> 2) + 8) * PITCH + c) = pack2(o[nt][2] * inv1, o[nt][3] * inv1);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int r = it * 4 + (lane >> 3), c8 = lane & 7;
            const int t = qt * ATTG_QROWS + r;
            if (t < T)
                *reinterpret_cast<uint4*>(obase + static_cast<size_t>(t) * D + c8 * 8) =
                    *reinterpret_cast<const uint4*>(sQ + r * PITCH + c8 * 8);
        }
        __syncwarp();                                             // my staging rows are refilled by the next query tile
    }
    if (!kv_ready) __syncthreads();                               // a warp without a query tile still meets the barrier
    cp_async_wait<0>();
}

}  // namespace

template <int HEADS, bool CAUSAL, int WARPS>
static int launch_general_w(const __nv_bfloat16* q, int n_images, int tokens, __nv_bfloat16* o, cudaStream_t stream) {
    const size_t smem = attg_smem_bytes(tokens, WARPS);
    if (smem > 113 * 1024) return CLIPPPO_ERR_UNSUPPORTED;          // 2 CTAs / SM; T <= 352
    static DeviceOnce configured;
    if (configured.first_use()) {
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(attention_general_kernel<HEADS, CAUSAL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    }
    CLIPPPO_CUDA_TRY(launch_pdl(attention_general_kernel<HEADS, CAUSAL, WARPS>, static_cast<unsigned>(n_images) * HEADS, WARPS * 32, smem,
                                stream, 1, q, tokens, o));
    return CLIPPPO_OK;
}

template <int HEADS, bool CAUSAL>
static int launch_general(const __nv_bfloat16* q, int n_images, int tokens, __nv_bfloat16* o, cudaStream_t stream) {
    static const int short_warps = [] { const char* e = getenv("CLIPPPO_ATT_SHORT_WARPS"); return e ? atoi(e) : 3; }();
    if (tokens <= ATTG_SHORT_T && short_warps == 3) return launch_general_w<HEADS, CAUSAL, 3>(q, n_images, tokens, o, stream);
    return launch_general_w<HEADS, CAUSAL, 6>(q, n_images, tokens, o, stream);
}

int attention_launch(const void* qkv_bf16, int n_images, int tokens, int heads, int head_dim, void* out_bf16,
                     cudaStream_t stream, bool causal) {
    if (!qkv_bf16 || !out_bf16) return CLIPPPO_ERR_NULL;
    if (n_images <= 0 || tokens <= 0 || heads <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (head_dim != DH) return CLIPPPO_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(qkv_bf16) % 16) || (reinterpret_cast<uintptr_t>(out_bf16) % 16)) return CLIPPPO_ERR_ALIGN;
    // ViT-L/14 (T = 257) and every other unmasked 64 < T <= 257: the tcgen05 / TMEM / TMA kernel (attention_tc.cu);
    // CLIPPPO_ATT_TC=0 keeps the mma.sync kernel below for A/B measurements
    static const bool use_tc = [] { const char* e = getenv("CLIPPPO_ATT_TC"); return !(e && e[0] == '0'); }();
    if (use_tc && attention_tc_supported(tokens, causal)) {
        const int st = attention_tc_launch(qkv_bf16, n_images, tokens, heads, out_bf16, stream);
        if (st == CLIPPPO_OK) prof_count_launch();
        return st;
    }
    // T <= 64 (ViT-B/32: T = 50): two images per tcgen05 tile; causal T <= 128 (text tower: T = 77): one sequence per tile.
    // CLIPPPO_ATT_TC50=0 keeps the mma.sync kernels below (A/B measurements); read per call so that tests can switch it.
    {
        const char* e = getenv("CLIPPPO_ATT_TC50");
        const bool on = use_tc && !(e && e[0] == '0');
        if (on && attention_tc_pair_supported(tokens, causal)) {
            const int st = attention_tc_pair_launch(qkv_bf16, n_images, tokens, heads, out_bf16, stream, causal);
            if (st == CLIPPPO_OK) prof_count_launch();
            return st;
        }
    }
    if (tokens > TP || causal) {                     // text tower (T = 77, causal), T > 257: one CTA per (sequence, head), K / V resident
        if (static_cast<long long>(n_images) * heads > 0x7fffffffLL) return CLIPPPO_ERR_BAD_SHAPE;
        const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv_bf16);
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
        int st = CLIPPPO_ERR_UNSUPPORTED;            // head counts of the CLIP towers: 8 (text), 12 (B/32, B/16), 16 (L/14)
        if (causal) {
            if (heads == 8) st = launch_general<8, true>(q, n_images, tokens, o, stream);
            else if (heads == 12) st = launch_general<12, true>(q, n_images, tokens, o, stream);
            else if (heads == 16) st = launch_general<16, true>(q, n_images, tokens, o, stream);
        } else {
            if (heads == 8) st = launch_general<8, false>(q, n_images, tokens, o, stream);
            else if (heads == 12) st = launch_general<12, false>(q, n_images, tokens, o, stream);
            else if (heads == 16) st = launch_general<16, false>(q, n_images, tokens, o, stream);
        }
        if (st == CLIPPPO_OK) prof_count_launch();
        return st;
    }
    const long long items = static_cast<long long>(n_images) * heads;
    if (items > 0x7fffffffLL) return CLIPPPO_ERR_BAD_SHAPE;
    const int grid = static_cast<int>(items < kNumSMs * ATT_CTAS_PER_SM ? items : kNumSMs * ATT_CTAS_PER_SM);
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv_bf16);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
    if (heads == 12) {
        static DeviceOnce configured;
        if (configured.first_use()) {
            CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(attention_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        }
        CLIPPPO_CUDA_TRY(launch_pdl(attention_kernel<12>, grid, ATT_THREADS, ATT_SMEM_BYTES, stream, 1, q, static_cast<int>(items), tokens, o));
    } else if (heads == 16) {
        static DeviceOnce configured;
        if (configured.first_use()) {
            CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(attention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        }
        CLIPPPO_CUDA_TRY(launch_pdl(attention_kernel<16>, grid, ATT_THREADS, ATT_SMEM_BYTES, stream, 1, q, static_cast<int>(items), tokens, o));
    } else {
        return CLIPPPO_ERR_UNSUPPORTED;                      // head counts of the CLIP towers: 12 (B/32, B/16), 16 (L/14)
    }
    prof_count_launch();
    return CLIPPPO_OK;
}

}  // namespace clipppo

extern "C" int clipppo_attention_bf16(const void* qkv_bf16, int n_images, int tokens, int heads, int head_dim,
                                      void* out_bf16, clipppo_stream_t stream) {
    return clipppo::attention_launch(qkv_bf16, n_images, tokens, heads, head_dim, out_bf16, clipppo::as_stream(stream), false);
}

extern "C" int clipppo_attention_causal_bf16(const void* qkv_bf16, int n_seqs, int tokens, int heads, int head_dim,
                                             void* out_bf16, clipppo_stream_t stream) {
    return clipppo::attention_launch(qkv_bf16, n_seqs, tokens, heads, head_dim, out_bf16, clipppo::as_stream(stream), true);
}
