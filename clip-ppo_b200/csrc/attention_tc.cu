// V4 (long sequences) - softmax(Q K^T / sqrt(d)) V on the 5th-generation tensor cores:
// tcgen05.mma with the accumulators in TMEM, operands staged by TMA.  Serves ViT-L/14 (T = 257, 16 heads)
// and any 64 < T <= 257 of the image tower; the SDPA call inside [clip] nn.MultiheadAttention that
// reference shared/clip_ppo_utils.py:163 / :213-215 reaches through encode_image.
//
// Work item = one (image, head).  One persistent CTA per SM; Q / K / V of an item (96 KB) are double-buffered
// in shared memory, so the TMA loads of item i+1 run under the arithmetic of item i.  The two 128-row query
// tiles of an item belong to two softmax warpgroups that ping-pong on the tensor core:
//   warps 0-3   softmax of query tile 0, warps 4-7 of query tile 1: thread = query row (TMEM lane).  Per 128-key
//               block: tcgen05.ld the 128 x 128 fp32 scores, row maximum, exp2, pack to bf16 and tcgen05.st P over
//               the first 64 columns of S; per item: read the tile's two output accumulators, merge, normalise,
//               stage in the (dead) Q tile under the TMA swizzle, TMA store.
//   warp 8      TMEM allocation (all 512 columns: S of both tiles 2 x 128, four 64-column output accumulators);
//               the warp walks the schedule converged and one elected lane issues every tcgen05.mma (operands in uniform
//               registers: no R2UR loops between the hand-offs, 749 -> 654 us at 1024 images x 16 heads):  S_w = Q_w K_kb^T  (A, B from shared memory, both K-major
//               SWIZZLE_128B);  O_w,kb = P_w V_kb  (A = P from TMEM, B = V from shared memory, MN-major: V stays
//               [key][d] exactly as the TMA box wrote it - no transpose anywhere).  While warpgroup 0 exponentiates
//               block (0, kb) the tensor core computes S of tile 1, and so on.
//   warp 9      issues the TMA loads (same converged / elected scheme), up to two items ahead.
//   warps 10-11 the "tail" query row (below), one key block each, on the CUDA cores.
// T = 257 = 2 x 128 + 1: the 256 x 256 part of the problem maps onto M = N = 128 tensor-core tiles without
// padding; the 257th token would cost a third, almost empty, 128-row tile and a third key block.  Instead its
// key enters every row's softmax as one extra score computed by the row's own thread (a 64-long dot product
// against a broadcast shared-memory row), its value row is added in the epilogue, and its query row is the job of
// warps 10-11 - 0.4 % of the FLOPs.
// Both key blocks of a tile accumulate into one TMEM accumulator under one running shift per row; the shift (and with
// it the accumulator row) is only touched when a later 64-key half beats it by more than 2^8 - see the softmax role.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "gemm.cuh"

namespace clipppo {

using namespace ptx;

namespace {

constexpr int DH = 64;
constexpr int ATC_THREADS = 384;                        // 8 softmax warps, MMA warp, TMA warp, 2 tail-row warps
constexpr int W_MMA = 8, W_TMA = 9, W_TAIL = 10;
constexpr int TILE = 128;                               // query rows per tile = keys per block
constexpr int TILE_BYTES = TILE * DH * 2;               // 16 KB: one TMA box of 128 rows x 128 B
constexpr int OFF_Q = 0, OFF_K = 2 * TILE_BYTES, OFF_V = 4 * TILE_BYTES;
constexpr int OFF_TAIL = 6 * TILE_BYTES;                // 3 x 1 KB: q / k / v of the tail token (row 0 of an 8-row box)
constexpr int STAGE_BYTES = OFF_TAIL + 3 * 1024;        // one item: 99 KB
constexpr int OFF_PART = 2 * STAGE_BYTES;               // 2 x 32 lanes x float4: warp 11's partial tail row (max, sum, o0, o1)
constexpr int OFF_BAR = OFF_PART + 2 * 32 * 16;
enum { B_QK = 0, B_V = 2, B_FREE = 4, B_S = 6, B_P = 8, B_O = 10, B_OREAD = 12, NUM_BARS = 14 };   // two of each
constexpr int ATC_SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;   // + tmem slot + alignment slack
constexpr uint32_t ATC_TMEM_COLS = 512;                 // S / P of tile w: 128 w;  O of tile w: 256 + 64 w;  tail-key scores of tile w: 384 + 16 w
constexpr int MAX_T = 2 * TILE + 1;

#ifndef ATC_TIMING
#define ATC_TIMING 0
#endif

struct AtcArgs {
    int n_items, T, heads;
    __nv_bfloat16* out;      // [n * T, heads * 64]: the tail row is written with plain stores
    uint32_t v_lbo, v_sbo;   // MN-major descriptor fields of the V operand (16-byte units)
    uint32_t p_kstep_cols;   // TMEM columns per 16-key step of the P operand
    uint32_t wg1_delay_ns;
    long long* dbg;          // ATC_TIMING builds: per-phase clock totals of CTA 0's first softmax warps
};

// mbarrier waits.  try_wait returns after a hardware-bounded nap (~50 ns on B200) whether or not the phase has
// completed, so every waiting warp keeps issuing poll instructions on a scheduler it shares with a softmax warp:
//   mbar_wait_hot   the hand-offs on the critical path (S ready, P ready, O ready): bare try_wait loop
//   mbar_wait_cold  everything that runs ahead of the consumers (TMA producer, tail-row warps, operand arrival):
//                   naps a microsecond between polls
// Both trap instead of hanging if a protocol bug leaves the barrier incomplete.
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_cold(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(500);
        if (++spins > (1u << 22)) __trap();
    }
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 3-D tiled TMA load / store (column, token, image)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major operand under the 128-byte swizzle: rows of 64 elements (128 B) = the MN extent, 8 consecutive
// K indices per 1024-byte group; LBO = distance between 64-element MN blocks (unused for N = 64),
// SBO = distance between 8-row K groups.  Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo16 & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(sbo16 & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// the row's dot product with a broadcast 64-element bf16 row
__device__ __forceinline__ float dot64_bf16(const uint8_t* row, int sw, const uint8_t* bcast) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 a = *reinterpret_cast<const uint4*>(row + ((c ^ sw) << 4));
        const uint4 b = *reinterpret_cast<const uint4*>(bcast + c * 16);
        acc0 = fmaf(bf16lo(a.x), bf16lo(b.x), acc0); acc1 = fmaf(bf16hi(a.x), bf16hi(b.x), acc1);
        acc2 = fmaf(bf16lo(a.y), bf16lo(b.y), acc2); acc3 = fmaf(bf16hi(a.y), bf16hi(b.y), acc3);
        acc0 = fmaf(bf16lo(a.z), bf16lo(b.z), acc0); acc1 = fmaf(bf16hi(a.z), bf16hi(b.z), acc1);
        acc2 = fmaf(bf16lo(a.w), bf16lo(b.w), acc2); acc3 = fmaf(bf16hi(a.w), bf16hi(b.w), acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

// softmax(Q K^T / 8) V for sequences of 65 .. 257 tokens; see the header comment for the roles.
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_tail,
                    const __grid_constant__ CUtensorMap tm_out, const AtcArgs g) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET from the __shared__ symbol: the pointer keeps its address space, so the staging-tile
    // accesses are STS / LDS instead of generic ST.E / LD.E
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NUM_BARS * 8);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform

    pdl_launch_dependents();
    const int T = g.T, heads = g.heads, D = heads * DH;
    const bool tail = (T == MAX_T);
    const int t_main = tail ? 2 * TILE : T;                       // tokens on the tensor-core path
    const int n_tiles = (t_main + TILE - 1) / TILE;               // query tiles = key blocks (1 or 2)
    const int last_keys = t_main - (n_tiles - 1) * TILE;          // valid keys of the last block
    const int last_n = (last_keys + 15) & ~15;                    // its MMA N / K extent

    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tm_qkv);
        prefetch_tmap(&tm_tail);
        prefetch_tmap(&tm_out);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_QK + i), 1);
            mbar_init(bar(B_V + i), 1);
            mbar_init(bar(B_FREE + i), 4 * n_tiles + (tail ? 1 : 0));
            mbar_init(bar(B_S + i), 1);
            mbar_init(bar(B_P + i), 4);
            mbar_init(bar(B_O + i), 1);
            mbar_init(bar(B_OREAD + i), 4);
        }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), ATC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();

    const float sl2 = 0.125f * 1.4426950408889634f;               // 1/sqrt(64) * log2(e)

    if (warp == W_TMA) {
        // ===================== TMA producer: two items ahead of the consumers =====================
        // (this warp and the MMA warp walk their schedules converged; only the issue sits under elect.sync, so the operands of
        // the single-thread instructions stay in uniform registers - see elect_one_sync() in tc_ptx.cuh)
        {
            int it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const int img = item / heads, head = item - img * heads;
                const int st = it & 1, use = it >> 1;
                const uint32_t base = sbase + st * STAGE_BYTES;
                if (use > 0) mbar_wait_cold(bar(B_FREE + st), (use - 1) & 1);     // every reader of the stage's previous item is done
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(bar(B_QK + st), 2u * n_tiles * TILE_BYTES + (tail ? 3 * 1024 : 0));
                    for (int t = 0; t < n_tiles; ++t) {
                        tma_load_3d(base + OFF_Q + t * TILE_BYTES, &tm_qkv, bar(B_QK + st), head * DH, t * TILE, img);
                        tma_load_3d(base + OFF_K + t * TILE_BYTES, &tm_qkv, bar(B_QK + st), D + head * DH, t * TILE, img);
                    }
                    if (tail) {
                        tma_load_3d(base + OFF_TAIL, &tm_tail, bar(B_QK + st), head * DH, 2 * TILE, img);
                        tma_load_3d(base + OFF_TAIL + 1024, &tm_tail, bar(B_QK + st), D + head * DH, 2 * TILE, img);
                        tma_load_3d(base + OFF_TAIL + 2048, &tm_tail, bar(B_QK + st), 2 * D + head * DH, 2 * TILE, img);
                    }
                    mbar_arrive_expect_tx(bar(B_V + st), static_cast<uint32_t>(n_tiles) * TILE_BYTES);
                    for (int t = 0; t < n_tiles; ++t)
                        tma_load_3d(base + OFF_V + t * TILE_BYTES, &tm_qkv, bar(B_V + st), 2 * D + head * DH, t * TILE, img);
                }
                __syncwarp();
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        {
            const uint32_t idesc_pv = make_idesc_bf16(TILE, DH) | (1u << 16);     // B (= V) is MN-major
            uint32_t ph_p0 = 0, ph_p1 = 0;
            int it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const int st = it & 1, use = it >> 1;
                const uint32_t base = sbase + st * STAGE_BYTES;
                auto issue_s = [&](int w, int kb) {                // S_w = Q_w K_kb^T
                    const int nk = (kb == n_tiles - 1) ? last_n : TILE;
                    const uint64_t dq = make_kmajor_sw128_desc(base + OFF_Q + w * TILE_BYTES);
                    const uint64_t dk = make_kmajor_sw128_desc(base + OFF_K + kb * TILE_BYTES);
                    const uint32_t idesc_s = make_idesc_bf16(TILE, nk);
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + 128 * w, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    umma_commit(bar(B_S + w));
                };
                mbar_wait_cold(bar(B_QK + st), use & 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    for (int w = 0; w < n_tiles; ++w) issue_s(w, 0);
                    if (tail) {
                        // the tail key's score for all 256 query rows: Q_w x (k_tail box)^T, N = 16 (row 0 of the box is the key, rows 1-7
                        // are zero-filled, the second 8-row group aliases the first), into 16 spare TMEM columns per tile.  Complete
                        // before the softmax asks for it: every later commit on B_S covers these MMAs.
                        const uint64_t dkt = make_kmajor_sw128_desc(base + OFF_TAIL + 1024) & ~(static_cast<uint64_t>(0x3FFF) << 32);   // SBO = 0
                        const uint32_t idesc_t = make_idesc_bf16(TILE, 16);
                        for (int w = 0; w < 2; ++w) {
                            const uint64_t dq = make_kmajor_sw128_desc(base + OFF_Q + w * TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + 384 + 16 * w, dq + 2 * k, dkt + 2 * k, idesc_t, k != 0);
                        }
                    }
                }
                __syncwarp();
                mbar_wait_cold(bar(B_V + st), use & 1);
                tc_fence_after();
                for (int kb = 0; kb < n_tiles; ++kb) {
                    const int nk = (kb == n_tiles - 1) ? last_n : TILE;
                    for (int w = 0; w < n_tiles; ++w) {
                        if (kb == 0 && it > 0) mbar_wait_hot(bar(B_OREAD + w), (it - 1) & 1);   // the previous item's accumulators were read
                        mbar_wait_hot(bar(B_P + w), w ? ph_p1 : ph_p0);                          // P_w is in TMEM
                        if (w) ph_p1 ^= 1; else ph_p0 ^= 1;
                        tc_fence_after();
                        const uint32_t v_addr = base + OFF_V + kb * TILE_BYTES;
                        if (elect_one_sync()) {
                            for (int k = 0; k < nk / 16; ++k)                       // O_w (+)= P_w V_kb: one accumulator for both key blocks
                                umma_bf16_ts(tmem_base + 256 + 64 * w, tmem_base + 128 * w + g.p_kstep_cols * k,
                                             make_mnmajor_sw128_desc(v_addr + k * 2048, g.v_lbo, g.v_sbo), idesc_pv, (kb | k) != 0);
                            if (kb + 1 < n_tiles) issue_s(w, kb + 1);          // in order behind the MMAs that read P_w
                            else {
                                if (tail)      // + p_t v_tail: A = the (p_t, 0, ...) columns, B = the v_tail box (row 0; both 8-row groups alias it)
                                    umma_bf16_ts(tmem_base + 256 + 64 * w, tmem_base + 128 * w + 64,
                                                 make_mnmajor_sw128_desc(base + OFF_TAIL + 2048, g.v_lbo, 0), idesc_pv, 1);
                                umma_commit(bar(B_O + w));
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp >= W_TAIL) {
        // ===================== tail query row (token 256) on the CUDA cores: warp 10 keys 0..127, warp 11 keys 128..256 =====================
        if (tail) {
            const int half = warp - W_TAIL;
            int it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const int img = item / heads, head = item - img * heads;
                const int st = it & 1, use = it >> 1;
                const uint8_t* sm = smem + st * STAGE_BYTES;
                mbar_wait_cold(bar(B_QK + st), use & 1);
                float s[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = lane + 32 * i;                  // key 128 half + j
                    s[i] = dot64_bf16(sm + OFF_K + half * TILE_BYTES + j * 128, j & 7, sm + OFF_TAIL);
                }
                float s_t = -INFINITY;
                if (half == 1) s_t = dot64_bf16(sm + OFF_TAIL + 1024, 0, sm + OFF_TAIL);      // the tail key (broadcast reads)
                float mx = s_t;
#pragma unroll
                for (int i = 0; i < 4; ++i) mx = fmaxf(mx, s[i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float mb = mx * sl2;
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) { s[i] = ex2f(fmaf(s[i], sl2, -mb)); sum += s[i]; }
                sum = warp_sum(sum);
                const float p_t = (half == 1) ? ex2f(fmaf(s_t, sl2, -mb)) : 0.f;
                sum += p_t;
                // O[d] = sum_j p_j V[j][d]; lane owns d = 2 lane, 2 lane + 1 (4 bytes of every V row)
                mbar_wait_cold(bar(B_V + st), use & 1);
                float o0 = 0.f, o1 = 0.f;
                const int cch = lane >> 2, cof = (lane & 3) * 4;
                const uint8_t* vb = sm + OFF_V + half * TILE_BYTES + cof;
                float o2 = 0.f, o3 = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int jo = 0; jo < 4; ++jo) {
#pragma unroll
                        for (int ji = 0; ji < 8; ++ji) {          // row j = 32 i + 8 jo + ji: j & 7 = ji is a compile-time constant
                            const float pj = __shfl_sync(0xffffffffu, s[i], 8 * jo + ji);
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(vb + (32 * i + 8 * jo + ji) * 128 + ((cch ^ ji) << 4));
                            if (ji & 1) { o2 = fmaf(pj, bf16lo(w), o2); o3 = fmaf(pj, bf16hi(w), o3); }
                            else { o0 = fmaf(pj, bf16lo(w), o0); o1 = fmaf(pj, bf16hi(w), o1); }
                        }
                    }
                }
                o0 += o2; o1 += o3;
                float4* part = reinterpret_cast<float4*>(smem + OFF_PART) + st * 32 + lane;
                if (half == 1) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(sm + OFF_TAIL + 2048 + (cch << 4) + cof);
                    o0 = fmaf(p_t, bf16lo(w), o0);
                    o1 = fmaf(p_t, bf16hi(w), o1);
                    *part = make_float4(mb, sum, o0, o1);
                }
                asm volatile("bar.sync 1, 64;" ::: "memory");       // warps 10 and 11
                if (half == 0) {
                    const float4 p = *part;
                    const float m = fmaxf(mb, p.x);
                    const float w0 = ex2f(mb - m), w1 = ex2f(p.x - m);
                    const float inv = 1.0f / (sum * w0 + p.y * w1);
                    __nv_bfloat16* dst = g.out + (static_cast<size_t>(img) * T + 2 * TILE) * D + head * DH;
                    reinterpret_cast<uint32_t*>(dst)[lane] = pack_bf16x2((o0 * w0 + p.z * w1) * inv, (o1 * w0 + p.w * w1) * inv);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(B_FREE + st));
                }
            }
        }
    } else if (warp < 4 * n_tiles) {
        // ===================== softmax + epilogue: warpgroup w owns query tile w, thread = query row =====================
        // One running shift m per row for the whole item (log2 domain, scores already scaled): the accumulator O_w
        // takes both key blocks.  m only moves when a 64-key half raises the maximum by more than 2^8 (P stays far
        // inside the bf16 / fp32 range up to there); then everything already produced under the old shift - the
        // packed first half of the block, the row sum, the accumulator row in TMEM - is scaled by f = bf16(2^(m - m'))
        // and m becomes m - log2(f), the shift for which those scaled values are exact.
        const int w = warp >> 2;
        const int r = (warp & 3) * 32 + lane;                      // row inside the tile = TMEM lane
        const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 128 * w;      // my S / P row
        const uint32_t orow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + 64 * w;  // my accumulator row
        constexpr float kLazy = 8.0f;
        uint32_t ph_s = 0;
        int it = 0;
#if ATC_TIMING
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#define ATC_TICK(i) do { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; } while (0)
#else
#define ATC_TICK(i) do { } while (0)
#endif
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
            const int img = item / heads, head = item - img * heads;
            const int st = it & 1, use = it >> 1;
            uint8_t* sm = smem + st * STAGE_BYTES;
            float s_t = 0.f;                                      // the tail key's score for my row
            float m = 0.f, l = 0.f;                               // running shift and row sum
            for (int kb = 0; kb < n_tiles; ++kb) {
                const bool last = (kb == n_tiles - 1);
                const int nvalid = last ? last_keys : TILE;
                const bool ragged = nvalid != TILE;
                ATC_TICK(0);
                mbar_wait_hot(bar(B_S + w), ph_s); ph_s ^= 1;
                tc_fence_after();
                ATC_TICK(1);
                if (tail && last) {                               // the tail key's score: column 0 of the 16-column MMA below S and O
                    uint32_t t1;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(t1) : "r"(trow - 128 * w + 384 + 16 * w) : "memory");
                    tmem_ld_wait();
                    s_t = __uint_as_float(t1) * sl2;
                }
                // Four 32-key chunks, double-buffered: chunk c + 1 streams out of TMEM while chunk c is exponentiated
                // against the running shift - no maximum has to be known first (lazy shift, see above).
                uint32_t v[2][32], pk[64];
                const int nch = (nvalid + 31) >> 5;
                float2 sum2 = make_float2(0.f, 0.f);
                const float2 sc2 = make_float2(sl2, sl2);
                float fo = 1.0f;                                  // factor owed to the accumulator row
                tmem_ld_32x32(trow, v[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nch) {
                        tmem_ld_wait();
                        if (c + 1 < nch) tmem_ld_32x32(trow + (c + 1) * 32, v[(c + 1) & 1]);
                        uint32_t (&vc)[32] = v[c & 1];
                        if (ragged) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j >= nvalid) vc[j] = 0xff800000u;                // -inf: keys beyond T
                        }
                        // the chunk's maximum: needed BEFORE the exponentials only for the very first chunk of the item; everywhere
                        // else the exponentials run against the current shift at once (nothing to wait for) and the maximum is
                        // checked afterwards - a chunk that beat the shift by more than 2^8 (rare) is redone
                        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int j = 0; j < 32; ++j) mxa[j & 3] = fmaxf(mxa[j & 3], __uint_as_float(vc[j]));
                        float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * sl2;
                        if (tail && last && c == nch - 1) mx = fmaxf(mx, s_t);
                        if (kb == 0 && c == 0) m = mx;
                        const float2 sum_before = sum2;
                        {
                            const float2 nm2 = make_float2(-m, -m);
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float2 x = __ffma2_rn(make_float2(__uint_as_float(vc[j]), __uint_as_float(vc[j + 1])), sc2, nm2);
                                const float2 e = make_float2(ex2f(x.x), ex2f(x.y));
                                sum2 = __fadd2_rn(sum2, e);
                                pk[16 * c + (j >> 1)] = pack_bf16x2(e.x, e.y);
                            }
                        }
                        if (!(kb == 0 && c == 0) && __any_sync(0xffffffffu, mx > m + kLazy)) {
                            // rare: move the shift; whatever this row already produced under the old one is scaled by f, this chunk redone
                            __nv_bfloat16 fb = __float2bfloat16_rn(1.0f);
                            if (mx > m + kLazy) fb = __float2bfloat16_rn(ex2f(m - mx));
                            const float ff = __bfloat162float(fb);
                            const __nv_bfloat162 f2 = __halves2bfloat162(fb, fb);
#pragma unroll
                            for (int j = 0; j < 16 * c; ++j) {
                                __nv_bfloat162 hv = *reinterpret_cast<__nv_bfloat162*>(&pk[j]);
                                hv = __hmul2(hv, f2);
                                pk[j] = *reinterpret_cast<uint32_t*>(&hv);
                            }
                            l *= ff; fo *= ff;
                            if (ff != 1.0f) {
                                m = (ff == 0.f) ? mx : m - __log2f(ff);
                                sum2 = make_float2(sum_before.x * ff, sum_before.y * ff);
                                const float2 nm2 = make_float2(-m, -m);
#pragma unroll
                                for (int j = 0; j < 32; j += 2) {
                                    const float2 x = __ffma2_rn(make_float2(__uint_as_float(vc[j]), __uint_as_float(vc[j + 1])), sc2, nm2);
                                    const float2 e = make_float2(ex2f(x.x), ex2f(x.y));
                                    sum2 = __fadd2_rn(sum2, e);
                                    pk[16 * c + (j >> 1)] = pack_bf16x2(e.x, e.y);
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[16 * c + j] = 0u;
                    }
                }
                ATC_TICK(2);
                {
                    uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[0]);
                    uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[32]);
                    tmem_st_32x32(trow, lo);
                    if (nvalid > 64) tmem_st_32x32(trow + 32, hi);
                }
                l += sum2.x + sum2.y;
                if (tail && last) {
                    // the tail key: its probability goes into 8 spare columns of my S row as the bf16 pair (p_t, 0) followed by
                    // zeros - a 16-key A operand whose only non-zero entry multiplies row 0 of the v_tail box (one extra MMA)
                    const float p_t = ex2f(s_t - m);
                    l += p_t;
                    uint32_t pt[8] = {pack_bf16x2(p_t, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                 ::"r"(trow + 64), "r"(pt[0]), "r"(pt[1]), "r"(pt[2]), "r"(pt[3]), "r"(pt[4]), "r"(pt[5]), "r"(pt[6]), "r"(pt[7])
                                 : "memory");
                }
                // the accumulator row (holding the first key block) follows the shift; the whole warp moves together
                if (kb > 0 && __any_sync(0xffffffffu, fo != 1.0f)) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t oo[32];
                        tmem_ld_32x32(orow + hh * 32, oo);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) oo[j] = __float_as_uint(__uint_as_float(oo[j]) * fo);
                        tmem_st_32x32(orow + hh * 32, oo);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_P + w));
                ATC_TICK(3);
            }
            // ---- epilogue of the tile: O / l ----
            mbar_wait_hot(bar(B_O + w), it & 1);
            tc_fence_after();
            ATC_TICK(4);
            const float inv = 1.0f / l;
            const float2 inv2 = make_float2(inv, inv);
            uint8_t* stage = sm + OFF_Q + w * TILE_BYTES + r * 128;         // my row of the (dead) Q tile
            uint32_t oa[2][32];
            tmem_ld_32x32(orow, oa[0]);
            tmem_ld_32x32(orow + 32, oa[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_OREAD + w));       // the accumulator is in registers
            ATC_TICK(5);
#pragma unroll
            for (int c = 0; c < 8; ++c) {                       // 16-byte pieces: columns 8 c .. 8 c + 7
                float2 o2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    o2[e] = __fmul2_rn(make_float2(__uint_as_float(oa[c >> 2][8 * (c & 3) + 2 * e]),
                                                   __uint_as_float(oa[c >> 2][8 * (c & 3) + 2 * e + 1])), inv2);
                const uint4 pk = make_uint4(pack_bf16x2(o2[0].x, o2[0].y), pack_bf16x2(o2[1].x, o2[1].y),
                                            pack_bf16x2(o2[2].x, o2[2].y), pack_bf16x2(o2[3].x, o2[3].y));
                *reinterpret_cast<uint4*>(stage + ((c ^ (r & 7)) << 4)) = pk;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&tm_out, sbase + st * STAGE_BYTES + OFF_Q + w * TILE_BYTES + (warp & 3) * 32 * 128, head * DH,
                             w * TILE + (warp & 3) * 32, img);
                bulk_commit_group();
                bulk_wait_group_read<0>();                      // the staged rows have left shared memory: the stage may be refilled
                mbar_arrive(bar(B_FREE + st));
            }
            __syncwarp();
            ATC_TICK(6);
        }
#if ATC_TIMING
        if (g.dbg && blockIdx.x == 0 && lane == 0 && (warp & 3) == 0) {
            for (int i = 0; i < 8; ++i) g.dbg[w * 8 + i] = tacc[i];
        }
#endif
        if (lane == 0) bulk_wait_group<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATC_TMEM_COLS);
    }
}

// =============================================================================================
// Short sequences (T <= 64; ViT-B/32: T = 50) on the same machinery, two images per tensor-core tile.
//
// At T = 50 a head is a 50 x 50 x 64 problem: the kernel is bound by HBM (25.6 KB per (image, head)) and, at the
// power-capped clock of a real step, by instruction issue - the mma.sync version (attention.cu) executes ~1600 warp
// instructions per item.  Here a work item is (image PAIR, head): the two images' Q / K / V slices land by TMA in
// rows 0 .. T-1 and 64 .. 64+T-1 of 128-row tiles (the other rows stay zero), ONE 128 x 128 x 64 tcgen05.mma
// computes both score blocks (the off-diagonal blocks are never read), thread = query row reads its own image's
// 64 columns straight out of TMEM, exponentiates T of them, writes P back as a block-diagonal bf16 operand
// (zeros in the other image's half), and O = P V is one more MMA chain over the 128 keys with V as the MN-major
// shared-memory operand.  ~400 warp instructions per item.
//   warps 0-3 / 4-7   two softmax warpgroups, one per TMEM slot: warpgroup g owns the CTA's items g, g + 2, ...
//   warp 8            lane 0 issues every tcgen05.mma: S(i), then P V of item i - 1
//   warp 9            lane 0 issues the TMA loads, up to four items ahead (4 x 48 KB stages)
constexpr int TP_STAGES = 4;
constexpr int TP_STAGE_BYTES = 3 * TILE_BYTES;                  // Q | K | V tiles of one item: 48 KB
constexpr int TP_OFF_BAR = TP_STAGES * TP_STAGE_BYTES;
enum { PB_LOAD = 0, PB_FREE = TP_STAGES, PB_S = 2 * TP_STAGES, PB_P = 2 * TP_STAGES + 2, PB_O = 2 * TP_STAGES + 4,
       PB_TFREE = 2 * TP_STAGES + 6, TP_NUM_BARS = 2 * TP_STAGES + 8 };
constexpr int TP_SMEM_BYTES = TP_OFF_BAR + TP_NUM_BARS * 8 + 16 + 1024;
constexpr int TP_THREADS = 320;
constexpr int TP_SLOT_COLS = 192;                               // per TMEM slot: S / P 128 columns, O 64 columns

struct TpArgs {
    int n_items, T, heads, causal;
    uint32_t v_lbo, v_sbo, p_kstep_cols;
};

// PAIR = true: T <= 64, work item = (image pair, head).  PAIR = false: 64 < T <= 128, work item = (sequence, head) in rows
// 0 .. T-1 of the tile - the text tower (T = 77) under its causal mask ([clip] build_attention_mask: query t sees keys 0 .. t).
template <bool PAIR>
__global__ void __launch_bounds__(TP_THREADS, 1)
attention_tc_pair_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, const TpArgs g) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET from the __shared__ symbol: the pointer keeps its address space, so the staging-tile
    // accesses are STS / LDS instead of generic ST.E / LD.E
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    auto bar = [&](int i) { return sbase + TP_OFF_BAR + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TP_OFF_BAR + TP_NUM_BARS * 8);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    constexpr int WP_MMA = 8, WP_TMA = 9;

    pdl_launch_dependents();
    const int T = g.T, heads = g.heads, D = heads * DH;
    if (warp == WP_TMA && lane == 0) {
        prefetch_tmap(&tm_in);
        prefetch_tmap(&tm_out);
        for (int i = 0; i < TP_STAGES; ++i) { mbar_init(bar(PB_LOAD + i), 1); mbar_init(bar(PB_FREE + i), 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(PB_S + i), 1); mbar_init(bar(PB_P + i), 4); mbar_init(bar(PB_O + i), 1); mbar_init(bar(PB_TFREE + i), 4);
        }
        fence_barrier_init();
    }
    if (warp == WP_MMA) tmem_alloc(smem_u32(tmem_slot), ATC_TMEM_COLS);
    // rows T .. 63 and 64 + T .. 127 of every tile are never written by a TMA box: zero everything once
    for (int i = threadIdx.x; i < TP_OFF_BAR / 16; i += TP_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();

    if (warp == WP_TMA) {
        {
            int it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const int pair = item / heads, head = item - pair * heads;
                const int stg = it % TP_STAGES, use = it / TP_STAGES;
                if (use > 0) mbar_wait_cold(bar(PB_FREE + stg), (use - 1) & 1);
                const uint32_t base = sbase + stg * TP_STAGE_BYTES;
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(bar(PB_LOAD + stg), (PAIR ? 6u : 3u) * static_cast<uint32_t>(T) * 128u);
                    if constexpr (PAIR) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {             // an image beyond the batch (odd n) is zero-filled by the TMA unit
#pragma unroll
                            for (int m = 0; m < 3; ++m)
                                tma_load_3d(base + m * TILE_BYTES + h * 64 * 128, &tm_in, bar(PB_LOAD + stg), m * D + head * DH, 0, 2 * pair + h);
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 3; ++m) tma_load_3d(base + m * TILE_BYTES, &tm_in, bar(PB_LOAD + stg), m * D + head * DH, 0, pair);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == WP_MMA) {
        {
            constexpr uint32_t idesc_s = make_idesc_bf16(TILE, TILE);
            const uint32_t idesc_pv = make_idesc_bf16(TILE, DH) | (1u << 16);     // B (= V) is MN-major
            auto issue_pv = [&](int j) {                       // O_j = P_j V_j over all 128 key rows of the tile
                const int slot = j & 1, stg = j % TP_STAGES;
                mbar_wait_hot(bar(PB_P + slot), (j >> 1) & 1);
                tc_fence_after();
                const uint32_t v_addr = sbase + stg * TP_STAGE_BYTES + 2 * TILE_BYTES;
                const uint32_t ts = tmem_base + TP_SLOT_COLS * slot;
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < TILE / 16; ++k)
                        umma_bf16_ts(ts + 128, ts + g.p_kstep_cols * k, make_mnmajor_sw128_desc(v_addr + k * 2048, g.v_lbo, g.v_sbo), idesc_pv, k != 0);
                    umma_commit(bar(PB_O + slot));
                }
                __syncwarp();
            };
            int it = 0;
            for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
                const int slot = it & 1, stg = it % TP_STAGES;
                mbar_wait_cold(bar(PB_LOAD + stg), (it / TP_STAGES) & 1);
                if (it >= 2) mbar_wait_hot(bar(PB_TFREE + slot), ((it >> 1) - 1) & 1);      // the slot's previous item has left TMEM
                tc_fence_after();
                const uint32_t base = sbase + stg * TP_STAGE_BYTES;
                const uint64_t dq = make_kmajor_sw128_desc(base), dk = make_kmajor_sw128_desc(base + TILE_BYTES);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + TP_SLOT_COLS * slot, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    umma_commit(bar(PB_S + slot));
                }
                __syncwarp();
                if (it >= 1) issue_pv(it - 1);
            }
            if (it >= 1) issue_pv(it - 1);
        }
    } else {
        // ===================== softmax + epilogue: warpgroup wg owns TMEM slot wg, thread = row of the tile =====================
        const int wg = warp >> 2, wq = warp & 3;
        const int r = wq * 32 + lane;                           // tile row = TMEM lane
        const int half = PAIR ? (wq >> 1) : 0;                  // which image of the pair
        const int t0 = PAIR ? (r & 63) : r;                     // token of my row
        const int t = t0;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + TP_SLOT_COLS * wg;
        const float sl2 = 0.125f * 1.4426950408889634f;
        int it = 0;
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++it) {
            if ((it & 1) != wg) continue;
            const int pair = item / heads, head = item - pair * heads;
            const int stg = it % TP_STAGES;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait_hot(bar(PB_S + wg), ph);
            tc_fence_after();
            // my keys: PAIR - the 64 columns of my image; else all 128.  Valid: the first T of them, under the causal mask the first t + 1.
            constexpr int NCH = PAIR ? 2 : 4;
            const uint32_t tkeys = trow + (PAIR ? 64 * half : 0);
            const int nvalid = g.causal ? min(t0 + 1, T) : T;
            float sum = 0.f;
            uint32_t pk[16 * NCH];
            {
                uint32_t v[NCH][32];
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld_32x32(tkeys + 32 * c, v[c]);
                tmem_ld_wait();
                float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int j = 0; j < 32 * NCH; ++j) {
                    if (j >= nvalid) v[j >> 5][j & 31] = 0xff800000u;   // keys beyond T (zero rows of the K tile) / above the diagonal
                    mxa[j & 3] = fmaxf(mxa[j & 3], __uint_as_float(v[j >> 5][j & 31]));
                }
                const float m = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * sl2;
                const float2 nm2 = make_float2(-m, -m), sc2 = make_float2(sl2, sl2);
                float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 32 * NCH; j += 2) {
                    const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[j >> 5][j & 31]), __uint_as_float(v[j >> 5][(j & 31) + 1])), sc2, nm2);
                    const float2 e = make_float2(ex2f(x.x), ex2f(x.y));
                    sum2 = __fadd2_rn(sum2, e);
                    pk[j >> 1] = pack_bf16x2(e.x, e.y);
                }
                sum = sum2.x + sum2.y;
            }
            if constexpr (PAIR) {   // P: keys 0 .. 63 in packed columns 0 .. 31, keys 64 .. 127 in 32 .. 63; the other image's half is zero
                uint32_t z[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) z[j] = 0u;
                tmem_st_32x32(trow + 32 * half, *reinterpret_cast<uint32_t (*)[32]>(&pk[0]));
                tmem_st_32x32(trow + 32 * (1 - half), z);
            } else {
                tmem_st_32x32(trow, *reinterpret_cast<uint32_t (*)[32]>(&pk[0]));
                tmem_st_32x32(trow + 32, *reinterpret_cast<uint32_t (*)[32]>(&pk[32]));
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(PB_P + wg));
            // ---- O / l ----
            mbar_wait_hot(bar(PB_O + wg), ph);
            tc_fence_after();
            uint32_t oa[2][32];
            tmem_ld_32x32(trow + 128, oa[0]);
            tmem_ld_32x32(trow + 160, oa[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(PB_TFREE + wg));      // S / P / O of this item are in registers: the slot may take item + 2
            const float inv = 1.0f / sum;
            const float2 inv2 = make_float2(inv, inv);
            uint8_t* stage = smem + stg * TP_STAGE_BYTES + r * 128;         // my row of the (dead) Q tile
            if (t < T) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float2 o2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        o2[e] = __fmul2_rn(make_float2(__uint_as_float(oa[c >> 2][8 * (c & 3) + 2 * e]),
                                                       __uint_as_float(oa[c >> 2][8 * (c & 3) + 2 * e + 1])), inv2);
                    *reinterpret_cast<uint4*>(stage + ((c ^ (r & 7)) << 4)) =
                        make_uint4(pack_bf16x2(o2[0].x, o2[0].y), pack_bf16x2(o2[1].x, o2[1].y),
                                   pack_bf16x2(o2[2].x, o2[2].y), pack_bf16x2(o2[3].x, o2[3].y));
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                // 32-row slab of image 2 pair + half; tokens >= T and an image beyond the batch are clipped by the TMA unit
                tma_store_3d(&tm_out, sbase + stg * TP_STAGE_BYTES + wq * 32 * 128, head * DH, PAIR ? (wq & 1) * 32 : wq * 32,
                             PAIR ? 2 * pair + half : pair);
                bulk_commit_group();
                bulk_wait_group_read<0>();
                mbar_arrive(bar(PB_FREE + stg));
            }
            __syncwarp();
        }
        if (lane == 0) bulk_wait_group<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATC_TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 [n, T, cols] (token rows of `cols` elements) as boxes of {64 columns, box_rows tokens, 1 image} under the
// 128-byte swizzle; tokens beyond T are zero-filled on loads and clipped on stores.
int make_seq_tmap(CUtensorMap* map, const void* ptr, int n, int T, int cols, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { last_cuda_error_ref() = static_cast<int>(cudaErrorNotSupported); return CLIPPPO_ERR_CUDA; }
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(n)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(T) * cols * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { last_cuda_error_ref() = 10000 + static_cast<int>(r); return CLIPPPO_ERR_CUDA; }
    return CLIPPPO_OK;
}

uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? static_cast<uint32_t>(strtoul(e, nullptr, 0)) : dflt;
}

}  // namespace

bool attention_tc_supported(int tokens, bool causal) { return !causal && tokens > 64 && tokens <= MAX_T; }

int attention_tc_launch(const void* qkv_bf16, int n_images, int tokens, int heads, void* out_bf16, cudaStream_t stream) {
    if (!attention_tc_supported(tokens, false)) return CLIPPPO_ERR_UNSUPPORTED;
    const long long items = static_cast<long long>(n_images) * heads;
    if (items > 0x7fffffffLL) return CLIPPPO_ERR_BAD_SHAPE;
    const int D = heads * DH;
    CUtensorMap tq, tt, to;
    int st = make_seq_tmap(&tq, qkv_bf16, n_images, tokens, 3 * D, TILE);
    if (!st) st = make_seq_tmap(&tt, qkv_bf16, n_images, tokens, 3 * D, 8);
    if (!st) st = make_seq_tmap(&to, out_bf16, n_images, tokens, D, 32);
    if (st) return st;
    static DeviceOnce configured;
    if (configured.first_use()) {
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
    }
    // operand-layout knobs, only ever changed by tools/ probes
    static const uint32_t v_lbo = env_u32("CLIPPPO_ATC_V_LBO", 64), v_sbo = env_u32("CLIPPPO_ATC_V_SBO", 64),
                          p_cols = env_u32("CLIPPPO_ATC_P_COLS", 8), wg1_delay = env_u32("CLIPPPO_ATC_WG1_DELAY", 0);
    AtcArgs g{static_cast<int>(items), tokens, heads, static_cast<__nv_bfloat16*>(out_bf16), v_lbo, v_sbo, p_cols, wg1_delay, nullptr};
#if ATC_TIMING
    static long long* dbg = nullptr;
    if (!dbg) { cudaMalloc(&dbg, 16 * 8); cudaMemset(dbg, 0, 16 * 8); }
    g.dbg = dbg;
#endif
    const int grid = static_cast<int>(items < kNumSMs ? items : kNumSMs);
    CLIPPPO_CUDA_TRY(launch_pdl(attention_tc_kernel, grid, ATC_THREADS, ATC_SMEM_BYTES, stream, 1, tq, tt, to, g));
#if ATC_TIMING
    {
        long long h[16];
        cudaMemcpy(h, g.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[8] = {"loop top", "wait S", "chunks", "st+arrive", "wait O", "ld O", "epilogue", "-"};
        for (int w = 0; w < 2; ++w) for (int i = 0; i < 7; ++i) printf("  wg%d %-10s %lld\n", w, names[i], h[w * 8 + i]);
    }
#endif
    return CLIPPPO_OK;
}

bool attention_tc_pair_supported(int tokens, bool causal) {
    return tokens >= 16 && (causal ? tokens <= TILE : tokens <= 64);      // 64 < T <= 128 without a mask: attention_tc_kernel
}

// T <= 64 (ViT-B/32: T = 50): two images per 128-row tensor-core tile.  64 < T <= 128 causal (text tower: T = 77): one sequence per tile.
int attention_tc_pair_launch(const void* qkv_bf16, int n_images, int tokens, int heads, void* out_bf16, cudaStream_t stream, bool causal) {
    if (!attention_tc_pair_supported(tokens, causal)) return CLIPPPO_ERR_UNSUPPORTED;
    const bool pair = tokens <= 64;
    const long long items = static_cast<long long>(pair ? (n_images + 1) / 2 : n_images) * heads;
    if (items > 0x7fffffffLL) return CLIPPPO_ERR_BAD_SHAPE;
    const int D = heads * DH;
    CUtensorMap ti, to;
    int st = make_seq_tmap(&ti, qkv_bf16, n_images, tokens, 3 * D, tokens);
    if (!st) st = make_seq_tmap(&to, out_bf16, n_images, tokens, D, 32);
    if (st) return st;
    static DeviceOnce configured;
    if (configured.first_use()) {
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(attention_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(attention_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
    }
    static const uint32_t v_lbo = env_u32("CLIPPPO_ATC_V_LBO", 64), v_sbo = env_u32("CLIPPPO_ATC_V_SBO", 64),
                          p_cols = env_u32("CLIPPPO_ATC_P_COLS", 8);
    TpArgs g{static_cast<int>(items), tokens, heads, causal ? 1 : 0, v_lbo, v_sbo, p_cols};
    const int grid = static_cast<int>(items < kNumSMs ? items : kNumSMs);
    if (pair) CLIPPPO_CUDA_TRY(launch_pdl(attention_tc_pair_kernel<true>, grid, TP_THREADS, TP_SMEM_BYTES, stream, 1, ti, to, g));
    else CLIPPPO_CUDA_TRY(launch_pdl(attention_tc_pair_kernel<false>, grid, TP_THREADS, TP_SMEM_BYTES, stream, 1, ti, to, g));
    return CLIPPPO_OK;
}

}  // namespace clipppo
