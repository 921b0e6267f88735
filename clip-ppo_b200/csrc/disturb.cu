// D1 - fused visual disturbance: noise -> contrast -> blur -> cutout, one launch, one HBM pass.
//
// Replaces DisturbanceWrapperGPU.apply_disturbances (reference shared/disturbances_gpu.py:66-73)
// and, through the `stages` mask, each of its four stage methods.  Numerical spec: SURVEY.md
// Appendix A (torchvision gaussian_noise_image / adjust_contrast / gaussian_blur + the
// reference's own cutout).
//
// Data layout.  An image is split into S horizontal stripes; the S CTAs that own them form one
// thread-block cluster.  Each CTA keeps its stripe (all channels, plus 2*(k/2) halo rows) in
// shared memory, so x and noise are read from HBM exactly once and `out` is written once:
// 12 algorithmic bytes per element.  The only cross-stripe dependencies - the per-image gray
// mean of the contrast stage and the halo rows of the vertical blur - travel over distributed
// shared memory inside the cluster, not through HBM.
//
//   phase 1  load x (+ sigma*noise, clamp) -> smem, accumulate the gray sum      [HBM read]
//   phase 2  block reduce, cluster reduce over DSMEM -> per-image mean
//   phase 3  contrast blend in place in smem
//   phase 4  halo rows (reflect at the image border) copied from the owning CTA's smem
//   phase 5  separable blur: horizontal taps from smem, vertical taps from a register ring;
//            cutout predicate on the store                                         [HBM write]
#include <stdlib.h>

#include <type_traits>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace clipppo {

struct DisturbParams {
    const void* x;
    const float* noise;
    void* out;
    long long xs[4], ns[4];
    int B, C, H, W;
    int stages;
    float sigma_n, c, omc;
    float taps[CLIPPPO_MAX_BLUR_TAPS];
    int sh, sw, ph, pw;
    int S, R;       // stripes per image (= cluster size), rows per stripe
    int nsplit;     // row splits of a stripe in the blur phase (balances the 4-column tasks over the CTA)
    int fast;       // x / noise: every image a contiguous [C,H,W] fp32 block (any batch stride that is a multiple of 4
                    // elements, e.g. one frame of an Atari stack), W % 4 == 0, 16-byte aligned
    int batch_contig;  // ... and the images back to back (what the general kernel's vector path assumes)
    int x_u8;       // fast path only: x holds uint8 pixels (contiguous NCHW), read as float(v) / 255
    int nhwc;       // fast path only: x and noise are fp32 NHWC-dense views of [B,3,H,W] (strides H*W*3, 1, W*3, 3: the
                    // MiniGrid call site and the reference benchmark); de-interleaved in registers while loading
    int io_mode;    // 0: fp32 strided in, fp32 NCHW out.  1: u8 NHWC in / u8 NHWC out.
                    // 2: fp32 (0..255) NHWC in / u8 NHWC out.
    int nthreads;   // fast path: CTA size (covers the blur tasks in one round when it can)
    int log2S;      // fast path: S is a power of two
    unsigned magic_nq, magic_nsplit, magic_rs;   // fast path: ceil(2^32 / d), so that __umulhi(n, magic) == n / d for n < 2^16
    float out_scale;   // fast path: the result is multiplied by this on the way out (1: off).  `apply_disturbances(...) * 255`
                       // of the call sites (clip_ppo_atari.py:584, the rollout buffers hold 0..255) without a pass of its own
    int philox;        // fast path, opt-in: no noise tensor - the N(0,1) draws are generated in the kernel (Philox4x32-10
                       // keyed by philox_seed, counter = (logical NCHW quad index, philox_offset), Box-Muller): 8 B / element
    unsigned long long philox_seed, philox_offset;
    long long philox_first;   // global index of image 0 of this call (a shard of a larger batch draws the whole batch's noise)
    __device__ __forceinline__ float* out_f32() const { return static_cast<float*>(out); }
};

constexpr int kDisturbThreads = 256;
constexpr int kSmemHeaderFloats = 64;   // [0,32) reduction scratch, [32] cluster partial

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// [tv] gaussian_noise_image: clamp(x + (0.0 + n*sigma), 0, 1); no FMA contraction so the
// result is bit-identical to the eager mul / add / clamp sequence.
__device__ __forceinline__ float noisy(float x, float n, float sigma) {
    return clamp01(__fadd_rn(x, __fmul_rn(n, sigma)));
}

// Horizontal K-tap filter of 4 adjacent columns of one smem row.  `rowq` points at column j0 of the
// row.  EDGE = false: the quad is interior, so columns j0-4 .. j0+7 exist and are fetched with three
// 128-bit loads.  EDGE = true: first / last quad of a row, scalar loads with reflect indexing.
template <int K, bool EDGE>
__device__ __forceinline__ void hfilter4(const float* __restrict__ rowq, int j0, int W, const float (&taps)[K], float (&h)[4]) {
    constexpr int P = K / 2;
    float v[12];                                   // columns j0-4 .. j0+7
    if constexpr (!EDGE) {
        const float4 a = *reinterpret_cast<const float4*>(rowq - 4);
        const float4 m = *reinterpret_cast<const float4*>(rowq);
        const float4 z = *reinterpret_cast<const float4*>(rowq + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = m.x; v[5] = m.y; v[6] = m.z; v[7] = m.w;
        v[8] = z.x; v[9] = z.y; v[10] = z.z; v[11] = z.w;
    } else {
#pragma unroll
        for (int t = 4 - P; t < 8 + P; ++t) {
            int jj = j0 - 4 + t;
            if (jj < 0) jj = -jj;
            if (jj >= W) jj = 2 * (W - 1) - jj;
            v[t] = rowq[jj - j0];
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float acc = 0.0f;
#pragma unroll
        for (int u = 0; u < K; ++u) acc = fmaf(taps[u], v[4 + o + u - P], acc);
        h[o] = acc;
    }
}

// One blur task: a 4-column x [ra, rb) block of channel plane `base`.  The horizontally filtered rows
// live in a K-deep register ring whose slots are compile-time (row loop unrolled by K): the vertical
// pass is K FMAs per output, no register moves, no address arithmetic beyond two pointer bumps.
template <int K, bool EDGE>
__device__ __forceinline__ void blur_task(const float* __restrict__ base, float* __restrict__ outq, int W, int j0,
                                          int ra, int rb, int ir0, const float (&taps)[K], unsigned cutmask, int sh, int sh_end) {
    const float* rowq = base + ra * W + j0;        // tile row `ra` <-> first input row of output row ra
    float ring[K][4];
#pragma unroll
    for (int i = 0; i < K - 1; ++i) { hfilter4<K, EDGE>(rowq, j0, W, taps, ring[i]); rowq += W; }
    int ir = ir0 + ra;
    outq += static_cast<size_t>(ra) * W;
    for (int r = ra; r < rb; r += K) {
#pragma unroll
        for (int rr = 0; rr < K; ++rr) {
            if (r + rr < rb) {
                hfilter4<K, EDGE>(rowq, j0, W, taps, ring[(rr + K - 1) % K]);
                rowq += W;
                float o4[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float acc = 0.0f;
#pragma unroll
                    for (int u = 0; u < K; ++u) acc = fmaf(taps[u], ring[(rr + u) % K][o], acc);
                    o4[o] = acc;
                }
                if (cutmask && ir >= sh && ir < sh_end) {
#pragma unroll
                    for (int o = 0; o < 4; ++o) if (cutmask & (1u << o)) o4[o] = 0.0f;
                }
                st_stream_f4(outq, make_float4(o4[0], o4[1], o4[2], o4[3]));
                outq += W;
                ++ir;
            }
        }
    }
}


// =============================================================================================
// Fast path: x / noise contiguous fp32 NCHW, W % 4 == 0, k <= 7, fp32 NCHW out - every shape the
// reference's call sites and BASELINE.json's configs use.  Same phase structure as the general
// kernel below, built for instruction count (the general kernel is issue-bound at ~100 thread
// instructions per element):
//   * smem rows carry kPad reflected columns on either side (pitch W + 8), filled by a tiny pass
//     of their own, so the horizontal filter is three aligned 128-bit loads for EVERY quad - no
//     edge variant, no reflect arithmetic in the hot loop;
//   * halo rows are re-read from global memory (L2) and perturbed again instead of being copied
//     over DSMEM: the only cluster-wide dependency left is the gray mean (one barrier; the second
//     one, which keeps CTAs resident until their partial sum has been read, is split around the blur);
//   * clamps are the .sat modifier of the add (2 instructions per element for noise, 2 for contrast);
//   * thread -> (row, quad) positions advance incrementally (no division in any loop);
//   * the cutout is a warp-uniform row test plus a per-thread column mask.
// =============================================================================================
constexpr int kPad = 4;
constexpr int kFastMaxThreads = 384;     // x 2 CTAs / SM => at most 85 registers per thread

// ---- in-kernel noise (opt-in; the default path reads the tensor torch.randn_like drew) ----------------------
// Philox4x32-10 (Salmon et al., the generator behind torch's CUDA randn) on counter (quad_lo, quad_hi, off_lo, off_hi)
// and key (seed_lo, seed_hi); the four 32-bit outputs become the four N(0,1) draws of the quad by two Box-Muller
// transforms.  A draw is a function of (seed, offset, logical NCHW element index) only: independent of the stripe
// decomposition, of the memory layout of x and of the number of GPUs a batch is sharded over (global image index
// goes in through the base quad index).  oracle/philox.py restates it in numpy.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
// (a, b) -> two N(0,1): r = sqrt(-2 ln u), u = (a + 0.5) 2^-32 in (0, 1); theta = pi * int32(b) * 2^-31 in [-pi, pi)
__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
    const float u = fmaf(static_cast<float>(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float r = sqrtf(-1.3862943611198906f * __log2f(u));           // -2 ln 2 * log2 u
    const float th = static_cast<float>(static_cast<int>(b)) * 1.4629180792671596e-9f;   // pi * 2^-31
    return make_float2(r * __cosf(th), r * __sinf(th));
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long quad, unsigned long long seed, unsigned long long offset) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<unsigned>(quad), static_cast<unsigned>(quad >> 32),
                                             static_cast<unsigned>(offset), static_cast<unsigned>(offset >> 32)),
                                  make_uint2(static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32)));
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

__device__ __forceinline__ float sat_add(float a, float b) {      // clamp(a + b, 0, 1)
    float r;
    asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// Horizontal K-tap filter of one quad.  Blackwell's packed fp32 pipe (FFMA2: two independent
// round-to-nearest FMAs per instruction, bit-identical to two FFMAs) halves the instruction count.
template <int K>
__device__ __forceinline__ void hfilter4p(const float* __restrict__ rowq, const float2 (&t2)[K], float2 (&h)[2]) {
    constexpr int P = K / 2;
    const float4 a = *reinterpret_cast<const float4*>(rowq - 4);
    const float4 m = *reinterpret_cast<const float4*>(rowq);
    const float4 z = *reinterpret_cast<const float4*>(rowq + 4);
    const float v[12] = {a.x, a.y, a.z, a.w, m.x, m.y, m.z, m.w, z.x, z.y, z.z, z.w};
    // Output pair o reads the pairs (v[i], v[i+1]), i = 4 + 2o - P + u.  Those with even i sit in an aligned
    // register pair of the 128-bit loads and go through FFMA2; the odd ones would cost two MOVs to re-pair,
    // so they are two scalar FFMAs on the halves of the accumulator instead (same arithmetic, bit for bit).
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        float2 acc;
        if constexpr ((P & 1) == 0) {
            acc = __fmul2_rn(t2[0], make_float2(v[4 + 2 * o - P], v[5 + 2 * o - P]));
        } else {
            acc.x = t2[0].x * v[4 + 2 * o - P];
            acc.y = t2[0].x * v[5 + 2 * o - P];
        }
#pragma unroll
        for (int u = 1; u < K; ++u) {
            if (((u - P) & 1) == 0) {
                acc = __ffma2_rn(t2[u], make_float2(v[4 + 2 * o + u - P], v[5 + 2 * o + u - P]), acc);
            } else {
                acc.x = fmaf(t2[u].x, v[4 + 2 * o + u - P], acc.x);
                acc.y = fmaf(t2[u].x, v[5 + 2 * o + u - P], acc.y);
            }
        }
        h[o] = acc;
    }
}

// WT = compile-time image width (84 / 224: the reference's frame sizes; strides, quad counts and the
// index divisions become immediates) or 0 for a run-time width.
//
// One image = one cluster of S stripe-CTAs (S = 1: a plain CTA).  The only cluster-wide dependency is
// the per-image gray mean of the contrast stage: one barrier, with the halo-row loads between its
// ARRIVE and its WAIT.
// VAR 0: the plain kernel.  VAR 1: + out_scale.  VAR 2: + in-kernel Philox noise (and out_scale).  Separate instantiations so
// that the plain kernel - at the 85-register limit for k = 7 - carries none of it.
template <int K, int WT, bool XU8, bool NHWC, int VAR = 0>
__global__ void __launch_bounds__(kFastMaxThreads, 2)
disturb_fast_kernel(const __grid_constant__ DisturbParams p) {
    constexpr bool PHILOX = (VAR == 2), SCALED = (VAR >= 1);
    constexpr int P = K / 2;
    extern __shared__ __align__(16) float smem[];
    float* red = smem;
    float* partial = smem + 32;
    float* tile = smem + kSmemHeaderFloats;

    cg::cluster_group cluster = cg::this_cluster();
    const int S = p.S, R = p.R, C = p.C, H = p.H, W = WT ? WT : p.W;
    const int b = blockIdx.x >> p.log2S, s = blockIdx.x & (S - 1);
    const int r0 = min(s * R, H), r1 = min(r0 + R, H), rows = r1 - r0;
    const int WP = W + 2 * kPad;              // smem row pitch (floats)
    const int RS = R + 2 * P;                 // smem rows per channel: P halo rows above, R own rows, P below
    const int plane = RS * WP;                // smem floats per channel
    const int nq = W >> 2;
    auto div_nq = [&](int i) { return WT ? i / (WT >> 2 ? WT >> 2 : 1) : static_cast<int>(__umulhi(i, p.magic_nq)); };
    const int tid = threadIdx.x, nth = blockDim.x;
    const bool do_noise = (p.stages & CLIPPPO_STAGE_NOISE) != 0;
    const bool do_contrast = (p.stages & CLIPPPO_STAGE_CONTRAST) != 0;
    const bool exchange = do_contrast && S > 1;                     // the gray mean spans several CTAs
    const float sigma = p.sigma_n;
    const int n4 = rows * nq;
    const int HW = H * W;
    // PHILOX: logical NCHW quad index of (image b, channel 0, row 0, column 0)
    [[maybe_unused]] const unsigned long long quad0 =
        (static_cast<unsigned long long>(b) + static_cast<unsigned long long>(p.philox_first)) * static_cast<unsigned>(C * (HW >> 2));

    auto noisy4 = [&](float4 v, const float4& nz) {     // [tv] gaussian_noise_image: mul, add, clamp - no FMA contraction (bit-exact)
        if (do_noise) {
            v.x = sat_add(v.x, __fmul_rn(nz.x, sigma));
            v.y = sat_add(v.y, __fmul_rn(nz.y, sigma));
            v.z = sat_add(v.z, __fmul_rn(nz.z, sigma));
            v.w = sat_add(v.w, __fmul_rn(nz.w, sigma));
        }
        return v;
    };

    // ---- load: x (+ sigma * noise, clamp) -> tile; returns this thread's share of the gray sum over the OWN rows.
    // The stripe of a channel is contiguous in global memory and is walked linearly, one quad per thread
    // and step; quad index -> (row, quad-in-row) is a multiply-high by a reciprocal (an immediate when the
    // width is a template constant); global offsets are 32-bit from one per-image base pointer.
    // One quad position j (row, quad-in-row) is shared by the channels: the index arithmetic is paid once per
    // position, the channels differ by constants.  UJ positions x CT channels = 6 quads of x and 6 of noise
    // per thread and trip (12 independent 128-bit loads in flight), no serial tail.
    auto load_own_c = [&](auto ct_c, auto uj_c, int c_first) {
        constexpr int CT = decltype(ct_c)::value, UJ = decltype(uj_c)::value;
        const size_t x_off = static_cast<size_t>(b) * p.xs[0], n_off = static_cast<size_t>(b) * p.ns[0];   // batch strides (elements)
        constexpr bool xu8 = XU8;                        // uint8 frames (a template flag: as a run-time branch it cost the fp32 path 10 %): element offsets are byte offsets
        const float* xs = xu8 ? reinterpret_cast<const float*>(static_cast<const uint8_t*>(p.x) + x_off + c_first * HW + r0 * W)
                              : static_cast<const float*>(p.x) + x_off + c_first * HW + r0 * W;
        const float* ns = do_noise ? p.noise + n_off + c_first * HW + r0 * W : static_cast<const float*>(p.x);
        asm volatile("" : "+l"(xs), "+l"(ns));           // keep the two bases in registers (no rematerialisation)
        float* const tile0 = tile + c_first * plane + P * WP + kPad;      // first own row, first real column
        float gs[CT];
#pragma unroll
        for (int c_ = 0; c_ < CT; ++c_) gs[c_] = 0.0f;
        for (int j0 = tid; j0 < n4; j0 += nth * UJ) {
            float4 xv[UJ][CT], nv[UJ][CT];
#pragma unroll
            for (int u = 0; u < UJ; ++u) {
                const int j = j0 + u * nth;
                if (j < n4) {
#pragma unroll
                    for (int c_ = 0; c_ < CT; ++c_) {
                        if constexpr (xu8) xv[u][c_].x = __uint_as_float(ld_stream_u32(reinterpret_cast<const uint8_t*>(xs) + c_ * HW + 4 * j));
                        else xv[u][c_] = ld_stream_f4(xs + c_ * HW + 4 * j);
                        if constexpr (!PHILOX) { if (do_noise) nv[u][c_] = ld_stream_f4(ns + c_ * HW + 4 * j); }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UJ; ++u) {
                const int j = j0 + u * nth;
                if (j < n4) {
                    float* dst = tile0 + 4 * j + 2 * kPad * div_nq(j);    // row * WP + 4 * quad, with 4 j = row * W + 4 * quad
#pragma unroll
                    for (int c_ = 0; c_ < CT; ++c_) {
                        if constexpr (xu8) xv[u][c_] = u8x4_over_255(__float_as_uint(xv[u][c_].x));
                        if constexpr (PHILOX) {
                            if (do_noise) nv[u][c_] = philox_normal4(quad0 + static_cast<unsigned>((c_first + c_) * (HW >> 2) + (r0 * W >> 2) + j),
                                                                     p.philox_seed, p.philox_offset);
                        }
                        const float4 v = noisy4(xv[u][c_], nv[u][c_]);
                        gs[c_] += (v.x + v.y) + (v.z + v.w);
                        *reinterpret_cast<float4*>(dst + c_ * plane) = v;
                    }
                }
            }
        }
        if constexpr (CT == 3) return fmaf(0.114f, gs[2], fmaf(0.587f, gs[1], 0.2989f * gs[0]));
        else return gs[0];
    };
    // NHWC-dense input: the stripe is still one contiguous run of rows * W * 3 floats; position j (row, quad) owns the 12
    // floats at 12 j - three 128-bit loads of x and three of noise - which are the quad's four pixels of all three channels.
    auto deinterleave = [](const float4& a, const float4& b, const float4& c, float4 (&o)[3]) {
        o[0] = make_float4(a.x, a.w, b.z, c.y);
        o[1] = make_float4(a.y, b.x, b.w, c.z);
        o[2] = make_float4(a.z, b.y, c.x, c.w);
    };
    auto load_own_nhwc = [&]() {
        constexpr int UJ = 2;
        const float* xs = static_cast<const float*>(p.x) + static_cast<size_t>(b) * p.xs[0] + 3 * r0 * W;
        const float* ns = do_noise ? p.noise + static_cast<size_t>(b) * p.ns[0] + 3 * r0 * W : xs;
        asm volatile("" : "+l"(xs), "+l"(ns));
        float* const tile0 = tile + P * WP + kPad;
        float gs[3] = {0.0f, 0.0f, 0.0f};
        for (int j0 = tid; j0 < n4; j0 += nth * UJ) {
            float4 xv[UJ][3], nv[UJ][3];
#pragma unroll
            for (int u = 0; u < UJ; ++u) {
                const int j = j0 + u * nth;
                if (j < n4) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        xv[u][t] = ld_stream_f4(xs + 12 * j + 4 * t);
                        if (do_noise) nv[u][t] = ld_stream_f4(ns + 12 * j + 4 * t);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UJ; ++u) {
                const int j = j0 + u * nth;
                if (j < n4) {
                    float* dst = tile0 + 4 * j + 2 * kPad * div_nq(j);
                    float4 xc[3], nc[3];
                    deinterleave(xv[u][0], xv[u][1], xv[u][2], xc);
                    if (do_noise) deinterleave(nv[u][0], nv[u][1], nv[u][2], nc);
#pragma unroll
                    for (int c_ = 0; c_ < 3; ++c_) {
                        const float4 v = noisy4(xc[c_], nc[c_]);
                        gs[c_] += (v.x + v.y) + (v.z + v.w);
                        *reinterpret_cast<float4*>(dst + c_ * plane) = v;
                    }
                }
            }
        }
        return fmaf(0.114f, gs[2], fmaf(0.587f, gs[1], 0.2989f * gs[0]));
    };
    auto load_own = [&]() {
        if constexpr (NHWC) return load_own_nhwc();
        if (C == 3) return load_own_c(std::integral_constant<int, 3>{}, std::integral_constant<int, 2>{}, 0);
        float g = 0.0f;                                   // C == 1 (or, without a contrast stage, any channel count)
        for (int c_ = 0; c_ < C; ++c_) g += load_own_c(std::integral_constant<int, 1>{}, std::integral_constant<int, 6>{}, c_);
        return g;
    };
    // Halo rows: the P image rows above and below the stripe (reflected at the image border) are
    // fetched and perturbed again by this CTA instead of being copied from the neighbour's shared
    // memory - 2P/R more (mostly L2-resident) reads, but no cluster barrier before the blur.  They are
    // not part of the gray sum, so they load between the ARRIVE and the WAIT of the mean barrier.
    auto load_halo = [&]() {
        const size_t x_off = static_cast<size_t>(b) * p.xs[0], n_off = static_cast<size_t>(b) * p.ns[0];   // batch strides (elements)
        constexpr bool xu8 = XU8;
        const float* __restrict__ xi = static_cast<const float*>(p.x) + (xu8 ? 0 : x_off);
        const uint8_t* __restrict__ xb = static_cast<const uint8_t*>(p.x) + x_off;
        const float* __restrict__ ni = do_noise ? p.noise + n_off : static_cast<const float*>(p.x);
        if constexpr (K > 1 && NHWC) {
            const int nh = rows > 0 ? 2 * P * nq : 0;               // (halo row, quad) positions; three channels each
            for (int i = tid; i < nh; i += nth) {
                const int hr = div_nq(i), quad = i - hr * nq;
                const int lr = hr < P ? hr : rows + hr;
                int ir = r0 - P + lr;
                if (ir < 0) ir = -ir;
                if (ir >= H) ir = 2 * (H - 1) - ir;
                const int goff = 3 * (ir * W + 4 * quad);
                float4 xv[3], nv[3], xc[3], nc[3];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    xv[t] = ld_stream_f4(xi + goff + 4 * t);
                    if (do_noise) nv[t] = ld_stream_f4(ni + goff + 4 * t);
                }
                deinterleave(xv[0], xv[1], xv[2], xc);
                if (do_noise) deinterleave(nv[0], nv[1], nv[2], nc);
#pragma unroll
                for (int c_ = 0; c_ < 3; ++c_)
                    *reinterpret_cast<float4*>(tile + c_ * plane + lr * WP + kPad + 4 * quad) = noisy4(xc[c_], nc[c_]);
            }
        } else if constexpr (K > 1) {
            const int nh4 = rows > 0 ? C * 2 * P * nq : 0;
            // every load of UH positions is in flight before the first one is used: one round trip to L2 / HBM per UH * nth quads
            constexpr int UH = 3;
            for (int i0 = tid; i0 < nh4; i0 += nth * UH) {
                float4 xv[UH], nv[UH];
                int soff[UH], gq[UH];
#pragma unroll
                for (int u = 0; u < UH; ++u) {
                    const int i = i0 + u * nth;
                    if (i < nh4) {
                        const int hr_c = div_nq(i), quad = i - hr_c * nq;
                        const int c_ = hr_c / (2 * P), hr = hr_c - c_ * (2 * P);
                        const int lr = hr < P ? hr : rows + hr;
                        int ir = r0 - P + lr;
                        if (ir < 0) ir = -ir;
                        if (ir >= H) ir = 2 * (H - 1) - ir;
                        const int goff = c_ * HW + ir * W + 4 * quad;
                        soff[u] = c_ * plane + lr * WP + kPad + 4 * quad;
                        gq[u] = goff >> 2;
                        if constexpr (xu8) xv[u].x = __uint_as_float(ld_stream_u32(xb + goff)); else xv[u] = ld_stream_f4(xi + goff);
                        if constexpr (!PHILOX) { if (do_noise) nv[u] = ld_stream_f4(ni + goff); }
                    }
                }
#pragma unroll
                for (int u = 0; u < UH; ++u) {
                    const int i = i0 + u * nth;
                    if (i < nh4) {
                        if constexpr (xu8) xv[u] = u8x4_over_255(__float_as_uint(xv[u].x));
                        if constexpr (PHILOX) { if (do_noise) nv[u] = philox_normal4(quad0 + static_cast<unsigned>(gq[u]), p.philox_seed, p.philox_offset); }
                        *reinterpret_cast<float4*>(tile + soff[u]) = noisy4(xv[u], nv[u]);
                    }
                }
            }
        }
    };

    // ---- contrast blend in place over own + halo rows: [tv] _blend: c*x + (1-c)*mean, clamp ----
    auto contrast_image = [&](float tot) {
        const float m = tot / static_cast<float>(H * W);
        const float cm = __fmul_rn(p.omc, m);
        const float cf = p.c;
        const int n43 = (rows > 0 ? rows + 2 * P : 0) * nq;
        for (int j = tid; j < n43; j += nth) {
            float* q = tile + kPad + 4 * j + 2 * kPad * div_nq(j);
#pragma unroll
            for (int c_ = 0; c_ < 3; ++c_) {
                if (c_ < C) {
                    float4* q4 = reinterpret_cast<float4*>(q + c_ * plane);
                    float4 v = *q4;
                    v.x = sat_add(__fmul_rn(cf, v.x), cm);
                    v.y = sat_add(__fmul_rn(cf, v.y), cm);
                    v.z = sat_add(__fmul_rn(cf, v.z), cm);
                    v.w = sat_add(__fmul_rn(cf, v.w), cm);
                    *q4 = v;
                }
            }
        }
    };

    // ---- reflected pad columns of every smem row (own + halo) ----
    auto pad_image = [&]() {
        if constexpr (K > 1) {
            for (int i = tid; i < C * RS * 2; i += nth) {
                const int side = i & 1, cr = i >> 1;
                const int c_ = __umulhi(cr, p.magic_rs), row = cr - c_ * RS;
                float* rowp = tile + c_ * plane + row * WP + kPad;
                if (side == 0) { rowp[-1] = rowp[1]; rowp[-2] = rowp[2]; rowp[-3] = rowp[3]; }            // x[-j] = x[j]
                else { rowp[W] = rowp[W - 2]; rowp[W + 1] = rowp[W - 3]; rowp[W + 2] = rowp[W - 4]; }     // x[W-1+j] = x[W-1-j]
            }
        }
    };

    // ---- separable blur (horizontal from smem, vertical in a register ring), cutout, store ----
    float2 t2[K];
#pragma unroll
    for (int u = 0; u < K; ++u) t2[u] = make_float2(p.taps[u], p.taps[u]);
    const int nsplit = p.nsplit;
    const int rps = (rows + nsplit - 1) / nsplit;
    const int ntasks = C * nsplit * nq;
    const bool do_cut = (p.stages & CLIPPPO_STAGE_CUTOUT) != 0;
    const int sw_end = p.sw + p.pw;
    [[maybe_unused]] const float2 sc2 = make_float2(p.out_scale, p.out_scale);
    auto blur_image = [&]() {
        for (int task = tid; task < ntasks; task += nth) {
            const int rest = div_nq(task), q = task - rest * nq;
            const int c_ = nsplit == 1 ? rest : __umulhi(rest, p.magic_nsplit), sp = rest - c_ * nsplit;
            const int ra = sp * rps, rb = min(ra + rps, rows);
            if (ra >= rb) continue;
            const int j0 = q * 4;
            float2 keep01 = make_float2(1.0f, 1.0f), keep23 = keep01;
            bool any_cut = false;
            if (do_cut) {
                if (j0 >= p.sw && j0 < sw_end) { keep01.x = 0.0f; any_cut = true; }
                if (j0 + 1 >= p.sw && j0 + 1 < sw_end) { keep01.y = 0.0f; any_cut = true; }
                if (j0 + 2 >= p.sw && j0 + 2 < sw_end) { keep23.x = 0.0f; any_cut = true; }
                if (j0 + 3 >= p.sw && j0 + 3 < sw_end) { keep23.y = 0.0f; any_cut = true; }
            }
            // image rows [cut0, cut1) of this task are inside the window (empty range if the quad is outside)
            const int cut0 = any_cut ? p.sh : 0x7fffffff, cut1 = any_cut ? p.sh + p.ph : 0;
            const float* rowq = tile + c_ * plane + ra * WP + kPad + j0;     // tile row ra = first input row of output row ra
            float* outq = p.out_f32() + ((static_cast<size_t>(b) * C + c_) * H + r0 + ra) * W + j0;
            float2 ring[K][2];                       // horizontally filtered rows, slots are compile-time
#pragma unroll
            for (int i = 0; i < K - 1; ++i) hfilter4p<K>(rowq + i * WP, t2, ring[i]);
            rowq += (K - 1) * WP;
            int ir = r0 + ra;
            // one output row: filter the newest input row into ring slot (rr + K - 1) % K, combine the K slots
            auto out_row = [&](auto rr_c) {
                constexpr int rr = decltype(rr_c)::value;
                hfilter4p<K>(rowq + rr * WP, t2, ring[(rr + K - 1) % K]);
                float2 o01 = __fmul2_rn(t2[0], ring[rr % K][0]), o23 = __fmul2_rn(t2[0], ring[rr % K][1]);
#pragma unroll
                for (int u = 1; u < K; ++u) {
                    o01 = __ffma2_rn(t2[u], ring[(rr + u) % K][0], o01);
                    o23 = __ffma2_rn(t2[u], ring[(rr + u) % K][1], o23);
                }
                const int irr = ir + rr;
                if (irr >= cut0 && irr < cut1) { o01 = __fmul2_rn(o01, keep01); o23 = __fmul2_rn(o23, keep23); }
                if constexpr (SCALED) {             // one fp32 multiply: == `out * scale` of torch, bit for bit
                    o01 = __fmul2_rn(o01, sc2); o23 = __fmul2_rn(o23, sc2);
                }
                st_stream_f4(outq + rr * W, make_float4(o01.x, o01.y, o23.x, o23.y));
            };
            int r = ra;
            for (; r + K <= rb; r += K) {            // full K-row blocks: the ring rotates back to slot 0, no row tests
                static_for<K>(out_row);
                rowq += K * WP; outq += K * W; ir += K;
            }
            if constexpr (K > 1) {                   // tail of < K rows
                const int left = rb - r;
                static_for<K - 1>([&](auto rr_c) { if ((decltype(rr_c)::value) < left) out_row(rr_c); });
            }
        }
    };

    // ---- the image ----
    const float g = load_own();
    float tot = 0.0f;
    if (do_contrast) {
        tot = block_sum(g, red);
        if (exchange) {
            // release of ONE shared-memory word: the writer fences, every thread arrives relaxed (an arrive.release is a
            // MEMBAR.ALL.GPU in each of the CTA's threads)
            if (tid == 0) { *partial = tot; asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        }
    }
    load_halo();                                     // not part of the sum: rides between ARRIVE and WAIT
    if (do_contrast) {
        if (exchange) {
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            tot = 0.0f;
            for (int q = 0; q < S; ++q) tot += *cluster.map_shared_rank(partial, q);
            // "my reads of the peers' sums are done": relaxed - there is nothing to release, and an arrive.release is a MEMBAR in
            // every thread.  The arrive must not overtake the loads, though: a (never taken) branch on their sum is a scoreboard
            // wait for all of them, and instructions issue in order.
            if (__float_as_uint(tot) == 0x7fc0deadu) asm volatile("nanosleep.u32 1;" ::: "memory");
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        }
        __syncthreads();
        contrast_image(tot);
    }
    __syncthreads();
    pad_image();
    __syncthreads();
    blur_image();
    // nobody exits while a peer may still read its partial sum: the second barrier phase, waited for here
    if (exchange) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int K>
__global__ void __launch_bounds__(kDisturbThreads)
disturb_kernel(const __grid_constant__ DisturbParams p) {
    constexpr int P = K / 2;
    extern __shared__ __align__(16) float smem[];
    float* red = smem;
    float* partial = smem + 32;
    float* tile = smem + kSmemHeaderFloats;

    cg::cluster_group cluster = cg::this_cluster();
    const int S = p.S, R = p.R, C = p.C, H = p.H, W = p.W;
    const int b = blockIdx.x / S, s = blockIdx.x % S;
    const int r0 = min(s * R, H), r1 = min(r0 + R, H), rows = r1 - r0;
    const int RS = R + 2 * P;                 // smem rows per channel
    const int plane = RS * W;                 // smem floats per channel
    const int tid = threadIdx.x, nth = blockDim.x;
    const bool do_noise = (p.stages & CLIPPPO_STAGE_NOISE) != 0;
    const bool do_contrast = (p.stages & CLIPPPO_STAGE_CONTRAST) != 0;
    const bool do_cut = (p.stages & CLIPPPO_STAGE_CUTOUT) != 0;
    const float sigma = p.sigma_n;
    const bool vec4 = (W & 3) == 0;           // smem rows are 16-byte aligned

    // ---- phase 1: stripe -> smem -----------------------------------------------------------
    float gsum = 0.0f;
    if (p.fast && p.batch_contig) {
        const int n4 = (rows * W) >> 2;       // float4 per channel of this stripe (contiguous in global)
        constexpr int UNR = 4;
        for (int c_ = 0; c_ < C; ++c_) {
            const size_t goff = (static_cast<size_t>(b) * C + c_) * H * W + static_cast<size_t>(r0) * W;
            const float* xs = static_cast<const float*>(p.x) + goff;
            const float* ns = do_noise ? p.noise + goff : nullptr;
            float4* dst = reinterpret_cast<float4*>(tile + c_ * plane + P * W);
            const float wc = (C == 3) ? (c_ == 0 ? 0.2989f : (c_ == 1 ? 0.587f : 0.114f)) : 1.0f;
            float csum = 0.0f;
            for (int i0 = tid; i0 < n4; i0 += nth * UNR) {
                float4 xv[UNR], nv[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int i = i0 + u * nth;
                    if (i < n4) {
                        xv[u] = ld_stream_f4(xs + 4 * i);
                        if (do_noise) nv[u] = ld_stream_f4(ns + 4 * i);
                    }
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int i = i0 + u * nth;
                    if (i < n4) {
                        float4 v = xv[u];
                        if (do_noise) {
                            v.x = noisy(v.x, nv[u].x, sigma);
                            v.y = noisy(v.y, nv[u].y, sigma);
                            v.z = noisy(v.z, nv[u].z, sigma);
                            v.w = noisy(v.w, nv[u].w, sigma);
                        }
                        csum += (v.x + v.y) + (v.z + v.w);
                        dst[i] = v;
                    }
                }
            }
            gsum = fmaf(wc, csum, gsum);
        }
    } else {
        const int n = rows * W;
        const int total = C * n;
        const bool chan_fastest = p.xs[1] < p.xs[3];     // NHWC-like memory order
        for (int i = tid; i < total; i += nth) {
            int c_, r, j;
            if (chan_fastest) { c_ = i % C; const int q = i / C; j = q % W; r = q / W; }
            else              { j = i % W;  const int q = i / W; r = q % rows; c_ = q / rows; }
            const long long xo = (long long)b * p.xs[0] + c_ * p.xs[1] + (long long)(r0 + r) * p.xs[2] + j * p.xs[3];
            float v;
            if (p.io_mode == 0)      v = static_cast<const float*>(p.x)[xo];
            else if (p.io_mode == 1) v = static_cast<float>(static_cast<const uint8_t*>(p.x)[xo]) / 255.0f;
            else                     v = static_cast<const float*>(p.x)[xo] / 255.0f;
            if (do_noise) {
                const long long no = (long long)b * p.ns[0] + c_ * p.ns[1] + (long long)(r0 + r) * p.ns[2] + j * p.ns[3];
                v = noisy(v, p.noise[no], sigma);
            }
            const float wc = (C == 3) ? (c_ == 0 ? 0.2989f : (c_ == 1 ? 0.587f : 0.114f)) : 1.0f;
            gsum += wc * v;
            tile[c_ * plane + (P + r) * W + j] = v;
        }
    }

    // ---- phase 2 + 3: per-image gray mean, contrast blend in place -------------------------
    if (do_contrast) {
        float tot = block_sum(gsum, red);
        if (S > 1) {
            if (tid == 0) *partial = tot;
            cluster.sync();
            tot = 0.0f;
            for (int q = 0; q < S; ++q) tot += *cluster.map_shared_rank(partial, q);
        }
        const float m = tot / static_cast<float>(H * W);
        const float cm = __fmul_rn(p.omc, m);
        const float cf = p.c;
        __syncthreads();
        if (vec4) {
            const int n4 = (rows * W) >> 2;
            for (int c_ = 0; c_ < C; ++c_) {
                float4* q4 = reinterpret_cast<float4*>(tile + c_ * plane + P * W);
                for (int i = tid; i < n4; i += nth) {
                    float4 v = q4[i];
                    v.x = clamp01(__fadd_rn(__fmul_rn(cf, v.x), cm));
                    v.y = clamp01(__fadd_rn(__fmul_rn(cf, v.y), cm));
                    v.z = clamp01(__fadd_rn(__fmul_rn(cf, v.z), cm));
                    v.w = clamp01(__fadd_rn(__fmul_rn(cf, v.w), cm));
                    q4[i] = v;
                }
            }
        } else {
            const int n = rows * W;
            for (int c_ = 0; c_ < C; ++c_) {
                float* q = tile + c_ * plane + P * W;
                for (int i = tid; i < n; i += nth) q[i] = clamp01(__fadd_rn(__fmul_rn(cf, q[i]), cm));
            }
        }
    }

    // ---- phase 4: halo rows over DSMEM -----------------------------------------------------
    if (K > 1) {
        if (S > 1) cluster.sync(); else __syncthreads();
        if (rows > 0) {
            // 2P halo rows per channel; one WARP per (channel, halo row), columns by lane - the
            // row bookkeeping (reflect, owner CTA) is paid once per row per warp, not per thread block
            const int hwarp = tid >> 5, hlane = tid & 31, hnw = nth >> 5;
            for (int hr_c = hwarp; hr_c < C * 2 * P; hr_c += hnw) {
                const int c_ = hr_c / (2 * P), hr = hr_c - c_ * (2 * P);
                const int lr = hr < P ? hr : rows + hr;              // below the last own row
                int ir = r0 - P + lr;                                // image row before reflection
                if (ir < 0) ir = -ir;
                if (ir >= H) ir = 2 * (H - 1) - ir;
                const int owner = ir / R;
                const int olr = ir - owner * R + P;
                const float* src = ((owner == s) ? tile : cluster.map_shared_rank(tile, owner)) + c_ * plane + olr * W;
                float* dstrow = tile + c_ * plane + lr * W;
                if (vec4) {
                    for (int j = hlane; j < (W >> 2); j += 32)
                        reinterpret_cast<float4*>(dstrow)[j] = reinterpret_cast<const float4*>(src)[j];
                } else {
                    for (int j = hlane; j < W; j += 32) dstrow[j] = src[j];
                }
            }
        }
        if (S > 1) cluster.sync(); else __syncthreads();         // also: nobody exits while peers read
    } else {
        __syncthreads();
    }

    // ---- phase 5: separable blur, cutout, store --------------------------------------------
    float taps[K];
#pragma unroll
    for (int u = 0; u < K; ++u) taps[u] = p.taps[u];
    if (p.io_mode == 0 && vec4 && P <= 4) {
        // Fast path: 4-column x rps-row tasks, consecutive threads on consecutive column quads.
        const int nq = W >> 2, nsplit = p.nsplit;
        const int rps = (rows + nsplit - 1) / nsplit;
        const int ntasks = C * nsplit * nq;
        const int sh_end = p.sh + p.ph, sw_end = p.sw + p.pw;
        // Task order: all interior quads first, the 2 edge quads of every (channel, split) last, so a
        // warp runs either the 128-bit interior code or the reflect-indexed edge code, never both.
        const int nqi = nq > 2 ? nq - 2 : 0, nint = C * nsplit * nqi;
        for (int task = tid; task < ntasks; task += nth) {
            int q, rest;
            if (task < nint) { rest = task / nqi; q = 1 + (task - rest * nqi); }
            else if (nq >= 2) { const int e = task - nint; rest = e >> 1; q = (e & 1) ? nq - 1 : 0; }
            else { rest = task - nint; q = 0; }
            const int sp = rest % nsplit, c_ = rest / nsplit;
            const int ra = sp * rps, rb = min(ra + rps, rows);
            if (ra >= rb) continue;
            const int j0 = q * 4;
            unsigned cutmask = 0;
            if (do_cut) {
#pragma unroll
                for (int o = 0; o < 4; ++o) if (j0 + o >= p.sw && j0 + o < sw_end) cutmask |= 1u << o;
            }
            const float* base = tile + c_ * plane;
            float* outq = static_cast<float*>(p.out) + ((static_cast<size_t>(b) * C + c_) * H + r0) * W + j0;
            if (q > 0 && q < nq - 1) blur_task<K, false>(base, outq, W, j0, ra, rb, r0, taps, cutmask, p.sh, sh_end);
            else                     blur_task<K, true>(base, outq, W, j0, ra, rb, r0, taps, cutmask, p.sh, sh_end);
        }
        return;
    }
    // Generic path (odd widths, wide kernels, uint8 NHWC output): one column per thread.
    for (int task = tid; task < C * W; task += nth) {
        const int c_ = task / W, j = task - c_ * W;
        int jidx[K];
#pragma unroll
        for (int u = 0; u < K; ++u) {
            int jj = j + u - P;
            if (jj < 0) jj = -jj;
            if (jj >= W) jj = 2 * (W - 1) - jj;
            jidx[u] = jj;
        }
        const float* base = tile + c_ * plane;
        float ring[K];
#pragma unroll
        for (int lr = 0; lr < K - 1; ++lr) {
            float h = 0.0f;
#pragma unroll
            for (int u = 0; u < K; ++u) h = fmaf(taps[u], base[lr * W + jidx[u]], h);
            ring[lr] = h;
        }
        const bool col_cut = do_cut && j >= p.sw && j < p.sw + p.pw;
        for (int r = 0; r < rows; ++r) {
            float h = 0.0f;
#pragma unroll
            for (int u = 0; u < K; ++u) h = fmaf(taps[u], base[(r + K - 1) * W + jidx[u]], h);
            ring[K - 1] = h;
            float v = 0.0f;
#pragma unroll
            for (int u = 0; u < K; ++u) v = fmaf(taps[u], ring[u], v);
#pragma unroll
            for (int u = 0; u < K - 1; ++u) ring[u] = ring[u + 1];
            const int ir = r0 + r;
            if (col_cut && ir >= p.sh && ir < p.sh + p.ph) v = 0.0f;
            if (p.io_mode == 0) {
                static_cast<float*>(p.out)[((static_cast<size_t>(b) * C + c_) * H + ir) * W + j] = v;
            } else {
                static_cast<uint8_t*>(p.out)[((static_cast<size_t>(b) * H + ir) * W + j) * C + c_] =
                    static_cast<uint8_t>(__fmul_rn(v, 255.0f));
            }
        }
    }
}

template <int K>
static int launch_disturb(const DisturbParams& p, size_t smem, cudaStream_t stream) {
    static DeviceOnce configured;     // opt in to > 48 KB dynamic smem once per instantiation
    if (configured.first_use()) {
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(disturb_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(disturb_kernel<K>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(p.B) * p.S);
    cfg.blockDim = dim3(kDisturbThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CLIPPPO_CUDA_TRY(cudaLaunchKernelEx(&cfg, disturb_kernel<K>, p));
    prof_count_launch();
    return CLIPPPO_OK;
}

template <int K, int WT, bool XU8, bool NHWC, int VAR = 0>
static int launch_disturb_fast_x(const DisturbParams& p, size_t smem, cudaStream_t stream) {
    static DeviceOnce configured;
    if (configured.first_use()) {
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(disturb_fast_kernel<K, WT, XU8, NHWC, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CLIPPPO_CUDA_TRY(cudaFuncSetAttribute(disturb_fast_kernel<K, WT, XU8, NHWC, VAR>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(p.B) * p.S);
    cfg.blockDim = dim3(p.nthreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    // without the contrast stage the stripes of an image are independent: plain CTAs, no gang scheduling
    cfg.numAttrs = ((p.stages & CLIPPPO_STAGE_CONTRAST) && p.S > 1) ? 1 : 0;
    CLIPPPO_CUDA_TRY(cudaLaunchKernelEx(&cfg, disturb_fast_kernel<K, WT, XU8, NHWC, VAR>, p));
    prof_count_launch();
    return CLIPPPO_OK;
}

template <int K, int WT>
static int launch_disturb_fast_w(const DisturbParams& p, size_t smem, cudaStream_t stream) {
    if (p.philox || p.out_scale != 1.0f) {
        // the two additive features: contiguous NCHW images (fp32 or uint8) of the reference's frame widths; Philox noise for
        // the blurring severities.  Anything else: UNSUPPORTED, the caller falls back to torch noise / a separate multiply.
        if constexpr (WT != 0) {
            if (p.nhwc) return CLIPPPO_ERR_UNSUPPORTED;
            if (p.philox) {
                if constexpr (K > 1) {
                    return p.x_u8 ? launch_disturb_fast_x<K, WT, true, false, 2>(p, smem, stream)
                                  : launch_disturb_fast_x<K, WT, false, false, 2>(p, smem, stream);
                } else {
                    return CLIPPPO_ERR_UNSUPPORTED;
                }
            }
            return p.x_u8 ? launch_disturb_fast_x<K, WT, true, false, 1>(p, smem, stream)
                          : launch_disturb_fast_x<K, WT, false, false, 1>(p, smem, stream);
        } else {
            return CLIPPPO_ERR_UNSUPPORTED;
        }
    }
    if (p.nhwc) return launch_disturb_fast_x<K, WT, false, true>(p, smem, stream);
    return p.x_u8 ? launch_disturb_fast_x<K, WT, true, false>(p, smem, stream) : launch_disturb_fast_x<K, WT, false, false>(p, smem, stream);
}

template <int K>
static int launch_disturb_fast(const DisturbParams& p, size_t smem, cudaStream_t stream) {
    if (p.W == 224) return launch_disturb_fast_w<K, 224>(p, smem, stream);
    if (p.W == 84) return launch_disturb_fast_w<K, 84>(p, smem, stream);
    return launch_disturb_fast_w<K, 0>(p, smem, stream);
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Env-step batches (E = 8 ... 64 frames of 84 x 84 per call) leave most SMs without a CTA when an image is one
// CTA, and the call's duration is then one CTA's latency (52 us for 64 NHWC frames): cut such images into
// stripes (a cluster) until the grid covers the SMs, as long as a stripe keeps >= max(8, 4P) rows (23 us).
// Only images that would otherwise be a single CTA are widened, so large frames keep one stripe layout - and
// with it a bitwise batch-size-independent gray mean - at every batch size; for small frames the mean's
// summation order (last bit) depends on whether the batch is below 148 images.  CLIPPPO_DISTURB_WIDEN=0: off.
static int widen_for_small_batches(int S, int B, int H, int P, int max_cluster, int target_ctas) {
    static const int enabled = env_int("CLIPPPO_DISTURB_WIDEN", 1);
    if (!enabled || S != 1) return S;
    const int min_rows = 4 * P > 8 ? 4 * P : 8;
    while (static_cast<long long>(B) * S < target_ctas && 2 * S <= max_cluster && (H + 2 * S - 1) / (2 * S) >= min_rows) S *= 2;
    return S;
}

static int run_disturb(DisturbParams& p, const float* k1d_host, int k, cudaStream_t stream, int max_cluster = 0) {
    static const int env_cl = env_int("CLIPPPO_DISTURB_MAXCL", 16);
    static const int env_budget = env_int("CLIPPPO_DISTURB_SMEM_KB", 113);
    if (max_cluster == 0) max_cluster = env_cl;
    if (p.B <= 0 || p.C <= 0 || p.H <= 0 || p.W <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (!p.x || !p.out) return CLIPPPO_ERR_NULL;
    if ((p.stages & CLIPPPO_STAGE_NOISE) && !p.noise && !p.philox) return CLIPPPO_ERR_NULL;
    if (p.out_scale == 0.0f) p.out_scale = 1.0f;
    if ((p.stages & CLIPPPO_STAGE_CONTRAST) && p.C != 1 && p.C != 3) return CLIPPPO_ERR_BAD_CHANNELS;
    int K = 1;
    if (p.stages & CLIPPPO_STAGE_BLUR) {
        if (!k1d_host) return CLIPPPO_ERR_NULL;
        if (k < 1 || k > CLIPPPO_MAX_BLUR_TAPS || (k & 1) == 0) return CLIPPPO_ERR_BAD_SHAPE;
        if (k / 2 >= p.H || k / 2 >= p.W) return CLIPPPO_ERR_BAD_PAD;
        K = k;
        for (int i = 0; i < k; ++i) p.taps[i] = k1d_host[i];
    } else {
        p.taps[0] = 1.0f;
    }
    if (p.stages & CLIPPPO_STAGE_CUTOUT) {
        if (p.sh < 0 || p.sw < 0 || p.ph < 0 || p.pw < 0) return CLIPPPO_ERR_BAD_SHAPE;
    }
    const int P = K / 2;
    if (p.x_u8 && !(p.fast && p.io_mode == 0 && K <= 7 && p.W >= 8)) return CLIPPPO_ERR_UNSUPPORTED;   // uint8 frames: fast kernel only
    if (p.fast && p.io_mode == 0 && K <= 7 && p.W >= 8) {
        // ---- fast path (disturb_fast_kernel): padded smem rows ----
        static const int env_nsplit = env_int("CLIPPPO_DISTURB_NSPLIT", 0);
        auto tile_bytes = [&](int S) {
            const int R = (p.H + S - 1) / S;
            return (size_t)p.C * (R + 2 * P) * (p.W + 2 * kPad) * sizeof(float);
        };
        const size_t hdr = kSmemHeaderFloats * sizeof(float);
        // Stripes per image: the smallest cluster whose stripe (+ halo rows) fits half an SM (113 KB: two
        // CTAs of up to 384 threads per SM), else a whole SM.  Measured (profiles/r01_disturb_experiments.txt):
        // 2 x 352 threads beat 4 x 192 at 224x224x3 (63 % vs 58 % of the HBM peak), and an 84x84x3 frame
        // that fits one CTA (no cluster, no barrier) reaches 74 % instead of 65 % as a 2-CTA cluster.
        int S = 0;
        {
            const size_t budgets[3] = {static_cast<size_t>(env_budget) * 1024, 113 * 1024, 227 * 1024};
            for (int bi = 0; bi < 3 && !S; ++bi)
                for (int cand = 1; cand <= max_cluster; cand *= 2)
                    if (tile_bytes(cand) + hdr <= budgets[bi]) { S = cand; break; }
        }
        if (!S) return CLIPPPO_ERR_UNSUPPORTED;
        S = widen_for_small_batches(S, p.B, p.H, P, max_cluster, kNumSMs);
        // k = 7 on a small RGB frame that one CTA could hold (84 x 84 x 3 SEVERE): the 7-tap filter phase is long and the kernel
        // needs 80 registers, so one CTA per frame leaves an SM with two CTAs of 8 - 10 warps.  Two stripes per frame with row
        // splits for 256 threads (two load trips instead of three or four, three CTAs per SM) reach 82 % of the HBM copy peak
        // where one CTA per frame reaches 67 % (profiles/r02_disturb_smem_sweep84.txt) - affordable since the cluster barrier
        // lost its memory barriers.  k <= 5 and C = 1 are faster with the whole frame in one CTA (83 - 87 %).
        static const int env_two = env_int("CLIPPPO_DISTURB_TWO_STRIPES", 1);
        const bool two_stripes = env_two && K == 7 && p.C == 3 && S == 1 && max_cluster >= 2 && env_budget == 113 && p.H >= 8 * P &&
                                 tile_bytes(2) + hdr <= 75 * 1024;
        if (two_stripes) S = 2;
        p.S = S;
        p.R = (p.H + S - 1) / S;
        // Blur tasks = C x nsplit x (W/4) column quads, one per thread.  A split costs 2P extra
        // horizontally filtered rows per task; it buys warps (latency hiding) when a stripe has few quads.
        const int nq = p.W / 4;
        int nsplit = env_nsplit;
        if (nsplit <= 0) {
            nsplit = 1;
            while (p.C * nq * nsplit < 160 && nsplit < p.R) ++nsplit;                     // at least ~5 warps of tasks
            // ~10 warps while a split keeps >= 12 rows (>= 6P beyond the first cut: every split pays 2P warm-up rows).
            // Measured at 84x84x3 (tools/disturb_nsplit_84.py): 5 splits 84-87 % of the HBM peak at k = 3 / 5 where 6
            // gave 81-84 %; k = 7: 4 splits 67 %, 6 splits 62-64 %.
            auto min_rows = [&](int n) { return n >= 2 && 6 * P > 12 ? 6 * P : 12; };
            while (p.C * nq * nsplit < 300 && p.R / (nsplit + 1) >= min_rows(nsplit)) ++nsplit;
            if (two_stripes) nsplit = (256 + p.C * nq / 2) / (p.C * nq) > 0 ? (256 + p.C * nq / 2) / (p.C * nq) : 1;   // ~256 tasks
        }
        if (nsplit > p.R) nsplit = p.R;
        p.nsplit = nsplit;
        int nthreads = (p.C * nq * nsplit + 31) / 32 * 32;
        if (nthreads > kFastMaxThreads) nthreads = kFastMaxThreads;
        if (nthreads < 64) nthreads = 64;
        static const int env_nthreads = env_int("CLIPPPO_DISTURB_NTHREADS", 0);   // measurement knob: CTA size (multiple of 32, <= 384)
        if (env_nthreads >= 64 && env_nthreads <= kFastMaxThreads && env_nthreads % 32 == 0) nthreads = env_nthreads;
        p.nthreads = nthreads;
        p.log2S = 0;
        while ((1 << p.log2S) < S) ++p.log2S;
        p.magic_nq = static_cast<unsigned>((0x100000000ull + nq - 1) / nq);
        p.magic_nsplit = static_cast<unsigned>((0x100000000ull + nsplit - 1) / nsplit);
        p.magic_rs = static_cast<unsigned>((0x100000000ull + (p.R + 2 * P) - 1) / (p.R + 2 * P));
        const size_t smem = hdr + tile_bytes(S);
        int st = CLIPPPO_ERR_UNSUPPORTED;
        switch (K) {
            case 1: st = launch_disturb_fast<1>(p, smem, stream); break;
            case 3: st = launch_disturb_fast<3>(p, smem, stream); break;
            case 5: st = launch_disturb_fast<5>(p, smem, stream); break;
            case 7: st = launch_disturb_fast<7>(p, smem, stream); break;
        }
        if (st == CLIPPPO_ERR_CUDA && S > 8) {      // 16-CTA cluster refused on this device / partition: portable size
            cudaGetLastError();
            return run_disturb(p, k1d_host, k, stream, 8);
        }
        return st;
    }
    if (p.philox || p.out_scale != 1.0f) return CLIPPPO_ERR_UNSUPPORTED;      // fast-kernel features: the caller falls back (torch noise / a `* scale` pass)
    // ---- general path (disturb_kernel): any strides, NHWC / uint8 I/O, odd widths, wide kernels ----
    // stripes per image: the smallest cluster whose stripe fits the occupancy target
    auto smem_for = [&](int S) {
        const int R = (p.H + S - 1) / S;
        return (size_t)(kSmemHeaderFloats + (size_t)p.C * (R + 2 * P) * p.W) * sizeof(float);
    };
    const size_t budgets[3] = {static_cast<size_t>(env_budget) * 1024, 113 * 1024, 227 * 1024};
    int S = 0;
    for (int bi = 0; bi < 3 && !S; ++bi)
        for (int cand = 1; cand <= max_cluster && cand <= 8; cand *= 2)
            if (smem_for(cand) <= budgets[bi]) { S = cand; break; }
    if (!S) return CLIPPPO_ERR_UNSUPPORTED;
    // (a larger target, 8 CTAs per SM, was measured: 64 frames 24.5 us instead of 23.4, 256 frames 71.6 instead of 61.6)
    S = widen_for_small_batches(S, p.B, p.H, P, max_cluster < 8 ? max_cluster : 8, kNumSMs);
    p.S = S;
    p.R = (p.H + S - 1) / S;
    {   // blur-phase task shape: minimise rounds x (rows per task + ring warm-up rows)
        const int nq = (p.W + 3) / 4;
        long best_cost = -1;
        p.nsplit = 1;
        for (int sp = 1; sp <= 8 && sp <= p.R; ++sp) {
            const long tasks = (long)p.C * nq * sp;
            const long rounds = (tasks + kDisturbThreads - 1) / kDisturbThreads;
            const long cost = rounds * ((p.R + sp - 1) / sp + 2 * P);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; p.nsplit = sp; }
        }
    }
    const size_t smem = smem_for(S);
    int st = CLIPPPO_ERR_UNSUPPORTED;
    switch (K) {
        case 1:  st = launch_disturb<1>(p, smem, stream); break;
        case 3:  st = launch_disturb<3>(p, smem, stream); break;
        case 5:  st = launch_disturb<5>(p, smem, stream); break;
        case 7:  st = launch_disturb<7>(p, smem, stream); break;
        case 9:  st = launch_disturb<9>(p, smem, stream); break;
        case 11: st = launch_disturb<11>(p, smem, stream); break;
        case 13: st = launch_disturb<13>(p, smem, stream); break;
        case 15: st = launch_disturb<15>(p, smem, stream); break;
    }
    if (st == CLIPPPO_ERR_CUDA && S > 8) {          // 16-CTA cluster refused on this device / partition: portable size
        cudaGetLastError();
        return run_disturb(p, k1d_host, k, stream, 8);
    }
    return st;
}

static bool is_contig_nchw(const long long s[4], int C, int H, int W) {
    return s[3] == 1 && s[2] == W && s[1] == (long long)H * W && s[0] == (long long)C * H * W;
}
// the [B,3,H,W] view of NHWC memory (x.permute(0, 3, 1, 2) of a contiguous [B,H,W,3] tensor)
static bool is_nhwc_dense(const long long s[4], int B, int H, int W) {
    return s[1] == 1 && s[3] == 3 && s[2] == 3LL * W && (B == 1 || (s[0] >= 3LL * H * W && s[0] % 4 == 0));
}
// every image one contiguous [C,H,W] block; the images may sit at any non-overlapping, 16-byte-aligned distance
static bool is_image_contig(const long long s[4], int B, int C, int H, int W) {
    return s[3] == 1 && s[2] == W && (C == 1 || s[1] == (long long)H * W) &&
           (B == 1 || (s[0] >= (long long)C * H * W && s[0] % 4 == 0));
}

}  // namespace clipppo

using namespace clipppo;

extern "C" int clipppo_disturb_f32(const float* x, const int64_t x_strides_host[4],
                                   const float* noise, const int64_t noise_strides_host[4],
                                   float* out, int B, int C, int H, int W, int stages,
                                   float noise_sigma, float contrast,
                                   const float* k1d_host, int k,
                                   int sh, int sw, int ph, int pw, clipppo_stream_t stream) {
    DisturbParams p = {};
    p.x = x; p.noise = noise; p.out = out;
    p.B = B; p.C = C; p.H = H; p.W = W;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    for (int i = 0; i < 4; ++i) {
        const long long dflt[4] = {(long long)C * H * W, (long long)H * W, W, 1};
        p.xs[i] = x_strides_host ? x_strides_host[i] : dflt[i];
        p.ns[i] = noise_strides_host ? noise_strides_host[i] : dflt[i];
    }
    p.stages = stages & CLIPPPO_STAGE_ALL;
    p.sigma_n = noise_sigma;
    p.c = contrast;
    p.omc = static_cast<float>(1.0 - static_cast<double>(contrast));
    p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw;
    p.io_mode = 0;
    const bool need_noise = (p.stages & CLIPPPO_STAGE_NOISE) != 0;
    p.batch_contig = is_contig_nchw(p.xs, C, H, W) && (!need_noise || is_contig_nchw(p.ns, C, H, W));
    const bool aligned = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                         (!need_noise || reinterpret_cast<uintptr_t>(noise) % 16 == 0);
    p.fast = is_image_contig(p.xs, B, C, H, W) && (!need_noise || is_image_contig(p.ns, B, C, H, W)) && aligned;
    if (!p.fast && C == 3 && aligned && is_nhwc_dense(p.xs, B, H, W) && (!need_noise || is_nhwc_dense(p.ns, B, H, W))) {
        p.fast = 1;                 // the NHWC-strided view of the MiniGrid call site / the reference benchmark
        p.nhwc = 1;
    }
    return run_disturb(p, k1d_host, k, as_stream(stream));
}

extern "C" int clipppo_disturb_u8_f32(const uint8_t* x, const float* noise, float* out, int B, int C, int H, int W, int stages,
                                      float noise_sigma, float contrast, const float* k1d_host, int k,
                                      int sh, int sw, int ph, int pw, clipppo_stream_t stream) {
    DisturbParams p = {};
    p.x = x; p.noise = noise; p.out = out;
    p.B = B; p.C = C; p.H = H; p.W = W;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    const long long dflt[4] = {(long long)C * H * W, (long long)H * W, W, 1};
    for (int i = 0; i < 4; ++i) { p.xs[i] = dflt[i]; p.ns[i] = dflt[i]; }
    p.stages = stages & CLIPPPO_STAGE_ALL;
    p.sigma_n = noise_sigma;
    p.c = contrast;
    p.omc = static_cast<float>(1.0 - static_cast<double>(contrast));
    p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw;
    p.io_mode = 0;
    p.x_u8 = 1;
    const bool need_noise = (p.stages & CLIPPPO_STAGE_NOISE) != 0;
    if ((W % 4) || (reinterpret_cast<uintptr_t>(x) % 4)) return CLIPPPO_ERR_UNSUPPORTED;
    if (need_noise && noise && (reinterpret_cast<uintptr_t>(noise) % 16)) return CLIPPPO_ERR_ALIGN;
    p.fast = 1;
    p.batch_contig = 1;
    return run_disturb(p, k1d_host, k, as_stream(stream));
}

extern "C" int clipppo_disturb_nhwc_u8(const void* obs, int obs_is_f32,
                                       const float* noise, const int64_t noise_strides_host[4],
                                       uint8_t* out_nhwc, int B, int H, int W, int C, int stages,
                                       float noise_sigma, float contrast, const float* k1d_host, int k,
                                       int sh, int sw, int ph, int pw, clipppo_stream_t stream) {
    DisturbParams p = {};
    p.x = obs; p.noise = noise; p.out = out_nhwc;
    p.B = B; p.C = C; p.H = H; p.W = W;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    // logical [B,C,H,W] view of NHWC memory
    p.xs[0] = (long long)H * W * C; p.xs[1] = 1; p.xs[2] = (long long)W * C; p.xs[3] = C;
    const long long dflt[4] = {(long long)C * H * W, (long long)H * W, W, 1};
    for (int i = 0; i < 4; ++i) p.ns[i] = noise_strides_host ? noise_strides_host[i] : dflt[i];
    p.stages = stages & CLIPPPO_STAGE_ALL;
    p.sigma_n = noise_sigma;
    p.c = contrast;
    p.omc = static_cast<float>(1.0 - static_cast<double>(contrast));
    p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw;
    p.io_mode = obs_is_f32 ? 2 : 1;
    p.fast = 0;
    return run_disturb(p, k1d_host, k, as_stream(stream));
}

// The extended form (additive): uint8 or fp32 frames, an output scale, optional in-kernel noise.
extern "C" int clipppo_disturb_ex(const clipppo_disturb_desc* d, clipppo_stream_t stream) {
    if (!d) return CLIPPPO_ERR_NULL;
    const bool philox = (d->flags & CLIPPPO_DISTURB_PHILOX) != 0;
    if (philox && d->noise) return CLIPPPO_ERR_BAD_SHAPE;                  // either a noise tensor or the generator
    DisturbParams p = {};
    p.x = d->x; p.noise = d->noise; p.out = d->out;
    p.B = d->B; p.C = d->C; p.H = d->H; p.W = d->W;
    const int B = d->B, C = d->C, H = d->H, W = d->W;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    const long long dflt[4] = {(long long)C * H * W, (long long)H * W, W, 1};
    for (int i = 0; i < 4; ++i) {
        p.xs[i] = d->x_strides_host ? d->x_strides_host[i] : dflt[i];
        p.ns[i] = d->noise_strides_host ? d->noise_strides_host[i] : dflt[i];
    }
    p.stages = d->stages & CLIPPPO_STAGE_ALL;
    p.sigma_n = d->noise_sigma;
    p.c = d->contrast;
    p.omc = static_cast<float>(1.0 - static_cast<double>(d->contrast));
    p.sh = d->sh; p.sw = d->sw; p.ph = d->ph; p.pw = d->pw;
    p.io_mode = 0;
    p.out_scale = d->out_scale == 0.0f ? 1.0f : d->out_scale;
    p.philox = philox ? 1 : 0;
    p.philox_seed = d->philox_seed;
    p.philox_offset = d->philox_offset;
    const bool need_noise = (p.stages & CLIPPPO_STAGE_NOISE) != 0 && !philox;
    if (d->x_dtype == CLIPPPO_IMG_U8) {
        if (!is_contig_nchw(p.xs, C, H, W) || (need_noise && !is_contig_nchw(p.ns, C, H, W))) return CLIPPPO_ERR_UNSUPPORTED;
        if ((W % 4) || (reinterpret_cast<uintptr_t>(d->x) % 4)) return CLIPPPO_ERR_UNSUPPORTED;
        if (need_noise && d->noise && (reinterpret_cast<uintptr_t>(d->noise) % 16)) return CLIPPPO_ERR_ALIGN;
        p.x_u8 = 1; p.fast = 1; p.batch_contig = 1;
    } else {
        p.batch_contig = is_contig_nchw(p.xs, C, H, W) && (!need_noise || is_contig_nchw(p.ns, C, H, W));
        const bool aligned = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(d->x) % 16 == 0) &&
                             (!need_noise || reinterpret_cast<uintptr_t>(d->noise) % 16 == 0);
        p.fast = is_image_contig(p.xs, B, C, H, W) && (!need_noise || is_image_contig(p.ns, B, C, H, W)) && aligned;
        if (!p.fast && C == 3 && aligned && is_nhwc_dense(p.xs, B, H, W) && (!need_noise || is_nhwc_dense(p.ns, B, H, W))) {
            p.fast = 1;
            p.nhwc = 1;
        }
    }
    // the generator indexes draws by GLOBAL image number, so a batch sharded over ranks sees the noise of the whole batch
    p.philox_first = d->first_image;
    return run_disturb(p, d->k1d_host, d->k, as_stream(stream));
}
