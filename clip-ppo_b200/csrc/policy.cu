// §8f-1 - the PPO encoder of the reference's Agent (NatureCNN: conv 8x8 s4 -> 4x4 s2 -> 3x3 s1 -> FC 3136 -> 512, ReLU after
// every layer; reference minigrid_experiments/clip_ppo/clip_ppo_minigrid.py:229-242, atari_experiments/clip_ppo/
// clip_ppo_atari.py:196-209), forward and backward, in fp32.
//
// The reference trains this network in fp32 (cuDNN may use TF32 for the convolutions, the Linear layer may not): the
// replacement keeps every product and every accumulation in fp32 FMA arithmetic - bit-for-bit deterministic, no atomics -
// so its outputs and all its gradients match PyTorch's fp32 path to summation-order noise (tests/test_policy_gpu.py: <= 1e-4
// relative, measured ~1e-6).  Tensor cores are deliberately not used: kind::tf32 would cost three decimal digits, and the
// whole network is 46 MFLOP per frame forward + backward against 8.8 GFLOP for the frozen tower next to it.
//
// Layout: activations stay NHWC (rows = output positions, columns = channels - exactly what a GEMM over unfolded rows
// produces); convolution weights are used in (oc, kh, kw, c) order and the FC weight in (h, w, c) column order, permuted
// from / back to PyTorch's layouts by tiny kernels, so every large tensor is read and written contiguously.
// Every convolution is an IMPLICIT GEMM: no im2col matrix is ever written - the GEMM's loader gathers each float4 group
// of a filter row straight from the observation / activation (a 144-entry offset table in the kernel parameters).
//   forward   act = relu(gather(in) W^T + b)               per convolution;  hidden = relu(act3 Wfc^T + bfc)
//   backward  dZ = dA * (A > 0) (in the epilogue of the GEMM that produces dA);
//             dW = dZ^T gather(in) split over the rows, partial sums added in a fixed order;  db = column sums;
//             dA_prev = transposed convolution of dZ as a gathered GEMM with border taps predicated off - for the
//             stride-2 layer one GEMM per parity class of the input grid, so no multiply is wasted on a zero tap.
#include <stdint.h>

#include "common.cuh"

namespace clipppo {
namespace {

// ---- implicit im2col: a GEMM operand gathered straight from the tensor it would be unfolded from ---------------------
// Rows m = (n, i, j) (j fastest, NI x NJ per image); row origin = n sn + i si + j sj.  K is cut into float4 groups q:
// koff[q] is the group's element offset from the row origin, ktap[q] the filter tap it belongs to.  A tap (ta, tb) is
// valid for a row when (i - ta, j - tb) lies in [0, VI) x [0, VJ) (VI = 0: every tap is valid - the forward geometry);
// invalid groups read as zero (the transposed convolution of the input-gradient pass at the image border).
struct Gather {
    int NI, NJ;
    long long sn;
    int si, sj;
    int VI, VJ;
    int koff[144];
    signed char ta[16], tb[16];
    unsigned char ktap[144];
};
// where output row m = (n, i, j) goes: base + n sn + i si + j sj (+ column).  The mask operand is addressed the same way.
struct RowOut {
    int NI, NJ;
    long long sn;
    int si, sj;
    long long base;
};

// out[i] = (a[i] > 0) ? g[i] : 0
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ a, long long n, float* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = a[i] > 0.f ? g[i] : 0.f;
}

// generic 4-index permutation of a small tensor: dst[(i0, i1, i2, i3) in dst order] = src[...]; used for
// (oc, c, kh, kw) <-> (oc, kh, kw, c) and the FC weight's (o, c, h, w) <-> (o, h, w, c)
__global__ void __launch_bounds__(256)
permute_0231_kernel(const float* __restrict__ src, float* __restrict__ dst, int d0, int d1, int d2, int d3, int inverse) {
    // forward: src [d0, d1, d2, d3] -> dst [d0, d2, d3, d1];   inverse: src [d0, d2, d3, d1] -> dst [d0, d1, d2, d3]
    const long long total = static_cast<long long>(d0) * d1 * d2 * d3;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int i3 = static_cast<int>(i % d3), i2 = static_cast<int>((i / d3) % d2), i1 = static_cast<int>((i / (static_cast<long long>(d3) * d2)) % d1);
        const long long i0 = i / (static_cast<long long>(d3) * d2 * d1);
        const long long j = ((i0 * d2 + i2) * d3 + i3) * d1 + i1;
        if (inverse) dst[i] = src[j]; else dst[j] = src[i];
    }
}

// dst [cols, rows] = src [rows, cols]^T (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    __shared__ float t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8)
        if (by + r < rows && bx + tx < cols) t[r][tx] = src[static_cast<long long>(by + r) * cols + bx + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (bx + r < cols && by + tx < rows) dst[static_cast<long long>(bx + r) * rows + by + tx] = t[tx][r];
}

// ---- C[rows, N] = epilogue(a_scale * A[M, K] B[N, K]^T): fp32 FMA, 128 x BN block tile, 8 x (BN/16) per thread -----------
// GATHER = 0: A is a plain row-major matrix;  GATHER = 1: A is gathered (implicit im2col, above) and the output rows are
// scattered through `ro`.
enum { EPI_NONE = 0, EPI_BIAS_RELU = 1, EPI_MASK = 2 };     // MASK: C = (mask[same address] > 0) ? acc : 0

template <int BN, int EPI, int GATHER, int RM = 8, bool CHUNKED = false>
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, const __grid_constant__ Gather ga, const float* __restrict__ B, float* __restrict__ Cm,
                const __grid_constant__ RowOut ro, long long M, int N, int K, float a_scale, const float* __restrict__ bias,
                const float* __restrict__ mask, int ksteps = 0, float* __restrict__ part = nullptr) {
    // ksteps > 0: the K loop is cut into canonical chunks of `ksteps` 16-wide k-steps whose sums are added in index order - the
    // SAME arithmetic whether one CTA walks all chunks (part == null: large batches) or blockIdx.z owns one chunk and writes its
    // raw partial tile to part[z][M][N] for splitk_finish_kernel to add (few output tiles: the env-step batches of the rollout,
    // where the K = 3136 loop of 8 CTAs would be the whole latency).  A row's result is bitwise independent of the batch size.
    // 256 threads as TY x TX; a thread owns RM rows (groups of 4) x 4 consecutive columns:
    //   BN = 64: 16 x 16 threads, 128 x 64 tile (RM = 8) or 64 x 64 (RM = 4, small M: twice the CTAs);   BN = 32: 32 x 8 threads, 256 x 32 tile
    constexpr int TX = BN / 4, TY = 256 / TX, BM = TY * RM, BK = 16, HALF = BM / 2;
    constexpr int A_F4 = BM * BK / 4 / 256;                            // float4 loads of the A tile per thread: 2 or 4
    __shared__ float sa[2][BK][BM + 4];
    __shared__ float sb[2][BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const long long m0 = static_cast<long long>(blockIdx.x) * BM;
    const int n0 = blockIdx.y * BN;
    // Accumulators as PACKED pairs of rows: acc2[ip][j] = (C[2 ip][j], C[2 ip + 1][j]).  On Blackwell a three-register FFMA issues
    // every other cycle per scheduler; FFMA2 (two independent round-to-nearest FMAs on 64-bit register pairs, bit-identical
    // to two FFMAs) is what reaches the fp32 pipe's full rate.  The a-fragment pairs are the aligned halves of the float4
    // shared-memory loads; a b value is broadcast into a pair with one move.
    constexpr int RP = RM / 2;
    float2 acc2[RP][4];
#pragma unroll
    for (int i = 0; i < RP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    // loaders: A tile = BM rows x 16 k (A_F4 float4 per thread, rows a_row + 64 h); B tile = BN rows x 16 k
    const int a_row = tid >> 2, a_k4 = (tid & 3) * 4;
    const float* arow[A_F4];
    int ai[A_F4], aj[A_F4];
    bool aok[A_F4];
#pragma unroll
    for (int h = 0; h < A_F4; ++h) {
        const long long m = m0 + a_row + 64 * h;
        aok[h] = m < M;
        ai[h] = aj[h] = 0;
        if (GATHER) {
            const long long per = static_cast<long long>(ga.NI) * ga.NJ;
            const long long n = m / per;
            const int rem = static_cast<int>(m - n * per);
            ai[h] = rem / ga.NJ; aj[h] = rem - ai[h] * ga.NJ;
            arow[h] = A + n * ga.sn + static_cast<long long>(ai[h]) * ga.si + static_cast<long long>(aj[h]) * ga.sj;
        } else {
            arow[h] = A + m * K;
        }
    }
    float4 ra[A_F4], rb;
    const int b_row = tid >> 2, b_k4 = (tid & 3) * 4;                  // B tile: BN * 4 float4, one per thread (threads < BN * 4)
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int h = 0; h < A_F4; ++h) {
            ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!aok[h]) continue;
            if (GATHER) {
                const int q = (k0 + a_k4) >> 2;
                bool ok = true;
                if (ga.VI) {
                    const int t = ga.ktap[q];
                    const int si = ai[h] - ga.ta[t], sj = aj[h] - ga.tb[t];
                    ok = si >= 0 && si < ga.VI && sj >= 0 && sj < ga.VJ;
                }
                if (ok) ra[h] = __ldg(reinterpret_cast<const float4*>(arow[h] + ga.koff[q]));
            } else {
                ra[h] = __ldg(reinterpret_cast<const float4*>(arow[h] + k0 + a_k4));       // K % 16 == 0 on every caller
            }
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b_row < BN && n0 + b_row < N) rb = __ldg(reinterpret_cast<const float4*>(B + static_cast<long long>(n0 + b_row) * K + k0 + b_k4));
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int h = 0; h < A_F4; ++h) {
            const int r = a_row + 64 * h;
            sa[buf][a_k4 + 0][r] = ra[h].x; sa[buf][a_k4 + 1][r] = ra[h].y; sa[buf][a_k4 + 2][r] = ra[h].z; sa[buf][a_k4 + 3][r] = ra[h].w;
        }
        if (b_row < BN) { sb[buf][b_k4 + 0][b_row] = rb.x; sb[buf][b_k4 + 1][b_row] = rb.y; sb[buf][b_k4 + 2][b_row] = rb.z; sb[buf][b_k4 + 3][b_row] = rb.w; }
    };
    const int nk_all = K / BK;
    if constexpr (!CHUNKED) ksteps = 0;                               // the backward's GEMMs: one plain K walk, no second accumulator set
    const bool split = CHUNKED && ksteps > 0 && part != nullptr;
    const int kb0 = split ? static_cast<int>(blockIdx.z) * ksteps : 0;
    const int nk = split ? min(nk_all, kb0 + ksteps) : nk_all;
    float2 tot2[CHUNKED ? RP : 1][4];                                  // chunk sums so far (chunked, unsplit walk only)
#pragma unroll
    for (int i = 0; i < (CHUNKED ? RP : 1); ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tot2[i][j] = make_float2(0.f, 0.f);
    int left = ksteps;                                                 // k-steps left in the current chunk
    load_tiles(kb0 * BK);
    store_tiles(kb0 & 1);
    __syncthreads();
    for (int kb = kb0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tiles((kb + 1) * BK);                    // global loads in flight under the FMAs
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sa[buf][k][ty * 4]);
            float4 a1 = a0;
            if constexpr (RM == 8) a1 = *reinterpret_cast<const float4*>(&sa[buf][k][HALF + ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&sb[buf][k][tx * 4]);
            const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
            const float2 bb[4] = {make_float2(b4.x, b4.x), make_float2(b4.y, b4.y), make_float2(b4.z, b4.z), make_float2(b4.w, b4.w)};
#pragma unroll
            for (int i = 0; i < RP; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc2[i][j] = __ffma2_rn(ap[i], bb[j], acc2[i][j]);
        }
        if constexpr (CHUNKED) {
            if (ksteps > 0 && !split && (--left == 0 || kb + 1 == nk)) {  // chunk boundary: fold the chunk's sum in, start the next one at zero
#pragma unroll
                for (int i = 0; i < RP; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) { tot2[i][j] = __fadd2_rn(tot2[i][j], acc2[i][j]); acc2[i][j] = make_float2(0.f, 0.f); }
                left = ksteps;
            }
        }
        if (kb + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
    if constexpr (CHUNKED) {
        if (ksteps > 0 && !split) {
#pragma unroll
            for (int i = 0; i < RP; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc2[i][j] = tot2[i][j];
        }
    }
    float acc[RM][4];                                                  // unpack: the epilogues address single rows
#pragma unroll
    for (int i = 0; i < RP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[2 * i][j] = acc2[i][j].x; acc[2 * i + 1][j] = acc2[i][j].y; }
    if (split) {
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const long long m = m0 + (i < 4 ? ty * 4 + i : HALF + ty * 4 + (i - 4));
            const int n = n0 + tx * 4;
            if (m < M && n < N)
                *reinterpret_cast<float4*>(part + (static_cast<long long>(blockIdx.z) * M + m) * N + n) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const long long m = m0 + (i < 4 ? ty * 4 + i : HALF + ty * 4 + (i - 4));
        if (m >= M) continue;
        long long obase = m * N;
        if (GATHER) {
            const long long per = static_cast<long long>(ro.NI) * ro.NJ;
            const long long n = m / per;
            const int rem = static_cast<int>(m - n * per);
            const int oi = rem / ro.NJ, oj = rem - oi * ro.NJ;
            obase = ro.base + n * ro.sn + static_cast<long long>(oi) * ro.si + static_cast<long long>(oj) * ro.sj;
        }
        const int n = n0 + tx * 4;
        if (n >= N) continue;                                          // N % 4 == 0 on every caller
        float4 v = make_float4(acc[i][0] * a_scale, acc[i][1] * a_scale, acc[i][2] * a_scale, acc[i][3] * a_scale);
        if (EPI == EPI_BIAS_RELU) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n));
            v = make_float4(fmaxf(v.x + bb.x, 0.f), fmaxf(v.y + bb.y, 0.f), fmaxf(v.z + bb.z, 0.f), fmaxf(v.w + bb.w, 0.f));
        }
        if (EPI == EPI_MASK) {
            const float4 mk = __ldg(reinterpret_cast<const float4*>(mask + obase + n));
            v = make_float4(mk.x > 0.f ? v.x : 0.f, mk.y > 0.f ? v.y : 0.f, mk.z > 0.f ? v.z : 0.f, mk.w > 0.f ? v.w : 0.f);
        }
        *reinterpret_cast<float4*>(Cm + obase + n) = v;
    }
}

// ---- part[s][P, Q] = sum over the rows r of split s of A[r, P]^T B[r, Q]  (weight gradients: reduction over positions) -------
// BP x (16 TQ) output tile, 16 reduction rows per step, (BP/16) x TQ per thread (TQ in groups of 4 consecutive columns);
// B is gathered (implicit im2col) when GATHER = 1.  The splits are added in index order by reduce_parts_kernel: deterministic.
template <int BP, int TQ, int GATHER>
__global__ void __launch_bounds__(256)
sgemm_tn_split_kernel(const float* __restrict__ A, const float* __restrict__ B, const __grid_constant__ Gather gb,
                      float* __restrict__ part, long long R, int P, int Q, long long rows_per_split) {
    constexpr int BQ = 16 * TQ, BR = 16, TP = BP / 16, QG = TQ / 4;
    __shared__ float sa[BR][BP + 4];
    __shared__ float sb[BR][BQ + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int p0 = blockIdx.y * BP, q0 = blockIdx.x * BQ;
    const long long r_begin = static_cast<long long>(blockIdx.z) * rows_per_split;
    const long long r_end = (r_begin + rows_per_split < R) ? r_begin + rows_per_split : R;
    constexpr int TPP = TP / 2;                                        // packed pairs of output rows (FFMA2, see sgemm_nt_kernel)
    float2 acc2[TPP][TQ];
#pragma unroll
    for (int i = 0; i < TPP; ++i)
#pragma unroll
        for (int j = 0; j < TQ; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    const int lr = tid >> 4, lc = (tid & 15) * 4;                      // B: row lr, float4 groups lc + 64 g
    const int alr = (BP == 64) ? lr : (tid >> 3), alc = (BP == 64) ? lc : (tid & 7) * 4;     // A: 16 x BP (BP = 32: threads 0..127)
    const long long per = GATHER ? static_cast<long long>(gb.NI) * gb.NJ : 1;
    float4 va, vb[QG];
    auto load_step = [&](long long r0) {
        va = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const long long r = r0 + alr;
            if ((BP == 64 || tid < 128) && r < r_end && p0 + alc < P) va = __ldg(reinterpret_cast<const float4*>(A + r * P + p0 + alc));
        }
        const long long r = r0 + lr;
        const float* brow = B + r * Q;
        if (GATHER) {
            const long long n = r / per;
            const int rem = static_cast<int>(r - n * per);
            const int i = rem / gb.NJ, j = rem - i * gb.NJ;
            brow = B + n * gb.sn + static_cast<long long>(i) * gb.si + static_cast<long long>(j) * gb.sj;
        }
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            const int q = q0 + lc + 64 * g;
            vb[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r_end && q < Q) vb[g] = __ldg(reinterpret_cast<const float4*>(brow + (GATHER ? gb.koff[q >> 2] : q)));
        }
    };
    if (r_begin < r_end) load_step(r_begin);
    for (long long r0 = r_begin; r0 < r_end; r0 += BR) {
        __syncthreads();                                               // the previous step's FMAs are done with the tiles
        if (BP == 64 || tid < 128) *reinterpret_cast<float4*>(&sa[alr][alc]) = va;
#pragma unroll
        for (int g = 0; g < QG; ++g) *reinterpret_cast<float4*>(&sb[lr][lc + 64 * g]) = vb[g];
        __syncthreads();
        if (r0 + BR < r_end) load_step(r0 + BR);                       // next step's rows travel under this step's FMAs
#pragma unroll
        for (int k = 0; k < BR; ++k) {
            float2 ap[2];
            if constexpr (TP == 4) {
                const float4 a = *reinterpret_cast<const float4*>(&sa[k][ty * 4]);
                ap[0] = make_float2(a.x, a.y); ap[1] = make_float2(a.z, a.w);
            } else {
                ap[0] = *reinterpret_cast<const float2*>(&sa[k][ty * 2]);
                ap[1] = ap[0];
            }
#pragma unroll
            for (int g = 0; g < QG; ++g) {
                const float4 b = *reinterpret_cast<const float4*>(&sb[k][tx * 4 + 64 * g]);
                const float2 bb[4] = {make_float2(b.x, b.x), make_float2(b.y, b.y), make_float2(b.z, b.z), make_float2(b.w, b.w)};
#pragma unroll
                for (int i = 0; i < TPP; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc2[i][4 * g + j] = __ffma2_rn(ap[i], bb[j], acc2[i][4 * g + j]);
            }
        }
    }
    float acc[TP][TQ];
#pragma unroll
    for (int i = 0; i < TPP; ++i)
#pragma unroll
        for (int j = 0; j < TQ; ++j) { acc[2 * i][j] = acc2[i][j].x; acc[2 * i + 1][j] = acc2[i][j].y; }
    float* out = part + static_cast<long long>(blockIdx.z) * P * Q;
#pragma unroll
    for (int i = 0; i < TP; ++i) {
        const int p = p0 + ty * TP + i;
        if (p >= P) continue;
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            const int q = q0 + tx * 4 + 64 * g;
            if (q < Q) *reinterpret_cast<float4*>(out + static_cast<long long>(p) * Q + q) =
                make_float4(acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
        }
    }
}

__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* __restrict__ part, int splits, long long n, float scale, float* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[static_cast<long long>(k) * n + i];
        out[i] = s * scale;
    }
}

// the transposed-convolution weights of an input-gradient pass: out[c][t][oc] = w[oc][c][kh_t][kw_t] for the listed taps
struct TapList { int n; int kh[16], kw[16]; };
__global__ void __launch_bounds__(256)
prep_dgrad_w_kernel(const float* __restrict__ w, int OC, int C, int KH, int KW, TapList taps, float* __restrict__ out) {
    const int total = C * taps.n * OC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int oc = i % OC, t = (i / OC) % taps.n, c = i / (OC * taps.n);
        out[i] = w[((static_cast<long long>(oc) * C + c) * KH + taps.kh[t]) * KW + taps.kw[t]];
    }
}

// part[s][n] = sum of column n over the rows of split s (bias gradients); reduced by reduce_parts_kernel
__global__ void __launch_bounds__(256)
colsum_split_kernel(const float* __restrict__ X, long long R, int N, long long rows_per_split, float* __restrict__ part) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const long long r_begin = static_cast<long long>(blockIdx.y) * rows_per_split;
    const long long r_end = (r_begin + rows_per_split < R) ? r_begin + rows_per_split : R;
    float s = 0.f;
    if (n < N)
        for (long long r = r_begin + wy; r < r_end; r += 8) s += X[r * N + n];
    red[wy][lane] = s;
    __syncthreads();
    if (wy == 0 && n < N) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        part[static_cast<long long>(blockIdx.y) * N + n] = t;
    }
}

inline unsigned grid_for(long long n) {
    const long long b = (n + 255) / 256;
    return static_cast<unsigned>(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

// C = epi(a_scale * sum_z part[z]) for the split-K launches of sgemm_nt_kernel: splits added in index order (deterministic)
template <int EPI, int GATHER>
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ part, int splits, float* __restrict__ Cm, const __grid_constant__ RowOut ro,
                     long long M, int N, float a_scale, const float* __restrict__ bias, const float* __restrict__ mask) {
    const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;      // float4 index into [M, N]
    const int n4 = N / 4;
    if (q >= M * n4) return;
    const long long m = q / n4;
    const int n = static_cast<int>(q - m * n4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
        const float4 pz = __ldg(reinterpret_cast<const float4*>(part + (static_cast<long long>(z) * M + m) * N + n));
        v.x += pz.x; v.y += pz.y; v.z += pz.z; v.w += pz.w;
    }
    v = make_float4(v.x * a_scale, v.y * a_scale, v.z * a_scale, v.w * a_scale);
    long long obase = m * N;
    if (GATHER) {
        const long long per = static_cast<long long>(ro.NI) * ro.NJ;
        const long long img = m / per;
        const int rem = static_cast<int>(m - img * per);
        const int oi = rem / ro.NJ, oj = rem - oi * ro.NJ;
        obase = ro.base + img * ro.sn + static_cast<long long>(oi) * ro.si + static_cast<long long>(oj) * ro.sj;
    }
    if (EPI == EPI_BIAS_RELU) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n));
        v = make_float4(fmaxf(v.x + bb.x, 0.f), fmaxf(v.y + bb.y, 0.f), fmaxf(v.z + bb.z, 0.f), fmaxf(v.w + bb.w, 0.f));
    }
    if (EPI == EPI_MASK) {
        const float4 mk = __ldg(reinterpret_cast<const float4*>(mask + obase + n));
        v = make_float4(mk.x > 0.f ? v.x : 0.f, mk.y > 0.f ? v.y : 0.f, mk.z > 0.f ? v.z : 0.f, mk.w > 0.f ? v.w : 0.f);
    }
    *reinterpret_cast<float4*>(Cm + obase + n) = v;
}

// Canonical chunk length (k-steps) of a forward GEMM with `nk` k-steps: a function of K only, so that every batch size adds
// the same chunk sums in the same order.  `*split` = launch one CTA per (tile, chunk): when the tiles alone leave most of the
// part idle and the partial tiles fit the buffer.
int plan_ksteps(long long ctas, int nk, long long M, int N, size_t part_floats, bool* split) {
    *split = false;
    if (nk < 8) return 0;
    const int ksteps = max(4, (nk + 27) / 28);
    const int splits = (nk + ksteps - 1) / ksteps;
    *split = ctas < 148 && static_cast<size_t>(splits) * static_cast<size_t>(M) * N <= part_floats;
    return ksteps;
}

const Gather kNoGather = {};
const RowOut kNoRowOut = {};

// C = epi(a_scale * A B^T) with a plain row-major A [M, K]
template <int EPI>
int launch_nt_plain(const float* A, const float* B, float* Cm, long long M, int N, int K, const float* bias, const float* mask, cudaStream_t st,
                    float* part = nullptr, size_t part_floats = 0) {
    if (M * ((N + 63) / 64) <= 128LL * 148 * 2) {                     // few row tiles (the FC layer): 64-row tiles, twice the CTAs
        dim3 grid(static_cast<unsigned>((M + 63) / 64), (N + 63) / 64);
        bool split = false;
        const int ksteps = part ? plan_ksteps(static_cast<long long>(grid.x) * grid.y, K / 16, M, N, part_floats, &split) : 0;
        if (split) {                                                  // an env-step batch: the K = 3136 loop of 8 CTAs would be the whole latency
            const int splits = (K / 16 + ksteps - 1) / ksteps;
            grid.z = splits;
            sgemm_nt_kernel<64, EPI, 0, 4, true><<<grid, 256, 0, st>>>(A, kNoGather, B, Cm, kNoRowOut, M, N, K, 1.0f, bias, mask, ksteps, part);
            CLIPPPO_CHECK_LAUNCH();
            splitk_finish_kernel<EPI, 0><<<static_cast<unsigned>((M * (N / 4) + 255) / 256), 256, 0, st>>>(part, splits, Cm, kNoRowOut, M, N, 1.0f, bias, mask);
            CLIPPPO_CHECK_LAUNCH();
            return CLIPPPO_OK;
        }
        if (ksteps > 0) sgemm_nt_kernel<64, EPI, 0, 4, true><<<grid, 256, 0, st>>>(A, kNoGather, B, Cm, kNoRowOut, M, N, K, 1.0f, bias, mask, ksteps, nullptr);
        else sgemm_nt_kernel<64, EPI, 0, 4><<<grid, 256, 0, st>>>(A, kNoGather, B, Cm, kNoRowOut, M, N, K, 1.0f, bias, mask);
    } else {
        dim3 grid(static_cast<unsigned>((M + 127) / 128), (N + 63) / 64);
        bool split = false;
        const int ksteps = part ? plan_ksteps(1 << 20, K / 16, M, N, part_floats, &split) : 0;      // same chunks, never split
        if (ksteps > 0) sgemm_nt_kernel<64, EPI, 0, 8, true><<<grid, 256, 0, st>>>(A, kNoGather, B, Cm, kNoRowOut, M, N, K, 1.0f, bias, mask, ksteps, nullptr);
        else sgemm_nt_kernel<64, EPI, 0><<<grid, 256, 0, st>>>(A, kNoGather, B, Cm, kNoRowOut, M, N, K, 1.0f, bias, mask);
    }
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}
// the same with A gathered through `ga` and the output rows placed through `ro`
template <int EPI>
int launch_nt_gather(const float* A, const Gather& ga, const float* B, float* Cm, const RowOut& ro, long long M, int N, int K,
                     float a_scale, const float* bias, const float* mask, cudaStream_t st, float* part = nullptr, size_t part_floats = 0) {
    if (part && N > 32) {
        dim3 grid(static_cast<unsigned>((M + 127) / 128), (N + 63) / 64);
        bool split = false;
        const int ksteps = plan_ksteps(static_cast<long long>(grid.x) * grid.y, K / 16, M, N, part_floats, &split);
        if (split) {
            const int splits = (K / 16 + ksteps - 1) / ksteps;
            grid.z = splits;
            sgemm_nt_kernel<64, EPI, 1, 8, true><<<grid, 256, 0, st>>>(A, ga, B, Cm, ro, M, N, K, a_scale, bias, mask, ksteps, part);
            CLIPPPO_CHECK_LAUNCH();
            splitk_finish_kernel<EPI, 1><<<static_cast<unsigned>((M * (N / 4) + 255) / 256), 256, 0, st>>>(part, splits, Cm, ro, M, N, a_scale, bias, mask);
            CLIPPPO_CHECK_LAUNCH();
            return CLIPPPO_OK;
        }
        sgemm_nt_kernel<64, EPI, 1, 8, true><<<grid, 256, 0, st>>>(A, ga, B, Cm, ro, M, N, K, a_scale, bias, mask, ksteps, nullptr);   // same chunks, one CTA per tile
        CLIPPPO_CHECK_LAUNCH();
        return CLIPPPO_OK;
    }
    if (N <= 32) {
        dim3 grid(static_cast<unsigned>((M + 255) / 256), 1);          // 256 x 32 tiles
        sgemm_nt_kernel<32, EPI, 1><<<grid, 256, 0, st>>>(A, ga, B, Cm, ro, M, N, K, a_scale, bias, mask);
    } else {
        dim3 grid(static_cast<unsigned>((M + 127) / 128), (N + 63) / 64);
        sgemm_nt_kernel<64, EPI, 1><<<grid, 256, 0, st>>>(A, ga, B, Cm, ro, M, N, K, a_scale, bias, mask);
    }
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

constexpr int kMaxSplits = 160;      // sizes the partial-sum buffer: 160 x 64 x 576 floats = 23.6 MB
// column-tile width of the weight-gradient GEMM for a given Q (the convolution's K): the widest of 256 / 192 / 128 that divides it
inline int pick_tq(int P, int Q) {
    if (P <= 32 && Q % 256 == 0) return 16;
    if (Q % 192 == 0) return 12;
    return 8;
}
constexpr size_t kPartFloats = static_cast<size_t>(kMaxSplits) * 64 * 576;     // capacity of the partial-sum buffer (>= 2 FC weights)
inline int pick_splits(long long R, int P, int Q, int bq) {
    const long long tiles = static_cast<long long>((P + 63) / 64) * ((Q + bq - 1) / bq);
    long long s = (148 * 4) / tiles;                                     // about four co-resident CTAs per SM, a whole number of waves
    const long long by_buffer = static_cast<long long>(kPartFloats / (static_cast<size_t>(P) * Q));
    if (s > by_buffer) s = by_buffer;
    const long long max_by_rows = (R + 255) / 256;
    if (s > max_by_rows) s = max_by_rows;
    return static_cast<int>(s < 1 ? 1 : s);
}

template <int BP, int TQ>
void launch_tn_t(const float* A, const float* B, const Gather* gb, float* part, long long R, int P, int Q, long long rps, int splits,
                 cudaStream_t st) {
    dim3 grid((Q + 16 * TQ - 1) / (16 * TQ), (P + BP - 1) / BP, splits);
    if (gb) sgemm_tn_split_kernel<BP, TQ, 1><<<grid, 256, 0, st>>>(A, B, *gb, part, R, P, Q, rps);
    else sgemm_tn_split_kernel<BP, TQ, 0><<<grid, 256, 0, st>>>(A, B, kNoGather, part, R, P, Q, rps);
}

// out[P, Q] = scale * A[R, P]^T B[R, Q] through `part`; B plain (gb == nullptr) or gathered
int launch_tn(const float* A, const float* B, const Gather* gb, float* out, float* part, long long R, int P, int Q, float scale,
              cudaStream_t st) {
    const int tq = pick_tq(P, Q);
    const int splits = pick_splits(R, P, Q, 16 * tq);
    const long long rps = ((R + splits - 1) / splits + 15) / 16 * 16;
    if (P <= 32) {
        if (tq == 16) launch_tn_t<32, 16>(A, B, gb, part, R, P, Q, rps, splits, st);
        else if (tq == 12) launch_tn_t<32, 12>(A, B, gb, part, R, P, Q, rps, splits, st);
        else launch_tn_t<32, 8>(A, B, gb, part, R, P, Q, rps, splits, st);
    } else {
        if (tq == 12) launch_tn_t<64, 12>(A, B, gb, part, R, P, Q, rps, splits, st);
        else launch_tn_t<64, 8>(A, B, gb, part, R, P, Q, rps, splits, st);
    }
    CLIPPPO_CHECK_LAUNCH();
    const long long n = static_cast<long long>(P) * Q;
    reduce_parts_kernel<<<grid_for(n), 256, 0, st>>>(part, splits, n, scale, out);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

int launch_colsum(const float* X, long long R, int N, float* out, float* part, cudaStream_t st) {
    const int col_tiles = (N + 31) / 32;
    int splits = static_cast<int>((R + 511) / 512);
    const int want = (148 * 4 + col_tiles - 1) / col_tiles;             // enough CTAs to fill the machine
    if (splits > want) splits = want;
    if (splits < 1) splits = 1;
    const long long rps = (R + splits - 1) / splits;
    dim3 grid((N + 31) / 32, splits);
    colsum_split_kernel<<<grid, 256, 0, st>>>(X, R, N, rps, part);
    CLIPPPO_CHECK_LAUNCH();
    reduce_parts_kernel<<<grid_for(N), 256, 0, st>>>(part, splits, N, 1.0f, out);
    CLIPPPO_CHECK_LAUNCH();
    return CLIPPPO_OK;
}

// forward geometry of a convolution over a tensor with element strides (sn, sc, sh, sw): rows (n, oh, ow), K in
// (kh, kw, c) order when the channels are innermost (sc == 1, sw == C) and in PyTorch's (c, kh, kw) order when the width is
// (sw == 1); every float4 group of K must be contiguous in memory.
bool conv_gather(Gather* g, int C, int KH, int KW, int S, int OH, int OW, long long sn, long long sc, long long sh, long long sw,
                 bool* khwc) {
    *g = Gather{};
    g->NI = OH; g->NJ = OW; g->sn = sn; g->si = static_cast<int>(S * sh); g->sj = static_cast<int>(S * sw);
    g->VI = 0; g->VJ = 0;
    const int K = C * KH * KW;
    if (K % 16 || K / 4 > 144) return false;
    if (sc == 1 && sw == C && (KW * C) % 4 == 0) {
        *khwc = true;
        for (int q = 0; q < K / 4; ++q) {
            const int k = 4 * q, kh = k / (KW * C), within = k - kh * KW * C;
            g->koff[q] = static_cast<int>(kh * sh + within);
            g->ktap[q] = 0;
        }
        return (sh % 4 == 0) && (sn % 4 == 0) && ((S * sw) % 4 == 0);
    }
    if (sw == 1 && KW % 4 == 0) {
        *khwc = false;
        for (int q = 0; q < K / 4; ++q) {
            const int k = 4 * q, c = k / (KH * KW), kh = (k / KW) % KH, kw = k % KW;
            g->koff[q] = static_cast<int>(c * sc + kh * sh + kw);
            g->ktap[q] = 0;
        }
        return (sh % 4 == 0) && (sn % 4 == 0) && (sc % 4 == 0) && (S % 4 == 0);
    }
    return false;
}

struct Plan {
    long long M1, M2, M3;
    int K1, K2, K3, KF;
    // workspace offsets (floats)
    size_t act1, act2, act3, w1p, w2p, w3p, wfp, wT, wd, dz1, dz2, dz3, part, gtmp, total;
};

Plan make_plan(int mb, int C) {
    Plan p;
    p.M1 = static_cast<long long>(mb) * 400; p.M2 = static_cast<long long>(mb) * 81; p.M3 = static_cast<long long>(mb) * 49;
    p.K1 = C * 64; p.K2 = 512; p.K3 = 576; p.KF = 3136;
    size_t off = 0;
    auto take = [&](size_t n) { const size_t o = off; off += (n + 63) / 64 * 64; return o; };
    p.act1 = take(p.M1 * 32);                                        // [mb, 20, 20, 32]
    p.act2 = take(p.M2 * 64);                                        // [mb, 9, 9, 64]
    p.act3 = take(p.M3 * 64);                                        // [mb, 7, 7, 64] = the FC input in (h, w, c) order
    p.w1p = take(32 * p.K1); p.w2p = take(64 * p.K2); p.w3p = take(64 * p.K3); p.wfp = take(512 * static_cast<size_t>(p.KF));
    p.wT = take(512 * static_cast<size_t>(p.KF));                    // FC weight transposed (input-gradient GEMM)
    p.wd = take(4 * 32 * 256 > 64 * 576 ? 4 * 32 * 256 : 64 * 576);  // transposed-convolution weights (4 parity classes of conv2 / conv3)
    p.dz1 = take(p.M1 * 32); p.dz2 = take(p.M2 * 64); p.dz3 = take(p.M3 * 64);      // masked output gradients of the three convolutions
    // partial sums of the split reductions: conv weights reach kMaxSplits x 64 x 576, the FC weight at most 2 x 512 x 3136
    p.part = take(kPartFloats);                                      // split partial sums (weight gradients, column sums)
    p.gtmp = take(512 * static_cast<size_t>(p.KF));
    p.total = off + 64;                                              // + the forward's record of conv1's K order
    return p;
}

#define POL_TRY(expr) do { int st__ = (expr); if (st__ != CLIPPPO_OK) return st__; } while (0)

}  // namespace
}  // namespace clipppo

using namespace clipppo;

extern "C" int clipppo_nature_workspace_bytes(int mb, int channels, size_t* bytes) {
    if (!bytes) return CLIPPPO_ERR_NULL;
    if (mb <= 0 || channels <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    *bytes = make_plan(mb, channels).total * sizeof(float);
    return CLIPPPO_OK;
}

// obs: fp32 observations, logical [mb, C, 84, 84] with the given element strides (NCHW stacks, or the NHWC-strided view of
// MiniGrid frames); in_scale multiplies them on the way in (1/255: `_pre` / `x / 255.0` of the reference, folded into the first
// GEMM's epilogue).  weights: PyTorch layouts - conv [OC, C, KH, KW], fc [512, 3136] with columns in (c, h, w) order, biases.
// hidden: [mb, 512] = network(x).  The workspace keeps the three activations for the backward; no im2col matrix exists.
extern "C" int clipppo_nature_forward(const float* obs, const int64_t obs_strides_host[4], float in_scale, int mb, int channels,
                                      const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                      const float* b3, const float* wfc, const float* bfc, float* hidden, void* workspace,
                                      size_t workspace_bytes, clipppo_stream_t stream) {
    if (!obs || !obs_strides_host || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !wfc || !bfc || !hidden || !workspace) return CLIPPPO_ERR_NULL;
    if (mb <= 0 || channels <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 || reinterpret_cast<uintptr_t>(obs) % 16) return CLIPPPO_ERR_ALIGN;
    const Plan p = make_plan(mb, channels);
    if (p.total * sizeof(float) > workspace_bytes) return CLIPPPO_ERR_WORKSPACE;
    float* ws = static_cast<float*>(workspace);
    cudaStream_t st = as_stream(stream);
    Gather g1, g2, g3;
    bool khwc1 = false, khwc = false;
    if (!conv_gather(&g1, channels, 8, 8, 4, 20, 20, obs_strides_host[0], obs_strides_host[1], obs_strides_host[2], obs_strides_host[3], &khwc1))
        return CLIPPPO_ERR_UNSUPPORTED;              // the caller makes the observations contiguous first
    conv_gather(&g2, 32, 4, 4, 2, 9, 9, 400LL * 32, 1, 20LL * 32, 32, &khwc);
    conv_gather(&g3, 64, 3, 3, 1, 7, 7, 81LL * 64, 1, 9LL * 64, 64, &khwc);
    // weights into the K orders the gathers produce
    const float* w1k = w1;
    if (khwc1) {
        permute_0231_kernel<<<grid_for(32 * p.K1), 256, 0, st>>>(w1, ws + p.w1p, 32, channels, 8, 8, 0);
        CLIPPPO_CHECK_LAUNCH();
        w1k = ws + p.w1p;
    }
    permute_0231_kernel<<<grid_for(64 * p.K2), 256, 0, st>>>(w2, ws + p.w2p, 64, 32, 4, 4, 0);
    CLIPPPO_CHECK_LAUNCH();
    permute_0231_kernel<<<grid_for(64 * p.K3), 256, 0, st>>>(w3, ws + p.w3p, 64, 64, 3, 3, 0);
    CLIPPPO_CHECK_LAUNCH();
    permute_0231_kernel<<<grid_for(512LL * p.KF), 256, 0, st>>>(wfc, ws + p.wfp, 512, 64, 7, 7, 0);
    CLIPPPO_CHECK_LAUNCH();
    const RowOut r1{20, 20, 400LL * 32, 20 * 32, 32, 0}, r2{9, 9, 81LL * 64, 9 * 64, 64, 0}, r3{7, 7, 49LL * 64, 7 * 64, 64, 0};
    POL_TRY(launch_nt_gather<EPI_BIAS_RELU>(obs, g1, w1k, ws + p.act1, r1, p.M1, 32, p.K1, in_scale, b1, nullptr, st));
    // env-step batches (tens of frames): conv2 / conv3 / FC are a handful of output tiles with a long K loop - split over K
    // (partial sums in the backward's partial buffer, added in index order)
    POL_TRY(launch_nt_gather<EPI_BIAS_RELU>(ws + p.act1, g2, ws + p.w2p, ws + p.act2, r2, p.M2, 64, p.K2, 1.0f, b2, nullptr, st, ws + p.part, kPartFloats));
    POL_TRY(launch_nt_gather<EPI_BIAS_RELU>(ws + p.act2, g3, ws + p.w3p, ws + p.act3, r3, p.M3, 64, p.K3, 1.0f, b3, nullptr, st, ws + p.part, kPartFloats));
    POL_TRY(launch_nt_plain<EPI_BIAS_RELU>(ws + p.act3, ws + p.wfp, hidden, mb, 512, p.KF, bfc, nullptr, st, ws + p.part, kPartFloats));
    return CLIPPPO_OK;
}

// Gradients of every weight and bias given grad_hidden = dL/d hidden ([mb, 512]); `obs`, `hidden` and the workspace are those
// of the matching clipppo_nature_forward call.  Outputs in PyTorch layouts.  No gradient flows to the observations.
extern "C" int clipppo_nature_backward(const float* grad_hidden, const float* hidden, const float* obs,
                                       const int64_t obs_strides_host[4], float in_scale, int mb, int channels, const float* w2,
                                       const float* w3, float* gw1, float* gb1, float* gw2, float* gb2, float* gw3, float* gb3,
                                       float* gwfc, float* gbfc, void* workspace, size_t workspace_bytes, clipppo_stream_t stream) {
    if (!grad_hidden || !hidden || !obs || !obs_strides_host || !w2 || !w3 || !gw1 || !gb1 || !gw2 || !gb2 || !gw3 || !gb3 || !gwfc ||
        !gbfc || !workspace)
        return CLIPPPO_ERR_NULL;
    if (mb <= 0 || channels <= 0) return CLIPPPO_ERR_BAD_SHAPE;
    const Plan p = make_plan(mb, channels);
    if (p.total * sizeof(float) > workspace_bytes) return CLIPPPO_ERR_WORKSPACE;
    float* ws = static_cast<float*>(workspace);
    cudaStream_t st = as_stream(stream);
    Gather g1, g2, g3;
    bool khwc1 = false, khwc = false;
    if (!conv_gather(&g1, channels, 8, 8, 4, 20, 20, obs_strides_host[0], obs_strides_host[1], obs_strides_host[2], obs_strides_host[3], &khwc1))
        return CLIPPPO_ERR_UNSUPPORTED;
    conv_gather(&g2, 32, 4, 4, 2, 9, 9, 400LL * 32, 1, 20LL * 32, 32, &khwc);
    conv_gather(&g3, 64, 3, 3, 1, 7, 7, 81LL * 64, 1, 9LL * 64, 64, &khwc);
    float* dz1 = ws + p.dz1; float* dz2 = ws + p.dz2; float* dz3 = ws + p.dz3;
    // ---- fc: dZ = dH * (H > 0) lives in dz1 for a moment ----
    float* dzf = dz1;
    relu_bwd_kernel<<<grid_for(mb * 512LL), 256, 0, st>>>(grad_hidden, hidden, mb * 512LL, dzf);
    CLIPPPO_CHECK_LAUNCH();
    POL_TRY(launch_tn(dzf, ws + p.act3, nullptr, ws + p.gtmp, ws + p.part, mb, 512, p.KF, 1.0f, st));     // dWfc in (h, w, c) column order
    permute_0231_kernel<<<grid_for(512LL * p.KF), 256, 0, st>>>(ws + p.gtmp, gwfc, 512, 64, 7, 7, 1);
    CLIPPPO_CHECK_LAUNCH();
    POL_TRY(launch_colsum(dzf, mb, 512, gbfc, ws + p.part, st));
    {   // dZ3 = (dZ Wfc) * (act3 > 0), [M3, 64] = [mb, 3136]
        dim3 tg((p.KF + 31) / 32, (512 + 31) / 32);
        transpose_kernel<<<tg, 256, 0, st>>>(ws + p.wfp, ws + p.wT, 512, p.KF);
        CLIPPPO_CHECK_LAUNCH();
        POL_TRY(launch_nt_plain<EPI_MASK>(dzf, ws + p.wT, dz3, mb, p.KF, 512, nullptr, ws + p.act3, st));
    }
    // ---- conv3: weight gradient over the gathered act2; input gradient = 3 x 3 transposed convolution of dZ3 ----
    POL_TRY(launch_tn(dz3, ws + p.act2, &g3, ws + p.gtmp, ws + p.part, p.M3, 64, p.K3, 1.0f, st));
    permute_0231_kernel<<<grid_for(64LL * p.K3), 256, 0, st>>>(ws + p.gtmp, gw3, 64, 64, 3, 3, 1);
    CLIPPPO_CHECK_LAUNCH();
    POL_TRY(launch_colsum(dz3, p.M3, 64, gb3, ws + p.part, st));
    {
        TapList taps; taps.n = 9;
        Gather gd = {};
        gd.NI = 9; gd.NJ = 9; gd.sn = 49LL * 64; gd.si = 7 * 64; gd.sj = 64; gd.VI = 7; gd.VJ = 7;
        for (int t = 0; t < 9; ++t) {
            taps.kh[t] = t / 3; taps.kw[t] = t % 3;
            gd.ta[t] = static_cast<signed char>(t / 3); gd.tb[t] = static_cast<signed char>(t % 3);
            for (int q4 = 0; q4 < 16; ++q4) {                           // 64 output channels = 16 float4 groups per tap
                gd.koff[t * 16 + q4] = -(t / 3) * 7 * 64 - (t % 3) * 64 + 4 * q4;
                gd.ktap[t * 16 + q4] = static_cast<unsigned char>(t);
            }
        }
        prep_dgrad_w_kernel<<<grid_for(64 * 9 * 64), 256, 0, st>>>(w3, 64, 64, 3, 3, taps, ws + p.wd);
        CLIPPPO_CHECK_LAUNCH();
        const RowOut ro{9, 9, 81LL * 64, 9 * 64, 64, 0};
        POL_TRY(launch_nt_gather<EPI_MASK>(dz3, gd, ws + p.wd, dz2, ro, p.M2, 64, 576, 1.0f, nullptr, ws + p.act2, st));
    }
    // ---- conv2: weight gradient over the gathered act1; input gradient per parity class of the stride-2 grid ----
    POL_TRY(launch_tn(dz2, ws + p.act1, &g2, ws + p.gtmp, ws + p.part, p.M2, 64, p.K2, 1.0f, st));
    permute_0231_kernel<<<grid_for(64LL * p.K2), 256, 0, st>>>(ws + p.gtmp, gw2, 64, 32, 4, 4, 1);
    CLIPPPO_CHECK_LAUNCH();
    POL_TRY(launch_colsum(dz2, p.M2, 64, gb2, ws + p.part, st));
    for (int cls = 0; cls < 4; ++cls) {
        const int ph = cls >> 1, pw = cls & 1;
        // input position (2 i + ph, 2 j + pw), i, j in [0, 10), receives taps kh = 2 a + ph, kw = 2 b + pw from output (i - a, j - b)
        TapList taps; taps.n = 4;
        Gather gd = {};
        gd.NI = 10; gd.NJ = 10; gd.sn = 81LL * 64; gd.si = 9 * 64; gd.sj = 64; gd.VI = 9; gd.VJ = 9;
        for (int t = 0; t < 4; ++t) {
            const int a = t >> 1, b = t & 1;
            taps.kh[t] = 2 * a + ph; taps.kw[t] = 2 * b + pw;
            gd.ta[t] = static_cast<signed char>(a); gd.tb[t] = static_cast<signed char>(b);
            for (int q4 = 0; q4 < 16; ++q4) {
                gd.koff[t * 16 + q4] = -a * 9 * 64 - b * 64 + 4 * q4;
                gd.ktap[t * 16 + q4] = static_cast<unsigned char>(t);
            }
        }
        float* wd = ws + p.wd + cls * 32 * 256;
        prep_dgrad_w_kernel<<<grid_for(32 * 4 * 64), 256, 0, st>>>(w2, 64, 32, 4, 4, taps, wd);
        CLIPPPO_CHECK_LAUNCH();
        const RowOut ro{10, 10, 400LL * 32, 2 * 20 * 32, 2 * 32, static_cast<long long>(ph) * 20 * 32 + pw * 32};
        POL_TRY(launch_nt_gather<EPI_MASK>(dz2, gd, wd, dz1, ro, static_cast<long long>(mb) * 100, 32, 256, 1.0f, nullptr, ws + p.act1, st));
    }
    // ---- conv1: weight gradient over the gathered observations (no input gradient) ----
    POL_TRY(launch_tn(dz1, obs, &g1, khwc1 ? ws + p.gtmp : gw1, ws + p.part, p.M1, 32, p.K1, in_scale, st));
    if (khwc1) {
        permute_0231_kernel<<<grid_for(32LL * p.K1), 256, 0, st>>>(ws + p.gtmp, gw1, 32, channels, 8, 8, 1);
        CLIPPPO_CHECK_LAUNCH();
    }
    POL_TRY(launch_colsum(dz1, p.M1, 32, gb1, ws + p.part, st));
    return CLIPPPO_OK;
}
