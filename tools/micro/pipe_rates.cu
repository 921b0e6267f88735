// Not a test: issue-rate probe of the instructions the attention softmax leans on (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
// Prints thread-instructions per clock per SM for each op (148 CTAs x 1024 threads, 8 independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int OP>
__global__ void __launch_bounds__(1024) probe(float* out, int iters, long long* clocks) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    const float b = out[0], c = out[1];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fmaxf(a[i], b);                                     // FMNMX
            if (OP == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));    // FMNMX3
            if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));   // MUFU.EX2
            if (OP == 3) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b)); a[i] = __uint_as_float(r); }   // F2FP
            if (OP == 4) a[i] = fmaf(a[i], b, c);                                   // FFMA
            if (OP == 5) a[i] = a[i] + b;                                           // FADD
            if (OP == 6) asm volatile("mul.rn.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(__float_as_uint(b)));   // HMUL2.BF16
            if (OP == 7) { unsigned r; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(u[i])); u[i] = r; }   // MUFU.EX2 packed
            if (OP == 8) u[i] = (u[i] << 16) ^ (unsigned)it;                        // shift / LOP3
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    if (s == 123.456f) out[2] = s;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, float* out, long long* clocks) {
    const int iters = 2000;
    probe<OP><<<148, 1024>>>(out, iters, clocks);
    cudaDeviceSynchronize();
    probe<OP><<<148, 1024>>>(out, iters, clocks);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clocks, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-28s %7.1f thread-instr / clk / SM   (%.2f clk per warp-instr per SMSP)\n", name, 1024.0 * 8 * iters / avg,
           avg / (8.0 * iters * 8));
}

int main() {
    float* out; long long* clocks;
    cudaMalloc(&out, 64); cudaMemset(out, 0, 64);
    cudaMalloc(&clocks, 148 * 8);
    run<0>("FMNMX  (max.f32 a,b)", out, clocks);
    run<1>("FMNMX3 (max.f32 a,b,c)", out, clocks);
    run<2>("MUFU.EX2 f32", out, clocks);
    run<3>("F2FP.BF16.F32.PACK_AB", out, clocks);
    run<4>("FFMA", out, clocks);
    run<5>("FADD", out, clocks);
    run<6>("HMUL2.BF16", out, clocks);
    run<7>("MUFU.EX2 bf16x2", out, clocks);
    run<8>("SHL+LOP3", out, clocks);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
