// Micro-benchmark: issue rate of the legacy tensor path (mma.sync.m16n8k16 bf16 -> fp32) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_peak mma_sync_peak.cu && ./mma_sync_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ACC>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float c[ACC][4];
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 0x3f803f80u, a3 = 0x3f803f80u, b0 = 0x3f803f80u, b1 = blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

template <int ACC>
void run(int blocks_per_sm, int threads) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 4096, grid = 148 * blocks_per_sm;
    k<ACC><<<grid, threads>>>(d, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ACC><<<grid, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * 16 * ACC * iters * (threads / 32) * grid;
    printf("acc=%2d  %d CTAs/SM x %3d threads: %8.3f ms  %7.1f TFLOP/s\n", ACC, blocks_per_sm, threads, ms, flops / ms / 1e9);
    cudaFree(d);
}

int main() {
    run<4>(1, 128); run<8>(1, 128); run<8>(2, 128); run<8>(2, 256); run<16>(2, 256); run<8>(4, 256); run<16>(1, 128);
    return 0;
}
