// Launch counting / GEMM event spans (prof.cu).
#pragma once
#include <cuda_runtime.h>

namespace clipppo {
void prof_count_launch(int n = 1);
bool prof_timing_enabled();
void prof_span_begin(cudaStream_t s, double flops, long long tag, void** token);
void prof_span_end(cudaStream_t s, void* token);
}  // namespace clipppo
