// wgrad_tc.cuh -- interface of the split-K tcgen05 weight-gradient GEMM (wgrad_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fqss {
namespace tcw {

// bytes of the partial-sum buffer `run` needs for this shape
size_t part_bytes(int B, int M, int O, int I);

// dWq[O][I] = (da * sum_{b,m} dY[b,o,m] X[b,i,m]) / dws[o] + min_a * db[o]
//   dY [B][O][ld], X [B][I][ld] bf16; amin/amax: quantiser of X's tensor (NULL: da = 1, min = 0); db: fp64 row sums of the
//   UNSCALED output gradient.
int run(const void* dY, const void* X, int B, int M, int64_t ld, int O, int I, float* part, size_t part_cap, const float* amin,
        const float* amax, const float* dws, const double* db, float* dWq, cudaStream_t s);

}  // namespace tcw
}  // namespace fqss
