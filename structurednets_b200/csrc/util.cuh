// Small shared device utilities.
#pragma once
#include "common.cuh"

namespace snb {

// out[c] += sum_r M[r][c]   (bias gradient = column sums of grad_y); coalesced along c, atomics across row slabs
static __global__ void colsum_accumulate_kernel(const float* __restrict__ M, long ld, long rows, int cols, float* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    long r0 = (long)blockIdx.y * 256;
    long r1 = r0 + 256 < rows ? r0 + 256 : rows;
    float s = 0.f;
    for (long r = r0; r < r1; ++r) s += __ldg(M + r * ld + c);
    atomicAdd(out + c, s);
}

inline int colsum_accumulate(const float* M, long ld, long rows, int cols, float* out, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid(ceil_div(cols, 128), (unsigned)((rows + 255) / 256));
    SN_LAUNCH("colsum_accumulate_kernel", stream, colsum_accumulate_kernel<<<grid, 128, 0, stream>>>(M, ld, rows, cols, out));
    return 0;
}

}  // namespace snb
