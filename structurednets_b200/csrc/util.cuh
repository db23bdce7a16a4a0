// Small shared device utilities.
#pragma once
#include "common.cuh"

namespace snb {

// out[c] += sum_r M[r][c]   (bias gradient = column sums of grad_y); coalesced along c, atomics across row slabs.
// CTA = 128 columns x 64 rows as 4 row groups of 16: four independent partial sums per thread keep loads in flight (a single
// serial sum over 256 rows was latency-bound: 28 us for an 8192 x 1000 matrix).
constexpr int COLSUM_ROWS = 64;
static __global__ void __launch_bounds__(512)
colsum_accumulate_kernel(const float* __restrict__ M, long ld, long rows, int cols, float* __restrict__ out) {
    __shared__ float part[4][128];
    const int cl = threadIdx.x & 127, rg = threadIdx.x >> 7;
    const int c = blockIdx.x * 128 + cl;
    const long r0 = (long)blockIdx.y * COLSUM_ROWS + rg * 16;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < cols) {
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            const long r = r0 + k;
            if (r < rows) s0 += __ldg(M + r * ld + c);
            if (r + 1 < rows) s1 += __ldg(M + (r + 1) * ld + c);
            if (r + 2 < rows) s2 += __ldg(M + (r + 2) * ld + c);
            if (r + 3 < rows) s3 += __ldg(M + (r + 3) * ld + c);
        }
    }
    part[rg][cl] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (rg == 0 && c < cols) atomicAdd(out + c, (part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl]));
}

inline int colsum_accumulate(const float* M, long ld, long rows, int cols, float* out, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid(ceil_div(cols, 128), (unsigned)((rows + COLSUM_ROWS - 1) / COLSUM_ROWS));
    SN_LAUNCH("colsum_accumulate_kernel", stream, colsum_accumulate_kernel<<<grid, 512, 0, stream>>>(M, ld, rows, cols, out));
    return 0;
}

}  // namespace snb
