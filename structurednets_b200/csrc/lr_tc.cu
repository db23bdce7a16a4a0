// Low-rank layer, bf16 tensor-core path (BASELINE config C2):  y = (x R^T) L^T + b  with bf16 operands, fp32
// accumulation in TMEM and fp32 master parameters / gradients.  Replaces LRLayer.forward (reference
// layers/lr_layer.py:38-46) and its autograd backward for bfloat16 features.
//
//   forward :  h  = x R^T            (B x in)(r x in)^T      tcgen05, K-major operands, bf16 out (kept for backward)
//              y  = h L^T + b        (B x r)(out x r)^T      tcgen05, bf16 out
//   backward:  gh = gy L             (B x out)(r x out)^T    tcgen05 on the pre-transposed L
//              dL += gy^T h          (out x B)(r x B)^T      tcgen05 on transposed copies, split-K over the batch, fp32 atomics
//              dR += gh^T x          (r x B)(in x B)^T       same
//              db += colsum(gy)
// The batch-reduction GEMMs run on bf16 transposes made by a tiled transpose kernel (ranks that are not a multiple of 64);
// otherwise the activations are consumed in their natural [sample][feature] layout through MN-major UMMA descriptors.
#include "gemm_tc.cuh"

namespace {

using bf16 = __nv_bfloat16;

// fp32 parameters -> bf16 copies: L (out x r), L^T (r x out), R (r x in)
__global__ void lr_cast_params_kernel(const float* __restrict__ left, const float* __restrict__ right, bf16* __restrict__ Lb, bf16* __restrict__ Ltb,
                                      bf16* __restrict__ Rb, int out_dim, int rank, int in_dim, int ldlt) {
    const size_t nl = (size_t)out_dim * rank, nr = (size_t)rank * in_dim;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nl + nr; e += (size_t)gridDim.x * blockDim.x) {
        if (e < nl) {
            const int o = (int)(e / rank), k = (int)(e - (size_t)o * rank);
            const bf16 v = __float2bfloat16(left[e]);
            Lb[e] = v;
            Ltb[(size_t)k * ldlt + o] = v;
        } else {
            Rb[e - nl] = __float2bfloat16(right[e - nl]);
        }
    }
}

// dst[c][r] = src[r][c]  (bf16, 32x32 tiles through shared memory)
__global__ void transpose_bf16_kernel(const bf16* __restrict__ src, long lds, bf16* __restrict__ dst, long ldd, long rows, int cols) {
    __shared__ bf16 tile[32][33];
    const long r0 = (long)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long r = r0 + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long r = r0 + threadIdx.x;
        if (c < cols && r < rows) dst[(size_t)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// out[c] += sum_r M[r][c]: a thread owns a pair of columns (one 4-byte load per row, coalesced along the row), 64-row slabs
__global__ void colsum_bf16_kernel(const bf16* __restrict__ M, long ld, long rows, int cols, float* __restrict__ out) {
    const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= cols) return;
    const long r0 = (long)blockIdx.y * 64;
    const long r1 = r0 + 64 < rows ? r0 + 64 : rows;
    float s0 = 0.f, s1 = 0.f;
    if (c + 1 < cols && (ld & 1) == 0 && ((reinterpret_cast<uintptr_t>(M) & 3) == 0)) {
#pragma unroll 8
        for (long r = r0; r < r1; ++r) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(M + r * ld + c);
            s0 += __bfloat162float(v.x);
            s1 += __bfloat162float(v.y);
        }
    } else {
        for (long r = r0; r < r1; ++r) {
            s0 += __bfloat162float(M[r * ld + c]);
            if (c + 1 < cols) s1 += __bfloat162float(M[r * ld + c + 1]);
        }
    }
    atomicAdd(out + c, s0);
    if (c + 1 < cols) atomicAdd(out + c + 1, s1);
}

int transpose_bf16(const bf16* src, long lds, bf16* dst, long ldd, long rows, int cols, cudaStream_t st) {
    dim3 grid(snb::ceil_div(cols, 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    SN_LAUNCH("transpose_bf16_kernel", st, transpose_bf16_kernel<<<grid, block, 0, st>>>(src, lds, dst, ldd, rows, cols));
    return 0;
}

inline int split_for(int tiles, long K) {
    int want = snb::ceil_div(148, tiles > 0 ? tiles : 1);
    long maxs = (K + 511) / 512;   // at least 512 of K per slice
    if (want > maxs) want = (int)maxs;
    return want < 1 ? 1 : want;
}

}  // namespace

extern "C" {

// bf16 copies of the parameters; lt_ld = row pitch (elements) of left_t_bf16 (>= out_dim, multiple of 8)
int sn_lr_tc_cast_params(const float* left, const float* right, void* left_bf16, void* left_t_bf16, int64_t lt_ld, void* right_bf16,
                         int in_dim, int out_dim, int rank, sn_stream_t stream) {
    SN_CHECK_ARG(left && right && left_bf16 && left_t_bf16 && right_bf16, "lr_tc_cast_params: NULL buffer");
    SN_LAUNCH("lr_cast_params_kernel", snb::as_stream(stream), lr_cast_params_kernel<<<296, 256, 0, snb::as_stream(stream)>>>(left, right, (bf16*)left_bf16, (bf16*)left_t_bf16, (bf16*)right_bf16, out_dim, rank,
                                                                 in_dim, (int)lt_ld));
    return 0;
}

// x (B x in) bf16, hidden (B x rank) bf16 [written], y (B x out) bf16 [written]; all row pitches multiples of 8 elements
int sn_lr_tc_forward(const void* x, int64_t ldx, const void* left_bf16, const void* right_bf16, const float* bias, void* hidden, void* y,
                     int64_t ldy, int64_t B, int in_dim, int out_dim, int rank, sn_stream_t stream) {
    SN_CHECK_ARG(x && left_bf16 && right_bf16 && hidden && y, "lr_tc_forward: NULL buffer");
    SN_CHECK_ARG(rank % 8 == 0 && rank >= 8, "lr_tc_forward: the tensor-core path needs rank %% 8 == 0 (got %d)", rank);
    if (B <= 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    using namespace snb::tc;
    if (int rc = gemm_bf16_tc<64, STORE_BF16>((int)B, rank, in_dim, x, ldx, right_bf16, in_dim, hidden, rank, nullptr, 1.f, 1, st)) return rc;
    return gemm_bf16_tc<128, STORE_BF16>((int)B, out_dim, rank, hidden, rank, left_bf16, rank, y, ldy, bias, 1.f, 1, st);
}

// grad_left (out x rank), grad_right (rank x in), grad_bias (out) are fp32 and accumulated.  Scratch (bf16): ghid (B x rank),
// gyt (out x ldt), ht (rank x ldt), ght (rank x ldt), xt (in x ldt) with ldt = B rounded up to a multiple of 8.
int sn_lr_tc_backward(const void* x, int64_t ldx, const void* grad_y, int64_t ldgy, const void* left_t_bf16, int64_t lt_ld, const void* hidden,
                      void* ghid, void* gyt, void* ht, void* ght, void* xt, int64_t ldt, float* grad_left, float* grad_right, float* grad_bias,
                      int64_t B, int in_dim, int out_dim, int rank, sn_stream_t stream) {
    SN_CHECK_ARG(x && grad_y && left_t_bf16 && hidden && ghid && gyt && ht && ght && xt, "lr_tc_backward: NULL buffer");
    SN_CHECK_ARG(ldt >= B && ldt % 8 == 0, "lr_tc_backward: ldt must be B rounded up to a multiple of 8");
    if (B <= 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    using namespace snb::tc;
    // gh = gy L : A = gy (B x out), B operand = L^T (rank x out), K = out
    if (int rc = gemm_bf16_tc<64, STORE_BF16>((int)B, rank, out_dim, grad_y, ldgy, left_t_bf16, lt_ld, ghid, rank, nullptr, 1.f, 1, st)) return rc;
    if (grad_bias) {
        dim3 grid(snb::ceil_div(out_dim, 256), (unsigned)((B + 63) / 64));
        SN_LAUNCH("colsum_bf16_kernel", st, colsum_bf16_kernel<<<grid, 128, 0, st>>>((const bf16*)grad_y, ldgy, B, out_dim, grad_bias));
    }
    const bool mn_ok = (rank % 64 == 0);   // MN-major operands are fetched in 64-element blocks along M / N
    if (grad_left) {   // dL += gy^T h
        const int tiles = snb::ceil_div(out_dim, 128) * snb::ceil_div(rank, 128);
        if (mn_ok) {   // operands read in their natural [sample][feature] layout through MN-major UMMA descriptors
            if (int rc = gemm_bf16_tc<128, ATOMIC_F32, true>(out_dim, rank, (int)B, grad_y, ldgy, hidden, rank, grad_left, rank, nullptr, 1.f, split_for(tiles, B), st)) return rc;
        } else {
            if (int rc = transpose_bf16((const bf16*)grad_y, ldgy, (bf16*)gyt, ldt, B, out_dim, st)) return rc;
            if (int rc = transpose_bf16((const bf16*)hidden, rank, (bf16*)ht, ldt, B, rank, st)) return rc;
            if (int rc = gemm_bf16_tc<128, ATOMIC_F32>(out_dim, rank, (int)B, gyt, ldt, ht, ldt, grad_left, rank, nullptr, 1.f, split_for(tiles, B), st)) return rc;
        }
    }
    if (grad_right) {  // dR += gh^T x
        const int tiles = snb::ceil_div(rank, 128) * snb::ceil_div(in_dim, 128);
        if (mn_ok) {
            if (int rc = gemm_bf16_tc<128, ATOMIC_F32, true>(rank, in_dim, (int)B, ghid, rank, x, ldx, grad_right, in_dim, nullptr, 1.f, split_for(tiles, B), st)) return rc;
        } else {
            if (int rc = transpose_bf16((const bf16*)ghid, rank, (bf16*)ght, ldt, B, rank, st)) return rc;
            if (int rc = transpose_bf16((const bf16*)x, ldx, (bf16*)xt, ldt, B, in_dim, st)) return rc;
            if (int rc = gemm_bf16_tc<128, ATOMIC_F32>(rank, in_dim, (int)B, ght, ldt, xt, ldt, grad_right, in_dim, nullptr, 1.f, split_for(tiles, B), st)) return rc;
        }
    }
    return 0;
}

}  // extern "C"
