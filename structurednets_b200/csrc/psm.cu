// Product-of-sparse-matrices layer:  y = S_0 (S_1 ( ... (S_{n-1} x))) + b.
// Replaces PSMLayer.forward / forward_sparse (reference layers/psm_layer.py:36-60) -- with the factor
// order the reference *intends* (W = S_0 S_1 ... S_{n-1}, approximators/psm_approximator.py:96-105); the
// literal loop order of psm_layer.py:52-54 is only correct for <= 2 factors (SURVEY.md finding F2), where
// both coincide -- and the autograd backward (sparse gradients on the COO pattern).
//
// The chain is applied per tile of NS samples with every intermediate activation kept in shared memory,
// feature-major [feature][NS]: only x and y touch HBM.  Each factor is walked in CSR order (one thread per
// output row, gathering NS-wide input rows from shared memory); values are read through a permutation from
// the parameter's own (uncoalesced) COO value array, so nothing is re-sorted per step.  The backward
// recomputes the intermediates, forms dS_k[j] = sum_s g[row_j][s] * in_k[col_j][s] per tile (atomic add to
// the COO-ordered gradient values) and pushes g through S_k^T in CSC order.
#include "common.cuh"
#include "util.cuh"

namespace {

constexpr int PSM_MAX_FACTORS = 8;
constexpr int PSM_THREADS = 512;

struct Factor {
    int rows, cols, nnz, pad;
    const int* rowptr;   // rows + 1
    const int* colidx;   // nnz, CSR order
    const int* perm;     // nnz, CSR position -> index into the COO value array
    const int* cscptr;   // cols + 1
    const int* rowidx;   // nnz, CSC order
    const int* permc;    // nnz, CSC position -> index into the COO value array
    const float* vals;   // COO value array (parameter storage)
    float* gvals;        // COO-ordered gradient values (backward only)
};

struct Chain {
    int nf, in_dim, out_dim, maxdim;
    Factor f[PSM_MAX_FACTORS];
};

template <int NS>
__device__ __forceinline__ void load_x(const float* __restrict__ x, long ldx, long t0, long B, int dim, float* dst) {
    for (int s = 0; s < NS; ++s) {
        const bool ok = t0 + s < B;
        const float* row = x + (size_t)(t0 + s) * ldx;
        for (int f = threadIdx.x; f < dim; f += PSM_THREADS) dst[f * NS + s] = ok ? __ldg(row + f) : 0.f;
    }
}

// out[r][:] = sum_j vals[perm[j]] * in[colidx[j]][:]   for all rows r of the factor
template <int NS>
__device__ __forceinline__ void spmm_csr(const Factor& F, const float* in, float* out) {
    for (int r = threadIdx.x; r < F.rows; r += PSM_THREADS) {
        float acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.f;
        const int j1 = __ldg(F.rowptr + r + 1);
        for (int j = __ldg(F.rowptr + r); j < j1; ++j) {
            const float v = __ldg(F.vals + __ldg(F.perm + j));
            const float* iv = in + (size_t)__ldg(F.colidx + j) * NS;
            if constexpr (NS == 4) {
                const float4 t = *reinterpret_cast<const float4*>(iv);
                acc[0] = fmaf(v, t.x, acc[0]); acc[1] = fmaf(v, t.y, acc[1]);
                acc[2] = fmaf(v, t.z, acc[2]); acc[3] = fmaf(v, t.w, acc[3]);
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) acc[s] = fmaf(v, iv[s], acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) out[(size_t)r * NS + s] = acc[s];
    }
}

template <int NS>
__global__ void __launch_bounds__(PSM_THREADS)
psm_fwd_kernel(Chain ch, const float* __restrict__ x, long ldx, float* __restrict__ y, long ldy, const float* __restrict__ bias, long B) {
    extern __shared__ __align__(16) float smem[];
    float* a = smem;
    float* b = smem + (size_t)ch.maxdim * NS;
    const long t0 = (long)blockIdx.x * NS;
    load_x<NS>(x, ldx, t0, B, ch.in_dim, a);
    __syncthreads();
    for (int k = ch.nf - 1; k >= 0; --k) {
        spmm_csr<NS>(ch.f[k], a, b);
        __syncthreads();
        float* t = a; a = b; b = t;
    }
    for (int s = 0; s < NS; ++s) {
        if (t0 + s >= B) break;
        float* row = y + (size_t)(t0 + s) * ldy;
        for (int o = threadIdx.x; o < ch.out_dim; o += PSM_THREADS) row[o] = a[(size_t)o * NS + s] + (bias ? __ldg(bias + o) : 0.f);
    }
}

template <int NS>
__global__ void __launch_bounds__(PSM_THREADS)
psm_bwd_kernel(Chain ch, const float* __restrict__ x, long ldx, const float* __restrict__ gy, long ldgy, long B) {
    extern __shared__ __align__(16) float smem[];
    const size_t vec = (size_t)ch.maxdim * NS;
    // act[k] = input of factor k (k = nf-1 is x); g0/g1 = gradient ping-pong
    float* act = smem;                        // nf vectors
    float* g0 = smem + (size_t)ch.nf * vec;
    float* g1 = g0 + vec;
    const long t0 = (long)blockIdx.x * NS;
    load_x<NS>(x, ldx, t0, B, ch.in_dim, act + (size_t)(ch.nf - 1) * vec);
    load_x<NS>(gy, ldgy, t0, B, ch.out_dim, g0);
    __syncthreads();
    for (int k = ch.nf - 1; k >= 1; --k) {   // recompute the inputs of factors nf-2 .. 0
        spmm_csr<NS>(ch.f[k], act + (size_t)k * vec, act + (size_t)(k - 1) * vec);
        __syncthreads();
    }
    for (int k = 0; k < ch.nf; ++k) {
        const Factor& F = ch.f[k];
        const float* in = act + (size_t)k * vec;
        // dS_k[j] += sum_s g[row][s] * in[col][s]
        for (int r = threadIdx.x; r < F.rows; r += PSM_THREADS) {
            float gr[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) gr[s] = g0[(size_t)r * NS + s];
            const int j1 = __ldg(F.rowptr + r + 1);
            for (int j = __ldg(F.rowptr + r); j < j1; ++j) {
                const float* iv = in + (size_t)__ldg(F.colidx + j) * NS;
                float d = 0.f;
#pragma unroll
                for (int s = 0; s < NS; ++s) d = fmaf(gr[s], iv[s], d);
                atomicAdd(F.gvals + __ldg(F.perm + j), d);
            }
        }
        // g_next[c][:] = sum_j vals[permc[j]] * g[rowidx[j]][:]   (S_k^T g), not needed after the last factor
        if (k + 1 < ch.nf) {
            for (int c = threadIdx.x; c < F.cols; c += PSM_THREADS) {
                float acc[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) acc[s] = 0.f;
                const int j1 = __ldg(F.cscptr + c + 1);
                for (int j = __ldg(F.cscptr + c); j < j1; ++j) {
                    const float v = __ldg(F.vals + __ldg(F.permc + j));
                    const float* gv = g0 + (size_t)__ldg(F.rowidx + j) * NS;
#pragma unroll
                    for (int s = 0; s < NS; ++s) acc[s] = fmaf(v, gv[s], acc[s]);
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) g1[(size_t)c * NS + s] = acc[s];
            }
        }
        __syncthreads();
        float* t = g0; g0 = g1; g1 = t;
    }
}

}  // namespace

extern "C" {

static int make_chain(const sn_psm_factor* f, int nf, int in_dim, int out_dim, Chain* ch) {
    SN_CHECK_ARG(f != nullptr && nf >= 1 && nf <= PSM_MAX_FACTORS, "psm: need 1..%d factors", PSM_MAX_FACTORS);
    ch->nf = nf; ch->in_dim = in_dim; ch->out_dim = out_dim; ch->maxdim = in_dim > out_dim ? in_dim : out_dim;
    for (int k = 0; k < nf; ++k) {
        SN_CHECK_ARG(f[k].rowptr && f[k].cscptr && (f[k].nnz == 0 || (f[k].colidx && f[k].perm && f[k].rowidx && f[k].permc && f[k].vals)),
                     "psm: factor %d has NULL arrays", k);
        ch->f[k] = Factor{f[k].rows, f[k].cols, f[k].nnz, 0, f[k].rowptr, f[k].colidx, f[k].perm, f[k].cscptr, f[k].rowidx, f[k].permc, f[k].vals, f[k].grad_vals};
        if (f[k].rows > ch->maxdim) ch->maxdim = f[k].rows;
        if (f[k].cols > ch->maxdim) ch->maxdim = f[k].cols;
        if (k > 0) SN_CHECK_ARG(f[k - 1].cols == f[k].rows, "psm: factor %d cols (%d) != factor %d rows (%d)", k - 1, f[k - 1].cols, k, f[k].rows);
    }
    SN_CHECK_ARG(f[0].rows == out_dim && f[nf - 1].cols == in_dim, "psm: chain maps %d -> %d, layer is %d -> %d", f[nf - 1].cols, f[0].rows, in_dim, out_dim);
    return 0;
}

int sn_psm_forward(const sn_psm_factor* factors_host, int nf, const float* x, int64_t ldx, float* y, int64_t ldy, const float* bias,
                   int64_t B, int in_dim, int out_dim, sn_stream_t stream) {
    Chain ch;
    if (int rc = make_chain(factors_host, nf, in_dim, out_dim, &ch)) return rc;
    SN_CHECK_ARG(x && y, "psm_forward: NULL buffer");
    if (B <= 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
#define PSM_FWD(NS)                                                                                                         \
    {                                                                                                                       \
        size_t smem = (size_t)2 * ch.maxdim * NS * sizeof(float);                                                           \
        if (smem <= 227 * 1024) {                                                                                           \
            SN_CHECK_CUDA(cudaFuncSetAttribute(psm_fwd_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            psm_fwd_kernel<NS><<<(unsigned)((B + NS - 1) / NS), PSM_THREADS, smem, st>>>(ch, x, ldx, y, ldy, bias, B);      \
            SN_CHECK_LAUNCH("psm_fwd_kernel");                                                                              \
            return 0;                                                                                                       \
        }                                                                                                                   \
    }
    PSM_FWD(4)
    PSM_FWD(2)
    PSM_FWD(1)
#undef PSM_FWD
    snb::set_error("psm_forward: factor dimension %d does not fit in shared memory", ch.maxdim);
    return 1;
}

int sn_psm_backward(const sn_psm_factor* factors_host, int nf, const float* x, int64_t ldx, const float* grad_y, int64_t ldgy,
                    float* grad_bias, int64_t B, int in_dim, int out_dim, sn_stream_t stream) {
    Chain ch;
    if (int rc = make_chain(factors_host, nf, in_dim, out_dim, &ch)) return rc;
    SN_CHECK_ARG(x && grad_y, "psm_backward: NULL buffer");
    for (int k = 0; k < nf; ++k) SN_CHECK_ARG(ch.f[k].gvals != nullptr || ch.f[k].nnz == 0, "psm_backward: factor %d has no gradient buffer", k);
    if (B <= 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    if (grad_bias)
        if (int rc = snb::colsum_accumulate(grad_y, ldgy, B, out_dim, grad_bias, st)) return rc;
#define PSM_BWD(NS)                                                                                                         \
    {                                                                                                                       \
        size_t smem = (size_t)(ch.nf + 2) * ch.maxdim * NS * sizeof(float);                                                 \
        if (smem <= 227 * 1024) {                                                                                           \
            SN_CHECK_CUDA(cudaFuncSetAttribute(psm_bwd_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            psm_bwd_kernel<NS><<<(unsigned)((B + NS - 1) / NS), PSM_THREADS, smem, st>>>(ch, x, ldx, grad_y, ldgy, B);      \
            SN_CHECK_LAUNCH("psm_bwd_kernel");                                                                              \
            return 0;                                                                                                       \
        }                                                                                                                   \
    }
    PSM_BWD(4)
    PSM_BWD(2)
    PSM_BWD(1)
#undef PSM_BWD
    snb::set_error("psm_backward: %d factors of dimension %d do not fit in shared memory", ch.nf, ch.maxdim);
    return 1;
}

}  // extern "C"
