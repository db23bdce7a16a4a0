// Low-displacement-rank (tridiagonal-plus-corners operators) and Toeplitz-like layers.
//
// LDR replaces LDRLayer.forward = build_weight_matrix_torch + matmul (reference layers/ldr_layer.py:44-61,
// approximators/ldr_approximator.py:29-39) and its autograd backward.  The reference materialises
//     W = sum_i K(A, g_i) K(B^T, h_i)^T ,   K(M, v) = [v, Mv, ..., M^{n-1} v]
// with matrix_power per column (O(r n^4 log n)).  Since K(A,g_i) K(B^T,h_i)^T = sum_j A^j g_i h_i^T B^j,
//     W = sum_{j=0}^{n-1} T_j ,   T_0 = G H^T ,   T_{j+1} = A T_j B ,
// and A, B are tridiagonal + two corners, so each step is two O(n^2) banded passes in float64 (the
// parameters are float64).  The series is summed until a term is numerically zero relative to
// W (or j = n-1): for Glorot-scale operators it converges after a few dozen terms instead of n.
// Backward (exact gradient of the summed series, J = last term):
//     S_0 = dW, S_{l+1} = A^T S_l B^T ;  dM = sum_l S_l ;  dG = dM H ;  dH = dM^T G
//     dA = sum_{l<J} pattern( S_l (P_{J-1-l} B)^T ) ;  dB = sum_{l<J} pattern( P_{J-1-l}^T (A^T S_l) )
// with the prefix sums P_k = sum_{m<=k} T_m saved by the forward.
//
// Toeplitz-like replaces TLLayer.forward (reference layers/tl_layer.py:11-18,51-68):
//     W = 1/2 sum_j Krylov(Z_1, G[:,j]) Krylov(Z_-1, flip(H[j,:]))
// built as one (n x rn)(rn x n) fp32 GEMM from the two explicitly formed f-circulant Krylov stacks.
#include "gemm_f32.cuh"
#include "util.cuh"

namespace {

struct Band {           // tridiagonal-plus-corners n x n operator M
    const double* lo;   // lo[p] = M[p][p-1]   (lo[0] unused = 0)
    const double* di;   // di[p] = M[p][p]
    const double* up;   // up[p] = M[p][p+1]   (up[n-1] unused = 0)
    const double* cn;   // cn[0] = M[0][n-1], cn[1] = M[n-1][0]  (0 when n <= 2)
};

// slot map entry: 0..n-1 -> lo, n..2n-1 -> di, 2n..3n-1 -> up, 3n -> corner(0,n-1), 3n+1 -> corner(n-1,0)
__global__ void band_gather_kernel(const double* __restrict__ vals, const int* __restrict__ slot, int nnz, double* __restrict__ band) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) atomicAdd(band + slot[e], vals[e]);
}
__global__ void band_scatter_grad_kernel(const double* __restrict__ gband, const int* __restrict__ slot, int nnz, double* __restrict__ gvals) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) gvals[e] += gband[slot[e]];
}
// band of M^T from band of M
__global__ void band_transpose_kernel(const double* __restrict__ b, double* __restrict__ t, int n) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) {
        t[p] = p > 0 ? b[2 * n + p - 1] : 0.0;          // lo'[p] = up[p-1]
        t[n + p] = b[n + p];
        t[2 * n + p] = p + 1 < n ? b[p + 1] : 0.0;      // up'[p] = lo[p+1]
    }
    if (p == 0) { t[3 * n] = b[3 * n + 1]; t[3 * n + 1] = b[3 * n]; }
}

__device__ __forceinline__ Band as_band(const double* b, int n) { return Band{b, b + n, b + 2 * n, b + 3 * n}; }

// out = M T   (rows mix)
__global__ void band_left_kernel(const double* __restrict__ band, const double* __restrict__ T, double* __restrict__ out, int n) {
    int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (q >= n) return;
    Band M = as_band(band, n);
    double v = M.di[p] * T[(size_t)p * n + q];
    if (p > 0) v += M.lo[p] * T[(size_t)(p - 1) * n + q];
    if (p + 1 < n) v += M.up[p] * T[(size_t)(p + 1) * n + q];
    if (n > 2) {
        if (p == 0) v += M.cn[0] * T[(size_t)(n - 1) * n + q];
        if (p == n - 1) v += M.cn[1] * T[q];
    }
    out[(size_t)p * n + q] = v;
}
// out = T M   (columns mix)
__global__ void band_right_kernel(const double* __restrict__ band, const double* __restrict__ T, double* __restrict__ out, int n) {
    int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (q >= n) return;
    Band M = as_band(band, n);
    const double* row = T + (size_t)p * n;
    double v = row[q] * M.di[q];
    if (q > 0) v += row[q - 1] * M.up[q - 1];
    if (q + 1 < n) v += row[q + 1] * M.lo[q + 1];
    if (n > 2) {
        if (q == n - 1) v += row[0] * M.cn[0];
        if (q == 0) v += row[n - 1] * M.cn[1];
    }
    out[(size_t)p * n + q] = v;
}
// acc += T ; optionally snapshot = acc ; block-wise max |T| -> atomicMax on the bit pattern (non-negative doubles order like uint64)
__global__ void accumulate_kernel(const double* __restrict__ T, double* __restrict__ acc, double* __restrict__ snapshot, size_t n2,
                                  unsigned long long* __restrict__ maxabs_bits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0.0;
    if (i < n2) {
        double t = T[i];
        double w = acc[i] + t;
        acc[i] = w;
        if (snapshot) snapshot[i] = w;
        a = fabs(t);
    }
    for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
    if ((threadIdx.x & 31) == 0 && a > 0.0) atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(a));
}
// C[n x n] = G H^T  (double)
__global__ void ght_kernel(const double* __restrict__ G, const double* __restrict__ H, double* __restrict__ C, int n, int r) {
    int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (q >= n) return;
    double v = 0.0;
    for (int i = 0; i < r; ++i) v += G[(size_t)p * r + i] * H[(size_t)q * r + i];
    C[(size_t)p * n + q] = v;
}
// dG[p][i] += sum_q dM[p][q] H[q][i] ; dH[q][i] += sum_p dM[p][q] G[p][i]
__global__ void dgh_kernel(const double* __restrict__ dM, const double* __restrict__ G, const double* __restrict__ H,
                           double* __restrict__ dG, double* __restrict__ dH, int n, int r) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (i >= r) return;
    double a = 0.0, b = 0.0;
    for (int q = 0; q < n; ++q) {
        a += dM[(size_t)p * n + q] * H[(size_t)q * r + i];
        b += dM[(size_t)q * n + p] * G[(size_t)q * r + i];
    }
    dG[(size_t)p * r + i] += a;
    dH[(size_t)p * r + i] += b;
}
__global__ void cast_f64_f32_kernel(const double* __restrict__ a, float* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}
__global__ void cast_f32_f64_kernel(const float* __restrict__ a, double* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
// gband(A) += pattern( S Q^T ):  dA[p][a] = sum_q S[p][q] Q[a][q]  for a in {p-1, p, p+1} and the corners; one warp per row p
__global__ void row_band_dots_kernel(const double* __restrict__ S, const double* __restrict__ Q, double* __restrict__ gband, int n) {
    int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n) return;
    const double* s = S + (size_t)p * n;
    const double* q0 = Q + (size_t)p * n;
    const double* qm = p > 0 ? q0 - n : nullptr;
    const double* qp = p + 1 < n ? q0 + n : nullptr;
    const double* qc = (n > 2 && p == 0) ? Q + (size_t)(n - 1) * n : ((n > 2 && p == n - 1) ? Q : nullptr);
    double a0 = 0, am = 0, ap = 0, ac = 0;
    for (int q = lane; q < n; q += 32) {
        double sv = s[q];
        a0 += sv * q0[q];
        if (qm) am += sv * qm[q];
        if (qp) ap += sv * qp[q];
        if (qc) ac += sv * qc[q];
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); am += __shfl_xor_sync(0xffffffffu, am, o);
        ap += __shfl_xor_sync(0xffffffffu, ap, o); ac += __shfl_xor_sync(0xffffffffu, ac, o);
    }
    if (lane == 0) {
        gband[n + p] += a0;
        if (qm) gband[p] += am;
        if (qp) gband[2 * n + p] += ap;
        if (qc) gband[3 * n + (p == 0 ? 0 : 1)] += ac;
    }
}
// gband(B) += pattern( P^T V ):  dB[b][q] = sum_p P[p][b] V[p][q] for b in {q-1,q,q+1} and corners; thread per column q
// (blockIdx.y splits the rows p; partial sums are combined with double atomics)
__global__ void col_band_dots_kernel(const double* __restrict__ P, const double* __restrict__ V, double* __restrict__ gband, int n) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double a0 = 0, am = 0, ap = 0, ac = 0;   // b = q, q-1, q+1, corner
    const bool hm = q > 0, hp = q + 1 < n, c_first = (n > 2 && q == n - 1), c_last = (n > 2 && q == 0);
    const int rows_per = (n + gridDim.y - 1) / gridDim.y;
    const int p_begin = blockIdx.y * rows_per, p_end = min(n, p_begin + rows_per);
    for (int p = p_begin; p < p_end; ++p) {
        const double* pr = P + (size_t)p * n;
        double v = V[(size_t)p * n + q];
        a0 += pr[q] * v;
        if (hm) am += pr[q - 1] * v;
        if (hp) ap += pr[q + 1] * v;
        if (c_first) ac += pr[0] * v;        // dB[0][n-1]
        if (c_last) ac += pr[n - 1] * v;     // dB[n-1][0]
    }
    atomicAdd(gband + n + q, a0);                       // di[q] = B[q][q]
    if (hm) atomicAdd(gband + 2 * n + q - 1, am);       // B[q-1][q] = up[q-1]
    if (hp) atomicAdd(gband + q + 1, ap);               // B[q+1][q] = lo[q+1]
    if (c_first) atomicAdd(gband + 3 * n, ac);
    if (c_last) atomicAdd(gband + 3 * n + 1, ac);
}

inline dim3 grid2(int n) { return dim3(snb::ceil_div(n, 128), n); }
inline unsigned grid1(size_t n) { return (unsigned)((n + 255) / 256); }

// ---- Toeplitz-like helpers ----------------------------------------------------------------------
// K1[p][j*n+i] = G[(p-i) mod n][j]                       (Krylov(Z_1, G[:,j]))
__global__ void tl_k1_kernel(const float* __restrict__ G, float* __restrict__ K1, int n, int r) {
    int c = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;   // c = j*n + i
    if (c >= n * r) return;
    int j = c / n, i = c - j * n;
    int m = p - i; if (m < 0) m += n;
    K1[(size_t)p * n * r + c] = G[(size_t)m * r + j];
}
// K2[j*n+i][q] = v_j[(i-q) mod n] * (i < q ? -1 : 1),  v_j[m] = H[j][n-1-m]     (Krylov(Z_-1, flip(H[j,:])))
__global__ void tl_k2_kernel(const float* __restrict__ H, float* __restrict__ K2, int n, int r) {
    int q = blockIdx.y * blockDim.x + threadIdx.x, row = blockIdx.x;  // row = j*n + i  (n*r rows: on grid.x, grid.y stops at 65 535)
    if (q >= n) return;
    int j = row / n, i = row - j * n;
    int m = i - q; if (m < 0) m += n;
    float v = H[(size_t)j * n + (n - 1 - m)];
    K2[(size_t)row * n + q] = i < q ? -v : v;
}
// dG[m][j] += sum_i dK1[(m+i) mod n][j*n+i]
__global__ void tl_dg_kernel(const float* __restrict__ dK1, float* __restrict__ dG, int n, int r) {
    int j = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (j >= r) return;
    float s = 0.f;
    for (int i = 0; i < n; ++i) {
        int p = m + i; if (p >= n) p -= n;
        s += dK1[(size_t)p * n * r + (size_t)j * n + i];
    }
    dG[(size_t)m * r + j] += s;
}
// dH[j][n-1-m] += sum_q sign * dK2[j*n + (m+q) mod n][q]
__global__ void tl_dh_kernel(const float* __restrict__ dK2, float* __restrict__ dH, int n, int r) {
    int m = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (m >= n) return;
    float s = 0.f;
    for (int q = 0; q < n; ++q) {
        int i = m + q; if (i >= n) i -= n;
        float v = dK2[((size_t)j * n + i) * n + q];
        s += i < q ? -v : v;
    }
    dH[(size_t)j * n + (n - 1 - m)] += s;
}

}  // namespace

extern "C" {

// Workspace (doubles) for the LDR weight build of size n with up to `max_terms` stored prefix sums.
size_t sn_ldr_workspace_doubles(int n, int max_terms) {
    size_t n2 = (size_t)n * n;
    return n2 * (4 + (size_t)max_terms) + 6 * (3 * (size_t)n + 2) + 8;
}

// Builds W (n x n float32, row-major) from the representation (A, B: COO value arrays + slot maps; G, H: n x r
// float64 row-major).  `slot` maps each COO entry to lo/di/up/corner (see band_gather_kernel).  Terms are summed
// until max|T_j| <= rel_tol * max_j max|T_j| or j = n-1 or j = max_terms-1 (then returns an error).  Synchronises
// `stream` every 8 terms to read the convergence flag.  terms_out (host) receives J+1, the number of summed terms;
// prefix sums P_0..P_J are left in the workspace for sn_ldr_backward.
int sn_ldr_build_weight(int n, int r, const double* A_vals, const int32_t* A_slot, int A_nnz, const double* B_vals,
                        const int32_t* B_slot, int B_nnz, const double* G, const double* H, double* ws, int max_terms,
                        double rel_tol, float* W_out, int* terms_out, sn_stream_t stream) {
    SN_CHECK_ARG(n > 0 && r >= 0 && ws && W_out && terms_out && max_terms >= 1, "ldr_build_weight: bad arguments");
    cudaStream_t st = snb::as_stream(stream);
    const size_t n2 = (size_t)n * n, nb = 3 * (size_t)n + 2;
    double* acc = ws;                 // running sum W
    double* T = acc + n2;             // current term
    double* tmp = T + n2;             // scratch
    double* tmp2 = tmp + n2;
    double* bandA = tmp2 + n2;
    double* bandB = bandA + nb;
    double* bandAt = bandB + nb;
    double* bandBt = bandAt + nb;
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(bandBt + nb);   // [0] = max|T_j| bits
    double* P = bandBt + nb + 8 + 2 * nb;   // stored prefix sums (after the flag words and the gradient bands)
    SN_CHECK_CUDA(cudaMemsetAsync(ws, 0, (4 * n2 + 4 * nb + 8) * sizeof(double), st));
    if (A_nnz) { band_gather_kernel<<<grid1(A_nnz), 256, 0, st>>>(A_vals, A_slot, A_nnz, bandA); SN_CHECK_LAUNCH("band_gather"); }
    if (B_nnz) { band_gather_kernel<<<grid1(B_nnz), 256, 0, st>>>(B_vals, B_slot, B_nnz, bandB); SN_CHECK_LAUNCH("band_gather"); }
    band_transpose_kernel<<<snb::ceil_div(n, 128), 128, 0, st>>>(bandA, bandAt, n); SN_CHECK_LAUNCH("band_transpose");
    band_transpose_kernel<<<snb::ceil_div(n, 128), 128, 0, st>>>(bandB, bandBt, n); SN_CHECK_LAUNCH("band_transpose");
    int terms = 0;
    if (r > 0) {
        ght_kernel<<<grid2(n), 128, 0, st>>>(G, H, T, n, r); SN_CHECK_LAUNCH("ght_kernel");
        double global_max = 0.0;
        const int limit = n < max_terms ? n : max_terms;
        bool converged = false;
        for (int j = 0; j < limit; ++j) {
            accumulate_kernel<<<grid1(n2), 256, 0, st>>>(T, acc, P + (size_t)j * n2, n2, flag); SN_CHECK_LAUNCH("accumulate_kernel");
            terms = j + 1;
            const bool last = (j + 1 == limit);
            if ((j & 7) == 7 || last) {
                unsigned long long bits = 0;
                SN_CHECK_CUDA(cudaMemcpyAsync(&bits, flag, sizeof(bits), cudaMemcpyDeviceToHost, st));
                SN_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(bits), st));
                SN_CHECK_CUDA(cudaStreamSynchronize(st));
                double m;
                memcpy(&m, &bits, sizeof(m));   // max |T| over the last <= 8 terms
                if (m > global_max) global_max = m;
                if (!(m > rel_tol * global_max) || m == 0.0) { converged = true; break; }
            }
            if (last) break;
            band_left_kernel<<<grid2(n), 128, 0, st>>>(bandA, T, tmp, n); SN_CHECK_LAUNCH("band_left_kernel");
            band_right_kernel<<<grid2(n), 128, 0, st>>>(bandB, tmp, T, n); SN_CHECK_LAUNCH("band_right_kernel");
        }
        SN_CHECK_ARG(converged || terms == n, "ldr_build_weight: the Krylov series did not converge within %d stored terms (n = %d); "
                     "raise max_terms (workspace) or reduce the norm of A, B", max_terms, n);
    }
    cast_f64_f32_kernel<<<grid1(n2), 256, 0, st>>>(acc, W_out, n2); SN_CHECK_LAUNCH("cast_f64_f32_kernel");
    *terms_out = terms;
    return 0;
}

// Backward of the weight build.  dW: n x n float32 (d loss / d W).  Accumulates into gA_vals / gB_vals (COO order,
// float64), gG, gH (n x r float64).  `ws` must be the workspace sn_ldr_build_weight left behind, `terms` its terms_out.
int sn_ldr_backward(int n, int r, const float* dW, const int32_t* A_slot, int A_nnz, const int32_t* B_slot, int B_nnz,
                    const double* G, const double* H, double* ws, int terms, double* gA_vals, double* gB_vals, double* gG,
                    double* gH, sn_stream_t stream) {
    SN_CHECK_ARG(n > 0 && dW && ws, "ldr_backward: bad arguments");
    if (r == 0 || terms == 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    const size_t n2 = (size_t)n * n, nb = 3 * (size_t)n + 2;
    double* dM = ws;                  // reuse: running sum of S_l
    double* S = dM + n2;
    double* tmp = S + n2;
    double* tmp2 = tmp + n2;
    double* bandA = tmp2 + n2;
    double* bandB = bandA + nb;
    double* bandAt = bandB + nb;
    double* bandBt = bandAt + nb;
    double* gA = bandBt + nb + 8;     // gradient bands of A and B
    double* gB = gA + nb;
    double* P = gB + nb;
    SN_CHECK_CUDA(cudaMemsetAsync(dM, 0, n2 * sizeof(double), st));
    SN_CHECK_CUDA(cudaMemsetAsync(gA, 0, 2 * nb * sizeof(double), st));
    cast_f32_f64_kernel<<<grid1(n2), 256, 0, st>>>(dW, S, n2); SN_CHECK_LAUNCH("cast_f32_f64_kernel");
    const int J = terms - 1;
    for (int l = 0; l <= J; ++l) {
        accumulate_kernel<<<grid1(n2), 256, 0, st>>>(S, dM, nullptr, n2, reinterpret_cast<unsigned long long*>(bandBt + nb)); SN_CHECK_LAUNCH("accumulate_kernel");
        if (l == J) break;
        const double* Pk = P + (size_t)(J - 1 - l) * n2;
        band_right_kernel<<<grid2(n), 128, 0, st>>>(bandB, Pk, tmp, n); SN_CHECK_LAUNCH("band_right_kernel");            // Q = P B
        row_band_dots_kernel<<<snb::ceil_div(n, 4), 128, 0, st>>>(S, tmp, gA, n); SN_CHECK_LAUNCH("row_band_dots_kernel");  // dA += pattern(S Q^T)
        band_left_kernel<<<grid2(n), 128, 0, st>>>(bandAt, S, tmp2, n); SN_CHECK_LAUNCH("band_left_kernel");             // V = A^T S
        col_band_dots_kernel<<<dim3(snb::ceil_div(n, 128), n >= 256 ? 64 : 1), 128, 0, st>>>(Pk, tmp2, gB, n); SN_CHECK_LAUNCH("col_band_dots_kernel"); // dB += pattern(P^T V)
        band_right_kernel<<<grid2(n), 128, 0, st>>>(bandBt, tmp2, S, n); SN_CHECK_LAUNCH("band_right_kernel");           // S = V B^T
    }
    if (gA_vals && A_nnz) { band_scatter_grad_kernel<<<grid1(A_nnz), 256, 0, st>>>(gA, A_slot, A_nnz, gA_vals); SN_CHECK_LAUNCH("band_scatter_grad"); }
    if (gB_vals && B_nnz) { band_scatter_grad_kernel<<<grid1(B_nnz), 256, 0, st>>>(gB, B_slot, B_nnz, gB_vals); SN_CHECK_LAUNCH("band_scatter_grad"); }
    if (gG && gH) { dgh_kernel<<<dim3(snb::ceil_div(r, 32), n), 32, 0, st>>>(dM, G, H, gG, gH, n, r); SN_CHECK_LAUNCH("dgh_kernel"); }
    return 0;
}

// y[B x n_out] = x[B x n_in] W^T + bias  and  dW[n_out x n_in] += gy^T x ; plain fp32 GEMMs shared by LDR and TL.
int sn_dense_apply(const float* W, int n_out, int n_in, const float* x, int64_t ldx, float* y, int64_t ldy, const float* bias,
                   int64_t B, sn_stream_t stream) {
    SN_CHECK_ARG(W && x && y, "dense_apply: NULL buffer");
    return snb::gemm_f32(false, true, (int)B, n_out, n_in, 1.f, x, ldx, W, n_in, 0.f, y, ldy, bias, snb::as_stream(stream));
}
int sn_dense_weight_grad(const float* x, int64_t ldx, const float* gy, int64_t ldgy, float* dW, int n_out, int n_in, float* grad_bias,
                         int64_t B, sn_stream_t stream) {
    SN_CHECK_ARG(x && gy && dW, "dense_weight_grad: NULL buffer");
    cudaStream_t st = snb::as_stream(stream);
    if (grad_bias)
        if (int rc = snb::colsum_accumulate(gy, ldgy, B, n_out, grad_bias, st)) return rc;
    return snb::gemm_f32(true, false, n_out, n_in, (int)B, 1.f, gy, ldgy, x, ldx, 1.f, dW, n_in, nullptr, st, true);
}

// Toeplitz-like weight: W (n x n) from G (n x r), H (r x n); K1 (n x rn) and K2 (rn x n) are caller-provided scratch
// that sn_tl_backward reuses.
int sn_tl_build_weight(int n, int r, const float* G, const float* H, float* K1, float* K2, float* W, sn_stream_t stream) {
    SN_CHECK_ARG(n > 0 && r > 0 && G && H && K1 && K2 && W, "tl_build_weight: bad arguments");
    cudaStream_t st = snb::as_stream(stream);
    tl_k1_kernel<<<dim3(snb::ceil_div(n * r, 128), n), 128, 0, st>>>(G, K1, n, r); SN_CHECK_LAUNCH("tl_k1_kernel");
    tl_k2_kernel<<<dim3(n * r, snb::ceil_div(n, 128)), 128, 0, st>>>(H, K2, n, r); SN_CHECK_LAUNCH("tl_k2_kernel");
    return snb::gemm_f32(false, false, n, n, n * r, 0.5f, K1, (long)n * r, K2, n, 0.f, W, n, nullptr, st);
}
// dK1 / dK2: scratch of the same sizes as K1 / K2; gG (n x r), gH (r x n) are accumulated into.
int sn_tl_backward(int n, int r, const float* dW, const float* K1, const float* K2, float* dK1, float* dK2, float* gG, float* gH,
                   sn_stream_t stream) {
    SN_CHECK_ARG(n > 0 && r > 0 && dW && K1 && K2 && dK1 && dK2 && gG && gH, "tl_backward: bad arguments");
    cudaStream_t st = snb::as_stream(stream);
    // dK1 = 0.5 dW K2^T  (n x n)(rn x n)^T ; dK2 = 0.5 K1^T dW  (n x rn)^T (n x n)
    if (int rc = snb::gemm_f32(false, true, n, n * r, n, 0.5f, dW, n, K2, n, 0.f, dK1, (long)n * r, nullptr, st)) return rc;
    if (int rc = snb::gemm_f32(true, false, n * r, n, n, 0.5f, K1, (long)n * r, dW, n, 0.f, dK2, n, nullptr, st)) return rc;
    tl_dg_kernel<<<dim3(snb::ceil_div(r, 32), n), 32, 0, st>>>(dK1, gG, n, r); SN_CHECK_LAUNCH("tl_dg_kernel");
    tl_dh_kernel<<<dim3(snb::ceil_div(n, 128), r), 128, 0, st>>>(dK2, gH, n, r); SN_CHECK_LAUNCH("tl_dh_kernel");
    return 0;
}

}  // extern "C"
