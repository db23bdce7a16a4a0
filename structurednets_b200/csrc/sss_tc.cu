// SSS layer on the tensor cores: chunked formulation of the stage recursion (reference layers/sss_layer.py:99-131)
// evaluated as 3xTF32 GEMMs on tcgen05 / TMEM / TMA, fp32-accurate (hi/lo operand split, fp32 accumulation in TMEM).
//
// A chunk j is a run of consecutive stages [k0, k1) with <= 32 outputs, <= 160 inputs and <= 16 stages.  Multiplying
// the stages of a chunk out gives dense chunk matrices (tests/sss_tc_emulator.py is the torch statement of this):
//       y_j     = T_j u_j + O_j s_j + O'_j e_{j+1}        s_j: causal state entering the chunk from below
//       s_{j+1} = R_j u_j + Phi_j s_j                      e_{j+1}: anticausal state entering it from above
//       e_j     = R'_j u_j + Phi'_j e_{j+1}
// Forward   1. sss_tc_build_kernel     params -> W_j = [T_j; R_j; R'_j] (64 x 160, hi and lo tf32 parts), SC_j = {Phi, Phi', O, O'}
//           2. sss_tc_local_gemm_kernel [yloc | r | r'] = u_j W_j^T   for every (128-sample tile, chunk): tcgen05.mma kind::tf32,
//                                       x tiles and W tiles by TMA (SWIZZLE_128B), hi/lo split of x by converter warps,
//                                       accumulators double-buffered in TMEM, persistent CTAs
//           3. sss_tc_scan_fwd_kernel  chunk-level state scans (one thread per sample), y = yloc + O s + O' e + bias;
//                                       saves the chunk-boundary states for the backward
// Backward  4. sss_tc_scan_bwd_kernel  adjoint chunk scans: lambda_j = Phi_j^T lambda_{j+1} + O_j^T gy_j, mu likewise; grad_bias
//           5. sss_tc_grad_gemm_kernel dM_j = [gy_j | lambda_{j+1} | mu_j]^T [u_j | s_j | e_{j+1}]  (K = samples, MN-major operands
//                                       straight from the natural [sample][feature] layouts), split over sample ranges
//           6. sss_tc_build_bwd_kernel chain rule through the chunk-matrix construction: dM -> dA .. dG
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace snb::umma;
using snb::ceil_div;

constexpr int DS = SN_SSS_TC_DS;           // 16 padded state dimension
constexpr int PO = SN_SSS_TC_PO;           // 32 outputs per chunk
constexpr int KBW = 32;                    // input columns per k-block (128 bytes)
constexpr int KB_MAX = SN_SSS_TC_KB_MAX;   // 5 k-blocks per chunk
constexpr int LMAX = SN_SSS_TC_LMAX;       // 16 stages per chunk
constexpr int SOUT_MAX = 16;               // outputs of one stage
constexpr int WROWS = 128;                 // 64 hi rows + 64 lo rows
constexpr int WCOLS = KBW * KB_MAX;        // 160
constexpr int SCF = 2 * DS * DS + 2 * PO * DS;   // 1536 floats: Phi, Phi', O, O'
constexpr int DMC = KBW * (KB_MAX + 1);    // 192 columns of dM: u (<=160) then s (16), e (16) right after the last u block
constexpr int COLT = WCOLS + DS;           // 176 columns handled by the build kernels (inputs + unit states)
constexpr int BUILD_THREADS = 192;


__device__ __forceinline__ const sn_sss_stage& stage_of(const sn_sss_stage* stages, int n, int dir, int k) {
    return stages[dir * n + (dir == 0 ? k : n - 1 - k)];
}

// One stage applied to one column of the chunk's "identity input": v (state entering) -> yv (the stage's outputs), v (state leaving)
__device__ __forceinline__ void stage_apply(const sn_sss_stage& st, const float* __restrict__ params, float (&v)[DS], float (&yv)[SOUT_MAX], bool mine,
                                            int local) {
    const int d_in = st.d_in, d_out = st.d_out;
#pragma unroll
    for (int r = 0; r < SOUT_MAX; ++r) {
        float acc = 0.f;
        if (r < st.out_dim) {
            const float* ys = params + st.off_ys + r * d_in;
#pragma unroll
            for (int a = 0; a < DS; ++a)
                if (a < d_in) acc = fmaf(__ldg(ys + a), v[a], acc);
            if (mine && st.off_yu >= 0) acc += __ldg(params + st.off_yu + r * st.in_dim + local);
        }
        yv[r] = acc;
    }
    float nv[DS];
#pragma unroll
    for (int b = 0; b < DS; ++b) {
        float acc = 0.f;
        if (b < d_out) {
            const float* ss = params + st.off_ss + b * d_in;
#pragma unroll
            for (int a = 0; a < DS; ++a)
                if (a < d_in) acc = fmaf(__ldg(ss + a), v[a], acc);
            if (mine) acc += __ldg(params + st.off_su + b * st.in_dim + local);
        }
        nv[b] = acc;
    }
#pragma unroll
    for (int b = 0; b < DS; ++b) v[b] = nv[b];
}

__device__ __forceinline__ void store_hi_lo(float* W, int row, int t, float val) {
    const float hi = tf32_hi(val);
    W[row * WCOLS + t] = hi;
    W[(64 + row) * WCOLS + t] = tf32_hi(val - hi);
}

// ------------------------------------------------------------------------------------------
// 1. chunk matrices.  grid (nchunks, 2 directions); thread t < ncols: input column t of the chunk, ncols <= t < ncols+16: unit state
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BUILD_THREADS)
sss_tc_build_kernel(const sn_sss_stage* __restrict__ stages, int n, const sn_sss_tc_chunk* __restrict__ chunks, const float* __restrict__ params,
                    float* __restrict__ Wall, float* __restrict__ SCall) {
    const sn_sss_tc_chunk c = chunks[blockIdx.x];
    const int dir = blockIdx.y, t = threadIdx.x;
    const bool is_in = t < c.ncols;
    const int sidx = t - c.ncols;
    if (!is_in && sidx >= DS) return;
    float* W = Wall + (size_t)blockIdx.x * WROWS * WCOLS;
    float* SC = SCall + (size_t)blockIdx.x * SCF;
    float* Phi = SC + dir * DS * DS;
    float* Omat = SC + 2 * DS * DS + dir * PO * DS;
    const int nst = c.k_end - c.k_begin;
    const int col = c.col0 + t;
    float v[DS], yv[SOUT_MAX];
    {
        const sn_sss_stage& st0 = stage_of(stages, n, dir, dir == 0 ? c.k_begin : c.k_end - 1);
#pragma unroll
        for (int a = 0; a < DS; ++a) v[a] = (!is_in && a == sidx && sidx < st0.d_in) ? 1.f : 0.f;
    }
    bool activated = false;
    for (int i = 0; i < nst; ++i) {
        const sn_sss_stage st = stage_of(stages, n, dir, dir == 0 ? c.k_begin + i : c.k_end - 1 - i);
        const int local = col - st.in_off;
        const bool mine = is_in && local >= 0 && local < st.in_dim;
        stage_apply(st, params, v, yv, mine, local);
        const int rbase = st.out_off - c.row0;
        const bool wr = dir == 0 ? (activated || mine) : activated;
#pragma unroll
        for (int r = 0; r < SOUT_MAX; ++r) {
            if (r < st.out_dim) {
                if (is_in) {
                    if (wr) store_hi_lo(W, rbase + r, t, yv[r]);
                } else {
                    Omat[(rbase + r) * DS + sidx] = yv[r];
                }
            }
        }
        if (mine) activated = true;
    }
#pragma unroll
    for (int b = 0; b < DS; ++b) {
        if (is_in) store_hi_lo(W, PO + dir * DS + b, t, v[b]);
        else Phi[b * DS + sidx] = v[b];
    }
}

// ------------------------------------------------------------------------------------------
// 2. local GEMM: [yloc | r | r'](128 samples x 64) = u_j (128 x 32 nkb) * W_j^T, 3xTF32
//    warp 0: TMA producer | warp 1: TMEM alloc + MMA issuer | warps 2-5: hi/lo converters | warps 6-9: epilogue
// ------------------------------------------------------------------------------------------
constexpr int G1_THREADS = 320;
constexpr int G1_STAGES = 4;
constexpr int G1_TILE_BYTES = 128 * 128;              // 128 rows x 128 B
constexpr int G1_STAGE_BYTES = 3 * G1_TILE_BYTES;     // x (hi in place) | x_lo | W
constexpr size_t G1_SMEM = (size_t)G1_STAGES * G1_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(G1_THREADS, 1)
sss_tc_local_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                         const sn_sss_tc_chunk* __restrict__ chunks, int nchunks, long B, int ntiles, float* __restrict__ rbuf) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G1_STAGES * G1_STAGE_BYTES);
    uint64_t* conv = full + G1_STAGES;
    uint64_t* empty = conv + G1_STAGES;
    uint64_t* acc_full = empty + G1_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long total = (long)ntiles * nchunks;
    const long per = (total + gridDim.x - 1) / gridDim.x;
    const long w0 = (long)blockIdx.x * per;
    const long w1 = w0 + per < total ? w0 + per : total;
    if (w0 >= w1) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G1_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(conv + s, 128); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, 128); }
        mbar_fence_init();
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (long w = w0; w < w1; ++w) {
                const int tile = (int)(w / nchunks), ch = (int)(w % nchunks);
                const int nkb = chunks[ch].nkb, col0 = chunks[ch].col0;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % G1_STAGES, round = it / G1_STAGES;
                    if (round > 0) mbar_wait(empty + s, (round - 1) & 1);
                    uint8_t* st = smem + s * G1_STAGE_BYTES;
                    mbar_expect_tx(full + s, 2 * G1_TILE_BYTES);
                    tma_load_2d(st, &map_x, col0 + kb * KBW, tile * 128, full + s);
                    tma_load_2d(st + 2 * G1_TILE_BYTES, &map_w, kb * KBW, ch * WROWS, full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc1 = idesc_tf32(128, 128, false, false);   // x_hi * [W_hi ; W_lo]
            constexpr uint32_t idesc2 = idesc_tf32(128, 64, false, false);    // x_lo * W_hi
            uint32_t it = 0, ai = 0;
            for (long w = w0; w < w1; ++w, ++ai) {
                const int ch = (int)(w % nchunks);
                const int nkb = chunks[ch].nkb;
                const uint32_t b = ai & 1;
                if (ai >= 2) mbar_wait(acc_empty + b, ((ai >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + b * 128;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % G1_STAGES, round = it / G1_STAGES;
                    mbar_wait(conv + s, round & 1);
                    tc_fence_after();
                    uint8_t* st = smem + s * G1_STAGE_BYTES;
                    const uint64_t dxh = desc_kmajor_sw128(st);
                    const uint64_t dxl = desc_kmajor_sw128(st + G1_TILE_BYTES);
                    const uint64_t dw = desc_kmajor_sw128(st + 2 * G1_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < KBW / 8; ++k) {   // one tf32 UMMA = 8 floats = 32 bytes of K (+2 in 16-byte units)
                        mma_tf32(acc, dxh + 2 * k, dw + 2 * k, idesc1, (kb | k) ? 1u : 0u);
                        mma_tf32(acc, dxl + 2 * k, dw + 2 * k, idesc2, 1u);
                    }
                    umma_commit(empty + s);
                }
                umma_commit(acc_full + b);
            }
        }
    } else if (warp < 6) {
        // converters: x -> hi (in place), lo (second tile); identical swizzled layout, so plain 16-byte chunks
        const int ct = threadIdx.x - 64;
        uint32_t it = 0;
        for (long w = w0; w < w1; ++w) {
            const int nkb = chunks[(int)(w % nchunks)].nkb;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t s = it % G1_STAGES, round = it / G1_STAGES;
                mbar_wait(full + s, round & 1);
                float4* xh = reinterpret_cast<float4*>(smem + s * G1_STAGE_BYTES);
                float4* xl = xh + G1_TILE_BYTES / 16;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = xh[ct + 128 * i];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 h, l;
                    h.x = tf32_hi(v[i].x); h.y = tf32_hi(v[i].y); h.z = tf32_hi(v[i].z); h.w = tf32_hi(v[i].w);
                    l.x = v[i].x - h.x; l.y = v[i].y - h.y; l.z = v[i].z - h.z; l.w = v[i].w - h.w;
                    xh[ct + 128 * i] = h;
                    xl[ct + 128 * i] = l;
                }
                fence_async_smem();
                mbar_arrive(conv + s);
            }
        }
    } else {
        // epilogue: TMEM lane quarter = warp % 4; thread = one sample row
        const int q = warp & 3;
        uint32_t ai = 0;
        for (long w = w0; w < w1; ++w, ++ai) {
            const int tile = (int)(w / nchunks), ch = (int)(w % nchunks);
            const uint32_t b = ai & 1;
            mbar_wait(acc_full + b, (ai >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + b * 128 + ((uint32_t)(q * 32) << 16);
            float out[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t m[16], x2[16];
                tmem_ld16_nowait(acc + c0, m);
                tmem_ld16_nowait(acc + 64 + c0, x2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) out[c0 + i] = __uint_as_float(m[i]) + __uint_as_float(x2[i]);
            }
            tc_fence_before();
            mbar_arrive(acc_empty + b);
            const long row = (long)tile * 128 + q * 32 + lane;
            if (row < B) {
                float4* dst = reinterpret_cast<float4*>(rbuf + ((size_t)ch * B + row) * 64);
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// 3. forward chunk scans + output fix-up; one thread per sample
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 128;

__device__ __forceinline__ void load_sc(float4* dst, const float* __restrict__ src, int nfloat4) {
    for (int i = threadIdx.x; i < nfloat4; i += SCAN_THREADS) dst[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
}
// out[b] += sum_a M[b][a] v[a]   (M row-major DS columns in shared memory, read as broadcast float4)
template <int ROWS>
__device__ __forceinline__ void matvec_acc(const float4* __restrict__ M, const float (&v)[DS], float* out) {
#pragma unroll
    for (int b = 0; b < ROWS; ++b) {
        float acc = out[b];
#pragma unroll
        for (int a4 = 0; a4 < DS / 4; ++a4) {
            const float4 m = M[b * (DS / 4) + a4];
            acc = fmaf(m.x, v[4 * a4], acc);
            acc = fmaf(m.y, v[4 * a4 + 1], acc);
            acc = fmaf(m.z, v[4 * a4 + 2], acc);
            acc = fmaf(m.w, v[4 * a4 + 3], acc);
        }
        out[b] = acc;
    }
}
// out[a] += sum_b M[b][a] w[b]   (transposed application)
template <int ROWS>
__device__ __forceinline__ void matvec_t_acc(const float4* __restrict__ M, const float* w, float (&out)[DS]) {
#pragma unroll
    for (int b = 0; b < ROWS; ++b) {
        const float wb = w[b];
#pragma unroll
        for (int a4 = 0; a4 < DS / 4; ++a4) {
            const float4 m = M[b * (DS / 4) + a4];
            out[4 * a4] = fmaf(m.x, wb, out[4 * a4]);
            out[4 * a4 + 1] = fmaf(m.y, wb, out[4 * a4 + 1]);
            out[4 * a4 + 2] = fmaf(m.z, wb, out[4 * a4 + 2]);
            out[4 * a4 + 3] = fmaf(m.w, wb, out[4 * a4 + 3]);
        }
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
sss_tc_scan_fwd_kernel(const sn_sss_tc_chunk* __restrict__ chunks, int nchunks, const float* __restrict__ SCall, const float* __restrict__ rbuf,
                       float* __restrict__ S, float* __restrict__ y, long ldy, const float* __restrict__ bias, long B, int aligned) {
    __shared__ float4 sc[SCF / 4];
    const long row = (long)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const bool valid = row < B;
    // anticausal states, top chunk first: S[j][row][16..31] = e_{j+1} (the state entering chunk j from above)
    float e[DS];
#pragma unroll
    for (int a = 0; a < DS; ++a) e[a] = 0.f;
    for (int j = nchunks - 1; j >= 0; --j) {
        __syncthreads();
        load_sc(sc, SCall + (size_t)j * SCF + DS * DS, DS * DS / 4);   // Phi'
        __syncthreads();
        if (valid) {
            const size_t base = (size_t)j * B + row;
            float4* sdst = reinterpret_cast<float4*>(S + base * 32 + DS);
#pragma unroll
            for (int i = 0; i < 4; ++i) sdst[i] = make_float4(e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]);
            float ne[DS];
            const float4* rp = reinterpret_cast<const float4*>(rbuf + base * 64 + 48);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 r4 = __ldg(rp + i);
                ne[4 * i] = r4.x; ne[4 * i + 1] = r4.y; ne[4 * i + 2] = r4.z; ne[4 * i + 3] = r4.w;
            }
            matvec_acc<DS>(sc, e, ne);
#pragma unroll
            for (int a = 0; a < DS; ++a) e[a] = ne[a];
        }
    }
    // causal states bottom up + outputs
    float s[DS];
#pragma unroll
    for (int a = 0; a < DS; ++a) s[a] = 0.f;
    for (int j = 0; j < nchunks; ++j) {
        __syncthreads();
        load_sc(sc, SCall + (size_t)j * SCF, SCF / 4);
        __syncthreads();
        if (valid) {
            const sn_sss_tc_chunk c = chunks[j];
            const size_t base = (size_t)j * B + row;
            float4* sdst = reinterpret_cast<float4*>(S + base * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) sdst[i] = make_float4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 e4 = sdst[4 + i];
                e[4 * i] = e4.x; e[4 * i + 1] = e4.y; e[4 * i + 2] = e4.z; e[4 * i + 3] = e4.w;
            }
            const float4* rp = reinterpret_cast<const float4*>(rbuf + base * 64);
            float* yrow = y + row * ldy + c.row0;
            const float4* Om = sc + 2 * DS * DS / 4;
            const float4* Opm = Om + PO * DS / 4;
#pragma unroll
            for (int g = 0; g < PO / 4; ++g) {
                if (4 * g < c.nrows) {
                    const float4 yl = __ldg(rp + g);
                    float acc[4] = {yl.x, yl.y, yl.z, yl.w};
                    matvec_acc<4>(Om + g * 4 * (DS / 4), s, acc);
                    matvec_acc<4>(Opm + g * 4 * (DS / 4), e, acc);
                    if (aligned && 4 * g + 4 <= c.nrows) {
                        if (bias != nullptr) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c.row0 + 4 * g));
                            acc[0] += b4.x; acc[1] += b4.y; acc[2] += b4.z; acc[3] += b4.w;
                        }
                        *reinterpret_cast<float4*>(yrow + 4 * g) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (4 * g + i < c.nrows) yrow[4 * g + i] = acc[i] + (bias != nullptr ? __ldg(bias + c.row0 + 4 * g + i) : 0.f);
                    }
                }
            }
            float ns[DS];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 r4 = __ldg(rp + 8 + i);
                ns[4 * i] = r4.x; ns[4 * i + 1] = r4.y; ns[4 * i + 2] = r4.z; ns[4 * i + 3] = r4.w;
            }
            matvec_acc<DS>(sc, s, ns);
#pragma unroll
            for (int a = 0; a < DS; ++a) s[a] = ns[a];
        }
    }
}

// ------------------------------------------------------------------------------------------
// 4. adjoint chunk scans.  L[j][row][0..15] = lambda_{j+1} (adjoint of the causal state leaving chunk j),
//    L[j][row][16..31] = mu_j (adjoint of the anticausal state leaving chunk j).  grad_bias += column sums of gy.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_gy(const float* __restrict__ gy, long ldgy, long row, const sn_sss_tc_chunk& c, bool valid, int aligned,
                                        float (&g)[PO]) {
    const float* src = gy + row * ldgy + c.row0;
#pragma unroll
    for (int q = 0; q < PO / 4; ++q) {
        if (valid && aligned && 4 * q + 4 <= c.nrows) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + 4 * q));
            g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) g[4 * q + i] = (valid && 4 * q + i < c.nrows) ? __ldg(src + 4 * q + i) : 0.f;
        }
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
sss_tc_scan_bwd_kernel(const sn_sss_tc_chunk* __restrict__ chunks, int nchunks, const float* __restrict__ SCall, const float* __restrict__ gy,
                       long ldgy, float* __restrict__ L, float* __restrict__ gbias, long B, int aligned) {
    __shared__ float4 sc[(DS * DS + PO * DS) / 4];
    const long row = (long)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const bool valid = row < B;
    const int lane = threadIdx.x & 31;
    float lam[DS], g[PO];
#pragma unroll
    for (int a = 0; a < DS; ++a) lam[a] = 0.f;
    for (int j = nchunks - 1; j >= 0; --j) {
        __syncthreads();
        load_sc(sc, SCall + (size_t)j * SCF, DS * DS / 4);                                       // Phi
        load_sc(sc + DS * DS / 4, SCall + (size_t)j * SCF + 2 * DS * DS, PO * DS / 4);           // O
        __syncthreads();
        const sn_sss_tc_chunk c = chunks[j];
        load_gy(gy, ldgy, row, c, valid, aligned, g);
        if (valid) {
            float4* dst = reinterpret_cast<float4*>(L + ((size_t)j * B + row) * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_float4(lam[4 * i], lam[4 * i + 1], lam[4 * i + 2], lam[4 * i + 3]);
            float nl[DS];
#pragma unroll
            for (int a = 0; a < DS; ++a) nl[a] = 0.f;
            matvec_t_acc<DS>(sc, lam, nl);
            matvec_t_acc<PO>(sc + DS * DS / 4, g, nl);
#pragma unroll
            for (int a = 0; a < DS; ++a) lam[a] = nl[a];
        }
    }
    float mu[DS];
#pragma unroll
    for (int a = 0; a < DS; ++a) mu[a] = 0.f;
    for (int j = 0; j < nchunks; ++j) {
        __syncthreads();
        load_sc(sc, SCall + (size_t)j * SCF + DS * DS, DS * DS / 4);                             // Phi'
        load_sc(sc + DS * DS / 4, SCall + (size_t)j * SCF + 2 * DS * DS + PO * DS, PO * DS / 4);  // O'
        __syncthreads();
        const sn_sss_tc_chunk c = chunks[j];
        load_gy(gy, ldgy, row, c, valid, aligned, g);
        if (valid) {
            float4* dst = reinterpret_cast<float4*>(L + ((size_t)j * B + row) * 32 + DS);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_float4(mu[4 * i], mu[4 * i + 1], mu[4 * i + 2], mu[4 * i + 3]);
            float nm[DS];
#pragma unroll
            for (int a = 0; a < DS; ++a) nm[a] = 0.f;
            matvec_t_acc<DS>(sc, mu, nm);
            matvec_t_acc<PO>(sc + DS * DS / 4, g, nm);
#pragma unroll
            for (int a = 0; a < DS; ++a) mu[a] = nm[a];
        }
        if (gbias != nullptr) {
            // column sums over the warp's 32 samples: lane l ends up with the sum of column l
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const bool up = (lane & off) != 0;
                    const float send = up ? g[i] : g[i + off];
                    const float keep = up ? g[i + off] : g[i];
                    g[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            if (lane < c.nrows) atomicAdd(gbias + c.row0 + lane, g[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// 5. gradient GEMM: dM_j (64 x N) += [gy_j | lambda | mu]^T [u_j | s | e] over a range of samples, 3xTF32 (all four hi/lo terms).
//    K = samples; both operands MN-major straight from TMA boxes of [32 samples][32 floats] (SWIZZLE_128B_ATOM_32B, the one
//    layout kind::tf32 accepts for MN-major operands).
//    A (M = 128): blocks gy_hi, L_hi, gy_lo, L_lo;  B (N = 32 (nkb+1)): x blocks then the state block; hi and lo tiles.
// ------------------------------------------------------------------------------------------
constexpr int G2_THREADS = 320;
constexpr int G2_STAGES = 3;
constexpr int G2_KS = 32;                        // samples per stage
constexpr int G2_BLK = G2_KS * 128;              // one 32-feature block: 4 KB
constexpr int G2_A_BYTES = 4 * G2_BLK;           // 16 KB
constexpr int G2_BH_BYTES = (KB_MAX + 1) * G2_BLK;   // 24 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + 2 * G2_BH_BYTES;   // 64 KB
constexpr size_t G2_SMEM = (size_t)G2_STAGES * G2_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(G2_THREADS, 1)
sss_tc_grad_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_l,
                        const __grid_constant__ CUtensorMap map_s, const sn_sss_tc_chunk* __restrict__ chunks, long B, float* __restrict__ dM) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
    uint64_t* conv = full + G2_STAGES;
    uint64_t* empty = conv + G2_STAGES;
    uint64_t* acc_full = empty + G2_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.y;
    const sn_sss_tc_chunk c = chunks[ch];
    const int nkb = c.nkb;
    const int nblk = nkb + 1;                       // N blocks of 32
    const int ntk = (int)((B + G2_KS - 1) / G2_KS);
    const int per = (ntk + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * per;
    const int t1 = t0 + per < ntk ? t0 + per : ntk;
    if (t0 >= t1) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G2_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(conv + s, 128); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_gy);
        tma_prefetch_desc(&map_l);
        tma_prefetch_desc(&map_s);
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int t = t0, it = 0; t < t1; ++t, ++it) {
                const int s = it % G2_STAGES, round = it / G2_STAGES;
                if (round > 0) mbar_wait(empty + s, (round - 1) & 1);
                uint8_t* st = smem + s * G2_STAGE_BYTES;
                uint8_t* bh = st + G2_A_BYTES;
                mbar_expect_tx(full + s, (2 + nblk) * G2_BLK);
                tma_load_2d(st, &map_gy, c.row0, t * G2_KS, full + s);
                tma_load_3d(st + G2_BLK, &map_l, 0, t * G2_KS, ch, full + s);
                for (int i = 0; i < nkb; ++i) tma_load_2d(bh + i * G2_BLK, &map_x, c.col0 + i * KBW, t * G2_KS, full + s);
                tma_load_3d(bh + nkb * G2_BLK, &map_s, 0, t * G2_KS, ch, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(128, 32 * nblk, true, true);
            for (int t = t0, it = 0; t < t1; ++t, ++it) {
                const int s = it % G2_STAGES, round = it / G2_STAGES;
                mbar_wait(conv + s, round & 1);
                tc_fence_after();
                uint8_t* st = smem + s * G2_STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < G2_KS / 8; ++k) {   // 8 samples = 8 rows of 128 bytes = 1024 bytes per UMMA
                    const uint64_t da = desc_mnmajor_sw128_32b(st + k * 1024, G2_BLK);
                    const uint64_t dbh = desc_mnmajor_sw128_32b(st + G2_A_BYTES + k * 1024, G2_BLK);
                    const uint64_t dbl = desc_mnmajor_sw128_32b(st + G2_A_BYTES + G2_BH_BYTES + k * 1024, G2_BLK);
                    mma_tf32(tmem_base, da, dbh, idesc, (it | k) ? 1u : 0u);
                    mma_tf32(tmem_base, da, dbl, idesc, 1u);
                }
                umma_commit(empty + s);
            }
            umma_commit(acc_full);
        }
    } else if (warp < 6) {
        const int ct = threadIdx.x - 64;
        for (int t = t0, it = 0; t < t1; ++t, ++it) {
            const int s = it % G2_STAGES, round = it / G2_STAGES;
            mbar_wait(full + s, round & 1);
            uint8_t* st = smem + s * G2_STAGE_BYTES;
            {   // A: gy, L (8 KB) -> lo at +8 KB
                float4* h = reinterpret_cast<float4*>(st);
                float4* l = h + 2 * G2_BLK / 16;
#pragma unroll
                for (int i = 0; i < 2 * G2_BLK / 16 / 128; ++i) {
                    const float4 v = h[ct + 128 * i];
                    float4 hh, ll;
                    hh.x = tf32_hi(v.x); hh.y = tf32_hi(v.y); hh.z = tf32_hi(v.z); hh.w = tf32_hi(v.w);
                    ll.x = v.x - hh.x; ll.y = v.y - hh.y; ll.z = v.z - hh.z; ll.w = v.w - hh.w;
                    h[ct + 128 * i] = hh;
                    l[ct + 128 * i] = ll;
                }
            }
            {   // B: x blocks + state block -> lo at +24 KB
                float4* h = reinterpret_cast<float4*>(st + G2_A_BYTES);
                float4* l = h + G2_BH_BYTES / 16;
                const int n16 = nblk * G2_BLK / 16;
#pragma unroll 4
                for (int i = ct; i < n16; i += 128) {
                    const float4 v = h[i];
                    float4 hh, ll;
                    hh.x = tf32_hi(v.x); hh.y = tf32_hi(v.y); hh.z = tf32_hi(v.z); hh.w = tf32_hi(v.w);
                    ll.x = v.x - hh.x; ll.y = v.y - hh.y; ll.z = v.z - hh.z; ll.w = v.w - hh.w;
                    h[i] = hh;
                    l[i] = ll;
                }
            }
            fence_async_smem();
            mbar_arrive(conv + s);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int m = q * 32 + lane;
        float* drow = dM + ((size_t)ch * 64 + (m & 63)) * DMC;
        const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < 32 * nblk; c0 += 16) {
            uint32_t v[16];
            tmem_ld16_nowait(acc + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + 4 * i), "f"(__uint_as_float(v[4 * i])),
                             "f"(__uint_as_float(v[4 * i + 1])), "f"(__uint_as_float(v[4 * i + 2])), "f"(__uint_as_float(v[4 * i + 3]))
                             : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------
// 6. chain rule through the chunk-matrix construction.  grid (nchunks, 2); thread = column as in the build kernel.
//    Phase 1 replays the construction and stores the state entering every stage; phase 2 walks the stages backwards.
// ------------------------------------------------------------------------------------------
constexpr int BB_LD = 193;   // shared-memory row stride (odd: the 16 'a' rows of a dot product fall into distinct banks)

__global__ void __launch_bounds__(BUILD_THREADS)
sss_tc_build_bwd_kernel(const sn_sss_stage* __restrict__ stages, int n, const sn_sss_tc_chunk* __restrict__ chunks, const float* __restrict__ params,
                        const float* __restrict__ dMall, float* __restrict__ scratch, float* __restrict__ gparams) {
    __shared__ float X[(DS + SOUT_MAX) * BB_LD];   // rows 0..15: lambda, 16..31: this stage's gy
    __shared__ float V[DS * BB_LD];                // state entering the stage
    const sn_sss_tc_chunk c = chunks[blockIdx.x];
    const int dir = blockIdx.y, t = threadIdx.x;
    const bool is_in = t < c.ncols;
    const int sidx = t - c.ncols;
    const bool active = is_in || sidx < DS;
    const int ncol_active = c.ncols + DS;
    const float* dM = dMall + (size_t)blockIdx.x * 64 * DMC;
    float* scr = scratch + ((size_t)blockIdx.x * 2 + dir) * LMAX * DS * BUILD_THREADS;
    const int nst = c.k_end - c.k_begin;
    const int col = c.col0 + t;
    const int cidx = is_in ? t : c.nkb * KBW + dir * DS + sidx;

    float v[DS], yv[SOUT_MAX];
    {
        const sn_sss_stage& st0 = stage_of(stages, n, dir, dir == 0 ? c.k_begin : c.k_end - 1);
#pragma unroll
        for (int a = 0; a < DS; ++a) v[a] = (active && !is_in && a == sidx && sidx < st0.d_in) ? 1.f : 0.f;
    }
    int my_i = -1;
    for (int i = 0; i < nst; ++i) {
        const sn_sss_stage st = stage_of(stages, n, dir, dir == 0 ? c.k_begin + i : c.k_end - 1 - i);
        const int local = col - st.in_off;
        const bool mine = is_in && local >= 0 && local < st.in_dim;
#pragma unroll
        for (int a = 0; a < DS; ++a) scr[(i * DS + a) * BUILD_THREADS + t] = v[a];
        if (active) stage_apply(st, params, v, yv, mine, local);
        if (mine) my_i = i;
    }
    // adjoint of the state leaving the last stage: dR / dPhi rows of dM
    float lam[DS];
#pragma unroll
    for (int b = 0; b < DS; ++b) lam[b] = active ? dM[(PO + dir * DS + b) * DMC + cidx] : 0.f;

    for (int i = nst - 1; i >= 0; --i) {
        const sn_sss_stage st = stage_of(stages, n, dir, dir == 0 ? c.k_begin + i : c.k_end - 1 - i);
        const int local = col - st.in_off;
        const bool mine = (i == my_i);
        const bool before = is_in ? (my_i >= 0 && my_i < i) : true;
        const bool support = active && (is_in ? (dir == 0 ? (before || mine) : before) : true);
        const int rbase = st.out_off - c.row0;
        float g[SOUT_MAX];
#pragma unroll
        for (int r = 0; r < SOUT_MAX; ++r) g[r] = (support && r < st.out_dim) ? dM[(rbase + r) * DMC + cidx] : 0.f;
#pragma unroll
        for (int a = 0; a < DS; ++a) v[a] = scr[(i * DS + a) * BUILD_THREADS + t];
        __syncthreads();   // previous iteration's dot products are done with X / V
#pragma unroll
        for (int b = 0; b < DS; ++b) X[b * BB_LD + t] = lam[b];
#pragma unroll
        for (int r = 0; r < SOUT_MAX; ++r) X[(DS + r) * BB_LD + t] = g[r];
#pragma unroll
        for (int a = 0; a < DS; ++a) V[a * BB_LD + t] = v[a];
        __syncthreads();
        // d_ss[b][a] = sum_t lam_t[b] v_t[a] ;  d_ys[r][a] = sum_t g_t[r] v_t[a]
        for (int o = t; o < (DS + SOUT_MAX) * DS; o += BUILD_THREADS) {
            const int rowi = o / DS, a = o % DS;
            const bool ok = a < st.d_in && (rowi < DS ? rowi < st.d_out : (rowi - DS) < st.out_dim);
            if (ok) {
                const float* xr = X + rowi * BB_LD;
                const float* vr = V + a * BB_LD;
                float acc = 0.f;
                for (int tt = 0; tt < ncol_active; ++tt) acc = fmaf(xr[tt], vr[tt], acc);
                if (rowi < DS) gparams[st.off_ss + rowi * st.d_in + a] += acc;
                else gparams[st.off_ys + (rowi - DS) * st.d_in + a] += acc;
            }
        }
        if (mine) {
#pragma unroll
            for (int b = 0; b < DS; ++b)
                if (b < st.d_out) gparams[st.off_su + b * st.in_dim + local] += lam[b];
            if (st.off_yu >= 0) {
#pragma unroll
                for (int r = 0; r < SOUT_MAX; ++r)
                    if (r < st.out_dim) gparams[st.off_yu + r * st.in_dim + local] += g[r];
            }
        }
        // lambda entering the stage: ss^T lam + ys^T g
        float nl[DS];
#pragma unroll
        for (int a = 0; a < DS; ++a) nl[a] = 0.f;
#pragma unroll
        for (int b = 0; b < DS; ++b) {
            if (b < st.d_out) {
                const float* ss = params + st.off_ss + b * st.d_in;
#pragma unroll
                for (int a = 0; a < DS; ++a)
                    if (a < st.d_in) nl[a] = fmaf(__ldg(ss + a), lam[b], nl[a]);
            }
        }
#pragma unroll
        for (int r = 0; r < SOUT_MAX; ++r) {
            if (r < st.out_dim) {
                const float* ys = params + st.off_ys + r * st.d_in;
#pragma unroll
                for (int a = 0; a < DS; ++a)
                    if (a < st.d_in) nl[a] = fmaf(__ldg(ys + a), g[r], nl[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < DS; ++a) lam[a] = nl[a];
    }
}

int check_tc_plan(const sn_sss_tc_plan* p) {
    SN_CHECK_ARG(p != nullptr, "sss_tc: NULL plan");
    SN_CHECK_ARG(p->nb_states > 0 && p->input_dim > 0 && p->output_dim > 0 && p->nchunks > 0, "sss_tc: non-positive plan dimension");
    SN_CHECK_ARG(p->stages != nullptr && p->chunks != nullptr, "sss_tc: plan tables missing");
    SN_CHECK_ARG(p->input_dim % 4 == 0, "sss_tc: input_dim must be a multiple of 4 (TMA row pitch)");
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

extern "C" {

size_t sn_sss_tc_coef_floats(const sn_sss_tc_plan* p) {
    if (p == nullptr) return 0;
    return (size_t)p->nchunks * (WROWS * WCOLS + SCF);
}
size_t sn_sss_tc_rbuf_floats(const sn_sss_tc_plan* p, int64_t B) { return p == nullptr || B <= 0 ? 0 : (size_t)p->nchunks * B * 64; }
size_t sn_sss_tc_states_floats(const sn_sss_tc_plan* p, int64_t B) { return p == nullptr || B <= 0 ? 0 : (size_t)p->nchunks * B * 32; }
size_t sn_sss_tc_backward_workspace_floats(const sn_sss_tc_plan* p, int64_t B) {
    if (p == nullptr || B <= 0) return 0;
    return (size_t)p->nchunks * B * 32 + (size_t)p->nchunks * 64 * DMC + (size_t)p->nchunks * 2 * LMAX * DS * BUILD_THREADS;
}

int sn_sss_tc_build(const sn_sss_tc_plan* p, const float* params, float* coef, sn_stream_t stream) {
    if (int rc = check_tc_plan(p)) return rc;
    SN_CHECK_ARG(params && coef, "sss_tc_build: NULL buffer");
    float* W = coef;
    float* SC = coef + (size_t)p->nchunks * WROWS * WCOLS;
    sss_tc_build_kernel<<<dim3(p->nchunks, 2), BUILD_THREADS, 0, snb::as_stream(stream)>>>(p->stages, p->nb_states, p->chunks, params, W, SC);
    SN_CHECK_LAUNCH("sss_tc_build_kernel");
    return 0;
}

int sn_sss_tc_forward(const sn_sss_tc_plan* p, const float* coef, const float* x, int64_t ldx, float* y, int64_t ldy, const float* bias,
                      float* rbuf, float* states, int64_t B, sn_stream_t stream) {
    if (int rc = check_tc_plan(p)) return rc;
    SN_CHECK_ARG(coef && x && y && rbuf && states, "sss_tc_forward: NULL buffer");
    SN_CHECK_ARG(ldx >= p->input_dim && ldy >= p->output_dim, "sss_tc_forward: leading dimension too small");
    if (B <= 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    const float* W = coef;
    const float* SC = coef + (size_t)p->nchunks * WROWS * WCOLS;
    CUtensorMap mx, mw;
    if (int rc = make_map_f32(&mx, x, (uint64_t)p->input_dim, (uint64_t)B, (uint64_t)ldx, 128)) return rc;
    if (int rc = make_map_f32(&mw, W, (uint64_t)WCOLS, (uint64_t)p->nchunks * WROWS, (uint64_t)WCOLS, 128)) return rc;
    const int ntiles = (int)((B + 127) / 128);
    const long total = (long)ntiles * p->nchunks;
    const int grid = (int)(total < sm_count() ? total : sm_count());
    SN_CHECK_CUDA(cudaFuncSetAttribute(sss_tc_local_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G1_SMEM));
    sss_tc_local_gemm_kernel<<<grid, G1_THREADS, G1_SMEM, st>>>(mx, mw, p->chunks, p->nchunks, (long)B, ntiles, rbuf);
    SN_CHECK_LAUNCH("sss_tc_local_gemm_kernel");
    const int aligned = ((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (ldy & 3) == 0 && (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0) &&
                         p->rows_aligned) ? 1 : 0;
    sss_tc_scan_fwd_kernel<<<(unsigned)((B + SCAN_THREADS - 1) / SCAN_THREADS), SCAN_THREADS, 0, st>>>(p->chunks, p->nchunks, SC, rbuf, states, y, (long)ldy,
                                                                                                      bias, (long)B, aligned);
    SN_CHECK_LAUNCH("sss_tc_scan_fwd_kernel");
    return 0;
}

int sn_sss_tc_backward(const sn_sss_tc_plan* p, const float* params, const float* coef, const float* x, int64_t ldx, const float* grad_y,
                       int64_t ldgy, const float* states, float* workspace, float* grad_params, float* grad_bias, int64_t B, sn_stream_t stream) {
    if (int rc = check_tc_plan(p)) return rc;
    SN_CHECK_ARG(params && coef && x && grad_y && states && workspace && grad_params, "sss_tc_backward: NULL buffer");
    SN_CHECK_ARG(ldx >= p->input_dim && ldgy >= p->output_dim, "sss_tc_backward: leading dimension too small");
    SN_CHECK_ARG((ldgy & 3) == 0 && (reinterpret_cast<uintptr_t>(grad_y) & 15) == 0, "sss_tc_backward: grad_y must be 16-byte aligned with a row pitch multiple of 4");
    if (B <= 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    const float* SC = coef + (size_t)p->nchunks * WROWS * WCOLS;
    float* L = workspace;
    float* dM = L + (size_t)p->nchunks * B * 32;
    float* scratch = dM + (size_t)p->nchunks * 64 * DMC;
    const int aligned = p->rows_aligned ? 1 : 0;
    sss_tc_scan_bwd_kernel<<<(unsigned)((B + SCAN_THREADS - 1) / SCAN_THREADS), SCAN_THREADS, 0, st>>>(p->chunks, p->nchunks, SC, grad_y, (long)ldgy, L,
                                                                                                      grad_bias, (long)B, aligned);
    SN_CHECK_LAUNCH("sss_tc_scan_bwd_kernel");
    SN_CHECK_CUDA(cudaMemsetAsync(dM, 0, (size_t)p->nchunks * 64 * DMC * sizeof(float), st));
    CUtensorMap mx, mg, ml, ms;
    if (int rc = make_map_f32(&mx, x, (uint64_t)p->input_dim, (uint64_t)B, (uint64_t)ldx, G2_KS, 0, 0, true)) return rc;
    if (int rc = make_map_f32(&mg, grad_y, (uint64_t)p->output_dim, (uint64_t)B, (uint64_t)ldgy, G2_KS, 0, 0, true)) return rc;
    if (int rc = make_map_f32(&ml, L, 32, (uint64_t)B, 32, G2_KS, (uint64_t)p->nchunks, (uint64_t)B * 32, true)) return rc;
    if (int rc = make_map_f32(&ms, states, 32, (uint64_t)B, 32, G2_KS, (uint64_t)p->nchunks, (uint64_t)B * 32, true)) return rc;
    const int ntk = (int)((B + G2_KS - 1) / G2_KS);
    int nsplit = ceil_div(2 * sm_count(), p->nchunks);
    if (nsplit > ntk) nsplit = ntk;
    if (nsplit < 1) nsplit = 1;
    SN_CHECK_CUDA(cudaFuncSetAttribute(sss_tc_grad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
    sss_tc_grad_gemm_kernel<<<dim3(nsplit, p->nchunks), G2_THREADS, G2_SMEM, st>>>(mx, mg, ml, ms, p->chunks, (long)B, dM);
    SN_CHECK_LAUNCH("sss_tc_grad_gemm_kernel");
    sss_tc_build_bwd_kernel<<<dim3(p->nchunks, 2), BUILD_THREADS, 0, st>>>(p->stages, p->nb_states, p->chunks, params, dM, scratch, grad_params);
    SN_CHECK_LAUNCH("sss_tc_build_bwd_kernel");
    return 0;
}

}  // extern "C"
