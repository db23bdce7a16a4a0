// SSS (sequentially semiseparable / time-varying state-space) layer: forward and backward.
//
// Replaces SSSLayer.forward (reference layers/sss_layer.py:99-131) and the autograd graph torch
// builds from it.  Math per sample:
//   causal     k = 0..n-1 :  y_k  = C_k x_k + D_k u_k ;  x_{k+1} = A_k x_k + B_k u_k
//   anticausal k = n-1..0 :  y_k += G_k z_{k+1}       ;  z_k     = E_k z_{k+1} + F_k u_k
//
// Data layout.  Both sweeps are expressed as the same "stage step"
//       [ s_out ; y ] = P_k [ s_in ; u_k ]            P_k = [ A B ; C D ]  or  [ E F ; G 0 ]
// on a per-direction stage table (sn_sss_stage).  sn_sss_pack() re-lays the flat parameter buffer
// out per stage as  Pt[i][rows_pad]  (input-major: the forward / recompute loops read one float4
// of 4 consecutive output rows per input i)  followed by  P[r][k_pad]  (row-major: the adjoint loop
// reads one float4 of 4 consecutive inputs per output row r).  Rows are ordered state rows first,
// then y rows, zero padded.
//
// Thread mapping (SIMT fp32; the arithmetic intensity of this layer, ~56 flop/B, is above the
// fp32 ridge, so the FMA pipe is the binding roof -- DESIGN.md).  Per-sample vectors live in shared
// memory FEATURE-MAJOR ([feature][sample]), so a lane reads 4 samples of one feature with one
// LDS.128 and 4 rows of one input with one broadcast float4, giving 16 FMAs per 2 loads:
//   lane = (row-group slot, sample quad);  a warp owns NSW = 4*NQ samples of ONE direction and walks
//   all stages; the state is exchanged between lanes through a ping-pong buffer + __syncwarp().
//
// y is produced by two sweeps that meet in the middle: the first visitor of a stage stores
// (bias + part), the second one adds to it after a block barrier (deterministic, no atomics).
#include "common.cuh"

namespace {

using snb::ceil_div;
using snb::round_up;

struct Geom {
    int rgs;    // row-group slots per warp
    int nq;     // sample quads per warp
    int nsw;    // samples per warp = 4*nq
    int nswp;   // padded row stride (floats) of per-warp buffers
};

// row stride such that consecutive rows start 4,12,20 or 28 banks apart (conflict-free float4 access
// for up to 8 consecutive rows)
__host__ __device__ inline int pad_stride(int n) {
    int s = round_up(n, 4);
    while (((s & 31) & 7) != 4) s += 4;
    return s;
}

inline Geom make_geom(int rows_pad) {
    Geom g;
    int nrg = rows_pad / 4;
    g.rgs = nrg < 8 ? nrg : 8;
    if (g.rgs < 1) g.rgs = 1;
    g.nq = 32 / g.rgs;
    g.nsw = 4 * g.nq;
    g.nswp = pad_stride(g.nsw);
    return g;
}

// ------------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------------
__global__ void sss_pack_kernel(const sn_sss_stage* __restrict__ stages, int total_stages, int RP, int KP,
                                const float* __restrict__ params, float* __restrict__ packed) {
    int sidx = blockIdx.x;
    if (sidx >= total_stages) return;
    sn_sss_stage st = stages[sidx];
    const int K = st.d_in + st.in_dim;
    const int rows = st.d_out + st.out_dim;
    float* Pt = packed + st.pack_off;
    float* P = Pt + (size_t)KP * RP;
    for (int e = threadIdx.x; e < KP * RP; e += blockDim.x) {
        int i = e / RP, r = e - i * RP;
        float v = 0.f;
        if (i < K && r < rows) {
            if (r < st.d_out) {
                v = (i < st.d_in) ? params[st.off_ss + r * st.d_in + i] : params[st.off_su + r * st.in_dim + (i - st.d_in)];
            } else {
                int ry = r - st.d_out;
                if (i < st.d_in) v = params[st.off_ys + ry * st.d_in + i];
                else if (st.off_yu >= 0) v = params[st.off_yu + ry * st.in_dim + (i - st.d_in)];
            }
        }
        Pt[e] = v;
        P[(size_t)r * KP + i] = v;
    }
}

__device__ __forceinline__ void fma16(float (&acc)[4][4], const float4& p, const float4& v) {
    acc[0][0] = fmaf(p.x, v.x, acc[0][0]); acc[0][1] = fmaf(p.x, v.y, acc[0][1]);
    acc[0][2] = fmaf(p.x, v.z, acc[0][2]); acc[0][3] = fmaf(p.x, v.w, acc[0][3]);
    acc[1][0] = fmaf(p.y, v.x, acc[1][0]); acc[1][1] = fmaf(p.y, v.y, acc[1][1]);
    acc[1][2] = fmaf(p.y, v.z, acc[1][2]); acc[1][3] = fmaf(p.y, v.w, acc[1][3]);
    acc[2][0] = fmaf(p.z, v.x, acc[2][0]); acc[2][1] = fmaf(p.z, v.y, acc[2][1]);
    acc[2][2] = fmaf(p.z, v.z, acc[2][2]); acc[2][3] = fmaf(p.z, v.w, acc[2][3]);
    acc[3][0] = fmaf(p.w, v.x, acc[3][0]); acc[3][1] = fmaf(p.w, v.y, acc[3][1]);
    acc[3][2] = fmaf(p.w, v.z, acc[3][2]); acc[3][3] = fmaf(p.w, v.w, acc[3][3]);
}

// One stage step for the lanes of one warp:  out rows [4*rg, 4*rg+4) for rg in [rg_slot, nrg) step rgs.
//   state rows -> xn (next state), y rows -> yc (chunk output buffer, may be null: rows skipped)
__device__ __forceinline__ void stage_step(const sn_sss_stage& st, const float* __restrict__ packed, int RP,
                                           const float* xs, float* xn, const float* uc, float* yc,
                                           int stride, int ucol, int yrow, int nrg, int rg_slot, int rgs, int q) {
    const float4* Pt4 = reinterpret_cast<const float4*>(packed + st.pack_off);
    const int RP4 = RP >> 2;
    for (int rg = rg_slot; rg < nrg; rg += rgs) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        const float4* prow = Pt4 + rg;
        const float* xv = xs + 4 * q;
#pragma unroll 4
        for (int i = 0; i < st.d_in; ++i) {
            float4 p = __ldg(prow + (size_t)i * RP4);
            float4 v = *reinterpret_cast<const float4*>(xv + i * stride);
            fma16(acc, p, v);
        }
        prow += (size_t)st.d_in * RP4;
        const float* uv = uc + (size_t)ucol * stride + 4 * q;
#pragma unroll 4
        for (int i = 0; i < st.in_dim; ++i) {
            float4 p = __ldg(prow + (size_t)i * RP4);
            float4 v = *reinterpret_cast<const float4*>(uv + i * stride);
            fma16(acc, p, v);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r = 4 * rg + j;
            float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            if (r < st.d_out) {
                *reinterpret_cast<float4*>(xn + r * stride + 4 * q) = o;
            } else if (yc != nullptr && r < st.d_out + st.out_dim) {
                *reinterpret_cast<float4*>(yc + (size_t)(yrow + r - st.d_out) * stride + 4 * q) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
constexpr int FWD_PAIRS = 2;  // (causal, anticausal) warp pairs per CTA

__global__ void __launch_bounds__(FWD_PAIRS * 64)
sss_fwd_kernel(sn_sss_plan plan, const float* __restrict__ packed, const float* __restrict__ x, long ldx,
               float* __restrict__ y, long ldy, const float* __restrict__ bias, float* __restrict__ ckpt, long B,
               Geom g, int per_warp_floats) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp >> 1, dir = warp & 1;
    const int n = plan.nb_states, DP = plan.d_pad, RP = plan.rows_pad;
    const long s0 = ((long)blockIdx.x * FWD_PAIRS + pair) * g.nsw;
    float* wbase = smem + (size_t)warp * per_warp_floats;
    float* xs = wbase;
    float* xn = xs + (size_t)DP * g.nswp;
    float* uc = xn + (size_t)DP * g.nswp;
    float* yc = uc + (size_t)plan.chunk_in_max * g.nswp;
    const sn_sss_stage* stages = plan.stages + (size_t)dir * n;
    const sn_sss_chunk* chunks = plan.chunks + (size_t)dir * plan.nchunks;
    const int rg_slot = lane / g.nq, q = lane - rg_slot * g.nq;
    const bool active = rg_slot < g.rgs;
    bool did_mid = false;

    for (int ch = 0; ch < plan.nchunks; ++ch) {
        const sn_sss_chunk c = chunks[ch];
        if (c.second_visit && !did_mid) {
            __syncthreads();  // the other direction's first-visit stores to y are now visible
            did_mid = true;
        }
        if (ckpt != nullptr) {
            const int d = stages[c.kk_begin].d_in;
            float* cbase = ckpt + ((size_t)(dir * plan.nchunks + ch) * DP) * B;
            for (int e = lane; e < d * g.nsw; e += 32) {
                int f = e / g.nsw, s = e - f * g.nsw;
                if (s0 + s < B) cbase[(size_t)f * B + s0 + s] = xs[f * g.nswp + s];
            }
        }
        for (int e = lane; e < g.nsw * c.ncols; e += 32) {
            int s = e / c.ncols, cc = e - s * c.ncols;
            float v = 0.f;
            if (s0 + s < B) v = __ldg(x + (size_t)(s0 + s) * ldx + c.col0 + cc);
            uc[cc * g.nswp + s] = v;
        }
        __syncwarp();
        for (int kk = c.kk_begin; kk < c.kk_end; ++kk) {
            const sn_sss_stage st = stages[kk];
            if (active) {
                const int nrg = ceil_div(st.d_out + st.out_dim, 4);
                stage_step(st, packed, RP, xs, xn, uc, yc, g.nswp, st.in_off - c.col0, st.out_off - c.row0, nrg,
                           rg_slot, g.rgs, q);
            }
            __syncwarp();
            float* t = xs; xs = xn; xn = t;
        }
        for (int e = lane; e < g.nsw * c.nrows; e += 32) {
            int s = e / c.nrows, rr = e - s * c.nrows;
            if (s0 + s < B) {
                float v = yc[rr * g.nswp + s];
                float* dst = y + (size_t)(s0 + s) * ldy + c.row0 + rr;
                if (c.second_visit) v += __ldcg(dst);
                else if (bias != nullptr) v += __ldg(bias + c.row0 + rr);
                *dst = v;
            }
        }
        __syncwarp();
    }
    if (!did_mid) __syncthreads();
}

// ------------------------------------------------------------------------------------------
// backward: per (sample tile, direction) CTA; chunks visited in reverse processing order
//   1. recompute the chunk's entry states from the forward checkpoint          (warp-own samples)
//   2. adjoint sweep  lam_in = P^T [lam_out ; gy]                               (warp-own samples)
//   3. parameter gradients  dP_k = [lam_out ; gy] [s_in ; u]^T summed over the tile's samples
//      (warp-own stages, 4x4 register tiles over interleaved rows/cols), atomically added to the
//      flat gradient buffer at the parameters' natural offsets
// ------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = 2;

struct BwdSmem {
    int nsp;        // padded sample stride of the CTA-wide buffers
    int xh, lh, lcar, uc, gc, zero, total;  // offsets in floats
};

inline BwdSmem make_bwd_smem(const sn_sss_plan& p, const Geom& g) {
    BwdSmem s;
    int ns = BWD_WARPS * g.nsw;
    s.nsp = pad_stride(ns);
    int o = 0;
    s.xh = o;   o += p.chunk_len_max * p.d_pad * s.nsp;
    s.lh = o;   o += p.chunk_len_max * p.d_pad * s.nsp;
    s.lcar = o; o += p.d_pad * s.nsp;
    s.uc = o;   o += p.chunk_in_max * s.nsp;
    s.gc = o;   o += p.chunk_out_max * s.nsp;
    s.zero = o; o += s.nsp;
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(BWD_WARPS * 32)
sss_bwd_kernel(sn_sss_plan plan, const float* __restrict__ packed, const float* __restrict__ x, long ldx,
               const float* __restrict__ gy, long ldgy, const float* __restrict__ ckpt,
               float* __restrict__ gparams, float* __restrict__ gbias, long B, Geom g, BwdSmem sm) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int n = plan.nb_states, DP = plan.d_pad, RP = plan.rows_pad, KP = plan.k_pad;
    const int NS = BWD_WARPS * g.nsw, NSP = sm.nsp;
    const long t0 = (long)blockIdx.x * NS;      // first sample of the tile
    const int w0 = warp * g.nsw;                // this warp's first column inside the tile
    float* xh = smem + sm.xh;
    float* lh = smem + sm.lh;
    float* lcar = smem + sm.lcar;
    float* uc = smem + sm.uc;
    float* gc = smem + sm.gc;
    float* zr = smem + sm.zero;
    const sn_sss_stage* stages = plan.stages + (size_t)dir * n;
    const sn_sss_chunk* chunks = plan.chunks + (size_t)dir * plan.nchunks;
    const int rg_slot = lane / g.nq, q = lane - rg_slot * g.nq;
    const int rgs_state = (DP / 4 < g.rgs) ? DP / 4 : g.rgs;   // slots used by state-only loops
    const size_t hstride = (size_t)DP * NSP;                    // one history entry

    for (int e = threadIdx.x; e < NSP; e += blockDim.x) zr[e] = 0.f;
    for (int e = threadIdx.x; e < DP * NSP; e += blockDim.x) lcar[e] = 0.f;
    __syncthreads();

    for (int ch = plan.nchunks - 1; ch >= 0; --ch) {
        const sn_sss_chunk c = chunks[ch];
        const int len = c.kk_end - c.kk_begin;
        // ---- loads (warp-own samples) --------------------------------------------------
        {
            const int d = stages[c.kk_begin].d_in;
            const float* cbase = ckpt + ((size_t)(dir * plan.nchunks + ch) * DP) * B;
            for (int e = lane; e < d * g.nsw; e += 32) {
                int f = e / g.nsw, s = e - f * g.nsw;
                float v = 0.f;
                if (t0 + w0 + s < B) v = __ldg(cbase + (size_t)f * B + t0 + w0 + s);
                xh[f * NSP + w0 + s] = v;
            }
            for (int e = lane; e < g.nsw * c.ncols; e += 32) {
                int s = e / c.ncols, cc = e - s * c.ncols;
                float v = 0.f;
                if (t0 + w0 + s < B) v = __ldg(x + (size_t)(t0 + w0 + s) * ldx + c.col0 + cc);
                uc[cc * NSP + w0 + s] = v;
            }
            for (int e = lane; e < g.nsw * c.nrows; e += 32) {
                int s = e / c.nrows, rr = e - s * c.nrows;
                float v = 0.f;
                if (t0 + w0 + s < B) v = __ldg(gy + (size_t)(t0 + w0 + s) * ldgy + c.row0 + rr);
                gc[rr * NSP + w0 + s] = v;
            }
            // adjoint of the state leaving the chunk's last stage = carry from the later chunk
            const int dl = stages[c.kk_end - 1].d_out;
            for (int e = lane; e < dl * g.nsw; e += 32) {
                int f = e / g.nsw, s = e - f * g.nsw;
                lh[(size_t)(len - 1) * hstride + f * NSP + w0 + s] = lcar[f * NSP + w0 + s];
            }
        }
        __syncwarp();
        // ---- 1. recompute entry states of stages 1..len-1 ------------------------------
        for (int j = 0; j + 1 < len; ++j) {
            const sn_sss_stage st = stages[c.kk_begin + j];
            if (rg_slot < rgs_state) {
                stage_step(st, packed, RP, xh + j * hstride + w0, xh + (j + 1) * hstride + w0, uc + w0, nullptr, NSP,
                           st.in_off - c.col0, 0, ceil_div(st.d_out, 4), rg_slot, rgs_state, q);
            }
            __syncwarp();
        }
        // ---- 2. adjoint sweep -----------------------------------------------------------
        for (int j = len - 1; j >= 0; --j) {
            const sn_sss_stage st = stages[c.kk_begin + j];
            const float4* P4 = reinterpret_cast<const float4*>(packed + st.pack_off + (size_t)KP * RP);
            const int KP4 = KP >> 2;
            const float* lout = lh + j * hstride + w0;                          // adjoint of s_out
            const float* gyj = gc + (size_t)(st.out_off - c.row0) * NSP + w0;   // adjoint of y_k
            float* lin = (j > 0) ? (lh + (j - 1) * hstride + w0) : (lcar + w0);
            const int nig = ceil_div(st.d_in, 4);
            if (rg_slot < rgs_state) {
                for (int ig = rg_slot; ig < nig; ig += rgs_state) {
                    float acc[4][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
                    for (int r = 0; r < st.d_out; ++r) {
                        float4 p = __ldg(P4 + (size_t)r * KP4 + ig);
                        float4 v = *reinterpret_cast<const float4*>(lout + r * NSP + 4 * q);
                        fma16(acc, p, v);
                    }
                    for (int r = 0; r < st.out_dim; ++r) {
                        float4 p = __ldg(P4 + (size_t)(st.d_out + r) * KP4 + ig);
                        float4 v = *reinterpret_cast<const float4*>(gyj + r * NSP + 4 * q);
                        fma16(acc, p, v);
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        int i = 4 * ig + jj;
                        if (i < st.d_in)
                            *reinterpret_cast<float4*>(lin + i * NSP + 4 * q) =
                                make_float4(acc[jj][0], acc[jj][1], acc[jj][2], acc[jj][3]);
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- 3. parameter gradients (warp-own stages, all samples of the tile) -------------
        if (gbias != nullptr && dir == 0) {
            for (int rr = threadIdx.x; rr < c.nrows; rr += blockDim.x) {
                float s = 0.f;
                for (int e = 0; e < NS; ++e) s += gc[rr * NSP + e];
                atomicAdd(gbias + c.row0 + rr, s);
            }
        }
        for (int j = warp; j < len; j += BWD_WARPS) {
            const sn_sss_stage st = stages[c.kk_begin + j];
            const int rows = st.d_out + st.out_dim, K = st.d_in + st.in_dim;
            const int rgn = ceil_div(rows, 4), cgn = ceil_div(K, 4);
            const float* lout = lh + j * hstride;
            const float* gyj = gc + (size_t)(st.out_off - c.row0) * NSP;
            const float* xin = xh + j * hstride;
            const float* uj = uc + (size_t)(st.in_off - c.col0) * NSP;
            for (int t = lane; t < rgn * cgn; t += 32) {
                const int rg = t / cgn, cg = t - rg * cgn;
                const float* gp[4];
                const float* ip[4];
                int rr[4], ii[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    int r = rg + a * rgn;
                    rr[a] = r;
                    gp[a] = (r < st.d_out) ? (lout + r * NSP) : (r < rows ? gyj + (r - st.d_out) * NSP : zr);
                    int i = cg + a * cgn;
                    ii[a] = i;
                    ip[a] = (i < st.d_in) ? (xin + i * NSP) : (i < K ? uj + (i - st.d_in) * NSP : zr);
                }
                float acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
                for (int sq = 0; sq < NS; sq += 4) {
                    float4 gv[4], iv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        gv[a] = *reinterpret_cast<const float4*>(gp[a] + sq);
                        iv[a] = *reinterpret_cast<const float4*>(ip[a] + sq);
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            acc[a][b] = fmaf(gv[a].x, iv[b].x, acc[a][b]);
                            acc[a][b] = fmaf(gv[a].y, iv[b].y, acc[a][b]);
                            acc[a][b] = fmaf(gv[a].z, iv[b].z, acc[a][b]);
                            acc[a][b] = fmaf(gv[a].w, iv[b].w, acc[a][b]);
                        }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = rr[a];
                    if (r >= rows) continue;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int i = ii[b];
                        if (i >= K) continue;
                        int off;
                        if (r < st.d_out) {
                            off = (i < st.d_in) ? st.off_ss + r * st.d_in + i : st.off_su + r * st.in_dim + (i - st.d_in);
                        } else {
                            const int ry = r - st.d_out;
                            if (i < st.d_in) off = st.off_ys + ry * st.d_in + i;
                            else if (st.off_yu >= 0) off = st.off_yu + ry * st.in_dim + (i - st.d_in);
                            else continue;
                        }
                        atomicAdd(gparams + off, acc[a][b]);
                    }
                }
            }
        }
        __syncthreads();
    }
}

int check_plan(const sn_sss_plan* p) {
    SN_CHECK_ARG(p != nullptr, "sss: plan is NULL");
    SN_CHECK_ARG(p->nb_states > 0 && p->input_dim > 0 && p->output_dim > 0, "sss: bad plan dims");
    SN_CHECK_ARG(p->rows_pad > 0 && p->rows_pad % 4 == 0 && p->k_pad > 0 && p->k_pad % 4 == 0 && p->d_pad % 4 == 0 && p->d_pad > 0,
                 "sss: rows_pad/k_pad/d_pad must be positive multiples of 4");
    SN_CHECK_ARG(p->stages != nullptr && p->chunks != nullptr && p->nchunks > 0, "sss: plan tables missing");
    return 0;
}

}  // namespace

extern "C" {

size_t sn_sss_packed_floats(const sn_sss_plan* p) {
    if (p == nullptr) return 0;
    return (size_t)2 * p->nb_states * 2 * p->k_pad * p->rows_pad;
}

size_t sn_sss_ckpt_floats(const sn_sss_plan* p, int64_t B) {
    if (p == nullptr || B <= 0) return 0;
    return (size_t)2 * p->nchunks * p->d_pad * (size_t)B;
}

int sn_sss_pack(const sn_sss_plan* p, const float* params, float* packed, sn_stream_t stream) {
    if (int rc = check_plan(p)) return rc;
    SN_CHECK_ARG(params != nullptr && packed != nullptr, "sss_pack: NULL buffer");
    sss_pack_kernel<<<2 * p->nb_states, 128, 0, snb::as_stream(stream)>>>(p->stages, 2 * p->nb_states, p->rows_pad,
                                                                           p->k_pad, params, packed);
    SN_CHECK_LAUNCH("sss_pack_kernel");
    return 0;
}

int sn_sss_forward(const sn_sss_plan* p, const float* packed, const float* x, int64_t ldx, float* y, int64_t ldy,
                   const float* bias, float* ckpt, int64_t B, sn_stream_t stream) {
    if (int rc = check_plan(p)) return rc;
    SN_CHECK_ARG(packed && x && y, "sss_forward: NULL buffer");
    SN_CHECK_ARG(ldx >= p->input_dim && ldy >= p->output_dim, "sss_forward: leading dimension too small");
    if (B <= 0) return 0;
    Geom g = make_geom(p->rows_pad);
    int per_warp = (2 * p->d_pad + p->chunk_in_max + p->chunk_out_max) * g.nswp;
    size_t smem = (size_t)per_warp * FWD_PAIRS * 2 * sizeof(float);
    SN_CHECK_ARG(smem <= 227 * 1024, "sss_forward: stage dims need %zu bytes of shared memory (> 227 KB)", smem);
    static size_t configured = 0;
    if (smem > configured) {
        SN_CHECK_CUDA(cudaFuncSetAttribute(sss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    long tile = (long)FWD_PAIRS * g.nsw;
    unsigned grid = (unsigned)((B + tile - 1) / tile);
    sss_fwd_kernel<<<grid, FWD_PAIRS * 64, smem, snb::as_stream(stream)>>>(*p, packed, x, (long)ldx, y, (long)ldy, bias,
                                                                           ckpt, (long)B, g, per_warp);
    SN_CHECK_LAUNCH("sss_fwd_kernel");
    return 0;
}

int sn_sss_backward(const sn_sss_plan* p, const float* packed, const float* x, int64_t ldx, const float* grad_y,
                    int64_t ldgy, const float* ckpt, float* grad_params, float* grad_bias, float* grad_x, int64_t ldgx,
                    int64_t B, sn_stream_t stream) {
    if (int rc = check_plan(p)) return rc;
    (void)ldgx;
    SN_CHECK_ARG(packed && x && grad_y && ckpt && grad_params, "sss_backward: NULL buffer");
    SN_CHECK_ARG(grad_x == nullptr, "sss_backward: grad_x is not implemented (the reference training loop never needs it)");
    SN_CHECK_ARG(ldx >= p->input_dim && ldgy >= p->output_dim, "sss_backward: leading dimension too small");
    if (B <= 0) return 0;
    Geom g = make_geom(p->rows_pad);
    BwdSmem sm = make_bwd_smem(*p, g);
    size_t smem = (size_t)sm.total * sizeof(float);
    SN_CHECK_ARG(smem <= 227 * 1024, "sss_backward: stage dims need %zu bytes of shared memory (> 227 KB)", smem);
    static size_t configured = 0;
    if (smem > configured) {
        SN_CHECK_CUDA(cudaFuncSetAttribute(sss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    long tile = (long)BWD_WARPS * g.nsw;
    dim3 grid((unsigned)((B + tile - 1) / tile), 2);
    sss_bwd_kernel<<<grid, BWD_WARPS * 32, smem, snb::as_stream(stream)>>>(*p, packed, x, (long)ldx, grad_y, (long)ldgy,
                                                                          ckpt, grad_params, grad_bias, (long)B, g, sm);
    SN_CHECK_LAUNCH("sss_bwd_kernel");
    return 0;
}

}  // extern "C"
