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

constexpr int SSS_HDR = 8;   // floats at the head of every packed stage block: the stage descriptor

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
    int* hdr = reinterpret_cast<int*>(packed + st.pack_off);
    if (threadIdx.x < SSS_HDR) {
        const int h[SSS_HDR] = {st.in_off, st.in_dim, st.out_off, st.out_dim, st.d_in, st.d_out, st.k, 0};
        hdr[threadIdx.x] = h[threadIdx.x];
    }
    float* Pt = packed + st.pack_off + SSS_HDR;
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
    const float4* Pt4 = reinterpret_cast<const float4*>(packed + st.pack_off + SSS_HDR);
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
// async-copy / mbarrier helpers (sm_90+ PTX; SASS: SYNCS.*, UBLKCP, LDGSTS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ------------------------------------------------------------------------------------------
// forward
//   CTA = FWD_PAIRS (causal, anticausal) consumer warp pairs + 1 producer warp.
//   producer : streams the packed stage blocks (descriptor header + Pt) of both directions through a
//              FWD_SLOTS-deep shared-memory ring with cp.async.bulk + mbarrier (full/empty per slot).
//   consumer : owns NSW = 4*NQ samples of one direction; input columns of half a chunk are prefetched
//              with cp.async (double buffered, natural [sample][col] layout, row stride = 4*odd floats so
//              the 4 scalar reads of a lane's interleaved samples are bank-conflict free); the state lives in a
//              feature-major ping-pong buffer; y is collected per chunk and flushed with 64 B row segments.
//   Sample <-> position map inside a warp: position p = 4*q + t  holds local sample  q + NQ*t.
// ------------------------------------------------------------------------------------------
constexpr int FWD_PAIRS = 3;
constexpr int FWD_SLOTS = 4;
constexpr int FWD_CONSUMERS = 2 * FWD_PAIRS;
constexpr int FWD_THREADS = (FWD_CONSUMERS + 1) * 32;

struct FwdSmem {
    int slot_floats;   // one ring slot: header + Pt
    int ring_off, bar_off, warp_off, per_warp;   // float offsets (bar_off is 8-byte aligned)
    int uw;            // row stride of the u buffers
    int xs, xn, ub0, ub1, yb;   // offsets inside a warp's region
    int total;
};

inline FwdSmem make_fwd_smem(const sn_sss_plan& p, const Geom& g) {
    FwdSmem s;
    s.slot_floats = round_up(SSS_HDR + p.k_pad * p.rows_pad, 4);
    s.ring_off = 0;
    s.bar_off = s.ring_off + 2 * FWD_SLOTS * s.slot_floats;
    s.warp_off = s.bar_off + 2 * (2 * FWD_SLOTS * 2);   // 2 dirs x SLOTS x {full, empty} x 8 bytes
    s.uw = pad_stride(p.half_in_max + 3);
    int o = 0;
    s.xs = o;  o += p.d_pad * g.nswp;
    s.xn = o;  o += p.d_pad * g.nswp;
    s.ub0 = o; o += g.nsw * s.uw;
    s.ub1 = o; o += g.nsw * s.uw;
    s.yb = o;  o += p.chunk_out_max * g.nswp;
    s.per_warp = round_up(o, 4);
    s.total = s.warp_off + FWD_CONSUMERS * s.per_warp;
    return s;
}

// input columns [col0, col0+ncols) of the warp's samples -> ub[ls][uoff + (col - col0)]; returns uoff
__device__ __forceinline__ int issue_u_load(const float* __restrict__ x, long ldx, long s0, long B, int in_dim, int nsw, int uw,
                                            bool aligned, int col0, int ncols, float* ub, int lane) {
    if (aligned) {
        const int col0a = col0 & ~3;
        const int ngran = (col0 - col0a + ncols + 3) >> 2;
        for (int r0 = 0; r0 < nsw; r0 += 8) {
            const int ls = r0 + (lane >> 2);
            const bool rok = ls < nsw && s0 + ls < B;
            const float* row = x + (size_t)(s0 + ls) * ldx + col0a;
            for (int gi = lane & 3; gi < ngran; gi += 4) {
                const bool ok = rok && (col0a + 4 * gi < in_dim);
                if (ls < nsw) cp_async16(ub + ls * uw + 4 * gi, ok ? (const void*)(row + 4 * gi) : (const void*)x, ok);
            }
        }
        return col0 - col0a;
    }
    for (int ls = 0; ls < nsw; ++ls) {
        const bool rok = s0 + ls < B;
        const float* row = x + (size_t)(s0 + ls) * ldx + col0;
        for (int c = lane; c < ncols; c += 32) cp_async4(ub + ls * uw + c, rok ? (const void*)(row + c) : (const void*)x, rok);
    }
    return 0;
}

__global__ void __launch_bounds__(FWD_THREADS)
sss_fwd_kernel(sn_sss_plan plan, const float* __restrict__ packed, const float* __restrict__ x, long ldx,
               float* __restrict__ y, long ldy, const float* __restrict__ bias, float* __restrict__ ckpt, long B,
               Geom g, FwdSmem sm, int x_aligned) {
    extern __shared__ __align__(128) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = plan.nb_states, DP = plan.d_pad, RP = plan.rows_pad;
    float* ring = smem + sm.ring_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bar_off);   // [dir][slot][full, empty]
    auto full_bar = [&](int dir, int slot) { return bars + ((dir * FWD_SLOTS + slot) * 2 + 0); };
    auto empty_bar = [&](int dir, int slot) { return bars + ((dir * FWD_SLOTS + slot) * 2 + 1); };

    if (threadIdx.x == 0) {
        for (int d = 0; d < 2; ++d)
            for (int sl = 0; sl < FWD_SLOTS; ++sl) {
                mbar_init(full_bar(d, sl), 1);
                mbar_init(empty_bar(d, sl), FWD_PAIRS);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == FWD_CONSUMERS) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)sm.slot_floats * 4u;
            const size_t blk = SSS_HDR + 2 * (size_t)plan.k_pad * RP;
            for (int kk = 0; kk < n; ++kk) {
                const int sl = kk % FWD_SLOTS, round = kk / FWD_SLOTS;
                for (int d = 0; d < 2; ++d) {
                    if (round > 0) mbar_wait(empty_bar(d, sl), (round - 1) & 1);
                    mbar_arrive_expect_tx(full_bar(d, sl), bytes);
                    bulk_g2s(ring + (size_t)(d * FWD_SLOTS + sl) * sm.slot_floats, packed + ((size_t)d * n + kk) * blk, bytes, full_bar(d, sl));
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int pair = warp >> 1, dir = warp & 1;
    const long s0 = ((long)blockIdx.x * FWD_PAIRS + pair) * g.nsw;
    float* wbase = smem + sm.warp_off + (size_t)warp * sm.per_warp;
    float* xs = wbase + sm.xs;
    float* xn = wbase + sm.xn;
    float* ubuf[2] = {wbase + sm.ub0, wbase + sm.ub1};
    float* yb = wbase + sm.yb;
    const sn_sss_chunk* chunks = plan.chunks + (size_t)dir * plan.nchunks;
    const int rg_slot = lane / g.nq, q = lane - rg_slot * g.nq;
    const bool active = rg_slot < g.rgs;
    const int nswp = g.nswp, uw = sm.uw, nq = g.nq, nsw = g.nsw;
    const int RP4 = RP >> 2;
    bool did_mid = false;
    int ubi = 0;         // which u buffer holds the half that is consumed next
    int uoff_cur;

    sn_sss_chunk c = chunks[0];
    uoff_cur = issue_u_load(x, ldx, s0, B, plan.input_dim, nsw, uw, x_aligned != 0, c.col0_a, c.ncols_a, ubuf[0], lane);
    cp_async_commit();

    for (int ch = 0; ch < plan.nchunks; ++ch) {
        if (c.second_visit && !did_mid) {
            named_bar_sync(1, FWD_CONSUMERS * 32);   // the other direction's first-visit stores to y are visible
            did_mid = true;
        }
        sn_sss_chunk cnext = c;
        if (ch + 1 < plan.nchunks) cnext = chunks[ch + 1];
        for (int half = 0; half < 2; ++half) {
            const int kb = half == 0 ? c.kk_begin : c.kk_mid;
            const int ke = half == 0 ? c.kk_mid : c.kk_end;
            const int hcol0 = half == 0 ? c.col0_a : c.col0_b;
            // prefetch the next half (or the first half of the next chunk) into the other buffer
            int uoff_next = 0;
            bool issued = false;
            if (half == 0) {
                if (c.kk_mid < c.kk_end) { uoff_next = issue_u_load(x, ldx, s0, B, plan.input_dim, nsw, uw, x_aligned != 0, c.col0_b, c.ncols_b, ubuf[ubi ^ 1], lane); issued = true; }
            } else if (ch + 1 < plan.nchunks) {
                uoff_next = issue_u_load(x, ldx, s0, B, plan.input_dim, nsw, uw, x_aligned != 0, cnext.col0_a, cnext.ncols_a, ubuf[ubi ^ 1], lane); issued = true;
            }
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            if (kb < ke) {
                const float* ub = ubuf[ubi];
                for (int kk = kb; kk < ke; ++kk) {
                    const int sl = kk % FWD_SLOTS;
                    mbar_wait(full_bar(dir, sl), (kk / FWD_SLOTS) & 1);
                    const float* slot = ring + (size_t)(dir * FWD_SLOTS + sl) * sm.slot_floats;
                    const int* hdr = reinterpret_cast<const int*>(slot);
                    const int in_off = hdr[0], in_dim = hdr[1], out_off = hdr[2], out_dim = hdr[3], d_in = hdr[4], d_out = hdr[5];
                    if (ckpt != nullptr && kk == c.kk_begin) {
                        // checkpoint of the state entering the chunk, in true sample order
                        float* cbase = ckpt + ((size_t)(dir * plan.nchunks + ch) * DP) * B;
                        for (int e = lane; e < d_in * nsw; e += 32) {
                            const int f = e / nsw, ls = e - f * nsw;
                            const int pos = 4 * (ls % nq) + ls / nq;
                            if (s0 + ls < B) cbase[(size_t)f * B + s0 + ls] = xs[f * nswp + pos];
                        }
                    }
                    if (active) {
                        const int nrg = (d_out + out_dim + 3) >> 2;
                        const float4* Pt4 = reinterpret_cast<const float4*>(slot + SSS_HDR);
                        const float* u0 = ub + (q)*uw + uoff_cur + (in_off - hcol0);
                        const float* u1 = u0 + nq * uw;
                        const float* u2 = u1 + nq * uw;
                        const float* u3 = u2 + nq * uw;
                        for (int rg = rg_slot; rg < nrg; rg += g.rgs) {
                            float acc[4][4];
#pragma unroll
                            for (int a = 0; a < 4; ++a)
#pragma unroll
                                for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
                            const float4* prow = Pt4 + rg;
                            const float* xv = xs + 4 * q;
#pragma unroll 4
                            for (int i = 0; i < d_in; ++i) {
                                const float4 p = prow[0];
                                const float4 v = *reinterpret_cast<const float4*>(xv);
                                prow += RP4;
                                xv += nswp;
                                fma16(acc, p, v);
                            }
#pragma unroll 4
                            for (int i = 0; i < in_dim; ++i) {
                                const float4 p = prow[0];
                                prow += RP4;
                                const float4 v = make_float4(u0[i], u1[i], u2[i], u3[i]);
                                fma16(acc, p, v);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int r = 4 * rg + j;
                                const float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                                if (r < d_out) *reinterpret_cast<float4*>(xn + r * nswp + 4 * q) = o;
                                else if (r < d_out + out_dim) *reinterpret_cast<float4*>(yb + (out_off - c.row0 + r - d_out) * nswp + 4 * q) = o;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar(dir, sl));
                    float* t = xs; xs = xn; xn = t;
                }
            }
            if (issued) { ubi ^= 1; uoff_cur = uoff_next; }
        }
        // flush the chunk's outputs: y[s0 + ls][row0 + row]
        for (int ls = 0; ls < nsw; ++ls) {
            if (s0 + ls >= B) break;
            const int pos = 4 * (ls % nq) + ls / nq;
            float* dst = y + (size_t)(s0 + ls) * ldy + c.row0;
            for (int row = lane; row < c.nrows; row += 32) {
                float v = yb[row * nswp + pos];
                if (c.second_visit) v += __ldcg(dst + row);
                else if (bias != nullptr) v += __ldg(bias + c.row0 + row);
                dst[row] = v;
            }
        }
        __syncwarp();
        c = cnext;
    }
    cp_async_wait<0>();
    if (!did_mid) named_bar_sync(1, FWD_CONSUMERS * 32);
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
            const float4* P4 = reinterpret_cast<const float4*>(packed + st.pack_off + SSS_HDR + (size_t)KP * RP);
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
    return (size_t)2 * p->nb_states * (SSS_HDR + 2 * (size_t)p->k_pad * p->rows_pad);
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
    FwdSmem sm = make_fwd_smem(*p, g);
    size_t smem = (size_t)sm.total * sizeof(float);
    SN_CHECK_ARG(smem <= 227 * 1024, "sss_forward: stage dims need %zu bytes of shared memory (> 227 KB)", smem);
    static size_t configured = 0;
    if (smem > configured) {
        SN_CHECK_CUDA(cudaFuncSetAttribute(sss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // 16-byte cp.async path needs 16-byte aligned rows whose length is a multiple of 4 floats
    const int aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx & 3) == 0 && (p->input_dim & 3) == 0) ? 1 : 0;
    long tile = (long)FWD_PAIRS * g.nsw;
    unsigned grid = (unsigned)((B + tile - 1) / tile);
    sss_fwd_kernel<<<grid, FWD_THREADS, smem, snb::as_stream(stream)>>>(*p, packed, x, (long)ldx, y, (long)ldy, bias, ckpt, (long)B,
                                                                        g, sm, aligned);
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
