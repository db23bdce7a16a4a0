// fp32-accurate GEMM on the tensor cores (3xTF32: operands split into tf32 hi + lo parts on the fly, three tcgen05.mma per
// k-step, fp32 accumulation in TMEM):
//     D[M x N] = alpha * op(A) op(B) (+ bias[N])        or, split-K,  D += alpha * op(A) op(B)  (red.global.add)
// fp32 operands straight from global memory by TMA, in either layout:
//     A given as [M][K] (K-major)  or as [K][M] (MN-major);   B given as [N][K] (K-major) or as [K][N] (MN-major).
// K-major tiles use SWIZZLE_128B, MN-major ones SWIZZLE_128B_ATOM_32B (the only layout kind::tf32 takes for MN-major operands).
//   warp 0: TMA producer | warp 1: TMEM alloc + MMA issuer | warps 2-5: hi/lo converters (in place + lo tile) | warps 6-9: epilogue
// Replaces the SIMT gemm_f32_kernel wherever the operands are 16-byte aligned with row pitches that are multiples of 4 floats.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace snb {
namespace t3 {

using namespace snb::umma;

constexpr int T3_STAGES = 3;
constexpr int T3_CONV = 128 * T3_STAGES;            // converter threads: one group of four warps per pipeline stage
constexpr int T3_THREADS = 64 + T3_CONV + 128;
constexpr int T3_TILE = 128 * 128;                 // bytes: 128 rows x 32 floats (K-major) or 4 blocks of 32 k-rows x 32 floats (MN-major)
constexpr int T3_STAGE = 4 * T3_TILE;              // A hi | A lo | B hi | B lo
constexpr size_t T3_SMEM = (size_t)T3_STAGES * T3_STAGE + 1024 + 256;

struct T3Args {
    int M, N, K, kslice;
    float* D;
    long ldd;
    const float* bias;
    float alpha;
    int atomic;
    // Extents known only on the device (the LDR layer's Krylov series length): *kdev caps K (and the K slices of a split-K launch
    // are cut from the capped extent), CTAs whose rows start at or beyond *mdev leave at once.  NULL = use M / K as given.
    const int* kdev;
    const int* mdev;
};

// hi = round-to-nearest tf32 of x, lo = round-to-nearest tf32 of (x - hi).  The tensor core TRUNCATES the low 13 mantissa bits
// of its operands; an unrounded lo part would make every product err towards zero, a bias that grows linearly with K.
__device__ __forceinline__ float t3_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// lo = rn_tf32(x - trunc_tf32(x)): the raw fp32 tile serves as the hi operand because the tensor core truncates (verified on
// sm_100a: parity unchanged when the in-place hi store is dropped), so the converters only write the lo tile.
__device__ __forceinline__ float t3_lo(float x) { return t3_rn(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)); }
__device__ __forceinline__ void t3_lo4(const float4& v, float4& l) {
    l.x = t3_lo(v.x); l.y = t3_lo(v.y); l.z = t3_lo(v.z); l.w = t3_lo(v.w);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(T3_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, T3Args args) {
    extern __shared__ uint8_t t3_smem_raw[];
    uint8_t* smem = t3_smem_raw + ((1024u - (smem_u32(t3_smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + T3_STAGES * T3_STAGE);
    uint64_t* conv = full + T3_STAGES;
    uint64_t* empty = conv + T3_STAGES;
    uint64_t* acc_full = empty + T3_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    if (args.mdev != nullptr && m0 >= __ldg(args.mdev)) return;     // uniform over the CTA, before any barrier or TMEM allocation
    int Kl = args.K, kslice = args.kslice;
    if (args.kdev != nullptr) {
        Kl = min(Kl, __ldg(args.kdev));
        kslice = round_up(ceil_div(Kl > 0 ? Kl : 1, (int)gridDim.z), 32);
    }
    const int kbeg = blockIdx.z * kslice;
    const int kend = min(Kl, kbeg + kslice);
    const int nkb = kend > kbeg ? (kend - kbeg + 31) / 32 : 0;
    if (args.kdev != nullptr && args.atomic && nkb == 0) return;    // an empty K slice of an accumulating launch adds nothing
    // The tensor core adds products into its fp32 accumulator with truncation, an error that grows linearly with the length of
    // the sum; K is therefore spread over up to four TMEM accumulators that the epilogue adds with ordinary fp32 rounding.
    const int chunk = (nkb + 3) / 4 > 0 ? (nkb + 3) / 4 : 1;
    const int nacc = (nkb + chunk - 1) / chunk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T3_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(conv + s, 128); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T3_STAGES, round = kb / T3_STAGES;
                if (round > 0) mbar_wait(empty + s, (round - 1) & 1);
                uint8_t* st = smem + s * T3_STAGE;
                const int k0 = kbeg + kb * 32;
                mbar_expect_tx(full + s, 2 * T3_TILE);
                if (A_MN) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) tma_load_2d(st + i * 4096, &map_a, m0 + 32 * i, k0, full + s);
                } else {
                    tma_load_2d(st, &map_a, k0, m0, full + s);
                }
                if (B_MN) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) tma_load_2d(st + 2 * T3_TILE + i * 4096, &map_b, n0 + 32 * i, k0, full + s);
                } else {
                    tma_load_2d(st + 2 * T3_TILE, &map_b, k0, n0, full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_tf32(128, 128, A_MN, B_MN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T3_STAGES, round = kb / T3_STAGES;
                mbar_wait(conv + s, round & 1);
                tc_fence_after();
                uint8_t* st = smem + s * T3_STAGE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ah = A_MN ? desc_mnmajor_sw128_32b(st + k * 1024, 4096) : desc_kmajor_sw128(st) + 2 * k;
                    const uint64_t al = A_MN ? desc_mnmajor_sw128_32b(st + T3_TILE + k * 1024, 4096) : desc_kmajor_sw128(st + T3_TILE) + 2 * k;
                    const uint64_t bh = B_MN ? desc_mnmajor_sw128_32b(st + 2 * T3_TILE + k * 1024, 4096) : desc_kmajor_sw128(st + 2 * T3_TILE) + 2 * k;
                    const uint64_t bl = B_MN ? desc_mnmajor_sw128_32b(st + 3 * T3_TILE + k * 1024, 4096) : desc_kmajor_sw128(st + 3 * T3_TILE) + 2 * k;
                    const uint32_t acc = tmem + (uint32_t)(kb / chunk) * 128;
                    mma_tf32(acc, ah, bh, idesc, ((kb % chunk) | k) ? 1u : 0u);
                    mma_tf32(acc, ah, bl, idesc, 1u);
                    mma_tf32(acc, al, bh, idesc, 1u);
                }
                umma_commit(empty + s);
            }
            umma_commit(acc_full);
        }
    } else if (warp < 2 + T3_CONV / 32) {
        // A stage costs a converter warp a full LDS -> STS -> proxy-fence round trip (~3 000 cycles for the two tiles, against ~800
        // cycles of MMA): one group of four warps per stage keeps every stage in conversion at once.  A group owns its stage, so it
        // never skips a phase of that stage's `full` barrier.
        const int grp = (warp - 2) >> 2, ct = (threadIdx.x - 64) & 127;
        for (int kb = grp; kb < nkb; kb += T3_STAGES) {
            const int s = grp, round = kb / T3_STAGES;
            mbar_wait(full + s, round & 1);
            uint8_t* st = smem + s * T3_STAGE;
#pragma unroll
            for (int op = 0; op < 2; ++op) {
                float4* h = reinterpret_cast<float4*>(st + op * 2 * T3_TILE);
                float4* l = h + T3_TILE / 16;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = h[ct + 128 * i];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 ll;
                    t3_lo4(v[i], ll);
                    l[ct + 128 * i] = ll;
                }
            }
            fence_async_smem();
            mbar_arrive(conv + s);
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        if (nkb > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        const uint32_t acc = tmem + ((uint32_t)(q * 32) << 16);
        float* drow = args.D + (size_t)row * args.ldd + n0;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(args.D) & 15) == 0) && (args.ldd % 4 == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 16) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = 0.f;
            for (int a = 0; a < nacc; ++a) {
                uint32_t v[16];
                tmem_ld16_nowait(acc + a * 128 + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += __uint_as_float(v[j]);
            }
            if (row >= args.M || n0 + c0 >= args.N) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] *= args.alpha;
            const bool full16 = n0 + c0 + 16 <= args.N;
            if (args.atomic) {
                if (full16 && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + 4 * j), "f"(f[4 * j]), "f"(f[4 * j + 1]), "f"(f[4 * j + 2]),
                                     "f"(f[4 * j + 3])
                                     : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < args.N) atomicAdd(drow + c0 + j, f[j]);
                }
            } else {
                if (args.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < args.N) f[j] += __ldg(args.bias + n0 + c0 + j);
                }
                if (full16 && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(drow + c0 + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < args.N) drow[c0 + j] = f[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

inline bool t3_eligible(const float* A, long lda, const float* B, long ldb) {
    return (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 && lda % 4 == 0 && ldb % 4 == 0 && lda > 0 && ldb > 0;
}

// a_mn: A is stored [K][M]; b_mn: B is stored [K][N].  ksplit > 1 (or atomic) accumulates into D with reductions.
template <bool A_MN, bool B_MN>
inline int gemm_tf32x3_launch(int M, int N, int K, float alpha, const float* A, long lda, const float* B, long ldb, float* D, long ldd, const float* bias,
                              int ksplit, bool atomic, cudaStream_t stream, const int* kdev = nullptr, const int* mdev = nullptr) {
    CUtensorMap ma, mb;
    if (A_MN) { if (int rc = make_map_f32(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, 0, 0, true)) return rc; }
    else      { if (int rc = make_map_f32(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 128)) return rc; }
    if (B_MN) { if (int rc = make_map_f32(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, 0, 0, true)) return rc; }
    else      { if (int rc = make_map_f32(&mb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 128)) return rc; }
    T3Args args;
    args.M = M; args.N = N; args.K = K;
    if (ksplit < 1) ksplit = 1;
    args.kslice = round_up(ceil_div(K > 0 ? K : 1, ksplit), 32);
    args.D = D; args.ldd = ldd; args.bias = bias; args.alpha = alpha; args.atomic = atomic ? 1 : 0;
    args.kdev = kdev; args.mdev = mdev;
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128), ceil_div(K > 0 ? K : 1, args.kslice));
    SN_CHECK_ARG(atomic || grid.z == 1, "gemm_tf32x3: split-K needs the accumulating epilogue");
    SN_SET_MAX_SMEM((int)T3_SMEM, gemm_tf32x3_kernel<A_MN, B_MN>);
    SN_LAUNCH("gemm_tf32x3_kernel", stream, gemm_tf32x3_kernel<A_MN, B_MN><<<grid, T3_THREADS, T3_SMEM, stream>>>(ma, mb, args));
    return 0;
}

}  // namespace t3
}  // namespace snb
