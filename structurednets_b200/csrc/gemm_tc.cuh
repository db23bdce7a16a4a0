// bf16 tensor-core GEMM building block on tcgen05 / TMEM / TMA (sm_100a):
//     D[M x N] (+)= A[M x K] * B[N x K]^T        A, B bf16, K-major (K contiguous); fp32 accumulation in TMEM
// One 128 x BN output tile per CTA, K walked in 64-element (128-byte) blocks through a STAGES-deep TMA pipeline.
//   warp 0      : TMA producer (one elected lane): cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier complete_tx
//   warp 1      : TMEM allocation + MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16,
//                 operands from shared-memory descriptors, accumulator in TMEM; tcgen05.commit frees smem slots / signals
//                 the epilogue
//   warps 2..5  : epilogue: tcgen05.ld 32x32b.x16 (each warp reads the TMEM lane quarter warp_id % 4), + bias, then either
//                 bf16 / fp32 stores or fp32 atomic accumulation (split-K over blockIdx.z for the batch-reduction GEMMs)
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA), SYNCS (mbarrier).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace snb {
namespace tc {

constexpr int BM = 128;       // UMMA M
constexpr int BK = 64;        // bf16 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int STAGES = 4;
constexpr int THREADS = 192;  // 6 warps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of a converged warp; the compiler then knows the region is single-lane and issues UTCHMMA / UTMALDG / UTCBAR without
// its per-lane election loop (which costs ~50 cycles per instruction on the chain's critical path)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major operand tile laid out by TMA with SWIZZLE_128B: rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(const void* smem_tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                   // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                         // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                   // version, bits [46,48)
    d |= (uint64_t)2 << 61;                                   // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// Same for an MN-major operand tile (the M or N index is the contiguous one): TMA boxes of [64 K-rows][64 elements = 128 B],
// SWIZZLE_128B; 8 K-rows are 1024 bytes apart (SBO), consecutive 64-element MN blocks are `mn_block_bytes` apart (LBO).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(const void* smem_tile, uint32_t mn_block_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4);
    d |= (uint64_t)((mn_block_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: D fp32, A/B bf16; bit 15 / 16 = A / B is MN-major; N at [17,23) in units of 8, M at [24,29) in units of 16.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool mn_major = false) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

enum Epilogue { STORE_BF16 = 0, STORE_F32 = 1, ATOMIC_F32 = 2 };

struct GemmArgs {
    int M, N, K;        // K = full reduction length; blockIdx.z owns [z*kslice, min(K, (z+1)*kslice))
    int kslice;         // multiple of BK
    void* D;            // bf16 or fp32, row-major
    long ldd;
    const float* bias;  // per-N, may be null (ignored by ATOMIC_F32)
    float alpha;
};

template <int BN>
constexpr size_t smem_bytes() { return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 /*align*/ + 256 /*barriers*/; }

// MN = false: A[M x K], B[N x K] (K contiguous).  MN = true: A given as At[K x M], B as Bt[K x N] (M / N contiguous), i.e.
// D = At^T Bt -- the batch-reduction GEMMs of a backward pass read activations in their natural [sample][feature] layout.
template <int BN, int EPI, bool MN>
__global__ void __launch_bounds__(THREADS)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs args) {
    static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [32, 256] here");
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B: 1024-byte alignment
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t* empty = full + STAGES;
    uint64_t* acc_ready = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * args.kslice;
    const int kend = min(args.K, kbeg + args.kslice);
    const int nkb = (kend - kbeg + BK - 1) / BK;     // k-blocks of this CTA (TMA zero-fills beyond K)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, round = kb / STAGES;
                if (round > 0) mbar_wait(empty + s, (round - 1) & 1);
                mbar_expect_tx(full + s, A_BYTES + B_BYTES);
                if (MN) {
#pragma unroll
                    for (int h = 0; h < BM / 64; ++h) tma_load_2d(sA + s * A_BYTES + h * (BK * 128), &map_a, m0 + 64 * h, kbeg + kb * BK, full + s);
#pragma unroll
                    for (int h = 0; h < BN / 64; ++h) tma_load_2d(sB + s * B_BYTES + h * (BK * 128), &map_b, n0 + 64 * h, kbeg + kb * BK, full + s);
                } else {
                    tma_load_2d(sA + s * A_BYTES, &map_a, kbeg + kb * BK, m0, full + s);
                    tma_load_2d(sB + s * B_BYTES, &map_b, kbeg + kb * BK, n0, full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, MN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES, round = kb / STAGES;
                mbar_wait(full + s, round & 1);
                tc_fence_after();
                const uint64_t da = MN ? make_mnmajor_sw128_desc(sA + s * A_BYTES, BK * 128) : make_kmajor_sw128_desc(sA + s * A_BYTES);
                const uint64_t db = MN ? make_mnmajor_sw128_desc(sB + s * B_BYTES, BK * 128) : make_kmajor_sw128_desc(sB + s * B_BYTES);
                // one UMMA K step = 16 bf16: K-major -> 32 bytes along the row (+2 in 16-byte units);
                //                            MN-major -> 16 rows of 128 bytes (+128 in 16-byte units)
                constexpr int KSTEP = MN ? 128 : 2;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tmem_acc, da + KSTEP * k, db + KSTEP * k, idesc, (kb | k) ? 1u : 0u);
                umma_commit(empty + s);             // frees the smem slot once these MMAs have read it
            }
            umma_commit(acc_ready);                 // accumulator complete
        }
    } else {
        // ---------------- epilogue: warps 2..5 -> TMEM lane quarter (warp % 4) ----------------
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        if (nkb > 0) {
            mbar_wait(acc_ready, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            if (nkb > 0) {
                tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0u;
            }
            if (row < args.M) {
                const int nbase = n0 + c0;
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * args.alpha;
                if (EPI == ATOMIC_F32) {
                    float* d = reinterpret_cast<float*>(args.D) + (size_t)row * args.ldd + nbase;
                    if (nbase + 16 <= args.N && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {   // four 16-byte reductions instead of 16 scalar atomics
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4 * j), "f"(f[4 * j]), "f"(f[4 * j + 1]), "f"(f[4 * j + 2]),
                                         "f"(f[4 * j + 3])
                                         : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nbase + j < args.N) atomicAdd(d + j, f[j]);
                    }
                } else {
                    if (args.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nbase + j < args.N) f[j] += __ldg(args.bias + nbase + j);
                    }
                    if (EPI == STORE_F32) {
                        float* d = reinterpret_cast<float*>(args.D) + (size_t)row * args.ldd + nbase;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nbase + j < args.N) d[j] = f[j];
                    } else {
                        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + (size_t)row * args.ldd + nbase;
                        if (nbase + 16 <= args.N && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
                            uint32_t pk[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                                pk[j] = *reinterpret_cast<uint32_t*>(&h);
                            }
                            reinterpret_cast<uint4*>(d)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            reinterpret_cast<uint4*>(d)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (nbase + j < args.N) d[j] = __float2bfloat16(f[j]);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 row-major [rows][cols] tensor, box = [box_rows][64 cols], SWIZZLE_128B, out-of-bounds reads return zero
inline int make_map_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    SN_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    SN_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 2) % 16 == 0,
                 "TMA needs a 16-byte aligned base and a row pitch that is a multiple of 16 bytes (ld = %llu bf16)", (unsigned long long)ld_elems);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SN_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

// D (+)= alpha * A[M x K] B[N x K]^T ; A, B bf16 K-major with leading dimensions lda, ldb (elements)          (MN = false)
// D (+)= alpha * At[K x M]^T Bt[K x N] ; At, Bt bf16 row-major with leading dimensions lda, ldb (elements)     (MN = true)
// ksplit > 1 requires EPI == ATOMIC_F32 (each K slice adds its partial sum).
template <int BN, int EPI, bool MN = false>
inline int gemm_bf16_tc(int M, int N, int K, const void* A, long lda, const void* B, long ldb, void* D, long ldd, const float* bias,
                        float alpha, int ksplit, cudaStream_t stream) {
    static_assert(!MN || BN % 64 == 0, "MN-major operands are loaded in 64-element blocks");
    if (M <= 0 || N <= 0) return 0;
    CUtensorMap ma, mb;
    if (MN) {
        if (int rc = make_map_bf16(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK)) return rc;   // [K rows][M cols], box 64 x 64
        if (int rc = make_map_bf16(&mb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK)) return rc;
    } else {
        if (int rc = make_map_bf16(&ma, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM)) return rc;
        if (int rc = make_map_bf16(&mb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN)) return rc;
    }
    GemmArgs args;
    args.M = M; args.N = N; args.K = K;
    if (ksplit < 1) ksplit = 1;
    args.kslice = round_up(ceil_div(K > 0 ? K : 1, ksplit), BK);
    args.D = D; args.ldd = ldd; args.bias = bias; args.alpha = alpha;
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), ceil_div(K > 0 ? K : 1, args.kslice));
    SN_CHECK_ARG(EPI == ATOMIC_F32 || grid.z == 1, "split-K needs the atomic epilogue");
    constexpr size_t smem = smem_bytes<BN>();
    SN_SET_MAX_SMEM((int)smem, gemm_bf16_tc_kernel<BN, EPI, MN>);
    SN_LAUNCH("gemm_bf16_tc_kernel", stream, gemm_bf16_tc_kernel<BN, EPI, MN><<<grid, THREADS, smem, stream>>>(ma, mb, args));
    return 0;
}

}  // namespace tc
}  // namespace snb
