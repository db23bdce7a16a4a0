// One-shot all-reduce of the flat gradient buffer over NVLink peer memory (the path's only exchange step).
//
// The buffer is small (1.7 MB for the 500-stage SSS layer) and the exchange is latency bound: a ring / tree all-reduce pays several
// hops, this kernel pays one.  Every rank owns a region allocated with cudaMalloc and mapped into its peers with CUDA IPC:
//     [ flags: P2P_MAX_BLOCKS x world uint32 | epochs: P2P_MAX_BLOCKS uint32 | parity 0: world slots | parity 1: world slots ]
// Block b of every rank looks after the same slice of the buffer:
//   1. PUSHES its slice of the local gradients into slot [rank] of every rank's region (remote stores do not wait for a round trip;
//      a first version pulled the peers' copies with remote loads and was no faster than NCCL: 24.6 vs 21.1 us on two GPUs),
//   2. tells block b of every peer (a release store of the call's epoch into the peer's flag [b][rank]) and waits until every peer
//      has told it (acquire loads of its own flags) -- a barrier among the blocks b of the ranks, nothing wider,
//   3. sums the slice over the slots of its OWN region in rank order (so every rank computes bit-identical sums), in place.
// Two sets of slots alternate by epoch parity, which makes a second barrier unnecessary: a rank can only overwrite the slots of epoch e
// in epoch e + 2, i.e. after its barrier of epoch e + 1, which a peer enters only after it has finished reading in epoch e.
// The epoch counters live in device memory, so the launch is identical every step and can be captured in a CUDA graph.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int P2P_MAX_BLOCKS = 256;
constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_THREADS = 1024;
constexpr size_t P2P_FLAGS_BYTES = (size_t)P2P_MAX_BLOCKS * P2P_MAX_WORLD * sizeof(uint32_t);
constexpr size_t P2P_EPOCH_OFF = P2P_FLAGS_BYTES;
constexpr size_t P2P_DATA_OFF = P2P_EPOCH_OFF + (size_t)P2P_MAX_BLOCKS * sizeof(uint32_t);     // 17 408 bytes: 16-byte aligned

struct PeerTable { char* base[P2P_MAX_WORLD]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// slots are written by the peers: never from this SM's L1
__device__ __forceinline__ float4 ld_slot4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_slot1(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(P2P_THREADS)
allreduce_oneshot_kernel(float* __restrict__ data, long n, const PeerTable peers, int rank, int world, long cap_floats, float scale) {
    __shared__ uint32_t s_epoch;
    const int blk = blockIdx.x, tid = threadIdx.x;
    char* own = peers.base[rank];
    if (tid == 0) {
        uint32_t* ep = reinterpret_cast<uint32_t*>(own + P2P_EPOCH_OFF) + blk;
        s_epoch = *ep + 1;
        *ep = s_epoch;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const size_t slot_bytes = (size_t)cap_floats * sizeof(float);
    const size_t par_off = P2P_DATA_OFF + (size_t)(epoch & 1u) * (size_t)world * slot_bytes;
    const long n4 = n >> 2;
    const long per = (n4 + gridDim.x - 1) / gridDim.x;
    const long lo = (long)blk * per, hi = lo + per < n4 ? lo + per : n4;
    const bool tail = blk == 0 && (n & 3) != 0 && tid < (n & 3);        // block 0 also looks after the last n % 4 floats
    // 1. local slice -> slot [rank] of every rank
    for (long i = lo + tid; i < hi; i += P2P_THREADS) {
        const float4 v = reinterpret_cast<const float4*>(data)[i];
        for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.base[r] + par_off + (size_t)rank * slot_bytes)[i] = v;
    }
    if (tail) {
        const float v = data[4 * n4 + tid];
        for (int r = 0; r < world; ++r) reinterpret_cast<float*>(peers.base[r] + par_off + (size_t)rank * slot_bytes)[4 * n4 + tid] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. barrier among the blocks `blk` of all ranks
    if (tid < world && tid != rank) {
        st_release_sys(reinterpret_cast<uint32_t*>(peers.base[tid]) + (size_t)blk * P2P_MAX_WORLD + rank, epoch);
        const uint32_t* f = reinterpret_cast<const uint32_t*>(own) + (size_t)blk * P2P_MAX_WORLD + tid;
        // a peer that died must not hang this GPU for good: after 2^24 polls (seconds) the kernel traps and the error surfaces at the
        // next synchronisation
        long polls = 0;
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if (++polls > (1L << 24)) {
                printf("allreduce_oneshot_kernel: rank %d block %d waited for rank %d (epoch %u) too long\n", rank, blk, tid, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
    // 3. sum the slots of the own region in rank order, in place
    for (long i = lo + tid; i < hi; i += P2P_THREADS) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 v = ld_slot4(reinterpret_cast<const float4*>(own + par_off + (size_t)r * slot_bytes) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(data)[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    }
    if (tail) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += ld_slot1(reinterpret_cast<const float*>(own + par_off + (size_t)r * slot_bytes) + 4 * n4 + tid);
        data[4 * n4 + tid] = acc * scale;
    }
}

}  // namespace

extern "C" {

size_t sn_p2p_region_bytes(int64_t capacity_floats, int world) {
    if (capacity_floats <= 0 || world <= 0 || world > P2P_MAX_WORLD) return 0;
    const size_t cap = ((size_t)capacity_floats + 3) & ~(size_t)3;
    return P2P_DATA_OFF + 2 * (size_t)world * cap * sizeof(float);
}

int sn_p2p_alloc(void** region, int64_t capacity_floats, int world) {
    SN_CHECK_ARG(region != nullptr && capacity_floats > 0 && world >= 1 && world <= P2P_MAX_WORLD, "p2p_alloc: bad argument");
    const size_t bytes = sn_p2p_region_bytes(capacity_floats, world);
    SN_CHECK_CUDA(cudaMalloc(region, bytes));
    SN_CHECK_CUDA(cudaMemset(*region, 0, bytes));
    SN_CHECK_CUDA(cudaDeviceSynchronize());
    return 0;
}

int sn_p2p_free(void* region) {
    if (region != nullptr) SN_CHECK_CUDA(cudaFree(region));
    return 0;
}

int sn_p2p_export(const void* region, unsigned char* handle64) {
    SN_CHECK_ARG(region != nullptr && handle64 != nullptr, "p2p_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    SN_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(region)));
    memcpy(handle64, &h, 64);
    return 0;
}

int sn_p2p_import(const unsigned char* handle64, void** region) {
    SN_CHECK_ARG(region != nullptr && handle64 != nullptr, "p2p_import: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SN_CHECK_CUDA(cudaIpcOpenMemHandle(region, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int sn_p2p_close(void* region) {
    if (region != nullptr) SN_CHECK_CUDA(cudaIpcCloseMemHandle(region));
    return 0;
}

// data: this rank's buffer (n floats, 16-byte aligned), summed over the ranks in place and multiplied by `scale`.
// regions[r]: rank r's region as mapped in this process (regions[rank] = the pointer sn_p2p_alloc returned), host array of `world` pointers.
// Every rank must call this with the same n and capacity, the same number of times.
int sn_allreduce_oneshot_f32(float* data, int64_t n, void* const* regions, int rank, int world, int64_t capacity_floats, float scale, sn_stream_t stream) {
    SN_CHECK_ARG(data != nullptr && regions != nullptr, "allreduce_oneshot: NULL argument");
    SN_CHECK_ARG(world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "allreduce_oneshot: bad rank / world size");
    SN_CHECK_ARG(n > 0 && n <= capacity_floats, "allreduce_oneshot: n exceeds the capacity of the regions");
    SN_CHECK_ARG((reinterpret_cast<uintptr_t>(data) & 15) == 0, "allreduce_oneshot: data must be 16-byte aligned");
    PeerTable t;
    for (int r = 0; r < P2P_MAX_WORLD; ++r) t.base[r] = r < world ? static_cast<char*>(regions[r]) : nullptr;
    for (int r = 0; r < world; ++r) SN_CHECK_ARG(t.base[r] != nullptr, "allreduce_oneshot: NULL region");
    const long cap = (long)(((size_t)capacity_floats + 3) & ~(size_t)3);
    const long n4 = n >> 2;
    long blocks = (n4 + P2P_THREADS - 1) / P2P_THREADS;     // one float4 per thread while the SMs last
    if (blocks < 1) blocks = 1;
    {
        int dev = 0, sms = 148;
        SN_CHECK_CUDA(cudaGetDevice(&dev));
        SN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (blocks > sms) blocks = sms;
    }
    if (blocks > P2P_MAX_BLOCKS) blocks = P2P_MAX_BLOCKS;
    cudaStream_t st = snb::as_stream(stream);
    SN_LAUNCH("allreduce_oneshot_kernel", st, allreduce_oneshot_kernel<<<(unsigned)blocks, P2P_THREADS, 0, st>>>(data, (long)n, t, rank, world, cap, scale));
    return 0;
}

}  // extern "C"
