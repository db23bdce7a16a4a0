// Shared helpers for libsnb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "snb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsnb200 is written for sm_100a (B200) only"
#endif

namespace snb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// optional per-kernel CUDA-event timing (sn_timing_enable / sn_timing_report; bench.py's roofline uses it)
void timing_begin(const char* name, cudaStream_t stream);
void timing_end(cudaStream_t stream);

inline cudaStream_t as_stream(sn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define SN_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            snb::set_error(__VA_ARGS__);        \
            return 1;                           \
        }                                       \
    } while (0)

#define SN_CHECK_CUDA(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            snb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

// cudaFuncSetAttribute(kernel, MaxDynamicSharedMemorySize, bytes) only when `bytes` exceeds what this call site has already set on the
// current device: the attribute is sticky, and the call costs about a microsecond of host time per launch otherwise (an eager SSS step
// is host bound at small batches).  The kernel comes last because template argument lists contain commas.
#define SN_MAX_DEVICES 64
#define SN_SET_MAX_SMEM(bytes, ...)                                                                              \
    do {                                                                                                         \
        static int sn_smem_set_[SN_MAX_DEVICES];                                                                 \
        int sn_dev_ = -1;                                                                                        \
        SN_CHECK_CUDA(cudaGetDevice(&sn_dev_));                                                                  \
        const int sn_b_ = (int)(bytes);                                                                          \
        if (sn_dev_ < 0 || sn_dev_ >= SN_MAX_DEVICES || sn_smem_set_[sn_dev_] < sn_b_) {                         \
            SN_CHECK_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, sn_b_)); \
            if (sn_dev_ >= 0 && sn_dev_ < SN_MAX_DEVICES) sn_smem_set_[sn_dev_] = sn_b_;                         \
        }                                                                                                        \
    } while (0)

#define SN_CHECK_LAUNCH(name)                                                                  \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            snb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));           \
            return 3;                                                                          \
        }                                                                                      \
        snb::count_launch();                                                                   \
    } while (0)

// launch + event pair around it when timing is on + launch-error check
#define SN_LAUNCH(name, stream, ...)        \
    do {                                    \
        snb::timing_begin(name, stream);    \
        __VA_ARGS__;                        \
        snb::timing_end(stream);            \
        SN_CHECK_LAUNCH(name);              \
    } while (0)

// A second stream for independent pieces of one call (short, partly filled grids overlap almost completely).  Fork / join with events
// keeps the caller's stream semantics and is capturable in a CUDA graph.  One per host thread and device.
struct SideStream {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
inline SideStream* side_stream() {
    static thread_local SideStream sd;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (sd.device != dev) {
        if (cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        sd.device = dev;
    }
    return &sd;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace snb
