// Call sites of the hot path inside the training loop (reference training_helpers.py:57-73, 139-153):
//   * fused CrossEntropy (mean) + accuracy + gradient of the logits: one pass over the (batch, classes) output of the layer
//     instead of ATen's log_softmax / nll_loss / argmax / eq / sum kernels and their backward;
//   * fused MSE (mean) + accuracy + gradient for one-hot / regression targets;
//   * SGD and Adam over a layer's FLAT parameter / gradient buffers (one launch instead of one per parameter tensor -- the SSS layer
//     has 3 501), with the data-parallel gradient scale (1 / world size after the all-reduce) folded in, and a device-side step counter
//     so that the whole training step stays capturable in a CUDA graph.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int CE_WARPS = 8;

// warp per sample row; loss_sum += -log_softmax(x)[target] / B ; correct += (argmax == target) ; grad = (softmax - onehot) * gscale
__global__ void __launch_bounds__(CE_WARPS * 32)
ce_loss_kernel(const float* __restrict__ logits, long ld, const long long* __restrict__ targets, long B, int C, float inv_b, float* __restrict__ loss_sum,
               int* __restrict__ correct, float* __restrict__ grad, long ldg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float s_loss[CE_WARPS];
    __shared__ int s_corr[CE_WARPS];
    float my_loss = 0.f;
    int my_corr = 0;
    for (long row = (long)blockIdx.x * CE_WARPS + warp; row < B; row += (long)gridDim.x * CE_WARPS) {
        const float* x = logits + row * ld;
        const int tgt = (int)targets[row];
        float mx = -INFINITY;
        int arg = 0;
        for (int c = lane; c < C; c += 32) {
            const float v = x[c];
            if (v > mx) { mx = v; arg = c; }      // first maximum of the lane's (increasing) indices
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }   // torch.argmax returns the first maximal index
        }
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float lse = mx + logf(se);
        if (lane == 0) {
            my_loss += lse - x[tgt];
            my_corr += (arg == tgt) ? 1 : 0;
        }
        if (grad != nullptr) {
            float* g = grad + row * ldg;
            for (int c = lane; c < C; c += 32) {
                const float p = expf(x[c] - lse);
                g[c] = (p - (c == tgt ? 1.f : 0.f)) * inv_b;
            }
        }
    }
    if (lane == 0) { s_loss[warp] = my_loss; s_corr[warp] = my_corr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f;
        int k = 0;
        for (int w = 0; w < CE_WARPS; ++w) { l += s_loss[w]; k += s_corr[w]; }
        atomicAdd(loss_sum, l * inv_b);
        if (correct != nullptr) atomicAdd(correct, k);
    }
}

// MSE (mean over all B*C entries) + gradient 2 (x - t) / (B C); accuracy by argmax(x) == argmax(t) when C > 1 (reference :66-70)
__global__ void __launch_bounds__(CE_WARPS * 32)
mse_loss_kernel(const float* __restrict__ x, long ld, const float* __restrict__ t, long ldt, long B, int C, float inv_n, float* __restrict__ loss_sum,
                int* __restrict__ correct, float* __restrict__ grad, long ldg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float s_loss[CE_WARPS];
    __shared__ int s_corr[CE_WARPS];
    float my_loss = 0.f;
    int my_corr = 0;
    for (long row = (long)blockIdx.x * CE_WARPS + warp; row < B; row += (long)gridDim.x * CE_WARPS) {
        const float* xr = x + row * ld;
        const float* tr = t + row * ldt;
        float mx = -INFINITY, mt = -INFINITY, se = 0.f;
        int ax = 0, at = 0;
        for (int c = lane; c < C; c += 32) {
            const float v = xr[c], w = tr[c], d = v - w;
            se = fmaf(d, d, se);
            if (v > mx) { mx = v; ax = c; }
            if (w > mt) { mt = w; at = c; }
            if (grad != nullptr) grad[row * ldg + c] = 2.f * d * inv_n;
        }
        for (int o = 16; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            const float om = __shfl_xor_sync(0xffffffffu, mx, o), ot = __shfl_xor_sync(0xffffffffu, mt, o);
            const int oa = __shfl_xor_sync(0xffffffffu, ax, o), ob = __shfl_xor_sync(0xffffffffu, at, o);
            if (om > mx || (om == mx && oa < ax)) { mx = om; ax = oa; }
            if (ot > mt || (ot == mt && ob < at)) { mt = ot; at = ob; }
        }
        if (lane == 0) { my_loss += se; my_corr += (ax == at) ? 1 : 0; }
    }
    if (lane == 0) { s_loss[warp] = my_loss; s_corr[warp] = my_corr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f;
        int k = 0;
        for (int w = 0; w < CE_WARPS; ++w) { l += s_loss[w]; k += s_corr[w]; }
        atomicAdd(loss_sum, l * inv_n);
        if (correct != nullptr) atomicAdd(correct, k);
    }
}

__global__ void flat_sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long n, float lr, float gscale) {
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15) == 0) {
        float4 pv = *reinterpret_cast<float4*>(p + i);
        const float4 gv = *reinterpret_cast<const float4*>(g + i);
        const float a = -lr * gscale;
        pv.x = fmaf(a, gv.x, pv.x); pv.y = fmaf(a, gv.y, pv.y); pv.z = fmaf(a, gv.z, pv.z); pv.w = fmaf(a, gv.w, pv.w);
        *reinterpret_cast<float4*>(p + i) = pv;
    } else {
        for (long k = i; k < n && k < i + 4; ++k) p[k] = fmaf(-lr * gscale, g[k], p[k]);
    }
}

// torch.optim.Adam semantics (no amsgrad, no weight decay): m, v exponential averages, bias-corrected step
__global__ void flat_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n, float lr, float beta1,
                                 float beta2, float eps, float gscale, const int* __restrict__ step_dev, int step_host) {
    const int step = step_dev != nullptr ? *step_dev + 1 : step_host;
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const long i0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    for (long k = i0; k < n && k < i0 + 4; ++k) {
        const float gk = g[k] * gscale;
        const float mk = beta1 * m[k] + (1.f - beta1) * gk;
        const float vk = beta2 * v[k] + (1.f - beta2) * gk * gk;
        m[k] = mk;
        v[k] = vk;
        p[k] -= step_size * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
    }
}
__global__ void increment_kernel(int* c) { *c += 1; }

}  // namespace

extern "C" {

// loss_sum (device float) and correct (device int, nullable) are ACCUMULATED into: zero them before the first batch of a pass.
// grad_logits (nullable): d(mean loss)/d(logits), written (not accumulated).
int sn_ce_loss(const float* logits, int64_t ld, const int64_t* targets, int64_t B, int C, float* loss_sum, int* correct, float* grad_logits,
               int64_t ldg, sn_stream_t stream) {
    SN_CHECK_ARG(logits && targets && loss_sum && B >= 0 && C > 0, "ce_loss: bad arguments");
    if (B == 0) return 0;
    const unsigned grid = (unsigned)((B + CE_WARPS - 1) / CE_WARPS < 148 * 8 ? (B + CE_WARPS - 1) / CE_WARPS : 148 * 8);
    SN_LAUNCH("ce_loss_kernel", snb::as_stream(stream), ce_loss_kernel<<<grid, CE_WARPS * 32, 0, snb::as_stream(stream)>>>(
        logits, ld, reinterpret_cast<const long long*>(targets), B, C, 1.f / (float)B, loss_sum, correct, grad_logits, ldg));
    return 0;
}
int sn_mse_loss(const float* x, int64_t ld, const float* target, int64_t ldt, int64_t B, int C, float* loss_sum, int* correct, float* grad_x, int64_t ldg,
                sn_stream_t stream) {
    SN_CHECK_ARG(x && target && loss_sum && B >= 0 && C > 0, "mse_loss: bad arguments");
    if (B == 0) return 0;
    const unsigned grid = (unsigned)((B + CE_WARPS - 1) / CE_WARPS < 148 * 8 ? (B + CE_WARPS - 1) / CE_WARPS : 148 * 8);
    SN_LAUNCH("mse_loss_kernel", snb::as_stream(stream), mse_loss_kernel<<<grid, CE_WARPS * 32, 0, snb::as_stream(stream)>>>(
        x, ld, target, ldt, B, C, 1.f / ((float)B * (float)C), loss_sum, correct, grad_x, ldg));
    return 0;
}
// p -= lr * grad_scale * g over n floats
int sn_flat_sgd(float* params, const float* grads, int64_t n, float lr, float grad_scale, sn_stream_t stream) {
    SN_CHECK_ARG(params && grads && n >= 0, "flat_sgd: bad arguments");
    if (n == 0) return 0;
    SN_LAUNCH("flat_sgd_kernel", snb::as_stream(stream), flat_sgd_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, snb::as_stream(stream)>>>(params, grads, n, lr, grad_scale));
    return 0;
}
// Adam step over n floats.  step_dev (nullable): device int holding the number of steps taken so far -- read by the kernel and
// incremented afterwards (graph-capturable); when NULL, step_host (1-based) is used.
int sn_flat_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                 float grad_scale, int* step_dev, int step_host, sn_stream_t stream) {
    SN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n >= 0, "flat_adam: bad arguments");
    if (n == 0) return 0;
    cudaStream_t st = snb::as_stream(stream);
    SN_LAUNCH("flat_adam_kernel", st, flat_adam_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                                      grad_scale, step_dev, step_host));
    if (step_dev != nullptr) { increment_kernel<<<1, 1, 0, st>>>(step_dev); SN_CHECK_LAUNCH("increment_kernel"); }
    return 0;
}

}  // extern "C"
