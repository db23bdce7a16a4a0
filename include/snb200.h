/*
 * snb200.h -- C ABI of libsnb200.so, the B200 (sm_100a) implementation of the structured-layer
 * hot path of MatthiasKi/structurednets.
 *
 * The reference has no FFI: its "operator interface" for this path is the forward()/autograd
 * backward of the nn.Modules in structurednets/layers/ (one .py per layer).  Each pair of entry points below
 * replaces one of those (file:line cited per function).  A maintainer of the reference binds them
 * with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     _host; the library owns no memory, allocates nothing and never synchronises the device;
 *   - every kernel is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = ok, non-zero = error, message in sn_last_error_string() (thread local);
 *   - matrices are row-major; "ld" arguments are row strides in elements;
 *   - gradients are ACCUMULATED (+=) into grad buffers, like torch autograd does.
 */
#ifndef SNB200_H
#define SNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sn_stream_t; /* cudaStream_t */

#define SN_VERSION 100

int sn_version(void);
const char* sn_last_error_string(void);
/* number of kernels this library has launched since load / last reset (bench.py's gpu_launches) */
long long sn_launch_count(void);
void sn_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * SSS layer -- replaces SSSLayer.forward (layers/sss_layer.py:99-131) and its autograd backward.
 *
 * One stage descriptor per (direction, processing step).  Direction 0 = causal sweep (stages
 * 0..n-1), direction 1 = anticausal sweep (stages n-1..0); entry [dir*n + kk] describes the
 * kk-th stage that direction visits.  "state in" / "state out" are the state vectors entering /
 * leaving the stage in processing order.  Offsets off_* index the flat parameter buffer (floats):
 *   off_ys: (out_dim x d_in)   C_k  (causal)  or G_k (anticausal)
 *   off_yu: (out_dim x in_dim) D_k  (causal)  or -1  (anticausal: y_k does not see u_k)
 *   off_ss: (d_out x d_in)     A_k  (causal)  or E_k (anticausal)
 *   off_su: (d_out x in_dim)   B_k  (causal)  or F_k (anticausal)
 * ------------------------------------------------------------------------------------------ */
typedef struct sn_sss_stage {
    int32_t in_off, in_dim, out_off, out_dim;
    int32_t d_in, d_out;
    int32_t off_ys, off_yu, off_ss, off_su;
    int32_t pack_off; /* offset (floats) of this stage's block in the packed buffer */
    int32_t k;        /* natural stage index */
    int32_t reserved[4];
} sn_sss_stage;

/* A contiguous run of processing steps [kk_begin, kk_end) handled as one unit (state checkpoint
 * at kk_begin).  Chunks never straddle first_visit -> second_visit. */
typedef struct sn_sss_chunk {
    int32_t kk_begin, kk_end;
    int32_t col0, ncols;   /* input columns covered by the chunk  */
    int32_t row0, nrows;   /* output columns covered by the chunk */
    int32_t second_visit;  /* 1: the other direction already wrote y for these stages */
    int32_t kk_mid;        /* the forward stages its input columns in two halves: [kk_begin,kk_mid) and [kk_mid,kk_end) */
    int32_t col0_a, ncols_a, col0_b, ncols_b; /* input columns of the two halves */
    int32_t reserved[4];
} sn_sss_chunk;

typedef struct sn_sss_plan {
    int32_t nb_states, input_dim, output_dim;
    int32_t rows_pad, k_pad, d_pad; /* multiples of 4: max rows (out_dim+d_out), max K (d_in+in_dim), max state dim */
    int32_t nchunks;                /* per direction */
    int32_t chunk_in_max, chunk_out_max, chunk_len_max;
    int32_t nparams;                /* floats in the flat parameter buffer */
    int32_t half_in_max;            /* max input columns of a half chunk */
    const sn_sss_stage* stages; /* device, [2][nb_states] */
    const sn_sss_chunk* chunks; /* device, [2][nchunks]   */
} sn_sss_plan;

/* floats needed for the packed parameter copy / for the state checkpoints of a batch of B samples */
size_t sn_sss_packed_floats(const sn_sss_plan* plan_host);
size_t sn_sss_ckpt_floats(const sn_sss_plan* plan_host, int64_t B);
/* re-lays the flat parameters out per stage (transposed + row-major, zero padded); call once per
 * parameter update, before forward/backward */
int sn_sss_pack(const sn_sss_plan* plan_host, const float* params, float* packed, sn_stream_t stream);
/* y[B x output_dim] = SSS(x[B x input_dim]) + bias.  bias may be NULL.  ckpt may be NULL (inference);
 * otherwise it receives the chunk-entry states backward() needs. */
int sn_sss_forward(const sn_sss_plan* plan_host, const float* packed, const float* x, int64_t ldx,
                   float* y, int64_t ldy, const float* bias, float* ckpt, int64_t B, sn_stream_t stream);
/* grad_params (flat, same layout as params) += d loss / d params ; grad_bias (may be NULL) += column
 * sums of grad_y.  grad_x (may be NULL) receives d loss / d x (overwritten).  workspace: scratch of
 * sn_sss_backward_workspace_floats() floats (packed per-stage gradient accumulators). */
size_t sn_sss_backward_workspace_floats(const sn_sss_plan* plan_host);
int sn_sss_backward(const sn_sss_plan* plan_host, const float* packed, const float* x, int64_t ldx,
                    const float* grad_y, int64_t ldgy, const float* ckpt, float* workspace, float* grad_params,
                    float* grad_bias, float* grad_x, int64_t ldgx, int64_t B, sn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB200_H */
