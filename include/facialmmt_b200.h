/* facialmmt_b200 -- C ABI of the B200-native FacialMMT inference forward path.
 *
 * The reference (NUSTM/FacialMMT) has no FFI: its boundary for this path is the Python nn.Module API of
 * src/models.py (SwinForAffwildClassification :14-37, MultiModalTransformerForClassification :41-188,
 * meld_utt_transformer :192-223) plus the eval glue of train.py:169-234. Each entry point below names the
 * reference call it replaces. All pointers are plain device (or, where stated, host) pointers; no torch types.
 * Every function returns 0 on success and a negative code on failure; fmmt_last_error() gives the message.
 * Calls are asynchronous on the given cudaStream_t (passed as void*), and never synchronise the device unless
 * stated. A handle belongs to one device and is not re-entrant.
 */
#ifndef FACIALMMT_B200_H
#define FACIALMMT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(FMMT_BUILD)
#define FMMT_API __attribute__((visibility("default")))
#else
#define FMMT_API
#endif

#define FMMT_OK 0
#define FMMT_ERR_INVALID (-1)   /* bad shape / pointer / alignment (the reference would assert) */
#define FMMT_ERR_CUDA (-2)      /* CUDA runtime / driver error */
#define FMMT_ERR_STATE (-3)     /* handle not finalized, weight missing, ... */
#define FMMT_ERR_NOGPU (-4)     /* no sm_100 device: there is NO CPU fallback */

/* Activation codes for fmmt_op_gemm. */
#define FMMT_ACT_NONE 0
#define FMMT_ACT_GELU 1 /* exact erf GELU (nn.GELU(), F.gelu, modules/Transformer.py:119-124) */
#define FMMT_ACT_RELU 2
#define FMMT_ACT_TANH 3

FMMT_API const char* fmmt_last_error(void);
FMMT_API const char* fmmt_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
FMMT_API int64_t fmmt_launch_count(void);

/* ---- operator-level entry points (one CUDA kernel each); used by the per-kernel parity tests ---- */

/* nn.Linear: out[dest(r),:] = act(A[r,:] @ W^T + bias) + residual[dest(r),:].  A [M,lda] bf16, W [N,ldw] bf16 (the
 * nn.Linear.weight layout), fp32 accumulate on tcgen05 tensor cores. row_map (device int32, length map_period) is
 * optional: dest(r) = (r / map_period) * map_period + row_map[r % map_period]. Either output may be NULL. */
FMMT_API int fmmt_op_gemm(const void* A_bf16, int lda, const void* W_bf16, int ldw, int M, int N, int K,
                          const float* bias, int act, const float* residual, int ldr, float* out_f32, int ldo32,
                          void* out_bf16, int ldo16, const int* row_map, int map_period, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FACIALMMT_B200_H */
