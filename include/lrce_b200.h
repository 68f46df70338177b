/* lrce_b200.h — C ABI of liblrce_b200.so: the B200-native (sm_100a) kernels of LRCE's forward hot path.
 *
 * The reference (Sejong-VLI/VQA-LRCE-KBS-2023) is pure PyTorch, so there is no existing FFI to bind against; each entry
 * point below replaces the sequence of PyTorch library ops at the cited reference location (paths relative to the
 * reference root). The Python host module (`vqa-lrce-kbs-2023_b200/`) binds these with ctypes; INTEGRATION.md shows
 * the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *  - activations / weights are bf16 (row-major), biases / LayerNorm parameters / logits are fp32;
 *  - stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronises, allocates or frees;
 *  - returns LRCE_OK (0) or a negative LRCE_E* code; the message is available from lrce_last_error() (thread local);
 *  - hard-fails with LRCE_EARCH on anything that is not compute capability 10.x: there is no fallback path.
 */
#ifndef LRCE_B200_H_
#define LRCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRCE_ABI_VERSION 1

#define LRCE_OK 0
#define LRCE_EINVAL (-1)  /* bad shape / alignment / null pointer (checked on the host before any launch) */
#define LRCE_EARCH (-2)   /* device is not sm_100 */
#define LRCE_ECUDA (-3)   /* CUDA runtime / launch error */
#define LRCE_EDRIVER (-4) /* driver entry point (TMA descriptor encode) unavailable or failed */

int lrce_abi_version(void);
const char* lrce_last_error(void);

/* GEMM epilogues */
#define LRCE_EPI_BIAS 0          /* out = A W^T (+ bias)                         nn.Linear                            */
#define LRCE_EPI_BIAS_GELU 1     /* out = gelu_erf(A W^T + bias)                 Mlp.fc1 + act, video_swin_ori.py:52-53 */
#define LRCE_EPI_BIAS_RESIDUAL 2 /* out = residual + A W^T + bias                x + proj(..) / x + fc2(..), :299,:304  */
#define LRCE_EPI_BIAS_LN 3       /* out = LayerNorm(A W^T + bias), N == 128      PatchEmbed3D proj + norm, :475-479    */

/* out[M,N] = epilogue(A[M,K] * W[N,K]^T): tcgen05 tensor-core GEMM, A/W bf16 row-major with row pitches lda/ldw
 * (elements). N % 128 == 0, K % 8 == 0. `bias` may be NULL. `residual` ([M, ldr] bf16) may alias `out`.
 * out is bf16 [M, ldo] unless out_fp32 != 0 (fp32, bias epilogue only).
 * Replaces nn.Linear at video_swin_ori.py:46-48,150-152,318 / fusionv3.py:154,160 and the K/V in-projections of the
 * nn.TransformerDecoderLayer built at fusionv3.py:8-17. */
int lrce_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                   const void* residual, int ldr, void* out, int ldo, int epilogue, int out_fp32,
                   const float* ln_gamma, const float* ln_beta, float ln_eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LRCE_B200_H_ */
