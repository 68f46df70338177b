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

#define LRCE_ABI_VERSION 2

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
 * nn.TransformerDecoderLayer built at fusionv3.py:8-17.
 *
 * LayerNorm folding (norm1 -> qkv, norm2 -> fc1; video_swin_ori.py:252,:285): instead of a separate LayerNorm pass,
 *  - a producing GEMM called with out_stats != NULL also writes float2 out_stats[N/cw][M] = (mean, M2) of every
 *    cw-column chunk of the rows it stores (taken just before the bf16 rounding); cw = 64 when N % 256 == 0, else 32;
 *  - a consuming GEMM called with in_stats != NULL (float2 [K/in_chunk][M], in_chunk = the producer's cw, at most 16
 *    chunks) takes A = the RAW rows, W = W * diag(gamma),
 *    in_colsum[n] = sum_k W'[n,k], bias = bias + W beta, and computes
 *    out[m,n] = epilogue(rstd[m] * (acc[m,n] - mean[m] * in_colsum[n]) + bias[n]) with (mean, rstd) of row m rebuilt
 *    from the partials (Chan's combination, eps = in_eps). Exactly LayerNorm(x) W^T + b in exact arithmetic.
 * in_stats requires the BIAS or BIAS_GELU epilogue; out_stats requires bf16 output. */
int lrce_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                   const void* residual, int ldr, void* out, int ldo, int epilogue, int out_fp32,
                   const float* ln_gamma, const float* ln_beta, float ln_eps, const float* in_stats, int in_chunk,
                   const float* in_colsum, float in_eps, float* out_stats, void* stream);

/* ---- HBM-bound row kernels (Video Swin) --------------------------------------------------------------------- */

/* y[rows, C] = LayerNorm(x[rows, C]) * gamma + beta, x bf16, y bf16 (or fp32 when out_fp32 != 0); C in
 * {128, 256, 512, 768, 1024, 2048}. Replaces nn.LayerNorm at video_swin_ori.py:252 (norm1), :285 (norm2), :684 (final
 * norm). */
int lrce_layernorm_bf16(const void* x, void* y, const float* gamma, const float* beta, float eps, long long rows, int C,
                        int out_fp32, void* stream);

/* PatchMerging front half: y[n_seg*D*(H/2)*(W/2), 4C] = LayerNorm(cat[x(2i,2j), x(2i+1,2j), x(2i,2j+1), x(2i+1,2j+1)])
 * on x bf16 [n_seg, D, H, W, C]; the bias-free reduction Linear is a following lrce_gemm_bf16. gamma/beta have 4C
 * entries. Replaces video_swin_ori.py:333-339. H and W must be even (always true on this path). */
int lrce_patch_merge_ln_bf16(const void* x, void* y, const float* gamma, const float* beta, float eps, int n_seg, int D,
                             int H, int W, int C, void* stream);

/* PatchEmbed3D front half: clips fp32 [n_seg, T, 3, Hin, Win] in [0,1] -> A bf16 [n_seg*ceil(T/2)*(Hin/4)*(Win/4), 96],
 * ImageNet-normalised, K ordered (c, kd, kh, kw) like the Conv3d weight, frames >= T zero (padding happens after
 * normalisation). Followed by lrce_gemm_bf16(..., LRCE_EPI_BIAS_LN). Replaces video.py:35-37 + video_swin_ori.py:472-475. */
int lrce_patch_gather_f32(const float* clips, void* A, int n_seg, int T, int Hin, int Win, void* stream);

/* Standalone cyclic shift + window partition on bf16 [n_seg, D*H*W, C] -> [n_seg*nWin*N, C] (inverse != 0: the exact
 * inverse, window_reverse + roll back). The production path fuses this map into lrce_window_attention_bf16; this entry
 * exists for the bit-exact remap test and the HBM roofline measurement. Replaces torch.roll + window_partition /
 * window_reverse + torch.roll, video_swin_ori.py:262-276, :60-88. */
int lrce_window_remap_bf16(const void* in, void* out, int n_seg, int D, int H, int W, int C, int wd, int wh, int ww, int sd,
                           int sh, int sw, int inverse, void* stream);

/* Integer tables for one segment (each may be NULL): gather[nWin*N] = source token of (window, token);
 * region[nWin*N] = shift-mask region id (tokens attend iff ids are equal); relpos[N] = f(t) with
 * relative_position_index[i][j] = f(i) - f(j) + 1267. Same device functions the attention kernel uses.
 * Replaces window_partition/roll index math, compute_mask (video_swin_ori.py:346-359) and :134-147. */
int lrce_remap_index(int* gather, int* region, int* relpos, int D, int H, int W, int wd, int wh, int ww, int sd, int sh,
                     int sw, void* stream);

/* ---- window attention ------------------------------------------------------------------------------------------ */

/* bias_dense bf16 [n_heads, 147, 152] = relative_position_bias_table[index(i,j)][h] * log2(e) (columns >= 147 zero) from
 * table fp32 [2535, n_heads]; done once per weight load. Replaces the gather at video_swin_ori.py:171-173. */
int lrce_window_bias_pack(const float* table, void* bias_dense, int n_heads, void* stream);

/* out[n_seg*D*H*W, C] = merge_heads(softmax(q k^T / sqrt(32) + bias + shift_mask) v) per (3,7,7) window, with the cyclic
 * shift (0, shift_h, shift_w), window partition and their inverses fused into the loads/stores: qkv and out are both in
 * natural token order. qkv bf16 [n_seg*D*H*W, 3C] laid out [q|k|v][head][32]; C = 32 * n_heads; D == 3, H % 7 == W % 7
 * == 0. Replaces video_swin_ori.py:262-276 around WindowAttention3D.forward :166-186. */
int lrce_window_attention_bf16(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H, int W, int C,
                               int n_heads, int shift_h, int shift_w, void* stream);

/* ---- recurrent cross-modal encoder ------------------------------------------------------------------------------ */

/* VideoPosEmbed: out bf16 [B, S, T*(P+1), 768] from proj bf16 [B*S*T*P, 768] (projection_layer output), CLS row per
 * frame, + emb_pos[P+1] + emb_len[T] + emb_clip[S], LayerNorm eps. Replaces embedding.py:47-63. */
int lrce_video_posembed_ln(const void* proj, const float* emb_cls, const float* emb_pos, const float* emb_len,
                           const float* emb_clip, const float* gamma, const float* beta, float eps, void* out, int B, int S,
                           int T, int P, void* stream);

/* TextPosEmbed: out bf16 [Bt, L+1, 768] from text [Bt, L, 768] (bf16, or fp32 when text_fp32 != 0).
 * Replaces embedding.py:17-23. */
int lrce_text_posembed_ln(const void* text, int text_fp32, const float* emb_cls, const float* emb_pos, const float* gamma,
                          const float* beta, float eps, void* out, int Bt, int L, void* stream);

#define LRCE_ACT_NONE 0
#define LRCE_ACT_GELU 1
#define LRCE_ACT_RELU 2

/* Y[rows, N] (pitch ldy; fp32, or bf16 when y_bf16 != 0) = act(X W^T + bias) with X = Xa (+ Xb), optionally
 * LayerNorm'ed (ln_gamma != NULL, K == 768; the normalised X is also written to Xout fp32 when non-NULL). Xa is fp32
 * [rows, K], or bf16 when xa_bf16 != 0 (then Xb and ln_gamma must be NULL); K in {768, 3072}; W bf16 with
 * ceil(N/8)*8 rows of K. The summarisation-token path of nn.TransformerDecoderLayer (post-norm) and final_fc:
 * replaces fusionv3.py:46 (per-layer linears + norms on the 1-token target) and :195. */
int lrce_skinny_linear(const void* Xa, int xa_bf16, const float* Xb, const float* ln_gamma, const float* ln_beta, float eps,
                       float* Xout, const void* W, const float* bias, void* Y, int y_bf16, int rows, int K, int N, int ldy,
                       int act, void* stream);

/* ctx[rows, 768] (bf16) = softmax(q K^T) V for one query token per row over the memory [video segment `seg` (Tv tokens) ; text
 * (Lt tokens)], 12 heads x 64; q fp32 already scaled by 1/8. kv_video bf16 [(rows/n_cand)*S*Tv, ld_kv], kv_text bf16
 * [rows*Lt, ld_kv], K at column layer*1536 + head*64, V at +768. Replaces the multihead_attn call inside
 * nn.TransformerDecoderLayer (fusionv3.py:45-46; candidate expansion fusionv3.py:259). */
int lrce_cross_attention(const float* q, const void* kv_video, const void* kv_text, void* ctx, int rows, int seg, int S,
                         int Tv, int Lt, int n_cand, int layer, int ld_kv, void* stream);

/* tok_out = LN_f(tok + LN_3(h + y)): closes the 12th decoder layer and the recurrent step (fusionv3.py:47-48). */
int lrce_recurrent_update(const float* tok, const float* h, const float* y, const float* g3, const float* b3,
                          const float* gf, const float* bf, float eps, float* tok_out, int rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LRCE_B200_H_ */
