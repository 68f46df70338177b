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
 *  - a producing GEMM called with out_stats != NULL also writes float2 out_stats[M][N/cw] = (mean, M2) of every
 *    cw-column chunk of the rows it stores (taken just before the bf16 rounding); cw = 64 when N % 256 == 0, else 32;
 *  - a consuming GEMM called with in_stats != NULL (float2 [M][K/in_chunk] — row-major, so a row range of the matrix owns a contiguous range of statistics —, in_chunk = the producer's cw, 4, 8 or 16
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
/* the same for uint8 frames [n_seg, T, 3, Hin, Win] (0..255): x / 255 (torchvision ToTensor, e2e_dataset.py) is applied in
 * the kernel, so a clip crosses PCIe as 1 byte per pixel instead of 4. */
int lrce_patch_gather_u8(const unsigned char* clips, void* A, int n_seg, int T, int Hin, int Win, void* stream);

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

/* bias_dense bf16 [n_heads, 160, 160] = relative_position_bias_table[index(i,j)][h] * log2(e), rows and columns in the
 * attention kernel's slot order (tokens grouped by shift-mask class, 13 pad slots; pad columns = -inf; csrc/remap.cuh
 * key_slot_377); rows 155-156 of every head hold float[160] row maxima. From table fp32 [2535, n_heads]; done once per
 * weight load. Replaces the gather at video_swin_ori.py:171-173. */
int lrce_window_bias_pack(const float* table, void* bias_dense, int n_heads, void* stream);

/* out[n_seg*D*H*W, C] = merge_heads(softmax(q k^T / sqrt(32) + bias + shift_mask) v) per (3,7,7) window, with the cyclic
 * shift (0, shift_h, shift_w), window partition and their inverses fused into the loads/stores: qkv and out are both in
 * natural token order. qkv bf16 [n_seg*D*H*W, 3C] laid out [q|k|v][head][32]; C = 32 * n_heads; D == 3, H % 7 == W % 7
 * == 0. Replaces video_swin_ori.py:262-276 around WindowAttention3D.forward :166-186. */
int lrce_window_attention_bf16(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H, int W, int C,
                               int n_heads, int shift_h, int shift_w, void* stream);

/* Instrumented instantiation of the same kernel for tools/ (no library-side state: the buffer is a per-call argument):
 * per-warp mbarrier-wait cycle counters of CTA 0 are accumulated into prof (device, 224 int64, zeroed by the caller; layout
 * in csrc/window_attn.cu); a wait longer than ~50 ms is reported there instead of hanging (watchdog). */
int lrce_window_attention_profile(const void* qkv, void* out, const void* bias_dense, int n_seg, int D, int H, int W, int C,
                                  int n_heads, int shift_h, int shift_w, void* stream, long long* prof);

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

/* The Swin MLP block of the stage-1 blocks (C = 128) as ONE kernel (csrc/mlp_fused.cu):
 *   out = x + fc2( gelu( fc1( LayerNorm(x) ) ) )        video_swin_ori.py:40-57 (Mlp), :284-285 and :304 (norm2, residual)
 * The hidden activations [M, 512] never leave the SM. x bf16 [M, ldx] raw rows; the LayerNorm is folded exactly as in
 * lrce_gemm_bf16: w1 bf16 [512, 128] = W1 diag(gamma), b1 = b + W1 beta, colsum1[n] = sum_k w1[n, k], in_stats = float2
 * [M][4] (mean, M2) partials of the four 32-column chunks of every row (the out_stats of the GEMM that produced x).
 * w2 bf16 [128, 512], b2 fp32 [128]. out bf16 [M, ldo] may alias x. out_stats NULL or float2 [M][4] partials of the rows
 * written (for the next block's folded norm1). Replaces two lrce_gemm_bf16 calls (EPI_BIAS_GELU + EPI_BIAS_RESIDUAL). */
int lrce_mlp_fused_bf16(const void* x, int ldx, const void* w1, const float* b1, const float* colsum1, const float* in_stats,
                        float in_eps, const void* w2, const float* b2, void* out, int ldo, float* out_stats, int M, int C,
                        void* stream);

/* The Swin MLP block of the stage-2 / stage-3 blocks (C = 256 / 512; C = 128 is accepted with in_chunk = 32, but
 * lrce_mlp_fused_bf16 is the faster kernel there) as ONE kernel with the hidden activations kept in L2
 * (csrc/gemm_tc.cu, mlp_l2_kernel): same contract as lrce_mlp_fused_bf16 (video_swin_ori.py:40-57, :284-285, :304), statistics
 * in chunks of `in_chunk` = 64 columns (32 for C = 128). Every CTA pair walks its 256-row tiles through fc1 (folded LayerNorm + bias + GELU) into
 * its private 256 x 4C bf16 slice of `scratch`, then through fc2 (K = 4C, bias + residual x + out_stats) reading that slice
 * back; the scratch (lrce_mlp_l2_scratch_bytes(C) bytes, 16-byte aligned, contents undefined afterwards) is overwritten by
 * every row tile and therefore stays in L2. Replaces two lrce_gemm_bf16 calls (EPI_BIAS_GELU + EPI_BIAS_RESIDUAL) and the
 * DRAM round trip of the hidden rows between them. */
size_t lrce_mlp_l2_scratch_bytes(int C);
int lrce_mlp_l2_bf16(const void* x, int ldx, const void* w1, const float* b1, const float* colsum1, const float* in_stats,
                     int in_chunk, float in_eps, const void* w2, const float* b2, void* out, int ldo, float* out_stats,
                     void* scratch, size_t scratch_bytes, int M, int C, void* stream);

#define LRCE_ACT_NONE 0
#define LRCE_ACT_GELU 1
#define LRCE_ACT_RELU 2

/* The whole summarisation-token walk of FusionTransformer.forward (fusionv3.py:41-51) plus final_fc (:195, ReLU :368) as
 * ONE kernel of row-sharded 16-CTA clusters (csrc/encoder_walk.cu): S segments x n_layers post-norm decoder layers on `rows`
 * tokens; every cluster owns <= 8 rows for the whole walk, streams the weights through TMA in a pre-packed order and runs the
 * products as swap-AB tcgen05 MMAs; all exchanges are cluster-local (no grid barrier).
 *
 * lrce_encoder_walk_pack (once per weight version) re-tiles the decoder weights into that streaming order. `layer_table` is
 * a DEVICE array of n_layers records of 16 device pointers to FP32 tensors, in this order:
 *   sa_w[768,768] (out_proj . v_proj of the length-1 self-attention, folded), q_w[768,768] (pre-scaled by 1/8),
 *   o_w[768,768], w1[3072,768], w2[768,3072]; sa_b, q_b, o_b, b1, b2, norm1 g/b, norm2 g/b, norm3 g/b.
 * fc_w fp32 [n_out, 768], fc_b fp32 [n_out] = final_fc. `packed`: lrce_encoder_walk_pack_bytes(n_layers, n_out) bytes,
 * 128-byte aligned; the LayerNorm in front of a product is folded into its weights and bias there (exact algebra).
 *
 * lrce_encoder_walk: kv_video bf16 [(rows/n_cand)*S*Tv, ld_kv], kv_text bf16 [rows*Lt, ld_kv] hold K at column
 * layer*1536 + head*64 and V at +768 (one lrce_gemm_bf16 per modality). tok0 fp32 [768] = summarization_token;
 * f_gamma/f_beta = fusion_layer_norm; act = LRCE_ACT_*; out fp32 [rows, n_out]; tokens_tap NULL or fp32 [S, rows, 768]
 * (the token after each segment). Tv + Lt <= 192 (every reference config: 150 + at most 41), n_out <= 4096. */
size_t lrce_encoder_walk_pack_bytes(int n_layers, int n_out);
int lrce_encoder_walk_pack(const void* layer_table, int n_layers, const float* fc_w, const float* fc_b, int n_out, void* packed,
                           void* stream);
int lrce_encoder_walk(const void* packed, int n_layers, const void* kv_video, const void* kv_text, int ld_kv, const float* tok0,
                      const float* f_gamma, const float* f_beta, float eps, int n_out, int act, float* out, float* tokens_tap,
                      int rows, int S, int Tv, int Lt, int n_cand, void* stream);
/* The host-side plan lrce_encoder_walk derives for `rows` rows when `max_clusters` 16-CTA clusters can be resident (7 on a
 * B200 at this kernel's shared-memory footprint): rows per cluster (<= 8), row groups, clusters launched, ring slots, dynamic
 * shared memory. Pure host arithmetic, no CUDA call. */
int lrce_encoder_walk_plan(int rows, int max_clusters, int* rows_per_cluster, int* n_groups, int* clusters, int* ring_slots,
                           int* smem_bytes);
/* Instrumented instantiation of the same kernel for tools/ (no library-side state: everything is a per-call argument):
 * prof = device int64 [CTAs][32] receiving the cycles thread 0 of each CTA spent per sub-step of the dependency chain;
 * max_clusters > 0 caps the number of clusters; variant = 1 runs the production code, other values select code variants kept
 * for same-box A/B measurements. */
int lrce_encoder_walk_profile(const void* packed, int n_layers, const void* kv_video, const void* kv_text, int ld_kv,
                              const float* tok0, const float* f_gamma, const float* f_beta, float eps, int n_out, int act,
                              float* out, float* tokens_tap, int rows, int S, int Tv, int Lt, int n_cand, void* stream,
                              long long* prof, int max_clusters, int variant);

/* ---- row / sequence kernels: BERT-base (text.py:5-17) and the encoder's training step (configs[4]) ----------------------
 * Dropout arguments: p (0 = off), a site id and a seed; the mask of element i is a counter-based hash of (seed, site, i), so
 * a backward call given the same triple regenerates the forward mask. `seed_dev` (device pointer or NULL) is XOR-ed into the
 * seed at run time, so launches captured in a CUDA graph draw a fresh mask on every replay. "T" outputs are transposed bf16 copies
 * T[col * ldT + rowT0 + row]: the K-major A / W operands of the weight-gradient GEMMs dW = dY^T X. */

/* y = drop_out(LayerNorm(res + drop_a(a))) on 768-wide fp32 rows; res_bcast != 0: res is ONE row used for every row. Any of
 * u_out (pre-LN sum), y_f32, y_bf16, yT may be NULL. Post-norm residual steps: nn.TransformerDecoderLayer norm1-3 and
 * fusion_layer_norm (fusionv3.py:8-17, :47-49), BertSelfOutput / BertOutput LayerNorm. */
int lrce_add_ln_768(const float* a, const float* res, int res_bcast, const float* gamma, const float* beta, float eps,
                    float* u_out, float* y_f32, void* y_bf16, void* yT, int ldT, int rowT0, long long n, float p_a, int site_a,
                    float p_out, int site_out, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* backward of the step above: dy = (dy_a + dy_b) * mask_out; du fp32 (residual path), dub / dubT = du * mask_a in bf16 (branch
 * path); dgamma / dbeta fp32 [768] are ACCUMULATED (atomics). u = the forward's u_out. */
int lrce_ln_bwd_768(const float* dy_a, const float* dy_b, const float* u, const float* gamma, float eps, float* du, void* dub,
                    void* dubT, int ldT, int rowT0, float* dgamma, float* dbeta, long long n, float p_a, int site_a, float p_out,
                    int site_out, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* fp32 [n, C] -> bf16 [n, C] (+ transposed copy): mode 0 cast, 1 GELU(erf) forward, 2 GELU backward (x = upstream gradient,
 * aux = forward pre-activation); dropout per element (group 1) or per `group` columns (64 = per attention head). */
int lrce_rows_f32_to_bf16(const float* x, const float* aux, void* y, void* yT, int ldT, int rowT0, long long n, int C, int mode,
                          int group, float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
int lrce_dropout_bf16(void* x, long long n_elems, float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* single-query cross attention of the summarisation token over [video segment `seg` ; text] (12 heads x 64; K at column
 * kcol + head*64 of the K/V GEMM output, V at + 768; q fp32 [R, 768] un-scaled). Forward keeps P fp32 [R*12, 256]. Backward
 * writes dq (bf16, x 1/8) and dK / dV into buffers laid out like kv_video / kv_text (video rows of a clip shared by
 * n_cand > 1 candidates are accumulated atomically: zero them first). nn.MultiheadAttention inside the decoder layer. */
int lrce_xattn_fwd(const float* q, const void* kv_video, const void* kv_text, int ld_kv, int kcol, int R, int S, int seg, int Tv,
                   int Lt, int n_cand, float* P, void* ctx, void* ctxT, int ldT, int rowT0, float p_drop, int site,
                   unsigned long long seed, const unsigned long long* seed_dev, void* stream);
int lrce_xattn_bwd(const float* q, const void* kv_video, const void* kv_text, int ld_kv, int kcol, int R, int S, int seg, int Tv,
                   int Lt, int n_cand, const float* P, const float* dctx, void* dq, void* dqT, int ldT, int rowT0, void* dkv_video,
                   void* dkv_text, float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* out[i] (+)= sum of the first `cols` entries of row i of src bf16 [rows, ld] (bias gradients from transposed operands) */
int lrce_rowsum_bf16(const void* src, int ld, int cols, float* out, int rows, int accumulate, void* stream);
/* out[c] += sum over rows of src[r, c] (fp32 or bf16 rows; out must be initialised) */
int lrce_colsum(const void* src, int is_bf16, long long rows, int cols, long long ld, float* out, void* stream);
int lrce_transpose_bf16(const void* src, long long rows, int cols, long long ld_src, void* dst, long long ld_dst, void* stream);
/* dst = a + b (+ c), bf16, c may be NULL */
int lrce_add_bf16(void* dst, const void* a, const void* b, const void* c, long long n_elems, void* stream);
/* backward of lrce_video_posembed_ln (is_text = 0) / lrce_text_posembed_ln (is_text = 1; B = rows, P = L, S = T = 1): dy fp32
 * gradient of the embedded rows -> dproj bf16 (video), table / LayerNorm gradients accumulated with atomics. */
int lrce_posembed_bwd(const float* dy, const void* proj, const void* text, int text_fp32, const float* emb_cls, const float* emb_pos,
                      const float* emb_len, const float* emb_clip, const float* gamma, float eps, void* dproj, float* d_cls,
                      float* d_pos, float* d_len, float* d_clip, float* dgamma, float* dbeta, int B, int S, int T, int P, int is_text,
                      float p_drop, int site, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* out fp32 [M, N] = A bf16 [M, K] x W bf16 [N, K]^T (+ bias): the token-path products of the training step (M = a few dozen
 * rows, K = 768 or 3072), bound by streaming W once; a 128-row tensor-core tile would be mostly padding. */
int lrce_gemm_skinny_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, float* out, int ldo,
                          void* stream);
/* BertEmbeddings: LayerNorm(word[ids] + position[l] + token_type[type_ids]) -> fp32 + bf16 [n, 768]; n = n_seq * L */
int lrce_bert_embed_ln(const long long* ids, const long long* type_ids, const float* word, const float* pos, const float* type,
                       const float* gamma, const float* beta, float eps, float* out_f32, void* out_bf16, long long n, int L,
                       int vocab, int n_types, int max_pos, void* stream);
/* BertSelfAttention core: qkv bf16 [n_seq*L, 3*64*n_heads] ([q|k|v]) -> out bf16 [n_seq*L, 64*n_heads]; keys with
 * mask[seq, j] == 0 are excluded (mask int64 [n_seq, L] or NULL); L <= 64 */
int lrce_bert_attention(const void* qkv, const long long* mask, void* out, int n_seq, int L, int n_heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LRCE_B200_H_ */
