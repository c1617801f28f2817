/*
 * mmvid_b200 - C ABI of the B200-native MMVID token-generation hot path.
 *
 * The reference (snap-research/MMVID) has NO operator / FFI boundary: every FLOP is a stock PyTorch call
 * made from three Python classes.  This header therefore *defines* the boundary the
 * replacement exports; each entry point cites the reference code whose compute it replaces.  The Python
 * classes in mmvid_b200/ (same names, signatures and state-dict keys as the reference's BERT / DALLE /
 * VQGanVAE1024 / OpenAICLIPTransformer) call these through ctypes with raw device pointers.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; pointers are DEVICE pointers unless named host_*;
 *   - no allocation, no ownership transfer: outputs / scratch are caller-allocated;
 *   - stream-ordered on `stream` (a cudaStream_t); re-entrant across streams; safe to capture in CUDA graphs;
 *   - return 0 on success, negative MMVID_E* otherwise; mmvid_last_error() gives a thread-local message;
 *   - activations are fp32 unless a dtype argument says otherwise; `precision` selects the math pipe:
 *       MMVID_FP32  CUDA-core FFMA, fp32 accumulate  (parity / bit-exact-index mode)
 *       MMVID_TF32  tcgen05.mma kind::tf32, fp32 accumulate in TMEM (<=1e-3 logits parity mode)
 *       MMVID_BF16  tcgen05.mma kind::f16 (bf16 operands), fp32 accumulate in TMEM (wide-range 16-bit mode; 8-bit mantissa)
 *       MMVID_F16   tcgen05.mma kind::f16 (fp16 operands: the SAME 10-bit mantissa as tf32 at twice its rate and half its
 *                   bytes), fp32 accumulate in TMEM; 16-bit stores saturate at +-65504 (default throughput mode)
 *   16-bit tensors carry MMVID_DT_BF16 or MMVID_DT_F16; operand dtypes must match the precision.
 */
#ifndef MMVID_B200_H
#define MMVID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mmvid_stream_t; /* cudaStream_t */

enum { MMVID_OK = 0, MMVID_EINVAL = -1, MMVID_ECUDA = -2, MMVID_EUNSUPPORTED = -3 };
enum { MMVID_FP32 = 0, MMVID_TF32 = 1, MMVID_BF16 = 2, MMVID_F16 = 3 };
enum { MMVID_ACT_NONE = 0, MMVID_ACT_QUICKGELU = 1, MMVID_ACT_SWISH = 2 };
enum { MMVID_MASK_NONE = 0, MMVID_MASK_CAUSAL = 1, MMVID_MASK_PREV = 2 };
enum { MMVID_DT_F32 = 0, MMVID_DT_BF16 = 1, MMVID_DT_F16 = 2 };

int mmvid_version(void);
const char* mmvid_last_error(void);
/* number of kernels this library has launched since load / last reset (bench.py "gpu_launches") */
long long mmvid_launch_count(void);
void mmvid_reset_launch_count(void);
/* accounts for n kernels of this library that were replayed from a captured CUDA graph (they bypass the entry points) */
void mmvid_add_launch_count(long long n);

/* ------------------------------------------------------------------------------------------------
 * K1  fused token + positional embedding gather  (reference: dalle_bert.py:903-972,1032-1036;
 *     dalle_artv.py:441-491: ~7 nn.Embedding gathers + adds + torch.cat)
 * Writes rows [seq_off, seq_off+n) of the residual stream out[B, S, D]:
 *     id = ids[b*ids_bstride + i];  if (id == pad_value) id = pad_base + i;      (unique pad ids, :917-919)
 *     out[b, seq_off+i, :] = table[id, :] + (pos ? pos[i, :] : 0) + (table2 ? table2[id, :] : 0)
 * ids: int64.  table2 covers special_emb + special_pos_emb indexed by the same id (:903-906).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const int64_t* ids; long long ids_bstride; int n; int seq_off;
  const float* table; const float* table2; const float* pos;
  long long pad_value; long long pad_base; int use_pad;
  long long table_rows; /* rows of table (and table2): an id outside [0, table_rows) traps the kernel, which surfaces as a
                           CUDA error at the next synchronisation (torch's nn.Embedding device assert); 0 = unchecked */
} mmvid_embed_segment;
int mmvid_embed_gather(float* out, int B, int S, int D, const mmvid_embed_segment* host_segments, int num_segments,
                       mmvid_stream_t stream);

/* axial positional table: out[i, :] = sum_a w_a[coord_a(i), :], i < n  (axial_positional_embedding, summed
 * in axis order ((w0 + w1) + w2) like the package's python sum()).  shape[] = axis lengths (naxes <= 3). */
int mmvid_axial_table(float* out, int n, int D, const float* w0, const float* w1, const float* w2,
                      const int* host_shape, int naxes, mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2  LayerNorm over the last dim (clip_model.py:188-193 eps 1e-5; to_logits.0 dalle_bert.py:414-425)
 *     out dtype fp32 or bf16 (bf16 feeds the kind::f16 GEMMs).  x row stride = ldx elements.
 * ---------------------------------------------------------------------------------------------- */
int mmvid_layernorm(const float* x, long long ldx, const float* gamma, const float* beta, void* out, int out_dtype,
                    long long rows, int D, float eps, mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3/K5/K6/K7  Linear:  C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N])
 *     (F.linear inside nn.MultiheadAttention in/out-proj clip_model.py:208,222; mlp c_fc/c_proj :210-213;
 *      to_logits.1 dalle_bert.py:416; 1x1 convs model.py:124,159-178, vqgan.py:41-43)
 * a_dtype / w_dtype: MMVID_DT_F32 or MMVID_DT_BF16 (bf16 only with MMVID_BF16).  c_dtype likewise.
 * All leading dimensions in elements.  residual may alias C.
 * ---------------------------------------------------------------------------------------------- */
int mmvid_linear(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                 const float* bias, const float* residual, long long ldr, void* C, int c_dtype, long long ldc,
                 long long M, int N, int K, int act, int precision, mmvid_stream_t stream);

/* Strided-batched fp32 GEMM on CUDA cores (parity path of attention / spatial attention):
 *   C[b1,b2][m,n] = alpha * sum_k A[b1,b2][m,k] * B[b1,b2](k,n),   B(k,n) at  B + n*ldb_n + k*ldb_k */
int mmvid_gemm_batched_f32(const float* A, long long lda, long long a_s1, long long a_s2,
                           const float* B, long long ldb_n, long long ldb_k, long long b_s1, long long b_s2,
                           float* C, long long ldc, long long c_s1, long long c_s2,
                           int M, int N, int K, int batch1, int batch2, float alpha, mmvid_stream_t stream);

/* In-place masked row softmax of scores[batch, rows, cols] (fp32 path of F.scaled_dot_product_attention /
 * AttnBlock softmax model.py:193): mask CAUSAL: col <= row; MASK_PREV: rows listed in prev_rows (device int32,
 * n_prev of them) see only cols >= row (clip_model.py:571-575); others see everything. */
int mmvid_softmax_rows(float* scores, long long batch, int rows, int cols, long long ld, int mask_kind,
                       const int* prev_rows, int n_prev, mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3-K5  multi-head attention core on tensor cores (F.scaled_dot_product_attention, clip_model.py:219-222):
 *   out[b, s, h*64:(h+1)*64] = softmax(Q K^T / 8 + mask) V        head_dim fixed at 64 (CLIP)
 *   q, k: [B, H, S_pad, 64]; vt: [B, H, 64, S_pad] (V transposed; produced by mmvid_linear's QKV epilogue
 *   via mmvid_qkv_split); S_pad = S rounded up to 128.  dtype fp32 (TF32) or bf16 (BF16).
 * ---------------------------------------------------------------------------------------------- */
int mmvid_attention(const void* q, const void* k, const void* vt, void* out, int out_dtype, long long ldo,
                    int B, int H, int S, int S_pad, int mask_kind, const int* host_prev_rows, int n_prev,
                    int precision, mmvid_stream_t stream);
/* Fused QKV projection (nn.MultiheadAttention in-proj, clip_model.py:208,222): A[B*S, H*64] . W[3*H*64, H*64]^T + b
 * scattered by the GEMM epilogue into q,k [B,H,S_pad,64] and vt [B,H,64,S_pad] (fp32 for TF32, bf16 for BF16).
 * Padding (rows >= S) is not written: keep the buffers zero-initialised. */
int mmvid_linear_qkv(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                     const float* bias, void* q, void* k, void* vt, int out_dtype, int B, int H, int S, int S_pad,
                     int precision, mmvid_stream_t stream);
/* qkv [B*S, 3*H*64] fp32 -> q,k [B,H,S_pad,64], vt [B,H,64,S_pad] in fp32 or bf16 (zero padded) */
int mmvid_qkv_split(const float* qkv, void* q, void* k, void* vt, int dtype, int B, int H, int S, int S_pad,
                    mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K9  ART-V KV-cache decode (new capability; the reference re-runs the full prefix, dalle_artv.py:258-281)
 *   single-query attention against cached K/V: q [B,H,64], kcache/vcache [B,H,S_max,64], len = #valid keys
 * ---------------------------------------------------------------------------------------------- */
int mmvid_decode_attention(const float* q, long long q_bstride, const float* kcache, const float* vcache,
                           float* out, long long o_bstride, int B, int H, int S_max, int len, mmvid_stream_t stream);
/* GEMV-style linear for tiny M (decode): C[M,N] = act(A[M,K].W[N,K]^T + bias) (+ residual); M <= 16 */
int mmvid_linear_small_m(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                         const float* residual, long long ldr, float* C, long long ldc, int M, int N, int K, int act,
                         mmvid_stream_t stream);
/* append k,v rows of qkv[B, 3*H*64] into the caches at position pos */
int mmvid_kv_append(const float* qkv, long long qkv_bstride, float* kcache, float* vcache, int B, int H, int S_max,
                    int pos, mmvid_stream_t stream);

/* One full decode step (all layers, B <= 16 new tokens) issued natively: h [B,D] updated in place.
 * ws: caller workspace of mmvid_artv_decode_workspace_floats(B, D, H) floats. */
typedef struct {
  const float *ln1_w, *ln1_b, *in_w, *in_b, *out_w, *out_b, *ln2_w, *ln2_b, *fc_w, *fc_b, *proj_w, *proj_b;
  float *kcache, *vcache; /* [B, H, S_max, 64] */
} mmvid_decode_layer;
long long mmvid_artv_decode_workspace_floats(int B, int D, int H);
int mmvid_artv_decode_step(const mmvid_decode_layer* host_layers, int n_layers, float* h, float* ws, int B, int D, int H,
                           int S_max, int pos, mmvid_stream_t stream);
/* Persistent variant: ONE cooperative launch for all layers + the LN/image-logit head (grid-wide barriers between
 * phases instead of ~110 launches per token).  B <= 8, n_layers <= 24.  head_w: [n_logits, D] rows of to_logits.1. */
int mmvid_artv_decode_persistent(const mmvid_decode_layer* host_layers, int n_layers, float* h, float* ws,
                                 const float* head_ln_w, const float* head_ln_b, const float* head_w,
                                 const float* head_b, float* logits, int n_logits, int B, int D, int H, int S_max,
                                 int pos, mmvid_stream_t stream);
/* Same contract, third generation (decode_pdl.cu): 5 fused launches per layer (LayerNorm, cache append, split-KV
 * combine, bias / QuickGELU / residual folded in) chained with programmatic dependent launch so that each kernel streams
 * its weights from HBM while its predecessor is still running.  ws must be ZERO-INITIALISED once by the caller. */
int mmvid_artv_decode_fused(const mmvid_decode_layer* host_layers, int n_layers, float* h, float* ws,
                                 const float* head_ln_w, const float* head_ln_b, const float* head_w,
                                 const float* head_b, float* logits, int n_logits, int B, int D, int H, int S_max,
                                 int pos, mmvid_stream_t stream);

/* Fourth generation (decode_stream.cu): ONE persistent cooperative launch per token.  Every CTA owns a fixed column slice of
 * every weight matrix and streams its slabs of the next phases into a shared-memory ring with bulk async copies while the
 * current phase computes; phases are separated by a light grid barrier.  Weights and K/V caches are 16-bit (f16 != 0: fp16,
 * else bf16), accumulation / LayerNorm / softmax / residual stream fp32.  ws: mmvid_artv_decode_stream_workspace_floats
 * floats, ZERO-INITIALISED once by the caller.  Returns 1 when the shape does not fit the kernel's shared-memory plan
 * (callers fall back to mmvid_artv_decode_fused on fp32 weights). */
typedef struct {
  const float *ln1_w, *ln1_b, *in_b, *out_b, *ln2_w, *ln2_b, *fc_b, *proj_b;
  const void *in_w, *out_w, *fc_w, *proj_w; /* 16-bit, [N, K] row-major */
  void *kcache, *vcache;                    /* [B, H, S_max, 64] 16-bit */
} mmvid_decode_layer16;
long long mmvid_artv_decode_stream_workspace_floats(int B, int D, int H);
int mmvid_artv_decode_stream(const mmvid_decode_layer16* host_layers, int n_layers, float* h, float* ws,
                             const float* head_ln_w, const float* head_ln_b, const void* head_w16, const float* head_b,
                             float* logits, int n_logits, int B, int D, int H, int S_max, int pos,
                             const int* pos_dev /* optional device step counter added to pos (graph replay) */, int f16,
                             mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K14  VQ nearest-codeword lookup (taming/modules/vqvae/quantize.py:302-311):
 *   d[t,j] = (sum z_t^2 + sum e_j^2) - 2 z_t.e_j   in fp32, this association; idx[t] = argmin_j (lowest index wins)
 *   z: [T, dim] (NHWC rows); codebook [n_codes, dim]; idx int64 [T]
 *   e2_scratch: caller-owned device floats [n_codes] (||e||^2; callers on different streams pass different buffers)
 * ---------------------------------------------------------------------------------------------- */
int mmvid_vq_argmin(const float* z, const float* codebook, int64_t* idx, float* e2_scratch, long long T, int n_codes,
                    int dim, mmvid_stream_t stream);

/* K15 codebook gather for decode (vae.py:50-52): out[t, :] = codebook[ids[t], :]  (NHWC rows) */
int mmvid_codebook_gather(const int64_t* ids, const float* codebook, float* out, long long T, int dim,
                          mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K10/K13  Conv2d on NHWC fp32 activations (taming/modules/diffusionmodules/model.py:102-115 etc.)
 *   w packed [Cout, KH, KW, Cin].  out[n,y,x,co] = bias[co] + sum in[n, y*stride+ky-pad_t, x*stride+kx-pad_l, ci] w
 *   (+ residual[n,y,x,co]).  upsample=1: input is read through a nearest x2 upsample (Upsample, model.py:56-62).
 *   Downsample (model.py:77-84) = stride 2, pad_t=pad_l=0 with zero fill beyond the right/bottom edge.
 *   in_nchw / out_nchw: read / write NCHW instead (first / last layer of the VQGAN; fuses vae.py:41 `2x-1`
 *   when pre_affine=1 and vae.py:55 clamp(-1,1)*0.5+0.5 when post_clamp=1).
 *   precision MMVID_TF32 / MMVID_F16: implicit GEMM on tcgen05 whose pixel tiles are fetched by 4-D TMA boxes (halo taps =
 *   TMA zero fill); MMVID_F16 takes fp16 activations (written by mmvid_groupnorm / mmvid_upsample2x) and fp16 packed
 *   weights, accumulates in fp32 and writes fp32 (bias / residual fp32): the tf32 mantissa at twice the MMA rate.
 *   gn_partial (optional): the tensor-core conv ALSO writes the GroupNorm partial statistics of its result (after bias and
 *   residual; model.py:33-42 `Normalize` is what consumes every conv output of the VQGAN): floats
 *   [N*Ho*Wo/32][gn_groups][2] = per 32-pixel slab and group (sum, sum of squared deviations from the slab mean); mmvid_groupnorm_from_partials turns them
 *   into the normalised tensor without another statistics pass over the activation.  Only where mmvid_conv2d_gn_fusable()
 *   returns 1 (tensor-core path, H*W % 128 == 0, Cout % 32 == 0, 4 / 8 / 16 channels per group); refused otherwise.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* in; const void* w; /* fp32; fp16 when precision == MMVID_F16 (tensor-core conv: stride 1, NHWC, Cin % 64 == 0) */
  const float* bias; const float* residual; float* out;
  int N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, upsample;
  int in_nchw, out_nchw, pre_affine, post_clamp, precision;
  float* gn_partial; int gn_groups;
} mmvid_conv_params;
int mmvid_conv2d(const mmvid_conv_params* p, mmvid_stream_t stream);
int mmvid_conv2d_gn_fusable(const mmvid_conv_params* p);

/* K11 with the statistics taken from a conv's fused partial sums (see mmvid_conv_params::gn_partial): partial holds
 *   [N*HW/32][groups][2] floats, stats is a caller-owned scratch of N*groups*2 floats; otherwise as mmvid_groupnorm. */
int mmvid_groupnorm_from_partials(const float* in, void* out, int out_dtype, const float* gamma, const float* beta,
                                  const float* partial, float* stats, int N, int HW, int C, int groups, float eps, int swish,
                                  mmvid_stream_t stream);

/* K11 GroupNorm(32 groups, eps) + optional swish on NHWC (model.py:38-42, 33-35):
 *   stats scratch: mmvid_groupnorm_scratch_floats(N, groups) floats.  out may alias in.
 *   swish: 0 = none, 1 = x*sigmoid(x) with expf and IEEE division (fp32 parity mode), 2 = MUFU ex2 / rcp (~1e-6 relative).
 *   out_dtype: MMVID_DT_F32, or MMVID_DT_F16 when the result feeds a kind::f16 conv (then out must not alias in). */
int mmvid_groupnorm(const float* in, void* out, int out_dtype, const float* gamma, const float* beta,
                    float* stats_scratch, int N, int HW, int C, int groups, float eps, int swish, mmvid_stream_t stream);

/* size (in floats) of the stats scratch mmvid_groupnorm / mmvid_conv_out_fused need for N images */
long long mmvid_groupnorm_scratch_floats(int N, int groups);
/* statistics only: stats[n, g] = (mean, rstd) at the start of the scratch buffer */
int mmvid_groupnorm_stats(const float* in, float* stats_scratch, int N, int HW, int C, int groups, float eps,
                          mmvid_stream_t stream);

/* Decoder tail fused (model.py:578-581 + vae.py:55): GroupNorm + swish + 3x3 conv to Cout <= 4 channels
 * (+ clamp(-1,1)*0.5+0.5 when post_clamp bit 0 is set), NHWC in, NCHW out.  w packed [Cout,3,3,C].
 * post_clamp bit 1: swish through MUFU ex2 / rcp instead of expf + IEEE division. */
int mmvid_conv_out_fused(const float* in, const float* gamma, const float* beta, const float* w, const float* bias,
                         float* out, float* stats_scratch, int N, int H, int W, int C, int Cout, int groups, float eps,
                         int post_clamp, mmvid_stream_t stream);
/* ... with norm_out's statistics taken from the producing conv's fused partial sums (mmvid_conv_params::gn_partial);
 * stats: caller-owned scratch of N*groups*2 floats */
int mmvid_conv_out_fused_from_partials(const float* in, const float* gamma, const float* beta, const float* w,
                                       const float* bias, float* out, const float* partial, float* stats, int N, int H,
                                       int W, int C, int Cout, int groups, float eps, int post_clamp, mmvid_stream_t stream);

/* nearest x2 upsample NHWC (used only when not fused into the conv) */
int mmvid_upsample2x(const float* in, void* out, int out_dtype /* F32 | F16 */, int N, int H, int W, int C,
                     mmvid_stream_t stream);

/* elementwise helpers: NCHW<->NHWC transposes of small tensors */
int mmvid_nchw_to_nhwc(const float* in, float* out, int N, int C, int HW, mmvid_stream_t stream);
int mmvid_nhwc_to_nchw(const float* in, float* out, int N, int C, int HW, mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K7/K8 sampling support: row softmax of logits [rows, n] -> probs (fp32), optional additive noise
 * (dalle_bert.py:527-531 `logits + temperature * gumbel`), temperature applied by caller via noise scale */
int mmvid_softmax_logits(const float* logits, const float* noise, float noise_scale, float* probs, long long rows,
                         int n, mmvid_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training path (BERT.forward(return_loss=True), dalle_bert.py:980-1127; optimiser loop train.py:320-325).
 * GEMM-shaped backward work reuses mmvid_linear / mmvid_gemm_batched_f32 on transposed operands; these are the
 * HBM-bound backward pieces.
 * ---------------------------------------------------------------------------------------------- */
int mmvid_act_forward(const float* z, float* y, long long n, int act, mmvid_stream_t stream);
int mmvid_act_backward(const float* z, const float* dy, float* dz, long long n, int act, mmvid_stream_t stream);
/* out[c] (+)= sum_r x[r, c]; scratch >= 64*cols floats; deterministic two-stage reduction (bias gradients) */
int mmvid_colsum(const float* x, float* out, float* scratch, long long rows, int cols, int accumulate,
                 mmvid_stream_t stream);
/* LayerNorm backward: dx, and xhat*dy rows (dgamma = colsum(xhat_dy), dbeta = colsum(dy)) */
int mmvid_layernorm_backward(const float* x, const float* gamma, const float* dy, float* dx, float* xhat_dy,
                             long long rows, int D, float eps, mmvid_stream_t stream);
/* in place: dp[r,c] = p[r,c] * (dp[r,c] - sum_c' dp[r,c'] p[r,c']) * scale */
int mmvid_softmax_backward(const float* p, float* dp, long long rows, int cols, long long ld, float scale,
                           mmvid_stream_t stream);
/* F.cross_entropy over selected rows (dalle_bert.py:1040): loss_rows[r] = lse - logit[target], dlogits = softmax - onehot */
int mmvid_cross_entropy(const float* logits, const int64_t* target, const uint8_t* sel, float* dlogits,
                        float* loss_rows, long long rows, int n, mmvid_stream_t stream);
/* embedding backward (scatter-add with atomics) for one gather segment */
int mmvid_embed_backward(const float* dx, int B, int S, int D, const mmvid_embed_segment* host_segment, float* d_table,
                         float* d_table2, float* d_pos, mmvid_stream_t stream);
int mmvid_transpose2d(const float* in, float* out, int R, int C, mmvid_stream_t stream);

/* Profiling hook (not part of the data path): CTA (0,0) of every following mmvid_attention launch writes clock64()
 * stamps of its pipeline events into dev_buf (>= 1024 uint64; NULL switches it off).  See scripts/att_trace3.py. */
int mmvid_debug_attention_trace(unsigned long long* dev_buf);

/* ------------------------------------------------------------------------------------------------
 * Device-resident mask-predict sampling (BERT.mask_predict, dalle_bert.py:527-538, 646-691; sampling_mode 'batched').
 * Philox4x32-10 keyed by (seed, row | item, offset): same seed / offset => same draw.
 *   mmvid_mp_sample: per row of n logits (n % 128 == 0, <= 1024):  probs = softmax(logits + noise_scale * gumbel);
 *     tok ~ Categorical(probs);  Y = probs[tok].  Rows with skip[row] != 0 keep their Y / tok (tokens kept from the
 *     previous iteration).  skip may be NULL.
 *   mmvid_mp_keep: per (sample, beam): keep k of the tokens with pmask == 0 WITHOUT replacement, probability proportional
 *     to Y (Gumbel-top-k == torch.multinomial(Y, k, replacement=False) in distribution); tokens with pmask != 0 are always
 *     kept.  keep [samples*beams, Ttot] (uint8), ids_in = keep ? I_tok : mask_id.  Y, I_tok: [samples, Ttot].
 * ---------------------------------------------------------------------------------------------- */
int mmvid_mp_sample(const float* logits, long long rows, int n, float noise_scale, const uint8_t* skip, float* Y,
                    int64_t* tok, unsigned long long seed, unsigned long long offset,
                    const int* step_dev /* optional device counter added to offset (CUDA-graph replays) */,
                    mmvid_stream_t stream);
int mmvid_mp_keep(const float* Y, const uint8_t* pmask, const int64_t* I_tok, int samples, int beams, int Ttot, int k,
                  long long mask_id, uint8_t* keep, int64_t* ids_in, unsigned long long seed, unsigned long long offset,
                  mmvid_stream_t stream);

/* Profiling hook: CTA 0 of every following mmvid_artv_decode_stream launch writes %globaltimer (ns) stamps of its phase
 * events into dev_buf (>= 512 uint64; NULL switches it off).  See scripts/decode_trace.py. */
int mmvid_debug_decode_trace(unsigned long long* dev_buf);

/* Profiling hook: CTA 0 of every following single-CTA tensor-core GEMM (mmvid_linear in TF32 / BF16 precision) writes
 * clock64() stamps of its first 8 tiles into dev_buf (>= 512 uint64; NULL switches it off).  See scripts/gemm_trace.py. */
int mmvid_debug_gemm_trace(unsigned long long* dev_buf);

/* Host-only hook (no launch): the tile mmvid_linear picks for a plain tensor-core GEMM of this shape: 2000 + BN = CTA-pair
 * kernel with a 256 x BN tile, 1000 + BN = single-CTA kernel with a 128 x BN tile.  precision: MMVID_TF32 | MMVID_BF16,
 * c_dtype: MMVID_DT_F32 | MMVID_DT_BF16. */
int mmvid_debug_pick_tile(long long M, int N, int K, int precision, int c_dtype);

/* Profiling hook: clock64() cost of n_mma back-to-back tcgen05.mma of one shape on one SM (flavors: see
 * csrc/debug_mma_rate.cu); dev_out[0] = issue cycles, dev_out[1] = cycles until the commit barrier fires. */
int mmvid_debug_mma_rate(int flavor, int n_mma, unsigned long long* dev_out, mmvid_stream_t stream);

/* K18 optimiser step (train.py:322-325: opt.zero_grad / loss.backward / clip_grad_norm_ / opt.step; optimisers
 * utils_train.py:167-181: torch.optim.Adam(lr, weight_decay) or AdamW(betas=(0.9, 0.95))).  Multi-tensor: the caller
 * builds, once, a device table with one record per parameter tensor and a chunk map (chunk c covers elements
 * [chunk_index[c]*chunk_elems, +chunk_elems) of tensor chunk_tensor[c]); every call is one launch over all chunks.
 * A record with g == NULL is skipped (parameter without gradient).  All tensors fp32, contiguous. */
typedef struct {
  void* p;        /* parameter                */
  const void* g;  /* gradient (may be NULL)   */
  void* m;        /* exp_avg                  */
  void* v;        /* exp_avg_sq               */
  long long n;    /* elements                 */
  long long skipped; /* optimiser steps this tensor had no gradient for: its own step is `step - skipped` */
} mmvid_adam_tensor;
/* total_sq[0] = sum over all gradients of g^2; deterministic (per-chunk partials, ordered finalize).
 * partial_dev: n_chunks floats of scratch. */
int mmvid_grad_sqnorm(const mmvid_adam_tensor* table_dev, const int* chunk_tensor_dev, const int* chunk_index_dev,
                      int n_chunks, int chunk_elems, float* partial_dev, float* total_sq_dev, mmvid_stream_t stream);
/* torch.nn.utils.clip_grad_norm_: g *= min(1, max_norm / (sqrt(total_sq) + 1e-6)), in place, no host sync */
int mmvid_grad_clip(const mmvid_adam_tensor* table_dev, const int* chunk_tensor_dev, const int* chunk_index_dev,
                    int n_chunks, int chunk_elems, const float* total_sq_dev, float max_norm, mmvid_stream_t stream);
/* Adam (decoupled_weight_decay = 0) / AdamW (1) update of step `step` (1-based), torch semantics (eps outside the
 * bias-corrected sqrt, no amsgrad) */
int mmvid_adam_step(const mmvid_adam_tensor* table_dev, const int* chunk_tensor_dev, const int* chunk_index_dev,
                    int n_chunks, int chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int decoupled_weight_decay, int step, mmvid_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMVID_B200_H */
