// ART-V KV-cache decode path (K9) - a capability the reference lacks: DALLE.generate_images re-runs the
// whole prefix through all 12 layers for every sampled token (dalle_artv.py:258-281).  With a causal mask and
// additive absolute position embeddings a K/V cache is exact, so one decode step only needs
//   * skinny linears  C[M<=16, N] = act(A W^T + b) (+res)   - HBM-bound on W (read once, coalesced rows)
//   * single-query attention against the cache             - HBM-bound on K/V
// Both are CUDA-core kernels laid out for streaming bandwidth, not tensor-core shapes.
#include "common.cuh"

using namespace mmvid;

namespace {

// One slice of a GEMV column: `units` of 128 floats (32 lanes x float4) of weight row w4, 8 units (= 8 independent
// 16-byte loads per lane, 4 KB per warp) issued before the first FMA so that the row streams at memory speed instead of
// one L2/HBM round trip per 512 bytes.  a(b, k4) returns the float4 of activation row b at float4 index k4.
template <int MAXB, typename ALoad>
__device__ __forceinline__ void gemv_slice(const float4* __restrict__ w4, int k4_end, int u0, int u1, int lane, int B,
                                           ALoad a, float (&acc)[MAXB]) {
  for (int u = u0; u < u1; u += 8) {
    float4 w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k4 = (u + j) * 32 + lane;
      w[j] = (u + j < u1 && k4 < k4_end) ? __ldg(w4 + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k4 = (u + j) * 32 + lane;
      if (u + j < u1 && k4 < k4_end) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
          if (b < B) {
            const float4 x = a(b, k4);
            acc[b] = fmaf(x.x, w[j].x, acc[b]); acc[b] = fmaf(x.y, w[j].y, acc[b]);
            acc[b] = fmaf(x.z, w[j].z, acc[b]); acc[b] = fmaf(x.w, w[j].w, acc[b]);
          }
        }
      }
    }
  }
}

// Small-M linear: KS warps of a block share one output column (K split KS ways, combined through shared memory in a
// fixed order), the other 8/KS column groups take neighbouring columns; every warp streams its weight slice with 8
// independent 16-byte loads in flight per lane.  The host picks KS so that narrow layers still fill the 148 SMs.
template <int MAXM>
__global__ void __launch_bounds__(256) linear_small_m_kernel(const float* __restrict__ A, long long lda,
                                                            const float* __restrict__ W, long long ldw,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ residual, long long ldr,
                                                            float* __restrict__ C, long long ldc, int M, int N, int K,
                                                            int act, int KS) {
  __shared__ float sm_part[8][MAXM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cols_per_block = 8 / KS;
  const int sub = warp % KS, cw = warp / KS;
  const int n = blockIdx.x * cols_per_block + cw;
  const int units = (K + 127) >> 7;
  const int upw = (units + KS - 1) / KS;
  float acc[MAXM];
#pragma unroll
  for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
  if (n < N) {
    gemv_slice<MAXM>(reinterpret_cast<const float4*>(W + (long long)n * ldw), K >> 2, sub * upw, min(units, (sub + 1) * upw),
                     lane, M, [&](int m, int k4) { return __ldg(reinterpret_cast<const float4*>(A + m * lda) + k4); }, acc);
  }
  float mine = 0.f;  // lane m keeps row m's total
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    const float v = warp_sum(acc[m]);
    if (lane == m) mine = v;
  }
  if (KS > 1) {
    if (lane < MAXM) sm_part[warp][lane] = mine;
    __syncthreads();
    if (sub == 0 && lane < M) {
      mine = 0.f;
      for (int s2 = 0; s2 < KS; ++s2) mine += sm_part[warp + s2][lane];
    }
  }
  if (sub == 0 && lane < M && n < N) {
    float v = mine + (bias ? bias[n] : 0.f);
    v = apply_act(v, act);
    if (residual) v += residual[lane * ldr + n];
    C[lane * ldc + n] = v;
  }
}

// q [B, 3*H*64] layout (row of the fused QKV projection): append K,V of the new token to the caches.
__global__ void kv_append_kernel(const float* __restrict__ qkv, long long qkv_bstride, float* __restrict__ kc,
                                 float* __restrict__ vc, int H, int S_max, int pos) {
  const int b = blockIdx.y, h = blockIdx.x, d = threadIdx.x;  // 64 threads
  const int D = H * 64;
  const float* row = qkv + b * qkv_bstride;
  const long long dst = (((long long)b * H + h) * S_max + pos) * 64 + d;
  kc[dst] = row[D + h * 64 + d];
  vc[dst] = row[2 * D + h * 64 + d];
}

// Single-query attention: one block per (b, h); 8 warps split the keys, each lane owns 2 of the 64 dims.
__global__ void __launch_bounds__(256) decode_attention_kernel(const float* __restrict__ q, long long q_bstride,
                                                              const float* __restrict__ kc, const float* __restrict__ vc,
                                                              float* __restrict__ out, long long o_bstride, int H,
                                                              int S_max, int len) {
  __shared__ float sm_m[8], sm_l[8];
  __shared__ float sm_o[8][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float2 qv = *reinterpret_cast<const float2*>(q + b * q_bstride + h * 64 + lane * 2);
  const float* kb = kc + ((long long)b * H + h) * S_max * 64;
  const float* vb = vc + ((long long)b * H + h) * S_max * 64;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int s = warp; s < len; s += 8) {
    const float2 kv = *reinterpret_cast<const float2*>(kb + (long long)s * 64 + lane * 2);
    float dot = warp_sum(qv.x * kv.x + qv.y * kv.y) * 0.125f;
    const float m_new = fmaxf(m, dot);
    const float alpha = expf(m - m_new), p = expf(dot - m_new);
    const float2 vv = *reinterpret_cast<const float2*>(vb + (long long)s * 64 + lane * 2);
    l = l * alpha + p;
    o0 = o0 * alpha + p * vv.x;
    o1 = o1 * alpha + p * vv.y;
    m = m_new;
  }
  if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
  sm_o[warp][lane * 2] = o0; sm_o[warp][lane * 2 + 1] = o1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float M = -INFINITY;
    for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w]);
    float L = 0.f, O = 0.f;
    for (int w = 0; w < 8; ++w) {
      const float sc = (sm_m[w] == -INFINITY) ? 0.f : expf(sm_m[w] - M);
      L += sm_l[w] * sc;
      O += sm_o[w][threadIdx.x] * sc;
    }
    out[b * o_bstride + h * 64 + threadIdx.x] = O / L;
  }
}

// split-KV variant: block (h, b, z) handles keys z, z+Z, ... and writes an unnormalised partial (m, l, o[64])
__global__ void __launch_bounds__(256) decode_attention_split_kernel(const float* __restrict__ q, long long q_bstride,
                                                                    const float* __restrict__ kc, const float* __restrict__ vc,
                                                                    float* __restrict__ part, int H, int S_max, int len,
                                                                    int Z) {
  __shared__ float sm_m[8], sm_l[8];
  __shared__ float sm_o[8][64];
  const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float2 qv = *reinterpret_cast<const float2*>(q + b * q_bstride + h * 64 + lane * 2);
  const float* kb = kc + ((long long)b * H + h) * S_max * 64;
  const float* vb = vc + ((long long)b * H + h) * S_max * 64;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int s = z * 8 + warp; s < len; s += 8 * Z) {
    const float2 kv = *reinterpret_cast<const float2*>(kb + (long long)s * 64 + lane * 2);
    const float dot = warp_sum(qv.x * kv.x + qv.y * kv.y) * 0.125f;
    const float m_new = fmaxf(m, dot);
    const float alpha = expf(m - m_new), p = expf(dot - m_new);
    const float2 vv = *reinterpret_cast<const float2*>(vb + (long long)s * 64 + lane * 2);
    l = l * alpha + p;
    o0 = o0 * alpha + p * vv.x;
    o1 = o1 * alpha + p * vv.y;
    m = m_new;
  }
  if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
  sm_o[warp][lane * 2] = o0; sm_o[warp][lane * 2 + 1] = o1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float M = -INFINITY;
    for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w]);
    float L = 0.f, O = 0.f;
    for (int w = 0; w < 8; ++w) {
      const float sc = (sm_m[w] == -INFINITY) ? 0.f : expf(sm_m[w] - M);
      L += sm_l[w] * sc;
      O += sm_o[w][threadIdx.x] * sc;
    }
    float* dst = part + ((((long long)b * H + h) * Z + z) * 66);
    dst[2 + threadIdx.x] = O;
    if (threadIdx.x == 0) { dst[0] = M; dst[1] = L; }
  }
}
__global__ void decode_attention_combine_kernel(const float* __restrict__ part, float* __restrict__ out, long long o_bstride,
                                                int H, int Z) {
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;  // 64 threads
  const float* p = part + (((long long)b * H + h) * Z) * 66;
  float M = -INFINITY;
  for (int z = 0; z < Z; ++z) M = fmaxf(M, p[z * 66]);
  float L = 0.f, O = 0.f;
  for (int z = 0; z < Z; ++z) {
    const float sc = (p[z * 66] == -INFINITY) ? 0.f : expf(p[z * 66] - M);
    L += p[z * 66 + 1] * sc;
    O += p[z * 66 + 2 + d] * sc;
  }
  out[b * o_bstride + h * 64 + d] = O / L;
}

}  // namespace

extern "C" int mmvid_linear_small_m(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                                    const float* residual, long long ldr, float* C, long long ldc, int M, int N, int K,
                                    int act, mmvid_stream_t stream) {
  MMVID_REQUIRE(M >= 1 && M <= 16, "1 <= M <= 16");
  MMVID_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "K, lda, ldw multiples of 4");
  // K split: enough (column, slice) warps for ~2 blocks per SM, at most one slice per 128-float unit
  const int units = (K + 127) / 128;
  int KS = 1;
  while (KS < 8 && (long long)N * KS < 8LL * 2 * 148 && units >= KS * 2) KS *= 2;
  dim3 grid(ceil_div(N, 8 / KS));
  cudaStream_t st = to_stream(stream);
  if (M <= 4) linear_small_m_kernel<4><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act, KS);
  else if (M <= 8) linear_small_m_kernel<8><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act, KS);
  else linear_small_m_kernel<16><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act, KS);
  return check_launch("linear_small_m");
}

extern "C" int mmvid_kv_append(const float* qkv, long long qkv_bstride, float* kcache, float* vcache, int B, int H,
                               int S_max, int pos, mmvid_stream_t stream) {
  MMVID_REQUIRE(pos >= 0 && pos < S_max, "pos in range");
  kv_append_kernel<<<dim3(H, B), 64, 0, to_stream(stream)>>>(qkv, qkv_bstride, kcache, vcache, H, S_max, pos);
  return check_launch("kv_append");
}

extern "C" int mmvid_decode_attention(const float* q, long long q_bstride, const float* kcache, const float* vcache,
                                      float* out, long long o_bstride, int B, int H, int S_max, int len,
                                      mmvid_stream_t stream) {
  MMVID_REQUIRE(len >= 1 && len <= S_max, "1 <= len <= S_max");
  decode_attention_kernel<<<dim3(H, B), 256, 0, to_stream(stream)>>>(q, q_bstride, kcache, vcache, out, o_bstride, H,
                                                                    S_max, len);
  return check_launch("decode_attention");
}



// ------------------------------------------------------------------------------------------------
// One whole ART-V decode step (all transformer layers for the B newly sampled tokens) issued natively: 8 launches per
// layer with no Python / allocator round trips in between.  h [B, D] is updated in place; ws is a caller-owned
// workspace of mmvid_artv_decode_workspace_floats(B, D, H) floats.  Requires B <= 16.
// ------------------------------------------------------------------------------------------------
extern "C" long long mmvid_artv_decode_workspace_floats(int B, int D, int H) {
  const long long items = (long long)B * H * 16 > 4096 ? (long long)B * H * 16 : 4096;  // split-KV partial slots
  return (long long)B * (D + 3 * D + D + 4 * D) + items * 66 + 1024;  // + split-KV arrival counters (decode_pdl.cu)
}

extern "C" int mmvid_artv_decode_step(const mmvid_decode_layer* layers, int n_layers, float* h, float* ws, int B, int D,
                                      int H, int S_max, int pos, mmvid_stream_t stream) {
  MMVID_REQUIRE(B >= 1 && B <= 16, "1 <= B <= 16");
  MMVID_REQUIRE(D == H * 64 && D % 4 == 0 && D <= 1024, "D = 64 H <= 1024");
  MMVID_REQUIRE(pos >= 0 && pos < S_max, "pos in range");
  float* a = ws;                       // [B, D]
  float* qkv = a + (long long)B * D;   // [B, 3D]
  float* att = qkv + (long long)B * 3 * D;  // [B, D]
  float* mid = att + (long long)B * D;      // [B, 4D]
  float* part = mid + (long long)B * 4 * D; // [B, H, Z, 66]
  const int len = pos + 1;
  const int Z = len >= 1024 ? 16 : (len >= 256 ? 8 : (len >= 64 ? 4 : 1));
  for (int li = 0; li < n_layers; ++li) {
    const mmvid_decode_layer& L = layers[li];
    int rc;
    if ((rc = mmvid_layernorm(h, D, L.ln1_w, L.ln1_b, a, MMVID_DT_F32, B, D, 1e-5f, stream))) return rc;
    if ((rc = mmvid_linear_small_m(a, D, L.in_w, D, L.in_b, nullptr, 0, qkv, 3 * D, B, 3 * D, D, MMVID_ACT_NONE, stream))) return rc;
    if ((rc = mmvid_kv_append(qkv, 3 * D, L.kcache, L.vcache, B, H, S_max, pos, stream))) return rc;
    cudaStream_t st = to_stream(stream);
    decode_attention_split_kernel<<<dim3(H, B, Z), 256, 0, st>>>(qkv, 3 * D, L.kcache, L.vcache, part, H, S_max, len, Z);
    if ((rc = check_launch("decode_attention_split"))) return rc;
    decode_attention_combine_kernel<<<dim3(H, B), 64, 0, st>>>(part, att, D, H, Z);
    if ((rc = check_launch("decode_attention_combine"))) return rc;
    if ((rc = mmvid_linear_small_m(att, D, L.out_w, D, L.out_b, h, D, h, D, B, D, D, MMVID_ACT_NONE, stream))) return rc;
    if ((rc = mmvid_layernorm(h, D, L.ln2_w, L.ln2_b, a, MMVID_DT_F32, B, D, 1e-5f, stream))) return rc;
    if ((rc = mmvid_linear_small_m(a, D, L.fc_w, D, L.fc_b, nullptr, 0, mid, 4 * D, B, 4 * D, D, MMVID_ACT_QUICKGELU, stream))) return rc;
    if ((rc = mmvid_linear_small_m(mid, 4 * D, L.proj_w, 4 * D, L.proj_b, h, D, h, D, B, D, 4 * D, MMVID_ACT_NONE, stream))) return rc;
  }
  return MMVID_OK;
}

// ------------------------------------------------------------------------------------------------
// Persistent ART-V decode kernel: ONE cooperative launch runs all transformer layers (and the image-token head)
// for the B newly sampled tokens.  Every SM streams its share of each weight matrix (the step is HBM-bound on the
// 340 MB of fp32 weights + the K/V cache); phases are separated by grid-wide barriers instead of kernel launches:
//   per layer:  [LN1 + QKV GEMV] | [cache append + single-query attention] | [out-proj + residual] |
//               [LN2 + c_fc GEMV + QuickGELU] | [c_proj + residual]            then [LN + image-logit GEMV]
// GEMV phases: the (LayerNorm-ed) B x K input rows are staged in shared memory by every block, each warp owns output
// columns n = warp_global, warp_global + total_warps, ... and streams W[n, :] with 128-bit loads.
// ------------------------------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int DEC_MAX_LAYERS = 24;
constexpr int DEC_MAX_B = 8;

struct DecodeParams {
  mmvid_decode_layer layers[DEC_MAX_LAYERS];
  int n_layers;
  float* h;        // [B, D] in/out
  float* qkv;      // [B, 3D]
  float* att;      // [B, D]
  float* mid;      // [B, 4D]
  float* part;     // [B, H, Z, 66] split-KV partials (m, l, o[64])
  const float *head_ln_w, *head_ln_b, *head_w, *head_b;  // head_w [n_logits, D] (image rows only), may be null
  float* logits;   // [B, n_logits]
  int n_logits;
  int B, D, H, S_max, pos;
};

// stage `rows` = B x K (optionally LayerNorm-ed with gamma/beta) into shared memory.  All rows are fetched in ONE pass
// (every load in flight at once); LayerNorm statistics are then computed from shared memory, one warp per row.
__device__ void stage_rows(float* sm, const float* __restrict__ src, int B, int K, const float* gamma, const float* beta,
                           float* red) {
  (void)red;
  const int n4 = (B * K) >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x)
    reinterpret_cast<float4*>(sm)[i] = reinterpret_cast<const float4*>(src)[i];
  __syncthreads();
  if (gamma != nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int b = warp; b < B; b += nw) {
      float* x = sm + b * K;
      float s = 0.f;
      for (int c = lane; c < K; c += 32) s += x[c];
      const float mean = warp_sum(s) / (float)K;
      float q = 0.f;
      for (int c = lane; c < K; c += 32) { const float d = x[c] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / (float)K + 1e-5f);
      for (int c = lane; c < K; c += 32) x[c] = (x[c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
    __syncthreads();
  }
}

// out[b, n] = act(rows[b, :] . W[n, :] + bias[n]) (+ residual[b, n]) for all N columns, spread over the whole grid.
// KS consecutive warps of a block share one column (K split KS ways, partial sums combined through shared memory in a
// fixed order: deterministic) so that narrow layers (N = D) still occupy every SM.
template <int MAXB>
__device__ void gemv_phase(const float* sm_rows, int B, int K, const float* __restrict__ W, const float* __restrict__ bias,
                           const float* residual, long long ldr, float* out, long long ldo, int N, int act,
                           float (*sm_part)[MAXB]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_warps = gridDim.x * 8;
  const int units = (K + 127) >> 7;
  int KS = 1;
  while (KS < 8 && N * (KS * 2) <= total_warps && units >= KS * 2) KS *= 2;
  const int cols_per_block = 8 / KS;
  const int upw = (units + KS - 1) / KS;
  const int sub = warp % KS, cw = warp / KS;
  for (int c0 = blockIdx.x * cols_per_block; c0 < N; c0 += gridDim.x * cols_per_block) {
    const int n = c0 + cw;
    float acc[MAXB];
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
    if (n < N) {
      gemv_slice<MAXB>(reinterpret_cast<const float4*>(W + (long long)n * K), K >> 2, sub * upw, min(units, (sub + 1) * upw),
                       lane, B,
                       [&](int b, int k4) { return *reinterpret_cast<const float4*>(sm_rows + b * K + k4 * 4); }, acc);
    }
    float mine = 0.f;  // lane b keeps row b's total
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      const float v = warp_sum(acc[b]);
      if (lane == b) mine = v;
    }
    if (KS > 1) {
      if (lane < MAXB) sm_part[warp][lane] = mine;
      __syncthreads();
      if (sub == 0 && lane < B) {
        mine = 0.f;
        for (int s2 = 0; s2 < KS; ++s2) mine += sm_part[warp + s2][lane];
      }
    }
    if (sub == 0 && lane < B && n < N) {
      float v = mine + (bias ? bias[n] : 0.f);
      v = apply_act(v, act);
      if (residual) v += residual[lane * ldr + n];
      out[lane * ldo + n] = v;
    }
    if (KS > 1) __syncthreads();
  }
}

__global__ void __launch_bounds__(256) artv_decode_persistent_kernel(DecodeParams p) {
  extern __shared__ float sm[];      // max(B * 4D staged rows, attention scratch)
  __shared__ float red[32];
  __shared__ float sm_m[8], sm_l[8];
  __shared__ float sm_o[8][64];
  __shared__ float sm_part[8][DEC_MAX_B];
  cg::grid_group grid = cg::this_grid();
  const int B = p.B, D = p.D, H = p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = p.pos + 1;
  for (int li = 0; li < p.n_layers; ++li) {
    const mmvid_decode_layer& L = p.layers[li];
    // ---- phase 1: qkv = LN1(h) W_in^T + b
    stage_rows(sm, p.h, B, D, L.ln1_w, L.ln1_b, red);
    gemv_phase<DEC_MAX_B>(sm, B, D, L.in_w, L.in_b, nullptr, 0, p.qkv, 3 * D, 3 * D, MMVID_ACT_NONE, sm_part);
    grid.sync();
    // ---- phase 2: append K,V of the new token; split-KV single-query attention: item = (b, h, z) owns keys
    //      z*8 + warp, + 8Z, ... so that ALL blocks stream the cache (a (b,h)-per-block mapping leaves 2/3 of the SMs
    //      idle and is latency-bound on dependent K -> softmax -> V loads); two keys are in flight per warp.
    const int Z = max(1, (int)gridDim.x / (B * H));
    for (int item = blockIdx.x; item < B * H * Z; item += gridDim.x) {
      const int z = item % Z, bh = item / Z;
      const int b = bh / H, hh = bh - b * H;
      const float* row = p.qkv + (long long)b * 3 * D;
      float* kb = L.kcache + ((long long)b * H + hh) * p.S_max * 64;
      float* vb = L.vcache + ((long long)b * H + hh) * p.S_max * 64;
      const float2 qv = *reinterpret_cast<const float2*>(row + hh * 64 + lane * 2);
      const float2 knew = *reinterpret_cast<const float2*>(row + D + hh * 64 + lane * 2);
      const float2 vnew = *reinterpret_cast<const float2*>(row + 2 * D + hh * 64 + lane * 2);
      if (z == 0 && warp == 0) {  // the new token's K/V row (read back from registers below, so no ordering hazard)
        *reinterpret_cast<float2*>(kb + (long long)p.pos * 64 + lane * 2) = knew;
        *reinterpret_cast<float2*>(vb + (long long)p.pos * 64 + lane * 2) = vnew;
      }
      float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
      const int step = 8 * Z;
      for (int s0 = z * 8 + warp; s0 < len; s0 += 4 * step) {
        float2 kk[4], vv[4];
        float dd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // 8 independent 256-byte row loads in flight per warp
          const int su = s0 + u * step;
          kk[u] = make_float2(0.f, 0.f); vv[u] = make_float2(0.f, 0.f);
          if (su < len) {
            kk[u] = (su == p.pos) ? knew : *reinterpret_cast<const float2*>(kb + (long long)su * 64 + lane * 2);
            vv[u] = (su == p.pos) ? vnew : *reinterpret_cast<const float2*>(vb + (long long)su * 64 + lane * 2);
          }
        }
        float mx = m;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dd[u] = (s0 + u * step < len) ? warp_sum(qv.x * kk[u].x + qv.y * kk[u].y) * 0.125f : -INFINITY;
          mx = fmaxf(mx, dd[u]);
        }
        const float alpha = expf(m - mx);
        l *= alpha; o0 *= alpha; o1 *= alpha;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float pu = (dd[u] == -INFINITY) ? 0.f : expf(dd[u] - mx);
          l += pu; o0 += pu * vv[u].x; o1 += pu * vv[u].y;
        }
        m = mx;
      }
      if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
      sm_o[warp][lane * 2] = o0; sm_o[warp][lane * 2 + 1] = o1;
      __syncthreads();
      if (threadIdx.x < 64) {
        float M = -INFINITY;
        for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w]);
        float Ls = 0.f, O = 0.f;
        for (int w = 0; w < 8; ++w) {
          const float sc = (sm_m[w] == -INFINITY) ? 0.f : expf(sm_m[w] - M);
          Ls += sm_l[w] * sc;
          O += sm_o[w][threadIdx.x] * sc;
        }
        float* dst = p.part + (long long)item * 66;
        dst[2 + threadIdx.x] = O;
        if (threadIdx.x == 0) { dst[0] = M; dst[1] = Ls; }
      }
      __syncthreads();
    }
    grid.sync();
    // ---- phase 3: h += att W_out^T + b   (att is combined from the split-KV partials while it is staged)
    for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
      const int b = i / D, c = i - b * D, hh = c >> 6, d = c & 63;
      const float* pp = p.part + ((long long)(b * H + hh) * Z) * 66;
      float M = -INFINITY;
      for (int z = 0; z < Z; ++z) M = fmaxf(M, pp[z * 66]);
      float Ls = 0.f, O = 0.f;
      for (int z = 0; z < Z; ++z) {
        const float sc = (pp[z * 66] == -INFINITY) ? 0.f : expf(pp[z * 66] - M);
        Ls += pp[z * 66 + 1] * sc;
        O += pp[z * 66 + 2 + d] * sc;
      }
      sm[i] = O / Ls;
    }
    __syncthreads();
    gemv_phase<DEC_MAX_B>(sm, B, D, L.out_w, L.out_b, p.h, D, p.h, D, D, MMVID_ACT_NONE, sm_part);
    grid.sync();
    // ---- phase 4: mid = QuickGELU(LN2(h) W_fc^T + b)
    stage_rows(sm, p.h, B, D, L.ln2_w, L.ln2_b, red);
    gemv_phase<DEC_MAX_B>(sm, B, D, L.fc_w, L.fc_b, nullptr, 0, p.mid, 4 * D, 4 * D, MMVID_ACT_QUICKGELU, sm_part);
    grid.sync();
    // ---- phase 5: h += mid W_proj^T + b
    stage_rows(sm, p.mid, B, 4 * D, nullptr, nullptr, red);
    gemv_phase<DEC_MAX_B>(sm, B, 4 * D, L.proj_w, L.proj_b, p.h, D, p.h, D, D, MMVID_ACT_NONE, sm_part);
    grid.sync();
  }
  if (p.head_w != nullptr) {
    stage_rows(sm, p.h, B, D, p.head_ln_w, p.head_ln_b, red);
    gemv_phase<DEC_MAX_B>(sm, B, D, p.head_w, p.head_b, nullptr, 0, p.logits, p.n_logits, p.n_logits, MMVID_ACT_NONE, sm_part);
  }
}

}  // namespace

// Persistent variant of mmvid_artv_decode_step (+ fused LN + image-logit head).  B <= 8.
extern "C" int mmvid_artv_decode_persistent(const mmvid_decode_layer* layers, int n_layers, float* h, float* ws,
                                            const float* head_ln_w, const float* head_ln_b, const float* head_w,
                                            const float* head_b, float* logits, int n_logits, int B, int D, int H,
                                            int S_max, int pos, mmvid_stream_t stream) {
  MMVID_REQUIRE(B >= 1 && B <= DEC_MAX_B, "1 <= B <= 8");
  MMVID_REQUIRE(n_layers >= 1 && n_layers <= DEC_MAX_LAYERS, "1..24 layers");
  MMVID_REQUIRE(D == H * 64 && D % 4 == 0, "D = 64 H");
  MMVID_REQUIRE(pos >= 0 && pos < S_max, "pos in range");
  DecodeParams p{};
  for (int i = 0; i < n_layers; ++i) p.layers[i] = layers[i];
  p.n_layers = n_layers;
  p.h = h;
  p.qkv = ws;
  p.att = p.qkv + (long long)B * 3 * D;
  p.mid = p.att + (long long)B * D;
  p.part = p.mid + (long long)B * 4 * D;
  p.head_ln_w = head_ln_w; p.head_ln_b = head_ln_b; p.head_w = head_w; p.head_b = head_b;
  p.logits = logits; p.n_logits = n_logits;
  p.B = B; p.D = D; p.H = H; p.S_max = S_max; p.pos = pos;
  const size_t smem = (size_t)B * 4 * D * sizeof(float);
  static int grid_blocks = 0;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t err = cudaFuncSetAttribute(artv_decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(artv_decode): %s", cudaGetErrorString(err));
    smem_set = smem;
    grid_blocks = 0;
  }
  if (grid_blocks == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, artv_decode_persistent_kernel, 256, smem);
    if (per_sm < 1) return fail(MMVID_ECUDA, "artv_decode: kernel does not fit on an SM%s");
    grid_blocks = sms * (per_sm > 2 ? 2 : per_sm);
  }
  void* args[] = {&p};
  cudaError_t err = cudaLaunchCooperativeKernel((void*)artv_decode_persistent_kernel, dim3(grid_blocks), dim3(256), args, smem,
                                                to_stream(stream));
  count_launch();
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "artv_decode_persistent launch: %s", cudaGetErrorString(err));
  return MMVID_OK;
}
