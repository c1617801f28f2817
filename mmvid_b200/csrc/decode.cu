// ART-V KV-cache decode path (K9) - a capability the reference lacks: DALLE.generate_images re-runs the
// whole prefix through all 12 layers for every sampled token (dalle_artv.py:258-281).  With a causal mask and
// additive absolute position embeddings a K/V cache is exact, so one decode step only needs
//   * skinny linears  C[M<=16, N] = act(A W^T + b) (+res)   - HBM-bound on W (read once, coalesced rows)
//   * single-query attention against the cache             - HBM-bound on K/V
// Both are CUDA-core kernels laid out for streaming bandwidth, not tensor-core shapes.
#include "common.cuh"

using namespace mmvid;

namespace {

// One warp per output column n: streams W[n, :] with 128-bit loads, keeps M accumulators.
template <int MAXM>
__global__ void __launch_bounds__(256) linear_small_m_kernel(const float* __restrict__ A, long long lda,
                                                            const float* __restrict__ W, long long ldw,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ residual, long long ldr,
                                                            float* __restrict__ C, long long ldc, int M, int N, int K,
                                                            int act) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[MAXM];
#pragma unroll
  for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
  const float4* w4 = reinterpret_cast<const float4*>(W + (long long)n * ldw);
  for (int k4 = lane; k4 < K / 4; k4 += 32) {
    const float4 w = __ldg(w4 + k4);
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < M) {
        const float4 a = *reinterpret_cast<const float4*>(A + m * lda + k4 * 4);
        acc[m] = fmaf(a.x, w.x, acc[m]); acc[m] = fmaf(a.y, w.y, acc[m]);
        acc[m] = fmaf(a.z, w.z, acc[m]); acc[m] = fmaf(a.w, w.w, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    float v = warp_sum(acc[m]);
    if (lane == 0 && m < M) {
      if (bias) v += bias[n];
      v = apply_act(v, act);
      if (residual) v += residual[m * ldr + n];
      C[m * ldc + n] = v;
    }
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
  dim3 grid(ceil_div(N, 8));
  cudaStream_t st = to_stream(stream);
  if (M <= 4) linear_small_m_kernel<4><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act);
  else if (M <= 8) linear_small_m_kernel<8><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act);
  else linear_small_m_kernel<16><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, act);
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
  return (long long)B * (D + 3 * D + D + 4 * D) + (long long)B * H * 16 * 66;
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
