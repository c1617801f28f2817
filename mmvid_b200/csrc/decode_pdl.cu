// ART-V decode step, third generation: 5 fused launches per transformer layer chained with PROGRAMMATIC DEPENDENT LAUNCH.
//
// The per-token step is HBM-bound (340 MB of fp32 weights + the K/V cache per token), but
// as ~100 tiny dependent kernels it ran at ~10 % of the HBM roofline: every kernel paid launch ramp + a cold weight
// fetch + a tail, serialised (profiles/r1_f_decode.md).  Here
//   * every kernel starts with `griddepcontrol.launch_dependents`, then issues the loads that do NOT depend on the
//     previous kernel (its slice of the weight matrix, bias; for attention the old K/V rows) and only then executes
//     `griddepcontrol.wait` - so kernel N+1 streams its weights from HBM while kernel N is still computing;
//   * LayerNorm is folded into the consumer GEMV (rows staged + normalised in shared memory by every block),
//     the K/V-cache append into the QKV GEMV's epilogue, the split-KV combine into the attention kernel (last block of
//     each (batch, head) reduces the partials), bias / QuickGELU / residual into the GEMV epilogues.
// Per layer: [LN1+QKV+append] [attention] [out-proj+residual] [LN2+c_fc+QuickGELU] [c_proj+residual]; then [LN+head].
// Summation orders are fixed (no floating-point atomics): results are deterministic.
#include "common.cuh"

using namespace mmvid;

namespace {

constexpr int PDL_MAX_B = 8;

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ptxas gives griddepcontrol.wait (ACQBULK) acquire semantics only: EARLIER loads may sink below it, and in practice it is
// hoisted to right behind launch_dependents (PREEXIT), above every load - volatile or not - which silently removes the
// overlap this file is about.  A full fence in front of the wait pins it: the prefetched weight / cache loads have been
// performed (their latency overlaps the tail of the producer kernel) before this kernel starts waiting for the producer.
__device__ __forceinline__ void pdl_wait_after_loads() {
  asm volatile("fence.acq_rel.cta;\n\tgriddepcontrol.wait;" ::: "memory");
}
// Loads that must be ISSUED before griddepcontrol.wait.  They are volatile asm on purpose: nvcc treats __ldg /
// ld.global.nc as an invariant load and sinks it to its first use - i.e. below the wait (seen in the SASS: ACQBULK
// directly after PREEXIT, every LDG.CONSTANT after it), which silently removes the overlap the whole scheme is about.
__device__ __forceinline__ float4 ldg_early_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_early_f2(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_early_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

struct FusedLinearArgs {
  const float* A; long long lda; int M, K;
  const float* ln_g; const float* ln_b;            // LayerNorm(A) over K before the product when non-null
  const float* W; long long ldw; int N;
  const float* bias; const float* residual; long long ldr;
  float* out; long long ldo; int act, KS;
  float* kcache; float* vcache; int H, S_max, pos, D;  // columns [D, 3D) are also appended to the caches when non-null
};

// out[m, n] = act(LN?(A)[m, :] . W[n, :] + bias[n]) (+ residual[m, n]); KS warps of a block split one column's K range.
template <int MAXM>
__global__ void __launch_bounds__(256) fused_rows_linear_kernel(FusedLinearArgs p) {
  extern __shared__ float sm_rows[];  // M x K staged activations
  __shared__ float sm_part[8][MAXM];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KS = p.KS, cols_per_block = 8 / KS;
  const int sub = warp % KS, cw = warp / KS;
  const int n = blockIdx.x * cols_per_block + cw;
  const bool col_ok = n < p.N;
  const int units = (p.K + 127) >> 7, upw = (units + KS - 1) / KS;
  const int u0 = sub * upw, u1 = min(units, (sub + 1) * upw);
  const int k4_end = p.K >> 2;
  const float4* w4 = reinterpret_cast<const float4*>(p.W + (long long)(col_ok ? n : 0) * p.ldw);
  // ---- independent of the producer kernel: first 8 x 512 B of this warp's weight slice, bias
  float4 w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k4 = (u0 + j) * 32 + lane;
    w[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && u0 + j < u1 && k4 < k4_end) w[j] = ldg_early_f4(w4 + k4);
  }
  const bool writer = (sub == 0) && col_ok && lane < p.M;
  float bias_v = 0.f;
  if (writer && p.bias != nullptr) bias_v = ldg_early_f1(p.bias + n);
  // ---- everything below reads what the previous kernel wrote
  pdl_wait_after_loads();
  float res_v = 0.f;
  if (writer && p.residual != nullptr) res_v = p.residual[lane * p.ldr + n];
  {
    const int row4 = p.K >> 2, n4 = p.M * row4;
    for (int i = threadIdx.x; i < n4; i += 256) {
      const int b = i / row4, k4 = i - b * row4;
      reinterpret_cast<float4*>(sm_rows)[i] = *(reinterpret_cast<const float4*>(p.A + b * p.lda) + k4);
    }
  }
  __syncthreads();
  if (p.ln_g != nullptr) {  // LayerNorm in place, one warp per row (clip_model.py:188-193, eps 1e-5)
    for (int b = warp; b < p.M; b += 8) {
      float* x = sm_rows + b * p.K;
      float s = 0.f;
      for (int c = lane; c < p.K; c += 32) s += x[c];
      const float mean = warp_sum(s) / (float)p.K;
      float q = 0.f;
      for (int c = lane; c < p.K; c += 32) { const float d = x[c] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / (float)p.K + 1e-5f);
      for (int c = lane; c < p.K; c += 32) x[c] = (x[c] - mean) * rstd * __ldg(p.ln_g + c) + __ldg(p.ln_b + c);
    }
    __syncthreads();
  }
  float acc[MAXM];
#pragma unroll
  for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
  auto fma_batch = [&](int ub) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k4 = (ub + j) * 32 + lane;
      if (ub + j < u1 && k4 < k4_end) {
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
          if (m < p.M) {
            const float4 x = *reinterpret_cast<const float4*>(sm_rows + m * p.K + k4 * 4);
            acc[m] = fmaf(x.x, w[j].x, acc[m]); acc[m] = fmaf(x.y, w[j].y, acc[m]);
            acc[m] = fmaf(x.z, w[j].z, acc[m]); acc[m] = fmaf(x.w, w[j].w, acc[m]);
          }
        }
      }
    }
  };
  if (col_ok) {
    fma_batch(u0);
    for (int ub = u0 + 8; ub < u1; ub += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k4 = (ub + j) * 32 + lane;
        w[j] = (ub + j < u1 && k4 < k4_end) ? __ldg(w4 + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      fma_batch(ub);
    }
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
    if (sub == 0 && lane < p.M) {
      mine = 0.f;
      for (int s2 = 0; s2 < KS; ++s2) mine += sm_part[warp + s2][lane];
    }
  }
  if (writer) {
    float v = apply_act(mine + bias_v, p.act);
    if (p.residual != nullptr) v += res_v;
    p.out[lane * p.ldo + n] = v;
    if (p.kcache != nullptr && n >= p.D) {  // K / V of the new token straight into the caches [B, H, S_max, 64]
      const int c = n - p.D, which = c / p.D, cc = c - which * p.D, hh = cc >> 6, d = cc & 63;
      float* dst = (which == 0 ? p.kcache : p.vcache) + (((long long)lane * p.H + hh) * p.S_max + p.pos) * 64 + d;
      *dst = v;
    }
  }
}

// Single-query attention over the cache, split-KV: block (h, b, z), 8 warps, warp w owns keys z*8 + w + i*8Z with four
// K rows and four V rows in flight; the last block of each (b, h) to finish combines the Z partials (fixed order).
__global__ void __launch_bounds__(256) decode_attention_fused_kernel(const float* __restrict__ qkv, long long q_bstride,
                                                                    const float* __restrict__ kc,
                                                                    const float* __restrict__ vc, float* __restrict__ part,
                                                                    int* __restrict__ counters, float* __restrict__ att,
                                                                    long long att_bstride, int H, int S_max, int len, int Z) {
  __shared__ float sm_m[8], sm_l[8];
  __shared__ float sm_o[8][64];
  __shared__ int sm_last;
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kb = kc + ((long long)b * H + h) * S_max * 64;
  const float* vb = vc + ((long long)b * H + h) * S_max * 64;
  const int step = 8 * Z;
  const int s_first = z * 8 + warp;
  // rows of earlier tokens do not depend on the previous kernel (only row len-1 was just appended): fetch the first
  // batch before waiting for it
  float2 kk[4], vv[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int su = s_first + u * step;
    kk[u] = make_float2(0.f, 0.f); vv[u] = make_float2(0.f, 0.f);
    if (su < len - 1) {
      kk[u] = ldg_early_f2(reinterpret_cast<const float2*>(kb + (long long)su * 64 + lane * 2));
      vv[u] = ldg_early_f2(reinterpret_cast<const float2*>(vb + (long long)su * 64 + lane * 2));
    }
  }
  pdl_wait_after_loads();
  const float2 qv = *reinterpret_cast<const float2*>(qkv + b * q_bstride + h * 64 + lane * 2);
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int s0 = s_first; s0 < len; s0 += 4 * step) {
    if (s0 != s_first) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int su = s0 + u * step;
        kk[u] = make_float2(0.f, 0.f); vv[u] = make_float2(0.f, 0.f);
        if (su < len) {
          kk[u] = *reinterpret_cast<const float2*>(kb + (long long)su * 64 + lane * 2);
          vv[u] = *reinterpret_cast<const float2*>(vb + (long long)su * 64 + lane * 2);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int su = s0 + u * step;
        if (su == len - 1) {  // the token appended by the QKV kernel: plain (coherent) loads after the wait
          kk[u] = *reinterpret_cast<const float2*>(kb + (long long)su * 64 + lane * 2);
          vv[u] = *reinterpret_cast<const float2*>(vb + (long long)su * 64 + lane * 2);
        }
      }
    }
    float dd[4];
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
  float M = -INFINITY, Ls = 0.f, O = 0.f;
  if (threadIdx.x < 64) {
    for (int w2 = 0; w2 < 8; ++w2) M = fmaxf(M, sm_m[w2]);
    for (int w2 = 0; w2 < 8; ++w2) {
      const float sc = (sm_m[w2] == -INFINITY) ? 0.f : expf(sm_m[w2] - M);
      Ls += sm_l[w2] * sc;
      O += sm_o[w2][threadIdx.x] * sc;
    }
  }
  if (Z == 1) {
    if (threadIdx.x < 64) att[b * att_bstride + h * 64 + threadIdx.x] = O / Ls;
    return;
  }
  float* dst = part + ((((long long)b * H + h) * Z + z) * 66);
  if (threadIdx.x < 64) {
    dst[2 + threadIdx.x] = O;
    if (threadIdx.x == 0) { dst[0] = M; dst[1] = Ls; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sm_last = (atomicAdd(&counters[b * H + h], 1) == Z - 1) ? 1 : 0;
  __syncthreads();
  if (sm_last == 0) return;
  __threadfence();
  if (threadIdx.x < 64) {
    const volatile float* pp = part + (((long long)b * H + h) * Z) * 66;
    float Mx = -INFINITY;
    for (int zz = 0; zz < Z; ++zz) Mx = fmaxf(Mx, pp[zz * 66]);
    float L2 = 0.f, O2 = 0.f;
    for (int zz = 0; zz < Z; ++zz) {
      const float mz = pp[zz * 66];
      const float sc = (mz == -INFINITY) ? 0.f : expf(mz - Mx);
      L2 += pp[zz * 66 + 1] * sc;
      O2 += pp[zz * 66 + 2 + threadIdx.x] * sc;
    }
    att[b * att_bstride + h * 64 + threadIdx.x] = O2 / L2;
    if (threadIdx.x == 0) counters[b * H + h] = 0;  // ready for the next layer / step
  }
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

int fused_linear(const FusedLinearArgs& in, cudaStream_t st) {
  FusedLinearArgs p = in;
  const int units = (p.K + 127) / 128;
  int KS = 1;  // enough (column, slice) warps for ~2 blocks per SM, at most one slice per 128-float unit
  while (KS < 8 && (long long)p.N * KS < 8LL * 2 * 148 && units >= KS * 2) KS *= 2;
  p.KS = KS;
  const size_t smem = (size_t)p.M * p.K * sizeof(float);
  static size_t smem_set4 = 0, smem_set8 = 0;  // static shared memory counts against the 48 KB default too: always opt in
  dim3 grid(ceil_div(p.N, 8 / KS));
  cudaError_t err;
  if (p.M <= 4) {
    if (smem > smem_set4) {
      err = cudaFuncSetAttribute(fused_rows_linear_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(fused_rows_linear<4>): %s", cudaGetErrorString(err));
      smem_set4 = smem;
    }
    err = launch_pdl(fused_rows_linear_kernel<4>, grid, dim3(256), smem, st, p);
  } else {
    if (smem > smem_set8) {
      err = cudaFuncSetAttribute(fused_rows_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(fused_rows_linear<8>): %s", cudaGetErrorString(err));
      smem_set8 = smem;
    }
    err = launch_pdl(fused_rows_linear_kernel<8>, grid, dim3(256), smem, st, p);
  }
  count_launch();
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "fused_rows_linear launch: %s", cudaGetErrorString(err));
  return MMVID_OK;
}

}  // namespace

// One ART-V decode step for the B <= 8 newly sampled tokens (h [B, D] updated in place) + final LayerNorm + image-logit
// head.  ws: mmvid_artv_decode_workspace_floats(B, D, H) floats, ZERO-INITIALISED once by the caller (it holds the
// split-KV arrival counters, which the kernels leave at zero).  Same contract as mmvid_artv_decode_persistent.
extern "C" int mmvid_artv_decode_fused(const mmvid_decode_layer* layers, int n_layers, float* h, float* ws,
                                       const float* head_ln_w, const float* head_ln_b, const float* head_w,
                                       const float* head_b, float* logits, int n_logits, int B, int D, int H, int S_max,
                                       int pos, mmvid_stream_t stream) {
  MMVID_REQUIRE(B >= 1 && B <= PDL_MAX_B, "1 <= B <= 8");
  MMVID_REQUIRE(n_layers >= 1, "n_layers >= 1");
  MMVID_REQUIRE(D == H * 64 && D % 4 == 0 && D <= 1024, "D = 64 H <= 1024");
  MMVID_REQUIRE(pos >= 0 && pos < S_max, "pos in range");
  cudaStream_t st = to_stream(stream);
  float* qkv = ws;                               // [B, 3D]
  float* att = qkv + (long long)B * 3 * D;       // [B, D]
  float* mid = att + (long long)B * D;           // [B, 4D]
  float* part = mid + (long long)B * 4 * D;      // [B, H, Z, 66]
  const long long items = (long long)B * H * 16 > 4096 ? (long long)B * H * 16 : 4096;
  int* counters = reinterpret_cast<int*>(part + items * 66);  // [B * H]
  const int len = pos + 1;
  const int Z = len >= 1024 ? 16 : (len >= 256 ? 8 : (len >= 64 ? 4 : 1));
  int rc;
  for (int li = 0; li < n_layers; ++li) {
    const mmvid_decode_layer& L = layers[li];
    FusedLinearArgs a{};
    a.A = h; a.lda = D; a.M = B; a.K = D; a.ln_g = L.ln1_w; a.ln_b = L.ln1_b;
    a.W = L.in_w; a.ldw = D; a.N = 3 * D; a.bias = L.in_b; a.out = qkv; a.ldo = 3 * D; a.act = MMVID_ACT_NONE;
    a.kcache = L.kcache; a.vcache = L.vcache; a.H = H; a.S_max = S_max; a.pos = pos; a.D = D;
    if ((rc = fused_linear(a, st))) return rc;
    cudaError_t err = launch_pdl(decode_attention_fused_kernel, dim3(H, B, Z), dim3(256), 0, st, (const float*)qkv,
                                 (long long)3 * D, (const float*)L.kcache, (const float*)L.vcache, part, counters, att,
                                 (long long)D, H, S_max, len, Z);
    count_launch();
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "decode_attention_fused launch: %s", cudaGetErrorString(err));
    FusedLinearArgs o{};
    o.A = att; o.lda = D; o.M = B; o.K = D; o.W = L.out_w; o.ldw = D; o.N = D; o.bias = L.out_b;
    o.residual = h; o.ldr = D; o.out = h; o.ldo = D; o.act = MMVID_ACT_NONE;
    if ((rc = fused_linear(o, st))) return rc;
    FusedLinearArgs f{};
    f.A = h; f.lda = D; f.M = B; f.K = D; f.ln_g = L.ln2_w; f.ln_b = L.ln2_b;
    f.W = L.fc_w; f.ldw = D; f.N = 4 * D; f.bias = L.fc_b; f.out = mid; f.ldo = 4 * D; f.act = MMVID_ACT_QUICKGELU;
    if ((rc = fused_linear(f, st))) return rc;
    FusedLinearArgs c{};
    c.A = mid; c.lda = 4 * D; c.M = B; c.K = 4 * D; c.W = L.proj_w; c.ldw = 4 * D; c.N = D; c.bias = L.proj_b;
    c.residual = h; c.ldr = D; c.out = h; c.ldo = D; c.act = MMVID_ACT_NONE;
    if ((rc = fused_linear(c, st))) return rc;
  }
  if (head_w != nullptr) {
    FusedLinearArgs hd{};
    hd.A = h; hd.lda = D; hd.M = B; hd.K = D; hd.ln_g = head_ln_w; hd.ln_b = head_ln_b;
    hd.W = head_w; hd.ldw = D; hd.N = n_logits; hd.bias = head_b; hd.out = logits; hd.ldo = n_logits; hd.act = MMVID_ACT_NONE;
    if ((rc = fused_linear(hd, st))) return rc;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MMVID_ECUDA, "artv_decode_fused: %s", cudaGetErrorString(e));
  return MMVID_OK;
}
