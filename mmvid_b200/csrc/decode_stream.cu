// ART-V decode step, fourth generation: ONE persistent cooperative kernel per token whose weight stream is decoupled
// from the dependency chain (north_star: "a persistent KV-cache decode kernel for the per-token ARTV sample step").
//
// The step is a chain of 5 dependent phases per layer (+ the head): [LN1+QKV+cache append] [attention over the cache]
// [out-proj+residual] [LN2+c_fc+QuickGELU] [c_proj+residual].  As separate launches (decode_pdl.cu) every phase pays a
// launch ramp, a cold weight fetch and a tail: ~8-11 us each, 0.5-0.7 ms per token at B = 4, ~15 % of the HBM bound.
// Here
//   * one CTA per SM stays resident for the whole token; phases are separated by a sense-reversing grid barrier
//     (one atomic + one acquire spin per CTA, ~1.5 us) instead of kernel boundaries;
//   * every CTA owns a FIXED contiguous column slice of every weight matrix, so its share of a matrix is one contiguous
//     block of rows: a single thread streams the slabs of the NEXT phases into a ring in shared memory with bulk async
//     copies (cp.async.bulk, mbarrier complete_tx) while the current phase computes.  The HBM stream never waits for a
//     barrier: weights arrive up to three phases ahead of their use (ring of 4 slabs);
//   * weights and the K/V cache are 16-bit (fp16: the tf32 mantissa; or bf16): 170 MB + B x 18.4 KB x len per token
//      instead of twice that; accumulation, LayerNorm, softmax and the residual stream
//     stay fp32;
//   * LayerNorm, bias, QuickGELU, residual, the cache append and the split-KV combine are fused into the phases.
// Summation orders are fixed (no floating-point atomics): results are deterministic.
// The fp32 path that reproduces the reference's sampled ids bit for bit stays in decode_pdl.cu / decode.cu.
#include "common.cuh"
#include "tc_common.cuh"

#include <cooperative_groups.h>
#include <cuda_fp16.h>

using namespace mmvid;
using namespace mmvid::tc;

namespace {

constexpr int DS_THREADS = 512;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_MAX_LAYERS = 24;
constexpr int DS_MAX_B = 8;
constexpr int DS_SLOTS = 4;

struct StreamLayer {
  const float *ln1_w, *ln1_b, *in_b, *out_b, *ln2_w, *ln2_b, *fc_b, *proj_b;
  const uint16_t *in_w, *out_w, *fc_w, *proj_w;  // 16-bit copies, [N, K] row-major
  uint16_t *kcache, *vcache;                     // [B, H, S_max, 64] 16-bit
};

struct StreamParams {
  StreamLayer layers[DS_MAX_LAYERS];
  int n_layers;
  float* h;          // [B, D] in/out (fp32 residual stream of the new tokens)
  float* q;          // [B, D]
  float* att;        // [B, D]
  float* mid;        // [B, 4D]
  float* part;       // [B*H, Z, 66] split-KV partials (m, l, o[64])
  int* counters;     // [B*H] arrival counters of the split-KV combine (left at zero)
  unsigned int* bar; // [2] grid barrier: arrival count, generation
  const float *head_ln_w, *head_ln_b, *head_b;
  const uint16_t* head_w;  // [n_logits, D]
  float* logits;     // [B, n_logits]
  int n_logits;
  int B, D, H, S_max, pos, Z;
  const int* pos_dev;  // optional: device-resident step counter added to pos (CUDA-graph replay: same launch, next token)
  int f16;           // 16-bit flavour of weights / cache: 1 = fp16, 0 = bf16
  int slot_bytes;
  unsigned long long* trace;  // debug timeline (mmvid_debug_decode_trace): globaltimer stamps of CTA 0, normally null
};

// CTA 0, thread 0 stamps %globaltimer (ns) at phase events when a trace buffer is installed (scripts/decode_trace.py):
// entry i = phase_index * 8 + {0 phase start, 1 activations staged, 2 LayerNorm done, 3 slab landed, 4 items done,
// 5 results stored, 6 barrier passed}
__device__ __forceinline__ void ds_stamp(const StreamParams& p, int phase, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && phase < 64) p.trace[phase * 8 + ev] = globaltimer_ns();
}

__device__ __forceinline__ float h16_to_f32(uint32_t bits16, int f16) {
  if (f16) return __half2float(__ushort_as_half((unsigned short)bits16));
  return __uint_as_float(bits16 << 16);
}
template <bool F16>
__device__ __forceinline__ void unpack2_t(uint32_t w, float& lo, float& hi) {
  if constexpr (F16) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = v.x; hi = v.y;
  } else {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
  }
}
__device__ __forceinline__ void unpack2_h16(uint32_t w, int f16, float& lo, float& hi) {
  if (f16) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = v.x; hi = v.y;
  } else {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
  }
}
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Grid barrier (all CTAs are co-resident: one per SM).  ONE monotonic counter per launch: the k-th barrier is passed when
// the counter reaches k * gridDim.x; per CTA one release-add and an acquire spin.  Every thread's global writes are ordered
// before its CTA's arrival by __syncthreads + the arriving thread's release; the spinning thread's acquire + the closing
// __syncthreads order the other CTAs' writes before this CTA's reads.  The kernel's exit path resets the counter.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int k) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int target = k * gridDim.x;
    unsigned int v;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(bar) : "memory");
    if (v + 1 < target) {
      const uint64_t t0 = globaltimer_ns();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (v < target && globaltimer_ns() - t0 > MBAR_TIMEOUT_NS) asm volatile("trap;");
      } while (v < target);
    } else {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  __syncthreads();
}

// columns [lo, hi) of an N-column matrix owned by this CTA
__device__ __forceinline__ void col_range(int N, int& lo, int& hi) {
  const int G = gridDim.x, b = blockIdx.x;
  const int base = N / G, extra = N - base * G;
  lo = b * base + min(b, extra);
  hi = lo + base + (b < extra ? 1 : 0);
}

struct PhaseW { const uint16_t* W; int N, K; };

// weight phase `w` of the token: 4 per layer (QKV, out-proj, c_fc, c_proj) then the head
__device__ __forceinline__ PhaseW phase_weights(const StreamParams& p, int w) {
  PhaseW r;
  const int li = w >> 2, k = w & 3;
  if (li >= p.n_layers) { r.W = p.head_w; r.N = p.n_logits; r.K = p.D; return r; }
  const StreamLayer& L = p.layers[li];
  if (k == 0) { r.W = L.in_w; r.N = 3 * p.D; r.K = p.D; }
  else if (k == 1) { r.W = L.out_w; r.N = p.D; r.K = p.D; }
  else if (k == 2) { r.W = L.fc_w; r.N = 4 * p.D; r.K = p.D; }
  else { r.W = L.proj_w; r.N = p.D; r.K = 4 * p.D; }
  return r;
}

// thread 0: start streaming this CTA's slab of weight phase w into ring slot w % DS_SLOTS
__device__ __forceinline__ void issue_slab(const StreamParams& p, int w, int n_wphases, uint8_t* ring, uint64_t* full) {
  if (w >= n_wphases) return;
  const PhaseW ph = phase_weights(p, w);
  if (ph.W == nullptr) return;
  int lo, hi;
  col_range(ph.N, lo, hi);
  const int slot = w % DS_SLOTS;
  const uint32_t row_bytes = (uint32_t)ph.K * 2u;
  const uint32_t bytes = (uint32_t)(hi - lo) * row_bytes;
  if (bytes == 0) { mbar_arrive(&full[slot]); return; }
  mbar_expect_tx(&full[slot], bytes);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(ph.W) + (size_t)lo * row_bytes;
  const uint32_t dst = smem_u32(ring + (size_t)slot * p.slot_bytes);
  // rows are contiguous in memory: a few large copies (<= 16 KB each keeps several in flight)
  for (uint32_t off = 0; off < bytes; off += 16384u) bulk_g2s(dst + off, src + off, min(16384u, bytes - off), &full[slot]);
}

// One GEMV phase on this CTA's column slice [lo, hi):
//   out[b, n] = act(LN?(A)[b, :] . W[n, :] + bias[n]) (+ residual[b, n])
// The B activation rows are staged (LayerNorm-ed) in shared memory; items (column, K slice) are dealt round robin to the
// 16 warps, each item loads its activation slice into registers, streams its weight slice from the ring slab and leaves
// one partial per row in shared memory; the K slices of a column are then summed in a fixed order.
template <int MAXB, bool F16>
__device__ void gemv_phase(const StreamParams& p, const uint8_t* slab, const float* __restrict__ A, long long lda, int K,
                           const float* ln_g, const float* ln_b, int N, const float* __restrict__ bias, int act,
                           const float* residual, float* out, long long ldo, bool qkv_mode, const StreamLayer* L, int pos,
                           float* sm_act, float* sm_part, int tphase, const uint64_t* slab_bar, uint32_t slab_par) {
  const int B = p.B, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ds_stamp(p, tphase, 0);
  int lo, hi;
  col_range(N, lo, hi);
  const int C = hi - lo;
  // ---- everything the epilogue needs from global memory is requested NOW (one L2 round trip hidden behind the phase)
  float bias_v = 0.f, res_v = 0.f;
  if ((int)threadIdx.x < C * B) {
    const int c = threadIdx.x / B, b = threadIdx.x - c * B;
    if (bias) bias_v = __ldg(bias + lo + c);
    if (residual) res_v = __ldcg(residual + b * ldo + lo + c);
  }
  // ---- stage A [B, K] (fp32, written by other CTAs before the barrier: L2 loads) into shared memory
  if (ln_g != nullptr && K <= 1024) {
    // LayerNorm (clip_model.py:188-193, eps 1e-5) while staging: one warp per row, the row lives in registers (two-pass
    // statistics without touching memory twice); gamma / beta are requested before the row arrives
    for (int b = warp; b < B; b += DS_WARPS) {
      float4 x[8], gm[8], bt[8];
      const int n4 = K >> 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k4 = lane + 32 * i;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        gm[i] = x[i];
        bt[i] = x[i];
        if (k4 < n4) {
          x[i] = ldcg_f4(A + b * lda + k4 * 4);
          gm[i] = __ldg(reinterpret_cast<const float4*>(ln_g) + k4);
          bt[i] = __ldg(reinterpret_cast<const float4*>(ln_b) + k4);
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += (x[i].x + x[i].y) + (x[i].z + x[i].w);
      const float mean = warp_sum(sum) / (float)K;
      float qq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (lane + 32 * i < n4) {
          const float a = x[i].x - mean, c2 = x[i].y - mean, d = x[i].z - mean, e2 = x[i].w - mean;
          qq += (a * a + c2 * c2) + (d * d + e2 * e2);
        }
      }
      const float rstd = rsqrtf(warp_sum(qq) / (float)K + 1e-5f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k4 = lane + 32 * i;
        if (k4 < n4) {
          float4 o;
          o.x = (x[i].x - mean) * rstd * gm[i].x + bt[i].x;
          o.y = (x[i].y - mean) * rstd * gm[i].y + bt[i].y;
          o.z = (x[i].z - mean) * rstd * gm[i].z + bt[i].z;
          o.w = (x[i].w - mean) * rstd * gm[i].w + bt[i].w;
          reinterpret_cast<float4*>(sm_act + b * K)[k4] = o;
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < B * (K >> 2); i += DS_THREADS) {
      const int b = i / (K >> 2), k4 = i - b * (K >> 2);
      reinterpret_cast<float4*>(sm_act)[i] = ldcg_f4(A + b * lda + k4 * 4);
    }
    if (ln_g != nullptr) {  // rows longer than 1024: statistics from shared memory
      __syncthreads();
      for (int b = warp; b < B; b += DS_WARPS) {
        float* x = sm_act + b * K;
        float sum = 0.f;
        for (int c = lane; c < K; c += 32) sum += x[c];
        const float mean = warp_sum(sum) / (float)K;
        float qq = 0.f;
        for (int c = lane; c < K; c += 32) { const float d = x[c] - mean; qq += d * d; }
        const float rstd = rsqrtf(warp_sum(qq) / (float)K + 1e-5f);
        for (int c = lane; c < K; c += 32) x[c] = (x[c] - mean) * rstd * __ldg(ln_g + c) + __ldg(ln_b + c);
      }
    }
  }
  __syncthreads();
  ds_stamp(p, tphase, 2);
  mbar_wait(const_cast<uint64_t*>(slab_bar), slab_par);  // this phase's weight slab has landed in the ring
  ds_stamp(p, tphase, 3);
  // K slices: MAXB x slice / 32 activation registers per lane (<= 96); more slices when the CTA has few columns
  const int slice_max = MAXB <= 4 ? 768 : 384;
  int KS = max(1, K / slice_max);
  while (C * KS < DS_WARPS && KS < 8 && (K / (KS * 2)) % 128 == 0) KS *= 2;
  const int slice = K / KS, gpl = slice >> 7;  // 4-element granules per lane (slice is a multiple of 128)
  const int items = C * KS;
  int c = warp / KS, sl = warp - c * KS;  // item `warp`; the loop advances by DS_WARPS items without divisions
  const int dc = DS_WARPS / KS, ds = DS_WARPS - dc * KS;
  for (int it = warp; it < items; it += DS_WARPS) {
    const int k0 = sl * slice;
    float acc[MAXB];
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
    const uint2* wrow = reinterpret_cast<const uint2*>(slab + ((size_t)c * K + k0) * 2);
#pragma unroll 2
    for (int g = 0; g < gpl; ++g) {
      const uint2 wv = wrow[g * 32 + lane];
      float w0, w1, w2, w3;
      unpack2_t<F16>(wv.x, w0, w1);
      unpack2_t<F16>(wv.y, w2, w3);
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        if (b < B) {
          const float4 x = *reinterpret_cast<const float4*>(sm_act + b * K + k0 + (g * 32 + lane) * 4);
          acc[b] = fmaf(x.x, w0, acc[b]); acc[b] = fmaf(x.y, w1, acc[b]);
          acc[b] = fmaf(x.z, w2, acc[b]); acc[b] = fmaf(x.w, w3, acc[b]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      const float v = warp_sum(acc[b]);
      if (lane == 0 && b < B) sm_part[it * MAXB + b] = v;
    }
    c += dc; sl += ds;
    if (sl >= KS) { sl -= KS; ++c; }
  }
  __syncthreads();
  ds_stamp(p, tphase, 4);
  if ((int)threadIdx.x < C * B) {  // C * B <= DS_THREADS (checked on the host)
    const int cc0 = threadIdx.x / B, b = threadIdx.x - cc0 * B, n = lo + cc0;
    float v = 0.f;
    for (int s2 = 0; s2 < KS; ++s2) v += sm_part[(cc0 * KS + s2) * MAXB + b];
    v = apply_act(v + bias_v, act);
    if (residual) v += res_v;
    if (qkv_mode && n >= p.D) {  // K / V of the new token straight into the 16-bit caches [B, H, S_max, 64]
      const int cc = n - p.D, which = cc / p.D, c2 = cc - which * p.D, hh = c2 >> 6, d = c2 & 63;
      uint16_t* dst = (which == 0 ? L->kcache : L->vcache) + (((long long)b * p.H + hh) * p.S_max + pos) * 64 + d;
      *dst = cvt_h16_rt(v, p.f16);
    } else {
      out[b * ldo + n] = v;
    }
  }
}

// Single-query attention over the 16-bit cache, split-KV over Z CTAs per (batch, head).  8 lanes x 16 bytes cover one
// 64-dim row, so a warp load fetches 4 keys; 4 such loads (16 keys) of K and of V are in flight per warp.
template <bool F16>
__device__ void attention_phase(const StreamParams& p, const StreamLayer& L, int pos, int Z, float* sm_red) {
  const int B = p.B, H = p.H, len = pos + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sm_m = sm_red;                 // [16]
  float* sm_l = sm_red + 16;            // [16]
  float* sm_o = sm_red + 32;            // [16][64]
  __shared__ int sm_last;
  for (int item = blockIdx.x; item < B * H * Z; item += gridDim.x) {
    const int z = item % Z, bh = item / Z;
    const int b = bh / H, hh = bh - b * H;
    const uint16_t* kb = L.kcache + (long long)bh * p.S_max * 64;
    const uint16_t* vb = L.vcache + (long long)bh * p.S_max * 64;
    const int kq = lane >> 3, dq = lane & 7;
    float qv[8];
    {
      const float4 a = ldcg_f4(p.q + b * p.D + hh * 64 + dq * 8), c = ldcg_f4(p.q + b * p.D + hh * 64 + dq * 8 + 4);
      qv[0] = a.x * 0.125f; qv[1] = a.y * 0.125f; qv[2] = a.z * 0.125f; qv[3] = a.w * 0.125f;
      qv[4] = c.x * 0.125f; qv[5] = c.y * 0.125f; qv[6] = c.z * 0.125f; qv[7] = c.w * 0.125f;
    }
    // keys of this CTA: [z0, z1); of this warp: [w0, w1)
    const int z0 = (int)((long long)len * z / Z), z1 = (int)((long long)len * (z + 1) / Z);
    const int per = (z1 - z0 + DS_WARPS - 1) / DS_WARPS;
    const int w0 = z0 + warp * per, w1 = min(z1, w0 + per);
    float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    // 32 keys per warp in flight: 8 K rows and 8 V rows per lane quartet are requested before the first one is used
    for (int s0 = w0; s0 < w1; s0 += 32) {
      uint4 kk[8], vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int su = s0 + u * 4 + kq;
        kk[u] = make_uint4(0, 0, 0, 0); vv[u] = make_uint4(0, 0, 0, 0);
        if (su < w1) {
          kk[u] = __ldcg(reinterpret_cast<const uint4*>(kb + (long long)su * 64 + dq * 8));
          vv[u] = __ldcg(reinterpret_cast<const uint4*>(vb + (long long)su * 64 + dq * 8));
        }
      }
      float sc[8];
      float mx = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float k0, k1, k2, k3, k4, k5, k6, k7;
        unpack2_t<F16>(kk[u].x, k0, k1); unpack2_t<F16>(kk[u].y, k2, k3);
        unpack2_t<F16>(kk[u].z, k4, k5); unpack2_t<F16>(kk[u].w, k6, k7);
        float d = qv[0] * k0 + qv[1] * k1 + qv[2] * k2 + qv[3] * k3 + qv[4] * k4 + qv[5] * k5 + qv[6] * k6 + qv[7] * k7;
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        sc[u] = (s0 + u * 4 + kq < w1) ? d : -INFINITY;
        mx = fmaxf(mx, sc[u]);
      }
      if (mx != -INFINITY) {
        const float alpha = (m == -INFINITY) ? 0.f : expf(m - mx);
        l *= alpha;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= alpha;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float pu = (sc[u] == -INFINITY) ? 0.f : expf(sc[u] - mx);
          l += pu;
          float v0, v1, v2, v3, v4, v5, v6, v7;
          unpack2_t<F16>(vv[u].x, v0, v1); unpack2_t<F16>(vv[u].y, v2, v3);
          unpack2_t<F16>(vv[u].z, v4, v5); unpack2_t<F16>(vv[u].w, v6, v7);
          o[0] = fmaf(pu, v0, o[0]); o[1] = fmaf(pu, v1, o[1]); o[2] = fmaf(pu, v2, o[2]); o[3] = fmaf(pu, v3, o[3]);
          o[4] = fmaf(pu, v4, o[4]); o[5] = fmaf(pu, v5, o[5]); o[6] = fmaf(pu, v6, o[6]); o[7] = fmaf(pu, v7, o[7]);
        }
        m = mx;
      }
    }
    // merge the 4 key groups of the warp (lanes with equal dq): xor 8, 16
#pragma unroll
    for (int sh = 8; sh <= 16; sh <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, sh), l2 = __shfl_xor_sync(0xffffffffu, l, sh);
      const float M = fmaxf(m, m2);
      const float a1 = (m == -INFINITY) ? 0.f : expf(m - M), a2 = (m2 == -INFINITY) ? 0.f : expf(m2 - M);
      l = l * a1 + l2 * a2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float o2 = __shfl_xor_sync(0xffffffffu, o[i], sh);
        o[i] = o[i] * a1 + o2 * a2;
      }
      m = M;
    }
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm_o[warp * 64 + lane * 8 + i] = o[i];
    }
    __syncthreads();
    float M = -INFINITY, Ls = 0.f, O = 0.f;
    if (threadIdx.x < 64) {
      for (int w2 = 0; w2 < DS_WARPS; ++w2) M = fmaxf(M, sm_m[w2]);
      for (int w2 = 0; w2 < DS_WARPS; ++w2) {
        const float scl = (sm_m[w2] == -INFINITY) ? 0.f : expf(sm_m[w2] - M);
        Ls += sm_l[w2] * scl;
        O += sm_o[w2 * 64 + threadIdx.x] * scl;
      }
    }
    if (Z == 1) {
      if (threadIdx.x < 64) p.att[b * p.D + hh * 64 + threadIdx.x] = O / Ls;
      __syncthreads();
      continue;
    }
    float* dst = p.part + ((long long)bh * Z + z) * 66;
    if (threadIdx.x < 64) {
      dst[2 + threadIdx.x] = O;
      if (threadIdx.x == 0) { dst[0] = M; dst[1] = Ls; }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sm_last = (atomicAdd(&p.counters[bh], 1) == Z - 1) ? 1 : 0;
    __syncthreads();
    if (sm_last) {  // the last of the Z CTAs combines the partials in a fixed order
      __threadfence();
      if (threadIdx.x < 64) {
        const float* pp = p.part + (long long)bh * Z * 66;
        float Mx = -INFINITY;
        for (int zz = 0; zz < Z; ++zz) Mx = fmaxf(Mx, __ldcg(pp + zz * 66));
        float L2 = 0.f, O2 = 0.f;
        for (int zz = 0; zz < Z; ++zz) {
          const float mz = __ldcg(pp + zz * 66);
          const float scl = (mz == -INFINITY) ? 0.f : expf(mz - Mx);
          L2 += __ldcg(pp + zz * 66 + 1) * scl;
          O2 += __ldcg(pp + zz * 66 + 2 + threadIdx.x) * scl;
        }
        p.att[b * p.D + hh * 64 + threadIdx.x] = O2 / L2;
        if (threadIdx.x == 0) p.counters[bh] = 0;  // ready for the next layer / token
      }
    }
    __syncthreads();
  }
}

template <int MAXB, bool F16>
__global__ void __launch_bounds__(DS_THREADS, 1) artv_decode_stream_kernel(const __grid_constant__ StreamParams p) {
  // the step counter of a graph replay lives in device memory
  const int pos = p.pos + (p.pos_dev != nullptr ? __ldcg(p.pos_dev) : 0);
  // split-KV factor: at most the host's Z (CTAs per (batch, head)), fewer while the cache is short
  int Z = p.Z;
  while (Z > 1 && pos + 1 < Z * 64) --Z;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  __shared__ uint64_t full[DS_SLOTS];
  // carve: ring of DS_SLOTS weight slabs | staged activations [B, 4D] fp32 | partials / attention scratch
  uint8_t* ring = sm_raw;
  float* sm_act = reinterpret_cast<float*>(sm_raw + (size_t)DS_SLOTS * p.slot_bytes);
  float* sm_part = sm_act + (size_t)p.B * 4 * p.D;
  const int n_wphases = 4 * p.n_layers + (p.head_w != nullptr ? 1 : 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < DS_SLOTS; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
    fence_proxy_async();
    for (int w = 0; w < DS_SLOTS - 1; ++w) issue_slab(p, w, n_wphases, ring, full);
  }
  __syncthreads();
  const int D = p.D;
  int w = 0;       // weight phase counter (4 per layer + head)
  int tp = 0;      // phase counter incl. attention (trace index)
  unsigned int nbar = 0;  // grid barriers passed so far in this launch
  // the ring is primed with phases 0 .. DS_SLOTS-2; phase w + DS_SLOTS - 1 is issued when phase w starts (its slot held
  // phase w - 1, which every warp of this CTA finished before the barrier it has just passed)
  auto gemv = [&](const float* A, long long lda, int K, const float* ln_g, const float* ln_b, int N, const float* bias, int act,
                  const float* residual, float* out, long long ldo, bool qkv_mode, const StreamLayer* L) {
    if (threadIdx.x == 0) { fence_proxy_async(); issue_slab(p, w + DS_SLOTS - 1, n_wphases, ring, full); }
    gemv_phase<MAXB, F16>(p, ring + (size_t)(w % DS_SLOTS) * p.slot_bytes, A, lda, K, ln_g, ln_b, N, bias, act, residual, out, ldo,
                     qkv_mode, L, pos, sm_act, sm_part, tp, &full[w % DS_SLOTS], (uint32_t)(w / DS_SLOTS) & 1u);
    ++w;
  };
  auto barrier = [&]() {
    ds_stamp(p, tp, 5);
    grid_barrier(p.bar, ++nbar);
    ds_stamp(p, tp, 6);
    ++tp;
  };
  for (int li = 0; li < p.n_layers; ++li) {
    const StreamLayer& L = p.layers[li];
    gemv(p.h, D, D, L.ln1_w, L.ln1_b, 3 * D, L.in_b, MMVID_ACT_NONE, nullptr, p.q, D, true, &L);  // q -> p.q, k / v -> caches
    barrier();
    ds_stamp(p, tp, 0);
    attention_phase<F16>(p, L, pos, Z, sm_part);  // no weights: the ring keeps filling meanwhile
    barrier();
    gemv(p.att, D, D, nullptr, nullptr, D, L.out_b, MMVID_ACT_NONE, p.h, p.h, D, false, &L);     // out-proj + residual
    barrier();
    gemv(p.h, D, D, L.ln2_w, L.ln2_b, 4 * D, L.fc_b, MMVID_ACT_QUICKGELU, nullptr, p.mid, 4 * D, false, &L);
    barrier();
    gemv(p.mid, 4 * D, 4 * D, nullptr, nullptr, D, L.proj_b, MMVID_ACT_NONE, p.h, p.h, D, false, &L);  // c_proj + residual
    barrier();
  }
  if (p.head_w != nullptr)
    gemv(p.h, D, D, p.head_ln_w, p.head_ln_b, p.n_logits, p.head_b, MMVID_ACT_NONE, nullptr, p.logits, p.n_logits, false, nullptr);
  ds_stamp(p, tp, 5);
  // leave the barrier counter at zero for the next launch: the last CTA to get here resets it (every CTA has passed its
  // last barrier wait before it arrives here, so nobody reads the counter any more)
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(p.bar + 1, 1u) == gridDim.x - 1) {
      atomicExch(p.bar, 0u);
      atomicExch(p.bar + 1, 0u);
    }
  }
}

unsigned long long* g_decode_trace = nullptr;

size_t stream_ws_floats(int B, int D, int H, int Z) {
  // q, att [B, D] | mid [B, 4D] | part [B*H, Z, 66] | counters [B*H] | barrier [2] (+ padding)
  return (size_t)B * D * 2 + (size_t)B * 4 * D + (size_t)B * H * Z * 66 + (size_t)B * H + 64;
}

}  // namespace

// Profiling hook: CTA 0 of every following mmvid_artv_decode_stream launch writes %globaltimer stamps of its phase events
// into dev_buf (>= 512 uint64; NULL switches it off).  Layout: ds_stamp.
extern "C" int mmvid_debug_decode_trace(unsigned long long* dev_buf) {
  g_decode_trace = dev_buf;
  return MMVID_OK;
}

extern "C" long long mmvid_artv_decode_stream_workspace_floats(int B, int D, int H) {
  return (long long)stream_ws_floats(B, D, H, 16);
}

// One ART-V decode step (B <= 8 new tokens, h [B, D] updated in place, logits [B, n_logits] out) as ONE persistent launch.
// Weights (mmvid_decode_layer16) and K/V caches are 16-bit (f16 != 0: fp16, else bf16).  ws: ZERO-INITIALISED once by the
// caller (barrier words and split-KV counters; the kernel leaves them at zero / consistent).  Returns 1 ("not applicable")
// when the shape does not fit the kernel's shared-memory plan, so that callers can fall back to mmvid_artv_decode_fused.
extern "C" int mmvid_artv_decode_stream(const mmvid_decode_layer16* layers, int n_layers, float* h, float* ws,
                                        const float* head_ln_w, const float* head_ln_b, const void* head_w16,
                                        const float* head_b, float* logits, int n_logits, int B, int D, int H, int S_max,
                                        int pos, const int* pos_dev, int f16, mmvid_stream_t stream) {
  MMVID_REQUIRE(B >= 1 && B <= DS_MAX_B, "1 <= B <= 8");
  MMVID_REQUIRE(n_layers >= 1 && n_layers <= DS_MAX_LAYERS, "1..24 layers");
  MMVID_REQUIRE(D == H * 64 && D % 128 == 0, "D = 64 H, multiple of 128");
  MMVID_REQUIRE(pos >= 0 && pos < S_max, "pos in range");
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int G = sms;
  StreamParams p{};
  auto cols = [&](int N) { return (N + G - 1) / G; };
  size_t slot = 0;
  {
    const size_t cand[5] = {(size_t)cols(3 * D) * D * 2, (size_t)cols(D) * D * 2, (size_t)cols(4 * D) * D * 2,
                            (size_t)cols(D) * 4 * D * 2, head_w16 ? (size_t)cols(n_logits) * D * 2 : 0};
    for (size_t c : cand) slot = c > slot ? c : slot;
    slot = (slot + 127) & ~(size_t)127;
  }
  const int maxb = B <= 4 ? 4 : 8;
  const int max_items = 16 * 8 + 64;  // C * KS upper bound used for the partial buffer (checked below)
  const size_t part_floats = (size_t)max_items * maxb > 32 + 16 * 64 ? (size_t)max_items * maxb : 32 + 16 * 64;
  const size_t smem = DS_SLOTS * slot + (size_t)B * 4 * D * 4 + part_floats * 4;
  if (smem > 220 * 1024) return 1;
  {
    // items per phase = C * KS must fit the partial buffer
    const int Cs[5] = {cols(3 * D), cols(D), cols(4 * D), cols(D), head_w16 ? cols(n_logits) : 0};
    for (int c : Cs)
      if (c * 8 > max_items) return 1;
  }
  for (int i = 0; i < n_layers; ++i) {
    const mmvid_decode_layer16& s = layers[i];
    StreamLayer& d = p.layers[i];
    d.ln1_w = s.ln1_w; d.ln1_b = s.ln1_b; d.in_b = s.in_b; d.out_b = s.out_b; d.ln2_w = s.ln2_w; d.ln2_b = s.ln2_b;
    d.fc_b = s.fc_b; d.proj_b = s.proj_b;
    d.in_w = (const uint16_t*)s.in_w; d.out_w = (const uint16_t*)s.out_w; d.fc_w = (const uint16_t*)s.fc_w;
    d.proj_w = (const uint16_t*)s.proj_w;
    d.kcache = (uint16_t*)s.kcache; d.vcache = (uint16_t*)s.vcache;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    MMVID_REQUIRE(al16(s.in_w) && al16(s.out_w) && al16(s.fc_w) && al16(s.proj_w) && al16(s.kcache) && al16(s.vcache),
                  "16-byte aligned weights and caches");
  }
  p.n_layers = n_layers;
  int Z = G / (B * H);
  if (Z < 1) Z = 1;
  if (Z > 16) Z = 16;
  p.Z = Z;  // upper bound; the kernel uses fewer, fuller splits while the cache is short
  p.pos_dev = pos_dev;
  p.h = h;
  p.q = ws;
  p.att = p.q + (size_t)B * D;
  p.mid = p.att + (size_t)B * D;
  p.part = p.mid + (size_t)B * 4 * D;
  p.counters = reinterpret_cast<int*>(p.part + (size_t)B * H * 16 * 66);
  p.bar = reinterpret_cast<unsigned int*>(p.counters + B * H + (16 - (B * H) % 16));
  p.head_ln_w = head_ln_w; p.head_ln_b = head_ln_b; p.head_b = head_b; p.head_w = (const uint16_t*)head_w16;
  p.logits = logits; p.n_logits = n_logits;
  p.B = B; p.D = D; p.H = H; p.S_max = S_max; p.pos = pos; p.f16 = f16 ? 1 : 0; p.slot_bytes = (int)slot;
  p.trace = g_decode_trace;
  static size_t smem_set[4] = {0, 0, 0, 0};
  const int ki = (maxb == 8 ? 2 : 0) + (p.f16 ? 1 : 0);
  void* kern = ki == 0 ? (void*)artv_decode_stream_kernel<4, false>
               : ki == 1 ? (void*)artv_decode_stream_kernel<4, true>
               : ki == 2 ? (void*)artv_decode_stream_kernel<8, false> : (void*)artv_decode_stream_kernel<8, true>;
  if (smem > smem_set[ki]) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(artv_decode_stream): %s", cudaGetErrorString(err));
    smem_set[ki] = smem;
  }
  // Co-residency of the G = #SMs CTAs (1 per SM by shared memory) is what the grid barrier needs.  A cooperative launch
  // asserts it; under stream capture (the per-token CUDA graph of DALLE.generate_images) a plain launch is used: the
  // grid still fits the device exactly, and CTAs that wait for an SM held by the predecessor's tail are merely late.
  void* args[] = {&p};
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(to_stream(stream), &cap);
  cudaError_t err;
  if (cap == cudaStreamCaptureStatusNone)
    err = cudaLaunchCooperativeKernel(kern, dim3(G), dim3(DS_THREADS), args, smem, to_stream(stream));
  else
    err = cudaLaunchKernel(kern, dim3(G), dim3(DS_THREADS), args, smem, to_stream(stream));
  count_launch();
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "artv_decode_stream launch: %s", cudaGetErrorString(err));
  return MMVID_OK;
}
