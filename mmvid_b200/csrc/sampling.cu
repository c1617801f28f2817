// Device-resident mask-predict sampling (BERT.mask_predict, dalle_bert.py:527-538, 646-691): the two random draws of a
// mask-predict iteration as ONE kernel each, with no host synchronisation and no intermediate [rows, 1024] tensors.
//
//   mmvid_mp_sample : per target token  probs = softmax(logits + T * gumbel);  tok ~ Categorical(probs);  Y = probs[tok]
//                     (sample_multinomial, :527-534) - softmax, inverse-CDF draw and gather fused, one warp per row;
//                     rows flagged in `skip` (tokens kept from the previous iteration, :682-684) are left untouched.
//   mmvid_mp_keep   : per sample  keep k of the not-preserved tokens WITHOUT replacement with probability proportional to
//                     their own probability Y (torch.multinomial(Y, k, replacement=False), :651) and build the next
//                     iteration's input ids (kept token or [MASK]).  Sampling without replacement from weights w is the
//                     Plackett-Luce draw, which equals taking the k largest of  log w_i + Gumbel_i  (Gumbel-top-k), so the
//                     draw becomes one radix select per sample instead of a sort + multinomial + scatter chain.
//
// These kernels serve sampling_mode = 'batched' (throughput): they follow the reference's DISTRIBUTIONS, not torch's RNG
// stream.  sampling_mode = 'reference' keeps torch.rand_like / torch.multinomial in the reference's call order (bit-exact
// ids against the reference under the same seed).  Random numbers: Philox4x32-10 (curand device API), keyed by
// (seed, row | sample, offset); a launch consumes one offset unit.
#include "common.cuh"

#include <curand_kernel.h>

using namespace mmvid;

namespace {

__device__ __forceinline__ float gumbel_from_uniform(float u) {
  // dalle_bert.py:536-538: -log(-log(U + eps) + eps), eps = 1e-20
  return -logf(-logf(u + 1e-20f) + 1e-20f);
}

// ------------------------------------------------------------------------------------------------ token draw
// One warp per row of n <= 1024 logits (n % 128 == 0: 4 consecutive values per lane per 128-column block).
template <int MAXV>  // MAXV = n / 32 values per lane
__global__ void __launch_bounds__(256) mp_sample_kernel(const float* __restrict__ logits, long long rows, int n,
                                                        float noise_scale, const uint8_t* __restrict__ skip,
                                                        float* __restrict__ Y, long long* __restrict__ tok,
                                                        unsigned long long seed, unsigned long long offset,
                                                        const int* __restrict__ step_dev) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (step_dev != nullptr) offset += (unsigned long long)__ldg(step_dev);  // CUDA-graph replay: a fresh draw per replay
  if (skip != nullptr && skip[row]) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(logits + row * n);
  float v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV / 4; ++i) {
    const float4 q = __ldg(src + i * 32 + lane);  // columns i*128 + lane*4 .. +3
    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
  }
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)row * 32ull + lane, offset, &st);
  if (noise_scale != 0.f) {
#pragma unroll
    for (int i = 0; i < MAXV / 4; ++i) {
      const float4 u = curand_uniform4(&st);
      // curand_uniform is in (0, 1]; the reference draws torch.rand in [0, 1): mirror so that the eps terms act alike
      v[4 * i] += noise_scale * gumbel_from_uniform(1.f - u.x);
      v[4 * i + 1] += noise_scale * gumbel_from_uniform(1.f - u.y);
      v[4 * i + 2] += noise_scale * gumbel_from_uniform(1.f - u.z);
      v[4 * i + 3] += noise_scale * gumbel_from_uniform(1.f - u.w);
    }
  }
  float m = v[0];
#pragma unroll
  for (int i = 1; i < MAXV; ++i) m = fmaxf(m, v[i]);
  m = warp_max(m);
  float loc = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { v[i] = __expf(v[i] - m); loc += v[i]; }
  // inclusive scan of the per-lane sums (lane order; within a lane the order is the lane's own value order)
  float inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  const float total = __shfl_sync(0xffffffffu, inc, 31);
  // ONE uniform per row (lane 0's stream), scaled to the un-normalised mass
  float u = 0.f;
  if (lane == 0) u = (1.f - curand_uniform(&st)) * total;  // [0, total)
  u = __shfl_sync(0xffffffffu, u, 0);
  const float exc = inc - loc;
  // the owning lane is the first whose inclusive sum exceeds u; rounding can leave u >= total: then the last lane that
  // holds any mass takes it
  const unsigned ballot = __ballot_sync(0xffffffffu, inc > u && loc > 0.f);
  const unsigned has = __ballot_sync(0xffffffffu, loc > 0.f);
  const int owner = ballot ? (__ffs(ballot) - 1) : (has ? 31 - __clz(has) : 0);
  if (lane == owner) {
    float c = exc, pv = 0.f;
    int pick = -1;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      c += v[i];
      if (pick < 0 && c > u && v[i] > 0.f) { pick = i; pv = v[i]; }
    }
    if (pick < 0) {  // rounding leftovers: the lane's last value with mass
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        if (v[i] > 0.f) { pick = i; pv = v[i]; }
      if (pick < 0) pick = 0;
    }
    const int col = (pick >> 2) * 128 + lane * 4 + (pick & 3);
    Y[row] = pv / total;
    tok[row] = col;
  }
}

// ------------------------------------------------------------------------------------------------ keep-mask draw
__device__ __forceinline__ unsigned int ordered_key(float f) {
  // monotone float -> uint map (larger float = larger uint); -inf maps to the smallest key among finite inputs' range
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// One CTA per (sample, beam).  keys[i] = log Y[i] + Gumbel_i for the selectable tokens, -inf otherwise; the k largest keys
// are kept (radix select over the ordered 32-bit keys: 4 passes of 8 bits), preserved tokens are always kept.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) mp_keep_kernel(const float* __restrict__ Y, const uint8_t* __restrict__ pmask,
                                                          const long long* __restrict__ I_tok, int beams, int Ttot, int k,
                                                          long long mask_id, uint8_t* __restrict__ keep,
                                                          long long* __restrict__ ids_in, unsigned long long seed,
                                                          unsigned long long offset) {
  extern __shared__ unsigned int sm_keys[];  // [Ttot]
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_need;
  const int item = blockIdx.x, sample = item / beams;
  const float* y = Y + (long long)sample * Ttot;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)item * THREADS + threadIdx.x, offset, &st);
  for (int i = threadIdx.x; i < Ttot; i += THREADS) {
    const float u = 1.f - curand_uniform(&st);  // [0, 1)
    float key = -INFINITY;
    if (!(pmask != nullptr && pmask[i]) && y[i] > 0.f) key = __logf(y[i]) + gumbel_from_uniform(u);
    sm_keys[i] = (key == -INFINITY) ? 0u : ordered_key(key);
  }
  if (threadIdx.x == 0) { s_prefix = 0u; s_need = (unsigned)k; }
  __syncthreads();
  // radix select: find the key value T such that exactly `need` selectable keys are >= T (ties at T resolved by index)
  unsigned int prefix = 0u, need = (unsigned)k;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += THREADS) hist[i] = 0u;
    __syncthreads();
    const unsigned int hi_mask = pass == 3 ? 0u : (0xffffffffu << ((pass + 1) * 8));
    for (int i = threadIdx.x; i < Ttot; i += THREADS) {
      const unsigned int kk = sm_keys[i];
      if (kk != 0u && (kk & hi_mask) == (prefix & hi_mask)) atomicAdd(&hist[(kk >> (pass * 8)) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int acc = 0u, nd = s_need;
      int b = 255;
      for (; b > 0; --b) {
        if (acc + hist[b] >= nd) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((unsigned)b << (pass * 8));
      s_need = nd - acc;  // how many keys equal to the (partial) threshold digit are still wanted
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    __syncthreads();
  }
  // keys > prefix are kept; of the keys == prefix the first `need` (by index) are kept.  Fewer than k selectable keys:
  // the threshold collapses to the smallest key and everything selectable is kept.
  // (serial tie resolution by thread 0 is fine: ties of continuous keys are practically singletons)
  __shared__ unsigned int tie_left;
  if (threadIdx.x == 0) tie_left = need;
  __syncthreads();
  uint8_t* kp = keep + (long long)item * Ttot;
  long long* ido = ids_in + (long long)item * Ttot;
  const long long* itk = I_tok + (long long)sample * Ttot;
  for (int i0 = 0; i0 < Ttot; i0 += THREADS) {
    const int i = i0 + threadIdx.x;
    bool kflag = false;
    if (i < Ttot) {
      const unsigned int kk = sm_keys[i];
      if (pmask != nullptr && pmask[i]) kflag = true;
      else if (kk != 0u && kk > prefix) kflag = true;
      else if (kk != 0u && kk == prefix) kflag = atomicAdd(&tie_left, 0xffffffffu) - 1u < 0x7fffffffu;  // old value >= 1
    }
    if (i < Ttot) {
      kp[i] = kflag ? 1 : 0;
      ido[i] = kflag ? itk[i] : mask_id;
    }
  }
}

}  // namespace

extern "C" int mmvid_mp_sample(const float* logits, long long rows, int n, float noise_scale, const uint8_t* skip,
                               float* Y, int64_t* tok, unsigned long long seed, unsigned long long offset,
                               const int* step_dev, mmvid_stream_t stream) {
  MMVID_REQUIRE(n % 128 == 0 && n >= 128 && n <= 1024, "n: multiple of 128, <= 1024");
  if (rows == 0) return MMVID_OK;
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div<long long>(rows, wpb));
  cudaStream_t st = to_stream(stream);
  long long* t = reinterpret_cast<long long*>(tok);
  switch (n / 32) {
    case 32: mp_sample_kernel<32><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    case 28: mp_sample_kernel<28><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    case 24: mp_sample_kernel<24><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    case 20: mp_sample_kernel<20><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    case 16: mp_sample_kernel<16><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    case 12: mp_sample_kernel<12><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    case 8: mp_sample_kernel<8><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
    default: mp_sample_kernel<4><<<grid, wpb * 32, 0, st>>>(logits, rows, n, noise_scale, skip, Y, t, seed, offset, step_dev); break;
  }
  return check_launch("mp_sample");
}

extern "C" int mmvid_mp_keep(const float* Y, const uint8_t* pmask, const int64_t* I_tok, int samples, int beams, int Ttot,
                             int k, long long mask_id, uint8_t* keep, int64_t* ids_in, unsigned long long seed,
                             unsigned long long offset, mmvid_stream_t stream) {
  MMVID_REQUIRE(Ttot >= 1 && Ttot <= 16384, "Ttot <= 16384");
  MMVID_REQUIRE(k >= 0 && beams >= 1, "k >= 0, beams >= 1");
  if (samples == 0) return MMVID_OK;
  constexpr int THREADS = 512;
  mp_keep_kernel<THREADS><<<samples * beams, THREADS, (size_t)Ttot * sizeof(unsigned int), to_stream(stream)>>>(
      Y, pmask, reinterpret_cast<const long long*>(I_tok), beams, Ttot, k, mask_id, keep, reinterpret_cast<long long*>(ids_in),
      seed, offset);
  return check_launch("mp_keep");
}
