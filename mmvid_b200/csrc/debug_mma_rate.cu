// Profiling hook (not on the data path): raw tcgen05.mma issue rates of the instruction shapes the attention and GEMM
// kernels use, measured with clock64() around n back-to-back MMAs + one commit on a single CTA.  Operand contents are
// whatever is in shared / tensor memory (the rate does not depend on data).
//   flavor 0: SS kind::tf32 128x128x8    (Q K^T, GEMM)       1: TS kind::tf32 128x64x8   (P V, A = P in TMEM)
//          2: SS kind::f16  128x128x16                        3: TS kind::f16  128x64x16
//          4: SS kind::tf32 128x256x8                         5: SS kind::f16  128x256x16
//          6: attention pattern tf32: 16 x TS 128x64 then 8 x SS 128x128, repeated   7: same, kind::f16 (8 + 4)
//          8: SS kind::f16 128x64x16 (Q K^T against a 64-key tile)
//   flavor 16 + k (k = 0..3): MUFU.EX2 issue rate seen by ONE warp while 1 / 4 / 8 / 12 warps (k = 0: one warp alone,
//          k = 1, 2, 3: k warps on every SM sub-partition) each run n_mma x 8 independent ex2.approx; dev_out[0] = clock64
//          cycles of warp 0, dev_out[1] = of the last warp.
//   flavor 20 + k (k = 0..3), 4 warps (one per sub-partition), per iteration: k = 0: 8 x cvt.rn.satfinite.f16x2.f32;
//          k = 1: 8 x ex2 + 4 x cvt (the softmax mix); k = 2: 8 x fma.rn.f32x2; k = 3: 8 x ex2 + 8 x fma.rn.f32x2 -
//          which of these share an execution pipe (clock64 cycles of warp 0 for n_mma iterations).
//   flavor 24 / 25: tanh.approx.f32 / rcp.approx.f32 issue rate, one warp per sub-partition (as flavor 17 for ex2).
#include "common.cuh"
#include "tc_common.cuh"

using namespace mmvid;
using namespace mmvid::tc;

namespace {

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int flavor, int n_mma, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_ptr;
  if (warp == 0 && elect_one()) {
    const uint64_t ad = make_smem_desc_sw128(smem_u32(tiles));
    const uint64_t bd = make_smem_desc_sw128(smem_u32(tiles + 32768));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t kk = (uint32_t)(i & 3) * 32;
      switch (flavor) {
        case 0: mma_ss<true>(tm, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<true>(128, 128), 1); break;
        case 1: mma_ts<true>(tm + 384, tm + (i & 15) * 8, desc_advance(bd, kk), make_idesc<true>(128, 64), 1); break;
        case 2: mma_ss<false>(tm, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<false>(128, 128), 1); break;
        case 3: mma_ts<false>(tm + 384, tm + (i & 7) * 8, desc_advance(bd, kk), make_idesc<false>(128, 64), 1); break;
        case 4: mma_ss<true>(tm, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<true>(128, 256), 1); break;
        case 5: mma_ss<false>(tm, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<false>(128, 256), 1); break;
        case 8: mma_ss<false>(tm, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<false>(128, 64), 1); break;
        case 6: {
          const int r = i % 24;
          if (r < 16) mma_ts<true>(tm + 384, tm + r * 8, desc_advance(bd, kk), make_idesc<true>(128, 64), 1);
          else mma_ss<true>(tm + 128, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<true>(128, 128), 1);
          break;
        }
        default: {
          const int r = i % 12;
          if (r < 8) mma_ts<false>(tm + 384, tm + r * 8, desc_advance(bd, kk), make_idesc<false>(128, 64), 1);
          else mma_ss<false>(tm + 128, desc_advance(ad, kk), desc_advance(bd, kk), make_idesc<false>(128, 128), 1);
          break;
        }
      }
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = (unsigned long long)(t1 - t0);  // issue time
    out[1] = (unsigned long long)(t2 - t0);  // issue + execution of all n MMAs
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int OP>  // 0: ex2.approx, 1: tanh.approx, 2: rcp.approx
__global__ void __launch_bounds__(384, 1) mufu_rate_kernel(int n, unsigned long long* out, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed * (float)(threadIdx.x + i) * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      else if (OP == 1) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      else asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += v[i];
  if (acc == 123.456f) out[2] = 1;  // keep the chains alive
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (threadIdx.x == blockDim.x - 32) out[1] = (unsigned long long)(t1 - t0);
}

__global__ void __launch_bounds__(128, 1) pipe_mix_kernel(int mode, int n, unsigned long long* out, float seed) {
  float v[8];
  unsigned long long w[8];
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = seed * (float)(threadIdx.x + i) * 1e-3f;
    w[i] = (unsigned long long)__float_as_uint(v[i]) * 0x100000001ull;
    pk[i] = 0;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    if (mode == 0 || mode == 1) {
      if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      }
#pragma unroll
      for (int i = 0; i < (mode == 0 ? 8 : 4); ++i)
        asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(v[i]), "f"(v[(i + 1) & 7]));
    } else {
      if (mode == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(w[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += v[i] + __uint_as_float(pk[i]) + __uint_as_float((uint32_t)w[i]);
  if (acc == 123.456f) out[2] = 1;
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (threadIdx.x == 96) out[1] = (unsigned long long)(t1 - t0);
}

}  // namespace

extern "C" int mmvid_debug_mma_rate(int flavor, int n_mma, unsigned long long* dev_out, mmvid_stream_t stream) {
  MMVID_REQUIRE(((flavor >= 0 && flavor <= 8) || (flavor >= 16 && flavor <= 25)) && n_mma > 0 && dev_out != nullptr,
                "flavor 0..8 or 16..25");
  if (flavor == 24 || flavor == 25) {  // tanh.approx / rcp.approx, one warp per sub-partition
    if (flavor == 24) mufu_rate_kernel<1><<<1, 128, 0, to_stream(stream)>>>(n_mma, dev_out, 0.37f);
    else mufu_rate_kernel<2><<<1, 128, 0, to_stream(stream)>>>(n_mma, dev_out, 0.37f);
    return check_launch("mufu_rate");
  }
  if (flavor >= 20) {
    pipe_mix_kernel<<<1, 128, 0, to_stream(stream)>>>(flavor - 20, n_mma, dev_out, 0.37f);
    return check_launch("pipe_mix");
  }
  if (flavor >= 16) {
    const int warps = flavor == 16 ? 1 : 4 * (flavor - 16);
    mufu_rate_kernel<0><<<1, 32 * warps, 0, to_stream(stream)>>>(n_mma, dev_out, 0.37f);
    return check_launch("mufu_rate");
  }
  static bool attr_set = false;
  const int smem = 98304 + 1024 + 256;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(mma_rate): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  mma_rate_kernel<<<1, 128, smem, to_stream(stream)>>>(flavor, n_mma, dev_out);
  return check_launch("mma_rate");
}
