// Shared host/device helpers for libmmvid_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/mmvid_b200.h"

namespace mmvid {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// check the launch we just made (does not synchronise)
inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return MMVID_ECUDA;
  }
  return MMVID_OK;
}

#define MMVID_REQUIRE(cond, msg)                                              \
  do {                                                                        \
    if (!(cond)) {                                                            \
      snprintf(::mmvid::g_err, sizeof(::mmvid::g_err), "%s: requirement failed: %s (%s)", __func__, #cond, msg); \
      return MMVID_EINVAL;                                                    \
    }                                                                         \
  } while (0)

static inline cudaStream_t to_stream(mmvid_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) for the kernels of the transformer forward --------------------------------
// A kernel launched through launch_chained() may start - block scheduling, barrier / TMEM / tensor-map setup - while
// its predecessor in the stream is still draining; it must execute chain_wait() before it touches global memory (reads
// of the predecessor's results AND writes the predecessor may still read).  chain_release() lets the successor's
// blocks be scheduled as soon as every block of this grid has started.  Inside a captured CUDA graph the edge becomes a
// programmatic dependency.  MMVID_PDL=0 launches the same kernels with full stream serialisation (the device-side
// instructions are then no-ops).
__device__ __forceinline__ void chain_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool chained_launch_enabled();  // MMVID_PDL, read once (common.cu)

template <typename... KArgs, typename... Args>
cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = chained_launch_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (all threads get the result)
__device__ __forceinline__ float block_sum(float v, float* smem /* >= 32 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? smem[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? smem[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == MMVID_ACT_QUICKGELU) return v / (1.f + expf(-1.702f * v));   // x*sigmoid(1.702x), clip_model.py:196
  if (act == MMVID_ACT_SWISH) return v / (1.f + expf(-v));                // x*sigmoid(x), model.py:33
  return v;
}

// tensor-core epilogues (tf32 / bf16 modes): sigmoid through MUFU ex2 + rcp (~1e-6 relative, far inside the mode's
// own rounding) instead of the ~40-instruction accurate expf + IEEE division
// internal activation code (never crosses the ABI): QuickGELU through ONE MUFU op, x * sigmoid(1.702 x) =
// x * (0.5 + 0.5 * tanh(0.851 x)) with tanh.approx.f32 (max relative error 2^-11: the size of the 16-bit store's own
// rounding; measured: logits parity unchanged to four digits, r2o).  Halves the MUFU time of the c_fc epilogue, which bounds
// that GEMM in the 16-bit modes.  Default for 16-bit results; MMVID_GELU_TANH=0 restores ex2 + rcp.
#define MMVID_ACT_QUICKGELU_TANH 17

__device__ __forceinline__ float apply_act_fast(float v, int act) {
  if (act == MMVID_ACT_NONE) return v;
  if (act == MMVID_ACT_QUICKGELU_TANH) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * v));
    const float hv = 0.5f * v;
    return fmaf(hv, t, hv);
  }
  const float k = act == MMVID_ACT_QUICKGELU ? -1.702f * 1.4426950408889634f : -1.4426950408889634f;
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(k * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return v * r;
}

}  // namespace mmvid
