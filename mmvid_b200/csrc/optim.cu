// K18: multi-tensor optimiser step for the reference's training loop (train.py:320-325, utils_train.py:167-181):
//   clip_grad_norm_(dalle.parameters(), max_norm)  ->  mmvid_grad_sqnorm + mmvid_grad_clip
//   torch.optim.Adam / AdamW .step()               ->  mmvid_adam_step
// One launch covers every parameter tensor of the model: a device-resident table of (param, grad, exp_avg, exp_avg_sq, n)
// records plus a chunk map (chunk -> tensor, chunk index) built once by the host; each block owns one chunk.
// HBM-bound: Adam moves 4 reads + 3 writes of 4 B per parameter (28 B / parameter), the norm 4 B / parameter.
#include "common.cuh"

using namespace mmvid;

namespace {

constexpr int OPT_THREADS = 256;

// deterministic: one partial per chunk (fixed in-block tree), then a single-block ordered finalize
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_partial_kernel(const mmvid_adam_tensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                           const int* __restrict__ chunk_index, int chunk_elems, float* __restrict__ partial) {
  __shared__ float red[32];
  const mmvid_adam_tensor t = table[chunk_tensor[blockIdx.x]];
  const long long lo = (long long)chunk_index[blockIdx.x] * chunk_elems;
  const long long hi = min(t.n, lo + chunk_elems);
  const float* g = reinterpret_cast<const float*>(t.g);
  float s = 0.f;
  if (g != nullptr) {
    const bool vec = ((reinterpret_cast<uintptr_t>(g) & 15) == 0) && ((lo & 3) == 0);
    if (vec) {
      const long long n4 = (hi - lo) >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(g + lo);
      for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
        const float4 v = g4[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      for (long long i = lo + (n4 << 2) + threadIdx.x; i < hi; i += OPT_THREADS) s += g[i] * g[i];
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += OPT_THREADS) s += g[i] * g[i];
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
grad_sqnorm_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ total_sq) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) total_sq[0] = s;
}

// grads *= min(1, max_norm / (sqrt(total_sq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void __launch_bounds__(OPT_THREADS)
grad_clip_kernel(const mmvid_adam_tensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                 const int* __restrict__ chunk_index, int chunk_elems, const float* __restrict__ total_sq, float max_norm) {
  const float coef = fminf(1.f, max_norm / (sqrtf(total_sq[0]) + 1e-6f));
  if (coef >= 1.f) return;
  const mmvid_adam_tensor t = table[chunk_tensor[blockIdx.x]];
  float* g = reinterpret_cast<float*>(const_cast<void*>(t.g));
  if (g == nullptr) return;
  const long long lo = (long long)chunk_index[blockIdx.x] * chunk_elems;
  const long long hi = min(t.n, lo + chunk_elems);
  for (long long i = lo + threadIdx.x; i < hi; i += OPT_THREADS) g[i] *= coef;
}

struct AdamScalars {
  float lr, beta1, beta2, eps, wd, step_size, sqrt_bc2;
  int decoupled, step;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& a) {
  if (a.wd != 0.f) {
    if (a.decoupled) p *= (1.f - a.lr * a.wd);  // AdamW
    else g = fmaf(a.wd, p, g);                  // Adam (L2 added to the gradient)
  }
  m = m + (g - m) * (1.f - a.beta1);            // torch: exp_avg.lerp_(grad, 1 - beta1)
  v = a.beta2 * v + (1.f - a.beta2) * g * g;    // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.sqrt_bc2 + a.eps;  // torch: (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - a.step_size * (m / denom);
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_step_kernel(const mmvid_adam_tensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                 const int* __restrict__ chunk_index, int chunk_elems, AdamScalars a) {
  const mmvid_adam_tensor t = table[chunk_tensor[blockIdx.x]];
  if (t.g == nullptr) return;
  if (t.skipped != 0) {
    // this tensor missed `skipped` earlier updates: torch keeps a step counter per parameter, so its bias corrections
    // differ from the launch-wide ones (block-uniform branch, double precision like the host path)
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const double st = (double)(a.step - (int)t.skipped);
      sh[0] = (float)((double)a.lr / (1.0 - pow((double)a.beta1, st)));
      sh[1] = (float)sqrt(1.0 - pow((double)a.beta2, st));
    }
    __syncthreads();
    a.step_size = sh[0];
    a.sqrt_bc2 = sh[1];
  }
  const long long lo = (long long)chunk_index[blockIdx.x] * chunk_elems;
  const long long hi = min(t.n, lo + chunk_elems);
  float* p = reinterpret_cast<float*>(t.p);
  const float* g = reinterpret_cast<const float*>(t.g);
  float* m = reinterpret_cast<float*>(t.m);
  float* v = reinterpret_cast<float*>(t.v);
  const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0) && ((lo & 3) == 0);
  long long tail = lo;
  if (vec) {
    const long long n4 = (hi - lo) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + lo);
    const float4* g4 = reinterpret_cast<const float4*>(g + lo);
    float4* m4 = reinterpret_cast<float4*>(m + lo);
    float4* v4 = reinterpret_cast<float4*>(v + lo);
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      const float4 gg = g4[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, a);
      adam_one(pp.y, gg.y, mm.y, vv.y, a);
      adam_one(pp.z, gg.z, mm.z, vv.z, a);
      adam_one(pp.w, gg.w, mm.w, vv.w, a);
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    tail = lo + (n4 << 2);
  }
  for (long long i = tail + threadIdx.x; i < hi; i += OPT_THREADS) adam_one(p[i], g[i], m[i], v[i], a);
}

}  // namespace

extern "C" {

int mmvid_grad_sqnorm(const mmvid_adam_tensor* table_dev, const int* chunk_tensor_dev, const int* chunk_index_dev,
                      int n_chunks, int chunk_elems, float* partial_dev, float* total_sq_dev, mmvid_stream_t stream) {
  MMVID_REQUIRE(table_dev && chunk_tensor_dev && chunk_index_dev && partial_dev && total_sq_dev, "grad_sqnorm: null pointer");
  MMVID_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0, "grad_sqnorm: bad chunking");
  cudaStream_t st = (cudaStream_t)stream;
  grad_sqnorm_partial_kernel<<<n_chunks, OPT_THREADS, 0, st>>>(table_dev, chunk_tensor_dev, chunk_index_dev, chunk_elems,
                                                               partial_dev);
  count_launch();
  grad_sqnorm_final_kernel<<<1, 1024, 0, st>>>(partial_dev, n_chunks, total_sq_dev);
  return check_launch("grad_sqnorm");
}

int mmvid_grad_clip(const mmvid_adam_tensor* table_dev, const int* chunk_tensor_dev, const int* chunk_index_dev,
                    int n_chunks, int chunk_elems, const float* total_sq_dev, float max_norm, mmvid_stream_t stream) {
  MMVID_REQUIRE(table_dev && chunk_tensor_dev && chunk_index_dev && total_sq_dev, "grad_clip: null pointer");
  MMVID_REQUIRE(n_chunks > 0 && chunk_elems > 0 && max_norm > 0.f, "grad_clip: bad arguments");
  grad_clip_kernel<<<n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(table_dev, chunk_tensor_dev, chunk_index_dev,
                                                                     chunk_elems, total_sq_dev, max_norm);
  return check_launch("grad_clip");
}

int mmvid_adam_step(const mmvid_adam_tensor* table_dev, const int* chunk_tensor_dev, const int* chunk_index_dev,
                    int n_chunks, int chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int decoupled_weight_decay, int step, mmvid_stream_t stream) {
  MMVID_REQUIRE(table_dev && chunk_tensor_dev && chunk_index_dev, "adam_step: null pointer");
  MMVID_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0 && step >= 1, "adam_step: bad arguments");
  AdamScalars a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.decoupled = decoupled_weight_decay; a.step = step;
  // bias corrections in double on the host, exactly as torch's single-tensor path computes them from the python step
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.sqrt_bc2 = (float)sqrt(bc2);
  adam_step_kernel<<<n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(table_dev, chunk_tensor_dev, chunk_index_dev,
                                                                     chunk_elems, a);
  return check_launch("adam_step");
}

}  // extern "C"
