// HBM-bound kernels of the MMVID hot path: embedding gather (K1), LayerNorm (K2), row softmax,
// QKV split/transposes, GroupNorm+swish (K11), VQ argmin (K14), codebook gather (K15), layout helpers.
// All are coalesced along the channel (innermost) dimension with 128-bit accesses where alignment allows.
#include "common.cuh"
#include <stdlib.h>
#include <cuda_fp16.h>
#include <type_traits>

namespace mmvid {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

bool chained_launch_enabled() {
  static const bool on = [] {
    const char* v = getenv("MMVID_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}
}  // namespace mmvid

using namespace mmvid;

extern "C" int mmvid_version(void) { return 100; }
extern "C" const char* mmvid_last_error(void) { return g_err; }
extern "C" long long mmvid_launch_count(void) { return g_launches.load(); }
extern "C" void mmvid_reset_launch_count(void) { g_launches.store(0); }
extern "C" void mmvid_add_launch_count(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// K1 embedding gather: one warp-group row per block row; segments in constant-size param array.
// ------------------------------------------------------------------------------------------------
#define MMVID_MAX_SEGMENTS 8
struct EmbedParams {
  mmvid_embed_segment seg[MMVID_MAX_SEGMENTS];
  int nseg;
};

__global__ void embed_gather_kernel(float* __restrict__ out, int S, int D, EmbedParams p) {
  // grid.x = covered rows of one batch element (sum of seg.n), grid.y = B
  int r = blockIdx.x;
  const int b = blockIdx.y;
  int s = 0;
  while (s < p.nseg - 1 && r >= p.seg[s].n) { r -= p.seg[s].n; ++s; }
  const mmvid_embed_segment& g = p.seg[s];
  long long id = g.ids[(long long)b * g.ids_bstride + r];
  if (g.use_pad && id == g.pad_value) id = g.pad_base + r;
  if (g.table_rows > 0 && (id < 0 || id >= g.table_rows)) __trap();  // nn.Embedding raises on such an id
  const float4* t = reinterpret_cast<const float4*>(g.table + id * (long long)D);
  const float4* t2 = g.table2 ? reinterpret_cast<const float4*>(g.table2 + id * (long long)D) : nullptr;
  const float4* ps = g.pos ? reinterpret_cast<const float4*>(g.pos + (long long)r * D) : nullptr;
  float4* o = reinterpret_cast<float4*>(out + ((long long)b * S + g.seq_off + r) * D);
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 v = __ldg(t + c);
    if (t2) { float4 w = __ldg(t2 + c); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    if (ps) { float4 w = __ldg(ps + c); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    o[c] = v;
  }
}

extern "C" int mmvid_embed_gather(float* out, int B, int S, int D, const mmvid_embed_segment* segs, int nseg,
                                  mmvid_stream_t stream) {
  MMVID_REQUIRE(nseg >= 1 && nseg <= MMVID_MAX_SEGMENTS, "1..8 segments");
  MMVID_REQUIRE(D % 4 == 0, "D must be a multiple of 4");
  EmbedParams p;
  p.nseg = nseg;
  int rows = 0;
  for (int i = 0; i < nseg; ++i) {
    p.seg[i] = segs[i];
    MMVID_REQUIRE(segs[i].n > 0 && segs[i].seq_off >= 0 && segs[i].seq_off + segs[i].n <= S, "segment range");
    rows += segs[i].n;
  }
  if (B == 0) return MMVID_OK;
  embed_gather_kernel<<<dim3(rows, B), 192, 0, to_stream(stream)>>>(out, S, D, p);
  return check_launch("embed_gather");
}

__global__ void axial_table_kernel(float* __restrict__ out, int n, int D, const float* __restrict__ w0,
                                   const float* __restrict__ w1, const float* __restrict__ w2, int s0, int s1,
                                   int s2, int naxes) {
  const int i = blockIdx.x;
  if (i >= n) return;
  int c2 = 0, c1 = 0, c0 = 0;
  if (naxes == 3) { c2 = i % s2; c1 = (i / s2) % s1; c0 = i / (s2 * s1); }
  else if (naxes == 2) { c1 = i % s1; c0 = i / s1; }
  else { c0 = i; }
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v = w0[(long long)c0 * D + c];                 // python sum(): ((0 + e0) + e1) + e2
    if (naxes >= 2) v = v + w1[(long long)c1 * D + c];
    if (naxes >= 3) v = v + w2[(long long)c2 * D + c];
    out[(long long)i * D + c] = v;
  }
}

extern "C" int mmvid_axial_table(float* out, int n, int D, const float* w0, const float* w1, const float* w2,
                                 const int* shape, int naxes, mmvid_stream_t stream) {
  MMVID_REQUIRE(naxes >= 1 && naxes <= 3, "1..3 axes");
  int s0 = shape[0], s1 = naxes > 1 ? shape[1] : 1, s2 = naxes > 2 ? shape[2] : 1;
  MMVID_REQUIRE(n <= s0 * s1 * s2, "n exceeds axial volume");
  axial_table_kernel<<<n, 256, 0, to_stream(stream)>>>(out, n, D, w0, w1, w2, s0, s1, s2, naxes);
  return check_launch("axial_table");
}

// ------------------------------------------------------------------------------------------------
// K2 LayerNorm: one warp per row when D <= 1024 (row kept in registers: one HBM read, one write).
// Two-pass (mean, then centered variance) in fp32 like ATen's CPU/CUDA kernels.
// ------------------------------------------------------------------------------------------------
template <int VEC_PER_LANE, typename OutT>
__global__ void layernorm_warp_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, OutT* __restrict__ out, long long rows, int D,
                                      float eps) {
  chain_release();
  chain_wait();  // x is the previous kernel's result
  const int warps_per_block = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int nvec = D >> 2;
  float4 v[VEC_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) { v[i] = xr[c]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if constexpr (sizeof(OutT) == 4) {
        reinterpret_cast<float4*>(out + row * (long long)D)[c] = o;
      } else {
        uint2 pk;  // 16-bit outputs saturate instead of overflowing to inf (fp16: +-65504)
        if constexpr (std::is_same<OutT, __half>::value) {
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(o.y), "f"(o.x));
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(o.w), "f"(o.z));
        } else {
          asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(o.y), "f"(o.x));
          asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(o.w), "f"(o.z));
        }
        reinterpret_cast<uint2*>(out + row * (long long)D)[c] = pk;
      }
    }
  }
}

extern "C" int mmvid_layernorm(const float* x, long long ldx, const float* gamma, const float* beta, void* out,
                               int out_dtype, long long rows, int D, float eps, mmvid_stream_t stream) {
  MMVID_REQUIRE(D % 4 == 0 && D <= 1024 && ldx % 4 == 0, "D multiple of 4, <= 1024");
  if (rows == 0) return MMVID_OK;
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div<long long>(rows, wpb));
  cudaStream_t st = to_stream(stream);
  cudaError_t err;
  if (out_dtype == MMVID_DT_F32)
    err = launch_chained(layernorm_warp_kernel<8, float>, grid, dim3(wpb * 32), 0, st, x, ldx, gamma, beta, (float*)out, rows, D, eps);
  else if (out_dtype == MMVID_DT_F16)
    err = launch_chained(layernorm_warp_kernel<8, __half>, grid, dim3(wpb * 32), 0, st, x, ldx, gamma, beta, (__half*)out, rows, D,
                         eps);
  else
    err = launch_chained(layernorm_warp_kernel<8, __nv_bfloat16>, grid, dim3(wpb * 32), 0, st, x, ldx, gamma, beta,
                         (__nv_bfloat16*)out, rows, D, eps);
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "layernorm launch: %s", cudaGetErrorString(err));
  return check_launch("layernorm");
}

// ------------------------------------------------------------------------------------------------
// masked row softmax (fp32 parity path of SDPA / AttnBlock).  One block per row.
// ------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float* __restrict__ s, int rows, int cols, long long ld, int mask_kind,
                                    const int* __restrict__ prev_rows, int n_prev) {
  __shared__ float red[32];
  const long long r_all = blockIdx.x;
  const int row = (int)(r_all % rows);
  float* p = s + r_all * ld;
  int lo = 0, hi = cols;  // visible columns [lo, hi)
  if (mask_kind == MMVID_MASK_CAUSAL) hi = min(cols, row + 1);
  else if (mask_kind == MMVID_MASK_PREV) {
    for (int i = 0; i < n_prev; ++i) if (prev_rows[i] == row) lo = row;
  }
  float m = -INFINITY;
  for (int c = lo + threadIdx.x; c < hi; c += blockDim.x) m = fmaxf(m, p[c]);
  m = block_max(m, red);
  float sum = 0.f;
  for (int c = lo + threadIdx.x; c < hi; c += blockDim.x) { float e = expf(p[c] - m); p[c] = e; sum += e; }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) p[c] = (c >= lo && c < hi) ? p[c] * inv : 0.f;
}

extern "C" int mmvid_softmax_rows(float* scores, long long batch, int rows, int cols, long long ld, int mask_kind,
                                  const int* prev_rows, int n_prev, mmvid_stream_t stream) {
  if (batch * rows == 0) return MMVID_OK;
  MMVID_REQUIRE(batch * rows < (1ll << 31), "too many rows");
  softmax_rows_kernel<<<(unsigned)(batch * rows), 256, 0, to_stream(stream)>>>(scores, rows, cols, ld, mask_kind,
                                                                                 prev_rows, n_prev);
  return check_launch("softmax_rows");
}

// logits -> probs with optional additive noise (sampling head, dalle_bert.py:527-531)
__global__ void softmax_logits_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                      float noise_scale, float* __restrict__ probs, int n) {
  __shared__ float red[32];
  const long long r = blockIdx.x;
  const float* l = logits + r * n;
  const float* z = noise ? noise + r * n : nullptr;
  float* p = probs + r * n;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    float v = l[c];
    if (z) v = v + noise_scale * z[c];
    p[c] = v;
    m = fmaxf(m, v);
  }
  m = block_max(m, red);
  float sum = 0.f;
  for (int c = threadIdx.x; c < n; c += blockDim.x) { float e = expf(p[c] - m); p[c] = e; sum += e; }
  sum = block_sum(sum, red);
  for (int c = threadIdx.x; c < n; c += blockDim.x) p[c] = p[c] / sum;
}

extern "C" int mmvid_softmax_logits(const float* logits, const float* noise, float noise_scale, float* probs,
                                    long long rows, int n, mmvid_stream_t stream) {
  if (rows == 0) return MMVID_OK;
  softmax_logits_kernel<<<(unsigned)rows, 256, 0, to_stream(stream)>>>(logits, noise, noise_scale, probs, n);
  return check_launch("softmax_logits");
}

// ------------------------------------------------------------------------------------------------
// qkv [B*S, 3*H*64] -> q,k [B,H,S_pad,64]  vt [B,H,64,S_pad]   (zero padded rows)
// block = 64 tokens x one head; V is transposed through shared memory so both sides stay coalesced.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }

template <typename T>
__global__ void qkv_split_kernel(const float* __restrict__ qkv, T* __restrict__ q, T* __restrict__ k,
                                 T* __restrict__ vt, int H, int S, int S_pad) {
  __shared__ float tile[64][65];
  const int s0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int D3 = 3 * H * 64, D = H * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 256 threads: 64 x 4
  const long long head = ((long long)b * H + h);
  for (int r = ty; r < 64; r += 4) {
    const int s = s0 + r;
    float vq = 0.f, vk = 0.f, vv = 0.f;
    if (s < S) {
      const float* row = qkv + ((long long)b * S + s) * D3 + h * 64 + tx;
      vq = row[0]; vk = row[D]; vv = row[2 * D];
    }
    q[(head * S_pad + s) * 64 + tx] = cvt_out<T>(vq);
    k[(head * S_pad + s) * 64 + tx] = cvt_out<T>(vk);
    tile[r][tx] = vv;
  }
  __syncthreads();
  for (int d = ty; d < 64; d += 4) vt[(head * 64 + d) * S_pad + s0 + tx] = cvt_out<T>(tile[tx][d]);
}

extern "C" int mmvid_qkv_split(const float* qkv, void* q, void* k, void* vt, int dtype, int B, int H, int S, int S_pad,
                               mmvid_stream_t stream) {
  MMVID_REQUIRE(S_pad % 64 == 0 && S_pad >= S, "S_pad multiple of 64");
  dim3 grid(S_pad / 64, H, B);
  if (dtype == MMVID_DT_F32)
    qkv_split_kernel<float><<<grid, 256, 0, to_stream(stream)>>>(qkv, (float*)q, (float*)k, (float*)vt, H, S, S_pad);
  else if (dtype == MMVID_DT_F16)
    qkv_split_kernel<__half><<<grid, 256, 0, to_stream(stream)>>>(qkv, (__half*)q, (__half*)k, (__half*)vt, H, S, S_pad);
  else
    qkv_split_kernel<__nv_bfloat16><<<grid, 256, 0, to_stream(stream)>>>(qkv, (__nv_bfloat16*)q, (__nv_bfloat16*)k,
                                                                        (__nv_bfloat16*)vt, H, S, S_pad);
  return check_launch("qkv_split");
}

// ------------------------------------------------------------------------------------------------
// NCHW <-> NHWC (32x32 smem tile transpose), nearest upsample
// ------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cn) {
  // in [batch, R, Cn] -> out [batch, Cn, R]
  __shared__ float t[32][33];
  const long long base = (long long)blockIdx.z * R * Cn;
  int r = blockIdx.y * 32 + threadIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  for (int i = 0; i < 32; i += 8)
    if (r + i < R && c < Cn) t[threadIdx.y + i][threadIdx.x] = in[base + (long long)(r + i) * Cn + c];
  __syncthreads();
  r = blockIdx.y * 32 + threadIdx.x; c = blockIdx.x * 32 + threadIdx.y;
  for (int i = 0; i < 32; i += 8)
    if (c + i < Cn && r < R) out[base + (long long)(c + i) * R + r] = t[threadIdx.x][threadIdx.y + i];
}

extern "C" int mmvid_nchw_to_nhwc(const float* in, float* out, int N, int C, int HW, mmvid_stream_t stream) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  transpose_kernel<<<grid, dim3(32, 8), 0, to_stream(stream)>>>(in, out, C, HW);
  return check_launch("nchw_to_nhwc");
}
extern "C" int mmvid_nhwc_to_nchw(const float* in, float* out, int N, int C, int HW, mmvid_stream_t stream) {
  dim3 grid(ceil_div(C, 32), ceil_div(HW, 32), N);
  transpose_kernel<<<grid, dim3(32, 8), 0, to_stream(stream)>>>(in, out, HW, C);
  return check_launch("nhwc_to_nchw");
}

// four floats -> four fp16 values (round to nearest, saturating) in one 8-byte word
__device__ __forceinline__ uint2 pack4_f16(float4 v) {
  uint2 pk;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(v.y), "f"(v.x));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(v.w), "f"(v.z));
  return pk;
}

template <bool OUT16>
__global__ void upsample2x_kernel(const float4* __restrict__ in, void* __restrict__ out, int H, int W, int C4,
                                  long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long p = i / C4;
    const int x = (int)(p % (2 * W)); p /= (2 * W);
    const int y = (int)(p % (2 * H));
    const long long n = p / (2 * H);
    const float4 v = in[((n * H + (y >> 1)) * W + (x >> 1)) * C4 + c];
    if (OUT16) reinterpret_cast<uint2*>(out)[i] = pack4_f16(v);
    else reinterpret_cast<float4*>(out)[i] = v;
  }
}
extern "C" int mmvid_upsample2x(const float* in, void* out, int out_dtype, int N, int H, int W, int C, mmvid_stream_t stream) {
  MMVID_REQUIRE(C % 4 == 0, "C multiple of 4");
  MMVID_REQUIRE(out_dtype == MMVID_DT_F32 || out_dtype == MMVID_DT_F16, "fp32 or fp16 output");
  const long long total = (long long)N * 4 * H * W * (C / 4);
  const int blocks = (int)std::min<long long>(ceil_div<long long>(total, 256), 148 * 16);
  if (out_dtype == MMVID_DT_F16)
    upsample2x_kernel<true><<<blocks, 256, 0, to_stream(stream)>>>((const float4*)in, out, H, W, C / 4, total);
  else
    upsample2x_kernel<false><<<blocks, 256, 0, to_stream(stream)>>>((const float4*)in, out, H, W, C / 4, total);
  return check_launch("upsample2x");
}

// ------------------------------------------------------------------------------------------------
// K11 GroupNorm(+swish) on NHWC.
// Pass 1 (one HBM read, fully coalesced): each block owns a contiguous slab of pixels of one image and walks
// it row by row ([pixels, C] is contiguous), every thread accumulating SHIFTED first/second moments of its
// own 4 channels: s1 = sum(x-K), s2 = sum((x-K)^2) with K = that channel's value at pixel 0 of the image
// (keeps the cancellation in var = E[(x-K)^2] - E[x-K]^2 mild).  Threads -> groups through shared memory;
// per-(image, slab, group) partials go to scratch and are combined in a FIXED order by the finalize kernel,
// so the statistics are deterministic (no atomics) - bit-exact VQ ids need run-to-run reproducibility.
// Pass 2: elementwise normalise * gamma + beta (+ swish).
// ------------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_SLABS = 64;

__global__ void __launch_bounds__(GN_THREADS) groupnorm_partial_kernel(const float* __restrict__ in,
                                                                      float* __restrict__ partial, int HW, int C, int G,
                                                                      int slabs) {
  extern __shared__ float sm[];  // [2][C]
  const int n = blockIdx.y, slab = blockIdx.x;
  const int C4 = C >> 2;
  const int lanes_per_row = C4;                       // threads needed to cover one pixel row
  const int rows_per_iter = GN_THREADS / lanes_per_row;  // C <= 1024 -> >= 1
  const int c4 = threadIdx.x % lanes_per_row, rsub = threadIdx.x / lanes_per_row;
  const long long img = (long long)n * HW * C;
  const int p_per_slab = (HW + slabs - 1) / slabs;
  const int p0 = slab * p_per_slab, p1 = min(HW, p0 + p_per_slab);
  const float4 K = *reinterpret_cast<const float4*>(in + img + c4 * 4);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (rsub < rows_per_iter) {
    for (int p = p0 + rsub; p < p1; p += rows_per_iter) {
      const float4 v = *reinterpret_cast<const float4*>(in + img + (long long)p * C + c4 * 4);
      float d;
      d = v.x - K.x; s1[0] += d; s2[0] = fmaf(d, d, s2[0]);
      d = v.y - K.y; s1[1] += d; s2[1] = fmaf(d, d, s2[1]);
      d = v.z - K.z; s1[2] += d; s2[2] = fmaf(d, d, s2[2]);
      d = v.w - K.w; s1[3] += d; s2[3] = fmaf(d, d, s2[3]);
    }
  }
  // reduce the row sub-lanes per channel, in a fixed order
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) sm[i] = 0.f;
  __syncthreads();
  for (int r = 0; r < rows_per_iter; ++r) {
    if (rsub == r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { sm[c4 * 4 + j] += s1[j]; sm[C + c4 * 4 + j] += s2[j]; }
    }
    __syncthreads();
  }
  // per group: combine channels.  With shift K_c per channel: sum(x) = s1_c + cnt*K_c ; sum over the group of
  // centered moments is finished in the finalize kernel from (cnt, sum x, sum (x-K_c)^2, K_c) -> we emit, per channel
  // quad, moments re-expressed about the GROUP shift Kg = K of the group's first channel.
  const int cpg = C / G;
  const float cnt = (float)(p1 - p0);
  for (int g = threadIdx.x; g < G; g += GN_THREADS) {
    const float Kg = in[img + g * cpg];
    float a1 = 0.f, a2 = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float Kc = in[img + c];
      const float dk = Kc - Kg;  // x - Kg = (x - Kc) + dk
      a1 += sm[c] + cnt * dk;
      a2 += sm[C + c] + 2.f * dk * sm[c] + cnt * dk * dk;
    }
    float* o = partial + (((long long)n * slabs + slab) * G + g) * 2;
    o[0] = a1; o[1] = a2;
  }
}

__global__ void groupnorm_finalize_kernel(const float* __restrict__ in, const float* __restrict__ partial,
                                          float* __restrict__ stats, int HW, int C, int G, int slabs, float eps) {
  const int n = blockIdx.x;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float a1 = 0.f, a2 = 0.f;
    for (int s = 0; s < slabs; ++s) {
      const float* o = partial + (((long long)n * slabs + s) * G + g) * 2;
      a1 += o[0]; a2 += o[1];
    }
    const int cpg = C / G;
    const float cnt = (float)HW * cpg;
    const float Kg = in[(long long)n * HW * C + g * cpg];
    const float m1 = a1 / cnt;
    const float var = fmaxf(a2 / cnt - m1 * m1, 0.f);
    stats[((long long)n * G + g) * 2 + 0] = Kg + m1;
    stats[((long long)n * G + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

__global__ void groupnorm_apply_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ stats, long long HW, int C, int G, int swish,
                                       long long total4) {
  const int C4 = C >> 2, cpg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long n = i / (C4 * HW);
    float4 v = in[i];
    float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;
      const float mean = stats[(n * G + g) * 2], rstd = stats[(n * G + g) * 2 + 1];
      float y = (r[j] - mean) * rstd * __ldg(gamma + c + j) + __ldg(beta + c + j);
      if (swish) y = y / (1.f + expf(-y));
      r[j] = y;
    }
    out[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// Streaming variant used when 256 % (C/4) == 0 (every VQGAN width): a thread keeps ONE channel quad for its whole life,
// so gamma / beta / the (frame, group) statistics are loaded once per frame instead of once per element and there is
// no 64-bit division in the loop; four pixels are in flight per thread.  ncu had the generic kernel above at 2.8 TB/s
// (instruction bound: 64-bit div / mod, per-element statistic loads, accurate expf + IEEE division).
// swish: 0 none, 1 x * sigmoid(x) with expf and IEEE division (fp32 parity mode), 2 MUFU ex2 + rcp (~1e-6 relative).
// OUT16: the result feeds a kind::f16 implicit-GEMM conv and is written as fp16 (half the store bytes).
template <int SWISH, bool OUT16>
__global__ void __launch_bounds__(256) groupnorm_apply2_kernel(const float4* __restrict__ in, void* __restrict__ out,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ stats, int HW, int C, int G) {
  const int C4 = C >> 2, cpg = C / G;
  const int c4 = threadIdx.x % C4, prow = threadIdx.x / C4, ppb = 256 / C4;
  const int n = blockIdx.y, c = c4 * 4;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c)), bt = __ldg(reinterpret_cast<const float4*>(beta + c));
  const float g4[4] = {gm.x, gm.y, gm.z, gm.w}, b4[4] = {bt.x, bt.y, bt.z, bt.w};
  float mean[4], rstd[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g = (c + j) / cpg;
    mean[j] = stats[((long long)n * G + g) * 2];
    rstd[j] = stats[((long long)n * G + g) * 2 + 1];
  }
  auto f = [&](float4 v) {
    float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y = (r[j] - mean[j]) * rstd[j] * g4[j] + b4[j];  // same operation order as the generic kernel
      if (SWISH == 1) y = y / (1.f + expf(-y));
      if (SWISH == 2) y = apply_act_fast(y, MMVID_ACT_SWISH);
      r[j] = y;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
  };
  const float4* src = in + (long long)n * HW * C4 + c4;
  const long long obase = (long long)n * HW * C4 + c4;
  auto put = [&](long long pix, float4 v) {
    if (OUT16) reinterpret_cast<uint2*>(out)[obase + pix * C4] = pack4_f16(v);
    else reinterpret_cast<float4*>(out)[obase + pix * C4] = v;
  };
  const int step = gridDim.x * ppb;
  int p = blockIdx.x * ppb + prow;
  for (; p + 3 * step < HW; p += 4 * step) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = src[(long long)(p + u * step) * C4];
#pragma unroll
    for (int u = 0; u < 4; ++u) put(p + u * step, f(v[u]));
  }
  for (; p < HW; p += step) put(p, f(src[(long long)p * C4]));
}

static int groupnorm_stats_launch(const float* in, float* stats, int N, int HW, int C, int groups, float eps,
                                  cudaStream_t st) {
  // scratch layout: [N*G*2 final stats][N*slabs*G*2 partials]
  int slabs = (int)std::min<long long>(GN_MAX_SLABS, std::max<long long>(1, (148LL * 4) / std::max(N, 1)));
  slabs = std::min(slabs, std::max(1, HW / 16));
  float* partial = stats + (long long)N * groups * 2;
  groupnorm_partial_kernel<<<dim3(slabs, N), GN_THREADS, 2 * C * sizeof(float), st>>>(in, partial, HW, C, groups, slabs);
  int rc = check_launch("groupnorm_partial");
  if (rc) return rc;
  groupnorm_finalize_kernel<<<N, 32, 0, st>>>(in, partial, stats, HW, C, groups, slabs, eps);
  return check_launch("groupnorm_finalize");
}

extern "C" long long mmvid_groupnorm_scratch_floats(int N, int groups) {
  return (long long)N * groups * 2 * (1 + GN_MAX_SLABS);
}

extern "C" int mmvid_groupnorm_stats(const float* in, float* stats, int N, int HW, int C, int groups, float eps,
                                     mmvid_stream_t stream) {
  MMVID_REQUIRE(C % groups == 0 && C % 4 == 0 && C <= 1024, "C divisible by groups and 4, <= 1024");
  if (N == 0) return MMVID_OK;
  return groupnorm_stats_launch(in, stats, N, HW, C, groups, eps, to_stream(stream));
}

static int groupnorm_apply_launch(const float* in, void* out, int out_dtype, const float* gamma, const float* beta,
                                  const float* stats, int N, int HW, int C, int groups, int swish, cudaStream_t st) {
  const int C4 = C / 4;
  if (C4 <= 256 && 256 % C4 == 0 && N <= 65535) {
    const int ppb = 256 / C4;
    int gx = std::max(1, std::min(ceil_div(HW, ppb * 4), ceil_div(148 * 8, N)));
    dim3 grid(gx, N);
    const float4* i4 = (const float4*)in;
    if (out_dtype == MMVID_DT_F16) {
      if (swish == 0) groupnorm_apply2_kernel<0, true><<<grid, 256, 0, st>>>(i4, out, gamma, beta, stats, HW, C, groups);
      else if (swish == 2) groupnorm_apply2_kernel<2, true><<<grid, 256, 0, st>>>(i4, out, gamma, beta, stats, HW, C, groups);
      else groupnorm_apply2_kernel<1, true><<<grid, 256, 0, st>>>(i4, out, gamma, beta, stats, HW, C, groups);
    } else {
      if (swish == 0) groupnorm_apply2_kernel<0, false><<<grid, 256, 0, st>>>(i4, out, gamma, beta, stats, HW, C, groups);
      else if (swish == 2) groupnorm_apply2_kernel<2, false><<<grid, 256, 0, st>>>(i4, out, gamma, beta, stats, HW, C, groups);
      else groupnorm_apply2_kernel<1, false><<<grid, 256, 0, st>>>(i4, out, gamma, beta, stats, HW, C, groups);
    }
    return check_launch("groupnorm_apply2");
  }
  MMVID_REQUIRE(out_dtype == MMVID_DT_F32, "fp16 GroupNorm output needs C / 4 to divide 256");
  const long long total4 = (long long)N * HW * (C / 4);
  const int blocks = (int)std::min<long long>(ceil_div<long long>(total4, 256), 148 * 16);
  groupnorm_apply_kernel<<<blocks, 256, 0, st>>>((const float4*)in, (float4*)out, gamma, beta, stats, HW, C, groups,
                                                  swish, total4);
  return check_launch("groupnorm_apply");
}

extern "C" int mmvid_groupnorm(const float* in, void* out, int out_dtype, const float* gamma, const float* beta,
                               float* stats, int N, int HW, int C, int groups, float eps, int swish,
                               mmvid_stream_t stream) {
  MMVID_REQUIRE(C % groups == 0 && C % 4 == 0 && C <= 1024, "C divisible by groups and 4, <= 1024");
  MMVID_REQUIRE(out_dtype == MMVID_DT_F32 || out_dtype == MMVID_DT_F16, "fp32 or fp16 output");
  if (N == 0) return MMVID_OK;
  cudaStream_t st = to_stream(stream);
  int rc = groupnorm_stats_launch(in, stats, N, HW, C, groups, eps, st);
  if (rc) return rc;
  return groupnorm_apply_launch(in, out, out_dtype, gamma, beta, stats, N, HW, C, groups, swish, st);
}

// Statistics from the partial statistics a tensor-core conv wrote next to its result (EpiArgs::gn_partial, tc_gemm.cu):
// partial[(slab * G + g) * 2 + {0, 1}] = (sum, sum of squared deviations from the slab's own mean) of 32 pixels x (C / G)
// channels, HW / 32 slabs per image.  One block per image, thread = (group, slab lane); the slabs are combined with the
// parallel-variance formula  M2 = sum_i [M2_i + n_i (mean_i - mean)^2]  in double and in a fixed order (deterministic).
__global__ void __launch_bounds__(1024) groupnorm_finalize_partials_kernel(const float* __restrict__ partial,
                                                                           float* __restrict__ stats, int spi, int G,
                                                                           int n_slab, float eps) {
  __shared__ double red[1024];
  __shared__ double mean_s[512];
  const int n = blockIdx.x, g = threadIdx.x % G, sl = threadIdx.x / G, nsl = blockDim.x / G;
  const float2* base = reinterpret_cast<const float2*>(partial) + (long long)n * spi * G + g;
  double acc = 0.0;
  for (int s = sl; s < spi; s += nsl) acc += (double)base[(long long)s * G].x;
  red[threadIdx.x] = acc;
  __syncthreads();
  if (sl == 0) {
    for (int k = 1; k < nsl; ++k) acc += red[k * G + g];
    mean_s[g] = acc / ((double)spi * n_slab);
  }
  __syncthreads();
  const double mean = mean_s[g], inv_n = 1.0 / n_slab;
  acc = 0.0;
  for (int s = sl; s < spi; s += nsl) {
    const float2 v = base[(long long)s * G];
    const double d = (double)v.x * inv_n - mean;
    acc += (double)v.y + n_slab * d * d;
  }
  __syncthreads();
  red[threadIdx.x] = acc;
  __syncthreads();
  if (sl == 0) {
    for (int k = 1; k < nsl; ++k) acc += red[k * G + g];
    const double var = acc / ((double)spi * n_slab);
    stats[((long long)n * G + g) * 2 + 0] = (float)mean;
    stats[((long long)n * G + g) * 2 + 1] = rsqrtf((float)var + eps);
  }
}

extern "C" int mmvid_groupnorm_from_partials(const float* in, void* out, int out_dtype, const float* gamma, const float* beta,
                                             const float* partial, float* stats, int N, int HW, int C, int groups, float eps,
                                             int swish, mmvid_stream_t stream) {
  MMVID_REQUIRE(C % groups == 0 && C % 4 == 0 && C <= 1024, "C divisible by groups and 4, <= 1024");
  MMVID_REQUIRE(out_dtype == MMVID_DT_F32 || out_dtype == MMVID_DT_F16, "fp32 or fp16 output");
  MMVID_REQUIRE(HW % 32 == 0 && groups >= 1 && groups <= 512 && 1024 % groups == 0, "HW % 32 == 0, groups divides 1024");
  if (N == 0) return MMVID_OK;
  cudaStream_t st = to_stream(stream);
  groupnorm_finalize_partials_kernel<<<N, 1024, 0, st>>>(partial, stats, HW / 32, groups, 32 * (C / groups), eps);
  int rc = check_launch("groupnorm_finalize_partials");
  if (rc) return rc;
  return groupnorm_apply_launch(in, out, out_dtype, gamma, beta, stats, N, HW, C, groups, swish, st);
}

// ------------------------------------------------------------------------------------------------
// Decoder tail, fused:  GroupNorm-apply + swish + 3x3 conv to <= 4 output channels + clamp/rescale, NCHW out
// (Decoder.norm_out -> nonlinearity -> conv_out, model.py:578-581, + vae.py:55).  The generic implicit GEMM wastes
// a 64-wide N tile on 3 output channels and the normalised activation would make an extra 2x-tensor HBM round
// trip; here the raw activation is read once (halo tile in smem, normalised on the fly) and 3 floats/pixel leave.
// Block = 16 x 32 output pixels, 256 threads, each thread owns 2 pixels; channels are processed 32 at a time.
// ------------------------------------------------------------------------------------------------
constexpr int CO_TH = 16, CO_TW = 32, CO_CH = 32, CO_LD = 36;

template <bool FAST>
__global__ void __launch_bounds__(256) conv_out_fused_kernel(const float* __restrict__ in, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ w /*[Cout][3][3][C]*/,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            int H, int W, int C, int G, int Cout, int post_clamp) {
  extern __shared__ float smem[];
  float* tile = smem;                                   // [(CO_TH+2)*(CO_TW+2)][CO_LD]
  float* wsm = smem + (CO_TH + 2) * (CO_TW + 2) * CO_LD;  // [4][9][CO_CH]
  const int n = blockIdx.z, y0 = blockIdx.y * CO_TH, x0 = blockIdx.x * CO_TW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps: rows ty and ty+8
  const int cpg = C / G;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int HP = CO_TH + 2, WP = CO_TW + 2;
  for (int c0 = 0; c0 < C; c0 += CO_CH) {
    __syncthreads();
    // halo tile, normalised + swish on load; zero outside the image (conv padding applies to the activated tensor)
    for (int i = threadIdx.x; i < HP * WP * (CO_CH / 4); i += 256) {
      const int c4 = i % (CO_CH / 4), p = i / (CO_CH / 4);
      const int py = p / WP, px = p - py * WP;
      const int y = y0 + py - 1, x = x0 + px - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < H && x >= 0 && x < W) {
        v = *reinterpret_cast<const float4*>(in + (((long long)n * H + y) * W + x) * C + c0 + c4 * 4);
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + c4 * 4 + j, g = c / cpg;
          const float mean = stats[((long long)n * G + g) * 2], rstd = stats[((long long)n * G + g) * 2 + 1];
          float t = (r[j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
          r[j] = FAST ? apply_act_fast(t, MMVID_ACT_SWISH) : t / (1.f + expf(-t));
        }
        v = make_float4(r[0], r[1], r[2], r[3]);
      }
      *reinterpret_cast<float4*>(tile + (size_t)p * CO_LD + c4 * 4) = v;
    }
    for (int i = threadIdx.x; i < 4 * 9 * CO_CH; i += 256) {
      const int c = i % CO_CH, tap = (i / CO_CH) % 9, co = i / (CO_CH * 9);
      wsm[i] = co < Cout ? w[((long long)co * 9 + tap) * C + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - ky * 3;
      const float* a0 = tile + (size_t)((ty + ky) * WP + tx + kx) * CO_LD;
      const float* a1 = tile + (size_t)((ty + 8 + ky) * WP + tx + kx) * CO_LD;
#pragma unroll
      for (int c4 = 0; c4 < CO_CH / 4; ++c4) {
        const float4 p0 = *reinterpret_cast<const float4*>(a0 + c4 * 4);
        const float4 p1 = *reinterpret_cast<const float4*>(a1 + c4 * 4);
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          const float4 wv = *reinterpret_cast<const float4*>(wsm + (co * 9 + tap) * CO_CH + c4 * 4);
          acc[0][co] = fmaf(p0.x, wv.x, acc[0][co]); acc[0][co] = fmaf(p0.y, wv.y, acc[0][co]);
          acc[0][co] = fmaf(p0.z, wv.z, acc[0][co]); acc[0][co] = fmaf(p0.w, wv.w, acc[0][co]);
          acc[1][co] = fmaf(p1.x, wv.x, acc[1][co]); acc[1][co] = fmaf(p1.y, wv.y, acc[1][co]);
          acc[1][co] = fmaf(p1.z, wv.z, acc[1][co]); acc[1][co] = fmaf(p1.w, wv.w, acc[1][co]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int y = y0 + ty + i * 8, x = x0 + tx;
    if (y >= H || x >= W) continue;
    for (int co = 0; co < Cout; ++co) {
      float v = acc[i][co] + (bias ? bias[co] : 0.f);
      if (post_clamp) v = (fminf(fmaxf(v, -1.f), 1.f) + 1.f) * 0.5f;
      out[(((long long)n * Cout + co) * H + y) * W + x] = v;  // NCHW: a warp writes 32 consecutive x
    }
  }
}

static int conv_out_fused_launch(const float* in, const float* gamma, const float* beta, const float* w, const float* bias,
                                 float* out, const float* stats, int N, int H, int W, int C, int Cout, int groups,
                                 int post_clamp, cudaStream_t st) {
  const size_t smem = ((size_t)(CO_TH + 2) * (CO_TW + 2) * CO_LD + 4 * 9 * CO_CH) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_out_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(conv_out_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  dim3 grid(ceil_div(W, CO_TW), ceil_div(H, CO_TH), N);
  // post_clamp bit 0: clamp + rescale to [0, 1]; bit 1: swish through MUFU ex2 / rcp instead of expf + IEEE division
  if (post_clamp & 2)
    conv_out_fused_kernel<true><<<grid, 256, smem, st>>>(in, stats, gamma, beta, w, bias, out, H, W, C, groups, Cout, post_clamp & 1);
  else
    conv_out_fused_kernel<false><<<grid, 256, smem, st>>>(in, stats, gamma, beta, w, bias, out, H, W, C, groups, Cout, post_clamp & 1);
  return check_launch("conv_out_fused");
}

extern "C" int mmvid_conv_out_fused(const float* in, const float* gamma, const float* beta, const float* w,
                                    const float* bias, float* out, float* stats_scratch, int N, int H, int W, int C,
                                    int Cout, int groups, float eps, int post_clamp, mmvid_stream_t stream) {
  MMVID_REQUIRE(Cout >= 1 && Cout <= 4, "1..4 output channels");
  MMVID_REQUIRE(C % CO_CH == 0 && C % groups == 0, "C multiple of 32 and of groups");
  if (N == 0) return MMVID_OK;
  cudaStream_t st = to_stream(stream);
  int rc = groupnorm_stats_launch(in, stats_scratch, N, H * W, C, groups, eps, st);
  if (rc) return rc;
  return conv_out_fused_launch(in, gamma, beta, w, bias, out, stats_scratch, N, H, W, C, Cout, groups, post_clamp, st);
}

// the same with the statistics taken from the producing conv's fused partial sums (mmvid_conv_params::gn_partial)
extern "C" int mmvid_conv_out_fused_from_partials(const float* in, const float* gamma, const float* beta, const float* w,
                                                  const float* bias, float* out, const float* partial, float* stats, int N,
                                                  int H, int W, int C, int Cout, int groups, float eps, int post_clamp,
                                                  mmvid_stream_t stream) {
  MMVID_REQUIRE(Cout >= 1 && Cout <= 4, "1..4 output channels");
  MMVID_REQUIRE(C % CO_CH == 0 && C % groups == 0, "C multiple of 32 and of groups");
  MMVID_REQUIRE((H * W) % 32 == 0 && groups <= 512 && 1024 % groups == 0, "H*W % 32 == 0, groups divides 1024");
  if (N == 0) return MMVID_OK;
  cudaStream_t st = to_stream(stream);
  groupnorm_finalize_partials_kernel<<<N, 1024, 0, st>>>(partial, stats, H * W / 32, groups, 32 * (C / groups), eps);
  int rc = check_launch("groupnorm_finalize_partials");
  if (rc) return rc;
  return conv_out_fused_launch(in, gamma, beta, w, bias, out, stats, N, H, W, C, Cout, groups, post_clamp, st);
}

// ------------------------------------------------------------------------------------------------
// K14 VQ argmin.  One warp per latent row; the codebook (1024 x 256 fp32 = 1 MB) streams from L2.
// Summation order restates quantize.py:306-308:  d = (sum z^2 + sum e^2) - 2 * dot ; lowest index wins ties.
// Each block stages 8 rows of z in smem; each warp owns one row and loops over codes; lanes split the
// 256-dim dot product (8 floats per lane), so codebook reads are fully coalesced 1 KB rows.
// ------------------------------------------------------------------------------------------------
__global__ void code_sqnorm_kernel(const float* __restrict__ cb, float* __restrict__ e2, int n_codes, int dim) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n_codes) return;
  float s = 0.f;
  for (int c = threadIdx.x & 31; c < dim; c += 32) { float v = cb[(long long)j * dim + c]; s += v * v; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) e2[j] = s;
}

template <int ROWS>
__global__ void vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                 const float* __restrict__ e2, int64_t* __restrict__ idx, long long T, int n_codes,
                                 int dim) {
  // block: 256 threads; thread t handles codes t, t+256, ... for ROWS rows staged in smem.
  extern __shared__ float zs[];  // [ROWS][dim]
  __shared__ float z2[ROWS];
  __shared__ float best_d[ROWS][8];
  __shared__ int best_i[ROWS][8];
  const long long row0 = (long long)blockIdx.x * ROWS;
  for (int i = threadIdx.x; i < ROWS * dim; i += blockDim.x) {
    const long long r = row0 + i / dim;
    zs[i] = r < T ? z[r * dim + (i % dim)] : 0.f;
  }
  __syncthreads();
  {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = w; r < ROWS; r += 8) {
      float s = 0.f;
      for (int c = lane; c < dim; c += 32) s += zs[r * dim + c] * zs[r * dim + c];
      s = warp_sum(s);
      if (lane == 0) z2[r] = s;
    }
  }
  __syncthreads();
  float bd[ROWS];
  int bi[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) { bd[r] = INFINITY; bi[r] = 0x7fffffff; }
  for (int j = threadIdx.x; j < n_codes; j += blockDim.x) {
    const float4* e = reinterpret_cast<const float4*>(cb + (long long)j * dim);
    float dot[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) dot[r] = 0.f;
    for (int c = 0; c < dim / 4; ++c) {
      const float4 ev = __ldg(e + c);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float4 zv = reinterpret_cast<const float4*>(zs + r * dim)[c];
        dot[r] = fmaf(zv.x, ev.x, dot[r]); dot[r] = fmaf(zv.y, ev.y, dot[r]);
        dot[r] = fmaf(zv.z, ev.z, dot[r]); dot[r] = fmaf(zv.w, ev.w, dot[r]);
      }
    }
    const float ej = e2[j];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float d = (z2[r] + ej) - 2.f * dot[r];
      if (d < bd[r]) { bd[r] = d; bi[r] = j; }   // j increases per thread: strict < keeps the lowest index
    }
  }
  // reduce (min d, then min index) across the block
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float d = bd[r]; int i = bi[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (od < d || (od == d && oi < i)) { d = od; i = oi; }
    }
    if (lane == 0) { best_d[r][w] = d; best_i[r][w] = i; }
  }
  __syncthreads();
  if (threadIdx.x < ROWS) {
    const int r = threadIdx.x;
    float d = best_d[r][0]; int i = best_i[r][0];
    for (int k = 1; k < 8; ++k) {
      if (best_d[r][k] < d || (best_d[r][k] == d && best_i[r][k] < i)) { d = best_d[r][k]; i = best_i[r][k]; }
    }
    if (row0 + r < T) idx[row0 + r] = i;
  }
}

extern "C" int mmvid_vq_argmin(const float* z, const float* codebook, int64_t* idx, float* e2, long long T, int n_codes,
                               int dim, mmvid_stream_t stream) {
  MMVID_REQUIRE(dim % 4 == 0 && dim <= 1024, "dim multiple of 4");
  MMVID_REQUIRE(e2 != nullptr, "e2_scratch [n_codes] is caller-owned");
  if (T == 0) return MMVID_OK;
  cudaStream_t st = to_stream(stream);
  code_sqnorm_kernel<<<ceil_div(n_codes, 8), 256, 0, st>>>(codebook, e2, n_codes, dim);
  int rc = check_launch("code_sqnorm");
  if (rc) return rc;
  constexpr int ROWS = 4;
  vq_argmin_kernel<ROWS><<<(unsigned)ceil_div<long long>(T, ROWS), 256, ROWS * dim * sizeof(float), st>>>(
      z, codebook, e2, idx, T, n_codes, dim);
  return check_launch("vq_argmin");
}

__global__ void codebook_gather_kernel(const int64_t* __restrict__ ids, const float4* __restrict__ cb,
                                       float4* __restrict__ out, long long T, int dim4) {
  const long long t = (long long)blockIdx.x * (blockDim.x / 64) + threadIdx.x / 64;
  if (t >= T) return;
  const long long id = ids[t];
  for (int c = threadIdx.x % 64; c < dim4; c += 64) out[t * dim4 + c] = __ldg(cb + id * dim4 + c);
}
extern "C" int mmvid_codebook_gather(const int64_t* ids, const float* codebook, float* out, long long T, int dim,
                                     mmvid_stream_t stream) {
  MMVID_REQUIRE(dim % 4 == 0, "dim multiple of 4");
  if (T == 0) return MMVID_OK;
  codebook_gather_kernel<<<(unsigned)ceil_div<long long>(T, 4), 256, 0, to_stream(stream)>>>(
      ids, (const float4*)codebook, (float4*)out, T, dim / 4);
  return check_launch("codebook_gather");
}

// ================================================================================================
// Training-path kernels (BERT.forward(return_loss=True), dalle_bert.py:980-1127): elementwise / reduction
// backward pieces.  The GEMM-shaped backward work reuses mmvid_linear / mmvid_gemm_batched_f32.
// ================================================================================================

// y = act(z) elementwise (kept separate from the GEMM in training so z is available to the backward)
__global__ void act_fwd_kernel(const float4* __restrict__ z, float4* __restrict__ y, long long n4, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = z[i];
    v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
    y[i] = v;
  }
}
// dz = dy * act'(z);  QuickGELU: s = sigmoid(1.702 z), d = s + 1.702 z s (1 - s)
__global__ void act_bwd_kernel(const float4* __restrict__ z, const float4* __restrict__ dy, float4* __restrict__ dz,
                               long long n4, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = z[i], g = dy[i];
    float zz[4] = {a.x, a.y, a.z, a.w}, gg[4] = {g.x, g.y, g.z, g.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float k = act == MMVID_ACT_QUICKGELU ? 1.702f : 1.f;
      const float s = 1.f / (1.f + expf(-k * zz[j]));
      const float d = act == MMVID_ACT_NONE ? 1.f : s + k * zz[j] * s * (1.f - s);
      o[j] = gg[j] * d;
    }
    dz[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}
extern "C" int mmvid_act_forward(const float* z, float* y, long long n, int act, mmvid_stream_t stream) {
  MMVID_REQUIRE(n % 4 == 0, "n multiple of 4");
  if (n == 0) return MMVID_OK;
  const int blocks = (int)std::min<long long>(ceil_div<long long>(n / 4, 256), 148 * 16);
  act_fwd_kernel<<<blocks, 256, 0, to_stream(stream)>>>((const float4*)z, (float4*)y, n / 4, act);
  return check_launch("act_forward");
}
extern "C" int mmvid_act_backward(const float* z, const float* dy, float* dz, long long n, int act,
                                  mmvid_stream_t stream) {
  MMVID_REQUIRE(n % 4 == 0, "n multiple of 4");
  if (n == 0) return MMVID_OK;
  const int blocks = (int)std::min<long long>(ceil_div<long long>(n / 4, 256), 148 * 16);
  act_bwd_kernel<<<blocks, 256, 0, to_stream(stream)>>>((const float4*)z, (const float4*)dy, (float4*)dz, n / 4, act);
  return check_launch("act_backward");
}

// column sums: out[c] = sum_r x[r, c]   (bias gradients).  Two-stage, deterministic.
__global__ void colsum_partial_kernel(const float* __restrict__ x, float* __restrict__ part, long long rows, int cols,
                                      int rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += x[r * cols + c];
  part[(long long)blockIdx.y * cols + c] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, int cols,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(long long)p * cols + c];
  out[c] = accumulate ? out[c] + s : s;
}
extern "C" int mmvid_colsum(const float* x, float* out, float* scratch /* >= 64*cols floats */, long long rows, int cols,
                            int accumulate, mmvid_stream_t stream) {
  if (cols == 0) return MMVID_OK;
  const int nparts = (int)std::max<long long>(1, std::min<long long>(64, rows / 64));
  const int rpb = (int)ceil_div<long long>(rows, nparts);
  cudaStream_t st = to_stream(stream);
  colsum_partial_kernel<<<dim3(ceil_div(cols, 128), nparts), 128, 0, st>>>(x, scratch, rows, cols, rpb);
  int rc = check_launch("colsum_partial");
  if (rc) return rc;
  colsum_final_kernel<<<ceil_div(cols, 128), 128, 0, st>>>(scratch, out, nparts, cols, accumulate);
  return check_launch("colsum_final");
}

// LayerNorm backward (one warp per row): dx, and per-block partial dgamma/dbeta reduced by mmvid_colsum-like stage
template <int VEC_PER_LANE>
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dgb_rows,
                                     long long rows, int D, float eps) {
  // dgb_rows: [rows, 2*D] would be too large; instead each row writes xhat*dy and dy into dx-sized scratch handled by caller
  const int wpb = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  const float4* gr = reinterpret_cast<const float4*>(dy + row * D);
  float4 v[VEC_PER_LANE], g[VEC_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) { v[i] = xr[c]; g[i] = gr[c]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    else { v[i] = make_float4(0, 0, 0, 0); g[i] = v[i]; }
  }
  const float mean = warp_sum(s) / (float)D;
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) { float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean; qv += (a * a + b * b) + (cc * cc + d * d); }
  }
  const float rstd = rsqrtf(warp_sum(qv) / (float)D + eps);
  // s1 = sum(g*gamma), s2 = sum(g*gamma*xhat)
  float s1 = 0.f, s2 = 0.f;
  const float4* gm = reinterpret_cast<const float4*>(gamma);
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 w = __ldg(gm + c);
      const float xh[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
      const float gg[4] = {g[i].x * w.x, g[i].y * w.y, g[i].z * w.z, g[i].w * w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { s1 += gg[j]; s2 += gg[j] * xh[j]; }
    }
  }
  s1 = warp_sum(s1) / (float)D; s2 = warp_sum(s2) / (float)D;
  float4* dxr = reinterpret_cast<float4*>(dx + row * D);
  float4* dgr = reinterpret_cast<float4*>(dgb_rows + row * D);  // xhat * dy, summed over rows by the caller -> dgamma
#pragma unroll
  for (int i = 0; i < VEC_PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 w = __ldg(gm + c);
      const float xh[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
      const float gy[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
      const float ww[4] = {w.x, w.y, w.z, w.w};
      float o[4], t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { o[j] = rstd * (gy[j] * ww[j] - s1 - xh[j] * s2); t[j] = gy[j] * xh[j]; }
      dxr[c] = make_float4(o[0], o[1], o[2], o[3]);
      dgr[c] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
}
extern "C" int mmvid_layernorm_backward(const float* x, const float* gamma, const float* dy, float* dx,
                                        float* xhat_dy /* [rows, D] scratch: dgamma = colsum(xhat_dy), dbeta = colsum(dy) */,
                                        long long rows, int D, float eps, mmvid_stream_t stream) {
  MMVID_REQUIRE(D % 4 == 0 && D <= 1024, "D multiple of 4, <= 1024");
  if (rows == 0) return MMVID_OK;
  layernorm_bwd_kernel<8><<<(unsigned)ceil_div<long long>(rows, 8), 256, 0, to_stream(stream)>>>(x, gamma, dy, dx, xhat_dy,
                                                                                                rows, D, eps);
  return check_launch("layernorm_backward");
}

// softmax backward in place: ds[r, c] = p[r, c] * (dp[r, c] - sum_c' dp[r, c'] p[r, c']) * scale
__global__ void softmax_bwd_kernel(const float* __restrict__ p, float* __restrict__ dp, int cols, long long ld, float scale) {
  __shared__ float red[32];
  const long long r = blockIdx.x;
  const float* pr = p + r * ld;
  float* dr = dp + r * ld;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s += pr[c] * dr[c];
  s = block_sum(s, red);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dr[c] = pr[c] * (dr[c] - s) * scale;
}
extern "C" int mmvid_softmax_backward(const float* p, float* dp, long long rows, int cols, long long ld, float scale,
                                      mmvid_stream_t stream) {
  if (rows == 0) return MMVID_OK;
  softmax_bwd_kernel<<<(unsigned)rows, 256, 0, to_stream(stream)>>>(p, dp, cols, ld, scale);
  return check_launch("softmax_backward");
}

// cross entropy over selected rows (F.cross_entropy(logits[~mask1], target[~mask1]), dalle_bert.py:1040):
//   loss_sum += lse(row) - row[target];  dlogits[row] = (softmax(row) - onehot) * sel[row]   (scaled by caller via grad_scale)
__global__ void ce_fwd_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                  const uint8_t* __restrict__ sel, float* __restrict__ dlogits,
                                  float* __restrict__ loss_rows, int n) {
  __shared__ float red[32];
  const long long r = blockIdx.x;
  const float* l = logits + r * n;
  float* d = dlogits + r * n;
  if (!sel[r]) {
    for (int c = threadIdx.x; c < n; c += blockDim.x) d[c] = 0.f;
    if (threadIdx.x == 0) loss_rows[r] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int c = threadIdx.x; c < n; c += blockDim.x) m = fmaxf(m, l[c]);
  m = block_max(m, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < n; c += blockDim.x) s += expf(l[c] - m);
  s = block_sum(s, red);
  const float lse = m + logf(s);
  const int t = (int)target[r];
  for (int c = threadIdx.x; c < n; c += blockDim.x) d[c] = expf(l[c] - lse) - (c == t ? 1.f : 0.f);
  if (threadIdx.x == 0) loss_rows[r] = lse - l[t];
}
extern "C" int mmvid_cross_entropy(const float* logits, const int64_t* target, const uint8_t* sel, float* dlogits,
                                   float* loss_rows, long long rows, int n, mmvid_stream_t stream) {
  if (rows == 0) return MMVID_OK;
  ce_fwd_bwd_kernel<<<(unsigned)rows, 256, 0, to_stream(stream)>>>(logits, target, sel, dlogits, loss_rows, n);
  return check_launch("cross_entropy");
}

// embedding backward: d_table[id] += dx[b, seq_off+i]; d_table2[id] += ...; d_pos[i] += ...   (atomics: training only)
__global__ void embed_bwd_kernel(const float* __restrict__ dx, int S, int D, mmvid_embed_segment g, float* d_table,
                                 float* d_table2, float* d_pos) {
  const int r = blockIdx.x, b = blockIdx.y;
  long long id = g.ids[(long long)b * g.ids_bstride + r];
  if (g.use_pad && id == g.pad_value) id = g.pad_base + r;
  if (g.table_rows > 0 && (id < 0 || id >= g.table_rows)) __trap();
  const float* src = dx + ((long long)b * S + g.seq_off + r) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = src[c];
    if (d_table) atomicAdd(d_table + id * (long long)D + c, v);
    if (d_table2) atomicAdd(d_table2 + id * (long long)D + c, v);
    if (d_pos) atomicAdd(d_pos + (long long)r * D + c, v);
  }
}
extern "C" int mmvid_embed_backward(const float* dx, int B, int S, int D, const mmvid_embed_segment* seg, float* d_table,
                                    float* d_table2, float* d_pos, mmvid_stream_t stream) {
  if (B == 0 || seg->n == 0) return MMVID_OK;
  embed_bwd_kernel<<<dim3(seg->n, B), 192, 0, to_stream(stream)>>>(dx, S, D, *seg, d_table, d_table2, d_pos);
  return check_launch("embed_backward");
}

// 2-D transpose [R, C] -> [C, R] (weight / activation transposes feeding the backward GEMMs)
extern "C" int mmvid_transpose2d(const float* in, float* out, int R, int Cn, mmvid_stream_t stream) {
  dim3 grid(ceil_div(Cn, 32), ceil_div(R, 32), 1);
  transpose_kernel<<<grid, dim3(32, 8), 0, to_stream(stream)>>>(in, out, R, Cn);
  return check_launch("transpose2d");
}
