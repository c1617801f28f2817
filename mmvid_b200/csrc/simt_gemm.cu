// CUDA-core fp32 GEMM family (MMVID_FP32 precision): the bit-faithful parity path.
//   * mmvid_linear (fp32)            C = act(A W^T + b) + residual
//   * mmvid_gemm_batched_f32         strided-batched, generic B strides (QK^T, PV, spatial attention)
//   * mmvid_conv2d (fp32)            implicit GEMM over NHWC activations: the A-tile loader gathers
//                                    (tap, channel) slices on the fly - no im2col buffer ever touches HBM;
//                                    nearest-x2 upsample, asymmetric stride-2 padding, NCHW in/out and the
//                                    VQGAN pixel pre/post-processing are folded into the loader / epilogue.
// Classic register-tiled design: block tile BM x BN x 16, 256 threads, TM x TN micro-tile per thread,
// global->register prefetch of the next k-slab while the current one is consumed from shared memory.
#include "common.cuh"

using namespace mmvid;

namespace {

enum AMode { A_DENSE = 0, A_CONV = 1 };
enum BMode { B_KCONTIG = 0, B_GENERIC = 1 };

struct GemmArgs {
  const float* A; long long lda, a_s1, a_s2;
  const float* B; long long ldb_n, ldb_k, b_s1, b_s2;
  float* C; long long ldc, c_s1, c_s2;
  const float* bias; const float* residual; long long ldr;
  long long M; int N, K; int batch2; float alpha; int act;
  // conv
  int cH, cW, cCin, cKW, cStride, cPadT, cPadL, cHo, cWo, cUps, cInNchw, cOutNchw, cPreAffine, cPostClamp;
};

constexpr int BK = 16;

template <int AM>
__device__ __forceinline__ void load_a4(const GemmArgs& g, const float* __restrict__ A, long long m, int k,
                                        bool vec_ok, float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (m >= g.M) return;
  if constexpr (AM == A_DENSE) {
    const float* p = A + m * g.lda + k;
    if (vec_ok && k + 3 < g.K) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (k + j < g.K) v[j] = p[j];
    }
  } else {
    // m -> (n, yo, xo)
    const int xo = (int)(m % g.cWo);
    const long long t1 = m / g.cWo;
    const int yo = (int)(t1 % g.cHo);
    const long long n = t1 / g.cHo;
    const int Hs = g.cUps ? 2 * g.cH : g.cH, Ws = g.cUps ? 2 * g.cW : g.cW;  // logical (upsampled) input size
    if (vec_ok && k + 3 < g.K) {
      const int tap = k / g.cCin, ci = k - tap * g.cCin;
      const int ky = tap / g.cKW, kx = tap - ky * g.cKW;
      int yi = yo * g.cStride + ky - g.cPadT, xi = xo * g.cStride + kx - g.cPadL;
      if (yi < 0 || yi >= Hs || xi < 0 || xi >= Ws) return;
      if (g.cUps) { yi >>= 1; xi >>= 1; }
      const float4 t = *reinterpret_cast<const float4*>(A + ((n * g.cH + yi) * g.cW + xi) * g.cCin + ci);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k + j;
        if (kk >= g.K) continue;
        const int tap = kk / g.cCin, ci = kk - tap * g.cCin;
        const int ky = tap / g.cKW, kx = tap - ky * g.cKW;
        int yi = yo * g.cStride + ky - g.cPadT, xi = xo * g.cStride + kx - g.cPadL;
        if (yi < 0 || yi >= Hs || xi < 0 || xi >= Ws) continue;
        if (g.cUps) { yi >>= 1; xi >>= 1; }
        float x = g.cInNchw ? A[((n * g.cCin + ci) * g.cH + yi) * g.cW + xi]
                            : A[((n * g.cH + yi) * g.cW + xi) * g.cCin + ci];
        if (g.cPreAffine) x = 2.f * x - 1.f;  // vae.py:41
        v[j] = x;
      }
    }
  }
}

template <int BMODE>
__device__ __forceinline__ void load_b4(const GemmArgs& g, const float* __restrict__ B, int n, int k, bool vec_ok,
                                        float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (n >= g.N) return;
  if constexpr (BMODE == B_KCONTIG) {
    const float* p = B + (long long)n * g.ldb_n + k;
    if (vec_ok && k + 3 < g.K) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (k + j < g.K) v[j] = p[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (k + j < g.K) v[j] = B[(long long)n * g.ldb_n + (long long)(k + j) * g.ldb_k];
  }
}

template <int BM, int BN, int TM, int TN, int AM, int BMODE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const int b1 = bz / g.batch2, b2 = bz - b1 * g.batch2;
  const float* A = g.A + b1 * g.a_s1 + b2 * g.a_s2;
  const float* B = g.B + b1 * g.b_s1 + b2 * g.b_s2;
  float* C = g.C + b1 * g.c_s1 + b2 * g.c_s2;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // loader mapping: each thread loads float4 along k.  rows per pass = 256/4 = 64.
  const int lk = (tid & 3) * 4, lr = tid >> 2;
  constexpr int A_PASSES = BM / 64, B_PASSES = BN / 64;
  bool a_vec, b_vec;
  if constexpr (AM == A_DENSE) a_vec = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  else a_vec = (g.cCin % 4 == 0) && !g.cInNchw && !g.cPreAffine && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  b_vec = (BMODE == B_KCONTIG) && (g.ldb_n % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  float ra[A_PASSES][4], rb[B_PASSES][4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i) load_a4<AM>(g, A, m0 + lr + i * 64, k0 + lk, a_vec, ra[i]);
#pragma unroll
    for (int i = 0; i < B_PASSES; ++i) load_b4<BMODE>(g, B, n0 + lr + i * 64, k0 + lk, b_vec, rb[i]);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) As[buf][lk + j][lr + i * 64] = ra[i][j];
#pragma unroll
    for (int i = 0; i < B_PASSES; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[buf][lk + j][lr + i * 64] = rb[i][j];
  };

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (g.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias) v += g.bias[n];
      v = apply_act(v, g.act);
      if constexpr (AM == A_CONV) {
        if (g.cOutNchw) {
          const int xo = (int)(m % g.cWo);
          const long long t1 = m / g.cWo;
          const int yo = (int)(t1 % g.cHo);
          const long long nb = t1 / g.cHo;
          if (g.cPostClamp) v = (fminf(fmaxf(v, -1.f), 1.f) + 1.f) * 0.5f;  // vae.py:55
          C[((nb * g.N + n) * g.cHo + yo) * g.cWo + xo] = v;
          continue;
        }
      }
      if (g.residual) v += g.residual[m * g.ldr + n];
      C[m * g.ldc + n] = v;
    }
  }
}

template <int AM, int BMODE>
int launch_gemm(const GemmArgs& g, int batch, cudaStream_t st) {
  if (g.M == 0 || g.N == 0 || batch == 0) return MMVID_OK;
  const long long tiles_big = ceil_div<long long>(g.M, 128) * ceil_div(g.N, 128) * batch;
  if (g.N >= 96 && tiles_big >= 148) {
    dim3 grid((unsigned)ceil_div<long long>(g.M, 128), ceil_div(g.N, 128), batch);
    gemm_simt_kernel<128, 128, 8, 8, AM, BMODE><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid((unsigned)ceil_div<long long>(g.M, 64), ceil_div(g.N, 64), batch);
    gemm_simt_kernel<64, 64, 4, 4, AM, BMODE><<<grid, 256, 0, st>>>(g);
  }
  return check_launch("gemm_simt");
}

}  // namespace

extern "C" int mmvid_linear_tc(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                               const float* bias, const float* residual, long long ldr, void* C, int c_dtype,
                               long long ldc, long long M, int N, int K, int act, int precision, cudaStream_t st);

extern "C" int mmvid_linear(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                            const float* bias, const float* residual, long long ldr, void* C, int c_dtype,
                            long long ldc, long long M, int N, int K, int act, int precision, mmvid_stream_t stream) {
  if (precision != MMVID_FP32)
    return mmvid_linear_tc(A, a_dtype, lda, W, w_dtype, ldw, bias, residual, ldr, C, c_dtype, ldc, M, N, K, act,
                           precision, to_stream(stream));
  MMVID_REQUIRE(a_dtype == MMVID_DT_F32 && w_dtype == MMVID_DT_F32 && c_dtype == MMVID_DT_F32, "fp32 path is fp32 only");
  GemmArgs g{};
  g.A = (const float*)A; g.lda = lda;
  g.B = (const float*)W; g.ldb_n = ldw; g.ldb_k = 1;
  g.C = (float*)C; g.ldc = ldc;
  g.bias = bias; g.residual = residual; g.ldr = ldr;
  g.M = M; g.N = N; g.K = K; g.batch2 = 1; g.alpha = 1.f; g.act = act;
  return launch_gemm<A_DENSE, B_KCONTIG>(g, 1, to_stream(stream));
}

extern "C" int mmvid_gemm_batched_f32(const float* A, long long lda, long long a_s1, long long a_s2, const float* B,
                                      long long ldb_n, long long ldb_k, long long b_s1, long long b_s2, float* C,
                                      long long ldc, long long c_s1, long long c_s2, int M, int N, int K, int batch1,
                                      int batch2, float alpha, mmvid_stream_t stream) {
  MMVID_REQUIRE((long long)batch1 * batch2 <= 65535, "batch1*batch2 <= 65535");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.a_s1 = a_s1; g.a_s2 = a_s2;
  g.B = B; g.ldb_n = ldb_n; g.ldb_k = ldb_k; g.b_s1 = b_s1; g.b_s2 = b_s2;
  g.C = C; g.ldc = ldc; g.c_s1 = c_s1; g.c_s2 = c_s2;
  g.M = M; g.N = N; g.K = K; g.batch2 = batch2; g.alpha = alpha; g.act = MMVID_ACT_NONE;
  if (ldb_k == 1) return launch_gemm<A_DENSE, B_KCONTIG>(g, batch1 * batch2, to_stream(stream));
  return launch_gemm<A_DENSE, B_GENERIC>(g, batch1 * batch2, to_stream(stream));
}

extern "C" int mmvid_conv2d_tc(const mmvid_conv_params* p, cudaStream_t st);

extern "C" int mmvid_conv2d(const mmvid_conv_params* p, mmvid_stream_t stream) {
  MMVID_REQUIRE(p->stride == 1 || p->stride == 2, "stride 1 or 2");
  MMVID_REQUIRE(!(p->out_nchw && p->residual), "residual unsupported with NCHW output");
  if (p->precision != MMVID_FP32) return mmvid_conv2d_tc(p, to_stream(stream));
  GemmArgs g{};
  g.A = static_cast<const float*>(p->in); g.B = static_cast<const float*>(p->w); g.ldb_n = (long long)p->KH * p->KW * p->Cin; g.ldb_k = 1;
  g.C = p->out; g.ldc = p->Cout; g.bias = p->bias; g.residual = p->residual; g.ldr = p->Cout;
  g.M = (long long)p->N * p->Ho * p->Wo; g.N = p->Cout; g.K = p->KH * p->KW * p->Cin;
  g.batch2 = 1; g.alpha = 1.f; g.act = MMVID_ACT_NONE;
  g.cH = p->H; g.cW = p->W; g.cCin = p->Cin; g.cKW = p->KW; g.cStride = p->stride; g.cPadT = p->pad_t;
  g.cPadL = p->pad_l; g.cHo = p->Ho; g.cWo = p->Wo; g.cUps = p->upsample; g.cInNchw = p->in_nchw;
  g.cOutNchw = p->out_nchw; g.cPreAffine = p->pre_affine; g.cPostClamp = p->post_clamp;
  return launch_gemm<A_CONV, B_KCONTIG>(g, 1, to_stream(stream));
}
