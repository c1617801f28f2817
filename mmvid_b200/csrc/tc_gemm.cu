// tcgen05 GEMM for every Linear / 1x1 conv of the hot path (MMVID_TF32 and MMVID_BF16 precision):
//     C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)
// Warp-specialised, one 128 x BN output tile per CTA, 2 CTAs co-resident per SM so one CTA's epilogue
// overlaps the other's main loop:
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> 3-stage smem ring, mbarrier full/empty
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma (kind::tf32 | kind::f16), accumulator in TMEM
//   warps 2..5  epilogue       tcgen05.ld TMEM -> registers -> padded smem -> fully coalesced global stores with
//                              bias / QuickGELU / residual fused (fp32 or bf16 output)
// Both operands are K-major (activations [M,K] row-major, nn.Linear weights [N,K] row-major), so a single
// descriptor flavour is needed.  fp32 operands are loaded with the TFLOAT32 tensor-map type (round-to-nearest
// to tf32 inside the TMA unit); M/N/K tails rely on TMA zero fill, stores are predicated.
#include "common.cuh"
#include "tc_common.cuh"

#include <mutex>

using namespace mmvid;
using namespace mmvid::tc;

namespace mmvid {
namespace tc {

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(MMVID_ECUDA, "cuTensorMapEncodeTiled entry point unavailable%s");
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapDataType dt = dtype == MMVID_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMVID_ECUDA, "cuTensorMapEncodeTiled failed (%s) code %lld", "", (long long)r);
  return MMVID_OK;
}

}  // namespace tc
}  // namespace mmvid

namespace {

constexpr int BM = 128;
constexpr int STAGES = 3;
constexpr int GEMM_THREADS = 192;

struct EpiArgs {
  const float* bias;
  const float* residual;
  long long ldr;
  void* C;
  long long ldc;
  int c_bf16;
  long long M;
  int N, K, act;
  // implicit-GEMM conv (CONV=true): 128-pixel M tile = box {BW, BH, BNI} of the NHWC input, K = (tap, channel block)
  int cCin, cKW, cPadT, cPadL, cBW, cBH, cBNI, cW, cH;
};

template <int BN>
constexpr size_t gemm_smem_bytes() {
  return (size_t)STAGES * (BM * 128 + BN * 128) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <bool TF32, int BN, bool CONV>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, EpiArgs e) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [barriers 256 B][pad to 1024][stages: A | B]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int BKE = TF32 ? 32 : 64;  // elements per 128-byte k-block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_k = (e.K + BKE - 1) / BKE;
  // conv: decompose the tile's first pixel into (image, row, col); tiles never straddle rows partially because
  // BW = min(W,128), BH = min(H, 128/BW), BNI = 128/(BW*BH) and W, H are powers of two
  int cx0 = 0, cy0 = 0, cn0 = 0, cblocks = 1;
  if constexpr (CONV) {
    const long long pix = (long long)m0;
    cx0 = (int)(pix % e.cW);
    cy0 = (int)((pix / e.cW) % e.cH);
    cn0 = (int)(pix / ((long long)e.cW * e.cH));
    cblocks = e.cCin / BKE;
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* a = tiles + s * STAGE_BYTES;
        if constexpr (CONV) {
          const int tap = kb / cblocks, cb = kb - tap * cblocks;
          const int ky = tap / e.cKW, kx = tap - ky * e.cKW;
          // halo taps use negative / past-the-edge coordinates: TMA zero-fills, which IS the conv zero padding
          tma_load_4d(a, &tmA, &full[s], cb * BKE, cx0 + kx - e.cPadL, cy0 + ky - e.cPadT, cn0);
        } else {
          tma_load_2d(a, &tmA, &full[s], kb * BKE, m0);
        }
        tma_load_2d(a + A_BYTES, &tmB, &full[s], kb * BKE, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc<TF32>(BM, BN);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * STAGE_BYTES);
        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // 4 x 32 bytes = UMMA_K (8 tf32 | 16 bf16) per instruction
          mma_ss<TF32>(tmem_base, desc_advance(a_desc, kk * 32), desc_advance(b_desc, kk * 32), idesc,
                       (kb | kk) != 0 ? 1u : 0u);
        tc_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
      }
      tc_commit(tmem_full);    // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quarter = warp % 4
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    constexpr int LDS = BN + 4;
    float* stage = reinterpret_cast<float*>(tiles);  // pipeline buffers are dead now
    float* my_rows = stage + (size_t)(q * 32) * LDS;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(my_rows + (size_t)lane * LDS + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
    }
    __syncwarp();
    const bool vec_ok = (e.N % 4 == 0) && (e.ldc % 4 == 0) && (!e.residual || e.ldr % 4 == 0);
#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
      const long long m = (long long)m0 + q * 32 + r;
      if (m >= e.M) break;
      const float* srow = my_rows + (size_t)r * LDS;
#pragma unroll
      for (int cc = 0; cc < BN / 128 + (BN % 128 != 0); ++cc) {
        const int col = cc * 128 + lane * 4;
        if (col >= BN) continue;
        const int n = n0 + col;
        if (n >= e.N) continue;
        float4 v = *reinterpret_cast<const float4*>(srow + col);
        float o[4] = {v.x, v.y, v.z, v.w};
        if (vec_ok) {
          if (e.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
            o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = apply_act(o[j], e.act);
          if (e.residual) {
            const float4 rr = *reinterpret_cast<const float4*>(e.residual + m * e.ldr + n);
            o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
          }
          if (e.c_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.C) + m * e.ldc + n) = pk;
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.C) + m * e.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j >= e.N) break;
            float x = o[j];
            if (e.bias) x += e.bias[n + j];
            x = apply_act(x, e.act);
            if (e.residual) x += e.residual[m * e.ldr + n + j];
            if (e.c_bf16) reinterpret_cast<__nv_bfloat16*>(e.C)[m * e.ldc + n + j] = __float2bfloat16_rn(x);
            else reinterpret_cast<float*>(e.C)[m * e.ldc + n + j] = x;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <bool TF32, int BN, bool CONV = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiArgs& e, cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t smem = gemm_smem_bytes<BN>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tc_kernel<TF32, BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div<long long>(e.M, BM), (unsigned)ceil_div(e.N, BN));
  gemm_tc_kernel<TF32, BN, CONV><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, e);
  return check_launch("gemm_tc");
}

}  // namespace

extern "C" int mmvid_linear_tc(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                               const float* bias, const float* residual, long long ldr, void* C, int c_dtype,
                               long long ldc, long long M, int N, int K, int act, int precision, cudaStream_t st) {
  const bool tf32 = precision == MMVID_TF32;
  MMVID_REQUIRE(precision == MMVID_TF32 || precision == MMVID_BF16, "precision");
  const int want = tf32 ? MMVID_DT_F32 : MMVID_DT_BF16;
  MMVID_REQUIRE(a_dtype == want && w_dtype == want, "operand dtype must match precision (fp32 for TF32, bf16 for BF16)");
  const int esz = tf32 ? 4 : 2;
  MMVID_REQUIRE((lda * esz) % 16 == 0 && (ldw * esz) % 16 == 0, "row strides must be multiples of 16 bytes");
  MMVID_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "16-byte alignment");
  if (M == 0 || N == 0) return MMVID_OK;
  const int BKE = tf32 ? 32 : 64;
  // tile width: keep >= ~1.5 waves of CTAs on 148 SMs (2 CTAs/SM) when the problem is small
  const long long tiles128 = ceil_div<long long>(M, BM) * ceil_div(N, 128);
  const int BN = (tiles128 < 200) ? 64 : 128;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * esz};
    uint32_t box[2] = {(uint32_t)BKE, (uint32_t)BM};
    int rc = make_tensor_map(&tmA, A, a_dtype, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * esz};
    uint32_t box[2] = {(uint32_t)BKE, (uint32_t)BN};
    int rc = make_tensor_map(&tmB, W, w_dtype, 2, dims, str, box);
    if (rc) return rc;
  }
  EpiArgs e{bias, residual, ldr, C, ldc, c_dtype == MMVID_DT_BF16, M, N, K, act, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (tf32) return BN == 64 ? launch<true, 64>(tmA, tmB, e, st) : launch<true, 128>(tmA, tmB, e, st);
  return BN == 64 ? launch<false, 64>(tmA, tmB, e, st) : launch<false, 128>(tmA, tmB, e, st);
}


// ------------------------------------------------------------------------------------------------
// Tensor-core conv2d (stride 1, NHWC fp32, kind::tf32): implicit GEMM whose A tiles are fetched by 4-D TMA
// boxes {32 channels, BW, BH, BNI} at tap-shifted coordinates; out-of-range halo elements are zero-filled by
// the TMA unit, so neither an im2col buffer nor a padded copy of the activation ever exists.
// Requirements: stride 1, Cin % 32 == 0, Cout % 4 == 0, H and W powers of two, NHWC in/out.
// Everything else (first 3-channel conv, stride-2 downsample, 3-channel output conv) stays on the fp32 path.
// ------------------------------------------------------------------------------------------------
extern "C" int mmvid_conv2d_tc(const mmvid_conv_params* p, cudaStream_t st) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  MMVID_REQUIRE(p->precision == MMVID_TF32, "tensor-core conv runs kind::tf32");
  MMVID_REQUIRE(p->stride == 1 && !p->in_nchw && !p->out_nchw && !p->pre_affine && !p->post_clamp && !p->upsample,
                "tc conv: stride 1, NHWC, no fused resampling");
  MMVID_REQUIRE(p->Cin % 32 == 0 && p->Cout % 4 == 0, "tc conv: Cin % 32 == 0, Cout % 4 == 0");
  MMVID_REQUIRE(pow2(p->H) && pow2(p->W) && p->Ho == p->H && p->Wo == p->W, "tc conv: power-of-two 'same' convolution");
  const int BW = p->W < 128 ? p->W : 128;
  const int BH = (128 / BW) < p->H ? (128 / BW) : p->H;
  const int BNI = 128 / (BW * BH);
  const long long M = (long long)p->N * p->H * p->W;
  const int K = p->KH * p->KW * p->Cin;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    uint64_t str[3] = {(uint64_t)p->Cin * 4, (uint64_t)p->W * p->Cin * 4, (uint64_t)p->H * p->W * p->Cin * 4};
    uint32_t box[4] = {32, (uint32_t)BW, (uint32_t)BH, (uint32_t)BNI};
    int rc = make_tensor_map(&tmA, p->in, MMVID_DT_F32, 4, dims, str, box);
    if (rc) return rc;
  }
  const long long tiles128 = ceil_div<long long>(M, BM) * ceil_div(p->Cout, 128);
  const int BN = (tiles128 < 200) ? 64 : 128;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)p->Cout};
    uint64_t str[1] = {(uint64_t)K * 4};
    uint32_t box[2] = {32, (uint32_t)BN};
    int rc = make_tensor_map(&tmB, p->w, MMVID_DT_F32, 2, dims, str, box);
    if (rc) return rc;
  }
  EpiArgs e{p->bias, p->residual, (long long)p->Cout, p->out, (long long)p->Cout, 0, M, p->Cout, K, MMVID_ACT_NONE,
            p->Cin, p->KW, p->pad_t, p->pad_l, BW, BH, BNI, p->W, p->H};
  return BN == 64 ? launch<true, 64, true>(tmA, tmB, e, st) : launch<true, 128, true>(tmA, tmB, e, st);
}
