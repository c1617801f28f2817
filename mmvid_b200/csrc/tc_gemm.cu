// tcgen05 GEMM for every Linear / 1x1 conv of the hot path (MMVID_TF32, MMVID_BF16 and MMVID_F16 precision):
//     C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)
// Persistent and warp-specialised: one CTA per SM walks 128 x BN output tiles (BN = 256 / 128 / 64):
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> 4..8-stage smem ring, mbarrier full/empty
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma (kind::tf32 | kind::f16) into one of TWO TMEM
//                              accumulators, so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..9  epilogue       tcgen05.ld TMEM -> registers -> per-warp smem transpose -> 128-byte coalesced global
//                              stores with bias / QuickGELU / residual fused (fp32 or bf16 output)
// Both operands are K-major (activations [M,K] row-major, nn.Linear weights [N,K] row-major), so a single
// descriptor flavour is needed.  fp32 operands are loaded with the TFLOAT32 tensor-map type (round-to-nearest
// to tf32 inside the TMA unit); M/N/K tails rely on TMA zero fill, stores are predicated.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

#include <mutex>
#include <stdlib.h>

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
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides, bool swizzle128) {
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
  // loads of fp32 operands use TFLOAT32 (round-to-nearest in the TMA unit); DT_F32_EXACT is for stores of fp32 results
  const CUtensorMapDataType dt = dtype == MMVID_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : dtype == MMVID_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                 : (dtype == DT_F32_EXACT ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32);
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMVID_ECUDA, "cuTensorMapEncodeTiled failed (%s) code %lld", "", (long long)r);
  return MMVID_OK;
}

}  // namespace tc
}  // namespace mmvid

namespace {

constexpr int BM = 128;
constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int EPI_WARPS = 8;
constexpr int EPI_LD = 32;  // floats per staged row; 16-byte units are XOR-swizzled with the row index (no padding)

struct EpiArgs {
  const float* bias;
  const float* residual;
  long long ldr;
  void* C;
  long long ldc;
  int c_h16;   // result type: 0 = fp32, 1 = bf16, 2 = fp16
  int op_f16;  // 16-bit kinds: operands are fp16 (1) or bf16 (0)
  long long M;
  int N, K, act;
  // implicit-GEMM conv (CONV=true): 128-pixel M tile = box {BW, BH, BNI} of the NHWC input, K = (tap, channel block)
  int cCin, cKW, cPadT, cPadL, cBW, cBH, cBNI, cW, cH;
  int num_m_tiles, num_n_tiles;
  // QKV mode (qkv_q != nullptr): the [M, 3*H*64] result is scattered straight into the attention layout
  //   Q,K -> [B,H,S_pad,64]   V -> V^T [B,H,64,S_pad]   (fp32 or 16-bit via c_h16); bias applied, no act/residual
  void* qkv_q; void* qkv_k; void* qkv_vt; int qS, qSpad, qH;
  // GroupNorm statistics of the RESULT, fused (conv path, fp32 TMA-store epilogue): every epilogue warp writes the sum and
  // the sum of squared deviations from its own mean of its 32 pixels x 32 channels per group to
  // gn_partial[((m_tile * 4 + q) * gn_G + group) * 2 + {0, 1}]
  // (each entry is written by exactly one warp: deterministic), mmvid_groupnorm_from_partials reduces them per image.
  float* gn_partial; int gn_cpg, gn_G;
  int tma_store;  // 1: result tiles leave through TMA bulk stores (tmC; tmC1 for lone 16-bit chunks), see the epilogue
  unsigned long long* trace;  // debug timeline of CTA 0 (mmvid_debug_gemm_trace), normally null
  int spin;    // 1: the TMA / MMA threads poll their ring barriers (mbar_wait_spin) instead of suspending in try_wait
  int raster;  // 0: m fastest; 1: n fastest; 2: 8-wide n groups (each wave covers ~8 weight tiles x ~18 row tiles)
};

// per-group partial statistics of one 32-pixel x 32-channel result chunk (lane = pixel, o = its 32 channels), see
// EpiArgs::gn_partial: (sum, sum of squared deviations from the SLAB's own mean) - the finaliser combines the slabs with the
// parallel-variance formula, so a group whose mean is large against its spread loses nothing to cancellation.
template <int CPG>
__device__ __forceinline__ void gn_chunk_partials(const float (&o)[32], float* dst, int lane) {
  constexpr int NG = 32 / CPG;
  constexpr float n = 32.f * CPG;
  // ONE butterfly: sums of (x - K) and (x - K)^2 around a shift K taken from the slab itself (lane 0's first channel of the
  // group), so that M2 = sum (x-K)^2 - (sum (x-K))^2 / n only cancels against the slab's own spread
  float kk[NG], s[NG], q2[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    kk[g] = __shfl_sync(0xffffffffu, o[g * CPG], 0);
    float a = 0.f, b2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const float d = o[g * CPG + c] - kk[g];
      a += d;
      b2 = fmaf(d, d, b2);
    }
    s[g] = a; q2[g] = b2;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      s[g] += __shfl_xor_sync(0xffffffffu, s[g], off);
      q2[g] += __shfl_xor_sync(0xffffffffu, q2[g], off);
    }
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < NG; ++g)
      *reinterpret_cast<float2*>(dst + 2 * g) = make_float2(fmaf(n, kk[g], s[g]), fmaxf(q2[g] - s[g] * s[g] * (1.f / n), 0.f));
  }
}

// tile index -> (m tile, n tile)
__device__ __forceinline__ void tile_coords(const EpiArgs& e, int tile, int& mt, int& nt) {
  if (e.raster == 0) { mt = tile % e.num_m_tiles; nt = tile / e.num_m_tiles; }
  else if (e.raster == 1) { nt = tile % e.num_n_tiles; mt = tile / e.num_n_tiles; }
  else {
    const int GW = 8;
    const int group_tiles = GW * e.num_m_tiles;
    const int g = tile / group_tiles, r = tile - g * group_tiles;
    const int n_first = g * GW;
    const int gw = min(GW, e.num_n_tiles - n_first);
    nt = n_first + r % gw;
    mt = r / gw;
  }
}

// CTA 0 stamps clock64() at pipeline events of its first 8 tiles when a trace buffer is installed (scripts/gemm_trace.py):
// [t*64 + 0] MMA: accumulator free, [+1] first k-block landed, [+2] last k-block landed, [+3] tile committed,
// [+8] epilogue warp 2: accumulator complete, [+9] accumulator in registers, [+10] tile stored,
// [+16] TMA: first k-block issued, [+17] last k-block issued, [+24 + kb] MMA: k-block kb landed (kb < 32)
__device__ __forceinline__ void gemm_stamp(const EpiArgs& e, uint32_t tile_iter, int idx) {
  if (e.trace != nullptr && blockIdx.x == 0 && tile_iter < 8) e.trace[tile_iter * 64 + idx] = clock64();
}

template <int BN>
constexpr int gemm_stages() { return BN == 256 ? 4 : (BN == 128 ? 6 : 8); }

template <int BN>
constexpr size_t gemm_smem_bytes() {
  return (size_t)gemm_stages<BN>() * (BM * 128 + BN * 128) + EPI_WARPS * 32 * EPI_LD * 4 + 1024 /*align slack*/ + 256 /*barriers*/;
}

// Persistent, warp-specialised tcgen05 GEMM.  grid = min(#tiles, #SMs); every CTA walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... (m fastest, so concurrently running CTAs share the same weight
// tile in L2).  Three pipelines: smem ring (TMA <-> MMA), two TMEM accumulators (MMA <-> epilogue), tile loop.
// H16: the TMA-store epilogue writes 16-bit results (compile-time so that the fp32 epilogue keeps its register budget)
template <bool TF32, int BN, bool CONV, bool SWAP = false, bool H16 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmC,
                                                                 const __grid_constant__ CUtensorMap tmC1, EpiArgs e) {
  constexpr int STAGES = gemm_stages<BN>();
  extern __shared__ uint8_t smem_raw[];
  // carve: [barriers 256 B][pad to 1024][stages: A | B][epilogue staging]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int BKE = TF32 ? 32 : 64;  // elements per 128-byte k-block
  float* epi_stage = reinterpret_cast<float*>(tiles + (size_t)STAGES * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (e.K + BKE - 1) / BKE;
  const int total_tiles = e.num_m_tiles * e.num_n_tiles;
  const int cblocks = CONV ? e.cCin / BKE : 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  chain_release();
  chain_wait();  // operands / residual come from the previous kernel; the setup above overlapped its tail

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0, tma_tile = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tma_tile) {
        int mt, nt;
        tile_coords(e, tile, mt, nt);
        const int m0 = mt * BM, n0 = nt * BN;
        int cx0 = 0, cy0 = 0, cn0 = 0;
        if constexpr (CONV) {
          // tiles never straddle rows partially: BW = min(W,128), BH = min(H,128/BW), BNI = 128/(BW*BH), W,H powers of 2
          // (SWAP: the pixels are the N side of the tile, BN = 256 of them per box)
          const int p0 = SWAP ? n0 : m0;
          cx0 = p0 % e.cW;
          cy0 = (p0 / e.cW) % e.cH;
          cn0 = p0 / (e.cW * e.cH);
        }
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (e.spin) mbar_wait_spin(&empty[s], ph ^ 1); else mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* a = tiles + s * STAGE_BYTES;
          if constexpr (CONV && SWAP) {
            // transposed tile: A slot = 128 output channels of the packed weights, B slot = 256 pixels through the 4-D box
            const int tap = kb / cblocks, cb = kb - tap * cblocks;
            const int ky = tap / e.cKW, kx = tap - ky * e.cKW;
            tma_load_2d(a, &tmB, &full[s], kb * BKE, m0);
            tma_load_4d(a + A_BYTES, &tmA, &full[s], cb * BKE, cx0 + kx - e.cPadL, cy0 + ky - e.cPadT, cn0);
          } else {
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
          if (kb == 0) gemm_stamp(e, tma_tile, 16);
          if (kb == num_k - 1) gemm_stamp(e, tma_tile, 17);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = TF32 ? make_idesc<true>(BM, BN) : make_idesc_h16(e.op_f16 != 0, BM, BN);
      uint32_t it = 0, tile_iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
        const uint32_t acc = tile_iter & 1, acc_ph = (tile_iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        gemm_stamp(e, tile_iter, 0);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (e.spin) mbar_wait_spin(&full[s], ph); else mbar_wait(&full[s], ph);
          tc_fence_after();
          if (e.trace != nullptr) {
            if (kb == 0) gemm_stamp(e, tile_iter, 1);
            if (kb == num_k - 1) gemm_stamp(e, tile_iter, 2);
            if (kb < 32) gemm_stamp(e, tile_iter, 24 + kb);
          }
          const uint32_t a_addr = smem_u32(tiles + s * STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // 4 x 32 bytes = UMMA_K (8 tf32 | 16 bf16) per instruction
            mma_ss<TF32>(d_tmem, desc_advance(a_desc, kk * 32), desc_advance(b_desc, kk * 32), idesc,
                         (kb | kk) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete
        gemm_stamp(e, tile_iter, 3);
      }
    }
  } else {
    // ---------------- epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4
    const int q = warp & 3;
    const int ew = warp - 2;              // 0..7
    const int chalf = ew >> 2;            // which half of the tile's 32-column chunks this warp drains
    if (e.tma_store) {
      // ---- TMA-store epilogue (fp32 result, N % 32 == 0, 16-byte aligned rows).  The r1p source-level profile showed the
      // transposing epilogue below to be INSTRUCTION bound: ~700 warp instructions per 32-column chunk (per-row
      // predicates, 64-bit address arithmetic, shared-memory round trip), ~2900 clk per chunk, which caps the kernel as
      // soon as the main loop gets faster (256-wide tiles).  Here a lane keeps its own output row: bias / activation /
      // residual are applied in registers, the 32 x 32 chunk is written once to shared memory in the SWIZZLE_128B
      // layout and one elected lane hands it to the TMA unit, which also clips the M tail.
      const uint32_t st_base = smem_u32(epi_stage) + (uint32_t)(ew * 4096);
      const uint32_t st_row = st_base + (uint32_t)(lane * 128);
      constexpr int NCH = BN / 64;
      constexpr int GRP = NCH < 2 ? NCH : 2;  // 64-wide tiles: one chunk per warp
      if (lane == 0) prefetch_tmap(&tmC);
      uint32_t tile_iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
        int mt, nt;
        tile_coords(e, tile, mt, nt);
        const int m0 = mt * BM, n0 = nt * BN + chalf * (BN / 2);
        const uint32_t acc = tile_iter & 1, acc_ph = (tile_iter >> 1) & 1;
        const long long m_row = (long long)m0 + q * 32 + lane;   // the output row this lane owns
        const bool row_ok = m_row < e.M;
        // residual of the first chunk pair is requested before the accumulator wait (latency off the critical path)
        float4 res[GRP][8];
        auto load_res = [&](int c0) {
#pragma unroll
          for (int cc = 0; cc < GRP; ++cc) {
            const int ncol = n0 + (c0 + cc) * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              res[cc][i] = (row_ok && ncol < e.N) ? *reinterpret_cast<const float4*>(e.residual + m_row * e.ldr + ncol + 4 * i)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        if (!SWAP && e.residual) load_res(0);
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
        const bool etr = (warp == 2 && lane == 0);
        if (etr) gemm_stamp(e, tile_iter, 8);
        const uint32_t t_src = tmem_base + acc * BN + chalf * (BN / 2) + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += GRP) {
          uint32_t racc[GRP][32];
#pragma unroll
          for (int cc = 0; cc < GRP; ++cc) tmem_ld32(t_src + (c0 + cc) * 32, racc[cc]);
          tmem_ld_wait();
          if (c0 + GRP >= NCH) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (etr) gemm_stamp(e, tile_iter, 9);
          }
          // 16-bit results (H16): the group's chunks are packed and staged one after the other and leave together as one
          // {64 x 32} box (128-byte rows); a lone chunk leaves as a {32 x 32} box (tc_epilogue.cuh)
          const int n_grp0 = n0 + c0 * 32;
          int nvalid = 0;
#pragma unroll
          for (int cc = 0; cc < GRP; ++cc)
            if (n_grp0 + cc * 32 < e.N) nvalid = cc + 1;
          if (H16 && nvalid > 0) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
#pragma unroll
          for (int cc = 0; cc < GRP; ++cc) {
            const int ncol = n0 + (c0 + cc) * 32;
            if (!SWAP && ncol >= e.N) continue;  // warp-uniform
            float o[32];
            if constexpr (SWAP) {
              // transposed conv tile: this lane owns OUTPUT CHANNEL m0 + q*32 + lane, the chunk's 32 columns are pixels
              // ncol .. ncol+31.  Bias is one scalar per lane, residual rows are read channel-contiguous (coalesced), and
              // the chunk is transposed on its way into shared memory so that the bulk store writes NHWC rows.
              const int cout = m0 + q * 32 + lane;
              const float bv = e.bias ? __ldg(e.bias + cout) : 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                o[i] = __uint_as_float(racc[cc][i]) + bv;
                if (e.residual) o[i] += e.residual[(long long)(ncol + i) * e.ldr + cout];
              }
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 32; ++i)  // element (pixel row i, channel lane): 16-byte unit (lane >> 2) ^ (i & 7) of row i
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_base + (uint32_t)(i * 128 + ((((lane >> 2) ^ (i & 7)) << 4) | ((lane & 3) << 2)))),
                             "f"(o[i])
                             : "memory");
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                 reinterpret_cast<uint64_t>(&tmC)),
                             "r"(st_base), "r"(m0 + q * 32), "r"(ncol)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
              continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + ncol + 4 * i));  // same address in every lane
              o[4 * i + 0] = __uint_as_float(racc[cc][4 * i + 0]) + b.x;
              o[4 * i + 1] = __uint_as_float(racc[cc][4 * i + 1]) + b.y;
              o[4 * i + 2] = __uint_as_float(racc[cc][4 * i + 2]) + b.z;
              o[4 * i + 3] = __uint_as_float(racc[cc][4 * i + 3]) + b.w;
            }
            if (e.act != MMVID_ACT_NONE) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = apply_act_fast(o[i], e.act);
            }
            if (e.residual) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                o[4 * i + 0] += res[cc][i].x; o[4 * i + 1] += res[cc][i].y;
                o[4 * i + 2] += res[cc][i].z; o[4 * i + 3] += res[cc][i].w;
              }
            }
            if constexpr (CONV && !H16) {
              if (e.gn_partial != nullptr) {  // warp-uniform; rows are always valid here (M % 128 == 0 is a precondition)
                float* dst = e.gn_partial + (((long long)mt * 4 + q) * e.gn_G + ncol / e.gn_cpg) * 2;
                if (e.gn_cpg == 4) gn_chunk_partials<4>(o, dst, lane);
                else if (e.gn_cpg == 8) gn_chunk_partials<8>(o, dst, lane);
                else gn_chunk_partials<16>(o, dst, lane);
              }
            }
            if constexpr (H16) {
              const bool f16 = e.c_h16 == 2;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t u0 = pack_h16_rt(o[8 * j + 0], o[8 * j + 1], f16), u1 = pack_h16_rt(o[8 * j + 2], o[8 * j + 3], f16);
                const uint32_t u2 = pack_h16_rt(o[8 * j + 4], o[8 * j + 5], f16), u3 = pack_h16_rt(o[8 * j + 6], o[8 * j + 7], f16);
                uint32_t addr;
                if (nvalid == 2) addr = st_row + (uint32_t)(((cc * 4 + j) ^ (lane & 7)) * 16);
                else addr = st_base + (uint32_t)(lane * 64 + j * 16);  // lone chunk: dense 64-byte rows (SWIZZLE_NONE map)
                sts128_u32(addr, u0, u1, u2, u3);
              }
              continue;
            } else {
            // the previous chunk's bulk store must have finished READING the staging buffer
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)  // 16-byte unit j of row r lives at unit j ^ (r & 7): the TMA 128-byte swizzle
              sts128(st_row + (uint32_t)((j ^ (lane & 7)) * 16), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_2d(&tmC, st_base, ncol, m0 + q * 32);
            }  // !H16
          }
          if (H16 && nvalid > 0) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_2d(nvalid == 2 ? &tmC : &tmC1, st_base, n_grp0, m0 + q * 32);
          }
          if (!SWAP && e.residual && c0 + GRP < NCH) load_res(c0 + GRP);
        }
        if (etr) gemm_stamp(e, tile_iter, 10);
      }
      if (lane == 0) tma_store_wait_all();  // all stores done before smem goes away
      __syncwarp();
    } else {
    const uint32_t st_base = smem_u32(epi_stage) + (uint32_t)(ew * 32 * EPI_LD * 4);
    const uint32_t st_wr = st_base + (uint32_t)(lane * EPI_LD * 4);                       // my row while transposing
    const int col = (lane & 7) * 4, rsub = lane >> 3;                                     // coalesced phase mapping
    // row r keeps its 16-byte unit u at physical unit (u ^ (r & 7)): conflict-free 128-bit writes (one row per lane)
    // and reads (8 lanes per row) without padding.  Reader rows are rsub + 4*i, so (row & 7) = (rsub + 4*i) & 7.
    const uint32_t st_rd_row = st_base + (uint32_t)(rsub * EPI_LD * 4);
    const bool vec_ok = (e.N % 4 == 0) && (e.ldc % 4 == 0) && (!e.residual || e.ldr % 4 == 0);
    constexpr int NCH = BN / 64;          // chunks per warp (half of the tile's BN/32)
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
      int mt, nt;
      tile_coords(e, tile, mt, nt);
      const int m0 = mt * BM, n0 = nt * BN + chalf * (BN / 2);
      const uint32_t acc = tile_iter & 1, acc_ph = (tile_iter >> 1) & 1;
      // bias for every chunk of this tile is fetched BEFORE waiting for the accumulator (latency off the critical path)
      float4 bias_r[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int n = n0 + c * 32 + col;
        bias_r[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.bias && n < e.N) {
          if (vec_ok) bias_r[c] = __ldg(reinterpret_cast<const float4*>(e.bias + n));
          else {
            bias_r[c].x = e.bias[n];
            if (n + 1 < e.N) bias_r[c].y = e.bias[n + 1];
            if (n + 2 < e.N) bias_r[c].z = e.bias[n + 2];
            if (n + 3 < e.N) bias_r[c].w = e.bias[n + 3];
          }
        }
      }
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const bool etr = (warp == 2 && lane == 0);
      if (etr) gemm_stamp(e, tile_iter, 8);
      const uint32_t t_src = tmem_base + acc * BN + chalf * (BN / 2) + ((uint32_t)(q * 32) << 16);
      const long long m_first = (long long)m0 + q * 32 + rsub;
      // this warp's accumulator columns are requested two 32-column chunks at a time and awaited once per pair (a
      // tcgen05.wait::ld per chunk exposes a few hundred cycles each; more than two chunks in flight would not fit the
      // register budget of the 256-wide tile)
      constexpr int GRP = NCH < 2 ? NCH : 2;  // 64-wide tiles: one chunk per warp
#pragma unroll
      for (int c0 = 0; c0 < NCH; c0 += GRP) {
      uint32_t racc[GRP][32];
#pragma unroll
      for (int cc = 0; cc < GRP; ++cc) tmem_ld32(t_src + (c0 + cc) * 32, racc[cc]);
      tmem_ld_wait();
      if (c0 + GRP >= NCH) {
        // the whole accumulator is in registers: hand the TMEM buffer back to the MMA warp right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (etr) gemm_stamp(e, tile_iter, 9);
      }
#pragma unroll
      for (int cc = 0; cc < GRP; ++cc) {
        const int c = c0 + cc;
        if (n0 + c * 32 >= e.N) continue;
        uint32_t (&r)[32] = racc[cc];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(st_wr + (uint32_t)(((j ^ (lane & 7)) * 16)), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int n = n0 + c * 32 + col;
        const bool n_ok = n < e.N;
        if (e.qkv_q != nullptr) {
          const int D = e.qH * 64;
          const int nc = n0 + c * 32;               // first column of this chunk: never straddles a head (64-aligned)
          const int which = nc / D, hh = (nc - which * D) >> 6, d0 = nc & 63;
          if (which < 2) {
            // Q / K: rows stay rows; 8 lanes x 16 B cover the chunk's 32 head-dim values of one token
            void* dst = which == 0 ? e.qkv_q : e.qkv_k;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const long long m = m_first + i * 4;
              if (m >= e.M) continue;
              const float4 v = lds128(st_rd_row + (uint32_t)(i * 4 * EPI_LD * 4) + (uint32_t)((((lane & 7) ^ ((rsub + 4 * i) & 7)) * 16)));
              const int bb = (int)(m / e.qS), ss = (int)(m - (long long)bb * e.qS);
              const long long off = (((long long)bb * e.qH + hh) * e.qSpad + ss) * 64 + d0 + col;
              const float o0 = v.x + bias_r[c].x, o1 = v.y + bias_r[c].y, o2 = v.z + bias_r[c].z, o3 = v.w + bias_r[c].w;
              if (e.c_h16) {
                uint2 pk;
                pk.x = pack_h16_rt(o0, o1, e.c_h16 == 2);
                pk.y = pack_h16_rt(o2, o3, e.c_h16 == 2);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dst) + off) = pk;
              } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + off) = make_float4(o0, o1, o2, o3);
              }
            }
          } else {
            // V: written transposed straight from the accumulator registers (lane == token row, so for a fixed
            // head-dim index the 32 lanes write 32 consecutive tokens: one coalesced 128-byte (64-byte bf16) store)
            const long long m = (long long)m0 + q * 32 + lane;
            if (m < e.M) {
              const int bb = (int)(m / e.qS), ss = (int)(m - (long long)bb * e.qS);
              const long long base = (((long long)bb * e.qH + hh) * 64 + d0) * e.qSpad + ss;
#pragma unroll
              for (int d = 0; d < 32; ++d) {
                const float o = __uint_as_float(r[d]) + (e.bias ? __ldg(e.bias + nc + d) : 0.f);
                if (e.c_h16) reinterpret_cast<uint16_t*>(e.qkv_vt)[base + (long long)d * e.qSpad] = cvt_h16_rt(o, e.c_h16 == 2);
                else reinterpret_cast<float*>(e.qkv_vt)[base + (long long)d * e.qSpad] = o;
              }
            }
          }
        } else if (vec_ok) {
          // 8 lanes cover one 128-byte row segment, 4 rows per instruction; all loads are issued before any store
          float4 res[8], v[8];
          if (e.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const long long m = m_first + i * 4;
              res[i] = (m < e.M && n_ok) ? *reinterpret_cast<const float4*>(e.residual + m * e.ldr + n)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = lds128(st_rd_row + (uint32_t)(i * 4 * EPI_LD * 4) + (uint32_t)((((lane & 7) ^ ((rsub + 4 * i) & 7)) * 16)));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long m = m_first + i * 4;
            float o[4] = {v[i].x + bias_r[c].x, v[i].y + bias_r[c].y, v[i].z + bias_r[c].z, v[i].w + bias_r[c].w};
            if (e.act != MMVID_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = apply_act_fast(o[j], e.act);
            }
            if (e.residual) { o[0] += res[i].x; o[1] += res[i].y; o[2] += res[i].z; o[3] += res[i].w; }
            if (m < e.M && n_ok) {
              if (e.c_h16) {
                uint2 pk;
                pk.x = pack_h16_rt(o[0], o[1], e.c_h16 == 2);
                pk.y = pack_h16_rt(o[2], o[3], e.c_h16 == 2);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(e.C) + m * e.ldc + n) = pk;
              } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.C) + m * e.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
              }
            }
          }
        } else {
          // unaligned N / leading dimensions: scalar path
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const long long m = m_first + i * 4;
            if (m >= e.M || !n_ok) continue;
            const float4 v = lds128(st_rd_row + (uint32_t)(i * 4 * EPI_LD * 4) + (uint32_t)((((lane & 7) ^ ((rsub + 4 * i) & 7)) * 16)));
            const float o[4] = {v.x + bias_r[c].x, v.y + bias_r[c].y, v.z + bias_r[c].z, v.w + bias_r[c].w};
            for (int j = 0; j < 4; ++j) {
              if (n + j >= e.N) break;
              float x = apply_act(o[j], e.act);
              if (e.residual) x += e.residual[m * e.ldr + n + j];
              if (e.c_h16) reinterpret_cast<uint16_t*>(e.C)[m * e.ldc + n + j] = cvt_h16_rt(x, e.c_h16 == 2);
              else reinterpret_cast<float*>(e.C)[m * e.ldc + n + j] = x;
            }
          }
        }
        __syncwarp();
      }
      }
      if (etr) gemm_stamp(e, tile_iter, 10);
    }
    }  // !tma_store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// pick the N tile that minimises (#rounds over the SMs) x (tile width): larger tiles halve the L2->SM operand
// traffic per FLOP, smaller ones quantise better on small problems
constexpr int GEMM_SPIN_DEFAULT = 0;
constexpr int GEMM_TMA_STORE_DEFAULT = 1;
}  // namespace
namespace mmvid { unsigned long long* g_gemm_trace = nullptr; }
namespace {
using mmvid::g_gemm_trace;
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int pick_bn(long long M, int N, bool allow256 = true) {
  const int forced = env_int("MMVID_GEMM_BN", 0);  // tuning / experiments only
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  const int sms = num_sms();
  const long long mt = ceil_div<long long>(M, BM);
  int best = 64;
  double best_cost = 1e30;
  // Per-tile cost in units of 128x128-tile k-loops, from the r1q / r1r pipeline traces: the 128-wide SS MMA is operand-fetch
  // bound (380 clk per k-block instead of 256), the 256-wide one runs at the nominal 512 but has a 4-stage ring, so a
  // 256-wide tile costs ~1.5x a 128-wide one for 2x the work; 64-wide tiles pay the same operand traffic as 128-wide ones.
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn == 256 && (N < 256 || !allow256)) continue;
    const long long tiles = mt * ceil_div(N, bn);
    const double rounds = (double)ceil_div<long long>(tiles, sms);
    const double per_tile = bn == 256 ? 212.0 : (bn == 128 ? 140.0 : 88.0);
    const double cost = rounds * per_tile;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <bool TF32, int BN, bool CONV, bool SWAP, bool H16>
int launch_k(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC1, const EpiArgs& e,
             int grid, cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t smem = gemm_smem_bytes<BN>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tc_kernel<TF32, BN, CONV, SWAP, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  cudaError_t err = launch_chained(gemm_tc_kernel<TF32, BN, CONV, SWAP, H16>, dim3(grid), dim3(GEMM_THREADS), smem, st, tmA, tmB, tmC,
                                   tmC1, e);
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "gemm_tc launch: %s", cudaGetErrorString(err));
  return check_launch("gemm_tc");
}

template <bool TF32, int BN, bool CONV = false, bool SWAP = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, EpiArgs e, cudaStream_t st) {
  e.num_m_tiles = (int)ceil_div<long long>(e.M, BM);
  e.num_n_tiles = ceil_div(e.N, BN);
  if (SWAP) {  // transposed conv tile: "m" tiles walk the output channels (128 each), "n" tiles the pixels (BN each)
    e.num_m_tiles = e.N / BM;
    e.num_n_tiles = (int)(e.M / BN);
  }
  e.spin = env_int("MMVID_GEMM_SPIN", GEMM_SPIN_DEFAULT);
  e.trace = g_gemm_trace;
  e.raster = env_int("MMVID_GEMM_RASTER", 1);  // n fastest: consecutive CTAs share the activation tile (measured best)
  // result tiles leave through TMA bulk stores when the output is plain row-major fp32 with 32-column granularity
  CUtensorMap tmC = tmA, tmC1 = tmA;  // placeholders when unused
  e.tma_store = 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int csz = e.c_h16 ? 2 : 4;
  if (env_int("MMVID_GEMM_TMA_STORE", GEMM_TMA_STORE_DEFAULT) && !((SWAP || CONV || TF32) && e.c_h16) && e.qkv_q == nullptr && e.N % 32 == 0 &&
      (e.ldc * csz) % 16 == 0 && al16(e.C) && (!e.residual || (e.ldr % 4 == 0 && al16(e.residual))) && (!e.bias || al16(e.bias))) {
    uint64_t dims[2] = {(uint64_t)e.N, (uint64_t)e.M};
    uint64_t str[1] = {(uint64_t)e.ldc * csz};
    if (e.c_h16) {
      const int dt = e.c_h16 == 2 ? MMVID_DT_F16 : MMVID_DT_BF16;
      uint32_t box[2] = {64, 32}, box1[2] = {32, 32};
      int rc = make_tensor_map(&tmC, e.C, dt, 2, dims, str, box);
      if (rc) return rc;
      rc = make_tensor_map(&tmC1, e.C, dt, 2, dims, str, box1, nullptr, /*swizzle128=*/false);
      if (rc) return rc;
    } else {
      uint32_t box[2] = {32, 32};
      int rc = make_tensor_map(&tmC, e.C, DT_F32_EXACT, 2, dims, str, box);
      if (rc) return rc;
    }
    e.tma_store = 1;
  }
  const long long tiles = (long long)e.num_m_tiles * e.num_n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  if (SWAP && !e.tma_store) return fail(MMVID_EINVAL, "transposed conv tile needs the TMA-store epilogue%s");
  if constexpr (!TF32 && !CONV && !SWAP) {
    if (e.c_h16 && e.tma_store) return launch_k<TF32, BN, CONV, SWAP, true>(tmA, tmB, tmC, tmC1, e, grid, st);
  }
  return launch_k<TF32, BN, CONV, SWAP, false>(tmA, tmB, tmC, tmC1, e, grid, st);
}

template <bool TF32, bool CONV>
int launch_bn(int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiArgs& e, cudaStream_t st) {
  if (BN == 256) return launch<TF32, 256, CONV>(tmA, tmB, e, st);
  if (BN == 128) return launch<TF32, 128, CONV>(tmA, tmB, e, st);
  return launch<TF32, 64, CONV>(tmA, tmB, e, st);
}

}  // namespace

namespace {
// CTA-pair (cta_group::2, tc_gemm2.cu) tile for this problem: 0 = none (single-CTA kernel), else 128 | 192 | 256 columns.
// MMVID_GEMM_2CTA=128|192|256 forces it everywhere, =1 disables it.
int pick_pair_bn(long long M, int N, int K, bool tf32, int c_dtype) {
  int bn2 = env_int("MMVID_GEMM_2CTA", 0);
  if (bn2 == 0 && !tf32 && N >= 512 && M >= 1024) {
    // kind::f16: the MMAs take half the time for the same operand bytes, so the widest tile (fewest bytes per FLOP: the
    // L2 -> SM stream is what bounds these GEMMs) wins unless it quantises badly over the 74 CTA pairs
    const int sms = num_sms();
    const long long mt2 = ceil_div<long long>(M, 2 * BM);
    const double p128 = (double)ceil_div<long long>(mt2 * ceil_div(N, 128), sms / 2) * 128.0;
    const double p192 = (double)ceil_div<long long>(mt2 * ceil_div(N, 192), sms / 2) * 180.0;
    const double p256 = (double)ceil_div<long long>(mt2 * ceil_div(N, 256), sms / 2) * 230.0;
    bn2 = (p256 <= p192 && p256 <= p128) ? 256 : (p192 <= p128 ? 192 : 128);
  }
  if (bn2 == 0 && tf32 && c_dtype == MMVID_DT_F32 && N >= 512 && M >= 1024) {
    // An SS tcgen05.mma costs ~100 clk whatever its N (r1q-r1v traces: 380-430 clk per 4-instruction k-block for N = 128,
    // 192 and 256 alike; the 128-row A slice fetch from shared memory is the floor), so only instructions with >= 100 clk
    // of work run the tensor pipe at its rate: 256 x 192 / 256 x 256 CTA-pair tiles.  With the TMA-store epilogue these
    // tiles are no longer epilogue bound (c_fc 80 -> 72 us, c_proj 72 -> 67 us).  Per-tile costs in k clk from the traces.
    const int sms = num_sms();
    const long long mt = ceil_div<long long>(M, BM), mt2 = ceil_div<long long>(M, 2 * BM);
    const double c128 = (double)ceil_div<long long>(mt * ceil_div(N, 128), sms) * 11.2;
    const double c256 = tf32 ? (double)ceil_div<long long>(mt * ceil_div(N, 256), sms) * 17.1 : 1e30;
    const double p192 = (double)ceil_div<long long>(mt2 * ceil_div(N, 192), sms / 2) * 11.8;
    const double p256 = (double)ceil_div<long long>(mt2 * ceil_div(N, 256), sms / 2) * 13.2;
    const double best1 = c128 < c256 ? c128 : c256;
    if (p256 <= p192 && p256 < best1) bn2 = 256;
    else if (p192 < p256 && p192 < best1) bn2 = 192;
  }
  // long-K GEMMs: the 256 x 128 pair tile measured 7-9 % faster than the single-CTA tile (deeper TMA ring)
  if (bn2 == 0 && K >= 2048 && M >= 1024 && N >= 256) bn2 = 128;
  return (bn2 == 128 || bn2 == 192 || bn2 == 256) ? bn2 : 0;
}

}  // namespace

// Debug hook (host only, no launch): which tile mmvid_linear would pick for a plain (non-QKV) tensor-core GEMM:
// 2000 + BN = CTA-pair kernel with a 256 x BN tile, 1000 + BN = single-CTA kernel with a 128 x BN tile.
extern "C" int mmvid_debug_pick_tile(long long M, int N, int K, int precision, int c_dtype) {
  const bool tf32 = precision == MMVID_TF32;
  const int bn2 = pick_pair_bn(M, N, K, tf32, c_dtype);
  return bn2 ? 2000 + bn2 : 1000 + pick_bn(M, N, tf32);
}

// Debug / profiling hook: CTA 0 of every following single-CTA tensor-core GEMM launch writes clock64() stamps of its
// first 8 tiles into dev_buf (>= 512 entries; NULL switches it off).  Layout: gemm_stamp above.
extern "C" int mmvid_debug_gemm_trace(unsigned long long* dev_buf) {
  g_gemm_trace = dev_buf;
  return MMVID_OK;
}

extern "C" int mmvid_linear_tc2(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                                const float* bias, const float* residual, long long ldr, void* C, int c_dtype,
                                long long ldc, long long M, int N, int K, int act, int precision, int BN, cudaStream_t st);

extern "C" int mmvid_linear_qkv_tc2(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* q,
                                    void* k, void* vt, int B, int H, int S, int S_pad, int precision, cudaStream_t st);

namespace {
struct QkvReq { void* q; void* k; void* vt; int S, Spad, H; };
thread_local QkvReq g_qkv{nullptr, nullptr, nullptr, 0, 0, 0};
}  // namespace

extern "C" int mmvid_linear_tc(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                               const float* bias, const float* residual, long long ldr, void* C, int c_dtype,
                               long long ldc, long long M, int N, int K, int act, int precision, cudaStream_t st) {
  const bool tf32 = precision == MMVID_TF32;
  MMVID_REQUIRE(precision == MMVID_TF32 || precision == MMVID_BF16 || precision == MMVID_F16, "precision");
  if (act == MMVID_ACT_QUICKGELU && !tf32 && c_dtype != MMVID_DT_F32 && env_int("MMVID_GELU_TANH", 1))
    act = MMVID_ACT_QUICKGELU_TANH;  // 16-bit result: one MUFU op per element instead of two (see common.cuh)
  const int want = tf32 ? MMVID_DT_F32 : (precision == MMVID_F16 ? MMVID_DT_F16 : MMVID_DT_BF16);
  MMVID_REQUIRE(a_dtype == want && w_dtype == want,
                "operand dtype must match precision (fp32 for TF32, bf16 for BF16, fp16 for F16)");
  const int esz = tf32 ? 4 : 2;
  MMVID_REQUIRE((lda * esz) % 16 == 0 && (ldw * esz) % 16 == 0, "row strides must be multiples of 16 bytes");
  MMVID_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "16-byte alignment");
  if (M == 0 || N == 0) return MMVID_OK;
  const int BKE = tf32 ? 32 : 64;
  // tile width: keep >= ~1.5 waves of CTAs on 148 SMs (2 CTAs/SM) when the problem is small
  {
    // CTA-pair (cta_group::2, tc_gemm2.cu) kernel for long-K GEMMs (c_proj: K = 3072 measured 7-9 % faster than the
    // single-CTA tile; short-K GEMMs are bounded by their output stream and gain nothing).  MMVID_GEMM_2CTA=128|256
    // forces it everywhere, =1 disables it.
    const int bn2 = pick_pair_bn(M, N, K, tf32, c_dtype);
    if (bn2 != 0 && g_qkv.q == nullptr) {
      const int rc = mmvid_linear_tc2(A, a_dtype, lda, W, w_dtype, ldw, bias, residual, ldr, C, c_dtype, ldc, M, N, K, act,
                                      precision, bn2, st);
      if (rc != 1) return rc;  // 1: not applicable (result cannot leave through TMA stores) -> single-CTA kernel below
    }
  }
  const int BN = pick_bn(M, N, tf32);  // 256-wide tiles only pay in tf32 (r1s: bf16 c_fc 56 -> 68 us with them)
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
  EpiArgs e{};
  e.bias = bias; e.residual = residual; e.ldr = ldr; e.C = C; e.ldc = ldc;
  e.c_h16 = c_dtype == MMVID_DT_BF16 ? 1 : (c_dtype == MMVID_DT_F16 ? 2 : 0);
  e.op_f16 = precision == MMVID_F16;
  e.M = M; e.N = N; e.K = K; e.act = act;
  if (g_qkv.q) { e.qkv_q = g_qkv.q; e.qkv_k = g_qkv.k; e.qkv_vt = g_qkv.vt; e.qS = g_qkv.S; e.qSpad = g_qkv.Spad; e.qH = g_qkv.H; }
  return tf32 ? launch_bn<true, false>(BN, tmA, tmB, e, st) : launch_bn<false, false>(BN, tmA, tmB, e, st);
}


// ------------------------------------------------------------------------------------------------
// Tensor-core conv2d (stride 1, NHWC fp32, kind::tf32): implicit GEMM whose A tiles are fetched by 4-D TMA
// boxes {32 channels, BW, BH, BNI} at tap-shifted coordinates; out-of-range halo elements are zero-filled by
// the TMA unit, so neither an im2col buffer nor a padded copy of the activation ever exists.
// Requirements: stride 1, Cin % 32 == 0, Cout % 4 == 0, H and W powers of two, NHWC in/out.
// Everything else (first 3-channel conv, stride-2 downsample, 3-channel output conv) stays on the fp32 path.
// ------------------------------------------------------------------------------------------------
extern "C" int mmvid_conv2d_gn_fusable(const mmvid_conv_params* p) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!(p->precision == MMVID_TF32 || p->precision == MMVID_F16)) return 0;
  if (p->stride != 1 || p->in_nchw || p->out_nchw || p->pre_affine || p->post_clamp || p->upsample || p->KH != 3 || p->KW != 3) return 0;
  if (!pow2(p->H) || !pow2(p->W) || p->Ho != p->H || p->Wo != p->W || ((long long)p->H * p->W) % 128 != 0) return 0;
  if (p->gn_groups <= 0 || p->Cout % p->gn_groups != 0 || p->Cout % 32 != 0) return 0;
  const int cpg = p->Cout / p->gn_groups;
  if (cpg != 4 && cpg != 8 && cpg != 16) return 0;
  if (p->precision == MMVID_TF32 && env_int("MMVID_CONV_SWAP", 0)) return 0;
  // the TMA-store epilogue's own preconditions (launch<>): 16-byte aligned result / residual / bias
  if (!env_int("MMVID_GEMM_TMA_STORE", GEMM_TMA_STORE_DEFAULT) || !al16(p->out) || (p->residual && !al16(p->residual)) ||
      (p->bias && !al16(p->bias)))
    return 0;
  return 1;
}

extern "C" int mmvid_conv2d_tc(const mmvid_conv_params* p, cudaStream_t st) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  MMVID_REQUIRE(p->precision == MMVID_TF32 || p->precision == MMVID_F16, "tensor-core conv runs kind::tf32 or kind::f16 (fp16)");
  const bool f16 = p->precision == MMVID_F16;
  const int esz = f16 ? 2 : 4, BKE = f16 ? 64 : 32, dt = f16 ? MMVID_DT_F16 : MMVID_DT_F32;
  MMVID_REQUIRE(p->stride == 1 && !p->in_nchw && !p->out_nchw && !p->pre_affine && !p->post_clamp && !p->upsample,
                "tc conv: stride 1, NHWC, no fused resampling");
  MMVID_REQUIRE(p->Cin % BKE == 0 && p->Cout % 4 == 0, "tc conv: Cin % 32 == 0 (fp16: 64), Cout % 4 == 0");
  MMVID_REQUIRE(pow2(p->H) && pow2(p->W) && p->Ho == p->H && p->Wo == p->W, "tc conv: power-of-two 'same' convolution");
  const long long M = (long long)p->N * p->H * p->W;
  const int K = p->KH * p->KW * p->Cin;
  // Transposed tile for the layers with Cout = 128 (MMVID_CONV_SWAP, tf32 only): A = 128 output channels of the packed
  // weights, B = 256 pixels, so that the MMA is 256 wide and leaves the ~100 clk SS floor
  // (profiles/r1_g_gemm_pipeline.md); the result chunk is transposed in shared memory before the bulk store.
  if (!f16 && env_int("MMVID_CONV_SWAP", 0) && p->Cout % 128 == 0 && M % 256 == 0) {
    const int BW2 = p->W < 256 ? p->W : 256;
    const int BH2 = (256 / BW2) < p->H ? (256 / BW2) : p->H;
    const int BNI2 = 256 / (BW2 * BH2);
    if (BNI2 >= 1 && p->N % BNI2 == 0) {
      CUtensorMap tmX, tmW;
      {
        uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
        uint64_t str[3] = {(uint64_t)p->Cin * 4, (uint64_t)p->W * p->Cin * 4, (uint64_t)p->H * p->W * p->Cin * 4};
        uint32_t box[4] = {32, (uint32_t)BW2, (uint32_t)BH2, (uint32_t)BNI2};
        int rc = make_tensor_map(&tmX, p->in, MMVID_DT_F32, 4, dims, str, box);
        if (rc) return rc;
      }
      {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)p->Cout};
        uint64_t str[1] = {(uint64_t)K * 4};
        uint32_t box[2] = {32, 128};
        int rc = make_tensor_map(&tmW, p->w, MMVID_DT_F32, 2, dims, str, box);
        if (rc) return rc;
      }
      EpiArgs e{};
      e.bias = p->bias; e.residual = p->residual; e.ldr = p->Cout; e.C = p->out; e.ldc = p->Cout; e.c_h16 = 0; e.op_f16 = 0;
      e.M = M; e.N = p->Cout; e.K = K; e.act = MMVID_ACT_NONE;
      e.cCin = p->Cin; e.cKW = p->KW; e.cPadT = p->pad_t; e.cPadL = p->pad_l; e.cBW = BW2; e.cBH = BH2; e.cBNI = BNI2;
      e.cW = p->W; e.cH = p->H;
      return launch<true, 256, true, true>(tmX, tmW, e, st);
    }
  }
  const int BW = p->W < 128 ? p->W : 128;
  const int BH = (128 / BW) < p->H ? (128 / BW) : p->H;
  const int BNI = 128 / (BW * BH);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    uint64_t str[3] = {(uint64_t)p->Cin * esz, (uint64_t)p->W * p->Cin * esz, (uint64_t)p->H * p->W * p->Cin * esz};
    uint32_t box[4] = {(uint32_t)BKE, (uint32_t)BW, (uint32_t)BH, (uint32_t)BNI};
    int rc = make_tensor_map(&tmA, p->in, dt, 4, dims, str, box);
    if (rc) return rc;
  }
  const int BN = pick_bn(M, p->Cout);
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)p->Cout};
    uint64_t str[1] = {(uint64_t)K * esz};
    uint32_t box[2] = {(uint32_t)BKE, (uint32_t)BN};
    int rc = make_tensor_map(&tmB, p->w, dt, 2, dims, str, box);
    if (rc) return rc;
  }
  EpiArgs e{};
  e.bias = p->bias; e.residual = p->residual; e.ldr = p->Cout; e.C = p->out; e.ldc = p->Cout; e.c_h16 = 0; e.op_f16 = f16;
  e.M = M; e.N = p->Cout; e.K = K; e.act = MMVID_ACT_NONE;
  e.cCin = p->Cin; e.cKW = p->KW; e.cPadT = p->pad_t; e.cPadL = p->pad_l; e.cBW = BW; e.cBH = BH; e.cBNI = BNI;
  e.cW = p->W; e.cH = p->H;
  if (p->gn_partial != nullptr) {
    // fused GroupNorm statistics of the result: one image per pixel tile, whole groups per 32-channel chunk, and the
    // fp32 TMA-store epilogue (the only one that carries the hook) - refused loudly otherwise (mmvid_conv2d_gn_fusable)
    const int cpg = p->gn_groups > 0 ? p->Cout / p->gn_groups : 0;
    MMVID_REQUIRE(mmvid_conv2d_gn_fusable(p) == 1, "conv: fused GroupNorm statistics need the tensor-core path, H*W % 128 == 0, "
                                                   "Cout % 32 == 0 and 4, 8 or 16 channels per group");
    e.gn_partial = p->gn_partial; e.gn_cpg = cpg; e.gn_G = p->gn_groups;
  }
  return f16 ? launch_bn<false, true>(BN, tmA, tmB, e, st) : launch_bn<true, true>(BN, tmA, tmB, e, st);
}


// Fused QKV projection: qkv = A W^T + b written directly as Q,K [B,H,S_pad,64] and V^T [B,H,64,S_pad]
// (no [M, 3D] intermediate, no separate split/transposition pass).  Padding rows/columns are NOT touched:
// the caller keeps the buffers zero-initialised.
extern "C" int mmvid_linear_qkv(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                                const float* bias, void* q, void* k, void* vt, int out_dtype, int B, int H, int S,
                                int S_pad, int precision, mmvid_stream_t stream) {
  MMVID_REQUIRE(precision == MMVID_TF32 || precision == MMVID_BF16 || precision == MMVID_F16, "tensor-core precision required");
  MMVID_REQUIRE(S_pad >= S && S_pad % 64 == 0, "S_pad");
  const int want = precision == MMVID_TF32 ? MMVID_DT_F32 : (precision == MMVID_F16 ? MMVID_DT_F16 : MMVID_DT_BF16);
  if (out_dtype == want && a_dtype == want && w_dtype == want && env_int("MMVID_QKV_PAIR", 1)) {
    // 256 x 256 CTA-pair tiles with the TMA-store scatter (tc_gemm2.cu); 1 = preconditions not met, use the kernel below
    const int rc = mmvid_linear_qkv_tc2(A, lda, W, ldw, bias, q, k, vt, B, H, S, S_pad, precision, to_stream(stream));
    if (rc != 1) return rc;
  }
  g_qkv = QkvReq{q, k, vt, S, S_pad, H};
  const int rc = mmvid_linear_tc(A, a_dtype, lda, W, w_dtype, ldw, bias, nullptr, 0, q /*unused*/, out_dtype, 0,
                                 (long long)B * S, 3 * H * 64, H * 64, MMVID_ACT_NONE, precision, to_stream(stream));
  g_qkv = QkvReq{nullptr, nullptr, nullptr, 0, 0, 0};
  return rc;
}
