// 2-CTA tcgen05 GEMM (cta_group::2): a CTA pair (one TPC, cluster 2x1x1) owns a 256 x BN output tile.
// Each CTA loads ITS 128 rows of A and ITS half (BN/2 rows) of the weight tile; one tcgen05.mma.cta_group::2 issued by
// the leader CTA drives both SMs' tensor cores (M = 256) reading both CTAs' shared memory, so per CTA the operand
// bytes per FLOP drop by 25 % (BN = 128) / 50 % (BN = 256) against the single-CTA 128 x 128 tile and the freed shared
// memory deepens the TMA ring (8 / 6 stages).  Everything else mirrors tc_gemm.cu: persistent CTAs, two TMEM
// accumulators, 8 epilogue warps per CTA with batched tcgen05.ld and swizzled smem transposes.
//   full[s]      leader's barrier: leader expect_tx(bytes of BOTH CTAs); both CTAs' TMA loads complete_tx on it
//   empty[s]     one per CTA: tcgen05.commit.cta_group::2 ... multicast arrives on both
//   tmem_full[a] one per CTA (multicast commit);  tmem_empty[a] leader's barrier, 16 arrivals (8 warps x 2 CTAs)
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

#include <stdlib.h>

using namespace mmvid;
using namespace mmvid::tc;

namespace mmvid { extern unsigned long long* g_gemm_trace; }

namespace {

constexpr int BM = 128;  // rows per CTA (256 per pair)
constexpr int G2_THREADS = 576;  // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue (four per TMEM lane quarter)
constexpr int EPI_WARPS = 16;
constexpr int EPI_LD = 32;

struct Epi2Args {
  const float* bias; const float* residual; long long ldr;
  void* C; long long ldc;
  int c_h16;   // result type: 0 = fp32, 1 = bf16, 2 = fp16
  int op_f16;  // 16-bit kinds: operands are fp16 (1) or bf16 (0)
  long long M; int N, K, act;
  int num_m_tiles /* 256-row tiles */, num_n_tiles;
  int spin;  // 1: TMA / MMA threads poll their ring barriers (see tc_gemm.cu)
  int tma_store;  // 1: result tiles leave through TMA bulk stores (same epilogue as tc_gemm.cu)
  // QKV mode (qS > 0, TMA-store epilogue only): the [M, 3*H*64] result goes straight into the attention layout through
  // tmQ / tmK (3-D {64, S_pad, B*H}); V^T chunks are written from registers; see the epilogue
  int qS, qH, qB, qSpad;
  void* qkv_q; void* qkv_k; void* qkv_vt;
  unsigned long long* trace;  // debug timeline of cluster 0's leader CTA (same layout as gemm_stamp in tc_gemm.cu)
};

__device__ __forceinline__ void g2_stamp(const Epi2Args& e, uint32_t tile_iter, int idx) {
  if (e.trace != nullptr && blockIdx.x == 0 && tile_iter < 8) e.trace[tile_iter * 64 + idx] = clock64();
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");  // non-.aligned: role lanes arrive late
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {  // arrives on this barrier offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <bool TF32>
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
        : "memory");
  }
}

template <int BN>
constexpr int g2_stages() { return BN == 256 ? 5 : (BN == 192 ? 5 : 6); }  // 64 KB of the 227 KB stage the results
// TMEM columns to allocate for two BN-wide accumulators (the allocator wants a power of two)
template <int BN>
constexpr int g2_tmem_cols() { return 2 * BN <= 256 ? 256 : 512; }
template <int BN>
constexpr size_t g2_smem_bytes() {
  return (size_t)g2_stages<BN>() * (BM * 128 + (BN / 2) * 128) + EPI_WARPS * 32 * EPI_LD * 4 + 1024 + 512;
}

// H16: the TMA-store epilogue writes 16-bit results (compile-time so that the fp32 epilogue keeps its register budget)
template <bool TF32, int BN, bool H16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmQ,
                    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, Epi2Args e) {
  constexpr int STAGES = g2_stages<BN>();
  extern __shared__ uint8_t smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint64_t* res_bar = tmem_empty + 2;     // [EPI_WARPS] residual chunk landed in the warp's staging buffer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 512 + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * 128, B_BYTES = (BN / 2) * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int BKE = TF32 ? 32 : 64;
  float* epi_stage = reinterpret_cast<float*>(tiles + (size_t)STAGES * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_k = (e.K + BKE - 1) / BKE;
  const int total_tiles = e.num_m_tiles * e.num_n_tiles;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * EPI_WARPS); }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_ptr, g2_tmem_cols<BN>());
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  chain_release();
  chain_wait();  // activations (and the residual) come from the previous kernel; the setup above overlapped its tail

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int nt = tile % e.num_n_tiles, mt = tile / e.num_n_tiles;
        const int m0 = mt * (2 * BM) + (int)rank * BM;          // my 128 rows of the pair's 256
        const int n0 = nt * BN + (int)rank * (BN / 2);          // my half of the weight tile
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (e.spin) mbar_wait_spin(&empty[s], ph ^ 1); else mbar_wait(&empty[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);  // bytes landing in BOTH CTAs
          const uint32_t full_leader = mapa_u32(smem_u32(&full[s]), 0);
          uint8_t* a = tiles + s * STAGE_BYTES;
          tma_load_2d_2sm(a, &tmA, full_leader, kb * BKE, m0);
          tma_load_2d_2sm(a + A_BYTES, &tmB, full_leader, kb * BKE, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = TF32 ? make_idesc<true>(2 * BM, BN) : make_idesc_h16(e.op_f16 != 0, 2 * BM, BN);
      uint32_t it = 0, tile_iter = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tile_iter) {
        const uint32_t acc = tile_iter & 1, acc_ph = (tile_iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        g2_stamp(e, tile_iter, 0);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (e.spin) mbar_wait_spin(&full[s], ph); else mbar_wait(&full[s], ph);
          tc_fence_after();
          if (e.trace != nullptr) {
            if (kb == 0) g2_stamp(e, tile_iter, 1);
            if (kb == num_k - 1) g2_stamp(e, tile_iter, 2);
            if (kb < 32) g2_stamp(e, tile_iter, 24 + kb);
          }
          const uint32_t a_addr = smem_u32(tiles + s * STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ss2<TF32>(d_tmem, desc_advance(a_desc, kk * 32), desc_advance(b_desc, kk * 32), idesc, (kb | kk) != 0 ? 1u : 0u);
          tc_commit2(&empty[s]);
        }
        tc_commit2(&tmem_full[acc]);
        g2_stamp(e, tile_iter, 3);
      }
    }
  } else {
    // ---------------- epilogue: warps 2..17.  TMEM lane quarter = warp % 4 (hardware rule), column part = (warp - 2) / 4:
    // FOUR warps per SM sub-partition instead of two.  The r2f trace showed the 8-warp epilogue to need 13.4 k clk per
    // 256 x 256 tile with QuickGELU (7.7 k without) against a 5.8 k clk fp16 main loop: two warps per sub-partition cannot
    // hide the tcgen05.ld / MUFU / bulk-store latencies of a chunk behind each other.  Each warp now owns 1-2 32-column
    // chunks of the tile (one 64-column group), a lane keeps its own output row: bias / QuickGELU / residual in registers,
    // one write of the group into the warp's 4 KB staging buffer (SWIZZLE_128B layout), one bulk store; the TMA unit clips
    // the M tail.  fp32 results leave per 32-column chunk, 16-bit results (H16) per group (tc_epilogue.cuh).
    const int q = warp & 3;
    const int ew = warp - 2;
    const int cp = ew >> 2;
    constexpr int TC = BN / 32;  // 32-column chunks per tile: 4 | 6 | 8
    const int ch0 = cp * TC / 4, nch = (cp + 1) * TC / 4 - ch0;  // this warp's chunks [ch0, ch0 + nch), nch = 1 | 2
    const uint32_t st_base = smem_u32(epi_stage) + (uint32_t)(ew * 4096);
    const uint32_t st_row = st_base + (uint32_t)(lane * 128);
    if (lane == 0) { prefetch_tmap(&tmC); prefetch_tmap(&tmQ); prefetch_tmap(&tmV); }
    uint32_t tile_iter = 0, res_ph = 0;
    float bias_nx0 = 0.f, bias_nx1 = 0.f;  // this lane's bias columns of the NEXT tile's two chunks
    if (H16 && cluster_id < total_tiles) {
      const int ng = (cluster_id % e.num_n_tiles) * BN + ch0 * 32 + lane;
      bias_nx0 = (e.bias && ng < e.N) ? __ldg(e.bias + ng) : 0.f;
      bias_nx1 = (e.bias && nch > 1 && ng + 32 < e.N) ? __ldg(e.bias + ng + 32) : 0.f;
    }
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tile_iter) {
      const int nt = tile % e.num_n_tiles, mt = tile / e.num_n_tiles;
      const int m0 = mt * (2 * BM) + (int)rank * BM;
      const int n_grp0 = nt * BN + ch0 * 32;
      const uint32_t acc = tile_iter & 1, acc_ph = (tile_iter >> 1) & 1;
      const long long m_row = (long long)m0 + q * 32 + lane;
      const bool row_ok = m_row < e.M;
      int nvalid = 0;
      for (int cc = 0; cc < nch; ++cc)
        if (n_grp0 + cc * 32 < e.N) nvalid = cc + 1;
      // fp32 results with a residual: the residual chunk is fetched by a TMA LOAD into the warp's staging buffer (coalesced,
      // asynchronous, no registers) - for the first chunk before the accumulator is even complete -, the lanes add their
      // own row in place and the buffer leaves again as the bulk store.  (Per-lane 16-byte loads of 32 different rows took
      // ~8 k clk per chunk in the r2g trace: with ~1 KB of L1 left beside 227 KB of shared memory every load goes to L2.)
      const bool res_tma = !H16 && e.residual != nullptr && e.qS == 0;
      if (res_tma && nvalid > 0 && lane == 0) {
        tma_store_wait_read();  // the buffer's previous bulk store has been read out
        mbar_expect_tx(&res_bar[ew], 4096);
        tma_load_2d(epi_stage + ew * 1024, &tmV, &res_bar[ew], n_grp0, m0 + q * 32);
      }
      // Bias: lane i keeps column i of each of the warp's two chunks (fetched ONE TILE AHEAD, see below) and the row owners
      // pick the values up by shuffle.  Beside 227 KB of shared memory only ~1 KB of L1 is left, so a __ldg behind
      // tcgen05.wait::ld was an exposed L2 round trip of ~800 clk per chunk.
      float bias_c0 = 0.f, bias_c1 = 0.f;
      if constexpr (H16) {
        bias_c0 = bias_nx0; bias_c1 = bias_nx1;
        const int tile_n = tile + num_clusters;
        if (tile_n < total_tiles) {
          const int ng = (tile_n % e.num_n_tiles) * BN + ch0 * 32 + lane;
          bias_nx0 = (e.bias && ng < e.N) ? __ldg(e.bias + ng) : 0.f;
          bias_nx1 = (e.bias && nch > 1 && ng + 32 < e.N) ? __ldg(e.bias + ng + 32) : 0.f;
        }
      }
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const bool etr = (warp == 2 && lane == 0);
      if (etr) g2_stamp(e, tile_iter, 8);
      const uint32_t t_src = tmem_base + acc * BN + ch0 * 32 + ((uint32_t)(q * 32) << 16);
      // One chunk at a time (576 threads leave 96 registers per thread): the second chunk's tcgen05.ld is issued after the
      // first has been processed, the accumulator is released as soon as this warp's last chunk is in registers.
      // fused QKV scatter: groups start at multiples of 64 columns, so a group is the 64 head-dim values of ONE head of
      // Q, K or V.  16-bit V^T groups are written straight from registers.
      const bool vt_grp = H16 && e.qS > 0 && n_grp0 >= 2 * e.qH * 64;
      if (H16 && !vt_grp && nvalid > 0) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (cc >= nch) continue;  // warp-uniform
        uint32_t racc[32];
        tmem_ld32(t_src + cc * 32, racc);
        tmem_ld_wait();
        if (cc == nch - 1) {
          // this warp's share of the accumulator is in registers: release it (leader's barrier, 2 x 16 arrivals)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
          if (etr) g2_stamp(e, tile_iter, 9);
        }
        if (cc >= nvalid) continue;  // columns beyond N (warp-uniform)
        const int ncol = n_grp0 + cc * 32;
        float o[32];
        if constexpr (H16) {
          const float bias_l = cc == 0 ? bias_c0 : bias_c1;
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(racc[i]) + __shfl_sync(0xffffffffu, bias_l, i);
        } else {  // fp32 results: these GEMMs are not epilogue-bound and have no registers to spare - plain loads
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + ncol + 4 * i));
            o[4 * i + 0] = __uint_as_float(racc[4 * i + 0]) + b.x;
            o[4 * i + 1] = __uint_as_float(racc[4 * i + 1]) + b.y;
            o[4 * i + 2] = __uint_as_float(racc[4 * i + 2]) + b.z;
            o[4 * i + 3] = __uint_as_float(racc[4 * i + 3]) + b.w;
          }
        }
        if (e.act != MMVID_ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = apply_act_fast(o[i], e.act);
        }
        if (res_tma) {
          if (cc == 1 && lane == 0) {  // second chunk: its residual can only land once the first chunk's store has been read out
            tma_store_wait_read();
            mbar_expect_tx(&res_bar[ew], 4096);
            tma_load_2d(epi_stage + ew * 1024, &tmV, &res_bar[ew], ncol, m0 + q * 32);
          }
          mbar_wait(&res_bar[ew], res_ph);
          res_ph ^= 1u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = lds128(st_row + (uint32_t)((j ^ (lane & 7)) * 16));
            o[4 * j + 0] += r4.x; o[4 * j + 1] += r4.y; o[4 * j + 2] += r4.z; o[4 * j + 3] += r4.w;
          }
        } else if (e.residual && row_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 r4 = *reinterpret_cast<const float4*>(e.residual + m_row * e.ldr + ncol + 4 * i);
            o[4 * i + 0] += r4.x; o[4 * i + 1] += r4.y; o[4 * i + 2] += r4.z; o[4 * i + 3] += r4.w;
          }
        }
        if constexpr (H16) {
          const bool f16 = e.c_h16 == 2;
          if (vt_grp) {
            // for a fixed head-dim index the 32 lanes write 32 consecutive tokens: one coalesced 64-byte store
            const int hh = (ncol - 2 * e.qH * 64) >> 6, d0 = ncol & 63;
            if (row_ok) {
              const int b2 = (int)(m_row / e.qS), s2 = (int)(m_row - (long long)b2 * e.qS);
              uint16_t* dst = reinterpret_cast<uint16_t*>(e.qkv_vt) + (((long long)b2 * e.qH + hh) * 64 + d0) * e.qSpad + s2;
#pragma unroll
              for (int d = 0; d < 32; ++d) dst[(long long)d * e.qSpad] = cvt_h16_rt(o[d], f16);
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t u0 = pack_h16_rt(o[8 * j + 0], o[8 * j + 1], f16), u1 = pack_h16_rt(o[8 * j + 2], o[8 * j + 3], f16);
            const uint32_t u2 = pack_h16_rt(o[8 * j + 4], o[8 * j + 5], f16), u3 = pack_h16_rt(o[8 * j + 6], o[8 * j + 7], f16);
            uint32_t addr;
            if (nvalid == 2) addr = st_row + (uint32_t)(((cc * 4 + j) ^ (lane & 7)) * 16);
            else addr = st_base + (uint32_t)(lane * 64 + j * 16);  // lone chunk: dense 64-byte rows (SWIZZLE_NONE map)
            sts128_u32(addr, u0, u1, u2, u3);
          }
        } else {
          if (!res_tma) {
            if (lane == 0) tma_store_wait_read();  // the previous chunk's bulk store has finished READING the staging buffer
            __syncwarp();
          }
          if (e.qS > 0) {
            // ---- fused QKV scatter (fp32).  A chunk is 32 tokens x 32 columns of one head: Q / K chunks are stored as they
            // are into [B*H, S_pad, 64] by one 3-D TMA store.  Chunks whose 32 tokens straddle a batch boundary (or the end
            // of the last batch) are written row by row from registers: a TMA box cannot be clipped at S < S_pad, and the
            // padding rows must stay zero.  V^T chunks always go out from registers: for a fixed head-dim index the 32 lanes
            // write 32 consecutive tokens (one coalesced 128-byte store), and a TMA box into V^T would need its first token
            // 16-byte aligned, which an odd S rules out for every batch but the first.
            const int D = e.qH * 64;
            const int which = ncol / D, hh = (ncol - which * D) >> 6, d0 = ncol & 63;
            const int mb = m0 + q * 32;
            const int bb = mb / e.qS, ss0 = mb - bb * e.qS;
            if (which == 2 || ss0 + 32 > e.qS || bb >= e.qB) {  // warp-uniform
              if (row_ok) {
                const int b2 = (int)(m_row / e.qS), s2 = (int)(m_row - (long long)b2 * e.qS);
                if (which < 2) {
                  float* dst = reinterpret_cast<float*>(which == 0 ? e.qkv_q : e.qkv_k) + (((long long)b2 * e.qH + hh) * e.qSpad + s2) * 64 + d0;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                } else {
                  float* dst = reinterpret_cast<float*>(e.qkv_vt) + (((long long)b2 * e.qH + hh) * 64 + d0) * e.qSpad + s2;
#pragma unroll
                  for (int d = 0; d < 32; ++d) dst[(long long)d * e.qSpad] = o[d];
                }
              }
              continue;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(st_row + (uint32_t)((j ^ (lane & 7)) * 16), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_3d(which == 0 ? &tmQ : &tmK, st_base, d0, ss0, bb * e.qH + hh);
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)  // 16-byte unit j of row r lives at unit j ^ (r & 7): the TMA 128-byte swizzle
            sts128(st_row + (uint32_t)((j ^ (lane & 7)) * 16), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmC, st_base, ncol, m0 + q * 32);
        }
      }
      if (H16 && !vt_grp && nvalid > 0) {
        if (e.qS > 0) {
          // Q / K group: one {64, 32, 1} box, or row by row where the 32 tokens straddle a batch boundary
          const int D = e.qH * 64;
          const int which = n_grp0 / D, hh = (n_grp0 - which * D) >> 6;
          const int mb = m0 + q * 32;
          const int bb = mb / e.qS, ss0 = mb - bb * e.qS;
          if (ss0 + 32 > e.qS || bb >= e.qB) {  // warp-uniform
            __syncwarp();
            if (row_ok) {
              const int b2 = (int)(m_row / e.qS), s2 = (int)(m_row - (long long)b2 * e.qS);
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(which == 0 ? e.qkv_q : e.qkv_k) +
                                                    (((long long)b2 * e.qH + hh) * e.qSpad + s2) * 64);
#pragma unroll
              for (int u = 0; u < 8; ++u) {  // this lane's own staged row
                const float4 v = lds128(st_row + (uint32_t)((u ^ (lane & 7)) * 16));
                dst[u] = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
              }
            }
            __syncwarp();
          } else {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_3d(which == 0 ? &tmQ : &tmK, st_base, 0, ss0, bb * e.qH + hh);
          }
        } else {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) tma_store_2d(nvalid == 2 ? &tmC : &tmQ /* {32 x 32} box */, st_base, n_grp0, m0 + q * 32);
        }
      }
      if (etr) g2_stamp(e, tile_iter, 10);
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, g2_tmem_cols<BN>());
  }
}

struct QkvMaps { CUtensorMap q, k, v; };

template <bool TF32, int BN, bool H16>
int launch2k(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmQ,
             const CUtensorMap& tmK, const CUtensorMap& tmR, const Epi2Args& e, int clusters, const char* what,
             cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t smem = g2_smem_bytes<BN>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tc2_kernel<TF32, BN, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(gemm_tc2): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  cudaError_t err = launch_chained(gemm_tc2_kernel<TF32, BN, H16>, dim3(2 * clusters), dim3(G2_THREADS), smem, st, tmA, tmB, tmC, tmQ,
                                   tmK, tmR, e);
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "gemm_tc2 launch: %s", cudaGetErrorString(err));
  return check_launch(what);
}

template <bool TF32, int BN>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, Epi2Args e, cudaStream_t st, const QkvMaps* qm = nullptr) {
  e.num_m_tiles = (int)ceil_div<long long>(e.M, 2 * BM);
  e.num_n_tiles = ceil_div(e.N, BN);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (long long)e.num_m_tiles * e.num_n_tiles;
  const int clusters = (int)(tiles < sms / 2 ? tiles : sms / 2);
  CUtensorMap tmC = tmA, tmC1 = tmA, tmR = tmA;  // placeholders when unused
  e.tma_store = 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int csz = e.c_h16 ? 2 : 4;
  if (qm == nullptr && !(TF32 && e.c_h16) && e.N % 32 == 0 && (e.ldc * csz) % 16 == 0 && al16(e.C) &&
      (!e.residual || (e.ldr % 4 == 0 && al16(e.residual))) && (!e.bias || al16(e.bias))) {
    uint64_t dims[2] = {(uint64_t)e.N, (uint64_t)e.M};
    uint64_t str[1] = {(uint64_t)e.ldc * csz};
    if (e.c_h16) {
      // 16-bit results: {64 x 32} boxes (two accumulator chunks, 128-byte rows) + {32 x 32} boxes for a lone chunk
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
      if (e.residual) {  // fetched chunk by chunk through the same kind of box (may be C itself: in-place residual add)
        uint64_t rstr[1] = {(uint64_t)e.ldr * 4};
        rc = make_tensor_map(&tmR, e.residual, DT_F32_EXACT, 2, dims, rstr, box);
        if (rc) return rc;
      }
    }
    e.tma_store = 1;
  }
  if (qm != nullptr) {
    e.tma_store = 1;  // the QKV scatter only exists in the TMA-store epilogue (the caller checked its preconditions)
    if constexpr (!TF32) {
      if (e.c_h16) return launch2k<TF32, BN, true>(tmA, tmB, tmA, qm->q, qm->k, tmA, e, clusters, "gemm_tc2_qkv", st);
    }
    return launch2k<TF32, BN, false>(tmA, tmB, tmA, qm->q, qm->k, tmA, e, clusters, "gemm_tc2_qkv", st);
  }
  if (!e.tma_store) return 1;  // results leave through TMA bulk stores only: the caller uses the single-CTA kernel instead
  if constexpr (!TF32) {
    if (e.c_h16) return launch2k<TF32, BN, true>(tmA, tmB, tmC, tmC1, tmA, tmR, e, clusters, "gemm_tc2", st);
  }
  return launch2k<TF32, BN, false>(tmA, tmB, tmC, tmC1, tmA, tmR, e, clusters, "gemm_tc2", st);
}

}  // namespace

// 2-CTA path of mmvid_linear (selected by mmvid_linear_tc).  Returns 1 ("not applicable") when the result cannot leave
// through TMA bulk stores (N % 32, alignment, fp32 operands with a 16-bit result): the caller then runs the single-CTA kernel.
extern "C" int mmvid_linear_tc2(const void* A, int a_dtype, long long lda, const void* W, int w_dtype, long long ldw,
                                const float* bias, const float* residual, long long ldr, void* C, int c_dtype,
                                long long ldc, long long M, int N, int K, int act, int precision, int BN, cudaStream_t st) {
  const bool tf32 = precision == MMVID_TF32;
  const int esz = tf32 ? 4 : 2;
  const int BKE = tf32 ? 32 : 64;
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
    uint32_t box[2] = {(uint32_t)BKE, (uint32_t)(BN / 2)};
    int rc = make_tensor_map(&tmB, W, w_dtype, 2, dims, str, box);
    if (rc) return rc;
  }
  Epi2Args e{};
  e.bias = bias; e.residual = residual; e.ldr = ldr; e.C = C; e.ldc = ldc;
  e.c_h16 = c_dtype == MMVID_DT_BF16 ? 1 : (c_dtype == MMVID_DT_F16 ? 2 : 0);
  e.op_f16 = precision == MMVID_F16;
  e.M = M; e.N = N; e.K = K; e.act = act;
  { const char* v = getenv("MMVID_GEMM_SPIN"); e.spin = v ? atoi(v) : 0; }
  e.trace = mmvid::g_gemm_trace;
  if (BN == 256) return tf32 ? launch2<true, 256>(tmA, tmB, e, st) : launch2<false, 256>(tmA, tmB, e, st);
  if (BN == 192) return tf32 ? launch2<true, 192>(tmA, tmB, e, st) : launch2<false, 192>(tmA, tmB, e, st);
  return tf32 ? launch2<true, 128>(tmA, tmB, e, st) : launch2<false, 128>(tmA, tmB, e, st);
}

// Fused QKV projection on the CTA-pair kernel: qkv = A W^T + b scattered straight into Q, K [B,H,S_pad,64] (TMA stores) and
// V^T [B,H,64,S_pad] (coalesced register stores).  tf32: fp32 operands and fp32 Q / K / V^T; 16-bit kinds: operands and
// outputs in the precision's 16-bit type.  Returns 1 ("not applicable") when the preconditions do not hold, so that
// mmvid_linear_qkv can fall back to the single-CTA kernel.
extern "C" int mmvid_linear_qkv_tc2(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* q,
                                    void* k, void* vt, int B, int H, int S, int S_pad, int precision, cudaStream_t st) {
  const long long M = (long long)B * S;
  const int N = 3 * H * 64, K = H * 64;
  const bool tf32 = precision == MMVID_TF32;
  const int esz = tf32 ? 4 : 2, BKE = tf32 ? 32 : 64;
  const int dt_op = tf32 ? MMVID_DT_F32 : (precision == MMVID_F16 ? MMVID_DT_F16 : MMVID_DT_BF16);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (S < 32 || N % 256 != 0 || M < 1024 || !al16(q) || !al16(k) || !al16(vt) || (bias && !al16(bias)) || S_pad % 8 != 0) return 1;
  CUtensorMap tmA, tmB;
  QkvMaps qm;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * esz};
    uint32_t box[2] = {(uint32_t)BKE, (uint32_t)BM};
    int rc = make_tensor_map(&tmA, A, dt_op, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * esz};
    uint32_t box[2] = {(uint32_t)BKE, 128};
    int rc = make_tensor_map(&tmB, W, dt_op, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    // Q / K boxes: tf32 one 32-column chunk of 32 tokens, 16-bit a whole head (two chunks) of 32 tokens
    uint64_t dims[3] = {64, (uint64_t)S_pad, (uint64_t)B * H};
    uint64_t str[2] = {(uint64_t)64 * esz, (uint64_t)S_pad * 64 * esz};
    uint32_t box[3] = {tf32 ? 32u : 64u, 32, 1};
    int rc = make_tensor_map(&qm.q, q, tf32 ? DT_F32_EXACT : dt_op, 3, dims, str, box);
    if (rc) return rc;
    rc = make_tensor_map(&qm.k, k, tf32 ? DT_F32_EXACT : dt_op, 3, dims, str, box);
    if (rc) return rc;
  }
  qm.v = qm.q;  // unused: V^T chunks are written from registers
  Epi2Args e{};
  e.bias = bias; e.residual = nullptr; e.ldr = 0; e.C = nullptr; e.ldc = 0;
  e.c_h16 = tf32 ? 0 : (precision == MMVID_F16 ? 2 : 1);
  e.op_f16 = precision == MMVID_F16;
  e.M = M; e.N = N; e.K = K; e.act = MMVID_ACT_NONE;
  e.qS = S; e.qH = H; e.qB = B; e.qSpad = S_pad;
  e.qkv_q = q; e.qkv_k = k; e.qkv_vt = vt;
  { const char* v = getenv("MMVID_GEMM_SPIN"); e.spin = v ? atoi(v) : 0; }
  e.trace = mmvid::g_gemm_trace;
  return tf32 ? launch2<true, 256>(tmA, tmB, e, st, &qm) : launch2<false, 256>(tmA, tmB, e, st, &qm);
}
