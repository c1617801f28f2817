// Flash attention v6 for sm_100a: the rotating-score-buffer kernel of tc_attention3.cu made PERSISTENT.
//
// ncu of v5 at the benchmark shape: 17 key steps x 2547 clk = 43 k clk of steady state inside ~63 k clk of CTA lifetime;
// TMEM allocation, the Q / K loads, the first QK^T, the pipeline ramp and the O read-out are a third of the kernel,
// and with 197 KB of shared memory and all 512 TMEM columns per CTA nothing of the next CTA can overlap them.  Here one
// CTA per SM walks a list of work items (one item = one (batch, head) and one pair of 128-query tiles) and treats the
// tile-steps of all its items as ONE sequence: global step index G = 2 j + g counted across items selects the score
// buffer (G mod 3), the barriers (G mod 6) and the issuing MMA thread (G mod 2), so the invariant "PV(G) and QK(G+3) come
// from the same thread in that order" holds across item boundaries and the first three QK^T of item i+1 are in flight
// while the softmax warps still read out O of item i.  Extra hand-shakes per item: q_empty (last QK^T of an item issued ->
// the Q tiles may be reloaded), o_drained[g] (O_g is in registers -> the first PV of the next item may overwrite it);
// o_full[g] replaces the end-of-kernel barrier.  O goes from registers straight to global memory (a lane owns 256
// contiguous bytes of one row): the K / V stages that v5 recycled as a staging area are busy with the next item.
// Requires at least two key tiles per item (S > 128); mmvid_attention falls back to v5 otherwise.
#include "common.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

using namespace mmvid;
using namespace mmvid::tc;

namespace {

constexpr int ATT4_THREADS = 384;
constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TMEM_COLS = 512;
__host__ __device__ constexpr int X_COL_OF(int buf) { return buf * 128; }
__host__ __device__ constexpr int O_COL_OF(int g) { return 384 + g * 64; }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// packed pairs of floats in one 64-bit register (Blackwell f32x2 FMA-pipe instructions)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float a, float b) {
  f2_t p;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
  return p;
}
__device__ __forceinline__ void unpack2(f2_t p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// 2^x for a pair on the FMA pipe.  x <= ~9 (lazy-rescale threshold); x = -inf / very negative -> 2^-126 (~1e-38: a
// masked key contributes nothing measurable; avoids an exponent-field borrow).  n = round(x) falls out of adding
// 1.5 * 2^23 (its low mantissa bits then hold n in two's complement), f = x - n in [-0.5, 0.5], 2^f by a minimax
// polynomial, and 2^n is applied by adding n << 23 to the exponent field (the magic constant's own bits shift out).
template <int DEG>
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& e0, float& e1) {
  const f2_t MAGIC = pack2(12582912.f, 12582912.f), NMAGIC = pack2(-12582912.f, -12582912.f), NONE = pack2(-1.f, -1.f);
  const f2_t X = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const f2_t T = add2(X, MAGIC);
  const f2_t F = fma2(add2(T, NMAGIC), NONE, X);
  f2_t P;
  if constexpr (DEG == 4) {
    P = fma2(pack2(0.009570101276040077f, 0.009570101276040077f), F, pack2(0.05591785907745361f, 0.05591785907745361f));
    P = fma2(P, F, pack2(0.240247443318367f, 0.240247443318367f));
    P = fma2(P, F, pack2(0.6931217908859253f, 0.6931217908859253f));
    P = fma2(P, F, pack2(0.9999992847442627f, 0.9999992847442627f));
  } else {
    P = fma2(pack2(0.0551716648042202f, 0.0551716648042202f), F, pack2(0.2426111251115799f, 0.2426111251115799f));
    P = fma2(P, F, pack2(0.6932609677314758f, 0.6932609677314758f));
    P = fma2(P, F, pack2(0.9999280571937561f, 0.9999280571937561f));
  }
  float t0, t1, p0, p1;
  unpack2(T, t0, t1);
  unpack2(P, p0, p1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

struct Att4Args {
  unsigned long long* trace;  // debug timeline (mmvid_debug_attention_trace), normally null
  void* out; long long ldo; int out_bf16;
  int B, H, S, S_pad, mask_kind;
  int prev_rows[4]; int n_prev;
  int n_pairs, n_items;
};

__device__ __forceinline__ void att4_stamp(const Att4Args& a, int idx) {
  if (a.trace != nullptr && blockIdx.x == 0) a.trace[idx] = clock64();
}

template <bool TF32>
constexpr size_t att4_smem_bytes() {
  return (size_t)6 * (TF32 ? 32768 : 16384) + 1024 + 512;
}

template <bool TF32, int POLY8>
__global__ void __launch_bounds__(ATT4_THREADS, 1) attention_tc4_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV, Att4Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]  (count 2: QK of tile A and of tile B both read the stage)
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]  (count 2)
  uint64_t* s_full = bars + 9;     // [6] QK(G) landed in X(G mod 3); indexed by G mod 6 (see tc_attention3.cu)
  uint64_t* p_ready = bars + 15;   // [6] P(G) written over it (count 4: one arrival per warp)
  uint64_t* o_full = bars + 21;    // [2] per query tile: PV retired (one completion per key tile, counted across items)
  uint64_t* q_empty = bars + 23;   // the last QK^T of both query tiles of an item is issued (count 2)
  uint64_t* o_drained = bars + 24; // [2] O_g of the finished item is in registers (count 4: one arrival per warp)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 512 + 1023) & ~(uintptr_t)1023);
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int BKE = 128 / ESZ;
  constexpr int QK_KB = HD / BKE;   // 2 | 1
  constexpr int PV_KB = BKV / BKE;  // 4 | 2
  constexpr int T_BYTES = BQ * HD * ESZ;
  auto sQ = [&](int g) { return tiles + g * T_BYTES; };
  auto sK = [&](int st) { return tiles + (2 + st) * T_BYTES; };
  auto sV = [&](int st) { return tiles + (4 + st) * T_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_all = (a.S + BKV - 1) / BKV;
  const int n_my = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // item I of this CTA -> (batch*head, first query row, key tiles).  Late query pairs first: under the causal mask they
  // are the long ones.
  auto item_of = [&](int I, int& bh, int& q0, int& nkv) {
    const int it = (int)blockIdx.x + I * (int)gridDim.x;
    bh = it / a.n_pairs;
    q0 = (a.n_pairs - 1 - (it - bh * a.n_pairs)) * (2 * BQ);
    nkv = n_kv_all;
    if (a.mask_kind == MMVID_MASK_CAUSAL) nkv = min(nkv, (q0 + 2 * BQ - 1) / BKV + 1);
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
      mbar_init(&o_full[i], 1); mbar_init(&o_drained[i], 4);
    }
    for (int i = 0; i < 6; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ Q + K producer
    if (elect_one()) {
      uint32_t J = 0;  // key tiles loaded so far (all items): stage J & 1, phase (J >> 1) & 1
      for (int I = 0; I < n_my; ++I) {
        int bh, q0, nkv;
        item_of(I, bh, q0, nkv);
        if (I > 0) mbar_wait(q_empty, (uint32_t)(I - 1) & 1);  // every QK^T of the previous item has been issued and retired
        mbar_expect_tx(q_full, 2 * T_BYTES);
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int kb = 0; kb < QK_KB; ++kb)
            tma_load_2d(sQ(g) + kb * (BQ * 128), &tmQ, q_full, kb * BKE, bh * a.S_pad + q0 + g * BQ);
        for (int j = 0; j < nkv; ++j, ++J) {
          const int st = J & 1;
          mbar_wait(&k_empty[st], ((J >> 1) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], T_BYTES);
#pragma unroll
          for (int kb = 0; kb < QK_KB; ++kb)
            tma_load_2d(sK(st) + kb * (BKV * 128), &tmK, &k_full[st], kb * BKE, bh * a.S_pad + j * BKV);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ V^T producer
    if (elect_one()) {
      uint32_t J = 0;
      for (int I = 0; I < n_my; ++I) {
        int bh, q0, nkv;
        item_of(I, bh, q0, nkv);
        for (int j = 0; j < nkv; ++j, ++J) {
          const int st = J & 1;
          mbar_wait(&v_empty[st], ((J >> 1) & 1) ^ 1);
          mbar_expect_tx(&v_full[st], T_BYTES);
#pragma unroll
          for (int kb = 0; kb < PV_KB; ++kb)
            tma_load_2d(sV(st) + kb * (HD * 128), &tmV, &v_full[st], j * BKV + kb * BKE, bh * HD);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 even global steps, warp 3 odd
    const int my = warp == 3 ? 1 : 0;
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc<TF32>(BQ, BKV);
      constexpr uint32_t idesc_pv = make_idesc<TF32>(BQ, HD);
      constexpr uint32_t TB16 = T_BYTES >> 4;
      const uint64_t dQ0 = make_smem_desc_sw128(smem_u32(sQ(0)));
      const uint64_t dK0 = make_smem_desc_sw128(smem_u32(sK(0)));
      const uint64_t dV0 = make_smem_desc_sw128(smem_u32(sV(0)));
      // QK of local step n of an item whose steps start at global index Gb, key tiles at Jb, and which has ns steps
      auto issue_qk = [&](int n, uint32_t Gb, uint32_t Jb, int ns) {
        const int j = n >> 1, g = n & 1;
        const uint32_t Gn = Gb + (uint32_t)n, Jg = Jb + (uint32_t)j;
        const int st = Jg & 1;
        mbar_wait(&k_full[st], (Jg >> 1) & 1);
        tc_fence_after();
        const uint32_t x = tmem_base + (Gn % 3) * 128;
        const uint64_t qd = dQ0 + (uint64_t)(g * TB16), kd = dK0 + (uint64_t)(st * TB16);
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ss<TF32>(x, qd + (uint64_t)(kb * (BQ * 128 / 16) + kk * 2), kd + (uint64_t)(kb * (BKV * 128 / 16) + kk * 2),
                         idesc_qk, (kb | kk) != 0);
        tc_commit(&k_empty[st]);
        tc_commit(&s_full[Gn % 6]);
        if (n + 2 >= ns) tc_commit(q_empty);  // last QK^T of this query tile: one of the two arrivals that free the Q tiles
      };
      auto issue_pv = [&](int n, uint32_t Gb, uint32_t Jb) {
        const int j = n >> 1, g = n & 1;
        const uint32_t Gn = Gb + (uint32_t)n, Jg = Jb + (uint32_t)j;
        const int st = Jg & 1;
        const uint32_t x = tmem_base + (Gn % 3) * 128, o = tmem_base + O_COL_OF(0) + g * 64;
        const uint64_t vd = dV0 + (uint64_t)(st * TB16);
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ts<TF32>(o, x + kb * 32 + kk * 8, vd + (uint64_t)(kb * (HD * 128 / 16) + kk * 2), idesc_pv,
                         (j != 0 || (kb | kk) != 0) ? 1u : 0u);
        tc_commit(&v_empty[st]);
        tc_commit(&o_full[g]);
      };
      uint32_t G = 0, J = 0;  // global step / key-tile index of the current item's first step / tile
      int bh, q0, nkv;
      item_of(0, bh, q0, nkv);
      mbar_wait(q_full, 0);
      tc_fence_after();
      if (my == 0) {  // prologue of the very first item: three QK^T fill the three score buffers
        issue_qk(0, 0, 0, 2 * nkv);
        issue_qk(1, 0, 0, 2 * nkv);
        issue_qk(2, 0, 0, 2 * nkv);
      }
      for (int I = 0; I < n_my; ++I) {
        const int ns = 2 * nkv;
        const bool has_next = I + 1 < n_my;
        int bh2 = 0, q02 = 0, nkv2 = 0;
        if (has_next) item_of(I + 1, bh2, q02, nkv2);
        for (int n = my; n < ns; n += 2) {
          const int j = n >> 1;
          const uint32_t Gn = G + (uint32_t)n, Jg = J + (uint32_t)j;
          mbar_wait(&v_full[Jg & 1], (Jg >> 1) & 1);
          if (j == 0 && I > 0) mbar_wait(&o_drained[my], (uint32_t)(I - 1) & 1);  // O of the previous item has been read out
          mbar_wait(&p_ready[Gn % 6], (Gn / 6) & 1);
          tc_fence_after();
          if (I == 0 && n < 64) att4_stamp(a, n * 2);
          issue_pv(n, G, J);
          if (n + 3 < ns) issue_qk(n + 3, G, J, ns);  // same score buffer: PV (same issue stream) consumes P first
          if (I == 0 && n < 64) att4_stamp(a, n * 2 + 1);
        }
        // the steps G+ns .. G+ns+2 are the first three QK^T of the NEXT item; they go out after this thread's last PV
        // of the current item so that waiting for the reloaded Q tiles cannot delay that PV (buffer order still holds: the
        // PV that used each buffer was issued earlier by this same thread)
        if (has_next) {
          mbar_wait(q_full, (uint32_t)(I + 1) & 1);
          tc_fence_after();
          for (int m = 0; m < 3; ++m)
            if (((ns + m) & 1) == ((my + 3) & 1)) issue_qk(m, G + (uint32_t)ns, J + (uint32_t)nkv, 2 * nkv2);
        }
        G += (uint32_t)ns; J += (uint32_t)nkv;
        bh = bh2; q0 = q02; nkv = nkv2;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = (warp - 4) >> 2;  // 0: tile A, 1: tile B
    const int qd = warp & 3;
    const int row_local = qd * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t t_o = t_row + O_COL_OF(g);
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    constexpr float RESCALE_THRESH = 8.f;
    uint32_t G = 0, J = 0;
    for (int I = 0; I < n_my; ++I) {
      int bh, q0, n_kv;
      item_of(I, bh, q0, n_kv);
      const int b = bh / a.H, h = bh - b * a.H;
      const int row = q0 + g * BQ + row_local;
      int lo = 0, hi = a.S;
      if (a.mask_kind == MMVID_MASK_CAUSAL) hi = min(a.S, row + 1);
      else if (a.mask_kind == MMVID_MASK_PREV) {
        for (int i = 0; i < a.n_prev; ++i) if (a.prev_rows[i] == row) lo = row;
      }
      float m_ref = -INFINITY, l = 0.f;
      const bool itr = (qd == 0 && lane == 0 && I < 4);  // item-level stamps: [480 + g*16 + I*4 + {0 start, 1 last P sent, 2 O complete, 3 stored}]
      if (itr) att4_stamp(a, 480 + g * 16 + I * 4 + 0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t Gn = G + (uint32_t)(2 * j + g);
        const int buf = Gn % 3;
        const int bi = Gn % 6;
        const uint32_t par = (Gn / 6) & 1;
        const uint32_t t_s = t_row + X_COL_OF(buf);
        const int kv0 = j * BKV;
        const bool tile_full = __all_sync(0xffffffffu, (kv0 >= lo) && (kv0 + BKV <= hi));
        mbar_wait(&s_full[bi], par);
        tc_fence_after();
        const bool tr = (I == 0 && qd == 0 && lane == 0 && j < 32);
        const int tb = 128 + g * 192 + j * 6;
        if (tr) att4_stamp(a, tb + 0);
        uint32_t r[4][32];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld32(t_s + ch * 32, r[ch]);
        tmem_ld_wait();
        if (tr) att4_stamp(a, tb + 1);
        if (!tile_full) {
          // the common partial tile is the LAST key tile of a bidirectional (BERT) sequence: no lower bound inside the tile and the
          // same upper bound for every row.  Then whole 32-column chunks are either kept, dropped or (one of them) compared
          // element by element - warp-uniform branches instead of 128 two-sided compares per thread.
          const int nvalid = hi - kv0;
          const int nvalid0 = __shfl_sync(0xffffffffu, nvalid, 0);  // (outside the && below: every lane must take part)
          const bool simple = __all_sync(0xffffffffu, lo <= kv0 && nvalid == nvalid0);
          if (simple) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              if ((ch + 1) * 32 <= nvalid) continue;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (ch * 32 + i >= nvalid) r[ch][i] = 0xff800000u;  // -inf
            }
          } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int col = kv0 + ch * 32 + i;
                if (!(col >= lo && col < hi)) r[ch][i] = 0xff800000u;  // -inf
              }
          }
        }
        float mx0 = __uint_as_float(r[0][0]), mx1 = __uint_as_float(r[0][1]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int i = (ch == 0 ? 2 : 0); i < 32; i += 4) {
            mx0 = fmax3(mx0, __uint_as_float(r[ch][i]), __uint_as_float(r[ch][i + 1]));
            if (i + 3 < 32) mx1 = fmax3(mx1, __uint_as_float(r[ch][i + 2]), __uint_as_float(r[ch][i + 3]));
          }
        const float mx = fmaxf(mx0, mx1);
        if (tr) att4_stamp(a, tb + 2);
        // move the reference only when needed (warp-uniform decision because TMEM ld/st are warp collectives)
        const bool need = (mx != -INFINITY) && (m_ref == -INFINITY || (mx - m_ref) * c > RESCALE_THRESH);
        float alpha = 1.f;
        bool resc = false;
        if (need) {
          if (m_ref != -INFINITY) { alpha = ex2_approx((m_ref - mx) * c); resc = true; }  // else: O row and l are still ~0
          m_ref = mx;
        }
        if (__any_sync(0xffffffffu, resc)) {
          // rare path: O_g *= alpha in TMEM (alpha = 1 for rows that keep their reference).  PV of this tile's previous
          // step must have retired; PV of this step is not issued before p_ready below.
          mbar_wait(&o_full[g], (uint32_t)(J + j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t t[32];
            tmem_ld32(t_o + hf * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st32(t_o + hf * 32, t);
          }
        }
        l *= alpha;
        const float nmc = (m_ref == -INFINITY) ? 0.f : -m_ref * c;
        const f2_t c2 = pack2(c, c), nmc2 = pack2(nmc, nmc);
        f2_t rs = pack2(0.f, 0.f);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1, e0, e1;
            unpack2(fma2(pack2(__uint_as_float(r[ch][i]), __uint_as_float(r[ch][i + 1])), c2, nmc2), a0, a1);
            if (((i >> 1) & 3) < POLY8 / 2) {
              exp2_poly2<TF32 ? 4 : 3>(a0, a1, e0, e1);  // FMA pipe
            } else {
              e0 = ex2_approx(a0); e1 = ex2_approx(a1);  // MUFU; exp2(-inf) = 0 for masked keys
            }
            rs = add2(rs, pack2(e0, e1));
            if constexpr (TF32) {
              r[ch][i] = __float_as_uint(e0); r[ch][i + 1] = __float_as_uint(e1);
            } else {
              __nv_bfloat162 v2 = __floats2bfloat162_rn(e0, e1);
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&v2);
            }
          }
          if constexpr (TF32) tmem_st32(t_s + ch * 32, r[ch]);
          else tmem_st16(t_s + ch * 16, pk);
        }
        if (tr) att4_stamp(a, tb + 3);
        tmem_st_wait();
        if (tr) att4_stamp(a, tb + 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[bi]);  // one arrival per warp (4 per group)
        if (tr) att4_stamp(a, tb + 5);
        float rs0, rs1;
        unpack2(rs, rs0, rs1);
        l += rs0 + rs1;
      }
      // ---- item epilogue: wait for this tile's last PV, pull O into registers, release O_g, normalise, store
      if (itr) att4_stamp(a, 480 + g * 16 + I * 4 + 1);
      mbar_wait(&o_full[g], (J + (uint32_t)n_kv - 1) & 1);
      tc_fence_after();
      if (itr) att4_stamp(a, 480 + g * 16 + I * 4 + 2);
      float o[HD];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t t[32];
        tmem_ld32(t_o + hf * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[hf * 32 + i] = __uint_as_float(t[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_drained[g]);
      const float inv = 1.f / l;
      if (row < a.S) {
        if (a.out_bf16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.out) + ((long long)b * a.S + row) * a.ldo + h * HD;
#pragma unroll
          for (int i = 0; i < HD; i += 8) {
            uint4 v;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv), p1 = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv), p3 = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
            v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
            v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dst + i) = v;
          }
        } else {
          float* dst = reinterpret_cast<float*>(a.out) + ((long long)b * a.S + row) * a.ldo + h * HD;
#pragma unroll
          for (int i = 0; i < HD; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
        }
      }
      if (itr) att4_stamp(a, 480 + g * 16 + I * 4 + 3);
      G += (uint32_t)(2 * n_kv); J += (uint32_t)n_kv;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <bool TF32, int POLY8>
int launch_att4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Att4Args& a, cudaStream_t st) {
  static bool attr_set = false;
  static int sms = 0;
  constexpr size_t smem = att4_smem_bytes<TF32>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(attention_tc4_kernel<TF32, POLY8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(attention_tc4): %s", cudaGetErrorString(err));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    attr_set = true;
  }
  const int grid = a.n_items < sms ? a.n_items : sms;
  attention_tc4_kernel<TF32, POLY8><<<grid, ATT4_THREADS, smem, st>>>(tq, tk, tv, a);
  return check_launch("attention_tc4");
}

}  // namespace

// Persistent rotating-score-buffer kernel.  Needs S > 128 (two key tiles per item); outputs need 16-byte aligned rows.
extern "C" int mmvid_attention_v6(const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, void* out,
                                  int out_bf16, long long ldo, int B, int H, int S, int S_pad, int mask_kind,
                                  const int* host_prev_rows, int n_prev, int tf32, int poly8, unsigned long long* trace,
                                  cudaStream_t st) {
  Att4Args a{};
  a.trace = trace;
  a.out = out; a.ldo = ldo; a.out_bf16 = out_bf16;
  a.B = B; a.H = H; a.S = S; a.S_pad = S_pad; a.mask_kind = mask_kind; a.n_prev = n_prev;
  for (int i = 0; i < n_prev; ++i) a.prev_rows[i] = host_prev_rows[i];
  a.n_pairs = (S_pad / BQ + 1) / 2;
  a.n_items = B * H * a.n_pairs;
  if (tf32) return poly8 == 2 ? launch_att4<true, 2>(*tq, *tk, *tv, a, st) : launch_att4<true, 0>(*tq, *tk, *tv, a, st);
  return poly8 == 0 ? launch_att4<false, 0>(*tq, *tk, *tv, a, st) : launch_att4<false, 2>(*tq, *tk, *tv, a, st);
}
