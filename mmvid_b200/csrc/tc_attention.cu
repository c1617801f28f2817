// Flash attention for sm_100a (F.scaled_dot_product_attention, clip_model.py:217-222): two 128-query tiles per CTA whose
// score tiles ROTATE through THREE TMEM buffers.
//
// With one S buffer per query tile and P written over S in place, QK^T(j+1) of a tile cannot be issued before PV(j) of
// the same tile: each tile is one serial chain  QK -> ld -> max -> exp -> st -> PV  and the tensor pipe and the MUFU
// pipe take turns (both ~51 % busy, profiles/r1_e_attention_stall_analysis.md).
// Here the 2 x n_kv tile-steps  n = 2 j + g  (g = query tile A/B, j = key tile) use score buffer n mod 3:
//
//   TMEM columns:  X0 [0,128)  X1 [128,256)  X2 [256,384)  O_A [384,448)  O_B [448,512)      (all 512 columns)
//   tensor pipe :  QK(0) QK(1) QK(2) | PV(0) QK(3) | PV(1) QK(4) | PV(2) QK(5) | ...   -> X(n mod 3)
//
// so S of step n+2 (the next step of the SAME softmax group) has been produced while that group is still busy with
// step n: the softmax warpgroups never wait for the tensor pipe in steady state, the tensor pipe only ever waits for
// P(n), and MUFU (16 ex2/clk/SM = 2048 clk per key step for the two tiles) runs back to back with the MMAs
// (2064 clk per key step in kind::tf32, 1032 in kind::f16).  PV(n) and QK(n+3) are issued by the same thread in that
// order and the tensor pipe executes in issue order, so QK(n+3) overwrites X(n mod 3) only after PV(n) has read P(n).
//
// Because MUFU is then the binding pipe (tf32: tie, 16-bit: 2x the MMA time), POLY8 of every 8 exponentials are
// evaluated on the FMA pipe instead (Cody-Waite range reduction + degree 4 / 3 polynomial, packed f32x2 FMAs; the
// FlashAttention-4 trick): relative error 2.7e-6 (tf32 path; P is truncated to 10 mantissa bits by the MMA anyway)
// / 7.5e-5 (16-bit path).
//
// Warp roles (384 threads): warp 0 = TMA producer for Q and K, warp 2 = TMA producer for V^T (independent rings: K
// tiles are consumed ~1.5 key steps ahead of V tiles), warps 1 and 3 = MMA issuers (one per query tile); warps 4-7 / 8-11 = softmax
// groups A / B with 224 registers (setmaxnreg), one query row per thread, whole 128-column row in registers.
// Lazy rescale (FA4): O stays in TMEM relative to a reference max that rarely moves; masks are analytic (per-row key
// interval), never a dense [S,S] tensor.
#include "common.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

using namespace mmvid;
using namespace mmvid::tc;

namespace {

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}

constexpr int ATT3_THREADS = 384;
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

struct Att3Args {
  unsigned long long* trace;  // debug timeline (mmvid_debug_attention_trace), normally null
  void* out; long long ldo; int out_h16;  // 1: 16-bit output in the kernel's own 16-bit flavour (bf16 | fp16)
  int B, H, S, S_pad, mask_kind;
  int prev_rows[4]; int n_prev;
  int spin;  // 1: the MMA threads poll p_ready with test_wait instead of try_wait
  int dual;  // 1: two MMA-issuing threads (warps 1 and 3), one per query tile; 0: warp 1 issues everything
  int spec;  // 16-bit kinds: 1 = speculative exponentials (no row max in the common case), see the softmax warps
};

template <bool TRACE>
__device__ __forceinline__ void stamp6(unsigned long long* trace, int idx) {
  if constexpr (TRACE) {
    if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) trace[idx] = clock64();
  }
}

template <bool TF32>
constexpr size_t att3_smem_bytes() {
  // Q_A, Q_B, 2 x K, 2 x V^T tiles (each BQ*HD elements) + barriers + alignment slack
  return (size_t)6 * (TF32 ? 32768 : 16384) + 1024 + 512;
}

// F16 (16-bit kinds only): fp16 operands / P / 16-bit output instead of bf16 - the tf32 mantissa at the kind::f16 rate.
// P <= 2^8 (lazy rescale) and softmax inputs of O(10) sit comfortably inside the fp16 range.
// SPEC (16-bit kinds): speculative exponentials, see the softmax warps
template <bool TF32, int POLY8, bool F16, bool TRACE, bool SPEC = false>
__global__ void __launch_bounds__(ATT3_THREADS, 1) attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV, Att3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]  (count 2: QK of tile A and of tile B both read the stage)
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]  (count 2)
  // s_full / p_ready are indexed by n mod 6 (the score BUFFER is n mod 3): barrier i then only ever serves one softmax
  // group and one MMA thread, which observe its completions one after the other.  With one barrier per buffer a waiter
  // would see every OTHER completion - always the same parity - and mistake the previous phase for its own.
  uint64_t* s_full = bars + 9;    // [6] QK(n) landed in X(n mod 3)
  uint64_t* p_ready = bars + 15;  // [6] P(n) written over it (count 4: one arrival per warp)
  // o_full[g][j & 1]: PV of query tile g, key step j retired.  TWO barriers per tile (even / odd key steps): the softmax
  // group only looks at them on the rare rescale path, by which time a single barrier could be one OR two phases ahead
  // of the last phase the group saw - indistinguishable for a parity wait.  Each of the two is at most one phase ahead.
  uint64_t* o_full = bars + 21;   // [2][2]
  uint64_t* all_done = bars + 25;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 34);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 512 + 1023) & ~(uintptr_t)1023);
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int BKE = 128 / ESZ;
  constexpr int QK_KB = HD / BKE;   // 2 | 1
  constexpr int PV_KB = BKV / BKE;  // 4 | 2
  constexpr int T_BYTES = BQ * HD * ESZ;  // every tile (Q, K, V^T) has the same byte size
  // tile addresses by arithmetic (pointer arrays indexed by a run-time stage would live in local memory)
  auto sQ = [&](int g) { return tiles + g * T_BYTES; };
  auto sK = [&](int st) { return tiles + (2 + st) * T_BYTES; };
  auto sV = [&](int st) { return tiles + (4 + st) * T_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int bh = blockIdx.y;
  const int b = bh / a.H, h = bh - b * a.H;
  int n_kv = (a.S + BKV - 1) / BKV;
  if (a.mask_kind == MMVID_MASK_CAUSAL) n_kv = min(n_kv, (q0 + 2 * BQ - 1) / BKV + 1);
  const int n_steps = 2 * n_kv;  // tile-steps n = 2 j + g

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
      mbar_init(&o_full[2 * i], 1); mbar_init(&o_full[2 * i + 1], 1);
    }
    for (int i = 0; i < 6; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); }
    mbar_init(all_done, a.dual ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  chain_release();
  chain_wait();  // Q / K / V^T come from the previous kernel; everything above overlapped its tail

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ Q + K producer
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * T_BYTES);
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
          tma_load_2d(sQ(g) + kb * (BQ * 128), &tmQ, q_full, kb * BKE, bh * a.S_pad + q0 + g * BQ);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], T_BYTES);
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
          tma_load_2d(sK(st) + kb * (BKV * 128), &tmK, &k_full[st], kb * BKE, bh * a.S_pad + j * BKV);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ V^T producer
    if (elect_one()) {
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], T_BYTES);
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb)
          tma_load_2d(sV(st) + kb * (HD * 128), &tmV, &v_full[st], j * BKV + kb * BKE, bh * HD);
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------ MMA issuers
    // dual = 1: warp 1 owns the even tile-steps (PV of query tile A and the QK^T that recycles the same score buffer,
    // which belongs to tile B), warp 3 the odd ones.  The r1l timeline showed ONE issuing thread to be the bottleneck:
    // ~1000 clk of back-pressured MMA issue per tile-step plus ~600 clk of its own barrier round trips and descriptor
    // arithmetic, during which the tensor pipe drains.  With two threads each one's overhead hides behind the other's
    // MMAs.  Ordering still holds: PV(n) and QK(n+3) (same score buffer) come from the same thread, in that order.
    const int my = a.dual ? (warp == 3 ? 1 : 0) : 0;
    const int stride = a.dual ? 2 : 1;
    if ((a.dual || warp == 1) && elect_one()) {
      constexpr uint32_t idesc_qk = TF32 ? make_idesc<true>(BQ, BKV) : make_idesc_h16(F16, BQ, BKV);
      constexpr uint32_t idesc_pv = TF32 ? make_idesc<true>(BQ, HD) : make_idesc_h16(F16, BQ, HD);
      constexpr uint32_t TB16 = T_BYTES >> 4;  // descriptor address units (16 B) per tile
      // all shared-memory descriptors are  base + small multiples: the address field cannot carry (smem < 256 KB)
      const uint64_t dQ0 = make_smem_desc_sw128(smem_u32(sQ(0)));
      const uint64_t dK0 = make_smem_desc_sw128(smem_u32(sK(0)));
      const uint64_t dV0 = make_smem_desc_sw128(smem_u32(sV(0)));
      // QK(n): S = Q_g K_j^T into score buffer `buf` (= n mod 3)
      auto issue_qk = [&](int n, uint32_t buf) {
        const int j = n >> 1, g = n & 1, st = j & 1;
        const uint32_t x = tmem_base + buf * 128;
        const uint64_t qd = dQ0 + (uint64_t)(g * TB16), kd = dK0 + (uint64_t)(st * TB16);
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ss<TF32>(x, qd + (uint64_t)(kb * (BQ * 128 / 16) + kk * 2), kd + (uint64_t)(kb * (BKV * 128 / 16) + kk * 2),
                         idesc_qk, (kb | kk) != 0);
        tc_commit(&k_empty[st]);
        tc_commit(&s_full[n % 6]);
      };
      // PV(n): O_g (+)= P(n) V_j with A = P straight from TMEM (score buffer `buf`), B = V^T tile
      auto issue_pv = [&](int n, uint32_t buf) {
        const int j = n >> 1, g = n & 1, st = j & 1;
        const uint32_t x = tmem_base + buf * 128, o = tmem_base + O_COL_OF(0) + g * 64;
        const uint64_t vd = dV0 + (uint64_t)(st * TB16);
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // O accumulates in TMEM over ALL key tiles (lazy rescale, see softmax warps)
            mma_ts<TF32>(o, x + kb * 32 + kk * 8, vd + (uint64_t)(kb * (HD * 128 / 16) + kk * 2), idesc_pv,
                         (j != 0 || (kb | kk) != 0) ? 1u : 0u);
        tc_commit(&v_empty[st]);
        tc_commit(&o_full[2 * g + (j & 1)]);
      };
      if (warp == 1) {  // prologue: the first three QK^T fill the three score buffers
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 1);
        if (n_steps > 2) {
          mbar_wait(&k_full[1], 0);
          tc_fence_after();
          issue_qk(2, 2);
        }
      } else {
        mbar_wait(q_full, 0);  // warp 3's first QK^T (step 4) reads Q_A
      }
      uint32_t buf = (uint32_t)my;
      for (int n = my; n < n_steps; n += stride) {
        const int j = n >> 1, n3 = n + 3, j3 = n3 >> 1;
        const bool more = n3 < n_steps;
        // V_j landed long ago in steady state: this wait returns at once and sits in front of the P(n) wait
        mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        const uint32_t par = (uint32_t)(n / 6) & 1;
        if (a.spin) mbar_wait_spin(&p_ready[n % 6], par);  // P(n) written (and O_g rescaled if it had to be)
        else mbar_wait(&p_ready[n % 6], par);
        tc_fence_after();
        if (n < 64) stamp6<TRACE>(a.trace, n * 2);
        issue_pv(n, buf);
        if (more) {
          // K tile of step n+3: its wait hides behind PV(n), which is already queued.  (Placed after the P(n) wait on
          // purpose: P(n) implies that the previous phase of this K stage completed, so the parity test cannot alias
          // for warp 3, which never consumed key tiles 0 and 1 itself.)
          mbar_wait(&k_full[j3 & 1], (j3 >> 1) & 1);
          tc_fence_after();
          issue_qk(n3, buf);  // the buffer is free: PV(n) (same issue stream) consumes P(n) first
        }
        if (n < 64) stamp6<TRACE>(a.trace, n * 2 + 1);
        buf += stride;
        if (buf >= 3) buf -= 3;
      }
      tc_commit(all_done);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = (warp - 4) >> 2;  // 0: tile A, 1: tile B
    const int qd = warp & 3;
    const int row_local = qd * 32 + lane;
    const int row = q0 + g * BQ + row_local;
    const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t t_o = t_row + O_COL_OF(g);
    int lo = 0, hi = a.S;
    if (a.mask_kind == MMVID_MASK_CAUSAL) hi = min(a.S, row + 1);
    else if (a.mask_kind == MMVID_MASK_PREV) {
      for (int i = 0; i < a.n_prev; ++i) if (a.prev_rows[i] == row) lo = row;
    }
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e): scores are handled in the log2 domain
    // Lazy rescale (FA4): O_g accumulates in TMEM across key tiles relative to a reference max m_ref that only moves
    // when the running row max exceeds it by more than 2^8 (P <= 256 is harmless in fp32/tf32/bf16).
    constexpr float RESCALE_THRESH = 8.f;
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int n = 2 * j + g;
      const int buf = n % 3;
      const int bi = n % 6;
      const uint32_t par = (uint32_t)(n / 6) & 1;
      const uint32_t t_s = t_row + X_COL_OF(buf);
      const int kv0 = j * BKV;
      const bool tile_full = __all_sync(0xffffffffu, (kv0 >= lo) && (kv0 + BKV <= hi));
      mbar_wait(&s_full[bi], par);
      tc_fence_after();
      const bool tr = (qd == 0 && lane == 0 && j < 32);
      const int tb = 128 + g * 192 + j * 6;
      if (tr) stamp6<TRACE>(a.trace, tb + 0);
      uint32_t r[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(t_s + ch * 32, r[ch]);
      tmem_ld_wait();
      if (tr) stamp6<TRACE>(a.trace, tb + 1);
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
      // exponentials of the whole row against reference `nmc` (= -m_ref c), P stored over S, returns the row sum
      auto exp_and_store = [&](float nmc) -> float {
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
              pk[i >> 1] = pack_h16<F16>(e0, e1);
            }
          }
          if constexpr (TF32) tmem_st32(t_s + ch * 32, r[ch]);
          else tmem_st16(t_s + ch * 16, pk);
        }
        float rs0, rs1;
        unpack2(rs, rs0, rs1);
        return rs0 + rs1;
      };
      // SPEC (16-bit kinds, where the score row survives in registers): from the second key tile on the exponentials are
      // taken against the CURRENT reference straight away - no row maximum (98 FMNMX and a dependent phase of ~280 clk per
      // step), no rescale test.  A row sum <= 2^12 proves that no P exceeded 2^12 (fp16-safe, O and l in fp32); otherwise
      // (new maximum far above the reference, or a row whose reference is still -inf) the warp falls back to the exact path
      // below, which moves the reference and recomputes the row.  Effective lazy-rescale threshold: between 2^8 and 2^12.
      float rsum = 0.f;
      bool exact = true;
      if constexpr (SPEC && !TF32) {
        if (j > 0) {
          rsum = exp_and_store((m_ref == -INFINITY) ? 0.f : -m_ref * c);
          exact = __any_sync(0xffffffffu, !(rsum <= 4096.f));
        }
      }
      if (exact) {
        float mx0 = __uint_as_float(r[0][0]), mx1 = __uint_as_float(r[0][1]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int i = (ch == 0 ? 2 : 0); i < 32; i += 4) {
            mx0 = fmax3(mx0, __uint_as_float(r[ch][i]), __uint_as_float(r[ch][i + 1]));
            if (i + 3 < 32) mx1 = fmax3(mx1, __uint_as_float(r[ch][i + 2]), __uint_as_float(r[ch][i + 3]));
          }
        const float mx = fmaxf(mx0, mx1);
        if (tr) stamp6<TRACE>(a.trace, tb + 2);
        // move the reference only when needed (warp-uniform decision because TMEM ld/st are warp collectives)
        const bool need = (mx != -INFINITY) && (m_ref == -INFINITY || (mx - m_ref) * c > RESCALE_THRESH);
        float alpha = 1.f;
        bool resc = false;
        if (need) {
          if (m_ref != -INFINITY) { alpha = ex2_approx((m_ref - mx) * c); resc = true; }  // else: O row and l are still ~0
          m_ref = mx;
        }
        if (__any_sync(0xffffffffu, resc)) {
          // rare path: O_g *= alpha in TMEM (alpha = 1 for rows that keep their reference).
          // PV_g(j-1) must have retired before O is read-modify-written.  It commits to barrier (j-1) & 1 of this tile as that
          // barrier's phase (j-1) >> 1.  s_full(n) observed => QK(n-3) retired => PV(n-6) = PV_g(j-3), issued earlier by the
          // same thread, retired: the barrier has completed every phase before this one, and cannot be past it (PV_g(j+1)
          // needs P of this very step), so the parity wait is unambiguous however long ago the group last looked.
          mbar_wait(&o_full[2 * g + ((j - 1) & 1)], (uint32_t)((j - 1) >> 1) & 1u);
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
        // (speculative pass already stored this row against the same reference: nothing to redo unless a reference moved)
        if (!(SPEC && !TF32) || j == 0 || __any_sync(0xffffffffu, need))
          rsum = exp_and_store((m_ref == -INFINITY) ? 0.f : -m_ref * c);
      }
      if (tr) stamp6<TRACE>(a.trace, tb + 3);
      tmem_st_wait();
      if (tr) stamp6<TRACE>(a.trace, tb + 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[bi]);  // one arrival per warp (4 per group)
      if (tr) stamp6<TRACE>(a.trace, tb + 5);
      if (lane == 0 && j < 32) stamp6<TRACE>(a.trace, 512 + (g * 4 + qd) * 32 + j);  // per-warp arrival (skew between the 4 warps)
      l += rsum;
    }
    // every MMA of BOTH groups must have retired before K/V smem is recycled as the output staging area
    mbar_wait(all_done, 0);
    tc_fence_after();
    float o[HD];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t t[32];
      tmem_ld32(t_o + hf * 32, t);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[hf * 32 + i] = __uint_as_float(t[i]);
    }
    const float inv = 1.f / l;
    uint8_t* stage_base = g == 0 ? sK(0) : sV(0);  // 2 contiguous tiles each: >= 128 x 68 floats
    const int q_tile0 = q0 + g * BQ;
    if (a.out_h16) {
      constexpr int LD = HD + 8;
      uint16_t* st = reinterpret_cast<uint16_t*>(stage_base) + (size_t)row_local * LD;
#pragma unroll
      for (int i = 0; i < HD; i += 2)
        *reinterpret_cast<uint32_t*>(st + i) = pack_h16<F16>(o[i] * inv, o[i + 1] * inv);
      __syncwarp();
      uint16_t* outp = reinterpret_cast<uint16_t*>(a.out);
      for (int r0 = 0; r0 < 32; r0 += 4) {
        const int rl = qd * 32 + r0 + (lane >> 3);
        const int s = q_tile0 + rl;
        if (s < a.S) {
          const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<uint16_t*>(stage_base) + (size_t)rl * LD + (lane & 7) * 8);
          *reinterpret_cast<uint4*>(outp + ((long long)b * a.S + s) * a.ldo + h * HD + (lane & 7) * 8) = v;
        }
      }
    } else {
      // tf32: two 32 KB K (V) tiles hold 128 x 68 floats; bf16 operands leave 2 x 16 KB = exactly 128 x 64 floats
      // (fp32 output from the bf16 kernel is not a model path: unpadded rows, bank conflicts accepted)
      constexpr int LD = TF32 ? HD + 4 : HD;
      float* st = reinterpret_cast<float*>(stage_base) + (size_t)row_local * LD;
#pragma unroll
      for (int i = 0; i < HD; i += 4)
        *reinterpret_cast<float4*>(st + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      __syncwarp();
      float* outp = reinterpret_cast<float*>(a.out);
      for (int r0 = 0; r0 < 32; r0 += 2) {
        const int rl = qd * 32 + r0 + (lane >> 4);
        const int s = q_tile0 + rl;
        if (s < a.S) {
          const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<float*>(stage_base) + (size_t)rl * LD + (lane & 15) * 4);
          *reinterpret_cast<float4*>(outp + ((long long)b * a.S + s) * a.ldo + h * HD + (lane & 15) * 4) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <bool TF32, int POLY8, bool F16, bool TRACE, bool SPEC = false>
int launch_att3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Att3Args& a, cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t smem = att3_smem_bytes<TF32>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(attention_tc3_kernel<TF32, POLY8, F16, TRACE, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(attention_tc3): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  dim3 grid((a.S_pad / BQ + 1) / 2, a.B * a.H);
  cudaError_t err = launch_chained(attention_tc3_kernel<TF32, POLY8, F16, TRACE, SPEC>, grid, dim3(ATT3_THREADS), smem, st, tq, tk, tv, a);
  if (err != cudaSuccess) return fail(MMVID_ECUDA, "attention_tc3 launch: %s", cudaGetErrorString(err));
  return check_launch("attention_tc3");
}

// The timeline build (TRACE) exists for the default POLY8 = 2 only.
template <bool TF32, bool F16>
int launch_att3_poly(int poly8, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Att3Args& a,
                     cudaStream_t st) {
  if (a.trace != nullptr) {
    constexpr int DP = 2;
    if (poly8 != DP) return fail(MMVID_EINVAL, "attention_tc3: the trace build exists for the default poly8 only%s", "");
    return launch_att3<TF32, DP, F16, true>(tq, tk, tv, a, st);
  }
  if (poly8 == 0) return launch_att3<TF32, 0, F16, false>(tq, tk, tv, a, st);
  if constexpr (!TF32) {
    if (poly8 == 2 && a.spec) return launch_att3<TF32, 2, F16, false, true>(tq, tk, tv, a, st);
  }
  if (poly8 == 2) return launch_att3<TF32, 2, F16, false>(tq, tk, tv, a, st);
  if (poly8 == 4) return launch_att3<TF32, 4, F16, false>(tq, tk, tv, a, st);
  return fail(MMVID_EINVAL, "attention_tc3: poly8 must be 0, 2 or 4%s", "");
}

}  // namespace

// Rotating-score-buffer kernel.  poly8: how many of every 8 exponentials run on the FMA pipe (0, 2, 4); spin: the
// MMA thread polls p_ready.  trace: see mmvid_debug_attention_trace below.
// kind: MMVID_TF32 | MMVID_BF16 | MMVID_F16
static int mmvid_attention_v5(const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, void* out,
                              int out_h16, long long ldo, int B, int H, int S, int S_pad, int mask_kind,
                              const int* host_prev_rows, int n_prev, int kind, int poly8, int spin, int dual,
                              unsigned long long* trace, cudaStream_t st) {
  Att3Args a{};
  a.trace = trace;
  a.out = out; a.ldo = ldo; a.out_h16 = out_h16;
  a.B = B; a.H = H; a.S = S; a.S_pad = S_pad; a.mask_kind = mask_kind; a.n_prev = n_prev;
  a.spin = spin; a.dual = dual;
  a.spec = env_int("MMVID_ATT_SPEC", 1);
  for (int i = 0; i < n_prev; ++i) a.prev_rows[i] = host_prev_rows[i];
  if (kind == MMVID_TF32) return launch_att3_poly<true, false>(poly8, *tq, *tk, *tv, a, st);
  if (kind == MMVID_F16) return launch_att3_poly<false, true>(poly8, *tq, *tk, *tv, a, st);
  return launch_att3_poly<false, false>(poly8, *tq, *tk, *tv, a, st);
}

namespace mmvid { unsigned long long* g_att_trace = nullptr; }
// Debug / profiling hook: CTA (0,0) of every following attention launch writes clock64() stamps of its pipeline events
// into `dev_buf` (>= 1024 entries; pass NULL to switch it off).  MMA threads: [2 n] = P(n) observed, [2 n + 1] = PV(n) /
// QK(n+3) issued, n = 2 j + g < 64; softmax warp 0 of tile g: [128 + g*192 + j*6 + {0: S ready, 1: S in registers,
// 2: row max, 3: exps done / P stores issued, 4: P stores landed, 5: p_ready signalled}]; every softmax warp (tile g,
// lane quarter qd): [512 + (g*4 + qd)*32 + j] = its p_ready arrival at key step j < 32.  v6 (persistent) stamps the first
// item of CTA 0 there and adds per item it < 8 of tile g: [768 + g*32 + it*4 + {0: first S seen, 1: last P signalled,
// 2: O in registers, 3: bulk store issued}].
extern "C" int mmvid_debug_attention_trace(unsigned long long* dev_buf) {
  mmvid::g_att_trace = dev_buf;
  return MMVID_OK;
}

extern "C" int mmvid_attention(const void* q, const void* k, const void* vt, void* out, int out_dtype, long long ldo,
                               int B, int H, int S, int S_pad, int mask_kind, const int* host_prev_rows, int n_prev,
                               int precision, mmvid_stream_t stream) {
  MMVID_REQUIRE(precision == MMVID_TF32 || precision == MMVID_BF16 || precision == MMVID_F16,
                "tensor-core attention needs TF32, BF16 or F16");
  MMVID_REQUIRE(S_pad % 128 == 0 && S_pad >= S && S > 0, "S_pad multiple of 128");
  MMVID_REQUIRE(n_prev >= 0 && n_prev <= 4, "at most 4 mask_prev rows");
  MMVID_REQUIRE((long long)B * H <= 65535, "B*H <= 65535");
  const bool tf32 = precision == MMVID_TF32;
  const int esz = tf32 ? 4 : 2, dt = tf32 ? MMVID_DT_F32 : (precision == MMVID_F16 ? MMVID_DT_F16 : MMVID_DT_BF16);
  MMVID_REQUIRE(out_dtype == MMVID_DT_F32 || (!tf32 && out_dtype == dt), "output: fp32 or the precision's own 16-bit type");
  const uint32_t BKE = 128 / esz;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[2] = {64, (uint64_t)B * H * S_pad};
    uint64_t str[1] = {(uint64_t)64 * esz};
    uint32_t box[2] = {BKE, 128};
    int rc = make_tensor_map(&tq, q, dt, 2, dims, str, box);
    if (rc) return rc;
    rc = make_tensor_map(&tk, k, dt, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)S_pad, (uint64_t)B * H * 64};
    uint64_t str[1] = {(uint64_t)S_pad * esz};
    uint32_t box[2] = {BKE, 64};
    int rc = make_tensor_map(&tv, vt, dt, 2, dims, str, box);
    if (rc) return rc;
  }
  // Measured default (gpurun r2r, trace-free builds): 2 of every 8 exponentials on the FMA pipe in every kind (tf32: 98.3 us
  // against 102.4 us with all of them on MUFU).  MMVID_ATT_POLY (0|2|4 of every 8 exponentials on the FMA pipe),
  // MMVID_ATT_SPEC (0|1 speculative exponentials in the 16-bit kinds), MMVID_ATT_SPIN (0|1) and MMVID_ATT_DUAL (0|1: one or
  // two MMA-issuing threads) are tuning switches.
  return mmvid_attention_v5(&tq, &tk, &tv, out, out_dtype != MMVID_DT_F32, ldo, B, H, S, S_pad, mask_kind, host_prev_rows,
                            n_prev, precision, env_int("MMVID_ATT_POLY", 2), env_int("MMVID_ATT_SPIN", 0),
                            env_int("MMVID_ATT_DUAL", 1), mmvid::g_att_trace,
                            to_stream(stream));
}
