// Flash-style multi-head attention core on tcgen05 tensor cores (head_dim 64, CLIP ViT-B/32 stack):
//     out[b, s, h*64:(h+1)*64] = softmax(Q K^T / 8 + mask) V          (clip_model.py:217-222 -> SDPA)
// One CTA = one 128-query tile of one (batch, head); 2 CTAs co-resident per SM (one's softmax overlaps the
// other's MMAs).  Roles:
//   warp 0      TMA producer: Q once, then K_j / V^T_j tiles (SWIZZLE_128B K-major boxes)
//   warp 1      MMA issuer:   S = Q K_j^T  (SS, 128x128, fp32 accumulate in TMEM cols [0,128))
//                             O_j = P_j V_j (TS: A = P read straight from TMEM, B = V^T tile; TMEM cols [128,192))
//   warps 2..5  softmax: one query row per thread (TMEM lane == row): tcgen05.ld S, online max / exp2 /
//               row-sum in fp32, P written back IN PLACE over S with tcgen05.st (tf32: 1 value / column,
//               bf16: 2 packed / column), then O_reg = O_reg * alpha + O_j from TMEM.
// The mask is analytic - per query row a visible key interval [lo, hi): causal: hi = row+1; mask_prev
// (BERT): lo = row for the two special rows [ST1]/[VID] (clip_model.py:571-575); padding keys >= S are
// cut by hi <= S.  No [S,S] mask tensor is ever read (the reference copies a dense fp32 one per layer).
// V is consumed transposed (V^T[d, s], written by the QKV split) so every operand is K-major.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

using namespace mmvid;
using namespace mmvid::tc;

namespace {

constexpr int ATT_THREADS = 192;
constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;

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
// (x0, x1) * s + b with one packed FFMA2
__device__ __forceinline__ void ffma2(float x0, float x1, float s, float b, float& y0, float& y1) {
  unsigned long long px, ps, pb, py;
  asm("mov.b64 %0, {%1, %2};" : "=l"(px) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(py) : "l"(px), "l"(ps), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(py));
}

struct AttArgs {
  void* out; long long ldo; int out_bf16;
  int B, H, S, S_pad, mask_kind;
  int prev_rows[4]; int n_prev;
};

template <bool TF32>
constexpr size_t att_smem_bytes() {
  return (size_t)(TF32 ? 3 * 32768 : 3 * 16384) + 1024 + 256;
}

template <bool TF32>
__global__ void __launch_bounds__(ATT_THREADS, 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV, AttArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;   // QK_j complete
  uint64_t* p_ready = bars + 6;  // 128 softmax threads wrote P_j
  uint64_t* o_full = bars + 7;   // PV_j complete
  uint64_t* o_free = bars + 8;   // 128 softmax threads consumed O_j
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~(uintptr_t)1023);
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int BKE = 128 / ESZ;                 // elements per 128-byte k-block (32 | 64)
  constexpr int QK_KB = HD / BKE;                // k-blocks over head_dim (2 | 1)
  constexpr int PV_KB = BKV / BKE;               // k-blocks over the kv tile (4 | 2)
  constexpr int Q_BYTES = BQ * HD * ESZ, K_BYTES = BKV * HD * ESZ, V_BYTES = HD * BKV * ESZ;
  uint8_t* sQ = tiles;
  uint8_t* sK = tiles + Q_BYTES;
  uint8_t* sV = sK + K_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int bh = blockIdx.y;  // b * H + h
  const int b = bh / a.H, h = bh - b * a.H;
  int n_kv = (a.S + BKV - 1) / BKV;
  if (a.mask_kind == MMVID_MASK_CAUSAL) n_kv = min(n_kv, (q0 + BQ - 1) / BKV + 1);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1); mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int kb = 0; kb < QK_KB; ++kb) tma_load_2d(sQ + kb * (BQ * 128), &tmQ, q_full, kb * BKE, bh * a.S_pad + q0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_empty, ph ^ 1);
        mbar_expect_tx(k_full, K_BYTES);
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
          tma_load_2d(sK + kb * (BKV * 128), &tmK, k_full, kb * BKE, bh * a.S_pad + j * BKV);
        mbar_wait(v_empty, ph ^ 1);
        mbar_expect_tx(v_full, V_BYTES);
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb)
          tma_load_2d(sV + kb * (HD * 128), &tmV, v_full, j * BKV + kb * BKE, bh * HD);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc<TF32>(BQ, BKV);
      constexpr uint32_t idesc_pv = make_idesc<TF32>(BQ, HD);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_full, ph);
        if (j > 0) mbar_wait(o_free, ph ^ 1);  // softmax threads are done with O_{j-1} (and S/P_{j-1})
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb) {
          const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ + kb * (BQ * 128)));
          const uint64_t kd = make_smem_desc_sw128(smem_u32(sK + kb * (BKV * 128)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ss<TF32>(tmem_base, desc_advance(qd, kk * 32), desc_advance(kd, kk * 32), idesc_qk, (kb | kk) != 0);
        }
        tc_commit(k_empty);
        tc_commit(s_full);
        mbar_wait(v_full, ph);
        mbar_wait(p_ready, ph);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb) {
          const uint64_t vd = make_smem_desc_sw128(smem_u32(sV + kb * (HD * 128)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ts<TF32>(tmem_base + O_COL, tmem_base + kb * 32 + kk * 8, desc_advance(vd, kk * 32), idesc_pv,
                         (kb | kk) != 0);
        }
        tc_commit(v_empty);
        tc_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output (warps 2..5)
    const int qd = warp & 3;
    const int row_local = qd * 32 + lane;
    const int row = q0 + row_local;
    const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16);
    int lo = 0, hi = a.S;
    if (a.mask_kind == MMVID_MASK_CAUSAL) hi = min(a.S, row + 1);
    else if (a.mask_kind == MMVID_MASK_PREV) {
      for (int i = 0; i < a.n_prev; ++i) if (a.prev_rows[i] == row) lo = row;
    }
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m = -INFINITY, l = 0.f;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int kv0 = j * BKV;
      // warp-uniform: does every row of this warp see all 128 keys of the tile?  (true for all interior tiles of
      // the bidirectional BERT mask; false only on the padded last tile, causal diagonals and the two special rows)
      const bool tile_full = __all_sync(0xffffffffu, (kv0 >= lo) && (kv0 + BKV <= hi));
      mbar_wait(s_full, ph);
      tc_fence_after();
      float mx = -INFINITY;
      // ---- pass 1: row max (64 columns in flight per wait)
#pragma unroll
      for (int ch = 0; ch < BKV / 32; ++ch) {
        uint32_t r0[32];
        tmem_ld32(t_row + ch * 32, r0);
        tmem_ld_wait();
        if (tile_full) {
          float m0 = __uint_as_float(r0[0]), m1 = __uint_as_float(r0[1]);
#pragma unroll
          for (int i = 2; i < 32; i += 4) {
            m0 = fmax3(m0, __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
            if (i + 3 < 32) m1 = fmax3(m1, __uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3]));
          }
          mx = fmax3(mx, m0, m1);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c0 = kv0 + ch * 32 + i;
            mx = fmaxf(mx, (c0 >= lo && c0 < hi) ? __uint_as_float(r0[i]) : -INFINITY);
          }
        }
      }
      const float m_new = fmaxf(m, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_approx((m - m_use) * c);  // m = -inf -> 0
      const float nmc = -m_use * c;
      float rs0 = 0.f, rs1 = 0.f;
      // ---- pass 2: P = exp2(s*c - m*c), written back in place
#pragma unroll 1
      for (int ch = 0; ch < BKV / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(t_row + ch * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1;
          ffma2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), c, nmc, a0, a1);
          float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
          if (!tile_full) {
            const int col = kv0 + ch * 32 + i;
            e0 = (col >= lo && col < hi) ? e0 : 0.f;
            e1 = (col + 1 >= lo && col + 1 < hi) ? e1 : 0.f;
          }
          if constexpr (TF32) {
            rs0 += e0; rs1 += e1;
            r[i] = __float_as_uint(e0); r[i + 1] = __float_as_uint(e1);
          } else {
            __nv_bfloat162 v2 = __floats2bfloat162_rn(e0, e1);
            const uint32_t w = *reinterpret_cast<uint32_t*>(&v2);
            pk[i >> 1] = w;
            // row sum from the ROUNDED probabilities so numerator and denominator agree
            rs0 += __uint_as_float(w << 16);
            rs1 += __uint_as_float(w & 0xffff0000u);
          }
        }
        if constexpr (TF32) tmem_st32(t_row + ch * 32, r);
        else tmem_st16(t_row + ch * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      l = l * alpha + (rs0 + rs1);
      m = m_new;
      // ---- O_j
      mbar_wait(o_full, ph);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r0[32];
        tmem_ld32(t_row + O_COL + hf * 32, r0);
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          mbar_arrive(o_free);  // O_j is in registers: the next QK^T may overwrite S/P, the next PV may overwrite O
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) o[hf * 32 + i] = fmaf(o[hf * 32 + i], alpha, __uint_as_float(r0[i]));
      }
    }
    // ---- normalise, stage through (dead) tile smem, coalesced store
    const float inv = 1.f / l;
    // all MMAs have completed (o_full of the last tile) => Q/K/V smem is dead
    if (a.out_bf16) {
      constexpr int LD = HD + 8;
      __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(tiles) + (size_t)row_local * LD;
#pragma unroll
      for (int i = 0; i < HD; i += 2) {
        __nv_bfloat162 v2 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv);
        *reinterpret_cast<__nv_bfloat162*>(st + i) = v2;
      }
      __syncwarp();
      // 8 lanes x 16 B per row, 4 rows per instruction
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out);
      for (int r0 = 0; r0 < 32; r0 += 4) {
        const int rl = qd * 32 + r0 + (lane >> 3);
        const int s = q0 + rl;
        if (s < a.S) {
          const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<__nv_bfloat16*>(tiles) + (size_t)rl * LD + (lane & 7) * 8);
          *reinterpret_cast<uint4*>(outp + ((long long)b * a.S + s) * a.ldo + h * HD + (lane & 7) * 8) = v;
        }
      }
    } else {
      constexpr int LD = HD + 4;
      float* st = reinterpret_cast<float*>(tiles) + (size_t)row_local * LD;
#pragma unroll
      for (int i = 0; i < HD; i += 4)
        *reinterpret_cast<float4*>(st + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      __syncwarp();
      float* outp = reinterpret_cast<float*>(a.out);
      for (int r0 = 0; r0 < 32; r0 += 2) {
        const int rl = qd * 32 + r0 + (lane >> 4);
        const int s = q0 + rl;
        if (s < a.S) {
          const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<float*>(tiles) + (size_t)rl * LD + (lane & 15) * 4);
          *reinterpret_cast<float4*>(outp + ((long long)b * a.S + s) * a.ldo + h * HD + (lane & 15) * 4) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <bool TF32>
int launch_att(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t smem = att_smem_bytes<TF32>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(attention_tc_kernel<TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  dim3 grid(a.S_pad / BQ, a.B * a.H);
  attention_tc_kernel<TF32><<<grid, ATT_THREADS, smem, st>>>(tq, tk, tv, a);
  return check_launch("attention_tc");
}

}  // namespace

extern "C" int mmvid_attention_v3(const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, void* out,
                                  int out_bf16, long long ldo, int B, int H, int S, int S_pad, int mask_kind,
                                  const int* host_prev_rows, int n_prev, int tf32, cudaStream_t st);

extern "C" int mmvid_attention_v5(const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, void* out,
                                  int out_bf16, long long ldo, int B, int H, int S, int S_pad, int mask_kind,
                                  const int* host_prev_rows, int n_prev, int tf32, int poly8, int spin, int dual,
                                  int pingpong, unsigned long long* trace, cudaStream_t st);
extern "C" int mmvid_attention_v6(const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, void* out,
                                  int out_bf16, long long ldo, int B, int H, int S, int S_pad, int mask_kind,
                                  const int* host_prev_rows, int n_prev, int tf32, int poly8, unsigned long long* trace,
                                  cudaStream_t st);
namespace mmvid { extern unsigned long long* g_att_trace; }

namespace {
constexpr int ATT_IMPL_DEFAULT = 3;
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}
}  // namespace

extern "C" int mmvid_attention(const void* q, const void* k, const void* vt, void* out, int out_dtype, long long ldo,
                               int B, int H, int S, int S_pad, int mask_kind, const int* host_prev_rows, int n_prev,
                               int precision, mmvid_stream_t stream) {
  MMVID_REQUIRE(precision == MMVID_TF32 || precision == MMVID_BF16, "tensor-core attention needs TF32 or BF16");
  MMVID_REQUIRE(S_pad % 128 == 0 && S_pad >= S && S > 0, "S_pad multiple of 128");
  MMVID_REQUIRE(n_prev >= 0 && n_prev <= 4, "at most 4 mask_prev rows");
  MMVID_REQUIRE((long long)B * H <= 65535, "B*H <= 65535");
  const bool tf32 = precision == MMVID_TF32;
  const int esz = tf32 ? 4 : 2, dt = tf32 ? MMVID_DT_F32 : MMVID_DT_BF16;
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
  {
    // MMVID_ATT_IMPL: 4 = persistent rotating-score-buffer kernel (tc_attention4.cu), 3 (default) = rotating-score-buffer
    // kernel (tc_attention3.cu), 2 = two-tile ping-pong kernel
    // (tc_attention2.cu), 1 = the one-tile kernel below.  Measured defaults (profiles/r1_f_attention_v5.md): tf32 keeps
    // every exponential on MUFU, bf16 moves 2 of 8 to the FMA pipe.  MMVID_ATT_POLY (0|2|4 of every 8 exponentials on the
    // FMA pipe), MMVID_ATT_PP (0|1 MUFU ping-pong token) MMVID_ATT_SPIN (0|1) and MMVID_ATT_DUAL (0|1: one or two MMA-issuing threads) tune kernel 3.
    int impl = env_int("MMVID_ATT_IMPL", ATT_IMPL_DEFAULT);
    const int poly = env_int("MMVID_ATT_POLY", tf32 ? 0 : 2);
    if (impl == 4 && S > 128 && (ldo * (out_dtype == MMVID_DT_BF16 ? 2 : 4)) % 16 == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (poly == 0 || poly == 2))
      return mmvid_attention_v6(&tq, &tk, &tv, out, out_dtype == MMVID_DT_BF16, ldo, B, H, S, S_pad, mask_kind, host_prev_rows,
                                n_prev, tf32 ? 1 : 0, poly, mmvid::g_att_trace, to_stream(stream));
    if (impl == 4) impl = 3;  // one key tile per item, odd output alignment or poly 4: the non-persistent kernel
    if (impl == 3)
      return mmvid_attention_v5(&tq, &tk, &tv, out, out_dtype == MMVID_DT_BF16, ldo, B, H, S, S_pad, mask_kind,
                                host_prev_rows, n_prev, tf32 ? 1 : 0, env_int("MMVID_ATT_POLY", tf32 ? 0 : 2),
                                env_int("MMVID_ATT_SPIN", 0), env_int("MMVID_ATT_DUAL", 1), env_int("MMVID_ATT_PP", 0),
                                mmvid::g_att_trace, to_stream(stream));
    if (impl != 1)
      return mmvid_attention_v3(&tq, &tk, &tv, out, out_dtype == MMVID_DT_BF16, ldo, B, H, S, S_pad, mask_kind,
                                host_prev_rows, n_prev, tf32 ? 1 : 0, to_stream(stream));
  }
  AttArgs a{};
  a.out = out; a.ldo = ldo; a.out_bf16 = out_dtype == MMVID_DT_BF16;
  a.B = B; a.H = H; a.S = S; a.S_pad = S_pad; a.mask_kind = mask_kind; a.n_prev = n_prev;
  for (int i = 0; i < n_prev; ++i) a.prev_rows[i] = host_prev_rows[i];
  return tf32 ? launch_att<true>(tq, tk, tv, a, to_stream(stream)) : launch_att<false>(tq, tk, tv, a, to_stream(stream));
}
