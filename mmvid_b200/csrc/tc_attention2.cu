// Flash attention v3 for sm_100a: TWO 128-query tiles per CTA ping-pong through the tensor core
// (FA4-style schedule) so that the MMA pipe and the softmax (MUFU/FMA) pipes overlap inside one SM:
//
//   tensor core :  QK_A0 . QK_B0 PV_A0 QK_A1 | PV_B0 QK_B1 | PV_A1 QK_A2 | PV_B1 QK_B2 | ...   (A and B in anti-phase)
//   softmax grp A:        [ P_A0 ......... ][O_A0][ P_A1 ......... ][O_A1] ...
//   softmax grp B:              [ P_B0 ......... ][O_B0][ P_B1 ......... ][O_B1] ...
//
// One CTA per SM, 384 threads = 3 warpgroups: warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps} shrinks to 56
// registers/thread with setmaxnreg.dec, the two softmax warpgroups (A: warps 4-7, B: warps 8-11) grow to 224 so a
// whole 128-column score row lives in registers (the register file is partitioned per SM sub-partition: 3 warps x 32
// lanes x (56 + 224 + 224) = 16128 <= 16384).  K / V^T tiles are double-buffered and shared by both query tiles
// (half the L2 operand traffic of the one-tile kernel).  TMEM: S_A [0,128) S_B [128,256) O_A [256,320)
// O_B [320,384).  Softmax, masking and the in-place P write-back are identical to tc_attention.cu.
#include "common.cuh"
#include "tc_common.cuh"

using namespace mmvid;
using namespace mmvid::tc;

namespace {

constexpr int ATT2_THREADS = 384;
constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TMEM_COLS = 512;
__host__ __device__ constexpr int S_COL_OF(int g) { return g * 128; }
__host__ __device__ constexpr int O_COL_OF(int g) { return 256 + g * 64; }

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
__device__ __forceinline__ void ffma2(float x0, float x1, float s, float b, float& y0, float& y1) {
  unsigned long long px, ps, pb, py;
  asm("mov.b64 %0, {%1, %2};" : "=l"(px) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(py) : "l"(px), "l"(ps), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(py));
}

struct Att2Args {
  unsigned long long* trace;  // debug timeline (mmvid_debug_attention_trace), normally null
  void* out; long long ldo; int out_bf16;
  int B, H, S, S_pad, mask_kind;
  int prev_rows[4]; int n_prev;
};

// CTA (0,0) stamps clock64() at pipeline events when a trace buffer is installed (scripts/att_trace.py)
__device__ __forceinline__ void att2_stamp(const Att2Args& a, int idx) {
  if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) a.trace[idx] = clock64();
}

template <bool TF32>
constexpr size_t att2_smem_bytes() {
  // Q_A, Q_B, 2 x K, 2 x V^T tiles (each BQ*HD elements) + barriers + alignment slack
  return (size_t)6 * (TF32 ? 32768 : 16384) + 1024 + 256;
}

template <bool TF32>
__global__ void __launch_bounds__(ATT2_THREADS, 1) attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV, Att2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]  (count 2: QK_A and QK_B both read the stage)
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]  (count 2)
  uint64_t* s_full = bars + 9;    // [2] per group
  uint64_t* p_ready = bars + 11;  // [2]
  uint64_t* o_full = bars + 13;   // [2]
  uint64_t* o_free = bars + 15;   // [2]
  uint64_t* all_done = bars + 17;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~(uintptr_t)1023);
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int BKE = 128 / ESZ;
  constexpr int QK_KB = HD / BKE;   // 2 | 1
  constexpr int PV_KB = BKV / BKE;  // 4 | 2
  constexpr int T_BYTES = BQ * HD * ESZ;  // every tile (Q, K, V^T) has the same byte size
  uint8_t* sQ[2] = {tiles, tiles + T_BYTES};
  uint8_t* sK[2] = {tiles + 2 * T_BYTES, tiles + 3 * T_BYTES};
  uint8_t* sV[2] = {tiles + 4 * T_BYTES, tiles + 5 * T_BYTES};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int bh = blockIdx.y;
  const int b = bh / a.H, h = bh - b * a.H;
  int n_kv = (a.S + BKV - 1) / BKV;
  if (a.mask_kind == MMVID_MASK_CAUSAL) n_kv = min(n_kv, (q0 + 2 * BQ - 1) / BKV + 1);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 128);
    }
    mbar_init(all_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * T_BYTES);
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
          tma_load_2d(sQ[g] + kb * (BQ * 128), &tmQ, q_full, kb * BKE, bh * a.S_pad + q0 + g * BQ);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], T_BYTES);
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb)
          tma_load_2d(sK[st] + kb * (BKV * 128), &tmK, &k_full[st], kb * BKE, bh * a.S_pad + j * BKV);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], T_BYTES);
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb)
          tma_load_2d(sV[st] + kb * (HD * 128), &tmV, &v_full[st], j * BKV + kb * BKE, bh * HD);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc<TF32>(BQ, BKV);
      constexpr uint32_t idesc_pv = make_idesc<TF32>(BQ, HD);
      auto issue_qk = [&](int g, int st) {
#pragma unroll
        for (int kb = 0; kb < QK_KB; ++kb) {
          const uint64_t qd = make_smem_desc_sw128(smem_u32(sQ[g] + kb * (BQ * 128)));
          const uint64_t kd = make_smem_desc_sw128(smem_u32(sK[st] + kb * (BKV * 128)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ss<TF32>(tmem_base + S_COL_OF(g), desc_advance(qd, kk * 32), desc_advance(kd, kk * 32), idesc_qk,
                         (kb | kk) != 0);
        }
        tc_commit(&k_empty[st]);
        tc_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int st, bool first_tile) {
#pragma unroll
        for (int kb = 0; kb < PV_KB; ++kb) {
          const uint64_t vd = make_smem_desc_sw128(smem_u32(sV[st] + kb * (HD * 128)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // O accumulates in TMEM over ALL kv tiles (lazy rescale, see softmax warps)
            mma_ts<TF32>(tmem_base + O_COL_OF(g), tmem_base + S_COL_OF(g) + kb * 32 + kk * 8, desc_advance(vd, kk * 32), idesc_pv,
                         (!first_tile || (kb | kk) != 0) ? 1u : 0u);
        }
        tc_commit(&v_empty[st]);
        tc_commit(&o_full[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      // Anti-phase start: only tile A's first QK^T is issued up front; tile B's first QK^T goes out once A has
      // finished its first softmax.  From then on A's exp2 phase overlaps B's MMAs and vice versa (if both tiles run
      // in lockstep they fight for the MUFU pipe and then queue behind each other on the tensor pipe).
      issue_qk(0, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t kvph = (j >> 1) & 1;   // phase of the K/V stage barriers
        const uint32_t ph = j & 1;            // phase of the per-tile group barriers (one completion per kv tile)
        const bool more = (j + 1 < n_kv);
        mbar_wait(&v_full[st], kvph);
        if (more) mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_ready[g], ph);                  // group g wrote P_j over S_g (and rescaled O_g if it had to)
          tc_fence_after();
          if (j < 32) att2_stamp(a, j * 4 + g * 2);
          if (j == 0 && g == 0) issue_qk(1, 0);        // delayed start of tile B (see above)
          issue_pv(g, st, j == 0);
          if (more) issue_qk(g, st ^ 1);               // S_g is free: PV_j (same issue stream) consumed P_j first
          if (j < 32) att2_stamp(a, j * 4 + g * 2 + 1);
        }
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
    const uint32_t t_s = t_row + S_COL_OF(g), t_o = t_row + O_COL_OF(g);
    int lo = 0, hi = a.S;
    if (a.mask_kind == MMVID_MASK_CAUSAL) hi = min(a.S, row + 1);
    else if (a.mask_kind == MMVID_MASK_PREV) {
      for (int i = 0; i < a.n_prev; ++i) if (a.prev_rows[i] == row) lo = row;
    }
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e): scores are handled in the log2 domain
    // Lazy rescale (FA4): O_g accumulates in TMEM across kv tiles relative to a reference max m_ref that is only moved
    // when the running row max exceeds it by more than 2^8 (P <= 256 is harmless in fp32/tf32/bf16).  The per-tile
    // critical path is then  S -> P  only: no O round trip through registers, and the whole 128-column score row
    // lives in registers between ONE tcgen05.wait::ld and ONE tcgen05.wait::st.
    constexpr float RESCALE_THRESH = 8.f;
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int kv0 = j * BKV;
      const bool tile_full = __all_sync(0xffffffffu, (kv0 >= lo) && (kv0 + BKV <= hi));
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      const bool tr = (qd == 0 && lane == 0 && j < 32);
      const int tb = 128 + g * 192 + j * 6;
      if (tr) att2_stamp(a, tb + 0);
      uint32_t r[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(t_s + ch * 32, r[ch]);
      tmem_ld_wait();
      if (tr) att2_stamp(a, tb + 1);
      if (!tile_full) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = kv0 + ch * 32 + i;
            if (!(col >= lo && col < hi)) r[ch][i] = 0xff800000u;  // -inf
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
      if (tr) att2_stamp(a, tb + 2);
      // move the reference only when needed (warp-uniform decision because TMEM ld/st are warp collectives)
      const bool need = (mx != -INFINITY) && (m_ref == -INFINITY || (mx - m_ref) * c > RESCALE_THRESH);
      float alpha = 1.f;
      bool resc = false;
      if (need) {
        if (m_ref != -INFINITY) { alpha = ex2_approx((m_ref - mx) * c); resc = true; }  // else: O row and l are still 0
        m_ref = mx;
      }
      if (__any_sync(0xffffffffu, resc)) {
        // rare path: O_g *= alpha in TMEM (alpha = 1 for rows that keep their reference).  PV_{j-1} must have retired.
        mbar_wait(&o_full[g], ph ^ 1);
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
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1;
          ffma2(__uint_as_float(r[ch][i]), __uint_as_float(r[ch][i + 1]), c, nmc, a0, a1);
          const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);  // exp2(-inf) = 0 for masked keys
          if constexpr (TF32) {
            rs0 += e0; rs1 += e1;
            r[ch][i] = __float_as_uint(e0); r[ch][i + 1] = __float_as_uint(e1);
          } else {
            __nv_bfloat162 v2 = __floats2bfloat162_rn(e0, e1);
            const uint32_t w = *reinterpret_cast<uint32_t*>(&v2);
            pk[i >> 1] = w;
            rs0 += __uint_as_float(w << 16);
            rs1 += __uint_as_float(w & 0xffff0000u);
          }
        }
        if constexpr (TF32) tmem_st32(t_s + ch * 32, r[ch]);
        else tmem_st16(t_s + ch * 16, pk);
      }
      if (tr) att2_stamp(a, tb + 3);
      tmem_st_wait();
      if (tr) att2_stamp(a, tb + 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);  // one arrival per warp (4 per group), not 128 serialised ones
      if (tr) att2_stamp(a, tb + 5);
      l += rs0 + rs1;
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
    uint8_t* stage_base = g == 0 ? sK[0] : sV[0];  // 2 contiguous tiles each: >= 128 x 68 floats
    const int q_tile0 = q0 + g * BQ;
    if (a.out_bf16) {
      constexpr int LD = HD + 8;
      __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(stage_base) + (size_t)row_local * LD;
#pragma unroll
      for (int i = 0; i < HD; i += 2)
        *reinterpret_cast<__nv_bfloat162*>(st + i) = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv);
      __syncwarp();
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out);
      for (int r0 = 0; r0 < 32; r0 += 4) {
        const int rl = qd * 32 + r0 + (lane >> 3);
        const int s = q_tile0 + rl;
        if (s < a.S) {
          const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<__nv_bfloat16*>(stage_base) + (size_t)rl * LD + (lane & 7) * 8);
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

template <bool TF32>
int launch_att2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Att2Args& a, cudaStream_t st) {
  static bool attr_set = false;
  constexpr size_t smem = att2_smem_bytes<TF32>();
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(attention_tc2_kernel<TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail(MMVID_ECUDA, "cudaFuncSetAttribute(attention_tc2): %s", cudaGetErrorString(err));
    attr_set = true;
  }
  dim3 grid((a.S_pad / BQ + 1) / 2, a.B * a.H);
  attention_tc2_kernel<TF32><<<grid, ATT2_THREADS, smem, st>>>(tq, tk, tv, a);
  return check_launch("attention_tc2");
}

}  // namespace

// called by mmvid_attention (tc_attention.cu) unless MMVID_ATT_IMPL=1 selects the one-tile kernel
namespace mmvid { unsigned long long* g_att_trace = nullptr; }
// Debug / profiling hook: CTA (0,0) of every following attention launch writes clock64() stamps of its pipeline events
// into `dev_buf` (>= 512 entries; pass NULL to switch it off).  MMA warp: [j*4 + g*2] p_ready_g observed,
// [+1] PV_g(j) / QK_g(j+1) issued; softmax warp of tile g: [128 + g*192 + j*6 + {0: S ready, 1: S in registers,
// 2: row max, 3: exps done / P stores issued, 4: P stores landed, 5: p_ready signalled}].
extern "C" int mmvid_debug_attention_trace(unsigned long long* dev_buf) {
  mmvid::g_att_trace = dev_buf;
  return MMVID_OK;
}

extern "C" int mmvid_attention_v3(const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, void* out,
                                  int out_bf16, long long ldo, int B, int H, int S, int S_pad, int mask_kind,
                                  const int* host_prev_rows, int n_prev, int tf32, cudaStream_t st) {
  Att2Args a{};
  a.trace = mmvid::g_att_trace;
  a.out = out; a.ldo = ldo; a.out_bf16 = out_bf16;
  a.B = B; a.H = H; a.S = S; a.S_pad = S_pad; a.mask_kind = mask_kind; a.n_prev = n_prev;
  for (int i = 0; i < n_prev; ++i) a.prev_rows[i] = host_prev_rows[i];
  return tf32 ? launch_att2<true>(*tq, *tk, *tv, a, st) : launch_att2<false>(*tq, *tk, *tv, a, st);
}
