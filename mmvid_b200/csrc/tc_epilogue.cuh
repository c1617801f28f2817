// 16-bit result tiles through TMA bulk stores (shared by tc_gemm.cu and tc_gemm2.cu).
//
// In the TMA-store epilogues a lane owns one output row of its warp's 32 x 32 accumulator chunks.  With 16-bit results
// two neighbouring chunks (64 columns) make one 128-byte row: both are packed in registers, written once to the warp's
// 4 KB staging buffer in the SWIZZLE_128B layout and leave as ONE {64 x 32} bulk store; a lone chunk (odd chunk count
// of the 192-wide tile, 64-wide tiles, the last columns of N) leaves as a {32 x 32} store with 64-byte rows.
#pragma once
#include "tc_common.cuh"

namespace mmvid {
namespace tc {

__device__ __forceinline__ void sts128_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// pack one 32-float chunk of this lane's row into 16 words (bf16 or fp16 pairs)
__device__ __forceinline__ void pack_chunk_h16(const float (&o)[32], uint32_t (&pk)[16], int f16) {
  if (f16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_h16<true>(o[2 * i], o[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_h16<false>(o[2 * i], o[2 * i + 1]);
  }
}

// Stage chunk `slot` (0 | 1) of a 2-chunk group: 128-byte rows, 16-byte unit u of row r at unit u ^ (r & 7).
__device__ __forceinline__ void stage_h16_pair(uint32_t st_base, int lane, int slot, const uint32_t (&pk)[16]) {
  const uint32_t row = st_base + (uint32_t)(lane * 128);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128_u32(row + (uint32_t)((((slot * 4 + j) ^ (lane & 7))) * 16), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}
// Stage a lone chunk: 64-byte rows in a SWIZZLE_128B box, i.e. byte offset x of the dense box lives at
// x ^ (((x >> 7) & 7) << 4).
__device__ __forceinline__ void stage_h16_single(uint32_t st_base, int lane, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t lin = (uint32_t)(lane * 64 + j * 16);
    sts128_u32(st_base + (lin ^ (((lin >> 7) & 7u) << 4)), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tc
}  // namespace mmvid
