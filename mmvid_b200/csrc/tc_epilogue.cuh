// 16-bit result tiles through TMA bulk stores (shared by tc_gemm.cu and tc_gemm2.cu).
//
// In the TMA-store epilogues a lane owns one output row of its warp's 32 x 32 accumulator chunks.  With 16-bit results
// two neighbouring chunks (64 columns) make one 128-byte row: both are packed in registers, written once to the warp's
// 4 KB staging buffer in the SWIZZLE_128B layout and leave as ONE {64 x 32} bulk store; a lone chunk (odd chunk count
// of the 192-wide tile, 64-wide tiles, the last columns of N) leaves as a {32 x 32} store with dense 64-byte rows
// through a SWIZZLE_NONE tensor map (a few bank conflicts on a rare path).
#pragma once
#include "tc_common.cuh"

namespace mmvid {
namespace tc {

__device__ __forceinline__ void sts128_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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
