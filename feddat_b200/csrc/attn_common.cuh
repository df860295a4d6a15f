// Shared by the short-sequence attention kernels (attn.cu forward, attn_bwd.cu backward).
#pragma once
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {
namespace {

constexpr int AD = 64;                 // head dimension
constexpr int AQ = 128;                // rows per tile (TMEM lanes)

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
[[maybe_unused]] __device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// [B, S, cols] bf16 with token stride `ld` elements, viewed by (column, token, batch); box = [rows x 64 columns]
[[maybe_unused]] int make_tmap_tokens(CUtensorMap* out, const void* gptr, int B, int S, int cols, int64_t ld, uint32_t box_rows) {
  using PFN = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static PFN enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FD_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    FD_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, FD_ERR_CUDA,
               "cuTensorMapEncodeTiled entry point not available");
    enc = reinterpret_cast<PFN>(fn);
  }
  FD_REQUIRE((reinterpret_cast<uintptr_t>(gptr) & 15) == 0 && (ld * 2) % 16 == 0 && cols % AD == 0 && ld >= cols,
             FD_ERR_INVALID, "attention operand %p: base / token stride %lld not 16-byte aligned or narrower than %d",
             gptr, (long long)ld, cols);
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * S};
  cuuint32_t box[3] = {AD, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, FD_ERR_CUDA, "cuTensorMapEncodeTiled (token view) failed (%d) B=%d S=%d ld=%lld", (int)r,
             B, S, (long long)ld);
  return FD_OK;
}

}  // namespace
}  // namespace fd
