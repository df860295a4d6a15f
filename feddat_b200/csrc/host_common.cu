#include "host_common.h"

#include <string.h>

#include <mutex>

namespace fd {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  FD_REQUIRE(enc != nullptr, FD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  FD_REQUIRE((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, FD_ERR_INVALID,
             "tensor base %p is not 16-byte aligned", gptr);
  FD_REQUIRE((row_stride_elems * 2) % 16 == 0, FD_ERR_INVALID,
             "row stride %llu elements is not a multiple of 16 bytes",
             (unsigned long long)row_stride_elems);
  FD_REQUIRE(box_cols * 2 == 128 && box_rows >= 1 && box_rows <= 256, FD_ERR_INVALID,
             "bad TMA box %u x %u", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, FD_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu stride=%llu box=%ux%u", (int)r,
             (unsigned long long)rows, (unsigned long long)cols,
             (unsigned long long)row_stride_elems, box_rows, box_cols);
  return FD_OK;
}

int make_tmap_bf16_kblocks(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols,
                           uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_blocks) {
  PFN_encodeTiled enc = get_encode_fn();
  FD_REQUIRE(enc != nullptr, FD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  FD_REQUIRE((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, FD_ERR_INVALID,
             "tensor base %p is not 16-byte aligned", gptr);
  FD_REQUIRE(cols % 64 == 0 && cols / 64 >= 1 && cols / 64 <= 256 && (row_stride_elems * 2) % 16 == 0 &&
                 box_rows >= 1 && box_rows <= 256 && box_blocks >= 1 && box_blocks <= cols / 64,
             FD_ERR_INVALID, "bad k-block tensor map: cols=%llu stride=%llu box=%u rows x %u blocks",
             (unsigned long long)cols, (unsigned long long)row_stride_elems, box_rows, box_blocks);
  cuuint64_t gdim[3] = {64, rows, cols / 64};
  cuuint64_t gstride[2] = {row_stride_elems * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_blocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FD_REQUIRE(r == CUDA_SUCCESS, FD_ERR_CUDA,
             "cuTensorMapEncodeTiled (3-D k-block view) failed (%d) rows=%llu cols=%llu stride=%llu",
             (int)r, (unsigned long long)rows, (unsigned long long)cols,
             (unsigned long long)row_stride_elems);
  return FD_OK;
}

#ifdef FEDDAT_DEBUG
unsigned long long* g_trace = nullptr;
#endif

int device_sm_count(int* out) {
  static int cached[64] = {0};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && cached[dev] > 0) {
    *out = cached[dev];
    return FD_OK;
  }
  int n = 0;
  FD_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 64) cached[dev] = n;
  *out = n;
  return FD_OK;
}

int check_device_sm100() {
  static int ok[64] = {0};
  // cuTensorMapEncodeTiled is a DRIVER call and needs a current context on the calling thread.  PyTorch's
  // autograd threads select their device lazily, so a backward entry point can be the first CUDA call of
  // its thread: bind the runtime's primary context once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    FD_CHECK_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && ok[dev]) return FD_OK;
  int major = 0;
  FD_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FD_REQUIRE(major == 10, FD_ERR_NO_DEVICE,
             "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
  if (dev < 64) ok[dev] = 1;
  return FD_OK;
}

}  // namespace fd
