// Non-overlapping patches of the ViLT patch embedding as GEMM rows: HF ``ViltPatchEmbeddings`` is
// Conv2d(3, 768, kernel 32, stride 32) (the backbone of reference src/modeling/vilt.py:19,127), i.e. ONE
// [B h w, C ps ps] x [C ps ps, 768] product once the image is cut into patches.  This kernel does the cut and the cast
// in one pass: pixel_values [B, C, H, W] (fp32 or bf16) -> patches [B h w, C ps ps] bf16, row = (b, patch row, patch
// column), column = (channel, y, x) -- the order of the convolution weight's [768, C, ps, ps] view.  torch's generic
// strided copy took 46 us for the 14 M elements of a 32 x 3 x 384 x 384 batch (plus a separate cast); this is one
// coalesced read of 128-byte patch rows and one 16-byte store per 8 pixels.
#include <cuda_bf16.h>

#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

template <typename TIn>
__global__ void __launch_bounds__(256)
patchify_kernel(const TIn* __restrict__ px, uint4* __restrict__ out, int C, int H, int W, int ps, int h, int w, int64_t n_vec) {
  const int cols8 = C * ps * ps / 8, ps8 = ps / 8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_vec; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int c8 = static_cast<int>(i % cols8);
    const int64_t row = i / cols8;
    const int x8 = c8 % ps8, y = (c8 / ps8) % ps, c = c8 / (ps8 * ps);
    const int pw = static_cast<int>(row % w), ph = static_cast<int>((row / w) % h);
    const int64_t b = row / (static_cast<int64_t>(w) * h);
    const TIn* src = px + ((b * C + c) * H + ph * ps + y) * static_cast<int64_t>(W) + pw * ps + x8 * 8;
    uint4 o;
    if constexpr (sizeof(TIn) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), d = __ldg(reinterpret_cast<const float4*>(src) + 1);
      __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x, a.y), t1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 t2 = __floats2bfloat162_rn(d.x, d.y), t3 = __floats2bfloat162_rn(d.z, d.w);
      o = make_uint4(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1), *reinterpret_cast<uint32_t*>(&t2),
                     *reinterpret_cast<uint32_t*>(&t3));
    } else {
      o = __ldg(reinterpret_cast<const uint4*>(src));
    }
    out[i] = o;
  }
}

}  // namespace
}  // namespace fd

extern "C" int feddat_patchify(const void* pixel_values, void* patches, int B, int C, int H, int W, int ps, int in_dtype,
                               void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(in_dtype == FEDDAT_DTYPE_F32 || in_dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED,
             "patchify: pixel_values must be fp32 or bf16 (dtype=%d)", in_dtype);
  FD_REQUIRE(pixel_values && patches, FD_ERR_INVALID, "patchify: null pointer argument");
  FD_REQUIRE(B >= 0 && C >= 1 && ps >= 8 && ps % 8 == 0 && H >= ps && W >= ps && W % 8 == 0, FD_ERR_INVALID,
             "patchify: bad geometry B=%d C=%d H=%d W=%d patch=%d", B, C, H, W, ps);
  FD_REQUIRE(((reinterpret_cast<uintptr_t>(pixel_values) | reinterpret_cast<uintptr_t>(patches)) & 31) == 0, FD_ERR_INVALID,
             "patchify: tensors must be 32-byte aligned");
  if (B == 0) return FD_OK;
  const int h = H / ps, w = W / ps;
  const int64_t n_vec = static_cast<int64_t>(B) * h * w * C * ps * ps / 8;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  int64_t blocks = (n_vec + 255) / 256;
  if (blocks > 16ll * sms) blocks = 16ll * sms;
  auto st = static_cast<cudaStream_t>(stream);
  if (in_dtype == FEDDAT_DTYPE_F32)
    patchify_kernel<float><<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const float*>(pixel_values),
                                                                      static_cast<uint4*>(patches), C, H, W, ps, h, w, n_vec);
  else
    patchify_kernel<__nv_bfloat16><<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(pixel_values),
                                                                              static_cast<uint4*>(patches), C, H, W, ps, h, w, n_vec);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
