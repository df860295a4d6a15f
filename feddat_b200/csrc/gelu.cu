// Exact (erf) GELU forward / backward over bf16 tensors: the activation of the frozen ViLT
// intermediate layer (HF ViltIntermediate: dense 768 -> 3072 + GELU), whose output feeds
// ViltOutput.dense and then every DAT site (reference src/modeling/adaptered_output.py:73-79).
// SURVEY.md section 8(f) n3.  [5920, 3072] per layer: torch's generic elementwise kernels took 22.5 us
// (fwd) and 31 us (bwd); this is pure streaming work -- 16-byte vectors, grid-stride over a grid
// sized to the SM count, fp32 math, one rounding.
//   y  = 0.5 x (1 + erf(x / sqrt 2))
//   dx = dy (0.5 (1 + erf(x / sqrt 2)) + x exp(-x^2 / 2) / sqrt(2 pi))
#include <cuda_bf16.h>

#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool kBwd>
__global__ void __launch_bounds__(256)
gelu_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ out, int64_t n_vec) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_vec; i += static_cast<int64_t>(gridDim.x) * 256) {
    float xf[8], o[8];
    unpack8(x[i], xf);
    if constexpr (kBwd) {
      float g[8];
      unpack8(dy[i], g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float cdf = 0.5f * (1.f + erff(xf[k] * 0.70710678118654752f));
        const float pdf = 0.3989422804014327f * __expf(-0.5f * xf[k] * xf[k]);
        o[k] = g[k] * (cdf + xf[k] * pdf);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.5f * xf[k] * (1.f + erff(xf[k] * 0.70710678118654752f));
    }
    out[i] = pack8(o);
  }
}

int launch_gelu(bool bwd, const void* x, const void* dy, void* out, int64_t n, int dtype, void* stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  const char* who = bwd ? "gelu_bwd" : "gelu_fwd";
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED, "%s: only bf16 is implemented (dtype=%d)", who, dtype);
  FD_REQUIRE(x && out && (!bwd || dy), FD_ERR_INVALID, "%s: null pointer argument", who);
  FD_REQUIRE(n >= 0 && n % 8 == 0, FD_ERR_INVALID, "%s: element count %lld must be a multiple of 8", who, (long long)n);
  FD_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0,
             FD_ERR_INVALID, "%s: tensors must be 16-byte aligned", who);
  if (n == 0) return FD_OK;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  const int64_t n_vec = n / 8;
  int64_t blocks = (n_vec + 255) / 256;
  if (blocks > 8ll * sms) blocks = 8ll * sms;       // 8 resident 256-thread blocks per SM, grid-stride beyond
  auto st = static_cast<cudaStream_t>(stream);
  if (bwd)
    gelu_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(dy),
                                                               static_cast<uint4*>(out), n_vec);
  else
    gelu_kernel<false><<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const uint4*>(x), nullptr,
                                                                static_cast<uint4*>(out), n_vec);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int feddat_gelu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream) {
  return fd::launch_gelu(false, x, nullptr, y, n, dtype, stream);
}
extern "C" int feddat_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, int dtype, void* stream) {
  return fd::launch_gelu(true, x, dy, dx, n, dtype, stream);
}
