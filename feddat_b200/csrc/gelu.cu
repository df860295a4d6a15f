// Exact (erf) GELU forward / backward over bf16 tensors: the activation of the frozen ViLT
// intermediate layer (HF ViltIntermediate: dense 768 -> 3072 + GELU), whose output feeds
// ViltOutput.dense and then every DAT site (reference src/modeling/adaptered_output.py:73-79).
// SURVEY.md section 8(f) n3.  [5920, 3072] per layer: torch's generic elementwise kernels took 22.5 us
// (fwd) and 31 us (bwd), and so did a straight port: both are bound by libdevice's erff (~30
// instructions).  Here Phi(x) comes from the complementary error function in the Abramowitz-Stegun
// 7.1.26 form, which shares ONE exponential with the density the backward needs:
//   z = |x| / sqrt 2,  t = 1 / (1 + 0.3275911 z),  e = exp(-z^2)
//   q = 0.5 erfc(z) = 0.5 t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) e        |abs error| < 1e-7
//   Phi(x) = x >= 0 ? 1 - q : q            (no cancellation in the negative tail)
//   y  = x Phi(x)                           dx = dy (Phi(x) + x e / sqrt(2 pi))
// i.e. one MUFU.RCP, one MUFU.EX2 and ~12 FMAs per element; the result differs from erff-based GELU by
// < 1e-6 absolute before the bf16 rounding (identical after it except at rounding boundaries).
// 16-byte vectors, grid-stride over a grid sized to the SM count, fp32 math, one rounding.
#include <cuda_bf16.h>

#include "feddat_b200.h"
#include "gelu_math.cuh"
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
  // four independent 16-byte loads per thread and trip: with one, a full SM holds only 32 KB in flight and
  // the kernel streamed at ~3 TB/s
  constexpr int U = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (int64_t i0 = blockIdx.x * 256ll + threadIdx.x; i0 < n_vec; i0 += U * stride) {
    uint4 xv[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n_vec) {
        xv[u] = x[i];
        if constexpr (kBwd) gv[u] = dy[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n_vec) break;
      float xf[8], o[8];
      unpack8(xv[u], xf);
      if constexpr (kBwd) {
        float g[8];
        unpack8(gv[u], g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float cdf, e;
          gelu_terms(xf[k], cdf, e);
          o[k] = g[k] * fmaf(xf[k], 0.3989422804014327f * e, cdf);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float cdf, e;
          gelu_terms(xf[k], cdf, e);
          o[k] = xf[k] * cdf;
        }
      }
      out[i] = pack8(o);
    }
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
  blocks = (blocks + 3) / 4;                        // four vectors per thread and trip
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
