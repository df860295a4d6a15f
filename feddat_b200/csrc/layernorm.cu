// Fused (residual add +) LayerNorm over the model dimension d = 768 for the frozen ViLT blocks that
// sit on either side of every DAT site (HF ViltLayer.forward: layernorm_before, "first residual
// connection" + layernorm_after; reference call sites src/modeling/vilt.py:127 and
// src/modeling/adaptered_output.py:73-79 receive exactly these tensors).  SURVEY.md section 8(f) n3
// ("backbone efficiency"): torch's row-per-block LayerNorm and separate add kernels took 17.7 + 6 us
// per [5920, 768] bf16 tensor, 2.9 ms of a 9.4 ms train step, for 18 MB of traffic each.
//
//   forward :  s = bf16(x + res)   (res optional; s written only when res is given)
//              s2 = bf16(s + bias2) (optional): the residual stream with the NEXT dense layer's bias already
//              added, so that layer's "dense + bias + residual" is one GEMM with a beta = 1 epilogue
//              y = bf16((s - mean) * rstd * w + b),   mean / rstd in fp32 per row (saved for backward)
//   backward:  dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)) + dsum,   g = dy * w,
//              xhat = (s - mean) * rstd;  w, b frozen (no affine gradients); dsum = the gradient
//              arriving at `s` from the second consumer of the residual stream (optional)
//
// HBM-bound byte work: one warp per row, 24 bf16 per lane as three 16-byte vectors (coalesced 512 B
// per warp and load), statistics by warp shuffles in fp32, everything else in registers; 8 rows per
// 256-thread block.  Same arithmetic order as torch (fp32 mean, fp32 biased variance, rsqrt).
#include <cuda_bf16.h>

#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

constexpr int kD = 768;
constexpr int VPL = kD / (32 * 8);   // 16-byte vectors per lane = 3

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
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

template <bool kHasRes>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, const uint4* __restrict__ w,
              const uint4* __restrict__ b, uint4* __restrict__ y, uint4* __restrict__ sum_out,
              const uint4* __restrict__ bias2, uint4* __restrict__ sum2_out,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t base = static_cast<size_t>(row) * (kD / 8);
  float v[VPL][8];
  float s1 = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = lane + 32 * k;
    unpack8(x[base + idx], v[k]);
    if constexpr (kHasRes) {
      float r[8];
      unpack8(res[base + idx], r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k][i] += r[i];
      const uint4 sv = pack8(v[k]);          // the residual stream is a bf16 tensor: LN sees the rounded sum
      sum_out[base + idx] = sv;
      unpack8(sv, v[k]);
    }
    if (sum2_out != nullptr) {
      float b2[8], o2[8];
      unpack8(bias2[idx], b2);
#pragma unroll
      for (int i = 0; i < 8; ++i) o2[i] = v[k][i] + b2[i];
      sum2_out[base + idx] = pack8(o2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s1 += v[k][i];
  }
  const float mean = warp_sum(s1) * (1.f / kD);
  float s2 = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[k][i] - mean;
      s2 += d * d;
    }
  const float rstd = rsqrtf(warp_sum(s2) * (1.f / kD) + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = lane + 32 * k;
    float wf[8], bf[8], o[8];
    unpack8(w[idx], wf);
    unpack8(b[idx], bf);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (v[k][i] - mean) * rstd * wf[i] + bf[i];
    y[base + idx] = pack8(o);
  }
}

template <bool kHasDsum>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ dsum, const uint4* __restrict__ s,
              const uint4* __restrict__ w, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, uint4* __restrict__ dx, int M) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const size_t base = static_cast<size_t>(row) * (kD / 8);
  const float mean = mean_in[row], rstd = rstd_in[row];
  float g[VPL][8], xh[VPL][8];
  float a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = lane + 32 * k;
    float wf[8];
    unpack8(dy[base + idx], g[k]);
    unpack8(s[base + idx], xh[k]);
    unpack8(w[idx], wf);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[k][i] *= wf[i];
      xh[k][i] = (xh[k][i] - mean) * rstd;
      a1 += g[k][i];
      a2 += g[k][i] * xh[k][i];
    }
  }
  a1 = warp_sum(a1) * (1.f / kD);
  a2 = warp_sum(a2) * (1.f / kD);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = lane + 32 * k;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rstd * (g[k][i] - a1 - xh[k][i] * a2);
    if constexpr (kHasDsum) {
      float e[8];
      unpack8(dsum[base + idx], e);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += e[i];
    }
    dx[base + idx] = pack8(o);
  }
}

int check_ln(const char* who, int64_t M, int d, int dtype) {
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED, "%s: only bf16 is implemented (dtype=%d)", who, dtype);
  FD_REQUIRE(d == kD, FD_ERR_UNSUPPORTED, "%s: model_dim must be 768 (got %d)", who, d);
  FD_REQUIRE(M >= 0 && M < (1ll << 31) - 8, FD_ERR_INVALID, "%s: bad row count %lld", who, (long long)M);
  return FD_OK;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace fd

extern "C" int feddat_ln_fwd(const void* x, const void* res, const void* weight, const void* bias, void* y,
                             void* sum_out, const void* bias2, void* sum2_out, float* mean, float* rstd,
                             int64_t M, int d, float eps, int dtype, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  if ((rc = check_ln("ln_fwd", M, d, dtype))) return rc;
  FD_REQUIRE(x && weight && bias && y && mean && rstd, FD_ERR_INVALID, "ln_fwd: null pointer argument");
  FD_REQUIRE((res == nullptr) == (sum_out == nullptr), FD_ERR_INVALID,
             "ln_fwd: res and sum_out must both be given or both be NULL");
  FD_REQUIRE((bias2 == nullptr) == (sum2_out == nullptr) && aligned16(bias2) && aligned16(sum2_out), FD_ERR_INVALID,
             "ln_fwd: bias2 and sum2_out must both be given (16-byte aligned) or both be NULL");
  FD_REQUIRE(aligned16(x) && aligned16(weight) && aligned16(bias) && aligned16(y) && aligned16(res) &&
                 aligned16(sum_out),
             FD_ERR_INVALID, "ln_fwd: tensors must be 16-byte aligned");
  if (M == 0) return FD_OK;
  const int blocks = static_cast<int>((M + 7) / 8);
  auto st = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const uint4*>(x);
  auto W = static_cast<const uint4*>(weight);
  auto B = static_cast<const uint4*>(bias);
  if (res)
    ln_fwd_kernel<true><<<blocks, 256, 0, st>>>(X, static_cast<const uint4*>(res), W, B, static_cast<uint4*>(y),
                                                static_cast<uint4*>(sum_out), static_cast<const uint4*>(bias2),
                                                static_cast<uint4*>(sum2_out), mean, rstd, static_cast<int>(M), eps);
  else
    ln_fwd_kernel<false><<<blocks, 256, 0, st>>>(X, nullptr, W, B, static_cast<uint4*>(y), nullptr,
                                                 static_cast<const uint4*>(bias2), static_cast<uint4*>(sum2_out),
                                                 mean, rstd, static_cast<int>(M), eps);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

extern "C" int feddat_ln_bwd(const void* dy, const void* dsum, const void* s, const void* weight,
                             const float* mean, const float* rstd, void* dx, int64_t M, int d, int dtype,
                             void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  if ((rc = check_ln("ln_bwd", M, d, dtype))) return rc;
  FD_REQUIRE(dy && s && weight && mean && rstd && dx, FD_ERR_INVALID, "ln_bwd: null pointer argument");
  FD_REQUIRE(aligned16(dy) && aligned16(dsum) && aligned16(s) && aligned16(weight) && aligned16(dx), FD_ERR_INVALID,
             "ln_bwd: tensors must be 16-byte aligned");
  if (M == 0) return FD_OK;
  const int blocks = static_cast<int>((M + 7) / 8);
  auto st = static_cast<cudaStream_t>(stream);
  if (dsum)
    ln_bwd_kernel<true><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(dy), static_cast<const uint4*>(dsum),
                                                static_cast<const uint4*>(s), static_cast<const uint4*>(weight), mean,
                                                rstd, static_cast<uint4*>(dx), static_cast<int>(M));
  else
    ln_bwd_kernel<false><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(dy), nullptr, static_cast<const uint4*>(s),
                                                 static_cast<const uint4*>(weight), mean, rstd,
                                                 static_cast<uint4*>(dx), static_cast<int>(M));
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
