// Packs the fp32 master weights of the active adapter branches (nn.Linear layout, reference
// adapter.py:35,41) into the bf16 operand layouts the tcgen05 kernels consume:
//   Wd_cat  [nR, d]  rows of down.weight stacked branch after branch      (fwd GEMM1, K-major B)
//   WdT_cat [d, nR]  its transpose                                        (bwd dX GEMM, K-major B)
//   Wu_cat  [d, nR]  up.weight of the branches side by side               (fwd GEMM2, K-major B)
//   WuT_cat [nR, d]  its transpose                                        (bwd dH GEMM, K-major B)
//   bd_cat  [nR] fp32, bu_cat [d] fp32 = sum of the branches' up biases.
// A few hundred KB per adapter site; one launch, element-per-thread, reads served by L2.
#include "feddat_b200.h"
#include "host_common.h"
#include <cuda_bf16.h>

namespace fd {
namespace {

struct PackParams {
  const float* down_w[2];
  const float* down_b[2];
  const float* up_w[2];
  const float* up_b[2];
  int nb, r, d;
  __nv_bfloat16 *Wd, *WdT, *Wu, *WuT;
  float *bd, *bu;
};

__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
  const int R = p.nb * p.r, d = p.d;
  const int total = R * d;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    {  // [R, d] indexing: idx = j * d + c
      const int j = idx / d, c = idx - j * d;
      const int b = j / p.r, jj = j - b * p.r;
      if (p.Wd) p.Wd[idx] = __float2bfloat16_rn(p.down_w[b][jj * d + c]);
      if (p.WuT) p.WuT[idx] = __float2bfloat16_rn(p.up_w[b][c * p.r + jj]);
    }
    {  // [d, R] indexing: idx = c * R + j
      const int c = idx / R, j = idx - c * R;
      const int b = j / p.r, jj = j - b * p.r;
      if (p.Wu) p.Wu[idx] = __float2bfloat16_rn(p.up_w[b][c * p.r + jj]);
      if (p.WdT) p.WdT[idx] = __float2bfloat16_rn(p.down_w[b][jj * d + c]);
    }
    if (idx < R && p.bd) {
      const int b = idx / p.r;
      p.bd[idx] = p.down_b[b][idx - b * p.r];
    }
    if (idx < d && p.bu) {
      float s = p.up_b[0][idx];
      if (p.nb == 2) s += p.up_b[1][idx];
      p.bu[idx] = s;
    }
  }
}

}  // namespace
}  // namespace fd

extern "C" int feddat_pack_weights(const float* const* down_w, const float* const* down_b,
                                   const float* const* up_w, const float* const* up_b, int n_branch,
                                   int r, int d, void* Wd_cat, void* WdT_cat, void* Wu_cat,
                                   void* WuT_cat, float* bd_cat, float* bu_cat, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(n_branch == 1 || n_branch == 2, FD_ERR_INVALID, "pack_weights: n_branch must be 1 or 2");
  FD_REQUIRE(r >= 1 && d >= 1 && d >= n_branch * r, FD_ERR_INVALID,
             "pack_weights: bad shape r=%d d=%d", r, d);
  FD_REQUIRE(down_w && down_b && up_w && up_b, FD_ERR_INVALID, "pack_weights: null pointer table");
  PackParams p{};
  for (int b = 0; b < n_branch; ++b) {
    FD_REQUIRE(down_w[b] && down_b[b] && up_w[b] && up_b[b], FD_ERR_INVALID,
               "pack_weights: null weight pointer for branch %d", b);
    p.down_w[b] = down_w[b]; p.down_b[b] = down_b[b]; p.up_w[b] = up_w[b]; p.up_b[b] = up_b[b];
  }
  p.nb = n_branch; p.r = r; p.d = d;
  p.Wd = static_cast<__nv_bfloat16*>(Wd_cat); p.WdT = static_cast<__nv_bfloat16*>(WdT_cat);
  p.Wu = static_cast<__nv_bfloat16*>(Wu_cat); p.WuT = static_cast<__nv_bfloat16*>(WuT_cat);
  p.bd = bd_cat; p.bu = bu_cat;
  const int total = n_branch * r * d;
  int blocks = (total + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  pack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
