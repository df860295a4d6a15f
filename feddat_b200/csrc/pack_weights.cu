// Packs the fp32 master weights of the active adapter branches (nn.Linear layout, reference
// adapter.py:35,41) into the bf16 operand layouts the tcgen05 kernels consume:
//   Wd_cat  [nR, d]  rows of down.weight stacked branch after branch      (fwd GEMM1, K-major B)
//   WdT_cat [d, nR]  its transpose                                        (bwd dX GEMM, K-major B)
//   Wu_cat  [d, nR]  up.weight of the branches side by side               (fwd GEMM2, K-major B)
//   WuT_cat [nR, d]  its transpose                                        (bwd dH GEMM, K-major B)
//   bd_cat  [nR] fp32, bu_cat [d] fp32 = sum of the branches' up biases.
// A few hundred KB per adapter site; one launch of 32 x 32 transpose tiles.
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

// One 32 x 32 tile of one source matrix per block (256 threads = 32 x 8): coalesced fp32 reads, the
// straight copy written coalesced, the transposed copy through a padded smem tile (also coalesced).
// Tiles [0, nt_d) cover the stacked down weights [R, d] (-> Wd_cat, WdT_cat), tiles [nt_d, 2 nt_d) the
// side-by-side up weights [d, R] (-> Wu_cat, WuT_cat); the last block also writes the biases.
// (The first version indexed element-per-thread with strided fp32 reads: 7.5 us per site, 24 sites per
// train step.)
__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
  __shared__ float tile[32][33];
  const int R = p.nb * p.r, d = p.d, r = p.r;
  const int tr_n = (R + 31) / 32, tc_n = (d + 31) / 32;      // tiles over [R, d]
  const int nt = tr_n * tc_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int t = blockIdx.x;
  if (t < 2 * nt) {
    const bool up = t >= nt;
    if (up) t -= nt;
    // "row" runs over the bottleneck dimension j, "col" over the model dimension c, for both halves
    const int j0 = (t / tc_n) * 32, c0 = (t % tc_n) * 32;
    if (!up) {
      // down_w[b][jj, c]: contiguous in c
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j0 + ty + 8 * k, c = c0 + tx;
        float v = 0.f;
        if (j < R && c < d) {
          const int b = j / r;
          v = p.down_w[b][static_cast<size_t>(j - b * r) * d + c];
          if (p.Wd) p.Wd[static_cast<size_t>(j) * d + c] = __float2bfloat16_rn(v);
        }
        tile[ty + 8 * k][tx] = v;
      }
      __syncthreads();
      if (p.WdT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + ty + 8 * k, j = j0 + tx;
          if (j < R && c < d) p.WdT[static_cast<size_t>(c) * R + j] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
        }
      }
    } else {
      // up_w[b][c, jj]: contiguous in jj
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, j = j0 + tx;
        float v = 0.f;
        if (j < R && c < d) {
          const int b = j / r;
          v = p.up_w[b][static_cast<size_t>(c) * r + (j - b * r)];
          if (p.Wu) p.Wu[static_cast<size_t>(c) * R + j] = __float2bfloat16_rn(v);
        }
        tile[ty + 8 * k][tx] = v;
      }
      __syncthreads();
      if (p.WuT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + ty + 8 * k, c = c0 + tx;
          if (j < R && c < d) p.WuT[static_cast<size_t>(j) * d + c] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
        }
      }
    }
    return;
  }
  // bias block
  for (int idx = threadIdx.x; idx < R; idx += blockDim.x)
    if (p.bd) {
      const int b = idx / r;
      p.bd[idx] = p.down_b[b][idx - b * r];
    }
  for (int idx = threadIdx.x; idx < d; idx += blockDim.x)
    if (p.bu) {
      float s2 = p.up_b[0][idx];
      if (p.nb == 2) s2 += p.up_b[1][idx];
      p.bu[idx] = s2;
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
  const int R = n_branch * r;
  const int blocks = 2 * ((R + 31) / 32) * ((d + 31) / 32) + 1;
  pack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
