// Packs the fp32 master weights of the active adapter branches (nn.Linear layout, reference
// adapter.py:35,41) into the bf16 operand layouts the tcgen05 kernels consume:
//   Wd_cat  [nR, d]  rows of down.weight stacked branch after branch      (fwd GEMM1, K-major B)
//   WdT_cat [d, nR]  its transpose                                        (bwd dX GEMM, K-major B)
//   Wu_cat  [d, nR]  up.weight of the branches side by side               (fwd GEMM2, K-major B)
//   WuT_cat [nR, d]  its transpose                                        (bwd dH GEMM, K-major B)
//   bd_cat  [nR] fp32, bu_cat [d] fp32 = sum of the given up biases.
// A "job" is one such operand set; feddat_pack_weights_batched runs up to kMaxJobs jobs per launch, so a
// train step packs all of its adapter sites in both modes (12 sites x {gating pair, adapter_1} for ViLT)
// with ONE launch when the weights change -- once per optimizer round trip, not once per forward.
#include "feddat_b200.h"
#include "host_common.h"
#include <cuda_bf16.h>

namespace fd {
namespace {

constexpr int kMaxJobs = 24;

struct PackJob {
  const float* down_w[2];   // [r, d] each (row stride d)
  const float* down_b[2];   // [r]
  const float* up_w[2];     // [d, r] each, row stride ld_up (a column slice of a wider matrix is allowed)
  const float* bu_src[2];   // up biases summed into bu_cat (either may be null)
  __nv_bfloat16 *Wd, *WdT, *Wu, *WuT;
  float *bd, *bu;
  int nb, r, ld_up;
};

struct PackParams {
  PackJob job[kMaxJobs];
  int first_block[kMaxJobs + 1];   // prefix sums of the jobs' block counts
  int n_jobs, d;
};

// One 32 x 32 tile of one source matrix per block (256 threads = 32 x 8): coalesced fp32 reads, the
// straight copy written coalesced, the transposed copy through a padded smem tile (also coalesced).
// Within a job, tiles [0, nt) cover the stacked down weights [R, d] (-> Wd_cat, WdT_cat), tiles [nt, 2 nt)
// the side-by-side up weights [d, R] (-> Wu_cat, WuT_cat); the job's last block writes the biases.
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackParams p) {
  __shared__ float tile[32][33];
  int ji = 0;
  while (ji + 1 < p.n_jobs && static_cast<int>(blockIdx.x) >= p.first_block[ji + 1]) ++ji;
  const PackJob& jb = p.job[ji];
  const int R = jb.nb * jb.r, d = p.d, r = jb.r;
  const int tr_n = (R + 31) / 32, tc_n = (d + 31) / 32;      // tiles over [R, d]
  const int nt = tr_n * tc_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int t = static_cast<int>(blockIdx.x) - p.first_block[ji];
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
          v = jb.down_w[b][static_cast<size_t>(j - b * r) * d + c];
          if (jb.Wd) jb.Wd[static_cast<size_t>(j) * d + c] = __float2bfloat16_rn(v);
        }
        tile[ty + 8 * k][tx] = v;
      }
      __syncthreads();
      if (jb.WdT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + ty + 8 * k, j = j0 + tx;
          if (j < R && c < d) jb.WdT[static_cast<size_t>(c) * R + j] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
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
          v = jb.up_w[b][static_cast<size_t>(c) * jb.ld_up + (j - b * r)];
          if (jb.Wu) jb.Wu[static_cast<size_t>(c) * R + j] = __float2bfloat16_rn(v);
        }
        tile[ty + 8 * k][tx] = v;
      }
      __syncthreads();
      if (jb.WuT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + ty + 8 * k, c = c0 + tx;
          if (j < R && c < d) jb.WuT[static_cast<size_t>(j) * d + c] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
        }
      }
    }
    return;
  }
  // bias block
  if (jb.bd)
    for (int idx = threadIdx.x; idx < R; idx += blockDim.x) {
      const int b = idx / r;
      jb.bd[idx] = jb.down_b[b][idx - b * r];
    }
  if (jb.bu)
    for (int idx = threadIdx.x; idx < d; idx += blockDim.x) {
      float s2 = 0.f;                       // same order as torch's `up_b[0] + up_b[1]`
      if (jb.bu_src[0]) s2 = jb.bu_src[0][idx];
      if (jb.bu_src[1]) s2 += jb.bu_src[1][idx];
      jb.bu[idx] = s2;
    }
}

int check_job(const FeddatPackJob& j, int d, int idx) {
  FD_REQUIRE(j.n_branch == 1 || j.n_branch == 2, FD_ERR_INVALID, "pack_weights: job %d: n_branch must be 1 or 2", idx);
  FD_REQUIRE(j.r >= 1 && d >= 1, FD_ERR_INVALID, "pack_weights: job %d: bad shape r=%d d=%d", idx, j.r, d);
  FD_REQUIRE(j.ld_up == 0 || j.ld_up >= j.r, FD_ERR_INVALID, "pack_weights: job %d: ld_up=%d < r=%d", idx, j.ld_up, j.r);
  for (int b = 0; b < j.n_branch; ++b)
    FD_REQUIRE(j.down_w[b] && j.down_b[b] && j.up_w[b], FD_ERR_INVALID,
               "pack_weights: job %d: null weight pointer for branch %d", idx, b);
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int feddat_pack_weights_batched(const FeddatPackJob* jobs, int n_jobs, int d, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(jobs != nullptr && n_jobs >= 0, FD_ERR_INVALID, "pack_weights_batched: null job table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int j0 = 0; j0 < n_jobs; j0 += kMaxJobs) {
    PackParams p{};
    p.d = d;
    p.n_jobs = n_jobs - j0 < kMaxJobs ? n_jobs - j0 : kMaxJobs;
    int blocks = 0;
    for (int i = 0; i < p.n_jobs; ++i) {
      const FeddatPackJob& s = jobs[j0 + i];
      if ((rc = check_job(s, d, j0 + i))) return rc;
      PackJob& t = p.job[i];
      for (int b = 0; b < 2; ++b) {
        t.down_w[b] = b < s.n_branch ? s.down_w[b] : nullptr;
        t.down_b[b] = b < s.n_branch ? s.down_b[b] : nullptr;
        t.up_w[b] = b < s.n_branch ? s.up_w[b] : nullptr;
        t.bu_src[b] = s.bu_src[b];
      }
      t.Wd = static_cast<__nv_bfloat16*>(s.Wd_cat); t.WdT = static_cast<__nv_bfloat16*>(s.WdT_cat);
      t.Wu = static_cast<__nv_bfloat16*>(s.Wu_cat); t.WuT = static_cast<__nv_bfloat16*>(s.WuT_cat);
      t.bd = s.bd_cat; t.bu = s.bu_cat;
      t.nb = s.n_branch; t.r = s.r; t.ld_up = s.ld_up ? s.ld_up : s.r;
      p.first_block[i] = blocks;
      const int R = s.n_branch * s.r;
      blocks += 2 * ((R + 31) / 32) * ((d + 31) / 32) + 1;
    }
    p.first_block[p.n_jobs] = blocks;
    if (blocks == 0) continue;
    pack_kernel<<<blocks, 256, 0, st>>>(p);
    FD_CHECK_CUDA(cudaGetLastError());
  }
  return FD_OK;
}

extern "C" int feddat_pack_weights(const float* const* down_w, const float* const* down_b,
                                   const float* const* up_w, const float* const* up_b, int n_branch,
                                   int r, int d, void* Wd_cat, void* WdT_cat, void* Wu_cat,
                                   void* WuT_cat, float* bd_cat, float* bu_cat, void* stream) {
  using namespace fd;
  FD_REQUIRE(n_branch == 1 || n_branch == 2, FD_ERR_INVALID, "pack_weights: n_branch must be 1 or 2");
  FD_REQUIRE(r >= 1 && d >= 1 && d >= n_branch * r, FD_ERR_INVALID,
             "pack_weights: bad shape r=%d d=%d", r, d);
  FD_REQUIRE(down_w && down_b && up_w && up_b, FD_ERR_INVALID, "pack_weights: null pointer table");
  FeddatPackJob j{};
  for (int b = 0; b < n_branch; ++b) {
    FD_REQUIRE(down_w[b] && down_b[b] && up_w[b] && up_b[b], FD_ERR_INVALID,
               "pack_weights: null weight pointer for branch %d", b);
    j.down_w[b] = down_w[b]; j.down_b[b] = down_b[b]; j.up_w[b] = up_w[b]; j.bu_src[b] = up_b[b];
  }
  j.n_branch = n_branch; j.r = r; j.ld_up = r;
  j.Wd_cat = Wd_cat; j.WdT_cat = WdT_cat; j.Wu_cat = Wu_cat; j.WuT_cat = WuT_cat;
  j.bd_cat = bd_cat; j.bu_cat = bu_cat;
  return feddat_pack_weights_batched(&j, 1, d, stream);
}
