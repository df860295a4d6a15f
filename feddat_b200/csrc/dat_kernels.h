// Internal launch entry points shared between the DAT translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fd {

// Epilogue 1 splits the R / 16 sixteen-column chunks of P between the two epilogue groups: group A packs chunks
// [0, nA), group B the rest.  nA is rounded up to a multiple of four (64 columns) so that both groups' shares start at
// a 64-column boundary: the packed hidden also travels through the 64-column staging buffers (saved by the forward /
// loaded and turned into dP by the backward with TMA; "hidden chunks" in dat_fused.cu).
__host__ __device__ __forceinline__ int split_a(int n16) {
  const int a = (((n16 + 1) / 2) + 3) & ~3;
  return a < n16 ? a : n16;
}

// dat_fwd_pipe.cu: forward software-pipelined across tiles (every CTA pair owns >= 2 super-tiles).
// `grid` = 2 x (CTA pairs to launch).
int launch_dat_fwd_pipe(const void* X, const void* Res, void* Y, const void* Wd_cat, const float* bd_cat,
                        const void* Wu_cat, const float* bu_cat, void* H_out, int64_t M, int r_total, float scale,
                        int act, int grid, cudaStream_t st);

// Saved-mode backward data gradient on the same pipeline (ReLU, dX requested).
int launch_dat_bwd_pipe(const void* dY, void* dX, const void* WuT_cat, const void* WdT_cat, const void* H_in,
                        void* dP_t, int ld_t, int r_lo, int r_hi, int64_t M, int r_total, float scale, int add_dy,
                        int grid, cudaStream_t st);

}  // namespace fd
