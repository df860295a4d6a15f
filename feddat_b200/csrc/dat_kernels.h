// Internal launch entry points shared between the DAT translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fd {

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
