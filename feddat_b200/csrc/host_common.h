// Host-side helpers shared by the C-ABI entry points: thread-local last-error string, CUDA error
// checks that never throw / exit, and TMA tensor-map construction through the driver entry point
// (resolved at run time, so the library links against cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace fd {

enum : int {
  FD_OK = 0,
  FD_ERR_INVALID = -1,      // bad argument (shape / alignment / null pointer)
  FD_ERR_UNSUPPORTED = -2,  // configuration the sm_100a kernels do not cover (no fallback exists)
  FD_ERR_CUDA = -3,         // CUDA runtime / driver failure
  FD_ERR_NO_DEVICE = -4,    // not an sm_100 device
};

char* last_error_buf();  // thread-local, 512 bytes
int set_error(int code, const char* fmt, ...);

#define FD_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fd::set_error(fd::FD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)

#define FD_REQUIRE(cond, code, ...)                      \
  do {                                                   \
    if (!(cond)) return fd::set_error((code), __VA_ARGS__); \
  } while (0)

// 2-D row-major bf16 tensor [rows, cols] (cols contiguous) -> tensor map with a
// [box_rows x box_cols] box and 128-byte swizzle (box_cols * 2 bytes must be 128).
int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);

// The same [rows, cols] tensor (cols a multiple of 64) viewed as 3-D (64, rows, cols / 64): one box =
// [box_rows x 64] of `box_blocks` consecutive 64-column blocks, landing in smem as that many
// consecutive 128B-swizzled [box_rows x 64] tiles -- a whole MMA operand in one TMA instruction.
int make_tmap_bf16_kblocks(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols,
                           uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_blocks);

#ifdef FEDDAT_DEBUG
// debug build only (libfeddat_sm100_dbg.so): device buffer of 256 uint64 receiving pipeline timestamps of
// CTA 0 (NULL = off).  The product library carries no process-global mutable state.
extern unsigned long long* g_trace;
#define FD_TRACE_PTR fd::g_trace
#else
#define FD_TRACE_PTR nullptr
#endif

int device_sm_count(int* out);
int check_device_sm100();

}  // namespace fd
