/* feddat_b200_debug.h -- bring-up probes and debug switches of libfeddat_sm100_dbg.so
 *
 * NOT part of the product ABI (include/feddat_b200.h).  The debug twin of the library is the same
 * sources compiled with -DFEDDAT_DEBUG: it adds the globaltimer trace hooks inside the DAT kernels, the
 * process-global trace pointer / kernel-selection switch below and the tcgen05 / TMA bring-up probes of
 * csrc/probe.cu.  Used by tests/test_kernels_gpu.py::test_tcgen05_probe and scripts/*.py only.
 */
#ifndef FEDDAT_B200_DEBUG_H_
#define FEDDAT_B200_DEBUG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A/B measurement switch (tests / profiling): 1 = always use the single-tile forward kernel, 0 = pick the
 * tile-pipelined forward kernel when there are more 256-row super-tiles than CTA pairs (default). */
int feddat_debug_force_fused_fwd(int on);

/* Bring-up probe (tests only): one 128 x N x K tcgen05 GEMM, see csrc/probe.cu. */
int feddat_probe_gemm(const void* A, const void* B, float* D, int N, int K, int a_mode,
                      int b_mode, const uint32_t* overrides, void* stream);

/* cta_group::2 bring-up probe (tests only): D[256,N] = A[256,K] B[N,K]^T on one CTA pair. */
int feddat_probe_pair(const void* A, const void* B, float* D, int N, int K, int a_tmem, int reps,
                      unsigned long long* ns_out /* device, nullable: MMA loop time */, void* stream);

/* L2 -> SM TMA streaming bandwidth probe (optionally multicast across a cluster), csrc/probe.cu. */
int feddat_probe_l2bw(const void* buf, int n_boxes, int iters, int grid, int cluster, void* stream);
/* HBM access-pattern probe: tiles of a [M, 768] bf16 tensor read as [128 x 64] TMA boxes and written back
 * the same way, no compute (profiling only). */
int feddat_probe_tilecopy(const void* src, void* dst, int64_t M, int grid, int ns, int read_only, void* stream);
/* Per-SM TMA ingest sweep (ring depth, box height, producer count, grid size); bring-up / profiling only. */
int feddat_probe_ingest(const void* buf, int n_rows, int row_stride, int iters, int grid, int ns,
                        int box_rows, int n_prod, long long* issue_clk, void* stream);

/* Debug (scripts/trace_kernel.py): device buffer of 256 uint64 that receives globaltimer stamps of
 * CTA 0's pipeline events in feddat_dat_fwd / feddat_dat_bwd_dgrad; NULL disables. */
int feddat_debug_set_trace(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* FEDDAT_B200_DEBUG_H_ */
