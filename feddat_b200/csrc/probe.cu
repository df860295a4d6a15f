// Bring-up probe: one CTA, one 128 x N x K bf16 GEMM on tcgen05 with every operand source the
// production kernels rely on (K-major smem, MN-major smem, A from TMEM).  The descriptor strides can
// be overridden from the host so a single GPU call can sweep encodings.  Test infrastructure for
// tests/test_probe_gpu.py; not part of the hot path.
#ifdef FEDDAT_DEBUG   // bring-up / profiling kernels: only in libfeddat_sm100_dbg.so
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {

struct ProbeParams {
  int N, K;
  int a_mode;  // 0 = smem K-major, 1 = smem MN-major, 2 = TMEM
  int b_mode;  // 0 = smem K-major, 1 = smem MN-major
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // 0 = default
  uint32_t a_kstep, b_kstep;            // descriptor advance in bytes per UMMA_K=16 (0 = default)
  const __nv_bfloat16* A;               // used by a_mode == 2: row-major [128, K]
  float* D;                             // [128, N] fp32
};

__global__ void __launch_bounds__(128, 1)
probe_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K;
  const uint32_t a_base = smem_u32(smem);
  const uint32_t a_bytes = 128u * K * 2u;
  const uint32_t b_base = a_base + ((a_bytes + 1023u) & ~1023u);
  const uint32_t b_bytes = static_cast<uint32_t>(N) * K * 2u;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_done), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;

  if (tid == 0) {
    uint32_t tx = b_bytes + (p.a_mode == 2 ? 0u : a_bytes);
    mbar_arrive_expect_tx(smem_u32(&bar_full), tx);
    if (p.a_mode == 0) {
      for (int kc = 0; kc < K / 64; ++kc)
        tma_load_2d(a_base + kc * 128 * 128, &tmA, smem_u32(&bar_full), kc * 64, 0);
    } else if (p.a_mode == 1) {
      for (int mb = 0; mb < 2; ++mb)
        tma_load_2d(a_base + mb * K * 128, &tmA, smem_u32(&bar_full), mb * 64, 0);
    }
    if (p.b_mode == 0) {
      for (int kc = 0; kc < K / 64; ++kc)
        tma_load_2d(b_base + kc * N * 128, &tmB, smem_u32(&bar_full), kc * 64, 0);
    } else {
      for (int nb = 0; nb < N / 64; ++nb)
        tma_load_2d(b_base + nb * K * 128, &tmB, smem_u32(&bar_full), nb * 64, 0);
    }
  }
  if (p.a_mode == 2) {
    // A tile into TMEM columns [256, 256 + K/2): thread = row, bf16 pairs packed per column
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint4* arow = reinterpret_cast<const uint4*>(p.A + static_cast<size_t>(tid) * K);
    for (int c = 0; c < K / 16; ++c) {
      uint4 v0 = arow[2 * c], v1 = arow[2 * c + 1];
      uint32_t v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      tmem_st8(tmem + lane_base + 256 + c * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();

  if (tid == 0) {
    mbar_wait(smem_u32(&bar_full), 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N, p.a_mode == 1, p.b_mode == 1);
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t bdesc;
      if (p.b_mode == 0) {
        uint32_t step = p.b_kstep ? p.b_kstep * kk : (kk / 4) * N * 128 + (kk % 4) * 32;
        bdesc = make_smem_desc(b_base + step, p.b_lbo ? p.b_lbo : 16, p.b_sbo ? p.b_sbo : 1024,
                               kLayoutSw128);
      } else {
        uint32_t step = (p.b_kstep ? p.b_kstep : 2048u) * kk;
        bdesc = make_smem_desc(b_base + step, p.b_lbo ? p.b_lbo : K * 128,
                               p.b_sbo ? p.b_sbo : 1024, kLayoutSw128);
      }
      if (p.a_mode == 2) {
        umma_ts(tmem, tmem + 256 + kk * 8, bdesc, idesc, kk > 0);
      } else {
        uint64_t adesc;
        if (p.a_mode == 0) {
          uint32_t step = p.a_kstep ? p.a_kstep * kk : (kk / 4) * 128 * 128 + (kk % 4) * 32;
          adesc = make_smem_desc(a_base + step, p.a_lbo ? p.a_lbo : 16, p.a_sbo ? p.a_sbo : 1024,
                                 kLayoutSw128);
        } else {
          uint32_t step = (p.a_kstep ? p.a_kstep : 2048u) * kk;
          adesc = make_smem_desc(a_base + step, p.a_lbo ? p.a_lbo : K * 128,
                                 p.a_sbo ? p.a_sbo : 1024, kLayoutSw128);
        }
        umma_ss(tmem, adesc, bdesc, idesc, kk > 0);
      }
    }
    umma_commit(smem_u32(&bar_done));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_done), 0);
  tc_fence_after();

  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  float* drow = p.D + static_cast<size_t>(tid) * N;
  for (int c = 0; c < N / 16; ++c) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_base + c * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      *reinterpret_cast<float4*>(drow + c * 16 + i) =
          make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                      __uint_as_float(v[i + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace fd

// A: a_mode 0/2 -> row-major [128, K]; a_mode 1 -> row-major [K, 128].
// B: b_mode 0 -> row-major [N, K];     b_mode 1 -> row-major [K, N].   D = A * B^T, [128, N] fp32.
extern "C" int feddat_probe_gemm(const void* A, const void* B, float* D, int N, int K, int a_mode,
                                 int b_mode, const uint32_t* overrides /* 6 values or NULL */,
                                 void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 64 == 0 && K >= 64 && K <= 192 &&
                 (b_mode == 0 || N % 64 == 0),
             FD_ERR_INVALID, "probe: N=%d K=%d out of range", N, K);
  CUtensorMap tmA, tmB;
  if (a_mode == 1)
    rc = make_tmap_bf16_2d(&tmA, A, K, 128, 128, K, 64);
  else
    rc = make_tmap_bf16_2d(&tmA, A, 128, K, K, 128, 64);
  if (rc) return rc;
  if (b_mode == 1)
    rc = make_tmap_bf16_2d(&tmB, B, K, N, N, K, 64);
  else
    rc = make_tmap_bf16_2d(&tmB, B, N, K, K, N, 64);
  if (rc) return rc;
  ProbeParams p{};
  p.N = N; p.K = K; p.a_mode = a_mode; p.b_mode = b_mode;
  if (overrides) {
    p.a_lbo = overrides[0]; p.a_sbo = overrides[1]; p.b_lbo = overrides[2]; p.b_sbo = overrides[3];
    p.a_kstep = overrides[4]; p.b_kstep = overrides[5];
  }
  p.A = reinterpret_cast<const __nv_bfloat16*>(A);
  p.D = D;
  size_t smem = 1024 + ((128 * K * 2 + 1023) & ~1023) + static_cast<size_t>(N) * K * 2;
  FD_CHECK_CUDA(cudaFuncSetAttribute(probe_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  probe_gemm_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------
// L2 -> SM streaming bandwidth probe: every CTA TMA-loads [128 x 64] bf16 boxes (16 KB) from a small
// (L2-resident) matrix through an 8-slot ring and discards them.  cluster > 1: each CTA fetches
// 1/cluster of every box and multicasts it to the whole cluster, so L2 is read once per cluster.
// ------------------------------------------------------------------------------------------------
namespace fd {

__global__ void __launch_bounds__(64, 1)
probe_l2bw_kernel(const __grid_constant__ CUtensorMap tm, int n_boxes, int iters, int cluster) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  constexpr int NS = 8;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t rank = cluster > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar0 + 8 * s, 1);                 // full
      mbar_init(bar0 + 8 * (NS + s), cluster);    // empty: one arrive per CTA of the cluster
    }
    fence_mbar_init();
  }
  if (cluster > 1) cluster_sync_all(); else __syncthreads();
  const int rows_per = 128 / cluster;
  if (threadIdx.x == 0) {          // producer
    for (int i = 0; i < iters; ++i) {
      const int s = i % NS;
      const uint32_t par = (i / NS) & 1;
      mbar_wait(bar0 + 8 * (NS + s), par ^ 1);
      mbar_arrive_expect_tx(bar0 + 8 * s, 16384);
      const int box = (i * 7 + blockIdx.x / cluster * 3) % n_boxes;
      if (cluster > 1)
        tma_load_2d_mcast(smem0 + s * 16384 + rank * rows_per * 128, &tm, bar0 + 8 * s, 0,
                          box * 128 + rank * rows_per, (uint16_t)((1u << cluster) - 1), kEvictLast);
      else
        tma_load_2d_hint(smem0 + s * 16384, &tm, bar0 + 8 * s, 0, box * 128, kEvictLast);
    }
  } else if (threadIdx.x == 32) {  // consumer
    for (int i = 0; i < iters; ++i) {
      const int s = i % NS;
      const uint32_t par = (i / NS) & 1;
      mbar_wait(bar0 + 8 * s, par);
      if (cluster > 1) {
        for (int r = 0; r < cluster; ++r) mbar_arrive_remote(bar0 + 8 * (NS + s), r);
      } else {
        mbar_arrive(bar0 + 8 * (NS + s));
      }
    }
  }
  if (cluster > 1) cluster_sync_all(); else __syncthreads();
}

}  // namespace fd

// buf: bf16 [n_boxes * 128, 64] row-major.  Returns after enqueueing `grid` CTAs x `iters` boxes.
extern "C" int feddat_probe_l2bw(const void* buf, int n_boxes, int iters, int grid, int cluster,
                                 void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8, FD_ERR_INVALID,
             "probe_l2bw: cluster must be 1, 2, 4 or 8");
  FD_REQUIRE(grid % cluster == 0, FD_ERR_INVALID, "probe_l2bw: grid must be a multiple of cluster");
  CUtensorMap tm;
  rc = make_tmap_bf16_2d(&tm, buf, static_cast<uint64_t>(n_boxes) * 128, 64, 64, 128 / cluster, 64);
  if (rc) return rc;
  const size_t smem = 1024 + 8 * 16384;
  FD_CHECK_CUDA(cudaFuncSetAttribute(probe_l2bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, probe_l2bw_kernel, tm, n_boxes, iters, cluster));
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------
// Per-SM TMA ingest probe: `ns` slots of [box_rows x 64] bf16 boxes (128-byte swizzle) streamed from a
// [n_rows, row_stride] tensor; n_prod producer threads (slot s belongs to producer s % n_prod), one
// consumer thread that releases slots in order.  Separates "ring too shallow" (rate grows with ns)
// from a per-SM or chip-wide bandwidth cap (rate independent of ns; per-SM rate vs grid size).
// ------------------------------------------------------------------------------------------------
namespace fd {

__global__ void __launch_bounds__(160, 1)
probe_ingest_kernel(const __grid_constant__ CUtensorMap tm, int n_rows, int n_colblk, int iters, int ns,
                    int box_rows, int n_prod, long long* issue_clk) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[32];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t slot_bytes = static_cast<uint32_t>(box_rows) * 128u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) {
      mbar_init(bar0 + 8 * s, 1);
      mbar_init(bar0 + 8 * (16 + s), 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rowblk = n_rows / box_rows;
  if (warp < n_prod && lane == 0) {
    // all indices advance incrementally: no division on the issue path
    int s = warp % ns;
    uint32_t par = 0;
    int rb = (blockIdx.x * 131 + warp * 7) % n_rowblk, cb = blockIdx.x % n_colblk;
    long long t_issue = 0;
    for (int i = warp; i < iters; i += n_prod) {
      mbar_wait(bar0 + 8 * (16 + s), par ^ 1);
      const long long c0 = clock64();
      mbar_arrive_expect_tx(bar0 + 8 * s, slot_bytes);
      tma_load_2d_hint(smem0 + s * slot_bytes, &tm, bar0 + 8 * s, cb * 64, rb * box_rows, kEvictLast);
      t_issue += clock64() - c0;
      s += n_prod;
      if (s >= ns) { s -= ns; par ^= 1; }
      rb += 7 * n_prod;
      if (rb >= n_rowblk) { rb -= n_rowblk; if (++cb == n_colblk) cb = 0; }
    }
    if (issue_clk != nullptr && blockIdx.x == 0 && warp == 0)
      issue_clk[0] = t_issue / ((iters + n_prod - 1) / n_prod);
  } else if (warp == 4 && lane == 0) {
    int cs = 0;
    uint32_t cpar = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bar0 + 8 * cs, cpar);
      mbar_arrive(bar0 + 8 * (16 + cs));
      if (++cs == ns) { cs = 0; cpar ^= 1; }
    }
  }
  __syncthreads();
}

}  // namespace fd

// buf: bf16 [n_rows, row_stride] row-major (row_stride a multiple of 64).
extern "C" int feddat_probe_ingest(const void* buf, int n_rows, int row_stride, int iters, int grid,
                                   int ns, int box_rows, int n_prod, long long* issue_clk,
                                   void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(ns >= 1 && ns <= 16 && n_prod >= 1 && n_prod <= 4 && ns % n_prod == 0 && box_rows >= 8 && box_rows <= 256 &&
                 row_stride % 64 == 0 && n_rows % box_rows == 0,
             FD_ERR_INVALID, "probe_ingest: bad arguments");
  const size_t smem = 1024 + static_cast<size_t>(ns) * box_rows * 128;
  FD_REQUIRE(smem <= 226 * 1024, FD_ERR_INVALID, "probe_ingest: ring does not fit in shared memory");
  CUtensorMap tm;
  rc = make_tmap_bf16_2d(&tm, buf, n_rows, row_stride, row_stride, box_rows, 64);
  if (rc) return rc;
  FD_CHECK_CUDA(cudaFuncSetAttribute(probe_ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     226 * 1024));
  probe_ingest_kernel<<<grid, 160, smem, static_cast<cudaStream_t>(stream)>>>(
      tm, n_rows, row_stride / 64, iters, ns, box_rows, n_prod, issue_clk);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------
// Tile-copy probe: the DAT kernels' HBM access pattern without any compute.  Persistent CTAs walk
// 128-row tiles of a [M, 768] bf16 tensor; each tile is read as twelve [128 x 64] TMA boxes (k-chunk
// order, like GEMM1) and written back to a second tensor as twelve [128 x 64] TMA stores (like the
// output chunks).  Separates "this access pattern cannot reach the copy bandwidth" from "the compute
// pipeline leaves HBM idle".  mode 1: read only.
// ------------------------------------------------------------------------------------------------
namespace fd {

__global__ void __launch_bounds__(96, 1)
probe_tilecopy_kernel(const __grid_constant__ CUtensorMap tmSrc, const __grid_constant__ CUtensorMap tmDst,
                      int num_tiles, int ns, int read_only) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[32];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) {
      mbar_init(bar0 + 8 * s, 1);
      mbar_init(bar0 + 8 * (16 + s), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmSrc);
    tma_prefetch_desc(&tmDst);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    int s = 0;
    uint32_t par = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x)
      for (int kc = 0; kc < 12; ++kc) {
        mbar_wait(bar0 + 8 * (16 + s), par ^ 1);
        mbar_arrive_expect_tx(bar0 + 8 * s, 16384);
        tma_load_2d(smem0 + s * 16384, &tmSrc, bar0 + 8 * s, kc * 64, t * 128);
        if (++s == ns) { s = 0; par ^= 1; }
      }
  } else if (warp == 1 && lane == 0) {
    int s = 0, prev = -1;
    uint32_t par = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x)
      for (int kc = 0; kc < 12; ++kc) {
        mbar_wait(bar0 + 8 * s, par);
        if (read_only) {
          mbar_arrive(bar0 + 8 * (16 + s));
        } else {
          tma_store_2d(&tmDst, smem0 + s * 16384, kc * 64, t * 128);
          tma_store_commit();
          if (prev >= 0) {
            tma_store_wait_read<1>();
            mbar_arrive(bar0 + 8 * (16 + prev));
          }
          prev = s;
        }
        if (++s == ns) { s = 0; par ^= 1; }
      }
    if (!read_only && prev >= 0) {
      tma_store_wait_read<0>();
      mbar_arrive(bar0 + 8 * (16 + prev));
      tma_store_wait_all<0>();
    }
  }
  __syncthreads();
}

}  // namespace fd

extern "C" int feddat_probe_tilecopy(const void* src, void* dst, int64_t M, int grid, int ns, int read_only,
                                     void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(ns >= 2 && ns <= 13 && M > 0 && grid > 0, FD_ERR_INVALID, "probe_tilecopy: bad arguments");
  CUtensorMap tmS, tmD;
  if ((rc = make_tmap_bf16_2d(&tmS, src, M, 768, 768, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmD, dst, M, 768, 768, 128, 64))) return rc;
  const size_t smem = 1024 + static_cast<size_t>(ns) * 16384;
  FD_CHECK_CUDA(cudaFuncSetAttribute(probe_tilecopy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  probe_tilecopy_kernel<<<grid, 96, smem, static_cast<cudaStream_t>(stream)>>>(
      tmS, tmD, static_cast<int>((M + 127) / 128), ns, read_only);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------
// cta_group::2 bring-up: one CTA pair, D[256 x N] = A[256 x K] * B[N x K]^T.  Each CTA holds its
// 128 rows of A (smem, or TMEM when a_tmem) and N/2 rows of B; the leader CTA issues the MMAs.
// ------------------------------------------------------------------------------------------------
namespace fd {

__global__ void __launch_bounds__(128, 1)
probe_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __nv_bfloat16* A, float* D, int N, int K, int a_tmem, int reps,
                  unsigned long long* ns_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem0;
  const uint32_t a_bytes = 128u * K * 2u;
  const uint32_t b_base = a_base + a_bytes;
  const int NH = N / 2;
  const uint32_t b_bytes = static_cast<uint32_t>(NH) * K * 2u;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_done), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  if (tid == 0) {
    const uint32_t leader_full = mapa_u32(smem_u32(&bar_full), 0);
    if (rank == 0)
      mbar_arrive_expect_tx(smem_u32(&bar_full), 2 * (b_bytes + (a_tmem ? 0u : a_bytes)));
    if (!a_tmem)
      for (int kc = 0; kc < K / 64; ++kc)
        tma_load_2d_pair(a_base + kc * 128 * 128, &tmA, leader_full, kc * 64, rank * 128,
                         kEvictNormal);
    for (int kc = 0; kc < K / 64; ++kc)
      tma_load_2d_pair(b_base + kc * NH * 128, &tmB, leader_full, kc * 64, rank * NH, kEvictNormal);
  }
  if (a_tmem) {
    const uint4* arow =
        reinterpret_cast<const uint4*>(A + static_cast<size_t>(rank * 128 + tid) * K);
    for (int c = 0; c < K / 16; ++c) {
      uint4 v0 = arow[2 * c], v1 = arow[2 * c + 1];
      uint32_t v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      tmem_st8(tmem + lane_base + 256 + c * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  cluster_sync_all();  // both CTAs' TMEM A operands are written before the leader issues

  if (rank == 0 && tid == 0) {
    mbar_wait(smem_u32(&bar_full), 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(256, N);
    const uint64_t t0 = globaltimer_ns();
    for (int rep = 0; rep < reps; ++rep)   // reps > 1: MMA-throughput timing (D is then reps * A B^T)
      for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t bdesc = desc_kmajor_sw128(b_base + (kk / 4) * NH * 128 + (kk % 4) * 32);
        if (a_tmem)
          umma_ts_pair(tmem, tmem + 256 + kk * 8, bdesc, idesc, (rep | kk) > 0);
        else
          umma_ss_pair(tmem, desc_kmajor_sw128(a_base + (kk / 4) * 128 * 128 + (kk % 4) * 32),
                       bdesc, idesc, (rep | kk) > 0);
      }
    umma_commit_pair(smem_u32(&bar_done), 0b11);
    mbar_wait(smem_u32(&bar_done), 0);
    if (ns_out) *ns_out = globaltimer_ns() - t0;
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_done), 0);
  tc_fence_after();
  float* drow = D + static_cast<size_t>(rank * 128 + tid) * N;
  for (int c = 0; c < N / 16; ++c) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_base + c * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      *reinterpret_cast<float4*>(drow + c * 16 + i) =
          make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                      __uint_as_float(v[i + 3]));
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 512);
}

}  // namespace fd

// A: bf16 [256, K] row-major; B: bf16 [N, K] row-major; D: fp32 [256, N].
extern "C" int feddat_probe_pair(const void* A, const void* B, float* D, int N, int K, int a_tmem,
                                 int reps, unsigned long long* ns_out, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 64 == 0 && K >= 64 && K <= 256,
             FD_ERR_INVALID, "probe_pair: N=%d K=%d out of range", N, K);
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, A, 256, K, K, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, B, N, K, K, N / 2, 64))) return rc;
  const size_t smem = 1024 + 128 * K * 2 + static_cast<size_t>(N / 2) * K * 2;
  FD_CHECK_CUDA(cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, probe_pair_kernel, tmA, tmB,
                                   static_cast<const __nv_bfloat16*>(A), D, N, K, a_tmem,
                                   reps < 1 ? 1 : reps, ns_out));
  return FD_OK;
}
#endif  // FEDDAT_DEBUG
