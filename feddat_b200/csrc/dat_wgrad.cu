// DAT bottleneck backward, weight gradients of the trainable branch (autograd of reference
// adapter.py:124-163; trainability per adapter.py:71-85).  Inputs are the activations X, dY and the
// bf16 hidden H_t / pre-activation gradient dP_t slices written by feddat_dat_bwd_dgrad:
//
//   dWu [768, r_t] = scale * dY^T H_t        dbu [768] = scale * sum_m dY
//   dWd [r_t, 768] = dP_t^T  X               dbd [r_t] = sum_m dP_t
//
// Both products contract over the token dimension, so every operand is MN-major for tcgen05:
// the TMA box is [64 rows x 64 columns] of the row-major activation and the UMMA descriptors walk
// K = rows (128 B apart), MN = columns (contiguous), 64-column blocks `LBO` apart.
//
// Grid = groups x 6 column chunks (128 of the 768 model columns) x row splits, at most one CTA per SM.
// A CTA accumulates
//   D1 [r_t(<=128 lanes) x 128] = H_t^T  dY[:, chunk]     (= dWu^T chunk)
//   D2 [r_t          x 128]    = dP_t^T X [:, chunk]     (= dWd chunk)
// in TMEM over all its 64-row blocks.  The bias gradients are column sums taken from the same smem
// stages by the otherwise idle epilogue warps.
//
// DETERMINISTIC two-stage reduction across the row splits (round 1 used fp32 atomics: run-to-run
// different summation order): every CTA stores its partial tiles to a workspace (coalesced float4
// layout), the S CTAs of one (group, chunk) meet at a counter in global memory -- all CTAs of the launch
// are co-resident: the grid never exceeds the SM count and a CTA takes a whole SM -- and then each sums
// 1/S of the tile over the S partials IN SPLIT ORDER and writes the final gradient (no atomics, no
// zero-initialised outputs).  Up to 24 groups per launch: the gating rows and the adapter_1 rows of one site
// (dat_fused.cu), or -- the train step's DEFERRED form -- both groups of ALL twelve sites at the end of the
// backward pass: 24 groups x 6 chunks = 144 CTAs, every CTA contracts over ALL rows of its (group, chunk), so
// there are no row splits, no partials and no second stage (n_splits == 1 skips the workspace altogether).
#include <stdlib.h>

#include "feddat_b200.h"
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {
namespace {

constexpr int kD = 768;
constexpr int KB = 64;                   // rows (GEMM K) per pipeline stage
constexpr int NCW = 128;                 // model columns per CTA
constexpr int NCHUNK = kD / NCW;         // 6
constexpr int BLK = KB * 128;            // one [64 rows x 64 cols] bf16 block = 8 KB
constexpr int OPER = 2 * BLK;            // one operand (two 64-column blocks) = 16 KB
constexpr int STAGE = 4 * OPER;          // H, dP, dY, X
constexpr int WG_STAGES = 3;
constexpr int NUM_THREADS = 192;

constexpr int MAX_GROUPS = 24;
constexpr int WS_HEADER_BYTES = 4096;    // counters: [MAX_GROUPS][NCHUNK][2] u32 = 1152 B
constexpr int TILE_F4 = 2 * 128 * NCW / 4;   // float4s of one CTA's two partial tiles (D1 | D2) = 8192
constexpr int PART_FLOATS = 2 * 128 * NCW + 2 * NCW;   // + the two bias-gradient partial vectors

struct WgradGroup {
  int M, rt, n_splits, n_rowblocks, a_blocks, a_3d;
  int ld_dwu;
  int first_cta;       // first CTA of this group in the launch (NCHUNK * n_splits CTAs)
  float scale;
  float* dWu;
  float* dbu;
  float* dWd;
  float* dbd;
};

struct WgradTmaps {
  CUtensorMap xk, dyk, h, dp;   // h / dp: the 3-D k-block view when a_3d, the plain 2-D view otherwise
};

struct WgradParams {
  WgradTmaps tm[MAX_GROUPS];    // 12 KB of tensor maps: kernel parameters may take 32 764 B since CUDA 12.1
  WgradGroup g[MAX_GROUPS];
  int n_groups;
  unsigned int* counters;   // workspace head: [group][chunk][2] = {partials stored, final slices written}
  float* partials;          // workspace: [CTA][PART_FLOATS]
  unsigned long long* trace;   // debug timeline of CTA 0 (events 200..) or null
};

// spin until *ctr >= target (all CTAs of the launch are co-resident; bounded like mbar_wait)
__device__ __forceinline__ void wait_counter(const unsigned int* ctr, unsigned int target) {
  const uint64_t t0 = globaltimer_ns();
  unsigned int v, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) return;
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > FD_MBAR_TIMEOUT_NS) {
      printf("[feddat] wgrad split barrier timeout: block %d counter %u / %u\n", blockIdx.x, v, target);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_wgrad_kernel(const __grid_constant__ WgradParams pp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red_smem[2][8][NCW];     // bias-gradient partials of the 8 row sets (8 KB)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef FEDDAT_DEBUG
#define WG_TRACE(ev)                                                              \
  do {                                                                            \
    if (pp.trace != nullptr && blockIdx.x == 0) pp.trace[(ev)] = globaltimer_ns();  \
  } while (0)
#else
#define WG_TRACE(ev) do { } while (0)
#endif
  int grp = 0;
  for (int i = 1; i < pp.n_groups; ++i)
    if (static_cast<int>(blockIdx.x) >= pp.g[i].first_cta) grp = i;
  const WgradGroup& p = pp.g[grp];
  const WgradTmaps& T = pp.tm[grp];
  const CUtensorMap &tmXk = T.xk, &tmDYk = T.dyk, &tmH = T.h, &tmDP = T.dp;
  if (tid == 0) WG_TRACE(200);
  const int rel = static_cast<int>(blockIdx.x) - p.first_cta;
  const int chunk = rel % NCHUNK, split = rel / NCHUNK;
  const int col0 = chunk * NCW;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (WG_STAGES + s); };
  const uint32_t bar_acc = bar0 + 8u * (2 * WG_STAGES);
  const uint32_t bar_red = bar0 + 8u * (2 * WG_STAGES + 1);   // stage-2 reduction: partial slices landed in smem

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1 + 4);  // MMA commit + one arrive per aux warp
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_red, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmXk);
    tma_prefetch_desc(&tmDYk);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmDP);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                         // the prologue above overlapped the previous kernel's tail (PDL)
  pdl_launch_dependents();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) WG_TRACE(201);       // prologue done
  const int ablk = p.a_blocks;  // 1 when r_t <= 64 (second 64-column block of H/dP never loaded)

  if (warp == 0) {
    // four producer lanes, one operand each (H_t, dP_t, dY, X): a lone thread sustains one TMA per
    // ~170 ns (scripts/ingest_probe.py), and a stage used to be eight 8 KB boxes issued by one thread.
    // dY / X (and H_t / dP_t when r_t % 64 == 0) arrive as ONE 3-D box covering both 64-column blocks.
    if (lane < 4) {
      int stage = 0, it_n = 0;
      uint32_t phase = 0;
      for (int rb = split; rb < p.n_rowblocks; rb += p.n_splits) {
        const int m0 = rb * KB;
        mbar_wait(bar_empty(stage), phase ^ 1);
        const uint32_t dst = smem0 + stage * STAGE + lane * OPER;
        // lanes 1-3 may complete bytes before lane 0's expect_tx: the phase cannot close before lane
        // 0's arrival, and a transiently negative tx-count is legal
        if (lane == 0) mbar_arrive_expect_tx(bar_full(stage), 2 * ablk * BLK + 2 * OPER);
        if (lane < 2) {
          if (p.a_3d) {
            tma_load_3d(dst, lane == 0 ? &tmH : &tmDP, bar_full(stage), 0, m0, 0);
          } else {
            for (int b = 0; b < ablk; ++b)
              tma_load_2d(dst + b * BLK, lane == 0 ? &tmH : &tmDP, bar_full(stage), b * 64, m0);
          }
        } else {
          tma_load_3d(dst, lane == 2 ? &tmDYk : &tmXk, bar_full(stage), 0, m0, col0 / 64);
        }
        if (lane == 0 && it_n < 10) WG_TRACE(210 + it_n);
        ++it_n;
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = make_idesc_bf16(128, NCW, 1, 1);
      const uint32_t a_lbo = ablk == 2 ? BLK : 0;  // r_t <= 64: lanes 64..127 alias block 0 (ignored)
      bool first = true;
      int it_m = 0;
      for (int rb = split; rb < p.n_rowblocks; rb += p.n_splits) {
        mbar_wait(bar_full(stage), phase);
        tc_fence_after();
        const uint32_t base = smem0 + stage * STAGE;
#pragma unroll
        for (int kk = 0; kk < KB / 16; ++kk) {
          const uint32_t ko = kk * 16 * 128;
          umma_ss(tmem, desc_mnmajor_sw128(base + ko, a_lbo),
                  desc_mnmajor_sw128(base + 2 * OPER + ko, BLK), idesc, !(first && kk == 0));
          umma_ss(tmem + NCW, desc_mnmajor_sw128(base + OPER + ko, a_lbo),
                  desc_mnmajor_sw128(base + 3 * OPER + ko, BLK), idesc, !(first && kk == 0));
        }
        first = false;
        umma_commit(bar_empty(stage));
        if (it_m < 10) WG_TRACE(220 + it_m);
        ++it_m;
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar_acc);
      WG_TRACE(202);                  // all MMAs issued
    }
    __syncwarp();
  } else {
    // aux: bias-gradient column sums from the staged tiles; then the reduction epilogue
    const uint32_t q = warp & 3;
    const uint32_t j = q * 32 + lane;  // column within the chunk == TMEM lane
    // Bias gradients = column sums of the staged dY / dP tiles.  Thread t of the 128 aux threads owns the
    // 16-byte column group g = t % 16 (8 columns) of the rows r == t / 16 (mod 8): eight ld.shared.v4 per
    // tile and stage instead of 64 two-byte loads (the scalar version took 0.9 us per stage and, since a
    // stage is only released when these warps are done with it, set the kernel's pace: 1.2 us per stage
    // against 0.26 us of MMA time).  Partial sums stay in registers until the end of the kernel.
    const uint32_t t = tid - 64;
    const uint32_t g = t & 15, rs = t >> 4;
    const uint32_t goff = (g >> 3) * BLK + rs * 128 + (((g & 7) ^ rs) << 4);   // (rs + 8 i) & 7 == rs
    const bool do_dbu = p.dbu != nullptr;
    const bool do_dbd = (chunk == 0) && p.dbd != nullptr && static_cast<int>(g * 8) < p.rt;
    float acc_dy[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float acc_dp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto accumulate = [&](uint32_t tile_base, float (&acc)[8]) {
#pragma unroll
      for (int i = 0; i < KB / 8; ++i) {
        const uint4 v = ld_shared_v4(tile_base + goff + i * 1024);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] += __uint_as_float(w[k] << 16);
          acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
        }
      }
    };
    int stage = 0, it_a = 0;
    uint32_t phase = 0;
    for (int rb = split; rb < p.n_rowblocks; rb += p.n_splits) {
      mbar_wait(bar_full(stage), phase);
      const uint32_t base = smem0 + stage * STAGE;
      if (do_dbu) accumulate(base + 2 * OPER, acc_dy);
      if (do_dbd) accumulate(base + OPER, acc_dp);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(stage));
      if (tid == 64 && it_a < 10) WG_TRACE(230 + it_a);
      ++it_a;
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- stage 1 of the reduction: this CTA's partials -> workspace (not needed without row splits)
    const int S = p.n_splits;
    const bool cta_dbd = (chunk == 0) && p.dbd != nullptr;
    float* part = pp.partials + static_cast<size_t>(blockIdx.x) * PART_FLOATS;
    {
      // eight row-set partials per column -> one sum per column and CTA through smem
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        red_smem[0][rs][g * 8 + k] = acc_dy[k];
        red_smem[1][rs][g * 8 + k] = acc_dp[k];
      }
      named_bar_sync(1, 128);
      float sdy = 0.f, sdp = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        sdy += red_smem[0][r][t];
        sdp += red_smem[1][r][t];
      }
      if (S == 1) {
        if (do_dbu) p.dbu[col0 + t] = p.scale * sdy;
        if (cta_dbd && static_cast<int>(t) < p.rt) p.dbd[t] = sdp;
      } else {
        part[2 * 128 * NCW + t] = sdy;
        part[2 * 128 * NCW + NCW + t] = sdp;
      }
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    if (tid == 64) WG_TRACE(203);   // accumulators complete
    // The pipeline stages are free now (every MMA that read them has completed): they stage the partial
    // tiles.  thread j = TMEM lane = bottleneck unit j; float4 i of the 32-column group c of matrix m goes
    // to float4 index ((m * 4 + c) * 8 + i) * 128 + j (conflict-free st.shared.v4), then ONE 128 KB bulk
    // copy moves the CTA's partials to the workspace (row-per-thread st.global took 4 us here).
    {
      const uint32_t lane_addr = (q * 32) << 16;
#pragma unroll 1
      for (int mc = 0; mc < 2 * (NCW / 32); ++mc) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + (mc >> 2) * NCW + (mc & 3) * 32, v);
        tmem_ld_wait32(v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st_shared_v4(smem0 + (((mc * 8 + i) * 128 + j) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
    unsigned int* ctr = pp.counters + (grp * NCHUNK + chunk) * 2;
    // ---- stage 2: wait for the partials of all row splits of this (group, chunk), then sum 1 / S of the
    // tile over the splits in split order (deterministic) and write the final gradients.  S == 1: the tile
    // staged in shared memory IS the sum -- same layout, read in place.
    const int per = (TILE_F4 + S - 1) / S;
    const int f_begin = split * per;
    const int f_end = (split + 1) * per < TILE_F4 ? (split + 1) * per : TILE_F4;
    const int n_f = f_end > f_begin ? f_end - f_begin : 0;
    if (S == 1) {
      named_bar_sync(1, 128);
    } else {
      const float* part0 = pp.partials + static_cast<size_t>(p.first_cta + chunk) * PART_FLOATS;   // split 0
      const size_t split_stride = static_cast<size_t>(NCHUNK) * PART_FLOATS;
      fence_proxy_async_smem();
      __threadfence();                 // the two bias partial vectors (plain stores above)
      named_bar_sync(1, 128);
      if (t == 0) {
        bulk_store_1d(part, smem0, TILE_F4 * 16);
        tma_store_commit();
        tma_store_wait_all<0>();       // the partials are written (not merely read out of smem)
        fence_proxy_async_all();
        __threadfence();
        atomicAdd(ctr, 1u);
        wait_counter(ctr, static_cast<unsigned int>(S));
        __threadfence();
        fence_proxy_async_all();
        if (tid == 64) WG_TRACE(204);
        // this CTA's slice [f_begin, f_end) of every split's partial tile -> smem, one bulk copy per split
        if (n_f > 0) {
          mbar_arrive_expect_tx(bar_red, static_cast<uint32_t>(S) * n_f * 16u);
          for (int sidx = 0; sidx < S; ++sidx)
            bulk_load_1d(smem0 + static_cast<uint32_t>(sidx) * per * 16u,
                         part0 + sidx * split_stride + static_cast<size_t>(f_begin) * 4, n_f * 16u, bar_red);
        } else {
          mbar_arrive(bar_red);
        }
      }
      if (split == 0) {   // bias gradients: one CTA per (group, chunk), fixed order; loads issued together
        named_bar_sync(2, 128);        // thread 0 has passed the split barrier
        __threadfence();
        float vy[24], vp[24];
#pragma unroll
        for (int sidx = 0; sidx < 24; ++sidx) {
          vy[sidx] = vp[sidx] = 0.f;
          if (sidx < S) {
            const float* ps = part0 + sidx * split_stride + 2 * 128 * NCW;
            vy[sidx] = __ldcg(ps + t);
            vp[sidx] = __ldcg(ps + NCW + t);
          }
        }
        float sdy = 0.f, sdp = 0.f;
#pragma unroll
        for (int sidx = 0; sidx < 24; ++sidx) {
          sdy += vy[sidx];
          sdp += vp[sidx];
        }
        if (do_dbu) p.dbu[col0 + t] = p.scale * sdy;
        if (cta_dbd && static_cast<int>(t) < p.rt) p.dbd[t] = sdp;
      }
      mbar_wait(bar_red, 0);
    }
    for (int k = static_cast<int>(t); k < n_f; k += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int sidx = 0; sidx < S; ++sidx) {
        const uint4 u = ld_shared_v4(smem0 + (static_cast<uint32_t>(sidx) * per + k) * 16u);
        acc.x += __uint_as_float(u.x); acc.y += __uint_as_float(u.y);
        acc.z += __uint_as_float(u.z); acc.w += __uint_as_float(u.w);
      }
      const int f = f_begin + k;
      const int jj = f & 127, i = (f >> 7) & 7, c = (f >> 10) & 3, m = f >> 12;
      if (jj < p.rt) {
        const int col = col0 + c * 32 + i * 4;
        if (m == 0) {   // D1 = dWu^T chunk: element (jj, col) -> dWu[col, jj]
          float* d = p.dWu + static_cast<size_t>(col) * p.ld_dwu + jj;
          d[0] = p.scale * acc.x;
          d[p.ld_dwu] = p.scale * acc.y;
          d[2 * static_cast<size_t>(p.ld_dwu)] = p.scale * acc.z;
          d[3 * static_cast<size_t>(p.ld_dwu)] = p.scale * acc.w;
        } else {        // D2 = dWd chunk
          *reinterpret_cast<float4*>(p.dWd + static_cast<size_t>(jj) * kD + col) = acc;
        }
      }
    }
    if (S > 1) {
      // the last CTA of this (group, chunk) to finish resets the counters for the next launch
      named_bar_sync(1, 128);
      if (t == 0) {
        const unsigned int done = atomicAdd(ctr + 1, 1u);
        if (done == static_cast<unsigned int>(S) - 1) {
          ctr[0] = 0u;
          ctr[1] = 0u;
          __threadfence();
        }
      }
    }
  }
  if (tid == 64) WG_TRACE(206);       // reductions done
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
  if (tid == 0) WG_TRACE(205);
}

}  // namespace
}  // namespace fd

namespace fd {
namespace {

size_t wgrad_ws_bytes(int sms) {
  return WS_HEADER_BYTES + static_cast<size_t>(sms) * PART_FLOATS * sizeof(float);
}

int check_group(const FeddatWgradGroup& G, int d, int dtype) {
  FD_REQUIRE(G.X && G.dY && G.H_t && G.dP_t && G.dWu && G.dWd, FD_ERR_INVALID,
             "dat_bwd_wgrad: null pointer argument");
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED,
             "dat_bwd_wgrad: only bf16 activations are implemented (dtype=%d)", dtype);
  FD_REQUIRE(d == kD, FD_ERR_UNSUPPORTED, "dat_bwd_wgrad: model_dim must be 768 (got %d)", d);
  FD_REQUIRE(G.r_t >= 16 && G.r_t <= 128 && G.r_t % 16 == 0, FD_ERR_UNSUPPORTED,
             "dat_bwd_wgrad: r_t must be a multiple of 16 in [16, 128] (got %d); call once per "
             "128-wide slice for wider bottlenecks", G.r_t);
  FD_REQUIRE(G.ld_ht >= G.r_t && G.ld_ht % 8 == 0 && G.ld_dwu >= G.r_t, FD_ERR_INVALID,
             "dat_bwd_wgrad: bad leading dimensions ld_ht=%d ld_dwu=%d", G.ld_ht, G.ld_dwu);
  FD_REQUIRE((reinterpret_cast<uintptr_t>(G.dWd) & 15) == 0, FD_ERR_INVALID,
             "dat_bwd_wgrad: dWd must be 16-byte aligned");
  FD_REQUIRE(G.M >= 1 && G.M < (1ll << 31) - 256, FD_ERR_INVALID, "dat_bwd_wgrad: bad row count %lld",
             (long long)G.M);
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" size_t feddat_dat_wgrad_workspace_bytes(void) {
  int sms = 0;
  if (fd::device_sm_count(&sms)) return 0;
  return fd::wgrad_ws_bytes(sms);
}

extern "C" int feddat_dat_bwd_wgrad_grouped(const FeddatWgradGroup* groups, int n_groups, int d, int dtype,
                                            void* workspace, size_t ws_bytes, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(groups != nullptr && n_groups >= 1 && n_groups <= MAX_GROUPS, FD_ERR_INVALID,
             "dat_bwd_wgrad_grouped: 1 to %d groups per launch (got %d)", MAX_GROUPS, n_groups);
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  FD_REQUIRE(workspace != nullptr && ws_bytes >= wgrad_ws_bytes(sms) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
             FD_ERR_INVALID, "dat_bwd_wgrad: workspace of feddat_dat_wgrad_workspace_bytes() = %zu bytes needed "
             "(16-byte aligned, zero-initialised once)", wgrad_ws_bytes(sms));

  static thread_local WgradParams p;    // 14 KB: kept off the stack of the (autograd) caller thread
  p = WgradParams{};
  WgradTmaps* tms = p.tm;
  p.n_groups = n_groups;
  p.counters = static_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + WS_HEADER_BYTES);
  p.trace = FD_TRACE_PTR;
  int max_splits = sms / (NCHUNK * n_groups);
  if (max_splits < 1) max_splits = 1;
  int ctas = 0;
  for (int gi = 0; gi < n_groups; ++gi) {
    const FeddatWgradGroup& G = groups[gi];
    if ((rc = check_group(G, d, dtype))) return rc;
    WgradGroup& w = p.g[gi];
    w.M = static_cast<int>(G.M);
    w.rt = G.r_t;
    w.n_rowblocks = static_cast<int>((G.M + KB - 1) / KB);
    w.a_blocks = G.r_t > 64 ? 2 : 1;
    w.ld_dwu = G.ld_dwu;
    w.scale = G.branch_scale;
    w.dWu = G.dWu; w.dbu = G.dbu; w.dWd = G.dWd; w.dbd = G.dbd;
    w.n_splits = max_splits < w.n_rowblocks ? max_splits : w.n_rowblocks;
    w.first_cta = ctas;
    ctas += NCHUNK * w.n_splits;
    w.a_3d = (G.r_t % 64 == 0) ? 1 : 0;
    WgradTmaps& T = tms[gi];
    if ((rc = make_tmap_bf16_kblocks(&T.xk, G.X, G.M, kD, kD, KB, 2))) return rc;
    if ((rc = make_tmap_bf16_kblocks(&T.dyk, G.dY, G.M, kD, kD, KB, 2))) return rc;
    if (w.a_3d) {
      if ((rc = make_tmap_bf16_kblocks(&T.h, G.H_t, G.M, G.r_t, G.ld_ht, KB, G.r_t / 64))) return rc;
      if ((rc = make_tmap_bf16_kblocks(&T.dp, G.dP_t, G.M, G.r_t, G.ld_ht, KB, G.r_t / 64))) return rc;
    } else {
      if ((rc = make_tmap_bf16_2d(&T.h, G.H_t, G.M, G.r_t, G.ld_ht, KB, 64))) return rc;
      if ((rc = make_tmap_bf16_2d(&T.dp, G.dP_t, G.M, G.r_t, G.ld_ht, KB, 64))) return rc;
    }
  }
  FD_REQUIRE(ctas <= sms, FD_ERR_UNSUPPORTED, "dat_bwd_wgrad: %d CTAs exceed the %d SMs (co-residency)", ctas, sms);

  const size_t smem = 1024 + static_cast<size_t>(WG_STAGES) * STAGE;
  static bool configured[64] = {false};   // idempotent "attribute already set" cache
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(dat_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    if (dev < 64) configured[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const char* e = getenv("FEDDAT_PDL");
  cfg.numAttrs = (e && e[0] == '0') ? 0 : 1;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dat_wgrad_kernel, p));
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

extern "C" int feddat_dat_bwd_wgrad(const void* X, const void* dY, const void* H_t,
                                    const void* dP_t, float* dWu, float* dbu, float* dWd, float* dbd,
                                    int64_t M, int d, int r_t, int ld_ht, int ld_dwu,
                                    float branch_scale, int dtype, void* workspace, size_t ws_bytes,
                                    void* stream) {
  if (M == 0) return FEDDAT_OK;
  FeddatWgradGroup G{};
  G.X = X; G.dY = dY; G.H_t = H_t; G.dP_t = dP_t; G.dWu = dWu; G.dbu = dbu; G.dWd = dWd; G.dbd = dbd;
  G.M = M; G.r_t = r_t; G.ld_ht = ld_ht; G.ld_dwu = ld_dwu; G.branch_scale = branch_scale;
  return feddat_dat_bwd_wgrad_grouped(&G, 1, d, dtype, workspace, ws_bytes, stream);
}
