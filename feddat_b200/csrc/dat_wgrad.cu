// DAT bottleneck backward, weight gradients of the trainable branch (autograd of reference
// adapter.py:124-163; trainability per adapter.py:71-85).  Inputs are the activations X, dY and the
// bf16 hidden H_t / pre-activation gradient dP_t slices written by feddat_dat_bwd_dgrad:
//
//   dWu [768, r_t] += scale * dY^T H_t        dbu [768] += scale * sum_m dY
//   dWd [r_t, 768] += dP_t^T  X               dbd [r_t] += sum_m dP_t
//
// Both products contract over the token dimension, so every operand is MN-major for tcgen05:
// the TMA box is [64 rows x 64 columns] of the row-major activation and the UMMA descriptors walk
// K = rows (128 B apart), MN = columns (contiguous), 64-column blocks `LBO` apart.
//
// Grid = 6 column chunks (128 of the 768 model columns) x row splits.  A CTA accumulates
//   D1 [r_t(<=128 lanes) x 128] = H_t^T  dY[:, chunk]     (= dWu^T chunk)
//   D2 [r_t          x 128]    = dP_t^T X [:, chunk]     (= dWd chunk)
// in TMEM over all its 64-row blocks, and reduces into the fp32 gradients with red.global.add at
// the end (fp32 atomics: summation order across row splits is not deterministic).  The bias
// gradients are column sums taken from the same smem stages by the otherwise idle epilogue warps.
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

struct WgradParams {
  int M, rt, n_splits, n_rowblocks, a_blocks, a_3d;
  int ld_dwu;
  float scale;
  float* dWu;
  float* dbu;
  float* dWd;
  float* dbd;
  unsigned long long* trace;   // debug timeline of CTA 0 (events 200..) or null
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_wgrad_kernel(const __grid_constant__ CUtensorMap tmXk, const __grid_constant__ CUtensorMap tmDYk,
                 const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmDP,
                 const __grid_constant__ CUtensorMap tmHk, const __grid_constant__ CUtensorMap tmDPk,
                 const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 1];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red_smem[2][8][NCW];     // bias-gradient partials of the 8 row sets (8 KB)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef FEDDAT_DEBUG
#define WG_TRACE(ev)                                                              \
  do {                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(ev)] = globaltimer_ns();  \
  } while (0)
#else
#define WG_TRACE(ev) do { } while (0)
#endif
  if (tid == 0) WG_TRACE(200);
  const int chunk = blockIdx.x % NCHUNK, split = blockIdx.x / NCHUNK;
  const int col0 = chunk * NCW;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (WG_STAGES + s); };
  const uint32_t bar_acc = bar0 + 8u * (2 * WG_STAGES);

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1 + 4);  // MMA commit + one arrive per aux warp
    }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmXk);
    tma_prefetch_desc(&tmDYk);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmDP);
    tma_prefetch_desc(&tmHk);
    tma_prefetch_desc(&tmDPk);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
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
            tma_load_3d(dst, lane == 0 ? &tmHk : &tmDPk, bar_full(stage), 0, m0, 0);
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
    // eight row-set partials per column -> one sum per column and CTA through smem (one global atomic
    // per column and CTA, as many-way contended as the weight-gradient reduction: the row splits)
    {
      const bool cta_dbd = (chunk == 0) && p.dbd != nullptr;
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
      if (do_dbu) atomicAdd(p.dbu + col0 + t, p.scale * sdy);
      if (cta_dbd && static_cast<int>(t) < p.rt) atomicAdd(p.dbd + t, sdp);
    }

    if (split < p.n_rowblocks) {  // this CTA accumulated at least one row block
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      if (tid == 64) WG_TRACE(203);   // accumulators complete
      const uint32_t lane_addr = (q * 32) << 16;
      const bool valid = static_cast<int>(j) < p.rt;  // lane j = bottleneck unit j
#pragma unroll 1
      for (int c = 0; c < NCW / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + c * 32, v);  // D1: dWu^T
        tmem_ld_wait32(v);
        if (valid) {
          float* dst = p.dWu + static_cast<size_t>(col0 + c * 32) * p.ld_dwu + j;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            atomicAdd(dst + static_cast<size_t>(i) * p.ld_dwu, p.scale * __uint_as_float(v[i]));
        }
        tmem_ld32(tmem + lane_addr + NCW + c * 32, v);  // D2: dWd
        tmem_ld_wait32(v);
        if (valid) {
          float* dst = p.dWd + static_cast<size_t>(j) * kD + col0 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                         "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                         "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                         : "memory");
        }
      }
    }
  }
  if (tid == 64) WG_TRACE(204);       // reductions issued
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
  if (tid == 0) WG_TRACE(205);
}

}  // namespace
}  // namespace fd

extern "C" int feddat_dat_bwd_wgrad(const void* X, const void* dY, const void* H_t,
                                    const void* dP_t, float* dWu, float* dbu, float* dWd, float* dbd,
                                    int64_t M, int d, int r_t, int ld_ht, int ld_dwu,
                                    float branch_scale, int dtype, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(X && dY && H_t && dP_t && dWu && dWd, FD_ERR_INVALID,
             "dat_bwd_wgrad: null pointer argument");
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED,
             "dat_bwd_wgrad: only bf16 activations are implemented (dtype=%d)", dtype);
  FD_REQUIRE(d == kD, FD_ERR_UNSUPPORTED, "dat_bwd_wgrad: model_dim must be 768 (got %d)", d);
  FD_REQUIRE(r_t >= 16 && r_t <= 128 && r_t % 16 == 0, FD_ERR_UNSUPPORTED,
             "dat_bwd_wgrad: r_t must be a multiple of 16 in [16, 128] (got %d); call once per "
             "128-wide slice for wider bottlenecks", r_t);
  FD_REQUIRE(ld_ht >= r_t && ld_ht % 8 == 0 && ld_dwu >= r_t, FD_ERR_INVALID,
             "dat_bwd_wgrad: bad leading dimensions ld_ht=%d ld_dwu=%d", ld_ht, ld_dwu);
  FD_REQUIRE((reinterpret_cast<uintptr_t>(dWd) & 15) == 0, FD_ERR_INVALID,
             "dat_bwd_wgrad: dWd must be 16-byte aligned");
  FD_REQUIRE(M >= 0 && M < (1ll << 31) - 256, FD_ERR_INVALID, "dat_bwd_wgrad: bad row count %lld",
             (long long)M);
  if (M == 0) return FD_OK;

  WgradParams p{};
  p.M = static_cast<int>(M);
  p.rt = r_t;
  p.n_rowblocks = static_cast<int>((M + KB - 1) / KB);
  p.a_blocks = r_t > 64 ? 2 : 1;
  p.ld_dwu = ld_dwu;
  p.scale = branch_scale;
  p.dWu = dWu; p.dbu = dbu; p.dWd = dWd; p.dbd = dbd;
  p.trace = FD_TRACE_PTR;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  int splits = sms / NCHUNK;
  if (splits > p.n_rowblocks) splits = p.n_rowblocks;
  if (splits < 1) splits = 1;
  p.n_splits = splits;

  p.a_3d = (r_t % 64 == 0) ? 1 : 0;
  CUtensorMap tmXk, tmDYk, tmH, tmDP, tmHk, tmDPk;
  if ((rc = make_tmap_bf16_kblocks(&tmXk, X, M, kD, kD, KB, 2))) return rc;
  if ((rc = make_tmap_bf16_kblocks(&tmDYk, dY, M, kD, kD, KB, 2))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmH, H_t, M, r_t, ld_ht, KB, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDP, dP_t, M, r_t, ld_ht, KB, 64))) return rc;
  tmHk = tmH;
  tmDPk = tmDP;
  if (p.a_3d) {
    if ((rc = make_tmap_bf16_kblocks(&tmHk, H_t, M, r_t, ld_ht, KB, r_t / 64))) return rc;
    if ((rc = make_tmap_bf16_kblocks(&tmDPk, dP_t, M, r_t, ld_ht, KB, r_t / 64))) return rc;
  }

  const size_t smem = 1024 + static_cast<size_t>(WG_STAGES) * STAGE;
  static bool configured[64] = {false};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(dat_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    if (dev < 64) configured[dev] = true;
  }
  dat_wgrad_kernel<<<NCHUNK * splits, NUM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tmXk, tmDYk, tmH, tmDP, tmHk, tmDPk, p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
