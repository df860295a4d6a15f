// DAT bottleneck forward (reference: src/modeling/models/adapter.py:124-163).
//
//   Y = Res + scale * ( act(X * Wd_cat^T + bd_cat) * Wu_cat^T + bu_cat )
//
// "cat" = the active branches concatenated along the bottleneck dimension: one branch
// (adapter.py:125-131, scale 1) or the two gating branches adapter_0 | adapter_2
// (adapter.py:133-146, scale 0.5 each), so the dual adapter is ONE GEMM pair with hidden width
// R = r_total.  One persistent CTA per SM walks 128-row tiles:
//
//   warp 0      TMA producer   X k-chunks + Wd_cat k-chunks, then Wu_cat (n-chunk, k-chunk) tiles
//   warp 1      tcgen05.mma issuer: GEMM1 (128 x R x 768) -> TMEM, GEMM2 (128 x 768 x R) in N2-wide
//               chunks -> TMEM, accumulators in a 2 x 256-column ring
//   warps 2-5   epilogue: (1) TMEM -> +bias, act -> bf16 hidden tile in swizzled smem (the A operand
//               of GEMM2, it never goes to HBM); (2) TMEM -> scale, +bias, +residual (residual tile
//               brought in by TMA) -> bf16 -> smem -> TMA store
//
// HBM traffic per row: read X (1536 B) + write Y (1536 B); the residual re-read hits L2.
#include "feddat_b200.h"
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {

namespace {

constexpr int kD = 768;
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int KC1 = kD / BK;  // 12 k-chunks for GEMM1
constexpr int A_SLOT = BM * 128;
constexpr int STG_BYTES = BM * 128;  // 128 rows x 64 bf16
constexpr int NSTG = 3;
constexpr int MAX_STAGES = 4;
constexpr int NUM_THREADS = 192;
constexpr int CHUNKS_PER_TILE = kD / 64;  // 12 output chunks of 64 columns

struct FwdParams {
  int M, R, num_tiles, stages, n2, act;
  uint32_t b_slot_bytes;
  float scale;
  const float* bd;
  const float* bu;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 0) return fmaxf(x, 0.f);
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmRes,
               const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWd,
               const __grid_constant__ CUtensorMap tmWu, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 2 + 2 + 2 + NSTG];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R, S = p.stages, N2 = p.n2;
  const int KC2 = (R + 63) / 64;
  const int NC2 = kD / N2;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = A_SLOT + p.b_slot_bytes;
  const uint32_t h_base = smem0 + S * stage_bytes;
  const uint32_t stg_base = h_base + KC2 * A_SLOT;
  const uint32_t bias_base = stg_base + NSTG * STG_BYTES;
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));

  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto bar_acc_full = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto bar_acc_empty = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 2 + b); };
  const uint32_t bar_h_full = bar0 + 8u * (2 * MAX_STAGES + 4);
  const uint32_t bar_h_empty = bar0 + 8u * (2 * MAX_STAGES + 5);
  auto bar_res_full = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 6 + b); };

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 128);
    }
    mbar_init(bar_h_full, 128);
    mbar_init(bar_h_empty, 1);
    for (int b = 0; b < NSTG; ++b) mbar_init(bar_res_full(b), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmRes);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmWd);
    tma_prefetch_desc(&tmWu);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  // biases -> smem (broadcast reads in the epilogues)
  for (int i = tid; i < R; i += NUM_THREADS) bias_smem[i] = p.bd[i];
  for (int i = tid; i < kD; i += NUM_THREADS) bias_smem[R + i] = p.bu[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = tile * BM;
        for (int kc = 0; kc < KC1; ++kc) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          const uint32_t a_dst = smem0 + stage * stage_bytes;
          mbar_arrive_expect_tx(bar_full(stage), A_SLOT + R * 128);
          tma_load_2d_hint(a_dst, &tmX, bar_full(stage), kc * BK, m0, kEvictNormal);
          tma_load_2d_hint(a_dst + A_SLOT, &tmWd, bar_full(stage), kc * BK, 0, kEvictLast);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        for (int nc = 0; nc < NC2; ++nc) {
          for (int kc = 0; kc < KC2; ++kc) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            const uint32_t b_dst = smem0 + stage * stage_bytes + A_SLOT;
            mbar_arrive_expect_tx(bar_full(stage), N2 * 128);
            tma_load_2d_hint(b_dst, &tmWu, bar_full(stage), kc * BK, nc * N2, kEvictLast);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_it = 0, tile_it = 0;
      const uint32_t idesc1 = make_idesc_bf16(BM, R);
      const uint32_t idesc2 = make_idesc_bf16(BM, N2);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_it) {
        {  // GEMM1: P = X * Wd_cat^T
          const uint32_t buf = acc_it & 1, par = (acc_it >> 1) & 1;
          mbar_wait(bar_acc_empty(buf), par ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem + buf * 256;
          for (int kc = 0; kc < KC1; ++kc) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            const uint32_t a_src = smem0 + stage * stage_bytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss(d_tmem, desc_kmajor_sw128(a_src + k * 32),
                      desc_kmajor_sw128(a_src + A_SLOT + k * 32), idesc1, (kc | k) != 0);
            umma_commit(bar_empty(stage));
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_acc_full(buf));
          ++acc_it;
        }
        mbar_wait(bar_h_full, tile_it & 1);
        tc_fence_after();
        for (int nc = 0; nc < NC2; ++nc) {  // GEMM2: Y[:, nc] = H * Wu_cat[nc]^T
          const uint32_t buf = acc_it & 1, par = (acc_it >> 1) & 1;
          mbar_wait(bar_acc_empty(buf), par ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem + buf * 256;
          for (int kc = 0; kc < KC2; ++kc) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            const uint32_t b_src = smem0 + stage * stage_bytes + A_SLOT;
            const int ksteps = min(4, (R - kc * 64) / 16);
            for (int k = 0; k < ksteps; ++k)
              umma_ss(d_tmem, desc_kmajor_sw128(h_base + kc * A_SLOT + k * 32),
                      desc_kmajor_sw128(b_src + k * 32), idesc2, (kc | k) != 0);
            umma_commit(bar_empty(stage));
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_acc_full(buf));
          ++acc_it;
        }
        umma_commit(bar_h_empty);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const uint32_t q = warp & 3;            // TMEM lane quarter this warp may touch
    const uint32_t row = q * 32 + lane;     // tile row == TMEM lane
    const uint32_t lane_addr = (q * 32) << 16;
    const bool leader = (warp == 2 && lane == 0);
    const float scale = p.scale;
    const int act = p.act;
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + gridDim.x - 1) / gridDim.x;
    const uint32_t total_chunks = static_cast<uint32_t>(my_tiles) * CHUNKS_PER_TILE;

    auto issue_res_load = [&](uint32_t g) {
      if (g >= total_chunks) return;
      const int tile = blockIdx.x + (g / CHUNKS_PER_TILE) * gridDim.x;
      const int c = g % CHUNKS_PER_TILE;
      const uint32_t sb = g % NSTG;
      mbar_arrive_expect_tx(bar_res_full(sb), STG_BYTES);
      tma_load_2d(stg_base + sb * STG_BYTES, &tmRes, bar_res_full(sb), c * 64, tile * BM);
    };
    if (leader)
      for (uint32_t g = 0; g < NSTG - 1; ++g) issue_res_load(g);

    uint32_t acc_it = 0, tile_it = 0, chunk_g = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_it) {
      const int m0 = tile * BM;
      {  // epilogue 1: hidden tile
        const uint32_t buf = acc_it & 1, par = (acc_it >> 1) & 1;
        mbar_wait(bar_acc_full(buf), par);
        mbar_wait(bar_h_empty, (tile_it & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_src = tmem + lane_addr + buf * 256;
        for (int c = 0; c < R / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(t_src + c * 16, v);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a = apply_act(__uint_as_float(v[2 * i]) + bias_smem[c * 16 + 2 * i], act);
            float b = apply_act(__uint_as_float(v[2 * i + 1]) + bias_smem[c * 16 + 2 * i + 1], act);
            w[i] = pack_bf16x2(a, b);
          }
          const uint32_t kc = c >> 2, j0 = (c & 3) * 2;
          st_shared_v4(h_base + kc * A_SLOT + sw128_offset(row, j0), w[0], w[1], w[2], w[3]);
          st_shared_v4(h_base + kc * A_SLOT + sw128_offset(row, j0 + 1), w[4], w[5], w[6], w[7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar_acc_empty(buf));
        mbar_arrive(bar_h_full);
        ++acc_it;
      }
      for (int nc = 0; nc < NC2; ++nc) {  // epilogue 2: output chunks
        const uint32_t buf = acc_it & 1, par = (acc_it >> 1) & 1;
        mbar_wait(bar_acc_full(buf), par);
        tc_fence_after();
        for (int j = 0; j < N2 / 64; ++j, ++chunk_g) {
          const uint32_t sb = chunk_g % NSTG, rpar = (chunk_g / NSTG) & 1;
          const int col0 = nc * N2 + j * 64;
          const uint32_t t_src = tmem + lane_addr + buf * 256 + j * 64;
          uint32_t v0[32], v1[32];
          tmem_ld32(t_src, v0);
          tmem_ld32(t_src + 32, v1);
          mbar_wait(bar_res_full(sb), rpar);
          tmem_ld_wait();
          const uint32_t sbuf = stg_base + sb * STG_BYTES;
          const float* bu = bias_smem + R + col0;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint32_t addr = sbuf + sw128_offset(row, c8);
            const uint4 rv = ld_shared_v4(addr);
            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = c8 * 8 + 2 * i;
              const float a0 = __uint_as_float(e < 32 ? v0[e & 31] : v1[e & 31]);
              const float a1 = __uint_as_float(e < 32 ? v0[(e + 1) & 31] : v1[(e + 1) & 31]);
              const float2 r2 = unpack_bf16x2(rr[i]);
              o[i] = pack_bf16x2(r2.x + scale * (a0 + bu[e]), r2.y + scale * (a1 + bu[e + 1]));
            }
            st_shared_v4(addr, o[0], o[1], o[2], o[3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (leader) {
            tma_store_2d(&tmY, sbuf, col0, m0);
            tma_store_commit();
            tma_store_wait_read<1>();
            issue_res_load(chunk_g + NSTG - 1);
          }
        }
        tc_fence_before();
        mbar_arrive(bar_acc_empty(buf));
        ++acc_it;
      }
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

}  // namespace fd

extern "C" int feddat_dat_fwd(const void* X, const void* Res, void* Y, const void* Wd_cat,
                              const float* bd_cat, const void* Wu_cat, const float* bu_cat,
                              int64_t M, int d, int r_total, float branch_scale, int act, int dtype,
                              void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(X && Res && Y && Wd_cat && bd_cat && Wu_cat && bu_cat, FD_ERR_INVALID,
             "dat_fwd: null pointer argument");
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED,
             "dat_fwd: only bf16 activations are implemented (dtype=%d)", dtype);
  FD_REQUIRE(d == kD, FD_ERR_UNSUPPORTED, "dat_fwd: model_dim must be 768 (got %d)", d);
  FD_REQUIRE(r_total >= 16 && r_total <= 256 && r_total % 16 == 0, FD_ERR_UNSUPPORTED,
             "dat_fwd: r_total must be a multiple of 16 in [16, 256] (got %d); split wider "
             "bottlenecks into several calls", r_total);
  FD_REQUIRE(act == FEDDAT_ACT_RELU || act == FEDDAT_ACT_GELU, FD_ERR_INVALID,
             "dat_fwd: unknown activation %d", act);
  FD_REQUIRE(M >= 0 && M < (1ll << 31) - 256, FD_ERR_INVALID, "dat_fwd: bad row count %lld",
             (long long)M);
  if (M == 0) return FD_OK;

  FwdParams p{};
  p.M = static_cast<int>(M);
  p.R = r_total;
  p.num_tiles = static_cast<int>((M + BM - 1) / BM);
  p.n2 = r_total > 128 ? 256 : 128;
  p.b_slot_bytes = static_cast<uint32_t>(p.n2 * 128);
  p.act = act;
  p.scale = branch_scale;
  p.bd = bd_cat;
  p.bu = bu_cat;
  const int kc2 = (r_total + 63) / 64;
  const size_t fixed = 1024 + static_cast<size_t>(kc2) * A_SLOT + NSTG * STG_BYTES +
                       (r_total + kD) * sizeof(float);
  const size_t max_smem = 227 * 1024 - 512;  // static smem (barriers) lives in the same budget
  const size_t stage_bytes = A_SLOT + p.b_slot_bytes;
  int stages = static_cast<int>((max_smem - fixed) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  FD_REQUIRE(stages >= 2, FD_ERR_UNSUPPORTED, "dat_fwd: shared-memory budget exceeded (R=%d)",
             r_total);
  p.stages = stages;
  const size_t smem = fixed + stages * stage_bytes;

  CUtensorMap tmX, tmRes, tmY, tmWd, tmWu;
  if ((rc = make_tmap_bf16_2d(&tmX, X, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmRes, Res, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmY, Y, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWd, Wd_cat, r_total, kD, kD, r_total, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWu, Wu_cat, kD, r_total, r_total, p.n2, 64))) return rc;

  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  static bool configured[64] = {false};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(dat_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)max_smem));
    if (dev < 64) configured[dev] = true;
  }
  dat_fwd_kernel<<<grid, NUM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmRes, tmY,
                                                                                 tmWd, tmWu, p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
