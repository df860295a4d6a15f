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
//
// The same kernel template, instantiated with kBwd = true, is the backward data-gradient pass
// (what torch autograd derives from adapter.py:124-163):
//   GEMM1  P  = X  * Wd_cat^T            (recompute, never stored)
//   GEMM1b dH = dY * Wu_cat              (B operand: WuT_cat, K-major)
//   epilogue 1: dP = scale * dH * act'(P + bd) -> bf16 tile in smem (A operand of GEMM3); for the
//               trainable slice [r_lo, r_hi) also H = act(P + bd) and dP to HBM for the wgrad kernel
//   GEMM3  dX = dP * Wd_cat              (B operand: WdT_cat, K-major), + dY when the residual
//               input is X itself (adaptered_output.py:78), -> bf16 -> TMA store
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

struct FusedParams {
  int M, R, num_tiles, stages, n2, act;
  uint32_t b_slot_bytes;
  float scale;
  const float* bd;
  const float* bu;        // fwd only
  // bwd only
  int has_out;            // 0: dX not requested (skip GEMM3 / epilogue 2)
  int has_res;            // add the residual tile (fwd: always; bwd: add_dy)
  int r_lo, r_hi;         // trainable slice of the bottleneck
  __nv_bfloat16* H_t;     // [M, r_hi - r_lo] or null
  __nv_bfloat16* dP_t;    // [M, r_hi - r_lo] or null
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 0) return fmaxf(x, 0.f);
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float act_grad(float x, int act) {
  if (act == 0) return x > 0.f ? 1.f : 0.f;
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) +
         x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// Tensor maps:            forward                      backward (kBwd)
//   tmX   [M, 768]       X                            X
//   tmRes [M, 768]       residual input               dY  (GEMM1b A operand and the optional +dY)
//   tmY   [M, 768]       Y                            dX
//   tmWd  [R, 768]       Wd_cat                       Wd_cat
//   tmW2  [768, R]       Wu_cat                       WdT_cat
//   tmW1b [R, 768]       (unused)                     WuT_cat
template <bool kBwd>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWd,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW1b,
                 const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 2 + 2 + 2 + NSTG];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R, S = p.stages, N2 = p.n2;
  const int KC2 = (R + 63) / 64;
  const int NC2 = (kBwd && !p.has_out) ? 0 : kD / N2;
  const int chunks_per_tile = NC2 * (N2 / 64);

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
    tma_prefetch_desc(&tmW2);
    if (kBwd) tma_prefetch_desc(&tmW1b);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  // biases -> smem (broadcast reads in the epilogues)
  for (int i = tid; i < R; i += NUM_THREADS) bias_smem[i] = p.bd[i];
  if (!kBwd)
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
        if (kBwd) {
          for (int kc = 0; kc < KC1; ++kc) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            const uint32_t a_dst = smem0 + stage * stage_bytes;
            mbar_arrive_expect_tx(bar_full(stage), A_SLOT + R * 128);
            tma_load_2d_hint(a_dst, &tmRes, bar_full(stage), kc * BK, m0, kEvictNormal);
            tma_load_2d_hint(a_dst + A_SLOT, &tmW1b, bar_full(stage), kc * BK, 0, kEvictLast);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
        for (int nc = 0; nc < NC2; ++nc) {
          for (int kc = 0; kc < KC2; ++kc) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            const uint32_t b_dst = smem0 + stage * stage_bytes + A_SLOT;
            mbar_arrive_expect_tx(bar_full(stage), N2 * 128);
            tma_load_2d_hint(b_dst, &tmW2, bar_full(stage), kc * BK, nc * N2, kEvictLast);
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
        for (int g1 = 0; g1 < (kBwd ? 2 : 1); ++g1) {  // GEMM1: P = X Wd_cat^T  (bwd: + dH = dY Wu_cat)
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
    const bool has_res = p.has_res != 0;
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + gridDim.x - 1) / gridDim.x;
    const uint32_t total_chunks = static_cast<uint32_t>(my_tiles) * chunks_per_tile;

    auto issue_res_load = [&](uint32_t g) {
      if (!has_res || g >= total_chunks) return;
      const int tile = blockIdx.x + (g / chunks_per_tile) * gridDim.x;
      const int c = g % chunks_per_tile;
      const uint32_t sb = g % NSTG;
      mbar_arrive_expect_tx(bar_res_full(sb), STG_BYTES);
      tma_load_2d(stg_base + sb * STG_BYTES, &tmRes, bar_res_full(sb), c * 64, tile * BM);
    };
    if (leader)
      for (uint32_t g = 0; g < NSTG - 1; ++g) issue_res_load(g);

    uint32_t acc_it = 0, tile_it = 0, chunk_g = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_it) {
      const int m0 = tile * BM;
      if constexpr (!kBwd) {  // epilogue 1 (forward): hidden tile H = act(P + bd)
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
      } else {  // epilogue 1 (backward): dP = scale * dH * act'(P + bd); H_t / dP_t slices to HBM
        const uint32_t buf_p = acc_it & 1, par_p = (acc_it >> 1) & 1;
        const uint32_t buf_g = (acc_it + 1) & 1, par_g = ((acc_it + 1) >> 1) & 1;
        mbar_wait(bar_acc_full(buf_p), par_p);
        mbar_wait(bar_acc_full(buf_g), par_g);
        mbar_wait(bar_h_empty, (tile_it & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_p = tmem + lane_addr + buf_p * 256;
        const uint32_t t_g = tmem + lane_addr + buf_g * 256;
        const int grow = m0 + static_cast<int>(row);
        const int rt = p.r_hi - p.r_lo;
        for (int c = 0; c < R / 16; ++c) {
          uint32_t v[16], u[16];
          tmem_ld16(t_p + c * 16, v);
          tmem_ld16(t_g + c * 16, u);
          tmem_ld_wait();
          uint32_t w[8], hh[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float p0 = __uint_as_float(v[2 * i]) + bias_smem[c * 16 + 2 * i];
            const float p1 = __uint_as_float(v[2 * i + 1]) + bias_smem[c * 16 + 2 * i + 1];
            w[i] = pack_bf16x2(scale * __uint_as_float(u[2 * i]) * act_grad(p0, act),
                               scale * __uint_as_float(u[2 * i + 1]) * act_grad(p1, act));
            hh[i] = pack_bf16x2(apply_act(p0, act), apply_act(p1, act));
          }
          const uint32_t kc = c >> 2, j0 = (c & 3) * 2;
          st_shared_v4(h_base + kc * A_SLOT + sw128_offset(row, j0), w[0], w[1], w[2], w[3]);
          st_shared_v4(h_base + kc * A_SLOT + sw128_offset(row, j0 + 1), w[4], w[5], w[6], w[7]);
          const int col = c * 16;
          if (p.H_t != nullptr && col >= p.r_lo && col < p.r_hi && grow < p.M) {
            const size_t off = static_cast<size_t>(grow) * rt + (col - p.r_lo);
            uint4* hd = reinterpret_cast<uint4*>(p.H_t + off);
            uint4* gd = reinterpret_cast<uint4*>(p.dP_t + off);
            hd[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            hd[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
            gd[0] = make_uint4(w[0], w[1], w[2], w[3]);
            gd[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar_acc_empty(buf_p));
        mbar_arrive(bar_acc_empty(buf_g));
        mbar_arrive(bar_h_full);
        acc_it += 2;
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
          if (has_res) mbar_wait(bar_res_full(sb), rpar);
          tmem_ld_wait();
          const uint32_t sbuf = stg_base + sb * STG_BYTES;
          const float* bu = bias_smem + R + col0;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint32_t addr = sbuf + sw128_offset(row, c8);
            uint4 rv = make_uint4(0u, 0u, 0u, 0u);
            if (has_res) rv = ld_shared_v4(addr);
            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = c8 * 8 + 2 * i;
              const float a0 = __uint_as_float(e < 32 ? v0[e & 31] : v1[e & 31]);
              const float a1 = __uint_as_float(e < 32 ? v0[(e + 1) & 31] : v1[(e + 1) & 31]);
              const float2 r2 = unpack_bf16x2(rr[i]);
              if constexpr (kBwd)
                o[i] = pack_bf16x2(r2.x + a0, r2.y + a1);
              else
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
          if (!has_res) {
            // no residual barrier paces the staging ring: the buffer about to be rewritten
            // ((chunk_g + 1) % NSTG) was last stored from NSTG - 1 chunks ago, which the leader's
            // wait_group.read<1> above has retired; make that visible to the other 127 threads.
            named_bar_sync(2, 128);
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

int launch_fused(bool bwd, const void* X, const void* Res, void* Out, const void* Wd_cat,
                 const void* W2, const void* W1b, FusedParams p, int64_t M, int r_total,
                 cudaStream_t st, const char* who) {
  int rc;
  p.M = static_cast<int>(M);
  p.R = r_total;
  p.num_tiles = static_cast<int>((M + BM - 1) / BM);
  p.n2 = r_total > 128 ? 256 : 128;
  p.b_slot_bytes = static_cast<uint32_t>(p.n2 * 128);
  const int kc2 = (r_total + 63) / 64;
  const size_t fixed = 1024 + static_cast<size_t>(kc2) * A_SLOT + NSTG * STG_BYTES +
                       (r_total + kD) * sizeof(float);
  const size_t max_smem = 227 * 1024 - 512;  // static smem (barriers) lives in the same budget
  const size_t stage_bytes = A_SLOT + p.b_slot_bytes;
  int stages = static_cast<int>((max_smem - fixed) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  FD_REQUIRE(stages >= 2, FD_ERR_UNSUPPORTED, "%s: shared-memory budget exceeded (R=%d)", who,
             r_total);
  p.stages = stages;
  const size_t smem = fixed + stages * stage_bytes;

  CUtensorMap tmX, tmRes, tmY, tmWd, tmW2, tmW1b;
  if ((rc = make_tmap_bf16_2d(&tmX, X, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmRes, Res, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmY, Out ? Out : X, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWd, Wd_cat, r_total, kD, kD, r_total, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW2, W2, kD, r_total, r_total, p.n2, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW1b, W1b ? W1b : Wd_cat, r_total, kD, kD, r_total, 64))) return rc;

  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  static bool configured[2][64] = {{false}};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[bwd][dev]) {
    if (bwd)
      FD_CHECK_CUDA(cudaFuncSetAttribute(dat_fused_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    else
      FD_CHECK_CUDA(cudaFuncSetAttribute(dat_fused_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    if (dev < 64) configured[bwd][dev] = true;
  }
  if (bwd)
    dat_fused_kernel<true><<<grid, NUM_THREADS, smem, st>>>(tmX, tmRes, tmY, tmWd, tmW2, tmW1b, p);
  else
    dat_fused_kernel<false><<<grid, NUM_THREADS, smem, st>>>(tmX, tmRes, tmY, tmWd, tmW2, tmW1b, p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

int check_common(const char* who, int64_t M, int d, int r_total, int act, int dtype) {
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED,
             "%s: only bf16 activations are implemented (dtype=%d)", who, dtype);
  FD_REQUIRE(d == kD, FD_ERR_UNSUPPORTED, "%s: model_dim must be 768 (got %d)", who, d);
  FD_REQUIRE(r_total >= 16 && r_total <= 256 && r_total % 16 == 0, FD_ERR_UNSUPPORTED,
             "%s: r_total must be a multiple of 16 in [16, 256] (got %d); split wider bottlenecks "
             "into several calls", who, r_total);
  FD_REQUIRE(act == FEDDAT_ACT_RELU || act == FEDDAT_ACT_GELU, FD_ERR_INVALID,
             "%s: unknown activation %d", who, act);
  FD_REQUIRE(M >= 0 && M < (1ll << 31) - 256, FD_ERR_INVALID, "%s: bad row count %lld", who,
             (long long)M);
  return FD_OK;
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
  if ((rc = check_common("dat_fwd", M, d, r_total, act, dtype))) return rc;
  if (M == 0) return FD_OK;
  FusedParams p{};
  p.act = act;
  p.scale = branch_scale;
  p.bd = bd_cat;
  p.bu = bu_cat;
  p.has_out = 1;
  p.has_res = 1;
  return launch_fused(false, X, Res, Y, Wd_cat, Wu_cat, nullptr, p, M, r_total,
                      static_cast<cudaStream_t>(stream), "dat_fwd");
}

extern "C" int feddat_dat_bwd_dgrad(const void* X, const void* dY, void* dX, const void* Wd_cat,
                                    const float* bd_cat, const void* WuT_cat, const void* WdT_cat,
                                    void* H_t, void* dP_t, int r_lo, int r_hi, int64_t M, int d,
                                    int r_total, float branch_scale, int act, int add_dy, int dtype,
                                    void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(X && dY && Wd_cat && bd_cat && WuT_cat && WdT_cat, FD_ERR_INVALID,
             "dat_bwd_dgrad: null pointer argument");
  if ((rc = check_common("dat_bwd_dgrad", M, d, r_total, act, dtype))) return rc;
  FD_REQUIRE((H_t == nullptr) == (dP_t == nullptr), FD_ERR_INVALID,
             "dat_bwd_dgrad: H_t and dP_t must both be given or both be NULL");
  FD_REQUIRE(dX != nullptr || H_t != nullptr, FD_ERR_INVALID,
             "dat_bwd_dgrad: nothing to compute (dX and H_t are both NULL)");
  if (H_t) {
    FD_REQUIRE(r_lo >= 0 && r_hi > r_lo && r_hi <= r_total && r_lo % 16 == 0 && r_hi % 16 == 0,
               FD_ERR_INVALID, "dat_bwd_dgrad: bad trainable slice [%d, %d) of %d", r_lo, r_hi,
               r_total);
    FD_REQUIRE(((reinterpret_cast<uintptr_t>(H_t) | reinterpret_cast<uintptr_t>(dP_t)) & 15) == 0,
               FD_ERR_INVALID, "dat_bwd_dgrad: H_t / dP_t must be 16-byte aligned");
  }
  if (M == 0) return FD_OK;
  FusedParams p{};
  p.act = act;
  p.scale = branch_scale;
  p.bd = bd_cat;
  p.bu = nullptr;
  p.has_out = dX != nullptr;
  p.has_res = (dX != nullptr && add_dy) ? 1 : 0;
  p.r_lo = H_t ? r_lo : 0;
  p.r_hi = H_t ? r_hi : 0;
  p.H_t = static_cast<__nv_bfloat16*>(H_t);
  p.dP_t = static_cast<__nv_bfloat16*>(dP_t);
  return launch_fused(true, X, dY, dX, Wd_cat, WdT_cat, WuT_cat, p, M, r_total,
                      static_cast<cudaStream_t>(stream), "dat_bwd_dgrad");
}
