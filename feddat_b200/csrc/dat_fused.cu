// DAT bottleneck forward (reference: src/modeling/models/adapter.py:124-163).
//
//   Y = Res + scale * ( act(X * Wd_cat^T + bd_cat) * Wu_cat^T + bu_cat )
//
// "cat" = the active branches concatenated along the bottleneck dimension: one branch
// (adapter.py:125-131, scale 1) or the two gating branches adapter_0 | adapter_2
// (adapter.py:133-146, scale 0.5 each), so the dual adapter is ONE GEMM pair with hidden width
// R = r_total.  Persistent CTA PAIRS (cluster of 2, tcgen05 cta_group::2): a pair walks 256-row
// super-tiles, each CTA owning 128 rows (its X chunks, its TMEM accumulators, its epilogues, its
// output) and HALF of every weight tile -- the pair's tensor cores read both halves, so each SM
// ingests only half the weight bytes (L2->SM traffic is what bounds this kernel: measured 8.95 TB/s
// chip-wide, scripts/l2bw.py).  12 warps per CTA, every hand-off an mbarrier:
//
//   warp 0      TMA producer: one ring of uniform 16 KB slots carries, in consumption order, the X
//               k-chunks [128 x 64], the Wd_cat k-chunks (one or two 128-row boxes) and the Wu_cat
//               [128 x 64] tiles; it runs ahead across tiles as far as the ring allows
//   warp 1      tcgen05.mma issuer (leader CTA of the pair only; M = 256 across the pair).  GEMM1 P = X Wd^T (SS, K-major SW128 smem operands) into TMEM
//               columns [0, R); GEMM2 in six 128-column chunks, A operand = the hidden tile IN TMEM
//               (bf16 pairs packed over P's own columns, "TS" MMA), accumulators in a two-buffer
//               ring in TMEM columns [256, 512)
//   warp 2      residual producer: TMA loads of the residual [128 x 64] chunks into the staging ring
//   warp 3      store issuer: TMA stores of finished staging buffers, recycles them
//   warps 4-7   epilogue group A: (1) P -> +bias, act -> bf16 pairs -> tcgen05.st over P's columns
//               (the hidden never leaves the SM); (2) even output chunks
//   warps 8-11  epilogue group B: odd output chunks.  An output chunk = tcgen05.ld, scale, +bias,
//               +residual (from the staging buffer), bf16 back into the same staging buffer
//
// HBM traffic per row: read X (1536 B) + write Y (1536 B); the residual re-read and all weights hit L2.
//
// The same template with kBwd = true is the backward data-gradient pass (what torch autograd derives
// from adapter.py:124-163):
//   GEMM1  P  = X  * Wd_cat^T            (recompute, never stored)          -> TMEM [0, R)
//   GEMM1b dH = dY * Wu_cat              (B operand: WuT_cat, K-major)      -> TMEM [256, 256 + R)
//   epilogue 1: dP = scale * dH * act'(P + bd) -> bf16 pairs over P's columns (A operand of GEMM3);
//               for the trainable slice [r_lo, r_hi) also H = act(P + bd) and dP to HBM (wgrad kernel)
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
constexpr int KC1 = kD / BK;          // 12 k-chunks for GEMM1
constexpr int SLOT = BM * 128;        // 16 KB: [128 rows x 64 bf16], 128-byte swizzle
constexpr int N2 = 128;               // GEMM2 / GEMM3 output chunk width
constexpr int NC2 = kD / N2;          // 6 chunks
constexpr int MAX_SLOTS = 12;
constexpr int MAX_STG = 8;
constexpr int NUM_THREADS = 384;
constexpr uint32_t TM_P = 0;          // TMEM column of P (and of the packed hidden aliasing it)
constexpr uint32_t TM_D = 256;        // TMEM column of the output ring (and of dH in backward)

struct FusedParams {
  int M, R, num_tiles, n_slots, n_stg, act;
  float scale;
  const float* bd;
  const float* bu;        // fwd only
  // bwd only
  int has_out;            // 0: dX not requested (skip GEMM3 / epilogue 2)
  int has_res;            // add the residual tile (fwd: always; bwd: add_dy)
  int r_lo, r_hi;         // trainable slice of the bottleneck
  __nv_bfloat16* H_t;     // [M, r_hi - r_lo] or null
  __nv_bfloat16* dP_t;    // [M, r_hi - r_lo] or null
  unsigned long long* trace;  // debug: globaltimer stamps of CTA 0's pipeline events (or null)
};

// debug timeline (scripts/trace_kernel.py): event e of CTA 0's tile `t` (t < 2) -> trace[t * 128 + e]
#define FD_TRACE(ev, t)                                                             \
  do {                                                                              \
    if (p.trace != nullptr && blockIdx.x == 0 && (t) < 2)                           \
      p.trace[(t) * 128 + (ev)] = globaltimer_ns();                                 \
  } while (0)

// activation is a template parameter: a run-time switch makes ptxas keep the erff path live in the
// epilogue-1 inner loop (measured: 6.8 us instead of ~1.5 us per 128 x 256 tile)
template <bool kGelu>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (!kGelu) return fmaxf(x, 0.f);
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
template <bool kGelu>
__device__ __forceinline__ float act_grad(float x) {
  if constexpr (!kGelu) return x > 0.f ? 1.f : 0.f;
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
template <bool kBwd, bool kGelu>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWd,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW1b,
                 const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  // barriers: slot full/empty, P full, dH full, hidden full, D full/empty x2, staging res/out/empty
  __shared__ __align__(8) uint64_t bars[2 * MAX_SLOTS + 3 + 4 + 3 * MAX_STG];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R, NS = p.n_slots, NSTG = p.n_stg;
  const int KC2 = (R + 63) / 64;
  const int nc2 = (kBwd && !p.has_out) ? 0 : NC2;
  const uint32_t rank = cluster_ctarank();            // 0 = leader of the CTA pair
  const int RH = R / 2;                               // rows of a Wd / WuT k-chunk this CTA holds
  const uint32_t w_half_bytes = static_cast<uint32_t>(RH) * 128u;
  constexpr uint32_t W2_HALF_BYTES = (N2 / 2) * 128u;  // 64 rows of a Wu / WdT tile

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem0 + NS * SLOT;
  const uint32_t bias_base = stg_base + NSTG * SLOT;
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));

  const uint32_t bar0 = smem_u32(bars);
  auto bar_slot_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_slot_empty = [&](int s) { return bar0 + 8u * (MAX_SLOTS + s); };
  const uint32_t bar_p_full = bar0 + 8u * (2 * MAX_SLOTS);
  const uint32_t bar_g_full = bar0 + 8u * (2 * MAX_SLOTS + 1);
  const uint32_t bar_h_full = bar0 + 8u * (2 * MAX_SLOTS + 2);
  auto bar_d_full = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 3 + b); };
  auto bar_d_empty = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 5 + b); };
  auto bar_res_full = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 7 + b); };
  auto bar_out_full = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 7 + MAX_STG + b); };
  auto bar_stg_empty = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 7 + 2 * MAX_STG + b); };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_slot_full(s), 1);
      mbar_init(bar_slot_empty(s), 1);
    }
    mbar_init(bar_p_full, 1);
    mbar_init(bar_g_full, 1);
    mbar_init(bar_h_full, 256);          // both CTAs' epilogue-1 threads arrive at the leader
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d_full(b), 1);
      mbar_init(bar_d_empty(b), 256);    // both CTAs' epilogue-2 threads arrive at the leader
    }
    for (int b = 0; b < NSTG; ++b) {
      mbar_init(bar_res_full(b), 1);
      mbar_init(bar_out_full(b), 128);
      mbar_init(bar_stg_empty(b), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmRes);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmWd);
    tma_prefetch_desc(&tmW2);
    if (kBwd) tma_prefetch_desc(&tmW1b);
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
  for (int i = tid; i < R; i += NUM_THREADS) bias_smem[i] = p.bd[i];
  if (!kBwd)
    for (int i = tid; i < kD; i += NUM_THREADS) bias_smem[R + i] = p.bu[i];
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible to both CTAs of the pair
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) FD_TRACE(0, 0);
  const int num_pairs = (p.num_tiles + 1) / 2, pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
  const int my_tiles = (num_pairs - pair0 + pair_stride - 1) / pair_stride;   // super-tiles of this pair
  const uint32_t total_chunks = static_cast<uint32_t>(my_tiles) * nc2 * 2;    // 64-column staging chunks
  // this CTA's tile of super-tile `it`: may lie beyond the tensor (odd tile count) -- TMA then
  // zero-fills the loads and clips the stores, so no role needs a special case
  auto tile_of = [&](int it) { return 2 * (pair0 + it * pair_stride) + static_cast<int>(rank); };

  if (warp == 0) {
    // ------------------------------------------------------------------ ring producer
    if (lane == 0) {
      uint32_t n = 0;  // slots issued so far
      // the leader's "full" barrier collects the bytes of BOTH CTAs' loads for a slot
      auto acquire = [&](uint32_t bytes) -> uint32_t {
        const uint32_t s = n % NS, par = (n / NS) & 1;
        mbar_wait(bar_slot_empty(s), par ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(bar_slot_full(s), 2 * bytes);
        ++n;
        return s;
      };
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t tile_it = it;
        const int m0 = tile_of(it) * BM;
        FD_TRACE(110, tile_it);
        for (int pass = 0; pass < (kBwd ? 2 : 1); ++pass) {
          const CUtensorMap* ta = pass == 0 ? &tmX : &tmRes;
          const CUtensorMap* tw = pass == 0 ? &tmWd : &tmW1b;
          for (int kc = 0; kc < KC1; ++kc) {
            uint32_t s = acquire(SLOT);
            tma_load_2d_pair(smem0 + s * SLOT, ta, mapa_u32(bar_slot_full(s), 0), kc * BK, m0,
                             kEvictNormal);
            s = acquire(w_half_bytes);
            tma_load_2d_pair(smem0 + s * SLOT, tw, mapa_u32(bar_slot_full(s), 0), kc * BK,
                             static_cast<int>(rank) * RH, kEvictLast);
          }
        }
        FD_TRACE(111, tile_it);
        for (int c = 0; c < nc2; ++c)
          for (int kc = 0; kc < KC2; kc += 2) {   // two 8 KB half tiles share one 16 KB slot
            const int nk = min(2, KC2 - kc);
            const uint32_t s = acquire(W2_HALF_BYTES * nk);
            for (int j = 0; j < nk; ++j)
              tma_load_2d_pair(smem0 + s * SLOT + j * W2_HALF_BYTES, &tmW2,
                               mapa_u32(bar_slot_full(s), 0), (kc + j) * BK,
                               c * N2 + static_cast<int>(rank) * (N2 / 2), kEvictLast);
          }
        FD_TRACE(112, tile_it);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0) {
      uint32_t n = 0;
      uint32_t de[2] = {0, 0};  // uses of each D buffer so far (parity of its "empty" barrier)
      const uint32_t idesc1 = make_idesc_bf16(2 * BM, R);
      const uint32_t idesc2 = make_idesc_bf16(2 * BM, N2);
      auto wait_slot = [&]() -> uint32_t {
        const uint32_t s = n % NS, par = (n / NS) & 1;
        mbar_wait(bar_slot_full(s), par);
        ++n;
        return s;
      };
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t tile_it = it;
        for (int pass = 0; pass < (kBwd ? 2 : 1); ++pass) {
          // pass 0: P = X Wd_cat^T -> [TM_P, +R);  pass 1 (bwd): dH = dY Wu_cat -> [TM_D, +R)
          const uint32_t d_tmem = tmem + (pass == 0 ? TM_P : TM_D);
          if (pass == 1) {  // dH overlays the output ring: both buffers must have been drained
            for (int b = 0; b < 2; ++b) {
              mbar_wait(bar_d_empty(b), (de[b] & 1) ^ 1);
              ++de[b];
            }
            tc_fence_after();
          }
          for (int kc = 0; kc < KC1; ++kc) {
            const uint32_t sa = wait_slot();
            const uint32_t sb = wait_slot();
            tc_fence_after();
            if (pass == 0) FD_TRACE(10 + kc, tile_it);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_pair(d_tmem, desc_kmajor_sw128(smem0 + sa * SLOT + k * 32),
                           desc_kmajor_sw128(smem0 + sb * SLOT + k * 32), idesc1, (kc | k) != 0);
            umma_commit_pair(bar_slot_empty(sa), 0b11);
            umma_commit_pair(bar_slot_empty(sb), 0b11);
          }
          umma_commit_pair(pass == 0 ? bar_p_full : bar_g_full, 0b11);
          FD_TRACE(22 + pass, tile_it);
        }
        // epilogue 1 done in BOTH CTAs: the packed hidden (dP) is in TMEM [TM_P, TM_P + R/2) and P
        // may be overwritten by the next tile's GEMM1 (waited even when no GEMM2/3 follows)
        mbar_wait(bar_h_full, tile_it & 1);
        tc_fence_after();
        FD_TRACE(24, tile_it);
        for (int c = 0; c < nc2; ++c) {
          const int b = c & 1;
          mbar_wait(bar_d_empty(b), (de[b] & 1) ^ 1);
          ++de[b];
          tc_fence_after();
          FD_TRACE(25 + c, tile_it);
          const uint32_t d_tmem = tmem + TM_D + b * N2;
          for (int kc = 0; kc < KC2; kc += 2) {
            const uint32_t s = wait_slot();
            tc_fence_after();
            const int nk = min(2, KC2 - kc);
            for (int j = 0; j < nk; ++j) {
              const int ksteps = min(4, (R - (kc + j) * 64) / 16);
              for (int k = 0; k < ksteps; ++k)
                umma_ts_pair(d_tmem, tmem + TM_P + ((kc + j) * 4 + k) * 8,
                             desc_kmajor_sw128(smem0 + s * SLOT + j * (N2 / 2) * 128 + k * 32),
                             idesc2, (kc | j | k) != 0);
            }
            umma_commit_pair(bar_slot_empty(s), 0b11);
          }
          umma_commit_pair(bar_d_full(b), 0b11);
          FD_TRACE(31 + c, tile_it);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual producer
    if (lane == 0) {
      const uint32_t per_tile = nc2 * 2;
      for (uint32_t g = 0; g < total_chunks; ++g) {
        const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
        mbar_wait(bar_stg_empty(sb), par ^ 1);
        if (p.has_res) {
          const int tile = tile_of(g / per_tile);
          const int c64 = g % per_tile;
          mbar_arrive_expect_tx(bar_res_full(sb), SLOT);
          tma_load_2d(stg_base + sb * SLOT, &tmRes, bar_res_full(sb), c64 * 64, tile * BM);
          if (lane == 0) FD_TRACE(90 + c64, g / per_tile);
        } else {
          mbar_arrive(bar_res_full(sb));
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ store issuer
    if (lane == 0) {
      const uint32_t per_tile = nc2 * 2;
      for (uint32_t g = 0; g < total_chunks; ++g) {
        const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
        mbar_wait(bar_out_full(sb), par);
        const int tile = tile_of(g / per_tile);
        const int c64 = g % per_tile;
        tma_store_2d(&tmY, stg_base + sb * SLOT, c64 * 64, tile * BM);
        tma_store_commit();
        FD_TRACE(104 + (c64 >> 1), g / per_tile);
        if (g > 0) {  // the previous store has finished reading its buffer: recycle it
          tma_store_wait_read<1>();
          mbar_arrive(bar_stg_empty((g - 1) % NSTG));
        }
      }
      if (total_chunks > 0) {
        tma_store_wait_read<0>();
        mbar_arrive(bar_stg_empty((total_chunks - 1) % NSTG));
        tma_store_wait_all<0>();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue groups A / B
    const int group = (warp - 4) >> 2;      // 0: warps 4-7, 1: warps 8-11
    const uint32_t q = warp & 3;            // TMEM lane quarter this warp may touch
    const uint32_t row = q * 32 + lane;     // tile row == TMEM lane
    const uint32_t lane_addr = (q * 32) << 16;
    const float scale = p.scale;
    const bool has_res = p.has_res != 0;
    uint32_t df = 0;  // chunk fills of this group's D buffer so far
    // the MMA issuer lives in the leader CTA: "hidden ready" / "D buffer drained" go to ITS barriers
    const uint32_t leader_h_full = mapa_u32(bar_h_full, 0);
    const uint32_t leader_d_empty[2] = {mapa_u32(bar_d_empty(0), 0), mapa_u32(bar_d_empty(1), 0)};

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t tile_it = it;
      const int m0 = tile_of(it) * BM;
      if (group == 0) {
        // ---------------- epilogue 1: P (and dH) -> packed bf16 hidden over P's own columns
        mbar_wait(bar_p_full, tile_it & 1);
        if (kBwd) mbar_wait(bar_g_full, tile_it & 1);
        tc_fence_after();
        if (tid == 128) FD_TRACE(40, tile_it);
        const uint32_t t_p = tmem + lane_addr + TM_P;
        const uint32_t t_g = tmem + lane_addr + TM_D;
        const int grow = m0 + static_cast<int>(row);
        const int rt = p.r_hi - p.r_lo;
        for (int c = 0; c < R / 16; ++c) {
          uint32_t v[16], u[16], w[8];
          tmem_ld16(t_p + c * 16, v);
          if (kBwd) tmem_ld16(t_g + c * 16, u);
          tmem_ld_wait();
          if constexpr (!kBwd) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              w[i] = pack_bf16x2(
                  apply_act<kGelu>(__uint_as_float(v[2 * i]) + bias_smem[c * 16 + 2 * i]),
                  apply_act<kGelu>(__uint_as_float(v[2 * i + 1]) + bias_smem[c * 16 + 2 * i + 1]));
          } else {
            uint32_t hh[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float p0 = __uint_as_float(v[2 * i]) + bias_smem[c * 16 + 2 * i];
              const float p1 = __uint_as_float(v[2 * i + 1]) + bias_smem[c * 16 + 2 * i + 1];
              w[i] = pack_bf16x2(scale * __uint_as_float(u[2 * i]) * act_grad<kGelu>(p0),
                                 scale * __uint_as_float(u[2 * i + 1]) * act_grad<kGelu>(p1));
              hh[i] = pack_bf16x2(apply_act<kGelu>(p0), apply_act<kGelu>(p1));
            }
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
          // columns [8c, 8c+8) were read (as fp32 P columns) in an earlier iteration: safe to reuse
          tmem_st8(t_p + c * 8, w);
        }
        tmem_st_wait();
        tc_fence_before();
        if (kBwd) {  // dH (which overlays the output ring) is consumed
          mbar_arrive_cluster_addr(leader_d_empty[0]);
          mbar_arrive_cluster_addr(leader_d_empty[1]);
        }
        mbar_arrive_cluster_addr(leader_h_full);
        if (tid == 128) FD_TRACE(41, tile_it);
      }
      // ---------------- epilogue 2: this group's output chunks (c = group, group + 2, group + 4)
      for (int c = group; c < nc2; c += 2) {
        const int b = c & 1;
        mbar_wait(bar_d_full(b), df & 1);
        ++df;
        tc_fence_after();
        if (lane == 0 && q == 0) FD_TRACE(42 + 4 * c, tile_it);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const uint32_t g = (tile_it * nc2 + c) * 2 + j;
          const uint32_t sb = g % NSTG, rpar = (g / NSTG) & 1;
          const int col0 = c * N2 + j * 64;
          const uint32_t t_src = tmem + lane_addr + TM_D + b * N2 + j * 64;
          uint32_t v0[32], v1[32];
          tmem_ld32(t_src, v0);
          tmem_ld32(t_src + 32, v1);
          mbar_wait(bar_res_full(sb), rpar);
          tmem_ld_wait();
          if (lane == 0 && q == 0) FD_TRACE(43 + 4 * c + j, tile_it);
          const uint32_t sbuf = stg_base + sb * SLOT;
          const float4* bu4 = reinterpret_cast<const float4*>(bias_smem + R + col0);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint32_t addr = sbuf + sw128_offset(row, c8);
            uint4 rv = make_uint4(0u, 0u, 0u, 0u);
            if (has_res) rv = ld_shared_v4(addr);
            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
            float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if constexpr (!kBwd) {
              const float4 b0 = bu4[2 * c8], b1 = bu4[2 * c8 + 1];
              bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
              bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
            }
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = c8 * 8 + 2 * i;
              const float a0 = __uint_as_float(e < 32 ? v0[e & 31] : v1[e & 31]);
              const float a1 = __uint_as_float(e < 32 ? v0[(e + 1) & 31] : v1[(e + 1) & 31]);
              const float2 r2 = unpack_bf16x2(rr[i]);
              if constexpr (kBwd)
                o[i] = pack_bf16x2(r2.x + a0, r2.y + a1);
              else   // res + scale * (acc + bias)
                o[i] = pack_bf16x2(fmaf(scale, a0 + bb[2 * i], r2.x), fmaf(scale, a1 + bb[2 * i + 1], r2.y));
            }
            st_shared_v4(addr, o[0], o[1], o[2], o[3]);
          }
          fence_proxy_async_smem();
          mbar_arrive(bar_out_full(sb));
        }
        tc_fence_before();
        mbar_arrive_cluster_addr(leader_d_empty[b]);
        if (lane == 0 && q == 0) FD_TRACE(45 + 4 * c, tile_it);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the partner's smem / barriers / TMEM stay valid until both CTAs are done
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
}

unsigned long long* g_trace = nullptr;

int launch_fused(bool bwd, const void* X, const void* Res, void* Out, const void* Wd_cat,
                 const void* W2, const void* W1b, FusedParams p, int64_t M, int r_total,
                 cudaStream_t st, const char* who) {
  int rc;
  p.M = static_cast<int>(M);
  p.R = r_total;
  p.num_tiles = static_cast<int>((M + BM - 1) / BM);
  p.trace = g_trace;
  p.n_slots = 7;    // 16 KB TMA ring slots (X chunks, weight half-chunks)
  p.n_stg = 6;      // 16 KB residual-in / output staging buffers
  const size_t max_smem = 227 * 1024 - 1024;  // static smem (barriers) lives in the same budget
  const size_t smem = 1024 + static_cast<size_t>(p.n_slots + p.n_stg) * SLOT +
                      (r_total + kD) * sizeof(float);
  FD_REQUIRE(smem <= max_smem, FD_ERR_UNSUPPORTED, "%s: shared-memory budget exceeded (R=%d)", who,
             r_total);

  CUtensorMap tmX, tmRes, tmY, tmWd, tmW2, tmW1b;
  if ((rc = make_tmap_bf16_2d(&tmX, X, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmRes, Res, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmY, Out ? Out : X, M, kD, kD, BM, 64))) return rc;
  const uint32_t w_box_rows = r_total / 2;   // each CTA of a pair holds half of every weight tile
  if ((rc = make_tmap_bf16_2d(&tmWd, Wd_cat, r_total, kD, kD, w_box_rows, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW2, W2, kD, r_total, r_total, N2 / 2, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW1b, W1b ? W1b : Wd_cat, r_total, kD, kD, w_box_rows, 64))) return rc;

  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  const int num_pairs = (p.num_tiles + 1) / 2;
  const int grid = 2 * (num_pairs < sms / 2 ? num_pairs : sms / 2);
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap,
                            const CUtensorMap, const CUtensorMap, const CUtensorMap,
                            const FusedParams);
  const bool gelu = p.act == FEDDAT_ACT_GELU;
  KernelFn fn = bwd ? (gelu ? dat_fused_kernel<true, true> : dat_fused_kernel<true, false>)
                    : (gelu ? dat_fused_kernel<false, true> : dat_fused_kernel<false, false>);
  static bool configured[2][2][64] = {{{false}}};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[bwd][gelu][dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)max_smem));
    if (dev < 64) configured[bwd][gelu][dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, tmX, tmRes, tmY, tmWd, tmW2, tmW1b, p));
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

// debug: device buffer of 256 uint64 receiving CTA 0's pipeline timestamps (NULL disables)
extern "C" int feddat_debug_set_trace(void* dev_buf) {
  fd::g_trace = static_cast<unsigned long long*>(dev_buf);
  return 0;
}
