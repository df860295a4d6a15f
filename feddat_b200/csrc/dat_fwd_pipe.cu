// DAT bottleneck forward for MANY tiles per SM (reference: src/modeling/models/adapter.py:124-163):
//
//   Y = Res + scale * ( act(X * Wd_cat^T + bd_cat) * Wu_cat^T + bu_cat )
//
// Same CTA-pair decomposition as dat_fused_kernel<false> (dat_fused.cu: a pair of CTAs walks 256-row
// super-tiles, each CTA owns 128 rows and half of every weight tile), but software-pipelined ACROSS
// tiles: in dat_fused_kernel the hidden is packed in place over P and the output ring takes the other
// half of TMEM, so GEMM1 of tile t+1 cannot start before GEMM2 of tile t has finished and the three
// phases of a tile (GEMM1 5.9 us, epilogue 1 0.9 us, GEMM2 + epilogue 2 7.5 us) run back to back.
// Here TMEM is laid out so that both GEMMs are in flight at once:
//
//   columns [  0, 256)  P      = X Wd_cat^T of tile t+1          (GEMM1, fp32)
//   columns [256, 384)  H      = packed bf16 hidden of tile t     (written OUT of place by epilogue 1)
//   columns [384, 512)  D ring = two 64-column output accumulators of tile t (GEMM2)
//
// and the MMA issuer alternates one 64-column output chunk of tile t with one GEMM1 k-chunk of tile
// t+1.  The price is 64-wide output chunks (twelve per tile, N = 64 MMAs) and separate rings:
//
//   G1 ring   3 stages x 32 KB   [X k-chunk 128 x 64 | Wd_cat half k-chunk R/2 x 64]
//   W2 ring   3 slots  x 16 KB   this CTA's half [32 x R] of the Wu_cat tile of one output chunk
//   staging   4 slots  x 16 KB   residual in / output out, [128 x 64]
// (3 / 3 / 4 is the measured optimum of the 208 KB: 4 / 2 / 3 runs 13 % slower, 2 / 3 / 6 5 % slower --
// every ring wants more bytes in flight than shared memory holds.)
//
//   warp 0      lanes 0 / 1: G1 ring producers (activations / weights), + L2 prefetch of the next tile
//   warp 1      MMA issuer (whole warp, one elected lane issues)
//   warp 2      store issuer (TMA stores of finished staging buffers)
//   warp 3      residual + W2 producer (one load of each per output chunk)
//   warps 4-11  epilogue groups A / B: (1) half of P each -> +bias, act -> bf16 pairs -> H;
//               (2) group b drains the output chunks with parity b: tcgen05.ld, D buffer handed back at
//               once, y = res + bf16(scale * (acc + bu)) as a packed bf16 add, staging, TMA store
//
// kBwd = true is the SAVED-mode backward data gradient on the same pipeline (ReLU; the forward saved the
// hidden, see dat_fused.cu): GEMM1 = dH = dY WuT_cat^T into "P", epilogue 1 = dP = scale * dH * (H_in > 0)
// packed into "H" (the trainable slice also goes to HBM for the weight-gradient kernel), GEMM2 =
// dX = dP WdT_cat^T, epilogue 2 = (+ dY) -> bf16 -> TMA store.
//
// Used by feddat_dat_fwd / feddat_dat_bwd_dgrad when there are more 256-row super-tiles than CTA pairs;
// single-site launches keep dat_fused_kernel with its column split.
#include "dat_kernels.h"
#include "feddat_b200.h"
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {
namespace {

constexpr int kD = 768;
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int KC1 = kD / BK;            // 12 k-chunks for GEMM1
constexpr int SLOT = BM * 128;          // 16 KB: [128 rows x 64 bf16], 128-byte swizzle
constexpr int G1STAGE = 2 * SLOT;       // 32 KB
constexpr int NG1 = 3;
constexpr int NW2 = 3;
constexpr int NSTG = 4;
constexpr int N2 = 64;                  // output chunk width
constexpr int NC2 = kD / N2;            // 12 chunks per tile
constexpr int NUM_THREADS = 384;
constexpr uint32_t TM_P = 0, TM_H = 256, TM_D = 384;
constexpr uint32_t W2_KB_BYTES = (N2 / 2) * 128u;   // one k-block [32 rows x 64] of a half W2 tile = 4 KB

struct PipeParams {
  int M, R, num_tiles, w2_3d;
  float scale;
  const float* bd;          // fwd
  const float* bu;          // fwd
  __nv_bfloat16* H_out;     // fwd: save the hidden [M, R] for a saved-mode backward, or null
  // bwd (saved mode)
  const __nv_bfloat16* H_in;   // the forward's hidden [M, R]
  __nv_bfloat16* dP_t;         // pre-activation gradient of the trainable slice (row stride ld_t) or null
  int ld_t, r_lo, r_hi;
  int has_res;                 // add dY to dX (the residual input was X itself)
  unsigned long long* trace;
};

#ifdef FEDDAT_DEBUG
#define FDP_TRACE(ev, t)                                                            \
  do {                                                                              \
    if (p.trace != nullptr && blockIdx.x == 0 && (t) < 2)                           \
      p.trace[(t) * 128 + (ev)] = globaltimer_ns();                                 \
  } while (0)
#else
#define FDP_TRACE(ev, t) do { (void)(t); } while (0)
#endif

template <bool kGelu>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (!kGelu) return fmaxf(x, 0.f);
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}

// Tensor maps:   forward            backward (kBwd)
//   tmX          X                  dY          (GEMM1 A operand)
//   tmRes        residual input     dY          (the optional + dY)
//   tmY          Y                  dX
//   tmWd         Wd_cat             WuT_cat     (GEMM1 B operand, [R, 768])
//   tmW2 / k     Wu_cat             WdT_cat     (GEMM2 B operand, [768, R])
template <bool kBwd, bool kGelu>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_pipe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmRes,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWd,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW2k,
                    const PipeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NG1 + 2 * NW2 + 3 + 4 + 3 * NSTG];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) FDP_TRACE(1, 0);
  const int R = p.R;
  const int KC2 = (R + 63) / 64;
  const int n16 = R / 16, nA = (n16 + 1) / 2;
  const uint32_t rank = cluster_ctarank();
  const int RH = R / 2;
  const uint32_t w_half_bytes = static_cast<uint32_t>(RH) * 128u;

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w2_base = smem0 + NG1 * G1STAGE;
  const uint32_t stg_base = w2_base + NW2 * SLOT;
  const uint32_t bias_base = stg_base + NSTG * SLOT;
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));

  const uint32_t bar0 = smem_u32(bars);
  auto bar_g1_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto bar_g1_empty = [&](uint32_t s) { return bar0 + 8u * (NG1 + s); };
  auto bar_w2_full = [&](uint32_t s) { return bar0 + 8u * (2 * NG1 + s); };
  auto bar_w2_empty = [&](uint32_t s) { return bar0 + 8u * (2 * NG1 + NW2 + s); };
  constexpr uint32_t B0 = 2 * NG1 + 2 * NW2;
  const uint32_t bar_p_full = bar0 + 8u * B0;
  const uint32_t bar_h_full = bar0 + 8u * (B0 + 1);
  const uint32_t bar_h_free = bar0 + 8u * (B0 + 2);
  auto bar_d_full = [&](int b) { return bar0 + 8u * (B0 + 3 + b); };
  auto bar_d_empty = [&](int b) { return bar0 + 8u * (B0 + 5 + b); };
  auto bar_res_full = [&](uint32_t b) { return bar0 + 8u * (B0 + 7 + b); };
  auto bar_out_full = [&](uint32_t b) { return bar0 + 8u * (B0 + 7 + NSTG + b); };
  auto bar_stg_empty = [&](uint32_t b) { return bar0 + 8u * (B0 + 7 + 2 * NSTG + b); };

  if (tid == 0) {
    for (int s = 0; s < NG1; ++s) {
      mbar_init(bar_g1_full(s), 1);
      mbar_init(bar_g1_empty(s), 1);
    }
    for (int s = 0; s < NW2; ++s) {
      mbar_init(bar_w2_full(s), 1);
      mbar_init(bar_w2_empty(s), 1);
    }
    mbar_init(bar_p_full, 1);
    mbar_init(bar_h_full, 16);           // one lane of each of the 8 epilogue warps, both CTAs
    mbar_init(bar_h_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d_full(b), 1);
      mbar_init(bar_d_empty(b), 8);      // the 4 warps of the group that drains buffer b, both CTAs
    }
    for (int b = 0; b < NSTG; ++b) {
      mbar_init(bar_res_full(b), 1);
      mbar_init(bar_out_full(b), 4);
      mbar_init(bar_stg_empty(b), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmRes);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmWd);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW2k);
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) FDP_TRACE(0, 0);
  const int num_pairs = (p.num_tiles + 1) / 2, pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
  const int my_tiles = (num_pairs - pair0 + pair_stride - 1) / pair_stride;
  auto tile_of = [&](int it) { return 2 * (pair0 + it * pair_stride) + static_cast<int>(rank); };
  const uint32_t leader_g1_full0 = mapa_u32(bar_g1_full(0), 0);
  const uint32_t leader_w2_full0 = mapa_u32(bar_w2_full(0), 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ G1 ring producers
    if (lane < 2) {
      uint32_t n = 0;
      const CUtensorMap* tm = lane == 0 ? &tmX : &tmWd;
      const uint64_t pol = kEvictLast;
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = tile_of(it) * BM;
        const int c1 = lane == 0 ? m0 : static_cast<int>(rank) * RH;
        if (lane == 0) FDP_TRACE(110, it);
        for (int kc = 0; kc < KC1; ++kc, ++n) {
          const uint32_t s = n % NG1, par = (n / NG1) & 1;
          mbar_wait(bar_g1_empty(s), par ^ 1);
          if (lane == 0 && rank == 0) mbar_arrive_expect_tx(bar_g1_full(s), 2 * (SLOT + w_half_bytes));
          tma_load_2d_pair(smem0 + s * G1STAGE + lane * SLOT, tm, leader_g1_full0 + 8u * s, kc * BK, c1, pol);
        }
        if (lane == 0) FDP_TRACE(111, it);
        if (lane == 0 && it + 1 < my_tiles) {
          const int m1 = tile_of(it + 1) * BM;
          for (int kc = 0; kc < KC1; ++kc) tma_prefetch_l2_2d(&tmX, kc * BK, m1);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      uint32_t n1 = 0, n2 = 0;        // G1 stages / W2 slots consumed so far
      uint32_t de[2] = {0, 0};
      const uint32_t idesc1 = make_idesc_bf16(2 * BM, R);
      const uint32_t idesc2 = make_idesc_bf16(2 * BM, N2);
      auto g1_stage = [&](int kc, bool last, int trace_tile) {
        const uint32_t s = n1 % NG1, par = (n1 / NG1) & 1;
        mbar_wait(bar_g1_full(s), par);
        tc_fence_after();
        if (elect_one()) {
          FDP_TRACE(10 + kc, trace_tile);
          const uint64_t adesc = desc_kmajor_sw128(smem0 + s * G1STAGE);
          const uint64_t bdesc = adesc + (SLOT >> 4);
          umma_ss_pair(tmem + TM_P, adesc, bdesc, idesc1, kc != 0);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_ss_pair(tmem + TM_P, adesc + 2 * k, bdesc + 2 * k, idesc1, 1);
          umma_commit_pair(bar_g1_empty(s), 0b11);
          if (last) umma_commit_pair(bar_p_full, 0b11);
        }
        __syncwarp();
        ++n1;
      };
      // prologue: GEMM1 of the first tile
      for (int kc = 0; kc < KC1; ++kc) g1_stage(kc, kc == KC1 - 1, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const bool has_next = it + 1 < my_tiles;
        // hidden(it) is in H (epilogue 1 done in BOTH CTAs) and P has been read: P is free for GEMM1(it+1)
        mbar_wait(bar_h_full, it & 1);
        tc_fence_after();
        if (elect_one()) FDP_TRACE(24, it);
        __syncwarp();
        for (int c = 0; c < NC2; ++c) {
          const int b = c & 1;
          mbar_wait(bar_d_empty(b), (de[b] & 1) ^ 1);
          ++de[b];
          const uint32_t s = n2 % NW2, par = (n2 / NW2) & 1;
          mbar_wait(bar_w2_full(s), par);
          tc_fence_after();
          if (elect_one()) {
            if (c < 6) FDP_TRACE(25 + c, it);
            const uint32_t d_tmem = tmem + TM_D + b * N2;
            const uint64_t bdesc = desc_kmajor_sw128(w2_base + s * SLOT);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
              if (kk < n16)
                umma_ts_pair(d_tmem, tmem + TM_H + 8 * kk,
                             bdesc + (((kk >> 2) * W2_KB_BYTES + (kk & 3) * 32) >> 4), idesc2, kk != 0 ? 1u : 0u);
            }
            umma_commit_pair(bar_w2_empty(s), 0b11);
            umma_commit_pair(bar_d_full(b), 0b11);
            if (c == NC2 - 1) umma_commit_pair(bar_h_free, 0b11);   // every read of H(it) has completed
            if (c < 6) FDP_TRACE(31 + c, it);
          }
          __syncwarp();
          ++n2;
          // one GEMM1 k-chunk of the NEXT tile between two output chunks of this one
          if (has_next) g1_stage(c, c == KC1 - 1, it + 1);
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ residual + W2 producer
    if (lane == 0) {
      uint32_t g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = tile_of(it) * BM;
        for (int c = 0; c < NC2; ++c, ++g) {
          {   // W2 tile of chunk c: this CTA's 32 rows, every k-block
            const uint32_t s = g % NW2, par = (g / NW2) & 1;
            mbar_wait(bar_w2_empty(s), par ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(bar_w2_full(s), 2 * KC2 * W2_KB_BYTES);
            const int row0 = c * N2 + static_cast<int>(rank) * (N2 / 2);
            if (p.w2_3d) {
              tma_load_3d_pair(w2_base + s * SLOT, &tmW2k, leader_w2_full0 + 8u * s, 0, row0, 0, kEvictLast);
            } else {
              for (int kb = 0; kb < KC2; ++kb)
                tma_load_2d_pair(w2_base + s * SLOT + kb * W2_KB_BYTES, &tmW2, leader_w2_full0 + 8u * s, kb * BK,
                                 row0, kEvictLast);
            }
          }
          {   // residual chunk c
            const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
            mbar_wait(bar_stg_empty(sb), par ^ 1);
            if (!kBwd || p.has_res) {
              mbar_arrive_expect_tx(bar_res_full(sb), SLOT);
              tma_load_2d(stg_base + sb * SLOT, &tmRes, bar_res_full(sb), c * N2, m0);
            } else {
              mbar_arrive(bar_res_full(sb));
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ store issuer
    if (lane == 0) {
      uint32_t g = 0;
      const uint32_t total = static_cast<uint32_t>(my_tiles) * NC2;
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = tile_of(it) * BM;
        for (int c = 0; c < NC2; ++c, ++g) {
          const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
          mbar_wait(bar_out_full(sb), par);
          // outputs are never re-read by this kernel: let them leave L2 first, the activations (re-read as the
          // residual a few microseconds after GEMM1 consumed them) last
          tma_store_2d_hint(&tmY, stg_base + sb * SLOT, c * N2, m0, kEvictFirst);
          tma_store_commit();
          if (g > 0) {
            tma_store_wait_read<1>();
            mbar_arrive(bar_stg_empty((g - 1) % NSTG));
          }
        }
      }
      if (total > 0) {
        tma_store_wait_read<0>();       // kernel completion covers the visibility of the writes
        mbar_arrive(bar_stg_empty((total - 1) % NSTG));
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue groups A / B
    const int group = (warp - 4) >> 2;
    const uint32_t q = warp & 3;
    const uint32_t row = q * 32 + lane;
    const uint32_t lane_addr = (q * 32) << 16;
    const float scale = p.scale;
    uint32_t df = 0;
    const uint32_t leader_h_full = mapa_u32(bar_h_full, 0);
    const uint32_t leader_d_empty = mapa_u32(bar_d_empty(group), 0);
    const int c_lo = group == 0 ? 0 : nA, c_hi = group == 0 ? nA : n16;
    if (!kBwd) {   // biases -> smem after the cluster sync (the cold reads overlap GEMM1); bu pre-scaled
      const int et = tid - 128;
      for (int i = et; i < R; i += 256) bias_smem[i] = p.bd[i];
      for (int i = et; i < kD; i += 256) bias_smem[R + i] = p.scale * p.bu[i];
      named_bar_sync(1, 256);
    }

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t tile_it = it;
      const int grow = tile_of(it) * BM + static_cast<int>(row);
      {
        // ---------------- epilogue 1: this group's half of P -> packed bf16 hidden in H
        uint4 hreg[8][2];         // kBwd: this row's saved hidden (issued before the wait on GEMM1)
        if constexpr (kBwd) {
          const uint4* hrow = reinterpret_cast<const uint4*>(p.H_in + static_cast<size_t>(grow) * R);
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
            hreg[ci][0] = hreg[ci][1] = make_uint4(0u, 0u, 0u, 0u);
            if (c_lo + ci < c_hi && grow < p.M) {
              hreg[ci][0] = __ldg(hrow + 2 * (c_lo + ci));
              hreg[ci][1] = __ldg(hrow + 2 * (c_lo + ci) + 1);
            }
          }
        }
        mbar_wait(bar_p_full, tile_it & 1);
        if (it > 0) mbar_wait(bar_h_free, (tile_it - 1) & 1);   // GEMM2 of the previous tile has read H
        tc_fence_after();
        if (tid == 128) FDP_TRACE(40, tile_it);
        const uint32_t t_p = tmem + lane_addr + TM_P;
        const uint32_t t_h = tmem + lane_addr + TM_H;
        uint32_t wall[8][8];      // the group's packed hidden / dP, kept for the (deferred) global store
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c = c_lo + ci;
          if (c < c_hi) {
            uint32_t v[16];
            tmem_ld16(t_p + c * 16, v);
            tmem_ld_wait16(v);
            if constexpr (!kBwd) {
              const float* bdv = bias_smem + c * 16;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                wall[ci][i] = pack_bf16x2(apply_act<kGelu>(__uint_as_float(v[2 * i]) + bdv[2 * i]),
                                          apply_act<kGelu>(__uint_as_float(v[2 * i + 1]) + bdv[2 * i + 1]));
            } else {
              const uint32_t hb[8] = {hreg[ci][0].x, hreg[ci][0].y, hreg[ci][0].z, hreg[ci][0].w,
                                      hreg[ci][1].x, hreg[ci][1].y, hreg[ci][1].z, hreg[ci][1].w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {      // relu'(P) == (H > 0); H is never negative
                const float g0 = (hb[i] & 0x00007fffu) ? scale * __uint_as_float(v[2 * i]) : 0.f;
                const float g1 = (hb[i] & 0x7fff0000u) ? scale * __uint_as_float(v[2 * i + 1]) : 0.f;
                wall[ci][i] = pack_bf16x2(g0, g1);
              }
            }
            tmem_st8(t_h + c * 8, wall[ci]);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_addr(leader_h_full);
        // global stores AFTER GEMM2 has been released (a row-per-thread store is 32 L1 transactions per
        // instruction): forward saves the hidden, backward the trainable slice of dP for the wgrad kernel
        if constexpr (!kBwd) {
          if (p.H_out != nullptr && grow < p.M) {
            uint4* hrow = reinterpret_cast<uint4*>(p.H_out + static_cast<size_t>(grow) * R);
#pragma unroll
            for (int ci = 0; ci < 8; ++ci)
              if (c_lo + ci < c_hi) {
                hrow[2 * (c_lo + ci)] = make_uint4(wall[ci][0], wall[ci][1], wall[ci][2], wall[ci][3]);
                hrow[2 * (c_lo + ci) + 1] = make_uint4(wall[ci][4], wall[ci][5], wall[ci][6], wall[ci][7]);
              }
          }
        } else {
          if (p.dP_t != nullptr && grow < p.M) {
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
              const int col = (c_lo + ci) * 16;
              if (c_lo + ci < c_hi && col >= p.r_lo && col < p.r_hi) {
                uint4* gd = reinterpret_cast<uint4*>(p.dP_t + static_cast<size_t>(grow) * p.ld_t + (col - p.r_lo));
                gd[0] = make_uint4(wall[ci][0], wall[ci][1], wall[ci][2], wall[ci][3]);
                gd[1] = make_uint4(wall[ci][4], wall[ci][5], wall[ci][6], wall[ci][7]);
              }
            }
          }
        }
        if (tid == 128) FDP_TRACE(41, tile_it);
      }
      // ---------------- epilogue 2: output chunks c == group (mod 2), 64 columns each
      for (int c = group; c < NC2; c += 2) {
        mbar_wait(bar_d_full(group), df & 1);
        ++df;
        tc_fence_after();
        const uint32_t g = tile_it * NC2 + c;
        const uint32_t sb = g % NSTG, rpar = (g / NSTG) & 1;
        const uint32_t t_src = tmem + lane_addr + TM_D + group * N2;
        uint32_t v0[32], v1[32];
        tmem_ld32(t_src, v0);
        tmem_ld32(t_src + 32, v1);
        tmem_ld_wait32(v0);
        tmem_ld_wait32(v1);
        // the accumulator is in registers: hand the D buffer back before anything else
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_addr(leader_d_empty);
        if (lane == 0 && q == 0 && c < 6) FDP_TRACE(42 + 4 * c, tile_it);
        mbar_wait(bar_res_full(sb), rpar);
        const uint32_t sbuf = stg_base + sb * SLOT;
        const float4* bu4 = reinterpret_cast<const float4*>(bias_smem + R + c * N2);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          uint4 rv[4];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            rv[i4] = make_uint4(0u, 0u, 0u, 0u);
            if (!kBwd || p.has_res) rv[i4] = ld_shared_v4(sbuf + sw128_offset(row, hb * 4 + i4));
          }
          uint32_t o[4][4];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const uint32_t rr[4] = {rv[i4].x, rv[i4].y, rv[i4].z, rv[i4].w};
            float sbv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if constexpr (!kBwd) {
              const float4 b0 = bu4[2 * (hb * 4 + i4)], b1 = bu4[2 * (hb * 4 + i4) + 1];
              sbv[0] = b0.x; sbv[1] = b0.y; sbv[2] = b0.z; sbv[3] = b0.w;
              sbv[4] = b1.x; sbv[5] = b1.y; sbv[6] = b1.z; sbv[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = i4 * 8 + 2 * i;
              const float a0 = __uint_as_float(hb == 0 ? v0[e] : v1[e]);
              const float a1 = __uint_as_float(hb == 0 ? v0[e + 1] : v1[e + 1]);
              // fwd: y = res + bf16(scale * (acc + bu)) (see dat_fused.cu); bwd: dX = (dY +) bf16(acc)
              const uint32_t t = kBwd ? pack_bf16x2(a0, a1)
                                      : pack_bf16x2(fmaf(scale, a0, sbv[2 * i]), fmaf(scale, a1, sbv[2 * i + 1]));
              o[i4][i] = hadd2_bf16(rr[i], t);
            }
          }
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
            st_shared_v4(sbuf + sw128_offset(row, hb * 4 + i4), o[i4][0], o[i4][1], o[i4][2], o[i4][3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_out_full(sb));
        if (lane == 0 && q == 0 && c < 6) FDP_TRACE(45 + 4 * c, tile_it);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
  if (tid == 64) FDP_TRACE(2, 0);
}


int launch_pipe(bool bwd, const void* A, const void* Res, void* Out, const void* W1, const void* W2, PipeParams p,
                int64_t M, int r_total, bool gelu, int grid, cudaStream_t st) {
  int rc;
  p.M = static_cast<int>(M);
  p.R = r_total;
  p.num_tiles = static_cast<int>((M + BM - 1) / BM);
  p.w2_3d = (r_total % 64 == 0) ? 1 : 0;
  p.trace = FD_TRACE_PTR;
  const size_t max_smem = 227 * 1024 - 1024;
  const size_t smem = 1024 + static_cast<size_t>(NG1) * G1STAGE + static_cast<size_t>(NW2 + NSTG) * SLOT +
                      (r_total + kD) * sizeof(float);
  FD_REQUIRE(smem <= max_smem, FD_ERR_UNSUPPORTED, "dat pipe kernel: shared-memory budget exceeded (R=%d)", r_total);

  CUtensorMap tmX, tmRes, tmY, tmWd, tmW2, tmW2k;
  if ((rc = make_tmap_bf16_2d(&tmX, A, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmRes, Res, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmY, Out, M, kD, kD, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWd, W1, r_total, kD, kD, r_total / 2, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW2, W2, kD, r_total, r_total, N2 / 2, 64))) return rc;
  tmW2k = tmW2;
  if (p.w2_3d && (rc = make_tmap_bf16_kblocks(&tmW2k, W2, kD, r_total, r_total, N2 / 2, r_total / 64)))
    return rc;

  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                            const CUtensorMap, const CUtensorMap, const PipeParams);
  KernelFn fn = bwd ? dat_pipe_kernel<true, false> : (gelu ? dat_pipe_kernel<false, true> : dat_pipe_kernel<false, false>);
  static bool configured[3][64] = {{false}};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  const int kidx = bwd ? 2 : (gelu ? 1 : 0);
  if (dev >= 64 || !configured[kidx][dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    if (dev < 64) configured[kidx][dev] = true;
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
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, tmX, tmRes, tmY, tmWd, tmW2, tmW2k, p));
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

}  // namespace

int launch_dat_fwd_pipe(const void* X, const void* Res, void* Y, const void* Wd_cat, const float* bd_cat,
                        const void* Wu_cat, const float* bu_cat, void* H_out, int64_t M, int r_total, float scale,
                        int act, int grid, cudaStream_t st) {
  PipeParams p{};
  p.scale = scale;
  p.bd = bd_cat;
  p.bu = bu_cat;
  p.H_out = static_cast<__nv_bfloat16*>(H_out);
  return launch_pipe(false, X, Res, Y, Wd_cat, Wu_cat, p, M, r_total, act == FEDDAT_ACT_GELU, grid, st);
}

int launch_dat_bwd_pipe(const void* dY, void* dX, const void* WuT_cat, const void* WdT_cat, const void* H_in,
                        void* dP_t, int ld_t, int r_lo, int r_hi, int64_t M, int r_total, float scale, int add_dy,
                        int grid, cudaStream_t st) {
  PipeParams p{};
  p.scale = scale;
  p.H_in = static_cast<const __nv_bfloat16*>(H_in);
  p.dP_t = static_cast<__nv_bfloat16*>(dP_t);
  p.ld_t = ld_t;
  p.r_lo = dP_t ? r_lo : 0;
  p.r_hi = dP_t ? r_hi : 0;
  p.has_res = add_dy ? 1 : 0;
  return launch_pipe(true, dY, dY, dX, WuT_cat, WdT_cat, p, M, r_total, false, grid, st);
}

}  // namespace fd
