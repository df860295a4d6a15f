// Backward of the short-sequence self-attention (attn.cu): dQ, dK, dV from dO, Q, K, V, O and the forward's logsumexp,
// for S <= 192 tokens and head dimension 64 (the ViLT shape: 40 text + 145 image tokens).
//
// Everything is computed TRANSPOSED -- keys on the TMEM lanes, queries along the columns:
//   S^T  = K_t Q^T                       [128 keys x QP queries]   (QP = S padded to 64)
//   P^T  = 2^(S^T c - lse2[q])           no row reductions: the forward's logsumexp is a per-COLUMN constant here
//   dP^T = V_t dO^T
//   dS^T = P^T (dP^T - delta[q])         delta[q] = sum_d dO[q, d] O[q, d]
//   dV_t = P^T dO,   dK_t = scale dS^T Q          A operand straight from TMEM (packed bf16 over the scores' columns)
//   dQ  += scale dS K_t                  the one product that reduces over the lanes: dS^T also goes to shared memory
//                                        (row = key, 64-query blocks, 128-byte swizzle) and is read as an M-major A
// so the cuDNN flash backward's three launches (dO . O pass, 128 x 128 tiles with fp32 dQ atomics, dQ conversion:
// 13 + 64 + 9 us at B = 64, S = 185) become one, with no intermediate in global memory.
//
// Work item = (batch, head), inner loop over the 128-key tiles; one persistent CTA per SM:
//   warps 0-15   "softmax": four threads per key row (warps w, w + 4, w + 8, w + 12 share a TMEM lane quarter and
//                split the query range)
//   warps 16-19  epilogue: delta / lse2 of the NEXT item, dV and dK of each key tile, dQ at the end of the item,
//                through one [32 x 64] staging slab per warp and TMA stores
//   warp 20      MMA issuer.  TMEM: X = [0, 192) S^T, then P^T and dS^T packed over each thread's own score columns;
//                Y = [192, 384) dP^T, then dV in [192, 256) and dK in [256, 320); dQ accumulators [384, 512).
//                S^T of the next key tile is issued right behind the MMAs that consume P^T / dS^T (in-order
//                execution makes that safe), dP^T once the epilogue warps have taken dV / dK out of Y.
//   warp 21      TMA producer: Q and dO of the item (two buffers);  warp 22: K and V tiles (two stages)
#include <stdlib.h>

#include "feddat_b200.h"
#include "attn_common.cuh"

namespace fd {
namespace {

constexpr int BTHREADS = 768;          // 16 softmax + 4 epilogue + MMA + 2 TMA producer + 1 statistics warp
constexpr uint32_t X_COL = 0, Y_COL = 192, DQ_COL = 384;
constexpr uint32_t KT_BYTES = AQ * 128;          // one [128 x 64] operand tile

struct AttnBwdTmaps {
  CUtensorMap q, d_o, k, v;      // loads: Q / dO [QP x 64] per item, K / V [128 x 64] per key tile
};
struct AttnBwdParams {
  int B, S, H;
  int n_kt;          // 128-key tiles per item
  int n_items;       // B * H
  float scale, scale_log2e;
  const float* stats; // [B * H, 2, 192] fp32 from attn_stats_kernel: lse * log2(e) (+inf past S) | delta (0 past S)
  __nv_bfloat16* dQ; // the gradients leave through plain, row-segment-coalesced stores
  __nv_bfloat16* dK;
  __nv_bfloat16* dV;
  int64_t lddq, lddk, lddv;
  unsigned long long* trace;   // debug twin only
};

#ifdef FEDDAT_DEBUG
#define AB_TRACE(ev, g)                                                                      \
  do {                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && (g) < 16) p.trace[(g) * 16 + (ev)] = globaltimer_ns(); \
  } while (0)
#else
#define AB_TRACE(ev, g) do { (void)(g); } while (0)
#endif

template <int kNC>   // QP / 64
__global__ void __launch_bounds__(BTHREADS, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdTmaps tm, const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[20];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float stats_s[384];           // [lse2 | delta] of the current item, from the statistics kernel
  float* const lse2_s = stats_s;
  float* const delta_s = stats_s + 192;

  constexpr int QP = kNC * 64;
  constexpr uint32_t QD_BYTES = static_cast<uint32_t>(QP) * 128u;     // Q or dO of one item
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto q_s = [&](uint32_t buf) { return smem0 + buf * 2 * QD_BYTES; };
  auto do_s = [&](uint32_t buf) { return smem0 + buf * 2 * QD_BYTES + QD_BYTES; };
  const uint32_t kv0 = smem0 + 4 * QD_BYTES;
  auto k_s = [&](uint32_t st) { return kv0 + st * 2 * KT_BYTES; };
  auto v_s = [&](uint32_t st) { return kv0 + st * 2 * KT_BYTES + KT_BYTES; };
  const uint32_t ds_s = kv0 + 4 * KT_BYTES;                          // dS^T: kNC blocks of [128 keys x 64 queries]
  const uint32_t slab0 = ds_s + kNC * KT_BYTES;                      // 4 staging slabs; also what an M tile past QP reads
  const uint32_t bar0 = smem_u32(bars);
  auto bar_qd_full = [&](uint32_t b) { return bar0 + 8 * b; };
  auto bar_qd_free = [&](uint32_t b) { return bar0 + 16 + 8 * b; };
  auto bar_kv_full = [&](uint32_t s) { return bar0 + 32 + 8 * s; };
  auto bar_kv_free = [&](uint32_t s) { return bar0 + 48 + 8 * s; };
  const uint32_t bar_st = bar0 + 64, bar_dp = bar0 + 72, bar_ds = bar0 + 88, bar_dvk = bar0 + 96,
                 bar_acc_free = bar0 + 104, bar_dq_free = bar0 + 112, bar_stats = bar0 + 120, bar_dq_done = bar0 + 128;

  if (tid == 0) {
    for (uint32_t i = 0; i < 2; ++i) {
      mbar_init(bar_qd_full(i), 1);
      mbar_init(bar_qd_free(i), 1);
      mbar_init(bar_kv_full(i), 1);
      mbar_init(bar_kv_free(i), 1);
    }
    mbar_init(bar_st, 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_ds, 16);
    mbar_init(bar_dvk, 1);
    mbar_init(bar_acc_free, 4);
    mbar_init(bar_dq_free, 4);
    mbar_init(bar_stats, 1);
    mbar_init(bar_dq_done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm.q);
    tma_prefetch_desc(&tm.d_o);
    tma_prefetch_desc(&tm.k);
    tma_prefetch_desc(&tm.v);
  }
  // dS^T rows of keys past the sequence are never written; they meet zero K rows in the dQ product and must be finite
  for (uint32_t off = tid * 16u; off < kNC * KT_BYTES; off += BTHREADS * 16u) st_shared_v4(ds_s + off, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 20) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem = tmem_base_smem;
  const int S = p.S, n_kt = p.n_kt;
  const int n_mine = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n_mtq = (S + AQ - 1) / AQ;               // 128-query tiles of dQ
  auto item_of = [&](int i) { return static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x); };

  if (warp == 20) {
    // ------------------------------------------------------------------ MMA issuer
    {
      // the whole warp runs the control flow (waits, address arithmetic) so that descriptors stay in uniform
      // registers; one elected lane issues
      const uint32_t idesc_t = make_idesc_bf16(AQ, QP);            // S^T, dP^T: both operands K-major
      const uint32_t idesc_v = make_idesc_bf16(AQ, AD, 0, 1);      // dV, dK: A in TMEM, B MN-major
      const uint32_t idesc_q = make_idesc_bf16(AQ, AD, 1, 1);      // dQ: A M-major (dS^T in smem), B MN-major
      // descriptors differ from tile to tile only in the start-address field (units of 16 bytes, no carry out of
      // it: every operand lies below 256 KB): one base per operand, then integer adds -- a single thread issues
      // ~50 MMAs per key tile and descriptor arithmetic was half of its time
      const uint64_t dk_q0 = desc_kmajor_sw128(q_s(0)), dk_q1 = desc_kmajor_sw128(q_s(1));
      const uint64_t dk_do0 = desc_kmajor_sw128(do_s(0)), dk_do1 = desc_kmajor_sw128(do_s(1));
      const uint64_t dm_q0 = desc_mnmajor_sw128(q_s(0), 1024), dm_q1 = desc_mnmajor_sw128(q_s(1), 1024);
      const uint64_t dm_do0 = desc_mnmajor_sw128(do_s(0), 1024), dm_do1 = desc_mnmajor_sw128(do_s(1), 1024);
      const uint64_t dk_k0 = desc_kmajor_sw128(k_s(0)), dk_k1 = desc_kmajor_sw128(k_s(1));
      const uint64_t dk_v0 = desc_kmajor_sw128(v_s(0)), dk_v1 = desc_kmajor_sw128(v_s(1));
      const uint64_t dm_k0 = desc_mnmajor_sw128(k_s(0), 1024), dm_k1 = desc_mnmajor_sw128(k_s(1), 1024);
      const uint64_t dm_ds = desc_mnmajor_sw128(ds_s, KT_BYTES);
      uint32_t g = 0;
      auto issue_st = [&](int i, int kt, uint32_t gg) {
        const uint32_t buf = static_cast<uint32_t>(i) & 1, st = gg & 1;
        if (kt == 0) mbar_wait(bar_qd_full(buf), (static_cast<uint32_t>(i) >> 1) & 1);
        mbar_wait(bar_kv_full(st), (gg >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AD / 16; ++k) umma_ss(tmem + X_COL, (st ? dk_k1 : dk_k0) + 2 * k, (buf ? dk_q1 : dk_q0) + 2 * k, idesc_t, k > 0);
          umma_commit(bar_st);
          AB_TRACE(0, gg);
        }
        __syncwarp();
      };
      if (n_mine > 0) issue_st(0, 0, 0);
      for (int i = 0; i < n_mine; ++i) {
        const uint32_t buf = static_cast<uint32_t>(i) & 1;
        for (int kt = 0; kt < n_kt; ++kt, ++g) {
          const uint32_t st = g & 1;
          // dP^T = V_t dO^T into Y once the epilogue warps have taken the previous tile's dV / dK out of it
          if (g > 0) mbar_wait(bar_acc_free, (g - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AD / 16; ++k) umma_ss(tmem + Y_COL, (st ? dk_v1 : dk_v0) + 2 * k, (buf ? dk_do1 : dk_do0) + 2 * k, idesc_t, k > 0);
            umma_commit(bar_dp);
            AB_TRACE(1, g);
          }
          __syncwarp();
          // P^T and dS^T are in TMEM (X), dS^T also in shared memory; dP^T has been read out of Y
          mbar_wait(bar_ds, g & 1);
          tc_fence_after();
          if (elect_one()) {
          AB_TRACE(2, g);
#pragma unroll
          for (int k = 0; k < QP / 16; ++k)        // dV_t = P^T dO  (16 query rows = 2048 bytes = 128 address units per step)
            umma_ts(tmem + Y_COL, tmem + X_COL + (k / kNC) * (QP / 4) + (k % kNC) * 8, (buf ? dm_do1 : dm_do0) + 128 * k, idesc_v, k > 0);
#pragma unroll
          for (int k = 0; k < QP / 16; ++k)        // dK_t = dS^T Q
            umma_ts(tmem + Y_COL + 64, tmem + X_COL + (k / kNC) * (QP / 4) + QP / 8 + (k % kNC) * 8, (buf ? dm_q1 : dm_q0) + 128 * k, idesc_v,
                    k > 0);
          umma_commit(bar_dvk);
          }
          __syncwarp();
          // S^T of the next key tile overwrites X right behind its two readers (tcgen05.mma executes in order)
          if (kt + 1 < n_kt) issue_st(i, kt + 1, g + 1);
          else if (i + 1 < n_mine) issue_st(i + 1, 0, g + 1);
          if (kt == 0 && i > 0) mbar_wait(bar_dq_free, (static_cast<uint32_t>(i) - 1) & 1);   // previous item's dQ is out
          tc_fence_after();
          if (elect_one()) {
          for (int mt = 0; mt < n_mtq; ++mt)       // dQ[128 mt ...] += dS K_t  (M = queries: two 64-query blocks per tile)
#pragma unroll
            for (int k = 0; k < AQ / 16; ++k)
              umma_ss(tmem + DQ_COL + mt * 64, dm_ds + (mt * 2 * KT_BYTES >> 4) + 128 * k, (st ? dm_k1 : dm_k0) + 128 * k, idesc_q,
                      !(kt == 0 && k == 0));
          umma_commit(bar_kv_free(st));
          if (kt == n_kt - 1) {
            umma_commit(bar_qd_free(buf));
            umma_commit(bar_dq_done);
          }
          AB_TRACE(3, g);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 21) {
    // ------------------------------------------------------------------ TMA producer: Q and dO of the item
    if (lane == 0) {
      for (int i = 0; i < n_mine; ++i) {
        const int bh = item_of(i), h = bh % p.H, b = bh / p.H;
        const uint32_t buf = static_cast<uint32_t>(i) & 1;
        if (i >= 2) mbar_wait(bar_qd_free(buf), ((static_cast<uint32_t>(i) >> 1) - 1) & 1);
        mbar_arrive_expect_tx(bar_qd_full(buf), 2 * QD_BYTES);
        tma_load_3d(q_s(buf), &tm.q, bar_qd_full(buf), h * AD, 0, b);
        tma_load_3d(do_s(buf), &tm.d_o, bar_qd_full(buf), h * AD, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 22) {
    // ------------------------------------------------------------------ TMA producer: K and V tiles
    if (lane == 0) {
      uint32_t g = 0;
      for (int i = 0; i < n_mine; ++i) {
        const int bh = item_of(i), h = bh % p.H, b = bh / p.H;
        for (int kt = 0; kt < n_kt; ++kt, ++g) {
          const uint32_t st = g & 1;
          if (g >= 2) mbar_wait(bar_kv_free(st), ((g >> 1) - 1) & 1);
          mbar_arrive_expect_tx(bar_kv_full(st), 2 * KT_BYTES);
          tma_load_3d(k_s(st), &tm.k, bar_kv_full(st), h * AD, kt * AQ, b);
          tma_load_3d(v_s(st), &tm.v, bar_kv_full(st), h * AD, kt * AQ, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 23) {
    // ------------------------------------------------------------------ statistics: one 1.5 KB bulk copy per item, issued as
    // soon as the softmax warps are through with the previous item's (its last dS^T is out)
    // (every phase of bar_ds is waited for in turn: a parity wait cannot tell phase g from phase g - 2)
    if (lane == 0) {
      uint32_t g = 0;
      for (int i = 0; i < n_mine; ++i) {
        mbar_arrive_expect_tx(bar_stats, 384 * 4);
        bulk_load_1d(smem_u32(stats_s), p.stats + static_cast<size_t>(item_of(i)) * 384, 384 * 4, bar_stats);
        for (int kt = 0; kt < n_kt; ++kt, ++g) mbar_wait(bar_ds, g & 1);
      }
    }
    __syncwarp();
  } else if (warp >= 16) {
    // ------------------------------------------------------------------ epilogue warps
    const uint32_t q = warp & 3;
    const uint32_t lane_addr = (q * 32u) << 16;
    const uint32_t slab = slab0 + q * 4096u;
    // one [32 x 64] fp32 accumulator block of this lane quarter -> x mul -> bf16 -> the warp's slab (row per thread)
    // -> global memory as full 128-byte row segments, eight lanes per row (four lines per store instruction; a row
    // per lane would occupy the load / store unit 8 x longer, which the softmax warps' shared-memory traffic feels)
    auto store_block = [&](uint32_t col, float mul, __nv_bfloat16* base, int64_t ld, int row0, int h, int b, bool arrive_acc) {
      const float2 m2 = make_float2(mul, mul);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + col + hf * 32, v);
        tmem_ld_wait32(v);
        if (hf == 1 && arrive_acc) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_free);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 u = __fmul2_rn(make_float2(__uint_as_float(v[ch * 8 + 2 * k]), __uint_as_float(v[ch * 8 + 2 * k + 1])), m2);
            o[k] = pack_bf16x2(u.x, u.y);
          }
          st_shared_v4(slab + sw128_offset(lane, hf * 4 + ch), o[0], o[1], o[2], o[3]);
        }
      }
      __syncwarp();
      const uint32_t j8 = lane & 7, r0 = lane >> 3;
      __nv_bfloat16* dst = base + (static_cast<size_t>(b) * S + row0 + r0) * ld + h * AD + j8 * 8;
#pragma unroll
      for (int i2 = 0; i2 < 8; ++i2) {
        const uint4 t = ld_shared_v4(slab + sw128_offset(r0 + 4 * i2, j8));
        if (row0 + static_cast<int>(r0) + 4 * i2 < S) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(4 * i2) * ld) = t;
      }
      __syncwarp();
    };

    uint32_t g = 0;
    for (int i = 0; i < n_mine; ++i) {
      const int bh = item_of(i), h = bh % p.H, b = bh / p.H;
      for (int kt = 0; kt < n_kt; ++kt, ++g) {
        mbar_wait(bar_dvk, g & 1);
        tc_fence_after();
        if (tid == 512) AB_TRACE(9, g);
        const int row0 = kt * AQ + static_cast<int>(q) * 32;
        if (row0 < S) {
          store_block(Y_COL, 1.f, p.dV, p.lddv, row0, h, b, false);
          store_block(Y_COL + 64, p.scale, p.dK, p.lddk, row0, h, b, true);
          if (tid == 512) AB_TRACE(10, g);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_free);
        }
        if (kt == n_kt - 1) {
          mbar_wait(bar_dq_done, i & 1);
          tc_fence_after();
          for (int mt = 0; mt < n_mtq; ++mt)
            if (mt * AQ + static_cast<int>(q) * 32 < S)
              store_block(DQ_COL + mt * 64, p.scale, p.dQ, p.lddq, mt * AQ + static_cast<int>(q) * 32, h, b, false);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dq_free);
          if (tid == 512) AB_TRACE(11, g);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ "softmax" warps: four threads per key row
    const uint32_t q = warp & 3, cq = warp >> 2;
    const uint32_t row = q * 32 + lane;                        // key row of the tile == TMEM lane
    const uint32_t lane_addr = (q * 32u) << 16;
    const float c = p.scale_log2e;
    constexpr int NQ = QP / 4;                                 // query columns per thread
    const int j0 = static_cast<int>(cq) * NQ;
    uint32_t g = 0;
    for (int i = 0; i < n_mine; ++i) {
      mbar_wait(bar_stats, i & 1);
      for (int kt = 0; kt < n_kt; ++kt, ++g) {
        const bool live = kt * AQ + static_cast<int>(q) * 32 < S;     // warp-uniform: some key of this lane quarter is real
        mbar_wait(bar_st, g & 1);
        tc_fence_after();
        if (tid == 0) AB_TRACE(5, g);
        uint32_t pk[NQ / 2];                                   // this thread's P^T, bf16 pairs
        if (live) {
          uint32_t v[kNC][16];
#pragma unroll
          for (int ch = 0; ch < kNC; ++ch) tmem_ld16(tmem + lane_addr + X_COL + j0 + ch * 16, v[ch]);
#pragma unroll
          for (int ch = 0; ch < kNC; ++ch) {
            tmem_ld_wait16(v[ch]);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 l = *reinterpret_cast<const float4*>(&lse2_s[j0 + ch * 16 + k4 * 4]);
              const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(v[ch][k4 * 4]), __uint_as_float(v[ch][k4 * 4 + 1])),
                                           make_float2(c, c), make_float2(-l.x, -l.y));
              const float2 a1 = __ffma2_rn(make_float2(__uint_as_float(v[ch][k4 * 4 + 2]), __uint_as_float(v[ch][k4 * 4 + 3])),
                                           make_float2(c, c), make_float2(-l.z, -l.w));
              pk[ch * 8 + k4 * 2] = pack_bf16x2(fast_ex2(a0.x), fast_ex2(a0.y));
              pk[ch * 8 + k4 * 2 + 1] = pack_bf16x2(fast_ex2(a1.x), fast_ex2(a1.y));
            }
          }
        }
        // P^T goes over the FIRST half of this thread's own score columns, dS^T (below) over the second half: no
        // thread writes a column another one still has to read, so no barrier; the MMAs take one TMEM address per
        // 16-query step and do not care that the steps are not contiguous
        if (live) {
#pragma unroll
          for (int ch = 0; ch < kNC; ++ch) {
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = pk[ch * 8 + k];
            tmem_st8(tmem + lane_addr + X_COL + j0 + ch * 8, w);
          }
        }
        if (tid == 0) AB_TRACE(6, g);
        mbar_wait(bar_dp, g & 1);
        tc_fence_after();
        if (tid == 0) AB_TRACE(7, g);
        if (live) {
          uint32_t u[2][16];                                   // the next chunk is requested before this one is used
          tmem_ld16(tmem + lane_addr + Y_COL + j0, u[0]);
#pragma unroll
          for (int ch = 0; ch < kNC; ++ch) {
            uint32_t w[8];
            tmem_ld_wait16(u[ch & 1]);
            if (ch + 1 < kNC) tmem_ld16(tmem + lane_addr + Y_COL + j0 + (ch + 1) * 16, u[(ch + 1) & 1]);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 d = *reinterpret_cast<const float4*>(&delta_s[j0 + ch * 16 + k4 * 4]);
              const float2 p0 = unpack_bf16x2(pk[ch * 8 + k4 * 2]), p1 = unpack_bf16x2(pk[ch * 8 + k4 * 2 + 1]);
              const float2 t0 = __fadd2_rn(make_float2(__uint_as_float(u[ch & 1][k4 * 4]), __uint_as_float(u[ch & 1][k4 * 4 + 1])), make_float2(-d.x, -d.y));
              const float2 t1 = __fadd2_rn(make_float2(__uint_as_float(u[ch & 1][k4 * 4 + 2]), __uint_as_float(u[ch & 1][k4 * 4 + 3])), make_float2(-d.z, -d.w));
              const float2 s0 = __fmul2_rn(p0, t0), s1 = __fmul2_rn(p1, t1);
              w[k4 * 2] = pack_bf16x2(s0.x, s0.y);
              w[k4 * 2 + 1] = pack_bf16x2(s1.x, s1.y);
            }
            // dS^T: packed over the second half of this thread's own score columns ...
            tmem_st8(tmem + lane_addr + X_COL + j0 + NQ / 2 + ch * 8, w);
            // ... and into the [keys x queries] shared-memory tile the dQ product reads
            const int qc = j0 + ch * 16;                       // first of these 16 query columns
            const uint32_t blk = ds_s + static_cast<uint32_t>(qc >> 6) * KT_BYTES;
            st_shared_v4(blk + sw128_offset(row, (qc & 63) >> 3), w[0], w[1], w[2], w[3]);
            st_shared_v4(blk + sw128_offset(row, ((qc & 63) >> 3) + 1), w[4], w[5], w[6], w[7]);
          }
          tmem_st_wait();
          fence_proxy_async_smem();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ds);
        if (tid == 0) AB_TRACE(8, g);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 20) tmem_dealloc(tmem, 512);
}

// delta[q] = dO[q, :] . O[q, :] and lse2[q] = lse[q] log2(e) per (batch, head), padded to 192 entries with (+inf, 0)
// so that padded queries get P^T = 0 and dS^T = 0.  One block per (batch, 8 tokens): it reads the tokens' whole rows of
// O and dO (all heads: H x 128 contiguous bytes per token -- a per-head walk touches 128 bytes every 1.5 KB and runs at
// a third of the HBM rate), eight lanes per (token, head), three shuffles.
__global__ void __launch_bounds__(256)
attn_stats_kernel(const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO, const float* __restrict__ lse,
                  float* __restrict__ stats, int S, int H, int64_t ldo, int64_t lddo) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y, t0 = blockIdx.x * 8;
  const int j = threadIdx.x & 7;
  for (int slot = threadIdx.x >> 3; slot < 8 * H; slot += 32) {      // slot = (token of the chunk, head)
    const int r = t0 + slot / H, h = slot % H;
    float acc = 0.f;
    if (r < S) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(O + (static_cast<size_t>(b) * S + r) * ldo + h * AD) + j);
      const uint4 d = __ldg(reinterpret_cast<const uint4*>(dO + (static_cast<size_t>(b) * S + r) * lddo + h * AD) + j);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc = fmaf(__uint_as_float(aw[k] << 16), __uint_as_float(dw[k] << 16), acc);
        acc = fmaf(__uint_as_float(aw[k] & 0xffff0000u), __uint_as_float(dw[k] & 0xffff0000u), acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (j == 0 && r < 192) {
      const size_t bh = static_cast<size_t>(b) * H + h;
      stats[bh * 384 + r] = r < S ? __ldg(lse + bh * S + r) * 1.4426950408889634f : INFINITY;
      stats[bh * 384 + 192 + r] = acc;
    }
  }
}

}  // namespace
}  // namespace fd

extern "C" size_t feddat_attn_bwd_workspace_bytes(int B, int H) { return static_cast<size_t>(B) * H * 384 * sizeof(float); }

extern "C" int feddat_attn_bwd(const void* dO, const void* Q, const void* K, const void* V, const void* O, const void* LSE,
                               void* dQ, void* dK, void* dV, int B, int S, int H, int D, int64_t lddo, int64_t ldq,
                               int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddq, int64_t lddk, int64_t lddv, float scale,
                               void* workspace, size_t ws_bytes, int dtype, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED, "attn_bwd: only bf16 is implemented (dtype=%d)", dtype);
  FD_REQUIRE(dO && Q && K && V && O && LSE && dQ && dK && dV, FD_ERR_INVALID, "attn_bwd: null pointer argument");
  FD_REQUIRE(D == AD && S >= 1 && S <= 192 && H >= 1 && B >= 0, FD_ERR_UNSUPPORTED,
             "attn_bwd: head dimension %d / sequence length %d outside the short-sequence kernel (D = 64, S <= 192)", D, S);
  FD_REQUIRE(ldo % 8 == 0 && lddo % 8 == 0 && ((reinterpret_cast<uintptr_t>(O) | reinterpret_cast<uintptr_t>(dO)) & 15) == 0,
             FD_ERR_INVALID, "attn_bwd: O / dO rows must be 16-byte aligned");
  FD_REQUIRE(workspace != nullptr && ws_bytes >= feddat_attn_bwd_workspace_bytes(B, H) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
             FD_ERR_INVALID, "attn_bwd: workspace of feddat_attn_bwd_workspace_bytes(B, H) bytes, 16-byte aligned, required");
  if (B == 0) return FD_OK;
  auto st = static_cast<cudaStream_t>(stream);
  const char* e = getenv("FEDDAT_PDL");
  const int n_attr = (e && e[0] == '0') ? 0 : 1;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  {
    cudaLaunchConfig_t c0{};
    c0.gridDim = dim3(192 / 8, B);
    c0.blockDim = dim3(256);
    c0.stream = st;
    c0.attrs = attr;
    c0.numAttrs = n_attr;
    FD_CHECK_CUDA(cudaLaunchKernelEx(&c0, attn_stats_kernel, static_cast<const __nv_bfloat16*>(O),
                                     static_cast<const __nv_bfloat16*>(dO), static_cast<const float*>(LSE),
                                     static_cast<float*>(workspace), S, H, ldo, lddo));
  }
  AttnBwdParams p{};
  p.B = B; p.S = S; p.H = H;
  p.n_kt = (S + AQ - 1) / AQ;
  p.n_items = B * H;
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.stats = static_cast<const float*>(workspace);
  p.dQ = static_cast<__nv_bfloat16*>(dQ);
  p.dK = static_cast<__nv_bfloat16*>(dK);
  p.dV = static_cast<__nv_bfloat16*>(dV);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.trace = FD_TRACE_PTR;
  const int nc = (S + 63) / 64, QP = nc * 64;
  AttnBwdTmaps tm;
  if ((rc = make_tmap_tokens(&tm.q, Q, B, S, H * D, ldq, QP))) return rc;
  if ((rc = make_tmap_tokens(&tm.d_o, dO, B, S, H * D, lddo, QP))) return rc;
  if ((rc = make_tmap_tokens(&tm.k, K, B, S, H * D, ldk, AQ))) return rc;
  if ((rc = make_tmap_tokens(&tm.v, V, B, S, H * D, ldv, AQ))) return rc;
  FD_REQUIRE(lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 &&
                 ((reinterpret_cast<uintptr_t>(dQ) | reinterpret_cast<uintptr_t>(dK) | reinterpret_cast<uintptr_t>(dV)) & 15) == 0,
             FD_ERR_INVALID, "attn_bwd: dQ / dK / dV rows must be 16-byte aligned");
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  // Q / dO double-buffered, K / V two stages, dS^T, four slabs (+ one operand tile of slack behind dS^T for the
  // M tile that reaches past QP)
  const size_t smem = 1024 + 4 * static_cast<size_t>(QP) * 128 + 4 * KT_BYTES + static_cast<size_t>(nc) * KT_BYTES + KT_BYTES;
  using KernelFn = void (*)(const AttnBwdTmaps, const AttnBwdParams);
  KernelFn fn = nc == 1 ? attn_bwd_kernel<1> : nc == 2 ? attn_bwd_kernel<2> : attn_bwd_kernel<3>;
  static bool configured[4][64] = {{false}};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[nc][dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev < 64) configured[nc][dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.n_items < sms ? p.n_items : sms);
  cfg.blockDim = dim3(BTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, tm, p));
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
