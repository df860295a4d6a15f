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
// ingests only half the weight bytes.  12 warps per CTA, every hand-off an mbarrier:
//
//   ring        NS = 4 stages of 32 KB.  GEMM1 stage kc = [X k-chunk 128 x 64 | Wd_cat half k-chunk
//               R/2 x 64]; GEMM2 stage c = this CTA's half [64 x R] of the Wu_cat tile of output chunk c
//               (one 3-D TMA box when R % 64 == 0).  Producers run ahead across tiles.
//   warp 0      ring producers: lane 0 issues the activation TMA (X, and dY in backward), lane 1 the
//               weight TMA of every GEMM1 stage and the GEMM2 tiles, in one converged loop
//               (a single issuing thread sustains one TMA per ~170 ns, scripts/ingest_probe.py: the
//               v3 kernel's one-thread / 16 KB-slot ring capped the SM at 62 GB/s of a possible 130+)
//   warp 1      tcgen05.mma issuer (leader CTA of the pair only; M = 256 across the pair).  GEMM1
//               P = X Wd^T (SS, K-major SW128 smem operands) into TMEM columns [0, R); GEMM2 in six
//               128-column chunks, A operand = the hidden tile IN TMEM (bf16 pairs packed over P's own
//               columns, "TS" MMA), accumulators in a two-buffer ring in TMEM columns [256, 512)
//   warp 3      residual producer: TMA loads of the residual [128 x 64] chunks into the staging ring
//   warp 2      store issuer: TMA stores of finished staging buffers, recycles them
//   warps 4-7   epilogue group A, warps 8-11 group B.  (1) each group packs HALF of P's columns ->
//               +bias, act -> bf16 pairs -> tcgen05.st over its own half of P (the hidden never leaves
//               the SM); (2) group A drains even output chunks, group B odd ones: tcgen05.ld, scale,
//               +bias, +residual (from the staging buffer), bf16 back into the same staging buffer.
//               One elected lane per warp signals each barrier.
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
//
// GROUPED launches (feddat_dat_fwd_grouped / feddat_dat_bwd_dgrad_grouped): up to two independent row
// groups, each with its own activations, weights, bottleneck width and scale, share ONE launch -- the MKD
// schedule's gating rows (adapter_0 | adapter_2, R = 2r, scale .5) and adapter_1 rows (R = r, scale 1) of
// one adapter site (task_trainer.py:280-330).  Every CTA pair walks "slots": slot -> (group, super-tile,
// output-column range); all per-group quantities (R, descriptors, biases, tensor maps) are looked up per
// slot.  A single site of one batch (2 x 47 tiles) thus fills 144 of the 148 SMs in one launch instead of
// two launches of 144 CTAs that each recompute the hidden three times.
#include <stdlib.h>

#include "dat_kernels.h"
#include "feddat_b200.h"
#ifdef FEDDAT_DEBUG
#include "feddat_b200_debug.h"
#endif
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {

namespace {

constexpr int kD = 768;
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int KC1 = kD / BK;          // 12 k-chunks for GEMM1
constexpr int SLOT = BM * 128;        // 16 KB: [128 rows x 64 bf16], 128-byte swizzle
constexpr int STAGE = 2 * SLOT;       // 32 KB ring stage
constexpr int N2 = 128;               // GEMM2 / GEMM3 output chunk width
constexpr int NC2 = kD / N2;          // 6 chunks
constexpr int NS = 4;                 // ring stages (power of two)
constexpr int NSTG = 5;               // 16 KB residual-in / output staging buffers
constexpr int NUM_THREADS = 384;
constexpr uint32_t W2_KB_BYTES = (N2 / 2) * 128u;  // one k-block [64 rows x 64] of a half W2 tile
constexpr uint32_t TM_P = 0;          // TMEM column of P (and of the packed hidden aliasing it)
constexpr uint32_t TM_D = 256;        // TMEM column of the output ring (and of dH in backward)

constexpr int MAX_GROUPS = 2;
constexpr int BIAS_STRIDE = 256 + kD;   // floats of bias staging per group: bd [256] | scale * bu [768]

struct GroupCfg {
  int M, R, num_tiles, w2_3d;
  int n_split;            // S: CTA pairs per super-tile, each owning NC2 / S of the output chunks (small M)
  float scale;
  const float* bd;
  const float* bu;        // fwd only
  // bwd only
  int has_out;            // 0: dX not requested (skip GEMM3 / epilogue 2)
  int has_res;            // add the residual tile (fwd: always; bwd: add_dy)
  int r_lo, r_hi;         // trainable slice of the bottleneck
  __nv_bfloat16* H_t;     // bwd, recompute mode: hidden of the trainable slice (row stride ld_t) or null
  __nv_bfloat16* dP_t;    // bwd: pre-activation gradient of the trainable slice (row stride ld_t) or null
  int ld_t;               // row stride (elements) of H_t / dP_t
  const __nv_bfloat16* H_in;   // bwd, saved mode: the forward's hidden [M, R] (no P recompute)
  __nv_bfloat16* H_out;        // fwd: where to save the hidden [M, R] for such a backward, or null
};

struct FusedParams {
  GroupCfg g[MAX_GROUPS];
  int n_groups;
  int slots0;             // slots of group 0 (= its super-tiles x its column split); group 1 follows
  int total_slots;
  int act;
  unsigned long long* trace;  // debug: globaltimer stamps of CTA 0's pipeline events (or null)
};

// the nine tensor maps of one group
struct TmapSet {
  CUtensorMap x, res, y, wd, w2, w2k, w1b, h, dp;
};

// Epilogue 1 splits the R / 16 sixteen-column chunks of P between the two epilogue groups: group A packs
// chunks [0, nA), group B the rest.  nA is rounded up to a multiple of four (64 columns) so that both groups'
// shares start at a 64-column boundary: the packed hidden also travels through the 64-column staging buffers
// (saved by the forward / loaded and turned into dP by the backward with TMA, see "hidden chunks" below).
__host__ __device__ __forceinline__ int split_a(int n16) {
  const int a = (((n16 + 1) / 2) + 3) & ~3;
  return a < n16 ? a : n16;
}

// slot -> what this CTA pair computes
struct Slot {
  int g;        // group
  int tile0;    // first 128-row tile of the super-tile (this CTA: tile0 + cluster rank)
  int c_base;   // first 128-column output chunk of this pair
  int nc2;      // output chunks of this pair
};

// debug timeline (scripts/trace_kernel.py): event e of CTA 0's tile `t` (t < 2) -> trace[t * 128 + e]
#ifdef FEDDAT_DEBUG
#define FD_TRACE(ev, t)                                                             \
  do {                                                                              \
    if (p.trace != nullptr && blockIdx.x == 0 && (t) < 2)                           \
      p.trace[(t) * 128 + (ev)] = globaltimer_ns();                                 \
  } while (0)
// the same for BOTH CTAs of pair 0: event ev of CTA rank c -> trace[t * 128 + ev + 8 c]
#define FD_TRACE_PAIR(ev, t)                                                        \
  do {                                                                              \
    if (p.trace != nullptr && blockIdx.x < 2 && (t) < 2)                            \
      p.trace[(t) * 128 + (ev) + 8 * blockIdx.x] = globaltimer_ns();                \
  } while (0)
// entry / exit stamps of EVERY CTA (launch ramp and tail skew): trace[256 + 2 * blockIdx.x + which]
#define FD_TRACE_CTA(which)                                                         \
  do {                                                                              \
    if (p.trace != nullptr && blockIdx.x < 384) p.trace[256 + 2 * blockIdx.x + (which)] = globaltimer_ns(); \
  } while (0)
#else   // product build: no trace code in the kernels
#define FD_TRACE(ev, t) do { (void)(t); } while (0)
#define FD_TRACE_PAIR(ev, t) do { (void)(t); } while (0)
#define FD_TRACE_CTA(which) do { } while (0)
#endif

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

// Tensor maps (per group): forward                      backward (kBwd)
//   x     [M, 768]       X                            X
//   res   [M, 768]       residual input               dY  (GEMM1b A operand and the optional +dY)
//   y     [M, 768]       Y                            dX
//   wd    [R, 768]       Wd_cat                       Wd_cat
//   w2    [768, R]       Wu_cat                       WdT_cat      (2-D, one [64 x 64] k-block per box)
//   w2k   [768, R]       the same tensor as a 3-D (64, 768, R/64) view: box = all k-blocks of 64 rows
//   w1b   [R, 768]       (unused)                     WuT_cat
//   h     [M, R]         H_out (hidden to save)       H_in (saved hidden)      box [128 x 64]
//   dp    [M, r_t]       (unused)                     dP_t (trainable slice, row stride ld_t)
// HIDDEN CHUNKS.  The staging ring (NSTG buffers of [128 rows x 64 columns]) carries, per tile, first the
// ceil(R / 64) chunks of the hidden and then the output chunks.  Forward with H_out: epilogue 1 writes the packed
// hidden into the buffers next to its tcgen05.st and the store issuer TMA-stores them.  Saved-hidden backward:
// the residual producer TMA-loads the H_in chunks, epilogue 1 reads relu' from them, writes dP back IN PLACE
// and the store issuer stores the chunks of the trainable slice.  (Row-per-thread global loads / stores of
// these tiles are 32 sectors per instruction: 4 096 LSU transactions per CTA that sat in the SM's in-order
// memory pipeline in front of every mbarrier operation of the producer and MMA warps -- 2.6 us between
// "hidden packed" and the first GEMM2 MMA in the forward, 2.9 us on GEMM1b in the backward.)
// kSaved (backward, ReLU only): the hidden saved by the forward replaces the recompute of P -- no X
// read, no GEMM1 pass 0; relu'(P) is read off the saved hidden (H > 0).
template <bool kBwd, bool kGelu, bool kSaved = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dat_fused_kernel(const __grid_constant__ TmapSet tm0, const __grid_constant__ TmapSet tm1,
                 const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  // barriers: stage full/empty, P full, dH full, hidden full, D full/empty x2, staging res/out/empty
  __shared__ __align__(8) uint64_t bars[2 * NS + 3 + 4 + 3 * NSTG];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) FD_TRACE(1, 0);      // kernel entry (before barrier init / TMEM alloc / bias staging)
  if (tid == 0) FD_TRACE_CTA(0);
  const uint32_t rank = cluster_ctarank();            // 0 = leader of the CTA pair

  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem0 + NS * STAGE;
  const uint32_t bias_base = stg_base + NSTG * SLOT;
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));

  const uint32_t bar0 = smem_u32(bars);
  auto bar_slot_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto bar_slot_empty = [&](uint32_t s) { return bar0 + 8u * (NS + s); };
  const uint32_t bar_p_full = bar0 + 8u * (2 * NS);
  const uint32_t bar_g_full = bar0 + 8u * (2 * NS + 1);
  const uint32_t bar_h_full = bar0 + 8u * (2 * NS + 2);
  auto bar_d_full = [&](int b) { return bar0 + 8u * (2 * NS + 3 + b); };
  auto bar_d_empty = [&](int b) { return bar0 + 8u * (2 * NS + 5 + b); };
  auto bar_res_full = [&](uint32_t b) { return bar0 + 8u * (2 * NS + 7 + b); };
  auto bar_out_full = [&](uint32_t b) { return bar0 + 8u * (2 * NS + 7 + NSTG + b); };
  auto bar_stg_empty = [&](uint32_t b) { return bar0 + 8u * (2 * NS + 7 + 2 * NSTG + b); };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_slot_full(s), 1);
      mbar_init(bar_slot_empty(s), 1);
    }
    mbar_init(bar_p_full, 1);
    mbar_init(bar_g_full, 1);
    mbar_init(bar_h_full, 16);           // one lane of each of the 8 epilogue warps, both CTAs
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d_full(b), 1);
      mbar_init(bar_d_empty(b), 8);      // the 4 warps of the group that drains buffer b, both CTAs
    }
    for (int b = 0; b < NSTG; ++b) {
      mbar_init(bar_res_full(b), 1);
      mbar_init(bar_out_full(b), 4);     // the 4 warps of one epilogue group
      mbar_init(bar_stg_empty(b), 1);
    }
    fence_mbar_init();
    for (int g = 0; g < p.n_groups; ++g) {
      const TmapSet& T = g ? tm1 : tm0;
      tma_prefetch_desc(&T.x);
      tma_prefetch_desc(&T.res);
      tma_prefetch_desc(&T.y);
      tma_prefetch_desc(&T.wd);
      tma_prefetch_desc(&T.w2);
      tma_prefetch_desc(&T.w2k);
      if (kBwd) tma_prefetch_desc(&T.w1b);
      if (!kBwd || kSaved) tma_prefetch_desc(&T.h);
      if (kSaved) tma_prefetch_desc(&T.dp);
    }
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible to both CTAs of the pair
  tc_fence_after();
  // Everything above overlapped the tail of the previous kernel in the stream (programmatic dependent
  // launch); from here on this kernel reads what that kernel wrote.
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) FD_TRACE(0, 0);

  // Slots.  Group g contributes (its super-tiles) x (its column split S_g) slots; group 1's slots follow
  // group 0's.  With few tiles (a single adapter site of one batch: 47 tiles per group for 148 SMs) the
  // launch has one CTA pair per slot and S_g pairs share a super-tile: each recomputes the whole hidden
  // (GEMM1 + epilogue 1; the SMs would idle otherwise, the extra X / W reads hit L2) and produces its own
  // NC2 / S_g output chunks [c_base, c_base + nc2).  With many tiles S_g = 1 and the pairs stride over the
  // slots.  A CTA's tile may lie beyond the tensor (odd tile count): TMA then zero-fills the loads and
  // clips the stores, so no role needs a special case.
  const int pid = blockIdx.x >> 1, n_pairs_launched = static_cast<int>(gridDim.x >> 1);
  auto decode = [&](int slot) -> Slot {
    Slot s;
    s.g = slot >= p.slots0 ? 1 : 0;
    const int rel = slot - (s.g ? p.slots0 : 0);
    const GroupCfg& G = p.g[s.g];
    const int np = (G.num_tiles + 1) / 2;
    s.nc2 = (kBwd && !G.has_out) ? 0 : NC2 / G.n_split;
    s.c_base = (rel / np) * s.nc2;
    s.tile0 = 2 * (rel % np);
    return s;
  };
  // hidden chunks of a slot in the staging sequence: total, and group A's share
  auto hidden_chunks = [&](const Slot& sl, int& nhA) -> int {
    const GroupCfg& G = p.g[sl.g];
    const bool on = kBwd ? kSaved : (G.H_out != nullptr && sl.c_base == 0);
    const int n16 = G.R / 16, nA = split_a(n16);
    nhA = on ? (nA + 3) / 4 : 0;
    return on ? nhA + (n16 - nA + 3) / 4 : 0;
  };
  // every TMA of the pair credits its bytes to the LEADER's stage barrier
  const uint32_t leader_full0 = mapa_u32(bar_slot_full(0), 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ ring producers
    // lane 0 = activations (X, dY), lane 1 = weights: the SAME loop, so the warp stays converged;
    // each lane issues its own TMA (a lone thread sustains only one TMA per ~170 ns)
    if (lane < 2) {
      uint32_t n = 0;  // ring stages consumed so far (all roles count the same sequence)
      uint32_t tile_it = 0;
      for (int slot = pid; slot < p.total_slots; slot += n_pairs_launched, ++tile_it) {
        const Slot sl = decode(slot);
        const GroupCfg& G = p.g[sl.g];
        const TmapSet& T = sl.g ? tm1 : tm0;
        const int R = G.R, RH = R / 2, KC2 = (R + 63) / 64;
        const uint32_t w_half_bytes = static_cast<uint32_t>(RH) * 128u;
        const int m0 = (sl.tile0 + static_cast<int>(rank)) * BM;
        if (lane == 0) FD_TRACE(110, tile_it);
        for (int pass = kSaved ? 1 : 0; pass < (kBwd ? 2 : 1); ++pass) {
          const CUtensorMap* tm = lane == 0 ? (pass == 0 ? &T.x : &T.res) : (pass == 0 ? &T.wd : &T.w1b);
          const int c1 = lane == 0 ? m0 : static_cast<int>(rank) * RH;
          const uint64_t pol = kEvictLast;   // activations too: they are re-read as the residual
          for (int kc = 0; kc < KC1; ++kc, ++n) {
            const uint32_t s = n & (NS - 1), par = (n / NS) & 1;
            mbar_wait(bar_slot_empty(s), par ^ 1);
            // the leader's "full" barrier collects the bytes of BOTH CTAs' loads (X and W) for a
            // stage.  Lane 1's bytes may land before lane 0's expect_tx: the phase cannot complete
            // before that arrival, and a transiently negative tx-count is legal.
            if (lane == 0 && rank == 0) mbar_arrive_expect_tx(bar_slot_full(s), 2 * (SLOT + w_half_bytes));
            tma_load_2d_pair(smem0 + s * STAGE + lane * SLOT, tm, leader_full0 + 8u * s, kc * BK, c1, pol);
          }
        }
        if (lane == 0) FD_TRACE(111, tile_it);
        // Next slot's activations -> L2 now, so that its GEMM1 streams from L2 instead of waiting on
        // HBM in lock-step with every other SM
        if (lane == 0 && slot + n_pairs_launched < p.total_slots) {
          const Slot nx = decode(slot + n_pairs_launched);
          const TmapSet& TN = nx.g ? tm1 : tm0;
          const int m1 = (nx.tile0 + static_cast<int>(rank)) * BM;
          for (int kc = 0; kc < KC1; ++kc) {
            if (!kSaved) tma_prefetch_l2_2d(&TN.x, kc * BK, m1);
            if (kBwd) tma_prefetch_l2_2d(&TN.res, kc * BK, m1);
          }
        }
        // GEMM2 stages: this CTA's half [64 x R] of the W2 tile of output chunk c (lane 1); lane 0
        // waits along (a parity wait is only meaningful within one lap of the ring)
        for (int c = 0; c < sl.nc2; ++c, ++n) {
          const uint32_t s = n & (NS - 1), par = (n / NS) & 1;
          mbar_wait(bar_slot_empty(s), par ^ 1);
          if (lane == 1) {
            if (rank == 0) mbar_arrive_expect_tx(bar_slot_full(s), 2 * KC2 * W2_KB_BYTES);
            const int row0 = (sl.c_base + c) * N2 + static_cast<int>(rank) * (N2 / 2);
            if (G.w2_3d) {   // all k-blocks of the half tile in one box
              tma_load_3d_pair(smem0 + s * STAGE, &T.w2k, leader_full0 + 8u * s, 0, row0, 0, kEvictLast);
            } else {
              for (int kb = 0; kb < KC2; ++kb)
                tma_load_2d_pair(smem0 + s * STAGE + kb * W2_KB_BYTES, &T.w2, leader_full0 + 8u * s,
                                 kb * BK, row0, kEvictLast);
            }
          }
        }
        if (lane == 1) FD_TRACE(112, tile_it);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The WHOLE warp runs the loops (uniform control flow keeps the descriptors in uniform registers);
    // one elected lane issues.  Under `if (lane == 0)` every tcgen05.mma cost ~15 extra vector ->
    // uniform register moves, ~70 ns per instruction against 33 ns of tensor time (N = 128).
    if (rank == 0) {
      uint32_t n = 0;
      uint32_t de[2] = {0, 0};  // uses of each D buffer so far (parity of its "empty" barrier)
      const uint32_t idesc2 = make_idesc_bf16(2 * BM, N2);
      uint32_t tile_it = 0;
      for (int slot = pid; slot < p.total_slots; slot += n_pairs_launched, ++tile_it) {
        const Slot sl = decode(slot);
        const int R = p.g[sl.g].R;
        const int n16 = R / 16, nA = split_a(n16);
        const uint32_t idesc1 = make_idesc_bf16(2 * BM, R);
        for (int pass = kSaved ? 1 : 0; pass < (kBwd ? 2 : 1); ++pass) {
          // pass 0: P = X Wd_cat^T -> [TM_P, +R);  pass 1 (bwd): dH = dY Wu_cat -> [TM_D, +R)
          const uint32_t d_tmem = tmem + (pass == 0 ? TM_P : TM_D);
          if (pass == 1) {
            // dH overlays the output ring: the previous tile's last chunks must have been drained.
            // No new "use" is registered: dH's own readers are covered by bar_h_full below.
            for (int b = 0; b < 2; ++b) mbar_wait(bar_d_empty(b), (de[b] & 1) ^ 1);
            tc_fence_after();
          }
          for (int kc = 0; kc < KC1; ++kc, ++n) {
            const uint32_t s = n & (NS - 1), par = (n / NS) & 1;
            mbar_wait(bar_slot_full(s), par);
            tc_fence_after();
            if (elect_one()) {
              if (pass == 0) FD_TRACE(10 + kc, tile_it);
              // descriptors differ only in the start-address field (bits [0,14) of addr >> 4, no
              // carry below 256 KB): one build per stage, +2 per 32-byte k-step
              const uint64_t adesc = desc_kmajor_sw128(smem0 + s * STAGE);
              const uint64_t bdesc = adesc + (SLOT >> 4);
              umma_ss_pair(d_tmem, adesc, bdesc, idesc1, kc != 0);
#pragma unroll
              for (int k = 1; k < 4; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, 1);
              umma_commit_pair(bar_slot_empty(s), 0b11);
              if (kc == KC1 - 1) {
                umma_commit_pair(pass == 0 ? bar_p_full : bar_g_full, 0b11);
                FD_TRACE(22 + pass, tile_it);
              }
            }
            __syncwarp();
          }
        }
        // epilogue 1 done in BOTH CTAs: the packed hidden (dP) is in TMEM and P may be overwritten
        // by the next tile's GEMM1 (waited even when no GEMM2/3 follows)
        mbar_wait(bar_h_full, tile_it & 1);
        tc_fence_after();
        if (lane == 0) FD_TRACE(24, tile_it);
        for (int c = 0; c < sl.nc2; ++c, ++n) {
          const int b = c & 1;
          mbar_wait(bar_d_empty(b), (de[b] & 1) ^ 1);
          ++de[b];
          const uint32_t s = n & (NS - 1), par = (n / NS) & 1;
          mbar_wait(bar_slot_full(s), par);
          tc_fence_after();
          if (elect_one()) {
            FD_TRACE(25 + c, tile_it);
            const uint32_t d_tmem = tmem + TM_D + b * N2;
            // fully unrolled and predicated on the (uniform) bottleneck width
            const uint64_t bdesc = desc_kmajor_sw128(smem0 + s * STAGE);
            const uint32_t a0 = tmem + TM_P, a_hi = static_cast<uint32_t>(8 * nA);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
              if (kk < n16)
                umma_ts_pair(d_tmem, a0 + 8 * kk + (kk >= nA ? a_hi : 0u),
                             bdesc + (((kk >> 2) * W2_KB_BYTES + (kk & 3) * 32) >> 4), idesc2,
                             kk != 0 ? 1u : 0u);
            }
            umma_commit_pair(bar_slot_empty(s), 0b11);
            umma_commit_pair(bar_d_full(b), 0b11);
            FD_TRACE(31 + c, tile_it);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ residual producer
    if (lane == 0) {
      uint32_t g = 0;
      uint32_t tile_it = 0;
      for (int slot = pid; slot < p.total_slots; slot += n_pairs_launched, ++tile_it) {
        const Slot sl = decode(slot);
        const TmapSet& T = sl.g ? tm1 : tm0;
        const bool has_res = p.g[sl.g].has_res != 0;
        const int m0 = (sl.tile0 + static_cast<int>(rank)) * BM;
        const uint32_t per_tile = sl.nc2 * 2;
        int nhA;
        const int nh = hidden_chunks(sl, nhA);
        for (int j = 0; j < nh; ++j, ++g) {      // hidden chunks: loaded (saved backward) or just granted (forward)
          const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
          mbar_wait(bar_stg_empty(sb), par ^ 1);
          if constexpr (kSaved) {
            mbar_arrive_expect_tx(bar_res_full(sb), SLOT);
            tma_load_2d(stg_base + sb * SLOT, &T.h, bar_res_full(sb), j * 64, m0);
          } else {
            mbar_arrive(bar_res_full(sb));
          }
        }
        for (uint32_t c64 = 0; c64 < per_tile; ++c64, ++g) {
          const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
          mbar_wait(bar_stg_empty(sb), par ^ 1);
          if (has_res) {
            mbar_arrive_expect_tx(bar_res_full(sb), SLOT);
            tma_load_2d(stg_base + sb * SLOT, &T.res, bar_res_full(sb), (sl.c_base * 2 + c64) * 64, m0);
            FD_TRACE(90 + c64, tile_it);
          } else {
            mbar_arrive(bar_res_full(sb));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ store issuer
    if (lane == 0) {
      uint32_t g = 0;
      uint32_t tile_it = 0;
      for (int slot = pid; slot < p.total_slots; slot += n_pairs_launched, ++tile_it) {
        const Slot sl = decode(slot);
        const TmapSet& T = sl.g ? tm1 : tm0;
        const int m0 = (sl.tile0 + static_cast<int>(rank)) * BM;
        const uint32_t per_tile = sl.nc2 * 2;
        int nhA;
        const int nh = hidden_chunks(sl, nhA);
        const GroupCfg& G = p.g[sl.g];
        for (int j = 0; j < nh; ++j, ++g) {      // hidden chunks: H_out (forward) / dP_t of the trainable slice
          const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
          mbar_wait(bar_out_full(sb), par);
          if constexpr (!kBwd) {
            tma_store_2d_hint(&T.h, stg_base + sb * SLOT, j * 64, m0, kEvictFirst);
          } else {
            if (G.dP_t != nullptr && sl.c_base == 0 && j * 64 < G.r_hi && j * 64 + 64 > G.r_lo)
              tma_store_2d_hint(&T.dp, stg_base + sb * SLOT, j * 64 - G.r_lo, m0, kEvictLast);
          }
          tma_store_commit();                    // (an empty group when nothing was stored: same bookkeeping)
          if (g > 0) {
            tma_store_wait_read<1>();
            mbar_arrive(bar_stg_empty((g - 1) % NSTG));
          }
        }
        for (uint32_t c64 = 0; c64 < per_tile; ++c64, ++g) {
          const uint32_t sb = g % NSTG, par = (g / NSTG) & 1;
          mbar_wait(bar_out_full(sb), par);
          // outputs are never re-read by this kernel: let them leave L2 first
          tma_store_2d_hint(&T.y, stg_base + sb * SLOT, (sl.c_base * 2 + c64) * 64, m0, kEvictFirst);
          tma_store_commit();
          FD_TRACE(104 + (c64 >> 1), tile_it);
          if (g > 0) {  // the previous store has finished reading its buffer: recycle it
            tma_store_wait_read<1>();
            mbar_arrive(bar_stg_empty((g - 1) % NSTG));
          }
        }
      }
      if (g > 0) {
        // smem has been read; kernel completion covers the visibility of the writes themselves
        tma_store_wait_read<0>();
        mbar_arrive(bar_stg_empty((g - 1) % NSTG));
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue groups A / B
    const int group = (warp - 4) >> 2;      // 0: warps 4-7, 1: warps 8-11
    const uint32_t q = warp & 3;            // TMEM lane quarter this warp may touch
    const uint32_t row = q * 32 + lane;     // tile row == TMEM lane
    const uint32_t lane_addr = (q * 32) << 16;
    uint32_t df = 0;  // chunk fills of this group's D buffer so far
    // the MMA issuer lives in the leader CTA: "hidden ready" / "D buffer drained" go to ITS barriers
    const uint32_t leader_h_full = mapa_u32(bar_h_full, 0);
    const uint32_t leader_d_empty = mapa_u32(bar_d_empty(group), 0);
    // Biases -> smem by the 256 epilogue threads, AFTER the cluster sync: the (cold) global reads overlap
    // GEMM1 instead of delaying every role.  Per group: bd as is at [0, R); bu PRE-SCALED by the branch
    // scale at [256, 256 + 768) (epilogue 2 is one FFMA per element).
    {
      const int et = tid - 128;
      for (int g = 0; g < p.n_groups; ++g) {
        const GroupCfg& G = p.g[g];
        float* bs = bias_smem + g * BIAS_STRIDE;
        if (!kSaved)
          for (int i = et; i < G.R; i += 256) bs[i] = G.bd[i];
        if (!kBwd)
          for (int i = et; i < kD; i += 256) bs[256 + i] = G.scale * G.bu[i];
      }
      named_bar_sync(1, 256);
    }

    uint32_t stg_n = 0;      // 64-column staging chunks of all previous slots
    uint32_t tile_it = 0;
    for (int slot = pid; slot < p.total_slots; slot += n_pairs_launched, ++tile_it) {
      const Slot sl = decode(slot);
      const GroupCfg& G = p.g[sl.g];
      const int R = G.R, nc2 = sl.nc2, c_base = sl.c_base;
      // epilogue 1 is split between the two epilogue groups by 16-column chunks of P: group A packs
      // chunks [0, nA), group B chunks [nA, n16).  Each packs IN PLACE over its own columns, so the
      // hidden's k-step k (16 bottleneck units = 8 packed columns) sits at column 8 k (k < nA) or
      // 16 nA + 8 (k - nA)
      const int n16 = R / 16, nA = split_a(n16);
      const int c_lo = group == 0 ? 0 : nA, c_hi = group == 0 ? nA : n16;
      const uint32_t w_base = group == 0 ? 0u : static_cast<uint32_t>(16 * nA);   // where its hidden goes
      const float scale = G.scale;
      const bool has_res = G.has_res != 0;
      const float* bias_g = bias_smem + sl.g * BIAS_STRIDE;
      const int m0 = (sl.tile0 + static_cast<int>(rank)) * BM;
      int nhA;
      const int nh = hidden_chunks(sl, nhA);
      {
        // ---------------- epilogue 1: this group's half of P (and dH) -> packed bf16 hidden
        const int grow = m0 + static_cast<int>(row);
        const uint32_t hg0 = stg_n + (group == 0 ? 0u : static_cast<uint32_t>(nhA));   // its first hidden chunk
        if (!kSaved) mbar_wait(bar_p_full, tile_it & 1);
        if (kBwd) mbar_wait(bar_g_full, tile_it & 1);
        tc_fence_after();
        if (tid == 128) FD_TRACE(40, tile_it);
        const uint32_t t_p = tmem + lane_addr + TM_P;
        const uint32_t t_g = tmem + lane_addr + TM_D;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c = c_lo + ci;
          if (c < c_hi) {
            uint32_t v[16], u[16], w[8];
            // the 64-column staging buffer this 16-column chunk belongs to (forward with H_out: to fill; saved
            // backward: holding the saved hidden)
            const uint32_t hgi = hg0 + (ci >> 2);
            const uint32_t hbuf = stg_base + (hgi % NSTG) * SLOT;
            if (nh > 0 && (ci & 3) == 0) mbar_wait(bar_res_full(hgi % NSTG), (hgi / NSTG) & 1);
            if (!kSaved) tmem_ld16(t_p + c * 16, v);
            if (kBwd) tmem_ld16(t_g + c * 16, u);
            const float* bdv = bias_g + c * 16;
            uint4 h0 = make_uint4(0u, 0u, 0u, 0u), h1 = h0;
            if constexpr (kSaved) {
              h0 = ld_shared_v4(hbuf + sw128_offset(row, 2 * (ci & 3)));
              h1 = ld_shared_v4(hbuf + sw128_offset(row, 2 * (ci & 3) + 1));
            }
            if (!kSaved) tmem_ld_wait16(v);
            if (kBwd) tmem_ld_wait16(u);
            const int col = c * 16;
            if constexpr (!kBwd) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                w[i] = pack_bf16x2(apply_act<kGelu>(__uint_as_float(v[2 * i]) + bdv[2 * i]),
                                   apply_act<kGelu>(__uint_as_float(v[2 * i + 1]) + bdv[2 * i + 1]));
            } else if constexpr (kSaved) {
              const uint32_t hb[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {      // relu'(P) == (H > 0); H is never negative
                const float g0 = (hb[i] & 0x00007fffu) ? scale * __uint_as_float(u[2 * i]) : 0.f;
                const float g1 = (hb[i] & 0x7fff0000u) ? scale * __uint_as_float(u[2 * i + 1]) : 0.f;
                w[i] = pack_bf16x2(g0, g1);
              }
            } else {
              uint32_t hh[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float p0 = __uint_as_float(v[2 * i]) + bdv[2 * i];
                const float p1 = __uint_as_float(v[2 * i + 1]) + bdv[2 * i + 1];
                w[i] = pack_bf16x2(scale * __uint_as_float(u[2 * i]) * act_grad<kGelu>(p0),
                                   scale * __uint_as_float(u[2 * i + 1]) * act_grad<kGelu>(p1));
                hh[i] = pack_bf16x2(apply_act<kGelu>(p0), apply_act<kGelu>(p1));
              }
              if (G.H_t != nullptr && c_base == 0 && col >= G.r_lo && col < G.r_hi && grow < G.M) {
                const size_t off = static_cast<size_t>(grow) * G.ld_t + (col - G.r_lo);
                uint4* hd = reinterpret_cast<uint4*>(G.H_t + off);
                uint4* gd = reinterpret_cast<uint4*>(G.dP_t + off);
                hd[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                hd[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
                gd[0] = make_uint4(w[0], w[1], w[2], w[3]);
                gd[1] = make_uint4(w[4], w[5], w[6], w[7]);
              }
            }
            // the target columns lie inside a P chunk of THIS group that it has already read (or, in
            // saved mode, in the unused P region)
            tmem_st8(t_p + w_base + (c - c_lo) * 8, w);
            if (nh > 0) {
              // the same 16 columns into the staging buffer (forward: the hidden to save; saved backward: dP over
              // the hidden it was derived from -- each thread rewrites only what it has read itself)
              st_shared_v4(hbuf + sw128_offset(row, 2 * (ci & 3)), w[0], w[1], w[2], w[3]);
              st_shared_v4(hbuf + sw128_offset(row, 2 * (ci & 3) + 1), w[4], w[5], w[6], w[7]);
              if ((ci & 3) == 3 || c == c_hi - 1) {     // the buffer's last chunk of this group: hand it to the store issuer
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_out_full(hgi % NSTG));
              }
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_addr(leader_h_full);
        if (lane == 0) FD_TRACE_PAIR(70 + (warp - 4), tile_it);     // this warp's half-quarter of the hidden is packed
        if (tid == 128) FD_TRACE(41, tile_it);
      }
      // ---------------- epilogue 2: this group's output chunks (c = group, group + 2, group + 4)
      for (int c = group; c < nc2; c += 2) {
        const int b = c & 1;   // == group
        mbar_wait(bar_d_full(b), df & 1);
        ++df;
        tc_fence_after();
        if (lane == 0 && q == 0) FD_TRACE(42 + 4 * c, tile_it);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const uint32_t g = stg_n + nh + c * 2 + j;
          const uint32_t sb = g % NSTG, rpar = (g / NSTG) & 1;
          const int col0 = (c_base + c) * N2 + j * 64;
          const uint32_t t_src = tmem + lane_addr + TM_D + b * N2 + j * 64;
          uint32_t v0[32], v1[32];
          const bool tr = (tid == 128) && c == 0 && j == 0;
          if (tr) FD_TRACE(120, tile_it);
          tmem_ld32(t_src, v0);
          tmem_ld32(t_src + 32, v1);
          if (tr) FD_TRACE(121, tile_it);
          mbar_wait(bar_res_full(sb), rpar);
          if (tr) FD_TRACE(122, tile_it);
          tmem_ld_wait32(v0);
          tmem_ld_wait32(v1);
          if (j == 1) {
            // the chunk's whole accumulator has been read: hand the D buffer back BEFORE the last
            // half's math, so the MMAs of chunk c + 2 run under it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_addr(leader_d_empty);
          }
          if (tr) FD_TRACE(123, tile_it);
          if (lane == 0 && q == 0) FD_TRACE(43 + 4 * c + j, tile_it);
          const uint32_t sbuf = stg_base + sb * SLOT;
          const float4* bu4 = reinterpret_cast<const float4*>(bias_g + 256 + col0);
          // Two batches of 32 columns: the four residual ld.shared of a batch are issued back to back,
          // then the math (bias loads included) is free code for the scheduler, then the four
          // st.shared.  (asm volatile keeps program order: interleaving load / math / store per 16-byte
          // chunk exposed the full ld.shared latency eight times per half chunk -- 950 clk measured.)
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            uint4 rv[4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              rv[i4] = make_uint4(0u, 0u, 0u, 0u);
              if (has_res) rv[i4] = ld_shared_v4(sbuf + sw128_offset(row, hb * 4 + i4));
            }
            uint32_t o[4][4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const uint32_t rr[4] = {rv[i4].x, rv[i4].y, rv[i4].z, rv[i4].w};
              // y = res + bf16(scale * (acc + bu)): the adapter output is rounded to bf16 and added to
              // the bf16 residual with ONE packed HADD2 per column pair -- the arithmetic of the
              // reference under autocast (adapter.py:129-131: `up` is a half tensor, `residual + up`
              // a half add), and 4 instructions per pair instead of 9 with an fp32 residual add
              // (epilogue 2 is ALU-throughput bound)
              float sbv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              if constexpr (!kBwd) {
                const float4 b0 = bu4[2 * (hb * 4 + i4)], b1 = bu4[2 * (hb * 4 + i4) + 1];
                sbv[0] = b0.x; sbv[1] = b0.y; sbv[2] = b0.z; sbv[3] = b0.w;
                sbv[4] = b1.x; sbv[5] = b1.y; sbv[6] = b1.z; sbv[7] = b1.w;
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = i4 * 8 + 2 * i;          // column inside this 32-column batch
                const float a0 = __uint_as_float(hb == 0 ? v0[e] : v1[e]);
                const float a1 = __uint_as_float(hb == 0 ? v0[e + 1] : v1[e + 1]);
                uint32_t t;
                if constexpr (kBwd)
                  t = pack_bf16x2(a0, a1);
                else
                  t = pack_bf16x2(fmaf(scale, a0, sbv[2 * i]), fmaf(scale, a1, sbv[2 * i + 1]));
                o[i4][i] = hadd2_bf16(rr[i], t);
              }
            }
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4)
              st_shared_v4(sbuf + sw128_offset(row, hb * 4 + i4), o[i4][0], o[i4][1], o[i4][2], o[i4][3]);
          }
          if (tr) FD_TRACE(124, tile_it);
          fence_proxy_async_smem();
          if (tr) FD_TRACE(125, tile_it);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_out_full(sb));
          if (tr) FD_TRACE(126, tile_it);
        }
        if (lane == 0 && q == 0) FD_TRACE(45 + 4 * c, tile_it);
      }
      stg_n += static_cast<uint32_t>(nh + nc2 * 2);
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the partner's smem / barriers / TMEM stay valid until both CTAs are done
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
  if (tid == 64) FD_TRACE(2, 0);     // CTA exit
  if (tid == 64) FD_TRACE_CTA(1);
}


#ifdef FEDDAT_DEBUG
bool g_force_fused = false;       // A/B switch of the debug build (feddat_debug_force_fused_fwd)
#else
constexpr bool g_force_fused = false;
#endif

// Programmatic dependent launch is on unless FEDDAT_PDL=0 (read once; an A/B switch for measurements).
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FEDDAT_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// One group's host-side description (the C ABI's FeddatDatGroup, already validated).
struct HostGroup {
  const void *X, *Res, *W1, *W2, *W1b;   // W1: Wd_cat (fwd / recompute) or WuT_cat (saved); W2: Wu_cat / WdT_cat
  void* Out;
  GroupCfg cfg;
};

// cost model of the column split: a pair recomputes GEMM1 (~R) and produces 1/S of GEMM2 (~R / S)
int pick_splits(const HostGroup* hg, int n_groups, int sms, bool allow_split, int* splits) {
  int best = -1, best_cost = 0, best_ctas = 0;
  const int smax = allow_split ? 3 : 1;
  for (int s0 = 1; s0 <= smax; ++s0)
    for (int s1 = 1; s1 <= (n_groups > 1 ? smax : 1); ++s1) {
      const int s[2] = {s0, s1};
      int ctas = 0, cost = 0;
      for (int g = 0; g < n_groups; ++g) {
        const int sg = hg[g].cfg.has_out ? s[g] : 1;
        if (!hg[g].cfg.has_out && s[g] != 1) { ctas = 1 << 30; break; }
        ctas += 2 * ((hg[g].cfg.num_tiles + 1) / 2) * sg;
        const int c = hg[g].cfg.R * (sg + 1) / sg;
        cost = c > cost ? c : cost;
      }
      if (ctas > sms) continue;
      if (best < 0 || cost < best_cost || (cost == best_cost && ctas < best_ctas)) {
        best = s0 * 4 + s1; best_cost = cost; best_ctas = ctas;
        splits[0] = s0; splits[1] = s1;
      }
    }
  return best;
}

int launch_fused(bool bwd, bool saved, int act, HostGroup* hg, int n_groups, cudaStream_t st, const char* who) {
  int rc;
  FusedParams p{};
  p.n_groups = n_groups;
  p.act = act;
  p.trace = FD_TRACE_PTR;
  const size_t max_smem = 227 * 1024 - 1024;  // static smem (barriers) lives in the same budget
  const size_t smem = 1024 + static_cast<size_t>(NS) * STAGE + static_cast<size_t>(NSTG) * SLOT +
                      MAX_GROUPS * BIAS_STRIDE * sizeof(float);
  static_assert(1024 + NS * STAGE + NSTG * SLOT + MAX_GROUPS * BIAS_STRIDE * 4 <= 227 * 1024 - 1024, "smem budget");

  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  int total_pairs = 0;
  for (int g = 0; g < n_groups; ++g) total_pairs += (hg[g].cfg.num_tiles + 1) / 2;
  int splits[2] = {1, 1};
  const bool one_wave = 2 * total_pairs <= sms;
  if (one_wave) pick_splits(hg, n_groups, sms, true, splits);
  FD_REQUIRE(one_wave || n_groups == 1, FD_ERR_UNSUPPORTED,
             "%s: a grouped launch covers at most %d row tiles in total (the caller launches large groups one by one)",
             who, sms);

  TmapSet tms[MAX_GROUPS];
  int slots[MAX_GROUPS] = {0, 0};
  for (int g = 0; g < n_groups; ++g) {
    GroupCfg& c = hg[g].cfg;
    c.n_split = splits[g];
    c.w2_3d = (c.R % 64 == 0) ? 1 : 0;
    slots[g] = ((c.num_tiles + 1) / 2) * c.n_split;
    const int64_t M = c.M;
    const int R = c.R;
    TmapSet& T = tms[g];
    if ((rc = make_tmap_bf16_2d(&T.x, hg[g].X, M, kD, kD, BM, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&T.res, hg[g].Res, M, kD, kD, BM, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&T.y, hg[g].Out ? hg[g].Out : hg[g].X, M, kD, kD, BM, 64))) return rc;
    const uint32_t w_box_rows = R / 2;   // each CTA of a pair holds half of every weight tile
    if ((rc = make_tmap_bf16_2d(&T.wd, hg[g].W1, R, kD, kD, w_box_rows, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&T.w2, hg[g].W2, kD, R, R, N2 / 2, 64))) return rc;
    T.w2k = T.w2;
    if (c.w2_3d && (rc = make_tmap_bf16_kblocks(&T.w2k, hg[g].W2, kD, R, R, N2 / 2, R / 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&T.w1b, hg[g].W1b ? hg[g].W1b : hg[g].W1, R, kD, kD, w_box_rows, 64))) return rc;
    // hidden chunks through the staging ring: H_out (forward), H_in and the dP_t slice (saved backward)
    const void* hid = bwd ? static_cast<const void*>(c.H_in) : static_cast<const void*>(c.H_out);
    T.h = T.x;
    T.dp = T.x;
    if (hid != nullptr && (rc = make_tmap_bf16_2d(&T.h, hid, M, R, R, BM, 64))) return rc;
    if (saved && c.dP_t != nullptr &&
        (rc = make_tmap_bf16_2d(&T.dp, c.dP_t, M, c.r_hi - c.r_lo, c.ld_t, BM, 64))) return rc;
    p.g[g] = c;
  }
  if (n_groups == 1) {
    tms[1] = tms[0];
    p.g[1] = p.g[0];
  }
  p.slots0 = slots[0];
  p.total_slots = slots[0] + slots[1];
  const int pairs_launched = one_wave ? p.total_slots : (p.total_slots < sms / 2 ? p.total_slots : sms / 2);
  const int grid = 2 * pairs_launched;

  using KernelFn = void (*)(const TmapSet, const TmapSet, const FusedParams);
  const bool gelu = act == FEDDAT_ACT_GELU;
  KernelFn fn = saved ? dat_fused_kernel<true, false, true>
                : bwd ? (gelu ? dat_fused_kernel<true, true> : dat_fused_kernel<true, false>)
                      : (gelu ? dat_fused_kernel<false, true> : dat_fused_kernel<false, false>);
  // idempotent per-device "attribute already set" cache (not state the results depend on)
  static bool configured[3][2][64] = {{{false}}};
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  const int kidx = saved ? 2 : (bwd ? 1 : 0);
  if (dev >= 64 || !configured[kidx][gelu][dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)max_smem));
    if (dev < 64) configured[kidx][gelu][dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, tms[0], tms[1], p));
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

int tiles_of(int64_t M) { return static_cast<int>((M + BM - 1) / BM); }

// forward group -> HostGroup
int fwd_group(const FeddatDatGroup& G, int d, int act, int dtype, HostGroup* out) {
  FD_REQUIRE(G.X && G.Res && G.Y && G.Wd_cat && G.bd_cat && G.Wu_cat && G.bu_cat, FD_ERR_INVALID,
             "dat_fwd: null pointer argument");
  int rc = check_common("dat_fwd", G.M, d, G.r_total, act, dtype);
  if (rc) return rc;
  FD_REQUIRE((reinterpret_cast<uintptr_t>(G.H_out) & 15) == 0, FD_ERR_INVALID, "dat_fwd: H_out must be 16-byte aligned");
  HostGroup h{};
  h.X = G.X; h.Res = G.Res; h.Out = G.Y; h.W1 = G.Wd_cat; h.W2 = G.Wu_cat; h.W1b = nullptr;
  h.cfg.M = static_cast<int>(G.M); h.cfg.R = G.r_total; h.cfg.num_tiles = tiles_of(G.M);
  h.cfg.scale = G.branch_scale; h.cfg.bd = G.bd_cat; h.cfg.bu = G.bu_cat;
  h.cfg.has_out = 1; h.cfg.has_res = 1;
  h.cfg.H_out = static_cast<__nv_bfloat16*>(G.H_out);
  *out = h;
  return FD_OK;
}

// backward (dgrad) group -> HostGroup; *saved = saved-hidden mode
int bwd_group(const FeddatDatGroup& G, int d, int act, int dtype, HostGroup* out, bool* saved) {
  FD_REQUIRE(G.dY && G.WuT_cat && G.WdT_cat, FD_ERR_INVALID, "dat_bwd_dgrad: null pointer argument");
  int rc = check_common("dat_bwd_dgrad", G.M, d, G.r_total, act, dtype);
  if (rc) return rc;
  if (G.H_in != nullptr) {
    FD_REQUIRE(act == FEDDAT_ACT_RELU, FD_ERR_UNSUPPORTED,
               "dat_bwd_dgrad: a saved hidden determines act' only for ReLU; pass H_in = NULL (recompute) for GELU");
    FD_REQUIRE(G.H_t == nullptr, FD_ERR_INVALID, "dat_bwd_dgrad: H_t is not produced in saved mode (H_in holds it)");
    FD_REQUIRE((reinterpret_cast<uintptr_t>(G.H_in) & 15) == 0, FD_ERR_INVALID, "dat_bwd_dgrad: H_in must be 16-byte aligned");
    FD_REQUIRE(G.dX != nullptr || G.dP_t != nullptr, FD_ERR_INVALID, "dat_bwd_dgrad: nothing to compute");
  } else {
    FD_REQUIRE(G.X && G.Wd_cat && G.bd_cat, FD_ERR_INVALID, "dat_bwd_dgrad: X, Wd_cat, bd_cat are needed to recompute P");
    FD_REQUIRE((G.H_t == nullptr) == (G.dP_t == nullptr), FD_ERR_INVALID,
               "dat_bwd_dgrad: H_t and dP_t must both be given or both be NULL");
    FD_REQUIRE(G.dX != nullptr || G.H_t != nullptr, FD_ERR_INVALID,
               "dat_bwd_dgrad: nothing to compute (dX and H_t are both NULL)");
  }
  if (G.dP_t) {
    FD_REQUIRE(G.r_lo >= 0 && G.r_hi > G.r_lo && G.r_hi <= G.r_total && G.r_lo % 16 == 0 && G.r_hi % 16 == 0,
               FD_ERR_INVALID, "dat_bwd_dgrad: bad trainable slice [%d, %d) of %d", G.r_lo, G.r_hi, G.r_total);
    FD_REQUIRE(G.ld_t >= G.r_hi - G.r_lo && G.ld_t % 8 == 0, FD_ERR_INVALID, "dat_bwd_dgrad: bad row stride ld_t=%d", G.ld_t);
    FD_REQUIRE(((reinterpret_cast<uintptr_t>(G.H_t) | reinterpret_cast<uintptr_t>(G.dP_t)) & 15) == 0,
               FD_ERR_INVALID, "dat_bwd_dgrad: H_t / dP_t must be 16-byte aligned");
  }
  *saved = G.H_in != nullptr;
  HostGroup h{};
  // saved mode never touches X / Wd_cat: any valid tensor keeps the (unused) tensor maps well formed
  h.X = *saved ? G.dY : G.X;
  h.Res = G.dY; h.Out = G.dX;
  h.W1 = *saved ? G.WuT_cat : G.Wd_cat;
  h.W2 = G.WdT_cat; h.W1b = G.WuT_cat;
  h.cfg.M = static_cast<int>(G.M); h.cfg.R = G.r_total; h.cfg.num_tiles = tiles_of(G.M);
  h.cfg.scale = G.branch_scale;
  h.cfg.bd = *saved ? nullptr : G.bd_cat;
  h.cfg.bu = nullptr;
  h.cfg.has_out = G.dX != nullptr;
  h.cfg.has_res = (G.dX != nullptr && G.add_dy) ? 1 : 0;
  h.cfg.r_lo = G.dP_t ? G.r_lo : 0;
  h.cfg.r_hi = G.dP_t ? G.r_hi : 0;
  h.cfg.ld_t = G.ld_t;
  h.cfg.H_t = static_cast<__nv_bfloat16*>(G.H_t);
  h.cfg.dP_t = static_cast<__nv_bfloat16*>(G.dP_t);
  h.cfg.H_in = static_cast<const __nv_bfloat16*>(G.H_in);
  *out = h;
  return FD_OK;
}

int fwd_one(const FeddatDatGroup& G, int d, int act, int dtype, cudaStream_t st) {
  HostGroup h;
  int rc = fwd_group(G, d, act, dtype, &h);
  if (rc) return rc;
  if (G.M == 0) return FD_OK;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  // more super-tiles than CTA pairs: the tile-pipelined kernel (dat_fwd_pipe.cu)
  if ((h.cfg.num_tiles + 1) / 2 > sms / 2 && !g_force_fused)
    return launch_dat_fwd_pipe(G.X, G.Res, G.Y, G.Wd_cat, G.bd_cat, G.Wu_cat, G.bu_cat, G.H_out, G.M, G.r_total,
                               G.branch_scale, act, 2 * (sms / 2), st);
  return launch_fused(false, false, act, &h, 1, st, "dat_fwd");
}

int bwd_one(const FeddatDatGroup& G, int d, int act, int dtype, cudaStream_t st) {
  HostGroup h;
  bool saved = false;
  int rc = bwd_group(G, d, act, dtype, &h, &saved);
  if (rc) return rc;
  if (G.M == 0) return FD_OK;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  // more super-tiles than CTA pairs: the tile-pipelined kernel (dat_fwd_pipe.cu, kBwd)
  if (saved && G.dX != nullptr && !g_force_fused && (h.cfg.num_tiles + 1) / 2 > sms / 2)
    return launch_dat_bwd_pipe(G.dY, G.dX, G.WuT_cat, G.WdT_cat, G.H_in, G.dP_t, G.ld_t, G.r_lo, G.r_hi, G.M,
                               G.r_total, G.branch_scale, G.add_dy, 2 * (sms / 2), st);
  return launch_fused(true, saved, act, &h, 1, st, "dat_bwd_dgrad");
}

}  // namespace
}  // namespace fd

extern "C" int feddat_dat_fwd_grouped(const FeddatDatGroup* groups, int n_groups, int d, int act, int dtype,
                                      void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(groups != nullptr && n_groups >= 1 && n_groups <= MAX_GROUPS, FD_ERR_INVALID,
             "dat_fwd_grouped: 1 or 2 groups (got %d)", n_groups);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  HostGroup hg[MAX_GROUPS];
  int pairs = 0, live = 0;
  for (int g = 0; g < n_groups; ++g) {
    if ((rc = fwd_group(groups[g], d, act, dtype, &hg[g]))) return rc;
    pairs += (hg[g].cfg.num_tiles + 1) / 2;
    live += groups[g].M > 0;
  }
  if (n_groups == 1 || live < n_groups || 2 * pairs > sms) {   // too big for one wave (or empty groups): one by one
    for (int g = 0; g < n_groups; ++g)
      if ((rc = fwd_one(groups[g], d, act, dtype, st))) return rc;
    return FD_OK;
  }
  return launch_fused(false, false, act, hg, n_groups, st, "dat_fwd_grouped");
}

extern "C" int feddat_dat_bwd_dgrad_grouped(const FeddatDatGroup* groups, int n_groups, int d, int act, int dtype,
                                            void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(groups != nullptr && n_groups >= 1 && n_groups <= MAX_GROUPS, FD_ERR_INVALID,
             "dat_bwd_dgrad_grouped: 1 or 2 groups (got %d)", n_groups);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  HostGroup hg[MAX_GROUPS];
  bool saved[MAX_GROUPS] = {false, false};
  int pairs = 0, live = 0;
  for (int g = 0; g < n_groups; ++g) {
    if ((rc = bwd_group(groups[g], d, act, dtype, &hg[g], &saved[g]))) return rc;
    pairs += (hg[g].cfg.num_tiles + 1) / 2;
    live += groups[g].M > 0;
  }
  const bool same_mode = n_groups == 1 || saved[0] == saved[1];
  if (n_groups == 1 || live < n_groups || !same_mode || 2 * pairs > sms) {
    for (int g = 0; g < n_groups; ++g)
      if ((rc = bwd_one(groups[g], d, act, dtype, st))) return rc;
    return FD_OK;
  }
  return launch_fused(true, saved[0], act, hg, n_groups, st, "dat_bwd_dgrad_grouped");
}

extern "C" int feddat_dat_fwd(const void* X, const void* Res, void* Y, const void* Wd_cat,
                              const float* bd_cat, const void* Wu_cat, const float* bu_cat,
                              void* H_out, int64_t M, int d, int r_total, float branch_scale, int act,
                              int dtype, void* stream) {
  FeddatDatGroup G{};
  G.X = X; G.Res = Res; G.Y = Y; G.Wd_cat = Wd_cat; G.bd_cat = bd_cat; G.Wu_cat = Wu_cat; G.bu_cat = bu_cat;
  G.H_out = H_out; G.M = M; G.r_total = r_total; G.branch_scale = branch_scale;
  return feddat_dat_fwd_grouped(&G, 1, d, act, dtype, stream);
}

extern "C" int feddat_dat_bwd_dgrad(const void* X, const void* dY, void* dX, const void* Wd_cat,
                                    const float* bd_cat, const void* WuT_cat, const void* WdT_cat,
                                    const void* H_in, void* H_t, void* dP_t, int ld_t, int r_lo, int r_hi,
                                    int64_t M, int d, int r_total, float branch_scale, int act, int add_dy,
                                    int dtype, void* stream) {
  FeddatDatGroup G{};
  G.X = X; G.dY = dY; G.dX = dX; G.Wd_cat = Wd_cat; G.bd_cat = bd_cat; G.WuT_cat = WuT_cat; G.WdT_cat = WdT_cat;
  G.H_in = H_in; G.H_t = H_t; G.dP_t = dP_t; G.ld_t = ld_t; G.r_lo = r_lo; G.r_hi = r_hi;
  G.M = M; G.r_total = r_total; G.branch_scale = branch_scale; G.add_dy = add_dy;
  return feddat_dat_bwd_dgrad_grouped(&G, 1, d, act, dtype, stream);
}

#ifdef FEDDAT_DEBUG
// debug / A-B measurement: route every forward through dat_fused_kernel (1) or choose by size (0)
extern "C" int feddat_debug_force_fused_fwd(int on) {
  fd::g_force_fused = on != 0;
  return 0;
}

// debug: device buffer of 256 uint64 receiving CTA 0's pipeline timestamps (NULL disables)
extern "C" int feddat_debug_set_trace(void* dev_buf) {
  fd::g_trace = static_cast<unsigned long long*>(dev_buf);
  return 0;
}
#endif  // FEDDAT_DEBUG
