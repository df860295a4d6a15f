// The frozen MLP's first GEMM with the exact-GELU fused into its epilogue:
//
//   forward   pre = A W^T + b,  act = gelu(pre)                 A [M, K] (LayerNorm-after output),  W [N, K]
//   backward  dpre = (dY W2T^T) * gelu'(pre)                    dY [M, K] (gradient at the block's second
//                                                               residual), W2T [N, K] = ViltOutput.dense.weight^T
//
// i.e. HF ``ViltIntermediate`` (dense 768 -> 3072 + GELU) of the block the reference wraps with
// ``Adaptered_ViltOutput`` (src/modeling/adaptered_output.py:67-79; the backbone itself is third-party
// ``transformers.ViltModel``, src/modeling/vilt.py:19,127) and the data gradient that autograd derives for it.
// Both are [M = 2 x batch x tokens, N = 3072, K = 768] products with K contiguous in both operands.  Unfused,
// each is a cuBLAS GEMM (47-52 us at M = 11 840) followed by a streaming GELU pass over the [M, 3072] tensor
// (30 / 36 us, HBM-bound); here the activation rides in the epilogue while the tensor cores run the next tile
// (60-62 us / 63 us for the fused pair of outputs; the bare main loop of this kernel takes 43-45 us).
//
// Persistent CTA pairs (cluster of 2, tcgen05 cta_group::2), one 256 x 256 output tile per pair and step:
//   ring       NS stages of 32 KB: [A k-chunk 128 x 64 | W half k-chunk 128 x 64] per CTA, K-major SW128
//   warp 0     lanes 0 / 1: TMA producers (activations / weights), running ahead across tiles
//   warp 1     tcgen05.mma issuer (leader CTA): M = 256 across the pair, N = 256, fp32 accumulators in TMEM;
//              TWO accumulator buffers (2 x 256 columns), so the MMAs of tile t + 1 run under the epilogue of t
//   warp 2     TMEM allocation
//   warps 4-19 epilogue: SIXTEEN warps, one per (TMEM lane quarter = 32 tile rows) x (64-column chunk of the 256),
//              each with ONE private 4 KB staging slab and its own stores, so no barrier couples the warps:
//              tcgen05.ld (32 columns at a time) -> math in packed fp32 (FFMA2 / FMUL2: ~10 issue slots per element
//              against ~21 scalar -- an 8-warp scalar epilogue took 8 us per tile against 5.4 us of MMA) -> bf16.
//     forward  pre goes into the slab and leaves as plain 16-byte stores (eight lanes per 128-byte row segment, so
//              every store instruction writes four full lines; one row per lane would write 32 half-used sectors);
//              act waits in 32 registers, then takes the slab and leaves as one TMA store.  The bias is fp32 in
//              global memory (L1-resident), requested one 8-column step ahead of its use.
//     backward the saved pre-activations are loaded BEFORE the accumulator wait as full row segments (streaming,
//              last use), turned into row-per-thread order through the slab, and each thread overwrites its own
//              chunk with dpre for the TMA store.
// What bounds it (profiles/r2_summary.md section E): L2 -> SM operand traffic of 256 x 256 tiles is 440 MB per launch
// (~10 TB/s while the main loop runs); the 73-145 MB of epilogue stores and loads share that path, and each costs
// about its share (+6 us per [M, 3072] tensor moved).  Math alone is hidden (bwd: +0.4 us).
// Rounding: every output is rounded once, from fp32: pre = bf16(acc + b), act = bf16(gelu(acc + b)),
// dpre = bf16(acc * gelu'(pre)).
#include <stdlib.h>

#include "feddat_b200.h"
#include "gelu_math.cuh"
#include "host_common.h"
#include "ptx_sm100.cuh"

namespace fd {
namespace {

constexpr int BM = 128;               // rows per CTA (256 per pair)
constexpr int TN = 256;               // tile columns
constexpr int BK = 64;
constexpr int SLOT = BM * 128;        // 16 KB: [128 rows x 64 bf16], 128-byte swizzle
constexpr int STAGE = 2 * SLOT;       // A chunk + W half chunk
constexpr int NUM_THREADS = 640;     // 4 control warps + 16 epilogue warps

constexpr int NS = 5;                 // ring stages (160 KB)
constexpr int NEW = 16;               // epilogue warps
constexpr int NBUF = 1;               // staging slabs per epilogue warp (64 KB in all)
constexpr int WSLOT = 32 * 128;       // 4 KB: [32 rows x 64 bf16] slab of one warp, 128-byte swizzle

struct GemmParams {
  int M, N, K;
  int m_blocks;          // ceil(M / 256)
  int n_tiles;           // m_blocks * (N / 256)
  __nv_bfloat16* pre_out;      // fwd: [M, N], written from registers
  const __nv_bfloat16* pre_in; // bwd: [M, N], read into registers
  const float* bias;           // fwd: [N], fp32
};

struct GemmTmaps {
  CUtensorMap a, w, out0;      // out0: the TMA-stored output (fwd: act, bwd: dpre)
};

template <bool kBwd>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_gemm_kernel(const __grid_constant__ GemmTmaps tm, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NS + 4];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem0 + NS * STAGE;            // [epilogue warp][NBUF] slabs of 4 KB
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto bar_empty = [&](uint32_t s) { return bar0 + 8u * (NS + s); };
  auto bar_acc_full = [&](uint32_t b) { return bar0 + 8u * (2 * NS + b); };
  auto bar_acc_empty = [&](uint32_t b) { return bar0 + 8u * (2 * NS + 2 + b); };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 2 * NEW);   // one lane of each epilogue warp, both CTAs
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.w);
    tma_prefetch_desc(&tm.out0);
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem = tmem_base_smem;

  const int pid = blockIdx.x >> 1, n_pairs = static_cast<int>(gridDim.x >> 1);
  const int KC = p.K / BK;
  // tile t -> (row block, column block): row blocks vary fastest, so the pairs running at one time share W tiles
  auto tile_m0 = [&](int t) { return (t % p.m_blocks) * 2 * BM + static_cast<int>(rank) * BM; };
  auto tile_n0 = [&](int t) { return (t / p.m_blocks) * TN; };
  const uint32_t leader_full0 = mapa_u32(bar_full(0), 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producers
    if (lane < 2) {
      uint32_t n = 0;
      for (int t = pid; t < p.n_tiles; t += n_pairs) {
        const int c1 = lane == 0 ? tile_m0(t) : tile_n0(t) + static_cast<int>(rank) * (TN / 2);
        const CUtensorMap* map = lane == 0 ? &tm.a : &tm.w;
        for (int kc = 0; kc < KC; ++kc, ++n) {
          const uint32_t s = n % NS, par = (n / NS) & 1;
          mbar_wait(bar_empty(s), par ^ 1);
          if (lane == 0 && rank == 0) mbar_arrive_expect_tx(bar_full(s), 2 * STAGE);
          tma_load_2d_pair(smem0 + s * STAGE + lane * SLOT, map, leader_full0 + 8u * s, kc * BK, c1, kEvictLast);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * BM, TN);
      uint32_t n = 0, it = 0;
      for (int t = pid; t < p.n_tiles; t += n_pairs, ++it) {
        const uint32_t b = it & 1, use = it >> 1;
        mbar_wait(bar_acc_empty(b), (use & 1) ^ 1);        // the epilogues of tile it - 2 have drained buffer b
        tc_fence_after();
        const uint32_t d_tmem = tmem + b * TN;
        for (int kc = 0; kc < KC; ++kc, ++n) {
          const uint32_t s = n % NS, par = (n / NS) & 1;
          mbar_wait(bar_full(s), par);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = desc_kmajor_sw128(smem0 + s * STAGE);
            const uint64_t bdesc = adesc + (SLOT >> 4);
            umma_ss_pair(d_tmem, adesc, bdesc, idesc, kc != 0);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
            umma_commit_pair(bar_empty(s), 0b11);
            if (kc == KC - 1) umma_commit_pair(bar_acc_full(b), 0b11);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps
    const uint32_t ew = warp - 4;                    // 0..15
    const uint32_t cc = ew >> 2;                     // 64-column chunk of the tile
    const uint32_t q = warp & 3;                     // TMEM lane quarter == 32-row slab of the tile
    const uint32_t lane_addr = (q * 32) << 16;
    const uint32_t wbuf = stg_base + ew * WSLOT;
    const uint32_t leader_acc_empty0 = mapa_u32(bar_acc_empty(0), 0);
    uint32_t it = 0;

    for (int t = pid; t < p.n_tiles; t += n_pairs, ++it) {
      const uint32_t b = it & 1, use = it >> 1;
      const int m0 = tile_m0(t) + static_cast<int>(q) * 32, col0 = tile_n0(t) + static_cast<int>(cc) * 64;
      const uint32_t t_src = tmem + lane_addr + b * TN + cc * 64;
      // backward: this warp's [32 rows x 64 columns] of saved pre-activations, loaded ahead of the accumulator as
      // full 128-byte row segments (eight lanes per row, four rows per instruction; last use -> streaming)
      const uint32_t j8 = lane & 7, r0 = lane >> 3;
      uint4 pv[8];
      if constexpr (kBwd) {
        const __nv_bfloat16* src = p.pre_in + static_cast<size_t>(m0 + r0) * p.N + col0 + j8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pv[i] = (m0 + static_cast<int>(r0) + 4 * i < p.M)
                      ? __ldcs(reinterpret_cast<const uint4*>(src + static_cast<size_t>(4 * i) * p.N))
                      : make_uint4(0u, 0u, 0u, 0u);
      }
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
      float4 bn0 = make_float4(0.f, 0.f, 0.f, 0.f), bn1 = bn0;
      if constexpr (!kBwd) {
        bn0 = __ldg(b4);
        bn1 = __ldg(b4 + 1);
      }
      mbar_wait(bar_acc_full(b), use & 1);
      tc_fence_after();
      if (lane == 0) tma_store_wait_read<0>();        // this warp's store of the previous tile has left the slab
      __syncwarp();
      if constexpr (kBwd) {
        // through the slab into row-per-thread order (the layout of the accumulator in registers)
#pragma unroll
        for (int i = 0; i < 8; ++i) st_shared_v4(wbuf + sw128_offset(r0 + 4 * i, j8), pv[i].x, pv[i].y, pv[i].z, pv[i].w);
        __syncwarp();
      }
      uint32_t act[kBwd ? 1 : 32];                    // forward: the packed activations wait for the slab
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        uint32_t v[32];
        tmem_ld32(t_src + st * 32, v);
        tmem_ld_wait32(v);
        if (st == 1) {
          // the accumulator chunk of this warp is in registers: hand the buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_addr(leader_acc_empty0 + 8u * b);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t oo[4];
          if constexpr (!kBwd) {
            // forward: pre = acc + b -> slab (flushed below with plain stores); act = gelu(pre) -> registers.
            // The bias of the NEXT eight columns is requested before this step's math (an L1 hit still takes ~40
            // cycles, and 96 registers leave the compiler no room to hoist the loads itself)
            const float2 bb[4] = {make_float2(bn0.x, bn0.y), make_float2(bn0.z, bn0.w), make_float2(bn1.x, bn1.y),
                                  make_float2(bn1.z, bn1.w)};
            if (st * 4 + c < 7) {
              bn0 = __ldg(b4 + 2 * (st * 4 + c + 1));
              bn1 = __ldg(b4 + 2 * (st * 4 + c + 1) + 1);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = c * 8 + 2 * i;
              const float2 a = __fadd2_rn(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), bb[i]);
              oo[i] = pack_bf16x2(a.x, a.y);
              const float2 g = gelu_fwd2(a);
              act[st * 16 + c * 4 + i] = pack_bf16x2(g.x, g.y);
            }
          } else {
            // backward: dpre = acc * gelu'(pre)
            const uint4 pq = ld_shared_v4(wbuf + sw128_offset(lane, st * 4 + c));
            const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = c * 8 + 2 * i;
              const float2 x = make_float2(__uint_as_float(pw[i] << 16), __uint_as_float(pw[i] & 0xffff0000u));
              const float2 d = __fmul2_rn(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), gelu_grad2(x));
              oo[i] = pack_bf16x2(d.x, d.y);
            }
          }
          st_shared_v4(wbuf + sw128_offset(lane, st * 4 + c), oo[0], oo[1], oo[2], oo[3]);
        }
      }
      if constexpr (!kBwd) {
        // pre leaves through plain 16-byte stores, eight lanes per 128-byte row segment (full lines); the slab
        // then takes the activations for the TMA store
        __syncwarp();
        uint4 row[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) row[i] = ld_shared_v4(wbuf + sw128_offset(r0 + 4 * i, j8));
        __nv_bfloat16* dst = p.pre_out + static_cast<size_t>(m0 + r0) * p.N + col0 + j8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i)   
          if (m0 + static_cast<int>(r0) + 4 * i < p.M) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(4 * i) * p.N) = row[i];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_v4(wbuf + sw128_offset(lane, c), act[c * 4], act[c * 4 + 1], act[c * 4 + 2], act[c * 4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm.out0, wbuf, col0, m0);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem, 512);
}

int launch_mlp_gemm(bool bwd, const void* A, const void* W, const void* bias, const void* pre_in, void* out0, void* out1,
                    int64_t M, int N, int K, cudaStream_t st, const char* who) {
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(A && W && out0 && (bwd ? pre_in != nullptr : (bias != nullptr && out1 != nullptr)), FD_ERR_INVALID,
             "%s: null pointer argument", who);
  FD_REQUIRE(M >= 0 && M < (1ll << 31) - 512, FD_ERR_INVALID, "%s: bad row count %lld", who, (long long)M);
  FD_REQUIRE(N >= TN && N % TN == 0 && K >= BK && K % BK == 0, FD_ERR_UNSUPPORTED,
             "%s: N must be a multiple of %d and K a multiple of %d (got N=%d K=%d)", who, TN, BK, N, K);
  FD_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, FD_ERR_INVALID, "%s: bias must be 16-byte aligned", who);
  if (M == 0) return FD_OK;
  GemmParams p{};
  p.M = static_cast<int>(M); p.N = N; p.K = K;
  p.m_blocks = static_cast<int>((M + 2 * BM - 1) / (2 * BM));
  p.n_tiles = p.m_blocks * (N / TN);
  p.pre_out = bwd ? nullptr : static_cast<__nv_bfloat16*>(out0);
  p.pre_in = static_cast<const __nv_bfloat16*>(pre_in);
  p.bias = static_cast<const float*>(bias);
  GemmTmaps tm;
  if ((rc = make_tmap_bf16_2d(&tm.a, A, M, K, K, BM, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.w, W, N, K, K, TN / 2, 64))) return rc;
  // the TMA-stored output (forward: act, backward: dpre) moves as [32 rows x 64 columns] slabs, one per epilogue warp
  if ((rc = make_tmap_bf16_2d(&tm.out0, bwd ? out0 : out1, M, N, N, 32, 64))) return rc;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  const int pairs = p.n_tiles < sms / 2 ? p.n_tiles : sms / 2;
  const size_t max_smem = 227 * 1024 - 1024;
  const size_t smem = 1024 + static_cast<size_t>(NS) * STAGE + static_cast<size_t>(NEW) * NBUF * WSLOT;
  FD_REQUIRE(smem <= max_smem, FD_ERR_UNSUPPORTED, "%s: shared-memory budget exceeded", who);
  using KernelFn = void (*)(const GemmTmaps, const GemmParams);
  KernelFn fn = bwd ? mlp_gemm_kernel<true> : mlp_gemm_kernel<false>;
  static bool configured[2][64] = {{false}};   // idempotent "attribute already set" cache
  int dev = 0;
  FD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[bwd][dev]) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    if (dev < 64) configured[bwd][dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
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
  const char* e = getenv("FEDDAT_PDL");
  cfg.numAttrs = (e && e[0] == '0') ? 1 : 2;
  FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, tm, p));
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int feddat_mlp_fc1_gelu_fwd(const void* A, const void* W, const void* bias, void* pre_out, void* act_out,
                                       int64_t M, int N, int K, int dtype, void* stream) {
  using namespace fd;
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED, "mlp_fc1_gelu_fwd: only bf16 is implemented (dtype=%d)", dtype);
  return launch_mlp_gemm(false, A, W, bias, nullptr, pre_out, act_out, M, N, K, static_cast<cudaStream_t>(stream),
                         "mlp_fc1_gelu_fwd");
}

extern "C" int feddat_mlp_fc2_dgelu_bwd(const void* dY, const void* W2T, const void* pre, void* dpre_out, int64_t M,
                                        int N, int K, int dtype, void* stream) {
  using namespace fd;
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16, FD_ERR_UNSUPPORTED, "mlp_fc2_dgelu_bwd: only bf16 is implemented (dtype=%d)", dtype);
  return launch_mlp_gemm(true, dY, W2T, nullptr, pre, dpre_out, nullptr, M, N, K, static_cast<cudaStream_t>(stream),
                         "mlp_fc2_dgelu_bwd");
}
