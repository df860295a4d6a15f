// MKD head (reference kl_loss, task_trainer.py:506-516, plus the task loss and their (a + b) / 2 combination,
// task_trainer.py:296-301 / 316-321), forward value and d/dlogits in one pass over the logits:
//
//   feddat_mkd_loss     ViLT:  T^2 KL(softmax(teacher/T) || softmax(logits/T)) 'batchmean'
//                              + BCEWithLogits('mean') * C                (task_trainer.py:299,319)
//   feddat_mkd_ce_loss  ALBEF: the same KL over the shifted decoder logits [n_seq, La-1, 30522] (batchmean over
//                              n_seq, pads not masked) + the answer loss  sum_s w_s * sum_p CE(logits[s,p],
//                              labels[s,p+1])  (xbert.py:1287-1297 with reduction='none', albef_model.py:142-143;
//                              ignore_index -100), reading the UNSHIFTED prediction scores [n_seq, La, C] in place
//                              (the reference copies the [:, :-1] slice twice) in bf16 or fp32.
//
// Layout: one warp per row when C <= 2048 (ViLT answer heads: C = 100), one 1024-thread CTA per row otherwise
// (ALBEF vocabulary: C = 30522).  Two passes over the row: (1) online max / sum-exp of the operands, (2) KL
// terms, task terms and the gradient; the second pass re-reads the row from L1/L2.  The byte floor is 2 reads +
// 1 write of the logits; at the ALBEF batch's size (198 rows x 30 522) the kernel is instruction / MUFU bound
// instead (six exponentials per element: three running sums and three probabilities; ncu: XU pipe 43 %, 18 M
// warp instructions, 41 us for 36 MB -- profiles/r2_summary.md).
//
// Deterministic: every row's (kl, task) pair goes to a caller-provided workspace and ONE block sums the rows in a
// fixed order (round 1 accumulated per-CTA partials with float atomics: run-to-run different low bits).
#include <cuda_bf16.h>
#include <math.h>

#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

struct MkdParams {
  const void* logits;
  const void* teacher;
  const float* target;       // BCE soft targets [rows, C] (ViLT) or null
  void* dlogits;
  float* row_ws;             // [rows, 2]: per-row (kl, task) sums, unscaled
  int64_t rows;
  int C;
  float inv_temp;
  float kl_grad_scale;       // kl_weight * T / batchmean_div
  float task_grad_scale;     // task_weight * task_scale
  // token-CE mode (ALBEF)
  const int64_t* labels;     // [n_seq, La] or null
  const float* seq_weight;   // [n_seq]
  int La, La_teacher;        // positions per sequence of logits / of teacher (La or La - 1)
};

struct OnlineLse {
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  // a batch whose maximum is `bm` and whose sum of exp(v - max(m, bm)) the caller provides through `add`:
  // ONE rescale per batch instead of a compare / branch / dependent exp per element
  __device__ __forceinline__ float begin_batch(float bm) {
    const float mn = fmaxf(m, bm);
    s *= __expf(m - mn);          // m = -inf at the start: exp(-inf) = 0
    m = mn;
    return mn;
  }
  __device__ __forceinline__ void merge(float om, float os) {
    if (om == -INFINITY) return;
    if (om > m) {
      s = s * __expf(m - om) + os;
      m = om;
    } else {
      s += os * __expf(om - m);
    }
  }
};

template <int GROUP>  // threads cooperating on one row: 32 (warp) or 256 (CTA)
__device__ __forceinline__ float group_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if constexpr (GROUP == 32) {
    return v;
  } else {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    float t = (l < GROUP / 32) ? scratch[l] : 0.f;   // GROUP <= 1024: at most 32 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  }
}

template <int GROUP>
__device__ __forceinline__ void group_merge_lse(OnlineLse& a, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, a.m, o);
    float os = __shfl_xor_sync(0xffffffffu, a.s, o);
    a.merge(om, os);
  }
  if constexpr (GROUP != 32) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) {             // scratch holds 64 floats: (m, s) of up to 32 warps
      scratch[2 * w] = a.m;
      scratch[2 * w + 1] = a.s;
    }
    __syncthreads();
    OnlineLse t;
    t.init();
    if (l < GROUP / 32) {
      t.m = scratch[2 * l];
      t.s = scratch[2 * l + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, t.m, o);
      float os = __shfl_xor_sync(0xffffffffu, t.s, o);
      t.merge(om, os);
    }
    a = t;
  }
}

// VEC consecutive elements of a row of T (float or bf16) <-> fp32 registers
template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* row, int i, float (&v)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC == 4) {
      const float4 a = reinterpret_cast<const float4*>(row)[i];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if constexpr (VEC == 2) {
      const float2 a = reinterpret_cast<const float2*>(row)[i];
      v[0] = a.x; v[1] = a.y;
    } else {
      v[0] = reinterpret_cast<const float*>(row)[i];
    }
  } else {
    const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(row);
    if constexpr (VEC == 8) {
      const uint4 a = reinterpret_cast<const uint4*>(r)[i];
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[2 * k] = __uint_as_float(w[k] << 16);
        v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
    } else if constexpr (VEC == 2) {
      const uint32_t w = reinterpret_cast<const uint32_t*>(r)[i];
      v[0] = __uint_as_float(w << 16);
      v[1] = __uint_as_float(w & 0xffff0000u);
    } else {
      v[0] = __bfloat162float(r[i]);
    }
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* row, int i, const float (&v)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC == 4)
      reinterpret_cast<float4*>(row)[i] = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VEC == 2)
      reinterpret_cast<float2*>(row)[i] = make_float2(v[0], v[1]);
    else
      reinterpret_cast<float*>(row)[i] = v[0];
  } else {
    __nv_bfloat16* r = reinterpret_cast<__nv_bfloat16*>(row);
    if constexpr (VEC == 8) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<uint32_t*>(&t);
      }
      reinterpret_cast<uint4*>(r)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    } else if constexpr (VEC == 2) {
      __nv_bfloat162 t = __floats2bfloat162_rn(v[0], v[1]);
      reinterpret_cast<__nv_bfloat162*>(r)[i] = t;
    } else {
      r[i] = __float2bfloat16_rn(v[0]);
    }
  }
}

// kCe = false: BCE task term from `target` (may be null: KL only); kCe = true: token cross-entropy from `labels`
template <int GROUP>
constexpr int block_threads() { return GROUP == 32 ? 256 : GROUP; }

template <typename T, int GROUP, int VEC, bool kCe>
__global__ void __launch_bounds__(block_threads<GROUP>()) mkd_loss_kernel(const MkdParams p) {
  // vectors in flight per thread and operand (64 registers per thread at 1024 threads per CTA)
  constexpr int U = GROUP == 32 ? 2 : (VEC >= 8 ? 1 : (VEC >= 4 ? 2 : 4));
  __shared__ float scratch[64];
  const int groups_per_block = block_threads<GROUP>() / GROUP;
  const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
  const T* logits = static_cast<const T*>(p.logits);
  const T* teacher = static_cast<const T*>(p.teacher);
  T* dlogits = static_cast<T*>(p.dlogits);

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * groups_per_block + g; row < p.rows;
       row += static_cast<int64_t>(gridDim.x) * groups_per_block) {
    const T* x = logits + row * p.C;
    T* dx = dlogits ? dlogits + row * p.C : nullptr;
    const int nvec = p.C / VEC;
    int64_t trow = row;
    int label = -100;
    float w_seq = 0.f;
    if constexpr (kCe) {
      const int64_t s = row / p.La;
      const int pos = static_cast<int>(row - s * p.La);
      if (pos == p.La - 1) {
        // the last position predicts nothing (prediction_scores[:, :-1], xbert.py:1289): no loss, zero gradient
        if (dx) {
          float z[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) z[k] = 0.f;
          for (int i = t; i < nvec; i += GROUP) store_vec<T, VEC>(dx, i, z);
        }
        if (t == 0) {
          p.row_ws[2 * row] = 0.f;
          p.row_ws[2 * row + 1] = 0.f;
        }
        continue;
      }
      trow = s * p.La_teacher + pos;
      label = static_cast<int>(p.labels[s * p.La + pos + 1]);
      w_seq = p.seq_weight[s];
    }
    const T* y = teacher + trow * p.C;
    const float* tg = (!kCe && p.target) ? p.target + row * p.C : nullptr;

    OnlineLse la, lb, lc;      // logits / T, teacher / T, logits (CE)
    la.init();
    lb.init();
    lc.init();
    // U vectors per thread are loaded before any is consumed (one 4-byte load in flight per thread ran at
    // 0.4 TB/s), and the log-sum-exps advance per BATCH: batch maximum first (independent FMNMX), one rescale of
    // the running sum, then independent exps -- the per-element online update (compare, branch, dependent exp)
    // made the kernel issue-bound at 9 % of HBM bandwidth with the 198 rows of an ALBEF batch.
    for (int i0 = t; i0 < nvec; i0 += GROUP * U) {
      float xv[U][VEC], yv[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * GROUP;
        if (i < nvec) {
          load_vec<T, VEC>(x, i, xv[u]);
          load_vec<T, VEC>(y, i, yv[u]);
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) xv[u][k] = yv[u][k] = -INFINITY;   // exp(-inf - m) = 0: no predicates below
        }
      }
      float mx = -INFINITY, my = -INFINITY;
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          mx = fmaxf(mx, xv[u][k]);
          my = fmaxf(my, yv[u][k]);
        }
      const float ma = la.begin_batch(mx * p.inv_temp), mb = lb.begin_batch(my * p.inv_temp);
      const float mc = kCe ? lc.begin_batch(mx) : 0.f;
      float sa = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          sa += __expf(fmaf(xv[u][k], p.inv_temp, -ma));
          sb += __expf(fmaf(yv[u][k], p.inv_temp, -mb));
          if constexpr (kCe) sc += __expf(xv[u][k] - mc);
        }
      la.s += sa;
      lb.s += sb;
      if constexpr (kCe) lc.s += sc;
    }
    group_merge_lse<GROUP>(la, scratch);
    group_merge_lse<GROUP>(lb, scratch);
    if constexpr (kCe) group_merge_lse<GROUP>(lc, scratch);
    const float lse_a = la.m + __logf(la.s);
    const float log_sb = __logf(lb.s);
    const float inv_sb = 1.f / lb.s;
    const float lse_c = kCe ? lc.m + __logf(lc.s) : 0.f;
    const bool ce_on = kCe && label >= 0;
    const float ce_g = ce_on ? p.task_grad_scale * w_seq : 0.f;

    float kl_row = 0.f, task_row = 0.f;
    for (int i0 = t; i0 < nvec; i0 += GROUP * U) {
      float xu[U][VEC], yu[U][VEC], tu[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * GROUP;
        if (i < nvec) {
          load_vec<T, VEC>(x, i, xu[u]);
          load_vec<T, VEC>(y, i, yu[u]);
          if (tg) load_vec<float, VEC>(tg, i, tu[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * GROUP;
        if (i >= nvec) continue;
        float gv[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float xx = xu[u][k];
          const float a = xx * p.inv_temp, b = yu[u][k] * p.inv_temp;
          const float logp = a - lse_a;
          const float bq = b - lb.m;
          const float q = __expf(bq) * inv_sb;
          const float logq = bq - log_sb;
          if (q > 0.f) kl_row += q * (logq - logp);  // xlogy convention of F.kl_div
          float grad = p.kl_grad_scale * (__expf(logp) - q);
          if constexpr (kCe) {
            if (ce_on) {
              const int col = i * VEC + k;
              const float sm = __expf(xx - lse_c);
              grad += ce_g * (sm - (col == label ? 1.f : 0.f));
              if (col == label) task_row += lse_c - xx;
            }
          } else if (tg) {
            const float e = __expf(-fabsf(xx));
            task_row += fmaxf(xx, 0.f) - xx * tu[u][k] + log1pf(e);
            const float sig = xx >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
            grad += p.task_grad_scale * (sig - tu[u][k]);
          }
          gv[k] = grad;
        }
        if (dx) store_vec<T, VEC>(dx, i, gv);
      }
    }
    kl_row = group_sum<GROUP>(kl_row, scratch);
    task_row = group_sum<GROUP>(task_row, scratch);
    if (t == 0) {
      p.row_ws[2 * row] = kl_row;
      p.row_ws[2 * row + 1] = kCe ? w_seq * task_row : task_row;
    }
  }
}

// one block, fixed summation order: loss_out = {kl_weight * kl + task_weight * task, kl, task}
__global__ void __launch_bounds__(256) mkd_finalize_kernel(const float* row_ws, int64_t rows, float kl_row_scale,
                                                          float task_scale, float kl_weight, float task_weight,
                                                          float* loss_out) {
  __shared__ float sh[2][256];
  float kl = 0.f, task = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += 256) {
    kl += row_ws[2 * r];
    task += row_ws[2 * r + 1];
  }
  sh[0][threadIdx.x] = kl;
  sh[1][threadIdx.x] = task;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float k = sh[0][0] * kl_row_scale, t = sh[1][0] * task_scale;
    loss_out[0] = kl_weight * k + task_weight * t;
    loss_out[1] = k;
    loss_out[2] = t;
  }
}

template <typename T, int GROUP, int VEC, bool kCe>
int launch(const MkdParams& p, int sms, cudaStream_t st) {
  const int groups_per_block = block_threads<GROUP>() / GROUP;
  int64_t blocks = (p.rows + groups_per_block - 1) / groups_per_block;
  const int64_t cap = static_cast<int64_t>(sms) * 8;
  if (blocks > cap) blocks = cap;
  mkd_loss_kernel<T, GROUP, VEC, kCe><<<static_cast<int>(blocks), block_threads<GROUP>(), 0, st>>>(p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

// widest vector (in elements) that keeps every row of every operand aligned
int common_vec(int C, size_t elem, std::initializer_list<const void*> ptrs, int widest) {
  for (int v = widest; v > 1; v >>= 1) {
    bool ok = C % v == 0 && (static_cast<size_t>(C) * elem) % (v * elem) == 0;
    for (const void* q : ptrs) ok = ok && (q == nullptr || reinterpret_cast<uintptr_t>(q) % (v * elem) == 0);
    if (ok) return v;
  }
  return 1;
}

}  // namespace
}  // namespace fd

extern "C" int feddat_mkd_loss(const float* logits, const float* teacher, const float* target,
                               float* loss_out, float* dlogits, int64_t rows, int C, float temp,
                               float kl_weight, float task_weight, float task_scale,
                               int64_t batchmean_div, float* row_ws, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(logits && teacher && loss_out && row_ws, FD_ERR_INVALID, "mkd_loss: null pointer argument");
  FD_REQUIRE(rows >= 0 && C >= 1, FD_ERR_INVALID, "mkd_loss: bad shape rows=%lld C=%d",
             (long long)rows, C);
  FD_REQUIRE(temp > 0.f && batchmean_div > 0, FD_ERR_INVALID,
             "mkd_loss: temp and batchmean_div must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MkdParams p{};
  p.logits = logits; p.teacher = teacher; p.target = target; p.dlogits = dlogits; p.row_ws = row_ws;
  p.rows = rows; p.C = C;
  p.inv_temp = 1.f / temp;
  p.kl_grad_scale = kl_weight * temp / static_cast<float>(batchmean_div);
  p.task_grad_scale = task_weight * task_scale;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  if (rows > 0) {
    const int vec = common_vec(C, 4, {logits, teacher, target, dlogits}, 4);
    if (C <= 2048) {
      rc = vec == 4 ? launch<float, 32, 4, false>(p, sms, st)
           : vec == 2 ? launch<float, 32, 2, false>(p, sms, st) : launch<float, 32, 1, false>(p, sms, st);
    } else {
      rc = vec == 4 ? launch<float, 1024, 4, false>(p, sms, st)
           : vec == 2 ? launch<float, 1024, 2, false>(p, sms, st) : launch<float, 1024, 1, false>(p, sms, st);
    }
    if (rc) return rc;
  }
  mkd_finalize_kernel<<<1, 256, 0, st>>>(row_ws, rows, temp * temp / static_cast<float>(batchmean_div), task_scale,
                                         kl_weight, task_weight, loss_out);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

extern "C" int feddat_mkd_ce_loss(const void* logits, const void* teacher, const int64_t* labels,
                                  const float* seq_weight, float* loss_out, void* dlogits, int64_t n_seq, int La,
                                  int La_teacher, int C, float temp, float kl_weight, float task_weight, int dtype,
                                  float* row_ws, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(logits && teacher && labels && seq_weight && loss_out && row_ws, FD_ERR_INVALID,
             "mkd_ce_loss: null pointer argument");
  FD_REQUIRE(n_seq >= 0 && La >= 2 && C >= 1 && (La_teacher == La || La_teacher == La - 1), FD_ERR_INVALID,
             "mkd_ce_loss: bad shape n_seq=%lld La=%d La_teacher=%d C=%d", (long long)n_seq, La, La_teacher, C);
  FD_REQUIRE(temp > 0.f, FD_ERR_INVALID, "mkd_ce_loss: temp must be positive");
  FD_REQUIRE(dtype == FEDDAT_DTYPE_BF16 || dtype == FEDDAT_DTYPE_F32, FD_ERR_UNSUPPORTED,
             "mkd_ce_loss: logits must be bf16 or fp32 (dtype=%d)", dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MkdParams p{};
  p.logits = logits; p.teacher = teacher; p.dlogits = dlogits; p.row_ws = row_ws;
  p.labels = labels; p.seq_weight = seq_weight; p.La = La; p.La_teacher = La_teacher;
  p.rows = n_seq * La; p.C = C;
  p.inv_temp = 1.f / temp;
  // batchmean of the reference divides by the FIRST dimension of the [n_seq, La-1, C] logits (pads not masked)
  const float div = static_cast<float>(n_seq > 0 ? n_seq : 1);
  p.kl_grad_scale = kl_weight * temp / div;
  p.task_grad_scale = task_weight;     // seq_weight carries weights / batch (albef_model.py:142-143)
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  if (p.rows > 0) {
    if (dtype == FEDDAT_DTYPE_BF16) {
      const int vec = common_vec(C, 2, {logits, teacher, dlogits}, 8);
      rc = vec == 8 ? launch<__nv_bfloat16, 1024, 8, true>(p, sms, st)
           : vec >= 2 ? launch<__nv_bfloat16, 1024, 2, true>(p, sms, st) : launch<__nv_bfloat16, 1024, 1, true>(p, sms, st);
    } else {
      const int vec = common_vec(C, 4, {logits, teacher, dlogits}, 4);
      rc = vec == 4 ? launch<float, 1024, 4, true>(p, sms, st)
           : vec == 2 ? launch<float, 1024, 2, true>(p, sms, st) : launch<float, 1024, 1, true>(p, sms, st);
    }
    if (rc) return rc;
  }
  mkd_finalize_kernel<<<1, 256, 0, st>>>(row_ws, p.rows, temp * temp / div, 1.f, kl_weight, task_weight, loss_out);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}
