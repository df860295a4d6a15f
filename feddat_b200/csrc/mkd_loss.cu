// MKD head: temperature-softmax KL (reference kl_loss, task_trainer.py:506-516) + the ViLT task loss
// BCEWithLogits('mean') * C (task_trainer.py:299,319) + their (a + b) / 2 combination
// (task_trainer.py:300-301), forward value and d/dlogits in one launch.
//
// Layout: logits / teacher / target / dlogits are [rows, C] fp32 row-major.  One warp per row when
// C <= 2048 (ViLT answer heads: C = 100), one 256-thread CTA per row otherwise (ALBEF decoder
// vocabulary: C = 30522).  Two passes over the row: (1) online max / sum-exp of both operands,
// (2) KL terms, BCE terms and the gradient.  The second pass re-reads the row from L1/L2.
// HBM-bound: 3 reads + 1 write of rows*C*4 bytes at most.
#include <math.h>

#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

struct MkdParams {
  const float* logits;
  const float* teacher;
  const float* target;
  float* loss_out;
  float* dlogits;
  int64_t rows;
  int C;
  float inv_temp;
  float kl_row_scale;   // T^2 / batchmean_div
  float kl_grad_scale;  // kl_weight * T / batchmean_div
  float kl_weight, task_weight, task_scale;
};

struct OnlineLse {
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void push(float x) {
    if (x > m) {
      s = s * __expf(m - x) + 1.f;
      m = x;
    } else {
      s += __expf(x - m);
    }
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
    float t = (l < GROUP / 32) ? scratch[l] : 0.f;
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
    if (l == 0) {
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

template <int GROUP, int VEC>
__global__ void __launch_bounds__(256) mkd_loss_kernel(const MkdParams p) {
  __shared__ float scratch[32];
  const int groups_per_block = 256 / GROUP;
  const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
  float kl_acc = 0.f, task_acc = 0.f;  // per-thread partials over all rows this group handles

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * groups_per_block + g; row < p.rows;
       row += static_cast<int64_t>(gridDim.x) * groups_per_block) {
    const float* x = p.logits + row * p.C;
    const float* y = p.teacher + row * p.C;
    const float* tg = p.target ? p.target + row * p.C : nullptr;
    float* dx = p.dlogits ? p.dlogits + row * p.C : nullptr;
    const int nvec = p.C / VEC;

    OnlineLse la, lb;
    la.init();
    lb.init();
    for (int i = t; i < nvec; i += GROUP) {
      float xv[VEC], yv[VEC];
      if constexpr (VEC == 4) {
        float4 a = reinterpret_cast<const float4*>(x)[i], b = reinterpret_cast<const float4*>(y)[i];
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        yv[0] = b.x; yv[1] = b.y; yv[2] = b.z; yv[3] = b.w;
      } else if constexpr (VEC == 2) {
        float2 a = reinterpret_cast<const float2*>(x)[i], b = reinterpret_cast<const float2*>(y)[i];
        xv[0] = a.x; xv[1] = a.y; yv[0] = b.x; yv[1] = b.y;
      } else {
        xv[0] = x[i]; yv[0] = y[i];
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        la.push(xv[k] * p.inv_temp);
        lb.push(yv[k] * p.inv_temp);
      }
    }
    group_merge_lse<GROUP>(la, scratch);
    group_merge_lse<GROUP>(lb, scratch);
    const float lse_a = la.m + __logf(la.s);
    const float log_sb = __logf(lb.s);
    const float inv_sb = 1.f / lb.s;

    float kl_row = 0.f, task_row = 0.f;
    for (int i = t; i < nvec; i += GROUP) {
      float xv[VEC], yv[VEC], tv[VEC], gv[VEC];
      if constexpr (VEC == 4) {
        float4 a = reinterpret_cast<const float4*>(x)[i], b = reinterpret_cast<const float4*>(y)[i];
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        yv[0] = b.x; yv[1] = b.y; yv[2] = b.z; yv[3] = b.w;
        if (tg) {
          float4 c = reinterpret_cast<const float4*>(tg)[i];
          tv[0] = c.x; tv[1] = c.y; tv[2] = c.z; tv[3] = c.w;
        }
      } else if constexpr (VEC == 2) {
        float2 a = reinterpret_cast<const float2*>(x)[i], b = reinterpret_cast<const float2*>(y)[i];
        xv[0] = a.x; xv[1] = a.y; yv[0] = b.x; yv[1] = b.y;
        if (tg) {
          float2 c = reinterpret_cast<const float2*>(tg)[i];
          tv[0] = c.x; tv[1] = c.y;
        }
      } else {
        xv[0] = x[i]; yv[0] = y[i];
        if (tg) tv[0] = tg[i];
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float a = xv[k] * p.inv_temp, b = yv[k] * p.inv_temp;
        const float logp = a - lse_a;
        const float bq = b - lb.m;
        const float q = __expf(bq) * inv_sb;
        const float logq = bq - log_sb;
        if (q > 0.f) kl_row += q * (logq - logp);  // xlogy convention of F.kl_div
        float grad = p.kl_grad_scale * (__expf(logp) - q);
        if (tg) {
          const float xx = xv[k];
          const float e = __expf(-fabsf(xx));
          task_row += fmaxf(xx, 0.f) - xx * tv[k] + log1pf(e);
          const float sig = xx >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
          grad += p.task_weight * p.task_scale * (sig - tv[k]);
        }
        gv[k] = grad;
      }
      if (dx) {
        if constexpr (VEC == 4)
          reinterpret_cast<float4*>(dx)[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
        else if constexpr (VEC == 2)
          reinterpret_cast<float2*>(dx)[i] = make_float2(gv[0], gv[1]);
        else
          dx[i] = gv[0];
      }
    }
    kl_acc += kl_row;
    task_acc += task_row;
  }

  // block-level reduction of the partial sums, then one atomic triple per CTA
  float kl = group_sum<256>(kl_acc, scratch);
  float task = group_sum<256>(task_acc, scratch);
  if (threadIdx.x == 0) {
    kl *= p.kl_row_scale;
    task *= p.task_scale;
    atomicAdd(p.loss_out + 0, p.kl_weight * kl + p.task_weight * task);
    atomicAdd(p.loss_out + 1, kl);
    atomicAdd(p.loss_out + 2, task);
  }
}

template <int GROUP, int VEC>
int launch(const MkdParams& p, int sms, cudaStream_t st) {
  const int groups_per_block = 256 / GROUP;
  int64_t blocks = (p.rows + groups_per_block - 1) / groups_per_block;
  const int64_t cap = static_cast<int64_t>(sms) * 8;
  if (blocks > cap) blocks = cap;
  mkd_loss_kernel<GROUP, VEC><<<static_cast<int>(blocks), 256, 0, st>>>(p);
  FD_CHECK_CUDA(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int feddat_mkd_loss(const float* logits, const float* teacher, const float* target,
                               float* loss_out, float* dlogits, int64_t rows, int C, float temp,
                               float kl_weight, float task_weight, float task_scale,
                               int64_t batchmean_div, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(logits && teacher && loss_out, FD_ERR_INVALID, "mkd_loss: null pointer argument");
  FD_REQUIRE(rows >= 0 && C >= 1, FD_ERR_INVALID, "mkd_loss: bad shape rows=%lld C=%d",
             (long long)rows, C);
  FD_REQUIRE(temp > 0.f && batchmean_div > 0, FD_ERR_INVALID,
             "mkd_loss: temp and batchmean_div must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FD_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), st));
  if (rows == 0) return FD_OK;
  MkdParams p{};
  p.logits = logits; p.teacher = teacher; p.target = target; p.loss_out = loss_out;
  p.dlogits = dlogits; p.rows = rows; p.C = C;
  p.inv_temp = 1.f / temp;
  p.kl_row_scale = temp * temp / static_cast<float>(batchmean_div);
  p.kl_grad_scale = kl_weight * temp / static_cast<float>(batchmean_div);
  p.kl_weight = kl_weight; p.task_weight = task_weight; p.task_scale = task_scale;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  auto aligned = [&](int v) {
    auto ok = [&](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % (4 * v)) == 0; };
    return C % v == 0 && ok(logits) && ok(teacher) && ok(target) && ok(dlogits);
  };
  const int vec = aligned(4) ? 4 : (aligned(2) ? 2 : 1);
  if (C <= 2048) {
    if (vec == 4) return launch<32, 4>(p, sms, st);
    if (vec == 2) return launch<32, 2>(p, sms, st);
    return launch<32, 1>(p, sms, st);
  }
  if (vec == 4) return launch<256, 4>(p, sms, st);
  if (vec == 2) return launch<256, 2>(p, sms, st);
  return launch<256, 1>(p, sms, st);
}
