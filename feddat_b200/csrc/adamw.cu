// AdamW over the trainable parameters of one optimizer step (reference create_optimizer,
// src/train/visionlanguage_tasks/task_trainer.py:477-504: torch.optim.AdamW, betas (0.9, 0.98), decoupled weight
// decay on everything but biases / LayerNorm weights; stepped twice per batch, :303-308 and :323-328).
//
// The MKD step touches ~50 small tensors per optimizer step (12 sites x {down, up} x {weight, bias} plus the task
// head).  torch's fused multi-tensor AdamW gives each 65 536-element chunk one block: 48-70 blocks on 148 SMs, 40-70 us
// per step for 66-230 MB of traffic.  Here the tensor list travels by value in the kernel parameters (nothing to
// upload, capturable in a CUDA graph), chunks are 4 096 elements, and one launch covers both weight-decay groups.
// Arithmetic = torch's fused kernel (ATen/native/cuda/fused_adam_utils.cuh), fp32:
//   p -= lr wd p;  m += (1 - b1)(g - m);  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with t = step + 1 read from the parameter's own device-side step counter, which a second tiny launch increments.
#include <math.h>

#include "feddat_b200.h"
#include "host_common.h"

namespace fd {
namespace {

constexpr int kMaxTensors = 48;
constexpr int kChunk = 4096;

struct AdamwList {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  float* step[kMaxTensors];
  const float* lr[kMaxTensors];
  float wd[kMaxTensors];
  int chunk_start[kMaxTensors + 1];     // prefix sums of ceil(numel / kChunk)
  long long numel[kMaxTensors];
  int n;
};

__global__ void __launch_bounds__(256)
adamw_kernel(const __grid_constant__ AdamwList L, float beta1, float beta2, float eps) {
  int lo = 0, hi = L.n;                 // tensor of this block's chunk
  const int c = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (L.chunk_start[mid] <= c) lo = mid; else hi = mid;
  }
  const int t = lo;
  const long long base = static_cast<long long>(c - L.chunk_start[t]) * kChunk, n = L.numel[t];
  float* __restrict__ p = L.p[t];
  const float* __restrict__ g = L.g[t];
  float* __restrict__ m = L.m[t];
  float* __restrict__ v = L.v[t];
  const float lr = *L.lr[t], wd = L.wd[t], stepn = *L.step[t] + 1.f;
  const float bc1 = 1.f - powf(beta1, stepn), bc2 = 1.f - powf(beta2, stepn);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2), decay = lr * wd;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    pp -= decay * pp;
    mm += (1.f - beta1) * (gg - mm);
    vv = beta2 * vv + (1.f - beta2) * gg * gg;
    pp -= step_size * mm / (sqrtf(vv) / bc2_sqrt + eps);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
#pragma unroll
  for (int it = 0; it < kChunk / 1024; ++it) {
    const long long i = base + it * 1024 + threadIdx.x * 4;
    if (vec && i + 3 < n) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      upd(pp.x, gg.x, mm.x, vv.x);
      upd(pp.y, gg.y, mm.y, vv.y);
      upd(pp.z, gg.z, mm.z, vv.z);
      upd(pp.w, gg.w, mm.w, vv.w);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int k = 0; k < 4; ++k)
        if (i + k < n) upd(p[i + k], g[i + k], m[i + k], v[i + k]);
    }
  }
}

__global__ void adamw_bump_kernel(const __grid_constant__ AdamwList L) {
  const int t = threadIdx.x;
  if (t < L.n) *L.step[t] += 1.f;
}

}  // namespace
}  // namespace fd

extern "C" int feddat_adamw_step(const FeddatAdamwTensor* tensors, int n, float beta1, float beta2, float eps, void* stream) {
  using namespace fd;
  int rc = check_device_sm100();
  if (rc) return rc;
  FD_REQUIRE(n >= 0 && (n == 0 || tensors != nullptr), FD_ERR_INVALID, "adamw_step: bad tensor list");
  auto st = static_cast<cudaStream_t>(stream);
  for (int i0 = 0; i0 < n; i0 += kMaxTensors) {
    AdamwList L{};
    L.n = n - i0 < kMaxTensors ? n - i0 : kMaxTensors;
    int chunks = 0;
    for (int t = 0; t < L.n; ++t) {
      const FeddatAdamwTensor& T = tensors[i0 + t];
      FD_REQUIRE(T.param && T.grad && T.exp_avg && T.exp_avg_sq && T.step && T.lr && T.numel >= 0, FD_ERR_INVALID,
                 "adamw_step: tensor %d has a null pointer or a negative size", i0 + t);
      FD_REQUIRE(((reinterpret_cast<uintptr_t>(T.param) | reinterpret_cast<uintptr_t>(T.grad) | reinterpret_cast<uintptr_t>(T.exp_avg) |
                   reinterpret_cast<uintptr_t>(T.exp_avg_sq)) & 3) == 0, FD_ERR_INVALID, "adamw_step: tensor %d is not fp32-aligned", i0 + t);
      L.p[t] = static_cast<float*>(T.param); L.g[t] = static_cast<const float*>(T.grad);
      L.m[t] = static_cast<float*>(T.exp_avg); L.v[t] = static_cast<float*>(T.exp_avg_sq);
      L.step[t] = T.step; L.lr[t] = T.lr; L.wd[t] = T.weight_decay; L.numel[t] = T.numel;
      L.chunk_start[t] = chunks;
      chunks += static_cast<int>((T.numel + kChunk - 1) / kChunk);
    }
    L.chunk_start[L.n] = chunks;
    if (chunks > 0) {
      adamw_kernel<<<chunks, 256, 0, st>>>(L, beta1, beta2, eps);
      FD_CHECK_CUDA(cudaGetLastError());
    }
    adamw_bump_kernel<<<1, kMaxTensors, 0, st>>>(L);
    FD_CHECK_CUDA(cudaGetLastError());
  }
  return FD_OK;
}
