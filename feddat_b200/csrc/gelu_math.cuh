// Exact-GELU terms shared by the streaming kernels (gelu.cu) and the fused GEMM epilogues (mlp_gemm.cu).
#pragma once
namespace fd {
// (Phi(x), exp(-x^2 / 2)) with Phi from the complementary error function in the Abramowitz-Stegun 7.1.26
// form (|abs error| < 1e-7): one MUFU.RCP, one MUFU.EX2 and ~12 FMAs per element; see gelu.cu.
__device__ __forceinline__ void gelu_terms(float x, float& cdf, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  // rcp.approx / ex2.approx (1-2 ulp): the IEEE-rounded forms compile to slow-path calls and ~40
  // instructions per element
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  float pl = fmaf(t, 1.061405429f, -1.453152027f);
  pl = fmaf(t, pl, 1.421413741f);
  pl = fmaf(t, pl, -0.284496736f);
  pl = fmaf(t, pl, 0.254829592f);
  const float q = 0.5f * t * pl * e;
  cdf = x >= 0.f ? 1.f - q : q;
}
__device__ __forceinline__ float gelu_fwd_f(float x) {
  float cdf, e;
  gelu_terms(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, e;
  gelu_terms(x, cdf, e);
  return fmaf(x, 0.3989422804014327f * e, cdf);
}

// ---- two elements at a time (sm_100 packed fp32: FFMA2 / FMUL2 / FADD2 issue once for a pair) -------------------
// The fused GEMM epilogues are issue-bound, so the same A-S 7.1.26 evaluation is arranged for the fewest issue
// slots per element (8.5 forward, 10.5 backward against ~18 scalar):
//   t = 1 / (1 + (p / sqrt 2) |x|),  e = exp(-x^2 / 2) = ex2(x^2 * (-log2(e) / 2))
//   h = Phi(|x|) - 1/2 = 1/2 - q = fma(t * P(t), e, 1/2)        with P's coefficients pre-multiplied by -1/2
//   gelu(x)  = x / 2 + |x| h                                      (= x (1 - q) for x >= 0, x q for x < 0)
//   gelu'(x) = 1/2 + copysign(h, x) + x e / sqrt(2 pi)
__device__ __forceinline__ float2 gm_f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ void gelu_terms2(const float2 x, float2& h, float2& e) {
  float2 t, arg;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(0.2316418882f, fabsf(x.x), 1.f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(0.2316418882f, fabsf(x.y), 1.f)));
  arg = __fmul2_rn(__fmul2_rn(x, x), gm_f2(-0.7213475204444817f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
  float2 pl = __ffma2_rn(t, gm_f2(-0.5307027145f), gm_f2(0.7265760135f));
  pl = __ffma2_rn(t, pl, gm_f2(-0.7107068705f));
  pl = __ffma2_rn(t, pl, gm_f2(0.142248368f));
  pl = __ffma2_rn(t, pl, gm_f2(-0.127414796f));
  h = __ffma2_rn(__fmul2_rn(t, pl), e, gm_f2(0.5f));
}
__device__ __forceinline__ float2 gelu_fwd2(const float2 x) {
  float2 h, e;
  gelu_terms2(x, h, e);
  const float2 half_x = __fmul2_rn(x, gm_f2(0.5f));
  return make_float2(fmaf(fabsf(x.x), h.x, half_x.x), fmaf(fabsf(x.y), h.y, half_x.y));
}
__device__ __forceinline__ float2 gelu_grad2(const float2 x) {
  float2 h, e;
  gelu_terms2(x, h, e);
  const float2 s = __fadd2_rn(make_float2(copysignf(h.x, x.x), copysignf(h.y, x.y)), gm_f2(0.5f));
  return __ffma2_rn(__fmul2_rn(x, e), gm_f2(0.3989422804014327f), s);
}
}  // namespace fd
