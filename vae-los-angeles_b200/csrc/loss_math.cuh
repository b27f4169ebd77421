// Per-element loss terms shared by the stand-alone loss kernel (elementwise_dev.cuh) and the loss-fused GEMM epilogue
// (gemm_tile.cuh): value, dL/dy and dL/d(pre-activation) of the MSE-sum and BCE-sum terms
// (src/utils/losses.py:31-35, src/utils/directional_losses.py:23-24, 48-49 of the reference; ATen's BCE clamps).
#pragma once
#include <cuda_runtime.h>

namespace vla {

template <bool BCE>
__device__ __forceinline__ float loss_elem(float y, float t, float gs, float& g_out, float& g_logit) {
  if (BCE) {
    // lg2.approx-based logs (absolute error ~1e-7 per term, far below the 1e-5 relative budget of the summed loss)
    const float ly = fmaxf(__logf(y), -100.0f);
    const float l1 = fmaxf(__logf(1.0f - y), -100.0f);
    const float yy = y * (1.0f - y);
    g_out = __fdividef(y - t, fmaxf(yy, 1e-12f)) * gs;   // dL/dy (ATen's backward floor)
    g_logit = g_out * yy;                           // dL/d(pre-sigmoid)
    return -(t * ly + (1.0f - t) * l1);
  } else {
    const float d = y - t;
    g_out = 2.0f * d * gs;
    g_logit = g_out;
    return d * d;
  }
}

// BCE-sum term evaluated from the pre-sigmoid value x (the fused epilogue has it): with y = sigmoid(x),
//   -(t log y + (1 - t) log(1 - y)) = max(x, 0) - t x + log1p(exp(-|x|))   and   dL/dx = y - t,
// identical to the reference's binary_cross_entropy(sigmoid(x), t) wherever its log clamp (-100) and backward floor (1e-12)
// are inactive.  In fp32 that is -27.6 < x < 16.6: above 16.6 sigmoid(x) rounds to exactly 1.0, the reference's log(1 - y)
// hits the clamp (loss term 100 (1 - t) instead of (1 - t) x) and its gradient through the sigmoid becomes 0 instead of
// y - t; below -27.6 y (1 - y) falls under the backward floor and the reference's gradient shrinks by y (1 - y) / 1e-12.  The
// fused train step keeps the exact expression (a saturated logit still gets its true gradient); the functional path
// (loss_elem<true>) follows ATen to the letter.  Logits of this model are O(1) at initialisation and during training; the
// difference for saturated logits is deliberate and pinned by tests/test_gpu_parity.py::test_saturated_logits_fused_vs_functional.
// 3 MUFU + ~10 ALU operations per element instead of ~40.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Branch-free (the 32 elements of a row chunk then interleave in the instruction stream): single MUFU instructions with
// 2^-22 relative error, operands in (0, 2] so no range handling is needed.
__device__ __forceinline__ float bce_from_logit(float x, float t, float& y, float& g_logit) {
  const float e = ex2_approx(-fabsf(x) * 1.4426950408889634f);   // exp(-|x|) in (0, 1]
  const float r = rcp_approx(1.0f + e);
  y = x >= 0.f ? r : e * r;
  g_logit = y - t;
  return fmaxf(x, 0.f) - t * x + lg2_approx(1.0f + e) * 0.6931471805599453f;
}

}  // namespace vla
