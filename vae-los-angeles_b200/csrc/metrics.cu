// Reconstruction metrics in one pass over (y_true, y_pred): what compute_metrics of the reference
// (compare_directional_imputation.py:167-210) obtains from five scikit-learn / scipy calls and an N x N cosine matrix --
// MAE, MSE, RMSE, R2 over the flattened arrays, cosine similarity and Pearson r per sample, their mean / std.
//
// HBM-bound streaming reduction: 8 B of input per element, nothing else.  One warp per row (rows are 2-3 KB), 128-bit or
// 64-bit loads with 128 bytes in flight per lane and array, seven fp32 lane sums per row reduced by an fp32 butterfly, all cross-row sums in double, per-block partials, and a fixed-order final reduction by the last block (deterministic).
#include "vla_internal.h"

#include <cmath>

namespace vla {

namespace {

constexpr int MT_THREADS = 256;
constexpr int MT_WARPS = MT_THREADS / 32;

struct Sums { float sx, sy, sxy, sxx, syy, sad, ssd; };
__device__ __forceinline__ void acc1(Sums& s, float x, float y) {
  const float d = y - x;
  s.sx += x; s.sy += y; s.sxy = fmaf(x, y, s.sxy); s.sxx = fmaf(x, x, s.sxx); s.syy = fmaf(y, y, s.syy);
  s.sad += fabsf(d); s.ssd = fmaf(d, d, s.ssd);
}
// fp32 butterfly (5 shuffles); the caller continues in double.  A row has a few hundred elements, so the fp32 row sums
// carry ~1e-7 relative error; everything that is accumulated ACROSS rows is kept in double.
__device__ __forceinline__ double warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return static_cast<double>(v);
}

// VEC: floats per load (4: dim % 4 == 0 and 16-byte aligned bases; 2: dim % 2 == 0 and 8-byte aligned; 1: anything).
// One instantiation per width keeps each at <= 64 registers (four blocks = 32 warps per SM, 4 KB in flight per warp).
template <int VEC>
__global__ void __launch_bounds__(MT_THREADS, 4) metrics_kernel(const MetricsArgs a) {
  __shared__ double sh[MT_WARPS][8];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * MT_WARPS + warp;
  const long long nw = static_cast<long long>(gridDim.x) * MT_WARPS;
  const int D = a.dim;
  constexpr int U = VEC == 4 ? 4 : 8;      // row segments in flight per lane and array (128 B per lane either way)
  if (lane < 8) sh[warp][lane] = 0.0;      // this warp's running sums over its rows (lane 0 accumulates; kept out of registers)
  __syncwarp();
  for (long long row = gw; row < a.rows; row += nw) {
    const float* __restrict__ t = a.yt + row * D;
    const float* __restrict__ p = a.yp + row * D;
    Sums s{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (VEC == 4) {
      const float4* t4 = reinterpret_cast<const float4*>(t);
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const int n4 = D >> 2;
      for (int i = lane; i < n4; i += 32 * U) {
        float4 x[U], y[U];
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + 32 * k < n4) { x[k] = __ldcs(t4 + i + 32 * k); y[k] = __ldcs(p4 + i + 32 * k); }
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (i + 32 * k < n4) { acc1(s, x[k].x, y[k].x); acc1(s, x[k].y, y[k].y); acc1(s, x[k].z, y[k].z); acc1(s, x[k].w, y[k].w); }
      }
    } else if (VEC == 2) {
      const float2* t2 = reinterpret_cast<const float2*>(t);
      const float2* p2 = reinterpret_cast<const float2*>(p);
      const int n2 = D >> 1;
      for (int i = lane; i < n2; i += 32 * U) {
        float2 x[U], y[U];
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + 32 * k < n2) { x[k] = __ldcs(t2 + i + 32 * k); y[k] = __ldcs(p2 + i + 32 * k); }
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + 32 * k < n2) { acc1(s, x[k].x, y[k].x); acc1(s, x[k].y, y[k].y); }
      }
    } else {
      for (int i = lane; i < D; i += 32 * U) {
        float x[U], y[U];
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + 32 * k < D) { x[k] = __ldcs(t + i + 32 * k); y[k] = __ldcs(p + i + 32 * k); }
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + 32 * k < D) acc1(s, x[k], y[k]);
      }
    }
    const double sx = warp_sum(s.sx), sy = warp_sum(s.sy), sxy = warp_sum(s.sxy), sxx = warp_sum(s.sxx), syy = warp_sum(s.syy);
    const double sad = warp_sum(s.sad), ssd = warp_sum(s.ssd);
    if (lane == 0) {
      // cosine: rows L2-normalised, a zero row stays zero (sklearn.preprocessing.normalize)
      const double nt = sqrt(sxx), np_ = sqrt(syy);
      const double cs = sxy / ((nt == 0 ? 1.0 : nt) * (np_ == 0 ? 1.0 : np_));
      // Pearson: NaN for a constant row (scipy.stats.pearsonr), clipped to [-1, 1]
      const double vt = sxx - sx * sx / D, vp = syy - sy * sy / D;
      // a constant row has vt == 0 in exact arithmetic; the fp32 lane sums leave rounding residue of relative size ~1e-7
      const bool ok = vt > 1e-6 * sxx + 1e-300 && vp > 1e-6 * syy + 1e-300;
      double r = nan("");
      if (ok) { r = (sxy - sx * sy / D) / sqrt(vt * vp); r = fmin(1.0, fmax(-1.0, r)); }
      if (a.cos_out) a.cos_out[row] = static_cast<float>(cs);
      if (a.pearson_out) a.pearson_out[row] = static_cast<float>(r);
      sh[warp][0] += sad; sh[warp][1] += ssd; sh[warp][2] += sx; sh[warp][3] += sxx; sh[warp][4] += cs;
      if (ok) { sh[warp][5] += r; sh[warp][6] += r * r; sh[warp][7] += 1.0; }
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0;
    for (int w = 0; w < MT_WARPS; ++w) v += sh[w][threadIdx.x];
    a.partials[static_cast<size_t>(blockIdx.x) * 8 + threadIdx.x] = v;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 8) {
    double v = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(a.partials + static_cast<size_t>(b) * 8 + threadIdx.x);   // fixed order
    sh[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n_el = static_cast<double>(a.rows) * D;
    const double mae = sh[0][0] / n_el, mse = sh[0][1] / n_el;
    const double ss_tot = sh[0][3] - sh[0][2] * sh[0][2] / n_el;
    double r2;
    if (ss_tot > 1e-12 * sh[0][3]) r2 = 1.0 - sh[0][1] / ss_tot;
    else r2 = sh[0][1] == 0 ? 1.0 : 0.0;                                     // sklearn's force_finite convention
    const double cnt = sh[0][7];
    const double pm = cnt > 0 ? sh[0][5] / cnt : 0.0;
    const double pv = cnt > 0 ? fmax(sh[0][6] / cnt - pm * pm, 0.0) : 0.0;
    a.out[0] = mae; a.out[1] = mse; a.out[2] = sqrt(mse); a.out[3] = r2;
    a.out[4] = sh[0][4] / static_cast<double>(a.rows); a.out[5] = pm; a.out[6] = sqrt(pv); a.out[7] = cnt;
    *a.counter = 0u;                                                         // re-arm (graph replays)
  }
}

}  // namespace

int metrics_grid(long long rows) {
  long long b = (rows + MT_WARPS - 1) / MT_WARPS;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return static_cast<int>(b);
}

cudaError_t launch_metrics(const MetricsArgs& a, cudaStream_t s) {
  const uintptr_t bases = reinterpret_cast<uintptr_t>(a.yt) | reinterpret_cast<uintptr_t>(a.yp);
  const dim3 grid(metrics_grid(a.rows)), block(MT_THREADS);
  if (a.dim % 4 == 0 && bases % 16 == 0) return launch_pdl(metrics_kernel<4>, grid, block, 0, s, a);
  if (a.dim % 2 == 0 && bases % 8 == 0) return launch_pdl(metrics_kernel<2>, grid, block, 0, s, a);
  return launch_pdl(metrics_kernel<1>, grid, block, 0, s, a);
}

}  // namespace vla
