// Head block: the small layers around the latent as one launch each way, on the CUDA cores (vla_internal.h "Head block").
//
// Forward  = encoders.py:12-19 tail (BatchNorm1d -> ReLU -> Dropout of the LAST hidden layer), :18-23 / :40-46 / :54-61 (fc_mu,
//            fc_logvar; Embedding -> heads), vae.py:11-15, 64-73 (mean over modalities, reparameterize), decoders.py first
//            Linear + ReLU of every decoder (fused [sum of widths, L] matrix), losses.py:42 (KL partial).
// Backward = the same lines through autograd: d(first decoder layer) -> d(mu, logvar) + beta dKL -> d(heads) with the ReLU /
//            dropout mask and the BatchNorm backward statistics of the row block; embedding-row gradient of the site encoder.
//
// A block owns 32 rows; thread = (row, group of output columns): activations of the rows sit in shared memory as fp32 with a
// row pitch whose quarter is odd (conflict-free 128-bit reads by the 32 rows of a warp), the weights of the layer in flight sit
// in shared memory in the orientation that makes the reduction index contiguous (128-bit broadcast reads).  Everything is fp32
// (master weights): these layers feed ReLUs / the latent, where operand rounding costs gradient accuracy (DESIGN.md "Precision").
#include "elementwise_dev.cuh"
#include "tc_ptx.cuh"
#include "vla_internal.h"

namespace vla {

namespace {

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }
// row pitch (floats) >= w, multiple of 4, with pitch / 4 odd: 32 rows x 16-byte reads hit every bank exactly once per 8 rows
__host__ __device__ inline int pitch_of(int w) { const int p = pad4(w); return ((p >> 2) & 1) ? p : p + 4; }

struct HbLayout {       // offsets in floats
  int hs, hs_ld, kh;            // forward: activations of all encoders side by side [32][hs_ld]; backward: g_d0 rows [32][hs_ld]
  int wh;                       // heads weights: forward [HW][in] per encoder back to back; backward transposed [in][HWq]
  int ml, ml_ld;                // forward: sum of the heads [32][ml_ld]; backward: d(mu | logvar) [32][ml_ld]
  int zs, zs_ld;                // z / dL/dz [32][zs_ld]
  int w0, lq;                   // forward: W0 [C][lq]; backward: W0 transposed [L][C]
  int vec;                      // per-column vectors: [3][vec_n] (mean | rstd * gamma or rstd | beta)
  int vec_n;
  int red;                      // block reduction scratch (64 floats)
  int total;
};
__host__ __device__ inline HbLayout hb_layout(const HbArgs& a, bool backward) {
  HbLayout L;
  int kh = 0, nmax = 4;
  for (int e = 0; e < a.n_enc; ++e) { kh += a.enc[e].in_dim; if (a.enc[e].kind == 0 && a.enc[e].in_dim > nmax) nmax = a.enc[e].in_dim; }
  L.kh = kh;
  const int hwq = pad4(a.HW);
  int o = 0;
  const int wide = backward ? (a.C > nmax ? a.C : nmax) : kh;      // backward: g_d0 rows, later reused for one encoder's pre-activations
  L.hs = o; L.hs_ld = pitch_of(wide); o += HB_ROWS * L.hs_ld;
  L.wh = o; o += backward ? kh * hwq : a.HW * kh;
  L.ml = o; L.ml_ld = pitch_of(a.HW); o += HB_ROWS * L.ml_ld;
  L.zs = o; L.zs_ld = pitch_of(a.L); o += HB_ROWS * L.zs_ld;
  L.lq = pad4(a.L);
  L.w0 = o; o += backward ? a.L * a.C : a.C * L.lq;
  L.vec_n = nmax; L.vec = o; o += 3 * (backward ? nmax : kh);
  L.red = o; o += 64;
  L.total = o;
  return L;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// column sums over the 32 lanes of v[0..15]: lane j (< 16) ends up with the total of column j
__device__ __forceinline__ float col_sums16(float (&v)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = upper ? v[j] : v[j + s];
      const float keep = upper ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

extern __shared__ __align__(16) float hb_smem[];

// straight copies into shared memory travel as asynchronous 16-byte copies (no register round trip: a load followed by its
// store would serialise the loop on the memory latency, one round trip per iteration)
__device__ __forceinline__ void hb_cp16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void hb_cp_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}


// acc[q] += sum_k x[k] * w_q[k] for Q outputs: x = this thread's row (128-bit reads, conflict-free pitch), w_q = Q weight rows
// (128-bit broadcast reads: every lane of the warp reads the same address).  Q is a template parameter: a runtime bound inside
// an unrolled loop leaves predicated-off FMAs in the instruction stream (measured: 2.7x the useful instructions).
template <int Q>
__device__ __forceinline__ void dot_rows(const float* __restrict__ x, const float* __restrict__ w, int w_pitch, int K, float (&acc)[8]) {
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    const float4 h = *reinterpret_cast<const float4*>(x + k);
    float4 wv[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) wv[q] = *reinterpret_cast<const float4*>(w + q * w_pitch + k);
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = fmaf(h.x, wv[q].x, fmaf(h.y, wv[q].y, fmaf(h.z, wv[q].z, fmaf(h.w, wv[q].w, acc[q]))));
  }
}
__device__ __forceinline__ void dot_rows_n(int Q, const float* x, const float* w, int w_pitch, int K, float (&acc)[8]) {
  switch (Q) {
    case 1: dot_rows<1>(x, w, w_pitch, K, acc); break;
    case 2: dot_rows<2>(x, w, w_pitch, K, acc); break;
    case 3: dot_rows<3>(x, w, w_pitch, K, acc); break;
    case 4: dot_rows<4>(x, w, w_pitch, K, acc); break;
    case 5: dot_rows<5>(x, w, w_pitch, K, acc); break;
    case 6: dot_rows<6>(x, w, w_pitch, K, acc); break;
    case 7: dot_rows<7>(x, w, w_pitch, K, acc); break;
    case 8: dot_rows<8>(x, w, w_pitch, K, acc); break;
    default: break;
  }
}

// =============================================================================================
// forward
// =============================================================================================
__global__ void __launch_bounds__(HB_THREADS) head_block_fwd_kernel(const __grid_constant__ HbArgs a) {
  const HbLayout Y = hb_layout(a, false);
  float* const sm = hb_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * HB_ROWS;
  const int L = a.L, HW = a.HW, C = a.C;
  pdl_wait();
  pdl_launch_dependents();
  const long long ds_row0 = a.n_batches > 1 ? static_cast<long long>(a.dyn->batch_index % a.n_batches) * a.rows : 0;
  const unsigned long long step_off = a.dyn ? static_cast<unsigned long long>(__ldcg(&a.dyn->step)) << 20 : 0ull;

  // ---- stage the weights (heads of every encoder as stored, [HW][in]; W0 [C][L] zero padded to lq) ----
  {
    int off = 0;
    for (int e = 0; e < a.n_enc; ++e) {
      const int n4 = HW * a.enc[e].in_dim / 4;
      const float* src = a.enc[e].Wh;
      float* dst = sm + Y.wh + off;
      for (int i = tid; i < n4; i += HB_THREADS) hb_cp16(dst + 4 * i, src + 4 * i);
      off += HW * a.enc[e].in_dim;
    }
    if (Y.lq == L) {
      for (int i = tid; i < C * L / 4; i += HB_THREADS) hb_cp16(sm + Y.w0 + 4 * i, a.W0 + 4 * i);
    } else {
#pragma unroll 8
      for (int i = tid; i < C * Y.lq; i += HB_THREADS) {
        const int c = i / Y.lq, l = i - c * Y.lq;
        sm[Y.w0 + i] = l < L ? __ldg(a.W0 + c * L + l) : 0.f;
      }
    }
  }
  // ---- BatchNorm statistics of the whole batch from the per-tile column sums (every block; block 0 publishes them) ----
  // All 512 threads take part: thread = (column, slice of the tiles), every load of a thread in flight at once, the slices
  // are combined in tile order through shared memory (ml / zs regions are free until the heads run).
  {
    int voff = 0;
    double* part = reinterpret_cast<double*>(sm + Y.hs);       // [slices][2][n] doubles: the activation region is free now
    for (int e = 0; e < a.n_enc; ++e) {
      const HbEnc& E = a.enc[e];
      if (E.kind == 0) {
        const int n = E.in_dim;
        if (E.train) {
          const int slices = HB_THREADS / n > 0 ? HB_THREADS / n : 1;          // n <= 256: 2 .. 8 slices
          const int per = (E.m_tiles + slices - 1) / slices;
          for (int idx = tid; idx < n * slices; idx += HB_THREADS) {
            const int col = idx % n, sl = idx / n;
            double s1 = 0, s2 = 0;
            for (int t0 = sl * per; t0 < min(E.m_tiles, (sl + 1) * per); t0 += 8) {
              float p1[8], p2[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const bool ok = t0 + q < min(E.m_tiles, (sl + 1) * per);
                p1[q] = ok ? __ldcg(E.stats + (static_cast<size_t>(t0 + q) * 2 + 0) * n + col) : 0.f;
                p2[q] = ok ? __ldcg(E.stats + (static_cast<size_t>(t0 + q) * 2 + 1) * n + col) : 0.f;
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) { s1 += p1[q]; s2 += p2[q]; }
            }
            part[(sl * 2 + 0) * n + col] = s1;
            part[(sl * 2 + 1) * n + col] = s2;
          }
          __syncthreads();
          for (int col = tid; col < n; col += HB_THREADS) {
            double s1 = 0, s2 = 0;
            for (int sl = 0; sl < slices; ++sl) { s1 += part[(sl * 2 + 0) * n + col]; s2 += part[(sl * 2 + 1) * n + col]; }
            const double m = s1 / a.rows;
            double var = s2 / a.rows - m * m;
            var = var < 0 ? 0 : var;
            const float mean = static_cast<float>(m);
            const float rstd = rsqrtf(static_cast<float>(var) + 1e-5f);
            if (blockIdx.x == 0) {
              const double unbiased = a.rows > 1 ? var * a.rows / (a.rows - 1) : var;
              E.running_mean[col] = 0.9f * E.running_mean[col] + 0.1f * mean;
              E.running_var[col] = 0.9f * E.running_var[col] + 0.1f * static_cast<float>(unbiased);
              E.save_mean[col] = mean; E.save_rstd[col] = rstd;
            }
            sm[Y.vec + voff + col] = mean;
            sm[Y.vec + Y.kh + voff + col] = rstd * E.gamma[col];
            sm[Y.vec + 2 * Y.kh + voff + col] = E.beta[col];
          }
          __syncthreads();                                     // `part` is reused by the next encoder / overwritten by hs
        } else {
          for (int col = tid; col < n; col += HB_THREADS) {
            const float mean = E.running_mean[col];
            const float rstd = 1.0f / sqrtf(E.running_var[col] + 1e-5f);
            sm[Y.vec + voff + col] = mean;
            sm[Y.vec + Y.kh + voff + col] = rstd * E.gamma[col];
            sm[Y.vec + 2 * Y.kh + voff + col] = E.beta[col];
            if (blockIdx.x == 0) { E.save_mean[col] = mean; E.save_rstd[col] = rstd; }
          }
        }
        if (blockIdx.x == 0 && tid == 0 && E.train && E.nbt) *E.nbt += 1;
      }
      voff += E.in_dim;
    }
  }
  hb_cp_wait();
  __syncthreads();
  // ---- activations of every encoder into hs ----
  {
    int off = 0;
    for (int e = 0; e < a.n_enc; ++e) {
      const HbEnc& E = a.enc[e];
      const int n = E.in_dim;
      if (E.kind == 0) {
        const bool drop = E.train && a.p_drop > 0.f;
        const float keep_scale = drop ? 1.0f / (1.0f - a.p_drop) : 1.0f;
        const unsigned long long offset = E.drop_offset + step_off;
        // lane = column pair of a 64-column block, warp = rows {w, w + 16}
        for (int cb = 0; cb < n; cb += 64) {
          const int col = cb + lane * 2;
          const bool col_ok = col < n;               // n is a multiple of 64 on this path (host), kept for safety
          const float m0 = sm[Y.vec + off + (col_ok ? col : 0)], m1 = sm[Y.vec + off + (col_ok ? col + 1 : 0)];
          const float s0 = sm[Y.vec + Y.kh + off + (col_ok ? col : 0)], s1 = sm[Y.vec + Y.kh + off + (col_ok ? col + 1 : 0)];
          const float b0 = sm[Y.vec + 2 * Y.kh + off + (col_ok ? col : 0)], b1 = sm[Y.vec + 2 * Y.kh + off + (col_ok ? col + 1 : 0)];
          constexpr int NW = HB_THREADS / 32, RPW = HB_ROWS / NW;     // rows per warp: {w, w + NW, ...}
          float2 x[RPW];
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            const int row = row0 + warp + NW * i;
            x[i] = (col_ok && row < a.rows) ? __ldg(reinterpret_cast<const float2*>(E.pre + static_cast<size_t>(row) * n + col)) : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            const int r = warp + NW * i, row = row0 + r;
            const bool ok = col_ok && row < a.rows;
            float y0 = fmaxf((x[i].x - m0) * s0 + b0, 0.f);
            float y1 = fmaxf((x[i].y - m1) * s1 + b1, 0.f);
            if (drop && ok) {
              bool k0, k1;
              if (E.keep_mask) {
                const uchar2 k = *reinterpret_cast<const uchar2*>(E.keep_mask + static_cast<size_t>(row) * n + col);
                k0 = k.x != 0; k1 = k.y != 0;
              } else {
                const unsigned long long idx = (static_cast<unsigned long long>(row) * n + col) >> 1;
                const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                                           static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                                make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32)));
                k0 = u01(rnd.x) >= a.p_drop; k1 = u01(rnd.y) >= a.p_drop;
              }
              y0 = k0 ? y0 * keep_scale : 0.f;
              y1 = k1 ? y1 * keep_scale : 0.f;
            }
            if (!ok) { y0 = 0.f; y1 = 0.f; }
            if (col_ok) { sm[Y.hs + r * Y.hs_ld + off + col] = y0; sm[Y.hs + r * Y.hs_ld + off + col + 1] = y1; }
            if (ok) *reinterpret_cast<__nv_bfloat162*>(E.act + static_cast<size_t>(row) * E.ld_act + col) = __floats2bfloat162_rn(y0, y1);
            if (E.bits) {
              const uint32_t be = __ballot_sync(0xffffffffu, y0 > 0.f), bo = __ballot_sync(0xffffffffu, y1 > 0.f);
              if (lane < 2 && row < a.rows) {
                const uint32_t e16 = lane ? (be >> 16) : (be & 0xFFFFu), o16 = lane ? (bo >> 16) : (bo & 0xFFFFu);
                E.bits[static_cast<size_t>((cb >> 5) + lane) * a.rows + row] = spread16(e16) | (spread16(o16) << 1);
              }
            }
          }
        }
      } else {
        // one label per lane first, then 128-bit pieces of the embedding rows (all loads of a thread before its stores)
        const long long lab = (row0 + lane < a.rows) ? __ldg(E.site + ds_row0 + row0 + lane) : 0;
        const int n4 = n / 4;
#pragma unroll 4
        for (int i = tid; i < HB_ROWS * n4; i += HB_THREADS) {
          const int r = i / n4, c4 = i - r * n4;
          const long long s_r = __shfl_sync(0xffffffffu, lab, r);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row0 + r < a.rows) v = __ldg(reinterpret_cast<const float4*>(E.emb + s_r * n) + c4);
          *reinterpret_cast<float4*>(sm + Y.hs + r * Y.hs_ld + off + 4 * c4) = v;
        }
      }
      off += n;
    }
  }
  __syncthreads();
  // ---- heads: thread = (row = lane, group of JPW output columns = warp); reduction index contiguous, 128-bit reads ----
  {
    constexpr int NW = HB_THREADS / 32;
    const int JPW = (HW + NW - 1) / NW;                 // <= 8 (HW <= 64)
    const int j0 = warp * JPW;
    const int nq = max(0, min(JPW, HW - j0));           // outputs of this warp (uniform)
    float tot[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) tot[q] = 0.f;
    int off = 0;
    for (int e = 0; e < a.n_enc; ++e) {
      const int n = a.enc[e].in_dim;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      dot_rows_n(nq, sm + Y.hs + lane * Y.hs_ld + off, sm + Y.wh + HW * off + j0 * n, n, n, acc);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nq) tot[q] += acc[q] + __ldg(a.enc[e].bh + j0 + q);       // every head adds its own bias
      off += n;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < nq) sm[Y.ml + lane * Y.ml_ld + j0 + q] = tot[q];
  }
  __syncthreads();
  // ---- latent: mean over the encoders, z = mu + eps * exp(logvar / 2), KL partial ----
  {
    const float inv_enc = 1.0f / a.n_enc;
    const unsigned long long offset = a.lat_offset + ((a.dyn && !a.eps_in) ? step_off : 0ull);
    float kl = 0.f;
    for (int i = tid; i < HB_ROWS * Y.zs_ld; i += HB_THREADS) {
      const int r = i / Y.zs_ld, j = i - r * Y.zs_ld;
      const int row = row0 + r;
      float z = 0.f;
      if (j < L && row < a.rows) {
        float mu = sm[Y.ml + r * Y.ml_ld + j];
        float lv = a.ae ? 0.f : sm[Y.ml + r * Y.ml_ld + L + j];
        if (a.n_enc > 1) { mu *= inv_enc; lv *= inv_enc; }
        const unsigned idx = static_cast<unsigned>(row) * L + j;
        if (a.ae) {
          z = mu;
          a.mu[idx] = mu; a.logvar[idx] = 0.f; a.eps_save[idx] = 0.f;
        } else {
          float eps;
          if (a.eps_in) eps = a.eps_in[idx];
          else {
            const uint4 rnd = philox4x32_10(make_uint4(idx, 0u, static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32) ^ 0x5EEDu),
                                            make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32)));
            eps = normal_from(rnd.x, rnd.y);
          }
          z = mu + eps * expf(0.5f * lv);
          a.mu[idx] = mu; a.logvar[idx] = lv; a.eps_save[idx] = eps;
          kl += 1.0f + lv - mu * mu - expf(lv);
        }
        a.z[static_cast<size_t>(row) * a.ld_z + j] = __float2bfloat16(z);
      }
      sm[Y.zs + i] = z;
    }
    kl = warp_sum_f(kl);
    if (lane == 0) sm[Y.red + warp] = kl;
  }
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < HB_THREADS / 32; ++w) t += sm[Y.red + w];
    a.kl_partials[blockIdx.x] = -0.5f * t;
  }
  // ---- fused first decoder layer: thread = (row = lane, 16-column piece(s) = warp), bias + ReLU -> bf16 hi (+ lo) + bits ----
  if (C > 0) {
    const float* zrow = sm + Y.zs + lane * Y.zs_ld;
    const int row = row0 + lane;
    for (int pc = warp; pc * 16 < C; pc += HB_THREADS / 32) {
      float acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll 2
      for (int l = 0; l < Y.lq; l += 4) {
        const float4 z4 = *reinterpret_cast<const float4*>(zrow + l);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(sm + Y.w0 + (pc * 16 + c) * Y.lq + l);
          acc[c] = fmaf(z4.x, w.x, fmaf(z4.y, w.y, fmaf(z4.z, w.z, fmaf(z4.w, w.w, acc[c]))));
        }
      }
      uint32_t bits = 0u, hi[8], lo[8];
      const float4* b4 = reinterpret_cast<const float4*>(a.b0 + pc * 16);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 bv = __ldg(b4 + c4);
        const float y0 = fmaxf(acc[4 * c4] + bv.x, 0.f), y1 = fmaxf(acc[4 * c4 + 1] + bv.y, 0.f);
        const float y2 = fmaxf(acc[4 * c4 + 2] + bv.z, 0.f), y3 = fmaxf(acc[4 * c4 + 3] + bv.w, 0.f);
        bits |= ((y0 > 0.f ? 1u : 0u) | (y1 > 0.f ? 2u : 0u) | (y2 > 0.f ? 4u : 0u) | (y3 > 0.f ? 8u : 0u)) << (4 * c4);
        hi[2 * c4] = pack_bf16x2_hi_lo(y0, y1, lo[2 * c4]);
        hi[2 * c4 + 1] = pack_bf16x2_hi_lo(y2, y3, lo[2 * c4 + 1]);
      }
      if (row < a.rows) {
        uint4* dst = reinterpret_cast<uint4*>(a.d0 + static_cast<size_t>(row) * a.ld_d0 + pc * 16);
        dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        if (a.d0_lo > 0) {
          uint4* dl = reinterpret_cast<uint4*>(a.d0 + static_cast<size_t>(row) * a.ld_d0 + a.d0_lo + pc * 16);
          dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
        // two 16-column pieces share a 32-bit mask word: each writes its half
        if (a.d0_bits) reinterpret_cast<unsigned short*>(a.d0_bits)[(static_cast<size_t>(pc >> 1) * a.rows + row) * 2 + (pc & 1)] = static_cast<unsigned short>(bits);
      }
    }
  }
}

// =============================================================================================
// backward
// =============================================================================================
__global__ void __launch_bounds__(HB_THREADS) head_block_bwd_kernel(const __grid_constant__ HbArgs a) {
  const HbLayout Y = hb_layout(a, true);
  float* const sm = hb_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * HB_ROWS;
  const int L = a.L, HW = a.HW, C = a.C;
  const int hwq = pad4(HW);
  pdl_wait();
  pdl_launch_dependents();
  const float beta = a.dyn ? a.dyn->beta_kl : 0.f;
  // ---- stage: heads weights transposed ([in][hwq], reduction index j contiguous), W0 transposed ([L][C]), g_d0 rows ----
  {
    int off = 0;
    for (int e = 0; e < a.n_enc; ++e) {
      const int n = a.enc[e].in_dim;
      float* dst = sm + Y.wh + off * hwq;
      // coalesced reads of W [HW][n] (k fastest), transposed stores; eight loads in flight per thread
#pragma unroll 8
      for (int i = tid; i < HW * n; i += HB_THREADS) {
        const int j = i / n, k = i - j * n;
        dst[k * hwq + j] = __ldg(a.enc[e].Wh + i);
      }
      if (hwq != HW)
        for (int i = tid; i < n * (hwq - HW); i += HB_THREADS) dst[(i / (hwq - HW)) * hwq + HW + i % (hwq - HW)] = 0.f;
      off += n;
    }
    if (a.has_dec) {
#pragma unroll 8
      for (int i = tid; i < C * L; i += HB_THREADS) {          // coalesced read of W0 [C][L]
        const int c = i / L, l = i - c * L;
        sm[Y.w0 + l * C + c] = __ldg(a.W0 + i);
      }
      // g_d0 rows: 16-byte pieces (8 bf16) -> fp32
      const int c8n = C / 8;
#pragma unroll 4
      for (int i = tid; i < HB_ROWS * c8n; i += HB_THREADS) {
        const int r = i / c8n, c8 = i - r * c8n;
        const int row = row0 + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < a.rows) v = __ldcg(reinterpret_cast<const uint4*>(a.g_d0 + static_cast<size_t>(row) * a.ld_gd0) + c8);
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x)), f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
        const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.z)), f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.w));
        float* d = sm + Y.hs + r * Y.hs_ld + 8 * c8;
        *reinterpret_cast<float4*>(d) = make_float4(f0.x, f0.y, f1.x, f1.y);
        *reinterpret_cast<float4*>(d + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
      }
    }
  }
  __syncthreads();
  // ---- dL/dz = g_d0 W0: thread = (row = lane, LPW latent columns = warp) ----
  {
    constexpr int NW = HB_THREADS / 32;
    const int LPW = (L + NW - 1) / NW;                  // <= 8 (L <= 64)
    const int l0 = warp * LPW;
    const int nq = max(0, min(LPW, L - l0));
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    if (a.has_dec) dot_rows_n(nq, sm + Y.hs + lane * Y.hs_ld, sm + Y.w0 + l0 * C, C, C, acc);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < nq) sm[Y.zs + lane * Y.zs_ld + l0 + q] = acc[q];
  }
  __syncthreads();
  // ---- d(mu | logvar) incl. beta * dKL and the autograd-supplied upstream terms, / number of encoders ----
  {
    const float inv_enc = 1.0f / a.n_enc;
    for (int i = tid; i < HB_ROWS * L; i += HB_THREADS) {
      const int r = i / L, j = i - r * L;
      const int row = row0 + r;
      float gmu = 0.f, glv = 0.f;
      if (row < a.rows) {
        const unsigned idx = static_cast<unsigned>(row) * L + j;
        const float gz = sm[Y.zs + r * Y.zs_ld + j];
        if (a.ae) {
          gmu = (gz + (a.gmu_in ? a.gmu_in[idx] : 0.f)) * inv_enc;
          a.gml[static_cast<size_t>(row) * a.ld_gml + j] = __float2bfloat16(gmu);
        } else {
          const float mu = a.mu[idx], lv = a.logvar[idx], eps = a.eps_save[idx];
          gmu = gz + beta * mu;
          glv = gz * eps * 0.5f * expf(0.5f * lv) + beta * 0.5f * (expf(lv) - 1.0f);
          if (a.gmu_in) gmu += a.gmu_in[idx];
          if (a.glv_in) glv += a.glv_in[idx];
          gmu *= inv_enc; glv *= inv_enc;
          a.gml[static_cast<size_t>(row) * a.ld_gml + j] = __float2bfloat16(gmu);
          a.gml[static_cast<size_t>(row) * a.ld_gml + L + j] = __float2bfloat16(glv);
        }
      }
      // the backward GEMMs of the weight gradients read the bf16 copy; the heads' data gradient below uses the same rounding
      sm[Y.ml + r * Y.ml_ld + j] = __bfloat162float(__float2bfloat16(gmu));
      if (!a.ae) sm[Y.ml + r * Y.ml_ld + L + j] = __bfloat162float(__float2bfloat16(glv));
    }
    for (int i = tid; i < HB_ROWS * (Y.ml_ld - HW); i += HB_THREADS) {      // zero the padding the 128-bit reads touch
      const int r = i / (Y.ml_ld - HW), j = HW + i - r * (Y.ml_ld - HW);
      sm[Y.ml + r * Y.ml_ld + j] = 0.f;
    }
  }
  __syncthreads();
  // ---- heads' data gradients, encoder by encoder: thread = (row = lane, 16-column piece(s) = warp) ----
  {
    int off = 0;
    const float* grow = sm + Y.ml + lane * Y.ml_ld;
    const int row = row0 + lane;
    const bool row_ok = row < a.rows;
    for (int e = 0; e < a.n_enc; ++e) {
      const HbEnc& E = a.enc[e];
      const int n = E.in_dim;
      const float* WT = sm + Y.wh + off * hwq;            // [n][hwq]
      if (E.kind == 0) {
        // pre-activations of the row block (for x_hat) take the place of g_d0, mean / rstd into vec
        __syncthreads();
        for (int i = tid; i < HB_ROWS * (n / 4); i += HB_THREADS) {
          const int r = i / (n / 4), c4 = i - r * (n / 4);
          float* d = sm + Y.hs + r * Y.hs_ld + 4 * c4;
          if (row0 + r < a.rows) hb_cp16(d, E.pre + static_cast<size_t>(row0 + r) * n + 4 * c4);
          else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int c = tid; c < n; c += HB_THREADS) { sm[Y.vec + c] = __ldcg(E.save_mean + c); sm[Y.vec + Y.vec_n + c] = __ldcg(E.save_rstd + c); }
        hb_cp_wait();
        __syncthreads();
      }
      for (int piece = warp; piece * 16 < n; piece += HB_THREADS / 32) {
        const int c0 = piece * 16;
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        for (int j = 0; j < hwq; j += 4) {
          const float4 g = *reinterpret_cast<const float4*>(grow + j);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c0 + c < n) {
              const float4 w = *reinterpret_cast<const float4*>(WT + (c0 + c) * hwq + j);
              acc[c] = fmaf(g.x, w.x, fmaf(g.y, w.y, fmaf(g.z, w.z, fmaf(g.w, w.w, acc[c]))));
            }
          }
        }
        if (E.kind == 0) {
          const uint32_t word = row_ok ? __ldcg(E.bits + static_cast<size_t>(piece >> 1) * a.rows + row) : 0u;
          const uint32_t bits = (piece & 1) ? (word >> 16) : (word & 0xFFFFu);
          float s2[16];
          uint32_t packed[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float v0 = ((bits >> c) & 1u) ? acc[c] * E.mask_scale : 0.f;
            float v1 = ((bits >> (c + 1)) & 1u) ? acc[c + 1] * E.mask_scale : 0.f;
            const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            packed[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            // (the statistics use the unrounded gradient, as the GEMM epilogue they replace does)
            acc[c] = v0; acc[c + 1] = v1;
            s2[c] = v0 * (sm[Y.hs + lane * Y.hs_ld + c0 + c] - sm[Y.vec + c0 + c]) * sm[Y.vec + Y.vec_n + c0 + c];
            s2[c + 1] = v1 * (sm[Y.hs + lane * Y.hs_ld + c0 + c + 1] - sm[Y.vec + c0 + c + 1]) * sm[Y.vec + Y.vec_n + c0 + c + 1];
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(E.gy + static_cast<size_t>(row) * n + c0);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
          const float t1 = col_sums16(acc, lane);
          const float t2 = col_sums16(s2, lane);
          if (lane < 16) {
            E.bstats[(static_cast<size_t>(blockIdx.x) * 2 + 0) * n + c0 + lane] = t1;
            E.bstats[(static_cast<size_t>(blockIdx.x) * 2 + 1) * n + c0 + lane] = t2;
          }
        } else if (row_ok) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c0 + c < n) E.g_x[static_cast<size_t>(row) * E.ld_gx + c0 + c] = __float2bfloat16(acc[c]);
        }
      }
      off += n;
    }
  }
}

}  // namespace

size_t hb_smem_bytes(const HbArgs& a, bool backward) { return static_cast<size_t>(hb_layout(a, backward).total) * sizeof(float); }

static cudaError_t hb_launch(void (*kernel)(const HbArgs), const HbArgs& a, bool backward, cudaStream_t s) {
  const size_t smem = hb_smem_bytes(a, backward);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static size_t set_fwd = 0, set_bwd = 0;
  size_t& have = backward ? set_bwd : set_fwd;
  if (smem > have) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    have = smem;
  }
  const int blocks = (a.rows + HB_ROWS - 1) / HB_ROWS;
  return launch_pdl(kernel, dim3(blocks), dim3(HB_THREADS), smem, s, a);
}
cudaError_t launch_head_block_fwd(const HbArgs& a, cudaStream_t s) { return hb_launch(head_block_fwd_kernel, a, false, s); }
cudaError_t launch_head_block_bwd(const HbArgs& a, cudaStream_t s) { return hb_launch(head_block_bwd_kernel, a, true, s); }

}  // namespace vla
