// Device bodies of the HBM / L2-bound element-wise and reduction steps of the VAE train step: input ingest (fp32 -> bf16
// MMA operands, embedding gather), BatchNorm apply + ReLU + dropout, the fused latent step (modality mean,
// reparameterisation, per-sample KL), the fused loss (MSE + BCE + weighted CE + KL with their gradients), BatchNorm
// backward, the latent backward and the fused multi-tensor AdamW.  Each body is one 256-thread block's worth of work; it
// is called by the stand-alone kernels (elementwise.cu) and by the chain kernel (chain_kernel.cu, MEGA = true, where
// the eight epilogue warps of the CTA form the block and synchronise on a named barrier).
// All arithmetic is fp32; reductions are deterministic (fixed-order partials, no float atomics).
#pragma once
#include "dp_frame.cuh"
#include "loss_math.cuh"
#include "tc_ptx.cuh"
#include "vla_internal.h"

#include <cfloat>

namespace vla {

constexpr int EW_THREADS = 256;
constexpr int EW_SCRATCH_BYTES = 8 * 64 * 2 * 8 + 2 * 64 * 4;   // BatchNorm: double [8][64][2] + 2 x float [64] (largest user)

// Optional progress stamps of thread 0 inside an element-wise unit of the whole-step kernel (slots 2..5 of its row).
__device__ __forceinline__ void ew_stamp(unsigned long long* row, int slot, int tid) {
  if (row != nullptr && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    row[slot] = t;
  }
}

template <bool MEGA>
__device__ __forceinline__ void ew_sync() {
  if (MEGA) named_bar_sync(2, EW_THREADS); else __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }   // [0,1)
__device__ __forceinline__ float normal_from(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777216.0f);               // (0,1]
  const float u2 = u01(b);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// hi = bf16(x, y) packed; lo = bf16 of what the rounding dropped (split-bf16 operands, DESIGN.md "Precision")
__device__ __forceinline__ uint32_t pack_bf16x2_hi_lo(float x, float y, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  lo = *reinterpret_cast<const uint32_t*>(&l);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ bf16 bf16_lo_of(float x, bf16 h) { return __float2bfloat16(x - __bfloat162float(h)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum over the 256 threads of an element-wise block; result valid in thread 0.
template <bool MEGA>
__device__ __forceinline__ float block_sum(float v, float* sh, int tid) {
  v = warp_sum(v);
  const int w = tid >> 5, l = tid & 31;
  if (l == 0) sh[w] = v;
  ew_sync<MEGA>();
  float t = 0.f;
  if (w == 0) {
    t = (l < EW_THREADS / 32) ? sh[l] : 0.f;
    t = warp_sum(t);
  }
  ew_sync<MEGA>();
  return t;
}

// ---------------------------------------------------------------------------------------------
// ingest: fp32 inputs -> bf16 padded operands; embedding gather; one-hot; step counter
// ---------------------------------------------------------------------------------------------
// Site encoder input for rows [r_begin, r_end): gathered embedding rows and the one-hot operand of the embedding
// gradient.  One element per thread and iteration (site and table loads of different rows are independent).
__device__ __forceinline__ void ingest_site(const IngestArgs& a, int r_begin, int r_end, int t, int nthreads, long long row0) {
  if (a.site == nullptr) return;
  // one warp per row, lane = column: one label load per row, no index division
  const int lane = t & 31, gwarp = t >> 5, nwarps = nthreads >> 5;
  for (int r = r_begin + gwarp; r < r_end; r += nwarps) {
    const long long s = __ldg(a.site + row0 + r);
    bf16* hrow = a.h_site + static_cast<size_t>(r) * a.ld_hsite;
    if (a.hsite_lo > 0) {                       // split layout: hi in [0, embed), lo in [hsite_lo, hsite_lo + embed), rest stays zero
      for (int c = lane; c < a.embed; c += 32) {
        const float x = __ldg(a.emb + s * a.embed + c);
        const bf16 h = __float2bfloat16(x);
        hrow[c] = h;
        hrow[a.hsite_lo + c] = bf16_lo_of(x, h);
      }
    } else {
      for (int c = lane; c < a.ld_hsite; c += 32) hrow[c] = __float2bfloat16(c < a.embed ? __ldg(a.emb + s * a.embed + c) : 0.f);
    }
    bf16* orow = a.onehot + static_cast<size_t>(r) * a.ld_onehot;
    for (int c = lane; c < a.ld_onehot; c += 32) orow[c] = __float2bfloat16(s == c ? 1.f : 0.f);
  }
}

// Rows [r_begin, r_end) are converted by `nwarps` warps of which this is number `gwarp` (one warp per row).
__device__ __forceinline__ void ingest_body(const IngestArgs& a, int r_begin, int r_end, int gwarp, int nwarps, int lane,
                                            bool first_thread) {
  if (a.bump_step && first_thread) {
    a.dyn->step += 1;
    a.dyn->dp_epoch += 1;
    a.dyn->b1pow *= static_cast<double>(a.beta1);
    a.dyn->b2pow *= static_cast<double>(a.beta2);
  }
  const long long row0 = a.n_batches > 1 ? static_cast<long long>(a.dyn->batch_index % a.n_batches) * a.rows : 0;
  for (int e = 0; e < a.n; ++e) {
    const int lo_off = a.lo_off[e];
    const float* __restrict__ src = a.src[e];
    const int w = a.width[e];
    if ((w % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0) && lo_off > 0) {
      // split layout [hi: 0 .. w | zeros | lo: lo_off .. lo_off + w | zeros], even width: one warp per row, lanes stride over
      // 8-byte pairs (rows are 8-byte aligned); 4-byte bf16x2 stores, 128 contiguous bytes per
      // warp instruction.  Only the columns that hold data are written (the padding stays zero from the allocation).
      const int pairs = w >> 1, lo_w = lo_off >> 1;
      for (int r = r_begin + gwarp; r < r_end; r += nwarps) {
        const float2* sp = reinterpret_cast<const float2*>(src + (row0 + r) * w);
        uint32_t* dp = reinterpret_cast<uint32_t*>(a.dst[e] + static_cast<size_t>(r) * a.ld_dst[e]);
        // (up to 14 pairs = 896 columns per lane in flight at once: the whole row of either modality in ONE round of loads --
        // the launch is a chain of dependent load rounds, not a bandwidth problem)
        constexpr int INFLIGHT = 14;
        for (int p0 = lane; p0 < pairs; p0 += 32 * INFLIGHT) {
          float2 v[INFLIGHT];
#pragma unroll
          for (int k = 0; k < INFLIGHT; ++k) v[k] = (p0 + 32 * k < pairs) ? __ldg(sp + p0 + 32 * k) : make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < INFLIGHT; ++k) {
            const int p = p0 + 32 * k;
            if (p < pairs) {
              uint32_t lo;
              dp[p] = pack_bf16x2_hi_lo(v[k].x, v[k].y, lo);
              dp[lo_w + p] = lo;
            }
          }
        }
      }
      continue;
    }
    // split layout [hi: 0 .. w | zeros | lo: lo_off .. lo_off + w | zeros]: only the quads that hold data are written
    const int quads = lo_off > 0 ? (a.width[e] + 3) >> 2 : a.ld_dst[e] >> 2;   // ld_dst is a multiple of 8
    const bool vec_ok = (w % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
    for (int r = r_begin + gwarp; r < r_end; r += nwarps) {    // one warp per row: no index division
      const float* sp = src + (row0 + r) * w;
      uint2* dp = reinterpret_cast<uint2*>(a.dst[e] + static_cast<size_t>(r) * a.ld_dst[e]);
      for (int base = lane; base < quads; base += 128) {       // batches of four quads: all loads before the stores
        float x[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = (base + 32 * k) * 4;
          x[k][0] = x[k][1] = x[k][2] = x[k][3] = 0.f;
          if (vec_ok) {
            if (c + 1 < w) { const float2 t = __ldg(reinterpret_cast<const float2*>(sp + c)); x[k][0] = t.x; x[k][1] = t.y; }
            if (c + 3 < w) { const float2 t = __ldg(reinterpret_cast<const float2*>(sp + c + 2)); x[k][2] = t.x; x[k][3] = t.y; }
          } else {
            if (c < w) x[k][0] = __ldg(sp + c);
            if (c + 1 < w) x[k][1] = __ldg(sp + c + 1);
            if (c + 2 < w) x[k][2] = __ldg(sp + c + 2);
            if (c + 3 < w) x[k][3] = __ldg(sp + c + 3);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int qd = base + 32 * k;
          if (qd < quads) {
            uint2 o, ol;
            o.x = pack_bf16x2_hi_lo(x[k][0], x[k][1], ol.x);
            o.y = pack_bf16x2_hi_lo(x[k][2], x[k][3], ol.y);
            dp[qd] = o;
            if (lo_off > 0) dp[(lo_off >> 2) + qd] = ol;
          }
        }
      }
    }
  }
  ingest_site(a, r_begin, r_end, gwarp * 32 + lane, nwarps * 32, row0);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm apply (+ReLU +dropout) and BatchNorm backward.  Tile: 64 columns x ROWS rows per CTA,
// 256 threads = 32 column pairs x 8 row lanes.
// ---------------------------------------------------------------------------------------------
constexpr int BN_COLS = 64;

template <bool MEGA>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ stats, int m_tiles, int n, int col,
                                                bool col_ok, double (*sh)[BN_COLS][2], double* out, int tid) {
  // out[0..3] = (sum0[col], sum0[col+1], sum1[col], sum1[col+1]); valid for threads with ty == 0
  const int lane = tid & 31, ty = tid >> 5;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  if (col_ok) {
    for (int tb = ty; tb < m_tiles; tb += 32) {       // batches of 4 tiles: 8 independent loads in flight per thread
      float2 s0[4], s1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = tb + 8 * k;
        s0[k] = s1[k] = make_float2(0.f, 0.f);
        if (t < m_tiles) {
          s0[k] = __ldcg(reinterpret_cast<const float2*>(stats + (static_cast<size_t>(t) * 2 + 0) * n + col));
          s1[k] = __ldcg(reinterpret_cast<const float2*>(stats + (static_cast<size_t>(t) * 2 + 1) * n + col));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { a0 += s0[k].x; a1 += s0[k].y; b0 += s1[k].x; b1 += s1[k].y; }
    }
  }
  sh[ty][lane * 2][0] = a0; sh[ty][lane * 2 + 1][0] = a1;
  sh[ty][lane * 2][1] = b0; sh[ty][lane * 2 + 1][1] = b1;
  ew_sync<MEGA>();
  if (ty == 0) {
    double r[4] = {0, 0, 0, 0};
    for (int t = 0; t < 8; ++t) {
      r[0] += sh[t][lane * 2][0]; r[1] += sh[t][lane * 2 + 1][0];
      r[2] += sh[t][lane * 2][1]; r[3] += sh[t][lane * 2 + 1][1];
    }
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; out[3] = r[3];
  }
}

// One block = 64 columns x rows_per_block rows.  scratch: EW_SCRATCH_BYTES of shared memory.
// MEGA (whole-step kernel): every block stores the saved mean / rstd (identical values), so that a later unit only needs
// the blocks of its own rows to have finished.
// bit i of x -> bit 2 i (x < 2^16)
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

template <bool MEGA>
__device__ __forceinline__ void bn_act_body(const BnActArgs& a, int rows_per_block, int bx, int by, int tid, void* scratch) {
  double (*sh)[BN_COLS][2] = reinterpret_cast<double (*)[BN_COLS][2]>(scratch);
  float* s_mean = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 8 * BN_COLS * 2 * sizeof(double));
  float* s_rstd = s_mean + BN_COLS;
  const int lane = tid & 31, ty = tid >> 5;
  const int col = bx * BN_COLS + lane * 2;
  const bool col_ok = col < a.n;     // n is even
  // this thread's rows are independent of the statistics: fetch them first so the two latencies overlap
  constexpr int PF = 4;
  float2 xpf[PF];
  const int row_first = by * rows_per_block + ty;
  const int row_end = min(a.rows, (by + 1) * rows_per_block);
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int row = row_first + 8 * i;
    xpf[i] = (col_ok && row < row_end) ? __ldg(reinterpret_cast<const float2*>(a.pre + static_cast<size_t>(row) * a.ld_pre + col))
                                       : make_float2(0.f, 0.f);
  }
  // ... and so are the affine parameters and the step counter: every load of the launch is issued in ONE round
  const float2 gam = col_ok ? __ldg(reinterpret_cast<const float2*>(a.gamma + col)) : make_float2(0.f, 0.f);
  const float2 bet = col_ok ? __ldg(reinterpret_cast<const float2*>(a.beta + col)) : make_float2(0.f, 0.f);
  unsigned long long offset = a.offset;
  if (a.dyn) offset += static_cast<unsigned long long>(__ldcg(&a.dyn->step)) << 20;   // bumped earlier in the same launch: bypass L1
  if (a.train) {
    double r[4];
    reduce_partials<MEGA>(a.stats, a.m_tiles, a.n, col, col_ok, sh, r, tid);
    if (ty == 0 && col_ok) {
      const int srows = a.stat_rows > 0 ? a.stat_rows : a.rows;
      for (int j = 0; j < 2; ++j) {
        const double mean = r[j] / srows;
        double var = r[2 + j] / srows - mean * mean;
        var = var < 0 ? 0 : var;
        const float rstd = rsqrtf(static_cast<float>(var) + 1e-5f);
        s_mean[lane * 2 + j] = static_cast<float>(mean);
        s_rstd[lane * 2 + j] = rstd;
        if (MEGA || by == 0) {
          a.save_mean[col + j] = static_cast<float>(mean);
          a.save_rstd[col + j] = rstd;
        }
        if (by == 0) {
          if (a.update_running) {
            const double unbiased = srows > 1 ? var * srows / (srows - 1) : var;
            a.running_mean[col + j] = 0.9f * a.running_mean[col + j] + 0.1f * static_cast<float>(mean);
            a.running_var[col + j] = 0.9f * a.running_var[col + j] + 0.1f * static_cast<float>(unbiased);
          }
        }
      }
    }
    if (a.update_running && bx == 0 && by == 0 && tid == 0 && a.num_batches_tracked)
      *a.num_batches_tracked += 1;
  } else if (ty == 0 && col_ok) {
    for (int j = 0; j < 2; ++j) {
      const float rstd = 1.0f / sqrtf(a.running_var[col + j] + 1e-5f);
      s_mean[lane * 2 + j] = a.running_mean[col + j];
      s_rstd[lane * 2 + j] = rstd;
      if (MEGA || by == 0) { a.save_mean[col + j] = a.running_mean[col + j]; a.save_rstd[col + j] = rstd; }
    }
  }
  ew_sync<MEGA>();
  if (!col_ok) return;
  const float m0 = s_mean[lane * 2], m1 = s_mean[lane * 2 + 1];
  const float r0 = s_rstd[lane * 2] * gam.x, r1 = s_rstd[lane * 2 + 1] * gam.y;
  const float b0 = bet.x, b1 = bet.y;
  const bool drop = a.train && a.p_drop > 0.f;
  const float keep_scale = drop ? 1.0f / (1.0f - a.p_drop) : 1.0f;
  auto do_row = [&](int row, float2 x) {
    float y0 = fmaxf((x.x - m0) * r0 + b0, 0.f);
    float y1 = fmaxf((x.y - m1) * r1 + b1, 0.f);
    if (drop) {
      bool k0, k1;
      if (a.keep_mask) {
        const uchar2 k = *reinterpret_cast<const uchar2*>(a.keep_mask + static_cast<size_t>(row) * a.n + col);
        k0 = k.x != 0; k1 = k.y != 0;
      } else {
        const unsigned long long idx = (static_cast<unsigned long long>(row) * a.n + col) >> 1;
        const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                                   static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                        make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32)));
        k0 = u01(rnd.x) >= a.p_drop; k1 = u01(rnd.y) >= a.p_drop;
      }
      y0 = k0 ? y0 * keep_scale : 0.f;
      y1 = k1 ? y1 * keep_scale : 0.f;
    }
    uint32_t lo;
    const uint32_t hi = pack_bf16x2_hi_lo(y0, y1, lo);
    uint32_t* dst = reinterpret_cast<uint32_t*>(a.out + static_cast<size_t>(row) * a.ld_out + col);
    *dst = hi;
    if (a.out_lo > 0) dst[a.out_lo >> 1] = lo;
  };
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int row = row_first + 8 * i;
    if (row < row_end) do_row(row, xpf[i]);
  }
  for (int rb = row_first + 8 * PF; rb < row_end; rb += 8 * PF) {      // further batches of PF rows: loads before the stores
    float2 xb[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int row = rb + 8 * i;
      xb[i] = row < row_end ? __ldg(reinterpret_cast<const float2*>(a.pre + static_cast<size_t>(row) * a.ld_pre + col)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int row = rb + 8 * i;
      if (row < row_end) do_row(row, xb[i]);
    }
  }
}

template <bool MEGA>
__device__ __forceinline__ void bn_bwd_body(const BnBwdArgs& a, int rows_per_block, int bx, int by, int tid, void* scratch) {
  double (*sh)[BN_COLS][2] = reinterpret_cast<double (*)[BN_COLS][2]>(scratch);
  float* s_s1 = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 8 * BN_COLS * 2 * sizeof(double));
  float* s_s2 = s_s1 + BN_COLS;
  const int lane = tid & 31, ty = tid >> 5;
  const int col = bx * BN_COLS + lane * 2;
  const bool col_ok = col < a.n;
  // everything that does not depend on the reduced statistics is loaded first (one round of loads for the launch): the saved
  // mean / rstd, gamma, and this thread's first batch of rows
  const float2 mean2 = col_ok ? __ldcg(reinterpret_cast<const float2*>(a.mean + col)) : make_float2(0.f, 0.f);
  const float2 rstd2 = col_ok ? __ldcg(reinterpret_cast<const float2*>(a.rstd + col)) : make_float2(0.f, 0.f);
  const float2 gam2 = col_ok ? __ldg(reinterpret_cast<const float2*>(a.gamma + col)) : make_float2(0.f, 0.f);
  const int row_end = min(a.rows, (by + 1) * rows_per_block);
  const int rb_first = by * rows_per_block + ty;
  float2 gy0[4], x0[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = rb_first + 8 * k;
    gy0[k] = x0[k] = make_float2(0.f, 0.f);
    if (col_ok && row < row_end) {
      gy0[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a.gy + static_cast<size_t>(row) * a.ld_gy + col));
      x0[k] = __ldg(reinterpret_cast<const float2*>(a.pre + static_cast<size_t>(row) * a.ld_pre + col));
    }
  }
  {
    double r[4];
    reduce_partials<MEGA>(a.stats, a.m_tiles, a.n, col, col_ok, sh, r, tid);
    if (ty == 0 && col_ok) {
      for (int j = 0; j < 2; ++j) {
        s_s1[lane * 2 + j] = static_cast<float>(r[j]);
        s_s2[lane * 2 + j] = static_cast<float>(r[2 + j]);
        if (by == 0) {
          const double gs = a.param_grad_scale > 0.f ? a.param_grad_scale : 1.0;
          a.dbeta[col + j] = static_cast<float>(r[j] * gs);
          a.dgamma[col + j] = static_cast<float>(r[2 + j] * gs);
        }
      }
    }
  }
  ew_sync<MEGA>();
  if (!col_ok) return;
  const float inv_n = 1.0f / (a.stat_rows > 0 ? a.stat_rows : a.rows);
  const float m0 = mean2.x, m1 = mean2.y;
  const float rs0 = rstd2.x, rs1 = rstd2.y;
  const float g0 = gam2.x * rs0, g1 = gam2.y * rs1;
  const float c10 = a.train ? s_s1[lane * 2] * inv_n : 0.f, c11 = a.train ? s_s1[lane * 2 + 1] * inv_n : 0.f;
  const float c20 = a.train ? s_s2[lane * 2] * inv_n : 0.f, c21 = a.train ? s_s2[lane * 2 + 1] * inv_n : 0.f;
  for (int rb = rb_first; rb < row_end; rb += 32) {   // batches of four rows: all loads before the stores
    float2 gy[4], x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = rb + 8 * k;
      gy[k] = x[k] = make_float2(0.f, 0.f);
      if (rb == rb_first) { gy[k] = gy0[k]; x[k] = x0[k]; }       // (fetched before the reduction)
      else if (row < row_end) {
        gy[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a.gy + static_cast<size_t>(row) * a.ld_gy + col));
        x[k] = __ldg(reinterpret_cast<const float2*>(a.pre + static_cast<size_t>(row) * a.ld_pre + col));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = rb + 8 * k;
      if (row < row_end) {
        const float xh0 = (x[k].x - m0) * rs0, xh1 = (x[k].y - m1) * rs1;
        const float o0 = g0 * (gy[k].x - c10 - xh0 * c20);
        const float o1 = g1 * (gy[k].y - c11 - xh1 * c21);
        *reinterpret_cast<__nv_bfloat162*>(a.gpre + static_cast<size_t>(row) * a.ld_gpre + col) = __floats2bfloat162_rn(o0, o1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Latent: mean over present encoders' (mu | logvar) heads, z = mu + eps * exp(logvar / 2), KL partials
// (vae.py:11-15, 64-73; losses.py:42)
// ---------------------------------------------------------------------------------------------
// Element (r, j) of the latent step; returns its KL summand 1 + logvar - mu^2 - exp(logvar) (0 for the autoencoders).
__device__ __forceinline__ float latent_fwd_elem(const LatentFwdArgs& a, int r, int j, unsigned long long offset) {
  const unsigned idx = static_cast<unsigned>(r) * a.L + j;             // rows * L < 2^31 (checked by the launcher)
  float mu = 0.f, lv = 0.f;
  for (int e = 0; e < a.n_enc; ++e) {
    const float* p = a.ml[e] + static_cast<size_t>(r) * a.ld_ml[e];
    mu += p[j];
    if (!a.ae) lv += p[a.L + j];
  }
  if (a.n_enc > 1) { mu /= a.n_enc; lv /= a.n_enc; }
  if (a.ae) {
    // directional autoencoders (directional_ae.py:46-59): the mean of the encoder outputs IS the decoder input
    a.mu[idx] = mu;
    a.logvar[idx] = 0.f;
    a.eps_save[idx] = 0.f;
    const bf16 h = __float2bfloat16(mu);
    a.z[static_cast<size_t>(r) * a.ld_z + j] = h;
    if (a.z_lo > 0) a.z[static_cast<size_t>(r) * a.ld_z + a.z_lo + j] = bf16_lo_of(mu, h);
    return 0.f;
  }
  float eps;
  if (a.eps_in) {
    eps = a.eps_in[idx];
  } else {
    const uint4 rnd = philox4x32_10(make_uint4(idx, 0u, static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32) ^ 0x5EEDu),
                                    make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32)));
    eps = normal_from(rnd.x, rnd.y);
  }
  const float sd = expf(0.5f * lv);
  const float z = mu + eps * sd;
  a.mu[idx] = mu;
  a.logvar[idx] = lv;
  a.eps_save[idx] = eps;
  const bf16 h = __float2bfloat16(z);
  a.z[static_cast<size_t>(r) * a.ld_z + j] = h;
  if (a.z_lo > 0) a.z[static_cast<size_t>(r) * a.ld_z + a.z_lo + j] = bf16_lo_of(z, h);
  return 1.0f + lv - mu * mu - expf(lv);
}
__device__ __forceinline__ unsigned long long latent_offset(const LatentFwdArgs& a) {
  unsigned long long offset = a.offset;
  if (a.dyn && !a.eps_in) offset += static_cast<unsigned long long>(__ldcg(&a.dyn->step)) << 20;   // bumped by an earlier launch of the step
  return offset;
}

// Block b covers elements [256 b, 256 b + 256) of the [rows, L] latent.
template <bool MEGA>
__device__ __forceinline__ void latent_fwd_body(const LatentFwdArgs& a, int b, int tid, void* scratch) {
  float* sh = reinterpret_cast<float*>(scratch);
  const unsigned idx = static_cast<unsigned>(b) * EW_THREADS + tid;
  const unsigned total = static_cast<unsigned>(a.rows) * a.L;
  float kl = 0.f;
  if (idx < total) {
    const int r = static_cast<int>(idx / static_cast<unsigned>(a.L));
    kl = latent_fwd_elem(a, r, static_cast<int>(idx - static_cast<unsigned>(r) * a.L), latent_offset(a));
  }
  const float t = block_sum<MEGA>(kl, sh, tid);
  if (tid == 0) a.kl_partials[b] = -0.5f * t;
}

// Chain kernel: rows [r0, r1) by the 256 element-wise threads of one CTA; one KL partial (index `part`).
template <bool MEGA>
__device__ __forceinline__ void latent_fwd_rows(const LatentFwdArgs& a, int r0, int r1, int part, int tid, void* scratch) {
  float* sh = reinterpret_cast<float*>(scratch);
  const unsigned long long offset = latent_offset(a);
  float kl = 0.f;
  const int n = (r1 - r0) * a.L;
  for (int i = tid; i < n; i += EW_THREADS) {
    const int rr = i / a.L;
    kl += latent_fwd_elem(a, r0 + rr, i - rr * a.L, offset);
  }
  const float t = block_sum<MEGA>(kl, sh, tid);
  if (tid == 0) a.kl_partials[part] = -0.5f * t;
}

__device__ __forceinline__ void latent_bwd_elem(const LatentBwdArgs& a, int r, int j, float beta) {
  const unsigned idx = static_cast<unsigned>(r) * a.L + j;
  const float gz = a.gz ? a.gz[static_cast<size_t>(r) * a.ld_gz + j] : 0.f;
  if (a.ae) {
    float g = gz + (a.gmu_in ? a.gmu_in[idx] : 0.f);
    if (a.n_modalities > 1) g /= a.n_modalities;
    a.gml[static_cast<size_t>(r) * a.ld_gml + j] = __float2bfloat16(g);
    return;
  }
  const float mu = a.mu[idx], lv = a.logvar[idx], eps = a.eps[idx];
  float gmu = gz + beta * mu;
  float glv = gz * eps * 0.5f * expf(0.5f * lv) + beta * 0.5f * (expf(lv) - 1.0f);
  if (a.gmu_in) gmu += a.gmu_in[idx];
  if (a.glv_in) glv += a.glv_in[idx];
  if (a.n_modalities > 1) { gmu /= a.n_modalities; glv /= a.n_modalities; }
  a.gml[static_cast<size_t>(r) * a.ld_gml + j] = __float2bfloat16(gmu);
  a.gml[static_cast<size_t>(r) * a.ld_gml + a.L + j] = __float2bfloat16(glv);
}

__device__ __forceinline__ void latent_bwd_body(const LatentBwdArgs& a, int b, int tid) {
  const unsigned idx = static_cast<unsigned>(b) * EW_THREADS + tid;
  const unsigned total = static_cast<unsigned>(a.rows) * a.L;
  if (idx >= total) return;
  const int r = static_cast<int>(idx / static_cast<unsigned>(a.L));
  latent_bwd_elem(a, r, static_cast<int>(idx - static_cast<unsigned>(r) * a.L), a.dyn ? a.dyn->beta_kl : a.beta);
}

__device__ __forceinline__ void latent_bwd_rows(const LatentBwdArgs& a, int r0, int r1, int tid) {
  const float beta = a.dyn ? a.dyn->beta_kl : a.beta;
  const int n = (r1 - r0) * a.L;
  for (int i = tid; i < n; i += EW_THREADS) {
    const int rr = i / a.L;
    latent_bwd_elem(a, r0 + rr, i - rr * a.L, beta);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused loss: MSE-sum + BCE-sum (ATen clamps) + weighted CE-sum + KL-sum, values and gradients
// (losses.py:27-46; directional_losses.py:23-30, 48-55).  Block roles by block index range.
// ---------------------------------------------------------------------------------------------
constexpr int LOSS_THREADS = 256;
constexpr int LOSS_WARPS = LOSS_THREADS / 32;
constexpr int LOSS_PER_THREAD = 16;
constexpr int LOSS_PER_BLOCK = LOSS_THREADS * LOSS_PER_THREAD;

__host__ __device__ inline int ceil_div_ll(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// Block roles by block index range: [MSE rows | BCE rows | CE rows | KL elements].  MSE / BCE: one warp per sample row,
// lanes stride over 8-byte pairs (both 782 and 572 are even), so no per-element index arithmetic is needed for the
// padded bf16 gradient rows.
struct LossGrid { int nb_a, nb_b, nb_c, nb_k; };
__host__ __device__ inline LossGrid loss_grid(const LossArgs& a) {
  LossGrid g;
  g.nb_a = a.recon_a ? ceil_div_ll(a.rows, LOSS_WARPS) : 0;
  g.nb_b = a.recon_b ? ceil_div_ll(a.rows, LOSS_WARPS) : 0;
  g.nb_c = a.logits ? ceil_div_ll(a.rows, LOSS_THREADS) : 0;
  g.nb_k = (a.mu && !a.kl_partials) ? ceil_div_ll(static_cast<long long>(a.rows) * a.L, LOSS_PER_BLOCK) : 0;
  return g;
}

template <bool BCE>
__device__ __forceinline__ float loss_row(const float* __restrict__ rp, const float* __restrict__ tp, int w, float gs,
                                          float* __restrict__ gf, bf16* __restrict__ gb, int lane) {
  float acc = 0.f;
  const bool vec = (w % 2 == 0) && (((reinterpret_cast<uintptr_t>(rp) | reinterpret_cast<uintptr_t>(tp)) & 7) == 0) &&
                   (gf == nullptr || (reinterpret_cast<uintptr_t>(gf) & 7) == 0);
  if (vec) {
    const int n2 = w >> 1;
    const float2* r2 = reinterpret_cast<const float2*>(rp);
    const float2* t2 = reinterpret_cast<const float2*>(tp);
    // batches of 4 independent load pairs per lane before any arithmetic (memory-level parallelism)
    for (int base = lane; base < n2; base += 128) {
      float2 y[4], t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = base + 32 * k;
        if (i < n2) { y[k] = r2[i]; t[k] = __ldg(t2 + i); }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = base + 32 * k;
        if (i < n2) {
          float g0, g1, l0, l1;
          acc += loss_elem<BCE>(y[k].x, t[k].x, gs, g0, l0);
          acc += loss_elem<BCE>(y[k].y, t[k].y, gs, g1, l1);
          if (gf) reinterpret_cast<float2*>(gf)[i] = make_float2(g0, g1);
          if (gb) reinterpret_cast<__nv_bfloat162*>(gb)[i] = __floats2bfloat162_rn(l0, l1);
        }
      }
    }
  } else {
    for (int i = lane; i < w; i += 32) {
      float g0, l0;
      acc += loss_elem<BCE>(rp[i], __ldg(tp + i), gs, g0, l0);
      if (gf) gf[i] = g0;
      if (gb) gb[i] = __float2bfloat16(l0);
    }
  }
  return acc;
}

// Block `block` of `grid` (roles by block range, see loss_grid).  The block that finishes last reduces every partial.
template <bool MEGA>
__device__ __forceinline__ void loss_body(const LossArgs& a, int block, int grid, int tid, void* scratch) {
  float* sh = reinterpret_cast<float*>(scratch);
  int* s_last = reinterpret_cast<int*>(sh + 32);
  double* dsh = reinterpret_cast<double*>(sh + 64);
  const LossGrid G = loss_grid(a);
  const float beta = a.dyn ? a.dyn->beta_kl : a.beta;
  const float gamma = a.dyn ? a.dyn->gamma : a.gamma;
  const float gs = a.grad_scale;
  const long long row0 = a.n_batches > 1 ? static_cast<long long>(a.dyn->batch_index % a.n_batches) * a.rows : 0;
  const int warp = tid >> 5, lane = tid & 31;
  int b = block;
  float acc = 0.f;
  if (b < G.nb_a) {
    const int r = b * LOSS_WARPS + warp;
    if (r < a.rows)
      acc = loss_row<false>(a.recon_a + static_cast<size_t>(r) * a.width_a, a.a + (row0 + r) * a.width_a, a.width_a, gs,
                            a.ga_f32 ? a.ga_f32 + static_cast<size_t>(r) * a.width_a : nullptr,
                            a.ga_bf16 ? a.ga_bf16 + static_cast<size_t>(r) * a.ld_ga : nullptr, lane);
  } else if ((b -= G.nb_a) < G.nb_b) {
    const int r = b * LOSS_WARPS + warp;
    if (r < a.rows)
      acc = loss_row<true>(a.recon_b + static_cast<size_t>(r) * a.width_b, a.b + (row0 + r) * a.width_b, a.width_b, gs,
                           a.gb_f32 ? a.gb_f32 + static_cast<size_t>(r) * a.width_b : nullptr,
                           a.gb_bf16 ? a.gb_bf16 + static_cast<size_t>(r) * a.ld_gb : nullptr, lane);
  } else if ((b -= G.nb_b) < G.nb_c) {
    // ---- weighted cross-entropy, one thread per sample ----
    const int r = b * LOSS_THREADS + tid;
    if (r < a.rows) {
      const float* x = a.logits + static_cast<size_t>(r) * a.n_sites;
      const int t = static_cast<int>(a.site[row0 + r]);
      float mx = -FLT_MAX;
      for (int j = 0; j < a.n_sites; ++j) mx = fmaxf(mx, x[j]);
      float se = 0.f;
      for (int j = 0; j < a.n_sites; ++j) se += expf(x[j] - mx);
      const float lse = logf(se) + mx;
      const float w = a.class_w ? a.class_w[t] : 1.0f;
      acc = -w * (x[t] - lse);
      if (a.gc_f32 || a.gc_bf16) {
        const float sc = w * gamma * gs;
        for (int j = 0; j < a.n_sites; ++j) {
          const float g = (expf(x[j] - lse) - (j == t ? 1.0f : 0.0f)) * sc;
          if (a.gc_f32) a.gc_f32[static_cast<size_t>(r) * a.n_sites + j] = g;
          if (a.gc_bf16) a.gc_bf16[static_cast<size_t>(r) * a.ld_gc + j] = __float2bfloat16(g);
        }
      }
    }
  } else if ((b -= G.nb_c) < G.nb_k) {
    // ---- KL directly from mu / logvar (functional loss API) ----
    const long long total = static_cast<long long>(a.rows) * a.L;
    const long long base = static_cast<long long>(b) * LOSS_PER_BLOCK;
#pragma unroll 4
    for (int i = 0; i < LOSS_PER_THREAD; ++i) {
      const long long idx = base + static_cast<long long>(i) * LOSS_THREADS + tid;
      if (idx < total) {
        const float mu = a.mu[idx], lv = a.logvar[idx];
        const float e = expf(lv);
        acc += -0.5f * (1.0f + lv - mu * mu - e);
        if (a.gmu_f32) a.gmu_f32[idx] = beta * mu * gs;
        if (a.glv_f32) a.glv_f32[idx] = beta * 0.5f * (e - 1.0f) * gs;
      }
    }
  }
  const float t = block_sum<MEGA>(acc, sh, tid);
  if (tid == 0) {
    a.partials[block] = t;
    __threadfence();
    const unsigned int ticket = atomicAdd(a.counter, 1u);
    *s_last = (ticket == static_cast<unsigned int>(grid) - 1u) ? 1 : 0;
  }
  ew_sync<MEGA>();
  const bool last = *s_last != 0;
  ew_sync<MEGA>();          // s_last may be rewritten by the next block this CTA processes
  if (!last) return;
  __threadfence();
  // ---- last block: fixed-order reduction of every role's partials ----
  double sums[4] = {0, 0, 0, 0};   // mse, bce, ce, kl
  const int starts[5] = {0, G.nb_a, G.nb_a + G.nb_b, G.nb_a + G.nb_b + G.nb_c, G.nb_a + G.nb_b + G.nb_c + G.nb_k};
  for (int role = 0; role < 4; ++role) {
    double s = 0;
    for (int i = starts[role] + tid; i < starts[role + 1]; i += LOSS_THREADS) s += __ldcg(a.partials + i);
    if (role == 3 && a.kl_partials)
      for (int i = tid; i < a.n_kl_partials; i += LOSS_THREADS) s += __ldcg(a.kl_partials + i);
    dsh[tid] = s;
    ew_sync<MEGA>();
    for (int o = LOSS_THREADS / 2; o > 0; o >>= 1) {
      if (tid < o) dsh[tid] += dsh[tid + o];
      ew_sync<MEGA>();
    }
    sums[role] = dsh[0];
    ew_sync<MEGA>();
  }
  if (tid == 0) {
    const double recon = sums[0] + sums[1];
    a.out[0] = static_cast<float>(recon + static_cast<double>(gamma) * sums[2] + static_cast<double>(beta) * sums[3]);
    a.out[1] = static_cast<float>(recon);
    a.out[2] = static_cast<float>(sums[2]);
    a.out[3] = static_cast<float>(sums[3]);
    *a.counter = 0;   // re-arm for the next launch (graph replays)
    if (a.dyn_bump) a.dyn_bump->batch_index += 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused multi-tensor AdamW over the flat arena + refresh of the bf16 MMA shadows (both orientations)
// (torch.optim.AdamW semantics; call sites train_rna2dna.py:94-96, 185-189)
// ---------------------------------------------------------------------------------------------
// g_given: the gradient of this thread's four elements when the caller already holds it (dp_adamw_kernel: shard owner).
__device__ __forceinline__ void adamw_body(const AdamArgs& a, int chunk, int tid, const float4* g_given = nullptr,
                                           long long off_hint = -1, long long arena_elems = 0) {
  // speculative loads at the hinted arena offset (AdamHints): in flight together with the chunk-table entry
  const long long gi_spec = off_hint >= 0 ? off_hint + 4 * tid : -1;
  const bool spec = gi_spec >= 0 && gi_spec + 4 <= arena_elems && a.update && a.gframed == nullptr && g_given == nullptr;
  float4 sp_p = make_float4(0.f, 0.f, 0.f, 0.f), sp_g = sp_p, sp_m = sp_p, sp_v = sp_p;
  if (spec) {
    sp_p = *reinterpret_cast<const float4*>(a.p + gi_spec);
    sp_g = *reinterpret_cast<const float4*>(a.g + gi_spec);
    sp_m = *reinterpret_cast<const float4*>(a.m + gi_spec);
    sp_v = *reinterpret_cast<const float4*>(a.v + gi_spec);
  }
  const AdamChunk ch = a.chunks[chunk];
  const int e = 4 * tid;                               // element index inside the chunk
  if (e >= ch.n) return;
  float lr = a.lr, wd = a.weight_decay, bc1 = a.bc1, inv_bc2s = a.inv_bc2_sqrt;
  if (a.dyn) {   // independent of the chunk-table read above: the loads overlap
    lr = a.dyn->lr; wd = a.dyn->weight_decay;
    bc1 = static_cast<float>(1.0 - __ldcg(&a.dyn->b1pow));   // advanced earlier in the same launch: bypass L1
    inv_bc2s = static_cast<float>(1.0 / sqrt(1.0 - __ldcg(&a.dyn->b2pow)));
  }
  const float step_size = lr / bc1;
  const float decay = 1.0f - lr * wd;
  const float b1 = a.beta1, b2 = a.beta2, eps = a.eps;
  const long long gi = ch.offset + e;                  // multiple of 4: 16-byte aligned in every arena
  const int nv = min(4, ch.n - e);
  float p[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f}, m[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
  const bool framed = a.update && a.gframed != nullptr && g_given == nullptr;
  uint4 f0 = make_uint4(0u, 0u, 0u, 0u), f1 = f0;
  unsigned int epoch = 0u;
  if (framed) {               // first attempt at the two framed words of this float4, in flight with the loads below
    epoch = static_cast<unsigned int>(__ldcg(&a.dyn->dp_epoch));
    f0 = ld_framed(a.gframed + (gi >> 1));
    f1 = ld_framed(a.gframed + (gi >> 1) + 1);
  }
  if (nv == 4 && spec && gi == gi_spec) {
    *reinterpret_cast<float4*>(p) = sp_p; *reinterpret_cast<float4*>(g) = sp_g;
    *reinterpret_cast<float4*>(m) = sp_m; *reinterpret_cast<float4*>(v) = sp_v;
  } else if (nv == 4) {
    *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(a.p + gi);
    if (a.update) {
      if (!framed && !g_given) *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(a.g + gi);
      *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(a.m + gi);
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(a.v + gi);
    }
  } else {
    for (int k = 0; k < nv; ++k) {
      p[k] = a.p[gi + k];
      if (a.update) { if (!framed && !g_given) g[k] = a.g[gi + k]; m[k] = a.m[gi + k]; v[k] = a.v[gi + k]; }
    }
  }
  if (g_given) { g[0] = g_given->x; g[1] = g_given->y; g[2] = g_given->z; g[3] = g_given->w; }
  if (framed) {               // (the arena is padded to multiples of 4: both words exist for every chunk tail)
    const float2 lo = finish_framed(a.gframed + (gi >> 1), f0, epoch);
    const float2 hi = finish_framed(a.gframed + (gi >> 1) + 1, f1, epoch);
    g[0] = lo.x; g[1] = lo.y; g[2] = hi.x; g[3] = hi.y;
  }
  if (a.update) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      p[k] *= decay;
      m[k] = b1 * m[k] + (1.0f - b1) * g[k];
      v[k] = b2 * v[k] + (1.0f - b2) * g[k] * g[k];
      p[k] -= step_size * __fdividef(m[k], sqrtf(v[k]) * inv_bc2s + eps);
    }
    if (nv == 4) {
      *reinterpret_cast<float4*>(a.p + gi) = *reinterpret_cast<const float4*>(p);
      *reinterpret_cast<float4*>(a.m + gi) = *reinterpret_cast<const float4*>(m);
      *reinterpret_cast<float4*>(a.v + gi) = *reinterpret_cast<const float4*>(v);
      if (a.zero_grad) *reinterpret_cast<float4*>(a.gclear + gi) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int k = 0; k < nv; ++k) {
        a.p[gi + k] = p[k]; a.m[gi + k] = m[k]; a.v[gi + k] = v[k];
        if (a.zero_grad) a.gclear[gi + k] = 0.f;
      }
    }
  }
  if (ch.shadow_off >= 0) {
    // bf16 copy [rows, ld_shadow] used as the tensor-core operand (K-major for forward, MN-major for data gradients)
    const unsigned idx = static_cast<unsigned>(ch.first + e);                                // one tensor < 2^31 elements
    int r = static_cast<int>(idx / static_cast<unsigned>(ch.cols));
    int c = static_cast<int>(idx - static_cast<unsigned>(r) * ch.cols);
    if (nv == 4 && !(ch.cols & 1) && !(ch.shadow_off & 1) && !(ch.sh_lo & 1)) {
      // even row width (idx is a multiple of 4, so c is even and a pair never straddles a row): two packed 4-byte stores per
      // copy instead of four 2-byte ones (the row pitch and the lo offset are multiples of 8)
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        uint32_t lo;
        const uint32_t hi = pack_bf16x2_hi_lo(p[k], p[k + 1], lo);
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.shadow + ch.shadow_off + static_cast<long long>(r) * ch.ld_shadow + c);
        *dst = hi;
        if (ch.sh_lo > 0) dst[ch.sh_lo >> 1] = lo;
        c += 2;
        if (c >= ch.cols) { c = 0; ++r; }
      }
    } else {
      for (int k = 0; k < nv; ++k) {
        const bf16 h = __float2bfloat16(p[k]);
        bf16* dst = a.shadow + ch.shadow_off + static_cast<long long>(r) * ch.ld_shadow + c;
        *dst = h;
        if (ch.sh_lo > 0) dst[ch.sh_lo] = bf16_lo_of(p[k], h);
        if (++c == ch.cols) { c = 0; ++r; }
      }
    }
  }
}

// Final reduction of the fused loss by block 0 of the step's AdamW launch (LossTail with counter == nullptr): the same fixed
// order as the in-tile reduction of gemm_tile.cuh (every thread a strided share of each partial array, one shuffle butterfly
// per term, thread 0 adds the eight warp results in order).  256 threads.
__device__ __forceinline__ void loss_tail_reduce(const LossTail& T, int tid, double* dsh /* [8 * 4] shared */) {
  const int starts[4] = {0, T.n_mse, T.n_mse + T.n_bce, T.n_mse + T.n_bce + T.n_ce};
  double acc4[4] = {0, 0, 0, 0};                             // mse, bce, ce, kl
  for (int role = 0; role < 3; ++role)
    for (int i = starts[role] + tid; i < starts[role + 1]; i += 256) acc4[role] += __ldcg(T.partials + i);
  for (int i = tid; i < T.n_kl; i += 256) acc4[3] += __ldcg(T.kl_partials + i);
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int role = 0; role < 4; ++role) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc4[role] += __shfl_xor_sync(0xffffffffu, acc4[role], o);
    if (lane == 0) dsh[warp * 4 + role] = acc4[role];
  }
  __syncthreads();
  if (tid == 0) {
    double sums[4] = {0, 0, 0, 0};
    for (int w = 0; w < 8; ++w)
      for (int role = 0; role < 4; ++role) sums[role] += dsh[w * 4 + role];
    const double beta = T.dyn->beta_kl, gamma = T.dyn->gamma;
    const double recon = sums[0] + sums[1];
    T.out[0] = static_cast<float>(recon + gamma * sums[2] + beta * sums[3]);
    T.out[1] = static_cast<float>(recon);
    T.out[2] = static_cast<float>(sums[2]);
    T.out[3] = static_cast<float>(sums[3]);
    if (T.dyn_bump) T.dyn_bump->batch_index += 1;
  }
}

}  // namespace vla
