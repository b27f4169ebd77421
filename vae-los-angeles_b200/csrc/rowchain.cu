// Row-chain kernel: the row-local middle of a directional model's train step, every activation on chip
// (vla_internal.h "Row-chain kernel" has the design; the plan is built by vla_api.cu build_rowchain_plan).
//
// Replaces, for RNA2DNAVAE at batch 4096, eleven launches -- BN apply, heads, latent, three decoder layers + loss, three
// decoder data gradients, latent backward, encoder data gradient: the loop body of the reference at train_rna2dna.py:86-95
// between the two BatchNorm reductions (encoders.py:12-19, 54-57; directional_vae.py:36-47; decoders.py:26-33;
// directional_losses.py:23-30) -- by ONE launch whose only global traffic is what the step has to read (BatchNorm
// pre-activations, loss targets, weights) and what the weight-gradient GEMMs have to find in memory afterwards.
//
// Warp roles: warps 0..7 = epilogue / element-wise, warp 8 = TMA producer (weight tiles, side-input tiles), warp 9 = MMA
// issuer + TMEM owner + TMA stores -- the two single-thread roles sit on the HIGHEST warp ids (the scheduler prefers them)
// (epilogue: (thread = row of the 128-row block, the two warps of a TMEM lane quarter take
// alternate 16-column pieces).
#include "elementwise_dev.cuh"
#include "gemm_tile.cuh"

namespace vla {

namespace {

constexpr int RC_THREADS = 320;
constexpr int RC_SLOT = 16384;                        // one [128 x 64] bf16 operand block (128-byte swizzle)
constexpr int RC_NSLOT = 9;                           // activation slots: the widest resident operand is 572 -> 9 blocks
constexpr int RC_RING = 5;                            // streaming slots (weight tiles, fp32 side tiles)
constexpr int RC_RING_OFF = RC_NSLOT * RC_SLOT;
constexpr int RC_MISC_OFF = RC_RING_OFF + RC_RING * RC_SLOT;
constexpr int RC_BAR_OFF = RC_MISC_OFF;               // full[5] empty[5] acc act lda | tmem slot (+120) | flag (+124)
constexpr int RC_VEC_OFF = RC_MISC_OFF + 128;         // BatchNorm mean[128] rstd[128] rstd * gamma [128] beta[128]
constexpr int RC_DSH_OFF = RC_VEC_OFF + 2048;         // final loss reduction: 8 warps x 4 doubles
constexpr int RC_SMEM_USED = RC_DSH_OFF + 256;
constexpr int RC_SMEM_BYTES = RC_SMEM_USED;           // no alignment slack: the dynamic shared memory is declared 1 KiB-aligned
constexpr int RC_BN_MAX = 128;
extern __shared__ __align__(1024) uint8_t rc_dyn_smem[];
static_assert(RC_SMEM_BYTES <= 227 * 1024, "row-chain shared memory budget");
static_assert((2 * RC_RING + 3) * 8 <= 120, "barrier block");

struct Ring { int stage; uint32_t phase; };
__device__ __forceinline__ void ring_adv(Ring& r, int n = 1) {
  const int s = r.stage + n;
  r.phase ^= static_cast<uint32_t>(s / RC_RING) & 1u;
  r.stage = s % RC_RING;
}
__device__ __forceinline__ int pad16(int x) { return (x + 15) & ~15; }

// TMA store of one [128 x 64] block (shared -> global, clipped at the tensor's extents)
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// Everything an epilogue function needs to know about its thread.
struct Epi {
  uint8_t* smem;
  uint64_t* full; uint64_t* empty; uint64_t* acc_bar; uint64_t* act_bar;
  uint32_t tmem;          // TMEM base + this warp's lane quarter
  int q, half, lane, et, warp8;
  int row;                // row inside the 128-row block
  int grow;               // row inside the batch
  bool row_ok;
  int mb, rows;
  long long tgt_row0;     // first dataset row of the batch in flight (resident dataset)
  Ring ring;
  uint32_t acc_par;
  unsigned long long* dbg_row;   // optional stamps of the op in flight
};

// byte offset of 16-byte chunk `ch` (0..7) of row `row` inside a 128-byte-swizzled [128 x 128 B] block
__device__ __forceinline__ int sw_off(int row, int ch) { return (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4); }

// 16 consecutive fp32 of this thread's row in a side tile [128 x 32] fp32: columns half * 16 ..
__device__ __forceinline__ void tile_row16(const uint8_t* tile, int row, int half, float* x) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(tile + sw_off(row, half * 4 + i));
    x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
  }
}
// 16 consecutive floats of a per-column vector (bias, gamma, ...) starting at col; zero beyond n
__device__ __forceinline__ void vec16(const float* __restrict__ p, int col, int n, float* x) {
  if (p == nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = 0.f;
  } else if (col + 16 <= n && ((reinterpret_cast<uintptr_t>(p + col) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + col) + i);
      x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = col + j < n ? __ldg(p + col + j) : 0.f;
  }
}
// 16 values of this thread's row -> bf16 (hi) [and what the rounding dropped (lo)] into operand slots: global column gcol
// (multiple of 16) of a tensor whose column 0 lives in slot `slot0`
__device__ __forceinline__ void put16(uint8_t* smem, int slot0, int lo_slot0, int row, int gcol, const float* v) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hi[i] = pack_bf16x2_hi_lo(v[2 * i], v[2 * i + 1], lo[i]);
  const int kb = gcol >> 6, ch = (gcol & 63) >> 3;
  uint8_t* dst = smem + (slot0 + kb) * RC_SLOT;
  *reinterpret_cast<uint4*>(dst + sw_off(row, ch)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + sw_off(row, ch + 1)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  if (lo_slot0 >= 0) {
    uint8_t* dl = smem + (lo_slot0 + kb) * RC_SLOT;
    *reinterpret_cast<uint4*>(dl + sw_off(row, ch)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(dl + sw_off(row, ch + 1)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
}
// zero columns [gcol, gcol + 16) of this thread's row
__device__ __forceinline__ void zero16(uint8_t* smem, int slot0, int row, int gcol) {
  const int kb = gcol >> 6, ch = (gcol & 63) >> 3;
  uint8_t* dst = smem + (slot0 + kb) * RC_SLOT;
  *reinterpret_cast<uint4*>(dst + sw_off(row, ch)) = make_uint4(0u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(dst + sw_off(row, ch + 1)) = make_uint4(0u, 0u, 0u, 0u);
}

// Waits of the eight epilogue warps back off between polls: a spinning warp competes for issue slots with the single producer
// and MMA-issuer threads on its scheduler (measured: the issue loops ran at 0.33 us per tile whatever the tile's work).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void epi_wait_acc(Epi& e) {
  mbar_wait_sleep(e.acc_bar, e.acc_par, 128);
  e.acc_par ^= 1u;
  tc_fence_after();
  if (e.dbg_row && e.et == 0) e.dbg_row[2] = gtime();      // accumulator ready: the epilogue proper starts here
}
// the epilogue's writes (operand slots) and TMEM reads are done: hand over to the MMA issuer
__device__ __forceinline__ void epi_done(Epi& e) {
  tc_fence_before();
  fence_async_smem();
  mbar_arrive(e.act_bar);
}
// a side tile has been read by all 256 epilogue threads: give its ring slot back
__device__ __forceinline__ void tile_release(Epi& e) {
  named_bar_sync(4, EPI_THREADS);
  if (e.et == 0) mbar_arrive(&e.empty[e.ring.stage]);
  ring_adv(e.ring);
}
__device__ __forceinline__ void acc16(const Epi& e, int col, float* v) {
  uint32_t r[16];
  tmem_ld16(e.tmem + static_cast<uint32_t>(col), r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// ---- BatchNorm apply + ReLU + dropout: fp32 pre-activation tiles (ring) -> split-bf16 operand slots + mask bits ----
// (encoders.py:12-19: Linear -> BatchNorm1d -> ReLU -> Dropout(0.1); the Linear ran in the previous launch)
__device__ __noinline__ void rc_bnact(const RcPlan& plan, const RcOp& op, Epi& e) {
  // (every plan / op / thread field the loops touch is copied into registers first: the plan lives in the kernel parameter
  // space and is reached through a generic pointer here -- a field read inside a loop costs a memory round trip per iteration)
  uint8_t* const smem = e.smem;
  const float* s_mean = reinterpret_cast<const float*>(smem + RC_VEC_OFF);
  const float* s_rstd = s_mean + RC_BN_MAX;
  const int n = plan.bn_n, row = e.row, half = e.half, grow = e.grow, et = e.et;
  const bool row_ok = e.row_ok;
  const float p_drop = plan.p_drop;
  const bool drop = plan.train && p_drop > 0.f;
  const float keep_scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  const unsigned long long seed = plan.seed;
  unsigned long long offset = plan.bn_offset;
  if (plan.dyn) offset += static_cast<unsigned long long>(__ldcg(&plan.dyn->step)) << 20;
  const unsigned char* bn_keep = plan.bn_keep;
  const float* s_scale = s_rstd + RC_BN_MAX; const float* s_shift = s_rstd + 2 * RC_BN_MAX;
  unsigned short* mask_out = static_cast<unsigned short*>(const_cast<void*>(op.p[1]));
  const int side_tiles = op.side_tiles, out_slot = op.out_slot, out_lo_slot = op.out_lo_slot;
  uint64_t* const full = e.full; uint64_t* const empty = e.empty;
  Ring ring = e.ring;
#pragma unroll 1
  for (int t = 0; t < side_tiles; ++t) {
    const int col = t * 32 + half * 16;
    mbar_wait_sleep(&full[ring.stage], ring.phase, 32);
    float x[16];
    tile_row16(smem + RC_RING_OFF + ring.stage * RC_SLOT, row, half, x);
    unsigned bits = 0u;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      // (x - mean) * (rstd * gamma) + beta, as the separate launches compute it
      float y0 = fmaxf((x[j] - s_mean[col + j]) * s_scale[col + j] + s_shift[col + j], 0.f);
      float y1 = fmaxf((x[j + 1] - s_mean[col + j + 1]) * s_scale[col + j + 1] + s_shift[col + j + 1], 0.f);
      if (drop) {
        bool k0, k1;
        if (bn_keep) {
          const uchar2 k = row_ok ? *reinterpret_cast<const uchar2*>(bn_keep + static_cast<size_t>(grow) * n + col + j) : make_uchar2(0, 0);
          k0 = k.x != 0; k1 = k.y != 0;
        } else {
          const unsigned long long idx = (static_cast<unsigned long long>(grow) * n + col + j) >> 1;
          const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                                     static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                          make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
          k0 = u01(rnd.x) >= p_drop; k1 = u01(rnd.y) >= p_drop;
        }
        y0 = k0 ? y0 * keep_scale : 0.f;
        y1 = k1 ? y1 * keep_scale : 0.f;
      }
      if (!row_ok) { y0 = 0.f; y1 = 0.f; }
      x[j] = y0; x[j + 1] = y1;
      bits |= (y0 > 0.f ? 1u : 0u) << j;
      bits |= (y1 > 0.f ? 1u : 0u) << (j + 1);
    }
    put16(smem, out_slot, out_lo_slot, row, col, x);
    if (mask_out && row_ok) mask_out[static_cast<size_t>(grow) * (n >> 4) + (col >> 4)] = static_cast<unsigned short>(bits);
    named_bar_sync(4, EPI_THREADS);
    if (et == 0) mbar_arrive(&empty[ring.stage]);
    ring_adv(ring);
  }
  e.ring = ring;
  epi_done(e);
}

// ---- bias (+ ReLU) -> bf16 operand slots (hi [+ lo]) + mask bits (decoders.py:27-31: Linear -> ReLU) ----
__device__ __noinline__ void rc_relu(const RcPlan& plan, const RcOp& op, Epi& e) {
  uint8_t* const smem = e.smem;
  const float* bias = static_cast<const float*>(op.p[0]);
  unsigned short* mask_out = static_cast<unsigned short*>(const_cast<void*>(op.p[1]));
  const int e_n = op.e_n, e_tmem = op.e_tmem, e_col0 = op.e_col0, out_slot = op.out_slot, out_lo_slot = op.out_lo_slot;
  const int pitch = op.n_total >> 4;
  const bool relu = op.relu != 0;
  const int nsub = pad16(e_n) >> 4;
  const int row = e.row, grow = e.grow, half = e.half;
  const bool row_ok = e.row_ok;
  const uint32_t tmem = e.tmem;
  float bnext[16];
  vec16(bias, half * 16, e_n, bnext);                   // (in flight while the accumulator is awaited)
  epi_wait_acc(e);
#pragma unroll 1
  for (int sc = half; sc < nsub; sc += 2) {
    const int col = sc * 16;
    float bv[16], v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bv[j] = bnext[j];
    if (sc + 2 < nsub) vec16(bias, col + 32, e_n, bnext);      // the next piece's bias travels while this one is processed
    {
      uint32_t r[16];
      tmem_ld16(tmem + static_cast<uint32_t>(e_tmem + col), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    }
    unsigned bits = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float y = v[j] + bv[j];
      if (relu) y = fmaxf(y, 0.f);
      if (col + j >= e_n || !row_ok) y = 0.f;
      v[j] = y;
      bits |= (y > 0.f ? 1u : 0u) << j;
    }
    put16(smem, out_slot, out_lo_slot, row, e_col0 + col, v);
    if (mask_out && row_ok) mask_out[static_cast<size_t>(grow) * pitch + ((e_col0 + col) >> 4)] = static_cast<unsigned short>(bits);
  }
  (void)plan;
  epi_done(e);
}

// ---- data gradient: keep where the forward activation was positive (ReLU / dropout backward) -> bf16 operand slots ----
// this thread's 16 mask bits for piece sc out of the row's halfwords held in four uint4 (piece = halfword index)
__device__ __forceinline__ unsigned pick_bits(const uint4 (&m)[4], int sc) {
  const int v = sc >> 3;                                        // which uint4 (8 halfwords each)
  const uint4 q = v == 0 ? m[0] : (v == 1 ? m[1] : (v == 2 ? m[2] : m[3]));
  const int w = (sc >> 1) & 3;
  const unsigned word = w == 0 ? q.x : (w == 1 ? q.y : (w == 2 ? q.z : q.w));
  return (sc & 1) ? (word >> 16) : (word & 0xFFFFu);
}
// the whole mask row of this thread (up to 512 columns = 32 halfwords = 64 bytes), fetched before the accumulator is awaited
__device__ __forceinline__ void load_mask_row(const unsigned short* mask_in, int grow, int pitch, bool row_ok, uint4 (&m)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok && i * 8 < pitch) {
      const unsigned short* src = mask_in + static_cast<size_t>(grow) * pitch + i * 8;
      if ((pitch & 7) == 0) m[i] = *reinterpret_cast<const uint4*>(src);
      else {
        unsigned w[4] = {0u, 0u, 0u, 0u};
        for (int k = 0; k < 8 && i * 8 + k < pitch; ++k) w[k >> 1] |= static_cast<unsigned>(src[k]) << ((k & 1) * 16);
        m[i] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

__device__ __noinline__ void rc_mask(const RcPlan& plan, const RcOp& op, Epi& e) {
  uint8_t* const smem = e.smem;
  const unsigned short* mask_in = static_cast<const unsigned short*>(op.p[1]);
  const int e_n = op.e_n, e_tmem = op.e_tmem, e_col0 = op.e_col0, out_slot = op.out_slot;
  const int nsub = pad16(e_n) >> 4;
  const int pitch = op.n_total >> 4;
  const float fscale = op.fscale;
  const int row = e.row, half = e.half;
  const uint32_t tmem = e.tmem;
  uint4 mrow[4];
  load_mask_row(mask_in, e.grow, pitch, e.row_ok, mrow);
  epi_wait_acc(e);
#pragma unroll 1
  for (int sc = half; sc < nsub; sc += 2) {
    const int col = sc * 16;
    const unsigned bits = pick_bits(mrow, sc);
    float v[16];
    {
      uint32_t r[16];
      tmem_ld16(tmem + static_cast<uint32_t>(e_tmem + col), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ((bits >> j) & 1u) && col + j < e_n ? v[j] * fscale : 0.f;
    put16(smem, out_slot, -1, row, e_col0 + col, v);
  }
  (void)plan;
  epi_done(e);
}

// ---- last decoder layer + loss: value partials and dL/d(pre-activation) as the next GEMM's operand ----
// (directional_losses.py:23-24 BCE-sum on the sigmoid output, :48-49 MSE-sum; targets arrive as [128 x 32] fp32 tiles)
__device__ __noinline__ void rc_loss(const RcPlan& plan, const RcOp& op, Epi& e) {
  uint8_t* const smem = e.smem;
  const float* bias = static_cast<const float*>(op.p[0]);
  const bool bce = plan.loss_kind == LOSS_BCE;
  const int side_tiles = op.side_tiles, e_col0 = op.e_col0, e_tmem = op.e_tmem, n_total = op.n_total, out_slot = op.out_slot;
  const int row = e.row, half = e.half, et = e.et;
  const bool row_ok = e.row_ok;
  const uint32_t tmem = e.tmem;
  uint64_t* const full = e.full; uint64_t* const empty = e.empty;
  Ring ring = e.ring;
  float loss_acc = 0.f;
  float bnext[16];
  vec16(bias, e_col0 + half * 16, n_total, bnext);     // (in flight while the accumulator is awaited)
  epi_wait_acc(e);
#pragma unroll 1
  for (int t = 0; t < side_tiles; ++t) {
    const int col = t * 32 + half * 16;               // accumulator column of this piece
    const int gcol = e_col0 + col;
    float bv[16], v[16], tg[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bv[j] = bnext[j];
    if (t + 1 < side_tiles) vec16(bias, gcol + 32, n_total, bnext);
    {
      uint32_t r[16];
      tmem_ld16(tmem + static_cast<uint32_t>(e_tmem + col), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    }
    mbar_wait_sleep(&full[ring.stage], ring.phase, 32);
    tile_row16(smem + RC_RING_OFF + ring.stage * RC_SLOT, row, half, tg);
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = v[j] + bv[j];
      float l, g;
      if (bce) { float y; l = bce_from_logit(x, tg[j], y, g); }
      else { float g0; l = loss_elem<false>(x, tg[j], 1.0f, g0, g); }
      const bool ok = gcol + j < n_total && row_ok;
      part += ok ? l : 0.f;
      v[j] = ok ? g : 0.f;
    }
    loss_acc += part;
    put16(smem, out_slot, -1, row, gcol, v);
    named_bar_sync(4, EPI_THREADS);
    if (et == 0) mbar_arrive(&empty[ring.stage]);
    ring_adv(ring);
  }
  e.ring = ring;
  loss_acc = warp_sum(loss_acc);
  if (e.lane == 0) plan.loss_partials[(static_cast<size_t>(e.mb) * 2 + (op.last_loss ? 1 : 0)) * 8 + e.warp8] = loss_acc;
  epi_done(e);
  if (!op.last_loss) return;
  // ---- ticket: the CTA that finishes the step's last row block sums every partial in a fixed order ----
  int* s_flag = reinterpret_cast<int*>(e.smem + RC_BAR_OFF + 124);
  double* dsh = reinterpret_cast<double*>(e.smem + RC_DSH_OFF);
  named_bar_sync(4, EPI_THREADS);
  if (e.et == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(plan.counter, 1u);
    *s_flag = (ticket == static_cast<unsigned int>(plan.m_blocks) - 1u) ? 1 : 0;
  }
  named_bar_sync(4, EPI_THREADS);
  if (!*s_flag) return;
  __threadfence();
  double a_loss = 0, a_kl = 0;
  for (int i = e.et; i < plan.m_blocks * 16; i += EPI_THREADS) a_loss += __ldcg(plan.loss_partials + i);
  for (int i = e.et; i < plan.m_blocks * 8; i += EPI_THREADS) a_kl += __ldcg(plan.kl_partials + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a_loss += __shfl_xor_sync(0xffffffffu, a_loss, o); a_kl += __shfl_xor_sync(0xffffffffu, a_kl, o); }
  if (e.lane == 0) { dsh[e.warp8 * 2] = a_loss; dsh[e.warp8 * 2 + 1] = a_kl; }
  named_bar_sync(4, EPI_THREADS);
  if (e.et == 0) {
    double recon = 0, kl = 0;
    for (int w = 0; w < 8; ++w) { recon += dsh[w * 2]; kl += dsh[w * 2 + 1]; }
    const double beta = plan.dyn->beta_kl;
    plan.loss_out[0] = static_cast<float>(recon + beta * kl);
    plan.loss_out[1] = static_cast<float>(recon);
    plan.loss_out[2] = 0.f;
    plan.loss_out[3] = static_cast<float>(kl);
    *plan.counter = 0u;
    if (plan.dyn_bump) plan.dyn_bump->batch_index += 1;
  }
}

// ---- heads -> latent: mean over the encoders, z = mu + eps * exp(logvar / 2), KL partial (directional_vae.py:36-47) ----
// Accumulators: encoder 0 at e_tmem (mu | logvar, 2L columns; autoencoders: L columns), encoder 1 at e_tmem2.
// Output columns are owned in 32-column groups alternately by the two warps of a lane quarter; columns >= L are zeroed.
__device__ __noinline__ void rc_latent(const RcPlan& plan, const RcOp& op, Epi& e) {
  uint8_t* const smem = e.smem;
  const int L = plan.L, n_enc = plan.n_enc;
  const bool ae = plan.ae != 0;
  const float* bias0 = static_cast<const float*>(op.p[0]);
  const float* bias1 = static_cast<const float*>(op.p[1]);
  float* mu_out = static_cast<float*>(const_cast<void*>(op.p[2]));
  float* lv_out = static_cast<float*>(const_cast<void*>(op.p[3]));
  float* eps_out = static_cast<float*>(const_cast<void*>(op.p[4]));
  const float* eps_in = static_cast<const float*>(op.p[5]);
  const unsigned long long seed = plan.seed;
  unsigned long long offset = plan.lat_offset;
  if (plan.dyn && !eps_in) offset += static_cast<unsigned long long>(__ldcg(&plan.dyn->step)) << 20;
  const int kpad = ((L + 63) >> 6) << 6;                       // operand width: whole k-blocks
  const float inv_enc = 1.0f / n_enc;
  const int e_tmem = op.e_tmem, e_tmem2 = op.e_tmem2, out_slot = op.out_slot, out_lo_slot = op.out_lo_slot;
  const int row = e.row, grow = e.grow, half = e.half;
  const bool row_ok = e.row_ok;
  const uint32_t tmem = e.tmem;
  float kl = 0.f;
  // mu, logvar and eps (the backward needs them) leave through a shared-memory stage so that the global stores are whole
  // lines: [3][128 x L] fp32 in the operand slots in front of z (free now: the heads' MMAs have completed)
  float* const stg = reinterpret_cast<float*>(smem);
  const int SL = GEMM_BM * L;
  epi_wait_acc(e);
#pragma unroll 1
  for (int c8 = 0; c8 < kpad; c8 += 8) {
    if (((c8 >> 5) & 1) != half) continue;                     // this warp owns every other 32-column group
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.f;
    if (c8 < L) {
      uint32_t m0[8], v0[8], m1[8], v1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { v0[j] = 0u; m1[j] = 0u; v1[j] = 0u; }
      tmem_ld8(tmem + static_cast<uint32_t>(e_tmem + c8), m0);
      if (!ae) tmem_ld8(tmem + static_cast<uint32_t>(e_tmem + L + c8), v0);
      if (n_enc > 1) {
        tmem_ld8(tmem + static_cast<uint32_t>(e_tmem2 + c8), m1);
        if (!ae) tmem_ld8(tmem + static_cast<uint32_t>(e_tmem2 + L + c8), v1);
      }
      tmem_ld_wait();
      // every global input of the eight elements first (one round trip), then the arithmetic
      float b_mu0[8], b_mu1[8], b_lv0[8], b_lv1[8], ein[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c8 + j;
        const bool ok = col < L;
        b_mu0[j] = ok ? __ldg(bias0 + col) : 0.f;
        b_mu1[j] = (ok && n_enc > 1) ? __ldg(bias1 + col) : 0.f;
        b_lv0[j] = (ok && !ae) ? __ldg(bias0 + L + col) : 0.f;
        b_lv1[j] = (ok && !ae && n_enc > 1) ? __ldg(bias1 + L + col) : 0.f;
        ein[j] = (ok && eps_in && row_ok) ? eps_in[static_cast<unsigned>(grow) * L + col] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c8 + j;
        if (col < L) {
          // (same order of operations as the separate launches: every head adds its own bias, then the mean over the encoders)
          float mu = __uint_as_float(m0[j]) + b_mu0[j];
          float lv = __uint_as_float(v0[j]) + b_lv0[j];
          if (n_enc > 1) {
            mu = (mu + (__uint_as_float(m1[j]) + b_mu1[j])) * inv_enc;
            lv = (lv + (__uint_as_float(v1[j]) + b_lv1[j])) * inv_enc;
          }
          const unsigned idx = static_cast<unsigned>(grow) * L + col;
          const int sidx = row * L + col;
          float zz;
          if (ae) {
            zz = mu;
            stg[sidx] = mu; stg[SL + sidx] = 0.f; stg[2 * SL + sidx] = 0.f;
          } else {
            float eps = ein[j];
            if (!eps_in) {
              const uint4 rnd = philox4x32_10(make_uint4(idx, 0u, static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32) ^ 0x5EEDu),
                                              make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
              eps = normal_from(rnd.x, rnd.y);
            }
            zz = mu + eps * expf(0.5f * lv);
            stg[sidx] = mu; stg[SL + sidx] = lv; stg[2 * SL + sidx] = eps;
            if (row_ok) kl += 1.0f + lv - mu * mu - expf(lv);
          }
          z[j] = row_ok ? zz : 0.f;
        }
      }
    }
    // one 16-byte chunk (8 bf16) of the hi and the lo operand
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) hi[i] = pack_bf16x2_hi_lo(z[2 * i], z[2 * i + 1], lo[i]);
    const int kb = c8 >> 6, ch = (c8 & 63) >> 3;
    *reinterpret_cast<uint4*>(smem + (out_slot + kb) * RC_SLOT + sw_off(row, ch)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (out_lo_slot >= 0)
      *reinterpret_cast<uint4*>(smem + (out_lo_slot + kb) * RC_SLOT + sw_off(row, ch)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  kl = warp_sum(kl);
  if (e.lane == 0) plan.kl_partials[static_cast<size_t>(e.mb) * 8 + e.warp8] = -0.5f * kl;
  named_bar_sync(4, EPI_THREADS);
  {
    const int row0 = e.mb * GEMM_BM;
    const int nval = min(GEMM_BM, e.rows - row0) * L;
    const size_t g0 = static_cast<size_t>(row0) * L;
    for (int i = e.et; i < nval; i += EPI_THREADS) {
      mu_out[g0 + i] = stg[i]; lv_out[g0 + i] = stg[SL + i]; eps_out[g0 + i] = stg[2 * SL + i];
    }
  }
  epi_done(e);
}

// ---- dL/dz -> dL/d(mu | logvar) incl. beta * dKL, / n_encoders -> bf16 operand (directional_vae.py:36-47 backward) ----
// The thread's dL/dz row goes through a shared-memory scratch (free operand slots) so that any output column can pick its j.
__device__ __noinline__ void rc_latent_bwd(const RcPlan& plan, const RcOp& op, Epi& e) {
  uint8_t* const smem = e.smem;
  const int L = plan.L, n_enc = plan.n_enc;
  const bool ae = plan.ae != 0;
  const float* mu_in = static_cast<const float*>(op.p[2]);
  const float* lv_in = static_cast<const float*>(op.p[3]);
  const float* eps_in = static_cast<const float*>(op.p[4]);
  const float beta = plan.dyn ? plan.dyn->beta_kl : 0.f;
  const float inv_enc = 1.0f / n_enc;
  const int ld = L + 1;                                        // scratch pitch (floats): conflict-free row accesses
  const int e_tmem = op.e_tmem, out_slot = op.out_slot;
  float* scratch = reinterpret_cast<float*>(smem + op.out_slot2 * RC_SLOT);
  const int row = e.row, grow = e.grow, half = e.half;
  const bool row_ok = e.row_ok;
  const uint32_t tmem = e.tmem;
  // mu, logvar, eps of the row block: whole-line loads into the scratch (behind dL/dz) while the accumulator is awaited
  const bool staged = op.e_n2 != 0 && !ae;
  float* const stg = scratch + GEMM_BM * ld;
  const int SL = GEMM_BM * L;
  if (staged) {
    const int row0 = e.mb * GEMM_BM;
    const int nval = min(GEMM_BM, e.rows - row0) * L;
    const size_t g0 = static_cast<size_t>(row0) * L;
    for (int i = e.et; i < nval; i += EPI_THREADS) {
      stg[i] = mu_in[g0 + i]; stg[SL + i] = lv_in[g0 + i]; stg[2 * SL + i] = eps_in[g0 + i];
    }
  }
  epi_wait_acc(e);
  // dL/dz: this warp copies its half of the columns
#pragma unroll 1
  for (int c8 = half * 8; c8 < L; c8 += 16) {
    uint32_t r[8];
    tmem_ld8(tmem + static_cast<uint32_t>(e_tmem + c8), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c8 + j < L) scratch[row * ld + c8 + j] = __uint_as_float(r[j]);
  }
  named_bar_sync(4, EPI_THREADS);
  const int width = ae ? L : 2 * L;
  const int kpad = ((width + 63) >> 6) << 6;
#pragma unroll 1
  for (int c8 = 0; c8 < kpad; c8 += 8) {
    if (((c8 >> 5) & 1) != half) continue;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    if (c8 < width && row_ok) {
      float a0[8], a1[8], gz[8];                               // the global inputs of the eight elements first
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c8 + j;
        const int jj = col < L ? col : col - L;
        const bool ok = col < width;
        const unsigned idx = static_cast<unsigned>(grow) * L + jj;
        gz[j] = ok ? scratch[row * ld + jj] : 0.f;
        if (staged) {
          a0[j] = ok ? (col < L ? stg[row * L + jj] : stg[SL + row * L + jj]) : 0.f;
          a1[j] = (ok && col >= L) ? stg[2 * SL + row * L + jj] : 0.f;
        } else {
          a0[j] = (ok && !ae) ? (col < L ? mu_in[idx] : lv_in[idx]) : 0.f;
          a1[j] = (ok && !ae && col >= L) ? eps_in[idx] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c8 + j;
        float val = 0.f;
        if (col < width) {
          if (ae) val = gz[j] * inv_enc;
          else if (col < L) val = (gz[j] + beta * a0[j]) * inv_enc;
          else val = (gz[j] * a1[j] * 0.5f * expf(0.5f * a0[j]) + beta * 0.5f * (expf(a0[j]) - 1.0f)) * inv_enc;
        }
        g[j] = val;
      }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) hi[i] = pack_bf16x2_hi_lo(g[2 * i], g[2 * i + 1], lo[i]);
    const int kb = c8 >> 6, ch = (c8 & 63) >> 3;
    *reinterpret_cast<uint4*>(smem + (out_slot + kb) * RC_SLOT + sw_off(row, ch)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  }
  epi_done(e);
}

// column sums over the 32 lanes of v[0..15]: lane j (< 16) and lane j + 16 end up with the total of column j
__device__ __forceinline__ float warp_column_sums16(float (&v)[16], int lane) {
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

// ---- encoder data gradient: ReLU / dropout mask, BatchNorm backward statistics of the row block, dL/dy as bf16; the site
// encoder's gradient w.r.t. the gathered embedding rows beside it (encoders.py:12-19, 54-57 backward) ----
__device__ __noinline__ void rc_dgrad_enc(const RcPlan& plan, const RcOp& op, Epi& e) {
  uint8_t* const smem = e.smem;
  const float* s_mean = reinterpret_cast<const float*>(smem + RC_VEC_OFF);
  const float* s_rstd = s_mean + RC_BN_MAX;
  const unsigned short* mask_in = static_cast<const unsigned short*>(op.p[1]);
  float* stats = static_cast<float*>(const_cast<void*>(op.p[0]));
  const int n = op.e_n;                                        // BatchNorm width
  const int e_tmem = op.e_tmem, e_tmem2 = op.e_tmem2, e_n2 = op.e_n2, out_slot = op.out_slot, side_tiles = op.side_tiles;
  const float fscale = op.fscale;
  float* stage = reinterpret_cast<float*>(smem + op.out_slot2 * RC_SLOT + 8192);   // [2][4][n] column partials (upper half of the slot)
  const int row = e.row, half = e.half, et = e.et, lane = e.lane, q = e.q, mb = e.mb;
  const bool row_ok = e.row_ok;
  const uint32_t tmem = e.tmem;
  uint64_t* const full = e.full; uint64_t* const empty = e.empty;
  Ring ring = e.ring;
  uint4 mrow[4];
  load_mask_row(mask_in, e.grow, n >> 4, row_ok, mrow);
  epi_wait_acc(e);
#pragma unroll 1
  for (int t = 0; t < side_tiles; ++t) {
    const int col = t * 32 + half * 16;
    const unsigned bits = pick_bits(mrow, col >> 4);
    float v[16], pre[16];
    {
      uint32_t r[16];
      tmem_ld16(tmem + static_cast<uint32_t>(e_tmem + col), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    }
    mbar_wait_sleep(&full[ring.stage], ring.phase, 32);
    tile_row16(smem + RC_RING_OFF + ring.stage * RC_SLOT, row, half, pre);
    float s2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = ((bits >> j) & 1u) ? v[j] * fscale : 0.f;
      s2[j] = v[j] * (pre[j] - s_mean[col + j]) * s_rstd[col + j];
    }
    put16(smem, out_slot, -1, row, col, v);
    const float t1 = warp_column_sums16(v, lane);
    const float t2 = warp_column_sums16(s2, lane);
    if (lane < 16) {
      stage[(0 * 4 + q) * n + col + lane] = t1;
      stage[(1 * 4 + q) * n + col + lane] = t2;
    }
    named_bar_sync(4, EPI_THREADS);
    if (et == 0) mbar_arrive(&empty[ring.stage]);
    ring_adv(ring);
  }
  e.ring = ring;
  // the site encoder's part: plain bf16, zero padded to a whole k-block
  if (e_n2 > 0) {
#pragma unroll 1
    for (int sc = half; sc < 4; sc += 2) {
      const int col = sc * 16;
      if (col < e_n2) {
        float v[16];
        uint32_t r[16];
        tmem_ld16(tmem + static_cast<uint32_t>(e_tmem2 + col), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (col + j >= e_n2 || !row_ok) ? 0.f : __uint_as_float(r[j]);
        put16(smem, out_slot + 2, -1, row, col, v);
      } else {
        zero16(smem, out_slot + 2, row, col);
      }
    }
  }
  named_bar_sync(4, EPI_THREADS);
  for (int i = et; i < 2 * n; i += EPI_THREADS) {
    const int which = i / n, c = i - which * n;
    const float* pp = stage + (which * 4) * n + c;
    stats[(static_cast<size_t>(mb) * 2 + which) * n + c] = pp[0] + pp[n] + pp[2 * n] + pp[3 * n];
  }
  (void)plan;
  epi_done(e);
}

__global__ void __launch_bounds__(RC_THREADS, 1) rowchain_kernel(const __grid_constant__ RcPlan plan) {
  uint8_t* smem = rc_dyn_smem;
  if (smem_u32(smem) & 1023u) __trap();                // the swizzled operand blocks need 1 KiB alignment
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + RC_BAR_OFF);
  uint64_t* empty = full + RC_RING;
  uint64_t* acc_bar = empty + RC_RING;
  uint64_t* act_bar = acc_bar + 1;
  uint64_t* lda_bar = acc_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + RC_BAR_OFF + 120);
  float* s_mean = reinterpret_cast<float*>(smem + RC_VEC_OFF);
  float* s_rstd = s_mean + RC_BN_MAX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int et = static_cast<int>(threadIdx.x);             // epilogue thread index (warps 0..7)
  if (threadIdx.x == 0) {
    for (int s = 0; s < RC_RING; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_bar, 1);
    mbar_init(act_bar, EPI_THREADS);
    mbar_init(lda_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) for (int i = lane; i < plan.n_tm; i += 32) tma_prefetch_desc(&plan.tm[i]);
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  // ---- BatchNorm statistics of the whole batch from the per-tile column sums the previous launch left ----
  if (warp < 8 && plan.bn_n > 0) {
    for (int col = et; col < plan.bn_n; col += EPI_THREADS) {
      float mean, rstd;
      if (plan.train) {
        double s1 = 0, s2 = 0;
        for (int t = 0; t < plan.bn_m_tiles; ++t) {
          s1 += __ldcg(plan.bn_stats + (static_cast<size_t>(t) * 2 + 0) * plan.bn_n + col);
          s2 += __ldcg(plan.bn_stats + (static_cast<size_t>(t) * 2 + 1) * plan.bn_n + col);
        }
        const double m = s1 / plan.rows;
        double var = s2 / plan.rows - m * m;
        var = var < 0 ? 0 : var;
        mean = static_cast<float>(m);
        rstd = rsqrtf(static_cast<float>(var) + 1e-5f);
        if (blockIdx.x == 0) {
          const double unbiased = plan.rows > 1 ? var * plan.rows / (plan.rows - 1) : var;
          plan.bn_running_mean[col] = 0.9f * plan.bn_running_mean[col] + 0.1f * mean;
          plan.bn_running_var[col] = 0.9f * plan.bn_running_var[col] + 0.1f * static_cast<float>(unbiased);
        }
      } else {
        mean = plan.bn_running_mean[col];
        rstd = 1.0f / sqrtf(plan.bn_running_var[col] + 1e-5f);
      }
      s_mean[col] = mean; s_rstd[col] = rstd;
      {
        const float sc = rstd * plan.bn_gamma[col];
        s_rstd[RC_BN_MAX + col] = sc;                               // scale
        s_rstd[2 * RC_BN_MAX + col] = plan.bn_beta[col];             // beta
      }
      if (blockIdx.x == 0) { plan.bn_save_mean[col] = mean; plan.bn_save_rstd[col] = rstd; }
    }
    if (blockIdx.x == 0 && et == 0 && plan.train && plan.bn_nbt) *plan.bn_nbt += 1;
  }
  const long long tgt_row0 = plan.n_batches > 1 ? static_cast<long long>(plan.dyn->batch_index % plan.n_batches) * plan.rows : 0;
  __syncthreads();

  Ring ring{0, 0u};
  uint32_t acc_par = 0u, act_par = 0u, lda_par = 0u;
  for (int mb = blockIdx.x; mb < plan.m_blocks; mb += gridDim.x) {
    const int row0 = mb * GEMM_BM;
    if (warp == 8) {
      // =========================== TMA producer ===========================
      // (op fields are copied into registers before the tile loops: one constant-bank read each, not one per tile)
      if (lane == 0) {
        for (int i = 0; i < plan.n_ops; ++i) {
          const RcOp& op = plan.ops[i];
          const int kind = op.kind;
          if (kind == RC_LOADA) {
            const int kbs = op.kb, a_slot = op.a_slot, a_lo_slot = op.a_lo_slot, b_lo = op.b_lo;
            const CUtensorMap* tm = &plan.tm[op.tm_b];
            mbar_expect_tx(lda_bar, kbs * (a_lo_slot >= 0 ? 2 : 1) * RC_SLOT);
            for (int kb = 0; kb < kbs; ++kb) {
              tma_load_2d(smem + (a_slot + kb) * RC_SLOT, tm, lda_bar, kb * 64, row0);
              if (a_lo_slot >= 0) tma_load_2d(smem + (a_lo_slot + kb) * RC_SLOT, tm, lda_bar, b_lo + kb * 64, row0);
            }
          } else if (kind == RC_GEMM) {
            const int n = op.n, n0 = op.n0, kbs = op.kb, b_lo = op.b_lo;
            const bool nn = op.nn != 0;
            const int passes = (!nn && op.a_lo_slot >= 0) ? 2 : 1;
            const CUtensorMap* tm = &plan.tm[op.tm_b];
            for (int c0 = 0; c0 < n; c0 += 128) {
              const int nb = (min(128, n - c0) + 63) >> 6;
              for (int kb = 0; kb < kbs; ++kb) {
                if (plan.pad2 & 2) {            // experiment: no weight loads (the MMAs run on whatever the slots hold)
                  for (int ps = 0; ps < passes; ++ps) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1u);
                    mbar_arrive(&full[ring.stage]);
                    ring_adv(ring);
                  }
                } else if (!nn) {
                  for (int ps = 0; ps < passes; ++ps) {
                    mbar_wait(&empty[ring.stage], ring.phase ^ 1u);
                    mbar_expect_tx(&full[ring.stage], RC_SLOT);
                    tma_load_2d(smem + RC_RING_OFF + ring.stage * RC_SLOT, tm, &full[ring.stage], kb * 64 + (ps ? b_lo : 0), n0 + c0);
                    ring_adv(ring);
                  }
                } else {
                  mbar_wait(&empty[ring.stage], ring.phase ^ 1u);
                  mbar_expect_tx(&full[ring.stage], nb * 8192);
                  for (int bx = 0; bx < nb; ++bx)
                    tma_load_2d(smem + RC_RING_OFF + ring.stage * RC_SLOT + bx * 8192, tm, &full[ring.stage], n0 + c0 + bx * 64, kb * 64);
                  ring_adv(ring);
                }
              }
            }
          } else if (op.side_tiles > 0) {
            const int side_tiles = op.side_tiles, e_col0 = op.e_col0;
            const CUtensorMap* tm = &plan.tm[op.tm_side];
            const int srow = static_cast<int>((kind == RC_EPI && op.sub == EP_LOSS) ? tgt_row0 + row0 : row0);
            for (int t = 0; t < side_tiles; ++t) {
              mbar_wait(&empty[ring.stage], ring.phase ^ 1u);
              mbar_expect_tx(&full[ring.stage], RC_SLOT);
              tma_load_2d(smem + RC_RING_OFF + ring.stage * RC_SLOT, tm, &full[ring.stage], e_col0 + t * 32, srow);
              ring_adv(ring);
            }
          }
        }
      }
    } else if (warp == 9) {
      // =========================== MMA issuer (+ TMA stores of finished operands) ===========================
      if (lane == 0) {
        for (int i = 0; i < plan.n_ops; ++i) {
          const RcOp& op = plan.ops[i];
          const int kind = op.kind;
          if (kind == RC_GEMM) {
            const int n = op.n, kbs = op.kb, a_slot = op.a_slot, a_lo_slot = op.a_lo_slot, tmem_col = op.tmem_col;
            const bool nn = op.nn != 0, commit = op.commit != 0;
            const bool split = !nn && a_lo_slot >= 0;
            if (op.wait_lda) { mbar_wait(lda_bar, lda_par); lda_par ^= 1u; }
            tc_fence_after();
            if (plan.dbg) plan.dbg[(static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 + 0] = gtime();
            // descriptors advance by plain additions to the address field (14 bits of address >> 4: no carry out below 256 KiB):
            // +2 per 16 bf16 of K inside a K-major swizzle atom, +128 per 16 K-rows of an MN-major tile
            const uint64_t dk = make_smem_desc(0u, 16, 1024), dmn = make_smem_desc(0u, 8192, 1024);
            const uint64_t a0 = dk + (smem_u32(smem + a_slot * RC_SLOT) >> 4);
            const uint64_t a0lo = dk + (smem_u32(smem + (split ? a_lo_slot : a_slot) * RC_SLOT) >> 4);
            const uint64_t b0 = (nn ? dmn : dk) + (smem_u32(smem + RC_RING_OFF) >> 4);
            const uint32_t bstep = nn ? 128u : 2u;
            for (int c0 = 0; c0 < n; c0 += 128) {
              const int cn = min(128, n - c0);
              const uint32_t idesc = make_idesc_bf16(GEMM_BM, pad16(cn), 0, nn ? 1 : 0);
              const uint32_t d = tmem_base + static_cast<uint32_t>(tmem_col + c0);
              for (int kb = 0; kb < kbs; ++kb) {
                const uint64_t a_hi = a0 + static_cast<uint32_t>(kb * (RC_SLOT >> 4));
                mbar_wait(&full[ring.stage], ring.phase);
                tc_fence_after();
                const uint64_t sb = b0 + static_cast<uint32_t>(ring.stage * (RC_SLOT >> 4));
                if (plan.pad2 & 1) {            // experiment: no MMAs (the tiles are only awaited and released; 4: by a plain arrive)
                  if (split) {
                    if (plan.pad2 & 4) mbar_arrive(&empty[ring.stage]); else umma_commit(&empty[ring.stage]);
                    ring_adv(ring); mbar_wait(&full[ring.stage], ring.phase);
                  }
                  if (plan.pad2 & 4) mbar_arrive(&empty[ring.stage]); else umma_commit(&empty[ring.stage]);
                  ring_adv(ring);
                  continue;
                }
                umma_bf16(d, a_hi, sb, idesc, kb > 0 ? 1u : 0u);
                umma_bf16(d, a_hi + 2, sb + bstep, idesc, 1u);
                umma_bf16(d, a_hi + 4, sb + 2 * bstep, idesc, 1u);
                umma_bf16(d, a_hi + 6, sb + 3 * bstep, idesc, 1u);
                if (split) {
                  const uint64_t a_lo = a0lo + static_cast<uint32_t>(kb * (RC_SLOT >> 4));
                  umma_bf16(d, a_lo, sb, idesc, 1u);
                  umma_bf16(d, a_lo + 2, sb + 2, idesc, 1u);
                  umma_bf16(d, a_lo + 4, sb + 4, idesc, 1u);
                  umma_bf16(d, a_lo + 6, sb + 6, idesc, 1u);
                  umma_commit(&empty[ring.stage]);
                  ring_adv(ring);
                  mbar_wait(&full[ring.stage], ring.phase);
                  tc_fence_after();
                  const uint64_t sl = b0 + static_cast<uint32_t>(ring.stage * (RC_SLOT >> 4));
                  umma_bf16(d, a_hi, sl, idesc, 1u);
                  umma_bf16(d, a_hi + 2, sl + 2, idesc, 1u);
                  umma_bf16(d, a_hi + 4, sl + 4, idesc, 1u);
                  umma_bf16(d, a_hi + 6, sl + 6, idesc, 1u);
                }
                umma_commit(&empty[ring.stage]);
                ring_adv(ring);
              }
            }
            if (plan.dbg) plan.dbg[(static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 + 1] = gtime();
            if (commit) {
              bulk_wait_read();                   // the stores of the previous operands have left shared memory
              umma_commit(acc_bar);               // -> the epilogue may read the accumulator and overwrite the slots
            }
          } else if (kind == RC_EPI || kind == RC_BNACT) {
            ring_adv(ring, op.side_tiles);
            RcStore st[3] = {op.st[0], op.st[1], op.st[2]};
            mbar_wait(act_bar, act_par);
            act_par ^= 1u;
            tc_fence_after();
            bool any = false;
            for (int sidx = 0; sidx < 3; ++sidx)
              for (int kb = 0; kb < st[sidx].kb; ++kb) { tma_store_2d(&plan.tm[st[sidx].tm], smem + (st[sidx].slot + kb) * RC_SLOT, kb * 64, row0); any = true; }
            if (any) bulk_commit();
          }
        }
        bulk_wait_read();
      }
    } else {
      // =========================== epilogue / element-wise warps ===========================
      Epi e;
      e.smem = smem; e.full = full; e.empty = empty; e.acc_bar = acc_bar; e.act_bar = act_bar;
      e.q = warp & 3; e.half = warp >> 2; e.lane = lane; e.et = et; e.warp8 = warp;
      e.tmem = tmem_base + (static_cast<uint32_t>(e.q * 32) << 16);
      e.row = e.q * 32 + lane; e.grow = row0 + e.row; e.row_ok = e.grow < plan.rows;
      e.mb = mb; e.rows = plan.rows; e.tgt_row0 = tgt_row0;
      e.ring = ring; e.acc_par = acc_par; e.dbg_row = nullptr;
      for (int i = 0; i < plan.n_ops; ++i) {
        const RcOp& op = plan.ops[i];
        if (op.kind == RC_GEMM) {
          int tiles = 0;
          for (int c0 = 0; c0 < op.n; c0 += 128) tiles += op.kb * ((!op.nn && op.a_lo_slot >= 0) ? 2 : 1);
          ring_adv(e.ring, tiles);
        } else if (op.kind == RC_BNACT) {
          if (plan.dbg && et == 0) plan.dbg[(static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 + 2] = gtime();
          rc_bnact(plan, op, e);
          if (plan.dbg && et == 0) plan.dbg[(static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 + 3] = gtime();
        } else if (op.kind == RC_EPI) {
          if (plan.dbg && et == 0) plan.dbg[(static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 + 2] = gtime();
          e.dbg_row = plan.dbg ? plan.dbg + (static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 : nullptr;
          switch (op.sub) {
            case EP_LATENT: rc_latent(plan, op, e); break;
            case EP_RELU: rc_relu(plan, op, e); break;
            case EP_LOSS: rc_loss(plan, op, e); break;
            case EP_MASK: rc_mask(plan, op, e); break;
            case EP_LATENT_BWD: rc_latent_bwd(plan, op, e); break;
            default: rc_dgrad_enc(plan, op, e); break;
          }
          if (plan.dbg && et == 0) plan.dbg[(static_cast<size_t>(blockIdx.x) * RC_MAX_OPS + i) * 4 + 3] = gtime();
        }
      }
      ring = e.ring; acc_par = e.acc_par;
    }
    // every role walked the same op list: re-synchronise the ring position for the roles that only skipped
    {
      int tiles = 0, accs = 0, acts = 0, ldas = 0;
      for (int i = 0; i < plan.n_ops; ++i) {
        const RcOp& op = plan.ops[i];
        if (op.kind == RC_GEMM) {
          for (int c0 = 0; c0 < op.n; c0 += 128) tiles += op.kb * ((!op.nn && op.a_lo_slot >= 0) ? 2 : 1);
          accs += op.commit ? 1 : 0;
          ldas += op.wait_lda ? 1 : 0;
        } else if (op.kind == RC_EPI || op.kind == RC_BNACT) {
          tiles += op.side_tiles; acts += 1;
        }
      }
      if (mb == static_cast<int>(blockIdx.x)) { /* first block: nothing carried over yet */ }
      // positions after this block, computed identically by every thread
      Ring r2{0, 0u};
      const int blocks_done = (mb - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1;
      const long long total = static_cast<long long>(tiles) * blocks_done;
      r2.stage = static_cast<int>(total % RC_RING);
      r2.phase = static_cast<uint32_t>((total / RC_RING) & 1);
      ring = r2;
      acc_par = static_cast<uint32_t>((static_cast<long long>(accs) * blocks_done) & 1);
      act_par = static_cast<uint32_t>((static_cast<long long>(acts) * blocks_done) & 1);
      lda_par = static_cast<uint32_t>((static_cast<long long>(ldas) * blocks_done) & 1);
    }
    tc_fence_before();
    __syncthreads();               // operand slots and TMEM are free for the next row block
    tc_fence_after();
  }
  if (warp == 9) {
    if (lane == 0) bulk_wait_all();
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

cudaError_t launch_rowchain(const RcPlan& plan, int n_ctas, cudaStream_t s) {
  static cudaError_t attr = cudaFuncSetAttribute(rowchain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RC_SMEM_BYTES);
  if (attr != cudaSuccess) return attr;
  return launch_pdl(rowchain_kernel, dim3(n_ctas), dim3(RC_THREADS), RC_SMEM_BYTES, s, plan);
}

}  // namespace vla
