// One 128 x BN output tile of the grouped bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator in
// TMEM), operands staged by TMA into 128-byte-swizzled shared memory.  The body is shared by the stand-alone grouped
// GEMM kernel (gemm_tc.cu: one tile per CTA) and the chain kernel (chain_kernel.cu: a CTA of a 4-CTA cluster walks the
// row-local layers of one 128-row block; the shared-memory ring, the barriers and the TMEM allocation live across tiles).
//
//   mode 0 NT: C[M,N] = A[M,K] * B[N,K]^T   both operands K-major       (Linear forward)
//   mode 1 TN: C[M,N] = A[K,M]^T * B[K,N]   both operands MN-major      (weight gradients, split-K + red.add)
//   mode 2 NN: C[M,N] = A[M,K] * B[K,N]     A K-major, B MN-major       (data gradients: B = the forward weight copy)
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue.
// Epilogue (8 warps, thread = accumulator row): TMEM -> registers, bias / mask / activation applied in registers with
// the global operands prefetched as 128-bit vectors, per-column statistics (BatchNorm forward and backward) by a
// shuffle butterfly, 128-bit row stores; split-K partials go through a smem transpose so every red.add is coalesced.
#pragma once
#include "loss_math.cuh"
#include "tc_ptx.cuh"
#include "vla_internal.h"

#include <cfloat>

namespace vla {

constexpr int A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;                 // 16 KiB
constexpr int B_STAGE_BYTES = GEMM_BN_MAX_NT * 128;                  // 18 KiB (>= 2 x 8 KiB boxes of the MN-major modes)
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;           // 34 KiB: one SLOT = one k-block of A and B
constexpr int ONES_OFFSET = GEMM_SLOTS * STAGE_BYTES;                // 2 KiB of bf16 1.0
constexpr int ONES_BYTES = 2048;
constexpr int PATCH_LD = 36;                                         // floats; 144-byte rows: 16-byte aligned, conflict-free
constexpr int CHUNK_OFFSET = 0;                                      // 8 warps x fp32 [32][36] transpose patches: they alias
constexpr int CHUNK_BYTES = 8 * 32 * PATCH_LD * 4;                   // the first slots, idle once the accumulator is ready
constexpr int TGT_OFFSET = 2 * STAGE_BYTES;                          // GF_LOSS: per-warp target patches, up to 3 chunks per warp,
constexpr int TGT_WARP_BYTES = 3 * 32 * PATCH_LD * 4;                // in the later slots (idle once the accumulator is ready)
constexpr int VEC_LD = GEMM_BN_MAX_NT;                               // widest tile of any mode
constexpr int VEC_OFFSET = ONES_OFFSET + ONES_BYTES;                 // bias | mean | rstd, fp32 [3][VEC_LD]
constexpr int VEC_BYTES = 3 * VEC_LD * 4;
constexpr int PART_OFFSET = VEC_OFFSET + VEC_BYTES;                  // column-stat partials fp32 [2][5][4][32]
constexpr int MAX_CHUNKS = (GEMM_BN_MAX_NT + 31) / 32;
constexpr int PART_BYTES = 2 * MAX_CHUNKS * 4 * 32 * 4;
constexpr int BAR_OFFSET = PART_OFFSET + PART_BYTES;                 // mbarriers: full[S] | empty[S] | acc | dep | ew | tmem slot
constexpr int SMEM_USED = BAR_OFFSET + 128;
constexpr int SMEM_BYTES = SMEM_USED + 1024;                         // slack for manual 1 KiB alignment
constexpr int EPI_THREADS = GEMM_THREADS - 64;                       // 8 warps

static_assert(GEMM_BN_MAX_TN <= GEMM_BN_MAX_NT && B_STAGE_BYTES >= (GEMM_BN_MAX_TN / 64) * 8192, "B slot too small for the MN-major tiles");
static_assert(CHUNK_BYTES <= TGT_OFFSET, "epilogue patches and loss-target patches must not overlap");
static_assert(TGT_OFFSET + 8 * TGT_WARP_BYTES <= GEMM_SLOTS * STAGE_BYTES, "loss-target patches must fit in the idle slots");
static_assert((GEMM_BN_MAX_NT / 32 + 1 + 1) / 2 <= 3, "a warp stages the targets of at most 3 chunks");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(STAGE_BYTES % 1024 == 0 && A_STAGE_BYTES % 1024 == 0, "swizzle atoms need 1 KiB alignment");
static_assert(GEMM_BN_MAX_NT % 16 == 0 && GEMM_BN_MAX_TN % 64 == 0, "tile limits");
static_assert((2 * GEMM_STAGES + 5) * 8 + 4 <= 128, "barrier block");
// Persistent form (gemm_tc_persist_kernel): a CTA walks a list of tiles with TWO accumulators in TMEM -- the MMA issuer fills
// one while the epilogue warps drain the other -- and the producer runs ahead into the next tile's operands.  The ring then has
// PERSIST_STAGES stages (slots 0..3); the epilogue's transpose patches move out of the ring into the space of slots 4..5.
constexpr int PERSIST_STAGES = 2;
constexpr int PERSIST_CHUNK_OFFSET = PERSIST_STAGES * GEMM_GROUP * STAGE_BYTES;
static_assert(PERSIST_CHUNK_OFFSET + CHUNK_BYTES <= GEMM_SLOTS * STAGE_BYTES, "persistent form: patches must fit behind the ring");

constexpr int FEATS_FWD_PLAIN = GF_BIAS | GF_RELU | GF_OUT_F32 | GF_OUT_BF16;                      // hidden decoder layers, heads
constexpr int FEATS_FWD_FULL = FEATS_FWD_PLAIN | GF_SIGMOID | GF_COLSTATS;                         // + BatchNorm statistics / sigmoid
constexpr int FEATS_FWD_LOSS = FEATS_FWD_PLAIN | GF_SIGMOID | GF_LOSS | GF_LK_MSE | GF_LK_BCE | GF_LK_CE;   // last decoder layers of a
                                                                                                   // train step, any mix of loss kinds
// The same epilogue with ONE loss kind and no fp32 output compiled in, for groups that are uniform (rna2dna: BCE, dna2rna:
// MSE).  The generic instantiation is 84 KB of SASS, the size at which the epilogue is instruction-cache bound.
constexpr int FEATS_FWD_LOSS_BCE = GF_BIAS | GF_SIGMOID | GF_OUT_BF16 | GF_LOSS | GF_LK_BCE;
constexpr int FEATS_FWD_LOSS_MSE = GF_BIAS | GF_OUT_BF16 | GF_LOSS | GF_LK_MSE;
constexpr int FEATS_DGRAD_PLAIN = GF_MASK | GF_OUT_F32 | GF_OUT_BF16;                              // decoder data gradients
constexpr int FEATS_DGRAD_FULL = FEATS_DGRAD_PLAIN | GF_BNSTATS;                                   // + BatchNorm backward statistics
constexpr int FEATS_DGRAD_LAT = GF_LATBWD;                                                         // dL/dz -> d(mu | logvar) in the epilogue
constexpr int FEATS_WGRAD = GF_RED | GF_BIASGRAD;

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Sum over the 32 lanes of v[j] for every j; the total of column j ends up in lane j (31 shuffles).
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
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

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// 1 KiB-aligned base of the dynamic shared memory (SWIZZLE_128B atoms).  Derived from the __shared__ symbol with an
// integer offset so that the compiler keeps the shared state space (LDS / STS instead of generic LD / ST).
extern __shared__ uint8_t vla_dyn_smem[];
__device__ __forceinline__ uint8_t* aligned_smem() {
  const uint32_t pad = (1024u - (smem_u32(vla_dyn_smem) & 1023u)) & 1023u;
  return vla_dyn_smem + pad;
}

// What persists across the tiles a CTA processes (no shared-memory pointers: they are re-derived from aligned_smem()).
// Barriers at BAR_OFFSET: full[GEMM_STAGES] | empty[GEMM_STAGES] | acc (MMA -> epilogue) | dep (whole-step kernel:
// dependencies of this tile resolved, producer -> epilogue) | TMEM slot.
struct TileCtx {
  uint32_t tmem_base;
  int stage; uint32_t phase;  // position in the shared-memory ring (producer and MMA roles advance identically)
  uint32_t tile_parity;       // parity of acc_bar / dep_bar for the tile in flight
  uint32_t tile_seq;          // tiles this CTA has processed (persistent form: accumulator = tile_seq & 1)
  uint32_t ew_parity;         // parity of the element-wise bulk-load barrier
  unsigned long long* dbg;    // optional [.][8] globaltimer stamps
  int dbg_row;
  int dbg_flags;              // test hook: 1 = skip epilogue stores, 2 = skip main loop, ...
};

__device__ __forceinline__ uint64_t* tile_bars(uint8_t* smem) { return reinterpret_cast<uint64_t*>(smem + BAR_OFFSET); }
__device__ __forceinline__ uint64_t* tile_ew_bar(uint8_t* smem) { return tile_bars(smem) + 2 * GEMM_STAGES + 2; }

// One-time CTA setup shared by both kernels: barriers, bf16 ones for the bias-gradient MMA, TMEM.
// Ends with a CTA-wide barrier; returns the context with the ring at its origin.
// Barrier block: full[S] | empty[S] | acc (persistent: acc_full[0]) | dep (acc_full[1]) | ew (acc_empty[0]) | acc_empty[1] | TMEM slot
__device__ __forceinline__ TileCtx tile_setup(bool alloc_tmem, bool fill_ones = true, uint32_t tmem_cols = GEMM_TMEM_COLS,
                                               uint32_t empty_count = 1) {
  TileCtx c;
  uint8_t* smem = aligned_smem();
  uint64_t* full_bar = tile_bars(smem);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* acc_bar = empty_bar + GEMM_STAGES;
  uint64_t* dep_bar = acc_bar + 1;
  uint64_t* ew_bar = dep_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ew_bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(ew_bar + 1, 1);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], empty_count);      // (multicast ring: one release per CTA of the cluster)
    }
    mbar_init(acc_bar, 1);
    mbar_init(dep_bar, 1);
    mbar_init(ew_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1 && alloc_tmem) tmem_alloc(tmem_slot, tmem_cols);
  if (warp >= 2 && fill_ones) {
    // 2 KiB of bf16 1.0: the B operand of the bias-gradient MMA (layout-invariant)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + ONES_OFFSET);
    for (int i = threadIdx.x - 64; i < ONES_BYTES / 4; i += EPI_THREADS) ones[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = alloc_tmem ? *tmem_slot : 0u;
  c.stage = 0; c.phase = 0; c.tile_parity = 0; c.ew_parity = 0; c.tile_seq = 0;
  c.dbg = nullptr; c.dbg_row = 0; c.dbg_flags = 0;
  return c;
}

#define VLA_STAMP(slot) do { if (ctx.dbg) ctx.dbg[static_cast<size_t>(ctx.dbg_row) * 8 + (slot)] = gtime(); } while (0)

// FEATS: compile-time superset of the epilogue flags that may occur; everything else is compiled out (smaller code:
// the epilogue is instruction-fetch sensitive).  All threads of the CTA call this; on return the tile's global writes
// have been issued by the epilogue threads (the caller orders them: barrier + fence) and ctx has advanced.
// CL > 1 (NT mode, gemm_tc_cluster_kernel): the CL CTAs of a cluster work on CL neighbouring n-tiles of ONE row block.  Each
// loads 1 / CL of the A tile (GEMM_BM / CL rows) and MULTICASTS it into every CTA's slot, so an SM takes in A / CL + B per
// k-block instead of A + B (the main loop is bound by the per-SM L2 -> shared-memory intake: profiles/r2_ring_depth_experiment.md).
// A slot is reused only when all CL CTAs have consumed it: every MMA issuer's commit releases the stage in all CL CTAs.
template <int MODE, int FEATS, bool PERSIST = false, int CL = 1>
__device__ __forceinline__ void gemm_tile(TileCtx& ctx, const GemmScalars& Pd, const CUtensorMap* tmA, const CUtensorMap* tmB,
                                          int m_tile, int n_tile, int k_split, const LossTail* tail_desc = nullptr) {
  uint8_t* smem = aligned_smem();
  uint64_t* full_bar = tile_bars(smem);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  static_assert(!PERSIST || !(FEATS & GF_LOSS), "the loss epilogue stages its targets in the ring's slots: not persistent");
  static_assert(CL == 1 || (MODE == 0 && !PERSIST && GEMM_BM % CL == 0), "multicast ring: NT mode, one tile per CTA");
  constexpr int NS = PERSIST ? PERSIST_STAGES : GEMM_STAGES;          // ring stages in use
  const uint32_t acc_buf = PERSIST ? (ctx.tile_seq & 1u) : 0u;        // which accumulator this tile uses
  const uint32_t acc_par = PERSIST ? ((ctx.tile_seq >> 1) & 1u) : ctx.tile_parity;
  uint64_t* acc_bar = empty_bar + GEMM_STAGES + acc_buf;              // accumulator complete (MMA -> epilogue)
  uint64_t* acc_free = empty_bar + GEMM_STAGES + 2 + acc_buf;         // persistent: accumulator drained (epilogue -> MMA)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Snapshot of the descriptor's scalar fields (kernel parameter space, or the chain kernel's shared-memory phase image):
  // the epilogue's stores could alias the descriptor, which would force a reload after every store.
  struct {
    int M, N, K, BN, m_tiles, n_tiles, kb_per_split, flags, ld_f32, ld_bf16, ld_mask, ld_pre;
    float mask_scale;
    const float* bias; float* out_f32; bf16* out_bf16; const bf16* mask_src; const float* pre; const float* mean;
    const float* rstd; float* stats; float* bias_grad;
    const float* aux0; const float* aux1; const long long* aux_site; float* aux_partials; const DynParams* dyn;
    int aux_n, loss_kind; float aux_scale;
    int a_lo, b_lo, out_lo;
    unsigned int* mask_bits_out; const unsigned int* mask_bits_in;
  } P;
  P.M = Pd.M; P.N = Pd.N; P.K = Pd.K; P.BN = Pd.BN; P.m_tiles = Pd.m_tiles; P.n_tiles = Pd.n_tiles;
  P.kb_per_split = Pd.kb_per_split; P.flags = Pd.flags; P.ld_f32 = Pd.ld_f32; P.ld_bf16 = Pd.ld_bf16; P.ld_mask = Pd.ld_mask;
  P.ld_pre = Pd.ld_pre; P.mask_scale = Pd.mask_scale; P.bias = Pd.bias; P.out_f32 = Pd.out_f32; P.out_bf16 = Pd.out_bf16;
  P.a_lo = (MODE == 0) ? Pd.a_lo : 0; P.b_lo = (MODE == 0) ? Pd.b_lo : 0; P.out_lo = (MODE == 0) ? Pd.out_lo : 0;
  P.mask_bits_out = Pd.mask_bits_out; P.mask_bits_in = Pd.mask_bits_in;
  P.mask_src = Pd.mask_src; P.pre = Pd.pre; P.mean = Pd.mean; P.rstd = Pd.rstd; P.stats = Pd.stats; P.bias_grad = Pd.bias_grad;
  if (FEATS & (GF_LOSS | GF_LATBWD)) {
    P.aux0 = Pd.aux0; P.aux1 = Pd.aux1; P.aux_site = Pd.aux_site; P.aux_partials = Pd.aux_partials; P.dyn = Pd.dyn;
    P.aux_n = Pd.aux_n; P.loss_kind = Pd.loss_kind; P.aux_scale = Pd.aux_scale;
  }
  const int local = (k_split * P.m_tiles + m_tile) * P.n_tiles + n_tile;      // index of the tile inside its problem
  const int m0 = m_tile * GEMM_BM;
  const int n0 = n_tile * P.BN;
  const int BN = P.BN;
  const int kb_total = (P.K + GEMM_BK - 1) / GEMM_BK;
  const int kb0 = k_split * P.kb_per_split;
  const int kb1 = min(kb0 + P.kb_per_split, kb_total);
  // split-bf16 operands (NT only): every k-block occupies TWO consecutive ring slots -- (A_hi, B_hi) then (A_lo, B_lo) --
  // and the tile accumulates A_hi B_hi + A_lo B_hi + A_hi B_lo (the lo x lo term is below fp32 rounding).
  const bool split = (MODE == 0) && P.a_lo > 0;
  const bool bias_mma = (MODE == 1) && (FEATS & GF_BIASGRAD) && (P.flags & GF_BIASGRAD) && n_tile == 0;
  const int dbgf = ctx.dbg_flags;                 // test hooks (0 outside vla_test_gemm)
  const uint32_t tmem_base = ctx.tmem_base + acc_buf * GEMM_TMEM_COLS;

  if (dbgf & 2) {
    // test hook: no main loop, no epilogue
  } else if (warp == 0) {
    // =========================== TMA producer ===========================
    // A load unit = one k-block of A and B (split operands: the hi or the lo copy of a k-block) = one shared-memory slot.
    // GEMM_GROUP consecutive units share ONE full / empty barrier pair (one hand-shake).  One elected thread (see the MMA issuer).
    if (elect_one()) {
      int stage = ctx.stage;
      uint32_t phase = ctx.phase;
      const int u0 = (split ? 2 : 1) * kb0, u1 = (split ? 2 : 1) * kb1;
      const int nb = BN >> 6;
      const uint32_t unit_bytes = A_STAGE_BYTES + (MODE == 0 ? BN * 128 : nb * 8192);
      for (int ug = u0; ug < u1; ug += GEMM_GROUP) {
        const int cnt = min(GEMM_GROUP, u1 - ug);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], cnt * unit_bytes);
        for (int j = 0; j < cnt; ++j) {
          const int kv = ug + j;
          const int kb = split ? (kv >> 1) : kv;
          const int lo = split ? (kv & 1) : 0;                                   // second unit of a split k-block: the lo copies
          uint8_t* sa = smem + (stage * GEMM_GROUP + j) * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (MODE == 1) {
            tma_load_2d(sa, tmA, &full_bar[stage], m0, kb * GEMM_BK);           // box 64 (M) x 64 (K)
            tma_load_2d(sa + 8192, tmA, &full_bar[stage], m0 + 64, kb * GEMM_BK);
          } else if (CL > 1) {
            // this CTA's share of the row block, to every CTA of the cluster (tmA's box is 64 (K) x GEMM_BM / CL rows)
            const uint32_t rk = cluster_ctarank();
            tma_load_2d_mc(sa + rk * (A_STAGE_BYTES / CL), tmA, &full_bar[stage], kb * GEMM_BK + (lo ? P.a_lo : 0),
                           m0 + static_cast<int>(rk) * (GEMM_BM / CL), static_cast<uint16_t>((1u << CL) - 1u));
          } else {
            tma_load_2d(sa, tmA, &full_bar[stage], kb * GEMM_BK + (lo ? P.a_lo : 0), m0);   // box 64 (K) x 128 (M)
          }
          if (MODE == 0) {
            tma_load_2d(sb, tmB, &full_bar[stage], kb * GEMM_BK + (lo ? P.b_lo : 0), n0);   // box 64 (K) x BN (N)
          } else {
            for (int i = 0; i < nb; ++i)
              tma_load_2d(sb + i * 8192, tmB, &full_bar[stage], n0 + i * 64, kb * GEMM_BK);   // box 64 (N) x 64 (K)
          }
        }
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // One ELECTED thread (elect.sync: the compiler then keeps the whole loop on the uniform datapath without a vote loop around
    // every tcgen05.mma); operand descriptors are built once per slot and ADVANCED along K by adding to the start-address field
    // (16-byte units, no carry: shared memory < 256 KB).  The issue loop used to cost ~116 cycles per MMA -- more than the MMA
    // itself for tiles narrower than 224 columns -- and bounded every main loop (profiles/r2_mma_issue.md).
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, MODE == 1 ? 1 : 0, MODE == 0 ? 0 : 1);
      const uint32_t idesc_ones = make_idesc_bf16(GEMM_BM, 16, 1, 0);
      const uint64_t odesc = make_smem_desc(smem_u32(smem + ONES_OFFSET), 16, 1024);
      int stage = ctx.stage;
      uint32_t phase = ctx.phase;
      const int u0 = (split ? 2 : 1) * kb0, u1 = (split ? 2 : 1) * kb1;
      if (PERSIST) {                     // the epilogue of the tile before last has drained this accumulator
        mbar_wait(acc_free, acc_par ^ 1u);
        tc_fence_after();
      }
      for (int ug = u0; ug < u1; ug += GEMM_GROUP) {
        const int cnt = min(GEMM_GROUP, u1 - ug);
        mbar_wait(&full_bar[stage], phase);
        if (ug == u0) VLA_STAMP(3);                                 // first operands landed
        tc_fence_after();
        if (split) {
          // the two slots of the stage = (A_hi, B_hi) and (A_lo, B_lo) of one k-block; three passes of four MMAs
          const uint32_t a_hi = smem_u32(smem + (stage * GEMM_GROUP) * STAGE_BYTES);
          const uint64_t dah = make_smem_desc(a_hi, 16, 1024), dbh = make_smem_desc(a_hi + A_STAGE_BYTES, 16, 1024);
          const uint64_t dal = dah + (STAGE_BYTES >> 4), dbl = dbh + (STAGE_BYTES >> 4);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t da = pass == 1 ? dal : dah, db = pass == 2 ? dbl : dbh;
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)        // +16 bf16 of K inside the swizzle atom = +32 bytes = +2 units
              umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (ug > u0 || k > 0 || pass > 0) ? 1u : 0u);
          }
        } else {
          for (int j = 0; j < cnt; ++j) {
            const uint32_t sa = smem_u32(smem + (stage * GEMM_GROUP + j) * STAGE_BYTES);
            const uint32_t sb = sa + A_STAGE_BYTES;
            // MN-major: +16 K-rows of 128 B = +128 units per MMA; K-major: +16 bf16 of K inside the swizzle atom = +2 units
            const uint64_t a0 = MODE == 1 ? make_smem_desc(sa, 8192, 1024) : make_smem_desc(sa, 16, 1024);
            const uint64_t b0 = MODE == 0 ? make_smem_desc(sb, 16, 1024) : make_smem_desc(sb, 8192, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              const uint64_t adesc = a0 + (MODE == 1 ? 128 : 2) * k;
              const uint64_t bdesc = b0 + (MODE == 0 ? 2 : 128) * k;
              const uint32_t acc = (ug > u0 || j > 0 || k > 0) ? 1u : 0u;
              umma_bf16(tmem_base, adesc, bdesc, idesc, acc);
              if (bias_mma) {
                umma_bf16(tmem_base + GEMM_BIAS_TMEM_COL, adesc, odesc, idesc_ones, acc);
              }
            }
          }
        }
        if (CL > 1) umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>((1u << CL) - 1u));
        else umma_commit(&empty_bar[stage]);   // frees the stage's slots when these MMAs retire
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_bar);           // accumulator complete
      VLA_STAMP(4);                                                // all MMAs issued
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    // Thread = one accumulator row (TMEM lane); the two warps that share a lane quarter take alternate
    // 32-column chunks.  Everything stays in registers; global operands are fetched as 128-bit vectors
    // before they are needed; column statistics use a shuffle butterfly.
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // 0: warps 2-5, 1: warps 6-9
    const int et = threadIdx.x - 64;        // 0..255
    float* vec = reinterpret_cast<float*>(smem + VEC_OFFSET);
    float* part = reinterpret_cast<float*>(smem + PART_OFFSET);
    const int flags = P.flags & FEATS;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < P.M;
    const int n_chunks = (BN + 31) >> 5;
    const bool want_stats = (flags & (GF_COLSTATS | GF_BNSTATS)) != 0;

    // per-column epilogue vectors of this tile (BatchNorm statistics come from an earlier unit of the step)
    for (int i = et; i < BN; i += EPI_THREADS) {
      const int col = n0 + i;
      const bool ok = col < P.N;
      vec[i] = (ok && (P.flags & GF_BIAS)) ? P.bias[col] : 0.f;
      if (FEATS & GF_BNSTATS) {
        vec[VEC_LD + i] = (ok && (P.flags & GF_BNSTATS)) ? __ldcg(P.mean + col) : 0.f;
        vec[2 * VEC_LD + i] = (ok && (P.flags & GF_BNSTATS)) ? __ldcg(P.rstd + col) : 0.f;
      }
    }
    named_bar_sync(1, EPI_THREADS);

    // GF_LOSS (MSE / BCE): the targets stream from HBM.  While the main loop runs, pull this warp's target lines into
    // L2 (no registers held): the epilogue's loads then see L2 latency instead of DRAM latency.
    long long tgt_row0 = 0;
    if ((FEATS & GF_LOSS) && (P.flags & GF_LOSS)) {
      if (P.aux_n > 1) tgt_row0 = static_cast<long long>(P.dyn->batch_index % P.aux_n) * P.M;
      if (P.loss_kind != LOSS_CE && row_ok) {
        const float* trow = P.aux0 + (tgt_row0 + row) * P.N;
        for (int c = half; c < n_chunks; c += 2) {
          const int c_lo = n0 + c * 32, c_hi = min(c_lo + 31, min(P.N, n0 + BN) - 1);
          if (c_lo <= c_hi) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(trow + c_lo));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(trow + c_hi));
          }
        }
      }
    }

    // GF_MASK with bit masks: this warp's words (one per chunk, at most 3) travel while the main loop runs
    uint32_t mbits[3] = {0u, 0u, 0u};
    const bool bit_mask = (FEATS & GF_MASK) && (P.flags & GF_MASK) && P.mask_bits_in != nullptr;
    if (bit_mask && row_ok) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c = half + 2 * k;
        if (c < n_chunks && n0 + c * 32 < P.N) mbits[k] = __ldcg(P.mask_bits_in + static_cast<size_t>((n0 >> 5) + c) * P.M + row);
      }
    }

    // GF_BNSTATS: the fp32 pre-activations this warp will read (32 rows x 128 bytes per chunk, L2-resident) are pulled into
    // L1 while the main loop runs -- no registers held (holding them across the wait spills: profiles/r2_mma_issue.md)
    if ((FEATS & GF_BNSTATS) && (P.flags & GF_BNSTATS) && !(dbgf & 1)) {
      const int prow = m0 + q * 32 + lane;
      if (prow < P.M)
        for (int c = half; c < n_chunks; c += 2)
          if (n0 + c * 32 < P.N)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(P.pre + static_cast<size_t>(prow) * P.ld_pre + n0 + c * 32));
    }

    // GF_LATBWD: d(mu) = dz + ca, d(logvar) = dz * cb + cc with ca = beta mu, cb = eps exp(logvar / 2) / 2,
    // cc = beta (exp(logvar) - 1) / 2 (latent_bwd_elem).  The two warps of a lane quarter SHARE every 32-column chunk: each reads
    // the whole accumulator chunk but works on 16 of its 32 rows, lane = column (every global access is a contiguous piece of
    // one row).  The coefficients of the first chunk are loaded and computed here, before the accumulator wait -- 48 registers
    // per thread, all the exponentials off the critical path; later chunks (latent widths above 32) do it inside the loop.
    constexpr bool LAT = (FEATS & GF_LATBWD) != 0;
    float lat_a[16], lat_b[16], lat_c[16];
    auto lat_load = [&](int c) {
      const int col0 = n0 + c * 32;
      const int nvalid = min(32, min(P.N, n0 + BN) - col0);
      if (nvalid <= 0 || P.loss_kind != 0) return;                   // (warp-uniform; autoencoder: no coefficients)
      // no predicates: rows past M and lanes past the chunk read a valid neighbour and are ignored later
      const int cc = col0 + min(lane, nvalid - 1);
      const int rb = m0 + q * 32 + 16 * half;
      const float beta = P.dyn ? P.dyn->beta_kl : P.aux_scale;
      float mu[16], lv[16], ep[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const unsigned off = static_cast<unsigned>(min(rb + i, P.M - 1)) * static_cast<unsigned>(P.N) + cc;
        mu[i] = __ldg(P.aux0 + off); lv[i] = __ldg(P.aux1 + off); ep[i] = __ldg(P.pre + off);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        lat_a[i] = beta * mu[i];
        lat_b[i] = ep[i] * 0.5f * expf(0.5f * lv[i]);
        lat_c[i] = beta * 0.5f * (expf(lv[i]) - 1.0f);
      }
    };
    if (LAT && (P.flags & GF_LATBWD)) lat_load(0);

    mbar_wait(acc_bar, acc_par);
    tc_fence_after();
    if (et == 0) VLA_STAMP(5);                                     // accumulator ready

    // Per-warp transpose patch: global traffic of the epilogue is always issued with consecutive lanes on consecutive
    // addresses of one row (1 L1 wavefront per 128 bytes) instead of 32 rows per instruction.
    float* patch = reinterpret_cast<float*>(smem + (PERSIST ? PERSIST_CHUNK_OFFSET : CHUNK_OFFSET)) + (warp - 2) * (32 * PATCH_LD);
    const int rbase = m0 + q * 32;                      // first accumulator row of this warp
    const bool f32_vec = ((reinterpret_cast<uintptr_t>(P.out_f32) & 15) == 0) && ((P.ld_f32 & 3) == 0);
    const bool bf16_vec = ((reinterpret_cast<uintptr_t>(P.out_bf16) & 15) == 0) && ((P.ld_bf16 & 7) == 0);

    float loss_acc = 0.f;                                // GF_LOSS: this thread's share of the loss value
    // GF_LOSS (MSE / BCE): stage the targets of ALL this warp's chunks now -- 4-byte asynchronous copies straight into the
    // transposed position of a per-chunk patch (coalesced: a warp instruction reads 128 contiguous bytes of one row; the
    // lines were prefetched into L2 during the main loop).  One commit group per chunk: the first chunk waits for an L2
    // round trip, the later ones find their targets in shared memory; no registers are held.
    float* tgt_patch = reinterpret_cast<float*>(smem + TGT_OFFSET + (warp - 2) * TGT_WARP_BYTES);
    int tgt_groups = 0;
    if ((FEATS & GF_LOSS) && (flags & GF_LOSS) && P.loss_kind != LOSS_CE && !(dbgf & 1)) {
      for (int c = half; c < n_chunks; c += 2, ++tgt_groups) {
        const int col0 = n0 + c * 32;
        const int nvalid = min(32, min(P.N, n0 + BN) - col0);
        if (nvalid > 0) {
          const float* tp = P.aux0 + (tgt_row0 + rbase) * P.N + col0 + lane;
          const uint32_t dst = smem_u32(tgt_patch + tgt_groups * (32 * PATCH_LD) + lane);
          if ((P.N & 3) == 0 && (reinterpret_cast<uintptr_t>(P.aux0) & 15) == 0) {
            // rows are 16-byte aligned: 16-byte copies, eight lanes per row, four rows per instruction (a quarter of the
            // instructions of the 4-byte path; nvalid is then a multiple of 4)
            const float* tp16 = P.aux0 + (tgt_row0 + rbase) * P.N + col0 + (lane & 7) * 4;
            const uint32_t dst16 = smem_u32(tgt_patch + tgt_groups * (32 * PATCH_LD) + (lane >> 3) * PATCH_LD + (lane & 7) * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              const bool ok = rbase + rr < P.M && (lane & 7) * 4 < nvalid;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst16 + i * (4 * PATCH_LD * 4)),
                           "l"(ok ? tp16 + static_cast<size_t>(rr) * P.N : P.aux0), "r"(ok ? 16 : 0) : "memory");
            }
          } else {
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
              const bool ok = rbase + i < P.M && lane < nvalid;
              cp_async4_zfill(dst + i * (PATCH_LD * 4), ok ? tp + static_cast<size_t>(i) * P.N : P.aux0, ok ? 4 : 0);
            }
          }
        }
        cp_async_commit();
      }
    }
    int tgt_k = 0;                                       // index of the chunk in flight among this warp's chunks

    for (int c = LAT ? 0 : half; c < ((dbgf & 1) ? 0 : n_chunks); c += LAT ? 1 : 2) {
      const int my_k = tgt_k++;                         // which of this warp's staged target patches belongs to this chunk
      const int col0 = n0 + c * 32;
      const int nvalid = min(32, min(P.N, n0 + BN) - col0);   // <= 0: nothing to store (tile padding); a tile never touches
                                                              // columns of its right-hand neighbour (BN % 32 may be 16)
      const bool full = nvalid == 32;
      uint32_t r[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32);
      if (!(dbgf & 16)) {
        tmem_ld16(taddr, r);
        tmem_ld16(taddr + 16, r + 16);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = lane + j;
      }
      // ---- per-row operands: coalesced loads -> patch -> this thread's row ----
      uint4 mk[4];
      float4 pr[8];
      const bool mask_fast = (flags & GF_MASK) != 0 && !bit_mask;     // host side guarantees N % 32 == 0 and 16-byte aligned rows
      const bool pre_fast = (flags & GF_BNSTATS) != 0;
      if (mask_fast) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {                   // 8 rows x 64 bytes per instruction
          const int rr = 8 * k + (lane >> 2);
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (rbase + rr < P.M)
            t = __ldg(reinterpret_cast<const uint4*>(P.mask_src + static_cast<size_t>(rbase + rr) * P.ld_mask + col0) + (lane & 3));
          *reinterpret_cast<uint4*>(patch + rr * PATCH_LD + (lane & 3) * 4) = t;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) mk[i] = *reinterpret_cast<const uint4*>(patch + lane * PATCH_LD + i * 4);
        __syncwarp();
      }
      if (pre_fast) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {                   // 4 rows x 128 bytes per instruction
          const int rr = 4 * k + (lane >> 3);
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rbase + rr < P.M)
            t = __ldg(reinterpret_cast<const float4*>(P.pre + static_cast<size_t>(rbase + rr) * P.ld_pre + col0) + (lane & 7));
          *reinterpret_cast<float4*>(patch + rr * PATCH_LD + (lane & 7) * 4) = t;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) pr[i] = *reinterpret_cast<const float4*>(patch + lane * PATCH_LD + i * 4);
        __syncwarp();
      }
      if (!(dbgf & 16)) tmem_ld_wait();
      // One flag-uniform pass over the register row per feature (compact code: no per-element branching).
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(vec + c * 32 + j);
        v[j] = __uint_as_float(r[j]) + bv.x;         v[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
        v[j + 2] = __uint_as_float(r[j + 2]) + bv.z; v[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
      }
      if ((FEATS & GF_LATBWD) && (flags & GF_LATBWD)) {
        // v = dL/dz of this row chunk.  Backward of z = mu + eps * exp(logvar / 2) and of the KL term, then of the mean over
        // the modalities (vae.py:11-15, 60-66; losses.py:44; directional_vae.py:47-52) -- the arithmetic of latent_bwd_elem.
        // The row goes through the warp's transpose patch; then lane = column, this warp's 16 rows (coefficients: lat_load).
        if (my_k > 0) lat_load(c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(patch + lane * PATCH_LD + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        __syncwarp();
        if (lane < nvalid) {
          const bool ae = P.loss_kind != 0;
          const float nmod = static_cast<float>(P.aux_n);
          const int r16 = rbase + 16 * half;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (r16 + i < P.M) {
              const size_t idx = static_cast<size_t>(r16 + i) * P.N + col0 + lane;
              bf16* gp = P.out_bf16 + static_cast<size_t>(r16 + i) * P.ld_bf16 + col0 + lane;
              const float gz = patch[(16 * half + i) * PATCH_LD + lane];
              if (ae) {
                float g = gz + (P.mean ? __ldg(P.mean + idx) : 0.f);
                if (P.aux_n > 1) g /= nmod;
                *gp = __float2bfloat16(g);
              } else {
                float gmu = gz + lat_a[i];
                float glv = gz * lat_b[i] + lat_c[i];
                if (P.mean) gmu += __ldg(P.mean + idx);
                if (P.rstd) glv += __ldg(P.rstd + idx);
                if (P.aux_n > 1) { gmu /= nmod; glv /= nmod; }
                gp[0] = __float2bfloat16(gmu);
                gp[P.N] = __float2bfloat16(glv);
              }
            }
          }
        }
        __syncwarp();
        continue;
      }
      if (bit_mask) {                                  // ReLU / dropout backward from the forward's (activation > 0) bits
        const float sc = P.mask_scale;
        const uint32_t word = my_k == 0 ? mbits[0] : (my_k == 1 ? mbits[1] : mbits[2]);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((word >> j) & 1u) ? v[j] * sc : 0.f;
      } else if (flags & GF_MASK) {                    // ... or from the saved bf16 activation itself
        const float sc = P.mask_scale;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint32_t word = reinterpret_cast<const uint32_t*>(mk)[j >> 1];
          const float m = __uint_as_float((j & 1) ? (word & 0xFFFF0000u) : (word << 16));
          v[j] = m > 0.f ? v[j] * sc : 0.f;
        }
      }
      if (want_stats) {
        // statistics are taken before the activation (BatchNorm forward) / on the masked gradient (backward)
        float s1[32], s2[32];
        if (flags & GF_BNSTATS) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 mv = *reinterpret_cast<const float4*>(vec + VEC_LD + c * 32 + j);
            const float4 rv = *reinterpret_cast<const float4*>(vec + 2 * VEC_LD + c * 32 + j);
            const float4 pv = pr[j >> 2];
            s2[j] = v[j] * (pv.x - mv.x) * rv.x;         s2[j + 1] = v[j + 1] * (pv.y - mv.y) * rv.y;
            s2[j + 2] = v[j + 2] * (pv.z - mv.z) * rv.z; s2[j + 3] = v[j + 3] * (pv.w - mv.w) * rv.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) s2[j] = v[j] * v[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1[j] = row_ok ? v[j] : 0.f; s2[j] = row_ok ? s2[j] : 0.f; }
        const float t1 = warp_column_sums(s1, lane);
        const float t2 = warp_column_sums(s2, lane);
        part[((0 * MAX_CHUNKS + c) * 4 + q) * 32 + lane] = t1;
        part[((1 * MAX_CHUNKS + c) * 4 + q) * 32 + lane] = t2;
      }
      if ((flags & GF_RELU) && !(dbgf & 32)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        if (MODE == 0 && P.mask_bits_out != nullptr && nvalid > 0 && row_ok) {
          // (value > 0) per element as one word per row and 32-column chunk, [chunk][M]: the ReLU mask of the backward
          uint32_t bits = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) bits |= (v[j] > 0.f && j < nvalid ? 1u : 0u) << j;
          P.mask_bits_out[static_cast<size_t>(col0 >> 5) * P.M + row] = bits;
        }
      }
      const bool bce_fused = (FEATS & GF_LOSS) && (flags & GF_LOSS) && P.loss_kind == LOSS_BCE;
      if ((flags & GF_SIGMOID) && !(dbgf & 32) && !(bce_fused && !(flags & GF_OUT_F32))) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j]);      // (a fused BCE loss works on the pre-sigmoid value)
      }
      if (nvalid <= 0) continue;
      if (dbgf & 8) {            // test hook: keep the math alive, skip patch + global stores
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
        if (acc == 123.456f) P.out_f32[0] = acc;
        continue;
      }
      // ---- this thread's row -> patch; then row-contiguous global accesses ----
      if (flags & (GF_RED | GF_OUT_F32 | GF_OUT_BF16) && !((FEATS & GF_LOSS) && (flags & GF_LOSS) && !(flags & GF_OUT_F32))) {
        // (a loss tile that stores no fp32 output writes the patch only once, with the gradient, below)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(patch + lane * PATCH_LD + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        __syncwarp();
      }
      if (flags & GF_RED) {
        // split-K partial sums: one 128-byte red per row
        if (lane < nvalid) {
#pragma unroll 4
          for (int i = 0; i < 32; ++i)
            if (rbase + i < P.M)
              red_add_f32(P.out_f32 + static_cast<size_t>(rbase + i) * P.ld_f32 + col0 + lane, patch[i * PATCH_LD + lane]);
        }
      } else {
        if (flags & GF_OUT_F32) {
          if (full && f32_vec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {               // 4 rows x 128 bytes per instruction
              const int rr = 4 * k + (lane >> 3);
              if (rbase + rr < P.M)
                reinterpret_cast<float4*>(P.out_f32 + static_cast<size_t>(rbase + rr) * P.ld_f32 + col0)[lane & 7] =
                    *reinterpret_cast<const float4*>(patch + rr * PATCH_LD + (lane & 7) * 4);
            }
          } else if (lane < nvalid) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i)
              if (rbase + i < P.M) P.out_f32[static_cast<size_t>(rbase + i) * P.ld_f32 + col0 + lane] = patch[i * PATCH_LD + lane];
          }
        }
        if ((FEATS & GF_LOSS) && (flags & GF_LOSS)) {
          // v holds the model output of this row chunk.  Loss value into loss_acc, dL/d(pre-activation) into v (then
          // stored as the bf16 operand of the backward GEMMs).  losses.py:27-42; directional_losses.py:23-24, 48-49.
          __syncwarp();
          if ((FEATS & GF_LK_CE) && P.loss_kind == LOSS_CE) {
            // weighted cross-entropy: the whole logit row (N <= 32) is in this thread
            float mx = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) mx = fmaxf(mx, v[j]);
            float se = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) se += expf(v[j] - mx);
            const float lse = logf(se) + mx;
            const int t = row_ok ? static_cast<int>(P.aux_site[tgt_row0 + row]) : 0;
            const float w = P.aux1 ? P.aux1[t] : 1.0f;
            const float gamma = P.dyn ? P.dyn->gamma : P.aux_scale;
            float xt = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j == t) xt = v[j];
              v[j] = (expf(v[j] - lse) - (j == t ? 1.0f : 0.0f)) * (w * gamma);
            }
            if (row_ok) loss_acc += -w * (xt - lse);
          } else {
            // targets: staged above by asynchronous copies, one commit group per chunk (oldest first)
            {
              const int pending = tgt_groups - 1 - my_k;
              if (pending >= 2) cp_async_wait<2>(); else if (pending == 1) cp_async_wait<1>(); else cp_async_wait<0>();
            }
            __syncwarp();
            float tg[32];
            {
              const float* tp_s = tgt_patch + my_k * (32 * PATCH_LD) + lane * PATCH_LD;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 q4 = *reinterpret_cast<const float4*>(tp_s + i * 4);
                tg[4 * i] = q4.x; tg[4 * i + 1] = q4.y; tg[4 * i + 2] = q4.z; tg[4 * i + 3] = q4.w;
              }
            }
            float part = 0.f;
            if ((FEATS & GF_LK_BCE) && P.loss_kind == LOSS_BCE && !(flags & GF_OUT_F32)) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {                      // v = pre-sigmoid value
                float y, g1;
                const float l = bce_from_logit(v[j], tg[j], y, g1);
                part += (j < nvalid) ? l : 0.f;
                v[j] = g1;
              }
            } else if ((FEATS & GF_LK_BCE) && (FEATS & GF_OUT_F32) && P.loss_kind == LOSS_BCE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {                      // v = sigmoid output (also stored as fp32 above)
                float g0, g1;
                const float l = loss_elem<true>(v[j], tg[j], 1.0f, g0, g1);
                part += (j < nvalid) ? l : 0.f;
                v[j] = g1;
              }
            } else if (FEATS & GF_LK_MSE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float g0, g1;
                const float l = loss_elem<false>(v[j], tg[j], 1.0f, g0, g1);
                part += (j < nvalid) ? l : 0.f;
                v[j] = g1;
              }
            }
            if (row_ok) loss_acc += part;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(patch + lane * PATCH_LD + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          __syncwarp();
        }
        if (flags & GF_OUT_BF16) {
          if (full && bf16_vec) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {               // 8 rows x 64 bytes per instruction
              const int rr = 8 * k + (lane >> 2);
              if (rbase + rr < P.M) {
                const float4 lo = *reinterpret_cast<const float4*>(patch + rr * PATCH_LD + (lane & 3) * 8);
                const float4 hi = *reinterpret_cast<const float4*>(patch + rr * PATCH_LD + (lane & 3) * 8 + 4);
                __nv_bfloat162 b0 = __floats2bfloat162_rn(lo.x, lo.y), b1 = __floats2bfloat162_rn(lo.z, lo.w);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(hi.x, hi.y), b3 = __floats2bfloat162_rn(hi.z, hi.w);
                uint4 o;
                o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
                o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
                reinterpret_cast<uint4*>(P.out_bf16 + static_cast<size_t>(rbase + rr) * P.ld_bf16 + col0)[lane & 3] = o;
                if (MODE == 0 && !(FEATS & GF_LOSS) && P.out_lo > 0) {      // lo copy: what bf16 rounding dropped
                  const float2 f0 = __bfloat1622float2(b0), f1 = __bfloat1622float2(b1);
                  const float2 f2 = __bfloat1622float2(b2), f3 = __bfloat1622float2(b3);
                  b0 = __floats2bfloat162_rn(lo.x - f0.x, lo.y - f0.y); b1 = __floats2bfloat162_rn(lo.z - f1.x, lo.w - f1.y);
                  b2 = __floats2bfloat162_rn(hi.x - f2.x, hi.y - f2.y); b3 = __floats2bfloat162_rn(hi.z - f3.x, hi.w - f3.y);
                  o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
                  o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
                  reinterpret_cast<uint4*>(P.out_bf16 + static_cast<size_t>(rbase + rr) * P.ld_bf16 + P.out_lo + col0)[lane & 3] = o;
                }
              }
            }
          } else if (lane < nvalid) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i)
              if (rbase + i < P.M) {
                const float x = patch[i * PATCH_LD + lane];
                const bf16 h = __float2bfloat16(x);
                bf16* dst = P.out_bf16 + static_cast<size_t>(rbase + i) * P.ld_bf16 + col0 + lane;
                *dst = h;
                if (MODE == 0 && !(FEATS & GF_LOSS) && P.out_lo > 0) dst[P.out_lo] = __float2bfloat16(x - __bfloat162float(h));
              }
          }
        }
      }
      __syncwarp();
    }
    if ((FEATS & GF_LOSS) && (flags & GF_LOSS)) {
      // ---- loss partials of this tile, ticket, and (last tile of the step) the fixed-order final reduction ----
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
      if (lane == 0) P.aux_partials[static_cast<size_t>(local) * 8 + (warp - 2)] = loss_acc;
      int* s_flag = reinterpret_cast<int*>(part);                  // column-stat scratch is unused by loss tiles
      double* dsh = reinterpret_cast<double*>(part + 8);
      const LossTail T = *tail_desc;
      if (T.counter != nullptr) {          // (nullptr: the step's AdamW launch sums the partials -- nothing more to do here)
      named_bar_sync(3, EPI_THREADS);
      if (et == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(T.counter, 1u);
        *s_flag = (ticket == static_cast<unsigned int>(T.total_tickets) - 1u) ? 1 : 0;
      }
      named_bar_sync(3, EPI_THREADS);
      if (*s_flag) {
        __threadfence();
        // All four terms at once: every thread sums its (fixed) share of each partial array, one shuffle butterfly per
        // term, then thread 0 adds the eight warp results in order -- deterministic, and one barrier instead of the 36 a
        // per-term shared-memory tree needs (this runs on the LAST tile of the step: it is pure critical path).
        const int starts[4] = {0, T.n_mse, T.n_mse + T.n_bce, T.n_mse + T.n_bce + T.n_ce};
        double acc4[4] = {0, 0, 0, 0};                             // mse, bce, ce, kl
        for (int role = 0; role < 3; ++role)
          for (int i = starts[role] + et; i < starts[role + 1]; i += EPI_THREADS) acc4[role] += __ldcg(T.partials + i);
        for (int i = et; i < T.n_kl; i += EPI_THREADS) acc4[3] += __ldcg(T.kl_partials + i);
#pragma unroll
        for (int role = 0; role < 4; ++role) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc4[role] += __shfl_xor_sync(0xffffffffu, acc4[role], o);
          if (lane == 0) dsh[(warp - 2) * 4 + role] = acc4[role];
        }
        named_bar_sync(3, EPI_THREADS);
        double sums[4] = {0, 0, 0, 0};
        if (et == 0)
          for (int w = 0; w < EPI_THREADS / 32; ++w)
            for (int role = 0; role < 4; ++role) sums[role] += dsh[w * 4 + role];
        if (et == 0) {
          const double beta = T.dyn->beta_kl, gamma = T.dyn->gamma;
          const double recon = sums[0] + sums[1];
          T.out[0] = static_cast<float>(recon + gamma * sums[2] + beta * sums[3]);
          T.out[1] = static_cast<float>(recon);
          T.out[2] = static_cast<float>(sums[2]);
          T.out[3] = static_cast<float>(sums[3]);
          *T.counter = 0u;                                         // re-arm for the next step (graph replays)
          if (T.dyn_bump) T.dyn_bump->batch_index += 1;
        }
      }
      }
    }
    if (et == 0) VLA_STAMP(6);                                     // this thread's chunks are done
    if (et == EPI_THREADS - 32) VLA_STAMP(7);                      // ... and the last epilogue warp's
    if (bias_mma && half == 0) {
      uint32_t r1[1];
      tmem_ld1(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + GEMM_BIAS_TMEM_COL, r1);
      tmem_ld_wait();
      if (row_ok) red_add_f32(P.bias_grad + row, __uint_as_float(r1[0]));
    }
    if (want_stats) {
      named_bar_sync(3, EPI_THREADS);
      for (int i = et; i < 2 * BN; i += EPI_THREADS) {       // (cc < BN: only this tile's own columns)
        const int which = i / BN, cc = i - which * BN;
        const int col = n0 + cc;
        if (col < P.N) {
          const float* pp = part + ((which * MAX_CHUNKS + (cc >> 5)) * 4) * 32 + (cc & 31);
          P.stats[(static_cast<size_t>(m_tile) * 2 + which) * P.N + col] = pp[0] + pp[32] + pp[64] + pp[96];
        }
      }
    }
    if (PERSIST) {
      // every epilogue thread has read its accumulator rows and is done with this tile's column vectors / partials: hand
      // the accumulator back to the MMA issuer; the next tile may overwrite vec / part
      tc_fence_before();
      named_bar_sync(2, EPI_THREADS);
      if (et == 0) mbar_arrive(acc_free);
    }
  }

  // ---- every thread: advance the ring position and the tile parity ----
  if (!(dbgf & 2)) {
    const int units = (split ? 2 : 1) * (kb1 - kb0);
    const int s = ctx.stage + (units + GEMM_GROUP - 1) / GEMM_GROUP;
    ctx.phase ^= static_cast<uint32_t>(s / NS) & 1u;
    ctx.stage = s % NS;
    ctx.tile_parity ^= 1u;
    ctx.tile_seq += 1u;
  }
}

}  // namespace vla
