// Whole-step kernel: the fused train step (ingest -> encoders -> latent -> decoders -> loss -> data gradients ->
// weight gradients -> AdamW) as ONE persistent cooperative launch.  The step is a list of phases -- exactly the
// launches of the per-call path, same device bodies (gemm_tile.cuh, elementwise_dev.cuh) -- and every phase is split
// into units (one 128 x BN GEMM tile, or a few element-wise blocks).  A CTA walks the phases in order and runs the
// units assigned to it; instead of a kernel boundary between phases a unit waits until the 128-row blocks it reads
// have been completed by the phases it depends on (per-row-block counters in global memory, released with
// fence + atomic, acquired by one polling warp).  Row blocks therefore stream through the layers independently, the
// ~3 us launch + ramp cost per phase disappears, and TMEM / barriers / tensor-map prefetches are set up once.
//
// Deadlock freedom: all CTAs are co-resident (cooperative launch, 1 CTA per SM), every CTA runs its units in phase
// order, and a unit only waits on units of earlier phases.
//
// Replaces the loop body at train_rna2dna.py:82-99 / optimize_hyperparameters.py:104-113 of the reference.
#include "elementwise_dev.cuh"
#include "gemm_tile.cuh"

namespace vla {

namespace {

constexpr int PLAN_OFFSET = SMEM_USED;                                  // StepPlan copy
constexpr int PLAN_BYTES = (sizeof(StepPlan) + 127) & ~127;
constexpr int SCRATCH_OFFSET = PLAN_OFFSET + PLAN_BYTES;                // element-wise scratch
constexpr int STEP_SMEM_USED = SCRATCH_OFFSET + ((EW_SCRATCH_BYTES + 127) & ~127);
constexpr int STEP_SMEM_BYTES = STEP_SMEM_USED + 1024;
static_assert(STEP_SMEM_BYTES <= 227 * 1024, "whole-step kernel shared memory budget");
static_assert(SMEM_USED % 16 == 0, "plan copy alignment");

__host__ __device__ inline const GemmProblem* find_problem(const GemmGroup* g, int u) {
  int pi = 0;
  for (int i = 1; i < g->nprob; ++i)
    if (u >= g->p[i].tile_begin) pi = i;
  return &g->p[pi];
}

__host__ __device__ inline void loss_block_rows(const LossGrid& G, int rows, int b, int* r0, int* r1) {
  if (b < G.nb_a) { *r0 = b * LOSS_WARPS; *r1 = *r0 + LOSS_WARPS; }
  else if ((b -= G.nb_a) < G.nb_b) { *r0 = b * LOSS_WARPS; *r1 = *r0 + LOSS_WARPS; }
  else if ((b -= G.nb_b) < G.nb_c) { *r0 = b * LOSS_THREADS; *r1 = *r0 + LOSS_THREADS; }
  else { *r0 = 0; *r1 = rows; }
  if (*r1 > rows) *r1 = rows;
}

__host__ __device__ inline void unit_rows(const StepPhase& ph, const void* args, int u, int* r0, int* r1) {
  switch (ph.kind) {
    case SK_GEMM_NT_PLAIN: case SK_GEMM_NT_FULL: case SK_GEMM_NT_LOSS: case SK_GEMM_NN_PLAIN: case SK_GEMM_NN_FULL: {
      const GemmProblem* P = find_problem(static_cast<const GemmGroup*>(args), u);
      const int local = u - P->tile_begin;
      const int m_tile = (local / P->n_tiles) % P->m_tiles;
      *r0 = m_tile * GEMM_BM; *r1 = min(*r0 + GEMM_BM, P->M);
      break;
    }
    case SK_GEMM_TN: {
      const GemmProblem* P = find_problem(static_cast<const GemmGroup*>(args), u);
      const int local = u - P->tile_begin;
      const int k_split = local / (P->n_tiles * P->m_tiles);
      *r0 = k_split * P->kb_per_split * GEMM_BK;
      *r1 = min(*r0 + P->kb_per_split * GEMM_BK, P->K);
      break;
    }
    case SK_INGEST:
      *r0 = u * ph.rpb; *r1 = min(*r0 + ph.rpb, ph.rows);
      break;
    case SK_BN_ACT: case SK_BN_BWD: {
      const int by = u / ph.gx;
      *r0 = by * ph.rpb; *r1 = min(*r0 + ph.rpb, ph.rows);
      break;
    }
    case SK_LATENT_FWD: case SK_LATENT_BWD: {
      const long long e0 = static_cast<long long>(u) * ph.sub * EW_THREADS;
      const long long e1 = e0 + static_cast<long long>(ph.sub) * EW_THREADS;
      *r0 = static_cast<int>(e0 / ph.L);
      const long long last = (e1 - 1) / ph.L + 1;
      *r1 = last > ph.rows ? ph.rows : static_cast<int>(last);
      if (*r0 > ph.rows) *r0 = ph.rows;
      break;
    }
    case SK_LOSS: {
      const LossGrid G = loss_grid(*static_cast<const LossArgs*>(args));
      int lo = ph.rows, hi = 0;
      for (int b = u * ph.sub; b < min((u + 1) * ph.sub, ph.n_blocks); ++b) {
        int a0, a1;
        loss_block_rows(G, ph.rows, b, &a0, &a1);
        lo = min(lo, a0); hi = max(hi, a1);
      }
      *r0 = lo; *r1 = hi;
      break;
    }
    default:   // SK_ADAMW: no rows
      *r0 = 0; *r1 = 0;
  }
  if (*r1 < *r0) *r1 = *r0;
}

__device__ __forceinline__ void red_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Dependencies of one unit, resolved by one warp: lane l polls counters l, l + 32, ... of each range.
struct UnitDeps {
  const unsigned int* counters;
  const unsigned int* targets;
  int n;
  int lo[STEP_MAX_DEPS], hi[STEP_MAX_DEPS];      // inclusive counter index ranges
  __device__ __forceinline__ void wait(int lane) const {
    for (int d = 0; d < n; ++d) {
      for (int i = lo[d] + lane; i <= hi[d]; i += 32) {
        const unsigned int tgt = targets[i];
        unsigned int spins = 0;
        while (ld_acquire(counters + i) < tgt) {
          if (++spins > (1u << 22)) __trap();    // a protocol bug traps (launch error on the host) instead of hanging
        }
      }
    }
    __syncwarp();
  }
};

__device__ __forceinline__ UnitDeps make_deps(const StepPlan& pl, const StepPhase& ph, int r0, int r1) {
  UnitDeps d;
  d.counters = pl.counters; d.targets = pl.targets; d.n = ph.n_deps;
  const int m_lo = r0 / STEP_ROW_BLOCK, m_hi = (r1 > r0 ? (r1 - 1) / STEP_ROW_BLOCK : m_lo - 1);
#pragma unroll
  for (int i = 0; i < STEP_MAX_DEPS; ++i) {
    if (i < ph.n_deps) {
      const int cb = pl.ph[ph.dep_phase[i]].cbase;
      if (ph.dep_all[i] || m_hi < m_lo) { d.lo[i] = cb + pl.mt; d.hi[i] = cb + pl.mt; }
      else { d.lo[i] = cb + m_lo; d.hi[i] = cb + m_hi; }
    } else { d.lo[i] = 0; d.hi[i] = -1; }
  }
  return d;
}

// Separate functions (not inlined): each GEMM variant keeps its own register allocation for the epilogue hot loop; the
// walker's state is saved around the call once per unit.
template <int MODE, int FEATS>
__device__ __noinline__ void gemm_unit(TileCtx& ctx, const GemmProblem& P, int local, const UnitDeps& deps, const LossTail* tail) {
  gemm_tile<MODE, FEATS, true>(ctx, P, local, deps, tail);
}

// The argument structs live in global memory (the plan image).  Every unit works on a by-value copy: field reads then
// come from registers instead of being re-fetched after each global store (the stores could alias the struct).
__device__ __noinline__ void ew_unit(const StepPhase& ph_s, const void* args, int u, int r0, int r1, int tid, uint32_t ew_parity,
                                   unsigned long long* dbg_row) {
  uint8_t* smem = aligned_smem();
  void* scratch = smem + SCRATCH_OFFSET;
  const int kind = ph_s.kind, sub = ph_s.sub, n_blocks = ph_s.n_blocks, gx = ph_s.gx, rpb = ph_s.rpb;
  const int b0 = u * sub, b1 = min((u + 1) * sub, n_blocks);
  switch (kind) {
    case SK_INGEST: {
      const IngestArgs a = *static_cast<const IngestArgs*>(args);
      // staging area: the GEMM operand ring (idle during element-wise units)
      ingest_body_bulk<true>(a, r0, r1, tid, u == 0, smem, GEMM_STAGES * STAGE_BYTES, tile_ew_bar(smem), ew_parity, dbg_row);
      break;
    }
    case SK_BN_ACT: {
      const BnActArgs a = *static_cast<const BnActArgs*>(args);
      bn_act_body<true>(a, rpb, u % gx, u / gx, tid, scratch);
      break;
    }
    case SK_BN_BWD: {
      const BnBwdArgs a = *static_cast<const BnBwdArgs*>(args);
      bn_bwd_body<true>(a, rpb, u % gx, u / gx, tid, scratch);
      break;
    }
    case SK_LATENT_FWD: {
      const LatentFwdArgs a = *static_cast<const LatentFwdArgs*>(args);
      for (int b = b0; b < b1; ++b) latent_fwd_body<true>(a, b, tid, scratch);
      break;
    }
    case SK_LATENT_BWD: {
      const LatentBwdArgs a = *static_cast<const LatentBwdArgs*>(args);
      for (int b = b0; b < b1; ++b) latent_bwd_body(a, b, tid);
      break;
    }
    case SK_LOSS: {
      const LossArgs a = *static_cast<const LossArgs*>(args);
      for (int b = b0; b < b1; ++b) loss_body<true>(a, b, n_blocks, tid, scratch);
      break;
    }
    default: {   // SK_ADAMW
      const AdamArgs a = *static_cast<const AdamArgs*>(args);
      adamw_unit(a, b0, b1, tid, dbg_row);
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) step_kernel(const StepPlan* __restrict__ plan_g) {
  uint8_t* smem = aligned_smem();
  // plan -> shared memory (phase descriptors are read by every role of every unit)
  {
    const uint4* src = reinterpret_cast<const uint4*>(plan_g);
    uint4* dst = reinterpret_cast<uint4*>(smem + PLAN_OFFSET);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(StepPlan) / 16); i += GEMM_THREADS) dst[i] = src[i];
  }
  TileCtx ctx = tile_setup(true);          // ends with __syncthreads
  const StepPlan& pl = *reinterpret_cast<const StepPlan*>(smem + PLAN_OFFSET);
  const char* base = reinterpret_cast<const char*>(plan_g);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  ctx.dbg = pl.dbg;

  for (int p = 0; p < pl.n_phases; ++p) {
    const StepPhase& ph = pl.ph[p];
    const void* args = base + ph.args_off;
    int u = (cta - ph.unit_rot % G + G) % G;
    for (; u < ph.n_units; u += G) {
      ctx.dbg_row = ph.unit_base + u;
      if (threadIdx.x == 0 && ctx.dbg) {
        ctx.dbg[static_cast<size_t>(ctx.dbg_row) * 8 + 0] = gtime();
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        ctx.dbg[static_cast<size_t>(ctx.dbg_row) * 8 + 7] = (static_cast<unsigned long long>(p) << 32) | smid;
      }
      int r0, r1;
      unit_rows(ph, args, u, &r0, &r1);
      const UnitDeps deps = make_deps(pl, ph, r0, r1);
      if (ph.kind <= SK_GEMM_TN) {
        const GemmGroup* grp = static_cast<const GemmGroup*>(args);
        const GemmProblem& P = *find_problem(grp, u);
        const int local = u - P.tile_begin;
        const LossTail* tail = &grp->tail;
        switch (ph.kind) {
          case SK_GEMM_NT_PLAIN:  gemm_unit<0, FEATS_FWD_PLAIN>(ctx, P, local, deps, tail); break;
          case SK_GEMM_NT_FULL:   gemm_unit<0, FEATS_FWD_FULL>(ctx, P, local, deps, tail); break;
          case SK_GEMM_NT_LOSS:   gemm_unit<0, FEATS_FWD_LOSS>(ctx, P, local, deps, tail); break;
          case SK_GEMM_NN_PLAIN:  gemm_unit<2, FEATS_DGRAD_PLAIN>(ctx, P, local, deps, tail); break;
          case SK_GEMM_NN_FULL:   gemm_unit<2, FEATS_DGRAD_FULL>(ctx, P, local, deps, tail); break;
          default:                gemm_unit<1, FEATS_WGRAD>(ctx, P, local, deps, tail); break;
        }
      } else if (warp >= 2) {
        // element-wise unit: the eight epilogue warps are the 256-thread block
        const int tid = threadIdx.x - 64;
        if (warp == 2) deps.wait(lane);
        ew_sync<true>();
        if (tid == 0 && ctx.dbg) ctx.dbg[static_cast<size_t>(ctx.dbg_row) * 8 + 1] = gtime();
        ew_unit(ph, args, u, r0, r1, tid, ctx.ew_parity, ctx.dbg ? ctx.dbg + static_cast<size_t>(ctx.dbg_row) * 8 : nullptr);
      }
      if (ph.kind == SK_INGEST) ctx.ew_parity ^= 1u;     // every thread tracks the bulk-load barrier's phase
      // ---- unit boundary: order this unit's shared-memory / TMEM use before the next unit, then publish ----
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes (patches) before later TMA writes
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (threadIdx.x == 64) {
        // release: the CTA's stores (ordered before this thread by the barrier) become visible before the counters move
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (r1 > r0)
          for (int m = r0 / STEP_ROW_BLOCK; m <= (r1 - 1) / STEP_ROW_BLOCK; ++m) red_add_u32(pl.counters + ph.cbase + m, 1u);
        red_add_u32(pl.counters + ph.cbase + pl.mt, 1u);
        if (ctx.dbg) ctx.dbg[static_cast<size_t>(ctx.dbg_row) * 8 + 6] = gtime();
      }
    }
  }

  // ---- exit: the last CTA re-arms the counters for the next launch ----
  __shared__ int s_last;
  if (threadIdx.x == 64) {
    __threadfence();
    const unsigned int t = atomicAdd(pl.finish, 1u);
    s_last = (t == static_cast<unsigned int>(G) - 1u) ? 1 : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(ctx.tmem_base, GEMM_TMEM_COLS);
  }
  if (s_last) {
    __threadfence();
    for (int i = threadIdx.x; i < pl.n_counters; i += GEMM_THREADS) pl.counters[i] = 0u;
    if (threadIdx.x == 0) *pl.finish = 0u;
  }
}

}  // namespace

size_t step_smem_bytes() { return STEP_SMEM_BYTES; }

void step_unit_rows_host(const StepPhase& ph, const void* args, int u, int* r0, int* r1) { unit_rows(ph, args, u, r0, r1); }

LossGridInfo loss_grid_info(const LossArgs& a) {
  const LossGrid g = loss_grid(a);
  return LossGridInfo{g.nb_a, g.nb_b, g.nb_c, g.nb_k};
}

int step_max_grid(cudaError_t* err) {
  cudaError_t e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM_BYTES);
  int per_sm = 0, dev = 0, sms = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel, GEMM_THREADS, STEP_SMEM_BYTES);
  if (e == cudaSuccess) e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (err) *err = e;
  if (e != cudaSuccess || per_sm < 1) return 0;
  return sms;            // one CTA per SM (TMEM and shared memory are sized for exactly that)
}

cudaError_t launch_step(const StepPlan* plan_dev, int grid, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = STEP_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, step_kernel, plan_dev);
}

}  // namespace vla
