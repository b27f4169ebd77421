// Chain kernel: the row-local stretches of the train step (and the whole eval-mode inference forward) as ONE launch.
//
// A stretch is a list of phases -- exactly the launches of the per-call path, same device bodies (gemm_tile.cuh,
// elementwise_dev.cuh) -- in which a 128-row block of the batch depends only on the same rows of the previous phase:
// everything between two BatchNorm-statistics boundaries.  One thread-block cluster of four CTAs owns a row block and
// walks the phases; a GEMM phase's tiles (the layer's N split over the cluster, host-assigned by cost) or an
// element-wise phase's rows (32 per CTA) are spread over the four CTAs, and a cluster barrier (release / acquire) replaces
// the kernel boundary.  Activations move through L2 -- the weight-gradient GEMMs need them in global memory anyway --
// while TMEM, the mbarrier ring and the tensor-map prefetches are set up once per launch.  Clusters are independent of
// one another (no inter-cluster waits: no co-residency requirement); with more row blocks than clusters a cluster loops.
//
// Replaces, for rna2dna at batch 4096, eleven launches (BN apply -> heads -> latent -> three decoder layers + loss ->
// three decoder data gradients -> latent backward -> encoder data gradient: the loop body of the reference at
// train_rna2dna.py:86-95 between the two BatchNorm reductions) by one, and ingest -> first encoder layer by another.
#include "elementwise_dev.cuh"
#include "gemm_tile.cuh"

namespace vla {

namespace {

// ---- shared-memory phase images ----
// Per-phase latency is what the chain pays eleven times per row block, so nothing on a phase's start-up path may be a
// dependent global load: the phase table (this rank's view) is copied into shared memory once per launch, and the
// argument structs of phase p + 1 -- the scalar part of every GEMM problem this CTA has a tile of, the loss tail, or the
// element-wise argument struct -- are staged by asynchronous copies while phase p runs (double-buffered).
struct PhaseRow {                      // 40 bytes
  int kind;
  int n_units;                         // GEMM phases: tiles of this cluster rank
  long long args_off;
  unsigned short code[CHAIN_MAX_UNITS];
};
constexpr int EW_ARG_BYTES = 176;
static_assert(sizeof(IngestArgs) <= EW_ARG_BYTES && sizeof(BnActArgs) <= EW_ARG_BYTES && sizeof(BnBwdArgs) <= EW_ARG_BYTES &&
              sizeof(LatentFwdArgs) <= EW_ARG_BYTES && sizeof(LatentBwdArgs) <= EW_ARG_BYTES, "element-wise argument staging");
struct alignas(16) PhaseImg {
  LossTail tail;                                            // 80
  const CUtensorMap* tm[CHAIN_MAX_UNITS][2];                // 192
  union {
    GemmScalars unit[CHAIN_MAX_UNITS];                      // 12 x 224
    char ew[EW_ARG_BYTES];
  };
};
constexpr size_t GROUP_P_OFFSET = sizeof(GemmGroup) - GEMM_MAX_PROBLEMS * sizeof(GemmProblem);
constexpr size_t GROUP_TAIL_OFFSET = 24;
constexpr size_t PROBLEM_TMA_OFFSET = sizeof(GemmProblem) - 2 * sizeof(CUtensorMap);
static_assert(sizeof(LossTail) % 16 == 0 && sizeof(PhaseRow) == 40, "staging copies are 16-byte pieces");

constexpr int CH_TABLE_OFFSET = SMEM_USED;
constexpr int CH_IMG_OFFSET = CH_TABLE_OFFSET + ((CHAIN_MAX_PHASES * static_cast<int>(sizeof(PhaseRow)) + 15) & ~15);
constexpr int CH_SCRATCH_OFFSET = 0;       // element-wise scratch aliases the first operand slot (no GEMM tile runs in an element-wise phase)
constexpr int CH_SMEM_USED = CH_IMG_OFFSET + 2 * static_cast<int>(sizeof(PhaseImg));
static_assert(EW_SCRATCH_BYTES <= STAGE_BYTES, "element-wise scratch aliases one operand slot");
constexpr int CH_SMEM_BYTES = CH_SMEM_USED + 1024;
static_assert(CH_SMEM_BYTES <= 227 * 1024, "chain kernel shared memory budget");
static_assert(CH_IMG_OFFSET % 16 == 0 && sizeof(PhaseImg) % 16 == 0, "phase images are filled by 16-byte asynchronous copies");

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// Phase boundary: every thread of the four CTAs.  The cluster barrier's release / acquire orders this cluster's global
// stores (generic proxy) before the next phase's loads; the proxy fences on both sides extend that to the next phase's
// TMA reads (async proxy).  __syncthreads first: the roles of a CTA reconverge before the aligned cluster barrier.
__device__ __forceinline__ void phase_barrier() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Between two tiles of one CTA inside a phase: shared-memory patches (generic proxy) before the next tile's TMA writes,
// TMEM reads before the next tile's MMAs.
__device__ __forceinline__ void tile_boundary() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// Separate functions (not inlined): each GEMM variant keeps its own register allocation for the epilogue hot loop.
template <int MODE, int FEATS>
__device__ __noinline__ void gemm_unit(TileCtx& ctx, const GemmScalars& S, const CUtensorMap* tmA, const CUtensorMap* tmB,
                                       int m_tile, int n_tile, const LossTail* tail) {
  gemm_tile<MODE, FEATS>(ctx, S, tmA, tmB, m_tile, n_tile, 0, tail);
}

// The 256 element-wise threads (et = 0..255) start the copies of phase `p`'s arguments into `img` (they complete
// asynchronously; the caller commits the group and waits for it in front of the phase barrier).
__device__ __forceinline__ void stage_phase(const PhaseRow* table, int p, PhaseImg* img, const char* base, int et) {
  const int kind = table[p].kind, nu = table[p].n_units;
  const char* args = base + table[p].args_off;
  if (kind <= CK_GEMM_LAST) {
    constexpr int PIECES = static_cast<int>(sizeof(GemmScalars)) / 16;
    for (int i = et; i < nu * PIECES; i += EW_THREADS) {
      const int u = i / PIECES, c = i - u * PIECES;
      const char* prob = args + GROUP_P_OFFSET + static_cast<size_t>(table[p].code[u] >> 8) * sizeof(GemmProblem);
      cp_async16(reinterpret_cast<char*>(&img->unit[u]) + c * 16, prob + c * 16);
    }
    if (et < nu) {
      const char* prob = args + GROUP_P_OFFSET + static_cast<size_t>(table[p].code[et] >> 8) * sizeof(GemmProblem);
      img->tm[et][0] = reinterpret_cast<const CUtensorMap*>(prob + PROBLEM_TMA_OFFSET);
      img->tm[et][1] = reinterpret_cast<const CUtensorMap*>(prob + PROBLEM_TMA_OFFSET + sizeof(CUtensorMap));
    }
    constexpr int TP = static_cast<int>(sizeof(LossTail)) / 8;      // (the tail sits at an 8-byte offset of the group)
    if (et >= EW_THREADS - TP) {
      const int i = et - (EW_THREADS - TP);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(reinterpret_cast<char*>(&img->tail) + i * 8)),
                   "l"(args + GROUP_TAIL_OFFSET + i * 8) : "memory");
    }
  } else if (et < EW_ARG_BYTES / 16) {
    cp_async16(img->ew + et * 16, args + et * 16);       // (the plan image is padded behind its last section)
  }
}

// Element-wise phase on rows [r0, r1) by the 256 threads of warps 2..9; args: the staged argument struct (shared memory).
__device__ __noinline__ void ew_phase(int kind, const void* args, int mb, int rank, int r0, int r1, int tid) {
  uint8_t* smem = aligned_smem();
  void* scratch = smem + CH_SCRATCH_OFFSET;
  const int by = mb * CHAIN_CLUSTER + rank;                  // index of this CTA's 32-row slice
  constexpr int RPB = CHAIN_ROWS / CHAIN_CLUSTER;
  switch (kind) {
    case CK_INGEST: {
      const IngestArgs a = *static_cast<const IngestArgs*>(args);
      ingest_body(a, r0, r1, tid >> 5, EW_THREADS / 32, tid & 31, mb == 0 && rank == 0 && tid == 0);
      break;
    }
    case CK_BN_ACT: {
      const BnActArgs a = *static_cast<const BnActArgs*>(args);
      for (int bx = 0; bx * BN_COLS < a.n; ++bx) bn_act_body<true>(a, RPB, bx, by, tid, scratch);
      break;
    }
    case CK_BN_BWD: {
      const BnBwdArgs a = *static_cast<const BnBwdArgs*>(args);
      for (int bx = 0; bx * BN_COLS < a.n; ++bx) { bn_bwd_body<true>(a, RPB, bx, by, tid, scratch); ew_sync<true>(); }
      break;
    }
    case CK_LATENT_FWD: {
      const LatentFwdArgs a = *static_cast<const LatentFwdArgs*>(args);
      latent_fwd_rows<true>(a, r0, r1, by, tid, scratch);
      break;
    }
    default: {   // CK_LATENT_BWD
      const LatentBwdArgs a = *static_cast<const LatentBwdArgs*>(args);
      latent_bwd_rows(a, r0, r1, tid);
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) chain_kernel(const ChainPlan* __restrict__ plan) {
  const int warp = threadIdx.x >> 5;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cid = static_cast<int>(cluster_id_x());
  const int n_clusters = static_cast<int>(cluster_count_x());
  const char* base = reinterpret_cast<const char*>(plan);
  uint8_t* smem = aligned_smem();
  PhaseRow* table = reinterpret_cast<PhaseRow*>(smem + CH_TABLE_OFFSET);
  PhaseImg* imgs = reinterpret_cast<PhaseImg*>(smem + CH_IMG_OFFSET);
  const int et = static_cast<int>(threadIdx.x) - 64;
  // ---- one-time setup (overlaps the previous kernel's tail; the plan image is written by the host before the step) ----
  const int n_phases = plan->n_phases, m_blocks = plan->m_blocks, rows = plan->rows;
  unsigned long long* dbg = plan->dbg;
  for (int p = static_cast<int>(threadIdx.x); p < n_phases; p += GEMM_THREADS) {
    const ChainPhase& ph = plan->ph[p];
    PhaseRow r;
    r.kind = ph.kind; r.n_units = ph.kind <= CK_GEMM_LAST ? ph.n_units[rank] : 0; r.args_off = ph.args_off;
#pragma unroll
    for (int u = 0; u < CHAIN_MAX_UNITS; ++u) r.code[u] = ph.units[rank][u];
    table[p] = r;
  }
  if (warp == 0) {
    // tensor maps of every tile this CTA will run: fetch the descriptors now
    for (int p = 0; p < n_phases; ++p) {
      const ChainPhase& ph = plan->ph[p];
      if (ph.kind > CK_GEMM_LAST) continue;
      const GemmGroup* grp = reinterpret_cast<const GemmGroup*>(base + ph.args_off);
      for (int u = static_cast<int>(threadIdx.x); u < ph.n_units[rank]; u += 32) {
        const GemmProblem& P = grp->p[ph.units[rank][u] >> 8];
        tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmB);
      }
    }
  }
  TileCtx ctx = tile_setup(true);          // barriers, TMEM, bf16 ones; ends with __syncthreads (the phase table is complete)
  if (warp >= 2) { stage_phase(table, 0, imgs, base, et); cp_async_commit(); cp_async_wait<0>(); }
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();

  int buf = 0;
  for (int mb = cid; mb < m_blocks; mb += n_clusters) {
    const int r0 = min(rows, mb * CHAIN_ROWS + rank * (CHAIN_ROWS / CHAIN_CLUSTER));
    const int r1 = min(rows, r0 + CHAIN_ROWS / CHAIN_CLUSTER);
    for (int p = 0; p < n_phases; ++p) {
      const int kind = table[p].kind;
      const int nu = table[p].n_units;
      PhaseImg* img = imgs + buf;
      const size_t drow = static_cast<size_t>(cid * CHAIN_CLUSTER + rank) * CHAIN_MAX_PHASES + p;
      ctx.dbg = dbg; ctx.dbg_row = static_cast<int>(drow);
      if (dbg && threadIdx.x == 0) dbg[drow * 8] = gtime();
      // the next phase's arguments travel while this one runs
      if (warp >= 2) {
        const bool more = p + 1 < n_phases || mb + n_clusters < m_blocks;
        if (more) stage_phase(table, p + 1 < n_phases ? p + 1 : 0, imgs + (buf ^ 1), base, et);
        cp_async_commit();
      }
      if (kind <= CK_GEMM_LAST) {
        for (int u = 0; u < nu; ++u) {
          const GemmScalars& S = img->unit[u];
          if (mb < S.m_tiles) {
            const int n_tile = table[p].code[u] & 255;
            const CUtensorMap* tmA = img->tm[u][0];
            const CUtensorMap* tmB = img->tm[u][1];
            switch (kind) {
              case CK_GEMM_NT_PLAIN:    gemm_unit<0, FEATS_FWD_PLAIN>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
              case CK_GEMM_NT_FULL:     gemm_unit<0, FEATS_FWD_FULL>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
              case CK_GEMM_NT_LOSS:     gemm_unit<0, FEATS_FWD_LOSS>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
              case CK_GEMM_NT_LOSS_BCE: gemm_unit<0, FEATS_FWD_LOSS_BCE>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
              case CK_GEMM_NT_LOSS_MSE: gemm_unit<0, FEATS_FWD_LOSS_MSE>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
              case CK_GEMM_NN_PLAIN:    gemm_unit<2, FEATS_DGRAD_PLAIN>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
              default:                  gemm_unit<2, FEATS_DGRAD_FULL>(ctx, S, tmA, tmB, mb, n_tile, &img->tail); break;
            }
            if (u + 1 < nu) tile_boundary();
          }
        }
      } else if (warp >= 2 && r1 > r0) {
        if (dbg && threadIdx.x == 64) dbg[drow * 8 + 5] = gtime();
        ew_phase(kind, img->ew, mb, rank, r0, r1, et);
        if (dbg && threadIdx.x == 64) dbg[drow * 8 + 6] = gtime();
        if (dbg && threadIdx.x == GEMM_THREADS - 32) dbg[drow * 8 + 7] = gtime();
      }
      if (dbg && threadIdx.x == 0) dbg[drow * 8 + 1] = gtime();
      if (warp >= 2) cp_async_wait<0>();
      phase_barrier();
      if (dbg && threadIdx.x == 0) dbg[drow * 8 + 2] = gtime();
      buf ^= 1;
    }
  }

  // ---- teardown ----
  if (warp == 1) tmem_dealloc(ctx.tmem_base, GEMM_TMEM_COLS);
}

}  // namespace

size_t chain_smem_bytes() { return CH_SMEM_BYTES; }

static cudaError_t chain_prepare() {
  static cudaError_t attr = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
  return attr;
}

// How many 4-CTA clusters of this kernel the device runs at once (the grid never needs to exceed that: clusters loop).
int chain_max_clusters(cudaError_t* err) {
  cudaError_t e = chain_prepare();
  int n = 0;
  if (e == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CHAIN_CLUSTER * 64); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = CH_SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CHAIN_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaOccupancyMaxActiveClusters(&n, chain_kernel, &cfg);
  }
  if (err) *err = e;
  return e == cudaSuccess ? n : 0;
}

cudaError_t launch_chain(const ChainPlan* plan_dev, int n_clusters, cudaStream_t s) {
  cudaError_t e = chain_prepare();
  if (e != cudaSuccess) return e;
  static const bool pdl_on = [] { const char* v = getenv("VLA_NO_PDL"); return !(v && v[0] == '1'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_clusters * CHAIN_CLUSTER); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = CH_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CHAIN_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_on ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, chain_kernel, plan_dev);
}

}  // namespace vla
