// Grouped bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA into 128-byte-swizzled shared memory, one 128 x BN output tile per CTA.
//
//   mode 0 NT: C[M,N] = A[M,K] * B[N,K]^T   both operands K-major       (Linear forward)
//   mode 1 TN: C[M,N] = A[K,M]^T * B[K,N]   both operands MN-major      (weight gradients, split-K + red.add)
//   mode 2 NN: C[M,N] = A[M,K] * B[K,N]     A K-major, B MN-major       (data gradients: B = the forward weight copy)
//
// One launch covers up to GEMM_MAX_PROBLEMS independent problems (e.g. both encoders' first layers,
// or every weight gradient of the model): CTA index -> (problem, tile) through GemmGroup.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue.
// Epilogue (8 warps, thread = accumulator row): TMEM -> registers, bias / mask / activation applied in registers with
// the global operands prefetched as 128-bit vectors, per-column statistics (BatchNorm forward and backward) by a
// shuffle butterfly, 128-bit row stores; split-K partials go through a smem transpose so every red.add is coalesced.
//
// Replaces, per layer, the ATen addmm / mm calls issued by nn.Linear in the reference
// (src/models/encoders.py:13-19,31-41,54-55; src/models/decoders.py:13-15,27-31,44-46) and their
// autograd backward.
#include "gemm_tile.cuh"

#include <algorithm>
#include <mutex>

namespace vla {

namespace {

template <int MODE, int FEATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmGroup grp) {
  if (threadIdx.x == 0 && grp.dbg) {
    grp.dbg[static_cast<size_t>(blockIdx.x) * 8 + 0] = gtime();                                  // kernel entry
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); grp.dbg[static_cast<size_t>(blockIdx.x) * 8 + 7] = smid;
  }
  // ---- which problem / tile ----
  int pi = 0;
#pragma unroll
  for (int i = 1; i < GEMM_MAX_PROBLEMS; ++i)
    if (i < grp.nprob && static_cast<int>(blockIdx.x) >= grp.p[i].tile_begin) pi = i;
  const GemmProblem& P = grp.p[pi];
  if (threadIdx.x == 0) { tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmB); }
  // ---- one-time setup (overlaps the previous kernel's tail) ----
  TileCtx ctx = tile_setup(!(grp.dbg_flags & 4), !(grp.dbg_flags & 2));
  ctx.dbg = grp.dbg; ctx.dbg_row = blockIdx.x; ctx.dbg_flags = grp.dbg_flags;
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0 && grp.dbg) grp.dbg[static_cast<size_t>(blockIdx.x) * 8 + 2] = gtime();  // setup done, dependency resolved
  {
    const int local = static_cast<int>(blockIdx.x) - P.tile_begin;
    const int n_tile = local % P.n_tiles, rest = local / P.n_tiles;
    gemm_tile<MODE, FEATS>(ctx, P, &P.tmA, &P.tmB, rest % P.m_tiles, n_tile, rest / P.m_tiles, &grp.tail);
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && grp.dbg) grp.dbg[static_cast<size_t>(blockIdx.x) * 8 + 6] = gtime();  // epilogue done
  if ((threadIdx.x >> 5) == 1 && !(grp.dbg_flags & 4)) {
    tc_fence_after();
    tmem_dealloc(ctx.tmem_base, GEMM_TMEM_COLS);
  }
}

// Lock-step population step: the same tile body, one launch for the groups of n members (tables in global memory).
template <int MODE, int FEATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_multi_kernel(const MultiHdr* __restrict__ hdr, const GemmGroup* __restrict__ groups, int n) {
  __shared__ int s_sel[3];
  const int bid = static_cast<int>(blockIdx.x);
  for (int t = threadIdx.x; t < n; t += GEMM_THREADS) {
    const int b = hdr[t].block_begin;
    const int e = t + 1 < n ? hdr[t + 1].block_begin : 0x7fffffff;
    if (bid >= b && bid < e) { s_sel[0] = t; s_sel[1] = bid - b; }
  }
  __syncthreads();
  const GemmGroup& grp = groups[s_sel[0]];
  const int tile = s_sel[1];
  if (threadIdx.x < GEMM_MAX_PROBLEMS) {
    const int i = threadIdx.x;
    if (i < grp.nprob) {
      const int b = grp.p[i].tile_begin;
      const int e = i + 1 < grp.nprob ? grp.p[i + 1].tile_begin : 0x7fffffff;
      if (tile >= b && tile < e) s_sel[2] = i;
    }
  }
  __syncthreads();
  const GemmProblem& P = grp.p[s_sel[2]];
  if (threadIdx.x == 0) { tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmB); }
  TileCtx ctx = tile_setup(true, true);
  pdl_wait();
  pdl_launch_dependents();
  {
    const int local = tile - P.tile_begin;
    const int n_tile = local % P.n_tiles, rest = local / P.n_tiles;
    gemm_tile<MODE, FEATS>(ctx, P, &P.tmA, &P.tmB, rest % P.m_tiles, n_tile, rest / P.m_tiles, &grp.tail);
  }
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tmem_dealloc(ctx.tmem_base, GEMM_TMEM_COLS);
  }
}

// Persistent form for launches of more than one wave (large batches, lock-step populations): one CTA per SM walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...; two accumulators in TMEM let the MMAs of tile i + 1 overlap the epilogue of tile i,
// and the producer streams the next tile's operands meanwhile (gemm_tile<.., PERSIST = true>).  n == 0: the tiles of `one`
// (kernel parameter); n > 0: the tiles of n members' groups (tables in global memory).
template <int MODE, int FEATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_persist_kernel(const __grid_constant__ GemmGroup one, const MultiHdr* __restrict__ hdr,
                                                                          const GemmGroup* __restrict__ groups, int n, int total_tiles) {
  TileCtx ctx = tile_setup(true, true, 2 * GEMM_TMEM_COLS);
  pdl_wait();
  pdl_launch_dependents();
  // A CTA takes a CONTIGUOUS range of tiles (n-tiles of one row block are neighbours: the row block's A operand is read from
  // HBM once and re-read from L2 by the same SM); the (member, problem) lookup is repeated only when the range leaves a problem.
  const int per = (total_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int t0 = static_cast<int>(blockIdx.x) * per, t1 = min(total_tiles, t0 + per);
  const GemmGroup* grp = &one;
  const GemmProblem* Pp = nullptr;
  int p_begin = 0, p_end = 0, p_nt = 1, p_mt = 1;       // tiles [p_begin, p_end) of the launch belong to *Pp
  for (int t = t0; t < t1; ++t) {
    if (t >= p_end) {
      int base = 0;
      if (n > 0) {
        int lo = 0, hi = n - 1;                    // last member whose block_begin <= t
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (hdr[mid].block_begin <= t) lo = mid; else hi = mid - 1;
        }
        grp = groups + lo;
        base = hdr[lo].block_begin;
      }
      const int np = grp->nprob;
      int pi = 0;
      for (int i = 1; i < np; ++i)
        if (t - base >= grp->p[i].tile_begin) pi = i;
      Pp = &grp->p[pi];
      p_nt = Pp->n_tiles; p_mt = Pp->m_tiles;
      p_begin = base + Pp->tile_begin;
      p_end = p_begin + p_mt * p_nt * Pp->k_splits;
    }
    const int local = t - p_begin;
    const int n_tile = local % p_nt, rest = local / p_nt;
    gemm_tile<MODE, FEATS, true>(ctx, *Pp, &Pp->tmA, &Pp->tmB, rest % p_mt, n_tile, rest / p_mt, &grp->tail);
  }
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tmem_dealloc(ctx.tmem_base, 2 * GEMM_TMEM_COLS);
  }
}

// NT launches whose problems all have a multiple of GEMM_CLUSTER n-tiles per row block: 4-CTA clusters, the A tile of a row
// block multicast across the cluster (gemm_tile<.., CL>).
constexpr int GEMM_CLUSTER = 4;
template <int FEATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_cluster_kernel(const __grid_constant__ GemmGroup grp) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < GEMM_MAX_PROBLEMS; ++i)
    if (i < grp.nprob && static_cast<int>(blockIdx.x) >= grp.p[i].tile_begin) pi = i;
  const GemmProblem& P = grp.p[pi];
  if (threadIdx.x == 0) { tma_prefetch_desc(&P.tmA); tma_prefetch_desc(&P.tmB); }
  TileCtx ctx = tile_setup(true, false, GEMM_TMEM_COLS, GEMM_CLUSTER);
  cluster_sync_all();              // every CTA's barriers exist before a peer's multicast can signal them
  pdl_wait();
  pdl_launch_dependents();
  {
    const int local = static_cast<int>(blockIdx.x) - P.tile_begin;
    const int n_tile = local % P.n_tiles, rest = local / P.n_tiles;
    gemm_tile<0, FEATS, false, GEMM_CLUSTER>(ctx, P, &P.tmA, &P.tmB, rest % P.m_tiles, n_tile, rest / P.m_tiles, &grp.tail);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA leaves while a peer may still write into its slots or arrive on its barriers
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tmem_dealloc(ctx.tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace

size_t gemm_smem_bytes() { return SMEM_BYTES; }

template <int MODE, int FEATS>
cudaError_t launch_one(const GemmGroup& g, cudaStream_t stream, size_t smem) {
  static cudaError_t attr = cudaFuncSetAttribute(gemm_tc_kernel<MODE, FEATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr != cudaSuccess) return attr;
  return launch_pdl(gemm_tc_kernel<MODE, FEATS>, dim3(g.total_tiles), dim3(GEMM_THREADS), smem, stream, g);
}

template <int MODE, int FEATS>
cudaError_t launch_one_multi(const MultiHdr* hdr, const GemmGroup* groups, int n, int total_blocks, cudaStream_t stream) {
  static cudaError_t attr = cudaFuncSetAttribute(gemm_tc_multi_kernel<MODE, FEATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr != cudaSuccess) return attr;
  return launch_pdl(gemm_tc_multi_kernel<MODE, FEATS>, dim3(total_blocks), dim3(GEMM_THREADS), static_cast<size_t>(SMEM_BYTES), stream,
                    hdr, groups, n);
}

template <int MODE, int FEATS>
cudaError_t launch_one_persist(const GemmGroup& one, const MultiHdr* hdr, const GemmGroup* groups, int n, int total_tiles, cudaStream_t stream) {
  static cudaError_t attr = cudaFuncSetAttribute(gemm_tc_persist_kernel<MODE, FEATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr != cudaSuccess) return attr;
  static const int sms = [] { int d = 0, v = 148; if (cudaGetDevice(&d) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v > 0 ? v : 148; }();
  return launch_pdl(gemm_tc_persist_kernel<MODE, FEATS>, dim3(total_tiles < sms ? total_tiles : sms), dim3(GEMM_THREADS),
                    static_cast<size_t>(SMEM_BYTES), stream, one, hdr, groups, n, total_tiles);
}
// More than this many tiles (one wave and a half) and no fused loss: the persistent form.  VLA_PERSIST=0 turns it off.
bool use_persist(int variant, int total_tiles) {
  static const int thr = [] { const char* e = getenv("VLA_PERSIST"); return e ? atoi(e) : 222; }();
  // not the loss epilogues (their target patches live in the ring), not the data gradients with BatchNorm-backward statistics
  // (measured 1.3-1.4x slower in the persistent form: profiles/r2_population_profile.md)
  return thr > 0 && total_tiles > thr && !(variant >= 2 && variant <= 4) && variant != 33 && variant != 34;
}
// The persistent form runs a two-stage ring (its transpose patches need the other slots): it pays where the epilogue is the
// larger part of a tile -- short main loops -- and loses where the main loop is (measured on the merged launches of 40 tri-modal
// models at batch 4096, profiles/r2_population_profile.md: 9 / 13 slots 1.7x / 1.2x faster, 16 slots equal, 26 slots (split
// operands, K = 782) 1.2x slower, the K = 4096 weight-gradient tiles 1.8x slower).  Longest main loop of a group, in ring slots:
int gemm_max_units(const GemmGroup& g) {
  int u = 0;
  for (int i = 0; i < g.nprob; ++i) {
    const GemmProblem& p = g.p[i];
    const int kb = p.k_splits > 1 ? p.kb_per_split : (p.K + GEMM_BK - 1) / GEMM_BK;
    u = std::max(u, kb * (p.a_lo > 0 ? 2 : 1));
  }
  return u;
}
static const int PERSIST_MAX_UNITS = [] { const char* e = getenv("VLA_PERSIST_UNITS"); return e ? atoi(e) : 14; }();

template <int FEATS>
cudaError_t launch_one_cluster(const GemmGroup& g, cudaStream_t stream) {
  static cudaError_t attr = cudaFuncSetAttribute(gemm_tc_cluster_kernel<FEATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr != cudaSuccess) return attr;
  static const bool pdl_on = [] { const char* e = getenv("VLA_NO_PDL"); return !(e && e[0] == '1'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g.total_tiles); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = GEMM_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_on ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc_cluster_kernel<FEATS>, g);
}
// How many 4-CTA clusters of the cluster kernel the device runs at once (0: clusters unavailable).
int gemm_cluster_capacity() {
  static const int cap = [] {
    if (cudaFuncSetAttribute(gemm_tc_cluster_kernel<FEATS_FWD_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(GEMM_CLUSTER * 64); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = GEMM_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_cluster_kernel<FEATS_FWD_PLAIN>, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
  }();
  return cap;
}

// Which instantiation of the tile body a group runs: mode * 16 + {0 plain, 1 full, 2 loss (any mix), 3 BCE only, 4 MSE only}.
int gemm_variant(const GemmGroup& g, int mode) {
  int used = 0;
  for (int i = 0; i < g.nprob; ++i) used |= g.p[i].flags;
  if (mode == 0) {
    if (!(used & ~FEATS_FWD_PLAIN)) return 0;
    if (used & GF_LOSS) {
      if (used & ~FEATS_FWD_LOSS) return -1;
      int kinds = 0;
      for (int i = 0; i < g.nprob; ++i) kinds |= (g.p[i].flags & GF_LOSS) ? 1 << g.p[i].loss_kind : 1 << LOSS_NONE;
      // uniform groups run the epilogue that contains only their loss kind (about half the code of the generic one)
      if (kinds == 1 << LOSS_BCE && !(used & ~FEATS_FWD_LOSS_BCE)) return 3;
      if (kinds == 1 << LOSS_MSE && !(used & ~FEATS_FWD_LOSS_MSE)) return 4;
      return 2;
    }
    return 1;
  }
  if (mode == 2) {
    if (used & GF_LATBWD) {                // the latent-backward epilogue has its own (small) instantiation: uniform groups only
      for (int i = 0; i < g.nprob; ++i) if (g.p[i].flags != GF_LATBWD) return -1;
      return 34;
    }
    return 32 + ((used & ~FEATS_DGRAD_PLAIN) ? 1 : 0);
  }
  return 16;
}

#define VLA_GEMM_DISPATCH(variant, CALL)                                    \
  switch (variant) {                                                        \
    case 0: return CALL(0, FEATS_FWD_PLAIN);                                \
    case 1: return CALL(0, FEATS_FWD_FULL);                                 \
    case 2: return CALL(0, FEATS_FWD_LOSS);                                 \
    case 3: return CALL(0, FEATS_FWD_LOSS_BCE);                             \
    case 4: return CALL(0, FEATS_FWD_LOSS_MSE);                             \
    case 16: return CALL(1, FEATS_WGRAD);                                   \
    case 32: return CALL(2, FEATS_DGRAD_PLAIN);                             \
    case 33: return CALL(2, FEATS_DGRAD_FULL);                              \
    case 34: return CALL(2, FEATS_DGRAD_LAT);                               \
    default: return cudaErrorInvalidValue;                                  \
  }

cudaError_t launch_gemm_group(const GemmGroup& g, int mode, cudaStream_t stream) {
  if (g.total_tiles <= 0) return cudaSuccess;
  const int variant = gemm_variant(g, mode);
  if (variant < 0) return cudaErrorInvalidValue;
  if (Recorder* r = recorder()) {          // lock-step population step: collect, do not launch
    if (g.dbg || g.dbg_flags) { r->unsupported = true; return cudaErrorNotSupported; }
    RecOp op;
    op.kind = RK_GEMM; op.variant = variant; op.blocks = g.total_tiles; op.gx = g.total_tiles;
    op.args.assign(reinterpret_cast<const char*>(&g), sizeof(GemmGroup));
    r->ops.push_back(std::move(op));
    return cudaSuccess;
  }
  if (g.pad[0] == GEMM_CLUSTER && (mode != 0 || g.dbg || g.dbg_flags)) return cudaErrorInvalidValue;   // A maps have 32-row boxes
  if (mode == 0 && g.pad[0] == GEMM_CLUSTER) {      // (finalize_group built the A maps for the multicast ring)
    switch (variant) {
      case 0: return launch_one_cluster<FEATS_FWD_PLAIN>(g, stream);
      case 1: return launch_one_cluster<FEATS_FWD_FULL>(g, stream);
      case 2: return launch_one_cluster<FEATS_FWD_LOSS>(g, stream);
      case 3: return launch_one_cluster<FEATS_FWD_LOSS_BCE>(g, stream);
      case 4: return launch_one_cluster<FEATS_FWD_LOSS_MSE>(g, stream);
      default: return cudaErrorInvalidValue;
    }
  }
  if (!g.dbg && !g.dbg_flags && use_persist(variant, g.total_tiles) && gemm_max_units(g) <= PERSIST_MAX_UNITS) {
#define VLA_CALL_PERSIST(M_, F_) launch_one_persist<M_, F_>(g, nullptr, nullptr, 0, g.total_tiles, stream)
    switch (variant) {
      case 0: return VLA_CALL_PERSIST(0, FEATS_FWD_PLAIN);
      case 1: return VLA_CALL_PERSIST(0, FEATS_FWD_FULL);
      case 16: return VLA_CALL_PERSIST(1, FEATS_WGRAD);
      case 32: return VLA_CALL_PERSIST(2, FEATS_DGRAD_PLAIN);
      case 33: return VLA_CALL_PERSIST(2, FEATS_DGRAD_FULL);
      default: return cudaErrorInvalidValue;
    }
#undef VLA_CALL_PERSIST
  }
  size_t smem = SMEM_BYTES;
  if (g.dbg_flags & 0xFFFF00) smem = static_cast<size_t>(g.dbg_flags >> 8);   // test hook (only valid with dbg_flags & 2)
#define VLA_CALL_ONE(M_, F_) launch_one<M_, F_>(g, stream, smem)
  VLA_GEMM_DISPATCH(variant, VLA_CALL_ONE)
#undef VLA_CALL_ONE
}

cudaError_t launch_gemm_multi(int variant, const MultiHdr* hdr, const GemmGroup* groups, int n, int total_blocks, int max_units,
                              cudaStream_t stream) {
  if (n <= 0 || total_blocks <= 0) return cudaSuccess;
  if (n > MULTI_MAX_MEMBERS) return cudaErrorInvalidValue;
  if (use_persist(variant, total_blocks) && max_units <= PERSIST_MAX_UNITS) {
    static GemmGroup none;      // unused kernel parameter of the multi form
#define VLA_CALL_PERSIST(M_, F_) launch_one_persist<M_, F_>(none, hdr, groups, n, total_blocks, stream)
    switch (variant) {
      case 0: return VLA_CALL_PERSIST(0, FEATS_FWD_PLAIN);
      case 1: return VLA_CALL_PERSIST(0, FEATS_FWD_FULL);
      case 16: return VLA_CALL_PERSIST(1, FEATS_WGRAD);
      case 32: return VLA_CALL_PERSIST(2, FEATS_DGRAD_PLAIN);
      case 33: return VLA_CALL_PERSIST(2, FEATS_DGRAD_FULL);
      default: return cudaErrorInvalidValue;
    }
#undef VLA_CALL_PERSIST
  }
#define VLA_CALL_MULTI(M_, F_) launch_one_multi<M_, F_>(hdr, groups, n, total_blocks, stream)
  VLA_GEMM_DISPATCH(variant, VLA_CALL_MULTI)
#undef VLA_CALL_MULTI
}

// ---------------------------------------------------------------------------------------------
// Tensor maps (driver entry point fetched at run time: the library links no libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  static std::string why;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      why = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
      (void)cudaGetLastError();
    } else {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  if (!fn && err) *err = why;
  return fn;
}

bool make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                    uint32_t box_inner, uint32_t box_outer, std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_bytes & 15) || box_inner != 64 || box_outer > 256 ||
      box_outer == 0 || inner == 0 || outer == 0) {
    if (err) *err = "make_tmap_bf16: unsupported geometry (base/pitch alignment or box)";
    return false;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

bool make_tmap_f32_tile(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_bytes & 15) || inner == 0 || outer == 0) {
    if (err) *err = "make_tmap_f32_tile: base and row pitch must be 16-byte aligned";
    return false;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled (fp32 tile) failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

}  // namespace vla
