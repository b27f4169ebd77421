// Stand-alone launches of the element-wise / reduction steps (bodies in elementwise_dev.cuh): used by the per-call
// autograd path (vla_forward / vla_backward / vla_loss / vla_adamw) and for the phases with a grid-wide dependency; the
// row-local phases of large batches run the same bodies inside the chain kernel (chain_kernel.cu).
#include "elementwise_dev.cuh"

#include <algorithm>

namespace vla {

namespace {

__global__ void __launch_bounds__(256) ingest_kernel(IngestArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  ingest_body(a, 0, a.rows, static_cast<int>(tid >> 5), static_cast<int>(nthr >> 5), threadIdx.x & 31,
              blockIdx.x == 0 && threadIdx.x == 0);
}

__global__ void __launch_bounds__(256) bn_act_kernel(BnActArgs a, int rows_per_block) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[EW_SCRATCH_BYTES];
  bn_act_body<false>(a, rows_per_block, blockIdx.x, blockIdx.y, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) bn_bwd_kernel(BnBwdArgs a, int rows_per_block) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[EW_SCRATCH_BYTES];
  bn_bwd_body<false>(a, rows_per_block, blockIdx.x, blockIdx.y, threadIdx.x, scratch);
}

// Two independent BatchNorm layers in one launch: blocks [0, nb0) work on the first, the rest on the second.
__global__ void __launch_bounds__(256) bn_act_pair_kernel(BnActArgs a0, int rpb0, int gx0, int nb0, BnActArgs a1, int rpb1, int gx1) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[EW_SCRATCH_BYTES];
  const int b = static_cast<int>(blockIdx.x);
  if (b < nb0) bn_act_body<false>(a0, rpb0, b % gx0, b / gx0, threadIdx.x, scratch);
  else bn_act_body<false>(a1, rpb1, (b - nb0) % gx1, (b - nb0) / gx1, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) bn_bwd_pair_kernel(BnBwdArgs a0, int rpb0, int gx0, int nb0, BnBwdArgs a1, int rpb1, int gx1) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[EW_SCRATCH_BYTES];
  const int b = static_cast<int>(blockIdx.x);
  if (b < nb0) bn_bwd_body<false>(a0, rpb0, b % gx0, b / gx0, threadIdx.x, scratch);
  else bn_bwd_body<false>(a1, rpb1, (b - nb0) % gx1, (b - nb0) / gx1, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) latent_fwd_kernel(LatentFwdArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[256];
  latent_fwd_body<false>(a, blockIdx.x, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) latent_fwd_rows_kernel(LatentFwdArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[256];
  constexpr int RPB = CHAIN_ROWS / CHAIN_CLUSTER;
  const int r0 = blockIdx.x * RPB;
  latent_fwd_rows<false>(a, r0, min(a.rows, r0 + RPB), blockIdx.x, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) latent_bwd_kernel(LatentBwdArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  latent_bwd_body(a, blockIdx.x, threadIdx.x);
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_kernel(LossArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[256 + LOSS_THREADS * 8];
  loss_body<false>(a, blockIdx.x, gridDim.x, threadIdx.x, scratch);
}

// ---------------------------------------------------------------------------------------------
// Autograd path: upstream fp32 dL/d(recon) -> bf16 operands of the backward GEMMs
// ---------------------------------------------------------------------------------------------
struct OutGradPack { OutGradArgs e[3]; int n; };
__global__ void out_grad_kernel(OutGradPack p) {
  pdl_wait();
  pdl_launch_dependents();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int e = 0; e < p.n; ++e) {
    const OutGradArgs& a = p.e[e];
    const long long total = static_cast<long long>(a.rows) * a.width;
    for (long long i = tid; i < total; i += nthr) {
      float g = a.g[i];
      if (a.y) { const float y = a.y[i]; g *= y * (1.0f - y); }
      const int r = static_cast<int>(i / a.width);
      const int c = static_cast<int>(i - static_cast<long long>(r) * a.width);
      a.dst[static_cast<size_t>(r) * a.ld_dst + c] = __float2bfloat16(g);
    }
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  adamw_body(a, blockIdx.x, threadIdx.x);
  if (a.has_tail && blockIdx.x == 0) {
    __shared__ double dsh[32];
    loss_tail_reduce(a.tail, threadIdx.x, dsh);
  }
  if (a.gframed && a.sums_out && blockIdx.x == 0 && threadIdx.x < 2) {
    // data parallel: the summed loss scalars travel behind the gradients (vla_b200.h, vla_dp_losses)
    const unsigned int epoch = static_cast<unsigned int>(__ldcg(&a.dyn->dp_epoch));
    const uint4* w = a.gframed + a.tail2 + threadIdx.x;
    const float2 v = finish_framed(w, ld_framed(w), epoch);
    a.sums_out[2 * threadIdx.x] = v.x; a.sums_out[2 * threadIdx.x + 1] = v.y;
  }
}

// The same with the chunk offsets as a kernel parameter (single model, no exchange): one round of loads.
__global__ void __launch_bounds__(256) adamw_hinted_kernel(AdamArgs a, const __grid_constant__ AdamHints h) {
  pdl_wait();
  pdl_launch_dependents();
  adamw_body(a, blockIdx.x, threadIdx.x, nullptr, static_cast<long long>(h.off4[blockIdx.x]) * 4, h.arena_elems);
  if (a.has_tail && blockIdx.x == 0) {
    __shared__ double dsh[32];
    loss_tail_reduce(a.tail, threadIdx.x, dsh);
  }
}

// ---------------------------------------------------------------------------------------------
// Lock-step population step: the same bodies, one launch for every member (vla_internal.h, "Lock-step population step").
// ---------------------------------------------------------------------------------------------
struct MultiSel { int member; int local; int gx; int aux; };

// Finds the member that owns this block and stages its argument structure in shared memory (the tables were written when
// the plan was built, not by an earlier kernel of the stream: they may be read before griddepcontrol.wait).
template <typename Args>
__device__ __forceinline__ const Args& multi_select(const MultiHdr* __restrict__ hdr, const void* __restrict__ args, int stride,
                                                    int n, MultiSel* sel, uint32_t* stage) {
  __shared__ MultiSel s_sel;
  const int bid = static_cast<int>(blockIdx.x);
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const MultiHdr h = hdr[t];
    const int e = t + 1 < n ? hdr[t + 1].block_begin : 0x7fffffff;
    if (bid >= h.block_begin && bid < e) { s_sel.member = t; s_sel.local = bid - h.block_begin; s_sel.gx = h.gx; s_sel.aux = h.aux; }
  }
  __syncthreads();
  *sel = s_sel;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(static_cast<const char*>(args) + static_cast<size_t>(sel->member) * stride);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(Args) / 4); i += blockDim.x) stage[i] = src[i];
  __syncthreads();
  return *reinterpret_cast<const Args*>(stage);
}
#define MULTI_STAGE(Args) __shared__ __align__(16) uint32_t stage[(sizeof(Args) + 15) / 16 * 4]; static_assert(sizeof(Args) % 4 == 0, "argument structures are copied in words")

__global__ void __launch_bounds__(256) ingest_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(IngestArgs);
  MultiSel sel;
  const IngestArgs& a = multi_select<IngestArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  const int tid = sel.local * blockDim.x + threadIdx.x;
  const int nthr = sel.gx * blockDim.x;                 // gx = blocks of this member
  ingest_body(a, 0, a.rows, tid >> 5, nthr >> 5, threadIdx.x & 31, sel.local == 0 && threadIdx.x == 0);
}

__global__ void __launch_bounds__(256) bn_act_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(BnActArgs);
  MultiSel sel;
  const BnActArgs& a = multi_select<BnActArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[EW_SCRATCH_BYTES];
  bn_act_body<false>(a, sel.aux, sel.local % sel.gx, sel.local / sel.gx, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) bn_bwd_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(BnBwdArgs);
  MultiSel sel;
  const BnBwdArgs& a = multi_select<BnBwdArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[EW_SCRATCH_BYTES];
  bn_bwd_body<false>(a, sel.aux, sel.local % sel.gx, sel.local / sel.gx, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) latent_fwd_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(LatentFwdArgs);
  MultiSel sel;
  const LatentFwdArgs& a = multi_select<LatentFwdArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[256];
  latent_fwd_body<false>(a, sel.local, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) latent_bwd_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(LatentBwdArgs);
  MultiSel sel;
  const LatentBwdArgs& a = multi_select<LatentBwdArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  latent_bwd_body(a, sel.local, threadIdx.x);
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(LossArgs);
  MultiSel sel;
  const LossArgs& a = multi_select<LossArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) unsigned char scratch[256 + LOSS_THREADS * 8];
  loss_body<false>(a, sel.local, sel.gx, threadIdx.x, scratch);
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const MultiHdr* hdr, const void* args, int stride, int n) {
  MULTI_STAGE(AdamArgs);
  MultiSel sel;
  const AdamArgs& a = multi_select<AdamArgs>(hdr, args, stride, n, &sel, stage);
  pdl_wait();
  pdl_launch_dependents();
  adamw_body(a, sel.local, threadIdx.x);
}

// Appends a launch to the thread's recorder (vla_train_step_group) instead of issuing it.
template <typename Args>
bool record_launch(int kind, const Args& a, int blocks, int gx, int aux) {
  Recorder* r = recorder();
  if (!r) return false;
  RecOp op;
  op.kind = kind; op.blocks = blocks; op.gx = gx; op.aux = aux;
  op.args.assign(reinterpret_cast<const char*>(&a), sizeof(Args));
  r->ops.push_back(std::move(op));
  return true;
}

inline int grid_for(long long work_items, int threads, int max_blocks) {
  long long b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace

cudaError_t launch_ingest(const IngestArgs& a, cudaStream_t s) {
  // one warp per row, 8 warps per block
  const int grid = grid_for(static_cast<long long>(a.rows) * 32, 256, 148 * 8);
  if (record_launch(RK_INGEST, a, grid, grid, 0)) return cudaSuccess;
  return launch_pdl(ingest_kernel, dim3(grid), dim3(256), 0, s, a);
}

// Rows per block of the BatchNorm apply / backward kernels (a block = 64 columns x this many rows).  Every block re-reduces
// the per-tile statistics of its 64 columns, and the kernels hold 76-112 registers (2-3 blocks of 256 threads per SM), so the
// grid is sized to ONE wave of about two blocks per SM: wide layers (512 columns = 8 column blocks) get 128-row blocks.
int bn_rows_per_block(int rows, int m_tiles, int n) {
  const int col_blocks = (n + BN_COLS - 1) / BN_COLS;
  long long rpb = (static_cast<long long>(rows) * col_blocks + 2 * 148 - 1) / (2 * 148);
  rpb = (rpb + 31) / 32 * 32;
  if (rpb < 32) rpb = 32;
  if (m_tiles > rpb) rpb = (m_tiles + 31) / 32 * 32;      // keep the per-block re-reduction of the tile partials below the tile's own traffic
  return static_cast<int>(rpb);
}

cudaError_t launch_bn_act(const BnActArgs& a, cudaStream_t s) {
  if (a.n % 2) return cudaErrorInvalidValue;
  const int rpb = bn_rows_per_block(a.rows, a.train ? a.m_tiles : 0, a.n);
  dim3 grid((a.n + BN_COLS - 1) / BN_COLS, (a.rows + rpb - 1) / rpb);
  if (record_launch(RK_BN_ACT, a, grid.x * grid.y, grid.x, rpb)) return cudaSuccess;
  return launch_pdl(bn_act_kernel, grid, dim3(256), 0, s, a, rpb);
}

cudaError_t launch_bn_act_pair(const BnActArgs& a0, const BnActArgs& a1, cudaStream_t s) {
  if ((a0.n % 2) || (a1.n % 2)) return cudaErrorInvalidValue;
  // (the pair shares one wave: size each layer's blocks as if the other's elements were its own)
  const int rpb0 = bn_rows_per_block(a0.rows, a0.train ? a0.m_tiles : 0, a0.n + a1.n);
  const int rpb1 = bn_rows_per_block(a1.rows, a1.train ? a1.m_tiles : 0, a0.n + a1.n);
  const int gx0 = (a0.n + BN_COLS - 1) / BN_COLS, gy0 = (a0.rows + rpb0 - 1) / rpb0;
  const int gx1 = (a1.n + BN_COLS - 1) / BN_COLS, gy1 = (a1.rows + rpb1 - 1) / rpb1;
  return launch_pdl(bn_act_pair_kernel, dim3(gx0 * gy0 + gx1 * gy1), dim3(256), 0, s, a0, rpb0, gx0, gx0 * gy0, a1, rpb1, gx1);
}

cudaError_t launch_bn_bwd_pair(const BnBwdArgs& a0, const BnBwdArgs& a1, cudaStream_t s) {
  if ((a0.n % 2) || (a1.n % 2)) return cudaErrorInvalidValue;
  const int rpb0 = bn_rows_per_block(a0.rows, a0.m_tiles, a0.n + a1.n);
  const int rpb1 = bn_rows_per_block(a1.rows, a1.m_tiles, a0.n + a1.n);
  const int gx0 = (a0.n + BN_COLS - 1) / BN_COLS, gy0 = (a0.rows + rpb0 - 1) / rpb0;
  const int gx1 = (a1.n + BN_COLS - 1) / BN_COLS, gy1 = (a1.rows + rpb1 - 1) / rpb1;
  return launch_pdl(bn_bwd_pair_kernel, dim3(gx0 * gy0 + gx1 * gy1), dim3(256), 0, s, a0, rpb0, gx0, gx0 * gy0, a1, rpb1, gx1);
}

cudaError_t launch_bn_bwd(const BnBwdArgs& a, cudaStream_t s) {
  if (a.n % 2) return cudaErrorInvalidValue;
  const int rpb = bn_rows_per_block(a.rows, a.m_tiles, a.n);
  dim3 grid((a.n + BN_COLS - 1) / BN_COLS, (a.rows + rpb - 1) / rpb);
  if (record_launch(RK_BN_BWD, a, grid.x * grid.y, grid.x, rpb)) return cudaSuccess;
  return launch_pdl(bn_bwd_kernel, grid, dim3(256), 0, s, a, rpb);
}

cudaError_t launch_latent_fwd(const LatentFwdArgs& a, int* grid_out, cudaStream_t s) {
  const long long total = static_cast<long long>(a.rows) * a.L;
  const int grid = static_cast<int>((total + 255) / 256);
  if (grid_out) *grid_out = grid;
  if (record_launch(RK_LATENT_FWD, a, grid, grid, 0)) return cudaSuccess;
  return launch_pdl(latent_fwd_kernel, dim3(grid), dim3(256), 0, s, a);
}

cudaError_t launch_latent_fwd_rows(const LatentFwdArgs& a, cudaStream_t s) {
  const int blocks = CHAIN_CLUSTER * ((a.rows + CHAIN_ROWS - 1) / CHAIN_ROWS);      // (slices past the last row write a zero partial)
  if (recorder()) { recorder()->unsupported = true; return cudaErrorNotSupported; }
  return launch_pdl(latent_fwd_rows_kernel, dim3(blocks), dim3(256), 0, s, a);
}

cudaError_t launch_latent_bwd(const LatentBwdArgs& a, cudaStream_t s) {
  const long long total = static_cast<long long>(a.rows) * a.L;
  const int grid = static_cast<int>((total + 255) / 256);
  if (record_launch(RK_LATENT_BWD, a, grid, grid, 0)) return cudaSuccess;
  return launch_pdl(latent_bwd_kernel, dim3(grid), dim3(256), 0, s, a);
}

int loss_grid_size(int rows, int width_a, int width_b, int n_sites) {
  LossArgs a{};
  a.rows = rows;
  a.recon_a = width_a ? reinterpret_cast<const float*>(1) : nullptr; a.width_a = width_a;
  a.recon_b = width_b ? reinterpret_cast<const float*>(1) : nullptr; a.width_b = width_b;
  a.logits = n_sites ? reinterpret_cast<const float*>(1) : nullptr; a.n_sites = n_sites;
  a.mu = reinterpret_cast<const float*>(1); a.L = 128;
  const LossGrid g = loss_grid(a);
  return g.nb_a + g.nb_b + g.nb_c + g.nb_k + 1;
}

cudaError_t launch_loss(const LossArgs& a, cudaStream_t s) {
  const LossGrid g = loss_grid(a);
  int grid = g.nb_a + g.nb_b + g.nb_c + g.nb_k;
  if (grid == 0) grid = 1;   // KL-from-partials only: one block does the final reduction
  if (record_launch(RK_LOSS, a, grid, grid, 0)) return cudaSuccess;
  return launch_pdl(loss_kernel, dim3(grid), dim3(LOSS_THREADS), 0, s, a);
}

cudaError_t launch_out_grad(const OutGradArgs* a, int n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (recorder()) { recorder()->unsupported = true; return cudaErrorNotSupported; }
  OutGradPack p{};
  p.n = n;
  long long work = 0;
  for (int i = 0; i < n; ++i) { p.e[i] = a[i]; work = std::max(work, static_cast<long long>(a[i].rows) * a[i].width); }
  return launch_pdl(out_grad_kernel, dim3(grid_for(work, 256, 148 * 8)), dim3(256), 0, s, p);
}

cudaError_t launch_adamw(const AdamArgs& a, cudaStream_t s, const AdamHints* hints) {
  if (a.n_chunks <= 0) return cudaSuccess;
  if (a.gframed == nullptr && record_launch(RK_ADAMW, a, a.n_chunks, a.n_chunks, 0)) return cudaSuccess;
  if (recorder()) { recorder()->unsupported = true; return cudaErrorNotSupported; }
  if (hints && hints->n == a.n_chunks && a.gframed == nullptr && a.update)
    return launch_pdl(adamw_hinted_kernel, dim3(a.n_chunks), dim3(256), 0, s, a, *hints);
  return launch_pdl(adamw_kernel, dim3(a.n_chunks), dim3(256), 0, s, a);
}

Recorder*& recorder() {
  static thread_local Recorder* r = nullptr;
  return r;
}

cudaError_t launch_multi(int kind, int variant, const MultiHdr* hdr, const void* args, int stride, int n, int total_blocks,
                         cudaStream_t s) {
  (void)variant;
  if (n <= 0 || total_blocks <= 0) return cudaSuccess;
  if (n > MULTI_MAX_MEMBERS) return cudaErrorInvalidValue;
  const dim3 grid(total_blocks);
  switch (kind) {
    case RK_INGEST: return launch_pdl(ingest_multi_kernel, grid, dim3(256), 0, s, hdr, args, stride, n);
    case RK_BN_ACT: return launch_pdl(bn_act_multi_kernel, grid, dim3(256), 0, s, hdr, args, stride, n);
    case RK_BN_BWD: return launch_pdl(bn_bwd_multi_kernel, grid, dim3(256), 0, s, hdr, args, stride, n);
    case RK_LATENT_FWD: return launch_pdl(latent_fwd_multi_kernel, grid, dim3(256), 0, s, hdr, args, stride, n);
    case RK_LATENT_BWD: return launch_pdl(latent_bwd_multi_kernel, grid, dim3(256), 0, s, hdr, args, stride, n);
    case RK_LOSS: return launch_pdl(loss_multi_kernel, grid, dim3(LOSS_THREADS), 0, s, hdr, args, stride, n);
    case RK_ADAMW: return launch_pdl(adamw_multi_kernel, grid, dim3(256), 0, s, hdr, args, stride, n);
    default: return cudaErrorInvalidValue;
  }
}

namespace {

// One warp per gathered row: the row's two feature vectors stream as 128-bit loads where the geometry allows (dim % 4 == 0
// keeps every row 16-byte aligned), else as scalars; lane 0 copies the label.
__global__ void __launch_bounds__(256) gather_rows_kernel(GatherArgs g) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < g.n; r += warps) {
    long long src = g.index[r];
    src = src < 0 ? 0 : (src >= g.rows ? g.rows - 1 : src);
    const float* pa = g.a + src * g.dim_a; float* qa = g.out_a + static_cast<long long>(r) * g.dim_a;
    const float* pb = g.b + src * g.dim_b; float* qb = g.out_b + static_cast<long long>(r) * g.dim_b;
    if ((g.dim_a & 3) == 0) { for (int i = lane; i < g.dim_a / 4; i += 32) reinterpret_cast<float4*>(qa)[i] = __ldg(reinterpret_cast<const float4*>(pa) + i); }
    else if ((g.dim_a & 1) == 0) { for (int i = lane; i < g.dim_a / 2; i += 32) reinterpret_cast<float2*>(qa)[i] = __ldg(reinterpret_cast<const float2*>(pa) + i); }
    else { for (int i = lane; i < g.dim_a; i += 32) qa[i] = __ldg(pa + i); }
    if ((g.dim_b & 3) == 0) { for (int i = lane; i < g.dim_b / 4; i += 32) reinterpret_cast<float4*>(qb)[i] = __ldg(reinterpret_cast<const float4*>(pb) + i); }
    else if ((g.dim_b & 1) == 0) { for (int i = lane; i < g.dim_b / 2; i += 32) reinterpret_cast<float2*>(qb)[i] = __ldg(reinterpret_cast<const float2*>(pb) + i); }
    else { for (int i = lane; i < g.dim_b; i += 32) qb[i] = __ldg(pb + i); }
    if (lane == 0) g.out_site[r] = g.site[src];
  }
}

__global__ void __launch_bounds__(256) scale_kernel(ScaleArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const float sc = *a.scale;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int t = 0; t < a.count; ++t)
    for (long long i = tid; i < a.n[t]; i += nthr) a.x[t][i] *= sc;
}

}  // namespace

cudaError_t launch_gather_rows(const GatherArgs& g, cudaStream_t s) {
  if (g.n <= 0) return cudaSuccess;
  const int blocks = std::min(148 * 8, (g.n + 7) / 8);
  return launch_pdl(gather_rows_kernel, dim3(blocks), dim3(256), 0, s, g);
}
cudaError_t launch_scale(const ScaleArgs& a, cudaStream_t s) {
  long long total = 0;
  for (int t = 0; t < a.count; ++t) total += a.n[t];
  if (total <= 0) return cudaSuccess;
  const int blocks = static_cast<int>(std::min<long long>(148 * 4, (total + 1023) / 1024));
  return launch_pdl(scale_kernel, dim3(blocks), dim3(256), 0, s, a);
}

}  // namespace vla
