// Internal declarations shared by the kernels and the host-side plan builder.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <string>
#include <vector>

namespace vla {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// Grouped tcgen05 GEMM
// ---------------------------------------------------------------------------------------------
enum GemmFlags : int {
  GF_BIAS = 1 << 0,       // v += bias[col]
  GF_RELU = 1 << 1,       // v = max(v, 0)
  GF_SIGMOID = 1 << 2,    // v = 1 / (1 + exp(-v))
  GF_COLSTATS = 1 << 3,   // per-tile column sum / sum of squares of v (after bias, before activation)
  GF_MASK = 1 << 4,       // v = mask_src[row, col] > 0 ? v * mask_scale : 0  (ReLU / dropout backward)
  GF_BNSTATS = 1 << 5,    // per-tile column sums of v and v * xhat, xhat = (pre - mean) * rstd
  GF_OUT_F32 = 1 << 6,    // store fp32
  GF_OUT_BF16 = 1 << 7,   // store bf16
  GF_RED = 1 << 8,        // accumulate into out_f32 with red.global.add (split-K weight gradients)
  GF_BIASGRAD = 1 << 9,   // TN mode: also produce sum_k A[m, k] into bias_grad[m] (ones-MMA)
  GF_LOSS = 1 << 11,      // NT mode, last decoder layer: loss value partials + dL/d(pre-activation) as bf16 (loss_kind)
  // compile-time only (never set in GemmProblem::flags): which loss kinds an instantiation of the loss epilogue contains
  GF_LK_MSE = 1 << 12, GF_LK_BCE = 1 << 13, GF_LK_CE = 1 << 14,
  GF_LATBWD = 1 << 15,    // NN mode, data gradient of the fused first decoder layer: the tile's result is dL/dz; the epilogue
                          // turns it into d(mu | logvar) (reparameterisation + KL backward, divided by the number of
                          // modalities) and stores the bf16 operand of the heads' backward -- no separate latent_bwd launch
};
enum LossKind : int { LOSS_NONE = 0, LOSS_MSE = 1, LOSS_BCE = 2, LOSS_CE = 3 };

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
// Pipeline: GEMM_STAGES hand-shakes in flight, each covering GEMM_GROUP consecutive 64-wide k-blocks (one shared-memory slot
// per k-block).  A hand-shake (mbarrier wait -> expect_tx -> TMA issue | mbarrier wait -> MMA issue -> tcgen05.commit) costs
// ~0.33 us whatever it carries (profiles/r2_rowchain_experiments.md); at one k-block per hand-shake the main loops ran at
// 40 % tensor occupancy, so a stage carries two (split-bf16 operands: the hi and the lo copy of one k-block).
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_GROUP = 2;
constexpr int GEMM_SLOTS = GEMM_STAGES * GEMM_GROUP;
constexpr int GEMM_BN_MAX_NT = 144;   // multiple of 16
constexpr int GEMM_BN_MAX_TN = 128;   // multiple of 64
constexpr int GEMM_MAX_PROBLEMS = 16;
constexpr int GEMM_THREADS = 320;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int GEMM_TMEM_COLS = 256;
constexpr int GEMM_BIAS_TMEM_COL = 224;

// Everything a tile needs besides its two tensor maps (the chain kernel stages these in shared memory, phase by phase).
struct GemmScalars {
  int M, N, K;
  int BN;
  int m_tiles, n_tiles, k_splits, kb_per_split;
  int tile_begin;
  int flags;
  int ld_f32, ld_bf16, ld_mask, ld_pre;
  float mask_scale;
  int pad0;
  const float* bias;
  float* out_f32;
  bf16* out_bf16;
  const bf16* mask_src;
  const float* pre;
  const float* mean;
  const float* rstd;
  float* stats;
  float* bias_grad;
  // ---- fused tails ----
  // GF_LOSS:       aux0 = target fp32 [rows, N] dense (MSE / BCE), aux1 = class weights or nullptr (CE), aux_site = labels
  //                (CE); the gradient goes to out_bf16; aux_partials[local_tile * 8 + warp] receives the loss partial sums;
  //                aux_n = resident batches (targets start at row (dyn->batch_index % aux_n) * M when aux_n > 1).
  const float* aux0; const float* aux1;
  const long long* aux_site;
  float* aux_partials;
  const struct DynParams* dyn;
  int aux_n, loss_kind;
  float aux_scale;                   // CE: gamma when dyn == nullptr
  // GF_LATBWD:     aux0 = mu, aux1 = logvar, pre = eps (fp32 [rows, N], N = latent width), mean / rstd = incoming d(mu) /
  //                d(logvar) of the autograd path or nullptr, out_bf16 / ld_bf16 = the [mu | logvar] gradient rows,
  //                aux_n = modalities averaged by the forward, loss_kind != 0: autoencoder (no logvar half),
  //                beta = dyn->beta_kl or aux_scale.
  // ---- split-bf16 operands (NT mode) ----
  // a_lo > 0: both operands carry a second bf16 copy ("lo" = bf16(x - bf16(x))) a_lo / b_lo elements further along K in the
  // same rows; the tile then accumulates A_hi B_hi + A_lo B_hi + A_hi B_lo, i.e. operands of ~16 mantissa bits.  Used by
  // every forward GEMM whose result reaches a ReLU (DESIGN.md "Precision").
  int a_lo, b_lo;
  int out_lo;                        // > 0: GF_OUT_BF16 also stores the lo copy of the result out_lo elements further
  // ---- ReLU / dropout masks as bits (one 32-bit word per row and 32-column chunk, [chunk][M] words) ----
  // GF_OUT_BF16 forward tiles with mask_bits_out != nullptr record (value > 0) per element; GF_MASK tiles with
  // mask_bits_in != nullptr read the bits instead of the bf16 activation (one coalesced word per thread and chunk).
  unsigned int* mask_bits_out;
  const unsigned int* mask_bits_in;
  int pad1[2];
};
static_assert(sizeof(GemmScalars) % 16 == 0 && sizeof(GemmScalars) <= 224, "GemmScalars is copied in 16-byte pieces into a 224-byte slot");

struct alignas(64) GemmProblem : GemmScalars {
  CUtensorMap tmA;
  CUtensorMap tmB;
};

// Final reduction of the fused loss: the epilogue of the tile that takes the last ticket sums every partial in a fixed
// order (deterministic) and writes out[4] = {total, recon, class, kld} (losses.py:27-46).
struct LossTail {
  unsigned int* counter;       // zero between steps (re-armed by the last tile); nullptr: the tiles only write their partials and
                               // the step's AdamW launch does the final reduction (AdamArgs::tail) -- no ticket on the chain
  int total_tickets;           // loss tiles of the whole step (possibly spread over several launches)
  int n_mse, n_bce, n_ce;      // partial floats per term, contiguous in `partials` in this order
  const float* partials;
  const float* kl_partials; int n_kl;
  float* out;
  const struct DynParams* dyn; // beta / gamma
  struct DynParams* dyn_bump;  // batch_index += 1 after the reduction
  int pad[2];
};

struct GemmGroup {
  int nprob;
  int total_tiles;
  unsigned long long* dbg;   // optional [total_tiles][8] globaltimer stamps (test hook only)
  int dbg_flags;             // test hook: 1 = skip epilogue stores, 2 = skip main loop, 4 = skip TMEM alloc (with 2)
  int pad0;
  LossTail tail;             // used by problems with GF_LOSS
  int pad[8];                // pad[0] = 4: NT group laid out for the 4-CTA multicast kernel (A maps with 32-row boxes)
  GemmProblem p[GEMM_MAX_PROBLEMS];
};
static_assert(sizeof(LossTail) == 80, "LossTail layout");
static_assert(offsetof(GemmGroup, p) % 64 == 0, "tensor maps need 64-byte alignment");

// mode 0: NT  C = A[M,K] B[N,K]^T   (both K-major; forward)
// mode 1: TN  C = A[K,M]^T B[K,N]   (both MN-major; weight gradients, split-K)
// mode 2: NN  C = A[M,K] B[K,N]     (A K-major, B MN-major; data gradients read the forward weight copy [out,in] directly)
cudaError_t launch_gemm_group(const GemmGroup& g, int mode, cudaStream_t stream);

// Launch with the programmatic-stream-serialization attribute (PDL); every kernel of this library calls
// griddepcontrol.wait before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static const bool pdl_on = [] { const char* e = getenv("VLA_NO_PDL"); return !(e && e[0] == '1'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
size_t gemm_smem_bytes();

// Builds a 2-D bf16 tensor map with 128-byte swizzle.  inner/outer are extents in elements,
// pitch_bytes the outer stride; box_inner must be 64 (=128 B).  Returns false on failure.
bool make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                    uint32_t box_inner, uint32_t box_outer, std::string* err);

// ---------------------------------------------------------------------------------------------
// Element-wise / reduction kernels (elementwise.cu)
// ---------------------------------------------------------------------------------------------
// Per-step scalars that live on the device so that a captured CUDA graph can be replayed while the
// learning rate, the KL weight (beta warm-up, train_rna2dna.py:80) and the step count change.
struct DynParams {
  float lr, weight_decay, beta_kl, gamma;
  int step;          // number of the optimizer step in flight (1-based); bumped by the ingest kernel
  int batch_index;   // which resident batch the step in flight trains on; bumped by the loss kernel's last block
  int dp_epoch;      // train steps started on this handle since creation (never reset): the epoch of the peer-memory
                     // gradient exchange (dp_exchange.cu); bumped together with `step`
  int pad;
  double b1pow, b2pow;   // beta1^step, beta2^step, advanced together with `step` (no powf in the AdamW kernel)
};

struct IngestArgs {           // fp32 [rows, width] (dense) -> bf16 [rows, ld_dst]
  const float* src[2];
  bf16* dst[2];
  int width[2];
  int ld_dst[2];
  int lo_off[2];              // > 0: also write the lo copy bf16(x - bf16(x)) lo_off elements further along the row
  int n;                      // number of active entries (0..2)
  int rows;
  // site encoder input: h_site[r, :] = bf16(emb[site[r], :]) and onehot[r, s] = (site[r] == s)
  const long long* site; const float* emb; bf16* h_site; int ld_hsite; bf16* onehot; int ld_onehot;
  int n_sites, embed;
  int hsite_lo;               // > 0: lo copy of the gathered embedding rows
  // train step bookkeeping: dyn->step += 1 (first kernel of a train step)
  struct DynParams* dyn; int bump_step;
  float beta1, beta2;         // Adam betas, to advance dyn->b1pow / b2pow with the step
  int n_batches;              // > 1: rows are read at offset (dyn->batch_index % n_batches) * rows
};
cudaError_t launch_ingest(const IngestArgs& a, cudaStream_t s);

struct BnActArgs {
  const float* pre; int ld_pre;         // [rows, n]
  const float* stats; int m_tiles;      // [m_tiles][2][n] (train)
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float* save_mean; float* save_rstd;   // [n]
  bf16* out; int ld_out;
  int out_lo;                           // > 0: lo copy of the activation out_lo elements further along the row
  const unsigned char* keep_mask;       // optional injected keep mask [rows, n]
  int rows, n;
  int stat_rows;                        // > 0: rows behind the statistics (SyncBN: the global batch); 0 = rows
  int train;                            // batch statistics + dropout
  int update_running;
  float p_drop;
  unsigned long long seed, offset;      // Philox stream for dropout when keep_mask == nullptr
  const struct DynParams* dyn;          // if set, the Philox offset also mixes in dyn->step
};
cudaError_t launch_bn_act(const BnActArgs& a, cudaStream_t s);
// Two independent BatchNorm layers (the two dense encoders' layers of one round) as ONE launch.
cudaError_t launch_bn_act_pair(const BnActArgs& a0, const BnActArgs& a1, cudaStream_t s);

struct BnBwdArgs {
  const bf16* gy; int ld_gy;            // dL/dy (already through dropout and ReLU), [rows, n]
  const float* pre; int ld_pre;
  const float* stats; int m_tiles;      // [m_tiles][2][n]: sum gy, sum gy * xhat
  const float* mean; const float* rstd; const float* gamma;
  float* dgamma; float* dbeta;
  bf16* gpre; int ld_gpre;
  int rows, n, train;
  int stat_rows;                        // > 0: rows behind the statistics (SyncBN: the global batch); 0 = rows
  float param_grad_scale;               // SyncBN: dgamma / dbeta come out of GLOBAL sums on every rank -> 1 / world, so that the
                                        // step's all-reduce(SUM) of the gradient arena delivers them once; else 1
};
cudaError_t launch_bn_bwd(const BnBwdArgs& a, cudaStream_t s);
cudaError_t launch_bn_bwd_pair(const BnBwdArgs& a0, const BnBwdArgs& a1, cudaStream_t s);

struct LatentFwdArgs {
  const float* ml[3]; int ld_ml[3]; int n_enc;    // present encoders' heads output [rows, 2L]
  const float* eps_in;                            // injected eps [rows, L] or nullptr -> Philox
  unsigned long long seed, offset; const struct DynParams* dyn;
  float* mu; float* logvar;                       // outputs [rows, L] dense
  float* eps_save;                                // [rows, L]
  bf16* z; int ld_z;                              // [rows, ld_z]
  int z_lo;                                       // > 0: lo copy of z
  float* kl_partials;                             // [grid]
  int rows, L;
  int ae;                                         // autoencoder: heads are [rows, L]; z = mu = mean, no sampling, KL = 0
};
cudaError_t launch_latent_fwd(const LatentFwdArgs& a, int* grid_out, cudaStream_t s);
// Same step with the chain kernel's partial layout: one block and one KL partial per 32-row slice.
cudaError_t launch_latent_fwd_rows(const LatentFwdArgs& a, cudaStream_t s);

struct LatentBwdArgs {
  const float* gz; int ld_gz;                     // [rows, L] (may be nullptr -> 0)
  const float* gmu_in; const float* glv_in;       // autograd-supplied dL/dmu, dL/dlogvar (optional)
  const float* mu; const float* logvar; const float* eps;
  float beta;                                     // engine mode: adds beta * dKL/d(mu, logvar)
  int n_modalities;
  bf16* gml; int ld_gml;                          // [rows, >= 2L]
  const struct DynParams* dyn;                    // if set, beta is read from dyn->beta_kl
  int rows, L;
  int ae;                                         // autoencoder: only d(latent) = (gz + gmu_in) / n_modalities, width L
};
cudaError_t launch_latent_bwd(const LatentBwdArgs& a, cudaStream_t s);

struct LossArgs {
  // MSE term
  const float* recon_a; const float* a; int width_a;
  // BCE term
  const float* recon_b; const float* b; int width_b;
  // CE term
  const float* logits; const long long* site; const float* class_w; int n_sites;
  // KL term: either from kl_partials (engine) or directly from mu / logvar
  const float* kl_partials; int n_kl_partials;
  const float* mu; const float* logvar; int L;
  float beta, gamma;
  const struct DynParams* dyn;   // if set, beta / gamma are read from it
  struct DynParams* dyn_bump;    // if set, the last block advances dyn_bump->batch_index
  int n_batches;                 // > 1: targets are read at row offset (dyn->batch_index % n_batches) * rows
  int rows;
  // gradient outputs (all optional).  bf16 outputs are what the backward GEMMs consume:
  bf16* ga_bf16; int ld_ga;      // 2 (recon_a - a)
  bf16* gb_bf16; int ld_gb;      // dL/dlogit_b = (y - t) * y(1-y) / max(y(1-y), 1e-12)
  bf16* gc_bf16; int ld_gc;      // gamma * w * (softmax - onehot)
  float* ga_f32; float* gb_f32; float* gc_f32;   // dense fp32 dL/d(recon) for the autograd path
  float* gmu_f32; float* glv_f32;
  float grad_scale;              // upstream dL/dtotal (autograd path), 1 for the engine
  float* partials;               // workspace [grid * 4]
  unsigned int* counter;         // zero-initialised ticket for the last-block reduction
  float* out;                    // [4]: total, recon, class, kld
};
cudaError_t launch_loss(const LossArgs& a, cudaStream_t s);
int loss_grid_size(int rows, int width_a, int width_b, int n_sites);

struct OutGradArgs {  // autograd path: fp32 upstream dL/d(recon) -> bf16 GEMM operands
  const float* g; int width;           // dense [rows, width]
  const float* y;                      // sigmoid output (multiply by y (1 - y)) or nullptr
  bf16* dst; int ld_dst;
  int rows;
};
cudaError_t launch_out_grad(const OutGradArgs* a, int n, cudaStream_t s);

// One 1024-element piece of a parameter tensor, self-contained so the kernel needs a single table read.
struct alignas(16) AdamChunk {
  long long offset;       // arena element offset of the chunk's first element (multiple of 4)
  long long shadow_off;   // bf16 copy [rows, ld_shadow] offset of the tensor, -1 = none
  int n;                  // valid elements in this chunk (<= ADAM_CHUNK)
  int first;              // element index of the chunk's first element inside its tensor
  int cols;               // tensor columns (vectors: 1)
  int ld_shadow;
  int sh_lo;              // > 0: the bf16 copy carries a lo part sh_lo elements further along the row
  int pad[3];
};
constexpr int ADAM_CHUNK = 1024;   // one float4 per thread
struct AdamArgs {
  float* p; float* g; float* m; float* v;
  bf16* shadow;
  const AdamChunk* chunks; int n_chunks;
  float lr, beta1, beta2, eps, weight_decay;
  float bc1, inv_bc2_sqrt;    // 1 - beta1^t and 1 / sqrt(1 - beta2^t), computed in double on the host
  const struct DynParams* dyn;// if set, lr / weight_decay / bias corrections are read from it
  int update;                 // 0: only refresh the bf16 shadows from p
  int zero_grad;              // 1: clear gclear after use
  float* gclear;              // what zero_grad clears (normally g itself)
  // data parallel (dp_exchange.cu): the gradients are the framed sums arriving from the shard owners -- word i holds
  // elements 2i, 2i + 1 -- polled as the kernel walks the parameters; epoch = dyn->dp_epoch.  Block 0 also copies the
  // summed loss scalars (framed words [tail2, tail2 + 2)) to sums_out[4].
  const uint4* gframed;
  long long tail2;
  float* sums_out;
  // deferred final reduction of the fused loss (single-GPU whole step): block 0 sums the loss tiles' partials in the fixed
  // order of the in-tile reduction, writes out[4] and advances the batch index
  LossTail tail; int has_tail; int pad_tail;
};
// Arena offsets of the chunks as a KERNEL PARAMETER (constant bank): a block can issue its p / g / m / v loads at once instead
// of after the round trip for its chunk-table entry (the launch is a chain of dependent load rounds, not a bandwidth problem).
constexpr int ADAM_HINT_CHUNKS = 2040;
struct AdamHints { unsigned int off4[ADAM_HINT_CHUNKS]; int n; long long arena_elems; };   // offset / 4; n = 0: no hints
cudaError_t launch_adamw(const AdamArgs& a, cudaStream_t s, const AdamHints* hints = nullptr);

// ---------------------------------------------------------------------------------------------
// Chain kernel (chain_kernel.cu): the ROW-LOCAL stretches of a step -- consecutive launches in which a 128-row block of
// the batch only ever depends on the same rows of the previous launch (everything between two BatchNorm statistics
// boundaries: BN apply -> heads -> latent -> decoders + loss -> data gradients -> latent backward -> encoder data
// gradients) -- run as ONE launch.  One 4-CTA thread-block cluster owns a 128-row block and walks the phases; each phase's
// tiles (N split over the cluster) or rows (element-wise phases, 32 per CTA) are spread over the four CTAs; a cluster
// barrier replaces the kernel boundary.  Activations travel through L2 (they are needed in global memory by the
// weight-gradient GEMMs anyway); TMEM, the mbarrier ring and the tensor-map prefetches are set up once per launch.
// ---------------------------------------------------------------------------------------------
enum ChainKind : int {
  CK_GEMM_NT_PLAIN = 0, CK_GEMM_NT_FULL, CK_GEMM_NT_LOSS, CK_GEMM_NT_LOSS_BCE, CK_GEMM_NT_LOSS_MSE, CK_GEMM_NN_PLAIN,
  CK_GEMM_NN_FULL, CK_GEMM_LAST = CK_GEMM_NN_FULL,
  CK_INGEST, CK_BN_ACT, CK_BN_BWD, CK_LATENT_FWD, CK_LATENT_BWD,
};
constexpr int CHAIN_CLUSTER = 4;
constexpr int CHAIN_MAX_PHASES = 24;
constexpr int CHAIN_MAX_UNITS = 12;       // tiles of one row block one CTA may be given in one phase
constexpr int CHAIN_ROWS = GEMM_BM;       // rows per cluster iteration

struct ChainPhase {
  int kind;
  int n_units[CHAIN_CLUSTER];                              // GEMM phases: tiles per cluster rank
  unsigned short units[CHAIN_CLUSTER][CHAIN_MAX_UNITS];    // (problem << 8) | n_tile
  long long args_off;                                      // byte offset of the argument struct from the plan base
};
struct ChainPlan {
  int n_phases;
  int rows;
  int m_blocks;                 // 128-row blocks
  int pad;
  unsigned long long* dbg;      // optional [clusters * 4][CHAIN_MAX_PHASES][8] %globaltimer stamps (include/vla_b200.h)
  ChainPhase ph[CHAIN_MAX_PHASES];
};
size_t chain_smem_bytes();
int chain_max_clusters(cudaError_t* err);
cudaError_t launch_chain(const ChainPlan* plan_dev, int n_clusters, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Row-chain kernel (rowchain.cu): the row-local middle of a directional model's train step with every activation ON CHIP.
// One CTA owns a 128-row block and runs BatchNorm apply -> heads -> latent -> decoder -> loss -> decoder data gradients ->
// latent backward -> encoder data gradients as one launch: A operands live in shared memory in the UMMA layout (nine
// [128 x 64] bf16 slots, written by the epilogue threads themselves), weights stream through a five-slot TMA ring that runs
// ahead across layers, accumulators live in TMEM (all 512 columns), ReLU / dropout masks travel as bits, the fp32 side
// inputs (loss targets, BatchNorm pre-activations) arrive as TMA tiles through the same ring, and every tensor the
// weight-gradient GEMMs need goes to global memory by TMA stores straight from the operand slots.  The plan is a list of ops
// (kernel parameter); producer, MMA issuer and epilogue warps walk it in lockstep through mbarriers -- no kernel boundary,
// no cluster barrier, no dependent global load between two layers.
// ---------------------------------------------------------------------------------------------
constexpr int RC_MAX_TMAPS = 28;
constexpr int RC_MAX_OPS = 32;
enum RcKind : int { RC_LOADA = 1, RC_BNACT, RC_GEMM, RC_EPI };
enum RcEpiKind : int { EP_LATENT = 1, EP_RELU, EP_LOSS, EP_MASK, EP_LATENT_BWD, EP_DGRAD_ENC };
struct RcStore { short tm, slot, kb, pad; };   // after the op: TMA-store kb [128 x 64] blocks, slot.. -> tensor map tm, columns 0..
struct RcOp {
  int kind, sub;
  // RC_GEMM / RC_LOADA: B (or the loaded A) = tensor map tm_b; nn = 1: B is MN-major (data gradients read the forward
  // weight copy); output columns [n0, n0 + n) of the layer in chunks of 128 into TMEM columns tmem_col..; kb k-blocks;
  // A = slots a_slot.. (hi) and a_lo_slot.. (lo, -1: plain bf16); b_lo = element offset of the lo copy along K.
  short tm_b, nn, n0, n, kb, a_slot, a_lo_slot, b_lo, tmem_col, commit, wait_lda, pad0;
  // RC_EPI / RC_BNACT: accumulator columns e_tmem.. (e_n valid), second accumulator e_tmem2 (e_n2); global column of
  // accumulator column 0 = e_col0; output slots (slot of global column 0): out_slot (hi), out_lo_slot (lo, -1 none),
  // out_slot2 (second output); side-input tiles [128 x 32] fp32 through the ring: tensor map tm_side, side_tiles of them.
  short e_tmem, e_tmem2, e_n, e_n2, e_col0, out_slot, out_lo_slot, out_slot2, tm_side, side_tiles, last_loss, relu;
  RcStore st[3];
  float fscale; int n_total;               // n_total: width of the whole tensor (mask pitch, validity of the last columns)
  const void* p[6];
};
struct alignas(64) RcPlan {
  CUtensorMap tm[RC_MAX_TMAPS];
  RcOp ops[RC_MAX_OPS];
  int n_ops, n_tm, rows, m_blocks;
  int n_batches, L, ae, n_enc;
  const DynParams* dyn; DynParams* dyn_bump;
  // the BatchNorm layer in front of the heads (RC_BNACT applies it, EP_DGRAD_ENC takes its backward statistics)
  const float* bn_stats; const float* bn_gamma; const float* bn_beta; float* bn_running_mean; float* bn_running_var;
  long long* bn_nbt; float* bn_save_mean; float* bn_save_rstd; const unsigned char* bn_keep;
  int bn_n, bn_m_tiles, train, pad0; float p_drop; int pad1;
  unsigned long long seed, bn_offset, lat_offset;
  // loss
  float* loss_partials; float* kl_partials; unsigned int* counter; float* loss_out; int loss_kind; int pad2;
  unsigned long long* dbg;                 // optional [CTAs][RC_MAX_OPS][4] %globaltimer stamps
};
cudaError_t launch_rowchain(const RcPlan& plan, int n_ctas, cudaStream_t s);
// fp32 tile map, box [32 columns x 128 rows], 128-byte swizzle (loss targets, BatchNorm pre-activations)
bool make_tmap_f32_tile(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, std::string* err);

// ---------------------------------------------------------------------------------------------
// Head block (headblock.cu): the small layers around the latent as ONE launch each way, on the CUDA cores.
// Forward:  BatchNorm apply + ReLU + dropout of every encoder's last hidden layer -> mu / logvar heads -> mean over the
//           encoders -> z = mu + eps * exp(logvar / 2), KL partial -> fused first decoder layer (+ ReLU).
// Backward: first decoder layer's data gradient -> d(mu | logvar) incl. beta * dKL -> heads' data gradients (ReLU / dropout
//           mask, BatchNorm backward statistics of the row block), site encoder's gradient w.r.t. the embedding rows.
// These layers have K or N <= 2 * latent: 6-12 k FMA per row.  As tensor-core launches they cost 5-10 us EACH of fixed
// latency (first operands, accumulator hand-over, epilogue) for 0.1 us of math; here a 32-row block keeps everything in
// shared memory and registers, fp32 throughout (master weights, no bf16 staging).
// ---------------------------------------------------------------------------------------------
constexpr int HB_ROWS = 32;
constexpr int HB_THREADS = 512;            // 16 warps per 32-row block: the small dot products need the latency hiding
struct HbEnc {
  int kind;                       // 0: dense encoder (last hidden layer behind a BatchNorm), 1: site embedding
  int in_dim;                     // BatchNorm width / embedding dimension (multiple of 4)
  // kind 0 -- forward
  const float* pre; const float* stats; int m_tiles; int train;
  const float* gamma; const float* beta; float* running_mean; float* running_var; long long* nbt;
  float* save_mean; float* save_rstd; const unsigned char* keep_mask; unsigned long long drop_offset;
  bf16* act; int ld_act; int pad0; unsigned int* bits;
  // kind 0 -- backward
  bf16* gy; float* bstats; float mask_scale; int pad1;
  // kind 1
  const long long* site; const float* emb; bf16* g_x; int ld_gx; int pad2;
  // heads (both kinds): W [HW, in_dim], b [HW] fp32
  const float* Wh; const float* bh;
};
struct HbArgs {
  HbEnc enc[3];
  int n_enc, rows, L, HW, ae, C, n_batches, pad0;
  float p_drop; int pad1;
  const float* eps_in; unsigned long long seed, lat_offset; const DynParams* dyn;
  float* mu; float* logvar; float* eps_save; bf16* z; int ld_z; int pad2; float* kl_partials;
  // fused first decoder layer: W0 [C, L], b0 [C]; forward output d0 (hi | lo) + bits, backward input g_d0 [rows, C]
  const float* W0; const float* b0; bf16* d0; int ld_d0, d0_lo; unsigned int* d0_bits;
  const bf16* g_d0; int ld_gd0; int pad3;
  // backward: upstream d/dmu, d/dlogvar (autograd path, optional), output d(mu | logvar) as bf16 [rows, ld_gml]
  const float* gmu_in; const float* glv_in; bf16* gml; int ld_gml; int has_dec;
};
size_t hb_smem_bytes(const HbArgs& a, bool backward);
cudaError_t launch_head_block_fwd(const HbArgs& a, cudaStream_t s);
cudaError_t launch_head_block_bwd(const HbArgs& a, cudaStream_t s);

int bn_rows_per_block(int rows, int m_tiles, int n);

// ---------------------------------------------------------------------------------------------
// Data-parallel gradient exchange over NVLink peer memory (dp_exchange.cu)
// ---------------------------------------------------------------------------------------------
constexpr int DP_MAX_WORLD = 16;
struct DpArgs {
  int world, rank;
  long long first2;                     // this launch exchanges float2s [first2, first2 + n2) of the flat buffer
  long long n2;
  long long per2;                       // float2 elements per shard of this range = ceil(n2 / world)
  float* g;                             // local gradients (whole flat buffer; cleared as they are read)
  uint4* recv[DP_MAX_WORLD];            // every rank's RECV region of this range, [world][per2] framed words (peer pointers)
  uint4* rsum[DP_MAX_WORLD];            // every rank's RSUM (whole flat buffer, framed words, indexed by float2 index)
  const DynParams* dyn;                 // epoch = dyn->dp_epoch
  unsigned long long* trace;            // optional [4] %globaltimer stamps of the last launch (block 0)
};
cudaError_t launch_dp_exchange(const DpArgs& a, cudaStream_t s, bool pdl);
// Opt-in SyncBN: all-reduce(SUM) over the ranks of a BatchNorm layer's column sums.  One block:
// reduces this rank's per-tile partials [m_tiles][2][n] (double), pushes the 2n sums framed into every rank's slot array,
// polls the peers' pushes, adds in rank order (bit-identical on every rank) and writes [2][n] over tile 0 of `partials`
// (each thread only touches its own columns).  The BatchNorm kernels then run with m_tiles = 1, stat_rows = world * rows.
constexpr int DP_SMALL_WORDS = 1024;      // framed words per slot: up to 1024 columns
constexpr int DP_SMALL_REGIONS = 8;       // (BatchNorm layer, direction) slots used within one step
struct DpSmallArgs {
  int world, rank;
  float* partials; int m_tiles, n;
  uint4* slots[DP_MAX_WORLD];           // every rank's slot array of this region: [world][DP_SMALL_WORDS]
  const DynParams* dyn;
};
cudaError_t launch_dp_small_allreduce(const DpSmallArgs& a, cudaStream_t s);
cudaError_t launch_dp_adamw(const DpArgs& x, const AdamArgs& a, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Reconstruction metrics (metrics.cu)
// ---------------------------------------------------------------------------------------------
struct MetricsArgs {
  const float* yt; const float* yp;     // y_true, y_pred: dense fp32 [rows, dim]
  long long rows; int dim;
  float* cos_out; float* pearson_out;   // optional per-sample outputs [rows] (Pearson: NaN where undefined)
  double* partials;                     // workspace [grid][8]
  unsigned int* counter;                // zero-initialised ticket (re-armed by the kernel)
  double* out;                          // [8]: MAE, MSE, RMSE, R2, mean cosine, Pearson mean, Pearson std, Pearson count
};
int metrics_grid(long long rows);
// Batch assembly by index (src/data/dataset.py:35-39 + DataLoader collate as one kernel) and in-place scaling of up to 8 arrays
struct GatherArgs {
  const float* a; const float* b; const long long* site; long long rows; int dim_a, dim_b;
  const long long* index; int n;
  float* out_a; float* out_b; long long* out_site;
};
cudaError_t launch_gather_rows(const GatherArgs& g, cudaStream_t s);
struct ScaleArgs { float* x[8]; long long n[8]; int count; const float* scale; };
cudaError_t launch_scale(const ScaleArgs& a, cudaStream_t s);
cudaError_t launch_metrics(const MetricsArgs& a, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Lock-step population step (vla_train_step_group): n independent models of one kind take their train steps as ONE launch
// per step of the sequence -- launch j covers launch j of every member.  The members' launches are recorded instead of
// issued (the launch_* functions below append to the thread's Recorder when one is installed), zipped, and issued through
// the *_multi kernels: block -> (member, block of the member's own grid) through a table in global memory; the member's
// argument structure (the same one its stand-alone launch takes) is read from global memory too.
// (optimize_hyperparameters.py:68-133 trials, vae_cross_modality_cv.py:113-283 folds: the reference runs them one by one.)
// ---------------------------------------------------------------------------------------------
enum RecKind : int { RK_GEMM = 0, RK_INGEST, RK_BN_ACT, RK_BN_BWD, RK_LATENT_FWD, RK_LATENT_BWD, RK_ADAMW, RK_LOSS, RK_COUNT };
struct alignas(16) MultiHdr { int block_begin; int gx; int aux; int pad; };   // gx: width of a 2-D grid; aux: rows per block
constexpr int MULTI_MAX_MEMBERS = 256;
struct RecOp {
  int kind = 0, variant = 0;       // RK_GEMM: variant = mode * 16 + instantiation
  int blocks = 0, gx = 0, aux = 0;
  std::string args;                // the launch's argument structure, byte for byte
};
struct Recorder {
  std::vector<RecOp> ops;
  bool unsupported = false;        // a launch that has no *_multi form was attempted
};
Recorder*& recorder();             // thread-local; nullptr = launch normally
// One merged launch: hdr[n] and args[n] (stride bytes apart) in device memory.
cudaError_t launch_multi(int kind, int variant, const MultiHdr* hdr, const void* args, int stride, int n, int total_blocks,
                         cudaStream_t s);
cudaError_t launch_gemm_multi(int variant, const MultiHdr* hdr, const GemmGroup* groups, int n, int total_blocks, int max_units,
                              cudaStream_t s);
int gemm_max_units(const GemmGroup& g);             // longest main loop of the group in ring slots (persistent-form heuristic)
int gemm_cluster_capacity();                         // 4-CTA clusters of the multicast GEMM kernel the device runs at once
int gemm_variant(const GemmGroup& g, int mode);     // which instantiation launch_gemm_group picks (-1: invalid flags)

}  // namespace vla
