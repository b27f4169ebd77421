/* libvla_b200 -- C ABI of the B200 (sm_100a) implementation of the vae-los-angeles hot path.
 *
 * The reference (marcin119a/vae-los-angeles) is pure Python on top of PyTorch; it has no FFI of its
 * own.  The boundary this library sits behind is therefore the reference's Python import surface
 * (SURVEY.md section 8b); each entry point below names the reference interface whose arithmetic it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless said otherwise;
 * all work is enqueued on `stream` (a cudaStream_t passed as void*) with no host synchronisation;
 * functions return 0 on success and a negative code on failure, with the message available from
 * vla_last_error() (thread-local).  fp32 tensors are dense row-major, `site` is int64.
 * There is no CPU fallback.
 */
#ifndef VLA_B200_H
#define VLA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vla_stream_t;
typedef struct vla_model vla_model_t;

enum { VLA_KIND_MULTIMODAL = 0, VLA_KIND_RNA2DNA = 1, VLA_KIND_DNA2RNA = 2, VLA_KIND_RNA2DNA_AE = 3, VLA_KIND_DNA2RNA_AE = 4 };
enum { VLA_OK = 0, VLA_ERR_INVALID = -1, VLA_ERR_CUDA = -2, VLA_ERR_STATE = -3 };
enum { VLA_TENSOR_PARAM = 0, VLA_TENSOR_BUFFER = 1, VLA_TENSOR_COUNTER = 2 };

/* Model geometry.  kind selects the encoder / decoder stacks:
 *   MULTIMODAL: MultiModalVAE.__init__ (src/models/vae.py:27-35)
 *   RNA2DNA / DNA2RNA: RNA2DNAVAE / DNA2RNAVAE.__init__ (src/models/directional_vae.py:19-23, 70-74)
 *   RNA2DNA_AE / DNA2RNA_AE: RNA2DNAAE / DNA2RNAAE.__init__ (src/models/directional_ae.py:17-35, 76-98): the same stacks with
 *     one head of width `latent` per encoder; forward (directional_ae.py:37-59, 100-123) returns the latent in `mu`
 *     (`logvar` is written as zeros), nothing is sampled and every KL term is zero (src/utils/ae_losses.py:8-39). */
typedef struct {
  int kind;
  int dim_a;     /* RNA features   (input_dim_a / rna_dim) */
  int dim_b;     /* DNA features   (input_dim_b / dna_dim) */
  int n_sites;
  int latent;
  int embed;     /* embed_dim, 32 by default in the reference */
} vla_config_t;

/* One state_dict entry.  PARAM offsets index the fp32 parameter arena,
 * BUFFER offsets the fp32 running-statistics arena, COUNTER offsets the int64 num_batches_tracked array. */
typedef struct {
  char name[96];
  int kind;
  long long offset;
  int ndim;
  int shape[2];
} vla_tensor_info_t;

const char* vla_last_error(void);
int vla_abi_version(void);

int vla_model_create(const vla_config_t* cfg, vla_model_t** out);
/* Same handle without any device state: only the layout queries below work (no GPU needed).  Used by host-side
 * tooling and the CPU tests of the data-parallel gradient packing. */
int vla_model_create_layout_only(const vla_config_t* cfg, vla_model_t** out);
void vla_model_destroy(vla_model_t* m);
/* Makes sure workspace for `batch` rows exists (also done lazily by the calls below). */
int vla_model_reserve(vla_model_t* m, int batch);

long long vla_param_count(const vla_model_t* m);    /* fp32 elements in the parameter arena */
long long vla_buffer_count(const vla_model_t* m);   /* fp32 elements in the running-statistics arena */
int vla_counter_count(const vla_model_t* m);        /* number of BatchNorm layers */
int vla_num_tensors(const vla_model_t* m);
int vla_tensor_info(const vla_model_t* m, int index, vla_tensor_info_t* out);

/* Forward pass.  Replaces MultiModalVAE.forward (src/models/vae.py:37-79), RNA2DNAVAE.forward /
 * DNA2RNAVAE.forward (src/models/directional_vae.py:25-60, 76-111) including EncoderA/B/C.forward
 * (src/models/encoders.py:21-23, 43-46, 57-61), reparameterize (src/models/vae.py:11-15) and
 * DecoderA/B/C.forward (src/models/decoders.py:18-19, 35-36, 49-50).
 * A null x_a / x_b / site means "modality absent"; at least one encoder input must be present.
 * train != 0: BatchNorm uses batch statistics and updates the running statistics, dropout is active, and
 * activations are kept for vla_backward.  eps == NULL draws epsilon from Philox(seed, offset).
 * keep_masks: optional array with one entry per dropout layer (model order), each NULL or a uint8 keep
 * mask [batch, width]; used by the parity tests to replay the reference's dropout decisions. */
typedef struct {
  const float* params;
  float* buffers;
  long long* counters;
  const float* x_a;
  const float* x_b;
  const long long* site;
  int batch;
  int train;
  int refresh_shadows;          /* != 0: re-derive the bf16 MMA copies of the weights from `params` first */
  const float* eps;
  const unsigned char* const* keep_masks;
  unsigned long long seed;
  unsigned long long offset;
  float* recon_a;               /* [batch, dim_a]   (kinds with a type-A decoder) */
  float* recon_b;               /* [batch, dim_b]   (type-B decoder, sigmoid output) */
  float* recon_c;               /* [batch, n_sites] (type-C decoder, logits) */
  float* mu;                    /* [batch, latent] */
  float* logvar;                /* [batch, latent] */
} vla_forward_args_t;
int vla_forward(vla_model_t* m, const vla_forward_args_t* a, vla_stream_t stream);

/* Re-derives the bf16 tensor-core copies of the weights from the fp32 parameter arena (needed after the
 * parameters were changed by anything other than vla_adamw / vla_train_step). */
int vla_refresh_shadows(vla_model_t* m, const float* params, vla_stream_t stream);

/* Backward of the last train-mode vla_forward on this handle: what autograd does for the reference
 * after loss.backward() (train_rna2dna.py:95).  g_* are dL/d(output) (NULL = zero); `recon_b` must be the
 * sigmoid output returned by that forward when g_recon_b is given.  Writes every parameter gradient into
 * grads[param_count] (gradients of absent stacks are left zero). */
typedef struct {
  const float* params;
  const float* g_recon_a;
  const float* g_recon_b;
  const float* g_recon_c;
  const float* g_mu;
  const float* g_logvar;
  const float* recon_b;
  float* grads;
} vla_backward_args_t;
int vla_backward(vla_model_t* m, const vla_backward_args_t* a, vla_stream_t stream);

/* Loss values and gradients w.r.t. the model outputs.  Replaces vae_loss (src/utils/losses.py:8-46),
 * rna2dna_loss / dna2rna_loss (src/utils/directional_losses.py:8-30, 33-55): any term whose recon pointer
 * is NULL is skipped.  out[4] = {total, recon, class, kld}.  Gradient outputs are optional.
 * workspace: vla_loss_workspace_bytes() bytes, zero-initialised once by the caller. */
typedef struct {
  const float* recon_a; const float* a; int dim_a;
  const float* recon_b; const float* b; int dim_b;
  const float* recon_c; const long long* site; const float* class_weights; int n_sites;
  const float* mu; const float* logvar; int latent;
  int batch;
  float beta, gamma;
  float* g_recon_a; float* g_recon_b; float* g_recon_c; float* g_mu; float* g_logvar;
  float* out;
  void* workspace;
} vla_loss_args_t;
long long vla_loss_workspace_bytes(int batch, int dim_a, int dim_b, int n_sites, int latent);
int vla_loss(const vla_loss_args_t* a, vla_stream_t stream);

/* Reconstruction metrics of an evaluation pass in one streaming kernel.  Replaces compute_metrics
 * (compare_directional_imputation.py:167-210), i.e. sklearn's mean_absolute_error / mean_squared_error / r2_score on the
 * flattened arrays, the diagonal of sklearn's cosine_similarity (the reference builds the full N x N matrix) and
 * scipy.stats.pearsonr per sample with NaN rows skipped.  out[8] (DEVICE doubles) = {MAE, MSE, RMSE, R2, mean cosine
 * similarity, Pearson mean, Pearson population std, number of samples with a defined Pearson r}; cosine / pearson are
 * optional per-sample outputs [rows] (pearson: NaN where a row is constant).  workspace: vla_metrics_workspace_bytes(rows)
 * bytes, zero-initialised once by the caller. */
typedef struct {
  const float* y_true; const float* y_pred;    /* dense fp32 [rows, dim] */
  long long rows; int dim;
  float* cosine; float* pearson;
  double* out;
  void* workspace;
} vla_metrics_args_t;
/* Batch assembly on the device: rows index[0..n) (int64, device) of the three dataset arrays (fp32 [rows, dim_a], fp32
 * [rows, dim_b], int64 [rows]) into contiguous batch buffers -- what DataLoader's sampler + default collate do per batch
 * (reference src/data/dataset.py:35-39 __getitem__, train_rna2dna.py:57-67 DataLoader(shuffle=True)) as one kernel.
 * An index outside [0, rows) is an error reported through the return value of the next call that synchronises (the kernel
 * clamps it, never reads out of bounds). */
int vla_gather_rows(const float* a, int dim_a, const float* b, int dim_b, const long long* site, long long rows,
                    const long long* index, int n, float* out_a, float* out_b, long long* out_site, vla_stream_t stream);
/* x[i] *= scale[0] for n_tensors fp32 device arrays in one launch (scale: one fp32 on the device): the upstream d/d(total) of
 * the fused loss applied to its stored gradients (loss.backward() at train_rna2dna.py:95 when the script scales the loss). */
int vla_scale_inplace(void* const* tensors, const long long* counts, int n_tensors, const float* scale, vla_stream_t stream);
long long vla_metrics_workspace_bytes(long long rows);
int vla_recon_metrics(const vla_metrics_args_t* a, vla_stream_t stream);

/* One fused multi-tensor AdamW step over the flat arena (torch.optim.AdamW semantics, the optimizer built at
 * train_rna2dna.py:185-189 and stepped at :94-96); also refreshes the bf16 weight copies. */
typedef struct {
  float* params; const float* grads; float* exp_avg; float* exp_avg_sq;
  float lr, beta1, beta2, eps, weight_decay;
  int step;                     /* 1-based */
} vla_adamw_args_t;
int vla_adamw(vla_model_t* m, const vla_adamw_args_t* a, vla_stream_t stream);

/* Whole training step: forward + loss + backward + AdamW, the loop body at train_rna2dna.py:82-99 /
 * optimize_hyperparameters.py:104-113, with the learning rate, weight decay, KL weight, gamma and the step
 * count held on the device (vla_set_hyper) so that a captured CUDA graph of this call can be replayed.
 * Targets are the same tensors as the inputs (a, b, site): x_a, x_b and site are all required (inputs and/or
 * targets of every kind).  loss_out[4] = {total, recon, class, kld}. */
typedef struct {
  float* params; float* grads; float* exp_avg; float* exp_avg_sq;
  float* buffers; long long* counters;
  const float* x_a; const float* x_b; const long long* site;
  const float* class_weights;
  int batch;
  long long dataset_rows;       /* rows behind x_a / x_b / site; > batch: step t trains on rows
                                   [(i mod n) * batch, +batch), i = device-side batch index, n = dataset_rows / batch */
  const float* eps;
  const unsigned char* const* keep_masks;
  unsigned long long seed;
  float beta1, beta2, adam_eps;
  float* recon_a; float* recon_b; float* recon_c; float* mu; float* logvar;   /* optional outputs */
  float* loss_out;
  int phases;                   /* 0 or 3: whole step; 1: forward + loss + backward only (gradients left in `grads`, e.g. for a
                                   data-parallel all-reduce); 2: AdamW only (consumes and clears `grads`) */
  void* dp;                     /* NULL, or a connected vla_dp_t*: data-parallel step -- between backward and AdamW the
                                   gradients (and the 4 loss scalars behind them) are summed over the ranks through NVLink
                                   peer memory.  grads must be vla_dp_grads(dp), loss_out its last 4 floats; the summed
                                   losses land in vla_dp_losses(dp).  Every rank must call in lockstep. */
  int sync_bn;                  /* with dp: 1 = BatchNorm statistics over the GLOBAL batch (forward mean / variance and the two
                                   backward sums of every BatchNorm layer, encoders.py:14,32,36, are all-reduced over the ranks
                                   through peer memory): the step then equals the reference on the concatenated batch.
                                   0 = per-shard statistics (the DistributedDataParallel convention; default). */
} vla_train_args_t;
int vla_train_step(vla_model_t* m, const vla_train_args_t* a, vla_stream_t stream);
int vla_set_hyper(vla_model_t* m, float lr, float weight_decay, float beta_kl, float gamma, vla_stream_t stream);
/* Resets the device-side step counter (and beta1^t, beta2^t for the bias corrections) and the resident-batch index. */
int vla_set_step(vla_model_t* m, int completed_steps, int batch_index, float beta1, float beta2, vla_stream_t stream);

/* Lock-step train step of a POPULATION: n independent models of the same kind (any latent / embedding widths, batch sizes,
 * datasets, hyper-parameters), each with its own vla_train_args_t exactly as for vla_train_step.  Launch j of the step is
 * issued ONCE for all members (grouped GEMMs: every member's tiles in one grid; the element-wise launches likewise), so
 * the population's step costs the launch latency of one model's.  Replaces the sequential trial / fold loops of the
 * reference (optimize_hyperparameters.py:68-133 `objective` called per trial by study.optimize, :101-113 its batch loop;
 * vae_cross_modality_cv.py:314-344 folds, :136-158 and :219-240 the batch loops of train_vae / train_ae).
 * Results are bit-identical to n separate vla_train_step calls.  The merged launch tables are cached on models[0] (keyed
 * by the argument image): the first call with a new combination must happen outside a stream capture.  dp must be NULL;
 * phases as in vla_train_step. */
int vla_train_step_group(vla_model_t* const* models, const vla_train_args_t* const* args, int n, vla_stream_t stream);
int vla_group_cached_plans(vla_model_t* lead);


/* Data-parallel gradient exchange over NVLink / NVSwitch peer memory (one process per GPU, one node).  The reference has
 * no distributed code; this replaces what DistributedDataParallel's all-reduce would do around
 * loss.backward() / optimizer.step() (train_rna2dna.py:94-96), with SUM semantics (src/utils/losses.py:31-42 are
 * reduction='sum').  Each rank creates its buffers ([param_count + 4] floats: gradient arena | 4 loss scalars), the ranks
 * exchange the 64-byte IPC handles out of band (HOST memory, e.g. torch.distributed.all_gather_object) and connect; after
 * that vla_train_step(dp = handle) runs the collective as one kernel of the step (csrc/dp_exchange.cu).
 * Destroy only after every rank has finished its last step (barrier first). */
typedef struct vla_dp vla_dp_t;
int vla_dp_create(int world, int rank, long long n_floats, vla_dp_t** out);
int vla_dp_ipc_handle(vla_dp_t* d, void* out64);             /* HOST pointer, 64 bytes */
int vla_dp_connect(vla_dp_t* d, const void* handles);        /* HOST pointer, world x 64 bytes in rank order */
void* vla_dp_grads(vla_dp_t* d);                             /* device: local gradients [n_floats] */
void* vla_dp_losses(vla_dp_t* d);                            /* device: float[4], the loss scalars summed over the ranks */
/* %globaltimer stamps (ns) of block 0 in the last step's two exchange launches on this rank, HOST out8[8]: [0..2] the
 * encoder part (main stream), [4..6] the decoder part (side stream, overlapping the encoder backward): kernel entry,
 * contributions pushed to the other ranks, own shard slice reduced and pushed.  Synchronises the device. */
int vla_dp_trace(vla_dp_t* d, unsigned long long* out8);
/* Teardown in two steps: every rank calls vla_dp_disconnect (unmaps the peers' buffers), the ranks meet in a barrier, then every
 * rank calls vla_dp_destroy (frees its own exported buffer: no importer has it mapped any more). */
int vla_dp_disconnect(vla_dp_t* d);
void vla_dp_destroy(vla_dp_t* d);

/* Chain kernel.  The row-local stretches of a call -- consecutive launches in which a 128-row block of the batch depends
 * only on the same rows of the previous launch, i.e. everything between two BatchNorm-statistics boundaries; in eval mode
 * the whole forward -- run as ONE launch each (csrc/chain_kernel.cu): a 4-CTA cluster per row block walks the phases with
 * cluster barriers instead of kernel boundaries.  Used by vla_train_step and by vla_forward / vla_backward from batch 1024.
 * The first call for a new argument set builds the plan (device allocation + upload, not capturable); later calls and
 * CUDA-graph replays only launch.  VLA_CHAIN=0 in the environment issues the same phases as separate launches.
 * Timeline: when enabled, every CTA writes eight %globaltimer stamps per phase: 0 phase start, 1 phase end (thread 0, in front of the
 * cluster barrier), 2 barrier passed, 3 first operands landed, 4 all MMAs issued, 5 accumulator ready (element-wise phases: start),
 * 6 first epilogue warp done, 7 last epilogue warp done. */
int vla_chain_timeline(vla_model_t* m, int enable);
/* Row-chain kernel (csrc/rowchain.cu: the row-local middle of a directional model's step, activations on chip).  With
 * VLA_RC_TIMELINE=1 in the environment every CTA stamps %globaltimer per op: out[148][32][4] = GEMM issue start, all MMAs
 * issued, accumulator ready (element-wise ops: start), epilogue done; kinds / subs[32] name the ops.  Returns the op count. */
int vla_rowchain_timeline(vla_model_t* m, unsigned long long* out, int* kinds, int* subs);
int vla_chain_count(vla_model_t* m);                     /* chain launches made by the last call on this handle */
int vla_chain_cached_plans(vla_model_t* m);              /* plan images built so far on this handle (a repeated call must not add one) */
int vla_chain_info(vla_model_t* m, int which, char* name48, int* n_phases, int* n_ctas, double* flops, double* bytes);
int vla_chain_phase_name(vla_model_t* m, int which, int phase, char* name48);
int vla_chain_timeline_read(vla_model_t* m, int which, unsigned long long* out);   /* out[n_ctas][24][8]; returns n_ctas */

/* A caller that captures calls on this handle into CUDA graphs (vla_b200.Trainer) pins the handle (+1) for as long as the
 * graphs live (-1 afterwards): while pinned, a call that would have to grow -- i.e. free and reallocate -- the workspace the
 * graphs reference fails with VLA_ERR_STATE instead of leaving them replaying against freed memory. */
int vla_model_pin(vla_model_t* m, int delta);

/* Per-launch device timing (CUDA events on `stream`, recorded around every kernel launch the library makes between
 * vla_profile_begin and vla_profile_collect).  flops / bytes are the ALGORITHMIC work of the launch.
 * vla_profile_collect synchronises the stream's events and returns how many entries it wrote. */
typedef struct {
  char name[48];
  float ms;
  double flops;
  double bytes;
} vla_prof_entry_t;
int vla_profile_begin(vla_model_t* m);
int vla_profile_collect(vla_model_t* m, vla_prof_entry_t* out, int max_entries);
/* Like vla_profile_collect but keeps the events (they may be nodes of a captured graph that is replayed again). */
int vla_profile_read(vla_model_t* m, vla_prof_entry_t* out, int max_entries);
/* Stops recording without touching the events already recorded (call after capturing a profiled graph). */
int vla_profile_pause(vla_model_t* m);

/* Test hook: C[M,N] = A[M,K] * B[N,K]^T (mode 0) or C[M,N] = A[K,M]^T * B[K,N] (mode 1) on the tcgen05 path.
 * A, B bf16 (uint16 storage) with element pitches lda / ldb (multiples of 8), C fp32 dense, zero-filled by
 * the caller in mode 1 (split-K accumulation).  bias_grad (mode 1, optional) receives sum_k A[k, m]. */
int vla_test_gemm(int mode, const void* A, int lda, const void* B, int ldb, float* C, int M, int N, int K,
                  int bn, int k_splits, float* bias_grad, vla_stream_t stream);

/* Test hook: device address and row pitch (elements) of a workspace buffer left by the last forward / train step on this
 * handle.  what 0: the epsilon of reparameterize (src/models/vae.py:13, fp32 [rows, latent]); what 1: the bf16 output of
 * Dropout(ReLU(BatchNorm(.))) of encoder i, layer j (src/models/encoders.py:14-16, 32-38). */
int vla_test_workspace(vla_model_t* m, int what, int i, int j, void** ptr, int* ld);

/* Test hook: device buffer [tiles][8] of %globaltimer stamps written by the following vla_test_gemm calls
 * (kernel entry, dependency resolved, setup done, first operands landed, MMAs issued, accumulator ready, epilogue done). */
int vla_test_set_timeline(unsigned long long* dbg);
int vla_test_set_flags(int flags);   /* debugging switches of the GEMM test hook (profiles/cta_turnaround.py) */

#ifdef __cplusplus
}
#endif
#endif
