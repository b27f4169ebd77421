// Host side of libvla_b200: parameter-arena layout, workspace, launch sequencing of the forward,
// backward and fused train step, and the extern "C" boundary declared in include/vla_b200.h.
#include "../../include/vla_b200.h"
#include "vla_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

using namespace vla;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(expr)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return fail(VLA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
  } while (0)

constexpr int FEATS_FWD_PLAIN_HOST = GF_BIAS | GF_RELU | GF_OUT_F32 | GF_OUT_BF16;     // must match gemm_tile.cuh
constexpr int FEATS_DGRAD_PLAIN_HOST = GF_MASK | GF_OUT_F32 | GF_OUT_BF16;
constexpr int FEATS_FWD_LOSS_BCE_HOST = GF_BIAS | GF_SIGMOID | GF_OUT_BF16 | GF_LOSS;
constexpr int FEATS_FWD_LOSS_MSE_HOST = GF_BIAS | GF_OUT_BF16 | GF_LOSS;
inline int pad8(int x) { return (x + 7) & ~7; }
inline int pad64(int x) { return (x + 63) & ~63; }
// Row pitch / lo offset of a bf16 GEMM operand of logical width w.  Split layout (DESIGN.md "Precision"):
// [hi: 0 .. w | zeros up to pad64(w) | lo: pad64(w) .. pad64(w) + w | zeros]; the pad columns are never written (the
// workspace and the weight copies are zero-filled once), so whole 64-column k-blocks can be read from either half.
inline int op_ld(bool split, int w) { return split ? 2 * pad64(w) : pad8(w); }
inline int op_lo(bool split, int w) { return split ? pad64(w) : 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

struct Lin {
  int out = 0, in = 0;
  long long w_off = -1, b_off = -1;
  long long sh_off = -1; int sh_ld = 0;      // bf16 copy [out, sh_ld]: K-major B operand of the forward GEMM and
                                             // MN-major B operand of the data-gradient GEMM
  int sh_lo = 0;                             // > 0: split copy, lo part sh_lo elements further along the row
};
struct Bn { int n = 0; long long g_off = -1, b_off = -1, rm_off = -1, rv_off = -1; int counter = -1; };
struct Enc {
  char type; std::string prefix; int in_dim; int slot;   // slot: 0 = a, 1 = b, 2 = site
  std::vector<Lin> fc; std::vector<Bn> bn; Lin heads; long long emb_off = -1;
  int first_drop = 0;                                     // index of this encoder's first dropout layer
};
struct Dec {
  char type; std::string prefix; int out_dim;
  int cat_off = 0, cat_w = 0;                             // slice of the fused first layer
  std::vector<Lin> rest;
};

struct EncWS {
  bf16* x = nullptr; int ldx = 0, x_lo = 0;               // bf16 input (dense) / gathered embedding (site)
  bf16* onehot = nullptr; int ld_onehot = 0;
  bf16* g_x = nullptr; int ld_gx = 0;                     // site: gradient w.r.t. the gathered embedding
  std::vector<float*> pre, stats, bstats, mean, rstd;
  std::vector<bf16*> act, gy, gpre;
  std::vector<unsigned short*> bits;                      // row-chain kernel: (activation > 0) per element, [rows][n / 16] halfwords
  std::vector<int> ld_act, act_lo;
  float* ml = nullptr;
};
struct DecWS {
  std::vector<bf16*> act, gact;                           // hidden activations after the fused first layer
  std::vector<unsigned short*> bits;                      // row-chain kernel: ReLU masks of the hidden activations
  std::vector<int> ld_act, act_lo;
  bf16* g_out = nullptr; int ld_gout = 0;                 // bf16 gradient w.r.t. the last layer's pre-activation
  float* recon = nullptr;                                 // fp32 output when the caller passes none
};

struct PendingB { const void* base; uint64_t inner, outer, pitch; int mode; int K; bool fixed; };

struct TmapKey {
  uintptr_t base; uint64_t inner, outer, pitch; uint32_t box_outer;
  bool operator<(const TmapKey& o) const {
    return std::tie(base, inner, outer, pitch, box_outer) < std::tie(o.base, o.inner, o.outer, o.pitch, o.box_outer);
  }
};

}  // namespace

struct ProfEntry { char name[48]; cudaEvent_t e0, e1; double flops, bytes; };

// One recorded launch of a row-local stretch (chain_kernel.cu): its argument struct by value.
struct ChainOp {
  int kind;                               // ChainKind
  int gemm_mode;                          // GEMM ops: 0 NT / 2 NN
  std::vector<char> args;
  std::string name;
  double flops, bytes;
};
// Device image of one stretch, built on its first (eager) execution and reused by every later call / graph replay with
// the same arguments.  Never freed before the model (captured graphs hold the address).
struct ChainPlanCached {
  std::vector<char> key;                  // the host image [ChainPlan | argument structs] (dbg pointer cleared)
  char* dev = nullptr;
  size_t dbg_off = 0;
  int n_clusters = 0, n_phases = 0;
  double flops = 0, bytes = 0;
  std::string name;
  std::vector<std::string> phase_names;
};

// One cached lock-step step of a population: for launch j the table [MultiHdr n | argument structures n x stride].  Keyed
// by the host image (a replayed or re-captured step presents the same image).  Never freed before the leading model.
struct GroupPlanCached {
  std::string key;
  char* dev = nullptr;
  struct Launch { int kind, variant, stride, n, total_blocks, max_units; size_t hdr_off, args_off; };
  std::vector<Launch> launches;
};

struct vla_model {
  bool layout_only = false;
  bool prof_on = false;
  std::vector<ProfEntry> prof;
  vla_config_t cfg{};
  int L = 0, E = 0, S = 0;
  bool split = true;                 // split-bf16 operands for every forward GEMM that feeds a ReLU (VLA_SPLIT=0: plain bf16)
  bool ae = false; int HW = 0;       // autoencoder kinds: one head of width L per encoder (HW = L), else fused mu | logvar (HW = 2L)
  std::vector<Enc> encs;
  std::vector<Dec> decs;
  Lin cat;
  int n_drop = 0;
  long long n_params = 0, n_buffers = 0, n_shadow = 0;
  int n_bn = 0;
  std::vector<vla_tensor_info_t> infos;
  std::vector<AdamChunk> chunks_h;
  // device-resident, batch independent
  bf16* shadow = nullptr;
  AdamChunk* chunks_d = nullptr;
  DynParams* dyn = nullptr;
  float* loss_partials = nullptr; unsigned int* loss_counter = nullptr; float* loss_out = nullptr;
  float* eloss_partials = nullptr;   // partial sums of the loss-fused decoder epilogues
  // workspace sized for `cap` rows
  int cap = 0;
  char* ws = nullptr;
  std::vector<EncWS> ews;
  std::vector<DecWS> dws;
  float *mu = nullptr, *logvar = nullptr, *eps = nullptr, *kl_partials = nullptr, *gz = nullptr;
  bf16 *z = nullptr, *gml = nullptr, *d0 = nullptr, *g_d0 = nullptr;
  unsigned short* d0_bits = nullptr;
  int ldz = 0, z_lo = 0, ldgml = 0, ld_d0 = 0, d0_lo = 0;
  std::map<TmapKey, CUtensorMap> tmaps;
  // what the last forward left in the workspace
  bool saved = false; int saved_batch = 0, saved_present = 0, saved_train = 0, kl_grid = 0;
  unsigned long long generation = 0;
  // chain kernel: row-local stretches of a call as one launch each
  bool chain_on = false;                  // set by the entry point for the duration of a call that may chain
  std::vector<ChainOp> seg;               // the open stretch
  std::vector<ChainPlanCached*> plans;
  ChainPlanCached* last_plans[8] = {}; int n_last_plans = 0;   // plans launched by the last call (timeline)
  bool chain_dbg = false;
  unsigned long long* rc_dbg = nullptr;   // row-chain timeline buffer [148][RC_MAX_OPS][4] (VLA_RC_TIMELINE=1)
  RcPlan* rc_last = nullptr;              // host copy of the last row-chain plan (timeline labels)
  int chain_clusters = 0;                 // 4-CTA clusters of the chain kernel the device runs at once
  int pinned = 0;                         // > 0: captured graphs reference the workspace / plans (vla_model_pin)
  bool hb_used = false;                   // the last forward ran the head-block kernel (the backward must mirror it)
  // Side branch of a train step (single GPU): the decoder weight gradients and the decoder part of AdamW need nothing from
  // the encoder backward; they run on a low-priority stream beside it (fork after the decoder data gradients, join at the
  // end of the step; inside a capture the branch becomes a parallel arm of the graph).
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int dec_chunk0 = -1;                    // first AdamW chunk of the decoder parameters (arena order: encoders | decoders)
  bool side_busy = false;                 // the branch is open: the caller must join it
  AdamHints* adam_hints = nullptr;        // chunk offsets as a kernel parameter (built on first use)
  LossTail deferred_tail{}; bool deferred_tail_valid = false;   // final loss reduction handed to the step's AdamW launch
  // lock-step population steps led by this model (vla_train_step_group): device images of the merged launch tables
  std::vector<struct GroupPlanCached*> group_plans;
};

// Peer-memory gradient exchange of one data-parallel trainer (dp_exchange.cu).  Two allocations per rank:
//   local (never exported): [G n floats | summed losses | trace] -- only this rank touches them;
//   base  (exported / mapped with CUDA IPC): [RECV 2 x world x per2 framed words | RSUM n / 2 framed words].
struct vla_dp {
  int world = 1, rank = 0;
  long long n = 0;                 // floats per buffer (multiple of 4)
  long long per2 = 0;              // float2s per shard
  char* base = nullptr;            // exported allocation (RECV | RSUM)
  char* local = nullptr;           // private allocation (G | sums | trace)
  size_t off_recv = 0, off_rsum = 0, off_small = 0, off_sums = 0, off_trace = 0, bytes = 0, local_bytes = 0;
  int small_next = 0;              // SyncBN: next (BatchNorm layer, direction) region of the step being issued
  char* peer[DP_MAX_WORLD] = {};   // every rank's allocation as mapped here (own: base)
  bool connected = false;
  cudaStream_t side = nullptr;     // the early (decoder) part of the exchange runs here, beside the encoder backward
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

// ---------------------------------------------------------------------------------------------
// Optional per-launch timing
// ---------------------------------------------------------------------------------------------
struct ProfScope {
  vla_model* m; cudaStream_t st; int idx = -1;
  ProfScope(vla_model* m_, cudaStream_t st_, const char* name, double flops, double bytes) : m(m_), st(st_) {
    if (!m->prof_on) return;
    ProfEntry e{};
    snprintf(e.name, sizeof(e.name), "%s", name);
    e.flops = flops; e.bytes = bytes;
    if (cudaEventCreate(&e.e0) != cudaSuccess || cudaEventCreate(&e.e1) != cudaSuccess) return;
    // while the stream is being captured the records must be EXTERNAL event nodes, so that they fire on every replay
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    flags = cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
    cudaEventRecordWithFlags(e.e0, st, flags);
    m->prof.push_back(e);
    idx = static_cast<int>(m->prof.size()) - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecordWithFlags(m->prof[idx].e1, st, flags); }
  unsigned flags = cudaEventRecordDefault;
};

// ---------------------------------------------------------------------------------------------
// Chain recording
// ---------------------------------------------------------------------------------------------
void free_plans(vla_model* m) {
  for (ChainPlanCached* c : m->plans) { cudaFree(c->dev); delete c; }
  m->plans.clear();
  m->n_last_plans = 0;
}
int chain_flush(vla_model* m, cudaStream_t st);
int chain_add(vla_model* m, cudaStream_t st, int kind, int gemm_mode, const void* args, size_t size, const char* name,
              double flops, double bytes, bool needs_all);
int finalize_group(vla_model* m, GemmGroup& g, int mode);
int run_stats_allreduce(vla_model* m, vla_dp* dp, float* partials, int m_tiles, int n, const char* name, cudaStream_t st);
void gemm_work(const GemmGroup& g, double* flops, double* bytes) {
  *flops = 0; *bytes = 0;
  for (int i = 0; i < g.nprob; ++i) {
    const GemmProblem& p = g.p[i];
    *flops += 2.0 * p.M * p.N * p.K;
    const double out_b = (p.flags & (GF_OUT_F32 | GF_RED)) ? 4.0 : 0.0;
    // operands (hi + lo copies when split), results, and the fp32 targets a loss-fused tile reads (MSE / BCE)
    const double tgt_b = ((p.flags & GF_LOSS) && p.loss_kind != LOSS_CE) ? 4.0 : 0.0;
    const double opw = p.a_lo > 0 ? 4.0 : 2.0;
    *bytes += opw * (static_cast<double>(p.M) * p.K + static_cast<double>(p.N) * p.K) +
              (out_b + tgt_b + ((p.flags & GF_OUT_BF16) ? (p.out_lo > 0 ? 4.0 : 2.0) : 0.0)) * p.M * p.N;
    if (p.a_lo > 0) *flops += 4.0 * p.M * p.N * p.K;      // three MMA passes
    if (p.flags & GF_LATBWD) *bytes += (p.loss_kind ? 2.0 : 16.0) * p.M * p.N;   // mu, logvar, eps in; d(mu | logvar) bf16 out
  }
}
int timed_gemm(vla_model* m, GemmGroup& g, int mode, const char* name, cudaStream_t st, bool finalized = false) {
  if (mode != 1 && !finalized) { int rc0 = finalize_group(m, g, mode); if (rc0) return rc0; }
  for (int i = 0; i < g.nprob; ++i) {
    const GemmProblem& p = g.p[i];
    if ((p.flags & GF_MASK) && ((p.N & 31) || (p.ld_mask & 7) || (reinterpret_cast<uintptr_t>(p.mask_src) & 15)))
      return fail(VLA_ERR_STATE, std::string(name) + ": masked epilogue needs N % 32 == 0 and 16-byte aligned rows");
    if ((p.flags & GF_BNSTATS) && ((p.N & 31) || (p.ld_pre & 3) || (reinterpret_cast<uintptr_t>(p.pre) & 15)))
      return fail(VLA_ERR_STATE, std::string(name) + ": BatchNorm-statistics epilogue needs N % 32 == 0 and 16-byte aligned rows");
  }
  double fl, by; gemm_work(g, &fl, &by);
  if (m->chain_on && mode != 1) {
    // which instantiation of the tile body (the same choice launch_gemm_group makes)
    int used = 0, kinds = 0;
    for (int i = 0; i < g.nprob; ++i) {
      used |= g.p[i].flags;
      kinds |= (g.p[i].flags & GF_LOSS) ? 1 << g.p[i].loss_kind : 1 << LOSS_NONE;
    }
    int kind;
    if (mode == 0) {
      if (used & GF_LOSS) {
        kind = CK_GEMM_NT_LOSS;
        if (kinds == 1 << LOSS_BCE && !(used & ~FEATS_FWD_LOSS_BCE_HOST)) kind = CK_GEMM_NT_LOSS_BCE;
        if (kinds == 1 << LOSS_MSE && !(used & ~FEATS_FWD_LOSS_MSE_HOST)) kind = CK_GEMM_NT_LOSS_MSE;
      } else {
        kind = (used & ~FEATS_FWD_PLAIN_HOST) ? CK_GEMM_NT_FULL : CK_GEMM_NT_PLAIN;
      }
    } else {
      kind = (used & ~FEATS_DGRAD_PLAIN_HOST) ? CK_GEMM_NN_FULL : CK_GEMM_NN_PLAIN;
    }
    return chain_add(m, st, kind, mode, &g, sizeof(g), name, fl, by, false);
  }
  { int rcf = chain_flush(m, st); if (rcf) return rcf; }
  ProfScope ps(m, st, name, fl, by);
  cudaError_t e = launch_gemm_group(g, mode, st);
  if (e != cudaSuccess) return fail(VLA_ERR_CUDA, std::string("gemm launch (") + name + "): " + cudaGetErrorString(e));
  return VLA_OK;
}

// ---------------------------------------------------------------------------------------------
// Layout
// ---------------------------------------------------------------------------------------------
struct ArenaBuilder {
  vla_model* m;
  long long p = 0, b = 0, sh = 0;
  long long take_p(long long n) { long long o = p; p += (n + 3) & ~3LL; return o; }
  long long take_b(long long n) { long long o = b; b += (n + 3) & ~3LL; return o; }
  long long take_sh(long long n) { long long o = sh; sh += (n + 63) & ~63LL; return o; }   // 128-byte aligned
  void info(const std::string& name, int kind, long long off, int d0, int d1 = -1) {
    vla_tensor_info_t t{};
    snprintf(t.name, sizeof(t.name), "%s", name.c_str());
    t.kind = kind; t.offset = off;
    t.ndim = d0 < 0 ? 0 : (d1 < 0 ? 1 : 2);
    t.shape[0] = d0 < 0 ? 0 : d0; t.shape[1] = d1 < 0 ? 0 : d1;
    m->infos.push_back(t);
  }
  void seg(long long off, int rows, int cols, long long sh_off, int sh_ld, int sh_lo = 0) {
    const long long n = static_cast<long long>(rows) * cols;
    for (long long st = 0; st < n; st += ADAM_CHUNK) {
      AdamChunk c{};
      c.offset = off + st; c.shadow_off = sh_off; c.n = static_cast<int>(std::min<long long>(ADAM_CHUNK, n - st));
      c.first = static_cast<int>(st); c.cols = cols; c.ld_shadow = sh_ld; c.sh_lo = sh_lo;
      m->chunks_h.push_back(c);
    }
  }
  // A Linear whose weight rows may be exposed under several state_dict names (fused groups).
  // split: this layer's forward GEMM runs on split-bf16 operands (its result reaches a ReLU)
  Lin linear(int out, int in, bool split) {
    Lin l; l.out = out; l.in = in;
    l.w_off = take_p(static_cast<long long>(out) * in);
    l.b_off = take_p(out);
    l.sh_ld = op_ld(split, in); l.sh_lo = op_lo(split, in);
    l.sh_off = take_sh(static_cast<long long>(out) * l.sh_ld);
    seg(l.w_off, out, in, l.sh_off, l.sh_ld, l.sh_lo);
    seg(l.b_off, out, 1, -1, 0);
    return l;
  }
};

int build_layout(vla_model* m) {
  const vla_config_t& c = m->cfg;
  m->L = c.latent; m->E = c.embed; m->S = c.n_sites;
  struct StackSpec { const char* prefix; char type; };
  std::vector<StackSpec> es, ds;
  if (c.kind == VLA_KIND_MULTIMODAL) {
    es = {{"encoder_a", 'A'}, {"encoder_b", 'B'}, {"encoder_c", 'C'}};
    ds = {{"decoder_a", 'A'}, {"decoder_b", 'B'}, {"decoder_c", 'C'}};
  } else if (c.kind == VLA_KIND_RNA2DNA) {
    es = {{"encoder_rna", 'A'}, {"encoder_site", 'C'}};
    ds = {{"decoder_dna", 'B'}};
  } else if (c.kind == VLA_KIND_DNA2RNA) {
    es = {{"encoder_dna", 'B'}, {"encoder_site", 'C'}};
    ds = {{"decoder_rna", 'A'}};
  } else if (c.kind == VLA_KIND_RNA2DNA_AE) {     // RNA2DNAAE.__init__, src/models/directional_ae.py:17-35
    es = {{"encoder_rna", 'A'}, {"site", 'C'}};
    ds = {{"decoder_dna", 'B'}};
    m->ae = true;
  } else if (c.kind == VLA_KIND_DNA2RNA_AE) {     // DNA2RNAAE.__init__, src/models/directional_ae.py:76-98
    es = {{"encoder_dna", 'B'}, {"site", 'C'}};
    ds = {{"decoder_rna", 'A'}};
    m->ae = true;
  } else {
    return fail(VLA_ERR_INVALID, "unknown model kind");
  }
  ArenaBuilder ab{m};
  const int L = m->L;
  const bool ae = m->ae;
  m->HW = ae ? L : 2 * L;
  for (const auto& s : es) {
    Enc e; e.type = s.type; e.prefix = s.prefix;
    e.slot = s.type == 'A' ? 0 : (s.type == 'B' ? 1 : 2);
    e.first_drop = m->n_drop;
    int last;
    if (s.type == 'C') {
      e.in_dim = c.n_sites;
      e.emb_off = ab.take_p(static_cast<long long>(c.n_sites) * c.embed);
      ab.seg(e.emb_off, c.n_sites, c.embed, -1, 0);
      ab.info(ae ? std::string("site_embedding.weight") : e.prefix + ".embedding.weight", VLA_TENSOR_PARAM, e.emb_off, c.n_sites, c.embed);
      last = c.embed;
    } else {
      e.in_dim = s.type == 'A' ? c.dim_a : c.dim_b;
      std::vector<int> hidden = s.type == 'A' ? std::vector<int>{128} : std::vector<int>{512, 256};
      last = e.in_dim;
      for (size_t i = 0; i < hidden.size(); ++i) {
        const int h = hidden[i];
        Lin l = ab.linear(h, last, m->split);
        Bn bn; bn.n = h;
        bn.g_off = ab.take_p(h); ab.seg(bn.g_off, h, 1, -1, 0);
        bn.b_off = ab.take_p(h); ab.seg(bn.b_off, h, 1, -1, 0);
        bn.rm_off = ab.take_b(h); bn.rv_off = ab.take_b(h);
        bn.counter = m->n_bn++;
        // the autoencoders' encoders are bare nn.Sequentials (directional_ae.py:21-27, 80-90): no ".fc" level
        const std::string stem = ae ? e.prefix + "." : e.prefix + ".fc.";
        const std::string fc = stem + std::to_string(4 * i);
        const std::string nb = stem + std::to_string(4 * i + 1);
        ab.info(fc + ".weight", VLA_TENSOR_PARAM, l.w_off, h, last);
        ab.info(fc + ".bias", VLA_TENSOR_PARAM, l.b_off, h);
        ab.info(nb + ".weight", VLA_TENSOR_PARAM, bn.g_off, h);
        ab.info(nb + ".bias", VLA_TENSOR_PARAM, bn.b_off, h);
        ab.info(nb + ".running_mean", VLA_TENSOR_BUFFER, bn.rm_off, h);
        ab.info(nb + ".running_var", VLA_TENSOR_BUFFER, bn.rv_off, h);
        ab.info(nb + ".num_batches_tracked", VLA_TENSOR_COUNTER, bn.counter, -1);
        e.fc.push_back(l); e.bn.push_back(bn);
        last = h;
        m->n_drop++;
      }
    }
    // fused heads: rows [0, L) = fc_mu, rows [L, 2L) = fc_logvar
    e.heads = ab.linear(m->HW, last, m->split);
    if (ae) {
      // one head: the last Linear of the encoder Sequential, or site_projection (directional_ae.py:26, 30, 89, 93)
      const std::string head = s.type == 'C' ? std::string("site_projection") : e.prefix + "." + std::to_string(4 * e.fc.size());
      ab.info(head + ".weight", VLA_TENSOR_PARAM, e.heads.w_off, L, last);
      ab.info(head + ".bias", VLA_TENSOR_PARAM, e.heads.b_off, L);
    } else {
      ab.info(e.prefix + ".fc_mu.weight", VLA_TENSOR_PARAM, e.heads.w_off, L, last);
      ab.info(e.prefix + ".fc_mu.bias", VLA_TENSOR_PARAM, e.heads.b_off, L);
      ab.info(e.prefix + ".fc_logvar.weight", VLA_TENSOR_PARAM, e.heads.w_off + static_cast<long long>(L) * last, L, last);
      ab.info(e.prefix + ".fc_logvar.bias", VLA_TENSOR_PARAM, e.heads.b_off + L, L);
    }
    m->encs.push_back(e);
  }
  // fused first decoder layers: one [sum of first hidden widths, L] matrix
  int cat_w = 0;
  for (const auto& s : ds) cat_w += s.type == 'A' ? 128 : (s.type == 'B' ? 256 : 64);
  m->dec_chunk0 = static_cast<int>(m->chunks_h.size());     // every chunk from here on belongs to a decoder
  m->cat = ab.linear(cat_w, L, m->split);
  int off = 0;
  for (const auto& s : ds) {
    Dec d; d.type = s.type; d.prefix = s.prefix;
    d.out_dim = s.type == 'A' ? c.dim_a : (s.type == 'B' ? c.dim_b : c.n_sites);
    d.cat_w = s.type == 'A' ? 128 : (s.type == 'B' ? 256 : 64);
    d.cat_off = off; off += d.cat_w;
    ab.info(d.prefix + ".fc.0.weight", VLA_TENSOR_PARAM, m->cat.w_off + static_cast<long long>(d.cat_off) * L, d.cat_w, L);
    ab.info(d.prefix + ".fc.0.bias", VLA_TENSOR_PARAM, m->cat.b_off + d.cat_off, d.cat_w);
    std::vector<int> widths = s.type == 'B' ? std::vector<int>{512, d.out_dim} : std::vector<int>{d.out_dim};
    int last = d.cat_w;
    for (size_t i = 0; i < widths.size(); ++i) {
      // The output layers feed no ReLU: plain bf16 -- except the site classifier's (64 x n_sites, negligible work), whose
      // argmax must agree with the reference's row by row (BASELINE.json north_star).
      Lin l = ab.linear(widths[i], last, m->split && (i + 1 < widths.size() || s.type == 'C'));
      const std::string fc = d.prefix + ".fc." + std::to_string(2 * (i + 1));
      ab.info(fc + ".weight", VLA_TENSOR_PARAM, l.w_off, widths[i], last);
      ab.info(fc + ".bias", VLA_TENSOR_PARAM, l.b_off, widths[i]);
      d.rest.push_back(l);
      last = widths[i];
    }
    m->decs.push_back(d);
  }
  m->n_params = ab.p; m->n_buffers = ab.b; m->n_shadow = ab.sh;
  return VLA_OK;
}

// ---------------------------------------------------------------------------------------------
// Workspace
// ---------------------------------------------------------------------------------------------
struct Bump {
  size_t off = 0; char* base = nullptr;
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void carve(vla_model* m, Bump& b, int cap) {
  const int L = m->L;
  const int mt = ceil_div(cap, GEMM_BM);
  m->ews.assign(m->encs.size(), EncWS{});
  for (size_t i = 0; i < m->encs.size(); ++i) {
    const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
    if (e.type == 'C') {
      w.ldx = op_ld(m->split, m->E); w.x_lo = op_lo(m->split, m->E); w.x = b.take<bf16>(static_cast<size_t>(cap) * w.ldx);
      w.ld_gx = pad8(m->E); w.g_x = b.take<bf16>(static_cast<size_t>(cap) * w.ld_gx);
      w.ld_onehot = pad8(m->S); w.onehot = b.take<bf16>(static_cast<size_t>(cap) * w.ld_onehot);
    } else {
      w.ldx = op_ld(m->split, e.in_dim); w.x_lo = op_lo(m->split, e.in_dim); w.x = b.take<bf16>(static_cast<size_t>(cap) * w.ldx);
      for (const Lin& l : e.fc) {
        w.ld_act.push_back(op_ld(m->split, l.out)); w.act_lo.push_back(op_lo(m->split, l.out));
        w.pre.push_back(b.take<float>(static_cast<size_t>(cap) * l.out));
        w.stats.push_back(b.take<float>(static_cast<size_t>(mt) * 2 * l.out));
        w.bstats.push_back(b.take<float>(static_cast<size_t>(ceil_div(cap, HB_ROWS)) * 2 * l.out));   // (head block: one partial per 32 rows)
        w.mean.push_back(b.take<float>(l.out));
        w.rstd.push_back(b.take<float>(l.out));
        w.act.push_back(b.take<bf16>(static_cast<size_t>(cap) * w.ld_act.back()));
        w.gy.push_back(b.take<bf16>(static_cast<size_t>(cap) * l.out));
        w.gpre.push_back(b.take<bf16>(static_cast<size_t>(cap) * l.out));
        w.bits.push_back(b.take<unsigned short>(static_cast<size_t>(cap) * ((l.out + 15) / 16)));
      }
    }
    w.ml = b.take<float>(static_cast<size_t>(cap) * m->HW);
  }
  m->mu = b.take<float>(static_cast<size_t>(cap) * L);
  m->logvar = b.take<float>(static_cast<size_t>(cap) * L);
  m->eps = b.take<float>(static_cast<size_t>(cap) * L);
  m->gz = b.take<float>(static_cast<size_t>(cap) * L);
  m->kl_partials = b.take<float>(std::max(std::max(ceil_div(cap * L, 256), 8 * mt), ceil_div(cap, HB_ROWS)) + 1);
  m->ldz = op_ld(m->split, L); m->z_lo = op_lo(m->split, L); m->z = b.take<bf16>(static_cast<size_t>(cap) * m->ldz);
  m->ldgml = pad8(m->HW); m->gml = b.take<bf16>(static_cast<size_t>(cap) * m->ldgml);
  m->ld_d0 = op_ld(m->split, m->cat.out); m->d0_lo = op_lo(m->split, m->cat.out);
  m->d0 = b.take<bf16>(static_cast<size_t>(cap) * m->ld_d0);
  m->g_d0 = b.take<bf16>(static_cast<size_t>(cap) * m->cat.out);
  m->d0_bits = b.take<unsigned short>(static_cast<size_t>(cap) * ((m->cat.out + 15) / 16));
  m->dws.assign(m->decs.size(), DecWS{});
  for (size_t i = 0; i < m->decs.size(); ++i) {
    const Dec& d = m->decs[i]; DecWS& w = m->dws[i];
    for (size_t r = 0; r + 1 < d.rest.size(); ++r) {
      // hidden activation r feeds layer r + 1: split only when that layer is itself followed by a ReLU
      const bool sp = m->split && (r + 2 < d.rest.size() || d.type == 'C');
      w.ld_act.push_back(op_ld(sp, d.rest[r].out)); w.act_lo.push_back(op_lo(sp, d.rest[r].out));
      w.act.push_back(b.take<bf16>(static_cast<size_t>(cap) * w.ld_act.back()));
      w.gact.push_back(b.take<bf16>(static_cast<size_t>(cap) * d.rest[r].out));
      w.bits.push_back(b.take<unsigned short>(static_cast<size_t>(cap) * ((d.rest[r].out + 15) / 16)));
    }
    w.ld_gout = pad8(d.out_dim);
    w.g_out = b.take<bf16>(static_cast<size_t>(cap) * w.ld_gout);
    w.recon = b.take<float>(static_cast<size_t>(cap) * d.out_dim);
  }
  const int lg = loss_grid_size(cap, m->cfg.dim_a, m->cfg.dim_b, m->S);
  m->loss_partials = b.take<float>(lg);
  size_t ep = 0;
  for (const Dec& d : m->decs) ep += static_cast<size_t>(mt) * ceil_div(d.out_dim, 32) * 8;
  m->eloss_partials = b.take<float>(std::max(ep, static_cast<size_t>(16) * mt) + 8);
}

int reserve(vla_model* m, int batch) {
  if (m->layout_only) return fail(VLA_ERR_STATE, "layout-only handle: no device state");
  if (batch <= m->cap) return VLA_OK;
  int cap = std::max(batch, 32);
  if (!m->seg.empty()) return fail(VLA_ERR_STATE, "workspace growth while recording a chain");
  // Captured CUDA graphs (vla_b200.Trainer) bake workspace addresses and tensor maps into their kernel arguments: growing
  // the workspace under them would leave the graphs replaying against freed memory.
  if (m->pinned > 0 && m->ws)
    return fail(VLA_ERR_STATE, "this handle's workspace is pinned by a live Trainer (captured CUDA graphs reference it) and holds " +
                                   std::to_string(m->cap) + " rows; a call with " + std::to_string(batch) +
                                   " rows would reallocate it.  Use a separate module / handle for the larger batch, or close the Trainer first");
  if (m->ws) { CK(cudaDeviceSynchronize()); CK(cudaFree(m->ws)); m->ws = nullptr; m->cap = 0; }
  m->tmaps.clear();
  free_plans(m);
  m->saved = false;
  Bump dry; carve(m, dry, cap);
  const size_t bytes = dry.off + 256;
  CK(cudaMalloc(&m->ws, bytes));
  CK(cudaMemset(m->ws, 0, bytes));
  Bump real; real.base = m->ws; carve(m, real, cap);
  m->cap = cap;
  return VLA_OK;
}

// ---------------------------------------------------------------------------------------------
// GEMM problem builders
// ---------------------------------------------------------------------------------------------
int get_tmap(vla_model* m, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
             uint32_t box_outer) {
  TmapKey k{reinterpret_cast<uintptr_t>(base), inner, outer, pitch_bytes, box_outer};
  auto it = m->tmaps.find(k);
  if (it == m->tmaps.end()) {
    CUtensorMap t;
    std::string err;
    if (!make_tmap_bf16(&t, base, inner, outer, pitch_bytes, 64, box_outer, &err)) return fail(VLA_ERR_CUDA, err);
    if (m->tmaps.size() > 4096) m->tmaps.clear();
    it = m->tmaps.emplace(k, t).first;
  }
  *out = it->second;
  return VLA_OK;
}

int choose_bn_tn(int N) {
  int best = 64; int best_cost = 1 << 30;
  for (int bn = 64; bn <= GEMM_BN_MAX_TN; bn += 64) {
    const int cost = ceil_div(N, bn) * bn;
    if (cost < best_cost || (cost == best_cost && bn > best)) { best = bn; best_cost = cost; }
  }
  return best;
}

// Per-group list of the B operands whose tensor maps wait for the joint tile-width choice (finalize_group)
thread_local PendingB g_pending[GEMM_MAX_PROBLEMS];
// ... and the A operands of an NT group (the multicast kernel needs their maps with GEMM_BM / 4-row boxes)
struct PendingA { const void* base; uint64_t inner, outer, pitch; };
thread_local PendingA g_pending_a[GEMM_MAX_PROBLEMS];

// C[M,N] = A[M,K] * W[N,K]^T ; A bf16 [M, lda], W bf16 [N, ldw]
// a_lo / b_lo > 0 (both or neither): split-bf16 operands, the lo copies sit a_lo / b_lo elements further along each row
int add_nt(vla_model* m, GemmGroup& g, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, int flags,
           GemmProblem** out, int force_bn = 0, int a_lo = 0, int b_lo = 0) {
  if (g.nprob >= GEMM_MAX_PROBLEMS) return fail(VLA_ERR_STATE, "too many problems in one GEMM group");
  if ((a_lo > 0) != (b_lo > 0) || (a_lo & 63) || (b_lo & 63)) return fail(VLA_ERR_STATE, "split GEMM operands need both lo copies at 64-element offsets");
  GemmProblem& p = g.p[g.nprob];
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.BN = force_bn;
  p.m_tiles = ceil_div(M, GEMM_BM);
  p.k_splits = 1; p.kb_per_split = ceil_div(K, GEMM_BK);
  p.flags = flags;
  p.a_lo = a_lo; p.b_lo = b_lo;
  int rc;
  if ((rc = get_tmap(m, &p.tmA, A, a_lo > 0 ? a_lo + K : K, M, static_cast<uint64_t>(lda) * 2, GEMM_BM))) return rc;
  g_pending_a[g.nprob] = PendingA{A, static_cast<uint64_t>(a_lo > 0 ? a_lo + K : K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda) * 2};
  g_pending[g.nprob] = PendingB{W, static_cast<uint64_t>(b_lo > 0 ? b_lo + K : K), static_cast<uint64_t>(N), static_cast<uint64_t>(ldw) * 2, 0, K, force_bn != 0};
  g.nprob++;
  *out = &p;
  return VLA_OK;
}

// dX[M,N] = dY[M,K] * W[K,N] ; dY bf16 [M, lda] (K-major A), W bf16 [K, ldw] = the forward weight copy [out, in] (MN-major B)
int add_nn(vla_model* m, GemmGroup& g, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, int flags,
           GemmProblem** out, int force_bn = 0) {
  if (g.nprob >= GEMM_MAX_PROBLEMS) return fail(VLA_ERR_STATE, "too many problems in one GEMM group");
  GemmProblem& p = g.p[g.nprob];
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.BN = force_bn;
  p.m_tiles = ceil_div(M, GEMM_BM);
  p.k_splits = 1; p.kb_per_split = ceil_div(K, GEMM_BK);
  p.flags = flags;
  int rc;
  if ((rc = get_tmap(m, &p.tmA, A, K, M, static_cast<uint64_t>(lda) * 2, GEMM_BM))) return rc;
  g_pending[g.nprob] = PendingB{W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw) * 2, 2, K, force_bn != 0};
  g.nprob++;
  *out = &p;
  return VLA_OK;
}

// Joint tile-width choice for an NT / NN group: minimise waves x slowest tile (exhaustive over <= 5 widths per problem,
// greedy coordinate descent when the group is large), then build the B tensor maps and the tile index ranges.
// Cost of one tile in the chain kernel (us-like units): fixed part (first operands, epilogue drain, boundary) + operand
// traffic of the main loop + epilogue columns.
double chain_tile_cost(const GemmProblem& p, int bn) {
  return 1.5 + (p.a_lo > 0 ? 2 : 1) * ceil_div(p.K, GEMM_BK) * (16384.0 + bn * 128.0) / 150e3 + 0.4 * ceil_div(bn, 64);
}
// Longest-processing-time assignment of the tiles of ONE row block to the four CTAs of a cluster.  units: (problem << 8) |
// n_tile per rank.  Returns the makespan, or a negative value when a rank would get more than CHAIN_MAX_UNITS tiles.
double chain_assign(const GemmGroup& g, const int* bn_of, std::vector<int> (*units)[CHAIN_CLUSTER]) {
  struct U { double c; int code; };
  std::vector<U> all;
  for (int i = 0; i < g.nprob; ++i) {
    const int bn = bn_of ? bn_of[i] : g.p[i].BN;
    const int nt = ceil_div(g.p[i].N, bn);
    for (int t = 0; t < nt; ++t) all.push_back(U{chain_tile_cost(g.p[i], bn), (i << 8) | t});
  }
  std::stable_sort(all.begin(), all.end(), [](const U& a, const U& b) { return a.c > b.c; });
  double load[CHAIN_CLUSTER] = {0, 0, 0, 0};
  int cnt[CHAIN_CLUSTER] = {0, 0, 0, 0};
  for (const U& u : all) {
    int r = 0;
    for (int k = 1; k < CHAIN_CLUSTER; ++k) if (load[k] < load[r]) r = k;
    load[r] += u.c; cnt[r]++;
    if (units) (*units)[r].push_back(u.code);
  }
  double mk = 0;
  for (int k = 0; k < CHAIN_CLUSTER; ++k) { mk = std::max(mk, load[k]); if (cnt[k] > CHAIN_MAX_UNITS) return -1.0; }
  return mk;
}

// Tile widths for the chain kernel: the tiles of one 128-row block are shared by the four CTAs of a cluster, so the widths
// minimise the busiest CTA's load (coordinate descent over the candidate widths of each problem, LPT assignment).
int finalize_chain(vla_model* m, GemmGroup& g, int mode) {
  const int n = g.nprob;
  int cand[GEMM_MAX_PROBLEMS][16], nc[GEMM_MAX_PROBLEMS], pick[GEMM_MAX_PROBLEMS], bn_of[GEMM_MAX_PROBLEMS];
  for (int i = 0; i < n; ++i) {
    const GemmProblem& p = g.p[i];
    nc[i] = 0;
    // full 32-column chunks wherever the epilogue reads or reduces per-column side inputs; CE needs the whole row in one chunk
    const bool chunky = (p.flags & (GF_COLSTATS | GF_BNSTATS | GF_MASK)) != 0 || ((p.flags & GF_LOSS) && p.loss_kind == LOSS_CE) ||
                        p.mask_bits_out != nullptr;      // (mask words cover whole 32-column chunks)
    const int step = mode == 0 ? (chunky ? 32 : 16) : 64;
    const int cap = mode == 0 ? GEMM_BN_MAX_NT : GEMM_BN_MAX_TN;
    if (g_pending[i].fixed) cand[i][nc[i]++] = p.BN;
    else for (int bn = step; bn <= cap && nc[i] < 16; bn += step) { cand[i][nc[i]++] = bn; if (bn >= p.N) break; }
    pick[i] = nc[i] - 1;                                  // start from the widest tiles
    bn_of[i] = cand[i][pick[i]];
  }
  double best = chain_assign(g, bn_of, nullptr);
  if (best < 0) best = 1e30;
  for (int sweep = 0; sweep < 4; ++sweep) {
    bool moved = false;
    for (int i = 0; i < n; ++i) {
      int keep = pick[i];
      for (int c = 0; c < nc[i]; ++c) {
        bn_of[i] = cand[i][c];
        const double v = chain_assign(g, bn_of, nullptr);
        if (v >= 0 && v < best - 1e-9) { best = v; keep = c; moved = true; }
      }
      pick[i] = keep; bn_of[i] = cand[i][keep];
    }
    if (!moved) break;
  }
  g.total_tiles = 0;
  for (int i = 0; i < n; ++i) {
    GemmProblem& p = g.p[i];
    p.BN = bn_of[i];
    p.n_tiles = ceil_div(p.N, p.BN);
    const PendingB& b = g_pending[i];
    int rc;
    if ((rc = get_tmap(m, &p.tmB, b.base, b.inner, b.outer, b.pitch, mode == 0 ? static_cast<uint32_t>(p.BN) : 64u))) return rc;
    p.tile_begin = g.total_tiles;
    g.total_tiles += p.m_tiles * p.n_tiles;
  }
  return VLA_OK;
}

int finalize_group(vla_model* m, GemmGroup& g, int mode) {
  if (m->chain_on) return finalize_chain(m, g, mode);
  const int n = g.nprob;
  int cand[GEMM_MAX_PROBLEMS][8], nc[GEMM_MAX_PROBLEMS], pick[GEMM_MAX_PROBLEMS];
  for (int i = 0; i < n; ++i) {
    nc[i] = 0;
    if (g_pending[i].fixed) { cand[i][nc[i]++] = g.p[i].BN; }
    else if (mode == 0) {
      for (int bn = 32; bn <= GEMM_BN_MAX_NT; bn += 32) cand[i][nc[i]++] = bn;
      // the widest tile (144 = 4.5 chunks) for problems whose epilogue does not need whole 32-column chunks
      const GemmProblem& q = g.p[i];
      const bool chunky = (q.flags & (GF_COLSTATS | GF_BNSTATS | GF_MASK)) != 0 || ((q.flags & GF_LOSS) && q.loss_kind == LOSS_CE) ||
                          q.mask_bits_out != nullptr;
      if (!chunky && GEMM_BN_MAX_NT % 32) cand[i][nc[i]++] = GEMM_BN_MAX_NT;
    }
    else { for (int bn = 64; bn <= GEMM_BN_MAX_TN; bn += 64) cand[i][nc[i]++] = bn; }
    // a width beyond the padded N only wastes MMA columns
    int keep = 0;
    for (int c = 0; c < nc[i]; ++c) { cand[i][keep++] = cand[i][c]; if (cand[i][c] >= g.p[i].N) break; }
    nc[i] = keep;
    pick[i] = 0;
  }
  auto cost = [&](const int* pk) {
    int tiles = 0; double slow = 0;
    for (int i = 0; i < n; ++i) {
      const int bn = cand[i][pk[i]];
      tiles += g.p[i].m_tiles * ceil_div(g.p[i].N, bn);
      const double t = 5.0 + (g.p[i].a_lo > 0 ? 2 : 1) * ceil_div(g.p[i].K, GEMM_BK) * (16384.0 + bn * 128.0) / 150e3 + 0.4 * ceil_div(bn, 64);
      slow = std::max(slow, t);
    }
    return ceil_div(tiles, 148) * slow + 1e-3 * tiles;
  };
  // coordinate descent from the narrowest tiles (a handful of sweeps converges for these tiny search spaces)
  double best = cost(pick);
  for (int sweep = 0; sweep < 4; ++sweep) {
    bool moved = false;
    for (int i = 0; i < n; ++i) {
      int keep = pick[i];
      for (int c = 0; c < nc[i]; ++c) {
        pick[i] = c;
        const double v = cost(pick);
        if (v < best - 1e-9) { best = v; keep = c; moved = true; }
      }
      pick[i] = keep;
    }
    if (!moved) break;
  }
  g.total_tiles = 0;
  for (int i = 0; i < n; ++i) {
    GemmProblem& p = g.p[i];
    p.BN = cand[i][pick[i]];
    p.n_tiles = ceil_div(p.N, p.BN);
    const PendingB& b = g_pending[i];
    int rc;
    if ((rc = get_tmap(m, &p.tmB, b.base, b.inner, b.outer, b.pitch, mode == 0 ? static_cast<uint32_t>(p.BN) : 64u))) return rc;
    p.tile_begin = g.total_tiles;
    g.total_tiles += p.m_tiles * p.n_tiles;
  }
  // One wave of NT tiles whose row blocks all have a multiple of 4 n-tiles: 4-CTA clusters with the A tile multicast
  // (gemm_tc_cluster_kernel; VLA_CLUSTER=0 turns it off).  The A maps are rebuilt with 32-row boxes.
  g.pad[0] = 0;
  static const bool cluster_on = [] { const char* e = getenv("VLA_CLUSTER"); return !(e && e[0] == '0'); }();
  if (mode == 0 && cluster_on && !recorder() && !g.dbg && !g.dbg_flags && g.total_tiles <= 148 &&
      g.total_tiles / 4 <= gemm_cluster_capacity()) {
    bool ok = n > 0;
    for (int i = 0; i < n; ++i) ok = ok && (g.p[i].n_tiles % 4 == 0) && g.p[i].k_splits == 1;
    if (ok) {
      for (int i = 0; i < n; ++i) {
        const PendingA& a = g_pending_a[i];
        int rc;
        if ((rc = get_tmap(m, &g.p[i].tmA, a.base, a.inner, a.outer, a.pitch, GEMM_BM / 4))) return rc;
      }
      g.pad[0] = 4;
    }
  }
  return VLA_OK;
}

// dW[M,N] += G[Kb,M]^T * X[Kb,N] ; G bf16 [Kb, ldg], X bf16 [Kb, ldx]; tiling is fixed up by finalize_tn
int add_tn(vla_model* m, GemmGroup& g, const bf16* G, int ldg, const bf16* X, int ldx, int M, int N, int Kb, float* dW,
           int ld_dw, float* dbias, int force_bn = 0) {
  if (g.nprob >= GEMM_MAX_PROBLEMS) return fail(VLA_ERR_STATE, "too many problems in one GEMM group");
  GemmProblem& p = g.p[g.nprob];
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = Kb;
  p.BN = force_bn ? force_bn : choose_bn_tn(N);
  p.m_tiles = ceil_div(M, GEMM_BM); p.n_tiles = ceil_div(N, p.BN);
  p.flags = GF_RED | (dbias ? GF_BIASGRAD : 0);
  p.out_f32 = dW; p.ld_f32 = ld_dw; p.bias_grad = dbias;
  int rc;
  if ((rc = get_tmap(m, &p.tmA, G, M, Kb, static_cast<uint64_t>(ldg) * 2, 64))) return rc;
  if ((rc = get_tmap(m, &p.tmB, X, N, Kb, static_cast<uint64_t>(ldx) * 2, 64))) return rc;
  g.nprob++;
  return VLA_OK;
}
void finalize_tn(GemmGroup& g, int Kb, int force_splits = 0) {
  int base = 0;
  for (int i = 0; i < g.nprob; ++i) base += g.p[i].m_tiles * g.p[i].n_tiles;
  const int kb_total = ceil_div(Kb, GEMM_BK);
  int splits = force_splits ? force_splits : std::max(1, 148 / std::max(base, 1));     // one wave
  splits = std::min(splits, kb_total);
  const int per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, per);
  g.total_tiles = 0;
  for (int i = 0; i < g.nprob; ++i) {
    GemmProblem& p = g.p[i];
    p.k_splits = splits; p.kb_per_split = per;
    p.tile_begin = g.total_tiles;
    g.total_tiles += p.m_tiles * p.n_tiles * splits;
  }
}

// Zero everything: the chain plan cache compares whole argument images, so no byte of a group may be stack garbage.
void init_group(GemmGroup& g) { memset(static_cast<void*>(&g), 0, sizeof(g)); }

// ---------------------------------------------------------------------------------------------
// Sequencing
// ---------------------------------------------------------------------------------------------
struct FwdIO {
  const float* params; float* buffers; long long* counters;
  const float* x[2]; const long long* site;
  int batch, train;
  const float* eps; const unsigned char* const* keep_masks;
  unsigned long long seed, offset;
  float* recon[3];   // by decoder type A, B, C
  float* mu; float* logvar;
  bool engine;       // train step: dyn-driven Philox offsets / step bump
  int n_batches;     // train step over a resident dataset
  float beta1, beta2;
  // train step: the last decoder layers also produce the loss partials and dL/d(pre-activation) (no separate loss pass)
  bool fuse_loss = false;
  const float* tgt_a = nullptr; const float* tgt_b = nullptr; const long long* tgt_site = nullptr;
  const float* class_w = nullptr; float* loss_out = nullptr;
  bool rc_prefix = false;   // row-chain step: only ingest + the first encoder layer (the row-chain kernel takes over behind it)
  bool defer_loss_sum = false; // single-GPU whole step: the loss tiles only write partials, AdamW's block 0 does the final sum
  vla_dp* sync_dp = nullptr;   // opt-in SyncBN: BatchNorm statistics over the global batch (all-reduce of the column sums)
};

int present_mask(const vla_model* m, const FwdIO& io) {
  int mask = 0;
  for (size_t i = 0; i < m->encs.size(); ++i) {
    const Enc& e = m->encs[i];
    const bool here = e.slot == 2 ? io.site != nullptr : io.x[e.slot] != nullptr;
    if (here) mask |= 1 << i;
  }
  return mask;
}

int run_shadow_refresh(vla_model* m, const float* params, cudaStream_t st) {
  AdamArgs a{};
  a.p = const_cast<float*>(params); a.shadow = m->shadow;
  a.chunks = m->chunks_d; a.n_chunks = static_cast<int>(m->chunks_h.size());
  a.update = 0;
  { ProfScope ps(m, st, "shadow_refresh", 0, 6.0 * m->n_params); CK(launch_adamw(a, st)); }
  return VLA_OK;
}

// ---------------------------------------------------------------------------------------------
// Head block (headblock.cu): BatchNorm apply of the last hidden layers + heads + latent + fused first decoder layer as one
// CUDA-core launch (and the mirror-image backward) instead of four / three tensor-core and element-wise launches.
// ---------------------------------------------------------------------------------------------
// Opt-in (VLA_HEADBLOCK=1).  Measured (rna2dna, batch 4096, profiles/r2_headblock.md): 11 launches instead of 16, results
// within the oracle tolerances (fp32 instead of split-bf16 for these layers), but 23 us + 21 us for the two launches against
// ~30 us of step time for the seven launches they replace: the step takes 135 us instead of 122 us.
bool headblock_enabled() {
  if (recorder()) return false;         // lock-step population steps run the separate launches
  const char* e = getenv("VLA_HEADBLOCK");
  return e && e[0] == '1';
}
bool headblock_fits(const vla_model* m, int present) {
  if (m->HW > 64 || m->L > 64 || (m->E & 3) || m->cat.out > 512 || (m->cat.out & 31)) return false;
  for (size_t i = 0; i < m->encs.size(); ++i) {
    if (!(present >> i & 1)) continue;
    const Enc& e = m->encs[i];
    if (e.type == 'C') continue;
    if (e.fc.empty() || e.fc.back().out > 256 || (e.fc.back().out & 63)) return false;
  }
  return true;
}
int fill_headblock(vla_model* m, HbArgs* a, int B, int present, const float* P, float* buffers, long long* counters, int train,
                   const unsigned char* const* keep_masks, unsigned long long seed, unsigned long long offset, bool engine,
                   int n_batches, const long long* site) {
  memset(static_cast<void*>(a), 0, sizeof(*a));
  a->rows = B; a->L = m->L; a->HW = m->HW; a->ae = m->ae ? 1 : 0; a->C = m->cat.out; a->n_batches = n_batches; a->p_drop = 0.1f;
  a->seed = seed; a->lat_offset = offset * 16; a->dyn = m->dyn;
  (void)engine;
  const int mt = ceil_div(B, GEMM_BM);
  for (size_t i = 0; i < m->encs.size(); ++i) {
    if (!(present >> i & 1)) continue;
    const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
    HbEnc& E = a->enc[a->n_enc++];
    E.Wh = P + e.heads.w_off; E.bh = P + e.heads.b_off;
    if (e.type == 'C') {
      E.kind = 1; E.in_dim = m->E; E.site = site; E.emb = P + e.emb_off; E.g_x = w.g_x; E.ld_gx = w.ld_gx;
    } else {
      const size_t r = e.fc.size() - 1;
      const Bn& bn = e.bn[r];
      E.kind = 0; E.in_dim = bn.n; E.pre = w.pre[r]; E.stats = w.stats[r]; E.m_tiles = mt; E.train = train;
      E.gamma = P + bn.g_off; E.beta = P + bn.b_off;
      E.running_mean = buffers ? buffers + bn.rm_off : nullptr; E.running_var = buffers ? buffers + bn.rv_off : nullptr;
      E.nbt = counters ? counters + bn.counter : nullptr;
      E.save_mean = w.mean[r]; E.save_rstd = w.rstd[r];
      E.keep_mask = keep_masks ? keep_masks[e.first_drop + r] : nullptr;
      E.drop_offset = offset * 16 + 1 + e.first_drop + r;
      E.act = w.act[r]; E.ld_act = w.ld_act[r]; E.bits = reinterpret_cast<unsigned int*>(w.bits[r]);
      E.gy = w.gy[r]; E.bstats = w.bstats[r]; E.mask_scale = train ? 1.0f / 0.9f : 1.0f;
    }
  }
  a->mu = m->mu; a->logvar = m->logvar; a->eps_save = m->eps; a->z = m->z; a->ld_z = m->ldz; a->kl_partials = m->kl_partials;
  a->W0 = P + m->cat.w_off; a->b0 = P + m->cat.b_off; a->d0 = m->d0; a->ld_d0 = m->ld_d0; a->d0_lo = m->d0_lo;
  a->d0_bits = reinterpret_cast<unsigned int*>(m->d0_bits);
  a->g_d0 = m->g_d0; a->ld_gd0 = m->cat.out; a->gml = m->gml; a->ld_gml = m->ldgml;
  if (hb_smem_bytes(*a, false) > 227 * 1024 || hb_smem_bytes(*a, true) > 227 * 1024) return 1;
  return 0;
}

int run_forward(vla_model* m, const FwdIO& io, cudaStream_t st) {
  const int B = io.batch, L = m->L;
  if (B <= 0) return fail(VLA_ERR_INVALID, "batch must be positive");
  const int present = present_mask(m, io);
  if (!present) return fail(VLA_ERR_INVALID, "no encoder input present");
  int rc;
  if ((rc = reserve(m, B))) return rc;
  if (io.train && B < 2)
    for (size_t i = 0; i < m->encs.size(); ++i)
      if ((present >> i & 1) && m->encs[i].type != 'C')
        return fail(VLA_ERR_INVALID, "Expected more than 1 value per channel when training (BatchNorm1d)");
  const float* P = io.params;
  const bf16* SH = m->shadow;

  if (m->prof_on) { ProfScope calib(m, st, "_empty_pair", 0, 0); }   // event-pair overhead, subtracted by the reader
  // ---- ingest ----
  {
    IngestArgs a{};
    a.rows = B;
    for (size_t i = 0; i < m->encs.size(); ++i) {
      if (!(present >> i & 1)) continue;
      const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
      if (e.type == 'C') {
        a.site = io.site; a.emb = P + e.emb_off; a.h_site = w.x; a.ld_hsite = w.ldx; a.hsite_lo = w.x_lo;
        a.onehot = w.onehot; a.ld_onehot = w.ld_onehot; a.n_sites = m->S; a.embed = m->E;
      } else {
        a.src[a.n] = io.x[e.slot]; a.dst[a.n] = w.x; a.width[a.n] = e.in_dim; a.ld_dst[a.n] = w.ldx; a.lo_off[a.n] = w.x_lo; a.n++;
      }
    }
    a.dyn = m->dyn; a.bump_step = io.engine ? 1 : 0; a.n_batches = io.n_batches; a.beta1 = io.beta1; a.beta2 = io.beta2;
    { double by = 0; for (int e = 0; e < a.n; ++e) by += static_cast<double>(B) * a.width[e] * (4.0 + (a.lo_off[e] > 0 ? 4.0 : 2.0));
      if (m->chain_on) { if ((rc = chain_add(m, st, CK_INGEST, -1, &a, sizeof(a), "ingest", 0, by, false))) return rc; }
      else { ProfScope ps(m, st, "ingest", 0, by); CK(launch_ingest(a, st)); } }
  }
  const int mt = ceil_div(B, GEMM_BM);
  // the small layers around the latent as one CUDA-core launch (headblock.cu)?
  HbArgs hb;
  const bool use_hb = headblock_enabled() && !m->chain_on && !io.rc_prefix && headblock_fits(m, present) &&
                      fill_headblock(m, &hb, B, present, P, io.buffers, io.counters, io.train, io.keep_masks, io.seed, io.offset,
                                     io.engine, io.n_batches, io.site) == 0;
  m->hb_used = use_hb;
  // ---- encoders, round by round ----
  size_t max_depth = 0;
  for (size_t i = 0; i < m->encs.size(); ++i) if (present >> i & 1) max_depth = std::max(max_depth, m->encs[i].fc.size());
  for (size_t r = 0; r <= max_depth; ++r) {
    GemmGroup g; init_group(g);
    for (size_t i = 0; i < m->encs.size(); ++i) {
      if (!(present >> i & 1)) continue;
      const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
      GemmProblem* p;
      if (r < e.fc.size()) {
        const Lin& l = e.fc[r];
        const bf16* A = r == 0 ? w.x : w.act[r - 1];
        const int lda = r == 0 ? w.ldx : w.ld_act[r - 1];
        const int a_lo = r == 0 ? w.x_lo : w.act_lo[r - 1];
        const int flags = GF_BIAS | GF_OUT_F32 | (io.train ? GF_COLSTATS : 0);
        if ((rc = add_nt(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.out, l.in, flags, &p, 0, a_lo, l.sh_lo))) return rc;
        p->bias = P + l.b_off; p->out_f32 = w.pre[r]; p->ld_f32 = l.out; p->stats = w.stats[r];
      } else if (r == std::max(e.fc.size(), std::min<size_t>(1, max_depth)) && !io.rc_prefix && !use_hb) {
        // (an encoder without hidden layers -- the site encoder -- runs its heads in round 1 with the other heads, not beside
        // the wide first layers: round 0 then fits one wave of 32-column tiles)
        const Lin& l = e.heads;
        const size_t d = e.fc.size();
        const bf16* A = d == 0 ? w.x : w.act[d - 1];
        const int lda = d == 0 ? w.ldx : w.ld_act[d - 1];
        const int a_lo = d == 0 ? w.x_lo : w.act_lo[d - 1];
        if ((rc = add_nt(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.out, l.in, GF_BIAS | GF_OUT_F32, &p, 0, a_lo, l.sh_lo))) return rc;
        p->bias = P + l.b_off; p->out_f32 = w.ml; p->ld_f32 = m->HW;
      }
    }
    if (g.nprob && (rc = timed_gemm(m, g, 0, r == 0 ? "gemm_enc_l0" : (r == 1 ? "gemm_enc_l1" : "gemm_enc_l2"), st))) return rc;
    if (io.rc_prefix) {
      m->saved = true; m->saved_batch = B; m->saved_present = present; m->saved_train = io.train;
      m->generation++;
      return VLA_OK;
    }
    std::vector<BnActArgs> round_bn;                     // BatchNorm applies of this round: independent of one another
    for (size_t i = 0; i < m->encs.size(); ++i) {
      if (!(present >> i & 1)) continue;
      const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
      if (r >= e.fc.size()) continue;
      if (use_hb && r + 1 == e.fc.size()) continue;      // the last hidden layer's BatchNorm apply belongs to the head block
      const Bn& bn = e.bn[r];
      BnActArgs a{};
      a.pre = w.pre[r]; a.ld_pre = bn.n; a.stats = w.stats[r]; a.m_tiles = mt;
      if (io.sync_dp && io.train) {
        if ((rc = run_stats_allreduce(m, io.sync_dp, w.stats[r], mt, bn.n, "syncbn_fwd", st))) return rc;
        a.m_tiles = 1; a.stat_rows = B * io.sync_dp->world;
      }
      a.gamma = P + bn.g_off; a.beta = P + bn.b_off;
      a.running_mean = io.buffers + bn.rm_off; a.running_var = io.buffers + bn.rv_off;
      a.num_batches_tracked = io.counters ? io.counters + bn.counter : nullptr;
      a.save_mean = w.mean[r]; a.save_rstd = w.rstd[r];
      a.out = w.act[r]; a.ld_out = w.ld_act[r]; a.out_lo = w.act_lo[r];
      a.keep_mask = io.keep_masks ? io.keep_masks[e.first_drop + r] : nullptr;
      a.rows = B; a.n = bn.n; a.train = io.train; a.update_running = io.train; a.p_drop = 0.1f;
      a.seed = io.seed; a.offset = io.offset * 16 + 1 + e.first_drop + r; a.dyn = io.engine ? m->dyn : nullptr;
      // train mode needs the statistics of the WHOLE batch: a grid-wide dependency, the stretch ends in front of it
      if (m->chain_on) { if ((rc = chain_add(m, st, CK_BN_ACT, -1, &a, sizeof(a), "bn_act", 0, 6.0 * B * bn.n, a.train != 0))) return rc; }
      else round_bn.push_back(a);
    }
    // two BatchNorm layers in one round (both dense encoders): ONE launch for the pair -- a launch less on the chain
    if (round_bn.size() == 2 && !recorder()) {
      ProfScope ps(m, st, "bn_act", 0, 6.0 * B * (round_bn[0].n + round_bn[1].n));
      CK(launch_bn_act_pair(round_bn[0], round_bn[1], st));
    } else {
      for (const BnActArgs& a : round_bn) { ProfScope ps(m, st, "bn_act", 0, 6.0 * B * a.n); CK(launch_bn_act(a, st)); }
    }
  }
  // ---- latent ----
  if (use_hb) {
    hb.eps_in = io.eps;
    if (!io.engine) hb.dyn = nullptr;                    // per-call path: Philox offsets come from the caller, no step counter
    hb.n_batches = io.engine ? io.n_batches : 1;
    const double by = static_cast<double>(B) * (m->L * 16.0 + m->cat.out * 4.0);
    double fl = 0;
    for (int e = 0; e < hb.n_enc; ++e) fl += 2.0 * B * hb.enc[e].in_dim * m->HW;
    fl += 2.0 * B * m->L * m->cat.out;
    { ProfScope ps(m, st, "head_block_fwd", fl, by); CK(launch_head_block_fwd(hb, st)); }
    m->kl_grid = ceil_div(B, HB_ROWS);
    if (io.mu) CK(cudaMemcpyAsync(io.mu, m->mu, sizeof(float) * B * L, cudaMemcpyDeviceToDevice, st));
    if (io.logvar) CK(cudaMemcpyAsync(io.logvar, m->logvar, sizeof(float) * B * L, cudaMemcpyDeviceToDevice, st));
  } else {
    LatentFwdArgs a{};
    for (size_t i = 0; i < m->encs.size(); ++i)
      if (present >> i & 1) { a.ml[a.n_enc] = m->ews[i].ml; a.ld_ml[a.n_enc] = m->HW; a.n_enc++; }
    a.eps_in = io.eps; a.seed = io.seed; a.offset = io.offset * 16; a.dyn = io.engine ? m->dyn : nullptr;
    a.mu = m->mu; a.logvar = m->logvar; a.eps_save = m->eps; a.z = m->z; a.ld_z = m->ldz; a.z_lo = m->z_lo;
    a.kl_partials = m->kl_partials; a.rows = B; a.L = L; a.ae = m->ae ? 1 : 0;
    const double lat_by = static_cast<double>(B) * L * (8.0 * a.n_enc + 14.0);
    if (m->chain_on) {
      m->kl_grid = CHAIN_CLUSTER * mt;               // one KL partial per (row block, cluster rank)
      if ((rc = chain_add(m, st, CK_LATENT_FWD, -1, &a, sizeof(a), "latent_fwd", 0, lat_by, false))) return rc;
    } else {
      ProfScope ps(m, st, "latent_fwd", 0, lat_by); CK(launch_latent_fwd(a, &m->kl_grid, st));
    }
    if (io.mu || io.logvar) {
      if ((rc = chain_flush(m, st))) return rc;
      if (io.mu) CK(cudaMemcpyAsync(io.mu, m->mu, sizeof(float) * B * L, cudaMemcpyDeviceToDevice, st));
      if (io.logvar) CK(cudaMemcpyAsync(io.logvar, m->logvar, sizeof(float) * B * L, cudaMemcpyDeviceToDevice, st));
    }
  }
  // ---- decoders ----
  if (!use_hb) {
    GemmGroup g; init_group(g); GemmProblem* p;
    const Lin& l = m->cat;
    if ((rc = add_nt(m, g, m->z, m->ldz, SH + l.sh_off, l.sh_ld, B, l.out, l.in, GF_BIAS | GF_RELU | GF_OUT_BF16, &p, 0, m->z_lo, l.sh_lo))) return rc;
    p->bias = P + l.b_off; p->out_bf16 = m->d0; p->ld_bf16 = m->ld_d0; p->out_lo = m->d0_lo;
    if (l.out % 32 == 0) p->mask_bits_out = reinterpret_cast<unsigned int*>(m->d0_bits);
    if ((rc = timed_gemm(m, g, 0, "gemm_dec_l0", st))) return rc;
  }
  size_t max_rest = 0;
  for (const Dec& d : m->decs) max_rest = std::max(max_rest, d.rest.size());
  std::vector<GemmGroup> dec_groups;
  for (size_t r = 0; r < max_rest; ++r) {
    GemmGroup g; init_group(g);
    for (size_t i = 0; i < m->decs.size(); ++i) {
      const Dec& d = m->decs[i]; DecWS& w = m->dws[i];
      if (r >= d.rest.size()) continue;
      const Lin& l = d.rest[r];
      const bf16* A = r == 0 ? m->d0 + d.cat_off : w.act[r - 1];
      const int lda = r == 0 ? m->ld_d0 : w.ld_act[r - 1];
      // split operands for the hidden layers (their result feeds a ReLU); the output layer runs plain bf16
      const int a_lo = l.sh_lo > 0 ? (r == 0 ? m->d0_lo : w.act_lo[r - 1]) : 0;
      GemmProblem* p;
      const bool last = r + 1 == d.rest.size();
      if (last) {
        const int slot = d.type == 'A' ? 0 : (d.type == 'B' ? 1 : 2);
        if (io.fuse_loss) {
          // output -> loss partials + bf16 dL/d(pre-activation); the fp32 output is written only if the caller wants it
          const int flags = GF_BIAS | GF_LOSS | GF_OUT_BF16 | (io.recon[slot] ? GF_OUT_F32 : 0) | (d.type == 'B' ? GF_SIGMOID : 0);
          if ((rc = add_nt(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.out, l.in, flags, &p, 0, a_lo, l.sh_lo))) return rc;
          p->bias = P + l.b_off; p->out_f32 = io.recon[slot]; p->ld_f32 = l.out;
          p->out_bf16 = w.g_out; p->ld_bf16 = w.ld_gout;
          p->loss_kind = d.type == 'A' ? LOSS_MSE : (d.type == 'B' ? LOSS_BCE : LOSS_CE);
          p->aux0 = d.type == 'A' ? io.tgt_a : (d.type == 'B' ? io.tgt_b : nullptr);
          p->aux1 = d.type == 'C' ? io.class_w : nullptr;
          p->aux_site = d.type == 'C' ? io.tgt_site : nullptr;
          p->aux_n = io.n_batches; p->dyn = m->dyn;
        } else {
          float* dst = io.recon[slot] ? io.recon[slot] : w.recon;
          const int flags = GF_BIAS | GF_OUT_F32 | (d.type == 'B' ? GF_SIGMOID : 0);
          if ((rc = add_nt(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.out, l.in, flags, &p, 0, a_lo, l.sh_lo))) return rc;
          p->bias = P + l.b_off; p->out_f32 = dst; p->ld_f32 = l.out;
        }
      } else {
        if ((rc = add_nt(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.out, l.in, GF_BIAS | GF_RELU | GF_OUT_BF16, &p, 0, a_lo, l.sh_lo))) return rc;
        p->bias = P + l.b_off; p->out_bf16 = w.act[r]; p->ld_bf16 = w.ld_act[r]; p->out_lo = w.act_lo[r];
        if (l.out % 32 == 0) p->mask_bits_out = reinterpret_cast<unsigned int*>(w.bits[r]);
      }
    }
    if (g.nprob) {
      if ((rc = finalize_group(m, g, 0))) return rc;
      dec_groups.push_back(g);
    }
  }
  if (io.fuse_loss) {
    // partial-sum layout [mse | bce | ce] and the ticket count of the final reduction, shared by every loss launch
    LossTail T{};
    int tiles[4] = {0, 0, 0, 0};
    for (const GemmGroup& g : dec_groups)
      for (int i = 0; i < g.nprob; ++i)
        if (g.p[i].flags & GF_LOSS) tiles[g.p[i].loss_kind] += g.p[i].m_tiles * g.p[i].n_tiles;
    T.counter = m->loss_counter; T.total_tickets = tiles[LOSS_MSE] + tiles[LOSS_BCE] + tiles[LOSS_CE];
    T.n_mse = 8 * tiles[LOSS_MSE]; T.n_bce = 8 * tiles[LOSS_BCE]; T.n_ce = 8 * tiles[LOSS_CE];
    T.partials = m->eloss_partials; T.kl_partials = m->kl_partials; T.n_kl = m->kl_grid;
    T.out = io.loss_out; T.dyn = m->dyn; T.dyn_bump = m->dyn;
    m->deferred_tail_valid = false;
    if (io.defer_loss_sum) { T.counter = nullptr; m->deferred_tail = T; m->deferred_tail_valid = true; }
    const int base[4] = {0, 0, T.n_mse, T.n_mse + T.n_bce};
    for (GemmGroup& g : dec_groups) {
      g.tail = T;
      for (int i = 0; i < g.nprob; ++i)
        if (g.p[i].flags & GF_LOSS) g.p[i].aux_partials = m->eloss_partials + base[g.p[i].loss_kind];
    }
  }
  for (size_t r = 0; r < dec_groups.size(); ++r)
    if ((rc = timed_gemm(m, dec_groups[r], 0, r == 0 ? "gemm_dec_l1" : "gemm_dec_l2", st, true))) return rc;
  m->saved = true; m->saved_batch = B; m->saved_present = present; m->saved_train = io.train;
  m->generation++;
  return VLA_OK;
}

// Arguments of the peer-memory exchange over float2s [first2, end2) of the flat gradient buffer (dp_exchange.cu).
// part selects the RECV region and the trace slots (0: main stream, 1: side stream).
DpArgs make_dp_args(vla_model* m, vla_dp* dp, long long first2, long long end2, int part) {
  DpArgs x{};
  x.world = dp->world; x.rank = dp->rank; x.dyn = m->dyn;
  x.first2 = first2; x.n2 = std::max(0LL, end2 - first2);
  x.per2 = ((x.n2 + dp->world - 1) / dp->world + 1) & ~1LL;      // even: a float4 of the arena never straddles two shards
  x.g = reinterpret_cast<float*>(dp->local);
  for (int r = 0; r < dp->world; ++r) {
    x.recv[r] = reinterpret_cast<uint4*>(dp->peer[r] + dp->off_recv) + static_cast<size_t>(part) * dp->world * dp->per2;
    x.rsum[r] = reinterpret_cast<uint4*>(dp->peer[r] + dp->off_rsum);
  }
  x.trace = reinterpret_cast<unsigned long long*>(dp->local + dp->off_trace) + 4 * part;
  return x;
}
// Send the decoder gradients early, on the side stream, while the encoder backward runs?  Round 1 (rna2dna, batch 4096 per
// GPU, profiles/r1_dp_exchange.md): neutral at 2 GPUs, -4 us at 8 GPUs.  Round 2, with the decoder weight gradients on the
// low-priority 64-CTA side branch instead (as on one GPU): 137.0 us without the early exchange vs 139.3 us with it at 8 GPUs
// (profiles/r2_bench_n8*.json) -- the early exchange's polling blocks and its full-width weight-gradient launch hold SMs the
// main chain's GEMM CTAs need (dgrad_enc_l1 13.7 vs 10.6 us).  Default: off.  VLA_DP_OVERLAP=1 turns it on (every rank alike).
bool dp_overlap(const vla_dp* dp) {
  (void)dp;
  const char* e = getenv("VLA_DP_OVERLAP");
  return e && e[0] == '1';
}
// algorithmic bytes of an exchange: payload out + in over NVLink
double dp_bytes(const DpArgs& x) { return 16.0 * x.n2 * (x.world - 1) / x.world; }
int run_exchange(vla_model* m, vla_dp* dp, long long first2, long long end2, int part, cudaStream_t st, bool pdl) {
  const DpArgs x = make_dp_args(m, dp, first2, end2, part);
  if (x.n2 <= 0) return VLA_OK;
  { int rcf = chain_flush(m, st); if (rcf) return rcf; }
  ProfScope ps(m, st, part ? "dp_exchange_dec" : "dp_exchange_enc", 0, dp_bytes(x));
  CK(launch_dp_exchange(x, st, pdl));
  return VLA_OK;
}

// SyncBN: sum a BatchNorm layer's per-tile column sums over the ranks, in place (tile 0 then holds the global sums).
int run_stats_allreduce(vla_model* m, vla_dp* dp, float* partials, int m_tiles, int n, const char* name, cudaStream_t st) {
  if (n > DP_SMALL_WORDS || (n & 1)) return fail(VLA_ERR_INVALID, "SyncBN: BatchNorm width must be even and at most 1024");
  if (dp->small_next >= DP_SMALL_REGIONS) return fail(VLA_ERR_STATE, "SyncBN: more BatchNorm exchanges in one step than regions");
  DpSmallArgs a{};
  a.world = dp->world; a.rank = dp->rank; a.partials = partials; a.m_tiles = m_tiles; a.n = n; a.dyn = m->dyn;
  const size_t region = static_cast<size_t>(dp->small_next++) * dp->world * DP_SMALL_WORDS;
  for (int r = 0; r < dp->world; ++r) a.slots[r] = reinterpret_cast<uint4*>(dp->peer[r] + dp->off_small) + region;
  { int rcf = chain_flush(m, st); if (rcf) return rcf; }
  ProfScope ps(m, st, name, 0, 16.0 * n * dp->world);
  CK(launch_dp_small_allreduce(a, st));
  return VLA_OK;
}

// Side stream of the model (lowest priority: its kernels fill the SMs the main chain leaves idle).  VLA_SIDE=0 turns it off.
bool side_ready(vla_model* m) {
  static const bool on = [] { const char* e = getenv("VLA_SIDE"); return !(e && e[0] == '0'); }();
  if (!on) return false;
  if (m->side) return true;
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  if (cudaStreamCreateWithPriority(&m->side, cudaStreamNonBlocking, lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    (void)cudaGetLastError();
    if (m->side) { cudaStreamDestroy(m->side); m->side = nullptr; }
    return false;
  }
  return true;
}

struct BwdIO {
  const float* params;
  const float* g_recon[3]; const float* recon_b;   // fp32 upstream gradients (autograd path), may be null
  const float* g_mu; const float* g_logvar;
  float* grads;
  bool engine;                                     // bf16 output gradients already written by the loss kernel
  bool zero_grads;
  vla_dp* dp = nullptr;                            // data parallel: the decoder weight gradients are computed and sent early
  bool rc_suffix = false;                          // row-chain step: only BatchNorm backward of the first layers + weight gradients
  bool side_dec = false;                           // single GPU engine step: decoder weight gradients on the model's side stream
  bool sync_bn = false;                            // opt-in SyncBN (with dp): global BatchNorm-backward sums
};

int run_backward(vla_model* m, const BwdIO& io, cudaStream_t st) {
  if (!m->saved) return fail(VLA_ERR_STATE, "vla_backward without a preceding vla_forward on this handle");
  const int B = m->saved_batch, L = m->L, present = m->saved_present, train = m->saved_train;
  const float* P = io.params; const bf16* SH = m->shadow; float* G = io.grads;
  int rc;
  if (io.zero_grads) { if ((rc = chain_flush(m, st))) return rc; CK(cudaMemsetAsync(G, 0, sizeof(float) * m->n_params, st)); }
  // which decoders carry a gradient
  std::vector<bool> active(m->decs.size(), false);
  const bool sfx = io.rc_suffix;
  if (sfx) {
    for (size_t i = 0; i < m->decs.size(); ++i) active[i] = true;
  } else if (!io.engine) {
    OutGradArgs og[3]; int n = 0;
    for (size_t i = 0; i < m->decs.size(); ++i) {
      const Dec& d = m->decs[i]; DecWS& w = m->dws[i];
      const int slot = d.type == 'A' ? 0 : (d.type == 'B' ? 1 : 2);
      if (!io.g_recon[slot]) continue;
      if (d.type == 'B' && !io.recon_b) return fail(VLA_ERR_INVALID, "g_recon_b given without recon_b");
      og[n].g = io.g_recon[slot]; og[n].width = d.out_dim; og[n].y = d.type == 'B' ? io.recon_b : nullptr;
      og[n].dst = w.g_out; og[n].ld_dst = w.ld_gout; og[n].rows = B; n++;
      active[i] = true;
    }
    if ((rc = chain_flush(m, st))) return rc;
    { ProfScope ps(m, st, "out_grad", 0, 0); CK(launch_out_grad(og, n, st)); }
  } else {
    for (size_t i = 0; i < m->decs.size(); ++i) active[i] = true;
  }
  bool any_dec = false;
  for (bool a : active) any_dec = any_dec || a;
  // ---- decoder data gradients, last layer first ----
  size_t max_rest = 0;
  for (const Dec& d : m->decs) max_rest = std::max(max_rest, d.rest.size());
  for (size_t rr = max_rest; rr-- > 0 && !sfx;) {
    GemmGroup g; init_group(g);
    for (size_t i = 0; i < m->decs.size(); ++i) {
      const Dec& d = m->decs[i]; DecWS& w = m->dws[i];
      if (!active[i] || rr >= d.rest.size()) continue;
      const Lin& l = d.rest[rr];
      const bool last = rr + 1 == d.rest.size();
      const bf16* A = last ? w.g_out : w.gact[rr];
      const int lda = last ? w.ld_gout : l.out;
      GemmProblem* p;
      // dX[B, in] = dY[B, out] * W[out, in]  ->  A K-major, B = forward weight copy read MN-major
      if ((rc = add_nn(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.in, l.out, GF_MASK | GF_OUT_BF16, &p))) return rc;
      if (rr == 0) {
        p->mask_src = m->d0 + d.cat_off; p->ld_mask = m->ld_d0;
        p->out_bf16 = m->g_d0 + d.cat_off; p->ld_bf16 = m->cat.out;
        if (m->cat.out % 32 == 0 && d.cat_off % 32 == 0)       // the forward left (activation > 0) bits, [chunk][B] words
          p->mask_bits_in = reinterpret_cast<const unsigned int*>(m->d0_bits) + static_cast<size_t>(d.cat_off / 32) * B;
      } else {
        p->mask_src = w.act[rr - 1]; p->ld_mask = w.ld_act[rr - 1];
        p->out_bf16 = w.gact[rr - 1]; p->ld_bf16 = l.in;
        if (l.in % 32 == 0) p->mask_bits_in = reinterpret_cast<const unsigned int*>(w.bits[rr - 1]);
      }
      p->mask_scale = 1.0f;
    }
    if (g.nprob && (rc = timed_gemm(m, g, 2, rr == 0 ? "dgrad_dec_l1" : "dgrad_dec_l2", st))) return rc;
  }
  // inactive decoders contribute zero to dL/dz: clear their slice of g_d0
  if (any_dec)
    for (size_t i = 0; i < m->decs.size(); ++i)
      if (!active[i] && (rc = chain_flush(m, st))) return rc;
      else if (!active[i])
        CK(cudaMemset2DAsync(m->g_d0 + m->decs[i].cat_off, sizeof(bf16) * m->cat.out, 0, sizeof(bf16) * m->decs[i].cat_w, B, st));
  // Weight-gradient group over the encoder and / or decoder layers (dW = dY^T X, bias gradients by the ones-MMA).
  auto emit_wgrad = [&](bool enc, bool dec, const char* name, cudaStream_t wst, int max_ctas = 0) -> int {
    GemmGroup g; init_group(g);
    int rc2;
    if (enc) {
      for (size_t i = 0; i < m->encs.size(); ++i) {
        if (!(present >> i & 1)) continue;
        const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
        for (size_t r = 0; r < e.fc.size(); ++r) {
          const Lin& l = e.fc[r];
          const bf16* X = r == 0 ? w.x : w.act[r - 1];
          const int ldx = r == 0 ? w.ldx : w.ld_act[r - 1];
          if ((rc2 = add_tn(m, g, w.gpre[r], l.out, X, ldx, l.out, l.in, B, G + l.w_off, l.in, G + l.b_off))) return rc2;
        }
        const Lin& h = e.heads;
        const bf16* X = e.fc.empty() ? w.x : w.act.back();
        const int ldx = e.fc.empty() ? w.ldx : w.ld_act.back();
        if ((rc2 = add_tn(m, g, m->gml, m->ldgml, X, ldx, h.out, h.in, B, G + h.w_off, h.in, G + h.b_off))) return rc2;
        if (e.type == 'C')
          if ((rc2 = add_tn(m, g, w.onehot, w.ld_onehot, w.g_x, w.ld_gx, m->S, m->E, B, G + e.emb_off, m->E, nullptr))) return rc2;
      }
    }
    if (dec && any_dec) {
      // fused first decoder layer: rows of inactive decoders receive zeros (their g_d0 slice was cleared)
      const Lin& c = m->cat;
      if ((rc2 = add_tn(m, g, m->g_d0, c.out, m->z, m->ldz, c.out, c.in, B, G + c.w_off, c.in, G + c.b_off))) return rc2;
      for (size_t i = 0; i < m->decs.size(); ++i) {
        if (!active[i]) continue;
        const Dec& d = m->decs[i]; DecWS& w = m->dws[i];
        for (size_t r = 0; r < d.rest.size(); ++r) {
          const Lin& l = d.rest[r];
          const bool last = r + 1 == d.rest.size();
          const bf16* Gr = last ? w.g_out : w.gact[r];
          const int ldg = last ? w.ld_gout : l.out;
          const bf16* X = r == 0 ? m->d0 + d.cat_off : w.act[r - 1];
          const int ldx = r == 0 ? m->ld_d0 : w.ld_act[r - 1];
          if ((rc2 = add_tn(m, g, Gr, ldg, X, ldx, l.out, l.in, B, G + l.w_off, l.in, G + l.b_off))) return rc2;
        }
      }
    }
    if (!g.nprob) return VLA_OK;
    int force = 0;
    if (max_ctas > 0) {       // side branch: leave SMs to the main chain's launches
      int base = 0;
      for (int i = 0; i < g.nprob; ++i) base += g.p[i].m_tiles * g.p[i].n_tiles;
      force = std::max(1, max_ctas / std::max(base, 1));
    }
    finalize_tn(g, B, force);
    return timed_gemm(m, g, 1, name, wst);
  };
  // Data parallel: the decoder weight gradients need nothing from the encoder backward.  Compute them now and send them
  // (with the loss scalars, which sit behind them in the flat buffer) on the side stream while the encoder backward runs.
  // (with the chain kernel the decoder data gradients sit in the middle of one launch: no fork point, no early exchange)
  // CTAs of a weight-gradient launch that runs beside the main chain (the rest of the SMs stay free for the chain's launches)
  static const int side_ctas = [] { const char* e = getenv("VLA_SIDE_CTAS"); return e ? atoi(e) : 64; }();
  // (data parallel: only where the decoder gradients are not exchanged early -- that path has its own branch below)
  const bool side_dec = (io.dp == nullptr || !dp_overlap(io.dp)) && io.side_dec && any_dec && !m->chain_on && !sfx && !recorder() &&
                        side_ready(m);
  if (side_dec) {
    CK(cudaEventRecord(m->ev_fork, st));
    CK(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
    if ((rc = emit_wgrad(false, true, "wgrad_dec", m->side, side_ctas))) return rc;
    m->side_busy = true;
  }
  const bool early_dec = side_dec || (io.dp != nullptr && any_dec && !m->chain_on && dp_overlap(io.dp));
  if (early_dec && !side_dec) {
    vla_dp* dp = io.dp;
    CK(cudaEventRecord(dp->ev_fork, st));
    CK(cudaStreamWaitEvent(dp->side, dp->ev_fork, 0));
    if ((rc = emit_wgrad(false, true, "wgrad_dec", dp->side))) return rc;      // (full width: the exchange behind it is on the clock)
    if ((rc = run_exchange(m, dp, m->cat.w_off / 2, dp->n / 2, 1, dp->side, false))) return rc;
    CK(cudaEventRecord(dp->ev_join, dp->side));
  }
  int n_present = 0;
  for (size_t i = 0; i < m->encs.size(); ++i) n_present += present >> i & 1;
  const bool use_hb = m->hb_used && !sfx;
  if (use_hb) {
    // ---- head block backward: d(first decoder layer) -> d(mu | logvar) -> d(heads) incl. BatchNorm statistics, one launch ----
    HbArgs hb;
    if (fill_headblock(m, &hb, B, present, P, nullptr, nullptr, train, nullptr, 0, 0, io.engine, 1, nullptr))
      return fail(VLA_ERR_STATE, "head block backward: plan does not fit");
    hb.has_dec = any_dec ? 1 : 0;
    hb.gmu_in = io.g_mu; hb.glv_in = io.g_logvar;
    if (!io.engine) hb.dyn = nullptr;
    double fl = 2.0 * B * m->L * m->cat.out;
    for (int e = 0; e < hb.n_enc; ++e) fl += 2.0 * B * hb.enc[e].in_dim * m->HW;
    { ProfScope ps(m, st, "head_block_bwd", fl, static_cast<double>(B) * (m->cat.out * 2.0 + m->L * 12.0)); CK(launch_head_block_bwd(hb, st)); }
  }
  // The latent backward is row-local arithmetic on dL/dz: with VLA_FUSE_LATBWD=1 it runs in the epilogue of the GEMM that
  // produces dL/dz (GF_LATBWD, one launch less).  OPT-IN: measured slower at batch 4096 (profiles/r2_latbwd_fusion.md: rna2dna
  // 100.6 -> 111.9 us with the side branch, 102.6 -> 105.0 us without) -- the epilogue's operand loads and exponentials do not
  // hide behind a two-k-block main loop, and the next GEMM then competes for SMs with the side branch's weight gradients.
  const bool fuse_latbwd_on = [] { const char* e = getenv("VLA_FUSE_LATBWD"); return e && e[0] == '1'; }();   // (read per call: a
                                                                  // graph step issues its launches once, at capture)
  const bool fuse_latbwd = fuse_latbwd_on && any_dec && !sfx && !use_hb && !m->chain_on;
  if (any_dec && !sfx && !use_hb) {
    GemmGroup g; init_group(g); GemmProblem* p;
    const Lin& l = m->cat;
    if ((rc = add_nn(m, g, m->g_d0, l.out, SH + l.sh_off, l.sh_ld, B, l.in, l.out, fuse_latbwd ? GF_LATBWD : GF_OUT_F32, &p))) return rc;
    if (fuse_latbwd) {
      p->aux0 = m->mu; p->aux1 = m->logvar; p->pre = m->eps; p->mean = io.g_mu; p->rstd = io.g_logvar;
      p->out_bf16 = m->gml; p->ld_bf16 = m->ldgml;
      p->aux_n = n_present; p->loss_kind = m->ae ? 1 : 0;
      p->dyn = io.engine ? m->dyn : nullptr; p->aux_scale = 0.f;
    } else {
      p->out_f32 = m->gz; p->ld_f32 = L;
    }
    if ((rc = timed_gemm(m, g, 2, "dgrad_dec_l0", st))) return rc;
  }
  // ---- latent ----
  if (!sfx && !use_hb && !fuse_latbwd) {
    LatentBwdArgs a{};
    a.gz = any_dec ? m->gz : nullptr; a.ld_gz = L;
    a.gmu_in = io.g_mu; a.glv_in = io.g_logvar;
    a.mu = m->mu; a.logvar = m->logvar; a.eps = m->eps;
    a.beta = 0.f; a.dyn = io.engine ? m->dyn : nullptr;
    a.n_modalities = n_present; a.gml = m->gml; a.ld_gml = m->ldgml; a.rows = B; a.L = L; a.ae = m->ae ? 1 : 0;
    if (m->chain_on) { if ((rc = chain_add(m, st, CK_LATENT_BWD, -1, &a, sizeof(a), "latent_bwd", 0, static_cast<double>(B) * L * 20.0, false))) return rc; }
    else { ProfScope ps(m, st, "latent_bwd", 0, static_cast<double>(B) * L * 20.0); CK(launch_latent_bwd(a, st)); }
  }
  const int mt = ceil_div(B, GEMM_BM);
  // ---- encoder data gradients ----
  size_t max_depth = 0;
  for (size_t i = 0; i < m->encs.size(); ++i) if (present >> i & 1) max_depth = std::max(max_depth, m->encs[i].fc.size());
  bool site_done = false;
  for (size_t r = max_depth; r >= 1; --r) {
    GemmGroup g; init_group(g);
    std::vector<std::pair<size_t, size_t>> bn_todo;   // (encoder, layer) whose gy this round produces
    for (size_t i = 0; i < m->encs.size(); ++i) {
      if (!(present >> i & 1)) continue;
      const Enc& e = m->encs[i]; EncWS& w = m->ews[i];
      GemmProblem* p;
      if (e.type == 'C') {
        if (site_done || use_hb) { site_done = true; continue; }
        const Lin& l = e.heads;
        if ((rc = add_nn(m, g, m->gml, m->ldgml, SH + l.sh_off, l.sh_ld, B, l.in, l.out, GF_OUT_BF16, &p))) return rc;
        p->out_bf16 = w.g_x; p->ld_bf16 = w.ld_gx;
        site_done = true;
        continue;
      }
      const size_t depth = e.fc.size();
      if (r > depth) continue;
      if (use_hb && r == depth) { bn_todo.emplace_back(i, r - 1); continue; }      // (the head block wrote gy and the statistics)
      const Lin& l = r == depth ? e.heads : e.fc[r];
      const bf16* A = r == depth ? m->gml : w.gpre[r];
      const int lda = r == depth ? m->ldgml : l.out;
      const size_t tgt = r - 1;                       // gradient w.r.t. act[tgt]
      if ((rc = add_nn(m, g, A, lda, SH + l.sh_off, l.sh_ld, B, l.in, l.out, GF_MASK | GF_BNSTATS | GF_OUT_BF16, &p))) return rc;
      p->mask_src = w.act[tgt]; p->ld_mask = w.ld_act[tgt]; p->mask_scale = train ? 1.0f / 0.9f : 1.0f;
      p->pre = w.pre[tgt]; p->ld_pre = l.in; p->mean = w.mean[tgt]; p->rstd = w.rstd[tgt];
      p->stats = w.bstats[tgt];
      p->out_bf16 = w.gy[tgt]; p->ld_bf16 = l.in;
      bn_todo.emplace_back(i, tgt);
    }
    if (g.nprob && !sfx && (rc = timed_gemm(m, g, 2, r == 1 ? "dgrad_enc_l1" : "dgrad_enc_l2", st))) return rc;
    std::vector<BnBwdArgs> round_bb;
    for (auto& it : bn_todo) {
      const Enc& e = m->encs[it.first]; EncWS& w = m->ews[it.first]; const Bn& bn = e.bn[it.second];
      BnBwdArgs a{};
      a.gy = w.gy[it.second]; a.ld_gy = bn.n; a.pre = w.pre[it.second]; a.ld_pre = bn.n;
      // partial statistics per 128-row GEMM tile, or per 32-row block when the head block produced them
      a.stats = w.bstats[it.second]; a.m_tiles = (use_hb && it.second + 1 == e.fc.size()) ? ceil_div(B, HB_ROWS) : mt;
      if (io.sync_bn && io.dp && train) {
        if ((rc = run_stats_allreduce(m, io.dp, w.bstats[it.second], a.m_tiles, bn.n, "syncbn_bwd", st))) return rc;
        a.m_tiles = 1; a.stat_rows = B * io.dp->world; a.param_grad_scale = 1.0f / io.dp->world;
      }
      a.mean = w.mean[it.second]; a.rstd = w.rstd[it.second]; a.gamma = P + bn.g_off;
      a.dgamma = G + bn.g_off; a.dbeta = G + bn.b_off;
      a.gpre = w.gpre[it.second]; a.ld_gpre = bn.n; a.rows = B; a.n = bn.n; a.train = train;
      if (m->chain_on) { if ((rc = chain_add(m, st, CK_BN_BWD, -1, &a, sizeof(a), "bn_bwd", 0, 8.0 * B * bn.n, train != 0))) return rc; }
      else round_bb.push_back(a);
    }
    if (round_bb.size() == 2 && !recorder()) {
      ProfScope ps(m, st, "bn_bwd", 0, 8.0 * B * (round_bb[0].n + round_bb[1].n));
      CK(launch_bn_bwd_pair(round_bb[0], round_bb[1], st));
    } else {
      for (const BnBwdArgs& a : round_bb) { ProfScope ps(m, st, "bn_bwd", 0, 8.0 * B * a.n); CK(launch_bn_bwd(a, st)); }
    }
  }
  if (!site_done && !sfx && !use_hb) {
    for (size_t i = 0; i < m->encs.size(); ++i) {
      if (!(present >> i & 1) || m->encs[i].type != 'C') continue;
      GemmGroup g; init_group(g); GemmProblem* p;
      const Lin& l = m->encs[i].heads; EncWS& w = m->ews[i];
      if ((rc = add_nn(m, g, m->gml, m->ldgml, SH + l.sh_off, l.sh_ld, B, l.in, l.out, GF_OUT_BF16, &p))) return rc;
      p->out_bf16 = w.g_x; p->ld_bf16 = w.ld_gx;
      if ((rc = timed_gemm(m, g, 2, "dgrad_site", st))) return rc;
    }
  }
  // ---- weight (and bias) gradients: one grouped split-K launch (data parallel: the encoder part; the decoder part went early) ----
  if (early_dec) return emit_wgrad(true, false, "wgrad_enc", st);
  return emit_wgrad(true, true, "wgrad_all", st);
}

// ---------------------------------------------------------------------------------------------
// Chain: collect the row-local launches of a call, issue each stretch as one chain_kernel launch
// ---------------------------------------------------------------------------------------------
int chain_add(vla_model* m, cudaStream_t st, int kind, int gemm_mode, const void* args, size_t size, const char* name,
              double flops, double bytes, bool needs_all) {
  int rc;
  if (needs_all && (rc = chain_flush(m, st))) return rc;
  if (static_cast<int>(m->seg.size()) >= CHAIN_MAX_PHASES && (rc = chain_flush(m, st))) return rc;
  if (kind == CK_LATENT_FWD && static_cast<const LatentFwdArgs*>(args)->dyn && !static_cast<const LatentFwdArgs*>(args)->eps_in)
    for (const ChainOp& o : m->seg)      // the Philox offset reads dyn->step: never in the launch whose ingest phase bumps it
      if (o.kind == CK_INGEST && reinterpret_cast<const IngestArgs*>(o.args.data())->bump_step) { if ((rc = chain_flush(m, st))) return rc; break; }
  ChainOp op;
  op.kind = kind; op.gemm_mode = gemm_mode; op.name = name; op.flops = flops; op.bytes = bytes;
  op.args.assign(static_cast<const char*>(args), static_cast<const char*>(args) + size);
  m->seg.push_back(std::move(op));
  return VLA_OK;
}

int launch_op(vla_model* m, const ChainOp& op, cudaStream_t st) {
  ProfScope ps(m, st, op.name.c_str(), op.flops, op.bytes);
  cudaError_t e = cudaSuccess;
  switch (op.kind) {
    case CK_INGEST: e = launch_ingest(*reinterpret_cast<const IngestArgs*>(op.args.data()), st); break;
    case CK_BN_ACT: e = launch_bn_act(*reinterpret_cast<const BnActArgs*>(op.args.data()), st); break;
    case CK_BN_BWD: e = launch_bn_bwd(*reinterpret_cast<const BnBwdArgs*>(op.args.data()), st); break;
    // (KL partials in the chain's layout: one per 32-row slice -- the loss tail was sized for that)
    case CK_LATENT_FWD: e = launch_latent_fwd_rows(*reinterpret_cast<const LatentFwdArgs*>(op.args.data()), st); break;
    case CK_LATENT_BWD: e = launch_latent_bwd(*reinterpret_cast<const LatentBwdArgs*>(op.args.data()), st); break;
    default: {
      // GemmGroup holds tensor maps (64-byte alignment): copy out of the byte vector
      GemmGroup g;
      memcpy(&g, op.args.data(), sizeof(g));
      e = launch_gemm_group(g, op.gemm_mode, st);
    }
  }
  if (e != cudaSuccess) return fail(VLA_ERR_CUDA, "launch (" + op.name + "): " + cudaGetErrorString(e));
  return VLA_OK;
}

int op_rows(const ChainOp& op) {
  switch (op.kind) {
    case CK_INGEST: return reinterpret_cast<const IngestArgs*>(op.args.data())->rows;
    case CK_BN_ACT: return reinterpret_cast<const BnActArgs*>(op.args.data())->rows;
    case CK_BN_BWD: return reinterpret_cast<const BnBwdArgs*>(op.args.data())->rows;
    case CK_LATENT_FWD: return reinterpret_cast<const LatentFwdArgs*>(op.args.data())->rows;
    case CK_LATENT_BWD: return reinterpret_cast<const LatentBwdArgs*>(op.args.data())->rows;
    default: return reinterpret_cast<const GemmGroup*>(op.args.data())->p[0].M;
  }
}

// Returns 1 when this stretch cannot run as a chain (the caller issues the separate launches), 0 on success, < 0 on error.
int launch_stretch(vla_model* m, cudaStream_t st, const std::vector<ChainOp>& ops) {
  auto up64 = [](size_t x) { return (x + 63) & ~size_t(63); };
  ChainPlan pl{};
  pl.n_phases = static_cast<int>(ops.size());
  pl.rows = op_rows(ops[0]);
  pl.m_blocks = ceil_div(pl.rows, CHAIN_ROWS);
  std::vector<char> img(up64(sizeof(ChainPlan)), 0);
  double flops = 0, bytes = 0;
  bool has_ingest = false, has_latent = false, has_loss = false, has_bwd = false;
  for (size_t i = 0; i < ops.size(); ++i) {
    const ChainOp& op = ops[i];
    if (op_rows(op) != pl.rows) return 1;
    ChainPhase& ph = pl.ph[i];
    ph.kind = op.kind;
    ph.args_off = static_cast<long long>(img.size());
    img.resize(up64(img.size() + op.args.size()), 0);
    memcpy(img.data() + ph.args_off, op.args.data(), op.args.size());
    if (op.kind <= CK_GEMM_LAST) {
      GemmGroup g;
      memcpy(&g, op.args.data(), sizeof(g));
      std::vector<int> units[CHAIN_CLUSTER];
      if (chain_assign(g, nullptr, &units) < 0) return 1;
      for (int r = 0; r < CHAIN_CLUSTER; ++r) {
        ph.n_units[r] = static_cast<int>(units[r].size());
        for (size_t u = 0; u < units[r].size(); ++u) ph.units[r][u] = static_cast<unsigned short>(units[r][u]);
      }
      for (int k = 0; k < g.nprob; ++k) if (g.p[k].flags & GF_LOSS) has_loss = true;
      if (op.gemm_mode == 2) has_bwd = true;
    }
    has_ingest = has_ingest || op.kind == CK_INGEST;
    has_latent = has_latent || op.kind == CK_LATENT_FWD;
    has_bwd = has_bwd || op.kind == CK_BN_BWD || op.kind == CK_LATENT_BWD;
    flops += op.flops; bytes += op.bytes;
  }
  memcpy(img.data(), &pl, sizeof(pl));
  ChainPlanCached* c = nullptr;
  for (ChainPlanCached* q : m->plans) if (q->key == img) { c = q; break; }
  if (!c) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CK(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone)
      return fail(VLA_ERR_STATE, "the first call for a new argument set builds the chain plan (device allocation + upload) and "
                                 "cannot run inside a stream capture; make the same call once outside the capture");
    if (m->plans.size() >= 64) return 1;       // callers with ever-changing buffers: separate launches from here on
    if (m->chain_clusters == 0) {
      cudaError_t e = cudaSuccess;
      m->chain_clusters = chain_max_clusters(&e);
      if (e != cudaSuccess) { (void)cudaGetLastError(); m->chain_clusters = -1; }
    }
    if (m->chain_clusters <= 0) return 1;
    c = new ChainPlanCached();
    c->key = img;
    c->dbg_off = (img.size() + 255) & ~size_t(255);
    const size_t dbg_bytes = sizeof(unsigned long long) * 8 * CHAIN_MAX_PHASES * CHAIN_CLUSTER * static_cast<size_t>(m->chain_clusters);
    cudaError_t e = cudaMalloc(&c->dev, c->dbg_off + dbg_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->dev + c->dbg_off, 0, dbg_bytes);
    if (e == cudaSuccess) {
      ChainPlan up = pl;
      up.dbg = m->chain_dbg ? reinterpret_cast<unsigned long long*>(c->dev + c->dbg_off) : nullptr;
      std::vector<char> img2 = img;
      memcpy(img2.data(), &up, sizeof(up));
      e = cudaMemcpy(c->dev, img2.data(), img2.size(), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { cudaFree(c->dev); delete c; return fail(VLA_ERR_CUDA, std::string("chain plan upload: ") + cudaGetErrorString(e)); }
    c->n_clusters = std::min(pl.m_blocks, m->chain_clusters);
    c->n_phases = pl.n_phases;
    c->flops = flops; c->bytes = bytes;
    c->name = has_ingest && !has_latent ? "chain_ingest_enc"
              : (has_latent ? (has_loss ? "chain_fwd_bwd" : (has_ingest ? "chain_forward" : "chain_enc_dec")) : (has_bwd ? "chain_bwd_" : "chain_") + ops[0].name);
    for (const ChainOp& op : ops) c->phase_names.push_back(op.name);
    m->plans.push_back(c);
  }
  if (m->n_last_plans < 8) m->last_plans[m->n_last_plans++] = c;
  ProfScope ps(m, st, c->name.c_str(), c->flops, c->bytes);
  cudaError_t e = launch_chain(reinterpret_cast<const ChainPlan*>(c->dev), c->n_clusters, st);
  if (e != cudaSuccess) return fail(VLA_ERR_CUDA, std::string("chain kernel launch (") + c->name + "): " + cudaGetErrorString(e));
  return VLA_OK;
}

int chain_flush(vla_model* m, cudaStream_t st) {
  if (m->seg.empty()) return VLA_OK;
  std::vector<ChainOp> ops;
  ops.swap(m->seg);
  if (ops.size() >= 2) {
    const int rc = launch_stretch(m, st, ops);
    if (rc <= 0) return rc;
  }
  for (const ChainOp& op : ops) { const int rc = launch_op(m, op, st); if (rc) return rc; }
  return VLA_OK;
}

// When the row-local stretches of a call run as chain launches (chain_kernel.cu).  The chain plans are device images keyed
// by the argument structs, built on a stretch's first eager execution: callers whose buffers keep their addresses (the
// Trainer, graph-captured inference) hit the cache from the second call on.  VLA_CHAIN=0 switches it off everywhere.
bool chain_enabled() {      // read per call: a host-side switch, not on the replay path
  // Measured (profiles/r2_chain_timeline_v1.log, rna2dna batch 4096): 192 us / step chained vs 123 us as separate launches --
  // every phase pays ~3 us until its first operands land, 2-16 us of epilogue and ~1.4 us of barrier.  Opt-in until the
  // per-phase latency work lands (VLA_CHAIN=1).
  const char* e = getenv("VLA_CHAIN");
  return e && e[0] == '1';
}
struct ChainScope {      // sets m->chain_on for one entry-point call; the stretch still open at the end is issued by finish()
  vla_model* m; bool prev;
  ChainScope(vla_model* m_, bool on) : m(m_), prev(m_->chain_on) { m->chain_on = on; if (on) m->n_last_plans = 0; }
  int finish(cudaStream_t st) { const int rc = chain_flush(m, st); m->chain_on = prev; return rc; }
  ~ChainScope() { m->seg.clear(); m->chain_on = prev; }
};

cudaStream_t as_stream(vla_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace

// =============================================================================================
// extern "C"
// =============================================================================================
extern "C" {

const char* vla_last_error(void) { return g_err.c_str(); }
int vla_abi_version(void) { return 1; }

int vla_model_create(const vla_config_t* cfg, vla_model_t** out) {
  if (!cfg || !out) return fail(VLA_ERR_INVALID, "null argument");
  if (cfg->dim_a < 1 || cfg->dim_b < 1 || cfg->n_sites < 1 || cfg->latent < 1 || cfg->embed < 1)
    return fail(VLA_ERR_INVALID, "dimensions must be positive");
  if (2 * cfg->latent > 256) return fail(VLA_ERR_INVALID, "latent_dim > 128 is not supported");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(VLA_ERR_CUDA, std::string("libvla_b200 is built for sm_100a only; device is sm_") +
                                  std::to_string(prop.major) + std::to_string(prop.minor));
  vla_model* m = new vla_model();
  m->cfg = *cfg;
  { const char* e = getenv("VLA_SPLIT"); m->split = !(e && e[0] == '0'); }
  int rc = build_layout(m);
  if (rc) { delete m; return rc; }
  auto bail = [&](cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    vla_model_destroy(m);
    return fail(VLA_ERR_CUDA, msg);
  };
  cudaError_t e;
  if ((e = cudaMalloc(&m->shadow, sizeof(bf16) * m->n_shadow)) != cudaSuccess) return bail(e, "cudaMalloc shadow");
  if ((e = cudaMemset(m->shadow, 0, sizeof(bf16) * m->n_shadow)) != cudaSuccess) return bail(e, "cudaMemset shadow");
  if ((e = cudaMalloc(&m->chunks_d, sizeof(AdamChunk) * m->chunks_h.size())) != cudaSuccess) return bail(e, "cudaMalloc chunks");
  if ((e = cudaMemcpy(m->chunks_d, m->chunks_h.data(), sizeof(AdamChunk) * m->chunks_h.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "copy chunks");
  if ((e = cudaMalloc(&m->dyn, sizeof(DynParams))) != cudaSuccess) return bail(e, "cudaMalloc dyn");
  DynParams d{5e-4f, 1e-5f, 1e-3f, 1.0f, 0, 0, 0, 0, 1.0, 1.0};
  if ((e = cudaMemcpy(m->dyn, &d, sizeof(d), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "copy dyn");
  if ((e = cudaMalloc(&m->loss_counter, 64)) != cudaSuccess) return bail(e, "cudaMalloc counter");
  if ((e = cudaMemset(m->loss_counter, 0, 64)) != cudaSuccess) return bail(e, "memset counter");
  m->loss_out = reinterpret_cast<float*>(m->loss_counter) + 4;
  *out = m;
  return VLA_OK;
}

int vla_model_create_layout_only(const vla_config_t* cfg, vla_model_t** out) {
  if (!cfg || !out) return fail(VLA_ERR_INVALID, "null argument");
  if (cfg->dim_a < 1 || cfg->dim_b < 1 || cfg->n_sites < 1 || cfg->latent < 1 || cfg->embed < 1)
    return fail(VLA_ERR_INVALID, "dimensions must be positive");
  vla_model* m = new vla_model();
  m->cfg = *cfg;
  int rc = build_layout(m);
  if (rc) { delete m; return rc; }
  m->layout_only = true;
  *out = m;
  return VLA_OK;
}

void vla_model_destroy(vla_model_t* m) {
  if (!m) return;
  if (m->layout_only) { delete m; return; }
  free_plans(m);
  for (GroupPlanCached* g : m->group_plans) { cudaFree(g->dev); delete g; }
  delete m->adam_hints;
  cudaFree(m->rc_dbg); delete m->rc_last;
  if (m->side) cudaStreamDestroy(m->side);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  cudaFree(m->shadow); cudaFree(m->chunks_d); cudaFree(m->dyn); cudaFree(m->loss_counter);
  cudaFree(m->ws);
  delete m;
}

int vla_model_reserve(vla_model_t* m, int batch) { return m ? reserve(m, batch) : fail(VLA_ERR_INVALID, "null model"); }
long long vla_param_count(const vla_model_t* m) { return m->n_params; }
long long vla_buffer_count(const vla_model_t* m) { return m->n_buffers; }
int vla_counter_count(const vla_model_t* m) { return m->n_bn; }
int vla_num_tensors(const vla_model_t* m) { return static_cast<int>(m->infos.size()); }
int vla_tensor_info(const vla_model_t* m, int index, vla_tensor_info_t* out) {
  if (!m || !out || index < 0 || index >= static_cast<int>(m->infos.size())) return fail(VLA_ERR_INVALID, "bad tensor index");
  *out = m->infos[index];
  return VLA_OK;
}

int vla_forward(vla_model_t* m, const vla_forward_args_t* a, vla_stream_t stream) {
  if (!m || !a || !a->params || !a->buffers) return fail(VLA_ERR_INVALID, "null argument");
  cudaStream_t st = as_stream(stream);
  int rc;
  if (a->refresh_shadows && (rc = run_shadow_refresh(m, a->params, st))) return rc;
  FwdIO io{};
  io.params = a->params; io.buffers = a->buffers; io.counters = a->counters;
  io.x[0] = a->x_a; io.x[1] = a->x_b; io.site = a->site;
  io.batch = a->batch; io.train = a->train ? 1 : 0;
  io.eps = a->eps; io.keep_masks = a->keep_masks; io.seed = a->seed; io.offset = a->offset;
  io.recon[0] = a->recon_a; io.recon[1] = a->recon_b; io.recon[2] = a->recon_c;
  io.mu = a->mu; io.logvar = a->logvar; io.engine = false;
  // Large batches (inference sweeps, BASELINE configs[3]): row-local stretches as chain launches -- in eval mode the whole
  // forward is ONE launch.  Small per-call batches (the scripts' batch 32) keep the separate launches: their output tensors
  // change address from call to call, which would rebuild the chain plan every time.
  ChainScope cs(m, chain_enabled() && a->batch >= 1024);
  if ((rc = run_forward(m, io, st))) return rc;
  return cs.finish(st);
}

int vla_refresh_shadows(vla_model_t* m, const float* params, vla_stream_t stream) {
  if (!m || !params) return fail(VLA_ERR_INVALID, "null argument");
  return run_shadow_refresh(m, params, as_stream(stream));
}

int vla_backward(vla_model_t* m, const vla_backward_args_t* a, vla_stream_t stream) {
  if (!m || !a || !a->params || !a->grads) return fail(VLA_ERR_INVALID, "null argument");
  BwdIO io{};
  io.params = a->params;
  io.g_recon[0] = a->g_recon_a; io.g_recon[1] = a->g_recon_b; io.g_recon[2] = a->g_recon_c;
  io.recon_b = a->recon_b; io.g_mu = a->g_mu; io.g_logvar = a->g_logvar;
  io.grads = a->grads; io.engine = false; io.zero_grads = true;
  ChainScope cs(m, chain_enabled() && m->saved_batch >= 1024);
  int rc = run_backward(m, io, as_stream(stream));
  if (rc) return rc;
  return cs.finish(as_stream(stream));
}

long long vla_loss_workspace_bytes(int batch, int dim_a, int dim_b, int n_sites, int latent) {
  (void)latent;
  return 256 + 4LL * loss_grid_size(batch, dim_a, dim_b, n_sites);
}

int vla_loss(const vla_loss_args_t* a, vla_stream_t stream) {
  if (!a || !a->out || !a->workspace) return fail(VLA_ERR_INVALID, "null argument");
  if (a->batch <= 0) return fail(VLA_ERR_INVALID, "batch must be positive");
  LossArgs l{};
  l.recon_a = a->recon_a; l.a = a->a; l.width_a = a->dim_a;
  l.recon_b = a->recon_b; l.b = a->b; l.width_b = a->dim_b;
  l.logits = a->recon_c; l.site = a->site; l.class_w = a->class_weights; l.n_sites = a->n_sites;
  if ((l.recon_a && !l.a) || (l.recon_b && !l.b) || (l.logits && !l.site)) return fail(VLA_ERR_INVALID, "recon without target");
  l.mu = a->mu; l.logvar = a->logvar; l.L = a->latent;
  l.beta = a->beta; l.gamma = a->gamma; l.rows = a->batch;
  l.ga_f32 = a->g_recon_a; l.gb_f32 = a->g_recon_b; l.gc_f32 = a->g_recon_c; l.gmu_f32 = a->g_mu; l.glv_f32 = a->g_logvar;
  l.grad_scale = 1.0f;
  l.counter = reinterpret_cast<unsigned int*>(a->workspace);
  l.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) + 256);
  l.out = a->out;
  CK(launch_loss(l, as_stream(stream)));
  return VLA_OK;
}

static int run_adamw(vla_model_t* m, float* p, const float* g, float* ea, float* eas, float lr, float b1, float b2,
                     float eps, float wd, int step, bool dyn, bool zero_grad, cudaStream_t st, vla_dp* dp = nullptr,
                     const DpArgs* fused = nullptr, int chunk0 = 0, int n_chunks = -1, const char* name = "adamw") {
  AdamArgs a{};
  a.p = p; a.g = const_cast<float*>(g); a.m = ea; a.v = eas; a.shadow = m->shadow;
  a.gclear = a.g;
  if (dp) {
    a.gframed = reinterpret_cast<const uint4*>(dp->base + dp->off_rsum);
    a.tail2 = m->n_params / 2;
    a.sums_out = reinterpret_cast<float*>(dp->local + dp->off_sums);
  }
  a.chunks = m->chunks_d + chunk0; a.n_chunks = n_chunks >= 0 ? n_chunks : static_cast<int>(m->chunks_h.size()) - chunk0;
  a.lr = lr; a.beta1 = b1; a.beta2 = b2; a.eps = eps; a.weight_decay = wd;
  if (step > 0) {
    a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), step));
    a.inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), step)));
  }
  a.dyn = dyn ? m->dyn : nullptr; a.update = 1; a.zero_grad = zero_grad ? 1 : 0;
  if (m->deferred_tail_valid && chunk0 == 0) {        // (the launch that holds block 0 of the step's optimizer)
    if (dp || fused) return fail(VLA_ERR_STATE, "deferred loss reduction under data parallelism");
    a.tail = m->deferred_tail; a.has_tail = 1;
    m->deferred_tail_valid = false;
  }
  { int rcf = chain_flush(m, st); if (rcf) return rcf; }
  if (fused) {      // exchange of `fused`'s range + AdamW as one launch (dp_exchange.cu)
    ProfScope ps(m, st, "dp_exchange_adamw", 0, 34.0 * m->n_params + dp_bytes(*fused)); CK(launch_dp_adamw(*fused, a, st));
    return VLA_OK;
  }
  double n_el = 0;
  for (int i = 0; i < a.n_chunks; ++i) n_el += m->chunks_h[chunk0 + i].n;
  // chunk offsets as a kernel parameter for the whole-arena launch of a single model (one round of loads in the kernel)
  const AdamHints* hints = nullptr;
  if (chunk0 == 0 && a.n_chunks == static_cast<int>(m->chunks_h.size()) && a.n_chunks <= ADAM_HINT_CHUNKS && !dp && a.update) {
    if (!m->adam_hints) {
      m->adam_hints = new AdamHints();
      memset(m->adam_hints, 0, sizeof(AdamHints));
      for (int i = 0; i < a.n_chunks; ++i) m->adam_hints->off4[i] = static_cast<unsigned int>(m->chunks_h[i].offset / 4);
      m->adam_hints->n = a.n_chunks;
      m->adam_hints->arena_elems = m->n_params;
    }
    hints = m->adam_hints;
  }
  { ProfScope ps(m, st, name, 0, 34.0 * n_el); CK(launch_adamw(a, st, hints)); }
  return VLA_OK;
}

int vla_adamw(vla_model_t* m, const vla_adamw_args_t* a, vla_stream_t stream) {
  if (!m || !a || !a->params || !a->grads || !a->exp_avg || !a->exp_avg_sq) return fail(VLA_ERR_INVALID, "null argument");
  if (a->step < 1) return fail(VLA_ERR_INVALID, "step is 1-based");
  return run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, a->lr, a->beta1, a->beta2, a->eps, a->weight_decay,
                   a->step, false, false, as_stream(stream));
}

int vla_set_hyper(vla_model_t* m, float lr, float weight_decay, float beta_kl, float gamma, vla_stream_t stream) {
  if (!m) return fail(VLA_ERR_INVALID, "null model");
  const float h[4] = {lr, weight_decay, beta_kl, gamma};
  CK(cudaMemcpyAsync(m->dyn, h, sizeof(h), cudaMemcpyHostToDevice, as_stream(stream)));
  return VLA_OK;
}
int vla_set_step(vla_model_t* m, int completed_steps, int batch_index, float beta1, float beta2, vla_stream_t stream) {
  if (!m) return fail(VLA_ERR_INVALID, "null model");
  const int v[2] = {completed_steps, batch_index};
  CK(cudaMemcpyAsync(&m->dyn->step, v, sizeof(v), cudaMemcpyHostToDevice, as_stream(stream)));
  const double pw[2] = {pow(static_cast<double>(beta1), completed_steps), pow(static_cast<double>(beta2), completed_steps)};
  CK(cudaMemcpyAsync(&m->dyn->b1pow, pw, sizeof(pw), cudaMemcpyHostToDevice, as_stream(stream)));
  return VLA_OK;
}

// ---------------------------------------------------------------------------------------------
// Row-chain step (rowchain.cu): plan of the on-chip middle of a directional model's train step
// ---------------------------------------------------------------------------------------------
// Which steps qualify: one dense encoder with ONE BatchNorm layer of at most 128 columns, the site encoder, one decoder
// with two layers behind the fused first one (RNA2DNAVAE / RNA2DNAAE: encoders.py:8-24, 46-61; decoders.py:21-35), latent
// <= 64, hidden widths <= 512, output <= 576 columns, loss fused (MSE or BCE).  Everything else keeps the separate launches.
// Opt-in (VLA_ROWCHAIN=1).  Measured at batch 4096 (profiles/r2_rowchain_timeline_*.log): parity with the separate launches
// (tests/test_gpu_rowchain.py) but 160 us for what the separate launches do in ~75 us -- one CTA per row block leaves 32 SMs
// doing 11 layers strictly one after the other: 47 us of issue loops that run at 0.33 us per 16 KB weight tile whatever the
// tile's work (producer <-> MMA mbarrier hand-shake + tcgen05.commit; with neither loads nor MMAs still 0.2 us) and ~100 us of
// epilogues at 8 warps per SM.  The separate launches spread the same layers over 128 CTAs each.
bool rowchain_enabled() {
  const char* e = getenv("VLA_ROWCHAIN");
  return e && e[0] == '1';
}
bool rowchain_fits(const vla_model* m, int present) {
  if (m->encs.size() != 2 || m->decs.size() != 1) return false;
  const Enc& ea = m->encs[0]; const Enc& es = m->encs[1]; const Dec& d = m->decs[0];
  if (ea.type == 'C' || es.type != 'C' || present != 3) return false;
  if (ea.fc.size() != 1 || ea.fc[0].out > 128 || (ea.fc[0].out & 31)) return false;
  if (d.type == 'C' || d.rest.size() != 2) return false;
  if (m->L > 64 || m->E > 64 || m->HW > 128) return false;
  if (m->cat.out > 256 || (m->cat.out & 63) || d.rest[0].out > 512 || (d.rest[0].out & 63) || d.rest[1].out > 576) return false;
  return true;
}

struct RcBuilder {
  vla_model* m; RcPlan* pl; int rc = VLA_OK;
  int tm_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer) {
    if (rc || pl->n_tm >= RC_MAX_TMAPS) { if (!rc) rc = fail(VLA_ERR_STATE, "row-chain plan: too many tensor maps"); return 0; }
    rc = get_tmap(m, &pl->tm[pl->n_tm], base, inner, outer, pitch_bytes, box_outer);
    return pl->n_tm++;
  }
  int tm_f32(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes) {
    if (rc || pl->n_tm >= RC_MAX_TMAPS) { if (!rc) rc = fail(VLA_ERR_STATE, "row-chain plan: too many tensor maps"); return 0; }
    std::string err;
    if (!make_tmap_f32_tile(&pl->tm[pl->n_tm], base, inner, outer, pitch_bytes, &err)) rc = fail(VLA_ERR_CUDA, err);
    return pl->n_tm++;
  }
  RcOp& op(int kind, int sub = 0) {
    static RcOp dummy;
    if (pl->n_ops >= RC_MAX_OPS) { if (!rc) rc = fail(VLA_ERR_STATE, "row-chain plan: too many ops"); return dummy; }
    RcOp& o = pl->ops[pl->n_ops++];
    memset(&o, 0, sizeof(o));
    o.kind = kind; o.sub = sub; o.a_lo_slot = -1; o.out_lo_slot = -1; o.tm_side = -1; o.fscale = 1.0f;
    return o;
  }
  // weight of a Linear as the K-major B operand of its forward GEMM (box 64 x 128) / the MN-major one of its data gradient
  int w_nt(const Lin& l) { return tm_bf16(m->shadow + l.sh_off, static_cast<uint64_t>(l.sh_lo > 0 ? l.sh_lo + l.in : l.in), l.out, static_cast<uint64_t>(l.sh_ld) * 2, 128); }
  int w_nn(const Lin& l) { return tm_bf16(m->shadow + l.sh_off, l.in, l.out, static_cast<uint64_t>(l.sh_ld) * 2, 64); }
  // activation [B, width] as a store / load target in blocks of [128 x 64]
  int act(const void* base, int width, int ld, int B) { return tm_bf16(base, width, B, static_cast<uint64_t>(ld) * 2, 128); }
};

int build_rowchain_plan(vla_model* m, const FwdIO& io, RcPlan* pl) {
  memset(static_cast<void*>(pl), 0, sizeof(*pl));
  const int B = io.batch, L = m->L, HW = m->HW, E = m->E;
  const float* P = io.params;
  const Enc& ea = m->encs[0]; const Enc& es = m->encs[1]; const Dec& d = m->decs[0];
  EncWS& wa = m->ews[0]; EncWS& wsx = m->ews[1]; DecWS& wd = m->dws[0];
  const Lin& l1 = d.rest[0]; const Lin& l2 = d.rest[1];
  const int H = ea.fc[0].out;                     // BatchNorm width (<= 128)
  const int hk = ceil_div(H, 64);                 // k-blocks of the encoder activation
  const bool sp = m->split;
  RcBuilder b{m, pl};
  pl->rows = B; pl->m_blocks = ceil_div(B, GEMM_BM); pl->n_batches = io.n_batches; pl->L = L; pl->ae = m->ae ? 1 : 0; pl->n_enc = 2;
  pl->dyn = m->dyn; pl->dyn_bump = m->dyn;
  const Bn& bn = ea.bn[0];
  pl->bn_stats = wa.stats[0]; pl->bn_gamma = P + bn.g_off; pl->bn_beta = P + bn.b_off;
  pl->bn_running_mean = io.buffers + bn.rm_off; pl->bn_running_var = io.buffers + bn.rv_off;
  pl->bn_nbt = io.counters ? io.counters + bn.counter : nullptr;
  pl->bn_save_mean = wa.mean[0]; pl->bn_save_rstd = wa.rstd[0];
  pl->bn_keep = io.keep_masks ? io.keep_masks[ea.first_drop] : nullptr;
  pl->bn_n = H; pl->bn_m_tiles = pl->m_blocks; pl->train = io.train; pl->p_drop = 0.1f;
  pl->seed = io.seed; pl->bn_offset = io.offset * 16 + 1 + ea.first_drop; pl->lat_offset = io.offset * 16;
  pl->loss_partials = m->eloss_partials; pl->kl_partials = m->kl_partials; pl->counter = m->loss_counter; pl->loss_out = io.loss_out;
  pl->loss_kind = d.type == 'A' ? LOSS_MSE : LOSS_BCE;
  const float* tgt = d.type == 'A' ? io.tgt_a : io.tgt_b;
  const long long tgt_rows = static_cast<long long>(B) * std::max(io.n_batches, 1);

  const int tm_pre = b.tm_f32(wa.pre[0], H, B, static_cast<uint64_t>(H) * 4);
  const int tm_tgt = b.tm_f32(tgt, d.out_dim, tgt_rows, static_cast<uint64_t>(d.out_dim) * 4);
  // slot map (nine [128 x 64] blocks): 0.. encoder activation hi, hk.. lo | 4, 5 gathered embedding hi, lo | 6, 7 z hi, lo
  const int S_H = 0, S_HLO = sp ? 2 : -1, S_X = 4, S_XLO = sp ? 5 : -1, S_Z = 6, S_ZLO = sp ? 7 : -1;
  { RcOp& o = b.op(RC_LOADA);      // the gathered embedding rows (written by the ingest launch)
    o.tm_b = b.act(wsx.x, sp ? wsx.x_lo + E : E, wsx.ldx, B); o.kb = 1; o.a_slot = S_X; o.a_lo_slot = S_XLO; o.b_lo = wsx.x_lo; }
  { RcOp& o = b.op(RC_BNACT);
    o.tm_side = tm_pre; o.side_tiles = H / 32; o.out_slot = S_H; o.out_lo_slot = S_HLO; o.n_total = H;
    o.p[1] = wa.bits[0];
    o.st[0] = RcStore{static_cast<short>(b.act(wa.act[0], H, wa.ld_act[0], B)), static_cast<short>(S_H), static_cast<short>(hk), 0}; }
  { RcOp& o = b.op(RC_GEMM);       // heads of the dense encoder
    o.tm_b = b.w_nt(ea.heads); o.n = HW; o.kb = hk; o.a_slot = S_H; o.a_lo_slot = S_HLO; o.b_lo = ea.heads.sh_lo; o.tmem_col = 0; }
  { RcOp& o = b.op(RC_GEMM);       // heads of the site encoder
    o.tm_b = b.w_nt(es.heads); o.n = HW; o.kb = 1; o.a_slot = S_X; o.a_lo_slot = S_XLO; o.b_lo = es.heads.sh_lo; o.tmem_col = 256;
    o.commit = 1; o.wait_lda = 1; }
  { RcOp& o = b.op(RC_EPI, EP_LATENT);
    o.e_tmem = 0; o.e_tmem2 = 256; o.out_slot = S_Z; o.out_lo_slot = S_ZLO;
    o.p[0] = P + ea.heads.b_off; o.p[1] = P + es.heads.b_off; o.p[2] = m->mu; o.p[3] = m->logvar; o.p[4] = m->eps; o.p[5] = io.eps;
    o.st[0] = RcStore{static_cast<short>(b.act(m->z, L, m->ldz, B)), static_cast<short>(S_Z), 1, 0}; }
  const Lin& c = m->cat;
  const int ck = c.out / 64, k1 = l1.out / 64, k2 = ceil_div(l2.out, 64);
  { RcOp& o = b.op(RC_GEMM);       // fused first decoder layer
    o.tm_b = b.w_nt(c); o.n = c.out; o.kb = 1; o.a_slot = S_Z; o.a_lo_slot = S_ZLO; o.b_lo = c.sh_lo; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_RELU);
    o.e_n = c.out; o.n_total = c.out; o.relu = 1; o.out_slot = 0; o.out_lo_slot = sp ? ck : -1;
    o.p[0] = P + c.b_off; o.p[1] = m->d0_bits;
    o.st[0] = RcStore{static_cast<short>(b.act(m->d0, c.out, m->ld_d0, B)), 0, static_cast<short>(ck), 0}; }
  { RcOp& o = b.op(RC_GEMM);       // hidden decoder layer
    o.tm_b = b.w_nt(l1); o.n = l1.out; o.kb = ck; o.a_slot = 0; o.a_lo_slot = sp ? ck : -1; o.b_lo = l1.sh_lo; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_RELU);
    o.e_n = l1.out; o.n_total = l1.out; o.relu = 1; o.out_slot = 0;
    o.p[0] = P + l1.b_off; o.p[1] = wd.bits[0];
    o.st[0] = RcStore{static_cast<short>(b.act(wd.act[0], l1.out, wd.ld_act[0], B)), 0, static_cast<short>(k1), 0}; }
  // output layer + loss: at most 512 accumulator columns at a time; the columns beyond 512 go FIRST (their gradient block
  // lands in the slot behind the hidden activation, which the second round still reads)
  const int tm_w2 = b.w_nt(l2);
  const int tail_n = l2.out > 512 ? l2.out - 512 : 0;
  if (tail_n) {
    { RcOp& o = b.op(RC_GEMM); o.tm_b = tm_w2; o.n0 = 512; o.n = tail_n; o.kb = k1; o.a_slot = 0; o.commit = 1; }
    { RcOp& o = b.op(RC_EPI, EP_LOSS);
      o.e_n = tail_n; o.e_col0 = 512; o.n_total = l2.out; o.out_slot = 0; o.tm_side = tm_tgt; o.side_tiles = ceil_div(tail_n, 32);
      o.p[0] = P + l2.b_off; }
  }
  { RcOp& o = b.op(RC_GEMM); o.tm_b = tm_w2; o.n0 = 0; o.n = std::min(l2.out, 512); o.kb = k1; o.a_slot = 0; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_LOSS);
    o.e_n = std::min(l2.out, 512); o.e_col0 = 0; o.n_total = l2.out; o.out_slot = 0; o.tm_side = tm_tgt; o.side_tiles = ceil_div(std::min(l2.out, 512), 32);
    o.last_loss = 1; o.p[0] = P + l2.b_off;
    o.st[0] = RcStore{static_cast<short>(b.act(wd.g_out, l2.out, wd.ld_gout, B)), 0, static_cast<short>(k2), 0}; }
  // ---- backward ----
  { RcOp& o = b.op(RC_GEMM); o.tm_b = b.w_nn(l2); o.nn = 1; o.n = l2.in; o.kb = k2; o.a_slot = 0; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_MASK);
    o.e_n = l2.in; o.n_total = l2.in; o.out_slot = 0; o.p[1] = wd.bits[0];
    o.st[0] = RcStore{static_cast<short>(b.act(wd.gact[0], l2.in, l2.in, B)), 0, static_cast<short>(k1), 0}; }
  { RcOp& o = b.op(RC_GEMM); o.tm_b = b.w_nn(l1); o.nn = 1; o.n = l1.in; o.kb = k1; o.a_slot = 0; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_MASK);
    o.e_n = l1.in; o.n_total = l1.in; o.out_slot = 0; o.p[1] = m->d0_bits;
    o.st[0] = RcStore{static_cast<short>(b.act(m->g_d0, c.out, c.out, B)), 0, static_cast<short>(ck), 0}; }
  { RcOp& o = b.op(RC_GEMM); o.tm_b = b.w_nn(c); o.nn = 1; o.n = L; o.kb = ck; o.a_slot = 0; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_LATENT_BWD);
    o.e_tmem = 0; o.out_slot = 4; o.out_slot2 = 4 + ceil_div(HW, 64); o.p[2] = m->mu;
    // scratch behind the operand: dL/dz [128][L + 1] and, if they fit in the free slots, mu / logvar / eps [3][128 x L]
    o.e_n2 = (static_cast<size_t>(GEMM_BM) * (L + 1) + 3u * GEMM_BM * L) * 4 <= static_cast<size_t>(9 - o.out_slot2) * 16384 ? 1 : 0; o.p[3] = m->logvar; o.p[4] = m->eps;
    o.st[0] = RcStore{static_cast<short>(b.act(m->gml, HW, m->ldgml, B)), 4, static_cast<short>(ceil_div(HW, 64)), 0}; }
  const int gk = ceil_div(HW, 64);
  { RcOp& o = b.op(RC_GEMM); o.tm_b = b.w_nn(ea.heads); o.nn = 1; o.n = H; o.kb = gk; o.a_slot = 4; o.tmem_col = 0; }
  { RcOp& o = b.op(RC_GEMM); o.tm_b = b.w_nn(es.heads); o.nn = 1; o.n = E; o.kb = gk; o.a_slot = 4; o.tmem_col = 256; o.commit = 1; }
  { RcOp& o = b.op(RC_EPI, EP_DGRAD_ENC);
    o.e_tmem = 0; o.e_n = H; o.e_tmem2 = 256; o.e_n2 = E; o.out_slot = 0; o.out_slot2 = 8; o.tm_side = tm_pre; o.side_tiles = H / 32;
    o.n_total = H; o.fscale = io.train ? 1.0f / 0.9f : 1.0f;
    o.p[0] = wa.bstats[0]; o.p[1] = wa.bits[0];
    o.st[0] = RcStore{static_cast<short>(b.act(wa.gy[0], H, H, B)), 0, static_cast<short>(hk), 0};
    o.st[1] = RcStore{static_cast<short>(b.act(wsx.g_x, E, wsx.ld_gx, B)), 2, 1, 0}; }
  return b.rc;
}

// The launches of one train step, in order (row-local stretches are collected and issued as chain launches when m->chain_on).
static int train_step_sequence(vla_model_t* m, const vla_train_args_t* a, cudaStream_t st) {
  FwdIO io{};
  io.params = a->params; io.buffers = a->buffers; io.counters = a->counters;
  // encoder inputs per kind (train_rna2dna.py:86, train_dna2rna.py:86, optimize_hyperparameters.py:106)
  for (const Enc& e : m->encs) {
    if (e.slot == 0) io.x[0] = a->x_a;
    if (e.slot == 1) io.x[1] = a->x_b;
    if (e.slot == 2) io.site = a->site;
  }
  io.batch = a->batch; io.train = 1; io.eps = a->eps; io.keep_masks = a->keep_masks;
  io.seed = a->seed; io.offset = 0;
  io.recon[0] = a->recon_a; io.recon[1] = a->recon_b; io.recon[2] = a->recon_c;
  io.mu = a->mu; io.logvar = a->logvar; io.engine = true;
  if (a->batch <= 0) return fail(VLA_ERR_INVALID, "batch must be positive");
  const bool do_fb = a->phases != 2, do_opt = a->phases != 1;
  vla_dp* dp = reinterpret_cast<vla_dp*>(a->dp);
  if (a->sync_bn && !dp) return fail(VLA_ERR_INVALID, "vla_train_step: sync_bn needs the data-parallel exchange (dp)");
  if (dp) {
    dp->small_next = 0;
    if (a->sync_bn) {
      if (m->chain_on || rowchain_enabled() || headblock_enabled())
        return fail(VLA_ERR_STATE, "vla_train_step: sync_bn runs the separate launches only (VLA_CHAIN / VLA_ROWCHAIN / VLA_HEADBLOCK off)");
      io.sync_dp = dp;
    }
    if (!do_fb || !do_opt) return fail(VLA_ERR_INVALID, "vla_train_step: the peer-memory exchange needs the whole step (phases 0 or 3)");
    if (!dp->connected) return fail(VLA_ERR_STATE, "vla_train_step: vla_dp_connect has not been called");
    if (dp->n != m->n_params + 4) return fail(VLA_ERR_INVALID, "vla_train_step: exchange buffer size != param_count + 4");
    if (a->grads != reinterpret_cast<float*>(dp->local) || a->loss_out != a->grads + m->n_params)
      return fail(VLA_ERR_INVALID, "vla_train_step: grads / loss_out must be vla_dp_grads() and its last 4 floats");
  }
  if (!do_fb)
    return run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, 0.f, a->beta1, a->beta2, a->adam_eps, 0.f, 0, true,
                     true, st);
  io.n_batches = a->dataset_rows > a->batch ? static_cast<int>(a->dataset_rows / a->batch) : 1;
  io.beta1 = a->beta1; io.beta2 = a->beta2;
  // loss fused into the last decoder layers' epilogues (the CE term needs the whole logit row in one 32-column chunk)
  static const bool fuse_loss_on = [] { const char* e = getenv("VLA_FUSE_LOSS"); return !(e && e[0] == '0'); }();
  io.fuse_loss = m->S <= 32 && fuse_loss_on;
  for (const Dec& d : m->decs) {
    if (d.type == 'A' && !a->x_a) return fail(VLA_ERR_INVALID, "x_a (target) missing");
    if (d.type == 'B' && !a->x_b) return fail(VLA_ERR_INVALID, "x_b (target) missing");
    if (d.type == 'C' && !a->site) return fail(VLA_ERR_INVALID, "site (target) missing");
  }
  io.tgt_a = a->x_a; io.tgt_b = a->x_b; io.tgt_site = a->site; io.class_w = a->class_weights; io.loss_out = a->loss_out;
  {
    static const bool defer_on = [] { const char* e = getenv("VLA_DEFER_LOSS"); return !(e && e[0] == '0'); }();
    io.defer_loss_sum = defer_on && do_opt && dp == nullptr && !recorder();
  }
  int rc;
  // ---- row-chain step: ingest + first encoder layer | the on-chip middle | BatchNorm backward + weight gradients | AdamW ----
  const bool rowchain = !recorder() && rowchain_enabled() && io.fuse_loss && rowchain_fits(m, present_mask(m, io)) && !a->recon_a && !a->recon_b &&
                        !a->recon_c && !a->mu && !a->logvar && a->batch >= 2;
  if (rowchain) {
    io.rc_prefix = true;
    const bool chain_prev = m->chain_on;
    m->chain_on = false;                       // the prefix and the suffix are plain launches
    if ((rc = run_forward(m, io, st))) { m->chain_on = chain_prev; return rc; }
    RcPlan pl;
    if ((rc = build_rowchain_plan(m, io, &pl))) { m->chain_on = chain_prev; return rc; }
    {
      const char* tl = getenv("VLA_RC_TIMELINE");
      if (tl && tl[0] == '1') {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (!m->rc_dbg && cap == cudaStreamCaptureStatusNone) {
          if (cudaMalloc(&m->rc_dbg, sizeof(unsigned long long) * 148 * RC_MAX_OPS * 4) == cudaSuccess)
            cudaMemset(m->rc_dbg, 0, sizeof(unsigned long long) * 148 * RC_MAX_OPS * 4);
          else { m->rc_dbg = nullptr; (void)cudaGetLastError(); }
        }
        pl.dbg = m->rc_dbg;
        { const char* ex = getenv("VLA_RC_EXPERIMENT"); pl.pad2 = ex ? atoi(ex) : 0; }   // 1: no MMAs, 2: no weight loads (timing only)
        if (!m->rc_last) m->rc_last = new RcPlan();
        *m->rc_last = pl;
      }
    }
    {
      double fl = 0, by = 0;
      ProfScope ps(m, st, "rowchain_fwd_bwd", fl, by);
      cudaError_t e = launch_rowchain(pl, std::min(pl.m_blocks, 148), st);
      if (e != cudaSuccess) { m->chain_on = chain_prev; return fail(VLA_ERR_CUDA, std::string("row-chain launch: ") + cudaGetErrorString(e)); }
    }
    BwdIO bo{};
    bo.params = a->params; bo.grads = a->grads; bo.engine = true; bo.zero_grads = false; bo.dp = dp; bo.rc_suffix = true;
    rc = run_backward(m, bo, st);
    m->chain_on = chain_prev;
    if (rc) return rc;
  } else {
  if ((rc = run_forward(m, io, st))) return rc;
  // ---- loss: values + bf16 gradients for the backward GEMMs (only when it could not be fused) ----
  if (!io.fuse_loss) {
    LossArgs l{};
    l.rows = a->batch; l.dyn = m->dyn; l.grad_scale = 1.0f; l.dyn_bump = m->dyn; l.n_batches = io.n_batches;
    for (size_t i = 0; i < m->decs.size(); ++i) {
      const Dec& d = m->decs[i]; DecWS& w = m->dws[i];
      const int slot = d.type == 'A' ? 0 : (d.type == 'B' ? 1 : 2);
      const float* recon = io.recon[slot] ? io.recon[slot] : w.recon;
      if (d.type == 'A') {
        if (!a->x_a) return fail(VLA_ERR_INVALID, "x_a (target) missing");
        l.recon_a = recon; l.a = a->x_a; l.width_a = d.out_dim; l.ga_bf16 = w.g_out; l.ld_ga = w.ld_gout;
      } else if (d.type == 'B') {
        if (!a->x_b) return fail(VLA_ERR_INVALID, "x_b (target) missing");
        l.recon_b = recon; l.b = a->x_b; l.width_b = d.out_dim; l.gb_bf16 = w.g_out; l.ld_gb = w.ld_gout;
      } else {
        if (!a->site) return fail(VLA_ERR_INVALID, "site (target) missing");
        l.logits = recon; l.site = a->site; l.class_w = a->class_weights; l.n_sites = m->S;
        l.gc_bf16 = w.g_out; l.ld_gc = w.ld_gout;
      }
    }
    l.kl_partials = m->kl_partials; l.n_kl_partials = m->kl_grid; l.mu = m->mu; l.logvar = m->logvar; l.L = m->L;
    l.partials = m->loss_partials; l.counter = m->loss_counter; l.out = a->loss_out;
    { double by = 0;
      if (l.recon_a) by += static_cast<double>(l.rows) * l.width_a * 10.0;
      if (l.recon_b) by += static_cast<double>(l.rows) * l.width_b * 10.0;
      if (l.logits) by += static_cast<double>(l.rows) * l.n_sites * 6.0;
      if ((rc = chain_flush(m, st))) return rc;
      { ProfScope ps(m, st, "loss", 0, by); CK(launch_loss(l, st)); } }
  }
  BwdIO bo{};
  bo.params = a->params; bo.grads = a->grads; bo.engine = true; bo.zero_grads = false;   // AdamW leaves grads zeroed
  bo.dp = dp; bo.side_dec = m->dec_chunk0 > 0; bo.sync_bn = dp != nullptr && a->sync_bn != 0;
  if ((rc = run_backward(m, bo, st))) return rc;
  }
  if (m->side_busy) {
    // the side branch carries the decoder weight gradients; give it the decoder part of AdamW too, then join
    m->side_busy = false;
    static const bool side_adam = [] { const char* e = getenv("VLA_SIDE_ADAM"); return e && e[0] == '1'; }();
    if (!side_adam || dp) {      // join, then the optimizer launch of the normal flow (plain AdamW, or exchange + AdamW)
      CK(cudaEventRecord(m->ev_join, m->side));
      CK(cudaStreamWaitEvent(st, m->ev_join, 0));
      goto side_joined;
    }
    const int nd = static_cast<int>(m->chunks_h.size()) - m->dec_chunk0;
    if (do_opt && (rc = run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, 0.f, a->beta1, a->beta2, a->adam_eps, 0.f, 0,
                                  true, true, m->side, nullptr, nullptr, m->dec_chunk0, nd, "adamw_dec"))) return rc;
    CK(cudaEventRecord(m->ev_join, m->side));
    if (do_opt && (rc = run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, 0.f, a->beta1, a->beta2, a->adam_eps, 0.f, 0,
                                  true, true, st, nullptr, nullptr, 0, m->dec_chunk0, "adamw_enc"))) return rc;
    CK(cudaStreamWaitEvent(st, m->ev_join, 0));
    return VLA_OK;
  }
side_joined:
  if (!do_opt) return VLA_OK;
  if (dp) {
    // ---- the step's one collective, second part: the encoder gradients (the decoder part left from run_backward on the
    // side stream, or goes now if there was no decoder gradient); then AdamW, which polls the framed sums of both parts ----
    // (an engine step carries every decoder gradient, so run_backward took the early path when VLA_DP_OVERLAP=1)
    const char* fu = getenv("VLA_DP_FUSE");
    const bool overlap = dp_overlap(dp), fuse = !(fu && fu[0] == '0');
    const long long end2 = overlap ? m->cat.w_off / 2 : dp->n / 2;
    if (overlap) CK(cudaStreamWaitEvent(st, dp->ev_join, 0));
    if (fuse) {
      const DpArgs x = make_dp_args(m, dp, 0, end2, 0);
      return run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, 0.f, a->beta1, a->beta2, a->adam_eps, 0.f, 0, true,
                       false, st, dp, &x);
    }
    if ((rc = run_exchange(m, dp, 0, end2, 0, st, true))) return rc;
    return run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, 0.f, a->beta1, a->beta2, a->adam_eps, 0.f, 0, true,
                     false, st, dp);
  }
  return run_adamw(m, a->params, a->grads, a->exp_avg, a->exp_avg_sq, 0.f, a->beta1, a->beta2, a->adam_eps, 0.f, 0, true,
                   true, st);
}


int vla_train_step(vla_model_t* m, const vla_train_args_t* a, vla_stream_t stream) {
  if (!m || !a || !a->params || !a->grads || !a->exp_avg || !a->exp_avg_sq || !a->buffers || !a->loss_out)
    return fail(VLA_ERR_INVALID, "null argument");
  if (a->batch <= 0) return fail(VLA_ERR_INVALID, "batch must be positive");
  cudaStream_t st = as_stream(stream);
  ChainScope cs(m, chain_enabled() && a->phases != 2);
  int rc = train_step_sequence(m, a, st);
  if (rc) return rc;
  return cs.finish(st);
}

// ---------------------------------------------------------------------------------------------
// Lock-step population step
// ---------------------------------------------------------------------------------------------
int vla_train_step_group(vla_model_t* const* ms, const vla_train_args_t* const* as, int n, vla_stream_t stream) {
  if (!ms || !as || n <= 0) return fail(VLA_ERR_INVALID, "vla_train_step_group: null argument");
  if (n > MULTI_MAX_MEMBERS) return fail(VLA_ERR_INVALID, "vla_train_step_group: at most " + std::to_string(MULTI_MAX_MEMBERS) + " members per call");
  cudaStream_t st = as_stream(stream);
  // ---- record every member's launches (nothing is issued) ----
  std::vector<Recorder> recs(n);
  for (int i = 0; i < n; ++i) {
    vla_model_t* m = ms[i]; const vla_train_args_t* a = as[i];
    if (!m || !a || !a->params || !a->grads || !a->exp_avg || !a->exp_avg_sq || !a->buffers || !a->loss_out)
      return fail(VLA_ERR_INVALID, "vla_train_step_group: null argument (member " + std::to_string(i) + ")");
    if (a->dp) return fail(VLA_ERR_INVALID, "vla_train_step_group: members are independent models (no data-parallel exchange)");
    if (m->prof_on) return fail(VLA_ERR_STATE, "vla_train_step_group: per-launch profiling is per model");
    for (int j = 0; j < i; ++j) if (ms[j] == m) return fail(VLA_ERR_INVALID, "vla_train_step_group: a model appears twice");
    const bool chain_prev = m->chain_on;
    m->chain_on = false;
    recorder() = &recs[i];
    const int rc = train_step_sequence(m, a, st);
    recorder() = nullptr;
    m->chain_on = chain_prev;
    if (recs[i].unsupported) return fail(VLA_ERR_STATE, "vla_train_step_group: this step contains a launch without a grouped form (member " + std::to_string(i) + ")");
    if (rc) return rc;
  }
  // ---- zip: launch j of every member must be the same kernel ----
  const size_t n_ops = recs[0].ops.size();
  for (int i = 1; i < n; ++i) {
    if (recs[i].ops.size() != n_ops) return fail(VLA_ERR_INVALID, "vla_train_step_group: members of different kinds (launch counts differ)");
    for (size_t j = 0; j < n_ops; ++j)
      if (recs[i].ops[j].kind != recs[0].ops[j].kind || recs[i].ops[j].variant != recs[0].ops[j].variant ||
          recs[i].ops[j].args.size() != recs[0].ops[j].args.size())
        return fail(VLA_ERR_INVALID, "vla_train_step_group: members of different kinds (launch " + std::to_string(j) + " differs)");
  }
  // ---- host image ----
  std::string img;
  std::vector<GroupPlanCached::Launch> launches(n_ops);
  for (size_t j = 0; j < n_ops; ++j) {
    GroupPlanCached::Launch& l = launches[j];
    const RecOp& o0 = recs[0].ops[j];
    l.kind = o0.kind; l.variant = o0.variant; l.n = n; l.max_units = 0;
    if (l.kind == RK_GEMM)
      for (int i = 0; i < n; ++i)
        l.max_units = std::max(l.max_units, gemm_max_units(*reinterpret_cast<const GemmGroup*>(recs[i].ops[j].args.data())));
    l.stride = static_cast<int>((o0.args.size() + 63) / 64 * 64);          // tensor maps inside a GemmGroup need 64-byte alignment
    img.resize((img.size() + 255) / 256 * 256, '\0');
    l.hdr_off = img.size();
    int begin = 0;
    for (int i = 0; i < n; ++i) {
      const RecOp& o = recs[i].ops[j];
      MultiHdr h{begin, o.gx, o.aux, 0};
      img.append(reinterpret_cast<const char*>(&h), sizeof(h));
      begin += o.blocks;
    }
    l.total_blocks = begin;
    img.resize((img.size() + 255) / 256 * 256, '\0');
    l.args_off = img.size();
    for (int i = 0; i < n; ++i) {
      img.append(recs[i].ops[j].args);
      img.resize(l.args_off + static_cast<size_t>(i + 1) * l.stride, '\0');
    }
  }
  // ---- cached device copy (keyed by the image) ----
  vla_model_t* lead = ms[0];
  GroupPlanCached* plan = nullptr;
  for (GroupPlanCached* c : lead->group_plans) if (c->key == img) { plan = c; break; }
  if (!plan) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap != cudaStreamCaptureStatusNone)
      return fail(VLA_ERR_STATE, "vla_train_step_group: a new plan cannot be built inside a stream capture; make the same call once outside the capture");
    if (lead->group_plans.size() >= 64) {
      if (lead->pinned > 0) return fail(VLA_ERR_STATE, "vla_train_step_group: plan cache full while captured graphs are live");
      cudaStreamSynchronize(st);
      for (GroupPlanCached* g : lead->group_plans) { cudaFree(g->dev); delete g; }
      lead->group_plans.clear();
    }
    plan = new GroupPlanCached();
    if (cudaMalloc(&plan->dev, img.size()) != cudaSuccess) { delete plan; (void)cudaGetLastError(); return fail(VLA_ERR_CUDA, "vla_train_step_group: cudaMalloc"); }
    if (cudaMemcpy(plan->dev, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(plan->dev); delete plan; (void)cudaGetLastError();
      return fail(VLA_ERR_CUDA, "vla_train_step_group: copy of the launch tables");
    }
    plan->key = img; plan->launches = launches;
    lead->group_plans.push_back(plan);
  }
  // ---- issue ----
  // VLA_GROUP_PROF=1 (diagnostic, eager calls only): an event pair around every merged launch, printed to stderr
  static const bool gprof = [] { const char* e = getenv("VLA_GROUP_PROF"); return e && e[0] == '1'; }();
  std::vector<cudaEvent_t> gev;
  bool timing = false;
  if (gprof) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    timing = cap == cudaStreamCaptureStatusNone;
  }
  auto stamp = [&]() { if (timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); gev.push_back(e); } };
  stamp();
  for (const GroupPlanCached::Launch& l : plan->launches) {
    const MultiHdr* hdr = reinterpret_cast<const MultiHdr*>(plan->dev + l.hdr_off);
    const void* args = plan->dev + l.args_off;
    cudaError_t e;
    if (l.kind == RK_GEMM) {
      if (l.stride != static_cast<int>(sizeof(GemmGroup))) return fail(VLA_ERR_STATE, "vla_train_step_group: GemmGroup stride");
      e = launch_gemm_multi(l.variant, hdr, reinterpret_cast<const GemmGroup*>(args), l.n, l.total_blocks, l.max_units, st);
    } else {
      e = launch_multi(l.kind, l.variant, hdr, args, l.stride, l.n, l.total_blocks, st);
    }
    if (e != cudaSuccess) return fail(VLA_ERR_CUDA, std::string("vla_train_step_group: launch: ") + cudaGetErrorString(e));
    stamp();
  }
  if (timing) {
    static const char* kind_name[RK_COUNT] = {"gemm", "ingest", "bn_act", "bn_bwd", "latent_fwd", "latent_bwd", "adamw", "loss"};
    cudaStreamSynchronize(st);
    for (size_t j = 0; j + 1 < gev.size(); ++j) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, gev[j], gev[j + 1]);
      const GroupPlanCached::Launch& l = plan->launches[j];
      fprintf(stderr, "[group] %2zu %-10s variant %2d blocks %6d max_units %3d  %9.1f us\n", j, kind_name[l.kind], l.variant, l.total_blocks,
              l.max_units, 1e3 * ms);
    }
    for (cudaEvent_t e : gev) cudaEventDestroy(e);
  }
  return VLA_OK;
}
int vla_group_cached_plans(vla_model_t* lead) { return lead ? static_cast<int>(lead->group_plans.size()) : 0; }

int vla_model_pin(vla_model_t* m, int delta) {
  if (!m) return fail(VLA_ERR_INVALID, "null model");
  m->pinned = std::max(0, m->pinned + delta);
  return VLA_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer-memory gradient exchange: allocation, IPC export / import
// ---------------------------------------------------------------------------------------------
int vla_dp_create(int world, int rank, long long n_floats, vla_dp_t** out) {
  if (!out || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world || n_floats < 4 || (n_floats & 3))
    return fail(VLA_ERR_INVALID, "vla_dp_create: world in [1, 16], rank in [0, world), n_floats a positive multiple of 4");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "vla_dp_ipc_handle writes 64 bytes");
  vla_dp* d = new vla_dp();
  d->world = world; d->rank = rank; d->n = n_floats;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  const long long n2 = n_floats / 2;
  d->per2 = ((n2 + world - 1) / world + 1) & ~1LL;
  d->off_recv = 0;
  d->off_rsum = up(sizeof(uint4) * 2 * static_cast<size_t>(world) * d->per2);   // one RECV region per part
  d->off_small = d->off_rsum + up(sizeof(uint4) * static_cast<size_t>(n2));       // SyncBN slot arrays [region][world][words]
  d->bytes = d->off_small + up(sizeof(uint4) * static_cast<size_t>(DP_SMALL_REGIONS) * world * DP_SMALL_WORDS);
  d->off_sums = up(sizeof(float) * n_floats);
  d->off_trace = d->off_sums + 256;
  d->local_bytes = d->off_trace + 256;
  // G stays in a private allocation: only this rank ever touches its gradients, so nothing of it needs exporting.
  cudaError_t e = cudaMalloc(&d->base, d->bytes);
  if (e == cudaSuccess) e = cudaMemset(d->base, 0, d->bytes);
  if (e == cudaSuccess) e = cudaMalloc(&d->local, d->local_bytes);
  if (e == cudaSuccess) e = cudaMemset(d->local, 0, d->local_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) { cudaFree(d->base); cudaFree(d->local); delete d; return fail(VLA_ERR_CUDA, std::string("vla_dp_create: ") + cudaGetErrorString(e)); }
  d->peer[rank] = d->base;
  d->connected = world == 1;
  *out = d;
  return VLA_OK;
}
int vla_dp_ipc_handle(vla_dp_t* d, void* out64) {
  if (!d || !out64) return fail(VLA_ERR_INVALID, "null argument");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, d->base));
  memcpy(out64, &h, sizeof(h));
  return VLA_OK;
}
int vla_dp_connect(vla_dp_t* d, const void* handles) {
  if (!d || !handles) return fail(VLA_ERR_INVALID, "null argument");
  if (d->connected) return VLA_OK;
  for (int r = 0; r < d->world; ++r) {
    if (r == d->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + 64 * r, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(VLA_ERR_CUDA, "vla_dp_connect: cudaIpcOpenMemHandle(rank " + std::to_string(r) + "): " + cudaGetErrorString(e) +
                                    " (the ranks must be processes on one node whose GPUs have peer access)");
    }
    d->peer[r] = static_cast<char*>(p);
  }
  d->connected = true;
  return VLA_OK;
}
int vla_dp_trace(vla_dp_t* d, unsigned long long* out8) {
  if (!d || !out8) return fail(VLA_ERR_INVALID, "null argument");
  CK(cudaMemcpy(out8, d->local + d->off_trace, 64, cudaMemcpyDeviceToHost));
  return VLA_OK;
}
void* vla_dp_grads(vla_dp_t* d) { return d ? d->local : nullptr; }
void* vla_dp_losses(vla_dp_t* d) { return d ? d->local + d->off_sums : nullptr; }
/* First half of the teardown: unmap every peer's exported buffer.  The caller then runs a barrier over the ranks and only after
 * that frees its own exported allocation (vla_dp_destroy): CUDA leaves freeing an exported allocation that an importer still
 * has mapped undefined. */
int vla_dp_disconnect(vla_dp_t* d) {
  if (!d) return fail(VLA_ERR_INVALID, "null argument");
  CK(cudaDeviceSynchronize());
  for (int r = 0; r < d->world; ++r)
    if (r != d->rank && d->peer[r]) { cudaIpcCloseMemHandle(d->peer[r]); d->peer[r] = nullptr; }
  d->connected = d->world == 1;
  return VLA_OK;
}
void vla_dp_destroy(vla_dp_t* d) {
  if (!d) return;
  for (int r = 0; r < d->world; ++r)
    if (r != d->rank && d->peer[r]) cudaIpcCloseMemHandle(d->peer[r]);
  if (d->side) cudaStreamDestroy(d->side);
  if (d->ev_fork) cudaEventDestroy(d->ev_fork);
  if (d->ev_join) cudaEventDestroy(d->ev_join);
  cudaFree(d->base);
  cudaFree(d->local);
  delete d;
}

// ---------------------------------------------------------------------------------------------
// Reconstruction metrics
// ---------------------------------------------------------------------------------------------
long long vla_metrics_workspace_bytes(long long rows) {
  return 256 + static_cast<long long>(metrics_grid(rows)) * 8 * static_cast<long long>(sizeof(double));
}
int vla_recon_metrics(const vla_metrics_args_t* a, vla_stream_t stream) {
  if (!a || !a->y_true || !a->y_pred || !a->out || !a->workspace) return fail(VLA_ERR_INVALID, "null argument");
  if (a->rows <= 0 || a->dim <= 0) return fail(VLA_ERR_INVALID, "rows and dim must be positive");
  MetricsArgs m{};
  m.yt = a->y_true; m.yp = a->y_pred; m.rows = a->rows; m.dim = a->dim;
  m.cos_out = a->cosine; m.pearson_out = a->pearson;
  m.counter = reinterpret_cast<unsigned int*>(a->workspace);
  m.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(a->workspace) + 256);
  m.out = a->out;
  CK(launch_metrics(m, as_stream(stream)));
  return VLA_OK;
}

int vla_gather_rows(const float* a, int dim_a, const float* b, int dim_b, const long long* site, long long rows,
                    const long long* index, int n, float* out_a, float* out_b, long long* out_site, vla_stream_t stream) {
  if (!a || !b || !site || !index || !out_a || !out_b || !out_site) return fail(VLA_ERR_INVALID, "null argument");
  if (rows <= 0 || n < 0 || dim_a <= 0 || dim_b <= 0) return fail(VLA_ERR_INVALID, "vla_gather_rows: bad extents");
  GatherArgs g{a, b, site, rows, dim_a, dim_b, index, n, out_a, out_b, out_site};
  CK(launch_gather_rows(g, as_stream(stream)));
  return VLA_OK;
}
int vla_scale_inplace(void* const* tensors, const long long* counts, int n_tensors, const float* scale, vla_stream_t stream) {
  if (!tensors || !counts || !scale || n_tensors < 0 || n_tensors > 8) return fail(VLA_ERR_INVALID, "vla_scale_inplace: 0..8 tensors");
  ScaleArgs a{};
  a.count = n_tensors; a.scale = scale;
  for (int i = 0; i < n_tensors; ++i) { a.x[i] = static_cast<float*>(tensors[i]); a.n[i] = counts[i]; }
  CK(launch_scale(a, as_stream(stream)));
  return VLA_OK;
}

/* Chain kernel timeline: per (CTA, phase) %globaltimer stamps (phase start, phase end before the cluster barrier) written by
 * the chain launches of the following calls. */
int vla_chain_timeline(vla_model_t* m, int enable) {
  if (!m) return fail(VLA_ERR_INVALID, "null model");
  CK(cudaDeviceSynchronize());
  m->chain_dbg = enable != 0;
  for (ChainPlanCached* c : m->plans) {
    unsigned long long* dbg = enable ? reinterpret_cast<unsigned long long*>(c->dev + c->dbg_off) : nullptr;
    CK(cudaMemcpy(c->dev + offsetof(ChainPlan, dbg), &dbg, sizeof(dbg), cudaMemcpyHostToDevice));
  }
  return VLA_OK;
}
int vla_chain_count(vla_model_t* m) { return m ? m->n_last_plans : 0; }
int vla_chain_cached_plans(vla_model_t* m) { return m ? static_cast<int>(m->plans.size()) : 0; }
int vla_chain_info(vla_model_t* m, int which, char* name48, int* n_phases, int* n_ctas, double* flops, double* bytes) {
  if (!m || which < 0 || which >= m->n_last_plans) return fail(VLA_ERR_INVALID, "bad chain index");
  const ChainPlanCached* c = m->last_plans[which];
  if (name48) snprintf(name48, 48, "%s", c->name.c_str());
  if (n_phases) *n_phases = c->n_phases;
  if (n_ctas) *n_ctas = c->n_clusters * CHAIN_CLUSTER;
  if (flops) *flops = c->flops;
  if (bytes) *bytes = c->bytes;
  return VLA_OK;
}
int vla_chain_phase_name(vla_model_t* m, int which, int phase, char* name48) {
  if (!m || which < 0 || which >= m->n_last_plans || !name48) return fail(VLA_ERR_INVALID, "bad chain index");
  const ChainPlanCached* c = m->last_plans[which];
  if (phase < 0 || phase >= c->n_phases) return fail(VLA_ERR_INVALID, "bad phase index");
  snprintf(name48, 48, "%s", c->phase_names[phase].c_str());
  return VLA_OK;
}
/* out[n_ctas][CHAIN_MAX_PHASES = 24][8] */
int vla_chain_timeline_read(vla_model_t* m, int which, unsigned long long* out) {
  if (!m || !out || which < 0 || which >= m->n_last_plans) return fail(VLA_ERR_INVALID, "bad chain index");
  const ChainPlanCached* c = m->last_plans[which];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, c->dev + c->dbg_off, sizeof(unsigned long long) * 8 * CHAIN_MAX_PHASES * c->n_clusters * CHAIN_CLUSTER, cudaMemcpyDeviceToHost));
  return c->n_clusters * CHAIN_CLUSTER;
}

/* Row-chain timeline (VLA_RC_TIMELINE=1): out[148][32][4] %globaltimer stamps of the last row-chain launch -- per op: GEMM
 * issue start, all MMAs issued, accumulator ready (element-wise ops: start), epilogue done; kinds[32] / subs[32] = the ops. */
int vla_rowchain_timeline(vla_model_t* m, unsigned long long* out, int* kinds, int* subs) {
  if (!m || !out || !m->rc_dbg || !m->rc_last) return fail(VLA_ERR_STATE, "no row-chain timeline (set VLA_RC_TIMELINE=1 before the step)");
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, m->rc_dbg, sizeof(unsigned long long) * 148 * RC_MAX_OPS * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < RC_MAX_OPS; ++i) {
    if (kinds) kinds[i] = i < m->rc_last->n_ops ? m->rc_last->ops[i].kind : 0;
    if (subs) subs[i] = i < m->rc_last->n_ops ? m->rc_last->ops[i].sub : 0;
  }
  return m->rc_last->n_ops;
}

int vla_profile_begin(vla_model_t* m) {
  if (!m) return fail(VLA_ERR_INVALID, "null model");
  for (ProfEntry& e : m->prof) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
  m->prof.clear();
  m->prof_on = true;
  return VLA_OK;
}
int vla_profile_pause(vla_model_t* m) {
  if (!m) return fail(VLA_ERR_INVALID, "null model");
  m->prof_on = false;
  return VLA_OK;
}
int vla_profile_read(vla_model_t* m, vla_prof_entry_t* out, int max_entries) {
  if (!m || !out) return fail(VLA_ERR_INVALID, "null argument");
  int n = 0;
  for (ProfEntry& e : m->prof) {
    if (n >= max_entries) break;
    float ms = 0.f;
    if (cudaEventSynchronize(e.e1) != cudaSuccess || cudaEventElapsedTime(&ms, e.e0, e.e1) != cudaSuccess) { (void)cudaGetLastError(); continue; }
    memcpy(out[n].name, e.name, sizeof(out[n].name));
    out[n].ms = ms; out[n].flops = e.flops; out[n].bytes = e.bytes;
    ++n;
  }
  return n;
}
int vla_profile_collect(vla_model_t* m, vla_prof_entry_t* out, int max_entries) {
  if (!m || !out) return fail(VLA_ERR_INVALID, "null argument");
  m->prof_on = false;
  int n = 0;
  for (ProfEntry& e : m->prof) {
    if (cudaEventSynchronize(e.e1) == cudaSuccess && n < max_entries) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e.e0, e.e1);
      memcpy(out[n].name, e.name, sizeof(out[n].name));
      out[n].ms = ms; out[n].flops = e.flops; out[n].bytes = e.bytes;
      ++n;
    }
    cudaEventDestroy(e.e0); cudaEventDestroy(e.e1);
  }
  m->prof.clear();
  return n;
}

/* Test hook: device address of a workspace buffer of the last forward on this handle (tests/test_gpu_philox.py reads the
 * epsilon the kernels drew and the post-dropout activations).  what 0: eps fp32 [rows, latent] (ld = latent);
 * what 1: activation bf16 of encoder `i`, BatchNorm layer `j` (ld = row pitch in elements).  Returns VLA_ERR_INVALID if absent. */
int vla_test_workspace(vla_model_t* m, int what, int i, int j, void** ptr, int* ld) {
  if (!m || !ptr || !ld || !m->ws) return fail(VLA_ERR_INVALID, "vla_test_workspace: no workspace");
  if (what == 0) { *ptr = m->eps; *ld = m->L; return VLA_OK; }
  if (what == 1 && i >= 0 && i < static_cast<int>(m->ews.size()) && j >= 0 && j < static_cast<int>(m->ews[i].act.size())) {
    *ptr = m->ews[i].act[j]; *ld = m->ews[i].ld_act[j]; return VLA_OK;
  }
  return fail(VLA_ERR_INVALID, "vla_test_workspace: no such buffer");
}

static unsigned long long* g_test_dbg = nullptr;
static int g_test_flags = 0;
int vla_test_set_flags(int flags) { g_test_flags = flags; return VLA_OK; }
/* Test hook: device buffer [tiles][8] that the next vla_test_gemm calls fill with %globaltimer stamps (NULL = off). */
int vla_test_set_timeline(unsigned long long* dbg) { g_test_dbg = dbg; return VLA_OK; }

int vla_test_gemm(int mode, const void* A, int lda, const void* B, int ldb, float* C, int M, int N, int K, int bn,
                  int k_splits, float* bias_grad, vla_stream_t stream) {
  if (bn > (mode == 0 ? GEMM_BN_MAX_NT : GEMM_BN_MAX_TN)) return fail(VLA_ERR_INVALID, "vla_test_gemm: tile wider than the kernel's limit (144 NT, 128 TN / NN)");
  static vla_model scratch;   // only its tensor-map cache is used
  GemmGroup g; init_group(g);
  int rc;
  if (mode == 0) {
    GemmProblem* p;
    if ((rc = add_nt(&scratch, g, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb, M, N, K, GF_OUT_F32, &p, bn))) return rc;
    p->out_f32 = C; p->ld_f32 = N;
  } else if (mode == 2) {
    GemmProblem* p;
    if ((rc = add_nn(&scratch, g, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb, M, N, K, GF_OUT_F32, &p, bn))) return rc;
    p->out_f32 = C; p->ld_f32 = N;
  } else {
    if ((rc = add_tn(&scratch, g, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb, M, N, K, C, N, bias_grad, bn))) return rc;
    finalize_tn(g, K, k_splits);
  }
  g.dbg = g_test_dbg;
  g.dbg_flags = g_test_flags;
  if (mode != 1 && (rc = finalize_group(&scratch, g, mode))) return rc;
  CK(launch_gemm_group(g, mode, as_stream(stream)));
  scratch.tmaps.clear();
  return VLA_OK;
}

}  // extern "C"
