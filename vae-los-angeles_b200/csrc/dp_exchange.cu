// Data-parallel gradient exchange over NVLink / NVSwitch peer memory: the one collective of the train step
// (SURVEY.md section 8e), written as a kernel over peer pointers instead of an NCCL call.
//
// Protocol: a two-shot all-reduce(SUM) in which every 8-byte word on the wire carries its own arrival flag
// ("low-latency" framing: {payload fp32, epoch} pairs, 16-byte stores of two such pairs).  An 8-byte store is delivered
// atomically, so a reader that sees the current epoch next to a payload knows the payload is the current one: there is
// no barrier, no memory fence and no second flag round trip anywhere in the exchange (a first version with "ready" /
// "written" flags and system-scope fences spent 6.5 us in each fence and 22 us per exchange, profiles/r1_dp_exchange.md).
//
// Every rank (one process per GPU, one node) owns one cudaMalloc allocation, exported with cudaIpcGetMemHandle and mapped
// by every other rank with cudaIpcOpenMemHandle:
//   G    fp32 [n]            local gradients: the backward kernels accumulate into it (flat gradient arena | 4 loss scalars)
//   RECV framed [W][per]     slot s: rank s's contribution to the shard this rank reduces
//   RSUM framed [n / 2]      the gradients summed over the ranks, as they arrive from the shard owners
// A launch exchanges one contiguous range of the flat buffer; rank r owns float2s [r * per, (r + 1) * per) of the range.
// The train step (vla_api.cu) runs the exchange INSIDE the optimizer launch (dp_adamw_kernel: phases A, B, then AdamW as
// phase C), so the collective adds no dependent launch to the step.  From 4 ranks on, the decoder gradients (80 % of the
// bytes, complete as soon as the decoder weight-gradient GEMMs are) leave earlier, through dp_exchange_kernel on a side
// stream WHILE the encoder backward runs, and the optimizer launch only exchanges the encoder gradients.
//   A. push: this rank's G values of every OTHER rank's shard go, framed, into that rank's RECV[r]; G is cleared as it is read;
//   B. reduce: for its own shard a rank adds, IN RANK ORDER (every rank computes bit-identical sums: the replicas cannot
//      drift), its own G values and the framed values polled from RECV[s], and pushes the framed sum into EVERY rank's RSUM;
//   C. consume: AdamW polls RSUM word by word as it walks the parameters (elementwise_dev.cuh, AdamArgs::gframed), so the
//      gather of the sums costs no pass of its own; the first block also copies the 4 summed loss scalars out.
// The epoch is DynParams::dp_epoch (bumped by the first kernel of every train step, never reset).  Buffer reuse is safe
// without further synchronisation: a rank starts pushing epoch e + 1 only after its own AdamW of epoch e, which needed
// every owner's sums of epoch e, which needed every rank's pushes of epoch e -- so all reads of epoch e data are over.
// Progress: block b of a rank waits (phase B) only for the pushes of block b of the other ranks, and phase A never waits,
// so blocks scheduled in index order cannot deadlock even when the launch shares the SMs with the encoder backward.
// The AdamW blocks wait only on remote progress.
// Wire traffic per rank and step: 2 x (W - 1) / W x 8n bytes out (and in) -- 7.5 MB at 8 GPUs for the rna2dna model.
#include "elementwise_dev.cuh"   // adamw_body, dp_frame.cuh
#include "tc_ptx.cuh"
#include "vla_internal.h"

namespace vla {

namespace {

constexpr int DP_THREADS = 256;

// Phases A and B for the calling block (every block of the launch calls this with the same `a`).
__device__ __forceinline__ void dp_exchange_body(const DpArgs& a) {
  const int tid = threadIdx.x;
  const unsigned int epoch = static_cast<unsigned int>(__ldcg(&a.dyn->dp_epoch));
  const bool tr = a.trace && blockIdx.x == 0 && tid == 0;
  if (tr) a.trace[0] = dp_now_ns();                                   // kernel entry
  const int W = a.world, me = a.rank;
  const long long per = a.per2;                                    // float2s per shard
  const long long gtid = static_cast<long long>(blockIdx.x) * DP_THREADS + tid;
  const long long stride = static_cast<long long>(gridDim.x) * DP_THREADS;
  float2* G2 = reinterpret_cast<float2*>(a.g) + a.first2;
  // A. push my contribution to every other rank's shard (framed), clearing G behind the read.  The loads of a batch of
  //    peers are all issued before the first (volatile) store, otherwise every peer costs a dependent L2 round trip.
  for (long long j = gtid; j < per; j += stride) {
    for (int d0 = 1; d0 < W; d0 += 8) {
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int s = (me + d0 + k) % W;                            // start at a different peer on every rank
        if (d0 + k < W && per * s + j < a.n2) v[k] = G2[per * s + j];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int s = (me + d0 + k) % W;
        if (d0 + k < W && per * s + j < a.n2) {
          G2[per * s + j] = make_float2(0.f, 0.f);
          st_framed(a.recv[s] + per * me + j, v[k].x, v[k].y, epoch);
        }
      }
    }
  }
  if (tr) a.trace[1] = dp_now_ns();                                   // block 0: pushes issued
  // B. reduce my shard in rank order, push the framed sums to every rank
  {
    const long long lo = per * me;
    const long long cnt = min(per, a.n2 - lo);
    const uint4* rcv = a.recv[me];
    for (long long j = gtid; j < cnt; j += stride) {
      const float2 mine = G2[lo + j];
      G2[lo + j] = make_float2(0.f, 0.f);
      float2 acc = make_float2(0.f, 0.f);
      for (int s0 = 0; s0 < W; s0 += 8) {
        uint4 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)                    // all polls of the batch in flight before the first check
          if (s0 + k < W && s0 + k != me) t[k] = ld_framed(rcv + per * (s0 + k) + j);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (s0 + k >= W) continue;
          float2 v = mine;
          if (s0 + k != me) {
            v = finish_framed(rcv + per * (s0 + k) + j, t[k], epoch);
          }
          acc.x += v.x; acc.y += v.y;
        }
      }
      for (int d = 0; d < W; ++d) st_framed(a.rsum[(me + d) % W] + a.first2 + lo + j, acc.x, acc.y, epoch);
    }
  }
  if (tr) a.trace[2] = dp_now_ns();                                   // block 0: shard slice reduced and pushed
}


__global__ void __launch_bounds__(DP_THREADS) dp_exchange_kernel(const DpArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  dp_exchange_body(a);
}

// The exchange and the optimizer step as ONE launch: a block pushes and reduces its slice (phases A, B) and then walks
// its AdamW chunks, polling the framed sums as they arrive from the shard owners (phase C).  One dependent launch less
// on the critical path of the step than exchange + AdamW.  Unlike the plain exchange, phase C of a block waits for
// phase B of arbitrary blocks of the other ranks, so every block of the launch must be resident: the grid is capped by
// the occupancy of this kernel (launch_dp_adamw) and the chunks are walked grid-stride.
__global__ void __launch_bounds__(DP_THREADS, 4) dp_adamw_kernel(const DpArgs x, const AdamArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  dp_exchange_body(x);
  for (int c = blockIdx.x; c < a.n_chunks; c += gridDim.x) adamw_body(a, c, threadIdx.x);
  if (a.sums_out && blockIdx.x == 0 && threadIdx.x < 2) {
    const unsigned int epoch = static_cast<unsigned int>(__ldcg(&a.dyn->dp_epoch));
    const uint4* w = a.gframed + a.tail2 + threadIdx.x;
    const float2 v = finish_framed(w, ld_framed(w), epoch);
    a.sums_out[2 * threadIdx.x] = v.x; a.sums_out[2 * threadIdx.x + 1] = v.y;
  }
}

// SyncBN: all-reduce(SUM) of one BatchNorm layer's column sums (vla_internal.h, DpSmallArgs).  Replaces what
// torch.nn.SyncBatchNorm would do for encoders.py:14,32,36 under data parallelism.
__global__ void __launch_bounds__(DP_THREADS) dp_small_allreduce_kernel(const DpSmallArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const unsigned int epoch = static_cast<unsigned int>(__ldcg(&a.dyn->dp_epoch));
  const int W = a.world, me = a.rank, n = a.n;
  // word i carries the flattened [2][n] values 2 i and 2 i + 1 (n is even)
  for (int i = threadIdx.x; i < n; i += DP_THREADS) {
    double s0 = 0, s1 = 0;
    for (int t = 0; t < a.m_tiles; ++t) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(a.partials + static_cast<size_t>(t) * 2 * n) + i);
      s0 += v.x; s1 += v.y;
    }
    const float f0 = static_cast<float>(s0), f1 = static_cast<float>(s1);
    for (int d = 0; d < W; ++d) st_framed(a.slots[(me + d) % W] + static_cast<size_t>(me) * DP_SMALL_WORDS + i, f0, f1, epoch);
    double t0 = 0, t1 = 0;
    for (int r = 0; r < W; ++r) {                         // rank order: every rank computes bit-identical sums
      const uint4* w = a.slots[me] + static_cast<size_t>(r) * DP_SMALL_WORDS + i;
      const float2 v = r == me ? make_float2(f0, f1) : finish_framed(w, ld_framed(w), epoch);
      t0 += v.x; t1 += v.y;
    }
    reinterpret_cast<float2*>(a.partials)[i] = make_float2(static_cast<float>(t0), static_cast<float>(t1));
  }
}

}  // namespace

cudaError_t launch_dp_small_allreduce(const DpSmallArgs& a, cudaStream_t s) {
  if (a.n <= 0 || (a.n & 1) || a.n > DP_SMALL_WORDS) return cudaErrorInvalidValue;
  return launch_pdl(dp_small_allreduce_kernel, dim3(1), dim3(DP_THREADS), 0, s, a);
}

cudaError_t launch_dp_exchange(const DpArgs& a, cudaStream_t s, bool pdl) {
  if (a.n2 <= 0) return cudaSuccess;
  long long blocks = (a.n2 + DP_THREADS - 1) / DP_THREADS;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;        // every rank uses the same grid: the block <-> element mapping must agree
  if (pdl) return launch_pdl(dp_exchange_kernel, dim3(static_cast<unsigned>(blocks)), dim3(DP_THREADS), 0, s, a);
  dp_exchange_kernel<<<static_cast<unsigned>(blocks), DP_THREADS, 0, s>>>(a);     // side stream: plain launch
  return cudaGetLastError();
}

// Exchange of float2s [x.first2, x.first2 + x.n2) fused with the AdamW step over the whole arena (a.gframed set).
// Phase C of a block waits for phase B of arbitrary blocks on the OTHER ranks, so every block of this launch must be
// resident at once.  The launch is therefore cooperative: if the blocks cannot all be co-resident (another stream, trainer or
// process occupies the SMs) it FAILS with an error instead of dead-locking across ranks.  Programmatic dependent launch is
// requested beside it where the driver accepts the pair; otherwise the launch is cooperative only.
cudaError_t launch_dp_adamw(const DpArgs& x, const AdamArgs& a, cudaStream_t s) {
  static int per_sm_of[64];                     // occupancy per device (0 = not queried yet)
  static int pdl_ok = -1;                       // cooperative + programmatic serialization accepted together?
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  int per_sm = dev >= 0 && dev < 64 ? per_sm_of[dev] : 0;
  if (per_sm == 0) {
    int v = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, dp_adamw_kernel, DP_THREADS, 0);
    if (e != cudaSuccess) return e;
    per_sm = v < 1 ? 1 : (v > 4 ? 4 : v);
    if (dev >= 0 && dev < 64) per_sm_of[dev] = per_sm;
  }
  long long blocks = a.n_chunks;
  if (blocks > static_cast<long long>(per_sm) * sms) blocks = static_cast<long long>(per_sm) * sms;   // all resident
  if (blocks < 1) blocks = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(blocks)); cfg.blockDim = dim3(DP_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  if (pdl_ok != 0) {
    cfg.numAttrs = 2;
    e = cudaLaunchKernelEx(&cfg, dp_adamw_kernel, x, a);
    if (e == cudaSuccess) { pdl_ok = 1; return e; }
    if (pdl_ok == 1) return e;                  // the pair works on this driver: a real failure (e.g. not co-resident)
    (void)cudaGetLastError();
    pdl_ok = 0;
  }
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, dp_adamw_kernel, x, a);
}

}  // namespace vla
