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
// One launch per step, between the weight-gradient GEMMs and AdamW; rank r owns float2s [r * per, (r + 1) * per):
//   A. push: this rank's G values of every OTHER rank's shard go, framed, into that rank's RECV[r]; G is cleared as it is read;
//   B. reduce: for its own shard a rank adds, IN RANK ORDER (every rank computes bit-identical sums: the replicas cannot
//      drift), its own G values and the framed values polled from RECV[s], and pushes the framed sum into EVERY rank's RSUM;
//   C. consume: the AdamW kernel launched behind this one polls RSUM word by word as it walks the parameters
//      (elementwise_dev.cuh, AdamArgs::gframed), so the gather of the sums costs no pass of its own; its first block also
//      copies the 4 summed loss scalars out.
// The epoch is DynParams::dp_epoch (bumped by the first kernel of every train step, never reset).  Buffer reuse is safe
// without further synchronisation: a rank starts pushing epoch e + 1 only after its own AdamW of epoch e, which needed
// every owner's sums of epoch e, which needed every rank's pushes of epoch e -- so all reads of epoch e data are over.
// All blocks of the launch must be co-resident (phase B of one rank waits for phase A of every block of its peers):
// the grid is at most four blocks per SM.  The AdamW blocks wait only on remote progress and need no such guarantee.
// Wire traffic per rank and step: 2 x (W - 1) / W x 8n bytes out (and in) -- 7.5 MB at 8 GPUs for the rna2dna model.
#include "dp_frame.cuh"
#include "tc_ptx.cuh"
#include "vla_internal.h"

namespace vla {

namespace {

constexpr int DP_THREADS = 256;

__global__ void __launch_bounds__(DP_THREADS) dp_exchange_kernel(const DpArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const unsigned int epoch = static_cast<unsigned int>(__ldcg(&a.dyn->dp_epoch));
  const bool tr = a.trace && blockIdx.x == 0 && tid == 0;
  if (tr) a.trace[0] = dp_now_ns();                                   // kernel entry
  const int W = a.world, me = a.rank;
  const long long per = a.per2;                                    // float2s per shard
  const long long gtid = static_cast<long long>(blockIdx.x) * DP_THREADS + tid;
  const long long stride = static_cast<long long>(gridDim.x) * DP_THREADS;
  float2* G2 = reinterpret_cast<float2*>(a.g);
  // A. push my contribution to every other rank's shard (framed), clearing G behind the read
  for (int d = 1; d < W; ++d) {
    const int s = (me + d) % W;                                     // start at a different peer on every rank
    const long long lo = per * s;
    const long long cnt = min(per, a.n2 - lo);
    uint4* dst = a.recv[s] + per * me;
    for (long long j = gtid; j < cnt; j += stride) {
      const float2 v = G2[lo + j];
      G2[lo + j] = make_float2(0.f, 0.f);
      st_framed(dst + j, v.x, v.y, epoch);
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
      for (int d = 0; d < W; ++d) st_framed(a.rsum[(me + d) % W] + lo + j, acc.x, acc.y, epoch);
    }
  }
  if (tr) a.trace[2] = dp_now_ns();                                   // block 0: shard slice reduced and pushed
}

}  // namespace

cudaError_t launch_dp_exchange(const DpArgs& a, cudaStream_t s) {
  long long blocks = (a.n2 + DP_THREADS - 1) / DP_THREADS;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;        // co-residency (see above)
  return launch_pdl(dp_exchange_kernel, dim3(static_cast<unsigned>(blocks)), dim3(DP_THREADS), 0, s, a);
}

}  // namespace vla
