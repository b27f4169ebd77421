// Framed words of the peer-memory gradient exchange (dp_exchange.cu): {payload fp32, epoch} pairs, two per 16-byte
// store.  An 8-byte store is delivered atomically, so a reader that finds the current epoch next to a payload knows the
// payload is current -- no fence, no separate flag.  Shared by the exchange kernel (producer / reducer) and the AdamW
// kernel (consumer of the reduced gradients).
#pragma once
#include <cuda_runtime.h>

namespace vla {

__device__ __forceinline__ void st_framed(uint4* p, float x, float y, unsigned int epoch) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(x)), "r"(epoch),
               "r"(__float_as_uint(y)), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ uint4 ld_framed(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool framed_ok(const uint4& v, unsigned int epoch) { return v.y == epoch && v.w == epoch; }
__device__ __forceinline__ unsigned long long dp_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Slow path of a poll, kept out of line: the callers are instruction-cache sensitive (an inlined spin loop per polled word
// made the exchange + AdamW kernel 40 KB of SASS and 10 us slower).  A peer that never arrives (crashed rank) traps after 60 s
// instead of hanging the GPU.
static __device__ __noinline__ uint4 spin_framed(const uint4* p, unsigned int epoch) {
  const unsigned long long t0 = dp_now_ns();
  int spins = 0;
  uint4 v;
  do {
    if ((++spins & 1023) == 0 && dp_now_ns() - t0 > 60ull * 1000000000ull) __trap();
    __nanosleep(64);                      // back off: thousands of threads polling back to back queue in front of real loads
    v = ld_framed(p);
  } while (!framed_ok(v, epoch));
  return v;
}
// Complete a poll: `v` is the first attempt at word `p`; re-polls until both halves carry `epoch`.
__device__ __forceinline__ float2 finish_framed(const uint4* p, uint4 v, unsigned int epoch) {
  if (!framed_ok(v, epoch)) v = spin_framed(p, epoch);
  return make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
}

}  // namespace vla
