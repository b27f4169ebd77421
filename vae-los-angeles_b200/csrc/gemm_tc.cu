// Grouped bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA into 128-byte-swizzled shared memory, one 128 x BN output tile per CTA.
//
//   mode NT: C[M,N] = A[M,K] * B[N,K]^T   both operands K-major    (Linear forward, data gradients)
//   mode TN: C[M,N] = A[K,M]^T * B[K,N]   both operands MN-major   (weight gradients, split-K + red.add)
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
#include "tc_ptx.cuh"
#include "vla_internal.h"

#include <mutex>

namespace vla {

namespace {

constexpr int A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;                 // 16 KiB
constexpr int B_STAGE_BYTES = GEMM_BN_MAX_TN * GEMM_BK * 2;          // 24 KiB (>= 160 * 128 B)
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;           // 40 KiB
constexpr int ONES_OFFSET = GEMM_STAGES * STAGE_BYTES;               // 2 KiB of bf16 1.0
constexpr int ONES_BYTES = 2048;
constexpr int STAGE_LD = 33;
constexpr int CHUNK_OFFSET = ONES_OFFSET + ONES_BYTES;               // 8 warps x fp32 [32][33] (split-K red path)
constexpr int CHUNK_BYTES = 8 * 32 * STAGE_LD * 4;
constexpr int VEC_OFFSET = CHUNK_OFFSET + CHUNK_BYTES;               // bias | mean | rstd, fp32 [3][192]
constexpr int VEC_BYTES = 3 * GEMM_BN_MAX_TN * 4;
constexpr int PART_OFFSET = VEC_OFFSET + VEC_BYTES;                  // column-stat partials fp32 [2][6][4][32]
constexpr int MAX_CHUNKS = GEMM_BN_MAX_TN / 32;
constexpr int PART_BYTES = 2 * MAX_CHUNKS * 4 * 32 * 4;
constexpr int BAR_OFFSET = PART_OFFSET + PART_BYTES;                 // mbarriers
constexpr int SMEM_USED = BAR_OFFSET + 128;
constexpr int SMEM_BYTES = SMEM_USED + 1024;                         // slack for manual 1 KiB alignment
constexpr int EPI_THREADS = GEMM_THREADS - 64;                       // 8 warps

static_assert(B_STAGE_BYTES >= GEMM_BN_MAX_NT * 128, "B stage too small for NT tiles");
static_assert(STAGE_BYTES % 1024 == 0 && A_STAGE_BYTES % 1024 == 0, "swizzle atoms need 1 KiB alignment");
static_assert(GEMM_BN_MAX_NT % 32 == 0 && GEMM_BN_MAX_NT <= GEMM_BN_MAX_TN, "tile limits");

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Sum over the 32 lanes of v[j] for every j; the total of column j ends up in lane j (31 shuffles).
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = upper ? v[j] : v[j + s];
      const float keep = upper ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmGroup grp) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KiB-aligned base (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* acc_bar = empty_bar + GEMM_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which problem / tile ----
  int pi = 0;
#pragma unroll
  for (int i = 1; i < GEMM_MAX_PROBLEMS; ++i)
    if (i < grp.nprob && static_cast<int>(blockIdx.x) >= grp.p[i].tile_begin) pi = i;
  const GemmProblem& P = grp.p[pi];
  const int local = blockIdx.x - P.tile_begin;
  const int n_tile = local % P.n_tiles;
  const int m_tile = (local / P.n_tiles) % P.m_tiles;
  const int k_split = local / (P.n_tiles * P.m_tiles);
  const int m0 = m_tile * GEMM_BM;
  const int n0 = n_tile * P.BN;
  const int BN = P.BN;
  const int kb_total = (P.K + GEMM_BK - 1) / GEMM_BK;
  const int kb0 = k_split * P.kb_per_split;
  const int kb1 = min(kb0 + P.kb_per_split, kb_total);
  const bool bias_mma = (MODE == 1) && (P.flags & GF_BIASGRAD) && n_tile == 0;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, GEMM_TMEM_COLS);
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    if (bias_mma) {
      // 2 KiB of bf16 1.0: the B operand of the bias-gradient MMA (layout-invariant)
      uint32_t* ones = reinterpret_cast<uint32_t*>(smem + ONES_OFFSET);
      for (int i = et; i < ONES_BYTES / 4; i += EPI_THREADS) ones[i] = 0x3F803F80u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // per-column epilogue vectors of this tile
    float* vec = reinterpret_cast<float*>(smem + VEC_OFFSET);
    for (int i = et; i < BN; i += EPI_THREADS) {
      const int col = n0 + i;
      const bool ok = col < P.N;
      vec[i] = (ok && (P.flags & GF_BIAS)) ? P.bias[col] : 0.f;
      vec[GEMM_BN_MAX_TN + i] = (ok && (P.flags & GF_BNSTATS)) ? P.mean[col] : 0.f;
      vec[2 * GEMM_BN_MAX_TN + i] = (ok && (P.flags & GF_BNSTATS)) ? P.rstd[col] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (MODE == 0) {
          mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + BN * 128);
          tma_load_2d(sa, &P.tmA, &full_bar[stage], kb * GEMM_BK, m0);   // box 64 (K) x 128 (M)
          tma_load_2d(sb, &P.tmB, &full_bar[stage], kb * GEMM_BK, n0);   // box 64 (K) x BN (N)
        } else {
          const int nb = BN >> 6;
          mbar_expect_tx(&full_bar[stage], (2 + nb) * 8192);
          tma_load_2d(sa, &P.tmA, &full_bar[stage], m0, kb * GEMM_BK);           // box 64 (M) x 64 (K)
          tma_load_2d(sa + 8192, &P.tmA, &full_bar[stage], m0 + 64, kb * GEMM_BK);
          for (int i = 0; i < nb; ++i)
            tma_load_2d(sb + i * 8192, &P.tmB, &full_bar[stage], n0 + i * 64, kb * GEMM_BK);
        }
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, MODE, MODE);
      const uint32_t idesc_ones = make_idesc_bf16(GEMM_BM, 16, 1, 0);
      const uint32_t ones_addr = smem_u32(smem + ONES_OFFSET);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          uint64_t adesc, bdesc;
          if (MODE == 0) {
            adesc = make_smem_desc(sa + k * 32, 16, 1024);     // +16 bf16 of K inside the swizzle atom
            bdesc = make_smem_desc(sb + k * 32, 16, 1024);
          } else {
            adesc = make_smem_desc(sa + k * 2048, 8192, 1024);  // +16 K-rows of 128 B
            bdesc = make_smem_desc(sb + k * 2048, 8192, 1024);
          }
          const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
          umma_bf16(tmem_base, adesc, bdesc, idesc, acc);
          if (bias_mma) {
            const uint64_t odesc = make_smem_desc(ones_addr, 16, 1024);
            umma_bf16(tmem_base + GEMM_BIAS_TMEM_COL, adesc, odesc, idesc_ones, acc);
          }
        }
        umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_bar);               // accumulator complete
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    // Thread = one accumulator row (TMEM lane); the two warps that share a lane quarter take alternate
    // 32-column chunks.  Everything stays in registers; global operands are fetched as 128-bit vectors
    // before they are needed; column statistics use a shuffle butterfly.
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // 0: warps 2-5, 1: warps 6-9
    const int et = threadIdx.x - 64;        // 0..255
    const float* vec = reinterpret_cast<const float*>(smem + VEC_OFFSET);
    float* part = reinterpret_cast<float*>(smem + PART_OFFSET);
    const int flags = P.flags;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < P.M;
    const int n_chunks = (BN + 31) >> 5;
    const bool want_stats = (flags & (GF_COLSTATS | GF_BNSTATS)) != 0;

    mbar_wait(acc_bar, 0);
    tc_fence_after();

    for (int c = half; c < n_chunks; c += 2) {
      const int col0 = n0 + c * 32;
      const int nvalid = min(32, P.N - col0);          // <= 0: nothing to store (tile padding)
      uint32_t r[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32);
      tmem_ld16(taddr, r);
      tmem_ld16(taddr + 16, r + 16);
      // ---- global operands of this row, issued before the TMEM data is needed ----
      uint4 mk[4];
      float4 pr[8];
      const bool full = nvalid == 32;
      if ((flags & GF_MASK) && row_ok && full) {
        const uint4* src = reinterpret_cast<const uint4*>(P.mask_src + static_cast<size_t>(row) * P.ld_mask + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) mk[i] = __ldg(src + i);
      }
      if ((flags & GF_BNSTATS) && row_ok && full) {
        const float4* src = reinterpret_cast<const float4*>(P.pre + static_cast<size_t>(row) * P.ld_pre + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) pr[i] = __ldg(src + i);
      }
      tmem_ld_wait();
      float v[32];
      float w[32];                                     // second statistic (v*v or v*xhat)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]) + vec[c * 32 + j];
        if (flags & GF_COLSTATS) w[j] = x * x;
        if (flags & GF_MASK) {
          float m;
          if (full) {
            const uint32_t word = reinterpret_cast<const uint32_t*>(mk)[j >> 1];
            m = __uint_as_float((j & 1) ? (word & 0xFFFF0000u) : (word << 16));
          } else {
            m = (row_ok && j < nvalid) ? __bfloat162float(P.mask_src[static_cast<size_t>(row) * P.ld_mask + col0 + j]) : 0.f;
          }
          x = m > 0.f ? x * P.mask_scale : 0.f;
        }
        if (flags & GF_BNSTATS) {
          float p;
          if (full) p = reinterpret_cast<const float*>(pr)[j];
          else p = (row_ok && j < nvalid) ? P.pre[static_cast<size_t>(row) * P.ld_pre + col0 + j] : 0.f;
          const float xh = (p - vec[GEMM_BN_MAX_TN + c * 32 + j]) * vec[2 * GEMM_BN_MAX_TN + c * 32 + j];
          w[j] = x * xh;
        }
        v[j] = x;
      }
      if (want_stats) {
        // statistics are taken before the activation (BatchNorm forward) / on the masked gradient (backward)
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1[j] = row_ok ? v[j] : 0.f; s2[j] = row_ok ? w[j] : 0.f; }
        const float t1 = warp_column_sums(s1, lane);
        const float t2 = warp_column_sums(s2, lane);
        part[((0 * MAX_CHUNKS + c) * 4 + q) * 32 + lane] = t1;
        part[((1 * MAX_CHUNKS + c) * 4 + q) * 32 + lane] = t2;
      }
      if (flags & (GF_RELU | GF_SIGMOID)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (flags & GF_RELU) v[j] = fmaxf(v[j], 0.f);
          if (flags & GF_SIGMOID) v[j] = sigmoidf_(v[j]);
        }
      }
      if (flags & GF_RED) {
        // split-K partial sums: transpose through this warp's smem patch so that each red covers 32 consecutive
        // floats of one row (one 128-byte L2 atomic request instead of 32 scattered ones)
        float* chunk = reinterpret_cast<float*>(smem + CHUNK_OFFSET) + (warp - 2) * (32 * STAGE_LD);
#pragma unroll
        for (int j = 0; j < 32; ++j) chunk[lane * STAGE_LD + j] = v[j];
        __syncwarp();
        if (lane < nvalid) {
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const int rr = m0 + q * 32 + i;
            if (rr < P.M) red_add_f32(P.out_f32 + static_cast<size_t>(rr) * P.ld_f32 + col0 + lane, chunk[i * STAGE_LD + lane]);
          }
        }
        __syncwarp();
      } else if (row_ok && nvalid > 0) {
        if (flags & GF_OUT_F32) {
          float* dst = P.out_f32 + static_cast<size_t>(row) * P.ld_f32 + col0;
          const bool a16 = full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
          const bool a8 = full && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
          if (a16) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else if (a8) {
#pragma unroll
            for (int i = 0; i < 16; ++i) reinterpret_cast<float2*>(dst)[i] = make_float2(v[2 * i], v[2 * i + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) dst[j] = v[j];
          }
        }
        if (flags & GF_OUT_BF16) {
          bf16* dst = P.out_bf16 + static_cast<size_t>(row) * P.ld_bf16 + col0;
          if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 o;
              __nv_bfloat162 b0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
              __nv_bfloat162 b1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
              __nv_bfloat162 b2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
              __nv_bfloat162 b3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
              o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
              o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
              reinterpret_cast<uint4*>(dst)[i] = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) dst[j] = __float2bfloat16(v[j]);
          }
        }
      }
    }
    if (bias_mma && half == 0) {
      uint32_t r1[1];
      tmem_ld1(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + GEMM_BIAS_TMEM_COL, r1);
      tmem_ld_wait();
      if (row_ok) red_add_f32(P.bias_grad + row, __uint_as_float(r1[0]));
    }
    if (want_stats) {
      named_bar_sync(3, EPI_THREADS);
      for (int i = et; i < 2 * BN; i += EPI_THREADS) {
        const int which = i / BN, cc = i - which * BN;
        const int col = n0 + cc;
        if (col < P.N) {
          const float* pp = part + ((which * MAX_CHUNKS + (cc >> 5)) * 4) * 32 + (cc & 31);
          P.stats[(static_cast<size_t>(m_tile) * 2 + which) * P.N + col] = pp[0] + pp[32] + pp[64] + pp[96];
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

std::once_flag g_attr_once;
cudaError_t g_attr_err = cudaSuccess;

}  // namespace

size_t gemm_smem_bytes() { return SMEM_BYTES; }

cudaError_t launch_gemm_group(const GemmGroup& g, int mode, cudaStream_t stream) {
  std::call_once(g_attr_once, [] {
    g_attr_err = cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (g_attr_err == cudaSuccess)
      g_attr_err = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (g_attr_err != cudaSuccess) return g_attr_err;
  if (g.total_tiles <= 0) return cudaSuccess;
  if (mode == 0)
    gemm_tc_kernel<0><<<g.total_tiles, GEMM_THREADS, SMEM_BYTES, stream>>>(g);
  else
    gemm_tc_kernel<1><<<g.total_tiles, GEMM_THREADS, SMEM_BYTES, stream>>>(g);
  return cudaGetLastError();
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

}  // namespace vla
