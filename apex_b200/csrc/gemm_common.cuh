// Shared pieces of the tcgen05 GEMM kernels: parameters, PTX wrappers, descriptors, the fused epilogue.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstdio>

namespace ab {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 x 16-bit = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 384;   // 12 warps: TMA, MMA, TMEM alloc, spare, 8 epilogue warps
constexpr int kEpiWarp0 = 4;

enum Epi { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_DGELU = 3, EPI_ACCUM = 4, EPI_BIAS_RELU = 5, EPI_BIAS_SIGMOID = 6,
           EPI_RELU = 7, EPI_SIGMOID = 8 };

struct Params {
  int M, N, K;
  void* D; long long ldd;          // output [M, N] row-major
  const void* bias;                // [N] (dtype of D) or null
  void* aux; long long ldaux;      // [M, N] pre-activation (dtype of D): written by BIAS_GELU, read by DGELU
  const void* C; long long ldc;    // accumulate source (EPI_ACCUM), same dtype as D
  int a_mn_major, b_mn_major;
  int a_mn3, b_mn3;                // MN-major operand staged by ONE 3-D TMA box per stage (extent % 128-byte row == 0) instead of 2 / 4 2-D boxes
  int epi;
  float alpha;                     // accumulator scale applied before the epilogue op
  const float* scale_a; const float* scale_b;  // optional DEVICE scalars multiplied into alpha (fp8 dequantisation scales: no host sync)
  // All-gather fused into the GEMM: B is a weight living in a ZeRO parameter buffer whose buckets are still being
  // pushed by the optimizer-step kernels of the D ranks (csrc/dist_adam.cu). Before the TMA producer loads a B tile it acquires the
  // ready flags [bucket][rank] of the buckets that hold the tile's rows; tiles whose buckets have landed start while later buckets are
  // still on the wire. Null => B is plain memory.
  const uint32_t* ready_flags; uint32_t ready_epoch; int ready_world;
  long long ready_w_off;           // element offset of B's first element inside the parameter buffer
  long long ready_bucket_elems;
  float* colsum;                   // optional fp32 [N], zero on entry: += column sums of the values this GEMM stores (fused bias gradient:
                                   // the reference's EPILOGUE_DGELU_BGRAD / BGRADB, csrc/fused_dense_cuda.cu:595,829)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (with a message) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("apex_b200 gemm: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Spin (bounded: traps after ~10 s) until every rank has released `epoch` for buckets [b_lo, b_hi]; then order the generic-proxy acquire
// before the async-proxy (TMA) reads of the data those flags guard.
__device__ __forceinline__ void acquire_buckets(const uint32_t* flags, uint32_t epoch, int world, int b_lo, int b_hi) {
  for (int b = b_lo; b <= b_hi; b++) {
    for (int r = 0; r < world; r++) {
      const uint32_t* slot = flags + (size_t)b * 8 + r;
      uint32_t v;
      long long t0 = clock64();
      while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > 20000000000LL) { printf("apex_b200 gemm: weight bucket %d of rank %d never became ready (epoch %u)\n", b, r, epoch); __trap(); }
        __nanosleep(128);
      }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
// layout: 2 = SWIZZLE_128B (16-byte swizzle atoms, 8-row period); 1 = SWIZZLE_128B_BASE32B (32-byte atoms, 4-row period) -- the only
// layout the tensor core accepts for MN-major 32-bit (tf32) operands; its TMA counterpart is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a_format [7,10), b_format [10,13),
// a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29).
// `is_bf16` doubles as the operand format code: kind::f16 0 = F16, 1 = BF16; kind::f8f6f4 0 = E4M3, 1 = E5M2.
// Bit 0 of the code is A's format, bit 1 set means "B's format differs from A's" (fp8 backward: E5M2 gradients x E4M3 weights / activations).
// fmt == kFmtTF32: both operands are fp32 words read as TF32 (kind::tf32 format code 2).
constexpr int kFmtTF32 = 8;
__host__ __device__ inline uint32_t make_idesc(int fmt, int a_mn, int b_mn, int m, int n) {
  uint32_t d = 0;
  uint32_t fa = (uint32_t)(fmt & 1), fb = ((fmt >> 1) & 1) ? (fa ^ 1u) : fa;
  if (fmt == kFmtTF32) fa = fb = 2u;
  d |= 1u << 4;
  d |= fa << 7;
  d |= fb << 10;
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// erf-based GELU with a 12-instruction erf (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 -- far below 16-bit output resolution):
// the libdevice erff costs ~3x as much and, at 32768 elements per 128 x 256 tile, made the epilogue as long as the tile's MMAs.
// With z = x/sqrt(2): erf(z) = sign * (1 - poly(t) * exp(-z^2)), t = 1/(1 + p*|z|), and exp(-z^2) = exp(-x^2/2) is also the
// Gaussian of gelu'(x), so the derivative needs no second exponential.
__device__ __forceinline__ float erf_core(float az, float e) {  // az = |z|, e = exp(-z*z) -> erf(|z|)
  const float t = __frcp_rn(fmaf(0.3275911f, az, 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  return fmaf(-poly, e, 1.f);
}
__device__ __forceinline__ float gelu_f(float x) {
  const float e = __expf(-0.5f * x * x);
  const float r = copysignf(erf_core(fabsf(x) * 0.70710678118654752f, e), x);
  return 0.5f * x * (1.f + r);
}
__device__ __forceinline__ float dgelu_f(float x) {
  const float e = __expf(-0.5f * x * x);
  const float r = copysignf(erf_core(fabsf(x) * 0.70710678118654752f, e), x);
  return fmaf(x * 0.39894228040143268f, e, 0.5f * (1.f + r));
}

// Epilogue of one accumulator tile for one thread-row: TMEM (32 columns at a time) -> registers -> fused op -> global.
// `half` in {0, 1}: two warps share a TMEM lane quarter q and take one half of the tile's columns each.
template <typename TOut, int BN>
__device__ __forceinline__ void epilogue_tile(const Params& p, uint32_t tmem_base, int acc, int row, int n_blk, int q, int half) {
  const bool row_ok = row < p.M;
  TOut* drow = reinterpret_cast<TOut*>(p.D) + (size_t)row * p.ldd;
  const float alpha = p.alpha * (p.scale_a ? __ldg(p.scale_a) : 1.f) * (p.scale_b ? __ldg(p.scale_b) : 1.f);
#pragma unroll 1
  for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
    tmem_ld_wait();
    const int col0 = n_blk * BN + c0;
    float v[32];
    const bool full = (col0 + 32 <= p.N);
    if (row_ok && col0 < p.N) {
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]) * alpha;
      if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_RELU || p.epi == EPI_BIAS_SIGMOID) {
        const TOut* b = reinterpret_cast<const TOut*>(p.bias) + col0;
        if (full && aligned16(b)) {
          constexpr int VB = 16 / sizeof(TOut);
#pragma unroll
          for (int j = 0; j < 32; j += VB) {
            float t[VB]; load_vec<TOut, VB>(t, b + j);
#pragma unroll
            for (int e = 0; e < VB; e++) v[j + e] += t[e];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) if (full || col0 + j < p.N) v[j] += to_f<TOut>(b[j]);
        }
      }
      if (p.epi == EPI_BIAS_GELU) {
        TOut* arow = reinterpret_cast<TOut*>(p.aux) + (size_t)row * p.ldaux + col0;
        if (full && (sizeof(TOut) * p.ldaux) % 16 == 0) {
          float tmp[8];
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
#pragma unroll
            for (int e = 0; e < 8; e++) tmp[e] = v[j + e];
            store_vec<TOut, 8>(arow + j, tmp);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) if (col0 + j < p.N) arow[j] = from_f<TOut>(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = gelu_f(to_f<TOut>(from_f<TOut>(v[j])));
      } else if (p.epi == EPI_DGELU) {
        const TOut* arow = reinterpret_cast<const TOut*>(p.aux) + (size_t)row * p.ldaux + col0;
        if (full && aligned16(arow)) {
          constexpr int VB = 16 / sizeof(TOut);
#pragma unroll
          for (int j = 0; j < 32; j += VB) {
            float t[VB]; load_vec<TOut, VB>(t, arow + j);
#pragma unroll
            for (int e = 0; e < VB; e++) v[j + e] *= dgelu_f(t[e]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) if (full || col0 + j < p.N) v[j] *= dgelu_f(to_f<TOut>(arow[j]));
        }
      } else if (p.epi == EPI_ACCUM) {
        const TOut* crow = reinterpret_cast<const TOut*>(p.C) + (size_t)row * p.ldc + col0;
        if (full && aligned16(crow)) {
          constexpr int VB = 16 / sizeof(TOut);
#pragma unroll
          for (int j = 0; j < 32; j += VB) {
            float t[VB]; load_vec<TOut, VB>(t, crow + j);
#pragma unroll
            for (int e = 0; e < VB; e++) v[j + e] += t[e];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) if (full || col0 + j < p.N) v[j] += to_f<TOut>(crow[j]);
        }
      } else if (p.epi == EPI_BIAS_RELU || p.epi == EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
      } else if (p.epi == EPI_BIAS_SIGMOID || p.epi == EPI_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 1.f / (1.f + __expf(-v[j]));
      }
      if (full && (sizeof(TOut) * p.ldd) % 16 == 0 && aligned16(p.D)) {
        float tmp[8];
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
#pragma unroll
          for (int e = 0; e < 8; e++) tmp[e] = v[j + e];
          store_vec<TOut, 8>(drow + col0 + j, tmp);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) if (col0 + j < p.N) drow[col0 + j] = from_f<TOut>(v[j]);
      }
    }
    if (p.colsum && col0 < p.N) {   // warp-uniform
      // Fused bias gradient: column sums of this 32-row x 32-column block. Recursive-halving transpose-reduce: after the step with
      // stride s every lane keeps the s columns whose index agrees with its lane bit, so 31 shuffles leave column `lane` on lane `lane`.
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int j = 0; j < 32; j++) if (!row_ok || (!full && col0 + j >= p.N)) v[j] = 0.f;
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; j++) {
          const float send = up ? v[j] : v[j + s];
          const float keep = up ? v[j + s] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, v[0]);
    }
  }
}

}  // namespace gemm
}  // namespace ab
