// Blackwell-native GEMM for the fused-dense family: D[M,N] = op(A) * op(B) with fused epilogues.
//
//   * operands staged in shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B), a KSTAGES-deep mbarrier ring;
//   * tcgen05.mma (kind::f16, cta_group::1, 128 x BN x 16) issued by ONE elected thread, fp32 accumulators in TMEM;
//   * two TMEM accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1;
//   * warp-specialised persistent CTAs (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//     warps 4..7 = epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global).
//
// Either operand may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) so that forward (x W^T), dgrad (dy W)
// and wgrad (dy^T x) all run without materialising a transpose.
//
// Epilogues (reference cuBLASLt call sites in csrc/fused_dense_cuda.cu):
//   EPI_NONE       plain store                               (dgrad, :1104-1129)
//   EPI_BIAS       + bias[N]                                 (EPILOGUE_BIAS, :81)
//   EPI_BIAS_GELU  + bias, GELU; pre-activation saved to aux (EPILOGUE_GELU_AUX_BIAS, :315,336)
//   EPI_DGELU      * gelu'(aux)                              (EPILOGUE_DGELU_BGRAD, :829; the bias-grad is a separate reduction)
//   EPI_ACCUM      + C (beta = 1), fp32 or 16-bit main-grad  (fused_weight_gradient_dense_cuda.cu:42-49)
//   EPI_BIAS_RELU / EPI_BIAS_SIGMOID / EPI_RELU / EPI_SIGMOID (mlp_cuda.cu:95-119,271-405)
#include "common.cuh"
#include <cuda.h>
#include <cstdio>

namespace ab {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 x 16-bit = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;   // 8 warps
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
  int epi;
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
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a_format [7,10), b_format [10,13),
// a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29).
__host__ __device__ inline uint32_t make_idesc(int is_bf16, int a_mn, int b_mn, int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(is_bf16 ? 1 : 0) << 7;
  d |= (uint32_t)(is_bf16 ? 1 : 0) << 10;
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int BN, int KSTAGES>
struct Smem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = KSTAGES * kStage;
  static constexpr int kTotal = kBarOff + 256 + 1024;  // barriers + tmem ptr + alignment slack
};

template <typename TOut, int BN, int KSTAGES>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                           Params p, int is_bf16) {
  using S = Smem<BN, KSTAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + KSTAGES;
  uint64_t* tmem_full = empty_bar + KSTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (p.K + BK - 1) / BK;
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: power of two

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < KSTAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (one elected lane)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t % num_m, n_blk = t / num_m;
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * S::kStage;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStage);
          if (!p.a_mn_major) {
            tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m_blk * BM);            // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; i++)                                              // box {64 m, 64 k} x2
              tma_load_2d(sa + i * (BK * 128), &map_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
          }
          if (!p.b_mn_major) {
            tma_load_2d(sb, &map_b, &full_bar[stage], kb * BK, n_blk * BN);            // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; i++)
              tma_load_2d(sb + i * (BK * 128), &map_b, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
          }
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(is_bf16, p.a_mn_major, p.b_mn_major, BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStage);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            // K-major: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); advance 32 B per UMMA_K inside the atom.
            // MN-major: 64-wide MN blocks BK*128 B apart (LBO), 8-k-row atoms 1024 B apart (SBO); advance 2 atoms per UMMA_K.
            const uint64_t adesc = p.a_mn_major ? make_desc(sa + k * 2048, BK * 128, 1024) : make_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = p.b_mn_major ? make_desc(sb + k * 2048, BK * 128, 1024) : make_desc(sb + k * 32, 16, 1024);
            umma_f16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                 // frees the smem slot when these MMAs have read it
          if (kb == num_k - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================== epilogue: TMEM -> registers -> fused op -> global
    const int q = warp - kEpiWarp0;  // == warp % 4: the TMEM lane quarter this warp may touch
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % num_m, n_blk = t / num_m;
      mbar_wait(&tmem_full[acc], acc_phase, 4);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      TOut* drow = reinterpret_cast<TOut*>(p.D) + (size_t)row * p.ldd;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c0;
        if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
          const bool full = (col0 + 32 <= p.N);
          if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_RELU || p.epi == EPI_BIAS_SIGMOID) {
            const TOut* b = reinterpret_cast<const TOut*>(p.bias) + col0;
#pragma unroll
            for (int j = 0; j < 32; j++) if (full || col0 + j < p.N) v[j] += to_f<TOut>(b[j]);
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
#pragma unroll
            for (int j = 0; j < 32; j++) if (full || col0 + j < p.N) v[j] *= dgelu_f(to_f<TOut>(arow[j]));
          } else if (p.epi == EPI_ACCUM) {
            const TOut* crow = reinterpret_cast<const TOut*>(p.C) + (size_t)row * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 32; j++) if (full || col0 + j < p.N) v[j] += to_f<TOut>(crow[j]);
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
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// 2-D row-major [rows, cols] 16-bit tensor with leading dimension ld (elements); box = {box_cols (inner), box_rows}.
static int make_map(CUtensorMap* m, const void* ptr, int is_bf16, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                    uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1001;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(2000 + (int)r);
}

template <typename TOut, int BN, int KSTAGES>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, int is_bf16, int sms, cudaStream_t st) {
  using S = Smem<BN, KSTAGES>;
  static bool attr_done = false;
  auto kern = gemm_kernel<TOut, BN, KSTAGES>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, kThreads, S::kTotal, st>>>(ma, mb, p, is_bf16);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

}  // namespace gemm
}  // namespace ab

using namespace ab;
using namespace ab::gemm;

// D[M,N] (dt_out) = op(A) op(B), A/B 16-bit (bf16 or fp16, same type).
//   a_mn_major == 0: A is row-major [M, K] with leading dim lda;  == 1: A is row-major [K, M].
//   b_mn_major == 0: B is row-major [N, K] with leading dim ldb;  == 1: B is row-major [K, N].
// Requirements (else returns -10 so the caller can fall back): inner extents and leading dims multiples of 8 elements,
// 16-byte aligned bases.
AB_API int ab_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, long long lda, long long ldb, long long ldd,
                        int a_mn_major, int b_mn_major, int dt_in, int dt_out, int epi, const void* bias, void* aux, long long ldaux,
                        const void* C, long long ldc, int sms, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (dt_in != kBF16 && dt_in != kF16) return -10;
  const int is_bf16 = dt_in == kBF16;
  if ((lda % 8) || (ldb % 8) || !aligned16(A) || !aligned16(B)) return -10;
  if ((!a_mn_major && (K % 8)) || (a_mn_major && (M % 8)) || (!b_mn_major && (K % 8)) || (b_mn_major && (N % 8))) return -10;
  const int BN = (N >= 192 || N > 128) ? 256 : 128;
  CUtensorMap ma, mb;
  int rc;
  if (!a_mn_major) rc = make_map(&ma, A, is_bf16, M, K, lda, BK, BM); else rc = make_map(&ma, A, is_bf16, K, M, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map(&mb, B, is_bf16, N, K, ldb, BK, BN); else rc = make_map(&mb, B, is_bf16, K, N, ldb, 64, BK);
  if (rc) return rc;
  Params p;
  p.M = M; p.N = N; p.K = K; p.D = D; p.ldd = ldd; p.bias = bias; p.aux = aux; p.ldaux = ldaux; p.C = C; p.ldc = ldc;
  p.a_mn_major = a_mn_major; p.b_mn_major = b_mn_major; p.epi = epi;
  if (sms <= 0) sms = kNumSMs;
#define GEMM_GO(T)                                                                      \
  return BN == 256 ? launch<T, 256, 4>(ma, mb, p, is_bf16, sms, st) : launch<T, 128, 6>(ma, mb, p, is_bf16, sms, st)
  if (dt_out == kBF16) { GEMM_GO(bf16); }
  if (dt_out == kF16) { GEMM_GO(f16); }
  if (dt_out == kF32) { GEMM_GO(float); }
  return -1;
}

// Column sums of a row-major [M, N] matrix (bias gradients): out[n] = sum_m x[m, n]. fp32 accumulation, deterministic.
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial(const T* __restrict__ x, float* __restrict__ part, int M, int N, long long ld,
                                                      int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  if (c >= N) return;
  float s = 0.f;
  for (int r = r0; r < r1; r++) s += to_f<T>(x[(size_t)r * ld + c]);
  part[(size_t)blockIdx.y * N + c] = s;
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_final(const float* __restrict__ part, T* __restrict__ out, int N, int nparts) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int k = 0; k < nparts; k++) s += part[(size_t)k * N + c];
  out[c] = from_f<T>(s);
}

// ws: >= 64 * N floats
AB_API int ab_colsum(const void* x, void* out, float* ws, int M, int N, long long ld, int dt, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  int parts = (M + 127) / 128;
  if (parts > 64) parts = 64;
  const int rpb = (M + parts - 1) / parts;
  dim3 grid((N + 255) / 256, parts);
  AB_DISPATCH_FLOAT3(dt, T, colsum_partial<T><<<grid, 256, 0, st>>>((const T*)x, ws, M, N, ld, rpb);
                     colsum_final<T><<<(N + 255) / 256, 256, 0, st>>>(ws, (T*)out, N, parts));
  AB_CHECK_LAUNCH();
  return 0;
}
