// Pieces shared by the tcgen05 attention forward / backward kernels (fmha_fwd_sm100.cu, fmha_bwd_sm100.cu).
#pragma once
#include "gemm_common.cuh"

namespace ab {
namespace fmha {
using namespace ab::gemm;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2: the softmax is exponential-bound, libdevice exp2f adds range fix-ups
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// registers -> TMEM (32 lanes x 32 columns): the accumulator rescale of the online softmax
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two floats -> one packed 16-bit pair (low half = first argument), a single cvt instruction
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<bf16>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <> __device__ __forceinline__ uint32_t pack2<f16>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// byte offset of element (row r, col c) inside a [rows x 64 cols] 16-bit block in the SWIZZLE_128B K-major layout
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
  const int chunk = (c >> 3) ^ (r & 7);
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (c & 7) * 2);
}

// ---- dropout: counter-based Philox4x32-7, one call per 2 x 2 block of the (query, key) score matrix so that both the
// query-per-lane kernels (forward, dQ) and the key-per-lane kernel (dK / dV) get two consecutive elements out of one call.
// counter = (query >> 1, key >> 1, batch * heads + head, offset), key = seed; component (query & 1) * 2 + (key & 1).
struct Dropout {
  uint32_t seed_lo, seed_hi, offset, thresh;  // keep an element iff its random word >= thresh  (thresh = p * 2^32)
  float rp;                                   // 1 / (1 - p)
};
static inline Dropout make_dropout(float p, unsigned long long seed, unsigned long long offset) {
  Dropout d;
  d.seed_lo = (uint32_t)seed; d.seed_hi = (uint32_t)(seed >> 32); d.offset = (uint32_t)offset;
  double t = (double)p * 4294967296.0;
  d.thresh = p <= 0.f ? 0u : (t >= 4294967295.0 ? 4294967295u : (uint32_t)t);
  d.rp = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return d;
}
__device__ __forceinline__ uint4 philox2x2(const Dropout& d, uint32_t q2, uint32_t k2, uint32_t bh) {
  uint32_t c0 = q2, c1 = k2, c2 = bh, c3 = d.offset, k0 = d.seed_lo, k1 = d.seed_hi;
#pragma unroll
  for (int r = 0; r < 7; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t philox_pick(const uint4& r, int comp) {
  return comp == 0 ? r.x : comp == 1 ? r.y : comp == 2 ? r.z : r.w;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D view [rows][heads][d] of a 16-bit tensor: strides in elements; box = {64 d, 1 head, box_rows rows}
static inline int make_map3(CUtensorMap* m, const void* ptr, int is_bf16, long long rows, int heads, int d, long long row_stride,
                            long long head_stride, int box_rows = 128) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) return -1001;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);  // binds the primary context on threads that have none (autograd workers)
  cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)head_stride * 2, (cuuint64_t)row_stride * 2};
  cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(f)(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                                                 const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(2000 + (int)r);
}

}  // namespace fmha
}  // namespace ab
