// EXPERIMENTAL: pieces shared by the tcgen05 attention forward / backward kernels (not yet validated on hardware).
#pragma once
#include "../gemm_common.cuh"

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

// byte offset of element (row r, col c) inside a [rows x 64 cols] 16-bit block in the SWIZZLE_128B K-major layout
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
  const int chunk = (c >> 3) ^ (r & 7);
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + chunk * 16 + (c & 7) * 2);
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
