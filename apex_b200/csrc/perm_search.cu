// 2:4 structured-sparsity channel-permutation search (ASP). Spec: reference
// apex/contrib/sparsity/permutation_search_kernels/CUDA_kernels/permutation_search_kernels.cu:46-632 (sum_after_2_to_4,
// build_permute_map = exhaustive search over stripe groups, build_swap_map, check_permutations).
// Design: |W| is computed once; a stripe group's columns (8 or 12 or 16 of them) are staged in shared memory for a tile of rows
// and every THREAD scores one candidate arrangement of those columns (top-2-of-4 magnitude kept per group of four, branch-free
// min/max network), so 35 / 5775 / 2.6 M candidates x all stripe groups are one launch; a block arg-max and a (group, chunk)
// partial table leave only a tiny final reduction. The same kernel scores single-column swaps (channel_swap strategy) by
// passing the 17-entry "identity + one swap" candidate list, and whole-matrix permutations are scored by perm_eval_kernel.
#include "common.cuh"

namespace ab {

__device__ __forceinline__ float top2of4(float a, float b, float c, float d) {
  const float x1 = fmaxf(a, b), n1 = fminf(a, b), x2 = fmaxf(c, d), n2 = fminf(c, d);
  return fmaxf(x1, x2) + fmaxf(fminf(x1, x2), fmaxf(n1, n2));
}

// out[p] = sum over rows and groups of four of the two largest magnitudes of row[perm[p][4g..4g+3]]
__global__ void __launch_bounds__(256) perm_eval_kernel(const float* __restrict__ m, int R, int C, const int* __restrict__ perms, float* __restrict__ out) {
  __shared__ float red[40];
  const int p = blockIdx.x;
  const int* pm = perms ? perms + (long long)p * C : nullptr;
  const int G = C / 4;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < (long long)R * G; i += 256) {
    const int r = (int)(i / G), g = (int)(i - (long long)r * G);
    const float* row = m + (long long)r * C;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = fabsf(row[pm ? pm[4 * g + k] : 4 * g + k]);
    acc += top2of4(v[0], v[1], v[2], v[3]);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[p] = acc;
}

constexpr int kMaxW = 16;      // columns per stripe group
constexpr int kRowTile = 256;  // rows staged per pass (16 KB of shared memory at W = 16)

// grid (num_groups, cand_chunks), 256 threads: thread t scores candidate chunk*256 + t of its stripe group.
// groups[g][0..S) are stripe ids (a stripe = 4 adjacent columns); cands[c][0..4S) index into the gathered 4S columns.
// part_val / part_idx [num_groups][cand_chunks]: best score of the chunk and its candidate id.
__global__ void __launch_bounds__(256) stripe_search_kernel(const float* __restrict__ m, int R, int C, const int* __restrict__ groups, int S,
                                                           const unsigned char* __restrict__ cands, int P, float* __restrict__ part_val,
                                                           int* __restrict__ part_idx) {
  __shared__ float tile[kRowTile * kMaxW];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int W = 4 * S, g = blockIdx.x, cand = blockIdx.y * 256 + threadIdx.x;
  const int* sg = groups + (long long)g * S;
  unsigned char pc[kMaxW];
#pragma unroll
  for (int k = 0; k < kMaxW; k++) pc[k] = (cand < P && k < W) ? cands[(long long)cand * W + k] : (unsigned char)k;
  float acc = 0.f;
  for (int r0 = 0; r0 < R; r0 += kRowTile) {
    const int rows = min(kRowTile, R - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * W; i += 256) {
      const int r = i / W, c = i - r * W;
      tile[r * kMaxW + c] = fabsf(m[(long long)(r0 + r) * C + sg[c >> 2] * 4 + (c & 3)]);
    }
    __syncthreads();
    if (cand < P) {
      for (int r = 0; r < rows; r++) {
        const float* row = tile + r * kMaxW;
#pragma unroll
        for (int q = 0; q < kMaxW / 4; q++)
          if (q < S) acc += top2of4(row[pc[4 * q]], row[pc[4 * q + 1]], row[pc[4 * q + 2]], row[pc[4 * q + 3]]);
      }
    }
  }
  // block arg-max (ties -> lowest candidate id, so the identity wins when nothing improves)
  float v = cand < P ? acc : -INFINITY;
  int id = cand;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, id, o);
    if (v2 > v || (v2 == v && i2 < id)) { v = v2; id = i2; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = v; s_idx[threadIdx.x >> 5] = id; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
      if (s_val[w] > v || (s_val[w] == v && s_idx[w] < id)) { v = s_val[w]; id = s_idx[w]; }
    part_val[(long long)g * gridDim.y + blockIdx.y] = v;
    part_idx[(long long)g * gridDim.y + blockIdx.y] = id;
  }
}

}  // namespace ab

using namespace ab;

AB_API int ab_perm_eval(const float* m, int R, int C, const int* perms, int P, float* out, cudaStream_t s) {
  if (C % 4 != 0) return -2;
  if (P <= 0) return 0;
  perm_eval_kernel<<<P, 256, 0, s>>>(m, R, C, perms, out);
  return (int)cudaGetLastError();
}

AB_API int ab_stripe_search(const float* m, int R, int C, const int* groups, int num_groups, int S, const unsigned char* cands, int P,
                            float* part_val, int* part_idx, cudaStream_t s) {
  if (C % 4 != 0 || S < 1 || 4 * S > kMaxW) return -2;
  if (num_groups <= 0 || P <= 0) return 0;
  const int chunks = (P + 255) / 256;
  if (chunks > 65535) return -3;
  stripe_search_kernel<<<dim3(num_groups, chunks), 256, 0, s>>>(m, R, C, groups, S, cands, P, part_val, part_idx);
  return (int)cudaGetLastError();
}
