// NHWC GroupNorm (+SiLU), "a group fits one CTA" path: when the H*W x (C/G) slab of one (image, group) fits in shared memory
// (<= ~200 KB: every shape of the reference's Blackwell table except the largest), ONE CTA owns the group end to end -- x is read
// from global memory exactly once, statistics are an exact two-pass over the shared-memory copy, the output is produced from it,
// and there is no inter-CTA synchronisation at all. Backward keeps x and dy in shared memory the same way; per-(image, channel)
// sums go to a small global array and the LAST CTA to finish (atomic ticket) folds them into dgamma / dbeta.
// Larger groups take the dependency-driven persistent kernel in group_norm.cu.
// Spec: apex/contrib/csrc/group_norm_v2/gn_cuda_kernel.cuh:195,596 (block-sync strategy when a group fits one CTA).
#include "common.cuh"

namespace ab {

constexpr int kGsThreads = 512;

template <typename T>
__device__ __forceinline__ float gs_ld_w(const void* p, int fp32, int c) {
  return fp32 ? reinterpret_cast<const float*>(p)[c] : to_f<T>(reinterpret_cast<const T*>(p)[c]);
}
__device__ __forceinline__ float gs_silu(float v) { return v * __frcp_rn(1.f + __expf(-v)); }
__device__ __forceinline__ float gs_dsilu(float v) { const float s = __frcp_rn(1.f + __expf(-v)); return s * (1.f + v * (1.f - s)); }

// 4 consecutive channels as one 8-byte (16-bit types) or 16-byte (fp32) chunk
template <typename T> struct Chunk4;
template <> struct Chunk4<float> { using Raw = uint4; };
template <> struct Chunk4<bf16> { using Raw = uint2; };
template <> struct Chunk4<f16> { using Raw = uint2; };
template <typename T>
__device__ __forceinline__ void unpack4(const typename Chunk4<T>::Raw& r, float (&v)[4]) {
  const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
  for (int j = 0; j < 4; j++) v[j] = to_f<T>(e[j]);
}
template <typename T>
__device__ __forceinline__ typename Chunk4<T>::Raw pack4(const float (&v)[4]) {
  typename Chunk4<T>::Raw r;
  T* e = reinterpret_cast<T*>(&r);
#pragma unroll
  for (int j = 0; j < 4; j++) e[j] = from_f<T>(v[j]);
  return r;
}

__device__ __forceinline__ float gs_block_sum(float v, float* red) {  // red: >= 32 floats; result broadcast to every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x & 31) < (kGsThreads / 32) ? red[threadIdx.x & 31] : 0.f;
  t = warp_sum(t);
  return __shfl_sync(0xffffffffu, t, 0);
}

// grid = N*G. Thread (ch, lane): ch = its 4-channel chunk of the group (fixed), lane strides over rows.
template <typename T, bool SILU>
__global__ void __launch_bounds__(kGsThreads, 1) gn_group_fwd(const T* __restrict__ x, T* __restrict__ y, const void* __restrict__ gamma,
                                                            const void* __restrict__ beta, int w_fp32, float* __restrict__ mean,
                                                            float* __restrict__ rstd, int HW, int C, int G, float eps) {
  using Raw = typename Chunk4<T>::Raw;
  extern __shared__ uint4 gs_smem_raw[];
  Raw* xs = reinterpret_cast<Raw*>(gs_smem_raw);
  __shared__ float red[32];
  const int Cg = C / G, nchunk = Cg >> 2, lanes = kGsThreads / nchunk;
  const int n = blockIdx.x / G, g = blockIdx.x - n * G;
  const int ch = threadIdx.x % nchunk, lane = threadIdx.x / nchunk;
  const bool active = lane < lanes;
  const T* xg = x + (size_t)n * HW * C + (size_t)g * Cg + ch * 4;
  T* yg = y + (size_t)n * HW * C + (size_t)g * Cg + ch * 4;
  float s = 0.f;
  if (active) {
#pragma unroll 4
    for (int r = lane; r < HW; r += lanes) {
      const Raw v = *reinterpret_cast<const Raw*>(xg + (size_t)r * C);
      xs[r * nchunk + ch] = v;
      float f[4]; unpack4<T>(v, f);
      s += (f[0] + f[1]) + (f[2] + f[3]);
    }
  }
  const float inv_m = 1.f / ((float)HW * (float)Cg);
  const float mu = gs_block_sum(s, red) * inv_m;
  float ss = 0.f;
  if (active) {
#pragma unroll 4
    for (int r = lane; r < HW; r += lanes) {
      float f[4]; unpack4<T>(xs[r * nchunk + ch], f);
#pragma unroll
      for (int j = 0; j < 4; j++) { const float d = f[j] - mu; ss = fmaf(d, d, ss); }
    }
  }
  const float rs = rsqrtf(gs_block_sum(ss, red) * inv_m + eps);
  if (threadIdx.x == 0) { mean[blockIdx.x] = mu; rstd[blockIdx.x] = rs; }
  if (active) {
    float A[4], B[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = g * Cg + ch * 4 + j;
      const float gm = gamma ? gs_ld_w<T>(gamma, w_fp32, c) : 1.f, bt = beta ? gs_ld_w<T>(beta, w_fp32, c) : 0.f;
      A[j] = rs * gm; B[j] = fmaf(-mu, A[j], bt);
    }
#pragma unroll 4
    for (int r = lane; r < HW; r += lanes) {
      float f[4], o[4]; unpack4<T>(xs[r * nchunk + ch], f);
#pragma unroll
      for (int j = 0; j < 4; j++) { const float v = fmaf(f[j], A[j], B[j]); o[j] = SILU ? gs_silu(v) : v; }
      *reinterpret_cast<Raw*>(yg + (size_t)r * C) = pack4<T>(o);
    }
  }
}

// chan: [N][C][2] (sum g, sum g*xhat per image and channel); ticket: zero on entry, left zero.
template <typename T, bool SILU>
__global__ void __launch_bounds__(kGsThreads, 1) gn_group_bwd(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                            const void* __restrict__ gamma, const void* __restrict__ beta, int w_fp32,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float* __restrict__ chan, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            unsigned int* __restrict__ ticket, int N, int HW, int C, int G) {
  using Raw = typename Chunk4<T>::Raw;
  extern __shared__ uint4 gs_smem_raw[];
  const int Cg = C / G, nchunk = Cg >> 2, lanes = kGsThreads / nchunk;
  Raw* xs = reinterpret_cast<Raw*>(gs_smem_raw);
  Raw* gs = xs + (size_t)HW * nchunk;
  float* part = reinterpret_cast<float*>(gs + (size_t)HW * nchunk);  // [lanes][Cg][2]
  __shared__ float red[32];
  __shared__ int s_last;
  const int n = blockIdx.x / G, g = blockIdx.x - n * G;
  const int ch = threadIdx.x % nchunk, lane = threadIdx.x / nchunk;
  const bool active = lane < lanes;
  const size_t base = (size_t)n * HW * C + (size_t)g * Cg + ch * 4;
  const float mu = mean[blockIdx.x], rs = rstd[blockIdx.x], mr = -mu * rs;
  float gm[4], bt[4], db[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = g * Cg + ch * 4 + j;
    gm[j] = (active && gamma) ? gs_ld_w<T>(gamma, w_fp32, c) : 1.f;
    bt[j] = (active && beta) ? gs_ld_w<T>(beta, w_fp32, c) : 0.f;
  }
  if (active) {
#pragma unroll 2
    for (int r = lane; r < HW; r += lanes) {
      const Raw xv = *reinterpret_cast<const Raw*>(x + base + (size_t)r * C);
      const Raw gv = *reinterpret_cast<const Raw*>(dy + base + (size_t)r * C);
      xs[r * nchunk + ch] = xv; gs[r * nchunk + ch] = gv;
      float xf[4], gf[4]; unpack4<T>(xv, xf); unpack4<T>(gv, gf);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float xh = fmaf(xf[j], rs, mr);
        float gg = gf[j];
        if (SILU) gg *= gs_dsilu(fmaf(xh, gm[j], bt[j]));
        db[j] += gg; dg[j] = fmaf(gg, xh, dg[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      part[((size_t)lane * Cg + ch * 4 + j) * 2] = db[j];
      part[((size_t)lane * Cg + ch * 4 + j) * 2 + 1] = dg[j];
    }
  }
  __syncthreads();
  float m1 = 0.f, m2 = 0.f;
  if ((int)threadIdx.x < Cg) {
    float s0 = 0.f, s1 = 0.f;
    for (int l = 0; l < lanes; l++) { s0 += part[((size_t)l * Cg + threadIdx.x) * 2]; s1 += part[((size_t)l * Cg + threadIdx.x) * 2 + 1]; }
    const int c = g * Cg + threadIdx.x;
    chan[((size_t)n * C + c) * 2] = s0;
    chan[((size_t)n * C + c) * 2 + 1] = s1;
    const float gmc = gamma ? gs_ld_w<T>(gamma, w_fp32, c) : 1.f;
    m1 = gmc * s0; m2 = gmc * s1;
  }
  const float inv_m = 1.f / ((float)HW * (float)Cg);
  m1 = gs_block_sum(m1, red) * inv_m;
  m2 = gs_block_sum(m2, red) * inv_m;
  if (active) {
    float RG[4];
#pragma unroll
    for (int j = 0; j < 4; j++) RG[j] = rs * gm[j];
    const float NM1 = -rs * m1, NM2 = -rs * m2;
#pragma unroll 2
    for (int r = lane; r < HW; r += lanes) {
      float xf[4], gf[4], o[4];
      unpack4<T>(xs[r * nchunk + ch], xf); unpack4<T>(gs[r * nchunk + ch], gf);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float xh = fmaf(xf[j], rs, mr);
        float gg = gf[j];
        if (SILU) gg *= gs_dsilu(fmaf(xh, gm[j], bt[j]));
        o[j] = fmaf(gg, RG[j], fmaf(xh, NM2, NM1));
      }
      *reinterpret_cast<Raw*>(dx + base + (size_t)r * C) = pack4<T>(o);
    }
  }
  // the last CTA to finish folds the per-image channel sums into dgamma / dbeta
  if (dgamma) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1); }
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int c = threadIdx.x; c < C; c += kGsThreads) {
        float b = 0.f, gsum = 0.f;
        for (int nn = 0; nn < N; nn++) { b += __ldcg(chan + ((size_t)nn * C + c) * 2); gsum += __ldcg(chan + ((size_t)nn * C + c) * 2 + 1); }
        dgamma[c] = gsum;
        if (dbeta) dbeta[c] = b;
      }
      if (threadIdx.x == 0) *ticket = 0u;
    }
  }
}

template <typename T>
static int gn_small_launch(int is_bwd, const void* x, const void* dy, void* out, const void* gamma, const void* beta, int w_fp32, float* mean,
                           float* rstd, float* dgamma, float* dbeta, float* chan, unsigned int* ticket, int N, int HW, int C, int G, float eps,
                           int silu, cudaStream_t st) {
  const int Cg = C / G, nchunk = Cg / 4, lanes = kGsThreads / nchunk;
  const size_t slab = (size_t)HW * nchunk * sizeof(typename Chunk4<T>::Raw);
  const size_t dyn = is_bwd ? 2 * slab + (size_t)lanes * Cg * 2 * sizeof(float) : slab;
#define GS_GO(KERN, ...)                                                                                  \
  do {                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);   \
    if (e != cudaSuccess) return (int)e;                                                                  \
    KERN<<<N * G, kGsThreads, dyn, st>>>(__VA_ARGS__);                                                    \
  } while (0)
  if (!is_bwd) {
    if (silu) GS_GO((gn_group_fwd<T, true>), (const T*)x, (T*)out, gamma, beta, w_fp32, mean, rstd, HW, C, G, eps);
    else GS_GO((gn_group_fwd<T, false>), (const T*)x, (T*)out, gamma, beta, w_fp32, mean, rstd, HW, C, G, eps);
  } else {
    if (silu) GS_GO((gn_group_bwd<T, true>), (const T*)x, (const T*)dy, (T*)out, gamma, beta, w_fp32, mean, rstd, chan, dgamma, dbeta, ticket, N, HW, C, G);
    else GS_GO((gn_group_bwd<T, false>), (const T*)x, (const T*)dy, (T*)out, gamma, beta, w_fp32, mean, rstd, chan, dgamma, dbeta, ticket, N, HW, C, G);
  }
  return (int)cudaGetLastError();
}

}  // namespace ab

using namespace ab;

// Returns 1 when (HW, C, G, dtype) qualifies for the one-CTA-per-group path (the caller then uses ab_group_norm_small), else 0.
AB_API int ab_group_norm_small_ok(int is_bwd, int HW, int C, int G, int dt) {
  if (C % G != 0) return 0;
  const int Cg = C / G;
  if (Cg % 4 != 0 || Cg > kGsThreads) return 0;
  const size_t esz = dt == kF32 ? 4 : 2;
  const int lanes = kGsThreads / (Cg / 4);
  const size_t slab = (size_t)HW * Cg * esz;
  const size_t dyn = is_bwd ? 2 * slab + (size_t)lanes * Cg * 2 * sizeof(float) : slab;
  return dyn <= 200 * 1024 && (size_t)C * esz % 16 == 0 ? 1 : 0;
}

// chan: N*C*2 floats of scratch (bwd); ticket: one zero-initialised uint (bwd).
AB_API int ab_group_norm_small(int is_bwd, const void* x, const void* dy, void* out, const void* gamma, const void* beta, int w_fp32, float* mean,
                               float* rstd, float* dgamma, float* dbeta, float* chan, unsigned int* ticket, int N, int HW, int C, int G,
                               float eps, int silu, int dt, cudaStream_t st) {
  if (N <= 0 || HW <= 0) return 0;
  if (!ab_group_norm_small_ok(is_bwd, HW, C, G, dt)) return -2;
  if (!aligned16(x) || !aligned16(out) || (is_bwd && !aligned16(dy))) return -3;
  int rc = -1;
  if (dt == kF32) rc = gn_small_launch<float>(is_bwd, x, dy, out, gamma, beta, w_fp32, mean, rstd, dgamma, dbeta, chan, ticket, N, HW, C, G, eps, silu, st);
  else if (dt == kF16) rc = gn_small_launch<f16>(is_bwd, x, dy, out, gamma, beta, w_fp32, mean, rstd, dgamma, dbeta, chan, ticket, N, HW, C, G, eps, silu, st);
  else if (dt == kBF16) rc = gn_small_launch<bf16>(is_bwd, x, dy, out, gamma, beta, w_fp32, mean, rstd, dgamma, dbeta, chan, ticket, N, HW, C, G, eps, silu, st);
  return rc;
}
