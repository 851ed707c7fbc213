// NHWC GroupNorm (+SiLU), "a group fits one CTA" path: when the H*W x (C/G) slab of one (image, group) fits in shared memory
// (<= ~200 KB: every shape of the reference's Blackwell table except the largest), ONE CTA owns the group end to end -- x is read
// from global memory exactly once, statistics are an exact two-pass over the shared-memory copy, the output is produced from it,
// and there is no inter-CTA synchronisation at all. Backward keeps x and dy in shared memory the same way; per-(image, channel)
// sums go to a small global array and the LAST CTA to finish (atomic ticket) folds them into dgamma / dbeta.
// Groups larger than one CTA's shared memory are split by rows over a thread-block CLUSTER of 2 or 4 CTAs: the partial sums travel
// through distributed shared memory (st.shared::cluster into every peer's slot + barrier.cluster), so x is still read once.
// Anything else (odd C/G, > 4 x 200 KB) takes the dependency-driven persistent kernel in group_norm.cu.
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

// VC (8, 4 or 2) consecutive channels as one 4 / 8 / 16-byte chunk
template <int BYTES> struct RawOf;
template <> struct RawOf<4> { using type = uint32_t; };
template <> struct RawOf<8> { using type = uint2; };
template <> struct RawOf<16> { using type = uint4; };
template <typename T, int VC> struct ChunkN { using Raw = typename RawOf<VC * sizeof(T)>::type; };
template <typename T, int VC>
__device__ __forceinline__ void unpackN(const typename ChunkN<T, VC>::Raw& r, float (&v)[VC]) {
  const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
  for (int j = 0; j < VC; j++) v[j] = to_f<T>(e[j]);
}
template <typename T, int VC>
__device__ __forceinline__ typename ChunkN<T, VC>::Raw packN(const float (&v)[VC]) {
  typename ChunkN<T, VC>::Raw r;
  T* e = reinterpret_cast<T*>(&r);
#pragma unroll
  for (int j = 0; j < VC; j++) e[j] = from_f<T>(v[j]);
  return r;
}

// ---- thread-block cluster helpers (cluster size 1 degenerates to the CTA itself)
__device__ __forceinline__ uint32_t gs_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t gs_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void gs_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// A CTA may store into a peer's shared memory only once the peer is running: every cluster kernel ARRIVES at a cluster barrier first thing
// and WAITS on it right before its first remote store (the loads in between hide the latency).
__device__ __forceinline__ void gs_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void gs_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void gs_st_cluster(float* local_addr, uint32_t cta, float v) {  // store into CTA `cta`'s copy of a shared variable
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_addr), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// sum of one value per CTA over the cluster; xslot: __shared__ float[8] (one slot per rank), a different array per use
__device__ __forceinline__ float gs_cluster_sum(float v, float* xslot, uint32_t K, uint32_t rank) {
  if (K == 1) return v;
  if (threadIdx.x < K) gs_st_cluster(xslot + rank, threadIdx.x, v);
  gs_cluster_sync();
  float t = 0.f;
  for (uint32_t r = 0; r < K; r++) t += xslot[r];
  return t;
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

// grid = N*G*K CTAs, cluster = K CTAs per (image, group), each owning a contiguous range of rows.
// Thread (ch, lane): ch = its VC-channel chunk of the group (fixed), lane strides over the CTA's rows.
template <typename T, int VC, bool SILU>
__global__ void __launch_bounds__(kGsThreads, 2) gn_group_fwd(const T* __restrict__ x, T* __restrict__ y, const void* __restrict__ gamma,
                                                            const void* __restrict__ beta, int w_fp32, float* __restrict__ mean,
                                                            float* __restrict__ rstd, int HW, int C, int G, float eps) {
  using Raw = typename ChunkN<T, VC>::Raw;
  extern __shared__ uint4 gs_smem_raw[];
  Raw* xs = reinterpret_cast<Raw*>(gs_smem_raw);
  __shared__ float red[32];
  __shared__ float xs0[8], xs1[8];
  const uint32_t K = gs_cluster_size(), rank = gs_cluster_rank();
  if (K > 1) gs_cluster_arrive();   // started: peers may write into this CTA's shared memory after their matching wait
  const int Cg = C / G, nchunk = Cg / VC, lanes = kGsThreads / nchunk;
  const int ng = blockIdx.x / K, n = ng / G, g = ng - n * G;
  const int row0 = (int)((long long)HW * rank / K), rows = (int)((long long)HW * (rank + 1) / K) - row0;
  const int ch = threadIdx.x % nchunk, lane = threadIdx.x / nchunk;
  const bool active = lane < lanes;
  const T* xg = x + ((size_t)n * HW + row0) * C + (size_t)g * Cg + ch * VC;
  T* yg = y + ((size_t)n * HW + row0) * C + (size_t)g * Cg + ch * VC;
  float s = 0.f;
  if (active) {
#pragma unroll 8
    for (int r = lane; r < rows; r += lanes) {
      const Raw v = *reinterpret_cast<const Raw*>(xg + (size_t)r * C);
      xs[r * nchunk + ch] = v;
      float f[VC]; unpackN<T, VC>(v, f);
#pragma unroll
      for (int j = 0; j < VC; j++) s += f[j];
    }
  }
  const float inv_m = 1.f / ((float)HW * (float)Cg);
  if (K > 1) gs_cluster_wait();
  const float mu = gs_cluster_sum(gs_block_sum(s, red), xs0, K, rank) * inv_m;
  float ss = 0.f;
  if (active) {
#pragma unroll 4
    for (int r = lane; r < rows; r += lanes) {
      float f[VC]; unpackN<T, VC>(xs[r * nchunk + ch], f);
#pragma unroll
      for (int j = 0; j < VC; j++) { const float d = f[j] - mu; ss = fmaf(d, d, ss); }
    }
  }
  const float rs = rsqrtf(gs_cluster_sum(gs_block_sum(ss, red), xs1, K, rank) * inv_m + eps);
  if (threadIdx.x == 0 && rank == 0) { mean[ng] = mu; rstd[ng] = rs; }
  if (active) {
    float A[VC], B[VC];
#pragma unroll
    for (int j = 0; j < VC; j++) {
      const int c = g * Cg + ch * VC + j;
      const float gm = gamma ? gs_ld_w<T>(gamma, w_fp32, c) : 1.f, bt = beta ? gs_ld_w<T>(beta, w_fp32, c) : 0.f;
      A[j] = rs * gm; B[j] = fmaf(-mu, A[j], bt);
    }
#pragma unroll 4
    for (int r = lane; r < rows; r += lanes) {
      float f[VC], o[VC]; unpackN<T, VC>(xs[r * nchunk + ch], f);
#pragma unroll
      for (int j = 0; j < VC; j++) { const float v = fmaf(f[j], A[j], B[j]); o[j] = SILU ? gs_silu(v) : v; }
      *reinterpret_cast<Raw*>(yg + (size_t)r * C) = packN<T, VC>(o);
    }
  }
  if (K > 1) gs_cluster_sync();  // nobody exits while a peer may still write into its shared memory
}

// chan: [N][C][2] (sum g, sum g*xhat per image and channel); ticket: zero on entry, left zero.
template <typename T, int VC, bool SILU>
__global__ void __launch_bounds__(kGsThreads, 2) gn_group_bwd(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                            const void* __restrict__ gamma, const void* __restrict__ beta, int w_fp32,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float* __restrict__ chan, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            unsigned int* __restrict__ ticket, int N, int HW, int C, int G, int max_rows) {
  using Raw = typename ChunkN<T, VC>::Raw;
  extern __shared__ uint4 gs_smem_raw[];
  const uint32_t K = gs_cluster_size(), rank = gs_cluster_rank();
  if (K > 1) gs_cluster_arrive();   // started: peers may write into this CTA's shared memory after their matching wait
  const int Cg = C / G, nchunk = Cg / VC, lanes = kGsThreads / nchunk;
  Raw* xs = reinterpret_cast<Raw*>(gs_smem_raw);
  Raw* gs = xs + (size_t)max_rows * nchunk;
  float* part = reinterpret_cast<float*>(gs + (size_t)max_rows * nchunk);  // [lanes][Cg][2]
  float* csum = part + (size_t)lanes * Cg * 2;                              // [K][Cg][2]: every cluster CTA's per-channel sums
  __shared__ float red[32];
  __shared__ int s_last;
  const int ng = blockIdx.x / K, n = ng / G, g = ng - n * G;
  const int row0 = (int)((long long)HW * rank / K), rows = (int)((long long)HW * (rank + 1) / K) - row0;
  const int ch = threadIdx.x % nchunk, lane = threadIdx.x / nchunk;
  const bool active = lane < lanes;
  const size_t base = ((size_t)n * HW + row0) * C + (size_t)g * Cg + ch * VC;
  const float mu = mean[ng], rs = rstd[ng], mr = -mu * rs;
  float gm[VC], bt[VC], db[VC], dg[VC];
#pragma unroll
  for (int j = 0; j < VC; j++) {
    const int c = g * Cg + ch * VC + j;
    gm[j] = (active && gamma) ? gs_ld_w<T>(gamma, w_fp32, c) : 1.f;
    bt[j] = (active && beta) ? gs_ld_w<T>(beta, w_fp32, c) : 0.f;
    db[j] = 0.f; dg[j] = 0.f;
  }
  if (active) {
#pragma unroll 4
    for (int r = lane; r < rows; r += lanes) {
      const Raw xv = *reinterpret_cast<const Raw*>(x + base + (size_t)r * C);
      const Raw gv = *reinterpret_cast<const Raw*>(dy + base + (size_t)r * C);
      xs[r * nchunk + ch] = xv; gs[r * nchunk + ch] = gv;
      float xf[VC], gf[VC]; unpackN<T, VC>(xv, xf); unpackN<T, VC>(gv, gf);
#pragma unroll
      for (int j = 0; j < VC; j++) {
        const float xh = fmaf(xf[j], rs, mr);
        float gg = gf[j];
        if (SILU) gg *= gs_dsilu(fmaf(xh, gm[j], bt[j]));
        db[j] += gg; dg[j] = fmaf(gg, xh, dg[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < VC; j++) {
      part[((size_t)lane * Cg + ch * VC + j) * 2] = db[j];
      part[((size_t)lane * Cg + ch * VC + j) * 2 + 1] = dg[j];
    }
  }
  __syncthreads();
  if (K > 1) gs_cluster_wait();
  // per-channel sums of this CTA's rows -> slot [rank] of every cluster CTA's csum
  for (int c = threadIdx.x; c < Cg; c += kGsThreads) {
    float s0 = 0.f, s1 = 0.f;
    for (int l = 0; l < lanes; l++) { s0 += part[((size_t)l * Cg + c) * 2]; s1 += part[((size_t)l * Cg + c) * 2 + 1]; }
    for (uint32_t r = 0; r < K; r++) {
      if (K == 1) { csum[c * 2] = s0; csum[c * 2 + 1] = s1; }
      else { gs_st_cluster(csum + ((size_t)rank * Cg + c) * 2, r, s0); gs_st_cluster(csum + ((size_t)rank * Cg + c) * 2 + 1, r, s1); }
    }
  }
  if (K > 1) gs_cluster_sync(); else __syncthreads();
  float m1 = 0.f, m2 = 0.f;
  for (int c = threadIdx.x; c < Cg; c += kGsThreads) {
    float s0 = 0.f, s1 = 0.f;
    for (uint32_t r = 0; r < K; r++) { s0 += csum[((size_t)r * Cg + c) * 2]; s1 += csum[((size_t)r * Cg + c) * 2 + 1]; }
    const int cc = g * Cg + c;
    if (rank == 0) { chan[((size_t)n * C + cc) * 2] = s0; chan[((size_t)n * C + cc) * 2 + 1] = s1; }
    const float gmc = gamma ? gs_ld_w<T>(gamma, w_fp32, cc) : 1.f;
    m1 += gmc * s0; m2 += gmc * s1;
  }
  const float inv_m = 1.f / ((float)HW * (float)Cg);
  m1 = gs_block_sum(m1, red) * inv_m;
  m2 = gs_block_sum(m2, red) * inv_m;
  if (active) {
    float RG[VC];
#pragma unroll
    for (int j = 0; j < VC; j++) RG[j] = rs * gm[j];
    const float NM1 = -rs * m1, NM2 = -rs * m2;
#pragma unroll 2
    for (int r = lane; r < rows; r += lanes) {
      float xf[VC], gf[VC], o[VC];
      unpackN<T, VC>(xs[r * nchunk + ch], xf); unpackN<T, VC>(gs[r * nchunk + ch], gf);
#pragma unroll
      for (int j = 0; j < VC; j++) {
        const float xh = fmaf(xf[j], rs, mr);
        float gg = gf[j];
        if (SILU) gg *= gs_dsilu(fmaf(xh, gm[j], bt[j]));
        o[j] = fmaf(gg, RG[j], fmaf(xh, NM2, NM1));
      }
      *reinterpret_cast<Raw*>(dx + base + (size_t)r * C) = packN<T, VC>(o);
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
        // parameter gradients in the parameters' own dtype (fp32 weights, or the activation type): no cast kernels afterwards
        if (w_fp32 || sizeof(T) == 4) { dgamma[c] = gsum; if (dbeta) dbeta[c] = b; }
        else { reinterpret_cast<T*>(dgamma)[c] = from_f<T>(gsum); if (dbeta) reinterpret_cast<T*>(dbeta)[c] = from_f<T>(b); }
      }
      if (threadIdx.x == 0) *ticket = 0u;
    }
  }
  if (K > 1) gs_cluster_sync();
}

// shared-memory plan for (HW, Cg): chunk width, cluster size, dynamic bytes. Returns false when the group does not qualify.
struct GsPlan { int vc, k, max_rows; size_t dyn; };
static bool gs_plan(int is_bwd, int HW, int C, int G, size_t esz, GsPlan& p) {
  if (G <= 0 || C % G != 0) return false;
  const int Cg = C / G;
  p.vc = (esz == 2 && Cg % 8 == 0) ? 8 : (Cg % 4 == 0) ? 4 : ((Cg % 2 == 0) ? 2 : 0);
  if (p.vc == 0 || Cg / p.vc > kGsThreads) return false;
  const int lanes = kGsThreads / (Cg / p.vc);
  // smallest cluster whose per-CTA footprint lets two CTAs share an SM (<= 100 KB); else the smallest that fits at all (<= 200 KB)
  for (size_t limit : {(size_t)100 * 1024, (size_t)200 * 1024}) {
    for (int k = 1; k <= 8; k *= 2) {
      if (HW < k) break;
      const int max_rows = (HW + k - 1) / k;
      const size_t slab = (size_t)max_rows * Cg * esz;
      const size_t dyn = is_bwd ? 2 * slab + ((size_t)lanes * Cg * 2 + (size_t)k * Cg * 2) * sizeof(float) : slab;
      if (dyn <= limit) { p.k = k; p.max_rows = max_rows; p.dyn = (dyn + 15) / 16 * 16; return true; }
    }
  }
  return false;
}

template <typename KernT, typename... Args>
static int gs_launch(KernT kern, int ctas, int k, size_t dyn, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(kGsThreads); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = k; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, args...);
  return (int)e;
}

template <typename T, int VC>
static int gn_small_launch(const GsPlan& p, int is_bwd, const void* x, const void* dy, void* out, const void* gamma, const void* beta, int w_fp32,
                           float* mean, float* rstd, float* dgamma, float* dbeta, float* chan, unsigned int* ticket, int N, int HW, int C,
                           int G, float eps, int silu, cudaStream_t st) {
  const int ctas = N * G * p.k;
  if (!is_bwd) {
    if (silu) return gs_launch(gn_group_fwd<T, VC, true>, ctas, p.k, p.dyn, st, (const T*)x, (T*)out, gamma, beta, w_fp32, mean, rstd, HW, C, G, eps);
    return gs_launch(gn_group_fwd<T, VC, false>, ctas, p.k, p.dyn, st, (const T*)x, (T*)out, gamma, beta, w_fp32, mean, rstd, HW, C, G, eps);
  }
  if (silu) return gs_launch(gn_group_bwd<T, VC, true>, ctas, p.k, p.dyn, st, (const T*)x, (const T*)dy, (T*)out, gamma, beta, w_fp32,
                             (const float*)mean, (const float*)rstd, chan, dgamma, dbeta, ticket, N, HW, C, G, p.max_rows);
  return gs_launch(gn_group_bwd<T, VC, false>, ctas, p.k, p.dyn, st, (const T*)x, (const T*)dy, (T*)out, gamma, beta, w_fp32,
                   (const float*)mean, (const float*)rstd, chan, dgamma, dbeta, ticket, N, HW, C, G, p.max_rows);
}

}  // namespace ab

using namespace ab;

// Returns 1 when (HW, C, G, dtype) qualifies for the group-per-CTA(-cluster) path (the caller then uses ab_group_norm_small), else 0.
AB_API int ab_group_norm_small_ok(int is_bwd, int HW, int C, int G, int dt) {
  GsPlan p;
  return gs_plan(is_bwd, HW, C, G, dt == kF32 ? 4 : 2, p) ? 1 : 0;
}

// chan: N*C*2 floats of scratch (bwd); ticket: one zero-initialised uint (bwd).
AB_API int ab_group_norm_small(int is_bwd, const void* x, const void* dy, void* out, const void* gamma, const void* beta, int w_fp32, float* mean,
                               float* rstd, float* dgamma, float* dbeta, float* chan, unsigned int* ticket, int N, int HW, int C, int G,
                               float eps, int silu, int dt, cudaStream_t st) {
  if (N <= 0 || HW <= 0) return 0;
  GsPlan p;
  if (!gs_plan(is_bwd, HW, C, G, dt == kF32 ? 4 : 2, p)) return -2;
  if (!aligned16(x) || !aligned16(out) || (is_bwd && !aligned16(dy))) return -3;
#define GS_VC(T, V) gn_small_launch<T, V>(p, is_bwd, x, dy, out, gamma, beta, w_fp32, mean, rstd, dgamma, dbeta, chan, ticket, N, HW, C, G, eps, silu, st)
  if (dt == kF32) return p.vc == 4 ? GS_VC(float, 4) : GS_VC(float, 2);
  if (dt == kF16) return p.vc == 8 ? GS_VC(f16, 8) : p.vc == 4 ? GS_VC(f16, 4) : GS_VC(f16, 2);
  if (dt == kBF16) return p.vc == 8 ? GS_VC(bf16, 8) : p.vc == 4 ? GS_VC(bf16, 4) : GS_VC(bf16, 2);
  return -1;
}
