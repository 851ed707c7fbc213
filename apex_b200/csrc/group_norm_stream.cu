// GroupNorm (+SiLU) for LARGE channels-last activations: two streaming passes per direction instead of one resident slab per group.
//
// The group-per-CTA(-cluster) kernels (group_norm_small.cu) keep an (image, group) slab in shared memory: above ~100 KB per slab they
// need clusters of 8 CTAs with one 143 KB CTA per SM and become latency bound (0.65 - 0.8x of the reference's group_norm_v2 on its 8
// largest shapes). Here the work is split by ROWS, not by groups, so every SM streams at full width whatever (HW, C, G) is:
//   forward : pass 1  per-(image, channel) sum and sum of squares   -> chan[n][2][C]          (atomics, fp32)
//             pass 2  per-group mean / rstd from the channel sums (prologue, shared memory), y = act(x * A_c + B_c)
//             x is read twice but the second read comes out of the 126 MB L2 for every shape the reference tunes for (<= 63 MB).
//   backward: pass 1  per-(image, channel) sums of dy' and dy' * x   (dy' = dy * silu'(z) when the activation is fused)
//             pass 2  dx = dy' * P_c + x * Q_g + R_g ; dgamma / dbeta accumulated per image by the first row-block
// Reference: apex/contrib/csrc/group_norm_v2/gn_cuda_kernel.cuh:195,596 (its Blackwell kernels: one pass with cooperative groups /
// clusters per group tile) and apex/contrib/csrc/group_norm/ (v1 two-pass NHWC kernels). Layout: [N, HW, C] contiguous.
#include "common.cuh"

namespace ab {

constexpr int kGsT = 256;
constexpr int kGsMaxG = 64;
constexpr int kGsU = 4;      // rows in flight per thread

struct GnsArgs {
  const void* x; const void* dy; void* out;
  const void* gamma; const void* beta; int w_fp32;
  float* mean; float* rstd;        // [N, G]
  float* chan;                     // [N][2][C] fp32 scratch, zero on entry of pass 1
  float* dgamma; float* dbeta;     // [C] fp32, zero on entry (bwd)
  int N, HW, C, G; float eps; int silu;
};

template <typename T> __device__ __forceinline__ float ldw(const void* p, int i, int w_fp32) {
  return w_fp32 ? reinterpret_cast<const float*>(p)[i] : to_f<T>(reinterpret_cast<const T*>(p)[i]);
}
template <typename T, int V> __device__ __forceinline__ void gs_load(float (&r)[V], const T* p) {
  if constexpr (V == 1) r[0] = to_f<T>(p[0]); else load_vec<T, V>(r, p);
}
template <typename T, int V> __device__ __forceinline__ void gs_store(T* p, const float (&r)[V]) {
  if constexpr (V == 1) p[0] = from_f<T>(r[0]); else store_vec<T, V>(p, r);
}
__host__ __device__ inline int gs_cols(int cvecs) { return cvecs < kGsT ? cvecs : kGsT; }

__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + __expf(-z)); }

// ---- pass 1 (both directions): per-(image, channel) sums. BWD: (dy', dy' * x); FWD: (x, x^2).
template <typename T, int V, bool BWD, bool SILU>
__global__ void __launch_bounds__(kGsT) gns_stats(const __grid_constant__ GnsArgs a) {
  __shared__ float red[2][kGsT][V + 1];
  __shared__ float s_mean[kGsMaxG], s_rstd[kGsMaxG];
  const int n = blockIdx.y;
  const int cvecs = a.C / V, cols = gs_cols(cvecs), rpc = kGsT / cols;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const bool active = tr < rpc;
  const int cpg = a.C / a.G;
  const T* x = reinterpret_cast<const T*>(a.x) + (size_t)n * a.HW * a.C;
  const T* dy = reinterpret_cast<const T*>(a.dy) + (size_t)n * a.HW * a.C;
  if (BWD && SILU) {
    if (threadIdx.x < a.G) { s_mean[threadIdx.x] = a.mean[n * a.G + threadIdx.x]; s_rstd[threadIdx.x] = a.rstd[n * a.G + threadIdx.x]; }
    __syncthreads();
  }
  for (int cv0 = 0; cv0 < cvecs; cv0 += cols) {
    const int cv = cv0 + tc;
    const bool on = active && cv < cvecs;
    const int c0 = cv * V;
    float s1[V], s2[V], A[V], B[V];
#pragma unroll
    for (int j = 0; j < V; j++) {
      s1[j] = 0.f; s2[j] = 0.f; A[j] = 1.f; B[j] = 0.f;
      if (BWD && SILU && on) {   // z = x * A + B
        const int g = (c0 + j) / cpg;
        const float ga = a.gamma ? ldw<T>(a.gamma, c0 + j, a.w_fp32) : 1.f, be = a.beta ? ldw<T>(a.beta, c0 + j, a.w_fp32) : 0.f;
        A[j] = s_rstd[g] * ga; B[j] = be - s_mean[g] * A[j];
      }
    }
    if (on) {
      const int step = gridDim.x * rpc;
      for (int r0 = blockIdx.x * rpc + tr; r0 < a.HW; r0 += kGsU * step) {
        float xv[kGsU][V], g[kGsU][V];
#pragma unroll
        for (int u = 0; u < kGsU; u++) {   // kGsU independent rows in flight per thread
          const int r = r0 + u * step;
          if (r < a.HW) {
            const size_t off = (size_t)r * a.C + c0;
            gs_load<T, V>(xv[u], x + off);
            if (BWD) gs_load<T, V>(g[u], dy + off);
          }
        }
#pragma unroll
        for (int u = 0; u < kGsU; u++) {
          if (r0 + u * step >= a.HW) continue;
          if (!BWD) {
#pragma unroll
            for (int j = 0; j < V; j++) { s1[j] += xv[u][j]; s2[j] = fmaf(xv[u][j], xv[u][j], s2[j]); }
          } else {
#pragma unroll
            for (int j = 0; j < V; j++) {
              float gj = g[u][j];
              if (SILU) { const float z = fmaf(xv[u][j], A[j], B[j]), sg = sigmoid_f(z); gj *= sg * (1.f + z * (1.f - sg)); }
              s1[j] += gj; s2[j] = fmaf(gj, xv[u][j], s2[j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < V; j++) { red[0][threadIdx.x][j] = s1[j]; red[1][threadIdx.x][j] = s2[j]; }
    __syncthreads();
    if (tr == 0 && cv < cvecs) {
      float* dst = a.chan + (size_t)n * 2 * a.C;
#pragma unroll
      for (int j = 0; j < V; j++) {
        float t1 = 0.f, t2 = 0.f;
        for (int k = 0; k < rpc; k++) { t1 += red[0][k * cols + tc][j]; t2 += red[1][k * cols + tc][j]; }
        atomicAdd(dst + c0 + j, t1);
        atomicAdd(dst + a.C + c0 + j, t2);
      }
    }
    __syncthreads();
  }
}

// ---- forward pass 2: group statistics from the channel sums, then y = act(x * A_c + B_c)
template <typename T, int V, bool SILU>
__global__ void __launch_bounds__(kGsT) gns_fwd_apply(const __grid_constant__ GnsArgs a) {
  __shared__ float s_mean[kGsMaxG], s_rstd[kGsMaxG];
  const int n = blockIdx.y;
  const int cpg = a.C / a.G;
  const float* ch = a.chan + (size_t)n * 2 * a.C;
  {  // one warp-slice per group: G <= 64 groups, 256 threads -> 4 threads per group at G = 64
    const int tpg = kGsT / kGsMaxG;                       // 4
    const int g = threadIdx.x / tpg, l = threadIdx.x % tpg;
    float s1 = 0.f, s2 = 0.f;
    if (g < a.G) for (int c = l; c < cpg; c += tpg) { s1 += __ldcg(ch + g * cpg + c); s2 += __ldcg(ch + a.C + g * cpg + c); }
#pragma unroll
    for (int o = tpg / 2; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (g < a.G && l == 0) {
      const float inv = 1.f / ((float)a.HW * (float)cpg), mu = s1 * inv;
      const float rs = rsqrtf(fmaxf(s2 * inv - mu * mu, 0.f) + a.eps);
      s_mean[g] = mu; s_rstd[g] = rs;
      if (blockIdx.x == 0) { a.mean[n * a.G + g] = mu; a.rstd[n * a.G + g] = rs; }
    }
  }
  __syncthreads();
  const int cvecs = a.C / V, cols = gs_cols(cvecs), rpc = kGsT / cols;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  if (tr >= rpc) return;
  const T* x = reinterpret_cast<const T*>(a.x) + (size_t)n * a.HW * a.C;
  T* y = reinterpret_cast<T*>(a.out) + (size_t)n * a.HW * a.C;
  for (int cv = tc; cv < cvecs; cv += cols) {
    const int c0 = cv * V;
    float A[V], B[V];
#pragma unroll
    for (int j = 0; j < V; j++) {
      const int g = (c0 + j) / cpg;
      const float ga = a.gamma ? ldw<T>(a.gamma, c0 + j, a.w_fp32) : 1.f, be = a.beta ? ldw<T>(a.beta, c0 + j, a.w_fp32) : 0.f;
      A[j] = s_rstd[g] * ga; B[j] = be - s_mean[g] * A[j];
    }
    const int step = gridDim.x * rpc;
    for (int r0 = blockIdx.x * rpc + tr; r0 < a.HW; r0 += kGsU * step) {
      float v[kGsU][V];
#pragma unroll
      for (int u = 0; u < kGsU; u++) if (r0 + u * step < a.HW) gs_load<T, V>(v[u], x + (size_t)(r0 + u * step) * a.C + c0);
#pragma unroll
      for (int u = 0; u < kGsU; u++) {
        if (r0 + u * step >= a.HW) continue;
#pragma unroll
        for (int j = 0; j < V; j++) {
          float z = fmaf(v[u][j], A[j], B[j]);
          if (SILU) z *= sigmoid_f(z);
          v[u][j] = z;
        }
        gs_store<T, V>(y + (size_t)(r0 + u * step) * a.C + c0, v[u]);
      }
    }
  }
}

// ---- backward pass 2: dx = dy' * P_c + x * Q_g + R_g; the first row-block of every image also folds its channel sums into dgamma / dbeta
template <typename T, int V, bool SILU>
__global__ void __launch_bounds__(kGsT) gns_bwd_apply(const __grid_constant__ GnsArgs a) {
  __shared__ float s_mean[kGsMaxG], s_rstd[kGsMaxG], s_q[kGsMaxG], s_r[kGsMaxG];
  const int n = blockIdx.y;
  const int cpg = a.C / a.G;
  const float* ch = a.chan + (size_t)n * 2 * a.C;   // [0]: sum dy', [1]: sum dy' * x
  {
    const int tpg = kGsT / kGsMaxG;
    const int g = threadIdx.x / tpg, l = threadIdx.x % tpg;
    float ag = 0.f, bg = 0.f;
    float mu = 0.f, rs = 0.f;
    if (g < a.G) {
      mu = a.mean[n * a.G + g]; rs = a.rstd[n * a.G + g];
      for (int c = l; c < cpg; c += tpg) {
        const int cc = g * cpg + c;
        const float ga = a.gamma ? ldw<T>(a.gamma, cc, a.w_fp32) : 1.f;
        const float A = __ldcg(ch + cc), B = __ldcg(ch + a.C + cc);
        ag += ga * A;
        bg += ga * (B - mu * A) * rs;
      }
    }
#pragma unroll
    for (int o = tpg / 2; o > 0; o >>= 1) { ag += __shfl_xor_sync(0xffffffffu, ag, o); bg += __shfl_xor_sync(0xffffffffu, bg, o); }
    if (g < a.G && l == 0) {
      const float invM = 1.f / ((float)a.HW * (float)cpg);
      s_mean[g] = mu; s_rstd[g] = rs;
      s_q[g] = -rs * rs * bg * invM;
      s_r[g] = -rs * ag * invM + mu * rs * rs * bg * invM;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && (a.dgamma || a.dbeta)) {
    for (int c = threadIdx.x; c < a.C; c += kGsT) {
      const int g = c / cpg;
      const float A = __ldcg(ch + c), B = __ldcg(ch + a.C + c);
      if (a.dbeta) atomicAdd(a.dbeta + c, A);
      if (a.dgamma) atomicAdd(a.dgamma + c, (B - s_mean[g] * A) * s_rstd[g]);
    }
  }
  const int cvecs = a.C / V, cols = gs_cols(cvecs), rpc = kGsT / cols;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  if (tr >= rpc) return;
  const T* x = reinterpret_cast<const T*>(a.x) + (size_t)n * a.HW * a.C;
  const T* dy = reinterpret_cast<const T*>(a.dy) + (size_t)n * a.HW * a.C;
  T* dx = reinterpret_cast<T*>(a.out) + (size_t)n * a.HW * a.C;
  for (int cv = tc; cv < cvecs; cv += cols) {
    const int c0 = cv * V;
    float P[V], Q[V], R[V], A[V], B[V];
#pragma unroll
    for (int j = 0; j < V; j++) {
      const int g = (c0 + j) / cpg;
      const float ga = a.gamma ? ldw<T>(a.gamma, c0 + j, a.w_fp32) : 1.f, be = a.beta ? ldw<T>(a.beta, c0 + j, a.w_fp32) : 0.f;
      P[j] = s_rstd[g] * ga; Q[j] = s_q[g]; R[j] = s_r[g];
      A[j] = P[j]; B[j] = be - s_mean[g] * P[j];
    }
    const int step = gridDim.x * rpc;
    for (int r0 = blockIdx.x * rpc + tr; r0 < a.HW; r0 += 2 * step) {
      float xv[2][V], g[2][V];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (r0 + u * step < a.HW) {
          const size_t off = (size_t)(r0 + u * step) * a.C + c0;
          gs_load<T, V>(xv[u], x + off);
          gs_load<T, V>(g[u], dy + off);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (r0 + u * step >= a.HW) continue;
#pragma unroll
        for (int j = 0; j < V; j++) {
          float gj = g[u][j];
          if (SILU) { const float z = fmaf(xv[u][j], A[j], B[j]), sg = sigmoid_f(z); gj *= sg * (1.f + z * (1.f - sg)); }
          g[u][j] = fmaf(gj, P[j], fmaf(xv[u][j], Q[j], R[j]));
        }
        gs_store<T, V>(dx + (size_t)(r0 + u * step) * a.C + c0, g[u]);
      }
    }
  }
}

// One FULL wave per kernel: the CTAs of these kernels live for a handful of loop iterations, so a grid that needs a second, partly
// filled wave costs almost 2x (the first version sized the grid for 4 CTAs per SM while the kernels hold 77 - 126 registers, i.e. 2 - 3
// resident CTAs: gns_stats took 44 us for 63 MB). Blocks per image = what the occupancy calculator says fits, divided over the batch.
template <auto KERN> static int gns_blocks_per_image(const GnsArgs& a, int rpc) {
  static int per_sm = 0;   // one copy per kernel (the kernel is the template argument)
  if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KERN, kGsT, 0) != cudaSuccess || per_sm < 1)) per_sm = 2;
  long long cap = (long long)kNumSMs * per_sm / (a.N > 0 ? a.N : 1);
  long long blocks = ((long long)a.HW + rpc - 1) / rpc;   // at least one row per row-slot
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

template <typename T, int V>
static int gns_launch(const GnsArgs& a, int is_bwd, cudaStream_t st) {
  const int cols = gs_cols(a.C / V), rpc = kGsT / cols;
  cudaError_t e = cudaMemsetAsync(a.chan, 0, sizeof(float) * 2 * (size_t)a.N * a.C, st);
  if (e != cudaSuccess) return (int)e;
#define GNS_GO(KERN) { const dim3 grid((unsigned)gns_blocks_per_image<KERN>(a, rpc), (unsigned)a.N); KERN<<<grid, kGsT, 0, st>>>(a); }
  if (!is_bwd) {
    GNS_GO((gns_stats<T, V, false, false>));
    if (a.silu) GNS_GO((gns_fwd_apply<T, V, true>)) else GNS_GO((gns_fwd_apply<T, V, false>))
  } else {
    if (a.dgamma && (e = cudaMemsetAsync(a.dgamma, 0, sizeof(float) * a.C, st)) != cudaSuccess) return (int)e;
    if (a.dbeta && (e = cudaMemsetAsync(a.dbeta, 0, sizeof(float) * a.C, st)) != cudaSuccess) return (int)e;
    if (a.silu) { GNS_GO((gns_stats<T, V, true, true>)); GNS_GO((gns_bwd_apply<T, V, true>)); }
    else { GNS_GO((gns_stats<T, V, true, false>)); GNS_GO((gns_bwd_apply<T, V, false>)); }
  }
#undef GNS_GO
  return (int)cudaGetLastError();
}

}  // namespace ab

using namespace ab;

// x / dy / out: [N, HW, C] contiguous (channels-last); gamma / beta: [C] of the activation dtype (or fp32 when w_fp32); mean / rstd: fp32
// [N, G] (forward: outputs; backward: inputs); chan: fp32 scratch of 2 * N * C floats; dgamma / dbeta: fp32 [C] (backward, may be null).
AB_API int ab_group_norm_stream(int is_bwd, const void* x, const void* dy, void* out, const void* gamma, const void* beta, int w_fp32,
                                float* mean, float* rstd, float* dgamma, float* dbeta, float* chan, int N, int HW, int C, int G, float eps,
                                int silu, int dt, cudaStream_t st) {
  if (N <= 0 || HW <= 0 || C <= 0) return 0;
  if (G <= 0 || G > kGsMaxG || C % G) return -2;
  GnsArgs a;
  a.x = x; a.dy = dy; a.out = out; a.gamma = gamma; a.beta = beta; a.w_fp32 = w_fp32; a.mean = mean; a.rstd = rstd; a.chan = chan;
  a.dgamma = dgamma; a.dbeta = dbeta; a.N = N; a.HW = HW; a.C = C; a.G = G; a.eps = eps; a.silu = silu;
  const bool al = aligned16(x) && aligned16(out) && (!is_bwd || aligned16(dy));
  if (dt == kF32) return (al && C % 4 == 0) ? gns_launch<float, 4>(a, is_bwd, st) : gns_launch<float, 1>(a, is_bwd, st);
  if (dt == kF16) return (al && C % 8 == 0) ? gns_launch<f16, 8>(a, is_bwd, st) : gns_launch<f16, 1>(a, is_bwd, st);
  if (dt == kBF16) return (al && C % 8 == 0) ? gns_launch<bf16, 8>(a, is_bwd, st) : gns_launch<bf16, 1>(a, is_bwd, st);
  return -1;
}
