// Fused softmax cross-entropy with label smoothing. One pass over the logits in forward (online max / sum-exp, sum of logits and
// the label logit in the same sweep), saving only max+log-sum-exp per row; backward recomputes the softmax from it.
// Spec: reference apex/contrib/csrc/xentropy/xentropy_kernel.cu:111-120,334-370 (forward, loss formula :368), :467-485 (backward,
// written in place over the logits), host :488-611.
//   loss = (mlse - sum(x)/C) * smoothing - (x[label] - mlse) * (1 - smoothing),   mlse = max + log(sum exp(x - max))
//   dlogits = dloss * (exp(x - mlse) - (1 - smoothing) * [i == label] - smoothing / C);   rows with label == padding_idx -> 0
#include "common.cuh"

namespace ab {

constexpr int kXeThreads = 512;

template <typename T>
__global__ void __launch_bounds__(kXeThreads) xentropy_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ labels,
                                                                  float* __restrict__ losses, float* __restrict__ mlse_out, int rows,
                                                                  int C, float smoothing, long long padding_idx) {
  constexpr int E = 16 / sizeof(T);
  __shared__ float red[40];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = logits + (size_t)row * C;
    const long long label = labels[row];
    float mx = -INFINITY, se = 0.f, sx = 0.f;
    const bool vec = (C % E == 0) && aligned16(xr);
    if (vec) {
      for (int i = threadIdx.x * E; i < C; i += kXeThreads * E) {
        float v[E];
        load_vec<T, E>(v, xr + i);
        float lm = v[0];
#pragma unroll
        for (int e = 1; e < E; e++) lm = fmaxf(lm, v[e]);
        const float nm = fmaxf(mx, lm);
        float add = 0.f;
#pragma unroll
        for (int e = 0; e < E; e++) { add += __expf(v[e] - nm); sx += v[e]; }
        se = se * __expf(mx - nm) + add;
        mx = nm;
      }
    } else {
      for (int i = threadIdx.x; i < C; i += kXeThreads) {
        const float v = to_f<T>(xr[i]);
        const float nm = fmaxf(mx, v);
        se = se * __expf(mx - nm) + __expf(v - nm);
        mx = nm; sx += v;
      }
    }
    const float gmx = block_max(mx, red);
    se = (mx == -INFINITY) ? 0.f : se * __expf(mx - gmx);
    se = block_sum(se, red);
    sx = block_sum(sx, red);
    if (threadIdx.x == 0) {
      const float mlse = gmx + logf(se);
      float loss = 0.f;
      if (label != padding_idx) {
        const float xl = (label >= 0 && label < C) ? to_f<T>(xr[label]) : 0.f;
        loss = (mlse - sx / (float)C) * smoothing - (xl - mlse) * (1.f - smoothing);
      }
      losses[row] = loss;
      mlse_out[row] = mlse;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kXeThreads) xentropy_bwd_kernel(const float* __restrict__ grad_loss, const T* __restrict__ logits,
                                                                  const float* __restrict__ mlse, const long long* __restrict__ labels,
                                                                  T* __restrict__ grad_logits, int rows, int C, float smoothing,
                                                                  long long padding_idx) {
  constexpr int E = 16 / sizeof(T);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = logits + (size_t)row * C;
    T* gr = grad_logits + (size_t)row * C;
    const long long label = labels[row];
    const float g = (label == padding_idx) ? 0.f : grad_loss[row];
    const float m = mlse[row];
    const float sm = smoothing / (float)C;
    const bool vec = (C % E == 0) && aligned16(xr) && aligned16(gr);
    if (vec) {
      for (int i = threadIdx.x * E; i < C; i += kXeThreads * E) {
        float v[E], o[E];
        load_vec<T, E>(v, xr + i);
#pragma unroll
        for (int e = 0; e < E; e++) o[e] = g * (__expf(v[e] - m) - ((i + e) == label ? (1.f - smoothing) : 0.f) - sm);
        store_vec<T, E>(gr + i, o);
      }
    } else {
      for (int i = threadIdx.x; i < C; i += kXeThreads)
        gr[i] = from_f<T>(g * (__expf(to_f<T>(xr[i]) - m) - (i == label ? (1.f - smoothing) : 0.f) - sm));
    }
  }
}

}  // namespace ab

using namespace ab;

AB_API int ab_xentropy_fwd(const void* logits, const long long* labels, float* losses, float* mlse, int rows, int C, float smoothing,
                           long long padding_idx, int dt, cudaStream_t st) {
  if (rows <= 0) return 0;
  const int grid = rows < kNumSMs * 4 ? rows : kNumSMs * 4;
  AB_DISPATCH_FLOAT3(dt, T, xentropy_fwd_kernel<T><<<grid, kXeThreads, 0, st>>>((const T*)logits, labels, losses, mlse, rows, C, smoothing, padding_idx));
  AB_CHECK_LAUNCH();
  return 0;
}

// grad_logits may alias logits (in-place, like the reference)
AB_API int ab_xentropy_bwd(const float* grad_loss, const void* logits, const float* mlse, const long long* labels, void* grad_logits,
                           int rows, int C, float smoothing, long long padding_idx, int dt, cudaStream_t st) {
  if (rows <= 0) return 0;
  const int grid = rows < kNumSMs * 4 ? rows : kNumSMs * 4;
  AB_DISPATCH_FLOAT3(dt, T, xentropy_bwd_kernel<T><<<grid, kXeThreads, 0, st>>>(grad_loss, (const T*)logits, mlse, labels, (T*)grad_logits, rows, C, smoothing, padding_idx));
  AB_CHECK_LAUNCH();
  return 0;
}
