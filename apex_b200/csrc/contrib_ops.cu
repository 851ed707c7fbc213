// Small fused contrib ops: sigmoid focal loss (forward + partial gradient in one pass, in-place backward scaling) and
// index_mul_2d (out = in1[idx] * in2 without materialising the gather; backward scatters with fp32 atomics).
// Spec: reference apex/contrib/csrc/focal_loss/focal_loss_cuda_kernel.cu:19-190 and
// apex/contrib/csrc/index_mul_2d/index_mul_2d_cuda_kernel.cu:7-207.
#include "common.cuh"

namespace ab {

// loss = sum over (example, real class) of alpha_t * (1 - p_t)^gamma * BCE(p, y_smoothed) / num_positives
// targets: class id per example, -1 = background (all negatives), -2 = ignore.
template <typename T>
__global__ void __launch_bounds__(256) focal_fwd_kernel(const T* __restrict__ x, const long long* __restrict__ tgt, const float* __restrict__ npos,
                                                       T* __restrict__ pgrad, float* __restrict__ loss_partial, long long n_ex, int n_cls,
                                                       int n_real, float alpha, float gamma, float smooth) {
  __shared__ float red[40];
  const float nn = 1.f - smooth * 0.5f, np = smooth * 0.5f, pn = smooth - smooth * 0.5f, pp = 1.f - smooth + smooth * 0.5f;
  float acc = 0.f;
  const long long total = n_ex * n_cls;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long ex = i / n_cls;
    const int c = (int)(i - ex * n_cls);
    const long long y = tgt[ex];
    float g = 0.f;
    if (y != -2 && c < n_real) {
      const float p = to_f<T>(x[i]);
      const float sigma = 1.f / (1.f + __expf(-p));
      const float off_a = (p >= 0.f ? 0.f : -p) + log1pf(__expf(-fabsf(p)));  // softplus(-p)
      float base = smooth > 0.f ? nn * p : p, off_b = (smooth > 0.f ? np : 0.f) - sigma;
      float f1 = 1.f - alpha, f2 = sigma, b1 = gamma, b2 = 1.f - sigma;
      if (y >= 0 && c == y) {
        base = smooth > 0.f ? pn * p : 0.f; off_b = (smooth > 0.f ? pp : 1.f) - sigma;
        f1 = alpha; f2 = 1.f - sigma; b1 = -gamma; b2 = sigma;
      }
      const float cf = f1 * powf(f2, gamma);
      acc += cf * (base + off_a);
      g = cf * (b1 * b2 * (base + off_a) - off_b);
    }
    pgrad[i] = from_f<T>(g);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) loss_partial[blockIdx.x] = s / npos[0];
}

__global__ void focal_sum_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  __shared__ float red[40];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

template <typename T>
__global__ void __launch_bounds__(256) focal_bwd_kernel(T* __restrict__ pgrad, const float* __restrict__ gloss, const float* __restrict__ npos,
                                                       long long total) {
  const float s = gloss[0] / npos[0];
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256)
    pgrad[i] = from_f<T>(to_f<T>(pgrad[i]) * s);
}

// ---- index_mul_2d: out[i, :] = in1[idx[i], :] * in2[i, :]
template <typename T>
__global__ void __launch_bounds__(256) index_mul_fwd_kernel(const T* __restrict__ in1, const T* __restrict__ in2, const long long* __restrict__ idx,
                                                           T* __restrict__ out, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    out[i] = from_f<T>(to_f<T>(in1[idx[r] * cols + c]) * to_f<T>(in2[i]));
  }
}
// grad_in2 = grad_out * in1[idx]; grad_in1[idx[i]] += grad_out[i] * in2[i]  (fp32 accumulation buffer)
template <typename T>
__global__ void __launch_bounds__(256) index_mul_bwd_kernel(const T* __restrict__ in1, const T* __restrict__ in2, const long long* __restrict__ idx,
                                                           const T* __restrict__ gout, float* __restrict__ gin1_f32, T* __restrict__ gin2,
                                                           long long rows, int cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float g = to_f<T>(gout[i]);
    const long long j = idx[r] * cols + c;
    gin2[i] = from_f<T>(g * to_f<T>(in1[j]));
    atomicAdd(gin1_f32 + j, g * to_f<T>(in2[i]));
  }
}

}  // namespace ab

using namespace ab;

AB_API int ab_focal_loss_fwd(const void* x, const long long* tgt, const float* npos, void* pgrad, float* loss_partial, float* loss,
                             long long n_ex, int n_cls, int n_real, float alpha, float gamma, float smooth, int dt, cudaStream_t st) {
  const long long total = n_ex * n_cls;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (grid < 1) grid = 1;
  AB_DISPATCH_FLOAT3(dt, T, focal_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, tgt, npos, (T*)pgrad, loss_partial, n_ex, n_cls, n_real, alpha, gamma, smooth));
  focal_sum_kernel<<<1, 256, 0, st>>>(loss_partial, grid, loss);
  AB_CHECK_LAUNCH();
  return 0;
}
AB_API int ab_focal_loss_bwd(void* pgrad, const float* gloss, const float* npos, long long total, int dt, cudaStream_t st) {
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (grid < 1) grid = 1;
  AB_DISPATCH_FLOAT3(dt, T, focal_bwd_kernel<T><<<grid, 256, 0, st>>>((T*)pgrad, gloss, npos, total));
  AB_CHECK_LAUNCH();
  return 0;
}
AB_API int ab_index_mul_2d_fwd(const void* in1, const void* in2, const long long* idx, void* out, long long rows, int cols, int dt,
                               cudaStream_t st) {
  int grid = (int)((rows * cols + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (grid < 1) grid = 1;
  AB_DISPATCH_FLOAT3(dt, T, index_mul_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)in1, (const T*)in2, idx, (T*)out, rows, cols));
  AB_CHECK_LAUNCH();
  return 0;
}
AB_API int ab_index_mul_2d_bwd(const void* in1, const void* in2, const long long* idx, const void* gout, float* gin1_f32, void* gin2,
                               long long rows, int cols, int dt, cudaStream_t st) {
  int grid = (int)((rows * cols + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (grid < 1) grid = 1;
  AB_DISPATCH_FLOAT3(dt, T, index_mul_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)in1, (const T*)in2, idx, (const T*)gout, gin1_f32, (T*)gin2, rows, cols));
  AB_CHECK_LAUNCH();
  return 0;
}
