// amp_C-equivalent elementwise / reduction multi-tensor ops on the device-table engine:
//   scale, axpby, l2norm (+per-tensor, +unscale, +mp early-exit), l2norm_scale, norm_out (L2 / Linf blend), cast (incl. e5m2)
// Behavioural spec: reference csrc/multi_tensor_scale_kernel.cu:29-111, multi_tensor_axpby_kernel.cu:27-122,
// multi_tensor_l2norm_kernel.cu:28-443, multi_tensor_l2norm_kernel_mp.cu, multi_tensor_l2norm_scale_kernel.cu.
#include "mt_engine.cuh"

namespace ab {

struct NoCtx {};

// ---------------------------------------------------------------- scale
struct ScaleOp {
  static constexpr unsigned kRead = 1u, kWrite = 2u;
  static constexpr int kAcc = 0;
  using Ctx = NoCtx;
  float scale; int* noop; const float* scale_ptr;  // effective scale = scale * (*scale_ptr) when a device scalar is given
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&)[2], int) const {
    bool fin = true;
    const float sc = scale_ptr ? scale * (*scale_ptr) : scale;
#pragma unroll
    for (int j = 0; j < V; j++) { fin = fin && finite_f(r[0][j]); r[1][j] = r[0][j] * sc; }
    if (!fin && noop) *noop = 1;  // benign race: every writer stores the same value
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- axpby
struct AxpbyOp {
  static constexpr unsigned kRead = 3u, kWrite = 4u;
  static constexpr int kAcc = 0;
  using Ctx = NoCtx;
  float a, b; int arg_to_check; int* noop;
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&)[2], int) const {
    bool fin = true;
#pragma unroll
    for (int j = 0; j < V; j++) {
      r[2][j] = a * r[0][j] + b * r[1][j];
      if (arg_to_check == -1) fin = fin && finite_f(r[0][j]) && finite_f(r[1][j]);
      else if (arg_to_check == 0) fin = fin && finite_f(r[0][j]);
      else fin = fin && finite_f(r[1][j]);
    }
    if (!fin) *noop = 1;
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- norms
// mode 0: sum of squares, mode 1: max |x|. Writes one partial per chunk; deterministic second stage below.
template <bool kMax>
struct NormOp {
  static constexpr unsigned kRead = 1u, kWrite = 0u;
  static constexpr int kAcc = 1;
  using Ctx = NoCtx;
  float* partial; const float* inv_scale; const int* noop_in; int* noop_out;
  __device__ bool skip() const { return noop_in != nullptr && *noop_in != 0; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&acc)[2], int) const {
    const float s = inv_scale ? *inv_scale : 1.f;
#pragma unroll
    for (int j = 0; j < V; j++) {
      float x = r[0][j] * s;
      if (kMax) acc[0] = fmaxf(acc[0], fabsf(x)); else acc[0] += x * x;
    }
  }
  __device__ void end(const Ctx&, int, int cid, float (&acc)[2], float* red) const {
    float s = kMax ? block_max(acc[0], red) : block_sum(acc[0], red);
    if (threadIdx.x == 0) {
      partial[cid] = s;
      if (noop_out && !finite_f(s)) *noop_out = 1;
    }
  }
};

// out = in*scale and sum(out^2)
struct NormScaleOp {
  static constexpr unsigned kRead = 1u, kWrite = 2u;
  static constexpr int kAcc = 1;
  using Ctx = NoCtx;
  float* partial; float scale; int* noop;
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&acc)[2], int) const {
    bool fin = true;
#pragma unroll
    for (int j = 0; j < V; j++) {
      fin = fin && finite_f(r[0][j]);
      float o = r[0][j] * scale;
      r[1][j] = o;
      acc[0] += o * o;
    }
    if (!fin) *noop = 1;
  }
  __device__ void end(const Ctx&, int, int cid, float (&acc)[2], float* red) const {
    float s = block_sum(acc[0], red);
    if (threadIdx.x == 0) partial[cid] = s;
  }
};

// Stage 2a: per-tensor. One warp per tensor; fixed summation order => deterministic.
// blend: out = sqrt(a*old^2 + b*n^2) (L2) or a*old + b*n (Linf); a<0 => plain norm.
__global__ void norm_per_tensor_kernel(const float* __restrict__ partial, const int* __restrict__ prefix, int n,
                                       float* __restrict__ per_tensor, int is_max, float a, float b, int blend) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int t = w; t < n; t += nw) {
    const int c0 = prefix[t], c1 = prefix[t + 1];
    float s = 0.f;
    for (int c = c0 + lane; c < c1; c += 32) s = is_max ? fmaxf(s, partial[c]) : s + partial[c];
    s = is_max ? warp_max(s) : warp_sum(s);
    if (lane == 0) {
      float nrm = is_max ? s : sqrtf(s);
      if (blend) {
        float old = per_tensor[t];
        nrm = is_max ? (a * old + b * nrm) : sqrtf(a * old * old + b * nrm * nrm);
      }
      per_tensor[t] = nrm;
    }
  }
}

// Stage 2b: global norm over all chunk partials (single CTA, fixed order).
__global__ void norm_global_kernel(const float* __restrict__ partial, int total, float* __restrict__ out, int is_max,
                                   const int* __restrict__ noop_in) {
  __shared__ float red[40];
  if (noop_in && *noop_in) return;
  float s = 0.f;
  for (int c = threadIdx.x; c < total; c += blockDim.x) s = is_max ? fmaxf(s, partial[c]) : s + partial[c];
  s = is_max ? block_max(s, red) : block_sum(s, red);
  if (threadIdx.x == 0) out[0] = is_max ? s : sqrtf(s);
}

// ---------------------------------------------------------------- cast (maybe_cast_mt equivalent incl. e5m2 <-> float)
template <typename T> struct Raw8 { uint8_t v; };

}  // namespace ab

using namespace ab;

#define TB make_table(arena, n, depth, total_chunks, chunk)

AB_API int ab_mt_scale(void* arena, int n, int depth, int total_chunks, int chunk, int dt_in, int dt_out, float scale,
                       int* noop, const float* scale_ptr, cudaStream_t st) {
  if (depth != 2) return -2;
  ScaleOp op{scale, noop, scale_ptr};
  AB_DISPATCH_FLOAT3(dt_in, TI, AB_DISPATCH_FLOAT3(dt_out, TO, return (mt_launch<4, ScaleOp, TI, TO>(TB, op, st))));
  return 0;
}

AB_API int ab_mt_axpby(void* arena, int n, int depth, int total_chunks, int chunk, int dt_x, int dt_y, int dt_o, float a,
                       float b, int arg_to_check, int* noop, cudaStream_t st) {
  if (depth != 3) return -2;
  AxpbyOp op{a, b, arg_to_check, noop};
  AB_DISPATCH_FLOAT3(
      dt_x, TX,
      AB_DISPATCH_FLOAT3(dt_y, TY, AB_DISPATCH_FLOAT3(dt_o, TO, return (mt_launch<4, AxpbyOp, TX, TY, TO>(TB, op, st)))));
  return 0;
}

// Generic norm: writes out[0] (global) and, if per_tensor != nullptr, per_tensor[t].
//   partial: scratch of >= total_chunks floats. inv_scale: optional device scalar (unscale_l2norm).
//   noop_in: optional early-exit flag (l2norm_mp). is_max: Linf. blend/a/b: novograd running-norm blend.
AB_API int ab_mt_norm(void* arena, int n, int depth, int total_chunks, int chunk, int dt, float* partial, float* out,
                      float* per_tensor, const float* inv_scale, const int* noop_in, int* noop_out, int is_max, int blend,
                      float a, float b, cudaStream_t st) {
  if (depth < 1) return -2;
  MTTable tb = TB;
  int rc = 0;
  if (is_max) {
    NormOp<true> op{partial, inv_scale, noop_in, noop_out};
    AB_DISPATCH_FLOAT3(dt, T, rc = (mt_launch<4, NormOp<true>, T>(tb, op, st)));
  } else {
    NormOp<false> op{partial, inv_scale, noop_in, noop_out};
    AB_DISPATCH_FLOAT3(dt, T, rc = (mt_launch<4, NormOp<false>, T>(tb, op, st)));
  }
  if (rc) return rc;
  if (per_tensor && n > 0) {
    int blocks = (n + 7) / 8; if (blocks > 1184) blocks = 1184;
    norm_per_tensor_kernel<<<blocks, 256, 0, st>>>(partial, tb.chunk_prefix, n, per_tensor, is_max, a, b, blend);
  }
  if (out) norm_global_kernel<<<1, 1024, 0, st>>>(partial, total_chunks, out, is_max, noop_in);
  AB_CHECK_LAUNCH();
  return 0;
}

AB_API int ab_mt_l2norm_scale(void* arena, int n, int depth, int total_chunks, int chunk, int dt_in, int dt_out,
                              float scale, float* partial, float* out, float* per_tensor, int* noop, cudaStream_t st) {
  if (depth != 2) return -2;
  MTTable tb = TB;
  NormScaleOp op{partial, scale, noop};
  int rc = 0;
  AB_DISPATCH_FLOAT3(dt_in, TI, AB_DISPATCH_FLOAT3(dt_out, TO, rc = (mt_launch<4, NormScaleOp, TI, TO>(tb, op, st))));
  if (rc) return rc;
  if (per_tensor && n > 0) {
    int blocks = (n + 7) / 8; if (blocks > 1184) blocks = 1184;
    norm_per_tensor_kernel<<<blocks, 256, 0, st>>>(partial, tb.chunk_prefix, n, per_tensor, 0, 0.f, 0.f, 0);
  }
  if (out) norm_global_kernel<<<1, 1024, 0, st>>>(partial, total_chunks, out, 0, nullptr);
  AB_CHECK_LAUNCH();
  return 0;
}
