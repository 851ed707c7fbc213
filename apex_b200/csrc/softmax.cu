// Scaled softmax family (Megatron): plain, padding-masked and causal (upper-triangular) — forward and backward.
// Spec: reference csrc/megatron/scaled_masked_softmax.h:105-461 (mask value -10000, fully masked rows produce 0),
// scaled_upper_triang_masked_softmax.h:129-362 (row q attends to keys 0..q, the rest is written as 0),
// generic_scaled_masked_softmax.h:60-344 (any key length). The reference keeps a whole row per WARP (sk <= 16384 / 4096 masked,
// template switch over log2(sk)); here a row is held by 8..512 threads chosen at run time, so one kernel covers every key
// length up to 32768 with 16-byte accesses, and longer rows stream through an online-softmax kernel.
#include "norm_common.cuh"

namespace ab {

enum { SM_PLAIN = 0, SM_MASKED = 1, SM_CAUSAL = 2 };

template <int MAXV, typename T, int MODE>
__global__ void __launch_bounds__(512) softmax_fwd_vec(const T* __restrict__ x, T* __restrict__ y, const uint8_t* __restrict__ mask,
                                                       float scale, long long rows, int sk, int sq, int heads, int mask_per_batch,
                                                       int tpr) {
  constexpr int E = 16 / sizeof(T);
  __shared__ float sred[128];
  RowReducer red(sred, tpr);
  const int rows_per_cta = blockDim.x / tpr;
  const int nvec = sk / E;
  for (long long row0 = (long long)blockIdx.x * rows_per_cta; row0 < rows; row0 += (long long)gridDim.x * rows_per_cta) {
    const long long row = row0 + red.rg;
    const bool valid = row < rows;
    const int q = (int)(row % sq);
    const int klen = (MODE == SM_CAUSAL) ? q + 1 : sk;  // keys this row may attend to
    const uint8_t* mrow = nullptr;
    if (MODE == SM_MASKED && valid) {
      const long long b = row / ((long long)sq * heads);
      mrow = mask + ((mask_per_batch ? b : 0) * sq + q) * (long long)sk;
    }
    float v[MAXV][E];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int idx = i * tpr + red.lane_r;
      const bool on = valid && idx < nvec && idx * E < klen;
      if (on) {
        load_vec<T, E>(v[i], x + row * sk + (long long)idx * E);
        uint8_t mk[E];
        if (MODE == SM_MASKED) {
          if (E == 8) *reinterpret_cast<uint2*>(mk) = *reinterpret_cast<const uint2*>(mrow + idx * E);
          else *reinterpret_cast<uint32_t*>(mk) = *reinterpret_cast<const uint32_t*>(mrow + idx * E);
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
          float t = v[i][e] * scale;
          if (MODE == SM_MASKED && mk[e] == 1) t = -10000.f;
          if (MODE == SM_CAUSAL && idx * E + e >= klen) t = -INFINITY;
          v[i][e] = t;
          mx = fmaxf(mx, t);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) v[i][e] = -INFINITY;
      }
    }
    mx = red.maxv(mx);
    const float keep = (MODE == SM_MASKED && mx == -10000.f) ? 0.f : 1.f;  // every key masked -> zeros
    float sum = 0.f;
    const float mxs = (mx == -INFINITY) ? 0.f : mx;
#pragma unroll
    for (int i = 0; i < MAXV; i++)
#pragma unroll
      for (int e = 0; e < E; e++) { v[i][e] = __expf(v[i][e] - mxs); sum += v[i][e]; }
    sum = red.sum(sum);
    const float inv = keep / sum;
    if (valid) {
#pragma unroll
      for (int i = 0; i < MAXV; i++) {
        const int idx = i * tpr + red.lane_r;
        if (idx < nvec) {
          float o[E];
#pragma unroll
          for (int e = 0; e < E; e++) o[e] = v[i][e] * inv;  // exp(-inf) = 0 beyond klen
          store_vec<T, E>(y + row * sk + (long long)idx * E, o);
        }
      }
    }
  }
}

template <int MAXV, typename T>
__global__ void __launch_bounds__(512) softmax_bwd_vec(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, float scale,
                                                       long long rows, int sk, int tpr) {
  constexpr int E = 16 / sizeof(T);
  __shared__ float sred[128];
  RowReducer red(sred, tpr);
  const int rows_per_cta = blockDim.x / tpr;
  const int nvec = sk / E;
  for (long long row0 = (long long)blockIdx.x * rows_per_cta; row0 < rows; row0 += (long long)gridDim.x * rows_per_cta) {
    const long long row = row0 + red.rg;
    const bool valid = row < rows;
    float g[MAXV][E], p[MAXV][E];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int idx = i * tpr + red.lane_r;
      if (valid && idx < nvec) {
        load_vec<T, E>(g[i], dy + row * sk + (long long)idx * E);
        load_vec<T, E>(p[i], y + row * sk + (long long)idx * E);
#pragma unroll
        for (int e = 0; e < E; e++) dot += g[i][e] * p[i][e];
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) { g[i][e] = 0.f; p[i][e] = 0.f; }
      }
    }
    dot = red.sum(dot);
    if (valid) {
#pragma unroll
      for (int i = 0; i < MAXV; i++) {
        const int idx = i * tpr + red.lane_r;
        if (idx < nvec) {
          float o[E];
#pragma unroll
          for (int e = 0; e < E; e++) o[e] = scale * p[i][e] * (g[i][e] - dot);
          store_vec<T, E>(dx + row * sk + (long long)idx * E, o);
        }
      }
    }
  }
}

// Any key length / alignment: one CTA per row, online softmax (single pass for max & sum), second pass writes.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) softmax_fwd_generic(const T* __restrict__ x, T* __restrict__ y, const uint8_t* __restrict__ mask,
                                                           float scale, long long rows, int sk, int sq, int heads, int mask_per_batch) {
  __shared__ float red[40];
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int q = (int)(row % sq);
    const int klen = (MODE == SM_CAUSAL) ? q + 1 : sk;
    const uint8_t* mrow = nullptr;
    if (MODE == SM_MASKED) { const long long b = row / ((long long)sq * heads); mrow = mask + ((mask_per_batch ? b : 0) * sq + q) * (long long)sk; }
    float mx = -INFINITY, sum = 0.f;
    for (int i = threadIdx.x; i < klen; i += blockDim.x) {
      float t = to_f<T>(x[row * sk + i]) * scale;
      if (MODE == SM_MASKED && mrow[i] == 1) t = -10000.f;
      const float nm = fmaxf(mx, t);
      sum = sum * __expf(mx - nm) + __expf(t - nm);
      mx = nm;
    }
    const float gmx = block_max(mx, red);
    sum = (mx == -INFINITY) ? 0.f : sum * __expf(mx - gmx);
    sum = block_sum(sum, red);
    const float keep = (MODE == SM_MASKED && gmx == -10000.f) ? 0.f : 1.f;
    const float inv = keep / sum;
    for (int i = threadIdx.x; i < sk; i += blockDim.x) {
      float o = 0.f;
      if (i < klen) {
        float t = to_f<T>(x[row * sk + i]) * scale;
        if (MODE == SM_MASKED && mrow[i] == 1) t = -10000.f;
        o = __expf(t - gmx) * inv;
      }
      y[row * sk + i] = from_f<T>(o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_generic(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, float scale,
                                                           long long rows, int sk) {
  __shared__ float red[40];
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    float dot = 0.f;
    for (int i = threadIdx.x; i < sk; i += blockDim.x) dot += to_f<T>(dy[row * sk + i]) * to_f<T>(y[row * sk + i]);
    dot = block_sum(dot, red);
    for (int i = threadIdx.x; i < sk; i += blockDim.x) {
      const float p = to_f<T>(y[row * sk + i]);
      dx[row * sk + i] = from_f<T>(scale * p * (to_f<T>(dy[row * sk + i]) - dot));
    }
  }
}

template <typename T, int MODE>
int softmax_fwd_launch(const void* x, void* y, const uint8_t* mask, float scale, long long rows, int sk, int sq, int heads,
                       int mask_per_batch, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  const bool vec_ok = (sk % E == 0) && aligned16(x) && aligned16(y) && (MODE != SM_MASKED || (reinterpret_cast<uintptr_t>(mask) % 8 == 0));
  NormCfg c = norm_cfg(vec_ok ? sk / E : 1, 4, 512);
  if (vec_ok && c.ok) {
    long long grid = (rows + c.rows_per_cta - 1) / c.rows_per_cta;
    const long long cap = (long long)kNumSMs * (2048 / c.threads);
    if (grid > cap) grid = cap;
#define SMF(MV) softmax_fwd_vec<MV, T, MODE><<<(int)grid, c.threads, 0, st>>>((const T*)x, (T*)y, mask, scale, rows, sk, sq, heads, mask_per_batch, c.tpr)
    switch (c.maxv) { case 1: SMF(1); break; case 2: SMF(2); break; case 4: SMF(4); break; default: SMF(8); break; }
  } else {
    long long grid = rows < kNumSMs * 8 ? rows : kNumSMs * 8;
    softmax_fwd_generic<T, MODE><<<(int)grid, 256, 0, st>>>((const T*)x, (T*)y, mask, scale, rows, sk, sq, heads, mask_per_batch);
  }
  AB_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int softmax_bwd_launch(const void* dy, const void* y, void* dx, float scale, long long rows, int sk, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  const bool vec_ok = (sk % E == 0) && aligned16(dy) && aligned16(y) && aligned16(dx);
  NormCfg c = norm_cfg(vec_ok ? sk / E : 1, 4, 512);
  if (vec_ok && c.ok && c.maxv <= 4) {
    long long grid = (rows + c.rows_per_cta - 1) / c.rows_per_cta;
    const long long cap = (long long)kNumSMs * (2048 / c.threads);
    if (grid > cap) grid = cap;
#define SMB(MV) softmax_bwd_vec<MV, T><<<(int)grid, c.threads, 0, st>>>((const T*)dy, (const T*)y, (T*)dx, scale, rows, sk, c.tpr)
    switch (c.maxv) { case 1: SMB(1); break; case 2: SMB(2); break; default: SMB(4); break; }
  } else {
    long long grid = rows < kNumSMs * 8 ? rows : kNumSMs * 8;
    softmax_bwd_generic<T><<<(int)grid, 256, 0, st>>>((const T*)dy, (const T*)y, (T*)dx, scale, rows, sk);
  }
  AB_CHECK_LAUNCH();
  return 0;
}

}  // namespace ab

using namespace ab;

// x, y: [rows, sk]; rows = b*heads*sq. mode 0 plain, 1 masked (mask [b or 1, 1, sq, sk] uint8, 1 = masked), 2 causal (sq == sk rows).
AB_API int ab_softmax_fwd(const void* x, void* y, const uint8_t* mask, float scale, long long rows, int sk, int sq, int heads,
                          int mask_per_batch, int mode, int dt, cudaStream_t st) {
  if (rows <= 0 || sk <= 0) return 0;
#define SM_MODE(T)                                                                                              \
  if (mode == SM_PLAIN) return softmax_fwd_launch<T, SM_PLAIN>(x, y, mask, scale, rows, sk, sq, heads, mask_per_batch, st);   \
  if (mode == SM_MASKED) return softmax_fwd_launch<T, SM_MASKED>(x, y, mask, scale, rows, sk, sq, heads, mask_per_batch, st); \
  return softmax_fwd_launch<T, SM_CAUSAL>(x, y, mask, scale, rows, sk, sq, heads, mask_per_batch, st)
  AB_DISPATCH_FLOAT3(dt, T, SM_MODE(T));
  return 0;
}

AB_API int ab_softmax_bwd(const void* dy, const void* y, void* dx, float scale, long long rows, int sk, int dt, cudaStream_t st) {
  if (rows <= 0 || sk <= 0) return 0;
  AB_DISPATCH_FLOAT3(dt, T, return (softmax_bwd_launch<T>(dy, y, dx, scale, rows, sk, st)));
  return 0;
}
