// RNN-T (transducer) joint and loss for B200.
// Spec: reference apex/contrib/csrc/transducer/transducer_joint_kernel.cu:157-525 (broadcast add f[b,t,:] + g[b,u,:], ReLU, dropout,
// packed output, masked backward reductions) and transducer_loss_kernel.cu:31-385 (alpha / beta lattice recursion, loss, backward
// fused with the log-softmax backward).
// Design: the joint is one grid-stride pass, 16-byte vectors along H, a row per (b,t,u) found by a B-entry binary search when the
// output is packed; the backward does the two reductions (over u for df, over t for dg) in ONE launch. The loss never
// materialises log-softmax: a warp-per-row pass stores only the log-sum-exp (4 bytes per lattice cell), the lattice recursion
// runs one CTA per (utterance, direction) along anti-diagonals, and the backward recomputes softmax from the logits + lse.
#include "common.cuh"

namespace ab {

__device__ __forceinline__ uint32_t hash_rng(uint64_t seed, uint64_t idx) {  // splitmix64 finaliser of (seed, counter)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

struct JointDims {
  int B, T, U, H, packed;
  const int* f_len; const int* g_len; const long long* batch_offset;  // batch_offset: inclusive cumsum of f_len*g_len (packed only)
  long long rows;                                                       // B*T*U or batch_offset[B-1]
};

// row -> (b, t, u, valid)
__device__ __forceinline__ bool joint_decode(const JointDims& d, long long row, int& b, int& t, int& u) {
  if (!d.packed) {
    const long long tu = (long long)d.T * d.U;
    b = (int)(row / tu);
    const int r = (int)(row - b * tu);
    t = r / d.U; u = r - t * d.U;
    return t < d.f_len[b] && u < d.g_len[b];
  }
  int lo = 0, hi = d.B - 1;  // first b with batch_offset[b] > row
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (d.batch_offset[mid] > row) hi = mid; else lo = mid + 1; }
  b = lo;
  const long long start = b ? d.batch_offset[b - 1] : 0;
  const int gl = d.g_len[b];
  const int r = (int)(row - start);
  t = r / gl; u = r - t * gl;
  return true;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) joint_fwd_kernel(const T* __restrict__ f, const T* __restrict__ g, T* __restrict__ out,
                                                       uint8_t* __restrict__ mask, JointDims d, int relu, float drop_p, uint64_t seed) {
  const int hv = d.H / V;
  const long long total = d.rows * hv;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thresh = (uint32_t)(drop_p * 4294967296.0);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long row = i / hv;
    const int h = (int)(i - row * hv) * V;
    int b, t, u;
    const bool ok = joint_decode(d, row, b, t, u);
    float o[V];
    uint8_t m[V];
    if (ok) {
      float fv[V], gv[V];
      if constexpr (V == 1) {
        fv[0] = to_f<T>(f[((long long)b * d.T + t) * d.H + h]);
        gv[0] = to_f<T>(g[((long long)b * d.U + u) * d.H + h]);
      } else {
        load_vec<T, V>(fv, f + ((long long)b * d.T + t) * d.H + h);
        load_vec<T, V>(gv, g + ((long long)b * d.U + u) * d.H + h);
      }
#pragma unroll
      for (int j = 0; j < V; j++) {
        float v = fv[j] + gv[j];
        bool keep = true;
        if (relu) { keep = v > 0.f; v = fmaxf(v, 0.f); }
        if (drop_p > 0.f) {
          const bool kd = hash_rng(seed, (uint64_t)(row * d.H + h + j)) >= thresh;
          v = kd ? v * keep_scale : 0.f;
          keep = keep && kd;
        }
        o[j] = v; m[j] = keep ? 1 : 0;
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; j++) { o[j] = 0.f; m[j] = 0; }
    }
    if constexpr (V == 1) out[row * d.H + h] = from_f<T>(o[0]);
    else store_vec<T, V>(out + row * d.H + h, o);
    if (mask) {
#pragma unroll
      for (int j = 0; j < V; j++) mask[row * d.H + h + j] = m[j];
    }
  }
}

// CTA x in [0, B*T): df[b,t,:] = sum_u dout[b,t,u,:]*m ; CTA x in [B*T, B*T + B*U): dg[b,u,:] = sum_t dout[b,t,u,:]*m
template <typename T>
__global__ void __launch_bounds__(256) joint_bwd_kernel(const T* __restrict__ dout, const uint8_t* __restrict__ mask, T* __restrict__ df,
                                                       T* __restrict__ dg, JointDims d, float scale) {
  const int BT = d.B * d.T;
  const bool is_f = (int)blockIdx.x < BT;
  const int idx = is_f ? blockIdx.x : blockIdx.x - BT;
  const int b = is_f ? idx / d.T : idx / d.U;
  const int a = is_f ? idx - b * d.T : idx - b * d.U;  // t (for df) or u (for dg)
  const int fl = d.f_len[b], gl = d.g_len[b];
  const int n_red = is_f ? gl : fl;
  const bool live = is_f ? (a < fl) : (a < gl);
  const long long start = d.packed ? (b ? d.batch_offset[b - 1] : 0) : (long long)b * d.T * d.U;
  const int ustride = d.packed ? gl : d.U;
  T* dst = is_f ? df + ((long long)b * d.T + a) * d.H : dg + ((long long)b * d.U + a) * d.H;
  for (int h = threadIdx.x; h < d.H; h += 256) {
    float acc = 0.f;
    if (live) {
#pragma unroll 4
      for (int r = 0; r < n_red; r++) {
        const long long row = start + (is_f ? (long long)a * ustride + r : (long long)r * ustride + a);
        float v = to_f<T>(dout[row * d.H + h]);
        if (mask) v = mask[row * d.H + h] ? v * scale : 0.f;
        acc += v;
      }
    }
    dst[h] = from_f<T>(acc);
  }
}

// ------------------------------------------------------------------------------------------------------------- loss
struct LossDims {
  int B, T, U, V, blank, packed;          // U = max label length + 1
  const int* f_len; const int* y_len; const int* label;  // label [B, U-1]
  const long long* batch_offset;          // packed: inclusive cumsum of f_len*(y_len+1)
};
__device__ __forceinline__ long long cell_row(const LossDims& d, int b, int t, int u) {
  if (!d.packed) return ((long long)b * d.T + t) * d.U + u;
  return (b ? d.batch_offset[b - 1] : 0) + (long long)t * (d.y_len[b] + 1) + u;
}
__device__ __forceinline__ float log_add(float a, float b) {
  const float m = fmaxf(a, b), n = fminf(a, b);
  return (n == -INFINITY) ? m : m + log1pf(__expf(n - m));
}

// one warp per lattice cell: lse over V
template <typename T>
__global__ void __launch_bounds__(256) lse_kernel(const T* __restrict__ x, float* __restrict__ lse, long long rows, int V) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * 256) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const T* p = x + r * V;
    float m = -INFINITY, s = 0.f;
    for (int v = lane; v < V; v += 32) {
      const float xv = to_f<T>(p[v]);
      if (xv > m) { s = s * __expf(m - xv) + 1.f; m = xv; } else { s += __expf(xv - m); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      const float mm = fmaxf(m, m2);
      s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
      m = mm;
    }
    if (lane == 0) lse[r] = m + __logf(s);
  }
}

// grid (B, 2): y = 0 alpha, y = 1 beta. alpha / beta are dense [B, T, U] fp32. Threads run along u, loop over anti-diagonals.
template <typename T>
__global__ void __launch_bounds__(1024) lattice_kernel(const T* __restrict__ x, const float* __restrict__ lse, float* __restrict__ alpha,
                                                      float* __restrict__ beta, float* __restrict__ loss, LossDims d) {
  const int b = blockIdx.x;
  const int Tb = d.f_len[b], Ub = d.y_len[b] + 1;
  const int* lab = d.label + (long long)b * (d.U - 1);
  auto lp_blank = [&](int t, int u) { const long long r = cell_row(d, b, t, u); return to_f<T>(x[r * d.V + d.blank]) - lse[r]; };
  auto lp_emit = [&](int t, int u) { const long long r = cell_row(d, b, t, u); return to_f<T>(x[r * d.V + lab[u]]) - lse[r]; };
  if (Tb <= 0 || Ub <= 0) { if (threadIdx.x == 0 && blockIdx.y == 1) loss[b] = 0.f; return; }
  if (blockIdx.y == 0) {
    float* A = alpha + (long long)b * d.T * d.U;
    for (int dg = 0; dg < Tb + Ub - 1; dg++) {
      for (int u = threadIdx.x; u < Ub; u += blockDim.x) {
        const int t = dg - u;
        if (t < 0 || t >= Tb) continue;
        float v;
        if (t == 0 && u == 0) v = 0.f;
        else {
          const float a = t > 0 ? A[(t - 1) * d.U + u] + lp_blank(t - 1, u) : -INFINITY;
          const float e = u > 0 ? A[t * d.U + u - 1] + lp_emit(t, u - 1) : -INFINITY;
          v = log_add(a, e);
        }
        A[t * d.U + u] = v;
      }
      __syncthreads();
    }
  } else {
    float* Bt = beta + (long long)b * d.T * d.U;
    for (int dg = Tb + Ub - 2; dg >= 0; dg--) {
      for (int u = threadIdx.x; u < Ub; u += blockDim.x) {
        const int t = dg - u;
        if (t < 0 || t >= Tb) continue;
        float v;
        if (t == Tb - 1 && u == Ub - 1) v = lp_blank(t, u);
        else {
          const float a = t < Tb - 1 ? Bt[(t + 1) * d.U + u] + lp_blank(t, u) : -INFINITY;
          const float e = u < Ub - 1 ? Bt[t * d.U + u + 1] + lp_emit(t, u) : -INFINITY;
          v = log_add(a, e);
        }
        Bt[t * d.U + u] = v;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) loss[b] = -Bt[0];
  }
}

// one warp per lattice cell: dx[v] = gl * (softmax_v * exp(alpha+beta-ll) - [v==blank] exp(alpha+lp_blank+beta(t+1,u)-ll)
//                                           - [v==label_u] exp(alpha+lp_emit+beta(t,u+1)-ll))
template <typename T>
__global__ void __launch_bounds__(256) loss_bwd_kernel(const T* __restrict__ x, const float* __restrict__ lse, const float* __restrict__ alpha,
                                                      const float* __restrict__ beta, const float* __restrict__ loss_grad, T* __restrict__ dx,
                                                      LossDims d, long long dense_cells) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * 256) >> 5;
  for (long long c = warp; c < dense_cells; c += nwarps) {
    const long long tu = (long long)d.T * d.U;
    const int b = (int)(c / tu);
    const int r = (int)(c - b * tu);
    const int t = r / d.U, u = r - t * d.U;
    const int Tb = d.f_len[b], Ub = d.y_len[b] + 1;
    const bool ok = t < Tb && u < Ub;
    if (!ok) {
      if (!d.packed) { T* o = dx + c * d.V; for (int v = lane; v < d.V; v += 32) o[v] = from_f<T>(0.f); }
      continue;
    }
    const long long row = cell_row(d, b, t, u);
    const T* p = x + row * d.V;
    T* o = dx + row * d.V;
    const float* A = alpha + (long long)b * tu;
    const float* Bt = beta + (long long)b * tu;
    const float ll = Bt[0], l = lse[row], gl = loss_grad[b];
    const float a = A[t * d.U + u];
    const float common = __expf(a + Bt[t * d.U + u] - ll);
    const float lpb = to_f<T>(p[d.blank]) - l;
    float gb;
    if (t == Tb - 1) gb = (u == Ub - 1) ? __expf(a + lpb - ll) : 0.f;
    else gb = __expf(a + lpb + Bt[(t + 1) * d.U + u] - ll);
    int lbl = -1; float ge = 0.f;
    if (u < Ub - 1) {
      lbl = d.label[(long long)b * (d.U - 1) + u];
      ge = __expf(a + to_f<T>(p[lbl]) - l + Bt[t * d.U + u + 1] - ll);
    }
    for (int v = lane; v < d.V; v += 32) {
      float gv = __expf(to_f<T>(p[v]) - l) * common;
      if (v == d.blank) gv -= gb;
      if (v == lbl) gv -= ge;
      o[v] = from_f<T>(gv * gl);
    }
  }
}

template <typename T>
static int joint_fwd_t(const void* f, const void* g, void* out, uint8_t* mask, const JointDims& d, int relu, float p, uint64_t seed, cudaStream_t s) {
  const bool v = (d.H % (16 / (int)sizeof(T)) == 0) && aligned16(f) && aligned16(g) && aligned16(out);
  const long long work = d.rows * (v ? d.H / (16 / (int)sizeof(T)) : d.H);
  const int grid = (int)((work + 255) / 256 < (long long)kNumSMs * 8 ? (work + 255) / 256 : kNumSMs * 8);
  if (grid <= 0) return 0;
  if (v) joint_fwd_kernel<T, 16 / (int)sizeof(T)><<<grid, 256, 0, s>>>((const T*)f, (const T*)g, (T*)out, mask, d, relu, p, seed);
  else joint_fwd_kernel<T, 1><<<grid, 256, 0, s>>>((const T*)f, (const T*)g, (T*)out, mask, d, relu, p, seed);
  return (int)cudaGetLastError();
}

}  // namespace ab

using namespace ab;

#define DISPATCH_FLOAT(dt, ...)                                                      \
  switch (dt) {                                                                      \
    case kF32: { using T = float; __VA_ARGS__; break; }                              \
    case kF16: { using T = f16; __VA_ARGS__; break; }                                \
    case kBF16: { using T = bf16; __VA_ARGS__; break; }                              \
    default: return -100;                                                            \
  }

AB_API int ab_transducer_joint_fwd(const void* f, const void* g, void* out, void* mask, const int* f_len, const int* g_len,
                                   const long long* batch_offset, int B, int T_, int U, int H, int packed, long long rows, int relu,
                                   float drop_p, unsigned long long seed, int dt, cudaStream_t s) {
  JointDims d{B, T_, U, H, packed, f_len, g_len, batch_offset, rows};
  int rc = 0;
  DISPATCH_FLOAT(dt, rc = joint_fwd_t<T>(f, g, out, (uint8_t*)mask, d, relu, drop_p, seed, s));
  return rc;
}

AB_API int ab_transducer_joint_bwd(const void* dout, const void* mask, void* df, void* dg, const int* f_len, const int* g_len,
                                   const long long* batch_offset, int B, int T_, int U, int H, int packed, float scale, int dt,
                                   cudaStream_t s) {
  JointDims d{B, T_, U, H, packed, f_len, g_len, batch_offset, 0};
  const int grid = B * T_ + B * U;
  if (grid <= 0) return 0;
  DISPATCH_FLOAT(dt, (joint_bwd_kernel<T><<<grid, 256, 0, s>>>((const T*)dout, (const uint8_t*)mask, (T*)df, (T*)dg, d, scale)));
  return (int)cudaGetLastError();
}

AB_API int ab_transducer_loss_fwd(const void* x, float* lse, float* alpha, float* beta, float* loss, const int* label, const int* f_len,
                                  const int* y_len, const long long* batch_offset, int B, int T_, int U, int V, int blank, int packed,
                                  long long rows, int dt, cudaStream_t s) {
  LossDims d{B, T_, U, V, blank, packed, f_len, y_len, label, batch_offset};
  if (rows <= 0 || B <= 0) return 0;
  const long long want = (rows + 7) / 8;
  const int grid = (int)(want < (long long)kNumSMs * 8 ? want : kNumSMs * 8);
  int thr = 32;
  while (thr < U && thr < 1024) thr <<= 1;
  DISPATCH_FLOAT(dt, (lse_kernel<T><<<grid, 256, 0, s>>>((const T*)x, lse, rows, V));
                 (lattice_kernel<T><<<dim3(B, 2), thr, 0, s>>>((const T*)x, lse, alpha, beta, loss, d)));
  return (int)cudaGetLastError();
}

AB_API int ab_transducer_loss_bwd(const void* x, const float* lse, const float* alpha, const float* beta, const float* loss_grad, void* dx,
                                  const int* label, const int* f_len, const int* y_len, const long long* batch_offset, int B, int T_, int U,
                                  int V, int blank, int packed, int dt, cudaStream_t s) {
  LossDims d{B, T_, U, V, blank, packed, f_len, y_len, label, batch_offset};
  const long long cells = (long long)B * T_ * U;
  if (cells <= 0) return 0;
  const long long want = (cells + 7) / 8;
  const int grid = (int)(want < (long long)kNumSMs * 8 ? want : kNumSMs * 8);
  DISPATCH_FLOAT(dt, (loss_bwd_kernel<T><<<grid, 256, 0, s>>>((const T*)x, lse, alpha, beta, loss_grad, (T*)dx, d, cells)));
  return (int)cudaGetLastError();
}
