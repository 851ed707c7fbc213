// LayerNorm / RMSNorm forward for sm_100a: the row lives in registers (one global read of x, one write of y), rows are
// spread over a persistent grid sized to the 148 SMs, statistics are a register two-pass (mean, then centred variance).
// Spec: reference csrc/layer_norm_cuda_kernel.cu:317-376,803-835 (cuApplyLayerNorm / cuApplyRMSNorm: outputs y, mean[n1],
// invvar[n1] fp32; gamma/beta have the OUTPUT dtype) and apex/contrib/csrc/layer_norm/ln_fwd_kernels.cuh:6-107.
#include "norm_common.cuh"
#include <cstdlib>

namespace ab {

// Instruction budget matters as much as bytes here (bf16: 4 B/element of traffic buys ~20 issue slots per element at the HBM
// roofline): x is decoded twice (not three times) thanks to single-pass shifted statistics, gamma / beta are converted to fp32
// ONCE per CTA into shared memory (16-byte LDS instead of a global load + convert per element and row), and the output is
// y = x*A + B with A = rstd*gamma, B = beta - mean*A. Narrow rows (<= 2 vectors per thread) also prefetch the next row group.
template <int MAXV, typename Tin, typename Tout, bool RMS>
__global__ void __launch_bounds__(512, MAXV <= 4 ? 2 : 1) ln_fwd_vec(const Tin* __restrict__ x, Tout* __restrict__ y, float* __restrict__ mean,
                                                   float* __restrict__ invvar, const Tout* __restrict__ gamma,
                                                   const Tout* __restrict__ beta, int n1, int n2, float eps, int tpr, int gb_smem) {
  constexpr int E = 16 / sizeof(Tin);
  constexpr bool PREFETCH = MAXV <= 2;
  __shared__ float sred[128];
  extern __shared__ float gb_sm[];  // [2][n2] fp32 gamma, beta when gb_smem
  RowReducer red(sred, tpr);
  const int rows_per_cta = blockDim.x / tpr;
  const int nvec = n2 / E;
  const float inv_n = 1.f / (float)n2;
  if (gb_smem && gamma) {
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      gb_sm[i] = to_f<Tout>(gamma[i]);
      gb_sm[n2 + i] = beta ? to_f<Tout>(beta[i]) : 0.f;
    }
    __syncthreads();
  }
  auto load_rows = [&](int row0, uint4 (&raw)[MAXV], float& shift) {
    const int row = row0 + red.rg;
    const bool valid = row < n1;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * n2);
#pragma unroll
    for (int v = 0; v < MAXV; v++) {
      const int idx = v * tpr + red.lane_r;
      raw[v] = (valid && idx < nvec) ? __ldg(xr + idx) : make_uint4(0, 0, 0, 0);
    }
    shift = (!RMS && valid) ? to_f<Tin>(__ldg(x + (size_t)row * n2)) : 0.f;  // a sample of the row: kills the cancellation in E[d^2]-E[d]^2
  };
  const int row_step = gridDim.x * rows_per_cta;
  uint4 raw[MAXV], raw_n[PREFETCH ? MAXV : 1];
  float shift = 0.f, shift_n = 0.f;
  if (PREFETCH && blockIdx.x * rows_per_cta < n1) load_rows(blockIdx.x * rows_per_cta, raw, shift);
  for (int row0 = blockIdx.x * rows_per_cta; row0 < n1; row0 += row_step) {
    const int row = row0 + red.rg;
    const bool valid = row < n1;
    const bool more = PREFETCH && row0 + row_step < n1;
    if (!PREFETCH) load_rows(row0, raw, shift);
    if constexpr (PREFETCH) { if (more) load_rows(row0 + row_step, raw_n, shift_n); }
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; v++) {
      const int idx = v * tpr + red.lane_r;
      if (idx < nvec) {
        float f[E]; unpack16<Tin>(raw[v], f);
#pragma unroll
        for (int e = 0; e < E; e++) {
          const float d = RMS ? f[e] : f[e] - shift;
          if (!RMS) s += d;
          ss = fmaf(d, d, ss);
        }
      }
    }
    float mu = 0.f, rstd;
    if (RMS) {
      rstd = rsqrtf(red.sum(ss) * inv_n + eps);
    } else {
      red.sum2(s, ss);
      const float md = s * inv_n;
      mu = shift + md;
      rstd = rsqrtf(fmaxf(ss * inv_n - md * md, 0.f) + eps);
    }
    if (valid) {
      if (red.lane_r == 0) {
        if (!RMS && mean) mean[row] = mu;
        invvar[row] = rstd;
      }
      Tout* yr = y + (size_t)row * n2;
      const float nmr = -mu * rstd;
#pragma unroll
      for (int v = 0; v < MAXV; v++) {
        const int idx = v * tpr + red.lane_r;
        if (idx < nvec) {
          float f[E]; unpack16<Tin>(raw[v], f);
          float o[E];
          if (gamma) {
            float g[E], bb[E];
            if (gb_smem) {
#pragma unroll
              for (int q = 0; q < E / 4; q++) {
                const float4 g4 = reinterpret_cast<const float4*>(gb_sm + (size_t)idx * E)[q];
                const float4 b4 = reinterpret_cast<const float4*>(gb_sm + n2 + (size_t)idx * E)[q];
                g[4 * q] = g4.x; g[4 * q + 1] = g4.y; g[4 * q + 2] = g4.z; g[4 * q + 3] = g4.w;
                bb[4 * q] = b4.x; bb[4 * q + 1] = b4.y; bb[4 * q + 2] = b4.z; bb[4 * q + 3] = b4.w;
              }
            } else {
              load_vec<Tout, E>(g, gamma + (size_t)idx * E);
              if (beta) load_vec<Tout, E>(bb, beta + (size_t)idx * E);
              else {
#pragma unroll
                for (int e = 0; e < E; e++) bb[e] = 0.f;
              }
            }
#pragma unroll
            for (int e = 0; e < E; e++) {
              const float A = rstd * g[e];
              o[e] = fmaf(f[e], A, fmaf(nmr, g[e], bb[e]));
            }
          } else {
#pragma unroll
            for (int e = 0; e < E; e++) o[e] = fmaf(f[e], rstd, nmr);
          }
          store_vec<Tout, E>(yr + (size_t)idx * E, o);
        }
      }
    }
    if constexpr (PREFETCH) {
      if (more) {
#pragma unroll
        for (int v = 0; v < MAXV; v++) raw[v] = raw_n[v];
        shift = shift_n;
      }
    }
  }
}

// Any n2 / any alignment: one CTA per row, strided scalar access, x re-read from L2 for the second pass.
template <typename Tin, typename Tout, bool RMS>
__global__ void __launch_bounds__(256) ln_fwd_generic(const Tin* __restrict__ x, Tout* __restrict__ y, float* __restrict__ mean,
                                                      float* __restrict__ invvar, const Tout* __restrict__ gamma,
                                                      const Tout* __restrict__ beta, int n1, int n2, float eps) {
  __shared__ float red[40];
  for (int row = blockIdx.x; row < n1; row += gridDim.x) {
    const Tin* xr = x + (size_t)row * n2;
    float mu = 0.f;
    if (!RMS) {
      float s = 0.f;
      for (int i = threadIdx.x; i < n2; i += blockDim.x) s += to_f<Tin>(xr[i]);
      mu = block_sum(s, red) / (float)n2;
    }
    float ss = 0.f;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) { const float d = to_f<Tin>(xr[i]) - mu; ss += d * d; }
    const float rstd = rsqrtf(block_sum(ss, red) / (float)n2 + eps);
    if (threadIdx.x == 0) { if (!RMS && mean) mean[row] = mu; invvar[row] = rstd; }
    Tout* yr = y + (size_t)row * n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      float o = (to_f<Tin>(xr[i]) - mu) * rstd;
      if (gamma) o *= to_f<Tout>(gamma[i]);
      if (beta) o += to_f<Tout>(beta[i]);
      yr[i] = from_f<Tout>(o);
    }
  }
}

template <typename Tin, typename Tout, bool RMS>
int ln_fwd_launch(const void* x, void* y, float* mean, float* invvar, const void* gamma, const void* beta, int n1, int n2,
                  float eps, cudaStream_t st) {
  constexpr int E = 16 / sizeof(Tin);
  const bool vec_ok = (n2 % E == 0) && aligned16(x) && ((size_t)n2 * sizeof(Tin)) % 16 == 0 && aligned16(y) &&
                      ((size_t)n2 * sizeof(Tout)) % 16 == 0 && (!gamma || aligned16(gamma)) && (!beta || aligned16(beta));
  static const int target_v = getenv("APEX_B200_LN_FWD_V") ? atoi(getenv("APEX_B200_LN_FWD_V")) : 4;  // tuning knob (vectors / thread)
  NormCfg c = norm_cfg(vec_ok ? n2 / E : 1, target_v, 512);
  if (vec_ok && c.ok) {
    int grid = (n1 + c.rows_per_cta - 1) / c.rows_per_cta;
    // fp32 gamma / beta staged in shared memory while 4 CTAs per SM still fit comfortably (<= 32 KB each: hidden <= 4096), else read through L1
    const size_t dyn = (gamma && (size_t)2 * n2 * sizeof(float) <= 32 * 1024) ? (size_t)2 * n2 * sizeof(float) : 0;
    const int cap = kNumSMs * (1024 / c.threads);
    if (grid > cap) grid = cap;
#define LN_FWD_GO(MV)                                                                                                 \
  if (dyn > 48 * 1024) cudaFuncSetAttribute(ln_fwd_vec<MV, Tin, Tout, RMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn); \
  ln_fwd_vec<MV, Tin, Tout, RMS><<<grid, c.threads, dyn, st>>>((const Tin*)x, (Tout*)y, mean, invvar, (const Tout*)gamma, \
                                                             (const Tout*)beta, n1, n2, eps, c.tpr, dyn ? 1 : 0)
    switch (c.maxv) {
      case 1: LN_FWD_GO(1); break;
      case 2: LN_FWD_GO(2); break;
      case 4: LN_FWD_GO(4); break;
      default: LN_FWD_GO(8); break;
    }
  } else {
    int grid = n1 < kNumSMs * 8 ? n1 : kNumSMs * 8;
    ln_fwd_generic<Tin, Tout, RMS><<<grid, 256, 0, st>>>((const Tin*)x, (Tout*)y, mean, invvar, (const Tout*)gamma,
                                                        (const Tout*)beta, n1, n2, eps);
  }
  AB_CHECK_LAUNCH();
  return 0;
}

}  // namespace ab

using namespace ab;

// x[n1,n2] (dt_in) -> y[n1,n2] (dt_out); gamma/beta (dt_out) may be null; mean may be null for RMSNorm.
AB_API int ab_layer_norm_fwd(const void* x, void* y, float* mean, float* invvar, const void* gamma, const void* beta, int n1,
                             int n2, float eps, int dt_in, int dt_out, int rms, cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0) return 0;
#define LN_PAIR(TI, TO)                                                                     \
  return rms ? ln_fwd_launch<TI, TO, true>(x, y, mean, invvar, gamma, beta, n1, n2, eps, st) \
             : ln_fwd_launch<TI, TO, false>(x, y, mean, invvar, gamma, beta, n1, n2, eps, st)
  if (dt_in == kF32 && dt_out == kF32) { LN_PAIR(float, float); }
  if (dt_in == kF16 && dt_out == kF16) { LN_PAIR(f16, f16); }
  if (dt_in == kBF16 && dt_out == kBF16) { LN_PAIR(bf16, bf16); }
  if (dt_in == kF16 && dt_out == kF32) { LN_PAIR(f16, float); }
  if (dt_in == kBF16 && dt_out == kF32) { LN_PAIR(bf16, float); }
  if (dt_in == kF32 && dt_out == kF16) { LN_PAIR(float, f16); }
  if (dt_in == kF32 && dt_out == kBF16) { LN_PAIR(float, bf16); }
  return -1;
}
