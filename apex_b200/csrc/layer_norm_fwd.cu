// LayerNorm / RMSNorm forward for sm_100a: the row lives in registers (one global read of x, one write of y), rows are
// spread over a persistent grid sized to the 148 SMs, statistics are a register two-pass (mean, then centred variance).
// Spec: reference csrc/layer_norm_cuda_kernel.cu:317-376,803-835 (cuApplyLayerNorm / cuApplyRMSNorm: outputs y, mean[n1],
// invvar[n1] fp32; gamma/beta have the OUTPUT dtype) and apex/contrib/csrc/layer_norm/ln_fwd_kernels.cuh:6-107.
#include "norm_common.cuh"
#include <cstdlib>
#include <type_traits>

namespace ab {

// 16-byte vector -> fp32. bf16 by hand: (w << 16) and (w & 0xffff0000) are one instruction per element (the library conversion
// of the high half is a PRMT + a shift). OPAQUE routes the words through an opaque move so that a second decode of the same registers
// is really executed instead of the compiler keeping the first pass's fp32 copies alive (register pressure).
template <typename T, bool OPAQUE> __device__ __forceinline__ void decode16(const uint4& raw, float (&f)[16 / sizeof(T)]) {
  uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  if constexpr (OPAQUE) {
#pragma unroll
    for (int q = 0; q < 4; q++) asm volatile("mov.b32 %0, %1;" : "=r"(w[q]) : "r"(w[q]));
  }
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int q = 0; q < 4; q++) f[q] = __uint_as_float(w[q]);
  } else if constexpr (std::is_same<T, bf16>::value) {
#pragma unroll
    for (int q = 0; q < 4; q++) { f[2 * q] = __uint_as_float(w[q] << 16); f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
      f[2 * q] = t.x; f[2 * q + 1] = t.y;
    }
  }
}

// Instruction budget matters as much as bytes here: ncu (profiles/ln_fwd_now.md) counted ~19 issued instructions per element at 58 % of
// DRAM, with the L1 as the busiest unit (80 %) -- every 16-byte vector of x cost four 16-byte shared-memory loads of fp32 gamma / beta.
// So: x is decoded twice (not three times) thanks to single-pass shifted statistics; 16-bit values are decoded by hand (one instruction
// per element, decode16); gamma / beta are staged ONCE per CTA in shared memory as raw 16-bit words (half the L1 traffic of fp32
// copies; mode 1 keeps the fp32 staging for fp32 parameters); the output is (x * rstd - mean * rstd) * gamma + beta, two FMAs.
// Measured effect (bf16, 4096-wide rows): 4.9 -> 5.5 TB/s. Narrow rows (<= 2 vectors per thread) also prefetch the next row group.
template <int MAXV, typename Tin, typename Tout, bool RMS, bool PF = (MAXV <= 2)>
__global__ void __launch_bounds__(512, MAXV <= 4 ? 2 : 1) ln_fwd_vec(const Tin* __restrict__ x, Tout* __restrict__ y, float* __restrict__ mean,
                                                   float* __restrict__ invvar, const Tout* __restrict__ gamma,
                                                   const Tout* __restrict__ beta, int n1, int n2, float eps, int tpr, int gb_smem) {
  constexpr int E = 16 / sizeof(Tin);
  constexpr bool PREFETCH = PF;
  __shared__ float sred[128];
  extern __shared__ float gb_sm[];  // [2][n2] fp32 gamma, beta when gb_smem
  RowReducer red(sred, tpr);
  const int rows_per_cta = blockDim.x / tpr;
  const int nvec = n2 / E;
  const float inv_n = 1.f / (float)n2;
  if (gb_smem == 2 && gamma) {
    const int words = n2 * (int)sizeof(Tout) / 4;
    uint32_t* w = reinterpret_cast<uint32_t*>(gb_sm);
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      w[i] = reinterpret_cast<const uint32_t*>(gamma)[i];
      w[words + i] = beta ? reinterpret_cast<const uint32_t*>(beta)[i] : 0u;
    }
    __syncthreads();
  } else if (gb_smem && gamma) {
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
        float f[E]; decode16<Tin, false>(raw[v], f);
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
      rstd = rsqrt_fast(red.sum(ss) * inv_n + eps);
    } else {
      red.sum2(s, ss);
      const float md = s * inv_n;
      mu = shift + md;
      rstd = rsqrt_fast(fmaxf(ss * inv_n - md * md, 0.f) + eps);
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
          float f[E]; decode16<Tin, false>(raw[v], f);
          float o[E];
          if (gamma) {
            float g[E], bb[E];
            if (gb_smem == 2) {   // raw Tout bits in shared memory: half the bytes of the fp32 staging (the L1 was the busiest unit: 80 %)
              constexpr int GW = E * sizeof(Tout) / 4;
              const uint32_t* gp = reinterpret_cast<const uint32_t*>(gb_sm) + (size_t)idx * GW;
              const uint32_t* bp = gp + (size_t)nvec * GW;
              uint32_t gw[GW], bw[GW];
              if constexpr (GW % 4 == 0) {
#pragma unroll
                for (int q = 0; q < GW / 4; q++) {
                  const uint4 a4 = reinterpret_cast<const uint4*>(gp)[q], b4 = reinterpret_cast<const uint4*>(bp)[q];
                  gw[4 * q] = a4.x; gw[4 * q + 1] = a4.y; gw[4 * q + 2] = a4.z; gw[4 * q + 3] = a4.w;
                  bw[4 * q] = b4.x; bw[4 * q + 1] = b4.y; bw[4 * q + 2] = b4.z; bw[4 * q + 3] = b4.w;
                }
              } else {
                const uint2 a2 = *reinterpret_cast<const uint2*>(gp), b2 = *reinterpret_cast<const uint2*>(bp);
                gw[0] = a2.x; gw[1] = a2.y; bw[0] = b2.x; bw[1] = b2.y;
              }
              if constexpr (sizeof(Tout) == 2 && GW == 4) {
                decode16<Tout, false>(make_uint4(gw[0], gw[1], gw[2], gw[3]), reinterpret_cast<float(&)[16 / sizeof(Tout)]>(g));
                decode16<Tout, false>(make_uint4(bw[0], bw[1], bw[2], bw[3]), reinterpret_cast<float(&)[16 / sizeof(Tout)]>(bb));
              } else {
                const Tout* ge = reinterpret_cast<const Tout*>(gw);
                const Tout* be = reinterpret_cast<const Tout*>(bw);
#pragma unroll
                for (int e = 0; e < E; e++) { g[e] = to_f<Tout>(ge[e]); bb[e] = to_f<Tout>(be[e]); }
              }
            } else if (gb_smem) {
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
            for (int e = 0; e < E; e++) o[e] = fmaf(fmaf(f[e], rstd, nmr), g[e], bb[e]);   // xhat * gamma + beta: two FMAs
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
  static const int pf4 = getenv("APEX_B200_LN_FWD_PF4") ? atoi(getenv("APEX_B200_LN_FWD_PF4")) : 0;    // prefetch the next row group at 4 vectors / thread
  static const int gb_kb = getenv("APEX_B200_LN_GB_SMEM_KB") ? atoi(getenv("APEX_B200_LN_GB_SMEM_KB")) : 32;
  NormCfg c = norm_cfg(vec_ok ? n2 / E : 1, target_v, 512);
  if (vec_ok && c.ok) {
    int grid = (n1 + c.rows_per_cta - 1) / c.rows_per_cta;
    // fp32 gamma / beta staged in shared memory while 4 CTAs per SM still fit comfortably (<= 32 KB each: hidden <= 4096), else read through L1
    // gamma / beta staging in shared memory: mode 2 = raw Tout bits (any width up to 64 KB), mode 1 = fp32 copies (<= gb_kb KB), 0 = global
    static const int gb_mode = getenv("APEX_B200_LN_GB_MODE") ? atoi(getenv("APEX_B200_LN_GB_MODE")) : 2;
    const size_t raw_bytes = (size_t)2 * n2 * sizeof(Tout);
    const bool raw_ok = gb_mode == 2 && gamma && sizeof(Tout) < 4 && raw_bytes <= 64 * 1024;
    const size_t dyn = raw_ok ? raw_bytes : ((gamma && (size_t)2 * n2 * sizeof(float) <= (size_t)gb_kb * 1024) ? (size_t)2 * n2 * sizeof(float) : 0);
    const int gb_flag = raw_ok ? 2 : (dyn ? 1 : 0);
    const int cap = kNumSMs * (1024 / c.threads);
    if (grid > cap) grid = cap;
#define LN_FWD_GO2(MV, PFV)                                                                                           \
  if (dyn > 40 * 1024) cudaFuncSetAttribute(ln_fwd_vec<MV, Tin, Tout, RMS, PFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn); \
  ln_fwd_vec<MV, Tin, Tout, RMS, PFV><<<grid, c.threads, dyn, st>>>((const Tin*)x, (Tout*)y, mean, invvar, (const Tout*)gamma, \
                                                                  (const Tout*)beta, n1, n2, eps, c.tpr, gb_flag)
#define LN_FWD_GO(MV) LN_FWD_GO2(MV, (MV <= 2))
    switch (c.maxv) {
      case 1: LN_FWD_GO(1); break;
      case 2: LN_FWD_GO(2); break;
      case 4: if (pf4) { LN_FWD_GO2(4, true); } else { LN_FWD_GO2(4, false); } break;
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
