// LayerNorm / RMSNorm backward for sm_100a. ONE pass over (dy, x): each persistent CTA keeps its rows in registers,
// produces dx, and accumulates its share of dgamma/dbeta in registers; a tiny second kernel folds the <=296 per-CTA partials.
// (The reference reads dy and x twice: cuComputePartGradGammaBeta + cuComputeGradInput, layer_norm_cuda_kernel.cu:481-801.)
// memory_efficient: the saved tensor is the OUTPUT y; xhat is rebuilt as (y-beta)/clamp(gamma) (reference :378-394,416,761).
#include "norm_common.cuh"
#include <cstdlib>

namespace ab {

__device__ __forceinline__ float clamp_mag(float g, float eps) {
  return fabsf(g) < eps ? copysignf(eps, g) : g;
}

// Tin: dtype of x and dx. Tout: dtype of dy, gamma, beta (and of y when MEMEFF).
template <int MAXV, typename Tin, typename Tout, bool RMS, bool MEMEFF>
__global__ void __launch_bounds__(512) ln_bwd_vec(const Tout* __restrict__ dy, const void* __restrict__ saved, const float* __restrict__ mean,
                           const float* __restrict__ invvar, const Tout* __restrict__ gamma, const Tout* __restrict__ beta,
                           Tin* __restrict__ dx, float* __restrict__ part_g, float* __restrict__ part_b, int n1, int n_full,
                           float eps, int tpr, int cl) {
  // cl == 2: a CLUSTER of two CTAs shares every row -- each owns one half of the columns, so rows of up to 2 x 1024 vectors keep the
  // register-accumulator / prefetching configuration (MAXV <= 2) instead of shared-memory accumulators; the two halves of a row's
  // (s1, s2) meet through distributed shared memory, one cluster barrier per row group. n2 below is the width THIS CTA works on.
  constexpr int E = 16 / sizeof(Tin);
  __shared__ float sred[128];
  __shared__ float sacc[2][4096];
  __shared__ float xchg[2][2][16];   // [parity][s1 / s2][row group]: the PEER's half-row sums, pushed through distributed shared memory
  RowReducer red(sred, tpr);
  const int rows_per_cta = blockDim.x / tpr;
  const int n2 = n_full / cl;
  uint32_t crank = 0;
  if (cl == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cid = blockIdx.x / cl, ncl = gridDim.x / cl;   // row distribution is per cluster
  {
    const size_t col0 = (size_t)crank * n2;
    dy += col0; dx += col0;
    saved = MEMEFF ? (const void*)(reinterpret_cast<const Tout*>(saved) + col0) : (const void*)(reinterpret_cast<const Tin*>(saved) + col0);
    if (gamma) gamma += col0;
    if (beta) beta += col0;
    if (part_g) part_g += col0;
    if (part_b) part_b += col0;
  }
  int xpar = 0;
  // a CTA may store into its peer's shared memory only once the peer is running (compute-sanitizer racecheck: "located in a block that
  // might not have entered yet"): one cluster barrier before the first exchange
  if (cl == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int nvec = n2 / E;
  const float inv_n = 1.f / (float)n_full;
  // per-thread dgamma / dbeta accumulators: registers, or (rows wider than 2 vectors per thread) this thread's private columns
  // of a dynamic shared-memory array -- 64 accumulator registers would push the row itself out to local memory
  constexpr bool SACC = MAXV >= 4;
  extern __shared__ float dyn_acc[];  // [2][n2] when SACC
  float acc_g[SACC ? 1 : MAXV][E], acc_b[SACC ? 1 : MAXV][E];
  if (SACC) {
    for (int i = threadIdx.x; i < 2 * n2; i += blockDim.x) dyn_acc[i] = 0.f;
    __syncthreads();
  } else {
#pragma unroll
    for (int v = 0; v < (SACC ? 1 : MAXV); v++)
#pragma unroll
      for (int e = 0; e < E; e++) { acc_g[v][e] = 0.f; acc_b[v][e] = 0.f; }
  }

  constexpr int DW = E * sizeof(Tout) / 4;  // 32-bit words of one dy / y vector
  constexpr int SW = MEMEFF ? DW : 4;
  // rows are kept as RAW bits (half the registers of fp32 copies) and decoded twice; the NEXT row group's loads are issued
  // before this one's reductions (register double buffering) so that every SM keeps ~2x the bytes in flight
  auto load_rows = [&](int row0, uint32_t (&draw)[MAXV][DW], uint32_t (&sraw)[MAXV][SW], float& mu, float& rstd) {
    const int row = row0 + red.rg;
    const bool valid = row < n1;
    mu = (valid && !RMS && !MEMEFF) ? mean[row] : 0.f;
    rstd = valid ? invvar[row] : 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; v++) {
      const int idx = v * tpr + red.lane_r;
      const bool on = valid && idx < nvec;
      const uint32_t* dp = reinterpret_cast<const uint32_t*>(dy + (size_t)row * n_full + (size_t)idx * E);
      if (DW >= 4) {
#pragma unroll
        for (int q = 0; q < DW / 4; q++) {
          uint4 t = on ? __ldg(reinterpret_cast<const uint4*>(dp) + q) : make_uint4(0, 0, 0, 0);
          draw[v][q * 4 + 0] = t.x; draw[v][q * 4 + 1] = t.y; draw[v][q * 4 + 2] = t.z; draw[v][q * 4 + 3] = t.w;
        }
      } else {
        uint2 t = on ? __ldg(reinterpret_cast<const uint2*>(dp)) : make_uint2(0, 0);
        draw[v][0] = t.x; draw[v][1] = t.y;
      }
      if (MEMEFF) {
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(reinterpret_cast<const Tout*>(saved) + (size_t)row * n_full + (size_t)idx * E);
        if (DW >= 4) {
#pragma unroll
          for (int q = 0; q < DW / 4; q++) {
            uint4 t = on ? __ldg(reinterpret_cast<const uint4*>(sp) + q) : make_uint4(0, 0, 0, 0);
            sraw[v][q * 4 + 0] = t.x; sraw[v][q * 4 + 1] = t.y; sraw[v][q * 4 + 2] = t.z; sraw[v][q * 4 + 3] = t.w;
          }
        } else {
          uint2 t = on ? __ldg(reinterpret_cast<const uint2*>(sp)) : make_uint2(0, 0);
          sraw[v][0] = t.x; sraw[v][1] = t.y;
        }
      } else {
        uint4 t = on ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const Tin*>(saved) + (size_t)row * n_full + (size_t)idx * E))
                     : make_uint4(0, 0, 0, 0);
        sraw[v][0] = t.x; sraw[v][1] = t.y; sraw[v][2] = t.z; sraw[v][3] = t.w;
      }
    }
  };
  // gamma (and beta for the memory-efficient variant) are loop-invariant per thread: packed raw bits in registers
  uint32_t graw[MAXV][DW], braw[MAXV][MEMEFF ? DW : 1];
#pragma unroll
  for (int v = 0; v < MAXV; v++) {
    const int idx = v * tpr + red.lane_r;
#pragma unroll
    for (int q = 0; q < DW; q++) {
      graw[v][q] = (gamma && idx < nvec) ? reinterpret_cast<const uint32_t*>(gamma + (size_t)idx * E)[q] : 0u;
      if (MEMEFF) braw[v][q] = (beta && idx < nvec) ? reinterpret_cast<const uint32_t*>(beta + (size_t)idx * E)[q] : 0u;
    }
  }
  const int row_step = ncl * rows_per_cta;
  uint32_t draw[MAXV][DW], sraw[MAXV][SW], draw_n[MAXV][DW], sraw_n[MAXV][SW];
  float mu = 0.f, rstd = 0.f, mu_n = 0.f, rstd_n = 0.f;
  constexpr bool PREFETCH = MAXV <= 2;  // wider rows would spill: they keep one row group in flight
  if (PREFETCH && cid * rows_per_cta < n1) load_rows(cid * rows_per_cta, draw, sraw, mu, rstd);
  for (int row0 = cid * rows_per_cta; row0 < n1; row0 += row_step) {
    const int row = row0 + red.rg;
    const bool valid = row < n1;
    const bool more = PREFETCH && row0 + row_step < n1;
    if (!PREFETCH) load_rows(row0, draw, sraw, mu, rstd);
    if (more) load_rows(row0 + row_step, draw_n, sraw_n, mu_n, rstd_n);
    const float nmr = -mu * rstd;
    // decode vector v -> dy and xhat (and the raw x for the non-memory-efficient pass 2)
    auto decode = [&](int v, float (&xh)[E], float (&d)[E], float (&xr)[E]) {
      decode_words<Tout, E>(draw[v], d);
      if (MEMEFF) {
        float yv[E], gv[E], bv[E];
        decode_words<Tout, E>(sraw[v], yv);
        decode_words<Tout, E>(graw[v], gv);
        decode_words<Tout, E>(braw[v], bv);
#pragma unroll
        for (int e = 0; e < E; e++) {
          const float yy = (!RMS && beta) ? yv[e] - bv[e] : yv[e];
          xh[e] = gamma ? yy / clamp_mag(gv[e], eps) : yy;
          xr[e] = xh[e];
        }
      } else {
        decode_words<Tin, E>(sraw[v], xr);
#pragma unroll
        for (int e = 0; e < E; e++) xh[e] = fmaf(xr[e], rstd, nmr);
      }
    };
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; v++) {
      const int idx = v * tpr + red.lane_r;
      if (valid && idx < nvec) {
        float xh[E], d[E], xr[E], gv[E];
        decode(v, xh, d, xr);
        decode_words<Tout, E>(graw[v], gv);
        if (SACC) {
          float* ag = dyn_acc + (size_t)idx * E;
          float* ab = ag + n2;
#pragma unroll
          for (int q = 0; q < E / 4; q++) {
            float4 g4 = reinterpret_cast<float4*>(ag)[q], b4 = reinterpret_cast<float4*>(ab)[q];
            g4.x = fmaf(d[4 * q], xh[4 * q], g4.x); g4.y = fmaf(d[4 * q + 1], xh[4 * q + 1], g4.y);
            g4.z = fmaf(d[4 * q + 2], xh[4 * q + 2], g4.z); g4.w = fmaf(d[4 * q + 3], xh[4 * q + 3], g4.w);
            b4.x += d[4 * q]; b4.y += d[4 * q + 1]; b4.z += d[4 * q + 2]; b4.w += d[4 * q + 3];
            reinterpret_cast<float4*>(ag)[q] = g4; reinterpret_cast<float4*>(ab)[q] = b4;
          }
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
          if (!SACC) { acc_g[v][e] = fmaf(d[e], xh[e], acc_g[v][e]); acc_b[v][e] += d[e]; }
          const float w = gamma ? d[e] * gv[e] : d[e];
          if (!RMS) s1 += w;
          s2 = fmaf(w, xh[e], s2);
        }
      }
    }
    if (!RMS) red.sum2(s1, s2); else s2 = red.sum(s2);
    if (cl == 2) {   // add the other half of the row: each CTA PUSHES its sums into the peer's slot, so after the barrier the read is local
      if (red.lane_r == 0) {
        uint32_t a1, a2;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a1) : "r"((uint32_t)__cvta_generic_to_shared(&xchg[xpar][0][red.rg])), "r"(crank ^ 1u));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a2) : "r"((uint32_t)__cvta_generic_to_shared(&xchg[xpar][1][red.rg])), "r"(crank ^ 1u));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a1), "f"(s1) : "memory");
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a2), "f"(s2) : "memory");
      }
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      const float p1 = xchg[xpar][0][red.rg], p2 = xchg[xpar][1][red.rg];
      s1 += p1; s2 += p2;
      xpar ^= 1;
    }
    if (valid) {
      // dx = rstd*(w - s1/n - xhat*s2/n) = rstd*w + A + B*xhat;  with xhat = x*rstd + nmr:  = rstd*w + x*(B*rstd) + (A + B*nmr)
      const float A = RMS ? 0.f : -rstd * s1 * inv_n, B = -rstd * s2 * inv_n;
      const float Br = MEMEFF ? B : B * rstd, C = MEMEFF ? A : fmaf(B, nmr, A);
#pragma unroll
      for (int v = 0; v < MAXV; v++) {
        const int idx = v * tpr + red.lane_r;
        if (idx < nvec) {
          float xh[E], d[E], xr[E], o[E], gv[E];
          if (MEMEFF) {
            decode(v, xh, d, xr);
          } else {
            decode_words<Tout, E>(draw[v], d);
            decode_words<Tin, E>(sraw[v], xr);
          }
          decode_words<Tout, E>(graw[v], gv);
#pragma unroll
          for (int e = 0; e < E; e++) {
            const float w = gamma ? d[e] * gv[e] : d[e];
            o[e] = fmaf(xr[e], Br, fmaf(rstd, w, C));
          }
          store_vec<Tin, E>(dx + (size_t)row * n_full + (size_t)idx * E, o);
        }
      }
    }
    if (more) {
#pragma unroll
      for (int v = 0; v < MAXV; v++) {
#pragma unroll
        for (int q = 0; q < DW; q++) draw[v][q] = draw_n[v][q];
#pragma unroll
        for (int q = 0; q < SW; q++) sraw[v][q] = sraw_n[v][q];
      }
      mu = mu_n; rstd = rstd_n;
    }
  }

  if (cl == 2)   // nobody leaves while the peer may still read this CTA's exchange slots
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (part_g == nullptr) return;  // no affine parameters
  // fold the row groups of this CTA, then one partial row per CTA
  if (SACC) {
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      part_g[(size_t)cid * n_full + i] = dyn_acc[i];
      if (part_b) part_b[(size_t)cid * n_full + i] = dyn_acc[n2 + i];
    }
  } else if (rows_per_cta > 1) {
    for (int i = threadIdx.x; i < n2; i += blockDim.x) { sacc[0][i] = 0.f; sacc[1][i] = 0.f; }
    __syncthreads();
    for (int g = 0; g < rows_per_cta; g++) {
      if (red.rg == g) {
#pragma unroll
        for (int v = 0; v < MAXV; v++) {
          const int idx = v * tpr + red.lane_r;
          if (idx < nvec) {
#pragma unroll
            for (int e = 0; e < E; e++) { sacc[0][idx * E + e] += acc_g[SACC ? 0 : v][e]; sacc[1][idx * E + e] += acc_b[SACC ? 0 : v][e]; }
          }
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      part_g[(size_t)cid * n_full + i] = sacc[0][i];
      if (part_b) part_b[(size_t)cid * n_full + i] = sacc[1][i];
    }
  } else {
#pragma unroll
    for (int v = 0; v < MAXV; v++) {
      const int idx = v * tpr + red.lane_r;
      if (idx < nvec) {
        store_vec<float, E>(part_g + (size_t)cid * n_full + (size_t)idx * E, acc_g[SACC ? 0 : v]);
        if (part_b) store_vec<float, E>(part_b + (size_t)cid * n_full + (size_t)idx * E, acc_b[SACC ? 0 : v]);
      }
    }
  }
}

// dgamma[c] = sum_k part_g[k][c] (fixed order => deterministic)
template <typename Tout>
__global__ void __launch_bounds__(512) ln_bwd_fold(const float* __restrict__ part_g, const float* __restrict__ part_b, int k, int n2,
                                                   Tout* __restrict__ dgamma, Tout* __restrict__ dbeta) {
  __shared__ float sg[16][33], sb[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float g = 0.f, b = 0.f;
  if (c < n2) {
    for (int r = threadIdx.y; r < k; r += 16) {
      g += part_g[(size_t)r * n2 + c];
      if (part_b) b += part_b[(size_t)r * n2 + c];
    }
  }
  sg[threadIdx.y][threadIdx.x] = g; sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < n2) {
#pragma unroll
    for (int r = 1; r < 16; r++) { g += sg[r][threadIdx.x]; b += sb[r][threadIdx.x]; }
    dgamma[c] = from_f<Tout>(g);
    if (dbeta) dbeta[c] = from_f<Tout>(b);
  }
}

// ---- generic fallback (any n2 / alignment): dx one CTA per row; dgamma/dbeta by column strips re-reading dy and x.
template <typename Tin, typename Tout, bool RMS, bool MEMEFF>
__global__ void __launch_bounds__(256) ln_bwd_dx_generic(const Tout* __restrict__ dy, const void* __restrict__ saved,
                                                         const float* __restrict__ mean, const float* __restrict__ invvar,
                                                         const Tout* __restrict__ gamma, const Tout* __restrict__ beta,
                                                         Tin* __restrict__ dx, int n1, int n2, float eps) {
  __shared__ float red[40];
  for (int row = blockIdx.x; row < n1; row += gridDim.x) {
    const float mu = (!RMS && !MEMEFF) ? mean[row] : 0.f, rstd = invvar[row];
    auto xhat = [&](int i) -> float {
      if (MEMEFF) {
        float yv = to_f<Tout>(reinterpret_cast<const Tout*>(saved)[(size_t)row * n2 + i]);
        if (!RMS && beta) yv -= to_f<Tout>(beta[i]);
        return gamma ? yv / clamp_mag(to_f<Tout>(gamma[i]), eps) : yv;
      }
      return (to_f<Tin>(reinterpret_cast<const Tin*>(saved)[(size_t)row * n2 + i]) - mu) * rstd;
    };
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const float w = to_f<Tout>(dy[(size_t)row * n2 + i]) * (gamma ? to_f<Tout>(gamma[i]) : 1.f);
      s1 += w; s2 += w * xhat(i);
    }
    s1 = RMS ? 0.f : block_sum(s1, red) / (float)n2;
    s2 = block_sum(s2, red) / (float)n2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      const float w = to_f<Tout>(dy[(size_t)row * n2 + i]) * (gamma ? to_f<Tout>(gamma[i]) : 1.f);
      dx[(size_t)row * n2 + i] = from_f<Tin>(rstd * (w - s1 - xhat(i) * s2));
    }
  }
}

template <typename Tin, typename Tout, bool RMS, bool MEMEFF>
__global__ void __launch_bounds__(512) ln_bwd_gamma_generic(const Tout* __restrict__ dy, const void* __restrict__ saved,
                                                            const float* __restrict__ mean, const float* __restrict__ invvar,
                                                            const Tout* __restrict__ gamma, const Tout* __restrict__ beta,
                                                            Tout* __restrict__ dgamma, Tout* __restrict__ dbeta, int n1, int n2,
                                                            float eps) {
  __shared__ float sg[16][33], sb[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float g = 0.f, b = 0.f;
  if (c < n2) {
    for (int r = threadIdx.y; r < n1; r += 16) {
      const float d = to_f<Tout>(dy[(size_t)r * n2 + c]);
      float xh;
      if (MEMEFF) {
        float yv = to_f<Tout>(reinterpret_cast<const Tout*>(saved)[(size_t)r * n2 + c]);
        if (!RMS && beta) yv -= to_f<Tout>(beta[c]);
        xh = gamma ? yv / clamp_mag(to_f<Tout>(gamma[c]), eps) : yv;
      } else {
        const float mu = RMS ? 0.f : mean[r];
        xh = (to_f<Tin>(reinterpret_cast<const Tin*>(saved)[(size_t)r * n2 + c]) - mu) * invvar[r];
      }
      g += d * xh; b += d;
    }
  }
  sg[threadIdx.y][threadIdx.x] = g; sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < n2) {
#pragma unroll
    for (int r = 1; r < 16; r++) { g += sg[r][threadIdx.x]; b += sb[r][threadIdx.x]; }
    dgamma[c] = from_f<Tout>(g);
    if (dbeta) dbeta[c] = from_f<Tout>(b);
  }
}

template <typename Tin, typename Tout, bool RMS, bool MEMEFF>
int ln_bwd_launch(const void* dy, const void* saved, const float* mean, const float* invvar, const void* gamma, const void* beta,
                  void* dx, void* dgamma, void* dbeta, float* ws, int n1, int n2, float eps, cudaStream_t st) {
  constexpr int E = 16 / sizeof(Tin);
  const bool vec_ok = (n2 % E == 0) && aligned16(dy) && aligned16(saved) && aligned16(dx) && ((size_t)n2 * sizeof(Tin)) % 16 == 0 &&
                      ((size_t)n2 * sizeof(Tout)) % 16 == 0 && (!gamma || aligned16(gamma)) && (!beta || aligned16(beta));
  NormCfg c = norm_cfg(vec_ok ? n2 / E : 1, 2, 512);
  // rows wider than 2 vectors per thread: split every row over a 2-CTA cluster when that brings each half back to <= 2 vectors per thread
  static const int use_cluster = getenv("APEX_B200_LN_BWD_CLUSTER") ? atoi(getenv("APEX_B200_LN_BWD_CLUSTER")) : 1;
  int cl = 1;
  if (vec_ok && c.ok && c.maxv >= 4 && use_cluster && (n2 / E) % 2 == 0) {
    NormCfg h = norm_cfg(n2 / E / 2, 2, 512);
    // measured (benchmarks/bench_ln_sweep.py, bf16): 16384-wide rows 0.52 -> 0.55 (LN) / 0.53 -> 0.585 (RMS) of the copy bandwidth; 12288-wide
    // rows leave a quarter of the half-row's lanes idle and lose (0.45 vs 0.47), so the split is used only when the halves fill their threads
    if (h.ok && h.maxv <= 2 && (use_cluster == 2 || h.tpr * h.maxv == n2 / E / 2)) { c = h; cl = 2; }
  }
  const bool small_rows_ok = c.rows_per_cta == 1 || n2 <= 4096;
  const size_t dyn = c.maxv >= 4 ? (size_t)2 * n2 * sizeof(float) : 0;  // shared-memory accumulators of the wide-row variants
  if (vec_ok && c.ok && c.maxv <= 4 && dyn <= 160 * 1024 && small_rows_ok && ws != nullptr) {
    int nclus = (n1 + c.rows_per_cta - 1) / c.rows_per_cta;   // clusters (cl == 2) or CTAs
    const int cap = kNumSMs * (512 / c.threads);
    if (nclus > cap / cl) nclus = cap / cl;
    const int grid = nclus * cl;
    float* part_g = dgamma ? ws : nullptr;
    float* part_b = (dgamma && dbeta) ? ws + (size_t)cap * n2 : nullptr;
#define LN_BWD_GO(MV)                                                                                                               \
  {                                                                                                                                 \
    auto kern = ln_bwd_vec<MV, Tin, Tout, RMS, MEMEFF>;                                                                             \
    if (dyn) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);                                     \
    cudaLaunchConfig_t lc = {};                                                                                                     \
    lc.gridDim = dim3(grid); lc.blockDim = dim3(c.threads); lc.dynamicSmemBytes = dyn; lc.stream = st;                              \
    cudaLaunchAttribute at[1];                                                                                                      \
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; \
    lc.attrs = at; lc.numAttrs = cl > 1 ? 1 : 0;                                                                                    \
    cudaError_t le = cudaLaunchKernelEx(&lc, kern, (const Tout*)dy, saved, mean, invvar, (const Tout*)gamma, (const Tout*)beta, (Tin*)dx, \
                                        part_g, part_b, n1, n2, eps, c.tpr, cl);                                                    \
    if (le != cudaSuccess) return (int)le;                                                                                          \
  }
    switch (c.maxv) {
      case 1: LN_BWD_GO(1); break;
      case 2: LN_BWD_GO(2); break;
      default: LN_BWD_GO(4); break;
    }
    if (dgamma)
      ln_bwd_fold<Tout><<<(n2 + 31) / 32, dim3(32, 16), 0, st>>>(part_g, part_b, nclus, n2, (Tout*)dgamma, (Tout*)dbeta);
  } else {
    int grid = n1 < kNumSMs * 8 ? n1 : kNumSMs * 8;
    ln_bwd_dx_generic<Tin, Tout, RMS, MEMEFF><<<grid, 256, 0, st>>>((const Tout*)dy, saved, mean, invvar, (const Tout*)gamma,
                                                                    (const Tout*)beta, (Tin*)dx, n1, n2, eps);
    if (dgamma)
      ln_bwd_gamma_generic<Tin, Tout, RMS, MEMEFF><<<(n2 + 31) / 32, dim3(32, 16), 0, st>>>(
          (const Tout*)dy, saved, mean, invvar, (const Tout*)gamma, (const Tout*)beta, (Tout*)dgamma, (Tout*)dbeta, n1, n2, eps);
  }
  AB_CHECK_LAUNCH();
  return 0;
}

}  // namespace ab

using namespace ab;

// workspace floats needed: 2 * (148*2) * n2
AB_API int64_t ab_layer_norm_bwd_ws_floats(int n2) { return (int64_t)2 * kNumSMs * 2 * n2; }

// dy[n1,n2] (dt_out), saved = x (dt_in) or y (dt_out, memory_efficient) -> dx (dt_in), dgamma/dbeta (dt_out, nullable).
AB_API int ab_layer_norm_bwd(const void* dy, const void* saved, const float* mean, const float* invvar, const void* gamma,
                             const void* beta, void* dx, void* dgamma, void* dbeta, float* ws, int n1, int n2, float eps, int dt_in,
                             int dt_out, int rms, int memory_efficient, cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0) return 0;
#define LNB(TI, TO)                                                                                                              \
  {                                                                                                                              \
    if (rms) return memory_efficient ? ln_bwd_launch<TI, TO, true, true>(dy, saved, mean, invvar, gamma, beta, dx, dgamma, dbeta, ws, n1, n2, eps, st)   \
                                     : ln_bwd_launch<TI, TO, true, false>(dy, saved, mean, invvar, gamma, beta, dx, dgamma, dbeta, ws, n1, n2, eps, st); \
    return memory_efficient ? ln_bwd_launch<TI, TO, false, true>(dy, saved, mean, invvar, gamma, beta, dx, dgamma, dbeta, ws, n1, n2, eps, st)          \
                            : ln_bwd_launch<TI, TO, false, false>(dy, saved, mean, invvar, gamma, beta, dx, dgamma, dbeta, ws, n1, n2, eps, st);        \
  }
  if (dt_in == kF32 && dt_out == kF32) LNB(float, float)
  if (dt_in == kF16 && dt_out == kF16) LNB(f16, f16)
  if (dt_in == kBF16 && dt_out == kBF16) LNB(bf16, bf16)
  if (dt_in == kF16 && dt_out == kF32) LNB(f16, float)
  if (dt_in == kBF16 && dt_out == kF32) LNB(bf16, float)
  if (dt_in == kF32 && dt_out == kF16) LNB(float, f16)
  if (dt_in == kF32 && dt_out == kBF16) LNB(float, bf16)
  return -1;
}
