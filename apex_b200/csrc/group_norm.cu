// NHWC GroupNorm (+ fused SiLU / swish), forward and backward, as ONE persistent, dependency-driven kernel per direction:
//   stage 1  per-(image, channel, row-split) partial statistics (threads own 16-byte channel vectors, rows 4x unrolled). Every work
//            item bumps its image's arrival counter; the item that arrives LAST folds the image's groups (mean / rstd per (n, g);
//            bwd: the gamma-weighted group sums and the per-(n, c) sums, and -- when the last image is folded -- dgamma / dbeta)
//            and publishes the image by writing the launch epoch into its ready flag.
//   stage 2  y = silu?(x*A + B) with per-(n, c) coefficients (bwd: dx = g*RG + xhat*NM2 + NM1); an item only waits for ITS image's
//            flag, so images are normalised while later images are still being reduced. The second read of x comes from the 126 MB L2.
// There is no grid-wide barrier: a CTA finishes all of its stage-1 items before it can wait, and the grid is sized to be co-resident.
// Any C / G / H*W (the reference supports 23 hard-coded channel counts for G in {16, 32} and, for its Blackwell path,
// 14 (HW, C) shapes: apex/contrib/group_norm/group_norm.py:247-289,390-416; kernels apex/contrib/csrc/group_norm_v2/
// gn_cuda_kernel.cuh:195,596 use hardware clusters of 2 + DSMEM or atom.add flip barriers for the same stats->apply dependency).
#include "common.cuh"
#include <cstdio>

namespace ab {

constexpr int kGnThreads = 256;
constexpr int kGnMaxN = 8192;  // images per launch (size of the control buffer: 1 + 2 * kGnMaxN words, see group_norm.py)

struct GnArgs {
  const void* x; const void* dy; void* out;
  const void* gamma; const void* beta;   // [C] in the activation dtype or fp32 (w_fp32)
  int w_fp32;
  float* mean; float* rstd;               // [N, G]
  float* dgamma; float* dbeta;            // [C] (bwd): fp32 when w_fp32 or T is float, else T
  float* partial;                         // [N*C][S][3]
  float* chan;                            // [N*C][3] per-(n,c) merged values
  float* gsum;                            // [N*G][2] bwd: gamma-weighted group sums / M
  unsigned int* ctrl;                     // [0] done-image counter, [1 .. 1+kGnMaxN) per-image arrival counters, [1+kGnMaxN ..) per-image ready epochs
  unsigned int epoch;                     // value that marks THIS launch in the ready flags (strictly increasing per ctrl buffer)
  int N, HW, C, G, splits, splits3, silu, is_bwd;
  float eps;
};

struct GW { float mean, m2, n; };
__device__ __forceinline__ GW gw_merge(const GW& a, const GW& b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  const float n = a.n + b.n, d = b.mean - a.mean, f = b.n / n;
  return GW{a.mean + d * f, a.m2 + b.m2 + d * d * a.n * f, n};
}

template <typename T>
__device__ __forceinline__ float ld_w(const void* p, int fp32, int c) {
  return fp32 ? reinterpret_cast<const float*>(p)[c] : to_f<T>(reinterpret_cast<const T*>(p)[c]);
}
__device__ __forceinline__ float silu_f(float v) { return v * __frcp_rn(1.f + __expf(-v)); }
__device__ __forceinline__ float dsilu_f(float v) { const float s = __frcp_rn(1.f + __expf(-v)); return s * (1.f + v * (1.f - s)); }

// IS_BWD / SILU are compile-time. Threads are laid out (cx, ry): cx owns V adjacent channels (one 16-byte vector), ry strides over
// rows, so everything that depends on the channel only -- shift, gamma, beta, the per-(n, c) normalisation coefficients -- is
// computed once per work item and the inner loops are loads + 2-3 FMAs per element (+ the SiLU exponentials).
template <typename T, bool IS_BWD, bool SILU>
__global__ void __launch_bounds__(kGnThreads, 2) group_norm_kernel(GnArgs a) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float csum[2][kGnThreads / 32][64];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ dy = reinterpret_cast<const T*>(a.dy);
  T* __restrict__ out = reinterpret_cast<T*>(a.out);
  const int C = a.C, HW = a.HW, G = a.G, Cg = C / G, S = a.splits;
  const bool vec = (C % V == 0) && aligned16(x) && aligned16(out) && (!IS_BWD || aligned16(dy));
  const int cw = vec ? V : 1, lanes_c = vec ? 8 : 32, lanes_r = kGnThreads / lanes_c, tile_c = lanes_c * cw;
  const int ctiles = (C + tile_c - 1) / tile_c;
  const int cx = tid % lanes_c, ry = tid / lanes_c;

  __shared__ int s_last;
  unsigned int* img_ctr = a.ctrl + 1;
  unsigned int* ready = a.ctrl + 1 + kGnMaxN;  // FIXED offsets: a layout that depended on N would alias stale flags with counters
  const int lane_ = tid & 31;
  // ------------------------------------------------------------------ stage 1: per-(n, c, split) partial sums; the LAST item of an
  // image to finish folds that image's groups and raises the image's ready flag (no grid-wide barrier anywhere in this kernel)
  {
    const int items = a.N * ctiles * S;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int s = it % S, ct = (it / S) % ctiles, n = it / (S * ctiles);
      const int r0 = (int)((long long)HW * s / S), r1 = (int)((long long)HW * (s + 1) / S);
      const int cbase = ct * tile_c + cx * cw;
      const bool c_ok = cbase < C;
      float acc0[V], acc1[V], shift[V], sc[V], sh[V], gm[V], bt[V];
#pragma unroll
      for (int j = 0; j < V; j++) { acc0[j] = 0.f; acc1[j] = 0.f; shift[j] = 0.f; sc[j] = 0.f; sh[j] = 0.f; gm[j] = 0.f; bt[j] = 0.f; }
      const T* xb = x + (size_t)n * HW * C;
      const T* gb = IS_BWD ? dy + (size_t)n * HW * C : nullptr;
      if (c_ok) {
#pragma unroll
        for (int j = 0; j < V; j++) {
          if (j < cw && cbase + j < C) {
            if (!IS_BWD) { if (r1 > r0) shift[j] = to_f<T>(xb[(size_t)r0 * C + cbase + j]); }
            else {
              const int g = (cbase + j) / Cg;
              const float mu = a.mean[n * G + g], r = a.rstd[n * G + g];
              sc[j] = r; sh[j] = -mu * r;  // xhat = x*sc + sh
              if (SILU) { gm[j] = ld_w<T>(a.gamma, a.w_fp32, cbase + j); bt[j] = a.beta ? ld_w<T>(a.beta, a.w_fp32, cbase + j) : 0.f; }
            }
          }
        }
#pragma unroll 4
        for (int r = r0 + ry; r < r1; r += lanes_r) {
          const size_t off = (size_t)r * C + cbase;
          float xv[V], gv[V];
          if (vec) { load_vec<T, V>(xv, xb + off); if (IS_BWD) load_vec<T, V>(gv, gb + off); }
          else { xv[0] = to_f<T>(xb[off]); if (IS_BWD) gv[0] = to_f<T>(gb[off]); }
#pragma unroll
          for (int j = 0; j < V; j++) {
            if (j < cw) {
              if (!IS_BWD) { const float d = xv[j] - shift[j]; acc0[j] += d; acc1[j] = fmaf(d, d, acc1[j]); }
              else {
                const float xh = fmaf(xv[j], sc[j], sh[j]);
                float g = gv[j];
                if (SILU) g *= dsilu_f(fmaf(xh, gm[j], bt[j]));
                acc0[j] += g; acc1[j] = fmaf(g, xh, acc1[j]);
              }
            }
          }
        }
      }
      // all row lanes of a channel share the shift -> plain sums: xor-shuffles inside the warp, one shared-memory hop across warps
      __syncthreads();
#pragma unroll
      for (int j = 0; j < V; j++) {
        if (j < cw) {
          float v0 = acc0[j], v1 = acc1[j];
          for (int o = lanes_c; o < 32; o <<= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
          if (lane < lanes_c) { csum[0][wid][cx * cw + j] = v0; csum[1][wid][cx * cw + j] = v1; }
        }
      }
      __syncthreads();
      if (tid < tile_c && ct * tile_c + tid < C) {
        const int c = ct * tile_c + tid;
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int w = 0; w < kGnThreads / 32; w++) { t0 += csum[0][w][tid]; t1 += csum[1][w][tid]; }
        float* p = a.partial + (((size_t)n * C + c) * S + s) * 3;
        if (!IS_BWD) {
          const float cnt = (float)(r1 - r0);
          float m = 0.f, m2 = 0.f;
          if (cnt > 0.f) { const float mm = t0 / cnt; m = to_f<T>(xb[(size_t)r0 * C + c]) + mm; m2 = fmaxf(t1 - t0 * mm, 0.f); }
          p[0] = m; p[1] = m2; p[2] = cnt;
        } else { p[0] = t0; p[1] = t1; p[2] = 0.f; }
      }
      // ---- arrival: the last item of image n folds its groups
      __syncthreads();
      if (tid == 0) { __threadfence(); s_last = (atomicAdd(img_ctr + n, 1u) == (unsigned)(ctiles * S - 1)); }
      __syncthreads();
      if (s_last) {
        __threadfence();
        for (int g = wid; g < G; g += kGnThreads / 32) {
          const int ng = n * G + g;
          if (!IS_BWD) {
            GW w{0.f, 0.f, 0.f};
#pragma unroll 2
            for (int q = lane_; q < Cg * S; q += 32) {
              const int c = g * Cg + q / S, sp = q % S;
              const float* p = a.partial + (((size_t)n * C + c) * S + sp) * 3;
              w = gw_merge(w, GW{__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)});
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
              w = gw_merge(w, GW{__shfl_xor_sync(0xffffffffu, w.mean, o), __shfl_xor_sync(0xffffffffu, w.m2, o), __shfl_xor_sync(0xffffffffu, w.n, o)});
            if (lane_ == 0) { a.mean[ng] = w.mean; a.rstd[ng] = rsqrtf((w.n > 0.f ? w.m2 / w.n : 0.f) + a.eps); }
          } else {
            float m1 = 0.f, m2 = 0.f;
            for (int q = lane_; q < Cg; q += 32) {
              const int c = g * Cg + q;
              float s1 = 0.f, s2 = 0.f;
              for (int sp = 0; sp < S; sp++) { const float* p = a.partial + (((size_t)n * C + c) * S + sp) * 3; s1 += __ldcg(p); s2 += __ldcg(p + 1); }
              float* cc = a.chan + ((size_t)n * C + c) * 3;
              cc[0] = s1; cc[1] = s2;
              const float gmv = ld_w<T>(a.gamma, a.w_fp32, c);
              m1 += gmv * s1; m2 += gmv * s2;
            }
            m1 = warp_sum(m1); m2 = warp_sum(m2);
            if (lane_ == 0) { const float invM = 1.f / ((float)HW * (float)Cg); a.gsum[ng * 2] = m1 * invM; a.gsum[ng * 2 + 1] = m2 * invM; }
          }
        }
        __syncthreads();
        if (tid == 0) {
          img_ctr[n] = 0u;
          __threadfence();
          atomicExch(ready + n, a.epoch);                       // image n may now be normalised
          s_last = IS_BWD && a.dgamma && (atomicAdd(a.ctrl, 1u) == (unsigned)(a.N - 1));
        }
        __syncthreads();
        if (IS_BWD && s_last) {                                   // every image is folded: dgamma / dbeta over the batch
          __threadfence();
          for (int c = tid; c < C; c += kGnThreads) {
            float dg = 0.f, db = 0.f;
            for (int nn = 0; nn < a.N; nn++) { const float* cc = a.chan + ((size_t)nn * C + c) * 3; db += __ldcg(cc); dg += __ldcg(cc + 1); }
            // parameter gradients in the parameters' own dtype (fp32 weights, or the activation type): no cast kernels afterwards
            if (a.w_fp32 || sizeof(T) == 4) { a.dgamma[c] = dg; if (a.dbeta) a.dbeta[c] = db; }
            else { reinterpret_cast<T*>(a.dgamma)[c] = from_f<T>(dg); if (a.dbeta) reinterpret_cast<T*>(a.dbeta)[c] = from_f<T>(db); }
          }
          if (tid == 0) a.ctrl[0] = 0u;
        }
      }
    }
  }

  // ------------------------------------------------------------------ phase 3: apply, same (cx, ry) layout, x re-read from L2
  //   fwd: y  = silu?(x*A + B)                          A = rstd*gamma, B = beta - mean*A
  //   bwd: dx = g*RG + xhat*NM2 + NM1, xhat = x*R + MR   g = dy (* dsilu(xhat*gamma + beta)), RG = rstd*gamma, NMk = -rstd*mk
  {
    const int S3 = a.splits3;
    const int items = a.N * ctiles * S3;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int s = it % S3, ct = (it / S3) % ctiles, n = it / (S3 * ctiles);
      // wait until image n's statistics are published (every CTA has already finished ALL of its stage-1 items, so the image's
      // remaining contributors are running on other, co-resident CTAs: no deadlock)
      if (tid == 0) {
        long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned int*>(ready + n) != a.epoch) {
          if (clock64() - t0 > 20000000000LL) { printf("apex_b200 group_norm: image %d never became ready\n", n); __trap(); }
        }
        __threadfence();
      }
      __syncthreads();
      const int r0 = (int)((long long)HW * s / S3), r1 = (int)((long long)HW * (s + 1) / S3);
      const int cbase = ct * tile_c + cx * cw;
      const bool c_ok3 = cbase < C;
      float A[V], B[V], R[V], MR[V], NM1[V], NM2[V], gm[V], bt[V];
#pragma unroll
      for (int j = 0; j < V; j++) {
        A[j] = B[j] = R[j] = MR[j] = NM1[j] = NM2[j] = gm[j] = bt[j] = 0.f;
        if (j < cw && cbase + j < C) {
          const int c = cbase + j, g = c / Cg;
          const float mu = __ldcg(a.mean + n * G + g), r = __ldcg(a.rstd + n * G + g);
          gm[j] = a.gamma ? ld_w<T>(a.gamma, a.w_fp32, c) : 1.f;
          bt[j] = a.beta ? ld_w<T>(a.beta, a.w_fp32, c) : 0.f;
          if (!IS_BWD) { A[j] = r * gm[j]; B[j] = fmaf(-mu, A[j], bt[j]); }
          else {
            const float m1 = __ldcg(a.gsum + (n * G + g) * 2), m2 = __ldcg(a.gsum + (n * G + g) * 2 + 1);
            R[j] = r; MR[j] = -mu * r; A[j] = r * gm[j]; NM1[j] = -r * m1; NM2[j] = -r * m2;
          }
        }
      }
      const T* xb = x + (size_t)n * HW * C;
      const T* gb = IS_BWD ? dy + (size_t)n * HW * C : nullptr;
      T* ob = out + (size_t)n * HW * C;
#pragma unroll 4
      for (int r = r0 + ry; c_ok3 && r < r1; r += lanes_r) {
        const size_t off = (size_t)r * C + cbase;
        float xv[V], gv[V], o[V];
        if (vec) { load_vec<T, V>(xv, xb + off); if (IS_BWD) load_vec<T, V>(gv, gb + off); }
        else { xv[0] = to_f<T>(xb[off]); if (IS_BWD) gv[0] = to_f<T>(gb[off]); }
#pragma unroll
        for (int j = 0; j < V; j++) {
          if (j < cw) {
            if (!IS_BWD) {
              const float yv = fmaf(xv[j], A[j], B[j]);
              o[j] = SILU ? silu_f(yv) : yv;
            } else {
              const float xh = fmaf(xv[j], R[j], MR[j]);
              float g2 = gv[j];
              if (SILU) g2 *= dsilu_f(fmaf(xh, gm[j], bt[j]));
              o[j] = fmaf(g2, A[j], fmaf(xh, NM2[j], NM1[j]));
            }
          }
        }
        if (vec) store_vec<T, V>(ob + off, o); else ob[off] = from_f<T>(o[0]);
      }
    }
  }
}

}  // namespace ab

using namespace ab;

// scratch floats: partial N*C*S*3 + chan N*C*3 + gsum N*G*2
AB_API long long ab_group_norm_scratch_floats(int N, int C, int G, int max_splits) {
  return (long long)N * C * max_splits * 3 + (long long)N * C * 3 + (long long)N * G * 2 + 64;
}

AB_API int ab_group_norm(int is_bwd, const void* x, const void* dy, void* out, const void* gamma, const void* beta, int w_fp32, float* mean,
                         float* rstd, float* dgamma, float* dbeta, float* scratch, long long scratch_floats, unsigned int* ctrl, unsigned int epoch, int N,
                         int HW, int C, int G, float eps, int silu, int dt, cudaStream_t st) {
  if (N <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G != 0 || N > kGnMaxN) return -2;
  GnArgs a;
  a.x = x; a.dy = dy; a.out = out; a.gamma = gamma; a.beta = beta; a.w_fp32 = w_fp32; a.mean = mean; a.rstd = rstd; a.dgamma = dgamma;
  a.dbeta = dbeta; a.ctrl = ctrl; a.epoch = epoch; a.N = N; a.HW = HW; a.C = C; a.G = G; a.eps = eps; a.silu = silu; a.is_bwd = is_bwd;
  int grid = kNumSMs * 2;
  const int ctiles = (C + 63) / 64;
  long long units = (long long)N * ctiles;
  long long splits = (2LL * grid + units - 1) / units;
  if (splits > HW / 64) splits = HW / 64;
  if (splits < 1) splits = 1;
  const long long fixed = (long long)N * C * 3 + (long long)N * G * 2 + 64;
  while (splits > 1 && (long long)N * C * splits * 3 + fixed > scratch_floats) splits--;
  if ((long long)N * C * splits * 3 + fixed > scratch_floats) return -5;
  a.splits = (int)splits;
  a.partial = scratch;
  a.chan = scratch + (size_t)N * C * splits * 3;
  a.gsum = a.chan + (size_t)N * C * 3;
  const long long total = (long long)N * HW * C;
  long long want = units * splits;
  (void)total;
  long long s3 = (4LL * kNumSMs * 2 + units - 1) / units;   // row splits of the apply phase: ~4 work items per CTA
  if (s3 > HW / 32) s3 = HW / 32;
  if (s3 < 1) s3 = 1;
  a.splits3 = (int)s3;
  if (units * s3 > want) want = units * s3;
  if (want < grid) grid = (int)want;
#define GN_GO(T)                                                                                              \
  do {                                                                                                        \
    if (is_bwd) { if (silu) group_norm_kernel<T, true, true><<<grid, kGnThreads, 0, st>>>(a);                 \
                  else group_norm_kernel<T, true, false><<<grid, kGnThreads, 0, st>>>(a); }                   \
    else        { if (silu) group_norm_kernel<T, false, true><<<grid, kGnThreads, 0, st>>>(a);                \
                  else group_norm_kernel<T, false, false><<<grid, kGnThreads, 0, st>>>(a); }                  \
  } while (0)
  AB_DISPATCH_FLOAT3(dt, T, GN_GO(T));
  AB_CHECK_LAUNCH();
  return 0;
}
