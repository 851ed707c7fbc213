// SyncBatchNorm for B200: ONE persistent kernel per direction that does local statistics, the cross-GPU reduction (P2P stores
// into every peer's exchange buffer over NVLink + epoch flags — no NCCL, no extra launches) and the normalisation.
//
// Forward : [1] per-(unit, split) partial (mean, M2, n); the LAST split of a unit to finish Chan-merges the unit's channels and
//               publishes (mean, M2, n) into every rank's exchange buffer                         -> grid barrier ->
//           [2] epoch signal / wait across ranks; every CTA merges the D contributions of a slice of channels in rank order:
//               mean, inv_std, running stats                                                       -> grid barrier ->
//           [4] y = (x - mean) * inv_std * w + b (+ z) (ReLU)
// Backward: same skeleton with (sum dy, sum dy*(x-mean)); grad_w / grad_b from the local sums;
//           dx = (dy - sum_dy/N - (x-mean) * inv_std^2 * sum_dy_xmu/N) * w * inv_std  (dy masked by the ReLU of the fused block).
// A `phases` bitmask selects sub-sets so the ten standalone entry points of the reference extension
// (csrc/syncbn.cpp:71-89: welford_mean_var, welford_parallel, batchnorm_forward, reduce_bn, batchnorm_backward and the
// _c_last variants, kernels csrc/welford.cu:217-788) are the same code. NCHW and channels-last both run coalesced,
// 16-byte vectorised, 4x unrolled. Uneven per-rank batch sizes are supported: counts travel with the statistics.
#include "symm_device.cuh"

namespace ab {

constexpr int kBnThreads = 512;
// Multi-GPU launches leave kSmMargin SMs completely free (one CTA per SM on the others). The kernel spins on its peers, so it must
// never hold EVERY SM: a concurrent NCCL kernel of the same process (DDP's gradient all-reduce during backward) that cannot get an
// SM here stalls its counterpart on the peer, whose SyncBN kernel then cannot become resident either -- a cross-rank deadlock that
// was observed at 2 and 8 GPUs with a 2-CTAs-per-SM grid. (The reference's group_norm_v2 has an sm_margin knob for the same reason.)
constexpr int kSmMargin = 32;

struct BnArgs {
  const void* x; const void* dy; const void* z; void* out; void* dz;  // out: y (fwd) or dx (bwd); dz: grad of residual (bwd)
  int N, C, HW, nhwc;
  const float* weight; const float* bias;
  float* mean; float* invstd;       // [C]  fwd: written in phase 2; bwd: inputs
  float* var_biased;                // [C]  optional output (welford_mean_var)
  float* running_mean; float* running_var; float momentum; float eps;
  float* grad_w; float* grad_b;     // [C]  bwd outputs
  float* sum_dy; float* sum_dy_xmu; // [C]  bwd: global sums (phase 2 out / phase 4 in)
  float* partial;                   // [C][splits][3] scratch
  float* merged;                    // [C][3] this rank's merged statistics
  unsigned int* unit_ctr;           // [units] arrival counters (zero on entry, left zero on exit)
  float* count_total;               // [1]  total element count over all ranks
  unsigned int* grid_bar;           // [2]  {arrivals, generation}
  int splits, phases, fuse_relu, is_bwd;
  Signal sig; int channel;          // cross-GPU epoch signalling (sig.world == 1 => local only)
  uint32_t* epoch_ctr; int xchg_region;  // device-resident epoch (graph-replayable): epoch = *ctr + 1, exchange half = epoch & 1
  PeerPtrs xchg; int xchg_off;      // every rank's exchange buffer (floats): slot [xchg_off + (r*C + c)*3 + k]
};

// software grid barrier; every CTA of the (co-resident) grid calls it
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* gen = bar + 1;
    const unsigned int g = *gen;
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      long long t0 = clock64();
      while (*gen == g) {
        if (clock64() - t0 > 20000000000LL) { printf("apex_b200 syncbn: grid barrier timeout\n"); __trap(); }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

struct Wf { float mean, m2, n; };
__device__ __forceinline__ Wf wf_merge(const Wf& a, const Wf& b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  const float n = a.n + b.n, d = b.mean - a.mean, f = b.n / n;
  return Wf{a.mean + d * f, a.m2 + b.m2 + d * d * a.n * f, n};
}
__device__ __forceinline__ Wf wf_from_sums(float shift, float s, float ss, float cnt) {
  if (cnt <= 0.f) return Wf{0.f, 0.f, 0.f};
  const float m = s / cnt;
  return Wf{shift + m, fmaxf(ss - s * m, 0.f), cnt};
}
__device__ __forceinline__ Wf wf_shfl_xor(const Wf& w, int o) {
  return Wf{__shfl_xor_sync(0xffffffffu, w.mean, o), __shfl_xor_sync(0xffffffffu, w.m2, o), __shfl_xor_sync(0xffffffffu, w.n, o)};
}

template <typename T> struct VecOf { static constexpr int V = 16 / sizeof(T); };

// relu-masked incoming gradient for the fused (bn + add + relu) block
__device__ __forceinline__ float masked_grad(const BnArgs& a, bool fuse_relu, float g, float xv, float zv, int c, float mu, float is) {
  if (!fuse_relu) return g;
  const float yv = (xv - mu) * is * (a.weight ? a.weight[c] : 1.f) + (a.bias ? a.bias[c] : 0.f) + zv;
  return yv > 0.f ? g : 0.f;
}

// this rank's merged per-channel value -> local `merged` and every peer's exchange slot; bwd also emits grad_w / grad_b
__device__ __forceinline__ void publish(const BnArgs& a, int xchg_off, int c, float v0, float v1, float v2) {
  if (a.is_bwd) {
    if (a.grad_w) a.grad_w[c] = v1 * a.invstd[c];
    if (a.grad_b) a.grad_b[c] = v0;
  }
  float* m = a.merged + (size_t)c * 3;
  m[0] = v0; m[1] = v1; m[2] = v2;
  const int D = a.sig.world;
  if (D > 1) {
    for (int r = 0; r < D; r++) {
      float* dst = reinterpret_cast<float*>(a.xchg.p[r]) + xchg_off + ((size_t)a.sig.rank * a.C + c) * 3;
      st_relaxed_sys_f32(dst, v0); st_relaxed_sys_f32(dst + 1, v1); st_relaxed_sys_f32(dst + 2, v2);
    }
  }
}

// IS_BWD / NHWC / FUSED (residual add and/or ReLU present) are compile-time: every instantiation carries only its own inner loops
// (the runtime-flag version was 13.7k SASS instructions at 114 registers and stalled on instruction fetch for small layers).
template <typename T, bool IS_BWD, bool NHWC, bool FUSED>
__global__ void __launch_bounds__(kBnThreads, 1) syncbn_kernel(const __grid_constant__ BnArgs a) {
  constexpr int V = VecOf<T>::V;
  // every CTA reads the counter here; it is advanced after the grid barrier that precedes the exchange. `a` stays read-only.
  const uint32_t epoch = a.epoch_ctr ? *reinterpret_cast<volatile uint32_t*>(a.epoch_ctr) + 1u : a.sig.epoch;
  const int xchg_off = a.epoch_ctr ? (int)(epoch & 1u) * a.xchg_region : a.xchg_off;
  __shared__ float sm[3][kBnThreads + 8];
  __shared__ int s_last;
  __shared__ float csum[2][kBnThreads / 32][64];  // NHWC: per-warp channel sums of one item
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ dy = reinterpret_cast<const T*>(a.dy);
  const T* __restrict__ z = FUSED ? reinterpret_cast<const T*>(a.z) : nullptr;
  const bool fuse_relu = FUSED && a.fuse_relu;
  const long long per_c = (long long)a.N * a.HW;
  const int S = a.splits;
  const int C = a.C, HW = a.HW;

  // ------------------------------------------------------------------ phase 1: partial statistics (+ per-unit merge/publish)
  if (a.phases & 1) {
    if (!NHWC) {
      // NCHW: unit = channel; a split is a range of the channel's N*HW elements; threads run along HW (coalesced)
      const bool vec = (HW % V == 0) && aligned16(x) && (!IS_BWD || aligned16(dy)) && (!z || aligned16(z));
      const int items = C * S;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int c = it / S, s = it - c * S;
        long long e0 = per_c * s / S, e1 = per_c * (s + 1) / S;
        if (vec) { e0 = e0 / V * V; e1 = (s == S - 1) ? per_c : e1 / V * V; }
        float acc0 = 0.f, acc1 = 0.f, cnt = 0.f, shift = 0.f, mu = 0.f, is = 0.f;
        if (!IS_BWD) {
          if (e1 > e0) { const long long n0 = e0 / HW; shift = to_f<T>(x[(n0 * C + c) * (long long)HW + (e0 - n0 * HW)]); }
        } else { mu = a.mean[c]; is = a.invstd[c]; }
        const int step = vec ? V : 1;
#pragma unroll 4
        for (long long e = e0 + (long long)tid * step; e < e1; e += (long long)kBnThreads * step) {
          const long long n = (per_c < 0x7fffffffLL) ? (long long)((unsigned)e / (unsigned)HW) : e / HW;
          const long long off = (n * C + c) * (long long)HW + (e - n * HW);
          float xv[V], gv[V], zv[V];
          if (vec) {
            load_vec<T, V>(xv, x + off);
            if (IS_BWD) load_vec<T, V>(gv, dy + off);
            if (IS_BWD && fuse_relu && z) load_vec<T, V>(zv, z + off);
          } else {
            xv[0] = to_f<T>(x[off]);
            if (IS_BWD) gv[0] = to_f<T>(dy[off]);
            if (IS_BWD && fuse_relu && z) zv[0] = to_f<T>(z[off]);
          }
#pragma unroll
          for (int j = 0; j < V; j++) {
            if (j < step) {
              if (!IS_BWD) { const float d = xv[j] - shift; acc0 += d; acc1 += d * d; cnt += 1.f; }
              else {
                const float g = masked_grad(a, fuse_relu, gv[j], xv[j], (fuse_relu && z) ? zv[j] : 0.f, c, mu, is);
                acc0 += g; acc1 += g * (xv[j] - mu);
              }
            }
          }
        }
        // block reduction -> partial[c][s]
        float* pp = a.partial + ((size_t)c * S + s) * 3;
        if (!IS_BWD) {
          Wf w = wf_from_sums(shift, acc0, acc1, cnt);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) w = wf_merge(w, wf_shfl_xor(w, o));
          __syncthreads();
          if (lane == 0) { sm[0][wid] = w.mean; sm[1][wid] = w.m2; sm[2][wid] = w.n; }
          __syncthreads();
          if (wid == 0) {
            Wf t = lane < kBnThreads / 32 ? Wf{sm[0][lane], sm[1][lane], sm[2][lane]} : Wf{0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t = wf_merge(t, wf_shfl_xor(t, o));
            if (lane == 0) { pp[0] = t.mean; pp[1] = t.m2; pp[2] = t.n; }
          }
        } else {
          acc0 = warp_sum(acc0); acc1 = warp_sum(acc1);
          __syncthreads();
          if (lane == 0) { sm[0][wid] = acc0; sm[1][wid] = acc1; }
          __syncthreads();
          if (wid == 0) {
            float t0 = lane < kBnThreads / 32 ? sm[0][lane] : 0.f, t1 = lane < kBnThreads / 32 ? sm[1][lane] : 0.f;
            t0 = warp_sum(t0); t1 = warp_sum(t1);
            if (lane == 0) { pp[0] = t0; pp[1] = t1; pp[2] = 0.f; }
          }
        }
        // last split of this channel merges and publishes
        __syncthreads();
        if (tid == 0) { __threadfence(); s_last = (atomicAdd(a.unit_ctr + c, 1u) == (unsigned)(S - 1)); }
        __syncthreads();
        if (s_last) {
          __threadfence();
          if (wid == 0) {
            const float* base = a.partial + (size_t)c * S * 3;
            if (!IS_BWD) {
              Wf w{0.f, 0.f, 0.f};
              for (int q = lane; q < S; q += 32) w = wf_merge(w, Wf{__ldcg(base + q * 3), __ldcg(base + q * 3 + 1), __ldcg(base + q * 3 + 2)});
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) w = wf_merge(w, wf_shfl_xor(w, o));
              if (lane == 0) publish(a, xchg_off, c, w.mean, w.m2, w.n);
            } else {
              float s0 = 0.f, s1 = 0.f;
              for (int q = lane; q < S; q += 32) { s0 += __ldcg(base + q * 3); s1 += __ldcg(base + q * 3 + 1); }
              s0 = warp_sum(s0); s1 = warp_sum(s1);
              if (lane == 0) publish(a, xchg_off, c, s0, s1, (float)per_c);
            }
            if (lane == 0) { a.unit_ctr[c] = 0u; __threadfence_system(); }
          }
        }
      }
    } else {
      // NHWC: unit = channel tile; thread (cx, ry): cx owns V (or 1) adjacent channels, ry strides over rows (4x unrolled)
      const bool vec = (C % V == 0) && aligned16(x) && (!IS_BWD || aligned16(dy)) && (!z || aligned16(z));
      const int cw = vec ? V : 1;            // channels per thread
      const int lanes_c = vec ? 8 : 32;      // threads along channels
      const int lanes_r = kBnThreads / lanes_c;
      const int tile_c = lanes_c * cw;       // 64 or 32 channels per tile
      const int ctiles = (C + tile_c - 1) / tile_c;
      const int cx = tid % lanes_c, ry = tid / lanes_c;
      const int items = ctiles * S;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int ct = it / S, s = it - ct * S;
        const long long r0 = per_c * s / S, r1 = per_c * (s + 1) / S;  // rows = (n, hw) positions
        const int cbase = ct * tile_c + cx * cw;
        const bool c_ok = cbase < C;
        float acc0[V], acc1[V], shift[V], mu[V], is[V], cnt = 0.f;
#pragma unroll
        for (int j = 0; j < V; j++) { acc0[j] = 0.f; acc1[j] = 0.f; shift[j] = 0.f; mu[j] = 0.f; is[j] = 0.f; }
        if (c_ok) {
#pragma unroll
          for (int j = 0; j < V; j++) {
            if (j < cw && cbase + j < C) {
              if (!IS_BWD) { if (r1 > r0) shift[j] = to_f<T>(x[r0 * C + cbase + j]); }
              else { mu[j] = a.mean[cbase + j]; is[j] = a.invstd[cbase + j]; }
            }
          }
#pragma unroll 4
          for (long long r = r0 + ry; r < r1; r += lanes_r) {
            const long long off = r * C + cbase;
            float xv[V], gv[V], zv[V];
            if (vec) {
              load_vec<T, V>(xv, x + off);
              if (IS_BWD) load_vec<T, V>(gv, dy + off);
              if (IS_BWD && fuse_relu && z) load_vec<T, V>(zv, z + off);
            } else {
              xv[0] = to_f<T>(x[off]);
              if (IS_BWD) gv[0] = to_f<T>(dy[off]);
              if (IS_BWD && fuse_relu && z) zv[0] = to_f<T>(z[off]);
            }
            cnt += 1.f;
#pragma unroll
            for (int j = 0; j < V; j++) {
              if (j < cw) {
                if (!IS_BWD) { const float d = xv[j] - shift[j]; acc0[j] += d; acc1[j] += d * d; }
                else {
                  const float g = masked_grad(a, fuse_relu, gv[j], xv[j], (fuse_relu && z) ? zv[j] : 0.f, cbase + j, mu[j], is[j]);
                  acc0[j] += g; acc1[j] += g * (xv[j] - mu[j]);
                }
              }
            }
          }
        }
        // reduce over the row lanes: every thread of a channel column used the same shift, so plain sums combine exactly --
        // xor-shuffles inside the warp (lanes with equal cx), then one shared-memory hop across the 8 warps
        __syncthreads();  // the previous item's readers are done with csum
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
          for (int w = 0; w < kBnThreads / 32; w++) { t0 += csum[0][w][tid]; t1 += csum[1][w][tid]; }
          float* p = a.partial + ((size_t)c * S + s) * 3;
          if (!IS_BWD) {
            const float sh = r1 > r0 ? to_f<T>(x[r0 * C + c]) : 0.f;
            const Wf w = wf_from_sums(sh, t0, t1, (float)(r1 - r0));
            p[0] = w.mean; p[1] = w.m2; p[2] = w.n;
          } else { p[0] = t0; p[1] = t1; p[2] = 0.f; }
        }
        // last split of this channel tile merges and publishes its channels: thread (ch = tid % tile_c, grp = tid / tile_c)
        __syncthreads();
        if (tid == 0) { __threadfence(); s_last = (atomicAdd(a.unit_ctr + ct, 1u) == (unsigned)(S - 1)); }
        __syncthreads();
        if (s_last) {
          __threadfence();
          // warp w merges channels w, w+8, ...: lanes stride over the S split partials (independent loads), shuffle tree at the end.
          // (A thread-per-(channel, group) loop made this one CTA's serial tail ~150 us long at S = 592 while 295 CTAs waited.)
          for (int ch = wid; ch < tile_c; ch += kBnThreads / 32) {
            const int c = ct * tile_c + ch;
            if (c >= C) continue;
            const float* base = a.partial + (size_t)c * S * 3;
            if (!IS_BWD) {
              Wf w{0.f, 0.f, 0.f};
#pragma unroll 4
              for (int q = lane; q < S; q += 32) w = wf_merge(w, Wf{__ldcg(base + q * 3), __ldcg(base + q * 3 + 1), __ldcg(base + q * 3 + 2)});
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) w = wf_merge(w, wf_shfl_xor(w, o));
              if (lane == 0) publish(a, xchg_off, c, w.mean, w.m2, w.n);
            } else {
              float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
              for (int q = lane; q < S; q += 32) { s0 += __ldcg(base + q * 3); s1 += __ldcg(base + q * 3 + 1); }
              s0 = warp_sum(s0); s1 = warp_sum(s1);
              if (lane == 0) publish(a, xchg_off, c, s0, s1, (float)per_c);
            }
          }
          __threadfence_system();
          __syncthreads();
          if (tid == 0) a.unit_ctr[ct] = 0u;
        }
      }
    }
    if (a.phases & 6) grid_barrier(a.grid_bar, gridDim.x);
  }

  // ------------------------------------------------------------------ phase 2: cross-GPU exchange, finalize (all CTAs)
  if (a.phases & 2) {
    const int D = a.sig.world, rank = a.sig.rank;
    if (D > 1) {
      if (blockIdx.x == 0) { __threadfence_system(); signal_all_e(a.sig, epoch, a.channel, tid); }
      wait_all_e(a.sig, epoch, a.channel, tid);
      __syncthreads();
      if (a.epoch_ctr && blockIdx.x == 0 && tid == 0) *a.epoch_ctr = epoch;
    }
    const float* mine = D > 1 ? reinterpret_cast<const float*>(a.xchg.p[rank]) + xchg_off : nullptr;
    for (int c = blockIdx.x * kBnThreads + tid; c < C; c += gridDim.x * kBnThreads) {
      if (!IS_BWD) {
        Wf w{0.f, 0.f, 0.f};
        if (D > 1) {
          for (int r = 0; r < D; r++) { const float* p = mine + ((size_t)r * C + c) * 3; w = wf_merge(w, Wf{ld_relaxed_sys_f32(p), ld_relaxed_sys_f32(p + 1), ld_relaxed_sys_f32(p + 2)}); }
        } else {
          const float* p = a.merged + (size_t)c * 3; w = Wf{__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)};
        }
        const float var_b = w.n > 0.f ? w.m2 / w.n : 0.f;
        a.mean[c] = w.mean;
        if (a.var_biased) a.var_biased[c] = var_b;
        if (a.invstd) a.invstd[c] = rsqrtf(var_b + a.eps);
        if (a.running_mean) {
          const float var_u = w.n > 1.f ? w.m2 / (w.n - 1.f) : var_b;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * w.mean;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * var_u;
        }
        if (c == 0 && a.count_total) a.count_total[0] = w.n;
      } else {
        float s0 = 0.f, s1 = 0.f, nt = 0.f;
        if (D > 1) {
          for (int r = 0; r < D; r++) { const float* p = mine + ((size_t)r * C + c) * 3; s0 += ld_relaxed_sys_f32(p); s1 += ld_relaxed_sys_f32(p + 1); nt += ld_relaxed_sys_f32(p + 2); }
        } else {
          const float* p = a.merged + (size_t)c * 3; s0 = __ldcg(p); s1 = __ldcg(p + 1); nt = __ldcg(p + 2);
        }
        a.sum_dy[c] = s0; a.sum_dy_xmu[c] = s1;
        if (c == 0 && a.count_total) a.count_total[0] = nt;
      }
    }
    if (a.phases & 4) grid_barrier(a.grid_bar, gridDim.x);
  }

  // ------------------------------------------------------------------ phase 3: elementwise, in memory order, 16-byte vectors
  // Per-channel coefficients so the inner loop is 1-3 FMAs per element:
  //   fwd: y  = x*sc + sh (+z)(relu)                 sc = inv_std*w, sh = b - mean*sc
  //   bwd: dx = g*A + x*Bc + Cc                      A = w*inv_std, Bc = -inv_std^3*w*sum_dy_xmu/N, Cc = -A*sum_dy/N - mean*Bc
  if (a.phases & 4) {
    T* __restrict__ out = reinterpret_cast<T*>(a.out);
    T* __restrict__ dz = (FUSED && IS_BWD) ? reinterpret_cast<T*>(a.dz) : nullptr;
    const float inv_n = IS_BWD ? 1.f / __ldcg(a.count_total) : 0.f;
    const bool vec = ((NHWC ? C : HW) % V == 0) && aligned16(x) && aligned16(out) && (!IS_BWD || aligned16(dy)) && (!z || aligned16(z)) &&
                     (!dz || aligned16(dz));
    const int step = vec ? V : 1;
    auto coef = [&](int c, float& sc, float& sh, float& A, float& Bc, float& Cc) {
      const float mu = __ldcg(a.mean + c), is = __ldcg(a.invstd + c), w = a.weight ? a.weight[c] : 1.f;
      sc = is * w; sh = (a.bias ? a.bias[c] : 0.f) - mu * sc;
      if (IS_BWD) {
        A = sc; Bc = -is * is * sc * __ldcg(a.sum_dy_xmu + c) * inv_n; Cc = -A * __ldcg(a.sum_dy + c) * inv_n - mu * Bc;
      } else { A = 0.f; Bc = 0.f; Cc = 0.f; }
    };
    auto apply = [&](long long i, const float (&sc)[V], const float (&sh)[V], const float (&A)[V], const float (&Bc)[V], const float (&Cc)[V]) {
      float xv[V], gv[V], zv[V], o[V], gz[V];
      if (vec) {
        load_vec<T, V>(xv, x + i);
        if (IS_BWD) load_vec<T, V>(gv, dy + i);
        if (z) load_vec<T, V>(zv, z + i);
      } else {
        xv[0] = to_f<T>(x[i]);
        if (IS_BWD) gv[0] = to_f<T>(dy[i]);
        if (z) zv[0] = to_f<T>(z[i]);
      }
#pragma unroll
      for (int j = 0; j < V; j++) {
        if (j < step) {
          if (!IS_BWD) {
            float yv = fmaf(xv[j], sc[j], sh[j]);
            if (z) yv += zv[j];
            if (fuse_relu) yv = fmaxf(yv, 0.f);
            o[j] = yv;
          } else {
            float g = gv[j];
            if (fuse_relu) { float yv = fmaf(xv[j], sc[j], sh[j]); if (z) yv += zv[j]; if (yv <= 0.f) g = 0.f; }
            gz[j] = g;
            o[j] = fmaf(g, A[j], fmaf(xv[j], Bc[j], Cc[j]));
          }
        }
      }
      if (vec) {
        store_vec<T, V>(out + i, o);
        if (IS_BWD && dz) store_vec<T, V>(dz + i, gz);
      } else {
        out[i] = from_f<T>(o[0]);
        if (IS_BWD && dz) dz[i] = from_f<T>(gz[0]);
      }
    };
    float sc[V], sh[V], A[V], Bc[V], Cc[V];
    if (NHWC) {
      const long long total = (long long)a.N * C * HW;
      const long long gthreads = (long long)gridDim.x * kBnThreads;
      const int cvecs = C / step;
      const long long gtid = (long long)blockIdx.x * kBnThreads + tid;
      if (gthreads % cvecs == 0) {
        // every vector this thread touches has the same channels: coefficients live in registers for the whole loop
        const int c0 = (int)(gtid % cvecs) * step;
#pragma unroll
        for (int j = 0; j < V; j++) if (j < step) coef(c0 + j, sc[j], sh[j], A[j], Bc[j], Cc[j]);
#pragma unroll 2
        for (long long i = gtid * step; i < total; i += gthreads * step) apply(i, sc, sh, A, Bc, Cc);
      } else {
        for (long long i = gtid * step; i < total; i += gthreads * step) {
          const int c0 = (int)(i % C);
#pragma unroll
          for (int j = 0; j < V; j++) if (j < step) coef(c0 + j, sc[j], sh[j], A[j], Bc[j], Cc[j]);
          apply(i, sc, sh, A, Bc, Cc);
        }
      }
    } else {
      // NCHW: tile = 256*step consecutive elements of one (n, c) plane -> one channel lookup per tile
      const int tile = kBnThreads * step;
      const int tpp = (HW + tile - 1) / tile;
      const long long tiles = (long long)a.N * C * tpp;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long plane = t / tpp;
        const int c = (int)(plane % C);
        const int in_plane = (int)(t - plane * tpp) * tile + tid * step;
        if (in_plane < HW) {
          float s0, s1, s2, s3, s4;
          coef(c, s0, s1, s2, s3, s4);
#pragma unroll
          for (int j = 0; j < V; j++) { sc[j] = s0; sh[j] = s1; A[j] = s2; Bc[j] = s3; Cc[j] = s4; }
          apply(plane * HW + in_plane, sc, sh, A, Bc, Cc);
        }
      }
    }
  }
}

}  // namespace ab

using namespace ab;

// One entry point for every SyncBN operation; see the header comment for `phases`. x is contiguous NCHW (nhwc=0) or
// contiguous channels-last (nhwc=1), viewed as [N, C, HW] / [N, HW, C].
// scratch: float buffer of `scratch_floats`: [merged: 3C][unit counters: C uint32][partials: the rest].
AB_API int ab_syncbn(int is_bwd, int phases, const void* x, const void* dy, const void* z, void* out, void* dz, int N, int C, int HW,
                     int nhwc, const float* weight, const float* bias, float* mean, float* invstd, float* var_biased,
                     float* running_mean, float* running_var, float momentum, float eps, float* grad_w, float* grad_b, float* sum_dy,
                     float* sum_dy_xmu, float* scratch, long long scratch_floats, float* count_total, unsigned int* grid_bar,
                     int fuse_relu, const uint64_t* pads, const uint64_t* xchg, int xchg_off, int rank, int world, unsigned int epoch,
                     unsigned int* epoch_ctr, int xchg_region, int sm_margin, int channel, int dt, cudaStream_t st) {
  if (C <= 0 || N <= 0 || HW <= 0) return 0;
  BnArgs a;
  a.x = x; a.dy = dy; a.z = z; a.out = out; a.dz = dz; a.N = N; a.C = C; a.HW = HW; a.nhwc = nhwc;
  a.weight = weight; a.bias = bias; a.mean = mean; a.invstd = invstd; a.var_biased = var_biased; a.running_mean = running_mean;
  a.running_var = running_var; a.momentum = momentum; a.eps = eps; a.grad_w = grad_w; a.grad_b = grad_b; a.sum_dy = sum_dy;
  a.sum_dy_xmu = sum_dy_xmu; a.count_total = count_total; a.grid_bar = grid_bar; a.phases = phases;
  a.fuse_relu = fuse_relu; a.is_bwd = is_bwd; a.channel = channel; a.xchg_off = xchg_off;
  a.merged = scratch;
  // the arrival counters live at a FIXED place (the tail of the scratch) so that calls with different C never alias them with
  // the merged / partial floats of an earlier call: they must read zero on entry
  constexpr long long kCtrCap = 16384;
  if (C > kCtrCap || scratch_floats < kCtrCap + 6LL * C) return -5;
  a.unit_ctr = reinterpret_cast<unsigned int*>(scratch + (scratch_floats - kCtrCap));
  a.partial = scratch + (size_t)3 * C;
  const long long partial_cap = scratch_floats - kCtrCap - (long long)3 * C;
  if (partial_cap < (long long)3 * C) return -5;
  for (int i = 0; i < kMaxPeers; i++) {
    a.sig.pads.p[i] = (pads && i < world) ? (void*)pads[i] : nullptr;
    a.xchg.p[i] = (xchg && i < world) ? (void*)xchg[i] : nullptr;
  }
  a.sig.rank = rank; a.sig.world = world; a.sig.epoch = epoch;
  a.epoch_ctr = world > 1 ? epoch_ctr : nullptr; a.xchg_region = xchg_region;
  const long long per_c = (long long)N * HW;
  // the software grid barrier needs every CTA resident: one 512-thread CTA per SM, minus the NCCL margin when peers are involved
  // (sm_margin < 0 => the default kSmMargin; the reference exposes the same knob: gn_cuda_host_template.cuh:50-61)
  if (sm_margin < 0) sm_margin = kSmMargin;
  if (sm_margin > kNumSMs - 8) sm_margin = kNumSMs - 8;
  int grid = world > 1 ? kNumSMs - sm_margin : kNumSMs;
  const int units = nhwc ? (C + 63) / 64 : C;
  long long splits = (2LL * grid + units - 1) / units;
  const long long min_per_split = nhwc ? 128 : 4096;  // rows / elements: keep every split a few full passes long
  if (splits > per_c / min_per_split) splits = per_c / min_per_split;
  if (splits > partial_cap / (3LL * C)) splits = partial_cap / (3LL * C);
  if (splits > 1024) splits = 1024;
  if (splits < 1) splits = 1;
  a.splits = (int)splits;
  const long long total = (long long)N * C * HW;
  long long want = 1;
  if (phases & 1) want = (long long)units * splits;
  if (phases & 2) { const long long w2 = ((long long)C + kBnThreads - 1) / kBnThreads; if (w2 > want) want = w2; }
  if (phases & 4) { const long long w4 = (total + kBnThreads * 16 - 1) / (kBnThreads * 16); if (w4 > want) want = w4; }
  if (want < grid) grid = (int)want;
  const bool fused = fuse_relu || z != nullptr || dz != nullptr;
  // Cooperative launch: the software grid barrier (and the spin on the peers) needs every CTA of the grid resident at once; the
  // driver guarantees all-or-nothing placement (or fails the launch) instead of this kernel hoping that no other stream holds SMs.
  void* kargs[] = {(void*)&a};
  cudaError_t lerr = cudaSuccess;
#define BN_GO4(T, B, H, F) lerr = cudaLaunchCooperativeKernel((const void*)syncbn_kernel<T, B, H, F>, dim3(grid), dim3(kBnThreads), kargs, 0, st)
#define BN_GO(T)                                                                                   \
  do {                                                                                             \
    if (is_bwd) { if (nhwc) { if (fused) BN_GO4(T, true, true, true); else BN_GO4(T, true, true, false); }        \
                  else      { if (fused) BN_GO4(T, true, false, true); else BN_GO4(T, true, false, false); } }    \
    else        { if (nhwc) { if (fused) BN_GO4(T, false, true, true); else BN_GO4(T, false, true, false); }      \
                  else      { if (fused) BN_GO4(T, false, false, true); else BN_GO4(T, false, false, false); } }  \
  } while (0)
  AB_DISPATCH_FLOAT3(dt, T, BN_GO(T));
  if (lerr != cudaSuccess) return (int)lerr;
  AB_CHECK_LAUNCH();
  return 0;
}
