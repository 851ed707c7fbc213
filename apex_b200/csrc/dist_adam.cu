// ZeRO-2 optimizer step as in-kernel collectives over NVLink 5 / NVSwitch (no NCCL on this path):
//
//   MODE_FUSED : reduce-scatter (pull peers' grad shards, or ONE multimem.ld_reduce through the switch) + scale + grad-norm
//                + Adam/AdamW on the fp32 shard + all-gather (push the new low-precision params to every rank, or ONE
//                multimem.st) — a single pass; HBM traffic of the update hides under the link traffic.
//   MODE_RS    : reduce-scatter + norm only (writes the fp32 reduced shard) — used when the update needs the global norm
//                first (clipping / GradScaler), and per bucket while backward is still running (overlap_grad_sync).
//   MODE_ADAM  : Adam on the reduced shard + all-gather push.
//   MODE_PUSH  : all-gather only: cast this rank's fp32 shard to the parameter dtype and push it to every rank (the in-kernel
//                all-gather of optimizers whose update runs in other kernels: DistributedFusedLAMB).
//
// Replaces the reference pipeline: _grad_copy into buckets (distributed_fused_adam.py:1600-1666) -> NCCL
// reduce_scatter_tensor (:1929-1947) -> multi_tensor_l2norm (:2216) -> DistAdamFunctor (multi_tensor_distopt_adam_kernel.cu:84-168)
// -> NCCL all_gather_into_tensor (:2067-2082) -> maybe_cast_mt copy-out (:1716-1767).
//
// Layout: the flat parameter space is cut into `n_buckets` buckets of `bucket_elems`; bucket b is sharded D ways, rank r owns
// flat range [b*B + r*Sb, b*B + (r+1)*Sb). Local state (p, m, v fp32) is bucket-major: local index = b*Sb + o.
// Cross-rank ordering uses epoch signals (symm_device.cuh): start = "my grads are final and my params buffer may be
// overwritten"; end = "I no longer read your grads and everything I pushed into your params is visible".
#include "symm_device.cuh"
#include <cstdlib>

namespace ab {

constexpr int kDChunk = 2048;   // elements per CTA work item (256 threads x 8)
constexpr int kDThreads = 256;

enum { MODE_FUSED = 0, MODE_RS = 1, MODE_ADAM = 2, MODE_PUSH = 3 };

struct DistArgs {
  PeerPtrs grads;    // every rank's full gradient buffer (TG) as mapped here
  PeerPtrs params;   // every rank's full parameter buffer (TP)
  const void* mc_grads;  // multicast alias of the gradient buffers (NVLS) or null
  void* mc_params;       // multicast alias of the parameter buffers or null
  float* p; float* m; float* v;  // local fp32 shards
  short* rem;                    // store_param_remainders: p == null, the fp32 master is (this rank's bf16 parameter << 16) + signed int16
                                 // remainder (reference multi_tensor_distopt_adam_kernel.cu:319-429); halves the master-weight traffic
  float* reduced;                // local fp32 reduced-gradient shard (MODE_RS out / MODE_ADAM in)
  long long bucket_elems;
  int shard_elems, bucket_begin, bucket_end;
  int lay_rank;  // which shard of every bucket this rank owns (== sig.rank on the fused path; sig.world may be 1 when the
                 // collectives are done by NCCL around the kernel and only the layout is sharded)
  Signal sig;
  uint32_t* epoch_ctr;  // device-resident epoch of this pad channel (graph-replayable): the launch uses *epoch_ctr + 1 and the
                        // closing CTA stores it back; null => sig.epoch was chosen by the host
  int chan_start, chan_end, norm_slot;
  unsigned int* done_ctr;
  float* norm_partials;  // [gridDim.x]
  float* norm_out;       // [0] = this shard's sum of squares, [1] = sum over all ranks
  const float* grad_scale;
  float pre_scale;
  float lr, beta1, beta2, bc1, bc2, eps, decay;
  int mode;
  const int* noop;
  const float* lr_ptr;   // capturable: learning rate and step count live on the device
  const int* step_ptr;
  int bias_correction;
  // overlap_param_sync (all-gather fused into the consuming GEMM): when every chunk of this rank's shard of bucket b has been pushed,
  // `ready_epoch` is released into slot [b][rank] of EVERY rank's flag array; consumers (csrc/gemm_sm100.cu TMA producer, or a
  // cuStreamWaitValue32) acquire the D slots of the buckets that hold the weight tile they are about to load.
  PeerPtrs ready;          // every rank's flag array [n_buckets][kMaxPeers] uint32 (null => feature off)
  uint32_t* bucket_ctr;    // local [n_buckets] chunk counters, zero between launches
  uint32_t ready_epoch;
  // world == 1 only: gradients read in place from wherever autograd left them (no copy into the contiguous buffer). Table of
  // {first chunk of the parameter in the flat space, gradient address, numel} sorted by chunk; every parameter starts on a chunk
  // boundary in this layout, so a chunk has exactly one source. Null => the contiguous buffer.
  const long long* src_tab; int src_n;
  int skip_start;          // the start barrier already ran as a 1-CTA gate kernel in front of this launch (ab_symm_gate)
};

template <typename T> __device__ __forceinline__ void unpack8(const uint4* raw, float (&f)[8]) {
  words_to_float<T, 8>(reinterpret_cast<const uint32_t*>(raw), f);
}

__device__ __forceinline__ float lerp_f(float t, float x, float y) { return fmaf(t, y, fmaf(-t, x, x)); }

// One 8-element vector of the update: Adam on fp32 (p, m, v), returns the new parameter values in p.
struct AdamScalars { float gs, lr, bc1, bc2, beta1, beta2, eps, decay; int mode; };

__device__ __forceinline__ void adam8(float (&p)[8], float (&m)[8], float (&v)[8], const float (&g)[8], const AdamScalars& h) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float sg = g[i] * h.gs;
    if (h.mode == 0) sg += h.decay * p[i];
    m[i] = lerp_f(h.beta1, sg, m[i]);
    v[i] = lerp_f(h.beta2, sg * sg, v[i]);
    float upd = (m[i] / h.bc1) / (sqrtf(v[i] / h.bc2) + h.eps);
    if (h.mode != 0) upd += h.decay * p[i];
    p[i] -= h.lr * upd;
  }
}
// evict-first variants: optimizer state is touched exactly once per step, keeping it in L2 only displaces lines that are about to be written
__device__ __forceinline__ void ld8_cs(const float* src, float (&d)[8]) {
  float4 a, b;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src));
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + 4));
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void st8_cs(float* dst, const float (&d)[8]) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(d[0]), "f"(d[1]), "f"(d[2]), "f"(d[3]) : "memory");
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4), "f"(d[4]), "f"(d[5]), "f"(d[6]), "f"(d[7]) : "memory");
}
__device__ __forceinline__ void ld8(const float* src, float (&d)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void st8(float* dst, const float (&d)[8]) {
  *reinterpret_cast<float4*>(dst) = make_float4(d[0], d[1], d[2], d[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(d[4], d[5], d[6], d[7]);
}

// DP: compile-time bound on the number of peers looped over (1, 2, 4, 8); U: chunks whose remote gradient loads are issued
// before any of them is consumed (NVLink round trips are ~2-4 us: bytes in flight per SM, not threads, set the bandwidth).
// MB: minimum resident CTAs per SM the register budget is set for; CS: evict-first loads / stores of the optimizer state.
template <typename TG, typename TP, int MODE, bool NVLS, int DP, int U, int MB = ((!NVLS && DP == 1) ? 3 : 2), bool CS = false>
__global__ void __launch_bounds__(kDThreads, MB) dist_step_kernel(const __grid_constant__ DistArgs a) {
  // every CTA reads the counter before the closing CTA can store it back; `a` itself stays read-only (constant bank, no stack copy)
  const uint32_t epoch = a.epoch_ctr ? *reinterpret_cast<volatile uint32_t*>(a.epoch_ctr) + 1u : a.sig.epoch;
  constexpr int GV = sizeof(TG) * 8 / 16;  // 16-byte vectors per 8 gradient elements
  constexpr int PV = sizeof(TP) * 8 / 16;
  constexpr int NP = NVLS ? 1 : DP;        // gradient sources read per element
  constexpr bool kReads = MODE == MODE_FUSED || MODE == MODE_RS;      // pulls / reduces gradients (start barrier, norm)
  constexpr bool kUpdates = MODE == MODE_FUSED || MODE == MODE_ADAM;  // applies Adam to (p, m, v)
  __shared__ float red[40];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int D = a.sig.world, rank = a.sig.rank;

  if (kReads && D > 1 && !a.skip_start) {
    if (blockIdx.x == 0) signal_all_e(a.sig, epoch, a.chan_start, tid);
    wait_all_e(a.sig, epoch, a.chan_start, tid);
    __syncthreads();
  }

  const bool skip = kUpdates && a.noop != nullptr && *a.noop != 0;
  AdamScalars h{1.f, a.lr, a.bc1, a.bc2, a.beta1, a.beta2, a.eps, a.decay, a.mode};
  if (MODE != MODE_RS) {
    if (a.grad_scale) h.gs = *a.grad_scale;
    if (a.lr_ptr) h.lr = *a.lr_ptr;
    if (a.step_ptr && a.bias_correction) {
      const float sf = (float)(*a.step_ptr);
      h.bc1 = 1.f - powf(a.beta1, sf);
      h.bc2 = 1.f - powf(a.beta2, sf);
    }
  }
  const int cpb = a.shard_elems / kDChunk;  // chunks per bucket shard
  const long long c0 = (long long)a.bucket_begin * cpb, c1 = (long long)a.bucket_end * cpb;
  float nsq = 0.f;

  if (!skip) {
    for (long long cb = c0 + blockIdx.x; cb < c1; cb += (long long)gridDim.x * U) {
      long long local[U], flat[U];
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long c = cb + (long long)u * gridDim.x;
        on[u] = c < c1;
        const long long b = c / cpb;
        const int oc = (int)(c - b * cpb);
        local[u] = c * kDChunk + tid * 8;
        flat[u] = b * a.bucket_elems + (long long)a.lay_rank * a.shard_elems + (long long)oc * kDChunk + tid * 8;
      }
      // ---- phase 1: every gradient load of the U chunks is in flight before the first use
      uint4 raw[U][NP][GV];
      float red_in[U][8];
      if (MODE == MODE_ADAM) {
#pragma unroll
        for (int u = 0; u < U; u++) if (on[u]) ld8(a.reduced + local[u], red_in[u]);
      } else if (kReads) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (on[u]) {
            if (NVLS) {
#pragma unroll
              for (int q = 0; q < GV; q++)
                raw[u][0][q] = multimem_ld_reduce16<TG>(reinterpret_cast<const char*>(a.mc_grads) + flat[u] * sizeof(TG) + q * 16);
            } else if (DP == 1 && a.src_tab != nullptr) {
              const long long fc = (flat[u] - tid * 8) / kDChunk;
              int lo = 0, hi = a.src_n - 1;
              while (lo < hi) {   // uniform across the CTA: broadcast loads out of L1
                const int mid = (lo + hi + 1) >> 1;
                if (__ldg(a.src_tab + 3 * mid) <= fc) lo = mid; else hi = mid - 1;
              }
              const char* base = reinterpret_cast<const char*>(__ldg(a.src_tab + 3 * lo + 1));
              const long long e = (fc - __ldg(a.src_tab + 3 * lo)) * kDChunk + tid * 8;
              const bool ok = base != nullptr && e + 8 <= __ldg(a.src_tab + 3 * lo + 2);
#pragma unroll
              for (int q = 0; q < GV; q++) raw[u][0][q] = ok ? ld_peer16(base + e * sizeof(TG) + q * 16) : make_uint4(0, 0, 0, 0);
            } else {
#pragma unroll
              for (int p = 0; p < DP; p++) {
                if (p < D) {
#pragma unroll
                  for (int q = 0; q < GV; q++)
                    raw[u][p][q] = ld_peer16(reinterpret_cast<const char*>(a.grads.p[p]) + flat[u] * sizeof(TG) + q * 16);
                }
              }
            }
          }
        }
      }
      // ---- phase 2: per chunk: reduce, (Adam), store / push
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (!on[u]) continue;
        float g[8];
        if (MODE == MODE_ADAM) {
#pragma unroll
          for (int i = 0; i < 8; i++) g[i] = red_in[u][i];
        } else if (kReads) {
#pragma unroll
          for (int i = 0; i < 8; i++) g[i] = 0.f;
#pragma unroll
          for (int p = 0; p < NP; p++) {  // fixed summation order => bitwise reproducible
            if (NVLS || p < D) {
              float f[8];
              unpack8<TG>(raw[u][p], f);
#pragma unroll
              for (int i = 0; i < 8; i++) g[i] += f[i];
            }
          }
#pragma unroll
          for (int i = 0; i < 8; i++) { g[i] *= a.pre_scale; nsq += g[i] * g[i]; }
        }
        if (MODE == MODE_RS) {
          st8(a.reduced + local[u], g);
        } else {
          float p[8];
          bool use_rem = false;
          if constexpr (sizeof(TP) == 2) use_rem = a.rem != nullptr;
          if (use_rem) {   // fp32 master = bf16 parameter bits (this rank's copy of its own shard) : int16 remainder
            const uint4 hraw = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a.params.p[rank]) + flat[u] * 2);
            const uint4 lraw = *reinterpret_cast<const uint4*>(a.rem + local[u]);
            const unsigned short* hh = reinterpret_cast<const unsigned short*>(&hraw);
            const short* ll = reinterpret_cast<const short*>(&lraw);
#pragma unroll
            for (int i = 0; i < 8; i++) {
              int hi = (int)(short)hh[i];
              const int lo = (int)ll[i];
              if (lo < 0) hi -= 1;   // undo the round-to-nearest carry of the split
              p[i] = __uint_as_float(((unsigned)(hi & 0xffff) << 16) | (unsigned)(lo & 0xffff));
            }
          } else if (CS) ld8_cs(a.p + local[u], p); else ld8(a.p + local[u], p);
          if (kUpdates) {
            float m[8], v[8];
            if (CS) { ld8_cs(a.m + local[u], m); ld8_cs(a.v + local[u], v); } else { ld8(a.m + local[u], m); ld8(a.v + local[u], v); }
            adam8(p, m, v, g, h);
            if (!use_rem) { if (CS) st8_cs(a.p + local[u], p); else st8(a.p + local[u], p); }
            if (CS) { st8_cs(a.m + local[u], m); st8_cs(a.v + local[u], v); } else { st8(a.m + local[u], m); st8(a.v + local[u], v); }
          }
          uint4 out[PV];
          if (use_rem) {   // split the updated master: signed low half stays local, the (carry-adjusted) high half IS the bf16 parameter
            uint4 lo8;
            short* lo = reinterpret_cast<short*>(&lo8);
            unsigned short* e = reinterpret_cast<unsigned short*>(out);
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const unsigned uu = __float_as_uint(p[i]);
              const int nlo = (int)(short)(uu & 0xffffu);
              int nhi = (int)(short)(uu >> 16);
              if (nlo < 0) nhi += 1;
              lo[i] = (short)nlo; e[i] = (unsigned short)nhi;
            }
            if (kUpdates) *reinterpret_cast<uint4*>(a.rem + local[u]) = lo8;
          } else {
            TP* e = reinterpret_cast<TP*>(out);
#pragma unroll
            for (int i = 0; i < 8; i++) e[i] = from_f<TP>(p[i]);
          }
          if (NVLS) {
#pragma unroll
            for (int q = 0; q < PV; q++) multimem_st16(reinterpret_cast<char*>(a.mc_params) + flat[u] * sizeof(TP) + q * 16, out[q]);
          } else {
#pragma unroll
            for (int p2 = 0; p2 < DP; p2++) {
              if (p2 < D) {
#pragma unroll
                for (int q = 0; q < PV; q++) st_peer16(reinterpret_cast<char*>(a.params.p[p2]) + flat[u] * sizeof(TP) + q * 16, out[q]);
              }
            }
          }
        }
      }
      if (MODE != MODE_RS && a.bucket_ctr != nullptr) {
        // this CTA's pushes of the iteration are issued; thread 0 orders them (cumulatively, through the barrier) before the counters
        __syncthreads();
        if (tid == 0) {
          __threadfence_system();
#pragma unroll
          for (int u = 0; u < U; u++) {
            if (!on[u]) continue;
            const int b = (int)((cb + (long long)u * gridDim.x) / cpb);
            if (atomicAdd(a.bucket_ctr + b, 1u) + 1u == (unsigned)cpb) {
              a.bucket_ctr[b] = 0u;
              __threadfence_system();
              for (int r = 0; r < D; r++)
                st_release_sys(reinterpret_cast<uint32_t*>(a.ready.p[r]) + (size_t)b * kMaxPeers + rank, a.ready_epoch);
            }
          }
        }
      }
    }
  } else if (a.bucket_ctr != nullptr && blockIdx.x == 0 && tid == 0) {
    // skipped step (overflow): the parameters are unchanged and therefore already "ready"
    for (int b = a.bucket_begin; b < a.bucket_end; b++)
      for (int r = 0; r < D; r++) st_release_sys(reinterpret_cast<uint32_t*>(a.ready.p[r]) + (size_t)b * kMaxPeers + rank, a.ready_epoch);
  }

  // ---- epilogue: per-CTA norm partial, then the last CTA to finish closes the collective
  if (kReads) {
    const float s = block_sum(nsq, red);
    if (tid == 0) a.norm_partials[blockIdx.x] = s;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned int ticket = atomicAdd(a.done_ctr, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (kReads) {
    float s = 0.f;
    for (int k = tid; k < (int)gridDim.x; k += kDThreads) s += __ldcg(a.norm_partials + k);
    s = block_sum(s, red);
    if (tid == 0) a.norm_out[0] = s;
    if (D > 1) {
      // publish this shard's sum of squares into slot [norm_slot][rank] of every pad
      if (tid < D) {
        float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint32_t*>(a.sig.pads.p[tid]) + kPadChannels * kMaxPeers);
        st_relaxed_sys_f32(scratch + a.norm_slot * kMaxPeers + rank, s);
      }
      __threadfence_system();
      __syncthreads();
    } else if (tid == 0) {
      a.norm_out[1] = s;
    }
  }
  if (D > 1) {
    signal_all_e(a.sig, epoch, a.chan_end, tid);
    wait_all_e(a.sig, epoch, a.chan_end, tid);
    __syncthreads();
    if (kReads && tid == 0) {
      const float* scratch = reinterpret_cast<const float*>(reinterpret_cast<const uint32_t*>(a.sig.pads.p[rank]) + kPadChannels * kMaxPeers);
      float tot = 0.f;
      for (int r = 0; r < D; r++) tot += ld_relaxed_sys_f32(scratch + a.norm_slot * kMaxPeers + r);
      a.norm_out[1] = tot;
    }
  }
  if (tid == 0) {
    *a.done_ctr = 0u;
    if (a.epoch_ctr) *a.epoch_ctr = epoch;
  }
}

template <typename TG, typename TP, int MODE>
int dist_launch_mode(const DistArgs& a, int nvls, int grid, cudaStream_t st) {
  const int D = a.sig.world;
#define DGO(N, DPV, UV) dist_step_kernel<TG, TP, MODE, N, DPV, UV><<<grid, kDThreads, 0, st>>>(a)
  if (nvls) DGO(true, 1, 8);
  else if (D <= 1) {
    // tuning knob for the single-GPU instantiation (APEX_B200_DIST_W1: 0 = default); see DESIGN.md for the A/B table
    static const int variant = getenv("APEX_B200_DIST_W1") ? atoi(getenv("APEX_B200_DIST_W1")) : 0;
#define DGW(UV, MBV, CSV) dist_step_kernel<TG, TP, MODE, false, 1, UV, MBV, CSV><<<grid, kDThreads, 0, st>>>(a)
    switch (variant) {   // A/B on B200 (gpurun_out/w1_variants.jsonl): U = 2 with 4 resident CTAs per SM and a 6-per-SM grid is the fastest
      case 1: DGW(4, 2, false); break;
      case 2: DGW(2, 3, false); break;
      case 3: DGW(2, 3, true); break;
      case 4: DGW(4, 2, true); break;
      case 5: DGW(1, 4, true); break;
      default: DGW(2, 4, false); break;
    }
#undef DGW
  }
  else if (D == 2) DGO(false, 2, 4);
  else if (D <= 4) DGO(false, 4, 2);
  else DGO(false, 8, 2);
#undef DGO
  AB_CHECK_LAUNCH();
  return 0;
}

template <typename TG, typename TP>
int dist_launch(const DistArgs& a, int mode, int nvls, int grid, cudaStream_t st) {
  if (mode == MODE_FUSED) return dist_launch_mode<TG, TP, MODE_FUSED>(a, nvls, grid, st);
  if (mode == MODE_RS) return dist_launch_mode<TG, TP, MODE_RS>(a, nvls, grid, st);
  if (mode == MODE_PUSH) return dist_launch_mode<TG, TP, MODE_PUSH>(a, nvls, grid, st);
  return dist_launch_mode<TG, TP, MODE_ADAM>(a, nvls, grid, st);
}

}  // namespace ab

using namespace ab;

// grads/params/pads: arrays of `world` pointers (this process's mappings of every rank's buffers).
AB_API int ab_dist_adam_step(int mode, int nvls, const uint64_t* grads, const uint64_t* params, const uint64_t* pads,
                             uint64_t mc_grads, uint64_t mc_params, float* p, short* rem, float* m, float* v, float* reduced,
                             long long bucket_elems, int shard_elems, int bucket_begin, int bucket_end, int lay_rank, int rank, int world,
                             unsigned int epoch, unsigned int* epoch_ctr, int chan_start, int chan_end, int norm_slot, unsigned int* done_ctr,
                             float* norm_partials, float* norm_out, const float* grad_scale, float pre_scale, float lr, float beta1,
                             float beta2, float eps, int step, int adam_mode, int bias_correction, float decay, const int* noop,
                             const float* lr_ptr, const int* step_ptr, const uint64_t* ready, unsigned int* bucket_ctr,
                             unsigned int ready_epoch, int skip_start, const long long* src_tab, int src_n, int dt_g, int dt_p, int grid,
                             cudaStream_t st) {
  if (world < 1 || world > kMaxPeers) return -3;
  if (shard_elems % kDChunk != 0) return -4;
  DistArgs a;
  for (int i = 0; i < kMaxPeers; i++) {
    a.grads.p[i] = i < world && grads ? (void*)grads[i] : nullptr;
    a.params.p[i] = i < world && params ? (void*)params[i] : nullptr;
    a.sig.pads.p[i] = i < world && pads ? (void*)pads[i] : nullptr;
    a.ready.p[i] = i < world && ready ? (void*)ready[i] : nullptr;
  }
  a.bucket_ctr = ready ? bucket_ctr : nullptr; a.ready_epoch = ready_epoch; a.skip_start = skip_start;
  a.src_tab = (world == 1 && src_n > 0) ? src_tab : nullptr; a.src_n = src_n;
  a.mc_grads = (const void*)mc_grads; a.mc_params = (void*)mc_params;
  a.p = p; a.rem = (p == nullptr && dt_p == kBF16) ? rem : nullptr; a.m = m; a.v = v; a.reduced = reduced;
  if (p == nullptr && a.rem == nullptr) return -5;
  a.bucket_elems = bucket_elems; a.shard_elems = shard_elems; a.bucket_begin = bucket_begin; a.bucket_end = bucket_end;
  a.sig.rank = rank; a.sig.world = world; a.sig.epoch = epoch; a.epoch_ctr = world > 1 ? epoch_ctr : nullptr; a.lay_rank = lay_rank;
  a.chan_start = chan_start; a.chan_end = chan_end; a.norm_slot = norm_slot;
  a.done_ctr = done_ctr; a.norm_partials = norm_partials; a.norm_out = norm_out;
  a.grad_scale = grad_scale; a.pre_scale = pre_scale;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.decay = decay; a.mode = adam_mode;
  a.bc1 = 1.f; a.bc2 = 1.f;
  if (bias_correction) {
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  }
  a.noop = noop; a.lr_ptr = lr_ptr; a.step_ptr = step_ptr; a.bias_correction = bias_correction;
  if (grid <= 0) grid = kNumSMs * 2;
  const long long chunks = (long long)(bucket_end - bucket_begin) * (shard_elems / kDChunk);
  if (chunks <= 0) return 0;
  if (chunks < grid) grid = (int)chunks;
  const bool use_nvls = nvls && world > 1 && ((mode == MODE_ADAM || mode == MODE_PUSH) || mc_grads) && ((mode == MODE_RS) || mc_params);
#define DPAIR(TG, TP) return dist_launch<TG, TP>(a, mode, use_nvls ? 1 : 0, grid, st)
  if (dt_g == kBF16 && dt_p == kBF16) DPAIR(bf16, bf16);
  if (dt_g == kF16 && dt_p == kF16) DPAIR(f16, f16);
  if (dt_g == kF32 && dt_p == kF32) DPAIR(float, float);
  if (dt_g == kF32 && dt_p == kBF16) DPAIR(float, bf16);
  return -1;
}

// Plain cross-rank barrier kernel on a channel (used by tests and by buffer (re)initialisation).
__global__ void symm_barrier_kernel(Signal s, int channel) {
  signal_all(s, channel, threadIdx.x);
  wait_all(s, channel, threadIdx.x);
}
AB_API int ab_symm_barrier(const uint64_t* pads, int rank, int world, unsigned int epoch, int channel, cudaStream_t st) {
  Signal s;
  for (int i = 0; i < kMaxPeers; i++) s.pads.p[i] = i < world ? (void*)pads[i] : nullptr;
  s.rank = rank; s.world = world; s.epoch = epoch;
  symm_barrier_kernel<<<1, 32, 0, st>>>(s, channel);
  AB_CHECK_LAUNCH();
  return 0;
}

// Start barrier of a collective as its own 1-warp kernel: while this rank waits for its peers (which may still be busy with their backward
// pass) only one warp is parked on the GPU instead of a whole grid of spinning CTAs; the collective kernel launched behind it on the same
// stream skips its own start barrier. Uses epoch = *epoch_ctr + 1 like the kernel that follows (which is the one that stores it back).
__global__ void symm_gate_kernel(Signal s, const uint32_t* epoch_ctr, int channel) {
  s.epoch = *reinterpret_cast<const volatile uint32_t*>(epoch_ctr) + 1u;
  signal_all(s, channel, threadIdx.x);
  wait_all(s, channel, threadIdx.x);
}
AB_API int ab_symm_gate(const uint64_t* pads, int rank, int world, const unsigned int* epoch_ctr, int channel, cudaStream_t st) {
  if (world <= 1) return 0;
  Signal s;
  for (int i = 0; i < kMaxPeers; i++) s.pads.p[i] = i < world ? (void*)pads[i] : nullptr;
  s.rank = rank; s.world = world; s.epoch = 0;
  symm_gate_kernel<<<1, 32, 0, st>>>(s, epoch_ctr, channel);
  AB_CHECK_LAUNCH();
  return 0;
}

AB_API int ab_symm_pad_words() { return kPadWords; }
