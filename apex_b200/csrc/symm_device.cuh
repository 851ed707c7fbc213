// Device-side primitives for in-kernel collectives over the symmetric heap (NVLink 5 P2P and NVSwitch multicast).
// Conceptual ancestors in the reference: peer_memory's flag-in-flit protocol (peer_memory_cuda.cu:146-253) and groupbn's
// magic-number butterfly (nhwc_batch_norm_kernel.h:358-460). Here: monotonically increasing epochs in per-peer signal
// slots, release/acquire at .sys scope, no resets, graph-replayable when the epoch lives in device memory.
#pragma once
#include "common.cuh"
#include <cstdio>

namespace ab {

constexpr int kMaxPeers = 8;  // one NVSwitch domain (HGX B200)

struct PeerPtrs { void* p[kMaxPeers]; };

// Signal pad (uint32 words) layout per rank: [channel][kMaxPeers]; a channel is one barrier stream.
constexpr int kPadChannels = 64;
constexpr int kPadWords = kPadChannels * kMaxPeers + 1024;  // + scratch floats for tiny payloads (norms, stats)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f32(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// 16-byte peer load that does not allocate in L1 (peer lines are not coherent with the producer's later writes)
__device__ __forceinline__ uint4 ld_peer16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_peer16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// NVSwitch in-fabric reduction: one load returns the sum over every rank's copy.
template <typename T> __device__ __forceinline__ uint4 multimem_ld_reduce16(const void* mc);
template <> __device__ __forceinline__ uint4 multimem_ld_reduce16<bf16>(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
template <> __device__ __forceinline__ uint4 multimem_ld_reduce16<f16>(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
template <> __device__ __forceinline__ uint4 multimem_ld_reduce16<float>(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
  return r;
}
// NVSwitch broadcast store: one store lands in every rank's copy.
__device__ __forceinline__ void multimem_st16(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_red_add_f32(float* mc, float v) {
  asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

struct Signal {
  PeerPtrs pads;   // pads.p[r] = rank r's signal pad as mapped in THIS process
  int rank, world;
  uint32_t epoch;  // value that marks "this use"; strictly increasing across uses of a channel
};

// Thread `t < world` of the calling warp/CTA tells peer t "rank `rank` reached `epoch` on `channel`".
// Everything this CTA wrote before (after a __syncthreads + __threadfence_system by the caller) is visible first.
__device__ __forceinline__ void signal_all(const Signal& s, int channel, int t) {
  if (t < s.world) {
    uint32_t* slot = reinterpret_cast<uint32_t*>(s.pads.p[t]) + channel * kMaxPeers + s.rank;
    st_release_sys(slot, s.epoch);
  }
}
// Thread `t < world` spins until peer t has signalled `epoch` on `channel`. Bounded: traps after ~10 s instead of hanging.
__device__ __forceinline__ void wait_all(const Signal& s, int channel, int t) {
  if (t < s.world) {
    const uint32_t* slot = reinterpret_cast<const uint32_t*>(s.pads.p[s.rank]) + channel * kMaxPeers + t;
    long long t0 = clock64();
    while ((int)(ld_acquire_sys(slot) - s.epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) { printf("apex_b200: peer %d never signalled channel %d (epoch %u)\n", t, channel, s.epoch); __trap(); }
      __nanosleep(64);
    }
  }
}

// Variants that take the epoch separately: kernels that derive it from a device counter keep their argument struct read-only (a
// modified __grid_constant__ / by-value parameter is copied to the stack and every later access goes through local memory).
__device__ __forceinline__ void signal_all_e(const Signal& s, uint32_t epoch, int channel, int t) {
  if (t < s.world) {
    uint32_t* slot = reinterpret_cast<uint32_t*>(s.pads.p[t]) + channel * kMaxPeers + s.rank;
    st_release_sys(slot, epoch);
  }
}
__device__ __forceinline__ void wait_all_e(const Signal& s, uint32_t epoch, int channel, int t) {
  if (t < s.world) {
    const uint32_t* slot = reinterpret_cast<const uint32_t*>(s.pads.p[s.rank]) + channel * kMaxPeers + t;
    long long t0 = clock64();
    while ((int)(ld_acquire_sys(slot) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) { printf("apex_b200: peer %d never signalled channel %d (epoch %u)\n", t, channel, epoch); __trap(); }
      __nanosleep(64);
    }
  }
}

}  // namespace ab
