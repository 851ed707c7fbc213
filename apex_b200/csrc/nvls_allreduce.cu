// EXPERIMENTAL (not part of the default build, not yet validated on hardware): one-shot NVSwitch all-reduce of a symmetric buffer.
// Every rank owns 1/D of the buffer: v = multimem.ld_reduce(mc + i) pulls the sum of all D copies through the switch,
// multimem.st(mc + i, v) broadcasts it back into every copy. Two epoch barriers (start: everybody's data is in place; end:
// everybody's slice has been written everywhere). The grid is deliberately small (<= 32 CTAs): this kernel spins on its peers and
// is meant to run NEXT TO compute kernels (DDP bucket all-reduce during backward) -- see DESIGN.md section 7.
// Spec: the NCCL all-reduce the reference's DistributedDataParallel issues per gradient bucket.
#include "symm_device.cuh"

namespace ab {

template <typename T>
__global__ void __launch_bounds__(512) nvls_allreduce_kernel(char* mc, long long n_elems, Signal sig, int chan_start, int chan_end,
                                                             unsigned int* ticket, float post_scale) {
  constexpr int V = 16 / sizeof(T);
  // start: my copy is complete (stream order) -> tell everyone; wait until everyone said the same
  if (blockIdx.x == 0) { __threadfence_system(); signal_all(sig, chan_start, threadIdx.x); }
  wait_all(sig, chan_start, threadIdx.x);
  __syncthreads();
  const long long nvec = n_elems / V;                                 // host guarantees n_elems % (V * world) == 0
  const long long per_rank = nvec / sig.world;
  const long long v0 = per_rank * sig.rank;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_rank; i += (long long)gridDim.x * blockDim.x) {
    char* p = mc + (v0 + i) * 16;
    uint4 r = multimem_ld_reduce16<T>(p);
    if (post_scale != 1.f) {
      T* e = reinterpret_cast<T*>(&r);
#pragma unroll
      for (int j = 0; j < V; j++) e[j] = from_f<T>(to_f<T>(e[j]) * post_scale);
    }
    multimem_st16(p, r);
  }
  // end: the last CTA to finish publishes "my slice is written everywhere" and waits for the same from all peers
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    signal_all(sig, chan_end, threadIdx.x);
    wait_all(sig, chan_end, threadIdx.x);
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace ab

using namespace ab;

// mc: multicast address of the symmetric staging buffer (SymmetricMemory.mc_ptr + offset); n_elems multiple of (16 / esize) * world.
AB_API int ab_nvls_allreduce(void* mc, long long n_elems, const uint64_t* pads, int rank, int world, unsigned int epoch, int chan_start,
                             int chan_end, void* ticket, float post_scale, int ctas, int dt, cudaStream_t st) {
  if (n_elems <= 0 || world <= 1) return 0;
  Signal s;
  for (int i = 0; i < kMaxPeers; i++) s.pads.p[i] = i < world ? (void*)pads[i] : nullptr;
  s.rank = rank; s.world = world; s.epoch = epoch;
  const int esz = dt == kF32 ? 4 : 2;
  if (n_elems % ((16 / esz) * world) != 0 || ((uintptr_t)mc % 16) != 0) return -3;
  if (ctas <= 0 || ctas > 32) ctas = 16;
  unsigned int* tk = reinterpret_cast<unsigned int*>(ticket);
  if (dt == kF32) nvls_allreduce_kernel<float><<<ctas, 512, 0, st>>>((char*)mc, n_elems, s, chan_start, chan_end, tk, post_scale);
  else if (dt == kBF16) nvls_allreduce_kernel<bf16><<<ctas, 512, 0, st>>>((char*)mc, n_elems, s, chan_start, chan_end, tk, post_scale);
  else if (dt == kF16) nvls_allreduce_kernel<f16><<<ctas, 512, 0, st>>>((char*)mc, n_elems, s, chan_start, chan_end, tk, post_scale);
  else return -2;
  return (int)cudaGetLastError();
}
