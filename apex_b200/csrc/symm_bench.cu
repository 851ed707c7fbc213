// Bandwidth probes for the symmetric heap -- the "8-GPU P2P bandwidth, NVLS
// multimem bandwidth, barrier latency" unit measurements of SURVEY.md section 7.2 step 5. Each probe moves `bytes` once per launch with
// 16-byte accesses and `unroll` independent accesses in flight per thread (bytes in flight per SM, not threads, set NVLink bandwidth).
//   op 0  peer read    local <- peer   (ld.global.L1::no_allocate from the peer mapping)
//   op 1  peer write   peer  <- local  (st to the peer mapping)
//   op 2  multicast ld_reduce (multimem.ld_reduce.add: the switch returns the sum of every rank's copy; result discarded into a checksum)
//   op 3  multicast st        (multimem.st: one store lands in every rank's copy)
//   op 4  multicast ld_reduce of src -> multimem.st into dst (both directions of every link busy at once: the all-reduce / fused ZeRO pattern)
// Driver: benchmarks/bench_symm.py (torchrun, CUDA events, max over ranks).
#include "symm_device.cuh"

namespace ab {

template <int U>
__global__ void __launch_bounds__(512) symm_bench_kernel(int op, const char* src, char* dst, long long nvec, float* sink) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (; i + (U - 1) * stride < nvec; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long k = i + u * stride;
      if (op == 0) v[u] = ld_peer16(src + k * 16);
      else if (op == 2 || op == 4) v[u] = multimem_ld_reduce16<float>(src + k * 16);
      else v[u] = *reinterpret_cast<const uint4*>(src + k * 16);   // local read feeding a remote / multicast store
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long k = i + u * stride;
      if (op == 0) *reinterpret_cast<uint4*>(dst + k * 16) = v[u];
      else if (op == 1) st_peer16(dst + k * 16, v[u]);
      else if (op == 3 || op == 4) multimem_st16(dst + k * 16, v[u]);
      else { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
  }
  for (; i < nvec; i += stride) {  // tail
    uint4 v = (op == 0) ? ld_peer16(src + i * 16) : (op == 2 || op == 4) ? multimem_ld_reduce16<float>(src + i * 16) : *reinterpret_cast<const uint4*>(src + i * 16);
    if (op == 0) *reinterpret_cast<uint4*>(dst + i * 16) = v;
    else if (op == 1) st_peer16(dst + i * 16, v);
    else if (op == 3 || op == 4) multimem_st16(dst + i * 16, v);
    else { acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
  }
  if (op == 2 && (acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x5eedf00du) *sink = 1.f;  // keeps the reduce loads alive
}

}  // namespace ab

using namespace ab;

// src / dst: device addresses valid in THIS process (local buffer, peer mapping or multicast mapping, as the op requires); bytes % 16 == 0.
AB_API int ab_symm_bench(int op, const void* src, void* dst, long long bytes, int ctas, int unroll, float* sink, cudaStream_t st) {
  if (op < 0 || op > 4 || bytes <= 0 || bytes % 16) return -3;
  if (ctas <= 0) ctas = 148 * 2;
  const long long nvec = bytes / 16;
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  switch (unroll) {
    case 1: symm_bench_kernel<1><<<ctas, 512, 0, st>>>(op, s, d, nvec, sink); break;
    case 2: symm_bench_kernel<2><<<ctas, 512, 0, st>>>(op, s, d, nvec, sink); break;
    case 4: symm_bench_kernel<4><<<ctas, 512, 0, st>>>(op, s, d, nvec, sink); break;
    case 8: symm_bench_kernel<8><<<ctas, 512, 0, st>>>(op, s, d, nvec, sink); break;
    default: return -4;
  }
  return (int)cudaGetLastError();
}
