// 1-D halo exchange between neighbouring GPUs in ONE kernel: pack the two outgoing halo slabs of a strided 4-D tensor straight
// into the neighbours' transfer buffers (16-byte P2P stores over NVLink), release-signal the neighbours, acquire-wait for theirs,
// and unpack the incoming slabs into the halo rows — no NCCL, no host synchronisation, no separate barrier launches.
// Spec: reference apex/contrib/csrc/peer_memory/peer_memory_cuda.cu:146-295,528 (push_pull_halos_1d: flag-in-flit volatile
// stores + cooperative launch). Differences: payload and flag are separate (full 16-byte payload vectors; one epoch word per
// neighbour per exchange), transfer buffers are double-buffered by exchange parity so there is no trailing barrier, and any
// strided layout (NCHW, channels-last, explicit NHWC; split along H or W) goes through the same index decode.
#include "symm_device.cuh"

namespace ab {

struct HaloArgs {
  char* y;                         // storage base of the padded tensor
  int shape[4]; long long stride[4];  // slab extents and tensor strides (elements), innermost last
  long long off[4];                // element offsets of low_out, low_in, high_out, high_in
  long long slab;                  // elements per slab
  char* tx_low; char* tx_high; char* tx_mine;  // [2][slab] transfer buffers (this parity): the low / high neighbour's and mine
  uint32_t* pad_low; uint32_t* pad_high; uint32_t* pad_mine;
  int rank, rank_low, rank_high, has_low, has_high, channel;
  uint32_t epoch;
  unsigned int* ticket;
  // device-resident epoch (graph-replayable): epoch = *epoch_ctr + 1; the transfer buffers are double-buffered by the epoch's parity
  // (tx_* point at parity 0, parity 1 lies `parity_stride` bytes further); the CTA that signals the neighbours stores the epoch back
  uint32_t* epoch_ctr; long long parity_stride;
};

template <typename E>
__device__ __forceinline__ long long slab_index(const HaloArgs& a, long long i, int vec) {
  // i counts vectors of `vec` elements along the innermost dim
  const int s3 = a.shape[3] / vec;
  const int i3 = (int)(i % s3); long long r = i / s3;
  const int i2 = (int)(r % a.shape[2]); r /= a.shape[2];
  const int i1 = (int)(r % a.shape[1]); const int i0 = (int)(r / a.shape[1]);
  return i0 * a.stride[0] + i1 * a.stride[1] + i2 * a.stride[2] + (long long)i3 * vec * a.stride[3];
}

template <typename E, typename VT>
__global__ void __launch_bounds__(256) halo_kernel(const __grid_constant__ HaloArgs a) {
  constexpr int VEC = sizeof(VT) / sizeof(E);
  __shared__ int s_last;
  const uint32_t epoch = a.epoch_ctr ? *reinterpret_cast<volatile uint32_t*>(a.epoch_ctr) + 1u : a.epoch;
  const long long poff = a.epoch_ctr ? (long long)(epoch & 1u) * a.parity_stride : 0;
  char* const tx_low = a.tx_low + poff; char* const tx_high = a.tx_high + poff; char* const tx_mine = a.tx_mine + poff;
  const long long nvec = a.slab / VEC;
  const long long gstride = (long long)gridDim.x * 256;
  const E* y = reinterpret_cast<const E*>(a.y);
  // ---- push: my low-side interior rows become the low neighbour's HIGH input (its slot 1), and symmetrically
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += gstride) {
    const long long o = slab_index<E>(a, i, VEC);
    if (a.has_low) {
      const VT v = *reinterpret_cast<const VT*>(y + a.off[0] + o);
      *reinterpret_cast<VT*>(tx_low + (a.slab + i * VEC) * sizeof(E)) = v;
    }
    if (a.has_high) {
      const VT v = *reinterpret_cast<const VT*>(y + a.off[2] + o);
      *reinterpret_cast<VT*>(tx_high + (i * VEC) * sizeof(E)) = v;
    }
  }
  // ---- the last CTA to finish its pushes tells the neighbours
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence_system();
    if (a.has_low) st_release_sys(a.pad_low + a.channel * kMaxPeers + a.rank, epoch);
    if (a.has_high) st_release_sys(a.pad_high + a.channel * kMaxPeers + a.rank, epoch);
    *a.ticket = 0u;
    if (a.epoch_ctr) *a.epoch_ctr = epoch;   // every CTA has read it (all tickets are in)
  }
  // ---- wait for the neighbours' pushes into my buffer
  if (threadIdx.x < 2) {
    const bool need = threadIdx.x == 0 ? a.has_low : a.has_high;
    if (need) {
      const uint32_t* slot = a.pad_mine + a.channel * kMaxPeers + (threadIdx.x == 0 ? a.rank_low : a.rank_high);
      long long t0 = clock64();
      while ((int)(ld_acquire_sys(slot) - epoch) < 0) {
        if (clock64() - t0 > 20000000000LL) { printf("apex_b200 halo: neighbour never signalled (epoch %u)\n", epoch); __trap(); }
        __nanosleep(64);
      }
    }
  }
  __syncthreads();
  // ---- unpack (zero halos at the ends of the chain)
  E* yw = reinterpret_cast<E*>(a.y);
  VT zero; memset(&zero, 0, sizeof(VT));
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += gstride) {
    const long long o = slab_index<E>(a, i, VEC);
    const VT lo = a.has_low ? __ldcg(reinterpret_cast<const VT*>(tx_mine + (i * VEC) * sizeof(E))) : zero;
    const VT hi = a.has_high ? __ldcg(reinterpret_cast<const VT*>(tx_mine + (a.slab + i * VEC) * sizeof(E))) : zero;
    *reinterpret_cast<VT*>(yw + a.off[1] + o) = lo;
    *reinterpret_cast<VT*>(yw + a.off[3] + o) = hi;
  }
}

}  // namespace ab

using namespace ab;

// meta: 4 slab extents, 4 strides, 4 offsets (low_out, low_in, high_out, high_in) as int64, strides/offsets in elements.
AB_API int ab_halo_exchange_1d(void* y, int esize, const long long* meta, void* tx_low, void* tx_high, void* tx_mine, void* pad_low,
                               void* pad_high, void* pad_mine, int rank, int rank_low, int rank_high, int has_low, int has_high,
                               int channel, unsigned int epoch, unsigned int* epoch_ctr, long long parity_stride, void* ticket, int max_ctas,
                               cudaStream_t s) {
  HaloArgs a;
  a.y = (char*)y;
  a.slab = 1;
  for (int i = 0; i < 4; i++) { a.shape[i] = (int)meta[i]; a.stride[i] = meta[4 + i]; a.off[i] = meta[8 + i]; a.slab *= meta[i]; }
  a.tx_low = (char*)tx_low; a.tx_high = (char*)tx_high; a.tx_mine = (char*)tx_mine;
  a.pad_low = (uint32_t*)pad_low; a.pad_high = (uint32_t*)pad_high; a.pad_mine = (uint32_t*)pad_mine;
  a.rank = rank; a.rank_low = rank_low; a.rank_high = rank_high; a.has_low = has_low; a.has_high = has_high;
  a.channel = channel; a.epoch = epoch; a.ticket = (unsigned int*)ticket; a.epoch_ctr = epoch_ctr; a.parity_stride = parity_stride;
  if (a.slab <= 0) return 0;
  if (esize != 2 && esize != 4) return -2;
  const int V = 16 / esize;
  bool vec = a.stride[3] == 1 && a.shape[3] % V == 0 && ((uintptr_t)y % 16 == 0);
  for (int i = 0; i < 3 && vec; i++) vec = a.stride[i] % V == 0;
  for (int i = 0; i < 4 && vec; i++) vec = a.off[i] % V == 0;
  const long long nvec = vec ? a.slab / V : a.slab;
  long long want = (nvec + 255) / 256;
  const int cap = max_ctas > 0 && max_ctas < kNumSMs ? max_ctas : kNumSMs;  // all CTAs must be co-resident (they spin)
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  if (esize == 2) {
    if (vec) halo_kernel<unsigned short, uint4><<<grid, 256, 0, s>>>(a); else halo_kernel<unsigned short, unsigned short><<<grid, 256, 0, s>>>(a);
  } else {
    if (vec) halo_kernel<unsigned int, uint4><<<grid, 256, 0, s>>>(a); else halo_kernel<unsigned int, unsigned int><<<grid, 256, 0, s>>>(a);
  }
  return (int)cudaGetLastError();
}
