// Multi-tensor engine v2 (B200): the tensor table lives in DEVICE memory (built once per parameter set and cached by the
// host runtime), one persistent launch walks every chunk of every tensor. Replaces the reference's by-value 4 KB kernel
// argument struct + one launch per <=110/64/48/36 tensors (reference: csrc/multi_tensor_apply.cuh:13-103).
#pragma once
#include <cstdlib>
#include "common.cuh"
#include <tuple>

namespace ab {

// Device-resident table. Arena layout (host packs it, see table_bytes()/binding.cpp):
//   [ptrs: depth*n void*][numel: n int64][chunk_prefix: (n+1) int32]
struct MTTable {
  void* const* ptrs;        // ptrs[d * n + t]
  const int64_t* numel;     // elements of tensor t
  const int* chunk_prefix;  // chunk_prefix[t] = first global chunk id of tensor t; [n] = total
  int n;
  int depth;
  int total_chunks;
  int chunk;  // elements per chunk (multiple of 32)
};

__host__ __device__ inline MTTable make_table(void* arena, int n, int depth, int total_chunks, int chunk) {
  MTTable t;
  char* b = reinterpret_cast<char*>(arena);
  t.ptrs = reinterpret_cast<void* const*>(b);
  t.numel = reinterpret_cast<const int64_t*>(b + sizeof(void*) * (size_t)depth * n);
  t.chunk_prefix = reinterpret_cast<const int*>(b + sizeof(void*) * (size_t)depth * n + sizeof(int64_t) * (size_t)n);
  t.n = n; t.depth = depth; t.total_chunks = total_chunks; t.chunk = chunk;
  return t;
}

// upper_bound over chunk_prefix: tensor that owns global chunk `cid`.
__device__ __forceinline__ int find_tensor(const int* __restrict__ prefix, int n, int cid) {
  int lo = 0, hi = n;  // invariant: prefix[lo] <= cid < prefix[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= cid) lo = mid; else hi = mid;
  }
  return lo;
}

constexpr int kMTThreads = 256;
constexpr int kMTUnroll = 2;

// Op concept:
//   static constexpr unsigned kRead, kWrite;      bit i => slot i is loaded / stored
//   static constexpr int kAcc;                    number of per-thread float accumulators (0..2)
//   __device__ bool skip() const;                 whole-kernel early exit (noop flag)
//   struct Ctx; __device__ Ctx begin(int t) const;   per-(tensor) scalars
//   template<int D,int V> __device__ void apply(float (&r)[D][V], const Ctx&, float (&acc)[2]) const;
//   __device__ void end(const Ctx&, int t, int cid, float (&acc)[2], float* red) const;   (block-uniform call)
template <int V, class Op, class... Ts>
struct MTImpl {
  static constexpr int D = sizeof...(Ts);
  using Tup = std::tuple<Ts...>;

  template <int I>
  static __device__ __forceinline__ void load_all(float (&r)[D][V], void* const (&p)[D], int64_t off, int nvalid, bool vec) {
    if constexpr (I < D) {
      using T = typename std::tuple_element<I, Tup>::type;
      if constexpr ((Op::kRead >> I) & 1u) {
        const T* q = reinterpret_cast<const T*>(p[I]) + off;
        if (vec) {
          load_vec<T, V>(r[I], q);
        } else {
#pragma unroll
          for (int j = 0; j < V; j++) r[I][j] = (j < nvalid) ? to_f<T>(q[j]) : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < V; j++) r[I][j] = 0.f;
      }
      load_all<I + 1>(r, p, off, nvalid, vec);
    }
  }
  template <int I>
  static __device__ __forceinline__ void store_all(const float (&r)[D][V], void* const (&p)[D], int64_t off, int nvalid, bool vec) {
    if constexpr (I < D) {
      using T = typename std::tuple_element<I, Tup>::type;
      if constexpr ((Op::kWrite >> I) & 1u) {
        T* q = reinterpret_cast<T*>(p[I]) + off;
        if (vec) {
          store_vec<T, V>(q, r[I]);
        } else {
#pragma unroll
          for (int j = 0; j < V; j++) if (j < nvalid) q[j] = from_f<T>(r[I][j]);
        }
      }
      store_all<I + 1>(r, p, off, nvalid, vec);
    }
  }
};

template <int V, class Op, class... Ts>
__global__ void __launch_bounds__(kMTThreads, 3) mt_kernel(MTTable tb, Op op) {
  using Impl = MTImpl<V, Op, Ts...>;
  constexpr int D = Impl::D;
  __shared__ float red[40];
  if (op.skip()) return;
  for (int cid = blockIdx.x; cid < tb.total_chunks; cid += gridDim.x) {
    const int t = find_tensor(tb.chunk_prefix, tb.n, cid);
    const int64_t n_t = tb.numel[t];
    const int64_t base = (int64_t)(cid - __ldg(tb.chunk_prefix + t)) * tb.chunk;
    const int len = (int)min((int64_t)tb.chunk, n_t - base);
    void* p[D];
    bool al = true;
#pragma unroll
    for (int d = 0; d < D; d++) {
      p[d] = tb.ptrs[(size_t)d * tb.n + t];
      if (((Op::kRead | Op::kWrite) >> d) & 1u) al = al && aligned16(p[d]);
    }
    typename Op::Ctx ctx = op.begin(t);
    float acc[2] = {0.f, 0.f};
    for (int i0 = threadIdx.x * V; i0 < len; i0 += kMTThreads * V * kMTUnroll) {
      float r[kMTUnroll][D][V];
      int nv[kMTUnroll];
#pragma unroll
      for (int u = 0; u < kMTUnroll; u++) {
        const int i = i0 + u * kMTThreads * V;
        nv[u] = min(V, len - i);  // <=0 when past the end
        if (nv[u] > 0) Impl::template load_all<0>(r[u], p, base + i, nv[u], al && nv[u] == V);
      }
#pragma unroll
      for (int u = 0; u < kMTUnroll; u++) {
        if (nv[u] > 0) {
          if (nv[u] < V) {  // neutralise padded lanes for reductions: they were zero-filled at load
          }
          op.template apply<D, V>(r[u], ctx, acc, nv[u]);
          Impl::template store_all<0>(r[u], p, base + i0 + u * kMTThreads * V, nv[u], al && nv[u] == V);
        }
      }
    }
    if constexpr (Op::kAcc > 0) op.end(ctx, t, cid, acc, red);
  }
}

inline int mt_grid(int total_chunks) {
  // CTAs per SM of the grid. 3 are resident; a finer grid lets the block scheduler even out the two dies / uneven tensors: measured
  // on 10k tensors (gpurun_out/mt_grid.jsonl) 12 per SM is 3 % (Adam) to 6 % (SGD) faster than 3, while 4 and 8 (partial waves) are slower.
  static const int mult = getenv("APEX_B200_MT_GRID_MULT") ? atoi(getenv("APEX_B200_MT_GRID_MULT")) : 12;
  int g = kNumSMs * (mult > 0 ? mult : 12);
  return total_chunks < g ? (total_chunks > 0 ? total_chunks : 1) : g;
}

template <int V, class Op, class... Ts>
inline int mt_launch(const MTTable& tb, const Op& op, cudaStream_t st) {
  if (tb.total_chunks <= 0) return 0;
  mt_kernel<V, Op, Ts...><<<mt_grid(tb.total_chunks), kMTThreads, 0, st>>>(tb, op);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

}  // namespace ab
