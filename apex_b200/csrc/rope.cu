// Fused rotary position embedding, forward and backward, for the four layouts of the reference extension
// (csrc/megatron/fused_rotary_positional_embedding.h:27-430, .cpp:42-193): sbhd with on-the-fly sin/cos from `freqs`,
// sbhd with cached cos/sin, packed thd (var-len, cu_seqlens) and 2-D image RoPE (height half / width half).
// One persistent kernel: a CTA owns a token (all heads), computes the token's cos/sin ONCE into shared memory and applies them
// to every head; arbitrary element strides on both sides (covers transpose_output).
//   fwd: y[i] = x[i] cos[i] + rot(x)[i] sin[i],  rot(x)[i] = -x[i+r/2] (i < r/2), x[i-r/2] (i >= r/2);  dims >= r pass through
//   bwd: dx[i] = dy[i] cos[i] + rot'(dy sin)[i]
#include "common.cuh"

namespace ab {

enum { ROPE_SBHD = 0, ROPE_THD = 1, ROPE_2D = 2 };

struct RopeArgs {
  const void* x; void* out;
  int mode, is_bwd, cached;
  int n_tokens, s, b, h, d, r;           // r = rotary dims (d2); for 2-D each half of d is rotated with r = d/2
  long long xs_tok0, xs_tok1, xs_h, xs_d;  // sbhd: (stride_s, stride_b); thd: (stride_t, 0); 2d: handled via ih/iw below
  long long os_tok0, os_tok1, os_h, os_d;
  long long xs_b2, xs_ih, xs_iw, os_b2, os_seq;  // 2-D: x strides (b, ih, iw); out strides (b, ih*iw flattened)
  int ih, iw;
  const float* freqs;   // [max_s, r] fp32 angles
  const void* cos0; const void* sin0; const void* cos1; const void* sin1;  // cached tables (dtype dt_cs): [pos, r]
  const int* cu_seqlens; int n_seqs;
  int dt_cs;
};

__device__ __forceinline__ float ld_cs(const void* p, int dt, long long i) {
  if (dt == kF32) return reinterpret_cast<const float*>(p)[i];
  if (dt == kF16) return __half2float(reinterpret_cast<const f16*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}

// VEC: unit element stride on both sides, every stride / offset a multiple of 8 elements, rotary half a multiple of 8: a work item is
// 8 consecutive dims of the first half and the matching 8 of the second half (two 16-byte loads, two 16-byte stores, one integer
// division) — the scalar version (2-byte accesses, a division per pair) was instruction-issue bound at 16 % of HBM (profiles/rope_fwd.md).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) rope_kernel(const __grid_constant__ RopeArgs a) {
  extern __shared__ float sc[];  // cos[nsc], sin[nsc]
  const int nsc = (a.mode == ROPE_2D) ? a.d : a.r;
  float* cs = sc;
  float* sn = sc + nsc;
  const T* x = reinterpret_cast<const T*>(a.x);
  T* out = reinterpret_cast<T*>(a.out);
  for (int tok = blockIdx.x; tok < a.n_tokens; tok += gridDim.x) {
    long long xoff, ooff;
    int pos0 = 0, pos1 = 0;
    if (a.mode == ROPE_SBHD) {
      const int si = tok / a.b, bi = tok - si * a.b;
      pos0 = si; xoff = si * a.xs_tok0 + bi * a.xs_tok1; ooff = si * a.os_tok0 + bi * a.os_tok1;
    } else if (a.mode == ROPE_THD) {
      int lo = 0, hi = a.n_seqs;  // cu_seqlens[lo] <= tok < cu_seqlens[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.cu_seqlens[mid] <= tok) lo = mid; else hi = mid; }
      pos0 = tok - a.cu_seqlens[lo]; xoff = tok * a.xs_tok0; ooff = tok * a.os_tok0;
    } else {
      const int per_img = a.ih * a.iw;
      const int bi = tok / per_img, rem = tok - bi * per_img;
      pos0 = rem / a.iw; pos1 = rem - pos0 * a.iw;
      xoff = bi * a.xs_b2 + pos0 * a.xs_ih + pos1 * a.xs_iw; ooff = bi * a.os_b2 + (long long)rem * a.os_seq;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsc; i += blockDim.x) {
      float c, s;
      if (a.mode == ROPE_2D) {
        const int half = a.d / 2;
        if (i < half) { c = ld_cs(a.cos0, a.dt_cs, (long long)pos0 * half + i); s = ld_cs(a.sin0, a.dt_cs, (long long)pos0 * half + i); }
        else { c = ld_cs(a.cos1, a.dt_cs, (long long)pos1 * half + (i - half)); s = ld_cs(a.sin1, a.dt_cs, (long long)pos1 * half + (i - half)); }
      } else if (a.cached) {
        c = ld_cs(a.cos0, a.dt_cs, (long long)pos0 * a.r + i); s = ld_cs(a.sin0, a.dt_cs, (long long)pos0 * a.r + i);
      } else {
        sincosf(a.freqs[(long long)pos0 * a.r + i], &s, &c);
      }
      cs[i] = c; sn[i] = s;
    }
    __syncthreads();
    const int parts = (a.mode == ROPE_2D) ? 2 : 1;
    const int r = (a.mode == ROPE_2D) ? a.d / 2 : a.r;
    const int hr = r / 2;
    const int pairs_per_head = parts * hr;
    if (VEC) {
      constexpr int V = 16 / sizeof(T);          // elements per 16-byte vector (8 for 16-bit types, 4 for fp32)
      const int vp = hr / V, per_head = parts * vp;
      for (int j = threadIdx.x; j < a.h * per_head; j += blockDim.x) {
        const int head = j / per_head, jj = j - head * per_head;
        const int part = jj / vp, i = (jj - part * vp) * V;
        const int base = part * r;
        const long long xi = xoff + head * a.xs_h + base + i, oi = ooff + head * a.os_h + base + i;
        float v0[V], v1[V], o0[V], o1[V];
        load_vec<T, V>(v0, x + xi);
        load_vec<T, V>(v1, x + xi + hr);
#pragma unroll
        for (int e = 0; e < V; e++) {
          const float c0 = cs[base + i + e], c1 = cs[base + i + e + hr], s0 = sn[base + i + e], s1 = sn[base + i + e + hr];
          if (!a.is_bwd) { o0[e] = v0[e] * c0 - v1[e] * s0; o1[e] = v1[e] * c1 + v0[e] * s1; }
          else { o0[e] = v0[e] * c0 + v1[e] * s1; o1[e] = v1[e] * c1 - v0[e] * s0; }
        }
        store_vec<T, V>(out + oi, o0);
        store_vec<T, V>(out + oi + hr, o1);
      }
      if (a.mode != ROPE_2D && a.d > a.r) {  // pass-through tail, 16 bytes at a time
        const int tv = (a.d - a.r) / V;
        for (int j = threadIdx.x; j < a.h * tv; j += blockDim.x) {
          const int head = j / tv, i = a.r + (j - head * tv) * V;
          *reinterpret_cast<uint4*>(out + ooff + head * a.os_h + i) = *reinterpret_cast<const uint4*>(x + xoff + head * a.xs_h + i);
        }
      }
      continue;
    }
    for (int j = threadIdx.x; j < a.h * pairs_per_head; j += blockDim.x) {
      const int head = j / pairs_per_head, jj = j - head * pairs_per_head;
      const int part = jj / hr, i = jj - part * hr;
      const int base = part * r;  // dim offset of this rotary block inside the head
      const long long xi = xoff + head * a.xs_h + (long long)(base + i) * a.xs_d, xj = xi + (long long)hr * a.xs_d;
      const long long oi = ooff + head * a.os_h + (long long)(base + i) * a.os_d, oj = oi + (long long)hr * a.os_d;
      const float v0 = to_f<T>(x[xi]), v1 = to_f<T>(x[xj]);
      const float c0 = cs[base + i], c1 = cs[base + i + hr], s0 = sn[base + i], s1 = sn[base + i + hr];
      float o0, o1;
      if (!a.is_bwd) { o0 = v0 * c0 - v1 * s0; o1 = v1 * c1 + v0 * s1; }
      else { o0 = v0 * c0 + v1 * s1; o1 = v1 * c1 - v0 * s0; }
      out[oi] = from_f<T>(o0); out[oj] = from_f<T>(o1);
    }
    if (a.mode != ROPE_2D && a.d > a.r) {  // pass-through tail
      const int tail = a.d - a.r;
      for (int j = threadIdx.x; j < a.h * tail; j += blockDim.x) {
        const int head = j / tail, i = a.r + (j - head * tail);
        out[ooff + head * a.os_h + (long long)i * a.os_d] = x[xoff + head * a.xs_h + (long long)i * a.xs_d];
      }
    }
  }
}

}  // namespace ab

using namespace ab;

AB_API int ab_rope(const void* x, void* out, int mode, int is_bwd, int cached, int n_tokens, int s, int b, int h, int d, int r,
                   long long xs0, long long xs1, long long xsh, long long xsd, long long os0, long long os1, long long osh, long long osd,
                   long long xs_b2, long long xs_ih, long long xs_iw, long long os_b2, long long os_seq, int ih, int iw, const float* freqs,
                   const void* cos0, const void* sin0, const void* cos1, const void* sin1, const int* cu_seqlens, int n_seqs, int dt_cs,
                   int dt, cudaStream_t st) {
  if (n_tokens <= 0) return 0;
  RopeArgs a;
  a.x = x; a.out = out; a.mode = mode; a.is_bwd = is_bwd; a.cached = cached; a.n_tokens = n_tokens; a.s = s; a.b = b; a.h = h; a.d = d; a.r = r;
  a.xs_tok0 = xs0; a.xs_tok1 = xs1; a.xs_h = xsh; a.xs_d = xsd; a.os_tok0 = os0; a.os_tok1 = os1; a.os_h = osh; a.os_d = osd;
  a.xs_b2 = xs_b2; a.xs_ih = xs_ih; a.xs_iw = xs_iw; a.os_b2 = os_b2; a.os_seq = os_seq; a.ih = ih; a.iw = iw;
  a.freqs = freqs; a.cos0 = cos0; a.sin0 = sin0; a.cos1 = cos1; a.sin1 = sin1; a.cu_seqlens = cu_seqlens; a.n_seqs = n_seqs; a.dt_cs = dt_cs;
  const int nsc = (mode == ROPE_2D) ? d : r;
  const int grid = n_tokens < kNumSMs * 8 ? n_tokens : kNumSMs * 8;
  const size_t smem = sizeof(float) * 2 * (size_t)nsc;
  const int esz = dt == kF32 ? 4 : 2, V = 16 / esz;
  const int rr = (mode == ROPE_2D) ? d / 2 : r;
  auto m8 = [&](long long v) { return v % V == 0; };
  const bool vec = xsd == 1 && osd == 1 && (rr / 2) % V == 0 && rr % 2 == 0 && (d - r) % V == 0 && m8(xs0) && m8(xs1) && m8(xsh) && m8(os0) &&
                   m8(os1) && m8(osh) && m8(xs_b2) && m8(xs_ih) && m8(xs_iw) && m8(os_b2) && m8(os_seq) && aligned16(x) && aligned16(out);
  if (vec) { AB_DISPATCH_FLOAT3(dt, T, (rope_kernel<T, true><<<grid, 256, smem, st>>>(a))); }
  else { AB_DISPATCH_FLOAT3(dt, T, (rope_kernel<T, false><<<grid, 256, smem, st>>>(a))); }
  AB_CHECK_LAUNCH();
  return 0;
}
