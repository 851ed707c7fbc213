// Fused pointwise tails of the convolution blocks (conv_bias_relu / bottleneck), channels-last:
//   forward :  out = act( (y * scale[c]) + bias[c] + z ) * mask          one pass, in place over the convolution output
//   backward:  g = dout * [out > 0] * mask ;  dz = g ;  dy = g * scale[c] ;  dbias[c] = sum_rows g ;  dscale[c] = sum_rows g * y
//              one pass over dout / out: the ReLU mask, the residual branch gradient, the gradient that feeds the convolution
//              backward and the per-channel reductions come out of the same read.
// Replaces the pointwise parts of the reference's cuDNN-frontend runtime-fused graphs (apex/contrib/csrc/conv_bias_relu/
// conv_bias_relu.cpp: conv+bias+relu forward, drelu+dbias backward graphs :1902-1911; apex/contrib/csrc/bottleneck/bottleneck.cpp:
// conv+scale+bias(+add)+relu forward, drelu-dscale-dbias backward :3558-3594). The convolution itself stays a library call (cuDNN),
// as in the reference.
//
// Tensors are viewed as [rows = N*H*W, C] row-major (channels-last memory). A thread owns VEC consecutive channels and walks rows
// with a grid stride, so per-channel coefficients are loaded once and the reductions are register accumulators; CTAs combine
// through shared memory and one atomicAdd per (CTA, channel) into fp32 outputs.
#include "common.cuh"

namespace ab {

constexpr int kCeThreads = 256;

struct ConvEpiArgs {
  const void* y;        // convolution output (forward input; backward: only for dscale)
  const void* z;        // residual (forward) or null
  const uint8_t* mask;  // [rows, C] 0 / 1 bytes or null
  const float* scale;   // [C] or null
  const float* bias;    // [C] or null
  void* out;            // forward output (may alias y)
  const void* dout;     // backward: incoming gradient
  const void* fout;     // backward: forward output (ReLU mask) or null when relu == 0
  void* dy;             // backward: gradient wrt the convolution output
  void* dz;             // backward: gradient wrt the residual (null: not needed, or aliases dy when scale == null)
  float* dbias;         // [C] fp32, zero on entry, or null
  float* dscale;        // [C] fp32, zero on entry, or null
  long long rows; int C; int relu;
};

template <typename T, int V> __device__ __forceinline__ void ce_load(float (&r)[V], const T* p) {
  if constexpr (V == 1) r[0] = to_f<T>(p[0]); else load_vec<T, V>(r, p);
}
template <typename T, int V> __device__ __forceinline__ void ce_store(T* p, const float (&r)[V]) {
  if constexpr (V == 1) p[0] = from_f<T>(r[0]); else store_vec<T, V>(p, r);
}

// helper: how the CTA's threads are laid over (row, channel-vector): min(cvecs, 256) columns, the rest of the threads stack rows
__host__ __device__ inline int ce_cols(int cvecs) { return cvecs < kCeThreads ? cvecs : kCeThreads; }

template <typename T, int VEC>
__global__ void __launch_bounds__(kCeThreads) conv_epi_fwd(ConvEpiArgs a) {
  const int cvecs = a.C / VEC;
  const int cols = ce_cols(cvecs), rows_per_cta = kCeThreads / cols;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  if (tr >= rows_per_cta) return;
  const T* y = reinterpret_cast<const T*>(a.y);
  const T* z = reinterpret_cast<const T*>(a.z);
  T* out = reinterpret_cast<T*>(a.out);
  for (int cv = tc; cv < cvecs; cv += cols) {
    const int c0 = cv * VEC;
    float sc[VEC], bi[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) { sc[j] = a.scale ? a.scale[c0 + j] : 1.f; bi[j] = a.bias ? a.bias[c0 + j] : 0.f; }
    for (long long r = (long long)blockIdx.x * rows_per_cta + tr; r < a.rows; r += (long long)gridDim.x * rows_per_cta) {
      const long long off = r * a.C + c0;
      float v[VEC], zv[VEC];
      ce_load<T, VEC>(v, y + off);
      if (z) ce_load<T, VEC>(zv, z + off);
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        float o = fmaf(v[j], sc[j], bi[j]);
        if (z) o += zv[j];
        if (a.mask) o *= (float)a.mask[off + j];
        if (a.relu) o = fmaxf(o, 0.f);
        v[j] = o;
      }
      ce_store<T, VEC>(out + off, v);
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kCeThreads) conv_epi_bwd(ConvEpiArgs a) {
  __shared__ float red[2][kCeThreads][VEC > 4 ? 9 : VEC + 1];
  const int cvecs = a.C / VEC;
  const int cols = ce_cols(cvecs), rows_per_cta = kCeThreads / cols;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const bool active = tr < rows_per_cta;
  const T* dout = reinterpret_cast<const T*>(a.dout);
  const T* fout = reinterpret_cast<const T*>(a.fout);
  const T* y = reinterpret_cast<const T*>(a.y);
  T* dy = reinterpret_cast<T*>(a.dy);
  T* dz = reinterpret_cast<T*>(a.dz);
  for (int cv0 = 0; cv0 < cvecs; cv0 += cols) {   // uniform trip count: the CTA-wide reduction below needs every thread
    const int cv = cv0 + tc;
    const bool on = active && cv < cvecs;
    const int c0 = cv * VEC;
    float sc[VEC], sb[VEC], ss[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) { sc[j] = (on && a.scale) ? a.scale[c0 + j] : 1.f; sb[j] = 0.f; ss[j] = 0.f; }
    if (on) {
      for (long long r = (long long)blockIdx.x * rows_per_cta + tr; r < a.rows; r += (long long)gridDim.x * rows_per_cta) {
        const long long off = r * a.C + c0;
        float g[VEC], fo[VEC], yv[VEC], gs[VEC];
        ce_load<T, VEC>(g, dout + off);
        if (a.relu) ce_load<T, VEC>(fo, fout + off);
        if (a.dscale) ce_load<T, VEC>(yv, y + off);
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          float gj = g[j];
          if (a.relu && !(fo[j] > 0.f)) gj = 0.f;
          if (a.mask) gj *= (float)a.mask[off + j];
          g[j] = gj;
          sb[j] += gj;
          if (a.dscale) ss[j] += gj * yv[j];
          gs[j] = gj * sc[j];
        }
        if (dz && dz != dy) ce_store<T, VEC>(dz + off, g);
        ce_store<T, VEC>(dy + off, gs);
      }
    }
    if (a.dbias || a.dscale) {
#pragma unroll
      for (int j = 0; j < VEC; j++) { red[0][threadIdx.x][j] = sb[j]; red[1][threadIdx.x][j] = ss[j]; }
      __syncthreads();
      if (tr == 0 && cv < cvecs) {
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          float tb = 0.f, ts = 0.f;
          for (int k = 0; k < rows_per_cta; k++) { tb += red[0][k * cols + tc][j]; ts += red[1][k * cols + tc][j]; }
          if (a.dbias) atomicAdd(a.dbias + c0 + j, tb);
          if (a.dscale) atomicAdd(a.dscale + c0 + j, ts);
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace ab

using namespace ab;

static int ce_grid(long long rows, int C, int vec) {
  const int cols = ce_cols(C / vec), rpc = kCeThreads / cols;
  long long want = (rows + rpc - 1) / rpc;
  const long long cap = (long long)kNumSMs * 8;
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}

// y / z / out: [rows, C] channels-last views of dtype dt; scale / bias: fp32 [C] or null; mask: uint8 [rows, C] or null.
AB_API int ab_conv_epilogue_fwd(const void* y, const void* z, const void* mask, const float* scale, const float* bias, void* out,
                                long long rows, int C, int relu, int dt, cudaStream_t st) {
  if (rows <= 0 || C <= 0) return 0;
  ConvEpiArgs a{};
  a.y = y; a.z = z; a.mask = (const uint8_t*)mask; a.scale = scale; a.bias = bias; a.out = out; a.rows = rows; a.C = C; a.relu = relu;
  const bool vec8 = (C % 8 == 0) && aligned16(y) && aligned16(out) && (!z || aligned16(z)) && dt != kF32;
  const bool vec4 = (C % 4 == 0) && aligned16(y) && aligned16(out) && (!z || aligned16(z)) && dt == kF32;
#define CE_FWD(T, V) conv_epi_fwd<T, V><<<ce_grid(rows, C, V), kCeThreads, 0, st>>>(a)
  if (dt == kBF16) { if (vec8) CE_FWD(bf16, 8); else CE_FWD(bf16, 1); }
  else if (dt == kF16) { if (vec8) CE_FWD(f16, 8); else CE_FWD(f16, 1); }
  else if (dt == kF32) { if (vec4) CE_FWD(float, 4); else CE_FWD(float, 1); }
  else return -1;
#undef CE_FWD
  AB_CHECK_LAUNCH();
  return 0;
}

// dout / fout / y / dy / dz: [rows, C]; dbias / dscale: fp32 [C], ZERO on entry. dz == dy is allowed when scale == null.
AB_API int ab_conv_epilogue_bwd(const void* dout, const void* fout, const void* y, const void* mask, const float* scale, void* dy, void* dz,
                                float* dbias, float* dscale, long long rows, int C, int relu, int dt, cudaStream_t st) {
  if (rows <= 0 || C <= 0) return 0;
  if (relu && !fout) return -2;
  if (dscale && !y) return -2;
  ConvEpiArgs a{};
  a.dout = dout; a.fout = fout; a.y = y; a.mask = (const uint8_t*)mask; a.scale = scale; a.dy = dy; a.dz = dz; a.dbias = dbias;
  a.dscale = dscale; a.rows = rows; a.C = C; a.relu = relu;
  const bool al = aligned16(dout) && aligned16(dy) && (!fout || aligned16(fout)) && (!y || aligned16(y)) && (!dz || aligned16(dz));
  const bool vec8 = (C % 8 == 0) && al && dt != kF32;
  const bool vec4 = (C % 4 == 0) && al && dt == kF32;
#define CE_BWD(T, V) conv_epi_bwd<T, V><<<ce_grid(rows, C, V), kCeThreads, 0, st>>>(a)
  if (dt == kBF16) { if (vec8) CE_BWD(bf16, 8); else CE_BWD(bf16, 1); }
  else if (dt == kF16) { if (vec8) CE_BWD(f16, 8); else CE_BWD(f16, 1); }
  else if (dt == kF32) { if (vec4) CE_BWD(float, 4); else CE_BWD(float, 1); }
  else return -1;
#undef CE_BWD
  AB_CHECK_LAUNCH();
  return 0;
}
