// LAMB (2 passes instead of the reference's 4 launches), sharded-optimizer Adam/LAMB update functors, multi-tensor cast.
// Behavioural spec: reference csrc/multi_tensor_lamb.cu:38-319, multi_tensor_lamb_mp.cu:38-388,
// multi_tensor_lamb_stage_1.cu / _stage_2.cu, apex/contrib/csrc/optimizers/multi_tensor_distopt_adam_kernel.cu:45-461,
// multi_tensor_distopt_lamb_kernel.cu:33-406, fused_adam_cuda_kernel.cu:565-721 (maybe_cast_mt).
#include "mt_engine.cuh"

namespace ab {

struct e5m2 { uint8_t b; };
struct e4m3 { uint8_t b; };
template <> __device__ __forceinline__ float to_f<e5m2>(e5m2 v) {
  __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v.b, __NV_E5M2);
  return __half2float(__half(h));
}
template <> __device__ __forceinline__ e5m2 from_f<e5m2>(float v) {
  return e5m2{(uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2)};
}
template <> __device__ __forceinline__ float to_f<e4m3>(e4m3 v) {
  __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v.b, __NV_E4M3);
  return __half2float(__half(h));
}
template <> __device__ __forceinline__ e4m3 from_f<e4m3>(float v) {
  return e4m3{(uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3)};
}
// raw int16 lanes (bf16 bit patterns / fp32 low halves) carried exactly in float registers
template <> __device__ __forceinline__ float to_f<int16_t>(int16_t v) { return (float)v; }
template <> __device__ __forceinline__ int16_t from_f<int16_t>(float v) { return (int16_t)(int)v; }

struct NoCtx3 {};

// ---------------------------------------------------------------- LAMB pass 1
// slots: 0=g 1=p 2=m 3=v [4=separate update out]. Writes the update term into g (or slot 4), updates m,v, and
// accumulates sum(p^2) and sum(update^2) per chunk in the SAME pass (the reference runs two extra l2norm launches).
template <bool kDev, int kUpdSlot>
struct LambStage1Op {
  static constexpr unsigned kRead = 0b01111u;
  static constexpr unsigned kWrite = 0b01100u | (1u << kUpdSlot);
  static constexpr int kAcc = 2;
  struct Ctx { float clip, bc1, bc2, inv_scale, decay; };
  float beta1, beta2, beta3, bc1, bc2, eps, decay; int mode;
  const float* global_grad_norm; float max_norm; const float* max_norm_ptr;
  const int* step_ptr; int bias_correction; const float* inv_scale_ptr; const int* noop;
  const float* per_tensor_decay;
  float* partial_p; float* partial_u;
  __device__ bool skip() const { return kDev && noop && *noop != 0; }
  __device__ Ctx begin(int t) const {
    Ctx c{1.f, bc1, bc2, 1.f, per_tensor_decay ? per_tensor_decay[t] : decay};
    const float mx = (kDev && max_norm_ptr) ? *max_norm_ptr : max_norm;
    if (global_grad_norm && mx > 0.f) {
      const float gn = *global_grad_norm;
      c.clip = gn > mx ? gn / mx : 1.f;
    }
    if (kDev) {
      c.inv_scale = inv_scale_ptr ? *inv_scale_ptr : 1.f;
      if (bias_correction) {
        const float s = (float)(*step_ptr);
        c.bc1 = 1.f - powf(beta1, s);
        c.bc2 = 1.f - powf(beta2, s);
      } else { c.bc1 = 1.f; c.bc2 = 1.f; }
    }
    return c;
  }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx& c, float (&acc)[2], int nv) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      float g = (r[0][j] * c.inv_scale) / c.clip, p = r[1][j], m = r[2][j], v = r[3][j];
      if (mode == 0) g += c.decay * p;
      m = m * beta1 + beta3 * g;
      v = v * beta2 + (1.f - beta2) * g * g;
      float upd = (m / c.bc1) / (sqrtf(v / c.bc2) + eps);
      if (mode != 0) upd += c.decay * p;
      if (j < nv) { acc[0] += p * p; acc[1] += upd * upd; }
      r[kUpdSlot][j] = upd; r[2][j] = m; r[3][j] = v;
    }
  }
  __device__ void end(const Ctx&, int, int cid, float (&acc)[2], float* red) const {
    float sp = block_sum(acc[0], red);
    float su = block_sum(acc[1], red);
    if (threadIdx.x == 0) { partial_p[cid] = sp; partial_u[cid] = su; }
  }
};

// ---------------------------------------------------------------- LAMB pass 2: p -= ratio * update
// slots: kU=update, kP=param [, kM = low-precision model copy]
template <bool kDev, int kU, int kP, int kM>
struct LambStage2Op {
  static constexpr unsigned kRead = (1u << kU) | (1u << kP);
  static constexpr unsigned kWrite = (1u << kP) | (kM >= 0 ? (1u << (kM < 0 ? 0 : kM)) : 0u);
  static constexpr int kAcc = 0;
  struct Ctx { float ratio; };
  const float* pnorm; const float* unorm; float lr; const float* lr_ptr; float decay; const float* per_tensor_decay;
  int use_nvlamb; const int* noop;
  __device__ bool skip() const { return kDev && noop && *noop != 0; }
  __device__ Ctx begin(int t) const {
    const float l = (kDev && lr_ptr) ? *lr_ptr : lr;
    const float d = per_tensor_decay ? per_tensor_decay[t] : decay;
    float ratio = l;
    if (use_nvlamb || d != 0.f) {
      const float pn = pnorm[t], un = unorm[t];
      if (un != 0.f && pn != 0.f) ratio = l * (pn / un);
    }
    return Ctx{ratio};
  }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx& c, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      const float p = r[kP][j] - c.ratio * r[kU][j];
      r[kP][j] = p;
      if constexpr (kM >= 0) r[kM][j] = p;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- sharded Adam: slots p_in, m, v, g, p_out
// lerp-form moments and a DEVICE grad_scale (folds loss-unscale and clipping), reference distopt kernel :31-34,:84-168.
template <bool kCapturable>
struct DistAdamOp {
  static constexpr unsigned kRead = 0b01111u, kWrite = 0b10111u;
  static constexpr int kAcc = 0;
  struct Ctx { float gs, lr, bc1, bc2; };
  const float* grad_scale; float beta1, beta2, bc1, bc2, eps, lr, decay; int mode;
  const float* lr_ptr; const int* step_ptr; int bias_correction; const int* noop;
  __device__ bool skip() const { return kCapturable && noop && *noop != 0; }
  __device__ Ctx begin(int) const {
    Ctx c{grad_scale ? *grad_scale : 1.f, lr, bc1, bc2};
    if (kCapturable) {
      c.lr = *lr_ptr;
      if (bias_correction) {
        const float s = (float)(*step_ptr);
        c.bc1 = 1.f - powf(beta1, s); c.bc2 = 1.f - powf(beta2, s);
      } else { c.bc1 = 1.f; c.bc2 = 1.f; }
    }
    return c;
  }
  static __device__ __forceinline__ float lerp_(float t, float x, float y) { return fmaf(t, y, fmaf(-t, x, x)); }
  __device__ __forceinline__ void step1(float& p, float& m, float& v, float g, const Ctx& c) const {
    float sg = g * c.gs;
    if (mode == 0) sg += decay * p;
    m = lerp_(beta1, sg, m);
    v = lerp_(beta2, sg * sg, v);
    float upd = (m / c.bc1) / (sqrtf(v / c.bc2) + eps);
    if (mode != 0) upd += decay * p;
    p -= c.lr * upd;
  }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx& c, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      step1(r[0][j], r[1][j], r[2][j], r[3][j], c);
      r[4][j] = r[0][j];
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// bf16 params + int16 remainders == exact fp32 master without storing it. slots: p_in(i16) rem(i16) m v g p_out(i16)
struct DistAdamRemOp {
  static constexpr unsigned kRead = 0b011111u, kWrite = 0b101110u;
  static constexpr int kAcc = 0;
  using Ctx = DistAdamOp<false>::Ctx;
  DistAdamOp<false> base;
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int t) const { return base.begin(t); }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx& c, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      int hi = (int)r[0][j], lo = (int)r[1][j];
      if (lo < 0) hi -= 1;  // undo round-to-nearest carry
      float p = __uint_as_float(((unsigned)(hi & 0xffff) << 16) | (unsigned)(lo & 0xffff));
      base.step1(p, r[2][j], r[3][j], r[4][j], c);
      const unsigned u = __float_as_uint(p);
      int nlo = (int)(int16_t)(u & 0xffffu), nhi = (int)(int16_t)(u >> 16);
      if (nlo < 0) nhi += 1;
      r[1][j] = (float)nlo;
      r[5][j] = (float)(int16_t)nhi;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- cast: out = (Tout) in
struct CastOp {
  static constexpr unsigned kRead = 1u, kWrite = 2u;
  static constexpr int kAcc = 0;
  using Ctx = NoCtx3;
  float scale;
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) r[1][j] = r[0][j] * scale;
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

}  // namespace ab

using namespace ab;
#define TB make_table(arena, n, depth, total_chunks, chunk)

#define AB_DISPATCH_CAST(code, NAME, ...)                           \
  switch (code) {                                                   \
    case ab::kF32: { using NAME = float; __VA_ARGS__; break; }      \
    case ab::kF16: { using NAME = ab::f16; __VA_ARGS__; break; }    \
    case ab::kBF16: { using NAME = ab::bf16; __VA_ARGS__; break; }  \
    case ab::kE5M2: { using NAME = ab::e5m2; __VA_ARGS__; break; }  \
    case ab::kE4M3: { using NAME = ab::e4m3; __VA_ARGS__; break; }  \
    default: return -1;                                             \
  }

// per-tensor sqrt(sum(partials)) — one warp per tensor, deterministic order
__global__ void ab_norm_pt_fwd(const float* __restrict__ partial, const int* __restrict__ prefix, int n,
                               float* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int t = w; t < n; t += nw) {
    float s = 0.f;
    for (int c = prefix[t] + lane; c < prefix[t + 1]; c += 32) s += partial[c];
    s = ab::warp_sum(s);
    if (lane == 0) out[t] = sqrtf(s);
  }
}

// LAMB pass 1. table depth 4/5: [g,p,m,v,(p_model or update_out)]. dt_g for g, dt_p for p/m/v.
// device_scalars!=0: step/inv_scale/max_norm/noop pointers are honoured (mixed-precision LAMB).
// upd_slot: 0 (write update into g) or 4 (write into 5th list; legacy stage1 API).
AB_API int ab_mt_lamb_stage1(void* arena, int n, int depth, int total_chunks, int chunk, int dt_g, int dt_p, int dt_s4,
                             float beta1, float beta2, float beta3, int step, int bias_correction, float eps, int mode,
                             float decay, const float* per_tensor_decay, const float* global_grad_norm, float max_norm,
                             const float* max_norm_ptr, int device_scalars, const int* step_ptr,
                             const float* inv_scale_ptr, const int* noop, float* partial_p, float* partial_u,
                             float* pnorm, float* unorm, int upd_slot, cudaStream_t st) {
  float bc1 = 1.f, bc2 = 1.f;
  if (bias_correction && !device_scalars) {
    bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  }
  MTTable tb = TB;
  int rc = -2;
#define LAMB1(DEV, SLOT, ...)                                                                                        \
  {                                                                                                                  \
    LambStage1Op<DEV, SLOT> op{beta1, beta2, beta3, bc1, bc2, eps, decay, mode, global_grad_norm, max_norm,          \
                               max_norm_ptr, step_ptr, bias_correction, inv_scale_ptr, noop, per_tensor_decay,       \
                               partial_p, partial_u};                                                                \
    AB_DISPATCH_FLOAT3(dt_g, TG, AB_DISPATCH_FLOAT3(dt_p, TP, rc = (mt_launch<4, LambStage1Op<DEV, SLOT>, __VA_ARGS__>(tb, op, st)))); \
  }
  if (upd_slot == 0 && depth == 4) {
    if (device_scalars) LAMB1(true, 0, TG, TP, TP, TP) else LAMB1(false, 0, TG, TP, TP, TP)
  } else if (upd_slot == 0 && depth == 5) {  // 5th list (model copy) untouched in pass 1
    if (device_scalars) LAMB1(true, 0, TG, TP, TP, TP, TG) else LAMB1(false, 0, TG, TP, TP, TP, TG)
  } else if (upd_slot == 4 && depth == 5) {
    AB_DISPATCH_FLOAT3(dt_s4, TU, if (device_scalars) LAMB1(true, 4, TG, TP, TP, TP, TU) else LAMB1(false, 4, TG, TP, TP, TP, TU));
  }
  if (rc) return rc;
  if (n > 0 && pnorm && unorm) {
    int blocks = (n + 7) / 8; if (blocks > 1184) blocks = 1184;
    ab_norm_pt_fwd<<<blocks, 256, 0, st>>>(partial_p, tb.chunk_prefix, n, pnorm);
    ab_norm_pt_fwd<<<blocks, 256, 0, st>>>(partial_u, tb.chunk_prefix, n, unorm);
  }
  AB_CHECK_LAUNCH();
  return 0;
}

// LAMB pass 2 on the SAME table as pass 1 (slots 0=update,1=p,[4=model copy]) or a legacy [p, update] table (legacy=1).
AB_API int ab_mt_lamb_stage2(void* arena, int n, int depth, int total_chunks, int chunk, int dt_g, int dt_p,
                             const float* pnorm, const float* unorm, float lr, const float* lr_ptr, float decay,
                             const float* per_tensor_decay, int use_nvlamb, int device_scalars, const int* noop,
                             int model_copy, int legacy, cudaStream_t st) {
  MTTable tb = TB;
#define LAMB2(DEV, U, P, M, ...)                                                                                      \
  {                                                                                                                   \
    LambStage2Op<DEV, U, P, M> op{pnorm, unorm, lr, lr_ptr, decay, per_tensor_decay, use_nvlamb, noop};              \
    AB_DISPATCH_FLOAT3(dt_g, TG, AB_DISPATCH_FLOAT3(dt_p, TP, return (mt_launch<4, LambStage2Op<DEV, U, P, M>, __VA_ARGS__>(tb, op, st)))); \
  }
  if (legacy) {  // [p, update]
    if (depth != 2) return -2;
    LAMB2(false, 1, 0, -1, TP, TG)
  } else if (depth == 4) {
    if (device_scalars) LAMB2(true, 0, 1, -1, TG, TP, TP, TP) else LAMB2(false, 0, 1, -1, TG, TP, TP, TP)
  } else if (depth == 5) {
    if (model_copy) { if (device_scalars) LAMB2(true, 0, 1, 4, TG, TP, TP, TP, TG) else LAMB2(false, 0, 1, 4, TG, TP, TP, TP, TG) }
    else { if (device_scalars) LAMB2(true, 0, 1, -1, TG, TP, TP, TP, TG) else LAMB2(false, 0, 1, -1, TG, TP, TP, TP, TG) }
  }
  return -2;
}

// Sharded Adam: [p_in(T), m(T), v(T), g(G), p_out(O)]
AB_API int ab_mt_dist_adam(void* arena, int n, int depth, int total_chunks, int chunk, int dt_state, int dt_g, int dt_out,
                           const float* grad_scale, float lr, float beta1, float beta2, float eps, int step, int mode,
                           int bias_correction, float decay, int capturable, const float* lr_ptr, const int* step_ptr,
                           const int* noop, cudaStream_t st) {
  if (depth != 5) return -2;
  float bc1 = 1.f, bc2 = 1.f;
  if (bias_correction && !capturable) {
    bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  }
  MTTable tb = TB;
#define DADAM(CAP)                                                                                                   \
  {                                                                                                                  \
    DistAdamOp<CAP> op{grad_scale, beta1, beta2, bc1, bc2, eps, lr, decay, mode, lr_ptr, step_ptr, bias_correction, noop}; \
    AB_DISPATCH_FLOAT3(dt_state, TS, AB_DISPATCH_FLOAT3(dt_g, TG, AB_DISPATCH_FLOAT3(dt_out, TO,                     \
        return (mt_launch<4, DistAdamOp<CAP>, TS, TS, TS, TG, TO>(tb, op, st)))));                                   \
  }
  if (capturable) DADAM(true) else DADAM(false)
  return -2;
}

// Sharded Adam with bf16 params + int16 remainders: [p_in(i16), rem(i16), m(f32), v(f32), g(G), p_out(i16)]
AB_API int ab_mt_dist_adam_remainders(void* arena, int n, int depth, int total_chunks, int chunk, int dt_g,
                                      const float* grad_scale, float lr, float beta1, float beta2, float eps, int step,
                                      int mode, int bias_correction, float decay, cudaStream_t st) {
  if (depth != 6) return -2;
  float bc1 = 1.f, bc2 = 1.f;
  if (bias_correction) {
    bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  }
  DistAdamRemOp op{DistAdamOp<false>{grad_scale, beta1, beta2, bc1, bc2, eps, lr, decay, mode, nullptr, nullptr, 0, nullptr}};
  AB_DISPATCH_FLOAT3(dt_g, TG, return (mt_launch<4, DistAdamRemOp, int16_t, int16_t, float, float, TG, int16_t>(TB, op, st)));
  return 0;
}

// Multi-tensor cast/scale copy: [in, out], any of f32/f16/bf16/e5m2/e4m3 on either side.
AB_API int ab_mt_cast(void* arena, int n, int depth, int total_chunks, int chunk, int dt_in, int dt_out, float scale,
                      cudaStream_t st) {
  if (depth != 2) return -2;
  CastOp op{scale};
  AB_DISPATCH_CAST(dt_in, TI, AB_DISPATCH_CAST(dt_out, TO, return (mt_launch<4, CastOp, TI, TO>(TB, op, st))));
  return 0;
}
