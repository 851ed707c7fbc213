// Multi-tensor optimizer updates on the device-table engine: Adam/AdamW (+capturable, +fp32 master), Adagrad, SGD,
// NovoGrad, update_scale_hysteresis.
// Behavioural spec: reference csrc/multi_tensor_adam.cu:23-376, multi_tensor_adagrad.cu:24-96,
// multi_tensor_sgd_kernel.cu:28-184, multi_tensor_novograd.cu:26-139, update_scale_hysteresis.cu:5-55.
#include "mt_engine.cuh"

namespace ab {

struct NoCtx2 {};

// ---------------------------------------------------------------- Adam
// slots: 0=g 1=p 2=m 3=v [4=p_master]. With a master slot the math runs on the master copy and p gets the cast.
// Host-scalar mode: lr/bc1/bc2 passed by value. Capturable mode: lr_ptr/step_ptr/inv_scale_ptr device scalars,
// kernel exits when *noop != 0, grads are unscaled in place (reference multi_tensor_adam.cu:156-157).
template <bool kMaster, bool kCapturable>
struct AdamOp {
  static constexpr unsigned kRead = kMaster ? 0b11101u : 0b01111u;
  static constexpr unsigned kWrite = kMaster ? (kCapturable ? 0b11111u : 0b11110u) : (kCapturable ? 0b01111u : 0b01110u);
  static constexpr int kAcc = 0;
  struct Ctx { float lr, bc1, bc2, inv_scale; };
  float beta1, beta2, eps, decay; int mode;  // mode 0: L2, 1: decoupled (AdamW)
  float lr, bc1, bc2;
  const float* lr_ptr; const int* step_ptr; const float* inv_scale_ptr; const int* noop; int bias_correction;
  __device__ bool skip() const { return kCapturable && noop && *noop != 0; }
  __device__ Ctx begin(int) const {
    Ctx c{lr, bc1, bc2, 1.f};
    if (kCapturable) {
      c.lr = *lr_ptr;
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
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx& c, float (&)[2], int) const {
    constexpr int P = kMaster ? 4 : 1;
#pragma unroll
    for (int j = 0; j < V; j++) {
      float g = r[0][j] * c.inv_scale;
      if (kCapturable) r[0][j] = g;
      float p = r[P][j], m = r[2][j], v = r[3][j];
      if (mode == 0) g += decay * p;
      m = beta1 * m + (1.f - beta1) * g;
      v = beta2 * v + (1.f - beta2) * g * g;
      const float denom = sqrtf(v / c.bc2) + eps;
      float upd = (m / c.bc1) / denom;
      if (mode != 0) upd += decay * p;
      p -= c.lr * upd;
      r[2][j] = m; r[3][j] = v; r[P][j] = p;
      if (kMaster) r[1][j] = p;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- Adam + stochastic weight average + low-precision compute copy in ONE pass
// slots: 0 g (bf16/fp16/fp32), 1 p (fp32), 2 m, 3 v, 4 swa (fp32), 5 compute copy (16-bit). Reference: the single Triton multi-tensor kernel of
// apex/contrib/openfold_triton/fused_adam_swa.py:209-400 (Adam math modes, swa = decay * swa + (1 - decay) * p, bf16 copy, clip scale).
struct AdamSwaOp {
  static constexpr unsigned kRead = 0b011111u, kWrite = 0b111110u;
  static constexpr int kAcc = 0;
  using Ctx = NoCtx2;
  float beta1, beta2, eps, decay, lr, bc1, bc2, swa_a, swa_b; int mode;  // mode 0: L2 (ApexAdam / PyTorchAdam), 1: decoupled (ApexAdamW)
  int torch_math;           // PyTorchAdam: denom = sqrt(v) / sqrt(bc2) + eps, step size lr / bc1 (same value, different rounding order)
  const float* clip_scale;  // optional device scalar multiplied into every gradient
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int) const { return Ctx{}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&)[2], int) const {
    const float gs = clip_scale ? *clip_scale : 1.f;
#pragma unroll
    for (int j = 0; j < V; j++) {
      float g = r[0][j] * gs, p = r[1][j], m = r[2][j], v = r[3][j];
      if (mode == 0) g += decay * p;
      m = beta1 * m + (1.f - beta1) * g;
      v = beta2 * v + (1.f - beta2) * g * g;
      float upd;
      if (torch_math) upd = (m / bc1) / (sqrtf(v) / sqrtf(bc2) + eps);
      else upd = (m / bc1) / (sqrtf(v / bc2) + eps);
      if (mode != 0) upd += decay * p;
      p -= lr * upd;
      r[1][j] = p; r[2][j] = m; r[3][j] = v;
      r[4][j] = swa_a * r[4][j] + swa_b * p;
      r[5][j] = p;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- Adagrad: slots g,p,h
struct AdagradOp {
  static constexpr unsigned kRead = 0b111u, kWrite = 0b110u;
  static constexpr int kAcc = 0;
  using Ctx = NoCtx2;
  float eps, lr, decay; int mode;
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      float g = r[0][j], p = r[1][j], h = r[2][j];
      if (mode == 0) {  // L2
        g += decay * p;
        h += g * g;
        p -= lr * (g / (sqrtf(h) + eps));
      } else {  // AdamW-style
        h += g * g;
        p -= lr * (g / (sqrtf(h) + eps) + decay * p);
      }
      r[1][j] = p; r[2][j] = h;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- SGD: slots g,p,mom,[p_model]
template <bool kModelCopy>
struct SgdOp {
  static constexpr unsigned kRead = 0b0111u;
  static constexpr unsigned kWrite = kModelCopy ? 0b1110u : 0b0110u;
  static constexpr int kAcc = 0;
  using Ctx = NoCtx2;
  float wd, momentum, dampening, lr, scale; int nesterov, first_run, wd_after_momentum; const int* noop;
  __device__ bool skip() const { return noop && *noop != 0; }
  __device__ Ctx begin(int) const { return {}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx&, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      float g = r[0][j] * scale, p = r[1][j], mom = r[2][j];
      if (wd != 0.f && !wd_after_momentum) g += wd * p;
      if (momentum != 0.f) {
        mom = first_run ? g : mom * momentum + (1.f - dampening) * g;
        g = nesterov ? g + momentum * mom : mom;
      }
      if (wd != 0.f && wd_after_momentum) g += wd * p;
      p -= lr * g;
      r[1][j] = p; r[2][j] = mom;
      if (kModelCopy) r[3][j] = p;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- NovoGrad: slots g,p,m + per-tensor grad norm
struct NovoGradOp {
  static constexpr unsigned kRead = 0b111u, kWrite = 0b110u;
  static constexpr int kAcc = 0;
  struct Ctx { float denom; };
  float beta1, beta3, bc1, bc2, eps, lr, decay; int m_mode; const float* per_tensor_norm;
  __device__ bool skip() const { return false; }
  __device__ Ctx begin(int t) const { return Ctx{per_tensor_norm[t] / bc2 + eps}; }
  template <int D, int V>
  __device__ __forceinline__ void apply(float (&r)[D][V], const Ctx& c, float (&)[2], int) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      float g = r[0][j], p = r[1][j], m = r[2][j];
      if (m_mode == 0) {  // regularisation inside the moment
        g = g / c.denom + decay * p;
        m = beta1 * m + beta3 * g;
        p -= lr * (m / bc1);
      } else {
        m = beta1 * m + beta3 * g;
        p -= lr * ((m / bc1) / c.denom + decay * p);
      }
      r[1][j] = p; r[2][j] = m;
    }
  }
  __device__ void end(const Ctx&, int, int, float (&)[2], float*) const {}
};

// ---------------------------------------------------------------- GradScaler update with hysteresis (device scalars)
// One warp does it (no reason for a <<<1,1>>> launch to be slower than that); semantics of
// reference csrc/update_scale_hysteresis.cu:5-41.
__global__ void update_scale_hysteresis_kernel(float* scale, int* growth_tracker, int* hysteresis_tracker,
                                               const float* found_inf, double growth_factor, double backoff_factor,
                                               int growth_interval, int hysteresis) {
  if (threadIdx.x != 0) return;
  if (*found_inf > 0.f) {
    *hysteresis_tracker -= 1;
    if (*hysteresis_tracker <= 0) {
      *scale = (float)((double)(*scale) * backoff_factor);
      *growth_tracker = 0;
    } else {
      *growth_tracker = 0;  // an overflow always restarts the clean-step counter
    }
    return;
  }
  int successful = *growth_tracker + 1;
  if (successful == growth_interval) {
    float ns = (float)((double)(*scale) * growth_factor);
    if (finite_f(ns)) *scale = ns;  // never grow to inf
    successful = 0;
  }
  *growth_tracker = successful;
  *hysteresis_tracker = hysteresis;  // a clean step re-arms the hysteresis budget
}

}  // namespace ab

using namespace ab;
#define TB make_table(arena, n, depth, total_chunks, chunk)

// Adam. depth 4: [g,p,m,v]; depth 5: [g,p,m,v,p_master(f32)]. m,v are fp32.
// capturable != 0 => lr_ptr/step_ptr/inv_scale_ptr/noop are device pointers.
AB_API int ab_mt_adam(void* arena, int n, int depth, int total_chunks, int chunk, int dt_g, int dt_p, float lr, float beta1,
                      float beta2, float eps, int step, int mode, int bias_correction, float decay, int capturable,
                      const float* lr_ptr, const int* step_ptr, const float* inv_scale_ptr, const int* noop,
                      cudaStream_t st) {
  float bc1 = 1.f, bc2 = 1.f;
  if (bias_correction && !capturable) {
    bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  }
  MTTable tb = TB;
#define ADAM_GO(MASTER, CAP, ...)                                                                              \
  {                                                                                                             \
    AdamOp<MASTER, CAP> op{beta1, beta2, eps, decay, mode, lr, bc1, bc2, lr_ptr, step_ptr, inv_scale_ptr, noop, \
                           bias_correction};                                                                    \
    AB_DISPATCH_FLOAT3(dt_g, TG, AB_DISPATCH_FLOAT3(dt_p, TP, return (mt_launch<4, AdamOp<MASTER, CAP>, __VA_ARGS__>(tb, op, st)))); \
  }
  if (depth == 4) {
    if (capturable) ADAM_GO(false, true, TG, TP, float, float) else ADAM_GO(false, false, TG, TP, float, float)
  } else if (depth == 5) {
    if (capturable) ADAM_GO(true, true, TG, TP, float, float, float) else ADAM_GO(true, false, TG, TP, float, float, float)
  }
  return -2;
}

// [g, p, m, v, swa, compute]: dt_g for g, dt_c for the compute copy; p / m / v / swa fp32.
AB_API int ab_mt_adam_swa(void* arena, int n, int depth, int total_chunks, int chunk, int dt_g, int dt_c, float lr, float beta1, float beta2,
                          float eps, int step, int mode, int torch_math, int bias_correction, float decay, float swa_a, float swa_b,
                          const float* clip_scale, cudaStream_t st) {
  if (depth != 6) return -2;
  float bc1 = 1.f, bc2 = 1.f;
  if (bias_correction) {
    bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  }
  AdamSwaOp op{beta1, beta2, eps, decay, lr, bc1, bc2, swa_a, swa_b, mode, torch_math, clip_scale};
  AB_DISPATCH_FLOAT3(dt_g, TG, AB_DISPATCH_FLOAT3(dt_c, TC, return (mt_launch<4, AdamSwaOp, TG, float, float, float, float, TC>(TB, op, st))));
  return -2;
}

AB_API int ab_mt_adagrad(void* arena, int n, int depth, int total_chunks, int chunk, int dt, float lr, float eps, int mode,
                         float decay, cudaStream_t st) {
  if (depth != 3) return -2;
  AdagradOp op{eps, lr, decay, mode};
  AB_DISPATCH_FLOAT3(dt, T, return (mt_launch<4, AdagradOp, T, T, T>(TB, op, st)));
  return 0;
}

// SGD. depth 3: [g,p,mom]; depth 4: [g,p(f32),mom(f32),p_model(dt_model)].
AB_API int ab_mt_sgd(void* arena, int n, int depth, int total_chunks, int chunk, int dt_g, int dt_p, int dt_model, float wd,
                     float momentum, float dampening, float lr, int nesterov, int first_run, int wd_after_momentum,
                     float scale, const int* noop, cudaStream_t st) {
  MTTable tb = TB;
  if (depth == 3) {
    SgdOp<false> op{wd, momentum, dampening, lr, scale, nesterov, first_run, wd_after_momentum, noop};
    AB_DISPATCH_FLOAT3(dt_g, TG, AB_DISPATCH_FLOAT3(dt_p, TP, return (mt_launch<4, SgdOp<false>, TG, TP, TP>(tb, op, st))));
  } else if (depth == 4) {
    SgdOp<true> op{wd, momentum, dampening, lr, scale, nesterov, first_run, wd_after_momentum, noop};
    AB_DISPATCH_FLOAT3(dt_g, TG,
                       AB_DISPATCH_FLOAT3(dt_model, TM, return (mt_launch<4, SgdOp<true>, TG, float, float, TM>(tb, op, st))));
  }
  return -2;
}

AB_API int ab_mt_novograd(void* arena, int n, int depth, int total_chunks, int chunk, int dt, float lr, float beta1,
                          float beta2, float eps, int step, int bias_correction, float decay, int grad_averaging,
                          int m_mode, const float* per_tensor_norm, cudaStream_t st) {
  if (depth != 3) return -2;
  float bc1 = 1.f, bc2 = 1.f;
  if (bias_correction) {
    bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  }
  const float beta3 = grad_averaging ? 1.f - beta1 : 1.f;
  NovoGradOp op{beta1, beta3, bc1, bc2, eps, lr, decay, m_mode, per_tensor_norm};
  AB_DISPATCH_FLOAT3(dt, T, return (mt_launch<4, NovoGradOp, T, T, T>(TB, op, st)));
  return 0;
}

AB_API int ab_update_scale_hysteresis(float* scale, int* growth_tracker, int* hysteresis_tracker, const float* found_inf,
                                      double growth_factor, double backoff_factor, int growth_interval, int hysteresis,
                                      cudaStream_t st) {
  update_scale_hysteresis_kernel<<<1, 32, 0, st>>>(scale, growth_tracker, hysteresis_tracker, found_inf, growth_factor,
                                                   backoff_factor, growth_interval, hysteresis);
  AB_CHECK_LAUNCH();
  return 0;
}
