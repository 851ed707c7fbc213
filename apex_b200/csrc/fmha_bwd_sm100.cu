// Attention backward on tcgen05 (validated on B200 against an fp32 reference: tests/test_gpu_fmha.py).
//   P = exp(scale * Q K^T + key_bias - lse),  dP = dO V^T,  dS = scale * P o (dP - delta),  delta = rowsum(dO o O)
//   dV = P^T dO,  dK = dS^T Q,  dQ = dS K        (with dropout: P -> P o mask / (1 - p) in dV, dP -> dP o mask / (1 - p) in dS;
//   the mask is regenerated from the forward's Philox (seed, offset), fmha_common.cuh)
// Spec: the backward of reference apex/contrib/csrc/fmha (fmha_dgrad_*: mma.sync, fp16, d = 64, seq <= 512) and of the softmax /
// batched-GEMM chain in apex/contrib/csrc/multihead_attn.
//
// ONE kernel template, two instantiations, no atomics (deterministic):
//   DKV = true : one CTA per (128-KEY tile, head, sequence); K and V stay resident in shared memory, 64-query tiles of Q and dO
//                stream through a 2-stage TMA ring.   S^T = K Q^T and dP^T = V dO^T land in TMEM with one KEY per lane;
//                dV += P^T dO and dK += dS^T Q accumulate in TMEM.
//   DKV = false: one CTA per (128-QUERY tile, head, sequence); Q and dO resident, 64-key tiles of K and V stream.
//                S = Q K^T, dP = dO V^T with one QUERY per lane; dQ += dS K.
// The two are the same program with the operand roles exchanged ("X" = resident pair, "Y" = streamed pair):
//   S' = X1 Y1^T,  dP' = X2 Y2^T,  [out1 += P' Y2],  out2 += dS' Y1
// S and dP are recomputed in both (one extra pair of QK-sized GEMMs) instead of a dQ reduction through global atomics.
// Warp roles as in the forward kernel: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM allocation, warps 4-11 element-wise + epilogue (two
// threads per resident row: warps w and w + 4 share a TMEM lane quarter and take one 32-column half of every 64-column tile each).
// TMEM columns: S'[2] at 0 / 64, dP'[2] at 128 / 192, out1 at 256, out2 at 384 (fp32, 128 lanes).
#include "fmha_common.cuh"

namespace ab {
namespace fmha {

constexpr int TO = 128;   // resident ("outer") rows per CTA
constexpr int TI = 64;    // streamed ("inner") rows per tile
constexpr int Y_STAGES = 2;

struct BwdParams {
  int heads, causal, is_bf16;
  const int* cu_seqlens_q; const int* cu_seqlens_k;
  int seq_q, seq_k;                      // fixed lengths when cu_seqlens_* are null
  float scale;
  const float* lse; const float* delta;  // [rows_q, heads] fp32: log-sum-exp of the forward (natural log), rowsum(dO o O)
  void* out1; long long out1_row_stride, out1_head_stride;  // dV (DKV only)
  void* out2; long long out2_row_stride, out2_head_stride;  // dK (DKV) or dQ
  const float* key_bias; long long key_bias_stride; int bias_div;  // bias row = bias_div ? head / bias_div : batch
          // as in the forward
  Dropout drop; int has_drop;
};

template <int D>
struct BwdSmem {
  static constexpr int kX = TO * D * 2;            // one resident operand
  static constexpr int kY = TI * D * 2;            // one streamed operand tile
  static constexpr int kE = TO * TI * 2;           // P' or dS' (16-bit)
  static constexpr int kX1 = 0, kX2 = kX;
  static constexpr int kYOff = 2 * kX;             // stage s: Y1 at kYOff + s * 2 * kY, Y2 right behind it
  static constexpr int kPOff = kYOff + Y_STAGES * 2 * kY;
  static constexpr int kDSOff = kPOff + kE;
  static constexpr int kBarOff = kDSOff + kE;
  static constexpr int kStatOff = kBarOff + 256;   // [2 buffers][2 (lse, delta)][TI] floats (DKV only)
  static constexpr int kTotal = kStatOff + 2 * 2 * TI * 4 + 1024;
};

template <typename T, int D, bool DKV>
__global__ void __launch_bounds__(384, 1)
fmha_bwd_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                const __grid_constant__ CUtensorMap map_y1, const __grid_constant__ CUtensorMap map_y2, BwdParams p) {
  using S = BwdSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* x_full = bars;                       // 1
  uint64_t* y_full = bars + 1;                   // Y_STAGES
  uint64_t* y_empty = y_full + Y_STAGES;         // Y_STAGES
  uint64_t* s_full = y_empty + Y_STAGES;         // 2
  uint64_t* s_empty = s_full + 2;                // 2
  uint64_t* e_full = s_empty + 2;                // 1: P' / dS' written to shared memory
  uint64_t* e_empty = e_full + 1;                // 1: the GEMMs that read them have completed
  uint64_t* o_full = e_empty + 1;                // 1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  float* stat = reinterpret_cast<float*>(smem + S::kStatOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal dQ: the last query tiles have the most key tiles -- heaviest first (dK / dV already is: key tile 0 meets every query tile)
  const int ot = (!DKV && p.causal) ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int q_row0 = p.cu_seqlens_q ? p.cu_seqlens_q[b] : b * p.seq_q;
  const int q_len = p.cu_seqlens_q ? p.cu_seqlens_q[b + 1] - q_row0 : p.seq_q;
  const int k_row0 = p.cu_seqlens_k ? p.cu_seqlens_k[b] : b * p.seq_k;
  const int k_len = p.cu_seqlens_k ? p.cu_seqlens_k[b + 1] - k_row0 : p.seq_k;
  const int outer_len = DKV ? k_len : q_len, outer_row0 = DKV ? k_row0 : q_row0;
  const int inner_len = DKV ? q_len : k_len, inner_row0 = DKV ? q_row0 : k_row0;
  if (ot * TO >= outer_len) return;  // whole CTA exits before any barrier is initialised
  // causal: query i attends keys <= i + diag. Inner tiles that are masked for every row of this CTA are skipped.
  const int diag = k_len - q_len;
  const int brow = p.bias_div > 0 ? head / p.bias_div : b;
  int j0 = 0, j1 = (inner_len + TI - 1) / TI;
  if (p.causal) {
    if (DKV) { const int q_min = max(0, ot * TO - diag); j0 = min(j1, q_min / TI); }
    else { const int last_key = min(k_len - 1, ot * TO + TO - 1 + diag); j1 = last_key < 0 ? 0 : last_key / TI + 1; }
  }
  const int n = j1 - j0;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) { prefetch_tmap(&map_x1); prefetch_tmap(&map_x2); prefetch_tmap(&map_y1); prefetch_tmap(&map_y2); }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < Y_STAGES; s++) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&s_full[a], 1); mbar_init(&s_empty[a], 256); }
    mbar_init(e_full, 256); mbar_init(e_empty, 1); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_o1 = tmem_base + 256, tmem_o2 = tmem_base + 384;

  if (warp == 0) {
    // ================================================= TMA producer
    if (lane == 0 && n > 0) {  // nothing may be in flight towards this CTA's shared memory when it exits
      mbar_expect_tx(x_full, 2 * S::kX);
      for (int i = 0; i < D / 64; i++) {
        tma_load_3d(smem + S::kX1 + i * (TO * 128), &map_x1, x_full, i * 64, head, outer_row0 + ot * TO);
        tma_load_3d(smem + S::kX2 + i * (TO * 128), &map_x2, x_full, i * 64, head, outer_row0 + ot * TO);
      }
      int ys = 0; uint32_t yph = 0;
      for (int j = j0; j < j1; j++) {
        mbar_wait(&y_empty[ys], yph ^ 1, 201);
        mbar_expect_tx(&y_full[ys], 2 * S::kY);
        uint8_t* y1 = smem + S::kYOff + ys * 2 * S::kY;
        for (int i = 0; i < D / 64; i++) {
          tma_load_3d(y1 + i * (TI * 128), &map_y1, &y_full[ys], i * 64, head, inner_row0 + j * TI);
          tma_load_3d(y1 + S::kY + i * (TI * 128), &map_y2, &y_full[ys], i * 64, head, inner_row0 + j * TI);
        }
        if (++ys == Y_STAGES) { ys = 0; yph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer (whole warp, one elected lane issues)
    if (n > 0) {
      const uint32_t idesc_s = make_idesc(p.is_bf16, 0, 0, TO, TI);  // S' / dP': both operands K-major (d contiguous), N = 64
      const uint32_t idesc_o = make_idesc(p.is_bf16, 0, 1, TO, D);   // out += E Y: E K-major (inner index contiguous), Y MN-major
      const uint32_t x1_addr = smem_u32(smem + S::kX1), x2_addr = smem_u32(smem + S::kX2);
      const uint32_t p_addr = smem_u32(smem + S::kPOff), ds_addr = smem_u32(smem + S::kDSOff);
      mbar_wait(x_full, 0, 210);
      int ys_s = 0; uint32_t yph_s = 0; int ys_o = 0; int sb = 0; uint32_t sph = 0; uint32_t eph = 0;
      auto issue_s = [&]() {  // S'[sb] = X1 Y1^T, dP'[sb] = X2 Y2^T
        mbar_wait(&y_full[ys_s], yph_s, 211);
        mbar_wait(&s_empty[sb], sph ^ 1, 212);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t y1_addr = smem_u32(smem + S::kYOff + ys_s * 2 * S::kY), y2_addr = y1_addr + S::kY;
#pragma unroll
          for (int k = 0; k < D / UMMA_K; k++) {  // 64-column blocks: [128 x 128 B] for X, [64 x 128 B] for Y
            const uint32_t xo = (k >> 2) * (TO * 128) + (k & 3) * 32, yo = (k >> 2) * (TI * 128) + (k & 3) * 32;
            umma_f16(tmem_s + sb * TI, make_desc(x1_addr + xo, 16, 1024), make_desc(y1_addr + yo, 16, 1024), idesc_s, k ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < D / UMMA_K; k++) {
            const uint32_t xo = (k >> 2) * (TO * 128) + (k & 3) * 32, yo = (k >> 2) * (TI * 128) + (k & 3) * 32;
            umma_f16(tmem_dp + sb * TI, make_desc(x2_addr + xo, 16, 1024), make_desc(y2_addr + yo, 16, 1024), idesc_s, k ? 1u : 0u);
          }
          umma_commit(&s_full[sb]);
        }
        __syncwarp();
        if (++ys_s == Y_STAGES) { ys_s = 0; yph_s ^= 1; }
        if (++sb == 2) { sb = 0; sph ^= 1; }
      };
      issue_s();
      for (int jj = 0; jj < n; jj++) {
        if (jj + 1 < n) issue_s();               // the next S' / dP' overlap the element-wise work on this tile
        mbar_wait(e_full, eph, 213);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t y1_addr = smem_u32(smem + S::kYOff + ys_o * 2 * S::kY), y2_addr = y1_addr + S::kY;
#pragma unroll
          for (int k = 0; k < TI / UMMA_K; k++) {  // 16 inner rows per step = two 8-row atoms of the MN-major Y tile (2048 B)
            if (DKV)
              umma_f16(tmem_o1, make_desc(p_addr + k * 32, 16, 1024), make_desc(y2_addr + k * 2048, TI * 128, 1024), idesc_o, (jj | k) ? 1u : 0u);
            umma_f16(tmem_o2, make_desc(ds_addr + k * 32, 16, 1024), make_desc(y1_addr + k * 2048, TI * 128, 1024), idesc_o, (jj | k) ? 1u : 0u);
          }
          umma_commit(&y_empty[ys_o]);
          umma_commit(e_empty);
          if (jj == n - 1) umma_commit(o_full);
        }
        __syncwarp();
        eph ^= 1;
        if (++ys_o == Y_STAGES) ys_o = 0;
      }
    }
  } else if (warp >= 4) {
    // ================================================= element-wise: thread <-> resident row (= TMEM lane)
    const int qd = (warp - 4) & 3, half = (warp - 4) >> 2, row = qd * 32 + lane, tid = threadIdx.x - 128;   // tid in [0, 256)
    const int oi = ot * TO + row;                            // index of the resident row inside its sequence
    const bool row_ok = oi < outer_len;
    const float sl2 = p.scale * 1.4426950408889634f;
    float my_lse2 = 0.f, my_delta = 0.f, my_bias2 = 0.f;
    const bool has_bias = p.key_bias != nullptr;
    if (DKV && has_bias && row_ok) my_bias2 = p.key_bias[(size_t)brow * p.key_bias_stride + oi] * 1.4426950408889634f;
    const uint32_t bh = (uint32_t)(b * p.heads + head);
    if (!DKV && row_ok) {
      my_lse2 = p.lse[(size_t)(q_row0 + oi) * p.heads + head] * 1.4426950408889634f;
      my_delta = p.delta[(size_t)(q_row0 + oi) * p.heads + head];
    }
    int sb = 0; uint32_t sph = 0; uint32_t eeph = 0;
    uint8_t* pbuf = smem + S::kPOff;
    uint8_t* dsbuf = smem + S::kDSOff;
    for (int jj = 0; jj < n; jj++) {
      const int inner0 = (j0 + jj) * TI;
      const float* st = stat + (jj & 1) * 2 * TI;
      if (DKV) {  // per-QUERY statistics of this tile -> shared memory (read as broadcasts below); buffers alternate
        float* sw = stat + (jj & 1) * 2 * TI;
        if (tid < 2 * TI) {
          const int qi = inner0 + (tid & (TI - 1));
          const float* src = tid < TI ? p.lse : p.delta;
          float val = qi < q_len ? src[(size_t)(q_row0 + qi) * p.heads + head] : 0.f;
          if (tid < TI) val *= 1.4426950408889634f;
          sw[tid] = val;                                      // tid < 64: lse * log2(e); 64 <= tid < 128: delta
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else if (has_bias) {  // per-KEY bias of this tile -> the (otherwise unused) statistics buffer
        float* sw = stat + (jj & 1) * 2 * TI;
        if (tid < TI) { const int ki = inner0 + tid; sw[tid] = ki < k_len ? p.key_bias[(size_t)brow * p.key_bias_stride + ki] * 1.4426950408889634f : 0.f; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      // does this (outer tile, inner tile) pair contain ANY masked (query, key) combination?
      const int key_lo = DKV ? ot * TO : inner0, key_hi = DKV ? ot * TO + TO - 1 : inner0 + TI - 1;
      const int q_lo = DKV ? inner0 : ot * TO, q_hi = DKV ? inner0 + TI - 1 : ot * TO + TO - 1;
      (void)key_lo;
      const bool masked = key_hi >= k_len || q_hi >= q_len || (p.causal && key_hi > q_lo + diag);
      mbar_wait(&s_full[sb], sph, 220);
      mbar_wait(e_empty, eeph ^ 1, 221);  // the GEMMs of the previous tile have finished reading P' / dS' (first use: passes)
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * (TI / 2); c0 < (half + 1) * (TI / 2); c0 += 32) {
        uint32_t rs[32], rd[32];
        tmem_ld32(tmem_s + ((uint32_t)(qd * 32) << 16) + (uint32_t)(sb * TI + c0), rs);
        tmem_ld32(tmem_dp + ((uint32_t)(qd * 32) << 16) + (uint32_t)(sb * TI + c0), rd);
        tmem_ld_wait();
        if (!masked && !p.has_drop && !has_bias) {
          // interior tile, no dropout: no per-element predicates or index arithmetic (they were half of the issued instructions)
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float pe8[8], ds8[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const float l2 = DKV ? st[c0 + i + e] : my_lse2, dl = DKV ? st[TI + c0 + i + e] : my_delta;
              const float pe = ex2_approx(fmaf(__uint_as_float(rs[i + e]), sl2, -l2));
              pe8[e] = pe;
              ds8[e] = pe * (__uint_as_float(rd[i + e]) - dl) * p.scale;
            }
            const uint32_t off = sw128_offset(row, c0 + i);
            if (DKV) *reinterpret_cast<uint4*>(pbuf + off) = make_uint4(pack2<T>(pe8[0], pe8[1]), pack2<T>(pe8[2], pe8[3]), pack2<T>(pe8[4], pe8[5]), pack2<T>(pe8[6], pe8[7]));
            *reinterpret_cast<uint4*>(dsbuf + off) = make_uint4(pack2<T>(ds8[0], ds8[1]), pack2<T>(ds8[2], ds8[3]), pack2<T>(ds8[4], ds8[5]), pack2<T>(ds8[6], ds8[7]));
          }
          continue;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          T p8[8], d8[8];
          uint4 rnd = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const int ii = inner0 + c0 + i + e;
            const int key = DKV ? oi : ii, qi = DKV ? ii : oi;
            const bool ok = key < k_len && qi < q_len && (!p.causal || key <= qi + diag);
            const float l2 = DKV ? st[c0 + i + e] : my_lse2, dl = DKV ? st[TI + c0 + i + e] : my_delta;
            const float b2 = DKV ? my_bias2 : (has_bias ? st[c0 + i + e] : 0.f);
            const float pe = ok ? ex2_approx(fmaf(__uint_as_float(rs[i + e]), sl2, b2 - l2)) : 0.f;
            float pd = pe, dp = __uint_as_float(rd[i + e]);
            if (p.has_drop) {
              if ((e & 1) == 0) {
                const int ii0 = inner0 + c0 + i + e;
                rnd = DKV ? philox2x2(p.drop, (uint32_t)(ii0 >> 1), (uint32_t)(oi >> 1), bh) : philox2x2(p.drop, (uint32_t)(oi >> 1), (uint32_t)(ii0 >> 1), bh);
              }
              const int comp = DKV ? (e & 1) * 2 + (oi & 1) : (oi & 1) * 2 + (e & 1);
              const bool keep = philox_pick(rnd, comp) >= p.drop.thresh;
              pd = keep ? pe * p.drop.rp : 0.f;
              dp = keep ? dp * p.drop.rp : 0.f;
            }
            p8[e] = from_f<T>(pd);
            d8[e] = from_f<T>(pe * (dp - dl) * p.scale);
          }
          const uint32_t off = sw128_offset(row, c0 + i);
          if (DKV) *reinterpret_cast<uint4*>(pbuf + off) = *reinterpret_cast<const uint4*>(p8);
          *reinterpret_cast<uint4*>(dsbuf + off) = *reinterpret_cast<const uint4*>(d8);
        }
      }
      tc_fence_before();
      fence_proxy_async();          // generic-proxy stores to shared memory -> visible to the tensor-core (async) proxy
      mbar_arrive(e_full);
      mbar_arrive(&s_empty[sb]);
      eeph ^= 1;
      if (++sb == 2) { sb = 0; sph ^= 1; }
    }
    // ---- epilogue: TMEM accumulators -> global (rows past the sequence end are not stored)
    if (n > 0) { mbar_wait(o_full, 0, 222); tc_fence_after(); }
    const size_t grow = (size_t)(outer_row0 + oi);
#pragma unroll 1
    for (int which = DKV ? 0 : 1; which < 2; which++) {
      T* orow = which == 0 ? reinterpret_cast<T*>(p.out1) + grow * p.out1_row_stride + (size_t)head * p.out1_head_stride
                           : reinterpret_cast<T*>(p.out2) + grow * p.out2_row_stride + (size_t)head * p.out2_head_stride;
      const uint32_t tm = which == 0 ? tmem_o1 : tmem_o2;
#pragma unroll 1
      for (int c0 = half * (D / 2); c0 < (half + 1) * (D / 2); c0 += 32) {
        uint32_t r[32];
        if (n > 0) { tmem_ld32(tm + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, r); tmem_ld_wait(); }
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; e++) o8[e] = n > 0 ? __uint_as_float(r[i + e]) : 0.f;
            store_vec<T, 8>(orow + c0 + i, o8);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace fmha
}  // namespace ab

using namespace ab;
using ab::fmha::fmha_bwd_kernel; using ab::fmha::make_map3;

// q / k / v / dout: 16-bit [rows, heads, d] views (element strides, multiples of 8); lse / delta: fp32 [rows_q, heads] contiguous;
// dq / dk / dv: outputs with their own strides. cu_seqlens_*: device int32 [batch + 1] or null (then rows = batch * seq).
AB_API int ab_fmha_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta, void* dq, void* dk,
                       void* dv, const int* cu_seqlens_q, const int* cu_seqlens_k, int batch, int heads, int d, long long rows_q,
                       long long rows_k, int max_seq_q, int max_seq_k, long long q_row_stride, long long q_head_stride,
                       long long k_row_stride, long long k_head_stride, long long v_row_stride, long long v_head_stride,
                       long long do_row_stride, long long do_head_stride, long long dq_row_stride, long long dq_head_stride,
                       long long dk_row_stride, long long dk_head_stride, long long dv_row_stride, long long dv_head_stride, float scale,
                       int causal, const float* key_bias, long long key_bias_stride, int bias_div, float p_drop, unsigned long long seed,
                       unsigned long long offset, int dt, cudaStream_t st) {
  if (batch <= 0 || heads <= 0 || rows_q <= 0 || rows_k <= 0) return 0;
  if ((d != 64 && d != 128) || (dt != kBF16 && dt != kF16)) return -10;
  if ((q_row_stride | q_head_stride | k_row_stride | k_head_stride | v_row_stride | v_head_stride | do_row_stride | do_head_stride |
       dq_row_stride | dq_head_stride | dk_row_stride | dk_head_stride | dv_row_stride | dv_head_stride) % 8) return -10;
  const int is_bf16 = dt == kBF16;
  using ab::fmha::TO; using ab::fmha::TI;
  CUtensorMap mq_o, mdo_o, mk_o, mv_o, mq_i, mdo_i, mk_i, mv_i;  // _o: 128-row boxes (resident operand), _i: 64-row boxes (streamed)
  int rc;
  if ((rc = make_map3(&mq_o, q, is_bf16, rows_q, heads, d, q_row_stride, q_head_stride, TO))) return rc;
  if ((rc = make_map3(&mdo_o, dout, is_bf16, rows_q, heads, d, do_row_stride, do_head_stride, TO))) return rc;
  if ((rc = make_map3(&mk_o, k, is_bf16, rows_k, heads, d, k_row_stride, k_head_stride, TO))) return rc;
  if ((rc = make_map3(&mv_o, v, is_bf16, rows_k, heads, d, v_row_stride, v_head_stride, TO))) return rc;
  if ((rc = make_map3(&mq_i, q, is_bf16, rows_q, heads, d, q_row_stride, q_head_stride, TI))) return rc;
  if ((rc = make_map3(&mdo_i, dout, is_bf16, rows_q, heads, d, do_row_stride, do_head_stride, TI))) return rc;
  if ((rc = make_map3(&mk_i, k, is_bf16, rows_k, heads, d, k_row_stride, k_head_stride, TI))) return rc;
  if ((rc = make_map3(&mv_i, v, is_bf16, rows_k, heads, d, v_row_stride, v_head_stride, TI))) return rc;
  ab::fmha::BwdParams p;
  p.heads = heads; p.causal = causal; p.is_bf16 = is_bf16; p.cu_seqlens_q = cu_seqlens_q; p.cu_seqlens_k = cu_seqlens_k;
  p.seq_q = max_seq_q; p.seq_k = max_seq_k; p.scale = scale; p.lse = lse; p.delta = delta;
  p.key_bias = key_bias; p.key_bias_stride = key_bias_stride; p.bias_div = bias_div; p.has_drop = p_drop > 0.f; p.drop = ab::fmha::make_dropout(p_drop, seed, offset);
  ab::fmha::BwdParams pkv = p, pq = p;
  pkv.out1 = dv; pkv.out1_row_stride = dv_row_stride; pkv.out1_head_stride = dv_head_stride;
  pkv.out2 = dk; pkv.out2_row_stride = dk_row_stride; pkv.out2_head_stride = dk_head_stride;
  pq.out1 = nullptr; pq.out1_row_stride = 0; pq.out1_head_stride = 0;
  pq.out2 = dq; pq.out2_row_stride = dq_row_stride; pq.out2_head_stride = dq_head_stride;
  const dim3 grid_kv((max_seq_k + TO - 1) / TO, heads, batch), grid_q((max_seq_q + TO - 1) / TO, heads, batch);
#define FMHA_BWD_GO(T, DD)                                                                                                   \
  do {                                                                                                                       \
    auto kkv = fmha_bwd_kernel<T, DD, true>;                                                                                 \
    auto kq = fmha_bwd_kernel<T, DD, false>;                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, ab::fmha::BwdSmem<DD>::kTotal);   \
    if (e != cudaSuccess) return (int)e;                                                                                     \
    e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, ab::fmha::BwdSmem<DD>::kTotal);                \
    if (e != cudaSuccess) return (int)e;                                                                                     \
    kkv<<<grid_kv, 384, ab::fmha::BwdSmem<DD>::kTotal, st>>>(mk_o, mv_o, mq_i, mdo_i, pkv);   /* X = (K, V), Y = (Q, dO) */     \
    kq<<<grid_q, 384, ab::fmha::BwdSmem<DD>::kTotal, st>>>(mq_o, mdo_o, mk_i, mv_i, pq);      /* X = (Q, dO), Y = (K, V) */     \
  } while (0)
  if (is_bf16) { if (d == 64) FMHA_BWD_GO(bf16, 64); else FMHA_BWD_GO(bf16, 128); }
  else { if (d == 64) FMHA_BWD_GO(f16, 64); else FMHA_BWD_GO(f16, 128); }
  return (int)cudaGetLastError();
}
