// Fused multi-head attention forward on tcgen05 (validated on B200 against an fp32 reference: tests/test_gpu_fmha.py).
//   O = dropout(softmax(scale * Q K^T + key_bias + causal mask)) V      Q, K, V: [rows, heads, d] views (any row / head stride: packed qkv, sbhd, bshd), d in {64, 128}
// Spec: reference apex/contrib/csrc/fmha (mma.sync kernels, fp16, d = 64, seq <= 512) and the batched-GEMM + softmax pipeline of
// apex/contrib/csrc/multihead_attn. Variable-length batches through cu_seqlens, optional causal mask.
//
// One CTA per (128-query tile, head, sequence), 12 warps. Warp roles as in gemm_sm100.cu: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM
// allocation, warps 4-11 softmax: TWO threads per query row (= TMEM lane; warps w and w + 4 share a lane quarter and take one 64-key
// half of every tile each) -- the softmax is instruction-issue bound, eight warps hide the TMEM-load / MUFU latencies that four could
// not (ncu: 38 % issue utilisation with four). P is double buffered so the softmax of tile j + 1 overlaps the P V of tile j.
// ONE pass over the keys, online softmax with LAZY rescaling (the first version made two passes: row maxima first, then a second
// Q K^T per tile; that cost one extra MMA and one extra TMEM sweep per tile):
//   * the two halves of a row are INDEPENDENT softmax streams: each thread keeps its own running maximum m and sum l over its 64-key
//     half-tiles, and each half has its own accumulator O_h in TMEM (S double buffer 256 columns + 2 x d columns <= 512), so the two
//     threads of a row never have to agree on a maximum; the epilogue merges O_0 * 2^(m_0 - m) + O_1 * 2^(m_1 - m);
//   * P = exp2(S * scale * log2e - m_used); m_used only moves when the tile maximum exceeds it by more than 8 (a factor 256, harmless
//     in fp32 accumulators and exact after the final division by l): then the thread waits for the previous P V to retire, multiplies
//     its row of O_h by 2^(m_used - m_new) with a tcgen05.ld / tcgen05.st round trip and carries on -- after the first tiles this
//     almost never happens, so the common tile costs one TMEM load of S, one maximum sweep, one exponential sweep.
// Optional per-key additive bias [batch, seq_k] (key-padding masks: -inf / -10000 on padded keys) and Philox dropout on P
// (fmha_common.cuh: the backward regenerates the same mask from (seed, offset)).
#include "fmha_common.cuh"

namespace ab {
namespace fmha {

constexpr int TQ = 128;   // query rows per CTA
constexpr int TK = 128;   // keys per tile
constexpr int KV_STAGES = 2;

struct Params {
  int heads, d, causal, is_bf16;
  const int* cu_seqlens_q; const int* cu_seqlens_k;  // [batch + 1] row offsets (varlen); null => fixed seq_q / seq_k per batch
  int seq_q, seq_k;
  float scale;
  void* out; long long out_row_stride, out_head_stride;  // elements
  float* lse;                                            // [batch? packed rows][heads] or null
  const float* key_bias; long long key_bias_stride; int bias_div;  // bias row = bias_div ? head / bias_div : batch
       // additive bias per (batch, key), natural-log domain; null => none
  Dropout drop; int has_drop;
};

template <int D>
struct Smem {
  static constexpr int kQ = TQ * D * 2;
  static constexpr int kKV = TK * D * 2;
  static constexpr int kP = TQ * TK * 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQ;
  static constexpr int kVOff = kKOff + KV_STAGES * kKV;
  static constexpr int kPOff = kVOff + KV_STAGES * kKV;   // two P buffers
  static constexpr int kBarOff = kPOff + 2 * kP;
  static constexpr int kBiasOff = kBarOff + 256;     // [2][TK] floats: key bias * log2(e) of the current / next tile
  static constexpr int kTotal = kBiasOff + 2 * TK * 4 + 1024;
};

template <typename T, int D>
__global__ void __launch_bounds__(384, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                ab::fmha::Params p) {
  using S = Smem<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* q_full = bars;                        // 1
  uint64_t* k_full = bars + 1;                    // KV_STAGES
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;         // 2
  uint64_t* s_empty = s_full + 2;                 // 2
  uint64_t* p_full = s_empty + 2;                 // 2
  uint64_t* p_empty = p_full + 2;                 // 2
  uint64_t* o_full = p_empty + 2;                 // 1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: the last query tiles have the most key tiles; launching them first (blocks are issued in increasing blockIdx.x) keeps the
  // light tiles for the tail of the grid. The query tiles of one (head, sequence) stay adjacent, so their K / V tiles still meet in L2.
  const int qt = p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int q_row0 = p.cu_seqlens_q ? p.cu_seqlens_q[b] : b * p.seq_q;
  const int q_len = p.cu_seqlens_q ? p.cu_seqlens_q[b + 1] - q_row0 : p.seq_q;
  const int k_row0 = p.cu_seqlens_k ? p.cu_seqlens_k[b] : b * p.seq_k;
  const int k_len = p.cu_seqlens_k ? p.cu_seqlens_k[b + 1] - k_row0 : p.seq_k;
  if (qt * TQ >= q_len) return;  // whole CTA exits before any barrier is initialised
  // causal: query i attends keys <= i + (k_len - q_len); tiles entirely above the diagonal are skipped
  const int diag = k_len - q_len;
  const int brow = p.bias_div > 0 ? head / p.bias_div : b;
  int n_kv = (k_len + TK - 1) / TK;
  if (p.causal) { const int last_key = min(k_len - 1, qt * TQ + TQ - 1 + diag); n_kv = last_key < 0 ? 0 : last_key / TK + 1; }
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) { prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v); }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; s++) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&s_full[a], 1); mbar_init(&s_empty[a], 256); mbar_init(&p_full[a], 256); mbar_init(&p_empty[a], 1); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ================================================= TMA producer
    if (lane == 0 && n_kv > 0) {  // nothing may be in flight towards this CTA's shared memory when it exits
      mbar_expect_tx(q_full, S::kQ);
      for (int i = 0; i < D / 64; i++) tma_load_3d(smem + S::kQOff + i * (TQ * 128), &map_q, q_full, i * 64, head, q_row0 + qt * TQ);
      int ks = 0; uint32_t kph = 0; int vs = 0; uint32_t vph = 0;
      for (int j = 0; j < n_kv; j++) {
        mbar_wait(&k_empty[ks], kph ^ 1, 101);
        mbar_expect_tx(&k_full[ks], S::kKV);
        for (int i = 0; i < D / 64; i++) tma_load_3d(smem + S::kKOff + ks * S::kKV + i * (TK * 128), &map_k, &k_full[ks], i * 64, head, k_row0 + j * TK);
        if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1, 102);
        mbar_expect_tx(&v_full[vs], S::kKV);
        for (int i = 0; i < D / 64; i++) tma_load_3d(smem + S::kVOff + vs * S::kKV + i * (TK * 128), &map_v, &v_full[vs], i * 64, head, k_row0 + j * TK);
        if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer (whole warp, one elected lane issues)
    const uint32_t idesc_s = make_idesc(p.is_bf16, 0, 0, TQ, TK);   // S = Q K^T : both operands K-major (d contiguous)
    const uint32_t idesc_o = make_idesc(p.is_bf16, 0, 1, TQ, D);    // O = P V   : P K-major (keys contiguous), V MN-major (d contiguous)
    const uint32_t q_addr = smem_u32(smem + S::kQOff), p_addr = smem_u32(smem + S::kPOff);
    if (n_kv > 0) mbar_wait(q_full, 0, 110);
    int ks = 0; uint32_t kph = 0; int vs = 0; uint32_t vph = 0; int sb = 0; uint32_t sph = 0; uint32_t pph[2] = {0, 0};
    auto issue_s = [&]() {  // S[sb] = Q K[ks]^T
      mbar_wait(&k_full[ks], kph, 111);
      mbar_wait(&s_empty[sb], sph ^ 1, 112);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_addr = smem_u32(smem + S::kKOff + ks * S::kKV);
#pragma unroll
        for (int k = 0; k < D / UMMA_K; k++) {
          const uint32_t off = (k >> 2) * (TQ * 128) + (k & 3) * 32;  // 64-column blocks are [128 x 128 B] regions
          umma_f16(tmem_s + sb * TK, make_desc(q_addr + off, 16, 1024), make_desc(k_addr + off, 16, 1024), idesc_s, k ? 1u : 0u);
        }
        umma_commit(&k_empty[ks]);
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
      if (++ks == KV_STAGES) { ks = 0; kph ^= 1; }
      if (++sb == 2) { sb = 0; sph ^= 1; }
    };
    if (n_kv > 0) issue_s();                             // tile 0
    for (int j = 0; j < n_kv; j++) {
      if (j + 1 < n_kv) issue_s();                       // S of the next tile overlaps the softmax of this one
      const int pb = j & 1;
      mbar_wait(&p_full[pb], pph[pb], 113);
      mbar_wait(&v_full[vs], vph, 114);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = smem_u32(smem + S::kVOff + vs * S::kKV);
#pragma unroll
        for (int k = 0; k < TK / UMMA_K; k++) {   // keys [0, 64) accumulate into O_0, keys [64, 128) into O_1 (independent softmax streams)
          const uint64_t adesc = make_desc(p_addr + pb * S::kP + (k >> 2) * (TQ * 128) + (k & 3) * 32, 16, 1024);
          const uint64_t bdesc = make_desc(v_addr + k * 2048, TK * 128, 1024);  // 16 key rows x 128 B per step; d blocks TK*128 B apart
          umma_f16(tmem_o + (uint32_t)((k >> 2) * D), adesc, bdesc, idesc_o, (j | (k & 3)) ? 1u : 0u);
        }
        umma_commit(&v_empty[vs]);
        umma_commit(&p_empty[pb]);
        if (j == n_kv - 1) umma_commit(o_full);
      }
      __syncwarp();
      pph[pb] ^= 1;
      if (++vs == KV_STAGES) { vs = 0; vph ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================================= softmax / epilogue: thread <-> query row
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;   // lane quarter (== warp % 4) and which 64-key half of every tile
    const int row = q * 32 + lane;                          // TMEM lane == row of the tile
    const int qi = qt * TQ + row;                           // index inside the sequence
    const bool row_ok = qi < q_len;
    const float sl2 = p.scale * 1.4426950408889634f;
    int sb = 0; uint32_t sph = 0;
    float m = -INFINITY;                                    // exponent offset in use for THIS thread's half: t = S * scale * log2(e) + bias * log2(e)
    float l = 0.f;                                          // sum of exp2(t - m) over this half's keys so far
    float* bias_s = reinterpret_cast<float*>(smem + S::kBiasOff);
    const bool has_bias = p.key_bias != nullptr;
    auto stage_bias = [&](int j, int slot) {                // the 128 threads of one half stage one key each; double buffered, one barrier per tile
      const int kidx = j * TK + row;
      if (half == 0) bias_s[(slot & 1) * TK + row] = kidx < k_len ? p.key_bias[(size_t)brow * p.key_bias_stride + kidx] * 1.4426950408889634f : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    auto key_ok = [&](int kidx) { return kidx < k_len && (!p.causal || kidx <= qi + diag); };
    // does key tile j contain a key that is masked for ANY row of this query tile (sequence tail, or the causal diagonal)?
    auto tile_masked = [&](int j) { return (j + 1) * TK > k_len || (p.causal && (j + 1) * TK - 1 > qt * TQ + diag); };
    uint32_t peph[2] = {0, 0};
    const uint32_t bh = (uint32_t)(b * p.heads + head);
    const int comp_row = (qi & 1) * 2;
    const uint32_t tmem_oh = tmem_o + (uint32_t)(half * D) + ((uint32_t)(q * 32) << 16);   // this thread's row of ITS half's accumulator
    constexpr float kLazy = 8.f;                            // rescale only when the maximum grows by more than 2^8
    for (int j = 0; j < n_kv; j++) {
      if (has_bias) stage_bias(j, j);
      const float* bj = bias_s + (j & 1) * TK;
      const int pb = j & 1;
      uint8_t* pbuf = smem + S::kPOff + pb * S::kP;
      const bool masked = tile_masked(j);
      const int c_lo = half * (TK / 2);
      mbar_wait(&s_full[sb], sph, 121);
      tc_fence_after();
      // ---- this thread's 64 scores, once
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_s + ((uint32_t)(q * 32) << 16) + (uint32_t)(sb * TK + c_lo), r0);
      tmem_ld32(tmem_s + ((uint32_t)(q * 32) << 16) + (uint32_t)(sb * TK + c_lo + 32), r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);                            // S is in registers: the next Q K^T may overwrite this buffer
      // ---- tile maximum (scaled domain)
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (!masked && !has_bias) {
#pragma unroll
        for (int i = 0; i < 32; i++) { m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r0[i])); m4[(i + 2) & 3] = fmaxf(m4[(i + 2) & 3], __uint_as_float(r1[i])); }
#pragma unroll
        for (int i = 0; i < 4; i++) m4[i] *= sl2;
      } else {
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float t0 = has_bias ? fmaf(__uint_as_float(r0[i]), sl2, bj[c_lo + i]) : __uint_as_float(r0[i]) * sl2;
          const float t1 = has_bias ? fmaf(__uint_as_float(r1[i]), sl2, bj[c_lo + 32 + i]) : __uint_as_float(r1[i]) * sl2;
          if (!masked || key_ok(j * TK + c_lo + i)) m4[i & 3] = fmaxf(m4[i & 3], t0);
          if (!masked || key_ok(j * TK + c_lo + 32 + i)) m4[(i + 2) & 3] = fmaxf(m4[(i + 2) & 3], t1);
        }
      }
      const float m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      // ---- lazy rescale of this half's accumulator row
      const bool grow = m_tile > m + kLazy || (m == -INFINITY && m_tile > -INFINITY);
      if (__any_sync(0xffffffffu, grow)) {
        const float f = grow ? (m == -INFINITY ? 0.f : ex2_approx(m - m_tile)) : 1.f;
        if (j > 0) {                                        // O_h holds the tiles before this one: wait until P V (j - 1) has retired
          mbar_wait(&p_empty[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1), 124);
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < D; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_oh + (uint32_t)c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(tmem_oh + (uint32_t)c0, o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        l *= f;
        if (grow) m = m_tile;
      }
      const float off = m == -INFINITY ? 0.f : m;           // nothing valid yet in this half: every p below is exp2(-inf) = 0
      mbar_wait(&p_empty[pb], peph[pb] ^ 1, 122);           // the P V that last read this P buffer has finished (first use: passes immediately)
      // ---- P = exp2(t - m), row sum, P -> shared memory (swizzled)
#pragma unroll
      for (int hseg = 0; hseg < 2; hseg++) {
        const uint32_t (&r)[32] = hseg == 0 ? r0 : r1;
        const int c0 = c_lo + hseg * 32;
        float pv[32];
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        if (!masked && !has_bias) {
          const float nm = -off;
#pragma unroll
          for (int i = 0; i < 32; i++) { const float e = ex2_approx(fmaf(__uint_as_float(r[i]), sl2, nm)); pv[i] = e; l4[i & 3] += e; }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float e = (!masked || key_ok(j * TK + c0 + i)) ? ex2_approx(fmaf(__uint_as_float(r[i]), sl2, (has_bias ? bj[c0 + i] : 0.f) - off)) : 0.f;
            pv[i] = e; l4[i & 3] += e;
          }
        }
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        if (p.has_drop) {                                   // the row sum keeps every element; only what feeds P V is dropped
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint4 rnd = philox2x2(p.drop, (uint32_t)(qi >> 1), (uint32_t)((j * TK + c0 + i) >> 1), bh);
            if (philox_pick(rnd, comp_row) < p.drop.thresh) pv[i] = 0.f;
            if (philox_pick(rnd, comp_row + 1) < p.drop.thresh) pv[i + 1] = 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 h8;   // one cvt.rn.{bf16x2,f16x2}.f32 per pair
          h8.x = pack2<T>(pv[i], pv[i + 1]); h8.y = pack2<T>(pv[i + 2], pv[i + 3]);
          h8.z = pack2<T>(pv[i + 4], pv[i + 5]); h8.w = pack2<T>(pv[i + 6], pv[i + 7]);
          const int c = c0 + i;
          *reinterpret_cast<uint4*>(pbuf + (c >> 6) * (TQ * 128) + sw128_offset(row, c & 63)) = h8;
        }
      }
      fence_proxy_async();          // generic-proxy stores to shared memory -> visible to the tensor-core (async) proxy
      mbar_arrive(&p_full[pb]);
      peph[pb] ^= 1;
      if (++sb == 2) { sb = 0; sph ^= 1; }
    }
    // ---- epilogue: merge the two halves' streams, O = (O_0 * 2^(m_0 - m) + O_1 * 2^(m_1 - m)) / (l_0 * 2^(m_0 - m) + l_1 * 2^(m_1 - m))
    if (n_kv > 0) { mbar_wait(o_full, 0, 123); tc_fence_after(); }
    // every MMA has completed: the P buffers are free again; the two threads of a row exchange (m, l) through one of them
    float* xch = reinterpret_cast<float*>(smem + S::kPOff);
    xch[half * TQ + row] = m;
    xch[2 * TQ + half * TQ + row] = l;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float m0 = xch[row], m1 = xch[TQ + row], l0 = xch[2 * TQ + row], l1 = xch[3 * TQ + row];
    const float mm = fmaxf(m0, m1);
    const float f0 = m0 == -INFINITY ? 0.f : ex2_approx(m0 - mm), f1 = m1 == -INFINITY ? 0.f : ex2_approx(m1 - mm);
    const float lt = l0 * f0 + l1 * f1;
    const float inv_l = lt > 0.f ? (p.has_drop ? p.drop.rp : 1.f) / lt : 0.f;
    const float g0 = f0 * inv_l, g1 = f1 * inv_l;
    T* orow = reinterpret_cast<T*>(p.out) + (size_t)(q_row0 + qi) * p.out_row_stride + (size_t)head * p.out_head_stride;
    const uint32_t tmem_row = tmem_o + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = half * (D / 2); c0 < (half + 1) * (D / 2); c0 += 32) {
      uint32_t ra[32], rb[32];
      if (n_kv > 0) { tmem_ld32(tmem_row + (uint32_t)c0, ra); tmem_ld32(tmem_row + (uint32_t)(D + c0), rb); tmem_ld_wait(); }
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; e++) o8[e] = n_kv > 0 ? fmaf(__uint_as_float(ra[i + e]), g0, __uint_as_float(rb[i + e]) * g1) : 0.f;
          store_vec<T, 8>(orow + c0 + i, o8);
        }
      }
    }
    if (row_ok && p.lse && half == 0) p.lse[(size_t)(q_row0 + qi) * p.heads + head] = lt > 0.f ? mm * 0.6931471805599453f + logf(lt) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace fmha
}  // namespace ab

using namespace ab;
using ab::fmha::fmha_fwd_kernel; using ab::fmha::make_map3; using ab::fmha::TQ;

// q / k / v: 16-bit tensors viewed as [rows, heads, d] with the given element strides (head stride and row stride multiples of 8).
// cu_seqlens_*: device int32 [batch + 1] or null (then rows = batch * seq). out: [rows_q, heads, d] with its own strides.
AB_API int ab_fmha_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* cu_seqlens_q, const int* cu_seqlens_k,
                       int batch, int heads, int d, long long rows_q, long long rows_k, int max_seq_q, int seq_k, long long q_row_stride,
                       long long q_head_stride, long long k_row_stride, long long k_head_stride, long long v_row_stride,
                       long long v_head_stride, long long out_row_stride, long long out_head_stride, float scale, int causal,
                       const float* key_bias, long long key_bias_stride, int bias_div, float p_drop, unsigned long long seed, unsigned long long offset,
                       int dt, cudaStream_t st) {
  if (batch <= 0 || heads <= 0 || rows_q <= 0) return 0;
  if ((d != 64 && d != 128) || (dt != kBF16 && dt != kF16)) return -10;
  if ((q_row_stride | q_head_stride | k_row_stride | k_head_stride | v_row_stride | v_head_stride | out_row_stride | out_head_stride) % 8) return -10;
  const int is_bf16 = dt == kBF16;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_map3(&mq, q, is_bf16, rows_q, heads, d, q_row_stride, q_head_stride))) return rc;
  if ((rc = make_map3(&mk, k, is_bf16, rows_k, heads, d, k_row_stride, k_head_stride))) return rc;
  if ((rc = make_map3(&mv, v, is_bf16, rows_k, heads, d, v_row_stride, v_head_stride))) return rc;
  ab::fmha::Params p;
  p.heads = heads; p.d = d; p.causal = causal; p.is_bf16 = is_bf16; p.cu_seqlens_q = cu_seqlens_q; p.cu_seqlens_k = cu_seqlens_k;
  p.seq_q = max_seq_q; p.seq_k = seq_k; p.scale = scale; p.out = out; p.out_row_stride = out_row_stride; p.out_head_stride = out_head_stride;
  p.lse = lse; p.key_bias = key_bias; p.key_bias_stride = key_bias_stride; p.bias_div = bias_div;
  p.has_drop = p_drop > 0.f;
  p.drop = ab::fmha::make_dropout(p_drop, seed, offset);
  const dim3 grid((max_seq_q + TQ - 1) / TQ, heads, batch);
#define FMHA_GO(T, DD)                                                                                              \
  do {                                                                                                              \
    auto kern = fmha_fwd_kernel<T, DD>;                                                                             \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ab::fmha::Smem<DD>::kTotal);      \
    if (e != cudaSuccess) return (int)e;                                                                            \
    kern<<<grid, 384, ab::fmha::Smem<DD>::kTotal, st>>>(mq, mk, mv, p);                                                       \
  } while (0)
  if (is_bf16) { if (d == 64) FMHA_GO(bf16, 64); else FMHA_GO(bf16, 128); }
  else { if (d == 64) FMHA_GO(f16, 64); else FMHA_GO(f16, 128); }
  return (int)cudaGetLastError();
}
