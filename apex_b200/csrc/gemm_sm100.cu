// Blackwell-native GEMM for the fused-dense family: D[M,N] = op(A) * op(B) with fused epilogues.
//
//   * operands staged in shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B), a KSTAGES-deep mbarrier ring;
//   * tcgen05.mma (kind::f16, cta_group::1, 128 x BN x 16) issued by ONE elected thread, fp32 accumulators in TMEM;
//   * two TMEM accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1;
//   * warp-specialised persistent CTAs (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//     warps 4..7 = epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global).
//
// Either operand may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) so that forward (x W^T), dgrad (dy W)
// and wgrad (dy^T x) all run without materialising a transpose.
//
// Epilogues (reference cuBLASLt call sites in csrc/fused_dense_cuda.cu):
//   EPI_NONE       plain store                               (dgrad, :1104-1129)
//   EPI_BIAS       + bias[N]                                 (EPILOGUE_BIAS, :81)
//   EPI_BIAS_GELU  + bias, GELU; pre-activation saved to aux (EPILOGUE_GELU_AUX_BIAS, :315,336)
//   EPI_DGELU      * gelu'(aux)                              (EPILOGUE_DGELU_BGRAD, :829)
//   any epilogue + Params::colsum: the bias gradient (column sums of the stored tile) accumulated from the epilogue registers
//   EPI_ACCUM      + C (beta = 1), fp32 or 16-bit main-grad  (fused_weight_gradient_dense_cuda.cu:42-49)
//   EPI_BIAS_RELU / EPI_BIAS_SIGMOID / EPI_RELU / EPI_SIGMOID (mlp_cuda.cu:95-119,271-405)
#include "gemm_common.cuh"
#include <cstdlib>

namespace ab {
namespace gemm {

template <int BN, int KSTAGES>
struct Smem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = KSTAGES * kStage;
  static constexpr int kTotal = kBarOff + 256 + 1024;  // barriers + tmem ptr + alignment slack
};

// Tile rasterisation: groups of kGroupM tile-rows are walked column by column, so the tiles that are in flight together
// (one per CTA / CTA pair) cover a compact ~8 x 9 block of the output and share A and B panels in L2 instead of streaming all of A.
constexpr int kGroupM = 8;
__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& m_blk, int& n_blk) {
  const int per_group = kGroupM * num_n;
  const int g = t / per_group, r = t - g * per_group;
  const int m0 = g * kGroupM;
  const int gm = min(kGroupM, num_m - m0);
  n_blk = r / gm;
  m_blk = m0 + (r - n_blk * gm);
}

template <typename TOut, int BN, int KSTAGES>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                           Params p, int is_bf16) {
  using S = Smem<BN, KSTAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + KSTAGES;
  uint64_t* tmem_full = empty_bar + KSTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (p.K + BK - 1) / BK;
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: power of two

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < KSTAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 256); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (one elected lane)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk; tile_coords(t, num_m, num_n, m_blk, n_blk);
        if (p.ready_flags) {   // B rows [n0, n1) x all of K: contiguous in the parameter buffer (K-major weight, ldb == K)
          const long long n0 = (long long)n_blk * BN, n1 = min((long long)p.N, n0 + BN);
          acquire_buckets(p.ready_flags, p.ready_epoch, p.ready_world, (int)((p.ready_w_off + n0 * p.K) / p.ready_bucket_elems),
                          (int)((p.ready_w_off + n1 * p.K - 1) / p.ready_bucket_elems));
        }
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * S::kStage;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStage);
          if (!p.a_mn_major) {
            tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m_blk * BM);            // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; i++)                                              // box {64 m, 64 k} x2
              tma_load_2d(sa + i * (BK * 128), &map_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
          }
          if (!p.b_mn_major) {
            tma_load_2d(sb, &map_b, &full_bar[stage], kb * BK, n_blk * BN);            // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; i++)
              tma_load_2d(sb + i * (BK * 128), &map_b, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
          }
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer. The WHOLE warp runs the (uniform) loop so that descriptor
    // arithmetic stays in the uniform datapath; one elected lane issues. Shared-memory descriptors are built once: per k-block only
    // a 32-bit add on the address field remains (the loop body must cost less than the 4 x 64 tensor cycles it feeds).
    {
      const uint32_t idesc = make_idesc(is_bf16, p.a_mn_major, p.b_mn_major, BM, BN);
      // K-major: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); advance 32 B per UMMA_K inside the atom.
      // MN-major: 64-wide MN blocks BK*128 B apart (LBO), 8-k-row atoms 1024 B apart (SBO); advance 2 atoms per UMMA_K.
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t a0 = p.a_mn_major ? make_desc(smem0, BK * 128, 1024) : make_desc(smem0, 16, 1024);
      const uint64_t b0 = p.b_mn_major ? make_desc(smem0 + S::kABytes, BK * 128, 1024) : make_desc(smem0 + S::kABytes, 16, 1024);
      const uint32_t a_step = p.a_mn_major ? (2048u >> 4) : (32u >> 4), b_step = p.b_mn_major ? (2048u >> 4) : (32u >> 4);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = a0 + (uint32_t)(stage * (S::kStage >> 4)), bd = b0 + (uint32_t)(stage * (S::kStage >> 4));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++)
              umma_f16(tmem_d, ad + k * a_step, bd + k * b_step, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);                 // frees the smem slot when these MMAs have read it
            if (kb == num_k - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================== epilogue: TMEM -> registers -> fused op -> global
    const int q = (warp - kEpiWarp0) & 3, half = (warp - kEpiWarp0) >> 2;  // q == warp % 4: the TMEM lane quarter this warp may touch
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk; tile_coords(t, num_m, num_n, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase, 4);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      epilogue_tile<TOut, BN>(p, tmem_base, acc, row, n_blk, q, half);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// =====================================================================================================================
// 2-CTA variant: a cluster of two CTAs (one SM pair) owns a 256 x 256 output tile. tcgen05.mma.cta_group::2 (M = 256) is issued by
// the leader CTA only; each CTA stages its own 128 rows of A and its own 128-row HALF of the B tile, so a pair pulls 64 KB per
// k-block from L2 instead of 2 x 48 KB (the 1-CTA kernel above is L2-bandwidth bound: 12 GB through L2 for the 8192x16384x4096
// FFN GEMM), and the freed shared memory deepens the ring to 6 stages. Accumulators: rows 0-127 in CTA0's TMEM, 128-255 in CTA1's.
//   full[s]   (leader's copy is the one waited on): both CTAs' TMA loads complete_tx on the LEADER's barrier
//   empty[s]  / tmem_full[a]: tcgen05.commit ... multicast::cluster to both CTAs
//   tmem_empty[a] (leader's copy): 256 local + 256 remote (mapa) arrivals from the two 8-warp epilogues
// =====================================================================================================================
constexpr int BN2 = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on the barrier at the same offset in the LEADER (even) CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_leader), "r"(c0), "r"(c1) : "memory");
}
// 3-D variant: an MN-major operand viewed as {128-byte row of MN elements, k, MN block}; one box {row, BKE k-rows, all of the CTA's MN
// blocks} lands in shared memory block after block -- the same image the 2-D boxes build, with one instruction instead of 2 / 4.
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int KSTAGES>
struct Smem2 {
  static constexpr int kABytes = BM * BK * 2;         // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN2 / 2) * BK * 2;  // this CTA's half of the B tile
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = KSTAGES * kStage;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

// KIND selects the operand element size; every BYTE quantity (128-byte swizzle rows, 16 KB per operand per stage, 32 bytes of K per MMA,
// four MMAs per k-block) is the same for all three, only the element counts differ:
//   KIND_16 : bf16 / fp16, kind::f16,     64 k-elements per block, K = 16 per MMA
//   KIND_8  : E4M3 / E5M2, kind::f8f6f4, 128 k-elements per block, K = 32 per MMA (K-major operands only)
//   KIND_32 : fp32 words read as TF32, kind::tf32, 32 k-elements per block, K = 8 per MMA
// MN-major operands are staged as boxes of {one 128-byte row of MN elements, BKE k-rows}: 64 / 32 elements wide for 16-bit / tf32.
constexpr int KIND_16 = 0, KIND_8 = 1, KIND_32 = 2;
template <typename TOut, int KSTAGES, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params p, int is_bf16) {
  using S = Smem2<KSTAGES>;
  extern __shared__ uint8_t smem_raw[];
  // dynamic shared memory starts at the same offset in both CTAs, so the rounded-up base is the same too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + KSTAGES;
  uint64_t* tmem_full = empty_bar + KSTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int num_m = (p.M + 2 * BM - 1) / (2 * BM), num_n = (p.N + BN2 - 1) / BN2;
  const int num_tiles = num_m * num_n;
  constexpr int ESZ = KIND == KIND_16 ? 2 : (KIND == KIND_8 ? 1 : 4);
  constexpr int BKE = 128 / ESZ;         // k-elements per block (128 bytes per K-major row)
  constexpr int MNB = 128 / ESZ;         // MN-elements per 128-byte row of an MN-major box
  constexpr int kMnBoxBytes = BKE * 128; // one MN-major box: BKE k-rows of 128 bytes
  const int num_k = (p.K + BKE - 1) / BKE;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  constexpr uint32_t kTmemCols = 2 * BN2;

  if (warp == 0 && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < KSTAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 512); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / TMA completion can target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (one lane in EACH CTA)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int m_blk, n_blk; tile_coords(t, num_m, num_n, m_blk, n_blk);
        const int m0 = m_blk * 2 * BM + (int)cta * BM;         // this CTA's A rows
        const int n0 = n_blk * BN2 + (int)cta * (BN2 / 2);     // this CTA's half of B
        if (p.ready_flags && n0 < p.N) {
          const long long r1 = min((long long)p.N, (long long)n0 + BN2 / 2);
          acquire_buckets(p.ready_flags, p.ready_epoch, p.ready_world, (int)((p.ready_w_off + (long long)n0 * p.K) / p.ready_bucket_elems),
                          (int)((p.ready_w_off + r1 * p.K - 1) / p.ready_bucket_elems));
        }
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 11);
          uint8_t* sa = smem + stage * S::kStage;
          uint8_t* sb = sa + S::kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::kStage);  // bytes of both CTAs land on the leader's barrier
          if (!p.a_mn_major) tma_load_2d_2sm(sa, &map_a, &full_bar[stage], kb * BKE, m0);
          else if (p.a_mn3) tma_load_3d_2sm(sa, &map_a, &full_bar[stage], 0, kb * BKE, m0 / MNB);
          else {
#pragma unroll
            for (int i = 0; i < BM / MNB; i++) tma_load_2d_2sm(sa + i * kMnBoxBytes, &map_a, &full_bar[stage], m0 + i * MNB, kb * BKE);
          }
          if (!p.b_mn_major) tma_load_2d_2sm(sb, &map_b, &full_bar[stage], kb * BKE, n0);
          else if (p.b_mn3) tma_load_3d_2sm(sb, &map_b, &full_bar[stage], 0, kb * BKE, n0 / MNB);
          else {
#pragma unroll
            for (int i = 0; i < (BN2 / 2) / MNB; i++) tma_load_2d_2sm(sb + i * kMnBoxBytes, &map_b, &full_bar[stage], n0 + i * MNB, kb * BKE);
          }
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only; whole warp, one elected lane issues)
    if (leader) {
      const uint32_t idesc = make_idesc(is_bf16, p.a_mn_major, p.b_mn_major, 2 * BM, BN2);
      const uint32_t smem0 = smem_u32(smem);
      // MN-major: boxes kMnBoxBytes apart along MN (LBO), 8-k-row atoms 1024 B apart (SBO); one MMA covers 32 / ESZ k-rows
      constexpr uint32_t kMnStep = (32u / ESZ) * 128u;
      // 32-bit MN-major operands: 32-byte swizzle atoms with a 4-row period (SWIZZLE_128B_BASE32B), so k-row groups are 512 B apart
      constexpr uint32_t kMnLayout = KIND == KIND_32 ? 1u : 2u, kMnSbo = KIND == KIND_32 ? 512u : 1024u;
      constexpr uint32_t lbo = kMnBoxBytes, sbo = kMnSbo, mstep = kMnStep;
      const uint64_t a0 = p.a_mn_major ? make_desc(smem0, lbo, sbo, kMnLayout) : make_desc(smem0, 16, 1024);
      const uint64_t b0 = p.b_mn_major ? make_desc(smem0 + S::kABytes, lbo, sbo, kMnLayout) : make_desc(smem0 + S::kABytes, 16, 1024);
      const uint32_t a_step = p.a_mn_major ? (mstep >> 4) : (32u >> 4), b_step = p.b_mn_major ? (mstep >> 4) : (32u >> 4);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 12);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN2);
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(&full_bar[stage], phase, 13);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = a0 + (uint32_t)(stage * (S::kStage >> 4)), bd = b0 + (uint32_t)(stage * (S::kStage >> 4));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++)
              if (KIND == KIND_8) umma_f8_2sm(tmem_d, ad + k * a_step, bd + k * b_step, idesc, (kb | k) != 0 ? 1u : 0u);
              else if (KIND == KIND_32) umma_tf32_2sm(tmem_d, ad + k * a_step, bd + k * b_step, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_f16_2sm(tmem_d, ad + k * a_step, bd + k * b_step, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == num_k - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == KSTAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================== epilogue (both CTAs: their own 128 rows)
    const int q = (warp - kEpiWarp0) & 3, half = (warp - kEpiWarp0) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int m_blk, n_blk; tile_coords(t, num_m, num_n, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase, 14);
      tc_fence_after();
      const int row = m_blk * 2 * BM + (int)cta * BM + q * 32 + lane;
      epilogue_tile<TOut, BN2>(p, tmem_base, acc, row, n_blk, q, half);
      tc_fence_before();
      if (leader) mbar_arrive(&tmem_empty[acc]);
      else mbar_arrive_remote(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still signal into this CTA's shared memory
  if (warp == 2) { tc_fence_after(); tmem_dealloc2(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is a DRIVER call and needs a context bound to the calling thread. A fresh thread (e.g. the autograd engine's
// backward thread) whose first CUDA activity is this call has none yet -- runtime calls bind the primary context lazily, driver calls
// do not (observed: CUDA_ERROR_INVALID_CONTEXT when the caching allocator served every allocation of the backward pass).
static void bind_primary_context_once() {
  static thread_local bool done = false;
  if (!done) {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess) cudaSetDevice(d);
    cudaFree(nullptr);
    done = true;
  }
}

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// Descriptor cache: a training loop issues the same GEMMs (same pointers, shapes, boxes) every step, and cuTensorMapEncodeTiled is a
// driver call (~1-2 us each, two per GEMM). Small direct-mapped cache keyed by everything that goes into the descriptor.
struct MapKey { const void* ptr; uint64_t rows, cols, ld; uint32_t box_cols, box_rows; int fmt, esize, atom32; };
struct MapSlot { MapKey k; CUtensorMap m; bool valid; };
static inline bool key_eq(const MapKey& a, const MapKey& b) {
  return a.ptr == b.ptr && a.rows == b.rows && a.cols == b.cols && a.ld == b.ld && a.box_cols == b.box_cols && a.box_rows == b.box_rows &&
         a.fmt == b.fmt && a.esize == b.esize && a.atom32 == b.atom32;
}
static int make_map_uncached(CUtensorMap* m, const void* ptr, int is_bf16, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                             uint32_t box_rows, int esize, int atom32);
// atom32: SWIZZLE_128B_ATOM_32B instead of SWIZZLE_128B (MN-major tf32 operands, see make_desc)
static int make_map(CUtensorMap* m, const void* ptr, int is_bf16, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                    uint32_t box_rows, int esize = 2, int atom32 = 0) {
  constexpr int kSlots = 256;
  static thread_local MapSlot cache[kSlots];   // thread_local: the autograd engine calls from its own threads, no locking needed
  const MapKey k{ptr, rows, cols, ld, box_cols, box_rows, is_bf16, esize, atom32};
  uint64_t h = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull ^ (rows * 0xC2B2AE3D27D4EB4Full) ^ (cols << 17) ^ (ld << 3) ^ box_rows ^ ((uint64_t)box_cols << 9);
  MapSlot& slot = cache[(h >> 32) % kSlots];
  if (slot.valid && key_eq(slot.k, k)) { *m = slot.m; return 0; }
  const int rc = make_map_uncached(m, ptr, is_bf16, rows, cols, ld, box_cols, box_rows, esize, atom32);
  if (rc == 0) { slot.k = k; slot.m = *m; slot.valid = true; }
  return rc;
}

// 2-D row-major [rows, cols] 16-bit tensor with leading dimension ld (elements); box = {box_cols (inner), box_rows}.
static int make_map_uncached(CUtensorMap* m, const void* ptr, int is_bf16, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                             uint32_t box_rows, int esize, int atom32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1001;
  bind_primary_context_once();
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (cuuint64_t)esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(2000 + (int)r);
}

// MN-major operand [k_rows, mn_cols] (row-major, leading dim ld) as a 3-D tensor {mnb (one 128-byte row), k_rows, mn_cols / mnb}
// with box {mnb, box_k, nblk}; requires mn_cols % mnb == 0. Cached like the 2-D maps (box_cols carries nblk << 16 | mnb).
static int make_map3(CUtensorMap* m, const void* ptr, int fmt, uint64_t k_rows, uint64_t mn_cols, uint64_t ld, uint32_t mnb, uint32_t box_k,
                     uint32_t nblk, int esize, int atom32) {
  constexpr int kSlots = 64;
  static thread_local MapSlot cache[kSlots];
  const MapKey k{ptr, k_rows, mn_cols, ld, (nblk << 16) | mnb, box_k, fmt, esize, atom32 | 2};
  uint64_t h = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull ^ (k_rows * 0xC2B2AE3D27D4EB4Full) ^ (mn_cols << 17) ^ (ld << 3) ^ box_k;
  MapSlot& slot = cache[(h >> 32) % kSlots];
  if (slot.valid && key_eq(slot.k, k)) { *m = slot.m; return 0; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1001;
  bind_primary_context_once();
  cuuint64_t dims[3] = {mnb, k_rows, mn_cols / mnb};
  cuuint64_t strides[2] = {ld * (cuuint64_t)esize, (cuuint64_t)mnb * esize};
  cuuint32_t box[3] = {mnb, box_k, nblk};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : ((fmt & 1) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUresult r = fn(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -(2000 + (int)r);
  slot.k = k; slot.m = *m; slot.valid = true;
  return 0;
}

template <typename TOut, int BN, int KSTAGES>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, int is_bf16, int sms, cudaStream_t st) {
  using S = Smem<BN, KSTAGES>;
  static bool attr_done = false;
  auto kern = gemm_kernel<TOut, BN, KSTAGES>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, kThreads, S::kTotal, st>>>(ma, mb, p, is_bf16);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

template <typename TOut, int KSTAGES, int KIND = KIND_16>
static int launch2(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, int is_bf16, int sms, cudaStream_t st) {
  using S = Smem2<KSTAGES>;
  static bool attr_done = false;
  auto kern = gemm2_kernel<TOut, KSTAGES, KIND>;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN2 - 1) / BN2);
  int clusters = sms / 2;
  if (tiles < clusters) clusters = tiles;
  kern<<<2 * clusters, kThreads, S::kTotal, st>>>(ma, mb, p, is_bf16);
  return (int)cudaGetLastError();
}

}  // namespace gemm
}  // namespace ab

using namespace ab;
using namespace ab::gemm;

// D[M,N] (dt_out) = op(A) op(B), A/B 16-bit (bf16 or fp16, same type).
//   a_mn_major == 0: A is row-major [M, K] with leading dim lda;  == 1: A is row-major [K, M].
//   b_mn_major == 0: B is row-major [N, K] with leading dim ldb;  == 1: B is row-major [K, N].
// Requirements (else returns -10 so the caller can fall back): inner extents and leading dims multiples of 8 elements,
// 16-byte aligned bases.
AB_API int ab_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, long long lda, long long ldb, long long ldd,
                        int a_mn_major, int b_mn_major, int dt_in, int dt_out, int epi, const void* bias, void* aux, long long ldaux,
                        const void* C, long long ldc, float* colsum, const void* ready_flags, unsigned int ready_epoch, int ready_world,
                        long long ready_w_off, long long ready_bucket_elems, int sms, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (dt_in != kBF16 && dt_in != kF16) return -10;
  const int is_bf16 = dt_in == kBF16;
  if ((lda % 8) || (ldb % 8) || !aligned16(A) || !aligned16(B)) return -10;
  if ((!a_mn_major && (K % 8)) || (a_mn_major && (M % 8)) || (!b_mn_major && (K % 8)) || (b_mn_major && (N % 8))) return -10;
  const int BN = (N >= 192 || N > 128) ? 256 : 128;
  CUtensorMap ma, mb;
  int rc;
  // 2-CTA (SM pair) kernel for anything with at least one full 256 x 256 tile; the single-CTA kernel for small / skinny problems
  const bool use2 = (M > 128) && (N > 128) && (getenv("APEX_B200_GEMM_1CTA") == nullptr);
  static const bool no3d = getenv("APEX_B200_GEMM_NO3D") != nullptr;
  const bool a3 = use2 && a_mn_major && (M % 64 == 0) && !no3d, b3 = use2 && b_mn_major && (N % 64 == 0) && !no3d;
  if (!a_mn_major) rc = make_map(&ma, A, is_bf16, M, K, lda, BK, BM);
  else if (a3) rc = make_map3(&ma, A, is_bf16, K, M, lda, 64, BK, BM / 64, 2, 0);
  else rc = make_map(&ma, A, is_bf16, K, M, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map(&mb, B, is_bf16, N, K, ldb, BK, use2 ? BN2 / 2 : BN);
  else if (b3) rc = make_map3(&mb, B, is_bf16, K, N, ldb, 64, BK, (BN2 / 2) / 64, 2, 0);
  else rc = make_map(&mb, B, is_bf16, K, N, ldb, 64, BK);
  if (rc) return rc;
  Params p;
  p.M = M; p.N = N; p.K = K; p.D = D; p.ldd = ldd; p.bias = bias; p.aux = aux; p.ldaux = ldaux; p.C = C; p.ldc = ldc;
  p.a_mn_major = a_mn_major; p.b_mn_major = b_mn_major; p.epi = epi; p.alpha = 1.f; p.scale_a = nullptr; p.scale_b = nullptr;
  p.a_mn3 = a3; p.b_mn3 = b3;
  p.colsum = colsum;
  // flag-guarded B: only the forward layout (K-major weight, dense rows) maps tile rows to a contiguous range of the parameter buffer
  const bool guarded = ready_flags != nullptr && !b_mn_major && ldb == K && ready_bucket_elems > 0;
  p.ready_flags = guarded ? (const uint32_t*)ready_flags : nullptr; p.ready_epoch = ready_epoch; p.ready_world = ready_world;
  p.ready_w_off = ready_w_off; p.ready_bucket_elems = ready_bucket_elems;
  if (ready_flags != nullptr && !guarded) return -11;   // the caller must wait for the whole buffer instead
  if (sms <= 0) sms = kNumSMs;
#define GEMM_GO(T)                                                                      \
  if (use2) return launch2<T, 6>(ma, mb, p, is_bf16, sms, st);                          \
  return BN == 256 ? launch<T, 256, 4>(ma, mb, p, is_bf16, sms, st) : launch<T, 128, 6>(ma, mb, p, is_bf16, sms, st)
  if (dt_out == kBF16) { GEMM_GO(bf16); }
  if (dt_out == kF16) { GEMM_GO(f16); }
  if (dt_out == kF32) { GEMM_GO(float); }
  return -1;
}

// D[M,N] (dt_out) = alpha * A B^T with A [M,K] (dt_in), B [N,K] (dt_b) row-major 8-bit floats (kE4M3 / kE5M2, may differ),
// fp32 accumulation in TMEM (tcgen05.mma.kind::f8f6f4, cta_group::2). K, lda, ldb must be multiples of 16 (TMA 16-byte rule).
// scale_a / scale_b: optional device floats multiplied into alpha in the epilogue (per-tensor dequantisation scales).
AB_API int ab_gemm_fp8(const void* A, const void* B, void* D, int M, int N, int K, long long lda, long long ldb, long long ldd, int dt_in,
                       int dt_b, int dt_out, int epi, const void* bias, void* aux, long long ldaux, float alpha, const float* scale_a, const float* scale_b,
                       int sms, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (dt_in != kE4M3 && dt_in != kE5M2) return -10;
  if (dt_b != kE4M3 && dt_b != kE5M2) return -10;
  if ((K % 16) || (lda % 16) || (ldb % 16) || !aligned16(A) || !aligned16(B)) return -10;
  if (epi == EPI_ACCUM) return -10;
  const int fmt = (dt_in == kE5M2 ? 1 : 0) | (dt_b != dt_in ? 2 : 0);   // bit 0: A is E5M2; bit 1: B has the other 8-bit format
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, fmt, M, K, lda, 2 * BK, BM, 1);
  if (rc) return rc;
  rc = make_map(&mb, B, fmt, N, K, ldb, 2 * BK, BN2 / 2, 1);
  if (rc) return rc;
  Params p;
  p.M = M; p.N = N; p.K = K; p.D = D; p.ldd = ldd; p.bias = bias; p.aux = aux; p.ldaux = ldaux; p.C = nullptr; p.ldc = 0;
  p.a_mn_major = 0; p.b_mn_major = 0; p.a_mn3 = 0; p.b_mn3 = 0; p.epi = epi; p.alpha = alpha; p.scale_a = scale_a; p.scale_b = scale_b; p.colsum = nullptr; p.ready_flags = nullptr;
  if (sms <= 0) sms = kNumSMs;
  if (dt_out == kBF16) return launch2<bf16, 6, KIND_8>(ma, mb, p, fmt, sms, st);
  if (dt_out == kF16) return launch2<f16, 6, KIND_8>(ma, mb, p, fmt, sms, st);
  if (dt_out == kF32) return launch2<float, 6, KIND_8>(ma, mb, p, fmt, sms, st);
  return -1;
}

// D[M,N] (fp32) = op(A) op(B) with fp32 operands multiplied as TF32 (tcgen05.mma.kind::tf32: 10-bit mantissa products, fp32 accumulation
// in TMEM) -- the tensor-core path for fp32 layers when the caller allows TF32 (torch.backends.cuda.matmul.allow_tf32 /
// float32_matmul_precision != "highest"); IEEE fp32 products have no tensor-core instruction and stay a library SGEMM.
// Operand layouts as in ab_gemm_bf16. Requirements (else -10): inner extents and leading dims multiples of 4 elements, 16-byte bases.
AB_API int ab_gemm_tf32(const void* A, const void* B, void* D, int M, int N, int K, long long lda, long long ldb, long long ldd,
                        int a_mn_major, int b_mn_major, int epi, const void* bias, void* aux, long long ldaux, const void* C, long long ldc,
                        float* colsum, int sms, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 4) || (ldb % 4) || !aligned16(A) || !aligned16(B)) return -10;
  if ((!a_mn_major && (K % 4)) || (a_mn_major && (M % 4)) || (!b_mn_major && (K % 4)) || (b_mn_major && (N % 4))) return -10;
  constexpr int BKE = 32, MNB = 32;
  CUtensorMap ma, mb;
  int rc;
  static const bool no3d = getenv("APEX_B200_GEMM_NO3D") != nullptr;
  const bool a3 = a_mn_major && (M % MNB == 0) && !no3d, b3 = b_mn_major && (N % MNB == 0) && !no3d;
  if (!a_mn_major) rc = make_map(&ma, A, kFmtTF32, M, K, lda, BKE, BM, 4);
  else if (a3) rc = make_map3(&ma, A, kFmtTF32, K, M, lda, MNB, BKE, BM / MNB, 4, 1);
  else rc = make_map(&ma, A, kFmtTF32, K, M, lda, MNB, BKE, 4, 1);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map(&mb, B, kFmtTF32, N, K, ldb, BKE, BN2 / 2, 4);
  else if (b3) rc = make_map3(&mb, B, kFmtTF32, K, N, ldb, MNB, BKE, (BN2 / 2) / MNB, 4, 1);
  else rc = make_map(&mb, B, kFmtTF32, K, N, ldb, MNB, BKE, 4, 1);
  if (rc) return rc;
  Params p;
  p.M = M; p.N = N; p.K = K; p.D = D; p.ldd = ldd; p.bias = bias; p.aux = aux; p.ldaux = ldaux; p.C = C; p.ldc = ldc;
  p.a_mn_major = a_mn_major; p.b_mn_major = b_mn_major; p.epi = epi; p.alpha = 1.f; p.scale_a = nullptr; p.scale_b = nullptr;
  p.a_mn3 = a3; p.b_mn3 = b3;
  p.colsum = colsum; p.ready_flags = nullptr; p.ready_epoch = 0; p.ready_world = 0; p.ready_w_off = 0; p.ready_bucket_elems = 0;
 
  if (sms <= 0) sms = kNumSMs;
  return launch2<float, 6, KIND_32>(ma, mb, p, kFmtTF32, sms, st);
}

// Column sums of a row-major [M, N] matrix (bias gradients): out[n] = sum_m x[m, n]. fp32 accumulation, deterministic.
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial(const T* __restrict__ x, float* __restrict__ part, int M, int N, long long ld,
                                                      int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  if (c >= N) return;
  float s = 0.f;
  for (int r = r0; r < r1; r++) s += to_f<T>(x[(size_t)r * ld + c]);
  part[(size_t)blockIdx.y * N + c] = s;
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_final(const float* __restrict__ part, T* __restrict__ out, int N, int nparts) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int k = 0; k < nparts; k++) s += part[(size_t)k * N + c];
  out[c] = from_f<T>(s);
}

// ws: >= 64 * N floats
AB_API int ab_colsum(const void* x, void* out, float* ws, int M, int N, long long ld, int dt, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  int parts = (M + 127) / 128;
  if (parts > 64) parts = 64;
  const int rpb = (M + parts - 1) / parts;
  dim3 grid((N + 255) / 256, parts);
  AB_DISPATCH_FLOAT3(dt, T, colsum_partial<T><<<grid, 256, 0, st>>>((const T*)x, ws, M, N, ld, rpb);
                     colsum_final<T><<<(N + 255) / 256, 256, 0, st>>>(ws, (T*)out, N, parts));
  AB_CHECK_LAUNCH();
  return 0;
}

// Stream-ordered wait (no kernel, no spinning SM): work enqueued on `st` after this call starts only once the 32-bit word at `addr`
// (device memory, e.g. a bucket-ready flag written by a peer's step kernel) is >= value. Used for consumers that are not flag-aware.
typedef CUresult (*StreamWaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
AB_API int ab_stream_wait_geq(const void* addr, unsigned int value, cudaStream_t st) {
  static StreamWaitFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return -1001;
    fn = reinterpret_cast<StreamWaitFn>(f);
  }
  bind_primary_context_once();
  CUresult r = fn((CUstream)st, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  return r == CUDA_SUCCESS ? 0 : -(2000 + (int)r);
}
