// sam_road_b200 :: 2-CTA tcgen05 GEMM (cta_group::2).  Same contract and epilogues as gemm_tc.cuh, but
// a cluster of two CTAs (one TPC) computes a 256 x 256 tile with UMMA M = 256: each CTA stages only
// its 128 rows of A and its 128 rows (N) of W per k-block (32 KB / stage instead of 48 KB for a
// 128 x 256 tile), halving L2 -> SM operand traffic per FLOP -- the 1-CTA kernel is bound by that
// traffic (~12.8 TB/s measured, ncu l1tex__m_xbar2l1tex_read_bytes) at ~1.0 PFLOP/s.
//
//   both CTAs : warp 0 TMA producer (bytes credited to the leader's full barrier), warps 2..9 epilogue
//               of their own 128 accumulator rows (TMEM is per SM)
//   leader    : warp 1 lane 0 issues tcgen05.mma.cta_group::2 and multicasts the commits
#pragma once

#include <type_traits>

#include "gemm_tc.cuh"

namespace srb {

constexpr int kGemm2Stages = 5;

constexpr int kGemm2OutBytes = 32 * 64 * 2;     // one 32-row x 64-column fp16 block (128 B rows)

struct Gemm2Smem {
  static constexpr int kABytes = 128 * kGemmBK * 2;
  static constexpr int kStageBytes = 2 * kABytes;                    // A half + B half
  static constexpr int kOutOffset = kGemm2Stages * kStageBytes;      // 1024-aligned
  static constexpr int kBarOffset = kOutOffset + kGemmEpiWarps * 2 * kGemm2OutBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

// fp16 epilogue of the 2-CTA kernel: out16 = act(acc + bias), one accumulator row per lane.  The
// 32 x 64 block is laid out in smem exactly as a 128B-swizzled TMA box (16 B piece j of row r at
// r*128 + ((j ^ (r & 7)) << 4): conflict-free 16 B stores) and leaves through a TMA store, so the
// epilogue costs 2 KB of smem writes + 2 KB of TMA reads per 32 x 32 accumulator block instead of
// the 8 KB of the transposing epilogue -- shared-memory bandwidth is what the K = 768 GEMMs of the
// encoder are short of (tools/gemm_probe.py: main loop alone 1.55 PFLOP/s).
__device__ __forceinline__ void epi_f16_pack_chunk(const float (&v)[32], const float* __restrict__ bias_n,
                                                   int act, uint8_t* row_base, uint32_t piece0,
                                                   uint32_t sw, bool valid = true) {
  const float4* bp = reinterpret_cast<const float4*>(bias_n);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 x[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias_n) bb = __ldg(bp + 2 * i + h);
      x[2 * h] = __fadd2_rn(make_float2(v[8 * i + 4 * h], v[8 * i + 4 * h + 1]), make_float2(bb.x, bb.y));
      x[2 * h + 1] = __fadd2_rn(make_float2(v[8 * i + 4 * h + 2], v[8 * i + 4 * h + 3]), make_float2(bb.z, bb.w));
    }
    if (act == ACT_GELU) {
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = gelu_erf_fast2(x[k]);
    } else if (act != ACT_NONE) {
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = make_float2(apply_act(x[k].x, act), apply_act(x[k].y, act));
    }
    uint4 u;
    u.x = pack_half2(x[0].x, x[0].y);
    u.y = pack_half2(x[1].x, x[1].y);
    u.z = pack_half2(x[2].x, x[2].y);
    u.w = pack_half2(x[3].x, x[3].y);
    if (valid) *reinterpret_cast<uint4*>(row_base + (((piece0 + static_cast<uint32_t>(i)) ^ sw) << 4)) = u;
  }
}

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, int M, int N, int K, typename Epi::Params ep,
                int reverse) {
  using SM = Gemm2Smem;
  constexpr int STAGES = kGemm2Stages;
  constexpr int BN = 256;
  constexpr bool kTmaOut = std::is_same<Epi, EpiF16>::value;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int num_m = (M + 255) / 256;
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + kGemmBK - 1) / kGemmBK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kTmaOut) tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);      // leader: arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);     // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * (Epi::kSplitCols ? 8 : 4));   // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int tt = reverse ? num_tiles - 1 - tile : tile;
        const int m_blk = tt / num_n, n_blk = tt % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * SM::kStageBytes;
          uint8_t* sb = sa + SM::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * SM::kStageBytes);
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * kGemmBK, m_blk * 256 + static_cast<int>(rank) * 128);
          tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * kGemmBK, n_blk * BN + static_cast<int>(rank) * 128);
          if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * SM::kStageBytes);
          const uint64_t adesc = umma_desc_k128(a_addr);
          const uint64_t bdesc = umma_desc_k128(a_addr + SM::kABytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_f16_ss_2sm(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                            idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2..9, both CTAs) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    if (Epi::kSplitCols || half == 0) {
      uint8_t* obuf = smem + SM::kOutOffset + (warp - 2) * (2 * kGemm2OutBytes);
      const uint32_t sw = static_cast<uint32_t>(lane & 7);
      int gs = 0;                         // 64-column blocks stored so far (smem double buffer)
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int tt = reverse ? num_tiles - 1 - tile : tile;
        const int m_blk = tt / num_n, n_blk = tt % num_n;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after_sync();
        const int m0 = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
        int col0 = 0, n_cols = BN;
        if (Epi::kSplitCols) { col0 = half * (BN / 2); n_cols = BN / 2; }
        TmemRow row{tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                    static_cast<uint32_t>(as * BN + col0)};
        if constexpr (kTmaOut) {
          if (ep.act != ACT_PROBE_SKIP) {
            // the TMEM load of chunk c+1 is in flight while chunk c is converted and stored
            uint32_t rb[2][32];
            tmem_ld_32x32_nowait(row.taddr, rb[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int h = c >> 1, cc = c & 1;
              uint8_t* buf = obuf + ((gs + h) & 1) * kGemm2OutBytes;
              const int n0 = n_blk * BN + col0 + h * 64;
              if (cc == 0) {
                if (lane == 0) bulk_wait_group_read<1>();   // the store two blocks ago has left buf
                __syncwarp();
              }
              tmem_ld_wait();
              if (c + 1 < 4) tmem_ld_32x32_nowait(row.taddr + static_cast<uint32_t>(c + 1) * 32u, rb[(c + 1) & 1]);
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[c & 1][i]);
              if (ep.act & ACT_PROBE_DIRECT) {     // A/B: registers -> st.global, no smem, no TMA
                const int m = m0 + lane;
                epi_f16_pack_chunk(v, ep.bias ? ep.bias + n0 + cc * 32 : nullptr, ep.act & 0x3f,
                                   reinterpret_cast<uint8_t*>(ep.out + static_cast<size_t>(m < M ? m : 0) * ep.ldo + n0),
                                   static_cast<uint32_t>(cc * 4), 0u, m < M);
                continue;
              }
              epi_f16_pack_chunk(v, ep.bias ? ep.bias + n0 + cc * 32 : nullptr, ep.act & 0x3f,
                                 buf + lane * 128, static_cast<uint32_t>(cc * 4), sw);
              if (cc == 1 && !(ep.act & ACT_PROBE_NOTMA)) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmO, buf, n0, m0);
                  bulk_commit_group();
                }
              }
            }
            gs += 2;
          }
        } else {
          Epi::run(ep, m0, M, n_blk * BN + col0, n_cols, row, nullptr, lane);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tempty_bar[as]);
          else mbar_arrive_remote(&tempty_bar[as], 0);
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
      if (kTmaOut && lane == 0) bulk_wait_group<0>();
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

template <class Epi>
int launch_gemm_tc2(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                    const typename Epi::Params& ep, cudaStream_t stream) {
  using SM = Gemm2Smem;
  SRB_REQUIRE(N % 256 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm2: bad shape N=%d K=%d",
              N, K);
  CUtensorMap tmA, tmB, tmO;
  if (int rc = make_tmap_f16_2d(&tmA, A, M, K, lda, 128)) return rc;
  if (int rc = make_tmap_f16_2d(&tmB, W, N, K, ldw, 128)) return rc;
  tmO = tmA;
  if constexpr (std::is_same<Epi, EpiF16>::value) {
    if (int rc = make_tmap_f16_2d(&tmO, ep.out, M, N, ep.ldo, 32)) return rc;
  }
  auto kern = gemm_tc2_kernel<Epi>;
  static uint64_t attr_devs = 0;          // one bit per CUDA device: function attributes are per device
  if (first_use_on_device(&attr_devs)) {
    SRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  }
  const int num_tiles = ((M + 255) / 256) * (N / 256);
  const int max_clusters = device_sm_count() / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  kern<<<2 * clusters, kGemmThreads, SM::kTotal, stream>>>(tmA, tmB, tmO, M, N, K, ep, traverse_reverse() ? 1 : 0);
  SRB_CUDA_OK(cudaGetLastError());
  note_launch(1);
  return 0;
}

}  // namespace srb
