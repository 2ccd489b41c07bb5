// sam_road_b200 :: 2-CTA tcgen05 GEMM with an in-place fp32 residual epilogue fed by TMA.
//
//   X[m, n] += A[m, :] . W[n, :] + bias[n]        (attention proj + shortcut, MLP lin2 + shortcut:
//                                                  image_encoder.py:179-180)
//
// Same main loop as gemm_tc2.cuh (cluster of two CTAs, UMMA M = 256, 256 x 256 cluster tile).  The
// residual stream is fp32, so for short K this GEMM is bound by the 8 B per output element the
// epilogue moves.  Here every epilogue warp streams its 32 x 128 slab of X as four 32 x 32 fp32 blocks
// (4 KB, 128B-swizzled) through a private 3-deep smem ring: the TMA load of block g+2 is in flight
// -- also across tile boundaries, before the accumulator of the next tile exists -- while block g is
// combined with the accumulator (one row per lane, conflict-free 16 B accesses thanks to the swizzle)
// and written back with a TMA store.  Loads and stores are whole 128 B lines.
#pragma once

#include "gemm_tc2.cuh"

namespace srb {

constexpr int kGemm2RBlockBytes = 32 * 32 * 4;

template <int STAGES, int NB>
struct Gemm2RSmem {
  static constexpr int kABytes = 128 * kGemmBK * 2;
  static constexpr int kStageBytes = 2 * kABytes;
  static constexpr int kBufOffset = STAGES * kStageBytes;                               // 1024-aligned
  static constexpr int kBarOffset = kBufOffset + kGemmEpiWarps * NB * kGemm2RBlockBytes;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};

template <int STAGES, int NB, bool kReduce>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc2_resid_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR,
                      int M, int N, int K, const float* __restrict__ bias, int load_rows, int reverse) {
  using SM = Gemm2RSmem<STAGES, NB>;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;                       // [8 warps][NB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + kGemmEpiWarps * NB);

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
    tma_prefetch_desc(&tmX);
    if (!kReduce) tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kGemmEpiWarps);
    }
    for (int s = 0; s < kGemmEpiWarps * NB; ++s) mbar_init(&rfull_bar[s], 1);
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
    const int q = warp & 3;            // TMEM lane quarter
    const int ew = warp - 2;
    const int half = ew >> 2;          // column half of the tile
    uint8_t* rbuf = smem + SM::kBufOffset + ew * (NB * kGemm2RBlockBytes);
    uint64_t* rfull = rfull_bar + ew * NB;
    const int my_tiles = cluster_id < num_tiles ? (num_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
    const int total = my_tiles * 4;    // 32x32 blocks this warp streams
    auto coords = [&](int g, int& x, int& y) {
      const int tile0 = cluster_id + (g >> 2) * num_clusters;
      const int tile = reverse ? num_tiles - 1 - tile0 : tile0;
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      y = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
      x = n_blk * BN + half * 128 + (g & 3) * 32;
    };
    if (!kReduce && lane == 0) {
      for (int g = 0; g < NB - 1 && g < total; ++g) {
        int x, y;
        coords(g, x, y);
        mbar_arrive_expect_tx(&rfull[g], kGemm2RBlockBytes);
        tma_load_2d(rbuf + g * kGemm2RBlockBytes, &tmR, &rfull[g], x, load_rows > 0 ? y % load_rows : y);
      }
    }
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint32_t rb[2][32];
    for (int t4 = 0; t4 < total; t4 += 4) {
#pragma unroll
     for (int c = 0; c < 4; ++c) {
      const int g = t4 + c;
      const int b = g % NB;
      int x, y;
      coords(g, x, y);
      if (c == 0) {
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after_sync();
      }
      // accumulator block c+1 is already on its way from TMEM while block c is processed
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                            static_cast<uint32_t>(as * BN + half * 128);
      if (c == 0) tmem_ld_32x32_nowait(trow, rb[0]);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[c & 1][i]);
      if (c < 3) tmem_ld_32x32_nowait(trow + static_cast<uint32_t>(c + 1) * 32u, rb[(c + 1) & 1]);
      float4* mine = reinterpret_cast<float4*>(rbuf + b * kGemm2RBlockBytes + lane * 128);
      const float4* bp = reinterpret_cast<const float4*>(bias + x);
      if (kReduce) {
        // X += (acc + bias): the addition to the shortcut happens in L2 (TMA reduce-add), the
        // epilogue never reads X.  Buffer b was last read by the reduce of block g - NB.
        if (lane == 0) bulk_wait_group_read<NB - 1>();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(bp + i);
          mine[static_cast<uint32_t>(i) ^ sw] =
              make_float4(v[4 * i] + bb.x, v[4 * i + 1] + bb.y, v[4 * i + 2] + bb.z, v[4 * i + 3] + bb.w);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmX, rbuf + b * kGemm2RBlockBytes, x, y);
          bulk_commit_group();
        }
      } else {
        mbar_wait(&rfull[b], static_cast<uint32_t>((g / NB) & 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4* sp = mine + (static_cast<uint32_t>(i) ^ sw);
          const float4 r = *sp;
          const float4 bb = __ldg(bp + i);
          // (acc + shortcut) + bias: the association of EpiF32, so both kernels agree bit for bit
          *sp = make_float4((v[4 * i] + r.x) + bb.x, (v[4 * i + 1] + r.y) + bb.y,
                            (v[4 * i + 2] + r.z) + bb.z, (v[4 * i + 3] + r.w) + bb.w);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmX, rbuf + b * kGemm2RBlockBytes, x, y);
          bulk_commit_group();
          const int gn = g + NB - 1;      // refill the buffer block g-1 was stored from
          if (gn < total) {
            bulk_wait_group_read<1>();
            int xn, yn;
            coords(gn, xn, yn);
            const int bn = gn % NB;
            mbar_arrive_expect_tx(&rfull[bn], kGemm2RBlockBytes);
            tma_load_2d(rbuf + bn * kGemm2RBlockBytes, &tmR, &rfull[bn], xn,
                        load_rows > 0 ? yn % load_rows : yn);
          }
        }
      }
      if (c == 3) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tempty_bar[as]);
          else mbar_arrive_remote(&tempty_bar[as], 0);
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
     }
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// X[M, ldx] (fp32, in place) += A.W^T + bias.   N % 256 == 0, bias != null.
// kReduce: X += A.W^T + bias (L2 reduce-add).  !kReduce: X = A.W^T + bias + R[m % load_rows] with R
// streamed through smem (R = X for the in-place shortcut, R = pos_embed for the patch embedding).
template <int STAGES, int NB, bool kReduce>
int launch_gemm_tc2_resid(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                          const float* bias, float* X, int ldx, cudaStream_t stream,
                          const float* R = nullptr, int ldr = 0, int load_rows = 0) {
  using SM = Gemm2RSmem<STAGES, NB>;
  SRB_REQUIRE(N % 256 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldx % 4 == 0,
              "gemm2r: bad shape N=%d K=%d", N, K);
  CUtensorMap tmA, tmB, tmX;
  if (int rc = make_tmap_f16_2d(&tmA, A, M, K, lda, 128)) return rc;
  if (int rc = make_tmap_f16_2d(&tmB, W, N, K, ldw, 128)) return rc;
  if (int rc = make_tmap_f32_2d(&tmX, X, M, N, ldx, 32)) return rc;
  CUtensorMap tmR = tmX;
  if (R != nullptr && R != X) {
    SRB_REQUIRE(!kReduce && load_rows > 0 && load_rows % 32 == 0, "gemm2r: bad addend rows %d", load_rows);
    if (int rc = make_tmap_f32_2d(&tmR, R, load_rows, N, ldr, 32)) return rc;
  } else {
    load_rows = 0;
  }
  auto kern = gemm_tc2_resid_kernel<STAGES, NB, kReduce>;
  static uint64_t attr_devs = 0;          // one bit per CUDA device: function attributes are per device
  if (first_use_on_device(&attr_devs)) {
    SRB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  }
  const int num_tiles = ((M + 255) / 256) * (N / 256);
  const int max_clusters = device_sm_count() / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  kern<<<2 * clusters, kGemmThreads, SM::kTotal, stream>>>(tmA, tmB, tmX, tmR, M, N, K, bias, load_rows,
                                                           traverse_reverse() ? 1 : 0);
  SRB_CUDA_OK(cudaGetLastError());
  note_launch(1);
  return 0;
}

}  // namespace srb
