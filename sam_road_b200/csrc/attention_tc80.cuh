// sam_road_b200 :: tcgen05 flash attention for head_dim 80 (ViT-H, 16 heads x 80; model.py:217).
//
// Same algorithm as attention_tc.cuh (read that header first: S = QK^T and O += PV on tcgen05, O in
// TMEM, online softmax with a lazily updated reference max, decomposed rel-pos bias through the
// extra projection T = Q [Rh;Rw]^T, window tokens through a 4-D TMA box with pad rows rewritten to
// the qkv bias).  What head_dim 80 changes:
//   * 80 fp16 = 160 B does not fit the 128 B swizzle atom, so every operand tile is two K-blocks of
//     64 columns: columns [0,64) and [64,128) of the head.  The second TMA box runs 48 columns past
//     the head (into the next head / the next q|k|v section, zero fill past the row end); QK^T and
//     the rel-pos projection use only its first 16-column k-step (5 k-steps of 16 = K 80), so those
//     columns never enter a result.  P V is computed for all 128 columns (two N = 64 UMMAs per
//     k-step); O columns [80,128) are finite garbage and are never read.
//   * shared memory then holds one softmax group (128 query rows) instead of two: a window unit
//     (image, window, head) loads K/V once and runs its one or two 128-row query slabs one after the
//     other; a global unit is (image, head, 128-row slab) with K/V blocks streamed through a 2-stage ring.
//   * TMEM: S [0,128) | O [128,256) | T [256,256+NTAB).
// 8 warps: 0 TMA producer, 1 MMA issuer (+ TMEM allocator), 2 idle, 3 window pad fix-up, 4-7 softmax.
#pragma once

#include "attention_tc.cuh"

namespace srb {

constexpr int kAtc80Threads = 256;
constexpr int kAtc80HD = 80;

template <bool kWindow, int WIN>
struct Atc80Smem {
  static constexpr int NTAB = (4 * WIN - 2 <= 64) ? 64 : 128;
  static constexpr int kQRows = kWindow ? 256 : 128;            // rows per Q k-block tile
  static constexpr int kQTile = kQRows * 128;
  static constexpr int kTabTile = NTAB * 128;
  static constexpr int kKVRows = kWindow ? 208 : 128;           // rows per K / V k-block tile
  static constexpr int kKVTile = kKVRows * 128;
  static constexpr int kKVStage = 4 * kKVTile;                  // K kb0 | K kb1 | V kb0 | V kb1
  static constexpr int kKVStages = kWindow ? 1 : 2;
  static constexpr int kOffQ = 0;
  static constexpr int kOffTab = 2 * kQTile;
  static constexpr int kOffKV = kOffTab + 2 * kTabTile;
  static constexpr int kOffP = kOffKV + kKVStages * kKVStage;   // 2 k-blocks x 16 KB
  static constexpr int kOffBar = kOffP + 32768;
  static constexpr int kBytes = kOffBar + 256 + 1024;
  static_assert(kBytes <= 232448, "attention_tc80: shared memory budget");
};

struct Atc80Unit {
  int b, head, wy, wx, slab;
};

template <bool kWindow, int WIN>
__device__ __forceinline__ Atc80Unit atc80_decode(int u, const AtcParams& p) {
  Atc80Unit r;
  if (p.reverse) u = p.num_units - 1 - u;
  r.wy = r.wx = r.slab = 0;
  r.head = u % p.heads; u /= p.heads;
  if constexpr (kWindow) {
    r.wx = u % p.nwin; u /= p.nwin;
    r.wy = u % p.nwin; u /= p.nwin;
  } else {
    constexpr int SLABS = (WIN * WIN) / 128;
    r.slab = u % SLABS; u /= SLABS;
  }
  r.b = u;
  return r;
}

// query slabs of a unit: a window has 196 rows = 2 slabs unless rows 128.. are all padding
template <bool kWindow, int WIN>
__device__ __forceinline__ int atc80_slabs(const Atc80Unit& un, const AtcParams& p) {
  if constexpr (kWindow) {
    const int ry = min(WIN, p.s - un.wy * WIN);
    return ry * WIN > 128 ? 2 : 1;
  }
  return 1;
}

template <bool kWindow, int WIN>
__global__ void __launch_bounds__(kAtc80Threads, 1)
attention_tc80_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTab,
                      AtcParams p) {
  constexpr int KEYS = WIN * WIN;
  constexpr int NBLK = (KEYS + 127) / 128;
  using SM = Atc80Smem<kWindow, WIN>;
  constexpr int NTAB = SM::NTAB;
  constexpr int HALF = NTAB / 2;
  constexpr int kLastKeys = KEYS - 128 * (NBLK - 1);
  constexpr int kLastMma = ((kLastKeys + 15) / 16) * 16;
  constexpr int kStages = SM::kKVStages;
  static_assert(2 * WIN - 1 <= HALF, "rel-pos table half too small");
  static_assert(kWindow || KEYS % 128 == 0, "global mode needs s*s % 128 == 0");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem + SM::kOffQ;
  uint8_t* sTab = smem + SM::kOffTab;
  uint8_t* sKV = smem + SM::kOffKV;
  uint8_t* sP = smem + SM::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* tab_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* kv_full = bars + 3;                 // [2]
  uint64_t* kv_empty = bars + 5;                // [2]
  uint64_t* kv_fixed = bars + 7;
  uint64_t* kv_seen = bars + 8;
  uint64_t* s_ready = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_ready = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint64_t* t_ready = bars + 13;
  uint64_t* t_free = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.s * p.s;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmTab);
    mbar_init(tab_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(kv_fixed, 1);
    mbar_init(kv_seen, 1);
    mbar_init(s_ready, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(pv_done, 1);
    mbar_init(t_ready, 1);
    mbar_init(t_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if constexpr (kWindow) {
    // rows >= 196 of the Q / K / V tiles are never written by TMA: zero them once
    for (int i = threadIdx.x; i < 2 * (256 - KEYS) * 8; i += kAtc80Threads) {
      const int kb = i / ((256 - KEYS) * 8);
      *reinterpret_cast<uint4*>(sQ + kb * SM::kQTile + KEYS * 128 + (i % ((256 - KEYS) * 8)) * 16) =
          make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < 4 * (208 - KEYS) * 8; i += kAtc80Threads) {
      const int part = i / ((208 - KEYS) * 8);
      *reinterpret_cast<uint4*>(sKV + part * SM::kKVTile + KEYS * 128 + (i % ((208 - KEYS) * 8)) * 16) =
          make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmS = tmem_base, tmO = tmem_base + 128, tmT = tmem_base + 256;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // =========================== TMA producer ===========================
      mbar_arrive_expect_tx(tab_full, 2 * SM::kTabTile);
      tma_load_2d(sTab, &tmTab, tab_full, 0, 0);
      tma_load_2d(sTab + SM::kTabTile, &tmTab, tab_full, 64, 0);
      int uc = 0, qc = 0, gb = 0;      // units, Q loads, K/V blocks (global ring) so far
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++uc) {
        const Atc80Unit un = atc80_decode<kWindow, WIN>(u, p);
        const int colQ = un.head * kAtc80HD, colK = p.D + colQ, colV = 2 * p.D + colQ;
        if constexpr (kWindow) {
          mbar_wait(&kv_empty[0], static_cast<uint32_t>((uc & 1) ^ 1));
          mbar_arrive_expect_tx(&kv_full[0], 4 * KEYS * 128);
          const int x0 = un.wx * WIN, y0 = un.wy * WIN;
          tma_load_4d(sKV + 0 * SM::kKVTile, &tmQKV, &kv_full[0], colK, x0, y0, un.b);
          tma_load_4d(sKV + 1 * SM::kKVTile, &tmQKV, &kv_full[0], colK + 64, x0, y0, un.b);
          tma_load_4d(sKV + 2 * SM::kKVTile, &tmQKV, &kv_full[0], colV, x0, y0, un.b);
          tma_load_4d(sKV + 3 * SM::kKVTile, &tmQKV, &kv_full[0], colV + 64, x0, y0, un.b);
          mbar_wait(q_empty, static_cast<uint32_t>((qc & 1) ^ 1));
          mbar_arrive_expect_tx(q_full, 2 * KEYS * 128);
          tma_load_4d(sQ, &tmQKV, q_full, colQ, x0, y0, un.b);
          tma_load_4d(sQ + SM::kQTile, &tmQKV, q_full, colQ + 64, x0, y0, un.b);
          ++qc;
        } else {
          const int row0 = un.b * T;
          mbar_wait(q_empty, static_cast<uint32_t>((qc & 1) ^ 1));
          mbar_arrive_expect_tx(q_full, 2 * 128 * 128);
          tma_load_2d(sQ, &tmQKV, q_full, colQ, row0 + un.slab * 128);
          tma_load_2d(sQ + SM::kQTile, &tmQKV, q_full, colQ + 64, row0 + un.slab * 128);
          ++qc;
          for (int jb = 0; jb < NBLK; ++jb, ++gb) {
            const int stage = gb % kStages;
            mbar_wait(&kv_empty[stage], static_cast<uint32_t>(((gb / kStages) & 1) ^ 1));
            uint8_t* dst = sKV + stage * SM::kKVStage;
            mbar_arrive_expect_tx(&kv_full[stage], 4 * 128 * 128);
            tma_load_2d(dst + 0 * SM::kKVTile, &tmQKV, &kv_full[stage], colK, row0 + jb * 128);
            tma_load_2d(dst + 1 * SM::kKVTile, &tmQKV, &kv_full[stage], colK + 64, row0 + jb * 128);
            tma_load_2d(dst + 2 * SM::kKVTile, &tmQKV, &kv_full[stage], colV, row0 + jb * 128);
            tma_load_2d(dst + 3 * SM::kKVTile, &tmQKV, &kv_full[stage], colV + 64, row0 + jb * 128);
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // =========================== MMA issuer ===========================
      mbar_wait(tab_full, 0);
      int uc = 0, qc = 0, gb = 0;
      int wcnt = 0;      // S MMAs issued so far (s_free phases)
      int tcnt = 0;      // T projections issued so far (t_free phases)
      int bcnt = 0;      // blocks completed so far (p_ready phases)
      constexpr uint32_t idT = umma_idesc_f16(128, NTAB);
      constexpr uint32_t idPV = umma_idesc_f16_bmn(128, 64);
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++uc) {
        const Atc80Unit un = atc80_decode<kWindow, WIN>(u, p);
        const int nslab = atc80_slabs<kWindow, WIN>(un, p);
        mbar_wait(q_full, static_cast<uint32_t>(qc & 1));
        ++qc;
        if constexpr (kWindow) {
          mbar_wait(&kv_full[0], static_cast<uint32_t>(uc & 1));
          mbar_wait(kv_fixed, static_cast<uint32_t>(uc & 1));
          mbar_arrive(kv_seen);
        }
        tc_fence_after_sync();
        for (int slab = 0; slab < nslab; ++slab) {
          const uint32_t qa0 = smem_u32(sQ + slab * 16384), qa1 = smem_u32(sQ + SM::kQTile + slab * 16384);
          auto issue_s = [&](int jb) {
            uint32_t kb0, kb1;
            if constexpr (kWindow) {
              kb0 = smem_u32(sKV + jb * 16384);
              kb1 = smem_u32(sKV + SM::kKVTile + jb * 16384);
            } else {
              const int gbk = gb + jb;
              mbar_wait(&kv_full[gbk % kStages], static_cast<uint32_t>((gbk / kStages) & 1));
              kb0 = smem_u32(sKV + (gbk % kStages) * SM::kKVStage);
              kb1 = kb0 + SM::kKVTile;
            }
            if (wcnt > 0) mbar_wait(s_free, static_cast<uint32_t>((wcnt - 1) & 1));
            tc_fence_after_sync();
            const int nkeys = (kWindow && jb == NBLK - 1) ? kLastMma : 128;
            const uint32_t idS = umma_idesc_f16(128, nkeys);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmS, umma_desc_k128(qa0) + 2 * k, umma_desc_k128(kb0) + 2 * k, idS, k != 0 ? 1u : 0u);
            umma_f16_ss(tmS, umma_desc_k128(qa1), umma_desc_k128(kb1), idS, 1u);   // columns 64..79
            umma_commit(s_ready);
            wcnt++;
          };
          {   // rel-pos projection T = Q_slab * Tab^T
            if (tcnt > 0) {
              mbar_wait(t_free, static_cast<uint32_t>((tcnt - 1) & 1));
              tc_fence_after_sync();
            }
            const uint32_t tb0 = smem_u32(sTab), tb1 = smem_u32(sTab + SM::kTabTile);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmT, umma_desc_k128(qa0) + 2 * k, umma_desc_k128(tb0) + 2 * k, idT, k != 0 ? 1u : 0u);
            umma_f16_ss(tmT, umma_desc_k128(qa1), umma_desc_k128(tb1), idT, 1u);
            tcnt++;
            umma_commit(t_ready);
          }
          issue_s(0);
          for (int jb = 0; jb < NBLK; ++jb) {
            if (jb + 1 < NBLK) issue_s(jb + 1);
            else if (kWindow ? slab == nslab - 1 : true) umma_commit(q_empty);   // Q reusable after these MMAs
            // ---- O += P V (all 128 columns: two N = 64 halves per k-step) ----
            mbar_wait(p_ready, static_cast<uint32_t>((bcnt + jb) & 1));
            tc_fence_after_sync();
            const int nkeys = (kWindow && jb == NBLK - 1) ? kLastMma : 128;
            const uint32_t pbase = smem_u32(sP);
            uint32_t v0;
            if constexpr (kWindow) v0 = smem_u32(sKV + 2 * SM::kKVTile + jb * 16384);
            else v0 = smem_u32(sKV + ((gb + jb) % kStages) * SM::kKVStage + 2 * SM::kKVTile);
            const uint32_t v1 = v0 + SM::kKVTile;
            for (int k = 0; k < nkeys / 16; ++k) {
              const uint64_t adesc = umma_desc_k128(pbase + (k >> 2) * 16384) + 2 * (k & 3);
              umma_f16_ss(tmO, adesc, umma_desc_k128(v0 + k * 2048), idPV, (jb | k) != 0 ? 1u : 0u);
              umma_f16_ss(tmO + 64, adesc, umma_desc_k128(v1 + k * 2048), idPV, (jb | k) != 0 ? 1u : 0u);
            }
            umma_commit(pv_done);
            if constexpr (!kWindow) umma_commit(&kv_empty[(gb + jb) % kStages]);
          }
          bcnt += NBLK;
          if constexpr (!kWindow) gb += NBLK;
        }
        if constexpr (kWindow) umma_commit(&kv_empty[0]);
      }
    } else if (warp == 3) {
      // =========================== window pad fix-up ===========================
      if constexpr (kWindow) {
        int uc = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++uc) {
          const Atc80Unit un = atc80_decode<kWindow, WIN>(u, p);
          const int ry = min(WIN, p.s - un.wy * WIN), rx = min(WIN, p.s - un.wx * WIN);
          mbar_wait(&kv_full[0], static_cast<uint32_t>(uc & 1));
          if (ry < WIN || rx < WIN) {
            // lane -> (16 B piece c of the 80 columns: 10 pieces, K or V); pieces 0..7 live in the
            // kb0 tile, 8..9 in the kb1 tile
            for (int pc = lane; pc < 20; pc += 32) {
              const int c = pc % 10;
              const bool isv = pc >= 10;
              const float* bp = p.qkv_bias + (isv ? 2 : 1) * p.D + un.head * kAtc80HD + c * 8;
              uint4 ub;
              ub.x = pack_half2(__ldg(bp + 0), __ldg(bp + 1));
              ub.y = pack_half2(__ldg(bp + 2), __ldg(bp + 3));
              ub.z = pack_half2(__ldg(bp + 4), __ldg(bp + 5));
              ub.w = pack_half2(__ldg(bp + 6), __ldg(bp + 7));
              uint8_t* tile = sKV + ((isv ? 2 : 0) + (c >= 8 ? 1 : 0)) * SM::kKVTile;
              const int cc = c & 7;
              for (int r = 0; r < KEYS; ++r)
                if (r / WIN >= ry || r % WIN >= rx)
                  *reinterpret_cast<uint4*>(tile + r * 128 + ((cc ^ (r & 7)) << 4)) = ub;
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) {
            if (uc > 0) mbar_wait(kv_seen, static_cast<uint32_t>((uc - 1) & 1));
            mbar_arrive(kv_fixed);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== softmax group (128 rows) ===========================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmS + tlane, tO = tmO + tlane, tT = tmT + tlane;
    uint8_t* myP = sP + row * 128;
    // Gather scratch of the rel-pos projection: thread-private, inside the thread's own two P rows, rotated
    // by the lane (see attention_tc.cuh: a group-wide scratch over the P buffer raced with early P stores)
    auto scr = [&](int j) -> float* {
      return reinterpret_cast<float*>(myP + (j >> 5) * 16384) + ((j + lane) & 31);
    };
    const int sw = row & 7;
    constexpr float kLog2e = 1.4426950408889634f;
    int bcnt = 0, tcnt = 0;

    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      const Atc80Unit un = atc80_decode<kWindow, WIN>(u, p);
      const int nslab = atc80_slabs<kWindow, WIN>(un, p);
      for (int slab = 0; slab < nslab; ++slab) {
        const int qrow = slab * 128 + row;            // row inside the window (window mode)
        int qy, qx, ry = WIN, rx = WIN;
        bool q_real;
        size_t out_tok;
        if constexpr (kWindow) {
          qy = qrow / WIN; qx = qrow % WIN;
          ry = min(WIN, p.s - un.wy * WIN); rx = min(WIN, p.s - un.wx * WIN);
          q_real = qrow < KEYS && qy < ry && qx < rx;
          out_tok = static_cast<size_t>(un.b) * T + (un.wy * WIN + qy) * p.s + (un.wx * WIN + qx);
        } else {
          const int tok = un.slab * 128 + row;
          qy = tok / WIN; qx = tok % WIN;
          q_real = true;
          out_tok = static_cast<size_t>(un.b) * T + tok;
        }
        const uint32_t tpar = static_cast<uint32_t>(tcnt & 1);
        ++tcnt;
        mbar_wait(t_ready, tpar);
        tc_fence_after_sync();
        if constexpr (kWindow) {
          // a warp whose 32 rows are all window padding keeps the barrier protocol moving only
          const int r0 = slab * 128 + quarter * 32;
          if (r0 >= KEYS || r0 / WIN >= ry) {
            mbar_arrive(t_free);
            for (int jb = 0; jb < NBLK; ++jb) {
              const uint32_t par = static_cast<uint32_t>((bcnt + jb) & 1);
              mbar_wait(s_ready, par);
              mbar_arrive(s_free);
              if (bcnt + jb > 0) mbar_wait(pv_done, par ^ 1u);
              mbar_arrive(p_ready);
            }
            bcnt += NBLK;
            continue;
          }
        }

        // ---- rel-pos rows of this query, pre-multiplied by log2(e) ----
        float rel_h[WIN], rel_w[WIN];
        {
          const int sy = (kWindow && qrow >= KEYS) ? 0 : qy;
          const int sx = (kWindow && qrow >= KEYS) ? 0 : qx;
          // the scratch lives in the P buffer: PV of the previous slab / unit must have retired
          if (bcnt > 0) {
            mbar_wait(pv_done, static_cast<uint32_t>((bcnt - 1) & 1));
            tc_fence_after_sync();
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int c = 0; c < HALF / 32; ++c) {
              uint32_t r32[32];
              tmem_ld_32x32(tT + half * HALF + c * 32, r32);
#pragma unroll
              for (int i = 0; i < 32; ++i) *scr(c * 32 + i) = __uint_as_float(r32[i]);
            }
            if (half == 1) {
              tc_fence_before_sync();
              mbar_arrive(t_free);
            }
            const int sh = (half == 0 ? sy : sx) + WIN - 1;
#pragma unroll
            for (int i = 0; i < WIN; ++i) {
              const float v = *scr(sh - i) * kLog2e;
              if (half == 0) rel_h[i] = v; else rel_w[i] = v;
            }
          }
        }
        float2 relw2[WIN / 2];
#pragma unroll
        for (int i = 0; i < WIN / 2; ++i) relw2[i] = make_float2(rel_w[2 * i], rel_w[2 * i + 1]);

        float m_ref = 0.f, l_run = 0.f;
#pragma unroll(kWindow ? NBLK : 1)
        for (int jb = 0; jb < NBLK; ++jb) {
          constexpr int kChunksLast = (kLastKeys + 31) / 32;
          const int nchunk = (kWindow && jb == NBLK - 1) ? kChunksLast : 4;
          const uint32_t par = static_cast<uint32_t>((bcnt + jb) & 1);
          mbar_wait(s_ready, par);
          tc_fence_after_sync();
          uint32_t sraw[128];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < nchunk) tmem_ld_32x32_nowait(tS + c * 32, sraw + c * 32);
          tmem_ld_wait();
          tc_fence_before_sync();
          mbar_arrive(s_free);
          float2 y2[64];
          float m_blk = -INFINITY;
          float rhc[4];
          const float2 sl2 = make_float2(p.scale_log2e, p.scale_log2e);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            rhc[c] = 0.f;
            if (c < nchunk) {
              float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
              if constexpr (!kWindow) {
                constexpr int NU = (32 + WIN - 1) / WIN;
                float rh[NU];
#pragma unroll
                for (int uu = 0; uu < NU; ++uu) rh[uu] = rel_h[jb * (128 / WIN) + (c * 32) / WIN + uu];
                rhc[c] = rh[0];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int i = 2 * j;
                  float2 v = __ffma2_rn(make_float2(__uint_as_float(sraw[c * 32 + i]),
                                                    __uint_as_float(sraw[c * 32 + i + 1])),
                                        sl2, relw2[((c * 32 + i) % WIN) / 2]);
                  if (NU > 1 && i / WIN > 0) {
                    const float d = rh[NU - 1] - rh[0];
                    v = __fadd2_rn(v, make_float2(d, d));
                  }
                  y2[c * 16 + j] = v;
                  mq[j & 3] = fmax3(mq[j & 3], v.x, v.y);
                }
                m_blk = fmaxf(m_blk, fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3])) + rh[0]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int key = jb * 128 + c * 32 + 2 * j;
                  float2 v = make_float2(-INFINITY, -INFINITY);
                  if (key < KEYS) {
                    const float rh = rel_h[key / WIN];
                    v = __ffma2_rn(make_float2(__uint_as_float(sraw[c * 32 + 2 * j]),
                                               __uint_as_float(sraw[c * 32 + 2 * j + 1])),
                                   sl2, __fadd2_rn(make_float2(rh, rh), relw2[(key % WIN) / 2]));
                  }
                  y2[c * 16 + j] = v;
                  mq[j & 3] = fmax3(mq[j & 3], v.x, v.y);
                }
                m_blk = fmaxf(m_blk, fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3])));
              }
            }
          }
          // ---- lazy reference max ----
          bool pv_waited = false;
          if (jb == 0) {
            m_ref = m_blk;
          } else {
            const bool grow = m_blk > m_ref + 8.0f;
            if (__any_sync(0xffffffffu, grow)) {
              mbar_wait(pv_done, par ^ 1u);
              tc_fence_after_sync();
              pv_waited = true;
              const float m_new = grow ? m_blk : m_ref;
              const float alpha = ex2_approx(m_ref - m_new);
              m_ref = m_new;
              l_run *= alpha;
#pragma unroll
              for (int c = 0; c < 3; ++c) {            // columns 0..95 cover the 80 real ones
                uint32_t r32[32];
                tmem_ld_32x32(tO + c * 32, r32);
#pragma unroll
                for (int i = 0; i < 32; ++i) r32[i] = __float_as_uint(__uint_as_float(r32[i]) * alpha);
                tmem_st_32x32(tO + c * 32, r32);
              }
              tmem_st_wait();
            }
          }
          // ---- p = 2^(y - m_ref), packed to fp16, row sum ----
          uint32_t pk[64];
          float2 lsa = make_float2(0.f, 0.f), lsb = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nchunk) {
              const float sub = rhc[c] - m_ref;
              const float2 sub2 = make_float2(sub, sub);
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float2 a0 = __fadd2_rn(y2[c * 16 + j], sub2);
                const float2 a1 = __fadd2_rn(y2[c * 16 + j + 1], sub2);
                const float2 p0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
                const float2 p1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
                lsa = __fadd2_rn(lsa, p0);
                lsb = __fadd2_rn(lsb, p1);
                pk[c * 16 + j] = pack_half2(p0.x, p0.y);
                pk[c * 16 + j + 1] = pack_half2(p1.x, p1.y);
              }
            }
          }
          l_run += (lsa.x + lsa.y) + (lsb.x + lsb.y);
          // ---- P -> smem (swizzled) once the previous PV has finished reading the buffer ----
          if (bcnt + jb > 0 && !pv_waited) {
            mbar_wait(pv_done, par ^ 1u);
            tc_fence_after_sync();
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nchunk) {
              uint8_t* dst = myP + (c >> 1) * 16384;
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 uo = make_uint4(pk[c * 16 + q4 * 4 + 0], pk[c * 16 + q4 * 4 + 1],
                                            pk[c * 16 + q4 * 4 + 2], pk[c * 16 + q4 * 4 + 3]);
                const int piece = (c & 1) * 4 + q4;
                *reinterpret_cast<uint4*>(dst + ((piece ^ sw) << 4)) = uo;
              }
            }
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          mbar_arrive(p_ready);
        }

        // ---- epilogue of the slab: O / l -> global (80 columns) ----
        mbar_wait(pv_done, static_cast<uint32_t>((bcnt + NBLK - 1) & 1));
        tc_fence_after_sync();
        {
          const float inv = 1.0f / l_run;
          __half* op = p.out + out_tok * p.D + un.head * kAtc80HD;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            uint32_t r32[32];
            tmem_ld_32x32(tO + c * 32, r32);
            if (q_real) {
#pragma unroll
              for (int q4 = 0; q4 < (c < 2 ? 4 : 2); ++q4) {
                uint4 uo;
                uo.x = pack_half2(__uint_as_float(r32[q4 * 8 + 0]) * inv, __uint_as_float(r32[q4 * 8 + 1]) * inv);
                uo.y = pack_half2(__uint_as_float(r32[q4 * 8 + 2]) * inv, __uint_as_float(r32[q4 * 8 + 3]) * inv);
                uo.z = pack_half2(__uint_as_float(r32[q4 * 8 + 4]) * inv, __uint_as_float(r32[q4 * 8 + 5]) * inv);
                uo.w = pack_half2(__uint_as_float(r32[q4 * 8 + 6]) * inv, __uint_as_float(r32[q4 * 8 + 7]) * inv);
                *reinterpret_cast<uint4*>(op + c * 32 + q4 * 8) = uo;
              }
            }
          }
          tc_fence_before_sync();
        }
        bcnt += NBLK;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace srb
