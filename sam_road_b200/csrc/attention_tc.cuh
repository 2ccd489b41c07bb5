// sam_road_b200 :: tcgen05 flash attention for the ViT encoder (head_dim 64), window and global,
// with the decomposed relative-position bias of image_encoder.py:325-361 fused in.
//
// Persistent CTAs (grid = #SMs) loop over work units (image, window, head) / (image, head, 256-query
// slab).  12 warps = 3 warpgroups:
//   warp 0        TMA producer: rel-pos table (once), per unit Q (256 rows) and K/V blocks of 128 keys
//   warp 1        TMEM allocator; lane 0 issues the tcgen05.mma of softmax group 0 (blocking waits)
//   warp 2        lane 0 issues the tcgen05.mma of softmax group 1
//   warp 3        window mode: rewrites the K/V rows of window-pad tokens with the qkv bias, one unit
//                 ahead of the softmax groups (K/V is double-buffered across units); idle otherwise
//   warps 4-7     softmax group 0 (query rows   0..127 of the unit, one row per thread)
//   warps 8-11    softmax group 1 (query rows 128..255)
//
// Per 128-key block and group:  S = Q K^T (UMMA 128x128x16, fp32 in TMEM) -> the thread loads its
// whole S row into registers with ONE sweep of tcgen05.ld (the S buffer is released immediately, so
// the next block's QK^T overlaps this block's softmax) -> logits in the log2 domain
// y = s*scale*log2e + (rel_h[q,kh] + rel_w[q,kw])*log2e -> p = 2^(y - m_ref) with a lazily updated
// reference max (O and l are rescaled only when the row max grows by more than 2^8; exact, since
// m_ref cancels in O/l) -> P (fp16) into 128B-swizzled smem -> O += P V accumulated in TMEM
// (UMMA, V consumed MN-major straight from its [key][hd] TMA tile).  O is read once per unit.
// The attention matrix never touches HBM.
//
// Rel-pos: T = Q * [rel_pos_h ; rel_pos_w]^T is one extra UMMA per unit (64 or 128 table rows); each
// thread gathers rel_h[kh] = T_h[qh-kh+K-1], rel_w[kw] = T_w[qw-kw+K-1] for its row through a smem
// scratch (get_rel_pos, image_encoder.py:292-322).  The rel-pos term uses the UNscaled q, as the
// reference (image_encoder.py:231-234).
//
// Window mode never materialises window_partition (image_encoder.py:243-264): a 4-D TMA box
// [64 ch, 14 x, 14 y, 1 img] pulls the window's tokens; out-of-image tokens arrive as zeros and are
// overwritten in smem with the qkv bias (pad tokens have q=k=v=bias because padding follows norm1,
// image_encoder.py:168-172).  Their query rows are never stored; a softmax group
// whose 128 rows are all padding skips the unit, a warp whose 32 rows are all padding only keeps the
// barrier protocol moving.
//
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) and, for 64-row rel-pos tables, T0 [384,448)
// T1 [448,512).  In global mode the two groups take turns on the MUFU (e_done barriers): while one
// group exponentiates a block the other loads / maximises / stores (tools/att_trace.py: -6 %).
#pragma once

#include "common.cuh"
#include "ops.h"

namespace srb {

constexpr int kAtcThreads = 384;

// shared memory map (bytes from a 1024-aligned base)
//   global: Q 32 KB | Tab 16 KB | K/V ring 3 x (16 K + 16 K) | P 64 KB
//   window: Q 32 KB | Tab 16 KB | K/V 2 x (K 208 rows + V 208 rows) double-buffered across units | P 64 KB
template <bool kWindow, int kTabBytes = 16384>
struct AtcSmem {
  static constexpr int kQStages = 1;
  static constexpr int kKVStages = kWindow ? 2 : 3;
  static constexpr int kOffQ = 0;
  static constexpr int kOffTab = kQStages * 32768;
  static constexpr int kOffKV = kOffTab + kTabBytes;      // 16 KB table (<= 128 rows) or 32 KB (256 rows, 64x64 grid)
  static constexpr int kKVHalf = kWindow ? 208 * 128 : 16384;     // bytes of the K (or V) part of a stage
  static constexpr int kKVStage = 2 * kKVHalf;
  static constexpr int kOffP = kOffKV + kKVStages * kKVStage;   // 2 groups x 2 k-blocks x 16 KB
  static constexpr int kOffBar = kOffP + 65536;
  static constexpr int kBytes = kOffBar + 512 + 1024;
};

struct AtcParams {
  const float* qkv_bias;   // [3D] fp32
  __half* out;             // [B*s*s, D]
  int B, s, heads, D;
  int nwin;                // windows per side (window mode)
  int num_units;
  float scale_log2e;       // head_dim^-0.5 * log2(e)
  int no_stagger;          // A/B hook: do not delay group 1 by half a block
  int alternate;           // A/B hook: strict alternation of the groups' exponential phases
  int reverse;             // units in descending order
  long long* trace;        // debug: 256 clock64 stamps of CTA 0 (softmax warp 4: 8 per block; MMA threads; unit phases), or null
};

struct AtcUnit {
  int b, head, wy, wx, slab;
};

// rows of the packed rel-pos table [rel_pos_h ; rel_pos_w]: each half must hold 2*WIN-1 rows
__host__ __device__ constexpr int atc_table_rows(int win) {
  return 4 * win - 2 <= 64 ? 64 : (4 * win - 2 <= 128 ? 128 : 256);
}
__host__ __device__ constexpr int atc_table_bytes(int win) { return atc_table_rows(win) == 256 ? 32768 : 16384; }

template <bool kWindow, int WIN>
__device__ __forceinline__ AtcUnit atc_decode(int u, const AtcParams& p) {
  AtcUnit r;
  if (p.reverse) u = p.num_units - 1 - u;      // walk the images backwards (set_traverse_reverse)
  r.wy = r.wx = r.slab = 0;
  if constexpr (kWindow) {
    r.head = u % p.heads; u /= p.heads;
    r.wx = u % p.nwin; u /= p.nwin;
    r.wy = u % p.nwin; u /= p.nwin;
  } else {
    constexpr int SLABS = (WIN * WIN) / 256;
    r.slab = u % SLABS; u /= SLABS;
    r.head = u % p.heads; u /= p.heads;
  }
  r.b = u;
  return r;
}

// window mode: group 1 (query rows 128..195) is all padding when the window has <= 9 real rows
template <bool kWindow, int WIN>
__device__ __forceinline__ bool atc_group1_active(const AtcUnit& un, const AtcParams& p) {
  if constexpr (kWindow) {
    const int ry = min(WIN, p.s - un.wy * WIN);
    return ry * WIN > 128;
  }
  return true;
}

template <bool kWindow, int WIN, bool kPoly>
__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTab,
                    AtcParams p) {
  constexpr int KEYS = WIN * WIN;
  constexpr int NBLK = (KEYS + 127) / 128;
  constexpr int NTAB = atc_table_rows(WIN);
  constexpr int HALF = NTAB / 2;
  // 64x64 token grid: the table has 2 x 127 rows, so the projection T = Q Tab^T is 256 columns wide.  It goes
  // through the group's 128-column S region in two phases (rel_pos_h half, then rel_pos_w half).
  constexpr bool kTwoPhaseT = NTAB == 256;
  constexpr int kLastKeys = KEYS - 128 * (NBLK - 1);            // real keys in the last block
  constexpr int kLastMma = ((kLastKeys + 15) / 16) * 16;        // keys the last UMMA covers
  static_assert(2 * WIN - 1 <= HALF, "rel-pos table half too small");
  static_assert(kWindow || KEYS % 256 == 0, "global mode needs s*s % 256 == 0");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  using SM = AtcSmem<kWindow, atc_table_bytes(WIN)>;
  constexpr int QST = SM::kQStages;
  constexpr int kAtcKVStages = SM::kKVStages;
  constexpr int KVH = SM::kKVHalf;
  uint8_t* sQ = smem + SM::kOffQ;
  uint8_t* sTab = smem + SM::kOffTab;
  uint8_t* sKV = smem + SM::kOffKV;
  uint8_t* sP = smem + SM::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* tab_full = bars + 0;
  uint64_t* kv_fixed = bars + 2;
  uint64_t* q_full = bars + 3;                  // [2]
  uint64_t* q_empty = bars + 5;                 // [2]
  uint64_t* kv_full = bars + 7;                 // [<=3]
  uint64_t* kv_empty = bars + 10;               // [<=3]
  uint64_t* s_ready = bars + 13;                // [2]
  uint64_t* s_free = bars + 15;                 // [2]
  uint64_t* p_ready = bars + 17;                // [2]
  uint64_t* pv_done = bars + 19;                // [2]
  uint64_t* t_ready = bars + 21;                // [2]
  uint64_t* stagger = bars + 23;                // group 0 -> group 1 issuer: half-block phase offset
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* kv_seen = bars + 25;                // both issuers are past kv_fixed(ui) (paces warp 3)
  uint64_t* t_free = bars + 26;                 // [2] rel-pos projection T_g gathered (kSepT)
  uint64_t* e_done = bars + 28;                 // [2] group g finished the exponentials of a block

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.s * p.s;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmTab);
    mbar_init(tab_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2); mbar_init(&t_ready[i], 1); }
    mbar_init(kv_fixed, 1);
    mbar_init(kv_seen, 2);
    mbar_init(stagger, 128);
    for (int i = 0; i < 2; ++i) { mbar_init(&t_free[i], 128); mbar_init(&e_done[i], 128); }
    for (int i = 0; i < kAtcKVStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if constexpr (kWindow) {
    // rows >= 196 of the Q/K/V tiles are never written by TMA: zero them once (V rows must be
    // finite because P = 0 there; Q/K rows only feed masked / unstored entries)
    for (int i = threadIdx.x; i < (256 - KEYS) * 8; i += kAtcThreads)
      *reinterpret_cast<uint4*>(sQ + KEYS * 128 + i * 16) = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (208 - KEYS) * 8 * 2 * kAtcKVStages; i += kAtcThreads) {
      const int part = i / ((208 - KEYS) * 8);          // (stage, K|V)
      const int off = part * KVH + KEYS * 128 + (i % ((208 - KEYS) * 8)) * 16;
      *reinterpret_cast<uint4*>(sKV + off) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384); a 64-row rel-pos table gets its
  // own T0 [384,448) T1 [448,512), so the projection of the next unit and its first QK^T no longer
  // queue behind the gather (a 128-row table, global 32x32, shares the S region)
  constexpr bool kSepT = NTAB == 64;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // =========================== TMA producer ===========================
      mbar_arrive_expect_tx(tab_full, NTAB * 128);
      tma_load_2d(sTab, &tmTab, tab_full, 0, 0);
      int ui = 0;
      int gb = 0;                                  // running K/V block counter (global mode ring)
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++ui) {
        const AtcUnit un = atc_decode<kWindow, WIN>(u, p);
        const int colQ = un.head * 64, colK = p.D + un.head * 64, colV = 2 * p.D + un.head * 64;
        const int qs = ui % QST;
        const uint32_t qpar = static_cast<uint32_t>((ui / QST) & 1);
        uint8_t* qdst = sQ + qs * 32768;
        if constexpr (kWindow) {
          // K/V of the next unit first (longest chain: TMA -> pad fix-up -> first QK^T)
          const int ks = ui % kAtcKVStages;
          mbar_wait(&kv_empty[ks], (((ui / kAtcKVStages) & 1) ^ 1u));
          mbar_arrive_expect_tx(&kv_full[ks], 2 * KEYS * 128);
          tma_load_4d(sKV + ks * SM::kKVStage, &tmQKV, &kv_full[ks], colK, un.wx * WIN, un.wy * WIN, un.b);
          tma_load_4d(sKV + ks * SM::kKVStage + KVH, &tmQKV, &kv_full[ks], colV, un.wx * WIN,
                      un.wy * WIN, un.b);
          mbar_wait(&q_empty[qs], qpar ^ 1u);
          mbar_arrive_expect_tx(&q_full[qs], KEYS * 128);
          tma_load_4d(qdst, &tmQKV, &q_full[qs], colQ, un.wx * WIN, un.wy * WIN, un.b);
        } else {
          const int row0 = un.b * T;
          mbar_wait(&q_empty[qs], qpar ^ 1u);
          mbar_arrive_expect_tx(&q_full[qs], 256 * 128);
          tma_load_2d(qdst, &tmQKV, &q_full[qs], colQ, row0 + un.slab * 256);
          tma_load_2d(qdst + 16384, &tmQKV, &q_full[qs], colQ, row0 + un.slab * 256 + 128);
          for (int jb = 0; jb < NBLK; ++jb, ++gb) {
            const int stage = gb % kAtcKVStages;
            mbar_wait(&kv_empty[stage], ((gb / kAtcKVStages) & 1) ^ 1u);
            uint8_t* dst = sKV + stage * SM::kKVStage;
            mbar_arrive_expect_tx(&kv_full[stage], 32768);
            tma_load_2d(dst, &tmQKV, &kv_full[stage], colK, row0 + jb * 128);
            tma_load_2d(dst + 16384, &tmQKV, &kv_full[stage], colV, row0 + jb * 128);
          }
        }
      }
    } else if ((warp == 1 || warp == 2) && lane == 0) {
      // =========================== MMA issuers: one thread per softmax group ===========================
      // For one group the order of events is fixed -- S(0), then for every block: S(jb+1) once the
      // group has pulled S(jb) into registers (s_free), PV(jb) once P(jb) is in smem (p_ready) -- so
      // each issuer uses plain blocking waits; nothing polls.
      const int g = warp - 1;
      mbar_wait(tab_full, 0);
      int ui = 0;
      int wcnt = 0;      // writes (T or S) issued into this group's S region so far
      int tcnt = 0;      // rel-pos projections issued into this group's own T region (kSepT)
      int bcnt = 0;      // blocks completed by this group in earlier units
      constexpr uint32_t idT = umma_idesc_f16(128, kTwoPhaseT ? 128 : NTAB);
      constexpr uint32_t idPV = umma_idesc_f16_bmn(128, 64);
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++ui) {
        const AtcUnit un = atc_decode<kWindow, WIN>(u, p);
        const bool active = g == 0 || atc_group1_active<kWindow, WIN>(un, p);
        const int qs = ui % QST;
        const uint8_t* sQu = sQ + qs * 32768;
        const int ks = ui % kAtcKVStages;                  // window mode: K/V stage of this unit
        const int gb0 = ui * NBLK;                         // global mode: ring position of block 0
        mbar_wait(&q_full[qs], (ui / QST) & 1);
        if (!active) {     // nothing to compute, but keep the 2-arrival barriers and the pacing moving
          mbar_arrive(&t_ready[g]);
          mbar_arrive(&q_empty[qs]);
          if constexpr (kWindow) {
            mbar_wait(kv_fixed, ui & 1);
            mbar_arrive(kv_seen);
            mbar_arrive(&kv_empty[ks]);
          }
          continue;
        }
        tc_fence_after_sync();
        auto wait_s_region = [&]() {       // previous contents of the S region consumed?
          if (wcnt > 0) {
            mbar_wait(&s_free[g], (wcnt - 1) & 1);
            tc_fence_after_sync();
          }
        };
        auto issue_s = [&](int jb) {
          if constexpr (kWindow) {
            if (jb == 0) mbar_wait(&kv_full[ks], (ui / kAtcKVStages) & 1);
          } else {
            const int gbk = gb0 + jb;
            mbar_wait(&kv_full[gbk % kAtcKVStages], (gbk / kAtcKVStages) & 1);
          }
          wait_s_region();
          tc_fence_after_sync();
          const int nkeys = (kWindow && jb == NBLK - 1) ? kLastMma : 128;
          const uint32_t kbase = kWindow ? smem_u32(sKV + ks * SM::kKVStage + jb * 16384)
                                         : smem_u32(sKV + ((gb0 + jb) % kAtcKVStages) * SM::kKVStage);
          const uint32_t idS = umma_idesc_f16(128, nkeys);
          const uint64_t adesc = umma_desc_k128(smem_u32(sQu + g * 16384));
          const uint64_t bdesc = umma_desc_k128(kbase);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + g * 128, adesc + 2 * k, bdesc + 2 * k, idS, k != 0 ? 1u : 0u);
          umma_commit(&s_ready[g]);
          if (p.trace && blockIdx.x == 0 && ui == 0) p.trace[64 + jb * 2 + g] = clock64();
          wcnt++;
          if (jb == NBLK - 1) umma_commit(&q_empty[qs]);   // Q tile reusable once these MMAs retire
        };
        if constexpr (kTwoPhaseT) {
          // T_h = Q_g * rel_pos_h^T, then (once the group has gathered it) T_w = Q_g * rel_pos_w^T, both
          // through the S region; t_ready completes twice per unit (phases 2*ui and 2*ui+1)
          constexpr uint32_t idT128 = umma_idesc_f16(128, 128);
          const uint64_t adesc = umma_desc_k128(smem_u32(sQu + g * 16384));
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            wait_s_region();
            const uint64_t bdesc = umma_desc_k128(smem_u32(sTab + half * 16384));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem_base + g * 128, adesc + 2 * k, bdesc + 2 * k, idT128, k != 0 ? 1u : 0u);
            wcnt++;
            umma_commit(&t_ready[g]);
          }
        } else {   // rel-pos projection T_g = Q_g * Tab^T (own TMEM region, or the S region)
          if constexpr (kSepT) {
            if (tcnt > 0) {
              mbar_wait(&t_free[g], (tcnt - 1) & 1);
              tc_fence_after_sync();
            }
          } else {
            wait_s_region();
          }
          const uint64_t bdesc = umma_desc_k128(smem_u32(sTab));
          const uint64_t adesc = umma_desc_k128(smem_u32(sQu + g * 16384));
          const uint32_t tT = kSepT ? tmem_base + 384 + g * 64 : tmem_base + g * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tT, adesc + 2 * k, bdesc + 2 * k, idT, k != 0 ? 1u : 0u);
          if constexpr (kSepT) tcnt++; else wcnt++;
          umma_commit(&t_ready[g]);
        }
        if constexpr (kWindow) {
          mbar_wait(kv_fixed, ui & 1);
          mbar_arrive(kv_seen);
          tc_fence_after_sync();
        }
        // A/B hook (off by default, superseded by the turn-taking below): start group 1 half a block
        // behind group 0 so that one group's exponentials overlap the other's TMEM loads / max / P
        // stores (8-block global units only)
        if (g == 1 && !kWindow && !p.no_stagger) mbar_wait(stagger, ui & 1);
        issue_s(0);
        for (int jb = 0; jb < NBLK; ++jb) {
          if (jb + 1 < NBLK) issue_s(jb + 1);
          // ---- O += P V ----
          mbar_wait(&p_ready[g], (bcnt + jb) & 1);
          tc_fence_after_sync();
          const int nkeys = (kWindow && jb == NBLK - 1) ? kLastMma : 128;
          const uint32_t pbase = smem_u32(sP + g * 32768);
          const uint32_t vbase = kWindow ? smem_u32(sKV + ks * SM::kKVStage + KVH + jb * 16384)
                                         : smem_u32(sKV + ((gb0 + jb) % kAtcKVStages) * SM::kKVStage + KVH);
          const uint32_t d = tmem_base + 256 + g * 64;
          for (int k = 0; k < nkeys / 16; ++k) {
            const uint64_t adesc = umma_desc_k128(pbase + (k >> 2) * 16384) + 2 * (k & 3);
            const uint64_t bdesc = umma_desc_k128(vbase + k * 2048);
            umma_f16_ss(d, adesc, bdesc, idPV, (jb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[g]);
          if (p.trace && blockIdx.x == 0 && ui == 0) p.trace[96 + jb * 2 + g] = clock64();
          if constexpr (!kWindow) umma_commit(&kv_empty[(gb0 + jb) % kAtcKVStages]);   // 1 of 2 arrivals
        }
        if constexpr (kWindow) umma_commit(&kv_empty[ks]);
        bcnt += NBLK;
      }
    }
    else if (warp == 3) {
      // =========================== window pad fix-up (warp 3) ===========================
      // pad tokens of the window: k = b_k, v = b_v (fp16) written into the swizzled tiles once the
      // K/V TMA has landed (it zero-fills them); then the tiles are handed to the MMA issuers.  K/V
      // is double-buffered across units, so this runs a unit ahead of the softmax groups.
      if constexpr (kWindow) {
        int ui = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++ui) {
          const AtcUnit un = atc_decode<kWindow, WIN>(u, p);
          const int ry = min(WIN, p.s - un.wy * WIN), rx = min(WIN, p.s - un.wx * WIN);
          const int ks = ui % kAtcKVStages;
          uint8_t* kvs = sKV + ks * SM::kKVStage;
          mbar_wait(&kv_full[ks], (ui / kAtcKVStages) & 1);
          if (ry < WIN || rx < WIN) {
            const float* bk = p.qkv_bias + p.D + un.head * 64;
            const float* bv = p.qkv_bias + 2 * p.D + un.head * 64;
            // lane = 16 B piece (8 channels) x {K, V} x 2 row phases; the bias piece is converted once
            const int c = lane & 7;
            const bool isv = (lane >> 3) & 1;
            const float* bp = (isv ? bv : bk) + c * 8;
            uint4 ub;
            ub.x = pack_half2(__ldg(bp + 0), __ldg(bp + 1));
            ub.y = pack_half2(__ldg(bp + 2), __ldg(bp + 3));
            ub.z = pack_half2(__ldg(bp + 4), __ldg(bp + 5));
            ub.w = pack_half2(__ldg(bp + 6), __ldg(bp + 7));
            uint8_t* dstbase = kvs + (isv ? KVH : 0);
            for (int r = lane >> 4; r < KEYS; r += 2) {
              if (r / WIN >= ry || r % WIN >= rx)
                *reinterpret_cast<uint4*>(dstbase + r * 128 + ((c ^ (r & 7)) << 4)) = ub;
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) {
            // never more than one kv_fixed phase ahead of the issuers' waits
            if (ui > 0) mbar_wait(kv_seen, (ui - 1) & 1);
            mbar_arrive(kv_fixed);
          }
          __syncwarp();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =========================== softmax groups ===========================
    const int g = (warp - 4) >> 2;                 // 0 or 1
    const int quarter = warp & 3;                  // TMEM lane quarter of this warp
    const int row = quarter * 32 + lane;           // row inside the group's 128-row tile
    const int qrow = g * 128 + row;                // row inside the 256-row unit
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + tlane + g * 128;
    const uint32_t tO = tmem_base + tlane + 256 + g * 64;
    const uint32_t tT = tmem_base + tlane + 384 + g * 64;      // kSepT only
    uint8_t* myP = sP + g * 32768 + row * 128;
    // Gather scratch of the rel-pos projection: THREAD-PRIVATE, inside the thread's own two P rows (row
    // `row` of k-block 0 holds table rows 0..31, of k-block 1 rows 32..63), rotated by the lane so that a
    // warp's accesses hit 32 banks.  (A group-wide [HALF][128] scratch over the P buffer let a warp that is
    // late into a unit scribble over the P rows an early warp had already stored for the unit's first
    // block -- a rare, schedule-dependent corruption; window units with idle warps made the skew large.)
    auto scr = [&](int j) -> float* {
      return reinterpret_cast<float*>(myP + (j >> 5) * 16384) + ((j + lane) & 31);
    };
    const int sw = row & 7;
    constexpr float kLog2e = 1.4426950408889634f;
    int bcnt = 0;                                  // blocks completed by this group in earlier units
    int ui = 0;

    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++ui) {
      if (p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && ui < 8) p.trace[112 + ui] = clock64();
      const AtcUnit un = atc_decode<kWindow, WIN>(u, p);
      const bool active = g == 0 || atc_group1_active<kWindow, WIN>(un, p);
      int qy, qx, ry = WIN, rx = WIN;
      bool q_real;
      size_t out_tok;
      if constexpr (kWindow) {
        qy = qrow / WIN; qx = qrow % WIN;
        ry = min(WIN, p.s - un.wy * WIN); rx = min(WIN, p.s - un.wx * WIN);
        q_real = qrow < KEYS && qy < ry && qx < rx;
        out_tok = static_cast<size_t>(un.b) * T + (un.wy * WIN + qy) * p.s + (un.wx * WIN + qx);
        // every thread, also of an inactive group 1 (which has nothing else to wait for), consumes
        // this unit's t_ready phase: nobody can run a phase ahead of the barrier parities
        mbar_wait(&t_ready[g], ui & 1);
      } else {
        const int tok = un.slab * 256 + qrow;
        qy = tok / WIN; qx = tok % WIN;
        q_real = true;
        out_tok = static_cast<size_t>(un.b) * T + tok;
      }
      if (!active) continue;
      if constexpr (kWindow) {
        // A warp whose 32 query rows are all window padding (rows >= 196, or below the image edge)
        // produces no output: it keeps the barrier protocol moving and computes nothing.  Its P rows
        // keep whatever they held -- O rows are independent and these are never stored.
        const int r0 = g * 128 + quarter * 32;
        if (r0 >= KEYS || r0 / WIN >= ry) {
          mbar_arrive(&t_free[g]);
          for (int jb = 0; jb < NBLK; ++jb) {
            const uint32_t par = static_cast<uint32_t>((bcnt + jb) & 1);
            mbar_wait(&s_ready[g], par);
            mbar_arrive(&s_free[g]);
            if (jb > 0) mbar_wait(&pv_done[g], par ^ 1u);   // p_ready phase jb-1 is complete
            mbar_arrive(&p_ready[g]);
          }
          bcnt += NBLK;
          continue;
        }
      }

      // ---- rel-pos rows of this query, pre-multiplied by log2(e) ----
      // table rows: rel_pos_h at [0, 2K-1), rel_pos_w at [HALF, HALF + 2K-1)  (pack_rel_table)
      // rel_h[kh] = T_h[qy - kh + WIN-1] ; rel_w[kw] = T_w[qx - kw + WIN-1]  (image_encoder.py:318-322)
      float rel_h[WIN], rel_w[WIN];
      {
        if constexpr (!kWindow && !kTwoPhaseT) mbar_wait(&t_ready[g], ui & 1);
        tc_fence_after_sync();
        const int sy = (kWindow && qrow >= KEYS) ? 0 : qy;
        const int sx = (kWindow && qrow >= KEYS) ? 0 : qx;
        // a half of the table (HALF candidate rows per query) goes through the thread's 64-float scratch
        // in rounds of 64 columns; each rel_*[i] = T[sh - i] is picked up in the round that holds its column
        constexpr int kRound = HALF < 64 ? HALF : 64;
        constexpr int kRounds = HALF / kRound;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if constexpr (kTwoPhaseT) {
            mbar_wait(&t_ready[g], static_cast<uint32_t>(half));     // phases 2*ui + half
            tc_fence_after_sync();
          }
          const uint32_t tsrc = (kSepT ? tT : tS) + (kTwoPhaseT ? 0 : half * HALF);
          const int sh = (half == 0 ? sy : sx) + WIN - 1;
#pragma unroll
          for (int r = 0; r < kRounds; ++r) {
#pragma unroll
            for (int c = 0; c < kRound / 32; ++c) {
              uint32_t r32[32];
              tmem_ld_32x32(tsrc + r * 64 + c * 32, r32);
#pragma unroll
              for (int i = 0; i < 32; ++i) *scr(c * 32 + i) = __uint_as_float(r32[i]);
            }
            if (r == kRounds - 1 && (half == 1 || kTwoPhaseT)) {
              tc_fence_before_sync();
              mbar_arrive(kSepT ? &t_free[g] : &s_free[g]);   // T consumed: the region may be overwritten
            }
#pragma unroll
            for (int i = 0; i < WIN; ++i) {
              const int j = sh - i;
              if (kRounds == 1 || (j >> 6) == r) {
                const float v = *scr(j & 63) * kLog2e;
                if (half == 0) rel_h[i] = v; else rel_w[i] = v;
              }
            }
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
        const bool tr = p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && ui == 0;
        long long* trp = p.trace + jb * 8;
        if (tr) trp[0] = clock64();
        if (jb == 0 && p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && ui < 8) p.trace[128 + ui] = clock64();
        mbar_wait(&s_ready[g], par);
        if (tr) trp[1] = clock64();
        tc_fence_after_sync();
        // ---- one sweep: the whole S row into registers, then release the S buffer ----
        uint32_t sraw[128];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nchunk) tmem_ld_32x32_nowait(tS + c * 32, sraw + c * 32);
        tmem_ld_wait();
        if (tr) trp[2] = clock64();
        tc_fence_before_sync();
        mbar_arrive(&s_free[g]);
        // ---- logits (log2 domain) and block max; packed fp32x2 math (FFMA2/FADD2) halves the issue
        // slots of the element-wise work, which bounds this kernel together with the MUFU pipe ----
        float2 y2[64];
        float m_blk = -INFINITY;
        float rhc[4];      // global mode: per-chunk key-row constant, folded in after the max
        const float2 sl2 = make_float2(p.scale_log2e, p.scale_log2e);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          rhc[c] = 0.f;
          if (c < nchunk) {
            float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if constexpr (!kWindow) {
              // keys of chunk c of block jb: kh = jb*(128/WIN) + (c*32+i)/WIN, kw = (c*32+i) % WIN
              constexpr int NU = (32 + WIN - 1) / WIN;   // key rows spanned by a 32-key chunk (1 or 2)
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
                if (NU > 1 && i / WIN > 0) {             // WIN = 16: second key row of the chunk
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
                const int key = jb * 128 + c * 32 + 2 * j;         // even; KEYS and WIN are even
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
        if (tr) trp[3] = clock64() + static_cast<long long>(m_blk > 1e30f);
        // ---- lazy reference max: rescale O / l only when the row max grew by more than 2^8 ----
        bool pv_waited = false;
        if (jb == 0) {
          m_ref = m_blk;
        } else {
          const bool grow = m_blk > m_ref + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            mbar_wait(&pv_done[g], par ^ 1u);      // PV(jb-1) retired: O may be read-modified
            tc_fence_after_sync();
            pv_waited = true;
            const float m_new = grow ? m_blk : m_ref;
            const float alpha = ex2_approx(m_ref - m_new);
            m_ref = m_new;
            l_run *= alpha;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r32[32];
              tmem_ld_32x32(tO + c * 32, r32);
#pragma unroll
              for (int i = 0; i < 32; ++i) r32[i] = __float_as_uint(__uint_as_float(r32[i]) * alpha);
              tmem_st_32x32(tO + c * 32, r32);
            }
            tmem_st_wait();
          }
        }
        // ---- the two groups take turns on the MUFU: g0 b0, g1 b0, g0 b1, ... (global mode) ----
        if constexpr (!kWindow) {
          if (p.alternate) {
            const int n = bcnt + jb;                 // this group's block ordinal over the whole launch
            if (g == 0) { if (n > 0) mbar_wait(&e_done[1], static_cast<uint32_t>((n - 1) & 1)); }
            else mbar_wait(&e_done[0], static_cast<uint32_t>(n & 1));
          }
        }
        // ---- p = 2^(y - m_ref) packed to fp16 in registers, row sum (packed partial sums) ----
        uint32_t pk[64];
        float2 lsa = make_float2(0.f, 0.f), lsb = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nchunk) {
            const float sub = rhc[c] - m_ref;
            const float2 sub2 = make_float2(sub, sub);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float2 a0 = __fadd2_rn(y2[c * 16 + j], sub2);
              const float2 a1 = __fadd2_rn(y2[c * 16 + j + 1], sub2);
              const float2 a2 = __fadd2_rn(y2[c * 16 + j + 2], sub2);
              const float2 a3 = __fadd2_rn(y2[c * 16 + j + 3], sub2);
              const float2 p0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
              const float2 p1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
              const float2 p2 = make_float2(ex2_approx(a2.x), ex2_approx(a2.y));
              // kPoly: 1 in 4 on the (packed) FMA pipe instead of the MUFU
              const float2 p3 = kPoly ? ex2_poly3_2(a3) : make_float2(ex2_approx(a3.x), ex2_approx(a3.y));
              lsa = __fadd2_rn(lsa, __fadd2_rn(p0, p1));
              lsb = __fadd2_rn(lsb, __fadd2_rn(p2, p3));
              pk[c * 16 + j] = pack_half2(p0.x, p0.y);
              pk[c * 16 + j + 1] = pack_half2(p1.x, p1.y);
              pk[c * 16 + j + 2] = pack_half2(p2.x, p2.y);
              pk[c * 16 + j + 3] = pack_half2(p3.x, p3.y);
            }
            if (g == 0 && jb == 0 && c == 1) mbar_arrive(stagger);   // ~half a block into the unit
          }
        }
        if constexpr (!kWindow) { if (p.alternate) mbar_arrive(&e_done[g]); }
        l_run += (lsa.x + lsa.y) + (lsb.x + lsb.y);
        if (tr) trp[4] = clock64() + static_cast<long long>(l_run > 1e30f);
        // ---- P -> smem (swizzled) once PV(jb-1) has finished reading the buffer ----
        if (jb > 0 && !pv_waited) {
          mbar_wait(&pv_done[g], par ^ 1u);
          tc_fence_after_sync();
        }
        if (tr) trp[5] = clock64();
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
        if (tr) trp[6] = clock64();
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(&p_ready[g]);
        if (tr) trp[7] = clock64();
      }

      // ---- epilogue of the unit: O / l -> global ----
      if (p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && ui < 8) p.trace[120 + ui] = clock64();
      mbar_wait(&pv_done[g], static_cast<uint32_t>((bcnt + NBLK - 1) & 1));
      tc_fence_after_sync();
      if constexpr (kWindow) {
        // window units are bound by their latency chain: rows go straight from the registers to `out`
        // (the transposed path below was measured 15-17 % slower here, tools/gpu/att_ab2.sh)
        const float inv = 1.0f / l_run;
        __half* op = p.out + out_tok * p.D + un.head * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r32[32];
          tmem_ld_32x32(tO + c * 32, r32);
          if (q_real) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
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
      } else {
        // A row of O is 128 B = one cache line of `out`.  Stored straight from the registers, every 16 B
        // store instruction of a warp touches 32 different lines, and the load/store unit takes one cycle
        // per line: with 8 warps per CTA in their epilogue that is ~2 k cycles per 256-query unit.  So
        // the row goes through the thread's own P row (swizzled, conflict-free), and the warp reads its
        // 32 rows back transposed: 8 lanes cover one row, a store instruction touches 4 lines (global
        // attention -9 %).  Warp-local: only the warp's own rows of the (idle) P buffer are involved.
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r32[32];
          tmem_ld_32x32(tO + c * 32, r32);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 uo;
            uo.x = pack_half2(__uint_as_float(r32[q4 * 8 + 0]) * inv, __uint_as_float(r32[q4 * 8 + 1]) * inv);
            uo.y = pack_half2(__uint_as_float(r32[q4 * 8 + 2]) * inv, __uint_as_float(r32[q4 * 8 + 3]) * inv);
            uo.z = pack_half2(__uint_as_float(r32[q4 * 8 + 4]) * inv, __uint_as_float(r32[q4 * 8 + 5]) * inv);
            uo.w = pack_half2(__uint_as_float(r32[q4 * 8 + 6]) * inv, __uint_as_float(r32[q4 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(myP + (((c * 4 + q4) ^ sw) << 4)) = uo;
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        const int piece = lane & 7;
        const uint8_t* wP = sP + g * 32768 + quarter * 32 * 128;       // the warp's 32 rows
        __half* ob = p.out + (static_cast<size_t>(un.b) * T + un.slab * 256) * p.D + un.head * 64 + piece * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + (lane >> 3);                           // row inside the warp
          const int qr = g * 128 + quarter * 32 + rl;                   // row inside the unit
          const uint4 v = *reinterpret_cast<const uint4*>(wP + rl * 128 + ((piece ^ (rl & 7)) << 4));
          *reinterpret_cast<uint4*>(ob + static_cast<size_t>(qr) * p.D) = v;
        }
        __syncwarp();       // every lane has read the warp's rows: the next unit may overwrite them
      }
      if (p.trace && blockIdx.x == 0 && warp == 4 && lane == 0 && ui < 8) p.trace[136 + ui] = clock64();
      bcnt += NBLK;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace srb
