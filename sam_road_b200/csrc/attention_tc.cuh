// sam_road_b200 :: tcgen05 flash attention for the ViT encoder (head_dim 64), window and global,
// with the decomposed relative-position bias of image_encoder.py:325-361 fused in.
//
// One CTA = one (image, window, head[, 256-query slab]) work unit, 12 warps (3 warpgroups):
//   warp 0        TMA producer: rel-pos table, Q (256 rows), K/V blocks of 128 keys
//   warp 1        TMEM allocator + single-thread tcgen05.mma issuer (event driven)
//   warps 2-3     idle (complete the control warpgroup, which gives its registers away)
//   warps 4-7     softmax group 0 (query rows   0..127, one row per thread)
//   warps 8-11    softmax group 1 (query rows 128..255)
//
// Per 128-key block and per group:  S = Q K^T (UMMA 128x128x16, fp32 in TMEM) -> the group's threads
// read their S row with tcgen05.ld, form logits x = s*scale + rel_h[q,kh] + rel_w[q,kw], run the
// online softmax in fp32 registers, write P (fp16) into 128B-swizzled smem -> O_blk = P V (UMMA, V
// consumed MN-major straight from its [key][hd] TMA tile) into a TMEM scratch -> registers
// o = o*alpha + O_blk.  The attention matrix never touches HBM.
//
// Rel-pos: T = Q * [rel_pos_h ; rel_pos_w]^T is one extra UMMA per slab (N = 64 or 128 table rows);
// every thread then gathers rel_h[kh] = T[qh-kh+K-1], rel_w[kw] = T[(2K-1)+qw-kw+K-1] for its row
// (get_rel_pos, image_encoder.py:292-322) -- the rel-pos term uses the UNscaled q, as the reference.
//
// Window mode never materialises window_partition (image_encoder.py:243-264): a 4-D TMA box
// [64 ch, 14 x, 14 y, 1 img] pulls the window's tokens; out-of-image tokens arrive as zeros and are
// overwritten in smem with the qkv bias (pad tokens have q=k=v=bias because padding follows norm1,
// image_encoder.py:168-172).  Their query rows are never stored.
#pragma once

#include "common.cuh"
#include "ops.h"

namespace srb {

constexpr int kAtcThreads = 384;   // 3 warpgroups: {TMA, MMA, 2 idle}, softmax 0, softmax 1
constexpr int kAtcKVStages = 3;

// shared memory map (bytes from a 1024-aligned base)
constexpr int kAtcOffQ = 0;                       // 256 rows x 128 B
constexpr int kAtcOffTab = 32768;                 // <=128 rows x 128 B
constexpr int kAtcOffKV = 49152;                  // global: 3 x (K 16K + V 16K); window: K 32K + V 32K
constexpr int kAtcOffP = kAtcOffKV + 98304;       // 2 groups x 2 k-blocks x 16 KB
constexpr int kAtcOffBar = kAtcOffP + 65536;
constexpr int kAtcSmemBytes = kAtcOffBar + 512 + 1024;

struct AtcParams {
  const float* qkv_bias;   // [3D] fp32
  __half* out;             // [B*s*s, D]
  int B, s, heads, D;
  int nwin;                // windows per side (window mode)
  float scale;             // head_dim^-0.5
};

// gather rel[i] = t[shift + K-1-i], i in [0,K), from a register/local array via a log shifter
template <int K, int L>
__device__ __forceinline__ void shift_down(float (&t)[L], int shift) {
#pragma unroll
  for (int bit = 32; bit >= 1; bit >>= 1) {
    if (bit < K) {
      const bool on = (shift & bit) != 0;
#pragma unroll
      for (int j = 0; j + bit < L; ++j) t[j] = on ? t[j + bit] : t[j];
    }
  }
}

// MODE_WIN: WIN = 14 (keys = 196, two blocks 128 + 80[68 real]); MODE_GLOBAL: WIN = s (16 or 32)
template <bool kWindow, int WIN>
__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTab,
                    AtcParams p) {
  constexpr int KEYS = WIN * WIN;
  constexpr int NBLK = (KEYS + 127) / 128;
  constexpr int NTAB = (4 * WIN - 2 <= 64) ? 64 : 128;
  constexpr int LREL = 2 * WIN - 1;
  static_assert(4 * WIN - 2 <= 128, "rel-pos table too large for one UMMA");
  static_assert(kWindow || KEYS % 256 == 0, "global mode needs s*s % 256 == 0");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem + kAtcOffQ;
  uint8_t* sTab = smem + kAtcOffTab;
  uint8_t* sKV = smem + kAtcOffKV;
  uint8_t* sP = smem + kAtcOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtcOffBar);
  uint64_t* tab_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* t_ready = bars + 2;
  uint64_t* kv_fixed = bars + 3;
  uint64_t* kv_full = bars + 4;                 // [kAtcKVStages]
  uint64_t* kv_empty = bars + 4 + kAtcKVStages; // [kAtcKVStages]
  uint64_t* s_ready = bars + 10;                // [2]
  uint64_t* s_free = bars + 12;                 // [2]
  uint64_t* p_ready = bars + 14;                // [2]
  uint64_t* pv_done = bars + 16;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work unit ----
  int b, head, wy = 0, wx = 0, slab = 0;
  {
    int u = blockIdx.x;
    head = u % p.heads; u /= p.heads;
    if constexpr (kWindow) {
      wx = u % p.nwin; u /= p.nwin;
      wy = u % p.nwin; u /= p.nwin;
    } else {
      constexpr int SLABS = KEYS / 256;
      slab = u % SLABS; u /= SLABS;
    }
    b = u;
  }
  const int T = p.s * p.s;
  const int colQ = head * 64, colK = p.D + head * 64, colV = 2 * p.D + head * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmTab);
    mbar_init(tab_full, 1);
    mbar_init(q_full, 1);
    mbar_init(t_ready, 1);
    mbar_init(kv_fixed, 256);
    for (int i = 0; i < kAtcKVStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
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
    // finite because P = 0 there; Q/K garbage only feeds masked / unstored entries, but keep clean)
    for (int i = threadIdx.x; i < (256 - KEYS) * 8; i += kAtcThreads) {
      const int off = KEYS * 128 + i * 16;
      *reinterpret_cast<uint4*>(sQ + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sKV + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sKV + 32768 + off) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(tab_full, NTAB * 128);
      tma_load_2d(sTab, &tmTab, tab_full, 0, 0);
      if constexpr (kWindow) {
        mbar_arrive_expect_tx(q_full, KEYS * 128);
        tma_load_4d(sQ, &tmQKV, q_full, colQ, wx * WIN, wy * WIN, b);
        mbar_arrive_expect_tx(&kv_full[0], 2 * KEYS * 128);
        tma_load_4d(sKV, &tmQKV, &kv_full[0], colK, wx * WIN, wy * WIN, b);
        tma_load_4d(sKV + 32768, &tmQKV, &kv_full[0], colV, wx * WIN, wy * WIN, b);
      } else {
        const int row0 = b * T;
        mbar_arrive_expect_tx(q_full, 256 * 128);
        tma_load_2d(sQ, &tmQKV, q_full, colQ, row0 + slab * 256);
        tma_load_2d(sQ + 16384, &tmQKV, q_full, colQ, row0 + slab * 256 + 128);
        int stage = 0;
        uint32_t phase = 0;
        for (int jb = 0; jb < NBLK; ++jb) {
          mbar_wait(&kv_empty[stage], phase ^ 1u);
          uint8_t* dst = sKV + stage * 32768;
          mbar_arrive_expect_tx(&kv_full[stage], 32768);
          tma_load_2d(dst, &tmQKV, &kv_full[stage], colK, row0 + jb * 128);
          tma_load_2d(dst + 16384, &tmQKV, &kv_full[stage], colV, row0 + jb * 128);
          if (++stage == kAtcKVStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // rel-pos projections T_g = Q_g * Tab^T into the S regions
      mbar_wait(tab_full, 0);
      mbar_wait(q_full, 0);
      tc_fence_after_sync();
      {
        constexpr uint32_t idT = umma_idesc_f16(128, NTAB);
        const uint64_t bdesc = umma_desc_k128(smem_u32(sTab));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint64_t adesc = umma_desc_k128(smem_u32(sQ + g * 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + g * 128, adesc + 2 * k, bdesc + 2 * k, idT, k != 0 ? 1u : 0u);
        }
        umma_commit(t_ready);
      }
      if constexpr (kWindow) {
        mbar_wait(kv_fixed, 0);
        tc_fence_after_sync();
      }
      int ns[2] = {0, 0}, npv[2] = {0, 0};      // next S / PV block per group
      int kv_seen = 0;                          // number of K/V blocks whose full barrier was observed
      int released = 0;                         // number of K/V blocks released back to the producer
      while (npv[0] < NBLK || npv[1] < NBLK) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // ---- O_blk = P V for the oldest pending block of this group ----
          if (npv[g] < ns[g] && mbar_try_wait(&p_ready[g], npv[g] & 1)) {
            tc_fence_after_sync();
            const int jb = npv[g];
            const int nkeys = (kWindow && jb == NBLK - 1) ? (((KEYS - 128 * (NBLK - 1)) + 15) / 16) * 16 : 128;
            const uint32_t pbase = smem_u32(sP + g * 32768);
            const uint32_t vbase = kWindow ? smem_u32(sKV + 32768 + jb * 16384)
                                           : smem_u32(sKV + (jb % kAtcKVStages) * 32768 + 16384);
            constexpr uint32_t idPV = umma_idesc_f16_bmn(128, 64);
            const uint32_t d = tmem_base + 256 + g * 64;
            for (int k = 0; k < nkeys / 16; ++k) {
              const uint64_t adesc = umma_desc_k128(pbase + (k >> 2) * 16384) + 2 * (k & 3);
              const uint64_t bdesc = umma_desc_k128(vbase + k * 2048);
              umma_f16_ss(d, adesc, bdesc, idPV, k != 0 ? 1u : 0u);
            }
            umma_commit(&pv_done[g]);
            npv[g]++;
            if constexpr (!kWindow) {
              const int done = npv[0] < npv[1] ? npv[0] : npv[1];
              while (released < done) {         // both groups finished with block `released`
                umma_commit(&kv_empty[released % kAtcKVStages]);
                released++;
              }
            }
          }
          // ---- S = Q K^T for the next block of this group ----
          if (ns[g] < NBLK) {
            const int jb = ns[g];
            bool ok = mbar_try_wait(&s_free[g], jb & 1);
            if (ok && jb >= kv_seen) {
              const int st = kWindow ? 0 : jb % kAtcKVStages;
              const uint32_t par = kWindow ? 0u : static_cast<uint32_t>((jb / kAtcKVStages) & 1);
              if (kWindow && jb > 0) {
                kv_seen = NBLK;                 // one TMA filled the whole window
              } else if (mbar_try_wait(&kv_full[st], par)) {
                kv_seen = kWindow ? NBLK : jb + 1;
              } else {
                ok = false;
              }
            }
            if (ok) {
              tc_fence_after_sync();
              const int nkeys = (kWindow && jb == NBLK - 1) ? (((KEYS - 128 * (NBLK - 1)) + 15) / 16) * 16 : 128;
              const uint32_t kbase = kWindow ? smem_u32(sKV + jb * 16384)
                                             : smem_u32(sKV + (jb % kAtcKVStages) * 32768);
              const uint32_t idS = umma_idesc_f16(128, nkeys);
              const uint64_t adesc = umma_desc_k128(smem_u32(sQ + g * 16384));
              const uint64_t bdesc = umma_desc_k128(kbase);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss(tmem_base + g * 128, adesc + 2 * k, bdesc + 2 * k, idS, k != 0 ? 1u : 0u);
              umma_commit(&s_ready[g]);
              ns[g]++;
            }
          }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =========================== softmax groups ===========================
    const int g = (warp - 4) >> 2;                 // 0 or 1
    const int quarter = warp & 3;                  // TMEM lane quarter of this warp
    const int row = quarter * 32 + lane;           // row inside the group's 128-row tile
    const int qrow = g * 128 + row;                // row inside the 256-row slab
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + tlane + g * 128;
    const uint32_t tPV = tmem_base + tlane + 256 + g * 64;

    // query position
    int qy, qx, ry = WIN, rx = WIN;
    bool q_real;
    size_t out_tok;
    if constexpr (kWindow) {
      qy = qrow / WIN; qx = qrow % WIN;
      ry = min(WIN, p.s - wy * WIN); rx = min(WIN, p.s - wx * WIN);
      q_real = qrow < KEYS && qy < ry && qx < rx;
      out_tok = static_cast<size_t>(b) * T + (wy * WIN + qy) * p.s + (wx * WIN + qx);
    } else {
      const int tok = slab * 256 + qrow;
      qy = tok / WIN; qx = tok % WIN;
      q_real = true;
      out_tok = static_cast<size_t>(b) * T + tok;
    }

    // ---- rel-pos rows of this query ----
    float rel_h[WIN], rel_w[WIN];
    {
      mbar_wait(t_ready, 0);
      tc_fence_after_sync();
      // table rows: rel_pos_h at [0, LREL), rel_pos_w at [NTAB/2, NTAB/2 + LREL)  (pack_rel_table)
      // rel_h[kh] = T_h[qy - kh + WIN-1] ; rel_w[kw] = T_w[qx - kw + WIN-1]   (image_encoder.py:318-322)
      const int sy = (kWindow && qrow >= KEYS) ? 0 : qy;
      const int sx = (kWindow && qrow >= KEYS) ? 0 : qx;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float t[NTAB / 2];
#pragma unroll
        for (int c = 0; c < NTAB / 64; ++c) {
          uint32_t r32[32];
          tmem_ld_32x32(tS + half * (NTAB / 2) + c * 32, r32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) t[c * 32 + i] = __uint_as_float(r32[i]);
        }
        shift_down<WIN, NTAB / 2>(t, half == 0 ? sy : sx);
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          if (half == 0) rel_h[i] = t[WIN - 1 - i];
          else rel_w[i] = t[WIN - 1 - i];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&s_free[g]);                      // T consumed: S(0) may overwrite it
    }

    if constexpr (kWindow) {
      // pad tokens of the window: k = b_k, v = b_v (fp16) written into the swizzled tiles once the
      // K/V TMA has landed (it zero-fills them); then hand the tiles to the MMA warp
      mbar_wait(&kv_full[0], 0);
      const int r = (warp - 4) * 32 + lane;
      if (r < KEYS && (r / WIN >= ry || r % WIN >= rx)) {
        const float* bk = p.qkv_bias + p.D + head * 64;
        const float* bv = p.qkv_bias + 2 * p.D + head * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 uk, uv;
          uk.x = pack_half2(__ldg(bk + c * 8 + 0), __ldg(bk + c * 8 + 1));
          uk.y = pack_half2(__ldg(bk + c * 8 + 2), __ldg(bk + c * 8 + 3));
          uk.z = pack_half2(__ldg(bk + c * 8 + 4), __ldg(bk + c * 8 + 5));
          uk.w = pack_half2(__ldg(bk + c * 8 + 6), __ldg(bk + c * 8 + 7));
          uv.x = pack_half2(__ldg(bv + c * 8 + 0), __ldg(bv + c * 8 + 1));
          uv.y = pack_half2(__ldg(bv + c * 8 + 2), __ldg(bv + c * 8 + 3));
          uv.z = pack_half2(__ldg(bv + c * 8 + 4), __ldg(bv + c * 8 + 5));
          uv.w = pack_half2(__ldg(bv + c * 8 + 6), __ldg(bv + c * 8 + 7));
          const int off = r * 128 + ((c ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(sKV + off) = uk;
          *reinterpret_cast<uint4*>(sKV + 32768 + off) = uv;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(kv_fixed);
    }

    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    constexpr float kLog2e = 1.4426950408889634f;
    uint8_t* myP = sP + g * 32768 + row * 128;
    const int sw = row & 7;

#pragma unroll(kWindow ? NBLK : 1)
    for (int jb = 0; jb < NBLK; ++jb) {
      constexpr int kLastKeys = KEYS - 128 * (NBLK - 1);             // real keys in the last block
      const int nchunk = (kWindow && jb == NBLK - 1) ? (kLastKeys + 31) / 32 : 4;
      mbar_wait(&s_ready[g], jb & 1);
      tc_fence_after_sync();
      // ---------------- pass 1: row max of the logits ----------------
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunk) {
          uint32_t r32[32];
          tmem_ld_32x32(tS + c * 32, r32);
          tmem_ld_wait();
          float rhv[2] = {0.f, 0.f};     // global mode: the (1 or 2) key rows this chunk spans
          if constexpr (!kWindow) {
#pragma unroll
            for (int u = 0; u < (32 + WIN - 1) / WIN; ++u)
              rhv[u] = rel_h[jb * (128 / WIN) + (c * 32) / WIN + u];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = jb * 128 + c * 32 + i;
            if (kWindow && key >= KEYS) continue;
            const float bias = kWindow ? rel_h[key / WIN] + rel_w[key % WIN]
                                       : rhv[i / WIN] + rel_w[(c * 32 + i) % WIN];
            mx = fmaxf(mx, fmaf(__uint_as_float(r32[i]), p.scale, bias));
          }
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * kLog2e);
      const float mneg = -m_new * kLog2e;
      // ---------------- pass 2: p = exp(x - m), row sum, P -> smem (fp16, swizzled) ----------------
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunk) {
          uint32_t r32[32];
          tmem_ld_32x32(tS + c * 32, r32);
          tmem_ld_wait();
          float pv[32];
          float rhv[2] = {0.f, 0.f};
          if constexpr (!kWindow) {
#pragma unroll
            for (int u = 0; u < (32 + WIN - 1) / WIN; ++u)
              rhv[u] = rel_h[jb * (128 / WIN) + (c * 32) / WIN + u];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = jb * 128 + c * 32 + i;
            if (kWindow && key >= KEYS) {
              pv[i] = 0.f;
            } else {
              const float bias = kWindow ? rel_h[key / WIN] + rel_w[key % WIN]
                                         : rhv[i / WIN] + rel_w[(c * 32 + i) % WIN];
              const float x = fmaf(__uint_as_float(r32[i]), p.scale, bias);
              pv[i] = exp2f(fmaf(x, kLog2e, mneg));
              lsum += pv[i];
            }
          }
          // chunk c = 32 keys = 64 B = four 16-byte pieces of k-block (c>>1), pieces (c&1)*4 ..
          uint8_t* dst = myP + (c >> 1) * 16384;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 u;
            u.x = pack_half2(pv[q4 * 8 + 0], pv[q4 * 8 + 1]);
            u.y = pack_half2(pv[q4 * 8 + 2], pv[q4 * 8 + 3]);
            u.z = pack_half2(pv[q4 * 8 + 4], pv[q4 * 8 + 5]);
            u.w = pack_half2(pv[q4 * 8 + 6], pv[q4 * 8 + 7]);
            const int piece = (c & 1) * 4 + q4;
            *reinterpret_cast<uint4*>(dst + ((piece ^ sw) << 4)) = u;
          }
        } else if (kWindow && c * 32 < ((kLastKeys + 15) / 16) * 16) {
          // keys covered by the UMMA K extent but beyond the last 32-chunk: zero P
          uint8_t* dst = myP + (c >> 1) * 16384;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + q4) ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[g]);
      mbar_arrive(&s_free[g]);
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      // ---------------- O update ----------------
      mbar_wait(&pv_done[g], jb & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r32[32];
        tmem_ld_32x32(tPV + c * 32, r32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(r32[i]));
      }
      tc_fence_before_sync();
    }

    if (q_real) {
      const float inv = 1.0f / l_run;
      __half* op = p.out + out_tok * p.D + head * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_half2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
        u.y = pack_half2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
        u.z = pack_half2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
        u.w = pack_half2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(op + c * 8) = u;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace srb
