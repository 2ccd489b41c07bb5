// sam_road_b200 :: fused TopoNet transformer (model.py:74-86,135-146): all three post-norm encoder
// layers + output_proj for a tile of 128 pair tokens (8 samples x 16 pairs) in ONE persistent kernel.
//
// Per layer (torch TransformerEncoderLayer, d=128, 4 heads, ff=128, relu, LN eps 1e-5, eval mode):
//   qkv = x W_in^T + b_in -> per-sample 16x16 masked attention per head -> x = LN1(x + att W_o^T + b_o)
//   -> x = LN2(x + relu(x W_1^T + b_1) W_2^T + b_2)
// The four GEMMs run on tcgen05 (UMMA 128x128x16, fp16 operands, fp32 in TMEM); everything between
// them stays on chip:
//   TMEM  cols [0,384)   GEMM accumulators (q|k|v, later out_proj / lin1 / lin2 in [0,128))
//         cols [384,512) the fp32 residual stream x of the tile (one token per lane)
//   smem  A buffer 32 KB  current fp16 GEMM A operand (x -> attention output -> x' -> hidden -> x'')
//         KV buffer 66 KB k and v of the tile as fp16 rows
//         weight ring 3 x 32 KB  the 18 [128x128] weight chunks streamed by TMA in consumption order
// Warps: 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..17 = 512 token threads, FOUR per token
// (warps w, w+4, w+8, w+12 share a TMEM lane quarter): epilogues, attention, LayerNorm, output_proj.
// Part p owns columns 32p..32p+31 = head p, and two of the eight 32-column chunks of k|v; a thread
// keeps its 32 columns of the row in registers through a whole LayerNorm (one TMEM read, one write).
// LayerNorm statistics and the output dot product are combined through a small smem exchange, always
// summed in part order.  The 16 x 16 x 32 attention of a (sample, head) is far too small for a UMMA
// tile and runs on warp-level mma.sync: a warp owns two samples of its head; q (bias, 1/sqrt(32),
// fp16) is staged in the warp's own rows of the A buffer, k / v fragments come from the k|v rows by
// ldmatrix / ldmatrix.trans, softmax in fp32 on the S fragments, P (fp16) feeds the PV mma from
// registers.  The token work, not the GEMMs, bounds this kernel, and it is bound by latency (one tile
// in flight per CTA, phases separated by the MMA round trips): four warps per scheduler hide twice
// the latency of two (one -> two threads per token was 1.58x in round 1, two -> four 1.34x, attention
// on mma.sync another 1.36x).  Key-padding semantics: masked keys are excluded from
// the softmax; masked slots report output_proj.bias.
#pragma once

#include "common.cuh"
#include "ops.h"

namespace srb {

constexpr int kTtcThreads = 576;          // 2 + 16 warps
constexpr int kTtcTokenThreads = 512;
constexpr int kTtcWStages = 3;
constexpr int kTtcOffA = 0;                       // 2 k-blocks x 16 KB
constexpr int kTtcOffKV = 32768;                  // 128 rows x 528 B (k|v fp16, padded: ldmatrix rows hit 32 banks)
constexpr int kTtcKVStride = 528;
constexpr int kTtcOffW = kTtcOffKV + 68608;       // 3 x 32 KB
constexpr int kTtcOffBar = kTtcOffW + kTtcWStages * 32768;
constexpr int kTtcOffXch = kTtcOffBar + 256;      // 3 slots x 4 parts x 128 floats: partial sums of the parts
constexpr int kTtcSmemBytes = kTtcOffXch + 3 * 2048 + 1024;
constexpr int kTtcChunksPerLayer = 6;             // Wq, Wk, Wv, Wo, W1, W2

struct TtcLayerParams {
  const float* in_b;    // [384]
  const float* out_b;   // [128]
  const float* l1_b;    // [128]
  const float* l2_b;    // [128]
  const float *n1_g, *n1_b, *n2_g, *n2_b;   // [128]
};

struct TtcParams {
  TtcLayerParams layer[3];
  // pair features (model.py:96-120), computed in the kernel: x = relu(PS[b,src] + PT[b,tgt] +
  // Wo (pt[tgt] - pt[src]) + bias) with the per-point projections pst [B*N, 256] (Ws f | Wt f)
  const float* pst;
  const float* w_off;     // [128][2]
  const float* pair_b;    // [128]
  const void* points;     // [B, N, 2] (x, y)
  const void* pairs;      // [B, Ns, Np, 2] indices into N
  int pts_dtype, pairs_dtype, N, tokens_per_b, zero_offset;
  const uint8_t* valid;   // [tokens] fixed validity (all-invalid rows already flipped), or null
  const float* out_w;     // [128]
  const float* out_b;     // [1]
  float* logits;          // [tokens] or null
  float* scores;          // [tokens] or null
  int tokens;
  int num_tiles;
};

__device__ __forceinline__ float ttc_load_coord(const void* p, int dtype, size_t idx) {
  if (dtype == 0) return static_cast<const float*>(p)[idx];
  if (dtype == 1) return static_cast<float>(static_cast<const long long*>(p)[idx]);
  return static_cast<float>(static_cast<const int*>(p)[idx]);
}
__device__ __forceinline__ long long ttc_load_index(const void* p, int dtype, size_t idx) {
  if (dtype == 1) return static_cast<const long long*>(p)[idx];
  return static_cast<long long>(static_cast<const int*>(p)[idx]);
}

// warp-level MMA for the 16 x 16 attention of one sample and head (far too small for a UMMA tile)
__device__ __forceinline__ void ttc_ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ttc_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
// d (16x8, fp32) += a (16x16, fp16, row) * b (16x8, fp16, col)
__device__ __forceinline__ void ttc_mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Everything the token threads exchange through shared memory stays inside a TMEM lane quarter: the four
// threads of a row (one per part) sit in the four warps of the same quarter, and a warp's two samples
// are rows of its own quarter.  So the barriers are per quarter (ids 1..4, 128 threads each), and the
// quarters only meet at the a_ready mbarrier of the next GEMM.
__device__ __forceinline__ void named_bar_sync_quarter(int quarter) {
  asm volatile("bar.sync %0, 128;" ::"r"(quarter + 1) : "memory");
}

__global__ void __launch_bounds__(kTtcThreads, 1)   // 18 warps: 5 on two of the sub-partitions -> 96 registers
toponet_tc_kernel(const __grid_constant__ CUtensorMap tmW, TtcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem + kTtcOffA;
  uint8_t* sKV = smem + kTtcOffKV;
  uint8_t* sW = smem + kTtcOffW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTtcOffBar);
  uint64_t* a_ready = bars + 2;     // 512 token threads: new A operand written
  uint64_t* acc_ready = bars + 3;   // MMA commit: GEMM result in TMEM
  uint64_t* w_full = bars + 4;      // [3]
  uint64_t* w_empty = bars + 7;     // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    mbar_init(a_ready, kTtcTokenThreads);
    mbar_init(acc_ready, 1);
    for (int i = 0; i < kTtcWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int wc = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int ch = 0; ch < 3 * kTtcChunksPerLayer; ++ch, ++wc) {
          const int st = wc % kTtcWStages;
          mbar_wait(&w_empty[st], ((wc / kTtcWStages) & 1) ^ 1u);
          mbar_arrive_expect_tx(&w_full[st], 32768);
          tma_load_2d(sW + st * 32768, &tmW, &w_full[st], 0, ch * 128);
          tma_load_2d(sW + st * 32768 + 16384, &tmW, &w_full[st], 64, ch * 128);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 128);
      int wc = 0, ac = 0;   // weight chunks consumed, a_ready completions consumed
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int l = 0; l < 3; ++l) {
          for (int gemm = 0; gemm < 4; ++gemm) {          // 0: in_proj (3 chunks), 1: out, 2: lin1, 3: lin2
            // every GEMM reads an A operand the token threads have just written (the first one of a
            // tile: the pair features) and overwrites accumulator columns they read in the previous
            // step: each waits for their a_ready arrival
            mbar_wait(a_ready, ac & 1);
            ++ac;
            tc_fence_after_sync();
            const int nch = gemm == 0 ? 3 : 1;
            for (int j = 0; j < nch; ++j, ++wc) {
              const int st = wc % kTtcWStages;
              mbar_wait(&w_full[st], (wc / kTtcWStages) & 1);
              tc_fence_after_sync();
              const uint32_t abase = smem_u32(sA), wbase = smem_u32(sW + st * 32768);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t adesc = umma_desc_k128(abase + (k >> 2) * 16384) + 2 * (k & 3);
                const uint64_t bdesc = umma_desc_k128(wbase + (k >> 2) * 16384) + 2 * (k & 3);
                umma_f16_ss(tmem_base + j * 128, adesc, bdesc, idesc, k != 0 ? 1u : 0u);
              }
              umma_commit(&w_empty[st]);
            }
            umma_commit(acc_ready);
          }
        }
      }
    }
  } else {
    // =========================== token threads ===========================
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                 // 0..3: columns 32*part .. 32*part+31, head `part`
    const int row = quarter * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tAcc = tmem_base + tlane;          // cols [0,384)
    const uint32_t tRes = tmem_base + tlane + 384;    // cols [384,512)
    const int sw = row & 7;
    uint8_t* myA = sA + row * 128;
    float* xch = reinterpret_cast<float*>(smem + kTtcOffXch);   // [3][4 parts][128 rows]
    int ti = 0, rc = 0;                               // tiles, acc_ready completions consumed

    auto write_a_chunk = [&](int c, const float (&v)[32]) {   // 32 fp32 -> fp16 into the swizzled A buffer
      uint8_t* dst = myA + (c >> 1) * 16384;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 u;
        u.x = pack_half2(v[q4 * 8 + 0], v[q4 * 8 + 1]);
        u.y = pack_half2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
        u.z = pack_half2(v[q4 * 8 + 4], v[q4 * 8 + 5]);
        u.w = pack_half2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
        *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + q4) ^ sw) << 4)) = u;
      }
    };
    auto ld_chunk = [&](uint32_t taddr, float (&v)[32]) {
      uint32_t r[32];
      tmem_ld_32x32(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    };
    auto st_chunk = [&](uint32_t taddr, const float (&v)[32]) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
      tmem_st_32x32(taddr, r);
    };
    // the four parts' partials of the same row, summed in part order (slot: 0 sum, 1 var, 2 dot): every
    // thread of a row gets the same value
    auto combine = [&](int slot, float mine) -> float {
      float* x = xch + slot * 512 + row;
      x[part * 128] = mine;
      named_bar_sync_quarter(quarter);
      return ((x[0] + x[128]) + x[256]) + x[384];
    };
    auto ldg32 = [&](const float* src, float (&v)[32]) {     // 32 consecutive floats (128 B aligned)
      const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(s4 + i);
        v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
    };
    // x = LayerNorm(res + acc + bias) ; res <- x ; optionally A buffer <- fp16(x); returns x.w_out
    // (this part's 32 columns, held in registers; exact two-pass statistics over the whole row)
    auto residual_layernorm = [&](const float* bias, const float* gamma, const float* beta,
                                  bool write_a, const float* wdot) -> float {
      const int c = part;
      float r[32];
      float sum = 0.f;
      {
        uint32_t ua[32], ur[32];
        tmem_ld_32x32_nowait(tAcc + c * 32, ua);
        tmem_ld_32x32_nowait(tRes + c * 32, ur);
        float bv[32];
        ldg32(bias + c * 32, bv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          r[i] = __uint_as_float(ur[i]) + (__uint_as_float(ua[i]) + bv[i]);
          sum += r[i];
        }
      }
      const float mean = combine(0, sum) * (1.0f / 128.0f);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = r[i] - mean;
        var = fmaf(d, d, var);
      }
      const float rstd = rsqrtf(combine(1, var) * (1.0f / 128.0f) + 1e-5f);
      float dot = 0.f;
      {
        float gv[32], bt[32];
        ldg32(gamma + c * 32, gv);
        ldg32(beta + c * 32, bt);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = (r[i] - mean) * rstd * gv[i] + bt[i];
        if (wdot) {
          ldg32(wdot + c * 32, gv);
#pragma unroll
          for (int i = 0; i < 32; ++i) dot = fmaf(r[i], gv[i], dot);
        }
      }
      st_chunk(tRes + c * 32, r);
      if (write_a) write_a_chunk(c, r);
      tmem_st_wait();
      return dot;
    };

    // source / target rows and offset of this thread's pair token in tile `t` (index -> address chain
    // of the gather; issued one tile ahead so that only the pst loads themselves are exposed)
    float nx_ox = 0.f, nx_oy = 0.f;
    size_t nx_ps = 0, nx_pt = 0;
    auto pair_lookup = [&](int t) {
      const long tk = static_cast<long>(t) * 128 + row;
      nx_ox = nx_oy = 0.f;
      nx_ps = nx_pt = 0;
      if (t < p.num_tiles && tk < p.tokens) {
        const size_t b = static_cast<size_t>(tk) / p.tokens_per_b;
        const long long nm1 = static_cast<long long>(p.N) - 1;    // clamp: no out-of-bounds gather on bad indices
        nx_ps = b * p.N + min(max(ttc_load_index(p.pairs, p.pairs_dtype, static_cast<size_t>(tk) * 2 + 0), 0LL), nm1);
        nx_pt = b * p.N + min(max(ttc_load_index(p.pairs, p.pairs_dtype, static_cast<size_t>(tk) * 2 + 1), 0LL), nm1);
        if (!p.zero_offset) {
          nx_ox = ttc_load_coord(p.points, p.pts_dtype, nx_pt * 2 + 0) -
                  ttc_load_coord(p.points, p.pts_dtype, nx_ps * 2 + 0);
          nx_oy = ttc_load_coord(p.points, p.pts_dtype, nx_pt * 2 + 1) -
                  ttc_load_coord(p.points, p.pts_dtype, nx_ps * 2 + 1);
        }
      }
    };
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
      const long tok = static_cast<long>(tile) * 128 + row;
      const bool tok_ok = tok < p.tokens;
      // key-validity bits of this token's sample (16 consecutive tokens)
      uint32_t kmask = 0xffffu;
      bool my_valid = true;
      if (p.valid) {
        kmask = 0;
        const long s0 = static_cast<long>(tile) * 128 + (row & ~15);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (s0 + j < p.tokens && p.valid[s0 + j]) kmask |= 1u << j;
        if (kmask == 0) kmask = 0xffffu;            // tail rows beyond `tokens`
        my_valid = tok_ok && p.valid[tok];
      }

      // ---- pair features of this token -> TMEM residual (fp32) and A buffer (fp16) ----
      {
        if (ti == 0) pair_lookup(tile);              // later tiles: looked up during the previous tile
        const float ox = nx_ox, oy = nx_oy;
        const size_t ps = nx_ps, pt = nx_pt;
        {
          const int c = part;
          float v[32];
          const float4* a4 = reinterpret_cast<const float4*>(p.pst + ps * 256 + c * 32);
          const float4* b4 = reinterpret_cast<const float4*>(p.pst + pt * 256 + 128 + c * 32);
          const float4* w4 = reinterpret_cast<const float4*>(p.w_off + c * 64);
          const float4* c4 = reinterpret_cast<const float4*>(p.pair_b + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = a4[i], bb = b4[i], w0 = __ldg(w4 + 2 * i), w1 = __ldg(w4 + 2 * i + 1),
                         cb = __ldg(c4 + i);
            float x0 = a.x + bb.x, x1 = a.y + bb.y, x2 = a.z + bb.z, x3 = a.w + bb.w;
            x0 += w0.x * ox + w0.y * oy + cb.x;
            x1 += w0.z * ox + w0.w * oy + cb.y;
            x2 += w1.x * ox + w1.y * oy + cb.z;
            x3 += w1.z * ox + w1.w * oy + cb.w;
            v[4 * i + 0] = tok_ok ? fmaxf(x0, 0.f) : 0.f;
            v[4 * i + 1] = tok_ok ? fmaxf(x1, 0.f) : 0.f;
            v[4 * i + 2] = tok_ok ? fmaxf(x2, 0.f) : 0.f;
            v[4 * i + 3] = tok_ok ? fmaxf(x3, 0.f) : 0.f;
          }
          st_chunk(tRes + c * 32, v);
          write_a_chunk(c, v);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(a_ready);
      }

      float dot = 0.f;
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        const TtcLayerParams& L = p.layer[l];
        // ================= qkv: k, v -> smem (fp16 rows), attention per head =================
        mbar_wait(acc_ready, rc & 1); ++rc;
        tc_fence_after_sync();
        {                                            // parts 0,1: k (cols 128..255), parts 2,3: v (cols 256..383)
          uint32_t kv[2][32];
          tmem_ld_32x32_nowait(tAcc + 128 + (part * 2 + 0) * 32, kv[0]);
          tmem_ld_32x32_nowait(tAcc + 128 + (part * 2 + 1) * 32, kv[1]);
          tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = part * 2 + cc;
            uint8_t* dst = sKV + row * kTtcKVStride + ((c * 64) ^ (((row >> 4) & 1) << 6));
            const float4* b4 = reinterpret_cast<const float4*>(L.in_b + 128 + c * 32);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 b0 = __ldg(b4 + 2 * q4), b1 = __ldg(b4 + 2 * q4 + 1);
              const uint32_t* v = kv[cc] + q4 * 8;
              uint4 u;
              u.x = pack_half2(__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y);
              u.y = pack_half2(__uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w);
              u.z = pack_half2(__uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y);
              u.w = pack_half2(__uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w);
              *reinterpret_cast<uint4*>(dst + q4 * 16) = u;
            }
          }
        }
        {
          // ---- attention of head `part` for the warp's two samples on mma.sync (16 queries x 16 keys x
          // 32 dims per sample): q (bias added, scaled, fp16) is staged in the warp's own rows of the A
          // buffer -- the 64 B chunk (row, h) the attention output overwrites afterwards --, k and v come
          // straight from their smem rows through ldmatrix, P stays in registers (S fragments -> A
          // fragments), O is normalised and stored as the out_proj A operand.
          const int h = part;
          {
            uint32_t q[32];
            tmem_ld_32x32_nowait(tAcc + h * 32, q);  // in flight across the barrier
            named_bar_sync_quarter(quarter);         // the k and v rows of this quarter's samples are in smem
            tmem_ld_wait();
            float qs[32];
            const float4* b4 = reinterpret_cast<const float4*>(L.in_b + h * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {            // torch MHA scales q by 1/sqrt(head_dim)
              const float4 bb = __ldg(b4 + i);
              qs[4 * i + 0] = (__uint_as_float(q[4 * i + 0]) + bb.x) * 0.17677669529663687f;
              qs[4 * i + 1] = (__uint_as_float(q[4 * i + 1]) + bb.y) * 0.17677669529663687f;
              qs[4 * i + 2] = (__uint_as_float(q[4 * i + 2]) + bb.z) * 0.17677669529663687f;
              qs[4 * i + 3] = (__uint_as_float(q[4 * i + 3]) + bb.w) * 0.17677669529663687f;
            }
            write_a_chunk(h, qs);
          }
          __syncwarp();
          const int g8 = lane >> 2, t4 = lane & 3;   // fragment coordinates: row g8 (+8), column pair t4
          const int mi = lane >> 3, rr = lane & 7;   // ldmatrix: this lane addresses row rr of matrix mi
          const uint32_t sA_h = smem_u32(sA) + (h >> 1) * 16384;
#pragma unroll
          for (int sh = 0; sh < 2; ++sh) {
            const int r0 = quarter * 32 + sh * 16;                 // first tile row of the sample
            const int sxs = ((r0 >> 4) & 1) << 6;                  // odd samples: k|v columns XOR 64 B
            const uint32_t km = __shfl_sync(0xffffffffu, kmask, sh * 16);
            // S = Q K^T : two 8-key column tiles, two 16-dim k-steps
            float sacc[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int i = 0; i < 4; ++i) sacc[nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              uint32_t qa[4], kb[4];
              {   // matrices: (rows 0-7 | 8-15) x (dims ks*16 + 0-7 | 8-15)
                const int qr = r0 + (mi & 1) * 8 + rr;
                const int piece = (h & 1) * 4 + ks * 2 + (mi >> 1);
                ttc_ldmatrix_x4(qa, sA_h + qr * 128 + ((piece ^ (qr & 7)) << 4));
              }
              {   // matrices: (keys 0-7 | 8-15) x (dims ks*16 + 0-7 | 8-15) -> b0, b1 of tile 0; of tile 1
                const int kr = r0 + (mi >> 1) * 8 + rr;
                ttc_ldmatrix_x4(kb, smem_u32(sKV) + kr * kTtcKVStride + ((h * 64) ^ sxs) +
                                        (ks * 16 + (mi & 1) * 8) * 2);
              }
              ttc_mma_16816(sacc[0], qa, kb[0], kb[1]);
              ttc_mma_16816(sacc[1], qa, kb[2], kb[3]);
            }
            // masked softmax of rows g8 (values [nt][0..1]) and g8 + 8 ([nt][2..3]); keys nt*8 + 2*t4 (+1)
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const bool on = (km >> (nt * 8 + 2 * t4 + e)) & 1u;
                sacc[nt][e] = on ? sacc[nt][e] : -INFINITY;
                sacc[nt][2 + e] = on ? sacc[nt][2 + e] : -INFINITY;
                mx0 = fmaxf(mx0, sacc[nt][e]);
                mx1 = fmaxf(mx1, sacc[nt][2 + e]);
              }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                sacc[nt][e] = __expf(sacc[nt][e] - mx0);           // exp(-inf) = 0 for masked keys
                sacc[nt][2 + e] = __expf(sacc[nt][2 + e] - mx1);
                l0 += sacc[nt][e];
                l1 += sacc[nt][2 + e];
              }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            uint32_t pa[4];
            pa[0] = pack_half2(sacc[0][0], sacc[0][1]);
            pa[1] = pack_half2(sacc[0][2], sacc[0][3]);
            pa[2] = pack_half2(sacc[1][0], sacc[1][1]);
            pa[3] = pack_half2(sacc[1][2], sacc[1][3]);
            // O = P V : four 8-dim column tiles, one 16-key k-step
            float oacc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int i = 0; i < 4; ++i) oacc[j][i] = 0.f;
#pragma unroll
            for (int dh = 0; dh < 2; ++dh) {
              uint32_t vb[4];   // transposed matrices: (keys 0-7 | 8-15) x (dims dh*16 + 0-7 | 8-15)
              const int vr = r0 + (mi & 1) * 8 + rr;
              ttc_ldmatrix_x4_trans(vb, smem_u32(sKV) + vr * kTtcKVStride + ((256 + h * 64) ^ sxs) +
                                            (dh * 16 + (mi >> 1) * 8) * 2);
              ttc_mma_16816(oacc[dh * 2 + 0], pa, vb[0], vb[1]);
              ttc_mma_16816(oacc[dh * 2 + 1], pa, vb[2], vb[3]);
            }
            const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
            __syncwarp();                            // every lane's Q fragments of this sample are loaded
            const int ra = r0 + g8, rb = r0 + g8 + 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {            // attention output columns h*32 + j*8 + 2*t4 (+1)
              const int piece = (h & 1) * 4 + j;
              *reinterpret_cast<uint32_t*>(sA + (h >> 1) * 16384 + ra * 128 + ((piece ^ (ra & 7)) << 4) + t4 * 4) =
                  pack_half2(oacc[j][0] * inv0, oacc[j][1] * inv0);
              *reinterpret_cast<uint32_t*>(sA + (h >> 1) * 16384 + rb * 128 + ((piece ^ (rb & 7)) << 4) + t4 * 4) =
                  pack_half2(oacc[j][2] * inv1, oacc[j][3] * inv1);
            }
          }
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(a_ready);
        // ================= out_proj + residual + LayerNorm1 =================
        mbar_wait(acc_ready, rc & 1); ++rc;
        tc_fence_after_sync();
        residual_layernorm(L.out_b, L.n1_g, L.n1_b, true, nullptr);
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(a_ready);
        // ================= linear1 + relu =================
        mbar_wait(acc_ready, rc & 1); ++rc;
        tc_fence_after_sync();
        {
          const int c = part;
          float v[32], bv[32];
          ldg32(L.l1_b + c * 32, bv);
          ld_chunk(tAcc + c * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bv[i], 0.f);
          write_a_chunk(c, v);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(a_ready);
        if (l == 2) pair_lookup(tile + static_cast<int>(gridDim.x));   // next tile's indices, off the critical path
        // ================= linear2 + residual + LayerNorm2 =================
        mbar_wait(acc_ready, rc & 1); ++rc;
        tc_fence_after_sync();
        dot = residual_layernorm(L.l2_b, L.n2_g, L.n2_b, l < 2, l == 2 ? p.out_w : nullptr);
        tc_fence_before_sync();
        fence_proxy_async_smem();
        if (l < 2) mbar_arrive(a_ready);   // after the last layer the next tile's pair features arrive
      }
      // ================= output_proj + sigmoid =================
      dot = combine(2, dot);
      if (tok_ok && part == 0) {
        const float b = __ldg(p.out_b);
        const float lg = my_valid ? dot + b : b;
        if (p.logits) p.logits[tok] = lg;
        if (p.scores) p.scores[tok] = 1.0f / (1.0f + expf(-lg));
      }
      tc_fence_before_sync();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace srb
