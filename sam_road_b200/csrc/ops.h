// sam_road_b200 :: internal C++ launcher declarations (host side).  Every launcher is stream-ordered,
// asynchronous, returns 0 on success and sets srb::set_last_error() otherwise.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace srb {

enum Act : int {
  ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SIGMOID = 3,
  ACT_GELU_SCALAR = 4,         // same function, one element per instruction (A/B timing only)
  // tools/gemm_probe.py only: epilogue ablations that do NOT produce the result
  ACT_PROBE_SKIP = 100,        // release the accumulator untouched (main-loop ceiling)
  ACT_PROBE_TMEM = 101,        // TMEM loads only
  ACT_PROBE_NOSTORE = 102,     // full GELU epilogue without the global stores
  ACT_PROBE_NOTMA = 0x80,      // flag (2-CTA fp16 epilogue): everything but the TMA store (no output)
  ACT_PROBE_DIRECT = 0x40      // flag (2-CTA fp16 epilogue): registers -> st.global instead of smem + TMA store
};

void set_last_error(const char* fmt, ...);
const char* get_last_error();
int device_sm_count();
// true the first time it is called with this mask on the current CUDA device (then sets the bit)
bool first_use_on_device(uint64_t* device_mask);
void note_launch(int n = 1);
// stream memory operations on a 32-bit flag word (no kernel): ordered write / wait until *addr >= value
int stream_write_value32(void* addr, uint32_t value, cudaStream_t st);
int stream_wait_value32_geq(void* addr, uint32_t value, cudaStream_t st);
uint64_t launch_count(bool reset);

// ---- GEMM family (gemm_ops.cu) : C = A[M,K] * W[N,K]^T with fused epilogues ------------------------
int gemm_f16out(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                const float* bias, int act, __half* out, int ldo, cudaStream_t st);
int gemm_f32out(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                const float* bias, const float* resid, const float* pos, int pos_rows, float* out,
                int ldo, cudaStream_t st);
int gemm_ln(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
            const float* bias, const float* resid, const float* gamma, const float* beta, float eps,
            int group, int act, __half* out16, float* out32, float* out_nchw, int tokens, int ldo,
            cudaStream_t st, int conv_s = 0);   // conv_s > 0: A is an NHWC image, implicit 3x3 / pad 1 conv
int gemm_dec_final(const __half* A, int lda, const __half* W, int ldw, int M, int K,
                   const float* bias3, const float* w4, const float* bias4, int s, int P,
                   float* scores, float* logits, cudaStream_t st);
// plain SIMT fp32-accumulate GEMM used only by the on-device unit tests as an independent checker
void gemm_disable_2cta(int mode);   // test hook: bit 0 forces the 1-CTA kernels, bit 1 the register-path fp32 epilogue
int gemm_ref_simt(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K,
                  float* out, int ldo, cudaStream_t st);

// ---- elementwise / data-movement kernels (kernels.cu) ------------------------------------------------
int layernorm_f16(const float* x, const float* gamma, const float* beta, float eps, int M, int D,
                  __half* out, cudaStream_t st);
// rgb: [B,P,P,3] fp32 (dtype 0) or uint8 (dtype 1); out: [B*(P/16)^2, 768] fp16, k = ky*48+kx*3+c
int im2col_patch16(const void* rgb, int dtype, int B, int P, const float* mean, const float* inv_std,
                   __half* out, cudaStream_t st);
// scene: uint8 [H,W,3]; tile_xy: device int32 [B,2] origins (x0,y0) inside the scene; out: uint8 [B,P,P,3]
int crop_tiles(const uint8_t* scene, int H, int W, const int* tile_xy, int B, int P, uint8_t* out,
               cudaStream_t st);
// x: [B*s*s, C] fp16 NHWC; out: [B*s*s, 9*C], k = (ky*3+kx)*C + c, zero padding 1
int im2col_3x3(const __half* x, int B, int s, int C, __half* out, cudaStream_t st);
int convert_f32_f16(const float* x, long n, __half* out, cudaStream_t st);

// ---- encoder attention (attention.cu) --------------------------------------------------------------------
// qkv: [B*s*s, 3*D] fp16, columns (q|k|v) x head x hd ; out: [B*s*s, D] fp16.
// win == s means global attention; otherwise window attention over zero-padded LN output, whose pad
// tokens have q=k=v=bias (image_encoder.py:168-172,227).  rel_h/rel_w: [2*win-1, hd] fp32.
// rel_tab (optional): fp16 [64 or 128, 64] = [rel_pos_h ; rel_pos_w ; 0] as packed by
// pack_rel_table(); when null the launcher packs it on the fly into a per-device scratch.
int encoder_attention(const __half* qkv, const float* qkv_bias, const float* rel_h,
                      const float* rel_w, const __half* rel_tab, int B, int s, int win, int heads,
                      int hd, __half* out, cudaStream_t st);
// rows of the packed table for a window size: each half holds 2*win-1 rows (64 / 128 / 256 rows in all)
inline int rel_table_rows(int win) { return 4 * win - 2 <= 64 ? 64 : (4 * win - 2 <= 128 ? 128 : 256); }
int pack_rel_table(const float* rel_h, const float* rel_w, int win, int hd, __half* tab,
                   cudaStream_t st);
// force the SIMT v1 attention kernel (tests use it as the independent on-device checker)
void attention_force_simt(int mode);   // bit 0: SIMT kernel, bit 1: tcgen05 kernel with 1-in-4 polynomial exps, bit 2: half-block stagger, bit 3: no MUFU turn-taking
void attention_set_trace(long long* device_buffer_128);   // debug: per-phase clock64 stamps of CTA 0

// ---- TopoNet pieces (toponet.cu) -----------------------------------------------------------------------------
// points dtype: 0 = float32, 1 = int64, 2 = int32 ; pairs dtype: 1 = int64, 2 = int32
int topo_sample_features(const float* feat_nchw, int B, int C, int s, int P, const void* points,
                         int pts_dtype, int N, __half* out, cudaStream_t st);
int topo_pair_features(const float* pst, const float* w_off, const float* bias, const void* points,
                       int pts_dtype, const void* pairs, int pairs_dtype, int B, int N, int Ns,
                       int Np, int zero_offset, float* x32, __half* x16, cudaStream_t st);
int topo_fix_valid(const uint8_t* valid, int rows, int Np, uint8_t* out, cudaStream_t st);
int topo_attention(const __half* qkv, const uint8_t* valid, int rows, int Np, __half* out,
                   cudaStream_t st);
int topo_output(const float* x32, const uint8_t* valid_fixed, const float* w, const float* b,
                int tokens, float* logits, float* scores, cudaStream_t st);

// fused 3-layer transformer + output_proj for n_pairs == 16 (toponet_tc.cuh)
struct TopoFusedParams {
  const float *in_b[3], *out_b[3], *l1_b[3], *l2_b[3], *n1_g[3], *n1_b[3], *n2_g[3], *n2_b[3];
  const float* out_w;        // [128]
  const float* out_b_final;  // [1]
};
struct TopoPairInputs {       // what topo_pair_features reads; the fused kernel forms x itself
  const float* pst;          // [B*N, 256] per-point projections (Ws f | Wt f)
  const float* w_off;        // [128][2]
  const float* bias;         // [128]
  const void* points;        // [B, N, 2]
  const void* pairs;         // [B, Ns, Np, 2]
  int pts_dtype, pairs_dtype, N, tokens_per_b, zero_offset;
};
int topo_transformer_fused(const TopoPairInputs& in, const __half* w_chunks, const TopoFusedParams& fp,
                           const uint8_t* valid_fixed, int tokens, float* logits, float* scores,
                           cudaStream_t st);

// ---- SAM mask-decoder path (sam_decoder.cu), USE_SAM_DECODER: True --------------------------------
struct SamAttnW { const float *qw, *qb, *kw, *kb, *vw, *vb, *ow, *ob; };
struct SamDecoderWeights {
  const float* tokens;        // [4][256] = [iou_token ; mask_tokens]
  const float* q0;            // [4][256] = norm1(self_attn(tokens)) of layer 0 (sam_decoder_prepare)
  const float* no_mask_embed; // [256]
  const float* dense_pe;      // [T][256]
  SamAttnW self_attn[2], t2i[2], i2t[2], final_attn;
  const float *n1g[2], *n1b[2], *n2g[2], *n2b[2], *n3g[2], *n3b[2], *n4g[2], *n4b[2];
  const float *l1w[2], *l1b[2], *l2w[2], *l2b[2];
  const float *nfg, *nfb;
  const float* hw[2][3];      // hypernetwork MLPs 1 and 2
  const float* hb[2][3];
  const __half *t2i_kw16[3], *t2i_vw16[3];   // image-side GEMM operands (layers 0, 1, final)
  const __half *i2t_qw16[2], *i2t_ow16[2];
  const __half *up1_w, *up2_w;
  const float *up1_b, *up1_g, *up1_beta, *up2_b;
};
size_t sam_decoder_ws_bytes(int B, int T);
int sam_decoder_prepare(const SamDecoderWeights& w, float* q0_out, cudaStream_t st);
int sam_decoder_forward(const SamDecoderWeights& w, const float* emb_nchw, int B, int s, int P, void* ws,
                        float* mask_scores, float* mask_logits, cudaStream_t st);

// ---- mask fusion (kernels.cu) : inferencer.py:79-110 --------------------------------------------------------
int fuse_masks(const float* scores, int n_tiles, int P, const int* tile_x0, const int* tile_y0,
               int H, int W, uint8_t* keypoint_u8, uint8_t* road_u8, cudaStream_t st);

}  // namespace srb
