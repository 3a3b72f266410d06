/*
 * medsam2_b200 — C-ABI of the B200-native (sm_100a) kernels behind Medical-SAM2's per-slice
 * inference hot path.  This library replaces the reference's only native boundary
 * (sam2_train/_C.so built from sam2_train/csrc/connected_components.cu:213-289) and, in
 * addition, every tensor op the reference delegates to torch (cuBLAS/cuDNN/SDPA) on that path.
 *
 * Conventions (all entry points):
 *   - plain pointers are DEVICE pointers unless named h_*; sizes are element counts;
 *   - `dt` arguments: MS2_F32 (0) or MS2_BF16 (1);
 *   - stream-ordered on `stream` (a cudaStream_t), no implicit synchronisation, no allocation
 *     (workspaces are passed in), no global state besides cached TMA descriptors;
 *   - return 0 on success, <0 on error; ms2_last_error() gives the message (thread-local);
 *   - inputs are borrowed and never modified unless documented "in place".
 * Layouts are token-major (NHWC / [tokens, C]) unless stated.
 */
#ifndef MEDSAM2_B200_H
#define MEDSAM2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS2_F32 0
#define MS2_BF16 1

#define MS2_ACT_NONE 0
#define MS2_ACT_GELU 1 /* exact erf GELU (nn.GELU default) */
#define MS2_ACT_RELU 2
#define MS2_ACT_SIGMOID 3

typedef void* ms2_stream_t; /* cudaStream_t */

int ms2_version(void);
const char* ms2_last_error(void);
/* 1 if the running device is compute capability 10.x (tcgen05 path usable) */
int ms2_device_is_sm100(void);

/* ---- connected components: replaces _C.get_connected_componnets
 *      (reference sam2_train/csrc/connected_components.cu:213-282; caller utils/misc.py:47-63).
 *      mask uint8 [N,1,H,W] (H,W even) -> labels,counts int32 [N,1,H,W].
 *      label(p) = 1 + min over p's 8-connected component of ((row&~1)*W + (col&~1)); 0 for bg.
 *      workspace: int32 [N*H*W], only read/written when H*W/4 exceeds the shared-memory path
 *      (may be NULL otherwise). */
int ms2_cc_label(const uint8_t* mask, int32_t* labels, int32_t* counts, int32_t* workspace,
                 int N, int H, int W, ms2_stream_t stream);
/* fused hole filling (reference utils/misc.py:247-258): out = (bg component of (in<=thresh) with
 * area<=max_area) ? fill_value : in.  in/out fp32 [N,1,H,W]; requires the shared-memory path
 * (H*W/4 <= 28672 blocks, e.g. 256x256). */
int ms2_fill_holes(const float* in, float* out, int N, int H, int W, float thresh, int max_area,
                   float fill_value, ms2_stream_t stream);

/* ---- LayerNorm over the last dim (nn.LayerNorm / LayerNorm2d in NHWC):
 *      y = act(LN(x [+ add]) * gamma + beta); x,add fp32 [M,C]; y dtype y_dt. */
int ms2_layernorm(const float* x, const float* add, const float* gamma, const float* beta, void* y,
                  int y_dt, int M, int C, float eps, int act, ms2_stream_t stream);

/* ---- GEMM with fused epilogue (every nn.Linear / 1x1 conv / ConvTranspose k2s2 / im2col conv):
 *      out[M,N] = residual + colscale * act(A[M,K] @ W[N,K]^T + bias)
 *      A dtype a_dt (row stride lda), W dtype w_dt (row stride K), bias/colscale fp32 [N] or NULL,
 *      residual fp32 (row stride ldr) or NULL, out dtype o_dt (row stride ldo).
 *      impl: 0 = auto, 1 = SIMT fp32-accumulate reference kernel, 2 = tcgen05/TMA (bf16 only),
 *      3 = small-M (M <= 64) latency kernel for the decoder's token-side MLPs. */
int ms2_gemm(const void* A, int a_dt, long lda, const void* W, int w_dt, const float* bias,
             const float* colscale, const float* residual, long ldr, void* out, int o_dt, long ldo,
             int M, int N, int K, int act, int impl, ms2_stream_t stream);

/* up to 8 independent small-M GEMMs (same M <= 64, same K, per-group N / activation) in one launch:
 * out_g[M,N_g] = act_g(A_g[M,K] @ W_g[N_g,K]^T + bias_g).  h_* are HOST arrays of `groups` entries holding device
 * pointers / strides; used for the six 3-layer head MLPs of the mask decoder (mask_decoder.py:75-92,238-267). */
int ms2_gemm_smallm_grouped(int groups, const void* const* h_A, const long* h_lda, const void* const* h_W,
                            const float* const* h_bias, void* const* h_out, const long* h_ldo,
                            const int* h_N, const int* h_act, int a_dt, int o_dt, int M, int K,
                            ms2_stream_t stream);

/* ---- scaled-dot-product attention, no mask (F.scaled_dot_product_attention call sites:
 *      hieradet.py:72-76, transformer.py:252-258, :318).  Dense mode:
 *      q [B,Hh,Lq,D] / k,v [B,Hh,Lk,D] / o [B,Hh,Lq,D] addressed through element strides
 *      (batch, head, token); D in {16,32,64,96,128,256}; dtype dt for q,k,v,o. */
int ms2_attention(const void* q, const void* k, const void* v, void* o, int dt,
                  long q_bs, long q_hs, long q_ts, long k_bs, long k_hs, long k_ts,
                  long v_bs, long v_hs, long v_ts, long o_bs, long o_hs, long o_ts,
                  int B, int Hh, int Lq, int Lk, int D, float scale, int impl, ms2_stream_t stream);
/* same, with a caller-provided scratch buffer that lets the tcgen05 path split the keys over several
 * CTAs per query tile (flash-decoding style) when B*Hh*ceil(Lq/128) alone cannot fill the SMs:
 * workspace_bytes >= nsplit * B*Hh*Lq*(D+2)*4 enables up to nsplit splits (the kernel picks the count);
 * impl: 0 = auto, 1 = SIMT fp32-accumulate kernel, 2 = tcgen05/TMA (bf16, D in {64,96,128,256}),
 * 3 = small-shape kernels of the mask decoder (D in {16,32}; Lk <= 32 keys or Lq <= 16 queries). */
int ms2_attention_ws(const void* q, const void* k, const void* v, void* o, int dt,
                     long q_bs, long q_hs, long q_ts, long k_bs, long k_hs, long k_ts,
                     long v_bs, long v_hs, long v_ts, long o_bs, long o_hs, long o_ts,
                     int B, int Hh, int Lq, int Lk, int D, float scale, int impl,
                     void* workspace, long workspace_bytes, ms2_stream_t stream);

/* memory cross-attention with un-projected values (tcgen05 only): q,k head dim D = 256, v/o head dim DV = 64.
 * Because softmax rows sum to one, softmax(QK^T)(M Wv^T + b) = (softmax(QK^T) M) Wv^T + b, so the kernel attends
 * over the raw 64-d memory values M and the caller applies v_proj to the 64-d result (memory_attention.py:73-79,
 * transformer.py:292-294,318): 4x less P.V work, V traffic and tensor memory. */
int ms2_attention_dv(const void* q, const void* k, const void* v, void* o, int dt,
                     long q_bs, long q_hs, long q_ts, long k_bs, long k_hs, long k_ts,
                     long v_bs, long v_hs, long v_ts, long o_bs, long o_hs, long o_ts,
                     int B, int Hh, int Lq, int Lk, int D, int DV, float scale,
                     void* workspace, long workspace_bytes, ms2_stream_t stream);

/* Split-KV memory cross-attention over several GPUs (the reference has no counterpart: it attends
 * over the whole bank on one GPU, memory_attention.py:73-79 -> transformer.py:288-331).  Every rank attends over its own
 * share of the bank and emits ONE un-normalised partial per query row:
 *   part_o [B*Lq, DV] fp32 = sum_j 2^(s_j - m) v_j,   part_ml [B*Lq, 2] fp32 = (m = row max of s in log2 units, l = sum_j 2^(s_j - m)).
 * ms2_attention_merge combines `nparts` partials (part r at parts_o + r*part_stride and parts_ml + r*part_stride, in
 * floats; an empty share is (0, -inf, 0)) into the softmax-normalised o [B,Lq,DV]. */
int ms2_attention_dv_partial(const void* q, const void* k, const void* v, float* part_o, float* part_ml, int dt,
                             long q_bs, long q_ts, long k_bs, long k_ts, long v_bs, long v_ts,
                             int B, int Lq, int Lk, int D, int DV, float scale,
                             void* workspace, long workspace_bytes, ms2_stream_t stream);
int ms2_attention_merge(const float* parts_o, const float* parts_ml, long part_stride, void* o, int dt,
                        long o_bs, long o_ts, int B, int Lq, int DV, int nparts, ms2_stream_t stream);

/* ---- the same exchange WITHOUT a collective call (fused compute + transfer over NVLink peer memory; replaces the
 *      dist.all_gather of the packed partial).  Every rank owns a gather buffer of `world` packed partials
 *      ([rows*64] un-normalised O then [rows*2] (m,l) per rank) and `world` uint32 flag words, both in peer-accessible
 *      (symmetric) device memory.  ms2_attention_dv_partial_push computes this rank's partial over its Lk keys (Lk == 0:
 *      the empty partial) and its last kernel stores it into slot `rank` of EVERY rank's buffer — h_dst[r] / h_flag[r] are
 *      HOST arrays of the DEVICE addresses of this rank's slot / flag word in rank r's memory — then publishes `step` in
 *      the flags (st.release.sys after a grid-wide fence; `counter` = one zeroed uint32 of local device memory).
 *      ms2_attention_merge_wait spins until all `world` local flags show `step` (ld.acquire.sys) and merges the local
 *      gather buffer `parts` (stride `part_stride` floats between ranks) into bf16 o [B,Lq,64].  Two buffers used
 *      alternately make back-to-back exchanges safe.  workspace: >= B*Lq*66*4 bytes. */
int ms2_attention_dv_partial_push(const void* q, const void* k, const void* v, int dt, long q_bs, long q_ts, long k_bs,
                                  long k_ts, long v_bs, long v_ts, int B, int Lq, int Lk, int D, int DV, float scale,
                                  void* workspace, long workspace_bytes, const void* const* h_dst, const void* const* h_flag,
                                  int world, int step, void* counter, ms2_stream_t stream);
int ms2_attention_merge_wait(const float* parts, long part_stride, const void* flags, int step, int world, void* o, int dt,
                             long o_bs, long o_ts, int B, int Lq, int DV, ms2_stream_t stream);

/* ---- Hiera windowed attention with window partition / zero-pad-as-bias-key / q max-pool /
 *      unpartition+crop folded into the loads and stores (hieradet.py:58-83,136-159,
 *      backbones/utils.py:16-62).  qkv [B,H,W,3,heads,D] dtype dt (the qkv Linear output on the
 *      UNPADDED tokens), qkv_bias fp32 [3*heads*D] (value of q/k/v at zero-padded positions),
 *      out [B,Ho,Wo,heads*D] dtype dt with Ho=H/(qpool?2:1).  ws = window size (on the input grid). */
int ms2_window_attention(const void* qkv, const float* qkv_bias, void* out, int dt, int B, int H, int W,
                         int heads, int D, int ws, int qpool, float scale, ms2_stream_t stream);
/* same with an explicit implementation choice: 0 = auto, 1 = SIMT fp32-accumulate kernel,
 * 2 = tcgen05 (bf16, D = 96, ws*ws <= 256: several windows packed per M=128 tile, block-diagonal mask). */
int ms2_window_attention_impl(const void* qkv, const float* qkv_bias, void* out, int dt, int B, int H, int W,
                              int heads, int D, int ws, int qpool, float scale, int impl, ms2_stream_t stream);

/* ---- 2x2 max pool, stride 2, NHWC fp32 (hieradet.py:23-34 on the shortcut). */
int ms2_maxpool2x2(const float* x, float* y, int B, int H, int W, int C, ms2_stream_t stream);

/* ---- PatchEmbed: conv 7x7 stride 4 pad 3, 3->Cout, NCHW fp32 image -> NHWC fp32 tokens,
 *      + precomputed pos-embed table [Ho,Wo,Cout] (backbones/utils.py:87-95, hieradet.py:269-284).
 *      w fp32 [Cout,3,7,7]. */
int ms2_patch_embed(const float* img, const float* w, const float* bias, const float* pos, float* out,
                    int B, int Hin, int Win, int Cout, ms2_stream_t stream);

/* bf16 path of the same conv: NCHW image (img_dt fp32 or bf16) -> bf16 im2col rows [B*Ho*Wo, ldk] (taps in (ky,kx,c) order, zero
 * beyond 147; ldk a multiple of 8 >= 152); the contraction then runs in ms2_gemm against the weight re-laid-out to
 * [Cout, ldk], with bias and the pos-embed table as the epilogue residual. */
int ms2_patch_im2col(const void* img, int img_dt, void* cols, int B, int Hin, int Win, int ldk, ms2_stream_t stream);

/* ---- elementwise family (fp32 unless noted) */
/* y = a*x + b*z + c; x fp32 [n]; z fp32 [zn] or NULL, broadcast as z[i mod zn] (zn divides n);
 * y dtype y_dt.  Covers residual adds, `curr + 0.1*curr_pos` (memory_attention.py:137-138),
 * `mask*20-10` (sam2_base.py:418-420) and add-then-cast to the GEMM operand type. */
int ms2_axpby(const float* x, float a, const float* z, float b, float c, void* y, int y_dt, long n, long zn,
              ms2_stream_t stream);
/* y[b,:] = gate[b] > 0 ? x[b,:] : fill   (NO_OBJ_SCORE gating, sam2_base.py:354-363); P elements per row */
int ms2_gate_rows(const float* x, const float* gate, float fill, float* y, int B, long P, ms2_stream_t stream);
/* y[b,:] = x[b, idx[b], :]  (best-IoU / stability mask selection, sam2_base.py:376-383,
 * mask_decoder.py:289-316); x fp32 [B,M,P], idx int32 [B] (clamped to [0,M-1]) */
int ms2_select_plane(const float* x, const int32_t* idx, float* y, int B, int M, long P, ms2_stream_t stream);
/* y[m,c] = x[m,c] + s*v[c]  (row-vector broadcast) */
int ms2_add_rowvec(const float* x, const float* v, float s, float* y, long M, int C, ms2_stream_t stream);
/* dtype cast between fp32 and dt (n elements) */
int ms2_cast(const void* x, int x_dt, void* y, int y_dt, long n, ms2_stream_t stream);
/* y = act(x) */
int ms2_activation(const float* x, float* y, long n, int act, ms2_stream_t stream);
/* FPN top-down: fine[b,y,x,c] += coarse[b,y/2,x/2,c]  (image_encoder.py:113-124, nearest x2) */
int ms2_upsample2x_add(float* fine, const float* coarse, int B, int H, int W, int C, ms2_stream_t stream);
/* NHWC <-> NCHW fp32 transposes for the reference-facing API tensors */
int ms2_nhwc_to_nchw(const float* x, float* y, int B, int H, int W, int C, ms2_stream_t stream);
int ms2_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, ms2_stream_t stream);

/* ---- axial RoPE, in place on a strided [rows, D] matrix of dtype dt (position_encoding.py:167-216).
 *      row r uses table position (r mod table_len); cos/sin fp32 [table_len, D/2]; rows >= n_rope_rows
 *      of each batch are left untouched (object-pointer tokens, transformer.py:309-315). */
int ms2_rope(void* x, int dt, long batch_stride, long row_stride, int B, int rows, int n_rope_rows, int D,
             const float* cos_t, const float* sin_t, int table_len, ms2_stream_t stream);

/* ---- im2col for the small strided convs (memory_encoder.py:38-58, prompt_encoder.py:54-62,
 *      sam2_base.py:108): x fp32 NHWC [B,H,W,Cin] -> cols dtype dt [B*Ho*Wo, k*k*Cin]
 *      (tap order ky,kx,ci), zero padding.  pre: 0 none, 1 sigmoid, 2 (x>0); then x*pre_scale+pre_bias
 *      is applied to in-range inputs before padding (sam2_base.py:686-696). */
int ms2_im2col(const float* x, void* cols, int dt, int B, int H, int W, int Cin, int k, int stride, int pad,
               int pre, float pre_scale, float pre_bias, ms2_stream_t stream);

/* ---- fused conv 3x3 stride 2 pad 1 + LayerNorm2d + GELU for the thin first layers of the mask down-sampler
 *      (memory_encoder.py:38-58; Cin->Cout in {1->4, 4->16, 16->64}): x fp32 NHWC [B,H,W,Cin], w fp32 [Cout,Cin,3,3],
 *      y dtype y_dt NHWC [B,Ho,Wo,Cout]; `pre` as in ms2_im2col. */
int ms2_conv3x3s2_ln_gelu(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                          void* y, int y_dt, int B, int H, int W, int Cin, int Cout, float eps, int pre,
                          float pre_scale, float pre_bias, ms2_stream_t stream);

/* ---- depthwise conv 7x7 pad 3, NHWC fp32, w fp32 [C,7,7] (memory_encoder.py:84-90). */
int ms2_dwconv7x7(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int C,
                  ms2_stream_t stream);

/* ---- ConvTranspose2d k2 s2 epilogue (mask_decoder.py:66-73,233-236): g fp32 [B,H,W,4*C]
 *      ((dy,dx,c) column order, bias NOT yet added) -> out fp32 [B,2H,2W,C] = act(g + bias + skip). */
int ms2_pixel_shuffle_add(const float* g, const float* bias, const float* skip, float* out, int B, int H,
                          int W, int C, int act, ms2_stream_t stream);

/* ---- hypernetwork mask product (mask_decoder.py:247-248): masks[b,m,p] = sum_c hyper[b,m,c]*up[b,p,c]
 *      up dtype fp32 [B,P,C], hyper fp32 [B,Mk,C], masks fp32 [B,Mk,P]; C<=64, Mk<=8. */
int ms2_hyper_mask(const float* up, const float* hyper, float* masks, int B, int P, int C, int Mk,
                   ms2_stream_t stream);

/* ---- bilinear resize, align_corners=False, fp32 planes [N,H,W]->[N,Ho,Wo]
 *      (F.interpolate call sites sam2_base.py:368, video_predictor:736,831,844, transforms.py:98);
 *      antialias!=0 selects the triangle-filter variant (sam2_base.py:321-327,421-427). */
int ms2_resize_bilinear(const float* x, float* y, int N, int H, int W, int Ho, int Wo, int antialias,
                        ms2_stream_t stream);

/* ---- prompt encoder random-Fourier features (position_encoding.py:130-136,151-158):
 *      coords fp32 [n,2] already normalised to [0,1]; gauss fp32 [2,F]; out fp32 [n,2F] = [sin|cos]. */
int ms2_fourier_pe(const float* coords, const float* gauss, float* out, int n, int F, ms2_stream_t stream);

/* ---- sparse prompt embedding in one launch (prompt_encoder.py:79-101 `_embed_points`): coords fp32 [B,N,2] in input
 *      pixels, labels int32 [B,N], gauss fp32 [2,F], table fp32 [5,2F] = [not_a_point, point_embeddings 0..3];
 *      out fp32 [B,n_prefix+N+pad,2F]; pad != 0 appends the padding point (label -1) used when no box is given;
 *      prefix fp32 [n_prefix,2F] (or NULL) is copied in front of the prompt rows of every batch entry (the mask decoder's
 *      output tokens, mask_decoder.py:186-196: torch.cat of the tokens and the sparse prompts). */
int ms2_point_embed(const float* coords, const int* labels, const float* gauss, const float* table, float* out,
                    int B, int N, int pad, int F, int image_w, int image_h, const float* prefix, int n_prefix,
                    ms2_stream_t stream);

/* ---- frame ingest (utils/misc.py:92-101 `_load_img_as_tensor`, :215-244 `load_video_frames_from_data`,
 *      transforms.py:28-42): out NCHW = (x/255 - mean)/std, dtype out_dt (fp32 = the reference's frame dtype; bf16 = the
 *      cast autocast applies in front of the patch-embed conv, fused here).  x: in_layout 0 = fp32 NCHW (video tensor),
 *      1 = uint8 NHWC (decoded JPEG / image predictor), 2 = uint8 NCHW (uint8 video tensor). */
int ms2_normalize_image(const void* x, int in_layout, void* out, int out_dt, int B, int H, int W, ms2_stream_t stream);

/* ---- per-slice glue of the tracking step (the reference: a dozen tiny torch ops each)
 *      ms2_bank_rows (reference sam2_train/modeling/sam2_base.py:566-637: torch.cat of the recent memories and the
 *      object-pointer tokens, `+ maskmem_pos_enc + maskmem_tpos_enc`, cast): n <= 80 sources (HOST arrays of DEVICE
 *      pointers: h_src[i] fp32 [B,h_rows[i],W]; h_pos[i] fp32 [h_rows[i],W] with h_pos_bs[i] = 0, or [B,h_rows[i],W] with
 *      its batch stride, or NULL) are written back to back, in order: k_in[b] rows = src + pos, m_out[b] rows = src, both
 *      of dtype out_dt with batch strides k_bs / m_bs (elements); either destination may be NULL.  W % 4 == 0. */
int ms2_bank_rows(const void* const* h_src, const void* const* h_pos, const long* h_pos_bs, const int* h_rows, int n, int W,
                  int B, void* k_in, long k_bs, void* m_out, long m_bs, int out_dt, ms2_stream_t stream);
/*      ms2_argmax_select_rows (sam2_base.py:383-395): idx_out[b] = first arg-max of scores[b*scores_rs + 0..M) (may be NULL);
 *      out[b,0..C) = rows[b*rows_bs + idx*rows_rs + 0..C) (rows/out may be NULL). */
int ms2_argmax_select_rows(const float* scores, long scores_rs, int B, int M, const float* rows, long rows_bs, long rows_rs, int C,
                           int32_t* idx_out, float* out, ms2_stream_t stream);
/*      ms2_obj_ptr_mix (sam2_base.py:397-408): lam = soft ? sigmoid(logit) : (logit > 0);
 *      out = (fixed ? lam*ptr : ptr) + (1-lam)*no_obj;  ptr/out fp32 [B,C], logits fp32 [B], no_obj fp32 [C]. */
int ms2_obj_ptr_mix(const float* ptr, const float* logits, const float* no_obj, float* out, int B, int C, int soft,
                    int fixed, ms2_stream_t stream);
/*      ms2_stability_select (modeling/sam/mask_decoder.py:269-317 `_dynamic_multimask_via_stability`): counts int32 [B,2]
 *      from ms2_mask_stability_counts, ious fp32 [B,M]: idx = (area_i/area_u >= thresh or area_u == 0) ? 0 :
 *      1 + argmax(ious[b,1:]); iou_out[b] = ious[b,idx]. */
int ms2_stability_select(const int32_t* counts, const float* ious, int B, int M, float thresh, int32_t* idx_out,
                         float* iou_out, ms2_stream_t stream);

/* ---- row-complete GEMM tiles for the 256-wide memory-attention stream (csrc/gemm_row.cu): A bf16 [M,K] (lda), W bf16 [N,K],
 *      bias fp32 [N] or NULL; M a multiple of 128, K of 8.
 *      ms2_gemm_res_ln (memory_attention.py:58-99,166 `tgt = tgt + ...; tgt2 = self.norm_k(tgt)`), N = 256:
 *          x_out = residual + A W^T + bias (fp32 [M,256], may alias residual) and
 *          t_out = LayerNorm(x_out) * gamma + beta (biased variance, eps), t_dt MS2_BF16 or MS2_F32.
 *      ms2_gemm_rope (transformer.py:288-318 q_proj / k_proj + position_encoding.py:185-216 apply_rotary_enc), N a multiple
 *          of 256: out bf16 [M,N] = A W^T + bias rounded to bf16; the 256-wide column tiles below `rope_cols` are then
 *          rotated like ms2_rope with D = 256: row r uses table position ((r mod L) mod table_len), cos/sin fp32
 *          [table_len,128]. */
int ms2_gemm_res_ln(const void* A, long lda, const void* W, const float* bias, const float* residual, long ldr,
                    float* x_out, long ldx, const float* gamma, const float* beta, float eps, void* t_out, int t_dt,
                    long ldt, int M, int K, ms2_stream_t stream);
int ms2_gemm_rope(const void* A, long lda, const void* W, const float* bias, void* out, long ldo, int M, int N, int K,
                  int L, int rope_cols, const float* cos_t, const float* sin_t, int table_len, ms2_stream_t stream);

/*      ms2_multi_copy: n device-to-device copies (h_src[i] -> h_dst[i], h_bytes[i] bytes; the three tables are HOST arrays)
 *      in one launch per 16 items.  No reference counterpart: it replaces the per-tensor `copy_` / `clone` that a CUDA-graph
 *      replay of the SAM heads (sam2_base.py:249-412) and the memory encoder (:720-760) needs around its static buffers. */
int ms2_multi_copy(const void* const* h_src, void* const* h_dst, const long* h_bytes, int n, ms2_stream_t stream);

/* ---- memory-attention stack driven from ONE call per tracked frame (reference sam2_train/modeling/memory_attention.py:15-169,
 *      modeling/sam/transformer.py:266-331): host-side orchestration only — these entry points launch the kernels above in
 *      the reference's order (LN -> fused qkv GEMM -> RoPE -> self-attention -> out-proj + residual -> LN -> q GEMM + RoPE ->
 *      cross-attention over the memory bank -> folded value/output projection + residual -> LN -> FFN), so that the
 *      latency-bound tracked-frame path is not bound by the caller's per-launch overhead.  Shipped configuration only
 *      (one head of C = 256 over Cm = 64-d memories, bf16 operands, ReLU FFN).  `layers` points to `n_layers` consecutive
 *      ms2_memattn_layer_w records (device pointers: weights bf16 [N,K] row-major, everything else fp32). */
typedef struct ms2_memattn_layer_w {
  const float *norm1_g, *norm1_b, *norm2_g, *norm2_b, *norm3_g, *norm3_b;
  const void* qkv_w; const float* qkv_b;   /* self-attention q|k|v fused [3C, C] */
  const void* so_w; const float* so_b;     /* self-attention out_proj [C, C] */
  const void* cq_w; const float* cq_b;     /* cross-attention q_proj [C, C] */
  const void* ck_w; const float* ck_b;     /* cross-attention k_proj [C, Cm] */
  const void* vo_w; const float* vo_b;     /* cross-attention out_proj . v_proj folded [C, Cm] */
  const void* f1_w; const float* f1_b;     /* FFN linear1 [F, C] */
  const void* f2_w; const float* f2_b;     /* FFN linear2 [C, F] */
  const float *rope_cos, *rope_sin;        /* axial RoPE table [rope_len, C/2] */
  float eps1, eps2, eps3;
  int C, Cm, F, rope_len;
} ms2_memattn_layer_w;
long ms2_memattn_workspace_bytes(int B, int L, int C, int Cm, int F);
/*      x fp32 [B,L,C] (in place) -> ... -> q bf16 [B,L,C] (RoPE applied) of the layer's cross-attention */
int ms2_memattn_layer_pre(const void* layer, float* x, void* q_out, void* workspace, long workspace_bytes, int B, int L,
                          ms2_stream_t stream);
/*      att bf16 [B,L,Cm] = softmax(q K^T) M  ->  x updated in place through the rest of the layer */
int ms2_memattn_layer_post(const void* layer, float* x, const void* att, void* workspace, long workspace_bytes, int B, int L,
                           ms2_stream_t stream);
/*      per-frame rows of the memory bank (reference sam2_base.py:566-637): sources as in ms2_bank_rows -> raw values into
 *      m_bank rows [row0, row0+n), per layer K = k_proj(src + pos) into h_K[l] rows [row0, row0+n) (HOST array of device
 *      pointers, batch stride k_bs elements), RoPE on the first n_rope_rows of them */
int ms2_memattn_bank_project(const void* layers, int n_layers, const void* const* h_src, const void* const* h_pos,
                             const long* h_pos_bs, const int* h_rows, int n_src, int B, int n_rope_rows, int Lq,
                             void* const* h_K, long k_bs, int row0, void* m_bank, long m_bs, void* workspace,
                             long workspace_bytes, ms2_stream_t stream);
/*      the whole stack on one GPU: x = curr + pos_scale * curr_pos (curr_pos may be NULL); layers; final LayerNorm -> out */
int ms2_memattn_forward(const void* layers, int n_layers, const float* curr, const float* curr_pos, float pos_scale,
                        void* const* h_K, long k_bs, const void* m_bank, long m_bs, int Lk, const float* norm_g,
                        const float* norm_b, float norm_eps, float* x, float* out, void* workspace, long workspace_bytes,
                        int B, int L, ms2_stream_t stream);

/* ---- mask statistics for the stability fallback (mask_decoder.py:269-317): per (b) plane of fp32
 *      logits [N,P]: counts[n,0] = #(x>delta), counts[n,1] = #(x>-delta). */
int ms2_mask_stability_counts(const float* x, int32_t* counts, int N, long P, float delta, ms2_stream_t stream);

/* ---- segmentation metric counts in one pass (reference repo root: func_3d/utils.py:139-202 `eval_seg`, :204-214 `iou`,
 *      :215-240 `dice_coeff`; call site func_3d/function.py:300): pred, gt fp32 [N,P] device planes, thr_host = T <= 8
 *      thresholds in HOST memory; counts int32 [N,T,3] = (#(pred>th & gt>th), #(pred>th), #(gt>th)), exact. */
int ms2_seg_counts(const float* pred, const float* gt, const float* thr_host, int T, int32_t* counts, int N, long P,
                   ms2_stream_t stream);

/* ---- validation loss in one pass (reference repo root: func_3d/function.py:35-36,299 `BCEWithLogitsLoss(pos_weight)`):
 *      pred (logits), gt fp32 [N,P] device planes; sums fp64 [N] (device) = per-plane SUM of the element losses
 *      (1-y)*x + (1+(pos_weight-1)*y)*(log1p(exp(-|x|)) + max(-x,0)); the caller divides by the element count. */
int ms2_bce_logits_sum(const float* pred, const float* gt, float pos_weight, double* sums, int N, long P,
                       ms2_stream_t stream);

/* ---- the same scoring straight from LOW-resolution logits (reference repo root: func_3d/function.py:283-305 scores the
 *      video-resolution logits that sam2_train/sam2_video_predictor.py:724-744 `_get_orig_video_res_output` up-samples
 *      with F.interpolate(bilinear, align_corners=False)): low fp32 [N,h,w] device planes are up-sampled to H x W on the
 *      fly (bit-identical to ms2_resize_bilinear), compared with gt fp32 [N,H,W]; counts int32 [N,T,3] as ms2_seg_counts,
 *      sums fp64 [N] as ms2_bce_logits_sum (NULL: skip the loss).  W % 4 == 0.  The video-resolution logits are never
 *      written: one pass over gt. */
int ms2_score_lowres(const float* low, const float* gt, const float* h_thr, int T, float pos_weight, int32_t* counts,
                     double* sums, int N, int h, int w, int H, int W, ms2_stream_t stream);

/* ---- non-overlapping constraints (reference sam2_train/modeling/sam2_base.py:812-830 `_apply_non_overlapping_constraints`,
 *      callers sam2_video_predictor.py:742-743, :850-851): in/out fp32 [n_obj, P] device planes (out != in); per pixel the
 *      FIRST arg-max object keeps its score, every other object gets min(score, -10). */
int ms2_non_overlap(const float* in, float* out, int n_obj, long P, ms2_stream_t stream);

/* ---- automatic mask generator, per-plane statistics in one pass (automatic_mask_generator.py:300-340; utils/amg.py:158-180
 *      `calculate_stability_score`, :296-348 `batched_mask_to_box`): x fp32 logits [N,H,W]; stats int32 [N,7] =
 *      (#(x>thr+off), #(x>thr-off), #(x>thr), min col, min row, max col, max row of x>thr); an empty plane gives
 *      (.., 0, W, H, -1, -1). */
int ms2_mask_stats(const float* x, int32_t* stats, int N, int H, int W, float thr, float off, ms2_stream_t stream);

/* ---- binarise + un-crop + transpose selected planes for run-length encoding (utils/amg.py:279-293 `uncrop_masks`,
 *      :107-134 `mask_to_rle_pytorch`): out uint8 [K,OW,OH], out[k, x0+c, y0+r] = x[sel[k], r, c] > thr, 0 elsewhere. */
int ms2_mask_binarize_t(const float* x, const int32_t* sel, uint8_t* out, int K, int H, int W, float thr, int OH, int OW,
                        int x0, int y0, ms2_stream_t stream);

/* ---- run-length encoding on the device (utils/amg.py:107-134 `mask_to_rle_pytorch`): m uint8 [K,L] (0/1 masks already in
 *      encoding order); pos int32 [K,cap] receives, in increasing order, the positions p >= 1 with m[k,p] != m[k,p-1];
 *      cnt int32 [K] = the TOTAL number of such positions (if cnt[k] > cap the tail was dropped: call again with a larger cap). */
int ms2_rle_transitions(const uint8_t* m, int32_t* pos, int32_t* cnt, int K, long L, int cap, ms2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
