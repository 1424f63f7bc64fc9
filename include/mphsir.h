/*
 * mphsir.h — C ABI of libmphsir.so, the B200 (sm_100a) kernel library behind the drop-in
 * MP_HSIR_Net module (mp_hsir_b200/model.py).
 *
 * The reference (ZhehuiWu/MP-HSIR) has no FFI: its hot path is the Python nn.Module
 * net/MP_HSIR.py, every op an ATen call.  Each entry point below therefore names the
 * reference *Python* interface it replaces (file:line of net/MP_HSIR.py); INTEGRATION.md
 * shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - Activations are fp32, token-major ("channels-last"): row n = (b*H + y)*W + x, `ld` floats
 *     per row.  This is the layout PGSSTB itself uses (net/MP_HSIR.py:665-668).
 *   - Every pointer is a CUDA device pointer owned by the caller (PyTorch allocates); the
 *     library allocates nothing and keeps no state besides the last error string.
 *   - Calls enqueue asynchronously on `stream` (a cudaStream_t passed as void*), never
 *     synchronise, are re-entrant and CUDA-graph capturable.
 *   - Return 0 on success; non-zero -> mphsir_last_error() describes it.  No CPU fallback.
 */
#ifndef MPHSIR_H_
#define MPHSIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPHSIR_VERSION 100 /* round 1 */

#if defined(__GNUC__)
#define MPHSIR_API __attribute__((visibility("default")))
#else
#define MPHSIR_API
#endif

enum { MPHSIR_OK = 0, MPHSIR_ERR_INVALID = 1, MPHSIR_ERR_CUDA = 2 };

MPHSIR_API int mphsir_version(void);
MPHSIR_API const char* mphsir_last_error(void);
/* sm_100 check + SM count; returns MPHSIR_ERR_CUDA if the device is not compute capability 10.x */
MPHSIR_API int mphsir_device_check(int device, int* sm_count);

/* ---------------------------------------------------------------------------------------
 * Layout conversion at the module boundary: NCHW image -> token-major rows, channels
 * [C, ld_out) zero-filled.  Replaces the implicit layout of `inp_img` (net/MP_HSIR.py:810-814).
 * ------------------------------------------------------------------------------------- */
MPHSIR_API int mphsir_nchw_to_tokens(const float* in, float* out, int B, int C, int HW, int ld_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * GEMM family  Y = epilogue( prologue(A) @ Bt )   A:[M,K] token-major, Bt:[Kp,ldb] (pre-packed
 * "in x out" weights, Kp = K rounded up to 16, zero padded).
 *
 * prologue : optional LayerNorm over the K columns of each row (ln_gamma/ln_beta != NULL),
 *            replacing nn.LayerNorm (net/MP_HSIR.py:618-619,667,719) and WithBias_LayerNorm
 *            (:354-357) in front of the projection that consumes it.
 * epilogue : see enum below.  Replaces nn.Linear / 1x1 nn.Conv2d call sites:
 *            Spatial_Attention.qkv/.proj (:195,:216), Spectral_Attention.qkv/.project_out
 *            (:98,:113), GatedMlp.fc1/.fc2 (:77-80), FeedForward/FFN project_in/out (:387-390,
 *            :261-264), CrossAttention q/kv/project_out (:236-248), PromptFusion.conv (:597),
 *            reduce_chan_level2 (:828), and the residual adds of PGSSTB.forward (:715-719),
 *            TransformerBlock.forward (:476-477), BaseBlock.forward (:760).
 * ------------------------------------------------------------------------------------- */
/* arithmetic of the contraction (accumulation is always fp32) */
enum {
  MPHSIR_PREC_FP32_SIMT = 0, /* FFMA, bit-for-bit fp32 products (gemm_simt.cu)                            */
  MPHSIR_PREC_BF16X3 = 1,    /* tcgen05, operands split hi+lo bf16: hi*hi + hi*lo + lo*hi (fp32-grade)    */
  MPHSIR_PREC_BF16 = 2       /* tcgen05, operands rounded to bf16                                          */
};

enum {
  MPHSIR_EPI_BIAS = 0,     /* Y = acc + bias                                                     */
  MPHSIR_EPI_RESIDUAL = 1, /* Y = res1 + scale_b*(acc + bias) [+ res2]                           */
  MPHSIR_EPI_GLU = 2,      /* packed cols (2j,2j+1)=(value_j,gate_j): Y[:,j] = v*gelu_erf(g)     */
  MPHSIR_EPI_SPECTRAL = 3, /* Y = res1 + scale_b*(gsrc*gate[window(row)] + acc)   (:715-718)     */
  /* attention proj with the global-spectral qkv 1x1 folded behind it (one GEMM over the window-attention
   * output, weight rows [W_proj ; W_sqkv*W_proj]):  cols n <  n_split: Y[m,n] = res1 + scale_b*(acc+bias)*gate[window(row),n]
   * (the shortcut plus the locally gated spatial branch, :715-718 / :153);  cols n >= n_split: Y2[m,n-n_split] = acc+bias */
  MPHSIR_EPI_PROJ = 4
};

typedef struct {
  const float* A;      /* [M, lda] */
  int lda;
  int a_row_mod;       /* >0: A row index = m % a_row_mod (operand shared by all samples)        */
  const float* Bt;     /* [Kp, ldb]; ldb multiple of 64, >= N rounded up to the tile              */
  int ldb;
  long long b_batch_stride; /* floats between per-sample weight matrices; 0 = shared            */
  int rows_per_batch;  /* H*W; needed when b_batch_stride != 0 or epi needs sample / window ids  */
  float* Y;
  int ldy;
  int M, N, K;         /* K = number of valid A columns (multiple of 4)                           */
  const float* ln_gamma; /* [K] or NULL */
  const float* ln_beta;
  const float* bias;   /* [N] (packed order) or NULL */
  int epi;
  const float* res1;
  int ldr1;
  const float* res2;   /* optional second residual (BaseBlock shortcut) */
  int ldr2;
  const float* gsrc;   /* SPECTRAL: spatial-attention output sa [M, ldg] */
  int ldg;
  const float* gate;   /* SPECTRAL: per-window channel gate [B*nW, N]    */
  int H, W, shift;     /* SPECTRAL: image size at this level and cyclic shift (0 or 4)            */
  const float* row_scale; /* [B] DropPath keep/keep_prob per sample, NULL = 1 (eval)              */
  int precision;       /* MPHSIR_PREC_*: FP32_SIMT uses Bt; the tensor-core modes use Bimg             */
  const void* Bimg;    /* packed bf16 hi/lo weight image (mphsir_pack_bimg) of the logical [N,K] matrix */
  long long bimg_batch_bytes; /* bytes between per-sample images; 0 = shared                          */
  float* Y2;           /* PROJ: second output [M, ldy2] for columns >= n_split                           */
  int ldy2;
  int n_split;         /* PROJ: multiple of 32; gate is [B*nW, n_split]                                  */
} mphsir_gemm_params;

/* Tensor-core weight image: logical W[N,K] fp32 (row n at W + n*ld, or W + k*ld + n when
 * `transposed`) -> bf16 hi/lo parts in the 128-byte-swizzled K-major layout the MMA reads
 * (mp_hsir_b200/csrc/gemm_tc.cuh: bimg_offset).  `batch` independent matrices, w_batch_stride floats
 * apart, produce images mphsir_bimg_bytes(N,K) bytes apart.  img must be 128-byte aligned. */
/* Debug: when buf != NULL every subsequent tensor-core GEMM launch writes per-CTA cycle counters of its
 * warp roles into buf[grid][16] (tools/gemm_bench.py --profile).  Pass NULL to switch it off. */
MPHSIR_API void mphsir_debug_tc_counters(long long* buf);
/* Debug: 0 disables the CTA-pair (cta_group::2) instantiation of the tensor-core GEMM (default 1: tensor-heavy shapes use it). */
MPHSIR_API void mphsir_debug_tc_cluster(int enabled);
MPHSIR_API void mphsir_debug_tc_reverse(int enabled);          /* 0: tensor-core GEMM launches never walk their row tiles backwards (default 1: launches over >= 131072 rows do — their input is then read starting with the part the producer wrote last, which is still in L2) */
/* How mphsir_gemm_fwd (tensor-core precisions) cuts a [M,K] x [K,N] launch into work on a device with `sm_count` SMs — pure host
 * arithmetic, no device needed: out6[8] = {CTAs per cluster (2 = cta_group::2 pairs), pass groups per SPLIT row tile, passes per
 * group, grid size, work-item iterations per CTA, 1 if the tiles are walked backwards, number of leading row tiles that stay whole,
 * accumulator columns per pass (256 or 128)}. */
MPHSIR_API int mphsir_gemm_plan(int M, int N, int K, int rows_per_batch, int per_sample_weights, int sm_count, int* out6);
MPHSIR_API void mphsir_debug_tc_psplit(int enabled);           /* 0: never hand the 256-column passes of a row tile to several CTAs (A/B switch; default 1: few-tile GEMMs do) */
MPHSIR_API void mphsir_debug_tc_ebox1(int enabled);         /* 0: two store boxes per epilogue warp everywhere (default 1: one box + 4-slot A ring for BIAS GEMMs with 64 < K <= 128) */
MPHSIR_API void mphsir_debug_pdl(int enabled);              /* 1: programmatic dependent launch of the persistent tcgen05 kernels (default 0: measured, no gain) */
MPHSIR_API void mphsir_debug_tc_tma_epilogue(int enabled);  /* 0: register-staged GEMM epilogue everywhere (default 1: TMA boxes) */
MPHSIR_API void mphsir_debug_window_attn_tc(int enabled);   /* 0: mma.sync window attention everywhere (default 1: TMA-fed tcgen05 kernel for head dims 32 / 64) */
MPHSIR_API void mphsir_debug_window_attn_tc_counters(long long* buf);   /* [grid][16] per-CTA role cycle counters of the tcgen05 window attention (NULL: off) */
MPHSIR_API void mphsir_debug_dwgram_tma(int enabled);       /* 0: use the direct-load dwconv+Gram kernel everywhere (default 1) */
MPHSIR_API void mphsir_debug_mlp_counters(long long* buf); /* same idea for the fused MLP kernel */
MPHSIR_API void mphsir_debug_mlp_flags(int flags);          /* timing experiments (wrong results!): 16 no weight copies, 32 no conversion, 64 no GLU work, 128 no final epilogue, 256 no X loads, 4096 no MMAs */
MPHSIR_API size_t mphsir_bimg_bytes(int N, int K);
MPHSIR_API int mphsir_pack_bimg(const float* W, int ld, int transposed, long long w_batch_stride, void* img,
                                int batch, int N, int K, void* stream);

/* The same for many (un-batched) matrices in one launch per 64 items: the trainer re-packs every weight after each
 * optimizer step. */
typedef struct {
  const float* W;
  void* img;
  int ld, transposed, N, K;
} mphsir_pack_item;
MPHSIR_API int mphsir_pack_bimg_multi(const mphsir_pack_item* items, int count, void* stream);

MPHSIR_API int mphsir_gemm_fwd(const mphsir_gemm_params* p, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused gated MLP (tensor-core precisions): Y = X + s_b * (fc2(value * gelu(gate)) + b2) [+ res2] with
 * [value|gate] = LayerNorm(X) W1^T + b1 — PGSSTB.forward :719 + GatedMlp.forward :76-82 in ONE kernel; the
 * 2*hidden intermediate stays in TMEM / shared memory.  W1img: image (mphsir_pack_bimg) of the interleaved
 * fc1 matrix [2*hid_pad, C] ((2j,2j+1) = (value_j, gate_j), as for MPHSIR_EPI_GLU), b1 in the same order;
 * W2img: image of fc2 [C, hid_pad].  Supported shapes: mphsir_mlp_supported(C, hid_pad).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const float* X;
  int ldx;
  const float *ln_gamma, *ln_beta;
  const void* W1img;
  const float* b1;
  const void* W2img;
  const float* b2;
  const float* res2; /* optional second residual (BaseBlock shortcut) */
  int ldr2;
  const float* row_scale; /* DropPath keep/keep_prob per sample or NULL */
  int rows_per_batch;
  float* Y;
  int ldy;
  int M, C, hid_pad;
  int precision; /* MPHSIR_PREC_BF16X3 or MPHSIR_PREC_BF16 */
} mphsir_mlp_params;

MPHSIR_API int mphsir_mlp_supported(int C, int hid_pad);
MPHSIR_API int mphsir_mlp_fwd(const mphsir_mlp_params* p, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense 3x3 convolution (zero pad 1, no bias) as an implicit GEMM over token-major input.
 * Replaces OverlapPatchEmbed.proj (:458), Downsample/Upsample bodies incl. PixelUnshuffle /
 * PixelShuffle (:436-437,:446-447), TVSP.conv_last (:581) and `output(x)+inp_img` (:841).
 * Wt is packed [9*Cin, ldb]: row = tap*Cin + c, tap = 3*(dy+1)+(dx+1).
 * ------------------------------------------------------------------------------------- */
enum {
  MPHSIR_CONV_TOKENS = 0,    /* Y[n, c]                                                          */
  MPHSIR_CONV_UNSHUFFLE = 1, /* Y[(b,y/2,x/2), c*4+2(y&1)+(x&1)]                                 */
  MPHSIR_CONV_SHUFFLE = 2,   /* packed col q*Cn+cn (q=2i+j) -> Y[(b,2y+i,2x+j), cn]              */
  MPHSIR_CONV_NCHW_RES = 3   /* Y[b,c,y,x] = acc + R[b,c,y,x]  (NCHW, c < N)                     */
};

typedef struct {
  const float* X;  /* [B*H*W, ldx] token-major, channels [0,Cin) valid (Cin multiple of 16)      */
  int ldx;
  const float* Wt; /* [9*Cin, ldb] */
  int ldb;
  float* Y;
  int ldy;
  int B, H, W, Cin, N; /* N = number of output channels */
  int out_mode;
  const float* R;  /* NCHW residual for MPHSIR_CONV_NCHW_RES */
  int precision;   /* MPHSIR_PREC_* */
  const void* Bimg; /* packed image of the logical [N, 9*Cin] matrix for the tensor-core modes */
} mphsir_conv3x3_params;

MPHSIR_API int mphsir_conv3x3_fwd(const mphsir_conv3x3_params* p, void* stream);

/* ---------------------------------------------------------------------------------------
 * Window attention core: for every (shifted) 8x8 window and head
 *     softmax(q k^T * hd^-0.5 + rel_pos_bias + shift_mask) v
 * qkv is the token-major [B*H*W, 3C] output of the LN+qkv GEMM in *image* order; the roll,
 * window_partition, window_reverse and un-roll of PGSSTB.forward (:671-678,:689-696) are
 * folded into the addressing, the Swin mask of calculate_mask (:639-660) is evaluated in
 * closed form, and the per-window token mean of the result (needed by the local spectral
 * branch, :135) is emitted as well.  Replaces Spatial_Attention.forward :195-215 (proj is a GEMM).
 *   bias : [heads,64,64] pre-gathered relative-position bias (:200-202)
 *   out  : [B*H*W, ldo] image order;  win_mean : [B*nW, C] in shifted-window order
 * ------------------------------------------------------------------------------------- */
MPHSIR_API int mphsir_window_attn_fwd(const float* qkv, int ldqkv, const float* bias, float* out, int ldo,
                           float* win_mean, int B, int H, int W, int C, int heads, int shift,
                           int precision /* MPHSIR_PREC_*: SIMT fp32, or tensor cores with bf16x3 / bf16 operands */,
                           void* stream);
/* Row band of a scene sharded over GPUs (mp_hsir_b200/sharded.py): the [H, W] image handed to the kernel is rows
 * [y0, y0 + H) (cyclic) of a scene of mask_H rows, and the Swin mask of calculate_mask (:643-658) is evaluated in
 * the shifted coordinates of the SCENE: local shifted row ys is scene row (ys + mask_y0) mod mask_H.  The roll wraps
 * inside the band (those windows belong to the halo and are recomputed by the neighbour rank).
 * mphsir_window_attn_fwd == mask_H = H, mask_y0 = 0.  Tensor-core precisions only. */
MPHSIR_API int mphsir_window_attn_band_fwd(const float* qkv, int ldqkv, const float* bias, float* out, int ldo,
                           float* win_mean, int B, int H, int W, int C, int heads, int shift, int precision,
                           int mask_H, int mask_y0, void* stream);
/* The same operator as a TMA-fed tcgen05 kernel (window_attn_tc.cu: two windows per 128-row MMA tile, S / P / O in tensor
 * memory, P the TMEM A operand of P V) for head dims 32 and 64 (mphsir_window_attn_tc_supported).  bias_t is the
 * relative-position bias TRANSPOSED, [heads][key 64][query 64], so that a warp's 32 queries read one line per key. */
MPHSIR_API int mphsir_window_attn_tc_supported(int head_dim);
MPHSIR_API int mphsir_window_attn_tc_enabled(void);   /* debug switch state (mphsir_debug_window_attn_tc) */
MPHSIR_API int mphsir_window_attn_tc_fwd(const float* qkv, int ldqkv, const float* bias_t, float* out, int ldo,
                           float* win_mean, int B, int H, int W, int C, int heads, int shift, int precision,
                           int mask_H, int mask_y0, void* stream);

/* ---------------------------------------------------------------------------------------
 * Local spectral branch (low-rank spectral-prompt gate), one fused kernel per window:
 * PG_Spectral_Attention.forward :135-152 collapsed to a per-window channel gate g[B_,C];
 * the caller applies sa*g in the MPHSIR_EPI_SPECTRAL epilogue (:153).
 *   core_mean [B_,C] : token mean of the attention core (before proj)
 *   all weights pre-transposed to "in x out".  Spatial_Attention.proj (the mean over tokens commutes
 *   with it) is folded into the two matrices that consume its output:
 *     promptT[C,128] = projT @ linear_prompt^T,  promptb[128] = linear_prompt @ projb
 *     downT[C,r]     = projT @ linear_down^T,    downb[r]     = linear_down @ projb
 *   param[128,r], qT[r,r], kvT[r,2r], p2T[r,r], p2b[r], upT[r,C]
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const float* core_mean;
  const float *promptT, *promptb, *downT, *downb, *param, *qT, *kvT, *p2T, *p2b, *upT;
  float* gate; /* [B_, C] */
  int B_, C, r;
} mphsir_local_gate_params;

MPHSIR_API int mphsir_local_gate_fwd(const mphsir_local_gate_params* p, void* stream);
/* Two-step form: the caller first computes logits[B_, ldl] = core_mean x [promptT | downT] + [promptb | downb]
 * (128 prompt logits, then the r low-rank projections; one mphsir_gemm_fwd call) and this kernel finishes
 * :136-152 with one warp per window.  core_mean, promptT/b and downT/b of *p are ignored. */
MPHSIR_API int mphsir_local_gate_tail_fwd(const float* logits, int ldl, const mphsir_local_gate_params* p, void* stream);
/* One-launch form of the same chain (8-32 windows per CTA: fp32 register-tiled C-long products, then the r-sized remainder in
 * the same warp); needs every field of *p like mphsir_local_gate_fwd, r a multiple of 4.  What the inference engine calls. */
MPHSIR_API int mphsir_local_gate2_fwd(const mphsir_local_gate_params* p, void* stream);

/* ---------------------------------------------------------------------------------------
 * Depthwise 3x3 conv (zero pad 1, no bias) on token-major data, optional GDFN gate.
 * Replaces qkv_dwconv (:98), q_dwconv/kv_dwconv (:236-237), FeedForward.dwconv + gelu gate
 * (:388-389, :262-263).   w9 packed [9, C].
 *   gate_half = 0 : Y[n,c] = dw(X)[n,c]                                 (Y has C columns)
 *   gate_half = h : Y[n,j] = gelu_erf(dw[n,j]) * dw[n,h+j],  j < h      (C == 2h)
 * ------------------------------------------------------------------------------------- */
MPHSIR_API int mphsir_dwconv3x3_fwd(const float* X, int ldx, const float* w9, float* Y, int ldy, int B, int H,
                         int W, int C, int gate_half, void* stream);

/* ---------------------------------------------------------------------------------------
 * Global spectral ("transposed") attention statistics and weight folding, replacing
 * Spectral_Attention.forward :101-113 / Attention.forward :412-426 / CrossAttention :239-248.
 *  1. gram_partial : per (sample, head, token chunk) partial  q^T k [c,c], sum q^2 [c], sum k^2 [c]
 *  2. gram_softmax : reduce partials, A = softmax_j( G_ij / (max(|q_i|,eps) max(|k_j|,eps)) * T_h )
 *  3. fold         : Mt_b[h*c+j, o] = sum_i WoutT[h*c+i, o] * A_h[i, j]   ("in x out" for the
 *                    apply GEMM, which then computes project_out(attn @ v) in one pass)
 * q,k are column slices of token-major buffers; *_shared != 0 means the operand has a single
 * sample that every batch element uses (TVSP visual prompt).
 * ------------------------------------------------------------------------------------- */
MPHSIR_API size_t mphsir_gram_partial_floats(int B, int heads, int c, int HW, int* n_chunks);
MPHSIR_API int mphsir_gram_partial_fwd(const float* q, int ldq, int q_shared, const float* k, int ldk,
                            int k_shared, float* partial, int B, int HW, int heads, int c,
                            void* stream);
MPHSIR_API int mphsir_gram_softmax_fwd(const float* partial, int n_chunks, const float* temperature,
                            float* attn /* [B,heads,c,c] */, float* scratch /* [B*heads*(c*c+2c)] */, int B,
                            int heads, int c, void* stream);
MPHSIR_API int mphsir_spectral_fold_fwd(const float* attn, const float* WoutT /* [C,C] in x out */,
                             float* Mt /* [B, Cp, ldm] */, int ldm, long long m_batch_stride, int B,
                             int heads, int c, void* stream);

/* Fused step 0+1 for the tensor-core precisions: depthwise 3x3 of the [q|k|v] 1x1 output X [B*H*W, >=3C] +
 * Gram partials; only v is written (V [B*H*W, ldv]); q and k never reach HBM.  partial layout as above with
 * n_chunks = CTAs per sample (mphsir_dwgram_partial_floats).  Supported (C, C/heads): see
 * mphsir_dwgram_supported; otherwise use dwconv3x3 + gram_partial. */
MPHSIR_API int mphsir_dwgram_supported(int C, int c);
MPHSIR_API size_t mphsir_dwgram_partial_floats(int B, int heads, int c, int H, int W, int* n_chunks);
MPHSIR_API int mphsir_dwgram_fwd(const float* X, int ldx, const float* w9 /* [9, 3C] */, float* V, int ldv,
                                 float* partial, int B, int H, int W, int C, int heads, int precision, void* stream);
/* Row band of a sharded scene: conv and V cover all H rows handed in (own rows + halo), but only 8x8 tiles whose first
 * row lies in [gram_y0, gram_y1) (multiples of 8) enter the Gram statistics — each rank contributes its OWN rows to the
 * reduction over the whole scene (net/MP_HSIR.py:104-110).  mphsir_dwgram_fwd == [0, H). */
MPHSIR_API int mphsir_dwgram_band_fwd(const float* X, int ldx, const float* w9 /* [9, 3C] */, float* V, int ldv,
                                      float* partial, int B, int H, int W, int C, int heads, int precision,
                                      int gram_y0, int gram_y1, void* stream);
/* Step 2a alone: reduced[B*heads, c*c+2c] = sum over chunks of partial — the per-rank statistics a sharded scene
 * sums across GPUs before mphsir_spectral_finish_fwd(n_chunks = 1). */
MPHSIR_API int mphsir_gram_reduce(const float* partial, int n_chunks, float* reduced, int B, int heads, int c, void* stream);

/* Steps 2+3 in one call (what the module uses): reduce the partials (into `scratch` when n_chunks > 1),
 * normalise + temperature + softmax, fold project_out, and write the per-sample matrix as fp32 "in x out"
 * Mt [B, Cp, ldm] (may be NULL) and/or as the tensor-core image (may be NULL; see mphsir_pack_bimg).
 * attn_out [B,heads,c,c] is optional (tests). */
MPHSIR_API int mphsir_spectral_finish_fwd(const float* partial, int n_chunks, float* scratch, const float* temperature,
                                          const float* WoutT, float* Mt, int ldm, long long m_batch_stride, void* bimg,
                                          long long bimg_batch_bytes, float* attn_out, int B, int heads, int c,
                                          void* stream);

/* ---------------------------------------------------------------------------------------
 * TVSP helpers.
 *  tvsp_query : Q[b,i,j,d] = tp[b,d] * clip_b[floor(i*B/ps), floor(j*512/ps)] with
 *               tp = (weights @ learnable)/T  — the broadcast + nearest interpolate of :575-577.
 *  bilinear   : F.interpolate(mode="bilinear", align_corners=False) on token-major data (:580).
 * ------------------------------------------------------------------------------------- */
MPHSIR_API int mphsir_tvsp_query_fwd(const float* clip_b /* [B,512] */, const float* weights /* [B,T] */,
                          const float* learnable /* [T,D] */, float* Q /* [B*ps*ps, D] */, int B,
                          int T, int D, int ps, void* stream);
MPHSIR_API int mphsir_bilinear_fwd(const float* X, int ldx, float* Y, int ldy, int B, int h, int w, int H,
                        int W, int C, void* stream);

/* Text_Prompt.forward :517-532: clip_b[B,512] = (weights @ clip)/T */
MPHSIR_API int mphsir_text_prompt_fwd(const float* weights /* [B,T] */, const float* clip /* [T,512] */,
                           float* clip_b, int B, int T, void* stream);


/* =======================================================================================
 * Training path: backward of MP_HSIR_Net.forward (what autograd derives from net/MP_HSIR.py
 * for train.py:50-67), clamp+L1 loss and AdamW (train.py:69).  Data-gradient GEMMs / convs reuse
 * mphsir_gemm_fwd / mphsir_conv3x3_fwd / mphsir_dwconv3x3_fwd with transposed (flipped) weights;
 * the entry points below are the operations that have no forward twin.  Gradient buffers are
 * caller-owned, pre-zeroed where the text says "+=" (fp32 atomics accumulate into them).
 * ===================================================================================== */

/* output-row maps: how a packed activation column o lands in the reference parameter layout */
enum {
  MPHSIR_MAP_IDENTITY = 0,   /* row o, valid o < a                                                        */
  MPHSIR_MAP_INTERLEAVE = 1, /* packed (2j,2j+1) = (first-half_j, second-half_j): row (o&1)*a + j, j < a  */
  MPHSIR_MAP_HALVES = 2      /* halves at [0,a) and [b,b+a): rows o / a + (o-b)                            */
};

/* Weight gradient on the tensor cores: dW[map(o)*so + i*si + tap*st] += sum_m dY[m,o] * X[src(m), i].
 *   plain          : src(m) = m                                  (nn.Linear / 1x1 conv weights)
 *   rows_per_batch : one output matrix per sample, dw_batch_stride floats apart (dU^T v of the spectral attention)
 *   taps = 9       : src(m) = pixel shifted by (dy,dx), zero outside the image (dense 3x3 conv weights [O,I,3,3]:
 *                    so = 9*I, si = 9, st = 1)
 *   x_row_mod > 0  : X has x_row_mod rows shared by every sample (TVSP visual prompt)
 * O, I multiples of 4; i >= i_valid (0 = I) is padding.  precision: MPHSIR_PREC_BF16X3 or MPHSIR_PREC_BF16. */
typedef struct {
  const float* dY;
  int lddy;
  const float* X;
  int ldx;
  float* dW;
  long long M;
  int O, I;
  int rows_per_batch;
  long long dw_batch_stride;
  int x_row_mod;
  int H, W, taps;
  long long so, si, st;
  int map_mode, map_a, map_b;
  int i_valid;
  int precision;
  float* dbias; /* optional: dbias[map(o)] += sum_m dY[m,o] (the bias gradient of the same layer, exact fp32 sums taken from the
                   dY tiles on their way through registers); plain mode only */
} mphsir_wgrad_params;
MPHSIR_API int mphsir_wgrad(const mphsir_wgrad_params* p, void* stream);
/* Up to 8 small plain-mode problems (no taps / per-sample / shared X, one precision) in ONE launch of the mma.sync kernel:
 * the seven r-sized weight gradients of the local spectral gate. */
MPHSIR_API int mphsir_wgrad_multi(const mphsir_wgrad_params* list, int count, void* stream);
/* Debug: force the weight-gradient engine: 0 = mma.sync (wgrad.cu), 1 = tcgen05 (wgrad_tc.cu), -1 = per-shape choice (default). */
MPHSIR_API void mphsir_debug_wgrad_tc(int enabled);

/* bias gradients: out[map(c)] += sum_m X[m,c] */
MPHSIR_API int mphsir_colsum(const float* X, int ldx, float* out, long long M, int C, int map_mode, int map_a, int map_b,
                             void* stream);

/* LayerNorm (nn.LayerNorm :618-619 / WithBias_LayerNorm :354-357) materialised for training: Y = LN(X), stats[m] =
 * (mean, rstd).  Backward: dX = add + dLN(G) (add may be NULL), dgamma/dbeta += . */
MPHSIR_API int mphsir_layernorm_fwd(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy,
                                    float* stats, long long M, int C, void* stream);
MPHSIR_API int mphsir_layernorm_bwd(const float* X, int ldx, const float* stats, const float* gamma, const float* G, int ldg,
                                    const float* add, int lda, float* dX, int lddx, float* dgamma, float* dbeta, long long M,
                                    int C, void* stream);

/* GatedMlp gate backward (:77-79) on the packed fc1 output H[:, (2j,2j+1)] = (value_j, gate_j): in place,
 * H <- dH, dHid <- hidden = value*gelu(gate) (the fc2 weight-gradient operand). */
MPHSIR_API int mphsir_glu_bwd(float* H, int ldh, float* dHid, int ldd, long long M, int hid_pad, void* stream);
/* GDFN gate (:388-389,:262-263) un-fused from the depthwise conv for training: T = [a | b] halves of hid_pad columns,
 * Y = gelu(a)*b; backward dT = [dY*b*gelu'(a) | dY*gelu(a)] (dT may alias T). */
MPHSIR_API int mphsir_gdfn_gate_fwd(const float* T, int ldt, float* Y, int ldy, long long M, int hid_pad, void* stream);
MPHSIR_API int mphsir_gdfn_gate_bwd(const float* T, int ldt, const float* dY, int ldy, float* dT, int lddt, long long M,
                                    int hid_pad, void* stream);

/* Y = alpha * row_scale[m / rows_per_batch] * X[m (mod x_row_mod)] + beta * Y  (DropPath :718-719, gradient sums) */
MPHSIR_API int mphsir_axpby(const float* X, int ldx, float* Y, int ldy, long long M, int C, float alpha, float beta,
                            const float* row_scale, int rows_per_batch, int x_row_mod, void* stream);
/* Y[n,:] = sum_b X[b*rows + n, :] */
MPHSIR_API int mphsir_batch_sum(const float* X, int ldx, float* Y, int ldy, int B, long long rows, int C, void* stream);

/* Window attention core backward (Spatial_Attention.forward :195-215 + roll/partition :671-696).  dqkv [B*H*W, >=3C]
 * image order like qkv; dbias_partial [groups, heads, 64, 64] (groups from mphsir_window_attn_bwd_groups) is reduced
 * with mphsir_colsum and scattered to the [225, heads] table gradient by mphsir_rpb_table_bwd (+=). */
MPHSIR_API int mphsir_window_attn_bwd_groups(int B, int H, int W, int heads);
MPHSIR_API int mphsir_window_attn_bwd(const float* qkv, int ldqkv, const float* bias, const float* dO, int ldo, float* dqkv,
                                      int lddq, float* dbias_partial, int groups, int B, int H, int W, int C, int heads,
                                      int shift, int precision /* MPHSIR_PREC_*: FFMA, or mma.sync with bf16x3 / bf16 operands */,
                                      void* stream);
MPHSIR_API int mphsir_rpb_table_bwd(const float* dbias /* [heads,64,64] */, float* dtable /* [225,heads] */, int heads,
                                    void* stream);

/* Local spectral branch backward (PG_Spectral_Attention.forward :135-153), windows = shifted 8x8 windows:
 *   window_reduce  : out[win,c] = scale * sum_{t in win} A[t,c] * (Bm ? Bm[t,c] : 1)     (window mean; dgate = sum dU*sa)
 *   local_gate_bwd : per-window chain backward -> record rows (layout in local_gate_bwd.cu); weights in the
 *                    reference's own [out,in] layouts
 *   gate_apply_bwd : dSA = dU * gate[win] + dMean[win]/64 */
MPHSIR_API int mphsir_window_reduce(const float* A, int lda, const float* Bm, int ldb, float* out, int B, int H, int W, int C,
                                    int shift, float scale, void* stream);
typedef struct {
  const float *param /* [128,r] */, *q /* [r,r] */, *kv /* [2r,r] */, *proj /* [r,r] */, *proj_bias /* [r] */, *up /* [C,r] */;
} mphsir_local_gate_bwd_weights;
MPHSIR_API int mphsir_local_gate_bwd_record_ld(int r);
MPHSIR_API int mphsir_local_gate_bwd(const float* LL /* [B_, >=128+r]: W_prompt m | W_down m */, int ldl, const float* dG,
                                     const mphsir_local_gate_bwd_weights* w, float* record, int ldr, int B_, int C, int r,
                                     void* stream);
/* forward twin used by the trainer: U = X + row_scale_b * SA * gate[win] (shortcut + locally gated spatial branch, :153,:715-718) */
MPHSIR_API int mphsir_gate_apply_fwd(const float* X, int ldx, const float* SA, int lds, const float* gate, const float* row_scale,
                                     float* U, int ldu, int B, int H, int W, int C, int shift, void* stream);
MPHSIR_API int mphsir_gate_apply_bwd(const float* dU, int ldu, const float* gate, const float* dMean, float* dSA, int lds,
                                     int B, int H, int W, int C, int shift, void* stream);

/* Global spectral attention backward, small-matrix part (:101-113): from P_b = dU_b^T v_b [B,C,C] (mphsir_wgrad per-sample),
 * the reduced Gram statistics gsum [B*heads, c*c+2c] of the forward and Wout [C,C] ([out,in]) produce dWout (+=),
 * dTemperature (+=) and Wb [B, 2C, ldwb] "in x out" with [dq | dk] = [q | k] Wb. */
MPHSIR_API int mphsir_spectral_bwd(const float* P, long long p_batch_stride, const float* Wout, const float* gsum,
                                   const float* temperature, float* Wb, int ldwb, long long wb_batch_stride, float* dWout,
                                   float* dTemperature, int B, int heads, int c, void* stream);

/* depthwise 3x3 weight gradient in the reference layout [C,1,3,3]: dW[map(c)*9 + tap] +=
 * (the data gradient is mphsir_dwconv3x3_fwd with flipped taps) */
MPHSIR_API int mphsir_dwconv3x3_wgrad(const float* X, int ldx, const float* dY, int ldy, float* dW, int B, int H, int W, int C,
                                      int map_mode, int map_a, int map_b, void* stream);

/* PixelUnshuffle(2)/PixelShuffle(2) (:437,:447) on token-major data, reference channel order c*4+2i+j.
 * unshuffle: in [B*H*W, C] -> out [B*H/2*W/2, 4C];  shuffle: in [B*H*W, 4C] -> out [B*2H*2W, C] */
MPHSIR_API int mphsir_pixel_unshuffle(const float* in, int ldi, float* out, int ldo, int B, int H, int W, int C, void* stream);
MPHSIR_API int mphsir_pixel_shuffle(const float* in, int ldi, float* out, int ldo, int B, int H, int W, int C, void* stream);
MPHSIR_API int mphsir_tokens_to_nchw(const float* in, int ld, float* out, int B, int C, int HW, void* stream);

/* bilinear resize backward (dX pre-zeroed, +=) and TVSP query backward (dLearnable [T,D] +=) */
MPHSIR_API int mphsir_bilinear_bwd(const float* dY, int ldy, float* dX, int ldx, int B, int h, int w, int H, int W, int C,
                                   void* stream);
MPHSIR_API int mphsir_tvsp_query_bwd(const float* dQ, int ldq, const float* clip_b, const float* weights, float* dLearnable,
                                     int B, int T, int D, int ps, void* stream);

/* loss[0] += mean|clamp(out,0,1) - clean| (train.py:58-61), dOut = d loss / d out * grad_scale */
MPHSIR_API int mphsir_l1_clamp_loss(const float* out, const float* clean, float* dOut, float* loss, long long numel,
                                    float grad_scale, void* stream);
/* torch.optim.AdamW step (train.py:69) over one flat fp32 range; g is multiplied by grad_scale first.  dyn (device, may
 * be NULL) = {lr, 1-beta1^step, 1-beta2^step} overrides lr/step so that a captured CUDA graph can be replayed every step. */
MPHSIR_API int mphsir_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, float grad_scale, const float* dyn, void* stream);

/* ---------------------------------------------------------------------------------------
 * Around the hot path: evaluation metrics and training-time degradations on the device.
 *
 * mphsir_psnr_ssim: utils/val_utils.py:49-69 (compute_psnr_ssim) per NCHW plane, on clip(.,0,1), data_range 1.
 *   sums[2*p]   = sum over the plane of (x-y)^2                         -> PSNR_p = 10 log10(H*W / sums[2p])
 *   sums[2*p+1] = sum of the SSIM index over the (H-6)*(W-6) whole 7x7 windows (skimage defaults: uniform window,
 *                 sample covariance, K1 0.01, K2 0.03; border crop 3)    -> SSIM_p = sums[2p+1] / ((H-6)(W-6))
 *   fp64 accumulation; sums [2*planes] is zeroed by the call.
 * mphsir_plane_nonzero: nonzero[p] = 1 if plane p has any non-zero element (compute_psnr_ssim2, :88: band-miss eval).
 * mphsir_degrade: out = clean * keep[b,c] * (u > mask_ratio[b]) + sigma[b,c] * n, u ~ U[0,1), n ~ N(0,1) drawn from
 *   Philox4x32-10 keyed by `seed`, counter = element pair index (utils/degradation_utils.py:25-39, :227-233, :275-284;
 *   one degradation per sample as utils/dataset_utils.py:128-146 — the per-sample / per-band parameters are drawn by
 *   the host and passed as the three small arrays).
 * ------------------------------------------------------------------------------------- */
MPHSIR_API int mphsir_psnr_ssim(const float* restored, const float* clean, int planes, int H, int W, double* sums, void* stream);
MPHSIR_API int mphsir_plane_nonzero(const float* X, int planes, long long hw, int* nonzero, void* stream);
MPHSIR_API int mphsir_degrade(const float* clean, float* out, int B, int C, long long hw, const float* sigma /* [B*C] */,
                              const float* keep /* [B*C] */, const float* mask_ratio /* [B] */, unsigned long long seed,
                              void* stream);
/* The structured half of 'complexN' (utils/degradation_utils.py:296-316), in place after mphsir_degrade, for the samples with
 * active[b] != 0:  x = x * colmul[b,c,col] + coladd[b,c,col]  (deadline columns: colmul 0, :57-68; stripes: coladd = -offset, :41-55),
 * then with probability impulse[b,c] a pixel becomes 1 (salt, half of the flips) or 0 (pepper) (:70-84); the per-pixel decisions
 * come from the Philox4x32-10 stream keyed by `seed` with counter word 2 = 1.  colmul / coladd [B*C*W], impulse [B*C], active [B]. */
MPHSIR_API int mphsir_degrade_structured(float* x, int B, int C, int H, int W, const float* colmul, const float* coladd,
                                         const float* impulse, const int* active, unsigned long long seed, void* stream);
/* Gaussian blur degradation (utils/degradation_utils.py:91-108): every band of sample b with ksize[b] > 0 (odd, <= 21; device
 * int array [B]) is convolved with the k x k outer product of the normalised 1-D Gaussian of sigma 0.3((k-1)/2 - 1) + 0.8,
 * zero padding k/2; planes of samples with ksize[b] == 0 are not touched.  kmax >= every ksize[b] (validated on the host). */
MPHSIR_API int mphsir_gaussian_blur(const float* in, float* out, const int* ksize, int B, int C, int H, int W, int kmax, void* stream);
/* Super-resolution degradation 'sr' (utils/degradation_utils.py:165-176 `_bicubic_downsample`, then :189-200 `_resize` through
 * single_degrade :431-432): every band of sample b with factor[b] > 0 (device int array [B]; the reference draws 2, 4 or 8) is
 * bicubically down-sampled to (H / f, W / f) with torch's align_corners=True convention (A = -0.75, clamped indices) and
 * replicated f x f back to H x W; planes of samples with factor[b] == 0 are not touched.  H and W need not be multiples of f
 * (trailing pixels repeat the last low-resolution row / column); f <= min(H, W) is checked on the host. */
MPHSIR_API int mphsir_sr_degrade(const float* in, float* out, const int* factor, int B, int C, int H, int W, void* stream);
/* Generic k x k blur degradations — circle blur (utils/degradation_utils.py:110-128), square blur (:150-163), motion blur
 * (:130-148): F.conv2d(x, kernel.repeat(C,1,1,1), padding=k//2, groups=C) with one host-built kernel.  taps: its k*k values,
 * row-major, DEVICE memory; k odd, <= 21; zero padding; planes of samples with active[b] == 0 (device int [B]) are not touched. */
MPHSIR_API int mphsir_blur2d(const float* in, float* out, const float* taps, const int* active, int B, int C, int H, int W, int k,
                             void* stream);
/* Poisson noise 'poissonN' (utils/degradation_utils.py:86-89): out = Poisson(max(x, 0) * scale[b]) / scale[b], one Philox4x32-10
 * uniform per element (key `seed`, counter word 2 = 2) inverted through the Poisson CDF in double precision; samples with
 * scale[b] <= 0 (device float [B]) are not touched.  chw = elements per sample.  in == out is allowed. */
MPHSIR_API int mphsir_poisson(const float* in, float* out, const float* scale, int B, long long chw, unsigned long long seed,
                              void* stream);
/* Haze degradation (utils/degradation_utils.py:235-273) for a caller-supplied cirrus-band map (the reference loads it from .mat
 * files and resizes it to the patch).  mphsir_topk_mean: mean[p] = mean of the k largest elements of plane p (the atmospheric
 * light of :257-261 with k = max(int(H*W*top_percent/100), 1); ties counted as a sort counts them).  mphsir_haze:
 * out = x * T + light[b,c] * (1 - T), T = exp(expo[c] * log(t1)), t1 = 1 - omega[b] * cirrus[b,pixel] (<= 0 -> 1e-10), evaluated
 * in double; expo[c] = (lambda_0 / lambda_c)^gamma from the host; samples with omega[b] <= 0 are not touched.  in == out allowed. */
MPHSIR_API int mphsir_topk_mean(const float* X, int planes, long long hw, int k, float* mean, void* stream);
MPHSIR_API int mphsir_haze(const float* in, float* out, const float* cirrus /* [B*hw] */, const float* omega /* [B] */,
                           const float* expo /* [C] */, const float* light /* [B*C] */, int B, int C, long long hw, void* stream);

/* ---------------------------------------------------------------------------------------
 * Collectives of the row-sharded scene over NVLink peer memory (mp_hsir_b200/csrc/peer.cu; one process per GPU of one
 * node).  Every rank allocates one WINDOW of mphsir_peer_window_bytes(halo_cap, ar_cap) bytes (256-byte aligned, zeroed
 * once), exports it through CUDA IPC and maps every peer's window; `windows[q]` is rank q's window as mapped in THIS
 * process.  All ranks must issue the same sequence of these calls (sequence numbers live in the windows, so CUDA-graph
 * replays stay in step).  Each call is ONE kernel: push to the peers' windows, publish flags, wait for the peers' flags,
 * consume.  halo_exchange: halo_top <- previous rank's own_last rows, halo_bottom <- next rank's own_first rows (cyclic),
 * n floats per direction (multiple of 4, <= halo_cap).  all_reduce: data[0..n) <- sum over ranks in rank order (the
 * same bits on every rank), n <= ar_cap.
 * ------------------------------------------------------------------------------------- */
MPHSIR_API size_t mphsir_peer_window_bytes(long long halo_cap, long long ar_cap);
/* window life cycle (the one place the library owns device memory: an IPC handle names a whole cudaMalloc allocation):
 * alloc + zero on the current device, export a 64-byte CUDA IPC handle, open a peer's handle (lazy P2P enable), close, free */
MPHSIR_API int mphsir_peer_window_alloc(size_t bytes, void** window);
MPHSIR_API int mphsir_peer_window_free(void* window);
MPHSIR_API int mphsir_peer_export(void* window, unsigned char* handle64);
MPHSIR_API int mphsir_peer_open(const unsigned char* handle64, void** mapped);
MPHSIR_API int mphsir_peer_close(void* mapped);
MPHSIR_API int mphsir_peer_halo_exchange(void* const* windows, int rank, int world, long long halo_cap, long long ar_cap,
                                         const float* own_first, const float* own_last, float* halo_top, float* halo_bottom,
                                         long long n, void* stream);
MPHSIR_API int mphsir_peer_all_reduce(void* const* windows, int rank, int world, long long halo_cap, long long ar_cap,
                                      float* data, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPHSIR_H_ */
