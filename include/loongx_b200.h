/*
 * loongx_b200 — C ABI of the B200-native LoongX denoising hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*); no torch types cross
 * this boundary.  All functions return 0 on success and a negative lx_status on error (message via
 * lx_last_error()).  Nothing here synchronises the stream or allocates device memory: the caller owns all
 * buffers (in the Python host layer they come from the torch caching allocator).
 *
 * The reference (LanceZPF/loongx) is pure Python with no FFI layer; each function below names
 * the reference function / call site whose arithmetic it replaces.
 */
#ifndef LOONGX_B200_H_
#define LOONGX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum lx_status { LX_OK = 0, LX_ERR_ARG = -1, LX_ERR_CUDA = -2, LX_ERR_UNSUPPORTED = -3 };

const char* lx_last_error(void);
int lx_version(void);
/* Device properties the host layer needs for grid sizing: out[0]=SM count, out[1]=cc major, out[2]=cc minor. */
int lx_device_info(int32_t* out3);

/* Optional exchange workspace for the split-work schedules of lx_attention and lx_gemm_bf16 (a tile whose work is cut
 * between two CTAs is finished by one of them from the other's fp32 partial).  `ptr`: 1024-byte-aligned device buffer of
 * `bytes` (64 MiB covers every shape of the DiT path), owned by the caller and bound to `stream`: only launches on that
 * stream use it, everything else (and every launch when no workspace is registered) runs the unsplit schedule with
 * identical results up to fp32 summation order.  The call enqueues a small memset on `stream`; ptr = NULL unregisters. */
int lx_set_workspace(void* ptr, int64_t bytes, void* stream);

/* Launch accounting (bench.py's gpu_launches / roofline): kernel classes 0 = tcgen05 GEMM, 1 = attention, 2 = DiT row
 * kernels, 3 = CS3/DGF kernels; cls < 0 = all.  lx_profile_begin() switches on CUDA-event timing of every launch (on
 * the stream it is launched on); lx_profile_end() synchronises and returns per-class totals in arrays of 4:
 * milliseconds, launches, algorithmic work (FLOPs for classes 0-1, bytes for classes 2-3). */
int64_t lx_launch_count(int32_t cls);
void lx_launch_count_reset(void);
int lx_profile_begin(void);
int lx_profile_end(double* ms, int64_t* launches, double* work);

/* ------------------------------------------------------------------------------------------------------
 * Row-tile metadata.  Activations are stored stream-major: rows = [txt(B*Nt) | img(B*Ni) | cond(B*Nc)],
 * every stream length a multiple of 128, so each 128-row tile belongs to one (stream, batch element).
 * ------------------------------------------------------------------------------------------------------ */
typedef struct lx_tile_meta {
  int32_t stream;  /* 0 = txt, 1 = img, 2 = cond */
  int32_t batch;   /* batch element */
  int32_t seq_row; /* batch*S + offset of the tile's first token inside the joint [txt|img|cond] sequence */
  int32_t reserved;
} lx_tile_meta_t;

/* ------------------------------------------------------------------------------------------------------
 * tcgen05 GEMM   C = epilogue(A[M,K] · W[N,K]^T)         (bf16 in, fp32 accumulate in TMEM)
 * Replaces every nn.Linear on the DiT path: attn.to_{q,k,v}/add_*_proj (block.py:27-29,46-48,81-83),
 * to_out/to_add_out (block.py:154-160), ff/ff_context (block.py:258-265), proj_mlp/proj_out
 * (block.py:302,328), x_embedder/context_embedder/time_text_embed/norm*.linear/proj_out
 * (transformer.py:92-115,243-244).
 * ------------------------------------------------------------------------------------------------------ */
enum lx_epilogue {
  LX_EPI_BIAS = 0,          /* out = acc + bias                                   -> bf16                */
  LX_EPI_BIAS_GELU = 1,     /* out = gelu_tanh(acc + bias)                        -> bf16                */
  LX_EPI_BIAS_SILU = 2,     /* out = silu(acc + bias)                             -> bf16                */
  LX_EPI_GATE_RESIDUAL = 3, /* out = residual + gate[stream,batch] * (acc + bias) -> bf16 (in place ok)  */
  LX_EPI_QKV = 4,           /* bias, per-head RMSNorm(q,k), RoPE, scatter to Q/K/V [B,H,S,128] -> bf16    */
  LX_EPI_BIAS_F32 = 5,      /* out = acc + bias                                   -> fp32                */
  /* training step (model.py:569-729): the GELU's neighbours fused into the GEMMs either side of it */
  LX_EPI_BIAS_GELU_DUAL = 6, /* out = gelu_tanh'(acc + bias) (the backward's factor) AND out2 = gelu_tanh(acc + bias) -> bf16 */
  LX_EPI_MUL_AUX = 7         /* out = (acc + bias) * residual[row, col]  -> bf16 (residual = BIAS_GELU_DUAL's `out`)     */
};

typedef struct lx_gemm_segment {
  int32_t mode;       /* lx_epilogue */
  int32_t col_offset; /* output column = (n - first column of the segment) + col_offset */
  void* out;          /* [M, ldo] */
  int64_t ldo;        /* elements */
} lx_gemm_segment_t;

/* Row group: M-tiles [m_begin/128, next group's m_begin/128) multiply by this weight panel.  Lets one launch run the
 * text rows against the *_context weights, the image rows against W and the condition rows against the LoRA-merged
 * W + (alpha/r) B A (lora_controller.py:5-43 applies LoRA to the condition branch only). */
typedef struct lx_gemm_group {
  const void* W; /* bf16 [N, K], row stride ldw */
  int64_t ldw;
  const float* bias; /* fp32 [N] or NULL */
  int32_t K;         /* multiple of 8 */
  int32_t m_begin;   /* first row of the group (multiple of 128; group 0 starts at 0) */
} lx_gemm_group_t;

typedef struct lx_gemm_desc {
  const void* A; /* bf16 [M, K], row stride lda (elements, multiple of 8) */
  int64_t lda;
  int32_t M, N;
  int32_t n_groups; /* 1..3 */
  int32_t n_split;  /* output columns >= n_split use seg[1] (multiple of 256); = N for a single segment */
  lx_gemm_group_t group[3];
  lx_gemm_segment_t seg[2];
  const lx_tile_meta_t* tile_meta; /* [ceil(M/128)]; required by GATE_RESIDUAL and QKV */
  /* LX_EPI_GATE_RESIDUAL / LX_EPI_MUL_AUX */
  const void* residual; /* bf16 [M, ldr]; column = (n - first column of the segment) + col_offset */
  int64_t ldr;
  const void* gate[3];    /* bf16 gate vectors per stream: gate[s] + batch*gate_stride[s] + n */
  int64_t gate_stride[3]; /* elements */
  /* LX_EPI_QKV: columns [0,D)=q, [D,2D)=k, [2D,3D)=v with D = heads*128 */
  void* q; /* bf16 [B, heads, seq_total, 128] */
  void* k;
  void* v;
  int32_t heads, seq_total;
  const float* rms_q[3]; /* fp32 [128] RMSNorm weight per stream (NULL = no norm) */
  const float* rms_k[3];
  const float* rope; /* fp32 [seq_total, 64, 2] (cos, sin) per rotary pair, NULL = no RoPE */
  float rms_eps;
  int32_t tile_n; /* 0 = choose the N tile (256 / 224 / 192) that minimises wave quantisation; else force it (128 is
                     only available forced: for outputs of <= 128 columns) */
  /* != 0: some W panel is an activation written by an earlier kernel on this stream (e.g. K / V^T of the VAE's
   * mid-block attention), not a constant weight: the kernel then requests no W tile ahead of its programmatic-dependency
   * wait.  0 (weights): the first W tiles are prefetched while the previous kernel is still draining. */
  int32_t w_dynamic;
  /* Second output (at most one segment uses it), column = (n - first column of the segment) + col_offset2:
   * LX_EPI_BIAS_GELU_DUAL: gelu_tanh(acc + bias); LX_EPI_GATE_RESIDUAL (optional, NULL = none): acc + bias, the pre-gate
   * projection the training backward needs for the gate gradient */
  int32_t col_offset2;
  void* out2; /* bf16 [M, ldo2] */
  int64_t ldo2;
  /* LX_EPI_QKV, optional (NULL = none): the projection before RMSNorm / RoPE (acc + bias) is also written here, row-major
   * [M, ld_qkv_pre] with q | k | v in columns [0, 3*heads*128) -- what the training backward differentiates through */
  void* qkv_pre;
  int64_t ld_qkv_pre;
} lx_gemm_desc_t;

int lx_gemm_bf16(const lx_gemm_desc_t* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Joint attention  out = softmax(Q K^T * scale [+ mask / bias]) V   over the [txt | img | cond] sequence.
 * Replaces F.scaled_dot_product_attention and the surrounding concat / transpose / split
 * (block.py:70-72, 102-135) plus the block masks (block.py:106-120) and the c_factor bias (block.py:121-128).
 * ------------------------------------------------------------------------------------------------------ */
typedef struct lx_attn_desc {
  const void* q; /* bf16 [B, H, S, 128] (written by the LX_EPI_QKV epilogue) */
  const void* k;
  const void* v;
  void* out; /* bf16 rows in the stream-major layout; head h occupies columns col_offset + [128h, 128h+128) */
  int64_t ldo;
  const int32_t* out_row_base; /* [B * S/128]: first output row of (batch, query tile) */
  int32_t col_offset;
  int32_t B, H, S;
  int32_t n_cond;    /* trailing condition tokens of the sequence (multiple of 128, 0 = none) */
  int32_t mask_mode; /* 0 = full; 1 = cond<->rest blocked (union_cond_attn=False); 2 = cond queries see only cond keys
                        (independent_condition) */
  float cross_bias;  /* log(c_factor) added to cond<->rest logits; non-zero overrides mask_mode like the reference */
  float scale;       /* 1/sqrt(128) */
  float* lse;        /* optional fp32 [B, H, S]: log2-domain log-sum-exp of every query row (for lx_attention_bwd) */
  /* Ragged streams: each of [txt | img | cond] is padded to a multiple of 128 tokens; stream_end[s] is the (padded) end
   * of stream s inside the sequence and pad[s] in [0,128) the number of padding tokens at its end.  Padding KEYS are
   * masked out; padding query rows produce don't-care output rows.  All zeros = no padding. */
  int32_t stream_end[3];
  int32_t pad[3];
  int32_t q_tiles;  /* number of leading 128-row query tiles to compute (0 = all S/128); keys always span the whole S */
  int32_t reserved;
} lx_attn_desc_t;

int lx_attention(const lx_attn_desc_t* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound).
 * ------------------------------------------------------------------------------------------------------ */
/* out = LayerNorm(x; no affine, eps) * (1 + scale[stream,batch]) + shift[stream,batch].
 * Replaces AdaLayerNormZero / -Single / -Continuous + norm2 FiLM (block.py:192-207, 238-253, 301-305;
 * transformer.py:243). */
typedef struct lx_lnmod_desc {
  const void* x; /* bf16 [rows, ldx] */
  int64_t ldx;
  void* out; /* bf16 [rows, ldo] */
  int64_t ldo;
  int32_t rows, D;
  const lx_tile_meta_t* tile_meta;
  const void* shift[3]; /* bf16, per stream: shift[s] + batch*stride[s] */
  const void* scale[3];
  int64_t stride[3];
  float eps;
  int32_t reserved;
} lx_lnmod_desc_t;
int lx_ln_modulate(const lx_lnmod_desc_t* desc, void* stream);

/* Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): out[m] = [cos(t*mult*f), sin(t*mult*f)] bf16. */
int lx_timestep_embed(const float* t, void* out, int64_t ldo, int32_t M, float mult, void* stream);
/* out[m, 0:D] = silu(a[m] + b[m] + c[m % c_rows]) for bf16 [M, D] inputs (b, c may be NULL): temb = t_emb + g_emb +
 * text_emb followed by the SiLU every AdaLN applies (transformer.py:102-114, App. A.2/A.5). */
int lx_add_silu_bcast(const void* a, const void* b, const void* c, int32_t c_rows, void* out, int64_t ldo, int32_t M,
                      int32_t D, void* stream);
/* x[m, 0:D] = bf16(float(x[m]) + float(r[m])), bf16 rows with strides ldx / ldr: `hidden_states + controlnet_block_samples[i]`
 * (transformer.py:172-181) and its single-block form on the image rows (transformer.py:230-239). */
int lx_add_rows(void* x, int64_t ldx, const void* r, int64_t ldr, int32_t rows, int32_t D, void* stream);
/* FlowMatchEulerDiscreteScheduler.step (generate.py:349): out = bf16(float(x) + dt*float(v)), n elements. */
int lx_euler_step(const void* x, const void* v, void* out, float dt, int64_t n, void* stream);
/* FluxPosEmbed (transformer.py:130-134): ids fp32 [S,3] -> table fp32 [S,64,2] (cos, sin), float64 internally. */
int lx_rope_table(const float* ids, float* table, int32_t S, int32_t d0, int32_t d1, int32_t d2, double theta,
                  void* stream);
/* FluxPipeline._pack_latents / _unpack_latents (generate.py:262, 375): [B,C,h,w] <-> [B,(h/2)(w/2),4C], bit-exact. */
int lx_pack_latents(const void* in, void* out, int32_t B, int32_t C, int32_t h, int32_t w, int32_t elem_bytes,
                    int32_t unpack, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * DiT engine: the whole tranformer_forward (transformer.py:47-252) as one native call sequence.
 * Weight layout: every nn.Linear is one lx_linear_t.  A LoRA-targeted Linear additionally carries w_lora =
 * W + (alpha/r) lora_B lora_A, merged once at load (weights are frozen at inference); rows on which LoRA is active
 * (the condition stream by default, lora_controller.py:5-43; also the image / text rows with latent_lora) are a
 * separate row group of the same GEMM launch that reads w_lora instead of w.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct lx_linear {
  const void* w; /* bf16 [n, ldw] */
  int64_t ldw;
  const float* bias;  /* fp32 [n] or NULL */
  const void* w_lora; /* bf16 [n, ldw] merged weight, or NULL when the layer is not a LoRA target */
  int32_t n, k;
} lx_linear_t;

typedef struct lx_double_block {
  lx_linear_t qkv;     /* [to_q;to_k;to_v]  (img + cond rows) */
  lx_linear_t qkv_ctx; /* [add_q_proj;add_k_proj;add_v_proj] (txt rows) */
  lx_linear_t out, out_ctx;
  lx_linear_t ff_up, ff_down, ff_ctx_up, ff_ctx_down;
  const float* norm_q; /* fp32 [128] */
  const float* norm_k;
  const float* norm_added_q;
  const float* norm_added_k;
} lx_double_block_t;

typedef struct lx_single_block {
  lx_linear_t qkv_mlp; /* [to_q;to_k;to_v;proj_mlp] */
  lx_linear_t proj_out;
  const float* norm_q;
  const float* norm_k;
} lx_single_block_t;

typedef struct lx_dit_model {
  int32_t num_layers, num_single_layers, heads, in_channels;
  int32_t joint_dim, pooled_dim, guidance_embeds, reserved;
  int32_t axes_dim[3];
  int32_t reserved2;
  lx_linear_t x_embedder, context_embedder;
  lx_linear_t time_1, time_2, guid_1, guid_2, text_1, text_2;
  lx_linear_t mod_img;    /* all double blocks' norm1.linear stacked:         [L*6D, D] */
  lx_linear_t mod_txt;    /* all double blocks' norm1_context.linear stacked: [L*6D, D]     */
  lx_linear_t mod_single; /* all single blocks' norm.linear stacked:          [Ls*3D, D] */
  lx_linear_t norm_out, proj_out;
  const lx_double_block_t* double_blocks; /* host array [num_layers] */
  const lx_single_block_t* single_blocks; /* host array [num_single_layers] */
} lx_dit_model_t;

/* Geometry + caller-owned device buffers of one batch of edits (R = B*(n_txt+n_img+n_cond) rows, D = heads*128,
 * T = number of denoise steps prepared). */
typedef struct lx_dit_plan {
  int32_t B, n_txt, n_img, n_cond;
  int32_t T;
  int32_t mask_mode;    /* see lx_attn_desc_t */
  int32_t latent_lora;  /* model_config["latent_lora"]: LoRA also on the image (and, in single blocks, text) rows */
  int32_t add_cond_attn;
  float cross_bias;
  int32_t reserved;
  const lx_tile_meta_t* tile_meta; /* device [R/128] */
  const int32_t* out_row_base;     /* device [B*S/128] */
  const float* rope;               /* device [S,64,2] for the joint [txt|img|cond] ids, or NULL */
  void* X;       /* bf16 [R, D] residual stream (stream-major rows) */
  void* XN;      /* bf16 [R, D] modulated operand */
  void* Q;       /* bf16 [B,H,S,128] */
  void* K;
  void* V;
  void* scratch; /* bf16 [R, 5D]: attention out / FF hidden / single-block concat */
  void* X0_txt;  /* bf16 [B*n_txt, D]  context_embedder(prompt_embeds)   (step invariant) */
  void* X0_cond; /* bf16 [B*n_cond, D] x_embedder(cond_latents)          (step invariant) */
  void* emb_tmp; /* bf16 [4, T*B + B, D] scratch of the timestep / guidance / text MLPs */
  void* sin_tmp; /* bf16 [T*B + B, 256] */
  void* silu_t;  /* bf16 [T*B, D] silu(temb) per (step, batch) */
  void* silu_c;  /* bf16 [B,   D] silu(cond_temb) */
  void* mod_img;      /* bf16 [T*B, L*6D] */
  void* mod_txt;      /* bf16 [T*B, L*6D] */
  void* mod_single;   /* bf16 [T*B, Ls*3D] */
  void* mod_out;      /* bf16 [T*B, 2D] */
  void* mod_cond_img;    /* bf16 [B, L*6D]  (cond stream, step invariant) */
  void* mod_cond_single; /* bf16 [B, Ls*3D] */
  float* t_dev;          /* fp32 [T*B + B] device scratch for timestep values */
  float* g_dev;          /* fp32 [T*B + B] device scratch for guidance values */
  int32_t pad[3];        /* padding tokens at the end of the txt / img / cond stream (n_* above are the PADDED lengths,
                            multiples of 128; the reference accepts any H, W divisible by 16) */
  /* Step-invariant condition branch: under model_config.independent_condition the condition queries see
   * only condition keys and the condition stream is modulated by cond_temb (c_t, constant), so its K / V of every block do
   * not depend on the denoise step.  With cond_cached != 0 the block calls process the text + image rows only and read
   * the condition keys / values that a full pass (cond_cached = 0, same kv_block_stride) left in each block's own K / V
   * buffer: block b uses K + b*kv_block_stride (double blocks first, then single blocks). */
  int32_t cond_cached;
  int64_t kv_block_stride; /* elements between consecutive blocks' K (and V) buffers; 0 = one shared buffer */
} lx_dit_plan_t;

/* Step-invariant work, once per edit (generate.py:168-306 hoisted): context_embedder, x_embedder(cond), temb for all
 * T timesteps + cond_temb (c_t), every block's AdaLN modulation vector.  timesteps / guidance are HOST arrays of
 * length T*B and B (timestep in (0,1], the *1000 of transformer.py:95-98 is applied inside). */
int lx_dit_prepare(const lx_dit_model_t* model, const lx_dit_plan_t* plan, const void* prompt_embeds,
                   const void* pooled, const void* cond_latents, const float* timesteps, const float* guidance,
                   float c_t, void* stream);
/* x_embedder(latents) + copies of the step-invariant txt / cond embeddings into plan->X (transformer.py:91-93,115). */
int lx_dit_embed(const lx_dit_model_t* model, const lx_dit_plan_t* plan, const void* latents, void* stream);
/* One DiT forward at prepared step `step`: latents bf16 [B, n_img, in_channels] -> noise_pred (same shape). */
int lx_dit_step(const lx_dit_model_t* model, const lx_dit_plan_t* plan, int32_t step, const void* latents,
                void* noise_pred, void* stream);
/* The tail of lx_dit_step on its own: AdaLayerNormContinuous (norm_out) on the image rows of plan->X + proj_out
 * (transformer.py:241-244).  lx_dit_embed + the block calls below + lx_dit_head = lx_dit_step; the split form is what a
 * caller with controlnet residuals between the blocks (transformer.py:172-181, 230-239; lx_add_rows) uses. */
int lx_dit_head(const lx_dit_model_t* model, const lx_dit_plan_t* plan, int32_t step, void* noise_pred, void* stream);
/* Reference-granularity entry points for parity tests (block.py:179-278 / 281-339): run ONE block of the prepared
 * plan in place on plan->X. */
int lx_dit_double_block(const lx_dit_model_t* model, const lx_dit_plan_t* plan, int32_t step, int32_t block,
                        void* stream);
int lx_dit_single_block(const lx_dit_model_t* model, const lx_dit_plan_t* plan, int32_t step, int32_t block,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------
 * CS3 (cross-scale state-space signal encoders) and DGF (DUAN dynamic gated fusion): fp32, once per edit.
 * Reference: src/train/model.py:16-373 (encoders), 479-511 (length normaliser), 731-779 (fuse_*), 947-1035 (DUAN);
 * the S4 layer itself is the third-party s4torch package (model.py:14), restated per SURVEY.md App. B.
 * ------------------------------------------------------------------------------------------------------ */
/* OminiModel.spatial_pyramid_pooling (model.py:479-511): zero-pad / truncate rows of length Lin to Lout. */
int lx_pad_truncate(const float* in, float* out, int32_t rows, int32_t Lin, int32_t Lout, void* stream);
/* S4 DPLR convolution kernel K[d, L] from (lambda, p, q)[n] complex64, (B, Ct)[d, n] complex64, log_step[d]:
 * Cauchy sums at the L roots of unity + inverse DFT, float64 internally.  workspace: L*d*16 bytes. */
int lx_s4_kernel_gen(const void* lam, const void* p, const void* q, const void* Bm, const void* Ct, const float* log_step,
                     float* K, void* workspace, int32_t d, int32_t n, int32_t L, void* stream);
/* y[b,c,l] = gelu(sum_{j<=l} K[c,j] u[b,c,l-j] + D[c] u[b,c,l]) for u [B, d, L] (S4Layer + GELU). */
int lx_s4_conv_gelu(const float* u, const float* K, const float* D, float* y, int32_t B, int32_t d, int32_t L,
                    void* stream);
/* out[b,j,l] = sum_c W[j,c] in[b,c,l] + bias[j] (+ residual[b,j,l]) (then LayerNorm over j with ln_w / ln_b). */
int lx_channel_linear(const float* in, const float* W, const float* bias, const float* residual, const float* ln_w,
                      const float* ln_b, float* out, int32_t B, int32_t d_in, int32_t d_out, int32_t L, float eps,
                      void* stream);
/* nn.AdaptiveAvgPool1d(O) on in [B, C, L]; out[b*out_bstride + c*cs + i*is + off] (model.py:83-103, 345-373). */
int lx_adaptive_pool(const float* in, float* out, int32_t B, int32_t C, int32_t L, int32_t O, int64_t out_bstride,
                     int32_t cs, int32_t is, int32_t off, void* stream);
/* Several nn.AdaptiveAvgPool1d sizes of the same input in ONE launch (FeaturePyramidPooling, model.py:345-373): bin i of
 * pool k -> out[b*out_bstride + c*cs + off[k] + i], n <= 8 (host arrays O, off). */
int lx_adaptive_pool_multi(const float* in, float* out, int32_t B, int32_t C, int32_t L, int32_t n, const int32_t* O,
                           const int32_t* off, int64_t out_bstride, int32_t cs, void* stream);
/* y[b, :] = W[n_out, n_in] x[b, :] + bias for B <= 8 rows (weights streamed once). */
int lx_gemv_f32(const float* W, const float* bias, const float* x, float* y, int32_t B, int32_t n_out, int32_t n_in,
                int64_t ldx, int64_t ldy, void* stream);
/* y = relu(LayerNorm(x) * w + b) per row of [rows, n]. */
int lx_ln_relu_rows(const float* x, const float* w, const float* b, float* y, int32_t rows, int32_t n, float eps,
                    void* stream);
/* Unflatten(tokens, 8) -> Linear(8, n_out): out[b, t, :] = W[n_out, 8] h[b, 8t:8t+8] + bias (model.py:70-71). */
int lx_token_linear(const float* h, const float* W, const float* bias, float* out, int32_t B, int32_t tokens,
                    int32_t n_out, int64_t out_bstride, void* stream);

/* Batched fp32 GEMM  C[b] = act(A[M,K] . Bm[b][K,N] + bias[m] (+ R[b]));  act 0 none / 1 relu / 2 sigmoid.  With
 * rowmean != NULL nothing is stored; mean_n(act(.)) is accumulated into rowmean[b, m] (must be zeroed). */
typedef struct lx_sgemm_desc {
  const float* A;
  int64_t lda;
  const float* Bm;
  int64_t ldb;
  int64_t b_bstride;
  const float* bias; /* [M] or NULL */
  const float* R;    /* residual with C's layout or NULL */
  float* C;
  int64_t ldc;
  int64_t c_bstride;
  float* rowmean; /* [batch, M] or NULL */
  int32_t M, N, K, batch;
  int32_t act;
  int32_t reserved;
  float* rowpart; /* with rowmean: scratch [batch, M, ceil(N / 64)] -> the mean is summed in a fixed order (bit-reproducible);
                     NULL = float atomics into a zeroed rowmean */
} lx_sgemm_desc_t;
int lx_sgemm_f32(const lx_sgemm_desc_t* desc, void* stream);

/* fp32 <-> bf16 element cast (to_bf16 != 0: fp32 -> bf16). */
int lx_cast(const void* in, void* out, int64_t n, int32_t to_bf16, void* stream);

/* DUAN (model.py:947-1035): gate / FiLM 1x1-conv weights in nn.Conv1d layout [out, in]. */
typedef struct lx_duan_weights {
  const float* gate_w1; /* [hidden, C] */
  const float* gate_b1;
  const float* gate_w2; /* [C, hidden] */
  const float* gate_b2;
  const float* mlp_w1; /* [hidden, C] */
  const float* mlp_b1;
  const float* mlp_w2; /* [2C, hidden] */
  const float* mlp_b2;
  int32_t hidden;
  float eps;
} lx_duan_weights_t;
/* y[b] (batch stride y_bstride) = DUAN(x, c) for fp32 x, c [B, C, L]; workspace: fp32
 * [B*(12*C + hidden*(L+1) + C*ceil(L/64))].  Bit-reproducible (no float atomics). */
int lx_duan_forward(const lx_duan_weights_t* w, const float* x, const float* c, float* y, int64_t y_bstride, int32_t B,
                    int32_t C, int32_t L, float keep_ratio, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Backward of CS3 / DGF (fp32): the gradients `loss.backward()` leaves on every encoder / fusion parameter of the
 * reference's OminiModel.step (src/train/model.py:656-701 runs inside the autograd graph; Lightning's DDP all-reduces
 * them together with the LoRA factors, train.py:181-183).  Weight / bias gradients are ACCUMULATED (+=).
 * ------------------------------------------------------------------------------------------------------ */
/* C[b] = alpha op(A[b]) op(B[b]) + beta C[b];  op(A) is [M, K] (stored [K, M] when trans_a), op(B) is [K, N] (stored
 * [N, K] when trans_b).  reduce_batch != 0: ONE C = alpha sum_b op(A[b]) op(B[b]) + beta C (weight gradients). */
typedef struct lx_sgemm_ex_desc {
  const float* A;
  int64_t lda, a_bstride;
  const float* Bm;
  int64_t ldb, b_bstride;
  float* C;
  int64_t ldc, c_bstride;
  int32_t M, N, K, batch;
  int32_t trans_a, trans_b, reduce_batch, reserved;
  float alpha, beta;
} lx_sgemm_ex_desc_t;
int lx_sgemm_ex(const lx_sgemm_ex_desc_t* desc, void* stream);
/* out[j] (+)= sum_r in[r*ld + j]  (bias gradient of a row-batched Linear). */
int lx_sum_rows_f32(const float* in, int64_t ld, int32_t rows, int32_t n, float* out, int32_t accumulate, void* stream);
/* out[c] (+)= sum_b sum_l a[b,c,l] (* b[b,c,l] when b != NULL) for [B, C, L] tensors (bias gradients of 1x1 convs /
 * channel Linears; the S4 skip gradient dD = sum ds * u). */
int lx_sum_last_f32(const float* a, const float* b, int32_t B, int32_t C, int32_t L, float* out, int32_t accumulate,
                    void* stream);
/* Backward of lx_ln_relu_rows: y = relu(LayerNorm(x) w + b) -> dx; dw, db += (model.py:60-72). */
int lx_ln_relu_rows_bwd(const float* x, const float* w, const float* b, const float* dy, float* dx, float* dw, float* db,
                        int32_t rows, int32_t n, float eps, void* stream);
/* nn.Dropout(p) in train mode (model.py:64,68): y[i] = keep(seed, i) ? x[i] / (1 - p) : 0 with a counter-based mask, so
 * the backward is the same call on the gradient with the same seed.  (The mask stream is not torch's Philox stream.) */
int lx_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, void* stream);
/* Backward of lx_token_linear: dh[b, 8t+k] = sum_o dout[b,t,o] W[o,k]; dW, dbias +=. */
int lx_token_linear_bwd(const float* h, const float* W, const float* dout, float* dh, float* dW, float* dbias, int32_t B,
                        int32_t tokens, int32_t n_out, int64_t out_bstride, void* stream);
/* Backward of lx_adaptive_pool (same index arguments): din[b,c,l] += dfeat[...] / bin length. */
int lx_adaptive_pool_bwd(const float* dfeat, float* din, int32_t B, int32_t C, int32_t L, int32_t O, int64_t f_bstride,
                         int32_t cs, int32_t is, int32_t off, void* stream);
/* LayerNorm over the d channels of v [B, d, L] (S4Block post-norm): dv from dy; dw, db +=. */
int lx_channel_ln_bwd(const float* v, const float* w, const float* dy, float* dv, float* dw, float* db, int32_t B, int32_t d,
                      int32_t L, float eps, void* stream);
/* ds = dg * gelu_erf'(s). */
int lx_gelu_erf_bwd(const float* s, const float* dg, float* ds, int64_t n, void* stream);
/* y = act(causal_conv(u, K) + D u) like lx_s4_conv_gelu, with: pre (optional) = the pre-activation; reverse != 0 =
 * u and y are indexed time-reversed (the adjoint of the causal convolution: du = conv_reversed(ds, K) + D ds);
 * gelu == 0 = no activation; D may be NULL. */
int lx_s4_conv(const float* u, const float* K, const float* D, float* y, float* pre, int32_t B, int32_t d, int32_t L,
               int32_t reverse, int32_t gelu, void* stream);
/* dK[c, j] (+)= sum_b sum_{l>=j} ds[b,c,l] u[b,c,l-j]. */
int lx_s4_conv_wgrad(const float* ds, const float* u, float* dK, int32_t B, int32_t d, int32_t L, int32_t accumulate,
                     void* stream);
/* Backward of lx_s4_kernel_gen: dK [d, L] -> dB, dCt (complex64 [d, n], +=; torch's dL/dRe + i dL/dIm convention),
 * dlog_step [d] (+=).  workspace: 72 * d * L bytes. */
int lx_s4_kernel_gen_bwd(const void* lam, const void* p, const void* q, const void* Bm, const void* Ct, const float* log_step,
                         const float* dK, void* dB, void* dCt, float* dlog_step, void* workspace, int32_t d, int32_t n,
                         int32_t L, void* stream);
/* Backward of lx_duan_forward.  fwd_ws: the workspace the forward filled for the same (x, c); dw: where the eight weight
 * gradients accumulate (same layout as w); dy has batch stride dy_bstride; dx / dc may be NULL, acc_* != 0 accumulates.
 * ws: fp32 scratch [B*C*L + B*hidden*L + B*(8*C + 2*hidden)].  The top-k channel mask is a constant (as in autograd). */
int lx_duan_backward(const lx_duan_weights_t* w, const lx_duan_weights_t* dw, const float* x, const float* c, const float* dy,
                     int64_t dy_bstride, float* dx, float* dc, int32_t acc_dx, int32_t acc_dc, int32_t B, int32_t C, int32_t L,
                     const float* fwd_ws, float* ws, void* stream);
/* out[b, k] += sum_n d[b, n] W[n, k] for a bf16 panel W [N, K] read once, B <= 8 (gradient of the conditioning vectors
 * through the DiT's AdaLN / time-text-embedding Linears, transformer.py:102-114). */
int lx_skinny_xw_bf16(const float* d, int64_t ldd, const void* W, int64_t ldw, float* out, int64_t ldo, int32_t B, int32_t N,
                      int32_t K, void* stream);
/* dpre = dy * silu'(a + b + c[m % c_rows]) for bf16 a, b, c (backward of lx_add_silu_bcast); and the fp32 form. */
int lx_silu_bwd_sum(const float* dy, const void* a, const void* b, const void* c, int32_t c_rows, float* dpre, int32_t M,
                    int32_t D, void* stream);
int lx_silu_bwd_f32(const float* dy, const float* pre, float* dpre, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Training step (OminiModel.step, src/train/model.py:569-729): rectified-flow objective around tranformer_forward with
 * gradients for the LoRA factors.  The forward runs block.py:179-339 un-fused where the backward needs an intermediate
 * (pre-norm q/k/v, pre-GELU hidden, pre-gate projection outputs); dX of every Linear is lx_gemm_bf16 against the
 * transposed weight panel.  Rows / tile_meta as above; per-stream vectors are `p[stream] + batch*stride[stream]`.
 * ------------------------------------------------------------------------------------------------------ */
/* Backward of lx_attention (F.scaled_dot_product_attention in block.py:129-131, with the same block masks / c_factor
 * bias): the five products of the FlashAttention backward on tcgen05, P^T / dS^T kept in tensor memory.
 * lx_attention_bwd_prep: dO rows (stream-major, head h at columns [128h, 128h+128)) + O rows -> dO head-major and
 * delta[b,h,s] = sum_d dO*O. */
typedef struct lx_attn_bwd_desc {
  const void* q; /* bf16 [B, H, S, 128] */
  const void* k;
  const void* v;
  const void* d_out;  /* bf16 [B, H, S, 128] */
  const float* lse;   /* fp32 [B, H, S] written by lx_attention (log2 domain) */
  const float* delta; /* fp32 [B, H, S] */
  float* dq;          /* fp32 [B, H, S, 128] accumulator, zero-initialised by the caller */
  void* dk;           /* bf16 [B, H, S, 128] */
  void* dv;
  int32_t B, H, S;
  int32_t n_cond, mask_mode;
  float cross_bias, scale;
  int32_t reserved;
  int32_t stream_end[3]; /* see lx_attn_desc_t */
  int32_t pad[3];
} lx_attn_bwd_desc_t;
int lx_attention_bwd_prep(const void* d_out_rows, int64_t ld_do, const void* out_rows, int64_t ld_o, int32_t rows,
                          int32_t heads, const lx_tile_meta_t* tile_meta, void* d_out_heads, float* delta,
                          int32_t seq_total, void* stream);
int lx_attention_bwd(const lx_attn_bwd_desc_t* desc, void* stream);
/* GELU(tanh) and its derivative on bf16 [rows, cols] views (FeedForward act / act_mlp, block.py:258-265, 302). */
int lx_gelu_fwd(const void* pre, int64_t ld_pre, void* out, int64_t ldo, int32_t rows, int32_t cols, void* stream);
int lx_gelu_bwd(const void* pre, int64_t ld_pre, const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int32_t rows,
                int32_t cols, void* stream);
/* out = res + gate[stream,batch] * y (block.py:224-234, 268-274, 328-334) and its backward:
 * dy = gate * dout; dgate[stream][batch, col] += sum_rows dout * y (fp32, skipped for NULL streams). */
int lx_gate_residual_fwd(const void* res, const void* y, void* out, int64_t ld, int32_t rows, int32_t D,
                         const lx_tile_meta_t* tile_meta, const void* const gate[3], const int64_t gate_stride[3],
                         void* stream);
int lx_gate_bwd(const void* dout, const void* y, void* dy, int64_t ld, int32_t rows, int32_t D,
                const lx_tile_meta_t* tile_meta, const void* const gate[3], const int64_t gate_stride[3],
                float* const dgate[3], const int64_t dgate_stride[3], void* stream);
/* Backward of lx_ln_modulate: dx = dres + dLN(dxn * (1 + scale)) (dres may be NULL); dscale / dshift[stream][batch, col]
 * accumulate in fp32 (NULL = skip).  stats_workspace: fp32 [rows, 2] (needed when any dscale / dshift is given). */
int lx_ln_modulate_bwd(const void* x, const void* dxn, const void* dres, void* dx, int64_t ld, int32_t rows, int32_t D,
                       const lx_tile_meta_t* tile_meta, const void* const scale[3], const int64_t scale_stride[3],
                       float* const dscale[3], float* const dshift[3], const int64_t dstride[3], float eps,
                       float* stats_workspace, void* stream);
/* q/k/v post-processing of attn_forward (block.py:34-41, 60-67, 74-99): qkv_pre rows [rows, >= 3*heads*128] ->
 * per-head RMSNorm(q,k)*w, RoPE, scatter to [B,H,S,128]; and its backward (dq/dk/dv head-major -> rows). */
int lx_qkv_post_fwd(const void* qkv_pre, int64_t ld, int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta, void* q,
                    void* k, void* v, int32_t seq_total, const float* const rms_q[3], const float* const rms_k[3],
                    const float* rope, float eps, void* stream);
int lx_qkv_post_bwd(const void* qkv_pre, int64_t ld, const void* dq, const void* dk, const void* dv, void* dqkv_pre,
                    int64_t ldo, int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta, int32_t seq_total,
                    const float* const rms_q[3], const float* const rms_k[3], const float* rope, float eps, void* stream);
/* the same with dq read straight from lx_attention_bwd's fp32 accumulation buffer [B,H,S,128] (no cast pass) */
int lx_qkv_post_bwd_f32dq(const void* qkv_pre, int64_t ld, const float* dq, const void* dk, const void* dv, void* dqkv_pre,
                          int64_t ldo, int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta, int32_t seq_total,
                          const float* const rms_q[3], const float* const rms_k[3], const float* rope, float eps, void* stream);
/* rows [rows, ld] with head h in columns [128h, 128h+128) -> [B,H,S,128] (inverse of the attention output layout). */
int lx_rows_to_heads(const void* rows_in, int64_t ld, void* heads_out, int32_t rows, int32_t heads,
                     const lx_tile_meta_t* tile_meta, int32_t seq_total, void* stream);
/* peft LoRA Linear gradients: dA[r,K] += s (dy B)^T x, dB[N,r] += s dy^T (x A^T) for bf16 x [M,K],
 * dy [M,N] and fp32 factors A [r,K], B [N,r] (r <= 16).  workspace: fp32 [2*M*r]. */
int lx_lora_grad(const void* x, int64_t ldx, const void* dy, int64_t ldy, const float* A, const float* Bw, float* dA,
                 float* dB, int32_t M, int32_t K, int32_t N, int32_t r, float scaling, float* workspace, void* stream);
/* The same gradients for up to 4 sub-Linears that share x and whose outputs are adjacent column blocks of one dy
 * (to_q | to_k | to_v (| proj_mlp), block.py:50-58, 292-300): four launches for the whole group, x and dy read twice each.
 * rank in {4, 8, 16}, groups * rank <= 16, widths multiples of 256, factors 16-byte aligned; workspace: fp32 [2*M*groups*r]. */
typedef struct {
  int32_t groups, r;
  int32_t width[4];  /* output columns of sub-Linear g inside dy */
  const float* A[4]; /* [r, K] */
  const float* B[4]; /* [width, r] */
  float* dA[4];
  float* dB[4];
  float scaling[4];
} lx_lora_stack_t;
int lx_lora_grad_stacked(const void* x, int64_t ldx, const void* dy, int64_t ldy, int32_t M, int32_t K,
                         const lx_lora_stack_t* st, float* workspace, void* stream);
/* out = bf16(W + s B A): the merged panel of the LoRA-active row group, rebuilt after every optimizer step. */
int lx_lora_merge(const void* W, int64_t ldw, const float* A, const float* Bw, void* out, int64_t ldo, int32_t N, int32_t K,
                  int32_t r, float scaling, void* stream);
/* the same merge writing the K-major copy too: outT[k, n] = out[n, k] (one pass over W) */
int lx_lora_merge_t(const void* W, int64_t ldw, const float* A, const float* Bw, void* out, int64_t ldo, void* outT,
                    int64_t ldt, int32_t N, int32_t K, int32_t r, float scaling, void* stream);
/* out[c, r] = in[r, c] (bf16): K-major W^T panels for the dX GEMMs. */
int lx_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols, void* stream);
/* Rectified-flow objective (model.py:590-594, 726): x_t = (1 - t_b) x_0 + t_b x_1;
 * loss += mean((pred - (x_1 - x_0))^2) (loss must be zeroed), dpred = grad_scale * 2 (pred - target) / n (may be NULL). */
int lx_flow_noise_mix(const void* x0, const void* x1, const float* t, void* xt, int32_t B, int64_t per_sample, void* stream);
int lx_flow_mse_loss(const void* pred, const void* x0, const void* x1, float* loss, void* dpred, int64_t n,
                     float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * VAE either side of the loop: pipeline_tools.py:7-12 `vae.encode(images).latent_dist.sample()`,
 * `(z - shift) * scale`; generate.py:375-380 `z / scale + shift`, `vae.decode(z)[0]`, `image_processor.postprocess`.
 * Third-party arithmetic (diffusers 0.31.0 AutoencoderKL, FLUX.1-dev vae/config.json) restated in oracle/vae.py.
 * Activations are bf16 rows [B*H*W, C] (NHWC); a convolution = lx_vae_im2col (GroupNorm affine + SiLU + nearest x2
 * up-sampling + stride folded into the panel write) followed by lx_gemm_bf16 against the [Cout, taps*C] weight panel
 * (columns ordered (ky, kx, c)); the residual add of ResnetBlock2D / the attention is the GEMM's LX_EPI_GATE_RESIDUAL
 * epilogue with a gate of ones.
 * ------------------------------------------------------------------------------------------------------ */
/* GroupNorm(groups, C, eps) of x [B, hw, C] as per-(sample, channel) affine coefficients:
 * coeff[b, c] = (a, b) with y = x * a + b, a = gamma[c] * rstd[b, g], b = beta[c] - mean[b, g] * a.
 * workspace: fp64, lx_vae_group_norm_workspace(B, hw, groups) elements (per-chunk partial sums; no zeroing needed, the
 * reduction order is fixed, so the result is bit-reproducible).  C = 8 * a power of two, <= 2048. */
int64_t lx_vae_group_norm_workspace(int32_t B, int64_t hw, int32_t groups);
int lx_vae_group_norm_coeffs(const void* x, int32_t B, int64_t hw, int32_t C, int32_t groups, const float* gamma,
                             const float* beta, float eps, double* workspace, float* coeff, void* stream);
typedef struct lx_vae_im2col_desc {
  const void* x;      /* bf16 [B, H, W, C] */
  void* out;          /* bf16 [B*Ho*Wo, ldk]; columns [taps*C, ldk) are written as zeros */
  const float* coeff; /* fp32 [B, C, 2] from lx_vae_group_norm_coeffs, or NULL (no normalisation) */
  int32_t B, H, W, C; /* C a multiple of 8 */
  int32_t upsample;   /* 1, or 2 = the convolution sees the nearest-neighbour x2 image (Upsample2D) */
  int32_t stride;     /* 1, or 2 (Downsample2D) */
  int32_t pad_lo;     /* zero rows/columns before the image: 1 for padding=1; 0 for Downsample2D's F.pad(x, (0,1,0,1)) */
  int32_t taps;       /* 9 = 3x3 kernel, 1 = 1x1 (normalised copy) */
  int32_t silu;       /* apply SiLU after the affine (only with coeff) */
  int32_t Ho, Wo;     /* output grid */
  int64_t ldk;        /* multiple of 8, >= taps*C */
} lx_vae_im2col_desc_t;
int lx_vae_im2col(const lx_vae_im2col_desc_t* desc, void* stream);
/* p[r, c] = softmax_c(s[r, :n] * scale) as bf16; columns [n, ldp) are written as zeros (mid-block attention, one head). */
int lx_vae_softmax_rows(const float* s, int64_t lds, void* p, int64_t ldp, int32_t rows, int32_t n, float scale, void* stream);
/* fp32 [B, C, hw] -> bf16 rows [B*hw, c_pad] = in * mul + add, channels >= C zero (image / latent entry:
 * VaeImageProcessor.normalize is mul 2, add -1; generate.py:376-378 is mul 1/scaling_factor, add shift_factor). */
int lx_vae_nchw_to_rows(const float* in, void* out, int32_t B, int32_t C, int64_t hw, int32_t c_pad, float mul, float add,
                        void* stream);
/* fp32 rows [B*hw, ld] -> fp32 [B, C, hw]; denormalize != 0 applies (x / 2 + 0.5).clamp(0, 1) (image_processor.postprocess). */
int lx_vae_rows_to_nchw(const float* in, int64_t ld, float* out, int32_t B, int32_t C, int64_t hw, int32_t denormalize,
                        void* stream);
/* moments rows [B*hw, ld] = [mean(L) | logvar(L)] -> out fp32 [B, L, hw] =
 * (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps - shift) * scale; eps fp32 [B, L, hw] or NULL (the mode). */
int lx_vae_sample_latents(const float* moments, int64_t ld, const float* eps, float* out, int32_t B, int32_t L, int64_t hw,
                          float shift, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Text encoders behind FluxPipeline.encode_prompt (generate.py:156-165, pipeline_tools.py:33-52):
 * transformers' T5EncoderModel (T5 v1.1 XXL) and CLIPTextModel (CLIP-L), restated and pinned in
 * oracle/text_encoders.py.  Every Linear is lx_gemm_bf16 (fused q|k|v and wi_0|wi_1 panels, residual adds through
 * LX_EPI_GATE_RESIDUAL); these are the remaining pieces.
 * ------------------------------------------------------------------------------------------------------ */
/* out[i, :] = table[ids[i], :] (+ pos[i % period, :]); bf16 tables [vocab, D] / [period, D]; pos may be NULL. */
int lx_embed_rows(const void* table, const int32_t* ids, const void* pos, int32_t period, void* out, int32_t n, int32_t D,
                  int32_t vocab, void* stream);
/* rms != 0: y = x * rsqrt(mean(x^2) + eps) * gamma (T5LayerNorm); else (x - mean) * rstd * gamma + beta (nn.LayerNorm);
 * bf16 rows in / out, fp32 gamma / beta [D]. */
int lx_norm_rows(const void* x, int64_t ldx, const float* gamma, const float* beta, void* out, int64_t ldo, int32_t rows,
                 int32_t D, float eps, int32_t rms, void* stream);
/* out = a * b over bf16 [rows, cols] views (T5DenseGatedActDense: gelu_new(wi_0 x) * (wi_1 x)). */
int lx_mul_rows(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int32_t rows, int32_t cols,
                void* stream);
typedef struct lx_small_attn_desc {
  const void* q;  /* bf16 rows [B*S, ld]; head h in columns [64h, 64h + 64) */
  const void* k;
  const void* v;
  int64_t ldq, ldk, ldv;
  void* out;         /* bf16 rows [B*S, ldo], same head layout */
  int64_t ldo;
  const float* bias; /* fp32 [H, S, S] added to the scaled logits (T5 relative position bias) or NULL */
  int32_t B, H, S;   /* S <= 512 */
  int32_t head_dim;  /* 64 */
  int32_t causal;    /* != 0: key j > query i is masked (CLIP text) */
  float scale;       /* logits = scale * q.k (T5: 1, CLIP: 1/8) */
  int32_t bias_relative; /* != 0: bias is fp32 [H, 2S-1], entry (key - query + S - 1): a bias that depends on the distance
                            only (T5), kept in shared memory instead of streaming H*S*S values per layer */
  int32_t reserved;
} lx_small_attn_desc_t;
int lx_attention_small(const lx_small_attn_desc_t* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOONGX_B200_H_ */
