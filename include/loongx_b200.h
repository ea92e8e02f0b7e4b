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
  LX_EPI_BIAS_F32 = 5       /* out = acc + bias                                   -> fp32                */
};

typedef struct lx_gemm_segment {
  int32_t mode;       /* lx_epilogue */
  int32_t col_offset; /* output column = (n - first column of the segment) + col_offset */
  void* out;          /* [M, ldo] */
  int64_t ldo;        /* elements */
} lx_gemm_segment_t;

typedef struct lx_gemm_desc {
  const void* A; /* bf16 [M, K], row stride lda (elements, multiple of 8) */
  int64_t lda;
  const void* W; /* bf16 [N, K], row stride ldw */
  int64_t ldw;
  const float* bias; /* fp32 [N] or NULL */
  int32_t M, N, K;
  int32_t n_split; /* output columns >= n_split use seg[1] (multiple of 256); = N for a single segment */
  lx_gemm_segment_t seg[2];
  const lx_tile_meta_t* tile_meta; /* [ceil(M/128)]; required by GATE_RESIDUAL and QKV */
  /* LX_EPI_GATE_RESIDUAL */
  const void* residual; /* bf16 [M, ldr] */
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
  int32_t reserved;
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
} lx_attn_desc_t;

int lx_attention(const lx_attn_desc_t* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound).
 * ------------------------------------------------------------------------------------------------------ */
/* out[:, 0:D] = LayerNorm(x; no affine, eps) * (1 + scale[stream,batch]) + shift[stream,batch];
 * out[:, D:D+ext] = (that bf16 row) . lora_a^T for rows whose stream is in lora_stream_mask, else 0.
 * Replaces AdaLayerNormZero / -Single / -Continuous + norm2 FiLM (block.py:192-207, 238-253, 301-305;
 * transformer.py:243) and peft's lora_A on the same operand. */
typedef struct lx_lnmod_desc {
  const void* x; /* bf16 [rows, ldx] */
  int64_t ldx;
  void* out; /* bf16 [rows, ldo], ldo >= D + ext */
  int64_t ldo;
  int32_t rows, D;
  int32_t ext;    /* K-extension columns to fill (multiple of 64, 0 = none) */
  int32_t lora_r; /* rows of lora_a (<= 16) */
  const lx_tile_meta_t* tile_meta;
  const void* shift[3]; /* bf16, per stream: shift[s] + batch*stride[s] */
  const void* scale[3];
  int64_t stride[3];
  const void* lora_a; /* bf16 [lora_r, D] or NULL */
  int32_t lora_stream_mask; /* bit s set = LoRA active on stream s (default: cond only = 4) */
  float eps;
} lx_lnmod_desc_t;
int lx_ln_modulate(const lx_lnmod_desc_t* desc, void* stream);

/* x[:, K:K+ext] = x[:, 0:K] . lora_a^T (active rows) or 0: the LoRA "down" half for operands produced by another
 * kernel (attention output, GELU hidden, packed latents). */
typedef struct lx_lora_down_desc {
  void* x; /* bf16 [rows, ldx], ldx >= K + ext */
  int64_t ldx;
  int32_t rows, K, ext, lora_r;
  const lx_tile_meta_t* tile_meta; /* NULL = every row is a condition row */
  const void* lora_a;              /* bf16 [lora_r, K] or NULL */
  int32_t lora_stream_mask;
  int32_t reserved;
} lx_lora_down_desc_t;
int lx_lora_down(const lx_lora_down_desc_t* desc, void* stream);

/* Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): out[m] = [cos(t*mult*f), sin(t*mult*f)] bf16. */
int lx_timestep_embed(const float* t, void* out, int64_t ldo, int32_t M, float mult, void* stream);
/* out[m, 0:D] = silu(a[m] + b[m] + c[m % c_rows]) for bf16 [M, D] inputs (b, c may be NULL): temb = t_emb + g_emb +
 * text_emb followed by the SiLU every AdaLN applies (transformer.py:102-114, App. A.2/A.5). */
int lx_add_silu_bcast(const void* a, const void* b, const void* c, int32_t c_rows, void* out, int64_t ldo, int32_t M,
                      int32_t D, void* stream);
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
 * Weight layout: every nn.Linear is one lx_linear_t.  A LoRA-targeted Linear stores its weight K-extended:
 * columns [0,k) = base W, columns [k, k+ext) = lora_B * (alpha/r) (zero padded to a multiple of 64), and
 * lora_a holds the stacked lora_A rows; the operand producers fill the matching ext columns with x.A^T on the
 * rows where LoRA is active (condition stream by default, lora_controller.py:5-43), so base + LoRA is ONE GEMM.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct lx_linear {
  const void* w; /* bf16 [n, ldw] */
  int64_t ldw;
  const float* bias;  /* fp32 [n] or NULL */
  const void* lora_a; /* bf16 [lora_r, k] or NULL */
  int32_t n, k, ext, lora_r;
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
  lx_linear_t mod_img;    /* all double blocks' norm1.linear stacked:         [L*6D, D+ext] */
  lx_linear_t mod_txt;    /* all double blocks' norm1_context.linear stacked: [L*6D, D]     */
  lx_linear_t mod_single; /* all single blocks' norm.linear stacked:          [Ls*3D, D+ext] */
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
  void* XN;      /* bf16 [R, D+64] modulated operand */
  void* Q;       /* bf16 [B,H,S,128] */
  void* K;
  void* V;
  void* scratch; /* bf16 [R, 5D+64]: attention out / FF hidden / single-block concat */
  void* XE;      /* bf16 [B*max(n_img,n_cond), in_channels+64] packed-latent operand of x_embedder */
  void* X0_txt;  /* bf16 [B*n_txt, D]  context_embedder(prompt_embeds)   (step invariant) */
  void* X0_cond; /* bf16 [B*n_cond, D] x_embedder(cond_latents)          (step invariant) */
  void* emb_tmp; /* bf16 [4, T*B + B, D] scratch of the timestep / guidance / text MLPs */
  void* sin_tmp; /* bf16 [T*B + B, 256] */
  void* silu_t;  /* bf16 [T*B, D + max ext] silu(temb) per (step, batch) */
  void* silu_c;  /* bf16 [B,   D + max ext] silu(cond_temb) */
  void* mod_img;      /* bf16 [T*B, L*6D] */
  void* mod_txt;      /* bf16 [T*B, L*6D] */
  void* mod_single;   /* bf16 [T*B, Ls*3D] */
  void* mod_out;      /* bf16 [T*B, 2D] */
  void* mod_cond_img;    /* bf16 [B, L*6D]  (cond stream, step invariant) */
  void* mod_cond_single; /* bf16 [B, Ls*3D] */
  float* t_dev;          /* fp32 [T*B + B] device scratch for timestep values */
  float* g_dev;          /* fp32 [T*B + B] device scratch for guidance values */
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
/* Reference-granularity entry points for parity tests (block.py:179-278 / 281-339): run ONE block of the prepared
 * plan in place on plan->X. */
int lx_dit_double_block(const lx_dit_model_t* model, const lx_dit_plan_t* plan, int32_t step, int32_t block,
                        void* stream);
int lx_dit_single_block(const lx_dit_model_t* model, const lx_dit_plan_t* plan, int32_t step, int32_t block,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOONGX_B200_H_ */
