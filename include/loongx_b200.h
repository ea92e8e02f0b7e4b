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

#ifdef __cplusplus
}
#endif
#endif /* LOONGX_B200_H_ */
