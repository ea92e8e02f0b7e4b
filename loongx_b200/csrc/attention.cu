// Joint [txt | img | cond] non-causal attention for sm_100a (head_dim 128), FlashAttention-style online softmax with
// both GEMMs on tcgen05 tensor cores and S, P and the running output accumulator all resident in TMEM.
//
//   one CTA = TWO 128-row query tiles (A, B) of one (batch, head), ping-ponged on the tensor pipe
//   warp 0      : TMA producer (both Q tiles once; K and V tiles through 2-stage mbarrier rings, 128-byte swizzle)
//   warp 1      : tcgen05.mma issuer.  Issue order  S_A0 S_B0 | PV_A0 S_A1 PV_B0 S_B1 | PV_A1 S_A2 PV_B1 S_B2 | ...
//                 S_g = Q_g K_j^T   (SS form, TMEM columns [128g, 128g+128))
//                 O_g += P_g V_j    (TS form: the bf16 probabilities are read from TMEM, where the softmax threads
//                                    packed them in place over the first 64 columns of S_g; V is the MN-major smem operand)
//   warps 4..7  : softmax of tile A, one thread per query row (tcgen05.ld 32x32b: no cross-lane reductions)
//   warps 8..11 : softmax of tile B   (setmaxnreg moves warpgroup 0's registers to these two: 240 regs / thread, the
//                 whole 128-column S row stays in registers between the max and the exp pass)
//                 exp2 (single MUFU.EX2) with the 1/sqrt(d)·log2(e) scale folded into one FFMA, lazy rescale of O (only
//                 when the running max grows by more than 8 in log2 units), P never touches shared memory.
//   While one tile's rows are in the MUFU-bound softmax (16 exp2/clk/SM = 1024 clk per 128x128 tile) the tensor pipe
//   runs the other tile's PV + next QK^T (2 x 512 clk), so the two pipes overlap instead of alternating.
//   A launch whose tile counts are odd (tiny test shapes) runs with one query tile per CTA (group B idle).
//
// Replaces F.scaled_dot_product_attention + the q/k/v concat + head transpose at block.py:70-72,102-104,129-135,
// including the optional block masks (block.py:106-120) and the log(c_factor) bias (block.py:121-128), which are
// uniform per 128x128 tile because every stream length is a multiple of 128.
#include "host_util.cuh"
#include "ptx.cuh"
#include "tmem_wide.cuh"

namespace lx {

constexpr int ATT_BQ = 128, ATT_BKV = 128, ATT_D = 128;
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;  // 32 KiB: two 128x64 swizzle atoms
constexpr int ATT_ATOM_BYTES = 128 * 64 * 2;   // 16 KiB
constexpr int ATT_KV_STAGES = 2;
#ifndef ATT_POLY_OF_8
#define ATT_POLY_OF_8 2
#endif
constexpr int ATT_THREADS = 384;  // warpgroup 0: TMA warp + MMA warp (+2 idle); warpgroups 1, 2: softmax of tiles A, B
constexpr int ATT_SMEM = ATT_TILE_BYTES * (2 + 2 * ATT_KV_STAGES) + 1024 + 256 + 2 * 2 * 2 * 128 * 4;

struct AttnParams {
  long long* dbg;  // optional timeline buffer (development aid, NULL in production)
  long long* cta_trace;  // optional [n_ctas][6]: smid, t_entry, t_setup_done, t_first_s, t_loop_end, t_exit (CTA-level)
  lx_attn_desc_t d;
  float scale_log2;  // scale * log2(e)
  float bias_log2;   // cross_bias * log2(e)
  int groups;        // query tiles per CTA (2, or 1 when the tile counts are odd)
};

// kPad: ragged streams (padding keys at the end of a stream's last tile are masked); a separate instantiation so that
// the aligned case pays nothing for it.
template <bool kPad>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ AttnParams p) {
  const long long t_entry = clock64();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 groups]
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;                // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;    // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* k_empty = bars + 5;     // [2] released by the commit after the last group's QK^T(it)
  uint64_t* v_empty = bars + 7;     // [2] released by the commit after the last group's PV(it)
  uint64_t* s_full = bars + 9;      // [2 groups] S_g(it) complete (and with it every earlier MMA, incl. PV_g(it-1))
  uint64_t* p_full = bars + 11;     // [2 groups] P_g(it) packed into TMEM by all 128 rows
  uint64_t* o_done = bars + 13;     // [2 groups] last PV_g complete
  uint64_t* p_half = bars + 20;     // [2 groups] first 64 key columns of P_g(it) packed (PV k-slices 0..3 may start)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const lx_attn_desc_t& d = p.d;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = p.groups;
  const int h = blockIdx.y, b = blockIdx.z;
  const int qt0 = blockIdx.x * G;  // first query tile of this CTA
  const int n_tiles = d.S / ATT_BKV;
  const int n_rest = (d.S - d.n_cond) / ATT_BKV;  // tiles of the non-condition part
  // all query tiles of a CTA are on the same side of the rest|cond boundary (G == 2 only when n_rest is even)
  const bool q_is_cond = qt0 >= n_rest;
  int kv_begin = 0, kv_end = n_tiles;  // kv tile range visible to this CTA's query tiles
  const bool use_bias = p.bias_log2 != 0.0f;
  if (!use_bias && d.n_cond > 0) {
    if (q_is_cond && (d.mask_mode == 1 || d.mask_mode == 2)) kv_begin = n_rest;
    if (!q_is_cond && d.mask_mode == 1) kv_end = n_rest;
  }
  const int n_it = kv_end - kv_begin;
  const int head_row0 = (b * d.H + h) * d.S;  // first row of this head in the [(B*H*S), 128] view

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&p_half[s], 128);
      mbar_init(&o_done[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long* trace = nullptr;
  if (p.cta_trace != nullptr && threadIdx.x == 128) {  // first softmax thread of tile A
    trace = p.cta_trace + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 6;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[0] = smid;
    trace[1] = t_entry;
    trace[2] = clock64();
  }
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: the prologue above overlapped the QKV GEMM's tail; Q / K / V are read from here on
  pdl_launch_dependents();
  // columns [128g, 128g+128): S_g fp32, its first 64 columns re-used for the packed bf16 P_g;  [256+128g, +128): O_g
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp < 4) {
  reg_dealloc<96>();  // setmaxnreg: hand the producer warpgroup's registers to the softmax warpgroups
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, G * ATT_TILE_BYTES);
      for (int g = 0; g < G; ++g) {
        tma_load_2d(sQ + g * ATT_TILE_BYTES, &tmQ, q_full, 0, head_row0 + (qt0 + g) * ATT_BQ);
        tma_load_2d(sQ + g * ATT_TILE_BYTES + ATT_ATOM_BYTES, &tmQ, q_full, 64, head_row0 + (qt0 + g) * ATT_BQ);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint32_t par = ((it >> 1) & 1) ^ 1;
        const int row = head_row0 + (kv_begin + it) * ATT_BKV;
        mbar_wait(&k_empty[st], par);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_2d(sK + st * ATT_TILE_BYTES, &tmK, &k_full[st], 0, row);
        tma_load_2d(sK + st * ATT_TILE_BYTES + ATT_ATOM_BYTES, &tmK, &k_full[st], 64, row);
        mbar_wait(&v_empty[st], par);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(sV + st * ATT_TILE_BYTES, &tmV, &v_full[st], 0, row);
        tma_load_2d(sV + st * ATT_TILE_BYTES + ATT_ATOM_BYTES, &tmV, &v_full[st], 64, row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);  // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);   // A = P (TMEM),    B = V (MN-major)
      // Base descriptors are built once; per MMA only the 14-bit address field (bytes >> 4) is advanced, so the issue
      // loop is UIADD + UTCHMMA (the tensor pipe needs a new 128x128x16 MMA every 64 clk).
      const uint64_t q_desc0 = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc0 = make_sdesc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc0 = make_sdesc_sw128(smem_u32(sV), ATT_ATOM_BYTES, 1024);
      auto issue_qk = [&](int g, int it) {
        const int st = it & 1;
        mbar_wait(&k_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t qd = q_desc0 + (uint64_t)(g * (ATT_TILE_BYTES >> 4));
        const uint64_t kd = k_desc0 + (uint64_t)(st * (ATT_TILE_BYTES >> 4));
        const uint32_t ts_g = tmem_S + g * 128;
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk) {
          const uint32_t off = ((kk >> 2) * ATT_ATOM_BYTES + (kk & 3) * 32) >> 4;
          umma_ss(ts_g, qd + off, kd + off, idesc_qk, kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
        if (g == G - 1) umma_commit(&k_empty[st]);
      };
      const bool mma_dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      mbar_wait(q_full, 0);
      for (int g = 0; g < G; ++g) issue_qk(g, 0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint64_t vd = v_desc0 + (uint64_t)(st * (ATT_TILE_BYTES >> 4));
        for (int g = 0; g < G; ++g) {
          const uint32_t to_g = tmem_O + g * 128, tp_g = tmem_S + g * 128;
          mbar_wait(&v_full[st], (it >> 1) & 1);
          // A: P_g[128 x 16] slice kk = 8 TMEM columns of packed bf16 pairs; B: V rows [16kk, 16kk+16) x 128 d
          // (MN-major: LBO = next 64-d atom).  The first four k-slices start as soon as the first 64 key columns of P
          // are packed, while the softmax threads are still exponentiating the other 64.
          mbar_wait(&p_half[g], it & 1);
          tc_fence_after();
          if (it == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ts(to_g, tp_g + kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ts(to_g, tp_g + kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, 1u);
          }
          mbar_wait(&p_full[g], it & 1);
          tc_fence_after();
          if (mma_dbg) p.dbg[(g * n_it + it) * 8 + 6] = clock64();
#pragma unroll
          for (int kk = 4; kk < 8; ++kk)
            umma_ts(to_g, tp_g + kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, 1u);
          if (g == G - 1) umma_commit(&v_empty[st]);
          // S_g(it+1) overwrites the columns P_g(it) is read from: the tensor pipe executes in issue order
          if (it + 1 < n_it) issue_qk(g, it + 1);
          else umma_commit(&o_done[g]);
          if (mma_dbg) p.dbg[(g * n_it + it) * 8 + 7] = clock64();
        }
      }
    }
    __syncwarp();
  }
  } else {
    reg_alloc<200>();
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int g = (warp - 4) >> 2;      // query tile of this warp
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    if (g < G) {
      const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t ts = tmem_S + g * 128 + lane_off;  // S_g row (fp32) / packed P_g row (first 64 columns)
      const uint32_t to = tmem_O + g * 128 + lane_off;
      float m_run = -INFINITY;  // running (possibly stale) max in log2 units
      float l_run = 0.f;
      const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                          quarter == 2;
#define DBG(slot) if (dbg_on) p.dbg[(g * n_it + it) * 8 + (slot)] = clock64();
      for (int it = 0; it < n_it; ++it) {
        const bool cross = use_bias && (q_is_cond != ((kv_begin + it) >= n_rest));
        const float bias = cross ? p.bias_log2 : 0.f;
        DBG(0)
        mbar_wait(&s_full[g], it & 1);
        tc_fence_after();
        DBG(1)
        if (trace != nullptr && it == 0) trace[3] = clock64();
        uint32_t v[128];
        {
          uint32_t(&v0)[64] = *reinterpret_cast<uint32_t(*)[64]>(&v[0]);
          uint32_t(&v1)[64] = *reinterpret_cast<uint32_t(*)[64]>(&v[64]);
          tmem_ld_32x32b_x64(ts, v0);
          tmem_ld_32x32b_x64(ts + 64, v1);
        }
        if (kPad) {  // ragged streams: the last key tile of a stream may end in padding tokens -> -inf logits
          const int kend = (kv_begin + it + 1) * ATT_BKV;
          int nvalid = ATT_BKV;
#pragma unroll
          for (int s3 = 0; s3 < 3; ++s3)
            if (kend == d.stream_end[s3]) nvalid -= d.pad[s3];
          if (nvalid < ATT_BKV) {
#pragma unroll
            for (int j = 0; j < 128; ++j)
              if (j >= nvalid) v[j] = 0xff800000u;
          }
        }
        DBG(2)
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 128; j += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(v[j]));
          mx1 = fmaxf(mx1, __uint_as_float(v[j + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(v[j + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(v[j + 3]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const float m_tile = mx * p.scale_log2 + bias;
        float alpha = 1.0f;
        bool rescale = false;
        if (m_tile > m_run + 8.0f) {  // also true on the first tile (m_run = -inf)
          alpha = ex2_approx(m_run - m_tile);  // 0 on the first tile
          m_run = m_tile;
          rescale = it > 0;
        }
        DBG(3)
        // s_full(it) also covers PV_g(it-1): O_g is stable here
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(to + c * 32, o);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32b_x32(to + c * 32, o);
          }
        }
        float2 ls = make_float2(0.f, 0.f);
        const float moff = bias - m_run;
        const float2 scale2 = make_float2(p.scale_log2, p.scale_log2), moff2 = make_float2(moff, moff);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // packed fp32x2 FMA / ADD halve the issue slots; ATT_POLY_OF_8 of every 8 pairs are exponentiated by a
            // polynomial on the FMA pipe, the rest by the MUFU
            float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c * 32 + 2 * j]), __uint_as_float(v[c * 32 + 2 * j + 1])),
                                  scale2, moff2);
            float2 e;
            if ((j & 7) < ATT_POLY_OF_8) {
              e = ex2_poly2(x);
            } else {
              e.x = ex2_approx_ordered(x.x);
              e.y = ex2_approx_ordered(x.y);
            }
            ls = __fadd2_rn(ls, e);
            pk[j] = pack_bf16(e.x, e.y);
          }
          tmem_st_32x32b_x16(ts + c * 16, pk);
          if (c == 1) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_half[g]);
          }
        }
        l_run = l_run * alpha + (ls.x + ls.y);
        DBG(4)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
        DBG(5)
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0)
          p.dbg[2 * n_it * 8 + (g * n_it + it) * 4 + quarter] = clock64();
      }
      // epilogue: O / l -> bf16 -> out rows
      if (trace != nullptr) trace[4] = clock64();
      const float inv_l = 1.0f / l_run;
      if (d.lse != nullptr) d.lse[head_row0 + (qt0 + g) * ATT_BQ + r] = m_run + log2f(l_run);  // for lx_attention_bwd
      mbar_wait(&o_done[g], 0);
      tc_fence_after();
      // O_g / l -> bf16 -> this tile's (now idle) Q buffer in the 128-byte-swizzled TMA layout -> two TMA stores of
      // 128 rows x 64 columns: full-line writes instead of 16-byte fragments at a 6 KB row stride
      uint8_t* stage = sQ + g * ATT_TILE_BYTES;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(to + c * 32, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[8 * j + 0]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
          u.y = pack_bf16(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
          u.z = pack_bf16(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
          u.w = pack_bf16(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
          const int c16 = c * 4 + j;  // 16-byte chunk of the 256-byte output row
          *reinterpret_cast<uint4*>(stage + (c16 >> 3) * ATT_ATOM_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
      if (warp == 4 + 4 * g && elect_one()) {
        const int out_row0 = d.out_row_base[b * n_tiles + qt0 + g];
        const int col0 = d.col_offset + h * ATT_D;
        tma_store_2d(&tmO, stage, col0, out_row0);
        tma_store_2d(&tmO, stage + ATT_ATOM_BYTES, col0 + 64, out_row0);
        tma_store_commit();
        tma_store_wait_read<0>();  // the staging buffer must outlive the bulk read; global visibility at kernel end
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (trace != nullptr) trace[5] = clock64();
}

}  // namespace lx

static long long* g_attn_dbg = nullptr;
static long long* g_attn_trace = nullptr;
extern "C" void lx_attention_debug_cta_trace(long long* device_buffer) { g_attn_trace = device_buffer; }
// development aid: per-iteration clock64 timeline of CTA (0,0,0): 8 slots per (query tile, KV iteration)
extern "C" void lx_attention_debug_timeline(long long* device_buffer) { g_attn_dbg = device_buffer; }

extern "C" int lx_attention(const lx_attn_desc_t* desc, void* stream) {
  using namespace lx;
  LX_CHECK_ARG(desc != nullptr, "lx_attention: null descriptor");
  const lx_attn_desc_t& d = *desc;
  LX_CHECK_ARG(d.q && d.k && d.v && d.out && d.out_row_base, "lx_attention: null pointer");
  LX_CHECK_ARG(d.B > 0 && d.H > 0 && d.S > 0 && d.S % 128 == 0, "lx_attention: S=%d must be a positive multiple of 128",
               d.S);
  LX_CHECK_ARG(d.n_cond >= 0 && d.n_cond % 128 == 0 && d.n_cond < d.S, "lx_attention: bad n_cond=%d", d.n_cond);
  LX_CHECK_ARG(d.mask_mode >= 0 && d.mask_mode <= 2, "lx_attention: bad mask_mode=%d", d.mask_mode);
  LX_CHECK_ARG(d.ldo % 8 == 0 && d.col_offset % 8 == 0, "lx_attention: ldo / col_offset must be multiples of 8");
  LX_CHECK_ARG(d.H <= 65535 && d.B <= 65535, "lx_attention: grid too large");
  for (int s3 = 0; s3 < 3; ++s3)
    LX_CHECK_ARG(d.pad[s3] >= 0 && d.pad[s3] < 128 && (d.pad[s3] == 0 || (d.stream_end[s3] > 0 && d.stream_end[s3] <= d.S &&
                                                                           d.stream_end[s3] % 128 == 0)),
                 "lx_attention: bad padding description for stream %d", s3);
  const uint64_t rows = (uint64_t)d.B * d.H * d.S;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmQ, d.q, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmK, d.k, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmV, d.v, rows, 128, 128, 128, 64))) return rc;
  CUtensorMap tmO;  // output rows in the stream-major activation layout: [B*S, ldo], head h at columns col_offset + 128 h
  if ((rc = make_tmap_2d_bf16(&tmO, d.out, (uint64_t)d.B * d.S, (uint64_t)d.ldo, (uint64_t)d.ldo, 128, 64))) return rc;
  AttnParams p;
  p.d = d;
  p.dbg = g_attn_dbg;
  p.cta_trace = g_attn_trace;
  const float log2e = 1.4426950408889634f;
  p.scale_log2 = d.scale * log2e;
  p.bias_log2 = d.cross_bias * log2e;
  const int n_tiles = d.S / 128, n_rest = (d.S - d.n_cond) / 128;
  p.groups = (n_tiles % 2 == 0 && n_rest % 2 == 0) ? 2 : 1;
  const bool has_pad = (d.pad[0] | d.pad[1] | d.pad[2]) != 0;
  static bool attr_set = false;
  if (!attr_set) {
    LX_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    LX_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  const int q_tiles = d.q_tiles > 0 ? d.q_tiles : n_tiles;
  LX_CHECK_ARG(q_tiles <= n_tiles, "lx_attention: q_tiles=%d exceeds S/128=%d", q_tiles, n_tiles);
  if (q_tiles % 2) p.groups = 1;
  dim3 grid(q_tiles / p.groups, d.H, d.B);
  double pairs = (double)d.S * d.S;  // visible (query, key) pairs per head
  if (d.cross_bias == 0.f && d.n_cond > 0) {
    const double nc = d.n_cond, nr = d.S - d.n_cond;
    if (d.mask_mode == 1) pairs = nc * nc + nr * nr;
    if (d.mask_mode == 2) pairs = nc * nc + nr * (double)d.S;
  }
  LaunchScope scope(KC_ATTENTION, stream, 4.0 * d.B * d.H * pairs * 128.0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = ATT_SMEM;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (has_pad) cudaLaunchKernelEx(&cfg, attention_kernel<true>, tmQ, tmK, tmV, tmO, p);
  else cudaLaunchKernelEx(&cfg, attention_kernel<false>, tmQ, tmK, tmV, tmO, p);
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      cudaFuncAttributes fa{};
      cudaFuncGetAttributes(&fa, attention_kernel<false>);
      set_error("lx_attention launch: %s (regs %d, max threads/block %d, local %zu B, dyn smem %d)", cudaGetErrorString(e),
                fa.numRegs, fa.maxThreadsPerBlock, fa.localSizeBytes, ATT_SMEM);
      return LX_ERR_CUDA;
    }
  }
  return LX_OK;
}
