// Joint [txt | img | cond] non-causal attention for sm_100a (head_dim 128), FlashAttention-style online softmax with
// both GEMMs on tcgen05 tensor cores and the running output accumulator resident in TMEM.
//
//   one CTA = one 128-row query tile of one (batch, head)
//   warp 0      : TMA producer (Q once; K/V tiles through a 2-stage mbarrier ring, 128-byte swizzle)
//   warp 1      : tcgen05.mma issuer:  S_j = Q K_j^T  (TMEM, double buffered)  and  O += P_j V_j  (TMEM)
//   warps 2..9  : softmax, TWO threads per query row (each owns 64 of the 128 key columns of the tile, tcgen05.ld
//                 32x32b; the two halves exchange their row max through shared memory + one named barrier);
//                 exp2 (single MUFU) with the 1/sqrt(d)·log2(e) scale folded in, lazy rescale of O (only when the
//                 running max grows by > 8 in log2 units), P_j written to shared memory as the bf16 A operand of the
//                 second GEMM.  The MUFU pipe (16 exp2/clk/SM = 1024 clk per 128x128 tile) then matches the tensor pipe
//                 (2 x 512 clk per tile) instead of trailing it by 3x.
//
// The issue order on the tensor pipe is S_0, [S_1, PV_0], [S_2, PV_1], ... so softmax(j) overlaps PV(j-1) and S(j+1).
//
// Replaces F.scaled_dot_product_attention + the q/k/v concat + head transpose at block.py:70-72,102-104,129-135,
// including the optional block masks (block.py:106-120) and the log(c_factor) bias (block.py:121-128), which are
// uniform per 128x128 tile because every stream length is a multiple of 128.
#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

constexpr int ATT_BQ = 128, ATT_BKV = 128, ATT_D = 128;
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;  // 32 KiB: two 128x64 swizzle atoms
constexpr int ATT_ATOM_BYTES = 128 * 64 * 2;   // 16 KiB
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_THREADS = 320;  // TMA warp + MMA warp + 8 softmax warps
constexpr int ATT_SOFTMAX_THREADS = 256;
constexpr int ATT_SMEM = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES + 1) + 1024 + 256 + 2 * 2 * 128 * 4;

struct AttnParams {
  long long* dbg;  // optional timeline buffer (development aid, NULL in production)
  lx_attn_desc_t d;
  float scale_log2;  // scale * log2(e)
  float bias_log2;   // cross_bias * log2(e)
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;                    // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;    // [stages]
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + ATT_TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* k_empty = bars + 5;     // [2] released by the commit after QK^T(it): K loads run a full iteration ahead
  uint64_t* v_empty = bars + 14;    // [2] released by the commit after PV(it)
  uint64_t* s_full = bars + 7;      // [2]
  uint64_t* s_empty = bars + 9;     // [2]
  uint64_t* p_full = bars + 11;     // 1
  uint64_t* pv_done = bars + 12;    // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* s_xchg = reinterpret_cast<float*>(bars + 32);  // [2 (parity)][2 (column half)][128 rows] row-max / row-sum exchange

  const lx_attn_desc_t& d = p.d;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = d.S / ATT_BKV;
  const int n_rest = (d.S - d.n_cond) / ATT_BKV;  // tiles of the non-condition part
  const bool q_is_cond = qt >= n_rest;
  // kv tile range visible to this query tile
  int kv_begin = 0, kv_end = n_tiles;
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
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], ATT_SOFTMAX_THREADS);
    }
    mbar_init(p_full, ATT_SOFTMAX_THREADS);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // two 128-column buffers
  const uint32_t tmem_O = tmem_base + 256;  // 128 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, 0, head_row0 + qt * ATT_BQ);
      tma_load_2d(sQ + ATT_ATOM_BYTES, &tmQ, q_full, 64, head_row0 + qt * ATT_BQ);
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
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);  // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);   // A = P (K-major), B = V (MN-major)
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_qk = [&](int it) {
        const int st = it & 1;
        mbar_wait(&k_full[st], (it >> 1) & 1);
        mbar_wait(&s_empty[st], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * ATT_ATOM_BYTES + (kk & 3) * 32;
          umma_ss(tmem_S + st * 128, make_sdesc_sw128(q_addr + off, 16, 1024), make_sdesc_sw128(k_addr + off, 16, 1024),
                  idesc_qk, kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_qk(it + 1);
        const int st = it & 1;
        mbar_wait(&v_full[st], (it >> 1) & 1);
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
          // A: P[128 x 16] slice kk (K-major); B: V rows [16kk, 16kk+16) x 128 d (MN-major: LBO = next 64-d atom)
          const uint64_t da = make_sdesc_sw128(p_addr + (kk >> 2) * ATT_ATOM_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t db = make_sdesc_sw128(v_addr + kk * 2048, ATT_ATOM_BYTES, 1024);
          umma_ss(tmem_O, da, db, idesc_pv, (it | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which 64 key columns (and which 64 output columns) this thread owns
    const int r = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float m_run = -INFINITY;  // running (possibly stale) max in log2 units, identical in both halves of a row
    float l_run = 0.f;        // partial row sum over this thread's columns
    const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
#define DBG(slot) if (dbg_on) p.dbg[it * 8 + (slot)] = clock64();
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const bool cross = use_bias && (q_is_cond != ((kv_begin + it) >= n_rest));
      DBG(0)
      const float bias = cross ? p.bias_log2 : 0.f;
      mbar_wait(&s_full[st], (it >> 1) & 1);
      tc_fence_after();
      DBG(1)
      const uint32_t ts = tmem_S + st * 128 + half * 64 + lane_off;
      uint32_t v[64];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld_32x32b_x32(ts, v0);
        tmem_ld_32x32b_x32(ts + 32, v1);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[st]);  // S is in registers: the buffer may be overwritten by QK(it+2)
      DBG(2)
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      float* xch = s_xchg + (it & 1) * 256;
      xch[half * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
      DBG(3)
      const float m_tile = mx * p.scale_log2 + bias;
      float alpha = 1.0f;
      bool rescale = false;
      if (m_tile > m_run + 8.0f) {  // also true on the first tile (m_run = -inf)
        alpha = ex2_approx(m_run - m_tile);  // 0 on the first tile
        m_run = m_tile;
        rescale = it > 0;
      }
      uint32_t pk[32];
      float lsum = 0.f;
      const float moff = bias - m_run;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, moff));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, moff));
        lsum += p0 + p1;
        pk[j] = pack_bf16(p0, p1);
      }
      l_run = l_run * alpha + lsum;
      DBG(4)
      if (it > 0) {
        mbar_wait(pv_done, (it - 1) & 1);  // PV(it-1) finished: O is stable and sP is free
        tc_fence_after();
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_O + lane_off + half * 64 + c * 32, o);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st_32x32b_x32(tmem_O + lane_off + half * 64 + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      DBG(5)
      // this thread's 64 P columns = one 128-byte-swizzled atom row
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        uint4 val = make_uint4(pk[4 * c16], pk[4 * c16 + 1], pk[4 * c16 + 2], pk[4 * c16 + 3]);
        *reinterpret_cast<uint4*>(sP + half * ATT_ATOM_BYTES + r * 128 + ((c16 ^ (r & 7)) << 4)) = val;
      }
      DBG(6)
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      DBG(7)
    }
    // epilogue: O / l -> bf16 -> out rows (each thread writes its 64 output columns)
    float* xch = s_xchg + (n_it & 1) * 256;
    xch[half * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv_l = 1.0f / (l_run + xch[(half ^ 1) * 128 + r]);
    mbar_wait(pv_done, (n_it - 1) & 1);
    tc_fence_after();
    const int out_row = d.out_row_base[b * n_tiles + qt] + r;
    __nv_bfloat16* out =
        reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)out_row * d.ldo + d.col_offset + h * ATT_D + half * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_O + lane_off + half * 64 + c * 32, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16(__uint_as_float(o[8 * j + 0]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
        u.y = pack_bf16(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
        u.z = pack_bf16(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
        u.w = pack_bf16(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
        *reinterpret_cast<uint4*>(out + c * 32 + j * 8) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lx

static long long* g_attn_dbg = nullptr;
// development aid: per-iteration clock64 timeline of CTA (0,0,0)'s first softmax thread, 8 slots per KV iteration
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
  const uint64_t rows = (uint64_t)d.B * d.H * d.S;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmQ, d.q, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmK, d.k, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmV, d.v, rows, 128, 128, 128, 64))) return rc;
  AttnParams p;
  p.d = d;
  p.dbg = g_attn_dbg;
  const float log2e = 1.4426950408889634f;
  p.scale_log2 = d.scale * log2e;
  p.bias_log2 = d.cross_bias * log2e;
  static bool attr_set = false;
  if (!attr_set) {
    LX_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid(d.S / 128, d.H, d.B);
  double pairs = (double)d.S * d.S;  // visible (query, key) pairs per head
  if (d.cross_bias == 0.f && d.n_cond > 0) {
    const double nc = d.n_cond, nr = d.S - d.n_cond;
    if (d.mask_mode == 1) pairs = nc * nc + nr * nr;
    if (d.mask_mode == 2) pairs = nc * nc + nr * (double)d.S;
  }
  LaunchScope scope(KC_ATTENTION, stream, 4.0 * d.B * d.H * pairs * 128.0);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
