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
#include <stdlib.h>
#include <string.h>

#include "host_util.cuh"
#include "ptx.cuh"
#include "tmem_wide.cuh"

namespace lx {

constexpr int ATT_BQ = 128, ATT_BKV = 128, ATT_D = 128;
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;  // 32 KiB: two 128x64 swizzle atoms
constexpr int ATT_ATOM_BYTES = 128 * 64 * 2;   // 16 KiB
constexpr int ATT_KV_STAGES = 2;
// pairs of every 8 whose exponentials run as a degree-3 polynomial on the FMA pipe instead of the MUFU.  Burst timing at
// 1.9 GHz preferred 2; at the power cap, where the loop runs, the 6 extra issue slots per element cost more energy than
// the MUFU slot they free: sustained 889 (2) / 914 (1) / 912 (0) / 876 (3) / 831 (4) TFLOP/s, in-loop 727 -> 752
// (scripts/attn_sustained.py, profiles/gpurun_logs/attn_poly_sweep_r2.log)
#ifndef ATT_POLY_OF_8
#define ATT_POLY_OF_8 1
#endif
constexpr int ATT_THREADS = 384;  // warpgroup 0: TMA warp + MMA warp (+2 idle); warpgroups 1, 2: softmax of tiles A, B
// Q (2 tiles) + K/V rings + ONE output staging tile shared by the two groups + barriers + alignment slack
constexpr int ATT_SMEM = ATT_TILE_BYTES * (2 + 2 * ATT_KV_STAGES + 1) + 1024 + 512;
static_assert(ATT_SMEM <= 227 * 1024, "attention shared memory budget");
// split-work exchange slot of one (CTA, query tile): un-normalised fp32 O[128 x 128] + (m, l)[128]
constexpr int ATT_SLOT_FLOATS = 128 * 128 + 2 * 128;
constexpr int ATT_WS_FLAG_BYTES = WS_FLAG_BYTES;  // flags [n_cta][2] at the start of the workspace region

struct AttnParams {
  long long* tl;  // development-aid timeline row (ptx.cuh::timeline_mark) or nullptr
  long long* cta_trace;  // optional [n_ctas][16] (development aid): smid, t_entry, t_setup_done, t_first_s, t_last_loop_end, t_exit,
                         // flag-wait clk, #segments, (loop end, epilogue end) of the first 3 segments, globaltimer entry / exit
  lx_attn_desc_t d;
  float scale_log2;  // scale * log2(e)
  float bias_log2;   // cross_bias * log2(e)
  int groups;        // query tiles per unit (2, or 1 when the tile counts are odd)
  // ---- work list: unit = (batch, head, `groups` consecutive query tiles); its work = the visible KV tiles, one
  //      iteration each.  Units are linearised (batch, head, unit) -> x in [0, total) counts KV iterations; CTA c owns
  //      the contiguous range [lo(c), lo(c+1)).  split = 1: lo(c) = c*total/n_cta (balanced to one iteration; a unit cut
  //      by a boundary is finished by the CTA that holds its first iteration, the others publish partials through the
  //      workspace); split = 0: boundaries fall on unit boundaries.
  int n_cta, split;
  int units_rest, units_cond;  // units per head on the two sides of the rest | cond boundary
  int it_rest, it_cond;        // KV iterations of a unit on each side
  int kvb_cond;                // first KV tile visible to condition queries
  int w_head;                  // iterations per (batch, head)
  long long total;             // B * H * w_head
  long long total_units;       // B * H * (units_rest + units_cond)
  int l2_prefetch;             // K / V tiles are prefetched into the L2 this many KV iterations ahead of their TMA load
  int* flags;                  // [n_cta][2]   (workspace; all zero between launches)
  float* slots;                // [n_cta][2][ATT_SLOT_FLOATS]
};

struct AttnSeg {
  int hh;      // batch * H + head
  int qt0;     // first query tile of the unit
  int kv0;     // first KV tile of this segment
  int n;       // KV iterations in this segment
  bool first;  // the segment starts at the unit's first iteration (this CTA finishes the unit)
  bool last;   // the segment reaches the unit's last iteration
  bool q_is_cond;
  long long x_end;  // linear position of the unit's end
};

__device__ __forceinline__ long long attn_lo(const AttnParams& p, int c) {
  if (c >= p.n_cta) return p.total;
  if (p.split) return (long long)c * p.total / p.n_cta;
  const long long U = (long long)c * p.total_units / p.n_cta;  // first unit of CTA c
  const int uph = p.units_rest + p.units_cond;
  const long long hh = U / uph;
  const int u = (int)(U - hh * uph);
  return hh * p.w_head + (u < p.units_rest ? u * p.it_rest : p.units_rest * p.it_rest + (u - p.units_rest) * p.it_cond);
}
// decode the segment that starts at linear position x and ends at the unit's end or at `hi`, whichever is first
__device__ __forceinline__ AttnSeg attn_seg(const AttnParams& p, long long x, long long hi) {
  AttnSeg s;
  const long long hh = x / p.w_head;
  const int rem = (int)(x - hh * p.w_head);
  const int w_rest = p.units_rest * p.it_rest;
  int u, it, n_it, kvb;
  if (rem < w_rest) {
    u = rem / p.it_rest; it = rem - u * p.it_rest; n_it = p.it_rest; kvb = 0;
    s.q_is_cond = false;
  } else {
    const int r2 = rem - w_rest;
    const int uc = r2 / p.it_cond;
    it = r2 - uc * p.it_cond; u = p.units_rest + uc; n_it = p.it_cond; kvb = p.kvb_cond;
    s.q_is_cond = true;
  }
  s.hh = (int)hh;
  s.qt0 = u * p.groups;
  s.kv0 = kvb + it;
  s.x_end = x + (n_it - it);
  s.n = (int)(min(s.x_end, hi) - x);
  s.first = it == 0;
  s.last = s.x_end <= hi;
  return s;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// kPad: ragged streams (padding keys at the end of a stream's last tile are masked); a separate instantiation so that
// the aligned case pays nothing for it.
template <bool kPad>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ AttnParams p) {
  const long long t_entry = clock64();
  if (threadIdx.x == 0) timeline_mark(p.tl, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 groups]
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;                // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;    // [stages]
  uint8_t* sO = sV + ATT_KV_STAGES * ATT_TILE_BYTES;    // output staging: one 128 x 64 atom per group
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + ATT_TILE_BYTES);
  uint64_t* q_full = bars;          // 1      Q tiles of the current segment landed
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* k_empty = bars + 5;     // [2] released by the commit after the last group's QK^T(j)
  uint64_t* v_empty = bars + 7;     // [2] released by the commit after the last group's PV(j)
  uint64_t* s_full = bars + 9;      // [2 groups] S_g(j) complete (and with it every earlier MMA, incl. PV_g(j-1))
  uint64_t* p_full = bars + 11;     // [2 groups] P_g(j) packed into TMEM by all 128 rows
  uint64_t* o_done = bars + 13;     // [2 groups] last PV_g of a segment complete
  uint64_t* p_half = bars + 15;     // [2 groups] first 64 key columns of P_g(j) packed (PV k-slices 0..3 may start)
  uint64_t* o_free = bars + 17;     // [2 groups] O_g of the finished segment has been read out of TMEM (128 arrivals)
  uint64_t* q_free = bars + 19;     // 1      every QK^T of the current segment has completed: Q may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const lx_attn_desc_t& d = p.d;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = p.groups;
  const int cta = blockIdx.x;
  const int n_tiles = d.S / ATT_BKV;
  const int n_rest = (d.S - d.n_cond) / ATT_BKV;  // tiles of the non-condition part
  const bool use_bias = p.bias_log2 != 0.0f;
  const long long x_lo = attn_lo(p, cta), x_hi = attn_lo(p, cta + 1);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmO);
    mbar_init(q_full, 1);
    mbar_init(q_free, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&p_half[s], 128);
      mbar_init(&o_done[s], 1);
      mbar_init(&o_free[s], 128);
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
    trace = p.cta_trace + (long long)cta * 16;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[0] = smid;
    trace[1] = t_entry;
    trace[2] = clock64();
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[14] = gt;
  }
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // PDL: the prologue above overlapped the QKV GEMM's tail; Q / K / V are read from here on
  pdl_launch_dependents();
  if (threadIdx.x == 0) timeline_mark(p.tl, 1);
  // columns [128g, 128g+128): S_g fp32, its first 64 columns re-used for the packed bf16 P_g;  [256+128g, +128): O_g
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp < 4) {
  reg_dealloc<96>();  // setmaxnreg: hand the producer warpgroup's registers to the softmax warpgroups
  if (warp == 0) {
    if (elect_one()) {
      int j = 0, si = 0;  // global KV iteration / segment counters of this CTA
      for (long long x = x_lo; x < x_hi; ++si) {
        const AttnSeg sg = attn_seg(p, x, x_hi);
        x += sg.n;
        const int head_row0 = sg.hh * d.S;  // first row of this head in the [(B*H*S), 128] view
        if (si > 0) mbar_wait(q_free, (si - 1) & 1);
        mbar_expect_tx(q_full, G * ATT_TILE_BYTES);
        for (int g = 0; g < G; ++g) {
          tma_load_2d(sQ + g * ATT_TILE_BYTES, &tmQ, q_full, 0, head_row0 + (sg.qt0 + g) * ATT_BQ);
          tma_load_2d(sQ + g * ATT_TILE_BYTES + ATT_ATOM_BYTES, &tmQ, q_full, 64, head_row0 + (sg.qt0 + g) * ATT_BQ);
        }
        const int pf = p.l2_prefetch;
        for (int it = 0; it < min(pf, sg.n); ++it) {  // the segment's first tiles
          const int row = head_row0 + (sg.kv0 + it) * ATT_BKV;
          tma_prefetch_2d(&tmK, 0, row);
          tma_prefetch_2d(&tmK, 64, row);
          tma_prefetch_2d(&tmV, 0, row);
          tma_prefetch_2d(&tmV, 64, row);
        }
        for (int it = 0; it < sg.n; ++it, ++j) {
          const int st = j & 1;
          const uint32_t par = ((j >> 1) & 1) ^ 1;
          const int row = head_row0 + (sg.kv0 + it) * ATT_BKV;
          if (pf > 0 && it + pf < sg.n) {
            const int prow = row + pf * ATT_BKV;
            tma_prefetch_2d(&tmK, 0, prow);
            tma_prefetch_2d(&tmK, 64, prow);
            tma_prefetch_2d(&tmV, 0, prow);
            tma_prefetch_2d(&tmV, 64, prow);
          }
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
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one() && x_lo < x_hi) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);  // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);   // A = P (TMEM),    B = V (MN-major)
      // Base descriptors are built once; per MMA only the 14-bit address field (bytes >> 4) is advanced, so the issue
      // loop is UIADD + UTCHMMA (the tensor pipe needs a new 128x128x16 MMA every 64 clk).
      const uint64_t q_desc0 = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t k_desc0 = make_sdesc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_desc0 = make_sdesc_sw128(smem_u32(sV), ATT_ATOM_BYTES, 1024);
      // S_g(j) = Q_g K_j^T.  seg_last_qk: j is the last iteration of its segment -> after the last group's QK^T the Q
      // tiles are dead (q_free lets the producer fetch the next segment's Q)
      auto issue_qk = [&](int g, int j, bool seg_last_qk) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
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
        if (g == G - 1) {
          umma_commit(&k_empty[st]);
          if (seg_last_qk) umma_commit(q_free);
        }
      };
      int j = 0, si = 0;
      long long x = x_lo;
      AttnSeg sg = attn_seg(p, x, x_hi);
      x += sg.n;
      mbar_wait(q_full, 0);
      for (int g = 0; g < G; ++g) issue_qk(g, 0, sg.n == 1);
      for (;;) {
        const bool has_next = x < x_hi;
        AttnSeg nx = sg;
        if (has_next) nx = attn_seg(p, x, x_hi);
        for (int it = 0; it < sg.n; ++it, ++j) {
          const int st = j & 1;
          const bool first = it == 0, last = it == sg.n - 1;
          const uint64_t vd = v_desc0 + (uint64_t)(st * (ATT_TILE_BYTES >> 4));
          for (int g = 0; g < G; ++g) {
            const uint32_t to_g = tmem_O + g * 128, tp_g = tmem_S + g * 128;
            mbar_wait(&v_full[st], (j >> 1) & 1);
            // a new segment overwrites O_g: the previous segment's epilogue must have read it out of TMEM
            if (first && si > 0) mbar_wait(&o_free[g], (si - 1) & 1);
            // A: P_g[128 x 16] slice kk = 8 TMEM columns of packed bf16 pairs; B: V rows [16kk, 16kk+16) x 128 d
            // (MN-major: LBO = next 64-d atom).  The first four k-slices start as soon as the first 64 key columns of P
            // are packed, while the softmax threads are still exponentiating the other 64.
            mbar_wait(&p_half[g], j & 1);
            tc_fence_after();
            if (first) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ts(to_g, tp_g + kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_ts(to_g, tp_g + kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, 1u);
            }
            mbar_wait(&p_full[g], j & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 4; kk < 8; ++kk)
              umma_ts(to_g, tp_g + kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, 1u);
            if (g == G - 1) umma_commit(&v_empty[st]);
            if (last) umma_commit(&o_done[g]);
            // S_g(j+1) overwrites the columns P_g(j) is read from: the tensor pipe executes in issue order
            if (!last) {
              issue_qk(g, j + 1, it + 1 == sg.n - 1);
            } else if (has_next) {
              if (g == 0) {
                mbar_wait(q_full, (si + 1) & 1);  // the next segment's Q tiles
                tc_fence_after();
              }
              issue_qk(g, j + 1, nx.n == 1);
            }
          }
        }
        if (!has_next) break;
        sg = nx;
        x += sg.n;
        ++si;
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
      int j = 0, si = 0;
      for (long long x = x_lo; x < x_hi; ++si) {
        const AttnSeg sg = attn_seg(p, x, x_hi);
        x += sg.n;
        float m_run = -INFINITY;  // running (possibly stale) max in log2 units
        float l_run = 0.f;
        for (int it = 0; it < sg.n; ++it, ++j) {
          const bool cross = use_bias && (sg.q_is_cond != ((sg.kv0 + it) >= n_rest));
          const float bias = cross ? p.bias_log2 : 0.f;
          mbar_wait(&s_full[g], j & 1);
          tc_fence_after();
          if (trace != nullptr && j == 0) trace[3] = clock64();
          uint32_t v[128];
          {
            uint32_t(&v0)[64] = *reinterpret_cast<uint32_t(*)[64]>(&v[0]);
            uint32_t(&v1)[64] = *reinterpret_cast<uint32_t(*)[64]>(&v[64]);
            tmem_ld_32x32b_x64(ts, v0);
            tmem_ld_32x32b_x64(ts + 64, v1);
          }
          if (kPad) {  // ragged streams: the last key tile of a stream may end in padding tokens -> -inf logits
            const int kend = (sg.kv0 + it + 1) * ATT_BKV;
            int nvalid = ATT_BKV;
#pragma unroll
            for (int s3 = 0; s3 < 3; ++s3)
              if (kend == d.stream_end[s3]) nvalid -= d.pad[s3];
            if (nvalid < ATT_BKV) {
#pragma unroll
              for (int jj = 0; jj < 128; ++jj)
                if (jj >= nvalid) v[jj] = 0xff800000u;
            }
          }
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int jj = 0; jj < 128; jj += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(v[jj]));
            mx1 = fmaxf(mx1, __uint_as_float(v[jj + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(v[jj + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(v[jj + 3]));
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
          // s_full(j) also covers PV_g(j-1): O_g is stable here
          if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(to + c * 32, o);
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
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
            for (int jj = 0; jj < 16; ++jj) {
              // packed fp32x2 FMA / ADD halve the issue slots; ATT_POLY_OF_8 of every 8 pairs are exponentiated by a
              // polynomial on the FMA pipe, the rest by the MUFU
              float2 xx = __ffma2_rn(make_float2(__uint_as_float(v[c * 32 + 2 * jj]), __uint_as_float(v[c * 32 + 2 * jj + 1])),
                                     scale2, moff2);
              float2 e;
              if ((jj & 7) < ATT_POLY_OF_8) {
                e = ex2_poly2(xx);
              } else {
                e.x = ex2_approx_ordered(xx.x);
                e.y = ex2_approx_ordered(xx.y);
              }
              ls = __fadd2_rn(ls, e);
              pk[jj] = pack_bf16(e.x, e.y);
            }
            tmem_st_32x32b_x16(ts + c * 16, pk);
            if (c == 1) {
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(&p_half[g]);
            }
          }
          l_run = l_run * alpha + (ls.x + ls.y);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[g]);
        }
        // ---------------------------------------------------------------- end of segment
        if (trace != nullptr) {
          if (x >= x_hi) trace[4] = clock64();
          if (si < 3) trace[8 + 2 * si] = clock64();
          trace[7] = si + 1;
        }
        mbar_wait(&o_done[g], si & 1);
        tc_fence_after();
        if (!sg.first) {
          // The unit's first iterations belong to an earlier CTA: publish (m, l, un-normalised O) for it.  Layout
          // [chunk][16-byte fragment][row]: a warp's 32 rows write 512 contiguous bytes per instruction.
          float* slot = p.slots + ((size_t)cta * 2 + g) * ATT_SLOT_FLOATS;
          float4* so = reinterpret_cast<float4*>(slot);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(to + c * 32, o);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              so[(c * 8 + q4) * 128 + r] = make_float4(__uint_as_float(o[4 * q4]), __uint_as_float(o[4 * q4 + 1]),
                                                      __uint_as_float(o[4 * q4 + 2]), __uint_as_float(o[4 * q4 + 3]));
          }
          tc_fence_before();
          mbar_arrive(&o_free[g]);
          reinterpret_cast<float2*>(slot + 128 * 128)[r] = make_float2(m_run, l_run);
          // one gpu-scope fence for the group: the barrier orders every row's stores before thread 0's fence
          // (cumulative), the release store of the flag after it; the other 127 threads go straight on
          asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
          if (r == 0) {
            __threadfence();
            st_release_gpu(p.flags + cta * 2 + g, 1);
          }
        } else {
          // This CTA finishes the unit.  If the unit continues in later CTAs' ranges, fold their partials in.
          float a_self = 1.0f;
          int c_end = cta + 1;
          if (!sg.last) {
            while (c_end < p.n_cta && attn_lo(p, c_end) < sg.x_end) ++c_end;
            float m_all = m_run;
            for (int cc = cta + 1; cc < c_end; ++cc) {
              const int* fl = p.flags + cc * 2 + g;
              if (ld_acquire_gpu(fl) == 0) {
                const long long t0 = clock64();
                if (trace != nullptr) trace[6] -= t0;
                while (ld_acquire_gpu(fl) == 0) {
                  __nanosleep(200);
                  if (clock64() - t0 > 8000000000LL) {
                    printf("lx: attention partial of CTA %d never arrived (CTA %d)\n", cc, cta);
                    __trap();
                  }
                }
                if (trace != nullptr) trace[6] += clock64();
              }
              const float2 ml = __ldcg(reinterpret_cast<const float2*>(p.slots + ((size_t)cc * 2 + g) * ATT_SLOT_FLOATS + 128 * 128) + r);
              m_all = fmaxf(m_all, ml.x);
            }
            a_self = ex2_approx(m_run - m_all);
            l_run *= a_self;
            for (int cc = cta + 1; cc < c_end; ++cc) {
              const float2 ml = __ldcg(reinterpret_cast<const float2*>(p.slots + ((size_t)cc * 2 + g) * ATT_SLOT_FLOATS + 128 * 128) + r);
              l_run += ml.y * ex2_approx(ml.x - m_all);
            }
            m_run = m_all;
          }
          const float inv_l = 1.0f / l_run;
          const int head_row0 = sg.hh * d.S;
          if (d.lse != nullptr) d.lse[head_row0 + (sg.qt0 + g) * ATT_BQ + r] = m_run + log2f(l_run);  // for lx_attention_bwd
          // O_g / l -> bf16 -> this group's 128 x 64 staging atom (128-byte-swizzled TMA layout) -> one TMA store per
          // 64-column half: full-line writes, and the two groups never wait for each other's stores.
          uint8_t* sOg = sO + g * ATT_ATOM_BYTES;
          const int bb_ = sg.hh / d.H, hd = sg.hh - bb_ * d.H;
          const int out_row0 = d.out_row_base[bb_ * n_tiles + sg.qt0 + g];
          const int col0 = d.col_offset + hd * ATT_D;
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            uint4 u[8];
#pragma unroll
            for (int cc2 = 0; cc2 < 2; ++cc2) {
              const int c = hf * 2 + cc2;
              uint32_t o[32];
              tmem_ld_32x32b_x32(to + c * 32, o);
              float f[32];
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) f[jj] = __uint_as_float(o[jj]) * a_self;
              for (int cc = cta + 1; cc < c_end; ++cc) {
                const float* slot = p.slots + ((size_t)cc * 2 + g) * ATT_SLOT_FLOATS;
                const float bb = ex2_approx(__ldcg(reinterpret_cast<const float2*>(slot + 128 * 128) + r).x - m_run);
                const float4* so = reinterpret_cast<const float4*>(slot);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                  const float4 t = __ldcg(so + (c * 8 + q4) * 128 + r);
                  f[4 * q4 + 0] += t.x * bb;
                  f[4 * q4 + 1] += t.y * bb;
                  f[4 * q4 + 2] += t.z * bb;
                  f[4 * q4 + 3] += t.w * bb;
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                u[cc2 * 4 + q].x = pack_bf16(f[8 * q + 0] * inv_l, f[8 * q + 1] * inv_l);
                u[cc2 * 4 + q].y = pack_bf16(f[8 * q + 2] * inv_l, f[8 * q + 3] * inv_l);
                u[cc2 * 4 + q].z = pack_bf16(f[8 * q + 4] * inv_l, f[8 * q + 5] * inv_l);
                u[cc2 * 4 + q].w = pack_bf16(f[8 * q + 6] * inv_l, f[8 * q + 7] * inv_l);
              }
            }
            if (hf == 1) {  // O_g is in registers: the next segment's first PV may overwrite it
              tc_fence_before();
              mbar_arrive(&o_free[g]);
            }
            // the previous store out of this atom (first half / previous segment) must have finished reading it
            if (r == 0) tma_store_wait_read<0>();
            asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
#pragma unroll
            for (int q = 0; q < 8; ++q)  // q = 16-byte chunk of the 128-byte half row
              *reinterpret_cast<uint4*>(sOg + r * 128 + ((q ^ (r & 7)) << 4)) = u[q];
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
            if (r == 0) {
              tma_store_2d(&tmO, sOg, col0 + 64 * hf, out_row0);
              tma_store_commit();
            }
          }
          if (r == 0)
            for (int cc = cta + 1; cc < c_end; ++cc) p.flags[cc * 2 + g] = 0;  // consumed: back to the idle state
        }
        if (trace != nullptr && si < 3) trace[9 + 2 * si] = clock64();
      }
      if (r == 0) tma_store_wait_read<0>();  // the staging atom must outlive the last bulk read
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) timeline_mark(p.tl, 2);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (trace != nullptr) {
    trace[5] = clock64();
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[15] = gt;
  }
}

}  // namespace lx

static long long* g_attn_trace = nullptr;
static int g_attn_force_ctas = 0;
static int g_attn_split = 1;
static int g_attn_l2_prefetch = -1;  // -1: read LX_ATT_PF (default 0) on first use
extern "C" void lx_attention_debug_cta_trace(long long* device_buffer) { g_attn_trace = device_buffer; }
// development / test aid: pretend the device has n SMs (0 = automatic), so that small shapes exercise the split-work path
extern "C" void lx_debug_attention_ctas(int n) { g_attn_force_ctas = n; }
// development aid: 0 = never cut units between CTAs (A/B timing of the split-work schedule)
extern "C" void lx_debug_attention_split(int on) { g_attn_split = on; }

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
  memset(&p, 0, sizeof(p));
  p.d = d;
  p.cta_trace = g_attn_trace;
  const float log2e = 1.4426950408889634f;
  p.scale_log2 = d.scale * log2e;
  p.bias_log2 = d.cross_bias * log2e;
  const int n_tiles = d.S / 128, n_rest = (d.S - d.n_cond) / 128;
  const int q_tiles = d.q_tiles > 0 ? d.q_tiles : n_tiles;
  LX_CHECK_ARG(q_tiles <= n_tiles, "lx_attention: q_tiles=%d exceeds S/128=%d", q_tiles, n_tiles);
  // both query tiles of a unit lie on the same side of the rest | cond boundary
  p.groups = (n_tiles % 2 == 0 && n_rest % 2 == 0 && q_tiles % 2 == 0) ? 2 : 1;
  const int G = p.groups;
  const bool masked = d.cross_bias == 0.f && d.n_cond > 0;
  const int kv_end_rest = (masked && d.mask_mode == 1) ? n_rest : n_tiles;
  p.kvb_cond = (masked && (d.mask_mode == 1 || d.mask_mode == 2)) ? n_rest : 0;
  p.units_rest = (q_tiles < n_rest ? q_tiles : n_rest) / G;
  p.units_cond = (q_tiles > n_rest ? q_tiles - n_rest : 0) / G;
  p.it_rest = kv_end_rest;
  p.it_cond = n_tiles - p.kvb_cond;
  p.w_head = p.units_rest * p.it_rest + p.units_cond * p.it_cond;
  p.total = (long long)d.B * d.H * p.w_head;
  p.total_units = (long long)d.B * d.H * (p.units_rest + p.units_cond);
  const int sms = g_attn_force_ctas > 0 ? g_attn_force_ctas : num_sms();
  p.n_cta = (int)(p.total_units < sms ? p.total_units : sms);
  // split the work to one KV iteration when there are more units than SMs (else: one unit per CTA is already balanced)
  // and the caller registered an exchange workspace for this stream (lx_set_workspace)
  p.split = 0;
  // batch invariance: with a CTA count that is a multiple of B every batch element's work is cut at the same relative
  // positions (lo(c + k n/B) = k W + lo(c)), so identical edits of one batch give bit-identical rows
  const int n_split = sms - sms % d.B;
  if (p.total_units > p.n_cta && g_attn_split && n_split > 0 && p.total_units > n_split) {
    const size_t need = ATT_WS_FLAG_BYTES + (size_t)n_split * 2 * ATT_SLOT_FLOATS * sizeof(float);
    char* ws = static_cast<char*>(workspace_region(stream, 0, need));
    if (ws != nullptr && n_split * 2 * sizeof(int) <= (size_t)ATT_WS_FLAG_BYTES) {
      p.split = 1;
      p.n_cta = n_split;
      p.flags = reinterpret_cast<int*>(ws);
      p.slots = reinterpret_cast<float*>(ws + ATT_WS_FLAG_BYTES);
    }
  }
  if (g_attn_l2_prefetch < 0) {
    const char* e = getenv("LX_ATT_PF");
    g_attn_l2_prefetch = e ? atoi(e) : 0;
  }
  p.l2_prefetch = g_attn_l2_prefetch;
  const bool has_pad = (d.pad[0] | d.pad[1] | d.pad[2]) != 0;
  static bool attr_set = false;
  if (!attr_set) {
    LX_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    LX_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid(p.n_cta);
  double pairs = (double)d.S * d.S;  // visible (query, key) pairs per head
  if (masked) {
    const double nc = d.n_cond, nr = d.S - d.n_cond;
    if (d.mask_mode == 1) pairs = nc * nc + nr * nr;
    if (d.mask_mode == 2) pairs = nc * nc + nr * (double)d.S;
  }
  if (debug_skip_mask() & 2) return LX_OK;  // timing experiments only (lx_debug_skip)
  p.tl = timeline_next(KC_ATTENTION);
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
