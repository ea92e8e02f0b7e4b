// Backward of the joint [txt | img | cond] attention (block.py:129-131) for sm_100a, head_dim 128: dQ, dK, dV from
// Q, K, V, dO and the forward's per-row log-sum-exp, all five GEMMs of the FlashAttention backward on tcgen05.
//
//   one CTA = one 128-key tile j of one (batch, head); it walks the query tiles i that see those keys.
//   Everything is computed TRANSPOSED (keys on the TMEM lanes), so that the probabilities never leave tensor memory
//   for the two products that accumulate over query tiles:
//       S^T  = K_j Q_i^T            SS   -> TMEM cols [0,128)     fp32
//       dP^T = V_j dO_i^T           SS   -> TMEM cols [128,256)   fp32
//       P^T  = exp2(S^T * scale*log2e + bias - lse_i)            (compute threads, one key row each; lse in log2 units)
//       dS^T = scale * P^T * (dP^T - delta_i)                    delta_i = rowsum(dO_i * O_i)
//              P^T, dS^T are packed to bf16 in place over TMEM cols [0,64) / [64,128); dS^T also goes to shared memory
//       dV_j += P^T  dO_i           TS   (A from TMEM, B = dO_i as MN-major smem operand)   -> TMEM cols [256,384)
//       dK_j += dS^T Q_i            TS   (B = Q_i MN-major)                                 -> TMEM cols [384,512)
//       dQ_i  = dS   K_j            SS   (A = the dS^T tile read MN-major, B = K_j MN-major) -> TMEM cols [128,256),
//               issued FIRST of the three, so that its read-back by the compute threads runs under dV / dK; staged in
//               shared memory in one round per iteration (half of it in the then-dead dS^T tile) and added to the fp32 dQ
//               accumulator in global memory by the TMA unit (cp.reduce.async.bulk.tensor .add, 128 x 16 chunks)
//   warp 0: TMA producer (K_j, V_j once; Q_i through a 2-stage ring, dO_i single-buffered: 192 KB of shared memory)
//   warp 1: tcgen05.mma issuer        warps 2..9: compute, two threads per key row (TMEM lane quarter = warp & 3,
//           64 of the 128 query columns each)
//   The tensor pipe executes in issue order, so S^T(i+1) may overwrite the columns P^T(i) / dS^T(i) are read from.
//
// Block masks (block.py:106-120) restrict the query-tile range of a key tile; the log(c_factor) bias (block.py:121-128)
// is uniform per 128x128 tile.
#include "host_util.cuh"
#include "ptx.cuh"
#include "tmem_wide.cuh"

namespace lx {

constexpr int AB_TILE_BYTES = 128 * 128 * 2;
constexpr int AB_ATOM_BYTES = 128 * 64 * 2;
constexpr int AB_THREADS = 320;  // TMA warp, MMA warp, 2 x 4 compute warps
constexpr int AB_DQ_SLOT = 128 * 16 * 4;  // 8 KiB: 128 query rows x 16 head dims fp32; two (double-buffered) per compute group
constexpr int AB_SMEM = 6 * AB_TILE_BYTES + 4 * AB_DQ_SLOT + 2 * 2 * 128 * 4 + 256;  // 226.25 KiB, base 1024-aligned

struct AttnBwdParams {
  const float* lse;    // [B*H*S] log2-domain log-sum-exp of the forward
  const float* delta;  // [B*H*S] rowsum(dO * O)
  float* dq;           // fp32 [B,H,S,128], zero-initialised by the launcher's caller
  __nv_bfloat16* dk;   // bf16 [B,H,S,128]
  __nv_bfloat16* dv;
  int B, H, S, n_cond, mask_mode;
  float scale, scale_log2, bias_log2;
  int stream_end[3], pad[3];  // ragged streams (see lx_attn_desc_t): padding keys get P = dS = 0
  int dbg_flags;  // development aid: bit 0 = skip the dQ reduction (timing experiments only); bit 1 = phase clocks -> prof
  long long* prof;  // [CTAs][4]: clocks the first compute thread spent waiting for S^T / dP^T, in the softmax phase, waiting
                    // for dQ, in the dQ read-back phase (lx_attention_bwd_debug_prof)
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ AttnBwdParams p) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // the 128-byte-swizzle atoms need a 1 KiB-aligned base (no slack left)
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE_BYTES;
  uint8_t* sQ = sV + AB_TILE_BYTES;         // [2 stages]
  uint8_t* sdO = sQ + 2 * AB_TILE_BYTES;
  uint8_t* sdS = sdO + AB_TILE_BYTES;
  uint8_t* sdQ = sdS + AB_TILE_BYTES;       // [2 compute groups] staging of dQ chunks for the TMA reduce
  float* sStat = reinterpret_cast<float*>(sdQ + 4 * AB_DQ_SLOT);  // [2 parities][128] (bias - lse, scale * delta)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 2 * 128);
  uint64_t* kv_full = bars;        // 1
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* do_full = bars + 5;
  uint64_t* do_empty = bars + 6;
  uint64_t* sp_full = bars + 7;    // S^T(i) and dP^T(i) complete
  uint64_t* p_ready = bars + 8;    // 128 arrivals: P^T / dS^T packed (TMEM + smem)
  uint64_t* dq_full = bars + 9;    // dQ_i complete (and dV, dK of this iteration)
  uint64_t* dq_read = bars + 10;   // 128 arrivals: dQ_i read out of TMEM
  uint64_t* acc_done = bars + 11;  // all MMAs of the CTA complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.S / 128;
  const int n_rest = (p.S - p.n_cond) / 128;
  const bool k_is_cond = j >= n_rest;
  const bool use_bias = p.bias_log2 != 0.0f;
  int q_begin = 0, q_end = n_tiles;  // query tiles that see this key tile
  if (!use_bias && p.n_cond > 0) {
    if (p.mask_mode == 1) {
      if (k_is_cond) q_begin = n_rest;
      else q_end = n_rest;
    } else if (p.mask_mode == 2 && !k_is_cond) {
      q_end = n_rest;  // condition queries see condition keys only
    }
  }
  const int n_it = q_end - q_begin;
  const int head_row0 = (b * p.H + h) * p.S;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmdO);
    prefetch_tmap(&tmdQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(sp_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_read, 256);
    mbar_init(acc_done, 1);
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
  const uint32_t tm_S = tmem_base;         // S^T fp32 -> P^T bf16 [0,64) | dS^T bf16 [64,128)
  const uint32_t tm_dP = tmem_base + 128;  // dP^T fp32, then the dQ_i accumulator
  const uint32_t tm_dV = tmem_base + 256;
  const uint32_t tm_dK = tmem_base + 384;

  if (warp == 0) {
    if (elect_one()) {
      const int krow = head_row0 + j * 128;
      mbar_expect_tx(kv_full, 2 * AB_TILE_BYTES);
      tma_load_2d(sK, &tmK, kv_full, 0, krow);
      tma_load_2d(sK + AB_ATOM_BYTES, &tmK, kv_full, 64, krow);
      tma_load_2d(sV, &tmV, kv_full, 0, krow);
      tma_load_2d(sV + AB_ATOM_BYTES, &tmV, kv_full, 64, krow);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const int row = head_row0 + (q_begin + it) * 128;
        mbar_wait(&q_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], AB_TILE_BYTES);
        tma_load_2d(sQ + st * AB_TILE_BYTES, &tmQ, &q_full[st], 0, row);
        tma_load_2d(sQ + st * AB_TILE_BYTES + AB_ATOM_BYTES, &tmQ, &q_full[st], 64, row);
        mbar_wait(do_empty, (it & 1) ^ 1);
        mbar_expect_tx(do_full, AB_TILE_BYTES);
        tma_load_2d(sdO, &tmdO, do_full, 0, row);
        tma_load_2d(sdO + AB_ATOM_BYTES, &tmdO, do_full, 64, row);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one() && n_it > 0) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, false, false);  // A K-major smem, B K-major smem
      constexpr uint32_t idesc_tm = make_idesc_bf16(128, 128, false, true);   // A TMEM, B MN-major smem
      constexpr uint32_t idesc_mm = make_idesc_bf16(128, 128, true, true);    // A MN-major smem, B MN-major smem
      const uint64_t k_kmaj = make_sdesc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t v_kmaj = make_sdesc_sw128(smem_u32(sV), 16, 1024);
      const uint64_t q_kmaj0 = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t do_kmaj = make_sdesc_sw128(smem_u32(sdO), 16, 1024);
      const uint64_t q_mn0 = make_sdesc_sw128(smem_u32(sQ), AB_ATOM_BYTES, 1024);
      const uint64_t do_mn = make_sdesc_sw128(smem_u32(sdO), AB_ATOM_BYTES, 1024);
      const uint64_t ds_mn = make_sdesc_sw128(smem_u32(sdS), AB_ATOM_BYTES, 1024);
      const uint64_t k_mn = make_sdesc_sw128(smem_u32(sK), AB_ATOM_BYTES, 1024);
      auto issue_s = [&](int it) {  // S^T(it) = K_j Q^T
        const int st = it & 1;
        mbar_wait(&q_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t qd = q_kmaj0 + (uint64_t)(st * (AB_TILE_BYTES >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = ((kk >> 2) * AB_ATOM_BYTES + (kk & 3) * 32) >> 4;
          umma_ss(tm_S, k_kmaj + off, qd + off, idesc_kk, kk != 0 ? 1u : 0u);
        }
      };
      auto issue_dp = [&](int it) {  // dP^T(it) = V_j dO^T
        mbar_wait(do_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = ((kk >> 2) * AB_ATOM_BYTES + (kk & 3) * 32) >> 4;
          umma_ss(tm_dP, v_kmaj + off, do_kmaj + off, idesc_kk, kk != 0 ? 1u : 0u);
        }
        umma_commit(sp_full);
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      issue_dp(0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint32_t acc = it > 0 ? 1u : 0u;
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        // dQ_i = dS K_j FIRST: A = dS^T tile [key rows][query cols] read MN-major (M = queries), k-slice = 16 key rows.
        // Its read-back / reduction by the compute threads then runs under the dV / dK products below instead of after
        // them (the tensor pipe executes in issue order: dq_full used to fire only when all three products were done).
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tm_dP, ds_mn + (uint64_t)(kk * (2048 >> 4)), k_mn + (uint64_t)(kk * (2048 >> 4)), idesc_mm,
                  kk != 0 ? 1u : 0u);
        umma_commit(dq_full);
        // dV += P^T dO_i : k-slice kk = 16 queries = 8 packed TMEM columns / 16 rows (2048 B) of the dO tile
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tm_dV, tm_S + kk * 8, do_mn + (uint64_t)(kk * (2048 >> 4)), idesc_tm, kk == 0 ? acc : 1u);
        umma_commit(do_empty);
        // dK += dS^T Q_i
        const uint64_t qm = q_mn0 + (uint64_t)(st * (AB_TILE_BYTES >> 4));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tm_dK, tm_S + 64 + kk * 8, qm + (uint64_t)(kk * (2048 >> 4)), idesc_tm, kk == 0 ? acc : 1u);
        umma_commit(&q_empty[st]);
        if (it + 1 < n_it) {
          issue_s(it + 1);
          mbar_wait(dq_read, it & 1);  // dQ_i has left TMEM cols [128,256)
          tc_fence_after();
          issue_dp(it + 1);
        } else {
          umma_commit(acc_done);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------- compute warps: one key row per thread PAIR (64 queries each)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;  // which 64 of the tile's 128 query columns this thread owns
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;  // 0..255
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int nvalid = 128;  // keys of this tile that are not stream padding
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3)
      if ((j + 1) * 128 == p.stream_end[s3]) nvalid -= p.pad[s3];
    const float keep = r < nvalid ? 1.0f : 0.0f;
    const bool has_pad = nvalid < 128;
    // this thread's per-query statistic: lse (threads 0..127) or delta (128..255) of query tid % 128 of the current tile
    const float* stat_src = (tid < 128 ? p.lse : p.delta) + head_row0 + (tid & 127);
    float stat = n_it > 0 ? stat_src[(size_t)q_begin * 128] : 0.f;
    const bool prof_on = (p.dbg_flags & 2) && p.prof != nullptr && threadIdx.x == 64;
    long long pc[4] = {0, 0, 0, 0}, t_prev = prof_on ? clock64() : 0;
    auto lap = [&](int k) {
      if (prof_on) {
        const long long t = clock64();
        pc[k] += t - t_prev;
        t_prev = t;
      }
    };
    for (int it = 0; it < n_it; ++it) {
      const int qi = q_begin + it;
      const bool cross = use_bias && (k_is_cond != (qi >= n_rest));
      const float bias = cross ? p.bias_log2 : 0.f;
      // per-query constants of this iteration, interleaved so that one 16-byte shared-memory load serves two queries:
      //   .x = bias - lse_q (log2 units)     .y = scale * delta_q
      float2* st_q = reinterpret_cast<float2*>(sStat + (it & 1) * 256);
      if (tid < 128) st_q[tid].x = bias - stat;
      else st_q[tid - 128].y = p.scale * stat;
      // the next iteration's statistic is requested now and consumed a whole iteration later (this global load used to sit
      // at the top of every iteration: one exposed L2 round trip per (key tile, query tile) pair)
      if (it + 1 < n_it) stat = stat_src[(size_t)(qi + 1) * 128];
      lap(3);  // (tail of the previous read-back phase + staging of the statistics)
      mbar_wait(sp_full, it & 1);
      lap(0);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld_32x32b_x64(tm_S + lane_off + half * 64, s);
      // every thread has its S^T values in registers (and the statistics are staged) before anyone overwrites the
      // S^T columns with packed P^T / dS^T; and the bulk reads of the two dQ chunks staged in the dS^T tile are complete
      if (threadIdx.x == 64 || threadIdx.x == 192) tma_store_wait_read<2>();
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;  // 32-query chunk of the tile
        uint32_t dp[32];
        tmem_ld_32x32b_x32(tm_dP + lane_off + c * 32, dp);
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float4 cq = *reinterpret_cast<const float4*>(st_q + c * 32 + 2 * jj);  // (nl0, sd0, nl1, sd1)
          // P = exp2(S scale log2e + bias - lse), dS = scale P (dP - delta) = P (scale dP - scale delta): four
          // FP32 instructions + one MUFU per element.  (Splitting this loop so that the exponentials start on S^T while
          // the tensor pipe is still on dP^T was measured: slower, 0.93 -> 1.07 ms at B = 4.)
          float p0 = ex2_approx(__uint_as_float(s[cc * 32 + 2 * jj]) * p.scale_log2 + cq.x);
          float p1 = ex2_approx(__uint_as_float(s[cc * 32 + 2 * jj + 1]) * p.scale_log2 + cq.z);
          if (has_pad) {  // (uniform over the CTA) padding keys of a ragged stream: P = dS = 0
            p0 *= keep;
            p1 *= keep;
          }
          const float d0 = p0 * (__uint_as_float(dp[2 * jj]) * p.scale - cq.y);
          const float d1 = p1 * (__uint_as_float(dp[2 * jj + 1]) * p.scale - cq.w);
          pk[jj] = pack_bf16(p0, p1);
          dk[jj] = pack_bf16(d0, d1);
        }
        tmem_st_32x32b_x16(tm_S + lane_off + c * 16, pk);
        tmem_st_32x32b_x16(tm_S + lane_off + 64 + c * 16, dk);
        // dS^T row r, query columns [32c, 32c+32) -> four 16-byte chunks of the 128B-swizzled [128 x 128] tile
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int c16 = c * 4 + q4;
          uint4 u = make_uint4(dk[4 * q4], dk[4 * q4 + 1], dk[4 * q4 + 2], dk[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(sdS + (c16 >> 3) * AB_ATOM_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4)) = u;
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      lap(1);
      mbar_wait(dq_full, it & 1);
      lap(2);
      tc_fence_after();
      // dQ_i (lanes = query rows, this group's 64 of the 128 head dims): TMEM -> registers -> four 128 x 16 fp32 staging tiles
      // (64-byte swizzle) -> four cp.reduce.async.bulk.tensor (+=): the L2 does the fp32 adds on whole lines.  ONE staging
      // round per iteration: chunks 0, 1 go into the dS^T tile, which is dead between the dQ product (dq_full) and the next
      // softmax phase, chunks 2, 3 into the group's own two slots; the next softmax phase only waits for the bulk reads of
      // chunks 0, 1 (issued first) before it overwrites the tile.  History: 8 KB at a time through two slots = four rounds
      // of wait / barrier / store / fence / barrier per iteration, 3000 of 5900 clk; the whole half in the dS^T tile = the
      // next softmax phase waits 2100 clk for 32 KB of reductions to drain.
      const bool issuer = (threadIdx.x == 64 + 128 * half);
      uint32_t o[64];
      tmem_ld_32x32b_x64(tm_dP + lane_off + half * 64, o);
      tc_fence_before();
      mbar_arrive(dq_read);  // this thread's part of dQ_i has left tensor memory
      if (issuer) tma_store_wait_read<0>();  // the previous iteration's bulk reads of the staging tiles (a whole iteration ago)
      asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
      if (!(p.dbg_flags & 1)) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint8_t* slot = (t < 2 ? sdS : sdQ) + (half * 2 + (t & 1)) * AB_DQ_SLOT;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint32_t a = static_cast<uint32_t>(r * 64 + ch * 16);
            *reinterpret_cast<uint4*>(slot + (a ^ (((a >> 7) & 3u) << 4))) =  // 64-byte swizzle of the TMA tile
                make_uint4(o[t * 16 + 4 * ch], o[t * 16 + 4 * ch + 1], o[t * 16 + 4 * ch + 2], o[t * 16 + 4 * ch + 3]);
          }
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
      if (issuer && !(p.dbg_flags & 1)) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // one bulk group per chunk: the softmax phase waits for the first two only
          tma_reduce_add_2d(&tmdQ, (t < 2 ? sdS : sdQ) + (half * 2 + (t & 1)) * AB_DQ_SLOT, half * 64 + t * 16,
                            head_row0 + qi * 128);
          tma_store_commit();
        }
      }
    }
    lap(3);
    if (prof_on) {
      long long* dst = p.prof + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 4;
      for (int k = 0; k < 4; ++k) dst[k] = pc[k];
    }
    if (threadIdx.x == 64 || threadIdx.x == 192) tma_store_wait_read<0>();  // staging must outlive the last bulk read
    // epilogue: dV_j, dK_j (lanes = key rows, this thread's 64 head dims) -> bf16 rows
    const size_t out_row = ((size_t)(head_row0 + j * 128 + r)) * 128 + half * 64;
    if (n_it > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* dst = (which == 0 ? p.dv : p.dk) + out_row;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        if (n_it > 0) tmem_ld_32x32b_x32((which == 0 ? tm_dV : tm_dK) + lane_off + half * 64 + c * 32, o);
        else {
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = 0u;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[8 * q4 + 0]), __uint_as_float(o[8 * q4 + 1]));
          u.y = pack_bf16(__uint_as_float(o[8 * q4 + 2]), __uint_as_float(o[8 * q4 + 3]));
          u.z = pack_bf16(__uint_as_float(o[8 * q4 + 4]), __uint_as_float(o[8 * q4 + 5]));
          u.w = pack_bf16(__uint_as_float(o[8 * q4 + 6]), __uint_as_float(o[8 * q4 + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + q4 * 8) = u;
        }
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

// dO rows [R, ld] + O rows [R, ld] (head h in columns [128h, 128h+128)) -> dO head-major [B,H,S,128] and
// delta[b,h,s] = sum_d dO * O.  One warp per (row, head), 4 elements per lane.
__global__ void __launch_bounds__(128) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ d_out, int64_t ld_do,
                                                            const __nv_bfloat16* __restrict__ out, int64_t ld_o, int rows,
                                                            int heads, const lx_tile_meta_t* __restrict__ tm,
                                                            __nv_bfloat16* __restrict__ do_heads, float* __restrict__ delta,
                                                            int seq_total) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const lx_tile_meta_t m = tm[row >> 7];
  const int seq = m.seq_row + (row & 127);
  const int b = seq / seq_total, s = seq % seq_total;
  for (int h = warp; h < heads; h += 4) {
    const uint2 ug = *reinterpret_cast<const uint2*>(d_out + (size_t)row * ld_do + h * 128 + lane * 4);
    const uint2 uo = *reinterpret_cast<const uint2*>(out + (size_t)row * ld_o + h * 128 + lane * 4);
    const float2 g0 = unpack_bf16(ug.x), g1 = unpack_bf16(ug.y), o0 = unpack_bf16(uo.x), o1 = unpack_bf16(uo.y);
    const float dot = warp_sum(g0.x * o0.x + g0.y * o0.y + g1.x * o1.x + g1.y * o1.y);
    const size_t hrow = ((size_t)b * heads + h) * seq_total + s;
    *reinterpret_cast<uint2*>(do_heads + hrow * 128 + lane * 4) = ug;
    if (lane == 0) delta[hrow] = dot;
  }
}

}  // namespace lx

using namespace lx;

extern "C" int lx_attention_bwd_prep(const void* d_out_rows, int64_t ld_do, const void* out_rows, int64_t ld_o,
                                     int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta, void* d_out_heads,
                                     float* delta, int32_t seq_total, void* stream) {
  LX_CHECK_ARG(d_out_rows && out_rows && tile_meta && d_out_heads && delta && rows > 0 && heads > 0 && seq_total > 0 &&
                   ld_do % 4 == 0 && ld_o % 4 == 0,
               "lx_attention_bwd_prep: bad arguments");
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * heads * 128);
  launch_pdl(attn_bwd_prep_kernel, dim3(rows), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(d_out_rows), ld_do, reinterpret_cast<const __nv_bfloat16*>(out_rows), ld_o, rows,
      heads, tile_meta, reinterpret_cast<__nv_bfloat16*>(d_out_heads), delta, seq_total);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

static int g_attn_bwd_dbg = 0;
extern "C" void lx_attention_bwd_debug_flags(int flags) { g_attn_bwd_dbg = flags; }
static long long* g_attn_bwd_prof = nullptr;
// development aid: device buffer of [B * H * S/128][4] int64 for the phase clocks (flags bit 1); NULL switches it off
extern "C" void lx_attention_bwd_debug_prof(long long* buf) { g_attn_bwd_prof = buf; }

extern "C" int lx_attention_bwd(const lx_attn_bwd_desc_t* desc, void* stream) {
  LX_CHECK_ARG(desc != nullptr, "lx_attention_bwd: null descriptor");
  const lx_attn_bwd_desc_t& d = *desc;
  LX_CHECK_ARG(d.q && d.k && d.v && d.d_out && d.lse && d.delta && d.dq && d.dk && d.dv, "lx_attention_bwd: null pointer");
  LX_CHECK_ARG(d.B > 0 && d.H > 0 && d.S > 0 && d.S % 128 == 0, "lx_attention_bwd: S=%d must be a positive multiple of 128",
               d.S);
  LX_CHECK_ARG(d.n_cond >= 0 && d.n_cond % 128 == 0 && d.n_cond < d.S, "lx_attention_bwd: bad n_cond=%d", d.n_cond);
  LX_CHECK_ARG(d.mask_mode >= 0 && d.mask_mode <= 2, "lx_attention_bwd: bad mask_mode=%d", d.mask_mode);
  LX_CHECK_ARG(d.H <= 65535 && d.B <= 65535, "lx_attention_bwd: grid too large");
  const uint64_t rows = (uint64_t)d.B * d.H * d.S;
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdQ;
  int rc;
  if ((rc = make_tmap_2d_f32(&tmdQ, d.dq, rows, 128, 128, 128, 16))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmQ, d.q, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmK, d.k, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmV, d.v, rows, 128, 128, 128, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmdO, d.d_out, rows, 128, 128, 128, 64))) return rc;
  AttnBwdParams p;
  p.lse = d.lse; p.delta = d.delta; p.dq = d.dq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(d.dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(d.dv);
  p.B = d.B; p.H = d.H; p.S = d.S; p.n_cond = d.n_cond; p.mask_mode = d.mask_mode;
  const float log2e = 1.4426950408889634f;
  p.scale = d.scale;
  p.scale_log2 = d.scale * log2e;
  p.bias_log2 = d.cross_bias * log2e;
  for (int s3 = 0; s3 < 3; ++s3) {
    LX_CHECK_ARG(d.pad[s3] >= 0 && d.pad[s3] < 128, "lx_attention_bwd: bad pad[%d]", s3);
    p.stream_end[s3] = d.pad[s3] ? d.stream_end[s3] : 0;
    p.pad[s3] = d.pad[s3];
  }
  p.dbg_flags = g_attn_bwd_dbg;
  p.prof = g_attn_bwd_prof;
  static bool attr_set = false;
  if (!attr_set) {
    LX_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_set = true;
  }
  double pairs = (double)d.S * d.S;
  if (d.cross_bias == 0.f && d.n_cond > 0) {
    const double nc = d.n_cond, nr = d.S - d.n_cond;
    if (d.mask_mode == 1) pairs = nc * nc + nr * nr;
    if (d.mask_mode == 2) pairs = nc * nc + nr * (double)d.S;
  }
  LaunchScope scope(KC_ATTENTION, stream, 10.0 * d.B * d.H * pairs * 128.0);  // five 2*S*S*128 products
  launch_pdl(attention_bwd_kernel, dim3(d.S / 128, d.H, d.B), dim3(AB_THREADS), AB_SMEM, static_cast<cudaStream_t>(stream), tmQ, tmK, tmV,
                                                                                                             tmdO, tmdQ, p);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
