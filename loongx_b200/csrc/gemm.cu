// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C = epilogue(A[M,K] · W[N,K]^T), bf16 in, fp32 accumulate.
//
//   warp 0      : TMA producer  (A tile 128x64, W tile 256x64 per k-block, 128-byte swizzle, 4-stage mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x256x16, accumulators in TMEM,
//                 two 256-column accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 2..5  : epilogue (tcgen05.ld 32x32b: one thread = one output row) with the fused element-wise tails of
//                 the DiT block: bias, GELU-tanh / SiLU, gate*x+residual, per-head RMSNorm + RoPE + head scatter.
//
// Tiles are scheduled M-fastest so that the CTAs resident at one time share a W panel (L2 reuse of the weights,
// A is small enough to stay L2-resident).
#include <stdlib.h>

#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int GEMM_THREADS = 192;

// N tile variants: 256 is the default; 224 / 192 exist to cut wave-quantisation loss (e.g. M=2560, N=3072 is
// 240 tiles = 1.62 waves of 148 CTAs at BN=256 but 280 tiles = 1.89 waves of *smaller* tiles at BN=224).
template <int BN, int NCTA>
struct TileCfg {
  // NCTA == 2: a CTA pair computes a 256 x BN tile with cta_group::2 MMAs; each CTA stages its own 128 rows of A and
  // half (BN/2 rows) of the W tile, so per-CTA L2->smem traffic and smem operand reads per FLOP drop by a third.
  static constexpr int B_BYTES = (BN / NCTA) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = NCTA == 2 ? (BN == 256 ? 6 : 7) : (BN == 256 ? 4 : 5);
  static constexpr int EPI_STAGE_BYTES = BN * 4 + BN * 2;  // bias fp32 + gate bf16 of one tile's columns
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * EPI_STAGE_BYTES;
  static_assert(B_BYTES % 1024 == 0, "B tile must keep 1024-byte swizzle-atom alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct GemmParams {
  lx_gemm_desc_t d;
  int tiles_m, tiles_n;
  int num_kb[3];
  int tile_begin[3];  // first M tile of group g (unused groups: tiles_m)
  int raster_g;       // M tiles per raster band (see tile_coords); = tiles_m when the whole A operand fits the L2 budget
  // Stream-K head (sk_tiles > 0): the first sk_tiles tiles (raster order) are not given to clusters whole; their
  // sk_tiles * num_kb[0] k-blocks are cut into one contiguous, equally long range per cluster, processed BEFORE the
  // remaining tiles, which form full waves of whole tiles.  A tile cut by a range boundary is finished by the cluster that
  // holds its first k-block; the others publish fp32 partial accumulators through the workspace.
  long long* tl;  // development-aid timeline row (ptx.cuh::timeline_mark) or nullptr
  int sk_tiles;
  int* flags;    // [gridDim.x]  (workspace; all zero between launches)
  float* slots;  // [gridDim.x][GEMM_SLOT_FLOATS]: partial accumulator of one CTA, layout [chunk][16-byte fragment][row]
};

constexpr int GEMM_SLOT_FLOATS = 256 * 128;

__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One unit of work of a cluster: k-blocks [kb0, kb0 + nk) of `tile`.
struct GemmWork {
  int tile, kb0, nk;
  bool publish;  // the tile's first k-blocks belong to an earlier cluster: the accumulator goes to the workspace
  bool fix;      // this cluster holds the tile's first k-blocks but not its last: fold the later clusters' partials in
  int x_end;     // stream-K position of the tile's end (fix only)
};
// Work list of one cluster: its stream-K range first, then whole tiles sk_tiles + cluster_id + i * num_clusters.
struct GemmSched {
  int x, hi, dp_tile;
  __device__ __forceinline__ GemmSched(const GemmParams& p, int cluster_id, int num_clusters) {
    const long long sk_total = (long long)p.sk_tiles * p.num_kb[0];
    x = (int)(cluster_id * sk_total / num_clusters);
    hi = (int)((cluster_id + 1) * sk_total / num_clusters);
    dp_tile = p.sk_tiles + cluster_id;
  }
  __device__ __forceinline__ bool next(const GemmParams& p, int num_clusters, int num_tiles, GemmWork& w) {
    if (x < hi) {
      const int nkb = p.num_kb[0];
      w.tile = x / nkb;
      w.kb0 = x - w.tile * nkb;
      w.nk = min(nkb - w.kb0, hi - x);
      w.publish = w.kb0 > 0;
      w.fix = w.kb0 == 0 && w.nk < nkb;
      w.x_end = (w.tile + 1) * nkb;
      x += w.nk;
      return true;
    }
    if (dp_tile < num_tiles) {
      w.tile = dp_tile;
      w.kb0 = 0;
      w.nk = -1;  // whole tile: the caller looks the group's k-block count up
      w.publish = w.fix = false;
      w.x_end = 0;
      dp_tile += num_clusters;
      return true;
    }
    return false;
  }
};
__device__ __forceinline__ int sk_lo(const GemmParams& p, int cluster, int num_clusters) {
  return (int)(cluster * ((long long)p.sk_tiles * p.num_kb[0]) / num_clusters);
}

// Tile order: bands of `raster_g` M tiles; inside a band M runs fastest, then N.  A wave of CTAs then works on one band's
// rows (A working set = raster_g * BM*NCTA * K * 2 bytes, sized for the L2) while sweeping the N tiles, instead of
// streaming the whole A operand once per couple of N tiles when M is large (B >= 4 per GPU, training).
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& tm_idx, int& tn_idx) {
  const int per_band = p.raster_g * p.tiles_n;
  const int band = tile / per_band;
  const int rem = tile - band * per_band;
  const int m_first = band * p.raster_g;
  const int g_size = min(p.raster_g, p.tiles_m - m_first);  // the last band may be shorter
  tn_idx = rem / g_size;
  tm_idx = m_first + rem - tn_idx * g_size;
}

__device__ __forceinline__ int group_of(const GemmParams& p, int tm) {
  return (tm >= p.tile_begin[1] ? 1 : 0) + (tm >= p.tile_begin[2] ? 1 : 0);
}

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]);
  u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]);
  u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

// acc (+bias) for one 32-column chunk -> x[32]
__device__ __forceinline__ void chunk_bias(const uint32_t (&r)[32], const float* __restrict__ bias, int n, float (&x)[32]) {
  if (bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = b4[j];
      x[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
      x[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
      x[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
      x[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
  }
}

// kTrain: the training-step epilogues (BIAS_GELU_DUAL, MUL_AUX, GATE_RESIDUAL + out2, QKV + qkv_pre) are compiled only into
// this variant; the inference kernels (kTrain = false) are instruction-for-instruction what they were without them (with the
// branches compiled in, the 255-register epilogue scheduled worse and the edit lost 2.3 %, A/B on one box).
template <int BN, int NCTA, bool kTrain>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ GemmParams p) {
  constexpr int STAGES = TileCfg<BN, NCTA>::STAGES;
  constexpr int STAGE_BYTES = TileCfg<BN, NCTA>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES + 256;  // [2][bias fp32 BN | gate bf16 BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const lx_gemm_desc_t& d = p.d;
  if (threadIdx.x == 64) timeline_mark(p.tl, 0);
  // CTA pair bookkeeping (NCTA == 1: rank 0, every CTA is its own "cluster")
  const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / NCTA;
  const int num_clusters = gridDim.x / NCTA;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB0);
    if (p.d.n_groups > 1) prefetch_tmap(&tmB1);
    if (p.d.n_groups > 2) prefetch_tmap(&tmB2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4 * NCTA);  // one arrival per epilogue warp of every CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (NCTA == 2) {
      tmem_alloc_2cta(tmem_slot, 512);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail; its outputs (our A operand / residual) are read below.
  // The TMA producer thread defers its wait: the W tiles of the first pipeline stages do not depend on the previous
  // kernel, so it requests them first (see the producer loop).
  const bool defer_pdl = (warp == 0);
  if (!defer_pdl) pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 64) timeline_mark(p.tl, 1);

  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {  // elect.sync: the compiler keeps descriptors in uniform registers (no per-MMA waterfall loop)
      int stage = 0;
      uint32_t phase = 0;
      bool first_tile = true;
      GemmSched sched(p, cluster_id, num_clusters);
      GemmWork wk;
      while (sched.next(p, num_clusters, num_tiles, wk)) {
        int tmi, tni;
        tile_coords(p, wk.tile, tmi, tni);
        const int tm = tmi * NCTA + rank;  // this CTA's 128-row tile
        const int m0 = tm * BM;
        const int n0 = tni * BN + rank * (BN / NCTA);  // this CTA's slice of the W tile
        const int g = group_of(p, tm);
        const CUtensorMap* tmB = g == 0 ? &tmB0 : (g == 1 ? &tmB1 : &tmB2);
        const int nkb = wk.nk < 0 ? p.num_kb[g] : wk.nk;
        const int kb_first = wk.kb0;
        // first tile: the W halves of the first min(STAGES, nkb) stages are requested BEFORE the PDL wait (weights do
        // not depend on the previous kernel), the A halves right after it
        // (w_dynamic: W is an earlier kernel's output, nothing may be requested before the wait)
        const int n_early = (first_tile && !p.d.w_dynamic) ? min(STAGES, nkb) : 0;
        if (first_tile) {
          for (int kb = 0; kb < n_early; ++kb) {  // stage == kb here, every slot is free
            uint8_t* sa = smem + kb * STAGE_BYTES;
            if (NCTA == 2) {
              if (rank == 0) mbar_expect_tx(&full[kb], 2 * STAGE_BYTES);
              tma_load_2d_2sm(sa + A_BYTES, tmB, mapa_shared(smem_u32(&full[kb]), 0), (kb_first + kb) * BK, n0);
            } else {
              mbar_expect_tx(&full[kb], STAGE_BYTES);
              tma_load_2d(sa + A_BYTES, tmB, &full[kb], (kb_first + kb) * BK, n0);
            }
          }
          pdl_wait();
          first_tile = false;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          const bool early = kb < n_early;  // W already requested and the expected bytes already posted
          const int kc = (kb_first + kb) * BK;
          if (NCTA == 2) {
            // both CTAs' bytes are counted on the LEADER's full barrier (the leader issues the pair's MMAs)
            if (rank == 0 && !early) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
            const uint32_t bar = mapa_shared(smem_u32(&full[stage]), 0);
            tma_load_2d_2sm(sa, &tmA, bar, kc, m0);
            if (!early) tma_load_2d_2sm(sa + A_BYTES, tmB, bar, kc, n0);
          } else {
            if (!early) mbar_expect_tx(&full[stage], STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full[stage], kc, m0);
            if (!early) tma_load_2d(sa + A_BYTES, tmB, &full[stage], kc, n0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
    pdl_wait();  // lanes that did not run the loop (and a CTA without tiles) still order themselves after the predecessor
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * NCTA, BN, false, false);
      const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(smem), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      GemmSched sched(p, cluster_id, num_clusters);
      GemmWork wk;
      while (sched.next(p, num_clusters, num_tiles, wk)) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        int tmi, tni;
        tile_coords(p, wk.tile, tmi, tni);
        const int nkb = wk.nk < 0 ? p.num_kb[group_of(p, tmi * NCTA)] : wk.nk;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          // base descriptors are built once; only the 14-bit address field (bytes >> 4) advances per stage / k-slice
          const uint64_t da = a_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
          const uint64_t db = da + (uint64_t)(A_BYTES >> 4);
          if (NCTA == 2) {
            if (kb == 0) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
            }
            umma_commit_2cta(&empty[stage], 3);  // frees the stage in both CTAs
          } else {
            if (kb == 0) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, 1u);
            }
            umma_commit(&empty[stage]);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (NCTA == 2) umma_commit_2cta(&tfull[acc], 3);
        else umma_commit(&tfull[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    GemmSched sched(p, cluster_id, num_clusters);
    GemmWork wk;
    while (sched.next(p, num_clusters, num_tiles, wk)) {
      if (wk.publish) {
        // stream-K: this cluster computed k-blocks [kb0, kb0 + nk) of a tile that an earlier cluster finishes.  The
        // fp32 partial goes to this CTA's workspace slot, [chunk][16-byte fragment][row]: a warp's 32 rows write 512
        // contiguous bytes per instruction; then the flag is raised (release) for the finishing CTA.
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr_p = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
        float4* slot = reinterpret_cast<float4*>(p.slots + (size_t)blockIdx.x * GEMM_SLOT_FLOATS);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr_p + c * 32, r);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            slot[(c * 8 + q4) * 128 + row_in_tile] = make_float4(__uint_as_float(r[4 * q4]), __uint_as_float(r[4 * q4 + 1]),
                                                                 __uint_as_float(r[4 * q4 + 2]), __uint_as_float(r[4 * q4 + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));
          else mbar_arrive(&tempty[acc]);
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) st_release_gpu_s32(p.flags + blockIdx.x, 1);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      int tmi, tni;
      tile_coords(p, wk.tile, tmi, tni);
      const int tm = tmi * NCTA + rank;
      const int m0 = tm * BM;
      const int n0 = tni * BN;
      const int row = m0 + row_in_tile;
      // stream-K: CTAs [fix_cta0, fix_cta0 + NCTA * fix_n) step NCTA (same rank in the following clusters) hold the
      // partial accumulators of this tile's remaining k-blocks
      int fix_n = 0;
      if (wk.fix) {
        int c_end = cluster_id + 1;
        while (c_end < num_clusters && sk_lo(p, c_end, num_clusters) < wk.x_end) ++c_end;
        fix_n = c_end - (cluster_id + 1);
      }
      const float* fix_slot0 = p.slots + (size_t)(blockIdx.x + NCTA) * GEMM_SLOT_FLOATS;
      // accumulator chunk (32 columns from tile column `col`) of this thread's row: TMEM + the published partials
      auto ld_acc = [&](uint32_t taddr_row, int col, uint32_t (&r)[32]) {
        tmem_ld_32x32b_x32(taddr_row + col, r);
        for (int f = 0; f < fix_n; ++f) {
          const float4* sp = reinterpret_cast<const float4*>(fix_slot0 + (size_t)f * NCTA * GEMM_SLOT_FLOATS);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 t = __ldcg(sp + ((col >> 5) * 8 + q4) * 128 + row_in_tile);
            r[4 * q4 + 0] = __float_as_uint(__uint_as_float(r[4 * q4 + 0]) + t.x);
            r[4 * q4 + 1] = __float_as_uint(__uint_as_float(r[4 * q4 + 1]) + t.y);
            r[4 * q4 + 2] = __float_as_uint(__uint_as_float(r[4 * q4 + 2]) + t.z);
            r[4 * q4 + 3] = __float_as_uint(__uint_as_float(r[4 * q4 + 3]) + t.w);
          }
        }
      };
      const bool row_ok = row < d.M;
      const int si = (n0 >= d.n_split) ? 1 : 0;
      const lx_gemm_segment_t& seg = d.seg[si];
      const int seg_n0 = si ? d.n_split : 0;
      const int mode = seg.mode;
      const float* bias_g = d.group[group_of(p, tm)].bias;
      // Stage the tile's bias (and gate) columns in shared memory before waiting for the accumulator: in the chunk loop
      // they are broadcast shared-memory reads instead of one exposed L2 round trip per 32 columns.  Double-buffered by
      // accumulator stage; the named barrier also keeps a fast warp from overwriting a buffer a slow warp still reads.
      float* s_bias = reinterpret_cast<float*>(epi_stage + acc * TileCfg<BN, NCTA>::EPI_STAGE_BYTES);
      __nv_bfloat16* s_gate = reinterpret_cast<__nv_bfloat16*>(s_bias + BN);
      {
        const int et = threadIdx.x - 64;  // 0..127 over the four epilogue warps
        if (bias_g != nullptr)
          for (int c = et; c < BN; c += 128) s_bias[c] = (n0 + c < d.N) ? __ldg(bias_g + n0 + c) : 0.f;
        if (mode == LX_EPI_GATE_RESIDUAL && m0 < d.M) {  // (the odd CTA of the last pair may own no rows at all)
          const lx_tile_meta_t mg = d.tile_meta[tm];
          const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(d.gate[mg.stream]) +
                                    (size_t)mg.batch * d.gate_stride[mg.stream];
          for (int c = et; c < BN; c += 128) s_gate[c] = (n0 + c < d.N) ? gp[n0 + c] : __float2bfloat16(0.f);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const float* bias = bias_g != nullptr ? s_bias : nullptr;

      // Gate-residual tiles: the residual row segment is fetched BEFORE waiting for the accumulator, so the (L2-latency,
      // one 16-byte fragment per row) loads complete while the tensor pipe is still working on this tile; in the
      // chunk loop below they would be one exposed round trip per 32 columns on the last tile of every CTA.
      uint4 rres[BN / 8];
      if ((mode == LX_EPI_GATE_RESIDUAL || (kTrain && mode == LX_EPI_MUL_AUX)) && row_ok) {
        const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(d.residual) + (size_t)row * d.ldr +
                                   (n0 - seg_n0 + seg.col_offset);
#pragma unroll
        for (int j = 0; j < BN / 8; ++j)
          if (n0 + j * 8 < d.N) rres[j] = *reinterpret_cast<const uint4*>(res + j * 8);
      }
      lx_tile_meta_t meta_pf = {0, 0, 0, 0};
      if (BN == 256 && mode == LX_EPI_QKV && m0 < d.M) meta_pf = d.tile_meta[tm];  // also fetched ahead of the wait
      if (BN == 256 && mode == LX_EPI_QKV && d.rope != nullptr && row_ok && n0 < 2 * d.heads * 128) {
        // same idea for the q / k tiles: this row's 64 (cos, sin) pairs (shared by both heads of the tile) are in
        // registers before the accumulator is ready
        const float4* rope_pf =
            reinterpret_cast<const float4*>(d.rope + (size_t)(meta_pf.seq_row % d.seq_total + row_in_tile) * 128);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 t = __ldg(rope_pf + j);
          rres[j % (BN / 8)] = make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w));
        }
      }

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int f = 0; f < fix_n; ++f) {  // the partials must have been published (acquire; normally long since)
        const int* fl = p.flags + blockIdx.x + NCTA * (f + 1);
        if (ld_acquire_gpu_s32(fl) == 0) {
          const long long t0 = clock64();
          while (ld_acquire_gpu_s32(fl) == 0) {
            __nanosleep(100);
            if (clock64() - t0 > 8000000000LL) {
              printf("lx: gemm partial of CTA %d never arrived (CTA %d)\n", (int)(blockIdx.x + NCTA * (f + 1)), (int)blockIdx.x);
              __trap();
            }
          }
        }
      }

      if (BN == 256 && mode == LX_EPI_QKV) {
        const lx_tile_meta_t meta = meta_pf;
        const int D = d.heads * 128;
        const int sec = n0 / D;  // 0 = q, 1 = k, 2 = v
        const int s_pos = meta.seq_row % d.seq_total + row_in_tile;
        __nv_bfloat16* dst_base = reinterpret_cast<__nv_bfloat16*>(sec == 0 ? d.q : (sec == 1 ? d.k : d.v));
        const float* rmsw = sec == 0 ? d.rms_q[meta.stream] : (sec == 1 ? d.rms_k[meta.stream] : nullptr);
        const float4* rope_row =
            (d.rope != nullptr && sec < 2) ? reinterpret_cast<const float4*>(d.rope + (size_t)s_pos * 128) : nullptr;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int nh = n0 + half * 128;
          if (nh >= d.n_split) break;
          const int head = (nh - sec * D) >> 7;
          __nv_bfloat16* dst = dst_base + (((size_t)meta.batch * d.heads + head) * d.seq_total + s_pos) * 128;
          float inv = 1.0f;
          if (rmsw != nullptr) {
            float ss = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t r[32];
              float x[32];
              ld_acc(taddr, half * 128 + c * 32, r);
                  chunk_bias(r, bias, half * 128 + c * 32, x);
#pragma unroll
              for (int j = 0; j < 32; ++j) ss += x[j] * x[j];
            }
            inv = rsqrtf(ss * (1.0f / 128.0f) + d.rms_eps);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            float x[32];
            ld_acc(taddr, half * 128 + c * 32, r);
              chunk_bias(r, bias, half * 128 + c * 32, x);
            if (kTrain && d.qkv_pre != nullptr && row_ok) {  // training forward: the pre-norm projection survives for the backward
              __nv_bfloat16* pre = reinterpret_cast<__nv_bfloat16*>(d.qkv_pre) + (size_t)row * d.ld_qkv_pre + nh + c * 32;
#pragma unroll
              for (int j = 0; j < 4; ++j) store_bf16x8(pre + j * 8, &x[8 * j]);
            }
            if (rmsw != nullptr) {
              const float4* w4 = reinterpret_cast<const float4*>(rmsw + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 w = __ldg(w4 + j);
                x[4 * j + 0] *= inv * w.x;
                x[4 * j + 1] *= inv * w.y;
                x[4 * j + 2] *= inv * w.z;
                x[4 * j + 3] *= inv * w.w;
              }
            }
            if (rope_row != nullptr && row_ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 cu = rres[(c * 8 + j) % (BN / 8)];  // prefetched (cos0, sin0, cos1, sin1) for two rotary pairs
                const float4 cs = make_float4(__uint_as_float(cu.x), __uint_as_float(cu.y), __uint_as_float(cu.z),
                                              __uint_as_float(cu.w));
                float a0 = x[4 * j + 0], b0 = x[4 * j + 1], a1 = x[4 * j + 2], b1 = x[4 * j + 3];
                x[4 * j + 0] = a0 * cs.x - b0 * cs.y;
                x[4 * j + 1] = b0 * cs.x + a0 * cs.y;
                x[4 * j + 2] = a1 * cs.z - b1 * cs.w;
                x[4 * j + 3] = b1 * cs.z + a1 * cs.w;
              }
            }
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) store_bf16x8(dst + c * 32 + j * 8, &x[8 * j]);
            }
          }
        }
      } else if (mode == LX_EPI_GATE_RESIDUAL || (kTrain && mode == LX_EPI_MUL_AUX)) {
        // MUL_AUX (training, dX through an activation): the "residual" rows hold the activation's derivative, written by
        // the forward's BIAS_GELU_DUAL epilogue -- one multiply per element here; evaluating gelu' in this fully unrolled
        // branch (255 registers, no room for instruction-level parallelism) cost +45 % on the whole GEMM
        const bool mul_aux = kTrain && mode == LX_EPI_MUL_AUX;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(seg.out) + (size_t)row * seg.ldo;
        // GATE_RESIDUAL with out2 (training forward): the pre-gate projection y = acc + bias survives for the gate gradient
        __nv_bfloat16* y_out = (kTrain && !mul_aux && d.out2 != nullptr)
                                   ? reinterpret_cast<__nv_bfloat16*>(d.out2) + (size_t)row * d.ldo2 + (d.col_offset2 - seg_n0)
                                   : nullptr;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int n = n0 + c * 32;
          if (n < d.N) {
            uint32_t r[32];
            float x[32];
            ld_acc(taddr, c * 32, r);
            chunk_bias(r, bias, c * 32, x);
            const int oc = n - seg_n0 + seg.col_offset;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n + j * 8 < d.N) {
                const uint4 g = *reinterpret_cast<const uint4*>(s_gate + c * 32 + j * 8);
                uint32_t gu[4] = {g.x, g.y, g.z, g.w};
                if (row_ok) {
                  if (y_out != nullptr) store_bf16x8(y_out + n + j * 8, &x[8 * j]);
                  const uint4 rr = rres[c * 4 + j];
                  uint32_t ru[4] = {rr.x, rr.y, rr.z, rr.w};
                  float o[8];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float2 gg = unpack_bf16(gu[e]);
                    float2 r2 = unpack_bf16(ru[e]);
                    if (mul_aux) {
                      o[2 * e + 0] = r2.x * x[8 * j + 2 * e + 0];
                      o[2 * e + 1] = r2.y * x[8 * j + 2 * e + 1];
                    } else {
                      o[2 * e + 0] = r2.x + gg.x * x[8 * j + 2 * e + 0];
                      o[2 * e + 1] = r2.y + gg.y * x[8 * j + 2 * e + 1];
                    }
                  }
                  store_bf16x8(out + oc + j * 8, o);
                }
              }
            }
          }
        }
      } else {
        // BIAS / GELU / SILU / F32
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int n = n0 + c * 32;
          if (n >= d.N) break;
          uint32_t r[32];
          float x[32];
          ld_acc(taddr, c * 32, r);
          chunk_bias(r, bias, c * 32, x);
          if (kTrain && mode == LX_EPI_BIAS_GELU_DUAL) {  // training forward: gelu' goes to `out` for the backward's MUL_AUX epilogue
            float gp[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) gelu_tanh_pair(x[j], x[j], gp[j]);
            if (row_ok) {
              __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(seg.out) + (size_t)row * seg.ldo + (n - seg_n0 + seg.col_offset);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j * 8 < d.N) store_bf16x8(out + j * 8, &gp[8 * j]);
            }
          } else if (mode == LX_EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_tanh(x[j]);
          } else if (mode == LX_EPI_BIAS_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = silu(x[j]);
          }
          const int oc = n - seg_n0 + seg.col_offset;
          if (row_ok) {
            if (mode == LX_EPI_BIAS_F32) {
              float* out = reinterpret_cast<float*>(seg.out) + (size_t)row * seg.ldo + oc;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n + j * 4 < d.N)
                  *reinterpret_cast<float4*>(out + j * 4) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            } else {
              __nv_bfloat16* out = (kTrain && mode == LX_EPI_BIAS_GELU_DUAL)
                                       ? reinterpret_cast<__nv_bfloat16*>(d.out2) + (size_t)row * d.ldo2 + (n - seg_n0 + d.col_offset2)
                                       : reinterpret_cast<__nv_bfloat16*>(seg.out) + (size_t)row * seg.ldo + oc;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j * 8 < d.N) store_bf16x8(out + j * 8, &x[8 * j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));  // the leader's barrier
        else mbar_arrive(&tempty[acc]);
      }
      if (fix_n > 0) {  // every epilogue thread has read the partials: the flags return to idle for the next launch
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64)
          for (int f = 0; f < fix_n; ++f) p.flags[blockIdx.x + NCTA * (f + 1)] = 0;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 64) timeline_mark(p.tl, 2);
  if (NCTA == 2) cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_2cta(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lx

namespace lx {

static long long g_stream_k_launches = 0;
// The stream-K head is OFF by default: measured on B200 (round 2, profiles/README.md) the fp32 partial-tile exchange
// through L2 costs more than the partial last wave it removes (M = 2560: N = K = 3072 36.4 -> 41.3 us, N = 3072 /
// K = 12288 136.9 -> 142.7 us, N = 12288 / K = 3072 134.1 -> 135.9 us; whole edit 962 -> 1037 ms).  lx_debug_gemm_stream_k(1)
// switches it on (tests, further tuning).
static int g_stream_k = 0;
static int g_fake_sms = 0;  // lx_debug_gemm_sms(n): pretend the device has n SMs (tests: small shapes exercise stream-K)
static long long g_raster_budget_mb = 40;  // L2 share given to the A rows of one raster band (the 126 MB L2 is two 63 MB
                                           // partitions; 40 MB measured best end to end at B = 4: -7.6 % per denoise step)

template <int BN, int NCTA, bool kTrain>
int launch_gemm_t(const lx_gemm_desc_t& d, void* stream) {
  CUtensorMap tmA, tmB[3];
  GemmParams p;
  p.d = d;
  p.tiles_m = (d.M + BM * NCTA - 1) / (BM * NCTA);  // (pairs of) 128-row tiles
  p.tiles_n = (d.N + BN - 1) / BN;
  {
    int k_all = 0;
    for (int g = 0; g < d.n_groups; ++g) k_all = max(k_all, (int)d.group[g].K);
    const long long band_bytes = (long long)BM * NCTA * k_all * 2;  // A rows of one M tile (pair)
    const long long budget = g_raster_budget_mb * (1LL << 20);
    p.raster_g = (int)max(1LL, min((long long)p.tiles_m, budget / max(band_bytes, 1LL)));
  }
  int kmax = 0;
  for (int g = 0; g < 3; ++g) {
    const int gi = g < d.n_groups ? g : 0;
    p.num_kb[g] = (d.group[gi].K + BK - 1) / BK;
    p.tile_begin[g] = g < d.n_groups ? d.group[g].m_begin / BM : (d.M + BM - 1) / BM + NCTA;
    kmax = max(kmax, d.group[gi].K);
    int rc = make_tmap_2d_bf16(&tmB[g], d.group[gi].W, (uint64_t)d.N, (uint64_t)d.group[gi].K,
                               (uint64_t)d.group[gi].ldw, BN / NCTA, BK);
    if (rc) return rc;
  }
  int rc = make_tmap_2d_bf16(&tmA, d.A, (uint64_t)d.M, (uint64_t)kmax, (uint64_t)d.lda, BM, BK);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    LX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, NCTA, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TileCfg<BN, NCTA>::SMEM));
    attr_set = true;
  }
  const int sms = g_fake_sms > 0 ? g_fake_sms : num_sms();
  const int grid = min(p.tiles_m * p.tiles_n, max(sms / NCTA, 1)) * NCTA;
  // Stream-K head: when the tile count is not a multiple of the cluster count, the `rem` tiles that would form a partial
  // last wave are cut along K into one equal range per cluster and processed first; the rest runs as full waves of
  // whole tiles.  Needs one k-block count for all row groups, enough k-blocks per cluster to amortise the partial-tile
  // exchange (128 KB per CTA through L2), a partial wave that is actually wasteful, and the caller's workspace.
  p.sk_tiles = 0;
  p.flags = nullptr;
  p.slots = nullptr;
  {
    const int clusters = grid / NCTA, tiles = p.tiles_m * p.tiles_n;
    const int rem = tiles % clusters;
    const bool same_k = p.num_kb[0] == p.num_kb[1] && p.num_kb[0] == p.num_kb[2];
    if (g_stream_k && tiles > clusters && rem > 0 && rem * 100 < clusters * 92 && same_k &&
        (long long)rem * p.num_kb[0] >= 8LL * clusters) {
      const size_t need = WS_FLAG_BYTES + (size_t)grid * GEMM_SLOT_FLOATS * sizeof(float);
      char* ws = static_cast<char*>(workspace_region(stream, 1, need));
      if (ws != nullptr && (size_t)grid * sizeof(int) <= (size_t)WS_FLAG_BYTES) {
        p.sk_tiles = rem;
        ++g_stream_k_launches;
        p.flags = reinterpret_cast<int*>(ws);
        p.slots = reinterpret_cast<float*>(ws + WS_FLAG_BYTES);
      }
    }
  }
  double flops = 0;
  for (int g = 0; g < d.n_groups; ++g) {
    const int m_end = g + 1 < d.n_groups ? d.group[g + 1].m_begin : d.M;
    flops += 2.0 * (m_end - d.group[g].m_begin) * (double)d.N * d.group[g].K;
  }
  p.tl = timeline_next(KC_GEMM);
  LaunchScope scope(KC_GEMM, stream, flops);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = TileCfg<BN, NCTA>::SMEM;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  LX_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, NCTA, kTrain>, tmA, tmB[0], tmB[1], tmB[2], p));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

template <int BN, int NCTA>
int launch_gemm(const lx_gemm_desc_t& d, void* stream) {
  bool train = d.out2 != nullptr || d.qkv_pre != nullptr;
  for (int s = 0; s < 2; ++s) train = train || d.seg[s].mode == LX_EPI_BIAS_GELU_DUAL || d.seg[s].mode == LX_EPI_MUL_AUX;
  return train ? launch_gemm_t<BN, NCTA, true>(d, stream) : launch_gemm_t<BN, NCTA, false>(d, stream);
}

// N tile that minimises (number of waves) x (tile width); ties go to the wider tile.  With the stream-K head available
// (`frac_waves`: the partial last wave costs only its share) the widest tile always wins.
int pick_tile_n(int M, int N, bool need_256, int ncta, bool frac_waves) {
  if (need_256 || (frac_waves && N % 256 == 0)) return 256;
  const int units = max((g_fake_sms > 0 ? g_fake_sms : num_sms()) / ncta, 1);
  const int tiles_m = (M + BM * ncta - 1) / (BM * ncta);
  int best = 256;
  long best_cost = -1;
  for (int bn : {256, 224, 192}) {
    const long tiles = (long)tiles_m * ((N + bn - 1) / bn);
    const long cost = ((tiles + units - 1) / units) * bn;
    if (best_cost < 0 || cost < best_cost) {
      best = bn;
      best_cost = cost;
    }
  }
  return best;
}

}  // namespace lx

static int g_force_ncta = 0;
// development knob (scripts/gemm_shapes.py): 1 = never pair CTAs, 0 = automatic
extern "C" void lx_debug_gemm_force_ncta(int ncta) { g_force_ncta = ncta; }
// development aid: L2 budget (MB) for one raster band of A rows; a huge value restores the plain M-fastest order
extern "C" void lx_debug_gemm_raster_budget_mb(int mb) { lx::g_raster_budget_mb = mb; }
extern "C" void lx_debug_gemm_stream_k(int on) { lx::g_stream_k = on; }
extern "C" void lx_debug_gemm_sms(int n) { lx::g_fake_sms = n; }
extern "C" long long lx_debug_gemm_stream_k_launches(void) { return lx::g_stream_k_launches; }

extern "C" int lx_gemm_bf16(const lx_gemm_desc_t* desc, void* stream) {
  using namespace lx;
  LX_CHECK_ARG(desc != nullptr, "lx_gemm_bf16: null descriptor");
  const lx_gemm_desc_t& d = *desc;
  LX_CHECK_ARG(d.M > 0 && d.N > 0 && d.N % 8 == 0, "lx_gemm_bf16: bad shape M=%d N=%d", d.M, d.N);
  LX_CHECK_ARG(d.A != nullptr, "lx_gemm_bf16: null A");
  LX_CHECK_ARG(d.n_groups >= 1 && d.n_groups <= 3, "lx_gemm_bf16: n_groups=%d outside [1,3]", d.n_groups);
  for (int g = 0; g < d.n_groups; ++g) {
    const lx_gemm_group_t& G = d.group[g];
    LX_CHECK_ARG(G.W != nullptr && G.K > 0 && G.K % 8 == 0, "lx_gemm_bf16: group %d needs W and K (multiple of 8), K=%d", g,
                 G.K);
    LX_CHECK_ARG(G.m_begin % 128 == 0 && (g == 0 ? G.m_begin == 0 : G.m_begin >= d.group[g - 1].m_begin),
                 "lx_gemm_bf16: group %d m_begin=%d must be an ascending multiple of 128 (group 0 at 0)", g, G.m_begin);
  }
  LX_CHECK_ARG(d.n_split > 0 && (d.n_split == d.N || d.n_split % 256 == 0) && d.n_split <= d.N,
               "lx_gemm_bf16: n_split=%d must be N or a multiple of 256", d.n_split);
  const int nseg = d.n_split < d.N ? 2 : 1;
  bool need_256 = nseg == 2;
  for (int s = 0; s < nseg; ++s) {
    const lx_gemm_segment_t& g = d.seg[s];
    LX_CHECK_ARG(g.mode >= LX_EPI_BIAS && g.mode <= LX_EPI_MUL_AUX, "lx_gemm_bf16: bad epilogue mode %d", g.mode);
    if (g.mode == LX_EPI_QKV) {
      need_256 = true;
      LX_CHECK_ARG(s == 0, "lx_gemm_bf16: QKV segment must be segment 0");
      LX_CHECK_ARG(d.q && d.k && d.v && d.heads > 0 && d.seq_total > 0 && d.tile_meta,
                   "lx_gemm_bf16: QKV epilogue needs q/k/v, heads, seq_total, tile_meta");
      LX_CHECK_ARG(d.n_split == 3 * d.heads * 128, "lx_gemm_bf16: QKV segment must span 3*heads*128 columns");
      LX_CHECK_ARG(d.seq_total % 128 == 0, "lx_gemm_bf16: seq_total must be a multiple of 128");
      LX_CHECK_ARG(d.heads % 2 == 0, "lx_gemm_bf16: QKV epilogue needs an even head count (256-column tiles)");
      LX_CHECK_ARG(d.M % 128 == 0, "lx_gemm_bf16: QKV epilogue needs M to be a multiple of 128");
      LX_CHECK_ARG(d.qkv_pre == nullptr || (d.ld_qkv_pre >= d.n_split && d.ld_qkv_pre % 8 == 0),
                   "lx_gemm_bf16: qkv_pre needs ld_qkv_pre >= 3*heads*128 (multiple of 8)");
    } else {
      LX_CHECK_ARG(g.out != nullptr && g.ldo > 0 && g.ldo % 8 == 0 && g.col_offset % 8 == 0,
                   "lx_gemm_bf16: segment %d needs out / ldo (multiple of 8)", s);
    }
    if (g.mode == LX_EPI_GATE_RESIDUAL) {
      LX_CHECK_ARG(d.residual && d.tile_meta && d.ldr % 8 == 0, "lx_gemm_bf16: GATE_RESIDUAL needs residual, tile_meta");
    }
    if (g.mode == LX_EPI_MUL_AUX) {
      LX_CHECK_ARG(d.residual && d.ldr % 8 == 0, "lx_gemm_bf16: MUL_AUX needs the factor rows in residual / ldr");
    }
    if (g.mode == LX_EPI_GATE_RESIDUAL && d.out2) {
      LX_CHECK_ARG(d.ldo2 > 0 && d.ldo2 % 8 == 0 && d.col_offset2 % 8 == 0, "lx_gemm_bf16: out2 needs ldo2 (multiple of 8)");
    }
    if (g.mode == LX_EPI_BIAS_GELU_DUAL) {
      LX_CHECK_ARG(d.out2 && d.ldo2 > 0 && d.ldo2 % 8 == 0 && d.col_offset2 % 8 == 0,
                   "lx_gemm_bf16: BIAS_GELU_DUAL needs out2 / ldo2 (multiple of 8)");
    }
  }
  // CTA pairs (256-row tiles) whenever every row group starts on a 256-row boundary; LX_GEMM_NCTA=1 forces single CTAs
  static const int env_ncta = [] {
    const char* e = getenv("LX_GEMM_NCTA");
    return e ? atoi(e) : 0;
  }();
  const int forced_ncta = g_force_ncta ? g_force_ncta : env_ncta;
  bool pair_ok = d.M >= 256 && forced_ncta != 1;
  for (int g = 0; g < d.n_groups; ++g) pair_ok = pair_ok && d.group[g].m_begin % 256 == 0;
  const int ncta = pair_ok ? 2 : 1;
  int bn = d.tile_n;
  if (bn == 0) {
    bool same_k = true;
    for (int g = 1; g < d.n_groups; ++g) same_k = same_k && (d.group[g].K + BK - 1) / BK == (d.group[0].K + BK - 1) / BK;
    const bool frac_waves = g_stream_k && same_k && workspace_region(stream, 1, WS_FLAG_BYTES) != nullptr &&
                            (long long)((d.M + BM * ncta - 1) / (BM * ncta)) * ((d.N + 255) / 256) >
                                (g_fake_sms > 0 ? g_fake_sms : num_sms()) / ncta;
    bn = pick_tile_n(d.M, d.N, need_256, ncta, frac_waves);
  }
  // 128 is never picked automatically (the DiT shapes are tuned on 256 / 224 / 192); callers with N <= 128 outputs (the
  // VAE's 128-channel convolutions) force it to avoid a third of idle tile columns
  LX_CHECK_ARG(bn == 256 || ((bn == 224 || bn == 192 || bn == 128) && !need_256), "lx_gemm_bf16: tile_n=%d not allowed here", bn);
  if (ncta == 2) {
    if (bn == 256) return launch_gemm<256, 2>(d, stream);
    if (bn == 224) return launch_gemm<224, 2>(d, stream);
    if (bn == 128) return launch_gemm<128, 2>(d, stream);
    return launch_gemm<192, 2>(d, stream);
  }
  if (bn == 256) return launch_gemm<256, 1>(d, stream);
  if (bn == 224) return launch_gemm<224, 1>(d, stream);
  if (bn == 128) return launch_gemm<128, 1>(d, stream);
  return launch_gemm<192, 1>(d, stream);
}
