// Grouped weight-gradient GEMM of the training step (BASELINE configs[2]):
//
//   out_g[n, k] (+)= sum_t dY_g[t, n] * X_g[t, k]        g = 0 .. count-1, t over ALL tokens, fp32 accumulation in TMEM
//
// One launch runs the weight gradients of several Linear layers (q|k|v, out_proj, fc1, fc2 of up to three encoder
// layers).  A single wgrad has only 4-16 output tiles of 256 x 256, far fewer than the 74 CTA pairs of a B200, which is
// why gemm.cu's wgrad had to split the token contraction 6-18 ways and write fp32 partials (108 MB per encoder layer,
// reduced by another kernel).  Pooled, the tiles of 12 gradients fill two whole waves with the FULL contraction per
// tile: no partials, no reduction, one long main loop per tile (16000 tokens = 250 k-blocks) instead of a short one
// with an exposed fp32 epilogue.  Both operands are read as stored (MN-major: memory is [token][feature]).
//
// Bias gradients ride along: db_g[n] = sum_t dY_g[t, n] is the same contraction against a vector of ones, so a group
// with a bias gets one extra "bias tile" per 256 output features whose B operand is a constant 16 x 64 tile of ones in
// shared memory (N = 16 UMMA: 1/16 of the tensor work, no W loads); the epilogue writes column 0.  These short tiles
// are scheduled after all regular tiles.  This replaces a column-sum kernel + partial reduction per Linear layer.
//
// Structure = gemm.cu (persistent, warp-specialised, cta_group::2, 256 x 256 tile per CTA pair, 2 x 256 TMEM columns):
//   warp 0 TMA producer, warp 1 MMA issuer (leader), warp 2 TMEM allocator, warps 4-7 epilogue (fp32 slabs -> TMA store;
//   ACC: the current gradient is TMA-loaded into the slab two slabs ahead and updated in place).
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"
#include <algorithm>
#include <utility>
#include <vector>

namespace mm {

constexpr int WG_MAX = MM_WGRAD_MAX_GROUPS;

struct WgMaps {
  CUtensorMap a[WG_MAX], w[WG_MAX], out[WG_MAX];
};
constexpr int WG_ORDER_MAX = 1536;    // entries of the host-made tile schedule (rounds x CTA pairs)
constexpr unsigned short WG_NO_TILE = 0xFFFF;

struct WgDev {
  int count, num_tiles, num_reg_tiles, accumulate;
  int num_kb[WG_MAX];           // 64-token blocks of group g's contraction (groups may differ in their token count)
  int tail_steps[WG_MAX];       // K = 16 steps of the last block
  // Tile schedule.  Pair q runs order[q], order[q + pairs], ... (WG_NO_TILE: nothing this round).  The host fills it by
  // greedy list scheduling over the tiles' costs (long contractions first), so that groups with different token counts
  // -- the encoder layers' 16 000 tokens and the image-side projections' 36 928 -- share ONE balanced launch.
  // sched_len == 0: tile i runs in round-robin position i.
  int sched_len;
  unsigned short order[WG_ORDER_MAX];
  int tile_start[WG_MAX + 1];   // regular tiles: prefix sum over groups
  int bias_start[WG_MAX + 1];   // bias tiles (one per 256 output features of a group with a bias): prefix sum
  int n_tiles[WG_MAX];          // column tiles of group g (0: bias only)
  int n[WG_MAX];                // output columns (k_in) of group g
  int n_out[WG_MAX];            // output rows (features) of group g
  float* bias[WG_MAX];          // bias gradient [n_out] or null
};

struct WgCfg {
  static constexpr int BM = 128, BN = 256, BK = 64, STAGES = 5, NB = 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int SLAB_BYTES = BM * 128;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int ONES_BYTES = 1024;    // 8 rows x 128 B of 16-bit ones: this CTA's half of the N = 16 bias operand
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + NB * SLAB_BYTES + ONES_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ uint4* wg_slab_chunk(uint8_t* slab, int row, int c) {
  return reinterpret_cast<uint4*>(slab + row * 128 + ((c ^ (row & 7)) << 4));
}

struct WgTile {
  int g, mp, n_tile;
  bool bias;
};
__device__ __forceinline__ WgTile wg_decode(const WgDev& p, int tile) {
  WgTile t;
  t.bias = tile >= p.num_reg_tiles;
  const int* start = t.bias ? p.bias_start : p.tile_start;
  const int tt = t.bias ? tile - p.num_reg_tiles : tile;
  int g = 0;
  while (g + 1 < p.count && tt >= start[g + 1]) ++g;
  t.g = g;
  const int local = tt - start[g];
  if (t.bias) {
    t.mp = local, t.n_tile = 0;
  } else {
    t.n_tile = local % p.n_tiles[g], t.mp = local / p.n_tiles[g];
  }
  return t;
}

__device__ __forceinline__ int wg_tile_at(const WgDev& p, int i) {
  if (p.sched_len == 0) return i;
  const unsigned short t = p.order[i];
  return t == WG_NO_TILE ? -1 : (int)t;
}

template <bool ACC, typename OpT>
__global__ void __launch_bounds__(256, 1)
wgrad_grouped_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgDev p) {
  using Cfg = WgCfg;
  constexpr int STAGES = Cfg::STAGES, NB = Cfg::NB, BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * Cfg::A_BYTES;
  uint8_t* sSlab = sB + STAGES * Cfg::B_BYTES;
  uint8_t* sOnes = sSlab + NB * Cfg::SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + Cfg::ONES_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* auxfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(auxfull + NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pid = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_sched = p.sched_len ? p.sched_len : p.num_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    for (int i = 0; i < NB; ++i) mbar_init(&auxfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  if (warp == 3) {   // the ones operand of the bias tiles (every element 1.0: any operand layout reads the same)
    const uint32_t one2 = OpTraits<OpT>::pack2(1.0f, 1.0f);
    for (int i = lane; i < Cfg::ONES_BYTES / 4; i += 32) reinterpret_cast<uint32_t*>(sOnes)[i] = one2;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int i = pid; i < n_sched; i += npairs) {
        const int tile = wg_tile_at(p, i);
        if (tile < 0) continue;
        const WgTile t = wg_decode(p, tile);
        const int row0 = t.mp * (2 * Cfg::BM) + rank * Cfg::BM;
        const int wrow0 = t.n_tile * BN + rank * (BN / 2);
        const CUtensorMap* mA = &maps.a[t.g];
        const CUtensorMap* mW = &maps.w[t.g];
        const int num_kb = p.num_kb[t.g];
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + (t.bias ? 0 : Cfg::B_BYTES)));
          uint8_t* dA = sA + stage * Cfg::A_BYTES;
          uint8_t* dB = sB + stage * Cfg::B_BYTES;
          const int kc = kb * Cfg::BK;
          // MN-major operands: two 64 (features) x 64 (tokens) boxes each; tokens / features past the end are zero-filled
          tma_load_3d_2sm(dA, mA, &full[stage], row0, kc, 0);
          tma_load_3d_2sm(dA + Cfg::A_BYTES / 2, mA, &full[stage], row0 + 64, kc, 0);
          if (!t.bias) {
            tma_load_3d_2sm(dB, mW, &full[stage], wrow0, kc, 0);
            tma_load_3d_2sm(dB + Cfg::B_BYTES / 2, mW, &full[stage], wrow0 + 64, kc, 0);
          }
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc(2 * Cfg::BM, BN, OpTraits<OpT>::fmt) | (1u << 15) | (1u << 16);   // A, B MN-major
      const uint32_t idesc_bias = umma_idesc(2 * Cfg::BM, 16, OpTraits<OpT>::fmt) | (1u << 15);           // B = ones, N = 16
      const uint64_t ones_desc = umma_desc_sw128(smem_u32(sOnes));
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int i = pid; i < n_sched; i += npairs) {
        const int tile = wg_tile_at(p, i);
        if (tile < 0) continue;
        const bool bias = tile >= p.num_reg_tiles;
        const int g = wg_decode(p, tile).g;
        const int num_kb = p.num_kb[g], tail_steps = p.tail_steps[g];
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128_mn(smem_u32(sA + stage * Cfg::A_BYTES), Cfg::A_BYTES / 2);
          const uint64_t bdesc = umma_desc_sw128_mn(smem_u32(sB + stage * Cfg::B_BYTES), Cfg::B_BYTES / 2);
          const int steps = (kb == num_kb - 1) ? tail_steps : 4;
          if (bias) {
            for (int kk = 0; kk < steps; ++kk)
              umma_f16_2sm(tmem_d, adesc + 128ull * kk, ones_desc + 2ull * kk, idesc_bias, (kb | kk) != 0);
          } else {
            for (int kk = 0; kk < steps; ++kk)
              umma_f16_2sm(tmem_d, adesc + 128ull * kk, bdesc + 128ull * kk, idesc, (kb | kk) != 0);
          }
          umma_commit_2sm(&empty[stage], 3);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        umma_commit_2sm(&tfull[as], 3);
        if (++as == 2) as = 0, aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs) =====================
    const int ew = warp - 4, et = threadIdx.x - 128, lrow = ew * 32 + lane;
    const uint32_t tempty_leader = mapa_u32(&tempty[0], 0);
    uint32_t as = 0, aphase = 0, slab_ctr = 0, aux_phase = 0;
    for (int i = pid; i < n_sched; i += npairs) {
      const int tile = wg_tile_at(p, i);
      if (tile < 0) continue;
      const WgTile t = wg_decode(p, tile);
      const int g = t.g;
      const int row0 = t.mp * (2 * Cfg::BM) + rank * Cfg::BM;
      const int col_tile = t.n_tile * BN, ncols = p.n[g];
      const CUtensorMap* mO = &maps.out[g];
      if (t.bias) {
        // column 0 of the N = 16 accumulator = sum over tokens of this thread's output feature
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        uint32_t rz[32];
        tmem_ld32(tmem_base + as * BN + (static_cast<uint32_t>(ew * 32) << 16), rz);
        tmem_ld_wait();
        const int r = row0 + lrow;
        if (r < p.n_out[g]) {
          float* dst = p.bias[g] + r;
          const float v = __uint_as_float(rz[0]);
          *dst = ACC ? *dst + v : v;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0)
            mbar_arrive(&tempty[as]);
          else
            mbar_arrive_cluster(tempty_leader + as * 8);
        }
        if (++as == 2) as = 0, aphase ^= 1;
        continue;
      }
      if constexpr (ACC) {
        if (et == 0) {   // gradient slabs 0 and 1 of this tile: buffers last used 4 and 3 slabs ago
          bulk_wait_read<2>();
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const uint32_t b = (slab_ctr + s) % NB;
            if (col_tile + s * 32 < ncols) {
              mbar_expect_tx(&auxfull[b], Cfg::SLAB_BYTES);
              tma_load_3d(sSlab + b * Cfg::SLAB_BYTES, mO, &auxfull[b], col_tile + s * 32, row0, 0);
            }
          }
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(ew * 32) << 16);
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr, ra);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {      // two 32-column slabs per step, the next TMEM load always in flight
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = col_tile + c0 + 32 * h;
          if (col >= ncols) break;                                          // uniform
          const uint32_t b = slab_ctr % NB;
          uint8_t* slab = sSlab + b * Cfg::SLAB_BYTES;
          uint32_t(&r)[32] = h == 0 ? ra : rb;
          if constexpr (ACC) {
            if (et == 0 && col + 64 < ncols && c0 + 32 * h + 64 < BN) {     // prefetch the gradient slab two ahead
              bulk_wait_read<1>();
              const uint32_t b2 = (slab_ctr + 2) % NB;
              mbar_expect_tx(&auxfull[b2], Cfg::SLAB_BYTES);
              tma_load_3d(sSlab + b2 * Cfg::SLAB_BYTES, mO, &auxfull[b2], col + 64, row0, 0);
            }
          } else {
            if (et == 0) bulk_wait_read<NB - 1>();
            asm volatile("bar.sync 1, 128;" ::: "memory");                  // slab free
          }
          tmem_ld_wait();
          if (col + 32 < ncols && c0 + 32 * h + 32 < BN) tmem_ld32(taddr + c0 + 32 * h + 32, h == 0 ? rb : ra);
          if constexpr (ACC) {
            mbar_wait(&auxfull[b], (aux_phase >> b) & 1);
            aux_phase ^= (1u << b);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4* ch = wg_slab_chunk(slab, lrow, i);
              uint4 q = *ch;
              q.x = __float_as_uint(__uint_as_float(q.x) + __uint_as_float(r[4 * i]));
              q.y = __float_as_uint(__uint_as_float(q.y) + __uint_as_float(r[4 * i + 1]));
              q.z = __float_as_uint(__uint_as_float(q.z) + __uint_as_float(r[4 * i + 2]));
              q.w = __float_as_uint(__uint_as_float(q.w) + __uint_as_float(r[4 * i + 3]));
              *ch = q;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *wg_slab_chunk(slab, lrow, i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(mO, slab, col, row0, 0);
            bulk_commit();
          }
          ++slab_ctr;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0)
          mbar_arrive(&tempty[as]);
        else
          mbar_arrive_cluster(tempty_leader + as * 8);
      }
      if (++as == 2) as = 0, aphase ^= 1;
    }
    if (et == 0) bulk_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// Greedy list scheduling on the host (groups with different token counts only; a uniform launch keeps the plain
// round-robin walk, whose neighbouring tiles share operand columns in L2).  Tiles in descending cost order -- cost = the
// tile's 64-token blocks + a fixed hand-off / epilogue share; a bias tile loads half the bytes -- each to the pair with
// the least work so far (ties: the lowest pair, i.e. round-robin among equals, so neighbours still run together).
static void wg_schedule(WgDev& p) {
  const int max_pairs = kNumSMs / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  if (p.num_tiles >= WG_NO_TILE) return;
  std::vector<std::pair<float, int>> tiles(p.num_tiles);
  for (int t = 0; t < p.num_tiles; ++t) {
    const bool bias = t >= p.num_reg_tiles;
    const int* start = bias ? p.bias_start : p.tile_start;
    const int tt = bias ? t - p.num_reg_tiles : t;
    int g = 0;
    while (g + 1 < p.count && tt >= start[g + 1]) ++g;
    tiles[t] = {bias ? 0.5f * p.num_kb[g] + 6.f : (float)p.num_kb[g] + 14.f, t};
  }
  std::stable_sort(tiles.begin(), tiles.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) {
    return a.first > b.first;
  });
  std::vector<float> load(pairs, 0.f);
  std::vector<std::vector<int>> lists(pairs);
  for (const auto& tc : tiles) {
    int best = 0;
    for (int q = 1; q < pairs; ++q)
      if (load[q] < load[best]) best = q;
    load[best] += tc.first;
    lists[best].push_back(tc.second);
  }
  size_t rounds = 0;
  for (const auto& l : lists) rounds = l.size() > rounds ? l.size() : rounds;
  if (rounds * pairs > (size_t)WG_ORDER_MAX) return;      // too many tiles for the table: round-robin
  for (size_t r = 0; r < rounds; ++r)
    for (int q = 0; q < pairs; ++q)
      p.order[r * pairs + q] = r < lists[q].size() ? (unsigned short)lists[q][r] : WG_NO_TILE;
  p.sched_len = (int)(rounds * pairs);
}

template <bool ACC, typename OpT>
static int launch_wgrad(const WgMaps& m, const WgDev& p, cudaStream_t s) {
  auto kern = wgrad_grouped_kernel<ACC, OpT>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(wgrad_grouped)");
    attr_set = true;
  }
  const int max_pairs = kNumSMs / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = WgCfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m, p);
  if (e != cudaSuccess) return fail(e, "wgrad_grouped_kernel launch");
  return 0;
}

}  // namespace mm

extern "C" int mm_wgrad_grouped(const mm_wgrad_group* groups, int32_t count, int64_t tokens_all, int32_t accumulate,
                                int32_t dtype, void* stream) {
  using namespace mm;
  if (!groups || count <= 0 || count > WG_MAX) return bad_arg("wgrad_grouped: 1 .. MM_WGRAD_MAX_GROUPS groups");
  if (dtype != MM_DTYPE_BF16 && dtype != MM_DTYPE_F16) return bad_arg("wgrad_grouped: dtype");
  const int kind = dtype == MM_DTYPE_F16 ? 1 : 0;
  static thread_local WgMaps m;
  static thread_local WgDev p;
  memset(&p, 0, sizeof(p));
  p.count = count;
  p.accumulate = accumulate != 0;
  int tiles = 0, btiles = 0;
  bool uniform = true;
  for (int g = 0; g < count; ++g) {
    const mm_wgrad_group& q = groups[g];
    const int64_t tokens = q.tokens > 0 ? q.tokens : tokens_all;
    if (tokens <= 0) return bad_arg("wgrad_grouped: tokens");
    p.num_kb[g] = (int)((tokens + 63) / 64);
    p.tail_steps[g] = (int)((tokens - (int64_t)(p.num_kb[g] - 1) * 64 + 15) >> 4);
    uniform = uniform && p.num_kb[g] == p.num_kb[0];
    if (!q.dy || q.n_out <= 0 || q.k_in < 0) return bad_arg("wgrad_grouped: group");
    if (q.k_in > 0 && (!q.x || !q.out)) return bad_arg("wgrad_grouped: group with k_in > 0 needs x and out");
    if (q.k_in == 0 && !q.bias) return bad_arg("wgrad_grouped: group computes nothing");
    if (q.k_in % 4) return bad_arg("wgrad_grouped: k_in must be a multiple of 4");
    int rc = make_tmap_3d_ex(&m.a[g], q.dy, kind, (uint64_t)q.n_out, (uint64_t)tokens, 1, (uint64_t)q.dy_ld, 0, 64, 64);
    if (rc) return rc;
    if (q.k_in > 0) {
      rc = make_tmap_3d_ex(&m.w[g], q.x, kind, (uint64_t)q.k_in, (uint64_t)tokens, 1, (uint64_t)q.x_ld, 0, 64, 64);
      if (rc) return rc;
      rc = make_tmap_3d_ex(&m.out[g], q.out, 2, (uint64_t)q.k_in, (uint64_t)q.n_out, 1, (uint64_t)q.out_ld, 0, 32, 128);
      if (rc) return rc;
    } else {
      m.w[g] = m.a[g], m.out[g] = m.a[g];   // never used
    }
    const int m_pairs = (q.n_out + 255) / 256;
    p.tile_start[g] = tiles;
    p.bias_start[g] = btiles;
    p.n_tiles[g] = (q.k_in + 255) / 256;
    p.n[g] = q.k_in;
    p.n_out[g] = q.n_out;
    p.bias[g] = q.bias;
    tiles += m_pairs * p.n_tiles[g];
    if (q.bias) btiles += m_pairs;
  }
  for (int g = count; g <= WG_MAX; ++g) p.tile_start[g] = tiles, p.bias_start[g] = btiles;
  for (int g = count; g < WG_MAX; ++g) m.a[g] = m.a[0], m.w[g] = m.w[0], m.out[g] = m.out[0];
  p.num_reg_tiles = tiles;
  p.num_tiles = tiles + btiles;
  if (!uniform) wg_schedule(p);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (accumulate)
    return kind ? launch_wgrad<true, __half>(m, p, s) : launch_wgrad<true, __nv_bfloat16>(m, p, s);
  return kind ? launch_wgrad<false, __half>(m, p, s) : launch_wgrad<false, __nv_bfloat16>(m, p, s);
}
