// Fused  dh = dY W ;  g <- g + LayerNorm'(dh) ;  g_op <- 16-bit(g) o dropout mask ;  (dgamma, dbeta) partial sums
//
// The backward twin of gemm_ln.cu: the input gradient of a Linear whose input was a LayerNorm output (the QKV projection
// after self_attn_layer_norm, fc1 after final_layer_norm, q_proj after encoder_attn_layer_norm -- fairseq
// TransformerEncoderLayer / TransformerDecoderLayer under autograd) together with the backward of that LayerNorm and the
// residual add.  Un-fused this is a dgrad GEMM writing fp32 dh (33 MB at 16 000 x 512) and a row kernel reading it back
// with x and g; fused, dh never leaves TMEM.
//
//     dh[m, n]  = sum_c dY[m, c] W[c, n]                     W = the Linear's weight [out = c, in = n] as stored (MN-major)
//     xhat      = (x - mean(x)) * rstd(x)                    x = the LayerNorm's fp32 input row (kept by the forward)
//     v         = dh * gamma ;  c1 = mean_n v ;  c2 = mean_n (v * xhat)
//     g[m, n]  += rstd * (v - c1 - xhat * c2)                g = the fp32 residual-stream gradient, updated in place
//     dgamma[n] = sum_m dh * xhat ;  dbeta[n] = sum_m dh     -> per-warp partial rows, summed by mm_reduce_partials_many
//
// A CTA pair owns 256 complete rows (each CTA 128 rows x 512 fp32 accumulator columns = all of TMEM), thread = row.
//   during the main loop   the six idle warps compute mean / rstd of the CTA's 128 rows of x (two-pass, from global)
//   pass 1  x slabs (128 rows x 32 fp32) stream through the idle operand ring; c1, c2 accumulate per row; the column
//           sums over each warp's 32 rows come from a register butterfly (31 shuffles per 32 x 32 block, no staging);
//           the accumulator cell is overwritten with the pair (dh, xhat) rounded to bf16 -- x is not read again
//   pass 2  g slabs stream in, are updated in place and leave by TMA store; the new g also replaces the cell in TMEM
//   pass 3  TMEM -> 16-bit slabs (x keep / (1 - p) of the branch that consumes them) -> TMA store
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

struct LbCfg {
  static constexpr int BM = 128, N = 512, BK = 64, STAGES = 4;
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr int B_BYTES = 256 * BK * 2;           // 32 KB: this CTA's 128 columns of each of the two N halves
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
  static constexpr int SLAB_BYTES = BM * 128;            // 16 KB
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * SLAB_BYTES + 512 + 1024;
};

struct LbDev {
  int rows, k, num_kb, num_tiles;
  const float* x;            // [rows, 512] fp32: the LayerNorm input
  const float* gamma;
  float eps;
  float* partials;           // [4 * gridDim.x][2][512]: (dgamma, dbeta) partial sums, one row per warp quadrant
  float drop_p;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  unsigned site;
};

__device__ __forceinline__ uint4* lb_slab_chunk(uint8_t* slab, int row, int c) {
  return reinterpret_cast<uint4*>(slab + row * 128 + ((c ^ (row & 7)) << 4));
}

// v[c] of lane r = element (row r, column c) of a 32 x 32 block; returns the sum over the 32 rows of column `lane`.
__device__ __forceinline__ float lb_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <typename OpT, bool DROP>
__global__ void __launch_bounds__(256, 1)
gemm_ln_bwd_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                   const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapG,
                   const __grid_constant__ CUtensorMap mapGop, const LbDev p) {
  using Cfg = LbCfg;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                                       // STAGES x (A | B0 | B1); 12 slabs in the epilogue
  uint8_t* extra = ring + STAGES * Cfg::STAGE_BYTES;          // 2 more slabs: row statistics and gamma
  uint64_t* bars = reinterpret_cast<uint64_t*>(extra + 2 * Cfg::SLAB_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = full + STAGES;        // [STAGES]
  uint64_t* tfull = empty + STAGES;       // [1]
  uint64_t* auxfull = tfull + 1;          // [12] slab landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(auxfull + 12);
  float* stat = reinterpret_cast<float*>(extra);              // [c1 | c2][2 halves][128 rows]
  float* s_gamma = stat + 512;                                // [512]
  float* s_mean = s_gamma + 512;                              // [128]
  float* s_rstd = s_mean + 128;                               // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pid = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapG);
    tma_prefetch_desc(&mapGop);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    for (int i = 0; i < 12; ++i) mbar_init(&auxfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  for (int i = threadIdx.x; i < 512; i += 256) s_gamma[i] = __ldg(p.gamma + i);   // a parameter: safe before pdl_wait()
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int h = warp >> 2;                          // column half [256 h, 256 h + 256)
  const int ht = threadIdx.x & 127;
  const int wq = warp & 3;                          // TMEM lane quadrant = rows [32 wq, 32 wq + 32) of the CTA's 128
  const int lrow = wq * 32 + lane;
  const uint32_t taddr = tmem_base + 256 * h + (static_cast<uint32_t>(wq * 32) << 16);
  auto slab_ptr = [&](int b) -> uint8_t* { return ring + (6 * h + b) * Cfg::SLAB_BYTES; };   // 6 slabs per half
  uint64_t* aux = auxfull + 6 * h;
  uint32_t stage = 0, phase = 0, it = 0;
  uint32_t aux_phase = 0;
  const uint64_t pol_keep = l2_policy_evict_last();
  const uint64_t pol_once = l2_policy_evict_first();
  float col_g[8], col_b[8];                         // this lane's column sums (column 256 h + 32 j + lane), all tiles
#pragma unroll
  for (int j = 0; j < 8; ++j) col_g[j] = 0.f, col_b[j] = 0.f;
  const unsigned drop_thr = dropout_threshold(p.drop_p);
  const float drop_inv = 1.0f / (1.0f - p.drop_p);
  const unsigned long long drop_seed = DROP ? p.seed + (p.seed_dev ? *p.seed_dev : 0ull) : 0ull;

  for (int tile = pid; tile < p.num_tiles; tile += npairs, ++it) {
    const int row0 = tile * 256 + rank * Cfg::BM;
    // ===================== phase A: main loop; the idle warps take the row statistics of x =====================
    if (warp == 0 && lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
        uint8_t* st = ring + stage * Cfg::STAGE_BYTES;
        tma_load_3d_2sm_hint(st, &mapA, &full[stage], kb * Cfg::BK, row0, 0, pol_once);
        // W [k, 512] as stored: boxes of 64 output columns x 64 contraction rows (8 KB); this CTA's 128 columns of
        // each 256-column half
        uint8_t* sb = st + Cfg::A_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          tma_load_3d_2sm(sb + q * 8192, &mapW, &full[stage], 256 * (q >> 1) + rank * 128 + 64 * (q & 1), kb * Cfg::BK, 0);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc(256, 256, OpTraits<OpT>::fmt) | (1u << 16);   // B MN-major
      const int k_tail = p.k - (p.num_kb - 1) * Cfg::BK;
      const int tail_steps = (k_tail + 15) >> 4;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        uint8_t* st = ring + stage * Cfg::STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(smem_u32(st));
        const uint64_t b0 = umma_desc_sw128_mn(smem_u32(st + Cfg::A_BYTES), 8192);
        const uint64_t b1 = umma_desc_sw128_mn(smem_u32(st + Cfg::A_BYTES + 16384), 8192);
        const int steps = (kb == p.num_kb - 1) ? tail_steps : 4;
        for (int kk = 0; kk < steps; ++kk) {
          umma_f16_2sm(tmem_base, adesc + 2 * kk, b0 + 128ull * kk, idesc, (kb | kk) != 0);
          umma_f16_2sm(tmem_base + 256, adesc + 2 * kk, b1 + 128ull * kk, idesc, (kb | kk) != 0);
        }
        umma_commit_2sm(&empty[stage], 3);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      umma_commit_2sm(tfull, 3);
    }
    if (warp == 1 && !(lane == 0 && rank == 0)) {
      for (int kb = 0; kb < p.num_kb; ++kb)
        if (++stage == STAGES) stage = 0, phase ^= 1;
    }
    if (warp >= 2) {
      // two-pass mean / rstd of the CTA's 128 rows: a warp takes four rows per step (16 float4 loads in flight per lane,
      // coalesced 512-byte reads), 16 elements of each row per lane
      for (int r0 = 4 * (warp - 2); r0 < Cfg::BM; r0 += 24) {
        float4 q[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int grow = row0 + r0 + u;
          const float4* xr = reinterpret_cast<const float4*>(p.x + (long long)(grow < p.rows ? grow : 0) * Cfg::N);
#pragma unroll
          for (int i = 0; i < 4; ++i) q[u][i] = __ldg(xr + lane + 32 * i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) s += (q[u][i].x + q[u][i].y) + (q[u][i].z + q[u][i].w);
#pragma unroll
          for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * (1.0f / Cfg::N);
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = q[u][i].x - mean, b = q[u][i].y - mean, c = q[u][i].z - mean, d = q[u][i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          const bool live = row0 + r0 + u < p.rows;      // rows past the end: xhat = 0, nothing is stored for them
          if (lane == 0) s_mean[r0 + u] = live ? mean : 0.f, s_rstd[r0 + u] = live ? rsqrtf(ss * (1.0f / Cfg::N) + p.eps) : 0.f;
        }
      }
    }
    __syncwarp();
    // ===================== phase B: epilogue on all 8 warps =====================
    mbar_wait(tfull, it & 1);
    tc_fence_after();
    asm volatile("bar.sync 3, 256;" ::: "memory");      // row statistics visible; every warp has left the main loop
    const float mean = s_mean[lrow], rstd = s_rstd[lrow];
    uint32_t ra[32], rb[32];

    // ---- pass 1: c1, c2, column sums; accumulator cell <- (dh, xhat) as a bf16 pair ----
    if (ht == 0) {
#pragma unroll 1
      for (int j = 0; j < 6; ++j) {
        mbar_expect_tx(&aux[j], Cfg::SLAB_BYTES);
        tma_load_3d_hint(slab_ptr(j), &mapX, &aux[j], 256 * h + 32 * j, row0, 0, pol_once);
      }
    }
    float c1 = 0.f, c2 = 0.f;
    auto pass1 = [&](int j, uint32_t (&r)[32]) {
      const int b = j < 6 ? j : j - 6;
      uint8_t* slab = slab_ptr(b);
      mbar_wait(&aux[b], (aux_phase >> b) & 1);
      aux_phase ^= (1u << b);
      float dy[32], dx[32];
      uint32_t lo[16], hi[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 xq = *lb_slab_chunk(slab, lrow, c);
        const float4 gq = *reinterpret_cast<const float4*>(s_gamma + 256 * h + 32 * j + 4 * c);
        const float xs[4] = {__uint_as_float(xq.x), __uint_as_float(xq.y), __uint_as_float(xq.z), __uint_as_float(xq.w)};
        const float gs[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = __uint_as_float(r[4 * c + e]);
          const float xh = (xs[e] - mean) * rstd;
          const float v = d * gs[e];
          c1 += v;
          c2 = fmaf(v, xh, c2);
          dy[4 * c + e] = d;
          dx[4 * c + e] = d * xh;
          const uint32_t pk = OpTraits<__nv_bfloat16>::pack2(d, xh);
          if (c < 4) lo[4 * c + e] = pk; else hi[4 * (c - 4) + e] = pk;
        }
      }
      tmem_st16(taddr + 32 * j, lo);
      tmem_st16(taddr + 32 * j + 16, hi);
      col_b[j] += lb_column_sums(dy, lane);
      col_g[j] += lb_column_sums(dx, lane);
    };
    tmem_ld32(taddr, ra);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      tmem_ld_wait();
      tmem_ld32(taddr + 32 * (j + 1), rb);
      pass1(j, ra);
      tmem_ld_wait();
      if (j + 2 < 8) tmem_ld32(taddr + 32 * (j + 2), ra);
      pass1(j + 1, rb);
      if (j == 0) {                        // every thread of the half has read slabs 0 and 1: reuse them for 6 and 7
        if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (ht == 0) {
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            mbar_expect_tx(&aux[q], Cfg::SLAB_BYTES);
            tma_load_3d_hint(slab_ptr(q), &mapX, &aux[q], 256 * h + 32 * (6 + q), row0, 0, pol_once);
          }
        }
      }
    }
    tmem_st_wait();
    stat[h * 128 + lrow] = c1;
    stat[256 + h * 128 + lrow] = c2;
    asm volatile("bar.sync 3, 256;" ::: "memory");      // also: every x slab has been consumed
    c1 = (c1 + stat[(h ^ 1) * 128 + lrow]) * (1.0f / Cfg::N);
    c2 = (c2 + stat[256 + (h ^ 1) * 128 + lrow]) * (1.0f / Cfg::N);

    // ---- pass 2: g <- g + rstd * (dh gamma - c1 - xhat c2), in place in the slab; the new g replaces the cell ----
    if (ht == 0) {
#pragma unroll 1
      for (int j = 0; j < 6; ++j) {
        mbar_expect_tx(&aux[j], Cfg::SLAB_BYTES);
        tma_load_3d_hint(slab_ptr(j), &mapG, &aux[j], 256 * h + 32 * j, row0, 0, pol_once);
      }
    }
    auto pass2 = [&](int j, uint32_t (&r)[32]) {
      const int b = j < 6 ? j : j - 6;
      uint8_t* slab = slab_ptr(b);
      mbar_wait(&aux[b], (aux_phase >> b) & 1);
      aux_phase ^= (1u << b);
      uint32_t lo[16], hi[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 gq = *lb_slab_chunk(slab, lrow, c);
        const float4 gm = *reinterpret_cast<const float4*>(s_gamma + 256 * h + 32 * j + 4 * c);
        const float gin[4] = {__uint_as_float(gq.x), __uint_as_float(gq.y), __uint_as_float(gq.z), __uint_as_float(gq.w)};
        const float gs[4] = {gm.x, gm.y, gm.z, gm.w};
        uint32_t* dst = c < 4 ? &lo[4 * c] : &hi[4 * (c - 4)];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t pk = r[4 * c + e];
          const float d = __uint_as_float(pk << 16), xh = __uint_as_float(pk & 0xffff0000u);
          const float out = fmaf(rstd, fmaf(d, gs[e], -c1) - xh * c2, gin[e]);
          dst[e] = __float_as_uint(out);
        }
        *lb_slab_chunk(slab, lrow, c) = make_uint4(dst[0], dst[1], dst[2], dst[3]);
      }
      tmem_st16(taddr + 32 * j, lo);
      tmem_st16(taddr + 32 * j + 16, hi);
    };
    tmem_ld32(taddr, ra);
#pragma unroll 1
    for (int j = 0; j < 8; j += 2) {
      if (j == 4 && ht == 0) {             // buffers 0, 1 (store group 0) are recycled for slabs 6, 7
        bulk_wait_read<1>();
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          mbar_expect_tx(&aux[q], Cfg::SLAB_BYTES);
          tma_load_3d_hint(slab_ptr(q), &mapG, &aux[q], 256 * h + 32 * (6 + q), row0, 0, pol_once);
        }
      }
      tmem_ld_wait();
      tmem_ld32(taddr + 32 * (j + 1), rb);
      pass2(j, ra);
      tmem_ld_wait();
      if (j + 2 < 8) tmem_ld32(taddr + 32 * (j + 2), ra);
      pass2(j + 1, rb);
      fence_proxy_async_smem();
      if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (ht == 0) {
        tma_store_3d_hint(&mapG, slab_ptr(j < 6 ? j : j - 6), 256 * h + 32 * j, row0, 0, pol_keep);
        tma_store_3d_hint(&mapG, slab_ptr(j + 1 < 6 ? j + 1 : j + 1 - 6), 256 * h + 32 * (j + 1), row0, 0, pol_keep);
        bulk_commit();
      }
    }
    tmem_st_wait();

    // ---- pass 3: 16-bit copy of the new g (x dropout mask of the branch that consumes it), 64 columns per slab ----
    // pass-2 store groups used buffers {0,1} {2,3} {4,5} {0,1}: step 0 takes buffer 3, only the last group may still read
    if (ht == 0) bulk_wait_read<1>();
    const unsigned long long drop_row4 = (unsigned long long)(row0 + lrow) * (Cfg::N / 4);
    auto pass3 = [&](int col, const uint32_t (&r)[32], uint8_t* slab, int chunk0) {
      float y[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(r[i]);
      if constexpr (DROP) {
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4)
          dropout_apply4(dropout_bits4(drop_seed, p.site, drop_row4 + (unsigned)((col >> 2) + g4)), drop_thr, drop_inv,
                         y[4 * g4], y[4 * g4 + 1], y[4 * g4 + 2], y[4 * g4 + 3]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 q;
        q.x = OpTraits<OpT>::pack2(y[8 * c + 0], y[8 * c + 1]);
        q.y = OpTraits<OpT>::pack2(y[8 * c + 2], y[8 * c + 3]);
        q.z = OpTraits<OpT>::pack2(y[8 * c + 4], y[8 * c + 5]);
        q.w = OpTraits<OpT>::pack2(y[8 * c + 6], y[8 * c + 7]);
        *lb_slab_chunk(slab, lrow, chunk0 + c) = q;
      }
    };
    tmem_ld32(taddr, ra);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      uint8_t* slab = slab_ptr(3 * ((j + 1) & 1));       // buffers 3 / 0 alternate
      if (ht == 0) bulk_wait_read<1>();
      if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      tmem_ld_wait();
      tmem_ld32(taddr + 64 * j + 32, rb);
      pass3(256 * h + 64 * j, ra, slab, 0);
      tmem_ld_wait();
      if (j + 1 < 4) tmem_ld32(taddr + 64 * (j + 1), ra);
      pass3(256 * h + 64 * j + 32, rb, slab, 4);
      fence_proxy_async_smem();
      if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (ht == 0) {
        tma_store_3d(&mapGop, slab, 256 * h + 64 * j, row0, 0);
        bulk_commit();
      }
    }
    // the operand ring and TMEM are handed back to the main loop of the next tile (both CTAs)
    if (ht == 0) bulk_wait_read<0>();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
  }
  if ((threadIdx.x & 127) == 0) bulk_wait<0>();
  // (dgamma, dbeta) partial sums of this warp's 32 rows (all tiles): row 4 * blockIdx.x + wq, columns of half h
  {
    float* pg = p.partials + ((long long)(4 * blockIdx.x + wq) * 2) * Cfg::N + 256 * h + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      pg[32 * j] = col_g[j];
      pg[Cfg::N + 32 * j] = col_b[j];
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <typename OpT, bool DROP>
static int launch_ln_bwd_gemm(const CUtensorMap& mA, const CUtensorMap& mW, const CUtensorMap& mX, const CUtensorMap& mG,
                              const CUtensorMap& mGop, const LbDev& p, int pairs, cudaStream_t s) {
  auto kern = gemm_ln_bwd_kernel<OpT, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LbCfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(gemm_ln_bwd)");
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = LbCfg::SMEM_BYTES;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA, mW, mX, mG, mGop, p);
  if (e != cudaSuccess) return fail(e, "gemm_ln_bwd_kernel launch");
  return 0;
}

static int lb_pairs(int64_t rows) {
  const int64_t tiles = (rows + 255) / 256;
  const int max_pairs = kNumSMs / 2;
  return (int)(tiles < max_pairs ? tiles : max_pairs);
}

}  // namespace mm

extern "C" int mm_gemm_ln_bwd_partial_rows(int64_t rows) { return rows > 0 ? 8 * mm::lb_pairs(rows) : 0; }

extern "C" int mm_gemm_ln_bwd(const void* dy, int64_t dy_ld, const void* w, int64_t w_ld, int64_t rows, int32_t k,
                              const float* x, const float* gamma, float eps, float* g, void* g_op, float* partials,
                              float drop_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site, int32_t dtype,
                              void* stream) {
  using namespace mm;
  if (!dy || !w || !x || !gamma || !g || !g_op || !partials) return bad_arg("gemm_ln_bwd: null pointer");
  if (rows <= 0 || k <= 0 || (k % 8) || (dy_ld % 8) || (w_ld % 8) || w_ld < LbCfg::N || dy_ld < k)
    return bad_arg("gemm_ln_bwd: extents (k, leading dimensions multiples of 8; w is [k, 512])");
  if (drop_p < 0.f || drop_p >= 1.f) return bad_arg("gemm_ln_bwd: dropout p in [0, 1)");
  if ((reinterpret_cast<uintptr_t>(x) & 15)) return bad_arg("gemm_ln_bwd: x must be 16-byte aligned");
  const int f16 = dtype == MM_DTYPE_F16;
  const uint64_t N = LbCfg::N;
  CUtensorMap mA, mW, mX, mG, mGop;
  int rc = make_tmap_3d(&mA, dy, f16, (uint64_t)k, (uint64_t)rows, 1, (uint64_t)dy_ld, 0, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mW, w, f16 ? 1 : 0, N, (uint64_t)k, 1, (uint64_t)w_ld, 0, 64, 64);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mX, x, 2, N, (uint64_t)rows, 1, N, 0, 32, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mG, g, 2, N, (uint64_t)rows, 1, N, 0, 32, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mGop, g_op, f16 ? 1 : 0, N, (uint64_t)rows, 1, N, 0, 64, 128);
  if (rc) return rc;
  LbDev p;
  memset(&p, 0, sizeof(p));
  p.rows = (int)rows, p.k = k, p.num_kb = (k + 63) / 64, p.num_tiles = (int)((rows + 255) / 256);
  p.x = x, p.gamma = gamma, p.eps = eps, p.partials = partials;
  p.drop_p = drop_p, p.seed = seed, p.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev), p.site = site;
  const int pairs = lb_pairs(rows);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (drop_p > 0.f)
    return f16 ? launch_ln_bwd_gemm<__half, true>(mA, mW, mX, mG, mGop, p, pairs, s)
               : launch_ln_bwd_gemm<__nv_bfloat16, true>(mA, mW, mX, mG, mGop, p, pairs, s);
  return f16 ? launch_ln_bwd_gemm<__half, false>(mA, mW, mX, mG, mGop, p, pairs, s)
             : launch_ln_bwd_gemm<__nv_bfloat16, false>(mA, mW, mX, mG, mGop, p, pairs, s);
}
