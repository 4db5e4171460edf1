// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues, CTA-pair (cta_group::2) version.
//
//   acc[r, c] = sum_k A[r, k] * W[c, k]          A, W 16-bit K-major; fp32 accumulators in TMEM
//
// A cluster of two CTAs (one SM each) owns a 256 x 256 output tile: each CTA stages ITS 128 rows of A and
// ITS 128 rows of W per k-block (32 KB instead of 48 KB per 128x256x64 MACs -> 1.5x less L2->SM traffic,
// which is what bounds the K=512 GEMMs of the encoder), the leader CTA issues one tcgen05.mma
// (M=256, N=256, K=16) per step for both, and each CTA drains its own 128 x 256 fp32 accumulator from TMEM.
//
// Roles per CTA (256 threads), grid = 2 * min(pair_tiles, 74):
//   warp 0    TMA producer (cp.async.bulk.tensor ... cta_group::2: bytes are counted on the LEADER's mbarrier)
//   warp 1    MMA issuer (leader CTA only); tcgen05.commit multicast frees the smem stage in BOTH CTAs and
//             publishes the accumulator to BOTH epilogues
//   warp 2    TMEM allocator (2 accumulator stages x 256 columns: epilogue overlaps the next tile's MMAs)
//   warps 4-7 epilogue: tcgen05.ld -> bias / scale / ReLU / GLU / residual / position / gate -> 128-byte-swizzled
//             16 KB slabs in shared memory -> TMA store (fully coalesced, clipped at the tensor bounds);
//             residual / gate inputs are TMA-loaded into the same slabs two slabs ahead and updated in place.
// The A operand is addressed through a 3-D tensor map (K, rows, batch) whose row stride may be smaller than
// K: that is how the stride-2 Conv1d layers of the subsampler run as GEMMs over a time-major buffer without
// materialising im2col.  Rows / columns past the tensor bounds are zero-filled by TMA on load.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

struct GemmDev {
  int rows, batches, n, k, kb_split, num_kb, w_batched;
  int m_pairs_per_batch, n_tiles, num_tiles;
  const float* bias;
  float scale;
  int scale_cols;
  const float* aux1;
  long long aux_ld;
  int rows_per_seq, out_row_offset;
  void* vt;
  int vt_col0, vt_rows;
  long long vt_ld;
  const float* pos;
  const int* seq_lens;
  // operand layout extensions (backward pass): MN-major operands (memory is [contraction][rows]), split-K batches that
  // advance the contraction coordinate, and (batch, head) decomposition of the batch index with per-head column offsets
  int a_mn, w_mn, a_kbatch, w_kbatch, a_hm, w_hm, out_hm, heads, head_stride;
  // MM_EPI_RELU_OP in the training forward: activation dropout on relu(acc + bias) (fairseq --activation-dropout /
  // --relu-dropout); element index = row * n + column, mask = dropout_keep(seed, site, index)
  float drop_p;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
  unsigned drop_site;
};

struct GemmCfg {
  static constexpr int BM = 128, BN = 256, BK = 64;
#ifndef MM_GEMM_STAGES
#define MM_GEMM_STAGES 5
#endif
#ifndef MM_GEMM_NB
#define MM_GEMM_NB 4
#endif
  static constexpr int STAGES = MM_GEMM_STAGES;
  static constexpr int NB = MM_GEMM_NB;             // epilogue staging slabs
  static constexpr int A_BYTES = BM * BK * 2;       // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 2; // 16 KB: this CTA's half of the W tile
  static constexpr int SLAB_BYTES = BM * 128;       // 128 rows x 128 B
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BIAS_BYTES = BN * 4;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + NB * SLAB_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
};

// ex2.approx + rcp.approx (2 ulp): an IEEE division here is a ~30-instruction branchy subroutine per element, and with
// one epilogue warp per scheduler its latency is fully exposed (the GLU / gate epilogues were 4x slower than the MMAs).
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// this thread's row inside a swizzled slab: 16-byte chunk c lives at position c ^ (row & 7)
__device__ __forceinline__ uint4* slab_chunk(uint8_t* slab, int row, int c) {
  return reinterpret_cast<uint4*>(slab + row * 128 + ((c ^ (row & 7)) << 4));
}
template <typename OpT>
__device__ __forceinline__ void slab_write_op32(uint8_t* slab, int row, int c0, const float* v) {  // 32 values = 4 chunks
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 q;
    q.x = OpTraits<OpT>::pack2(v[8 * i + 0], v[8 * i + 1]);
    q.y = OpTraits<OpT>::pack2(v[8 * i + 2], v[8 * i + 3]);
    q.z = OpTraits<OpT>::pack2(v[8 * i + 4], v[8 * i + 5]);
    q.w = OpTraits<OpT>::pack2(v[8 * i + 6], v[8 * i + 7]);
    *slab_chunk(slab, row, c0 + i) = q;
  }
}
__device__ __forceinline__ void slab_write_f32(uint8_t* slab, int row, const float* v) {  // 32 floats = 8 chunks
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *slab_chunk(slab, row, i) = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                                           __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
}
__device__ __forceinline__ void slab_read_f32(uint8_t* slab, int row, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 q = *slab_chunk(slab, row, i);
    v[4 * i] = __uint_as_float(q.x), v[4 * i + 1] = __uint_as_float(q.y);
    v[4 * i + 2] = __uint_as_float(q.z), v[4 * i + 3] = __uint_as_float(q.w);
  }
}

#ifdef MM_GEMM_TRACE
__device__ long long g_gemm_trace[148 * 4];   // per CTA: clock64 start / end, globaltimer start / end
#endif

template <int MODE, typename OpT>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
            const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapOut0,
            const __grid_constant__ CUtensorMap mapOut1, const __grid_constant__ CUtensorMap mapAux0,
            const GemmDev p) {
  using Cfg = GemmCfg;
  constexpr int STAGES = Cfg::STAGES, NB = Cfg::NB, BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * Cfg::A_BYTES;
  uint8_t* sSlab = sB + STAGES * Cfg::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlab + NB * Cfg::SLAB_BYTES);
  uint64_t* full = bars;                   // [STAGES] TMA -> MMA   (leader's copy is the live one)
  uint64_t* empty = full + STAGES;         // [STAGES] MMA -> TMA   (per CTA, multicast commit)
  uint64_t* tfull = empty + STAGES;        // [2] MMA -> epilogue   (per CTA, multicast commit)
  uint64_t* tempty = tfull + 2;            // [2] epilogues of both CTAs -> MMA (leader's copy)
  uint64_t* auxfull = tempty + 2;          // [NB] TMA aux load -> epilogue (per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(auxfull + NB);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);  // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pid = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapOut0);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);     // leader's expect_tx covers both CTAs' bytes; the peer only issues its TMA
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);   // 4 epilogue warps x 2 CTAs
    }
    for (int i = 0; i < NB; ++i) mbar_init(&auxfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may start its prologue on SMs this grid no longer needs
  pdl_wait();                // everything above overlapped the previous kernel's tail; its outputs are visible now
#ifdef MM_GEMM_TRACE
  if (threadIdx.x == 128) {
    g_gemm_trace[blockIdx.x * 4 + 0] = clock64();
    g_gemm_trace[blockIdx.x * 4 + 2] = (long long)globaltimer_ns();
  }
#endif

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = pid; tile < p.num_tiles; tile += npairs) {
        const int n_tile = tile % p.n_tiles;
        const int mp = tile / p.n_tiles;
        const int bi = mp / p.m_pairs_per_batch;
        const int row0 = (mp % p.m_pairs_per_batch) * (2 * Cfg::BM) + rank * Cfg::BM;
        const int wrow0 = n_tile * BN + rank * (BN / 2);
        const int hb = p.heads > 0 ? bi / p.heads : bi, hcol = p.heads > 0 ? (bi % p.heads) * p.head_stride : 0;
        const int a_b = p.a_hm ? hb : bi, a_c = p.a_hm ? hcol : 0;
        const int w_b = p.w_batched ? (p.w_hm ? hb : bi) : 0, w_c = p.w_hm ? hcol : 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          uint8_t* dA = sA + stage * Cfg::A_BYTES;
          uint8_t* dB = sB + stage * Cfg::B_BYTES;
          if (p.a_mn) {
            // MN-major A: two 64 (rows) x 64 (contraction) boxes, 8 KB each; memory rows are contraction indices
            const int kc = kb * Cfg::BK + (p.a_kbatch ? bi * p.k : 0);
            const int bz = p.a_kbatch ? 0 : a_b;
            tma_load_3d_2sm(dA, &mapA0, &full[stage], row0 + a_c, kc, bz);
            tma_load_3d_2sm(dA + Cfg::A_BYTES / 2, &mapA0, &full[stage], row0 + 64 + a_c, kc, bz);
          } else if (kb < p.kb_split) {
            tma_load_3d_2sm(dA, &mapA0, &full[stage], kb * Cfg::BK + a_c, row0, a_b);
          } else {
            tma_load_3d_2sm(dA, &mapA1, &full[stage], (kb - p.kb_split) * Cfg::BK, row0, bi);
          }
          if (p.w_mn) {
            const int kc = kb * Cfg::BK + (p.w_kbatch ? bi * p.k : 0);
            const int bz = p.w_kbatch ? 0 : w_b;
            tma_load_3d_2sm(dB, &mapW, &full[stage], wrow0 + w_c, kc, bz);
            tma_load_3d_2sm(dB + Cfg::B_BYTES / 2, &mapW, &full[stage], wrow0 + 64 + w_c, kc, bz);
          } else {
            tma_load_3d_2sm(dB, &mapW, &full[stage], kb * Cfg::BK + w_c, wrow0, w_b);
          }
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      // bit 15 / 16: A / B operand is MN-major
      const uint32_t idesc = umma_idesc(2 * Cfg::BM, BN, OpTraits<OpT>::fmt) | (p.a_mn ? (1u << 15) : 0u) |
                             (p.w_mn ? (1u << 16) : 0u);
      // per K=16 step: +32 B inside the 128 B swizzle row (K-major) or +16 rows of 128 B (MN-major)
      const uint64_t a_step = p.a_mn ? 128 : 2, b_step = p.w_mn ? 128 : 2;
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      const int k_tail = p.k - (p.num_kb - 1) * Cfg::BK;          // valid K in the last k-block
      const int tail_steps = (k_tail + 15) >> 4;
      for (int tile = pid; tile < p.num_tiles; tile += npairs) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t aaddr = smem_u32(sA + stage * Cfg::A_BYTES), baddr = smem_u32(sB + stage * Cfg::B_BYTES);
          const uint64_t adesc = p.a_mn ? umma_desc_sw128_mn(aaddr, Cfg::A_BYTES / 2) : umma_desc_sw128(aaddr);
          const uint64_t bdesc = p.w_mn ? umma_desc_sw128_mn(baddr, Cfg::B_BYTES / 2) : umma_desc_sw128(baddr);
          const int steps = (kb == p.num_kb - 1) ? tail_steps : 4;
          for (int kk = 0; kk < steps; ++kk)
            umma_f16_2sm(tmem_d, adesc + a_step * kk, bdesc + b_step * kk, idesc, (kb | kk) != 0);
          umma_commit_2sm(&empty[stage], 3);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        umma_commit_2sm(&tfull[as], 3);
        if (++as == 2) as = 0, aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs) =====================
    const int ew = warp - 4;                 // == warp % 4 -> TMEM lane quadrant
    const int et = threadIdx.x - 128;        // 0..127
    const int lrow = ew * 32 + lane;         // accumulator row (TMEM lane) of this thread
    const uint32_t tempty_leader = mapa_u32(&tempty[0], 0);
    uint32_t as = 0, aphase = 0;
    uint32_t slab_ctr = 0;                   // running slab counter -> staging buffer ring
    uint32_t aux_phase = 0;                  // bit b = parity of auxfull[b]
    constexpr bool kAux = (MODE == MM_EPI_RESID_F32 || MODE == MM_EPI_GATE || MODE == MM_EPI_MASK_OP);
    constexpr int kAuxCols = (MODE == MM_EPI_MASK_OP) ? 64 : 32;   // columns per aux slab (128 B rows)

    for (int tile = pid; tile < p.num_tiles; tile += npairs) {
      const int n_tile = tile % p.n_tiles;
      const int mp = tile / p.n_tiles;
      const int bi = mp / p.m_pairs_per_batch;
      const int row0 = (mp % p.m_pairs_per_batch) * (2 * Cfg::BM) + rank * Cfg::BM;
      const int r = row0 + lrow;                                         // row within the batch
      const bool rvalid = r < p.rows;
      const int col_tile = n_tile * BN;
      const int ob = p.out_hm ? bi / p.heads : bi;                       // output batch / column offset (head mode)
      const int ocol = p.out_hm ? (bi % p.heads) * p.head_stride : 0;

      asm volatile("bar.sync 1, 128;" ::: "memory");   // every thread is done with the previous tile's bias
      for (int i = et; i < BN; i += 128) {
        const int c = col_tile + i;
        sBias[i] = (p.bias != nullptr && c < p.n) ? __ldg(p.bias + c) : 0.f;
      }
      if constexpr (kAux) {
        if (et == 0) {   // residual / text slabs 0 and 1 of this tile: buffers last used 4 and 3 slabs ago
          bulk_wait_read<2>();
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const uint32_t b = (slab_ctr + s) % NB;
            if (col_tile + s * kAuxCols < p.n) {
              mbar_expect_tx(&auxfull[b], Cfg::SLAB_BYTES);
              tma_load_3d(sSlab + b * Cfg::SLAB_BYTES, &mapAux0, &auxfull[b], col_tile + s * kAuxCols, row0, bi);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // bias visible

      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(ew * 32) << 16);

      if constexpr (MODE == MM_EPI_OP || MODE == MM_EPI_RELU_OP) {
        int b = bi, t = r;
        if (p.rows_per_seq > 0) b = r / p.rows_per_seq, t = r - b * p.rows_per_seq;
        // the two 32-column TMEM loads of the NEXT slab are always in flight while the current one is processed
        uint32_t ra0[32], ra1[32];
        tmem_ld32(taddr, ra0);
        tmem_ld32(taddr + 32, ra1);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          const int col = col_tile + c0;
          if (col >= p.n) break;                                          // uniform
          const bool to_vt = (MODE == MM_EPI_OP) && p.vt != nullptr && col >= p.vt_col0;   // uniform
          uint8_t* slab = sSlab + (slab_ctr % NB) * Cfg::SLAB_BYTES;
          tmem_ld_wait();
          float v0[32], v1[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = __uint_as_float(ra0[i]) + sBias[c0 + i];
            v1[i] = __uint_as_float(ra1[i]) + sBias[c0 + 32 + i];
          }
          if (c0 + 64 < BN && col + 64 < p.n) {
            tmem_ld32(taddr + c0 + 64, ra0);
            tmem_ld32(taddr + c0 + 96, ra1);
          }
          if constexpr (MODE == MM_EPI_OP) {
            if (col < p.scale_cols) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v0[i] *= p.scale;
            }
            if (col + 32 < p.scale_cols) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v1[i] *= p.scale;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v0[i] = fmaxf(v0[i], 0.f), v1[i] = fmaxf(v1[i], 0.f);
            if (p.drop_p > 0.f) {     // n % 4 == 0 (checked on the host): a group of four columns shares one hash
              const unsigned thr = dropout_threshold(p.drop_p);
              const float inv = 1.0f / (1.0f - p.drop_p);
              const unsigned long long seed = p.drop_seed + (p.drop_seed_dev ? *p.drop_seed_dev : 0ull);
              const unsigned long long i4 = ((unsigned long long)((long long)bi * p.rows + r) * (unsigned)p.n + (unsigned)col) >> 2;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                dropout_apply4(dropout_bits4(seed, p.drop_site, i4 + g), thr, inv, v0[4 * g], v0[4 * g + 1], v0[4 * g + 2],
                               v0[4 * g + 3]);
                dropout_apply4(dropout_bits4(seed, p.drop_site, i4 + 8 + g), thr, inv, v1[4 * g], v1[4 * g + 1],
                               v1[4 * g + 2], v1[4 * g + 3]);
              }
            }
          }
          if (to_vt) {
            // transposed store: lanes hold consecutive t -> coalesced 2-byte stores per column
            if (rvalid) {
              const int vc = col - p.vt_col0;
              OpT* dst = reinterpret_cast<OpT*>(p.vt) + ((long long)b * p.vt_rows + vc) * p.vt_ld + t;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (col + i < p.n) dst[(long long)i * p.vt_ld] = OpTraits<OpT>::cvt(v0[i]);
                if (col + 32 + i < p.n) dst[(long long)(32 + i) * p.vt_ld] = OpTraits<OpT>::cvt(v1[i]);
              }
            }
          } else {
            if (et == 0) bulk_wait_read<NB - 1>();
            asm volatile("bar.sync 1, 128;" ::: "memory");                // slab free
            slab_write_op32<OpT>(slab, lrow, 0, v0);
            slab_write_op32<OpT>(slab, lrow, 4, v1);
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
              tma_store_3d(&mapOut0, slab, col + ocol, row0 + p.out_row_offset, ob);
              bulk_commit();
            }
            ++slab_ctr;
          }
          __syncwarp();
        }
      } else if constexpr (MODE == MM_EPI_MASK_OP) {
        // dgrad through an activation: out = mask > 0 ? (acc + bias) * scale : 0, where `mask` is the 16-bit activation
        // the forward pass kept (ReLU output, already activation-dropped).  Its 128 x 64 slabs arrive by TMA two slabs
        // ahead and are overwritten in place with the masked gradient, which leaves by TMA store.
        uint32_t ra0[32], ra1[32];
        tmem_ld32(taddr, ra0);
        tmem_ld32(taddr + 32, ra1);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          const int col = col_tile + c0;
          if (col >= p.n) break;                                          // uniform
          const uint32_t b = slab_ctr % NB;
          uint8_t* slab = sSlab + b * Cfg::SLAB_BYTES;
          if (et == 0 && col + 128 < p.n && c0 + 128 < BN) {              // prefetch the mask slab two ahead
            bulk_wait_read<1>();
            const uint32_t b2 = (slab_ctr + 2) % NB;
            mbar_expect_tx(&auxfull[b2], Cfg::SLAB_BYTES);
            tma_load_3d(sSlab + b2 * Cfg::SLAB_BYTES, &mapAux0, &auxfull[b2], col + 128, row0, bi);
          }
          tmem_ld_wait();
          float v[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = (__uint_as_float(ra0[i]) + sBias[c0 + i]) * p.scale;
            v[32 + i] = (__uint_as_float(ra1[i]) + sBias[c0 + 32 + i]) * p.scale;
          }
          if (c0 + 64 < BN && col + 64 < p.n) {
            tmem_ld32(taddr + c0 + 64, ra0);
            tmem_ld32(taddr + c0 + 96, ra1);
          }
          mbar_wait(&auxfull[b], (aux_phase >> b) & 1);
          aux_phase ^= (1u << b);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4* ch = slab_chunk(slab, lrow, i);
            const uint4 m = *ch;
            const OpT* e = reinterpret_cast<const OpT*>(&m);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (!(OpTraits<OpT>::to_float(e[j]) > 0.f)) v[8 * i + j] = 0.f;
            uint4 q;
            q.x = OpTraits<OpT>::pack2(v[8 * i + 0], v[8 * i + 1]);
            q.y = OpTraits<OpT>::pack2(v[8 * i + 2], v[8 * i + 3]);
            q.z = OpTraits<OpT>::pack2(v[8 * i + 4], v[8 * i + 5]);
            q.w = OpTraits<OpT>::pack2(v[8 * i + 6], v[8 * i + 7]);
            *ch = q;
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&mapOut0, slab, col, row0, bi);
            bulk_commit();
          }
          ++slab_ctr;
        }
      } else if constexpr (MODE == MM_EPI_GLU_OP) {
        constexpr int HALF = BN / 2;
#pragma unroll 1
        for (int j0 = 0; j0 < HALF; j0 += 64) {
          uint8_t* slab = sSlab + (slab_ctr % NB) * Cfg::SLAB_BYTES;
          if (et == 0) bulk_wait_read<NB - 1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t ra[32], rg[32];
            tmem_ld32(taddr + j0 + 32 * h, ra);
            tmem_ld32(taddr + HALF + j0 + 32 * h, rg);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float a = __uint_as_float(ra[i]) + sBias[j0 + 32 * h + i];
              const float g = __uint_as_float(rg[i]) + sBias[HALF + j0 + 32 * h + i];
              v[i] = a * sigmoidf_(g);
            }
            slab_write_op32<OpT>(slab, lrow, 4 * h, v);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&mapOut0, slab, n_tile * HALF + j0, row0 + p.out_row_offset, bi);
            bulk_commit();
          }
          ++slab_ctr;
        }
      } else if constexpr (MODE == MM_EPI_GLU_POS_F32) {
        constexpr int HALF = BN / 2;
        const int n_out = p.n >> 1;
        const int slen = (rvalid && p.seq_lens) ? p.seq_lens[bi] : 0x7fffffff;
#pragma unroll 1
        for (int j0 = 0; j0 < HALF; j0 += 32) {
          uint8_t* slab = sSlab + (slab_ctr % NB) * Cfg::SLAB_BYTES;
          if (et == 0) bulk_wait_read<NB - 1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          uint32_t ra[32], rg[32];
          tmem_ld32(taddr + j0, ra);
          tmem_ld32(taddr + HALF + j0, rg);
          const int oc = n_tile * HALF + j0;
          // valid position t -> sinusoidal row t + 2; padded positions get the zero row.  The table reads are issued
          // before the TMEM wait so their L2 latency hides behind it and the sigmoid arithmetic.
          float4 pq[8];
          if (rvalid && r < slen) {
            const float4* pe = reinterpret_cast<const float4*>(p.pos + (long long)(r + 2) * n_out + oc);
#pragma unroll
            for (int i = 0; i < 8; ++i) pq[i] = __ldg(pe + i);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) pq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = __uint_as_float(ra[i]) + sBias[j0 + i];
            const float g = __uint_as_float(rg[i]) + sBias[HALF + j0 + i];
            v[i] = a * sigmoidf_(g) * p.scale;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v[4 * i] += pq[i].x, v[4 * i + 1] += pq[i].y, v[4 * i + 2] += pq[i].z, v[4 * i + 3] += pq[i].w;
          slab_write_f32(slab, lrow, v);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&mapOut0, slab, oc, row0, bi);
            bulk_commit();
          }
          ++slab_ctr;
        }
      } else if constexpr (MODE == MM_EPI_F32) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int col = col_tile + c0;
          if (col >= p.n) break;                                          // uniform
          uint8_t* slab = sSlab + (slab_ctr % NB) * Cfg::SLAB_BYTES;
          if (et == 0) bulk_wait_read<NB - 1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          uint32_t ra[32];
          tmem_ld32(taddr + c0, ra);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]) + sBias[c0 + i];
          slab_write_f32(slab, lrow, v);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&mapOut0, slab, col, row0, bi);
            bulk_commit();
          }
          ++slab_ctr;
        }
      } else if constexpr (MODE == MM_EPI_F32_OP) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          const int col = col_tile + c0;
          if (col >= p.n) break;                                          // uniform
          // three slabs: fp32 columns [0,32), fp32 columns [32,64), 16-bit columns [0,64)
          uint8_t* s0 = sSlab + ((slab_ctr + 0) % NB) * Cfg::SLAB_BYTES;
          uint8_t* s1 = sSlab + ((slab_ctr + 1) % NB) * Cfg::SLAB_BYTES;
          uint8_t* s2 = sSlab + ((slab_ctr + 2) % NB) * Cfg::SLAB_BYTES;
          if (et == 0) bulk_wait_read<0>();   // one group per iteration covers 3 of the 4 slabs
          asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t ra[32];
            tmem_ld32(taddr + c0 + 32 * h, ra);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]) + sBias[c0 + 32 * h + i];
            slab_write_f32(h == 0 ? s0 : s1, lrow, v);
            slab_write_op32<OpT>(s2, lrow, 4 * h, v);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&mapOut0, s0, col, row0, bi);
            if (col + 32 < p.n) tma_store_3d(&mapOut0, s1, col + 32, row0, bi);
            tma_store_3d(&mapOut1, s2, col, row0, bi);
            bulk_commit();
          }
          slab_ctr += 3;
        }
      } else {  // MM_EPI_RESID_F32 / MM_EPI_GATE: in-place update of TMA-loaded fp32 slabs
        const long long arow = (long long)bi * p.rows + r;
        // gate: the attention output `o` of this row comes straight from global memory (each thread its own 128 B);
        // the loads for chunk c + 1 are issued while chunk c is processed, a full chunk ahead of their use
        float o[32], o_next[32];
        auto load_o = [&](int col, float (&dst)[32]) {
          if (rvalid && col < p.n) {
            const float4* op = reinterpret_cast<const float4*>(p.aux1 + arow * p.aux_ld + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = __ldg(op + i);
              dst[4 * i] = q.x, dst[4 * i + 1] = q.y, dst[4 * i + 2] = q.z, dst[4 * i + 3] = q.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[i] = 0.f;
          }
        };
        if constexpr (MODE == MM_EPI_GATE) load_o(col_tile, o_next);
#pragma unroll 1
        for (int c0 = 0, s = 0; c0 < BN; c0 += 32, ++s) {
          const int col = col_tile + c0;
          if (col >= p.n) break;                                          // uniform
          const uint32_t b = slab_ctr % NB;
          uint8_t* slab = sSlab + b * Cfg::SLAB_BYTES;
          if (et == 0 && col + 64 < p.n && c0 + 64 < BN) {                // prefetch the slab two ahead
            bulk_wait_read<1>();
            const uint32_t b2 = (slab_ctr + 2) % NB;
            mbar_expect_tx(&auxfull[b2], Cfg::SLAB_BYTES);
            tma_load_3d(sSlab + b2 * Cfg::SLAB_BYTES, &mapAux0, &auxfull[b2], col + 64, row0, bi);
          }
          uint32_t ra[32];
          tmem_ld32(taddr + c0, ra);
          if constexpr (MODE == MM_EPI_GATE) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = o_next[i];
            if (c0 + 32 < BN) load_o(col + 32, o_next);
          }
          tmem_ld_wait();
          mbar_wait(&auxfull[b], (aux_phase >> b) & 1);
          aux_phase ^= (1u << b);
          float v[32], x[32];
          slab_read_f32(slab, lrow, x);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]) + sBias[c0 + i];
          if constexpr (MODE == MM_EPI_RESID_F32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += x[i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float g = sigmoidf_(v[i]);
              v[i] = (1.0f - g) * x[i] + g * o[i];
            }
          }
          slab_write_f32(slab, lrow, v);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            tma_store_3d(&mapOut0, slab, col, row0, bi);
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
    if (et == 0) bulk_wait<0>();   // all TMA stores of this CTA complete before its smem goes away
#ifdef MM_GEMM_TRACE
    if (et == 0) {
      g_gemm_trace[blockIdx.x * 4 + 1] = clock64();
      g_gemm_trace[blockIdx.x * 4 + 3] = (long long)globaltimer_ns();
    }
#endif
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
thread_local char g_last_error[256] = {0};

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_3d_ex(CUtensorMap* out, const void* base, int kind, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                    uint64_t stride1, uint64_t stride2, uint32_t box0, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return bad_arg("cuTensorMapEncodeTiled entry point not available");
  const uint64_t es = kind == 2 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1 * es) % 16 || (stride2 * es) % 16)
    return bad_arg("TMA operand base/strides must be 16-byte aligned");
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1 * es, stride2 * es};
  if (dim2 == 1 && strides[1] == 0) strides[1] = strides[0] * dim1;
  cuuint32_t box[3] = {box0, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                        : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error),
             "cuTensorMapEncodeTiled failed (%d): kind %d dims %llu %llu %llu strides %llu %llu box %u %u", (int)r,
             kind, (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2,
             (unsigned long long)strides[0], (unsigned long long)strides[1], box0, box_rows);
    return static_cast<int>(cudaErrorInvalidValue);
  }
  return 0;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int is_f16, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                 uint64_t stride1, uint64_t stride2, uint32_t box_rows) {
  return make_tmap_3d_ex(out, base, is_f16 ? 1 : 0, dim0, dim1, dim2, stride1, stride2, 64, box_rows);
}

struct GemmMaps {
  CUtensorMap a0, a1, w, out0, out1, aux0;
};

template <int MODE, typename OpT>
static int launch_gemm(const GemmMaps& m, const GemmDev& p, cudaStream_t s) {
  using Cfg = GemmCfg;
  auto kern = gemm_kernel<MODE, OpT>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(gemm)");
    attr_set = true;
  }
  const int max_pairs = kNumSMs / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m.a0, m.a1, m.w, m.out0, m.out1, m.aux0, p);
  if (e != cudaSuccess) return fail(e, "gemm_kernel launch");
  return 0;
}

template <typename OpT>
static int dispatch_mode(int mode, const GemmMaps& m, const GemmDev& p, cudaStream_t s) {
  switch (mode) {
    case MM_EPI_OP: return launch_gemm<MM_EPI_OP, OpT>(m, p, s);
    case MM_EPI_RELU_OP: return launch_gemm<MM_EPI_RELU_OP, OpT>(m, p, s);
    case MM_EPI_RESID_F32: return launch_gemm<MM_EPI_RESID_F32, OpT>(m, p, s);
    case MM_EPI_GLU_OP: return launch_gemm<MM_EPI_GLU_OP, OpT>(m, p, s);
    case MM_EPI_GLU_POS_F32: return launch_gemm<MM_EPI_GLU_POS_F32, OpT>(m, p, s);
    case MM_EPI_F32_OP: return launch_gemm<MM_EPI_F32_OP, OpT>(m, p, s);
    case MM_EPI_GATE: return launch_gemm<MM_EPI_GATE, OpT>(m, p, s);
    case MM_EPI_F32: return launch_gemm<MM_EPI_F32, OpT>(m, p, s);
    case MM_EPI_MASK_OP: return launch_gemm<MM_EPI_MASK_OP, OpT>(m, p, s);
  }
  return bad_arg("unknown epilogue mode");
}

}  // namespace mm

#ifdef MM_GEMM_TRACE
extern "C" int mm_debug_gemm_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, mm::g_gemm_trace, sizeof(long long) * 148 * 4);
}
#endif

extern "C" int mm_gemm(const mm_gemm_args* a, void* stream) {
  using namespace mm;
  if (!a || !a->a0 || !a->w || !a->out0) return bad_arg("null operand");
  if (a->rows <= 0 || a->batches <= 0 || a->n <= 0 || a->k <= 0) return bad_arg("non-positive extent");
  if (a->dtype != MM_DTYPE_BF16 && a->dtype != MM_DTYPE_F16) return bad_arg("dtype");
  const int k0 = a->a1 ? a->k_split : a->k;
  if (a->a1 && (k0 <= 0 || k0 % 64 != 0 || k0 >= a->k)) return bad_arg("k_split must be a multiple of 64 inside (0,k)");
  if (a->rows_per_seq > 0 && a->batches != 1) return bad_arg("rows_per_seq requires batches == 1");
  if (a->block_n != 0 && a->block_n != 256) return bad_arg("block_n must be 0 or 256");
  const int mode = a->mode;
  const bool glu = mode == MM_EPI_GLU_OP || mode == MM_EPI_GLU_POS_F32;
  if (glu && (a->n % 256 != 0)) return bad_arg("GLU epilogue needs n % 256 == 0");
  if (mode == MM_EPI_RESID_F32 && !a->aux0) return bad_arg("RESID needs aux0");
  if (mode == MM_EPI_GATE && (!a->aux0 || !a->aux1)) return bad_arg("GATE needs aux0 and aux1");
  if (mode == MM_EPI_F32_OP && !a->out1) return bad_arg("F32_OP needs out1");
  if (mode == MM_EPI_MASK_OP && (!a->aux0 || a->out_tbc || a->vt)) return bad_arg("MASK_OP needs aux0 (16-bit), no tbc / vt");
  if (mode == MM_EPI_GLU_POS_F32 && !a->pos) return bad_arg("GLU_POS needs pos");
  if (a->out_tbc && (a->n_seqs <= 0 || a->batches != a->n_seqs))
    return bad_arg("out_tbc needs batched rows (batches == n_seqs)");
  if (a->vt && mode != MM_EPI_OP) return bad_arg("vt only with MM_EPI_OP");
  if (a->vt && (a->vt_col0 % 64)) return bad_arg("vt_col0 must be a multiple of 64");
  if (a->rows_per_seq > 0 && !(mode == MM_EPI_OP && a->vt)) {
    // (b, t) decomposition of flat rows is only needed by the transposed V store
  }
  if ((mode == MM_EPI_RESID_F32 || mode == MM_EPI_GATE) && (a->n % 32)) return bad_arg("RESID/GATE need n % 32 == 0");

  const bool layouts = a->a_mn || a->w_mn || a->a_hm || a->w_hm || a->out_hm;
  if (layouts) {
    if (a->a1) return bad_arg("MN-major / head-mode operands exclude the split A operand");
    if ((a->a_hm || a->w_hm || a->out_hm) && (a->heads <= 0 || a->head_stride <= 0 || a->batches % a->heads))
      return bad_arg("head mode needs heads > 0, head_stride > 0 and batches % heads == 0");
    if (a->out_hm && (mode != MM_EPI_OP || a->vt || a->n % 64)) return bad_arg("out_hm needs MM_EPI_OP and n % 64 == 0");
    if ((a->a_kbatch && !a->a_mn) || (a->w_kbatch && !a->w_mn)) return bad_arg("kbatch needs an MN-major operand");
    if ((a->a_kbatch || a->w_kbatch) && a->k % 64) return bad_arg("split-K batches need k % 64 == 0");
    if (a->w_hm && !a->w_batched) return bad_arg("w_hm needs w_batched");
  }
  GemmMaps m;
  const int f16 = a->dtype == MM_DTYPE_F16;
  const uint64_t rows = (uint64_t)a->rows, nb = (uint64_t)a->batches;
  const uint64_t hw = (uint64_t)a->heads * (uint64_t)a->head_stride;      // column extent of a head-mode tensor
  const uint64_t nseq = a->heads > 0 ? nb / (uint64_t)a->heads : nb;
  int rc;
  if (a->a_mn) {
    // memory [batch][contraction][rows]: dim0 = rows (or all heads' columns), dim1 = contraction indices
    const uint64_t d0 = a->a_hm ? hw : rows;
    if (a->a_kbatch)
      rc = make_tmap_3d_ex(&m.a0, a->a0, f16 ? 1 : 0, d0, (uint64_t)(a->a_k_total > 0 ? a->a_k_total : nb * k0), 1,
                           (uint64_t)a->a0_ld, 0, 64, 64);
    else
      rc = make_tmap_3d_ex(&m.a0, a->a0, f16 ? 1 : 0, d0, (uint64_t)k0, a->a_hm ? nseq : nb, (uint64_t)a->a0_ld,
                           (uint64_t)a->a0_bs, 64, 64);
  } else {
    rc = make_tmap_3d(&m.a0, a->a0, f16, a->a_hm ? hw : (uint64_t)k0, rows, a->a_hm ? nseq : nb, (uint64_t)a->a0_ld,
                      (uint64_t)a->a0_bs, 128);
  }
  if (rc) return rc;
  if (a->a1) {
    rc = make_tmap_3d(&m.a1, a->a1, f16, (uint64_t)(a->k - k0), rows, nb, (uint64_t)a->a1_ld, (uint64_t)a->a1_bs, 128);
    if (rc) return rc;
  } else {
    m.a1 = m.a0;
  }
  const uint64_t wb = a->w_batched ? (a->w_hm ? nseq : nb) : 1;
  if (a->w_mn) {
    const uint64_t d0 = a->w_hm ? hw : (uint64_t)a->n;
    if (a->w_kbatch)
      rc = make_tmap_3d_ex(&m.w, a->w, f16 ? 1 : 0, d0, (uint64_t)(a->w_k_total > 0 ? a->w_k_total : nb * a->k), 1,
                           (uint64_t)a->w_ld, 0, 64, 64);
    else
      rc = make_tmap_3d_ex(&m.w, a->w, f16 ? 1 : 0, d0, (uint64_t)a->k, wb, (uint64_t)a->w_ld, (uint64_t)a->w_bs, 64,
                           64);
  } else {
    rc = make_tmap_3d(&m.w, a->w, f16, a->w_hm ? hw : (uint64_t)a->k, (uint64_t)a->n, wb, (uint64_t)a->w_ld,
                      (uint64_t)a->w_bs, 128);
  }
  if (rc) return rc;

  // outputs: 3-D maps (columns, rows, batch); T x B x C stores are just another stride pair
  const bool out_op = mode == MM_EPI_OP || mode == MM_EPI_RELU_OP || mode == MM_EPI_GLU_OP || mode == MM_EPI_MASK_OP;
  uint64_t n_out = glu ? (uint64_t)a->n / 2 : (uint64_t)a->n;
  if (mode == MM_EPI_OP && a->vt) n_out = (uint64_t)a->vt_col0;
  const uint64_t out_rows = rows + (uint64_t)a->out_row_offset;
  uint64_t s1 = (uint64_t)a->out0_ld, s2 = (uint64_t)a->out0_bs;
  if (a->out_tbc) s1 = (uint64_t)a->n_seqs * a->out0_ld, s2 = (uint64_t)a->out0_ld;
  if (n_out > 0) {
    rc = make_tmap_3d_ex(&m.out0, a->out0, out_op ? (f16 ? 1 : 0) : 2, a->out_hm ? hw : n_out, out_rows,
                         a->out_hm ? nseq : nb, s1, s2, out_op ? 64 : 32, 128);
    if (rc) return rc;
  } else {
    m.out0 = m.a0;  // every column goes to vt; never used
  }
  if (mode == MM_EPI_F32_OP) {
    rc = make_tmap_3d_ex(&m.out1, a->out1, f16 ? 1 : 0, (uint64_t)a->n, rows, nb, (uint64_t)a->out1_ld,
                         (uint64_t)a->out1_bs, 64, 128);
    if (rc) return rc;
  } else {
    m.out1 = m.out0;
  }
  if (mode == MM_EPI_RESID_F32 || mode == MM_EPI_GATE) {
    rc = make_tmap_3d_ex(&m.aux0, a->aux0, 2, (uint64_t)a->n, rows, nb, (uint64_t)a->aux_ld,
                         (uint64_t)a->rows * a->aux_ld, 32, 128);
    if (rc) return rc;
  } else if (mode == MM_EPI_MASK_OP) {
    rc = make_tmap_3d_ex(&m.aux0, a->aux0, f16 ? 1 : 0, (uint64_t)a->n, rows, nb, (uint64_t)a->aux_ld,
                         (uint64_t)a->rows * a->aux_ld, 64, 128);
    if (rc) return rc;
  } else {
    m.aux0 = m.out0;
  }

  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.rows = a->rows, p.batches = a->batches, p.n = a->n, p.k = a->k;
  p.num_kb = (a->k + 63) / 64;
  p.kb_split = a->a1 ? k0 / 64 : p.num_kb;
  p.w_batched = a->w_batched;
  p.m_pairs_per_batch = (a->rows + 255) / 256;
  p.n_tiles = (a->n + 255) / 256;
  p.num_tiles = p.m_pairs_per_batch * a->batches * p.n_tiles;
  p.bias = a->bias, p.scale = a->scale, p.scale_cols = a->scale_cols;
  p.aux1 = a->aux1, p.aux_ld = a->aux_ld;
  p.rows_per_seq = a->rows_per_seq, p.out_row_offset = a->out_row_offset;
  p.vt = a->vt, p.vt_col0 = a->vt_col0, p.vt_rows = a->vt_rows, p.vt_ld = a->vt_ld;
  p.pos = a->pos, p.seq_lens = a->seq_lens;
  p.a_mn = a->a_mn, p.w_mn = a->w_mn, p.a_kbatch = a->a_kbatch, p.w_kbatch = a->w_kbatch;
  p.a_hm = a->a_hm, p.w_hm = a->w_hm, p.out_hm = a->out_hm, p.heads = a->heads, p.head_stride = a->head_stride;
  if (a->drop_p != 0.f) {
    if (mode != MM_EPI_RELU_OP || a->drop_p < 0.f || a->drop_p >= 1.f || (a->n % 4))
      return bad_arg("drop_p needs MM_EPI_RELU_OP, p in [0, 1) and n % 4 == 0");
    p.drop_p = a->drop_p, p.drop_seed = a->drop_seed, p.drop_site = a->drop_site;
    p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(a->drop_seed_dev);
  }

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return f16 ? dispatch_mode<__half>(mode, m, p, s) : dispatch_mode<__nv_bfloat16>(mode, m, p, s);
}
