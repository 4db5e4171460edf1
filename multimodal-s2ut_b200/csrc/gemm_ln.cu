// Fused  x <- x + A W^T + b ;  h <- LayerNorm(x) * gamma + beta      (N = d_model = 512 full rows)
//
// This is the out_proj / fc2 step of a pre-LN transformer layer TOGETHER with the LayerNorm that opens the
// next sub-layer (fairseq TransformerEncoderLayer: `x = residual + dropout(...)` followed by
// `self_attn_layer_norm` / `final_layer_norm` / the encoder's last `layer_norm`).  Fusing them removes one
// full read of the fp32 residual stream and one kernel launch per sub-layer.
//
// A CTA pair (cta_group::2) owns 256 complete rows: each CTA accumulates 128 rows x 512 fp32 columns = all 512
// TMEM columns (two N=256 UMMAs per K step).  Because a row is complete inside one thread's TMEM lane, the
// LayerNorm statistics need no cross-thread reduction at all.  Epilogue per CTA (thread = row):
//   sweep 1  residual slabs (128 rows x 32 fp32) arrive by TMA into the operand ring, which is idle once the last
//            MMA has completed (12 + 2 slabs in flight); v = acc + bias + x is written back to TMEM and to the slab,
//            the slab is TMA-stored as the new fp32 residual; row sum / sum of squares accumulate in registers
//   sweep 2  (v - mean) * rstd * gamma + beta -> 16-bit (and optionally fp32) slabs -> TMA store
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

struct LnCfg {
  static constexpr int BM = 128, N = 512, BK = 64, STAGES = 4;
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr int B_BYTES = 256 * BK * 2;           // 32 KB: this CTA's 128 rows of each of the two N halves
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
  static constexpr int SLAB_BYTES = BM * 128;            // 16 KB
  static constexpr int RING_SLABS = STAGES * STAGE_BYTES / SLAB_BYTES;  // 12
  static constexpr int NSLAB = RING_SLABS + 2;           // 14 slab buffers during the epilogue
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * SLAB_BYTES + 512 + 1024;
};

struct LnDev {
  int rows, k, num_kb, num_tiles;
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  int want_f32;
  int l2_hints;   // bit 0: residual stream evict_last, bit 1: A operand evict_first
  // training: dropout on the sub-layer output before the residual add, x <- x + dropout(a W^T + b) (fairseq
  // TransformerEncoderLayer dropout_module); element index = row * 512 + column, mask = dropout_keep(seed, site, index)
  float drop_p;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  unsigned site;
};

__device__ __forceinline__ uint4* ln_slab_chunk(uint8_t* slab, int row, int c) {
  return reinterpret_cast<uint4*>(slab + row * 128 + ((c ^ (row & 7)) << 4));
}

#ifdef MM_LN_TRACE
__device__ long long g_ln_trace[148 * 2 * 16];
#define LN_TRACE(slot)                                                                      \
  do {                                                                                      \
    if (ht == 0 && it == 0) g_ln_trace[(blockIdx.x * 2 + h) * 16 + (slot)] = clock64();       \
  } while (0)
#else
#define LN_TRACE(slot)
#endif

template <typename OpT>
__global__ void __launch_bounds__(256, 1)
gemm_resid_ln_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                     const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapH,
                     const __grid_constant__ CUtensorMap mapHf, const __grid_constant__ CUtensorMap mapXo,
                     const LnDev p) {
  using Cfg = LnCfg;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* ring = smem;                                       // STAGES x (A | B0 | B1); 12 slabs in the epilogue
  uint8_t* extra = ring + STAGES * Cfg::STAGE_BYTES;          // 2 more slabs (row statistics live here)
  uint64_t* bars = reinterpret_cast<uint64_t*>(extra + 2 * Cfg::SLAB_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = full + STAGES;        // [STAGES]
  uint64_t* tfull = empty + STAGES;       // [1] accumulator complete (multicast commit)
  uint64_t* auxfull = tfull + 1;          // [12] residual slab landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(auxfull + 12);
  float* stat = reinterpret_cast<float*>(extra);              // [sum | sumsq][2 halves][128 rows]
  float* s_bias = stat + 512;                                 // [512] bias, gamma, beta: read once from global; the
  float* s_gamma = s_bias + 512;                              // epilogue then takes them as shared-memory broadcasts
  float* s_beta = s_gamma + 512;                              // (their L2 latency was exposed in every 32-column step)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pid = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapH);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    for (int i = 0; i < 12; ++i) mbar_init(&auxfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  for (int i = threadIdx.x; i < 512; i += 256) {   // parameters, not activations: safe to read before pdl_wait()
    s_bias[i] = __ldg(p.bias + i);
    s_gamma[i] = __ldg(p.gamma + i);
    s_beta[i] = __ldg(p.beta + i);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may start its prologue on SMs this grid no longer needs
  pdl_wait();                // everything above overlapped the previous kernel's tail; its outputs are visible now

  // epilogue geometry: warps w and w+4 share TMEM lane quadrant (w & 3); half h owns columns [256 h, 256 h + 256)
  const int h = warp >> 2;
  const int ht = threadIdx.x & 127;                 // thread index inside the half
  const int lrow = (warp & 3) * 32 + lane;
  const uint32_t taddr = tmem_base + 256 * h + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  auto slab_ptr = [&](int b) -> uint8_t* { return ring + (6 * h + b) * Cfg::SLAB_BYTES; };   // 6 slabs per half
  uint64_t* aux = auxfull + 6 * h;
  uint32_t stage = 0, phase = 0, it = 0;
  uint32_t aux_phase = 0;                           // bit b = parity of aux[b]
  // L2 residency: the fp32 residual stream (rows x 512 x 4 B) is re-read by the next fused GEMM two kernels later,
  // while the activations between them are read exactly once; ask L2 to keep the former and drop the latter first.
  const uint64_t pol_x = (p.l2_hints & 1) ? l2_policy_evict_last() : l2_policy_evict_normal();
  const uint64_t pol_a = (p.l2_hints & 2) ? l2_policy_evict_first() : l2_policy_evict_normal();
#ifdef MM_LN_TRACE
  if (ht == 0) g_ln_trace[(blockIdx.x * 2 + h) * 16 + 0] = clock64();
#endif

  for (int tile = pid; tile < p.num_tiles; tile += npairs, ++it) {
    const int row0 = tile * 256 + rank * Cfg::BM;
    // ===================== phase A: mainloop (one producer thread per CTA, one MMA thread per pair) ==========
    if (warp == 0 && lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
        uint8_t* st = ring + stage * Cfg::STAGE_BYTES;
        tma_load_3d_2sm_hint(st, &mapA, &full[stage], kb * Cfg::BK, row0, 0, pol_a);
        tma_load_3d_2sm(st + Cfg::A_BYTES, &mapW, &full[stage], kb * Cfg::BK, rank * 128, 0);
        tma_load_3d_2sm(st + Cfg::A_BYTES + 16384, &mapW, &full[stage], kb * Cfg::BK, 256 + rank * 128, 0);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc(256, 256, OpTraits<OpT>::fmt);
      const int k_tail = p.k - (p.num_kb - 1) * Cfg::BK;
      const int tail_steps = (k_tail + 15) >> 4;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        uint8_t* st = ring + stage * Cfg::STAGE_BYTES;
        const uint64_t adesc = umma_desc_sw128(smem_u32(st));
        const uint64_t b0 = umma_desc_sw128(smem_u32(st + Cfg::A_BYTES));
        const uint64_t b1 = umma_desc_sw128(smem_u32(st + Cfg::A_BYTES + 16384));
        const int steps = (kb == p.num_kb - 1) ? tail_steps : 4;
        for (int kk = 0; kk < steps; ++kk) {
          umma_f16_2sm(tmem_base, adesc + 2 * kk, b0 + 2 * kk, idesc, (kb | kk) != 0);
          umma_f16_2sm(tmem_base + 256, adesc + 2 * kk, b1 + 2 * kk, idesc, (kb | kk) != 0);
        }
        umma_commit_2sm(&empty[stage], 3);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      umma_commit_2sm(tfull, 3);
    }
    if (warp == 1 && !(lane == 0 && rank == 0)) {   // keep the (stage, phase) bookkeeping of warp 1 consistent
      for (int kb = 0; kb < p.num_kb; ++kb)
        if (++stage == STAGES) stage = 0, phase ^= 1;
    }
    __syncwarp();
    // ===================== phase B: epilogue on all 8 warps =====================
    mbar_wait(tfull, it & 1);     // accumulator complete => every MMA has finished reading the operand ring
    tc_fence_after();
    LN_TRACE(1);
    if (ht == 0) {                // this half's residual slabs 0..5 (slabs 6, 7 recycle buffers 0, 1)
#pragma unroll 1
      for (int j = 0; j < 6; ++j) {
        mbar_expect_tx(&aux[j], Cfg::SLAB_BYTES);
        tma_load_3d_hint(slab_ptr(j), &mapX, &aux[j], 256 * h + 32 * j, row0, 0, pol_x);
      }
    }
    // ---- sweep 1: v = acc + bias + x ; new residual out ; row sum ----
    float sum = 0.f, sumsq = 0.f;
    uint32_t ra[32], rb[32];
    const bool drop = p.drop_p > 0.f;
    const unsigned drop_thr = dropout_threshold(p.drop_p);
    const float drop_inv = 1.0f / (1.0f - p.drop_p);
    const unsigned long long drop_seed = p.seed + ((drop && p.seed_dev) ? *p.seed_dev : 0ull);
    const unsigned long long drop_row4 = (unsigned long long)(row0 + lrow) * (Cfg::N / 4);
    auto sweep1 = [&](int j, uint32_t (&r)[32]) {
      const int b = j < 6 ? j : j - 6;
      uint8_t* slab = slab_ptr(b);
      const float4* b4 = reinterpret_cast<const float4*>(s_bias + 256 * h + 32 * j);
      float4 bq[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) bq[c] = b4[c];
      mbar_wait(&aux[b], (aux_phase >> b) & 1);
      aux_phase ^= (1u << b);
      uint4 xq[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) xq[c] = *ln_slab_chunk(slab, lrow, c);
      uint32_t lo[16], hi[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float a0 = __uint_as_float(r[4 * c + 0]) + bq[c].x, a1 = __uint_as_float(r[4 * c + 1]) + bq[c].y;
        float a2 = __uint_as_float(r[4 * c + 2]) + bq[c].z, a3 = __uint_as_float(r[4 * c + 3]) + bq[c].w;
        if (drop)
          dropout_apply4(dropout_bits4(drop_seed, p.site, drop_row4 + (unsigned)(64 * h + 8 * j + c)), drop_thr, drop_inv, a0,
                         a1, a2, a3);
        const float v0 = a0 + __uint_as_float(xq[c].x);
        const float v1 = a1 + __uint_as_float(xq[c].y);
        const float v2 = a2 + __uint_as_float(xq[c].z);
        const float v3 = a3 + __uint_as_float(xq[c].w);
        sum += (v0 + v1) + (v2 + v3);
        sumsq += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
        uint32_t* dst = c < 4 ? &lo[4 * c] : &hi[4 * (c - 4)];
        dst[0] = __float_as_uint(v0), dst[1] = __float_as_uint(v1);
        dst[2] = __float_as_uint(v2), dst[3] = __float_as_uint(v3);
        *ln_slab_chunk(slab, lrow, c) = make_uint4(dst[0], dst[1], dst[2], dst[3]);
      }
      tmem_st16(taddr + 32 * j, lo);          // keep v in TMEM for sweeps 2 and 3
      tmem_st16(taddr + 32 * j + 16, hi);
    };
    tmem_ld32(taddr, ra);
#pragma unroll 1
    for (int j = 0; j < 8; j += 2) {
      if (j == 4 && ht == 0) {   // buffers 0, 1 (store group 0) are recycled for slabs 6, 7
        bulk_wait_read<1>();
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          mbar_expect_tx(&aux[q], Cfg::SLAB_BYTES);
          tma_load_3d_hint(slab_ptr(q), &mapX, &aux[q], 256 * h + 32 * (6 + q), row0, 0, pol_x);
        }
      }
      tmem_ld_wait();
      tmem_ld32(taddr + 32 * (j + 1), rb);
      sweep1(j, ra);
      tmem_ld_wait();
      if (j + 2 < 8) tmem_ld32(taddr + 32 * (j + 2), ra);
      sweep1(j + 1, rb);
      fence_proxy_async_smem();
      if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (ht == 0) {
        // mapXo == mapX: residual stream updated in place; a separate output keeps the input (training: activations kept)
        tma_store_3d_hint(&mapXo, slab_ptr(j < 6 ? j : j - 6), 256 * h + 32 * j, row0, 0, pol_x);
        tma_store_3d_hint(&mapXo, slab_ptr(j + 1 < 6 ? j + 1 : j + 1 - 6), 256 * h + 32 * (j + 1), row0, 0, pol_x);
        bulk_commit();
      }
    }
    tmem_st_wait();
    LN_TRACE(2);
    // Row statistics: the two halves exchange (sum, sum of squares).  Raw-moment variance in fp32 is accurate to
    // ~1e-6 here because the residual stream has |mean| <~ std (checked against a two-pass fp64 LayerNorm);
    // it saves a full extra sweep over the accumulator in TMEM.
    stat[h * 128 + lrow] = sum;
    stat[256 + h * 128 + lrow] = sumsq;
    asm volatile("bar.sync 3, 256;" ::: "memory");
    const float mean = (sum + stat[(h ^ 1) * 128 + lrow]) * (1.0f / Cfg::N);
    const float ex2 = (sumsq + stat[256 + (h ^ 1) * 128 + lrow]) * (1.0f / Cfg::N);
    const float rstd = rsqrtf(fmaxf(ex2 - mean * mean, 0.f) + p.eps);
    LN_TRACE(3);
    // ---- sweep 3: normalise -> 16-bit slabs (64 columns each), optional fp32 copy ----
    // Sweep-1 store groups used buffers {0,1} {2,3} {4,5} {0,1}.  Without the fp32 copy a step needs one buffer, so
    // step 0 takes buffer 3 and only the last group may still be reading; with the copy it needs three: drain all.
    if (ht == 0) {
      if (p.want_f32) bulk_wait_read<0>(); else bulk_wait_read<1>();
    }
    auto sweep3 = [&](int col, const uint32_t (&r)[32], uint8_t* slab, int chunk0, uint8_t* fs) {
      const float4* g4 = reinterpret_cast<const float4*>(s_gamma + col);
      const float4* be4 = reinterpret_cast<const float4*>(s_beta + col);
      float y[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 g = g4[c], be = be4[c];
        y[4 * c + 0] = fmaf((__uint_as_float(r[4 * c + 0]) - mean) * rstd, g.x, be.x);
        y[4 * c + 1] = fmaf((__uint_as_float(r[4 * c + 1]) - mean) * rstd, g.y, be.y);
        y[4 * c + 2] = fmaf((__uint_as_float(r[4 * c + 2]) - mean) * rstd, g.z, be.z);
        y[4 * c + 3] = fmaf((__uint_as_float(r[4 * c + 3]) - mean) * rstd, g.w, be.w);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 q;
        q.x = OpTraits<OpT>::pack2(y[8 * c + 0], y[8 * c + 1]);
        q.y = OpTraits<OpT>::pack2(y[8 * c + 2], y[8 * c + 3]);
        q.z = OpTraits<OpT>::pack2(y[8 * c + 4], y[8 * c + 5]);
        q.w = OpTraits<OpT>::pack2(y[8 * c + 6], y[8 * c + 7]);
        *ln_slab_chunk(slab, lrow, chunk0 + c) = q;
      }
      if (p.want_f32) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *ln_slab_chunk(fs, lrow, c) = make_uint4(__float_as_uint(y[4 * c]), __float_as_uint(y[4 * c + 1]),
                                                   __float_as_uint(y[4 * c + 2]), __float_as_uint(y[4 * c + 3]));
      }
    };
    tmem_ld32(taddr, ra);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int sb = 3 * ((j + 1) & 1);       // buffers {3,4,5} / {0,1,2} alternate
      uint8_t* slab = slab_ptr(sb);
      uint8_t* f0 = slab_ptr(sb + 1);
      uint8_t* f1 = slab_ptr(sb + 2);
      if (j == 1) LN_TRACE(7);
      if (ht == 0) bulk_wait_read<1>();       // the group that used these buffers two iterations ago has drained
      if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (j == 1) LN_TRACE(8);
      tmem_ld_wait();
      tmem_ld32(taddr + 64 * j + 32, rb);
      if (j == 1) LN_TRACE(9);
      sweep3(256 * h + 64 * j, ra, slab, 0, f0);
      if (j == 1) LN_TRACE(10);
      tmem_ld_wait();
      if (j + 1 < 4) tmem_ld32(taddr + 64 * (j + 1), ra);
      sweep3(256 * h + 64 * j + 32, rb, slab, 4, f1);
      if (j == 1) LN_TRACE(11);
      fence_proxy_async_smem();
      if (h == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (j == 1) LN_TRACE(12);
      if (ht == 0) {
        tma_store_3d(&mapH, slab, 256 * h + 64 * j, row0, 0);
        if (p.want_f32) {
          tma_store_3d(&mapHf, f0, 256 * h + 64 * j, row0, 0);
          tma_store_3d(&mapHf, f1, 256 * h + 64 * j + 32, row0, 0);
        }
        bulk_commit();
      }
    }
    // the operand ring and TMEM are handed back to the mainloop of the next tile (both CTAs)
    LN_TRACE(4);
    if (ht == 0) bulk_wait_read<0>();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    LN_TRACE(5);
  }
  if ((threadIdx.x & 127) == 0) bulk_wait<0>();
#ifdef MM_LN_TRACE
  if (ht == 0) g_ln_trace[(blockIdx.x * 2 + h) * 16 + 6] = clock64();
#endif

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <typename OpT>
static int launch_ln_gemm(const CUtensorMap& mA, const CUtensorMap& mW, const CUtensorMap& mX, const CUtensorMap& mH,
                          const CUtensorMap& mHf, const CUtensorMap& mXo, const LnDev& p, cudaStream_t s) {
  auto kern = gemm_resid_ln_kernel<OpT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LnCfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(gemm_resid_ln)");
    attr_set = true;
  }
  const int max_pairs = kNumSMs / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = LnCfg::SMEM_BYTES;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA, mW, mX, mH, mHf, mXo, p);
  if (e != cudaSuccess) return fail(e, "gemm_resid_ln_kernel launch");
  return 0;
}

}  // namespace mm

#ifdef MM_LN_TRACE
extern "C" int mm_debug_ln_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, mm::g_ln_trace, sizeof(long long) * 148 * 2 * 16);
}
#endif

extern "C" int mm_gemm_resid_ln(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int32_t rows, int32_t k,
                                int32_t n, const float* bias, float* x, const float* gamma, const float* beta,
                                float eps, void* h_op, float* h_f32, int32_t dtype, void* stream) {
  return mm_gemm_resid_ln_out(a, a_ld, w, w_ld, rows, k, n, bias, x, x, gamma, beta, eps, h_op, h_f32, dtype, stream);
}

extern "C" int mm_gemm_resid_ln_out(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int32_t rows, int32_t k,
                                    int32_t n, const float* bias, const float* x, float* x_out, const float* gamma,
                                    const float* beta, float eps, void* h_op, float* h_f32, int32_t dtype,
                                    void* stream) {
  return mm_gemm_resid_ln_drop(a, a_ld, w, w_ld, rows, k, n, bias, x, x_out, gamma, beta, eps, h_op, h_f32, 0.f, 0, nullptr,
                               0, dtype, stream);
}

extern "C" int mm_gemm_resid_ln_drop(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int32_t rows, int32_t k,
                                     int32_t n, const float* bias, const float* x, float* x_out, const float* gamma,
                                     const float* beta, float eps, void* h_op, float* h_f32, float drop_p, uint64_t seed,
                                     const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream) {
  using namespace mm;
  if (drop_p < 0.f || drop_p >= 1.f) return bad_arg("gemm_resid_ln: dropout p in [0, 1)");
  if (!x_out) return bad_arg("gemm_resid_ln: null x_out");
  if (!a || !w || !bias || !x || !gamma || !beta || !h_op) return bad_arg("gemm_resid_ln: null pointer");
  if (n != LnCfg::N) return bad_arg("gemm_resid_ln: n must be 512 (full rows in one accumulator)");
  if (rows <= 0 || k <= 0) return bad_arg("gemm_resid_ln: extents");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap mA, mW, mX, mH, mHf, mXo;
  int rc = make_tmap_3d(&mA, a, f16, (uint64_t)k, (uint64_t)rows, 1, (uint64_t)a_ld, 0, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mW, w, f16, (uint64_t)k, (uint64_t)n, 1, (uint64_t)w_ld, 0, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mX, x, 2, (uint64_t)n, (uint64_t)rows, 1, (uint64_t)n, 0, 32, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mXo, x_out, 2, (uint64_t)n, (uint64_t)rows, 1, (uint64_t)n, 0, 32, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mH, h_op, f16 ? 1 : 0, (uint64_t)n, (uint64_t)rows, 1, (uint64_t)n, 0, 64, 128);
  if (rc) return rc;
  if (h_f32) {
    rc = make_tmap_3d_ex(&mHf, h_f32, 2, (uint64_t)n, (uint64_t)rows, 1, (uint64_t)n, 0, 32, 128);
    if (rc) return rc;
  } else {
    mHf = mX;
  }
  LnDev p;
  memset(&p, 0, sizeof(p));
  p.rows = rows, p.k = k, p.num_kb = (k + 63) / 64, p.num_tiles = (rows + 255) / 256;
  p.bias = bias, p.gamma = gamma, p.beta = beta, p.eps = eps, p.want_f32 = h_f32 != nullptr;
  p.drop_p = drop_p, p.seed = seed, p.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev), p.site = site;
  static const int l2_hints = getenv("MM_LN_L2_HINTS") ? atoi(getenv("MM_LN_L2_HINTS")) : 3;
  p.l2_hints = l2_hints;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return f16 ? launch_ln_gemm<__half>(mA, mW, mX, mH, mHf, mXo, p, s)
             : launch_ln_gemm<__nv_bfloat16>(mA, mW, mX, mH, mHf, mXo, p, s);
}
