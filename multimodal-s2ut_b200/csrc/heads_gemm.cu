// Per-(sequence, head) contraction with a 64-wide output (head_dim = 64), the shape of attention backward's three
// output products:
//
//     dV = P^T dO      dK = dS^T q      dQ = dS k * head_dim^-0.5          (per sequence b and head h)
//
//   out[b][r][out_col0 + 64 h + c] = scale * sum_j A_bh[r, j] * W[b][j][w_col0 + 64 h + c]
//
// A_bh is the [rows, k] probability / score-gradient matrix of one (b, h) (P or dS as mm_attention_bwd_scores wrote
// them, read as stored or transposed), W is q, k or dO straight from the token-major [tokens, heads * 64] layout and the
// result lands in the q | k | v gradient layout.  gemm.cu ran these as 256 x 256 pair tiles of which 64 columns were
// real (the other 192 were the next heads' columns: 4x the tensor work and 4x the W traffic, 150 TFLOP/s); here a tile
// is 128 rows x 64 columns:
//   warp 0   TMA producer: A k-block (16 KB) + the head's W k-block (one {64 features, 64 tokens} box, 8 KB), 4-stage ring
//   warp 1   MMA issuer: tcgen05.mma cta_group::1, M = 128, N = 64, K = 16; A K-major or MN-major, W MN-major;
//            2 x 64 TMEM columns double-buffer the accumulator
//   warps 2-5 epilogue: tcgen05.ld -> scale -> 16-bit -> swizzled slab -> TMA store (clipped at the tensor bounds)
// 113 KB of shared memory: two CTAs per SM, so one CTA's loads overlap the other's epilogue.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

struct HgCfg {
  static constexpr int BM = 128, BN = 64, BK = 64, STAGES = 4;
  static constexpr int A_BYTES = BM * BK * 2;      // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;      // 8 KB
  static constexpr int SLAB_BYTES = BM * 128;      // 16 KB: 128 rows x 64 16-bit columns
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + SLAB_BYTES + 256 + 1024;
  static constexpr int THREADS = 192;
};
static_assert(2 * HgCfg::SMEM_BYTES <= 227 * 1024, "two CTAs per SM");

struct HgDev {
  int rows, k, num_kb, tail_steps, m_tiles, num_tiles, heads, a_mn, w_col0, out_col0;
  float scale;
};

template <typename OpT>
__global__ void __launch_bounds__(HgCfg::THREADS, 2)
heads_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                  const __grid_constant__ CUtensorMap mapOut, const HgDev p) {
  using Cfg = HgCfg;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * Cfg::A_BYTES;
  uint8_t* sSlab = sB + STAGES * Cfg::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlab + Cfg::SLAB_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = full + STAGES;       // [STAGES]
  uint64_t* tfull = empty + STAGES;      // [2]
  uint64_t* tempty = tfull + 2;          // [2] 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapOut);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * Cfg::BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int bh = tile / p.m_tiles, row0 = (tile - bh * p.m_tiles) * Cfg::BM;
        const int b = bh / p.heads, h = bh - b * p.heads;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          uint8_t* dA = sA + stage * Cfg::A_BYTES;
          if (p.a_mn) {   // memory [j][r]: two {64 rows, 64 contraction indices} boxes
            tma_load_3d(dA, &mapA, &full[stage], row0, kb * Cfg::BK, bh);
            tma_load_3d(dA + Cfg::A_BYTES / 2, &mapA, &full[stage], row0 + 64, kb * Cfg::BK, bh);
          } else {
            tma_load_3d(dA, &mapA, &full[stage], kb * Cfg::BK, row0, bh);
          }
          tma_load_3d(sB + stage * Cfg::B_BYTES, &mapW, &full[stage], p.w_col0 + h * Cfg::BN, kb * Cfg::BK, b);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(Cfg::BM, Cfg::BN, OpTraits<OpT>::fmt) | (p.a_mn ? (1u << 15) : 0u) | (1u << 16);
      const uint64_t a_step = p.a_mn ? 128 : 2;
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t aaddr = smem_u32(sA + stage * Cfg::A_BYTES), baddr = smem_u32(sB + stage * Cfg::B_BYTES);
          const uint64_t adesc = p.a_mn ? umma_desc_sw128_mn(aaddr, Cfg::A_BYTES / 2) : umma_desc_sw128(aaddr);
          const uint64_t bdesc = umma_desc_sw128_mn(baddr, Cfg::B_BYTES);
          const int steps = (kb == p.num_kb - 1) ? p.tail_steps : 4;
          for (int kk = 0; kk < steps; ++kk)
            umma_f16(tmem_d, adesc + a_step * kk, bdesc + 128ull * kk, idesc, (kb | kk) != 0);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        umma_commit(&tfull[as]);
        if (++as == 2) as = 0, aphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5: TMEM lane quadrant = warp & 3) =====================
    const int quad = warp & 3, lrow = quad * 32 + lane;
    const bool elected = threadIdx.x == 64;
    uint32_t as = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int bh = tile / p.m_tiles, row0 = (tile - bh * p.m_tiles) * Cfg::BM;
      const int b = bh / p.heads, h = bh - b * p.heads;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * Cfg::BN + (static_cast<uint32_t>(quad * 32) << 16);
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      if (elected) bulk_wait_read<0>();                         // the previous tile's store has read the slab
      asm volatile("bar.sync 1, 128;" ::: "memory");
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);                  // accumulator drained: the next tile's MMAs may start
      if (++as == 2) as = 0, aphase ^= 1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t* r = i < 4 ? &r0[8 * i] : &r1[8 * (i - 4)];
        uint4 q;
        q.x = OpTraits<OpT>::pack2(__uint_as_float(r[0]) * p.scale, __uint_as_float(r[1]) * p.scale);
        q.y = OpTraits<OpT>::pack2(__uint_as_float(r[2]) * p.scale, __uint_as_float(r[3]) * p.scale);
        q.z = OpTraits<OpT>::pack2(__uint_as_float(r[4]) * p.scale, __uint_as_float(r[5]) * p.scale);
        q.w = OpTraits<OpT>::pack2(__uint_as_float(r[6]) * p.scale, __uint_as_float(r[7]) * p.scale);
        *reinterpret_cast<uint4*>(sSlab + lrow * 128 + ((i ^ (lrow & 7)) << 4)) = q;
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (elected) {
        tma_store_3d(&mapOut, sSlab, p.out_col0 + h * Cfg::BN, row0, b);
        bulk_commit();
      }
    }
    if (elected) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * Cfg::BN);
  }
}

template <typename OpT>
static int launch_heads_gemm(const CUtensorMap& mA, const CUtensorMap& mW, const CUtensorMap& mO, const HgDev& p,
                             cudaStream_t s) {
  auto kern = heads_gemm_kernel<OpT>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HgCfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(heads_gemm)");
    attr_set = true;
  }
  const int max_ctas = 2 * kNumSMs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.num_tiles < max_ctas ? p.num_tiles : max_ctas);
  cfg.blockDim = dim3(HgCfg::THREADS);
  cfg.dynamicSmemBytes = HgCfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA, mW, mO, p);
  if (e != cudaSuccess) return fail(e, "heads_gemm_kernel launch");
  return 0;
}

}  // namespace mm

extern "C" int mm_heads_gemm(const void* a, int64_t a_ld, int64_t a_bs, int32_t a_transposed, const void* w, int64_t w_ld,
                             int64_t w_bs, int32_t w_col0, void* out, int64_t out_ld, int64_t out_bs, int32_t out_col0,
                             int32_t rows, int32_t k, int32_t batch, int32_t heads, float scale, int32_t dtype,
                             void* stream) {
  using namespace mm;
  if (!a || !w || !out) return bad_arg("heads_gemm: null pointer");
  if (rows <= 0 || k <= 0 || batch <= 0 || heads <= 0) return bad_arg("heads_gemm: extents");
  if (dtype != MM_DTYPE_BF16 && dtype != MM_DTYPE_F16) return bad_arg("heads_gemm: dtype");
  const int d = heads * HgCfg::BN;
  if (w_col0 % 8 || out_col0 % 8 || w_ld < w_col0 + d || out_ld < out_col0 + d)
    return bad_arg("heads_gemm: column offsets / leading dims (head_dim must be 64)");
  const int kind = dtype == MM_DTYPE_F16 ? 1 : 0;
  const uint64_t bh = (uint64_t)batch * heads;
  CUtensorMap mA, mW, mO;
  int rc;
  if (a_transposed)   // memory [bh][j][r]
    rc = make_tmap_3d_ex(&mA, a, kind, (uint64_t)rows, (uint64_t)k, bh, (uint64_t)a_ld, (uint64_t)a_bs, 64, 64);
  else                // memory [bh][r][j]
    rc = make_tmap_3d_ex(&mA, a, kind, (uint64_t)k, (uint64_t)rows, bh, (uint64_t)a_ld, (uint64_t)a_bs, 64, 128);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mW, w, kind, (uint64_t)(w_col0 + d), (uint64_t)k, (uint64_t)batch, (uint64_t)w_ld, (uint64_t)w_bs,
                       64, 64);
  if (rc) return rc;
  rc = make_tmap_3d_ex(&mO, out, kind, (uint64_t)(out_col0 + d), (uint64_t)rows, (uint64_t)batch, (uint64_t)out_ld,
                       (uint64_t)out_bs, 64, 128);
  if (rc) return rc;
  HgDev p;
  memset(&p, 0, sizeof(p));
  p.rows = rows, p.k = k, p.heads = heads, p.a_mn = a_transposed != 0, p.w_col0 = w_col0, p.out_col0 = out_col0;
  p.num_kb = (k + 63) / 64;
  p.tail_steps = (k - (p.num_kb - 1) * 64 + 15) >> 4;
  p.m_tiles = (rows + HgCfg::BM - 1) / HgCfg::BM;
  p.num_tiles = (int)bh * p.m_tiles;
  p.scale = scale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return kind ? launch_heads_gemm<__half>(mA, mW, mO, p, s) : launch_heads_gemm<__nv_bfloat16>(mA, mW, mO, p, s);
}
