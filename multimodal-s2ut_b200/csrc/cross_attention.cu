// Fused speech -> image attention (flash style) on tcgen05: O = softmax(q K^T + key mask) V for ONE head of width
// d_model (SelectiveAttention, mm_s2ut/models/fuse.py:80-113, constructed with num_heads = 1 at
// mm_s2s_transformer.py:132-137; MultimodalAttention, fuse.py:145-167, whose learned bias_k / bias_v is simply key /
// value number Tk - 1 of every utterance).  The scores and the probabilities never leave the SM: S lives in TMEM,
// P goes back to TMEM as packed 16-bit over the consumed S columns and feeds the P V UMMA from there.
//
// The head is d_model wide (512), so a 128-query tile needs a 128 x 512 fp32 output accumulator = all 512 TMEM
// columns and S would not fit beside it.  An item is therefore (utterance, 128-query tile, 256-column block of O):
// S (128 queries x 128 keys, contraction over all of d_model) is double-buffered in TMEM columns [0,256), the item's
// O block sits in [256,512).  Column blocks of one query tile run on neighbouring SMs at the same time and share the
// Q / K tiles through L2; the redundant S UMMAs cost less than the S / P round trip through L2 / HBM they replace
// (37 MB fp32 + 19 MB at the bench shape, three launches).
//
// Roles (320 threads, one CTA per SM, persistent over the item list):
//   warp 8      TMA producer: one 4-stage ring of 32 KB stages; an S stage = Q k-block (128 x 64) | K_c k-block
//               (128 keys x 64), a P V stage = two {64 columns x 128 keys} blocks of V_c (MN-major B operand)
//   warp 9      MMA issuer:   S_{g+1} = Q K_{g+1}^T (SS, M128 N128) is issued BEFORE P_g V_g (TS, two N = 128 halves),
//               so the tensor pipe computes the next chunk's scores while the softmax warps work on this one
//   warps 0-7   online softmax, two threads per query row (64 keys each); the running maximum is only moved when it
//               grew by more than 2^8 (the final 1 / l normalisation makes that exact), so the O accumulator is
//               rescaled in TMEM a handful of times per row at most; epilogue: O / l -> 16-bit -> swizzled staging
//               slabs -> TMA store; optional log-sum-exp per query row for a backward pass.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int XA_BM = 128;                       // query rows per item
constexpr int XA_KC = 128;                       // keys per chunk
constexpr int XA_NB = 256;                       // output columns per item
constexpr int XA_TILE_BYTES = 128 * 64 * 2;      // every TMA box is {64 elements, 128 rows} of 16-bit: 16 KB
constexpr int XA_STAGE_BYTES = 2 * XA_TILE_BYTES;
constexpr int XA_STAGES = 4;
constexpr int XA_OUT_BYTES = 4 * XA_TILE_BYTES;  // 4 slabs of 64 output columns
constexpr int XA_XCH_FLOATS = 2 * 2 * XA_BM + 2 * XA_BM;   // [chunk parity][key half][row] maxima + [key half][row] sums
constexpr int XA_SMEM_BYTES = XA_STAGES * XA_STAGE_BYTES + XA_OUT_BYTES + 256 + XA_XCH_FLOATS * 4 + 1024;
constexpr int XA_THREADS = 320;
static_assert(XA_SMEM_BYTES <= 232448, "shared memory budget");

__device__ __forceinline__ float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// maximum of the scores whose bit is set
__device__ __forceinline__ float xa_max32(const uint32_t (&r)[32], uint32_t bits) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  if (bits == 0xffffffffu) {
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      m0 = fmaxf(m0, __uint_as_float(r[k])), m1 = fmaxf(m1, __uint_as_float(r[k + 1]));
      m2 = fmaxf(m2, __uint_as_float(r[k + 2])), m3 = fmaxf(m3, __uint_as_float(r[k + 3]));
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      m0 = fmaxf(m0, (bits >> k) & 1u ? __uint_as_float(r[k]) : -INFINITY);
      m1 = fmaxf(m1, (bits >> (k + 1)) & 1u ? __uint_as_float(r[k + 1]) : -INFINITY);
      m2 = fmaxf(m2, (bits >> (k + 2)) & 1u ? __uint_as_float(r[k + 2]) : -INFINITY);
      m3 = fmaxf(m3, (bits >> (k + 3)) & 1u ? __uint_as_float(r[k + 3]) : -INFINITY);
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// p = exp2(s log2e - mb) for the scores whose bit is set (0 elsewhere) -> 16 packed pairs; returns their sum
template <typename OpT>
__device__ __forceinline__ float xa_exp32(const uint32_t (&r)[32], uint32_t (&pk)[16], float mb, uint32_t bits) {
  constexpr float L2E = 1.4426950408889634f;
  float l0 = 0.f, l1 = 0.f;
  if (bits == 0xffffffffu) {
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      const float p0 = xa_ex2(fmaf(__uint_as_float(r[k]), L2E, -mb));
      const float p1 = xa_ex2(fmaf(__uint_as_float(r[k + 1]), L2E, -mb));
      pk[k >> 1] = OpTraits<OpT>::pack2(p0, p1);
      l0 += p0, l1 += p1;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
      const float p0 = (bits >> k) & 1u ? xa_ex2(fmaf(__uint_as_float(r[k]), L2E, -mb)) : 0.f;
      const float p1 = (bits >> (k + 1)) & 1u ? xa_ex2(fmaf(__uint_as_float(r[k + 1]), L2E, -mb)) : 0.f;
      pk[k >> 1] = OpTraits<OpT>::pack2(p0, p1);
      l0 += p0, l1 += p1;
    }
  }
  return l0 + l1;
}

struct XaDev {
  int Tq, Tk, d, k_col0, v_col0;
  int nqt, ncb, n_items, num_kb, nc;
  const uint8_t* key_mask;   // [batch][mask_ld], non-zero = key masked out; may be NULL
  long long mask_ld;
  float* lse;                // [batch * Tq] natural-log sum-exp of every query row's scores; may be NULL
};

template <typename OpT>
__global__ void __launch_bounds__(XA_THREADS, 1)
cross_attention_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                       const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapOut,
                       const XaDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sRing = smem;
  uint8_t* sOut = sRing + XA_STAGES * XA_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + XA_OUT_BYTES);
  uint64_t* full = bars;                  // [STAGES] TMA -> MMA
  uint64_t* empty = full + XA_STAGES;     // [STAGES] MMA -> TMA
  uint64_t* s_full = empty + XA_STAGES;   // [2] S chunk in TMEM buffer (g & 1) -> softmax warps
  uint64_t* p_full = s_full + 2;          // [1] 8 softmax warps: P stored (and O rescaled) -> MMA
  uint64_t* o_done = p_full + 1;          // [1] P V of a chunk complete -> softmax warps
  uint64_t* o_free = o_done + 1;          // [1] 8 softmax warps hold the item's O in registers -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  float* x_max = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][2][128]
  float* x_sum = x_max + 2 * 2 * XA_BM;                                               // [2][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int G = n_local * p.nc;           // chunks this CTA walks, in order (item-major)
  constexpr float L2E = 1.4426950408889634f;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapOut);
    for (int i = 0; i < XA_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 8);
    mbar_init(o_done, 1);
    mbar_init(o_free, 8);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // the q / k|v projections (previous kernels) are complete and visible

  auto item_coords = [&](int i, int& cb, int& qt, int& b) {
    const int item = blockIdx.x + i * gridDim.x;
    cb = item % p.ncb, qt = (item / p.ncb) % p.nqt, b = item / (p.ncb * p.nqt);
  };

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0 && G > 0) {
      uint32_t n = 0;   // stage fills issued so far
      auto acquire = [&]() -> uint8_t* {
        const uint32_t st = n % XA_STAGES;
        mbar_wait(&empty[st], ((n / XA_STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], XA_STAGE_BYTES);
        return sRing + st * XA_STAGE_BYTES;
      };
      auto fill_s = [&](int g) {
        int cb, qt, b;
        item_coords(g / p.nc, cb, qt, b);
        const int c = g % p.nc;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          uint8_t* st = acquire();
          uint64_t* bar = &full[n % XA_STAGES];
          tma_load_3d(st, &mapQ, bar, kb * 64, qt * XA_BM, b);
          tma_load_3d(st + XA_TILE_BYTES, &mapK, bar, p.k_col0 + kb * 64, c * XA_KC, b);
          ++n;
        }
      };
      auto fill_v = [&](int g) {
        int cb, qt, b;
        item_coords(g / p.nc, cb, qt, b);
        const int c = g % p.nc;
        for (int j = 0; j < 2; ++j) {
          uint8_t* st = acquire();
          uint64_t* bar = &full[n % XA_STAGES];
          const int col = p.v_col0 + cb * XA_NB + j * 128;
          tma_load_3d(st, &mapV, bar, col, c * XA_KC, b);
          tma_load_3d(st + XA_TILE_BYTES, &mapV, bar, col + 64, c * XA_KC, b);
          ++n;
        }
      };
      fill_s(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) fill_s(g + 1);
        fill_v(g);
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0 && G > 0) {
      constexpr uint32_t idesc_s = umma_idesc(XA_BM, XA_KC, OpTraits<OpT>::fmt);
      constexpr uint32_t idesc_o = umma_idesc(XA_BM, 128, OpTraits<OpT>::fmt) | (1u << 16);   // B (= V) is MN-major
      uint32_t n = 0;
      auto issue_s = [&](int g) {
        const uint32_t t_s = tmem_base + 128 * (g & 1);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const uint32_t st = n % XA_STAGES;
          mbar_wait(&full[st], (n / XA_STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(sRing + st * XA_STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(base), bdesc = umma_desc_sw128(base + XA_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(t_s, adesc + 2 * kk, bdesc + 2 * kk, idesc_s, (kb | kk) != 0);
          umma_commit(&empty[st]);
          ++n;
        }
        umma_commit(&s_full[g & 1]);
      };
      auto issue_pv = [&](int g) {
        const uint32_t t_p = tmem_base + 128 * (g & 1);
        const int c = g % p.nc;
        for (int j = 0; j < 2; ++j) {
          const uint32_t st = n % XA_STAGES;
          mbar_wait(&full[st], (n / XA_STAGES) & 1);
          tc_fence_after();
          const uint64_t vdesc = umma_desc_sw128_mn(smem_u32(sRing + st * XA_STAGE_BYTES), XA_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // 16 keys per step: 8 packed P columns, 16 V rows of 128 B
            umma_f16_ts(tmem_base + 256 + 128 * j, t_p + 8 * ks, vdesc + 128ull * ks, idesc_o, (c | ks) != 0);
          umma_commit(&empty[st]);
          ++n;
        }
        umma_commit(o_done);
      };
      issue_s(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) issue_s(g + 1);     // buffer (g + 1) & 1 last held P_{g-1}, read by the P V issued before
        mbar_wait(p_full, g & 1);
        const int i = g / p.nc;
        if (g % p.nc == 0 && i > 0) mbar_wait(o_free, (i - 1) & 1);   // the previous item's O has been read out
        tc_fence_after();
        issue_pv(g);
      }
    }
  } else {
    // ---------------- softmax warps + epilogue ----------------
    const int hf = warp >> 2;                  // which 64 keys of a chunk / which 128 output columns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_o = tmem_base + 256 + t_lane + 128 * hf;
    auto group_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const bool elected = tid == 0;
    float m_run = -INFINITY, m_ref = -INFINITY, l = 0.f;
    for (int g = 0; g < G; ++g) {
      const int i = g / p.nc, c = g % p.nc;
      int cb, qt, b;
      item_coords(i, cb, qt, b);
      if (c == 0) m_run = -INFINITY, m_ref = -INFINITY, l = 0.f;
      // which of my 64 keys exist / are visible
      const int k0 = c * XA_KC + 64 * hf;
      const int nv = min(64, max(0, p.Tk - k0));
      uint32_t bits0 = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
      uint32_t bits1 = nv >= 64 ? 0xffffffffu : (nv > 32 ? ((1u << (nv - 32)) - 1u) : 0u);
      if (p.key_mask != nullptr) {
        const uint8_t* mk = p.key_mask + (long long)b * p.mask_ld + k0;
        for (int k = 0; k < nv; ++k)
          if (mk[k]) {
            if (k < 32) bits0 &= ~(1u << k); else bits1 &= ~(1u << (k - 32));
          }
      }
      const uint32_t t_s = tmem_base + 128 * (g & 1) + t_lane;
      mbar_wait(&s_full[g & 1], (g >> 1) & 1);
      tc_fence_after();
      uint32_t ra[32], rb[32];
      tmem_ld32(t_s + 64 * hf, ra);
      tmem_ld32(t_s + 64 * hf + 32, rb);
      tmem_ld_wait();
      const float mloc = fmaxf(xa_max32(ra, bits0), xa_max32(rb, bits1));
      float* xm = x_max + (g & 1) * 2 * XA_BM;
      xm[hf * XA_BM + row] = mloc;
      group_sync();                            // both halves hold their S in registers: P may overwrite it
      m_run = fmaxf(m_run, fmaxf(mloc, xm[(hf ^ 1) * XA_BM + row]));
      float alpha = 1.0f;
      if (c == 0) {
        m_ref = m_run;
      } else if ((m_run - m_ref) * L2E > 8.0f) {     // lazy rescale: only when the maximum grew by more than 2^8
        alpha = xa_ex2((m_ref - m_run) * L2E);       // (0 when m_ref was -inf)
        m_ref = m_run;
      }
      const float mb = m_ref * L2E;
      uint32_t pk[16];
      float lc = xa_exp32<OpT>(ra, pk, mb, bits0);
      tmem_st16(t_s + 32 * hf, pk);
      lc += xa_exp32<OpT>(rb, pk, mb, bits1);
      tmem_st16(t_s + 32 * hf + 16, pk);
      l = l * alpha + lc;
      if (c > 0) {
        // O <- alpha O in the warps where some row's reference maximum moved; P V of the previous chunk has landed
        mbar_wait(o_done, (g - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            tmem_ld32(t_o + 32 * q, ra);
            tmem_ld_wait();
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              lo[k] = __float_as_uint(__uint_as_float(ra[k]) * alpha);
              hi[k] = __float_as_uint(__uint_as_float(ra[16 + k]) * alpha);
            }
            tmem_st16(t_o + 32 * q, lo);
            tmem_st16(t_o + 32 * q + 16, hi);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (c + 1 < p.nc) continue;

      // ---- item epilogue: O / l -> 16-bit -> staging slabs -> TMA store ----
      x_sum[hf * XA_BM + row] = l;
      mbar_wait(o_done, g & 1);
      tc_fence_after();
      if (elected) bulk_wait_read<0>();        // the previous item's stores have finished reading the slabs
      group_sync();                            // ... and both halves of the row sum are visible
      const float lt = l + x_sum[(hf ^ 1) * XA_BM + row];
      const float inv = lt > 0.f ? 1.0f / lt : 0.f;
      if (p.lse != nullptr && hf == 0 && cb == 0 && qt * XA_BM + row < p.Tq)
        p.lse[(long long)b * p.Tq + qt * XA_BM + row] = m_ref + __logf(lt);
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {            // my 128 columns = slabs 2 hf and 2 hf + 1
        tmem_ld32(t_o + 64 * s, ra);
        tmem_ld32(t_o + 64 * s + 32, rb);
        tmem_ld_wait();
        if (s == 1) {                          // O is in registers: the next item's P V may overwrite the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
        uint8_t* so = sOut + (2 * hf + s) * XA_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u, v;
          u.x = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 0]) * inv, __uint_as_float(ra[8 * k + 1]) * inv);
          u.y = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 2]) * inv, __uint_as_float(ra[8 * k + 3]) * inv);
          u.z = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 4]) * inv, __uint_as_float(ra[8 * k + 5]) * inv);
          u.w = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 6]) * inv, __uint_as_float(ra[8 * k + 7]) * inv);
          v.x = OpTraits<OpT>::pack2(__uint_as_float(rb[8 * k + 0]) * inv, __uint_as_float(rb[8 * k + 1]) * inv);
          v.y = OpTraits<OpT>::pack2(__uint_as_float(rb[8 * k + 2]) * inv, __uint_as_float(rb[8 * k + 3]) * inv);
          v.z = OpTraits<OpT>::pack2(__uint_as_float(rb[8 * k + 4]) * inv, __uint_as_float(rb[8 * k + 5]) * inv);
          v.w = OpTraits<OpT>::pack2(__uint_as_float(rb[8 * k + 6]) * inv, __uint_as_float(rb[8 * k + 7]) * inv);
          *reinterpret_cast<uint4*>(so + row * 128 + ((k ^ (row & 7)) << 4)) = u;
          *reinterpret_cast<uint4*>(so + row * 128 + (((4 + k) ^ (row & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      group_sync();
      if (elected) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          tma_store_3d(&mapOut, sOut + s * XA_TILE_BYTES, cb * XA_NB + 64 * s, qt * XA_BM, b);
        bulk_commit();
      }
    }
    if (elected) bulk_wait<0>();   // the last stores have landed before the CTA's smem goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename OpT>
static int launch_cross_attention(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                                  const CUtensorMap& mout, const XaDev& p, cudaStream_t s) {
  auto kern = cross_attention_kernel<OpT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(cross_attention)");
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.n_items < kNumSMs ? p.n_items : kNumSMs);
  cfg.blockDim = dim3(XA_THREADS);
  cfg.dynamicSmemBytes = XA_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, mout, p);
  if (e != cudaSuccess) return fail(e, "cross_attention_kernel launch");
  return 0;
}

}  // namespace mm

using namespace mm;

extern "C" int mm_cross_attention(const void* q, int64_t q_ld, int32_t q_len, const void* k, int64_t k_ld,
                                  int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                                  int64_t kv_batch_stride, const uint8_t* key_mask, int64_t mask_ld, int32_t batch,
                                  int32_t d_model, void* out, int64_t out_ld, float* lse, int32_t dtype,
                                  void* stream) {
  if (!q || !k || !v || !out) return bad_arg("cross_attention: null pointer");
  if (batch <= 0 || q_len <= 0 || kv_len <= 0) return bad_arg("cross_attention: extents");
  if (d_model <= 0 || d_model % XA_NB) return bad_arg("cross_attention: d_model must be a multiple of 256");
  if ((q_ld % 8) || (k_ld % 8) || (v_ld % 8) || (out_ld % 8) || q_ld < d_model || out_ld < d_model ||
      k_ld < k_col0 + d_model || v_ld < v_col0 + d_model || (k_col0 % 8) || (v_col0 % 8) || (kv_batch_stride % 8))
    return bad_arg("cross_attention: leading dims / column offsets");
  if (key_mask && mask_ld < kv_len) return bad_arg("cross_attention: mask_ld < kv_len");
  const int f16 = dtype == MM_DTYPE_F16;
  const uint64_t k_bs = kv_batch_stride > 0 ? (uint64_t)kv_batch_stride : (uint64_t)kv_len * k_ld;
  const uint64_t v_bs = kv_batch_stride > 0 ? (uint64_t)kv_batch_stride : (uint64_t)kv_len * v_ld;
  CUtensorMap mq, mk, mv, mout;
  int rc = make_tmap_3d(&mq, q, f16, (uint64_t)d_model, (uint64_t)q_len, (uint64_t)batch, (uint64_t)q_ld,
                        (uint64_t)q_len * q_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mk, k, f16, (uint64_t)k_ld, (uint64_t)kv_len, (uint64_t)batch, (uint64_t)k_ld, k_bs, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mv, v, f16, (uint64_t)v_ld, (uint64_t)kv_len, (uint64_t)batch, (uint64_t)v_ld, v_bs, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mout, out, f16, (uint64_t)d_model, (uint64_t)q_len, (uint64_t)batch, (uint64_t)out_ld,
                    (uint64_t)q_len * out_ld, 128);
  if (rc) return rc;
  XaDev p;
  memset(&p, 0, sizeof(p));
  p.Tq = q_len, p.Tk = kv_len, p.d = d_model, p.k_col0 = k_col0, p.v_col0 = v_col0;
  p.nqt = (q_len + XA_BM - 1) / XA_BM;
  p.ncb = d_model / XA_NB;
  p.n_items = batch * p.nqt * p.ncb;
  p.num_kb = d_model / 64;
  p.nc = (kv_len + XA_KC - 1) / XA_KC;
  p.key_mask = key_mask, p.mask_ld = mask_ld, p.lse = lse;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return f16 ? launch_cross_attention<__half>(mq, mk, mv, mout, p, s)
             : launch_cross_attention<__nv_bfloat16>(mq, mk, mv, mout, p, s);
}
