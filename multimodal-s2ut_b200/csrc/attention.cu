// Multi-head self-attention core on tcgen05: O = softmax_fp32(Q K^T + key-padding mask) V, head_dim 64.
// q, k, v are column blocks of ONE [B*T, 3d] tensor (the QKV GEMM output); q arrives pre-scaled by head_dim^-0.5
// (fused into that GEMM's epilogue).  Reference: fairseq/modules/multihead_attention.py (encoder self-attention,
// key_padding_mask), as called from TransformerEncoderLayerBase.forward.
//
// Two kernels:
//  * self_attention_t256_kernel (T <= 256, i.e. utterances up to ~10 s at 4x subsampling): persistent, warp-specialised,
//    S / P / O all live in TMEM; described in front of the kernel below.  This is the one the bench runs.
//  * self_attention_kernel (any T): one CTA (128 threads) per (utterance, head, 128-query tile), keys in chunks of 256:
//      S = Q K^T      tcgen05.mma M=128 N=256 K=64, fp32 scores in TMEM columns [0,256)
//      softmax        thread i owns score row i (TMEM lane i): no cross-thread reductions
//      P              written as 16-bit into shared memory in the 128B-swizzled K-major layout UMMA expects
//      O += P V       tcgen05.mma M=128 N=64 K=256 (A = P from smem, B = V rows as an MN-major operand)
//    with a two-sweep schedule (sweep 1: row maxima only; sweep 2: exp / P V) so O never needs rescaling.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int AT_BM = 128, AT_HD = 64, AT_KC = 256;
constexpr int AT_Q_BYTES = AT_BM * AT_HD * 2;        // 16 KB
constexpr int AT_K_BYTES = AT_KC * AT_HD * 2;        // 32 KB
constexpr int AT_V_BYTES = AT_HD * AT_KC * 2;        // 32 KB (4 blocks of 64 keys)
constexpr int AT_P_BYTES = AT_BM * AT_KC * 2;        // 64 KB (4 blocks of 64 keys)
constexpr int AT_SMEM_BYTES = AT_Q_BYTES + AT_K_BYTES + AT_V_BYTES + AT_P_BYTES + 64 + 1024;
constexpr int AT_TMEM_COLS = 512;
constexpr int AT_O_COL = 256;

template <typename OpT>
__global__ void __launch_bounds__(128, 1)
self_attention_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapVT,
                      const int* __restrict__ seq_lens, int T, int d_model, OpT* __restrict__ out, long long out_ld) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;
  uint8_t* sV = sK + AT_K_BYTES;
  uint8_t* sP = sV + AT_V_BYTES;
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(sP + AT_P_BYTES);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int len = min(seq_lens[b], T);
  const int nchunks = max(1, (len + AT_KC - 1) / AT_KC);
  constexpr float L2E = 1.4426950408889634f;

  if (tid == 0) {
    tma_prefetch_desc(&mapQK);
    tma_prefetch_desc(&mapVT);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, AT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  uint32_t tma_phase = 0, mma_phase = 0;
  constexpr uint32_t idesc_s = umma_idesc(AT_BM, AT_KC, OpTraits<OpT>::fmt);
  constexpr uint32_t idesc_o = umma_idesc(AT_BM, AT_HD, OpTraits<OpT>::fmt) | (1u << 16);  // B (= V) is MN-major

  auto load_and_scores = [&](int c, bool with_q, bool with_v) {
    if (tid == 0) {
      mbar_expect_tx(bar_tma, (with_q ? AT_Q_BYTES : 0) + AT_K_BYTES + (with_v ? AT_V_BYTES : 0));
      if (with_q) tma_load_3d(sQ, &mapQK, bar_tma, h * AT_HD, qt * AT_BM, b);
      tma_load_3d(sK, &mapQK, bar_tma, d_model + h * AT_HD, c * AT_KC, b);
      tma_load_3d(sK + AT_K_BYTES / 2, &mapQK, bar_tma, d_model + h * AT_HD, c * AT_KC + 128, b);
      if (with_v) {   // V rows = keys, 64 values (128 B) each: MN-major B operand of the PV MMA
        tma_load_3d(sV, &mapQK, bar_tma, 2 * d_model + h * AT_HD, c * AT_KC, b);
        tma_load_3d(sV + AT_V_BYTES / 2, &mapQK, bar_tma, 2 * d_model + h * AT_HD, c * AT_KC + 128, b);
      }
    }
    mbar_wait(bar_tma, tma_phase);
    tma_phase ^= 1;
    if (tid == 0) {
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t bdesc = umma_desc_sw128(smem_u32(sK));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc_s, kk != 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  };

  auto row_max_of_chunk = [&](int c, float m) {
#pragma unroll 1
    for (int cc = 0; cc < AT_KC / 32; ++cc) {
      if (c * AT_KC + cc * 32 >= len) break;  // uniform
      uint32_t r[32];
      tmem_ld32(t_row + cc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int key = c * AT_KC + cc * 32 + i;
        if (key < len) m = fmaxf(m, __uint_as_float(r[i]));
      }
    }
    return m;
  };

  float m = -INFINITY, l = 0.f;
  bool q_loaded = false;
  if (nchunks > 1) {
    for (int c = 0; c < nchunks; ++c) {
      load_and_scores(c, !q_loaded, false);
      q_loaded = true;
      m = row_max_of_chunk(c, m);
      tc_fence_before();
      __syncthreads();
    }
  }
  for (int c = 0; c < nchunks; ++c) {
    load_and_scores(c, !q_loaded, true);
    q_loaded = true;
    if (nchunks == 1) m = row_max_of_chunk(0, m);
    const float mb = (m == -INFINITY) ? 0.f : m * L2E;
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < AT_KC / 32; ++cc) {
      uint32_t r[32];
      uint32_t pk[16];
      if (c * AT_KC + cc * 32 < len) {  // uniform
        tmem_ld32(t_row + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int key = c * AT_KC + cc * 32 + i;
          float p0 = key < len ? exp2f(fmaf(__uint_as_float(r[i]), L2E, -mb)) : 0.f;
          float p1 = key + 1 < len ? exp2f(fmaf(__uint_as_float(r[i + 1]), L2E, -mb)) : 0.f;
          pk[i >> 1] = OpTraits<OpT>::pack2(p0, p1);
          l += p0 + p1;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      uint8_t* prow = sP + (cc >> 1) * (AT_P_BYTES / 4) + row * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk16 = (cc & 1) * 4 + q4;
        *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (row & 7)) << 4)) =
            make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(smem_u32(sP));
      const uint64_t bdesc = umma_desc_sw128(smem_u32(sV));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_f16(tmem_base + AT_O_COL, adesc + (uint64_t)((j * (AT_P_BYTES / 4)) >> 4) + 2 * kk,
                   bdesc + (uint64_t)(((j * 64 + kk * 16) * 128) >> 4), idesc_o, (c | j | kk) != 0);
        }
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
  }

  // ---- epilogue: O / l -> 16-bit, 128 contiguous bytes per row ----
  {
    const int row = warp * 32 + lane;
    const int t = qt * AT_BM + row;
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    uint32_t r0[32], r1[32];
    tmem_ld32(t_row + AT_O_COL, r0);
    tmem_ld32(t_row + AT_O_COL + 32, r1);
    tmem_ld_wait();
    if (t < T) {
      uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * T + t) * out_ld + h * AT_HD);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q;
        q.x = OpTraits<OpT>::pack2(__uint_as_float(r0[8 * i + 0]) * inv, __uint_as_float(r0[8 * i + 1]) * inv);
        q.y = OpTraits<OpT>::pack2(__uint_as_float(r0[8 * i + 2]) * inv, __uint_as_float(r0[8 * i + 3]) * inv);
        q.z = OpTraits<OpT>::pack2(__uint_as_float(r0[8 * i + 4]) * inv, __uint_as_float(r0[8 * i + 5]) * inv);
        q.w = OpTraits<OpT>::pack2(__uint_as_float(r0[8 * i + 6]) * inv, __uint_as_float(r0[8 * i + 7]) * inv);
        dst[i] = q;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q;
        q.x = OpTraits<OpT>::pack2(__uint_as_float(r1[8 * i + 0]) * inv, __uint_as_float(r1[8 * i + 1]) * inv);
        q.y = OpTraits<OpT>::pack2(__uint_as_float(r1[8 * i + 2]) * inv, __uint_as_float(r1[8 * i + 3]) * inv);
        q.z = OpTraits<OpT>::pack2(__uint_as_float(r1[8 * i + 4]) * inv, __uint_as_float(r1[8 * i + 5]) * inv);
        q.w = OpTraits<OpT>::pack2(__uint_as_float(r1[8 * i + 6]) * inv, __uint_as_float(r1[8 * i + 7]) * inv);
        dst[4 + i] = q;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

template <typename OpT>
static int launch_attn(const CUtensorMap& mqk, const CUtensorMap& mvt, const int* lens, int B, int T, int H, int d,
                       void* out, long long out_ld, cudaStream_t s) {
  auto kern = self_attention_kernel<OpT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(self_attention)");
    attr_set = true;
  }
  dim3 grid((T + AT_BM - 1) / AT_BM, H, B);
  kern<<<grid, 128, AT_SMEM_BYTES, s>>>(mqk, mvt, lens, T, d, reinterpret_cast<OpT*>(out), out_ld);
  MM_CHECK_LAUNCH("self_attention_kernel launch");
  return 0;
}

// ===================================================================================================
// Persistent, warp-specialised variant for T <= 256 (utterances up to ~10 s: one key chunk per query tile).
//
// One CTA per SM walks a list of (utterance, head, 128-query tile) items.  Roles (576 threads):
//   warp 16     TMA producer: Q / K / V^T of item i into a 2-stage shared-memory ring
//   warp 17     MMA issuer:   S_i = Q K^T (SS) as soon as stage i has landed and TMEM region (i & 1) is free;
//                             O_i = P_i V (TS: P is read from TMEM) once softmax group (i & 1) has published P_i
//   warps 0-7   softmax group 0 (even items)      } two threads per query row (= TMEM lane): warps 0-3 / 8-11 own
//   warps 8-15  softmax group 1 (odd items)       } keys 0-127, warps 4-7 / 12-15 keys 128-255.  Row max and row sum
//                                                   are combined through shared memory; P goes back to TMEM as packed
//                                                   16-bit over the already-consumed S columns, then O / l -> HBM
// The two groups ping-pong, and four softmax warps per scheduler hide the TMEM / MUFU latencies of one another
// (the per-item dependency chain, not bandwidth, bounded the 8-warp version); the producer runs one item ahead.
// TMEM: region g = columns [256 g, 256 g + 256): S (fp32, 256 keys) -> P(keys 0-127) in [0,64), O in [64,128),
// P(keys 128-255) in [128,192).  P V for the first 128 keys is issued as soon as that half of P is in TMEM, while the
// group is still exponentiating the second half.  O leaves through a swizzled 16 KB staging tile and one TMA store
// per item (rows beyond T are clipped by the [B][T][d] tensor map).
// ===================================================================================================
constexpr int PA_STAGE_BYTES = AT_Q_BYTES + AT_K_BYTES + AT_V_BYTES;   // 80 KB
constexpr int PA_MAX_LENS = 1024;   // utterance lengths staged in shared memory (larger batches read them from global)
constexpr int PA_OUT_BYTES = AT_BM * AT_HD * 2;   // 16 KB staging tile per softmax group
constexpr int PA_XCH_BYTES = 2 * 2 * 2 * AT_BM * 4;   // [group][max | sum][key half][row] floats
constexpr int PA_SMEM_BYTES = 2 * PA_STAGE_BYTES + 2 * PA_OUT_BYTES + 256 + PA_MAX_LENS * 4 + PA_XCH_BYTES + 1024;
constexpr int PA_THREADS = 576;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2(s * log2e - m * log2e) for 32 scores -> 16 packed 16-bit pairs; returns the row-sum contribution
template <typename OpT, bool MASKED>
__device__ __forceinline__ float softmax_chunk(const uint32_t (&r)[32], uint32_t (&pk)[16], float mb, int valid) {
  constexpr float L2E = 1.4426950408889634f;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(r[k]), L2E, -mb));
    float p1 = ex2_approx(fmaf(__uint_as_float(r[k + 1]), L2E, -mb));
    if (MASKED) {
      p0 = k < valid ? p0 : 0.f;
      p1 = k + 1 < valid ? p1 : 0.f;
    }
    pk[k >> 1] = OpTraits<OpT>::pack2(p0, p1);
    l0 += p0;
    l1 += p1;
  }
  return l0 + l1;
}
template <bool MASKED>
__device__ __forceinline__ float max_chunk(const uint32_t (&r)[32], int valid) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float a = (!MASKED || k + 0 < valid) ? __uint_as_float(r[k + 0]) : -INFINITY;
    const float b = (!MASKED || k + 1 < valid) ? __uint_as_float(r[k + 1]) : -INFINITY;
    const float c = (!MASKED || k + 2 < valid) ? __uint_as_float(r[k + 2]) : -INFINITY;
    const float d = (!MASKED || k + 3 < valid) ? __uint_as_float(r[k + 3]) : -INFINITY;
    m0 = fmaxf(m0, a), m1 = fmaxf(m1, b), m2 = fmaxf(m2, c), m3 = fmaxf(m3, d);
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

#ifdef MM_ATT_TRACE
__device__ long long g_att_trace[148 * 2 * 8 * 6];
#define ATT_TRACE(slot)                                                                              \
  do {                                                                                               \
    if ((warp & 7) == 0 && lane == 0 && (i >> 1) < 8)                                                \
      g_att_trace[((blockIdx.x * 2 + g) * 8 + (i >> 1)) * 6 + (slot)] = clock64();                   \
  } while (0)
#else
#define ATT_TRACE(slot)
#endif

template <typename OpT>
__global__ void __launch_bounds__(PA_THREADS, 1)
self_attention_t256_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapOut,
                           const int* __restrict__ seq_lens, int T, int d_model, int H, int nqt, int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sOut = smem + 2 * PA_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * PA_OUT_BYTES);
  uint64_t* qk_full = bars;         // [2] TMA (Q, K) -> MMA
  uint64_t* qk_empty = bars + 2;    // [2] S MMA done -> TMA
  uint64_t* v_full = bars + 4;      // [2] TMA (V^T) -> MMA
  uint64_t* v_empty = bars + 6;     // [2] PV MMA done -> TMA
  uint64_t* s_full = bars + 8;      // [2] S MMA done -> softmax group
  uint64_t* p_half = bars + 10;     // [2 groups][2 halves] softmax group (4 warps) -> MMA: 128 keys of P are in TMEM
  uint64_t* o_full = bars + 14;     // [2] PV MMA done -> softmax group
  uint64_t* reg_free = bars + 16;   // [2] softmax group has read O -> MMA may overwrite the TMEM region
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  int* s_lens = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);
  float* s_xch = reinterpret_cast<float*>(s_lens + PA_MAX_LENS);
  const int n_batch = n_items / (nqt * H);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr float L2E = 1.4426950408889634f;

  if (tid == 0) {
    tma_prefetch_desc(&mapQK);
    tma_prefetch_desc(&mapOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_half[2 * i], 4);
      mbar_init(&p_half[2 * i + 1], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&reg_free[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // the QKV projection (previous kernel) is complete and visible
  for (int i = threadIdx.x; i < n_batch && i < PA_MAX_LENS; i += PA_THREADS) s_lens[i] = seq_lens[i];
  __syncthreads();

  if (warp == 16) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        const int qt = item % nqt, h = (item / nqt) % H, b = item / (nqt * H);
        const int s = i & 1, u = i >> 1;
        uint8_t* sQ = smem + s * PA_STAGE_BYTES;
        uint8_t* sK = sQ + AT_Q_BYTES;
        uint8_t* sV = sK + AT_K_BYTES;
        mbar_wait(&qk_empty[s], (u & 1) ^ 1);
        mbar_expect_tx(&qk_full[s], AT_Q_BYTES + AT_K_BYTES);
        tma_load_3d(sQ, &mapQK, &qk_full[s], h * AT_HD, qt * AT_BM, b);
        tma_load_3d(sK, &mapQK, &qk_full[s], d_model + h * AT_HD, 0, b);
        tma_load_3d(sK + AT_K_BYTES / 2, &mapQK, &qk_full[s], d_model + h * AT_HD, 128, b);
        mbar_wait(&v_empty[s], (u & 1) ^ 1);
        mbar_expect_tx(&v_full[s], AT_V_BYTES);   // V rows = keys, 64 values (128 B) each: MN-major B operand
        tma_load_3d(sV, &mapQK, &v_full[s], 2 * d_model + h * AT_HD, 0, b);
        tma_load_3d(sV + AT_V_BYTES / 2, &mapQK, &v_full[s], 2 * d_model + h * AT_HD, 128, b);
      }
    }
  } else if (warp == 17) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(AT_BM, AT_KC, OpTraits<OpT>::fmt);
      constexpr uint32_t idesc_o = umma_idesc(AT_BM, AT_HD, OpTraits<OpT>::fmt) | (1u << 16);  // B (= V) is MN-major
      // Two work queues (next S = Q K^T, next O = P V) served by whichever is ready first: the issuer never blocks
      // on one softmax group while the other group's MMA could be issued, so the groups settle into ping-pong.
      int s_next = 0, pva_next = 0, pv_next = 0;   // next S, next first-half P V, next second-half P V
      const uint64_t t0 = globaltimer_ns();
      while (pv_next < n_local) {
        bool progressed = false;
        if (s_next < n_local && s_next <= pv_next + 1) {
          const int s = s_next & 1, u = s_next >> 1;
          if (mbar_test(&qk_full[s], u & 1) && mbar_test(&reg_free[s], (u & 1) ^ 1)) {
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem + s * PA_STAGE_BYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + s * PA_STAGE_BYTES + AT_Q_BYTES));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_base + 256 * s, adesc + 2 * kk, bdesc + 2 * kk, idesc_s, kk != 0);
            umma_commit(&s_full[s]);
            umma_commit(&qk_empty[s]);
            ++s_next;
            progressed = true;
          }
        }
        if (pva_next < s_next) {
          const int s = pva_next & 1, u = pva_next >> 1;
          if (mbar_test(&p_half[2 * s], u & 1) && mbar_test(&v_full[s], u & 1)) {
            tc_fence_after();
            const uint64_t vdesc = umma_desc_sw128(smem_u32(smem + s * PA_STAGE_BYTES + AT_Q_BYTES + AT_K_BYTES));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)   // 16 keys per step: 8 packed P columns, 16 V rows of 128 B
                umma_f16_ts(tmem_base + 256 * s + 64, tmem_base + 256 * s + 8 * (4 * j + kk),
                            vdesc + (uint64_t)(((j * 64 + kk * 16) * 128) >> 4), idesc_o, (j | kk) != 0);
            }
            ++pva_next;
            progressed = true;
          }
        }
        if (pv_next < pva_next) {
          const int s = pv_next & 1, u = pv_next >> 1;
          if (mbar_test(&p_half[2 * s + 1], u & 1)) {
            tc_fence_after();
            const uint64_t vdesc = umma_desc_sw128(smem_u32(smem + s * PA_STAGE_BYTES + AT_Q_BYTES + AT_K_BYTES));
#pragma unroll
            for (int j = 2; j < 4; ++j) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ts(tmem_base + 256 * s + 64, tmem_base + 256 * s + 128 + 8 * (4 * (j - 2) + kk),
                            vdesc + (uint64_t)(((j * 64 + kk * 16) * 128) >> 4), idesc_o, 1);
            }
            umma_commit(&o_full[s]);
            umma_commit(&v_empty[s]);
            ++pv_next;
            progressed = true;
          }
        }
        if (!progressed && globaltimer_ns() - t0 > 8000000000ull) {
          printf("mm: attention issuer timeout block %d s_next %d pva_next %d pv_next %d\n", blockIdx.x, s_next,
                 pva_next, pv_next);
          __trap();
        }
      }
    }
  } else {
    // ---------------- softmax + epilogue groups ----------------
    const int g = warp >> 3;                   // item parity served by this group
    const int hf = (warp >> 2) & 1;            // key half: chunks 4 hf .. 4 hf + 3 of the 8 32-key chunks
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem_base + 256 * g + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float* x_max = s_xch + g * 4 * AT_BM;      // [2 halves][128 rows]
    float* x_sum = x_max + 2 * AT_BM;
    auto group_sync = [&]() {
      if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    for (int i = g; i < n_local; i += 2) {
      const int item = blockIdx.x + i * gridDim.x;
      const int qt = item % nqt, h = (item / nqt) % H, b = item / (nqt * H);
      const int u = i >> 1;
      const int len = max(1, min(b < PA_MAX_LENS ? s_lens[b] : seq_lens[b], T));
      const int nch = (len + 31) >> 5;        // chunks holding at least one valid key
      const int nfull = len >> 5;             // chunks that need no masking
      const int rem = len & 31;               // valid keys in chunk `nfull` when rem > 0
      const int c0 = 4 * hf;                  // my chunks: c0 .. c0 + 3
      ATT_TRACE(0);
      mbar_wait(&s_full[g], u & 1);
      tc_fence_after();
      ATT_TRACE(1);
      uint32_t ra[32], rb[32];
      // ---- sweep 1: maximum over my 128 keys, then over the row ----
      float m = -INFINITY;
#pragma unroll 1
      for (int cc = c0; cc < c0 + 4 && cc < nch; cc += 2) {
        tmem_ld32(t_row + cc * 32, ra);
        if (cc + 1 < nch) tmem_ld32(t_row + (cc + 1) * 32, rb);
        tmem_ld_wait();
        m = fmaxf(m, cc < nfull ? max_chunk<false>(ra, 32) : max_chunk<true>(ra, rem));
        if (cc + 1 < nch) m = fmaxf(m, cc + 1 < nfull ? max_chunk<false>(rb, 32) : max_chunk<true>(rb, rem));
      }
      if (c0 < nch) {   // prefetch my first chunk for sweep 2 across the exchange
        tmem_ld32(t_row + c0 * 32, ra);
      }
      x_max[hf * AT_BM + row] = m;
      group_sync();
      m = fmaxf(m, x_max[(hf ^ 1) * AT_BM + row]);
      const float mb = m * L2E;
      ATT_TRACE(2);
      // ---- sweep 2: probabilities of my 128 keys (next chunk prefetched while the current one goes through MUFU) ----
      float l = 0.f;
      tmem_ld_wait();
      const uint32_t p_col = t_row + 128 * hf;   // P(keys 0-127) -> columns [0,64), P(keys 128-255) -> [128,192)
#pragma unroll 1
      for (int cc = c0; cc < c0 + 4; cc += 2) {
        uint32_t pk[16];
        if (cc + 1 < nch) tmem_ld32(t_row + (cc + 1) * 32, rb);
        if (cc < nch) {
          l += cc < nfull ? softmax_chunk<OpT, false>(ra, pk, mb, 32) : softmax_chunk<OpT, true>(ra, pk, mb, rem);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = 0u;
        }
        tmem_ld_wait();
        tmem_st16(p_col + 16 * (cc - c0), pk);    // P over S columns that have already been consumed
        if (cc + 2 < c0 + 4 && cc + 2 < nch) tmem_ld32(t_row + (cc + 2) * 32, ra);
        if (cc + 1 < nch) {
          l += cc + 1 < nfull ? softmax_chunk<OpT, false>(rb, pk, mb, 32) : softmax_chunk<OpT, true>(rb, pk, mb, rem);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = 0u;
        }
        tmem_ld_wait();
        tmem_st16(p_col + 16 * (cc + 1 - c0), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_half[2 * g + hf]);   // my 128 keys of P are in TMEM: their P V can start
      x_sum[hf * AT_BM + row] = l;
      ATT_TRACE(3);

      mbar_wait(&o_full[g], u & 1);
      tc_fence_after();
      ATT_TRACE(4);
      {
        tmem_ld32(t_row + 64 + 32 * hf, ra);        // my 32 of the 64 output columns
        uint8_t* so = sOut + g * PA_OUT_BYTES;
        const bool elected = (warp & 7) == 0 && lane == 0;
        if (elected) bulk_wait_read<0>();           // the previous item's store has finished reading the tile
        group_sync();                               // ... and both halves of the row sum are visible
        const float inv = 1.0f / (l + x_sum[(hf ^ 1) * AT_BM + row]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&reg_free[g]);   // O is in registers: the TMEM region may be reused
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 q;
          q.x = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 0]) * inv, __uint_as_float(ra[8 * k + 1]) * inv);
          q.y = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 2]) * inv, __uint_as_float(ra[8 * k + 3]) * inv);
          q.z = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 4]) * inv, __uint_as_float(ra[8 * k + 5]) * inv);
          q.w = OpTraits<OpT>::pack2(__uint_as_float(ra[8 * k + 6]) * inv, __uint_as_float(ra[8 * k + 7]) * inv);
          *reinterpret_cast<uint4*>(so + row * 128 + (((4 * hf + k) ^ (row & 7)) << 4)) = q;
        }
        fence_proxy_async_smem();
        group_sync();
        if (elected) {
          tma_store_3d(&mapOut, so, h * AT_HD, qt * AT_BM, b);
          bulk_commit();
        }
      }
      ATT_TRACE(5);
    }
    if ((warp & 7) == 0 && lane == 0) bulk_wait<0>();   // the last stores have landed before the CTA's smem goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename OpT>
static int launch_attn_t256(const CUtensorMap& mqk, const CUtensorMap& mout, const int* lens, int B, int T, int H,
                            int d, cudaStream_t s) {
  auto kern = self_attention_t256_kernel<OpT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PA_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(self_attention_t256)");
    attr_set = true;
  }
  const int nqt = (T + AT_BM - 1) / AT_BM;
  const int n_items = B * H * nqt;
  const int grid = n_items < kNumSMs ? n_items : kNumSMs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(PA_THREADS);
  cfg.dynamicSmemBytes = PA_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mqk, mout, lens, T, d, H, nqt, n_items);
  if (e != cudaSuccess) return fail(e, "self_attention_t256_kernel launch");
  return 0;
}

}  // namespace mm

using namespace mm;

#ifdef MM_ATT_TRACE
extern "C" int mm_debug_att_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, g_att_trace, sizeof(long long) * 148 * 2 * 8 * 6);
}
#endif

extern "C" int mm_self_attention(const void* qkv, int64_t qkv_ld, const int32_t* seq_lens, int32_t batch, int32_t seq,
                                 int32_t heads, void* out, int64_t out_ld, int32_t dtype, void* stream) {
  if (!qkv || !seq_lens || !out) return bad_arg("self_attention: null pointer");
  if (batch <= 0 || seq <= 0 || heads <= 0) return bad_arg("self_attention: extents");
  const int d = heads * AT_HD;
  if (qkv_ld < 3 * d || (qkv_ld % 8) || (out_ld % 8) || out_ld < d)
    return bad_arg("self_attention: leading dims (head_dim must be 64)");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap mqk;
  int rc = make_tmap_3d(&mqk, qkv, f16, (uint64_t)(3 * d), (uint64_t)seq, (uint64_t)batch, (uint64_t)qkv_ld,
                        (uint64_t)seq * qkv_ld, 128);
  if (rc) return rc;
  const CUtensorMap& mvt = mqk;   // V is read from the same [B*T, 3d] tensor (columns [2d, 3d)) as an MN-major operand
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (seq <= AT_KC) {   // single key chunk: persistent warp-specialised kernel, output through TMA ([B][T][d] view)
    CUtensorMap mout;
    rc = make_tmap_3d(&mout, out, f16, (uint64_t)d, (uint64_t)seq, (uint64_t)batch, (uint64_t)out_ld,
                      (uint64_t)seq * out_ld, 128);
    if (rc) return rc;
    return f16 ? launch_attn_t256<__half>(mqk, mout, seq_lens, batch, seq, heads, d, s)
               : launch_attn_t256<__nv_bfloat16>(mqk, mout, seq_lens, batch, seq, heads, d, s);
  }
  return f16 ? launch_attn<__half>(mqk, mvt, seq_lens, batch, seq, heads, d, out, out_ld, s)
             : launch_attn<__nv_bfloat16>(mqk, mvt, seq_lens, batch, seq, heads, d, out, out_ld, s);
}
