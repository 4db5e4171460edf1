// Multi-head self-attention core on tcgen05: O = softmax_fp32(Q K^T + key-padding mask) V, head_dim 64.
// q, k, v are column blocks of ONE [B*T, 3d] tensor (the QKV GEMM output); q arrives pre-scaled by head_dim^-0.5
// (fused into that GEMM's epilogue).  Reference: fairseq/modules/multihead_attention.py (encoder self-attention,
// key_padding_mask), as called from TransformerEncoderLayerBase.forward.
//
// Two persistent, warp-specialised kernels (one CTA per SM walks a list of (utterance, head, 128-query tile) items;
// TMA producer warp, MMA warp, two ping-pong softmax groups of 8 warps; S, P and O live in TMEM; O leaves by TMA store):
//  * self_attention_t256_kernel  T <= 256 (utterances up to ~10 s at 4x subsampling): one 256-key chunk, plain softmax.
//    This is the one the bench runs.
//  * self_attention_long_kernel  any T: 128-key chunks, online softmax (running maximum, O rescaled in TMEM), chunks
//    beyond an utterance's length skipped.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int AT_BM = 128, AT_HD = 64, AT_KC = 256;
constexpr int AT_Q_BYTES = AT_BM * AT_HD * 2;        // 16 KB
constexpr int AT_K_BYTES = AT_KC * AT_HD * 2;        // 32 KB
constexpr int AT_V_BYTES = AT_HD * AT_KC * 2;        // 32 KB (4 blocks of 64 keys)

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
// The same with attention dropout (training forward; fairseq MultiheadAttention dropout_module on the probabilities):
// the row sum (the softmax denominator) is taken over the UN-dropped probabilities, the packed operand of the P V product
// carries keep / (1 - p).  i4 = (element index of the chunk's first key in the [batch*heads][Tp][Tp] mask) / 4.
struct AttnDrop {
  float p;
  unsigned site;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  int tp;       // row length of the mask index space: round_up(T, 64)
};
template <typename OpT>
__device__ __forceinline__ float softmax_chunk_drop(const uint32_t (&r)[32], uint32_t (&pk)[16], float mb, int valid,
                                                    unsigned long long seed, unsigned site, unsigned long long i4,
                                                    unsigned thr, float inv) {
  constexpr float L2E = 1.4426950408889634f;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int k = 4 * g;
    float p0 = ex2_approx(fmaf(__uint_as_float(r[k]), L2E, -mb));
    float p1 = ex2_approx(fmaf(__uint_as_float(r[k + 1]), L2E, -mb));
    float p2 = ex2_approx(fmaf(__uint_as_float(r[k + 2]), L2E, -mb));
    float p3 = ex2_approx(fmaf(__uint_as_float(r[k + 3]), L2E, -mb));
    p0 = k < valid ? p0 : 0.f;
    p1 = k + 1 < valid ? p1 : 0.f;
    p2 = k + 2 < valid ? p2 : 0.f;
    p3 = k + 3 < valid ? p3 : 0.f;
    l0 += p0 + p2;
    l1 += p1 + p3;
    dropout_apply4(dropout_bits4(seed, site, i4 + g), thr, inv, p0, p1, p2, p3);
    pk[2 * g] = OpTraits<OpT>::pack2(p0, p1);
    pk[2 * g + 1] = OpTraits<OpT>::pack2(p2, p3);
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

template <typename OpT, bool DROP>
__global__ void __launch_bounds__(PA_THREADS, 1)
self_attention_t256_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapOut,
                           const int* __restrict__ seq_lens, int T, int d_model, int H, int nqt, int n_items,
                           float* __restrict__ lse, const AttnDrop dr) {
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
    const unsigned long long drop_seed = DROP ? dr.seed + (dr.seed_dev ? *dr.seed_dev : 0ull) : 0ull;
    const unsigned drop_thr = dropout_threshold(dr.p);
    const float drop_inv = 1.0f / (1.0f - dr.p);
    for (int i = g; i < n_local; i += 2) {
      const int item = blockIdx.x + i * gridDim.x;
      const int qt = item % nqt, h = (item / nqt) % H, b = item / (nqt * H);
      const int u = i >> 1;
      // mask index of (this thread's query row, key 0): ((b H + h) Tp + q) Tp, in units of four elements
      const unsigned long long i4row =
          DROP ? ((unsigned long long)((long long)(b * H + h) * dr.tp + qt * AT_BM + row) * (unsigned)dr.tp) >> 2 : 0ull;
      auto chunk = [&](const uint32_t (&r)[32], uint32_t (&pk)[16], float mb_, int cc, int nfull_, int rem_) -> float {
        if constexpr (DROP)
          return softmax_chunk_drop<OpT>(r, pk, mb_, cc < nfull_ ? 32 : rem_, drop_seed, dr.site, i4row + 8 * cc, drop_thr,
                                         drop_inv);
        else
          return cc < nfull_ ? softmax_chunk<OpT, false>(r, pk, mb_, 32) : softmax_chunk<OpT, true>(r, pk, mb_, rem_);
      };
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
#ifndef MM_ATT_NO_TURNS
      // The exponentials of the two groups take turns in item order (0, 1, 2, ...: named barriers 3 / 4, 256 waiting +
      // 256 arriving threads).  Left alone the groups drift into phase: both sit in sweep 2 at once, each at half the
      // MUFU rate, and then both leave the MUFU pipe idle while they wait for O and store.  With turns a group has the
      // pipe to itself and the other group's loads / maximum / P V wait / store run beside it.
      if (i > 0) {
        if (g == 0) asm volatile("bar.sync 3, 512;" ::: "memory"); else asm volatile("bar.sync 4, 512;" ::: "memory");
      }
#endif
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
          l += chunk(ra, pk, mb, cc, nfull, rem);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = 0u;
        }
        tmem_ld_wait();
        tmem_st16(p_col + 16 * (cc - c0), pk);    // P over S columns that have already been consumed
        if (cc + 2 < c0 + 4 && cc + 2 < nch) tmem_ld32(t_row + (cc + 2) * 32, ra);
        if (cc + 1 < nch) {
          l += chunk(rb, pk, mb, cc + 1, nfull, rem);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = 0u;
        }
        tmem_ld_wait();
        tmem_st16(p_col + 16 * (cc + 1 - c0), pk);
      }
#ifndef MM_ATT_NO_TURNS
      if (i + 1 < n_local) {   // item i + 1 (the other group's) may exponentiate
        if (g == 0) asm volatile("bar.arrive 4, 512;" ::: "memory"); else asm volatile("bar.arrive 3, 512;" ::: "memory");
      }
#endif
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
        const float lt = l + x_sum[(hf ^ 1) * AT_BM + row];
        const float inv = 1.0f / lt;
        // log-sum-exp of the row's scores (natural log), [batch][head][query]: what a backward pass needs to rebuild P
        if (lse != nullptr && hf == 0 && qt * AT_BM + row < T)
          lse[((long long)b * H + h) * T + qt * AT_BM + row] = m + __logf(lt);
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

template <typename OpT, bool DROP = false>
static int launch_attn_t256(const CUtensorMap& mqk, const CUtensorMap& mout, const int* lens, int B, int T, int H,
                            int d, float* lse, cudaStream_t s, AttnDrop dr = AttnDrop{0.f, 0u, 0ull, nullptr, 0}) {
  auto kern = self_attention_t256_kernel<OpT, DROP>;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mqk, mout, lens, T, d, H, nqt, n_items, lse, dr);
  if (e != cudaSuccess) return fail(e, "self_attention_t256_kernel launch");
  return 0;
}


// ===================================================================================================
// Persistent variant for T > 256: the same roles and the same two ping-pong softmax groups, but keys arrive in chunks
// of 128 and the softmax is ONLINE (running row maximum m, running sum l, O rescaled by exp2(m_old - m_new) when the
// maximum moves), so S / P / O of a group fit in 192 TMEM columns for any T:
//   region g = columns [256 g, 256 g + 256):  S chunk (fp32, 128 keys) in [0,128) -> P (packed 16-bit) in [0,64),
//   O (fp32, 64 values) in [128,192), accumulated over the chunks by the P V UMMAs.
// Per chunk and group: S = Q K_c^T (SS UMMA) -> each of the 256 threads takes 64 scores of its row (two threads per
// row, row maximum combined through shared memory) -> P to TMEM, O rescaled in TMEM only by warps that saw their
// maximum move -> O += P V_c (TS UMMA, V_c rows as MN-major B).  The next chunk's S UMMA is issued right behind the
// P V UMMAs (the tensor pipe executes in order, so it cannot overtake their reads of P).  Chunks that lie entirely
// beyond an utterance's length are skipped, so ragged batches cost what their real lengths cost.
// Shared memory: per group a Q tile (16 KB) and a 2-stage ring of (K_c | V_c) (32 KB per stage), plus the staging
// tile for the TMA store of O.
// ===================================================================================================
constexpr int PL_KC = 128;                                   // keys per chunk
constexpr int PL_KV_BYTES = 2 * PL_KC * AT_HD * 2;           // K_c + V_c: 32 KB
constexpr int PL_SMEM_BYTES = 2 * AT_Q_BYTES + 2 * 2 * PL_KV_BYTES + 2 * PA_OUT_BYTES + 256 + PA_MAX_LENS * 4 +
                              PA_XCH_BYTES + 1024;
static_assert(PL_SMEM_BYTES <= 232448, "shared memory budget");

template <typename OpT, bool DROP>
__global__ void __launch_bounds__(PA_THREADS, 1)
self_attention_long_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                           const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapOut,
                           const int* __restrict__ seq_lens, int T, int q_col0, int k_col0, int v_col0, int causal,
                           int H, int nqt, int n_items, int q_len, float* __restrict__ lse, const AttnDrop dr) {
  // Q, K, V may be three different tensors (decoder cross-attention: queries from the decoder states, keys / values
  // from the projected encoder states) or column blocks of one (self-attention: q | k | v of the QKV GEMM).  T is
  // the key extent, seq_lens[b] (NULL: T) the number of valid keys of utterance b; the number of query rows only
  // enters through nqt and the bounds of mapQ / mapOut.  causal: key k is visible to query row q iff k <= q.
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // [2 groups] 16 KB
  uint8_t* sKV = sQ + 2 * AT_Q_BYTES;                  // [2 groups][2 stages] (K_c 16 KB | V_c 16 KB)
  uint8_t* sOut = sKV + 4 * PL_KV_BYTES;               // [2 groups] 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * PA_OUT_BYTES);
  uint64_t* q_full = bars;            // [2]     TMA Q -> MMA
  uint64_t* q_empty = bars + 2;       // [2]     last S UMMA of the item done -> TMA
  uint64_t* kv_full = bars + 4;       // [2][2]  TMA (K_c, V_c) -> MMA
  uint64_t* kv_empty = bars + 8;      // [2][2]  P V_c done -> TMA
  uint64_t* s_full = bars + 12;       // [2]     S_c done -> softmax group
  uint64_t* p_full = bars + 14;       // [2]     softmax group (8 warps): P_c stored, O rescaled -> MMA
  uint64_t* o_done = bars + 16;       // [2]     P V_c done -> softmax group
  uint64_t* reg_free = bars + 18;     // [2]     group has read the final O -> MMA may start the next item there
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  int* s_lens = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);
  float* s_xch = reinterpret_cast<float*>(s_lens + PA_MAX_LENS);
  const int n_batch = n_items / (nqt * H);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr float L2E = 1.4426950408889634f;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&o_done[i], 1);
      mbar_init(&reg_free[i], 8);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < n_batch && i < PA_MAX_LENS; i += PA_THREADS) s_lens[i] = seq_lens ? seq_lens[i] : T;
  __syncthreads();
  auto len_of = [&](int b) { return max(1, min(b < PA_MAX_LENS ? s_lens[b] : (seq_lens ? seq_lens[b] : T), T)); };
  auto item_coords = [&](int i, int& qt, int& h, int& b) {
    const int item = blockIdx.x + i * gridDim.x;
    qt = item % nqt, h = (item / nqt) % H, b = item / (nqt * H);
  };
  // key chunks of an item: up to the utterance length, and under a causal mask up to the query tile's diagonal chunk
  auto chunks_of = [&](int b, int qt) {
    const int nc = (len_of(b) + PL_KC - 1) / PL_KC;
    return causal ? min(nc, qt + 1) : nc;
  };

  if (warp == 16) {
    // ---------------- TMA producer: the two groups' (item, chunk) streams, interleaved chunk by chunk ----------------
    if (lane == 0) {
      int it[2] = {0, 1}, ch[2] = {0, 0};          // next item / chunk per group
      uint32_t kv_n[2] = {0, 0}, q_n[2] = {0, 0};   // loads issued so far (-> stage and parity)
      while (it[0] < n_local || it[1] < n_local) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (it[g] >= n_local) continue;
          int qt, h, b;
          item_coords(it[g], qt, h, b);
          const int nc = chunks_of(b, qt);
          if (ch[g] == 0) {
            mbar_wait(&q_empty[g], (q_n[g] & 1) ^ 1);
            mbar_expect_tx(&q_full[g], AT_Q_BYTES);
            tma_load_3d(sQ + g * AT_Q_BYTES, &mapQ, &q_full[g], q_col0 + h * AT_HD, qt * AT_BM, b);
            ++q_n[g];
          }
          const uint32_t st = kv_n[g] & 1;
          uint8_t* kv = sKV + (2 * g + st) * PL_KV_BYTES;
          mbar_wait(&kv_empty[2 * g + st], ((kv_n[g] >> 1) & 1) ^ 1);
          mbar_expect_tx(&kv_full[2 * g + st], PL_KV_BYTES);
          tma_load_3d(kv, &mapK, &kv_full[2 * g + st], k_col0 + h * AT_HD, ch[g] * PL_KC, b);
          tma_load_3d(kv + PL_KV_BYTES / 2, &mapV, &kv_full[2 * g + st], v_col0 + h * AT_HD, ch[g] * PL_KC, b);
          ++kv_n[g];
          if (++ch[g] == nc) ch[g] = 0, it[g] += 2;
        }
      }
    }
  } else if (warp == 17) {
    // ---------------- MMA issuer: per group a two-state machine (S of chunk c, then P V of chunk c) ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(AT_BM, PL_KC, OpTraits<OpT>::fmt);
      constexpr uint32_t idesc_o = umma_idesc(AT_BM, AT_HD, OpTraits<OpT>::fmt) | (1u << 16);  // B (= V) is MN-major
      int it[2] = {0, 1}, ch[2] = {0, 0}, nc[2] = {0, 0};
      bool need_pv[2] = {false, false};
      uint32_t kv_n[2] = {0, 0}, q_n[2] = {0, 0}, pv_n[2] = {0, 0};
      const uint64_t t0 = globaltimer_ns();
      while (it[0] < n_local || it[1] < n_local) {
        bool progressed = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (it[g] >= n_local) continue;
          const uint32_t st = kv_n[g] & 1;
          uint8_t* kv = sKV + (2 * g + st) * PL_KV_BYTES;
          if (!need_pv[g]) {
            if (ch[g] == 0) {
              if (!mbar_test(&q_full[g], q_n[g] & 1) || !mbar_test(&reg_free[g], (q_n[g] & 1) ^ 1)) continue;
              int qt, h, b;
              item_coords(it[g], qt, h, b);
              nc[g] = chunks_of(b, qt);
            }
            if (!mbar_test(&kv_full[2 * g + st], (kv_n[g] >> 1) & 1)) continue;
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(sQ + g * AT_Q_BYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(kv));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_base + 256 * g, adesc + 2 * kk, bdesc + 2 * kk, idesc_s, kk != 0);
            umma_commit(&s_full[g]);
            if (ch[g] + 1 == nc[g]) {   // last S of the item: the Q tile may be replaced
              umma_commit(&q_empty[g]);
              ++q_n[g];
            }
            need_pv[g] = true;
            progressed = true;
          } else {
            if (!mbar_test(&p_full[g], pv_n[g] & 1)) continue;
            tc_fence_after();
            const uint64_t vdesc = umma_desc_sw128(smem_u32(kv + PL_KV_BYTES / 2));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)   // 16 keys per step: 8 packed P columns, 16 V rows of 128 B
                umma_f16_ts(tmem_base + 256 * g + 128, tmem_base + 256 * g + 8 * (4 * j + kk),
                            vdesc + (uint64_t)(((j * 64 + kk * 16) * 128) >> 4), idesc_o, (ch[g] | j | kk) != 0);
            }
            umma_commit(&o_done[g]);
            umma_commit(&kv_empty[2 * g + st]);
            ++pv_n[g];
            ++kv_n[g];
            need_pv[g] = false;
            if (++ch[g] == nc[g]) ch[g] = 0, it[g] += 2;
            progressed = true;
          }
        }
        if (!progressed && globaltimer_ns() - t0 > 8000000000ull) {
          printf("mm: long attention issuer timeout block %d items %d %d chunks %d %d\n", blockIdx.x, it[0], it[1], ch[0],
                 ch[1]);
          __trap();
        }
      }
    }
  } else {
    // ---------------- softmax + epilogue groups ----------------
    const int g = warp >> 3;
    const int hf = (warp >> 2) & 1;            // which 64 keys of a chunk / which 32 output columns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem_base + 256 * g + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float* x_max = s_xch + g * 4 * AT_BM;      // [2 halves][128 rows]
    float* x_sum = x_max + 2 * AT_BM;
    auto group_sync = [&]() {
      if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    uint32_t s_n = 0, o_n = 0;                 // S chunks / P V chunks seen by this group (-> parities)
    const unsigned long long drop_seed = DROP ? dr.seed + (dr.seed_dev ? *dr.seed_dev : 0ull) : 0ull;
    const unsigned drop_thr = dropout_threshold(dr.p);
    const float drop_inv = 1.0f / (1.0f - dr.p);
    const int drop_lp = (q_len + 63) / 64 * 64;
    for (int i = g; i < n_local; i += 2) {
      int qt, h, b;
      item_coords(i, qt, h, b);
      const int len = len_of(b);
      const int nc = chunks_of(b, qt);
      float m = -INFINITY, l = 0.f;
      for (int c = 0; c < nc; ++c) {
        int valid = min(64, max(0, len - c * PL_KC - 64 * hf));   // valid keys among my 64
        if (causal && c == qt) valid = min(valid, max(0, row + 1 - 64 * hf));   // diagonal chunk: keys <= my query row
        mbar_wait(&s_full[g], s_n & 1);
        ++s_n;
        tc_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld32(t_row + 64 * hf, ra);
        tmem_ld32(t_row + 64 * hf + 32, rb);
        tmem_ld_wait();
        float mloc = valid >= 32 ? max_chunk<false>(ra, 32) : max_chunk<true>(ra, valid);
        if (valid > 32) mloc = fmaxf(mloc, valid >= 64 ? max_chunk<false>(rb, 32) : max_chunk<true>(rb, valid - 32));
        x_max[hf * AT_BM + row] = mloc;
        group_sync();                            // both halves hold their S in registers: P may overwrite it
        const float m_new = fmaxf(m, fmaxf(mloc, x_max[(hf ^ 1) * AT_BM + row]));
        const float alpha = ex2_approx((m - m_new) * L2E);   // 0 for the first chunk (m = -inf)
        m = m_new;
        const float mb = m * L2E;
        uint32_t pk[16];
        float lc;
        if constexpr (DROP) {
          // mask index of (query row, first of my 64 keys): ((b H + h) Lp + q) Tp + key, four elements per hash
          const unsigned long long i4 =
              ((unsigned long long)((long long)(b * H + h) * drop_lp + qt * AT_BM + row) * (unsigned)dr.tp +
               (unsigned)(c * PL_KC + 64 * hf)) >> 2;
          lc = softmax_chunk_drop<OpT>(ra, pk, mb, min(valid, 32), drop_seed, dr.site, i4, drop_thr, drop_inv);
          tmem_st16(t_row + 32 * hf, pk);
          lc += softmax_chunk_drop<OpT>(rb, pk, mb, max(valid - 32, 0), drop_seed, dr.site, i4 + 8, drop_thr, drop_inv);
          tmem_st16(t_row + 32 * hf + 16, pk);
        } else {
        if (valid >= 32) lc = softmax_chunk<OpT, false>(ra, pk, mb, 32);
        else lc = softmax_chunk<OpT, true>(ra, pk, mb, valid);
        tmem_st16(t_row + 32 * hf, pk);
        if (valid >= 64) lc += softmax_chunk<OpT, false>(rb, pk, mb, 32);
        else lc += softmax_chunk<OpT, true>(rb, pk, mb, max(valid - 32, 0));
        tmem_st16(t_row + 32 * hf + 16, pk);
        }
        l = l * alpha + lc;
        // O <- alpha O, only in warps where some row's maximum moved; P V of the previous chunk must have landed
        if (c > 0) {
          mbar_wait(&o_done[g], o_n & 1);
          ++o_n;
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
            tmem_ld32(t_row + 128 + 32 * hf, ra);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) ra[k] = __float_as_uint(__uint_as_float(ra[k]) * alpha);
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) lo[k] = ra[k], hi[k] = ra[16 + k];
            tmem_st16(t_row + 128 + 32 * hf, lo);
            tmem_st16(t_row + 128 + 32 * hf + 16, hi);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        // the next chunk's x_max write happens after the next S wait; the read above finished before p_full fired
        // for this warp, but the OTHER half may still be reading: order through the group barrier of the next chunk
      }
      x_sum[hf * AT_BM + row] = l;
      mbar_wait(&o_done[g], o_n & 1);
      ++o_n;
      tc_fence_after();
      {
        uint32_t ra[32];
        tmem_ld32(t_row + 128 + 32 * hf, ra);
        uint8_t* so = sOut + g * PA_OUT_BYTES;
        const bool elected = (warp & 7) == 0 && lane == 0;
        if (elected) bulk_wait_read<0>();
        group_sync();
        const float lt = l + x_sum[(hf ^ 1) * AT_BM + row];
        const float inv = 1.0f / lt;
        if (lse != nullptr && hf == 0 && qt * AT_BM + row < q_len)      // [batch][head][query], natural log
          lse[((long long)b * H + h) * q_len + qt * AT_BM + row] = m + __logf(lt);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&reg_free[g]);
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
    }
    if ((warp & 7) == 0 && lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename OpT, bool DROP = false>
static int launch_attn_long(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mout,
                            const int* lens, int B, int Tq, int Tk, int q_col0, int k_col0, int v_col0, int causal,
                            int H, float* lse, cudaStream_t s, AttnDrop dr = AttnDrop{0.f, 0u, 0ull, nullptr, 0}) {
  auto kern = self_attention_long_kernel<OpT, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(self_attention_long)");
    attr_set = true;
  }
  const int nqt = (Tq + AT_BM - 1) / AT_BM;
  const int n_items = B * H * nqt;
  const int grid = n_items < kNumSMs ? n_items : kNumSMs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(PA_THREADS);
  cfg.dynamicSmemBytes = PL_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, mout, lens, Tk, q_col0, k_col0, v_col0, causal, H, nqt,
                                     n_items, Tq, lse, dr);
  if (e != cudaSuccess) return fail(e, "self_attention_long_kernel launch");
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
  return mm_self_attention_lse(qkv, qkv_ld, seq_lens, batch, seq, heads, out, out_ld, nullptr, dtype, stream);
}

extern "C" int mm_self_attention_lse(const void* qkv, int64_t qkv_ld, const int32_t* seq_lens, int32_t batch, int32_t seq,
                                     int32_t heads, void* out, int64_t out_ld, float* lse, int32_t dtype, void* stream) {
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
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // 129..256 keys: one 256-key chunk is cheapest (23.5 vs 27 us at the bench shape); up to 128 keys the chunked kernel
  // needs a single 128-key chunk (16.5 vs 20 us at T = 125), beyond 256 it is the only one
  if (seq > PL_KC && seq <= AT_KC) {   // single 256-key chunk: persistent warp-specialised kernel, output through TMA ([B][T][d] view)
    CUtensorMap mout;
    rc = make_tmap_3d(&mout, out, f16, (uint64_t)d, (uint64_t)seq, (uint64_t)batch, (uint64_t)out_ld,
                      (uint64_t)seq * out_ld, 128);
    if (rc) return rc;
    return f16 ? launch_attn_t256<__half>(mqk, mout, seq_lens, batch, seq, heads, d, lse, s)
               : launch_attn_t256<__nv_bfloat16>(mqk, mout, seq_lens, batch, seq, heads, d, lse, s);
  }
  // T > 256: persistent online-softmax kernel, 128-key chunks; mqk (box 64 x 128) serves Q, K and V chunks alike
  CUtensorMap mout;
  rc = make_tmap_3d(&mout, out, f16, (uint64_t)d, (uint64_t)seq, (uint64_t)batch, (uint64_t)out_ld,
                    (uint64_t)seq * out_ld, 128);
  if (rc) return rc;
  return f16 ? launch_attn_long<__half>(mqk, mqk, mqk, mout, seq_lens, batch, seq, seq, 0, d, 2 * d, 0, heads, lse, s)
             : launch_attn_long<__nv_bfloat16>(mqk, mqk, mqk, mout, seq_lens, batch, seq, seq, 0, d, 2 * d, 0, heads, lse, s);
}

extern "C" int mm_self_attention_drop(const void* qkv, int64_t qkv_ld, const int32_t* seq_lens, int32_t batch, int32_t seq,
                                      int32_t heads, void* out, int64_t out_ld, float* lse, float drop_p, uint64_t seed,
                                      const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream) {
  if (!qkv || !seq_lens || !out) return bad_arg("self_attention_drop: null pointer");
  if (batch <= 0 || heads <= 0) return bad_arg("self_attention_drop: extents");
  if (!(drop_p > 0.f && drop_p < 1.f)) return bad_arg("self_attention_drop: p in (0, 1)");
  const int d = heads * AT_HD;
  if (qkv_ld < 3 * d || (qkv_ld % 8) || (out_ld % 8) || out_ld < d)
    return bad_arg("self_attention_drop: leading dims (head_dim must be 64)");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap mqk, mout;
  int rc = make_tmap_3d(&mqk, qkv, f16, (uint64_t)(3 * d), (uint64_t)seq, (uint64_t)batch, (uint64_t)qkv_ld,
                        (uint64_t)seq * qkv_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mout, out, f16, (uint64_t)d, (uint64_t)seq, (uint64_t)batch, (uint64_t)out_ld, (uint64_t)seq * out_ld, 128);
  if (rc) return rc;
  AttnDrop dr{drop_p, site, seed, reinterpret_cast<const unsigned long long*>(seed_dev), (seq + 63) / 64 * 64};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!(seq > PL_KC && seq <= AT_KC))     // the chunked kernel (same choice as mm_self_attention_lse)
    return f16 ? launch_attn_long<__half, true>(mqk, mqk, mqk, mout, seq_lens, batch, seq, seq, 0, d, 2 * d, 0, heads, lse, s, dr)
               : launch_attn_long<__nv_bfloat16, true>(mqk, mqk, mqk, mout, seq_lens, batch, seq, seq, 0, d, 2 * d, 0, heads,
                                                         lse, s, dr);
  return f16 ? launch_attn_t256<__half, true>(mqk, mout, seq_lens, batch, seq, heads, d, lse, s, dr)
             : launch_attn_t256<__nv_bfloat16, true>(mqk, mout, seq_lens, batch, seq, heads, d, lse, s, dr);
}

extern "C" int mm_attention(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                            int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                            const int32_t* kv_lens, int32_t batch, int32_t heads, int32_t causal, void* out,
                            int64_t out_ld, int32_t dtype, void* stream) {
  return mm_attention_lse(q, q_ld, q_col0, q_len, k, k_ld, k_col0, v, v_ld, v_col0, kv_len, kv_lens, batch, heads, causal,
                          out, out_ld, nullptr, dtype, stream);
}

extern "C" int mm_attention_lse(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                                int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                                const int32_t* kv_lens, int32_t batch, int32_t heads, int32_t causal, void* out,
                                int64_t out_ld, float* lse, int32_t dtype, void* stream) {
  return mm_attention_drop(q, q_ld, q_col0, q_len, k, k_ld, k_col0, v, v_ld, v_col0, kv_len, kv_lens, batch, heads, causal,
                           out, out_ld, lse, 0.f, 0, nullptr, 0, dtype, stream);
}

extern "C" int mm_attention_drop(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                                 int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                                 const int32_t* kv_lens, int32_t batch, int32_t heads, int32_t causal, void* out,
                                 int64_t out_ld, float* lse, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                                 uint32_t site, int32_t dtype, void* stream) {
  if (!q || !k || !v || !out) return bad_arg("attention: null pointer");
  if (drop_p < 0.f || drop_p >= 1.f) return bad_arg("attention: dropout p in [0, 1)");
  if (batch <= 0 || q_len <= 0 || kv_len <= 0 || heads <= 0) return bad_arg("attention: extents");
  const int d = heads * AT_HD;
  if ((q_ld % 8) || (k_ld % 8) || (v_ld % 8) || (out_ld % 8) || out_ld < d || q_ld < q_col0 + d || k_ld < k_col0 + d ||
      v_ld < v_col0 + d || (q_col0 % 8) || (k_col0 % 8) || (v_col0 % 8))
    return bad_arg("attention: leading dims / column offsets (head_dim must be 64)");
  if (causal && q_len != kv_len) return bad_arg("attention: a causal mask needs q_len == kv_len");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap mq, mk, mv, mout;
  int rc = make_tmap_3d(&mq, q, f16, (uint64_t)q_ld, (uint64_t)q_len, (uint64_t)batch, (uint64_t)q_ld,
                        (uint64_t)q_len * q_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mk, k, f16, (uint64_t)k_ld, (uint64_t)kv_len, (uint64_t)batch, (uint64_t)k_ld,
                    (uint64_t)kv_len * k_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mv, v, f16, (uint64_t)v_ld, (uint64_t)kv_len, (uint64_t)batch, (uint64_t)v_ld,
                    (uint64_t)kv_len * v_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mout, out, f16, (uint64_t)d, (uint64_t)q_len, (uint64_t)batch, (uint64_t)out_ld,
                    (uint64_t)q_len * out_ld, 128);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (drop_p > 0.f) {
    AttnDrop dr{drop_p, site, seed, reinterpret_cast<const unsigned long long*>(seed_dev), (kv_len + 63) / 64 * 64};
    return f16 ? launch_attn_long<__half, true>(mq, mk, mv, mout, kv_lens, batch, q_len, kv_len, q_col0, k_col0, v_col0,
                                                 causal != 0, heads, lse, s, dr)
               : launch_attn_long<__nv_bfloat16, true>(mq, mk, mv, mout, kv_lens, batch, q_len, kv_len, q_col0, k_col0,
                                                         v_col0, causal != 0, heads, lse, s, dr);
  }
  return f16 ? launch_attn_long<__half>(mq, mk, mv, mout, kv_lens, batch, q_len, kv_len, q_col0, k_col0, v_col0,
                                         causal != 0, heads, lse, s)
             : launch_attn_long<__nv_bfloat16>(mq, mk, mv, mout, kv_lens, batch, q_len, kv_len, q_col0, k_col0,
                                                 v_col0, causal != 0, heads, lse, s);
}
