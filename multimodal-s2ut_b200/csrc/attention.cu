// Multi-head self-attention core on tcgen05: O = softmax_fp32(Q K^T + key-padding mask) V, head_dim 64.
//
// One CTA (128 threads) = one (utterance, head, 128-query tile).  Keys are processed in chunks of 256:
//   S = Q K^T      tcgen05.mma M=128 N=256 K=64, fp32 scores in TMEM columns [0,256)
//   softmax        thread i owns score row i (TMEM lane i): no cross-thread reductions at all
//   P              written as 16-bit into shared memory in the 128B-swizzled K-major layout UMMA expects
//   O += P V       tcgen05.mma M=128 N=64 K=256 (A = P from smem, B = V^T tile), O in TMEM columns [256,320)
// For T <= 256 (utterances up to ~10 s) there is a single chunk and S is computed once.  Longer sequences
// use a two-sweep schedule (sweep 1: row maxima only; sweep 2: exp / P V) so O never needs rescaling.
// q arrives pre-scaled by head_dim^-0.5 (fused into the QKV GEMM epilogue), V arrives transposed
// ([B][d][T_pad], zero beyond T) so both MMAs use K-major operands loaded by TMA.
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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
  constexpr uint32_t idesc_o = umma_idesc(AT_BM, AT_HD, OpTraits<OpT>::fmt);

  auto load_and_scores = [&](int c, bool with_q, bool with_v) {
    if (tid == 0) {
      mbar_expect_tx(bar_tma, (with_q ? AT_Q_BYTES : 0) + AT_K_BYTES + (with_v ? AT_V_BYTES : 0));
      if (with_q) tma_load_3d(sQ, &mapQK, bar_tma, h * AT_HD, qt * AT_BM, b);
      tma_load_3d(sK, &mapQK, bar_tma, d_model + h * AT_HD, c * AT_KC, b);
      tma_load_3d(sK + AT_K_BYTES / 2, &mapQK, bar_tma, d_model + h * AT_HD, c * AT_KC + 128, b);
      if (with_v) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tma_load_3d(sV + j * (AT_V_BYTES / 4), &mapVT, bar_tma, c * AT_KC + j * 64, h * AT_HD, b);
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
                   bdesc + (uint64_t)((j * (AT_V_BYTES / 4)) >> 4) + 2 * kk, idesc_o, (c | j | kk) != 0);
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

}  // namespace mm

using namespace mm;

extern "C" int mm_self_attention(const void* qk, int64_t qk_ld, const void* vt, int64_t vt_ld, const int32_t* seq_lens,
                                 int32_t batch, int32_t seq, int32_t heads, void* out, int64_t out_ld, int32_t dtype,
                                 void* stream) {
  if (!qk || !vt || !seq_lens || !out) return bad_arg("self_attention: null pointer");
  if (batch <= 0 || seq <= 0 || heads <= 0) return bad_arg("self_attention: extents");
  const int d = heads * AT_HD;
  if (qk_ld < 2 * d || (qk_ld % 8) || (vt_ld % 8) || vt_ld < seq || (out_ld % 8) || out_ld < d)
    return bad_arg("self_attention: leading dims (head_dim must be 64)");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap mqk, mvt;
  int rc = make_tmap_3d(&mqk, qk, f16, (uint64_t)(2 * d), (uint64_t)seq, (uint64_t)batch, (uint64_t)qk_ld,
                        (uint64_t)seq * qk_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&mvt, vt, f16, (uint64_t)vt_ld, (uint64_t)d, (uint64_t)batch, (uint64_t)vt_ld,
                    (uint64_t)d * vt_ld, 64);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return f16 ? launch_attn<__half>(mqk, mvt, seq_lens, batch, seq, heads, d, out, out_ld, s)
             : launch_attn<__nv_bfloat16>(mqk, mvt, seq_lens, batch, seq, heads, d, out, out_ld, s);
}
