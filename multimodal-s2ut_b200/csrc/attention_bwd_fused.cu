// Fused attention backward for sequences of up to 256 positions (the encoder's self-attention at 10 s utterances:
// T = 250), one (sequence, head) per work item, everything between the inputs and dq / dk / dv on chip:
//
//     S = q k^T         P = exp(S - lse)  (masked keys: 0)         dV = P^T dO
//     dP = dO v^T       dS = P o (dP - rowsum(dO o O))             dK = dS^T q        dQ = dS k * head_dim^-0.5
//
// (autograd of fairseq's MultiheadAttention in the reference's training step).  mm_attention_bwd_scores
// + 3 x mm_heads_gemm wrote P and dS to HBM as 16-bit matrices and read them back three times (134 MB written + 201 MB
// read per encoder layer at 64 x 10 s); here P and dS live in shared memory for the length of one 128 x 128 step and
// the three output products accumulate in TMEM.
//
// TMEM (all 512 columns): S 128 | dP 128 | dV_c 64 | dK_c 64 | dQ_0 64 | dQ_1 64  -- which is why the key chunks are
// the OUTER loop (dV_c / dK_c of one 128-key chunk are complete after its <= 2 query tiles, dQ of both query tiles
// accumulates over the <= 2 chunks) and why longer sequences stay on the two-kernel path.
//
// Roles (320 threads, persistent over (sequence, head) items):
//   warp 8     TMA producer: q, dO, O tiles of both query tiles and k, v of both key chunks, once per item (O lands
//              in the P / dS staging area: it is only needed for delta before the first step)
//   warp 9     MMA issuer.  Step (c, qt): S, dP (SS, M128 N128 K64) -> consumers -> P, dS slabs in shared memory ->
//              dV_c += P^T dO_qt, dK_c += dS^T q_qt (A = the slabs read MN-major, B = the dO / q tile read MN-major),
//              dQ_qt += dS k_c (A = the slabs K-major, B = the k tile MN-major), all M128 N64.  The S / dP products of
//              step i+1 are issued before the output products of step i, so they run under the consumers' arithmetic.
//   warps 0-7  consumers, two threads per query row (64 keys each): delta, then per step tcgen05.ld -> exp2 / multiply
//              -> 16-bit slabs; epilogues: dV_c / dK_c after a chunk's last query tile, dQ at the end of the item
//              (tcgen05.ld -> 16-bit -> staging slab -> TMA store into the q | k | v gradient layout).
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int AF_BM = 128, AF_HD = 64, AF_KC = 128;
constexpr int AF_TILE = AF_BM * AF_HD * 2;      // 16 KB
constexpr int AF_THREADS = 320;
// q[2] | dO[2] | k[2] | v[2] | P slab 0,1 | dS slab 0,1 | out slab | barriers + row constants
constexpr int AF_SMEM_BYTES = 8 * AF_TILE + 4 * AF_TILE + AF_TILE + 2048 + 1024;
static_assert(AF_SMEM_BYTES <= 232448, "shared memory budget");

struct AfDev {
  int len, H, nq, nc, n_items;
  int q_col0, k_col0, v_col0;
  const int* kv_lens;        // [batch] valid keys, or NULL
  const float* lse;          // [batch][H][len]
  float scale;               // head_dim^-0.5 on dQ
  // attention dropout (training): the forward pass multiplied P by keep / (1 - p) (mm_self_attention_drop); the same
  // mask is regenerated here: dV takes the dropped P, dS = P o (dP o keep / (1 - p) - delta)
  float drop_p;
  unsigned drop_site;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
  int drop_tp;               // row length of the mask index space: round_up(len, 64)
};

__device__ __forceinline__ float af_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename OpT, bool DROP>
__global__ void __launch_bounds__(AF_THREADS, 1)
attention_bwd_fused_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                           const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO,
                           const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapDQ,
                           const __grid_constant__ CUtensorMap mapDK, const __grid_constant__ CUtensorMap mapDV,
                           const AfDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // [2 query tiles]
  uint8_t* sDO = sQ + 2 * AF_TILE;          // [2]
  uint8_t* sK = sDO + 2 * AF_TILE;          // [2 key chunks]
  uint8_t* sV = sK + 2 * AF_TILE;           // [2]
  uint8_t* sStage = sV + 2 * AF_TILE;       // P slab 0, P slab 1, dS slab 0, dS slab 1 (O tiles 0, 1 before the first step)
  uint8_t* sOut = sStage + 4 * AF_TILE;     // epilogue staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + AF_TILE);
  uint64_t* in_full = bars;           // TMA -> MMA + consumers (per item)
  uint64_t* in_empty = bars + 1;      // last MMAs of the item done -> TMA
  uint64_t* sdp_full = bars + 2;      // S, dP of a step in TMEM -> consumers
  uint64_t* sdp_empty = bars + 3;     // 8 consumer warps hold S, dP in registers -> MMA
  uint64_t* slab_full = bars + 4;     // 8 consumer warps wrote P, dS of a step -> MMA
  uint64_t* slab_empty = bars + 5;    // output MMAs of a step done reading the slabs -> consumers
  uint64_t* acc_full = bars + 6;      // dV_c, dK_c complete -> consumers
  uint64_t* acc_empty = bars + 7;     // 8 consumer warps drained dV_c, dK_c -> MMA
  uint64_t* dq_full = bars + 8;       // dQ of the item complete -> consumers
  uint64_t* dq_empty = bars + 9;      // 8 consumer warps drained dQ -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* x_delta = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);   // [2 halves][128 rows]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr float L2E = 1.4426950408889634f;
  // TMEM columns
  constexpr uint32_t C_S = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapDO);
    tma_prefetch_desc(&mapO);
    mbar_init(in_full, 1);
    mbar_init(in_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 8);
    mbar_init(slab_full, 8);
    mbar_init(slab_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        const int h = item % p.H, b = item / p.H;
        mbar_wait(in_empty, (i & 1) ^ 1);
        mbar_expect_tx(in_full, (3 * p.nq + 2 * p.nc) * AF_TILE);
        for (int qt = 0; qt < p.nq; ++qt) {
          tma_load_3d(sQ + qt * AF_TILE, &mapQ, in_full, p.q_col0 + h * AF_HD, qt * AF_BM, b);
          tma_load_3d(sDO + qt * AF_TILE, &mapDO, in_full, h * AF_HD, qt * AF_BM, b);
          tma_load_3d(sStage + qt * AF_TILE, &mapO, in_full, h * AF_HD, qt * AF_BM, b);
        }
        for (int c = 0; c < p.nc; ++c) {
          tma_load_3d(sK + c * AF_TILE, &mapK, in_full, p.k_col0 + h * AF_HD, c * AF_KC, b);
          tma_load_3d(sV + c * AF_TILE, &mapV, in_full, p.v_col0 + h * AF_HD, c * AF_KC, b);
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(AF_BM, AF_KC, OpTraits<OpT>::fmt);                              // S, dP
      constexpr uint32_t idesc_kv = umma_idesc(AF_BM, AF_HD, OpTraits<OpT>::fmt) | (1u << 15) | (1u << 16);   // dV, dK
      constexpr uint32_t idesc_q = umma_idesc(AF_BM, AF_HD, OpTraits<OpT>::fmt) | (1u << 16);                 // dQ
      uint32_t steps = 0, accs = 0;     // running counters -> mbarrier parities
      for (int i = 0; i < n_local; ++i) {
        mbar_wait(in_full, i & 1);
        mbar_wait(dq_empty, (i & 1) ^ 1);          // the previous item's dQ has been drained
        tc_fence_after();
        int pc = -1, pqt = 0;
        uint32_t pstep = 0;
        // the output products of step (c, qt) from its P / dS slabs
        auto outputs = [&](int c, int qt, uint32_t st) {
          mbar_wait(slab_full, st & 1);
          if (qt == 0) mbar_wait(acc_empty, (accs & 1) ^ 1);      // the previous chunk's dV / dK have been drained
          tc_fence_after();
          const uint64_t p_mn = umma_desc_sw128_mn(smem_u32(sStage), AF_TILE);
          const uint64_t ds_mn = umma_desc_sw128_mn(smem_u32(sStage + 2 * AF_TILE), AF_TILE);
          const uint64_t do_b = umma_desc_sw128_mn(smem_u32(sDO + qt * AF_TILE), AF_TILE);
          const uint64_t q_b = umma_desc_sw128_mn(smem_u32(sQ + qt * AF_TILE), AF_TILE);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)      // contraction over the 128 query rows of the tile
            umma_f16(tmem_base + C_DV, p_mn + 128ull * kk, do_b + 128ull * kk, idesc_kv, (qt | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16(tmem_base + C_DK, ds_mn + 128ull * kk, q_b + 128ull * kk, idesc_kv, (qt | kk) != 0);
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {    // contraction over the 128 keys of the chunk: two 64-key slabs
            const uint64_t ds_k = umma_desc_sw128(smem_u32(sStage + (2 + s2) * AF_TILE));
            const uint64_t k_b = umma_desc_sw128_mn(smem_u32(sK + c * AF_TILE + s2 * 64 * 128), AF_TILE);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_base + C_DQ + qt * AF_HD, ds_k + 2ull * kk, k_b + 128ull * kk, idesc_q, (c | s2 | kk) != 0);
          }
          umma_commit(slab_empty);
          if (qt == p.nq - 1) {
            umma_commit(acc_full);
            ++accs;
          }
        };
        for (int c = 0; c < p.nc; ++c) {
          for (int qt = 0; qt < p.nq; ++qt) {
            mbar_wait(sdp_empty, (steps & 1) ^ 1);       // the consumers hold the previous step's S / dP in registers
            tc_fence_after();
            const uint64_t qd = umma_desc_sw128(smem_u32(sQ + qt * AF_TILE));
            const uint64_t dod = umma_desc_sw128(smem_u32(sDO + qt * AF_TILE));
            const uint64_t kd = umma_desc_sw128(smem_u32(sK + c * AF_TILE));
            const uint64_t vd = umma_desc_sw128(smem_u32(sV + c * AF_TILE));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + C_S, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + C_DP, dod + 2 * kk, vd + 2 * kk, idesc_s, kk != 0);
            umma_commit(sdp_full);
            if (pc >= 0) outputs(pc, pqt, pstep);
            pc = c, pqt = qt, pstep = steps;
            ++steps;
          }
        }
        outputs(pc, pqt, pstep);
        umma_commit(dq_full);
        umma_commit(in_empty);      // every MMA that reads this item's tiles and slabs has completed
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int hf = warp >> 2;                  // which 64 keys of a chunk / which 32 of the 64 head columns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>((warp & 3) * 32) << 16;
    auto group_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const bool elected = tid == 0;
    uint32_t steps = 0, accs = 0;
    const unsigned long long drop_seed = DROP ? p.drop_seed + (p.drop_seed_dev ? *p.drop_seed_dev : 0ull) : 0ull;
    const unsigned drop_thr = dropout_threshold(p.drop_p);
    const float drop_inv = 1.0f / (1.0f - p.drop_p);
    // 64 columns of a 128-row fp32 accumulator -> 16-bit -> staging slab -> TMA store at (col, row0, b)
    auto store_tile = [&](uint32_t tcol, float scale, const CUtensorMap* m, int col, int row0, int b, uint64_t* drained) {
      uint32_t r[32];
      tmem_ld32(tmem_base + tcol + t_lane + 32 * hf, r);
      if (elected) bulk_wait_read<0>();         // the previous store has finished reading the staging slab
      tmem_ld_wait();
      if (drained != nullptr) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(drained);
      }
      group_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 q;
        q.x = OpTraits<OpT>::pack2(__uint_as_float(r[8 * k + 0]) * scale, __uint_as_float(r[8 * k + 1]) * scale);
        q.y = OpTraits<OpT>::pack2(__uint_as_float(r[8 * k + 2]) * scale, __uint_as_float(r[8 * k + 3]) * scale);
        q.z = OpTraits<OpT>::pack2(__uint_as_float(r[8 * k + 4]) * scale, __uint_as_float(r[8 * k + 5]) * scale);
        q.w = OpTraits<OpT>::pack2(__uint_as_float(r[8 * k + 6]) * scale, __uint_as_float(r[8 * k + 7]) * scale);
        *reinterpret_cast<uint4*>(sOut + row * 128 + (((4 * hf + k) ^ (row & 7)) << 4)) = q;
      }
      fence_proxy_async_smem();
      group_sync();
      if (elected) {
        tma_store_3d(m, sOut, col, row0, b);
        bulk_commit();
      }
    };
    for (int i = 0; i < n_local; ++i) {
      const int item = blockIdx.x + i * gridDim.x;
      const int h = item % p.H, b = item / p.H;
      const int len = p.kv_lens ? max(0, min(p.kv_lens[b], p.len)) : p.len;
      // ---- delta = rowsum(dO o O) over the head's 64 columns (this thread: 32 of them) and lse, per query tile ----
      mbar_wait(in_full, i & 1);
      float delta[2] = {0.f, 0.f}, lb[2] = {0.f, 0.f};
      for (int qt = 0; qt < p.nq; ++qt) {
        float dsum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int chunk = ((4 * hf + k) ^ (row & 7)) << 4;
          const uint4 a = *reinterpret_cast<const uint4*>(sDO + qt * AF_TILE + row * 128 + chunk);
          const uint4 o = *reinterpret_cast<const uint4*>(sStage + qt * AF_TILE + row * 128 + chunk);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const OpT* ap = reinterpret_cast<const OpT*>(&aw[j]);
            const OpT* op = reinterpret_cast<const OpT*>(&ow[j]);
            dsum = fmaf(OpTraits<OpT>::to_float(ap[0]), OpTraits<OpT>::to_float(op[0]), dsum);
            dsum = fmaf(OpTraits<OpT>::to_float(ap[1]), OpTraits<OpT>::to_float(op[1]), dsum);
          }
        }
        x_delta[hf * AF_BM + row] = dsum;
        group_sync();
        delta[qt] = dsum + x_delta[(hf ^ 1) * AF_BM + row];
        group_sync();                               // x_delta is rewritten by the next query tile
        const int qrow = qt * AF_BM + row;
        lb[qt] = (qrow < p.len ? __ldg(p.lse + ((long long)b * p.H + h) * p.len + qrow) : 0.f) * L2E;
      }
      // (the O tiles in the staging area are dead from here on: the first step's P / dS overwrite them)

      for (int c = 0; c < p.nc; ++c) {
        for (int qt = 0; qt < p.nq; ++qt) {
          const int qrow = qt * AF_BM + row;
          const int k0 = c * AF_KC + 64 * hf;
          const int nv = qrow < p.len ? min(64, max(0, len - k0)) : 0;     // visible keys among my 64
          const float dl = delta[qt], lbq = lb[qt];
          uint8_t* sP = sStage + hf * AF_TILE;
          uint8_t* sD = sStage + 2 * AF_TILE + hf * AF_TILE;
          mbar_wait(sdp_full, steps & 1);
          mbar_wait(slab_empty, (steps & 1) ^ 1);     // the previous step's output MMAs have read the slabs
          tc_fence_after();
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            uint32_t pk[16], dk[16];
            uint32_t rs[32], rd[32];
            const uint32_t t_s = tmem_base + C_S + t_lane + 64 * hf + 32 * half;
            tmem_ld32(t_s, rs);
            tmem_ld32(t_s + (C_DP - C_S), rd);
            tmem_ld_wait();
            if (half == 1) {        // S and dP of this step are in registers: the next step's products may overwrite them
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(sdp_empty);
            }
            const int nvh = nv - 32 * half;
            if constexpr (DROP) {
              // mask index of (query row, first key of this half): ((b H + h) Tp + q) Tp + key, four elements per hash
              const unsigned long long i4 =
                  ((unsigned long long)((long long)(b * p.H + h) * p.drop_tp + qrow) * (unsigned)p.drop_tp + (unsigned)(k0 + 32 * half)) >> 2;
#pragma unroll
              for (int g4 = 0; g4 < 8; ++g4) {
                const int k = 4 * g4;
                float pr[4], m4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  pr[e] = af_ex2(fmaf(__uint_as_float(rs[k + e]), L2E, -lbq));
                  pr[e] = k + e < nvh ? pr[e] : 0.f;
                }
                dropout_apply4(dropout_bits4(drop_seed, p.drop_site, i4 + g4), drop_thr, drop_inv, m4[0], m4[1], m4[2], m4[3]);
                float dd[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) dd[e] = pr[e] * (__uint_as_float(rd[k + e]) * m4[e] - dl);
                pk[2 * g4] = OpTraits<OpT>::pack2(pr[0] * m4[0], pr[1] * m4[1]);
                pk[2 * g4 + 1] = OpTraits<OpT>::pack2(pr[2] * m4[2], pr[3] * m4[3]);
                dk[2 * g4] = OpTraits<OpT>::pack2(dd[0], dd[1]);
                dk[2 * g4 + 1] = OpTraits<OpT>::pack2(dd[2], dd[3]);
              }
            } else {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              float p0 = af_ex2(fmaf(__uint_as_float(rs[k]), L2E, -lbq));
              float p1 = af_ex2(fmaf(__uint_as_float(rs[k + 1]), L2E, -lbq));
              p0 = k < nvh ? p0 : 0.f;
              p1 = k + 1 < nvh ? p1 : 0.f;
              const float d0 = p0 * (__uint_as_float(rd[k]) - dl);
              const float d1 = p1 * (__uint_as_float(rd[k + 1]) - dl);
              pk[k >> 1] = OpTraits<OpT>::pack2(p0, p1);
              dk[k >> 1] = OpTraits<OpT>::pack2(d0, d1);
            }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int chunk = ((4 * half + k) ^ (row & 7)) << 4;
              *reinterpret_cast<uint4*>(sP + row * 128 + chunk) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
              *reinterpret_cast<uint4*>(sD + row * 128 + chunk) = make_uint4(dk[4 * k], dk[4 * k + 1], dk[4 * k + 2], dk[4 * k + 3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(slab_full);
          ++steps;
          if (qt == p.nq - 1) {
            // ---- dV_c, dK_c: rows = the chunk's keys ----
            mbar_wait(acc_full, accs & 1);
            tc_fence_after();
            store_tile(C_DV, 1.0f, &mapDV, p.v_col0 + h * AF_HD, c * AF_KC, b, nullptr);
            store_tile(C_DK, 1.0f, &mapDK, p.k_col0 + h * AF_HD, c * AF_KC, b, acc_empty);
            ++accs;
          }
        }
      }
      // ---- dQ of both query tiles ----
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      for (int qt = 0; qt < p.nq; ++qt)
        store_tile(C_DQ + qt * AF_HD, p.scale, &mapDQ, p.q_col0 + h * AF_HD, qt * AF_BM, b,
                   qt == p.nq - 1 ? dq_empty : nullptr);
    }
    if (elected) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// General variant: any query / key length, causal mask, q and k|v in different tensors (the decoder's causal
// self-attention and its encoder attention; the encoder beyond 256 positions).  Same steps and the same TMEM plan, but
//   * query tiles are taken in PAIRS (dQ of two tiles is what fits TMEM): for pair: for key chunk: for tile in pair
//   * dV_c / dK_c are complete only after the last pair: pairs before the last leave their fp32 partial in a scratch
//     buffer of the CTA (L2-resident, 64 KB per chunk), later pairs add it back -- the same CTA in a fixed order, so the
//     result stays deterministic
//   * q / dO / O tiles are loaded per pair, (k_c | v_c) through a 2-stage ring per (pair, chunk)
//   * causal: steps whose chunk lies entirely above the tile's diagonal are skipped by all three roles alike.
// =====================================================================================================================
struct AgDev {
  int q_len, kv_len, H, nq, nc, npairs, n_items, causal;
  int q_col0, k_col0, v_col0, dq_col0, dk_col0, dv_col0;
  const int* kv_lens;        // [batch] valid keys, or NULL
  const float* lse;          // [batch][H][q_len]
  float* scratch;            // [gridDim.x][nc][2][128 * 64] fp32
  float scale;
  // attention dropout (mm_attention_drop wrote O with the mask): index ((b H + h) drop_lp + q) drop_tp + key
  float drop_p;
  unsigned drop_site;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
  int drop_lp, drop_tp;      // round_up(q_len, 64), round_up(kv_len, 64)
};

template <typename OpT, bool DROP>
__global__ void __launch_bounds__(AF_THREADS, 1)
attention_bwd_general_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                             const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO,
                             const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapDQ,
                             const __grid_constant__ CUtensorMap mapDK, const __grid_constant__ CUtensorMap mapDV,
                             const AgDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // [2 tiles of the pair]
  uint8_t* sDO = sQ + 2 * AF_TILE;          // [2]
  uint8_t* sKV = sDO + 2 * AF_TILE;         // [2 stages] (k_c 16 KB | v_c 16 KB)
  uint8_t* sStage = sKV + 4 * AF_TILE;      // P slab 0, 1, dS slab 0, 1 (O tiles of the pair before its first step)
  uint8_t* sOut = sStage + 4 * AF_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + AF_TILE);
  uint64_t* pair_full = bars;         // TMA (q, dO, O of a pair) -> MMA + consumers
  uint64_t* pair_empty = bars + 1;    // last MMAs of the pair done -> TMA
  uint64_t* sdp_full = bars + 2;
  uint64_t* sdp_empty = bars + 3;
  uint64_t* slab_full = bars + 4;
  uint64_t* slab_empty = bars + 5;
  uint64_t* acc_full = bars + 6;
  uint64_t* acc_empty = bars + 7;
  uint64_t* dq_full = bars + 8;
  uint64_t* dq_empty = bars + 9;
  uint64_t* kv_full = bars + 10;      // [2]
  uint64_t* kv_empty = bars + 12;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* x_delta = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr float L2E = 1.4426950408889634f;
  constexpr uint32_t C_S = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapDO);
    tma_prefetch_desc(&mapO);
    mbar_init(pair_full, 1);
    mbar_init(pair_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 8);
    mbar_init(slab_full, 8);
    mbar_init(slab_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  // step (chunk c, query tile qt) carries at least one visible (query, key) pair
  auto step_live = [&](int c, int qt) {
    if (qt >= p.nq) return false;
    return !(p.causal && c * AF_KC > min(qt * AF_BM + AF_BM, p.q_len) - 1);
  };
  auto chunk_live = [&](int pr, int c) { return step_live(c, 2 * pr) || step_live(c, 2 * pr + 1); };
  auto last_tile = [&](int pr, int c) { return step_live(c, 2 * pr + 1) ? 1 : 0; };     // last live tile of (pair, chunk)
  auto first_tile = [&](int pr, int c) { return step_live(c, 2 * pr) ? 0 : 1; };

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t pn = 0, kvn = 0;
      for (int i = 0; i < n_local; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        const int h = item % p.H, b = item / p.H;
        for (int pr = 0; pr < p.npairs; ++pr, ++pn) {
          const int nt = min(2, p.nq - 2 * pr);
          mbar_wait(pair_empty, (pn & 1) ^ 1);
          mbar_expect_tx(pair_full, 3 * nt * AF_TILE);
          for (int t = 0; t < nt; ++t) {
            const int qt = 2 * pr + t;
            tma_load_3d(sQ + t * AF_TILE, &mapQ, pair_full, p.q_col0 + h * AF_HD, qt * AF_BM, b);
            tma_load_3d(sDO + t * AF_TILE, &mapDO, pair_full, h * AF_HD, qt * AF_BM, b);
            tma_load_3d(sStage + t * AF_TILE, &mapO, pair_full, h * AF_HD, qt * AF_BM, b);
          }
          for (int c = 0; c < p.nc; ++c) {
            if (!chunk_live(pr, c)) continue;
            const uint32_t st = kvn & 1;
            mbar_wait(&kv_empty[st], ((kvn >> 1) & 1) ^ 1);
            mbar_expect_tx(&kv_full[st], 2 * AF_TILE);
            tma_load_3d(sKV + st * 2 * AF_TILE, &mapK, &kv_full[st], p.k_col0 + h * AF_HD, c * AF_KC, b);
            tma_load_3d(sKV + st * 2 * AF_TILE + AF_TILE, &mapV, &kv_full[st], p.v_col0 + h * AF_HD, c * AF_KC, b);
            ++kvn;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(AF_BM, AF_KC, OpTraits<OpT>::fmt);
      constexpr uint32_t idesc_kv = umma_idesc(AF_BM, AF_HD, OpTraits<OpT>::fmt) | (1u << 15) | (1u << 16);
      constexpr uint32_t idesc_q = umma_idesc(AF_BM, AF_HD, OpTraits<OpT>::fmt) | (1u << 16);
      uint32_t steps = 0, accs = 0, pn = 0, kvn = 0;
      for (int i = 0; i < n_local; ++i) {
        for (int pr = 0; pr < p.npairs; ++pr, ++pn) {
          mbar_wait(pair_full, pn & 1);
          mbar_wait(dq_empty, (pn & 1) ^ 1);         // the previous pair's dQ has been drained
          tc_fence_after();
          bool dq_started[2] = {false, false};
          int pc = -1, pt = 0, pfirst = 0, plast = 0;
          uint32_t pstep = 0, pst = 0;
          auto outputs = [&](int t, bool first, bool last, uint32_t step, uint32_t st) {
            mbar_wait(slab_full, step & 1);
            if (first) mbar_wait(acc_empty, (accs & 1) ^ 1);
            tc_fence_after();
            const uint64_t p_mn = umma_desc_sw128_mn(smem_u32(sStage), AF_TILE);
            const uint64_t ds_mn = umma_desc_sw128_mn(smem_u32(sStage + 2 * AF_TILE), AF_TILE);
            const uint64_t do_b = umma_desc_sw128_mn(smem_u32(sDO + t * AF_TILE), AF_TILE);
            const uint64_t q_b = umma_desc_sw128_mn(smem_u32(sQ + t * AF_TILE), AF_TILE);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16(tmem_base + C_DV, p_mn + 128ull * kk, do_b + 128ull * kk, idesc_kv, (!first) || kk != 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16(tmem_base + C_DK, ds_mn + 128ull * kk, q_b + 128ull * kk, idesc_kv, (!first) || kk != 0);
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const uint64_t ds_k = umma_desc_sw128(smem_u32(sStage + (2 + s2) * AF_TILE));
              const uint64_t k_b = umma_desc_sw128_mn(smem_u32(sKV + st * 2 * AF_TILE + s2 * 64 * 128), AF_TILE);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(tmem_base + C_DQ + t * AF_HD, ds_k + 2ull * kk, k_b + 128ull * kk, idesc_q,
                         dq_started[t] || (s2 | kk) != 0);
            }
            dq_started[t] = true;
            umma_commit(slab_empty);
            if (last) {
              umma_commit(acc_full);
              ++accs;
              umma_commit(&kv_empty[st]);        // the chunk's k / v tiles are no longer read
            }
          };
          for (int c = 0; c < p.nc; ++c) {
            if (!chunk_live(pr, c)) continue;
            const uint32_t st = kvn & 1;
            mbar_wait(&kv_full[st], (kvn >> 1) & 1);
            ++kvn;
            const int ft = first_tile(pr, c), lt = last_tile(pr, c);
            for (int t = ft; t <= lt; ++t) {
              mbar_wait(sdp_empty, (steps & 1) ^ 1);
              tc_fence_after();
              const uint64_t qd = umma_desc_sw128(smem_u32(sQ + t * AF_TILE));
              const uint64_t dod = umma_desc_sw128(smem_u32(sDO + t * AF_TILE));
              const uint64_t kd = umma_desc_sw128(smem_u32(sKV + st * 2 * AF_TILE));
              const uint64_t vd = umma_desc_sw128(smem_u32(sKV + st * 2 * AF_TILE + AF_TILE));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + C_S, qd + 2 * kk, kd + 2 * kk, idesc_s, kk != 0);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base + C_DP, dod + 2 * kk, vd + 2 * kk, idesc_s, kk != 0);
              umma_commit(sdp_full);
              if (pc >= 0) outputs(pt, pfirst != 0, plast != 0, pstep, pst);
              pc = c, pt = t, pfirst = (t == ft), plast = (t == lt), pstep = steps, pst = st;
              ++steps;
            }
          }
          if (pc >= 0) outputs(pt, pfirst != 0, plast != 0, pstep, pst);
          umma_commit(dq_full);
          umma_commit(pair_empty);
        }
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int hf = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>((warp & 3) * 32) << 16;
    auto group_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const bool elected = tid == 0;
    uint32_t steps = 0, accs = 0, pn = 0;
    const unsigned long long drop_seed = DROP ? p.drop_seed + (p.drop_seed_dev ? *p.drop_seed_dev : 0ull) : 0ull;
    const unsigned drop_thr = dropout_threshold(p.drop_p);
    const float drop_inv = 1.0f / (1.0f - p.drop_p);
    float* my_scratch = p.scratch + (size_t)blockIdx.x * p.nc * 2 * (AF_BM * AF_HD);
    // 32 of the 64 columns of this thread's accumulator row (+ the scratch partial) -> scratch (fp32) or 16-bit store
    auto drain_tile = [&](uint32_t tcol, float scale, float* part, bool add_part, bool to_scratch, const CUtensorMap* m,
                          int col, int row0, int b, uint64_t* drained) {
      uint32_t r[32];
      tmem_ld32(tmem_base + tcol + t_lane + 32 * hf, r);
      if (!to_scratch && elected) bulk_wait_read<0>();
      tmem_ld_wait();
      if (drained != nullptr) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(drained);
      }
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
      float4* pp = reinterpret_cast<float4*>(part + row * AF_HD + 32 * hf);
      if (add_part) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 q = __ldcg(pp + k);
          v[4 * k] += q.x, v[4 * k + 1] += q.y, v[4 * k + 2] += q.z, v[4 * k + 3] += q.w;
        }
      }
      if (to_scratch) {
#pragma unroll
        for (int k = 0; k < 8; ++k) __stcg(pp + k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
        return;
      }
      group_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 q;
        q.x = OpTraits<OpT>::pack2(v[8 * k + 0] * scale, v[8 * k + 1] * scale);
        q.y = OpTraits<OpT>::pack2(v[8 * k + 2] * scale, v[8 * k + 3] * scale);
        q.z = OpTraits<OpT>::pack2(v[8 * k + 4] * scale, v[8 * k + 5] * scale);
        q.w = OpTraits<OpT>::pack2(v[8 * k + 6] * scale, v[8 * k + 7] * scale);
        *reinterpret_cast<uint4*>(sOut + row * 128 + (((4 * hf + k) ^ (row & 7)) << 4)) = q;
      }
      fence_proxy_async_smem();
      group_sync();
      if (elected) {
        tma_store_3d(m, sOut, col, row0, b);
        bulk_commit();
      }
    };
    for (int i = 0; i < n_local; ++i) {
      const int item = blockIdx.x + i * gridDim.x;
      const int h = item % p.H, b = item / p.H;
      const int len = p.kv_lens ? max(0, min(p.kv_lens[b], p.kv_len)) : p.kv_len;
      for (int pr = 0; pr < p.npairs; ++pr, ++pn) {
        const int nt = min(2, p.nq - 2 * pr);
        mbar_wait(pair_full, pn & 1);
        float delta[2] = {0.f, 0.f}, lb[2] = {0.f, 0.f};
        for (int t = 0; t < nt; ++t) {
          float dsum = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int chunk = ((4 * hf + k) ^ (row & 7)) << 4;
            const uint4 a = *reinterpret_cast<const uint4*>(sDO + t * AF_TILE + row * 128 + chunk);
            const uint4 o = *reinterpret_cast<const uint4*>(sStage + t * AF_TILE + row * 128 + chunk);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const OpT* ap = reinterpret_cast<const OpT*>(&aw[j]);
              const OpT* op = reinterpret_cast<const OpT*>(&ow[j]);
              dsum = fmaf(OpTraits<OpT>::to_float(ap[0]), OpTraits<OpT>::to_float(op[0]), dsum);
              dsum = fmaf(OpTraits<OpT>::to_float(ap[1]), OpTraits<OpT>::to_float(op[1]), dsum);
            }
          }
          x_delta[hf * AF_BM + row] = dsum;
          group_sync();
          delta[t] = dsum + x_delta[(hf ^ 1) * AF_BM + row];
          group_sync();
          const int qrow = (2 * pr + t) * AF_BM + row;
          lb[t] = (qrow < p.q_len ? __ldg(p.lse + ((long long)b * p.H + h) * p.q_len + qrow) : 0.f) * L2E;
        }
        for (int c = 0; c < p.nc; ++c) {
          if (!chunk_live(pr, c)) continue;
          const int ft = first_tile(pr, c), lt = last_tile(pr, c);
          for (int t = ft; t <= lt; ++t) {
            const int qrow = (2 * pr + t) * AF_BM + row;
            const int k0 = c * AF_KC + 64 * hf;
            int nv = qrow < p.q_len ? min(64, max(0, len - k0)) : 0;
            if (p.causal) nv = min(nv, max(0, qrow + 1 - k0));
            const float dl = delta[t], lbq = lb[t];
            uint8_t* sP = sStage + hf * AF_TILE;
            uint8_t* sD = sStage + 2 * AF_TILE + hf * AF_TILE;
            mbar_wait(sdp_full, steps & 1);
            mbar_wait(slab_empty, (steps & 1) ^ 1);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              uint32_t pk[16], dk[16];
              uint32_t rs[32], rd[32];
              const uint32_t t_s = tmem_base + C_S + t_lane + 64 * hf + 32 * half;
              tmem_ld32(t_s, rs);
              tmem_ld32(t_s + (C_DP - C_S), rd);
              tmem_ld_wait();
              if (half == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sdp_empty);
              }
              const int nvh = nv - 32 * half;
              if constexpr (DROP) {
                const unsigned long long i4 =
                    ((unsigned long long)((long long)(b * p.H + h) * p.drop_lp + qrow) * (unsigned)p.drop_tp +
                     (unsigned)(k0 + 32 * half)) >> 2;
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                  const int k = 4 * g4;
                  float pr4[4], m4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    pr4[e] = af_ex2(fmaf(__uint_as_float(rs[k + e]), L2E, -lbq));
                    pr4[e] = k + e < nvh ? pr4[e] : 0.f;
                  }
                  dropout_apply4(dropout_bits4(drop_seed, p.drop_site, i4 + g4), drop_thr, drop_inv, m4[0], m4[1], m4[2], m4[3]);
                  float dd[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) dd[e] = pr4[e] * (__uint_as_float(rd[k + e]) * m4[e] - dl);
                  pk[2 * g4] = OpTraits<OpT>::pack2(pr4[0] * m4[0], pr4[1] * m4[1]);
                  pk[2 * g4 + 1] = OpTraits<OpT>::pack2(pr4[2] * m4[2], pr4[3] * m4[3]);
                  dk[2 * g4] = OpTraits<OpT>::pack2(dd[0], dd[1]);
                  dk[2 * g4 + 1] = OpTraits<OpT>::pack2(dd[2], dd[3]);
                }
              } else {
#pragma unroll
              for (int k = 0; k < 32; k += 2) {
                float p0 = af_ex2(fmaf(__uint_as_float(rs[k]), L2E, -lbq));
                float p1 = af_ex2(fmaf(__uint_as_float(rs[k + 1]), L2E, -lbq));
                p0 = k < nvh ? p0 : 0.f;
                p1 = k + 1 < nvh ? p1 : 0.f;
                const float d0 = p0 * (__uint_as_float(rd[k]) - dl);
                const float d1 = p1 * (__uint_as_float(rd[k + 1]) - dl);
                pk[k >> 1] = OpTraits<OpT>::pack2(p0, p1);
                dk[k >> 1] = OpTraits<OpT>::pack2(d0, d1);
              }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int chunk = ((4 * half + k) ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(sP + row * 128 + chunk) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                *reinterpret_cast<uint4*>(sD + row * 128 + chunk) = make_uint4(dk[4 * k], dk[4 * k + 1], dk[4 * k + 2], dk[4 * k + 3]);
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(slab_full);
            ++steps;
            if (t == lt) {
              // ---- dV_c, dK_c of this pair: add the earlier pairs' partial, store (last pair) or keep as the partial ----
              int first_pr = 0;
              while (!chunk_live(first_pr, c)) ++first_pr;
              const bool add_part = pr > first_pr, to_scratch = pr < p.npairs - 1;
              float* part = my_scratch + (size_t)c * 2 * (AF_BM * AF_HD);
              mbar_wait(acc_full, accs & 1);
              tc_fence_after();
              drain_tile(C_DV, 1.0f, part, add_part, to_scratch, &mapDV, p.dv_col0 + h * AF_HD, c * AF_KC, b, nullptr);
              drain_tile(C_DK, 1.0f, part + AF_BM * AF_HD, add_part, to_scratch, &mapDK, p.dk_col0 + h * AF_HD, c * AF_KC, b,
                         acc_empty);
              ++accs;
            }
          }
        }
        // ---- dQ of the pair's tiles ----
        mbar_wait(dq_full, pn & 1);
        tc_fence_after();
        for (int t = 0; t < nt; ++t)
          drain_tile(C_DQ + t * AF_HD, p.scale, nullptr, false, false, &mapDQ, p.dq_col0 + h * AF_HD, (2 * pr + t) * AF_BM, b,
                     t == nt - 1 ? dq_empty : nullptr);
      }
    }
    if (elected) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename OpT, bool DROP>
static int launch_attention_bwd_general(const CUtensorMap (&m)[8], const AgDev& p, int grid, cudaStream_t s) {
  auto kern = attention_bwd_general_kernel<OpT, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(attention_bwd_general)");
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(AF_THREADS);
  cfg.dynamicSmemBytes = AF_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], p);
  if (e != cudaSuccess) return fail(e, "attention_bwd_general_kernel launch");
  return 0;
}

template <typename OpT, bool DROP>
static int launch_attention_bwd_fused(const CUtensorMap (&m)[8], const AfDev& p, cudaStream_t s) {
  auto kern = attention_bwd_fused_kernel<OpT, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(attention_bwd_fused)");
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.n_items < kNumSMs ? p.n_items : kNumSMs);
  cfg.blockDim = dim3(AF_THREADS);
  cfg.dynamicSmemBytes = AF_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], p);
  if (e != cudaSuccess) return fail(e, "attention_bwd_fused_kernel launch");
  return 0;
}

}  // namespace mm

using namespace mm;

extern "C" int mm_attention_bwd_fused(const void* qkv, int64_t qkv_ld, int32_t q_col0, int32_t k_col0, int32_t v_col0,
                                      int32_t seq_len, const int32_t* kv_lens, int32_t batch, int32_t heads,
                                      const void* dout, int64_t do_ld, const void* out, int64_t o_ld, const float* lse,
                                      void* dqkv, int64_t dqkv_ld, int32_t dtype, void* stream) {
  return mm_attention_bwd_fused_drop(qkv, qkv_ld, q_col0, k_col0, v_col0, seq_len, kv_lens, batch, heads, dout, do_ld, out,
                                     o_ld, lse, dqkv, dqkv_ld, 0.f, 0, nullptr, 0, dtype, stream);
}

extern "C" int mm_attention_bwd_fused_drop(const void* qkv, int64_t qkv_ld, int32_t q_col0, int32_t k_col0,
                                           int32_t v_col0, int32_t seq_len, const int32_t* kv_lens, int32_t batch,
                                           int32_t heads, const void* dout, int64_t do_ld, const void* out, int64_t o_ld,
                                           const float* lse, void* dqkv, int64_t dqkv_ld, float drop_p, uint64_t seed,
                                           const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream) {
  if (drop_p < 0.f || drop_p >= 1.f) return bad_arg("attention_bwd_fused: dropout p in [0, 1)");
  if (!qkv || !dout || !out || !lse || !dqkv) return bad_arg("attention_bwd_fused: null pointer");
  if (batch <= 0 || seq_len <= 0 || heads <= 0) return bad_arg("attention_bwd_fused: extents");
  if (seq_len > 2 * AF_BM) return bad_arg("attention_bwd_fused: seq_len must be <= 256 (dQ of two query tiles in TMEM)");
  const int d = heads * AF_HD;
  const int hi = (q_col0 > k_col0 ? q_col0 : k_col0) > v_col0 ? (q_col0 > k_col0 ? q_col0 : k_col0) : v_col0;
  if ((qkv_ld % 8) || (dqkv_ld % 8) || (do_ld % 8) || (o_ld % 8) || qkv_ld < hi + d || dqkv_ld < hi + d || do_ld < d ||
      o_ld < d || (q_col0 % 8) || (k_col0 % 8) || (v_col0 % 8))
    return bad_arg("attention_bwd_fused: leading dims / column offsets (head_dim must be 64)");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap m[8];
  const uint64_t T = (uint64_t)seq_len, B = (uint64_t)batch;
  int rc = make_tmap_3d(&m[0], qkv, f16, (uint64_t)qkv_ld, T, B, (uint64_t)qkv_ld, T * qkv_ld, 128);
  if (rc) return rc;
  m[1] = m[0], m[2] = m[0];
  rc = make_tmap_3d(&m[3], dout, f16, (uint64_t)d, T, B, (uint64_t)do_ld, T * do_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[4], out, f16, (uint64_t)d, T, B, (uint64_t)o_ld, T * o_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[5], dqkv, f16, (uint64_t)dqkv_ld, T, B, (uint64_t)dqkv_ld, T * dqkv_ld, 128);
  if (rc) return rc;
  m[6] = m[5], m[7] = m[5];
  AfDev p;
  memset(&p, 0, sizeof(p));
  p.len = seq_len, p.H = heads;
  p.nq = (seq_len + AF_BM - 1) / AF_BM;
  p.nc = (seq_len + AF_KC - 1) / AF_KC;
  p.n_items = batch * heads;
  p.q_col0 = q_col0, p.k_col0 = k_col0, p.v_col0 = v_col0;
  p.kv_lens = kv_lens, p.lse = lse;
  p.scale = 0.125f;
  p.drop_p = drop_p, p.drop_site = site, p.drop_seed = seed;
  p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev), p.drop_tp = (seq_len + 63) / 64 * 64;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (drop_p > 0.f)
    return f16 ? launch_attention_bwd_fused<__half, true>(m, p, s) : launch_attention_bwd_fused<__nv_bfloat16, true>(m, p, s);
  return f16 ? launch_attention_bwd_fused<__half, false>(m, p, s) : launch_attention_bwd_fused<__nv_bfloat16, false>(m, p, s);
}

extern "C" int64_t mm_attention_bwd_general_scratch_floats(int32_t kv_len) {
  const int nc = (kv_len + AF_KC - 1) / AF_KC;
  return (int64_t)kNumSMs * nc * 2 * (AF_BM * AF_HD);
}

extern "C" int mm_attention_bwd_general(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k,
                                        int64_t k_ld, int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0,
                                        int32_t kv_len, const int32_t* kv_lens, int32_t batch, int32_t heads,
                                        int32_t causal, const void* dout, int64_t do_ld, const void* out, int64_t o_ld,
                                        const float* lse, void* dq, int64_t dq_ld, int32_t dq_col0, void* dk,
                                        int64_t dk_ld, int32_t dk_col0, void* dv, int64_t dv_ld, int32_t dv_col0,
                                        float* scratch, int32_t dtype, void* stream) {
  return mm_attention_bwd_general_drop(q, q_ld, q_col0, q_len, k, k_ld, k_col0, v, v_ld, v_col0, kv_len, kv_lens, batch, heads,
                                       causal, dout, do_ld, out, o_ld, lse, dq, dq_ld, dq_col0, dk, dk_ld, dk_col0, dv, dv_ld,
                                       dv_col0, scratch, 0.f, 0, nullptr, 0, dtype, stream);
}

extern "C" int mm_attention_bwd_general_drop(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k,
                                             int64_t k_ld, int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0,
                                             int32_t kv_len, const int32_t* kv_lens, int32_t batch, int32_t heads,
                                             int32_t causal, const void* dout, int64_t do_ld, const void* out, int64_t o_ld,
                                             const float* lse, void* dq, int64_t dq_ld, int32_t dq_col0, void* dk,
                                             int64_t dk_ld, int32_t dk_col0, void* dv, int64_t dv_ld, int32_t dv_col0,
                                             float* scratch, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                                             uint32_t site, int32_t dtype, void* stream) {
  if (!q || !k || !v || !dout || !out || !lse || !dq || !dk || !dv || !scratch) return bad_arg("attention_bwd_general: null pointer");
  if (drop_p < 0.f || drop_p >= 1.f) return bad_arg("attention_bwd_general: dropout p in [0, 1)");
  if (batch <= 0 || q_len <= 0 || kv_len <= 0 || heads <= 0) return bad_arg("attention_bwd_general: extents");
  if (causal && q_len != kv_len) return bad_arg("attention_bwd_general: a causal mask needs q_len == kv_len");
  const int d = heads * AF_HD;
  if ((q_ld % 8) || (k_ld % 8) || (v_ld % 8) || (do_ld % 8) || (o_ld % 8) || (dq_ld % 8) || (dk_ld % 8) || (dv_ld % 8) ||
      q_ld < q_col0 + d || k_ld < k_col0 + d || v_ld < v_col0 + d || do_ld < d || o_ld < d || dq_ld < dq_col0 + d ||
      dk_ld < dk_col0 + d || dv_ld < dv_col0 + d || (q_col0 % 8) || (k_col0 % 8) || (v_col0 % 8) || (dq_col0 % 8) ||
      (dk_col0 % 8) || (dv_col0 % 8))
    return bad_arg("attention_bwd_general: leading dims / column offsets (head_dim must be 64)");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap m[8];
  const uint64_t Lq = (uint64_t)q_len, Tk = (uint64_t)kv_len, B = (uint64_t)batch;
  int rc = make_tmap_3d(&m[0], q, f16, (uint64_t)q_ld, Lq, B, (uint64_t)q_ld, Lq * q_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[1], k, f16, (uint64_t)k_ld, Tk, B, (uint64_t)k_ld, Tk * k_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[2], v, f16, (uint64_t)v_ld, Tk, B, (uint64_t)v_ld, Tk * v_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[3], dout, f16, (uint64_t)d, Lq, B, (uint64_t)do_ld, Lq * do_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[4], out, f16, (uint64_t)d, Lq, B, (uint64_t)o_ld, Lq * o_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[5], dq, f16, (uint64_t)dq_ld, Lq, B, (uint64_t)dq_ld, Lq * dq_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[6], dk, f16, (uint64_t)dk_ld, Tk, B, (uint64_t)dk_ld, Tk * dk_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[7], dv, f16, (uint64_t)dv_ld, Tk, B, (uint64_t)dv_ld, Tk * dv_ld, 128);
  if (rc) return rc;
  AgDev p;
  memset(&p, 0, sizeof(p));
  p.q_len = q_len, p.kv_len = kv_len, p.H = heads, p.causal = causal != 0;
  p.nq = (q_len + AF_BM - 1) / AF_BM;
  p.nc = (kv_len + AF_KC - 1) / AF_KC;
  p.npairs = (p.nq + 1) / 2;
  p.n_items = batch * heads;
  p.q_col0 = q_col0, p.k_col0 = k_col0, p.v_col0 = v_col0, p.dq_col0 = dq_col0, p.dk_col0 = dk_col0, p.dv_col0 = dv_col0;
  p.kv_lens = kv_lens, p.lse = lse, p.scratch = scratch;
  p.scale = 0.125f;
  p.drop_p = drop_p, p.drop_site = site, p.drop_seed = seed;
  p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  p.drop_lp = (q_len + 63) / 64 * 64, p.drop_tp = (kv_len + 63) / 64 * 64;
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (drop_p > 0.f)
    return f16 ? launch_attention_bwd_general<__half, true>(m, p, grid, s)
               : launch_attention_bwd_general<__nv_bfloat16, true>(m, p, grid, s);
  return f16 ? launch_attention_bwd_general<__half, false>(m, p, grid, s)
             : launch_attention_bwd_general<__nv_bfloat16, false>(m, p, grid, s);
}
