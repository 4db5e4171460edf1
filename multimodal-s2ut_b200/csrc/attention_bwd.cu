// Attention backward, score side, on tcgen05: from q, k, v, dO, O and the forward pass's per-row log-sum-exp this
// kernel forms, per (sequence, head),
//     S = q k^T            P = exp(S - lse)   (masked keys: 0)
//     dP = dO v^T          dS = P o (dP - delta),   delta = rowsum(dO o O)
// and writes P and dS as 16-bit [batch * heads][Lp][Tp] matrices -- the operands of the three remaining contractions
// (dV = P^T dO, dK = dS^T q, dQ = dS k, on the GEMM kernel).  The fp32 scores, dP and the softmax backward pass never
// touch HBM: they replaced a K = 64 scores GEMM writing fp32 (bound by its stores, 5 % tensor pipe), a second one for
// dP and a row-wise softmax-backward kernel (fairseq MultiheadAttention under autograd).
//
// The flash-attention identities make the kernel chunk-local: with lse and delta per query row there is no row
// reduction over the keys, so any key length, key-length masks and the causal mask are per-element predicates.
//
// Roles (320 threads, persistent over (sequence, head, 128-query tile) items):
//   warp 8     TMA producer: q, dO, O tiles once per item; (k_c | v_c) 128-key chunks through a 2-stage ring
//   warp 9     MMA issuer:   S_c = q k_c^T and dP_c = dO v_c^T (SS, M128 N128, K = 64) into TMEM stage c & 1
//              (2 x 256 columns: the next chunk's products run while the consumers work on this one)
//   warps 0-7  consumers, two threads per query row (64 keys each): delta from the dO / O tiles, then per chunk
//              tcgen05.ld -> exp2 / multiply -> 16-bit -> swizzled staging slabs -> TMA store of P_c and dS_c.
// Chunks that lie entirely behind an utterance's key length or above the causal diagonal skip the loads and the MMAs;
// their P / dS tiles are still written (as zeros): the GEMMs that follow contract over every key.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int AB_BM = 128, AB_HD = 64, AB_KC = 128;
constexpr int AB_TILE = AB_BM * AB_HD * 2;                 // 16 KB: every TMA box is {64 elements, 128 rows}
constexpr int AB_SMEM_BYTES = 3 * AB_TILE /* q, dO, O */ + 2 * 2 * AB_TILE /* 2 x (k_c | v_c) */ +
                              4 * AB_TILE /* P_c, dS_c staging: 2 slabs of 64 keys each */ + 256 + 2 * AB_BM * 4 + 1024;
constexpr int AB_THREADS = 320;
static_assert(AB_SMEM_BYTES <= 232448, "shared memory budget");

struct AbDev {
  int q_len, kv_len, H, nqt, nc, n_items, causal;
  int q_col0, k_col0, v_col0;
  const int* kv_lens;        // [batch] valid keys, or NULL
  const float* lse;          // [batch][H][q_len]
};

__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename OpT>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_scores_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                            const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO,
                            const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapP,
                            const __grid_constant__ CUtensorMap mapDS, const AbDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + AB_TILE;
  uint8_t* sO = sDO + AB_TILE;
  uint8_t* sKV = sO + AB_TILE;                  // [2 stages] (k_c 16 KB | v_c 16 KB)
  uint8_t* sStage = sKV + 4 * AB_TILE;          // P slab 0, P slab 1, dS slab 0, dS slab 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 4 * AB_TILE);
  uint64_t* q_full = bars;            // [1] TMA (q, dO, O) -> MMA + consumers
  uint64_t* q_empty = bars + 1;       // [1] last MMAs of the item done + consumers have read dO / O -> TMA (count 1 + 8)
  uint64_t* kv_full = bars + 2;       // [2] TMA (k_c, v_c) -> MMA
  uint64_t* kv_empty = bars + 4;      // [2] MMAs of the chunk done -> TMA
  uint64_t* t_full = bars + 6;        // [2] S_c, dP_c in TMEM stage -> consumers
  uint64_t* t_empty = bars + 8;       // [2] 8 consumer warps have read the stage -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* x_delta = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 halves][128 rows]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_local = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr float L2E = 1.4426950408889634f;

  if (tid == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    tma_prefetch_desc(&mapDO);
    tma_prefetch_desc(&mapO);
    tma_prefetch_desc(&mapP);
    tma_prefetch_desc(&mapDS);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 9);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
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

  auto item_coords = [&](int i, int& qt, int& h, int& b) {
    const int item = blockIdx.x + i * gridDim.x;
    qt = item % p.nqt, h = (item / p.nqt) % p.H, b = item / (p.nqt * p.H);
  };
  auto len_of = [&](int b) { return p.kv_lens ? max(0, min(p.kv_lens[b], p.kv_len)) : p.kv_len; };
  // chunk c of an item carries at least one visible (query, key) pair
  auto live = [&](int b, int qt, int c) {
    if (c * AB_KC >= len_of(b)) return false;
    if (p.causal && c * AB_KC > min(qt * AB_BM + AB_BM, p.q_len) - 1) return false;
    return true;
  };

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t kv_n = 0;
      for (int i = 0; i < n_local; ++i) {
        int qt, h, b;
        item_coords(i, qt, h, b);
        mbar_wait(q_empty, (i & 1) ^ 1);
        mbar_expect_tx(q_full, 3 * AB_TILE);
        tma_load_3d(sQ, &mapQ, q_full, p.q_col0 + h * AB_HD, qt * AB_BM, b);
        tma_load_3d(sDO, &mapDO, q_full, h * AB_HD, qt * AB_BM, b);
        tma_load_3d(sO, &mapO, q_full, h * AB_HD, qt * AB_BM, b);
        for (int c = 0; c < p.nc; ++c) {
          if (!live(b, qt, c)) continue;
          const uint32_t st = kv_n & 1;
          mbar_wait(&kv_empty[st], ((kv_n >> 1) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * AB_TILE);
          tma_load_3d(sKV + st * 2 * AB_TILE, &mapK, &kv_full[st], p.k_col0 + h * AB_HD, c * AB_KC, b);
          tma_load_3d(sKV + st * 2 * AB_TILE + AB_TILE, &mapV, &kv_full[st], p.v_col0 + h * AB_HD, c * AB_KC, b);
          ++kv_n;
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(AB_BM, AB_KC, OpTraits<OpT>::fmt);
      uint32_t kv_n = 0;
      for (int i = 0; i < n_local; ++i) {
        int qt, h, b;
        item_coords(i, qt, h, b);
        mbar_wait(q_full, i & 1);
        tc_fence_after();
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ)), dodesc = umma_desc_sw128(smem_u32(sDO));
        for (int c = 0; c < p.nc; ++c) {
          if (!live(b, qt, c)) continue;
          const uint32_t st = kv_n & 1, ph = (kv_n >> 1) & 1;
          mbar_wait(&t_empty[st], ph ^ 1);          // the consumers have drained this TMEM stage
          mbar_wait(&kv_full[st], ph);
          tc_fence_after();
          const uint64_t kdesc = umma_desc_sw128(smem_u32(sKV + st * 2 * AB_TILE));
          const uint64_t vdesc = umma_desc_sw128(smem_u32(sKV + st * 2 * AB_TILE + AB_TILE));
          const uint32_t t_s = tmem_base + 256 * st, t_dp = t_s + 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(t_s, qdesc + 2 * kk, kdesc + 2 * kk, idesc, kk != 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16(t_dp, dodesc + 2 * kk, vdesc + 2 * kk, idesc, kk != 0);
          umma_commit(&t_full[st]);
          umma_commit(&kv_empty[st]);
          ++kv_n;
        }
        umma_commit(q_empty);     // every MMA that reads the q / dO tiles of this item has completed
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int hf = warp >> 2;                  // which 64 keys of a chunk / which 32 head columns for delta
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>((warp & 3) * 32) << 16;
    auto group_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const bool elected = tid == 0;
    uint32_t kv_n = 0;
    for (int i = 0; i < n_local; ++i) {
      int qt, h, b;
      item_coords(i, qt, h, b);
      const int qrow = qt * AB_BM + row;
      const int len = len_of(b);
      // ---- delta = rowsum(dO o O) over the head's 64 columns (this thread: 32 of them), lse of the row ----
      mbar_wait(q_full, i & 1);
      float dsum = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int chunk = ((4 * hf + k) ^ (row & 7)) << 4;
        const uint4 a = *reinterpret_cast<const uint4*>(sDO + row * 128 + chunk);
        const uint4 o = *reinterpret_cast<const uint4*>(sO + row * 128 + chunk);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const OpT* ap = reinterpret_cast<const OpT*>(&aw[j]);
          const OpT* op = reinterpret_cast<const OpT*>(&ow[j]);
          dsum = fmaf(OpTraits<OpT>::to_float(ap[0]), OpTraits<OpT>::to_float(op[0]), dsum);
          dsum = fmaf(OpTraits<OpT>::to_float(ap[1]), OpTraits<OpT>::to_float(op[1]), dsum);
        }
      }
      x_delta[hf * AB_BM + row] = dsum;
      group_sync();
      const float delta = dsum + x_delta[(hf ^ 1) * AB_BM + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(q_empty);      // this warp is done with the dO / O tiles (q_empty also waits for the MMAs)
      const float lb = (qrow < p.q_len ? __ldg(p.lse + ((long long)b * p.H + h) * p.q_len + qrow) : 0.f) * L2E;
      group_sync();                             // x_delta may be rewritten by the next item only after everyone read it

      for (int c = 0; c < p.nc; ++c) {
        const bool lv = live(b, qt, c);
        // key k of my 64 is visible iff k < len and (not causal or k <= my query row)
        const int k0 = c * AB_KC + 64 * hf;
        int nv = min(64, max(0, len - k0));
        if (p.causal) nv = min(nv, max(0, qrow + 1 - k0));
        uint8_t* sP = sStage + hf * AB_TILE;              // slab hf of P_c (my 64 keys)
        uint8_t* sD = sStage + 2 * AB_TILE + hf * AB_TILE;
        uint32_t st = 0;
        if (lv) {
          st = kv_n & 1;
          mbar_wait(&t_full[st], (kv_n >> 1) & 1);
          tc_fence_after();
        }
        if (elected) bulk_wait_read<0>();       // the previous chunk's stores have finished reading the slabs
        group_sync();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t pk[16], dk[16];
          if (lv) {
            uint32_t rs[32], rd[32];
            const uint32_t t_s = tmem_base + 256 * st + t_lane + 64 * hf + 32 * half;
            tmem_ld32(t_s, rs);
            tmem_ld32(t_s + 128, rd);
            tmem_ld_wait();
            const int nvh = nv - 32 * half;     // visible keys among these 32
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              float p0 = ab_ex2(fmaf(__uint_as_float(rs[k]), L2E, -lb));
              float p1 = ab_ex2(fmaf(__uint_as_float(rs[k + 1]), L2E, -lb));
              p0 = k < nvh ? p0 : 0.f;
              p1 = k + 1 < nvh ? p1 : 0.f;
              const float d0 = p0 * (__uint_as_float(rd[k]) - delta);
              const float d1 = p1 * (__uint_as_float(rd[k + 1]) - delta);
              pk[k >> 1] = OpTraits<OpT>::pack2(p0, p1);
              dk[k >> 1] = OpTraits<OpT>::pack2(d0, d1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) pk[k] = 0u, dk[k] = 0u;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int chunk = ((4 * half + k) ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(sP + row * 128 + chunk) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
            *reinterpret_cast<uint4*>(sD + row * 128 + chunk) = make_uint4(dk[4 * k], dk[4 * k + 1], dk[4 * k + 2], dk[4 * k + 3]);
          }
        }
        if (lv) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[st]);      // S_c / dP_c are consumed: the stage may be overwritten
          ++kv_n;
        }
        fence_proxy_async_smem();
        group_sync();
        if (elected) {
          const int bh = b * p.H + h;
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            tma_store_3d(&mapP, sStage + s2 * AB_TILE, c * AB_KC + 64 * s2, qt * AB_BM, bh);
            tma_store_3d(&mapDS, sStage + (2 + s2) * AB_TILE, c * AB_KC + 64 * s2, qt * AB_BM, bh);
          }
          bulk_commit();
        }
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

template <typename OpT>
static int launch_attention_bwd_scores(const CUtensorMap (&m)[7], const AbDev& p, cudaStream_t s) {
  auto kern = attention_bwd_scores_kernel<OpT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(attention_bwd_scores)");
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.n_items < kNumSMs ? p.n_items : kNumSMs);
  cfg.blockDim = dim3(AB_THREADS);
  cfg.dynamicSmemBytes = AB_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, m[0], m[1], m[2], m[3], m[4], m[5], m[6], p);
  if (e != cudaSuccess) return fail(e, "attention_bwd_scores_kernel launch");
  return 0;
}

}  // namespace mm

using namespace mm;

extern "C" int mm_attention_bwd_scores(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k,
                                       int64_t k_ld, int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0,
                                       int32_t kv_len, const int32_t* kv_lens, int32_t batch, int32_t heads,
                                       int32_t causal, const void* dout, int64_t do_ld, const void* out, int64_t o_ld,
                                       const float* lse, void* probs, void* dscores, int64_t pd_ld, int64_t pd_bs,
                                       int32_t dtype, void* stream) {
  if (!q || !k || !v || !dout || !out || !lse || !probs || !dscores) return bad_arg("attention_bwd_scores: null pointer");
  if (batch <= 0 || q_len <= 0 || kv_len <= 0 || heads <= 0) return bad_arg("attention_bwd_scores: extents");
  const int d = heads * AB_HD;
  if ((q_ld % 8) || (k_ld % 8) || (v_ld % 8) || (do_ld % 8) || (o_ld % 8) || (pd_ld % 8) || (pd_bs % 8) ||
      q_ld < q_col0 + d || k_ld < k_col0 + d || v_ld < v_col0 + d || do_ld < d || o_ld < d || pd_ld < kv_len ||
      pd_bs < (int64_t)q_len * pd_ld || (q_col0 % 8) || (k_col0 % 8) || (v_col0 % 8))
    return bad_arg("attention_bwd_scores: leading dims / column offsets (head_dim must be 64)");
  if (causal && q_len != kv_len) return bad_arg("attention_bwd_scores: a causal mask needs q_len == kv_len");
  const int f16 = dtype == MM_DTYPE_F16;
  CUtensorMap m[7];
  int rc = make_tmap_3d(&m[0], q, f16, (uint64_t)q_ld, (uint64_t)q_len, (uint64_t)batch, (uint64_t)q_ld,
                        (uint64_t)q_len * q_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[1], k, f16, (uint64_t)k_ld, (uint64_t)kv_len, (uint64_t)batch, (uint64_t)k_ld,
                    (uint64_t)kv_len * k_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[2], v, f16, (uint64_t)v_ld, (uint64_t)kv_len, (uint64_t)batch, (uint64_t)v_ld,
                    (uint64_t)kv_len * v_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[3], dout, f16, (uint64_t)d, (uint64_t)q_len, (uint64_t)batch, (uint64_t)do_ld,
                    (uint64_t)q_len * do_ld, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[4], out, f16, (uint64_t)d, (uint64_t)q_len, (uint64_t)batch, (uint64_t)o_ld,
                    (uint64_t)q_len * o_ld, 128);
  if (rc) return rc;
  // P / dS: [batch * heads][q_len][kv_len] windows of the [..][Lp][Tp] buffers; stores are clipped at these bounds
  rc = make_tmap_3d(&m[5], probs, f16, (uint64_t)kv_len, (uint64_t)q_len, (uint64_t)batch * heads, (uint64_t)pd_ld,
                    (uint64_t)pd_bs, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&m[6], dscores, f16, (uint64_t)kv_len, (uint64_t)q_len, (uint64_t)batch * heads, (uint64_t)pd_ld,
                    (uint64_t)pd_bs, 128);
  if (rc) return rc;
  AbDev p;
  memset(&p, 0, sizeof(p));
  p.q_len = q_len, p.kv_len = kv_len, p.H = heads, p.causal = causal != 0;
  p.nqt = (q_len + AB_BM - 1) / AB_BM;
  p.nc = (kv_len + AB_KC - 1) / AB_KC;
  p.n_items = batch * heads * p.nqt;
  p.q_col0 = q_col0, p.k_col0 = k_col0, p.v_col0 = v_col0;
  p.kv_lens = kv_lens, p.lse = lse;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return f16 ? launch_attention_bwd_scores<__half>(m, p, s) : launch_attention_bwd_scores<__nv_bfloat16>(m, p, s);
}
