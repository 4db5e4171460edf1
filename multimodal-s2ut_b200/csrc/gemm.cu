// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   acc[r, c] = sum_k A[r, k] * W[c, k]          A, W 16-bit K-major; fp32 accumulators in TMEM
//
// Roles (256 threads, one CTA per SM, grid = min(tiles, 148)):
//   warp 0    TMA producer: A (128 x 64) and W (BN x 64) tiles, 128B swizzle, STAGES-deep mbarrier ring
//   warp 1    MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 per k-block;
//             tcgen05.commit frees the smem stage / publishes the accumulator
//   warp 2    TMEM allocator (2 accumulator stages of BN columns -> epilogue overlaps the next tile's MMAs)
//   warps 4-7 epilogue: tcgen05.ld 32 lanes x 32 columns, fused bias / scale / ReLU / GLU / residual /
//             positional add / sigmoid gate, vectorised global stores
// The A operand is addressed through a 3-D tensor map (K, rows, batch) whose row stride may be smaller
// than K: that is how the stride-2 Conv1d layers of the subsampler run as GEMMs over a time-major buffer
// without materialising im2col.  Rows / columns past the tensor bounds are zero-filled by TMA and masked
// in the epilogue.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

struct GemmDev {
  int rows, batches, n, k, kb_split, num_kb, w_batched;
  int m_tiles_per_batch, n_tiles, num_tiles;
  const float* bias;
  float scale;
  int scale_cols;
  void* out0;
  long long out0_ld, out0_bs;
  void* out1;
  long long out1_ld, out1_bs;
  const float* aux0;
  const float* aux1;
  long long aux_ld;
  int rows_per_seq, out_tbc, n_seqs, out_row_offset;
  void* vt;
  int vt_col0, vt_rows;
  long long vt_ld;
  const float* pos;
  const int* seq_lens;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512 (power of two)
  static constexpr int BAR_BYTES = 256;
  static constexpr int BIAS_BYTES = 2 * BN * 4;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + BIAS_BYTES + 1024;  // +align slack
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

template <typename OpT>
__device__ __forceinline__ void store_op(OpT* dst, const float* v, int nvalid) {
  if (nvalid == 32) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q;
      q.x = OpTraits<OpT>::pack2(v[8 * i + 0], v[8 * i + 1]);
      q.y = OpTraits<OpT>::pack2(v[8 * i + 2], v[8 * i + 3]);
      q.z = OpTraits<OpT>::pack2(v[8 * i + 4], v[8 * i + 5]);
      q.w = OpTraits<OpT>::pack2(v[8 * i + 6], v[8 * i + 7]);
      d4[i] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = OpTraits<OpT>::cvt(v[i]);
  }
}
__device__ __forceinline__ void store_f32(float* dst, const float* v, int nvalid) {
  if (nvalid == 32) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}
__device__ __forceinline__ void load_f32(const float* src, float* v, int nvalid) {
  if (nvalid == 32) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 q = s4[i];
      v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < nvalid) ? src[i] : 0.f;
  }
}

template <int BN, int MODE, typename OpT>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
            const __grid_constant__ CUtensorMap mapW, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* full = bars;                 // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES] MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;   // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;          // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapW);
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
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int bi = m_tile / p.m_tiles_per_batch;
        const int row0 = (m_tile % p.m_tiles_per_batch) * Cfg::BM;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          if (kb < p.kb_split)
            tma_load_3d(sA + stage * Cfg::A_BYTES, &mapA0, &full[stage], kb * Cfg::BK, row0, bi);
          else
            tma_load_3d(sA + stage * Cfg::A_BYTES, &mapA1, &full[stage], (kb - p.kb_split) * Cfg::BK, row0, bi);
          tma_load_3d(sB + stage * Cfg::B_BYTES, &mapW, &full[stage], kb * Cfg::BK, n_tile * BN,
                      p.w_batched ? bi : 0);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(Cfg::BM, BN, OpTraits<OpT>::fmt);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      const int k_tail = p.k - (p.num_kb - 1) * Cfg::BK;          // valid K in the last k-block
      const int tail_steps = (k_tail + 15) >> 4;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
          const int steps = (kb == p.num_kb - 1) ? tail_steps : 4;
          for (int kk = 0; kk < steps; ++kk)  // +32 B per K=16 step inside the 128 B swizzle row
            umma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        umma_commit(&tfull[as]);
        if (++as == 2) as = 0, aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;                 // == warp % 4 -> TMEM lane quadrant
    const int et = threadIdx.x - 128;        // 0..127
    const int lrow = ew * 32 + lane;         // accumulator row (TMEM lane) of this thread
    uint32_t as = 0, aphase = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int bi = m_tile / p.m_tiles_per_batch;
      const int r = (m_tile % p.m_tiles_per_batch) * Cfg::BM + lrow;   // row within the batch
      const bool rvalid = r < p.rows;
      int b = bi, t = r;
      if (p.rows_per_seq > 0) b = r / p.rows_per_seq, t = r - b * p.rows_per_seq;
      const long long arow = (p.rows_per_seq > 0) ? r : (long long)bi * p.rows + r;   // aux row
      const int col_tile = n_tile * BN;

      float* sb = sBias + (it & 1) * BN;
      for (int i = et; i < BN; i += 128) {
        const int c = col_tile + i;
        sb[i] = (p.bias != nullptr && c < p.n) ? __ldg(p.bias + c) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(ew * 32) << 16);

      if constexpr (MODE == MM_EPI_GLU_OP || MODE == MM_EPI_GLU_POS_F32) {
        constexpr int HALF = BN / 2;
        const int n_out = p.n >> 1;
        int slen = 0;
        if constexpr (MODE == MM_EPI_GLU_POS_F32) slen = (rvalid && p.seq_lens) ? p.seq_lens[b] : 0x7fffffff;
#pragma unroll 1
        for (int j0 = 0; j0 < HALF; j0 += 32) {
          uint32_t ra[32], rg[32];
          tmem_ld32(taddr + j0, ra);
          tmem_ld32(taddr + HALF + j0, rg);
          tmem_ld_wait();
          const int oc = n_tile * HALF + j0;
          if (rvalid && oc < n_out) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float a = __uint_as_float(ra[i]) + sb[j0 + i];
              const float g = __uint_as_float(rg[i]) + sb[HALF + j0 + i];
              v[i] = a * sigmoidf_(g);
            }
            const int nvalid = min(32, n_out - oc);
            if constexpr (MODE == MM_EPI_GLU_OP) {
              OpT* dst = reinterpret_cast<OpT*>(p.out0) + b * p.out0_bs + (long long)(t + p.out_row_offset) * p.out0_ld + oc;
              store_op<OpT>(dst, v, nvalid);
            } else {
              if (t < slen) {
                float pe[32];
                load_f32(p.pos + (long long)(t + 2) * n_out + oc, pe, nvalid);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], p.scale, pe[i]);
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= p.scale;
              }
              float* dst = reinterpret_cast<float*>(p.out0) + b * p.out0_bs + (long long)t * p.out0_ld + oc;
              store_f32(dst, v, nvalid);
            }
          }
          __syncwarp();
        }
      } else {
        const long long orow = p.out_tbc ? ((long long)t * p.n_seqs + b) : -1;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t ra[32];
          tmem_ld32(taddr + c0, ra);
          tmem_ld_wait();
          const int col = col_tile + c0;
          if (rvalid && col < p.n) {
          const int nvalid = min(32, p.n - col);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]) + sb[c0 + i];

          if constexpr (MODE == MM_EPI_OP) {
            if (col < p.scale_cols) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= p.scale;
            }
            if (p.vt != nullptr && col >= p.vt_col0) {
              // transposed store: lanes hold consecutive t -> coalesced 2-byte stores per column
              OpT* dst = reinterpret_cast<OpT*>(p.vt) + ((long long)b * p.vt_rows + (col - p.vt_col0)) * p.vt_ld + t;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nvalid) dst[(long long)i * p.vt_ld] = OpTraits<OpT>::cvt(v[i]);
            } else {
              OpT* dst = reinterpret_cast<OpT*>(p.out0) + b * p.out0_bs + (long long)t * p.out0_ld + col;
              store_op<OpT>(dst, v, nvalid);
            }
          } else if constexpr (MODE == MM_EPI_RELU_OP) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            OpT* dst = reinterpret_cast<OpT*>(p.out0) + b * p.out0_bs + (long long)t * p.out0_ld + col;
            store_op<OpT>(dst, v, nvalid);
          } else if constexpr (MODE == MM_EPI_RESID_F32) {
            float x[32];
            load_f32(p.aux0 + arow * p.aux_ld + col, x, nvalid);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += x[i];
            float* dst = reinterpret_cast<float*>(p.out0) +
                         (p.out_tbc ? orow * p.out0_ld : b * p.out0_bs + (long long)t * p.out0_ld) + col;
            store_f32(dst, v, nvalid);
          } else if constexpr (MODE == MM_EPI_F32_OP) {
            float* d0 = reinterpret_cast<float*>(p.out0) + b * p.out0_bs + (long long)t * p.out0_ld + col;
            store_f32(d0, v, nvalid);
            OpT* d1 = reinterpret_cast<OpT*>(p.out1) + b * p.out1_bs + (long long)t * p.out1_ld + col;
            store_op<OpT>(d1, v, nvalid);
          } else if constexpr (MODE == MM_EPI_GATE) {
            float x[32], o[32];
            load_f32(p.aux0 + arow * p.aux_ld + col, x, nvalid);
            load_f32(p.aux1 + arow * p.aux_ld + col, o, nvalid);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float g = sigmoidf_(v[i]);
              v[i] = (1.0f - g) * x[i] + g * o[i];
            }
            float* dst = reinterpret_cast<float*>(p.out0) +
                         (p.out_tbc ? orow * p.out0_ld : b * p.out0_bs + (long long)t * p.out0_ld) + col;
            store_f32(dst, v, nvalid);
          } else {  // MM_EPI_F32
            float* dst = reinterpret_cast<float*>(p.out0) + b * p.out0_bs + (long long)t * p.out0_ld + col;
            store_f32(dst, v, nvalid);
          }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) as = 0, aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

int make_tmap_3d(CUtensorMap* out, const void* base, int is_f16, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                 uint64_t stride1, uint64_t stride2, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return bad_arg("cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1 * 2) % 16 || (stride2 * 2) % 16)
    return bad_arg("TMA operand base/strides must be 16-byte aligned");
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};
  if (dim2 == 1 && strides[1] == 0) strides[1] = strides[0] * dim1;
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error),
             "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu strides %llu %llu box %u", (int)r,
             (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2,
             (unsigned long long)strides[0], (unsigned long long)strides[1], box_rows);
    return static_cast<int>(cudaErrorInvalidValue);
  }
  return 0;
}

template <int BN, int MODE, typename OpT>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const GemmDev& p,
                       cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, MODE, OpT>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(gemm)");
    attr_set = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  kern<<<grid, 256, Cfg::SMEM_BYTES, s>>>(a0, a1, w, p);
  MM_CHECK_LAUNCH("gemm_kernel launch");
  return 0;
}

template <int BN, typename OpT>
static int dispatch_mode(int mode, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w,
                         const GemmDev& p, cudaStream_t s) {
  switch (mode) {
    case MM_EPI_OP: return launch_gemm<BN, MM_EPI_OP, OpT>(a0, a1, w, p, s);
    case MM_EPI_RELU_OP: return launch_gemm<BN, MM_EPI_RELU_OP, OpT>(a0, a1, w, p, s);
    case MM_EPI_RESID_F32: return launch_gemm<BN, MM_EPI_RESID_F32, OpT>(a0, a1, w, p, s);
    case MM_EPI_GLU_OP: return launch_gemm<BN, MM_EPI_GLU_OP, OpT>(a0, a1, w, p, s);
    case MM_EPI_GLU_POS_F32: return launch_gemm<BN, MM_EPI_GLU_POS_F32, OpT>(a0, a1, w, p, s);
    case MM_EPI_F32_OP: return launch_gemm<BN, MM_EPI_F32_OP, OpT>(a0, a1, w, p, s);
    case MM_EPI_GATE: return launch_gemm<BN, MM_EPI_GATE, OpT>(a0, a1, w, p, s);
    case MM_EPI_F32: return launch_gemm<BN, MM_EPI_F32, OpT>(a0, a1, w, p, s);
  }
  return bad_arg("unknown epilogue mode");
}

}  // namespace mm

extern "C" int mm_gemm(const mm_gemm_args* a, void* stream) {
  using namespace mm;
  if (!a || !a->a0 || !a->w || !a->out0) return bad_arg("null operand");
  if (a->rows <= 0 || a->batches <= 0 || a->n <= 0 || a->k <= 0) return bad_arg("non-positive extent");
  if (a->dtype != MM_DTYPE_BF16 && a->dtype != MM_DTYPE_F16) return bad_arg("dtype");
  const int k0 = a->a1 ? a->k_split : a->k;
  if (a->a1 && (k0 <= 0 || k0 % 64 != 0 || k0 >= a->k)) return bad_arg("k_split must be a multiple of 64 inside (0,k)");
  if (a->rows_per_seq > 0 && a->batches != 1) return bad_arg("rows_per_seq requires batches == 1");
  int bn = a->block_n ? a->block_n : 256;
  if (bn != 128 && bn != 256) return bad_arg("block_n must be 128 or 256");
  const bool glu = a->mode == MM_EPI_GLU_OP || a->mode == MM_EPI_GLU_POS_F32;
  if (glu && (a->n % bn != 0)) return bad_arg("GLU epilogue needs n % block_n == 0");
  if ((a->out0_ld % 8) || (a->out1 && a->out1_ld % 8) || (a->aux0 && a->aux_ld % 4))
    return bad_arg("output leading dims must be multiples of 8 elements");
  if (a->mode == MM_EPI_RESID_F32 && !a->aux0) return bad_arg("RESID needs aux0");
  if (a->mode == MM_EPI_GATE && (!a->aux0 || !a->aux1)) return bad_arg("GATE needs aux0 and aux1");
  if (a->mode == MM_EPI_F32_OP && !a->out1) return bad_arg("F32_OP needs out1");
  if (a->mode == MM_EPI_GLU_POS_F32 && !a->pos) return bad_arg("GLU_POS needs pos");
  if (a->out_tbc && a->n_seqs <= 0) return bad_arg("out_tbc needs n_seqs");

  CUtensorMap mA0, mA1, mW;
  const int f16 = a->dtype == MM_DTYPE_F16;
  int rc = make_tmap_3d(&mA0, a->a0, f16, (uint64_t)k0, (uint64_t)a->rows, (uint64_t)a->batches, (uint64_t)a->a0_ld,
                        (uint64_t)a->a0_bs, 128);
  if (rc) return rc;
  if (a->a1) {
    rc = make_tmap_3d(&mA1, a->a1, f16, (uint64_t)(a->k - k0), (uint64_t)a->rows, (uint64_t)a->batches,
                      (uint64_t)a->a1_ld, (uint64_t)a->a1_bs, 128);
    if (rc) return rc;
  } else {
    mA1 = mA0;
  }
  rc = make_tmap_3d(&mW, a->w, f16, (uint64_t)a->k, (uint64_t)a->n, (uint64_t)(a->w_batched ? a->batches : 1),
                    (uint64_t)a->w_ld, (uint64_t)a->w_bs, (uint32_t)bn);
  if (rc) return rc;

  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.rows = a->rows, p.batches = a->batches, p.n = a->n, p.k = a->k;
  p.num_kb = (a->k + 63) / 64;
  p.kb_split = a->a1 ? k0 / 64 : p.num_kb;
  p.w_batched = a->w_batched;
  p.m_tiles_per_batch = (a->rows + 127) / 128;
  p.n_tiles = (a->n + bn - 1) / bn;
  p.num_tiles = p.m_tiles_per_batch * a->batches * p.n_tiles;
  p.bias = a->bias, p.scale = a->scale, p.scale_cols = a->scale_cols;
  p.out0 = a->out0, p.out0_ld = a->out0_ld, p.out0_bs = a->out0_bs;
  p.out1 = a->out1, p.out1_ld = a->out1_ld, p.out1_bs = a->out1_bs;
  p.aux0 = a->aux0, p.aux1 = a->aux1, p.aux_ld = a->aux_ld;
  p.rows_per_seq = a->rows_per_seq, p.out_tbc = a->out_tbc, p.n_seqs = a->n_seqs;
  p.out_row_offset = a->out_row_offset;
  p.vt = a->vt, p.vt_col0 = a->vt_col0, p.vt_rows = a->vt_rows, p.vt_ld = a->vt_ld;
  p.pos = a->pos, p.seq_lens = a->seq_lens;

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bn == 256) {
    return f16 ? dispatch_mode<256, __half>(a->mode, mA0, mA1, mW, p, s)
               : dispatch_mode<256, __nv_bfloat16>(a->mode, mA0, mA1, mW, p, s);
  }
  return f16 ? dispatch_mode<128, __half>(a->mode, mA0, mA1, mW, p, s)
             : dispatch_mode<128, __nv_bfloat16>(a->mode, mA0, mA1, mW, p, s);
}
