// Backward-pass and optimizer kernels of the training-step variant (BASELINE configs[2]): the HBM-bound pieces that sit
// between the tcgen05 GEMMs (gemm.cu serves every dgrad / wgrad contraction):
//   pack_t           16-bit operand copies of a gradient / activation, straight and TRANSPOSED (the wgrad GEMMs contract
//                    over tokens, so both operands are needed token-contiguous), optional ReLU mask, head split / merge
//   rowsum           bias gradients = row sums of the transposed gradient
//   reduce_partials  deterministic split-K / per-block partial reduction into the fp32 gradient
//   layernorm_bwd    dx (+ residual gradient), per-block dgamma / dbeta partials
//   softmax_bwd      P and dS = P o (dP - rowsum(P o dP)) from recomputed scores
//   glu_bwd, gate_bwd, col2im_k5s2   element-wise backward of GLU, the selective gate and the stride-2 k=5 conv gather
//   adam, sumsq, clip_coef           fairseq Adam on fp32 master parameters with global-norm clipping
// All plain CUDA-core kernels: warp-per-row or tile-per-block, fp32 arithmetic, deterministic reductions.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------
// pack_t: tile 64 rows x 64 cols through shared memory, 256 threads, 16-byte global loads / stores where aligned.
// ---------------------------------------------------------------------------------------------------
struct PackArgs {
  const void* in;
  const void* mask;
  void* out_n;
  void* out_t;
  long long in_ld, in_bs0, in_bs1, mask_ld;
  long long n_ld, n_bs0, n_bs1;
  long long t_ld, t_bs0, t_bs1;
  int rows, cols, nb1, t_cols_pad, in_is_f32;
  float scale;
};

template <typename OpT>
__global__ void __launch_bounds__(256) pack_t_kernel(PackArgs a) {
  // 64 x 64 tile of 16-bit values, row stride 66 (132 B): the column gathers of the transposed store are 2-way
  // bank-conflicted at worst; rows are written as 32-bit words.
  __shared__ __align__(16) uint32_t tile[64 * 33];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int b0 = blockIdx.z / a.nb1, b1 = blockIdx.z % a.nb1;
  const long long in_off = (long long)b0 * a.in_bs0 + (long long)b1 * a.in_bs1;
  const long long n_off = (long long)b0 * a.n_bs0 + (long long)b1 * a.n_bs1;
  const bool in_vec = (((a.in_ld | a.in_bs0 | a.in_bs1) & 7) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
  const bool mask_vec = !a.mask || (((a.mask_ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.mask) & 15) == 0));
  const bool n_vec = a.out_n && (((a.n_ld | a.n_bs0 | a.n_bs1) & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.out_n) & 15) == 0);
  // ---- load: 8 consecutive columns per thread, 2 rows per thread
  const int ch = threadIdx.x & 7;
  const int c = c0 + ch * 8;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int rl = (threadIdx.x >> 3) + 32 * h, r = r0 + rl;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (r < a.rows && c < a.cols) {
      const long long idx = in_off + (long long)r * a.in_ld + c;
      const bool full = c + 8 <= a.cols;
      if (full && in_vec) {
        if (a.in_is_f32) {
          const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.in) + idx);
          const float4 x0 = __ldcs(p), x1 = __ldcs(p + 1);
          v[0] = x0.x, v[1] = x0.y, v[2] = x0.z, v[3] = x0.w, v[4] = x1.x, v[5] = x1.y, v[6] = x1.z, v[7] = x1.w;
        } else {
          const uint4 q = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const OpT*>(a.in) + idx));
          const OpT* e = reinterpret_cast<const OpT*>(&q);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = OpTraits<OpT>::to_float(e[i]);
        }
      } else {
        for (int i = 0; i < 8 && c + i < a.cols; ++i)
          v[i] = a.in_is_f32 ? reinterpret_cast<const float*>(a.in)[idx + i]
                             : OpTraits<OpT>::to_float(reinterpret_cast<const OpT*>(a.in)[idx + i]);
      }
      if (a.scale != 1.0f) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= a.scale;
      }
      if (a.mask) {
        const OpT* m = reinterpret_cast<const OpT*>(a.mask) + (long long)r * a.mask_ld + c;
        if (full && mask_vec) {
          const uint4 q = __ldcs(reinterpret_cast<const uint4*>(m));
          const OpT* e = reinterpret_cast<const OpT*>(&q);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (!(OpTraits<OpT>::to_float(e[i]) > 0.f)) v[i] = 0.f;
        } else {
          for (int i = 0; i < 8 && c + i < a.cols; ++i)
            if (!(OpTraits<OpT>::to_float(m[i]) > 0.f)) v[i] = 0.f;
        }
      }
    }
    uint4 pk;
    pk.x = OpTraits<OpT>::pack2(v[0], v[1]);
    pk.y = OpTraits<OpT>::pack2(v[2], v[3]);
    pk.z = OpTraits<OpT>::pack2(v[4], v[5]);
    pk.w = OpTraits<OpT>::pack2(v[6], v[7]);
    if (a.out_n && r < a.rows && c < a.cols) {
      OpT* o = reinterpret_cast<OpT*>(a.out_n) + n_off + (long long)r * a.n_ld + c;
      if (c + 8 <= a.cols && n_vec) {
        *reinterpret_cast<uint4*>(o) = pk;
      } else {
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
        for (int i = 0; i < 8 && c + i < a.cols; i += 2) *reinterpret_cast<uint32_t*>(o + i) = w[i >> 1];
      }
    }
    uint32_t* t = tile + rl * 33 + ch * 4;
    t[0] = pk.x, t[1] = pk.y, t[2] = pk.z, t[3] = pk.w;
  }
  if (!a.out_t) return;
  __syncthreads();
  // ---- transposed store: output row = input column; a thread gathers 8 consecutive input rows of one column
  OpT* ot = reinterpret_cast<OpT*>(a.out_t) + (long long)b0 * a.t_bs0 + (long long)b1 * a.t_bs1;
  const bool t_vec = (((a.t_ld | a.t_bs0 | a.t_bs1) & 7) == 0) && ((a.t_cols_pad & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.out_t) & 15) == 0);
  const uint16_t* t16 = reinterpret_cast<const uint16_t*>(tile);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane & 7;
  const int r = r0 + 8 * j;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int cl = warp * 8 + pass * 4 + (lane >> 3), cc = c0 + cl;
    if (cc >= a.cols || r >= a.t_cols_pad) continue;
    uint16_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = t16[(8 * j + i) * 66 + cl];
    OpT* dst = ot + (long long)cc * a.t_ld + r;
    if (t_vec) {
      uint4 pk;
      pk.x = e[0] | ((uint32_t)e[1] << 16), pk.y = e[2] | ((uint32_t)e[3] << 16);
      pk.z = e[4] | ((uint32_t)e[5] << 16), pk.w = e[6] | ((uint32_t)e[7] << 16);
      *reinterpret_cast<uint4*>(dst) = pk;
    } else {
      for (int i = 0; i < 8 && r + i < a.t_cols_pad; i += 2)
        *reinterpret_cast<uint32_t*>(dst + i) = e[i] | ((uint32_t)e[i + 1] << 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// rowsum: out[r] (+)= sum_c in[r, c]; one 256-thread block per row, 16-byte loads, fixed-order block reduction.
// ---------------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256) rowsum_kernel(const OpT* __restrict__ in, long long ld, int rows, int cols,
                                                      float* __restrict__ out, int accumulate) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const OpT* p = in + (long long)row * ld;
  float s = 0.f;
  const bool vec = ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const int cols8 = vec ? (cols & ~7) : 0;
  for (int c = threadIdx.x * 8; c < cols8; c += 256 * 8) {
    const uint4 q = *reinterpret_cast<const uint4*>(p + c);
    const OpT* e = reinterpret_cast<const OpT*>(&q);
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += OpTraits<OpT>::to_float(e[i]);
    s += t;
  }
  for (int c = cols8 + threadIdx.x; c < cols; c += 256) s += OpTraits<OpT>::to_float(p[c]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    out[row] = accumulate ? out[row] + t : t;
  }
}

// out[i] (+)= sum_s part[s * stride + i]
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int S, long long stride,
                                                               long long n, float* __restrict__ out, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += part[(long long)k * stride + i];
  out[i] = accumulate ? out[i] + s : s;
}

// several reductions in one launch (blockIdx.y = job): the backward pass defers the ~10 partial reductions of a layer
// (split-K wgrad partials, bias column sums, LayerNorm parameter partials) and runs them together.  A block covers
// 256 / slices float4 columns and cuts the S partials of a job into `slices` interleaved slices (slice s sums partials
// s, s + slices, ... with four 128-bit loads in flight), so a job with few outputs and many partials (a bias gradient:
// 512 outputs x 125 partials) is as parallel as one with many outputs and few partials (a weight gradient).  The slices
// are combined through shared memory in a fixed order: the result depends only on (S, slices), never on the launch.
struct ReduceJobs {
  const float* part[16];
  float* out[16];
  long long stride[16], n[16];
  int S[16], accumulate[16], slices[16];
};
__global__ void __launch_bounds__(256) reduce_many_kernel(ReduceJobs t) {
  __shared__ float4 red[2][256];
  const int j = blockIdx.y;
  const float* __restrict__ part = t.part[j];
  float* __restrict__ out = t.out[j];
  const long long n = t.n[j], stride = t.stride[j];
  const int S = t.S[j], acc = t.accumulate[j], sl = t.slices[j];
  const int cols4 = 256 / sl;                       // float4 columns per half tile
  const int c4 = threadIdx.x % cols4, slice = threadIdx.x / cols4;
  const bool vec = ((stride & 3) == 0) && ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(part) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long n4 = (n + 3) >> 2;                // float4 columns (the last one may be partial when !vec)
  const long long tiles = (n4 + 2 * cols4 - 1) / (2 * cols4);     // a tile = two half tiles: 8 loads per thread in flight
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    long long i4[2];
    float4 s[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) i4[u] = (2 * tile + u) * cols4 + c4, s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec) {
      const long long s4 = stride >> 2;
      const float4* p4 = reinterpret_cast<const float4*>(part);
      int k = slice;
      for (; k + 3 * sl < S; k += 4 * sl) {
        float4 v[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            v[u][q] = i4[u] < n4 ? __ldcs(p4 + (long long)(k + q * sl) * s4 + i4[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int q = 0; q < 4; ++q) s[u].x += v[u][q].x, s[u].y += v[u][q].y, s[u].z += v[u][q].z, s[u].w += v[u][q].w;
      }
      for (; k < S; k += sl) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (i4[u] < n4) {
            const float4 a = __ldcs(p4 + (long long)k * s4 + i4[u]);
            s[u].x += a.x, s[u].y += a.y, s[u].z += a.z, s[u].w += a.w;
          }
      }
    } else {
      for (int u = 0; u < 2; ++u) {
        float* sv = reinterpret_cast<float*>(&s[u]);
        for (int e = 0; e < 4; ++e) {
          const long long i = 4 * i4[u] + e;
          if (i >= n) break;
          for (int k = slice; k < S; k += sl) sv[e] += part[(long long)k * stride + i];
        }
      }
    }
    if (sl > 1) {
      __syncthreads();                              // the previous tile's combine has read `red`
      red[0][threadIdx.x] = s[0], red[1][threadIdx.x] = s[1];
      __syncthreads();
    }
    if (slice == 0) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (i4[u] >= n4) continue;
        float4 r = s[u];
        for (int q = 1; q < sl; ++q) {
          const float4 o = red[u][q * cols4 + c4];
          r.x += o.x, r.y += o.y, r.z += o.z, r.w += o.w;
        }
        if (vec) {
          float4* o4 = reinterpret_cast<float4*>(out) + i4[u];
          if (acc) {
            const float4 o = *o4;
            r.x += o.x, r.y += o.y, r.z += o.z, r.w += o.w;
          }
          *o4 = r;
        } else {
          const float* sv = reinterpret_cast<const float*>(&r);
          for (int e = 0; e < 4; ++e) {
            const long long i = 4 * i4[u] + e;
            if (i < n) out[i] = acc ? out[i] + sv[e] : sv[e];
          }
        }
      }
    }
  }
}

// column sums of a 16-bit [rows, cols] matrix (bias gradients): a block owns 256 columns x COLSUM_ROWS rows.  A lane
// owns 8 consecutive columns (one 128-bit word per row) and the 8 warps take the block's rows round robin.  The words
// travel global -> shared memory by cp.async, 16 per lane issued back to back (ptxas interleaved plain 128-bit loads
// with their consumption and kept only ~5 in flight: 2 TB/s instead of 5), and each lane then sums the words it
// fetched itself.  Partials [gridDim.y][cols] are summed by reduce_partials.  Rows with (r % period) >= valid are
// skipped when period > 0.
constexpr int COLSUM_ROWS = 128;
constexpr int COLSUM_SMEM = COLSUM_ROWS * 32 * 16;     // 64 KB
template <typename OpT, bool PERIODIC>
__global__ void __launch_bounds__(256) colsum_kernel(const OpT* __restrict__ in, long long ld, int rows, int cols,
                                                      int period, int valid, float* __restrict__ partials) {
  extern __shared__ __align__(16) uint4 cs_tile[];     // [COLSUM_ROWS][32 lanes]
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + 8 * lane;
  const int r_begin = blockIdx.y * COLSUM_ROWS;
  const int r_end = min(rows, r_begin + COLSUM_ROWS);
  const bool vec = ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (c + 8 <= cols);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (vec) {
    const OpT* p0 = in + (long long)(r_begin + warp) * ld + c;
    const uint32_t d0 = smem_u32(cs_tile + warp * 32 + lane);
#pragma unroll
    for (int j = 0; j < COLSUM_ROWS / 8; ++j)
      if (r_begin + warp + 8 * j < r_end)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + j * 8 * 32 * 16), "l"(p0 + (long long)j * 8 * ld)
                     : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int j = 0; j < COLSUM_ROWS / 8; ++j) {
      const int r = r_begin + warp + 8 * j;
      bool ok = r < r_end;
      if (PERIODIC) ok = ok && (r % period) < valid;
      if (ok) {
        const uint4 q = cs_tile[(warp + 8 * j) * 32 + lane];
        const OpT* e = reinterpret_cast<const OpT*>(&q);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += OpTraits<OpT>::to_float(e[i]);
      }
    }
  } else if (c < cols) {
    for (int r = r_begin + warp; r < r_end; r += 8) {
      if (PERIODIC && (r % period) >= valid) continue;
      for (int i = 0; i < 8 && c + i < cols; ++i) s[i] += OpTraits<OpT>::to_float(in[(long long)r * ld + c + i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][8 * lane + i] = s[i];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partials[(long long)blockIdx.y * cols + cc] = t;
  }
}

template <typename OpT, bool PERIODIC>
static int launch_colsum(const void* in, long long ld, int rows, int cols, int period, int valid, float* partials,
                         cudaStream_t s) {
  auto kern = colsum_kernel<OpT, PERIODIC>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, COLSUM_SMEM);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(colsum)");
    attr_set = true;
  }
  dim3 grid((cols + 255) / 256, (rows + COLSUM_ROWS - 1) / COLSUM_ROWS);
  if (grid.y > 65535) return bad_arg("colsum: too many rows");
  kern<<<grid, 256, COLSUM_SMEM, s>>>(reinterpret_cast<const OpT*>(in), ld, rows, cols, period, valid, partials);
  MM_CHECK_LAUNCH("colsum_kernel launch");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward: warp per row (grid-stride), row in registers.
//   xhat = (x - mean) * rstd ; dyg = dy * gamma ; dx = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat))
//   dx_out = (resid ? resid : 0) + dx ; partials[block] = (sum_r dy * xhat, sum_r dy)
// ---------------------------------------------------------------------------------------------------
template <int DIM, typename OpT, bool DROP>
__global__ void __launch_bounds__(256, DIM <= 512 ? 2 : 1) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ dy, long long rows, float eps,
                                                             const float* resid, float* dx_out,
                                                             float* __restrict__ partials, OpT* __restrict__ dx_op,
                                                             float drop_p, unsigned long long seed,
                                                             const unsigned long long* seed_dev, unsigned site) {
  constexpr int V = DIM / 128;
  // drop_p > 0: the 16-bit copy carries the dropout mask of the sub-layer branch it feeds (gradient entering
  // out_proj / fc2 = g o mask / (1 - p)); the fp32 gradient (the residual path) stays unmasked
  if (DROP && seed_dev) seed += *seed_dev;
  const unsigned thr = dropout_threshold(drop_p);
  const float dinv = 1.0f / (1.0f - drop_p);
  __shared__ float red[8][DIM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) ag[i] = make_float4(0.f, 0.f, 0.f, 0.f), ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * DIM);
    const float4* dr = reinterpret_cast<const float4*>(dy + row * DIM);
    float4 v[V], d[V], rs[V];
    float s = 0.f;
    // every load of the row (x, dy and the residual gradient) is issued before the first reduction: one memory round
    // trip per row instead of two
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = xr[lane + 32 * i];
      d[i] = __ldcs(dr + lane + 32 * i);
      if (dx_out && resid) rs[i] = reinterpret_cast<const float4*>(resid + row * DIM)[lane + 32 * i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / DIM);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / DIM) + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 g = __ldg(g4 + lane + 32 * i);
      v[i].x *= rstd, v[i].y *= rstd, v[i].z *= rstd, v[i].w *= rstd;   // xhat
      ag[i].x += d[i].x * v[i].x, ag[i].y += d[i].y * v[i].y, ag[i].z += d[i].z * v[i].z, ag[i].w += d[i].w * v[i].w;
      ab[i].x += d[i].x, ab[i].y += d[i].y, ab[i].z += d[i].z, ab[i].w += d[i].w;
      d[i].x *= g.x, d[i].y *= g.y, d[i].z *= g.z, d[i].w *= g.w;       // dyg
      c1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      c2 += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    c1 = warp_sum(c1) * (1.0f / DIM);
    c2 = warp_sum(c2) * (1.0f / DIM);
    if (dx_out) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float4 o;
        o.x = rstd * (d[i].x - c1 - v[i].x * c2);
        o.y = rstd * (d[i].y - c1 - v[i].y * c2);
        o.z = rstd * (d[i].z - c1 - v[i].z * c2);
        o.w = rstd * (d[i].w - c1 - v[i].w * c2);
        if (resid) o.x += rs[i].x, o.y += rs[i].y, o.z += rs[i].z, o.w += rs[i].w;
        reinterpret_cast<float4*>(dx_out + row * DIM)[lane + 32 * i] = o;
        if (dx_op) {   // 16-bit copy of the gradient: the A operand of the next dgrad / wgrad GEMMs
          if constexpr (DROP)
            dropout_apply4(dropout_bits4(seed, site, (unsigned long long)(row * (DIM / 4) + lane + 32 * i)), thr, dinv, o.x,
                           o.y, o.z, o.w);
          uint2 pk;
          pk.x = OpTraits<OpT>::pack2(o.x, o.y);
          pk.y = OpTraits<OpT>::pack2(o.z, o.w);
          reinterpret_cast<uint2*>(dx_op + row * DIM)[lane + 32 * i] = pk;
        }
      }
    }
  }
  // block reduction of the parameter-gradient partials (fixed order: warp 0..7)
  float* out = partials + (long long)blockIdx.x * 2 * DIM;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i)
      reinterpret_cast<float4*>(&red[warp][0])[lane + 32 * i] = pass == 0 ? ag[i] : ab[i];
    __syncthreads();
    for (int c = threadIdx.x; c < DIM; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][c];
      out[pass * DIM + c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// softmax backward from recomputed scores: warp per query row.
//   P = softmax(S[:, :valid]) ; D = sum_k P dP ; dS = P (dP - D)   (keys >= valid: P = dS = 0 up to ld_out)
// rows are laid out [batch][rows_per_batch]; valid keys of a batch = kv_lens[batch / heads] (or n_keys).
// ---------------------------------------------------------------------------------------------------
// (counter-based dropout mask: dropout_keep in common.cuh)

// NPL = keys per lane held in registers (n_keys <= 32 * NPL): scores and dP are read once.
// With drop_p > 0 the probabilities carry attention dropout: m = keep(seed, site, row * ld_out + k) / (1 - p);
// P_out = P m (what multiplied V in the forward pass), dP' = dP m, dS = P (dP' - sum_k P dP').  dPv == NULL: forward
// use (softmax + dropout -> P_out only).
template <int NPL, typename OpT, bool DROP, bool BWD>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ S, const void* __restrict__ dPv,
                                                           int dp_is_op, long long ld_dp,
                                                           long long ld_in, long long rows, int rows_per_batch,
                                                           int n_keys, const int* __restrict__ kv_lens, int heads,
                                                           OpT* __restrict__ P, OpT* __restrict__ dS, long long ld_out,
                                                           int valid_rows, int causal, float drop_p,
                                                           unsigned long long seed,
                                                           const unsigned long long* __restrict__ seed_dev,
                                                           unsigned site) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows || (int)(row % rows_per_batch) >= valid_rows) return;
  if (DROP && seed_dev) seed += *seed_dev;
  int valid = n_keys;
  if (kv_lens) {
    const int v = kv_lens[(row / rows_per_batch) / heads];
    valid = v < n_keys ? v : n_keys;
  }
  if (causal) valid = min(valid, (int)(row % rows_per_batch) + 1);   // key j visible to query i iff j <= i
  const float* s = S + row * ld_in;
  // dP is fp32, or 16-bit (written by the GEMM's 16-bit epilogue: half the traffic of the dP = dO v^T round trip)
  const float* d = reinterpret_cast<const float*>(dPv) + row * ld_dp;
  const OpT* d16 = reinterpret_cast<const OpT*>(dPv) + row * ld_dp;
  // lane owns key pairs (2 * (lane + 32 i), +1): 8-byte loads, 4-byte stores
  float sv[NPL], dv[NPL];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NPL / 2; ++i) {
    const int k = 2 * (lane + 32 * i);
    float2 a = make_float2(-INFINITY, -INFINITY), g = make_float2(0.f, 0.f);
    if (k + 1 < valid) {
      a = __ldcs(reinterpret_cast<const float2*>(s + k));
      if constexpr (!BWD) {
      } else if (dp_is_op) {
        const uint32_t q = __ldcs(reinterpret_cast<const uint32_t*>(d16 + k));
        const OpT* e = reinterpret_cast<const OpT*>(&q);
        g = make_float2(OpTraits<OpT>::to_float(e[0]), OpTraits<OpT>::to_float(e[1]));
      } else {
        g = __ldcs(reinterpret_cast<const float2*>(d + k));
      }
    } else if (k < valid) {
      a.x = s[k];
      if constexpr (BWD) g.x = dp_is_op ? OpTraits<OpT>::to_float(d16[k]) : d[k];
    }
    sv[2 * i] = a.x, sv[2 * i + 1] = a.y, dv[2 * i] = g.x, dv[2 * i + 1] = g.y;
    mx = fmaxf(mx, fmaxf(a.x, a.y));
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    sv[i] = (sv[i] == -INFINITY) ? 0.f : __expf(sv[i] - mx);
    sum += sv[i];
  }
  sum = warp_sum(sum);
  const float inv = valid > 0 ? 1.0f / sum : 0.f;
  const float dinv = 1.0f / (1.0f - drop_p);
  float dot = 0.f;
  float mk[DROP ? NPL : 1];     // attention-dropout multiplier of each element (DROP instantiation only)
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    sv[i] *= inv;
    if constexpr (DROP) {
      const int k = 2 * (lane + 32 * (i >> 1)) + (i & 1);
      mk[i] = dropout_keep(seed, site, (unsigned long long)(row * ld_out + k), drop_p) ? dinv : 0.f;
      dv[i] *= mk[i];
    }
    dot += sv[i] * dv[i];
  }
  dot = warp_sum(dot);
  OpT* po = P ? P + row * ld_out : nullptr;
  OpT* go = BWD ? dS + row * ld_out : nullptr;
#pragma unroll
  for (int i = 0; i < NPL / 2; ++i) {
    const int k = 2 * (lane + 32 * i);
    if (k + 1 < ld_out) {
      if (po) {
        if constexpr (DROP)
          *reinterpret_cast<uint32_t*>(po + k) = OpTraits<OpT>::pack2(sv[2 * i] * mk[2 * i], sv[2 * i + 1] * mk[2 * i + 1]);
        else
          *reinterpret_cast<uint32_t*>(po + k) = OpTraits<OpT>::pack2(sv[2 * i], sv[2 * i + 1]);
      }
      if constexpr (BWD)
        *reinterpret_cast<uint32_t*>(go + k) =
            OpTraits<OpT>::pack2(sv[2 * i] * (dv[2 * i] - dot), sv[2 * i + 1] * (dv[2 * i + 1] - dot));
    } else if (k < ld_out) {
      if (po) po[k] = OpTraits<OpT>::cvt(DROP ? sv[2 * i] * mk[DROP ? 2 * i : 0] : sv[2 * i]);
      if constexpr (BWD) go[k] = OpTraits<OpT>::cvt(sv[2 * i] * (dv[2 * i] - dot));
    }
  }
  // columns beyond the register span (ld_out > 32 * NPL cannot happen: checked on the host)
}

template <typename OpT>
static int launch_softmax_bwd(const float* S, const void* dP, int dp_is_op, long long ld_dp, long long ld_in,
                              long long rows, int rpb, int n_keys,
                              const int* kv_lens, int heads, void* P, void* dS, long long ld_out, int valid_rows,
                              int causal, float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                              unsigned site, cudaStream_t s) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  OpT* p = reinterpret_cast<OpT*>(P);
  OpT* g = reinterpret_cast<OpT*>(dS);
#define MM_SMB(NPL, DROP)                                                                                                \
  do {                                                                                                                   \
    if (dP)                                                                                                              \
      softmax_bwd_kernel<NPL, OpT, DROP, true><<<grid, 256, 0, s>>>(S, dP, dp_is_op, ld_dp, ld_in, rows, rpb, n_keys,     \
                                                                     kv_lens, heads, p, g, ld_out, valid_rows, causal,    \
                                                                     drop_p, seed, seed_dev, site);                       \
    else                                                                                                                 \
      softmax_bwd_kernel<NPL, OpT, DROP, false><<<grid, 256, 0, s>>>(S, dP, dp_is_op, ld_dp, ld_in, rows, rpb, n_keys,    \
                                                                      kv_lens, heads, p, g, ld_out, valid_rows, causal,   \
                                                                      drop_p, seed, seed_dev, site);                      \
  } while (0)
  const bool drop = drop_p > 0.f;
  if (ld_out <= 256) {
    if (drop) MM_SMB(8, true); else MM_SMB(8, false);
  } else if (ld_out <= 640) {
    if (drop) MM_SMB(20, true); else MM_SMB(20, false);
  } else if (ld_out <= 2048) {
    if (drop) MM_SMB(64, true); else MM_SMB(64, false);
  } else {
    return bad_arg("softmax_bwd: at most 2048 keys");
  }
#undef MM_SMB
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// GLU backward: y = a * sigmoid(b), pre = [a | b] (fp32, 2n columns); dpre = [dy s sig(b) | dy s a sig(b)(1-sig(b))]
// ---------------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256) glu_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dy,
                                                       long long rows, int n, float scale, OpT* __restrict__ dpre) {
  const long long total = rows * (long long)(n / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (n / 2);
    const int c = (int)(i - r * (n / 2)) * 2;
    const float2 a = *reinterpret_cast<const float2*>(pre + r * 2 * n + c);
    const float2 b = *reinterpret_cast<const float2*>(pre + r * 2 * n + n + c);
    float2 g = *reinterpret_cast<const float2*>(dy + r * n + c);
    g.x *= scale, g.y *= scale;
    const float s0 = sigmoid_exact(b.x), s1 = sigmoid_exact(b.y);
    *reinterpret_cast<uint32_t*>(dpre + r * 2 * n + c) = OpTraits<OpT>::pack2(g.x * s0, g.y * s1);
    *reinterpret_cast<uint32_t*>(dpre + r * 2 * n + n + c) =
        OpTraits<OpT>::pack2(g.x * a.x * s0 * (1.f - s0), g.y * a.y * s1 * (1.f - s1));
  }
}

// ---------------------------------------------------------------------------------------------------
// selective-gate backward: res = (1-g) text + g attn, g = sigmoid(z)  (z includes the bias)
//   dz = dres (attn - text) g (1-g) ; dcat = [dres g | dres (1-g)]   (dcat then receives += dz Wg by the RESID GEMM)
// dres is T x B x C (fairseq layout of the fused states), everything else token-major [B*T, C].
// ---------------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dres,
                                                        const float* __restrict__ text, const float* __restrict__ attn,
                                                        int B, int T, int d, OpT* __restrict__ dz,
                                                        float* __restrict__ dcat) {
  const long long total = (long long)B * T * (d / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (d / 2);
    const int c = (int)(i - r * (d / 2)) * 2;
    const int b = (int)(r / T), t = (int)(r - (long long)b * T);
    const float2 g_ = *reinterpret_cast<const float2*>(dres + ((long long)t * B + b) * d + c);
    const float2 zz = *reinterpret_cast<const float2*>(z + r * d + c);
    const float2 tx = *reinterpret_cast<const float2*>(text + r * d + c);
    const float2 at = *reinterpret_cast<const float2*>(attn + r * d + c);
    const float g0 = sigmoid_exact(zz.x), g1 = sigmoid_exact(zz.y);
    *reinterpret_cast<uint32_t*>(dz + r * d + c) =
        OpTraits<OpT>::pack2(g_.x * (at.x - tx.x) * g0 * (1.f - g0), g_.y * (at.y - tx.y) * g1 * (1.f - g1));
    *reinterpret_cast<float2*>(dcat + r * 2 * d + c) = make_float2(g_.x * g0, g_.y * g1);
    *reinterpret_cast<float2*>(dcat + r * 2 * d + d + c) = make_float2(g_.x * (1.f - g0), g_.y * (1.f - g1));
  }
}

// T x B x C fp32 -> token-major [B*T, C] fp32 (gradient of the un-gated `text + attn` residual mode and of the
// no-fusion path, where the fused states are a plain transpose of the final LayerNorm output)
__global__ void __launch_bounds__(256) tbc_to_btc_kernel(const float* __restrict__ in, int B, int T, int d,
                                                          float* __restrict__ out) {
  const long long total = (long long)B * T * (d / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (d / 4);
    const int c = (int)(i - r * (d / 4)) * 4;
    const int b = (int)(r / T), t = (int)(r - (long long)b * T);
    *reinterpret_cast<float4*>(out + r * d + c) = *reinterpret_cast<const float4*>(in + ((long long)t * B + b) * d + c);
  }
}

// ---------------------------------------------------------------------------------------------------
// col2im of a k=5, stride-2, pad-2 Conv1d: dx[b, s, c] = sum over taps k with (s + 2 - k) even, t = (s+2-k)/2 in range
// of dcol[b, t, k, c].   dcol fp32 [B, T_out, 5*C], dx fp32 [B, T_in, C].
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col2im_k5s2_kernel(const float* __restrict__ dcol, int B, int T_out, int T_in,
                                                           int C, float* __restrict__ dx) {
  const long long total = (long long)B * T_in * (C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C / 4);
    const int c = (int)(i - r * (C / 4)) * 4;
    const int b = (int)(r / T_in), s = (int)(r - (long long)b * T_in);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int u = s + 2 - k;
      if (u < 0 || (u & 1)) continue;
      const int t = u >> 1;
      if (t >= T_out) continue;
      const float4 v = *reinterpret_cast<const float4*>(dcol + (((long long)b * T_out + t) * 5 + k) * C + c);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dx + r * C + c) = acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// Optimizer: sum of squares (per-block partials), clip coefficient, fairseq Adam.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

// out[0] = grad norm (after grad_scale), out[1] = multiplier to apply to the raw gradient: grad_scale * min(1, max_norm/(norm+1e-6))
__global__ void clip_coef_kernel(const float* __restrict__ partials, int S, float grad_scale, float max_norm,
                                 float* __restrict__ out, int dev_hyper, const float* __restrict__ extra_norm) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (dev_hyper) grad_scale = out[4], max_norm = out[5];   // per-step values written by the host before a graph replay
  double t = 0.0;
  for (int i = 0; i < S; ++i) t += (double)partials[i];
  float norm = sqrtf((float)t) * grad_scale;
  if (extra_norm) norm = sqrtf(norm * norm + extra_norm[0] * extra_norm[0]);   // joint norm with another flat buffer
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(1.0f, max_norm / (norm + 1e-6f));
  out[0] = norm;
  out[1] = coef;
}

// fairseq.optim.adam.Adam.step: m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; denom = sqrt(v) + eps ;
// step_size = lr sqrt(1-b2^t) / (1-b1^t) ; p -= wd lr p ; p -= step_size m / denom
template <typename OpT>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                    float beta1, float beta2, float eps, float weight_decay,
                                                    float step_size, const float* __restrict__ coef,
                                                    OpT* __restrict__ p_op, int dev_hyper) {
  const float gs = coef ? coef[1] : 1.0f;
  float wd_lr = weight_decay * lr;
  if (dev_hyper) step_size = coef[2], wd_lr = coef[3];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi, v[i] = vi;
    float pi = p[i];
    if (wd_lr != 0.f) pi -= wd_lr * pi;
    pi -= step_size * mi / (sqrtf(vi) + eps);
    p[i] = pi;
    if (p_op) p_op[i] = OpTraits<OpT>::cvt(pi);
  }
}

// ---------------------------------------------------------------------------------------------------
// label-smoothed cross entropy backward (fairseq label_smoothed_nll_loss, reduce=True):
//   loss = (1 - eps - eps_i) * nll + eps_i * smooth, eps_i = eps / (V - 1)
//   d loss / d logit_j = (1 - eps - eps_i) (p_j - [j == t]) + eps_i (V p_j - 1)   (0 for rows whose target is padding)
// warp per row; dlogits 16-bit [rows, ld_out], columns [vocab, ld_out) = 0.
// ---------------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, long long ld, int vocab,
                                                      const long long* __restrict__ target, int padding_idx, long long rows,
                                                      float eps, float grad_scale, OpT* __restrict__ dlogits,
                                                      long long ld_out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* x = logits + row * ld;
  OpT* o = dlogits + row * ld_out;
  const long long t = target[row];
  if (t == padding_idx || t < 0 || t >= vocab) {   // the same rows mm_label_smoothed_nll ignores
    for (int j = lane; j < ld_out; j += 32) o[j] = OpTraits<OpT>::cvt(0.f);
    return;
  }
  float mx = -INFINITY;
  for (int j = lane; j < vocab; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < vocab; j += 32) sum += __expf(x[j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  const float eps_i = eps / (float)(vocab - 1);
  const float a = 1.0f - eps - eps_i;
  for (int j = lane; j < ld_out; j += 32) {
    float g = 0.f;
    if (j < vocab) {
      const float p = __expf(x[j] - mx) * inv;
      g = a * (p - (j == t ? 1.0f : 0.0f)) + eps_i * ((float)vocab * p - 1.0f);
    }
    o[j] = OpTraits<OpT>::cvt(g * grad_scale);
  }
}

// embedding backward, deterministic: one block per vocabulary row scans the token list in order and sums the gradient
// rows of its token (no atomics: the result does not depend on scheduling).  table_grad[v] += scale * sum_{tokens[r]==v} dx[r]
// (nn.Embedding with padding_idx gives that row no gradient).  The token list is staged through shared memory.
__global__ void __launch_bounds__(256) embed_bwd_kernel(const long long* __restrict__ tokens, int padding_idx,
                                                         const float* __restrict__ dx, long long rows, int dim,
                                                         float scale, float* __restrict__ table_grad) {
  __shared__ int hits[1024];
  __shared__ int warp_cnt[8];
  __shared__ int n_hits;
  const int v = blockIdx.x;
  if (v == padding_idx) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};      // dim <= 1024: columns threadIdx.x + 256 j
  for (long long base = 0; base < rows; base += 1024) {
    // ordered compaction of the chunk's matching rows: warp w owns rows [128 w, 128 w + 128), four ballots each
    unsigned m[4];
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long r = base + warp * 128 + q * 32 + lane;
      m[q] = __ballot_sync(0xffffffffu, r < rows && tokens[r] == v);
      cnt += __popc(m[q]);
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (threadIdx.x == 255) n_hits = off + cnt;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if ((m[q] >> lane) & 1u) hits[off + __popc(m[q] & ((1u << lane) - 1u))] = warp * 128 + q * 32 + lane;
      off += __popc(m[q]);
    }
    __syncthreads();
    const int n = n_hits;
    for (int h = 0; h < n; ++h) {
      const float* src = dx + (base + hits[h]) * dim;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = threadIdx.x + 256 * j;
        if (c < dim) acc[j] += src[c];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = threadIdx.x + 256 * j;
    if (c < dim) table_grad[(long long)v * dim + c] += scale * acc[j];
  }
}

// ---------------------------------------------------------------------------------------------------
// Element-wise dropout (fairseq FairseqDropout = F.dropout in training): y = keep ? x / (1 - p) : 0 with a counter-based
// mask: keep(seed, site, i) is a pure function, so the backward pass (and the parity tests) regenerate the same mask
// from (seed, site) instead of storing it.  Not torch's Philox stream (a drop-in cannot share torch's generator state
// with a fused kernel anyway); splitmix64 finaliser of the element counter.
// ---------------------------------------------------------------------------------------------------
// out = resid + dropout(x)  (resid optional); x / out fp32 or 16-bit (same type), n elements, in place allowed
template <typename T, typename OpT>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, const float* __restrict__ resid,
                                                       T* __restrict__ out, long long n, float p,
                                                       unsigned long long seed, const unsigned long long* seed_dev,
                                                       unsigned site) {
  if (seed_dev) seed += *seed_dev;      // per-step seed held in device memory: CUDA-graph replays get fresh masks
  const float inv = 1.0f / (1.0f - p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v;
    if constexpr (sizeof(T) == 4) v = (float)x[i]; else v = OpTraits<OpT>::to_float(reinterpret_cast<const OpT*>(x)[i]);
    v = dropout_keep(seed, site, (unsigned long long)i, p) ? v * inv : 0.f;
    if (resid) v += resid[i];
    if constexpr (sizeof(T) == 4) out[i] = (T)v; else reinterpret_cast<OpT*>(out)[i] = OpTraits<OpT>::cvt(v);
  }
}

static inline unsigned grid_for(long long total, int per_block = 256) {
  long long g = (total + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mm

using namespace mm;

extern "C" int mm_pack_t(const void* in, int32_t in_is_f32, int64_t in_ld, int64_t in_bs0, int64_t in_bs1, int32_t nb1,
                         const void* mask, int64_t mask_ld, int32_t rows, int32_t cols, int32_t batches, float scale,
                         void* out_n, int64_t n_ld, int64_t n_bs0, int64_t n_bs1, void* out_t, int64_t t_ld,
                         int64_t t_bs0, int64_t t_bs1, int32_t t_cols_pad, int32_t dtype, void* stream) {
  if (!in || (!out_n && !out_t)) return bad_arg("pack_t: null operand");
  if (rows <= 0 || cols <= 0 || batches <= 0 || nb1 <= 0) return bad_arg("pack_t: non-positive extent");
  if ((cols & 1) || (in_ld & 1) || (in_bs0 & 1) || (in_bs1 & 1) || (n_ld & 1) || (t_ld & 1) || (mask_ld & 1) ||
      (n_bs0 & 1) || (n_bs1 & 1) || (t_bs0 & 1) || (t_bs1 & 1))
    return bad_arg("pack_t: cols and strides must be even");
  if (mask && batches != 1) return bad_arg("pack_t: mask needs batches == 1");
  if (dtype != MM_DTYPE_BF16 && dtype != MM_DTYPE_F16) return bad_arg("dtype");
  const int rows_even = rows + (rows & 1);
  if (out_t) {
    if (t_cols_pad < rows_even) t_cols_pad = rows_even;
    if ((t_cols_pad & 1) || t_ld < t_cols_pad) return bad_arg("pack_t: t_cols_pad must be even and <= t_ld");
  } else {
    t_cols_pad = rows;
  }
  PackArgs a;
  a.in = in, a.mask = mask, a.out_n = out_n, a.out_t = out_t;
  a.in_ld = in_ld, a.in_bs0 = in_bs0, a.in_bs1 = in_bs1, a.mask_ld = mask_ld;
  a.n_ld = n_ld, a.n_bs0 = n_bs0, a.n_bs1 = n_bs1, a.t_ld = t_ld, a.t_bs0 = t_bs0, a.t_bs1 = t_bs1;
  a.rows = rows, a.cols = cols, a.nb1 = nb1, a.t_cols_pad = t_cols_pad, a.in_is_f32 = in_is_f32, a.scale = scale;
  const int span = t_cols_pad > rows ? t_cols_pad : rows;
  dim3 grid((span + 63) / 64, (cols + 63) / 64, batches);
  if (grid.y > 65535 || grid.z > 65535) return bad_arg("pack_t: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MM_DTYPE_F16)
    pack_t_kernel<__half><<<grid, 256, 0, s>>>(a);
  else
    pack_t_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a);
  MM_CHECK_LAUNCH("pack_t_kernel launch");
  return 0;
}

extern "C" int mm_rowsum(const void* in, int64_t ld, int32_t rows, int32_t cols, float* out, int32_t accumulate,
                         int32_t dtype, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0) return bad_arg("rowsum");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = rows;
  if (dtype == MM_DTYPE_F16)
    rowsum_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(in), ld, rows, cols, out, accumulate);
  else
    rowsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld, rows, cols, out,
                                                       accumulate);
  MM_CHECK_LAUNCH("rowsum_kernel launch");
  return 0;
}

extern "C" int mm_colsum_blocks(int32_t rows) { return (rows + mm::COLSUM_ROWS - 1) / mm::COLSUM_ROWS; }

extern "C" int mm_colsum(const void* in, int64_t ld, int32_t rows, int32_t cols, int32_t period, int32_t valid,
                         float* partials, int32_t dtype, void* stream) {
  if (!in || !partials || rows <= 0 || cols <= 0 || (cols & 1) || (ld & 1)) return bad_arg("colsum");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool per = period > 0;
  if (per && valid <= 0) return bad_arg("colsum: period needs valid > 0");
  if (dtype == MM_DTYPE_F16)
    return per ? launch_colsum<__half, true>(in, ld, rows, cols, period, valid, partials, s)
               : launch_colsum<__half, false>(in, ld, rows, cols, period, valid, partials, s);
  return per ? launch_colsum<__nv_bfloat16, true>(in, ld, rows, cols, period, valid, partials, s)
             : launch_colsum<__nv_bfloat16, false>(in, ld, rows, cols, period, valid, partials, s);
}

extern "C" int mm_reduce_partials(const float* part, int32_t n_partials, int64_t stride, int64_t n, float* out,
                                  int32_t accumulate, void* stream) {
  if (!part || !out || n_partials <= 0 || n <= 0) return bad_arg("reduce_partials");
  reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      part, n_partials, stride, n, out, accumulate);
  MM_CHECK_LAUNCH("reduce_partials_kernel launch");
  return 0;
}

extern "C" int mm_reduce_partials_many(const mm_reduce_job* jobs, int32_t count, void* stream) {
  if (!jobs || count <= 0) return bad_arg("reduce_partials_many");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int base = 0; base < count; base += 16) {
    ReduceJobs t;
    memset(&t, 0, sizeof(t));
    const int m = count - base < 16 ? count - base : 16;
    long long gx = 1;
    for (int i = 0; i < m; ++i) {
      const mm_reduce_job& j = jobs[base + i];
      if (!j.part || !j.out || j.n_partials <= 0 || j.n <= 0) return bad_arg("reduce_partials_many: job");
      t.part[i] = j.part, t.out[i] = j.out, t.stride[i] = j.stride, t.n[i] = j.n, t.S[i] = j.n_partials;
      t.accumulate[i] = j.accumulate;
      int sl = 1;                                   // at least two partials per slice, at most 16 slices
      while (sl < 16 && 4 * sl <= j.n_partials) sl *= 2;
      t.slices[i] = sl;
      const long long tiles = ((j.n + 3) / 4 + 512 / sl - 1) / (512 / sl);
      if (tiles > gx) gx = tiles;
    }
    if (gx > 592) gx = 592;                         // 4 blocks per SM per job; a block walks its tiles
    reduce_many_kernel<<<dim3((unsigned)gx, (unsigned)m), 256, 0, s>>>(t);
    MM_CHECK_LAUNCH("reduce_many_kernel launch");
  }
  return 0;
}

extern "C" int mm_layernorm_bwd_blocks(void) { return 2 * kNumSMs; }

template <typename OpT>
static int launch_ln_bwd(const float* x, const float* gamma, const float* dy, long long rows, int dim, float eps,
                         const float* resid, float* dx, float* partials, void* dx_op, float drop_p,
                         unsigned long long seed, const unsigned long long* seed_dev, unsigned site, cudaStream_t s) {
  const unsigned grid = 2 * kNumSMs;
  OpT* o = reinterpret_cast<OpT*>(dx_op);
#define MM_LNB(D)                                                                                                        \
  do {                                                                                                                   \
    if (drop_p > 0.f)                                                                                                    \
      layernorm_bwd_kernel<D, OpT, true><<<grid, 256, 0, s>>>(x, gamma, dy, rows, eps, resid, dx, partials, o, drop_p,  \
                                                              seed, seed_dev, site);                                     \
    else                                                                                                                 \
      layernorm_bwd_kernel<D, OpT, false><<<grid, 256, 0, s>>>(x, gamma, dy, rows, eps, resid, dx, partials, o, drop_p, \
                                                               seed, seed_dev, site);                                    \
  } while (0)
  switch (dim) {
    case 256: MM_LNB(256); break;
    case 512: MM_LNB(512); break;
    case 768: MM_LNB(768); break;
    case 1024: MM_LNB(1024); break;
    default: return bad_arg("layernorm_bwd dim must be 256, 512, 768 or 1024");
  }
#undef MM_LNB
  return 0;
}

extern "C" int mm_layernorm_bwd_drop(const float* x, const float* gamma, const float* dy, int64_t rows, int32_t dim,
                                     float eps, const float* resid, float* dx, float* partials, void* dx_op,
                                     float drop_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                                     int32_t dtype, void* stream) {
  if (!x || !gamma || !dy || !partials || rows <= 0) return bad_arg("layernorm_bwd");
  if (dx_op && !dx) return bad_arg("layernorm_bwd: dx_op needs dx");
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !dx_op)) return bad_arg("layernorm_bwd: dropout needs dx_op and p in [0, 1)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  int rc = dtype == MM_DTYPE_F16
               ? launch_ln_bwd<__half>(x, gamma, dy, rows, dim, eps, resid, dx, partials, dx_op, drop_p, seed, sd, site, s)
               : launch_ln_bwd<__nv_bfloat16>(x, gamma, dy, rows, dim, eps, resid, dx, partials, dx_op, drop_p, seed, sd, site, s);
  if (rc) return rc;
  MM_CHECK_LAUNCH("layernorm_bwd_kernel launch");
  return 0;
}

extern "C" int mm_layernorm_bwd(const float* x, const float* gamma, const float* dy, int64_t rows, int32_t dim, float eps,
                                const float* resid, float* dx, float* partials, void* dx_op, int32_t dtype,
                                void* stream) {
  return mm_layernorm_bwd_drop(x, gamma, dy, rows, dim, eps, resid, dx, partials, dx_op, 0.f, 0, nullptr, 0, dtype, stream);
}

extern "C" int mm_softmax_bwd(const float* scores, const void* dprobs, int32_t dprobs_is_op, int64_t ld_dprobs,
                              int64_t ld_in, int64_t rows, int32_t rows_per_batch, int32_t n_keys, const int32_t* kv_lens,
                              int32_t heads, void* probs, void* dscores, int64_t ld_out, int32_t valid_rows, int32_t causal,
                              int32_t dtype, void* stream) {
  return mm_softmax_dropout_bwd(scores, dprobs, dprobs_is_op, ld_dprobs, ld_in, rows, rows_per_batch, n_keys, kv_lens,
                                heads, probs, dscores, ld_out, valid_rows, causal, 0.f, 0, nullptr, 0, dtype, stream);
}

extern "C" int mm_softmax_dropout_bwd(const float* scores, const void* dprobs, int32_t dprobs_is_op, int64_t ld_dprobs,
                                      int64_t ld_in, int64_t rows, int32_t rows_per_batch, int32_t n_keys,
                                      const int32_t* kv_lens, int32_t heads, void* probs, void* dscores, int64_t ld_out,
                                      int32_t valid_rows, int32_t causal, float drop_p, uint64_t seed,
                                      const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream) {
  if (valid_rows <= 0) valid_rows = rows_per_batch;
  if (!scores || (!dscores && !probs) || (dprobs && !dscores) || (!dprobs && dscores) || rows <= 0 || n_keys <= 0 ||
      rows_per_batch <= 0 || heads <= 0 || drop_p < 0.f || drop_p >= 1.f)
    return bad_arg("softmax_bwd");
  if (ld_out < n_keys || ld_in < n_keys) return bad_arg("softmax_bwd: leading dimensions");
  if (!dprobs) ld_dprobs = ld_in;
  if ((ld_in & 1) || (ld_out & 1) || (ld_dprobs & 1) || ld_dprobs < n_keys ||
      (reinterpret_cast<uintptr_t>(scores) & 7) || (reinterpret_cast<uintptr_t>(dprobs) & 7))
    return bad_arg("softmax_bwd: leading dimensions must be even and the inputs 8-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  const int rc = dtype == MM_DTYPE_F16
                     ? launch_softmax_bwd<__half>(scores, dprobs, dprobs_is_op, ld_dprobs, ld_in, rows, rows_per_batch,
                                                  n_keys, kv_lens, heads, probs, dscores, ld_out, valid_rows, causal,
                                                  drop_p, seed, sd, site, s)
                     : launch_softmax_bwd<__nv_bfloat16>(scores, dprobs, dprobs_is_op, ld_dprobs, ld_in, rows,
                                                         rows_per_batch, n_keys, kv_lens, heads, probs, dscores, ld_out,
                                                         valid_rows, causal, drop_p, seed, sd, site, s);
  if (rc) return rc;
  MM_CHECK_LAUNCH("softmax_bwd_kernel launch");
  return 0;
}

extern "C" int mm_label_smoothed_nll_bwd(const float* logits, int64_t ld, int32_t vocab, const int64_t* target,
                                         int32_t padding_idx, int64_t rows, float epsilon, float grad_scale,
                                         void* dlogits, int64_t ld_out, int32_t dtype, void* stream) {
  if (!logits || !target || !dlogits || rows <= 0 || vocab <= 1 || ld < vocab || ld_out < vocab)
    return bad_arg("label_smoothed_nll_bwd");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const long long* t = reinterpret_cast<const long long*>(target);
  if (dtype == MM_DTYPE_F16)
    ce_bwd_kernel<__half><<<grid, 256, 0, s>>>(logits, ld, vocab, t, padding_idx, rows, epsilon, grad_scale,
                                               reinterpret_cast<__half*>(dlogits), ld_out);
  else
    ce_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(logits, ld, vocab, t, padding_idx, rows, epsilon, grad_scale,
                                                      reinterpret_cast<__nv_bfloat16*>(dlogits), ld_out);
  MM_CHECK_LAUNCH("ce_bwd_kernel launch");
  return 0;
}

extern "C" int mm_embed_tokens_bwd(const int64_t* tokens, int32_t padding_idx, const float* dx, int64_t rows, int32_t dim,
                                   float scale, float* table_grad, int32_t vocab, void* stream) {
  if (!tokens || !dx || !table_grad || rows <= 0 || dim <= 0 || dim > 1024 || vocab <= 0) return bad_arg("embed_tokens_bwd");
  embed_bwd_kernel<<<(unsigned)vocab, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(tokens), padding_idx, dx, rows, dim, scale, table_grad);
  MM_CHECK_LAUNCH("embed_bwd_kernel launch");
  return 0;
}

extern "C" int mm_dropout(const void* x, int32_t x_is_f32, const float* resid, void* out, int64_t n, float p,
                          uint64_t seed, const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream) {
  if (!x || !out || n <= 0 || p < 0.f || p >= 1.f) return bad_arg("dropout");
  if (resid && !x_is_f32) return bad_arg("dropout: the residual form is fp32");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = grid_for(n);
  if (x_is_f32)
    dropout_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), resid,
                                                              reinterpret_cast<float*>(out), n, p, seed,
                                                              reinterpret_cast<const unsigned long long*>(seed_dev), site);
  else if (dtype == MM_DTYPE_F16)
    dropout_kernel<__half, __half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(x), nullptr,
                                                        reinterpret_cast<__half*>(out), n, p, seed,
                                                        reinterpret_cast<const unsigned long long*>(seed_dev), site);
  else
    dropout_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), nullptr,
                                                                      reinterpret_cast<__nv_bfloat16*>(out), n, p, seed,
                                                                      reinterpret_cast<const unsigned long long*>(seed_dev),
                                                                      site);
  MM_CHECK_LAUNCH("dropout_kernel launch");
  return 0;
}

extern "C" int mm_glu_bwd(const float* pre, const float* dy, int64_t rows, int32_t n, float scale, void* dpre,
                          int32_t dtype, void* stream) {
  if (!pre || !dy || !dpre || rows <= 0 || n <= 0 || (n & 1)) return bad_arg("glu_bwd");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = grid_for(rows * (long long)(n / 2));
  if (dtype == MM_DTYPE_F16)
    glu_bwd_kernel<__half><<<grid, 256, 0, s>>>(pre, dy, rows, n, scale, reinterpret_cast<__half*>(dpre));
  else
    glu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(pre, dy, rows, n, scale, reinterpret_cast<__nv_bfloat16*>(dpre));
  MM_CHECK_LAUNCH("glu_bwd_kernel launch");
  return 0;
}

extern "C" int mm_gate_bwd(const float* z, const float* dres_tbc, const float* text, const float* attn, int32_t batch,
                           int32_t seq, int32_t dim, void* dz, float* dcat, int32_t dtype, void* stream) {
  if (!z || !dres_tbc || !text || !attn || !dz || !dcat || batch <= 0 || seq <= 0 || dim <= 0 || (dim & 1))
    return bad_arg("gate_bwd");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = grid_for((long long)batch * seq * (dim / 2));
  if (dtype == MM_DTYPE_F16)
    gate_bwd_kernel<__half><<<grid, 256, 0, s>>>(z, dres_tbc, text, attn, batch, seq, dim,
                                                 reinterpret_cast<__half*>(dz), dcat);
  else
    gate_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(z, dres_tbc, text, attn, batch, seq, dim,
                                                        reinterpret_cast<__nv_bfloat16*>(dz), dcat);
  MM_CHECK_LAUNCH("gate_bwd_kernel launch");
  return 0;
}

extern "C" int mm_tbc_to_btc(const float* in_tbc, int32_t batch, int32_t seq, int32_t dim, float* out, void* stream) {
  if (!in_tbc || !out || batch <= 0 || seq <= 0 || dim <= 0 || (dim & 3)) return bad_arg("tbc_to_btc");
  tbc_to_btc_kernel<<<grid_for((long long)batch * seq * (dim / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in_tbc, batch, seq, dim, out);
  MM_CHECK_LAUNCH("tbc_to_btc_kernel launch");
  return 0;
}

extern "C" int mm_col2im_k5s2(const float* dcol, int32_t batch, int32_t t_out, int32_t t_in, int32_t channels, float* dx,
                              void* stream) {
  if (!dcol || !dx || batch <= 0 || t_out <= 0 || t_in <= 0 || channels <= 0 || (channels & 3)) return bad_arg("col2im");
  col2im_k5s2_kernel<<<grid_for((long long)batch * t_in * (channels / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dcol, batch, t_out, t_in, channels, dx);
  MM_CHECK_LAUNCH("col2im_k5s2_kernel launch");
  return 0;
}

extern "C" int mm_sumsq_blocks(void) { return 4 * kNumSMs; }

extern "C" int mm_grad_clip_coef(const float* grad, int64_t n, float grad_scale, float max_norm, float* partials,
                                 float* norm_coef, int32_t dev_hyper, const float* extra_norm, void* stream) {
  if (!grad || !partials || !norm_coef || n <= 0) return bad_arg("grad_clip_coef");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = 4 * kNumSMs;
  sumsq_kernel<<<blocks, 256, 0, s>>>(grad, n, partials);
  MM_CHECK_LAUNCH("sumsq_kernel launch");
  clip_coef_kernel<<<1, 32, 0, s>>>(partials, blocks, grad_scale, max_norm, norm_coef, dev_hyper, extra_norm);
  MM_CHECK_LAUNCH("clip_coef_kernel launch");
  return 0;
}

extern "C" int mm_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int32_t step, const float* norm_coef,
                       void* param_op, int32_t dtype, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 0) return bad_arg("adam");
  if (step == 0 && !norm_coef) return bad_arg("adam: step 0 reads step_size / wd*lr from norm_coef[2..3]");
  const int dev_hyper = step == 0;
  float step_size = 0.f;
  if (step > 0) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    step_size = (float)((double)lr * sqrt(bc2) / bc1);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MM_DTYPE_F16)
    adam_kernel<__half><<<grid_for(n), 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                    weight_decay, step_size, norm_coef,
                                                    reinterpret_cast<__half*>(param_op), dev_hyper);
  else
    adam_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                           weight_decay, step_size, norm_coef,
                                                           reinterpret_cast<__nv_bfloat16*>(param_op), dev_hyper);
  MM_CHECK_LAUNCH("adam_kernel launch");
  return 0;
}
