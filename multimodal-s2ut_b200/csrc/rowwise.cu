// HBM-bound row-wise kernels: LayerNorm, key-softmax for the speech->image attention, fp32->16-bit
// conversion, CMVN application / zero padding, subsampled sequence lengths.
// All are one-warp-per-row (or grid-stride) kernels with 128-bit loads/stores and warp-shuffle reductions.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

// ---------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (DIM/32 floats per lane), two-pass statistics.
// ---------------------------------------------------------------------------------------------------
template <int DIM, typename OpT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, long long rows,
                                                         OpT* __restrict__ out_op, float* __restrict__ out_f32,
                                                         float eps) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  constexpr int V = DIM / 128;  // float4 per lane
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * DIM);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = __ldcs(xr + lane + 32 * i);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / DIM);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / DIM) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    float4 y;
    y.x = fmaf(v[i].x * rstd, g.x, b.x);
    y.y = fmaf(v[i].y * rstd, g.y, b.y);
    y.z = fmaf(v[i].z * rstd, g.z, b.z);
    y.w = fmaf(v[i].w * rstd, g.w, b.w);
    if (out_f32) reinterpret_cast<float4*>(out_f32 + row * DIM)[lane + 32 * i] = y;
    if (out_op) {
      uint2 pk;
      pk.x = OpTraits<OpT>::pack2(y.x, y.y);
      pk.y = OpTraits<OpT>::pack2(y.z, y.w);
      reinterpret_cast<uint2*>(out_op + row * DIM)[lane + 32 * i] = pk;
    }
  }
}

template <typename OpT>
static int launch_ln(const float* x, const float* g, const float* b, long long rows, int dim, void* out_op,
                     float* out_f32, float eps, cudaStream_t s) {
  const int rows_per_block = 8;
  const unsigned grid = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
  OpT* o = reinterpret_cast<OpT*>(out_op);
  switch (dim) {
    case 256: launch_pdl(layernorm_kernel<256, OpT>, dim3(grid), dim3(256), 0, s, x, g, b, rows, o, out_f32, eps); break;
    case 512: launch_pdl(layernorm_kernel<512, OpT>, dim3(grid), dim3(256), 0, s, x, g, b, rows, o, out_f32, eps); break;
    case 768: launch_pdl(layernorm_kernel<768, OpT>, dim3(grid), dim3(256), 0, s, x, g, b, rows, o, out_f32, eps); break;
    case 1024: launch_pdl(layernorm_kernel<1024, OpT>, dim3(grid), dim3(256), 0, s, x, g, b, rows, o, out_f32, eps); break;
    default: return bad_arg("layernorm dim must be 256, 512, 768 or 1024");
  }
  MM_CHECK_LAUNCH("layernorm_kernel launch");
  return 0;
}

// LayerNorm over rows GATHERED from a device-resident 16-bit feature store: output row r is input row
// index[r / rows_per_index] * rows_per_index + r % rows_per_index.  This is the image pre-norm of a batch whose ViT
// features live on the GPU (ImageFeatureStore) - the batch is never materialised, the store is read once, in 16 bit.
template <int DIM, typename InT, typename OpT>
__global__ void __launch_bounds__(256) layernorm_gather_kernel(const InT* __restrict__ x,
                                                                const long long* __restrict__ index,
                                                                int rows_per_index, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, long long rows,
                                                                OpT* __restrict__ out_op, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int V = DIM / 256;  // 8 values (one 128-bit load) per lane per step
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long src = index ? index[row / rows_per_index] * rows_per_index + row % rows_per_index : row;
  const uint4* xr = reinterpret_cast<const uint4*>(x + src * DIM);
  float v[V][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const uint4 q = __ldcs(xr + lane + 32 * i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const InT lo = reinterpret_cast<const InT*>(&w[j])[0], hi = reinterpret_cast<const InT*>(&w[j])[1];
      v[i][2 * j] = OpTraits<InT>::to_float(lo);
      v[i][2 * j + 1] = OpTraits<InT>::to_float(hi);
      s += v[i][2 * j] + v[i][2 * j + 1];
    }
  }
  const float mean = warp_sum(s) * (1.0f / DIM);
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[i][j] -= mean;
      q2 += v[i][j] * v[i][j];
    }
  const float rstd = rsqrtf(warp_sum(q2) * (1.0f / DIM) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint4* orow = reinterpret_cast<uint4*>(out_op + row * DIM);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c8 = lane + 32 * i;   // group of 8 columns
    const float4 g0 = __ldg(g4 + 2 * c8), g1 = __ldg(g4 + 2 * c8 + 1), b0 = __ldg(b4 + 2 * c8), b1 = __ldg(b4 + 2 * c8 + 1);
    uint4 o;
    o.x = OpTraits<OpT>::pack2(fmaf(v[i][0] * rstd, g0.x, b0.x), fmaf(v[i][1] * rstd, g0.y, b0.y));
    o.y = OpTraits<OpT>::pack2(fmaf(v[i][2] * rstd, g0.z, b0.z), fmaf(v[i][3] * rstd, g0.w, b0.w));
    o.z = OpTraits<OpT>::pack2(fmaf(v[i][4] * rstd, g1.x, b1.x), fmaf(v[i][5] * rstd, g1.y, b1.y));
    o.w = OpTraits<OpT>::pack2(fmaf(v[i][6] * rstd, g1.z, b1.z), fmaf(v[i][7] * rstd, g1.w, b1.w));
    orow[c8] = o;
  }
}

template <typename InT, typename OpT>
static int launch_ln_gather(const void* x, const long long* index, int rows_per_index, const float* g, const float* b,
                            long long rows, int dim, void* out_op, float eps, cudaStream_t s) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const InT* xi = reinterpret_cast<const InT*>(x);
  OpT* o = reinterpret_cast<OpT*>(out_op);
  switch (dim) {
    case 256: launch_pdl(layernorm_gather_kernel<256, InT, OpT>, dim3(grid), dim3(256), 0, s, xi, index, rows_per_index, g, b, rows, o, eps); break;
    case 512: launch_pdl(layernorm_gather_kernel<512, InT, OpT>, dim3(grid), dim3(256), 0, s, xi, index, rows_per_index, g, b, rows, o, eps); break;
    case 768: launch_pdl(layernorm_gather_kernel<768, InT, OpT>, dim3(grid), dim3(256), 0, s, xi, index, rows_per_index, g, b, rows, o, eps); break;
    case 1024: launch_pdl(layernorm_gather_kernel<1024, InT, OpT>, dim3(grid), dim3(256), 0, s, xi, index, rows_per_index, g, b, rows, o, eps); break;
    default: return bad_arg("layernorm_gather dim must be 256, 512, 768 or 1024");
  }
  MM_CHECK_LAUNCH("layernorm_gather_kernel launch");
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Row softmax over image keys: scores fp32 [rows, ld_in] -> probabilities 16-bit [rows, ld_out].
// One warp per row; the row (<= 1024 keys) lives in registers.
// ---------------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ scores, long long ld_in,
                                                            long long rows, int n_keys,
                                                            const uint8_t* __restrict__ key_mask, int rows_per_seq,
                                                            OpT* __restrict__ probs, long long ld_out) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  constexpr int MAXV = 32;  // up to 1024 keys
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = scores + row * ld_in;
  const uint8_t* mk = key_mask ? key_mask + (row / rows_per_seq) * (long long)n_keys : nullptr;
  float v[MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    float s = -INFINITY;
    if (c < n_keys) {
      s = __ldcs(sr + c);
      if (mk && mk[c]) s = -INFINITY;
    }
    v[i] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  const float mref = (mx == -INFINITY) ? 0.f : mx;  // fully masked row -> NaN like the reference (0/0)
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    v[i] = __expf(v[i] - mref);
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  OpT* pr = probs + row * ld_out;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < ld_out) pr[c] = OpTraits<OpT>::cvt(c < n_keys ? v[i] * inv : 0.f);
  }
}

// Vector version for 16-byte aligned rows (ld_in, ld_out multiples of 4): each lane owns V4 groups of four adjacent
// keys -> 128-bit score loads, 64-bit probability stores, and no work beyond ceil(ld_out / 128) groups.
template <int V4, typename OpT>
__global__ void __launch_bounds__(256) softmax_rows_vec_kernel(const float* __restrict__ scores, long long ld_in,
                                                                long long rows, int n_keys,
                                                                const uint8_t* __restrict__ key_mask, int rows_per_seq,
                                                                OpT* __restrict__ probs, long long ld_out) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* sr = reinterpret_cast<const float4*>(scores + row * ld_in);
  const uint8_t* mk = key_mask ? key_mask + (row / rows_per_seq) * (long long)n_keys : nullptr;
  float v[V4][4];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int c = 4 * (lane + 32 * i);
    float4 s = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c < n_keys) s = __ldcs(sr + lane + 32 * i);   // c + 3 < ld_in: the pad columns are readable
    v[i][0] = s.x, v[i][1] = s.y, v[i][2] = s.z, v[i][3] = s.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c + e >= n_keys || (mk && mk[c + e])) v[i][e] = -INFINITY;
      mx = fmaxf(mx, v[i][e]);
    }
  }
  mx = warp_max(mx);
  const float mref = (mx == -INFINITY) ? 0.f : mx;  // fully masked row -> NaN like the reference (0/0)
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[i][e] = __expf(v[i][e] - mref);
      sum += v[i][e];
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  uint2* pr = reinterpret_cast<uint2*>(probs + row * ld_out);
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c < ld_out) {
      uint2 q;
      q.x = OpTraits<OpT>::pack2(c < n_keys ? v[i][0] * inv : 0.f, c + 1 < n_keys ? v[i][1] * inv : 0.f);
      q.y = OpTraits<OpT>::pack2(c + 2 < n_keys ? v[i][2] * inv : 0.f, c + 3 < n_keys ? v[i][3] * inv : 0.f);
      pr[lane + 32 * i] = q;
    }
  }
}

template <int V4, typename OpT>
static void launch_softmax_vec(const float* scores, long long ld_in, long long rows, int n_keys, const uint8_t* key_mask,
                               int rows_per_seq, void* probs, long long ld_out, cudaStream_t s) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  launch_pdl(softmax_rows_vec_kernel<V4, OpT>, dim3(grid), dim3(256), 0, s, scores, ld_in, rows, n_keys, key_mask, rows_per_seq,
                                                        reinterpret_cast<OpT*>(probs), ld_out);
}
template <typename OpT>
static void dispatch_softmax_vec(const float* scores, long long ld_in, long long rows, int n_keys,
                                 const uint8_t* key_mask, int rows_per_seq, void* probs, long long ld_out,
                                 cudaStream_t s) {
  switch ((ld_out + 127) / 128) {
    case 1: launch_softmax_vec<1, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    case 2: launch_softmax_vec<2, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    case 3: launch_softmax_vec<3, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    case 4: launch_softmax_vec<4, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    case 5: launch_softmax_vec<5, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    case 6: launch_softmax_vec<6, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    case 7: launch_softmax_vec<7, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
    default: launch_softmax_vec<8, OpT>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s); break;
  }
}

template <typename OpT>
__global__ void __launch_bounds__(256) convert_kernel(const float* __restrict__ x, OpT* __restrict__ out,
                                                       long long n) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
    uint2 pk;
    pk.x = OpTraits<OpT>::pack2(v.x, v.y);
    pk.y = OpTraits<OpT>::pack2(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = pk;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = OpTraits<OpT>::cvt(x[n4 * 4 + threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------
// CMVN apply + zero padding.  One thread per 4 mel bins of one (utterance, output row).
// mean_std: [B, 2, 80] fp32 (mean, std) from cmvn_stats_kernel; y = (x - mean) / std with IEEE fp32
// subtract / divide, i.e. exactly numpy's np.subtract / np.divide in fairseq UtteranceCMVN.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int frames_of(long long n, int lengths_are_samples) {
  if (!lengths_are_samples) return (int)n;
  return n < 400 ? 0 : (int)(1 + (n - 400) / 160);
}

__device__ __forceinline__ float4 cmvn4(const float4 v, const float* s_mean, const float* s_std, int c4) {
  float4 y;
  y.x = __fdiv_rn(__fsub_rn(v.x, s_mean[4 * c4 + 0]), s_std[4 * c4 + 0]);
  y.y = __fdiv_rn(__fsub_rn(v.y, s_mean[4 * c4 + 1]), s_std[4 * c4 + 1]);
  y.z = __fdiv_rn(__fsub_rn(v.z, s_mean[4 * c4 + 2]), s_std[4 * c4 + 2]);
  y.w = __fdiv_rn(__fsub_rn(v.w, s_mean[4 * c4 + 3]), s_std[4 * c4 + 3]);
  return y;
}

// F32 / OP: which of the two outputs exist (compile-time, so the variant the encoder runs -- 16-bit conv operand only --
// does not carry the registers of the fp32 path: 4 blocks per SM instead of 3, i.e. two waves of blocks at the bench
// shape instead of three, each wave one memory round trip deep)
template <typename OpT, bool F32, bool OP>
__global__ void __launch_bounds__(256, (F32 && OP) ? 3 : 4) cmvn_apply_kernel(const float* __restrict__ feats,
                                                          const float* __restrict__ mean_std,
                                                          const long long* __restrict__ lens, int lengths_are_samples,
                                                          int max_frames, float* __restrict__ out_f32,
                                                          OpT* __restrict__ out_op, int op_frames, int op_row_offset,
                                                          int rows_per_block, const int* __restrict__ spec, int n_fmask,
                                                          int n_tmask, float mask_value) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  __shared__ float s_mean[80], s_std[80];
  // SpecAugment (fairseq SpecAugmentTransform, applied after CMVN): per utterance n_fmask (f0, f) frequency bands and
  // n_tmask (t0, t) frame ranges, drawn on the host, overwritten with mask_value while the features are written
  const int* sp = spec ? spec + (long long)blockIdx.y * 2 * (n_fmask + n_tmask) : nullptr;
  auto specaug = [&](float4 y, int frame, int c4) {
    for (int i = 0; i < n_tmask; ++i) {
      const int t0 = sp[2 * (n_fmask + i)], t = sp[2 * (n_fmask + i) + 1];
      if (frame >= t0 && frame < t0 + t) return make_float4(mask_value, mask_value, mask_value, mask_value);
    }
    for (int i = 0; i < n_fmask; ++i) {
      const int f0 = sp[2 * i], f1 = f0 + sp[2 * i + 1], c = 4 * c4;
      if (c >= f0 && c < f1) y.x = mask_value;
      if (c + 1 >= f0 && c + 1 < f1) y.y = mask_value;
      if (c + 2 >= f0 && c + 2 < f1) y.z = mask_value;
      if (c + 3 >= f0 && c + 3 < f1) y.w = mask_value;
    }
    return y;
  };
  const int b = blockIdx.y;
  const int nfr = min(frames_of(lens[b], lengths_are_samples), max_frames);
  const bool ident = mean_std == nullptr;   // input already normalised: copy / pad only
  // rows handled by this block: [row0, row0 + rows_per_block) of the LARGER of the two output extents
  const int total_rows = max(op_frames, max_frames);
  const int row0 = blockIdx.x * rows_per_block;
  // Batches of five (row, float4-column) items per thread.  All feature loads of a batch are issued before anything
  // else - the statistics are fetched while they are in flight - so the kernel is one memory round trip deep
  // (one load at a time had left it latency-bound at a third of the HBM rate).
  constexpr int NI = 5;
  const int items = rows_per_block * 20;
  bool stats_ready = false;
  for (int base = threadIdx.x; base < items; base += NI * blockDim.x) {
    float4 vf[NI], vo[NI];
    int rowi[NI];
#pragma unroll
    for (int u = 0; u < NI; ++u) {
      const int idx = base + u * blockDim.x;
      const int row = row0 + idx / 20, c4 = idx % 20;
      rowi[u] = (idx < items && row < total_rows) ? row : -1;
      vf[u] = vo[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowi[u] < 0) continue;
      // fp32 output row `row` <-> utterance frame `row`; operand row `row` <-> frame `row - op_row_offset`
      if (F32 && row < nfr)
        vf[u] = __ldg(reinterpret_cast<const float4*>(feats + ((long long)b * max_frames + row) * 80) + c4);
      const int fr = row - op_row_offset;
      if (OP && fr >= 0 && fr < nfr)
        vo[u] = __ldg(reinterpret_cast<const float4*>(feats + ((long long)b * max_frames + fr) * 80) + c4);
    }
    if (!stats_ready) {   // uniform: first trip only
      if (threadIdx.x < 80) {
        s_mean[threadIdx.x] = ident ? 0.f : mean_std[(long long)b * 160 + threadIdx.x];
        s_std[threadIdx.x] = ident ? 1.f : mean_std[(long long)b * 160 + 80 + threadIdx.x];
      }
      __syncthreads();
      stats_ready = true;
    }
#pragma unroll
    for (int u = 0; u < NI; ++u) {
      if (rowi[u] < 0) continue;
      const int idx = base + u * blockDim.x;
      const int row = rowi[u], c4 = idx % 20;
      if (F32 && row < max_frames) {
        float4 y = (row < nfr && !ident) ? cmvn4(vf[u], s_mean, s_std, c4) : vf[u];
        if (sp && row < nfr) y = specaug(y, row, c4);
        reinterpret_cast<float4*>(out_f32 + ((long long)b * max_frames + row) * 80)[c4] = y;
      }
      if (OP && row < op_frames) {
        const int fr = row - op_row_offset;
        float4 y = (fr >= 0 && fr < nfr && !ident) ? cmvn4(vo[u], s_mean, s_std, c4) : vo[u];
        if (sp && fr >= 0 && fr < nfr) y = specaug(y, fr, c4);
        uint2 pk;
        pk.x = OpTraits<OpT>::pack2(y.x, y.y);
        pk.y = OpTraits<OpT>::pack2(y.z, y.w);
        reinterpret_cast<uint2*>(out_op + ((long long)b * op_frames + row) * 80)[c4] = pk;
      }
    }
  }
}

__global__ void seq_lens_kernel(const long long* __restrict__ lens, int lengths_are_samples, int batch, int n_layers,
                                int* __restrict__ out) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int n = frames_of(lens[b], lengths_are_samples);
  for (int i = 0; i < n_layers; ++i) n = n <= 0 ? 0 : (n - 1) / 2 + 1;  // floor((L-1)/2 + 1)
  out[b] = n;
}

}  // namespace mm

using namespace mm;

extern "C" int mm_layernorm(const float* x, const float* gamma, const float* beta, int64_t rows, int32_t dim,
                            void* out_op, float* out_f32, int32_t dtype, float eps, void* stream) {
  if (!x || !gamma || !beta || (!out_op && !out_f32)) return bad_arg("layernorm: null pointer");
  if (rows <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == MM_DTYPE_F16 ? launch_ln<__half>(x, gamma, beta, rows, dim, out_op, out_f32, eps, s)
                               : launch_ln<__nv_bfloat16>(x, gamma, beta, rows, dim, out_op, out_f32, eps, s);
}

namespace mm {
// Decoder input embedding (fairseq TransformerDecoderBase.extract_features_scriptable): x = scale * E[token] +
// sinusoidal[position], position = padding_idx + (number of non-pad tokens up to and including this one) for real
// tokens and padding_idx (the all-zero table row) for pad tokens (fairseq utils.make_positions).
// One block per (position, sequence): the 128 threads first count the non-pad tokens of the prefix, then write the row.
__global__ void __launch_bounds__(128) embed_tokens_kernel(const long long* __restrict__ tokens, int padding_idx,
                                                           const float* __restrict__ table, int vocab, float scale,
                                                           const float* __restrict__ pos_table, int pos_rows, int L,
                                                           int dim, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int s_cnt[4];
  const int t = blockIdx.x, b = blockIdx.y;
  const long long* row = tokens + (long long)b * L;
  int cnt = 0;
  for (int i = threadIdx.x; i <= t; i += 128) cnt += row[i] != padding_idx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  const long long tok = row[t];
  const int total = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
  int pos = tok != padding_idx ? padding_idx + total : padding_idx;
  pos = min(pos, pos_rows - 1);
  const long long tk = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const float4* e = reinterpret_cast<const float4*>(table + tk * dim);
  const float4* pe = reinterpret_cast<const float4*>(pos_table + (long long)pos * dim);
  float4* o4 = reinterpret_cast<float4*>(out + ((long long)b * L + t) * dim);
  for (int i = threadIdx.x; i < dim / 4; i += 128) {
    const float4 a = __ldg(e + i), p = __ldg(pe + i);
    o4[i] = make_float4(fmaf(scale, a.x, p.x), fmaf(scale, a.y, p.y), fmaf(scale, a.z, p.z), fmaf(scale, a.w, p.w));
  }
}
}  // namespace mm

extern "C" int mm_embed_tokens(const int64_t* tokens, int32_t padding_idx, const float* table, int32_t vocab, float scale,
                               const float* pos_table, int32_t pos_rows, int32_t batch, int32_t length, int32_t dim,
                               float* out, void* stream) {
  if (!tokens || !table || !pos_table || !out) return bad_arg("embed_tokens: null pointer");
  if (dim % 4 || vocab <= 0 || pos_rows <= padding_idx) return bad_arg("embed_tokens: dims");
  if (batch <= 0 || length <= 0) return 0;
  mm::launch_pdl(mm::embed_tokens_kernel, dim3(length, batch), dim3(128), 0, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const long long*>(tokens), padding_idx, table, vocab, scale, pos_table, pos_rows,
                 length, dim, out);
  MM_CHECK_LAUNCH("embed_tokens_kernel launch");
  return 0;
}

namespace mm {
// Label-smoothed cross entropy of the unit logits (fairseq label_smoothed_nll_loss behind the reference's criterion,
// mm_s2ut/criterions/speech_to_speech_criterion.py:58-72 -> RdropLabelSmoothedCrossEntropyCriterion.compute_loss):
// per row  lprobs = log_softmax(logits) in fp32,  nll = -lprobs[target],  smooth = -sum_v lprobs[v];  rows whose target
// is the padding index contribute 0.  One warp per row (V <= a few thousand), two-pass log-sum-exp from registers.
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, long long ld, int vocab,
                                                      const long long* __restrict__ target, int padding_idx,
                                                      long long rows, float* __restrict__ row_nll,
                                                      float* __restrict__ row_smooth) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* x = logits + row * ld;
  const long long tgt = target[row];
  float mx = -INFINITY, sum = 0.f;
  for (int c = lane; c < vocab; c += 32) {
    const float v = x[c];
    mx = fmaxf(mx, v);
    sum += v;
  }
  mx = warp_max(mx);
  sum = warp_sum(sum);
  float se = 0.f;
  for (int c = lane; c < vocab; c += 32) se += __expf(x[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  if (lane == 0) {
    const bool pad = tgt == padding_idx || tgt < 0 || tgt >= vocab;
    row_nll[row] = pad ? 0.f : lse - x[tgt];
    row_smooth[row] = pad ? 0.f : (float)vocab * lse - sum;
  }
}
// deterministic sum of the per-row terms (fixed tree order) -> out[0] = sum nll, out[1] = sum smooth
__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float* __restrict__ row_nll,
                                                         const float* __restrict__ row_smooth, long long rows,
                                                         float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double s_a[1024], s_b[1024];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 1024) a += row_nll[i], b += row_smooth[i];
  s_a[threadIdx.x] = a, s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_a[threadIdx.x] += s_a[threadIdx.x + o], s_b[threadIdx.x] += s_b[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)s_a[0], out[1] = (float)s_b[0];
}
}  // namespace mm

extern "C" int mm_label_smoothed_nll(const float* logits, int64_t ld, int32_t vocab, const int64_t* target,
                                     int32_t padding_idx, int64_t rows, float* row_nll, float* row_smooth,
                                     float* sums, void* stream) {
  if (!logits || !target || !row_nll || !row_smooth || !sums) return bad_arg("label_smoothed_nll: null pointer");
  if (vocab <= 0 || ld < vocab) return bad_arg("label_smoothed_nll: vocab / ld");
  if (rows <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  mm::launch_pdl(mm::ce_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, logits, (long long)ld, vocab,
                 reinterpret_cast<const long long*>(target), padding_idx, (long long)rows, row_nll, row_smooth);
  MM_CHECK_LAUNCH("ce_rows_kernel launch");
  mm::launch_pdl(mm::ce_reduce_kernel, dim3(1), dim3(1024), 0, s, (const float*)row_nll, (const float*)row_smooth,
                 (long long)rows, sums);
  MM_CHECK_LAUNCH("ce_reduce_kernel launch");
  return 0;
}

extern "C" int mm_layernorm_gather(const void* x, int32_t x_dtype, const int64_t* index, int32_t rows_per_index,
                                   const float* gamma, const float* beta, int64_t rows, int32_t dim, void* out_op,
                                   int32_t dtype, float eps, void* stream) {
  if (!x || !gamma || !beta || !out_op) return bad_arg("layernorm_gather: null pointer");
  if (index && rows_per_index <= 0) return bad_arg("layernorm_gather: rows_per_index");
  if (x_dtype != MM_DTYPE_F16 && x_dtype != MM_DTYPE_BF16) return bad_arg("layernorm_gather: 16-bit input only");
  if (rows <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long* idx = reinterpret_cast<const long long*>(index);
  const bool in16 = x_dtype == MM_DTYPE_F16, out16 = dtype == MM_DTYPE_F16;
  if (in16)
    return out16 ? launch_ln_gather<__half, __half>(x, idx, rows_per_index, gamma, beta, rows, dim, out_op, eps, s)
                 : launch_ln_gather<__half, __nv_bfloat16>(x, idx, rows_per_index, gamma, beta, rows, dim, out_op, eps, s);
  return out16 ? launch_ln_gather<__nv_bfloat16, __half>(x, idx, rows_per_index, gamma, beta, rows, dim, out_op, eps, s)
               : launch_ln_gather<__nv_bfloat16, __nv_bfloat16>(x, idx, rows_per_index, gamma, beta, rows, dim, out_op, eps, s);
}

extern "C" int mm_softmax_rows(const float* scores, int64_t ld_in, int64_t rows, int32_t n_keys,
                               const uint8_t* key_mask, int32_t rows_per_seq, void* probs, int64_t ld_out,
                               int32_t dtype, void* stream) {
  if (!scores || !probs) return bad_arg("softmax: null pointer");
  if (n_keys <= 0 || n_keys > 1024 || ld_out > 1024 || ld_out < n_keys) return bad_arg("softmax: 1 <= n_keys <= ld_out <= 1024");
  if (key_mask && rows_per_seq <= 0) return bad_arg("softmax: key_mask needs rows_per_seq");
  if (rows <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (ld_in % 4 == 0 && ld_out % 4 == 0 && ld_in >= ((n_keys + 3) & ~3) &&
      (reinterpret_cast<uintptr_t>(scores) & 15) == 0 && (reinterpret_cast<uintptr_t>(probs) & 7) == 0) {
    if (dtype == MM_DTYPE_F16)
      mm::dispatch_softmax_vec<__half>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s);
    else
      mm::dispatch_softmax_vec<__nv_bfloat16>(scores, ld_in, rows, n_keys, key_mask, rows_per_seq, probs, ld_out, s);
    MM_CHECK_LAUNCH("softmax_rows_vec_kernel launch");
    return 0;
  }
  if (dtype == MM_DTYPE_F16)
    launch_pdl(softmax_rows_kernel<__half>, dim3(grid), dim3(256), 0, s, scores, ld_in, rows, n_keys, key_mask, rows_per_seq,
                                                     reinterpret_cast<__half*>(probs), ld_out);
  else
    launch_pdl(softmax_rows_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, scores, ld_in, rows, n_keys, key_mask, rows_per_seq,
                                                            reinterpret_cast<__nv_bfloat16*>(probs), ld_out);
  MM_CHECK_LAUNCH("softmax_rows_kernel launch");
  return 0;
}

// scores[r, k] = -inf where key_mask[r / rows_per_seq][k] != 0: one warp per row, writes only the masked cells
namespace mm {
__global__ void __launch_bounds__(256) mask_scores_kernel(float* __restrict__ scores, long long ld, long long rows, int n_keys,
                                                          const uint8_t* __restrict__ key_mask, long long mask_ld,
                                                          int rows_per_seq) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint8_t* mk = key_mask + (row / rows_per_seq) * mask_ld;
  float* srow = scores + row * ld;
  for (int k = threadIdx.x & 31; k < n_keys; k += 32)
    if (mk[k]) srow[k] = -INFINITY;
}
}  // namespace mm

extern "C" int mm_mask_scores(float* scores, int64_t ld, int64_t rows, int32_t n_keys, const uint8_t* key_mask,
                              int64_t mask_ld, int32_t rows_per_seq, void* stream) {
  if (!scores || !key_mask) return bad_arg("mask_scores: null pointer");
  if (n_keys <= 0 || ld < n_keys || mask_ld < n_keys || rows_per_seq <= 0) return bad_arg("mask_scores: extents");
  if (rows <= 0) return 0;
  launch_pdl(mm::mask_scores_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             scores, (long long)ld, (long long)rows, n_keys, key_mask, (long long)mask_ld, rows_per_seq);
  MM_CHECK_LAUNCH("mask_scores_kernel launch");
  return 0;
}

extern "C" int mm_convert_f32(const float* x, void* out, int64_t n, int32_t dtype, void* stream) {
  if (!x || !out) return bad_arg("convert: null pointer");
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) return bad_arg("convert: alignment");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (dtype == MM_DTYPE_F16)
    launch_pdl(convert_kernel<__half>, dim3((unsigned)blocks), dim3(256), 0, s, x, reinterpret_cast<__half*>(out), n);
  else
    launch_pdl(convert_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), 0, s, x, reinterpret_cast<__nv_bfloat16*>(out), n);
  MM_CHECK_LAUNCH("convert_kernel launch");
  return 0;
}

extern "C" int mm_cmvn_apply(const float* feats, const float* mean_std, const int64_t* lens,
                             int32_t lengths_are_samples, int32_t batch, int32_t max_frames, float* out_f32,
                             void* out_op, int32_t op_frames, int32_t op_row_offset, int32_t dtype, void* stream) {
  return mm_cmvn_apply_specaug(feats, mean_std, lens, lengths_are_samples, batch, max_frames, out_f32, out_op, op_frames,
                               op_row_offset, dtype, nullptr, 0, 0, 0.f, stream);
}

extern "C" int mm_cmvn_apply_specaug(const float* feats, const float* mean_std, const int64_t* lens,
                                     int32_t lengths_are_samples, int32_t batch, int32_t max_frames, float* out_f32,
                                     void* out_op, int32_t op_frames, int32_t op_row_offset, int32_t dtype,
                                     const int32_t* spec_masks, int32_t n_fmask, int32_t n_tmask, float mask_value,
                                     void* stream) {
  if (!feats || !lens || (!out_f32 && !out_op)) return bad_arg("cmvn: null pointer");
  if (spec_masks && (n_fmask < 0 || n_tmask < 0 || n_fmask + n_tmask <= 0)) return bad_arg("cmvn: specaugment mask counts");
  if (!spec_masks) n_fmask = n_tmask = 0;
  if (batch <= 0 || max_frames <= 0) return 0;
  if (!out_op) op_frames = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows_per_block = 64;   // 64 rows x 20 float4 = 1280 items: one batch of five per thread, 1024 blocks = one wave
  const int total_rows = op_frames > max_frames ? op_frames : max_frames;
  dim3 grid((total_rows + rows_per_block - 1) / rows_per_block, batch);
  const long long* l = reinterpret_cast<const long long*>(lens);
  auto go = [&](auto kern, auto* op) {
    launch_pdl(kern, dim3(grid), dim3(256), 0, s, feats, mean_std, l, lengths_are_samples, max_frames, out_f32, op, op_frames,
               op_row_offset, rows_per_block, spec_masks, n_fmask, n_tmask, mask_value);
  };
  const bool f32 = out_f32 != nullptr, op = out_op != nullptr;
  if (dtype == MM_DTYPE_F16) {
    __half* o = reinterpret_cast<__half*>(out_op);
    if (f32 && op) go(cmvn_apply_kernel<__half, true, true>, o);
    else if (op) go(cmvn_apply_kernel<__half, false, true>, o);
    else go(cmvn_apply_kernel<__half, true, false>, o);
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_op);
    if (f32 && op) go(cmvn_apply_kernel<__nv_bfloat16, true, true>, o);
    else if (op) go(cmvn_apply_kernel<__nv_bfloat16, false, true>, o);
    else go(cmvn_apply_kernel<__nv_bfloat16, true, false>, o);
  }
  MM_CHECK_LAUNCH("cmvn_apply_kernel launch");
  return 0;
}

namespace mm {
// encoder_padding_mask[b, t] = t >= seq_lens[b]  (fairseq lengths_to_padding_mask on the subsampled lengths)
__global__ void __launch_bounds__(256) padding_mask_kernel(const int* __restrict__ seq_lens, int T, long long total,
                                                           uint8_t* __restrict__ mask) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  mask[i] = (int)(i % T) >= seq_lens[i / T];
}
}  // namespace mm

extern "C" int mm_padding_mask(const int32_t* seq_lens, int32_t batch, int32_t T, uint8_t* mask, void* stream) {
  if (!seq_lens || !mask) return bad_arg("padding_mask: null pointer");
  if (batch <= 0 || T <= 0) return 0;
  const long long total = (long long)batch * T;
  mm::launch_pdl(mm::padding_mask_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                 static_cast<cudaStream_t>(stream), seq_lens, T, total, mask);
  MM_CHECK_LAUNCH("padding_mask_kernel launch");
  return 0;
}

// seq_lens + padding mask in one launch: block b computes its utterance's subsampled length and writes mask row b
namespace mm {
__global__ void __launch_bounds__(128) seq_lens_mask_kernel(const long long* __restrict__ lens, int lengths_are_samples,
                                                            int n_layers, int T, int* __restrict__ out,
                                                            uint8_t* __restrict__ mask) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  int n = frames_of(lens[b], lengths_are_samples);
  for (int i = 0; i < n_layers; ++i) n = n <= 0 ? 0 : (n - 1) / 2 + 1;
  if (threadIdx.x == 0) out[b] = n;
  for (int t = threadIdx.x; t < T; t += 128) mask[(long long)b * T + t] = t >= n;
}
}  // namespace mm

extern "C" int mm_seq_lens_mask(const int64_t* lens, int32_t lengths_are_samples, int32_t batch, int32_t n_layers, int32_t T,
                                int32_t* out_lens, uint8_t* mask, void* stream) {
  if (!lens || !out_lens || !mask) return bad_arg("seq_lens_mask: null pointer");
  if (batch <= 0 || T <= 0) return 0;
  mm::launch_pdl(mm::seq_lens_mask_kernel, dim3(batch), dim3(128), 0, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const long long*>(lens), lengths_are_samples, n_layers, T, out_lens, mask);
  MM_CHECK_LAUNCH("seq_lens_mask_kernel launch");
  return 0;
}

extern "C" int mm_seq_lens(const int64_t* lens, int32_t lengths_are_samples, int32_t batch, int32_t n_layers,
                           int32_t* out_lens, void* stream) {
  if (!lens || !out_lens) return bad_arg("seq_lens: null pointer");
  if (batch <= 0) return 0;
  launch_pdl(seq_lens_kernel, dim3((batch + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(lens), lengths_are_samples, batch, n_layers, out_lens);
  MM_CHECK_LAUNCH("seq_lens_kernel launch");
  return 0;
}
