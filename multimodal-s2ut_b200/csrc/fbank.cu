// 80-bin Kaldi-compatible log-mel filterbank (torchaudio.compliance.kaldi.fbank defaults: 25 ms povey
// window, 10 ms shift, snip_edges, DC removal, pre-emphasis 0.97, 512-point FFT, 20 Hz .. Nyquist mel bank,
// log floor eps) + per-chunk CMVN statistics, fp32 throughout.
//
// One CTA = 32 consecutive frames of one utterance.  The 5360 samples they span are staged once into shared
// memory with 128-bit coalesced loads (frames overlap 2.5x, HBM sees each sample ~once).  Each HALF-warp
// owns a frame: the 512-point real FFT runs as a 256-point complex FFT (16 x 16 Cooley-Tukey, one
// 16-point DFT per lane entirely in registers, a single padded shared-memory transpose between the two
// passes), followed by the real-FFT split, |X|^2, the sparse triangular mel projection (<= 2 filters per
// FFT bin) and log.  Per-(utterance, chunk) sums / sums of squares for utterance CMVN are reduced in
// registers -> shared memory -> one deterministic store (no atomics).
#include <math.h>

#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_BINS = 80;
constexpr int FB_FRAMES_PER_CTA = 32;
constexpr int FB_WAVE = FB_WIN + (FB_FRAMES_PER_CTA - 1) * FB_SHIFT;  // 5360 samples
constexpr int FB_MAX_NNZ = 1024;
// table layout (floats)
constexpr int TB_WIN = 0;                         // [400]
constexpr int TB_TW256 = TB_WIN + 400;            // [256][2]  exp(-2 pi i m / 256)
constexpr int TB_TW512 = TB_TW256 + 512;          // [256][2]  exp(-2 pi i k / 512)
constexpr int TB_MELW = TB_TW512 + 512;           // [FB_MAX_NNZ]
constexpr int TB_K0 = TB_MELW + FB_MAX_NNZ;       // [80] int: first FFT bin of filter m
constexpr int TB_CNT = TB_K0 + 80;                // [80] int: number of bins
constexpr int TB_OFF = TB_CNT + 80;               // [80] int: offset into MELW
constexpr int TB_TOTAL = TB_OFF + 80;             // 2688 floats

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// forward 4-point DFT (W4 = -i), natural order in/out
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
  const float2 s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));
  a0 = cadd(s02, s13);
  a1 = cadd(d02, d13);
  a2 = csub(s02, s13);
  a3 = csub(d02, d13);
}

// forward 16-point DFT in registers: n = n1 + 4 n2, k = 4 k1 + k2
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508978f, R2 = 0.70710678118654752f;
  // W16^m = (cos, -sin)(2 pi m / 16)
  const float2 W[10] = {{1.f, 0.f},  {C1, -S1}, {R2, -R2}, {S1, -C1}, {0.f, -1.f},
                        {-S1, -C1}, {-R2, -R2}, {-C1, -S1}, {-1.f, 0.f}, {-C1, S1}};
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    dft4(a[n1], a[n1 + 4], a[n1 + 8], a[n1 + 12]);  // over n2 -> k2 at slot n1 + 4 k2
#pragma unroll
    for (int k2 = 1; k2 < 4; ++k2)
      if (n1 * k2 != 0) a[n1 + 4 * k2] = cmul(a[n1 + 4 * k2], W[n1 * k2]);
  }
  float2 t[16];
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    float2 b0 = a[0 + 4 * k2], b1 = a[1 + 4 * k2], b2 = a[2 + 4 * k2], b3 = a[3 + 4 * k2];
    dft4(b0, b1, b2, b3);  // over n1 -> k1
    t[k2] = b0, t[4 + k2] = b1, t[8 + k2] = b2, t[12 + k2] = b3;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = t[i];
}

constexpr int FB_THREADS = 256;
constexpr int FB_HW = FB_THREADS / 16;                 // 16 half-warps
constexpr int FB_XBUF = 16 * 17 * 2;                   // floats per half-warp: padded 16x17 complex transpose / spectrum
constexpr int FB_SMEM_FLOATS = FB_WAVE + FB_HW * FB_XBUF;
constexpr int FB_SMEM_BYTES = FB_SMEM_FLOATS * 4;      // 56 KB -> 3-4 CTAs per SM

// WavT = float (samples already in int16 range) or int16_t (raw PCM: halves the dominant HBM / PCIe read; the
// int16 -> fp32 conversion is exact, so both inputs give bit-identical features for integer-valued audio)
template <typename WavT>
__global__ void __launch_bounds__(FB_THREADS, 3) fbank_kernel(const WavT* __restrict__ wav,
                                                              const long long* __restrict__ n_samples,
                                                              long long wav_stride, float* __restrict__ feats,
                                                              int max_frames, const float* __restrict__ tables) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  extern __shared__ __align__(16) float fsm[];
  float* s_wave = fsm;
  float* s_work = s_wave + FB_WAVE;

  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const long long n = n_samples[b];
  int nfr = n < FB_WIN ? 0 : (int)(1 + (n - FB_WIN) / FB_SHIFT);
  nfr = min(nfr, max_frames);
  const int f0 = chunk * FB_FRAMES_PER_CTA;
  if (f0 >= nfr) return;  // whole CTA out of range (uniform)
  const int nf_here = min(FB_FRAMES_PER_CTA, nfr - f0);

  // ---- stage the samples of this CTA's 32 frames (5360 values; frames overlap 2.5x) ----
  const long long s0 = (long long)f0 * FB_SHIFT;
  const int n_need = FB_WIN + (nf_here - 1) * FB_SHIFT;  // <= n - s0 by construction
  const WavT* wsrc = wav + (long long)b * wav_stride + s0;
  if constexpr (sizeof(WavT) == 4) {
    if ((reinterpret_cast<uintptr_t>(wsrc) & 15) == 0) {
      for (int i = threadIdx.x * 4; i < n_need; i += FB_THREADS * 4) {
        if (i + 3 < n_need) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(wsrc + i));
          *reinterpret_cast<float4*>(s_wave + i) = v;
        } else {
          for (int j = i; j < n_need; ++j) s_wave[j] = wsrc[j];
        }
      }
    } else {
      for (int i = threadIdx.x; i < n_need; i += FB_THREADS) s_wave[i] = wsrc[i];
    }
  } else {
    if ((reinterpret_cast<uintptr_t>(wsrc) & 15) == 0) {     // 8 PCM samples per 128-bit load
      for (int i = threadIdx.x * 8; i < n_need; i += FB_THREADS * 8) {
        if (i + 7 < n_need) {
          const uint4 q = __ldcs(reinterpret_cast<const uint4*>(wsrc + i));
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s_wave[i + 2 * j] = (float)(short)(w[j] & 0xFFFFu);
            s_wave[i + 2 * j + 1] = (float)(short)(w[j] >> 16);
          }
        } else {
          for (int j = i; j < n_need; ++j) s_wave[j] = (float)wsrc[j];
        }
      }
    } else {
      for (int i = threadIdx.x; i < n_need; i += FB_THREADS) s_wave[i] = (float)wsrc[i];
    }
  }
  __syncthreads();

  // constant tables stay in global memory (10.7 KB, L1/L2 resident) and are read through the read-only path
  const float2* g_win2 = reinterpret_cast<const float2*>(tables + TB_WIN);
  const float2* g_tw256 = reinterpret_cast<const float2*>(tables + TB_TW256);
  const float2* g_tw512 = reinterpret_cast<const float2*>(tables + TB_TW512);
  const float* g_melw = tables + TB_MELW;
  const int* g_k0 = reinterpret_cast<const int*>(tables + TB_K0);
  const int* g_cnt = reinterpret_cast<const int*>(tables + TB_CNT);
  const int* g_off = reinterpret_cast<const int*>(tables + TB_OFF);

  const int hw = threadIdx.x >> 4;       // half-warp id 0..15
  const int l = threadIdx.x & 15;        // lane in half-warp
  const int hsel = (threadIdx.x >> 4) & 1;
  const unsigned hmask = 0xFFFFu << (16 * hsel);
  float* xbuf = s_work + hw * FB_XBUF;
  float2* x2 = reinterpret_cast<float2*>(xbuf);
  // mel filters owned by this lane (l + 16 j): start bin, length, weight offset
  int mk0[5], mcnt[5], moff[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    mk0[j] = __ldg(g_k0 + l + 16 * j);
    mcnt[j] = __ldg(g_cnt + l + 16 * j);
    moff[j] = __ldg(g_off + l + 16 * j);
  }

  for (int fl = hw; fl < nf_here; fl += FB_HW) {
    const float* x = s_wave + fl * FB_SHIFT;
    // ---- DC removal: mean over the 400 samples (lanes stride 16; second half-warp rotated by 16 banks) ----
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 25; ++j) {
      int i = l + 16 * j + 16 * hsel;
      i = i >= FB_WIN ? i - FB_WIN : i;
      s += x[i];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(hmask, s, o);
    const float mean = s * (1.0f / FB_WIN);
    // ---- pass 1 input straight from the sample tile: z[n] = (y[2n], y[2n+1]), n = l + 16 n2, with
    //      y[i] = ((x[i]-mean) - 0.97 (x[max(i-1,0)]-mean)) * povey[i] for i < 400, else 0 ----
    float2 a[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      const int nn = l + 16 * n2;              // complex index; samples 2 nn, 2 nn + 1
      if (nn < FB_WIN / 2) {
        const float2 xv = *reinterpret_cast<const float2*>(x + 2 * nn);
        const float xm = x[nn > 0 ? 2 * nn - 1 : 0];
        const float2 w = __ldg(g_win2 + nn);
        const float c0 = xv.x - mean, c1 = xv.y - mean, cm = xm - mean;
        a[n2].x = __fmul_rn(__fsub_rn(c0, __fmul_rn(0.97f, cm)), w.x);
        a[n2].y = __fmul_rn(__fsub_rn(c1, __fmul_rn(0.97f, c0)), w.y);
      } else {
        a[n2] = make_float2(0.f, 0.f);
      }
    }
    // ---- pass 1: lane n1 = l, 16-point DFT over n2; twiddle W256^(n1 k2); padded transpose through smem ----
    dft16(a);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float2 v = (k2 == 0) ? a[0] : cmul(a[k2], __ldg(g_tw256 + l * k2));
      x2[k2 * 17 + l] = v;
    }
    __syncwarp(hmask);
    // ---- pass 2: lane k2 = l, 16-point DFT over n1 -> a[k1] = Z[16 k1 + l] ----
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) a[n1] = x2[l * 17 + n1];
    dft16(a);
    __syncwarp(hmask);   // all lanes have read the transpose: xbuf is reused for the power spectrum
    // ---- real-FFT split: X[k] needs Z[k] and conj(Z[256-k]); 256-k = 16 (15-k1) + (16-l) lives in lane 16-l
    //      (lane 0: own slot (16-k1) & 15) -> one shuffle pair per bin instead of a trip through smem ----
    const int src = ((16 - l) & 15) + 16 * hsel;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float px = __shfl_sync(hmask, a[15 - k1].x, src);
      const float py = __shfl_sync(hmask, a[15 - k1].y, src);
      const float2 zk = a[k1];
      float2 zm;
      zm.x = (l == 0) ? a[(16 - k1) & 15].x : px;
      zm.y = -((l == 0) ? a[(16 - k1) & 15].y : py);
      const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
      const float2 d = csub(zk, zm);
      const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // -i/2 * (zk - zm)
      const float2 xk = cadd(e, cmul(__ldg(g_tw512 + 16 * k1 + l), o));
      xbuf[16 * k1 + l] = xk.x * xk.x + xk.y * xk.y;
    }
    __syncwarp(hmask);
    // ---- sparse mel projection + log ----
    float* orow = feats + ((long long)b * max_frames + (f0 + fl)) * FB_BINS;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float e = 0.f;
      for (int i = 0; i < mcnt[j]; ++i) e = fmaf(__ldg(g_melw + moff[j] + i), xbuf[mk0[j] + i], e);
      orow[l + 16 * j] = logf(fmaxf(e, 1.1920928955078125e-07f));
    }
    __syncwarp(hmask);
  }
}

// Utterance CMVN statistics with the REFERENCE's arithmetic: fairseq UtteranceCMVN runs numpy fp32
// `x.mean(0)` and `(x**2).sum(0)`, which accumulate sequentially over frames in fp32; the raw-moment
// variance E[x^2]-mean^2 cancels catastrophically, so its rounding noise (up to a few 1e-2 after
// normalisation in low-variance bins) is part of the reference's result.  One thread per (utterance, bin)
// replays that exact order (loads are batched 8 deep and independent; only the adds are a chain).
constexpr int CS_THREADS = 256;
constexpr int CS_TILE = 64;                              // frames per shared-memory tile (20 KB), double buffered
__global__ void __launch_bounds__(CS_THREADS) cmvn_stats_kernel(const float* __restrict__ feats,
                                                                const long long* __restrict__ lens,
                                                                int lengths_are_samples, int max_frames,
                                                                float* __restrict__ mean_std) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  pdl_wait();
  // All 256 threads stream [64 frames x 80 bins] tiles into shared memory with 16-byte cp.async (coalesced, next
  // tile in flight while the current one is consumed); threads 0..79 then replay numpy's frame-sequential fp32
  // accumulation from shared memory, so the only serial chain is two dependent adds per frame.
  __shared__ __align__(16) float tile[2][CS_TILE * FB_BINS];
  const int b = blockIdx.x, bin = threadIdx.x;
  const long long n = lens[b];
  int nfr = lengths_are_samples ? (n < FB_WIN ? 0 : (int)(1 + (n - FB_WIN) / FB_SHIFT)) : (int)n;
  nfr = min(nfr, max_frames);
  const float* base = feats + (long long)b * max_frames * FB_BINS;
  const int n_tiles = (nfr + CS_TILE - 1) / CS_TILE;
  auto issue = [&](int t) {
    const int rows = min(CS_TILE, nfr - t * CS_TILE);
    const float4* src = reinterpret_cast<const float4*>(base + (long long)t * CS_TILE * FB_BINS);
    const uint32_t dst = smem_u32(&tile[t & 1][0]);
    for (int i = threadIdx.x; i < rows * (FB_BINS / 4); i += CS_THREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float s = 0.f, q = 0.f;
  if (n_tiles > 0) issue(0);
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) {
      issue(t + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (bin < FB_BINS) {
      const int rows = min(CS_TILE, nfr - t * CS_TILE);
      const float* x = &tile[t & 1][bin];
      int i = 0;
      for (; i + 8 <= rows; i += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = x[(i + j) * FB_BINS];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s = __fadd_rn(s, v[j]);
          q = __fadd_rn(q, __fmul_rn(v[j], v[j]));
        }
      }
      for (; i < rows; ++i) {
        const float v = x[i * FB_BINS];
        s = __fadd_rn(s, v);
        q = __fadd_rn(q, __fmul_rn(v, v));
      }
    }
    __syncthreads();   // the tile may be overwritten by the load issued two iterations later
  }
  if (bin >= FB_BINS) return;
  float mean = 0.f, sd = 1.f;
  if (nfr > 0) {
    const float T = (float)nfr;
    mean = __fdiv_rn(s, T);
    const float var = __fsub_rn(__fdiv_rn(q, T), __fmul_rn(mean, mean));
    sd = __fsqrt_rn(fmaxf(var, 1e-10f));
  }
  mean_std[(long long)b * 160 + bin] = mean;
  mean_std[(long long)b * 160 + 80 + bin] = sd;
}

}  // namespace mm

using namespace mm;

extern "C" int mm_fbank_table_floats(void) { return TB_TOTAL; }

// Host-side table construction.  The mel bank repeats torchaudio's fp32 arithmetic (get_mel_banks) so the
// filter weights agree with the reference to the last bits; window and twiddles are rounded from double.
extern "C" int mm_fbank_build_tables(float* out) {
  if (!out) return bad_arg("fbank tables: null");
  memset(out, 0, sizeof(float) * TB_TOTAL);
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < FB_WIN; ++i) {
    const float hann = (float)(0.5 - 0.5 * cos(2.0 * PI * i / (FB_WIN - 1)));
    out[TB_WIN + i] = powf(hann, 0.85f);
  }
  for (int m = 0; m < 256; ++m) {
    out[TB_TW256 + 2 * m] = (float)cos(2.0 * PI * m / 256.0);
    out[TB_TW256 + 2 * m + 1] = (float)(-sin(2.0 * PI * m / 256.0));
    out[TB_TW512 + 2 * m] = (float)cos(2.0 * PI * m / 512.0);
    out[TB_TW512 + 2 * m + 1] = (float)(-sin(2.0 * PI * m / 512.0));
  }
  const double sample_freq = 16000.0, low_freq = 20.0, high_freq = 8000.0;
  const double fft_bin_width = sample_freq / FB_NFFT;
  const double mel_low = 1127.0 * log(1.0 + low_freq / 700.0);
  const double mel_high = 1127.0 * log(1.0 + high_freq / 700.0);
  const double delta = (mel_high - mel_low) / (FB_BINS + 1);
  int* k0 = reinterpret_cast<int*>(out + TB_K0);
  int* cnt = reinterpret_cast<int*>(out + TB_CNT);
  int* off = reinterpret_cast<int*>(out + TB_OFF);
  int nnz = 0;
  for (int m = 0; m < FB_BINS; ++m) {
    // torch: python-float scalars are applied in the tensor dtype (fp32)
    const float left = (float)mel_low + (float)m * (float)delta;
    const float center = (float)mel_low + ((float)m + 1.0f) * (float)delta;
    const float right = (float)mel_low + ((float)m + 2.0f) * (float)delta;
    int first = -1, last = -1;
    float w[256];
    for (int k = 0; k < 256; ++k) {
      const float f = (float)fft_bin_width * (float)k;
      const float mel = 1127.0f * logf(1.0f + f / 700.0f);
      const float up = (mel - left) / (center - left);
      const float down = (right - mel) / (right - center);
      const float v = fmaxf(0.0f, fminf(up, down));
      w[k] = v;
      if (v > 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    k0[m] = first < 0 ? 0 : first;
    cnt[m] = first < 0 ? 0 : last - first + 1;
    off[m] = nnz;
    if (nnz + cnt[m] > FB_MAX_NNZ) return bad_arg("fbank tables: mel bank too dense");
    for (int i = 0; i < cnt[m]; ++i) out[TB_MELW + nnz + i] = w[first + i];
    nnz += cnt[m];
  }
  return 0;
}

template <typename WavT>
static int launch_fbank(const WavT* wav, const int64_t* n_samples, int32_t batch, int64_t wav_stride, float* feats,
                        int32_t max_frames, const float* tables, void* stream) {
  if (!wav || !n_samples || !feats || !tables) return bad_arg("fbank: null pointer");
  if (batch <= 0 || max_frames <= 0) return 0;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(fbank_kernel<WavT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(fbank)");
    attr_set = true;
  }
  const int n_chunks = (max_frames + FB_FRAMES_PER_CTA - 1) / FB_FRAMES_PER_CTA;
  dim3 grid(n_chunks, batch);
  launch_pdl(fbank_kernel<WavT>, dim3(grid), dim3(FB_THREADS), FB_SMEM_BYTES, static_cast<cudaStream_t>(stream), 
      wav, reinterpret_cast<const long long*>(n_samples), wav_stride, feats, max_frames, tables);
  MM_CHECK_LAUNCH("fbank_kernel launch");
  return 0;
}

extern "C" int mm_fbank_f32(const float* wav, const int64_t* n_samples, int32_t batch, int64_t wav_stride, float* feats,
                            int32_t max_frames, const float* tables, void* stream) {
  return launch_fbank<float>(wav, n_samples, batch, wav_stride, feats, max_frames, tables, stream);
}

extern "C" int mm_fbank_i16(const int16_t* wav, const int64_t* n_samples, int32_t batch, int64_t wav_stride,
                            float* feats, int32_t max_frames, const float* tables, void* stream) {
  return launch_fbank<int16_t>(wav, n_samples, batch, wav_stride, feats, max_frames, tables, stream);
}

extern "C" int mm_cmvn_stats(const float* feats, const int64_t* lens, int32_t lengths_are_samples, int32_t batch,
                             int32_t max_frames, float* mean_std, void* stream) {
  if (!feats || !lens || !mean_std) return bad_arg("cmvn_stats: null pointer");
  if (batch <= 0 || max_frames <= 0) return 0;
  launch_pdl(cmvn_stats_kernel, dim3(batch), dim3(CS_THREADS), 0, static_cast<cudaStream_t>(stream), 
      feats, reinterpret_cast<const long long*>(lens), lengths_are_samples, max_frames, mean_std);
  MM_CHECK_LAUNCH("cmvn_stats_kernel launch");
  return 0;
}
