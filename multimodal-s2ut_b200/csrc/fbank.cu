// 80-bin Kaldi-compatible log-mel filterbank (torchaudio.compliance.kaldi.fbank defaults: 25 ms povey
// window, 10 ms shift, snip_edges, DC removal, pre-emphasis 0.97, 512-point FFT, 20 Hz .. Nyquist mel bank,
// log floor eps) and the utterance CMVN statistics, fp32 throughout.
//
// One work item = 24 consecutive frames of one utterance.  The 4080 samples they span are staged once into shared
// memory with 128-bit coalesced loads (frames overlap 2.5x, HBM sees each sample ~once).  Each HALF-warp
// owns a frame: the 512-point real FFT runs as a 256-point complex FFT (16 x 16 Cooley-Tukey, one
// 16-point DFT per lane entirely in registers, a single padded shared-memory transpose between the two
// passes), followed by the real-FFT split, |X|^2, the sparse triangular mel projection (<= 2 filters per
// FFT bin) and log.  The utterance CMVN statistics are a separate kernel (cmvn_stats_kernel below): they replay
// numpy's frame-sequential fp32 accumulation bit for bit, which a per-chunk partial sum would not.
#include <math.h>

#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_BINS = 80;
constexpr int FB_FRAMES_PER_CTA = 24;   // frames per work item: 3 rounds of 8 half-warps; the sample tile is 16 KB
constexpr int FB_WAVE = FB_WIN + (FB_FRAMES_PER_CTA - 1) * FB_SHIFT;  // 4080 samples
// Mel weights: filters are taken in five groups of 16 (one per lane of a half-warp), every filter zero-padded to its
// group's longest (rounded up to even).  For the 80-bin 20 Hz .. Nyquist bank at 512 points these lengths are fixed
// numbers (mm_fbank_build_tables checks them), so the projection loop unrolls completely.  Weight i of lane l of group g
// lives at FB_GOFF(g) + 16 i + l: a half-warp reads 16 consecutive floats per step (no bank conflicts; both half-warps
// of a warp read the same address).  The per-filter layout this replaces had lane strides of 4 / 4 / 6 / 10 / 16
// floats = up to 8-way conflicts: 160 shared-memory wavefronts per frame instead of 40.
__host__ __device__ constexpr int FB_GMAX(int g) { return g == 0 ? 4 : g == 1 ? 4 : g == 2 ? 6 : g == 3 ? 10 : 16; }
// (closed form on purpose: a recursive constexpr function called with an unrolled loop's index is emitted as a run-time call)
__host__ __device__ constexpr int FB_GOFF(int g) { return 16 * (g == 0 ? 0 : g == 1 ? 4 : g == 2 ? 8 : g == 3 ? 14 : g == 4 ? 24 : 40); }
static_assert(FB_GOFF(1) == 16 * FB_GMAX(0) && FB_GOFF(2) == FB_GOFF(1) + 16 * FB_GMAX(1) && FB_GOFF(3) == FB_GOFF(2) + 16 * FB_GMAX(2) &&
              FB_GOFF(4) == FB_GOFF(3) + 16 * FB_GMAX(3) && FB_GOFF(5) == FB_GOFF(4) + 16 * FB_GMAX(4), "group offsets");
constexpr int FB_MAX_NNZ = FB_GOFF(5);   // 640
// table layout (floats)
constexpr int TB_WIN = 0;                         // [400]
constexpr int TB_TW256 = TB_WIN + 400;            // [256][2]  exp(-2 pi i m / 256)
constexpr int TB_TW512 = TB_TW256 + 512;          // [256][2]  exp(-2 pi i k / 512)
constexpr int TB_MELW = TB_TW512 + 512;           // [FB_MAX_NNZ]
constexpr int TB_K0 = TB_MELW + FB_MAX_NNZ;       // [80] int: first FFT bin of filter m
constexpr int TB_CNT = TB_K0 + 80;                // [80] int: number of bins
constexpr int TB_OFF = TB_CNT + 80;               // [80] int: offset into MELW
constexpr int TB_GMAX = TB_OFF + 80;              // [8] int: longest filter of each 16-filter group (rounded up to even)
constexpr int TB_TOTAL = TB_GMAX + 8;

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

constexpr int FB_THREADS = 128;
constexpr int FB_HW = FB_THREADS / 16;                 // 8 half-warps: 4 frames each per 32-frame item
// floats per half-warp: padded 16x17 complex transpose / spectrum (544) + 16, so that the two half-warps of a warp sit
// 16 banks apart: their 16-float power-spectrum stores share one wavefront and the mel gather's reads collide half as often
#ifndef MM_FB_XPAD
#define MM_FB_XPAD 16
#endif
constexpr int FB_XBUF = 16 * 17 * 2 + MM_FB_XPAD;
constexpr int FB_CTAS_PER_SM = 5;
// shared memory (floats): sample tile | per-half-warp transpose / spectrum | povey window | mel weights | mel meta
constexpr int FS_WAVE = 0;
constexpr int FS_WORK = FS_WAVE + FB_WAVE;
constexpr int FS_WIN = FS_WORK + FB_HW * FB_XBUF;
constexpr int FS_MELW = FS_WIN + FB_WIN;
constexpr int FS_META = FS_MELW + FB_MAX_NNZ;          // k0[80] | cnt[80] | off[80] | group max[8] (ints)
constexpr int FS_TW1 = FS_META + 248;                    // [16 k2][16 lanes] float2: W256^(l k2), lane-contiguous
constexpr int FB_SMEM_FLOATS = FS_TW1 + 512;
constexpr int FB_SMEM_BYTES = FB_SMEM_FLOATS * 4;      // 40.5 KB -> 5 CTAs per SM
static_assert((FB_SMEM_BYTES + 1024) * FB_CTAS_PER_SM <= 228 * 1024, "shared memory per SM");

// W32^k = exp(-2 pi i k / 32), k = 0..15: the bin-block part of the real-FFT split twiddle W512^(16 k1 + l)
__device__ __forceinline__ float2 w32(int k) {
  constexpr float C[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f,
                           0.0f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f,
                           -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  constexpr float S[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                           0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                           1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f};
  return make_float2(C[k], -S[k]);
}

// WavT = float (samples already in int16 range) or int16_t (raw PCM: halves the dominant HBM / PCIe read; the
// int16 -> fp32 conversion is exact, so both inputs give bit-identical features for integer-valued audio)
//
// Persistent: a CTA walks (utterance, 24-frame chunk) items.  What the previous version of this kernel fetched from
// global memory per frame -- FFT twiddles at lane-strided addresses (up to 16 L1 sectors per request; ncu: L1/shared
// pipe 76 % busy, DRAM 5 %), window, mel weights -- now lives in registers (W512^l) and shared memory
// (window, mel bank, pass-1 twiddles in a lane-contiguous layout), loaded once per CTA; the DC sum reuses the sample pairs the FFT loads anyway.
template <typename WavT>
__global__ void __launch_bounds__(FB_THREADS, FB_CTAS_PER_SM) fbank_kernel(const WavT* __restrict__ wav,
                                                                           const long long* __restrict__ n_samples,
                                                                           long long wav_stride, float* __restrict__ feats,
                                                                           int max_frames, int n_chunks, int n_items,
                                                                           const float* __restrict__ tables) {
  pdl_launch_dependents();   // programmatic dependent launch: see host.cuh launch_pdl
  extern __shared__ __align__(16) float fsm[];
  float* s_wave = fsm + FS_WAVE;
  float* s_work = fsm + FS_WORK;
  float2* s_win2 = reinterpret_cast<float2*>(fsm + FS_WIN);
  float* s_melw = fsm + FS_MELW;
  int* s_k0 = reinterpret_cast<int*>(fsm + FS_META);

  // ---- constants (tables are host-written once: nothing here depends on the previous kernel) ----
  // (window | mel weights + meta: two contiguous table regions copied as 128-bit words, all loads in flight at once)
  {
    static_assert(TB_WIN % 4 == 0 && TB_MELW % 4 == 0 && FB_WIN % 4 == 0 && (FB_MAX_NNZ + 248) % 4 == 0 &&
                  FS_WIN % 4 == 0 && FS_MELW % 4 == 0 && TB_K0 == TB_MELW + FB_MAX_NNZ && TB_TOTAL == TB_K0 + 248, "layout");
    constexpr int NW = FB_WIN / 4, NM = (FB_MAX_NNZ + 248) / 4, R = (NW + NM + FB_THREADS - 1) / FB_THREADS;
    const float4* gw = reinterpret_cast<const float4*>(tables + TB_WIN);
    const float4* gm = reinterpret_cast<const float4*>(tables + TB_MELW);
    float4* sw = reinterpret_cast<float4*>(fsm + FS_WIN);
    float4* sm = reinterpret_cast<float4*>(fsm + FS_MELW);
    float4 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = r * FB_THREADS + threadIdx.x;
      if (i < NW) v[r] = __ldg(gw + i);
      else if (i < NW + NM) v[r] = __ldg(gm + (i - NW));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = r * FB_THREADS + threadIdx.x;
      if (i < NW) sw[i] = v[r];
      else if (i < NW + NM) sm[i - NW] = v[r];
    }
  }
  const int hw = threadIdx.x >> 4;       // half-warp id 0..7
  const int l = threadIdx.x & 15;        // lane in half-warp
  const int hsel = hw & 1;
  // W256^(l k2), the pass-1 twiddles: a [k2][lane] table in shared memory (lane-contiguous: conflict-free 64-bit reads)
  float2* s_tw1 = reinterpret_cast<float2*>(fsm + FS_TW1);
  for (int i = threadIdx.x; i < 256; i += FB_THREADS)
    s_tw1[i] = __ldg(reinterpret_cast<const float2*>(tables + TB_TW256) + (((i & 15) * (i >> 4)) & 255));
  const float2 wl = __ldg(reinterpret_cast<const float2*>(tables + TB_TW512) + l);   // W512^l
  float* xbuf = s_work + hw * FB_XBUF;
  float2* x2 = reinterpret_cast<float2*>(xbuf);
  pdl_wait();
  __syncthreads();

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / n_chunks, chunk = item - b * n_chunks;
    const long long n = n_samples[b];
    int nfr = n < FB_WIN ? 0 : (int)(1 + (n - FB_WIN) / FB_SHIFT);
    nfr = min(nfr, max_frames);
    const int f0 = chunk * FB_FRAMES_PER_CTA;
    if (f0 >= nfr) continue;  // whole item out of range (uniform)
    const int nf_here = min(FB_FRAMES_PER_CTA, nfr - f0);

    // ---- stage the samples of this item's 24 frames (4080 values; frames overlap 2.5x) ----
    const long long s0 = (long long)f0 * FB_SHIFT;
    const int n_need = FB_WIN + (nf_here - 1) * FB_SHIFT;  // <= n - s0 by construction
    const WavT* wsrc = wav + (long long)b * wav_stride + s0;
    // All 128-bit loads of the tile are issued before the first one is consumed (one HBM round trip per item, not
    // one per loop iteration: ncu had 19 % of the stall samples on the load -> store dependency of a rolled loop).
    if constexpr (sizeof(WavT) == 4) {
      if ((reinterpret_cast<uintptr_t>(wsrc) & 15) == 0) {
        constexpr int R = (FB_WAVE + FB_THREADS * 4 - 1) / (FB_THREADS * 4);   // 8 rounds of 512 samples
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = (r * FB_THREADS + threadIdx.x) * 4;
          if (i + 3 < n_need) v[r] = __ldcs(reinterpret_cast<const float4*>(wsrc + i));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = (r * FB_THREADS + threadIdx.x) * 4;
          if (i + 3 < n_need) *reinterpret_cast<float4*>(s_wave + i) = v[r];
        }
        for (int j = (n_need & ~3) + threadIdx.x; j < n_need; j += FB_THREADS) s_wave[j] = wsrc[j];
      } else {
        for (int i = threadIdx.x; i < n_need; i += FB_THREADS) s_wave[i] = wsrc[i];
      }
    } else {
      if ((reinterpret_cast<uintptr_t>(wsrc) & 15) == 0) {     // 8 PCM samples per 128-bit load
        constexpr int R = (FB_WAVE + FB_THREADS * 8 - 1) / (FB_THREADS * 8);   // 4 rounds of 1024 samples
        uint4 q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = (r * FB_THREADS + threadIdx.x) * 8;
          if (i + 7 < n_need) q[r] = __ldcs(reinterpret_cast<const uint4*>(wsrc + i));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = (r * FB_THREADS + threadIdx.x) * 8;
          if (i + 7 < n_need) {
            const uint32_t w[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
            float4 lo, hi;
            lo.x = (float)(short)(w[0] & 0xFFFFu), lo.y = (float)(short)(w[0] >> 16);
            lo.z = (float)(short)(w[1] & 0xFFFFu), lo.w = (float)(short)(w[1] >> 16);
            hi.x = (float)(short)(w[2] & 0xFFFFu), hi.y = (float)(short)(w[2] >> 16);
            hi.z = (float)(short)(w[3] & 0xFFFFu), hi.w = (float)(short)(w[3] >> 16);
            *reinterpret_cast<float4*>(s_wave + i) = lo;
            *reinterpret_cast<float4*>(s_wave + i + 4) = hi;
          }
        }
        for (int j = (n_need & ~7) + threadIdx.x; j < n_need; j += FB_THREADS) s_wave[j] = (float)wsrc[j];
      } else {
        for (int i = threadIdx.x; i < n_need; i += FB_THREADS) s_wave[i] = (float)wsrc[i];
      }
    }
    __syncthreads();

    // Both half-warps of a warp run the same instruction stream on two different frames: the loop is uniform per
    // CTA (a half-warp without a frame in the last round recomputes the last one and skips the store), so every
    // shuffle / syncwarp uses the full mask -- a run-time half-warp mask costs a convergence barrier per shuffle.
    const int rounds = (nf_here + FB_HW - 1) / FB_HW;
    for (int it = 0; it < rounds; ++it) {
      const int fl_raw = hw + it * FB_HW;
      const bool live = fl_raw < nf_here;
      const int fl = live ? fl_raw : nf_here - 1;
      const float* x = s_wave + fl * FB_SHIFT;
      // ---- the lane's 13 sample pairs z[n] = (x[2n], x[2n+1]), n = l + 16 n2 < 200; their sum gives the DC term ----
      float2 a[16];
      float s = 0.f;
#pragma unroll
      for (int n2 = 0; n2 < 13; ++n2) {
        const int nn = l + 16 * n2;
        if (n2 < 12 || nn < FB_WIN / 2) {
          a[n2] = *reinterpret_cast<const float2*>(x + 2 * nn);
          s += a[n2].x + a[n2].y;
        } else {
          a[n2] = make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.0f / FB_WIN);
      // ---- y[i] = ((x[i]-mean) - 0.97 (x[max(i-1,0)]-mean)) * povey[i] for i < 400, else 0 ----
#pragma unroll
      for (int n2 = 0; n2 < 13; ++n2) {
        const int nn = l + 16 * n2;
        if (n2 < 12 || nn < FB_WIN / 2) {
          const float xm = x[nn > 0 ? 2 * nn - 1 : 0];
          const float2 w = s_win2[nn];
          const float c0 = a[n2].x - mean, c1 = a[n2].y - mean, cm = xm - mean;
          a[n2].x = __fmul_rn(__fsub_rn(c0, __fmul_rn(0.97f, cm)), w.x);
          a[n2].y = __fmul_rn(__fsub_rn(c1, __fmul_rn(0.97f, c0)), w.y);
        }
      }
      a[13] = a[14] = a[15] = make_float2(0.f, 0.f);
      // ---- pass 1: lane n1 = l, 16-point DFT over n2; twiddle W256^(n1 k2); padded transpose through smem ----
      dft16(a);
      x2[l] = a[0];
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2) x2[k2 * 17 + l] = cmul(a[k2], s_tw1[k2 * 16 + l]);
      __syncwarp();
      // ---- pass 2: lane k2 = l, 16-point DFT over n1 -> a[k1] = Z[16 k1 + l] ----
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) a[n1] = x2[l * 17 + n1];
      dft16(a);
      __syncwarp();   // all lanes have read the transpose: xbuf is reused for the power spectrum
      // ---- real-FFT split.  With E = (Z[k] + conj(Z[256-k])) / 2, O = -i (Z[k] - conj(Z[256-k])) / 2 and
      //      t = W512^k O:  |X[k]|^2 = |E + t|^2 and |X[256-k]|^2 = |E - t|^2, so each (k, 256-k) pair is formed once.
      //      k = 16 k1 + l lives in slot k1 of lane l, 256-k = 16 (15-k1) + (16-l) in slot 15-k1 of lane 16-l (lane 0:
      //      own slot 16-k1): lane l does k1 = 0..7 for its own bins and thereby lane 16-l's bins k1 = 8..15; only
      //      k = 128 (lane 0, k1 = 8, its own partner) is left over.  W512^(16 k1 + l) = W512^l W32^k1. ----
      const int src = ((16 - l) & 15) + 16 * hsel;
#pragma unroll
      for (int k1 = 0; k1 < 9; ++k1) {
        const float px = __shfl_sync(0xffffffffu, a[15 - k1].x, src);
        const float py = __shfl_sync(0xffffffffu, a[15 - k1].y, src);
        const float2 zk = a[k1];
        float2 zm;
        zm.x = (l == 0) ? a[(16 - k1) & 15].x : px;
        zm.y = -((l == 0) ? a[(16 - k1) & 15].y : py);
        const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
        const float2 d = csub(zk, zm);
        const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // -i/2 * (zk - zm)
        const float2 t = cmul(cmul(wl, w32(k1)), o);
        const float2 xp = cadd(e, t), xn = csub(e, t);
        if (k1 < 8) {
          xbuf[16 * k1 + l] = xp.x * xp.x + xp.y * xp.y;
          if (l != 0 || k1 != 0) xbuf[256 - 16 * k1 - l] = xn.x * xn.x + xn.y * xn.y;   // (bin 256 = Nyquist: dropped)
        } else if (l == 0) {
          xbuf[128] = xp.x * xp.x + xp.y * xp.y;
        }
      }
      __syncwarp();
      // ---- sparse mel projection + log ----
      float* orow = feats + ((long long)b * max_frames + (f0 + fl)) * FB_BINS;
      // filters 16 j .. 16 j + 15 have similar widths: every lane runs the group's longest filter (weights are
      // zero-padded to it in the table): compile-time trip counts, and the 80 loads of the five groups are independent
      // of each other (the logs and stores come after the last group, so nothing serialises the groups)
      int kk[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) kk[j] = s_k0[l + 16 * j];
      float e[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float* wgt = s_melw + FB_GOFF(j) + l;
        const float* pw = xbuf + kk[j];
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int i = 0; i < FB_GMAX(j); i += 2) {
          e0 = fmaf(wgt[16 * i], pw[i], e0);
          e1 = fmaf(wgt[16 * i + 16], pw[i + 1], e1);
        }
        e[j] = logf(fmaxf(e0 + e1, 1.1920928955078125e-07f));
      }
      if (live) {
#pragma unroll
        for (int j = 0; j < 5; ++j) orow[l + 16 * j] = e[j];
      }
      __syncwarp();
    }
    __syncthreads();   // every half-warp is done with the sample tile before the next item overwrites it
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
  int* gmax = reinterpret_cast<int*>(out + TB_GMAX);
  static float wall[FB_BINS][256];
  for (int m = 0; m < FB_BINS; ++m) {
    // torch: python-float scalars are applied in the tensor dtype (fp32)
    const float left = (float)mel_low + (float)m * (float)delta;
    const float center = (float)mel_low + ((float)m + 1.0f) * (float)delta;
    const float right = (float)mel_low + ((float)m + 2.0f) * (float)delta;
    int first = -1, last = -1;
    for (int k = 0; k < 256; ++k) {
      const float f = (float)fft_bin_width * (float)k;
      const float mel = 1127.0f * logf(1.0f + f / 700.0f);
      const float up = (mel - left) / (center - left);
      const float down = (right - mel) / (right - center);
      const float v = fmaxf(0.0f, fminf(up, down));
      wall[m][k] = v;
      if (v > 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    k0[m] = first < 0 ? 0 : first;
    cnt[m] = first < 0 ? 0 : last - first + 1;
  }
  for (int g = 0; g < FB_BINS / 16; ++g) {
    int mx = 0;
    for (int m = 16 * g; m < 16 * g + 16; ++m) mx = cnt[m] > mx ? cnt[m] : mx;
    mx = (mx + 1) & ~1;
    gmax[g] = mx;
    if (mx != FB_GMAX(g)) return bad_arg("fbank tables: mel bank group widths differ from the compiled-in ones");
    for (int m = 16 * g; m < 16 * g + 16; ++m) {
      off[m] = FB_GOFF(g) + (m - 16 * g);
      // zero padding up to the group's longest filter; k0 + i may run past bin 255 for the top filters: the spectrum
      // buffer is 560 floats long and the weight there is 0
      for (int i = 0; i < mx; ++i)
        out[TB_MELW + off[m] + 16 * i] = (i < cnt[m] && k0[m] + i < 256) ? wall[m][k0[m] + i] : 0.0f;
    }
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
  const long long n_items = (long long)n_chunks * batch;
  if (n_items > 0x7fffffffLL) return bad_arg("fbank: too many frames");
  const int grid = (int)(n_items < (long long)kNumSMs * FB_CTAS_PER_SM ? n_items : (long long)kNumSMs * FB_CTAS_PER_SM);
  launch_pdl(fbank_kernel<WavT>, dim3(grid), dim3(FB_THREADS), FB_SMEM_BYTES, static_cast<cudaStream_t>(stream),
      wav, reinterpret_cast<const long long*>(n_samples), wav_stride, feats, max_frames, n_chunks, (int)n_items, tables);
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
