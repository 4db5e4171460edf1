// Gradient exchange of the training step over NVLink peer memory (BASELINE configs[2]: the path's one collective).
//
// One process per GPU; every rank maps the other ranks' flat gradient buffers (CUDA IPC) once.  The exchange is one
// kernel per rank, two-shot and in place:
//     rank r owns slice r of the buffer: it LOADS that slice from every rank over NVLink (its own from HBM), adds the
//     world's values in rank order (so the sum is bit-identical on every rank and from run to run), and STORES the
//     result into slice r of every rank's buffer.
// Slice r of any buffer is read and written by rank r alone, so the exchange kernel itself needs no flags; it is
// bracketed by two barriers over the ranks -- "every rank's gradients are complete" before, "every rank's stores have
// landed" after.  The barrier is a kernel too (mm_p2p_barrier: one warp; rank r stores a growing epoch number into slot r
// of every rank's flag array and spins until every slot of its own array has reached it), so the whole exchange is
// plain kernel launches: it is captured into the backward pass's CUDA graph on a side stream and overlaps the rest of
// the backward pass (an NCCL kernel cannot run beside the persistent GEMMs, which leave it no shared memory; these
// kernels use none).  Inbound loads and outbound stores use
// the two directions of the links at the same time: (N-1)/N of the buffer each way per GPU, the same bytes as a ring
// all-reduce, without its 2 (N-1) dependent steps.
#include "common.cuh"
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

namespace mm {

constexpr int P2P_MAX_RANKS = MM_P2P_MAX_RANKS;
constexpr int P2P_THREADS = 128;      // 128 threads x <= 64 registers and no shared memory: a CTA (8 K registers) fits beside
                                      // a persistent GEMM CTA (256 threads x 199 registers) on the same SM

struct BarrierArgs {
  unsigned int* flags[P2P_MAX_RANKS];  // flags[p] = rank p's flag array [P2P_MAX_RANKS] as mapped here
  unsigned int* epoch;                 // local counter (device memory), bumped once per barrier
  int world, rank;
};

// One warp.  Lane p < world: store the new epoch into slot `rank` of rank p's array, then wait until slot p of the local
// array has reached it.  Everything earlier in this stream (on every rank) is visible to everything later (on every rank).
__global__ void __launch_bounds__(32) p2p_barrier_kernel(const BarrierArgs a) {
  unsigned int e = 0;
  if (threadIdx.x == 0) {
    e = *a.epoch + 1u;
    *a.epoch = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  __threadfence_system();
  const int p = threadIdx.x;
  if (p < a.world) {
    volatile unsigned int* remote = a.flags[p] + a.rank;
    *remote = e;
    volatile unsigned int* mine = a.flags[a.rank] + p;
    while ((int)(*mine - e) < 0) {
    }
  }
  __syncwarp();
  __threadfence_system();
}

struct P2PArgs {
  float* buf[P2P_MAX_RANKS];   // buf[p] = rank p's buffer as mapped into THIS process (buf[rank] = the local one)
  int world, rank;
  long long lo, hi;            // this rank's slice [lo, hi), multiples of 4 elements except hi == n
};

template <int WORLD>
__global__ void __launch_bounds__(P2P_THREADS, 8) p2p_allreduce_kernel(const P2PArgs a) {
  constexpr int U = 8 / WORLD > 0 ? 8 / WORLD : 1;   // float4 positions per thread per round: 8 loads (128 B) in flight per thread
  const long long n4 = (a.hi - a.lo) >> 2;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += step * U) {
    float4 v[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * step;
      if (i < n4) {
#pragma unroll
        for (int p = 0; p < WORLD; ++p) v[u][p] = __ldcg(reinterpret_cast<const float4*>(a.buf[p] + a.lo + 4 * i));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * step;
      if (i < n4) {
        float4 s = v[u][0];
#pragma unroll
        for (int p = 1; p < WORLD; ++p) s.x += v[u][p].x, s.y += v[u][p].y, s.z += v[u][p].z, s.w += v[u][p].w;
#pragma unroll
        for (int p = 0; p < WORLD; ++p) __stcg(reinterpret_cast<float4*>(a.buf[p] + a.lo + 4 * i), s);
      }
    }
  }
  // tail (slice length not a multiple of 4: only the last rank's)
  const long long t0 = a.lo + 4 * n4;
  for (long long e = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.hi; e += step) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < WORLD; ++p) s += __ldcg(a.buf[p] + e);
#pragma unroll
    for (int p = 0; p < WORLD; ++p) __stcg(a.buf[p] + e, s);
  }
}

// 16-bit variant (the reference recipe trains with --fp16, scripts/textless/1_train.sh:125: fairseq exchanges 16-bit
// gradients): every rank first packs its fp32 gradients into a bf16 staging buffer (p2p_pack_kernel), the exchange runs on
// the staging buffers -- half the NVLink bytes -- with fp32 accumulation in rank order and ONE rounding of the sum, and
// p2p_unpack_kernel widens the (identical on every rank) result back into the fp32 gradient buffer.
struct P2PArgs16 {
  __nv_bfloat16* buf[P2P_MAX_RANKS];
  int world, rank;
  long long lo, hi;            // this rank's slice [lo, hi), multiples of 8 elements
};

__device__ __forceinline__ void bf16x8_add(float (&acc)[8], const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[2 * j] += __uint_as_float(w[j] << 16);
    acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
  }
}

template <int WORLD>
__global__ void __launch_bounds__(P2P_THREADS, 8) p2p_allreduce_bf16_kernel(const P2PArgs16 a) {
  constexpr int U = 8 / WORLD > 0 ? 8 / WORLD : 1;
  const long long n8 = (a.hi - a.lo) >> 3;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += step * U) {
    uint4 v[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * step;
      if (i < n8) {
#pragma unroll
        for (int p = 0; p < WORLD; ++p) v[u][p] = __ldcg(reinterpret_cast<const uint4*>(a.buf[p] + a.lo + 8 * i));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * step;
      if (i < n8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int p = 0; p < WORLD; ++p) bf16x8_add(acc, v[u][p]);
        uint4 r;
        r.x = OpTraits<__nv_bfloat16>::pack2(acc[0], acc[1]);
        r.y = OpTraits<__nv_bfloat16>::pack2(acc[2], acc[3]);
        r.z = OpTraits<__nv_bfloat16>::pack2(acc[4], acc[5]);
        r.w = OpTraits<__nv_bfloat16>::pack2(acc[6], acc[7]);
#pragma unroll
        for (int p = 0; p < WORLD; ++p) __stcg(reinterpret_cast<uint4*>(a.buf[p] + a.lo + 8 * i), r);
      }
    }
  }
}

// fp32 -> bf16 (n8 groups of 8 elements; the tail of the staging buffer beyond n is zero-filled) and back
__global__ void __launch_bounds__(256) p2p_pack_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       long long n, long long n_pad) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; 8 * i < n_pad; i += step) {
    float v[8];
    if (8 * i + 8 <= n) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(src + 8 * i));
      const float4 b = __ldcs(reinterpret_cast<const float4*>(src + 8 * i) + 1);
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 8 * i + j < n ? src[8 * i + j] : 0.f;
    }
    uint4 r;
    r.x = OpTraits<__nv_bfloat16>::pack2(v[0], v[1]);
    r.y = OpTraits<__nv_bfloat16>::pack2(v[2], v[3]);
    r.z = OpTraits<__nv_bfloat16>::pack2(v[4], v[5]);
    r.w = OpTraits<__nv_bfloat16>::pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + 8 * i) = r;
  }
}

__global__ void __launch_bounds__(256) p2p_unpack_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                                         long long n) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; 8 * i < n; i += step) {
    const uint4 r = __ldcg(reinterpret_cast<const uint4*>(src + 8 * i));      // written by the peers: L2, not L1
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[2 * j] = __uint_as_float(w[j] << 16), v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    if (8 * i + 8 <= n) {
      reinterpret_cast<float4*>(dst + 8 * i)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst + 8 * i)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (8 * i + j < n) dst[8 * i + j] = v[j];
    }
  }
}

typedef CUresult (*MemGetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static MemGetAddressRangeFn get_address_range_fn() {
  static MemGetAddressRangeFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<MemGetAddressRangeFn>(p);
  return fn;
}

}  // namespace mm

using namespace mm;

extern "C" int mm_ipc_get_handle(const void* ptr, uint8_t* handle64, int64_t* offset) {
  if (!ptr || !handle64 || !offset) return bad_arg("ipc_get_handle: null pointer");
  MemGetAddressRangeFn range = get_address_range_fn();
  if (!range) return bad_arg("cuMemGetAddressRange entry point not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return bad_arg("ipc_get_handle: not a device allocation");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) return fail(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (int64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return 0;
}

extern "C" int mm_ipc_open_handle(const uint8_t* handle64, void** mapped_base) {
  if (!handle64 || !mapped_base) return bad_arg("ipc_open_handle: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(mapped_base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(e, "cudaIpcOpenMemHandle");
  return 0;
}

extern "C" int mm_ipc_close_handle(void* mapped_base) {
  if (!mapped_base) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(mapped_base);
  if (e != cudaSuccess) return fail(e, "cudaIpcCloseMemHandle");
  return 0;
}

extern "C" int mm_p2p_allreduce_f32(float* const* bufs, int32_t world, int32_t rank, int64_t n, void* stream) {
  if (!bufs || world < 2 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || n <= 0)
    return bad_arg("p2p_allreduce: 2 .. MM_P2P_MAX_RANKS ranks");
  P2PArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    if (!bufs[p] || (reinterpret_cast<uintptr_t>(bufs[p]) & 15)) return bad_arg("p2p_allreduce: buffers must be 16-byte aligned");
    a.buf[p] = bufs[p];
  }
  a.world = world, a.rank = rank;
  const long long per = (((n + world - 1) / world) + 3) & ~3LL;     // slice length, a multiple of 4 elements
  a.lo = per * rank < n ? per * rank : n;
  a.hi = per * (rank + 1) < n ? per * (rank + 1) : n;
  if (a.hi <= a.lo) a.hi = a.lo;       // an empty slice still launches (graph capture keeps the same node set on every rank)
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n4 = (a.hi - a.lo + 3) >> 2;
  long long blocks = (n4 + P2P_THREADS - 1) / P2P_THREADS;
  if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned)blocks), block(P2P_THREADS);
  switch (world) {
    case 2: p2p_allreduce_kernel<2><<<grid, block, 0, s>>>(a); break;
    case 3: p2p_allreduce_kernel<3><<<grid, block, 0, s>>>(a); break;
    case 4: p2p_allreduce_kernel<4><<<grid, block, 0, s>>>(a); break;
    case 5: p2p_allreduce_kernel<5><<<grid, block, 0, s>>>(a); break;
    case 6: p2p_allreduce_kernel<6><<<grid, block, 0, s>>>(a); break;
    case 7: p2p_allreduce_kernel<7><<<grid, block, 0, s>>>(a); break;
    case 8: p2p_allreduce_kernel<8><<<grid, block, 0, s>>>(a); break;
    default: return bad_arg("p2p_allreduce: world size");
  }
  MM_CHECK_LAUNCH("p2p_allreduce_kernel launch");
  return 0;
}

extern "C" int mm_p2p_barrier(unsigned int* const* flags, unsigned int* epoch, int32_t world, int32_t rank, void* stream) {
  if (!flags || !epoch || world < 2 || world > P2P_MAX_RANKS || rank < 0 || rank >= world) return bad_arg("p2p_barrier");
  BarrierArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    if (!flags[p]) return bad_arg("p2p_barrier: null flag array");
    a.flags[p] = flags[p];
  }
  a.epoch = epoch, a.world = world, a.rank = rank;
  p2p_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MM_CHECK_LAUNCH("p2p_barrier_kernel launch");
  return 0;
}

extern "C" int mm_p2p_pack_bf16(const float* src, void* stage, int64_t n, int64_t n_pad, void* stream) {
  if (!src || !stage || n <= 0 || n_pad < n || (n_pad % 8) || (reinterpret_cast<uintptr_t>(src) & 15) ||
      (reinterpret_cast<uintptr_t>(stage) & 15))
    return bad_arg("p2p_pack_bf16: 16-byte aligned buffers, n_pad a multiple of 8 and >= n");
  long long blocks = (n_pad / 8 + 255) / 256;
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  p2p_pack_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(stage), n, n_pad);
  MM_CHECK_LAUNCH("p2p_pack_kernel launch");
  return 0;
}

extern "C" int mm_p2p_unpack_bf16(const void* stage, float* dst, int64_t n, void* stream) {
  if (!dst || !stage || n <= 0 || (reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(stage) & 15))
    return bad_arg("p2p_unpack_bf16: 16-byte aligned buffers");
  long long blocks = ((n + 7) / 8 + 255) / 256;
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  p2p_unpack_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(stage), dst, n);
  MM_CHECK_LAUNCH("p2p_unpack_kernel launch");
  return 0;
}

extern "C" int mm_p2p_allreduce_bf16(void* const* bufs, int32_t world, int32_t rank, int64_t n, void* stream) {
  if (!bufs || world < 2 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || n <= 0 || (n % 8))
    return bad_arg("p2p_allreduce_bf16: 2 .. MM_P2P_MAX_RANKS ranks, n a multiple of 8");
  P2PArgs16 a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    if (!bufs[p] || (reinterpret_cast<uintptr_t>(bufs[p]) & 15)) return bad_arg("p2p_allreduce_bf16: buffers must be 16-byte aligned");
    a.buf[p] = static_cast<__nv_bfloat16*>(bufs[p]);
  }
  a.world = world, a.rank = rank;
  const long long per = (((n + world - 1) / world) + 7) & ~7LL;
  a.lo = per * rank < n ? per * rank : n;
  a.hi = per * (rank + 1) < n ? per * (rank + 1) : n;
  if (a.hi <= a.lo) a.hi = a.lo;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n8 = (a.hi - a.lo) >> 3;
  long long blocks = (n8 + P2P_THREADS - 1) / P2P_THREADS;
  if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned)blocks), block(P2P_THREADS);
  switch (world) {
    case 2: p2p_allreduce_bf16_kernel<2><<<grid, block, 0, s>>>(a); break;
    case 3: p2p_allreduce_bf16_kernel<3><<<grid, block, 0, s>>>(a); break;
    case 4: p2p_allreduce_bf16_kernel<4><<<grid, block, 0, s>>>(a); break;
    case 5: p2p_allreduce_bf16_kernel<5><<<grid, block, 0, s>>>(a); break;
    case 6: p2p_allreduce_bf16_kernel<6><<<grid, block, 0, s>>>(a); break;
    case 7: p2p_allreduce_bf16_kernel<7><<<grid, block, 0, s>>>(a); break;
    case 8: p2p_allreduce_bf16_kernel<8><<<grid, block, 0, s>>>(a); break;
    default: return bad_arg("p2p_allreduce_bf16: world size");
  }
  MM_CHECK_LAUNCH("p2p_allreduce_bf16_kernel launch");
  return 0;
}
