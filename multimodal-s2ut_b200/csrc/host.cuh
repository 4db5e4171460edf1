// Host-side helpers shared by the C-ABI entry points: error capture and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mm {

extern thread_local char g_last_error[256];

inline int fail(cudaError_t e, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}
inline int bad_arg(const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "invalid argument: %s", what);
  return static_cast<int>(cudaErrorInvalidValue);
}
#define MM_CHECK_LAUNCH(name)                              \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return ::mm::fail(e__, name);  \
  } while (0)

// Launch with the programmatic-stream-serialization attribute: the kernel may become resident while its predecessor
// drains.  Every kernel launched this way calls pdl_wait() (griddepcontrol.wait) before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled();

// 3-D tensor map over a 16-bit tensor [dim2][dim1][dim0] (dim0 contiguous), 128B swizzle, zero OOB fill.
// Strides in ELEMENTS; box = {64, box_rows, 1}.
int make_tmap_3d(CUtensorMap* out, const void* base, int is_f16, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                 uint64_t stride1, uint64_t stride2, uint32_t box_rows);

// General form: kind 0 = bf16, 1 = f16, 2 = f32; box = {box0, box_rows, 1} (box0 * elem size must be 128 B).
int make_tmap_3d_ex(CUtensorMap* out, const void* base, int kind, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                    uint64_t stride1, uint64_t stride2, uint32_t box0, uint32_t box_rows);

}  // namespace mm
