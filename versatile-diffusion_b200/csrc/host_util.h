// Host-side helpers shared by the vdb200 translation units: error reporting, launch counting,
// SM count, TMA tensor-map construction through the driver entry point (no libcuda link needed).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace vdb {

int set_error(int code, const char* fmt, ...);   // records message, returns code
int num_sms();
void count_launch(int n = 1);

#define VDB_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::vdb::set_error(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// bf16 tensor maps, SWIZZLE_128B, zero OOB fill. Strides in bytes (dim0 is contiguous).
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t stride1, uint32_t box0,
                 uint32_t box1);
int make_tmap_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                 uint64_t stride1, uint64_t stride2, uint64_t stride3, uint32_t box0, uint32_t box1,
                 uint32_t box2, uint32_t box3);

}  // namespace vdb
