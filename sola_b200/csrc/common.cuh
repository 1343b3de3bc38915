// Shared device/host helpers for the sola_maskpath kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define SOLA_OK 0
#define SOLA_ERR_INVALID (-1)   // bad argument (null pointer, non-positive size, misaligned buffer)
#define SOLA_ERR_CUDA (-2)      // a CUDA runtime call / launch failed; see sola_last_error_string()
#define SOLA_ERR_UNSUPPORTED (-3)

namespace sola {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SOLA_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      sola::set_error(__VA_ARGS__);             \
      return SOLA_ERR_INVALID;                  \
    }                                           \
  } while (0)

#define SOLA_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      sola::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
      return SOLA_ERR_CUDA;                                                    \
    }                                                                          \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// Streaming 128-bit load: read-once data must not displace the packed planes kept in L1/L2.
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cached[dev]) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

}  // namespace sola
