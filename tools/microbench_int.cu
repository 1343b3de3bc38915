// Integer-pipe micro-benchmark for the K2 roofline denominator: sustained per-SM rates of
// POPC, LOP3, IADD3 and the two candidate inner loops (AND+POPC+IADD vs carry-save adder tree + 1 POPC / 8 words).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_int.bin tools/microbench_int.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define UNROLL 16

__global__ void k_popc(uint32_t* out, uint32_t seed) {
  uint32_t x[UNROLL], acc[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { x[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u; acc[i] = 0; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { acc[i] = __popc(x[i] ^ acc[i]); }      // POPC + LOP3 (dependent chain per i)
  }
  uint32_t s = 0;
  for (int i = 0; i < UNROLL; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lop3(uint32_t* out, uint32_t seed) {
  uint32_t x[UNROLL], y[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { x[i] = seed * (threadIdx.x + 1) + i; y[i] = x[i] * 3u; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { x[i] = (x[i] & y[i]) ^ (y[i] | 0x5555u) ^ (uint32_t)it; y[i] = (y[i] ^ x[i]) | (x[i] & 0xff00ffu); }
  }
  uint32_t s = 0;
  for (int i = 0; i < UNROLL; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_iadd(uint32_t* out, uint32_t seed) {
  uint32_t x[UNROLL], y[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { x[i] = seed * (threadIdx.x + 1) + i; y[i] = x[i] * 3u; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { x[i] = x[i] + y[i] + (uint32_t)it; y[i] = y[i] + x[i] + 7u; }
  }
  uint32_t s = 0;
  for (int i = 0; i < UNROLL; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain inner loop: acc += popc(a & b), 16 independent pairs, operands mutate cheaply so nothing hoists
__global__ void k_and_popc_add(uint32_t* out, uint32_t seed) {
  uint32_t a[4], b[4];
  int acc[16];
  for (int i = 0; i < 4; ++i) { a[i] = seed * (threadIdx.x + 1) + i; b[i] = a[i] * 2654435761u; }
  for (int i = 0; i < 16; ++i) acc[i] = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = a[i] * 5u + 1u; b[i] = b[i] + 0x9e3779b9u; }   // 8 cheap ops per 16 pair-words
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i * 4 + j] += __popc(a[i] & b[j]);
  }
  int s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#define CSA(h, l, a, b, c) { uint32_t u_ = (a) ^ (b); h = ((a) & (b)) | (u_ & (c)); l = u_ ^ (c); }

// carry-save (Harley-Seal) inner loop: 8 words of one pair per group -> 7 CSA + 1 POPC (weight 8); 4 pairs per thread
__global__ void k_csa8(uint32_t* out, uint32_t seed) {
  uint32_t a[8], b[4];
  uint32_t ones[4], twos[4], fours[4];
  int acc[4];
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i;
  for (int j = 0; j < 4; ++j) { b[j] = a[j] * 2654435761u; ones[j] = twos[j] = fours[j] = 0; acc[j] = 0; }
  for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * 5u + 1u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      b[j] += 0x9e3779b9u;
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = a[i] & (b[j] ^ (uint32_t)i);
      uint32_t tA, tB, fA, fB, e;
      CSA(tA, ones[j], ones[j], v[0], v[1]);
      CSA(tB, ones[j], ones[j], v[2], v[3]);
      CSA(fA, twos[j], twos[j], tA, tB);
      CSA(tA, ones[j], ones[j], v[4], v[5]);
      CSA(tB, ones[j], ones[j], v[6], v[7]);
      CSA(fB, twos[j], twos[j], tA, tB);
      CSA(e, fours[j], fours[j], fA, fB);
      acc[j] += __popc(e);
    }
  }
  int s = 0;
  for (int j = 0; j < 4; ++j) s += 8 * acc[j] + 4 * __popc(fours[j]) + 2 * __popc(twos[j]) + __popc(ones[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static double run(K kern, const char* name, double ops_per_thread, int sms, uint32_t* out, double clock_hz) {
  const int threads = 256, blocks = sms * 8;
  kern<<<blocks, threads>>>(out, 12345u);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(out, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double total = ops_per_thread * threads * (double)blocks;
  const double rate = total / (best * 1e-3);
  printf("{\"kernel\": \"%s\", \"ms\": %.4f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_%.0fMHz\": %.2f}\n", name, best, rate,
         clock_hz / 1e6, rate / clock_hz / sms);
  return rate;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clock_hz = clk_khz * 1e3;
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f}\n", p.name, p.multiProcessorCount, clock_hz / 1e6);
  uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * 256 * p.multiProcessorCount * 8);
  const int sms = p.multiProcessorCount;
  run(k_popc, "popc(+xor) dependent x16", (double)ITERS * UNROLL, sms, out, clock_hz);
  run(k_lop3, "lop3 x4 per slot", (double)ITERS * UNROLL * 4, sms, out, clock_hz);
  run(k_iadd, "iadd3 x2 per slot", (double)ITERS * UNROLL * 2, sms, out, clock_hz);
  run(k_and_popc_add, "pair-words: and+popc+iadd", (double)ITERS * 16, sms, out, clock_hz);
  run(k_csa8, "pair-words: csa8 (7 CSA + 1 popc / 8 words)", (double)(ITERS / 2) * 4 * 8, sms, out, clock_hz);
  cudaFree(out);
  return 0;
}
