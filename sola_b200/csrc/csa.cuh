// Carry-save popcount accumulation shared by K2 (pair_iou.cu) and the batched label counts (counts.cu).
#pragma once
#include "common.cuh"

namespace sola {

// Carry-save accumulation (Harley-Seal): POPC issues at 16 lanes/clk/SM on sm_100 (measured, profiles/r1_microbench_int_b200.jsonl)
// against 64 for LOP3, so the plain AND+POPC+IADD loop is POPC-bound.  Per pair we keep bit-sliced counters `ones`, `twos`
// and feed the four AND-ed words of a k-quad through three 3:2 compressors; only the weight-4 carry word is popcounted:
//   4 AND + 6 LOP3 + 1 POPC per 4 words  (2.5 alu ops and 0.25 POPC per word instead of 1 and 1).
// The compressors are written as explicit 3-input LOP3s (majority 0xE8, parity 0x96): left to itself the compiler fuses the ANDs
// into a chain of half adders (a ^ (b & c), a & b & c, or) that costs 12 LOP3 per quad instead of 10.  (A hybrid that sends some
// k-quads down the plain POPC route to use the idle POPC pipe measured 3-5 % slower, profiles/r1_k2_variants.jsonl.)
// total = acc + 2 * popc(twos) + popc(ones).
struct Csa { uint32_t ones, twos; int acc; };

__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// four more words into the bit-sliced counters
__device__ __forceinline__ void csa_add4(Csa& st, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
  const uint32_t t1 = lop3_maj(st.ones, x0, x1), s1 = lop3_xor3(st.ones, x0, x1);
  const uint32_t t2 = lop3_maj(s1, x2, x3);
  st.ones = lop3_xor3(s1, x2, x3);
  const uint32_t f = lop3_maj(st.twos, t1, t2);
  st.twos = lop3_xor3(st.twos, t1, t2);
  st.acc = __popc(f) * 4 + st.acc;
}

__device__ __forceinline__ void csa_quad(Csa& st, const uint4& a, const uint4& b) { csa_add4(st, a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }

__device__ __forceinline__ int csa_total(const Csa& st) { return st.acc + 2 * __popc(st.twos) + __popc(st.ones); }

}  // namespace sola
