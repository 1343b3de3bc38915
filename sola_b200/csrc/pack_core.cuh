// Device building blocks of K1 (shared by binarize_pack.cu and fused_pack_resize.cu): per-vector predicate extraction
// and the in-warp slot transpose that turns lane-strided predicate bits into finished 32-pixel words.
#pragma once
#include "common.cuh"
#include <type_traits>
#include <string.h>

namespace sola {

enum PackMode { MODE_THRESH3 = 0, MODE_THRESH1 = 1, MODE_NONZERO = 2 };

struct Thresholds {
  float mid, hi, lo;
  uint32_t hi_bf2, lo_bf2, mid_bf2;   // floor-to-bf16 of hi / lo / mid, replicated in both halves (bf16 packed-compare path)
};

// largest bf16 value <= t, as its 16 bits: for a bf16 x, (x > t) == (x > bf16_floor(t)), so the packed bf16 compare is exact
inline uint32_t bf16_floor_bits(float t) {
  uint32_t b;
  memcpy(&b, &t, 4);
  uint32_t h = b >> 16;
  if ((b & 0xffffu) && (b >> 31)) h += 1;        // negative and inexact: one step towards -inf
  return h & 0xffffu;
}

// numpy / torch compare float32 data against the Python scalar cast to float32 (prompt_generator.py:177,182).  A threshold of
// -0.0 is normalised to +0.0: `x > -0.0` and `x > +0.0` are the same predicate, but the sign-of-(t - x) extraction below would
// see (-0.0) - (+0.0) = -0.0 as "greater".
inline Thresholds make_thresholds(double thr, double off) {
  Thresholds t;
  t.mid = (float)thr + 0.0f;
  t.hi = (float)(thr + off) + 0.0f;
  t.lo = (float)(thr - off) + 0.0f;
  t.hi_bf2 = bf16_floor_bits(t.hi) * 0x00010001u;
  t.lo_bf2 = bf16_floor_bits(t.lo) * 0x00010001u;
  t.mid_bf2 = bf16_floor_bits(t.mid) * 0x00010001u;
  return t;
}

template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> { static constexpr int E = 4; };
template <> struct ElemTraits<__nv_bfloat16> { static constexpr int E = 8; };
template <> struct ElemTraits<uint8_t> { static constexpr int E = 16; };

__device__ __forceinline__ uint32_t gt_bit(float x, float t) { return x > t ? 1u : 0u; }   // NaN -> 0

// ---- per-128-bit-load predicate extraction -------------------------------------------------------------------
// Threshold modes use the sign of (t - x): with IEEE subtraction (denormals kept, canonical positive NaN on sm_100)
// sign(t - x) == 1  <=>  x > t, for every input including +-0, +-inf, denormals and NaN (-> 0, as `NaN > t` is False).
// That is one FADD on the fma pipe plus one funnel shift on the alu pipe per (element, threshold) — the shift pushes the
// sign bit into an accumulator — instead of FSETP + SEL + shift/or, which kept the alu pipe ~70 % busy at HBM speed.
// Elements are pushed last-to-first so that element 0 of the first vector ends up in bit 0.
__device__ __forceinline__ uint32_t push_gt(uint32_t acc, float x, float t) {
  return __funnelshift_l(__float_as_uint(__fsub_rn(t, x)), acc, 1);
}


// Returns E bits (element c of the vector -> bit c) for mid, and (MODE_THRESH3 only) hi / lo.   [MODE_NONZERO path]
template <int MODE>
__device__ __forceinline__ void vec_bits(const uint4& raw, const Thresholds& th, float, uint32_t& mid, uint32_t& hi, uint32_t& lo) {
  const float v[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w)};
  mid = hi = lo = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (MODE == MODE_NONZERO) {
      mid |= (v[c] != 0.0f ? 1u : 0u) << c;
    } else {
      mid |= gt_bit(v[c], th.mid) << c;
      if (MODE == MODE_THRESH3) {
        hi |= gt_bit(v[c], th.hi) << c;
        lo |= gt_bit(v[c], th.lo) << c;
      }
    }
  }
}

// push the E elements of one vector (last element first) into the three accumulators
template <int MODE>
__device__ __forceinline__ void vec_push(const uint4& raw, const Thresholds& th, float, uint32_t& mid, uint32_t& hi, uint32_t& lo) {
  const float v[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w)};
#pragma unroll
  for (int c = 3; c >= 0; --c) {
    mid = push_gt(mid, v[c], th.mid);
    if (MODE == MODE_THRESH3) { hi = push_gt(hi, v[c], th.hi); lo = push_gt(lo, v[c], th.lo); }
  }
}

template <int MODE>
__device__ __forceinline__ void vec_push(const uint4& raw, const Thresholds& th, __nv_bfloat16, uint32_t& mid, uint32_t& hi, uint32_t& lo) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int c = 7; c >= 0; --c) {
    const float x = __uint_as_float((c & 1) ? (w[c >> 1] & 0xffff0000u) : (w[c >> 1] << 16));   // bf16 -> fp32 is exact
    mid = push_gt(mid, x, th.mid);
    if (MODE == MODE_THRESH3) { hi = push_gt(hi, x, th.hi); lo = push_gt(lo, x, th.lo); }
  }
}

template <int MODE>
__device__ __forceinline__ void vec_push(const uint4&, const Thresholds&, uint8_t, uint32_t&, uint32_t&, uint32_t&) {}

template <int MODE>
__device__ __forceinline__ void vec_bits(const uint4& raw, const Thresholds& th, __nv_bfloat16, uint32_t& mid, uint32_t& hi, uint32_t& lo) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  mid = hi = lo = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    // bf16 -> fp32 is exact: the 16 bits are the high half of the fp32 pattern
    const float x = __uint_as_float((c & 1) ? (w[c >> 1] & 0xffff0000u) : (w[c >> 1] << 16));
    if (MODE == MODE_NONZERO) {
      mid |= (x != 0.0f ? 1u : 0u) << c;
    } else {
      mid |= gt_bit(x, th.mid) << c;
      if (MODE == MODE_THRESH3) {
        hi |= gt_bit(x, th.hi) << c;
        lo |= gt_bit(x, th.lo) << c;
      }
    }
  }
}

// 4 bytes -> 4 bits "byte != 0" (byte k -> bit k).
__device__ __forceinline__ uint32_t nonzero_bytes4(uint32_t w) {
  const uint32_t m = __vcmpne4(w, 0u) & 0x01010101u;
  return (m * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t gt_bytes4(uint32_t w, uint32_t thr_rep) {
  const uint32_t m = __vcmpgtu4(w, thr_rep) & 0x01010101u;
  return (m * 0x01020408u) >> 24;
}

template <int MODE>
__device__ __forceinline__ void vec_bits(const uint4& raw, const Thresholds& th, uint8_t, uint32_t& mid, uint32_t& hi, uint32_t& lo) {
  hi = lo = 0;
  if (MODE == MODE_NONZERO) {
    mid = nonzero_bytes4(raw.x) | (nonzero_bytes4(raw.y) << 4) | (nonzero_bytes4(raw.z) << 8) | (nonzero_bytes4(raw.w) << 12);
  } else {
    // u8 "logits": integer compare against floor(threshold), clamped to [0,255]; x > t  <=>  x > floor(t) for integers x
    const int ti = th.mid < 0.f ? -1 : (th.mid >= 255.f ? 255 : (int)floorf(th.mid));
    if (ti < 0) { mid = 0xffffu; return; }
    const uint32_t rep = 0x01010101u * (uint32_t)ti;
    mid = gt_bytes4(raw.x, rep) | (gt_bytes4(raw.y, rep) << 4) | (gt_bytes4(raw.z, rep) << 8) | (gt_bytes4(raw.w, rep) << 12);
  }
}

// ---- L x L slot transpose inside groups of L lanes -----------------------------------------------------------
template <int E>
__host__ __device__ constexpr uint32_t keep_mask(int d) {
  uint32_t m = 0;
  const uint32_t slot = (E >= 32) ? 0xffffffffu : ((1u << E) - 1u);
  for (int j = 0; j < 32 / E; ++j)
    if (!(j & d)) m |= slot << (E * j);
  return m;
}

template <int E>
__device__ __forceinline__ uint32_t transpose_slots(uint32_t x, int lane) {
  constexpr int L = 32 / E;
#pragma unroll
  for (int d = L / 2; d >= 1; d >>= 1) {
    const uint32_t lo = keep_mask<E>(d);
    const uint32_t v = __shfl_xor_sync(FULL, x, d);
    x = (lane & d) ? ((x & ~lo) | ((v >> (E * d)) & lo)) : ((x & lo) | ((v << (E * d)) & ~lo));
  }
  return x;
}

// ---- bf16 packed-compare path ---------------------------------------------------------------------------------------------
// With bf16 logits a 32-bit word holds two elements and `set.gt.bf16x2` with a bit-mask result (SASS HSET2.BF16_V2.BM) compares both
// halves in ONE fma/half-pipe instruction: 0xFFFF per half that passes, NaN -> 0 (ordered compare), -0 > +0 false.  The
// thresholds are floored to bf16 first, which is exact: for a bf16 x, (x > t) == (x > bf16_floor(t)).
//   * stability counts: the masks are SUBTRACTED from a 32-bit accumulator, two masks per IADD3.  Subtracting 0x0000FFFF adds
//     1 - 65536 and subtracting 0xFFFF0000 adds 65536 (mod 2^32), so after a low-half hits and b high-half hits
//     acc = a + 65536 * (b - a)  — decoded once per frame by stab_decode (a < 65536 per thread per frame by a wide margin).
//     0.5 HSET2 + 0.25 IADD3 per element per threshold instead of unpack + FADD + funnel shift per element.
//   * stored plane: PRMT gathers one byte of each of four masks (4 elements -> bytes 0x00 / 0xFF), two LOP3 keep bit k of byte k
//     (elements 0-3) and bit 4+k of byte k (elements 4-7), and a multiply by 0x01010101 sums the four bytes into the top byte =
//     the 8 predicate bits in element order, which one funnel shift pushes into the private word.
__device__ __forceinline__ uint32_t gt2_mask(uint32_t x2, uint32_t t2) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&x2), *reinterpret_cast<const __nv_bfloat162*>(&t2));
}

__device__ __forceinline__ uint32_t vec_bits8_bf16(const uint4& raw, uint32_t t2) {
  const uint32_t m0 = gt2_mask(raw.x, t2), m1 = gt2_mask(raw.y, t2), m2 = gt2_mask(raw.z, t2), m3 = gt2_mask(raw.w, t2);
  const uint32_t p_lo = __byte_perm(m0, m1, 0x6420);         // bytes: elements 0,1,2,3 as 0x00 / 0xFF
  const uint32_t p_hi = __byte_perm(m2, m3, 0x6420);         //        elements 4,5,6,7
  const uint32_t t = (p_lo & 0x08040201u) | (p_hi & 0x80402010u);
  return t * 0x01010101u;                                    // top byte = predicate bits of elements 0..7
}

__device__ __forceinline__ void vec_count2(const uint4& raw, const Thresholds& th, int& acc_hi, int& acc_lo) {
  acc_hi = acc_hi - (int)gt2_mask(raw.x, th.hi_bf2) - (int)gt2_mask(raw.y, th.hi_bf2);
  acc_hi = acc_hi - (int)gt2_mask(raw.z, th.hi_bf2) - (int)gt2_mask(raw.w, th.hi_bf2);
  acc_lo = acc_lo - (int)gt2_mask(raw.x, th.lo_bf2) - (int)gt2_mask(raw.y, th.lo_bf2);
  acc_lo = acc_lo - (int)gt2_mask(raw.z, th.lo_bf2) - (int)gt2_mask(raw.w, th.lo_bf2);
}

// Stability accumulators as chunk_extract3<.., WANT_BITS = false> leaves them: plain counts for fp32, the packed form above for bf16.
template <typename T> __device__ __forceinline__ int stab_decode(int acc) { return acc; }
template <> __device__ __forceinline__ int stab_decode<__nv_bfloat16>(int acc) {
  const int a = acc & 0xffff;
  return 2 * a + ((acc - a) >> 16);
}

// One chunk's predicate extraction.  Returns the stored plane's private word in xm; the two stability thresholds either as bit
// words (xh, xl — WANT_BITS, needed when a range mask must be applied) or accumulated into acc_hi / acc_lo in the per-dtype
// encoding that stab_decode<T> undoes (never mix plain counts into those accumulators).
template <typename T, int L, bool WANT_BITS>
__device__ __forceinline__ void chunk_extract3(const uint4 (&raw)[L], const Thresholds& th, uint32_t& xm, uint32_t& xh, uint32_t& xl,
                                               int& acc_hi, int& acc_lo) {
  xm = xh = xl = 0;
  if (sizeof(T) == 2 && !WANT_BITS) {
#pragma unroll
    for (int j = L - 1; j >= 0; --j) {
      xm = __funnelshift_l(vec_bits8_bf16(raw[j], th.mid_bf2), xm, 8);
      vec_count2(raw[j], th, acc_hi, acc_lo);
    }
  } else {
#pragma unroll
    for (int j = L - 1; j >= 0; --j) vec_push<MODE_THRESH3>(raw[j], th, T(), xm, xh, xl);
    if (!WANT_BITS) { acc_hi += __popc(xh); acc_lo += __popc(xl); }
  }
}

// any-width stand-alone K1 (flat run + re-cut), defined in fused_pack_resize.cu
template <typename T>
int launch_band_pack(const T* in, long long n_frames, int H, int W, Thresholds th, uint32_t* packed, int* cnt_hi, int* cnt_mid, int* cnt_lo,
                     cudaStream_t stream);

}  // namespace sola
