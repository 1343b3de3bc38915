// Device building blocks of R1 (shared by resize.cu and fused_pack_resize.cu): ATen-exact bilinear geometry, the
// per-CTA shared-memory tables, and the two-phase "uniform window / edge word" evaluation of a tile held in smem.
#pragma once
#include "common.cuh"

namespace sola {

struct Axis { int i0, i1; float l0, l1; };

__device__ __forceinline__ Axis bilinear_axis(int dst, float scale, int in_size) {
  float src = __fmaf_rn((float)dst + 0.5f, scale, -0.5f);
  src = (src >= 0.f) ? src : 0.f;
  Axis a;
  a.i0 = (int)src;
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = __fsub_rn(src, (float)a.i0);
  a.l0 = __fsub_rn(1.f, a.l1);
  return a;
}

__device__ __forceinline__ float bilinear_val(const Axis& ax, const Axis& ay, float v00, float v01, float v10, float v11) {
  const float top = __fmaf_rn(ax.l0, v00, __fmul_rn(ax.l1, v01));
  const float bot = __fmaf_rn(ax.l0, v10, __fmul_rn(ax.l1, v11));
  return __fmaf_rn(ay.l0, top, __fmul_rn(ay.l1, bot));
}

__device__ __forceinline__ uint32_t get_bit(const uint32_t* __restrict__ row, int x) { return (__ldg(row + (x >> 5)) >> (x & 31)) & 1u; }

struct __align__(16) XParam { int x0, x1; float w0, w1; };
struct __align__(16) YParam { int y0, y1; float h0, h1; };     // rows relative to the tile's first input row


constexpr int R1_TR = 32;          // output rows per tile
constexpr int R1_THREADS = 256;

// Shared-memory tables of one tile: per output pixel, per output row, per output word column.
struct R1Tables {
  XParam* xtab;   // [owp * 32]
  YParam* ytab;   // [R1_TR]      rows relative to the tile's first input row
  int4* ctab;     // [owp]        first / last source word, care masks of the first / last word
  __device__ __forceinline__ static size_t bytes(int owp) { return (size_t)owp * 32 * sizeof(XParam) + R1_TR * sizeof(YParam) + (size_t)owp * sizeof(int4); }
  __device__ __forceinline__ void carve(unsigned char* base, int owp) {
    xtab = reinterpret_cast<XParam*>(base);
    ytab = reinterpret_cast<YParam*>(xtab + owp * 32);
    ctab = reinterpret_cast<int4*>(ytab + R1_TR);
  }
  // to be followed by __syncthreads()
  __device__ __forceinline__ void build(int oy0, int ylo, int H, int W, int oh, int ow, float sy, float sx) {
    const int owp = (ow + 31) >> 5;
    for (int i = threadIdx.x; i < owp * 32; i += blockDim.x) {
      const Axis a = bilinear_axis(min(i, ow - 1), sx, W);
      xtab[i] = XParam{a.i0, a.i1, a.l0, a.l1};
    }
    if (threadIdx.x < R1_TR) {
      const Axis a = bilinear_axis(min(oy0 + (int)threadIdx.x, oh - 1), sy, H);
      ytab[threadIdx.x] = YParam{a.i0 - ylo, a.i1 - ylo, a.l0, a.l1};
    }
    for (int c = threadIdx.x; c < owp; c += blockDim.x) {
      const int xa = bilinear_axis(c * 32, sx, W).i0;                       // first source pixel any lane of this word reads
      const int xb = bilinear_axis(min(c * 32 + 31, ow - 1), sx, W).i1;     // last one
      ctab[c] = make_int4(xa >> 5, xb >> 5, (int)(0xffffffffu << (xa & 31)), (int)(0xffffffffu >> (31 - (xb & 31))));
    }
  }
};

// Resize one tile whose source rows sit in shared memory (`t`, row pitch Wp words, first row = the tile's ylo).
//   phase A, one THREAD per output word: OR / AND of the cared-for source bits of both rows; an all-0 or all-1
//            window resolves the whole word (background / interior);
//   phase B, one WARP per remaining (edge) word, lane = output pixel: ATen's exact fma sequence on the 4 bits.
// Returns this thread's popcount of the words it wrote (for the per-frame area).
__device__ __forceinline__ int resize_tile_from_smem(const uint32_t* __restrict__ t, int Wp, const R1Tables& tb, int nrows, int ow,
                                                     uint32_t* __restrict__ out_rows /* first output row of the tile */) {
  const int owp = (ow + 31) >> 5;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_words = nrows * owp;
  int area_acc = 0;
  for (int base = 0; base < n_words; base += blockDim.x) {
    const int i = base + tid;
    const bool have = i < n_words;
    uint32_t word = 0;
    bool edge = false;
    int r = 0, c = 0;
    if (have) {
      r = i / owp; c = i - r * owp;
      const YParam yp = tb.ytab[r];
      const int px_first = c * 32, px_last = min(c * 32 + 31, ow - 1);
      const int4 win = tb.ctab[c];
      const int wlo = win.x, whi = win.y;
      const uint32_t* r0 = t + yp.y0 * Wp;
      const uint32_t* r1 = t + yp.y1 * Wp;
      uint32_t any1 = 0u, all1 = 0xffffffffu;
      for (int w = wlo; w <= whi; ++w) {
        // only the source pixels [xa, xb] matter: bits outside are forced to "don't care" for both tests
        uint32_t care = 0xffffffffu;
        if (w == wlo) care &= (uint32_t)win.z;
        if (w == whi) care &= (uint32_t)win.w;
        const uint32_t v0 = r0[w], v1 = r1[w];
        any1 |= (v0 | v1) & care;
        all1 &= (v0 & v1) | ~care;
      }
      if (any1 == 0u) word = 0u;
      else if (all1 == 0xffffffffu) word = (px_last - px_first == 31) ? 0xffffffffu : ((1u << (px_last - px_first + 1)) - 1u);
      else edge = true;
    }
    unsigned pending = __ballot_sync(FULL, edge);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const int rr = __shfl_sync(FULL, r, src), cc = __shfl_sync(FULL, c, src);
      const int ox = cc * 32 + lane;
      const XParam xp = tb.xtab[ox];
      const YParam yp = tb.ytab[rr];
      const uint32_t* r0 = t + yp.y0 * Wp;
      const uint32_t* r1 = t + yp.y1 * Wp;
      // v in {0,1}: w*v is w or +0 exactly, and fma(w0, v00, t) is fl(w0*v00 + t) = fl((w0 & m00) + t)
      const uint32_t m00 = 0u - ((r0[xp.x0 >> 5] >> (xp.x0 & 31)) & 1u), m01 = 0u - ((r0[xp.x1 >> 5] >> (xp.x1 & 31)) & 1u);
      const uint32_t m10 = 0u - ((r1[xp.x0 >> 5] >> (xp.x0 & 31)) & 1u), m11 = 0u - ((r1[xp.x1 >> 5] >> (xp.x1 & 31)) & 1u);
      const uint32_t w0b = __float_as_uint(xp.w0), w1b = __float_as_uint(xp.w1);
      const float top = __fadd_rn(__uint_as_float(w0b & m00), __uint_as_float(w1b & m01));
      const float bot = __fadd_rn(__uint_as_float(w0b & m10), __uint_as_float(w1b & m11));
      const float val = __fmaf_rn(yp.h0, top, __fmul_rn(yp.h1, bot));
      const uint32_t wv = __ballot_sync(FULL, ox < ow && val > 0.5f);
      if (lane == src) word = wv;
    }
    if (have) {
      out_rows[(long long)r * owp + c] = word;
      area_acc += __popc(word);
    }
  }
  return area_acc;
}

}  // namespace sola
