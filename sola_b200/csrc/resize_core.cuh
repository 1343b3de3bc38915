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
struct __align__(16) YParam { int y0, y1; float h0, h1; };     // WORD offsets (row * pitch) of the two source rows inside the tile


constexpr int R1_TR = 32;          // output rows per tile
constexpr int R1_THREADS = 256;

// Shared-memory tables of one tile: per output pixel, per output row, per output word column.
struct R1Tables {
  XParam* xtab;   // [owp * 32]
  YParam* ytab;   // [R1_TR]      word offsets of the two source rows relative to the tile's first input row
  int4* ctab;     // [owp]        win3: first source word + care masks of words wlo, wlo+1, wlo+2 (0 = unused);
                  //              otherwise: first / last source word, care masks of the first / last word
  bool win3;      // every output word's source window spans <= 3 words (any scale factor < 2): branch-free phase A
  __device__ __forceinline__ static size_t bytes(int owp) { return (size_t)owp * 32 * sizeof(XParam) + R1_TR * sizeof(YParam) + (size_t)owp * sizeof(int4); }
  __device__ __forceinline__ void carve(unsigned char* base, int owp) {
    xtab = reinterpret_cast<XParam*>(base);
    ytab = reinterpret_cast<YParam*>(xtab + owp * 32);
    ctab = reinterpret_cast<int4*>(ytab + R1_TR);
  }
  // to be followed by __syncthreads()
  __device__ __forceinline__ void build(int oy0, int ylo, int H, int W, int oh, int ow, float sy, float sx, int pitch) {
    const int owp = (ow + 31) >> 5;
    win3 = sx < 1.99f;                 // window <= 31 * sx + 3 source pixels -> at most 3 words
    for (int i = threadIdx.x; i < owp * 32; i += blockDim.x) {
      const Axis a = bilinear_axis(min(i, ow - 1), sx, W);
      xtab[i] = XParam{a.i0, a.i1, a.l0, a.l1};
    }
    if (threadIdx.x < R1_TR) {
      const Axis a = bilinear_axis(min(oy0 + (int)threadIdx.x, oh - 1), sy, H);
      ytab[threadIdx.x] = YParam{(a.i0 - ylo) * pitch, (a.i1 - ylo) * pitch, a.l0, a.l1};
    }
    for (int c = threadIdx.x; c < owp; c += blockDim.x) {
      const int xa = bilinear_axis(c * 32, sx, W).i0;                       // first source pixel any lane of this word reads
      const int xb = bilinear_axis(min(c * 32 + 31, ow - 1), sx, W).i1;     // last one
      const int wlo = xa >> 5, whi = xb >> 5;
      const uint32_t mf = 0xffffffffu << (xa & 31), ml = 0xffffffffu >> (31 - (xb & 31));
      if (win3) {
        const uint32_t m0 = whi == wlo ? (mf & ml) : mf;
        const uint32_t m1 = whi == wlo ? 0u : (whi == wlo + 1 ? ml : 0xffffffffu);
        const uint32_t m2 = whi == wlo + 2 ? ml : 0u;
        ctab[c] = make_int4(wlo, (int)m0, (int)m1, (int)m2);
      } else {
        ctab[c] = make_int4(wlo, whi, (int)mf, (int)ml);
      }
    }
  }
};

// Resize one tile whose source rows sit in shared memory (`t`, row pitch Wp words, first row = the tile's ylo; with tb.win3 the
// buffer must be readable up to 2 words past its last row — those reads are masked out).
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
  // (r, c) of this thread's word, stepped by blockDim per iteration instead of divided out every time
  const int step_r = (int)blockDim.x / owp, step_c = (int)blockDim.x - step_r * owp;
  int r = tid / owp, c = tid - r * owp;
  for (int base = 0; base < n_words; base += blockDim.x) {
    const int i = base + tid;
    const bool have = i < n_words;
    uint32_t word = 0;
    bool edge = false;
    if (have) {
      const YParam yp = tb.ytab[r];
      const int4 win = tb.ctab[c];
      const uint32_t* r0 = t + yp.y0 + win.x;
      const uint32_t* r1 = t + yp.y1 + win.x;
      uint32_t any1, all1;
      if (tb.win3) {
        const uint32_t m0 = (uint32_t)win.y, m1 = (uint32_t)win.z, m2 = (uint32_t)win.w;
        const uint32_t a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2];
        any1 = ((a0 | b0) & m0) | ((a1 | b1) & m1) | ((a2 | b2) & m2);
        all1 = ((a0 & b0) | ~m0) & ((a1 & b1) | ~m1) & ((a2 & b2) | ~m2);
      } else {
        any1 = 0u; all1 = 0xffffffffu;
        const int n = win.y - win.x;
        for (int w = 0; w <= n; ++w) {
          // only the source pixels [xa, xb] matter: bits outside are forced to "don't care" for both tests
          uint32_t care = 0xffffffffu;
          if (w == 0) care &= (uint32_t)win.z;
          if (w == n) care &= (uint32_t)win.w;
          const uint32_t v0 = r0[w], v1 = r1[w];
          any1 |= (v0 | v1) & care;
          all1 &= (v0 & v1) | ~care;
        }
      }
      if (any1 == 0u) word = 0u;
      else if (all1 == 0xffffffffu) {
        const int n_px = min(32, ow - c * 32);
        word = n_px == 32 ? 0xffffffffu : ((1u << n_px) - 1u);
      } else edge = true;
    }
    unsigned pending = __ballot_sync(FULL, edge);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const int rr = __shfl_sync(FULL, r, src), cc = __shfl_sync(FULL, c, src);
      const int ox = cc * 32 + lane;
      const XParam xp = tb.xtab[ox];
      const YParam yp = tb.ytab[rr];
      const uint32_t* r0 = t + yp.y0;
      const uint32_t* r1 = t + yp.y1;
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
      out_rows[r * owp + c] = word;
      area_acc += __popc(word);
    }
    r += step_r; c += step_c;
    if (c >= owp) { c -= owp; ++r; }
  }
  return area_acc;
}

}  // namespace sola
