// Fused J & F kernel — region counts AND boundary-match counts of (pred, gt) frame pairs from ONE staged tile of the packed planes.
//
// Per frame pair it emits the seven integers every J / F flavour on this path is a function of:
//   [0] |pred ∩ gt|   [1] |pred|   [2] |gt|                    -> Evaluator.compute_J (evaluator.py:227-237: inter / union per frame),
//                                                                 Evaluator.compute_F (evaluator.py:239-247: tp/fp/fn = Σ[0], Σ[1]-Σ[0], Σ[2]-Σ[0]),
//                                                                 utils.compute_mask_metrics (track_generation/utils.py:132-174)
//   [3] |b(pred)|  [4] |b(gt)|  [5] |b(pred) ∩ dil(b(gt))|  [6] |b(gt) ∩ dil(b(pred))|
//                                                              -> the boundary F-measure BASELINE.json's north-star names (seg2bmap + disk
//                                                                 dilation + match; no reference implementation — DAVIS definition restated
//                                                                 in oracle/boundary_oracle.py, parity unpinned)
// It replaces raw_counts / packed_counts + pack + boundary_counts_kernel (three launches, the planes read three times) in the J&F sweep.
//
// Work item = (unit, frame, row band).  A unit is one (video, expression): T frame pairs of one shape; a sweep mixes shapes (MeViS:
// 360p ... 1080p), so items are enumerated over a device table of units and a persistent grid walks them.  Per item:
//   load     the band's rows (+ r halo rows each side, + 1 row for the south neighbour) of both planes are ONE contiguous run of
//            words each: a single elected thread fetches them with two TMA bulk copies (cp.async.bulk, mbarrier completion) — no
//            registers, no per-thread address math, and the next item's copy is issued as soon as the current tile has been
//            consumed, so it overlaps the dilation phase;
//   phase 1  "column walkers": thread = (row sub-band, word column) walks DOWN its rows holding the current row's word and its
//            east-shifted copy in registers, so every boundary word costs two shared loads per plane (the row below and its
//            east neighbour), one funnel shift and four LOP3 (b = (s^e)|(s^s')|(s^se), last-row / last-column / corner rules);
//            region popcounts and boundary popcounts are taken on the way for the rows the band owns; b goes to shared memory;
//   phase 2  boundary pixels are sparse (object contours), and a word of b(pred) that is zero needs no dilation of b(gt): each warp
//            scans the owned words, compacts the non-zero ones into a per-warp queue (ballot) and dilates ONLY those, 32 items per
//            pass with all lanes busy.  The disk is decomposed by column offset k: dil = OR_k shift(±k)(vertical OR of half-height
//            v[k]); walking k from r down to 0 the vertical extent only grows, so an item costs 6r shared loads + 3r LOP3 for
//            the running ORs and 2 funnel shifts + 1 LOP3 per k.  Items whose pixels all match inside a radius-3 disk (the common
//            case when pred ≈ gt) stop after 7 rows.
// Everything is exact integer arithmetic; pad bits (x >= W) stay zero through every step.
#include "tma.cuh"
#include "csa.cuh"
#include "jf_unit.h"
#include <string.h>

namespace sola {

constexpr int JF_THREADS = 256;                // 384 / 512 threads (two CTAs per SM at 80 / 64 registers, no spills) measured 8 % / 18 % slower in
                                               // boundary mode on every shape (profiles/r3_build_constants.json)
constexpr int JF_WARPS = JF_THREADS / 32;
constexpr int JF_MAX_R = 31;
constexpr int JF_MAX_WP = 256;                 // one walker per word column: W <= 8192
constexpr int JF_QUEUE = 96;                   // < 32 left over + at most 64 pushed per scan step
#ifndef JF_PRE_R_VALUE
#define JF_PRE_R_VALUE 3
#endif
constexpr int JF_PRE_R = JF_PRE_R_VALUE;       // radius of the early-exit pre-test.  Measured 2 / 3 / 4 (profiles/r2_jf_pretest_radius_experiment.json):
                                               // object-like pairs gain 2-6 % with the larger radii, speckled ones lose 2-10 %; 3 is the compromise

// row pitch of the boundary maps: the Wp words of a row + one zero word each side, made odd so that vertically adjacent items of
// phase 2 fall into different banks
__host__ __device__ constexpr int jf_pitch(int Wp) { return (Wp + 2) | 1; }

struct JfGeo {
  const uint32_t* pred;
  const uint32_t* gt;
  long long g0, g1;          // words [g0, g1) of the unit buffers hold the rows this item reads
  long long total_words;     // T * H * Wp
  long long out_col;
  int H, W, Wp, r;           // r < 0: region counts only
  int y0, y1;                // owned rows
  int ra;                    // first staged row
  int NB;                    // boundary-map rows kept: (y1 - y0) + 2r, first one is frame row y0 - r
  bool bulk;                 // both bases 16-byte aligned -> TMA bulk copies
};

__device__ __forceinline__ void jf_decode(const sola_jf_unit* __restrict__ units, int n_units, const sola_jf_unit& single,
                                          long long item, JfGeo& g) {
  sola_jf_unit u;
  if (units) {
    int lo = 0, hi = n_units - 1;                       // last unit with item0 <= item
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(&units[mid].item0) <= item) lo = mid; else hi = mid - 1;
    }
    u = units[lo];
  } else {
    u = single;
  }
  const long long local = item - u.item0;
  const int f = (int)(local / u.n_bands), band = (int)(local - (long long)f * u.n_bands);
  g.pred = u.pred; g.gt = u.gt;
  g.H = u.H; g.W = u.W; g.Wp = (u.W + 31) >> 5; g.r = u.radius;
  const int r = u.radius < 0 ? 0 : u.radius;
  g.y0 = band * u.band_rows;
  g.y1 = min(u.H, g.y0 + u.band_rows);
  g.ra = u.radius < 0 ? g.y0 : max(0, g.y0 - r);
  const int rb = u.radius < 0 ? g.y1 : min(u.H, g.y1 + r + 1);
  g.NB = u.radius < 0 ? 0 : (g.y1 - g.y0) + 2 * r;
  const long long FW = (long long)u.H * g.Wp;
  g.g0 = (long long)f * FW + (long long)g.ra * g.Wp;
  g.g1 = (long long)f * FW + (long long)rb * g.Wp;
  g.total_words = (long long)u.T * FW;
  g.out_col = u.out_off + f;
  g.bulk = aligned16(u.pred) && aligned16(u.gt);
}

// Stage words [g0, g1) of both planes at raw[(w - (g0 & ~3))]: bulk copies over the 16-byte aligned part, plain loads for what is left.
__device__ __forceinline__ void jf_issue_load(const JfGeo& g, uint32_t* rawP, uint32_t* rawG, uint64_t* bar, int tid) {
  const long long gs = g.g0 & ~3ll;
  if (g.bulk) {
    long long bulk_end = (g.g1 + 3) & ~3ll;                               // may run up to 3 words into the next rows: harmless, in bounds
    const long long lim = g.total_words & ~3ll;                           // ... unless the buffer ends first
    if (bulk_end > lim) bulk_end = lim;
    const long long nbulk = bulk_end > gs ? bulk_end - gs : 0;
    if (tid == 0) {
      mbar_expect_tx(bar, (unsigned)(2 * nbulk * 4));
      if (nbulk > 0) {
        bulk_load_1d(rawP, g.pred + gs, (unsigned)(nbulk * 4), bar);
        bulk_load_1d(rawG, g.gt + gs, (unsigned)(nbulk * 4), bar);
      }
    }
    const long long t0 = gs + nbulk;                                      // tail: at most 3 words (plus the < 4-word buffers)
    for (long long w = t0 + tid; w < g.g1; w += JF_THREADS) {
      rawP[w - gs] = __ldg(g.pred + w);
      rawG[w - gs] = __ldg(g.gt + w);
    }
  } else {
    if (tid == 0) mbar_expect_tx(bar, 0u);                                // keep the phase protocol uniform
    for (long long w = g.g0 + tid; w < g.g1; w += JF_THREADS) {
      rawP[w - gs] = __ldg(g.pred + w);
      rawG[w - gs] = __ldg(g.gt + w);
    }
  }
}

// ---- disk dilation of one word ---------------------------------------------------------------------------------------------------
// dil(word) = OR_k shift(±k)( OR of rows within ±v[k] ), v[k] = floor(sqrt(r^2 - k^2)).  Walking k from r down to 0 the vertical extent
// only grows, so three running words (left / centre / right column) are kept and each extra row pair costs 6 shared loads + 3 LOP3;
// each k costs 2 funnel shifts + 1 LOP3.  For the radii of the usual frame sizes (360p .. 1080p: 6, 8, 9, 12, 18) and for the
// radius-3 pre-test the whole walk is unrolled at compile time (no loop control, no table reads, constant shift amounts).
__host__ __device__ constexpr int jf_isqrt(int x) {
  int v = 0;
  while ((v + 1) * (v + 1) <= x) ++v;
  return v;
}

// Hot-loop shared-memory accesses go through explicit 32-bit shared-window addresses: the generic pointers an `extern __shared__`
// array decays to cost a window-base computation (S2UR SR_CgaCtaId / ULEA ...) and 64-bit address arithmetic per access group.
template <int OFF = 0>
__device__ __forceinline__ uint32_t lds32(unsigned addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ void sts32_if(unsigned addr, uint32_t v, bool pred) {
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p st.shared.u32 [%0], %1;\n}" ::"r"(addr), "r"(v), "r"((unsigned)pred) : "memory");
}

struct DilState {
  unsigned up, dn;                           // shared byte addresses of rows -vh / +vh of the item's word
  uint32_t L, C, R, dil;
};

template <int N>
__device__ __forceinline__ void dil_grow(DilState& d, unsigned pitch_bytes) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    d.up -= pitch_bytes; d.dn += pitch_bytes;
    d.L |= lds32<-4>(d.up) | lds32<-4>(d.dn);
    d.C |= lds32<0>(d.up) | lds32<0>(d.dn);
    d.R |= lds32<4>(d.up) | lds32<4>(d.dn);
  }
}

template <int RAD, int K, int VH>
struct DilStep {
  static __device__ __forceinline__ void run(DilState& d, unsigned pitch_bytes) {
    constexpr int VK = jf_isqrt(RAD * RAD - K * K);
    dil_grow<VK - VH>(d, pitch_bytes);
    // a source pixel at x-k lands on x (shift towards higher bits, refilled from the left word) and one at x+k on x (the mirror)
    if constexpr (K > 0) {
      d.dil |= __funnelshift_l(d.L, d.C, K) | __funnelshift_r(d.C, d.R, K);
      DilStep<RAD, K - 1, VK>::run(d, pitch_bytes);
    } else {
      d.dil |= d.C;
    }
  }
};

template <int RAD>
__device__ __forceinline__ uint32_t dilate_word_fixed(unsigned src, unsigned pitch_bytes) {
  DilState d{src, src, lds32<-4>(src), lds32<0>(src), lds32<4>(src), 0u};
  DilStep<RAD, RAD, 0>::run(d, pitch_bytes);
  return d.dil;
}

__device__ __forceinline__ uint32_t dilate_word_generic(unsigned src, unsigned pitch_bytes, int r, const unsigned char* __restrict__ v) {
  DilState d{src, src, lds32<-4>(src), lds32<0>(src), lds32<4>(src), 0u};
  int vh = 0;
  for (int k = r; k >= 0; --k) {
    const int vk = v[k];
    for (; vh < vk; ++vh) dil_grow<1>(d, pitch_bytes);
    d.dil |= k ? (__funnelshift_l(d.L, d.C, k) | __funnelshift_r(d.C, d.R, k)) : d.C;
  }
  return d.dil;
}

// RAD > 0: compile-time radius; RAD == 0: run-time radius r with the table v
template <int RAD>
__device__ __forceinline__ uint32_t dilate_word(unsigned src, unsigned pitch_bytes, int r, const unsigned char* __restrict__ v) {
  if constexpr (RAD > 0) return dilate_word_fixed<RAD>(src, pitch_bytes);
  else return dilate_word_generic(src, pitch_bytes, r, v);
}

// ---- phase 2: match counting over the owned rows of one item --------------------------------------------------------------------
// Two-level compaction keeps the lanes busy: non-zero boundary words -> queue 1; a radius-3 pre-test (7 rows; a subset of the real
// disk, so a pixel it matches IS matched) settles the words whose pixels all have a partner within 3 px — nearly all of them when
// pred ≈ gt; only the rest go to queue 2 and pay for the full (2r+1)-row dilation, again 32 at a time.
template <int RAD>
__device__ __forceinline__ void jf_phase2(unsigned sF, unsigned sG /* shared byte addresses of the two boundary maps */, int BP, int r, int n_steps,
                                          int my_o0 /* this lane's boundary-map offset at step 0 */, const uint2* __restrict__ masks,
                                          const unsigned char* __restrict__ vtab, uint32_t* __restrict__ q1, uint32_t* __restrict__ q2,
                                          int lane, int& fm, int& gm) {
  constexpr bool PRE = (RAD == 0) || (RAD > JF_PRE_R);        // run-time radii <= 2 take the pre-test too: harmless, it is exact
  const unsigned lt = (1u << lane) - 1u;
  int n1 = 0, n2 = 0;
  auto full_pass = [&](bool active) {
    if (active) {
      const uint32_t e = q2[n2 + lane];
      const int eo = (int)(e & 0x7fffffffu);
      const bool sel = (e >> 31) != 0u;
      const uint32_t need = lds32((sel ? sG : sF) + 4u * eo);
      const int m = __popc(need & dilate_word<RAD>((sel ? sF : sG) + 4u * eo, 4u * BP, r, vtab));
      if (sel) gm += m; else fm += m;
    }
    __syncwarp();
  };
  auto pre_pass = [&](bool active) {
    uint32_t e = 0u;
    bool fail = false;
    if (active) {
      e = q1[n1 + lane];
      const int eo = (int)(e & 0x7fffffffu);
      const bool sel = (e >> 31) != 0u;
      const uint32_t need = lds32((sel ? sG : sF) + 4u * eo);
      if (PRE && r > JF_PRE_R) {
        const uint32_t pre = dilate_word_fixed<JF_PRE_R>((sel ? sF : sG) + 4u * eo, 4u * BP);
        fail = (need & ~pre) != 0u;
        if (!fail) { if (sel) gm += __popc(need); else fm += __popc(need); }
      } else {
        fail = true;
      }
    }
    const unsigned mfail = __ballot_sync(FULL, fail);
    if (fail) q2[n2 + __popc(mfail & lt)] = e;
    n2 += __popc(mfail);
    __syncwarp();
    if (n2 >= 32) { n2 -= 32; full_pass(true); }
  };
  // the warp walks the steps of ITS OWN walkers again: masks[k] = which lanes produced a non-zero owned boundary word at step k.
  // Consecutive steps are consecutive rows of the same columns, so a batch of 32 items is a compact patch of the contour and its
  // shared loads spread over the banks (the row pitch BP is odd).
  for (int k = 0; k < n_steps; ++k) {
    const uint2 m = masks[k];
    if ((m.x | m.y) == 0u) continue;
    const uint32_t o = (uint32_t)(my_o0 + k * BP);
    if ((m.x >> lane) & 1u) q1[n1 + __popc(m.x & lt)] = o;
    n1 += __popc(m.x);
    if ((m.y >> lane) & 1u) q1[n1 + __popc(m.y & lt)] = o | 0x80000000u;
    n1 += __popc(m.y);
    __syncwarp();
    while (n1 >= 32) { n1 -= 32; pre_pass(true); }
  }
  if (n1 > 0) { const int n = n1; n1 = 0; pre_pass(lane < n); }
  if (n2 > 0) { const int n = n2; n2 = 0; full_pass(lane < n); }
}

// BOUNDARY = false is the region-counts-only build (J and F exactly as evaluator.py:227-247 defines them): no boundary maps, queues
// or disk tables, few registers, small tiles — many resident CTAs per SM keep enough bulk copies in flight for HBM speed.
template <int MIN_CTAS, bool BOUNDARY>
__global__ void __launch_bounds__(JF_THREADS, MIN_CTAS)
jf_fused_kernel(const sola_jf_unit* __restrict__ units, int n_units, const sola_jf_unit single, long long n_items, int raw_cap,
                int bm_cap, int mask_steps, int* __restrict__ counts, long long total_frames) {
  extern __shared__ __align__(16) uint32_t jf_smem[];
  uint32_t* rawP = jf_smem;
  uint32_t* rawG = rawP + raw_cap;
  uint32_t* bmF = rawG + raw_cap;
  uint32_t* bmG = bmF + bm_cap;
  uint2* masks = reinterpret_cast<uint2*>(bmG + bm_cap);          // [JF_WARPS][mask_steps]: phase 1's ballots = phase 2's work list
  const unsigned sF = smem_u32(bmF), sG = smem_u32(bmG);
  __shared__ uint64_t bar;
  __shared__ uint32_t queue1[JF_WARPS][JF_QUEUE], queue2[JF_WARPS][64];
  __shared__ int red[7][JF_WARPS];
  __shared__ unsigned char vtab[JF_MAX_R + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  long long item = blockIdx.x;
  if (item >= n_items) return;
  JfGeo g, gn;
  jf_decode(units, n_units, single, item, g);
  jf_issue_load(g, rawP, rawG, &bar, tid);
  unsigned parity = 0;
  int r_cached = -1;

  for (; item < n_items; item += gridDim.x) {
    const long long next = item + gridDim.x;
    const bool has_next = next < n_items;
    if (has_next) jf_decode(units, n_units, single, next, gn);
    const int Wp = g.Wp, BP = jf_pitch(Wp), r = g.r < 0 ? 0 : g.r;
    const int off = (int)(g.g0 & 3ll);
    if (BOUNDARY && g.r >= 0) {
      if (r != r_cached && tid <= r) vtab[tid] = (unsigned char)jf_isqrt(r * r - tid * tid);     // disk table for run-time radii
      r_cached = r;
      for (int j = tid; j < g.NB; j += JF_THREADS) {       // zero halo columns of both boundary maps
        bmF[j * BP] = 0u; bmF[j * BP + Wp + 1] = 0u;
        bmG[j * BP] = 0u; bmG[j * BP + Wp + 1] = 0u;
      }
    }
    __syncthreads();                                        // plain-load words of the tile are visible; previous item fully retired
    mbar_wait(&bar, parity);
    parity ^= 1u;

    int n_i = 0, n_p = 0, n_g = 0, n_bf = 0, n_bg = 0, fm = 0, gm = 0;
    int n_steps = 0, my_o0 = 0;
    if (!BOUNDARY || g.r < 0) {
      // region counts only: flat pass over the owned words [off, off + n_words) of the staged tile — whole 16-byte quads with
      // 128-bit shared loads, the few words before the first / after the last whole quad one by one
      const int n_words = (g.y1 - g.y0) * Wp;
      const int q_lo = (off + 3) >> 2, q_hi = (off + n_words) >> 2;
      auto count = [&](uint32_t p, uint32_t q) { n_p += __popc(p); n_g += __popc(q); n_i += __popc(p & q); };
      if (q_hi > q_lo) {
        const uint4* P4 = reinterpret_cast<const uint4*>(rawP);
        const uint4* G4 = reinterpret_cast<const uint4*>(rawG);
        // carry-save counters (csa.cuh): 3 POPC per 16-byte quad pair instead of 12 — the xu pipe was 58 % busy (ncu); +9 % on 360p
        // frames, +2 % on the mixed sweep, neutral elsewhere (profiles/r3_build_constants.json)
        Csa cp{0u, 0u, 0}, cg{0u, 0u, 0}, ci{0u, 0u, 0};
        for (int i = q_lo + tid; i < q_hi; i += JF_THREADS) {
          const uint4 p = P4[i], q = G4[i];
          csa_add4(cp, p.x, p.y, p.z, p.w); csa_add4(cg, q.x, q.y, q.z, q.w); csa_quad(ci, p, q);
        }
        n_p += csa_total(cp); n_g += csa_total(cg); n_i += csa_total(ci);
        const int head = 4 * q_lo - off, tail = off + n_words - 4 * q_hi;          // 0..3 words each
        if (tid < head) count(rawP[off + tid], rawG[off + tid]);
        else if (tid >= 32 && tid - 32 < tail) count(rawP[4 * q_hi + tid - 32], rawG[4 * q_hi + tid - 32]);
      } else {
        for (int i = tid; i < n_words; i += JF_THREADS) count(rawP[off + i], rawG[off + i]);
      }
    } else {
      // ---- phase 1: column walkers build both boundary maps ------------------------------------------------------------------
      // Warp-uniform loop over the walker steps, per-lane predicates instead of per-lane trip counts: every step ends in two ballots
      // ("which lanes produced a non-zero boundary word in a row the band owns"), the work list of phase 2.
      const int c = tid % Wp, s = tid / Wp;
      const int n_sub = min(JF_THREADS / Wp, g.NB);
      n_steps = (g.NB + n_sub - 1) / n_sub;
      const bool walker = s < n_sub;
      const int j0 = walker ? (int)((long long)s * g.NB / n_sub) : 0, j1 = walker ? (int)((long long)(s + 1) * g.NB / n_sub) : 0;
      const int ya = g.y0 - r + j0;                                  // frame row of this walker's step 0
      const uint32_t lastbit = (c == Wp - 1) ? (1u << ((g.W - 1) & 31)) : 0u;
      const bool east = c + 1 < Wp;
      my_o0 = j0 * BP + c + 1;
      // shared byte addresses: row y of the raw tile = ap + (y - ya) * Wp * 4 (dereferenced in-frame only); boundary words at ab
      unsigned ap = smem_u32(rawP) + 4u * (unsigned)(off + (ya - g.ra) * Wp + c);
      const unsigned dPG = 4u * (unsigned)raw_cap, dFG = sG - sF, Wp4 = 4u * Wp, BP4 = 4u * BP;
      unsigned ab = sF + 4u * (unsigned)my_o0;
      // Branch-free pipeline: every step loads row clamp(y + 1) of the staged tile (always a valid address), the registers shift
      // unconditionally, and validity lives in three masks — in-frame, has-a-row-below, owned — derived from per-lane step ranges.
      const int n_act = walker ? j1 - j0 : 0;
      const int k_in0 = max(0, -ya), k_in1 = min(n_act, g.H - ya), k_s1 = min(k_in1, g.H - 1 - ya);
      const int k_o0 = max(k_in0, g.y0 - ya), k_o1 = min(k_in1, g.y1 - ya);
      const int row_hi = (int)((g.g1 - g.g0) / Wp) - 1;            // last staged row (tile-relative)
      const int row0 = ya - g.ra;                                   // tile-relative row of step 0 (may be negative / past the end: clamped)
      const uint32_t emask = east ? 0xffffffffu : 0u;
      auto row_addr = [&](int row) { return ap + (unsigned)(min(max(row, 0), row_hi) - row0) * Wp4; };
      uint32_t p0, pe0, q0, qe0;
      {
        const unsigned a0 = row_addr(row0);
        p0 = lds32(a0); q0 = lds32(a0 + dPG);
        pe0 = __funnelshift_r(p0, lds32<4>(a0) & emask, 1);
        qe0 = __funnelshift_r(q0, lds32<4>(a0 + dPG) & emask, 1);
      }
      uint2* mk = masks + warp * mask_steps;
#pragma unroll 2
      for (int k = 0; k < n_steps; ++k) {
        const uint32_t fmask = (k >= k_in0 && k < k_in1) ? 0xffffffffu : 0u;
        const uint32_t smask = (k >= k_in0 && k < k_s1) ? 0xffffffffu : 0u;
        const uint32_t omask = (k >= k_o0 && k < k_o1) ? 0xffffffffu : 0u;
        const unsigned a1 = row_addr(row0 + k + 1);
        const uint32_t p1 = lds32(a1), q1w = lds32(a1 + dPG);
        const uint32_t pe1 = __funnelshift_r(p1, lds32<4>(a1) & emask, 1);
        const uint32_t qe1 = __funnelshift_r(q1w, lds32<4>(a1 + dPG) & emask, 1);
        const uint32_t ps = p0 ^ p1, qs = q0 ^ q1w, lbs = lastbit & smask;
        // rows with a row below: (s ^ e) | (s ^ south) | (s ^ south-east), last column: s ^ south only; last row: s ^ e, corner 0
        const uint32_t bp = ((((p0 ^ pe0) | ((ps | (p0 ^ pe1)) & smask)) & ~lastbit) | (ps & lbs)) & fmask;
        const uint32_t bq = ((((q0 ^ qe0) | ((qs | (q0 ^ qe1)) & smask)) & ~lastbit) | (qs & lbs)) & fmask;
        const uint32_t pm = p0 & omask, qm = q0 & omask, bpm = bp & omask, bqm = bq & omask;
        n_p += __popc(pm); n_g += __popc(qm); n_i += __popc(pm & qm);
        n_bf += __popc(bpm); n_bg += __popc(bqm);
        sts32_if(ab, bp, k < n_act);
        sts32_if(ab + dFG, bq, k < n_act);
        ab += BP4;
        const unsigned mF = __ballot_sync(FULL, bpm != 0u), mG = __ballot_sync(FULL, bqm != 0u);
        if (lane == 0) mk[k] = make_uint2(mF, mG);
        p0 = p1; pe0 = pe1; q0 = q1w; qe0 = qe1;
      }
    }
    __syncthreads();                                        // boundary maps complete; the raw tile is dead
    if (has_next) jf_issue_load(gn, rawP, rawG, &bar, tid);  // next tile streams in while this one is matched

    if (BOUNDARY && g.r >= 0) {
      uint32_t* q1 = queue1[warp];
      uint32_t* q2 = queue2[warp];
      const uint2* mk = masks + warp * mask_steps;
      switch (r) {
        case 6: jf_phase2<6>(sF, sG, BP, r, n_steps, my_o0, mk, vtab, q1, q2, lane, fm, gm); break;       // 360 x 640
        case 8: jf_phase2<8>(sF, sG, BP, r, n_steps, my_o0, mk, vtab, q1, q2, lane, fm, gm); break;       // 480 x 854
        case 9: jf_phase2<9>(sF, sG, BP, r, n_steps, my_o0, mk, vtab, q1, q2, lane, fm, gm); break;       // 540 x 960
        case 12: jf_phase2<12>(sF, sG, BP, r, n_steps, my_o0, mk, vtab, q1, q2, lane, fm, gm); break;     // 720 x 1280
        case 18: jf_phase2<18>(sF, sG, BP, r, n_steps, my_o0, mk, vtab, q1, q2, lane, fm, gm); break;     // 1080 x 1920
        default: jf_phase2<0>(sF, sG, BP, r, n_steps, my_o0, mk, vtab, q1, q2, lane, fm, gm); break;
      }
    }

    n_i = warp_sum(n_i); n_p = warp_sum(n_p); n_g = warp_sum(n_g);
    if (BOUNDARY && g.r >= 0) { n_bf = warp_sum(n_bf); n_bg = warp_sum(n_bg); fm = warp_sum(fm); gm = warp_sum(gm); }
    if (lane == 0) {
      red[0][warp] = n_i; red[1][warp] = n_p; red[2][warp] = n_g; red[3][warp] = n_bf; red[4][warp] = n_bg; red[5][warp] = fm; red[6][warp] = gm;
    }
    __syncthreads();
    if (tid < ((BOUNDARY && g.r >= 0) ? 7 : 3)) {
      int sum = 0;
#pragma unroll
      for (int w = 0; w < JF_WARPS; ++w) sum += red[tid][w];
      if (sum) atomicAdd(counts + (long long)tid * total_frames + g.out_col, sum);
    }
    g = gn;
  }
}

// ---- host side: row-band plan ------------------------------------------------------------------------------------------------
static size_t jf_unit_smem(int Wp, int r, int rows, bool boundary, int* raw_cap, int* bm_cap, int* steps) {
  const int rr = boundary ? r : 0;
  const int rc = (((rows + (boundary ? 2 * rr + 1 : 0)) * Wp + 8) + 3) & ~3;
  const int NB = rows + 2 * rr;
  const int bc = boundary ? ((NB * jf_pitch(Wp) + 1) & ~1) : 0;             // even: the uint2 mask table behind it stays 8-byte aligned
  int n_sub = JF_THREADS / Wp;
  if (n_sub > NB) n_sub = NB;
  const int st = boundary ? (NB + n_sub - 1) / n_sub : 0;
  if (raw_cap) *raw_cap = rc;
  if (bm_cap) *bm_cap = bc;
  if (steps) *steps = st;
  return (size_t)(2 * rc + 2 * bc + 2 * JF_WARPS * st) * sizeof(uint32_t);
}

// Largest band height whose tile fits `budget` bytes; 0 if not even one row fits.
static int jf_max_band_rows(int H, int Wp, int r, bool boundary, size_t budget) {
  int lo = 0, hi = H;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (jf_unit_smem(Wp, r, mid, boundary, nullptr, nullptr, nullptr) <= budget) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Tile budgets by resident CTAs per SM.  More CTAs hide the latency-bound phases better, fewer allow taller bands (less halo
// redundancy).  The 3-CTA build is capped at 80 registers (no spills), the other takes 128.
constexpr int JF_FIRST_CLASS = 1;                 // 0: try the 3-CTA class first; 1: start at 2 CTAs per SM
constexpr size_t JF_BUDGET_3CTA = 72 * 1024;
constexpr size_t JF_BUDGET_2CTA = 100 * 1024;
constexpr size_t JF_BUDGET_1CTA = 200 * 1024;     // tall halos (1080p: r = 18)
// sweeps without a single boundary unit run the region-only build: JF_REGION_CTAS resident CTAs per SM, each with a small tile.
// Measured on the bench's mixed sweep (profiles/r3_jf_region_ctas.jsonl): 8 / 6 / 4 / 3 / 2 CTAs -> 5.27 / 6.02 / 6.19 / 6.04 / 5.01 TB/s
// (the 128-register build with two 100 KB tiles: 5.4 TB/s)
#ifndef JF_REGION_CTAS_VALUE
#define JF_REGION_CTAS_VALUE 4
#endif
constexpr int JF_REGION_CTAS = JF_REGION_CTAS_VALUE;
constexpr size_t JF_BUDGET_REGION = (size_t)(216 / JF_REGION_CTAS - 1) * 1024;

extern "C" {
typedef struct sola_jf_plan {
  long long n_items, total_frames;
  int raw_cap, bm_cap, mask_steps, reserved;
} sola_jf_plan;
}

static int jf_plan(sola_jf_unit* units, int n_units, sola_jf_plan* plan, int force_ctas = 0) {
  // one budget for the whole launch: two CTAs per SM unless a unit's tile does not fit, then one.  Measured (profiles/
  // r2_jf_fused_bench.json): two resident CTAs beat one even at 1080p, where the 100 KB tile spends 37 of 101 rows on halo (1.20 M vs
  // 1.05 M frame pairs/s); bands for three CTAs lose everywhere except where the 2-CTA tile already fits three times (360p).
  // (force_ctas = 1 / 2 / 3 pins the class: experiments, tools/jf_fused_bench.py)
  bool all_region = n_units > 0;
  for (int i = 0; i < n_units; ++i) all_region = all_region && units[i].radius < 0;
  const bool region_class = all_region && (force_ctas == 0 || force_ctas == JF_REGION_CTAS);
  const size_t budgets[3] = {region_class ? JF_BUDGET_REGION : JF_BUDGET_3CTA, JF_BUDGET_2CTA, JF_BUDGET_1CTA};
  const int ctas[3] = {region_class ? JF_REGION_CTAS : 3, 2, 1};
  int first = region_class ? 0 : JF_FIRST_CLASS, last = 2;
  if (!region_class && force_ctas >= 1 && force_ctas <= 3) first = last = 3 - force_ctas;
  for (int cls = first; cls <= last; ++cls) {
    const size_t budget = budgets[cls];
    bool retry = false;
    long long items = 0, frames = 0;
    int raw_cap = 4, bm_cap = 0, steps = 0;
    for (int i = 0; i < n_units; ++i) {
      sola_jf_unit& u = units[i];
      SOLA_REQUIRE(u.T >= 0 && u.H > 0 && u.W > 0, "jf_sweep: unit %d has a bad shape T=%d H=%d W=%d", i, u.T, u.H, u.W);
      SOLA_REQUIRE(u.T == 0 || (u.pred && u.gt), "jf_sweep: unit %d has a null plane pointer", i);
      const int Wp = (u.W + 31) >> 5;
      const bool boundary = u.radius >= 0;
      if (u.radius > JF_MAX_R || Wp > JF_MAX_WP) {
        set_error("jf_sweep: unit %d unsupported (radius %d > %d or width %d > %d)", i, u.radius, JF_MAX_R, u.W, JF_MAX_WP * 32);
        return SOLA_ERR_UNSUPPORTED;
      }
      const int bmax = jf_max_band_rows(u.H, Wp, u.radius, boundary, budget);
      if (bmax < 1) {
        if (cls < last) { retry = true; break; }
        set_error("jf_sweep: unit %d (%dx%d, radius %d) does not fit shared memory", i, u.H, u.W, u.radius);
        return SOLA_ERR_UNSUPPORTED;
      }
      u.n_bands = (u.H + bmax - 1) / bmax;
      u.band_rows = (u.H + u.n_bands - 1) / u.n_bands;
      u.n_bands = (u.H + u.band_rows - 1) / u.band_rows;
      u.item0 = items;
      u.out_off = frames;
      items += (long long)u.T * u.n_bands;
      frames += u.T;
      int rc, bc, st;
      jf_unit_smem(Wp, u.radius, u.band_rows, boundary, &rc, &bc, &st);
      if (rc > raw_cap) raw_cap = rc;
      if (bc > bm_cap) bm_cap = bc;
      if (st > steps) steps = st;
    }
    if (retry) continue;
    plan->n_items = items; plan->total_frames = frames; plan->raw_cap = raw_cap; plan->bm_cap = bm_cap; plan->mask_steps = steps;
    plan->reserved = ctas[cls];
    // small frames: the 2-CTA plan's tile already leaves room for a third CTA — take the 80-register build so that it can be resident
    if (force_ctas == 0 && !region_class && cls == 1 && (size_t)(2 * raw_cap + 2 * bm_cap + 2 * JF_WARPS * steps) * sizeof(uint32_t) <= JF_BUDGET_3CTA) plan->reserved = 3;
    return SOLA_OK;
  }
  return SOLA_ERR_UNSUPPORTED;
}

static int jf_launch(const sola_jf_unit* units_dev, int n_units, const sola_jf_unit& single, const sola_jf_plan& p, int* counts_out,
                     cudaStream_t stream) {
  if (p.total_frames <= 0) return SOLA_OK;
  SOLA_REQUIRE(counts_out, "jf_sweep: null output");
  SOLA_CUDA(cudaMemsetAsync(counts_out, 0, sizeof(int) * 7 * (size_t)p.total_frames, stream));
  if (p.n_items <= 0) return SOLA_OK;
  const size_t smem = (size_t)(2 * p.raw_cap + 2 * p.bm_cap + 2 * JF_WARPS * p.mask_steps) * sizeof(uint32_t);
  SOLA_REQUIRE(p.raw_cap > 0 && p.raw_cap % 4 == 0 && p.bm_cap >= 0 && p.bm_cap % 2 == 0 && p.mask_steps >= 0 && smem <= 220 * 1024,
               "jf_sweep: bad shared-memory plan (raw_cap %d, bm_cap %d, mask_steps %d)", p.raw_cap, p.bm_cap, p.mask_steps);
  auto launch = [&](auto kernel) -> int {
    SOLA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, JF_THREADS, smem);
    if (occ < 1) occ = 1;
    long long grid = (long long)num_sms() * occ;
    if (grid > p.n_items) grid = p.n_items;
    kernel<<<(unsigned)grid, JF_THREADS, smem, stream>>>(units_dev, n_units, single, p.n_items, p.raw_cap, p.bm_cap, p.mask_steps, counts_out,
                                                         p.total_frames);
    return check_launch("jf_fused kernel");
  };
  // the plan's tile class decides the build: 3 CTAs per SM need the 80-register build; the region class has its own lean build
  if (p.reserved == JF_REGION_CTAS) return launch(jf_fused_kernel<JF_REGION_CTAS, false>);
  return p.reserved == 3 ? launch(jf_fused_kernel<3, true>) : launch(jf_fused_kernel<2, true>);
}

}  // namespace sola

using namespace sola;

extern "C" {

// Host-only: fills band_rows / n_bands / item0 / out_off of every unit (frames are numbered in unit order) and reports the launch plan.
int sola_jf_sweep_plan(sola_jf_unit* units_host, int n_units, sola_jf_plan* plan_out) {
  SOLA_REQUIRE(n_units >= 0 && (n_units == 0 || units_host), "jf_sweep_plan: bad arguments");
  SOLA_REQUIRE(plan_out, "jf_sweep_plan: null output");
  sola_jf_plan p{0, 0, 4, 0, 0, 0};
  if (n_units > 0) {
    const int rc = jf_plan(units_host, n_units, &p, plan_out->reserved);     // reserved on input: 0 = automatic tile class
    if (rc != SOLA_OK) return rc;
  }
  *plan_out = p;
  return SOLA_OK;
}

// counts_out int32 (7, total_frames): rows = inter, |pred|, |gt|, |b(pred)|, |b(gt)|, fg_match, gt_match (rows 3..6 stay 0 for units
// planned with radius < 0).  units_dev = the planned table copied to the device; plan = what sola_jf_sweep_plan returned (HOST pointer).
int sola_jf_sweep(const sola_jf_unit* units_dev, int n_units, const sola_jf_plan* plan, int* counts_out, cudaStream_t stream) {
  SOLA_REQUIRE(n_units >= 0 && plan && plan->n_items >= 0 && plan->total_frames >= 0, "jf_sweep: bad arguments");
  if (n_units == 0 || plan->total_frames == 0) return SOLA_OK;
  SOLA_REQUIRE(units_dev, "jf_sweep: null unit table");
  sola_jf_unit none;
  memset(&none, 0, sizeof(none));
  return jf_launch(units_dev, n_units, none, *plan, counts_out, stream);
}

// One unit, no table: pred, gt (n_frames, H, Wp) -> counts_out int32 (7, n_frames).  radius < 0: region counts only.
int sola_jf_boundary_packed(const uint32_t* pred, const uint32_t* gt, long long n_frames, int H, int W, int radius, int* counts_out,
                            cudaStream_t stream) {
  SOLA_REQUIRE(n_frames >= 0 && n_frames < (1ll << 31) && H > 0 && W > 0, "jf_boundary_packed: bad shape");
  if (n_frames == 0) return SOLA_OK;
  SOLA_REQUIRE(pred && gt && counts_out, "jf_boundary_packed: null pointer");
  sola_jf_unit u;
  memset(&u, 0, sizeof(u));
  u.pred = pred; u.gt = gt; u.T = (int)n_frames; u.H = H; u.W = W; u.radius = radius;
  sola_jf_plan p{0, 0, 4, 0, 0, 0};
  const int rc = jf_plan(&u, 1, &p);
  if (rc != SOLA_OK) return rc;
  return jf_launch(nullptr, 1, u, p, counts_out, stream);
}

}  // extern "C"
