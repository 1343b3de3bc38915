// K1+R1 fused — binarise + bit-pack + stability counts AND the bilinear-resized packed planes, in ONE pass over the logits.
//
// K1 alone runs at the HBM roofline with ~45 % of the issue slots used; R1 alone is ALU-bound and barely touches HBM.
// Fusing them hides R1's integer work under K1's memory time: the freshly packed rows never leave the SM before they are
// resized.  Replaces, in one launch,
//   (out_mask_logits > 0.0).float() + torch.cat            generate_tokens_grid.py:215-224
//   PromptGenerator.get_stability_score                    prompt_generator.py:169-186
//   seg_utils.reshape_masklet                              seg_utils.py:145-160   (called at generate_tokens_grid.py:248-250)
//
// CTA = one band of R1_TR output rows of the resized plane, walked over a slice of the frames.  Per frame:
//   phase 1  stream the band's input rows [tlo, yend] (the rows its output rows read, i.e. its own rows plus a 1-2 row halo
//            shared with the next band: ~2 % redundant HBM reads) exactly like K1 — 128-bit streaming loads, sign-bit
//            predicate extraction, in-warp slot transpose — writing every finished word to shared memory, and the rows the
//            band OWNS ([tlo, next band's first row)) also to the full-resolution packed output; popcounts only over owned words;
//   phase 2  resize the shared-memory tile (resize_core.cuh: uniform-window words by one thread, edge words by one warp).
// Results are bit-identical to running K1 and R1 separately (tests/test_gpu_fused.py).
#include "pack_core.cuh"
#include "resize_core.cuh"
#include <math.h>

namespace sola {

constexpr int FU_THREADS = 256;
constexpr int FU_WARPS = FU_THREADS / 32;
constexpr int FU_CHUNK_PX = 1024;

template <typename T, int MIN_CTAS = 0>
__global__ void __launch_bounds__(FU_THREADS, MIN_CTAS)
fused_pack_resize_kernel(const T* __restrict__ logits, int n_frames, int frames_per_slice, int H, int W, int oh, int ow,
                         float sy, float sx, int max_tile_rows, Thresholds th,
                         uint32_t* __restrict__ packed, uint32_t* __restrict__ resized,
                         int* __restrict__ cnt_hi, int* __restrict__ cnt_mid, int* __restrict__ cnt_lo, int* __restrict__ area_resized) {
  constexpr int E = ElemTraits<T>::E;
  constexpr int L = 32 / E;
  extern __shared__ __align__(16) unsigned char fu_smem[];
  const int Wp = W >> 5, owp = (ow + 31) >> 5;                       // W % 32 == 0 on this path
  R1Tables tb;
  tb.carve(fu_smem, owp);
  uint32_t* tile = reinterpret_cast<uint32_t*>(fu_smem + (size_t)owp * 32 * sizeof(XParam) + R1_TR * sizeof(YParam) + (size_t)owp * sizeof(int4));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_bands = (oh + R1_TR - 1) / R1_TR;
  const int band = blockIdx.x;
  const int oy0 = band * R1_TR;
  const int nrows = min(R1_TR, oh - oy0);
  const int f_begin = blockIdx.y * frames_per_slice;
  const int f_end = min(n_frames, f_begin + frames_per_slice);

  // rows: the band reads [ylo, yhi]; it owns [tlo, own_end) where own_end is the next band's first row (H for the last band)
  const int ylo = bilinear_axis(oy0, sy, H).i0;
  const int yhi = bilinear_axis(oy0 + nrows - 1, sy, H).i1;
  const int tlo = band == 0 ? 0 : ylo;
  const int own_end = band == n_bands - 1 ? H : bilinear_axis(oy0 + R1_TR, sy, H).i0;
  const int yend = max(yhi, own_end - 1);
  const int n_tile_words = (yend - tlo + 1) * Wp;
  const int owned_words = (own_end - tlo) * Wp;
  const int n_px = (yend - tlo + 1) * W;
  const int n_chunks = (n_px + FU_CHUNK_PX - 1) / FU_CHUNK_PX;
  tb.build(oy0, tlo, H, W, oh, ow, sy, sx, Wp);
  const long long FW = (long long)H * Wp, oFW = (long long)oh * owp;
  const int out_word_in_chunk = E * (lane % L) + lane / L;
  const uint32_t slot = (E >= 32) ? 0xffffffffu : ((1u << E) - 1u);
  __shared__ int red[4][FU_WARPS];

  for (int f = f_begin; f < f_end; ++f) {
    const T* src = logits + ((long long)f * H + tlo) * W;
    uint32_t* dst = packed ? packed + f * FW + (long long)tlo * Wp : nullptr;
    int n_mid = 0, n_hi = 0, n_lo = 0;      // n_hi / n_lo in chunk_extract3's per-dtype encoding (stab_decode)
    int b_hi = 0, b_lo = 0;                 // plain counts from the ownership-boundary chunks

    auto do_chunk = [&](int c, auto full_tag) {
      constexpr bool FULLCHUNK = decltype(full_tag)::value;
      const int px0 = c * FU_CHUNK_PX;
      uint4 raw[L];
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int px = px0 + E * (j * 32 + lane);
        if (FULLCHUNK || px < n_px) {
          raw[j] = ld_stream_u4(src + px);
        } else {
          const uint32_t ninf = sizeof(T) == 4 ? 0xff800000u : 0xff80ff80u;      // -inf: passes no threshold
          raw[j] = make_uint4(ninf, ninf, ninf, ninf);
        }
      }
      uint32_t xm = 0, xh = 0, xl = 0;
      // stability / area counts cover owned words only (halo rows are counted by the band that owns them)
      const int w0 = c * 32;
      if (w0 + 32 <= owned_words) {
        chunk_extract3<T, L, false>(raw, th, xm, xh, xl, n_hi, n_lo);
        n_mid += __popc(xm);
      } else if (w0 >= owned_words) {
        int unused_hi = 0, unused_lo = 0;
        chunk_extract3<T, L, false>(raw, th, xm, xh, xl, unused_hi, unused_lo);     // halo chunk: plane bits only
      } else {
        int unused_hi = 0, unused_lo = 0;
        chunk_extract3<T, L, true>(raw, th, xm, xh, xl, unused_hi, unused_lo);
        uint32_t own = 0;
#pragma unroll
        for (int j = 0; j < L; ++j)
          if (w0 + E * j + lane / L < owned_words) own |= slot << (E * j);
        n_mid += __popc(xm & own); b_hi += __popc(xh & own); b_lo += __popc(xl & own);
      }
      const uint32_t word = transpose_slots<E>(xm, lane);
      const int wi = w0 + out_word_in_chunk;
      if (FULLCHUNK || wi < n_tile_words) tile[wi] = word;
      if (dst && wi < owned_words) dst[wi] = word;
    };
    for (int c = warp; c < n_chunks; c += FU_WARPS) {
      if ((c + 1) * FU_CHUNK_PX <= n_px) do_chunk(c, std::true_type());
      else do_chunk(c, std::false_type());
    }
    __syncthreads();                                                 // tile (and, first time round, the tables) complete

    const int n_area = resize_tile_from_smem(tile, Wp, tb, nrows, ow, resized + f * oFW + (long long)oy0 * owp);

    n_mid = warp_sum(n_mid); n_hi = warp_sum(stab_decode<T>(n_hi) + b_hi); n_lo = warp_sum(stab_decode<T>(n_lo) + b_lo);
    const int s_area = warp_sum(n_area);
    if (lane == 0) { red[0][warp] = n_mid; red[1][warp] = n_hi; red[2][warp] = n_lo; red[3][warp] = s_area; }
    __syncthreads();                                                 // also: everyone finished reading the tile
    if (tid < 4) {
      int s = 0;
#pragma unroll
      for (int w = 0; w < FU_WARPS; ++w) s += red[tid][w];
      int* out = tid == 0 ? cnt_mid : (tid == 1 ? cnt_hi : (tid == 2 ? cnt_lo : area_resized));
      if (out && s) atomicAdd(out + f, s);
    }
  }
}

// ---- any-width variant ---------------------------------------------------------------------------------------------
// When W % 32 != 0 (480x854 DAVIS / MeViS frames) a row of the packed plane is not a whole number of 1024-pixel chunks and
// rows do not start on 16-byte boundaries, so the loads cannot be row-aligned.  Instead the band's pixels are streamed as ONE
// flat, 16-byte-aligned run (starting up to E-1 elements before the band, ending up to E-1 after it) into flat 32-pixel words
// in shared memory, and a second, cheap pass re-cuts the flat bit string into row-padded words with one funnel shift each.
// Counts are taken on the flat words with range masks on the first / last / ownership-boundary chunks.
// RESIZE = false gives the stand-alone K1 for these shapes (bands of 32 input rows, everything owned).
constexpr int BAND_ROWS = 32;

template <typename T, bool RESIZE>
__global__ void __launch_bounds__(FU_THREADS)
band_pack_generic_kernel(const T* __restrict__ logits, int n_frames, int frames_per_slice, int H, int W, int oh, int ow,
                         float sy, float sx, int max_tile_rows, int max_flat_words, Thresholds th,
                         uint32_t* __restrict__ packed, uint32_t* __restrict__ resized,
                         int* __restrict__ cnt_hi, int* __restrict__ cnt_mid, int* __restrict__ cnt_lo, int* __restrict__ area_resized) {
  constexpr int E = ElemTraits<T>::E;
  constexpr int L = 32 / E;
  extern __shared__ __align__(16) unsigned char gb_smem[];
  const int Wp = (W + 31) >> 5, owp = RESIZE ? (ow + 31) >> 5 : 0;
  R1Tables tb;
  unsigned char* cursor = gb_smem;
  if (RESIZE) {
    tb.carve(cursor, owp);
    cursor += (size_t)owp * 32 * sizeof(XParam) + R1_TR * sizeof(YParam) + (size_t)owp * sizeof(int4);
  }
  uint32_t* flat = reinterpret_cast<uint32_t*>(cursor);                       // [max_flat_words]
  uint32_t* tile = flat + max_flat_words;                                     // [max_tile_rows * Wp] (RESIZE only)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int band = blockIdx.x;
  const int f_begin = blockIdx.y * frames_per_slice;
  const int f_end = min(n_frames, f_begin + frames_per_slice);

  int tlo, yend, own_end, oy0 = 0, nrows = 0;
  if (RESIZE) {
    const int n_bands = (oh + R1_TR - 1) / R1_TR;
    oy0 = band * R1_TR;
    nrows = min(R1_TR, oh - oy0);
    const int ylo = bilinear_axis(oy0, sy, H).i0;
    const int yhi = bilinear_axis(oy0 + nrows - 1, sy, H).i1;
    tlo = band == 0 ? 0 : ylo;
    own_end = band == n_bands - 1 ? H : bilinear_axis(oy0 + R1_TR, sy, H).i0;
    yend = max(yhi, own_end - 1);
    tb.build(oy0, tlo, H, W, oh, ow, sy, sx, Wp);
  } else {
    tlo = band * BAND_ROWS;
    own_end = min(H, tlo + BAND_ROWS);
    yend = own_end - 1;
  }
  const int n_tile_rows = yend - tlo + 1;
  const int owned_rows = own_end - tlo;
  const long long FW = (long long)H * Wp, oFW = RESIZE ? (long long)oh * owp : 0;
  const int out_word_in_chunk = E * (lane % L) + lane / L;
  __shared__ int red[4][FU_WARPS];

  for (int f = f_begin; f < f_end; ++f) {
    const long long g0 = ((long long)f * H + tlo) * W;                        // first pixel of the tile (global element index)
    const long long g1 = g0 + (long long)n_tile_rows * W;
    const long long s0 = g0 - (g0 % E);                                       // 16-byte aligned start of the flat run
    const int off0 = (int)(g0 - s0);                                          // tile starts at this bit of the flat string
    const int n_flat_px = (int)(((g1 + E - 1) / E) * E - s0);                 // multiple of E; in-bounds (total elements % E == 0)
    const int n_flat_words = (n_flat_px + 31) >> 5;
    const int n_chunks = (n_flat_px + FU_CHUNK_PX - 1) / FU_CHUNK_PX;
    const int cnt_lo_px = off0, cnt_hi_px = off0 + owned_rows * W;             // pixels whose counts belong to this band
    const T* src = logits + s0;
    int n_mid = 0, n_hi = 0, n_lo = 0, b_hi = 0, b_lo = 0;      // as in the fused kernel: encoded / plain stability counts
    for (int c = warp; c < n_chunks; c += FU_WARPS) {
      const int px0 = c * FU_CHUNK_PX;
      uint4 raw[L];
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int px = px0 + E * (j * 32 + lane);
        if (px < n_flat_px) {
          raw[j] = ld_stream_u4(src + px);
        } else {
          const uint32_t ninf = sizeof(T) == 4 ? 0xff800000u : 0xff80ff80u;
          raw[j] = make_uint4(ninf, ninf, ninf, ninf);
        }
      }
      uint32_t xm = 0, xh = 0, xl = 0;
      if (px0 >= cnt_lo_px && px0 + FU_CHUNK_PX <= cnt_hi_px) {
        chunk_extract3<T, L, false>(raw, th, xm, xh, xl, n_hi, n_lo);
        n_mid += __popc(xm);
      } else if (!(px0 < cnt_hi_px && px0 + FU_CHUNK_PX > cnt_lo_px)) {
        int unused_hi = 0, unused_lo = 0;
        chunk_extract3<T, L, false>(raw, th, xm, xh, xl, unused_hi, unused_lo);     // halo chunk: plane bits only
      } else {
        int unused_hi = 0, unused_lo = 0;
        chunk_extract3<T, L, true>(raw, th, xm, xh, xl, unused_hi, unused_lo);
        uint32_t own = 0;
#pragma unroll
        for (int j = 0; j < L; ++j) {
          const int p = px0 + E * (j * 32 + lane);                            // first pixel of this vector
          const int a = min(max(cnt_lo_px - p, 0), E), b = min(max(cnt_hi_px - p, 0), E);
          if (b > a) own |= ((b - a >= 32 ? 0xffffffffu : ((1u << (b - a)) - 1u)) << a) << (E * j);
        }
        n_mid += __popc(xm & own); b_hi += __popc(xh & own); b_lo += __popc(xl & own);
      }
      const uint32_t word = transpose_slots<E>(xm, lane);
      const int k = (px0 >> 5) + out_word_in_chunk;
      if (k < n_flat_words) flat[k] = word;
    }
    if (tid < 2) flat[n_flat_words + tid] = 0u;                               // the re-cut may read up to two words past the end (pad bits only)
    __syncthreads();

    // re-cut the flat bit string into row-padded words
    uint32_t* dst = packed ? packed + f * FW + (long long)tlo * Wp : nullptr;
    for (int i = tid; i < n_tile_rows * Wp; i += FU_THREADS) {
      const int r = i / Wp, w = i - r * Wp;
      const int o = off0 + r * W + 32 * w;
      uint32_t v = __funnelshift_r(flat[o >> 5], flat[(o >> 5) + 1], o & 31);
      const int nvalid = W - 32 * w;
      if (nvalid < 32) v &= (1u << nvalid) - 1u;
      if (RESIZE) tile[i] = v;
      if (dst && r < owned_rows) dst[i] = v;
    }
    int n_area = 0;
    if (RESIZE) {
      __syncthreads();
      n_area = resize_tile_from_smem(tile, Wp, tb, nrows, ow, resized + f * oFW + (long long)oy0 * owp);
    }
    n_mid = warp_sum(n_mid); n_hi = warp_sum(stab_decode<T>(n_hi) + b_hi); n_lo = warp_sum(stab_decode<T>(n_lo) + b_lo);
    const int s_area = RESIZE ? warp_sum(n_area) : 0;
    if (lane == 0) { red[0][warp] = n_mid; red[1][warp] = n_hi; red[2][warp] = n_lo; red[3][warp] = s_area; }
    __syncthreads();                                                           // also: flat / tile free for the next frame
    if (tid < 4) {
      int s = 0;
#pragma unroll
      for (int w = 0; w < FU_WARPS; ++w) s += red[tid][w];
      int* out = tid == 0 ? cnt_mid : (tid == 1 ? cnt_hi : (tid == 2 ? cnt_lo : area_resized));
      if (out && s) atomicAdd(out + f, s);
    }
  }
}

struct FusedPlan {
  bool ok, generic;
  int max_tile_rows, max_flat_words, n_bands, slices, frames_per_slice;
  size_t smem;
};

// The fp32 kernel is built with a 5-CTA/SM register cap (48 registers, no spills): left alone ptxas takes 64 registers (4 CTAs/SM),
// which measured 1.1 % slower on the bench (2.968 vs 2.935 ms, profiles/r1_fused_prefetch_waves_experiment.log).
template <typename T> struct FusedBuild { static constexpr int MIN_CTAS = 0; };
template <> struct FusedBuild<float> { static constexpr int MIN_CTAS = 5; };

static FusedPlan plan_fused(const void* base, long long n_frames, int H, int W, int oh, int ow, int elem_size) {
  FusedPlan p{};
  p.ok = false;
  const int E = 16 / elem_size;
  if (!aligned16(base) || n_frames <= 0 || n_frames >= (1ll << 31)) return p;
  p.generic = (W % 32 != 0);
  if (p.generic && (n_frames * H * W) % E != 0) return p;               // the flat run must end on a vector boundary inside the buffer
  const float sy = (float)H / (float)oh;
  const int Wp = (W + 31) >> 5, owp = (ow + 31) >> 5;
  p.n_bands = (oh + R1_TR - 1) / R1_TR;
  // rows of the largest band tile, from the same fp32 index arithmetic the kernel uses (fmaf is the correctly rounded fma)
  auto i0 = [&](int dst) {
    float src = fmaf((float)dst + 0.5f, sy, -0.5f);
    if (!(src >= 0.f)) src = 0.f;
    return (int)src;
  };
  long long rows = 1;
  for (int b = 0; b < p.n_bands; ++b) {
    const int oy0 = b * R1_TR, nrows = (oh - oy0 < R1_TR) ? oh - oy0 : R1_TR;
    const int ylo = i0(oy0);
    int yhi = i0(oy0 + nrows - 1);
    yhi += (yhi < H - 1) ? 1 : 0;
    const int tlo = b == 0 ? 0 : ylo;
    const int own_end = b == p.n_bands - 1 ? H : i0(oy0 + R1_TR);
    const int yend = yhi > own_end - 1 ? yhi : own_end - 1;
    if (yend - tlo + 1 > rows) rows = yend - tlo + 1;
  }
  p.max_tile_rows = (int)rows;
  p.max_flat_words = (int)((rows * W + 2 * E + 31) / 32) + 2;
  p.smem = (size_t)owp * 32 * sizeof(XParam) + R1_TR * sizeof(YParam) + (size_t)owp * sizeof(int4) + (size_t)rows * Wp * sizeof(uint32_t)
           + 16;                                                         // phase A of R1 may read 2 (masked-out) words past the tile
  if (p.generic) p.smem += (size_t)p.max_flat_words * sizeof(uint32_t);
  if (p.smem > 160 * 1024) return p;
  int occ = 4;
  if (p.generic) {
    if (elem_size == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, band_pack_generic_kernel<float, true>, FU_THREADS, p.smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, band_pack_generic_kernel<__nv_bfloat16, true>, FU_THREADS, p.smem);
  } else if (elem_size == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_pack_resize_kernel<float, FusedBuild<float>::MIN_CTAS>, FU_THREADS, p.smem);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_pack_resize_kernel<__nv_bfloat16, FusedBuild<__nv_bfloat16>::MIN_CTAS>, FU_THREADS, p.smem);
  if (occ < 1) occ = 1;
  // several waves of resident CTAs: finer slices cost a table rebuild per CTA (~5 % of one frame's work) but shrink the
  // tail where the last CTAs run alone (12 and 24 waves measured 0.7-1 % slower than 16)
  const int waves = 16;
  const int target_ctas = num_sms() * occ * waves;
  int slices = target_ctas / p.n_bands;
  if (slices < 1) slices = 1;
  if (slices > n_frames) slices = (int)n_frames;
  if (slices > 65535) slices = 65535;
  p.frames_per_slice = (int)((n_frames + slices - 1) / slices);
  p.slices = (int)((n_frames + p.frames_per_slice - 1) / p.frames_per_slice);
  p.ok = true;
  return p;
}

}  // namespace sola

using namespace sola;

extern "C" {
int sola_binarize_pack_f32(const float*, long long, int, int, double, double, uint32_t*, int*, int*, int*, cudaStream_t);
int sola_binarize_pack_bf16(const void*, long long, int, int, double, double, uint32_t*, int*, int*, int*, cudaStream_t);
int sola_resize_bilinear_bin_packed(const uint32_t*, long long, int, int, int, int, uint32_t*, int*, cudaStream_t);
}

template <typename T>
static int launch_fused(const T* logits, long long n_frames, int H, int W, int oh, int ow, double thr, double off,
                        uint32_t* packed_out, uint32_t* resized_out, int* cnt_hi, int* cnt_mid, int* cnt_lo, int* area_resized,
                        cudaStream_t stream) {
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "binarize_pack_resize: bad shape");
  if (n_frames == 0) return SOLA_OK;
  SOLA_REQUIRE(logits && resized_out, "binarize_pack_resize: null pointer");
  const FusedPlan p = plan_fused(logits, n_frames, H, W, oh, ow, (int)sizeof(T));
  if (!p.ok) {
    // shapes the fused kernel does not cover (W % 32 != 0, unaligned base, huge downscale): same results from the two kernels
    SOLA_REQUIRE(packed_out != nullptr, "binarize_pack_resize: packed_out is required when the unfused path is taken");
    int rc = sizeof(T) == 4
                 ? sola_binarize_pack_f32(reinterpret_cast<const float*>(logits), n_frames, H, W, thr, off, packed_out, cnt_hi, cnt_mid, cnt_lo, stream)
                 : sola_binarize_pack_bf16(logits, n_frames, H, W, thr, off, packed_out, cnt_hi, cnt_mid, cnt_lo, stream);
    if (rc != SOLA_OK) return rc;
    return sola_resize_bilinear_bin_packed(packed_out, n_frames, H, W, oh, ow, resized_out, area_resized, stream);
  }
  for (int* c : {cnt_hi, cnt_mid, cnt_lo, area_resized})
    if (c) SOLA_CUDA(cudaMemsetAsync(c, 0, sizeof(int) * n_frames, stream));
  const Thresholds th = make_thresholds(thr, off);
  dim3 grid(p.n_bands, p.slices);
  if (p.generic) {
    SOLA_CUDA(cudaFuncSetAttribute(band_pack_generic_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    band_pack_generic_kernel<T, true><<<grid, FU_THREADS, p.smem, stream>>>(logits, (int)n_frames, p.frames_per_slice, H, W, oh, ow,
                                                                            (float)H / (float)oh, (float)W / (float)ow, p.max_tile_rows,
                                                                            p.max_flat_words, th, packed_out, resized_out, cnt_hi, cnt_mid,
                                                                            cnt_lo, area_resized);
    return check_launch("band_pack_generic<resize> kernel");
  }
  constexpr int MIN_CTAS = FusedBuild<T>::MIN_CTAS;
  SOLA_CUDA(cudaFuncSetAttribute(fused_pack_resize_kernel<T, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  fused_pack_resize_kernel<T, MIN_CTAS><<<grid, FU_THREADS, p.smem, stream>>>(logits, (int)n_frames, p.frames_per_slice, H, W, oh, ow,
                                                                             (float)H / (float)oh, (float)W / (float)ow, p.max_tile_rows, th,
                                                                             packed_out, resized_out, cnt_hi, cnt_mid, cnt_lo, area_resized);
  return check_launch("fused_pack_resize kernel");
}

namespace sola {
// Stand-alone K1 for W % 32 != 0 (called from binarize_pack.cu): bands of 32 input rows, flat run + re-cut, no resize.
template <typename T>
int launch_band_pack(const T* in, long long n_frames, int H, int W, Thresholds th, uint32_t* packed, int* cnt_hi, int* cnt_mid, int* cnt_lo,
                     cudaStream_t stream) {
  constexpr int E = ElemTraits<T>::E;
  const int n_bands = (H + BAND_ROWS - 1) / BAND_ROWS;
  const int max_flat_words = (BAND_ROWS * W + 2 * E + 31) / 32 + 2;
  const size_t smem = (size_t)max_flat_words * sizeof(uint32_t);
  if (smem > 160 * 1024) return SOLA_ERR_UNSUPPORTED;
  SOLA_CUDA(cudaFuncSetAttribute(band_pack_generic_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 4;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, band_pack_generic_kernel<T, false>, FU_THREADS, smem);
  if (occ < 1) occ = 1;
  int slices = num_sms() * occ * 16 / n_bands;
  if (slices < 1) slices = 1;
  if (slices > n_frames) slices = (int)n_frames;
  if (slices > 65535) slices = 65535;
  const int frames_per_slice = (int)((n_frames + slices - 1) / slices);
  slices = (int)((n_frames + frames_per_slice - 1) / frames_per_slice);
  dim3 grid(n_bands, slices);
  band_pack_generic_kernel<T, false><<<grid, FU_THREADS, smem, stream>>>(in, (int)n_frames, frames_per_slice, H, W, 0, 0, 1.f, 1.f, BAND_ROWS,
                                                                         max_flat_words, th, packed, nullptr, cnt_hi, cnt_mid, cnt_lo, nullptr);
  return check_launch("band_pack_generic kernel");
}
template int launch_band_pack<float>(const float*, long long, int, int, Thresholds, uint32_t*, int*, int*, int*, cudaStream_t);
template int launch_band_pack<__nv_bfloat16>(const __nv_bfloat16*, long long, int, int, Thresholds, uint32_t*, int*, int*, int*, cudaStream_t);
}  // namespace sola

extern "C" {

int sola_binarize_pack_resize_f32(const float* logits, long long n_frames, int H, int W, int oh, int ow, double thr, double off,
                                  uint32_t* packed_out, uint32_t* resized_out, int* cnt_hi, int* cnt_mid, int* cnt_lo,
                                  int* area_resized, cudaStream_t stream) {
  return launch_fused<float>(logits, n_frames, H, W, oh, ow, thr, off, packed_out, resized_out, cnt_hi, cnt_mid, cnt_lo, area_resized, stream);
}

int sola_binarize_pack_resize_bf16(const void* logits, long long n_frames, int H, int W, int oh, int ow, double thr, double off,
                                   uint32_t* packed_out, uint32_t* resized_out, int* cnt_hi, int* cnt_mid, int* cnt_lo,
                                   int* area_resized, cudaStream_t stream) {
  return launch_fused<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(logits), n_frames, H, W, oh, ow, thr, off, packed_out,
                                     resized_out, cnt_hi, cnt_mid, cnt_lo, area_resized, stream);
}

}  // extern "C"
