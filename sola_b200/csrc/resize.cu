// R1 / R2 — the two resizes in front of the greedy filter, producing bit-packed planes.
//
// R1  seg_utils.reshape_masklet (track_generation/seg_utils.py:145-160):
//       F.interpolate(masklet[None], (oh, ow), mode='bilinear') > 0.5      (align_corners=False, no antialias)
//     The reference runs this on CUDA tensors, so the arithmetic reproduced here is ATen's CUDA kernel
//     `upsample_bilinear2d_out_frame<float,float>` as compiled into torch's sm_100 cubin (read from its SASS):
//       scale   = (float)in / out                              (host, correctly rounded fp32 division)
//       src     = fma(dst + 0.5f, scale, -0.5f);  src < 0 -> 0
//       i0      = (int)src;  i1 = i0 + (i0 < in-1);  l1 = src - i0;  l0 = 1 - l1
//       top     = fma(w0, v00, w1 * v01);   bot = fma(w0, v10, w1 * v11)
//       val     = fma(h0, top, h1 * bot)
//     With 0/1 inputs ~1-3 % of output pixels land within 1e-6 of 0.5, so this order is part of
//     the contract: kept-track sets flip otherwise.  tests/test_gpu_resize.py checks bit-equality with torch-CUDA.
// R2  F.interpolate(prompt[None,None], (h, w), mode='nearest') (generate_tokens_grid.py:271-272):
//       src = min((int)floorf(dst * scale), in - 1), scale = (float)in / out.
#include "resize_core.cuh"
#include <math.h>

namespace sola {

// ---- R1 from packed planes -----------------------------------------------------------------------------------
// CTA = one tile of R1_TR output rows x the full output width, walked over a slice of the frames.
//   * per-CTA tables in shared memory, built once: per output pixel (x0, x1, w0, w1), per output row (y0, y1, h0, h1);
//   * per frame the contiguous block of input rows the tile needs is staged global -> shared with cp.async
//     (double-buffered: frame k+1 lands while frame k is computed);
//   * phase A, one THREAD per output word: OR / AND of the source words of both rows; an all-0 or all-1 window
//     resolves the whole word (background / interior) with ~1 instruction per output pixel-row of the warp;
//   * phase B, one WARP per remaining (edge) word, lane = output pixel: ATen's exact fma sequence on the 4 bits.
// Output words are written by consecutive threads -> coalesced.
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async16_ca(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}

__global__ void __launch_bounds__(R1_THREADS)
resize_bilinear_tiled_kernel(const uint32_t* __restrict__ in, int n_frames, int frames_per_slice, int H, int W, int oh, int ow,
                             float sy, float sx, int max_in_rows, uint32_t* __restrict__ out, int* __restrict__ area) {
  extern __shared__ __align__(16) unsigned char r1_smem[];
  const int Wp = (W + 31) >> 5, owp = (ow + 31) >> 5;
  R1Tables tb;
  tb.carve(r1_smem, owp);
  uint32_t* tile = reinterpret_cast<uint32_t*>(r1_smem + R1Tables::bytes(owp));   // [2][max_in_rows * Wp]
  const int tile_words_max = max_in_rows * Wp + 4;       // + 4 words: phase A may read 2 (masked-out) words past the last row, and those
                                                         //   must not alias the other buffer while cp.async is filling it
  const int tid = threadIdx.x, lane = tid & 31;
  const int oy0 = blockIdx.x * R1_TR;
  const int nrows = min(R1_TR, oh - oy0);
  const int f_begin = blockIdx.y * frames_per_slice;
  const int f_end = min(n_frames, f_begin + frames_per_slice);

  const int ylo = bilinear_axis(oy0, sy, H).i0;
  const int yhi = bilinear_axis(oy0 + nrows - 1, sy, H).i1;
  const int n_in_words = (yhi - ylo + 1) * Wp;
  tb.build(oy0, ylo, H, W, oh, ow, sy, sx, Wp);
  const long long FW = (long long)H * Wp, oFW = (long long)oh * owp;
  const bool vec16 = ((Wp & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);   // every row block starts 16-byte aligned

  auto prefetch = [&](int f, int buf) {
    const uint32_t* src = in + f * FW + (long long)ylo * Wp;
    uint32_t* dst = tile + buf * tile_words_max;
    if (vec16) {
      for (int i = tid * 4; i < n_in_words; i += R1_THREADS * 4) cp_async16_ca(dst + i, src + i);
    } else {
      for (int i = tid; i < n_in_words; i += R1_THREADS) cp_async4(dst + i, src + i);
    }
    asm volatile("cp.async.commit_group;");
  };

  if (f_begin < f_end) prefetch(f_begin, 0);
  for (int f = f_begin; f < f_end; ++f) {
    const int buf = (f - f_begin) & 1;
    if (f + 1 < f_end) {
      prefetch(f + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;");
    } else {
      asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();                                   // tile of frame f (and, first time round, the tables) visible
    const int area_acc = resize_tile_from_smem(tile + buf * tile_words_max, Wp, tb, nrows, ow, out + f * oFW + (long long)oy0 * owp);
    if (area) {
      __shared__ int red[R1_THREADS / 32];
      const int s = warp_sum(area_acc);
      if (lane == 0) red[tid >> 5] = s;
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < R1_THREADS / 32; ++w) tot += red[w];
        if (tot) atomicAdd(area + f, tot);
      }
    }
    __syncthreads();                                   // everyone done with `buf` before frame f+2 overwrites it
  }
}

// Generic fallback (any scale factor): one warp per output word walking the frames, geometry in registers.
constexpr int R1_WARPS = 8;

__global__ void __launch_bounds__(R1_WARPS * 32)
resize_bilinear_packed_kernel(const uint32_t* __restrict__ in, int n_frames, int frames_per_slice, int H, int W, int oh, int ow,
                              float sy, float sx, uint32_t* __restrict__ out, int* __restrict__ area) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Wp = (W + 31) >> 5, owp = (ow + 31) >> 5;
  const int word_id = blockIdx.x * R1_WARPS + warp;           // over oh * owp
  const int f_begin = blockIdx.y * frames_per_slice;
  const int f_end = min(n_frames, f_begin + frames_per_slice);
  if (word_id >= oh * owp) return;
  const int oy = word_id / owp, owx = word_id - oy * owp;
  const int ox = owx * 32 + lane;
  const bool live = ox < ow;
  const Axis ay = bilinear_axis(oy, sy, H);
  const Axis ax = bilinear_axis(live ? ox : ow - 1, sx, W);
  const long long FW = (long long)H * Wp, oFW = (long long)oh * owp;
  for (int f = f_begin; f < f_end; ++f) {
    const uint32_t* r0 = in + f * FW + (long long)ay.i0 * Wp;
    const uint32_t* r1 = in + f * FW + (long long)ay.i1 * Wp;
    const float v00 = (float)get_bit(r0, ax.i0), v01 = (float)get_bit(r0, ax.i1);
    const float v10 = (float)get_bit(r1, ax.i0), v11 = (float)get_bit(r1, ax.i1);
    const float val = bilinear_val(ax, ay, v00, v01, v10, v11);
    const uint32_t word = __ballot_sync(FULL, live && val > 0.5f);
    if (lane == 0) {
      out[f * oFW + word_id] = word;
      if (area && word) atomicAdd(area + f, __popc(word));
    }
  }
}

// ---- R1 from fp32 planes (general float input; drop-in for reshape_masklet on arbitrary tensors) --------------
__global__ void __launch_bounds__(256)
resize_bilinear_f32_kernel(const float* __restrict__ in, long long n_frames, int H, int W, int oh, int ow, float sy, float sx,
                           uint32_t* __restrict__ out_packed, float* __restrict__ out_f32, int* __restrict__ area) {
  const int lane = threadIdx.x & 31;
  const int owp = (ow + 31) >> 5;
  const long long words_per_frame = (long long)oh * owp;
  const long long total = n_frames * words_per_frame;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long wi = warp0; wi < total; wi += n_warps) {
    const long long f = wi / words_per_frame;
    const int rem = (int)(wi - f * words_per_frame);
    const int oy = rem / owp, owx = rem - oy * owp;
    const int ox = owx * 32 + lane;
    const bool live = ox < ow;
    bool bit = false;
    if (live) {
      const Axis ay = bilinear_axis(oy, sy, H);
      const Axis ax = bilinear_axis(ox, sx, W);
      const float* p0 = in + (f * H + ay.i0) * (long long)W;
      const float* p1 = in + (f * H + ay.i1) * (long long)W;
      const float val = bilinear_val(ax, ay, __ldg(p0 + ax.i0), __ldg(p0 + ax.i1), __ldg(p1 + ax.i0), __ldg(p1 + ax.i1));
      bit = val > 0.5f;
      if (out_f32) out_f32[(f * oh + oy) * (long long)ow + ox] = bit ? 1.f : 0.f;
    }
    const uint32_t word = __ballot_sync(FULL, bit);
    if (lane == 0) {
      if (out_packed) out_packed[wi] = word;
      if (area && word) atomicAdd(area + f, __popc(word));
    }
  }
}

// ---- R2 nearest ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  return min((int)floorf(__fmul_rn((float)dst, scale)), in_size - 1);
}

constexpr int NN_ROWS = 4;   // output rows per warp: 4 independent loads in flight per lane

template <bool PACKED_IN>
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const void* __restrict__ in_, int H, int W, int oh, int ow, float sy, float sx,
                      uint32_t* __restrict__ out_packed, int* __restrict__ area) {
  // grid: x = CTAs of 8 warps striding over the (row group x word column) items of one output plane, y = plane; 32-bit index math
  // only.  The kernel is issue-bound (ncu: 72 % issue active at ~17 instructions per 32-pixel output word).
  const int lane = threadIdx.x & 31;
  const int owp = (ow + 31) >> 5, Wp = (W + 31) >> 5;
  const int row_groups = (oh + NN_ROWS - 1) / NN_ROWS;
  const long long f = blockIdx.y;
  int pop = 0;
  for (int item = blockIdx.x * 8 + (threadIdx.x >> 5); item < row_groups * owp; item += gridDim.x * 8) {
    const int rg = item / owp, owx = item - rg * owp;
    const int ox = owx * 32 + lane;
    const bool live = ox < ow;
    const int x = nearest_src(live ? ox : 0, sx, W);
    bool bit[NN_ROWS];
#pragma unroll
    for (int k = 0; k < NN_ROWS; ++k) {
      const int oy = rg * NN_ROWS + k;
      bit[k] = false;
      if (live && oy < oh) {
        const int y = nearest_src(oy, sy, H);
        if (PACKED_IN) bit[k] = get_bit(reinterpret_cast<const uint32_t*>(in_) + (f * H + y) * Wp, x) != 0u;
        else bit[k] = __ldg(reinterpret_cast<const uint8_t*>(in_) + (f * H + y) * W + x) != 0;
      }
    }
#pragma unroll
    for (int k = 0; k < NN_ROWS; ++k) {
      const int oy = rg * NN_ROWS + k;
      const uint32_t word = __ballot_sync(FULL, bit[k]);
      if (lane == 0 && oy < oh) {
        out_packed[f * (long long)(oh * owp) + oy * owp + owx] = word;
        pop += __popc(word);
      }
    }
  }
  if (lane == 0 && area && pop) atomicAdd(area + f, pop);
}

template <bool PACKED_IN>
static int launch_nearest(const void* in, long long n_frames, int H, int W, int oh, int ow, uint32_t* out_packed, int* area, cudaStream_t stream) {
  const int owp = (ow + 31) >> 5;
  const float sy = (float)H / (float)oh, sx = (float)W / (float)ow;
  const long long gx_all = (((oh + NN_ROWS - 1) / NN_ROWS) * (long long)owp + 7) / 8;       // one item per warp
  for (long long f0 = 0; f0 < n_frames; f0 += 65535) {
    const long long nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
    long long gx_cap = ((long long)num_sms() * 16 + nf - 1) / nf;                             // ~16 CTAs per SM over the whole grid
    if (gx_cap < 1) gx_cap = 1;
    const unsigned gx = (unsigned)(gx_all < gx_cap ? gx_all : gx_cap);
    const char* src = reinterpret_cast<const char*>(in) + (PACKED_IN ? f0 * H * ((W + 31) >> 5) * 4 : f0 * H * (long long)W);
    resize_nearest_kernel<PACKED_IN><<<dim3(gx, (unsigned)nf), 256, 0, stream>>>(src, H, W, oh, ow, sy, sx, out_packed + f0 * oh * owp,
                                                                                area ? area + f0 : nullptr);
    int rc = check_launch("resize_nearest kernel");
    if (rc != SOLA_OK) return rc;
  }
  return SOLA_OK;
}

static inline float host_scale(int in_size, int out_size) { return (float)in_size / (float)out_size; }

static int grid_for_warps(long long warps) {
  long long blocks = (warps + 7) / 8;
  const long long cap = 1ll << 20;                   // one warp per output word up to ~8M words, grid-stride beyond
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace sola

using namespace sola;

extern "C" {

int sola_resize_bilinear_bin_packed(const uint32_t* in_packed, long long n_frames, int H, int W, int oh, int ow,
                                    uint32_t* out_packed, int* area, cudaStream_t stream) {
  SOLA_REQUIRE(in_packed && out_packed, "resize_bilinear_bin_packed: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && n_frames < (1ll << 31) && H > 0 && W > 0 && oh > 0 && ow > 0, "resize_bilinear_bin_packed: bad shape");
  if (n_frames == 0) return SOLA_OK;
  if (area) SOLA_CUDA(cudaMemsetAsync(area, 0, sizeof(int) * n_frames, stream));
  const int Wp = (W + 31) >> 5, owp = (ow + 31) >> 5;
  const float sy = host_scale(H, oh), sx = host_scale(W, ow);
  // input rows one tile of R1_TR output rows can touch (+2 for the y1 row and rounding)
  long long max_in_rows = (long long)ceil((double)R1_TR * (double)sy) + 3;
  if (max_in_rows > H) max_in_rows = H;
  const size_t smem = (size_t)owp * 32 * sizeof(XParam) + R1_TR * sizeof(YParam) + (size_t)owp * sizeof(int4) +
                      2 * ((size_t)max_in_rows * Wp + 4) * sizeof(uint32_t);  // == R1Tables::bytes(owp) + double-buffered tile, each
                                                                              //    buffer padded by 4 words (see the kernel)
  if (smem <= 200 * 1024) {
    SOLA_CUDA(cudaFuncSetAttribute(resize_bilinear_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (oh + R1_TR - 1) / R1_TR;
    int slices = (num_sms() * 16 + tiles - 1) / tiles;            // ~16 CTAs per SM in flight over the launch
    if (slices > n_frames) slices = (int)n_frames;
    if (slices > 65535) slices = 65535;
    if (slices < 1) slices = 1;
    const int frames_per_slice = (int)((n_frames + slices - 1) / slices);
    slices = (int)((n_frames + frames_per_slice - 1) / frames_per_slice);
    dim3 grid(tiles, slices);
    resize_bilinear_tiled_kernel<<<grid, R1_THREADS, smem, stream>>>(in_packed, (int)n_frames, frames_per_slice, H, W, oh, ow, sy, sx,
                                                                     (int)max_in_rows, out_packed, area);
    return check_launch("resize_bilinear_tiled kernel");
  }
  const int word_blocks = (oh * owp + R1_WARPS - 1) / R1_WARPS;
  int slices = (int)((num_sms() * 8 + word_blocks - 1) / word_blocks);
  if (slices > n_frames) slices = (int)n_frames;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  const int frames_per_slice = (int)((n_frames + slices - 1) / slices);
  slices = (int)((n_frames + frames_per_slice - 1) / frames_per_slice);
  dim3 grid(word_blocks, slices);
  resize_bilinear_packed_kernel<<<grid, R1_WARPS * 32, 0, stream>>>(in_packed, (int)n_frames, frames_per_slice, H, W, oh, ow, sy, sx,
                                                                    out_packed, area);
  return check_launch("resize_bilinear_packed kernel");
}

int sola_resize_bilinear_bin_f32(const float* in, long long n_frames, int H, int W, int oh, int ow,
                                 uint32_t* out_packed, float* out_f32, int* area, cudaStream_t stream) {
  SOLA_REQUIRE(in && (out_packed || out_f32), "resize_bilinear_bin_f32: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "resize_bilinear_bin_f32: bad shape");
  if (n_frames == 0) return SOLA_OK;
  if (area) SOLA_CUDA(cudaMemsetAsync(area, 0, sizeof(int) * n_frames, stream));
  const long long warps = n_frames * oh * ((ow + 31) >> 5);
  resize_bilinear_f32_kernel<<<grid_for_warps(warps), 256, 0, stream>>>(in, n_frames, H, W, oh, ow, host_scale(H, oh), host_scale(W, ow),
                                                                        out_packed, out_f32, area);
  return check_launch("resize_bilinear_f32 kernel");
}

int sola_resize_nearest_u8(const uint8_t* in, long long n_frames, int H, int W, int oh, int ow,
                           uint32_t* out_packed, int* area, cudaStream_t stream) {
  SOLA_REQUIRE(in && out_packed, "resize_nearest_u8: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "resize_nearest_u8: bad shape");
  if (n_frames == 0) return SOLA_OK;
  if (area) SOLA_CUDA(cudaMemsetAsync(area, 0, sizeof(int) * n_frames, stream));
  return launch_nearest<false>(in, n_frames, H, W, oh, ow, out_packed, area, stream);
}

int sola_resize_nearest_packed(const uint32_t* in_packed, long long n_frames, int H, int W, int oh, int ow,
                               uint32_t* out_packed, int* area, cudaStream_t stream) {
  SOLA_REQUIRE(in_packed && out_packed, "resize_nearest_packed: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "resize_nearest_packed: bad shape");
  if (n_frames == 0) return SOLA_OK;
  if (area) SOLA_CUDA(cudaMemsetAsync(area, 0, sizeof(int) * n_frames, stream));
  return launch_nearest<true>(in_packed, n_frames, H, W, oh, ow, out_packed, area, stream);
}

}  // extern "C"
