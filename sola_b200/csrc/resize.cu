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

// ---- R1 from packed planes -----------------------------------------------------------------------------------
// One warp per output word (oy, owx); lane = output pixel.  The per-pixel geometry is frame-invariant, so each
// warp keeps it in registers and walks a slice of the frames.  Words whose 2-row source window is uniformly 0 or
// uniformly 1 are resolved with one vote (background / interior), only edge words evaluate the interpolation.
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
  // source word window of this output word: [wlo, whi] covers x0 of lane 0 .. x1 of the last live lane
  const int x_first = __shfl_sync(FULL, ax.i0, 0);
  const int x_last = __shfl_sync(FULL, ax.i1, 31);            // dead lanes replicate ow-1, so lane 31 is the max
  const int wlo = x_first >> 5, whi = x_last >> 5;
  const int nwin = whi - wlo + 1;                             // <= 32 for any downscale factor below ~31
  const long long FW = (long long)H * Wp, oFW = (long long)oh * owp;
  const bool window_in_one_vote = nwin <= 16;
  for (int f = f_begin; f < f_end; ++f) {
    const uint32_t* r0 = in + f * FW + (long long)ay.i0 * Wp;
    const uint32_t* r1 = in + f * FW + (long long)ay.i1 * Wp;
    uint32_t word;
    bool resolved = false;
    if (window_in_one_vote) {
      // lanes 0..nwin-1 fetch row y0's window, lanes 16..16+nwin-1 row y1's
      const int k = lane & 15;
      uint32_t wv = 0;
      const bool fetch = k < nwin;
      if (fetch) wv = __ldg(((lane < 16) ? r0 : r1) + wlo + k);
      // pixels outside [x_first, x_last] in the edge words do not matter; mask them to the neighbouring value
      uint32_t care = 0xffffffffu;
      if (fetch) {
        if (k == 0) care &= 0xffffffffu << (x_first & 31);
        if (k == nwin - 1) care &= 0xffffffffu >> (31 - (x_last & 31));
      }
      const bool any1 = fetch && (wv & care) != 0u;
      const bool any0 = fetch && ((~wv) & care) != 0u;
      const bool has1 = __any_sync(FULL, any1), has0 = __any_sync(FULL, any0);
      if (!has1) { word = 0u; resolved = true; }
      else if (!has0) { word = live ? 0xffffffffu : 0u; resolved = true; }
    }
    if (resolved) {
      // all-ones: every live pixel interpolates 1-valued neighbours -> val = fma(h0, fl(w0+w1), h1*fl(w0+w1)) ~ 1 > 0.5
      word = __ballot_sync(FULL, live && word != 0u);
    } else {
      const float v00 = (float)get_bit(r0, ax.i0), v01 = (float)get_bit(r0, ax.i1);
      const float v10 = (float)get_bit(r1, ax.i0), v11 = (float)get_bit(r1, ax.i1);
      const float val = bilinear_val(ax, ay, v00, v01, v10, v11);
      word = __ballot_sync(FULL, live && val > 0.5f);
    }
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

template <bool PACKED_IN>
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const void* __restrict__ in_, long long n_frames, int H, int W, int oh, int ow, float sy, float sx,
                      uint32_t* __restrict__ out_packed, int* __restrict__ area) {
  const int lane = threadIdx.x & 31;
  const int owp = (ow + 31) >> 5, Wp = (W + 31) >> 5;
  const long long words_per_frame = (long long)oh * owp;
  const long long total = n_frames * words_per_frame;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long wi = warp0; wi < total; wi += n_warps) {
    const long long f = wi / words_per_frame;
    const int rem = (int)(wi - f * words_per_frame);
    const int oy = rem / owp, owx = rem - oy * owp;
    const int ox = owx * 32 + lane;
    bool bit = false;
    if (ox < ow) {
      const int y = nearest_src(oy, sy, H), x = nearest_src(ox, sx, W);
      if (PACKED_IN) bit = get_bit(reinterpret_cast<const uint32_t*>(in_) + (f * H + y) * (long long)Wp, x) != 0u;
      else bit = __ldg(reinterpret_cast<const uint8_t*>(in_) + (f * H + y) * (long long)W + x) != 0;
    }
    const uint32_t word = __ballot_sync(FULL, bit);
    if (lane == 0) {
      out_packed[wi] = word;
      if (area && word) atomicAdd(area + f, __popc(word));
    }
  }
}

static inline float host_scale(int in_size, int out_size) { return (float)in_size / (float)out_size; }

static int grid_for_warps(long long warps) {
  long long blocks = (warps + 7) / 8;
  const long long cap = (long long)num_sms() * 32;
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
  const int owp = (ow + 31) >> 5;
  const int word_blocks = (oh * owp + R1_WARPS - 1) / R1_WARPS;
  // enough CTAs for ~8 per SM: split the frame axis when the plane alone is too small
  int slices = (int)((num_sms() * 8 + word_blocks - 1) / word_blocks);
  if (slices > n_frames) slices = (int)n_frames;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  const int frames_per_slice = (int)((n_frames + slices - 1) / slices);
  slices = (int)((n_frames + frames_per_slice - 1) / frames_per_slice);
  dim3 grid(word_blocks, slices);
  resize_bilinear_packed_kernel<<<grid, R1_WARPS * 32, 0, stream>>>(in_packed, (int)n_frames, frames_per_slice, H, W, oh, ow,
                                                                    host_scale(H, oh), host_scale(W, ow), out_packed, area);
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
  const long long warps = n_frames * oh * ((ow + 31) >> 5);
  resize_nearest_kernel<false><<<grid_for_warps(warps), 256, 0, stream>>>(in, n_frames, H, W, oh, ow, host_scale(H, oh), host_scale(W, ow),
                                                                          out_packed, area);
  return check_launch("resize_nearest_u8 kernel");
}

int sola_resize_nearest_packed(const uint32_t* in_packed, long long n_frames, int H, int W, int oh, int ow,
                               uint32_t* out_packed, int* area, cudaStream_t stream) {
  SOLA_REQUIRE(in_packed && out_packed, "resize_nearest_packed: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "resize_nearest_packed: bad shape");
  if (n_frames == 0) return SOLA_OK;
  if (area) SOLA_CUDA(cudaMemsetAsync(area, 0, sizeof(int) * n_frames, stream));
  const long long warps = n_frames * oh * ((ow + 31) >> 5);
  resize_nearest_kernel<true><<<grid_for_warps(warps), 256, 0, stream>>>(in_packed, n_frames, H, W, oh, ow, host_scale(H, oh), host_scale(W, ow),
                                                                         out_packed, area);
  return check_launch("resize_nearest_packed kernel");
}

}  // extern "C"
