// K1 — read-once binarise + bit-pack + stability popcounts, and the {0,1}-mask packers / unpackers.
//
// Replaces, on the device and in one pass over the logits:
//   (out_mask_logits > 0.0).float()                 generate_tokens_grid.py:215,219,222 ; generate_tokens_gdino.py:232,237,240
//   torch.cat(list, 0)                              generate_tokens_grid.py:224
//   PromptGenerator.get_stability_score             track_generation/prompt_generator.py:169-186
//
// Layout: a mask plane (H, W) becomes (H, Wp) uint32 words, Wp = ceil(W/32); bit b of word w is
// pixel 32*w + b; pad bits of the last word of a row are zero.  When W % 32 == 0 the plane is simply the flat
// bit string of the pixels and the "flat" kernel below is used; otherwise the per-row kernel.
//
// Flat kernel, per warp iteration ("chunk" = 1024 pixels = 32 output words):
//   * every lane issues L = 32/E 128-bit streaming loads (E = elements per 16 B: 4 fp32 / 8 bf16 / 16 u8), lane-
//     contiguous so each load instruction covers 512 contiguous bytes;
//   * each load yields E predicate bits per threshold; L loads fill one private 32-bit register per threshold.
//     For the two stability thresholds only the popcount matters, so those registers are consumed as they are;
//   * for the stored plane the L x L slot matrix held by each group of L lanes is transposed with log2(L)
//     shuffle+select steps, after which every lane owns one finished word, and the warp stores 128 contiguous bytes.
// HBM traffic is exactly the algorithmic minimum: each logit is read once, each packed word written once.
#include "pack_core.cuh"

namespace sola {

// ---- flat kernel ---------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_CHUNK_PX = 1024;
constexpr int K1_CHUNKS_PER_CTA = 32;

template <typename T, int MODE>
__global__ void __launch_bounds__(K1_THREADS)
pack_flat_kernel(const T* __restrict__ in, int frame_px, int ctas_per_frame, Thresholds th,
                 uint32_t* __restrict__ packed, int* __restrict__ cnt_hi, int* __restrict__ cnt_mid, int* __restrict__ cnt_lo) {
  constexpr int E = ElemTraits<T>::E;
  constexpr int L = 32 / E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long frame = blockIdx.x / ctas_per_frame;
  const int cta = blockIdx.x - (int)(frame * ctas_per_frame);
  const int n_chunks = (frame_px + K1_CHUNK_PX - 1) / K1_CHUNK_PX;
  const int c_begin = (int)((long long)cta * n_chunks / ctas_per_frame);
  const int c_end = (int)((long long)(cta + 1) * n_chunks / ctas_per_frame);
  const T* src = in + frame * (long long)frame_px;
  uint32_t* dst = packed ? packed + frame * (long long)(frame_px >> 5) : nullptr;
  const int frame_words = frame_px >> 5;
  const int out_word_in_chunk = E * (lane % L) + lane / L;

  int n_mid = 0, n_hi = 0, n_lo = 0;
  // One chunk.  FULLCHUNK = every vector of the chunk lies inside the frame (all but possibly the last chunk of a
  // frame): no predicates anywhere.  The tail variant loads out-of-frame vectors as "-inf" so no threshold passes.
  auto do_chunk = [&](int c, auto full_tag) {
    constexpr bool FULLCHUNK = decltype(full_tag)::value;
    const int px0 = c * K1_CHUNK_PX;
    uint4 raw[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int px = px0 + E * (j * 32 + lane);
      if (FULLCHUNK || px < frame_px) {
        raw[j] = ld_stream_u4(src + px);
      } else {
        const uint32_t ninf = (MODE == MODE_NONZERO || sizeof(T) == 1) ? 0u : (sizeof(T) == 4 ? 0xff800000u : 0xff80ff80u);
        raw[j] = make_uint4(ninf, ninf, ninf, ninf);
      }
    }
    uint32_t xm = 0, xh = 0, xl = 0;
    if (MODE == MODE_NONZERO || sizeof(T) == 1) {
#pragma unroll
      for (int j = 0; j < L; ++j) {
        uint32_t m, h, l;
        vec_bits<MODE>(raw[j], th, T(), m, h, l);
        xm |= m << (E * j);
      }
    } else if (MODE == MODE_THRESH3) {
      chunk_extract3<T, L, false>(raw, th, xm, xh, xl, n_hi, n_lo);      // hi / lo counted inside (bf16: packed compares)
    } else {
#pragma unroll
      for (int j = L - 1; j >= 0; --j) vec_push<MODE>(raw[j], th, T(), xm, xh, xl);
    }
    n_mid += __popc(xm);
    if (dst) {
      const uint32_t word = transpose_slots<E>(xm, lane);
      const int wi = (px0 >> 5) + out_word_in_chunk;
      if (FULLCHUNK || wi < frame_words) dst[wi] = word;
    }
  };
  for (int c = c_begin + warp; c < c_end; c += K1_WARPS) {
    if ((c + 1) * K1_CHUNK_PX <= frame_px) do_chunk(c, std::true_type());
    else do_chunk(c, std::false_type());
  }

  __shared__ int red[3][K1_WARPS];
  n_mid = warp_sum(n_mid);
  if (MODE == MODE_THRESH3) { n_hi = warp_sum(stab_decode<T>(n_hi)); n_lo = warp_sum(stab_decode<T>(n_lo)); }
  if (lane == 0) { red[0][warp] = n_mid; red[1][warp] = n_hi; red[2][warp] = n_lo; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < K1_WARPS; ++w) s += red[threadIdx.x][w];
    int* out = threadIdx.x == 0 ? cnt_mid : (threadIdx.x == 1 ? cnt_hi : cnt_lo);
    if (out && (threadIdx.x == 0 || MODE == MODE_THRESH3) && s) atomicAdd(out + frame, s);
  }
}

// ---- per-row kernel (any W, any alignment): one warp per image row, ballot per 32 pixels ---------------------
template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((uint32_t) __ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
template <> __device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

constexpr int ROWS_PER_CTA = 8;

template <typename T, int MODE>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
pack_rows_kernel(const T* __restrict__ in, int H, int W, int row_blocks, Thresholds th,
                 uint32_t* __restrict__ packed, int* __restrict__ cnt_hi, int* __restrict__ cnt_mid, int* __restrict__ cnt_lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long frame = blockIdx.x / row_blocks;
  const int row = (blockIdx.x - (int)(frame * row_blocks)) * ROWS_PER_CTA + warp;
  const int Wp = (W + 31) >> 5;
  int n_mid = 0, n_hi = 0, n_lo = 0;
  if (row < H) {
    const T* src = in + (frame * H + row) * (long long)W;
    uint32_t* dst = packed ? packed + (frame * H + row) * (long long)Wp : nullptr;
    for (int w0 = 0; w0 < Wp; w0 += 4) {
      float v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int x = (w0 + u) * 32 + lane;
        ok[u] = x < W;
        v[u] = ok[u] ? load_as_float(src + x) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (w0 + u >= Wp) break;                       // warp-uniform
        const bool pm = ok[u] && (MODE == MODE_NONZERO ? (v[u] != 0.f) : (v[u] > th.mid));
        const uint32_t bm = __ballot_sync(FULL, pm);
        n_mid += __popc(bm);
        if (MODE == MODE_THRESH3) {
          n_hi += __popc(__ballot_sync(FULL, ok[u] && v[u] > th.hi));
          n_lo += __popc(__ballot_sync(FULL, ok[u] && v[u] > th.lo));
        }
        if (dst && lane == u) dst[w0 + u] = bm;
      }
    }
  }
  // counts are warp-uniform here (every lane saw the same ballots)
  __shared__ int red[3][ROWS_PER_CTA];
  if (lane == 0) { red[0][warp] = n_mid; red[1][warp] = n_hi; red[2][warp] = n_lo; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < ROWS_PER_CTA; ++w) s += red[threadIdx.x][w];
    int* out = threadIdx.x == 0 ? cnt_mid : (threadIdx.x == 1 ? cnt_hi : cnt_lo);
    if (out && (threadIdx.x == 0 || MODE == MODE_THRESH3) && s) atomicAdd(out + frame, s);
  }
}

// only float / bf16 logits have the any-width band kernel
template <typename T>
static int band_pack_dispatch(const T* in, long long n, int H, int W, Thresholds th, uint32_t* p, int* a, int* b, int* c, cudaStream_t s) {
  return launch_band_pack<T>(in, n, H, W, th, p, a, b, c, s);
}
static int band_pack_dispatch(const uint8_t*, long long, int, int, Thresholds, uint32_t*, int*, int*, int*, cudaStream_t) { return SOLA_ERR_UNSUPPORTED; }

template <typename T, int MODE>
static int launch_pack(const T* in, long long n_frames, int H, int W, Thresholds th, uint32_t* packed,
                       int* cnt_hi, int* cnt_mid, int* cnt_lo, cudaStream_t stream) {
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0, "pack: bad shape n_frames=%lld H=%d W=%d", n_frames, H, W);
  SOLA_REQUIRE((long long)H * W < (1ll << 31), "pack: frame larger than 2^31 pixels");
  if (n_frames == 0) return SOLA_OK;
  SOLA_REQUIRE(in != nullptr, "pack: input pointer is null");
  if (cnt_mid) SOLA_CUDA(cudaMemsetAsync(cnt_mid, 0, sizeof(int) * n_frames, stream));
  if (MODE == MODE_THRESH3) {
    if (cnt_hi) SOLA_CUDA(cudaMemsetAsync(cnt_hi, 0, sizeof(int) * n_frames, stream));
    if (cnt_lo) SOLA_CUDA(cudaMemsetAsync(cnt_lo, 0, sizeof(int) * n_frames, stream));
  }
  const int frame_px = H * W;
  const bool flat = (W % 32 == 0) && aligned16(in);
  if (!flat && MODE == MODE_THRESH3 && sizeof(T) > 1 && aligned16(in) && (n_frames * frame_px) % ElemTraits<T>::E == 0 && n_frames < (1ll << 31)) {
    // W % 32 != 0 (e.g. 480 x 854): flat 16-byte-aligned run per band + re-cut into row-padded words (fused_pack_resize.cu)
    const int rc = band_pack_dispatch(in, n_frames, H, W, th, packed, cnt_hi, cnt_mid, cnt_lo, stream);
    if (rc != SOLA_ERR_UNSUPPORTED) return rc;
  }
  if (flat) {
    const int n_chunks = (frame_px + K1_CHUNK_PX - 1) / K1_CHUNK_PX;
    const int ctas_per_frame = (n_chunks + K1_CHUNKS_PER_CTA - 1) / K1_CHUNKS_PER_CTA;
    const long long grid = n_frames * ctas_per_frame;
    SOLA_REQUIRE(grid < (1ll << 31), "pack: grid too large (%lld CTAs); split the batch", grid);
    pack_flat_kernel<T, MODE><<<(unsigned)grid, K1_THREADS, 0, stream>>>(in, frame_px, ctas_per_frame, th, packed, cnt_hi, cnt_mid, cnt_lo);
  } else {
    const int row_blocks = (H + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    const long long grid = n_frames * row_blocks;
    SOLA_REQUIRE(grid < (1ll << 31), "pack: grid too large (%lld CTAs); split the batch", grid);
    pack_rows_kernel<T, MODE><<<(unsigned)grid, ROWS_PER_CTA * 32, 0, stream>>>(in, H, W, row_blocks, th, packed, cnt_hi, cnt_mid, cnt_lo);
  }
  return check_launch("pack kernel");
}

// ---- unpack: packed (n, H, Wp) -> fp32 / u8 {0,1} planes (drop-in return types) ------------------------------
// warp-cooperative variant: each warp takes one word at a time, lane b writes pixel b (coalesced 128 B fp32 stores)
template <typename T>
__global__ void __launch_bounds__(256)
unpack_warp_kernel(const uint32_t* __restrict__ packed, long long n_rows, int W, int Wp, T one, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long total_words = n_rows * Wp;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp0 * 32; base < total_words; base += n_warps * 32) {
    const long long mine = base + lane;
    const uint32_t my_bits = mine < total_words ? packed[mine] : 0u;
    const int n_here = (int)min(32ll, total_words - base);
    for (int k = 0; k < n_here; ++k) {
      const uint32_t bits = __shfl_sync(FULL, my_bits, k);
      const long long wi = base + k;
      const long long row = wi / Wp;
      const int w = (int)(wi - row * Wp);
      const int x = w * 32 + lane;
      if (x < W) out[row * W + x] = ((bits >> lane) & 1u) ? one : (T)0;
    }
  }
}

template <typename T>
static int launch_unpack(const uint32_t* packed, long long n_frames, int H, int W, T one, T* out, cudaStream_t stream) {
  SOLA_REQUIRE(packed && out, "unpack: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0, "unpack: bad shape");
  if (n_frames == 0) return SOLA_OK;
  const int Wp = (W + 31) >> 5;
  const long long n_rows = n_frames * H;
  const long long warps_needed = (n_rows * Wp + 31) / 32;
  long long blocks = (warps_needed + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  unpack_warp_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(packed, n_rows, W, Wp, one, out);
  return check_launch("unpack kernel");
}

}  // namespace sola

using namespace sola;

extern "C" {

int sola_binarize_pack_f32(const float* logits, long long n_frames, int H, int W, double thr, double off,
                           uint32_t* packed_out, int* cnt_hi, int* cnt_mid, int* cnt_lo, cudaStream_t stream) {
  return launch_pack<float, MODE_THRESH3>(logits, n_frames, H, W, make_thresholds(thr, off), packed_out, cnt_hi, cnt_mid, cnt_lo, stream);
}

int sola_binarize_pack_bf16(const void* logits, long long n_frames, int H, int W, double thr, double off,
                            uint32_t* packed_out, int* cnt_hi, int* cnt_mid, int* cnt_lo, cudaStream_t stream) {
  return launch_pack<__nv_bfloat16, MODE_THRESH3>(reinterpret_cast<const __nv_bfloat16*>(logits), n_frames, H, W,
                                                  make_thresholds(thr, off), packed_out, cnt_hi, cnt_mid, cnt_lo, stream);
}

int sola_threshold_pack_f32(const float* x, long long n_frames, int H, int W, double thr,
                            uint32_t* packed_out, int* area, cudaStream_t stream) {
  return launch_pack<float, MODE_THRESH1>(x, n_frames, H, W, make_thresholds(thr, 0.0), packed_out, nullptr, area, nullptr, stream);
}

int sola_pack_mask_f32(const float* mask, long long n_frames, int H, int W, uint32_t* packed_out, int* area, cudaStream_t stream) {
  return launch_pack<float, MODE_NONZERO>(mask, n_frames, H, W, Thresholds{0, 0, 0}, packed_out, nullptr, area, nullptr, stream);
}

int sola_pack_mask_u8(const uint8_t* mask, long long n_frames, int H, int W, uint32_t* packed_out, int* area, cudaStream_t stream) {
  return launch_pack<uint8_t, MODE_NONZERO>(mask, n_frames, H, W, Thresholds{0, 0, 0}, packed_out, nullptr, area, nullptr, stream);
}

int sola_unpack_f32(const uint32_t* packed, long long n_frames, int H, int W, float* out, cudaStream_t stream) {
  return launch_unpack<float>(packed, n_frames, H, W, 1.0f, out, stream);
}

int sola_unpack_u8(const uint32_t* packed, long long n_frames, int H, int W, uint8_t* out, cudaStream_t stream) {
  return launch_unpack<uint8_t>(packed, n_frames, H, W, (uint8_t)1, out, stream);
}

// foreground written as `one_value` (255 for the PNG planes of inference.py:90)
int sola_unpack_u8_value(const uint32_t* packed, long long n_frames, int H, int W, int one_value, uint8_t* out, cudaStream_t stream) {
  SOLA_REQUIRE(one_value >= 0 && one_value <= 255, "unpack_u8_value: one_value out of range");
  return launch_unpack<uint8_t>(packed, n_frames, H, W, (uint8_t)one_value, out, stream);
}

}  // extern "C"
