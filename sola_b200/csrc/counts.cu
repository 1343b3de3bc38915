// K3 — per-frame intersection / area counts: the integer core of J, F, IoU and the label metrics.
//
// Replaces the ATen mul/add/sum + .item() chains of
//   Evaluator.compute_J / compute_F                  evaluator.py:227-247
//   seg_utils.compute_mask_iou / compute_masklet_iou track_generation/seg_utils.py:110-142
//   utils.compute_mask_iou_torch / compute_mask_metrics  track_generation/utils.py:65-75,132-174
// Every caller needs, per frame, only three integers: |A∩B|, |A|, |B| (union = |A|+|B|-|A∩B|; F's
// tp/fp/fn = Σ inter, Σ|pred| - tp, Σ|gt| - tp).  Counts are exact int32 per frame; the host sums in int64/float64.
//
// Two input forms:
//   * raw planes (fp32 / u8, "nonzero = foreground"): the drop-in signatures hand us fp32 {0,1} tensors; this
//     kernel reads each plane exactly once (HBM bound, 2*sizeof(T) bytes per pixel pair);
//   * bit-packed planes: batched (Na tracks x Nb objects x T frames) and ragged (a J&F sweep over units of
//     different shape, concatenated) variants.
#include "common.cuh"
#include "csa.cuh"

namespace sola {

// ---- raw planes ----------------------------------------------------------------------------------------------
template <typename T> struct RawTraits;
template <> struct RawTraits<float> { static constexpr int E = 4; };
template <> struct RawTraits<uint8_t> { static constexpr int E = 16; };

__device__ __forceinline__ uint32_t nz_bits(const uint4& r, float) {
  return (__uint_as_float(r.x) != 0.f ? 1u : 0u) | (__uint_as_float(r.y) != 0.f ? 2u : 0u) |
         (__uint_as_float(r.z) != 0.f ? 4u : 0u) | (__uint_as_float(r.w) != 0.f ? 8u : 0u);
}
__device__ __forceinline__ uint32_t nz4(uint32_t w) { return ((__vcmpne4(w, 0u) & 0x01010101u) * 0x01020408u) >> 24; }
__device__ __forceinline__ uint32_t nz_bits(const uint4& r, uint8_t) {
  return nz4(r.x) | (nz4(r.y) << 4) | (nz4(r.z) << 8) | (nz4(r.w) << 12);
}

constexpr int K3_THREADS = 256;
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int K3_CHUNK_PX = 1024;          // pixels per warp iteration
constexpr int K3_CHUNKS_PER_CTA = 32;

__device__ __forceinline__ void block_reduce3_atomic(int a, int b, int c, int* oa, int* ob, int* oc, long long idx) {
  __shared__ int red[3][K3_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < K3_WARPS; ++w) s += red[threadIdx.x][w];
    int* out = threadIdx.x == 0 ? oa : (threadIdx.x == 1 ? ob : oc);
    if (out && s) atomicAdd(out + idx, s);
  }
}

// Vector path: frame_px % E == 0 and both bases 16-byte aligned.
template <typename T>
__global__ void __launch_bounds__(K3_THREADS)
raw_counts_vec_kernel(const T* __restrict__ a, const T* __restrict__ b, int frame_px, int ctas_per_frame,
                      int* __restrict__ inter, int* __restrict__ area_a, int* __restrict__ area_b) {
  constexpr int E = RawTraits<T>::E;
  constexpr int L = 32 / E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long frame = blockIdx.x / ctas_per_frame;
  const int cta = blockIdx.x - (int)(frame * ctas_per_frame);
  const int n_chunks = (frame_px + K3_CHUNK_PX - 1) / K3_CHUNK_PX;
  const int c_begin = (int)((long long)cta * n_chunks / ctas_per_frame);
  const int c_end = (int)((long long)(cta + 1) * n_chunks / ctas_per_frame);
  const T* pa = a + frame * (long long)frame_px;
  const T* pb = b + frame * (long long)frame_px;
  int n_i = 0, n_a = 0, n_b = 0;
  for (int c = c_begin + warp; c < c_end; c += K3_WARPS) {
    const int px0 = c * K3_CHUNK_PX;
    uint4 ra[L], rb[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int px = px0 + E * (j * 32 + lane);
      const bool ok = px < frame_px;
      ra[j] = ok ? ld_stream_u4(pa + px) : make_uint4(0, 0, 0, 0);
      rb[j] = ok ? ld_stream_u4(pb + px) : make_uint4(0, 0, 0, 0);
    }
    uint32_t xa = 0, xb = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      xa |= nz_bits(ra[j], T()) << (E * j);
      xb |= nz_bits(rb[j], T()) << (E * j);
    }
    n_a += __popc(xa);
    n_b += __popc(xb);
    n_i += __popc(xa & xb);
  }
  block_reduce3_atomic(n_i, n_a, n_b, inter, area_a, area_b, frame);
}

// Scalar path: any frame_px / alignment.
template <typename T>
__global__ void __launch_bounds__(K3_THREADS)
raw_counts_scalar_kernel(const T* __restrict__ a, const T* __restrict__ b, int frame_px, int ctas_per_frame,
                         int* __restrict__ inter, int* __restrict__ area_a, int* __restrict__ area_b) {
  const long long frame = blockIdx.x / ctas_per_frame;
  const int cta = blockIdx.x - (int)(frame * ctas_per_frame);
  const T* pa = a + frame * (long long)frame_px;
  const T* pb = b + frame * (long long)frame_px;
  int n_i = 0, n_a = 0, n_b = 0;
  for (int px = cta * K3_THREADS + threadIdx.x; px < frame_px; px += ctas_per_frame * K3_THREADS) {
    const bool fa = pa[px] != (T)0, fb = pb[px] != (T)0;
    n_a += fa; n_b += fb; n_i += (fa && fb);
  }
  block_reduce3_atomic(n_i, n_a, n_b, inter, area_a, area_b, frame);
}

template <typename T>
static int launch_raw_counts(const T* a, const T* b, long long n_frames, long long frame_px, int* inter, int* area_a, int* area_b,
                             cudaStream_t stream) {
  SOLA_REQUIRE(a && b && inter && area_a && area_b, "frame_counts: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && frame_px > 0 && frame_px < (1ll << 31), "frame_counts: bad shape n_frames=%lld frame_px=%lld", n_frames, frame_px);
  if (n_frames == 0) return SOLA_OK;
  if (area_a == inter + n_frames && area_b == area_a + n_frames) {      // the usual (3, n) tensor: one memset instead of three
    SOLA_CUDA(cudaMemsetAsync(inter, 0, sizeof(int) * 3 * n_frames, stream));
  } else {
    SOLA_CUDA(cudaMemsetAsync(inter, 0, sizeof(int) * n_frames, stream));
    SOLA_CUDA(cudaMemsetAsync(area_a, 0, sizeof(int) * n_frames, stream));
    SOLA_CUDA(cudaMemsetAsync(area_b, 0, sizeof(int) * n_frames, stream));
  }
  constexpr int E = RawTraits<T>::E;
  const int n_chunks = (int)((frame_px + K3_CHUNK_PX - 1) / K3_CHUNK_PX);
  const int ctas_per_frame = (n_chunks + K3_CHUNKS_PER_CTA - 1) / K3_CHUNKS_PER_CTA;
  const long long grid = n_frames * ctas_per_frame;
  SOLA_REQUIRE(grid < (1ll << 31), "frame_counts: grid too large; split the batch");
  if (frame_px % E == 0 && aligned16(a) && aligned16(b))
    raw_counts_vec_kernel<T><<<(unsigned)grid, K3_THREADS, 0, stream>>>(a, b, (int)frame_px, ctas_per_frame, inter, area_a, area_b);
  else
    raw_counts_scalar_kernel<T><<<(unsigned)grid, K3_THREADS, 0, stream>>>(a, b, (int)frame_px, ctas_per_frame, inter, area_a, area_b);
  return check_launch("frame_counts kernel");
}

// ---- packed planes, batched: inter[Na][Nb][T], area_a[Na][T], area_b[Nb][T] ----------------------------------
constexpr int NB_TILE = 4;     // objects (B planes) per CTA
constexpr int NA_TILE = 2;     // tracks (A planes) per CTA: every B word fetched from L2 serves NA_TILE tracks (4 tracks need 96+
                               // registers and measured slower, 89-112 vs 83 us: profiles/r3_build_constants.json)
#ifndef LABEL_MIN_CTAS
#define LABEL_MIN_CTAS 4
#endif

// CTA = one frame x NA_TILE tracks x up to NB_TILE objects.  History of the bound: plain POPCs made it xu-pipe bound (16 lanes/clk/SM;
// a predicated-off POPC still occupies the pipe, so the live-slot choices are COMPILED in as template flags), carry-save counters
// (csa.cuh) moved that work to the alu pipe, and what remained was L2 traffic — with one track per CTA every track word pulled
// NK object words out of L2 (4 words moved per word counted), hence the track tile.  |b| is stored once per (object, frame) — by the
// CTAs of the first track tile, with plain POPCs —, |a| once per (track, frame) — by the first object tile.
template <int VEC, int NK, bool WANT_A, bool WANT_B>
__device__ __forceinline__ void packed_count_loop(const uint32_t* const (&pa)[NA_TILE], const uint32_t* const (&pb)[NB_TILE], int FW,
                                                  int (&acca)[NA_TILE], int (&acc)[NA_TILE][NB_TILE], int (&accb)[NB_TILE]) {
  if (VEC == 4) {                                  // planes 16-byte aligned, FW % 4 == 0: 128-bit loads, NA_TILE + NK in flight per iteration
    const int nq = FW >> 2;
    Csa ca[NA_TILE], ci[NA_TILE][NK];
#pragma unroll
    for (int i = 0; i < NA_TILE; ++i) {
      ca[i] = Csa{0u, 0u, 0};
#pragma unroll
      for (int k = 0; k < NK; ++k) ci[i][k] = Csa{0u, 0u, 0};
    }
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
      uint4 x[NA_TILE], y[NK];
#pragma unroll
      for (int i = 0; i < NA_TILE; ++i) x[i] = __ldg(reinterpret_cast<const uint4*>(pa[i]) + q);
#pragma unroll
      for (int k = 0; k < NK; ++k) y[k] = __ldg(reinterpret_cast<const uint4*>(pb[k]) + q);
#pragma unroll
      for (int i = 0; i < NA_TILE; ++i) {
        if (WANT_A) csa_add4(ca[i], x[i].x, x[i].y, x[i].z, x[i].w);
#pragma unroll
        for (int k = 0; k < NK; ++k) csa_quad(ci[i][k], x[i], y[k]);
      }
      if (WANT_B) {
#pragma unroll
        for (int k = 0; k < NK; ++k) accb[k] += __popc(y[k].x) + __popc(y[k].y) + __popc(y[k].z) + __popc(y[k].w);
      }
    }
#pragma unroll
    for (int i = 0; i < NA_TILE; ++i) {
      if (WANT_A) acca[i] += csa_total(ca[i]);
#pragma unroll
      for (int k = 0; k < NK; ++k) acc[i][k] += csa_total(ci[i][k]);
    }
  } else {
    for (int w = threadIdx.x; w < FW; w += blockDim.x) {
      uint32_t y[NK];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        y[k] = pb[k][w];
        if (WANT_B) accb[k] += __popc(y[k]);
      }
#pragma unroll
      for (int i = 0; i < NA_TILE; ++i) {
        const uint32_t x = pa[i][w];
        if (WANT_A) acca[i] += __popc(x);
#pragma unroll
        for (int k = 0; k < NK; ++k) acc[i][k] += __popc(x & y[k]);
      }
    }
  }
}

template <int VEC, int NK>
__device__ __forceinline__ void packed_count_dispatch(bool want_a, bool want_b, const uint32_t* const (&pa)[NA_TILE],
                                                      const uint32_t* const (&pb)[NB_TILE], int FW, int (&acca)[NA_TILE],
                                                      int (&acc)[NA_TILE][NB_TILE], int (&accb)[NB_TILE]) {
  if (want_a) {
    if (want_b) packed_count_loop<VEC, NK, true, true>(pa, pb, FW, acca, acc, accb);
    else packed_count_loop<VEC, NK, true, false>(pa, pb, FW, acca, acc, accb);
  } else {
    if (want_b) packed_count_loop<VEC, NK, false, true>(pa, pb, FW, acca, acc, accb);
    else packed_count_loop<VEC, NK, false, false>(pa, pb, FW, acca, acc, accb);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256, LABEL_MIN_CTAS)
packed_counts_kernel(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, int Na, int Nb, int T, int FW,
                     int* __restrict__ inter, int* __restrict__ area_a, int* __restrict__ area_b) {
  const int t = blockIdx.x % T, ia0 = (blockIdx.x / T) * NA_TILE;
  const int jb0 = blockIdx.y * NB_TILE;
  const int nk = min(NB_TILE, Nb - jb0);                 // live object slots of this tile (CTA-uniform)
  const bool want_b = ia0 == 0, want_a = blockIdx.y == 0; // CTA-uniform
  const uint32_t* pa[NA_TILE];
  const uint32_t* pb[NB_TILE];
#pragma unroll
  for (int i = 0; i < NA_TILE; ++i) pa[i] = A + ((long long)min(ia0 + i, Na - 1) * T + t) * FW;     // past Na: the last track again, not stored
#pragma unroll
  for (int k = 0; k < NB_TILE; ++k) pb[k] = B + ((long long)min(jb0 + k, Nb - 1) * T + t) * FW;
  int acc[NA_TILE][NB_TILE], accb[NB_TILE] = {0, 0, 0, 0}, acca[NA_TILE];
#pragma unroll
  for (int i = 0; i < NA_TILE; ++i) {
    acca[i] = 0;
#pragma unroll
    for (int k = 0; k < NB_TILE; ++k) acc[i][k] = 0;
  }
  switch (nk) {
    case 1: packed_count_dispatch<VEC, 1>(want_a, want_b, pa, pb, FW, acca, acc, accb); break;
    case 2: packed_count_dispatch<VEC, 2>(want_a, want_b, pa, pb, FW, acca, acc, accb); break;
    case 3: packed_count_dispatch<VEC, 3>(want_a, want_b, pa, pb, FW, acca, acc, accb); break;
    default: packed_count_dispatch<VEC, 4>(want_a, want_b, pa, pb, FW, acca, acc, accb); break;
  }
  constexpr int NRED = NA_TILE * NB_TILE + NB_TILE + NA_TILE;      // [i][k] intersections, |b_k|, |a_i|
  __shared__ int red[NRED][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NA_TILE; ++i) {
    acca[i] = warp_sum(acca[i]);
#pragma unroll
    for (int k = 0; k < NB_TILE; ++k) acc[i][k] = warp_sum(acc[i][k]);
  }
#pragma unroll
  for (int k = 0; k < NB_TILE; ++k) accb[k] = warp_sum(accb[k]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NA_TILE; ++i) {
      red[NA_TILE * NB_TILE + NB_TILE + i][warp] = acca[i];
#pragma unroll
      for (int k = 0; k < NB_TILE; ++k) red[i * NB_TILE + k][warp] = acc[i][k];
    }
#pragma unroll
    for (int k = 0; k < NB_TILE; ++k) red[NA_TILE * NB_TILE + k][warp] = accb[k];
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    int s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    const int e = threadIdx.x;
    if (e < NA_TILE * NB_TILE) {
      const int i = e / NB_TILE, k = e % NB_TILE;
      if (ia0 + i < Na && jb0 + k < Nb) inter[((long long)(ia0 + i) * Nb + jb0 + k) * T + t] = s;
    } else if (e < NA_TILE * NB_TILE + NB_TILE) {
      const int k = e - NA_TILE * NB_TILE;
      if (want_b && jb0 + k < Nb) area_b[(long long)(jb0 + k) * T + t] = s;
    } else {
      const int i = e - NA_TILE * NB_TILE - NB_TILE;
      if (want_a && ia0 + i < Na) area_a[(long long)(ia0 + i) * T + t] = s;
    }
  }
}

// ---- packed planes, ragged: frame f occupies words [off[f], off[f+1]) of both buffers ------------------------
__global__ void __launch_bounds__(256)
packed_counts_ragged_kernel(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, const long long* __restrict__ off,
                            int n_frames, int* __restrict__ inter, int* __restrict__ area_a, int* __restrict__ area_b) {
  __shared__ int red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
    const long long w0 = off[f], w1 = off[f + 1];
    int n_i = 0, n_a = 0, n_b = 0;
    for (long long w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
      const uint32_t x = A[w], y = B[w];
      n_a += __popc(x); n_b += __popc(y); n_i += __popc(x & y);
    }
    n_i = warp_sum(n_i); n_a = warp_sum(n_a); n_b = warp_sum(n_b);
    if (lane == 0) { red[0][warp] = n_i; red[1][warp] = n_a; red[2][warp] = n_b; }
    __syncthreads();
    if (threadIdx.x < 3) {
      int s = 0;
      for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
      (threadIdx.x == 0 ? inter : (threadIdx.x == 1 ? area_a : area_b))[f] = s;
    }
    __syncthreads();
  }
}

// ---- OR-merge of selected packed tracks (dataloader.py:319-350 get_sam2_masklet, :285-299 get_gt_masklet) -----
__global__ void __launch_bounds__(256)
or_merge_kernel(const uint32_t* __restrict__ tracks, const uint8_t* __restrict__ select, int K, long long words,
                uint32_t* __restrict__ out) {
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
    uint32_t acc = 0;
    for (int k = 0; k < K; ++k)
      if (select == nullptr || select[k]) acc |= tracks[(long long)k * words + w];
    out[w] = acc;
  }
}

// J&F accumulators from the three per-frame counts: uni[t] holds |pred_t| on entry and |pred_t ∪ gt_t| on exit;
// totals[0..2] += (tp, fp, fn) over all frames — the volume sums of Evaluator.compute_F (evaluator.py:239-247) as exact integers.
__global__ void __launch_bounds__(256)
jf_finalize_kernel(const int* __restrict__ inter, int* __restrict__ uni, const int* __restrict__ area_gt, long long T,
                   unsigned long long* __restrict__ totals) {
  long long tp = 0, fp = 0, fn = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
    const int i = inter[t], a = uni[t], b = area_gt[t];
    uni[t] = a + b - i;
    tp += i; fp += a - i; fn += b - i;
  }
  __shared__ long long red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    tp += __shfl_xor_sync(FULL, tp, d); fp += __shfl_xor_sync(FULL, fp, d); fn += __shfl_xor_sync(FULL, fn, d);
  }
  if (lane == 0) { red[0][warp] = tp; red[1][warp] = fp; red[2][warp] = fn; }
  __syncthreads();
  if (threadIdx.x < 3) {
    long long s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s) atomicAdd(totals + threadIdx.x, (unsigned long long)s);
  }
}

}  // namespace sola

using namespace sola;

extern "C" {

int sola_frame_counts_f32(const float* a, const float* b, long long n_frames, long long frame_px,
                          int* inter, int* area_a, int* area_b, cudaStream_t stream) {
  return launch_raw_counts<float>(a, b, n_frames, frame_px, inter, area_a, area_b, stream);
}

int sola_frame_counts_u8(const uint8_t* a, const uint8_t* b, long long n_frames, long long frame_px,
                         int* inter, int* area_a, int* area_b, cudaStream_t stream) {
  return launch_raw_counts<uint8_t>(a, b, n_frames, frame_px, inter, area_a, area_b, stream);
}

int sola_frame_counts_packed(const uint32_t* a, const uint32_t* b, int Na, int Nb, int T, long long frame_words,
                             int* inter, int* area_a, int* area_b, cudaStream_t stream) {
  SOLA_REQUIRE(a && b && inter && area_a && area_b, "frame_counts_packed: null pointer");
  SOLA_REQUIRE(Na >= 0 && Nb >= 0 && T >= 0 && frame_words > 0 && frame_words < (1ll << 26),
               "frame_counts_packed: bad shape Na=%d Nb=%d T=%d frame_words=%lld", Na, Nb, T, frame_words);
  if (Na == 0 || Nb == 0 || T == 0) return SOLA_OK;
  SOLA_REQUIRE((long long)Na * T < (1ll << 31) && (Nb + NB_TILE - 1) / NB_TILE <= 65535, "frame_counts_packed: grid too large");
  dim3 grid((unsigned)((long long)((Na + NA_TILE - 1) / NA_TILE) * T), (unsigned)((Nb + NB_TILE - 1) / NB_TILE));
  if (frame_words % 4 == 0 && aligned16(a) && aligned16(b))
    packed_counts_kernel<4><<<grid, 256, 0, stream>>>(a, b, Na, Nb, T, (int)frame_words, inter, area_a, area_b);
  else
    packed_counts_kernel<1><<<grid, 256, 0, stream>>>(a, b, Na, Nb, T, (int)frame_words, inter, area_a, area_b);
  return check_launch("packed_counts kernel");
}

int sola_frame_counts_packed_ragged(const uint32_t* a, const uint32_t* b, const long long* word_offsets, int n_frames,
                                    int* inter, int* area_a, int* area_b, cudaStream_t stream) {
  SOLA_REQUIRE(a && b && word_offsets && inter && area_a && area_b, "frame_counts_packed_ragged: null pointer");
  SOLA_REQUIRE(n_frames >= 0, "frame_counts_packed_ragged: negative frame count");
  if (n_frames == 0) return SOLA_OK;
  const int grid = min(n_frames, num_sms() * 8);
  packed_counts_ragged_kernel<<<grid, 256, 0, stream>>>(a, b, word_offsets, n_frames, inter, area_a, area_b);
  return check_launch("packed_counts_ragged kernel");
}

// ---- sola_jf_*: the J&F accumulators of one (video, expression) unit in the shape SURVEY.md §8(b) names -----------------------
// inter[t] = |pred_t ∩ gt_t|, uni[t] = |pred_t ∪ gt_t| (J_t = 1 if uni == 0 else inter / uni, evaluator.py:227-237) and
// tp_fp_fn[0..2] = volume sums for the reference's Dice-style F (evaluator.py:239-247).  One count pass + a T-element finalize;
// the |gt_t| scratch is a stream-ordered allocation.
}  // extern "C"

template <typename CountFn>
static int jf_common(long long T, int* inter, int* uni, long long* tp_fp_fn, cudaStream_t stream, CountFn count) {
  SOLA_REQUIRE(T >= 0 && T < (1ll << 31), "jf: bad frame count");
  SOLA_REQUIRE(tp_fp_fn, "jf: null totals pointer");
  SOLA_CUDA(cudaMemsetAsync(tp_fp_fn, 0, 3 * sizeof(long long), stream));
  if (T == 0) return SOLA_OK;
  SOLA_REQUIRE(inter && uni, "jf: null pointer");
  int* area_gt = nullptr;
  SOLA_CUDA(cudaMallocAsync(&area_gt, sizeof(int) * (size_t)T, stream));
  int rc = count(inter, uni, area_gt);
  if (rc == SOLA_OK) {
    const int blocks = (int)((T + 255) / 256 < 1024 ? (T + 255) / 256 : 1024);
    jf_finalize_kernel<<<blocks, 256, 0, stream>>>(inter, uni, area_gt, T, reinterpret_cast<unsigned long long*>(tp_fp_fn));
    rc = check_launch("jf_finalize kernel");
  }
  cudaFreeAsync(area_gt, stream);
  return rc;
}

extern "C" {

int sola_jf_f32(const float* pred, const float* gt, long long T, long long frame_px, int* inter, int* uni, long long* tp_fp_fn,
                cudaStream_t stream) {
  return jf_common(T, inter, uni, tp_fp_fn, stream,
                   [&](int* i, int* a, int* b) { return launch_raw_counts<float>(pred, gt, T, frame_px, i, a, b, stream); });
}

int sola_jf_u8(const uint8_t* pred, const uint8_t* gt, long long T, long long frame_px, int* inter, int* uni, long long* tp_fp_fn,
               cudaStream_t stream) {
  return jf_common(T, inter, uni, tp_fp_fn, stream,
                   [&](int* i, int* a, int* b) { return launch_raw_counts<uint8_t>(pred, gt, T, frame_px, i, a, b, stream); });
}

int sola_jf_packed(const uint32_t* pred, const uint32_t* gt, long long T, long long frame_words, int* inter, int* uni,
                   long long* tp_fp_fn, cudaStream_t stream) {
  return jf_common(T, inter, uni, tp_fp_fn, stream, [&](int* i, int* a, int* b) {
    return sola_frame_counts_packed(pred, gt, 1, 1, (int)T, frame_words, i, a, b, stream);
  });
}

int sola_or_merge(const uint32_t* tracks, const uint8_t* select, int K, long long words, uint32_t* out, cudaStream_t stream) {
  SOLA_REQUIRE(tracks && out, "or_merge: null pointer");
  SOLA_REQUIRE(K >= 0 && words >= 0, "or_merge: bad shape");
  if (words == 0) return SOLA_OK;
  long long blocks = (words + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  or_merge_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tracks, select, K, words, out);
  return check_launch("or_merge kernel");
}

}  // extern "C"
