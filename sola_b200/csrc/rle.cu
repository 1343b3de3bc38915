// COCO-RLE <-> bit-packed planes on the device.
//
// Replaces the CPU pycocotools hops around the hot path:
//   decode  AlignDataset.rle_masklet_decode (dataloader.py:353-369), seg_utils.decode_rle_masklet (seg_utils.py:70-75)
//           -> the uint8 (T, H, W) arrays that get OR-merged and then copied H2D at 4 B/px (evaluator.py:199-200)
//   encode  seg_utils.encode_rle_masklet_torch (seg_utils.py:93-106): a 4 B/px D2H copy followed by per-frame encode
// COCO runs are in column-major order while the packed planes are row-major, so both directions go through a
// column-major bit plane (W columns x Hp words) and a 32x32 bit-tile transpose done with warp ballots:
//   decode: 1-runs (host-prefix-summed starts / ends, a few hundred per mask) -> word fills with atomicOr -> transpose
//   encode: transpose -> transition bits (b[q] ^ b[q-1] along the column-major order) -> ordered compaction of their
//           positions; the host turns consecutive positions into counts and the varint string (tiny, sequential).
// Only 1/32 of the mask bytes ever cross PCIe.
#include "common.cuh"

namespace sola {

// ---- 32x32 bit-tile transpose: in (n, A, Bw) words = A x B bits per plane  ->  out (n, B, Aw) -----------------------
__global__ void __launch_bounds__(256)
bit_transpose_kernel(const uint32_t* __restrict__ in, long long n_planes, int A, int B, uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int Bw = (B + 31) >> 5, Aw = (A + 31) >> 5;
  const long long tiles_per_plane = (long long)Aw * Bw;
  const long long total = n_planes * tiles_per_plane;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp0; t < total; t += n_warps) {
    const long long p = t / tiles_per_plane;
    const int rem = (int)(t - p * tiles_per_plane);
    const int ta = rem / Bw, tb = rem - ta * Bw;               // tile: rows a in [32 ta, +32), bit columns b in [32 tb, +32)
    const int a = ta * 32 + lane;
    const uint32_t w = (a < A) ? __ldg(in + (p * A + a) * Bw + tb) : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t col = __ballot_sync(FULL, (w >> k) & 1u);  // bit k of every row = output row (32 tb + k), bits = rows a
      if (lane == k) mine = col;
    }
    const int b = tb * 32 + lane;
    if (b < B) out[(p * B + b) * Aw + ta] = mine;
  }
}

// ---- decode: fill 1-runs into column-major planes (n, W, Hp) ---------------------------------------------------------
// run r of plane run_plane[r] covers flat column-major pixel indices [run_start[r], run_end[r])  (q = x * H + y)
__global__ void __launch_bounds__(256)
rle_fill_runs_kernel(const int* __restrict__ run_plane, const int* __restrict__ run_start, const int* __restrict__ run_end,
                     long long n_runs, int H, int W, uint32_t* __restrict__ colmajor) {
  const int Hp = (H + 31) >> 5;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_runs; r += (long long)gridDim.x * blockDim.x) {
    int q = run_start[r];
    const int qe = run_end[r];
    uint32_t* plane = colmajor + (long long)run_plane[r] * W * Hp;
    while (q < qe) {
      const int x = q / H, y0 = q - x * H;
      const int y1 = min(H, y0 + (qe - q));                      // segment [y0, y1) inside column x
      uint32_t* col = plane + (long long)x * Hp;
      for (int wi = y0 >> 5; wi <= (y1 - 1) >> 5; ++wi) {
        const int lo = max(y0, wi * 32) - wi * 32, hi = min(y1, wi * 32 + 32) - wi * 32;     // bits [lo, hi) of word wi
        const uint32_t m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        atomicOr(col + wi, m);
      }
      q += y1 - y0;
    }
  }
}

// ---- encode: ordered positions of the transition bits of each plane --------------------------------------------------
// colmajor (n, W, Hp); positions are flat column-major pixel indices q = x * H + y where b[q] != b[q-1] (b[-1] = 0).
// One CTA per plane; out_pos (n, cap) int32, out_n (n) = number of transitions found (may exceed cap: then only the first
// cap are stored and the host must fall back for that plane).
__global__ void __launch_bounds__(256)
rle_transitions_kernel(const uint32_t* __restrict__ colmajor, int H, int W, int cap, int* __restrict__ out_pos, int* __restrict__ out_n) {
  const int Hp = (H + 31) >> 5;
  const long long p = blockIdx.x;
  const uint32_t* plane = colmajor + p * W * Hp;
  int* pos = out_pos + p * cap;
  const int n_words = W * Hp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int warp_tot[8];
  __shared__ int base_shared;
  if (tid == 0) base_shared = 0;
  __syncthreads();
  const int last_bits = H - (Hp - 1) * 32;                       // valid bits in the last word of a column (1..32)
  for (int w0 = 0; w0 < n_words; w0 += 256) {
    const int wi = w0 + tid;
    uint32_t trans = 0;
    int q0 = 0;
    if (wi < n_words) {
      const int x = wi / Hp, k = wi - x * Hp;
      const uint32_t cur = plane[wi];
      // bit preceding this word in column-major order: last valid bit of the previous word (or of the previous column)
      uint32_t prev_bit = 0;
      if (k > 0) prev_bit = plane[wi - 1] >> 31;
      else if (x > 0) prev_bit = (plane[wi - 1] >> (last_bits - 1)) & 1u;
      const uint32_t valid = (k == Hp - 1 && last_bits < 32) ? ((1u << last_bits) - 1u) : 0xffffffffu;
      trans = (cur ^ ((cur << 1) | prev_bit)) & valid;
      q0 = x * H + k * 32;
    }
    // ordered compaction: exclusive scan of popcounts over the 256 words of this chunk
    const int cnt = __popc(trans);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int warp_base = 0;
    for (int k = 0; k < warp; ++k) warp_base += warp_tot[k];
    int o = base_shared + warp_base + incl - cnt;
    uint32_t t = trans;
    while (t) {
      const int b = __ffs(t) - 1;
      t &= t - 1;
      if (o < cap) pos[o] = q0 + b;
      ++o;
    }
    __syncthreads();
    if (tid == 255) base_shared = o;                              // thread 255 holds the chunk's inclusive total
    __syncthreads();
  }
  if (tid == 0) out_n[p] = base_shared;
}

static int blocks_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b > (1ll << 20)) b = 1ll << 20;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace sola

using namespace sola;

extern "C" {

// planes (n, A, Bw) holding A x B bits each -> (n, B, Aw)
int sola_bit_transpose(const uint32_t* in, long long n_planes, int A, int B, uint32_t* out, cudaStream_t stream) {
  SOLA_REQUIRE(n_planes >= 0 && A > 0 && B > 0, "bit_transpose: bad shape");
  if (n_planes == 0) return SOLA_OK;
  SOLA_REQUIRE(in && out, "bit_transpose: null pointer");
  const long long tiles = n_planes * ((A + 31) >> 5) * ((B + 31) >> 5);
  bit_transpose_kernel<<<blocks_for(tiles, 8), 256, 0, stream>>>(in, n_planes, A, B, out);
  return check_launch("bit_transpose kernel");
}

// decode: runs -> packed row-major planes (n, H, Wp).  scratch: (n, W, Hp) words, zeroed here.
int sola_rle_decode_runs(const int* run_plane, const int* run_start, const int* run_end, long long n_runs,
                         long long n_planes, int H, int W, uint32_t* scratch_colmajor, uint32_t* packed_out, cudaStream_t stream) {
  SOLA_REQUIRE(n_planes >= 0 && n_runs >= 0 && H > 0 && W > 0, "rle_decode_runs: bad shape");
  if (n_planes == 0) return SOLA_OK;
  SOLA_REQUIRE(scratch_colmajor && packed_out && (n_runs == 0 || (run_plane && run_start && run_end)), "rle_decode_runs: null pointer");
  const int Hp = (H + 31) >> 5;
  SOLA_CUDA(cudaMemsetAsync(scratch_colmajor, 0, sizeof(uint32_t) * (size_t)n_planes * W * Hp, stream));
  if (n_runs > 0) {
    rle_fill_runs_kernel<<<blocks_for(n_runs, 256), 256, 0, stream>>>(run_plane, run_start, run_end, n_runs, H, W, scratch_colmajor);
    int rc = check_launch("rle_fill_runs kernel");
    if (rc != SOLA_OK) return rc;
  }
  return sola_bit_transpose(scratch_colmajor, n_planes, W, H, packed_out, stream);   // (W rows x H bits) -> (H rows x W bits)
}

// encode: packed row-major planes -> ordered transition positions per plane.  scratch: (n, W, Hp) words.
int sola_rle_encode_transitions(const uint32_t* packed, long long n_planes, int H, int W, uint32_t* scratch_colmajor,
                                int cap, int* out_pos, int* out_n, cudaStream_t stream) {
  SOLA_REQUIRE(n_planes >= 0 && H > 0 && W > 0 && cap > 0, "rle_encode_transitions: bad shape");
  if (n_planes == 0) return SOLA_OK;
  SOLA_REQUIRE(packed && scratch_colmajor && out_pos && out_n, "rle_encode_transitions: null pointer");
  SOLA_REQUIRE(n_planes < (1ll << 31) && (long long)H * W < (1ll << 31), "rle_encode_transitions: too large");
  int rc = sola_bit_transpose(packed, n_planes, H, W, scratch_colmajor, stream);     // (H rows x W bits) -> (W rows x H bits)
  if (rc != SOLA_OK) return rc;
  rle_transitions_kernel<<<(unsigned)n_planes, 256, 0, stream>>>(scratch_colmajor, H, W, cap, out_pos, out_n);
  return check_launch("rle_transitions kernel");
}


// Host-only: COCO compressed-RLE strings -> ones-runs.  The reference decodes with pycocotools' C loop (dataloader.py:360,
// cocoapi maskApi.c rleFrString + rleDecode), which writes H*W bytes per frame; here the same sequential varint parse stops at the
// run list ([start, end) in flat column-major pixel index), which is what crosses PCIe and what rle_fill_runs_kernel consumes.
//   strings[f] / lens[f]: the `counts` string of frame f (ASCII, not NUL-terminated); plane_ids[f]: the output plane its runs belong to;
//   n_pixels = H*W (every frame must cover exactly that many pixels); run_* are HOST arrays of capacity `cap`.
// Returns SOLA_ERR_INVALID on a malformed string / wrong coverage / overflow of cap (*n_runs_out then holds the runs needed so far).
int sola_rle_strings_to_runs(const char* const* strings, const long long* lens, const int* plane_ids, long long n_frames, long long n_pixels,
                             int* run_plane, int* run_start, int* run_end, long long cap, long long* n_runs_out) {
  SOLA_REQUIRE(n_frames >= 0 && n_pixels >= 0 && n_pixels < (1ll << 31) && cap >= 0 && n_runs_out, "rle_strings_to_runs: bad arguments");
  SOLA_REQUIRE(n_frames == 0 || (strings && lens && plane_ids), "rle_strings_to_runs: null pointer");
  long long n = 0;
  for (long long f = 0; f < n_frames; ++f) {
    const unsigned char* s = reinterpret_cast<const unsigned char*>(strings[f]);
    const long long len = lens[f];
    long long prev1 = 0, prev2 = 0;          // counts[i-1], counts[i-2]
    long long pos = 0;                       // pixels covered so far
    long long i = 0, p = 0;
    while (p < len) {
      long long x = 0;
      int k = 0;
      bool more = true;
      while (more) {
        if (p >= len || k > 12) { set_error("rle_strings_to_runs: frame %lld: truncated / oversized varint", f); *n_runs_out = n; return SOLA_ERR_INVALID; }
        const long long c = (long long)s[p] - 48;
        x |= (c & 0x1f) << (5 * k);
        more = (c & 0x20) != 0;
        ++p; ++k;
        if (!more && (c & 0x10)) x |= -1ll << (5 * k);
      }
      if (i > 2) x += prev2;
      if (x < 0 || pos + x > n_pixels) { set_error("rle_strings_to_runs: frame %lld: run %lld overruns the frame", f, i); *n_runs_out = n; return SOLA_ERR_INVALID; }
      if ((i & 1) && x > 0) {                // odd-indexed counts are ones-runs
        if (n < cap) { run_plane[n] = plane_ids[f]; run_start[n] = (int)pos; run_end[n] = (int)(pos + x); }
        ++n;
      }
      pos += x;
      prev2 = prev1; prev1 = x;
      ++i;
    }
    if (pos != n_pixels) { set_error("rle_strings_to_runs: frame %lld covers %lld pixels, expected %lld", f, pos, n_pixels); *n_runs_out = n; return SOLA_ERR_INVALID; }
  }
  *n_runs_out = n;
  if (n > cap) { set_error("rle_strings_to_runs: %lld runs, capacity %lld", n, cap); return SOLA_ERR_INVALID; }
  return SOLA_OK;
}

}  // extern "C"
