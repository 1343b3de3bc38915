// Boundary F-measure counts on bit-packed planes (extension: BASELINE.json's "seg2bmap boundary plus disk-dilation match").
//
// The reference tree has no boundary measure (its F is the volumetric pixel F1, evaluator.py:239-247 -> counts.cu); this
// kernel implements the public DAVIS definition restated in oracle/boundary_oracle.py:
//   bmap   = (seg ^ east) | (seg ^ south) | (seg ^ south-east), with the last-row / last-column / corner rules;
//   dil    = dilate(bmap, disk(r)) with zero padding;   r = ceil(0.008 * hypot(H, W))
//   counts = |bmap_fg|, |bmap_gt|, |bmap_fg & dil_gt|, |bmap_gt & dil_fg|      (per frame, exact int32)
//
// Everything is bit-parallel on 32-pixel words:
//   * phase 1: each CTA builds the boundary maps of its tile (32 rows x 30 word-columns, plus a halo of r rows and
//     one word-column on each side) in shared memory straight from the packed masks;
//   * phase 2: the disk is decomposed by column offset: pixel (dx, dy) is inside iff |dy| <= v[|dx|],
//     v[k] = floor(sqrt(r^2 - k^2)).  So dil = OR_k shift(+-k)( vertical_dilate(bmap, v[k]) ).  Walking k from r
//     down to 0 the vertical dilation only ever grows, so each thread keeps three running words (left, centre,
//     right neighbour columns) and spends 2 row-ORs per extra row plus two funnel shifts per k.
#include "common.cuh"

namespace sola {

constexpr int BT_ROWS = 32;       // output rows per CTA
constexpr int BT_COLS = 30;       // output word-columns per CTA (+2 halo = 32 staged)
constexpr int BT_SCOLS = BT_COLS + 2;
constexpr int B_MAX_R = 31;

struct DiskSpec {
  int r;
  unsigned char v[B_MAX_R + 1];   // v[k] = vertical half-extent at column offset k
};

__device__ __forceinline__ uint32_t ldw(const uint32_t* __restrict__ plane, int H, int Wp, int y, int c) {
  return (y >= 0 && y < H && c >= 0 && c < Wp) ? __ldg(plane + (long long)y * Wp + c) : 0u;
}

// boundary word at (y, c) from the packed mask plane
__device__ __forceinline__ uint32_t bmap_word(const uint32_t* __restrict__ seg, int H, int W, int Wp, int y, int c) {
  if (y < 0 || y >= H || c < 0 || c >= Wp) return 0u;
  const uint32_t s0 = ldw(seg, H, Wp, y, c), s0n = ldw(seg, H, Wp, y, c + 1);
  const uint32_t s1 = ldw(seg, H, Wp, y + 1, c), s1n = ldw(seg, H, Wp, y + 1, c + 1);
  const uint32_t e = (s0 >> 1) | (s0n << 31);      // east neighbour  (x+1, y)   ; zero beyond the last column (pad bits are 0)
  const uint32_t s = s1;                           // south neighbour (x, y+1)   ; zero below the last row
  const uint32_t se = (s1 >> 1) | (s1n << 31);     // south-east      (x+1, y+1)
  const int last_x = W - 1;
  const uint32_t last_col_bit = (c == (last_x >> 5)) ? (1u << (last_x & 31)) : 0u;
  uint32_t b;
  if (y == H - 1) {
    b = (s0 ^ e) & ~last_col_bit;                  // last row: seg ^ east ; corner forced to 0
  } else {
    b = (s0 ^ e) | (s0 ^ s) | (s0 ^ se);
    b = (b & ~last_col_bit) | ((s0 ^ s) & last_col_bit);   // last column: seg ^ south
  }
  // pad bits (x >= W) must stay clear: seg pads are 0 so b pads are 0 already
  return b;
}

__global__ void __launch_bounds__(256)
boundary_counts_kernel(const uint32_t* __restrict__ pred, const uint32_t* __restrict__ gt, int H, int W, DiskSpec disk,
                       int* __restrict__ n_fg, int* __restrict__ n_gt, int* __restrict__ fg_match, int* __restrict__ gt_match) {
  extern __shared__ uint32_t sm[];
  const int r = disk.r;
  const int Wp = (W + 31) >> 5;
  const int srows = BT_ROWS + 2 * r;
  uint32_t* b_fg = sm;                             // [srows][BT_SCOLS]
  uint32_t* b_gt = sm + srows * BT_SCOLS;
  const long long frame = blockIdx.z;
  const int y0 = blockIdx.y * BT_ROWS, c0 = blockIdx.x * BT_COLS;
  const uint32_t* pf = pred + frame * (long long)H * Wp;
  const uint32_t* pg = gt + frame * (long long)H * Wp;

  for (int i = threadIdx.x; i < srows * BT_SCOLS; i += blockDim.x) {
    const int sr = i / BT_SCOLS, sc = i - sr * BT_SCOLS;
    const int y = y0 - r + sr, c = c0 - 1 + sc;
    b_fg[i] = bmap_word(pf, H, W, Wp, y, c);
    b_gt[i] = bmap_word(pg, H, W, Wp, y, c);
  }
  __syncthreads();

  int c_fg = 0, c_gt = 0, c_fm = 0, c_gm = 0;
  for (int i = threadIdx.x; i < BT_ROWS * BT_COLS; i += blockDim.x) {
    const int ty = i / BT_COLS, tc = i - ty * BT_COLS;
    const int y = y0 + ty, c = c0 + tc;
    if (y >= H || c >= Wp) continue;
    const int sr = ty + r, sc = tc + 1;
    const uint32_t* rf = b_fg + sr * BT_SCOLS + sc;
    const uint32_t* rg = b_gt + sr * BT_SCOLS + sc;
    // running vertical dilations (left, centre, right) for both planes, currently of half-height `vh`
    uint32_t fl = rf[-1], fc = rf[0], fr = rf[1];
    uint32_t gl = rg[-1], gc = rg[0], gr = rg[1];
    uint32_t dil_f = 0, dil_g = 0;
    int vh = 0;
    for (int k = r; k >= 0; --k) {
      const int vk = disk.v[k];
      while (vh < vk) {
        ++vh;
        const uint32_t* uf = rf - vh * BT_SCOLS; const uint32_t* df = rf + vh * BT_SCOLS;
        const uint32_t* ug = rg - vh * BT_SCOLS; const uint32_t* dg = rg + vh * BT_SCOLS;
        fl |= uf[-1] | df[-1]; fc |= uf[0] | df[0]; fr |= uf[1] | df[1];
        gl |= ug[-1] | dg[-1]; gc |= ug[0] | dg[0]; gr |= ug[1] | dg[1];
      }
      if (k == 0) {
        dil_f |= fc; dil_g |= gc;
      } else {
        // source at x-k lands on x (shift towards higher bits, refill from the left word) and x+k (the mirror)
        dil_f |= __funnelshift_l(fl, fc, k) | __funnelshift_r(fc, fr, k);
        dil_g |= __funnelshift_l(gl, gc, k) | __funnelshift_r(gc, gr, k);
      }
    }
    const uint32_t bf = rf[0], bg = rg[0];
    c_fg += __popc(bf); c_gt += __popc(bg);
    c_fm += __popc(bf & dil_g); c_gm += __popc(bg & dil_f);
  }

  __shared__ int red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  c_fg = warp_sum(c_fg); c_gt = warp_sum(c_gt); c_fm = warp_sum(c_fm); c_gm = warp_sum(c_gm);
  if (lane == 0) { red[0][warp] = c_fg; red[1][warp] = c_gt; red[2][warp] = c_fm; red[3][warp] = c_gm; }
  __syncthreads();
  if (threadIdx.x < 4) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    int* out = threadIdx.x == 0 ? n_fg : (threadIdx.x == 1 ? n_gt : (threadIdx.x == 2 ? fg_match : gt_match));
    if (s) atomicAdd(out + frame, s);
  }
}

}  // namespace sola

using namespace sola;

extern "C" int sola_boundary_counts(const uint32_t* pred, const uint32_t* gt, long long n_frames, int H, int W, int radius,
                                    int* n_fg, int* n_gt, int* fg_match, int* gt_match, cudaStream_t stream) {
  SOLA_REQUIRE(pred && gt && n_fg && n_gt && fg_match && gt_match, "boundary_counts: null pointer");
  SOLA_REQUIRE(n_frames >= 0 && H > 0 && W > 0, "boundary_counts: bad shape");
  SOLA_REQUIRE(radius >= 0, "boundary_counts: negative radius");
  if (radius > B_MAX_R) {
    set_error("boundary_counts: radius %d > %d not supported (frame diagonal above ~3900 px)", radius, B_MAX_R);
    return SOLA_ERR_UNSUPPORTED;
  }
  if (n_frames == 0) return SOLA_OK;
  SOLA_REQUIRE(n_frames <= 65535, "boundary_counts: at most 65535 frames per launch");
  for (int* p : {n_fg, n_gt, fg_match, gt_match}) SOLA_CUDA(cudaMemsetAsync(p, 0, sizeof(int) * n_frames, stream));
  DiskSpec disk;
  disk.r = radius;
  for (int k = 0; k <= B_MAX_R; ++k) {
    int v = 0;
    if (k <= radius) {
      const int rem = radius * radius - k * k;
      while ((v + 1) * (v + 1) <= rem) ++v;
    }
    disk.v[k] = (unsigned char)v;
  }
  const int Wp = (W + 31) >> 5;
  dim3 grid((Wp + BT_COLS - 1) / BT_COLS, (H + BT_ROWS - 1) / BT_ROWS, (unsigned)n_frames);
  const size_t smem = (size_t)2 * (BT_ROWS + 2 * radius) * BT_SCOLS * sizeof(uint32_t);
  boundary_counts_kernel<<<grid, 256, smem, stream>>>(pred, gt, H, W, disk, n_fg, n_gt, fg_match, gt_match);
  return check_launch("boundary_counts kernel");
}
