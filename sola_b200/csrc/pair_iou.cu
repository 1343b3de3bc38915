// K2 — AND-popcount pairwise mask IoU on bit-packed tracks.
//
//   sola_pair_iou_st      full spatio-temporal N x N intersection matrix; semantics of
//                         seg_utils.compute_masklet_iou (track_generation/seg_utils.py:110-125) for every pair,
//                         with exact int64 counts instead of the reference's fp32 sums.
//   sola_pair_iou_gather  the matrix the reference greedy actually walks (generate_tokens_grid.py:266-278,
//                         generate_tokens_gdino.py:288-300): inter[i][j] = |track_i[frame_idx[j]] ∩ prompt_j|.
//
// The N x N kernel is integer-pipe bound for N >~ 16 (each word is used by N-1 pairs), so it is organised like a
// register-tiled GEMM whose multiply-add is LOP3+POPC+IADD:
//   * a CTA owns one 64 x 64 tile of pairs and every splits-th 32-word stage of the word (k) axis (a strided walk, so that every
//     CTA sees the same mix of background and object rows); the K axis is split so that the grid fills every SM several times
//     even when N = 64 gives a single tile;
//   * rows are staged global -> shared with 16-byte cp.async in a 4-stage ring, stored k-quad-major
//     ([k/4][row] uint4) so that the 16 lanes that read 16 consecutive rows hit 16 distinct bank groups;
//   * each of the 512 consumer threads accumulates a strided 4 x 2 micro-tile (rows warp+16r, cols lane+32c) in carry-save
//     registers; diagonal tiles stage their 64 rows once and skip the blocks below the diagonal;
//   * partial sums leave the CTA as 64-bit atomics on the N x N output (a few hundred adds per address).
#include "tma.cuh"
#include "csa.cuh"
#include <cuda.h>
#include <stdlib.h>          // CUtensorMap + enums only; the encoder is fetched at run time through cudaGetDriverEntryPoint

namespace sola {

constexpr int PT = 64;            // tile side (tracks)
constexpr int KQ = 8;             // uint4 per row per stage -> 32 words = 128 B per row per stage
constexpr int STAGE_WORDS = KQ * 4;
constexpr int K2_CTAS_PER_SM_TOTAL = 4;   // CTAs launched per SM over all tiles (two are resident at a time); 2 / 6 / 8 measured the same within 1 %
constexpr int NSTAGE = 4;         // ring depth; 5 and 6 stages measured 2 % slower with the 16-warp layout (profiles/r3_build_constants.json)
constexpr int ST_THREADS = 512;   // consumer threads: 16 warps (ty) x 32 lanes (tx), each a strided 4 x 2 micro-tile of pairs

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// tile list: (ti, tj) with ti <= tj, enumerated row-major over the upper triangle
__device__ __forceinline__ void tile_from_index(int idx, int nt, int& ti, int& tj) {
  int i = 0;
  while (idx >= nt - i) { idx -= nt - i; ++i; }
  ti = i; tj = i + idx;
}

// ---- the one consumer: a stage (64 or 2 x 64 rows x 32 words, in shared memory) into the thread's 4 x 4 carry-save accumulators ----
// Where a stage's 16-byte chunk q of operand row r lives depends on who staged it:
struct StageKQ {        // cp.async loader: k-quad major, [q][row]; rows of operand B follow the 64 rows of A (same rows when DIAG)
  const uint4* buf; int rows, b_off;
  __device__ __forceinline__ uint4 a(int q, int r) const { return buf[q * rows + r]; }
  __device__ __forceinline__ uint4 b(int q, int r) const { return buf[q * rows + b_off + r]; }
};
struct StageSwz {       // TMA loader, SWIZZLE_128B: chunk q of row r sits at chunk q ^ (r & 7) of the row's 128 bytes
  const uint4* A; const uint4* B;
  __device__ __forceinline__ uint4 a(int q, int r) const { return A[r * KQ + (q ^ (r & 7))]; }
  __device__ __forceinline__ uint4 b(int q, int r) const { return B[r * KQ + (q ^ (r & 7))]; }
};

// Thread (ty = warp, tx = lane) owns the pairs (row ty + 16 r, column tx + 32 c), r < 4, c < 2: the four operand rows of a warp are
// the same for all of its lanes (broadcast shared loads), its 32 lanes read 32 consecutive column rows (512 contiguous or
// swizzle-spread bytes: conflict-free).  Sixteen consumer warps per CTA (two CTAs per SM: 8 warps per scheduler) hide the
// fixed-latency LOP3 chains and the shared loads that the former 8-warp / 4 x 4 layout left exposed.
struct K2Acc {
  Csa c[4][2];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) c[a][b] = Csa{0u, 0u, 0};
  }
};

// On a diagonal tile the block (rows 16 r .., columns 32 c ..) is needed only where column >= row: 6 of the 8 blocks.  (Remapping the
// diagonal tile's 2080 pairs onto 5 slots with per-lane operand rows executes fewer chains but needs non-broadcast shared loads and
// measured slower, 0.32 vs 0.26 ms: profiles/r3_k2_variants.json.)
__host__ __device__ constexpr bool k2_block_needed(bool diag, int r, int c) { return !diag || 32 * c + 31 >= 16 * r; }

template <bool DIAG, class Stage>
__device__ __forceinline__ void consume_stage(const Stage& st, int tx, int ty, K2Acc& acc) {
  // Masklets are mostly background: an all-zero quad of operand row i contributes nothing to any of its pairs (the carry-save state
  // is unchanged by zero inputs), so its compressor chains are skipped.  The rows are warp-uniform, so the 4 rows x 8 quads = 32
  // tests of a stage are ONE per lane and a ballot: bit 8 r + q <=> quad q of row ty + 16 r is non-zero.  Skips never diverge.
  const uint4 probe = st.a(tx & 7, ty + 16 * (tx >> 3));
  const unsigned nz = __ballot_sync(FULL, (probe.x | probe.y | probe.z | probe.w) != 0u);
#pragma unroll
  for (int q = 0; q < KQ; ++q) {
    if (((nz >> q) & 0x01010101u) == 0u) continue;           // no row of this warp has anything in quad q
    uint4 b[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) b[c] = st.b(q, tx + 32 * c);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (!((nz >> (8 * r + q)) & 1u)) continue;
      const uint4 a = st.a(q, ty + 16 * r);
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (k2_block_needed(DIAG, r, c)) csa_quad(acc.c[r][c], a, b[c]);
    }
  }
}

// partial sums leave the CTA as 64-bit atomics on the N x N output (both mirror halves)
template <bool DIAG>
__device__ __forceinline__ void store_tile(const K2Acc& acc, int N, int ti, int tj, int tx, int ty, unsigned long long* __restrict__ inter) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (!k2_block_needed(DIAG, r, c)) continue;
      const int i = ti * PT + ty + 16 * r, j = tj * PT + tx + 32 * c;
      if (i >= N || j >= N) continue;
      if (DIAG && j < i) continue;                     // below the diagonal: the mirror entry is produced by another thread
      const unsigned long long v = (unsigned long long)csa_total(acc.c[r][c]);
      if (v == 0) continue;
      atomicAdd(inter + (long long)i * N + j, v);
      if (i != j) atomicAdd(inter + (long long)j * N + i, v);
    }
}

// Which stages of [stage_lo, stage_hi) the CTA `split` of `splits` reduces: every splits-th one.  Masklets are objects on an empty
// background, so a CONTIGUOUS K-range per CTA gives some CTAs only all-zero rows (top / bottom of a frame) and others the object
// centres — with two waves of CTAs the slow ones set the kernel time.  A strided walk hands every CTA the same mix of rows; each
// stage is its own TMA box / cp.async batch, so the order costs nothing.
struct StageWalk { long long first, step, count; };
__device__ __forceinline__ StageWalk stage_walk(long long stage_lo, long long stage_hi, int split, int splits) {
  const long long stages = stage_hi - stage_lo;
  return StageWalk{stage_lo + split, (long long)splits, stages > split ? (stages - split + splits - 1) / splits : 0};
}

// ---- cp.async-staged kernel: rows come from one (N, words) buffer or from a table of per-track pointers (peer GPUs' memory) --------
template <bool DIAG>
__device__ __forceinline__ void st_tile_cpasync(const uint32_t* __restrict__ packed, const uint32_t* const* __restrict__ row_ptrs, int N,
                                                long long words, int ti, int tj, long long s_first, long long s_step, long long n_st, uint4* smem,
                                                unsigned long long* __restrict__ inter) {
  constexpr int ROWS = DIAG ? PT : 2 * PT;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  K2Acc acc;
  acc.clear();
  // stage loader: ROWS*KQ 16-byte copies, thread t copies (row = c / KQ, q = c % KQ) for c = t, t+256, ...
  auto issue = [&](long long stage, int buf) {
    uint4* dst = smem + (size_t)buf * (2 * PT) * KQ;
    const long long w0 = stage * STAGE_WORDS;
#pragma unroll
    for (int c = tid; c < ROWS * KQ; c += ST_THREADS) {
      const int row = c / KQ, q = c % KQ;
      const int track = (row < PT) ? ti * PT + row : tj * PT + (row - PT);
      const long long w = w0 + q * 4;
      long long remain = (words - w) * 4;               // bytes left in this track row
      int nbytes = (track < N && remain > 0) ? (int)(remain < 16 ? remain : 16) : 0;
      const uint32_t* base = row_ptrs ? row_ptrs[track < N ? track : 0] : packed + (long long)(track < N ? track : 0) * words;
      const uint32_t* src = base + (nbytes ? w : 0);
      cp_async16(dst + q * ROWS + row, src, nbytes);
    }
  };
#pragma unroll
  for (int p = 0; p < NSTAGE - 1; ++p) {
    if (p < n_st) issue(s_first + p * s_step, p);
    cp_async_commit();
  }
  for (long long s = 0; s < n_st; ++s) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    if (s + NSTAGE - 1 < n_st) issue(s_first + (s + NSTAGE - 1) * s_step, (int)((s + NSTAGE - 1) % NSTAGE));
    cp_async_commit();
    consume_stage<DIAG>(StageKQ{smem + (size_t)(s % NSTAGE) * (2 * PT) * KQ, ROWS, DIAG ? 0 : PT}, tx, ty, acc);
  }
  cp_async_wait<0>();
  store_tile<DIAG>(acc, N, ti, tj, tx, ty, inter);
}

__global__ void __launch_bounds__(ST_THREADS)
pair_iou_st_kernel(const uint32_t* __restrict__ packed, const uint32_t* const* __restrict__ row_ptrs, int N, long long words, int nt,
                   int n_tiles, int splits, int tile_first, int tile_step, long long stage_lo, long long stage_hi,
                   unsigned long long* __restrict__ inter) {
  extern __shared__ uint4 smem_st[];
  // n_tiles = tiles this launch computes: global tile ids tile_first, tile_first + tile_step, ... (a rank's share);
  // [stage_lo, stage_hi) = the slice of the word axis this launch covers (a rank's share when the K axis is partitioned)
  const int tile = tile_first + (blockIdx.x % n_tiles) * tile_step, split = blockIdx.x / n_tiles;
  int ti, tj;
  tile_from_index(tile, nt, ti, tj);
  const StageWalk sw = stage_walk(stage_lo, stage_hi, split, splits);
  if (ti == tj) st_tile_cpasync<true>(packed, row_ptrs, N, words, ti, tj, sw.first, sw.step, sw.count, smem_st, inter);
  else st_tile_cpasync<false>(packed, row_ptrs, N, words, ti, tj, sw.first, sw.step, sw.count, smem_st, inter);
}

// ---- TMA-staged, warp-specialised ring (default) ---------------------------------------------------------------------------------
// The (rows x 32 words) stage tiles are regular, so they are fetched by the TMA unit: a rank-2 tensor map over packed[N][words]
// (uint32), box = 32 words x 64 rows = 8 KB, SWIZZLE_128B, hardware zero fill for rows >= N and for the K tail.  A 17th warp's elected
// lane is the producer (arms the stage's `full` mbarrier with the byte count, issues `cp.async.bulk.tensor.2d`); the 16 consumer
// warps hand each stage buffer back through an `empty` mbarrier (one arrive per warp), so there is no CTA-wide barrier in the stage
// loop and a warp that skipped many all-zero quads runs up to NSTAGE-1 stages ahead of its slowest sibling.
constexpr int RING_THREADS = ST_THREADS + 32;
constexpr int TMA_BOX_BYTES = PT * STAGE_WORDS * 4;      // 64 rows x 128 B

// `load_stage(dst, w0, bar)` issues the TMA loads of one stage (operand A at dst, B at dst + TMA_BOX_BYTES) after the expect_tx.
template <bool DIAG, class LoadStage>
__device__ __forceinline__ void st_tile_ring(LoadStage load_stage, int N, int ti, int tj, const StageWalk sw,
                                             unsigned char* smem, uint64_t* full, uint64_t* empty, unsigned long long* __restrict__ inter) {
  const int tid = threadIdx.x;
  const long long n_st = sw.count;
  if (tid >= ST_THREADS) {                              // producer warp: one lane drives the TMA unit
    if (tid == ST_THREADS) {
      for (long long s = 0; s < n_st; ++s) {
        const int buf = (int)(s % NSTAGE);
        if (s >= NSTAGE) mbar_wait(empty + buf, (unsigned)(((s / NSTAGE) - 1) & 1));     // all 16 consumer warps have left this buffer
        mbar_expect_tx(full + buf, DIAG ? TMA_BOX_BYTES : 2 * TMA_BOX_BYTES);
        load_stage(smem + (size_t)buf * 2 * TMA_BOX_BYTES, (int)((sw.first + s * sw.step) * STAGE_WORDS), full + buf);
      }
    }
    return;
  }
  const int tx = tid & 31, ty = tid >> 5, lane = tx;
  K2Acc acc;
  acc.clear();
  for (long long s = 0; s < n_st; ++s) {
    const int bufi = (int)(s % NSTAGE);
    mbar_wait(full + bufi, (unsigned)((s / NSTAGE) & 1));
    const uint4* A = reinterpret_cast<const uint4*>(smem + (size_t)bufi * 2 * TMA_BOX_BYTES);
    consume_stage<DIAG>(StageSwz{A, DIAG ? A : A + TMA_BOX_BYTES / 16}, tx, ty, acc);
    __syncwarp();                                       // every lane of this warp is done reading the buffer
    if (lane == 0) mbar_arrive(empty + bufi);
  }
  store_tile<DIAG>(acc, N, ti, tj, tx, ty, inter);
}

struct RingSetup { unsigned char* boxes; int ti, tj; StageWalk sw; };

__device__ __forceinline__ RingSetup ring_setup(unsigned char* dyn_smem, uint64_t* full, uint64_t* empty, int nt, int n_tiles, int splits,
                                                int tile_first, int tile_step, long long stage_lo, long long stage_hi) {
  if (threadIdx.x == 0) {
    for (int b = 0; b < NSTAGE; ++b) { mbar_init(full + b, 1); mbar_init(empty + b, ST_THREADS / 32); }
    mbar_fence_init();
  }
  __syncthreads();
  RingSetup r;
  const int tile = tile_first + (blockIdx.x % n_tiles) * tile_step, split = blockIdx.x / n_tiles;
  tile_from_index(tile, nt, r.ti, r.tj);
  r.sw = stage_walk(stage_lo, stage_hi, split, splits);
  // round the dynamic-smem base up to 1024 B by OFFSET (SWIZZLE_128B boxes; a pointer cast would demote the tile reads to generic loads)
  r.boxes = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
  return r;
}

__global__ void __launch_bounds__(RING_THREADS, 2)
pair_iou_st_ring_kernel(const __grid_constant__ CUtensorMap map, int N, long long words, int nt, int n_tiles, int splits,
                        int tile_first, int tile_step, unsigned long long* __restrict__ inter) {
  extern __shared__ __align__(1024) unsigned char smem_ring[];
  __shared__ uint64_t full[NSTAGE], empty[NSTAGE];
  const RingSetup r = ring_setup(smem_ring, full, empty, nt, n_tiles, splits, tile_first, tile_step, 0, (words + STAGE_WORDS - 1) / STAGE_WORDS);
  const CUtensorMap* m = &map;
  const int ti = r.ti, tj = r.tj;
  if (ti == tj) {
    st_tile_ring<true>([=](unsigned char* dst, int w0, uint64_t* bar) { tma_load_2d(dst, m, w0, ti * PT, bar); },
                       N, ti, tj, r.sw, r.boxes, full, empty, inter);
  } else {
    st_tile_ring<false>([=](unsigned char* dst, int w0, uint64_t* bar) {
      tma_load_2d(dst, m, w0, ti * PT, bar);
      tma_load_2d(dst + TMA_BOX_BYTES, m, w0, tj * PT, bar);
    }, N, ti, tj, r.sw, r.boxes, full, empty, inter);
  }
}

// ---- the same ring with a producer that loads the tiles straight out of PEER GPUs' memory ------------------------------------------
// One kernel that is both the exchange and the math of BASELINE config 5: every rank's packed planes sit in an NVLink-mapped buffer
// (n_local tracks x words), one rank-2 tensor map per rank; the producer lane splits each 64-row operand tile into pieces of
// `box_rows` rows (a divisor of 64 and of n_local, so a piece never straddles two ranks) and issues one `cp.async.bulk.tensor.2d`
// per piece against the owner's map — the TMA unit pulls the rows over NVLink while the 16 consumer warps reduce the previous stages.
// Rows beyond a rank's n_local (and the K tail) are zero-filled by the hardware and still count towards the stage's byte total.
struct PeerMaps { CUtensorMap m[8]; };

__global__ void __launch_bounds__(RING_THREADS, 2)
pair_iou_st_ring_peer_kernel(const __grid_constant__ PeerMaps maps, int world, int n_local, int box_rows, int nt, int n_tiles, int splits,
                             long long stage_lo, long long stage_hi, unsigned long long* __restrict__ inter) {
  extern __shared__ __align__(1024) unsigned char smem_peer[];
  __shared__ uint64_t full[NSTAGE], empty[NSTAGE];
  const RingSetup r = ring_setup(smem_peer, full, empty, nt, n_tiles, splits, 0, 1, stage_lo, stage_hi);
  const int N = world * n_local, ti = r.ti, tj = r.tj;
  const PeerMaps* pm = &maps;
  auto load_operand = [=](unsigned char* dst, int w0, int tile, uint64_t* bar) {
    const int pieces = PT / box_rows;
    for (int p = 0; p < pieces; ++p) {
      const int g0 = tile * PT + p * box_rows;                    // global index of the piece's first track
      int owner = g0 / n_local;
      if (owner > world - 1) owner = world - 1;                   // past the last track: out-of-range rows -> hardware zero fill
      tma_load_2d(dst + (size_t)p * box_rows * STAGE_WORDS * 4, &pm->m[owner], w0, g0 - owner * n_local, bar);
    }
  };
  if (ti == tj) {
    st_tile_ring<true>([=](unsigned char* dst, int w0, uint64_t* bar) { load_operand(dst, w0, ti, bar); },
                       N, ti, tj, r.sw, r.boxes, full, empty, inter);
  } else {
    st_tile_ring<false>([=](unsigned char* dst, int w0, uint64_t* bar) {
      load_operand(dst, w0, ti, bar);
      load_operand(dst + TMA_BOX_BYTES, w0, tj, bar);
    }, N, ti, tj, r.sw, r.boxes, full, empty, inter);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rank-2 map over packed[N][words] (uint32): box = STAGE_WORDS x PT, 128-byte swizzle, zero fill out of range
static bool make_track_map(const uint32_t* packed, int N, long long words, CUtensorMap* map) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || words >= (1ll << 31)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)words, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)words * 4};
  const cuuint32_t box[2] = {STAGE_WORDS, PT};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(packed), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Fallback for track rows that are not 16-byte aligned (words % 4 != 0 or odd base): one CTA per pair.
__global__ void __launch_bounds__(256)
pair_iou_st_simple_kernel(const uint32_t* __restrict__ packed, int N, long long words, unsigned long long* __restrict__ inter) {
  const int i = blockIdx.y, j = blockIdx.x;
  if (j < i) return;
  const uint32_t* a = packed + (long long)i * words;
  const uint32_t* b = packed + (long long)j * words;
  unsigned long long acc = 0;
  for (long long w = threadIdx.x; w < words; w += blockDim.x) acc += __popc(a[w] & b[w]);
  __shared__ unsigned long long red[8];
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    inter[(long long)i * N + j] = s;
    inter[(long long)j * N + i] = s;
  }
}

// ---- gathered single-frame IoU ------------------------------------------------------------------------------
constexpr int GT_TILE = 4;   // tracks per CTA
constexpr int GT_THREADS = 128;

// VEC = 4: planes are 16-byte aligned and FW % 4 == 0 -> 128-bit loads, 5 independent loads in flight per iteration
template <int VEC>
__global__ void __launch_bounds__(GT_THREADS)
pair_iou_gather_kernel(const uint32_t* __restrict__ tracks, const uint32_t* __restrict__ prompts, const int* __restrict__ frame_idx,
                       int N, int P, int T, int FW, int* __restrict__ inter, int* __restrict__ area_t, int* __restrict__ area_p) {
  const int j = blockIdx.x, i0 = blockIdx.y * GT_TILE;
  int f = frame_idx[j];
  f = f < 0 ? 0 : (f >= T ? T - 1 : f);
  const uint32_t* pp = prompts + (long long)j * FW;
  const uint32_t* pt[GT_TILE];
#pragma unroll
  for (int k = 0; k < GT_TILE; ++k) pt[k] = tracks + ((long long)min(i0 + k, N - 1) * T + f) * FW;
  int ni[GT_TILE] = {0, 0, 0, 0}, na[GT_TILE] = {0, 0, 0, 0}, np = 0;
  if (VEC == 4) {
    // (carry-save counters were measured here too: 59 vs 54 us — at 4050 quads per plane the kernel is latency-bound, not POPC-bound)
    const int nq = FW >> 2;
#pragma unroll 2
    for (int q = threadIdx.x; q < nq; q += GT_THREADS) {
      const uint4 p = __ldg(reinterpret_cast<const uint4*>(pp) + q);
      uint4 t[GT_TILE];
#pragma unroll
      for (int k = 0; k < GT_TILE; ++k) t[k] = __ldg(reinterpret_cast<const uint4*>(pt[k]) + q);
      np += __popc(p.x) + __popc(p.y) + __popc(p.z) + __popc(p.w);
#pragma unroll
      for (int k = 0; k < GT_TILE; ++k) {
        ni[k] += __popc(p.x & t[k].x) + __popc(p.y & t[k].y) + __popc(p.z & t[k].z) + __popc(p.w & t[k].w);
        na[k] += __popc(t[k].x) + __popc(t[k].y) + __popc(t[k].z) + __popc(t[k].w);
      }
    }
  } else {
    for (int w = threadIdx.x; w < FW; w += GT_THREADS) {
      const uint32_t p = pp[w];
      np += __popc(p);
#pragma unroll
      for (int k = 0; k < GT_TILE; ++k) {
        const uint32_t t = pt[k][w];
        ni[k] += __popc(p & t);
        na[k] += __popc(t);
      }
    }
  }
  __shared__ int red[2 * GT_TILE + 1][GT_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  np = warp_sum(np);
#pragma unroll
  for (int k = 0; k < GT_TILE; ++k) { ni[k] = warp_sum(ni[k]); na[k] = warp_sum(na[k]); }
  if (lane == 0) {
    red[2 * GT_TILE][warp] = np;
#pragma unroll
    for (int k = 0; k < GT_TILE; ++k) { red[k][warp] = ni[k]; red[GT_TILE + k][warp] = na[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * GT_TILE + 1) {
    const int k = threadIdx.x;
    int s = 0;
#pragma unroll
    for (int w = 0; w < GT_THREADS / 32; ++w) s += red[k][w];
    if (k < GT_TILE) {
      if (i0 + k < N) inter[(long long)(i0 + k) * P + j] = s;
    } else if (k < 2 * GT_TILE) {
      if (i0 + k - GT_TILE < N) area_t[(long long)(i0 + k - GT_TILE) * P + j] = s;
    } else if (blockIdx.y == 0) {
      area_p[j] = s;
    }
  }
}

}  // namespace sola

using namespace sola;

extern "C" {

static int launch_pair_iou_st(const uint32_t* packed, int N, long long words_per_track, long long* inter_out, long long* area_out,
                              int part, int n_parts, cudaStream_t stream, bool accumulate = false) {
  SOLA_REQUIRE(N >= 0 && words_per_track > 0, "pair_iou_st: bad shape N=%d words=%lld", N, words_per_track);
  SOLA_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "pair_iou_st: bad partition %d / %d", part, n_parts);
  if (N == 0) return SOLA_OK;
  SOLA_REQUIRE(packed && inter_out, "pair_iou_st: null pointer");
  if (!accumulate) SOLA_CUDA(cudaMemsetAsync(inter_out, 0, sizeof(long long) * (size_t)N * N, stream));
  const bool fast = (words_per_track % 4 == 0) && aligned16(packed);
  if (fast) {
    const int nt = (N + PT - 1) / PT;
    const int all_tiles = nt * (nt + 1) / 2;
    const int n_tiles = (all_tiles - part + n_parts - 1) / n_parts;          // tiles part, part + n_parts, ...
    if (n_tiles > 0) {
      const long long stages = (words_per_track + STAGE_WORDS - 1) / STAGE_WORDS;
      long long splits = ((long long)num_sms() * K2_CTAS_PER_SM_TOTAL + n_tiles - 1) / n_tiles;
      if (splits > stages) splits = stages;
      const long long min_splits = (stages + (1 << 20) - 1) >> 20;     // keep int32 partial sums (4 * acc4 + ...) below 2^31
      if (splits < min_splits) splits = min_splits;
      if (splits < 1) splits = 1;
      const size_t smem = (size_t)NSTAGE * (2 * PT) * KQ * sizeof(uint4);
      SOLA_REQUIRE(splits * n_tiles < (1ll << 31), "pair_iou_st: grid too large");
      CUtensorMap map;
      if (make_track_map(packed, N, words_per_track, &map)) {
        // TMA-staged tiles (UTMALDG), warp-specialised ring
        SOLA_CUDA(cudaFuncSetAttribute(pair_iou_st_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
        pair_iou_st_ring_kernel<<<(unsigned)(splits * n_tiles), RING_THREADS, smem + 1024, stream>>>(
            map, N, words_per_track, nt, n_tiles, (int)splits, part, n_parts, reinterpret_cast<unsigned long long*>(inter_out));
      } else {
        SOLA_CUDA(cudaFuncSetAttribute(pair_iou_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pair_iou_st_kernel<<<(unsigned)(splits * n_tiles), ST_THREADS, smem, stream>>>(
            packed, nullptr, N, words_per_track, nt, n_tiles, (int)splits, part, n_parts, 0, stages, reinterpret_cast<unsigned long long*>(inter_out));
      }
    }
  } else {
    SOLA_REQUIRE(n_parts == 1 && !accumulate, "pair_iou_st: the unaligned fallback does not support partitioning / accumulation");
    SOLA_REQUIRE(N <= 65535, "pair_iou_st: unaligned fallback supports N <= 65535");
    dim3 grid(N, N);
    pair_iou_st_simple_kernel<<<grid, 256, 0, stream>>>(packed, N, words_per_track, reinterpret_cast<unsigned long long*>(inter_out));
  }
  int rc = check_launch("pair_iou_st kernel");
  if (rc != SOLA_OK) return rc;
  if (area_out)   // area[i] = inter[i][i]; strided device-to-device copy of the diagonal
    SOLA_CUDA(cudaMemcpy2DAsync(area_out, sizeof(long long), inter_out, sizeof(long long) * ((size_t)N + 1), sizeof(long long), N,
                                cudaMemcpyDeviceToDevice, stream));
  return SOLA_OK;
}

int sola_pair_iou_st(const uint32_t* packed, int N, long long words_per_track, long long* inter_out, long long* area_out,
                     cudaStream_t stream) {
  return launch_pair_iou_st(packed, N, words_per_track, inter_out, area_out, 0, 1, stream);
}

// One rank's share of the N x N matrix: 64 x 64 tiles part, part + n_parts, ... of the upper triangle (both mirror halves are
// written); every other entry of inter_out is left 0, so summing the outputs of all parts (an all-reduce) gives the full matrix.
int sola_pair_iou_st_part(const uint32_t* packed, int N, long long words_per_track, int part, int n_parts, long long* inter_out,
                          cudaStream_t stream) {
  return launch_pair_iou_st(packed, N, words_per_track, inter_out, nullptr, part, n_parts, stream);
}

// Fused exchange + K2 for tracks spread over the GPUs of one box: row_ptrs[i] (device array of N device pointers) is the base of
// track i's packed planes, which may live in a PEER GPU's memory (mapped through NVLink / NVSwitch).  The kernel's cp.async stage
// loads read the peers directly, so the transfer overlaps the AND-popcount math stage by stage instead of preceding it as an NCCL
// all-gather.  The WORD axis is partitioned: part p computes every pair tile over words [stages*p/n, stages*(p+1)/n) * 32 — perfectly
// balanced for any N, and each rank pulls only 1/n_parts of every peer track (an all-to-all's volume, not an all-gather's).
// Intersections are sums over words, so the sum of all parts' outputs (one all-reduce) is the full matrix.  Rows must be 16-byte aligned.
int sola_pair_iou_st_rows(const uint32_t* const* row_ptrs, int N, long long words_per_track, int part, int n_parts, long long* inter_out,
                          cudaStream_t stream) {
  SOLA_REQUIRE(N >= 0 && words_per_track > 0 && words_per_track % 4 == 0, "pair_iou_st_rows: bad shape (words must be a multiple of 4)");
  SOLA_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "pair_iou_st_rows: bad partition %d / %d", part, n_parts);
  if (N == 0) return SOLA_OK;
  SOLA_REQUIRE(row_ptrs && inter_out, "pair_iou_st_rows: null pointer");
  SOLA_CUDA(cudaMemsetAsync(inter_out, 0, sizeof(long long) * (size_t)N * N, stream));
  const int nt = (N + PT - 1) / PT;
  const int n_tiles = nt * (nt + 1) / 2;
  const long long all_stages = (words_per_track + STAGE_WORDS - 1) / STAGE_WORDS;
  const long long stage_lo = all_stages * part / n_parts, stage_hi = all_stages * (part + 1) / n_parts;
  const long long stages = stage_hi - stage_lo;
  if (stages <= 0) return SOLA_OK;
  long long splits = ((long long)num_sms() * K2_CTAS_PER_SM_TOTAL + n_tiles - 1) / n_tiles;
  if (splits > stages) splits = stages;
  const long long min_splits = (stages + (1 << 20) - 1) >> 20;
  if (splits < min_splits) splits = min_splits;
  if (splits < 1) splits = 1;
  const size_t smem = (size_t)NSTAGE * (2 * PT) * KQ * sizeof(uint4);
  SOLA_CUDA(cudaFuncSetAttribute(pair_iou_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_iou_st_kernel<<<(unsigned)(splits * n_tiles), ST_THREADS, smem, stream>>>(
      nullptr, row_ptrs, N, words_per_track, nt, n_tiles, (int)splits, 0, 1, stage_lo, stage_hi,
      reinterpret_cast<unsigned long long*>(inter_out));
  return check_launch("pair_iou_st_rows kernel");
}

// Exchange + K2 in ONE kernel (validated on 2 x B200: tests/test_gpu_multirank.py, profiles/r2_cfg5_2gpu_tma.log).  bases_host = HOST array of `world` device pointers,
// rank r's (n_local, words_per_track) packed planes (peer-mapped symmetric memory); part p of n_parts covers its slice of the word
// axis as in sola_pair_iou_st_rows.  Returns SOLA_ERR_UNSUPPORTED when the shape cannot be tiled (gcd(64, n_local) < 8, world > 8).
int sola_pair_iou_st_peer(const uint32_t* const* bases_host, int world, int n_local, long long words_per_track, int part, int n_parts,
                          long long* inter_out, cudaStream_t stream) {
  SOLA_REQUIRE(bases_host && inter_out && world >= 1 && n_local >= 1 && words_per_track > 0 && words_per_track % 4 == 0,
               "pair_iou_st_peer: bad arguments");
  SOLA_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "pair_iou_st_peer: bad partition %d / %d", part, n_parts);
  int box_rows = PT;
  while (box_rows > 1 && n_local % box_rows != 0) box_rows >>= 1;
  EncodeTiledFn enc = tensor_map_encoder();
  if (world > 8 || box_rows < 8 || !enc || words_per_track >= (1ll << 31)) {
    set_error("pair_iou_st_peer: unsupported shape (world=%d, n_local=%d) or no TMA encoder", world, n_local);
    return SOLA_ERR_UNSUPPORTED;
  }
  const int N = world * n_local;
  SOLA_CUDA(cudaMemsetAsync(inter_out, 0, sizeof(long long) * (size_t)N * N, stream));
  PeerMaps maps;
  for (int r = 0; r < 8; ++r) {
    const uint32_t* base = bases_host[r < world ? r : world - 1];
    SOLA_REQUIRE(base && aligned16(base), "pair_iou_st_peer: rank %d base pointer is null or misaligned", r);
    const cuuint64_t dims[2] = {(cuuint64_t)words_per_track, (cuuint64_t)n_local};
    const cuuint64_t strides[1] = {(cuuint64_t)words_per_track * 4};
    const cuuint32_t box[2] = {STAGE_WORDS, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&maps.m[r], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("pair_iou_st_peer: cuTensorMapEncodeTiled failed for rank %d", r);
      return SOLA_ERR_CUDA;
    }
  }
  const int nt = (N + PT - 1) / PT;
  const int n_tiles = nt * (nt + 1) / 2;
  const long long all_stages = (words_per_track + STAGE_WORDS - 1) / STAGE_WORDS;
  const long long stage_lo = all_stages * part / n_parts, stage_hi = all_stages * (part + 1) / n_parts;
  const long long stages = stage_hi - stage_lo;
  if (stages <= 0) return SOLA_OK;
  long long splits = ((long long)num_sms() * K2_CTAS_PER_SM_TOTAL + n_tiles - 1) / n_tiles;
  if (splits > stages) splits = stages;
  const long long min_splits = (stages + (1 << 20) - 1) >> 20;
  if (splits < min_splits) splits = min_splits;
  if (splits < 1) splits = 1;
  const size_t smem = (size_t)NSTAGE * (2 * PT) * KQ * sizeof(uint4) + 1024;
  SOLA_CUDA(cudaFuncSetAttribute(pair_iou_st_ring_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_iou_st_ring_peer_kernel<<<(unsigned)(splits * n_tiles), RING_THREADS, smem, stream>>>(
      maps, world, n_local, box_rows, nt, n_tiles, (int)splits, stage_lo, stage_hi, reinterpret_cast<unsigned long long*>(inter_out));
  return check_launch("pair_iou_st_ring_peer kernel");
}

// inter_inout += the N x N intersections over these words (no memset): lets a caller walk the word axis in chunks — e.g. chunks
// pulled from peer GPUs on one stream while the previous chunk is being reduced on another (sharding.PeerPlanes).
int sola_pair_iou_st_accumulate(const uint32_t* packed, int N, long long words_per_track, long long* inter_inout, cudaStream_t stream) {
  return launch_pair_iou_st(packed, N, words_per_track, inter_inout, nullptr, 0, 1, stream, true);
}

namespace sola {
// dst (N, n_words) <- words [word_lo, word_lo + n_words) of every row in the pointer table: 128-bit streaming loads, which go over
// NVLink when the row lives in a peer GPU's memory.  grid.y = row, grid.x strides over the row's 16-byte vectors.
__global__ void __launch_bounds__(256)
pull_rows_kernel(const uint32_t* const* __restrict__ row_ptrs, long long word_lo, long long n_vec, uint4* __restrict__ dst) {
  const uint4* src = reinterpret_cast<const uint4*>(row_ptrs[blockIdx.y] + word_lo);
  uint4* out = dst + (long long)blockIdx.y * n_vec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {          // 4 independent loads in flight per thread (NVLink latency)
    const uint4 a = ld_stream_u4(src + i), b = ld_stream_u4(src + i + stride), c = ld_stream_u4(src + i + 2 * stride),
                d = ld_stream_u4(src + i + 3 * stride);
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n_vec; i += stride) out[i] = ld_stream_u4(src + i);
}
}  // namespace sola

int sola_pull_rows(const uint32_t* const* row_ptrs, int N, long long word_lo, long long n_words, uint32_t* dst, cudaStream_t stream) {
  SOLA_REQUIRE(N >= 0 && word_lo >= 0 && n_words >= 0 && word_lo % 4 == 0 && n_words % 4 == 0, "pull_rows: word range must be 16-byte aligned");
  if (N == 0 || n_words == 0) return SOLA_OK;
  SOLA_REQUIRE(row_ptrs && dst && aligned16(dst), "pull_rows: null / misaligned pointer");
  SOLA_REQUIRE(N <= 65535, "pull_rows: too many rows");
  const long long n_vec = n_words / 4;
  long long bx = (n_vec + 256 * 4 - 1) / (256 * 4);
  const long long cap = ((long long)num_sms() * 8 + N - 1) / N;         // ~8 CTAs per SM over the whole grid
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)N);
  sola::pull_rows_kernel<<<grid, 256, 0, stream>>>(row_ptrs, word_lo, n_vec, reinterpret_cast<uint4*>(dst));
  return check_launch("pull_rows kernel");
}

int sola_pair_iou_gather(const uint32_t* tracks, const uint32_t* prompts, const int* frame_idx, int N, int P, int T,
                         long long frame_words, int* inter, int* area_t, int* area_p, cudaStream_t stream) {
  SOLA_REQUIRE(tracks && prompts && frame_idx && inter && area_t && area_p, "pair_iou_gather: null pointer");
  SOLA_REQUIRE(N >= 0 && P >= 0 && T > 0 && frame_words > 0 && frame_words < (1ll << 26), "pair_iou_gather: bad shape");
  if (N == 0 || P == 0) return SOLA_OK;
  SOLA_REQUIRE((N + GT_TILE - 1) / GT_TILE <= 65535, "pair_iou_gather: too many tracks for one launch");
  dim3 grid(P, (N + GT_TILE - 1) / GT_TILE);
  if (frame_words % 4 == 0 && aligned16(tracks) && aligned16(prompts))
    pair_iou_gather_kernel<4><<<grid, GT_THREADS, 0, stream>>>(tracks, prompts, frame_idx, N, P, T, (int)frame_words, inter, area_t, area_p);
  else
    pair_iou_gather_kernel<1><<<grid, GT_THREADS, 0, stream>>>(tracks, prompts, frame_idx, N, P, T, (int)frame_words, inter, area_t, area_p);
  return check_launch("pair_iou_gather kernel");
}

}  // extern "C"
