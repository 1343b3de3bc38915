/* sola_maskpath.h — C ABI of libsola_maskpath.so: B200 (sm_100a) kernels for SOLA's masklet-scoring path.
 *
 * The reference (cvlab-kaist/SOLA) has no FFI layer: the path is plain Python/ATen.  Each entry point below
 * names the reference code it replaces (file:line under the reference tree); the Python mirror in
 * sola_b200/ keeps the reference signatures and calls these through ctypes.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless it says "host"; outputs are pre-allocated by the caller;
 *   - all work is enqueued on `stream` (a cudaStream_t; 0 = legacy default stream); no implicit synchronisation,
 *     no global state, re-entrant across streams / devices (the current CUDA device must own the pointers);
 *   - return value: 0 = ok, <0 = error (SOLA_ERR_*); sola_last_error_string() describes the last error of the
 *     calling thread; nothing throws;
 *   - bit-packed plane layout: a mask (H, W) is (H, Wp) uint32, Wp = (W + 31) / 32, bit b of word w of a row is
 *     pixel 32*w + b, pad bits of the last word are 0.  frame_words = H * Wp.
 *   - counts are exact integers: int32 per frame, int64 per volume.
 */
#ifndef SOLA_MASKPATH_H
#define SOLA_MASKPATH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOLA_OK 0
#define SOLA_ERR_INVALID (-1)
#define SOLA_ERR_CUDA (-2)
#define SOLA_ERR_UNSUPPORTED (-3)

typedef void* sola_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------------------- */
int sola_version(void);
const char* sola_last_error_string(void);
const char* sola_build_arch(void);               /* "sm_100a" */
const char* sola_build_digest(void);             /* sha256 of the CUDA sources + flags the library was built from */
unsigned long long sola_launch_count(void);      /* kernels launched through this library so far (process-wide) */

/* ---- K1: binarise + bit-pack + stability popcounts ---------------------------------------------------------
 * replaces  (out_mask_logits > 0.0).float() + torch.cat      track_generation/generate_tokens_grid.py:215,219,222,224
 *                                                            track_generation/generate_tokens_gdino.py:232,237,240,241
 *           PromptGenerator.get_stability_score              track_generation/prompt_generator.py:169-186
 * logits (n_frames, H, W); packed_out (n_frames, H, Wp) = logits > thr (strict; NaN -> 0);
 * cnt_hi/cnt_mid/cnt_lo [n_frames] = #(logit > thr+off), #(logit > thr), #(logit > thr-off); any output may be NULL.
 * Thresholds are rounded to fp32 before comparing, as numpy/torch do for a Python scalar. */
int sola_binarize_pack_f32(const float* logits, long long n_frames, int H, int W, double thr, double off,
                           uint32_t* packed_out, int* cnt_hi, int* cnt_mid, int* cnt_lo, sola_stream_t stream);
int sola_binarize_pack_bf16(const void* logits_bf16, long long n_frames, int H, int W, double thr, double off,
                            uint32_t* packed_out, int* cnt_hi, int* cnt_mid, int* cnt_lo, sola_stream_t stream);
/* single threshold (x > thr) -> packed + area; used for `interpolate(...) > 0.5` style binarisation of float planes */
int sola_threshold_pack_f32(const float* x, long long n_frames, int H, int W, double thr,
                            uint32_t* packed_out, int* area, sola_stream_t stream);
/* {0,1} masks (nonzero = foreground) -> packed + area.   fp32: masks as every reference function receives them;
 * u8: masks as dataloader.py:353-369 rle_masklet_decode yields them */
int sola_pack_mask_f32(const float* mask, long long n_frames, int H, int W, uint32_t* packed_out, int* area, sola_stream_t stream);
int sola_pack_mask_u8(const uint8_t* mask, long long n_frames, int H, int W, uint32_t* packed_out, int* area, sola_stream_t stream);
/* packed -> {0,1} planes, for drop-in return types (reshape_masklet returns fp32; get_sam2_masklet returns uint8) */
int sola_unpack_f32(const uint32_t* packed, long long n_frames, int H, int W, float* out, sola_stream_t stream);
int sola_unpack_u8(const uint32_t* packed, long long n_frames, int H, int W, uint8_t* out, sola_stream_t stream);
/* same with foreground = one_value (255 -> the PNG planes written by inference.py:89-91) */
int sola_unpack_u8_value(const uint32_t* packed, long long n_frames, int H, int W, int one_value, uint8_t* out, sola_stream_t stream);

/* ---- K3: per-frame |A∩B|, |A|, |B| ------------------------------------------------------------------------
 * replaces the mul/add/sum/.item() chains of
 *           Evaluator.compute_J / compute_F                  evaluator.py:227-247
 *           compute_mask_iou / compute_masklet_iou           track_generation/seg_utils.py:110-142
 *           compute_mask_iou_torch / compute_mask_metrics    track_generation/utils.py:65-75,132-174
 * raw planes: a, b (n_frames, frame_px) with nonzero = foreground; outputs int32 [n_frames]. */
int sola_frame_counts_f32(const float* a, const float* b, long long n_frames, long long frame_px,
                          int* inter, int* area_a, int* area_b, sola_stream_t stream);
int sola_frame_counts_u8(const uint8_t* a, const uint8_t* b, long long n_frames, long long frame_px,
                         int* inter, int* area_a, int* area_b, sola_stream_t stream);
/* packed, batched: a (Na, T, frame_words), b (Nb, T, frame_words) -> inter (Na, Nb, T), area_a (Na, T), area_b (Nb, T).
 * One launch labels every track against every GT object (generate_tokens_grid.py:253-264). */
int sola_frame_counts_packed(const uint32_t* a, const uint32_t* b, int Na, int Nb, int T, long long frame_words,
                             int* inter, int* area_a, int* area_b, sola_stream_t stream);
/* packed, ragged: frame f = words [word_offsets[f], word_offsets[f+1]) of both buffers (device int64 [n_frames+1]).
 * One launch covers a whole J&F sweep of differently-shaped units (evaluator.py:174-225). */
int sola_frame_counts_packed_ragged(const uint32_t* a, const uint32_t* b, const long long* word_offsets, int n_frames,
                                    int* inter, int* area_a, int* area_b, sola_stream_t stream);
/* The J&F accumulators of one (video, expression) unit: inter[t] = |pred_t ∩ gt_t|, uni[t] = |pred_t ∪ gt_t| (int32 [T]) and
 * tp_fp_fn int64 [3] = the volume sums behind Evaluator.compute_F (evaluator.py:239-247), exact.  J = mean_t(uni ? inter/uni : 1)
 * (evaluator.py:227-237).  Inputs: raw {0,1} planes (T, frame_px) fp32 / uint8 with nonzero = foreground, or bit-packed
 * (T, frame_words). */
int sola_jf_f32(const float* pred, const float* gt, long long T, long long frame_px, int* inter, int* uni, long long* tp_fp_fn,
                sola_stream_t stream);
int sola_jf_u8(const uint8_t* pred, const uint8_t* gt, long long T, long long frame_px, int* inter, int* uni, long long* tp_fp_fn,
               sola_stream_t stream);
int sola_jf_packed(const uint32_t* pred, const uint32_t* gt, long long T, long long frame_words, int* inter, int* uni,
                   long long* tp_fp_fn, sola_stream_t stream);
/* OR over the selected packed tracks: tracks (K, words), select u8 [K] (NULL = all) -> out (words).
 * replaces np.logical_or accumulation in dataloader.py:285-299 (GT objects) and :319-350 (selected SAM2 tracks). */
int sola_or_merge(const uint32_t* tracks, const uint8_t* select, int K, long long words, uint32_t* out, sola_stream_t stream);

/* ---- K2: pairwise mask IoU ---------------------------------------------------------------------------------
 * spatio-temporal N x N: packed (N, words_per_track) -> inter_out int64 (N, N) (symmetric, diagonal = area),
 * area_out int64 [N] (may be NULL).  Semantics of seg_utils.compute_masklet_iou (seg_utils.py:110-125) per pair. */
int sola_pair_iou_st(const uint32_t* packed, int N, long long words_per_track, long long* inter_out, long long* area_out,
                     sola_stream_t stream);
/* one rank's share of the same matrix (BASELINE config 5: tracks all-gathered, pair tiles split over the ranks): computes the
 * 64 x 64 tiles part, part + n_parts, ... of the upper triangle (mirrors written too), zeros elsewhere; the sum over parts
 * (one all-reduce) is the full matrix. */
int sola_pair_iou_st_part(const uint32_t* packed, int N, long long words_per_track, int part, int n_parts, long long* inter_out,
                          sola_stream_t stream);
/* fused exchange + K2: row_ptrs (device array of N device pointers) gives each track's packed planes, possibly in a PEER GPU's
 * memory; the kernel reads peers directly over NVLink so the transfer overlaps the math (no NCCL all-gather in front).  Here the
 * WORD axis is partitioned (part p covers 1/n_parts of every track), so the parts are balanced for any N and each rank pulls only
 * 1/n_parts of its peers' planes; the sum of all parts' outputs is the full matrix. */
int sola_pair_iou_st_rows(const uint32_t* const* row_ptrs, int N, long long words_per_track, int part, int n_parts, long long* inter_out,
                          sola_stream_t stream);
/* exchange + K2 in ONE kernel (validated on 2 x B200, tests/test_gpu_multirank.py).  bases_host is a HOST array of `world` device
 * pointers, rank r's (n_local, words_per_track) packed planes in peer-mapped memory; the K2 producer lane issues its TMA tile loads
 * against one tensor map per rank, so the rows cross NVLink inside the kernel that reduces them.  part / n_parts partition the
 * word axis as in sola_pair_iou_st_rows.  SOLA_ERR_UNSUPPORTED if gcd(64, n_local) < 8 or world > 8. */
int sola_pair_iou_st_peer(const uint32_t* const* bases_host, int world, int n_local, long long words_per_track, int part, int n_parts,
                          long long* inter_out, sola_stream_t stream);
/* inter_inout += intersections over these words (no memset): walk the word axis in chunks (e.g. pulled from peers, see below). */
int sola_pair_iou_st_accumulate(const uint32_t* packed, int N, long long words_per_track, long long* inter_inout, sola_stream_t stream);
/* dst (N, n_words) <- words [word_lo, word_lo + n_words) of every row of the pointer table (rows may live in peer GPUs' memory:
 * the loads then travel over NVLink).  With sola_pair_iou_st_accumulate on a second stream this is the pipelined exchange of
 * BASELINE config 5: chunk c+1 is pulled while chunk c is reduced by the TMA-staged K2 kernel. */
int sola_pull_rows(const uint32_t* const* row_ptrs, int N, long long word_lo, long long n_words, uint32_t* dst, sola_stream_t stream);
/* gathered single-frame: tracks (N, T, frame_words), prompts (P, frame_words), frame_idx int32 [P] ->
 * inter (N, P) = |track_i[frame_idx[j]] ∩ prompt_j|, area_t (N, P) = |track_i[frame_idx[j]]|, area_p [P].
 * This is the matrix walked by generate_tokens_grid.py:266-278 / generate_tokens_gdino.py:288-300. */
int sola_pair_iou_gather(const uint32_t* tracks, const uint32_t* prompts, const int* frame_idx, int N, int P, int T,
                         long long frame_words, int* inter, int* area_t, int* area_p, sola_stream_t stream);

/* ---- R1 / R2: resizes that feed the greedy filter ----------------------------------------------------------
 * R1 replaces seg_utils.reshape_masklet (seg_utils.py:145-160): bilinear (align_corners=False) to (oh, ow) then > 0.5,
 * reproducing ATen's CUDA upsample_bilinear2d fp32 arithmetic.  Input either bit-packed or fp32 planes.
 * out_packed (n_frames, oh, owp) and/or out_f32 (n_frames, oh, ow) {0,1}; area int32 [n_frames] (may be NULL). */
int sola_resize_bilinear_bin_packed(const uint32_t* in_packed, long long n_frames, int H, int W, int oh, int ow,
                                    uint32_t* out_packed, int* area, sola_stream_t stream);
int sola_resize_bilinear_bin_f32(const float* in, long long n_frames, int H, int W, int oh, int ow,
                                 uint32_t* out_packed, float* out_f32, int* area, sola_stream_t stream);
/* R2 replaces F.interpolate(prompt[None,None], (h,w), 'nearest') (generate_tokens_grid.py:271-272,
 * generate_tokens_gdino.py:293-294): legacy nearest, src = min(floor(dst * in/out), in-1); u8 or packed input. */
int sola_resize_nearest_u8(const uint8_t* in, long long n_frames, int H, int W, int oh, int ow,
                           uint32_t* out_packed, int* area, sola_stream_t stream);
int sola_resize_nearest_packed(const uint32_t* in_packed, long long n_frames, int H, int W, int oh, int ow,
                               uint32_t* out_packed, int* area, sola_stream_t stream);

/* K1 + R1 fused: one pass over the logits yields the full-resolution packed planes, the stability counts AND the
 * bilinear-resized packed planes (replaces generate_tokens_grid.py:215-224 + prompt_generator.py:169-186 + seg_utils.py:145-160).
 * resized_out (n_frames, oh, owp) is required; packed_out / counts / area_resized may be NULL.  Falls back internally to the
 * two separate kernels (same results) when W % 32 != 0 or the base pointer is not 16-byte aligned. */
int sola_binarize_pack_resize_f32(const float* logits, long long n_frames, int H, int W, int oh, int ow, double thr, double off,
                                  uint32_t* packed_out, uint32_t* resized_out, int* cnt_hi, int* cnt_mid, int* cnt_lo,
                                  int* area_resized, sola_stream_t stream);
int sola_binarize_pack_resize_bf16(const void* logits_bf16, long long n_frames, int H, int W, int oh, int ow, double thr, double off,
                                   uint32_t* packed_out, uint32_t* resized_out, int* cnt_hi, int* cnt_mid, int* cnt_lo,
                                   int* area_resized, sola_stream_t stream);

/* ---- COCO RLE <-> packed planes (replaces the pycocotools hops: dataloader.py:353-369, seg_utils.py:70-75,93-106) --------
 * sola_bit_transpose: planes (n, A, Bw) of A x B bits -> (n, B, Aw) (row-major <-> column-major bit planes).
 * sola_rle_decode_runs: ones-runs [run_start, run_end) in flat column-major pixel index, run_plane = plane of each run
 *   (device int32 arrays) -> packed_out (n, H, Wp); scratch_colmajor (n, W, Hp) words.
 * sola_rle_encode_transitions: packed (n, H, Wp) -> per plane the ordered column-major positions where the bit value changes
 *   (out_pos (n, cap) int32, out_n (n) = number found, may exceed cap); the host turns them into counts + the varint string. */
int sola_bit_transpose(const uint32_t* in, long long n_planes, int A, int B, uint32_t* out, sola_stream_t stream);
int sola_rle_decode_runs(const int* run_plane, const int* run_start, const int* run_end, long long n_runs,
                         long long n_planes, int H, int W, uint32_t* scratch_colmajor, uint32_t* packed_out, sola_stream_t stream);
/* host-only: the `counts` strings of n_frames RLE frames -> their ones-runs ([start, end) in flat column-major pixel index) in HOST
 * arrays of capacity cap; plane_ids[f] = output plane of frame f's runs (several tracks may share planes: sola_rle_decode_runs ORs them).
 * The sequential varint parse of cocoapi's rleFrString, stopped before the H*W-byte fill (dataloader.py:360). */
int sola_rle_strings_to_runs(const char* const* strings, const long long* lens, const int* plane_ids, long long n_frames, long long n_pixels,
                             int* run_plane, int* run_start, int* run_end, long long cap, long long* n_runs_out);
int sola_rle_encode_transitions(const uint32_t* packed, long long n_planes, int H, int W, uint32_t* scratch_colmajor,
                                int cap, int* out_pos, int* out_n, sola_stream_t stream);

/* ---- fused J & F: region counts + boundary-match counts in one pass over the packed planes ---------------------------------
 * replaces  Evaluator.compute_J / compute_F                  evaluator.py:227-247   (rows 0..2: |pred ∩ gt|, |pred|, |gt| per frame)
 *           the per-expression loop of compute_JF_metrics    evaluator.py:183-218   (one launch for a whole sweep of units)
 * and adds the boundary measure BASELINE.json's north-star names (seg2bmap + disk dilation + match counting; the reference tree has
 * no implementation — DAVIS definition, oracle/boundary_oracle.py):
 *           rows 3..6: |b(pred)|, |b(gt)|, |b(pred) & dilate(b(gt))|, |b(gt) & dilate(b(pred))|, radius = bound_pix.
 * counts_out is int32 (7, total_frames), frames numbered in unit order.  radius < 0 skips the boundary part (rows 3..6 stay 0).
 * Limits: radius <= 31, W <= 8192. */
typedef struct sola_jf_unit {
  const uint32_t* pred;   /* device, (T, H, Wp) bit-packed prediction planes */
  const uint32_t* gt;     /* device, (T, H, Wp) bit-packed ground-truth planes */
  long long out_off;      /* [filled by sola_jf_sweep_plan] column of this unit's frame 0 in counts_out */
  long long item0;        /* [filled by sola_jf_sweep_plan] first work item of this unit */
  int T, H, W;
  int radius;
  int band_rows, n_bands; /* [filled by sola_jf_sweep_plan] */
  int reserved0, reserved1;
} sola_jf_unit;           /* 64 bytes */
typedef struct sola_jf_plan {
  long long n_items, total_frames;      /* work items (frame x row band) and output columns of the sweep */
  int raw_cap, bm_cap, mask_steps;      /* shared-memory layout of the launch */
  int reserved;                         /* in: 0 (automatic) or the CTAs per SM to plan for (1..3); out: the class chosen — 4 = the
                                           region-only build (every unit has radius < 0: small tiles, four CTAs per SM) */
} sola_jf_plan;           /* 32 bytes */
/* host-only: plans the row-band split of every unit (HOST array, updated in place) and the launch's shared-memory layout */
int sola_jf_sweep_plan(sola_jf_unit* units_host, int n_units, sola_jf_plan* plan_out);
/* units_dev: the planned table copied to the device; plan: HOST pointer to what sola_jf_sweep_plan returned */
int sola_jf_sweep(const sola_jf_unit* units_dev, int n_units, const sola_jf_plan* plan, int* counts_out, sola_stream_t stream);
/* one unit, no table: pred, gt (n_frames, H, Wp) -> counts_out int32 (7, n_frames) */
int sola_jf_boundary_packed(const uint32_t* pred, const uint32_t* gt, long long n_frames, int H, int W, int radius,
                            int* counts_out, sola_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SOLA_MASKPATH_H */
