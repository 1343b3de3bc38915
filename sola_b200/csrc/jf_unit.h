// One (video, expression) unit of a J&F sweep — the same 64-byte struct include/sola_maskpath.h declares as `sola_jf_unit`
// (the public header is not included here: it spells the stream type as void*).
#pragma once
#include <stdint.h>

extern "C" {
typedef struct sola_jf_unit {
  const uint32_t* pred;   // device, (T, H, Wp) bit-packed prediction planes
  const uint32_t* gt;     // device, (T, H, Wp) bit-packed ground-truth planes
  long long out_off;      // column of this unit's frame 0 in the (7, total_frames) output   [filled by sola_jf_sweep_plan]
  long long item0;        // first work item (frame x row band) of this unit                 [filled by sola_jf_sweep_plan]
  int T, H, W;
  int radius;             // boundary disk radius (bound_pix); < 0: region counts only
  int band_rows, n_bands; // row-band split of one frame                                      [filled by sola_jf_sweep_plan]
  int reserved0, reserved1;
} sola_jf_unit;
}
static_assert(sizeof(sola_jf_unit) == 64, "sola_jf_unit must stay 64 bytes (include/sola_maskpath.h, sola_b200/packed.py)");
