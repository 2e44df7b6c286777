/* mofreak_b200 — test and measurement instruments of the C ABI (NOT part of the drop-in surface of mofreak_b200.h):
 * per-kernel accounting, stage dumps for the parity tests, the FREAK scale-index probe.  The synthetic clip
 * generator of tests/bench is a separate library with no device code: mofreak_synth.h / libmofreak_synth.so. */
#ifndef MOFREAK_B200_DEBUG_H_
#define MOFREAK_B200_DEBUG_H_

#include "mofreak_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- measurement ------------------------------------------------------------------------------------ */

/* Per-kernel accounting.  enable != 0 brackets every kernel launch with events from then on.
 * mfk_profile_read syncs and returns, per kernel family, launches and total device ms since enable. */
#define MFK_PROF_MAX 16
typedef struct mfk_prof_entry { const char* name; uint32_t launches; float total_ms; } mfk_prof_entry;
int  mfk_profile_enable(mfk_ctx* ctx, int enable);
int  mfk_profile_read(mfk_ctx* ctx, mfk_prof_entry* entries, int cap, int* n_entries);
/* Timeline of the spans recorded since mfk_profile_enable and not yet consumed by mfk_profile_read: one entry per
 * launcher call (kernel family), start/end in device ms relative to the last mfk_timer_start.  With the lanes overlapping
 * this is the Gantt chart of the launch sequences (the reference has no tracing).  Syncs. */
typedef struct mfk_span { const char* name; float start_ms, end_ms; } mfk_span;
int  mfk_profile_timeline(mfk_ctx* ctx, mfk_span* spans, int cap, int* n_spans);
/* serial != 0: consecutive launch sequences no longer overlap (one buffer set, the integral image on the main stream).
 * Throughput runs leave this off; the per-kernel accounting of mfk_profile_* is only a kernel's own time when it is on. */
int  mfk_set_serial(mfk_ctx* ctx, int serial);

/* ---- stage access for parity tests (device -> host copies of intermediates of the LAST batch) ------- */
/* what: 0 = pyramid image, 1 = score map S_full; layer 0..5; pair index within the last batch.
 * out must hold the layer's width*height bytes; dims (optional) receives {width, height}. */
int  mfk_debug_layer(mfk_ctx* ctx, int what, int layer, int pair, uint8_t* out, int* dims);
/* Keypoints of pair `pair` of the last batch BEFORE FREAK's border filter: n x {x, y, size, response,
 * layer} floats in reference order; writes at most cap, *n gets the full count. */
int  mfk_debug_keypoints(mfk_ctx* ctx, int pair, float* out, int cap, int* n);

/* Raw device counters of the context: [0] overflow flags, [1] rows accumulated, [2] pairs accumulated, [3] live candidates of
 * the last batch, [4..7] optional statistics (builds with -DMFK_PB_STATS: tie-resolution rounds, ties, max rounds, tie visits). */
int  mfk_debug_counters(mfk_ctx* ctx, uint32_t out[8]);

/* ---- utilities that are not device work ------------------------------------------------------------- */

/* FREAK scale index of a keypoint size through the context-independent threshold table the device
 * uses (exposed so tests can compare it with the libm expression it was derived from). */
int  mfk_freak_scale_index(float kp_size);

#ifdef __cplusplus
}
#endif
#endif /* MOFREAK_B200_DEBUG_H_ */
