/* Synthetic gray video for tests, smoke() and bench.py: libmofreak_synth.so (mofreak_b200/synth/synth.cpp, plain C++,
 * no CUDA).  A library of its own so that processes which must not map the product (the reference arm of bench.py)
 * can still generate the same frames. */
#ifndef MOFREAK_SYNTH_H_
#define MOFREAK_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Deterministic "moving squares" clip: n_frames of w x h gray u8 starting at time t0 (definition in synth.cpp). */
void mfk_synth_clip(uint8_t* out, int w, int h, int n_frames, uint32_t seed, int t0);
#ifdef __cplusplus
}
#endif
#endif
