/*
 * wn_synth_host.c -- host instantiation of the synthetic frame generator
 * (include/wn_synth.h) for tests and the CPU baseline.  Not part of the oracle
 * algorithm; lives in the oracle library so CPU-only tests have a generator.
 */
#include <stddef.h>
#include "wn_synth.h"

void wno_synth_box(int32_t n_waters, uint64_t seed, int32_t model, wn_synth_box* out)
{
    *out = wn_synth_make_box(n_waters, seed, model);
}

/* frames [frame0, frame0+n_frames) into out[n_frames][3*n_waters][3] float32 */
void wno_synth_frames(const wn_synth_box* box, int64_t frame0, int32_t n_frames, float* out)
{
    int32_t f, w;
    for (f = 0; f < n_frames; ++f)
        for (w = 0; w < box->n_waters; ++w)
            wn_synth_water(box, (uint64_t)(frame0 + f), w,
                           out + ((size_t)f * (size_t)box->n_waters + (size_t)w) * 9);
}
