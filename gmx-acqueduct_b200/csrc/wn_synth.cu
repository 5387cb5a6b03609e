/*
 * wn_synth.cu -- device instantiation of the synthetic water-box generator
 * (include/wn_synth.h).  Compiled with -fmad=false so that
 * every a*b+c stays two IEEE roundings and the float32 coordinates are
 * bit-identical to the host build (gcc -ffp-contract=off).  Benchmark and test
 * plumbing: the reference ships no trajectory.
 */
#include "wn_internal.cuh"
#include "wn_synth.h"

namespace {

__global__ void __launch_bounds__(256)
wn_k_synth(wn_synth_box box, long long frame0, int n_frames, float* __restrict__ frames)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_frames * box.n_waters;
    if (t >= total) return;
    const int f = (int)(t / box.n_waters);
    const int w = (int)(t % box.n_waters);
    float out9[9];
    wn_synth_water(&box, (uint64_t)(frame0 + f), w, out9);
    float* dst = frames + (size_t)t * 9;
#pragma unroll
    for (int c = 0; c < 9; ++c) dst[c] = out9[c];
}

}  // namespace

void wn_launch_synth(int n_waters, uint64_t seed, int model, int64_t frame0, int n_frames,
                     float* frames, cudaStream_t st)
{
    const wn_synth_box box = wn_synth_make_box(n_waters, seed, model);
    const long long total = (long long)n_frames * n_waters;
    if (total <= 0) return;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    wn_k_synth<<<blocks, 256, 0, st>>>(box, (long long)frame0, n_frames, frames);
}
