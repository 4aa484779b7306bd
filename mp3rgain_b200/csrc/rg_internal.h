// rg_internal.h -- structs shared by the host orchestration and the sm_100a kernels.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define RG_HIST_BINS 12000
#define RG_HIST_OFF 2000

// One serial IIR chain = (track, channel[, time block]).  Built on the host once frame counts are known.
struct RgChain {
    const float* src;   // first sample fed to the filter (start of warm-up, or of the track)
    double* wsum;       // out: sum of squares of each 50 ms window this chain owns, in order
    float* wmax;        // out: max |input sample| of each of those windows (normalised units)
    uint32_t n_warm;    // samples run through the filter from zero state before accumulation starts
    uint32_t n_main;    // samples accumulated (whole windows, plus the track's final partial one)
    uint32_t win;       // window length in samples (sample_rate*50/1000, src/replaygain.rs:704)
    uint32_t rate_idx;  // row of the coefficient table
    uint32_t flags;     // bit0: the chain starts at the first sample of its track (reference state: all-zero histories)
    uint32_t pad;
};

// One track as the histogram / percentile kernel sees it.
struct RgTrackWin {
    const double* wsum0;   // per-window sums, channel 0
    const double* wsum1;   // channel 1 (== wsum0 for mono: add_mono_sample adds sq to both sums)
    const float* wmax0;    // per-window max |input|, channel 0 / 1
    const float* wmax1;
    uint32_t n_full;       // complete windows
    uint32_t rem;          // samples in the final partial window (0 = none)
    uint32_t win;
    uint32_t chain_begin;  // [chain_begin, chain_end) index the per-chain peak array
    uint32_t chain_end;
    uint32_t flags;        // bit0: add this track to the album accumulators
};

// Device-side per-track result (mirrors rg_track_result's computed fields).
struct RgTrackOut {
    double loudness_db;
    double gain_db;
    double peak;
    uint32_t n_windows;
    uint32_t n_dropped;
    uint32_t n_edge;
    uint32_t pad;
};
