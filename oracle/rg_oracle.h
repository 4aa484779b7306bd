/*
 * rg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's ReplayGain 1.0 analysis path
 * (M-Igashi/mp3rgain v1.5.0, src/replaygain.rs).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (mp3rgain_b200/) never links, imports or calls it.
 *
 * Stage B (PCM -> histogram -> loudness) follows the reference line by line and is
 * pinned by the known answers derived from the reference's own unit tests
 * (src/replaygain.rs:1298-1365) -- see tests/test_oracle_stage_b.py.
 *
 * Stage A (MP3 bytes -> PCM) lives in mp3_oracle.c; the reference delegates it to the
 * third-party crate symphonia-bundle-mp3 0.5.5 (Cargo.lock:244-247) whose source is not
 * in /root/reference, so that stage restates ISO/IEC 11172-3 Layer III and its
 * parity with symphonia is UNPINNED (cross-checked against FFmpeg mp3float instead).
 */
#ifndef RG_ORACLE_H
#define RG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGO_HIST_SIZE 12000   /* src/replaygain.rs:630 HISTOGRAM_SIZE   */
#define RGO_HIST_OFFSET 2000  /* src/replaygain.rs:635 HISTOGRAM_OFFSET */

/* EqualLoudnessFilter (src/replaygain.rs:533-551): direct-form-I histories, newest first. */
typedef struct {
    double ya[11], yb[11], ba[3], bb[3];
    double yule_x[11], yule_y[11], butter_x[3], butter_y[3];
} rgo_filter;

/* ReplayGainAnalyzer + LoudnessHistogram (src/replaygain.rs:644-771) plus the per-track
 * bits of analyze_track_internal (peak, filters). */
typedef struct {
    rgo_filter f[2];
    double lsum, rsum;
    size_t totsamp, window_samples;
    double peak;
    uint32_t hist[RGO_HIST_SIZE];
    uint32_t n_windows;   /* windows that reached finish_window with totsamp > 0 */
    uint32_t n_dropped;   /* of those, windows whose index fell outside the histogram */
    unsigned sample_rate;
    int channels;
} rgo_state;

typedef struct {
    double loudness_db, gain_db, peak;
    uint32_t sample_rate;
    uint32_t n_windows, n_dropped;
} rgo_result;

/* EqualLoudnessFilter::new (src/replaygain.rs:555-584). Returns 0, or -1 for an unsupported rate. */
int rgo_filter_init(rgo_filter* f, unsigned sample_rate);
/* EqualLoudnessFilter::process (src/replaygain.rs:586-616). */
double rgo_filter_process(rgo_filter* f, double sample);

/* ReplayGainAnalyzer::new + filters (src/replaygain.rs:866-878). channels: 1 or 2. */
int rgo_state_init(rgo_state* s, unsigned sample_rate, int channels);
/* process_audio_buffer, F32 arm (src/replaygain.rs:960-980); planar normalised f32. ch1 ignored for mono. */
void rgo_state_feed_f32(rgo_state* s, const float* ch0, const float* ch1, size_t frames);
/* Feed samples that are ALREADY in 16-bit scale as f64 and bypass peak tracking: the path the
 * reference's unit tests drive (src/replaygain.rs:1311-1316: filter.process + add_mono_sample). */
void rgo_state_feed_mono_f64_raw(rgo_state* s, const double* x, size_t frames);
/* analyzer.finish_window() + get_loudness + PINK_REF (src/replaygain.rs:907-919). */
void rgo_state_finish(rgo_state* s, rgo_result* out);

/* LoudnessHistogram::get_loudness (src/replaygain.rs:665-682) on 32- or 64-bit counters. */
double rgo_hist_loudness_u32(const uint32_t* hist);
double rgo_hist_loudness_u64(const uint64_t* hist);

/* Whole-buffer convenience: init + feed + finish. hist_out may be NULL. Returns 0 / -1 (bad rate). */
int rgo_analyze_pcm_f32(const float* ch0, const float* ch1, int channels, size_t frames,
                        unsigned sample_rate, rgo_result* out, uint32_t* hist_out);

/* analyze_album_with_index merge (src/replaygain.rs:1048-1073): sums n histograms (u32 each, like the
 * reference's u32 accumulate), max of peaks, loudness of the sum. */
void rgo_album_merge(const uint32_t* hists, const double* peaks, size_t n,
                     uint32_t* album_hist, double* album_loudness_db, double* album_gain_db, double* album_peak);

/* ReplayGainResult::gain_steps (src/replaygain.rs:72-74): round-half-away-from-zero of gain/1.5. */
int rgo_gain_steps(double gain_db);

#ifdef __cplusplus
}
#endif
#endif
