/*
 * rg_oracle.c -- CPU ORACLE, stage B (test infrastructure, NOT product code).
 * See rg_oracle.h.  Compile with -O2 -ffp-contract=off -fno-fast-math: every multiply and
 * add below must round separately, exactly like the reference's (uncontracted) Rust f64 code.
 */
#include "rg_oracle.h"
#include "rg_coeffs_oracle.h"

#include <math.h>
#include <string.h>

#define DENORMAL_PREVENTION 1e-10 /* src/replaygain.rs:530 */
#define PINK_REF 64.82            /* src/replaygain.rs:44  */
#define SAMPLE_SCALE_16BIT 32768.0 /* src/replaygain.rs:949 */

int rgo_filter_init(rgo_filter* f, unsigned sample_rate) {
    /* src/replaygain.rs:558-572: match on the 12 supported rates, None otherwise */
    for (int r = 0; r < RG_NUM_RATES; r++) {
        if (RG_RATES[r] == sample_rate) {
            memset(f, 0, sizeof(*f));
            memcpy(f->ya, &RG_COEFS[r][0], sizeof(f->ya));
            memcpy(f->yb, &RG_COEFS[r][11], sizeof(f->yb));
            memcpy(f->ba, &RG_COEFS[r][22], sizeof(f->ba));
            memcpy(f->bb, &RG_COEFS[r][25], sizeof(f->bb));
            return 0;
        }
    }
    return -1;
}

double rgo_filter_process(rgo_filter* f, double sample) {
    /* src/replaygain.rs:588-590: shift histories by one, newest at [0] */
    memmove(&f->yule_x[1], &f->yule_x[0], 10 * sizeof(double));
    memmove(&f->yule_y[1], &f->yule_y[0], 10 * sizeof(double));
    f->yule_x[0] = sample;

    /* src/replaygain.rs:595-599:  (1e-10 + b0*x0) + sum_{i=1..10}(b_i*x_i - a_i*y_i),
     * the sum folded left to right from 0.0 */
    double s = 0.0;
    for (int i = 1; i < 11; i++) {
        double t = f->yb[i] * f->yule_x[i] - f->ya[i] * f->yule_y[i];
        s = s + t;
    }
    double yule_out = (DENORMAL_PREVENTION + f->yb[0] * f->yule_x[0]) + s;
    f->yule_y[0] = yule_out;

    /* src/replaygain.rs:603-605 */
    memmove(&f->butter_x[1], &f->butter_x[0], 2 * sizeof(double));
    memmove(&f->butter_y[1], &f->butter_y[0], 2 * sizeof(double));
    f->butter_x[0] = yule_out;

    /* src/replaygain.rs:608-612 */
    double sb = 0.0;
    for (int i = 1; i < 3; i++) {
        double t = f->bb[i] * f->butter_x[i] - f->ba[i] * f->butter_y[i];
        sb = sb + t;
    }
    double butter_out = (DENORMAL_PREVENTION + f->bb[0] * f->butter_x[0]) + sb;
    f->butter_y[0] = butter_out;
    return butter_out;
}

/* Rust `f64 as i32`: saturating, NaN -> 0 (the cast at src/replaygain.rs:755). */
static int32_t rust_f64_as_i32(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v; /* truncation toward zero */
}

/* ReplayGainAnalyzer::finish_window (src/replaygain.rs:743-765) */
static void finish_window(rgo_state* s) {
    if (s->totsamp == 0) return;
    double mean_square = (s->lsum + s->rsum) / (double)s->totsamp * 0.5;
    double val = (100.0 * 10.0) * log10(mean_square + 1e-37);
    /* (val as i32 + 2000) as usize: a negative sum wraps to a huge usize => not < 12000 */
    int64_t idx = (int64_t)rust_f64_as_i32(val) + RGO_HIST_OFFSET;
    s->n_windows++;
    if (idx >= 0 && idx < RGO_HIST_SIZE) s->hist[idx] += 1;
    else s->n_dropped++;
    s->lsum = 0.0;
    s->rsum = 0.0;
    s->totsamp = 0;
}

int rgo_state_init(rgo_state* s, unsigned sample_rate, int channels) {
    memset(s, 0, sizeof(*s));
    if (rgo_filter_init(&s->f[0], sample_rate) != 0) return -1;
    rgo_filter_init(&s->f[1], sample_rate);
    s->window_samples = ((size_t)sample_rate * 50) / 1000; /* src/replaygain.rs:704 */
    s->sample_rate = sample_rate;
    s->channels = channels >= 2 ? 2 : 1;
    return 0;
}

void rgo_state_feed_f32(rgo_state* s, const float* ch0, const float* ch1, size_t frames) {
    for (size_t i = 0; i < frames; i++) {
        double l = (double)ch0[i];
        if (fabs(l) > s->peak) s->peak = fabs(l);               /* peak.max(|l|), :967 */
        double lf = rgo_filter_process(&s->f[0], l * SAMPLE_SCALE_16BIT);
        if (s->channels >= 2) {
            double r = (double)ch1[i];
            if (fabs(r) > s->peak) s->peak = fabs(r);
            double rf = rgo_filter_process(&s->f[1], r * SAMPLE_SCALE_16BIT);
            s->lsum += lf * lf;                                 /* add_sample, :720-728 */
            s->rsum += rf * rf;
        } else {
            double sq = lf * lf;                                /* add_mono_sample, :731-740 */
            s->lsum += sq;
            s->rsum += sq;
        }
        s->totsamp += 1;
        if (s->totsamp >= s->window_samples) finish_window(s);
    }
}

void rgo_state_feed_mono_f64_raw(rgo_state* s, const double* x, size_t frames) {
    for (size_t i = 0; i < frames; i++) {
        double lf = rgo_filter_process(&s->f[0], x[i]);
        double sq = lf * lf;
        s->lsum += sq;
        s->rsum += sq;
        s->totsamp += 1;
        if (s->totsamp >= s->window_samples) finish_window(s);
    }
}

#define HIST_LOUDNESS_BODY                                                            \
    uint64_t total = 0;                                                               \
    for (int i = 0; i < RGO_HIST_SIZE; i++) total += (uint64_t)hist[i];               \
    if (total == 0) return -20.0;                                                     \
    /* ceil(total * (1.0 - 0.95)) with the f64 product, src/replaygain.rs:671 */      \
    uint64_t threshold = (uint64_t)ceil((double)total * (1.0 - 0.95));                \
    uint64_t count = 0;                                                               \
    for (int i = RGO_HIST_SIZE - 1; i >= 0; i--) {                                    \
        count += (uint64_t)hist[i];                                                   \
        if (count >= threshold) return (double)(i - RGO_HIST_OFFSET) / 100.0;         \
    }                                                                                 \
    return -20.0;

double rgo_hist_loudness_u32(const uint32_t* hist) { HIST_LOUDNESS_BODY }
double rgo_hist_loudness_u64(const uint64_t* hist) { HIST_LOUDNESS_BODY }

void rgo_state_finish(rgo_state* s, rgo_result* out) {
    finish_window(s); /* flush the partial window, src/replaygain.rs:907 */
    out->loudness_db = rgo_hist_loudness_u32(s->hist);
    out->gain_db = PINK_REF - out->loudness_db;
    out->peak = s->peak;
    out->sample_rate = s->sample_rate;
    out->n_windows = s->n_windows;
    out->n_dropped = s->n_dropped;
}

int rgo_analyze_pcm_f32(const float* ch0, const float* ch1, int channels, size_t frames,
                        unsigned sample_rate, rgo_result* out, uint32_t* hist_out) {
    static _Thread_local rgo_state st;
    if (rgo_state_init(&st, sample_rate, channels) != 0) return -1;
    rgo_state_feed_f32(&st, ch0, ch1, frames);
    rgo_state_finish(&st, out);
    if (hist_out) memcpy(hist_out, st.hist, sizeof(st.hist));
    return 0;
}

void rgo_album_merge(const uint32_t* hists, const double* peaks, size_t n,
                     uint32_t* album_hist, double* album_loudness_db, double* album_gain_db, double* album_peak) {
    double pk = 0.0;
    memset(album_hist, 0, RGO_HIST_SIZE * sizeof(uint32_t));
    for (size_t t = 0; t < n; t++) {
        if (peaks[t] > pk) pk = peaks[t];
        for (int i = 0; i < RGO_HIST_SIZE; i++) album_hist[i] += hists[t * RGO_HIST_SIZE + i];
    }
    *album_loudness_db = rgo_hist_loudness_u32(album_hist);
    *album_gain_db = PINK_REF - *album_loudness_db;
    *album_peak = pk;
}

int rgo_gain_steps(double gain_db) {
    return (int)round(gain_db / 1.5); /* C round() == Rust f64::round: half away from zero */
}
