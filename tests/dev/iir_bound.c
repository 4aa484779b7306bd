/* tests/dev/iir_bound.c -- TEST INFRASTRUCTURE.  How far can the time-blocked evaluation of the equal-loudness cascade (the
 * arithmetic of K4's k_iir_fast: transposed direct form II with FMAs, chains restarted from zero state `warm` windows early)
 * drift from the reference's serial direct-form-I arithmetic (oracle/rg_oracle.c), in units of val = 1000*log10(mean square),
 * relative to the tolerance with which K5 flags a window for the exact re-run?  tests/test_iir_bound.py asserts a 10x margin
 * over stress signals at the edge of what a Layer III decoder can emit, at every MP3 sample rate.
 *   gcc -O2 -ffp-contract=off -mfma -I oracle -o iir_bound tests/dev/iir_bound.c oracle/rg_oracle.c -lm
 *   iir_bound <sample_rate> <block_windows> <warm_windows>   -> one line per signal: id max|dval| max(|dval|/tol) flagged_windows
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rg_oracle.h"
#include "rg_coeffs_oracle.h"

typedef struct { double a[11], b[11], A[3], B[3], c, c2, s0[10], u0[2]; } tdf_coef;

static void tdf_init(tdf_coef* k, unsigned sr) {
    int r = -1;
    for (int i = 0; i < RG_NUM_RATES; i++) if (RG_RATES[i] == sr) r = i;
    const double* row = RG_COEFS[r];
    for (int i = 0; i < 11; i++) { k->a[i] = row[i]; k->b[i] = row[11 + i] * 32768.0; }
    for (int i = 0; i < 3; i++) { k->A[i] = row[22 + i]; k->B[i] = row[25 + i]; }
    double sa = 1.0, sA = 1.0, sB = 0.0;
    for (int i = 1; i <= 10; i++) sa += k->a[i];
    for (int i = 1; i <= 2; i++) sA += k->A[i];
    for (int i = 0; i <= 2; i++) sB += k->B[i];
    k->c = 1e-10 / sa;
    k->c2 = (1e-10 + sB * k->c) / sA;
    for (int i = 1; i <= 10; i++) { double s = 0; for (int j = i; j <= 10; j++) s += k->a[j]; k->s0[i - 1] = k->c * s; }
    for (int i = 1; i <= 2; i++) { double s = 0; for (int j = i; j <= 2; j++) s += -k->c * k->B[j] + k->c2 * k->A[j]; k->u0[i - 1] = s; }
}
static inline double tdf_step(const tdf_coef* k, double* s, double* u, double x) {
    double y = fma(k->b[0], x, s[0]);
    for (int i = 1; i <= 9; i++) s[i - 1] = fma(-k->a[i], y, fma(k->b[i], x, s[i]));
    s[9] = fma(-k->a[10], y, k->b[10] * x);
    double z = fma(k->B[0], y, u[0]);
    u[0] = fma(-k->A[1], z, fma(k->B[1], y, u[1]));
    u[1] = fma(-k->A[2], z, k->B[2] * y);
    return z;
}
static unsigned lcg;
static double noise(void) { lcg = lcg * 1664525u + 1013904223u; return (((int)((lcg >> 16) & 0xFFFF)) - 32768) / 32768.0; }

int main(int argc, char** argv) {
    const unsigned sr = argc > 1 ? (unsigned)atoi(argv[1]) : 44100;
    const int bw = argc > 2 ? atoi(argv[2]) : 16, warm = argc > 3 ? atoi(argv[3]) : 1;
    const int win = (int)(sr * 50 / 1000), nwin = 40 * bw;
    const size_t n = (size_t)win * nwin;
    float* x = malloc(n * sizeof(float));
    double* ze = malloc(n * sizeof(double));
    double* blk = malloc((size_t)bw * sizeof(double));
    tdf_coef k; tdf_init(&k, sr);
    for (int sig = 0; sig < 10; sig++) {
        lcg = 777 + sig;
        double lp = 0;
        for (size_t i = 0; i < n; i++) {
            const int w = (int)(i / win);
            double v = noise(), amp = 0.25;
            switch (sig) {
                case 0: break;                                                                   /* noise at -12 dBFS */
                case 1: amp = 1e-3; break;                                                       /* quiet noise */
                case 2: lp += 0.01 * (v - lp); v = lp * 8; break;                                /* mostly below 150 Hz */
                case 3: v = sin(2 * M_PI * 20.0 * i / sr); amp = 1.0; break;                     /* 20 Hz tone, full scale */
                case 4: v = 0.5 * sin(2 * M_PI * 1000.0 * i / sr) + 0.3; amp = 1.0; break;       /* tone + DC */
                case 5: v = (w / 7) % 2 ? v : 1e-4 * v; break;                                   /* loud / near-silent alternation */
                case 6: amp = (w % bw == bw - 1) ? 3.0e4 : 1e-3; break;                          /* +90 dBFS burst in the LAST window of a block (the next chain's warm-up) */
                case 7: amp = (w % bw == bw - 2) ? 3.0e4 : 1e-3; break;                          /* ... two windows before a block start (before the warm-up, if warm = 1) */
                case 8: v = (w % bw >= bw - 2) ? 1.0 : -1.0; amp = 3.0e4; break;                 /* DC steps of +-3e4 (|PCM| ~ 1e9) around block starts */
                case 9: v = sin(2 * M_PI * 20.0 * i / sr) * ((w % bw == bw - 1) ? 3.0e4 : 1.0); amp = 1.0; break;   /* 20 Hz, huge then normal */
            }
            x[i] = (float)(amp * v);
        }
        rgo_filter f; rgo_filter_init(&f, sr);
        for (size_t i = 0; i < n; i++) ze[i] = rgo_filter_process(&f, (double)x[i] * 32768.0);
        double worst = 0, worst_ratio = 0;
        int flagged = 0;
        double s[10], u[2];
        float hist[3] = {0, 0, 0};
        for (int w = 0; w < nwin; w++) {
            if (w % bw == 0) {   /* the chain of this block: zero state `warm` windows early (head chain: the track's equivalent initial state) */
                if (w == 0) { memcpy(s, k.s0, sizeof s); memcpy(u, k.u0, sizeof u); }
                else {
                    /* (the kernel carries the output offset c2 inside the Butterworth states; here it is added to the output) */
                    memset(s, 0, sizeof s); memset(u, 0, sizeof u);
                    for (size_t i = (size_t)(w - warm) * win; i < (size_t)w * win; i++) tdf_step(&k, s, u, (double)x[i]);
                }
                for (int ww = w; ww < w + bw && ww < nwin; ww++) {
                    double acc = 0;
                    for (int i = 0; i < win; i++) { double z = tdf_step(&k, s, u, (double)x[(size_t)ww * win + i]) + k.c2; acc = fma(z, z, acc); }
                    blk[ww - w] = acc;
                }
            }
            double e = 0;
            float amax = 0;
            for (int i = 0; i < win; i++) { double z = ze[(size_t)w * win + i]; e = e + z * z; float a = fabsf(x[(size_t)w * win + i]); if (a > amax) amax = a; }
            /* one channel feeding both sums, like add_mono_sample: (l + r) / tot * 0.5 = e / win */
            const double ms_e = (e + e) / win * 0.5, ms_b = (blk[w % bw] + blk[w % bw]) / win * 0.5;
            const double ve = 1000.0 * log10(ms_e + 1e-37), vb = 1000.0 * log10(ms_b + 1e-37);
            float a = amax;                                   /* K5: the loudest of this window and the three before it */
            for (int j = 0; j < 3; j++) if (hist[j] > a) a = hist[j];
            hist[2] = hist[1]; hist[1] = hist[0]; hist[0] = amax;
            const double tol = 1e-9 * fmax(1.0, 0.1 * 32768.0 * (double)a / sqrt(ms_b + 1e-300));
            const int hard_flag = !(a <= 1.0e6f) || !(ms_b - ms_b == 0.0);   /* K5 hands such tracks to the serial kernel outright */
            if (hard_flag) { flagged++; continue; }
            if (vb > -2002 && vb < 10001) {
                const double d = fabs(vb - ve);
                if (d > worst) worst = d;
                if (d / tol > worst_ratio) worst_ratio = d / tol;
            }
        }
        printf("%d %.6e %.6e %d\n", sig, worst, worst_ratio, flagged);
    }
    return 0;
}
