/* tests/dev/tdf2_check.c -- development check (test infrastructure): how far is the transposed-direct-form-II, FMA,
 * time-blocked evaluation of the equal-loudness cascade (the arithmetic of the FAST K4 kernel) from the reference's
 * direct-form-I unfused arithmetic (oracle/rg_oracle.c), measured in units of val = 1000*log10(mean square) per window?
 *   gcc -O2 -ffp-contract=off -mfma -I oracle -o /tmp/tdf2_check tests/dev/tdf2_check.c oracle/rg_oracle.c -lm
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
    k->c = 1e-10 / sa;                         /* steady-state offset of the Yule stage output */
    k->c2 = (1e-10 + sB * k->c) / sA;          /* steady-state offset of the Butterworth output (sB == 0 for a true high-pass) */
    /* states equivalent to the reference's all-zero histories: y'(-k) = -c, x(-k) = 0; butter in(-k) = -c (+c added = 0), z'(-k) = -c2 */
    for (int i = 1; i <= 10; i++) { double s = 0; for (int j = i; j <= 10; j++) s += k->a[j]; k->s0[i - 1] = k->c * s; }
    for (int i = 1; i <= 2; i++) { double s = 0; for (int j = i; j <= 2; j++) s += -k->c * k->B[j] + k->c2 * k->A[j]; k->u0[i - 1] = s; }
}

/* one sample of the linear cascade (no offsets inside): returns z' */
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
    const int win = (int)(sr * 50 / 1000), nwin = 400, bw = 16;
    const size_t n = (size_t)win * nwin;
    float* x = malloc(n * sizeof(float));
    double* ze = malloc(n * sizeof(double));
    tdf_coef k; tdf_init(&k, sr);
    printf("sr %u c %.3e c2 %.3e\n", sr, k.c, k.c2);
    for (int sig = 0; sig < 6; sig++) {
        lcg = 12345 + sig;
        double lp = 0, amp = sig == 1 ? 1e-3 : 0.25;
        for (size_t i = 0; i < n; i++) {
            double v = noise();
            if (sig == 2) { lp += 0.01 * (v - lp); v = lp * 8; }                           /* heavy low-pass: mostly < 150 Hz */
            if (sig == 3) v = sin(2 * M_PI * 40.0 * i / sr);                               /* 40 Hz sine: strongly attenuated */
            if (sig == 4) v = 0.5 * sin(2 * M_PI * 1000.0 * i / sr) + 0.3;                  /* tone + DC */
            if (sig == 5) v = (i / (win * 7)) % 2 ? v : 1e-4 * v;                           /* loud / near-silent alternation */
            x[i] = (float)(amp * v);
        }
        /* exact: the oracle's filter, sample by sample */
        rgo_filter f; rgo_filter_init(&f, sr);
        for (size_t i = 0; i < n; i++) ze[i] = rgo_filter_process(&f, (double)x[i] * 32768.0);
        double worst_c = 0, worst_b = 0, worst_ratio_c = 0, worst_ratio_b = 0;
        /* continuous TDF-II from the track start with equivalent initial state */
        double s[10], u[2];
        memcpy(s, k.s0, sizeof s); memcpy(u, k.u0, sizeof u);
        double* valc = malloc(nwin * sizeof(double));
        for (int w = 0; w < nwin; w++) {
            double acc = 0;
            for (int i = 0; i < win; i++) { double z = tdf_step(&k, s, u, (double)x[(size_t)w * win + i]) + k.c2; acc = fma(z, z, acc); }
            valc[w] = acc;
        }
        for (int w = 0; w < nwin; w++) {
            double e = 0, amax = 0;
            for (int i = 0; i < win; i++) { double z = ze[(size_t)w * win + i]; e = e + z * z; double a = fabs(x[(size_t)w * win + i]); if (a > amax) amax = a; }
            /* blocked: zero state one window early (first block of the track: equivalent initial state, no warm-up) */
            int w0 = w / bw * bw;
            static double blk[64]; static int blk_w0 = -1;
            if (w == w0) {
                blk_w0 = w0;
                if (w0 == 0) { memcpy(s, k.s0, sizeof s); memcpy(u, k.u0, sizeof u); }
                else { memset(s, 0, sizeof s); memset(u, 0, sizeof u); for (int i = 0; i < win; i++) tdf_step(&k, s, u, (double)x[(size_t)(w0 - 1) * win + i]); }
                for (int ww = w0; ww < w0 + bw && ww < nwin; ww++) {
                    double acc = 0;
                    for (int i = 0; i < win; i++) { double z = tdf_step(&k, s, u, (double)x[(size_t)ww * win + i]) + k.c2; acc = fma(z, z, acc); }
                    blk[ww - w0] = acc;
                }
            }
            double ms_e = e / win * 0.5 * 2, ms_c = valc[w] / win, ms_b = blk[w - blk_w0] / win;  /* mono-equivalent: both channels equal */
            double ve = 1000 * log10(ms_e + 1e-37), vc = 1000 * log10(ms_c + 1e-37), vb = 1000 * log10(ms_b + 1e-37);
            double tol = 1e-9 * fmax(1.0, 0.1 * 32768.0 * amax / sqrt(ms_e + 1e-300));
            if (ve > -2002 && ve < 10001) {
                if (fabs(vc - ve) > worst_c) worst_c = fabs(vc - ve);
                if (fabs(vb - ve) > worst_b) worst_b = fabs(vb - ve);
                if (fabs(vc - ve) / tol > worst_ratio_c) worst_ratio_c = fabs(vc - ve) / tol;
                if (fabs(vb - ve) / tol > worst_ratio_b) worst_ratio_b = fabs(vb - ve) / tol;
            }
        }
        printf("signal %d: max|dval| continuous %.3e (%.3f of tol)   blocked %.3e (%.3f of tol)\n", sig, worst_c, worst_ratio_c, worst_b, worst_ratio_b);
        free(valc);
    }
    return 0;
}
