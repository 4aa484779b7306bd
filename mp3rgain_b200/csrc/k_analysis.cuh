// k_analysis.cuh -- stage B on the device: equal-loudness IIR cascade (K4) and
// window -> bin -> histogram -> 95th percentile (K5).
//
// K4 replaces EqualLoudnessFilter::process + the sum-of-squares half of ReplayGainAnalyzer
// (reference src/replaygain.rs:586-616, 720-740, 966-975).  K5 replaces finish_window,
// LoudnessHistogram::get_loudness and the album accumulate (src/replaygain.rs:743-765, 665-682, 1056-1059).
//
// Arithmetic contract (this file is compiled with -fmad=false; every f64 op below is spelled with an
// _rn intrinsic so nothing is contracted or reassociated):
//   EXACT mode: per sample, t_i = b_i*x[n-i] - a_i*y[n-i] (two products, one subtract), S = t_1+...+t_10
//   left to right, y = (1e-10 + b_0*x[n]) + S -- the reference's order.  The *32768 input scale
//   (src/replaygain.rs:969) is folded into b_i: multiplying by 2^15 commutes with rounding, so
//   (2^15*b_i)*x is bit-identical to b_i*(2^15*x).  The leading "0.0 +" of Rust's iterator sum is dropped:
//   it can only change the sign of an all-zero partial sum, which the final add of a non-negative-zero
//   term erases.
//   FAST mode: the cascade as a linear transposed-direct-form-II filter evaluated with FMAs (see iir_step_fast);
//   chains start from zero state n_warm samples early.  Not bit-identical to the reference; K5 flags every
//   window whose bin decision is within `edge_tol` of flipping and the host re-decodes and re-runs those tracks
//   through the EXACT kernel at the end of the call.
#pragma once
#include "rg_internal.h"
#include "k_ptx.cuh"

// Coefficients travel as a by-value kernel parameter (constant bank 0), so every coefficient is a uniform operand at a
// compile-time offset -- no coefficient registers, no shared device symbol between engines.
struct RgCoefExact {      // reference order, direct form I
    double a[10];         // a_1..a_10 (Yule)
    double b[11];         // 32768*b_0..b_10 (Yule; the *32768 input scale of src/replaygain.rs:969 folded in)
    double A[2];          // a_1,a_2 (Butterworth)
    double B[3];          // b_0..b_2 (Butterworth)
};
struct RgCoefFast {       // transposed direct form II of the same cascade
    double a[10], b[11], A[2], B[3];
    double k2;            // c2*(1+A1+A2), c2 = steady-state output offset produced by the two "+1e-10" terms of the reference;
                          // the offset is carried inside the Butterworth states (u0+c2, u1+c2*(1+A1)), so the output needs no add
    double s0[10], u0[2]; // filter states equivalent to the reference's all-zero histories at the start of a track
    double uz[2];         // Butterworth states of a chain that starts from rest: {c2, c2*(1+A1)}
};

struct IirState {
    double xh[10], yh[10];  // Yule histories: xh[i] = x[n-1-i], yh[i] = y[n-1-i]
    double bx[2], by[2];    // Butterworth histories
};

// EXACT: per sample, t_i = b_i*x[n-i] - a_i*y[n-i] (two products, one subtract), S = t_1+...+t_10 left to right,
// y = (1e-10 + b_0*x[n]) + S -- the reference's order (src/replaygain.rs:595-599, 608-612), nothing fused.
__device__ __forceinline__ double iir_step_exact(IirState& s, const RgCoefExact& c, float xf) {
    const double x0 = (double)xf;
    double S = __dsub_rn(__dmul_rn(c.b[1], s.xh[0]), __dmul_rn(c.a[0], s.yh[0]));
#pragma unroll
    for (int i = 1; i < 10; i++) {
        double t = __dsub_rn(__dmul_rn(c.b[1 + i], s.xh[i]), __dmul_rn(c.a[i], s.yh[i]));
        S = __dadd_rn(S, t);
    }
    const double y0 = __dadd_rn(__dadd_rn(1e-10, __dmul_rn(c.b[0], x0)), S);
#pragma unroll
    for (int i = 9; i >= 1; i--) { s.xh[i] = s.xh[i - 1]; s.yh[i] = s.yh[i - 1]; }
    s.xh[0] = x0;
    s.yh[0] = y0;
    const double t1 = __dsub_rn(__dmul_rn(c.B[1], s.bx[0]), __dmul_rn(c.A[0], s.by[0]));
    const double t2 = __dsub_rn(__dmul_rn(c.B[2], s.bx[1]), __dmul_rn(c.A[1], s.by[1]));
    const double z0 = __dadd_rn(__dadd_rn(1e-10, __dmul_rn(c.B[0], y0)), __dadd_rn(t1, t2));
    s.bx[1] = s.bx[0]; s.bx[0] = y0;
    s.by[1] = s.by[0]; s.by[0] = z0;
    return z0;
}

// FAST: the same cascade as a *linear* filter in transposed direct form II, evaluated with FMAs:
//     y = b0*x + s1;   s_k' = (b_k*x + s_{k+1}) - a_k*y   (k = 1..10, s_11 = 0)
// Every state update is an independent 2-FMA chain and only y -> s1' -> y is loop-carried, so a single warp keeps ~10
// DFMAs in flight, and the state is 12 doubles instead of 24.  The reference's two "+1e-10" terms make its filter affine;
// an affine filter equals the linear one plus a constant (c2 at the output, carried in the Butterworth states -- the same
// argument that lets a chain start from zero state one window early) and, at the start of a track, the linear filter
// started from the non-zero state {s0,u0} that corresponds to the reference's all-zero histories.  Not bit-identical
// arithmetic: K5 flags every window whose bin decision is within edge_tol of flipping and those tracks are re-run EXACT.
struct TdfState { double s[10], u[2]; };
__device__ __forceinline__ double iir_step_fast(TdfState& st, const RgCoefFast& c, float xf) {
    const double x = (double)xf;
    double t[10];
#pragma unroll
    for (int k = 1; k <= 9; k++) t[k - 1] = __fma_rn(c.b[k], x, st.s[k]);
    t[9] = __dmul_rn(c.b[10], x);
    const double y = __fma_rn(c.b[0], x, st.s[0]);
#pragma unroll
    for (int k = 1; k <= 10; k++) st.s[k - 1] = __fma_rn(-c.a[k - 1], y, t[k - 1]);
    const double v1 = __fma_rn(c.B[1], y, st.u[1]);
    const double v2 = __fma_rn(c.B[2], y, c.k2);
    const double z = __fma_rn(c.B[0], y, st.u[0]);   // already includes the affine offset c2
    st.u[0] = __fma_rn(-c.A[0], z, v1);
    st.u[1] = __fma_rn(-c.A[1], z, v2);
    return z;
}

// Per-chain window accounting.
struct WinAcc {
    double sum;
    double* out;
    float* mout;
    float wmax;
    uint32_t tot, win;
};

template <bool FAST, bool ACCUM, class State, class Coef>
__device__ __forceinline__ void iir_sample(State& s, const Coef& c, WinAcc& w, float& peak, float xf) {
    double z;
    if constexpr (FAST) z = iir_step_fast(s, c, xf);
    else z = iir_step_exact(s, c, xf);
    if (ACCUM) {
        w.wmax = fmaxf(w.wmax, fabsf(xf));
        if (FAST) w.sum = __fma_rn(z, z, w.sum);
        else w.sum = __dadd_rn(w.sum, __dmul_rn(z, z));
        if (++w.tot == w.win) {
            *w.out++ = w.sum;
            *w.mout++ = w.wmax;
            peak = fmaxf(peak, w.wmax);
            w.sum = 0.0;
            w.wmax = 0.0f;
            w.tot = 0;
        }
    }
}

// Feed n consecutive samples starting at p: scalar head up to 16-byte alignment, float4 body (NV float4 per iteration,
// the next iteration's loads issued before this iteration's arithmetic), scalar tail.
template <bool FAST, bool ACCUM, int NV, class State, class Coef>
__device__ __forceinline__ void iir_run(State& s, const Coef& c, WinAcc& w, float& peak, const float* __restrict__ p, uint32_t n) {
    while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 15u)) {
        iir_sample<FAST, ACCUM>(s, c, w, peak, __ldg(p));
        p++; n--;
    }
    const float4* p4 = reinterpret_cast<const float4*>(p);
    uint32_t nit = n / (4 * NV);
    if (nit) {
        float4 buf[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) buf[k] = __ldg(p4 + k);
        for (uint32_t it = 0; it < nit; it++) {
            float4 cur[NV];
#pragma unroll
            for (int k = 0; k < NV; k++) cur[k] = buf[k];
            p4 += NV;
            if (it + 1 < nit) {
#pragma unroll
                for (int k = 0; k < NV; k++) buf[k] = __ldg(p4 + k);
            }
            if (ACCUM && w.tot + 4 * NV < w.win) {
                // the window cannot end inside this group: no per-sample bookkeeping, so the FP64 work of consecutive samples
                // overlaps freely
                float m = w.wmax;
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    const float xs[4] = {cur[k].x, cur[k].y, cur[k].z, cur[k].w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        double z;
                        if constexpr (FAST) z = iir_step_fast(s, c, xs[q]);
                        else z = iir_step_exact(s, c, xs[q]);
                        m = fmaxf(m, fabsf(xs[q]));
                        if (FAST) w.sum = __fma_rn(z, z, w.sum);
                        else w.sum = __dadd_rn(w.sum, __dmul_rn(z, z));
                    }
                }
                w.wmax = m;
                w.tot += 4 * NV;
            } else {
#pragma unroll
                for (int k = 0; k < NV; k++) {
                    iir_sample<FAST, ACCUM>(s, c, w, peak, cur[k].x);
                    iir_sample<FAST, ACCUM>(s, c, w, peak, cur[k].y);
                    iir_sample<FAST, ACCUM>(s, c, w, peak, cur[k].z);
                    iir_sample<FAST, ACCUM>(s, c, w, peak, cur[k].w);
                }
            }
        }
    }
    p = reinterpret_cast<const float*>(p4);
    n -= nit * (4 * NV);
    while (n > 0) {
        iir_sample<FAST, ACCUM>(s, c, w, peak, __ldg(p));
        p++; n--;
    }
}

// K4: one thread per chain; all chains of one launch share a sample rate.
#define K4_EXACT_THREADS 32
__global__ void __launch_bounds__(K4_EXACT_THREADS)
k_iir_exact(const RgChain* __restrict__ chains, uint32_t n_chains, float* __restrict__ peaks, uint32_t chain_base,
            const __grid_constant__ RgCoefExact c) {
    uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_chains) return;
    const RgChain ch = chains[ci];
    IirState s;
#pragma unroll
    for (int i = 0; i < 10; i++) { s.xh[i] = 0.0; s.yh[i] = 0.0; }
    s.bx[0] = s.bx[1] = s.by[0] = s.by[1] = 0.0;
    WinAcc w;
    w.sum = 0.0; w.out = ch.wsum; w.mout = ch.wmax; w.wmax = 0.0f; w.tot = 0; w.win = ch.win;
    float peak = 0.0f;
    iir_run<false, true, 5>(s, c, w, peak, ch.src, ch.n_main);
    if (w.tot) {  // the track's final partial window (src/replaygain.rs:907)
        *w.out = w.sum;
        *w.mout = w.wmax;
        peak = fmaxf(peak, w.wmax);
    }
    peaks[chain_base + ci] = peak;
}

template <int NV, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_iir_fast_t(const RgChain* __restrict__ chains, uint32_t n_chains, float* __restrict__ peaks, uint32_t chain_base,
             const __grid_constant__ RgCoefFast c) {
    uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_chains) return;
    const RgChain ch = chains[ci];
    TdfState s;
    const bool head = ch.flags & 1u;   // chain starts at the first sample of the track
#pragma unroll
    for (int i = 0; i < 10; i++) s.s[i] = head ? c.s0[i] : 0.0;
    s.u[0] = head ? c.u0[0] : c.uz[0];
    s.u[1] = head ? c.u0[1] : c.uz[1];
    WinAcc w;
    w.sum = 0.0; w.out = ch.wsum; w.mout = ch.wmax; w.wmax = 0.0f; w.tot = 0; w.win = ch.win;
    float peak = 0.0f;
    if (ch.n_warm) iir_run<true, false, NV>(s, c, w, peak, ch.src, ch.n_warm);
    iir_run<true, true, NV>(s, c, w, peak, ch.src + ch.n_warm, ch.n_main);
    if (w.tot) {
        *w.out = w.sum;
        *w.mout = w.wmax;
        peak = fmaxf(peak, w.wmax);
    }
    peaks[chain_base + ci] = peak;
}

// The time-blocked kernel of the END-TO-END path: the same chains and the same arithmetic, but the samples reach the thread
// through a private ring of two 128-byte lines in shared memory filled by cp.async (eight 16-byte copies per line, issued back
// to back, one line ahead) instead of float4 loads that rely on the line surviving in L1 from one load to the next.  While file
// images are still arriving, the frame-index kernels of the later groups run beside K4 and stream those images through the same
// L1: the register-prefetch kernel then took 6.1 ms per batch instead of 3.6, this one 4.8 (profiles/r2_timeline_e2e.log:
// 112.5 -> 106.0 ms per 1000 tracks end to end).  Alone it is 6 % slower than k_iir_fast (28.6 vs 26.9 ms: four samples per
// loop trip instead of twelve), so resident corpora keep that one.
#define K4_RING_BYTES (2 * 8 * 256 * 16)
template <bool ACCUM>
__device__ __forceinline__ void iir_run_staged(TdfState& s, const RgCoefFast& c, WinAcc& w, float& peak, const float* __restrict__ p,
                                               uint32_t n, float4* __restrict__ ring /* + threadIdx.x */) {
    // head: up to the next 128-byte line
    uint32_t head = (uint32_t)((128u - (reinterpret_cast<uintptr_t>(p) & 127u)) & 127u) >> 2;
    if (head > n) head = n;
#pragma unroll 1
    for (uint32_t i = 0; i < head; i++) iir_sample<true, ACCUM>(s, c, w, peak, __ldg(p + i));
    p += head; n -= head;
    const uint32_t nlines = n >> 5;
    if (nlines) {
        const float4* g = reinterpret_cast<const float4*>(p);
        auto issue = [&](uint32_t line) {
            float4* dst = ring + (line & 1u) * (8 * 256);
            const float4* src = g + (size_t)line * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) k2_cp_async16(dst + j * 256, src + j);
        };
        issue(0);
        k2_cp_async_commit();
        if (nlines > 1) issue(1);
        k2_cp_async_commit();
#pragma unroll 1
        for (uint32_t l = 0; l < nlines; l++) {
            k2_cp_async_wait1();   // every group but the newest has landed: line l is in its slot
            const float4* cur4 = ring + (l & 1u) * (8 * 256);
#pragma unroll 1   // (measured: unrolling this loop by 2 or 4 makes the kernel 17 % slower)
            for (int j = 0; j < 8; j++) {
                const float4 a = cur4[j * 256];
                const float xs[4] = {a.x, a.y, a.z, a.w};
                if (!ACCUM || w.tot + 4 < w.win) {
                    // the window cannot end inside this group: no per-sample bookkeeping
                    float m = w.wmax;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const double z = iir_step_fast(s, c, xs[q]);
                        if (ACCUM) { m = fmaxf(m, fabsf(xs[q])); w.sum = __fma_rn(z, z, w.sum); }
                    }
                    if (ACCUM) { w.wmax = m; w.tot += 4; }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) iir_sample<true, ACCUM>(s, c, w, peak, xs[q]);
                }
            }
            if (l + 2 < nlines) issue(l + 2);   // into the slot that was just consumed
            k2_cp_async_commit();
        }
        k2_cp_async_wait0();
    }
    p += (size_t)nlines * 32; n -= nlines * 32;
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) iir_sample<true, ACCUM>(s, c, w, peak, __ldg(p + i));
}

__global__ void __launch_bounds__(256, 1)
k_iir_staged(const RgChain* __restrict__ chains, uint32_t n_chains, float* __restrict__ peaks, uint32_t chain_base,
             const __grid_constant__ RgCoefFast c) {
    extern __shared__ __align__(16) unsigned char k4_raw[];
    float4* ring = reinterpret_cast<float4*>(k4_raw) + threadIdx.x;
    uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_chains) return;
    const RgChain ch = chains[ci];
    TdfState s;
    const bool head = ch.flags & 1u;   // chain starts at the first sample of the track
#pragma unroll
    for (int i = 0; i < 10; i++) s.s[i] = head ? c.s0[i] : 0.0;
    s.u[0] = head ? c.u0[0] : c.uz[0];
    s.u[1] = head ? c.u0[1] : c.uz[1];
    WinAcc w;
    w.sum = 0.0; w.out = ch.wsum; w.mout = ch.wmax; w.wmax = 0.0f; w.tot = 0; w.win = ch.win;
    float peak = 0.0f;
    if (ch.n_warm) iir_run_staged<false>(s, c, w, peak, ch.src, ch.n_warm, ring);
    iir_run_staged<true>(s, c, w, peak, ch.src + ch.n_warm, ch.n_main, ring);
    if (w.tot) {
        *w.out = w.sum;
        *w.mout = w.wmax;
        peak = fmaxf(peak, w.wmax);
    }
    peaks[chain_base + ci] = peak;
}

// the register-prefetch instantiations (tools/iir_bench.cu times the alternatives; `k4_slim` launches the second)
#define K4_FAST_THREADS 256
#define k_iir_fast k_iir_fast_t<3, K4_FAST_THREADS, 2>
// the "slim" launch: 128-thread CTAs, launched one wave of n_sms CTAs at a time -- one CTA per SM keeps most of the FP64 pipe
// busy with a quarter of the register file, so that the decode kernels of the next batch stay resident beside it
#define K4_SLIM_THREADS 128
#define k_iir_slim k_iir_fast_t<3, K4_SLIM_THREADS, 4>

// ------------------------------------------------------------------------------------------------------------
// K5: one CTA per track.  Window sums -> bin index -> shared 12000-bin histogram (warp-aggregated atomics)
// -> suffix scan for the 95th percentile -> optional flush to the global per-track / album histograms.

__device__ __forceinline__ void hist_add_aggregated(uint32_t* hist, int idx, bool valid) {
    // lanes of a warp that hit the same bin elect one leader which adds the group's population
    unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    unsigned peers = __match_any_sync(active, idx);
    int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[idx], (uint32_t)__popc(peers));
}

#define K5_THREADS 256
__global__ void __launch_bounds__(K5_THREADS)
k_hist_percentile(const RgTrackWin* __restrict__ tracks, const float* __restrict__ chain_peaks,
                  RgTrackOut* __restrict__ out, uint32_t* __restrict__ hist_out /* n*12000 or null */,
                  unsigned long long* __restrict__ album_hist /* 12000 or null */,
                  unsigned long long* __restrict__ album_peak_bits /* or null */, double edge_tol,
                  int defer_album_on_edge /* skip the album adds when this track has edge windows (it will be re-run) */,
                  int scale_tol /* time-blocked pass: scale edge_tol by the window's input-amplitude / filtered-RMS ratio */) {
    __shared__ uint32_t hist[RG_HIST_BINS];
    __shared__ uint32_t s_cnt[3];           // counted, dropped, edge
    __shared__ uint32_t s_chunk[K5_THREADS];   // per-track window counts fit 32 bits
    __shared__ float s_peak[K5_THREADS / 32];
    __shared__ int s_found;

    const RgTrackWin t = tracks[blockIdx.x];
    const int tid = threadIdx.x;
    for (int i = tid; i < RG_HIST_BINS; i += K5_THREADS) hist[i] = 0;
    if (tid < 3) s_cnt[tid] = 0;
    if (tid == 0) s_found = -1;
    __syncthreads();

    const uint32_t n_win = t.n_full + (t.rem ? 1u : 0u);
    uint32_t dropped = 0, edge = 0, counted = 0;
    const uint32_t n_round = (n_win + K5_THREADS - 1) / K5_THREADS * K5_THREADS;
    for (uint32_t w = tid; w < n_round; w += K5_THREADS) {
        bool valid = false;
        int idx = 0;
        if (w < n_win) {
            const double tot = (double)(w < t.n_full ? t.win : t.rem);
            const double l = t.wsum0[w], r = t.wsum1[w];
            // (lsum + rsum) / totsamp * 0.5 ; 1000*log10(ms + 1e-37) ; `as i32` (saturating, NaN -> 0)
            const double ms = __dmul_rn(__ddiv_rn(__dadd_rn(l, r), tot), 0.5);
            const double val = __dmul_rn(1000.0, log10(__dadd_rn(ms, 1e-37)));
            // Rust `val as i32` (src/replaygain.rs:755): truncates, saturates, NaN -> 0.  Spelled out: cvt.rzi.s32.f64 maps a
            // NaN *f64* source to MININT (PTX ISA, cvt: "zero if the source is not .f64 ..., otherwise 1 << (bits-1)"), which
            // round 1's hardware run showed as windows missing from bin 2000.
            const int v = (val != val) ? 0 : (val >= 2147483647.0) ? 2147483647 : (val <= -2147483648.0) ? (-2147483647 - 1) : (int)val;
            const long long li = (long long)v + RG_HIST_OFF;
            // Edge test: would a perturbation of the window's mean square at the level of the time-blocked kernel's
            // rounding differences move trunc(val) across a bin boundary that matters?  The tolerance grows with the
            // ratio of input amplitude to filtered RMS (heavily attenuated windows carry larger relative error), and
            // windows far outside the histogram can never change bin membership.
            // The filter state entering a window carries the level of the windows before it, and a chain that restarts from
            // zero state one (two, below 32 kHz) windows early forgets what came before its warm-up only down to
            // pole^(warm-up samples) of that level: take the loudest of this window and the three before it
            // (tests/test_iir_bound.py: a +90 dBFS burst two windows before a block start is otherwise 150x over tolerance).
            float a = 0.0f;
            if (scale_tol) {
                a = fmaxf(t.wmax0[w], t.wmax1[w]);
#pragma unroll
                for (uint32_t j = 1; j <= 3; j++)
                    if (w >= j) a = fmaxf(a, fmaxf(t.wmax0[w - j], t.wmax1[w - j]));
                // Outside what one warm-up window can forget: samples beyond +120 dBFS (no Layer III stream decodes to them
                // short of global_gain 255 throughout; caller-supplied PCM can hold anything), infinities, and NaNs, which the
                // serial filter carries to the end of the track.  Flag the track: it is filtered again serially.
                if (!(a <= 1.0e6f) || !(ms - ms == 0.0)) edge++;
            }
            if (val > -2002.0 && val < 10001.0) {
                double tol = edge_tol;
                if (scale_tol) {
                    const double amp = 32768.0 * (double)a;
                    const double ratio = amp / sqrt(ms + 1e-300);
                    tol = edge_tol * fmax(1.0, 0.1 * ratio);
                }
                if (fabs(val - rint(val)) < tol) edge++;
            }
            if (li >= 0 && li < RG_HIST_BINS) { valid = true; idx = (int)li; counted++; }
            else dropped++;
        }
        hist_add_aggregated(hist, idx, valid);
    }
    // peak over this track's chains
    float pk = 0.0f;
    for (uint32_t c = t.chain_begin + tid; c < t.chain_end; c += K5_THREADS) pk = fmaxf(pk, chain_peaks[c]);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
        edge += __shfl_xor_sync(0xffffffffu, edge, o);
        counted += __shfl_xor_sync(0xffffffffu, counted, o);
    }
    if ((tid & 31) == 0) {
        s_peak[tid >> 5] = pk;
        atomicAdd(&s_cnt[0], counted);
        atomicAdd(&s_cnt[1], dropped);
        atomicAdd(&s_cnt[2], edge);
    }
    __syncthreads();

    // get_loudness: threshold = ceil(total * (1.0 - 0.95)); scan from the top bin down.
    const unsigned long long total = s_cnt[0];
    constexpr int CH = (RG_HIST_BINS + K5_THREADS - 1) / K5_THREADS;  // bins per thread, thread 0 owns the TOP chunk
    const int hi = RG_HIST_BINS - 1 - tid * CH;                       // my highest bin
    unsigned long long mine = 0;
    for (int k = 0; k < CH; k++) {
        int b = hi - k;
        if (b >= 0) mine += hist[b];
    }
    s_chunk[tid] = (uint32_t)mine;
    __syncthreads();
    if (total > 0) {
        const unsigned long long threshold = (unsigned long long)ceil((double)total * (1.0 - 0.95));
        unsigned long long before = 0;  // counts in chunks above mine
        for (int k = 0; k < tid; k++) before += s_chunk[k];
        if (before < threshold && before + mine >= threshold) {
            unsigned long long cnt = before;
            for (int k = 0; k < CH; k++) {
                int b = hi - k;
                if (b < 0) break;
                cnt += hist[b];
                if (cnt >= threshold) { s_found = b; break; }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        float p = 0.0f;
        for (int k = 0; k < K5_THREADS / 32; k++) p = fmaxf(p, s_peak[k]);
        RgTrackOut o;
        o.loudness_db = (total == 0 || s_found < 0) ? -20.0 : (double)(s_found - RG_HIST_OFF) / 100.0;
        o.gain_db = __dsub_rn(64.82, o.loudness_db);
        o.peak = (double)p;
        o.n_windows = n_win;
        o.n_dropped = s_cnt[1];
        o.n_edge = s_cnt[2];
        o.pad = 0;
        out[blockIdx.x] = o;
        if (album_peak_bits && (t.flags & 1u) && !(defer_album_on_edge && s_cnt[2] > 0)) atomicMax(album_peak_bits, (unsigned long long)__double_as_longlong((double)p));
    }
    if (hist_out) {
        uint32_t* dst = hist_out + (size_t)blockIdx.x * RG_HIST_BINS;
        for (int i = tid; i < RG_HIST_BINS; i += K5_THREADS) dst[i] = hist[i];
    }
    if (album_hist && (t.flags & 1u) && !(defer_album_on_edge && s_cnt[2] > 0)) {
        for (int i = tid; i < RG_HIST_BINS; i += K5_THREADS) {
            uint32_t v = hist[i];
            if (v) atomicAdd(&album_hist[i], (unsigned long long)v);
        }
    }
}

__global__ void k_scatter_f32(const float* __restrict__ in, const uint32_t* __restrict__ pos, float* __restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[pos[i]] = in[i];
}

// get_loudness on a merged u64 album histogram (src/replaygain.rs:1065-1066); single CTA.
__global__ void __launch_bounds__(K5_THREADS)
k_album_percentile(const unsigned long long* __restrict__ album_hist, const unsigned long long* __restrict__ album_peak_bits,
                   double* __restrict__ out3 /* loudness, gain, peak */) {
    __shared__ unsigned long long s_chunk[K5_THREADS];
    __shared__ int s_found;
    const int tid = threadIdx.x;
    constexpr int CH = (RG_HIST_BINS + K5_THREADS - 1) / K5_THREADS;
    const int hi = RG_HIST_BINS - 1 - tid * CH;
    unsigned long long mine = 0;
    for (int k = 0; k < CH; k++) {
        int b = hi - k;
        if (b >= 0) mine += album_hist[b];
    }
    s_chunk[tid] = mine;
    if (tid == 0) s_found = -1;
    __syncthreads();
    unsigned long long total = 0;
    for (int k = 0; k < K5_THREADS; k++) total += s_chunk[k];
    if (total > 0) {
        const unsigned long long threshold = (unsigned long long)ceil((double)total * (1.0 - 0.95));
        unsigned long long before = 0;
        for (int k = 0; k < tid; k++) before += s_chunk[k];
        if (before < threshold && before + mine >= threshold) {
            unsigned long long cnt = before;
            for (int k = 0; k < CH; k++) {
                int b = hi - k;
                if (b < 0) break;
                cnt += album_hist[b];
                if (cnt >= threshold) { s_found = b; break; }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        double l = (total == 0 || s_found < 0) ? -20.0 : (double)(s_found - RG_HIST_OFF) / 100.0;
        out3[0] = l;
        out3[1] = __dsub_rn(64.82, l);
        out3[2] = __longlong_as_double((long long)*album_peak_bits);
    }
}
