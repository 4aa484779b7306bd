// k_hybrid_synth.cuh -- K2: requantise -> MS stereo -> reorder -> alias reduction -> IMDCT/overlap -> polyphase -> PCM.
//
// One CTA owns a run of consecutive granules of one track (both channels) and walks it in time order, keeping the IMDCT
// overlap and the last 15 DCT-32 vectors of the polyphase filterbank in shared memory.  Both are finite-memory, so a run
// that does not start at the head of the track first replays the two granules before it (PCM writes suppressed) to
// rebuild that state exactly: runs are independent and the whole corpus decodes in parallel.
//
// Arithmetic follows DESIGN.md "f32 decode contract" operation for operation.
#pragma once

#define K2_THREADS 256
#define SB_STRIDE 33     // padded row of 32 subband values
#define U_STRIDE 37      // padded row of 36 IMDCT outputs

struct GranInfo {
    int global_gain, sfc, bt, mixed, preflag, sf_shift /* 2 or 4 */, sbg[3];
    uint32_t part2_bit;   // absolute bit position of part 2 in the main-data stream
    uint32_t md_lim;      // bytes readable (later bytes read as zero)
    uint32_t nz;          // lines decoded by K1 (0 = silent)
    int scfsi;            // 4 bits, group 0 = bit 3 (only meaningful for gr 1)
};

template <int N, int OFF>
struct Dct2 {
    static __device__ __forceinline__ void run(float (&x)[N]) {
        float g[N / 2], h[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            g[i] = __fadd_rn(x[i], x[N - 1 - i]);
            h[i] = __fmul_rn(__fsub_rn(x[i], x[N - 1 - i]), c_dct_sec[OFF + i]);
        }
        Dct2<N / 2, OFF + N / 2>::run(g);
        Dct2<N / 2, OFF + N / 2>::run(h);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            x[2 * k] = g[k];
            x[2 * k + 1] = (k + 1 < N / 2) ? __fadd_rn(h[k], h[k + 1]) : h[k];
        }
    }
};
template <int OFF>
struct Dct2<1, OFF> {
    static __device__ __forceinline__ void run(float (&)[1]) {}
};

__device__ __forceinline__ uint32_t md_bits(const uint8_t* __restrict__ mdp, uint32_t lim, uint32_t bit, int n) {
    // n <= 8 bits at absolute bit position `bit`; bytes at or beyond lim read as zero
    if (n == 0) return 0;
    const uint32_t by = bit >> 3;
    const uint32_t b0 = by < lim ? __ldg(mdp + by) : 0u, b1 = by + 1 < lim ? __ldg(mdp + by + 1) : 0u;
    const uint32_t w = (b0 << 8) | b1;
    return (w >> (16 - n - (bit & 7))) & ((1u << n) - 1);
}

__global__ void __launch_bounds__(K2_THREADS)
k_hybrid_synth(const uint8_t* __restrict__ corpus, const uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
               const DecTrack* __restrict__ tracks, const RunDesc* __restrict__ runs, const L3Dev* __restrict__ T,
               const int16_t* __restrict__ spectra, const uint16_t* __restrict__ nz_in, float* __restrict__ pcm, uint32_t flags) {
    __shared__ float s_xr[2][576];
    __shared__ float s_u[2][32][U_STRIDE];
    __shared__ float s_ovl[2][18][32];
    __shared__ float s_sb[2][18][SB_STRIDE];
    __shared__ float s_ring[2][33][SB_STRIDE];
    __shared__ float s_dct4[18][19];
    __shared__ float s_win[4][36];
    __shared__ float s_im12[12][6];
    __shared__ float s_wa[8][32], s_wb[8][32];
    __shared__ float s_gain[L3_GAIN_MAX - L3_GAIN_MIN + 1];
    __shared__ float s_cs[8], s_ca[8];
    __shared__ uint8_t s_idx36[36], s_ia[32], s_ib[32];
    __shared__ int s_sfl[2][22];
    __shared__ int s_sfs[2][39];
    __shared__ GranInfo s_gi[2];

    const int tid = threadIdx.x;
    const RunDesc run = runs[blockIdx.x];
    const DecTrack tr = tracks[run.track];
    const int nch = (int)tr.channels;
    const int sr = (int)tr.sr_idx;
    const uint8_t* mdp = md + tr.md_off;

    // ---- one-time table staging and state reset
    for (int i = tid; i < 18 * 18; i += K2_THREADS) s_dct4[i / 18][i % 18] = T->dct4_18[i / 18][i % 18];
    for (int i = tid; i < 4 * 36; i += K2_THREADS) s_win[i / 36][i % 36] = T->winsgn[i / 36][i % 36];
    for (int i = tid; i < 72; i += K2_THREADS) s_im12[i / 6][i % 6] = T->imdct12[i / 6][i % 6];
    for (int i = tid; i < 256; i += K2_THREADS) { s_wa[i / 32][i % 32] = T->synth_wa[i / 32][i % 32]; s_wb[i / 32][i % 32] = T->synth_wb[i / 32][i % 32]; }
    for (int i = tid; i < L3_GAIN_MAX - L3_GAIN_MIN + 1; i += K2_THREADS) s_gain[i] = T->gain[i];
    if (tid < 8) { s_cs[tid] = T->alias_cs[tid]; s_ca[tid] = T->alias_ca[tid]; }
    if (tid < 36) s_idx36[tid] = T->imdct36_idx[tid];
    if (tid < 32) { s_ia[tid] = T->synth_ia[tid]; s_ib[tid] = T->synth_ib[tid]; }
    for (int i = tid; i < 2 * 18 * 32; i += K2_THREADS) (&s_ovl[0][0][0])[i] = 0.0f;
    for (int i = tid; i < 2 * 33 * SB_STRIDE; i += K2_THREADS) (&s_ring[0][0][0])[i] = 0.0f;
    for (int i = tid; i < 2 * 22; i += K2_THREADS) (&s_sfl[0][0])[i] = 0;
    for (int i = tid; i < 2 * 39; i += K2_THREADS) (&s_sfs[0][0])[i] = 0;
    __syncthreads();

    const uint32_t warm = run.g0 < 2 ? run.g0 : 2;
    int ring_base = 0;  // row of slot 0 of the current granule is (ring_base + 15) % 33; see row_of()

    for (uint32_t g = run.g0 - warm; g < run.g0 + run.ng; g++) {
        const uint32_t frame = g >> 1, gr = g & 1;
        const FrameRec rec = frames[tr.frame_base + frame];
        const bool ms = nch == 2 && ((rec.hdr >> 6) & 3) == 1 && ((rec.hdr >> 4) & 2);
        const size_t gc_first = (size_t)tr.gc_base + (size_t)frame * 2 * nch + gr * nch;

        // ---- A: side info of this granule (one thread per channel)
        if (tid < nch) {
            const int ch = tid;
            const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
            const int gbase = (nch == 1 ? 18 : 20);
            const int rem = gr * nch + ch;
            uint32_t bit = rec.md_start * 8;
            for (int k = 0; k < rem; k++) bit += side_field(s, gbase + 59 * k, 12);
            const int b0 = gbase + 59 * rem;
            GranInfo gi;
            gi.global_gain = side_field(s, b0 + 21, 8);
            gi.sfc = side_field(s, b0 + 29, 4);
            const int ws = side_field(s, b0 + 33, 1);
            gi.bt = 0; gi.mixed = 0; gi.sbg[0] = gi.sbg[1] = gi.sbg[2] = 0;
            if (ws) {
                gi.bt = side_field(s, b0 + 34, 2);
                gi.mixed = side_field(s, b0 + 36, 1);
                gi.sbg[0] = side_field(s, b0 + 47, 3);
                gi.sbg[1] = side_field(s, b0 + 50, 3);
                gi.sbg[2] = side_field(s, b0 + 53, 3);
            }
            if (gi.bt != 2) gi.mixed = 0;
            gi.preflag = side_field(s, b0 + 56, 1);
            gi.sf_shift = side_field(s, b0 + 57, 1) ? 4 : 2;
            gi.part2_bit = bit;
            gi.md_lim = rec.md_end;
            gi.nz = nz_in[gc_first + ch];
            gi.scfsi = (gr == 1) ? (int)side_field(s, (nch == 1 ? 14 : 12) + 4 * ch, 4) : 0;
            s_gi[ch] = gi;
        }
        __syncthreads();

        // ---- B: scalefactors straight from the bitstream (part 2 has a closed-form layout)
        if (tid < 64 * nch) {
            const int ch = tid >> 6, j = tid & 63;
            const GranInfo gi = s_gi[ch];
            const int slen1 = T->slen[0][gi.sfc], slen2 = T->slen[1][gi.sfc];
            if (gi.bt == 2) {
                if (j < 22) {  // long part: only the first 8 bands of a mixed block carry scalefactors, the rest read as 0
                    int v = 0;
                    if (gi.mixed && j < 8) v = md_bits(mdp, gi.md_lim, gi.part2_bit + j * slen1, slen1);
                    s_sfl[ch][j] = v;
                } else if (j < 22 + 39) {
                    const int q = j - 22, sfb = q / 3;
                    int v = 0;
                    if (sfb < 12 && !(gi.mixed && sfb < 3)) {
                        uint32_t off;
                        if (gi.mixed) off = (sfb < 6) ? (8 + (q - 9)) * slen1 : 17 * slen1 + (q - 18) * slen2;
                        else off = (sfb < 6) ? q * slen1 : 18 * slen1 + (q - 18) * slen2;
                        v = md_bits(mdp, gi.md_lim, gi.part2_bit + off, sfb < 6 ? slen1 : slen2);
                    }
                    s_sfs[ch][q] = v;
                }
            } else if (j < 21) {
                const int grp = j < 6 ? 0 : (j < 11 ? 1 : (j < 16 ? 2 : 3));
                const int lo = grp == 0 ? 0 : (grp == 1 ? 6 : (grp == 2 ? 11 : 16));
                if (!(gi.scfsi & (8 >> grp))) {  // otherwise keep granule 0's value (scfsi)
                    uint32_t off = 0;
                    if (grp > 0 && !(gi.scfsi & 8)) off += 6 * slen1;
                    if (grp > 1 && !(gi.scfsi & 4)) off += 5 * slen1;
                    if (grp > 2 && !(gi.scfsi & 2)) off += 5 * slen2;
                    const int sl = grp < 2 ? slen1 : slen2;
                    off += (j - lo) * sl;
                    s_sfl[ch][j] = md_bits(mdp, gi.md_lim, gi.part2_bit + off, sl);
                }
            } else if (j == 21) s_sfl[ch][21] = 0;
        }
        __syncthreads();

        // ---- C: requantise (+ MS stereo, + short-block reorder folded into the store index)
        for (int line = tid; line < 576; line += K2_THREADS) {
            float v[2] = {0.0f, 0.0f};
            int dst[2] = {line, line};
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                if (ch >= nch) break;
                const GranInfo& gi = s_gi[ch];
                int q = 0;
                if ((uint32_t)line < gi.nz) q = spectra[(gc_first + ch) * 576 + line];
                int e;
                if (gi.bt != 2) {
                    const int sfb = T->line_sfb_long[sr][line];
                    e = gi.global_gain - 210 - gi.sf_shift * (s_sfl[ch][sfb] + (gi.preflag ? T->pretab[sfb] : 0));
                } else {
                    const int code = gi.mixed ? T->line_sfb_mixed[sr][line] : T->line_sfb_short[sr][line];
                    if (gi.mixed && code < 32) e = gi.global_gain - 210 - gi.sf_shift * s_sfl[ch][code];
                    else {
                        const int c2 = gi.mixed ? code - 32 : code;
                        e = gi.global_gain - 210 - 8 * gi.sbg[c2 % 3] - gi.sf_shift * s_sfs[ch][c2];
                    }
                    // where this bitstream-order line lands after the reorder
                    dst[ch] = gi.mixed ? T->scatter_mixed[sr][line] : T->scatter_short[sr][line];
                }
                const int aq = q < 0 ? -q : q;
                const float m = __fmul_rn(__ldg(&T->pow43[aq]), s_gain[e - L3_GAIN_MIN]);
                v[ch] = q < 0 ? -m : m;
            }
            if (ms) {
                const float a = v[0], b = v[1];
                v[0] = __fmul_rn(__fadd_rn(a, b), L3_INV_SQRT2);
                v[1] = __fmul_rn(__fsub_rn(a, b), L3_INV_SQRT2);
            }
            s_xr[0][dst[0]] = v[0];
            if (nch == 2) s_xr[1][dst[1]] = v[1];
        }
        __syncthreads();

        // ---- D: alias reduction
        for (int it = tid; it < nch * 31 * 8; it += K2_THREADS) {
            const int ch = it / 248, r = it % 248, sb = r / 8 + 1, i = r % 8;
            const GranInfo& gi = s_gi[ch];
            const bool doit = gi.bt != 2 || (gi.mixed && sb == 1);
            if (doit) {
                const float lo = s_xr[ch][18 * sb - 1 - i], hi = s_xr[ch][18 * sb + i];
                s_xr[ch][18 * sb - 1 - i] = __fsub_rn(__fmul_rn(lo, s_cs[i]), __fmul_rn(hi, s_ca[i]));
                s_xr[ch][18 * sb + i] = __fadd_rn(__fmul_rn(hi, s_cs[i]), __fmul_rn(lo, s_ca[i]));
            }
        }
        __syncthreads();

        // ---- E: IMDCT (18-point DCT-IV for long blocks, 3 x 12-point for short blocks)
        for (int it = tid; it < nch * 32 * 18; it += K2_THREADS) {
            const int ch = it / 576, r = it % 576, sb = r / 18, m = r % 18;
            const GranInfo& gi = s_gi[ch];
            const float* X = &s_xr[ch][18 * sb];
            const bool is_short = gi.bt == 2 && !(gi.mixed && sb < 2);
            if (!is_short) {
                float a = 0.0f;
#pragma unroll
                for (int k = 0; k < 18; k++) a = fmaf(X[k], s_dct4[m][k], a);
                s_u[ch][sb][m] = a;
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int o = 2 * m + h, w = o / 12, i = o % 12;
                    float a = 0.0f;
#pragma unroll
                    for (int k = 0; k < 6; k++) a = fmaf(X[3 * k + w], s_im12[i][k], a);
                    s_u[ch][sb][o] = __fmul_rn(a, s_win[2][i]);
                }
            }
        }
        __syncthreads();

        // ---- F: window, overlap-add, frequency inversion -> subband samples [slot][subband]
        for (int it = tid; it < nch * 18 * 32; it += K2_THREADS) {
            const int ch = it / 576, r = it % 576, i = r / 32, sb = r % 32;
            const GranInfo& gi = s_gi[ch];
            const bool is_short = gi.bt == 2 && !(gi.mixed && sb < 2);
            float lo, hi;
            if (!is_short) {
                const int bt = (gi.bt == 2) ? 0 : gi.bt;  // long subbands of a mixed block use the normal window
                lo = __fmul_rn(s_u[ch][sb][s_idx36[i]], s_win[bt][i]);
                hi = __fmul_rn(s_u[ch][sb][s_idx36[18 + i]], s_win[bt][18 + i]);
            } else {
                // raw[j] = sum over windows w (ascending) of y_w[j - 6 - 6w]; j = i (lo) and 18 + i (hi)
                const float* y = s_u[ch][sb];
                // lo: j = i
                if (i < 6) lo = 0.0f;
                else if (i < 12) lo = y[i - 6];
                else lo = __fadd_rn(y[i - 6], y[12 + (i - 12)]);
                // hi: j = 18 + i
                if (i < 6) hi = __fadd_rn(y[12 + i + 6], y[24 + i]);
                else if (i < 12) hi = y[24 + i];
                else hi = 0.0f;
            }
            float vv = __fadd_rn(lo, s_ovl[ch][i][sb]);
            s_ovl[ch][i][sb] = hi;
            if ((sb & 1) && (i & 1)) vv = -vv;
            s_sb[ch][i][sb] = vv;
        }
        __syncthreads();

        // ---- G: 32-point DCT-II per time slot (Lee), result into the ring of the last 33 slots
        if (tid < nch * 18) {
            const int ch = tid / 18, t = tid % 18;
            float x[32];
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = s_sb[ch][t][k];
            Dct2<32, 0>::run(x);
            float* dstrow = s_ring[ch][(ring_base + 15 + t) % 33];
#pragma unroll
            for (int k = 0; k < 32; k++) dstrow[k] = x[k];
        }
        __syncthreads();

        // ---- H: polyphase window: 16 taps per output sample, PCM to HBM (coalesced 128-byte rows)
        if (g >= run.g0) {
            for (int it = tid; it < nch * 18 * 32; it += K2_THREADS) {
                const int ch = it / 576, r = it % 576, t = r / 32, j = r % 32;
                const int ia = s_ia[j], ib = s_ib[j];
                const int row0 = ring_base + 15 + t + 33;
                float a = 0.0f;
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    a = fmaf(s_ring[ch][(row0 - 2 * p) % 33][ia], s_wa[p][j], a);
                    a = fmaf(s_ring[ch][(row0 - 2 * p - 1) % 33][ib], s_wb[p][j], a);
                }
                if (flags & 0x4u) a = fminf(1.0f, fmaxf(-1.0f, a));
                pcm[tr.pcm_off + (size_t)ch * tr.pcm_plane + (size_t)g * 576 + 32 * t + j] = a;
            }
        }
        ring_base = (ring_base + 18) % 33;
        __syncthreads();
    }
}
