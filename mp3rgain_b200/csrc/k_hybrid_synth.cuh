// k_hybrid_synth.cuh -- K2: requantise -> MS stereo -> reorder -> alias reduction -> IMDCT/overlap -> polyphase -> PCM.
//
// One CTA owns a run of consecutive granules of one track (both channels) and walks it in time order, keeping the IMDCT
// overlap and the last 15 DCT-32 vectors of the polyphase filterbank in shared memory.  Both are finite-memory, so a run
// that does not start at the head of the track first replays the two granules before it (PCM writes suppressed) to
// rebuild that state exactly: runs are independent and the whole corpus decodes in parallel.
//
// Arithmetic follows DESIGN.md "f32 decode contract" operation for operation.
//
// Instruction budget notes (the kernel is issue-bound, see profiles/):
//  * side info, scalefactors and the per-band gain-table index of the whole run are computed once in a parallel prologue;
//  * the IMDCT is mapped one (channel, subband, half) per thread: half 1 evaluates DCT-IV rows 9..17, which are exactly the
//    18 outputs added to the saved overlap, half 0 rows 0..8, which are exactly the 18 outputs saved for the next granule,
//    so windowing, overlap-add and frequency inversion fuse into the same thread with the spectrum in registers and the
//    matrix rows arriving as warp-wide 128-bit shared-memory broadcasts;
//  * subbands above the last coded line (+1 for the alias butterflies) are skipped (their IMDCT input is exactly zero);
//  * the polyphase window keeps its 16 coefficients and a sliding set of DCT values in registers; the DCT ring uses two
//    static positions plus a mirrored tail so no index is reduced modulo the ring size;
//  * the CTA is warp-specialised: warps 0-3 (front) run the spectral half of granule g+1 (requantise, alias, IMDCT) while
//    warps 4-7 (back) run the filterbank half of granule g (DCT-32, polyphase window, PCM store); the two halves hand the
//    subband samples over through a double buffer guarded by named barriers (bar.arrive / bar.sync), so neither half
//    idles at a CTA-wide barrier while the other works;
//  * each 32-point DCT is split over two threads along Lee's first butterfly (even outputs from the sum branch, odd outputs
//    from the difference branch): same operations per element, half the dependent chain.
#pragma once

#define K2_THREADS 256
#define SB_STRIDE 33       // padded row of 32 subband values
#define OVL_STRIDE 19      // padded row of 18 overlap values
#define K2_MAX_G 66        // RUN_GRANULES (64) + 2 warm-up granules
#define RING_ROWS 51       // rows 0..50 of the DCT-32 ring (see synth step)

struct __align__(16) GranInfo {
    uint32_t part2_bit;   // absolute bit position of part 2 in the main-data stream
    uint32_t md_lim;      // bytes readable (later bytes read as zero)
    uint16_t nz;          // lines decoded by K1 (0 = silent)
    uint8_t global_gain, sfc, bt, mixed, preflag, sf_shift /* 2 or 4 */;
    uint8_t sbg[3], scfsi, ms, sb_limit /* subbands with possibly non-zero IMDCT input */, pad[2];
};

struct K2Smem {
    alignas(16) float xr[2][576];       // requantised, reordered spectrum
    float sb[2][2][18][SB_STRIDE];      // subband samples [buffer][channel][slot][subband] (front -> back hand-off)
    float ovl[2][2][32][OVL_STRIDE];    // IMDCT overlap, ping-pong by granule parity
    float ring[2][RING_ROWS][SB_STRIDE];
    alignas(16) float dct4[18][20];     // 18-point DCT-IV matrix, rows padded to 5 float4
    float win[4][36];
    alignas(16) float im12[12][8];      // 12-point IMDCT matrix, rows padded to 2 float4
    float wa[8][32], wb[8][32];
    float gain[L3_GAIN_MAX - L3_GAIN_MIN + 1];
    float cs[8], ca[8];
    GranInfo gi[K2_MAX_G][2];
    uint16_t eidx[K2_MAX_G][2][64];     // per band: index into gain[] ([0..21] long bands, [22..60] short sfb*3+win)
    uint16_t scat[2][576];              // bitstream-order line -> reordered line: [0] pure short, [1] mixed
    uint8_t code[3][576];               // line -> band code: [0] long, [1] pure short, [2] mixed
    uint8_t ia[32], ib[32];
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

// long-block scalefactor `sfb` (0..20) of a granule-channel whose part 2 starts at gi.part2_bit; `scfsi` = bands skipped
__device__ __forceinline__ int read_long_sf(const uint8_t* __restrict__ mdp, const GranInfo& gi, int scfsi, int sfb, int slen1, int slen2) {
    const int grp = sfb < 6 ? 0 : (sfb < 11 ? 1 : (sfb < 16 ? 2 : 3));
    const int lo = grp == 0 ? 0 : (grp == 1 ? 6 : (grp == 2 ? 11 : 16));
    uint32_t off = 0;
    if (grp > 0 && !(scfsi & 8)) off += 6 * slen1;
    if (grp > 1 && !(scfsi & 4)) off += 5 * slen1;
    if (grp > 2 && !(scfsi & 2)) off += 5 * slen2;
    const int sl = grp < 2 ? slen1 : slen2;
    off += (sfb - lo) * sl;
    return (int)md_bits(mdp, gi.md_lim, gi.part2_bit + off, sl);
}

// position of IMDCT-36 output i in the 18-point DCT-IV vector u: i<9: u[i+9]; 9..26: u[26-i]; 27..35: u[i-27]
__device__ __forceinline__ constexpr int idx36(int i) { return i < 9 ? i + 9 : (i < 27 ? 26 - i : i - 27); }

__global__ void __launch_bounds__(K2_THREADS, 3)
k_hybrid_synth(const uint8_t* __restrict__ corpus, const uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
               const DecTrack* __restrict__ tracks, const RunDesc* __restrict__ runs, const L3Dev* __restrict__ T,
               const int16_t* __restrict__ spectra, const uint16_t* __restrict__ nz_in, float* __restrict__ pcm, uint32_t flags) {
    extern __shared__ __align__(16) unsigned char k2_raw[];
    K2Smem& S = *reinterpret_cast<K2Smem*>(k2_raw);

    const int tid = threadIdx.x;
    const RunDesc run = runs[blockIdx.x];
    const DecTrack tr = tracks[run.track];
    const int nch = (int)tr.channels;
    const int sr = (int)tr.sr_idx;
    const uint8_t* mdp = md + tr.md_off;
    const uint32_t warm = run.g0 < 2 ? run.g0 : 2;
    const uint32_t gfirst = run.g0 - warm, gcount = run.ng + warm;

    // ---- one-time table staging and state reset
    for (int i = tid; i < 18 * 20; i += K2_THREADS) { const int m = i / 20, k = i - 20 * m; S.dct4[m][k] = k < 18 ? T->dct4_18[m][k] : 0.0f; }
    for (int i = tid; i < 4 * 36; i += K2_THREADS) (&S.win[0][0])[i] = (&T->winsgn[0][0])[i];
    for (int i = tid; i < 12 * 8; i += K2_THREADS) { const int r = i >> 3, k = i & 7; S.im12[r][k] = k < 6 ? T->imdct12[r][k] : 0.0f; }
    for (int i = tid; i < 256; i += K2_THREADS) { (&S.wa[0][0])[i] = (&T->synth_wa[0][0])[i]; (&S.wb[0][0])[i] = (&T->synth_wb[0][0])[i]; }
    for (int i = tid; i < L3_GAIN_MAX - L3_GAIN_MIN + 1; i += K2_THREADS) S.gain[i] = T->gain[i];
    for (int i = tid; i < 576; i += K2_THREADS) {
        S.code[0][i] = T->line_sfb_long[sr][i];
        S.code[1][i] = 22 + T->line_sfb_short[sr][i];
        const uint8_t c = T->line_sfb_mixed[sr][i];
        S.code[2][i] = c < 32 ? c : 22 + (c - 32);
        S.scat[0][i] = T->scatter_short[sr][i];
        S.scat[1][i] = T->scatter_mixed[sr][i];
    }
    if (tid < 8) { S.cs[tid] = T->alias_cs[tid]; S.ca[tid] = T->alias_ca[tid]; }
    if (tid < 32) { S.ia[tid] = T->synth_ia[tid]; S.ib[tid] = T->synth_ib[tid]; }
    for (int i = tid; i < 2 * 2 * 32 * OVL_STRIDE; i += K2_THREADS) (&S.ovl[0][0][0][0])[i] = 0.0f;
    for (int i = tid; i < 2 * RING_ROWS * SB_STRIDE; i += K2_THREADS) (&S.ring[0][0][0])[i] = 0.0f;

    // ---- prologue 1: side info of every granule-channel of the run
    for (uint32_t it = tid; it < gcount * nch; it += K2_THREADS) {
        const uint32_t gl = it / nch, ch = it % nch, g = gfirst + gl;
        const uint32_t frame = g >> 1, gr = g & 1;
        const FrameRec rec = frames[tr.frame_base + frame];
        const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
        const int gbase = (nch == 1 ? 18 : 20);
        const int rem = gr * nch + ch;
        uint32_t bit = rec.md_start * 8;
        for (int k = 0; k < rem; k++) bit += side_field(s, gbase + 59 * k, 12);
        const int b0 = gbase + 59 * rem;
        GranInfo gi;
        gi.global_gain = (uint8_t)side_field(s, b0 + 21, 8);
        gi.sfc = (uint8_t)side_field(s, b0 + 29, 4);
        const int ws = side_field(s, b0 + 33, 1);
        gi.bt = 0; gi.mixed = 0; gi.sbg[0] = gi.sbg[1] = gi.sbg[2] = 0;
        if (ws) {
            gi.bt = (uint8_t)side_field(s, b0 + 34, 2);
            gi.mixed = (uint8_t)side_field(s, b0 + 36, 1);
            gi.sbg[0] = (uint8_t)side_field(s, b0 + 47, 3);
            gi.sbg[1] = (uint8_t)side_field(s, b0 + 50, 3);
            gi.sbg[2] = (uint8_t)side_field(s, b0 + 53, 3);
        }
        if (gi.bt != 2) gi.mixed = 0;
        gi.preflag = (uint8_t)side_field(s, b0 + 56, 1);
        gi.sf_shift = side_field(s, b0 + 57, 1) ? 4 : 2;
        gi.part2_bit = bit;
        gi.md_lim = rec.md_end;
        const size_t gcb = (size_t)tr.gc_base + (size_t)frame * 2 * nch + gr * nch;
        gi.nz = nz_in[gcb + ch];
        gi.scfsi = (gr == 1) ? (uint8_t)side_field(s, (nch == 1 ? 14 : 12) + 4 * ch, 4) : 0;
        gi.ms = (nch == 2 && ((rec.hdr >> 6) & 3) == 1 && ((rec.hdr >> 4) & 2)) ? 1 : 0;
        // subbands whose IMDCT input can be non-zero: the coded lines (of either channel under MS), one more subband for
        // the alias butterflies; short blocks scatter lines across the whole band, so they keep all 32
        uint32_t nzm = gi.nz;
        if (gi.ms) nzm = max(nzm, (uint32_t)nz_in[gcb + (ch ^ 1)]);
        gi.sb_limit = nzm == 0 ? 0 : (gi.bt == 2 ? 32 : (uint8_t)min(32u, (nzm + 17) / 18 + 1));
        gi.pad[0] = gi.pad[1] = 0;
        S.gi[gl][ch] = gi;
    }
    __syncthreads();

    // ---- prologue 2: scalefactors straight from the bitstream (part 2 has a closed-form layout) -> gain-table index per
    // band.  One warp per granule-channel, one lane per band: every branch on the block type is warp-uniform.
    {
        const int lane = tid & 31;
        for (uint32_t gc = tid >> 5; gc < gcount * nch; gc += K2_THREADS / 32) {
            const uint32_t gl = gc / nch, ch = gc - gl * nch;
            const GranInfo gi = S.gi[gl][ch];
            const int slen1 = T->slen[0][gi.sfc], slen2 = T->slen[1][gi.sfc];
            const int base_e = (int)gi.global_gain - 210 - L3_GAIN_MIN;
            if (gi.bt != 2) {
                if (lane < 22) {
                    int v = 0;
                    if (lane < 21) {
                        const int grp = lane < 6 ? 0 : (lane < 11 ? 1 : (lane < 16 ? 2 : 3));
                        if (!(gi.scfsi & (8 >> grp))) v = read_long_sf(mdp, gi, gi.scfsi, lane, slen1, slen2);
                        else {
                            // scfsi: reuse granule 0's scalefactor of the same frame (granule gl-1 of this run; runs start
                            // on an even granule).  A short granule 0 has no long scalefactors (they read as 0) except
                            // the 8 of a mixed block.
                            const GranInfo g0 = S.gi[gl - 1][ch];
                            const int a1 = T->slen[0][g0.sfc], a2 = T->slen[1][g0.sfc];
                            if (g0.bt != 2) v = read_long_sf(mdp, g0, 0, lane, a1, a2);
                            else if (g0.mixed && lane < 8) v = md_bits(mdp, g0.md_lim, g0.part2_bit + lane * a1, a1);
                        }
                    }
                    S.eidx[gl][ch][lane] = (uint16_t)(base_e - gi.sf_shift * (v + (gi.preflag ? T->pretab[lane] : 0)));
                }
            } else {
                if (lane < 22) {  // long part: only the first 8 bands of a mixed block carry scalefactors
                    int v = 0;
                    if (gi.mixed && lane < 8) v = md_bits(mdp, gi.md_lim, gi.part2_bit + lane * slen1, slen1);
                    S.eidx[gl][ch][lane] = (uint16_t)(base_e - gi.sf_shift * (v + (gi.preflag ? T->pretab[lane] : 0)));
                }
                for (int q = lane; q < 39; q += 32) {
                    const int sfb = q / 3, win = q - 3 * sfb;
                    int v = 0;
                    if (sfb < 12 && !(gi.mixed && sfb < 3)) {
                        uint32_t off;
                        if (gi.mixed) off = (sfb < 6) ? (8 + (q - 9)) * slen1 : 17 * slen1 + (q - 18) * slen2;
                        else off = (sfb < 6) ? q * slen1 : 18 * slen1 + (q - 18) * slen2;
                        v = md_bits(mdp, gi.md_lim, gi.part2_bit + off, sfb < 6 ? slen1 : slen2);
                    }
                    S.eidx[gl][ch][22 + q] = (uint16_t)(base_e - 8 * gi.sbg[win] - gi.sf_shift * v);
                }
            }
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------------------------------------------------------
    // Warp specialisation.  Named barriers: 1 = front-internal, 2 = back-internal, 3+b = "sb[b] full", 5+b = "sb[b] empty".
    auto bar_sync = [](int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); };
    auto bar_arrive = [](int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); };

    if (tid < 128) {
        // ===================================== FRONT: requantise -> alias -> IMDCT/overlap =============================
        // The requantiser's input (packed int16 line pairs tid, tid+128, tid+256 per channel) is fetched one granule ahead,
        // so the HBM latency of the spectra overlaps a whole granule of compute.
        uint32_t pre[3][2];
        auto fetch = [&](uint32_t gl_, uint32_t (&dstv)[3][2]) {
            const uint32_t g_ = gfirst + gl_;
            const size_t gcf = (size_t)tr.gc_base + (size_t)(g_ >> 1) * 2 * nch + (g_ & 1) * nch;
#pragma unroll
            for (int sl = 0; sl < 3; sl++) {
                const int pr = tid + 128 * sl;
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    uint32_t v = 0;
                    if (pr < 288 && ch < nch && (uint32_t)(2 * pr) < S.gi[gl_][ch].nz)
                        v = __ldg(reinterpret_cast<const uint32_t*>(spectra + (gcf + ch) * 576 + 2 * pr));
                    dstv[sl][ch] = v;
                }
            }
        };
        fetch(0, pre);
        for (uint32_t gl = 0; gl < gcount; gl++) {
            uint32_t cur[3][2];
#pragma unroll
            for (int sl = 0; sl < 3; sl++) { cur[sl][0] = pre[sl][0]; cur[sl][1] = pre[sl][1]; }
            if (gl + 1 < gcount) fetch(gl + 1, pre);
            const int par = (int)(gl & 1);
            const int lim0 = S.gi[gl][0].sb_limit, lim1 = nch == 2 ? S.gi[gl][1].sb_limit : 0;
            const int pair_lim = 9 * max(lim0, lim1);        // line pairs that any active subband reads

            // ---- C: requantise (+ MS stereo, + short-block reorder folded into the store index), two lines per item
#pragma unroll
            for (int sl = 0; sl < 3; sl++) {
                const int pr = tid + 128 * sl;
                if (pr >= pair_lim) break;
                float v[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
                int dst[2][2] = {{2 * pr, 2 * pr + 1}, {2 * pr, 2 * pr + 1}};
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    if (ch >= nch) break;
                    const GranInfo& gi = S.gi[gl][ch];
                    const uint32_t packed = cur[sl][ch];
                    const int tsel = gi.bt != 2 ? 0 : (gi.mixed ? 2 : 1);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int line = 2 * pr + h;
                        const int q = (int)(int16_t)(h ? (packed >> 16) : (packed & 0xFFFFu));
                        const int aq = q < 0 ? -q : q;
                        const float m = __fmul_rn(__ldg(&T->pow43[aq]), S.gain[S.eidx[gl][ch][S.code[tsel][line]]]);
                        v[ch][h] = q < 0 ? -m : m;
                        if (tsel) dst[ch][h] = S.scat[tsel - 1][line];
                    }
                }
                if (S.gi[gl][0].ms) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float a = v[0][h], b = v[1][h];
                        v[0][h] = __fmul_rn(__fadd_rn(a, b), L3_INV_SQRT2);
                        v[1][h] = __fmul_rn(__fsub_rn(a, b), L3_INV_SQRT2);
                    }
                }
                S.xr[0][dst[0][0]] = v[0][0];
                S.xr[0][dst[0][1]] = v[0][1];
                if (nch == 2) { S.xr[1][dst[1][0]] = v[1][0]; S.xr[1][dst[1][1]] = v[1][1]; }
            }
            bar_sync(1, 128);

            // ---- D: alias reduction (8 butterflies per boundary between active subbands); it = ch*256 + (sb-1)*8 + i
            for (int it = tid; it < nch * 256; it += 128) {
                const int ch = it >> 8, r = it & 255, sb = (r >> 3) + 1, i = r & 7;
                const GranInfo& gi = S.gi[gl][ch];
                if (sb < (int)gi.sb_limit && (gi.bt != 2 || (gi.mixed && sb == 1))) {
                    const float lo = S.xr[ch][18 * sb - 1 - i], hi = S.xr[ch][18 * sb + i];
                    S.xr[ch][18 * sb - 1 - i] = __fsub_rn(__fmul_rn(lo, S.cs[i]), __fmul_rn(hi, S.ca[i]));
                    S.xr[ch][18 * sb + i] = __fadd_rn(__fmul_rn(hi, S.cs[i]), __fmul_rn(lo, S.ca[i]));
                }
            }
            bar_sync(1, 128);
            if (gl >= 2) bar_sync(5 + par, 256);   // the back half has finished reading sb[par] (granule gl-2)

            // ---- E+F: IMDCT + window + overlap-add + frequency inversion, one (channel, half, subband) per thread.
            // half 1 -> the 18 outputs of this granule (DCT-IV rows 9..17), half 0 -> the 18 outputs saved as overlap (rows 0..8).
            if (tid < 64 * nch) {
                const int ch = tid >> 6, half = (tid >> 5) & 1, sb = tid & 31;
                const GranInfo& gi = S.gi[gl][ch];
                const bool active = sb < (int)gi.sb_limit;
                const bool is_short = gi.bt == 2 && !(gi.mixed && sb < 2);
                float o[18];   // half 1: raw[0..17]; half 0: raw[18..35]
                if (!active) {
#pragma unroll
                    for (int i = 0; i < 18; i++) o[i] = 0.0f;
                } else {
                    float X[18];
                    const float2* X2 = reinterpret_cast<const float2*>(&S.xr[ch][18 * sb]);
#pragma unroll
                    for (int k = 0; k < 9; k++) { const float2 x = X2[k]; X[2 * k] = x.x; X[2 * k + 1] = x.y; }
                    if (!is_short) {
                        const int bt = (gi.bt == 2) ? 0 : gi.bt;  // long subbands of a mixed block use the normal window
                        float u[9];
#pragma unroll
                        for (int mm = 0; mm < 9; mm++) {
                            const float4* C4 = reinterpret_cast<const float4*>(S.dct4[9 * half + mm]);
                            float a = 0.0f;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const float4 c = C4[q];
                                a = fmaf(X[4 * q], c.x, a); a = fmaf(X[4 * q + 1], c.y, a);
                                a = fmaf(X[4 * q + 2], c.z, a); a = fmaf(X[4 * q + 3], c.w, a);
                            }
                            const float4 c = C4[4];
                            a = fmaf(X[16], c.x, a); a = fmaf(X[17], c.y, a);
                            u[mm] = a;
                        }
                        const float* wrow = &S.win[bt][18 * (1 - half)];
#pragma unroll
                        for (int i = 0; i < 18; i++) {
                            // half 1: output i uses u[idx36(i)] (9..17); half 0: output 18+i uses u[idx36(18+i)] (0..8)
                            const float uv = half ? u[idx36(i) - 9] : u[idx36(18 + i)];
                            o[i] = __fmul_rn(uv, wrow[i]);
                        }
                    } else {
                        // three 12-point IMDCTs y_w (w = 0,1,2) placed at raw[6+6w ..]; overlaps added in ascending window
                        // order.  half 1 needs y0[0..11], y1[0..5]; half 0 needs y1[6..11], y2[0..11]: 18 dot products each.
                        float y[18];
#pragma unroll
                        for (int d = 0; d < 18; d++) {
                            const int w = half ? (d < 12 ? 0 : 1) : (d < 6 ? 1 : 2);
                            const int i = half ? (d < 12 ? d : d - 12) : (d < 6 ? d + 6 : d - 6);
                            const float4 c0 = *reinterpret_cast<const float4*>(&S.im12[i][0]);
                            const float2 c1 = *reinterpret_cast<const float2*>(&S.im12[i][4]);
                            float a = 0.0f;
                            a = fmaf(X[w], c0.x, a); a = fmaf(X[3 + w], c0.y, a); a = fmaf(X[6 + w], c0.z, a);
                            a = fmaf(X[9 + w], c0.w, a); a = fmaf(X[12 + w], c1.x, a); a = fmaf(X[15 + w], c1.y, a);
                            y[d] = __fmul_rn(a, S.win[2][i]);
                        }
                        if (half) {
#pragma unroll
                            for (int i = 0; i < 18; i++) o[i] = i < 6 ? 0.0f : (i < 12 ? y[i - 6] : __fadd_rn(y[i - 6], y[12 + (i - 12)]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 18; i++) o[i] = i < 6 ? __fadd_rn(y[i], y[6 + i]) : (i < 12 ? y[6 + i] : 0.0f);
                        }
                    }
                }
                if (half) {
                    const float* prev = S.ovl[par][ch][sb];
#pragma unroll
                    for (int i = 0; i < 18; i++) {
                        float vv = __fadd_rn(o[i], prev[i]);
                        if ((sb & 1) && (i & 1)) vv = -vv;
                        S.sb[par][ch][i][sb] = vv;
                    }
                } else {
                    float* nxt = S.ovl[par ^ 1][ch][sb];
#pragma unroll
                    for (int i = 0; i < 18; i++) nxt[i] = o[i];
                }
            }
            __threadfence_block();
            bar_arrive(3 + par, 256);   // sb[par] is full
            bar_sync(1, 128);           // all reads of xr / writes of ovl done before the next granule overwrites them
        }
    } else {
        // ===================================== BACK: DCT-32 -> polyphase window -> PCM ====================================
        const int bt_ = tid - 128;
        // polyphase window step: thread (ch, half, j) keeps its 16 window coefficients in registers
        const int w_ch = bt_ >> 6, w_j = bt_ & 31, w_half = (bt_ >> 5) & 1;
        float wa[8], wb[8];
#pragma unroll
        for (int p = 0; p < 8; p++) { wa[p] = S.wa[p][w_j]; wb[p] = S.wb[p][w_j]; }
        const int w_ia = S.ia[w_j], w_ib = S.ib[w_j];
        // DCT step: thread (d, br): d = (channel, slot), br = 0 even outputs (sum branch) / 1 odd outputs (difference branch)
        const int d_br = (bt_ >> 5) & 1, d_idx = (bt_ & 31) + 32 * (bt_ >> 6);
        const bool d_on = d_idx < 18 * nch && d_idx < 36 && (bt_ & 31) < (bt_ < 64 ? 32 : 4);
        const int d_ch = d_idx / 18, d_t = d_idx - 18 * d_ch;

        for (uint32_t gl = 0; gl < gcount; gl++) {
            const uint32_t g = gfirst + gl;
            const int par = (int)(gl & 1);
            const int ring_base = 15 + 18 * par;   // row of slot 0: granules alternate between rows 15.. and 33..
            bar_sync(3 + par, 256);                // wait for the front half to fill sb[par]

            // ---- G: 32-point DCT-II per time slot (Lee), two threads per transform.  Ring rows: this granule's slots at
            // ring_base + t; in the upper position slots 3..17 are mirrored into rows 0..14 so the next granule (lower
            // position) finds its 15 predecessors at negative offsets without any modular arithmetic.
            if (d_on) {
                const float* x = S.sb[par][d_ch][d_t];
                float z[16];
                if (d_br == 0) {
#pragma unroll
                    for (int i = 0; i < 16; i++) z[i] = __fadd_rn(x[i], x[31 - i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) z[i] = __fmul_rn(__fsub_rn(x[i], x[31 - i]), c_dct_sec[i]);
                }
                Dct2<16, 16>::run(z);
                float* dstrow = S.ring[d_ch][ring_base + d_t];
                float* mir = S.ring[d_ch][d_t >= 3 ? d_t - 3 : 0];
                const bool do_mir = par && d_t >= 3;
                if (d_br == 0) {
#pragma unroll
                    for (int k = 0; k < 16; k++) { dstrow[2 * k] = z[k]; if (do_mir) mir[2 * k] = z[k]; }
                } else {
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const float v = k < 15 ? __fadd_rn(z[k], z[k + 1]) : z[15];
                        dstrow[2 * k + 1] = v;
                        if (do_mir) mir[2 * k + 1] = v;
                    }
                }
            }
            __threadfence_block();
            if (gl + 2 < gcount) bar_arrive(5 + par, 256);   // sb[par] may be overwritten (by granule gl+2)
            bar_sync(2, 128);                                // ring rows of this granule are complete

            // ---- H: polyphase window.  Thread (ch, half, j) produces 9 consecutive slots of output column j, sliding a
            // register window over the ring: out[t][j] = sum_p A[t-2p]*wa[p] + B[t-2p-1]*wb[p]  (p ascending, one fmaf chain)
            if (g >= run.g0 && w_ch < nch) {
                const float* colA = &S.ring[w_ch][ring_base + 9 * w_half][w_ia];
                const float* colB = &S.ring[w_ch][ring_base + 9 * w_half][w_ib];
                float A[24], B[24];   // rows t0-15 .. t0+8 relative to this thread's first slot t0
#pragma unroll
                for (int r = 0; r < 15; r++) { A[r] = colA[(r - 15) * SB_STRIDE]; B[r] = colB[(r - 15) * SB_STRIDE]; }
                float* dst = pcm + tr.pcm_off + (size_t)w_ch * tr.pcm_plane + (size_t)g * 576 + 32 * (9 * w_half) + w_j;
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    A[15 + t] = colA[t * SB_STRIDE];
                    B[15 + t] = colB[t * SB_STRIDE];
                    float a = 0.0f;
#pragma unroll
                    for (int p = 0; p < 8; p++) {
                        a = fmaf(A[15 + t - 2 * p], wa[p], a);
                        a = fmaf(B[15 + t - 2 * p - 1], wb[p], a);
                    }
                    if (flags & 0x4u) a = fminf(1.0f, fmaxf(-1.0f, a));
                    dst[32 * t] = a;
                }
            }
            bar_sync(2, 128);   // ring reads done before the next granule's DCT overwrites the other position / mirror rows
        }
    }
}
