// k_hybrid_synth.cuh -- K2: requantise -> MS stereo -> reorder -> alias reduction -> IMDCT/overlap -> polyphase -> PCM.
//
// One CTA owns a run of consecutive granules of one track (both channels) and walks it in time order in BATCHES of K2_G
// granules.  Inside a batch every stage is data-parallel over (granule, channel, subband | slot | column), so each stage is
// one phase in which all warps do the same kind of work, separated by one CTA barrier:
//
//   P0  side info + scalefactors -> per-band gain-table index      one warp per granule-channel (runs one batch ahead,
//                                                                   on the two warps that have no P4 work)
//   P1  requantise + MS + short-block reorder -> xr                one line pair (both channels) per thread-iteration
//   P2  alias butterflies + IMDCT + window -> F (first 18 outputs) one (granule, channel, subband) per thread
//                                            S2 (last 18 outputs = the overlap the NEXT granule adds)
//   P3  x = F + S2(prev granule); 32-point DCT-II (Lee) -> ring    two threads per transform (sum / difference branch),
//                                                                   the two branches on different warps
//   P4  16-tap polyphase window over the ring -> PCM               one (granule, channel, output column) per thread, 18 slots
//
// The overlap-add is performed by the reader (P3), so P2 has no dependency between granules; the overlap and the
// filterbank history travel between batches in 5-slot rings (S2, ring) indexed by granule number modulo 5, so nothing is
// copied.  Both are finite-memory: a run that does not start at the head of the track first replays the two granules
// before it (PCM writes suppressed) to rebuild that state exactly, so runs are independent and the corpus decodes in
// parallel.
//
// Arithmetic follows DESIGN.md "f32 decode contract" operation for operation (frequency inversion is applied to both
// addends instead of the sum: -(a+b) == (-a)+(-b) in IEEE arithmetic).
#pragma once

#define K2_THREADS 192
#define K2_G 2             // granules per batch
#define K2_RING (K2_G + 1) // slots of the S2 / DCT rings
#define K2_MAX_G 64        // granules per run, including the two replayed ones
#define ROW 33             // padded row of 32 values

struct __align__(8) GranInfo {
    uint32_t part2_bit;   // absolute bit position of part 2 in the main-data stream
    uint32_t md_lim;      // bytes readable (later bytes read as zero)
    uint16_t nz;          // lines decoded by K1 (0 = silent)
    uint8_t global_gain, sfc, bt, mixed, preflag, sf_shift /* 2 or 4 */;
    uint8_t sbg[3], scfsi, ms, sb_limit /* subbands with possibly non-zero IMDCT input */, pad[2];
};

struct K2Smem {
    // DCT-32 output rows; granule n owns rows 18*(n%3) .. +17 of each channel.  The same slot first holds the granule's
    // requantised spectrum xr (576 floats per channel): xr is dead once P2 has read it, before P3 writes the rows.
    alignas(16) float ring[2][K2_RING * 18][ROW];
    float F[K2_G][2][18][ROW];                // IMDCT outputs 0..17  [slot][subband]
    float S2[K2_RING][2][18][ROW];            // IMDCT outputs 18..35 of granule n live in slot (n+1)%3
    alignas(16) float dct4[18][20];           // 18-point DCT-IV matrix, rows padded to 5 float4
    float win[4][36];
    alignas(16) float im12[12][8];            // 12-point IMDCT matrix, rows padded to 2 float4
    float wa[8][32], wb[8][32];
    float gain[L3_GAIN_MAX - L3_GAIN_MIN + 1];
    GranInfo gi[K2_MAX_G][2];                 // side info of every granule-channel of the run
    uint16_t eidx[2][K2_G][2][64];            // [batch parity] per band: index into gain[] ([0..21] long, [22..60] short sfb*3+win)
    uint16_t scat[2][576];                    // bitstream-order line -> reordered line: [0] pure short, [1] mixed
    uint8_t code[3][576];                     // line -> band code: [0] long, [1] pure short, [2] mixed
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

// side info of granule g (0 = first granule of the track), channel ch
__device__ __forceinline__ GranInfo k2_parse_gi(const uint8_t* __restrict__ corpus, const DecTrack& tr, const FrameRec* __restrict__ frames,
                                                const uint16_t* __restrict__ nz_in, uint32_t g, uint32_t ch, int nch) {
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
    return gi;
}

__global__ void __launch_bounds__(K2_THREADS, 4)
k_hybrid_synth(const uint8_t* __restrict__ corpus, const uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
               const DecTrack* __restrict__ tracks, const RunDesc* __restrict__ runs, const L3Dev* __restrict__ T,
               const int16_t* __restrict__ spectra, const uint16_t* __restrict__ nz_in, float* __restrict__ pcm, uint32_t flags) {
    extern __shared__ __align__(16) unsigned char k2_raw[];
    K2Smem& S = *reinterpret_cast<K2Smem*>(k2_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const RunDesc run = runs[blockIdx.x];
    const DecTrack tr = tracks[run.track];
    const int nch = (int)tr.channels;
    const int sr = (int)tr.sr_idx;
    const uint8_t* mdp = md + tr.md_off;
    const uint32_t warm = run.g0 < 2 ? run.g0 : 2;
    const uint32_t gfirst = run.g0 - warm, gcount = run.ng + warm;   // gcount <= K2_MAX_G
    const uint32_t n_batches = (gcount + K2_G - 1) / K2_G;

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
    if (tid < 32) { S.ia[tid] = T->synth_ia[tid]; S.ib[tid] = T->synth_ib[tid]; }
    for (int i = tid; i < K2_RING * 2 * 18 * ROW; i += K2_THREADS) { (&S.S2[0][0][0][0])[i] = 0.0f; (&S.ring[0][0][0])[i] = 0.0f; }
    // side info of every granule-channel of the run, one thread each
    for (uint32_t it = tid; it < gcount * nch; it += K2_THREADS) {
        const uint32_t gl = it / nch, ch = it - gl * nch;
        S.gi[gl][ch] = k2_parse_gi(corpus, tr, frames, nz_in, gfirst + gl, ch, nch);
    }
    __syncthreads();

    // ---- P0: scalefactors of granule-channel q of batch b -> eidx[b&1].  One warp per granule-channel; every branch on the
    // block type is warp-uniform.  Part 2 of the main data has a closed-form layout, so each lane reads its own band's
    // scalefactor straight from the bitstream.
    auto p0_gc = [&](uint32_t b, int q) {
        const int gi_ = q / nch, ch = q - gi_ * nch;
        const uint32_t gl = b * K2_G + gi_;
        if (gl >= gcount) return;
        const GranInfo gi = S.gi[gl][ch];
        const int slen1 = T->slen[0][gi.sfc], slen2 = T->slen[1][gi.sfc];
        const int base_e = (int)gi.global_gain - 210 - L3_GAIN_MIN;
        uint16_t* eidx = S.eidx[b & 1][gi_][ch];
        if (gi.bt != 2) {
            if (lane < 22) {
                int v = 0;
                if (lane < 21) {
                    const int grp = lane < 6 ? 0 : (lane < 11 ? 1 : (lane < 16 ? 2 : 3));
                    if (!(gi.scfsi & (8 >> grp))) v = read_long_sf(mdp, gi, gi.scfsi, lane, slen1, slen2);
                    else {
                        // scfsi: reuse granule 0's scalefactor of the same frame (the previous granule of the run: runs start
                        // on an even granule).  A short granule 0 has no long scalefactors (they read as 0) except the 8
                        // of a mixed block.
                        const GranInfo g0 = S.gi[gl - 1][ch];
                        const int a1 = T->slen[0][g0.sfc], a2 = T->slen[1][g0.sfc];
                        if (g0.bt != 2) v = read_long_sf(mdp, g0, 0, lane, a1, a2);
                        else if (g0.mixed && lane < 8) v = md_bits(mdp, g0.md_lim, g0.part2_bit + lane * a1, a1);
                    }
                }
                eidx[lane] = (uint16_t)(base_e - gi.sf_shift * (v + (gi.preflag ? T->pretab[lane] : 0)));
            }
        } else {
            if (lane < 22) {  // long part: only the first 8 bands of a mixed block carry scalefactors
                int v = 0;
                if (gi.mixed && lane < 8) v = md_bits(mdp, gi.md_lim, gi.part2_bit + lane * slen1, slen1);
                eidx[lane] = (uint16_t)(base_e - gi.sf_shift * (v + (gi.preflag ? T->pretab[lane] : 0)));
            }
            for (int q2 = lane; q2 < 39; q2 += 32) {
                const int sfb = q2 / 3, win = q2 - 3 * sfb;
                int v = 0;
                if (sfb < 12 && !(gi.mixed && sfb < 3)) {
                    uint32_t off;
                    if (gi.mixed) off = (sfb < 6) ? (8 + (q2 - 9)) * slen1 : 17 * slen1 + (q2 - 18) * slen2;
                    else off = (sfb < 6) ? q2 * slen1 : 18 * slen1 + (q2 - 18) * slen2;
                    v = md_bits(mdp, gi.md_lim, gi.part2_bit + off, sfb < 6 ? slen1 : slen2);
                }
                eidx[22 + q2] = (uint16_t)(base_e - 8 * gi.sbg[win] - gi.sf_shift * v);
            }
        }
    };
    // The requantiser's input (packed int16 line pairs, K2_G*288 pairs per channel) is fetched one batch ahead into
    // registers, so its HBM latency overlaps a whole batch of compute.  Pairs at or above nz are zero and are not read.
    constexpr int P1_IT = K2_G * 288 / K2_THREADS;   // 3
    uint32_t pre[P1_IT][2];
    auto fetch = [&](uint32_t b) {
#pragma unroll
        for (int r = 0; r < P1_IT; r++) {
            const int it = tid + K2_THREADS * r;
            const int gi_ = it / 288, pr = it - 288 * gi_;
            const uint32_t gl = b * K2_G + gi_;
            pre[r][0] = pre[r][1] = 0;
            if (gl < gcount) {
                const uint32_t g = gfirst + gl;
                const size_t gcf = (size_t)tr.gc_base + (size_t)(g >> 1) * 2 * nch + (g & 1) * nch;
#pragma unroll
                for (int ch = 0; ch < 2; ch++)
                    if (ch < nch && (uint32_t)(2 * pr) < S.gi[gl][ch].nz)
                        pre[r][ch] = __ldg(reinterpret_cast<const uint32_t*>(spectra + (gcf + ch) * 576 + 2 * pr));
            }
        }
    };
    fetch(0);
    for (int q = warp; q < K2_G * nch; q += K2_THREADS / 32) p0_gc(0, q);
    __syncthreads();

    for (uint32_t b = 0; b < n_batches; b++) {
        const int par = (int)(b & 1);
        const uint32_t gl0 = b * K2_G;   // first granule of the batch, relative to gfirst

        // ================= P1: requantise (+ MS stereo, + short-block reorder folded into the store index) =================
        // One line pair of both channels per iteration; scalefactor bands start on even lines, so a pair shares its band.
#pragma unroll
        for (int r = 0; r < P1_IT; r++) {
            const int it = tid + K2_THREADS * r;
            const int gi_ = it / 288, pr = it - 288 * gi_;
            const uint32_t gl = gl0 + gi_;
            if (gl >= gcount) continue;
            const GranInfo& ga = S.gi[gl][0];
            const GranInfo& gb = S.gi[gl][nch - 1];
            if (pr >= 9 * max((int)ga.sb_limit, (int)gb.sb_limit)) continue;   // no active subband reads these lines
            const int slot = (int)(gl % K2_RING);
            float v[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
            int dst[2][2] = {{2 * pr, 2 * pr + 1}, {2 * pr, 2 * pr + 1}};
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                if (ch >= nch) break;
                const GranInfo& gi = S.gi[gl][ch];
                const uint32_t packed = pre[r][ch];
                const int tsel = gi.bt != 2 ? 0 : (gi.mixed ? 2 : 1);
                const float gn = S.gain[S.eidx[par][gi_][ch][S.code[tsel][2 * pr]]];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int q = (int)(int16_t)(h ? (packed >> 16) : (packed & 0xFFFFu));
                    const int aq = q < 0 ? -q : q;
                    const float m = __fmul_rn(__ldg(&T->pow43[aq]), gn);
                    v[ch][h] = q < 0 ? -m : m;
                    if (tsel) dst[ch][h] = S.scat[tsel - 1][2 * pr + h];
                }
            }
            if (ga.ms) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float a = v[0][h], c = v[1][h];
                    v[0][h] = __fmul_rn(__fadd_rn(a, c), L3_INV_SQRT2);
                    v[1][h] = __fmul_rn(__fsub_rn(a, c), L3_INV_SQRT2);
                }
            }
            float* x0 = &S.ring[0][18 * slot][0];
            x0[dst[0][0]] = v[0][0];
            x0[dst[0][1]] = v[0][1];
            if (nch == 2) { float* x1 = &S.ring[1][18 * slot][0]; x1[dst[1][0]] = v[1][0]; x1[dst[1][1]] = v[1][1]; }
        }
        __syncthreads();
        if (b + 1 < n_batches) fetch(b + 1);

        // ================= P2: alias butterflies + IMDCT + window ==========================================================
        // Tasks = (granule-channel q, half, subband) for the ACTIVE subbands only, compacted so that warps stay full:
        // half 1 evaluates DCT-IV rows 9..17 = the 18 outputs of this granule (F), half 0 rows 0..8 = the 18 outputs the next
        // granule adds (S2).  Inactive subbands are zero-filled by all threads.
        {
            int lim[2 * K2_G], base[2 * K2_G + 1];
            base[0] = 0;
#pragma unroll
            for (int q = 0; q < 2 * K2_G; q++) {
                const int gi_ = q >> 1, ch = q & 1;
                lim[q] = (ch < nch && gl0 + gi_ < gcount) ? (int)S.gi[gl0 + gi_][ch].sb_limit : -1;   // -1: no such granule-channel
                base[q + 1] = base[q] + (lim[q] > 0 ? lim[q] : 0);
            }
            const int total = base[2 * K2_G];
#pragma unroll
            for (int q = 0; q < 2 * K2_G; q++) {
                if (lim[q] < 0 || lim[q] >= 32) continue;
                const int gi_ = q >> 1, ch = q & 1, nsb = 32 - lim[q];
                float* Fq = &S.F[gi_][ch][0][0];
                float* Sq = &S.S2[(gl0 + gi_ + 1) % K2_RING][ch][0][0];
                for (int idx = tid; idx < 18 * nsb; idx += K2_THREADS) {
                    const int i = idx / nsb, sb = lim[q] + (idx - i * nsb);
                    Fq[i * ROW + sb] = 0.0f;
                    Sq[i * ROW + sb] = 0.0f;
                }
            }
            for (int t = tid; t < 2 * total; t += K2_THREADS) {
                const int half = t >= total ? 0 : 1;       // F producers first
                const int rr = t >= total ? t - total : t;
                int q = 0;
#pragma unroll
                for (int k = 1; k < 2 * K2_G; k++) q += rr >= base[k] ? 1 : 0;
                int bq = 0;
#pragma unroll
                for (int k = 1; k < 2 * K2_G; k++) bq = (q >= k) ? base[k] : bq;
                const int sb = rr - bq;
                const int gi_ = q >> 1, ch = q & 1;
                const uint32_t gl = gl0 + gi_;
                const GranInfo& gi = S.gi[gl][ch];
                const int limit = (int)gi.sb_limit;
                const float* xr = &S.ring[ch][18 * (gl % K2_RING)][0];
                float X[18];
                const float2* X2 = reinterpret_cast<const float2*>(&xr[18 * sb]);
#pragma unroll
                for (int k = 0; k < 9; k++) { const float2 x = X2[k]; X[2 * k] = x.x; X[2 * k + 1] = x.y; }
                // alias reduction: boundary s (between subbands s-1 and s) is processed when s < limit and the block is long
                // (or s == 1 of a mixed block).  The butterflies of neighbouring boundaries touch disjoint lines, so both of
                // this subband's boundaries use the un-aliased neighbour lines still in xr.
                const bool lo_b = sb >= 1 && (gi.bt != 2 || (gi.mixed && sb == 1));
                const bool hi_b = sb + 1 < limit && (gi.bt != 2 || (gi.mixed && sb == 0));
                if (lo_b) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float lo = xr[18 * sb - 1 - i], hi = X[i];
                        X[i] = __fadd_rn(__fmul_rn(hi, L3_CS(i)), __fmul_rn(lo, L3_CA(i)));
                    }
                }
                if (hi_b) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float lo = X[17 - i], hi = xr[18 * (sb + 1) + i];
                        X[17 - i] = __fsub_rn(__fmul_rn(lo, L3_CS(i)), __fmul_rn(hi, L3_CA(i)));
                    }
                }
                const bool is_short = gi.bt == 2 && !(gi.mixed && sb < 2);
                float o[18];   // half 1: raw[0..17]; half 0: raw[18..35]
                if (!is_short) {
                    const int bt = (gi.bt == 2) ? 0 : gi.bt;  // long subbands of a mixed block use the normal window
                    float u[9];
#pragma unroll
                    for (int mm = 0; mm < 9; mm++) {
                        const float4* C4 = reinterpret_cast<const float4*>(S.dct4[9 * half + mm]);
                        float a = 0.0f;
#pragma unroll
                        for (int k4 = 0; k4 < 4; k4++) {
                            const float4 c = C4[k4];
                            a = fmaf(X[4 * k4], c.x, a); a = fmaf(X[4 * k4 + 1], c.y, a);
                            a = fmaf(X[4 * k4 + 2], c.z, a); a = fmaf(X[4 * k4 + 3], c.w, a);
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
                // frequency inversion is applied to both addends of the overlap-add: -(a+b) == (-a)+(-b)
                float* dstp = half ? &S.F[gi_][ch][0][sb] : &S.S2[(gl + 1) % K2_RING][ch][0][sb];
                const bool odd = sb & 1;
#pragma unroll
                for (int i = 0; i < 18; i++) dstp[i * ROW] = (odd && (i & 1)) ? -o[i] : o[i];
            }
        }
        __syncthreads();

        // ================= P3: overlap-add + 32-point DCT-II per time slot (Lee), two threads per transform ================
        // Warps 0..2 take the sum branch (even outputs), warps 3..5 the difference branch (odd outputs).
        {
            const int br = tid >= K2_THREADS / 2;
            const int task = br ? tid - K2_THREADS / 2 : tid;     // (granule, channel, slot)
            const int per_g = 18 * nch;
            const int gi_ = task / per_g, rem = task - gi_ * per_g;
            const int ch = rem / 18, t = rem - 18 * ch;
            const uint32_t gl = gl0 + gi_;
            if (gi_ < K2_G && gl < gcount) {
                const float* xa = S.F[gi_][ch][t];
                const float* xb = S.S2[gl % K2_RING][ch][t];
                float z[16];
                if (br == 0) {
#pragma unroll
                    for (int i = 0; i < 16; i++) z[i] = __fadd_rn(__fadd_rn(xa[i], xb[i]), __fadd_rn(xa[31 - i], xb[31 - i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        z[i] = __fmul_rn(__fsub_rn(__fadd_rn(xa[i], xb[i]), __fadd_rn(xa[31 - i], xb[31 - i])), c_dct_sec[i]);
                }
                Dct2<16, 16>::run(z);
                float* dstrow = S.ring[ch][18 * (gl % K2_RING) + t];
                if (br == 0) {
#pragma unroll
                    for (int k = 0; k < 16; k++) dstrow[2 * k] = z[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 16; k++) dstrow[2 * k + 1] = k < 15 ? __fadd_rn(z[k], z[k + 1]) : z[15];
                }
            }
        }
        __syncthreads();

        // ================= P4: polyphase window -> PCM; the two spare warps run P0 of the next batch ======================
        if (tid < K2_G * 64) {
            const int gi_ = tid >> 6, ch = (tid >> 5) & 1, j = lane;
            const uint32_t gl = gl0 + gi_;
            const uint32_t g = gfirst + gl;
            if (ch < nch && gl < gcount && g >= run.g0) {
                // out[t][j] = sum_p A[t-2p]*wa[p] + B[t-2p-1]*wb[p]  (p ascending, one fmaf chain); rows t-15..t-1 of slot 0
                // live in the previous granule's ring slot (rows 3..17)
                float wa[8], wb[8];
#pragma unroll
                for (int p = 0; p < 8; p++) { wa[p] = S.wa[p][j]; wb[p] = S.wb[p][j]; }
                const int ia = S.ia[j], ib = S.ib[j];
                const float* cur = S.ring[ch][18 * (gl % K2_RING)];
                const float* prv = S.ring[ch][18 * ((gl + K2_RING - 1) % K2_RING) + 3];
                float A[33], B[33];   // rows -15 .. 17 relative to this granule's slot 0
#pragma unroll
                for (int r = 0; r < 15; r++) { A[r] = prv[r * ROW + ia]; B[r] = prv[r * ROW + ib]; }
                float* dst = pcm + tr.pcm_off + (size_t)ch * tr.pcm_plane + (size_t)g * 576 + j;
#pragma unroll
                for (int t = 0; t < 18; t++) {
                    A[15 + t] = cur[t * ROW + ia];
                    B[15 + t] = cur[t * ROW + ib];
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
        } else if (b + 1 < n_batches) {
            for (int q = warp - K2_G * 2; q < K2_G * nch; q += K2_THREADS / 32 - K2_G * 2) p0_gc(b + 1, q);
        }
        __syncthreads();
    }
}
