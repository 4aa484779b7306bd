// k_hybrid_synth.cuh -- K2: requantise -> MS / intensity stereo -> reorder -> alias reduction -> IMDCT/overlap -> polyphase -> PCM
// (MPEG-1, 2 and 2.5 Layer III; or, in peak mode, max |sample| per track instead of PCM).
//
// One CTA (8 warps) owns a run of consecutive granules of one track and walks it in time order in BATCHES of 8
// granule-channels (4 stereo granules or 8 mono granules): one granule-channel per warp.  Every stage uses the mapping in
// which its natural index is the lane, so there is no index arithmetic in the inner loops:
//
//   TMA  the batch's int16 spectra (8 x 1152 B, contiguous in HBM) arrive in shared memory as ONE bulk copy
//        (cp.async.bulk + mbarrier), issued a batch ahead by one thread,
//        together with the gain of every scalefactor band that K1 derived from part 2 (8 x 256 B)
//   P1   requantise + MS / intensity + short-block reorder -> xr (slot A) the warps of one granule, lane = line pair;
//                                                                        handed to P2 by a 64-thread named barrier
//   P2   alias butterflies + IMDCT + window + frequency inversion        warp = granule-channel, lane = SUBBAND: the 18
//        -> F (outputs 0..17, slot A in place), S2 (outputs 18..35)      lines live in registers, the DCT-IV runs as a
//                                                                        9-point FFT whose twiddles are FFMA immediates
//   P3   x = F + S2(previous granule); 32-point DCT-II (Lee) in place    thread = (granule-channel, time slot)
//   P4   16-tap polyphase window over the DCT rows -> PCM                warp = granule-channel, lane = output column
//
// The overlap (S2) and the DCT rows travel between batches in slots indexed by granule number modulo (G+1), so nothing is
// copied.  Both are finite-memory: a run that does not start at the head of the track first replays the two granules
// before it (PCM writes suppressed) to rebuild that state exactly, so runs are independent and the corpus decodes in
// parallel.
//
// Arithmetic follows DESIGN.md "f32 decode contract" operation for operation (frequency inversion is applied to both
// addends instead of the sum: -(a+b) == (-a)+(-b) in IEEE arithmetic).
#pragma once

#define K2_THREADS 256
#define K2_GC 8            // granule-channels per batch, one per warp
#define K2_SLOTS 10        // slots of A / S2: (G+1) granules x channels = 5x2 (stereo) or 9x1 (mono)
#define K2_MAX_G 256       // granules per run, including the two replayed ones
#define ROW 34             // padded row of 32 values: 8-byte aligned for the float2 accesses of P3, whose threads (one row each, 136 bytes
                           // apart, slots 2448 bytes apart) then fall on distinct bank pairs; P2 and P4 access a row across the lanes
#define SLOTF (18 * ROW)   // floats per slot

struct __align__(2) GranInfo {
    uint16_t nz;          // lines decoded by K1 (0 = silent)
    uint16_t bvp;         // pairs of the big-values region: int16 pairs in the spectrum row; the lines from there to nz are count1 quads
    uint8_t bt, mixed, ms /* bit0 MS stereo, bit1 intensity stereo (frame level) */, sb_limit /* subbands with possibly non-zero IMDCT input */;
};

struct K2Smem {
    alignas(16) float A[K2_SLOTS][SLOTF];     // per granule-channel: xr[576] -> F[18][ROW] -> DCT rows [18][ROW], in place
    float S2[K2_SLOTS][SLOTF];                // IMDCT outputs 18..35 of the granule-channel (added by the NEXT granule)
    alignas(16) uint32_t stage[K2_GC][288];   // the batch's int16 line pairs, written by the TMA unit
    alignas(16) uint8_t c1q[K2_GC][K1_C1_ROW];   // the batch's count1 quads, one byte each (v w x y as 2-bit two's complement, v in bits 0-1)
    alignas(16) uint8_t banddelta[K2_GC][K1_GAIN_ROW];   // per band: what to subtract from the granule-channel's gain exponent
                                              // ([0..21] long sfb, [22..60] short sfb*3+win, [62..63] the exponent itself), by TMA
    float gaintab[L3_GAIN_MAX - L3_GAIN_MIN + 1];   // 2^(e/4)
    float bgain[K2_GC][64];                   // the batch's band gains, looked up from the bytes above at the start of P1
    float wa[8][32], wb[8][32];
    GranInfo gi[K2_MAX_G][2];                 // side info of every granule-channel of the run
    uint32_t p1w[K2_MAX_G][2];                // tsel | ms<<2 | (nz!=0)<<3 | intensity<<4 | (9*max sb_limit of the granule)<<8
    uint16_t scat[2][576];                    // bitstream-order line -> reordered line: [0] pure short, [1] mixed
    uint8_t code[3][576];                     // line -> band code: [0] long, [1] pure short, [2] mixed
    uint8_t ia[32], ib[32];
    uint8_t istab[K2_GC / 2][64];             // intensity stereo: is_pos per band of the right channel's partition, 255 = not intensity coded
    alignas(16) int istop[K2_GC / 2][4];      // highest non-zero band of the right channel: [0] long sfb, [1..3] short sfb per window
                                              // (16-byte aligned: ptxas reads a pair's four words with ONE LDS.128; unaligned, that load reached
                                              // into the neighbouring table and racecheck reported it against another pair's store)
    float isr[2][8];                          // MPEG-1: is_pos -> left / right factor
    uint8_t iscale[K2_GC / 2];                // MPEG-2: intensity_scale of the granule's right channel
    alignas(8) unsigned long long mbar;       // completion barrier of the spectra copy
};
static_assert(3 * (sizeof(K2Smem) + 1024) <= 228 * 1024, "three CTAs of K2 must fit one SM's shared memory");

// Compile-time loop: the body receives its index as an integral_constant, so table entries indexed with it are constant
// expressions and reach the SASS as instruction immediates (FFMA/FMUL imm32) -- no constant-bank or uniform-register
// traffic for the 324 matrix entries of the IMDCT, the window shapes or the DCT secants.
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// Lee's recursive DCT-II split on float2: two independent transforms of the same size side by side (the two halves of a
// 32-point transform after its first split use the same secants), every operation one packed instruction with the secant as
// the broadcast operand:  g[i] = x[i] + x[N-1-i],  h[i] = (x[i] - x[N-1-i]) * sec[i],  X[2k] = G[k],  X[2k+1] = H[k] + H[k+1].
template <int N, int OFF>
struct Dct2x2 {
    static __device__ __forceinline__ void run(float2 (&x)[N]) {
        float2 g[N / 2], h[N / 2];
        static_for<0, N / 2>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr float sec = L3_DCT_SEC[OFF + i];
            g[i] = k2_fadd2(x[i], x[N - 1 - i]);
            h[i] = k2_fmul2(k2_fsub2(x[i], x[N - 1 - i]), make_float2(sec, sec));
        });
        Dct2x2<N / 2, OFF + N / 2>::run(g);
        Dct2x2<N / 2, OFF + N / 2>::run(h);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            x[2 * k] = g[k];
            x[2 * k + 1] = (k + 1 < N / 2) ? k2_fadd2(h[k], h[k + 1]) : h[k];
        }
    }
};
template <int OFF>
struct Dct2x2<1, OFF> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};
// 32-point transform of one row: the first split in scalar form, then both 16-point halves as ONE packed transform
// (the halves never mix before the final interleave, so each half of a packed operation is the scalar recursion's operation)
__device__ __forceinline__ void k2_dct32(float (&x)[32]) {
    float2 z[16];
    static_for<0, 16>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr float sec = L3_DCT_SEC[i];
        z[i].x = __fadd_rn(x[i], x[31 - i]);
        z[i].y = __fmul_rn(__fsub_rn(x[i], x[31 - i]), sec);
    });
    Dct2x2<16, 16>::run(z);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        x[2 * k] = z[k].x;
        x[2 * k + 1] = (k + 1 < 16) ? __fadd_rn(z[k].y, z[k + 1].y) : z[k].y;
    }
}

// position of IMDCT-36 output i in the 18-point DCT-IV vector u: i<9: u[i+9]; 9..26: u[26-i]; 27..35: u[i-27]
__device__ __forceinline__ constexpr int idx36(int i) { return i < 9 ? i + 9 : (i < 27 ? 26 - i : i - 27); }

// the side-info fields K2 still needs of granule g (0 = first granule of the track), channel ch
__device__ __forceinline__ GranInfo k2_parse_gi(const uint8_t* __restrict__ corpus, const DecTrack& tr, const FrameRec* __restrict__ frames,
                                                const uint16_t* __restrict__ nz_in, uint32_t g, uint32_t ch, int nch) {
    const uint32_t gpf = tr.gpf, frame = g / gpf, gr = g % gpf;
    const FrameRec rec = frames[tr.frame_base + frame];
    const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
    // MPEG-1: 59 bits per granule-channel from bit 18 / 20, window switching at +33; MPEG-2 / 2.5: 63 bits from bit 9 / 10, at +38
    const int b0 = gpf == 1 ? (nch == 1 ? 9 : 10) + 63 * (int)ch + 5 : (nch == 1 ? 18 : 20) + 59 * (int)(gr * nch + ch);
    GranInfo gi;
    const uint32_t f = side_field(s, b0 + 33, 4);   // window_switching(1) block_type(2) mixed(1)
    gi.bt = (f & 8) ? (uint8_t)((f >> 1) & 3) : 0;
    gi.mixed = (gi.bt == 2) ? (uint8_t)(f & 1) : 0;
    const size_t gcb = (size_t)tr.gc_base + (size_t)g * nch;
    gi.nz = nz_in[gcb + ch];
    gi.bvp = (uint16_t)min(side_field(s, b0 + 12 - (gpf == 1 ? 5 : 0), 9), (uint32_t)gi.nz >> 1);   // (b0 carries the 5 extra bits of the MPEG-2 scalefac_compress)
    const bool joint = nch == 2 && ((rec.hdr >> 6) & 3) == 1;
    gi.ms = (uint8_t)((joint && ((rec.hdr >> 4) & 2) ? 1 : 0) | (joint && ((rec.hdr >> 4) & 1) ? 2 : 0));
    // subbands whose IMDCT input can be non-zero: the coded lines (of either channel under MS / intensity stereo), one more
    // subband for the alias butterflies; short blocks scatter lines across the whole band, so they keep all 32
    uint32_t nzm = gi.nz;
    if (gi.ms) nzm = max(nzm, (uint32_t)nz_in[gcb + (ch ^ 1)]);
    gi.sb_limit = nzm == 0 ? 0 : (gi.bt == 2 ? 32 : (uint8_t)min(32u, (nzm + 17) / 18 + 1));
    return gi;
}

// Complex numbers are float2 {re, im}; every operation below is one packed f32x2 instruction on both components.  Swapping the
// halves of an operand and negating one of them (multiplication by -i or +i) are operand modifiers of those instructions, not
// instructions of their own.
// 3-point DFT (W3 = e^{-2 pi i/3}): X0 = x0 + s, X1 = m - i q, X2 = m + i q with s = x1 + x2, m = x0 - s/2 (one fma per
// component), q = (sqrt3/2)(x1 - x2).  Operation for operation the f32 decode contract's dft3.
__device__ __forceinline__ void k2_dft3(float2 x0, float2 x1, float2 x2, float2& y0, float2& y1, float2& y2) {
    const float2 s = k2_fadd2(x1, x2), d = k2_fsub2(x1, x2);
    y0 = k2_fadd2(x0, s);
    const float2 m = k2_ffma2(make_float2(-0.5f, -0.5f), s, x0);
    const float2 q = k2_fmul2(make_float2(L3_SQRT3_2, L3_SQRT3_2), d);
    const float2 miq = make_float2(q.y, -q.x);   // -i q
    y1 = k2_fadd2(m, miq);                       // (mr + qi, mi - qr)
    y2 = k2_fsub2(m, miq);                       // (mr - qi, mi + qr)
}

// 18-point DCT-IV u[m] = sum_k X[k] cos(pi/72 (2m+1)(2k+1)) through a 9-point complex FFT (3 x 3 Cooley-Tukey), 84 packed
// instructions (172 in scalar form, 324 FFMAs in matrix form); all twiddles are instruction immediates:
//   v[n] = (X[2n] + i X[17-2n]) e^{-i pi (4n+1)/72},  V = FFT9(v),  y[k] = V[k] e^{-i pi k/18},  u[2k] = Re y[k], u[17-2k] = -Im y[k]
// A complex product (a + i b)(c + i d) is evaluated as (fmaf(-b, d, a*c), fmaf(a, d, b*c)): one packed multiplication by c and
// one packed fma of (-b, a) with d.
__device__ __forceinline__ void k2_dct4_18(const float (&X)[18], float (&u)[18]) {
    float2 v[9];
    static_for<0, 9>([&](auto nc) {
        constexpr int n = decltype(nc)::value;
        constexpr float c = L3_D4_PRE[0][n], s = L3_D4_PRE[1][n];
        const float a = X[2 * n], b = X[17 - 2 * n];
        v[n] = k2_ffma2(make_float2(b, -a), make_float2(s, s), k2_fmul2(make_float2(a, b), make_float2(c, c)));   // (a + i b)(c - i s)
    });
    float2 t[3][3];   // [b][c]
    static_for<0, 3>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        k2_dft3(v[b], v[3 + b], v[6 + b], t[b][0], t[b][1], t[b][2]);
    });
    static_for<0, 4>([&](auto ic) {
        constexpr int b = 1 + decltype(ic)::value / 2, c = 1 + decltype(ic)::value % 2;
        constexpr int e = b * c == 1 ? 0 : (b * c == 2 ? 1 : 2);   // W9^{bc}: exponents 1, 2, 4
        constexpr float wr = L3_D4_W9[0][e], wi = L3_D4_W9[1][e];
        const float2 z = t[b][c];
        t[b][c] = k2_ffma2(make_float2(-z.y, z.x), make_float2(wi, wi), k2_fmul2(z, make_float2(wr, wr)));
    });
    static_for<0, 3>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        float2 y[3];
        k2_dft3(t[0][c], t[1][c], t[2][c], y[0], y[1], y[2]);
        static_for<0, 3>([&](auto dc) {
            constexpr int d = decltype(dc)::value, k = c + 3 * d;
            float2 r = y[d];
            if constexpr (k != 0) {
                constexpr float pc = L3_D4_POST[0][k], ps = L3_D4_POST[1][k];
                r = k2_ffma2(make_float2(y[d].y, -y[d].x), make_float2(ps, ps), k2_fmul2(y[d], make_float2(pc, pc)));   // y (pc - i ps)
            }
            u[2 * k] = r.x;
            u[17 - 2 * k] = -r.y;
        });
    });
}

// long-block window: o[i] = u[idx36(i)] * win[BT][i]
template <int BT>
__device__ __forceinline__ void k2_window36(const float (&u)[18], float (&o)[36]) {
    static_for<0, 36>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr float w = L3_WINSGN[BT][i];
        o[i] = __fmul_rn(u[idx36(i)], w);
    });
}

// P4 of one granule-channel, lane = output column j:
//   out[t][j] = sum_p A[t-2p]*wa[p] + B[t-2p-1]*wb[p]   (p ascending, one fmaf chain)
// rows t-15..t-1 of slot 0 live in the previous granule's slot (rows 3..17)
// PEAK: no PCM is stored; the granule-channel's max |sample| is folded into *peak_bits (float bits, atomicMax) instead --
// the peak-only path of find_peak_amplitude (src/replaygain.rs:1140-1249).
template <bool CLAMP, bool PEAK>
__device__ __forceinline__ void k2_polyphase(const K2Smem& S, const float* __restrict__ cur, const float* __restrict__ prv,
                                             float* __restrict__ dst, uint32_t* __restrict__ peak_bits, int j) {
    float wa[8], wb[8];
#pragma unroll
    for (int p = 0; p < 8; p++) { wa[p] = S.wa[p][j]; wb[p] = S.wb[p][j]; }
    const int ia = S.ia[j], ib = S.ib[j];
    float Av[33], Bv[33];   // rows -15 .. 17 relative to this granule's slot 0
    float mx = 0.0f;
#pragma unroll
    for (int r = 0; r < 15; r++) { Av[r] = prv[r * ROW + ia]; Bv[r] = prv[r * ROW + ib]; }
    // two output rows (t, t + 1) per chain, as the two halves of packed f32x2 operations: each half is the scalar fmaf chain of
    // its row, p ascending (FFMA2 with the window weight as the broadcast operand: 144 instructions for the 288 multiply-adds)
#pragma unroll
    for (int t = 0; t < 18; t += 2) {
        Av[15 + t] = cur[t * ROW + ia];
        Bv[15 + t] = cur[t * ROW + ib];
        Av[16 + t] = cur[(t + 1) * ROW + ia];
        Bv[16 + t] = cur[(t + 1) * ROW + ib];
        float2 a = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int p = 0; p < 8; p++) {
            a = k2_ffma2(make_float2(Av[15 + t - 2 * p], Av[16 + t - 2 * p]), make_float2(wa[p], wa[p]), a);
            a = k2_ffma2(make_float2(Bv[14 + t - 2 * p], Bv[15 + t - 2 * p]), make_float2(wb[p], wb[p]), a);
        }
        if (CLAMP) { a.x = fminf(1.0f, fmaxf(-1.0f, a.x)); a.y = fminf(1.0f, fmaxf(-1.0f, a.y)); }
        if (PEAK) mx = fmaxf(mx, fmaxf(fabsf(a.x), fabsf(a.y)));
        else { dst[32 * t] = a.x; dst[32 * t + 32] = a.y; }
    }
    if (PEAK) {
        const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));   // non-negative floats order like their bit patterns
        if (j == 0) atomicMax(peak_bits, m);
    }
}

template <int NCH>
__device__ __forceinline__ void k2_body(K2Smem& S, const uint8_t* __restrict__ corpus,
                                        const FrameRec* __restrict__ frames, const DecTrack& tr, const RunDesc& run,
                                        const L3Dev* __restrict__ T, const int16_t* __restrict__ spectra,
                                        const uint8_t* __restrict__ c1q, const uint16_t* __restrict__ nz_in, const uint8_t* __restrict__ gains,
                                        const uint8_t* __restrict__ sfraw, float* __restrict__ pcm, uint32_t* __restrict__ peak_bits,
                                        uint32_t flags) {
    constexpr int G = K2_GC / NCH;      // granules per batch
    constexpr int NS = G + 1;           // granule slots
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sr = (int)tr.sr_idx;
    const uint32_t warm = run.g0 < 2 ? run.g0 : 2;
    const uint32_t gfirst = run.g0 - warm, gcount = run.ng + warm;   // gcount <= K2_MAX_G
    const uint32_t n_batches = (gcount + G - 1) / G;
    const int my_gi = warp / NCH, my_ch = warp % NCH;                // this warp's granule-channel within a batch

    // ---- one-time table staging and state reset
    for (int i = tid; i < 256; i += K2_THREADS) { (&S.wa[0][0])[i] = (&T->synth_wa[0][0])[i]; (&S.wb[0][0])[i] = (&T->synth_wb[0][0])[i]; }
    for (int i = tid; i < L3_GAIN_MAX - L3_GAIN_MIN + 1; i += K2_THREADS) S.gaintab[i] = T->gain[i];
    for (int i = tid; i < 576; i += K2_THREADS) {
        S.code[0][i] = T->line_sfb_long[sr][i];
        S.code[1][i] = 22 + T->line_sfb_short[sr][i];
        const uint8_t c = T->line_sfb_mixed[sr][i];
        S.code[2][i] = c < 32 ? c : 22 + (c - 32);
        S.scat[0][i] = T->scatter_short[sr][i];
        S.scat[1][i] = T->scatter_mixed[sr][i];
    }
    if (tid < 32) { S.ia[tid] = T->synth_ia[tid]; S.ib[tid] = T->synth_ib[tid]; }
    if (tid < 14) S.isr[tid / 7][tid % 7] = T->is_ratio[tid / 7][tid % 7];
    for (int i = tid; i < K2_SLOTS * SLOTF; i += K2_THREADS) { (&S.A[0][0])[i] = 0.0f; (&S.S2[0][0])[i] = 0.0f; }
    if (tid == 0) k2_mbar_init(&S.mbar, 1);
    // side info of every granule-channel of the run, one thread each
    for (uint32_t it = tid; it < gcount * NCH; it += K2_THREADS) {
        const uint32_t gl = it / NCH, ch = it % NCH;
        S.gi[gl][ch] = k2_parse_gi(corpus, tr, frames, nz_in, gfirst + gl, ch, NCH);
    }
    __syncthreads();
    for (uint32_t it = tid; it < gcount * NCH; it += K2_THREADS) {
        const uint32_t gl = it / NCH, ch = it % NCH;
        const GranInfo gi = S.gi[gl][ch];
        const int lim = max((int)gi.sb_limit, (int)S.gi[gl][NCH - 1 - ch].sb_limit);
        const uint32_t tsel = gi.bt != 2 ? 0u : (gi.mixed ? 2u : 1u);
        S.p1w[gl][ch] = tsel | ((uint32_t)(gi.ms & 1) << 2) | ((gi.nz != 0 ? 1u : 0u) << 3) | ((uint32_t)(gi.ms >> 1) << 4) | ((uint32_t)(9 * lim) << 8);
    }

    // long-block band of this lane's line pairs (P1), one byte per pass
    constexpr int NIT = (288 + 32 * NCH - 1) / (32 * NCH);
    unsigned long long codeL = 0;
    uint32_t codeH = 0;   // (mono: nine passes)
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int pr = lane + 32 * my_ch + 32 * NCH * it;
        const uint32_t cd = pr < 288 ? S.code[0][2 * pr] : 0u;
        if (it < 8) codeL |= (unsigned long long)cd << (8 * (it & 7));
        else codeH = cd;
    }

    // spectra and band gains of batch b: granule-channels are contiguous in HBM, 1152 + 256 bytes each
    auto issue_copy = [&](uint32_t b) {
        const uint32_t gl0 = b * G;
        const uint32_t ngc = min((uint32_t)G, gcount - gl0) * NCH;
        const size_t gc0 = (size_t)tr.gc_base + (size_t)(gfirst + gl0) * NCH;
        k2_mbar_expect(&S.mbar, ngc * (1152u + K1_GAIN_ROW + K1_C1_ROW));
        k2_bulk_load(&S.stage[0][0], spectra + gc0 * 576, ngc * 1152u, &S.mbar);
        k2_bulk_load(&S.c1q[0][0], c1q + gc0 * K1_C1_ROW, ngc * K1_C1_ROW, &S.mbar);
        k2_bulk_load(&S.banddelta[0][0], gains + gc0 * K1_GAIN_ROW, ngc * K1_GAIN_ROW, &S.mbar);
    };

    if (tid == 0) issue_copy(0);
    __syncthreads();

    for (uint32_t b = 0; b < n_batches; b++) {
        const uint32_t gl0 = b * G;
        const uint32_t gl = gl0 + my_gi;          // this warp's granule (relative to gfirst)
        const bool have = gl < gcount;
        const int slot = (int)(gl % NS) * NCH + my_ch;
        const int slot_prev = (int)((gl + NS - 1) % NS) * NCH + my_ch;

        // ================= P1: requantise (+ MS stereo, + short-block reorder folded into the store index) =================
        // The warps of one granule share its line pairs; scalefactor bands start on even lines, so a pair shares its band.
        k2_mbar_wait(&S.mbar, b & 1);
        if (have) {
            const uint32_t w0 = S.p1w[gl][0], w1 = S.p1w[gl][NCH - 1];
            const int lim9 = (int)(w0 >> 8);
            const int tsel0 = w0 & 3, tsel1 = w1 & 3;
            const bool ms = (w0 >> 2) & 1;
            // K1 writes int16 pairs up to big_values, one byte per count1 quad from there to its last decoded line (nz), and
            // nothing beyond: the rest reads as zero
            const int nzp0 = S.gi[gl][0].nz >> 1, nzp1 = S.gi[gl][NCH - 1].nz >> 1;
            const int bvp0 = S.gi[gl][0].bvp, bvp1 = S.gi[gl][NCH - 1].bvp;
            const uint8_t* cq0 = S.c1q[my_gi * NCH];
            const uint8_t* cq1 = S.c1q[my_gi * NCH + NCH - 1];
            const uint32_t* st0 = S.stage[my_gi * NCH];
            const uint32_t* st1 = S.stage[my_gi * NCH + NCH - 1];
            const float* bg0 = S.bgain[my_gi * NCH];
            const float* bg1 = S.bgain[my_gi * NCH + NCH - 1];
            {   // this warp's granule-channel: band byte -> band gain, 2^((exponent - byte) / 4)
                const uint8_t* bd = S.banddelta[my_gi * NCH + my_ch];
                const int be = *reinterpret_cast<const uint16_t*>(bd + 62);
                float* bg = S.bgain[my_gi * NCH + my_ch];
                bg[lane] = S.gaintab[be - (int)bd[lane]];
                bg[lane + 32] = S.gaintab[be - (int)bd[lane + 32]];   // (entries 61..63 are never read)
            }
            if (NCH == 2) k2_pair_sync(my_gi); else __syncwarp();
            float* x0 = S.A[(gl % NS) * NCH];
            float* x1 = S.A[(gl % NS) * NCH + NCH - 1];
            const bool ist = NCH == 2 && ((w0 >> 4) & 1);
            if (NCH == 2 && ist) {
                // intensity stereo (pair-uniform, rare): bands of the right channel's partition above its last non-zero line
                // (per window for short blocks) take L*is_l, L*is_r with is_pos = the right channel's raw scalefactor
                // (K1's sfraw row); is_pos 7 marks a band that is not intensity coded.  ISO 11172-3 2.4.3.4.9.3.
                int* top = S.istop[my_gi];
                if (my_ch == 0 && lane < 4) top[lane] = -1;
                k2_pair_sync(my_gi);
                for (int pr = lane + 32 * my_ch; pr < 288; pr += 64) {
                    if (pr < nzp1 && (pr < bvp1 ? st1[pr] != 0u : ((cq1[(pr - bvp1) >> 1] >> (4 * ((pr - bvp1) & 1))) & 15u) != 0u)) {
                        const int code = S.code[tsel1][2 * pr];
                        if (code < 22) atomicMax(&top[0], code);
                        else atomicMax(&top[1 + (code - 22) % 3], (code - 22) / 3);
                    }
                }
                k2_pair_sync(my_gi);
                const int j = lane + 32 * my_ch;
                if (j < 61) {
                    const uint8_t* row = sfraw + ((size_t)tr.gc_base + (size_t)(gfirst + gl) * 2 + 1) * K1_GAIN_ROW;
                    uint32_t pos = 255;   // K1 wrote 255 for "not intensity coded" (MPEG-1: is_pos >= 7, MPEG-2: the all-ones position)
                    if (j < 22) {
                        const bool short_nz = top[1] >= 0 || top[2] >= 0 || top[3] >= 0;
                        if (tsel1 != 1 && !short_nz && j > top[0]) pos = row[j == 21 ? 20 : j];
                    } else if (tsel1 != 0) {
                        const int sfb = (j - 22) / 3, win = (j - 22) % 3;
                        if (sfb > top[1 + win]) pos = row[22 + (sfb == 12 ? 11 : sfb) * 3 + win];
                    }
                    S.istab[my_gi][j] = (uint8_t)pos;
                    if (j == 0) S.iscale[my_gi] = row[63];
                }
                k2_pair_sync(my_gi);
            }
            // Two passes over the lane's (at most NIT) line pairs, both fully unrolled: every load first, then the arithmetic, so
            // that the dependent shared-memory lookups (line -> band -> gain) of all pairs are in flight together.  PLAIN = both
            // channels are long blocks and the frame has no intensity stereo (97 % of all granules): the band of a pair is then
            // a per-lane constant (codeL, set up once per CTA) and the stores are in line order.
            auto pairs = [&](auto plain_c) {
                constexpr bool PLAIN = decltype(plain_c)::value;
                uint32_t pk[NIT][2];
                float gq[NIT][2];
#pragma unroll
                for (int it = 0; it < NIT; it++) {
                    const int pr = lane + 32 * my_ch + 32 * NCH * it;
                    const uint32_t codeP = it < 8 ? (uint32_t)(codeL >> (8 * (it & 7))) & 0xFFu : codeH;
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) {
                        const int nzp = ch ? nzp1 : nzp0, bvp = ch ? bvp1 : bvp0;
                        uint32_t packed = 0u;
                        float gn = 0.0f;
                        if (pr < nzp) {   // (nz never reaches beyond the granule's last subband: nzp <= lim9)
                            gn = (ch ? bg1 : bg0)[PLAIN ? codeP : (uint32_t)S.code[ch ? tsel1 : tsel0][2 * pr]];
                            if (pr < bvp) packed = (ch ? st1 : st0)[pr];
                            else {
                                // count1 pair: K1 left one byte per quad, v w x y as 2-bit two's complement (0, +1, -1) from bit 0 up
                                const int t = pr - bvp;
                                const uint32_t nib = ((ch ? cq1 : cq0)[t >> 1] >> (4 * (t & 1))) & 15u;
                                const uint32_t c0 = nib & 3u, c1 = nib >> 2;
                                packed = (((c0 ^ 2u) - 2u) & 0xFFFFu) | (((c1 ^ 2u) - 2u) << 16);
                            }
                        }
                        pk[it][ch] = packed;
                        gq[it][ch] = gn;
                    }
                }
#pragma unroll
                for (int it = 0; it < NIT; it++) {
                    const int pr = lane + 32 * my_ch + 32 * NCH * it;
                    if (pr >= lim9) break;
                    // the two lines of a pair are the halves of packed f32x2 operations; the sign goes onto the table value
                    // before the multiplication ((-p) * g == -(p * g) in IEEE arithmetic)
                    float2 v[2];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) {
                        const uint32_t packed = pk[it][ch];
                        const float gn = gq[it][ch];
                        const int q0 = (int)(int16_t)(packed & 0xFFFFu), q1 = (int)packed >> 16;
                        const float p0 = __uint_as_float(__float_as_uint(__ldg(&T->pow43[abs(q0)])) | ((packed << 16) & 0x80000000u));
                        const float p1 = __uint_as_float(__float_as_uint(__ldg(&T->pow43[abs(q1)])) | (packed & 0x80000000u));
                        v[ch] = k2_fmul2(make_float2(p0, p1), make_float2(gn, gn));
                    }
                    uint32_t ipos = 255;
                    if (!PLAIN && NCH == 2 && ist) ipos = S.istab[my_gi][S.code[tsel1][2 * pr]];
                    if (!PLAIN && NCH == 2 && ipos != 255) {
                        const float kl = tr.gpf == 1 ? __ldg(&T->is_lsf[S.iscale[my_gi]][0][ipos & 31]) : S.isr[0][ipos & 7];   // (MPEG-2 intensity stereo is rare: its table stays in global memory)
                        const float kr = tr.gpf == 1 ? __ldg(&T->is_lsf[S.iscale[my_gi]][1][ipos & 31]) : S.isr[1][ipos & 7];
                        const float2 l = v[0];
                        v[0] = k2_fmul2(l, make_float2(kl, kl));
                        v[NCH - 1] = k2_fmul2(l, make_float2(kr, kr));
                    } else if (NCH == 2 && ms) {
                        const float2 l = v[0], r = v[NCH - 1];
                        v[0] = k2_fmul2(k2_fadd2(l, r), make_float2(L3_INV_SQRT2, L3_INV_SQRT2));
                        v[NCH - 1] = k2_fmul2(k2_fsub2(l, r), make_float2(L3_INV_SQRT2, L3_INV_SQRT2));
                    }
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) {
                        float* x = ch ? x1 : x0;
                        const int tsel = ch ? tsel1 : tsel0;
                        if (PLAIN || tsel == 0) *reinterpret_cast<float2*>(x + 2 * pr) = v[ch];
                        else { x[S.scat[tsel - 1][2 * pr]] = v[ch].x; x[S.scat[tsel - 1][2 * pr + 1]] = v[ch].y; }
                    }
                }
            };
            if (!ist && tsel0 == 0 && tsel1 == 0) pairs(std::true_type{});
            else pairs(std::false_type{});
        }
        if (NCH == 2) k2_pair_sync(my_gi); else __syncwarp();

        // ================= P2: alias butterflies + IMDCT + window, lane = subband ==========================================
        if (have) {
            const GranInfo gi = S.gi[gl][my_ch];
            float* Aq = S.A[slot];
            float* Sq = S.S2[slot];
            const int sb = lane, limit = (int)gi.sb_limit;
            if (limit == 0) {
#pragma unroll
                for (int i = 0; i < 18; i++) { Aq[i * ROW + sb] = 0.0f; Sq[i * ROW + sb] = 0.0f; }
            } else {
                float X[18], lo[8], hi[8];
                const bool act = sb < limit;
                {
                    const float2* X2 = reinterpret_cast<const float2*>(Aq + 18 * sb);
#pragma unroll
                    for (int k = 0; k < 9; k++) { const float2 x = X2[k]; X[2 * k] = act ? x.x : 0.0f; X[2 * k + 1] = act ? x.y : 0.0f; }
                    const float* pl = Aq + 18 * max(sb, 1) - 1;
                    const float* ph = Aq + 18 * min(sb + 1, 31);
#pragma unroll
                    for (int i = 0; i < 8; i++) { lo[i] = pl[-i]; hi[i] = ph[i]; }
                }
                __syncwarp();   // every lane holds its inputs: the slot may now be overwritten in place
                // alias reduction: boundary s (between subbands s-1 and s) is processed when s < limit and the block is long
                // (or s == 1 of a mixed block).  The butterflies of neighbouring boundaries touch disjoint lines, so both of
                // this subband's boundaries use the un-aliased neighbour lines loaded above.
                const bool lo_b = act && sb >= 1 && (gi.bt != 2 || (gi.mixed && sb == 1));
                const bool hi_b = sb + 1 < limit && (gi.bt != 2 || (gi.mixed && sb == 0));
                if (lo_b) {
                    static_for<0, 8>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        constexpr float cs = L3_ALIAS_CS[i], ca = L3_ALIAS_CA[i];
                        X[i] = __fadd_rn(__fmul_rn(X[i], cs), __fmul_rn(lo[i], ca));
                    });
                }
                if (hi_b) {
                    static_for<0, 8>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        constexpr float cs = L3_ALIAS_CS[i], ca = L3_ALIAS_CA[i];
                        X[17 - i] = __fsub_rn(__fmul_rn(X[17 - i], cs), __fmul_rn(hi[i], ca));
                    });
                }
                const bool is_short = gi.bt == 2 && !(gi.mixed && sb < 2);
                float o[36];
                if (!is_short) {
                    float u[18];
                    k2_dct4_18(X, u);
                    const int bt = (gi.bt == 2) ? 0 : gi.bt;  // long subbands of a mixed block use the normal window
                    if (bt == 0) k2_window36<0>(u, o);
                    else if (bt == 1) k2_window36<1>(u, o);
                    else k2_window36<3>(u, o);
                } else {
                    // three 12-point IMDCTs y_w (w = 0,1,2) placed at raw[6+6w ..]; overlaps added in ascending window order
                    float y[3][12];
                    static_for<0, 36>([&](auto wc) {
                        constexpr int w = decltype(wc)::value / 12, i = decltype(wc)::value % 12;
                        constexpr float win = L3_WINSGN[2][i];
                        float a = 0.0f;
                        static_for<0, 6>([&](auto kc) {
                            constexpr int k = decltype(kc)::value;
                            constexpr float c = L3_IMDCT12[i][k];
                            a = fmaf(X[3 * k + w], c, a);
                        });
                        y[w][i] = __fmul_rn(a, win);
                    });
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        o[i] = 0.0f;
                        o[6 + i] = y[0][i];
                        o[12 + i] = __fadd_rn(y[0][6 + i], y[1][i]);
                        o[18 + i] = __fadd_rn(y[1][6 + i], y[2][i]);
                        o[24 + i] = y[2][6 + i];
                        o[30 + i] = 0.0f;
                    }
                }
                // frequency inversion (odd samples of odd subbands), applied to both addends of the overlap-add
                const uint32_t sgn = (sb & 1) ? 0x80000000u : 0u;
#pragma unroll
                for (int i = 0; i < 18; i++) {
                    const float f = (i & 1) ? __uint_as_float(__float_as_uint(o[i]) ^ sgn) : o[i];
                    const float s = (i & 1) ? __uint_as_float(__float_as_uint(o[18 + i]) ^ sgn) : o[18 + i];
                    Aq[i * ROW + sb] = f;
                    Sq[i * ROW + sb] = s;
                }
            }
        }
        __syncthreads();

        // ================= P3: overlap-add + 32-point DCT-II per time slot (Lee), in place ================================
        // Threads 0..143 take one (granule-channel, slot) each; one thread of an idle warp starts the next batch's copy.
        if (tid < 18 * K2_GC) {
            const int q = tid / 18, t = tid - 18 * q;
            const uint32_t g3 = gl0 + q / NCH;
            if (g3 < gcount) {
                float* a = S.A[(g3 % NS) * NCH + q % NCH] + t * ROW;
                const float* s = S.S2[((g3 + NS - 1) % NS) * NCH + q % NCH] + t * ROW;
                float2* a2 = reinterpret_cast<float2*>(a);
                const float2* s2 = reinterpret_cast<const float2*>(s);
                float x[32];
#pragma unroll
                for (int i = 0; i < 16; i++) { const float2 v = k2_fadd2(a2[i], s2[i]); x[2 * i] = v.x; x[2 * i + 1] = v.y; }
                k2_dct32(x);
#pragma unroll
                for (int i = 0; i < 16; i++) a2[i] = make_float2(x[2 * i], x[2 * i + 1]);
            }
        } else if (tid == 5 * 32 && b + 1 < n_batches) {
            issue_copy(b + 1);   // P1 of this batch has consumed the staging buffers (barrier above)
        }
        __syncthreads();

        // ================= P4: polyphase window -> PCM, lane = output column ==============================================
        if (have && gfirst + gl >= run.g0) {
            const uint32_t g = gfirst + gl;
            if (peak_bits) {   // peak-only path: nothing is stored
                if (flags & 0x4u) k2_polyphase<true, true>(S, S.A[slot], S.A[slot_prev] + 3 * ROW, nullptr, peak_bits + run.track, lane);
                else k2_polyphase<false, true>(S, S.A[slot], S.A[slot_prev] + 3 * ROW, nullptr, peak_bits + run.track, lane);
            } else {
                float* dst = pcm + tr.pcm_off + (size_t)my_ch * tr.pcm_plane + (size_t)g * 576 + lane;
                if (flags & 0x4u) k2_polyphase<true, false>(S, S.A[slot], S.A[slot_prev] + 3 * ROW, dst, nullptr, lane);
                else k2_polyphase<false, false>(S, S.A[slot], S.A[slot_prev] + 3 * ROW, dst, nullptr, lane);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(K2_THREADS, 3)
k_hybrid_synth(const uint8_t* __restrict__ corpus, const FrameRec* __restrict__ frames,
               const DecTrack* __restrict__ tracks, const RunDesc* __restrict__ runs, const L3Dev* __restrict__ T,
               const int16_t* __restrict__ spectra, const uint8_t* __restrict__ c1q, const uint16_t* __restrict__ nz_in,
               const uint8_t* __restrict__ gains,
               const uint8_t* __restrict__ sfraw, float* __restrict__ pcm, uint32_t* __restrict__ peak_bits /* [track] or null */,
               uint32_t flags) {
    extern __shared__ __align__(16) unsigned char k2_raw[];
    K2Smem& S = *reinterpret_cast<K2Smem*>(k2_raw);
    const RunDesc run = runs[blockIdx.x];
    const DecTrack tr = tracks[run.track];
    if (tr.channels == 1) k2_body<1>(S, corpus, frames, tr, run, T, spectra, c1q, nz_in, gains, sfraw, pcm, peak_bits, flags);
    else k2_body<2>(S, corpus, frames, tr, run, T, spectra, c1q, nz_in, gains, sfraw, pcm, peak_bits, flags);
}
