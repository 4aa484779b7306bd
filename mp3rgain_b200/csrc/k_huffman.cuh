// k_huffman.cuh -- K1: Layer III part 2 (scalefactors) and part 3 (Huffman) decode, one thread per granule-channel.
//
// The bit offset of every granule-channel is known without decoding anything: frame f's data starts md_start bytes into the
// track's compacted main-data stream (K0) and granule-channels follow each other at part2_3_length strides, so all
// granule-channels of all frames decode in parallel.
//
// Part 2: the thread reads its scalefactors (an irregular, serial bit-field walk that this mapping runs 32-wide), resolves
// scfsi against granule 0's thread in the same warp, and writes the GAIN of every scalefactor band -- 2^((global_gain-210
// - ...)/4) looked up in the frozen table -- as 64 floats per granule-channel, so that K2 receives them with its bulk copy.
//
// Part 3: all lanes of a warp advance through the 288 line pairs in lockstep.  The bit reader is a 64-bit accumulator with
// a one-word look-ahead: one refill + one table lookup per code word, and the linbits / sign bits of a pair (or the four
// sign bits of a count1 quad) are peeled off a single 32-bit peek without branches.  Each lane parks its pair in a 16-pair
// shared-memory row and every 16 pairs the warp flushes 32 rows with coalesced 64-byte segments, so the int16 spectra
// reach HBM as full sectors although every lane owns a different 1152-byte row.
#pragma once

#define K1_THREADS 256
#define K1_GAIN_ROW 64   // floats per granule-channel: [0..21] long sfb, [22..60] short sfb*3+win

struct BitBuf {
    const uint4* q;      // main-data stream as 16-byte quads (16-byte aligned base)
    uint32_t lim;        // bytes readable for this frame (later bytes read as zero, like the oracle's reservoir)
    uint32_t wi;         // index of the next word that will enter acc
    uint32_t c0, c1, c2, c3;   // words wi, wi+1, ... still to enter acc (raw, little-endian as loaded); c0 is next
    uint4 nx;            // the quad after those: loaded a whole quad ahead, so its latency hides behind ~100 decoded bits
    unsigned long long acc;   // unread bits, left-aligned
    int nb;              // valid bits in acc (>= 33 after refill())
    // word `i` of the stream, big-endian, with the bytes at or beyond lim zeroed
    __device__ __forceinline__ uint32_t fix(uint32_t raw, uint32_t i) const {
        uint32_t v = __byte_perm(raw, 0, 0x0123);
        if (i >= (lim >> 2)) {   // the word reaches past the readable bytes (rare: only at the very end of a frame's data)
            const uint32_t b = i * 4;
            v = b >= lim ? 0u : (v & (0xFFFFFFFFu << (8 * (b + 4 - lim))));
        }
        return v;
    }
    __device__ __forceinline__ uint32_t take() {   // next word; keeps the quad buffers rolling
        const uint32_t v = fix(c0, wi);
        wi++;
        c0 = c1; c1 = c2; c2 = c3;
        if ((wi & 3) == 0) {
            c0 = nx.x; c1 = nx.y; c2 = nx.z; c3 = nx.w;
            nx = __ldg(q + (wi >> 2) + 1);
        }
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t* base, uint32_t lim_bytes, uint32_t bitpos) {
        q = reinterpret_cast<const uint4*>(base);
        lim = lim_bytes;
        wi = bitpos >> 5;
        const uint4 cur = __ldg(q + (wi >> 2));
        nx = __ldg(q + (wi >> 2) + 1);
        c0 = cur.x; c1 = cur.y; c2 = cur.z; c3 = cur.w;
        for (uint32_t k = 0; k < (wi & 3); k++) { c0 = c1; c1 = c2; c2 = c3; }
        const uint32_t sh = bitpos & 31;
        const uint32_t hi = take(), lo = take();
        acc = ((((unsigned long long)hi) << 32) | lo) << sh;
        nb = 64 - (int)sh;
    }
    // bit position of the next unread bit: words 0..wi-1 have entered acc, nb of their bits are still unread
    __device__ __forceinline__ uint32_t pos() const { return (wi << 5) - (uint32_t)nb; }
    __device__ __forceinline__ void refill() {
        if (nb <= 32) {
            acc |= (unsigned long long)take() << (32 - nb);
            nb += 32;
        }
    }
    __device__ __forceinline__ uint32_t top() const { return (uint32_t)(acc >> 32); }
    __device__ __forceinline__ void skip(int n) { acc <<= n; nb -= n; }
    __device__ __forceinline__ uint32_t get(int n) {   // n <= 16
        refill();
        const uint32_t v = (top() >> 1) >> (31 - n);
        skip(n);
        return v;
    }
};

__device__ __forceinline__ uint32_t side_field(const uint8_t* __restrict__ s, int bit, int n) {
    // n <= 12 bits starting at `bit` of the side-info bytes (big-endian bit order)
    const int by = bit >> 3;
    const uint32_t w = ((uint32_t)__ldg(s + by) << 16) | ((uint32_t)__ldg(s + by + 1) << 8) | __ldg(s + by + 2);
    return (w >> (24 - n - (bit & 7))) & ((1u << n) - 1);
}

__global__ void __launch_bounds__(K1_THREADS, 4)
k_huffman(const uint8_t* __restrict__ corpus, const uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
          const DecTrack* __restrict__ tracks, const HuffBlock* __restrict__ blocks, const L3Dev* __restrict__ T,
          int16_t* __restrict__ spectra /* [gc][576] */, uint16_t* __restrict__ nz_out /* [gc] */,
          float* __restrict__ gains /* [gc][K1_GAIN_ROW] */,
          uint8_t* __restrict__ sfraw /* [gc][K1_GAIN_ROW]: raw scalefactors = is_pos, right channel of intensity-stereo frames only */) {
    __shared__ uint16_t s_huff[8192];
    __shared__ uint32_t s_stage[K1_THREADS / 32][32][17];
    __shared__ uint8_t s_sfl[K1_THREADS][24];     // long scalefactors of every thread's granule-channel (scfsi source)
    __shared__ uint8_t s_c1[2][64];               // count1 tables A / B: 6-bit peek -> (len<<4)|quad
    __shared__ uint32_t s_cnt[512];               // counting sort of the block's granule-channels by big_values: bin sizes, then bin starts
    __shared__ uint32_t s_wsum[K1_THREADS / 32];
    __shared__ uint8_t s_perm[K1_THREADS];        // worker thread -> granule-channel of the block, ordered by big_values
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(T->huff);
        uint4* dst = reinterpret_cast<uint4*>(s_huff);
        for (int i = tid; i < 8192 * 2 / 16; i += K1_THREADS) dst[i] = src[i];
        if (tid < 64) {
            s_c1[0][tid] = T->c1a_lut[tid];
            s_c1[1][tid] = (uint8_t)((4u << 4) | (15u - (tid >> 2)));
        }
    }
    const HuffBlock blk = blocks[blockIdx.x];
    const DecTrack tr = tracks[blk.track];
    const uint32_t nch = tr.channels, gpf = tr.gpf;
    const bool lsf = gpf == 1;   // MPEG-2 / 2.5: one granule per frame, 63-bit side info per granule-channel (ISO 13818-3 2.4.1.7)
    const uint32_t n_gc = tr.n_frames * gpf * nch;
    // ---- order the block's granule-channels by big_values, so that the lanes of a warp leave the big-values decoder
    // (and enter the zero tail) at about the same pair: the lockstep loop below pays for the slowest lane of the warp
    {
        const uint32_t g = blk.gc0 + tid;
        uint32_t key = 0x1FF;   // inactive slots sort last
        if (g < n_gc) {
            const uint32_t frame = g / (gpf * nch), rem = g % (gpf * nch);
            const FrameRec rec = frames[tr.frame_base + frame];
            const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
            key = side_field(s, (lsf ? (nch == 1 ? 9 : 10) + 63 * rem : (nch == 1 ? 18 : 20) + 59 * rem) + 12, 9);
        }
        // counting sort over the 9-bit key: my slot inside my bin, then an exclusive prefix over the 512 bins (two per thread)
        if (tid < 256) { s_cnt[tid] = 0; s_cnt[tid + 256] = 0; }
        __syncthreads();
        const uint32_t slot = atomicAdd(&s_cnt[key], 1u);
        __syncthreads();
        const uint32_t c0 = s_cnt[2 * tid], c1 = s_cnt[2 * tid + 1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += s_wsum[w];
        s_cnt[2 * tid] = base + incl - c0 - c1;        // bins 2*tid, 2*tid+1: where their first member goes
        s_cnt[2 * tid + 1] = base + incl - c1;
        __syncthreads();
        s_perm[s_cnt[key] + slot] = (uint8_t)tid;
    }
    __syncthreads();
    const int gl = s_perm[tid];                   // my granule-channel within the block
    const uint32_t gc = blk.gc0 + gl;
    const bool active = gc < n_gc;

    // ---- side info of my granule-channel
    int bv = 0, r0 = 0, r1 = 0, c1tab = 0;
    uint32_t reg0 = 0, reg1 = 0, reg2 = 0;   // per region: first-level table base << 8 | linbits
    uint32_t end = 0;
    bool silent = true;
    BitBuf br;
    br.q = reinterpret_cast<const uint4*>(md + tr.md_off);
    br.lim = 0; br.acc = 0; br.wi = 2; br.nb = 64;
    br.c0 = br.c1 = br.c2 = br.c3 = 0; br.nx = make_uint4(0, 0, 0, 0);
    float* grow = gains + ((size_t)tr.gc_base + gc) * K1_GAIN_ROW;
    uint8_t* srow = sfraw + ((size_t)tr.gc_base + gc) * K1_GAIN_ROW;
    bool is_right = false;   // right channel of an intensity-stereo frame: K2 needs its raw scalefactors (is_pos)
    uint8_t* sfl = s_sfl[gl];
#pragma unroll
    for (int i = 0; i < 6; i++) reinterpret_cast<uint32_t*>(sfl)[i] = 0;
    // part-2 parameters, kept for the gain pass below
    int bt = 0, mixed = 0, slen1 = 0, slen2 = 0, scfsi = 0, base_e = 0, sf_shift = 2, preflag = 0, sbg0 = 0, sbg1 = 0, sbg2 = 0;
    uint32_t gr = 0;
    if (active) {
        const uint32_t frame = gc / (gpf * nch), rem = gc % (gpf * nch), ch = rem % nch;
        gr = rem / nch;
        const FrameRec rec = frames[tr.frame_base + frame];
        const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
        is_right = ch == 1 && ((rec.hdr >> 6) & 3) == 1 && ((rec.hdr >> 4) & 1);
        // field offsets inside one granule-channel: MPEG-1 has a 4-bit scalefac_compress and a preflag bit (59 bits in all),
        // MPEG-2 a 9-bit scalefac_compress and no preflag bit (63 bits)
        const int gbase = lsf ? (nch == 1 ? 9 : 10) : (nch == 1 ? 18 : 20), gstride = lsf ? 63 : 59, x5 = lsf ? 5 : 0;
        uint32_t bit = rec.md_start * 8;
        for (uint32_t k = 0; k < rem; k++) bit += side_field(s, gbase + gstride * k, 12);
        const int b0 = gbase + gstride * rem;
        const uint32_t p23 = side_field(s, b0, 12);
        bv = side_field(s, b0 + 12, 9);
        base_e = (int)side_field(s, b0 + 21, 8) - 210 - L3_GAIN_MIN;
        const int sfc = side_field(s, b0 + 29, lsf ? 9 : 4);
        const int ws = side_field(s, b0 + 33 + x5, 1);
        int sel0, sel1, sel2 = 0;
        if (ws) {
            bt = side_field(s, b0 + 34 + x5, 2);
            mixed = side_field(s, b0 + 36 + x5, 1);
            sel0 = side_field(s, b0 + 37 + x5, 5);
            sel1 = side_field(s, b0 + 42 + x5, 5);
            sbg0 = side_field(s, b0 + 47 + x5, 3); sbg1 = side_field(s, b0 + 50 + x5, 3); sbg2 = side_field(s, b0 + 53 + x5, 3);
            // region 0 = the first 8 long bands, or the first 3 short bands x 3 windows: 36 lines at every MPEG-1 rate,
            // 54 / 36 at the MPEG-2 rates, 108 / 72 at 8 kHz
            r0 = bt == 2 ? 3 * T->short_start[tr.sr_idx][3] : T->long_start[tr.sr_idx][8];
            r1 = 576;
        } else {
            sel0 = side_field(s, b0 + 34 + x5, 5);
            sel1 = side_field(s, b0 + 39 + x5, 5);
            sel2 = side_field(s, b0 + 44 + x5, 5);
            const int rc0 = side_field(s, b0 + 49 + x5, 4), rc1 = side_field(s, b0 + 53 + x5, 3);
            r0 = T->long_start[tr.sr_idx][rc0 + 1];
            const int k = rc0 + rc1 + 2;
            r1 = T->long_start[tr.sr_idx][k > 22 ? 22 : k];
        }
        if (bt != 2) mixed = 0;
        preflag = lsf ? (sfc >= 500 && !is_right) : (int)side_field(s, b0 + 56, 1);
        sf_shift = side_field(s, b0 + (lsf ? 61 : 57), 1) ? 4 : 2;
        c1tab = side_field(s, b0 + (lsf ? 62 : 58), 1);
        // slot 0 ("no data") reads the all-zero first-level table at huff_base[0]: length 0, x = y = 0
        reg0 = ((uint32_t)T->huff_base[T->tabsel[sel0][0]] << 8) | T->tabsel[sel0][1];
        reg1 = ((uint32_t)T->huff_base[T->tabsel[sel1][0]] << 8) | T->tabsel[sel1][1];
        reg2 = ((uint32_t)T->huff_base[T->tabsel[sel2][0]] << 8) | T->tabsel[sel2][1];
        // part-2 length
        int part2 = 0;
        int ls[4] = {0, 0, 0, 0}, ltbl = 0;   // MPEG-2: field widths and partition table from scalefac_compress
        const int cls = bt != 2 ? 0 : (mixed ? 2 : 1);
        if (lsf) {
            int c = sfc;
            if (is_right) {   // the scalefactors of an intensity-stereo right channel are is_pos values: other partition tables
                c >>= 1;
                if (c < 180) { ls[0] = c / 36; ls[1] = (c % 36) / 6; ls[2] = (c % 36) % 6; ltbl = 3; }
                else if (c < 244) { c -= 180; ls[0] = (c % 64) >> 4; ls[1] = (c % 16) >> 2; ls[2] = c % 4; ltbl = 4; }
                else { c -= 244; ls[0] = c / 3; ls[1] = c % 3; ltbl = 5; }
            } else if (c < 400) { ls[0] = (c >> 4) / 5; ls[1] = (c >> 4) % 5; ls[2] = (c % 16) >> 2; ls[3] = c % 4; ltbl = 0; }
            else if (c < 500) { c -= 400; ls[0] = (c >> 2) / 5; ls[1] = (c >> 2) % 5; ls[2] = c % 4; ltbl = 1; }
            else { c -= 500; ls[0] = c / 3; ls[1] = c % 3; ltbl = 2; }
            for (int i = 0; i < 4; i++) part2 += T->lsf_nsf[ltbl][cls][i] * ls[i];
        } else {
            slen1 = T->slen[0][sfc]; slen2 = T->slen[1][sfc];
            if (bt == 2) part2 = (mixed ? 17 : 18) * slen1 + 18 * slen2;
            else {
                scfsi = (gr == 1) ? (int)side_field(s, (nch == 1 ? 14 : 12) + 4 * ch, 4) : 0;  // band group 0 first
                part2 += (scfsi & 8) ? 0 : 6 * slen1;
                part2 += (scfsi & 4) ? 0 : 5 * slen1;
                part2 += (scfsi & 2) ? 0 : 5 * slen2;
                part2 += (scfsi & 1) ? 0 : 5 * slen2;
            }
        }
        if (is_right) srow[63] = (uint8_t)(lsf ? (sfc & 1) : 0);   // MPEG-2: intensity_scale
        end = bit + p23;
        silent = (uint32_t)part2 > p23;
        br.init(md + tr.md_off, rec.md_end, bit);

        // ---- part 2 (ISO 11172-3 2.4.2.7): my own scalefactors.  Long ones go to shared memory (granule 1 may need them),
        // short ones straight into their band gain.
        if (lsf) {
            // ISO 13818-3 2.4.3.2: the values come in band order, lsf_nsf[..][i] of them ls[i] bits wide: long sfb 0..20, or
            // (mixed: 6 long bands, then) short sfb x 3 windows
            int k = 0, fld = 0, left = T->lsf_nsf[ltbl][cls][0];
            const int n_long = cls == 0 ? 21 : (cls == 2 ? 6 : 0);
            const int n_all = n_long + (cls == 0 ? 0 : (cls == 2 ? 27 : 36));
            for (; k < n_all; k++) {
                while (left == 0 && fld < 3) left = T->lsf_nsf[ltbl][cls][++fld];
                const int v = left ? (int)br.get(ls[fld]) : 0;
                if (left) left--;
                // intensity position for K2: 255 = the illegal (all-ones) position of a non-empty field -> band not intensity coded
                const uint8_t ipos = (ls[fld] > 0 && v == (1 << ls[fld]) - 1) ? 255 : (uint8_t)v;
                if (k < n_long) { sfl[k] = (uint8_t)v; if (is_right) srow[k] = ipos; }
                else {
                    const int q2 = k - n_long + (cls == 2 ? 9 : 0), win = q2 % 3;
                    if (is_right) srow[22 + q2] = ipos;
                    grow[22 + q2] = __ldg(&T->gain[base_e - 8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2)) - sf_shift * v]);
                }
            }
            if (cls)
#pragma unroll
                for (int win = 0; win < 3; win++) grow[22 + 36 + win] = __ldg(&T->gain[base_e - 8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2))]);
        } else if (bt == 2) {
            if (mixed) for (int sfb = 0; sfb < 8; sfb++) sfl[sfb] = (uint8_t)br.get(slen1);
            for (int q2 = mixed ? 9 : 0; q2 < 36; q2++) {
                const int v = (int)br.get(q2 < 18 ? slen1 : slen2);
                const int win = q2 % 3;
                if (is_right) srow[22 + q2] = (uint8_t)(v < 7 ? v : 255);
                grow[22 + q2] = __ldg(&T->gain[base_e - 8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2)) - sf_shift * v]);
            }
#pragma unroll
            for (int win = 0; win < 3; win++)   // sfb 12 carries no scalefactor; a mixed block has no short sfb 0..2
                grow[22 + 36 + win] = __ldg(&T->gain[base_e - 8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2))]);
        } else {
            for (int sfb = 0; sfb < 21; sfb++) {
                const int grp = sfb < 6 ? 0 : (sfb < 11 ? 1 : (sfb < 16 ? 2 : 3));
                if (!(scfsi & (8 >> grp))) sfl[sfb] = (uint8_t)br.get(grp < 2 ? slen1 : slen2);
            }
        }
    }
    __syncthreads();
    if (active) {
        // scfsi: granule 1 reuses granule 0's scalefactors of the same channel (granule-channel gl - nch: same frame, same block).
        // A short granule 0 has no long scalefactors (zero) except the 8 of a mixed block -- exactly what its row holds.
        if (bt != 2 && scfsi) {
            const uint8_t* g0 = s_sfl[gl - nch];
            for (int sfb = 0; sfb < 21; sfb++) {
                const int grp = sfb < 6 ? 0 : (sfb < 11 ? 1 : (sfb < 16 ? 2 : 3));
                if (scfsi & (8 >> grp)) sfl[sfb] = g0[sfb];
            }
        }
        if (bt != 2 || mixed) {
            const int nl = bt != 2 ? 22 : 8;
            for (int sfb = 0; sfb < nl; sfb++) {
                grow[sfb] = __ldg(&T->gain[base_e - sf_shift * ((int)sfl[sfb] + (preflag ? (int)T->pretab[sfb] : 0))]);
                if (is_right && !lsf) srow[sfb] = sfl[sfb] < 7 ? sfl[sfb] : 255;   // MPEG-1: is_pos 7 (and above) = not intensity coded
            }
        }
        // part 3 starts where part 2 ends by the side info's arithmetic (not where the reads above stopped: scfsi)
    }
    if (silent) bv = 0;

    // ---- lockstep pair loop.  Every iteration each lane emits one pair.  Count1 quads are decoded only on even p (all
    // lanes of the warp together), a lane keeps up to three decoded pairs queued as 2-bit values; a lane that enters the
    // count1 region on an odd p decodes its first quad at once.
    // The loop runs in blocks of 16 pairs (one flush each).  Because the block's granule-channels are sorted by big_values,
    // most blocks are uniform for the whole warp and run a specialised body with no per-lane state tests:
    //   blocks below the warp's smallest big_values   -> big-values decoder only
    //   blocks above the warp's largest big_values    -> quad decoder (even p) + queue drain only
    //   the few blocks in between                     -> the general body
    uint32_t cq = 0;              // queued count1 pairs, 4 bits each (x | y<<2, two's complement), next pair in the low nibble
    int cqn = 0;
    int state = silent ? 2 : 0;   // 0 big values, 1 count1, 2 no more decoding (queued pairs still drain)
    bool overrun = false;
    uint32_t nz = 0;
    const uint8_t* c1 = s_c1[c1tab];
    // word offset of the spectrum row that lane r of this warp decodes (0xFFFFFFFF: no such granule-channel)
    __shared__ uint32_t s_rowoff[K1_THREADS];
    s_rowoff[tid] = active ? (uint32_t)gl * 288u : 0xFFFFFFFFu;
    uint32_t* blk_words = reinterpret_cast<uint32_t*>(spectra + ((size_t)tr.gc_base + blk.gc0) * 576);
    const int wmin = __reduce_min_sync(0xffffffffu, state == 0 ? bv : 0);
    const int wmax = __reduce_max_sync(0xffffffffu, state == 0 ? bv : 0);
    __syncwarp();

    auto big_pair = [&](int p, int& x, int& y) {
        const int line = 2 * p;
        const uint32_t reg = line < r0 ? reg0 : (line < r1 ? reg1 : reg2);
        const int linbits = (int)(reg & 0xFFu);
        br.refill();
        uint32_t w = br.top();
        uint32_t e = s_huff[(reg >> 8) + (w >> (32 - HUF_L1))];
        if (e & 0x8000u) {
            const uint32_t nb2 = (e >> 11) & 15;
            e = s_huff[((e & 0x7FFu) << 2) + ((w << HUF_L1) >> (32 - nb2))];
        }
        br.skip((int)(e >> 8));
        x = (e >> 4) & 15;
        y = e & 15;
        // linbits(x) sign(x) linbits(y) sign(y): at most 28 bits, peeled off one peek
        br.refill();
        w = br.top();
        const int lbx = (x == 15) ? linbits : 0;
        x += (int)((w >> 1) >> (31 - lbx));
        w <<= lbx;
        const int sx = x != 0;
        const bool negx = sx && (w >> 31);
        w <<= sx;
        const int lby = (y == 15) ? linbits : 0;
        y += (int)((w >> 1) >> (31 - lby));
        w <<= lby;
        const int sy = y != 0;
        const bool negy = sy && (w >> 31);
        br.skip(lbx + sx + lby + sy);
        x = negx ? -x : x;
        y = negy ? -y : y;
    };
    auto quad = [&](int p) {
        // the quad lands at pair p + cqn (oracle: while (pos < end && i <= 572))
        if (br.pos() < end && p + cqn <= 286) {
            br.refill();
            const uint32_t e = c1[br.top() >> 26];
            br.skip((int)(e >> 4));
            // four sign bits at most, present only for the non-zero values, in order v w x y
            uint32_t w = br.top();
            const uint32_t a0 = (e >> 3) & 1, a1 = (e >> 2) & 1, a2 = (e >> 1) & 1, a3 = e & 1;
            const uint32_t n0 = a0 & (w >> 31); w <<= a0;
            const uint32_t n1 = a1 & (w >> 31); w <<= a1;
            const uint32_t n2 = a2 & (w >> 31); w <<= a2;
            const uint32_t n3 = a3 & (w >> 31);
            br.skip((int)(a0 + a1 + a2 + a3));
            if (br.pos() > end) state = 2;  // quad straddles the end: discarded
            else {
                const uint32_t nibs = a0 | (n0 << 1) | (a1 << 2) | (n1 << 3) | (a2 << 4) | (n2 << 5) | (a3 << 6) | (n3 << 7);
                cq |= nibs << (4 * cqn);
                cqn += 2;
                nz += 4;
            }
        } else state = 2;
    };
    auto pop = [&](int& x, int& y) {
        if (cqn) {
            x = ((int)(cq << 30)) >> 30;
            y = ((int)(cq << 28)) >> 30;
            cq >>= 4;
            cqn--;
        }
    };
    auto park = [&](int pp, int x, int y) {
        s_stage[warp][lane][pp] = ((uint32_t)(uint16_t)(int16_t)x) | ((uint32_t)(uint16_t)(int16_t)y << 16);
    };

    for (int p0 = 0; p0 < 288; p0 += 16) {
        if (p0 + 16 <= wmin) {
            // every lane is inside its big-values region for the whole block
#pragma unroll 1
            for (int pp = 0; pp < 16; pp++) {
                int x, y;
                big_pair(p0 + pp, x, y);
                park(pp, x, y);
            }
        } else if (p0 > wmax) {
            // no lane is in its big-values region any more (and every lane has made its transition)
            if (__any_sync(0xffffffffu, state != 2 || cqn != 0)) {
#pragma unroll 1
                for (int pp = 0; pp < 16; pp += 2) {
                    if (state == 1 && cqn <= 1) quad(p0 + pp);
                    int x = 0, y = 0;
                    pop(x, y);
                    park(pp, x, y);
                    x = 0; y = 0;
                    if (state == 1 && cqn == 0) quad(p0 + pp + 1);   // (only a lane whose previous quad came up short)
                    pop(x, y);
                    park(pp + 1, x, y);
                }
            } else {
#pragma unroll
                for (int pp = 0; pp < 16; pp++) s_stage[warp][lane][pp] = 0;
            }
        } else {
#pragma unroll 1
            for (int pp = 0; pp < 16; pp++) {
                const int p = p0 + pp;
                int x = 0, y = 0;
                if (state == 0 && p >= bv) {
                    // end of the big-values region: an overrun silences the granule-channel (oracle: read_huffman)
                    if (br.pos() > end) { overrun = true; state = 2; }
                    else { state = 1; nz = 2 * bv; }
                }
                // warp-uniform guards (real branches, not predication): the big-values and the quad decoders are skipped
                // whenever no lane of the warp needs them in this iteration
                const bool big = state == 0;
                const bool wantq = state == 1 && (cqn == 0 || (cqn == 1 && !(p & 1)));
                if (__any_sync(0xffffffffu, big)) {
                    if (big) big_pair(p, x, y);
                }
                if (__any_sync(0xffffffffu, wantq)) {
                    if (wantq) quad(p);
                }
                if (!big) pop(x, y);
                park(pp, x, y);
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int row = 2 * k + (lane >> 4), col = lane & 15;
            const uint32_t v = s_stage[warp][row][col];
            const uint32_t ro = s_rowoff[warp * 32 + row];
            if (ro != 0xFFFFFFFFu) blk_words[ro + p0 + col] = v;
        }
        __syncwarp();
    }
    if (state == 0) {  // big_values == 288: the region check never ran inside the loop
        if (br.pos() > end) overrun = true; else nz = 576;
    }
    if (active) nz_out[tr.gc_base + gc] = (overrun || silent) ? 0 : (uint16_t)nz;
}
