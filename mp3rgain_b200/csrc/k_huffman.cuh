// k_huffman.cuh -- K1: Layer III part 2 (scalefactors) and part 3 (Huffman) decode, one thread per granule-channel.
//
// The bit offset of every granule-channel is known without decoding anything: frame f's data starts md_start bytes into the
// track's compacted main-data stream (K0) and granule-channels follow each other at part2_3_length strides, so all
// granule-channels of all frames decode in parallel.  A CTA owns 256 consecutive granule-channels of one track:
//
//   pre-pass   (thread = granule-channel in stream order) the 59 / 63 side-info bits of the granule-channel are pulled into
//              three registers with four aligned word loads; part2_3_length is prefix-summed over the frame's
//              granule-channels with shuffles -> first and last bit of the thread's stretch of the stream
//   stage      the CTA's slice of the main-data stream (24 KB at 128 kbps) arrives in shared memory as ONE bulk copy
//              (cp.async.bulk + mbarrier); a slice that exceeds the 28 KB window (high bitrates) is walked in several windows
//   sort       the 256 granule-channels are ordered by big_values (counting sort; by staging window first when the slice needs
//              several), so that the lanes of a warp leave the
//              big-values loop together
//   part 2     scalefactors -> one BYTE per scalefactor band (what to subtract from the granule's gain exponent; K2 looks
//              the power of two up), resolved against granule 0 for scfsi; the 64-byte rows leave through the warp's tile
//   phase 1    big values: every lane decodes one pair per iteration (two-level table lookup, linbits and signs peeled off
//              a second 32-bit peek without branches); the bit reader is a 3-word register window over the staged bytes
//              (one funnel shift per peek, one LDS per 32 consumed bits)
//   phase 2    count1: every lane decodes one quad per iteration; code -> pattern, pattern + the next four bits -> the two
//              output words, both by table
//   flush      each lane parks its words in a 16-word row of the warp's shared-memory tile; full rows leave as coalesced
//              64-byte segments.  Nothing is written beyond the last decoded line: K2 reads `nz` and treats the rest of
//              the row as zero.
//
// Follows ISO/IEC 11172-3 2.4.2.7 / 2.4.3.4 and 13818-3 2.4.3.2 exactly as the CPU restatement the tests check against (read_scalefactors,
// read_huffman) does; the reference delegates this stage to symphonia-bundle-mp3 (src/replaygain.rs:861-863, 896).
#pragma once

#define K1_THREADS 256
#define K1_GAIN_ROW 64        // BYTES per granule-channel: [0..21] long sfb, [22..60] short sfb*3+win, [62..63] gain exponent base (u16)
#define K1_C1_ROW 144         // bytes per granule-channel: at most (576 - 2 * big_values) / 4 quads
#define K1_STAGE 28672u       // bytes of main data staged per window (256 granule-channels at 128 kbps: 24.4 KB)
#define K1_SLACK 16u          // the bit window prefetches up to 12 bytes past the last bit it needs

struct K1Pre {                       // pre-pass -> main pass, indexed by granule-channel of the block in stream order
    uint32_t side[K1_THREADS][3];    //   its side-info bits, left-aligned
    uint32_t start[K1_THREADS];      //   first bit of the granule-channel in the track's main-data stream
    uint32_t mdend[K1_THREADS];      //   bytes of main data readable for its frame
    uint32_t info[K1_THREADS];       //   active | ch<<1 | gr<<2 | is_right<<3 | scfsi<<4 | part2_3_length<<8
    uint32_t cnt[512];               // counting sort: bin sizes, then bin starts
    uint32_t wsum[K1_THREADS / 32];
    uint8_t perm[K1_THREADS];        // worker thread -> granule-channel of the block, ordered by big_values
};
struct K1Smem {
    alignas(16) uint8_t stage[K1_STAGE];
    uint16_t huff[8192];
    union {
        alignas(16) uint32_t tile[K1_THREADS / 32][32 * 17];   // per warp: 32 gain rows of 17 words, then 32 spectrum rows of 16 (swizzled)
        K1Pre pre;                   // (every worker has taken its values into registers before the first tile row is written)
    };
    uint32_t rowoff[K1_THREADS];     // (worker order) word offset in the block of the next 16 words the worker's row flushes, 0xFFFFFFFF = none
    uint32_t lim[K1_THREADS];        // (worker order) words of that stretch that are complete
    uint8_t sfl[K1_THREADS][24];     // long scalefactors of every granule-channel (scfsi source)
    uint8_t c1[2][64];               // count1 tables A / B: 6-bit peek -> (len << 4) | pattern
    uint8_t c1res[256];              // (pattern << 4 | the four bits after the code word) -> v w x y as 2-bit two's complement, v in bits 0-1
    uint32_t base;                   // first main-data byte wanted by a granule-channel that is still pending
    alignas(8) unsigned long long mbar;
};
static_assert(sizeof(K1Pre) <= sizeof(uint32_t) * (K1_THREADS / 32) * 32 * 17, "pre-pass scratch must fit under the tiles");

#ifdef __CUDACC__
#define K1_NOINLINE __noinline__
#else
#define K1_NOINLINE __attribute__((noinline))
#endif
__device__ K1_NOINLINE uint32_t k1_mask_tail(uint32_t v, uint32_t off, uint32_t lim_rel) {
    return off >= lim_rel ? 0u : (v & (0xFFFFFFFFu << (8 * (off + 4 - lim_rel))));
}

// Big-endian bit window over the staged main data: the words at pos>>5, +1, +2 (hi, lo, nx) in registers.
struct BitWin {
    const uint8_t* st;     // the staged bytes
    uint32_t hi, lo, nx;
    uint32_t off;          // offset in `st` of the word that follows nx
    uint32_t pos;          // absolute position of the next unread bit
    uint32_t lim_rel;      // bytes of `st` readable for this frame (later bytes read as zero, like the oracle's reservoir)
    __device__ __forceinline__ uint32_t next() {
        uint32_t v = __byte_perm(*reinterpret_cast<const uint32_t*>(st + off), 0, 0x0123);
        if (off + 4 > lim_rel) v = k1_mask_tail(v, off, lim_rel);   // (a real branch: never taken by a sound stream)
        off += 4;
        return v;
    }
    // Bytes at or beyond lim_bytes read as zero -- which can only matter to a granule-channel whose part2_3_length claims
    // bits beyond them (a damaged stream: end_bit > 8 * lim_bytes).  Every other granule-channel either stops at its last bit
    // or is discarded / silenced once it has run past it, whatever the bits beyond were, so it reads the staged bytes as they are.
    __device__ __forceinline__ void init(const uint8_t* staged, uint32_t base_byte, uint32_t lim_bytes, uint32_t bitpos, uint32_t end_bit) {
        st = staged; lim_rel = end_bit > 8u * lim_bytes ? lim_bytes - base_byte : 0xFFFFFFF0u; pos = bitpos;
        off = ((bitpos >> 5) << 2) - base_byte;
        hi = next(); lo = next(); nx = next();
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(lo, hi, pos); }    // bits pos .. pos+31
    __device__ __forceinline__ uint32_t peek2() const { return __funnelshift_l(nx, lo, pos); }   // bits pos+32 .. pos+63
    // The refill is written without a branch: lanes cross their word boundaries at different symbols, so a branch would run
    // its body for a handful of lanes at nearly every symbol.  The word after nx is loaded every time (a shared-memory read)
    // and taken only by the lanes that crossed.
    __device__ __forceinline__ void advance_if(bool cross) {
        uint32_t v = __byte_perm(*reinterpret_cast<const uint32_t*>(st + off), 0, 0x0123);
        if (cross && off + 4 > lim_rel) v = k1_mask_tail(v, off, lim_rel);   // (a real branch: never taken by a sound stream)
        hi = cross ? lo : hi;
        lo = cross ? nx : lo;
        nx = cross ? v : nx;
        off += cross ? 4u : 0u;
    }
    __device__ __forceinline__ void skip(uint32_t n) {   // n <= 32: at most one word boundary is crossed
        const uint32_t np = pos + n;
        advance_if(((np ^ pos) & 32u) != 0);
        pos = np;
    }
    __device__ __forceinline__ void skip64(uint32_t n) {   // n <= 64
        const uint32_t np = pos + n, adv = (np >> 5) - (pos >> 5);
        advance_if(adv != 0);
        if (adv > 1) advance_if(true);   // (rare: more than 32 bits in one symbol takes linbits on both values)
        pos = np;
    }
    __device__ __forceinline__ uint32_t get(int n) {     // n <= 16, 0 allowed
        const uint32_t v = (peek() >> 1) >> (31 - n);
        skip((uint32_t)n);
        return v;
    }
};

__device__ __forceinline__ uint32_t side_field(const uint8_t* __restrict__ s, int bit, int n) {
    // n <= 12 bits starting at `bit` of the side-info bytes (big-endian bit order)
    const int by = bit >> 3;
    const uint32_t w = ((uint32_t)__ldg(s + by) << 16) | ((uint32_t)__ldg(s + by + 1) << 8) | __ldg(s + by + 2);
    return (w >> (24 - n - (bit & 7))) & ((1u << n) - 1);
}

// n (1..12) bits at offset o (0..83) of a 96-bit big-endian window held in three registers
__device__ __forceinline__ uint32_t k1_bits(uint32_t W0, uint32_t W1, uint32_t W2, int o, int n) {
    const uint32_t a = o < 32 ? W0 : (o < 64 ? W1 : W2);
    const uint32_t b = o < 32 ? W1 : (o < 64 ? W2 : 0u);
    return __funnelshift_l(b, a, o) >> (32 - n);
}

__global__ void __launch_bounds__(K1_THREADS, 3)
k_huffman(const uint8_t* __restrict__ corpus, const uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
          const DecTrack* __restrict__ tracks, const HuffBlock* __restrict__ blocks, const L3Dev* __restrict__ T,
          int16_t* __restrict__ spectra /* [gc][576] */, uint16_t* __restrict__ nz_out /* [gc] */,
          uint8_t* __restrict__ c1q /* [gc][K1_C1_ROW]: one byte per count1 quad, v w x y as 2-bit two's complement */,
          uint8_t* __restrict__ gains /* [gc][K1_GAIN_ROW] */,
          uint8_t* __restrict__ sfraw /* [gc][K1_GAIN_ROW]: raw scalefactors = is_pos, right channel of intensity-stereo frames only */) {
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem& S = *reinterpret_cast<K1Smem*>(k1_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(T->huff);
        uint4* dst = reinterpret_cast<uint4*>(S.huff);
        for (int i = tid; i < 8192 * 2 / 16; i += K1_THREADS) dst[i] = src[i];
        if (tid < 64) {
            S.c1[0][tid] = T->c1a_lut[tid];
            S.c1[1][tid] = (uint8_t)((4u << 4) | (15u - (tid >> 2)));
        }
        {   // pattern bits v w x y (v = bit 3); the sign bits of the non-zero values follow the code word in that order
            uint32_t res = 0, used = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((tid >> (7 - k)) & 1) { res |= (((tid >> (3 - used)) & 1) ? 3u : 1u) << (2 * k); used++; }
            S.c1res[tid] = (uint8_t)res;
        }
        if (tid == 0) k2_mbar_init(&S.mbar, 1);
    }
    const HuffBlock blk = blocks[blockIdx.x];
    const DecTrack tr = tracks[blk.track];
    const uint32_t nch = tr.channels, gpf = tr.gpf, gcf = gpf * nch;
    const bool lsf = gpf == 1;   // MPEG-2 / 2.5: one granule per frame, 63-bit side info per granule-channel (ISO 13818-3 2.4.1.7)
    const uint32_t n_gc = tr.n_frames * gcf;
    // field offsets inside one granule-channel: MPEG-1 has a 4-bit scalefac_compress and a preflag bit (59 bits in all),
    // MPEG-2 a 9-bit scalefac_compress and no preflag bit (63 bits)
    const int gbase = lsf ? (nch == 1 ? 9 : 10) : (nch == 1 ? 18 : 20), gstride = lsf ? 63 : 59, x5 = lsf ? 5 : 0;

    // ---- pre-pass, thread = granule-channel in stream order: side-info bits, position in the stream, sort key
    uint32_t key = 0x1FF;   // inactive slots sort last
    uint32_t pre_start = 0;  // my granule-channel's first bit in the track's stream
    {
        const uint32_t g = blk.gc0 + tid;
        uint32_t p23 = 0, md_start = 0, inf = 0, W0 = 0, W1 = 0, W2 = 0, md_end = 0;
        const uint32_t frame = g / gcf, rem = g % gcf, ch = rem % nch, gr = rem / nch;
        if (g < n_gc) {
            const FrameRec rec = frames[tr.frame_base + frame];
            const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
            const int b0 = gbase + gstride * (int)rem;
            const uint8_t* a = s + (b0 >> 3);
            const uint32_t* a4 = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)3);
            const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(a) & 3) + (uint32_t)(b0 & 7);
            const uint32_t w0 = __byte_perm(__ldg(a4), 0, 0x0123), w1 = __byte_perm(__ldg(a4 + 1), 0, 0x0123),
                           w2 = __byte_perm(__ldg(a4 + 2), 0, 0x0123), w3 = __byte_perm(__ldg(a4 + 3), 0, 0x0123);
            W0 = __funnelshift_l(w1, w0, sh); W1 = __funnelshift_l(w2, w1, sh); W2 = __funnelshift_l(w3, w2, sh);
            p23 = W0 >> 20;
            key = (W0 >> 11) & 0x1FFu;
            md_start = rec.md_start; md_end = rec.md_end;
            const uint32_t is_right = (ch == 1 && ((rec.hdr >> 6) & 3) == 1 && ((rec.hdr >> 4) & 1)) ? 1u : 0u;
            const uint32_t scfsi = (!lsf && gr == 1) ? side_field(s, (nch == 1 ? 14 : 12) + 4 * (int)ch, 4) : 0u;  // band group 0 first
            inf = 1u | (ch << 1) | (gr << 2) | (is_right << 3) | (scfsi << 4) | (p23 << 8);
        }
        // granule-channels of one frame sit in gcf consecutive lanes (256 and 32 are multiples of gcf = 1, 2 or 4)
        uint32_t before = 0;
#pragma unroll
        for (int o = 1; o < 4; o++) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, p23, o);
            if ((uint32_t)o <= rem) before += up;
        }
        S.pre.side[tid][0] = W0; S.pre.side[tid][1] = W1; S.pre.side[tid][2] = W2;
        pre_start = md_start * 8 + before;
        S.pre.start[tid] = pre_start;
        S.pre.mdend[tid] = md_end;
        S.pre.info[tid] = inf;
#pragma unroll
        for (int i = 0; i < 6; i++) reinterpret_cast<uint32_t*>(S.sfl[tid])[i] = 0;
    }
    // ---- counting sort of the block's granule-channels over the 9-bit big_values: my slot inside my bin, then an exclusive
    // prefix over the 512 bins (two per thread)
    {
        S.pre.cnt[tid] = 0; S.pre.cnt[tid + 256] = 0;
        __syncthreads();
        // A slice longer than the staging window (256-320 kbps) is walked in several windows, each serving the granule-channels
        // whose bits lie inside it -- consecutive ones in stream order.  Sorted by big_values alone they would be spread over all
        // eight warps, and every warp would run every window's pass at a third of its lanes; with the (estimated) window as the
        // major key a warp belongs to one window and sits the other passes out.  (The estimate only orders the workers: which
        // window actually serves a granule-channel is decided in the loop below.)
        {
            const uint32_t n_act = min((uint32_t)K1_THREADS, n_gc - blk.gc0);
            const uint32_t first = S.pre.start[0] >> 3, span = (S.pre.start[n_act - 1] >> 3) + 512u - first;   // part2_3_length < 4096 bits
            if (span + K1_SLACK > K1_STAGE && key != 0x1FFu)
                key = (min(3u, ((pre_start >> 3) - first) / (K1_STAGE - 1024u)) << 7) | (key >> 2);
        }
        const uint32_t slot = atomicAdd(&S.pre.cnt[key], 1u);
        __syncthreads();
        const uint32_t c0 = S.pre.cnt[2 * tid], c1 = S.pre.cnt[2 * tid + 1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) S.pre.wsum[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += S.pre.wsum[w];
        S.pre.cnt[2 * tid] = base + incl - c0 - c1;        // bins 2*tid, 2*tid+1: where their first member goes
        S.pre.cnt[2 * tid + 1] = base + incl - c1;
        __syncthreads();
        S.pre.perm[S.pre.cnt[key] + slot] = (uint8_t)tid;
    }
    __syncthreads();

    // ---- main pass, thread = worker: side info of my granule-channel
    const int gl = S.pre.perm[tid];                   // my granule-channel within the block
    const uint32_t gc = blk.gc0 + gl;
    const uint32_t inf = S.pre.info[gl];
    const bool active = inf & 1u;
    const uint32_t ch = (inf >> 1) & 1, gr = (inf >> 2) & 1;
    const bool is_right = (inf >> 3) & 1;   // right channel of an intensity-stereo frame: K2 needs its raw scalefactors (is_pos)
    const int scfsi = (int)((inf >> 4) & 15);
    const uint32_t my_start = S.pre.start[gl], my_end = my_start + (inf >> 8), my_mdend = S.pre.mdend[gl];
    int bv = 0, r0 = 0, r1 = 0, c1tab = 0;
    uint32_t reg0 = 0, reg1 = 0, reg2 = 0;   // per region: first-level table base << 8 | linbits
    bool silent = true;
    int bt = 0, mixed = 0, slen1 = 0, slen2 = 0, base_e = 0, sf_shift = 2, preflag = 0, sbg0 = 0, sbg1 = 0, sbg2 = 0;
    int ls[4] = {0, 0, 0, 0}, ltbl = 0, cls = 0;   // MPEG-2: field widths and partition table from scalefac_compress
    if (active) {
        const uint32_t W0 = S.pre.side[gl][0], W1 = S.pre.side[gl][1], W2 = S.pre.side[gl][2];
        auto fld = [&](int o, int n) { return (int)k1_bits(W0, W1, W2, o, n); };
        const uint32_t p23 = inf >> 8;
        bv = fld(12, 9);
        base_e = fld(21, 8) - 210 - L3_GAIN_MIN;
        const int sfc = fld(29, lsf ? 9 : 4);
        const int ws = fld(33 + x5, 1);
        int sel0, sel1, sel2 = 0;
        if (ws) {
            bt = fld(34 + x5, 2);
            mixed = fld(36 + x5, 1);
            sel0 = fld(37 + x5, 5);
            sel1 = fld(42 + x5, 5);
            sbg0 = fld(47 + x5, 3); sbg1 = fld(50 + x5, 3); sbg2 = fld(53 + x5, 3);
            // region 0 = the first 8 long bands, or the first 3 short bands x 3 windows: 36 lines at every MPEG-1 rate,
            // 54 / 36 at the MPEG-2 rates, 108 / 72 at 8 kHz
            r0 = bt == 2 ? 3 * T->short_start[tr.sr_idx][3] : T->long_start[tr.sr_idx][8];
            r1 = 576;
        } else {
            sel0 = fld(34 + x5, 5);
            sel1 = fld(39 + x5, 5);
            sel2 = fld(44 + x5, 5);
            const int rc0 = fld(49 + x5, 4), rc1 = fld(53 + x5, 3);
            r0 = T->long_start[tr.sr_idx][rc0 + 1];
            const int k = rc0 + rc1 + 2;
            r1 = T->long_start[tr.sr_idx][k > 22 ? 22 : k];
        }
        if (bt != 2) mixed = 0;
        preflag = lsf ? (sfc >= 500 && !is_right) : fld(56, 1);
        sf_shift = fld(lsf ? 61 : 57, 1) ? 4 : 2;
        c1tab = fld(lsf ? 62 : 58, 1);
        // slot 0 ("no data") reads the all-zero first-level table at huff_base[0]: length 0, x = y = 0
        reg0 = ((uint32_t)T->huff_base[T->tabsel[sel0][0]] << 8) | T->tabsel[sel0][1];
        reg1 = ((uint32_t)T->huff_base[T->tabsel[sel1][0]] << 8) | T->tabsel[sel1][1];
        reg2 = ((uint32_t)T->huff_base[T->tabsel[sel2][0]] << 8) | T->tabsel[sel2][1];
        // part-2 length
        int part2 = 0;
        cls = bt != 2 ? 0 : (mixed ? 2 : 1);
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
            if (is_right) sfraw[((size_t)tr.gc_base + gc) * K1_GAIN_ROW + 63] = (uint8_t)(sfc & 1);   // intensity_scale
        } else {
            slen1 = T->slen[0][sfc]; slen2 = T->slen[1][sfc];
            if (bt == 2) part2 = (mixed ? 17 : 18) * slen1 + 18 * slen2;
            else {
                part2 += (scfsi & 8) ? 0 : 6 * slen1;
                part2 += (scfsi & 4) ? 0 : 5 * slen1;
                part2 += (scfsi & 2) ? 0 : 5 * slen2;
                part2 += (scfsi & 1) ? 0 : 5 * slen2;
            }
            if (is_right) sfraw[((size_t)tr.gc_base + gc) * K1_GAIN_ROW + 63] = 0;
        }
        silent = (uint32_t)part2 > p23;
    }
    if (silent) bv = 0;
    uint32_t* blk_words = reinterpret_cast<uint32_t*>(spectra + ((size_t)tr.gc_base + blk.gc0) * 576);
    uint32_t* blk_gains = reinterpret_cast<uint32_t*>(gains + ((size_t)tr.gc_base + blk.gc0) * K1_GAIN_ROW);
    uint8_t* srow = sfraw + ((size_t)tr.gc_base + gc) * K1_GAIN_ROW;
    uint8_t* sfl = S.sfl[gl];
    uint8_t* grow = reinterpret_cast<uint8_t*>(&S.tile[warp][lane * 17]);   // part 2 parks the 64-byte gain row in the warp's tile
    uint4* t4 = reinterpret_cast<uint4*>(S.tile[warp]);   // parts 3: 32 rows of four 16-byte chunks, chunk c of row r at c ^ ((r >> 1) & 3)
    const int swz = (lane >> 1) & 3;
    const uint8_t* c1 = S.c1[c1tab];
    const uint32_t avail = (tr.md_bytes + 16 + 15) & ~15u;   // bytes of the track's stream that may be copied (its region is longer)
    const uint8_t* md_track = md + tr.md_off;
    const uint32_t my_row = (uint32_t)gl * 288u;

    // ---- windows over the block's slice of the stream: all of it at once unless the slice exceeds K1_STAGE bytes
    bool pending = active;
    for (uint32_t pass = 0;; pass++) {
        if (tid == 0) S.base = 0xFFFFFFFFu;
        __syncthreads();
        if (pending) atomicMin(&S.base, my_start >> 3);
        __syncthreads();
        const uint32_t first = S.base;
        if (first == 0xFFFFFFFFu) break;
        const uint32_t base = first & ~15u;
        // the granule-channel with the smallest start always fits: part2_3_length < 4096 bits
        const bool in = pending && (my_end >> 3) + K1_SLACK <= base + K1_STAGE;
        const int last_pass = __syncthreads_and(!pending || in);   // (also: everybody has read S.base)
        if (tid == 0) {
            const uint32_t bytes = base < avail ? min(K1_STAGE, avail - base) : 16u;
            k2_mbar_expect(&S.mbar, bytes);
            k2_bulk_load(S.stage, md_track + base, bytes, &S.mbar);
        }
        S.rowoff[tid] = in ? my_row : 0xFFFFFFFFu;
        k2_mbar_wait(&S.mbar, pass & 1);

        BitWin br;
        br.st = S.stage; br.lim_rel = 0; br.pos = 0; br.off = 0; br.hi = br.lo = br.nx = 0;
        if (in) {
            br.init(S.stage, base, my_mdend, my_start, my_end);
            // ---- part 2 (ISO 11172-3 2.4.2.7): my own scalefactors.  Long ones go to shared memory (granule 1 may need them),
            // short ones straight into their band's byte.
            if (lsf) {
                // ISO 13818-3 2.4.3.2: the values come in band order, lsf_nsf[..][i] of them ls[i] bits wide: long sfb 0..20, or
                // (mixed: 6 long bands, then) short sfb x 3 windows
                int k = 0, f = 0, left = T->lsf_nsf[ltbl][cls][0];
                const int n_long = cls == 0 ? 21 : (cls == 2 ? 6 : 0);
                const int n_all = n_long + (cls == 0 ? 0 : (cls == 2 ? 27 : 36));
                for (; k < n_all; k++) {
                    while (left == 0 && f < 3) left = T->lsf_nsf[ltbl][cls][++f];
                    const int v = left ? (int)br.get(ls[f]) : 0;
                    if (left) left--;
                    // intensity position for K2: 255 = the illegal (all-ones) position of a non-empty field -> band not intensity coded
                    const uint8_t ipos = (ls[f] > 0 && v == (1 << ls[f]) - 1) ? 255 : (uint8_t)v;
                    if (k < n_long) { sfl[k] = (uint8_t)v; if (is_right) srow[k] = ipos; }
                    else {
                        const int q2 = k - n_long + (cls == 2 ? 9 : 0), win = q2 % 3;
                        if (is_right) srow[22 + q2] = ipos;
                        grow[22 + q2] = (uint8_t)(8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2)) + sf_shift * v);
                    }
                }
                if (cls)
#pragma unroll
                    for (int win = 0; win < 3; win++) grow[22 + 36 + win] = (uint8_t)(8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2)));
            } else if (bt == 2) {
                if (mixed) for (int sfb = 0; sfb < 8; sfb++) sfl[sfb] = (uint8_t)br.get(slen1);
                for (int q2 = mixed ? 9 : 0; q2 < 36; q2++) {
                    const int v = (int)br.get(q2 < 18 ? slen1 : slen2);
                    const int win = q2 % 3;
                    if (is_right) srow[22 + q2] = (uint8_t)(v < 7 ? v : 255);
                    grow[22 + q2] = (uint8_t)(8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2)) + sf_shift * v);
                }
#pragma unroll
                for (int win = 0; win < 3; win++)   // sfb 12 carries no scalefactor; a mixed block has no short sfb 0..2
                    grow[22 + 36 + win] = (uint8_t)(8 * (win == 0 ? sbg0 : (win == 1 ? sbg1 : sbg2)));
            } else {
                // long blocks: up to 11 fields of slen1 bits (<= 44 bits), then up to 10 of slen2 (<= 30): two 64-bit peeks, the
                // fields sliced out with shifts (a band group that scfsi copies from granule 0 occupies no bits)
                {
                    const unsigned long long w64 = ((unsigned long long)br.peek() << 32) | br.peek2();
                    uint32_t o = 0;
                    if (slen1) {
#pragma unroll 1
                        for (int sfb = 0; sfb < 11; sfb++)
                            if (!(scfsi & (sfb < 6 ? 8 : 4))) { sfl[sfb] = (uint8_t)((w64 << o) >> (64 - slen1)); o += slen1; }
                    }
                    br.skip64(o);
                }
                {
                    const unsigned long long w64 = ((unsigned long long)br.peek() << 32) | br.peek2();
                    uint32_t o = 0;
                    if (slen2) {
#pragma unroll 1
                        for (int sfb = 11; sfb < 21; sfb++)
                            if (!(scfsi & (sfb < 16 ? 2 : 1))) { sfl[sfb] = (uint8_t)((w64 << o) >> (64 - slen2)); o += slen2; }
                    }
                    br.skip64(o);
                }
            }
        }
        __syncthreads();
        if (in) {
            // scfsi: granule 1 reuses granule 0's scalefactors of the same channel (granule-channel gl - nch: same frame, same
            // block, decoded in this window or an earlier one).  A short granule 0 has no long scalefactors (zero) except the 8
            // of a mixed block -- exactly what its row holds.
            if (bt != 2 && scfsi) {
                const uint8_t* g0 = S.sfl[gl - nch];
                for (int sfb = 0; sfb < 21; sfb++) {
                    const int grp = sfb < 6 ? 0 : (sfb < 11 ? 1 : (sfb < 16 ? 2 : 3));
                    if (scfsi & (8 >> grp)) sfl[sfb] = g0[sfb];
                }
            }
            if (bt != 2 || mixed) {
                const int nl = bt != 2 ? 22 : 8;
                for (int sfb = 0; sfb < nl; sfb++) {
                    grow[sfb] = (uint8_t)(sf_shift * ((int)sfl[sfb] + (preflag ? (int)T->pretab[sfb] : 0)));
                    if (is_right && !lsf) srow[sfb] = sfl[sfb] < 7 ? sfl[sfb] : 255;   // MPEG-1: is_pos 7 (and above) = not intensity coded
                }
            }
            reinterpret_cast<uint16_t*>(grow)[31] = (uint16_t)base_e;
            // part 3 starts where part 2 ends by the side info's arithmetic (not where the reads above stopped: scfsi)
        }
        __syncwarp();
        // the warp's 32 gain rows (64 bytes each) leave as coalesced 64-byte segments
#pragma unroll 4
        for (int k = 0; k < 16; k++) {
            const int row = 2 * k + (lane >> 4), col = lane & 15;
            const uint32_t ro = S.rowoff[warp * 32 + row];
            if (ro != 0xFFFFFFFFu) blk_gains[(ro / 288u) * (K1_GAIN_ROW / 4) + col] = S.tile[warp][row * 17 + col];
        }
        __syncwarp();

        // ---- phase 1: big values.  Every lane decodes one pair per step while it has any; four pairs make a 16-byte chunk of
        // the lane's tile row, four chunks (16 pairs) a flush (a lane's last chunk may be partly zero: K2 knows big_values).
        const int mybv = in ? bv : 0;
        const int wmin = __reduce_min_sync(0xffffffffu, mybv);
        const int wmax = __reduce_max_sync(0xffffffffu, mybv);
        S.lim[tid] = (uint32_t)mybv;
        __syncwarp();
        auto big_pair = [&](int p) -> uint32_t {
            const int line = 2 * p;
            const uint32_t reg = line < r0 ? reg0 : (line < r1 ? reg1 : reg2);
            const int linbits = (int)(reg & 0xFFu);
            const uint32_t wh = br.peek(), wl = br.peek2();
            uint32_t e = S.huff[(reg >> 8) + (wh >> (32 - HUF_L1))];
            if (e & 0x8000u) {
                const uint32_t nb2 = (e >> 11) & 15;
                e = S.huff[((e & 0x7FFu) << 2) + ((wh << HUF_L1) >> (32 - nb2))];
            }
            const uint32_t len = e >> 8;   // <= 19
            int x = (e >> 4) & 15, y = e & 15;
            // linbits(x) sign(x) linbits(y) sign(y): at most 28 bits after the code word, all inside the 64-bit peek
            uint32_t w = __funnelshift_l(wl, wh, len);
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
            br.skip64(len + (uint32_t)(lbx + sx + lby + sy));
            x = negx ? -x : x;
            y = negy ? -y : y;
            return ((uint32_t)(uint16_t)(int16_t)x) | ((uint32_t)(uint16_t)(int16_t)y << 16);
        };
        for (int p0 = 0; p0 < wmax; p0 += 16) {
            const bool full = p0 + 16 <= wmin;   // every lane is inside its big-values region for the whole block of 16
            if (full) {
#pragma unroll 1
                for (int g = 0; g < 4; g++) {
                    const int pb = p0 + 4 * g;
                    const uint32_t v0 = big_pair(pb), v1 = big_pair(pb + 1), v2 = big_pair(pb + 2), v3 = big_pair(pb + 3);
                    t4[lane * 4 + (g ^ swz)] = make_uint4(v0, v1, v2, v3);
                }
            } else {
#pragma unroll 1
                for (int g = 0; g < 4; g++) {
                    const int pb = p0 + 4 * g;
                    if (pb >= wmax) break;
                    uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll 1
                    for (int k = 0; k < 4; k++) {
                        const uint32_t w = pb + k < mybv ? big_pair(pb + k) : 0u;
                        v[0] = k == 0 ? w : v[0]; v[1] = k == 1 ? w : v[1]; v[2] = k == 2 ? w : v[2]; v[3] = k == 3 ? w : v[3];
                    }
                    t4[lane * 4 + (g ^ swz)] = make_uint4(v[0], v[1], v[2], v[3]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int row = 8 * k + (lane >> 2), chunk = lane & 3;
                const uint32_t ro = S.rowoff[warp * 32 + row];
                if (ro != 0xFFFFFFFFu && (full || (uint32_t)(p0 + 4 * chunk) < S.lim[warp * 32 + row]))
                    *reinterpret_cast<uint4*>(blk_words + ro + p0 + 4 * chunk) = t4[row * 4 + (chunk ^ ((row >> 1) & 3))];
            }
            __syncwarp();
        }

        // ---- phase 2: count1 quads from pair `bv` on (oracle: while (pos < end && line <= 572)); a quad that straddles the end of
        // the granule-channel's bits is discarded.  An overrun of the big-values region silences the granule-channel (oracle:
        // read_huffman).  A quad leaves as ONE byte -- v w x y as 2-bit two's complement, through a 256-entry table indexed
        // by its pattern and the four bits after the code word -- sixteen quads per 16-byte store; K2 expands them.
        bool overrun = false;
        uint32_t nz = 0;
        bool live = false;
        int q = mybv;
        if (in && !silent) {
            if (br.pos > my_end) overrun = true;
            else { nz = 2u * (uint32_t)bv; live = true; }
        }
        uint4* qrow = reinterpret_cast<uint4*>(c1q + ((size_t)tr.gc_base + gc) * K1_C1_ROW);
        while (__any_sync(0xffffffffu, live)) {
            uint32_t acc[4] = {0, 0, 0, 0};
            const int q_before = q;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (live) {
                    live = false;
                    if (br.pos < my_end && q <= 286) {
                        const uint32_t w = br.peek();
                        const uint32_t e = c1[w >> 26];
                        const uint32_t len = e >> 4, pat = e & 15u;
                        const uint32_t tot = len + (uint32_t)__popc(pat);
                        if (br.pos + tot <= my_end) {   // (else: straddles the end, discarded)
                            acc[k >> 2] |= (uint32_t)S.c1res[(pat << 4) | ((w << len) >> 28)] << (8 * (k & 3));
                            br.skip(tot);
                            q += 2; nz += 4;
                            live = true;
                        }
                    }
                }
            }
            // (only a group that holds a quad is stored: a granule-channel whose count1 region ends exactly on a group boundary
            // -- big_values 0 and all 144 quads present fills its row to the last byte -- must not write a tenth, empty group
            // into its neighbour's row)
            if (q != q_before) *qrow++ = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        }
        if (in) nz_out[tr.gc_base + gc] = (overrun || silent) ? 0 : (uint16_t)nz;
        if (last_pass) break;
        pending = pending && !in;
        __syncthreads();   // the staging buffer and S.base are reused by the next window
    }
}
