// k_huffman.cuh -- K1: Layer III part-3 (Huffman) decode, one thread per granule-channel.
//
// The bit offset of every granule-channel is known without decoding anything: frame f's data starts md_start bytes into the
// track's compacted main-data stream (K0) and granule-channels follow each other at part2_3_length strides, so all
// granule-channels of all frames decode in parallel.  Scalefactor bits (part 2) are skipped here -- their length follows
// from the side info -- and extracted directly by K2.
//
// All lanes of a warp advance through the 288 line pairs in lockstep; each lane parks its pair in a 16-pair shared-memory
// row and every 16 pairs the warp flushes 32 rows with coalesced 64-byte segments, so the int16 spectra reach HBM as full
// sectors although every lane owns a different 1152-byte row.
#pragma once

#define K1_THREADS 128

struct BitRd {
    const uint32_t* w;   // main-data stream as 32-bit words (16-byte aligned base)
    uint32_t lim;        // bytes readable for this frame (later bytes read as zero, like the oracle's reservoir)
    uint32_t cw0, cw1;   // big-endian words at index wi, wi+1
    uint32_t wi;
    uint32_t pos;        // bit position
    __device__ __forceinline__ uint32_t ld(uint32_t i) const {
        const uint32_t b = i * 4;
        if (b >= lim) return 0;
        uint32_t v = __byte_perm(__ldg(w + i), 0, 0x0123);
        if (b + 4 > lim) v &= 0xFFFFFFFFu << (8 * (b + 4 - lim));
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t* base, uint32_t lim_bytes, uint32_t bitpos) {
        w = reinterpret_cast<const uint32_t*>(base);
        lim = lim_bytes;
        pos = bitpos;
        wi = pos >> 5;
        cw0 = ld(wi);
        cw1 = ld(wi + 1);
    }
    __device__ __forceinline__ uint32_t peek() {
        const uint32_t nwi = pos >> 5;
        if (nwi != wi) {
            if (nwi == wi + 1) { cw0 = cw1; cw1 = ld(nwi + 1); }
            else { cw0 = ld(nwi); cw1 = ld(nwi + 1); }
            wi = nwi;
        }
        return __funnelshift_l(cw1, cw0, pos & 31);
    }
};

__device__ __forceinline__ uint32_t side_field(const uint8_t* __restrict__ s, int bit, int n) {
    // n <= 12 bits starting at `bit` of the side-info bytes (big-endian bit order)
    const int by = bit >> 3;
    const uint32_t w = ((uint32_t)__ldg(s + by) << 16) | ((uint32_t)__ldg(s + by + 1) << 8) | __ldg(s + by + 2);
    return (w >> (24 - n - (bit & 7))) & ((1u << n) - 1);
}

__global__ void __launch_bounds__(K1_THREADS)
k_huffman(const uint8_t* __restrict__ corpus, const uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
          const DecTrack* __restrict__ tracks, const HuffBlock* __restrict__ blocks, const L3Dev* __restrict__ T,
          int16_t* __restrict__ spectra /* [gc][576] */, uint16_t* __restrict__ nz_out /* [gc] */) {
    __shared__ uint16_t s_huff[8192];
    __shared__ uint32_t s_stage[K1_THREADS / 32][32][17];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(T->huff);
        uint4* dst = reinterpret_cast<uint4*>(s_huff);
        for (int i = tid; i < 8192 * 2 / 16; i += K1_THREADS) dst[i] = src[i];
    }
    const HuffBlock blk = blocks[blockIdx.x];
    const DecTrack tr = tracks[blk.track];
    const uint32_t nch = tr.channels;
    const uint32_t n_gc = tr.n_frames * 2 * nch;
    const uint32_t gc = blk.gc0 + tid;
    const bool active = gc < n_gc;
    __syncthreads();

    // ---- side info of my granule-channel
    int bv = 0, r0 = 0, r1 = 0, c1tab = 0;
    int sel0 = 0, sel1 = 0, sel2 = 0;
    int slot0 = 0, slot1 = 0, slot2 = 0, lb0 = 0, lb1 = 0, lb2 = 0;
    uint32_t end = 0;
    bool silent = true;
    BitRd br;
    br.w = reinterpret_cast<const uint32_t*>(md + tr.md_off);
    br.lim = 0; br.cw0 = br.cw1 = 0; br.wi = 0; br.pos = 0;
    if (active) {
        const uint32_t frame = gc / (2 * nch), rem = gc % (2 * nch), gr = rem / nch, ch = rem % nch;
        const FrameRec rec = frames[tr.frame_base + frame];
        const uint8_t* s = corpus + tr.corpus_off + rec.file_off + 4 + ((rec.hdr & 0x10000u) ? 0 : 2);
        const int gbase = (nch == 1 ? 18 : 20);
        uint32_t bit = rec.md_start * 8;
        for (uint32_t k = 0; k < rem; k++) bit += side_field(s, gbase + 59 * k, 12);
        const int b0 = gbase + 59 * rem;
        const uint32_t p23 = side_field(s, b0, 12);
        bv = side_field(s, b0 + 12, 9);
        const int sfc = side_field(s, b0 + 29, 4);
        const int ws = side_field(s, b0 + 33, 1);
        int bt = 0, mixed = 0;
        if (ws) {
            bt = side_field(s, b0 + 34, 2);
            mixed = side_field(s, b0 + 36, 1);
            sel0 = side_field(s, b0 + 37, 5);
            sel1 = side_field(s, b0 + 42, 5);
            r0 = 36; r1 = 576;
        } else {
            sel0 = side_field(s, b0 + 34, 5);
            sel1 = side_field(s, b0 + 39, 5);
            sel2 = side_field(s, b0 + 44, 5);
            const int rc0 = side_field(s, b0 + 49, 4), rc1 = side_field(s, b0 + 53, 3);
            r0 = T->long_start[tr.sr_idx][rc0 + 1];
            const int k = rc0 + rc1 + 2;
            r1 = T->long_start[tr.sr_idx][k > 22 ? 22 : k];
        }
        c1tab = side_field(s, b0 + 58, 1);
        slot0 = T->tabsel[sel0][0]; lb0 = T->tabsel[sel0][1];
        slot1 = T->tabsel[sel1][0]; lb1 = T->tabsel[sel1][1];
        slot2 = T->tabsel[sel2][0]; lb2 = T->tabsel[sel2][1];
        // part-2 length
        const int slen1 = T->slen[0][sfc], slen2 = T->slen[1][sfc];
        int part2;
        if (bt == 2) part2 = (mixed ? 17 : 18) * slen1 + 18 * slen2;
        else {
            part2 = 0;
            const int sc = (gr == 1) ? (int)side_field(s, (nch == 1 ? 14 : 12) + 4 * ch, 4) : 0;  // scfsi bits, band group 0 first
            part2 += (sc & 8) ? 0 : 6 * slen1;
            part2 += (sc & 4) ? 0 : 5 * slen1;
            part2 += (sc & 2) ? 0 : 5 * slen2;
            part2 += (sc & 1) ? 0 : 5 * slen2;
        }
        end = bit + p23;
        silent = (uint32_t)part2 > p23;
        br.init(md + tr.md_off, rec.md_end, bit + part2);
    }
    if (silent) bv = 0;

    // ---- lockstep pair loop
    const uint16_t* hb = T->huff_base;
    int16_t q2 = 0, q3 = 0;
    bool have_second = false;
    int state = silent ? 2 : 0;   // 0 big values, 1 count1, 2 zero tail
    bool overrun = false;
    uint32_t nz = 0;
    int16_t* my_rows = spectra + ((size_t)tr.gc_base + blk.gc0 + warp * 32) * 576;
    const uint32_t warp_gc0 = blk.gc0 + warp * 32;

    for (int p0 = 0; p0 < 288; p0 += 16) {
        if (__any_sync(0xffffffffu, state != 2)) {
#pragma unroll 1
            for (int pp = 0; pp < 16; pp++) {
                const int p = p0 + pp;
                int x = 0, y = 0;
                if (state == 0 && p >= bv) {
                    // end of the big-values region: an overrun silences the granule-channel (oracle: read_huffman)
                    if (br.pos > end) { overrun = true; state = 2; }
                    else { state = 1; nz = 2 * bv; }
                }
                if (state == 0) {
                    const int line = 2 * p;
                    const int slot = line < r0 ? slot0 : (line < r1 ? slot1 : slot2);
                    const int linbits = line < r0 ? lb0 : (line < r1 ? lb1 : lb2);
                    if (slot) {
                        const uint32_t w = br.peek();
                        uint32_t e = s_huff[hb[slot] + (w >> (32 - HUF_L1))];
                        if (e & 0x8000u) {
                            const uint32_t nb = (e >> 11) & 15;
                            e = s_huff[((e & 0x7FFu) << 2) + ((w << HUF_L1) >> (32 - nb))];
                        }
                        br.pos += e >> 8;
                        x = (e >> 4) & 15;
                        y = e & 15;
                        if (x == 15 && linbits) { x += br.peek() >> (32 - linbits); br.pos += linbits; }
                        if (x) { if (br.peek() >> 31) x = -x; br.pos++; }
                        if (y == 15 && linbits) { y += br.peek() >> (32 - linbits); br.pos += linbits; }
                        if (y) { if (br.peek() >> 31) y = -y; br.pos++; }
                    }
                } else if (state == 1) {
                    if (have_second) { x = q2; y = q3; have_second = false; }
                    else if (br.pos < end && p <= 286) {
                        const uint32_t w = br.peek();
                        uint32_t e;
                        if (c1tab) e = (4u << 4) | (15u - (w >> 28));
                        else e = T->c1a_lut[w >> 26];
                        br.pos += e >> 4;
                        int v0 = (e >> 3) & 1, v1 = (e >> 2) & 1, v2 = (e >> 1) & 1, v3 = e & 1;
                        if (v0) { if (br.peek() >> 31) v0 = -1; br.pos++; }
                        if (v1) { if (br.peek() >> 31) v1 = -1; br.pos++; }
                        if (v2) { if (br.peek() >> 31) v2 = -1; br.pos++; }
                        if (v3) { if (br.peek() >> 31) v3 = -1; br.pos++; }
                        if (br.pos > end) state = 2;  // quad straddles the end: discarded
                        else { x = v0; y = v1; q2 = (int16_t)v2; q3 = (int16_t)v3; have_second = true; nz += 4; }
                    } else state = 2;
                }
                s_stage[warp][lane][pp] = ((uint32_t)(uint16_t)(int16_t)x) | ((uint32_t)(uint16_t)(int16_t)y << 16);
            }
        } else {
#pragma unroll
            for (int pp = 0; pp < 16; pp++) s_stage[warp][lane][pp] = 0;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int row = 2 * k + (lane >> 4), col = lane & 15;
            const uint32_t v = s_stage[warp][row][col];
            if (warp_gc0 + row < n_gc)
                *reinterpret_cast<uint32_t*>(my_rows + (size_t)row * 576 + 2 * (p0 + col)) = v;
        }
        __syncwarp();
    }
    if (state == 0) {  // big_values == 288: the region check never ran inside the loop
        if (br.pos > end) overrun = true; else nz = 576;
    }
    if (active) nz_out[tr.gc_base + gc] = (overrun || silent) ? 0 : (uint16_t)nz;
}
