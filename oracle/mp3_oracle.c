/*
 * mp3_oracle.c -- CPU ORACLE, stage A: MPEG-1 Layer III bytes -> planar f32 PCM (test infrastructure, NOT product code).
 *
 * The reference (mp3rgain v1.5.0) does not contain this stage: src/replaygain.rs:807-822, 861-863, 882, 896 call the
 * third-party crate symphonia-bundle-mp3 0.5.5 (Cargo.toml:26, Cargo.lock:244-247), whose source is absent from
 * /root/reference.  This file therefore restates the published algorithm, ISO/IEC 11172-3 Layer III, not any file of
 * the reference; bit parity with symphonia is UNPINNED.  It is cross-checked (tolerance) against FFmpeg's mp3float
 * decoder in tests/test_oracle_mp3.py.
 *
 * Demuxer behaviour mirrors what the reference's call sites imply (see DESIGN.md "symphonia switches"):
 *   - a leading ID3v2 tag is skipped; a first frame carrying a Xing/Info/VBRI tag is not emitted (unless flag 0x2);
 *   - FormatOptions::default() => no gapless trimming: every decoded sample is analysed;
 *   - a frame whose main_data_begin reaches before the start of the reservoir, or whose side info is invalid,
 *     is a DecodeError => the reference `continue`s (src/replaygain.rs:896-900): its PCM is dropped;
 *   - a truncated final frame is EOF (src/replaygain.rs:884-888);
 *   - output is unclamped f32 (flag 0x4 clamps to [-1,1]).
 *
 * Arithmetic contract (mirrored operation-for-operation by the CUDA kernels; compile with -ffp-contract=off):
 *   all float tables come from l3_tables_oracle.h; products and sums round separately except where fmaf() is
 *   written, and every dot product is a sequential fmaf chain in ascending index order starting from +0.0f.
 */
#include "mp3_oracle.h"
#include "l3_tables_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ lookup structures built once */
#define HUFF_L1_BITS 9
typedef struct { uint16_t* tab; int size; } hufftab_t; /* entry: leaf = (len<<8)|sym ; link = 0x8000 | (bits<<10)... see build */
static hufftab_t g_huff[16];
static int g_long_start[9][23], g_short_start[9][14];
static int g_init_done = 0;

/* two-level table: first level indexed by HUFF_L1_BITS bits.  leaf: bit15=0, bits 8..12 = code length, bits 0..7 = sym.
 * link: bit15=1, bits 10..14 = extra bits n (sub-table indexed by the next n bits), bits 0..9.. offset/4 is not enough for
 * big books, so links store an index into a side array instead. */
typedef struct { uint32_t off; uint8_t nbits; } hlink_t;
static hlink_t g_links[16][512];

static void build_huff(int slot) {
    int n = L3_HUFF_SIZE[slot - 1], off = L3_HUFF_OFF[slot - 1];
    const uint8_t* len = &L3_HUFF_LEN[off];
    const uint8_t* sym = &L3_HUFF_SYM[off];
    /* codes, left-aligned to 32 bits, in tree order */
    uint32_t* code = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t c = 0;
    for (int i = 0; i < n; i++) { code[i] = c; c += 1u << (32 - len[i]); }
    /* sub-table sizes per first-level prefix */
    int maxsub[512];
    memset(maxsub, 0, sizeof maxsub);
    for (int i = 0; i < n; i++)
        if (len[i] > HUFF_L1_BITS) {
            int p = code[i] >> (32 - HUFF_L1_BITS);
            if (len[i] - HUFF_L1_BITS > maxsub[p]) maxsub[p] = len[i] - HUFF_L1_BITS;
        }
    int total = 1 << HUFF_L1_BITS;
    for (int p = 0; p < 512; p++) if (maxsub[p]) { g_links[slot][p].off = total; g_links[slot][p].nbits = (uint8_t)maxsub[p]; total += 1 << maxsub[p]; }
    uint16_t* tab = (uint16_t*)calloc(total, sizeof(uint16_t));
    for (int i = 0; i < n; i++) {
        uint16_t leaf = (uint16_t)((len[i] << 8) | sym[i]);
        if (len[i] <= HUFF_L1_BITS) {
            int base = code[i] >> (32 - HUFF_L1_BITS), cnt = 1 << (HUFF_L1_BITS - len[i]);
            for (int k = 0; k < cnt; k++) tab[base + k] = leaf;
        } else {
            int p = code[i] >> (32 - HUFF_L1_BITS);
            int nb = g_links[slot][p].nbits;
            int sub = (code[i] << HUFF_L1_BITS) >> (32 - nb);
            int cnt = 1 << (nb - (len[i] - HUFF_L1_BITS));
            for (int k = 0; k < cnt; k++) tab[g_links[slot][p].off + sub + k] = leaf;
            tab[p] = 0x8000;
        }
    }
    g_huff[slot].tab = tab;
    g_huff[slot].size = total;
    free(code);
}

static void init_once(void) {
    if (g_init_done) return;
    for (int s = 1; s <= L3_NUM_HUFF; s++) build_huff(s);
    for (int r = 0; r < 9; r++) {
        int a = 0;
        for (int i = 0; i < 22; i++) { g_long_start[r][i] = a; a += L3_SFB_LONG_W[r][i]; }
        g_long_start[r][22] = a;
        a = 0;
        for (int i = 0; i < 13; i++) { g_short_start[r][i] = a; a += L3_SFB_SHORT_W[r][i]; }
        g_short_start[r][13] = a;
    }
    g_init_done = 1;
}

/* ------------------------------------------------------------------ bit reader (MSB first) */
typedef struct { const uint8_t* p; size_t nbytes; size_t pos; } bits_t; /* pos in bits */
static inline uint32_t peek32(const bits_t* b) {
    size_t by = b->pos >> 3;
    uint64_t v = 0;
    for (int i = 0; i < 5; i++) v = (v << 8) | (by + i < b->nbytes ? b->p[by + i] : 0);
    return (uint32_t)(v >> (8 - (b->pos & 7)));
}
static inline uint32_t getbits(bits_t* b, int n) {
    if (n == 0) return 0;
    uint32_t v = peek32(b) >> (32 - n);
    b->pos += n;
    return v;
}

/* ------------------------------------------------------------------ headers */
typedef struct { int crc, br_idx, sr_idx, pad, mode, mode_ext, sr, channels, frame_len, side_len, ver; } hdr_t;
static const int BR_V1_L3[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
static const int SR_V1[3] = {44100, 48000, 32000};

/* Valid Layer III header (any MPEG version so that LSF streams are recognised and reported as unsupported). */
static int parse_hdr(const uint8_t* p, hdr_t* h) {
    if (p[0] != 0xFF || (p[1] & 0xE0) != 0xE0) return 0;
    int ver = (p[1] >> 3) & 3;      /* 3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5, 1 = reserved */
    int layer = (p[1] >> 1) & 3;    /* 1 = Layer III */
    if (ver == 1 || layer != 1) return 0;
    h->ver = ver;
    h->crc = !(p[1] & 1);
    h->br_idx = p[2] >> 4;
    h->sr_idx = (p[2] >> 2) & 3;
    if (h->br_idx == 0 || h->br_idx == 15 || h->sr_idx == 3) return 0;
    h->pad = (p[2] >> 1) & 1;
    h->mode = p[3] >> 6;
    h->mode_ext = (p[3] >> 4) & 3;
    h->channels = h->mode == 3 ? 1 : 2;
    if (ver == 3) {
        h->sr = SR_V1[h->sr_idx];
        h->frame_len = 144 * BR_V1_L3[h->br_idx] * 1000 / h->sr + h->pad;
        h->side_len = h->channels == 1 ? 17 : 32;
    } else {
        static const int BR_V2[15] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};
        h->sr = SR_V1[h->sr_idx] >> (ver == 2 ? 1 : 2);
        h->frame_len = 72 * BR_V2[h->br_idx] * 1000 / h->sr + h->pad;
        h->side_len = h->channels == 1 ? 9 : 17;
    }
    return 1;
}
static int compatible(const hdr_t* a, const hdr_t* ref) {
    return a->ver == ref->ver && a->sr == ref->sr && a->channels == ref->channels;
}

static size_t skip_id3v2(const uint8_t* d, size_t len) {
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3' && d[3] != 0xFF && d[4] != 0xFF &&
        !(d[6] & 0x80) && !(d[7] & 0x80) && !(d[8] & 0x80) && !(d[9] & 0x80)) {
        size_t sz = ((size_t)d[6] << 21) | ((size_t)d[7] << 14) | ((size_t)d[8] << 7) | d[9];
        sz += 10;
        if (d[5] & 0x10) sz += 10; /* footer */
        return sz <= len ? sz : len;
    }
    return 0;
}

/* ------------------------------------------------------------------ side info */
typedef struct {
    int part2_3_length, big_values, global_gain, scalefac_compress, window_switching, block_type, mixed;
    int table_select[3], subblock_gain[3], region0_count, region1_count, preflag, scalefac_scale, count1table;
} gr_t;
typedef struct { int main_data_begin; int scfsi[2][4]; gr_t gr[2][2]; int valid; } side_t;

/* MPEG-1 (ISO 11172-3 2.4.1.7): 9-bit main_data_begin, scfsi, two granules, 4-bit scalefac_compress, preflag bit.
 * MPEG-2 / 2.5 "LSF" (ISO 13818-3 2.4.1.7): 8-bit main_data_begin, no scfsi, ONE granule, 9-bit scalefac_compress, no preflag bit
 * (preflag follows from scalefac_compress). */
static void parse_side(const uint8_t* p, int nch, int lsf, side_t* s) {
    bits_t b = {p, 32, 0};
    s->valid = 1;
    memset(s->scfsi, 0, sizeof s->scfsi);
    if (lsf) {
        s->main_data_begin = getbits(&b, 8);
        getbits(&b, nch == 1 ? 1 : 2);
    } else {
        s->main_data_begin = getbits(&b, 9);
        getbits(&b, nch == 1 ? 5 : 3);
        for (int ch = 0; ch < nch; ch++) for (int i = 0; i < 4; i++) s->scfsi[ch][i] = getbits(&b, 1);
    }
    for (int g = 0; g < (lsf ? 1 : 2); g++)
        for (int ch = 0; ch < nch; ch++) {
            gr_t* q = &s->gr[g][ch];
            q->part2_3_length = getbits(&b, 12);
            q->big_values = getbits(&b, 9);
            q->global_gain = getbits(&b, 8);
            q->scalefac_compress = getbits(&b, lsf ? 9 : 4);
            q->window_switching = getbits(&b, 1);
            if (q->window_switching) {
                q->block_type = getbits(&b, 2);
                q->mixed = getbits(&b, 1);
                q->table_select[0] = getbits(&b, 5);
                q->table_select[1] = getbits(&b, 5);
                q->table_select[2] = 0;
                for (int i = 0; i < 3; i++) q->subblock_gain[i] = getbits(&b, 3);
                q->region0_count = (q->block_type == 2 && !q->mixed) ? 8 : 7;
                q->region1_count = 20 - q->region0_count;
                if (q->block_type == 0) s->valid = 0;
            } else {
                for (int i = 0; i < 3; i++) q->table_select[i] = getbits(&b, 5);
                q->region0_count = getbits(&b, 4);
                q->region1_count = getbits(&b, 3);
                q->block_type = 0;
                q->mixed = 0;
                q->subblock_gain[0] = q->subblock_gain[1] = q->subblock_gain[2] = 0;
            }
            q->preflag = lsf ? (q->scalefac_compress >= 500) : (int)getbits(&b, 1);
            q->scalefac_scale = getbits(&b, 1);
            q->count1table = getbits(&b, 1);
            if (q->big_values > 288) s->valid = 0;
        }
}

/* ------------------------------------------------------------------ decoder state */
typedef struct {
    float prev[2][32][18];    /* IMDCT overlap */
    float ring[2][16][32];    /* last 16 DCT-32 vectors per channel, ring[head] newest */
    int head[2];
    int sf_l[2][2][22];       /* [gr][ch] long scalefactors */
    int sf_s[2][2][13][3];    /* [gr][ch] short scalefactors */
    int ill_l[22], ill_s[13][3]; /* MPEG-2 intensity stereo, right channel of the current frame: band carries the "illegal" is_pos */
} dec_t;

/* part 2: scalefactors (ISO 11172-3 2.4.2.7) */
static void read_scalefactors(bits_t* b, const side_t* s, int g, int ch, dec_t* d) {
    const gr_t* q = &s->gr[g][ch];
    int slen1 = L3_SLEN[0][q->scalefac_compress], slen2 = L3_SLEN[1][q->scalefac_compress];
    int* sl = d->sf_l[g][ch];
    int(*ss)[3] = d->sf_s[g][ch];
    memset(sl, 0, 22 * sizeof(int));
    memset(ss, 0, 13 * 3 * sizeof(int));
    if (q->block_type == 2) {
        int s0 = 0;
        if (q->mixed) {
            for (int sfb = 0; sfb < 8; sfb++) sl[sfb] = getbits(b, slen1);
            s0 = 3;
        }
        for (int sfb = s0; sfb < 6; sfb++) for (int w = 0; w < 3; w++) ss[sfb][w] = getbits(b, slen1);
        for (int sfb = 6; sfb < 12; sfb++) for (int w = 0; w < 3; w++) ss[sfb][w] = getbits(b, slen2);
    } else {
        static const int lo[4] = {0, 6, 11, 16}, hi[4] = {6, 11, 16, 21};
        for (int grp = 0; grp < 4; grp++) {
            int sl_bits = grp < 2 ? slen1 : slen2;
            for (int sfb = lo[grp]; sfb < hi[grp]; sfb++) {
                if (g == 1 && s->scfsi[ch][grp]) sl[sfb] = d->sf_l[0][ch][sfb];
                else sl[sfb] = getbits(b, sl_bits);
            }
        }
    }
}

/* part 2, MPEG-2 / 2.5 (ISO 13818-3 2.4.3.2): scalefac_compress -> four field widths and, through L3_LSF_NSF, how many
 * scalefactors use each; the values come in band order (long: sfb 0..20; short: sfb 0..11 x 3 windows; mixed: 6 long
 * bands, then short sfb 3..11 x 3 windows).  (Intensity-stereo right channels use other tables: such streams are rejected.) */
static void read_scalefactors_lsf(bits_t* b, const gr_t* q, int is_right, int* sl, int (*ss)[3], int* ill_l, int (*ill_s)[3]) {
    memset(sl, 0, 22 * sizeof(int));
    memset(ss, 0, 13 * 3 * sizeof(int));
    memset(ill_l, 0, 22 * sizeof(int));
    memset(ill_s, 0, 13 * 3 * sizeof(int));
    int sfc = q->scalefac_compress, slen[4] = {0, 0, 0, 0}, tbl;
    if (!is_right) {
        if (sfc < 400) { slen[0] = (sfc >> 4) / 5; slen[1] = (sfc >> 4) % 5; slen[2] = (sfc % 16) >> 2; slen[3] = sfc % 4; tbl = 0; }
        else if (sfc < 500) { sfc -= 400; slen[0] = (sfc >> 2) / 5; slen[1] = (sfc >> 2) % 5; slen[2] = sfc % 4; tbl = 1; }
        else { sfc -= 500; slen[0] = sfc / 3; slen[1] = sfc % 3; tbl = 2; }
    } else {   /* right channel of an intensity-stereo frame: the scalefactors are is_pos values, other partition tables */
        int isc = sfc >> 1;
        if (isc < 180) { slen[0] = isc / 36; slen[1] = (isc % 36) / 6; slen[2] = (isc % 36) % 6; tbl = 3; }
        else if (isc < 244) { isc -= 180; slen[0] = (isc % 64) >> 4; slen[1] = (isc % 16) >> 2; slen[2] = isc % 4; tbl = 4; }
        else { isc -= 244; slen[0] = isc / 3; slen[1] = isc % 3; tbl = 5; }
    }
    int cls = q->block_type != 2 ? 0 : (q->mixed ? 2 : 1);
    int v[40], il[40], n = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < L3_LSF_NSF[tbl][cls][i]; j++) {
            v[n] = getbits(b, slen[i]);
            /* the illegal intensity position (all ones): this band is not intensity coded.  A 0-bit field reads as position 0,
             * which FFmpeg and minimp3 take as a valid position (libmad as the illegal one); this restatement follows the former. */
            il[n] = is_right && slen[i] > 0 && v[n] == (1 << slen[i]) - 1;
            n++;
        }
    int k = 0;
    if (cls == 0) { for (int sfb = 0; sfb < 21 && k < n; sfb++) { sl[sfb] = v[k]; ill_l[sfb] = il[k]; k++; } }
    else {
        int s0 = 0;
        if (cls == 2) { for (int sfb = 0; sfb < 6; sfb++) { sl[sfb] = v[k]; ill_l[sfb] = il[k]; k++; } s0 = 3; }
        for (int sfb = s0; sfb < 12; sfb++) for (int w = 0; w < 3; w++) if (k < n) { ss[sfb][w] = v[k]; ill_s[sfb][w] = il[k]; k++; }
    }
}

/* part 3: Huffman.  Returns the number of lines decoded (start of the all-zero tail). */
static int read_huffman(bits_t* b, size_t end, const gr_t* q, int sr_row, int16_t* is) {
    memset(is, 0, 576 * sizeof(int16_t));
    int bv2 = q->big_values * 2;
    int r0, r1;
    /* window switching: region 0 = region0_count 7 (8 for pure short) bands, i.e. the first 8 long bands or the first 3 short
     * bands x 3 windows -- 36 lines at every MPEG-1 rate, 54 / 36 at the MPEG-2 rates, 108 / 72 at 8 kHz */
    if (q->window_switching) { r0 = q->block_type == 2 ? 3 * g_short_start[sr_row][3] : g_long_start[sr_row][8]; r1 = 576; }
    else {
        r0 = g_long_start[sr_row][q->region0_count + 1];
        int k = q->region0_count + q->region1_count + 2;
        r1 = g_long_start[sr_row][k > 22 ? 22 : k];
    }
    int i = 0;
    for (; i < bv2; i += 2) {
        int region = i < r0 ? 0 : (i < r1 ? 1 : 2);
        int slot = L3_TABSEL[q->table_select[region]][0], linbits = L3_TABSEL[q->table_select[region]][1];
        int x = 0, y = 0;
        if (slot) {
            uint32_t w = peek32(b);
            uint16_t e = g_huff[slot].tab[w >> (32 - HUFF_L1_BITS)];
            if (e & 0x8000) {
                const hlink_t* l = &g_links[slot][w >> (32 - HUFF_L1_BITS)];
                e = g_huff[slot].tab[l->off + ((w << HUFF_L1_BITS) >> (32 - l->nbits))];
            }
            b->pos += e >> 8;
            x = (e >> 4) & 15;
            y = e & 15;
            if (x == 15 && linbits) x += getbits(b, linbits);
            if (x && getbits(b, 1)) x = -x;
            if (y == 15 && linbits) y += getbits(b, linbits);
            if (y && getbits(b, 1)) y = -y;
        }
        is[i] = (int16_t)x;
        is[i + 1] = (int16_t)y;
    }
    if (b->pos > end) { /* big_values overran part2_3_length: treat the granule-channel as silence */
        memset(is, 0, 576 * sizeof(int16_t));
        b->pos = end;
        return 0;
    }
    const uint8_t* cb = L3_C1_BITS[q->count1table];
    const uint8_t* cl = L3_C1_LENS[q->count1table];
    while (b->pos < end && i <= 572) {
        uint32_t w = peek32(b);
        int v = -1;
        for (int k = 0; k < 16; k++)
            if ((w >> (32 - cl[k])) == cb[k]) { v = k; break; }
        b->pos += cl[v];
        int q4[4] = {(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1};
        for (int k = 0; k < 4; k++) if (q4[k] && getbits(b, 1)) q4[k] = -1;
        if (b->pos > end) break; /* quad straddles the end of the granule: discard it */
        for (int k = 0; k < 4; k++) is[i + k] = (int16_t)q4[k];
        i += 4;
    }
    b->pos = end;
    return i;
}

/* requantisation (ISO 2.4.3.4) in bitstream order */
static void requantize(const gr_t* q, int sr_row, const int* sf_l, int (*sf_s)[3], const int16_t* is, float* xr) {
    int mult = q->scalefac_scale ? 4 : 2;
    int i = 0;
    if (q->block_type != 2 || q->mixed) {
        int nl = 22;
        /* mixed: 8 long bands in MPEG-1, 6 in LSF (what the scalefactor partitions carry) = the lines below 36 at every rate
         * but 8 kHz, whose 6 long bands span 72 lines and meet short band 3 there (FFmpeg's compute_band_indexes: "For
         * 8000Hz, we handle the 72 first exponents as long blocks"; counting the bands below line 36 instead gave 3 at 8 kHz
         * and left the short part 36 lines early) */
        if (q->block_type == 2) nl = sr_row < 3 ? 8 : 6;
        for (int sfb = 0; sfb < nl; sfb++) {
            int e = q->global_gain - 210 - mult * (sf_l[sfb] + (q->preflag ? L3_PRETAB[sfb] : 0));
            float g = L3_GAIN[e - L3_GAIN_MIN];
            for (int k = 0; k < L3_SFB_LONG_W[sr_row][sfb]; k++, i++) {
                int v = is[i];
                float m = L3_POW43[v < 0 ? -v : v] * g;
                xr[i] = v < 0 ? -m : m;
            }
        }
    }
    if (q->block_type == 2) {
        for (int sfb = q->mixed ? 3 : 0; sfb < 13; sfb++) {
            int wdt = L3_SFB_SHORT_W[sr_row][sfb];
            for (int w = 0; w < 3; w++) {
                int e = q->global_gain - 210 - 8 * q->subblock_gain[w] - mult * sf_s[sfb][w];
                float g = L3_GAIN[e - L3_GAIN_MIN];
                for (int k = 0; k < wdt; k++, i++) {
                    int v = is[i];
                    float m = L3_POW43[v < 0 ? -v : v] * g;
                    xr[i] = v < 0 ? -m : m;
                }
            }
        }
    }
}

/* joint stereo in bitstream order (before the short-block reorder), ISO 11172-3 2.4.3.4.9.3.
 * MS: (m + s)/sqrt2, (m - s)/sqrt2.  Intensity: scalefactor bands of the RIGHT channel's partition that lie above its last
 * non-zero line (per window for short blocks) are intensity coded; the right channel's scalefactor there is is_pos, the
 * last band (long 21 / short 12) borrows the previous band's; is_pos 7 marks a band that is not intensity coded (it gets the
 * MS treatment if MS is on, like every band below the bound).  L' = L*is_l[is_pos], R' = L*is_r[is_pos], one multiply each. */
static void ms_band(float xr[2][576], int i0, int n) {
    for (int i = i0; i < i0 + n; i++) {
        float m = xr[0][i], sd = xr[1][i];
        xr[0][i] = (m + sd) * L3_INV_SQRT2;
        xr[1][i] = (m - sd) * L3_INV_SQRT2;
    }
}
/* MPEG-2 (ISO 13818-3 2.4.3.2): same band rule; the factors come from L3_IS_LSF[intensity_scale = scalefac_compress & 1] and a
 * band is excluded when it carries the illegal position (1 << slen) - 1 of its scalefactor field instead of is_pos 7. */
static void joint_stereo(const gr_t* q1, int sr_row, int ms, int intensity, const int* sf_l, int (*sf_s)[3], const int16_t* is1,
                         float xr[2][576], int lsf, const int* ill_l, int (*ill_s)[3]) {
    if (!intensity) {
        if (ms) ms_band(xr, 0, 576);
        return;
    }
    const float* kl = lsf ? L3_IS_LSF[q1->scalefac_compress & 1][0] : L3_IS_RATIO[0];
    const float* kr = lsf ? L3_IS_LSF[q1->scalefac_compress & 1][1] : L3_IS_RATIO[1];
    int short_start = 13, long_end = 22;
    if (q1->block_type == 2) { short_start = q1->mixed ? 3 : 0; long_end = q1->mixed ? (lsf ? 6 : 8) : 0; }
    int end = 576;
    int nz_short[3] = {0, 0, 0};
    for (int sfb = 12; sfb >= short_start; sfb--) {
        int wdt = L3_SFB_SHORT_W[sr_row][sfb];
        int src = sfb == 12 ? 11 : sfb;
        for (int w = 2; w >= 0; w--) {
            end -= wdt;
            int pos = -1;
            if (!nz_short[w]) {
                for (int j = 0; j < wdt; j++) if (is1[end + j] != 0) { nz_short[w] = 1; break; }
                if (!nz_short[w] && (lsf ? !ill_s[src][w] : sf_s[src][w] < 7)) pos = sf_s[src][w];
            }
            if (pos >= 0) {
                for (int j = end; j < end + wdt; j++) { float l = xr[0][j]; xr[0][j] = l * kl[pos]; xr[1][j] = l * kr[pos]; }
            } else if (ms) ms_band(xr, end, wdt);
        }
    }
    int nz = nz_short[0] | nz_short[1] | nz_short[2];
    for (int sfb = long_end - 1; sfb >= 0; sfb--) {
        int wdt = L3_SFB_LONG_W[sr_row][sfb];
        end -= wdt;
        int pos = -1;
        if (!nz) {
            for (int j = 0; j < wdt; j++) if (is1[end + j] != 0) { nz = 1; break; }
            int src = sfb == 21 ? 20 : sfb;
            if (!nz && (lsf ? !ill_l[src] : sf_l[src] < 7)) pos = sf_l[src];
        }
        if (pos >= 0) {
            for (int j = end; j < end + wdt; j++) { float l = xr[0][j]; xr[0][j] = l * kl[pos]; xr[1][j] = l * kr[pos]; }
        } else if (ms) ms_band(xr, end, wdt);
    }
}

/* short-block reorder [sfb][win][k] -> [sfb][k][win] */
static void reorder(const gr_t* q, int sr_row, float* xr) {
    float tmp[576];
    memcpy(tmp, xr, sizeof tmp);
    for (int sfb = q->mixed ? 3 : 0; sfb < 13; sfb++) {
        int st = g_short_start[sr_row][sfb], wdt = L3_SFB_SHORT_W[sr_row][sfb];
        for (int w = 0; w < 3; w++)
            for (int k = 0; k < wdt; k++) xr[3 * st + 3 * k + w] = tmp[3 * st + w * wdt + k];
    }
}

static void alias_reduce(const gr_t* q, float* xr) {
    int nb = q->block_type == 2 ? (q->mixed ? 1 : 0) : 31;
    for (int sb = 1; sb <= nb; sb++)
        for (int i = 0; i < 8; i++) {
            float lo = xr[18 * sb - 1 - i], hi = xr[18 * sb + i];
            xr[18 * sb - 1 - i] = lo * L3_ALIAS_CS[i] - hi * L3_ALIAS_CA[i];
            xr[18 * sb + i] = hi * L3_ALIAS_CS[i] + lo * L3_ALIAS_CA[i];
        }
}

/* IMDCT + window + overlap-add + frequency inversion for one subband; out[t] = subband sample at time slot t */
/* 3-point DFT of complex numbers (W3 = e^{-2 pi i/3}): X0 = x0 + s, X1 = m - i q, X2 = m + i q with s = x1 + x2,
 * m = x0 - s/2 (one fmaf per component), q = (sqrt3/2)(x1 - x2).  Every operation rounds separately. */
static void dft3(const float* xr, const float* xi, float* yr, float* yi) {
    float sr = xr[1] + xr[2], si = xi[1] + xi[2], dr = xr[1] - xr[2], di = xi[1] - xi[2];
    yr[0] = xr[0] + sr; yi[0] = xi[0] + si;
    float mr = fmaf(-0.5f, sr, xr[0]), mi = fmaf(-0.5f, si, xi[0]);
    float qr = L3_SQRT3_2 * dr, qi = L3_SQRT3_2 * di;
    yr[1] = mr + qi; yi[1] = mi - qr;
    yr[2] = mr - qi; yi[2] = mi + qr;
}

/* 18-point DCT-IV u[m] = sum_k X[k] cos(pi/72 (2m+1)(2k+1)) through a 9-point complex FFT (3 x 3 Cooley-Tukey):
 *   v[n] = (X[2n] + i X[17-2n]) e^{-i pi (4n+1)/72},  V = FFT9(v),  y[k] = V[k] e^{-i pi k/18},  u[2k] = Re y[k], u[17-2k] = -Im y[k].
 * A complex product (a + i b)(c + i d) is evaluated as (fmaf(-b, d, a*c), fmaf(a, d, b*c)). */
static void dct4_18(const float* X, float* u) {
    float vr[9], vi[9];
    for (int n = 0; n < 9; n++) {
        float a = X[2 * n], b = X[17 - 2 * n], c = L3_D4_PRE[0][n], s = L3_D4_PRE[1][n];
        vr[n] = fmaf(b, s, a * c);      /* (a + i b)(c - i s) */
        vi[n] = fmaf(-a, s, b * c);
    }
    float tr[3][3], ti[3][3];           /* [b][c]: 3-point DFTs over a of v[3a + b], then times W9^{bc} */
    for (int b = 0; b < 3; b++) {
        float xr[3] = {vr[b], vr[3 + b], vr[6 + b]}, xi[3] = {vi[b], vi[3 + b], vi[6 + b]};
        dft3(xr, xi, tr[b], ti[b]);
    }
    static const int tw[3][3] = {{-1, -1, -1}, {-1, 0, 1}, {-1, 1, 2}};   /* index into L3_D4_W9 (exponents 1, 2, 4), -1 = 1 */
    for (int b = 1; b < 3; b++)
        for (int c = 1; c < 3; c++) {
            float wr = L3_D4_W9[0][tw[b][c]], wi = L3_D4_W9[1][tw[b][c]], zr = tr[b][c], zi = ti[b][c];
            tr[b][c] = fmaf(-zi, wi, zr * wr);
            ti[b][c] = fmaf(zr, wi, zi * wr);
        }
    for (int c = 0; c < 3; c++) {
        float xr[3] = {tr[0][c], tr[1][c], tr[2][c]}, xi[3] = {ti[0][c], ti[1][c], ti[2][c]}, yr[3], yi[3];
        dft3(xr, xi, yr, yi);
        for (int d = 0; d < 3; d++) {
            int k = c + 3 * d;
            float re = yr[d], im = yi[d];
            if (k) {                     /* y = V (pc - i ps) */
                float pc = L3_D4_POST[0][k], ps = L3_D4_POST[1][k];
                re = fmaf(yi[d], ps, yr[d] * pc);
                im = fmaf(-yr[d], ps, yi[d] * pc);
            }
            u[2 * k] = re;
            u[17 - 2 * k] = -im;
        }
    }
}

static void hybrid_sb(const float* X, int bt, float* prev, int sb, float* out) {
    float raw[36];
    if (bt != 2) {
        float u[18];
        dct4_18(X, u);
        for (int i = 0; i < 36; i++) raw[i] = u[L3_IMDCT36_IDX[i]] * L3_WINSGN[bt][i];
    } else {
        for (int i = 0; i < 36; i++) raw[i] = 0.0f;
        for (int w = 0; w < 3; w++)
            for (int i = 0; i < 12; i++) {
                float a = 0.0f;
                for (int k = 0; k < 6; k++) a = fmaf(X[3 * k + w], L3_IMDCT12[i][k], a);
                raw[6 + 6 * w + i] = raw[6 + 6 * w + i] + a * L3_WINSGN[2][i];
            }
    }
    for (int i = 0; i < 18; i++) {
        float v = raw[i] + prev[i];
        prev[i] = raw[18 + i];
        out[i] = ((sb & 1) && (i & 1)) ? -v : v;
    }
}

/* Lee's 32-point DCT-II: B[j] = sum_k S[k] cos((2k+1) j pi / 64) */
static void dct2_rec(float* x, int n, const float* sec) {
    if (n == 1) return;
    int h = n / 2;
    float g[16], hh[16];
    for (int i = 0; i < h; i++) {
        g[i] = x[i] + x[n - 1 - i];
        hh[i] = (x[i] - x[n - 1 - i]) * sec[i];
    }
    dct2_rec(g, h, sec + h);
    dct2_rec(hh, h, sec + h);
    for (int k = 0; k < h; k++) {
        x[2 * k] = g[k];
        x[2 * k + 1] = (k + 1 < h) ? hh[k] + hh[k + 1] : hh[k];
    }
}

/* polyphase synthesis of one time slot: 32 subband samples -> 32 PCM samples */
static void synth_slot(dec_t* d, int ch, const float* S, float* pcm) {
    d->head[ch] = (d->head[ch] + 15) & 15;
    float* B = d->ring[ch][d->head[ch]];
    memcpy(B, S, 32 * sizeof(float));
    dct2_rec(B, 32, L3_DCT_SEC);
    for (int j = 0; j < 32; j++) {
        float a = 0.0f;
        for (int p = 0; p < 8; p++) {
            const float* Ba = d->ring[ch][(d->head[ch] + 2 * p) & 15];
            const float* Bb = d->ring[ch][(d->head[ch] + 2 * p + 1) & 15];
            a = fmaf(Ba[L3_SYNTH_IA[j]], L3_SYNTH_WA[p][j], a);
            a = fmaf(Bb[L3_SYNTH_IB[j]], L3_SYNTH_WB[p][j], a);
        }
        pcm[j] = a;
    }
}

/* ------------------------------------------------------------------ frame walk shared by decode/analyze */
typedef struct { size_t pos; hdr_t h; } frame_ref;

static int is_info_frame(const uint8_t* p, const hdr_t* h) {
    size_t o = 4 + (h->crc ? 2 : 0) + h->side_len;
    if (o + 4 <= (size_t)h->frame_len && (!memcmp(p + o, "Xing", 4) || !memcmp(p + o, "Info", 4))) return 1;
    if (36 + 4 <= h->frame_len && !memcmp(p + 36, "VBRI", 4)) return 1;
    return 0;
}

typedef void (*pcm_sink)(void* ctx, const float* l, const float* r, int n);

static int decode_stream(const uint8_t* data, size_t len, uint32_t flags, pcm_sink sink, void* ctx, mp3o_info* info) {
    init_once();
    memset(info, 0, sizeof *info);
    size_t pos = skip_id3v2(data, len);
    hdr_t ref;
    int have_ref = 0, in_sync = 0, first = 1;
    dec_t* d = (dec_t*)calloc(1, sizeof(dec_t));
    uint8_t* md = (uint8_t*)malloc(len + 64);
    size_t md_len = 0;
    static _Thread_local float xr[2][576], sbs[2][32][18], pcm[2][1152];
    static _Thread_local int16_t is[2][576];
    int status = MP3O_ERR_PROBE;

    while (pos + 4 <= len) {
        hdr_t h;
        int ok = parse_hdr(data + pos, &h) && (!have_ref || compatible(&h, &ref));
        if (ok && pos + (size_t)h.frame_len > len) break; /* truncated final frame = EOF */
        if (ok && !in_sync) {                             /* (re)sync needs the next header to confirm */
            size_t nx = pos + h.frame_len;
            hdr_t h2;
            if (nx + 4 <= len && !(parse_hdr(data + nx, &h2) && compatible(&h2, &h))) ok = 0;
        }
        if (!ok) { in_sync = 0; pos++; continue; }
        in_sync = 1;
        if (!have_ref) {
            ref = h; have_ref = 1;
            info->sample_rate = h.sr; info->channels = h.channels;
            status = 0;
        }
        const uint8_t* fp = data + pos;
        pos += h.frame_len;
        if (first) {
            first = 0;
            if (!(flags & MP3O_FLAG_KEEP_INFO_FRAME) && is_info_frame(fp, &h)) { info->skipped_info_frame = 1; continue; }
        }
        int nch = h.channels;
        size_t so = 4 + (h.crc ? 2 : 0);
        const int lsf = h.ver != 3, ngr = lsf ? 1 : 2;
        side_t s;
        parse_side(fp + so, nch, lsf, &s);
        size_t mdo = so + h.side_len;
        size_t md_start_ok = s.main_data_begin <= (int)md_len;
        size_t start = md_len - (md_start_ok ? s.main_data_begin : 0);
        if ((size_t)h.frame_len > mdo) { memcpy(md + md_len, fp + mdo, h.frame_len - mdo); md_len += h.frame_len - mdo; }
        if (!s.valid || !md_start_ok) { info->n_dropped_frames++; continue; }
        memset(md + md_len, 0, 32);

        /* rows of the band tables: 44100, 48000, 32000 | 22050, 24000, 16000 | 11025, 12000, 8000 */
        int sr_row = (h.ver == 3 ? 0 : (h.ver == 2 ? 3 : 6)) + h.sr_idx;
        bits_t b = {md, md_len + 32, start * 8};
        if (lsf && nch == 2 && h.mode == 1 && (h.mode_ext & 1)) s.gr[0][1].preflag = 0;   /* intensity right channel: no preflag */
        for (int g = 0; g < ngr; g++) {
            int nz[2] = {0, 0};
            for (int ch = 0; ch < nch; ch++) {
                const gr_t* q = &s.gr[g][ch];
                size_t end = b.pos + q->part2_3_length;
                if (lsf) read_scalefactors_lsf(&b, q, ch == 1 && h.mode == 1 && (h.mode_ext & 1), d->sf_l[g][ch], d->sf_s[g][ch], d->ill_l, d->ill_s);
                else read_scalefactors(&b, &s, g, ch, d);
                if (b.pos > end) { memset(is[ch], 0, sizeof is[ch]); b.pos = end; nz[ch] = 0; }
                else nz[ch] = read_huffman(&b, end, q, sr_row, is[ch]);
                requantize(q, sr_row, d->sf_l[g][ch], d->sf_s[g][ch], is[ch], xr[ch]);
            }
            if (nch == 2 && h.mode == 1 && h.mode_ext)
                joint_stereo(&s.gr[g][1], sr_row, h.mode_ext & 2, h.mode_ext & 1, d->sf_l[g][1], d->sf_s[g][1], is[1], xr, lsf, d->ill_l, d->ill_s);
            for (int ch = 0; ch < nch; ch++) {
                const gr_t* q = &s.gr[g][ch];
                if (q->block_type == 2) reorder(q, sr_row, xr[ch]);
                alias_reduce(q, xr[ch]);
                for (int sb = 0; sb < 32; sb++) {
                    int bt = (q->block_type == 2 && q->mixed && sb < 2) ? 0 : q->block_type;
                    hybrid_sb(&xr[ch][18 * sb], bt, d->prev[ch][sb], sb, sbs[ch][sb]);
                }
                for (int t = 0; t < 18; t++) {
                    float S[32];
                    for (int sb = 0; sb < 32; sb++) S[sb] = sbs[ch][sb][t];
                    synth_slot(d, ch, S, &pcm[ch][576 * g + 32 * t]);
                }
            }
        }
        if (flags & MP3O_FLAG_CLAMP_PCM)
            for (int ch = 0; ch < nch; ch++)
                for (int i = 0; i < 576 * ngr; i++) pcm[ch][i] = pcm[ch][i] > 1.0f ? 1.0f : (pcm[ch][i] < -1.0f ? -1.0f : pcm[ch][i]);
        sink(ctx, pcm[0], nch == 2 ? pcm[1] : NULL, 576 * ngr);
        info->n_frames++;
        info->n_samples += 576 * ngr;
    }
    free(md);
    free(d);
    info->status = status;
    return status;
}

/* ------------------------------------------------------------------ frame index only (the demux half of decode_stream)
 * The same walk as decode_stream, statement for statement, without the decoding: which frames become audio, where each
 * frame's granule data starts in the concatenated main-data stream (the bit reservoir unrolled) and what that stream holds.
 * Checker for the device's speculative frame index (k_frame_index.cuh) and its CPU model (tests/dev/k0_index_model.c). */
int mp3o_index(const uint8_t* data, size_t len, uint32_t flags, mp3o_frame* recs, size_t max_recs, uint8_t* md, size_t* md_len_out,
               mp3o_info* info) {
    init_once();
    memset(info, 0, sizeof *info);
    size_t pos = skip_id3v2(data, len);
    hdr_t ref;
    int have_ref = 0, in_sync = 0, first = 1;
    size_t md_len = 0, n = 0;
    int status = MP3O_ERR_PROBE;
    while (pos + 4 <= len) {
        hdr_t h;
        int ok = parse_hdr(data + pos, &h) && (!have_ref || compatible(&h, &ref));
        if (ok && pos + (size_t)h.frame_len > len) break;
        if (ok && !in_sync) {
            size_t nx = pos + h.frame_len;
            hdr_t h2;
            if (nx + 4 <= len && !(parse_hdr(data + nx, &h2) && compatible(&h2, &h))) ok = 0;
        }
        if (!ok) { in_sync = 0; pos++; continue; }
        in_sync = 1;
        if (!have_ref) { ref = h; have_ref = 1; info->sample_rate = h.sr; info->channels = h.channels; status = 0; }
        const uint8_t* fp = data + pos;
        const size_t fpos = pos;
        pos += h.frame_len;
        if (first) {
            first = 0;
            if (!(flags & MP3O_FLAG_KEEP_INFO_FRAME) && is_info_frame(fp, &h)) { info->skipped_info_frame = 1; continue; }
        }
        size_t so = 4 + (h.crc ? 2 : 0);
        side_t s;
        parse_side(fp + so, h.channels, h.ver != 3, &s);
        size_t mdo = so + h.side_len;
        size_t md_start_ok = s.main_data_begin <= (int)md_len;
        size_t start = md_len - (md_start_ok ? s.main_data_begin : 0);
        if ((size_t)h.frame_len > mdo) { memcpy(md + md_len, fp + mdo, h.frame_len - mdo); md_len += h.frame_len - mdo; }
        if (!s.valid || !md_start_ok) { info->n_dropped_frames++; continue; }
        if (n < max_recs) {
            recs[n].file_off = (uint32_t)fpos; recs[n].md_start = (uint32_t)start; recs[n].md_end = (uint32_t)md_len;
            recs[n].hdr = ((uint32_t)fp[0] << 24) | ((uint32_t)fp[1] << 16) | ((uint32_t)fp[2] << 8) | fp[3];
        }
        n++;
        info->n_frames++;
    }
    *md_len_out = md_len;
    info->status = status;
    return status;
}

/* ------------------------------------------------------------------ public entry points */
typedef struct { float* pcm; size_t max_frames, n; } copy_ctx;
static void sink_copy(void* c, const float* l, const float* r, int n) {
    copy_ctx* k = (copy_ctx*)c;
    if (k->n + n <= k->max_frames) {
        memcpy(k->pcm + k->n, l, n * sizeof(float));
        if (r) memcpy(k->pcm + k->max_frames + k->n, r, n * sizeof(float));
    }
    k->n += n;
}
int mp3o_decode(const uint8_t* data, size_t len, uint32_t flags, float* pcm, size_t max_frames, mp3o_info* info) {
    copy_ctx c = {pcm, max_frames, 0};
    return decode_stream(data, len, flags, sink_copy, &c, info);
}

typedef struct { rgo_state* st; int inited; unsigned sr; int nch; int bad_rate; const mp3o_info* info; } an_ctx;
static void sink_analyze(void* c, const float* l, const float* r, int n) {
    an_ctx* a = (an_ctx*)c;
    if (!a->inited) {
        a->inited = 1;
        if (rgo_state_init(a->st, a->info->sample_rate, a->info->channels) != 0) a->bad_rate = 1;
    }
    if (!a->bad_rate) rgo_state_feed_f32(a->st, l, r, (size_t)n);
}
int mp3o_analyze(const uint8_t* data, size_t len, uint32_t flags, rgo_result* out, uint32_t* hist_out, mp3o_info* info) {
    rgo_state* st = (rgo_state*)calloc(1, sizeof(rgo_state));
    an_ctx a = {st, 0, 0, 0, 0, info};
    int rc = decode_stream(data, len, flags, sink_analyze, &a, info);
    if (rc == 0) {
        if (!a.inited) rgo_state_init(st, info->sample_rate, info->channels); /* zero decodable frames: empty histogram */
        rgo_state_finish(st, out);
        if (hist_out) memcpy(hist_out, st->hist, sizeof st->hist);
    }
    free(st);
    return rc;
}
