/*
 * gf_kernel_model.c -- TEST INFRASTRUCTURE: a sequential C model of the ALGORITHM of mp3rgain_b200/csrc/k_gain_fields.cuh
 * (speculative walk of 4 KiB sub-segments, merge by entry position, patch from the entry positions), statement by statement
 * the same arithmetic as the kernels (16-byte SCAN step with the carry-free byte tricks, frame-length LUT, funnel-shift
 * fetches, in_range / bound handling) minus the warp machinery.  tests/test_gain_model.py runs it against
 * oracle/gain_fields_oracle.c on a few hundred thousand random, corrupted and adversarial images -- far more than the GPU
 * parity tests can afford -- so that the exactness argument of the speculation is checked at scale and a
 * future edit of the kernel's logic can be tried here first.  Bytes beyond `len` are filled by the caller with garbage: the
 * walk may load them but must never let them influence a result.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GF_SUB 4096u
#define GF_NONE 0xFFFFFFFFu

typedef struct { uint32_t first_jump, end_pos, n_frames, gsum, gcount, gminmax, first_hdr; } seg_t;
typedef struct { uint64_t frame_count, gain_sum, gain_count; uint32_t first_hdr, gminmax, rewalks; } track_t;

/* k0_hdr_valid / k0_frame_len of k_frame_index.cuh */
static int hdr_valid(uint32_t h) {
    if ((h & 0xFFE00000u) != 0xFFE00000u) return 0;
    const int ver = (h >> 19) & 3, layer = (h >> 17) & 3, br = (h >> 12) & 15, sr = (h >> 10) & 3;
    return ver != 1 && layer == 1 && br != 0 && br != 15 && sr != 3;
}
static int frame_len(uint32_t h) {
    const int ver = (h >> 19) & 3, br_idx = (h >> 12) & 15, sr_idx = (h >> 10) & 3, pad = (h >> 9) & 1;
    const int sr1 = sr_idx == 0 ? 44100 : (sr_idx == 1 ? 48000 : 32000);
    if (ver == 3) {
        const int kbps = (br_idx < 5) ? 24 + 8 * br_idx : (br_idx < 9) ? 16 * br_idx - 16 : (br_idx < 13) ? 32 * br_idx - 160 : (br_idx == 13 ? 256 : 320);
        return 144 * kbps * 1000 / sr1 + pad;
    }
    const int kbps = (br_idx < 9) ? 8 * br_idx : (br_idx < 13 ? 16 * br_idx - 64 : 144 + 16 * (br_idx - 13));
    return 72 * kbps * 1000 / (sr1 >> (ver == 2 ? 1 : 2)) + pad;
}
static uint16_t g_lut[256];
static int g_lut_ready = 0;
static uint32_t lut_index(uint32_t h) { return ((h >> 13) & 0xC0u) | ((h >> 10) & 0x3Fu); }
static void fill_lut(void) {
    for (uint32_t i = 0; i < 256; i++) {
        const uint32_t ver = i >> 6, br = (i >> 2) & 15, sr = i & 3;
        const uint32_t h = 0xFFE20000u | (ver << 19) | (br << 12) | (sr << 10);
        g_lut[i] = hdr_valid(h) ? (uint16_t)frame_len(h) : (uint16_t)0;
    }
    g_lut_ready = 1;
}

/* gf_bounds */
static void bounds(const uint8_t* d, uint32_t len, uint32_t* a0, uint32_t* audio_end) {
    *a0 = 0;
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3')
        *a0 = 10 + (((uint32_t)(d[6] & 0x7F) << 21) | ((uint32_t)(d[7] & 0x7F) << 14) | ((uint32_t)(d[8] & 0x7F) << 7) | (uint32_t)(d[9] & 0x7F));
    uint32_t end = len;
    if (end >= 128 && d[end - 128] == 'T' && d[end - 127] == 'A' && d[end - 126] == 'G') end -= 128;
    if (end >= 32) {
        const uint8_t* f = d + end - 32;
        if (memcmp(f, "APETAGEX", 8) == 0) {
            const uint64_t tag_size = (uint64_t)f[12] | ((uint64_t)f[13] << 8) | ((uint64_t)f[14] << 16) | ((uint64_t)f[15] << 24);
            const uint64_t header = (f[23] & 0x80) ? 32 : 0;
            if ((uint64_t)end >= tag_size + header) end = (uint32_t)(end - tag_size - header);
        }
    }
    *audio_end = end;
}

static uint32_t adjust(uint32_t cur, int steps, int wrap) {
    if (!wrap) {
        if (steps > 0) { const uint32_t a = (uint32_t)(steps < 255 ? steps : 255); return cur + a > 255u ? 255u : cur + a; }
        const long long n = -(long long)steps;
        const int a = (int)(n < 255 ? n : 255);
        return (uint32_t)((int)cur - a < 0 ? 0 : (int)cur - a);
    }
    const int ng = (int)(((long long)cur + steps) % 256);
    return (uint32_t)((ng + 256) % 256);
}

static uint32_t ldw(const uint8_t* img, uint32_t a) { uint32_t v; memcpy(&v, img + a, 4); return v; }   /* little-endian host */
static uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static uint32_t ld32(const uint8_t* img, uint32_t g) { const uint32_t a = g & ~3u; return funnel_r(ldw(img, a), ldw(img, a + 4), 8 * (g & 3)); }
static uint32_t bswap(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }
static uint32_t ffs32(uint32_t m) { uint32_t i = 0; while (!((m >> i) & 1)) i++; return i + 1; }

typedef struct { const uint8_t* img; uint32_t len, audio_end, bound; uint32_t h, next; } ctx_t;
static int in_range(const ctx_t* c, uint32_t p) { return (uint64_t)p + 4 <= c->audio_end && p < c->bound; }
static int check(ctx_t* c, uint32_t p) {
    c->h = bswap(ld32(c->img, p));
    const uint32_t f = g_lut[lut_index(c->h)];
    if ((c->h & 0xFFE60000u) != 0xFFE20000u || f == 0) return 0;
    c->next = p + f + ((c->h >> 9) & 1);
    if (c->next + 2 <= c->audio_end) return (ld32(c->img, c->next) & 0xE0FFu) == 0xE0FFu;
    return c->next <= c->audio_end;
}

/* gf_walk<PATCH>: reads img only; when patch != 0 the rewritten gain fields go to outp (a copy of the image) */
static seg_t walk(const uint8_t* img, uint8_t* outp, uint32_t len, uint32_t audio_end, uint32_t pos, uint32_t bound, int patch, int steps, int wrap, int channel, int run) {
    seg_t o;
    o.first_jump = GF_NONE; o.n_frames = 0; o.gcount = 0; o.first_hdr = 0;
    uint32_t gsum = 0, gmin = 255, gmax = 0;
    ctx_t c; c.img = img; c.len = len; c.audio_end = audio_end; c.bound = bound; c.h = 0; c.next = 0;
    int st = (run && in_range(&c, pos)) ? 0 : 2;   /* 0 SCAN, 1 FRAME, 2 DONE */
    while (st != 2) {
        if (st == 0) {
            const uint32_t a = pos & ~15u;
            const uint32_t w[5] = {ldw(img, a), ldw(img, a + 4), ldw(img, a + 8), ldw(img, a + 12), ldw(img, a + 16)};
            uint32_t m = 0;
            for (int k = 0; k < 4; k++) {
                const uint32_t ff = ((w[k] & 0x7F7F7F7Fu) + 0x01010101u) & w[k];
                const uint32_t u = funnel_r(w[k], w[k + 1], 8);
                const uint32_t cc = (ff & u & (u << 1) & (u << 2) & 0x80808080u) >> 7;
                m |= ((cc * 0x01020408u) >> 24) << (4 * k);
            }
            m &= ~((1u << (pos & 15u)) - 1u);
            if (!m) pos = a + 16;
            else {
                pos = a + (ffs32(m) - 1);
                if (in_range(&c, pos)) {
                    if (check(&c, pos)) st = 1;
                    else pos++;
                }
            }
            if (st == 0 && !in_range(&c, pos)) st = 2;
        } else {
            const uint32_t h = c.h;
            if (o.first_jump == GF_NONE) o.first_jump = pos;
            const int mono = ((h >> 6) & 3) == 3, lsf = ((h >> 19) & 3) != 3;
            const uint32_t so = (h & 0x10000u) ? 4 : 6;
            const uint32_t x = pos + so + (lsf ? (mono ? 9 : 17) : (mono ? 17 : 32));
            uint32_t mk = 0;
            if (x + 4 <= len) mk = ld32(img, x);
            if (!(x + 4 <= len && (mk == 0x676E6958u || mk == 0x6F666E49u))) {
                const int nch = mono ? 1 : 2, n = lsf ? nch : 2 * nch;
                const uint32_t bit0 = (lsf ? (mono ? 9u : 10u) : (mono ? 18u : 20u)) + 21u, per = lsf ? 63u : 59u;
                for (int k = 0; k < 4; k++) {
                    if (k < n) {
                        const uint32_t bit = bit0 + k * per;
                        const uint32_t idx = pos + so + (bit >> 3), sh = bit & 7;
                        uint32_t g = 0, b0 = 0, b1 = 0;
                        if (idx < len) {
                            const uint32_t v = ld32(img, idx);
                            b0 = v & 0xFF;
                            if (sh == 0) g = b0;
                            else if (idx + 1 < len) { b1 = (v >> 8) & 0xFF; g = ((b0 << sh) | (b1 >> (8 - sh))) & 0xFF; }
                            else g = (b0 << sh) & 0xFF;
                        }
                        gsum += g; if (g < gmin) gmin = g; if (g > gmax) gmax = g;
                        if (patch && idx < len && (channel < 0 || (k >= channel && (k - channel) % nch == 0))) {
                            const uint32_t v = adjust(g, steps, channel < 0 ? wrap : 0);
                            if (sh == 0) outp[idx] = (uint8_t)v;
                            else {
                                outp[idx] = (uint8_t)((b0 & (0xFFu << (8 - sh))) | (v >> sh));
                                if (idx + 1 < len) outp[idx + 1] = (uint8_t)((b1 & (0xFFu >> sh)) | ((v << (8 - sh)) & 0xFF));
                            }
                        }
                    }
                }
                if (o.n_frames == 0) o.first_hdr = h;
                o.n_frames++;
                o.gcount += (uint32_t)n;
            }
            pos = c.next;
            if (!in_range(&c, pos)) st = 2;
            else if (!check(&c, pos)) { pos++; st = in_range(&c, pos) ? 0 : 2; }
        }
    }
    o.end_pos = pos;
    o.gsum = gsum;
    o.gminmax = gmin | (gmax << 8);
    return o;
}

static void seg_range(uint32_t seg, uint32_t n_seg, uint32_t* b, uint32_t* e) {
    *b = seg * GF_SUB;
    *e = (seg + 1 == n_seg) ? 0xFFFFFFFFu : (seg + 1) * GF_SUB;
}

/* k_gf_scan + k_gf_merge (+ k_gf_patch when patch != 0) over one image.  `img` must have >= 32 readable bytes beyond len. */
int gfm_run(uint8_t* img, uint32_t len, int patch, int steps, int wrap, int channel, track_t* out) {
    if (!g_lut_ready) fill_lut();
    const uint32_t n_seg = len ? (len + GF_SUB - 1) / GF_SUB : 1;
    uint32_t a0, audio_end;
    bounds(img, len, &a0, &audio_end);
    seg_t* segs = (seg_t*)malloc(sizeof(seg_t) * n_seg);
    uint32_t* entry = (uint32_t*)malloc(sizeof(uint32_t) * n_seg);
    for (uint32_t s = 0; s < n_seg; s++) {          /* k_gf_scan */
        uint32_t b, e;
        seg_range(s, n_seg, &b, &e);
        if (b < a0) b = a0;
        segs[s] = walk(img, NULL, len, audio_end, b, e, 0, 0, 0, -1, b < e);
    }
    memset(out, 0, sizeof *out);
    uint32_t P = a0, gmin = 255, gmax = 0, first_s = GF_NONE;
    for (uint32_t s = 0; s < n_seg; s++) {          /* k_gf_merge */
        uint32_t b, e;
        seg_range(s, n_seg, &b, &e);
        entry[s] = GF_NONE;
        const int active = P < e && (uint64_t)P + 4 <= audio_end;
        if (!active) continue;
        seg_t o = segs[s];
        if (!(o.first_jump == GF_NONE || P <= o.first_jump)) { o = walk(img, NULL, len, audio_end, P, e, 0, 0, 0, -1, 1); out->rewalks++; }
        entry[s] = P;
        if (o.n_frames) {
            if (first_s == GF_NONE) { first_s = s; out->first_hdr = o.first_hdr; }
            out->frame_count += o.n_frames; out->gain_sum += o.gsum; out->gain_count += o.gcount;
            if ((o.gminmax & 0xFF) < gmin) gmin = o.gminmax & 0xFF;
            if ((o.gminmax >> 8) > gmax) gmax = o.gminmax >> 8;
        }
        P = o.end_pos;
    }
    out->gminmax = gmin | (gmax << 8);
    if (patch) {                                     /* k_gf_patch: any order -- here last sub-segment first; original in, copy out */
        uint8_t* copy = (uint8_t*)malloc(len ? len : 1);
        memcpy(copy, img, len);
        for (uint32_t k = n_seg; k-- > 0;) {
            uint32_t b, e;
            seg_range(k, n_seg, &b, &e);
            walk(img, copy, len, audio_end, entry[k], e, 1, steps, wrap, channel, entry[k] != GF_NONE);
        }
        memcpy(img, copy, len);
        free(copy);
    }
    free(segs);
    free(entry);
    return 0;
}
