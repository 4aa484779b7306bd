/*
 * gain_fields_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) for the reference's first-party frame walk
 * over the global_gain fields: mp3rgain::analyze() and apply_gain*() (M-Igashi/mp3rgain v1.5.0, src/lib.rs).
 *
 * Every function restates the reference function it cites, byte for byte.  Pinned by the reference's own integration
 * tests (tests/integration_tests.rs:32-80, 88-200, 416-428) run against this restatement on the reference's four fixtures
 * (tests/test_oracle_gain_fields.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 */
#include "gain_fields_oracle.h"

#include <string.h>

typedef struct {
    int version;      /* 0 = MPEG1, 1 = MPEG2, 2 = MPEG2.5 */
    int has_crc;
    int channel_mode; /* 0 Stereo, 1 Joint Stereo, 2 Dual Channel, 3 Mono */
    size_t frame_size;
} hdr_t;

/* parse_header, src/lib.rs:169-252 */
static int parse_header(const uint8_t* h, size_t avail, hdr_t* out) {
    static const unsigned br1[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
    static const unsigned br2[15] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};
    static const unsigned sr[3][3] = {{44100, 48000, 32000}, {22050, 24000, 16000}, {11025, 12000, 8000}};
    if (avail < 4) return 0;
    if (h[0] != 0xFF || (h[1] & 0xE0) != 0xE0) return 0;
    const unsigned vb = (h[1] >> 3) & 3;
    int version;
    if (vb == 0) version = 2; else if (vb == 2) version = 1; else if (vb == 3) version = 0; else return 0;
    if (((h[1] >> 1) & 3) != 1) return 0;                /* Layer III only */
    const unsigned bi = (h[2] >> 4) & 15;
    if (bi == 0 || bi == 15) return 0;
    const unsigned kbps = version == 0 ? br1[bi] : br2[bi];
    const unsigned si = (h[2] >> 2) & 3;
    if (si == 3) return 0;
    const unsigned rate = sr[version][si];
    const size_t spf = version == 0 ? 1152 : 576;
    out->version = version;
    out->has_crc = (h[1] & 1) == 0;
    out->channel_mode = (h[3] >> 6) & 3;
    out->frame_size = (spf * kbps * 125) / rate + ((h[2] & 2) ? 1 : 0);
    return 1;
}

/* skip_id3v2, src/lib.rs:343-354 */
static size_t skip_id3v2(const uint8_t* d, size_t len) {
    if (len < 10 || memcmp(d, "ID3", 3) != 0) return 0;
    const size_t size = ((size_t)(d[6] & 0x7F) << 21) | ((size_t)(d[7] & 0x7F) << 14) | ((size_t)(d[8] & 0x7F) << 7) | (size_t)(d[9] & 0x7F);
    return 10 + size;
}

static uint32_t u32le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* find_audio_end, src/lib.rs:358-383 (APE_FLAG_HEADER_PRESENT = 1 << 31, :844) */
static size_t find_audio_end(const uint8_t* d, size_t len) {
    size_t end = len;
    if (end >= 128 && memcmp(d + end - 128, "TAG", 3) == 0) end -= 128;
    if (end >= 32 && memcmp(d + end - 32, "APETAGEX", 8) == 0) {
        const size_t footer = end - 32;
        const size_t tag_size = u32le(d + footer + 12);
        const uint32_t flags = u32le(d + footer + 20);
        const size_t header = (flags & 0x80000000u) ? 32 : 0;
        if (footer + 32 >= tag_size + header) end = footer + 32 - tag_size - header;
    }
    return end;
}

/* is_xing_frame, src/lib.rs:388-408 */
static int is_xing_frame(const uint8_t* d, size_t len, size_t pos, const hdr_t* h) {
    const size_t side = h->version == 0 ? (h->channel_mode == 3 ? 17 : 32) : (h->channel_mode == 3 ? 9 : 17);
    const size_t x = pos + (h->has_crc ? 6 : 4) + side;
    if (x + 4 > len) return 0;
    return memcmp(d + x, "Xing", 4) == 0 || memcmp(d + x, "Info", 4) == 0;
}

typedef struct { size_t byte; unsigned bit; } loc_t;

/* calculate_gain_locations, src/lib.rs:262-296 */
static int gain_locations(size_t pos, const hdr_t* h, loc_t* loc) {
    const size_t side = pos + (h->has_crc ? 6 : 4);
    const int nch = h->channel_mode == 3 ? 1 : 2, ngr = h->version == 0 ? 2 : 1;
    const size_t before = h->version == 0 ? (nch == 1 ? 18 : 20) : (nch == 1 ? 9 : 10);
    const size_t per = h->version == 0 ? 59 : 63;
    int n = 0;
    for (int gr = 0; gr < ngr; gr++)
        for (int ch = 0; ch < nch; ch++) {
            const size_t bit = before + (size_t)(gr * nch + ch) * per + 21;
            loc[n].byte = side + bit / 8;
            loc[n].bit = (unsigned)(bit % 8);
            n++;
        }
    return n;
}

/* read_gain_at, src/lib.rs:299-315 */
static uint8_t read_gain_at(const uint8_t* d, size_t len, const loc_t* l) {
    const size_t i = l->byte;
    if (i >= len) return 0;
    if (l->bit == 0) return d[i];
    if (i + 1 < len) return (uint8_t)((uint8_t)(d[i] << l->bit) | (uint8_t)(d[i + 1] >> (8 - l->bit)));
    return (uint8_t)(d[i] << l->bit);
}

/* write_gain_at, src/lib.rs:318-340 */
static void write_gain_at(uint8_t* d, size_t len, const loc_t* l, uint8_t v) {
    const size_t i = l->byte;
    if (i >= len) return;
    if (l->bit == 0) { d[i] = v; return; }
    const unsigned s = l->bit;
    const uint8_t mask_high = (uint8_t)(0xFFu << (8 - s)), mask_low = (uint8_t)(0xFFu >> s);
    d[i] = (uint8_t)((d[i] & mask_high) | (v >> s));
    if (i + 1 < len) d[i + 1] = (uint8_t)((d[i + 1] & mask_low) | (uint8_t)(v << (8 - s)));
}

/* adjust_gain_value, src/lib.rs:526-541 */
static uint8_t adjust_gain_value(uint8_t cur, int32_t steps, int wrap) {
    if (!wrap) {
        if (steps > 0) { const int a = steps < 255 ? steps : 255; const int v = cur + a; return (uint8_t)(v > 255 ? 255 : v); }
        const int64_t neg = -(int64_t)steps;
        const int a = neg < 255 ? (int)neg : 255;
        const int v = cur - a;
        return (uint8_t)(v < 0 ? 0 : v);
    }
    const int32_t ng = (int32_t)(((int64_t)cur + steps) % 256);   /* Rust % truncates toward zero, like C */
    return (uint8_t)((ng + 256) % 256);
}

/* The walk shared by iterate_frames (src/lib.rs:413-466), apply_gain_to_data (:544-596) and apply_gain_to_channel_data
 * (:668-735).  `image` may be patched in place when steps != 0. */
static size_t walk(uint8_t* d, size_t len, gfo_stats* st, int patch, int32_t steps, int wrap, int channel) {
    const size_t audio_end = find_audio_end(d, len);
    size_t pos = skip_id3v2(d, len);
    size_t frames = 0;
    while (pos + 4 <= audio_end) {
        hdr_t h;
        if (!parse_header(d + pos, len - pos, &h)) { pos++; continue; }
        const size_t next = pos + h.frame_size;
        int valid;
        if (next + 2 <= audio_end) valid = d[next] == 0xFF && (d[next + 1] & 0xE0) == 0xE0;
        else valid = next <= audio_end;
        if (!valid) { pos++; continue; }
        if (is_xing_frame(d, len, pos, &h)) { pos = next; continue; }
        loc_t loc[4];
        const int n = gain_locations(pos, &h, loc);
        if (st) {
            if (frames == 0) { st->mpeg_version = h.version; st->channel_mode = h.channel_mode; }
            for (int i = 0; i < n; i++) {
                const uint8_t g = read_gain_at(d, len, &loc[i]);
                if (g < st->min_gain) st->min_gain = g;
                if (g > st->max_gain) st->max_gain = g;
                st->gain_sum += g;
                st->gain_count++;
            }
        }
        if (patch) {
            const int nch = h.channel_mode == 3 ? 1 : 2, ngr = h.version == 0 ? 2 : 1;
            if (channel < 0) {
                for (int i = 0; i < n; i++) write_gain_at(d, len, &loc[i], adjust_gain_value(read_gain_at(d, len, &loc[i]), steps, wrap));
            } else {
                for (int gr = 0; gr < ngr; gr++) {
                    const int li = gr * nch + channel;
                    if (li < n) write_gain_at(d, len, &loc[li], adjust_gain_value(read_gain_at(d, len, &loc[li]), steps, 0));
                }
            }
        }
        frames++;
        pos = next;
    }
    return frames;
}

/* analyze, src/lib.rs:470-514.  Returns 0, or -1 for "No valid MP3 frames found". */
int gfo_analyze(const uint8_t* image, size_t len, gfo_stats* out) {
    memset(out, 0, sizeof *out);
    out->min_gain = 255;
    out->max_gain = 0;
    out->frame_count = walk((uint8_t*)image, len, out, 0, 0, 0, -1);
    if (out->frame_count == 0) return -1;
    out->avg_gain = (double)out->gain_sum / (double)out->gain_count;
    out->headroom_steps = 255 - (int32_t)out->max_gain;
    out->headroom_db = (double)out->headroom_steps * 1.5;   /* GAIN_STEP_DB */
    return 0;
}

/* apply_gain_to_data / apply_gain_to_channel_data on a caller-owned copy of the file image.  channel < 0: all channels
 * (wrap selects GainMode::Wrapping); channel 0 / 1: that channel only, saturating.  Returns the number of modified frames. */
size_t gfo_apply_gain(uint8_t* image, size_t len, int32_t steps, int wrap, int channel) {
    return walk(image, len, NULL, 1, steps, wrap, channel);
}

/* ---- hooks for the reference's own unit tests (src/lib.rs:1344-1442) ---------------------------------------------- */
int gfo_parse_header(const uint8_t* h4, int* version, int* channel_mode, size_t* frame_size) {
    hdr_t h;
    if (!parse_header(h4, 4, &h)) return 0;
    *version = h.version; *channel_mode = h.channel_mode; *frame_size = h.frame_size;
    return 1;
}
uint8_t gfo_read_gain_at(const uint8_t* d, size_t len, size_t byte_offset, unsigned bit_offset) {
    const loc_t l = {byte_offset, bit_offset};
    return read_gain_at(d, len, &l);
}
void gfo_write_gain_at(uint8_t* d, size_t len, size_t byte_offset, unsigned bit_offset, uint8_t value) {
    const loc_t l = {byte_offset, bit_offset};
    write_gain_at(d, len, &l, value);
}
size_t gfo_skip_id3v2(const uint8_t* d, size_t len) { return skip_id3v2(d, len); }
size_t gfo_find_audio_end(const uint8_t* d, size_t len) { return find_audio_end(d, len); }
int gfo_is_xing_frame(const uint8_t* d, size_t len, size_t pos) {
    hdr_t h;
    if (len < pos + 4 || !parse_header(d + pos, len - pos, &h)) return 0;
    return is_xing_frame(d, len, pos, &h);
}
uint8_t gfo_adjust_gain_value(uint8_t cur, int32_t steps, int wrap) { return adjust_gain_value(cur, steps, wrap); }
