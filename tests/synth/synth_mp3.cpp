// synth_mp3.cpp -- deterministic synthetic MPEG-1 Layer III bitstream generator (TEST / BENCH INFRASTRUCTURE).
//
// There is no MP3 encoder in the build image, and BASELINE.json's configs are "synthetic ... MP3 corpora", so this file
// synthesises valid Layer III streams directly in the quantised domain: seeded random quantised spectra with a decaying
// envelope are Huffman-coded with the standard code books, scalefactors / side info / bit reservoir are packed like an
// encoder would, and the result is a legal CBR or VBR stream (checked against FFmpeg's mp3float and the oracle decoder
// in tests/test_oracle_mp3.py).  Streams are built from self-contained GROUPS of frames (first frame has
// main_data_begin = 0) so that any concatenation of groups is again a valid stream.
//
// Not part of the product; nothing under mp3rgain_b200/ links it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../oracle/l3_tables_oracle.h"

namespace {

struct Rng {  // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(next() % (uint64_t)n); }
    bool chance(double p) { return uni() < p; }
};

struct BitWriter {
    std::vector<uint8_t> buf;
    size_t nbits = 0;
    void put(uint32_t v, int n) {
        for (int i = n - 1; i >= 0; i--) {
            if ((nbits & 7) == 0) buf.push_back(0);
            if ((v >> i) & 1) buf.back() |= (uint8_t)(0x80 >> (nbits & 7));
            nbits++;
        }
    }
};

struct HuffEnc {
    uint32_t code[16][16];
    uint8_t len[16][16];
    int side;
};
HuffEnc g_enc[16];
int g_long_start[9][23], g_short_start[9][14];
bool g_init = false;
double g_is_prob = 0.25;  // share of joint-stereo frames that use intensity stereo (tests: high, for coverage; bench: 2 %, encoders at 128 kbps rarely use it)
int g_crc = 0;            // 1: frames carry the 16-bit CRC word after the header (protection_bit = 0); its value is not computed
int g_version = 3;        // MPEG version field of the streams: 3 = MPEG-1, 2 = MPEG-2 (LSF), 0 = MPEG-2.5 (LSF)
unsigned g_disable = 0;  // debug: bit0 short blocks, bit1 mixed, bit2 scfsi, bit3 preflag, bit4 sf_scale, bit5 subblock_gain, bit6 escapes, bit7 random tables, bit8 MS, bit9 intensity stereo

void init() {
    if (g_init) return;
    for (int s = 1; s <= L3_NUM_HUFF; s++) {
        int n = L3_HUFF_SIZE[s - 1], off = L3_HUFF_OFF[s - 1];
        uint32_t c = 0;
        g_enc[s].side = (int)std::lround(std::sqrt((double)n));
        for (int i = 0; i < n; i++) {
            int l = L3_HUFF_LEN[off + i], sym = L3_HUFF_SYM[off + i];
            g_enc[s].code[sym >> 4][sym & 15] = c >> (32 - l);
            g_enc[s].len[sym >> 4][sym & 15] = (uint8_t)l;
            c += 1u << (32 - l);
        }
    }
    for (int r = 0; r < 9; r++) {
        int a = 0;
        for (int i = 0; i < 22; i++) { g_long_start[r][i] = a; a += L3_SFB_LONG_W[r][i]; }
        g_long_start[r][22] = a;
        a = 0;
        for (int i = 0; i < 13; i++) { g_short_start[r][i] = a; a += L3_SFB_SHORT_W[r][i]; }
        g_short_start[r][13] = a;
    }
    g_init = true;
}

const int BR_V1_L3[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
const int SR_V1[3] = {44100, 48000, 32000};
const int BR_V2_L3[15] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160};
const int SR_ROW_HZ[9] = {44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000};

struct Granule {
    int block_type = 0, mixed = 0, window_switching = 0;
    int global_gain = 150, scalefac_compress = 0, preflag = 0, scalefac_scale = 0, count1table = 0;
    int table_select[3] = {0, 0, 0}, subblock_gain[3] = {0, 0, 0}, region0_count = 0, region1_count = 0;
    int big_values = 0, count1_end = 0;  // lines [0,2*big_values) big, [2*bv, count1_end) quads
    int sf_l[22] = {0}, sf_s[13][3] = {{0}};
    int16_t is[576] = {0};
    int part2_bits = 0, part3_bits = 0;
    int lsf_slen[4] = {0, 0, 0, 0}, lsf_tbl = 0;   // MPEG-2: field widths and the L3_LSF_NSF table chosen by scalefac_compress
};

int max_of_table(int sel) {  // largest codable value with table_select `sel`
    int slot = L3_TABSEL[sel][0], lb = L3_TABSEL[sel][1];
    if (!slot) return 0;
    int side = g_enc[slot].side;
    return (side == 16 && lb) ? 15 + (1 << lb) - 1 : side - 1;
}

int pair_bits(int sel, int x, int y) {
    int slot = L3_TABSEL[sel][0], lb = L3_TABSEL[sel][1];
    if (!slot) return 0;
    int ax = std::abs(x), ay = std::abs(y);
    int cx = std::min(ax, 15), cy = std::min(ay, 15);
    int bits = g_enc[slot].len[cx][cy];
    if (cx == 15 && lb) bits += lb;
    if (ax) bits++;
    if (cy == 15 && lb) bits += lb;
    if (ay) bits++;
    return bits;
}

void put_pair(BitWriter& w, int sel, int x, int y) {
    int slot = L3_TABSEL[sel][0], lb = L3_TABSEL[sel][1];
    if (!slot) return;
    int ax = std::abs(x), ay = std::abs(y);
    int cx = std::min(ax, 15), cy = std::min(ay, 15);
    w.put(g_enc[slot].code[cx][cy], g_enc[slot].len[cx][cy]);
    if (cx == 15 && lb) w.put((uint32_t)(ax - 15), lb);
    if (ax) w.put(x < 0, 1);
    if (cy == 15 && lb) w.put((uint32_t)(ay - 15), lb);
    if (ay) w.put(y < 0, 1);
}

void region_bounds(const Granule& g, int sr_row, int& r0, int& r1) {
    if (g.window_switching) { r0 = g.block_type == 2 ? 3 * g_short_start[sr_row][3] : g_long_start[sr_row][8]; r1 = 576; return; }
    r0 = g_long_start[sr_row][g.region0_count + 1];
    int k = g.region0_count + g.region1_count + 2;
    r1 = g_long_start[sr_row][k > 22 ? 22 : k];
}

// choose tables for the three regions and count the part-3 bits
void choose_tables_and_count(Granule& g, int sr_row, Rng& rng) {
    int r0, r1;
    region_bounds(g, sr_row, r0, r1);
    int bv2 = g.big_values * 2;
    int lim[4] = {0, std::min(r0, bv2), std::min(r1, bv2), bv2};
    int bits = 0;
    int nreg = g.window_switching ? 2 : 3;
    if (g.window_switching) lim[2] = bv2;
    for (int r = 0; r < nreg; r++) {
        int a = lim[r], b = lim[r + 1];
        int mx = 0;
        for (int i = a; i < b; i++) mx = std::max(mx, std::abs((int)g.is[i]));
        // candidates: every selector whose range covers mx (selectors 4 and 14 do not exist)
        int best = -1, best_bits = 1 << 30;
        std::vector<int> cands;
        for (int sel = 0; sel < 32; sel++) {
            if (sel == 4 || sel == 14) continue;
            if (sel == 0 && mx > 0) continue;
            if (max_of_table(sel) < mx) continue;
            if (a == b && sel != 0 && !rng.chance(0.05)) continue;  // empty region: usually table 0, sometimes junk selector
            int bb = 0;
            for (int i = a; i < b; i += 2) bb += pair_bits(sel, g.is[i], g.is[i + 1]);
            cands.push_back(sel);
            if (bb < best_bits) { best_bits = bb; best = sel; }
        }
        int sel = best;
        if (cands.size() > 1 && rng.chance(0.25) && !(g_disable & 128)) {  // sometimes a non-optimal but valid table, to exercise every code book
            int pick = cands[rng.below((int)cands.size())];
            if (max_of_table(pick) <= std::max(4 * mx + 3, 15) || rng.chance(0.2)) sel = pick;
        }
        g.table_select[r] = sel;
        for (int i = a; i < b; i += 2) bits += pair_bits(sel, g.is[i], g.is[i + 1]);
    }
    // count1: pick the cheaper quad table most of the time
    int cb[2] = {0, 0};
    for (int t = 0; t < 2; t++)
        for (int i = bv2; i < g.count1_end; i += 4) {
            int v = (g.is[i] != 0) << 3 | (g.is[i + 1] != 0) << 2 | (g.is[i + 2] != 0) << 1 | (g.is[i + 3] != 0);
            cb[t] += L3_C1_LENS[t][v] + __builtin_popcount(v);
        }
    g.count1table = (cb[1] < cb[0]) ? 1 : 0;
    if (rng.chance(0.1)) g.count1table ^= 1;
    bits += cb[g.count1table];
    g.part3_bits = bits;
}

// MPEG-2 scalefactors in bitstream order (long: sfb 0..20; short: sfb 0..11 x 3; mixed: 6 long bands, short sfb 3..11 x 3)
int lsf_values(const Granule& g, int* v) {
    int n = 0;
    if (g.block_type != 2) { for (int sfb = 0; sfb < 21; sfb++) v[n++] = g.sf_l[sfb]; }
    else {
        int s0 = 0;
        if (g.mixed) { for (int sfb = 0; sfb < 6; sfb++) v[n++] = g.sf_l[sfb]; s0 = 3; }
        for (int sfb = s0; sfb < 12; sfb++) for (int k = 0; k < 3; k++) v[n++] = g.sf_s[sfb][k];
    }
    return n;
}

int part2_bits_of(const Granule& g, const int* scfsi, int gr) {
    if (g_version != 3) {
        int cls = g.block_type != 2 ? 0 : (g.mixed ? 2 : 1), bits = 0;
        for (int i = 0; i < 4; i++) bits += L3_LSF_NSF[g.lsf_tbl][cls][i] * g.lsf_slen[i];
        return bits;
    }
    int s1 = L3_SLEN[0][g.scalefac_compress], s2 = L3_SLEN[1][g.scalefac_compress];
    if (g.block_type == 2) {
        if (g.mixed) return 8 * s1 + 9 * s1 + 18 * s2;
        return 18 * s1 + 18 * s2;
    }
    int bits = 0;
    static const int cnt[4] = {6, 5, 5, 5};
    for (int grp = 0; grp < 4; grp++)
        if (!(gr == 1 && scfsi[grp])) bits += cnt[grp] * (grp < 2 ? s1 : s2);
    return bits;
}

void write_granule_data(BitWriter& w, const Granule& g, const int* scfsi, int gr, int sr_row) {
    int s1 = L3_SLEN[0][g.scalefac_compress & 15], s2 = L3_SLEN[1][g.scalefac_compress & 15];
    if (g_version != 3) {
        int v[40], cls = g.block_type != 2 ? 0 : (g.mixed ? 2 : 1), k = 0;
        lsf_values(g, v);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < L3_LSF_NSF[g.lsf_tbl][cls][i]; j++) w.put((uint32_t)v[k++], g.lsf_slen[i]);
    } else if (g.block_type == 2) {
        int s0 = 0;
        if (g.mixed) { for (int sfb = 0; sfb < 8; sfb++) w.put(g.sf_l[sfb], s1); s0 = 3; }
        for (int sfb = s0; sfb < 6; sfb++) for (int k = 0; k < 3; k++) w.put(g.sf_s[sfb][k], s1);
        for (int sfb = 6; sfb < 12; sfb++) for (int k = 0; k < 3; k++) w.put(g.sf_s[sfb][k], s2);
    } else {
        static const int lo[4] = {0, 6, 11, 16}, hi[4] = {6, 11, 16, 21};
        for (int grp = 0; grp < 4; grp++) {
            if (gr == 1 && scfsi[grp]) continue;
            for (int sfb = lo[grp]; sfb < hi[grp]; sfb++) w.put(g.sf_l[sfb], grp < 2 ? s1 : s2);
        }
    }
    int r0, r1;
    region_bounds(g, sr_row, r0, r1);
    int bv2 = g.big_values * 2;
    for (int i = 0; i < bv2; i += 2) {
        int region = i < r0 ? 0 : (i < r1 ? 1 : 2);
        put_pair(w, g.table_select[region], g.is[i], g.is[i + 1]);
    }
    for (int i = bv2; i < g.count1_end; i += 4) {
        int v = (g.is[i] != 0) << 3 | (g.is[i + 1] != 0) << 2 | (g.is[i + 2] != 0) << 1 | (g.is[i + 3] != 0);
        w.put(L3_C1_BITS[g.count1table][v], L3_C1_LENS[g.count1table][v]);
        for (int k = 0; k < 4; k++) if (g.is[i + k]) w.put(g.is[i + k] < 0, 1);
    }
}

// Limit the spectrum to `cutoff` lines, derive big_values / count1 partition.
void partition(Granule& g, int cutoff) {
    for (int i = cutoff; i < 576; i++) g.is[i] = 0;
    int last = 576;
    while (last > 0 && g.is[last - 1] == 0) last--;
    int lastbig = last;
    while (lastbig > 0 && std::abs((int)g.is[lastbig - 1]) <= 1) lastbig--;
    int bv2 = (lastbig + 1) & ~1;
    if (bv2 > 576) bv2 = 576;
    int c1 = bv2 + ((last - bv2 + 3) / 4) * 4;
    if (last <= bv2) c1 = bv2;
    if (c1 > 576) {  // quads cannot cross 576: fold the tail into big_values instead
        bv2 = (last + 1) & ~1;
        c1 = bv2;
    }
    g.big_values = bv2 / 2;
    g.count1_end = c1;
}

struct StreamCfg {
    int sr_idx, br_idx /* 0 = VBR */, channels;
    double level_db;  // target RMS of the decoded PCM in dBFS
};

// energy of the requantised spectrum in units where global_gain = 210
double xr_energy_unit(const Granule& g, int sr_row) {
    double e = 0;
    int mult = g.scalefac_scale ? 4 : 2;
    int i = 0;
    if (g.block_type != 2 || g.mixed) {
        int nl = 22;
        if (g.block_type == 2) for (nl = 0; g_long_start[sr_row][nl] < 36; nl++) {}
        for (int sfb = 0; sfb < nl; sfb++) {
            double sc = std::exp2(-0.25 * mult * (g.sf_l[sfb] + (g.preflag ? L3_PRETAB[sfb] : 0)));
            for (int k = 0; k < L3_SFB_LONG_W[sr_row][sfb]; k++, i++) {
                double v = std::pow(std::abs((double)g.is[i]), 4.0 / 3.0) * sc;
                e += v * v;
            }
        }
    }
    if (g.block_type == 2)
        for (int sfb = g.mixed ? 3 : 0; sfb < 13; sfb++)
            for (int w = 0; w < 3; w++) {
                double sc = std::exp2(-0.25 * (8 * g.subblock_gain[w] + mult * g.sf_s[sfb][w]));
                for (int k = 0; k < L3_SFB_SHORT_W[sr_row][sfb]; k++, i++) {
                    double v = std::pow(std::abs((double)g.is[i]), 4.0 / 3.0) * sc;
                    e += v * v;
                }
            }
    return e;
}

void make_granule(Granule& g, Rng& rng, int sr_row, int bt, int mixed, double amp, double tilt, int target_bits,
                  const int* scfsi, int gr, double level_db, double is_cut = 1e9) {
    g = Granule();
    g.block_type = bt;
    g.window_switching = bt != 0;
    g.mixed = (bt == 2) ? mixed : 0;
    g.scalefac_scale = rng.chance(0.2) && !(g_disable & 16);
    g.preflag = (bt != 2) && rng.chance(0.15) && !(g_disable & 8);
    g.scalefac_compress = rng.below(16);
    int s1 = L3_SLEN[0][g.scalefac_compress], s2 = L3_SLEN[1][g.scalefac_compress];
    auto sfv = [&](int bits) { if (!bits) return 0; int m = (1 << bits) - 1; int v = (int)(-std::log(1.0 - rng.uni()) * 1.2); return std::min(v, m); };
    const bool lsf_is_right = g_version != 3 && is_cut < 1e8;   // right channel of an MPEG-2 intensity-stereo frame
    if (g_version != 3) {
        // MPEG-2: pick one of the three scalefac_compress ranges, then field widths, then values band by band
        int pick = rng.below(10);
        int* sl = g.lsf_slen;
        if (lsf_is_right) {
            // intensity partition, first range only (every field at least 1 bit wide, at most 4: the other ranges leave
            // bands without a position, which decoders treat differently); bit 0 of scalefac_compress = intensity_scale
            sl[0] = 1 + rng.below(4); sl[1] = 1 + rng.below(4); sl[2] = 1 + rng.below(4); sl[3] = 0;
            g.scalefac_compress = ((sl[0] * 36 + sl[1] * 6 + sl[2]) << 1) | rng.below(2); g.lsf_tbl = 3; g.preflag = 0;
        } else if (pick < 7) { sl[0] = rng.below(5); sl[1] = rng.below(5); sl[2] = rng.below(4); sl[3] = rng.below(4);
                        g.scalefac_compress = ((sl[0] * 5 + sl[1]) << 4) | (sl[2] << 2) | sl[3]; g.lsf_tbl = 0; g.preflag = 0; }
        else if (pick < 9) { sl[0] = rng.below(5); sl[1] = rng.below(5); sl[2] = rng.below(4); sl[3] = 0;
                             g.scalefac_compress = 400 + (((sl[0] * 5 + sl[1]) << 2) | sl[2]); g.lsf_tbl = 1; g.preflag = 0; }
        else { sl[0] = rng.below(4); sl[1] = rng.below(3); sl[2] = sl[3] = 0;
               g.scalefac_compress = 500 + sl[0] * 3 + sl[1]; g.lsf_tbl = 2; g.preflag = 1; }
        int cls = bt != 2 ? 0 : (g.mixed ? 2 : 1), v[40], k = 0;
        for (int i = 0; i < 4; i++) for (int j = 0; j < L3_LSF_NSF[g.lsf_tbl][cls][i]; j++) {
            v[k] = sfv(sl[i]);
            if (lsf_is_right && sl[i] && v[k] == (1 << sl[i]) - 1) v[k]--;   // never the illegal position (decoders disagree on it)
            k++;
        }
        k = 0;
        if (cls == 0) { for (int sfb = 0; sfb < 21; sfb++) g.sf_l[sfb] = v[k++]; }
        else {
            int s0 = 0;
            if (cls == 2) { for (int sfb = 0; sfb < 6; sfb++) g.sf_l[sfb] = v[k++]; s0 = 3; }
            for (int sfb = s0; sfb < 12; sfb++) for (int w3 = 0; w3 < 3; w3++) g.sf_s[sfb][w3] = v[k++];
        }
        for (int w3 = 0; w3 < 3; w3++) g.subblock_gain[w3] = (bt == 2 && rng.chance(0.3) && !(g_disable & 32)) ? rng.below(4) : 0;
    } else if (bt == 2) {
        if (g.mixed) for (int sfb = 0; sfb < 8; sfb++) g.sf_l[sfb] = sfv(s1);
        for (int sfb = g.mixed ? 3 : 0; sfb < 12; sfb++) for (int k = 0; k < 3; k++) g.sf_s[sfb][k] = sfv(sfb < 6 ? s1 : s2);
        for (int k = 0; k < 3; k++) g.subblock_gain[k] = (rng.chance(0.3) && !(g_disable & 32)) ? rng.below(4) : 0;
    } else {
        for (int sfb = 0; sfb < 21; sfb++) g.sf_l[sfb] = sfv(sfb < 11 ? s1 : s2);
    }
    if (bt == 0) { g.region0_count = rng.below(16); g.region1_count = rng.below(8); }
    // spectrum: Laplacian magnitudes under a decaying envelope; beyond `knee` only 0/+-1 survive (count1 material) and
    // nothing above ~16 kHz (an encoder's low-pass at this bitrate).  The overall scale of the magnitudes -- the encoder's
    // quantiser step -- is then chosen so that the coded granule FILLS its share of the frame's bits, like a real CBR
    // stream: most of the 576 lines carry data, big_values covers a few hundred lines, the rest is count1 material.
    int knee = 60 + rng.below(260);
    const int lowpass = (int)std::min(576.0, 576.0 * 16000.0 / (0.5 * SR_ROW_HZ[sr_row]));
    double base[576];
    bool neg[576], capped[576];
    int escape[576];
    for (int i = 0; i < 576; i++) {
        double f = (double)i;  // rough frequency of the line
        if (bt == 2 && !(g.mixed && i < 36)) {
            // bitstream order is [sfb][win][k]; approximate the line's frequency by its sfb start
            int p = i, sfb = 0;
            while (sfb < 12 && p >= 3 * g_short_start[sr_row][sfb + 1]) sfb++;
            f = 3.0 * g_short_start[sr_row][sfb];
        }
        base[i] = (f >= lowpass || f >= is_cut) ? 0.0 : -std::log(1.0 - rng.uni()) * std::pow(1.0 + f / 6.0, -tilt);
        capped[i] = f > knee;
        escape[i] = (rng.chance(0.0008) && f < 40 && f < is_cut && !(g_disable & 64)) ? 16 + rng.below(600) : 0;  // occasional linbits escapes
        neg[i] = rng.chance(0.5);
    }
    auto build = [&](double scale) {
        for (int i = 0; i < 576; i++) {
            int v = (int)std::min(8206.0, base[i] * scale);
            if (capped[i]) v = std::min(v, 1);
            if (escape[i]) v = escape[i];
            g.is[i] = (int16_t)(neg[i] ? -v : v);
        }
    };
    g.part2_bits = part2_bits_of(g, scfsi, gr);
    {
        const int want = (int)(0.97 * std::min(target_bits, 4095));
        double lo_s = 0.05, hi_s = 400.0 * std::max(1.0, amp / 20.0);
        const Rng trial0 = rng;
        for (int it = 0; it < 14; it++) {
            const double mid = std::sqrt(lo_s * hi_s);
            build(mid);
            partition(g, 576);
            Rng r2 = trial0;
            choose_tables_and_count(g, sr_row, r2);
            if (g.part2_bits + g.part3_bits <= want) lo_s = mid; else hi_s = mid;
        }
        build(lo_s);
    }
    int16_t full[576];
    memcpy(full, g.is, sizeof full);
    // largest cutoff whose coded size fits the target
    int lo = 0, hi = 576;
    g.part2_bits = part2_bits_of(g, scfsi, gr);
    if (g.part2_bits > std::min(target_bits, 4095)) {  // no room even for scalefactors: slen 0/0
        g.scalefac_compress = 0;
        g.lsf_slen[0] = g.lsf_slen[1] = g.lsf_slen[2] = g.lsf_slen[3] = 0; g.lsf_tbl = 0;
        if (g_version != 3) g.preflag = 0;
        memset(g.sf_l, 0, sizeof g.sf_l);
        memset(g.sf_s, 0, sizeof g.sf_s);
        g.part2_bits = 0;
    }
    Rng trial = rng;
    while (lo < hi) {
        int mid = (lo + hi + 1) / 2;
        memcpy(g.is, full, sizeof full);
        partition(g, mid);
        Rng r2 = trial;
        choose_tables_and_count(g, sr_row, r2);
        if (g.part2_bits + g.part3_bits <= std::min(target_bits, 4095)) lo = mid; else hi = mid - 1;
    }
    memcpy(g.is, full, sizeof full);
    partition(g, lo);
    {
        Rng r2 = trial;  // same state as the accepted trial => same table choices => same size
        choose_tables_and_count(g, sr_row, r2);
        rng = r2;
    }
    // level: choose global_gain so that the granule lands near level_db (PCM energy ~ K * spectral energy)
    double e = xr_energy_unit(g, sr_row);
    int gg = 140;
    if (e > 0) {
        const double K_DB = 26.3;  // measured: 10log10(mean PCM^2) - 10log10(sum xr^2 / 576), see tests/test_oracle_mp3.py
        double cur_db = 10.0 * std::log10(e / 576.0) + K_DB;
        gg = 210 + (int)std::lround((level_db - cur_db) / 1.50515);  // one global_gain step = 20log10(2^(1/4)) dB
    }
    g.global_gain = std::max(0, std::min(255, gg));
}

void write_side(BitWriter& w, int nch, int main_data_begin, const int scfsi[2][4], const Granule gr[2][2]) {
    const bool lsf = g_version != 3;
    if (lsf) {
        w.put((uint32_t)main_data_begin, 8);
        w.put(0, nch == 1 ? 1 : 2);
    } else {
        w.put((uint32_t)main_data_begin, 9);
        w.put(0, nch == 1 ? 5 : 3);
        for (int ch = 0; ch < nch; ch++) for (int i = 0; i < 4; i++) w.put(scfsi[ch][i], 1);
    }
    for (int g = 0; g < (lsf ? 1 : 2); g++)
        for (int ch = 0; ch < nch; ch++) {
            const Granule& q = gr[g][ch];
            w.put((uint32_t)(q.part2_bits + q.part3_bits), 12);
            w.put((uint32_t)q.big_values, 9);
            w.put((uint32_t)q.global_gain, 8);
            w.put((uint32_t)q.scalefac_compress, lsf ? 9 : 4);
            w.put((uint32_t)q.window_switching, 1);
            if (q.window_switching) {
                w.put((uint32_t)q.block_type, 2);
                w.put((uint32_t)q.mixed, 1);
                w.put((uint32_t)q.table_select[0], 5);
                w.put((uint32_t)q.table_select[1], 5);
                for (int i = 0; i < 3; i++) w.put((uint32_t)q.subblock_gain[i], 3);
            } else {
                for (int i = 0; i < 3; i++) w.put((uint32_t)q.table_select[i], 5);
                w.put((uint32_t)q.region0_count, 4);
                w.put((uint32_t)q.region1_count, 3);
            }
            if (!lsf) w.put((uint32_t)q.preflag, 1);
            w.put((uint32_t)q.scalefac_scale, 1);
            w.put((uint32_t)q.count1table, 1);
        }
}

}  // namespace

extern "C" {

void synth_set_disable(unsigned mask) { g_disable = mask; }
void synth_set_intensity_prob(double p) { g_is_prob = p; }
void synth_set_version(int v) { g_version = v; }
void synth_set_crc(int on) { g_crc = on; }   // 3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5

// One self-contained group of n_frames frames.  br_idx: 1..14 CBR, 0 = VBR (uniform bitrate index per frame).
// level_db: target loudness of the PCM (dBFS RMS); the group wanders around it.  Returns bytes written (0 if cap too small).
size_t synth_group(uint64_t seed, int n_frames, int sr_idx, int br_idx, int channels, double level_db, uint8_t* out, size_t cap) {
    init();
    Rng rng(seed);
    const bool lsf = g_version != 3;
    const int sr_row = (g_version == 3 ? 0 : (g_version == 2 ? 3 : 6)) + sr_idx, sr = SR_ROW_HZ[sr_row];
    const int side_len = lsf ? (channels == 1 ? 9 : 17) : (channels == 1 ? 17 : 32);
    const int ngr = lsf ? 1 : 2;
    const size_t mdb_max = lsf ? 255 : 511;
    std::vector<uint8_t> frames;        // finished frame bytes (header + side + payload area)
    std::vector<size_t> frame_off, payload_off, payload_len;
    std::vector<uint8_t> md;            // main-data stream (concatenated payload areas) written by position
    size_t M = 0;                       // start of the current frame's payload area in md
    size_t E = 0;                       // end (bytes) of the previous frame's granule data
    double pad_acc = 0.0;
    int bt_state = 0;                   // block type sequence 0 -> 1 -> 2 (xN) -> 3 -> 0
    int run_mixed = 0;
    double amp = 3.0 + 25.0 * rng.uni(), tilt = 0.9 + 0.8 * rng.uni();
    double lvl = level_db + 6.0 * (rng.uni() - 0.5);
    for (int f = 0; f < n_frames; f++) {
        int bri = br_idx ? br_idx : 1 + rng.below(14);
        double exact = (lsf ? 72.0 * BR_V2_L3[bri] : 144.0 * BR_V1_L3[bri]) * 1000.0 / sr;
        int base = (int)exact;
        pad_acc += exact - base;
        int pad = 0;
        if (pad_acc >= 1.0) { pad = 1; pad_acc -= 1.0; }
        int flen = base + pad;
        const int hdr_len = 4 + (g_crc ? 2 : 0);
        int cap_f = flen - hdr_len - side_len;
        size_t P = std::max<size_t>(E, M > mdb_max ? M - mdb_max : 0);
        if (f == 0) P = M;
        int budget = (int)((M + cap_f - P) * 8);
        // slow random walks of spectral shape and level so that 50 ms windows spread over many histogram bins
        amp = std::min(60.0, std::max(1.5, amp * std::exp(0.25 * (rng.uni() - 0.5))));
        tilt = std::min(2.0, std::max(0.7, tilt + 0.1 * (rng.uni() - 0.5)));
        lvl = std::min(level_db + 9.0, std::max(level_db - 14.0, lvl + 1.6 * (rng.uni() - 0.5)));
        bool js = channels == 2 && rng.chance(0.6);
        bool ms = js && rng.chance(0.7) && !(g_disable & 256);
        // intensity stereo: the right channel is empty above a random bound and its scalefactors there are is_pos values
        bool ist = js && rng.chance(g_is_prob) && !(g_disable & 512);
        double is_cut = ist ? 20.0 + 300.0 * rng.uni() : 1e9;
        Granule gr[2][2];
        int scfsi[2][4] = {{0}};
        int bts[2], mixed_g[2];
        bts[1] = 0; mixed_g[1] = 0;
        for (int g = 0; g < ngr; g++) {   // the window sequence 0 -> 1 -> 2.. -> 3 -> 0 advances once per coded granule
            if (bt_state == 0) { bt_state = (rng.chance(0.03) && !(g_disable & 1)) ? 1 : 0; if (bt_state) run_mixed = rng.chance(0.25) && !(g_disable & 2) && sr_row != 8; }   // (8 kHz mixed blocks: decoders disagree on the long/short split, none generated)
            else if (bt_state == 1) bt_state = 2;
            else if (bt_state == 2) bt_state = rng.chance(0.5) ? 2 : 3;
            else bt_state = 0;
            bts[g] = bt_state;
            mixed_g[g] = run_mixed;
        }
        // mixed_block_flag is constant over a run of short blocks (an encoder keeps subbands 0-1 on long windows for the whole
        // run; flipping it mid-run is legal syntax but decoders disagree on the resulting overlap, FFmpeg included)
        int mixed[2] = {mixed_g[0], mixed_g[1]};
        int used = 0;
        for (int g = 0; g < ngr; g++)
            for (int ch = 0; ch < channels; ch++) {
                if (g == 1 && bts[0] != 2 && bts[1] != 2)
                    for (int i = 0; i < 4; i++) scfsi[ch][i] = rng.chance(0.3) && !(g_disable & 4);
                int left = (ngr - g) * channels - ch;
                int remain = budget - used;
                int target = (int)(remain / left * (0.7 + 0.6 * rng.uni()));
                if (left == 1) target = remain;
                target = std::max(0, std::min(target, remain));
                if (ms && ch == 1 && left > 1) target = (int)(target * 0.6);   // the side channel gets the smaller share
                double a = amp * ((ms && ch == 1) ? 0.35 : 1.0);
                make_granule(gr[g][ch], rng, sr_row, bts[g], mixed[g], a, tilt, target, scfsi[ch], g, lvl, (ist && ch == 1) ? is_cut : 1e9);
                used += gr[g][ch].part2_bits + gr[g][ch].part3_bits;
            }
        // header
        BitWriter hw;
        hw.put(0x7FF, 11); hw.put((uint32_t)g_version, 2);                          // sync, MPEG-1 (11) / MPEG-2 (10) / MPEG-2.5 (00)
        hw.put(1, 2); hw.put(g_crc ? 0u : 1u, 1);                                   // Layer III, protection bit (0 = CRC word follows)
        hw.put((uint32_t)bri, 4); hw.put((uint32_t)sr_idx, 2); hw.put((uint32_t)pad, 1); hw.put(0, 1);
        hw.put(channels == 1 ? 3u : (js ? 1u : 0u), 2);
        hw.put(js ? ((ms ? 2u : 0u) | (ist ? 1u : 0u)) : 0u, 2);
        hw.put(0, 1); hw.put(1, 1); hw.put(0, 2);
        if (g_crc) hw.put(0xA55A, 16);                                              // the CRC word (decoders do not verify it by default)
        write_side(hw, channels, (int)(M - P), scfsi, gr);
        // granule data into the main-data stream at P
        BitWriter dw;
        for (int g = 0; g < ngr; g++) for (int ch = 0; ch < channels; ch++) write_granule_data(dw, gr[g][ch], scfsi[ch], g, sr_row);
        size_t need = P + dw.buf.size();
        if (md.size() < M + cap_f) md.resize(M + cap_f, 0);
        if (need > md.size()) return 0;  // cannot happen: budget enforced
        memcpy(md.data() + P, dw.buf.data(), dw.buf.size());
        E = need;
        frame_off.push_back(frames.size());
        frames.insert(frames.end(), hw.buf.begin(), hw.buf.end());
        payload_off.push_back(M);
        payload_len.push_back((size_t)cap_f);
        frames.resize(frames.size() + cap_f, 0);
        M += cap_f;
    }
    // scatter the main-data stream back into the frames' payload areas
    for (size_t f = 0; f < frame_off.size(); f++)
        memcpy(frames.data() + frame_off[f] + 4 + (g_crc ? 2 : 0) + side_len, md.data() + payload_off[f], payload_len[f]);
    if (frames.size() > cap) return 0;
    memcpy(out, frames.data(), frames.size());
    return frames.size();
}

// A track = consecutive groups seeded from (seed, group index).  Returns bytes written.
size_t synth_track(uint64_t seed, int n_frames, int group_len, int sr_idx, int br_idx, int channels, double level_db,
                   uint8_t* out, size_t cap) {
    size_t n = 0;
    int g = 0;
    Rng tr(seed ^ 0xABCDEF);
    double lvl = level_db;
    for (int f = 0; f < n_frames; f += group_len, g++) {
        int k = std::min(group_len, n_frames - f);
        lvl = std::min(level_db + 4.0, std::max(level_db - 8.0, lvl + 3.0 * (tr.uni() - 0.5)));
        size_t w = synth_group(seed * 1000003ull + (uint64_t)g, k, sr_idx, br_idx, channels, lvl, out + n, cap - n);
        if (!w) return 0;
        n += w;
    }
    return n;
}

}  // extern "C"
