// rg_decode.cuh -- stage A on the device: MPEG-1 / 2 / 2.5 Layer III file images -> planar f32 PCM in HBM.
//
// The reference delegates this stage to symphonia-bundle-mp3 (src/replaygain.rs:807-822, 861-863, 882, 896); the kernels
// below restate ISO/IEC 11172-3 and 13818-3 with the arithmetic contract of DESIGN.md ("f32 decode contract"): every table
// is the frozen float32 data of l3_tables.h, products and sums round separately except where fmaf is written.  Compiled
// with -fmad=false.
//
//   K0  k_index_scan / k_index_merge / k_index_finalize   frame index: ID3v2 skip, sync / resync, Xing-Info skip, side-info
//                       validity, bit-reservoir bookkeeping -- 64 KiB segments walked speculatively in parallel, stitched per track
//       k_md_gather     every frame's main data into a contiguous per-track stream (the bit reservoir unrolled)
//   K1  k_huffman       one thread per granule-channel: scalefactors -> band gains, Huffman decode -> int16 spectra
//   K2  k_hybrid_synth  one CTA per run of granules of one track: requantise, MS / intensity stereo, reorder, alias
//                       reduction, IMDCT + overlap, frequency inversion, 32-point DCT + polyphase window -> PCM (or peak)
#pragma once
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>
#include "rg_internal.h"
#include "l3_tables.h"

// ---------------------------------------------------------------------------------------------------------------
// Device-resident tables (one copy per engine, built on the host from l3_tables.h)
#define HUF_L1 8
struct L3Dev {
    uint16_t huff[8192];        // two-level lookup, all 15 books in one array (see build_huff_lookup)
    uint16_t huff_base[16];     // first-level base of book slot s (1..15); [0] = an all-zero table
    uint32_t huff_total;        // entries used
    uint8_t tabsel[32][2];      // table_select -> {slot, linbits}
    uint8_t c1a_lut[64];        // count1 table A: 6-bit peek -> (len<<4)|quad
    uint16_t long_start[9][24];   // rows: 44100, 48000, 32000 | 22050, 24000, 16000 | 11025, 12000, 8000
    uint16_t short_start[9][16];
    uint8_t lsf_nsf[6][3][4];      // MPEG-2 scalefactor partition: [scalefac_compress range][long/short/mixed][field] -> count
    uint8_t slen[2][16];
    uint8_t pretab[24];
    // per sample-rate row line maps for requantisation / reorder
    uint8_t line_sfb_long[9][576];        // line -> long sfb
    uint8_t line_sfb_short[9][576];       // bitstream-order line (pure short) -> sfb*3+win
    uint8_t line_sfb_mixed[9][576];       // mixed: lines <36 -> long sfb (0..7), else 32 + sfb*3+win
    uint16_t scatter_short[9][576];       // bitstream-order line -> line after the short-block reorder (pure short)
    uint16_t scatter_mixed[9][576];       // same for mixed blocks (first 36 lines stay)
    float pow43[L3_POW43_N];
    float gain[L3_GAIN_MAX - L3_GAIN_MIN + 1];
    float synth_wa[8][32], synth_wb[8][32];
    uint8_t synth_ia[32], synth_ib[32];
    float is_ratio[2][8];       // intensity stereo: is_pos -> left / right factor
    float is_lsf[2][2][32];     // MPEG-2 intensity stereo: [intensity_scale][left / right][is_pos]
};


struct TrackIn { uint64_t off, len; };            // byte range of one file image inside the device corpus buffer
struct TrackMeta {                                 // K0 output
    int32_t status;
    uint32_t sr_idx, channels, n_frames, md_bytes, n_dropped, skipped_info, n_total, n_rewalks;
};
struct FrameRec {                                  // one per PCM-producing frame
    uint32_t file_off;   // frame start, relative to the track's first byte
    uint32_t md_start;   // byte offset of this frame's granule data in the track's main-data stream
    uint32_t hdr;        // the 4 header bytes, big-endian
    uint32_t md_end;     // bytes of main data available once this frame is read
};

// Run descriptor for K2: a contiguous range of granules of one track
struct RunDesc { uint32_t track; uint32_t g0; uint32_t ng; uint32_t pad; };
// Block descriptor for K1: 128 consecutive granule-channels of one track
struct HuffBlock { uint32_t track; uint32_t gc0; };

struct DecTrack {   // per-track addressing shared by K1/K2 (device array)
    uint64_t corpus_off;     // file image start in the corpus buffer
    uint64_t md_off;         // main-data stream start in the md buffer
    uint64_t frame_base;     // first FrameRec of this track
    uint64_t gc_base;        // first granule-channel slot (int16 spectra + nz) of this track within the batch
    uint64_t pcm_off;        // first PCM float of plane 0 (plane 1 follows at +pcm_plane)
    uint64_t pcm_plane;      // floats per plane
    uint32_t n_frames, channels, sr_idx /* row of the band tables, 0..8 */, md_bytes;
    uint32_t gpf;            // granules per frame: 2 (MPEG-1) or 1 (MPEG-2 / 2.5)
    uint32_t pad;
};

#include "k_ptx.cuh"
#include "k_frame_index.cuh"
#include "k_huffman.cuh"
#include "k_hybrid_synth.cuh"

// ---------------------------------------------------------------------------------------------------------------
struct DecodeScratch {
    L3Dev* d_tab = nullptr;
    void release() { if (d_tab) cudaFree(d_tab); d_tab = nullptr; }
};

static void build_huff_lookup(L3Dev& T) {
    // entry: leaf  = (len<<8)|sym            (bit15 = 0, len <= 19)
    //        link  = 0x8000 | nbits<<11 | (offset>>2)   -> second level at offset + next nbits bits
    memset(T.huff, 0, sizeof T.huff);
    uint32_t total = 0;
    for (int s = 1; s <= L3_NUM_HUFF; s++) { T.huff_base[s] = (uint16_t)total; total += 1u << HUF_L1; }
    T.huff_base[0] = 0;
    for (int s = 1; s <= L3_NUM_HUFF; s++) {
        int n = L3_HUFF_SIZE[s - 1], off = L3_HUFF_OFF[s - 1];
        std::vector<uint32_t> code(n);
        uint32_t c = 0;
        for (int i = 0; i < n; i++) { code[i] = c; c += 1u << (32 - L3_HUFF_LEN[off + i]); }
        int maxsub[1 << HUF_L1];
        memset(maxsub, 0, sizeof maxsub);
        for (int i = 0; i < n; i++) {
            int l = L3_HUFF_LEN[off + i];
            if (l > HUF_L1) { int p = code[i] >> (32 - HUF_L1); if (l - HUF_L1 > maxsub[p]) maxsub[p] = l - HUF_L1; }
        }
        uint32_t suboff[1 << HUF_L1];
        for (int p = 0; p < (1 << HUF_L1); p++)
            if (maxsub[p]) {
                total = (total + 3) & ~3u;
                suboff[p] = total;
                T.huff[T.huff_base[s] + p] = (uint16_t)(0x8000u | ((uint32_t)maxsub[p] << 11) | (total >> 2));
                total += 1u << maxsub[p];
            }
        for (int i = 0; i < n; i++) {
            int l = L3_HUFF_LEN[off + i];
            uint16_t leaf = (uint16_t)((l << 8) | L3_HUFF_SYM[off + i]);
            if (l <= HUF_L1) {
                uint32_t base = code[i] >> (32 - HUF_L1), cnt = 1u << (HUF_L1 - l);
                for (uint32_t k = 0; k < cnt; k++) T.huff[T.huff_base[s] + base + k] = leaf;
            } else {
                int p = code[i] >> (32 - HUF_L1), nb = maxsub[p];
                uint32_t sub = (code[i] << HUF_L1) >> (32 - nb), cnt = 1u << (nb - (l - HUF_L1));
                for (uint32_t k = 0; k < cnt; k++) T.huff[suboff[p] + sub + k] = leaf;
            }
        }
    }
    // slot 0 ("no data in this region") = an all-zero first-level table: length 0, x = y = 0, so K1 needs no special case.
    // total must stay within 8192 entries (11-bit offset/4): checked by the caller
    T.huff_base[0] = (uint16_t)total;
    total += 1u << HUF_L1;
    T.huff_total = total;
}

static int decode_init_tables(DecodeScratch& ds) {
    static L3Dev T;  // ~80 KB, host staging
    memset(&T, 0, sizeof T);
    build_huff_lookup(T);
    if (T.huff_total > 8192) return -1;
    memcpy(T.tabsel, L3_TABSEL, sizeof T.tabsel);
    // count1 A: 6-bit peek lookup; B: 4-bit
    for (int v = 0; v < 16; v++) {
        int len = L3_C1_LENS[0][v], bits = L3_C1_BITS[0][v];
        for (int k = 0; k < (1 << (6 - len)); k++) T.c1a_lut[(bits << (6 - len)) + k] = (uint8_t)((len << 4) | v);
    }
    memcpy(T.lsf_nsf, L3_LSF_NSF, sizeof T.lsf_nsf);
    for (int r = 0; r < 9; r++) {
        int a = 0;
        for (int i = 0; i < 22; i++) { T.long_start[r][i] = (uint16_t)a; for (int k = 0; k < L3_SFB_LONG_W[r][i]; k++) T.line_sfb_long[r][a + k] = (uint8_t)i; a += L3_SFB_LONG_W[r][i]; }
        T.long_start[r][22] = (uint16_t)a; T.long_start[r][23] = (uint16_t)a;
        a = 0;
        for (int i = 0; i < 13; i++) { T.short_start[r][i] = (uint16_t)a; a += L3_SFB_SHORT_W[r][i]; }
        T.short_start[r][13] = (uint16_t)a;
        for (int mixed = 0; mixed < 2; mixed++) {
            uint8_t* ls = mixed ? T.line_sfb_mixed[r] : T.line_sfb_short[r];
            uint16_t* ro = mixed ? T.scatter_mixed[r] : T.scatter_short[r];
            for (int i = 0; i < 576; i++) ro[i] = (uint16_t)i;
            // mixed: 8 long bands in MPEG-1, 6 in LSF (what the scalefactor partitions carry): the lines below 36 at every rate
            // but 8 kHz, whose 6 long bands span 72 lines and meet short band 3 there (FFmpeg's compute_band_indexes does the same)
            if (mixed) for (int i = 0; i < T.long_start[r][r < 3 ? 8 : 6]; i++) ls[i] = T.line_sfb_long[r][i];
            for (int sfb = mixed ? 3 : 0; sfb < 13; sfb++) {
                int st = T.short_start[r][sfb], w = L3_SFB_SHORT_W[r][sfb];
                for (int win = 0; win < 3; win++)
                    for (int k = 0; k < w; k++) {
                        int src = 3 * st + win * w + k, dst = 3 * st + 3 * k + win;
                        ls[src] = (uint8_t)((mixed ? 32 : 0) + sfb * 3 + win);
                        ro[src] = (uint16_t)dst;
                    }
            }
        }
    }
    memcpy(T.slen, L3_SLEN, sizeof T.slen);
    memcpy(T.pretab, L3_PRETAB, 22);
    memcpy(T.pow43, L3_POW43, sizeof T.pow43);
    memcpy(T.gain, L3_GAIN, sizeof T.gain);
    memcpy(T.synth_wa, L3_SYNTH_WA, sizeof T.synth_wa);
    memcpy(T.synth_wb, L3_SYNTH_WB, sizeof T.synth_wb);
    memcpy(T.synth_ia, L3_SYNTH_IA, sizeof T.synth_ia);
    memcpy(T.synth_ib, L3_SYNTH_IB, sizeof T.synth_ib);
    for (int r = 0; r < 2; r++) for (int k = 0; k < 7; k++) T.is_ratio[r][k] = L3_IS_RATIO[r][k];
    memcpy(T.is_lsf, L3_IS_LSF, sizeof T.is_lsf);
    if (cudaMalloc(&ds.d_tab, sizeof(L3Dev)) != cudaSuccess) return -3;
    if (cudaMemcpy(ds.d_tab, &T, sizeof(L3Dev), cudaMemcpyHostToDevice) != cudaSuccess) return -3;
    return 0;
}
