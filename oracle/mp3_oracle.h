/*
 * mp3_oracle.h -- CPU ORACLE, stage A (test infrastructure, NOT product code).  See mp3_oracle.c.
 */
#ifndef MP3_ORACLE_H
#define MP3_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "rg_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

#define MP3O_FLAG_KEEP_INFO_FRAME 0x2u /* same bit values as RG_FLAG_* in include/rg_engine.h */
#define MP3O_FLAG_CLAMP_PCM 0x4u
#define MP3O_ERR_PROBE (-5)
#define MP3O_ERR_UNSUPPORTED_FORMAT (-10)

typedef struct {
    int32_t status;
    uint32_t sample_rate, channels, n_frames;
    uint64_t n_samples;           /* per channel */
    uint32_t skipped_info_frame;  /* 1 if a leading Xing/Info/VBRI frame was swallowed */
    uint32_t n_dropped_frames;    /* frames dropped as DecodeError (reservoir underflow / invalid side info) */
} mp3o_info;

/* One PCM-producing frame of the index: where it starts in the file, where its granule data starts / how many bytes the
 * concatenated main-data stream holds once the frame is read, and its 4 header bytes (big-endian). */
typedef struct { uint32_t file_off, md_start, hdr, md_end; } mp3o_frame;
/* The demux half of mp3o_decode only.  md: len + 64 bytes. */
int mp3o_index(const uint8_t* data, size_t len, uint32_t flags, mp3o_frame* recs, size_t max_recs, uint8_t* md, size_t* md_len_out,
               mp3o_info* info);

/* MP3 file image -> planar f32 PCM.  pcm holds 2 planes of max_frames floats (plane 1 unused for mono). */
int mp3o_decode(const uint8_t* data, size_t len, uint32_t flags, float* pcm, size_t max_frames, mp3o_info* info);
/* analyze_track_internal (src/replaygain.rs:796-925): decode + stage B. hist_out NULL or 12000 u32. */
int mp3o_analyze(const uint8_t* data, size_t len, uint32_t flags, rgo_result* out, uint32_t* hist_out, mp3o_info* info);

#ifdef __cplusplus
}
#endif
#endif
