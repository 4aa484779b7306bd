/*
 * gain_fields_oracle.h -- CPU ORACLE (test infrastructure, NOT product code): the reference's frame walk over the
 * global_gain fields (M-Igashi/mp3rgain v1.5.0 src/lib.rs:169-639, 668-735).  See gain_fields_oracle.c.
 */
#ifndef GAIN_FIELDS_ORACLE_H
#define GAIN_FIELDS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mp3Analysis (src/lib.rs:57-75) with the strings as enums: mpeg_version 0 "MPEG1", 1 "MPEG2", 2 "MPEG2.5";
 * channel_mode 0 "Stereo", 1 "Joint Stereo", 2 "Dual Channel", 3 "Mono". */
typedef struct {
    uint64_t frame_count;
    uint64_t gain_sum, gain_count;
    double avg_gain, headroom_db;
    int32_t headroom_steps;
    int32_t mpeg_version, channel_mode;
    uint8_t min_gain, max_gain;
} gfo_stats;

int gfo_analyze(const uint8_t* image, size_t len, gfo_stats* out);
size_t gfo_apply_gain(uint8_t* image, size_t len, int32_t steps, int wrap, int channel);

/* hooks for the reference's own unit tests (src/lib.rs:1344-1442) */
int gfo_parse_header(const uint8_t* h4, int* version, int* channel_mode, size_t* frame_size);
uint8_t gfo_read_gain_at(const uint8_t* d, size_t len, size_t byte_offset, unsigned bit_offset);
void gfo_write_gain_at(uint8_t* d, size_t len, size_t byte_offset, unsigned bit_offset, uint8_t value);
size_t gfo_skip_id3v2(const uint8_t* d, size_t len);
size_t gfo_find_audio_end(const uint8_t* d, size_t len);
int gfo_is_xing_frame(const uint8_t* d, size_t len, size_t pos);
uint8_t gfo_adjust_gain_value(uint8_t cur, int32_t steps, int wrap);

#ifdef __cplusplus
}
#endif
#endif
