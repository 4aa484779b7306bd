/*
 * rg_engine.h -- C ABI of the B200-native ReplayGain 1.0 analysis engine.
 *
 * This is the drop-in boundary for the analysis path of M-Igashi/mp3rgain v1.5.0.  The
 * reference has no FFI of its own: its boundary is the Rust module `mp3rgain::replaygain`
 * (src/replaygain.rs), so each entry point below names the Rust item it stands in for.  A Rust
 * shim maps those pub fns 1:1 onto these symbols.
 *
 * Conventions: plain pointers and sizes, no torch / CUDA types; every call returns 0 (RG_OK) or
 * a negative rg_status and never unwinds; all outputs are caller-allocated; no callbacks
 * (one exception: a host that brings its own transport for the album exchange attaches a function, see rg_exchange_fn).
 * Host pointers unless a parameter is explicitly documented as a DEVICE pointer.
 * The library has no CPU compute path: without a usable CUDA device rg_engine_create fails with
 * RG_ERR_NO_DEVICE and nothing else can be called.
 *
 * Threading: the reference's functions are re-entrant (every call owns its decoder and analyser).  An rg_engine owns one
 * set of streams and device buffers, so calls on the SAME engine serialise on an internal lock (any thread may call at
 * any time; they queue); callers that want calls to overlap create one engine per thread.  rg_last_error is thread-local.
 */
#ifndef MP3RGAIN_B200_RG_ENGINE_H
#define MP3RGAIN_B200_RG_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_HIST_SIZE 12000           /* src/replaygain.rs:630 HISTOGRAM_SIZE */
#define RG_HIST_OFFSET 2000          /* src/replaygain.rs:635 HISTOGRAM_OFFSET */
#define RG_PINK_REF 64.82            /* src/replaygain.rs:44  PINK_REF */
#define RG_REPLAYGAIN_REFERENCE_DB 89.0 /* src/replaygain.rs:37 (display only) */

/* Status codes. The strings returned by rg_status_str() are the reference's anyhow messages
 * (src/replaygain.rs:805,822,834,842-846,857,863,869-872). */
typedef enum {
    RG_OK = 0,
    RG_ERR_INVALID_ARG = -1,
    RG_ERR_NO_DEVICE = -2,          /* no CUDA device / driver: the engine has no CPU fallback */
    RG_ERR_CUDA = -3,               /* a CUDA call failed; rg_last_error() has the detail */
    RG_ERR_OPEN = -4,               /* "Failed to open: {path}"            (host wrappers only) */
    RG_ERR_PROBE = -5,              /* "Failed to probe format"  -- no MPEG audio frame found */
    RG_ERR_NO_AUDIO_TRACK = -6,     /* "No audio track found" */
    RG_ERR_TRACK_INDEX = -7,        /* "Track index {i} out of range (file has 1 audio track(s))" */
    RG_ERR_UNKNOWN_SAMPLE_RATE = -8,/* "Unknown sample rate" */
    RG_ERR_UNSUPPORTED_RATE = -9,   /* "Unsupported sample rate: {sr} Hz. Supported rates: ..." */
    RG_ERR_UNSUPPORTED_FORMAT = -10,/* MP4/AAC container (the engine decodes MPEG-1 / 2 / 2.5 Layer III) */
    RG_ERR_NO_MEMORY = -11,
    RG_ERR_NO_FRAMES = -12          /* "No valid MP3 frames found" (mp3rgain::analyze, src/lib.rs:498-500) */
} rg_status;

/* rg_config.flags */
#define RG_FLAG_EXACT_IIR      0x1u /* bit-exact serial IIR for every track (parity mode).  Default is the
                                       time-blocked IIR which re-runs edge-flagged tracks through the exact kernel. */
#define RG_FLAG_KEEP_INFO_FRAME 0x2u /* decode a leading Xing/Info/VBRI frame as audio instead of skipping it */
#define RG_FLAG_CLAMP_PCM      0x4u /* clamp decoded PCM to [-1,1] before analysis (symphonia-behaviour switch) */

typedef struct rg_engine rg_engine;   /* opaque: CUDA context, streams, device scratch, pinned staging */
typedef struct rg_corpus rg_corpus;   /* opaque: a batch of file images resident in device memory */

typedef struct {
    int32_t  device;            /* CUDA ordinal (one engine per GPU; one process per GPU when sharded) */
    uint32_t flags;             /* RG_FLAG_* */
    uint64_t scratch_bytes;     /* device scratch budget for decoded PCM + intermediates; 0 = default (48 GiB) */
} rg_config;

/* One file image.  Stands in for the `&Path` argument of analyze_track (src/replaygain.rs:929):
 * the caller has read (or mmapped / pinned) the file; the engine borrows the bytes until the call returns.
 * Pinned memory is copied at the full bus rate, pageable memory through the driver's staging.  When the non-empty buffers of a
 * call ascend through memory at most 64 bytes apart, at 16-byte multiples from the first (files read into one arena), the
 * engine copies them in runs and also READS the bytes between them (never writes; rg_engine_set_option("arena", 0) turns
 * that off). */
typedef struct {
    const uint8_t* data;
    uint64_t       len;
    int32_t        track_index; /* -1 = Option::None; MP3 has exactly one audio track, so only -1 and 0 succeed */
    int32_t        reserved;
} rg_input;

/* ReplayGainResult (src/replaygain.rs:56-68) + per-entry status + diagnostics. */
typedef struct {
    double   loudness_db;
    double   gain_db;           /* RG_PINK_REF - loudness_db */
    double   peak;              /* normalised units, not clamped */
    uint32_t sample_rate;
    int32_t  file_type;         /* 0 = Mp3, 1 = Aac (AudioFileType, src/replaygain.rs:47-53) */
    int32_t  status;            /* rg_status for this entry */
    uint32_t channels;
    uint32_t n_frames;          /* MPEG audio frames decoded */
    uint32_t n_windows;         /* 50 ms windows that reached finish_window */
    uint32_t n_dropped_windows; /* windows whose bin index fell outside [0,12000) (src/replaygain.rs:757) */
    uint32_t n_edge_windows;    /* windows whose 1000*log10 value sat within the edge tolerance of an integer */
    uint32_t exact_rerun;       /* 1 if the time-blocked pass flagged this track and the exact kernel re-ran it */
    uint32_t index_rewalks;     /* 64 KiB segments of the speculative frame index that the merge had to walk again (diagnostic) */
} rg_track_result;

/* AlbumGainResult minus the per-track vector (src/replaygain.rs:78-88). */
typedef struct {
    double album_loudness_db;
    double album_gain_db;
    double album_peak;
} rg_album_result;

/* PeakAmplitudeResult (src/replaygain.rs:1125-1132). */
typedef struct {
    double   peak;
    double   peak_pcm;          /* peak * 32768 */
    uint32_t sample_rate;
    int32_t  status;
} rg_peak_result;

/* ---- lifecycle ------------------------------------------------------------------------------- */
/* replaygain::is_available (src/replaygain.rs:1119): 1 when a CUDA device is usable, else 0. */
int  rg_is_available(void);
int  rg_engine_create(const rg_config* cfg, rg_engine** out);
void rg_engine_destroy(rg_engine* e);
const char* rg_status_str(int status);
/* Tuning knobs (not part of the reference surface). Keys: "iir_block_windows" (50 ms windows per time block of
 * the time-blocked IIR; default 0 = chosen per batch so that the equally long chains fill whole waves of the device), "iir_warm_windows" (zero-state warm-up windows before each block, default 1),
 * "edge_tol" (distance of 1000*log10(ms) from an integer below which the time-blocked pass hands the track to the
 * exact kernel; scaled by max(1, 0.1 * input amplitude / filtered RMS) per window, default 1e-9), "pipeline" (default 1:
 * copy / index / decode / analysis overlap on four streams with two alternating batches; 0 = one batch at a time, every
 * kernel alone on the device, for profiling), "group_bytes_min" (smallest copy / index group of the end-to-end path, default
 * 32 MiB), "k4_slim" (default 0; n > 0 launches the time-blocked IIR as waves of n 128-thread CTAs per SM instead of 256-thread
 * CTAs -- an experiment in co-scheduling with the decode kernels that did not pay).  Returns RG_ERR_INVALID_ARG for an unknown key. */
int rg_engine_set_option(rg_engine* e, const char* key, double value);
/* Thread-local detail string for the last failing call (CUDA error text etc.). */
const char* rg_last_error(void);

/* ---- analysis from host buffers (the end-to-end path: H2D copies happen inside) -------------- */
/* analyze_track / analyze_track_with_index (src/replaygain.rs:929-941), batched: n independent files.
 * Per-entry failures are reported in out[i].status (track mode: src/main.rs:1988-1999 keeps going).
 * hist_out: NULL or n*RG_HIST_SIZE u32 counters (the reference's private TrackAnalysisInternal.histogram). */
int rg_analyze_tracks(rg_engine* e, const rg_input* in, size_t n, rg_track_result* out, uint32_t* hist_out);

/* analyze_album / analyze_album_with_index (src/replaygain.rs:1033-1074).  Fails the call with the first
 * bad entry's status (the `?` at :1055).  album_hist_out: NULL or RG_HIST_SIZE u64 counters. */
int rg_analyze_album(rg_engine* e, const rg_input* in, size_t n, rg_track_result* out,
                     rg_album_result* album, uint64_t* album_hist_out);

/* find_peak_amplitude (src/replaygain.rs:1140-1249), batched. */
int rg_find_peak(rg_engine* e, const rg_input* in, size_t n, rg_peak_result* out);

/* process_audio_buffer + analyzer on caller-supplied planar f32 PCM (src/replaygain.rs:953-980, 702-771):
 * the stage-B entry point (decoder bypassed, like the reference's unit tests at :1298-1365).
 * planes[c] = channel c, frames samples each; nch 1 or 2.  hist_out NULL or RG_HIST_SIZE u32. */
int rg_analyze_pcm_f32(rg_engine* e, const float* const* planes, int nch, size_t frames, uint32_t sample_rate,
                       rg_track_result* out, uint32_t* hist_out);

/* ---- device-resident corpora (inputs already in HBM; what the kernel-only throughput times) -- */
int  rg_corpus_upload(rg_engine* e, const rg_input* in, size_t n, rg_corpus** out);
void rg_corpus_free(rg_engine* e, rg_corpus* c);
uint64_t rg_corpus_bytes(const rg_corpus* c);
int rg_analyze_corpus(rg_engine* e, rg_corpus* c, rg_track_result* out, uint32_t* hist_out);

/* ---- album mode across GPUs: one engine per rank, the engine owns the exchange ---------------------------------- */
/* analyze_album (src/replaygain.rs:1033-1074) over a file list that is split across `nranks` processes, one GPU each:
 * every rank analyses its shard and accumulates its tracks' histograms on its device; ONE exchange follows -- all-reduce
 * (sum) of the RG_HIST_SIZE u64 counters and all-reduce(max) of {peak, error key}, one NCCL group on the engine's stream
 * -- and every rank computes the album loudness from the merged histogram.  Every rank returns the same album result (or
 * the same error: a failing file anywhere fails the album everywhere, with the status of the failing file that comes
 * first in the global list, :1055).
 *
 * The communicator: rank 0 calls rg_comm_unique_id and hands the 128 bytes to the other ranks by any means (the host
 * program's own launcher, a file, torch.distributed ...); every rank then calls rg_engine_comm_init (ncclCommInitRank on
 * the engine's device).  A host that already holds an ncclComm_t for this device passes it to rg_engine_comm_attach
 * instead (borrowed, not destroyed by the engine).  NCCL is loaded at run time (libnccl.so.2); a host with another
 * transport attaches an exchange function (rg_engine_comm_attach_host): it is called once per album with the two host
 * buffers and must return them reduced over all ranks (sum / max), 0 on success. */
#define RG_COMM_ID_BYTES 128
typedef int (*rg_exchange_fn)(void* ctx, uint64_t* sum_u64, size_t n_sum, uint64_t* max_u64, size_t n_max);
int rg_comm_unique_id(uint8_t id[RG_COMM_ID_BYTES]);
int rg_engine_comm_init(rg_engine* e, int rank, int nranks, const uint8_t id[RG_COMM_ID_BYTES]);
int rg_engine_comm_attach(rg_engine* e, void* nccl_comm, int rank, int nranks);
int rg_engine_comm_attach_host(rg_engine* e, int rank, int nranks, rg_exchange_fn fn, void* ctx);
int rg_engine_comm_release(rg_engine* e);
/* `in` / `c` = THIS rank's shard (contiguous in the global file list, starting at global index shard_offset).
 * album_hist_out: NULL or RG_HIST_SIZE u64 (the merged histogram).  first_bad_global: NULL or the global index of the file
 * that failed the album (-1 if none).  With nranks == 1 (no communicator) this is rg_analyze_album. */
int rg_analyze_album_sharded(rg_engine* e, const rg_input* in, size_t n, uint64_t shard_offset, rg_track_result* out,
                             rg_album_result* album, uint64_t* album_hist_out, int64_t* first_bad_global);
int rg_analyze_album_sharded_corpus(rg_engine* e, rg_corpus* c, uint64_t shard_offset, rg_track_result* out,
                                    rg_album_result* album, uint64_t* album_hist_out, int64_t* first_bad_global);

/* ---- the same in two steps, for a caller that runs the exchange itself (device buffers) -------------------------- */
/* Analyse this rank's shard and ADD its tracks' histograms into d_album_hist (DEVICE pointer, RG_HIST_SIZE u64,
 * caller-zeroed) and fold its peak into *d_album_peak (DEVICE pointer, one f64, caller-zeroed).  The caller then
 * all-reduces both buffers (sum / max) over NCCL and calls rg_album_finish on every rank: the accumulate loop of
 * src/replaygain.rs:1053-1062 with the cross-device exchange in the middle. */
int rg_analyze_album_shard(rg_engine* e, rg_corpus* c, rg_track_result* out,
                           uint64_t* d_album_hist, double* d_album_peak);
/* get_loudness on the merged histogram + PINK_REF (src/replaygain.rs:1065-1066), computed on the device. */
int rg_album_finish(rg_engine* e, const uint64_t* d_album_hist, const double* d_album_peak, rg_album_result* album);

/* ---- the global_gain fields: mp3rgain::analyze() and apply_gain*() -------------------------- */
/* Mp3Analysis (src/lib.rs:57-75).  The reference's strings are enums here: mpeg_version 0 "MPEG1", 1 "MPEG2",
 * 2 "MPEG2.5" (MpegVersion::as_str, :85-93); channel_mode 0 "Stereo", 1 "Joint Stereo", 2 "Dual Channel", 3 "Mono"
 * (ChannelMode::as_str, :111-118). */
typedef struct {
    uint64_t frame_count;       /* frames the walk counted (Xing/Info frames excluded, :446-451) */
    uint64_t gain_sum;          /* sum / number of the global_gain values read: avg_gain = gain_sum / gain_count */
    uint64_t gain_count;
    double   avg_gain;
    double   headroom_db;       /* headroom_steps * 1.5 */
    int32_t  headroom_steps;    /* 255 - max_gain */
    int32_t  mpeg_version;
    int32_t  channel_mode;
    int32_t  status;            /* RG_OK, or RG_ERR_NO_FRAMES */
    uint8_t  min_gain, max_gain;
    uint8_t  reserved[2];
    uint32_t walk_rewalks;      /* 4 KiB sub-segments (GF_SUB) the merge had to walk again (diagnostic) */
} rg_gain_stats;

/* mp3rgain::analyze (src/lib.rs:470-514), batched over n file images: the reference's own frame walk (iterate_frames,
 * :413-466) run as a speculative segmented walk on the device.  Per-entry status in out[i].status. */
int rg_gain_analyze(rg_engine* e, const rg_input* in, size_t n, rg_gain_stats* out);
/* The same over a device-resident corpus. */
int rg_gain_analyze_corpus(rg_engine* e, rg_corpus* c, rg_gain_stats* out);

/* rg_gain_apply.mode */
#define RG_GAIN_SATURATE 0           /* GainMode::Saturating: apply_gain (src/lib.rs:602-617) */
#define RG_GAIN_WRAP     1           /* GainMode::Wrapping:   apply_gain_wrap (:1232-1246) */
/* apply_gain_to_data (src/lib.rs:544-596) / apply_gain_to_channel_data (:668-735), batched: patches the global_gain of
 * every granule-channel (channel < 0) or of one channel (0 = left, 1 = right; always saturating) by gain_steps and
 * writes the patched image of file i to out_images[i] (in[i].len bytes; may alias in[i].data).  frames_modified[i] =
 * the reference's return value.  gain_steps == 0 mirrors the early return of :603-605: images are copied through
 * unchanged and frames_modified[i] = 0.  The file read / write around it stays in the host wrapper, as in the reference. */
int rg_gain_apply(rg_engine* e, const rg_input* in, size_t n, int32_t gain_steps, int32_t mode, int32_t channel,
                  uint8_t* const* out_images, uint32_t* frames_modified);

/* ---- introspection for measurement and parity tests ------------------------------------------- */
typedef struct {
    uint64_t kernel_launches;   /* kernels launched by this engine since creation */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t tracks, frames, exact_reruns;
    float    last_ms_index, last_ms_huffman, last_ms_hybrid, last_ms_synth, last_ms_iir, last_ms_hist;
    float    last_ms_total;     /* CUDA-event time of the last analyse call's device work */
    uint32_t last_batches;      /* decode/analysis batches (= launches of each of K1, K2, K4, K5) in the last analyse call */
    uint32_t last_block_windows;/* windows per time block the time-blocked IIR used in the last batch (0 = exact mode) */
    uint64_t last_iir_samples;  /* channel-samples the time-blocked IIR ran in the last analyse call, warm-up included */
    float    last_ms_gain_scan, last_ms_gain_merge, last_ms_gain_patch;   /* CUDA-event times of the last rg_gain_* call */
    uint32_t last_gain_groups;  /* copy / walk groups (= launches of each gain-field kernel) in the last rg_gain_* call */
    float    last_ms_exchange;  /* CUDA-event time of the album exchange (the NCCL group) in the last rg_analyze_album_sharded* call */
    uint32_t reserved;
} rg_stats;
int rg_engine_stats(rg_engine* e, rg_stats* out);

/* Decode only: planar f32 PCM of one file image into caller buffers (parity hook for the decoder stages).
 * pcm: nch_max*max_frames floats (plane c at pcm + c*max_frames). Returns samples per channel in *frames_out. */
int rg_decode_pcm(rg_engine* e, const rg_input* in, float* pcm, size_t max_frames,
                  size_t* frames_out, uint32_t* channels_out, uint32_t* sample_rate_out);

#ifdef __cplusplus
}
#endif
#endif /* MP3RGAIN_B200_RG_ENGINE_H */
