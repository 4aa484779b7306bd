// k_gain_fields.cuh -- GF: the reference's first-party walk over the global_gain fields, on the device.
//
// mp3rgain::analyze() (src/lib.rs:470-514) and apply_gain*() (:544-639, :668-735) iterate the frames of a file image with
// their own small walker (iterate_frames, :413-466): skip_id3v2, stop before an ID3v1 / APEv2 trailer (find_audio_end,
// :358-383), accept a header when parse_header (:169-252) succeeds and either a sync word follows the frame or the frame
// ends the audio, jump over Xing/Info frames (:388-408), and read / patch the 8-bit global_gain of every granule-channel
// (calculate_gain_locations :262-296, read_gain_at / write_gain_at :299-340).  The walker keeps NO state besides its
// position, so any two walks that ever stand on the same byte agree from there on.  That makes it exactly parallel:
//
//   k_gf_scan    one THREAD per 4 KiB sub-segment walks speculatively from the sub-segment's first byte; records the first
//                position it JUMPED from (everything before it was a byte-by-byte scan), where it left the sub-segment,
//                and the gain statistics of the frames it counted
//   k_gf_merge   one warp per track stitches the sub-segments, 32 at a time (lane = sub-segment, a serial resolve over the
//                lanes by shuffles): the true walk enters a sub-segment at P (where its predecessor left); if P <= the
//                sub-segment's first jump, P is on the speculative trajectory and the result stands; otherwise (a false
//                sync inside the straddling frame) that lane walks the sub-segment again from P.  Emits Mp3Analysis per
//                track and every sub-segment's entry position
//   k_gf_patch   one thread per sub-segment walks the ORIGINAL image from its entry position (no speculation left) and writes
//                the patched gain fields (saturating / wrapping / one channel only) into a COPY of the image
//
// Why a copy: the reference patches in place while it walks, but its cursor only ever reads bytes at or ahead of itself and
// only ever writes inside frames it has passed, so every decision it takes is a decision about the original bytes.  A
// parallel in-place patch does not have that property: a false sync at the end of one sub-segment is confirmed or rejected
// by the two bytes at its frame length, which may be gain bytes that the neighbouring sub-segment's thread has already
// rewritten (saturation to 255 manufactures 0xFF 0xE? patterns).  tests/dev/gf_kernel_model.c found such a case in its
// 10,000th image; reading the original and writing the copy makes every walk see what the reference's cursor saw.
// Within the copy no byte is written twice: gain fields of one frame are 59 / 63 bits apart (different bytes), lie inside
// the side info of their own frame (last gain byte <= offset 34 of an MPEG-1 frame, <= 18 of an LSF frame; the smallest
// frames are 96 and 24 bytes), and the frames of one trajectory are disjoint.
//
// The walker touches ~48 bytes per frame (header, CRC, side info, Xing marker, the next sync word), not the payload.  A
// first version streamed every file byte through shared memory, one warp per 64 KiB segment, and was bound by instruction
// issue (~185 warp-instructions per frame, 2.1 TB/s of file bytes).  With one thread per sub-segment the same ~100
// instructions per frame serve 32 sub-segments at once, the loads go straight to the 32-byte sectors that hold the headers
// (about a quarter of the file's sectors), and the kernel is bound by scattered-sector HBM bandwidth instead.
#pragma once

#define GF_SUB 4096u            // bytes per speculative sub-segment (one thread each)
#define GF_THREADS 256
#define GF_NONE 0xFFFFFFFFu

struct GfSeg {
    uint32_t first_jump;   // first position the walk jumped from by a frame length (GF_NONE: it only scanned)
    uint32_t end_pos;      // first trajectory position at or beyond the sub-segment's end (or where the audio ended)
    uint32_t n_frames;     // counted frames (Xing/Info frames are jumped over, not counted)
    uint32_t gsum, gcount; // sum / number of the global_gain values read
    uint32_t gminmax;      // min | max << 8
    uint32_t first_hdr;    // header of the first counted frame
    uint32_t pad;
};
struct GfTrackOut {
    uint32_t frame_count, first_hdr, gminmax, n_rewalks;
    unsigned long long gain_sum, gain_count;
};

// skip_id3v2 (src/lib.rs:343-354) and find_audio_end (:358-383; APE_FLAG_HEADER_PRESENT = 1 << 31, :844)
__device__ __forceinline__ void gf_bounds(const uint8_t* __restrict__ d, uint32_t len, uint32_t& a0, uint32_t& audio_end) {
    a0 = 0;
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3')
        a0 = 10 + (((uint32_t)(d[6] & 0x7F) << 21) | ((uint32_t)(d[7] & 0x7F) << 14) | ((uint32_t)(d[8] & 0x7F) << 7) | (uint32_t)(d[9] & 0x7F));
    uint32_t end = len;
    if (end >= 128 && d[end - 128] == 'T' && d[end - 127] == 'A' && d[end - 126] == 'G') end -= 128;
    if (end >= 32) {
        const uint8_t* f = d + end - 32;
        if (f[0] == 'A' && f[1] == 'P' && f[2] == 'E' && f[3] == 'T' && f[4] == 'A' && f[5] == 'G' && f[6] == 'E' && f[7] == 'X') {
            const uint64_t tag_size = (uint64_t)f[12] | ((uint64_t)f[13] << 8) | ((uint64_t)f[14] << 16) | ((uint64_t)f[15] << 24);
            const uint64_t header = (f[23] & 0x80) ? 32 : 0;
            if ((uint64_t)end >= tag_size + header) end = (uint32_t)(end - tag_size - header);
        }
    }
    audio_end = end;
}

// adjust_gain_value (src/lib.rs:526-541)
__device__ __forceinline__ uint32_t gf_adjust(uint32_t cur, int steps, int wrap) {
    if (!wrap) {
        if (steps > 0) return min(255u, cur + (uint32_t)min(steps, 255));
        const int a = (int)min((long long)255, -(long long)steps);
        return (uint32_t)max(0, (int)cur - a);
    }
    const int ng = (int)(((long long)cur + steps) % 256);
    return (uint32_t)((ng + 256) % 256);
}

// Frame length without padding by (version, bitrate index, sample-rate index); 0 = not a valid Layer III header
// (parse_header, src/lib.rs:169-252).  Index: version << 6 | bitrate index << 2 | sample-rate index.
__device__ __forceinline__ uint32_t gf_lut_index(uint32_t h) { return ((h >> 13) & 0xC0u) | ((h >> 10) & 0x3Fu); }
__device__ __forceinline__ void gf_fill_lut(uint16_t* lut) {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t ver = i >> 6, br = (i >> 2) & 15, sr = i & 3;
        const uint32_t h = 0xFFE20000u | (ver << 19) | (br << 12) | (sr << 10);
        lut[i] = k0_hdr_valid(h) ? (uint16_t)k0_frame_len(h) : (uint16_t)0;
    }
    __syncthreads();
}

// Image loads.  Images start 16-byte aligned and are followed by >= 16 bytes of slack inside the corpus buffer
// (corpus_describe), so an aligned 16-byte load that starts below `len`, and the word after it, stay inside the allocation; bytes beyond `len` are
// never consumed by the walk (every read is guarded by len / audio_end).  The image is read-only for every kernel here
// (the patch goes to a copy), so all loads take the non-coherent path.
template <bool PATCH> __device__ __forceinline__ uint32_t gf_ldw(const uint8_t* img, uint32_t a) {   // a % 4 == 0
    return __ldg(reinterpret_cast<const uint32_t*>(img + a));
}
template <bool PATCH> __device__ __forceinline__ uint4 gf_ld16(const uint8_t* img, uint32_t a) {     // a % 16 == 0
    return __ldg(reinterpret_cast<const uint4*>(img + a));
}
// 4 bytes at any offset g, little-endian (first byte in bits 0-7)
template <bool PATCH> __device__ __forceinline__ uint32_t gf_ld32(const uint8_t* img, uint32_t g) {
    const uint32_t a = g & ~3u;
    return __funnelshift_r(gf_ldw<PATCH>(img, a), gf_ldw<PATCH>(img, a + 4), 8 * (g & 3));
}

// The walk over [pos, bound) of one file image by ONE thread -- but written for the warp: all 32 lanes call it together
// (`run` = this lane has a sub-segment to walk) and move through two warp-synchronous phases, so that lanes at different
// points of their walks still execute the same instructions:
//   SCAN   every scanning lane advances 16 aligned bytes per step: bytes that are 0xFF and followed by a byte with its top
//          three bits set are the only positions parse_header (src/lib.rs:176-178) can accept; such a candidate gets the
//          full header + next-sync check and either becomes the lane's current frame or the scan resumes one byte later
//   FRAME  every lane that stands on a validated frame processes it (Xing/Info test, gain fields), jumps, and validates
//          the header it lands on; a lane that lands on something else goes back to SCAN (one byte later, as the
//          reference does) once the others have finished their frames
// A first thread-per-sub-segment version without this structure ran with 5 of 32 lanes active per instruction.
// PATCH: write the rewritten gain fields to `outp` (a copy of the image; img itself is only read).  channel < 0: every
// granule-channel (wrap selects GainMode::Wrapping); channel 0 / 1: that channel's fields only, saturating
// (apply_gain_to_channel_data).
#define GF_ST_SCAN 0u
#define GF_ST_FRAME 1u
#define GF_ST_DONE 2u
template <bool PATCH>
__device__ GfSeg gf_walk(const uint8_t* img, uint8_t* outp, uint32_t len, uint32_t audio_end, const uint16_t* lut, uint32_t pos,
                         uint32_t bound, int steps, int wrap, int channel, bool run) {
    GfSeg o;
    o.first_jump = GF_NONE; o.n_frames = 0; o.gcount = 0; o.first_hdr = 0; o.pad = 0;
    uint32_t gsum = 0, gmin = 255, gmax = 0;
    uint32_t h = 0, next = 0;                        // FRAME state: header at pos, start of the following frame
    auto in_range = [&](uint32_t p) { return (uint64_t)p + 4 <= audio_end && p < bound; };
    // header + validity check of the reference's loop body (src/lib.rs:424-443) at p; fills h / next
    auto check = [&](uint32_t p) -> bool {
        h = __byte_perm(gf_ld32<PATCH>(img, p), 0, 0x0123);
        const uint32_t f = lut[gf_lut_index(h)];
        if ((h & 0xFFE60000u) != 0xFFE20000u || f == 0) return false;
        next = p + f + ((h >> 9) & 1);
        if (next + 2 <= audio_end) return (gf_ld32<PATCH>(img, next) & 0xE0FFu) == 0xE0FFu;
        return next <= audio_end;
    };
    uint32_t st = (run && in_range(pos)) ? GF_ST_SCAN : GF_ST_DONE;
    while (__any_sync(0xffffffffu, st != GF_ST_DONE)) {
        while (__any_sync(0xffffffffu, st == GF_ST_SCAN)) {
            if (st == GF_ST_SCAN) {
                // 16 aligned bytes per step.  Bit k of m: byte k is 0xFF and byte k+1 has its top three bits set
                const uint32_t a = pos & ~15u;
                const uint4 v = gf_ld16<PATCH>(img, a);
                const uint32_t w[5] = {v.x, v.y, v.z, v.w, gf_ldw<PATCH>(img, a + 16)};
                uint32_t m = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t ff = ((w[k] & 0x7F7F7F7Fu) + 0x01010101u) & w[k];       // bit 7 of byte i: byte i == 0xFF
                    const uint32_t u = __funnelshift_r(w[k], w[k + 1], 8);                  // byte i = byte i+1 of the stream
                    const uint32_t c = (ff & u & (u << 1) & (u << 2) & 0x80808080u) >> 7;    // bit 0 of byte i: candidate
                    m |= ((c * 0x01020408u) >> 24) << (4 * k);                               // gather the four flags
                }
                m &= ~((1u << (pos & 15u)) - 1u);                                            // at or after pos
                if (!m) pos = a + 16;
                else {
                    pos = a + (__ffs(m) - 1);
                    if (in_range(pos)) {
                        if (check(pos)) st = GF_ST_FRAME;
                        else pos++;
                    }
                }
                if (st == GF_ST_SCAN && !in_range(pos)) st = GF_ST_DONE;   // a scan step of the next sub-segment / the end of the audio
            }
        }
        while (__any_sync(0xffffffffu, st == GF_ST_FRAME)) {
            if (st == GF_ST_FRAME) {
                if (o.first_jump == GF_NONE) o.first_jump = pos;
                const bool mono = ((h >> 6) & 3) == 3, lsf = ((h >> 19) & 3) != 3;
                const uint32_t so = (h & 0x10000u) ? 4 : 6;
                const uint32_t x = pos + so + (lsf ? (mono ? 9 : 17) : (mono ? 17 : 32));
                uint32_t m = 0;
                if (x + 4 <= len) m = gf_ld32<PATCH>(img, x);
                if (!(x + 4 <= len && (m == 0x676E6958u /* "Xing" */ || m == 0x6F666E49u /* "Info" */))) {
                    const int nch = mono ? 1 : 2, n = lsf ? nch : 2 * nch;
                    const uint32_t bit0 = (lsf ? (mono ? 9u : 10u) : (mono ? 18u : 20u)) + 21u, per = lsf ? 63u : 59u;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k < n) {
                            const uint32_t bit = bit0 + k * per;
                            const uint32_t idx = pos + so + (bit >> 3), sh = bit & 7;
                            uint32_t g = 0, b0 = 0, b1 = 0;
                            if (idx < len) {
                                const uint32_t v = gf_ld32<PATCH>(img, idx);
                                b0 = v & 0xFF;
                                if (sh == 0) g = b0;
                                else if (idx + 1 < len) { b1 = (v >> 8) & 0xFF; g = ((b0 << sh) | (b1 >> (8 - sh))) & 0xFF; }
                                else g = (b0 << sh) & 0xFF;
                            }
                            gsum += g; gmin = min(gmin, g); gmax = max(gmax, g);
                            if (PATCH && idx < len && (channel < 0 || (k >= channel && (k - channel) % nch == 0))) {
                                const uint32_t v = gf_adjust(g, steps, channel < 0 ? wrap : 0);
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
                    o.gcount += n;
                }
                pos = next;
                if (!in_range(pos)) st = GF_ST_DONE;
                else if (!check(pos)) { pos++; st = in_range(pos) ? GF_ST_SCAN : GF_ST_DONE; }
            }
        }
    }
    o.end_pos = pos;
    o.gsum = gsum;
    o.gminmax = gmin | (gmax << 8);
    return o;
}

// byte range [b, e) of sub-segment `seg` (the last one is open-ended: the walk stops at the end of the audio)
__device__ __forceinline__ void gf_seg_range(uint32_t seg, uint32_t n_seg, uint32_t& b, uint32_t& e) {
    b = seg * GF_SUB;
    e = (seg + 1 == n_seg) ? 0xFFFFFFFFu : (seg + 1) * GF_SUB;
}
// the track of [t0, t0 + nt) that owns sub-segment s (TrackSeg.seg_base ascending)
__device__ __forceinline__ uint32_t gf_find_track(const TrackSeg* __restrict__ tseg, uint32_t nt, uint32_t s) {
    uint32_t lo = 0, hi = nt;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tseg[mid].seg_base <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// tin / tseg point at the first track of the group; seg0 = its first sub-segment
__global__ void __launch_bounds__(GF_THREADS)
k_gf_scan(const uint8_t* corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg, uint32_t n_tracks, uint32_t seg0,
          uint32_t n_segs, GfSeg* __restrict__ segout) {
    __shared__ uint16_t s_lut[256];
    gf_fill_lut(s_lut);
    const uint32_t si = blockIdx.x * GF_THREADS + threadIdx.x;
    const bool have = si < n_segs;
    const uint32_t t = gf_find_track(tseg, n_tracks, seg0 + (have ? si : 0));
    const TrackSeg ts = tseg[t];
    const uint8_t* img = corpus + tin[t].off;
    const uint32_t len = (uint32_t)tin[t].len;
    uint32_t a0, audio_end, b, e;
    gf_bounds(img, len, a0, audio_end);
    gf_seg_range(seg0 + (have ? si : 0) - ts.seg_base, ts.n_seg, b, e);
    if (b < a0) b = a0;
    const GfSeg o = gf_walk<false>(img, nullptr, len, audio_end, s_lut, b, e, 0, 0, -1, have && b < e);
    if (have) segout[seg0 + si] = o;
}

__global__ void __launch_bounds__(GF_THREADS)
k_gf_merge(const uint8_t* corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg, const GfSeg* __restrict__ segout,
           uint32_t* __restrict__ entry, GfTrackOut* __restrict__ out, uint32_t n_tracks) {
    __shared__ uint16_t s_lut[256];
    gf_fill_lut(s_lut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * (GF_THREADS / 32) + warp;
    if (t >= n_tracks) return;
    const uint8_t* img = corpus + tin[t].off;
    const uint32_t len = (uint32_t)tin[t].len;
    const TrackSeg ts = tseg[t];
    uint32_t a0, audio_end;
    gf_bounds(img, len, a0, audio_end);
    uint32_t P = a0;                                      // where the true walk stands (warp-uniform)
    unsigned long long frames = 0, gsum = 0, gcount = 0;  // this lane's sub-segments
    uint32_t gmin = 255, gmax = 0, rewalks = 0, first_s = GF_NONE, first_hdr = 0;
    for (uint32_t base = 0; base < ts.n_seg; base += 32) {
        const uint32_t s = base + lane;
        GfSeg o;
        o.first_jump = GF_NONE; o.end_pos = 0; o.n_frames = 0; o.gsum = 0; o.gcount = 0; o.gminmax = 255; o.first_hdr = 0; o.pad = 0;
        if (s < ts.n_seg) o = segout[ts.seg_base + s];
        uint32_t ent = GF_NONE;
        const uint32_t nj = min(32u, ts.n_seg - base);
        for (uint32_t j = 0; j < nj; j++) {               // serial resolve: lane j's sub-segment against the current P
            uint32_t b, e;
            gf_seg_range(base + j, ts.n_seg, b, e);
            const bool active = P < e && (uint64_t)P + 4 <= audio_end;
            const uint32_t fj = __shfl_sync(0xffffffffu, o.first_jump, j);
            if (active && !(fj == GF_NONE || P <= fj)) {  // P is not on the speculative trajectory: lane j walks again from P
                const GfSeg w = gf_walk<false>(img, nullptr, len, audio_end, s_lut, P, e, 0, 0, -1, lane == (int)j);
                if (lane == (int)j) { o = w; rewalks++; }
            }
            if (lane == (int)j) {
                if (active) ent = P;
                else { o.n_frames = 0; o.gsum = 0; o.gcount = 0; }
            }
            const uint32_t ep = __shfl_sync(0xffffffffu, o.end_pos, j);
            if (active) P = ep;
        }
        if (s < ts.n_seg) entry[ts.seg_base + s] = ent;
        if (o.n_frames) {
            if (first_s == GF_NONE) { first_s = s; first_hdr = o.first_hdr; }
            frames += o.n_frames; gsum += o.gsum; gcount += o.gcount;
            gmin = min(gmin, o.gminmax & 0xFF); gmax = max(gmax, o.gminmax >> 8);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        frames += __shfl_xor_sync(0xffffffffu, frames, d);
        gsum += __shfl_xor_sync(0xffffffffu, gsum, d);
        gcount += __shfl_xor_sync(0xffffffffu, gcount, d);
    }
    gmin = __reduce_min_sync(0xffffffffu, gmin);
    gmax = __reduce_max_sync(0xffffffffu, gmax);
    rewalks = __reduce_add_sync(0xffffffffu, rewalks);
    const uint32_t fs = __reduce_min_sync(0xffffffffu, first_s);   // the sub-segment holding the first counted frame
    const uint32_t src = __ffs(__ballot_sync(0xffffffffu, first_s == fs)) - 1;
    first_hdr = __shfl_sync(0xffffffffu, first_hdr, src);
    if (lane == 0) {
        GfTrackOut r;
        r.frame_count = (uint32_t)frames; r.first_hdr = fs == GF_NONE ? 0 : first_hdr; r.gminmax = gmin | (gmax << 8); r.n_rewalks = rewalks;
        r.gain_sum = gsum; r.gain_count = gcount;
        out[t] = r;
    }
}

__global__ void __launch_bounds__(GF_THREADS)
k_gf_patch(const uint8_t* corpus, uint8_t* out_corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg, uint32_t n_tracks, uint32_t seg0,
           uint32_t n_segs, const uint32_t* __restrict__ entry, int steps, int wrap, int channel) {
    __shared__ uint16_t s_lut[256];
    gf_fill_lut(s_lut);
    const uint32_t si = blockIdx.x * GF_THREADS + threadIdx.x;
    const bool have = si < n_segs;
    const uint32_t ent = have ? entry[seg0 + si] : GF_NONE;
    const uint32_t t = gf_find_track(tseg, n_tracks, seg0 + (have ? si : 0));
    const TrackSeg ts = tseg[t];
    const uint8_t* img = corpus + tin[t].off;
    const uint32_t len = (uint32_t)tin[t].len;
    uint32_t a0, audio_end, b, e;
    gf_bounds(img, len, a0, audio_end);
    gf_seg_range(seg0 + (have ? si : 0) - ts.seg_base, ts.n_seg, b, e);
    gf_walk<true>(img, out_corpus + tin[t].off, len, audio_end, s_lut, ent, e, steps, wrap, channel, ent != GF_NONE);
}
