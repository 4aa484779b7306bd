// k_gain_fields.cuh -- GF: the reference's first-party walk over the global_gain fields, on the device.
//
// mp3rgain::analyze() (src/lib.rs:470-514) and apply_gain*() (:544-639, :668-735) iterate the frames of a file image with
// their own small walker (iterate_frames, :413-466): skip_id3v2, stop before an ID3v1 / APEv2 trailer (find_audio_end,
// :358-383), accept a header when parse_header (:169-252) succeeds and either a sync word follows the frame or the frame
// ends the audio, jump over Xing/Info frames (:388-408), and read / patch the 8-bit global_gain of every granule-channel
// (calculate_gain_locations :262-296, read_gain_at / write_gain_at :299-340).  The walker keeps NO state besides its
// position, so any two walks that ever stand on the same byte agree from there on.  That makes it exactly parallel:
//
//   k_gf_scan    one warp per 64 KiB segment walks speculatively from the segment's first byte; records the first position
//                it JUMPED from (everything before it was a byte-by-byte scan), where it left the segment, and the gain
//                statistics of the frames it counted
//   k_gf_merge   one warp per track stitches the segments: the true walk enters a segment at P (where its predecessor
//                left); if P <= the segment's first jump, P is on the speculative trajectory and the segment's result
//                stands; otherwise (a false sync inside the straddling frame) the warp walks the segment again from P.
//                Emits Mp3Analysis per track and every segment's entry position
//   k_gf_patch   one warp per segment walks from its entry position (no speculation left) and patches the gain fields
//                in place (saturating / wrapping / one channel only)
//
// A write never changes what the walk reads later: gain bytes lie inside the side info of their own frame (last gain
// byte <= offset 34 of an MPEG-1 frame, <= 18 of an LSF frame; the smallest frames are 96 and 24 bytes) and frames of one
// trajectory are disjoint, so walking first and patching in parallel equals the reference's serial read-modify-write.
//
// The scan streams the segment through a 4 KiB shared-memory window with coalesced 16-byte loads: the kernel reads every
// file byte exactly once, its algorithmic traffic, and is bound by HBM bandwidth / load latency.  The byte-by-byte resync
// is lane-parallel (lane = candidate position, first valid header by ballot).
#pragma once

#define GF_SEG (64u * 1024u)
#define GF_CHUNK 4096
#define GF_WARPS 4
#define GF_NONE 0xFFFFFFFFu
#define GF_SPAN 1480            // pos .. pos+31 (candidates) + largest frame (1441) + next sync (2), rounded up

struct GfSeg {
    uint32_t first_jump;   // first position the walk jumped from by a frame length (GF_NONE: it only scanned)
    uint32_t end_pos;      // first trajectory position at or beyond the segment's end (or where the audio ended)
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
__device__ __forceinline__ void gf_bounds(const uint8_t* __restrict__ d, uint64_t len, uint64_t& a0, uint64_t& audio_end) {
    a0 = 0;
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3')
        a0 = 10 + (((uint64_t)(d[6] & 0x7F) << 21) | ((uint64_t)(d[7] & 0x7F) << 14) | ((uint64_t)(d[8] & 0x7F) << 7) | (uint64_t)(d[9] & 0x7F));
    uint64_t end = len;
    if (end >= 128 && d[end - 128] == 'T' && d[end - 127] == 'A' && d[end - 126] == 'G') end -= 128;
    if (end >= 32) {
        const uint8_t* f = d + end - 32;
        if (f[0] == 'A' && f[1] == 'P' && f[2] == 'E' && f[3] == 'T' && f[4] == 'A' && f[5] == 'G' && f[6] == 'E' && f[7] == 'X') {
            const uint64_t tag_size = (uint64_t)f[12] | ((uint64_t)f[13] << 8) | ((uint64_t)f[14] << 16) | ((uint64_t)f[15] << 24);
            const uint64_t header = (f[23] & 0x80) ? 32 : 0;
            if (end >= tag_size + header) end = end - tag_size - header;
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

// The walk over [pos, bound) of one file image; every lane of the warp executes it with uniform control flow.
// PATCH: rewrite the gain fields in place (img must be writable).  channel < 0: every granule-channel (wrap selects
// GainMode::Wrapping); channel 0 / 1: that channel's fields only, saturating (apply_gain_to_channel_data).
template <bool PATCH>
__device__ GfSeg gf_walk(uint8_t* img, uint64_t len, uint64_t audio_end, uint8_t* ck, uint64_t pos, uint64_t bound, int lane,
                         int steps, int wrap, int channel) {
    GfSeg o;
    o.first_jump = GF_NONE; o.n_frames = 0; o.gcount = 0; o.first_hdr = 0; o.pad = 0;
    uint32_t lsum = 0, lmin = 255, lmax = 0;   // this lane's granule-channel (lanes 0..3)
    uint64_t cs = ~0ull, ce = 0;               // bytes [cs, ce) of the image are in the window
    while (pos + 4 <= audio_end && pos < bound) {
        const uint64_t avail = ce < len ? ce : len;
        if (pos < cs || pos >= ce || (pos + GF_SPAN > avail && avail < len)) {
            cs = pos & ~15ull;
            ce = cs + GF_CHUNK;
            __syncwarp();
            for (int i = lane * 16; i < GF_CHUNK; i += 32 * 16) {
                uint4 v = make_uint4(0, 0, 0, 0);
                const uint64_t g = cs + i;
                if (g + 16 <= len) v = *reinterpret_cast<const uint4*>(img + g);
                else if (g < len) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (int b = 0; b < 16; b++) if (g + b < len) w[b >> 2] |= (uint32_t)img[g + b] << (8 * (b & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                *reinterpret_cast<uint4*>(ck + i) = v;
            }
            __syncwarp();
        }
        // lane-parallel resync: the first of the next 32 positions whose 4 bytes parse as a header
        uint32_t h = 0;
        if (pos + lane + 4 <= audio_end) {
            const uint8_t* p = ck + (pos + lane - cs);
            h = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        }
        const uint32_t cand = __ballot_sync(0xffffffffu, k0_hdr_valid(h));
        if (!cand) { pos += 32; continue; }
        const int l = __ffs(cand) - 1;
        pos += l;
        if (pos >= bound) break;                      // a scan step of the next segment
        h = __shfl_sync(0xffffffffu, h, l);
        const uint64_t next = pos + (uint64_t)k0_frame_len(h);
        bool valid;
        if (next + 2 <= audio_end) { const uint8_t* q = ck + (next - cs); valid = q[0] == 0xFF && (q[1] & 0xE0) == 0xE0; }
        else valid = next <= audio_end;
        if (!valid) { pos++; continue; }
        if (o.first_jump == GF_NONE) o.first_jump = (uint32_t)pos;
        const bool mono = ((h >> 6) & 3) == 3, lsf = ((h >> 19) & 3) != 3;
        const int so = 4 + ((h & 0x10000u) ? 0 : 2);
        const uint64_t x = pos + so + k0_side_len(h);
        if (x + 4 <= len) {
            const uint8_t* p = ck + (x - cs);
            if ((p[0] == 'X' && p[1] == 'i' && p[2] == 'n' && p[3] == 'g') || (p[0] == 'I' && p[1] == 'n' && p[2] == 'f' && p[3] == 'o')) { pos = next; continue; }
        }
        const int nch = mono ? 1 : 2, n = nch * (lsf ? 1 : 2);
        if (lane < n) {
            const int bit = (lsf ? (mono ? 9 : 10) : (mono ? 18 : 20)) + lane * (lsf ? 63 : 59) + 21;
            const uint64_t idx = pos + so + (bit >> 3);
            const int sh = bit & 7;
            uint32_t g = 0, b0 = 0, b1 = 0;
            if (idx < len) {
                const uint8_t* p = ck + (idx - cs);
                b0 = p[0];
                if (sh == 0) g = b0;
                else if (idx + 1 < len) { b1 = p[1]; g = ((b0 << sh) | (b1 >> (8 - sh))) & 0xFF; }
                else g = (b0 << sh) & 0xFF;
            }
            lsum += g; lmin = min(lmin, g); lmax = max(lmax, g);
            if (PATCH && idx < len && (channel < 0 || (lane >= channel && (lane - channel) % nch == 0))) {
                const uint32_t v = gf_adjust(g, steps, channel < 0 ? wrap : 0);
                if (sh == 0) img[idx] = (uint8_t)v;
                else {
                    img[idx] = (uint8_t)((b0 & (0xFFu << (8 - sh))) | (v >> sh));
                    if (idx + 1 < len) img[idx + 1] = (uint8_t)((b1 & (0xFFu >> sh)) | ((v << (8 - sh)) & 0xFF));
                }
            }
        }
        if (o.n_frames == 0) o.first_hdr = h;
        o.n_frames++;
        o.gcount += n;
        pos = next;
    }
    o.end_pos = (uint32_t)(pos < 0xFFFFFFFFull ? pos : 0xFFFFFFFFull);
    o.gsum = __reduce_add_sync(0xffffffffu, lsum);
    o.gminmax = __reduce_min_sync(0xffffffffu, lmin) | (__reduce_max_sync(0xffffffffu, lmax) << 8);
    return o;
}

// byte range [b, e) of segment `seg` (the last one is open-ended: the walk stops at the end of the audio)
__device__ __forceinline__ void gf_seg_range(uint32_t seg, uint32_t n_seg, uint64_t& b, uint64_t& e) {
    b = (uint64_t)seg * GF_SEG;
    e = (seg + 1 == n_seg) ? ~0ull : (uint64_t)(seg + 1) * GF_SEG;
}

__global__ void __launch_bounds__(32 * GF_WARPS)
k_gf_scan(uint8_t* corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg, const SegDesc* __restrict__ segs,
          uint32_t seg0, uint32_t n_segs, GfSeg* __restrict__ segout) {
    __shared__ __align__(16) uint8_t s_chunk[GF_WARPS][GF_CHUNK + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t si = blockIdx.x * GF_WARPS + warp;
    if (si >= n_segs) return;
    const SegDesc sd = segs[seg0 + si];
    uint8_t* img = corpus + tin[sd.track].off;
    const uint64_t len = tin[sd.track].len;
    uint64_t a0, audio_end, b, e;
    gf_bounds(img, len, a0, audio_end);
    gf_seg_range(sd.seg, tseg[sd.track].n_seg, b, e);
    if (b < a0) b = a0;
    GfSeg o;
    if (b >= e) { o.first_jump = GF_NONE; o.end_pos = 0; o.n_frames = 0; o.gsum = 0; o.gcount = 0; o.gminmax = 255; o.first_hdr = 0; o.pad = 0; }
    else o = gf_walk<false>(img, len, audio_end, s_chunk[warp], b, e, lane, 0, 0, -1);
    if (lane == 0) segout[seg0 + si] = o;
}

__global__ void __launch_bounds__(32 * GF_WARPS)
k_gf_merge(uint8_t* corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg, const GfSeg* __restrict__ segout,
           uint32_t* __restrict__ entry, GfTrackOut* __restrict__ out, uint32_t n_tracks) {
    __shared__ __align__(16) uint8_t s_chunk[GF_WARPS][GF_CHUNK + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * GF_WARPS + warp;
    if (t >= n_tracks) return;
    uint8_t* img = corpus + tin[t].off;
    const uint64_t len = tin[t].len;
    const TrackSeg ts = tseg[t];
    uint64_t a0, audio_end;
    gf_bounds(img, len, a0, audio_end);
    uint64_t P = a0;
    GfTrackOut r;
    r.frame_count = 0; r.first_hdr = 0; r.n_rewalks = 0; r.gain_sum = 0; r.gain_count = 0;
    uint32_t gmin = 255, gmax = 0;
    for (uint32_t s = 0; s < ts.n_seg; s++) {
        uint64_t b, e;
        gf_seg_range(s, ts.n_seg, b, e);
        uint32_t ent = GF_NONE;
        if (P < e && P + 4 <= audio_end) {
            GfSeg o = segout[ts.seg_base + s];
            if (!(o.first_jump == GF_NONE || P <= (uint64_t)o.first_jump)) {   // P is not on the speculative trajectory
                o = gf_walk<false>(img, len, audio_end, s_chunk[warp], P, e, lane, 0, 0, -1);
                r.n_rewalks++;
            }
            ent = (uint32_t)P;
            if (r.frame_count == 0 && o.n_frames) r.first_hdr = o.first_hdr;
            r.frame_count += o.n_frames;
            r.gain_sum += o.gsum; r.gain_count += o.gcount;
            if (o.n_frames) { gmin = min(gmin, o.gminmax & 0xFF); gmax = max(gmax, o.gminmax >> 8); }
            P = o.end_pos;
        }
        if (lane == 0) entry[ts.seg_base + s] = ent;
    }
    r.gminmax = gmin | (gmax << 8);
    if (lane == 0) out[t] = r;
}

__global__ void __launch_bounds__(32 * GF_WARPS)
k_gf_patch(uint8_t* corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg, const SegDesc* __restrict__ segs,
           uint32_t seg0, uint32_t n_segs, const uint32_t* __restrict__ entry, int steps, int wrap, int channel) {
    __shared__ __align__(16) uint8_t s_chunk[GF_WARPS][GF_CHUNK + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t si = blockIdx.x * GF_WARPS + warp;
    if (si >= n_segs) return;
    const uint32_t ent = entry[seg0 + si];
    if (ent == GF_NONE) return;
    const SegDesc sd = segs[seg0 + si];
    uint8_t* img = corpus + tin[sd.track].off;
    const uint64_t len = tin[sd.track].len;
    uint64_t a0, audio_end, b, e;
    gf_bounds(img, len, a0, audio_end);
    gf_seg_range(sd.seg, tseg[sd.track].n_seg, b, e);
    gf_walk<true>(img, len, audio_end, s_chunk[warp], ent, e, lane, steps, wrap, channel);
}
