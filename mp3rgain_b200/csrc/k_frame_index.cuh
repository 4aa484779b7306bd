// k_frame_index.cuh -- K0: frame index of every track (demux + bit-reservoir bookkeeping) and main-data compaction.
//
// Finding frame n+1 needs frame n's header (padding bit, bitrate), so a track is one serial chain -- ~6.9 k frames for three
// minutes, ~1 us each for a lone warp.  The chain is cut into 64 KiB SEGMENTS that are walked SPECULATIVELY in parallel:
//
//   k_index_scan      one warp per segment: (re)synchronise at the segment's first byte exactly as the serial walk does after
//                     losing sync (valid header confirmed by a compatible next header), then walk frames until one starts
//                     beyond the segment; stage {position, header, main_data_begin, payload size, validity} per frame
//   k_index_merge     one warp per track: stitches the segments in order.  A segment is accepted if it starts where its
//                     predecessor's walk ended (and agrees on the stream's version / rate / channel mode); otherwise --
//                     a false sync inside a frame's payload, ~1e-5 of segments -- the warp re-walks just that segment from
//                     the known position.  The result is exactly the serial walk's.  Emits per-segment bases (frame number,
//                     main-data offset) and the track's TrackMeta
//   k_index_finalize  one warp per segment: prefix sums over the staged frames -> FrameRec; frames dropped from the index
//                     (invalid side info, reservoir underflow) still feed the reservoir and are copied here
//   k_md_gather       (decode stream, per batch) one warp per frame: main data -> the track's contiguous stream
//
// The walk streams the file through an 8 KiB shared-memory window (coalesced 16-byte loads) and is warp-uniform scalar code
// (every lane reads the same bytes: smem broadcasts); it only touches the 4 header bytes and the side info of each frame.
// Replaces symphonia's MpaReader + bit reservoir (reference call sites src/replaygain.rs:815-822, 882); the header
// arithmetic matches the reference's own frame walker (src/lib.rs:169-252: frame_size = 144*kbps*1000/sr + pad).
#pragma once

#define K0_CHUNK 4096
#define K0_WARPS 4

__device__ __forceinline__ int k0_frame_len(uint32_t h) {
    // h = header bytes big-endian, already validated by k0_hdr_valid
    const int ver = (h >> 19) & 3, br_idx = (h >> 12) & 15, sr_idx = (h >> 10) & 3, pad = (h >> 9) & 1;
    const int sr1 = sr_idx == 0 ? 44100 : (sr_idx == 1 ? 48000 : 32000);
    if (ver == 3) {
        const int kbps = (br_idx < 5) ? 24 + 8 * br_idx : (br_idx < 9) ? 16 * br_idx - 16 : (br_idx < 13) ? 32 * br_idx - 160 : (br_idx == 13 ? 256 : 320);
        return 144 * kbps * 1000 / sr1 + pad;
    }
    // MPEG-2 / 2.5 (only to recognise the stream; it is then reported as unsupported)
    const int kbps = (br_idx < 9) ? 8 * br_idx : (br_idx < 13 ? 16 * br_idx - 64 : 144 + 16 * (br_idx - 13));
    return 72 * kbps * 1000 / (sr1 >> (ver == 2 ? 1 : 2)) + pad;
}

// bytes of side info: MPEG-1 17 / 32, MPEG-2 and 2.5 (LSF) 9 / 17 (mono / not)
__device__ __forceinline__ int k0_side_len(uint32_t h) {
    const bool mono = ((h >> 6) & 3) == 3;
    return (((h >> 19) & 3) == 3) ? (mono ? 17 : 32) : (mono ? 9 : 17);
}

// valid Layer III header of any MPEG version (so LSF streams are recognised and reported, like the oracle's parse_hdr)
__device__ __forceinline__ bool k0_hdr_valid(uint32_t h) {
    if ((h & 0xFFE00000u) != 0xFFE00000u) return false;
    const int ver = (h >> 19) & 3, layer = (h >> 17) & 3, br = (h >> 12) & 15, sr = (h >> 10) & 3;
    return ver != 1 && layer == 1 && br != 0 && br != 15 && sr != 3;
}
// same version, sample rate and channel count (mono vs not)
__device__ __forceinline__ bool k0_compatible(uint32_t a, uint32_t ref) {
    return ((a ^ ref) & 0x00180C00u) == 0 && ((((a >> 6) & 3) == 3) == (((ref >> 6) & 3) == 3));
}

#define K0_SEG (64u * 1024u)            // bytes per speculative segment
#define K0_SEG_RECS (K0_SEG / 24 + 18)  // staged frames per segment (smallest Layer III frame: 24 bytes -- MPEG-2, 8 kbps at 24 kHz;
                                        // with the MPEG-1 minimum of 96 bytes, low-bitrate LSF streams overflowed the staging)

struct TrackSeg { uint32_t seg_base, n_seg; };     // a track's segments: [seg_base, seg_base + n_seg) of the call's segment arrays
struct SegDesc { uint32_t track, seg; };           // segment -> (track of the launch's group, index within the track)
struct SegRec { uint32_t fpos, hdr, info, pad; };  // staged frame: info = main_data_begin | payload bytes << 9 | indexed << 20
struct SegOut {
    uint32_t first_pos;    // first accepted frame (if FOUND)
    uint32_t end_pos;      // where the walk stopped: first frame starting at or beyond the segment's end, or the EOF position
    uint32_t n_rec;        // staged frames (every accepted frame except a skipped Xing/Info frame)
    uint32_t n_indexed;    // of those, frames that become FrameRecs
    uint32_t m_local;      // main-data bytes of the staged frames
    uint32_t ref;          // header of the first accepted frame
    uint32_t flags;        // K0F_*
    uint32_t n_total;      // accepted frames including a skipped info frame
};
struct SegFin { uint32_t m_base, f_base, n_rec, pad; };   // per segment, from the merge: where its frames go
#define K0F_FOUND 1u
#define K0F_INSYNC 2u      // the walk was in sync when it stopped
#define K0F_STOP 4u        // truncated final frame or unsupported stream: nothing after this segment counts
#define K0F_LSF 8u
#define K0F_INFO 16u       // skipped a leading Xing/Info/VBRI frame

__device__ __forceinline__ uint64_t k0_id3_skip(const uint8_t* __restrict__ d, uint64_t len) {
    // ID3v2 skip (syncsafe size, optional footer), like symphonia's probe / reference src/lib.rs:343-354
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3' && d[3] != 0xFF && d[4] != 0xFF && !((d[6] | d[7] | d[8] | d[9]) & 0x80)) {
        uint64_t sz = ((uint64_t)d[6] << 21) | ((uint64_t)d[7] << 14) | ((uint64_t)d[8] << 7) | d[9];
        sz += 10;
        if (d[5] & 0x10) sz += 10;
        return sz <= len ? sz : len;
    }
    return 0;
}

// The serial frame walk over [pos, bound): the whole warp executes it (uniform control flow).  `m_known`: the main-data
// offset M0 at `pos` is exact, so reservoir underflow (main_data_begin > bytes available) is decided here; otherwise it is
// assumed not to happen (true once >= 511 bytes precede, which the merge checks).  `first`: the next accepted frame is the
// track's first (Xing/Info/VBRI check).
__device__ SegOut k0_walk(const uint8_t* __restrict__ src, uint64_t len, uint8_t* ck, const uint16_t (*s_flen)[16], uint64_t pos,
                          uint64_t bound, bool in_sync, bool have_ref, uint32_t ref, bool first, bool m_known, uint32_t M0,
                          uint32_t flags, SegRec* __restrict__ recs, int lane) {
    SegOut o;
    o.first_pos = 0; o.n_rec = 0; o.n_indexed = 0; o.m_local = 0; o.ref = ref; o.flags = 0; o.n_total = 0;
    uint32_t M = 0;
    uint64_t chunk_start = ~0ull, chunk_end = 0;
    while (pos + 4 <= len && pos < bound) {
        // keep [pos, pos + 1448) inside the window (largest frame 1441 B + the next header), or up to EOF
        const uint64_t avail_end0 = chunk_end < len ? chunk_end : len;
        if (pos < chunk_start || pos >= chunk_end || (pos + 1448 > avail_end0 && avail_end0 < len)) {
            chunk_start = pos & ~15ull;
            chunk_end = chunk_start + K0_CHUNK;
            __syncwarp();
            for (int i = lane * 16; i < K0_CHUNK; i += 32 * 16) {
                uint4 v = make_uint4(0, 0, 0, 0);
                const uint64_t g = chunk_start + i;
                if (g + 16 <= len) v = *reinterpret_cast<const uint4*>(src + g);
                else if (g < len) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (int b = 0; b < 16; b++) if (g + b < len) w[b >> 2] |= (uint32_t)src[g + b] << (8 * (b & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                *reinterpret_cast<uint4*>(ck + i) = v;
            }
            __syncwarp();
        }
        const uint8_t* p = ck + (pos - chunk_start);
        const uint32_t h = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        bool ok = k0_hdr_valid(h) && (!have_ref || k0_compatible(h, ref));
        int flen = 0;
        if (ok) {
            flen = (((h >> 19) & 3) == 3) ? (int)s_flen[(h >> 10) & 3][(h >> 12) & 15] + (int)((h >> 9) & 1) : k0_frame_len(h);
            if (pos + flen > len) { o.flags |= K0F_STOP; break; }  // truncated final frame = EOF (src/replaygain.rs:884-888)
        }
        if (ok && !in_sync) {  // (re)sync: the next header must confirm
            const uint64_t nx = pos + flen;
            if (nx + 4 <= len) {
                const uint8_t* q = ck + (nx - chunk_start);
                const uint32_t h2 = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
                if (!(k0_hdr_valid(h2) && k0_compatible(h2, h))) ok = false;
            }
        }
        if (!ok) { in_sync = false; pos++; continue; }
        in_sync = true;
        if (!(o.flags & K0F_FOUND)) { o.flags |= K0F_FOUND; o.first_pos = (uint32_t)pos; if (!have_ref) o.ref = h; }
        if (!have_ref) { ref = h; have_ref = true; }
        const int mono = ((h >> 6) & 3) == 3;
        const bool lsf = ((h >> 19) & 3) != 3;
        const int side_len = k0_side_len(h);
        const int so = 4 + ((h & 0x10000u) ? 0 : 2);
        const uint32_t fpos = (uint32_t)pos;
        pos += flen;
        o.n_total++;
        if (first) {
            first = false;
            if (!(flags & 0x2u)) {
                const uint8_t* x = p + so + side_len;
                bool tag = (so + side_len + 4 <= flen) &&
                           ((x[0] == 'X' && x[1] == 'i' && x[2] == 'n' && x[3] == 'g') || (x[0] == 'I' && x[1] == 'n' && x[2] == 'f' && x[3] == 'o'));
                if (!tag && flen >= 40) tag = p[36] == 'V' && p[37] == 'B' && p[38] == 'R' && p[39] == 'I';
                if (tag) { o.flags |= K0F_INFO; continue; }
            }
        }
        // side info: main_data_begin + validity (big_values <= 288, block_type != 0 when window switching); lane = granule-channel
        // MPEG-1: 9-bit main_data_begin, 59 bits per granule-channel from bit 18 / 20, two granules; MPEG-2 / 2.5: 8-bit
        // main_data_begin, 63 bits per granule-channel from bit 9 / 10, one granule (9-bit scalefac_compress, no preflag bit)
        const uint8_t* s = p + so;
        const int mdb = lsf ? (int)s[0] : (((int)s[0] << 1) | (s[1] >> 7));
        bool bad = false;
        if (lane < (mono ? 1 : 2) * (lsf ? 1 : 2)) {
            const int b0 = lsf ? (mono ? 9 : 10) + 63 * lane : (mono ? 18 : 20) + 59 * lane;
            const int bb = b0 + 12;  // big_values: 9 bits
            const uint32_t w = ((uint32_t)s[bb >> 3] << 16) | ((uint32_t)s[(bb >> 3) + 1] << 8) | s[(bb >> 3) + 2];
            const int bv = (w >> (24 - 9 - (bb & 7))) & 0x1FF;
            const int wb = b0 + (lsf ? 38 : 33);  // window_switching, block_type(2)
            const uint32_t w2 = ((uint32_t)s[wb >> 3] << 8) | s[(wb >> 3) + 1];
            const int ws_bt = (w2 >> (16 - 3 - (wb & 7))) & 7;
            bad = bv > 288 || ((ws_bt & 4) && !(ws_bt & 3));
        }
        const bool valid = !__any_sync(0xffffffffu, bad);
        const int mdo_off = so + side_len;
        const int cap = flen > mdo_off ? flen - mdo_off : 0;
        const bool start_ok = !m_known || (uint32_t)mdb <= M0 + M;
        const bool indexed = valid && start_ok;
        if (lane == 0 && o.n_rec < K0_SEG_RECS) {
            SegRec r;
            r.fpos = fpos; r.hdr = h; r.info = (uint32_t)mdb | ((uint32_t)cap << 9) | (indexed ? 1u << 20 : 0u); r.pad = 0;
            recs[o.n_rec] = r;
        }
        o.n_rec++;
        o.n_indexed += indexed ? 1 : 0;
        M += cap;
    }
    o.m_local = M;
    o.end_pos = (uint32_t)pos;
    if (in_sync) o.flags |= K0F_INSYNC;
    return o;
}

__device__ __forceinline__ void k0_fill_flen(uint16_t (*s_flen)[16]) {
    if (threadIdx.x < 48) {
        const int sr = threadIdx.x >> 4, br = threadIdx.x & 15;
        s_flen[sr][br] = (br == 0 || br == 15) ? 0 : (uint16_t)k0_frame_len(0xFFFA0000u | (br << 12) | (sr << 10));
    }
    __syncthreads();
}

// byte range of segment `seg` of a track whose audio starts at a0
__device__ __forceinline__ void k0_seg_range(uint32_t seg, uint32_t n_seg, uint64_t len, uint64_t a0, uint64_t& b, uint64_t& e) {
    b = (uint64_t)seg * K0_SEG;
    e = (seg + 1 == n_seg) ? len : (uint64_t)(seg + 1) * K0_SEG;
    if (b < a0) b = a0;
}

__global__ void __launch_bounds__(32 * K0_WARPS, 10)
k_index_scan(const uint8_t* __restrict__ corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg,
             const SegDesc* __restrict__ segs, uint32_t seg0, uint32_t n_segs, SegRec* __restrict__ stage,
             SegOut* __restrict__ segout, uint32_t flags) {
    __shared__ __align__(16) uint8_t s_chunk[K0_WARPS][K0_CHUNK + 16];
    __shared__ uint16_t s_flen[3][16];   // MPEG-1 Layer III frame length without padding, by sample-rate and bitrate index
    k0_fill_flen(s_flen);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t si = blockIdx.x * K0_WARPS + warp;
    if (si >= n_segs) return;
    const SegDesc sd = segs[seg0 + si];
    const uint8_t* src = corpus + tin[sd.track].off;   // 16-byte aligned by construction
    const uint64_t len = tin[sd.track].len;
    const uint64_t a0 = k0_id3_skip(src, len);
    uint64_t b, e;
    k0_seg_range(sd.seg, tseg[sd.track].n_seg, len, a0, b, e);
    SegOut o;
    if (b >= e) { o.first_pos = 0; o.end_pos = (uint32_t)b; o.n_rec = 0; o.n_indexed = 0; o.m_local = 0; o.ref = 0; o.flags = 0; o.n_total = 0; }
    else {
        const bool head = (uint64_t)sd.seg * K0_SEG <= a0;   // the segment the audio starts in: exact reservoir accounting from M = 0
        o = k0_walk(src, len, s_chunk[warp], s_flen, b, e, false, false, 0, head, head, 0, flags,
                    stage + (size_t)(seg0 + si) * K0_SEG_RECS, lane);
    }
    if (lane == 0) segout[seg0 + si] = o;
}

__global__ void __launch_bounds__(32 * K0_WARPS)
k_index_merge(const uint8_t* __restrict__ corpus, const TrackIn* __restrict__ tin, const TrackSeg* __restrict__ tseg,
              const uint64_t* __restrict__ md_base, SegRec* __restrict__ stage, SegOut* __restrict__ segout,
              SegFin* __restrict__ segfin, uint8_t* __restrict__ md, TrackMeta* __restrict__ meta, uint32_t n_tracks, uint32_t flags) {
    __shared__ __align__(16) uint8_t s_chunk[K0_WARPS][K0_CHUNK + 16];
    __shared__ uint16_t s_flen[3][16];
    k0_fill_flen(s_flen);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * K0_WARPS + warp;
    if (t >= n_tracks) return;
    const uint8_t* src = corpus + tin[t].off;
    const uint64_t len = tin[t].len;
    const uint64_t a0 = k0_id3_skip(src, len);
    const TrackSeg ts = tseg[t];
    bool have_ref = false, in_sync = false, stopped = false;
    uint32_t ref = 0, M = 0, nf = 0, n_total = 0, n_rec_all = 0, skipped = 0, n_rewalks = 0;
    uint64_t P = a0;
    int status = -5;   // RG_ERR_PROBE until a frame is found
    for (uint32_t s = 0; s < ts.n_seg; s++) {
        const uint32_t gi = ts.seg_base + s;
        uint64_t b, e;
        k0_seg_range(s, ts.n_seg, len, a0, b, e);
        SegFin fin;
        fin.m_base = M; fin.f_base = nf; fin.n_rec = 0; fin.pad = 0;
        if (!stopped && b < e && !(in_sync && P >= e)) {
            SegOut o = segout[gi];
            const bool head = (uint64_t)s * K0_SEG <= a0;
            bool redo = false;
            if (!head) {
                if (in_sync && !((o.flags & K0F_FOUND) && o.first_pos == (uint32_t)P)) redo = true;          // false sync (or none) in the scan
                if ((o.flags & K0F_FOUND) && have_ref && !k0_compatible(o.ref, ref)) redo = true;            // scanned against the wrong stream type
                if ((o.flags & K0F_FOUND) && n_total == 0) redo = true;                                      // it holds the track's first frame
                if (o.n_rec > 0 && M < 511) redo = true;                                                     // reservoir may still underflow
            }
            if (redo) n_rewalks++;
            if (redo)
                o = k0_walk(src, len, s_chunk[warp], s_flen, in_sync ? P : b, e, in_sync, have_ref, ref, n_total == 0, true, M, flags,
                            stage + (size_t)gi * K0_SEG_RECS, lane);
            if ((o.flags & K0F_FOUND) && !have_ref) { ref = o.ref; have_ref = true; status = 0; }
            if (o.flags & K0F_LSF) { status = -10; stopped = true; }
            else {
                fin.n_rec = o.n_rec < K0_SEG_RECS ? o.n_rec : K0_SEG_RECS;
                M += o.m_local; nf += o.n_indexed; n_total += o.n_total; n_rec_all += o.n_rec;
                if (o.flags & K0F_INFO) skipped = 1;
                in_sync = (o.flags & K0F_INSYNC) != 0;
                P = o.end_pos;
                if (o.flags & K0F_STOP) stopped = true;
            }
        }
        if (lane == 0) segfin[gi] = fin;
    }
    // zero padding after the stream so bit readers may over-read
    uint8_t* mdo = md + md_base[t];
    for (int bb = lane; bb < 16; bb += 32) mdo[M + bb] = 0;
    if (lane == 0) {
        TrackMeta m;
        m.status = status;
        m.sr_idx = have_ref ? (((ref >> 10) & 3) | (((ref >> 19) & 3) << 4)) : 0;   // sample-rate index | MPEG version << 4
        m.channels = have_ref ? ((((ref >> 6) & 3) == 3) ? 1 : 2) : 0;
        m.n_frames = nf; m.md_bytes = M; m.n_dropped = n_rec_all - nf; m.skipped_info = skipped; m.n_total = n_total; m.n_rewalks = n_rewalks;
        meta[t] = m;
    }
}

__global__ void __launch_bounds__(32 * K0_WARPS)
k_index_finalize(const uint8_t* __restrict__ corpus, const TrackIn* __restrict__ tin, const SegDesc* __restrict__ segs, uint32_t seg0,
                 uint32_t n_segs, const uint64_t* __restrict__ frame_base, const uint64_t* __restrict__ md_base,
                 const SegRec* __restrict__ stage, const SegFin* __restrict__ segfin, FrameRec* __restrict__ frames,
                 uint8_t* __restrict__ md) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t si = blockIdx.x * K0_WARPS + warp;
    if (si >= n_segs) return;
    const SegDesc sd = segs[seg0 + si];
    const SegFin fin = segfin[seg0 + si];
    if (!fin.n_rec) return;
    const SegRec* recs = stage + (size_t)(seg0 + si) * K0_SEG_RECS;
    FrameRec* fr = frames + frame_base[sd.track];
    uint32_t M = fin.m_base, nf = fin.f_base;
    for (uint32_t r0 = 0; r0 < fin.n_rec; r0 += 32) {
        const uint32_t r = r0 + lane;
        SegRec rec;
        rec.fpos = 0; rec.hdr = 0; rec.info = 0; rec.pad = 0;
        if (r < fin.n_rec) rec = recs[r];
        const uint32_t mdb = rec.info & 0x1FF, cap = (rec.info >> 9) & 0x7FF, indexed = (rec.info >> 20) & 1;
        uint32_t pm = cap, pf = indexed;   // inclusive warp prefix sums
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t am = __shfl_up_sync(0xffffffffu, pm, o), af = __shfl_up_sync(0xffffffffu, pf, o);
            if (lane >= o) { pm += am; pf += af; }
        }
        const uint32_t m_before = M + pm - cap;
        if (r < fin.n_rec && indexed) {
            FrameRec f;
            f.file_off = rec.fpos; f.md_start = m_before - mdb; f.hdr = rec.hdr; f.md_end = m_before + cap;
            fr[nf + pf - 1] = f;
        }
        // frames dropped from the index still feed the reservoir: copy their main data here (k_md_gather only sees FrameRecs)
        uint32_t dropped = __ballot_sync(0xffffffffu, r < fin.n_rec && !indexed && cap > 0);
        while (dropped) {
            const int l = __ffs(dropped) - 1;
            dropped &= dropped - 1;
            const uint32_t fpos = __shfl_sync(0xffffffffu, rec.fpos, l), hdr = __shfl_sync(0xffffffffu, rec.hdr, l);
            const uint32_t c = __shfl_sync(0xffffffffu, cap, l), mb = __shfl_sync(0xffffffffu, m_before, l);
            const int off = 4 + ((hdr & 0x10000u) ? 0 : 2) + k0_side_len(hdr);
            const uint8_t* sp = corpus + tin[sd.track].off + fpos + off;
            uint8_t* dp = md + md_base[sd.track] + mb;
            for (uint32_t bb = lane; bb < c; bb += 32) dp[bb] = sp[bb];
        }
        M += __shfl_sync(0xffffffffu, pm, 31);
        nf += __shfl_sync(0xffffffffu, pf, 31);
    }
}

// K0b: gathers the main data of every indexed frame into the track's contiguous stream (the bit reservoir unrolled).
// One CTA per run of frames (the K2 run descriptors are reused: granules g0 .. g0+ng of a track), one warp per frame.
#define K0B_THREADS 128
__global__ void __launch_bounds__(K0B_THREADS)
k_md_gather(const uint8_t* __restrict__ corpus, uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
            const DecTrack* __restrict__ tracks, const RunDesc* __restrict__ runs) {
    __shared__ uint16_t s_flen[3][16];   // MPEG-1 frame lengths by table (k0_frame_len divides: a fifth of this kernel's instructions)
    k0_fill_flen(s_flen);
    const RunDesc run = runs[blockIdx.x];
    const DecTrack tr = tracks[run.track];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t f0 = run.g0 / tr.gpf, f1 = (run.g0 + run.ng + tr.gpf - 1) / tr.gpf;
    for (uint32_t f = f0 + warp; f < f1; f += K0B_THREADS / 32) {
        const FrameRec rec = frames[tr.frame_base + f];
        const uint32_t h = rec.hdr;
        const int off = 4 + ((h & 0x10000u) ? 0 : 2) + k0_side_len(h);
        const int flen = (((h >> 19) & 3) == 3) ? (int)s_flen[(h >> 10) & 3][(h >> 12) & 15] + (int)((h >> 9) & 1) : k0_frame_len(h);
        const int cap = flen > off ? flen - off : 0;
        const uint8_t* src = corpus + tr.corpus_off + rec.file_off + off;
        uint8_t* dst = md + tr.md_off + rec.md_end - cap;
        // head bytes up to the first 4-byte boundary of dst, aligned words, tail bytes; the source is read bytewise
        // through L1 (each sector is reused by neighbouring lanes)
        const int head = min(cap, (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
        if (lane < head) dst[lane] = __ldg(src + lane);
        const int nw = (cap - head) >> 2;
        const uint8_t* s4 = src + head;
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
        // aligned destination words from two aligned source words each (funnel shift by the source's misalignment)
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(s4) & 3);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s4 - mis);
        for (int w = lane; w < nw; w += 32) {
            const uint32_t lo = __ldg(sw + w), hi = mis ? __ldg(sw + w + 1) : 0u;
            d4[w] = __funnelshift_r(lo, hi, 8 * mis);
        }
        const int tail0 = head + 4 * nw;
        if (tail0 + lane < cap) dst[tail0 + lane] = __ldg(src + tail0 + lane);
    }
}
