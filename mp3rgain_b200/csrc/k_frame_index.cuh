// k_frame_index.cuh -- K0: per-track frame walk + main-data compaction.
//
// One warp per track.  The warp streams the file image through an 8 KiB shared-memory window (coalesced 16-byte loads),
// walks the frame headers in it with warp-uniform scalar code (every lane reads the same bytes: smem broadcasts), and,
// frame by frame, all 32 lanes copy the frame's main-data bytes into the track's contiguous main-data stream.
// Replaces symphonia's MpaReader + bit reservoir (reference call sites src/replaygain.rs:815-822, 882); the header
// arithmetic matches the reference's own frame walker (src/lib.rs:169-252: frame_size = 144*kbps*1000/sr + pad).
#pragma once

#define K0_CHUNK 8192
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

__global__ void __launch_bounds__(32 * K0_WARPS)
k_frame_index(const uint8_t* __restrict__ corpus, const TrackIn* __restrict__ tin, const uint64_t* __restrict__ frame_base,
              const uint64_t* __restrict__ md_base, FrameRec* __restrict__ frames, uint8_t* __restrict__ md,
              TrackMeta* __restrict__ meta, uint32_t n_tracks, uint32_t flags) {
    __shared__ __align__(16) uint8_t s_chunk[K0_WARPS][K0_CHUNK + 16];
    __shared__ uint16_t s_flen[3][16];   // MPEG-1 Layer III frame length without padding, by sample-rate and bitrate index
    if (threadIdx.x < 48) {
        const int sr = threadIdx.x >> 4, br = threadIdx.x & 15;
        s_flen[sr][br] = (br == 0 || br == 15) ? 0 : (uint16_t)k0_frame_len(0xFFFA0000u | (br << 12) | (sr << 10));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * K0_WARPS + warp;
    if (t >= n_tracks) return;
    uint8_t* ck = s_chunk[warp];
    const uint8_t* src = corpus + tin[t].off;   // 16-byte aligned by construction
    const uint64_t len = tin[t].len;
    FrameRec* fr = frames + frame_base[t];
    uint8_t* mdo = md + md_base[t];

    // ID3v2 skip (syncsafe size, optional footer), like symphonia's probe / reference src/lib.rs:343-354
    uint64_t pos = 0;
    if (len >= 10) {
        const uint8_t* d = src;
        if (d[0] == 'I' && d[1] == 'D' && d[2] == '3' && d[3] != 0xFF && d[4] != 0xFF && !((d[6] | d[7] | d[8] | d[9]) & 0x80)) {
            uint64_t sz = ((uint64_t)d[6] << 21) | ((uint64_t)d[7] << 14) | ((uint64_t)d[8] << 7) | d[9];
            sz += 10;
            if (d[5] & 0x10) sz += 10;
            pos = sz <= len ? sz : len;
        }
    }

    uint32_t ref = 0;
    bool have_ref = false, in_sync = false, first = true;
    int status = -5;  // RG_ERR_PROBE until a frame is found
    uint32_t n_frames = 0, n_dropped = 0, skipped_info = 0, n_total = 0;
    uint32_t M = 0;   // main-data bytes accounted for so far
    uint64_t chunk_start = ~0ull, chunk_end = 0;

    // The walk itself only touches the 4 header bytes and the side info of each frame; the main data of every indexed frame
    // is gathered into the track's contiguous stream afterwards by k_md_gather (fully parallel).  Only frames that are
    // dropped from the index (they still feed the bit reservoir) are copied here, inline.
    while (pos + 4 <= len) {
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
            if (pos + flen > len) break;  // truncated final frame = EOF (src/replaygain.rs:884-888)
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
        if (!have_ref) {
            ref = h; have_ref = true; status = 0;
            if (((h >> 19) & 3) != 3) { status = -10; break; }  // MPEG-2/2.5: recognised, not decoded by this engine
        }
        const int mono = ((h >> 6) & 3) == 3;
        const int side_len = mono ? 17 : 32;
        const int so = 4 + ((h & 0x10000u) ? 0 : 2);
        const uint32_t fpos = (uint32_t)pos;
        pos += flen;
        n_total++;
        if (first) {
            first = false;
            if (!(flags & 0x2u)) {
                const uint8_t* x = p + so + side_len;
                bool tag = (so + side_len + 4 <= flen) &&
                           ((x[0] == 'X' && x[1] == 'i' && x[2] == 'n' && x[3] == 'g') || (x[0] == 'I' && x[1] == 'n' && x[2] == 'f' && x[3] == 'o'));
                if (!tag && flen >= 40) tag = p[36] == 'V' && p[37] == 'B' && p[38] == 'R' && p[39] == 'I';
                if (tag) { skipped_info = 1; continue; }
            }
        }
        // side info: main_data_begin + validity (big_values <= 288, block_type != 0 when window switching); lane = granule-channel
        const uint8_t* s = p + so;
        const int mdb = ((int)s[0] << 1) | (s[1] >> 7);
        bool bad = false;
        if (lane < (mono ? 2 : 4)) {
            const int b0 = (mono ? 18 : 20) + 59 * lane;
            const int bb = b0 + 12;  // big_values: 9 bits
            const uint32_t w = ((uint32_t)s[bb >> 3] << 16) | ((uint32_t)s[(bb >> 3) + 1] << 8) | s[(bb >> 3) + 2];
            const int bv = (w >> (24 - 9 - (bb & 7))) & 0x1FF;
            const int wb = b0 + 33;  // window_switching, block_type(2)
            const uint32_t w2 = ((uint32_t)s[wb >> 3] << 8) | s[(wb >> 3) + 1];
            const int ws_bt = (w2 >> (16 - 3 - (wb & 7))) & 7;
            bad = bv > 288 || ((ws_bt & 4) && !(ws_bt & 3));
        }
        const bool valid = !__any_sync(0xffffffffu, bad);
        const int mdo_off = so + side_len;
        const int cap = flen > mdo_off ? flen - mdo_off : 0;
        const bool start_ok = (uint32_t)mdb <= M;
        const uint32_t start = M - (start_ok ? mdb : 0);
        if (valid && start_ok) {
            M += cap;
            if (lane == 0) {
                FrameRec r;
                r.file_off = fpos; r.md_start = start; r.hdr = h; r.md_end = M;
                fr[n_frames] = r;
            }
            n_frames++;
        } else {
            for (int b = lane; b < cap; b += 32) mdo[M + b] = p[mdo_off + b];  // dropped frame: its main data still enters the reservoir
            M += cap;
            n_dropped++;
        }
    }
    // zero padding after the stream so bit readers may over-read
    for (int b = lane; b < 16; b += 32) mdo[M + b] = 0;
    if (lane == 0) {
        TrackMeta m;
        m.status = status;
        m.sr_idx = have_ref ? ((ref >> 10) & 3) : 0;
        m.channels = have_ref ? ((((ref >> 6) & 3) == 3) ? 1 : 2) : 0;
        if (have_ref && ((ref >> 19) & 3) != 3) {  // LSF: report the real rate index via n_total for the host (unsupported anyway)
            m.sr_idx = ((ref >> 10) & 3) | (((ref >> 19) & 3) << 4);
        }
        m.n_frames = n_frames; m.md_bytes = M; m.n_dropped = n_dropped; m.skipped_info = skipped_info; m.n_total = n_total;
        meta[t] = m;
    }
}

// K0b: gathers the main data of every indexed frame into the track's contiguous stream (the bit reservoir unrolled).
// One CTA per run of frames (the K2 run descriptors are reused: granules g0 .. g0+ng of a track), one warp per frame.
#define K0B_THREADS 128
__global__ void __launch_bounds__(K0B_THREADS)
k_md_gather(const uint8_t* __restrict__ corpus, uint8_t* __restrict__ md, const FrameRec* __restrict__ frames,
            const DecTrack* __restrict__ tracks, const RunDesc* __restrict__ runs) {
    const RunDesc run = runs[blockIdx.x];
    const DecTrack tr = tracks[run.track];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t f0 = run.g0 >> 1, f1 = (run.g0 + run.ng + 1) >> 1;
    for (uint32_t f = f0 + warp; f < f1; f += K0B_THREADS / 32) {
        const FrameRec rec = frames[tr.frame_base + f];
        const uint32_t h = rec.hdr;
        const int off = 4 + ((h & 0x10000u) ? 0 : 2) + ((((h >> 6) & 3) == 3) ? 17 : 32);
        const int flen = k0_frame_len(h);
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
        for (int w = lane; w < nw; w += 32) {
            const uint8_t* q = s4 + 4 * w;
            d4[w] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) | ((uint32_t)__ldg(q + 3) << 24);
        }
        const int tail0 = head + 4 * nw;
        if (tail0 + lane < cap) dst[tail0 + lane] = __ldg(src + tail0 + lane);
    }
}
