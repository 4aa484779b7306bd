// rg_engine_decode.inl -- entry points that run the MP3 decode kernels (stage A) and feed stage B.
// Included at the end of rg_engine.cu.

#define RUN_GRANULES 62   // granules per K2 CTA (even: runs start on granule 0 of a frame; +2 replayed = K2_MAX_G)

static const uint32_t MPEG1_RATES[3] = {44100, 48000, 32000};

// mp4meta::is_mp4_file (src/mp4meta.rs:872-889): ftyp box with a known brand => AudioFileType::Aac
static bool sniff_mp4(const uint8_t* d, uint64_t len) {
    if (len < 12) return false;
    const uint32_t size = ((uint32_t)d[0] << 24) | ((uint32_t)d[1] << 16) | ((uint32_t)d[2] << 8) | d[3];
    if (memcmp(d + 4, "ftyp", 4) != 0 || size < 12) return false;
    static const char* brands[] = {"M4A ", "M4B ", "M4P ", "M4V ", "mp41", "mp42", "isom", "iso2"};
    for (const char* b : brands) if (!memcmp(d + 8, b, 4)) return true;
    return false;
}

extern "C" int rg_corpus_upload(rg_engine* e, const rg_input* in, size_t n, rg_corpus** out) {
    if (!e || !out || (n && !in)) return RG_ERR_INVALID_ARG;
    *out = nullptr;
    CK(cudaSetDevice(e->device));
    rg_corpus* c = new rg_corpus();
    c->off.resize(n); c->len.resize(n); c->track_index.resize(n); c->is_mp4.resize(n);
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (in[i].len && !in[i].data) { delete c; return RG_ERR_INVALID_ARG; }
        c->off[i] = total;
        c->len[i] = in[i].len;
        c->track_index[i] = in[i].track_index;
        c->is_mp4[i] = sniff_mp4(in[i].data, in[i].len) ? 1 : 0;
        total += (in[i].len + 16 + 15) & ~15ull;   // 16-byte aligned start, >= 16 bytes of slack
    }
    c->total = total;
    cudaError_t err = c->d_bytes.ensure(total + 64);
    if (err != cudaSuccess) { delete c; return fail_cuda(err, "cudaMalloc(corpus)", __LINE__); }
    for (size_t i = 0; i < n; i++) {
        if (!in[i].len) continue;
        err = cudaMemcpyAsync(c->d_bytes.as<uint8_t>() + c->off[i], in[i].data, in[i].len, cudaMemcpyHostToDevice, e->stream);
        if (err != cudaSuccess) { c->d_bytes.release(); delete c; return fail_cuda(err, "cudaMemcpyAsync(corpus)", __LINE__); }
        e->stats.h2d_bytes += in[i].len;
    }
    err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) { c->d_bytes.release(); delete c; return fail_cuda(err, "cudaStreamSynchronize", __LINE__); }
    *out = c;
    return RG_OK;
}

extern "C" void rg_corpus_free(rg_engine* e, rg_corpus* c) {
    if (!c) return;
    if (e) cudaSetDevice(e->device);
    c->d_bytes.release();
    delete c;
}

extern "C" uint64_t rg_corpus_bytes(const rg_corpus* c) {
    uint64_t s = 0;
    if (c) for (uint64_t l : c->len) s += l;
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
struct DecodeOut {          // what a decode pass leaves behind for its caller
    std::vector<TrackMeta> meta;
    std::vector<int> status;            // rg_status per track after host-side checks
    std::vector<uint32_t> sample_rate;
};

// K0 over the whole corpus.  Leaves FrameRec / main-data / meta on the device (e->dec_*), meta on the host.
static int run_frame_index(rg_engine* e, rg_corpus* c, DecodeOut& d, std::vector<uint64_t>& frame_base, std::vector<uint64_t>& md_base) {
    const size_t n = c->len.size();
    d.meta.assign(n, TrackMeta{});
    d.status.assign(n, RG_OK);
    d.sample_rate.assign(n, 0);
    frame_base.resize(n + 1);
    md_base.resize(n + 1);
    std::vector<TrackIn> tin(n);
    uint64_t fb = 0, mb = 0;
    for (size_t i = 0; i < n; i++) {
        tin[i].off = c->off[i]; tin[i].len = c->len[i];
        frame_base[i] = fb; md_base[i] = mb;
        fb += c->len[i] / 96 + 2;                      // smallest MPEG-1 Layer III frame is 96 bytes
        mb += (c->len[i] + 32 + 15) & ~15ull;
    }
    frame_base[n] = fb; md_base[n] = mb;
    if (!n) return RG_OK;
    CK(e->d_tin.ensure(n * sizeof(TrackIn)));
    CK(e->d_fbase.ensure((n + 1) * sizeof(uint64_t) * 2));
    CK(e->d_frames.ensure(fb * sizeof(FrameRec)));
    CK(e->d_md.ensure(mb + 64));
    CK(e->d_meta.ensure(n * sizeof(TrackMeta)));
    CK(cudaMemcpyAsync(e->d_tin.p, tin.data(), n * sizeof(TrackIn), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_fbase.p, frame_base.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_fbase.as<uint64_t>() + (n + 1), md_base.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
    e->stats.h2d_bytes += n * sizeof(TrackIn) + 2 * (n + 1) * sizeof(uint64_t);
    CK(cudaEventRecord(e->ev[0], e->stream));
    k_frame_index<<<(unsigned)((n + K0_WARPS - 1) / K0_WARPS), 32 * K0_WARPS, 0, e->stream>>>(
        c->d_bytes.as<uint8_t>(), e->d_tin.as<TrackIn>(), e->d_fbase.as<uint64_t>(), e->d_fbase.as<uint64_t>() + (n + 1),
        e->d_frames.as<FrameRec>(), e->d_md.as<uint8_t>(), e->d_meta.as<TrackMeta>(), (uint32_t)n, e->flags);
    e->stats.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev[1], e->stream));
    CK(cudaMemcpyAsync(d.meta.data(), e->d_meta.p, n * sizeof(TrackMeta), cudaMemcpyDeviceToHost, e->stream));
    e->stats.d2h_bytes += n * sizeof(TrackMeta);
    CK(cudaStreamSynchronize(e->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]) == cudaSuccess) e->stats.last_ms_index += ms;
    for (size_t i = 0; i < n; i++) {
        const TrackMeta& m = d.meta[i];
        int st = m.status;
        if (c->is_mp4[i]) st = RG_ERR_UNSUPPORTED_FORMAT;          // AAC/M4A: out of scope for this engine
        else if (st == RG_OK) {
            d.sample_rate[i] = MPEG1_RATES[m.sr_idx & 3];
            // MP3 carries exactly one audio track (src/replaygain.rs:838-851)
            if (c->track_index[i] > 0) st = RG_ERR_TRACK_INDEX;
            else if (c->track_index[i] < -1) st = RG_ERR_INVALID_ARG;
        } else if (st == RG_ERR_UNSUPPORTED_FORMAT) {
            const uint32_t ver = (m.sr_idx >> 4) & 3;
            d.sample_rate[i] = MPEG1_RATES[m.sr_idx & 3] >> (ver == 2 ? 1 : 2);
        }
        d.status[i] = st;
    }
    return RG_OK;
}

// Decode tracks [t0,t1) of the corpus (K1+K2) into e->d_pcm; fills `pt` with the PCM views.
static int run_decode_batch(rg_engine* e, rg_corpus* c, const DecodeOut& d, const std::vector<uint64_t>& frame_base,
                            const std::vector<uint64_t>& md_base, size_t t0, size_t t1, std::vector<PcmTrack>& pt) {
    const size_t nb = t1 - t0;
    pt.assign(nb, PcmTrack{});
    std::vector<DecTrack> dt(nb);
    std::vector<HuffBlock> hb;
    std::vector<RunDesc> runs;
    uint64_t gc_total = 0, pcm_total = 0;
    for (size_t k = 0; k < nb; k++) {
        const size_t i = t0 + k;
        const TrackMeta& m = d.meta[i];
        DecTrack& t = dt[k];
        memset(&t, 0, sizeof t);
        PcmTrack& p = pt[k];
        p.status = d.status[i];
        p.sample_rate = d.sample_rate[i];
        p.rate_idx = rate_index(p.sample_rate);
        p.nch = (int)m.channels;
        if (p.status != RG_OK) continue;
        t.corpus_off = c->off[i];
        t.md_off = md_base[i];
        t.frame_base = frame_base[i];
        t.gc_base = gc_total;
        t.n_frames = m.n_frames; t.channels = m.channels; t.sr_idx = m.sr_idx & 3; t.md_bytes = m.md_bytes;
        const uint64_t plane = (uint64_t)m.n_frames * 1152;
        t.pcm_off = pcm_total; t.pcm_plane = plane;
        const uint32_t n_gc = m.n_frames * 2 * m.channels;
        for (uint32_t g0 = 0; g0 < n_gc; g0 += K1_THREADS) hb.push_back(HuffBlock{(uint32_t)k, g0});
        for (uint32_t g0 = 0; g0 < m.n_frames * 2; g0 += RUN_GRANULES)
            runs.push_back(RunDesc{(uint32_t)k, g0, std::min<uint32_t>(RUN_GRANULES, m.n_frames * 2 - g0), 0});
        gc_total += n_gc;
        pcm_total += plane * m.channels;
        p.frames = plane;
    }
    CK(e->d_dtracks.ensure(std::max<size_t>(nb, 1) * sizeof(DecTrack)));
    CK(e->d_hblocks.ensure(std::max<size_t>(hb.size(), 1) * sizeof(HuffBlock)));
    CK(e->d_runs.ensure(std::max<size_t>(runs.size(), 1) * sizeof(RunDesc)));
    CK(e->d_spectra.ensure(std::max<uint64_t>(gc_total, 1) * 576 * sizeof(int16_t) + 64));
    CK(e->d_nz.ensure(std::max<uint64_t>(gc_total, 1) * sizeof(uint16_t) + 64));
    CK(e->d_pcm.ensure(std::max<uint64_t>(pcm_total, 4) * sizeof(float) + 64));
    for (size_t k = 0; k < nb; k++) {
        if (pt[k].status != RG_OK) continue;
        pt[k].d_ch[0] = e->d_pcm.as<float>() + dt[k].pcm_off;
        pt[k].d_ch[1] = pt[k].nch >= 2 ? pt[k].d_ch[0] + dt[k].pcm_plane : nullptr;
    }
    CK(cudaMemcpyAsync(e->d_dtracks.p, dt.data(), nb * sizeof(DecTrack), cudaMemcpyHostToDevice, e->stream));
    if (!hb.empty()) CK(cudaMemcpyAsync(e->d_hblocks.p, hb.data(), hb.size() * sizeof(HuffBlock), cudaMemcpyHostToDevice, e->stream));
    if (!runs.empty()) CK(cudaMemcpyAsync(e->d_runs.p, runs.data(), runs.size() * sizeof(RunDesc), cudaMemcpyHostToDevice, e->stream));
    e->stats.h2d_bytes += nb * sizeof(DecTrack) + hb.size() * sizeof(HuffBlock) + runs.size() * sizeof(RunDesc);
    CK(cudaEventRecord(e->ev[1], e->stream));
    if (!hb.empty()) {
        k_huffman<<<(unsigned)hb.size(), K1_THREADS, 0, e->stream>>>(c->d_bytes.as<uint8_t>(), e->d_md.as<uint8_t>(), e->d_frames.as<FrameRec>(),
                                                                  e->d_dtracks.as<DecTrack>(), e->d_hblocks.as<HuffBlock>(), e->dec.d_tab,
                                                                  e->d_spectra.as<int16_t>(), e->d_nz.as<uint16_t>());
        e->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e->ev[2], e->stream));
    if (!runs.empty()) {
        k_hybrid_synth<<<(unsigned)runs.size(), K2_THREADS, sizeof(K2Smem), e->stream>>>(c->d_bytes.as<uint8_t>(), e->d_md.as<uint8_t>(), e->d_frames.as<FrameRec>(),
                                                                         e->d_dtracks.as<DecTrack>(), e->d_runs.as<RunDesc>(), e->dec.d_tab,
                                                                         e->d_spectra.as<int16_t>(), e->d_nz.as<uint16_t>(), e->d_pcm.as<float>(), e->flags);
        e->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e->ev[3], e->stream));
    e->pending_decode_timing = true;
    return RG_OK;
}

static void collect_decode_timing(rg_engine* e) {
    if (!e->pending_decode_timing) return;
    e->pending_decode_timing = false;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]) == cudaSuccess) e->stats.last_ms_huffman += ms;
    if (cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]) == cudaSuccess) e->stats.last_ms_hybrid += ms;
}

// Bytes of device scratch one track needs in a batch (spectra + nz + PCM)
static uint64_t track_scratch_bytes(const TrackMeta& m) {
    return (uint64_t)m.n_frames * m.channels * (2ull * (576 * 2 + 2) + 1152ull * 4) + 4096;
}

// The whole pipeline over a device-resident corpus.
static int analyze_corpus_impl(rg_engine* e, rg_corpus* c, rg_track_result* out, uint32_t* hist_out,
                               unsigned long long* d_album_hist, unsigned long long* d_album_peak, bool album_fail_fast,
                               int* first_bad) {
    const size_t n = c->len.size();
    CK(cudaSetDevice(e->device));
    e->stats.last_ms_index = e->stats.last_ms_huffman = e->stats.last_ms_hybrid = e->stats.last_ms_synth = 0;
    e->stats.last_ms_iir = e->stats.last_ms_hist = e->stats.last_ms_total = 0;
    if (first_bad) *first_bad = -1;
    if (!n) return RG_OK;
    CK(cudaEventRecord(e->ev[7], e->stream));
    DecodeOut d;
    std::vector<uint64_t> frame_base, md_base;
    int rc = run_frame_index(e, c, d, frame_base, md_base);
    if (rc != RG_OK) return rc;
    for (size_t i = 0; i < n; i++) {
        if (d.status[i] == RG_OK && rate_index(d.sample_rate[i]) < 0) d.status[i] = RG_ERR_UNSUPPORTED_RATE;
        if (d.status[i] != RG_OK && first_bad && *first_bad < 0) *first_bad = (int)i;
    }
    if (album_fail_fast && first_bad && *first_bad >= 0) {
        // album mode: the first failing file aborts everything (src/replaygain.rs:1055)
        for (size_t i = 0; i < n; i++) {
            memset(&out[i], 0, sizeof out[i]);
            out[i].status = d.status[i];
            out[i].sample_rate = d.sample_rate[i];
            out[i].file_type = c->is_mp4[i] ? 1 : 0;
        }
        return d.status[*first_bad];
    }
    size_t t0 = 0;
    while (t0 < n) {
        size_t t1 = t0;
        uint64_t bytes = 0;
        while (t1 < n) {
            const uint64_t b = d.status[t1] == RG_OK ? track_scratch_bytes(d.meta[t1]) : 0;
            if (t1 > t0 && bytes + b > e->scratch_bytes) break;
            bytes += b;
            t1++;
        }
        std::vector<PcmTrack> pt;
        rc = run_decode_batch(e, c, d, frame_base, md_base, t0, t1, pt);
        if (rc != RG_OK) return rc;
        std::vector<RgTrackOut> res;
        std::vector<uint8_t> rerun;
        rc = run_stage_b(e, pt, res, rerun, hist_out ? hist_out + t0 * RG_HIST_BINS : nullptr, d_album_hist, d_album_peak);
        if (rc != RG_OK) return rc;
        collect_decode_timing(e);
        for (size_t k = 0; k < pt.size(); k++) {
            fill_result(&out[t0 + k], pt[k], res[k], rerun[k], d.meta[t0 + k].n_frames);
            out[t0 + k].file_type = c->is_mp4[t0 + k] ? 1 : 0;
            if (pt[k].status == RG_OK) { e->stats.tracks++; e->stats.frames += d.meta[t0 + k].n_frames; }
        }
        t0 = t1;
    }
    CK(cudaEventRecord(e->ev[0], e->stream));
    CK(cudaEventSynchronize(e->ev[0]));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev[7], e->ev[0]) == cudaSuccess) e->stats.last_ms_total = ms;
    return RG_OK;
}

extern "C" int rg_analyze_corpus(rg_engine* e, rg_corpus* c, rg_track_result* out, uint32_t* hist_out) {
    if (!e || !c || (!out && !c->len.empty())) return RG_ERR_INVALID_ARG;
    return analyze_corpus_impl(e, c, out, hist_out, nullptr, nullptr, false, nullptr);
}

extern "C" int rg_analyze_tracks(rg_engine* e, const rg_input* in, size_t n, rg_track_result* out, uint32_t* hist_out) {
    if (!e || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    rg_corpus* c = nullptr;
    int rc = rg_corpus_upload(e, in, n, &c);
    if (rc != RG_OK) return rc;
    rc = analyze_corpus_impl(e, c, out, hist_out, nullptr, nullptr, false, nullptr);
    rg_corpus_free(e, c);
    return rc;
}

static int album_buffers(rg_engine* e) {
    CK(e->d_album.ensure(RG_HIST_BINS * sizeof(unsigned long long) + 64));
    return RG_OK;
}

extern "C" int rg_album_finish(rg_engine* e, const uint64_t* d_album_hist, const double* d_album_peak, rg_album_result* album) {
    if (!e || !d_album_hist || !d_album_peak || !album) return RG_ERR_INVALID_ARG;
    CK(cudaSetDevice(e->device));
    CK(e->d_misc.ensure(64));
    k_album_percentile<<<1, K5_THREADS, 0, e->stream>>>(reinterpret_cast<const unsigned long long*>(d_album_hist),
                                                        reinterpret_cast<const unsigned long long*>(d_album_peak), e->d_misc.as<double>());
    e->stats.kernel_launches++;
    CK(cudaGetLastError());
    double h[3];
    CK(cudaMemcpyAsync(h, e->d_misc.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += sizeof h;
    album->album_loudness_db = h[0];
    album->album_gain_db = h[1];
    album->album_peak = h[2];
    return RG_OK;
}

extern "C" int rg_analyze_album_shard(rg_engine* e, rg_corpus* c, rg_track_result* out, uint64_t* d_album_hist, double* d_album_peak) {
    if (!e || !c || !d_album_hist || !d_album_peak || (!out && !c->len.empty())) return RG_ERR_INVALID_ARG;
    int bad = -1;
    return analyze_corpus_impl(e, c, out, nullptr, reinterpret_cast<unsigned long long*>(d_album_hist),
                               reinterpret_cast<unsigned long long*>(d_album_peak), true, &bad);
}

extern "C" int rg_analyze_album(rg_engine* e, const rg_input* in, size_t n, rg_track_result* out, rg_album_result* album,
                                uint64_t* album_hist_out) {
    if (!e || !album || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    int rc = album_buffers(e);
    if (rc != RG_OK) return rc;
    unsigned long long* d_hist = e->d_album.as<unsigned long long>();
    CK(e->d_album_peak.ensure(64));
    unsigned long long* d_peak = e->d_album_peak.as<unsigned long long>();
    CK(cudaMemsetAsync(d_hist, 0, RG_HIST_BINS * sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(d_peak, 0, sizeof(unsigned long long), e->stream));
    rg_corpus* c = nullptr;
    rc = rg_corpus_upload(e, in, n, &c);
    if (rc != RG_OK) return rc;
    int bad = -1;
    rc = analyze_corpus_impl(e, c, out, nullptr, d_hist, d_peak, true, &bad);
    rg_corpus_free(e, c);
    if (rc != RG_OK) return rc;
    rc = rg_album_finish(e, reinterpret_cast<const uint64_t*>(d_hist), reinterpret_cast<const double*>(d_peak), album);
    if (rc != RG_OK) return rc;
    if (album_hist_out) {
        CK(cudaMemcpyAsync(album_hist_out, d_hist, RG_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->stats.d2h_bytes += RG_HIST_BINS * sizeof(uint64_t);
    }
    return RG_OK;
}

extern "C" int rg_find_peak(rg_engine* e, const rg_input* in, size_t n, rg_peak_result* out) {
    // find_peak_amplitude (src/replaygain.rs:1140-1249): max |sample| over all channels of the decoded stream.  The
    // analysis pass already tracks exactly that, so this is the same device pipeline with the loudness fields dropped.
    if (!e || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    std::vector<rg_track_result> res(n);
    int rc = rg_analyze_tracks(e, in, n, res.data(), nullptr);
    if (rc != RG_OK) return rc;
    for (size_t i = 0; i < n; i++) {
        out[i].status = res[i].status;
        out[i].sample_rate = res[i].sample_rate;
        out[i].peak = res[i].status == RG_OK ? res[i].peak : 0.0;
        out[i].peak_pcm = out[i].peak * 32768.0;
    }
    return RG_OK;
}

extern "C" int rg_decode_pcm(rg_engine* e, const rg_input* in, float* pcm, size_t max_frames, size_t* frames_out,
                             uint32_t* channels_out, uint32_t* sample_rate_out) {
    if (!e || !in || !pcm || !frames_out) return RG_ERR_INVALID_ARG;
    rg_corpus* c = nullptr;
    int rc = rg_corpus_upload(e, in, 1, &c);
    if (rc != RG_OK) return rc;
    DecodeOut d;
    std::vector<uint64_t> frame_base, md_base;
    rc = run_frame_index(e, c, d, frame_base, md_base);
    if (rc == RG_OK && d.status[0] != RG_OK) rc = d.status[0];
    std::vector<PcmTrack> pt;
    if (rc == RG_OK) rc = run_decode_batch(e, c, d, frame_base, md_base, 0, 1, pt);
    if (rc == RG_OK) {
        const size_t frames = pt[0].frames, ncopy = std::min(frames, max_frames);
        for (int ch = 0; ch < pt[0].nch && rc == RG_OK; ch++)
            if (ncopy && cudaMemcpyAsync(pcm + (size_t)ch * max_frames, pt[0].d_ch[ch], ncopy * sizeof(float), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess)
                rc = RG_ERR_CUDA;
        if (cudaStreamSynchronize(e->stream) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "rg_decode_pcm sync", __LINE__);
        collect_decode_timing(e);
        *frames_out = frames;
        if (channels_out) *channels_out = (uint32_t)pt[0].nch;
        if (sample_rate_out) *sample_rate_out = pt[0].sample_rate;
        e->stats.d2h_bytes += ncopy * sizeof(float) * pt[0].nch;
    }
    rg_corpus_free(e, c);
    return rc;
}
