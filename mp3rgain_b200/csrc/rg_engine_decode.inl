// rg_engine_decode.inl -- entry points that run the MP3 decode kernels (stage A) and feed stage B.
// Included at the end of rg_engine.cu.
//
// The whole path is a pipeline over four streams:
//   copy    file images host -> HBM, group by group                       (end-to-end entry points only)
//   index   K0 per group as soon as its bytes are resident; frame counts back to pinned memory
//   decode  K1 + K2 of batch k+1                                           (integer + f32 pipes)
//   main    K4 + K5 of batch k, results back to pinned memory              (FP64 pipe)
// The host plans batch k+1 (run / chain descriptors need K0's frame counts) while the device works on batch k; two
// BatchCtx alternate, each with half of the scratch budget.

#define RUN_GRANULES 254  // default granules per K2 CTA (even: runs start on granule 0 of a frame; +2 replayed = K2_MAX_G)

static const uint32_t MPEG1_RATES[3] = {44100, 48000, 32000};
// TrackMeta.sr_idx = sample-rate index | MPEG version << 4 (3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5)
static inline uint32_t meta_version(const TrackMeta& m) { return (m.sr_idx >> 4) & 3; }
static inline uint32_t meta_gpf(const TrackMeta& m) { return meta_version(m) == 3 ? 2u : 1u; }           // granules per frame
static inline uint32_t meta_sr_row(const TrackMeta& m) { return (meta_version(m) == 3 ? 0u : (meta_version(m) == 2 ? 3u : 6u)) + (m.sr_idx & 3); }
static inline uint32_t meta_rate(const TrackMeta& m) { return MPEG1_RATES[m.sr_idx & 3] >> (meta_version(m) == 3 ? 0 : (meta_version(m) == 2 ? 1 : 2)); }

// mp4meta::is_mp4_file (src/mp4meta.rs:872-889): ftyp box with a known brand => AudioFileType::Aac
static bool sniff_mp4(const uint8_t* d, uint64_t len) {
    if (len < 12) return false;
    const uint32_t size = ((uint32_t)d[0] << 24) | ((uint32_t)d[1] << 16) | ((uint32_t)d[2] << 8) | d[3];
    if (memcmp(d + 4, "ftyp", 4) != 0 || size < 12) return false;
    static const char* brands[] = {"M4A ", "M4B ", "M4P ", "M4V ", "mp41", "mp42", "isom", "iso2"};
    for (const char* b : brands) if (!memcmp(d + 8, b, 4)) return true;
    return false;
}

// Host-side description of n file images laid out back to back (16-byte aligned starts, >= 16 bytes of slack)
static int corpus_describe(rg_corpus* c, const rg_input* in, size_t n, bool allow_arena = true) {
    c->off.resize(n); c->len.resize(n); c->track_index.resize(n); c->is_mp4.resize(n);
    // An ARENA: the non-empty buffers ascend through host memory, at most 64 bytes apart (every byte of such a gap lies on a page
    // that holds the end of one buffer or the start of the next), at 16-byte multiples from the first.  The device then keeps the
    // caller's spacing, and a group of images crosses the bus as ONE copy instead of one per file (a loader that reads files
    // into a ring of pinned memory produces exactly this; 6000 cudaMemcpyAsync calls are ~20 ms of host time per call).
    bool arena = allow_arena && n >= 2;
    const uint8_t* base = nullptr;
    const uint8_t* prev_end = nullptr;
    for (size_t i = 0; i < n && arena; i++) {
        if (!in[i].len) continue;
        const uint8_t* p = static_cast<const uint8_t*>(in[i].data);
        if (!p) break;   // (reported below)
        if (!base) base = p;
        else if (p < prev_end || p - prev_end > 64 || ((p - base) & 15)) arena = false;
        prev_end = p + in[i].len;
    }
    c->arena = arena && base != nullptr;
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (in[i].len && !in[i].data) return RG_ERR_INVALID_ARG;
        if (c->arena) {
            if (in[i].len) { c->off[i] = (uint64_t)(static_cast<const uint8_t*>(in[i].data) - base); total = c->off[i] + in[i].len; }
            else c->off[i] = (total + 15) & ~15ull;
        } else {
            c->off[i] = total;
            total += (in[i].len + 16 + 15) & ~15ull;
        }
        c->len[i] = in[i].len;
        c->track_index[i] = in[i].track_index;
        c->is_mp4[i] = sniff_mp4(in[i].data, in[i].len) ? 1 : 0;
    }
    if (c->arena) total = ((total + 15) & ~15ull) + 16;   // (every image is followed by >= 16 bytes inside the allocation)
    c->total = total;
    return RG_OK;
}

// Host -> device copy of the file images of tracks [t0, t1) on stream st: one cudaMemcpyAsync per image, or -- when the
// buffers form an arena -- one per run of images of up to 4 MB.  (One copy for a whole group was 1.5 ms per call SLOWER for
// 2.9 MB files than one per file; what hurts is the call count of small files: 6000 clips of 481 KB.)
static int copy_images(rg_engine* e, rg_corpus* c, const rg_input* in, size_t t0, size_t t1, cudaStream_t st) {
    if (c->arena) {
        const uint64_t run_max = 4ull << 20;
        size_t i = t0;
        while (i < t1) {
            if (!c->len[i]) { i++; continue; }
            const size_t first = i;
            const uint64_t lo = c->off[i];
            uint64_t hi = lo + c->len[i];
            for (i++; i < t1; i++) {
                if (!c->len[i]) continue;
                if (c->off[i] + c->len[i] - lo > run_max) break;
                hi = c->off[i] + c->len[i];
            }
            CK(cudaMemcpyAsync(c->d_bytes + lo, in[first].data, hi - lo, cudaMemcpyHostToDevice, st));
            e->stats.h2d_bytes += hi - lo;
        }
        return RG_OK;
    }
    for (size_t i = t0; i < t1; i++) {
        if (!c->len[i]) continue;
        CK(cudaMemcpyAsync(c->d_bytes + c->off[i], in[i].data, c->len[i], cudaMemcpyHostToDevice, st));
        e->stats.h2d_bytes += c->len[i];
    }
    return RG_OK;
}

extern "C" int rg_corpus_upload(rg_engine* e, const rg_input* in, size_t n, rg_corpus** out) {
    if (!e || !out || (n && !in)) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    *out = nullptr;
    CK(cudaSetDevice(e->device));
    rg_corpus* c = new rg_corpus();
    int rc = corpus_describe(c, in, n, e->arena);
    if (rc != RG_OK) { delete c; return rc; }
    cudaError_t err = c->d_own.ensure(c->total + 64);
    if (err != cudaSuccess) { delete c; return fail_cuda(err, "cudaMalloc(corpus)", __LINE__); }
    c->d_bytes = c->d_own.as<uint8_t>();
    rc = copy_images(e, c, in, 0, n, e->stream);
    if (rc != RG_OK) { c->d_own.release(); delete c; return rc; }
    err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) { c->d_own.release(); delete c; return fail_cuda(err, "cudaStreamSynchronize", __LINE__); }
    *out = c;
    return RG_OK;
}

extern "C" void rg_corpus_free(rg_engine* e, rg_corpus* c) {
    if (!c) return;
    std::unique_lock<std::recursive_mutex> lock;
    if (e) { lock = std::unique_lock<std::recursive_mutex>(e->mu); cudaSetDevice(e->device); }
    c->d_own.release();
    delete c;
}

extern "C" uint64_t rg_corpus_bytes(const rg_corpus* c) {
    uint64_t s = 0;
    if (c) for (uint64_t l : c->len) s += l;
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
struct IndexPlan {                      // K0 bookkeeping of one analyse call
    std::vector<uint64_t> frame_base, md_base;
    std::vector<size_t> group_begin;    // tracks [group_begin[g], group_begin[g+1]) form index group g
    std::vector<int> status;            // rg_status per track after host-side checks (valid once its group is absorbed)
    std::vector<uint32_t> sample_rate;
    size_t groups_absorbed = 0, groups_enqueued = 0;
    std::vector<TrackSeg> tseg;         // per track: its speculative segments
    size_t n_segs = 0;
    const TrackMeta* meta = nullptr;    // pinned, written by the index stream
};

// Enqueues the copy (if the corpus still lives in host memory) and K0 of every group.  Returns without waiting.
static int index_enqueue(rg_engine* e, rg_corpus* c, IndexPlan& ip) {
    const size_t n = c->len.size();
    ip.status.assign(n, RG_OK);
    ip.sample_rate.assign(n, 0);
    ip.frame_base.resize(n + 1);
    ip.md_base.resize(n + 1);
    ip.group_begin.clear();
    ip.groups_absorbed = ip.groups_enqueued = 0;
    std::vector<TrackIn> tin(n);
    std::vector<TrackSeg>& tseg = ip.tseg;
    tseg.assign(n, TrackSeg{});
    std::vector<SegDesc> segs;
    uint64_t fb = 0, mb = 0, bytes = 0;
    for (size_t i = 0; i < n; i++) {
        tin[i].off = c->off[i]; tin[i].len = c->len[i];
        tseg[i].seg_base = (uint32_t)segs.size();
        tseg[i].n_seg = (uint32_t)std::max<uint64_t>(1, (c->len[i] + K0_SEG - 1) / K0_SEG);
        for (uint32_t sg = 0; sg < tseg[i].n_seg; sg++) segs.push_back(SegDesc{(uint32_t)i, sg});
        ip.frame_base[i] = fb; ip.md_base[i] = mb;
        fb += c->len[i] / 24 + 2;                      // smallest Layer III frame: 24 bytes (MPEG-2, 8 kbps at 24 kHz)
        mb += (c->len[i] + 32 + 15) & ~15ull;
        bytes += c->len[i];
    }
    ip.frame_base[n] = fb; ip.md_base[n] = mb;
    // index groups: one when the corpus is already resident (K0 is a serial walk per track, so its time does not depend on
    // the number of tracks); about eight, never smaller than 32 MiB, when the bytes are still arriving from the host, so
    // that indexing starts as soon as the first group has landed
    const uint64_t gbytes = (e->pipeline && c->pending_in) ? std::max<uint64_t>(bytes / 8 + 1, e->group_bytes_min) : ~0ull;
    // The first groups are smaller -- gbytes / group_first_div, each next one group_growth times its predecessor until gbytes is
    // reached -- so that decoding starts as soon as a few files have landed.  The bytes arrive about 1.7 times faster than the
    // kernels consume them, so a group may be up to that much larger than the one being decoded without starving the decode
    // stream; doubling (the first schedule: 1/4, 1/2, 1) left 1.7 and 3.0 ms of bubbles behind the first two batches
    // (profiles/r2_timeline_e2e_groups.log).
    // Shrinking the LAST groups as well -- so that a bus-bound call ends sooner after its last byte -- changed nothing, neither on
    // one GPU (pinned or pageable: profiles/r2_e2e_group_sweep.log) nor with eight ranks pulling through one host at once
    // (profiles/r2_e2e_8gpu_tail_groups.log), and was removed again.
    uint64_t acc = 0;
    double lim = gbytes == ~0ull ? 0.0 : std::max<double>((double)gbytes / e->group_first_div, 1.0);
    for (size_t i = 0; i < n; i++) {
        if (i == 0 || (gbytes != ~0ull && (double)acc >= lim)) {
            if (i) lim = std::min<double>((double)gbytes, lim * e->group_growth);
            ip.group_begin.push_back(i);
            acc = 0;
        }
        acc += c->len[i];
    }
    ip.group_begin.push_back(n);
    if (!n) return RG_OK;
    CK(e->d_tin.ensure(n * sizeof(TrackIn)));
    CK(e->d_fbase.ensure((n + 1) * sizeof(uint64_t) * 2));
    CK(e->d_frames.ensure(fb * sizeof(FrameRec)));
    CK(e->d_md.ensure(mb + 64));
    CK(e->d_meta.ensure(n * sizeof(TrackMeta)));
    CK(e->h_meta.ensure(n * sizeof(TrackMeta)));
    ip.meta = e->h_meta.as<TrackMeta>();
    CK(e->d_tseg.ensure(n * sizeof(TrackSeg)));
    CK(e->d_segs.ensure(segs.size() * sizeof(SegDesc)));
    CK(e->d_segout.ensure(segs.size() * sizeof(SegOut)));
    CK(e->d_segfin.ensure(segs.size() * sizeof(SegFin)));
    CK(e->d_stage.ensure(segs.size() * (size_t)K0_SEG_RECS * sizeof(SegRec)));
    CK(cudaMemcpyAsync(e->d_tseg.p, tseg.data(), n * sizeof(TrackSeg), cudaMemcpyHostToDevice, e->s_index));
    CK(cudaMemcpyAsync(e->d_segs.p, segs.data(), segs.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, e->s_index));
    CK(cudaMemcpyAsync(e->d_tin.p, tin.data(), n * sizeof(TrackIn), cudaMemcpyHostToDevice, e->s_index));
    CK(cudaMemcpyAsync(e->d_fbase.p, ip.frame_base.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, e->s_index));
    CK(cudaMemcpyAsync(e->d_fbase.as<uint64_t>() + (n + 1), ip.md_base.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, e->s_index));
    CK(cudaStreamSynchronize(e->s_index));   // the three sources above are stack / heap temporaries
    e->stats.h2d_bytes += n * (sizeof(TrackIn) + sizeof(TrackSeg)) + 2 * (n + 1) * sizeof(uint64_t) + segs.size() * sizeof(SegDesc);
    ip.n_segs = segs.size();
    CK(cudaEventRecord(e->ev[1], e->s_index));   // the descriptor tables above are in place
    return RG_OK;
}

// Enqueues the copy (if the group's bytes still live in host memory) and the frame index of group g.  Groups are enqueued a
// few ahead of the batch loop rather than all at once, so that the host reaches the first decode launch without first paying
// for every cudaMemcpyAsync call of the corpus.
static int index_enqueue_group(rg_engine* e, rg_corpus* c, IndexPlan& ip, size_t g) {
    const size_t n = c->len.size();
    const std::vector<TrackSeg>& tseg = ip.tseg;
    {
        const size_t t0 = ip.group_begin[g], t1 = ip.group_begin[g + 1];
        cudaStream_t si = (g & 3) == 0 ? e->s_index : e->s_index_x[(g & 3) - 1];
        if (si != e->s_index) CK(cudaStreamWaitEvent(si, e->ev[1], 0));
        CK(e->reserve_events(4 * g + 4));
        cudaEvent_t ev_copied = e->pool_event(4 * g), ev_k0a = e->pool_event(4 * g + 1), ev_k0b = e->pool_event(4 * g + 2),
                    ev_done = e->pool_event(4 * g + 3);
        if (c->pending_in) {
            if (int r = copy_images(e, c, c->pending_in, t0, t1, e->s_copy)) return r;
            CK(cudaEventRecord(ev_copied, e->s_copy));
            CK(cudaStreamWaitEvent(si, ev_copied, 0));
        }
        CK(cudaEventRecord(ev_k0a, si));
        {
            const uint32_t seg0 = tseg[t0].seg_base;
            const uint32_t nsg = (t1 < n ? tseg[t1].seg_base : (uint32_t)ip.n_segs) - seg0;
            const unsigned gs = (nsg + K0_WARPS - 1) / K0_WARPS, gt = (unsigned)((t1 - t0 + K0_WARPS - 1) / K0_WARPS);
            k_index_scan<<<gs, 32 * K0_WARPS, 0, si>>>(c->d_bytes, e->d_tin.as<TrackIn>(), e->d_tseg.as<TrackSeg>(), e->d_segs.as<SegDesc>(),
                                                       seg0, nsg, e->d_stage.as<SegRec>(), e->d_segout.as<SegOut>(), e->flags);
            k_index_merge<<<gt, 32 * K0_WARPS, 0, si>>>(c->d_bytes, e->d_tin.as<TrackIn>() + t0, e->d_tseg.as<TrackSeg>() + t0,
                                                        e->d_fbase.as<uint64_t>() + (n + 1) + t0, e->d_stage.as<SegRec>(),
                                                        e->d_segout.as<SegOut>(), e->d_segfin.as<SegFin>(), e->d_md.as<uint8_t>(),
                                                        e->d_meta.as<TrackMeta>() + t0, (uint32_t)(t1 - t0), e->flags);
            k_index_finalize<<<gs, 32 * K0_WARPS, 0, si>>>(c->d_bytes, e->d_tin.as<TrackIn>(), e->d_segs.as<SegDesc>(), seg0, nsg,
                                                           e->d_fbase.as<uint64_t>(), e->d_fbase.as<uint64_t>() + (n + 1),
                                                           e->d_stage.as<SegRec>(), e->d_segfin.as<SegFin>(), e->d_frames.as<FrameRec>(),
                                                           e->d_md.as<uint8_t>());
            e->stats.kernel_launches += 3;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev_k0b, si));
        CK(cudaMemcpyAsync(e->h_meta.as<TrackMeta>() + t0, e->d_meta.as<TrackMeta>() + t0, (t1 - t0) * sizeof(TrackMeta),
                           cudaMemcpyDeviceToHost, si));
        e->stats.d2h_bytes += (t1 - t0) * sizeof(TrackMeta);
        CK(cudaEventRecord(ev_done, si));
    }
    return RG_OK;
}

static int index_enqueue_upto(rg_engine* e, rg_corpus* c, IndexPlan& ip, size_t g_end) {
    const size_t ng = ip.group_begin.size() - 1;
    while (ip.groups_enqueued < std::min(ng, g_end)) {
        int rc = index_enqueue_group(e, c, ip, ip.groups_enqueued);
        if (rc != RG_OK) return rc;
        ip.groups_enqueued++;
    }
    return RG_OK;
}

// Waits until track `upto - 1` has been indexed and derives the per-track status of every newly finished group.
// With `opportunistic`, also absorbs later groups that happen to be finished already.
static int index_absorb(rg_engine* e, rg_corpus* c, IndexPlan& ip, size_t upto, bool opportunistic) {
    const size_t ng = ip.group_begin.size() - 1;
    while (ip.groups_absorbed < ng) {
        const size_t g = ip.groups_absorbed;
        const bool needed = ip.group_begin[g] < upto;
        if (int rc = index_enqueue_upto(e, c, ip, g + (needed ? 4 : 0))) return rc;
        if (g >= ip.groups_enqueued) break;
        if (!needed) {
            if (!opportunistic || cudaEventQuery(e->pool_event(4 * g + 3)) != cudaSuccess) break;
        }
        CK(cudaEventSynchronize(e->pool_event(4 * g + 3)));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->pool_event(4 * g + 1), e->pool_event(4 * g + 2)) == cudaSuccess) e->stats.last_ms_index += ms;
        if (e->timeline) {
            float t[4] = {};
            if (c->pending_in) cudaEventElapsedTime(&t[0], e->ev[7], e->pool_event(4 * g));
            for (int q = 1; q < 4; q++) cudaEventElapsedTime(&t[q], e->ev[7], e->pool_event(4 * g + q));
            fprintf(stderr, "[rg timeline] group %2zu (%4zu tracks): copied %6.2f  K0 %6.2f-%6.2f  meta on host %6.2f  absorbed at host time %6.2f\n",
                    g, ip.group_begin[g + 1] - ip.group_begin[g], t[0], t[1], t[2], t[3], e->host_ms());
        }
        for (size_t i = ip.group_begin[g]; i < ip.group_begin[g + 1]; i++) {
            const TrackMeta& m = ip.meta[i];
            int st = m.status;
            if (c->is_mp4[i]) st = RG_ERR_UNSUPPORTED_FORMAT;          // AAC/M4A: out of scope for this engine
            else if (st == RG_OK) {
                ip.sample_rate[i] = meta_rate(m);
                // MP3 carries exactly one audio track (src/replaygain.rs:838-851)
                if (c->track_index[i] > 0) st = RG_ERR_TRACK_INDEX;
                else if (c->track_index[i] < -1) st = RG_ERR_INVALID_ARG;
                else if (rate_index(ip.sample_rate[i]) < 0) st = RG_ERR_UNSUPPORTED_RATE;
            }
            ip.status[i] = st;
        }
        ip.groups_absorbed++;
    }
    return RG_OK;
}

// Enqueues K1 + K2 for the tracks `x.ids` of the corpus on the decode stream; fills x.pt with the PCM views.
static int decode_enqueue(rg_engine* e, rg_corpus* c, const IndexPlan& ip, BatchCtx& x, bool peak_only = false) {
    const size_t nb = x.ids.size();
    x.pt.assign(nb, PcmTrack{});
    x.decoded = false;
    std::vector<DecTrack> dt(nb);
    std::vector<HuffBlock> hb;
    std::vector<RunDesc> runs;
    uint64_t gc_total = 0, pcm_total = 0;
    for (size_t k = 0; k < nb; k++) {
        const size_t i = x.ids[k];
        const TrackMeta& m = ip.meta[i];
        DecTrack& t = dt[k];
        memset(&t, 0, sizeof t);
        PcmTrack& p = x.pt[k];
        p.status = ip.status[i];
        p.sample_rate = ip.sample_rate[i];
        p.rate_idx = rate_index(p.sample_rate);
        p.nch = (int)m.channels;
        if (p.status != RG_OK) continue;
        t.corpus_off = c->off[i];
        t.md_off = ip.md_base[i];
        t.frame_base = ip.frame_base[i];
        t.gc_base = gc_total;
        const uint32_t gpf = meta_gpf(m), n_gran = m.n_frames * gpf;
        t.n_frames = m.n_frames; t.channels = m.channels; t.sr_idx = meta_sr_row(m); t.md_bytes = m.md_bytes; t.gpf = gpf;
        const uint64_t plane = (uint64_t)n_gran * 576;
        t.pcm_off = pcm_total; t.pcm_plane = plane;
        const uint32_t n_gc = n_gran * m.channels;
        for (uint32_t g0 = 0; g0 < n_gc; g0 += K1_THREADS) hb.push_back(HuffBlock{(uint32_t)k, g0});
        for (uint32_t g0 = 0; g0 < n_gran; g0 += e->run_granules)
            runs.push_back(RunDesc{(uint32_t)k, g0, std::min<uint32_t>(e->run_granules, n_gran - g0), 0});
        gc_total += n_gc;
        pcm_total += plane * m.channels;
        p.frames = plane;
    }
    CK(x.d_dtracks.ensure(std::max<size_t>(nb, 1) * sizeof(DecTrack)));
    CK(x.d_hblocks.ensure(std::max<size_t>(hb.size(), 1) * sizeof(HuffBlock)));
    CK(x.d_runs.ensure(std::max<size_t>(runs.size(), 1) * sizeof(RunDesc)));
    CK(x.d_spectra.ensure(std::max<uint64_t>(gc_total, 1) * 576 * sizeof(int16_t) + 64));
    CK(x.d_nz.ensure(std::max<uint64_t>(gc_total, 1) * sizeof(uint16_t) + 64));
    CK(x.d_c1q.ensure(std::max<uint64_t>(gc_total, 1) * K1_C1_ROW + 64));
    CK(x.d_gains.ensure(std::max<uint64_t>(gc_total, 1) * K1_GAIN_ROW + 64));
    CK(x.d_sfraw.ensure(std::max<uint64_t>(gc_total, 1) * K1_GAIN_ROW + 64));
    if (!peak_only) CK(x.d_pcm.ensure(std::max<uint64_t>(pcm_total, 4) * sizeof(float) + 64));
    if (peak_only) {
        CK(x.d_tpeak.ensure(std::max<size_t>(nb, 1) * sizeof(uint32_t)));
        CK(x.h_tout.ensure(std::max<size_t>(nb, 1) * sizeof(uint32_t)));
        CK(cudaMemsetAsync(x.d_tpeak.p, 0, std::max<size_t>(nb, 1) * sizeof(uint32_t), e->s_dec));
    }
    for (size_t k = 0; k < nb; k++) {
        if (x.pt[k].status != RG_OK) continue;
        if (peak_only) continue;
        x.pt[k].d_ch[0] = x.d_pcm.as<float>() + dt[k].pcm_off;
        x.pt[k].d_ch[1] = x.pt[k].nch >= 2 ? x.pt[k].d_ch[0] + dt[k].pcm_plane : nullptr;
    }
    // descriptors through pinned staging: [dtracks] [hblocks] [runs]
    const size_t b_dt = up16(nb * sizeof(DecTrack)), b_hb = up16(hb.size() * sizeof(HuffBlock)), b_rn = up16(runs.size() * sizeof(RunDesc));
    CK(x.h_desc_a.ensure(b_dt + b_hb + b_rn + 64));
    uint8_t* hd = x.h_desc_a.as<uint8_t>();
    memcpy(hd, dt.data(), nb * sizeof(DecTrack));
    if (b_hb) memcpy(hd + b_dt, hb.data(), hb.size() * sizeof(HuffBlock));
    if (b_rn) memcpy(hd + b_dt + b_hb, runs.data(), runs.size() * sizeof(RunDesc));
    cudaStream_t st = e->s_dec;
    if (int r = fetch_desc(e, st, x.d_dtracks.p, hd, b_dt)) return r;
    if (int r = fetch_desc(e, st, x.d_hblocks.p, hd + b_dt, b_hb)) return r;
    if (int r = fetch_desc(e, st, x.d_runs.p, hd + b_dt + b_hb, b_rn)) return r;
    CK(cudaEventRecord(x.ev_a[0], st));
    if (!runs.empty()) {
        k_md_gather<<<(unsigned)runs.size(), K0B_THREADS, 0, st>>>(c->d_bytes, e->d_md.as<uint8_t>(), e->d_frames.as<FrameRec>(),
                                                                 x.d_dtracks.as<DecTrack>(), x.d_runs.as<RunDesc>());
        e->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(x.ev_a[3], st));
    if (!hb.empty()) {
        k_huffman<<<(unsigned)hb.size(), K1_THREADS, sizeof(K1Smem), st>>>(c->d_bytes, e->d_md.as<uint8_t>(), e->d_frames.as<FrameRec>(),
                                                           x.d_dtracks.as<DecTrack>(), x.d_hblocks.as<HuffBlock>(), e->dec.d_tab,
                                                           x.d_spectra.as<int16_t>(), x.d_nz.as<uint16_t>(), x.d_c1q.as<uint8_t>(), x.d_gains.as<uint8_t>(), x.d_sfraw.as<uint8_t>());
        e->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(x.ev_a[1], st));
    if (!runs.empty()) {
        k_hybrid_synth<<<(unsigned)runs.size(), K2_THREADS, sizeof(K2Smem), st>>>(c->d_bytes, e->d_frames.as<FrameRec>(),
                                                                               x.d_dtracks.as<DecTrack>(), x.d_runs.as<RunDesc>(), e->dec.d_tab,
                                                                               x.d_spectra.as<int16_t>(), x.d_c1q.as<uint8_t>(), x.d_nz.as<uint16_t>(),
                                                                               x.d_gains.as<uint8_t>(), x.d_sfraw.as<uint8_t>(), peak_only ? nullptr : x.d_pcm.as<float>(),
                                                                               peak_only ? x.d_tpeak.as<uint32_t>() : nullptr, e->flags);
        e->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(x.ev_a[2], st));
    if (peak_only) {
        CK(cudaMemcpyAsync(x.h_tout.p, x.d_tpeak.p, std::max<size_t>(nb, 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        e->stats.d2h_bytes += nb * sizeof(uint32_t);
        CK(cudaEventRecord(x.ev_b[3], st));
    }
    x.decoded = true;
    return RG_OK;
}

static void collect_decode_timing(rg_engine* e, BatchCtx& x) {
    if (!x.decoded) return;
    x.decoded = false;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, x.ev_a[0], x.ev_a[3]) == cudaSuccess) e->stats.last_ms_index += ms;
    if (cudaEventElapsedTime(&ms, x.ev_a[3], x.ev_a[1]) == cudaSuccess) e->stats.last_ms_huffman += ms;
    if (cudaEventElapsedTime(&ms, x.ev_a[1], x.ev_a[2]) == cudaSuccess) e->stats.last_ms_hybrid += ms;
    if (e->timeline) {   // development aid: where the batch's kernels sat in the call (ms since the call's first event)
        float t[8] = {};
        cudaEventElapsedTime(&t[0], e->ev[7], x.ev_a[0]); cudaEventElapsedTime(&t[1], e->ev[7], x.ev_a[3]);
        cudaEventElapsedTime(&t[2], e->ev[7], x.ev_a[1]); cudaEventElapsedTime(&t[3], e->ev[7], x.ev_a[2]);
        cudaEventElapsedTime(&t[4], e->ev[7], x.ev_b[0]); cudaEventElapsedTime(&t[5], e->ev[7], x.ev_b[1]);
        cudaEventElapsedTime(&t[6], e->ev[7], x.ev_b[2]); cudaEventElapsedTime(&t[7], e->ev[7], x.ev_b[3]);
        fprintf(stderr, "[rg timeline] batch of %4zu tracks: gather %6.2f K1 %6.2f-%6.2f K2 -%6.2f | K4 %6.2f-%6.2f K5 -%6.2f results %6.2f\n",
                x.ids.size(), t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
    }
}

// Bytes of device scratch one track needs in a batch (spectra + nz + band gains + PCM)
static uint64_t track_scratch_bytes(const TrackMeta& m) {
    return (uint64_t)m.n_frames * meta_gpf(m) * m.channels * ((576 * 2 + 2 + 2 * K1_GAIN_ROW + K1_C1_ROW) + 576ull * 4) + 4096;
}

static void drain_streams(rg_engine* e) {
    for (cudaStream_t st : {e->s_copy, e->s_index, e->s_index_x[0], e->s_index_x[1], e->s_index_x[2], e->s_dec, e->s_k4x[0], e->s_k4x[1], e->s_k4x[2], e->stream}) cudaStreamSynchronize(st);
    for (auto& x : e->ctx) { x.busy = false; x.decoded = false; x.hist_host = nullptr; }
}

// The whole pipeline over a corpus (resident, or still arriving from c->pending_in).
static int analyze_corpus_impl(rg_engine* e, rg_corpus* c, rg_track_result* out, uint32_t* hist_out,
                               unsigned long long* d_album_hist, unsigned long long* d_album_peak, bool album_fail_fast,
                               int* first_bad) {
    const size_t n = c->len.size();
    CK(cudaSetDevice(e->device));
    e->stats.last_ms_index = e->stats.last_ms_huffman = e->stats.last_ms_hybrid = e->stats.last_ms_synth = 0;
    e->stats.last_ms_iir = e->stats.last_ms_hist = e->stats.last_ms_total = 0;
    e->stats.last_batches = 0;
    e->stats.last_iir_samples = 0;
    if (first_bad) *first_bad = -1;
    if (!n) return RG_OK;
    CK(cudaEventRecord(e->ev[7], e->stream));
    e->t0_host = std::chrono::steady_clock::now();
    const bool exact_all = (e->flags & RG_FLAG_EXACT_IIR) != 0;
    IndexPlan ip;
    int rc = index_enqueue(e, c, ip);
    if (rc != RG_OK) { drain_streams(e); return rc; }
    if (album_fail_fast) {
        // album mode: the first failing file aborts everything (src/replaygain.rs:1055), before any histogram is merged
        rc = index_absorb(e, c, ip, n, false);
        if (rc != RG_OK) { drain_streams(e); return rc; }
        int bad = -1;
        for (size_t i = 0; i < n && bad < 0; i++) if (ip.status[i] != RG_OK) bad = (int)i;
        if (first_bad) *first_bad = bad;
        if (bad >= 0) {
            for (size_t i = 0; i < n; i++) {
                memset(&out[i], 0, sizeof out[i]);
                out[i].status = ip.status[i];
                out[i].sample_rate = ip.sample_rate[i];
                out[i].file_type = c->is_mp4[i] ? 1 : 0;
            }
            drain_streams(e);
            return ip.status[bad];
        }
    }
    const uint64_t budget = e->pipeline ? e->scratch_bytes / 2 : e->scratch_bytes;
    std::vector<uint32_t> flagged;
    std::vector<uint32_t> hist_tmp;   // histograms of a re-run batch, scattered back to the caller's rows afterwards

    auto collect = [&](BatchCtx& x, bool rerun_batch) -> int {
        if (!x.busy) return RG_OK;
        std::vector<RgTrackOut> res;
        int r = stage_b_collect(e, x, res);
        if (r != RG_OK) return r;
        collect_decode_timing(e, x);
        for (size_t k = 0; k < x.ids.size(); k++) {
            const uint32_t i = x.ids[k];
            if (!x.exact && x.pt[k].status == RG_OK && res[k].n_edge > 0) { flagged.push_back(i); }
            fill_result(&out[i], x.pt[k], res[k], rerun_batch ? 1 : 0, ip.meta[i].n_frames);
            out[i].file_type = c->is_mp4[i] ? 1 : 0;
            out[i].index_rewalks = ip.meta[i].n_rewalks;
            if (rerun_batch && hist_out) memcpy(hist_out + (size_t)i * RG_HIST_BINS, hist_tmp.data() + k * RG_HIST_BINS, RG_HIST_BINS * sizeof(uint32_t));
            if (!rerun_batch && x.pt[k].status == RG_OK) { e->stats.tracks++; e->stats.frames += ip.meta[i].n_frames; }
        }
        x.busy = false;
        x.hist_host = nullptr;
        return RG_OK;
    };
    auto fail = [&](int r) { drain_streams(e); return r; };

    size_t next = 0, k = 0;
    while (next < n) {
        rc = index_absorb(e, c, ip, next + 1, true);
        if (rc != RG_OK) return fail(rc);
        const size_t indexed_end = ip.group_begin[ip.groups_absorbed];
        // Equal batches instead of full ones plus a remainder: the scratch the rest of the call needs (exact for indexed
        // tracks, by the bytes-to-scratch ratio seen so far for the ones still arriving) is split evenly over the number of
        // batches it takes anyway, so that no call ends on a sliver of a batch that cannot fill the device.
        uint64_t need = 0, seen_scratch = 0, seen_len = 0, rest_len = 0;
        for (size_t i = next; i < n; i++) {
            if (i < indexed_end) {
                const uint64_t b = ip.status[i] == RG_OK ? track_scratch_bytes(ip.meta[i]) : 0;
                need += b; seen_scratch += b; seen_len += c->len[i];
            } else rest_len += c->len[i];
        }
        if (rest_len) need += seen_len ? (uint64_t)((double)seen_scratch / (double)seen_len * (double)rest_len) : rest_len * 37;
        const uint64_t nb_rest = std::max<uint64_t>(1, (need + budget - 1) / budget);
        const uint64_t even = std::min<uint64_t>(budget, need / nb_rest + need / nb_rest / 64 + (1ull << 20));
        size_t t1 = next;
        uint64_t bytes = 0;
        while (t1 < indexed_end) {
            const uint64_t b = ip.status[t1] == RG_OK ? track_scratch_bytes(ip.meta[t1]) : 0;
            if (t1 > next && bytes + b > even) break;
            bytes += b;
            t1++;
        }
        BatchCtx& x = e->ctx[e->pipeline ? (k & 1) : 0];
        rc = collect(x, false);
        if (rc != RG_OK) return fail(rc);
        x.ids.resize(t1 - next);
        for (size_t i = next; i < t1; i++) x.ids[i - next] = (uint32_t)i;
        x.hist_host = hist_out ? hist_out + next * RG_HIST_BINS : nullptr;
        x.images_arriving = c->pending_in != nullptr && ip.groups_absorbed + 1 < ip.group_begin.size();   // (frame-index kernels may still run beside this batch)
        const double h0 = e->timeline ? e->host_ms() : 0.0;
        rc = decode_enqueue(e, c, ip, x);
        const double h1 = e->timeline ? e->host_ms() : 0.0;
        if (rc == RG_OK) rc = stage_b_enqueue(e, x, exact_all, hist_out != nullptr, d_album_hist, d_album_peak, !exact_all, true);
        if (e->timeline) fprintf(stderr, "[rg timeline] host: batch %zu (%zu tracks) decode_enqueue %6.2f-%6.2f stage_b_enqueue -%6.2f\n", k, t1 - next, h0, h1, e->host_ms());
        if (rc != RG_OK) return fail(rc);
        x.busy = true;
        e->stats.last_batches++;
        if (!e->pipeline) { rc = collect(x, false); if (rc != RG_OK) return fail(rc); }
        next = t1;
        k++;
    }
    for (size_t j = 0; j < 2; j++) {   // oldest first
        rc = collect(e->ctx[(k + j) & 1], false);
        if (rc != RG_OK) return fail(rc);
    }
    // tracks whose time-blocked pass left a window on a bin edge: decode again, exact kernel (rare: ~1e-5 of tracks)
    size_t f0 = 0;
    while (f0 < flagged.size()) {
        size_t f1 = f0;
        uint64_t bytes = 0;
        while (f1 < flagged.size()) {
            const uint64_t b = track_scratch_bytes(ip.meta[flagged[f1]]);
            if (f1 > f0 && bytes + b > budget) break;
            bytes += b;
            f1++;
        }
        BatchCtx& x = e->ctx[0];
        x.ids.assign(flagged.begin() + f0, flagged.begin() + f1);
        if (hist_out) { hist_tmp.resize((f1 - f0) * RG_HIST_BINS); x.hist_host = hist_tmp.data(); }
        rc = decode_enqueue(e, c, ip, x);
        if (rc == RG_OK) rc = stage_b_enqueue(e, x, true, hist_out != nullptr, d_album_hist, d_album_peak, false, true);
        if (rc != RG_OK) return fail(rc);
        x.busy = true;
        rc = collect(x, true);
        if (rc != RG_OK) return fail(rc);
        e->stats.exact_reruns += f1 - f0;
        f0 = f1;
    }
    CK(cudaEventRecord(e->ev[0], e->stream));
    CK(cudaEventSynchronize(e->ev[0]));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev[7], e->ev[0]) == cudaSuccess) e->stats.last_ms_total = ms;
    return RG_OK;
}

extern "C" int rg_analyze_corpus(rg_engine* e, rg_corpus* c, rg_track_result* out, uint32_t* hist_out) {
    if (!e || !c || (!out && !c->len.empty())) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    return analyze_corpus_impl(e, c, out, hist_out, nullptr, nullptr, false, nullptr);
}

// A corpus whose bytes are still in the caller's (ideally pinned) buffers: the copy into the engine's own device buffer is
// part of the pipeline, group by group, so decoding starts while later files are still crossing the bus.
static int host_corpus(rg_engine* e, const rg_input* in, size_t n, rg_corpus& c) {
    CK(cudaSetDevice(e->device));
    int rc = corpus_describe(&c, in, n, e->arena);
    if (rc != RG_OK) return rc;
    CK(e->d_corpus.ensure(c.total + 64));
    c.d_bytes = e->d_corpus.as<uint8_t>();
    c.pending_in = in;
    return RG_OK;
}

extern "C" int rg_analyze_tracks(rg_engine* e, const rg_input* in, size_t n, rg_track_result* out, uint32_t* hist_out) {
    if (!e || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    rg_corpus c;
    int rc = host_corpus(e, in, n, c);
    if (rc != RG_OK) return rc;
    return analyze_corpus_impl(e, &c, out, hist_out, nullptr, nullptr, false, nullptr);
}

static int album_buffers(rg_engine* e) {
    CK(e->d_album.ensure(RG_HIST_BINS * sizeof(unsigned long long) + 64));
    return RG_OK;
}

extern "C" int rg_album_finish(rg_engine* e, const uint64_t* d_album_hist, const double* d_album_peak, rg_album_result* album) {
    if (!e || !d_album_hist || !d_album_peak || !album) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
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
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    int bad = -1;
    return analyze_corpus_impl(e, c, out, nullptr, reinterpret_cast<unsigned long long*>(d_album_hist),
                               reinterpret_cast<unsigned long long*>(d_album_peak), true, &bad);
}

extern "C" int rg_analyze_album(rg_engine* e, const rg_input* in, size_t n, rg_track_result* out, rg_album_result* album,
                                uint64_t* album_hist_out) {
    if (!e || !album || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    int rc = album_buffers(e);
    if (rc != RG_OK) return rc;
    unsigned long long* d_hist = e->d_album.as<unsigned long long>();
    CK(e->d_album_peak.ensure(64));
    unsigned long long* d_peak = e->d_album_peak.as<unsigned long long>();
    CK(cudaMemsetAsync(d_hist, 0, RG_HIST_BINS * sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(d_peak, 0, sizeof(unsigned long long), e->stream));
    rg_corpus c;
    rc = host_corpus(e, in, n, c);
    if (rc != RG_OK) return rc;
    int bad = -1;
    rc = analyze_corpus_impl(e, &c, out, nullptr, d_hist, d_peak, true, &bad);
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

// Bytes of device scratch one track needs in a peak-only batch (no PCM is stored)
static uint64_t track_scratch_bytes_peak(const TrackMeta& m) {
    return (uint64_t)m.n_frames * meta_gpf(m) * m.channels * (576 * 2 + 2 + 2 * K1_GAIN_ROW + K1_C1_ROW) + 4096;
}

extern "C" int rg_find_peak(rg_engine* e, const rg_input* in, size_t n, rg_peak_result* out) {
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    // find_peak_amplitude (src/replaygain.rs:1140-1249): max |sample| over all channels of the decoded stream.  The same
    // copy / index / decode pipeline as the analysis, but K2 folds |sample| into one word per track instead of storing PCM,
    // and there is no filter / histogram stage.
    if (!e || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    rg_corpus c;
    int rc = host_corpus(e, in, n, c);
    if (rc != RG_OK) return rc;
    e->stats.last_ms_index = e->stats.last_ms_huffman = e->stats.last_ms_hybrid = e->stats.last_ms_iir = e->stats.last_ms_hist = 0;
    e->stats.last_batches = 0;
    if (!n) return RG_OK;
    IndexPlan ip;
    rc = index_enqueue(e, &c, ip);
    if (rc != RG_OK) { drain_streams(e); return rc; }
    const uint64_t budget = e->pipeline ? e->scratch_bytes / 2 : e->scratch_bytes;
    auto collect = [&](BatchCtx& x) -> int {
        if (!x.busy) return RG_OK;
        CK(cudaEventSynchronize(x.ev_b[3]));
        collect_decode_timing(e, x);
        const uint32_t* bits = x.h_tout.as<uint32_t>();
        for (size_t k = 0; k < x.ids.size(); k++) {
            const uint32_t i = x.ids[k];
            float pk;
            memcpy(&pk, &bits[k], sizeof pk);
            out[i].status = x.pt[k].status;
            out[i].sample_rate = x.pt[k].sample_rate;
            out[i].peak = x.pt[k].status == RG_OK ? (double)pk : 0.0;
            out[i].peak_pcm = out[i].peak * 32768.0;
        }
        x.busy = false;
        return RG_OK;
    };
    size_t next = 0, k = 0;
    while (next < n) {
        rc = index_absorb(e, &c, ip, next + 1, true);
        if (rc != RG_OK) { drain_streams(e); return rc; }
        const size_t indexed_end = ip.group_begin[ip.groups_absorbed];
        size_t t1 = next;
        uint64_t bytes = 0;
        while (t1 < indexed_end) {
            const uint64_t b = ip.status[t1] == RG_OK ? track_scratch_bytes_peak(ip.meta[t1]) : 0;
            if (t1 > next && bytes + b > budget) break;
            bytes += b;
            t1++;
        }
        BatchCtx& x = e->ctx[e->pipeline ? (k & 1) : 0];
        rc = collect(x);
        if (rc != RG_OK) { drain_streams(e); return rc; }
        x.ids.resize(t1 - next);
        for (size_t i = next; i < t1; i++) x.ids[i - next] = (uint32_t)i;
        rc = decode_enqueue(e, &c, ip, x, true);
        if (rc != RG_OK) { drain_streams(e); return rc; }
        x.busy = true;
        e->stats.last_batches++;
        next = t1;
        k++;
    }
    for (size_t j = 0; j < 2; j++) {
        rc = collect(e->ctx[(k + j) & 1]);
        if (rc != RG_OK) { drain_streams(e); return rc; }
    }
    return RG_OK;
}

extern "C" int rg_decode_pcm(rg_engine* e, const rg_input* in, float* pcm, size_t max_frames, size_t* frames_out,
                             uint32_t* channels_out, uint32_t* sample_rate_out) {
    if (!e || !in || !pcm || !frames_out) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    rg_corpus c;
    int rc = host_corpus(e, in, 1, c);
    if (rc != RG_OK) return rc;
    e->stats.last_ms_index = e->stats.last_ms_huffman = e->stats.last_ms_hybrid = 0;
    IndexPlan ip;
    rc = index_enqueue(e, &c, ip);
    if (rc == RG_OK) rc = index_absorb(e, &c, ip, 1, false);
    if (rc == RG_OK && ip.status[0] != RG_OK) rc = ip.status[0];
    BatchCtx& x = e->ctx[0];
    if (rc == RG_OK) { x.ids.assign(1, 0u); rc = decode_enqueue(e, &c, ip, x); }
    if (rc == RG_OK) {
        const size_t frames = x.pt[0].frames, ncopy = std::min(frames, max_frames);
        for (int ch = 0; ch < x.pt[0].nch && rc == RG_OK; ch++)
            if (ncopy && cudaMemcpyAsync(pcm + (size_t)ch * max_frames, x.pt[0].d_ch[ch], ncopy * sizeof(float), cudaMemcpyDeviceToHost, e->s_dec) != cudaSuccess)
                rc = RG_ERR_CUDA;
        if (cudaStreamSynchronize(e->s_dec) != cudaSuccess) rc = fail_cuda(cudaGetLastError(), "rg_decode_pcm sync", __LINE__);
        collect_decode_timing(e, x);
        *frames_out = frames;
        if (channels_out) *channels_out = (uint32_t)x.pt[0].nch;
        if (sample_rate_out) *sample_rate_out = x.pt[0].sample_rate;
        e->stats.d2h_bytes += ncopy * sizeof(float) * x.pt[0].nch;
    }
    drain_streams(e);
    return rc;
}
