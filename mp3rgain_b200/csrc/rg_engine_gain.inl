// rg_engine_gain.inl -- entry points over the global_gain fields (k_gain_fields.cuh): mp3rgain::analyze() and
// apply_gain*() of the reference (src/lib.rs:470-514, 544-639, 668-735, 1232-1246).  Included at the end of rg_engine.cu.
//
// Host images arrive in copy groups; group g+1 crosses the bus while the walk kernels of group g run, and (when patching)
// the patched images of group g travel back on a third stream -- the bus is full-duplex, the kernels are ~2 % of the time.

static int gain_impl(rg_engine* e, rg_corpus* c, rg_gain_stats* out, bool patch, int steps, int wrap, int channel,
                     uint8_t* const* out_images, uint32_t* frames_modified) {
    const size_t n = c->len.size();
    CK(cudaSetDevice(e->device));
    e->stats.last_ms_gain_scan = e->stats.last_ms_gain_merge = e->stats.last_ms_gain_patch = 0;
    e->stats.last_gain_groups = 0;
    if (!n) return RG_OK;
    std::vector<TrackIn> tin(n);
    std::vector<TrackSeg> tseg(n);
    uint64_t n_sub = 0, bytes = 0;
    for (size_t i = 0; i < n; i++) {
        if (c->len[i] >= 0xFFFF0000ull) return RG_ERR_INVALID_ARG;   // positions are 32-bit on the device
        tin[i].off = c->off[i]; tin[i].len = c->len[i];
        tseg[i].seg_base = (uint32_t)n_sub;
        tseg[i].n_seg = (uint32_t)std::max<uint64_t>(1, (c->len[i] + GF_SUB - 1) / GF_SUB);
        n_sub += tseg[i].n_seg;
        bytes += c->len[i];
    }
    if (n_sub >= 0xFFFFFFFFull) return RG_ERR_INVALID_ARG;          // sub-segment ids are 32-bit (16 TB of file images)
    std::vector<size_t> group_begin;
    const uint64_t gbytes = c->pending_in ? std::max<uint64_t>(bytes / 8 + 1, e->group_bytes_min) : ~0ull;
    uint64_t acc = 0;
    for (size_t i = 0; i < n; i++) {
        if (i == 0 || acc >= gbytes) { group_begin.push_back(i); acc = 0; }
        acc += c->len[i];
    }
    group_begin.push_back(n);
    const size_t ng = group_begin.size() - 1;
    CK(e->d_tin.ensure(n * sizeof(TrackIn)));
    CK(e->d_tseg.ensure(n * sizeof(TrackSeg)));
    CK(e->d_gfseg.ensure(n_sub * sizeof(GfSeg)));
    CK(e->d_gfent.ensure(n_sub * sizeof(uint32_t)));
    CK(e->d_gfout.ensure(n * sizeof(GfTrackOut)));
    CK(e->h_gfout.ensure(n * sizeof(GfTrackOut)));
    if (patch) CK(e->d_gfpatch.ensure(c->total + 64));   // the patched copy (k_gf_patch reads the original, see k_gain_fields.cuh)
    cudaStream_t st = e->stream;
    CK(cudaMemcpyAsync(e->d_tin.p, tin.data(), n * sizeof(TrackIn), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->d_tseg.p, tseg.data(), n * sizeof(TrackSeg), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // the sources are stack / heap temporaries
    e->stats.h2d_bytes += n * (sizeof(TrackIn) + sizeof(TrackSeg));
    for (size_t g = 0; g < ng; g++) {
        const size_t t0 = group_begin[g], t1 = group_begin[g + 1];
        CK(e->reserve_events(6 * g + 6));
        cudaEvent_t ev_copied = e->pool_event(6 * g), ev0 = e->pool_event(6 * g + 1), ev1 = e->pool_event(6 * g + 2),
                    ev2 = e->pool_event(6 * g + 3), ev3 = e->pool_event(6 * g + 4);
        if (c->pending_in) {
            if (int r = copy_images(e, c, c->pending_in, t0, t1, e->s_copy)) return r;
            CK(cudaEventRecord(ev_copied, e->s_copy));
            CK(cudaStreamWaitEvent(st, ev_copied, 0));
        }
        const uint32_t seg0 = tseg[t0].seg_base;
        const uint32_t nsg = (t1 < n ? tseg[t1].seg_base : (uint32_t)n_sub) - seg0;
        const unsigned gs = (nsg + GF_THREADS - 1) / GF_THREADS, gt = (unsigned)((t1 - t0 + GF_THREADS / 32 - 1) / (GF_THREADS / 32));
        CK(cudaEventRecord(ev0, st));
        k_gf_scan<<<gs, GF_THREADS, 0, st>>>(c->d_bytes, e->d_tin.as<TrackIn>() + t0, e->d_tseg.as<TrackSeg>() + t0, (uint32_t)(t1 - t0), seg0, nsg,
                                             e->d_gfseg.as<GfSeg>());
        CK(cudaEventRecord(ev1, st));
        k_gf_merge<<<gt, GF_THREADS, 0, st>>>(c->d_bytes, e->d_tin.as<TrackIn>() + t0, e->d_tseg.as<TrackSeg>() + t0, e->d_gfseg.as<GfSeg>(),
                                                 e->d_gfent.as<uint32_t>(), e->d_gfout.as<GfTrackOut>() + t0, (uint32_t)(t1 - t0));
        CK(cudaEventRecord(ev2, st));
        e->stats.kernel_launches += 2;
        if (patch) {
            const uint64_t off0 = c->off[t0], span = (t1 < n ? c->off[t1] : c->total) - off0;
            CK(cudaMemcpyAsync(e->d_gfpatch.as<uint8_t>() + off0, c->d_bytes + off0, span, cudaMemcpyDeviceToDevice, st));
            k_gf_patch<<<gs, GF_THREADS, 0, st>>>(c->d_bytes, e->d_gfpatch.as<uint8_t>(), e->d_tin.as<TrackIn>() + t0, e->d_tseg.as<TrackSeg>() + t0, (uint32_t)(t1 - t0), seg0, nsg,
                                                  e->d_gfent.as<uint32_t>(), steps, wrap, channel);
            e->stats.kernel_launches++;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ev3, st));
        CK(cudaMemcpyAsync(e->h_gfout.as<GfTrackOut>() + t0, e->d_gfout.as<GfTrackOut>() + t0, (t1 - t0) * sizeof(GfTrackOut), cudaMemcpyDeviceToHost, st));
        e->stats.d2h_bytes += (t1 - t0) * sizeof(GfTrackOut);
        if (patch) {
            CK(cudaStreamWaitEvent(e->s_dec, ev3, 0));
            for (size_t i = t0; i < t1; i++) {
                if (!c->len[i]) continue;
                CK(cudaMemcpyAsync(out_images[i], e->d_gfpatch.as<uint8_t>() + c->off[i], c->len[i], cudaMemcpyDeviceToHost, e->s_dec));
                e->stats.d2h_bytes += c->len[i];
            }
        }
        e->stats.last_gain_groups++;
    }
    CK(cudaStreamSynchronize(st));
    if (patch) CK(cudaStreamSynchronize(e->s_dec));
    for (size_t g = 0; g < ng; g++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->pool_event(6 * g + 1), e->pool_event(6 * g + 2)) == cudaSuccess) e->stats.last_ms_gain_scan += ms;
        if (cudaEventElapsedTime(&ms, e->pool_event(6 * g + 2), e->pool_event(6 * g + 3)) == cudaSuccess) e->stats.last_ms_gain_merge += ms;
        if (patch && cudaEventElapsedTime(&ms, e->pool_event(6 * g + 3), e->pool_event(6 * g + 4)) == cudaSuccess) e->stats.last_ms_gain_patch += ms;
    }
    const GfTrackOut* r = e->h_gfout.as<GfTrackOut>();
    for (size_t i = 0; i < n; i++) {
        if (frames_modified) frames_modified[i] = r[i].frame_count;
        if (!out) continue;
        rg_gain_stats& o = out[i];
        memset(&o, 0, sizeof o);
        o.frame_count = r[i].frame_count;
        o.walk_rewalks = r[i].n_rewalks;
        if (!r[i].frame_count) { o.status = RG_ERR_NO_FRAMES; continue; }   // src/lib.rs:498-500
        o.status = RG_OK;
        o.gain_sum = r[i].gain_sum; o.gain_count = r[i].gain_count;
        o.avg_gain = (double)r[i].gain_sum / (double)r[i].gain_count;      // :502
        o.min_gain = (uint8_t)(r[i].gminmax & 0xFF); o.max_gain = (uint8_t)(r[i].gminmax >> 8);
        o.headroom_steps = 255 - (int32_t)o.max_gain;                      // :503
        o.headroom_db = (double)o.headroom_steps * 1.5;                     // :504, GAIN_STEP_DB
        const uint32_t ver = (r[i].first_hdr >> 19) & 3;
        o.mpeg_version = ver == 3 ? 0 : (ver == 2 ? 1 : 2);
        o.channel_mode = (int32_t)((r[i].first_hdr >> 6) & 3);
    }
    return RG_OK;
}

extern "C" int rg_gain_analyze(rg_engine* e, const rg_input* in, size_t n, rg_gain_stats* out) {
    if (!e || (n && (!in || !out))) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    rg_corpus c;
    int rc = host_corpus(e, in, n, c);
    if (rc != RG_OK) return rc;
    rc = gain_impl(e, &c, out, false, 0, 0, -1, nullptr, nullptr);
    if (rc != RG_OK) drain_streams(e);
    return rc;
}

extern "C" int rg_gain_analyze_corpus(rg_engine* e, rg_corpus* c, rg_gain_stats* out) {
    if (!e || !c || (!out && !c->len.empty())) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    int rc = gain_impl(e, c, out, false, 0, 0, -1, nullptr, nullptr);
    if (rc != RG_OK) drain_streams(e);
    return rc;
}

extern "C" int rg_gain_apply(rg_engine* e, const rg_input* in, size_t n, int32_t gain_steps, int32_t mode, int32_t channel,
                             uint8_t* const* out_images, uint32_t* frames_modified) {
    if (!e || (n && (!in || !out_images)) || channel > 1 || (mode != RG_GAIN_SATURATE && mode != RG_GAIN_WRAP)) return RG_ERR_INVALID_ARG;
    std::lock_guard<std::recursive_mutex> lock(e->mu);
    for (size_t i = 0; i < n; i++) if (in[i].len && (!in[i].data || !out_images[i])) return RG_ERR_INVALID_ARG;
    if (gain_steps == 0) {   // apply_gain returns Ok(0) before it reads the file (src/lib.rs:603-605)
        for (size_t i = 0; i < n; i++) {
            if (in[i].len && out_images[i] != in[i].data) memmove(out_images[i], in[i].data, in[i].len);
            if (frames_modified) frames_modified[i] = 0;
        }
        return RG_OK;
    }
    rg_corpus c;
    int rc = host_corpus(e, in, n, c);
    if (rc != RG_OK) return rc;
    rc = gain_impl(e, &c, nullptr, true, gain_steps, mode == RG_GAIN_WRAP ? 1 : 0, channel < 0 ? -1 : channel, out_images, frames_modified);
    if (rc != RG_OK) drain_streams(e);
    return rc;
}
