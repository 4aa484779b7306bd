"""GPU parity tests, stage A + B: the CUDA decode kernels (K0-K2) and the whole analyze path through the C ABI against
the CPU oracle on the same inputs.  Bar: decoded PCM equal sample for sample (float ==), identical 12000-bin histogram,
identical loudness / gain / peak."""
import json
import os

import numpy as np
import pytest

import mp3rgain_b200 as m
import oracle_lib
import synth_corpus

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXPECTED = json.load(open(os.path.join(GOLD, "fixtures_expected.json")))


def _img(name):
    return np.fromfile(os.path.join(GOLD, name + ".mp3"), np.uint8)


def _check_pcm(engine, oracle, img):
    rc, opcm, info = oracle.decode(img)
    assert rc == 0
    gpcm, sr = engine.decode_pcm(img)
    assert sr == info.sample_rate
    assert gpcm.shape == opcm.shape
    bad = np.nonzero(gpcm != opcm)
    assert len(bad[0]) == 0, (len(bad[0]), bad[0][:5], bad[1][:5], np.abs(gpcm - opcm).max())
    return info


def _check_analysis(engine, oracle, img):
    rc, ores, ohist, info = oracle.analyze_mp3(img)
    assert rc == 0
    res, sts = engine.analyze_images([img], want_hist=True)
    assert sts[0] == 0
    r = res[0]
    assert (r.histogram == ohist).all()
    assert r.loudness_db == ores.loudness_db and r.gain_db == ores.gain_db and r.peak == ores.peak
    assert r.n_windows == ores.n_windows and r.n_dropped_windows == ores.n_dropped
    assert r.n_frames == info.n_frames and r.sample_rate == info.sample_rate and r.channels == info.channels
    return r


@pytest.mark.parametrize("name", ["test_joint_stereo", "test_mono", "test_vbr", "test_stereo"])
def test_reference_fixtures(name, engine_exact, engine_fast, oracle):
    img = _img(name)
    _check_pcm(engine_exact, oracle, img)
    for eng in (engine_exact, engine_fast):
        r = _check_analysis(eng, oracle, img)
        exp = EXPECTED[name]          # independent expectation (FFmpeg PCM -> stage B): +-0.01 dB bar of the north star
        assert abs(r.loudness_db - exp["loudness_db"]) <= 0.01 + 1e-12
        assert {str(int(i)): int(r.histogram[i]) for i in np.nonzero(r.histogram)[0]} == exp["histogram"]


SYNTH = [dict(seed=3), dict(seed=4, channels=1), dict(seed=5, sr_idx=1), dict(seed=6, br_idx=0),
         dict(seed=7, sr_idx=2, br_idx=5), dict(seed=8, br_idx=14), dict(seed=9, br_idx=1),
         dict(seed=10, sr_idx=1, br_idx=0, channels=1)]


@pytest.mark.parametrize("kw", SYNTH)
def test_synthetic_streams_pcm_bit_exact(kw, engine_exact, oracle):
    img = synth_corpus.make_track(n_frames=300, **kw)
    info = _check_pcm(engine_exact, oracle, img)
    assert info.n_frames == 300


@pytest.mark.parametrize("kw", [dict(seed=21), dict(seed=22, sr_idx=1, br_idx=5), dict(seed=23, br_idx=0)])
def test_intensity_stereo_pcm_bit_exact(kw, engine_exact, oracle):
    """Streams whose joint-stereo frames all carry intensity stereo (the oracle's handling is pinned against FFmpeg in
    test_oracle_mp3.py): the device PCM equals the oracle's sample for sample."""
    img = synth_corpus.make_track(n_frames=300, intensity_prob=1.0, **kw)
    info = _check_pcm(engine_exact, oracle, img)
    assert info.n_frames == 300


LSF = [dict(version=2, seed=41), dict(version=2, seed=42, channels=1, sr_idx=1, br_idx=0), dict(version=2, seed=43, sr_idx=2, br_idx=4),
       dict(version=0, seed=44), dict(version=0, seed=45, sr_idx=1, br_idx=0), dict(version=0, seed=46, sr_idx=2, br_idx=2, channels=1)]


@pytest.mark.parametrize("kw", LSF)
def test_mpeg2_streams_pcm_bit_exact_and_analysis(kw, engine_exact, engine_fast, oracle):
    """MPEG-2 / 2.5 Layer III at all six low sample rates (the oracle's handling is pinned against FFmpeg in
    test_oracle_mp3.py): device PCM equals the oracle's sample for sample, and the analysis (50 ms windows of 1102 / 1200 /
    800 / 551 / 600 / 400 samples, the low-rate filter coefficients of src/replaygain.rs:320-525) gives identical histograms
    in exact and time-blocked mode."""
    img = synth_corpus.make_track(n_frames=500, **kw)
    info = _check_pcm(engine_exact, oracle, img)
    assert info.n_frames == 500 and info.sample_rate == synth_corpus.RATES_BY_VERSION[kw["version"]][kw.get("sr_idx", 0)]
    _check_analysis(engine_exact, oracle, img)
    _check_analysis(engine_fast, oracle, img)


@pytest.mark.parametrize("kw", [dict(version=v, sr_idx=sr, br_idx=6 if v == 2 else 2, seed=77 + sr + 3 * v, channels=c)
                                for v, sr, c in [(2, 0, 2), (2, 1, 1), (2, 2, 2), (0, 0, 2), (0, 1, 2), (0, 2, 2), (0, 2, 1)]])
def test_lsf_mixed_blocks(kw, engine_exact, oracle):
    """Mixed blocks at every MPEG-2 / 2.5 rate (6 long bands, then short band 3 onwards; at 8 kHz the long part is 72 lines):
    the band-code / reorder tables of K2 against the oracle, bit-exact PCM and histogram."""
    img, n = synth_corpus.set_mixed_block_flags(synth_corpus.make_track(n_frames=160, **kw))
    assert n >= 2
    _check_pcm(engine_exact, oracle, img)
    _check_analysis(engine_exact, oracle, img)


@pytest.mark.parametrize("kw", [dict(version=2, seed=47, channels=1, br_idx=2), dict(version=2, seed=48, channels=1, br_idx=1, sr_idx=1),
                                dict(version=0, seed=49, channels=1, br_idx=1, sr_idx=2), dict(version=2, seed=50, channels=2, br_idx=3)])
def test_low_bitrate_lsf_streams_with_frames_below_96_bytes(kw, engine_exact, engine_fast, oracle):
    """8-24 kbps MPEG-2 / 2.5: frames of 52, 24, 72 and 78 bytes, several 64 KiB index segments each.  The frame-record
    capacities used to assume the MPEG-1 minimum of 96 bytes per frame (len/96 records per track, 64Ki/96 staged frames per
    segment); these streams hold up to four times as many."""
    img = synth_corpus.make_track(n_frames=7000, **kw)
    assert len(img) / 7000 < 96 and len(img) > 2 * 65536
    info = _check_pcm(engine_exact, oracle, img)
    assert info.n_frames == 7000
    _check_analysis(engine_exact, oracle, img)
    _check_analysis(engine_fast, oracle, img)
    # and next to other tracks in one call (the records of one track must not run into the next one's)
    other = synth_corpus.make_track(seed=3, n_frames=200)
    res, sts = engine_fast.analyze_images([img, other, img], want_hist=True)
    rc, ores, ohist, _ = oracle.analyze_mp3(other)
    assert sts == [0, 0, 0] and (res[1].histogram == ohist).all() and (res[0].histogram == res[2].histogram).all()


@pytest.mark.parametrize("kw", [dict(version=2, seed=51), dict(version=2, seed=52, sr_idx=1, br_idx=0), dict(version=0, seed=53, sr_idx=1),
                                dict(version=0, seed=55, sr_idx=2, br_idx=0)])
def test_mpeg2_intensity_stereo_pcm_bit_exact(kw, engine_exact, oracle):
    """MPEG-2 / 2.5 streams whose joint-stereo frames all carry intensity stereo: device PCM == oracle PCM."""
    img = synth_corpus.make_track(n_frames=400, intensity_prob=1.0, **kw)
    info = _check_pcm(engine_exact, oracle, img)
    assert info.n_frames == 400


@pytest.mark.parametrize("kw", SYNTH[:4])
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_synthetic_streams_analysis(kw, mode, engine_exact, engine_fast, oracle):
    img = synth_corpus.make_track(n_frames=700, **kw)
    _check_analysis(engine_exact if mode == "exact" else engine_fast, oracle, img)


def _frame_starts(img):
    """Byte offsets of the frames of a clean synthetic stream (header walk)."""
    br = synth_corpus.BITRATES
    pos, out, b = 0, [], bytes(img)
    while pos + 4 <= len(b):
        bi, si, pad = b[pos + 2] >> 4, (b[pos + 2] >> 2) & 3, (b[pos + 2] >> 1) & 1
        out.append(pos)
        pos += 144 * br[bi] * 1000 // synth_corpus.RATES[si] + pad
    return out


def test_segmented_frame_index_equals_serial_walk(engine_exact, oracle):
    """K0 walks 64 KiB segments speculatively and stitches them (k_frame_index.cuh).  These streams span many segments and
    put the hard cases AT the segment boundaries: a tag larger than a segment, junk across a boundary, and -- the case the
    stitching exists for -- bytes inside a frame's payload just after a boundary that look like two consecutive valid
    headers, so that the segment's own scan locks onto a false sync and must be re-walked from the true position."""
    SEG = 64 * 1024
    base = synth_corpus.make_track(seed=31, n_frames=1500)           # ~ 627 kB = 10 segments
    assert len(base) > 9 * SEG
    starts = _frame_starts(base)
    hdr = bytes(base[:4])
    flen = starts[1] - starts[0]
    # (a) a 150 kB ID3v2 tag in front: the audio starts in the third segment
    tagged = np.frombuffer(synth_corpus.id3v2_tag(150_000) + bytes(base), np.uint8)
    # (b) junk that straddles a segment boundary (sync lost before the boundary, regained after it)
    cut = max(p for p in starts if p < 3 * SEG)
    nxt = min(p for p in starts if p >= 3 * SEG + 700)
    junk = np.concatenate([base[:cut], np.full(1234, 0x00, np.uint8), base[nxt:]])
    # (c) false syncs: right after the boundaries of segments 2..8, inside whatever frame's payload lies there, plant a
    # valid header followed, one (fake) frame length later, by another one
    fake = base.copy()
    planted = 0
    for k in range(2, 9):
        b0 = k * SEG
        f0 = max(p for p in starts if p <= b0)              # the frame that contains the boundary
        p1 = b0 + 40                                         # well inside its payload? (keep header + side info intact)
        if p1 < f0 + 40 or p1 + flen + 4 > len(fake):
            continue
        fake[p1:p1 + 4] = np.frombuffer(hdr, np.uint8)
        fake[p1 + flen:p1 + flen + 4] = np.frombuffer(hdr, np.uint8)   # may land in the next frame: also payload or header-safe
        planted += 1
    assert planted >= 5
    rewalks = {}
    for name, img in (("base", base), ("tagged", tagged), ("junk", junk), ("fake", fake)):
        rc, ores, ohist, info = oracle.analyze_mp3(img)
        assert rc == 0
        res, sts = engine_exact.analyze_images([img], want_hist=True)
        assert sts[0] == 0
        assert res[0].n_frames == info.n_frames, (name, res[0].n_frames, info.n_frames)
        assert (res[0].histogram == ohist).all() and res[0].peak == ores.peak, name
        rewalks[name] = res[0].index_rewalks
    # the clean stream stitches without a single re-walk; every planted false sync costs (at least) one
    assert rewalks["base"] == 0 and rewalks["fake"] >= planted, rewalks


def test_container_edge_cases(engine_exact, oracle):
    base = synth_corpus.make_track(seed=11, n_frames=50)
    variants = [
        np.frombuffer(synth_corpus.id3v2_tag(300) + synth_corpus.info_frame(bytes(base[:4])) + bytes(base) +
                      synth_corpus.id3v1_trailer(), np.uint8),
        base[:-7],
        np.concatenate([base, np.full(333, 0x55, np.uint8), base]),
        np.concatenate([np.full(77, 0xFF, np.uint8), base]),
        base[:500],
    ]
    grp = synth_corpus.make_group(5, 8)
    # cut in the middle of the bit reservoir: leading frames are dropped as DecodeError
    pos, cuts = 0, []
    b = bytes(grp)
    while pos + 4 <= len(b):
        bi, si, pad = b[pos + 2] >> 4, (b[pos + 2] >> 2) & 3, (b[pos + 2] >> 1) & 1
        cuts.append(pos)
        pos += 144 * synth_corpus.BITRATES[bi] * 1000 // synth_corpus.RATES[si] + pad
    variants.append(grp[cuts[3]:])
    for img in variants:
        _check_pcm(engine_exact, oracle, img)
        _check_analysis(engine_exact, oracle, img)


def test_keep_info_frame_flag(oracle):
    import mp3rgain_b200 as m
    img = _img("test_joint_stereo")
    rc, opcm, info = oracle.decode(img, flags=0x2)
    with m.Engine(device=0, exact=True, flags=0x2) as eng:
        gpcm, sr = eng.decode_pcm(img)
    assert info.n_frames == 41 and gpcm.shape == opcm.shape and (gpcm == opcm).all()


@pytest.mark.parametrize("kw", [dict(seed=61), dict(seed=62, channels=1, br_idx=0), dict(seed=63, version=2, sr_idx=1)])
def test_crc_protected_frames(kw, engine_exact, oracle):
    img = synth_corpus.make_track(n_frames=200, crc=True, **kw)
    info = _check_pcm(engine_exact, oracle, img)
    assert info.n_frames == 200
    _check_analysis(engine_exact, oracle, img)


def test_clamp_pcm_flag(oracle):
    """RG_FLAG_CLAMP_PCM (the symphonia-behaviour switch of DESIGN.md 4): a loud stream whose samples exceed +-1.0 is clamped
    identically by the oracle and by K2, in the PCM and therefore in peak and histogram."""
    import mp3rgain_b200 as m
    img = synth_corpus.make_track(seed=77, n_frames=200, level_db=-1.0)
    rc, opcm, info = oracle.decode(img)
    assert np.abs(opcm).max() > 1.0                       # the unclamped decode really overshoots
    rc, cpcm, info = oracle.decode(img, flags=0x4)
    assert np.abs(cpcm).max() == 1.0
    with m.Engine(device=0, exact=True, flags=0x4) as eng:
        gpcm, sr = eng.decode_pcm(img)
        res, sts = eng.analyze_images([img], want_hist=True)
    assert gpcm.shape == cpcm.shape and (gpcm == cpcm).all()
    rc, ores, ohist, info = oracle.analyze_mp3(img, flags=0x4)
    assert sts == [0] and (res[0].histogram == ohist).all() and res[0].peak == ores.peak == 1.0


def test_pipeline_off_gives_identical_results(oracle):
    """The "pipeline" option only changes scheduling (one batch at a time, every kernel alone): results are identical."""
    import mp3rgain_b200 as m
    imgs = [synth_corpus.make_track(seed=500 + i, n_frames=150 + 20 * i, channels=1 + i % 2) for i in range(6)]
    outs = []
    for pipe in (1, 0):
        with m.Engine(device=0, scratch_bytes=16 << 20) as eng:
            eng.set_option("pipeline", pipe)
            res, sts = eng.analyze_images(imgs, want_hist=True)
            outs.append([(r.loudness_db, r.peak, r.n_frames, r.histogram.tobytes()) for r in res])
    assert outs[0] == outs[1]


def test_errors_match_reference_messages(engine_exact):
    import mp3rgain_b200 as m
    res, sts = engine_exact.analyze_images([np.zeros(5000, np.uint8), np.zeros(0, np.uint8), _img("test_mono")])
    assert sts == [-5, -5, 0]                      # "Failed to probe format", twice; a bad entry does not poison the batch
    res, sts = engine_exact.analyze_images([_img("test_mono")], track_index=3)
    assert sts == [-7]                             # "Track index 3 out of range (file has 1 audio track(s))"
    with pytest.raises(m.ReplayGainError) as ei:
        engine_exact.analyze_images([_img("test_mono")], track_index=3, raise_on_error=True)
    assert "Track index 3 out of range (file has 1 audio track(s))" in str(ei.value)
    mp4 = np.frombuffer(bytes([0, 0, 0, 24]) + b"ftypM4A " + bytes(3000), np.uint8)
    res, sts = engine_exact.analyze_images([mp4])
    assert sts == [-10] and res[0].file_type == m.AudioFileType.Aac   # AAC/M4A containers are recognised, not decoded


def test_module_level_api_on_files(oracle, tmp_path):
    """The reference-shaped entry points: analyze_track / analyze_album / find_peak_amplitude on paths."""
    import mp3rgain_b200 as m
    paths = [os.path.join(GOLD, n + ".mp3") for n in ("test_joint_stereo", "test_mono", "test_vbr")]
    r = m.analyze_track(paths[0])
    assert r.loudness_db == 58.11 and abs(r.gain_db - 6.71) < 1e-9 and r.gain_steps() == 4 and r.sample_rate == 44100
    assert r.file_type == m.AudioFileType.Mp3
    alb = m.analyze_album(paths)
    hists, peaks = [], []
    for p in paths:
        rc, ores, ohist, info = oracle.analyze_mp3(np.fromfile(p, np.uint8))
        hists.append(ohist)
        peaks.append(ores.peak)
    merged, loud, gain, peak = oracle.album_merge(np.stack(hists), peaks)
    assert alb.album_loudness_db == loud and alb.album_gain_db == gain and alb.album_peak == peak
    assert [t.loudness_db for t in alb.tracks] == [58.11, 61.12, 58.56]
    pk = m.find_peak_amplitude(paths[1])
    assert pk.peak == peaks[1] and pk.peak_pcm == peaks[1] * 32768.0 and pk.sample_rate == 44100
    with pytest.raises(m.ReplayGainError) as ei:
        m.analyze_track(str(tmp_path / "missing.mp3"))
    assert str(ei.value).startswith("Failed to open: ")
    bad = tmp_path / "bad.mp3"
    bad.write_bytes(bytes(4000))
    with pytest.raises(m.ReplayGainError) as ei:
        m.analyze_album([paths[0], str(bad), paths[1]])      # first failing file aborts the album (src/replaygain.rs:1055)
    assert str(ei.value).startswith("Failed to probe format")


def test_album_histogram_is_sum_of_track_histograms(engine_fast, oracle):
    imgs = [synth_corpus.make_track(seed=20 + i, n_frames=200 + 37 * i, channels=1 + (i % 2)) for i in range(5)]
    alb = engine_fast.analyze_album_images(imgs, want_hist=True)
    hists, peaks = [], []
    for im in imgs:
        rc, ores, ohist, info = oracle.analyze_mp3(im)
        hists.append(ohist)
        peaks.append(ores.peak)
    merged, loud, gain, peak = oracle.album_merge(np.stack(hists), peaks)
    assert (alb.album_histogram == merged.astype(np.uint64)).all()
    assert alb.album_loudness_db == loud and alb.album_peak == peak


def test_batch_of_many_tracks_and_small_scratch(oracle):
    """Many tracks in one call, forced through several scratch batches, results independent of batching."""
    import mp3rgain_b200 as m
    pool = synth_corpus.GroupPool(seed=2, n_groups=24, group_len=16)
    imgs = [pool.track(100 + i, 160 + 16 * (i % 5)) for i in range(40)]
    with m.Engine(device=0, exact=False, scratch_bytes=8 << 20) as eng:
        res, sts = eng.analyze_images(imgs, want_hist=True)
    assert all(s == 0 for s in sts)
    for i in (0, 7, 19, 39):
        rc, ores, ohist, info = oracle.analyze_mp3(imgs[i])
        assert (res[i].histogram == ohist).all() and res[i].loudness_db == ores.loudness_db and res[i].peak == ores.peak


def _arena(imgs, gaps):
    """The images as slices of ONE host buffer, image k starting `gaps[k % len(gaps)]` + alignment padding after its predecessor."""
    pos, offs = 0, []
    for k, im in enumerate(imgs):
        pos = (pos + 15) & ~15
        if k:
            pos += gaps[k % len(gaps)]
        offs.append(pos)
        pos += len(im)
    buf = np.full(pos + 64, 0xA5, np.uint8)
    for o, im in zip(offs, imgs):
        buf[o:o + len(im)] = im
    return buf, [buf[o:o + len(im)] for o, im in zip(offs, imgs)]


def test_host_buffers_that_form_an_arena_cross_the_bus_group_by_group(oracle):
    """Callers that keep their files in one run of host memory (<= 64 bytes apart, 16-byte multiples from the first) get one
    copy per group instead of one per file and the caller's spacing on the device (corpus_describe).  Results -- analysis,
    Mp3Analysis fields, patched images -- are those of separately allocated buffers; an empty file and a file that is not an
    MP3 sit in the middle; a layout that does not qualify (a 96-byte gap) takes the per-file path."""
    import mp3rgain_b200 as m
    pool = synth_corpus.GroupPool(seed=5, n_groups=24, group_len=16)
    imgs = [pool.track(700 + i, 64 + 16 * (i % 4)) for i in range(14)]
    imgs[4] = np.zeros(0, np.uint8)
    imgs[9] = np.full(5000, 0x11, np.uint8)
    imgs[11] = synth_corpus.make_track(seed=77, n_frames=80, version=2, channels=1, br_idx=6, sr_idx=0)
    with m.Engine(device=0, exact=False, scratch_bytes=16 << 20) as eng:
        eng.set_option("group_bytes_min", 60_000)
        want, wst = eng.analyze_images(imgs, want_hist=True)
        wgain = eng.gain_analyze_images(imgs)
        wpatched, wn = eng.gain_apply_images(imgs, 3)
        for gaps in ((0, 16, 32, 48), (0,), (48, 0, 96)):
            buf, views = _arena(imgs, gaps)
            got, gst = eng.analyze_images(views, want_hist=True)
            assert gst == wst
            for a, b, st in zip(got, want, wst):
                if st == 0:
                    assert (a.histogram == b.histogram).all() and a.loudness_db == b.loudness_db and a.peak == b.peak and a.n_frames == b.n_frames
            ggain = eng.gain_analyze_images(views)
            assert [(g.status, g.frame_count, g.gain_sum, g.min_gain, g.max_gain) for g in ggain] == [
                (g.status, g.frame_count, g.gain_sum, g.min_gain, g.max_gain) for g in wgain]
            gpatched, gn = eng.gain_apply_images(views, 3)
            assert list(gn) == list(wn) and all((a == b).all() for a, b in zip(gpatched, wpatched))
            assert (buf[-64:] == 0xA5).all()     # the caller's memory is only read
    rc, ores, ohist, info = oracle.analyze_mp3(imgs[11])
    assert rc == 0 and (want[11].histogram == ohist).all()


def test_mixed_corpus_through_the_pipeline(oracle):
    """BASELINE configs[3] in miniature: VBR 32-320 kbps, 44.1 / 48 / 32 kHz, mono and stereo, different lengths, plus
    files that fail (garbage, MPEG-2, empty) in the middle of the list -- all in ONE rg_analyze_tracks call that is forced
    through many copy/index groups and many scratch batches.  Every per-file result equals the oracle's and a failing file
    does not disturb its neighbours (track mode keeps going, src/main.rs:1988-1999)."""
    import mp3rgain_b200 as m
    imgs = []
    for i in range(36):
        imgs.append(synth_corpus.make_track(seed=300 + i, n_frames=90 + 41 * (i % 7), sr_idx=i % 3, br_idx=(0 if i % 4 == 0 else 3 + i % 11),
                                            channels=1 + (i % 5 != 0)))
    lsf = np.frombuffer((bytes([0xFF, 0xF3, 0x80, 0x00]) + bytes(72 * 64000 // 22050 - 4)) * 4, np.uint8)
    imgs[5] = np.zeros(4000, np.uint8)            # not an MP3: "Failed to probe format"
    imgs[17] = lsf                                 # MPEG-2 Layer III, four empty frames: decodes to digital silence
    imgs[29] = np.zeros(0, np.uint8)              # empty file
    for exact in (False, True):
        with m.Engine(device=0, exact=exact, scratch_bytes=24 << 20) as eng:
            eng.set_option("group_bytes_min", 100_000)
            res, sts = eng.analyze_images(imgs, want_hist=True)
            assert eng.stats().last_batches >= 4
        for i, im in enumerate(imgs):
            rc, ores, ohist, info = oracle.analyze_mp3(im)
            assert sts[i] == rc, (i, sts[i], rc)
            if rc == 0:
                assert (res[i].histogram == ohist).all(), i
                assert res[i].loudness_db == ores.loudness_db and res[i].peak == ores.peak and res[i].n_frames == info.n_frames, i
                assert res[i].sample_rate == info.sample_rate and res[i].channels == info.channels, i


def test_peak_only_path_equals_the_analysis_peak(oracle):
    """rg_find_peak runs the decode kernels with K2 in peak mode (no PCM stored, no filter / histogram stage): per file the
    peak equals the oracle's max |sample| (and so the analysis path's), failing files report their status, and the call
    goes through several batches."""
    import mp3rgain_b200 as m
    imgs = [synth_corpus.make_track(seed=400 + i, n_frames=120 + 30 * (i % 4), sr_idx=i % 3, channels=1 + (i % 3 != 0)) for i in range(14)]
    imgs[6] = np.zeros(3000, np.uint8)
    with m.Engine(device=0, scratch_bytes=6 << 20) as eng:
        eng.set_option("group_bytes_min", 150_000)
        peaks = eng.find_peak_images(imgs)
        assert eng.stats().last_batches >= 3
    for i, im in enumerate(imgs):
        rc, ores, ohist, info = oracle.analyze_mp3(im)
        pk, st = peaks[i]
        assert st == rc, (i, st, rc)
        if rc == 0:
            assert pk.peak == ores.peak and pk.peak_pcm == ores.peak * 32768.0 and pk.sample_rate == info.sample_rate, i


def test_full_length_tracks_time_blocked_equals_exact():
    """BASELINE-sized tracks (3 minutes, 6891 frames, 3612 windows each): the default time-blocked IIR (with its edge-window
    re-run) and the bit-exact serial IIR give identical per-track histograms, loudness and peaks on every track, and the album
    histogram equals the sum of the track histograms -- the size-independent properties of the full-size workload."""
    import mp3rgain_b200 as m
    pool = synth_corpus.GroupPool(seed=9, n_groups=96, group_len=32, intensity_prob=0.05)
    imgs = [pool.track(7000 + i, 6891) for i in range(48)]
    with m.Engine(device=0, exact=False) as fast, m.Engine(device=0, exact=True) as exact:
        rf, sf = fast.analyze_images(imgs, want_hist=True)
        re_, se = exact.analyze_images(imgs, want_hist=True)
        alb = fast.analyze_album_images(imgs, want_hist=True)
    assert sf == se == [0] * len(imgs)
    total = np.zeros(12000, np.uint64)
    for a, b in zip(rf, re_):
        assert a.n_windows == b.n_windows == 3612 and a.n_frames == 6912   # the pool deals whole 32-frame groups
        assert (a.histogram == b.histogram).all() and a.loudness_db == b.loudness_db and a.peak == b.peak
        total += a.histogram.astype(np.uint64)
    assert (alb.album_histogram == total).all() and int(total.sum()) == 48 * 3612
    assert alb.album_peak == max(r.peak for r in rf)


def test_differential_fuzz_against_the_oracle():
    """tools/fuzz_parity.py: seeded corruptions (bit flips, header / side-info flips, truncation, inserted junk, removed
    chunks) of valid streams of every MPEG version; status, frame count, histogram, peak and PCM must equal the oracle's.
    (3,000 cases over seeds 1-7 ran clean on B200 in round 1; the suite keeps a short run.)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "90", "11"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "90 cases, 0 mismatches" in p.stdout


def test_edge_flagged_tracks_are_decoded_again_with_the_exact_filter(oracle):
    """The time-blocked IIR flags a track whose window lands within edge_tol of a bin boundary and the call decodes it again
    with the serial filter (analyze_corpus_impl's `flagged` loop).  Real corpora flag ~1e-5 of their tracks, so neither the
    bench nor any other test ever entered that loop: a tolerance of 1 flags every window, which must turn the time-blocked
    engine into the exact one -- per-track histograms, album histogram (the flagged tracks' album adds are deferred to the
    second pass and must be counted exactly once) and statistics."""
    imgs = [synth_corpus.make_track(seed=300 + i, n_frames=70 + 9 * i, **kw) for i, kw in enumerate(
        [dict(), dict(channels=1, br_idx=5), dict(version=2, sr_idx=1), dict(sr_idx=1), dict(version=0, sr_idx=2, br_idx=2)])]
    want = [oracle.analyze_mp3(im) for im in imgs]
    with m.Engine(device=0, exact=False) as eng:
        eng.set_option("edge_tol", 1.0)
        eng.set_option("group_bytes_min", 8192)
        before = eng.stats().exact_reruns
        res, sts = eng.analyze_images(imgs, want_hist=True)
        assert sts == [0] * len(imgs)
        assert eng.stats().exact_reruns - before == len(imgs)
        for r, (rc, ores, ohist, info) in zip(res, want):
            assert r.exact_rerun == 1 and (r.histogram == ohist).all() and r.loudness_db == ores.loudness_db and r.peak == ores.peak
        # album mode: the sum of the track histograms, every flagged track counted once
        album = eng.analyze_album_images(imgs, want_hist=True)
        total = sum(w[2].astype(np.uint64) for w in want)
        assert (album.album_histogram == total).all()
        assert album.album_peak == max(w[1].peak for w in want)
        # a resident corpus takes the same path
        with eng.upload(imgs) as cor:
            res2, sts2 = cor.analyze(want_hist=True)
        assert sts2 == sts and all((a.histogram == b.histogram).all() for a, b in zip(res, res2))
        # and with the tolerance back to its default nothing is flagged on this corpus, with the same results
        eng.set_option("edge_tol", 1e-9)
        res3, _ = eng.analyze_images(imgs, want_hist=True)
        assert all(r.exact_rerun == 0 for r in res3) and all((a.histogram == b.histogram).all() for a, b in zip(res, res3))


def test_count1_region_that_fills_its_row_to_the_last_byte(engine_exact, oracle):
    """K1 hands the count1 quads to K2 as one byte each, sixteen to a 16-byte store, in a 144-byte row per granule-channel.  A
    granule-channel with big_values = 0 and all 144 quads present fills its row exactly -- and round 2's first version then
    stored a tenth, empty group into the NEXT granule-channel's row, zeroing its first sixteen quads (found by the differential
    fuzz on a 320 kbps stream with sync-like junk spliced in: PCM off by 7e-15, histogram unchanged).  The image is that case."""
    img = np.fromfile(os.path.join(GOLD, "count1_fills_its_row.mp3"), np.uint8)
    info = _check_pcm(engine_exact, oracle, img)
    assert info.sample_rate == 32000 and info.n_frames == 72
    _check_analysis(engine_exact, oracle, img)


def test_a_realistic_share_of_flagged_tracks_is_rerun_beside_unflagged_ones(oracle):
    """The same loop at a flag rate between none and all: with the tolerance at 5e-4 a track of ~200 windows is flagged with
    probability ~0.2, so one call holds tracks that keep their time-blocked result and tracks that are decoded again with the
    serial filter (at the default 1e-9 that mix needs ~1e5 tracks).  Every track must equal the oracle either way, in track
    mode and in the album sum."""
    imgs = [synth_corpus.make_track(seed=700 + i, n_frames=380 + 7 * (i % 5), channels=1 + (i % 2)) for i in range(40)]
    want = [oracle.analyze_mp3(im) for im in imgs]
    with m.Engine(device=0, exact=False) as eng:
        eng.set_option("edge_tol", 5e-4)
        res, sts = eng.analyze_images(imgs, want_hist=True)
        n_rerun = sum(r.exact_rerun for r in res)
        assert sts == [0] * len(imgs) and 0 < n_rerun < len(imgs), n_rerun
        for r, (rc, ores, ohist, info) in zip(res, want):
            assert (r.histogram == ohist).all() and r.loudness_db == ores.loudness_db and r.peak == ores.peak
            assert r.exact_rerun == (1 if r.n_edge_windows else 0) or r.exact_rerun == 1   # a re-run track reports the serial pass's own count
        album = eng.analyze_album_images(imgs, want_hist=True)
        assert (album.album_histogram == sum(w[2].astype(np.uint64) for w in want)).all()


def test_slim_iir_launch_shape_gives_the_same_results(oracle):
    """k4_slim (k_iir_fast as waves of 128-thread CTAs, an option kept from the co-scheduling experiments): same chains, same
    arithmetic, other launch shape."""
    imgs = [synth_corpus.make_track(seed=320 + i, n_frames=90) for i in range(3)]
    with m.Engine(device=0, exact=False) as eng:
        ref, sts = eng.analyze_images(imgs, want_hist=True)
        for slim in (1, 2):
            eng.set_option("k4_slim", slim)
            res, sts2 = eng.analyze_images(imgs, want_hist=True)
            assert sts2 == sts == [0, 0, 0]
            assert all((a.histogram == b.histogram).all() and a.peak == b.peak for a, b in zip(ref, res))
        for r, im in zip(ref, imgs):
            assert (r.histogram == oracle.analyze_mp3(im)[2]).all()


@pytest.mark.parametrize("kw", [dict(version=2, sr_idx=0, br_idx=6, seed=90), dict(version=2, sr_idx=1, br_idx=6, seed=91),
                                dict(version=0, sr_idx=2, br_idx=2, seed=92)])
def test_lsf_intensity_partitions_of_every_range(kw, engine_exact, oracle):
    """ISO 13818-3 2.4.3.2, right channel of an intensity frame: the scalefactor partition comes from scalefac_compress >> 1
    in three ranges (< 180, < 244, <= 255).  The synthesiser only emits the first; the other two arms of K1 had never run
    (line coverage of the interpreted kernels)."""
    base = synth_corpus.make_track(n_frames=120, channels=2, intensity_prob=0.9, **kw)
    img, n = synth_corpus.set_intensity_scalefac_compress(base, [360, 401, 447, 487, 488, 495, 503, 511, 366, 490])
    assert n >= 40
    _check_pcm(engine_exact, oracle, img)
    _check_analysis(engine_exact, oracle, img)
