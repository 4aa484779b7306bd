"""GPU parity tests for the global_gain path: mp3rgain::analyze() and apply_gain*() through the C ABI
(rg_gain_analyze / rg_gain_apply -> k_gain_fields.cuh) against oracle/gain_fields_oracle.c.  Bar: every Mp3Analysis field
equal, the patched file image byte-identical, the modified-frame count equal."""
import os
import shutil

import numpy as np
import pytest

import oracle_lib
import synth_corpus

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["test_stereo", "test_mono", "test_vbr", "test_joint_stereo"]


def _img(name):
    return np.fromfile(os.path.join(GOLD, name + ".mp3"), np.uint8)


def _b(x):
    return np.frombuffer(bytes(x), np.uint8).copy()


def _check_stats(engine, oracle, images):
    got = engine.gain_analyze_images(images)
    for img, g in zip(images, got):
        rc, o = oracle.gain_analyze(img)
        assert (g.status == 0) == (rc == 0), (g.status, rc)
        assert g.frame_count == o.frame_count
        if rc != 0:
            assert g.status == -12
            continue
        assert (g.mpeg_version, g.channel_mode, g.min_gain, g.max_gain) == (o.mpeg_version, o.channel_mode, o.min_gain, o.max_gain)
        assert (g.gain_sum, g.gain_count) == (o.gain_sum, o.gain_count)
        assert g.avg_gain == o.avg_gain and g.headroom_steps == o.headroom_steps and g.headroom_db == o.headroom_db
    return got


def _check_apply(engine, oracle, images, steps, wrap=False, channel=-1):
    outs, mod = engine.gain_apply_images(images, steps, wrap, channel)
    for img, out, m in zip(images, outs, mod):
        ref, n = oracle.gain_apply(img, steps, wrap, channel)
        assert m == n
        bad = np.nonzero(out != ref)[0]
        assert bad.size == 0, (bad.size, bad[:8])
    return outs


@pytest.mark.parametrize("name", FIXTURES)
def test_reference_fixtures(name, engine_fast, oracle):
    img = _img(name)
    _check_stats(engine_fast, oracle, [img])
    for steps in (2, -2, 200, -255, 1):
        _check_apply(engine_fast, oracle, [img], steps)
    _check_apply(engine_fast, oracle, [img], 77, wrap=True)
    _check_apply(engine_fast, oracle, [img], -300, wrap=True)
    _check_apply(engine_fast, oracle, [img], 3, channel=0)
    _check_apply(engine_fast, oracle, [img], -3, channel=1)   # on the mono fixture this patches granule 1 (src/lib.rs:721-724)


def test_zero_steps_is_a_no_op(engine_fast):   # src/lib.rs:603-605
    img = _img("test_vbr")
    outs, mod = engine_fast.gain_apply_images([img], 0)
    assert mod == [0] and (outs[0] == img).all()


STREAMS = [dict(seed=101), dict(seed=102, channels=1, br_idx=5), dict(seed=103, br_idx=0, sr_idx=1), dict(seed=104, crc=True),
           dict(seed=105, version=2), dict(seed=106, version=2, channels=1, sr_idx=1, crc=True), dict(seed=107, version=0, sr_idx=2),
           dict(seed=108, version=0, channels=1, br_idx=0), dict(seed=109, br_idx=14, sr_idx=2)]


@pytest.mark.parametrize("kw", STREAMS)
def test_synthetic_streams_every_version(kw, engine_fast, oracle):
    img = synth_corpus.make_track(n_frames=400, **kw)     # several 64 KiB segments
    _check_stats(engine_fast, oracle, [img])
    _check_apply(engine_fast, oracle, [img], 5)
    _check_apply(engine_fast, oracle, [img], -40, wrap=True)
    _check_apply(engine_fast, oracle, [img], 9, channel=1)


def test_tags_trailers_and_info_frames(engine_fast, oracle):
    t = synth_corpus.make_track(seed=120, n_frames=300)
    fake = bytes(t[:417])
    ape = fake + b"APETAGEX" + (2000).to_bytes(4, "little") + (len(fake) + 32).to_bytes(4, "little") + bytes(16)
    ape_hdr = bytes(32) + fake + b"APETAGEX" + (2000).to_bytes(4, "little") + (len(fake) + 32).to_bytes(4, "little") + bytes(4) + (1 << 31).to_bytes(4, "little") + bytes(8)
    images = [
        np.concatenate([_b(synth_corpus.id3v2_tag(70000)), _b(synth_corpus.info_frame(bytes(t[:4]))), t, _b(synth_corpus.id3v1_trailer())]),
        np.concatenate([_b(synth_corpus.info_frame(bytes(t[:4]), b"Xing")), t]),
        np.concatenate([t, _b(ape), _b(synth_corpus.id3v1_trailer())]),
        np.concatenate([t, _b(ape_hdr)]),
        np.concatenate([_b(b"ID3\x04\x00\x00\x7f\x7f\x7f\x7f"), t]),            # tag size beyond the file: nothing to walk
        np.concatenate([t[:50000], _b(synth_corpus.info_frame(bytes(t[:4]))), t[50000:]]),   # Info marker mid-stream at a non-frame position
        t[:1], t[:3], t[:4], t[:417], t[:418], t[:419], t[:420], np.zeros(0, np.uint8), np.zeros(70000, np.uint8),
        np.full(200000, 0xFF, np.uint8),
    ]
    got = _check_stats(engine_fast, oracle, images)
    assert got[0].frame_count == 300 and got[4].status == -12
    _check_apply(engine_fast, oracle, images, 4)
    _check_apply(engine_fast, oracle, images, -6, channel=0)


def test_false_syncs_after_segment_boundaries_force_rewalks(engine_fast, oracle):
    """A valid-looking header right after a 64 KiB boundary, inside the payload of the frame straddling it and confirmed
    by a sync word at its own frame length: the speculative walk of that segment locks onto it, the merge must notice
    that the true walk enters the segment later and walk the segment again."""
    t = synth_corpus.make_track(seed=130, n_frames=2000).copy()
    hdr = bytes([0xFF, 0xFB, 0x50, 0x00])   # MPEG-1 64 kbps 44.1 kHz: 208 bytes
    planted = 0
    for s in range(1, len(t) // 65536):
        p = s * 65536 + 3
        t[p:p + 4] = list(hdr)
        t[p + 208:p + 210] = [0xFF, 0xE0]
        planted += 1
    got = _check_stats(engine_fast, oracle, [t])
    assert planted >= 10 and got[0].walk_rewalks >= planted // 2
    _check_apply(engine_fast, oracle, [t], 7)


def test_patch_decisions_are_taken_on_the_original_bytes(engine_fast, oracle):
    """Found by tests/test_gain_model.py (seed 9, case 1259): a false header at the end of sub-segment 0 whose frame length
    points at gain bytes of a frame in sub-segment 1; saturating them to 255 manufactures the sync word that would confirm
    the false header.  The reference's cursor never sees patched bytes, so neither may the patch kernel (it reads the
    original image and writes a copy)."""
    img = np.load(os.path.join(GOLD, "gain_patch_false_sync_case.npy"))
    for _ in range(5):      # (the in-place version was a race: repeat)
        _check_apply(engine_fast, oracle, [img] * 8, 80, channel=0)
        _check_apply(engine_fast, oracle, [img] * 8, 80)
        _check_apply(engine_fast, oracle, [img] * 8, 255)


def test_differential_fuzz(engine_fast, oracle):
    rng = np.random.default_rng(7)
    bases = [synth_corpus.make_track(seed=140 + i, n_frames=500, version=v, channels=c, br_idx=b, sr_idx=s, crc=cr)
             for i, (v, c, b, s, cr) in enumerate([(3, 2, 9, 0, False), (3, 1, 0, 1, False), (2, 2, 6, 0, True), (0, 2, 0, 2, False)])]
    images = []
    for case in range(120):
        img = bases[case % len(bases)].copy()
        kind = case % 5
        if kind == 0:
            for _ in range(int(rng.integers(1, 200))):
                img[int(rng.integers(0, len(img)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            for _ in range(int(rng.integers(1, 60))):      # plant sync-like bytes
                p = int(rng.integers(0, len(img) - 4))
                img[p:p + 4] = [0xFF, 0xFB, int(rng.integers(0x10, 0xF0)) & 0xF4, int(rng.integers(0, 256))]
        elif kind == 2:
            img = img[: int(rng.integers(1, len(img)))]
        elif kind == 3:
            p = int(rng.integers(0, len(img)))
            junk = rng.integers(0, 256, int(rng.integers(1, 90000)), dtype=np.uint8)
            junk[rng.integers(0, len(junk), len(junk) // 50)] = 0xFF
            img = np.concatenate([img[:p], junk, img[p:]])
        else:
            p = int(rng.integers(0, len(img) - 10))
            img = np.concatenate([img[:p], img[min(len(img), p + int(rng.integers(1, 5000))):]])
        images.append(img)
    _check_stats(engine_fast, oracle, images)
    _check_apply(engine_fast, oracle, images, 3)
    _check_apply(engine_fast, oracle, images, -130, wrap=True)
    _check_apply(engine_fast, oracle, images, 2, channel=1)


def test_many_tracks_in_several_copy_groups(oracle):
    import mp3rgain_b200 as m
    pool = synth_corpus.GroupPool(seed=9, n_groups=32)
    images = [pool.track(seed=200 + i, n_frames=int(200 + 37 * (i % 11))) for i in range(96)]
    images[17] = np.zeros(0, np.uint8)
    images[40] = np.zeros(1000, np.uint8)
    with m.Engine(device=0) as eng:
        eng.set_option("group_bytes_min", 1 << 20)
        got = _check_stats(eng, oracle, images)
        assert eng.stats().last_gain_groups > 1
        assert got[17].status == -12 and got[40].status == -12
        outs = _check_apply(eng, oracle, images, 6)
        # +6 then -6 restores every field that did not saturate: the synthetic corpus stays far below 249
        back, _ = eng.gain_apply_images(outs, -6)
        for a, b in zip(back, images):
            assert (a == b).all()


def test_three_minute_track_and_analysis_consistency(engine_fast, oracle):
    """Full-size property: the walk of a 3-minute CBR-128 track (44 segments) counts the frames the decoder path decodes
    (no tags, no Info frame => same frames), and the ReplayGain result moves by exactly 1.5 dB x steps after a patch."""
    pool = synth_corpus.GroupPool(seed=3, n_groups=64)
    img = pool.track(seed=1, n_frames=6891)
    (g,) = _check_stats(engine_fast, oracle, [img])
    res, sts = engine_fast.analyze_images([img])
    assert sts == [0] and g.frame_count == res[0].n_frames
    (up,), (n,) = engine_fast.gain_apply_images([img], 4)
    assert n == g.frame_count
    res2, sts2 = engine_fast.analyze_images([up])
    assert sts2 == [0] and abs((res2[0].loudness_db - res[0].loudness_db) - 6.02) <= 0.011   # 4 steps = x2 in amplitude = +6.02 dB


def test_module_level_api_mirrors_the_reference_integration_tests(tmp_path):
    """tests/integration_tests.rs:32-200, 288-350, 416-450 through mp3rgain_b200.gain on copies of the fixtures."""
    from mp3rgain_b200 import gain

    def copy(name):
        dst = tmp_path / name
        shutil.copy(os.path.join(GOLD, name), dst)
        return dst

    info = gain.analyze(os.path.join(GOLD, "test_stereo.mp3"))
    assert info.frame_count > 0 and info.mpeg_version == "MPEG1" and info.channel_mode in ("Stereo", "Joint Stereo")
    assert info.min_gain <= info.avg_gain <= info.max_gain
    assert info.headroom_steps == 255 - info.max_gain and abs(info.headroom_db - info.headroom_steps * 1.5) < 0.01
    assert gain.analyze(os.path.join(GOLD, "test_mono.mp3")).channel_mode == "Mono"
    assert gain.analyze(os.path.join(GOLD, "test_vbr.mp3")).frame_count > 0
    with pytest.raises(gain.GainError):
        gain.analyze(tmp_path / "nonexistent.mp3")
    for name in ("test_vbr.mp3", "test_joint_stereo.mp3", "test_mono.mp3"):
        p = copy(name)
        o = gain.analyze(p)
        assert gain.apply_gain(p, 2) > 0
        a = gain.analyze(p)
        assert a.max_gain >= o.max_gain and a.min_gain >= o.min_gain
        assert gain.apply_gain(p, -2) > 0
        assert open(p, "rb").read() == open(os.path.join(GOLD, name), "rb").read()
    p = copy("test_stereo.mp3")
    before = open(p, "rb").read()
    assert gain.apply_gain(p, 0) == 0 and open(p, "rb").read() == before                    # :157-166, :431-450
    gain.apply_gain(p, 200)
    assert gain.analyze(p).max_gain == 255
    gain.apply_gain(p, -255)
    assert gain.analyze(p).min_gain == 0
    p = copy("test_joint_stereo.mp3")
    assert gain.apply_gain_channel(p, gain.Channel.Left, 2) > 0 and gain.apply_gain_channel(p, gain.Channel.Right, -2) > 0
    assert gain.apply_gain_channel(p, gain.Channel.Left, 0) == 0
    with pytest.raises(gain.GainError) as ei:
        gain.apply_gain_channel(copy("test_mono.mp3"), gain.Channel.Left, 2)
    assert "mono" in str(ei.value)
    assert gain.db_to_steps(2.25) == 2 and gain.db_to_steps(-1.5) == -1 and gain.steps_to_db(-2) == -3.0    # src/lib.rs:1345-1358
    amp, mx, mn = gain.find_max_amplitude(os.path.join(GOLD, "test_vbr.mp3"))
    assert amp > 0 and mx >= mn
