"""CPU tests pinning oracle/gain_fields_oracle.c: the reference's own unit tests (src/lib.rs:1344-1442) and integration
tests (tests/integration_tests.rs:32-80, 88-200, 288-428) run against the restatement on the reference's four fixtures."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib
import synth_corpus

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VERS = ("MPEG1", "MPEG2", "MPEG2.5")
MODES = ("Stereo", "Joint Stereo", "Dual Channel", "Mono")


@pytest.fixture(scope="module")
def orc():
    return oracle_lib.load()


def fixture(name):
    return np.fromfile(os.path.join(GOLD, name), np.uint8)


def buf(b):
    return np.frombuffer(bytes(b), np.uint8).copy()


# ---- src/lib.rs unit tests --------------------------------------------------------------------------------------------
def test_parse_valid_header(orc):   # :1361-1370
    v, m, fs = C.c_int(), C.c_int(), C.c_size_t()
    h = buf([0xFF, 0xFB, 0x90, 0x00])
    assert orc.lib.gfo_parse_header(h.ctypes.data, C.byref(v), C.byref(m), C.byref(fs)) == 1
    assert VERS[v.value] == "MPEG1" and fs.value == 1152 * 128 * 125 // 44100   # 128 kbps, 44100 Hz


def test_parse_invalid_header(orc):   # :1372-1376
    v, m, fs = C.c_int(), C.c_int(), C.c_size_t()
    for h in ([0, 0, 0, 0], [0xFF, 0xFF, 0x90, 0x00]):
        a = buf(h)
        assert orc.lib.gfo_parse_header(a.ctypes.data, C.byref(v), C.byref(m), C.byref(fs)) == 0


def test_bit_operations(orc):   # :1378-1403
    lib = orc.lib
    lib.gfo_read_gain_at.restype = C.c_uint8
    lib.gfo_read_gain_at.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint]
    lib.gfo_write_gain_at.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint8]
    d = buf([0xAB, 0xCD, 0xEF, 0x12, 0x34])
    assert lib.gfo_read_gain_at(d.ctypes.data, 5, 1, 0) == 0xCD
    assert lib.gfo_read_gain_at(d.ctypes.data, 5, 1, 4) == 0xDE
    lib.gfo_write_gain_at(d.ctypes.data, 5, 1, 0, 0x42)
    assert d[1] == 0x42
    d = buf([0xAB, 0xCD, 0xEF, 0x12, 0x34])
    lib.gfo_write_gain_at(d.ctypes.data, 5, 1, 4, 0x99)
    assert d[1] == 0xC9 and d[2] == 0x9F


def test_skip_id3v2(orc):   # :1405-1412
    orc.lib.gfo_skip_id3v2.restype = C.c_size_t
    orc.lib.gfo_skip_id3v2.argtypes = [C.c_void_p, C.c_size_t]
    a = buf([0xFF, 0xFB, 0x90, 0x00])
    assert orc.lib.gfo_skip_id3v2(a.ctypes.data, a.size) == 0
    b = buf(b"ID3\x04\x00\x00\x00\x00\x00\x00")
    assert orc.lib.gfo_skip_id3v2(b.ctypes.data, b.size) == 10


def test_is_xing_frame(orc):   # :1414-1442
    orc.lib.gfo_is_xing_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    d = np.zeros(100, np.uint8)
    d[:4] = [0xFF, 0xFB, 0x90, 0x00]
    d[36:40] = list(b"Xing")
    assert orc.lib.gfo_is_xing_frame(d.ctypes.data, d.size, 0) == 1
    d[36:40] = list(b"Info")
    assert orc.lib.gfo_is_xing_frame(d.ctypes.data, d.size, 0) == 1
    d[36:40] = 0
    assert orc.lib.gfo_is_xing_frame(d.ctypes.data, d.size, 0) == 0


def test_adjust_gain_value(orc):   # :526-541
    f = orc.lib.gfo_adjust_gain_value
    f.restype = C.c_uint8
    f.argtypes = [C.c_uint8, C.c_int32, C.c_int]
    assert f(250, 10, 0) == 255 and f(3, -10, 0) == 0 and f(100, 0, 0) == 100 and f(100, 1000, 0) == 255 and f(100, -1000, 0) == 0
    assert f(250, 10, 1) == 4 and f(3, -10, 1) == 249 and f(7, 256 * 3, 1) == 7 and f(7, -256 * 3 - 8, 1) == 255


# ---- tests/integration_tests.rs ---------------------------------------------------------------------------------------
def test_analyze_stereo_file(orc):   # :32-49
    rc, st = orc.gain_analyze(fixture("test_stereo.mp3"))
    assert rc == 0 and st.frame_count > 0 and VERS[st.mpeg_version] == "MPEG1"
    assert MODES[st.channel_mode] in ("Stereo", "Joint Stereo")
    assert st.min_gain <= st.max_gain and st.min_gain <= st.avg_gain <= st.max_gain


def test_analyze_mono_file(orc):   # :51-64
    rc, st = orc.gain_analyze(fixture("test_mono.mp3"))
    assert rc == 0 and st.frame_count > 0 and MODES[st.channel_mode] == "Mono"


def test_analyze_vbr_file(orc):   # :66-78
    rc, st = orc.gain_analyze(fixture("test_vbr.mp3"))
    assert rc == 0 and st.frame_count > 0


def test_no_frames(orc):   # src/lib.rs:498-500
    rc, st = orc.gain_analyze(np.zeros(5000, np.uint8))
    assert rc == -1 and st.frame_count == 0
    rc, st = orc.gain_analyze(np.zeros(0, np.uint8))
    assert rc == -1


@pytest.mark.parametrize("name", ["test_stereo.mp3", "test_mono.mp3", "test_vbr.mp3", "test_joint_stereo.mp3"])
def test_apply_gain_directions_and_saturation(orc, name):   # :88-200, 352-414
    img = fixture(name)
    _, o = orc.gain_analyze(img)
    up, n = orc.gain_apply(img, 2)
    assert n == o.frame_count > 0
    _, a = orc.gain_analyze(up)
    assert a.max_gain >= o.max_gain and a.min_gain >= o.min_gain and a.frame_count == o.frame_count
    dn, _ = orc.gain_apply(img, -2)
    _, b = orc.gain_analyze(dn)
    assert b.max_gain <= o.max_gain and b.min_gain <= o.min_gain
    hi, _ = orc.gain_apply(img, 200)
    assert orc.gain_analyze(hi)[1].max_gain == 255
    lo, _ = orc.gain_apply(img, -255)
    assert orc.gain_analyze(lo)[1].min_gain == 0
    # only gain fields change, never the length; +k then -k restores every unsaturated field
    assert up.size == img.size
    if o.max_gain <= 253:
        back, _ = orc.gain_apply(up, -2)
        assert (back == img).all()
    # wrap mode is a bijection
    w, _ = orc.gain_apply(img, 77, wrap=True)
    w2, _ = orc.gain_apply(w, -77, wrap=True)
    assert (w2 == img).all()


def test_headroom_calculation(orc):   # :416-428
    _, st = orc.gain_analyze(fixture("test_stereo.mp3"))
    assert st.headroom_steps == 255 - st.max_gain and abs(st.headroom_db - st.headroom_steps * 1.5) < 0.01


def test_channel_gain_touches_only_that_channel(orc):   # :288-318, src/lib.rs:668-735
    img = fixture("test_joint_stereo.mp3")
    left, n = orc.gain_apply(img, 2, channel=0)
    right, _ = orc.gain_apply(img, 2, channel=1)
    both, _ = orc.gain_apply(img, 2)
    assert n > 0 and (left != img).any() and (right != img).any()
    lr, _ = orc.gain_apply(left, 2, channel=1)
    assert (lr == both).all()


def test_trailers_and_tags(orc):   # find_audio_end :358-383, skip_id3v2 :343-354, Xing skip :446-451
    t = synth_corpus.make_track(seed=5, n_frames=40)
    _, base = orc.gain_analyze(t)
    assert base.frame_count == 40
    tagged = np.concatenate([buf(synth_corpus.id3v2_tag(300)), buf(synth_corpus.info_frame(bytes(t[:4]))), t,
                             buf(synth_corpus.id3v1_trailer())])
    _, st = orc.gain_analyze(tagged)
    assert st.frame_count == 40 and st.gain_sum == base.gain_sum
    # APEv2 footer (32 bytes, tag_size counts items + footer) directly before the ID3v1 tag: bytes inside it that look
    # like frames are not walked
    fake = bytes(t[:417])
    ape = fake + b"APETAGEX" + (2000).to_bytes(4, "little") + (len(fake) + 32).to_bytes(4, "little") + bytes(16)
    with_ape = np.concatenate([t, buf(ape), buf(synth_corpus.id3v1_trailer())])
    _, st2 = orc.gain_analyze(with_ape)
    assert st2.frame_count == 40 and st2.gain_sum == base.gain_sum
