"""Synthetic MP3 corpora for tests and bench (ctypes wrapper of tests/synth/synth_mp3.cpp).  TEST/BENCH INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "synth", "synth_mp3.cpp")
LIB = os.path.join(HERE, "synth", "libsynth_mp3.so")
_lib = None

BITRATES = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
RATES = [44100, 48000, 32000]


def build():
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC], check=True, capture_output=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
            build()
        _lib = C.CDLL(LIB)
        _lib.synth_group.restype = C.c_size_t
        _lib.synth_group.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_size_t]
        _lib.synth_track.restype = C.c_size_t
        _lib.synth_track.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_size_t]
        _lib.synth_set_intensity_prob.argtypes = [C.c_double]
        _lib.synth_set_version.argtypes = [C.c_int]
        _lib.synth_set_crc.argtypes = [C.c_int]
    return _lib


RATES_BY_VERSION = {3: [44100, 48000, 32000], 2: [22050, 24000, 16000], 0: [11025, 12000, 8000]}


def make_group(seed, n_frames=32, sr_idx=0, br_idx=9, channels=2, level_db=-20.0):
    lib = _load()
    cap = n_frames * 1500 + 64
    buf = np.zeros(cap, np.uint8)
    n = lib.synth_group(seed, n_frames, sr_idx, br_idx, channels, level_db, buf.ctypes.data, cap)
    assert n > 0
    return buf[:n].copy()


def make_track(seed, n_frames=6891, group_len=32, sr_idx=0, br_idx=9, channels=2, level_db=-20.0, intensity_prob=0.25,
               version=3, crc=False):
    """One unique track (every group freshly generated). br_idx 0 = VBR. Returns a numpy uint8 array.
    intensity_prob: share of the joint-stereo frames (60 % of stereo frames) that carry intensity stereo.
    version: 3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5 (half / quarter sample rates, 576-sample frames).
    crc: frames carry the 16-bit CRC word (protection_bit = 0)."""
    lib = _load()
    lib.synth_set_intensity_prob(float(intensity_prob))
    lib.synth_set_version(int(version))
    lib.synth_set_crc(1 if crc else 0)
    try:
        return _make_track(lib, seed, n_frames, group_len, sr_idx, br_idx, channels, level_db)
    finally:
        lib.synth_set_version(3)
        lib.synth_set_crc(0)


def _make_track(lib, seed, n_frames, group_len, sr_idx, br_idx, channels, level_db):
    cap = n_frames * 1500 + 64
    buf = np.zeros(cap, np.uint8)
    n = lib.synth_track(seed, n_frames, group_len, sr_idx, br_idx, channels, level_db, buf.ctypes.data, cap)
    assert n > 0
    return buf[:n].copy()


class GroupPool:
    """A pool of self-contained frame groups; tracks are seeded random sequences of pool groups (each group starts with
    main_data_begin = 0, so any concatenation is a valid stream).  Lets bench.py build 1,000 distinct 3-minute tracks in
    seconds: every track is still a different bitstream and decodes to different PCM / filter state."""

    def __init__(self, seed=1, n_groups=256, group_len=32, sr_idx=0, br_idx=9, channels=2, level_db=-20.0,
                 intensity_prob=0.25):
        rng = np.random.default_rng(seed)
        self.group_len = group_len
        _load().synth_set_intensity_prob(float(intensity_prob))
        try:
            self.groups = [make_group(int(seed) * 7919 + g, group_len, sr_idx, br_idx, channels,
                                      level_db + float(rng.uniform(-8, 4))) for g in range(n_groups)]
        finally:
            _load().synth_set_intensity_prob(0.25)

    def track(self, seed, n_frames):
        rng = np.random.default_rng(seed)
        n_groups = (n_frames + self.group_len - 1) // self.group_len
        idx = rng.integers(0, len(self.groups), n_groups)
        parts = [self.groups[i] for i in idx]
        return np.concatenate(parts)


def id3v2_tag(size=100):
    body = bytes(size)
    sz = len(body)
    return b"ID3\x04\x00\x00" + bytes([(sz >> 21) & 0x7F, (sz >> 14) & 0x7F, (sz >> 7) & 0x7F, sz & 0x7F]) + body


def info_frame(first_frame_header, tag=b"Info"):
    """A Xing/Info frame matching the stream's version/rate/channel mode (128 kbps slot, zero side info)."""
    h = bytearray(first_frame_header[:4])
    h[2] = (9 << 4) | (h[2] & 0x0C)  # bitrate index 9, same sample rate, no padding
    sr = RATES[(h[2] >> 2) & 3]
    flen = 144 * 128000 // sr
    mono = (h[3] >> 6) == 3
    fr = bytearray(flen)
    fr[:4] = h
    o = 4 + (17 if mono else 32)
    fr[o:o + 4] = tag
    return bytes(fr)


def id3v1_trailer():
    return b"TAG" + bytes(125)


def set_mixed_block_flags(img):
    """Sets mixed_block_flag in every short-block granule-channel of an MPEG-2 / 2.5 stream (the synthesiser emits pure short
    blocks only at those versions).  The granule's scalefactor partition and band layout change with the flag, so the payload
    decodes to different -- but still well-defined -- values.  Returns (image, granule-channels changed)."""
    import ffmpeg_ref   # the pure-Python frame splitter of the harness
    out, n = [], 0
    for fb in ffmpeg_ref.split_frames(img):
        fr = bytearray(fb)
        h = int.from_bytes(fr[:4], "big")
        assert ((h >> 19) & 3) != 3, "MPEG-2 / 2.5 streams only"
        mono = ((h >> 6) & 3) == 3
        base = 4 + (0 if (h >> 16) & 1 else 2)
        for ch in range(1 if mono else 2):
            b0 = (9 if mono else 10) + 63 * ch          # this channel's 63 granule bits (main_data_begin 8 + private 1 / 2)

            def get(o, w):
                v = 0
                for k in range(w):
                    bit = b0 + o + k
                    v = (v << 1) | ((fr[base + bit // 8] >> (7 - bit % 8)) & 1)
                return v
            if get(38, 1) and get(39, 2) == 2:            # window_switching_flag, block_type == 2
                bit = b0 + 41                             # mixed_block_flag
                if not (fr[base + bit // 8] >> (7 - bit % 8)) & 1:
                    fr[base + bit // 8] |= 1 << (7 - bit % 8)
                    n += 1
        out.append(bytes(fr))
    return np.frombuffer(b"".join(out), np.uint8).copy(), n


def set_intensity_scalefac_compress(img, values):
    """Overwrites scalefac_compress of the RIGHT channel in every intensity-stereo frame of an MPEG-2 / 2.5 stream with the
    next entry of `values` (cycled): reaches the three intensity partition ranges of ISO 13818-3 2.4.3.2 (sfc >> 1 below
    180, below 244, up to 255), of which the synthesiser only emits the first.  The payload is re-interpreted, as in
    set_mixed_block_flags.  Returns (image, frames changed)."""
    import ffmpeg_ref
    out, n = [], 0
    for fb in ffmpeg_ref.split_frames(img):
        fr = bytearray(fb)
        h = int.from_bytes(fr[:4], "big")
        assert ((h >> 19) & 3) != 3, "MPEG-2 / 2.5 streams only"
        if ((h >> 6) & 3) == 1 and ((h >> 4) & 1):          # joint stereo with the intensity bit
            base = 4 + (0 if (h >> 16) & 1 else 2)
            v = int(values[n % len(values)])
            for k in range(9):
                bit = 10 + 63 + 29 + k                       # main_data_begin 8, private 2, channel 0's 63 bits, 12 + 9 + 8
                byte, m = base + bit // 8, 1 << (7 - bit % 8)
                fr[byte] = (fr[byte] | m) if (v >> (8 - k)) & 1 else (fr[byte] & ~m)
            n += 1
        out.append(bytes(fr))
    return np.frombuffer(b"".join(out), np.uint8).copy(), n
