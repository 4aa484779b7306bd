"""ctypes wrapper of the CPU oracle (oracle/librg_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by mp3rgain_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "librg_oracle.so")
HIST = 12000


class Res(C.Structure):
    _fields_ = [("loudness_db", C.c_double), ("gain_db", C.c_double), ("peak", C.c_double),
                ("sample_rate", C.c_uint32), ("n_windows", C.c_uint32), ("n_dropped", C.c_uint32)]


class Mp3Info(C.Structure):
    _fields_ = [("status", C.c_int32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("n_frames", C.c_uint32), ("n_samples", C.c_uint64), ("skipped_info_frame", C.c_uint32),
                ("n_dropped_frames", C.c_uint32)]


class GainStats(C.Structure):
    """gfo_stats (oracle/gain_fields_oracle.h): Mp3Analysis of src/lib.rs:57-75"""
    _fields_ = [("frame_count", C.c_uint64), ("gain_sum", C.c_uint64), ("gain_count", C.c_uint64),
                ("avg_gain", C.c_double), ("headroom_db", C.c_double), ("headroom_steps", C.c_int32),
                ("mpeg_version", C.c_int32), ("channel_mode", C.c_int32), ("min_gain", C.c_uint8), ("max_gain", C.c_uint8)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.gfo_analyze.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(GainStats)]
        lib.gfo_analyze.restype = C.c_int
        lib.gfo_apply_gain.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.c_int, C.c_int]
        lib.gfo_apply_gain.restype = C.c_size_t
        lib.rgo_analyze_pcm_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_uint, C.POINTER(Res), C.c_void_p]
        lib.rgo_analyze_pcm_f32.restype = C.c_int
        lib.rgo_state_init.argtypes = [C.c_void_p, C.c_uint, C.c_int]
        lib.rgo_state_feed_mono_f64_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.rgo_state_feed_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.rgo_state_finish.argtypes = [C.c_void_p, C.POINTER(Res)]
        lib.rgo_hist_loudness_u32.argtypes = [C.c_void_p]
        lib.rgo_hist_loudness_u32.restype = C.c_double
        lib.rgo_hist_loudness_u64.argtypes = [C.c_void_p]
        lib.rgo_hist_loudness_u64.restype = C.c_double
        lib.rgo_filter_init.argtypes = [C.c_void_p, C.c_uint]
        lib.rgo_filter_init.restype = C.c_int
        lib.rgo_gain_steps.argtypes = [C.c_double]
        lib.rgo_gain_steps.restype = C.c_int
        lib.rgo_album_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self.has_mp3 = hasattr(lib, "mp3o_decode")
        if self.has_mp3:
            lib.mp3o_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(Mp3Info)]
            lib.mp3o_decode.restype = C.c_int
            lib.mp3o_analyze.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(Res), C.c_void_p, C.POINTER(Mp3Info)]
            lib.mp3o_analyze.restype = C.c_int

    # ---- stage B ----
    def analyze_pcm(self, planes, sample_rate):
        planes = [np.ascontiguousarray(p, np.float32) for p in planes]
        res = Res()
        hist = np.zeros(HIST, np.uint32)
        rc = self.lib.rgo_analyze_pcm_f32(planes[0].ctypes.data, planes[1].ctypes.data if len(planes) > 1 else None,
                                          len(planes), planes[0].size, sample_rate, C.byref(res), hist.ctypes.data)
        return rc, res, hist

    def analyze_mono_raw_f64(self, x, sample_rate):
        """The reference unit tests' path: filter.process(x) + add_mono_sample, x already in 16-bit scale."""
        x = np.ascontiguousarray(x, np.float64)
        st = C.create_string_buffer(1 << 17)
        assert self.lib.rgo_state_init(st, sample_rate, 1) == 0
        self.lib.rgo_state_feed_mono_f64_raw(st, x.ctypes.data, x.size)
        res = Res()
        self.lib.rgo_state_finish(st, C.byref(res))
        return res

    def hist_loudness(self, hist):
        hist = np.ascontiguousarray(hist)
        if hist.dtype == np.uint64:
            return self.lib.rgo_hist_loudness_u64(hist.ctypes.data)
        hist = hist.astype(np.uint32)
        return self.lib.rgo_hist_loudness_u32(hist.ctypes.data)

    def album_merge(self, hists, peaks):
        hists = np.ascontiguousarray(hists, np.uint32)
        peaks = np.ascontiguousarray(peaks, np.float64)
        out = np.zeros(HIST, np.uint32)
        l, g, p = C.c_double(), C.c_double(), C.c_double()
        self.lib.rgo_album_merge(hists.ctypes.data, peaks.ctypes.data, len(peaks), out.ctypes.data, C.byref(l), C.byref(g), C.byref(p))
        return out, l.value, g.value, p.value

    # ---- stage A ----
    def decode(self, image, flags=0):
        a = np.ascontiguousarray(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image, np.uint8)
        max_frames = max(1152, (a.size // 24 + 2) * 1152)
        pcm = np.zeros((2, max_frames), np.float32)
        info = Mp3Info()
        rc = self.lib.mp3o_decode(a.ctypes.data, a.size, flags, pcm.ctypes.data, max_frames, C.byref(info))
        n = int(info.n_samples)
        return rc, pcm[: max(info.channels, 1), :n].copy(), info

    def analyze_mp3(self, image, flags=0):
        a = np.ascontiguousarray(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image, np.uint8)
        res = Res()
        hist = np.zeros(HIST, np.uint32)
        info = Mp3Info()
        rc = self.lib.mp3o_analyze(a.ctypes.data, a.size, flags, C.byref(res), hist.ctypes.data, C.byref(info))
        return rc, res, hist, info


    # ---- the first-party frame walk over the global_gain fields (src/lib.rs) ----
    def gain_analyze(self, image):
        a = np.ascontiguousarray(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image, np.uint8)
        st = GainStats()
        rc = self.lib.gfo_analyze(a.ctypes.data, a.size, C.byref(st))
        return rc, st

    def gain_apply(self, image, steps, wrap=False, channel=-1):
        """Returns (patched copy of the image, modified frames)."""
        a = np.array(np.frombuffer(image, np.uint8) if not isinstance(image, np.ndarray) else image, np.uint8, copy=True)
        n = self.lib.gfo_apply_gain(a.ctypes.data, a.size, steps, 1 if wrap else 0, channel)
        return a, int(n)


_cached = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


def load():
    global _cached
    if _cached is None:
        if not os.path.exists(LIB):
            build()
        _cached = Oracle(C.CDLL(LIB))
    return _cached


def lcg_pcm(n, seed, shift, channels):
    """SURVEY Appendix B noise: s <- s*1664525+1013904223 mod 2^32; sample = (((s>>16)&0xFFFF)-32768)>>shift,
    interleaved L,R per sample, as f32 v/32768.  Returns a list of planar float32 arrays."""
    m = n * channels
    out = np.empty(m, np.int64)
    s = seed & 0xFFFFFFFF
    # vectorised LCG: s_k = A^k s + C (A^k-1)/(A-1); simple loop in chunks keeps it exact and fast enough
    A, Cc = 1664525, 1013904223
    for i in range(m):
        s = (s * A + Cc) & 0xFFFFFFFF
        out[i] = (((s >> 16) & 0xFFFF) - 32768) >> shift
    v = (out.astype(np.float64) / 32768.0).astype(np.float32)
    if channels == 2:
        return [np.ascontiguousarray(v[0::2]), np.ascontiguousarray(v[1::2])]
    return [v]
