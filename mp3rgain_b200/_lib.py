"""ctypes binding of the engine's C ABI (include/rg_engine.h).

The shared library is built in-tree by `__graft_entry__.build()` / `make -C mp3rgain_b200/csrc`.
There is no Python or CPU implementation behind this module: if the library is missing, or no CUDA
device is usable, creating an Engine raises.
"""
import ctypes as C
import os

HIST_SIZE = 12000
HIST_OFFSET = 2000
PINK_REF = 64.82

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librg_engine.so")

RG_FLAG_EXACT_IIR = 0x1
RG_FLAG_KEEP_INFO_FRAME = 0x2
RG_FLAG_CLAMP_PCM = 0x4

RG_OK = 0
RG_ERR_INVALID_ARG = -1
RG_ERR_NO_DEVICE = -2
RG_ERR_CUDA = -3
RG_ERR_OPEN = -4
RG_ERR_PROBE = -5
RG_ERR_NO_AUDIO_TRACK = -6
RG_ERR_TRACK_INDEX = -7
RG_ERR_UNKNOWN_SAMPLE_RATE = -8
RG_ERR_UNSUPPORTED_RATE = -9
RG_ERR_UNSUPPORTED_FORMAT = -10
RG_ERR_NO_MEMORY = -11
RG_ERR_NO_FRAMES = -12

RG_GAIN_SATURATE = 0
RG_GAIN_WRAP = 1


class RgConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("scratch_bytes", C.c_uint64)]


class RgInput(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint64), ("track_index", C.c_int32), ("reserved", C.c_int32)]


class RgTrackResult(C.Structure):
    _fields_ = [
        ("loudness_db", C.c_double), ("gain_db", C.c_double), ("peak", C.c_double),
        ("sample_rate", C.c_uint32), ("file_type", C.c_int32), ("status", C.c_int32),
        ("channels", C.c_uint32), ("n_frames", C.c_uint32), ("n_windows", C.c_uint32),
        ("n_dropped_windows", C.c_uint32), ("n_edge_windows", C.c_uint32),
        ("exact_rerun", C.c_uint32), ("index_rewalks", C.c_uint32),
    ]


class RgAlbumResult(C.Structure):
    _fields_ = [("album_loudness_db", C.c_double), ("album_gain_db", C.c_double), ("album_peak", C.c_double)]


class RgPeakResult(C.Structure):
    _fields_ = [("peak", C.c_double), ("peak_pcm", C.c_double), ("sample_rate", C.c_uint32), ("status", C.c_int32)]


class RgStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("tracks", C.c_uint64), ("frames", C.c_uint64), ("exact_reruns", C.c_uint64),
        ("last_ms_index", C.c_float), ("last_ms_huffman", C.c_float), ("last_ms_hybrid", C.c_float),
        ("last_ms_synth", C.c_float), ("last_ms_iir", C.c_float), ("last_ms_hist", C.c_float),
        ("last_ms_total", C.c_float), ("last_batches", C.c_uint32), ("last_block_windows", C.c_uint32), ("last_iir_samples", C.c_uint64),
        ("last_ms_gain_scan", C.c_float), ("last_ms_gain_merge", C.c_float), ("last_ms_gain_patch", C.c_float),
        ("last_gain_groups", C.c_uint32), ("last_ms_exchange", C.c_float), ("reserved", C.c_uint32),
    ]


class RgGainStats(C.Structure):
    _fields_ = [
        ("frame_count", C.c_uint64), ("gain_sum", C.c_uint64), ("gain_count", C.c_uint64),
        ("avg_gain", C.c_double), ("headroom_db", C.c_double), ("headroom_steps", C.c_int32),
        ("mpeg_version", C.c_int32), ("channel_mode", C.c_int32), ("status", C.c_int32),
        ("min_gain", C.c_uint8), ("max_gain", C.c_uint8), ("reserved", C.c_uint8 * 2), ("walk_rewalks", C.c_uint32),
    ]


# every symbol include/rg_engine.h declares: name -> (restype, argtypes)
_P = C.c_void_p
COMM_ID_BYTES = 128
ExchangeFn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t)
SYMBOLS = {
    "rg_is_available": (C.c_int, []),
    "rg_engine_create": (C.c_int, [C.POINTER(RgConfig), C.POINTER(_P)]),
    "rg_engine_destroy": (None, [_P]),
    "rg_status_str": (C.c_char_p, [C.c_int]),
    "rg_engine_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "rg_last_error": (C.c_char_p, []),
    "rg_analyze_tracks": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.POINTER(RgTrackResult), _P]),
    "rg_analyze_album": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.POINTER(RgTrackResult), C.POINTER(RgAlbumResult), _P]),
    "rg_find_peak": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.POINTER(RgPeakResult)]),
    "rg_analyze_pcm_f32": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_size_t, C.c_uint32, C.POINTER(RgTrackResult), _P]),
    "rg_corpus_upload": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.POINTER(_P)]),
    "rg_corpus_free": (None, [_P, _P]),
    "rg_corpus_bytes": (C.c_uint64, [_P]),
    "rg_analyze_corpus": (C.c_int, [_P, _P, C.POINTER(RgTrackResult), _P]),
    "rg_analyze_album_shard": (C.c_int, [_P, _P, C.POINTER(RgTrackResult), _P, _P]),
    "rg_album_finish": (C.c_int, [_P, _P, _P, C.POINTER(RgAlbumResult)]),
    "rg_comm_unique_id": (C.c_int, [_P]),
    "rg_engine_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "rg_engine_comm_attach": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "rg_engine_comm_attach_host": (C.c_int, [_P, C.c_int, C.c_int, ExchangeFn, _P]),
    "rg_engine_comm_release": (C.c_int, [_P]),
    "rg_analyze_album_sharded": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.c_uint64, C.POINTER(RgTrackResult), C.POINTER(RgAlbumResult), _P, C.POINTER(C.c_int64)]),
    "rg_analyze_album_sharded_corpus": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(RgTrackResult), C.POINTER(RgAlbumResult), _P, C.POINTER(C.c_int64)]),
    "rg_gain_analyze": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.POINTER(RgGainStats)]),
    "rg_gain_analyze_corpus": (C.c_int, [_P, _P, C.POINTER(RgGainStats)]),
    "rg_gain_apply": (C.c_int, [_P, C.POINTER(RgInput), C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P), _P]),
    "rg_engine_stats": (C.c_int, [_P, C.POINTER(RgStats)]),
    "rg_decode_pcm": (C.c_int, [_P, C.POINTER(RgInput), _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}

_lib = None


def load():
    """Load librg_engine.so (once). Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). mp3rgain_b200 has no CPU implementation to fall back to.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header ever drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
