"""Host-side mirror of the reference's `mp3rgain::replaygain` module (src/replaygain.rs) on top of the
engine's C ABI.  Same names, argument meaning and error behaviour:

    analyze_track(path) / analyze_track_with_index(path, idx)   -> ReplayGainResult      (:929, :935)
    analyze_album(paths) / analyze_album_with_index(paths, idx) -> AlbumGainResult       (:1033, :1044)
    find_peak_amplitude(path)                                   -> PeakAmplitudeResult   (:1140)
    is_available()                                              -> bool                  (:1119)

Errors are raised as ReplayGainError carrying the reference's anyhow message.  All compute happens in
librg_engine.so on the GPU; this module only reads files, marshals pointers and formats results.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum
import os
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import HIST_SIZE, PINK_REF

REPLAYGAIN_REFERENCE_DB = 89.0  # src/replaygain.rs:37
GAIN_STEP_DB = 1.5              # src/lib.rs:48


class ReplayGainError(RuntimeError):
    def __init__(self, status: int, message: str, index: Optional[int] = None):
        super().__init__(message)
        self.status = status
        self.index = index      # album mode: position in the input list of the file that failed the album


class AudioFileType(enum.Enum):  # src/replaygain.rs:47-53
    Mp3 = 0
    Aac = 1


def _round_half_away(x: float) -> int:
    return int(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


@dataclasses.dataclass
class ReplayGainResult:  # src/replaygain.rs:56-75
    loudness_db: float
    gain_db: float
    peak: float
    sample_rate: int
    file_type: AudioFileType = AudioFileType.Mp3
    # engine diagnostics (not in the reference struct)
    channels: int = 0
    n_frames: int = 0
    n_windows: int = 0
    n_dropped_windows: int = 0
    n_edge_windows: int = 0
    exact_rerun: int = 0
    index_rewalks: int = 0
    histogram: Optional[np.ndarray] = None

    def gain_steps(self) -> int:
        return _round_half_away(self.gain_db / GAIN_STEP_DB)


@dataclasses.dataclass
class AlbumGainResult:  # src/replaygain.rs:78-95
    tracks: List[ReplayGainResult]
    album_loudness_db: float
    album_gain_db: float
    album_peak: float
    album_histogram: Optional[np.ndarray] = None

    def album_gain_steps(self) -> int:
        return _round_half_away(self.album_gain_db / GAIN_STEP_DB)


@dataclasses.dataclass
class PeakAmplitudeResult:  # src/replaygain.rs:1125-1132
    peak: float
    peak_pcm: float
    sample_rate: int


def _message(status: int, path: str = "", sample_rate: int = 0, track_index: int = -1) -> str:
    # the strings the reference's CLI prints verbatim (src/replaygain.rs:805,822,834,842-846,857,869-872)
    if status == _lib.RG_ERR_OPEN:
        return f"Failed to open: {path}"
    if status == _lib.RG_ERR_PROBE:
        return f"Failed to probe format: {path}"
    if status == _lib.RG_ERR_NO_AUDIO_TRACK:
        return "No audio track found"
    if status == _lib.RG_ERR_TRACK_INDEX:
        return f"Track index {track_index} out of range (file has 1 audio track(s))"
    if status == _lib.RG_ERR_UNKNOWN_SAMPLE_RATE:
        return "Unknown sample rate"
    if status == _lib.RG_ERR_UNSUPPORTED_RATE:
        return (f"Unsupported sample rate: {sample_rate} Hz. Supported rates: 96000, 88200, 64000, 48000, 44100, "
                "32000, 24000, 22050, 16000, 12000, 11025, 8000")
    lib = _lib.load()
    detail = lib.rg_last_error().decode() if status == _lib.RG_ERR_CUDA else ""
    return lib.rg_status_str(status).decode() + (f": {detail}" if detail else "")


class Engine:
    """One engine per GPU (rg_engine).  Batch entry points take raw file images (bytes / numpy uint8)."""

    def __init__(self, device: int = 0, exact: bool = False, flags: int = 0, scratch_bytes: int = 0):
        self._lib = _lib.load()
        cfg = _lib.RgConfig(device=device, flags=flags | (_lib.RG_FLAG_EXACT_IIR if exact else 0), scratch_bytes=scratch_bytes)
        h = C.c_void_p()
        rc = self._lib.rg_engine_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rg_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key: str, value: float):
        rc = self._lib.rg_engine_set_option(self._h, key.encode(), float(value))
        if rc != 0:
            raise ReplayGainError(rc, f"unknown engine option {key!r}")

    def stats(self) -> _lib.RgStats:
        st = _lib.RgStats()
        self._lib.rg_engine_stats(self._h, C.byref(st))
        return st

    # ---- marshalling helpers -------------------------------------------------------------------
    @staticmethod
    def _inputs(images: Sequence, track_index: Optional[int]):
        n = len(images)
        arr = (_lib.RgInput * n)()
        keep = []
        for i, img in enumerate(images):
            a = img if isinstance(img, np.ndarray) else np.frombuffer(img, dtype=np.uint8)
            a = np.ascontiguousarray(a, dtype=np.uint8)
            keep.append(a)
            arr[i].data = a.ctypes.data if a.size else None
            arr[i].len = a.size
            arr[i].track_index = -1 if track_index is None else int(track_index)
        return arr, keep

    @staticmethod
    def _result(r: _lib.RgTrackResult, hist=None) -> ReplayGainResult:
        return ReplayGainResult(
            loudness_db=r.loudness_db, gain_db=r.gain_db, peak=r.peak, sample_rate=r.sample_rate,
            file_type=AudioFileType(r.file_type), channels=r.channels, n_frames=r.n_frames, n_windows=r.n_windows,
            n_dropped_windows=r.n_dropped_windows, n_edge_windows=r.n_edge_windows, exact_rerun=r.exact_rerun,
            index_rewalks=r.index_rewalks, histogram=hist)

    # ---- batch API --------------------------------------------------------------------------------
    def analyze_images(self, images: Sequence, track_index: Optional[int] = None, want_hist: bool = False,
                       raise_on_error: bool = False):
        """rg_analyze_tracks over in-memory file images. Returns (results, statuses)."""
        n = len(images)
        arr, keep = self._inputs(images, track_index)
        out = (_lib.RgTrackResult * n)()
        hist = np.zeros((n, HIST_SIZE), np.uint32) if want_hist else None
        rc = self._lib.rg_analyze_tracks(self._h, arr, n, out, hist.ctypes.data if want_hist else None)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        res, sts = [], []
        for i in range(n):
            sts.append(out[i].status)
            if out[i].status != 0 and raise_on_error:
                raise ReplayGainError(out[i].status, _message(out[i].status, sample_rate=out[i].sample_rate,
                                                              track_index=-1 if track_index is None else track_index))
            res.append(self._result(out[i], hist[i] if want_hist else None))
        return res, sts

    def analyze_album_images(self, images: Sequence, track_index: Optional[int] = None, want_hist: bool = False,
                             names: Optional[Sequence] = None) -> AlbumGainResult:
        """rg_analyze_album.  `names` (the files' paths) only serve the error message: the reference fails the album with the
        first failing file's error, which names that file (src/replaygain.rs:1055, :822)."""
        n = len(images)
        arr, keep = self._inputs(images, track_index)
        out = (_lib.RgTrackResult * max(n, 1))()
        alb = _lib.RgAlbumResult()
        ah = np.zeros(HIST_SIZE, np.uint64) if want_hist else None
        rc = self._lib.rg_analyze_album(self._h, arr, n, out, C.byref(alb), ah.ctypes.data if want_hist else None)
        if rc != 0:
            bad = next((i for i in range(n) if out[i].status != 0), None)
            raise ReplayGainError(rc, _message(rc, path=os.fspath(names[bad]) if names is not None and bad is not None else "",
                                               sample_rate=out[bad].sample_rate if bad is not None else 0,
                                               track_index=-1 if track_index is None else track_index), index=bad)
        return AlbumGainResult([self._result(out[i]) for i in range(n)], alb.album_loudness_db, alb.album_gain_db,
                               alb.album_peak, ah)

    # ---- album mode across GPUs: this engine is one rank of several (the engine owns the exchange) ----------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """rg_comm_unique_id (ncclGetUniqueId): rank 0 creates it and hands it to the other ranks."""
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
        rc = _lib.load().rg_comm_unique_id(buf)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return bytes(buf)

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(unique_id)
        rc = self._lib.rg_engine_comm_init(self._h, rank, nranks, buf)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))

    def comm_attach_host(self, rank: int, nranks: int, fn):
        """Another transport for the exchange: fn(sum_array_u64, max_array_u64) reduces both numpy arrays in place over all ranks."""
        def tramp(ctx, psum, nsum, pmax, nmax):
            try:
                fn(np.ctypeslib.as_array(psum, (nsum,)), np.ctypeslib.as_array(pmax, (nmax,)))
                return 0
            except Exception:   # noqa: BLE001 -- must not unwind through the C frame
                return 1
        self._xchg = _lib.ExchangeFn(tramp)   # keep the trampoline alive as long as the engine uses it
        rc = self._lib.rg_engine_comm_attach_host(self._h, rank, nranks, self._xchg, None)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))

    def comm_release(self):
        self._lib.rg_engine_comm_release(self._h)
        self._xchg = None

    def analyze_album_sharded_images(self, images: Sequence, shard_offset: int, track_index: Optional[int] = None,
                                     want_hist: bool = True, names: Optional[Sequence] = None) -> AlbumGainResult:
        """rg_analyze_album_sharded: `images` = this rank's shard, which starts at `shard_offset` in the global file list;
        `names` = the GLOBAL list of paths (for the error message).  Every rank gets the same album numbers or the same error."""
        n = len(images)
        arr, keep = self._inputs(images, track_index)
        out = (_lib.RgTrackResult * max(n, 1))()
        alb = _lib.RgAlbumResult()
        ah = np.zeros(HIST_SIZE, np.uint64) if want_hist else None
        bad = C.c_int64(-1)
        rc = self._lib.rg_analyze_album_sharded(self._h, arr, n, int(shard_offset), out, C.byref(alb),
                                                ah.ctypes.data if want_hist else None, C.byref(bad))
        if rc != 0:
            g = int(bad.value)
            local = g - int(shard_offset)
            raise ReplayGainError(rc, _message(rc, path=os.fspath(names[g]) if names is not None and g >= 0 else "",
                                               sample_rate=out[local].sample_rate if 0 <= local < n else 0,
                                               track_index=-1 if track_index is None else track_index), index=g if g >= 0 else None)
        return AlbumGainResult([self._result(out[i]) for i in range(n)], alb.album_loudness_db, alb.album_gain_db,
                               alb.album_peak, ah)

    def find_peak_images(self, images: Sequence):
        n = len(images)
        arr, keep = self._inputs(images, None)
        out = (_lib.RgPeakResult * n)()
        rc = self._lib.rg_find_peak(self._h, arr, n, out)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return [(PeakAmplitudeResult(o.peak, o.peak_pcm, o.sample_rate), o.status) for o in out]

    def analyze_pcm_f32(self, planes: Sequence[np.ndarray], sample_rate: int, want_hist: bool = True) -> ReplayGainResult:
        """Stage-B entry: planar normalised f32 PCM (rg_analyze_pcm_f32)."""
        planes = [np.ascontiguousarray(p, dtype=np.float32) for p in planes]
        nch = len(planes)
        frames = planes[0].size
        ptrs = (C.c_void_p * nch)(*[p.ctypes.data for p in planes])
        out = _lib.RgTrackResult()
        hist = np.zeros(HIST_SIZE, np.uint32) if want_hist else None
        rc = self._lib.rg_analyze_pcm_f32(self._h, ptrs, nch, frames, sample_rate, C.byref(out),
                                          hist.ctypes.data if want_hist else None)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc, sample_rate=sample_rate))
        return self._result(out, hist)

    # ---- device-resident corpora: copy the file images once, run several passes over them ---------------
    def upload(self, images: Sequence, track_index: Optional[int] = None) -> "Corpus":
        """rg_corpus_upload: the images cross the bus once; Corpus.analyze() / Corpus.gain_analyze() then work from HBM."""
        return Corpus(self, images, track_index)

    # ---- the global_gain fields (mp3rgain::analyze / apply_gain*, src/lib.rs) -------------------------
    def gain_analyze_images(self, images: Sequence):
        """rg_gain_analyze over in-memory file images.  Returns the list of RgGainStats (per-entry status inside)."""
        n = len(images)
        arr, keep = self._inputs(images, None)
        out = (_lib.RgGainStats * max(n, 1))()
        rc = self._lib.rg_gain_analyze(self._h, arr, n, out)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return [out[i] for i in range(n)]

    def gain_apply_images(self, images: Sequence, gain_steps: int, wrap: bool = False, channel: int = -1):
        """rg_gain_apply: returns (patched images as new uint8 arrays, modified frames per image)."""
        n = len(images)
        arr, keep = self._inputs(images, None)
        outs = [np.empty(a.size, np.uint8) for a in keep]
        ptrs = (C.c_void_p * max(n, 1))(*[o.ctypes.data if o.size else None for o in outs])
        mod = np.zeros(max(n, 1), np.uint32)
        rc = self._lib.rg_gain_apply(self._h, arr, n, int(gain_steps), _lib.RG_GAIN_WRAP if wrap else _lib.RG_GAIN_SATURATE,
                                     int(channel), ptrs, mod.ctypes.data)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return outs, [int(x) for x in mod[:n]]

    def decode_pcm(self, image) -> tuple:
        """Stage-A parity hook: (planes[nch, frames] float32, sample_rate)."""
        arr, keep = self._inputs([image], None)
        max_frames = max(1152, (len(keep[0]) // 24 + 2) * 1152)
        buf = np.zeros((2, max_frames), np.float32)
        frames = C.c_size_t()
        nch = C.c_uint32()
        sr = C.c_uint32()
        rc = self._lib.rg_decode_pcm(self._h, arr, buf.ctypes.data, max_frames, C.byref(frames), C.byref(nch), C.byref(sr))
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return buf[: nch.value, : frames.value].copy(), sr.value


class Corpus:
    """A batch of file images resident in device memory (rg_corpus).  Use as a context manager."""

    def __init__(self, engine: Engine, images: Sequence, track_index: Optional[int] = None):
        self._eng, self._n, self._track_index = engine, len(images), track_index
        arr, keep = Engine._inputs(images, track_index)
        self._h = C.c_void_p()
        rc = engine._lib.rg_corpus_upload(engine._h, arr, self._n, C.byref(self._h))
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))

    def close(self):
        if getattr(self, "_h", None):
            self._eng._lib.rg_corpus_free(self._eng._h, self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def analyze(self, want_hist: bool = False):
        """rg_analyze_corpus.  Returns (results, statuses)."""
        n = self._n
        out = (_lib.RgTrackResult * max(n, 1))()
        hist = np.zeros((n, HIST_SIZE), np.uint32) if want_hist else None
        rc = self._eng._lib.rg_analyze_corpus(self._eng._h, self._h, out, hist.ctypes.data if want_hist else None)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return [Engine._result(out[i], hist[i] if want_hist else None) for i in range(n)], [out[i].status for i in range(n)]

    def gain_analyze(self):
        """rg_gain_analyze_corpus.  Returns the list of RgGainStats."""
        out = (_lib.RgGainStats * max(self._n, 1))()
        rc = self._eng._lib.rg_gain_analyze_corpus(self._eng._h, self._h, out)
        if rc != 0:
            raise ReplayGainError(rc, _message(rc))
        return [out[i] for i in range(self._n)]


# ---- module-level functions with the reference's names ------------------------------------------------
_default_engine: Optional[Engine] = None
_default_engine_lock = threading.Lock()


def _engine() -> Engine:
    """The module-level functions share one engine; calls from several threads queue on the engine's own lock."""
    global _default_engine
    with _default_engine_lock:
        if _default_engine is None:
            _default_engine = Engine(device=int(os.environ.get("LOCAL_RANK", "0")))
        return _default_engine


def is_available() -> bool:
    """replaygain::is_available (src/replaygain.rs:1119)."""
    try:
        return bool(_lib.load().rg_is_available())
    except (RuntimeError, OSError):
        return False


def _read(path) -> np.ndarray:
    try:
        return np.fromfile(os.fspath(path), dtype=np.uint8)
    except OSError:
        raise ReplayGainError(_lib.RG_ERR_OPEN, f"Failed to open: {os.fspath(path)}") from None


def analyze_track_with_index(file_path, track_index: Optional[int] = None) -> ReplayGainResult:
    img = _read(file_path)
    res, sts = _engine().analyze_images([img], track_index)
    if sts[0] != 0:
        raise ReplayGainError(sts[0], _message(sts[0], os.fspath(file_path), res[0].sample_rate,
                                               -1 if track_index is None else track_index))
    return res[0]


def analyze_track(file_path) -> ReplayGainResult:
    return analyze_track_with_index(file_path, None)


def analyze_album_with_index(files: Sequence, track_index: Optional[int] = None) -> AlbumGainResult:
    images = [_read(f) for f in files]  # an unreadable file aborts the album, like the `?` at :1055
    return _engine().analyze_album_images(images, track_index, names=list(files))


def analyze_album(files: Sequence) -> AlbumGainResult:
    return analyze_album_with_index(files, None)


def find_peak_amplitude(file_path) -> PeakAmplitudeResult:
    img = _read(file_path)
    (res, st), = _engine().find_peak_images([img])
    if st != 0:
        raise ReplayGainError(st, _message(st, os.fspath(file_path), res.sample_rate))
    return res
