"""Host-side mirror of the reference crate's lossless-gain surface (`mp3rgain::analyze`, `apply_gain*`, src/lib.rs) on top
of the engine's C ABI.  Same names, argument meaning and error behaviour:

    analyze(path)                                  -> Mp3Analysis       (:470-514)
    apply_gain(path, steps) / apply_gain_db        -> modified frames   (:602-634)
    apply_gain_wrap(path, steps)                   -> modified frames   (:1232-1246)
    apply_gain_channel(path, channel, steps)       -> modified frames   (:750-771)
    is_mono(path)                                  -> bool              (:661-664)
    find_max_amplitude(path)                       -> (max_amplitude, max_gain, min_gain)   (:1172-1199)
    db_to_steps / steps_to_db                                           (:637-644)

The frame walk and the bit-field patch run in librg_engine.so on the GPU (k_gain_fields.cuh); this module reads and writes
the files, as the reference's functions do around their in-memory walk.  The APEv2 undo tags (`*_with_undo`, `undo_gain`)
are tag I/O and stay with the crate.
"""
from __future__ import annotations

import dataclasses
import enum
import os

import numpy as np

from . import _lib
from .replaygain import GAIN_STEP_DB, ReplayGainError, _engine, _round_half_away, find_peak_amplitude

MAX_GAIN = 255  # src/lib.rs:51
MIN_GAIN = 0    # src/lib.rs:54
_VERSIONS = ("MPEG1", "MPEG2", "MPEG2.5")                          # MpegVersion::as_str, src/lib.rs:85-93
_MODES = ("Stereo", "Joint Stereo", "Dual Channel", "Mono")        # ChannelMode::as_str, :111-118


class GainError(ReplayGainError):
    pass


class Channel(enum.Enum):  # src/lib.rs:647-657
    Left = 0
    Right = 1

    def index(self) -> int:
        return self.value

    @staticmethod
    def from_index(index: int):
        return Channel(index) if index in (0, 1) else None


@dataclasses.dataclass
class Mp3Analysis:  # src/lib.rs:57-75
    frame_count: int
    mpeg_version: str
    channel_mode: str
    min_gain: int
    max_gain: int
    avg_gain: float
    headroom_steps: int
    headroom_db: float


def db_to_steps(db: float) -> int:
    return _round_half_away(db / GAIN_STEP_DB)


def steps_to_db(steps: int) -> float:
    return steps * GAIN_STEP_DB


def _read(path) -> np.ndarray:
    try:
        return np.fromfile(os.fspath(path), dtype=np.uint8)
    except OSError:
        raise GainError(_lib.RG_ERR_OPEN, f"Failed to read: {os.fspath(path)}") from None


def _write(path, image: np.ndarray):
    try:
        image.tofile(os.fspath(path))
    except OSError:
        raise GainError(_lib.RG_ERR_OPEN, f"Failed to write: {os.fspath(path)}") from None


def analysis_from_stats(st) -> Mp3Analysis:
    if st.status != 0:
        raise GainError(st.status, "No valid MP3 frames found")
    return Mp3Analysis(int(st.frame_count), _VERSIONS[st.mpeg_version], _MODES[st.channel_mode], int(st.min_gain),
                       int(st.max_gain), float(st.avg_gain), int(st.headroom_steps), float(st.headroom_db))


def analyze(file_path) -> Mp3Analysis:
    (st,) = _engine().gain_analyze_images([_read(file_path)])
    return analysis_from_stats(st)


def _apply(file_path, gain_steps: int, wrap: bool, channel: int) -> int:
    if gain_steps == 0:
        return 0
    (image,), (frames,) = _engine().gain_apply_images([_read(file_path)], gain_steps, wrap, channel)
    _write(file_path, image)
    return frames


def apply_gain(file_path, gain_steps: int) -> int:
    return _apply(file_path, gain_steps, False, -1)


def apply_gain_db(file_path, gain_db: float) -> int:
    return apply_gain(file_path, db_to_steps(gain_db))


def apply_gain_wrap(file_path, gain_steps: int) -> int:
    return _apply(file_path, gain_steps, True, -1)


def is_mono(file_path) -> bool:
    return analyze(file_path).channel_mode == "Mono"


def apply_gain_channel(file_path, channel: Channel, gain_steps: int) -> int:
    if gain_steps == 0:
        return 0
    if analyze(file_path).channel_mode == "Mono":
        raise GainError(_lib.RG_ERR_INVALID_ARG, "Cannot apply channel-specific gain to mono file. Use -g for mono files.")
    return _apply(file_path, gain_steps, False, channel.index())


def find_max_amplitude(file_path):
    a = analyze(file_path)
    return find_peak_amplitude(file_path).peak, a.max_gain, a.min_gain
