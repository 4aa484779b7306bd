"""mp3rgain_b200 -- B200-native ReplayGain 1.0 analysis engine (drop-in for mp3rgain's analysis path)."""
from . import gain, replaygain  # noqa: F401
from .replaygain import (  # noqa: F401
    AlbumGainResult, AudioFileType, Corpus, Engine, PeakAmplitudeResult, ReplayGainError, ReplayGainResult,
    analyze_album, analyze_album_with_index, analyze_track, analyze_track_with_index, find_peak_amplitude,
    is_available,
)
