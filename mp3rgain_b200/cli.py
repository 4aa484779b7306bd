"""CLI glue for the analysis path: the reference CLI's per-file analysis loops, batched.

The reference's `cmd_info` / `cmd_track_gain` / `cmd_album_gain` (src/main.rs:1119-1150, 1213-1290, 1296-1340) call
`analyze_track` / `find_max_amplitude` once per file; here the whole file list is uploaded once and goes through ONE `rg_analyze_corpus`
and ONE `rg_gain_analyze_corpus` call (the form that fills a GPU), and the lines are formatted exactly as the reference prints them to a
pipe (the `colored` crate emits no escape codes when stdout is not a terminal):

    info_tsv()          `-o` / `-o tsv` (mp3gain-compatible, what beets parses)   src/main.rs:1120-1124, 1703-1722, 1736
    track_gain_lines()  `-r` analysis lines                                       src/main.rs:1941-1968
    album_gain_lines()  `-a` analysis lines                                       src/main.rs:1296-1340

Applying the gain afterwards is `mp3rgain_b200.gain.apply_gain` (tag writing stays with the crate).
    python -m mp3rgain_b200.cli -o [tsv] [-d <dB>] [-i <track>] FILE...
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence, Tuple

from . import _lib
from .gain import db_to_steps
from .replaygain import REPLAYGAIN_REFERENCE_DB, ReplayGainError, _engine, _message, _read

TSV_HEADER = "File\tMP3 gain\tdB gain\tMax Amplitude\tMax global_gain\tMin global_gain"   # src/main.rs:1123


def get_filename(path) -> str:   # src/main.rs:28-32
    return os.path.basename(os.fspath(path)) or "unknown"


def format_tsv_line(filename: str, gain_db: float, peak: float, max_gain: int, min_gain: int) -> str:
    """`"{}\\t{}\\t{:.6}\\t{:.6}\\t{}\\t{}"` (src/main.rs:1719-1722); gain_db already includes the -d modifier (:1709)."""
    return "%s\t%d\t%.6f\t%.6f\t%d\t%d" % (filename, db_to_steps(gain_db), gain_db, peak * 32768.0, max_gain, min_gain)


def format_track_line(loudness_db: float, gain_db: float, base_steps: int, gain_modifier: int, peak: float) -> str:
    """src/main.rs:1957-1968"""
    mod = " + %d = %d" % (gain_modifier, base_steps + gain_modifier) if gain_modifier != 0 else ""
    return "      Loudness: %.1f dB, Gain: %+.1f dB (%d steps%s), Peak: %.4f" % (loudness_db, gain_db, base_steps, mod, peak)


def format_album_lines(album_loudness_db: float, album_gain_db: float, album_steps: int, gain_modifier: int, album_peak: float) -> List[str]:
    """src/main.rs:1322-1338"""
    mod = " + %d = %d" % (gain_modifier, album_steps + gain_modifier) if gain_modifier != 0 else ""
    return ["", "  Album loudness: %.1f dB" % album_loudness_db,
            "  Album gain:     %+.1f dB (%d steps%s)" % (album_gain_db, album_steps, mod),
            "  Album peak:     %.4f" % album_peak, ""]


def _load(files: Sequence) -> Tuple[list, list, list]:
    """Reads every file; returns (indices that were read, their images, [(index, error message)])."""
    idx, images, errors = [], [], []
    for i, f in enumerate(files):
        try:
            images.append(_read(f))
            idx.append(i)
        except ReplayGainError as e:
            errors.append((i, str(e)))
    return idx, images, errors


def info_tsv(files: Sequence, gain_modifier_db: float = 0.0, track_index: Optional[int] = None):
    """cmd_info in TSV mode.  Returns (stdout lines incl. the header, stderr lines `"{filename} - {error}"`)."""
    eng = _engine()
    idx, images, errors = _load(files)
    out_lines: List[Optional[str]] = [None] * len(files)
    err_lines: List[Optional[str]] = [None] * len(files)
    for i, msg in errors:
        err_lines[i] = "%s - %s" % (get_filename(files[i]), msg)
    if images:
        with eng.upload(images, track_index) as cor:   # the images cross the bus once for both passes
            res, sts = cor.analyze()
            gst = cor.gain_analyze()
        for k, i in enumerate(idx):
            if sts[k] != 0:
                err_lines[i] = "%s - %s" % (get_filename(files[i]), _message(sts[k], os.fspath(files[i]), res[k].sample_rate,
                                                                            -1 if track_index is None else track_index))
                continue
            # find_max_amplitude(file).unwrap_or((1.0, 255, 0)) (src/main.rs:1706-1707)
            mx, mn = (gst[k].max_gain, gst[k].min_gain) if gst[k].status == 0 else (255, 0)
            out_lines[i] = format_tsv_line(get_filename(files[i]), res[k].gain_db + gain_modifier_db, res[k].peak, mx, mn)
    return [TSV_HEADER] + [x for x in out_lines if x is not None], [x for x in err_lines if x is not None]


def track_gain_lines(files: Sequence, gain_modifier: int = 0, track_index: Optional[int] = None, dry_run: bool = True):
    """The analysis part of cmd_track_gain / process_track_gain in text mode.  Returns (stdout lines, stderr lines,
    per-file suggested steps or None)."""
    eng = _engine()
    pre = "[DRY RUN] " if dry_run else ""
    out = ["%smp3rgain Analyzing and %s track gain to %d file(s)" % (pre, "would apply" if dry_run else "applying", len(files)),
           "  Target: %s dB (ReplayGain 1.0)" % _rust_f64(REPLAYGAIN_REFERENCE_DB)]
    if gain_modifier != 0:
        out.append("  Gain modifier: %+d steps" % gain_modifier)
    out.append("")
    err: List[str] = []
    steps: List[Optional[int]] = [None] * len(files)
    idx, images, errors = _load(files)
    emap = dict(errors)
    res, sts = eng.analyze_images(images, track_index) if images else ([], [])
    pos = {i: k for k, i in enumerate(idx)}
    for i, f in enumerate(files):
        name = get_filename(f)
        out.append("  -> %sAnalyzing %s..." % (pre, name))
        if i in emap:
            err.append("  x %s - %s" % (name, emap[i]))
            continue
        k = pos[i]
        if sts[k] != 0:
            err.append("  x %s - %s" % (name, _message(sts[k], os.fspath(f), res[k].sample_rate, -1 if track_index is None else track_index)))
            continue
        base = res[k].gain_steps()
        out.append(format_track_line(res[k].loudness_db, res[k].gain_db, base, gain_modifier, res[k].peak))
        steps[i] = base + gain_modifier
        if steps[i] == 0:
            out.append("  . %s (no adjustment needed)" % name)
    return out, err, steps


def album_gain_lines(files: Sequence, gain_modifier: int = 0, track_index: Optional[int] = None, dry_run: bool = True):
    """The analysis part of cmd_album_gain in text mode.  Returns (stdout lines, album steps incl. the modifier)."""
    pre = "[DRY RUN] " if dry_run else ""
    out = ["%smp3rgain Analyzing album gain for %d file(s)" % (pre, len(files)),
           "  Target: %s dB (ReplayGain 1.0)" % _rust_f64(REPLAYGAIN_REFERENCE_DB)]
    if gain_modifier != 0:
        out.append("  Gain modifier: %+d steps" % gain_modifier)
    out += ["", "  -> Analyzing tracks..."]
    images = [_read(f) for f in files]   # the first unreadable file aborts the album (src/replaygain.rs:1055)
    alb = _engine().analyze_album_images(images, track_index)
    out += format_album_lines(alb.album_loudness_db, alb.album_gain_db, alb.album_gain_steps(), gain_modifier, alb.album_peak)
    return out, alb.album_gain_steps() + gain_modifier


def _rust_f64(x: float) -> str:
    """Rust `{}` of an f64: shortest round-trip digits, no exponent for ordinary magnitudes, integral values without '.0'
    only for integers -- `89.0_f64` prints as `89`."""
    return str(int(x)) if float(x).is_integer() else repr(float(x))


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = list(sys.argv[1:] if argv is None else argv)
    fmt, mod_db, track_index, files = None, 0.0, None, []
    i = 0
    while i < len(args):
        a = args[i]
        if a == "-o":
            fmt = "tsv"
            if i + 1 < len(args) and args[i + 1].lower() in ("json", "text", "tsv", "db"):   # src/main.rs:273-296
                i += 1
                fmt = "tsv" if args[i].lower() in ("tsv", "db") else args[i].lower()
        elif a == "-d":
            i += 1
            mod_db = float(args[i])
        elif a == "-i":
            i += 1
            track_index = int(args[i])
        elif a.startswith("-") and len(a) > 1:
            sys.stderr.write("error: this front end only carries the batched TSV report (-o [tsv] [-d dB] [-i track] FILE...)\n")
            return 1
        else:
            files.append(a)
        i += 1
    if fmt != "tsv" or not files:
        sys.stderr.write(__doc__.strip().splitlines()[-1].strip() + "\n")
        return 1
    out, err = info_tsv(files, mod_db, track_index)
    sys.stdout.write("\n".join(out) + "\n")
    if err:
        sys.stderr.write("\n".join(err) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
