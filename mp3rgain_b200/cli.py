"""CLI glue for the analysis path: the reference CLI's per-file analysis loops, batched.

The reference's `cmd_info` / `cmd_track_gain` / `cmd_album_gain` (src/main.rs:1119-1150, 1213-1290, 1296-1340) call
`analyze_track` / `find_max_amplitude` once per file; here the whole file list is uploaded once and goes through ONE `rg_analyze_corpus`
and ONE `rg_gain_analyze_corpus` call (the form that fills a GPU), and the lines are formatted exactly as the reference prints them to a
pipe (the `colored` crate emits no escape codes when stdout is not a terminal):

    info_tsv()          `-o` / `-o tsv` (mp3gain-compatible, what beets parses)   src/main.rs:1120-1124, 1703-1722, 1736
    info_json()         `-o json`: the JsonOutput envelope of cmd_info            src/main.rs:101-165, 1119-1150, 1815-1852
    track_gain_lines()  `-r` analysis lines                                       src/main.rs:1941-1968
    track_gain_json()   `-r -o json` (dry run: this front end analyses, it never writes)   src/main.rs:1213-1290, 1925-2000, 2012-2105
    album_gain_lines()  `-a` analysis lines                                       src/main.rs:1296-1340
    album_gain_json()   `-a -o json` (dry run)                                    src/main.rs:1296-1455
    expand_files_recursive()  `-R`                                                src/main.rs:436-470

Applying the gain afterwards is `mp3rgain_b200.gain.apply_gain` (tag writing -- the `-s c|d|i|a` modes, undo -- stays with the crate).
    python -m mp3rgain_b200.cli [-r | -a] [-o [tsv|json|text]] [-R] [-d <dB>] [-m <steps>] [-i <track>] [-k] [-c] [-s r|s] FILE...
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence, Tuple

import math

from . import _lib
from .gain import _MODES, _VERSIONS, db_to_steps, steps_to_db
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


# ---- `-R`: expand_files_recursive / collect_audio_files (src/main.rs:436-470) ---------------------------------------
AUDIO_EXTENSIONS = ("mp3", "m4a", "aac", "mp4")


def expand_files_recursive(paths: Sequence) -> List[str]:
    """Directories are walked (every level) for files whose extension is mp3 / m4a / aac / mp4 in any case; anything that is
    not a directory is kept as given; the result is sorted the way Rust sorts PathBufs (component by component)."""
    result: List[str] = []

    def collect(d):
        for name in os.listdir(d):
            path = os.path.join(d, name)
            if os.path.isdir(path):
                collect(path)
            else:
                ext = os.path.splitext(name)[1]
                # Path::extension: ".mp3" alone is a file name without extension
                if ext and name != ext and ext[1:].lower() in AUDIO_EXTENSIONS:
                    result.append(path)

    for p in paths:
        p = os.fspath(p)
        if os.path.isdir(p):
            collect(p)
        else:
            result.append(p)
    result.sort(key=lambda q: tuple(c for c in q.split(os.sep)))
    return result


# ---- `-o json`: the JsonOutput envelope (src/main.rs:101-165), printed with serde_json::to_string_pretty --------------
_FILE_FIELDS = ("file", "status", "frames", "mpeg_version", "channel_mode", "min_gain", "max_gain", "avg_gain", "headroom_steps",
                "headroom_db", "gain_applied_steps", "gain_applied_db", "loudness_db", "peak", "max_amplitude", "error", "warning", "dry_run")


def _json_f64(x: float) -> str:
    """serde_json's f64: shortest round-trip digits, always with a fraction or an exponent, exponents as `1e-7` / `1e21`."""
    if not math.isfinite(x):
        return "null"
    r = repr(float(x))
    if "e" in r:
        mant, exp = r.split("e")
        sign = "-" if exp.startswith("-") else ""
        return mant + "e" + sign + exp.lstrip("+-").lstrip("0")
    return r


def _json_str(v: str) -> str:
    out = ['"']
    for ch in v:
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ord(ch) < 0x20:
            out.append("\\u%04x" % ord(ch))
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def json_pretty(v, indent: int = 0) -> str:
    """serde_json::to_string_pretty: two-space indent, `"key": value`, empty containers as [] / {}."""
    pad, pad2 = "  " * indent, "  " * (indent + 1)
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return _json_f64(v)
    if isinstance(v, str):
        return _json_str(v)
    if isinstance(v, dict):
        if not v:
            return "{}"
        return "{\n" + ",\n".join(pad2 + _json_str(k) + ": " + json_pretty(x, indent + 1) for k, x in v.items()) + "\n" + pad + "}"
    if isinstance(v, (list, tuple)):
        if not v:
            return "[]"
        return "[\n" + ",\n".join(pad2 + json_pretty(x, indent + 1) for x in v) + "\n" + pad + "]"
    raise TypeError(type(v))


def json_file_result(file, **fields) -> dict:
    """JsonFileResult: `file` first, then the Some(..) fields in declaration order (None fields are not serialised)."""
    d = {"file": os.fspath(file)}
    for k in _FILE_FIELDS[1:]:
        if fields.get(k) is not None:
            d[k] = fields[k]
    assert not set(fields) - set(_FILE_FIELDS), set(fields) - set(_FILE_FIELDS)
    return d


def json_output(files=None, album=None, summary=None) -> str:
    d = {}
    if files is not None:
        d["files"] = files
    if album is not None:
        d["album"] = album
    if summary is not None:
        d["summary"] = summary
    return json_pretty(d)


def json_summary(total: int, successful: int, failed: int, dry_run: bool) -> dict:   # create_json_summary, src/main.rs:927-939
    d = {"total_files": total, "successful": successful, "failed": failed}
    if dry_run:
        d["dry_run"] = True
    return d


def clipping_check(steps: int, gain_db: float, peak: float, prevent_clipping: bool = False, ignore_clipping: bool = False,
                   wrap_gain: bool = False, quiet: bool = False) -> Tuple[int, Optional[str]]:
    """process_apply_replaygain_with_album's clipping logic (src/main.rs:2030-2075): (actual steps, warning or None)."""
    if steps > 0 and not wrap_gain:
        new_peak = peak * 10.0 ** (gain_db / 20.0)
        if new_peak > 1.0:
            if prevent_clipping:
                actual = max(db_to_steps(-20.0 * math.log10(peak)), 0)
                return actual, "gain reduced from %d to %d steps to prevent clipping (peak: %.4f)" % (steps, actual, peak)
            if not ignore_clipping and not quiet:
                return steps, "clipping warning: peak would be %.2f (>1.00)" % new_peak
    return steps, None


def info_json(files: Sequence) -> str:
    """cmd_info with `-o json` (src/main.rs:1119-1150): per MP3 the fields of mp3rgain::analyze() (process_info, :1815-1852),
    "info" for an MP4 container, "error" + the message otherwise.  One rg_gain_analyze call for the whole list."""
    eng = _engine()
    idx, images, errors = _load(files)
    emap = dict(errors)
    gst = eng.gain_analyze_images(images) if images else []
    pos = {i: k for k, i in enumerate(idx)}
    res = []
    for i, f in enumerate(files):
        if i in emap:
            res.append(json_file_result(f, status="error", error=emap[i]))
            continue
        g, img = gst[pos[i]], images[pos[i]]
        if len(img) >= 12 and bytes(img[4:8]) == b"ftyp" and bytes(img[8:12]) in (b"M4A ", b"M4B ", b"M4P ", b"M4V ", b"mp41", b"mp42", b"isom", b"iso2"):
            res.append(json_file_result(f, status="info"))     # mp4meta::is_mp4_file (src/mp4meta.rs:872-889)
        elif g.status != 0:
            res.append(json_file_result(f, status="error", error="No valid MP3 frames found"))   # src/lib.rs:498-500
        else:
            res.append(json_file_result(f, frames=int(g.frame_count), mpeg_version=_VERSIONS[g.mpeg_version], channel_mode=_MODES[g.channel_mode],
                                        min_gain=int(g.min_gain), max_gain=int(g.max_gain), avg_gain=float(g.avg_gain),
                                        headroom_steps=int(g.headroom_steps), headroom_db=float(g.headroom_db)))
    return json_output(files=res)


def track_gain_json(files: Sequence, gain_modifier: int = 0, track_index: Optional[int] = None, prevent_clipping: bool = False,
                    ignore_clipping: bool = False) -> str:
    """cmd_track_gain with `-o json` in dry-run mode: ONE rg_analyze_tracks call for the list, then per file what
    process_track_gain / process_apply_replaygain_with_album report without touching the file (src/main.rs:1925-2000, 2012-2105)."""
    eng = _engine()
    idx, images, errors = _load(files)
    emap = dict(errors)
    res_t, sts = eng.analyze_images(images, track_index) if images else ([], [])
    pos = {i: k for k, i in enumerate(idx)}
    out, failed = [], 0
    for i, f in enumerate(files):
        if i in emap or sts[pos[i]] != 0:
            msg = emap[i] if i in emap else _message(sts[pos[i]], os.fspath(f), res_t[pos[i]].sample_rate, -1 if track_index is None else track_index)
            out.append(json_file_result(f, status="error", error=msg))
            failed += 1
            continue
        r = res_t[pos[i]]
        steps = r.gain_steps() + gain_modifier
        if steps == 0:
            out.append(json_file_result(f, status="skipped", loudness_db=r.loudness_db, peak=r.peak, gain_applied_steps=0, gain_applied_db=0.0))
            continue
        actual, warn = clipping_check(steps, r.gain_db, r.peak, prevent_clipping, ignore_clipping)
        out.append(json_file_result(f, status="dry_run", loudness_db=r.loudness_db, peak=r.peak, gain_applied_steps=actual,
                                    gain_applied_db=steps_to_db(actual), warning=warn, dry_run=True))
    return json_output(files=out, summary=json_summary(len(files), 0, failed, True))


def album_gain_json(files: Sequence, gain_modifier: int = 0, track_index: Optional[int] = None, prevent_clipping: bool = False,
                    ignore_clipping: bool = False) -> Tuple[str, int]:
    """cmd_album_gain with `-o json` in dry-run mode (src/main.rs:1296-1455).  Returns (the document, the process exit code:
    1 when the album analysis failed)."""
    try:
        images = [_read(f) for f in files]
        alb = _engine().analyze_album_images(images, track_index, names=list(files))
    except ReplayGainError:
        return json_output(summary=json_summary(len(files), 0, len(files), True)), 1
    steps = alb.album_gain_steps() + gain_modifier
    album = {"loudness_db": alb.album_loudness_db, "gain_db": alb.album_gain_db, "gain_steps": steps, "peak": alb.album_peak}
    out = []
    for f, t in zip(files, alb.tracks):
        if steps == 0:
            out.append(json_file_result(f, status="skipped", loudness_db=t.loudness_db, peak=t.peak, gain_applied_steps=0, gain_applied_db=0.0))
        else:
            actual, warn = clipping_check(steps, t.gain_db, t.peak, prevent_clipping, ignore_clipping)
            out.append(json_file_result(f, status="dry_run", loudness_db=t.loudness_db, peak=t.peak, gain_applied_steps=actual,
                                        gain_applied_db=steps_to_db(actual), warning=warn, dry_run=True))
    return json_output(files=out, album=album, summary=json_summary(len(files), 0, 0, True)), 0


def _rust_f64(x: float) -> str:
    """Rust `{}` of an f64: shortest round-trip digits, no exponent for ordinary magnitudes, integral values without '.0'
    only for integers -- `89.0_f64` prints as `89`."""
    return str(int(x)) if float(x).is_integer() else repr(float(x))


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = list(sys.argv[1:] if argv is None else argv)
    fmt, mod_db, mod_steps, track_index, files = None, 0.0, 0, None, []
    mode, recursive, prevent, ignore = "info", False, False, False
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
        elif a == "-m":
            i += 1
            mod_steps = int(args[i])
        elif a == "-i":
            i += 1
            track_index = int(args[i])
        elif a == "-s":                      # stored-tag modes (src/main.rs:246-266): only the ones that mean "analyse" apply here
            i += 1
            if i >= len(args) or args[i] not in ("r", "s"):
                sys.stderr.write("error: -s %s is tag I/O (check / delete / ID3v2 / APEv2), which stays with the crate; this front end takes -s r and -s s\n"
                                 % (args[i] if i < len(args) else ""))
                return 1
        elif a in ("-r", "-a"):
            mode = "track" if a == "-r" else "album"
        elif a == "-R":
            recursive = True
        elif a == "-k":
            prevent = True
        elif a == "-c":
            ignore = True
        elif a in ("-n", "-q"):              # dry run is what this front end does anyway; quiet has nothing to silence in these reports
            pass
        elif a.startswith("-") and len(a) > 1:
            sys.stderr.write("error: option %s is not part of the batched analysis front end (see python -m mp3rgain_b200.cli with no arguments)\n" % a)
            return 1
        else:
            files.append(a)
        i += 1
    if recursive:
        files = expand_files_recursive(files)
    if not files or (mode == "info" and fmt not in ("tsv", "json")):
        sys.stderr.write(__doc__.strip().splitlines()[-1].strip() + "\n")
        return 1
    if mode == "info" and fmt == "tsv":
        out, err = info_tsv(files, mod_db, track_index)
    elif mode == "info":
        out, err = [info_json(files)], []
    elif mode == "track" and fmt == "json":
        out, err = [track_gain_json(files, mod_steps, track_index, prevent, ignore)], []
    elif mode == "track":
        out, err, _ = track_gain_lines(files, mod_steps, track_index)
    elif fmt == "json":
        doc, rc = album_gain_json(files, mod_steps, track_index, prevent, ignore)
        sys.stdout.write(doc + "\n")
        return rc
    else:
        try:
            out, _ = album_gain_lines(files, mod_steps, track_index)
            err = []
        except ReplayGainError as e:
            sys.stderr.write("error: Failed to analyze album: %s\n" % e)   # src/main.rs:1448
            return 1
    sys.stdout.write("\n".join(out) + "\n")
    if err:
        sys.stderr.write("\n".join(err) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
