"""GPU test of the batched CLI glue: the TSV / text lines for the reference's fixtures, built from ONE
rg_analyze_tracks + ONE rg_gain_analyze call, equal the lines the reference's per-file loop would print from the oracle's
numbers (src/main.rs:1703-1722, 1941-1968, 1296-1340)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["test_joint_stereo.mp3", "test_mono.mp3", "test_vbr.mp3", "test_stereo.mp3"]


def _round_half_away(x):
    return int(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def test_info_tsv_matches_the_per_file_loop(oracle, tmp_path, capsys):
    from mp3rgain_b200 import cli
    files = [os.path.join(GOLD, n) for n in NAMES] + [str(tmp_path / "missing.mp3")]
    junk = tmp_path / "junk.mp3"
    junk.write_bytes(bytes(5000))
    files.insert(2, str(junk))
    out, err = cli.info_tsv(files, gain_modifier_db=1.5)
    exp = [cli.TSV_HEADER]
    for n in NAMES:
        img = np.fromfile(os.path.join(GOLD, n), np.uint8)
        rc, res, hist, info = oracle.analyze_mp3(img)
        rc2, st = oracle.gain_analyze(img)
        assert rc == 0 and rc2 == 0
        g = res.gain_db + 1.5
        exp.append("{}\t{}\t{:.6f}\t{:.6f}\t{}\t{}".format(n, _round_half_away(g / 1.5), g, res.peak * 32768.0, st.max_gain, st.min_gain))
    assert out == exp
    assert err == ["junk.mp3 - Failed to probe format: " + str(junk), "missing.mp3 - Failed to open: " + files[-1]]
    assert cli.main(["-o", "-d", "1.5"] + files) == 0
    cap = capsys.readouterr()
    assert cap.out == "\n".join(exp) + "\n" and cap.err == "\n".join(err) + "\n"


def test_track_and_album_lines(oracle):
    from mp3rgain_b200 import cli
    files = [os.path.join(GOLD, n) for n in NAMES[:3]]
    out, err, steps = cli.track_gain_lines(files, gain_modifier=1)
    assert out[:4] == ["[DRY RUN] mp3rgain Analyzing and would apply track gain to 3 file(s)", "  Target: 89 dB (ReplayGain 1.0)",
                       "  Gain modifier: +1 steps", ""] and not err
    hists, peaks = [], []
    for n, f in zip(NAMES, files):
        rc, res, hist, info = oracle.analyze_mp3(np.fromfile(f, np.uint8))
        base = _round_half_away(res.gain_db / 1.5)
        assert "  -> [DRY RUN] Analyzing %s..." % n in out
        assert "      Loudness: {:.1f} dB, Gain: {:+.1f} dB ({} steps + 1 = {}), Peak: {:.4f}".format(res.loudness_db, res.gain_db, base, base + 1, res.peak) in out
        assert steps[NAMES.index(n)] == base + 1
        hists.append(hist)
        peaks.append(res.peak)
    lines, asteps = cli.album_gain_lines(files)
    _, l, g, p = oracle.album_merge(np.stack(hists), np.array(peaks))
    assert lines[-5:] == ["", "  Album loudness: {:.1f} dB".format(l), "  Album gain:     {:+.1f} dB ({} steps)".format(g, _round_half_away(g / 1.5)),
                          "  Album peak:     {:.4f}".format(p), ""]
    assert asteps == _round_half_away(g / 1.5)


def test_resident_corpus_gives_the_same_answers_as_the_host_entry_points(engine_fast):
    """Engine.upload(): the images cross the bus once, rg_analyze_corpus and rg_gain_analyze_corpus both work from HBM."""
    images = [np.fromfile(os.path.join(GOLD, n), np.uint8) for n in NAMES] + [np.zeros(3000, np.uint8), np.zeros(0, np.uint8)]
    res, sts = engine_fast.analyze_images(images, want_hist=True)
    gst = engine_fast.gain_analyze_images(images)
    with engine_fast.upload(images) as cor:
        res2, sts2 = cor.analyze(want_hist=True)
        gst2 = cor.gain_analyze()
        res3, _ = cor.analyze()          # a second pass over the same resident images
    assert sts == sts2 and sts[:4] == [0, 0, 0, 0] and sts[4] != 0
    for a, b, c in zip(res, res2, res3):
        assert (a.loudness_db, a.gain_db, a.peak, a.n_frames) == (b.loudness_db, b.gain_db, b.peak, b.n_frames) == (c.loudness_db, c.gain_db, c.peak, c.n_frames)
        assert (a.histogram == b.histogram).all()
    for a, b in zip(gst, gst2):
        assert (a.status, a.frame_count, a.gain_sum, a.min_gain, a.max_gain) == (b.status, b.frame_count, b.gain_sum, b.min_gain, b.max_gain)


def test_json_reports_match_the_per_file_loop(oracle, tmp_path, capsys):
    """`-o json` for cmd_info, `-r` and `-a` (src/main.rs:101-165, 1119-1150, 1213-1290, 1296-1455): the documents built from ONE
    batched engine call each equal the documents the reference's loops would serialise from the oracle's numbers."""
    import json
    from mp3rgain_b200 import cli
    files = [os.path.join(GOLD, n) for n in NAMES[:3]]
    junk = tmp_path / "junk.mp3"
    junk.write_bytes(bytes(3000))
    doc = json.loads(cli.info_json(files + [str(junk)]))
    assert list(doc) == ["files"] and len(doc["files"]) == 4
    for f, d in zip(files, doc["files"]):
        rc, st = oracle.gain_analyze(np.fromfile(f, np.uint8))
        assert rc == 0
        assert d == {"file": f, "frames": st.frame_count, "mpeg_version": ["MPEG1", "MPEG2", "MPEG2.5"][st.mpeg_version],
                     "channel_mode": ["Stereo", "Joint Stereo", "Dual Channel", "Mono"][st.channel_mode], "min_gain": st.min_gain,
                     "max_gain": st.max_gain, "avg_gain": st.gain_sum / st.gain_count, "headroom_steps": 255 - st.max_gain,
                     "headroom_db": (255 - st.max_gain) * 1.5}
    assert doc["files"][3] == {"file": str(junk), "status": "error", "error": "No valid MP3 frames found"}

    doc = json.loads(cli.track_gain_json(files + [str(junk)], gain_modifier=1))
    hists, peaks = [], []
    for f, d in zip(files, doc["files"]):
        rc, res, hist, info = oracle.analyze_mp3(np.fromfile(f, np.uint8))
        steps = _round_half_away(res.gain_db / 1.5) + 1
        assert d["file"] == f and d["loudness_db"] == res.loudness_db and d["peak"] == res.peak and d["status"] == "dry_run" and d["dry_run"] is True
        assert d["gain_applied_steps"] == steps and d["gain_applied_db"] == steps * 1.5
        assert ("warning" in d) == (steps > 0 and res.peak * 10 ** (res.gain_db / 20) > 1.0)
        hists.append(hist)
        peaks.append(res.peak)
    assert doc["files"][3] == {"file": str(junk), "status": "error", "error": "Failed to probe format: " + str(junk)}
    assert doc["summary"] == {"total_files": 4, "successful": 0, "failed": 1, "dry_run": True}

    text, rc = cli.album_gain_json(files)
    doc = json.loads(text)
    _, l, g, p = oracle.album_merge(np.stack(hists), np.array(peaks))
    assert rc == 0 and doc["album"] == {"loudness_db": l, "gain_db": g, "gain_steps": _round_half_away(g / 1.5), "peak": p}
    assert [d["file"] for d in doc["files"]] == files and doc["summary"] == {"total_files": 3, "successful": 0, "failed": 0, "dry_run": True}
    text, rc = cli.album_gain_json(files + [str(junk)])      # the first failing file fails the album (src/main.rs:1428-1450)
    assert rc == 1 and json.loads(text) == {"summary": {"total_files": 4, "successful": 0, "failed": 4, "dry_run": True}}

    # -R over a directory, through main()
    d = tmp_path / "lib" / "sub"
    d.mkdir(parents=True)
    import shutil
    shutil.copy(files[0], d / "b.MP3")
    shutil.copy(files[1], tmp_path / "lib" / "a.mp3")
    (tmp_path / "lib" / "cover.jpg").write_bytes(b"x")
    assert cli.main(["-R", "-o", "json", str(tmp_path / "lib")]) == 0
    got = json.loads(capsys.readouterr().out)
    assert [x["file"] for x in got["files"]] == [str(tmp_path / "lib" / "a.mp3"), str(d / "b.MP3")]
