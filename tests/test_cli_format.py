"""CPU tests of the CLI glue's formatting (mp3rgain_b200/cli.py) against the reference's format strings
(src/main.rs:1123, 1719-1722, 1957-1968, 1322-1338).  No GPU: pure string functions."""
from mp3rgain_b200 import cli


def test_tsv_header_and_line():
    assert cli.TSV_HEADER == "File\tMP3 gain\tdB gain\tMax Amplitude\tMax global_gain\tMin global_gain"
    # println!("{}\t{}\t{:.6}\t{:.6}\t{}\t{}", filename, gain_steps, gain_db, peak * 32768.0, max_gain, min_gain)
    assert cli.format_tsv_line("a.mp3", 6.71, 0.5, 210, 110) == "a.mp3\t4\t6.710000\t16384.000000\t210\t110"
    assert cli.format_tsv_line("b c.mp3", -0.76, 1.0000305, 255, 0) == "b c.mp3\t-1\t-0.760000\t32768.999424\t255\t0"
    assert cli.format_tsv_line("x", 0.75, 0.0, 1, 1).split("\t")[1] == "1"       # Rust round(): half away from zero
    assert cli.format_tsv_line("x", -0.75, 0.0, 1, 1).split("\t")[1] == "-1"
    assert cli.format_tsv_line("x", 0.74, 0.0, 1, 1).split("\t")[1] == "0"


def test_track_and_album_lines():
    assert cli.format_track_line(58.11, 6.71, 4, 0, 0.98765) == "      Loudness: 58.1 dB, Gain: +6.7 dB (4 steps), Peak: 0.9877"
    assert cli.format_track_line(70.0, -5.18, -3, 2, 1.2) == "      Loudness: 70.0 dB, Gain: -5.2 dB (-3 steps + 2 = -1), Peak: 1.2000"
    assert cli.format_album_lines(61.12, 3.7, 2, 0, 0.5) == ["", "  Album loudness: 61.1 dB", "  Album gain:     +3.7 dB (2 steps)",
                                                             "  Album peak:     0.5000", ""]
    assert cli.format_album_lines(61.12, 3.7, 2, -1, 0.5)[2] == "  Album gain:     +3.7 dB (2 steps + -1 = 1)"


def test_helpers():
    assert cli.get_filename("/a/b/c.mp3") == "c.mp3" and cli.get_filename("c.mp3") == "c.mp3"
    assert cli._rust_f64(89.0) == "89" and cli._rust_f64(64.82) == "64.82"
    assert cli.main([]) == 1 and cli.main(["-x", "a.mp3"]) == 1
    assert cli.main(["-s", "d", "a.mp3"]) == 1       # tag modes stay with the crate


def test_json_envelope_is_serde_json_pretty():
    """JsonOutput / JsonFileResult / JsonAlbumResult / JsonSummary (src/main.rs:101-165) through serde_json::to_string_pretty:
    declaration order, None fields skipped, two-space indent, f64 always with a fraction or an exponent."""
    f = cli.json_file_result("dir/a \"q\".mp3", status="dry_run", loudness_db=58.11, peak=0.5, gain_applied_steps=4, gain_applied_db=6.0,
                             warning="clipping warning: peak would be 1.23 (>1.00)", dry_run=True)
    assert list(f) == ["file", "status", "gain_applied_steps", "gain_applied_db", "loudness_db", "peak", "warning", "dry_run"]
    doc = cli.json_output(files=[f], album={"loudness_db": 61.1, "gain_db": 3.0, "gain_steps": 2, "peak": 1e-7}, summary=cli.json_summary(1, 0, 0, True))
    assert doc == """{
  "files": [
    {
      "file": "dir/a \\"q\\".mp3",
      "status": "dry_run",
      "gain_applied_steps": 4,
      "gain_applied_db": 6.0,
      "loudness_db": 58.11,
      "peak": 0.5,
      "warning": "clipping warning: peak would be 1.23 (>1.00)",
      "dry_run": true
    }
  ],
  "album": {
    "loudness_db": 61.1,
    "gain_db": 3.0,
    "gain_steps": 2,
    "peak": 1e-7
  },
  "summary": {
    "total_files": 1,
    "successful": 0,
    "failed": 0,
    "dry_run": true
  }
}"""
    assert cli.json_output(summary=cli.json_summary(3, 0, 3, False)) == '{\n  "summary": {\n    "total_files": 3,\n    "successful": 0,\n    "failed": 3\n  }\n}'
    assert cli.json_output(files=[]) == '{\n  "files": []\n}'
    assert [cli._json_f64(x) for x in (0.0, 89.0, -2.5e-9, 1e22, 64.82)] == ["0.0", "89.0", "-2.5e-9", "1e22", "64.82"]
    import json
    assert json.loads(doc)["files"][0]["file"] == 'dir/a "q".mp3'


def test_clipping_check_messages():
    """src/main.rs:2030-2075"""
    assert cli.clipping_check(4, 6.0, 0.4) == (4, None)                                    # 0.4 * 10^(6/20) = 0.798
    assert cli.clipping_check(4, 6.0, 0.9) == (4, "clipping warning: peak would be 1.80 (>1.00)")
    assert cli.clipping_check(4, 6.0, 0.9, ignore_clipping=True) == (4, None)
    assert cli.clipping_check(4, 6.0, 0.9, prevent_clipping=True) == (1, "gain reduced from 4 to 1 steps to prevent clipping (peak: 0.9000)")
    assert cli.clipping_check(-2, -3.0, 2.0) == (-2, None) and cli.clipping_check(4, 6.0, 0.9, wrap_gain=True) == (4, None)


def test_expand_files_recursive(tmp_path):
    """src/main.rs:436-470: directories walked at every level, mp3 / m4a / aac / mp4 in any case, other paths kept, sorted."""
    (tmp_path / "b" / "deep").mkdir(parents=True)
    (tmp_path / "a").mkdir()
    for rel in ("b/x.MP3", "b/deep/y.m4a", "b/notes.txt", "a/z.aac", "a/.mp3", "a/v.Mp4", "top.mp3"):
        (tmp_path / rel).write_bytes(b"")
    got = cli.expand_files_recursive([str(tmp_path / "top.mp3"), str(tmp_path / "b"), str(tmp_path / "a"), "missing.flac"])
    want = sorted([str(tmp_path / r) for r in ("top.mp3", "b/x.MP3", "b/deep/y.m4a", "a/z.aac", "a/v.Mp4")] + ["missing.flac"],
                  key=lambda q: tuple(q.split("/")))
    assert got == want
