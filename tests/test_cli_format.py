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
    assert cli.main(["-r", "x.mp3"]) == 1 and cli.main([]) == 1
