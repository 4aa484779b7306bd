"""CPU tests: pin the oracle's stage B (oracle/rg_oracle.c) to everything the reference's own tests hold for this
path (src/replaygain.rs:1275-1365) and to the exact known answers of SURVEY.md Appendix B (derived independently
from the reference's literal operation order)."""
import math

import numpy as np
import pytest

import oracle_lib

RATES = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000]


def test_filter_creation_all_rates(oracle):
    # src/replaygain.rs:1275-1294 test_filter_creation
    import ctypes as C
    buf = C.create_string_buffer(4096)
    for r in RATES:
        assert oracle.lib.rgo_filter_init(buf, r) == 0, r
    assert oracle.lib.rgo_filter_init(buf, 99999) == -1


def _sine(amp):
    sr = 44100
    return np.array([amp * math.sin(2.0 * math.pi * 1000.0 * (i / sr)) for i in range(sr)], np.float64)


def test_rms_calculation_reference_unit_test(oracle):
    # src/replaygain.rs:1298-1332: 1 kHz, amplitude 0.5*32768, 1 s mono -> 50 < loudness < 100
    res = oracle.analyze_mono_raw_f64(_sine(0.5 * 32768.0), 44100)
    assert 50.0 < res.loudness_db < 100.0
    # exact value (Appendix B): 20 windows, all in bin 9297
    assert res.n_windows == 20 and res.n_dropped == 0
    assert res.loudness_db == 72.97
    assert abs(res.gain_db - (-8.15)) < 1e-9


def test_loudness_calculation_reference_unit_test(oracle):
    # src/replaygain.rs:1336-1365: amplitude 0.1*32768 -> 50 < loudness < 80
    res = oracle.analyze_mono_raw_f64(_sine(0.1 * 32768.0), 44100)
    assert 50.0 < res.loudness_db < 80.0
    assert res.n_windows == 20
    assert res.loudness_db == 58.99


LCG_CASES = [
    # sr, ch, n, seed, shift -> windows, dropped, peak, lo bin, hi bin, loudness, gain
    (44100, 2, 88200, 12345, 2, 40, 0, 0.25, 8383, 8471, 64.67, 0.15),
    (48000, 1, 72000, 777, 4, 30, 0, 0.0625, 7131, 7230, 52.26, 12.56),
    (32000, 2, 32320, 42, 0, 21, 0, 1.0, 9744, 9796, 77.90, -13.08),
    (44100, 2, 4410, 1, 16, 2, 0, None, 493, 500, -15.00, 79.82),
]


@pytest.mark.parametrize("case", LCG_CASES)
def test_lcg_noise_known_answers(oracle, case):
    sr, ch, n, seed, shift, nwin, ndrop, peak, lo, hi, loud, gain = case
    planes = oracle_lib.lcg_pcm(n, seed, shift, ch)
    rc, res, hist = oracle.analyze_pcm(planes, sr)
    assert rc == 0
    assert res.n_windows == nwin and res.n_dropped == ndrop
    nz = np.nonzero(hist)[0]
    assert nz.min() == lo and nz.max() == hi
    assert hist.sum() == nwin - ndrop
    assert res.loudness_db == loud
    assert abs(res.gain_db - gain) < 1e-9
    if peak is not None:
        assert res.peak == peak


def test_silence_drops_every_window(oracle):
    # finish_window: idx < 0 -> not counted (src/replaygain.rs:755-759); empty histogram -> -20 (:667-669)
    z = np.zeros(44100, np.float32)
    rc, res, hist = oracle.analyze_pcm([z, z], 44100)
    assert rc == 0 and hist.sum() == 0
    assert res.n_windows == 20 and res.n_dropped == 20
    assert res.loudness_db == -20.0 and abs(res.gain_db - 84.82) < 1e-12


def test_percentile_threshold_uses_f64_product(oracle):
    # ceil(total * (1.0 - 0.95)): 1.0-0.95 = 0.050000000000000044 so total=20 -> 2, 40 -> 3, 3600 -> 181
    for total, thr in ((20, 2), (40, 3), (3600, 181), (1, 1), (19, 1)):
        h = np.zeros(12000, np.uint32)
        h[5000:5000 + total] = 1  # one window per bin, ascending
        # the answer is the bin holding the thr-th loudest window
        want = (5000 + total - thr - 2000) / 100.0
        assert oracle.hist_loudness(h) == want, total
    assert oracle.hist_loudness(np.zeros(12000, np.uint32)) == -20.0


def test_unsupported_rate(oracle):
    rc, res, hist = oracle.analyze_pcm([np.zeros(10, np.float32)], 99999)
    assert rc == -1


def test_gain_steps_round_half_away(oracle):
    # ReplayGainResult::gain_steps (src/replaygain.rs:72-74); Rust round() is half away from zero
    assert oracle.lib.rgo_gain_steps(0.75) == 1
    assert oracle.lib.rgo_gain_steps(-0.75) == -1
    assert oracle.lib.rgo_gain_steps(6.71) == 4
    assert oracle.lib.rgo_gain_steps(84.82) == 57
    assert oracle.lib.rgo_gain_steps(-8.15) == -5


def test_album_merge_is_sum_of_histograms(oracle):
    planes_a = oracle_lib.lcg_pcm(44100, 5, 2, 2)
    planes_b = oracle_lib.lcg_pcm(44100, 6, 6, 2)
    _, ra, ha = oracle.analyze_pcm(planes_a, 44100)
    _, rb, hb = oracle.analyze_pcm(planes_b, 44100)
    merged, loud, gain, peak = oracle.album_merge(np.stack([ha, hb]), [ra.peak, rb.peak])
    assert (merged == ha + hb).all()
    assert peak == max(ra.peak, rb.peak)
    assert loud == oracle.hist_loudness((ha + hb).astype(np.uint64))
    assert abs(gain - (64.82 - loud)) < 1e-12
