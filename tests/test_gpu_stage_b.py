"""GPU parity tests, stage B: rg_analyze_pcm_f32 (CUDA K4+K5 through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar: identical 12000-bin histogram, identical loudness/gain/peak (f64 ==)."""
import math

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu


def _check(engine, oracle, planes, sr):
    rc, ores, ohist = oracle.analyze_pcm(planes, sr)
    assert rc == 0
    g = engine.analyze_pcm_f32(planes, sr)
    assert (g.histogram == ohist).all(), np.nonzero(g.histogram != ohist)[0][:10]
    assert g.loudness_db == ores.loudness_db
    assert g.gain_db == ores.gain_db
    assert g.peak == ores.peak
    assert g.n_windows == ores.n_windows
    assert g.n_dropped_windows == ores.n_dropped
    return g


CASES = [(44100, 2, 88200, 12345, 2), (48000, 1, 72000, 777, 4), (32000, 2, 32320, 42, 0), (44100, 2, 4410, 1, 16),
         (22050, 2, 30000, 9, 3), (8000, 1, 8000, 3, 1), (96000, 2, 50000, 11, 5), (11025, 2, 12000, 4, 2)]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_lcg_noise_matches_oracle(case, mode, engine_exact, engine_fast, oracle):
    sr, ch, n, seed, shift = case
    planes = oracle_lib.lcg_pcm(n, seed, shift, ch)
    _check(engine_exact if mode == "exact" else engine_fast, oracle, planes, sr)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_reference_unit_test_sines(mode, engine_exact, engine_fast, oracle):
    # src/replaygain.rs:1298-1365 drives the filter with f64 sines; through the f32 PCM entry we feed the f32-rounded
    # sine and require equality with the oracle, plus the reference's own range assertions.
    eng = engine_exact if mode == "exact" else engine_fast
    for amp, lo, hi in ((0.5, 50.0, 100.0), (0.1, 50.0, 80.0)):
        x = np.array([amp * math.sin(2.0 * math.pi * 1000.0 * (i / 44100.0)) for i in range(44100)], np.float32)
        g = _check(eng, oracle, [x], 44100)
        assert lo < g.loudness_db < hi


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_edge_cases(mode, engine_exact, engine_fast, oracle):
    eng = engine_exact if mode == "exact" else engine_fast
    z = np.zeros(44100, np.float32)
    g = _check(eng, oracle, [z, z], 44100)          # silence: every window dropped, -20 dB
    assert g.loudness_db == -20.0 and g.n_dropped_windows == 20
    _check(eng, oracle, [np.zeros(0, np.float32)], 44100)          # empty input
    _check(eng, oracle, oracle_lib.lcg_pcm(7, 1, 0, 2), 44100)     # shorter than one window
    _check(eng, oracle, oracle_lib.lcg_pcm(2205, 2, 0, 1), 44100)  # exactly one window
    _check(eng, oracle, oracle_lib.lcg_pcm(2206, 2, 0, 1), 44100)  # one window + 1 sample
    big = (np.array(oracle_lib.lcg_pcm(30000, 8, 0, 1)[0]) * 9.0e5).astype(np.float32)  # > +100 dB: all dropped
    g = _check(eng, oracle, [big], 44100)
    assert g.n_dropped_windows == g.n_windows


def test_unsupported_rate_is_an_error(engine_exact):
    import mp3rgain_b200 as m
    with pytest.raises(m.ReplayGainError) as ei:
        engine_exact.analyze_pcm_f32([np.zeros(100, np.float32)], 99999)
    assert ei.value.status == -9
    assert "Unsupported sample rate: 99999 Hz" in str(ei.value)


def test_long_track_fast_equals_exact(engine_exact, engine_fast):
    """Full-size property: a 3-minute stereo track through the time-blocked kernel (with edge re-run) yields the same
    histogram as the exact serial kernel."""
    rng = np.random.default_rng(7)
    n = 44100 * 180
    t = np.arange(n, dtype=np.float64) / 44100.0
    env = 0.05 + 0.3 * (0.5 + 0.5 * np.sin(2 * np.pi * 0.11 * t))
    l = (env * (0.6 * np.sin(2 * np.pi * 220.0 * t) + 0.4 * rng.standard_normal(n))).astype(np.float32)
    r = (env * (0.6 * np.sin(2 * np.pi * 331.0 * t) + 0.4 * rng.standard_normal(n))).astype(np.float32)
    a = engine_exact.analyze_pcm_f32([l, r], 44100)
    b = engine_fast.analyze_pcm_f32([l, r], 44100)
    assert a.n_windows == 3600
    assert (a.histogram == b.histogram).all()
    assert a.loudness_db == b.loudness_db and a.peak == b.peak


@pytest.mark.parametrize("spike", [np.inf, -np.inf, 3.0e38, 2.0e7])
def test_caller_pcm_beyond_what_a_warm_up_window_forgets(spike, engine_exact, engine_fast, oracle):
    """rg_analyze_pcm_f32 takes whatever the caller holds.  One sample at +146 dBFS or beyond leaves the serial filter ringing
    for several 50 ms windows, and an infinity leaves it at NaN to the end of the track (every later window then lands in
    bin 2000: `NaN as i32` = 0) -- more than the single warm-up window of the time-blocked kernel can reproduce, so K5 flags
    such tracks and they are filtered again serially.  Found by tests/dev/simt_emu/fuzz_stage_b_emu.py (the decoder cannot
    produce such samples: |PCM| stays below 1e9 even with every global_gain at 255)."""
    l, r = oracle_lib.lcg_pcm(44100, 11, 0, 2)
    l = np.array(l, np.float32)
    l[9000] = spike
    for eng in (engine_exact, engine_fast):
        g = _check(eng, oracle, [l, np.array(r, np.float32)], 44100)
    assert g.exact_rerun == 1


def test_nan_samples_poison_the_rest_of_the_track_like_the_reference(engine_exact, engine_fast, oracle):
    x = np.array(oracle_lib.lcg_pcm(3 * 16000, 5, 0, 1)[0], np.float32)
    x[20000] = np.nan          # f64::max ignores the NaN for the peak; the filter state does not recover
    for eng in (engine_exact, engine_fast):
        g = _check(eng, oracle, [x], 16000)
        assert g.histogram[2000] == g.n_windows - 20000 // 800 and np.isfinite(g.peak)
