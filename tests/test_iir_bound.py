"""CPU test: the bound on the time-blocked IIR (K4's default mode).  tests/dev/iir_bound.c evaluates the cascade the way
k_iir_fast does (transposed direct form II, FMAs, chains restarted from zero state one window early -- two at 24 kHz and
below, the engine's rule) next to the oracle's serial direct-form-I filter, window by window, over stress signals at the edge
of what a Layer III decoder can emit (|PCM| up to 3e4 = global_gain 255 throughout, +-DC steps, full-scale 20 Hz tones, a
+90 dBFS burst one and two windows before a block start, loud / near-silent alternation), and reports the largest
|val_blocked - val_serial| relative to the tolerance with which K5 flags a window for the serial re-run
(edge_tol * max(1, 0.1 * 32768 * loudest of this window and the three before it / filtered RMS)).  The engine's claim --
a window whose bin could differ from the reference's is always flagged -- holds while that ratio stays below 1; this test
fails when the margin drops under 10x at any of the nine MP3 sample rates.  (Reference: src/replaygain.rs:586-616, 743-765.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RATES = [44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("gcc") is None or os.uname().machine != "x86_64":
        pytest.skip("needs gcc on x86-64 (-mfma)")
    exe = str(tmp_path_factory.mktemp("iir") / "iir_bound")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "oracle"), "-o", exe,
                    os.path.join(ROOT, "tests", "dev", "iir_bound.c"), os.path.join(ROOT, "oracle", "rg_oracle.c"), "-lm"], check=True)
    return exe


def _run(exe, rate, block_windows, warm):
    out = subprocess.run([exe, str(rate), str(block_windows), str(warm)], capture_output=True, text=True, check=True).stdout
    rows = [ln.split() for ln in out.splitlines()]
    return {int(r[0]): (float(r[1]), float(r[2]), int(r[3])) for r in rows}


def engine_warm_windows(rate):
    return 2 if rate <= 24000 else 1      # rg_engine.cu: build_chains


@pytest.mark.parametrize("rate", RATES)
def test_blocked_iir_stays_ten_times_inside_the_flag_tolerance(harness, rate):
    for bw in (8, 16, 38):
        res = _run(harness, rate, bw, engine_warm_windows(rate))
        assert len(res) == 10
        worst = max(v[1] for v in res.values())
        assert worst < 0.1, (rate, bw, res)


def test_one_warm_up_window_is_not_enough_below_32_khz(harness):
    """What the rule is for: with a single warm-up window the 8 kHz chains (400 samples per window, pole radius 0.946) are off
    by more than the tolerance on band-limited noise -- unflagged windows could land in the wrong bin."""
    res = _run(harness, 8000, 16, 1)
    assert max(v[1] for v in res.values()) > 1.0
