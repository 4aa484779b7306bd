#!/usr/bin/env python3
"""Differential fuzz of stage B (K4 exact and time-blocked + K5) on the interpreted kernel sources against the oracle: random
lengths around window multiples, all 12 rates, mono / stereo, level steps and sweeps (bin-edge hunting), silence gaps,
denormals, infinities and NaNs.  Development aid.   usage: fuzz_stage_b_emu.py [n_cases=300] [seed=1]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from mp3rgain_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.environ.get("RG_SIMT_EMU_LIB") or build_emu.build()
import mp3rgain_b200 as m  # noqa: E402
import oracle_lib  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
orc = oracle_lib.load()
RATES = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000]
bad = 0
with m.Engine(device=0, exact=True) as ex, m.Engine(device=0, exact=False) as fa:
    for case in range(n_cases):
        sr = RATES[int(rng.integers(0, 12))]
        win = sr * 50 // 1000
        nwin = int(rng.integers(0, 40))
        n = max(0, nwin * win + int(rng.choice([-1, 0, 1, 7, win // 2, int(rng.integers(0, win))])))
        nch = int(rng.integers(1, 3))
        t = np.arange(n) / sr
        planes = []
        for c in range(nch):
            kind = int(rng.integers(0, 6))
            x = rng.standard_normal(n) * 10 ** rng.uniform(-6, 0)
            if kind == 1:      # level sweep: walks the window levels across many bin edges
                x = x * np.logspace(rng.uniform(-3, 0), rng.uniform(-3, 0), n) if n else x
            elif kind == 2:    # tone + silence gaps
                x = 10 ** rng.uniform(-4, 0) * np.sin(2 * np.pi * rng.uniform(20, sr / 2.1) * t)
                if n:
                    a, b = sorted(rng.integers(0, n, 2)); x[a:b] = 0.0
            elif kind == 3:    # denormal / tiny levels (windows below the histogram)
                x = x * 1e-30
            elif kind == 4:    # far above full scale (windows above the histogram), with a step
                x = x * 10 ** rng.uniform(2, 8)
                if n: x[: int(rng.integers(0, n))] *= 1e-6
            x = x.astype(np.float32)
            if kind == 5 and n:  # non-finite samples
                for _ in range(int(rng.integers(1, 4))):
                    x[int(rng.integers(0, n))] = rng.choice(np.array([np.inf, -np.inf, np.nan, 3.0e38, -3.0e38], np.float32))
            planes.append(x)
        rc, ores, ohist = orc.analyze_pcm(planes, sr)
        for name, eng in (("exact", ex), ("fast", fa)):
            g = eng.analyze_pcm_f32(planes, sr)
            same_peak = g.peak == ores.peak or (np.isnan(g.peak) and np.isnan(ores.peak))
            if not ((g.histogram == ohist).all() and g.loudness_db == ores.loudness_db and same_peak and g.n_windows == ores.n_windows
                    and g.n_dropped_windows == ores.n_dropped):
                bad += 1
                print("MISMATCH case", case, name, "sr", sr, "n", n, "nch", nch, "loud", g.loudness_db, ores.loudness_db, "peak", g.peak, ores.peak,
                      "windows", g.n_windows, ores.n_windows, "dropped", g.n_dropped_windows, ores.n_dropped,
                      "hist diff bins", np.nonzero(g.histogram != ohist)[0][:6], flush=True)
                os.makedirs("/tmp/fuzz_emu", exist_ok=True)
                np.save(f"/tmp/fuzz_emu/stageb_{seed}_{case}.npy", np.stack(planes) if n else np.zeros((nch, 0), np.float32))
print("fuzz_stage_b_emu: %d cases, %d mismatches (seed %d)" % (n_cases, bad, seed))
sys.exit(1 if bad else 0)
