#!/usr/bin/env python3
"""Regenerate tests/golden/*.json.  Run in the build container only (needs FFmpeg's libavcodec via tests/ffmpeg_ref.py)."""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ffmpeg_ref  # noqa: E402
import oracle_lib  # noqa: E402

o = oracle_lib.load()
out = {}
for name in ("test_joint_stereo", "test_mono", "test_vbr", "test_stereo"):
    img = open(os.path.join(HERE, name + ".mp3"), "rb").read()
    ff = ffmpeg_ref.decode_frames(ffmpeg_ref.split_frames(img))
    rc, res, hist = o.analyze_pcm(list(ff), 44100)
    nz = np.nonzero(hist)[0]
    out[name] = {
        "decoder": "ffmpeg mp3float (libavcodec 62.11.100), flags2=+skip_manual, Info frame skipped",
        "channels": int(ff.shape[0]), "samples": int(ff.shape[1]),
        "loudness_db": res.loudness_db, "gain_db": round(res.gain_db, 2), "peak": res.peak,
        "n_windows": res.n_windows, "n_dropped": res.n_dropped,
        "histogram": {str(int(i)): int(hist[i]) for i in nz},
    }
json.dump(out, open(os.path.join(HERE, "fixtures_expected.json"), "w"), indent=1, sort_keys=True)

ka = {"sines": [], "lcg": []}
for amp in (0.5, 0.1):
    x = np.array([amp * 32768.0 * math.sin(2.0 * math.pi * 1000.0 * (i / 44100.0)) for i in range(44100)], np.float64)
    r = o.analyze_mono_raw_f64(x, 44100)
    ka["sines"].append({"amp": amp, "loudness_db": r.loudness_db, "n_windows": r.n_windows})
for sr, ch, n, seed, shift in ((44100, 2, 88200, 12345, 2), (48000, 1, 72000, 777, 4), (32000, 2, 32320, 42, 0), (44100, 2, 4410, 1, 16)):
    rc, r, hist = o.analyze_pcm(oracle_lib.lcg_pcm(n, seed, shift, ch), sr)
    nz = np.nonzero(hist)[0]
    ka["lcg"].append({"sr": sr, "ch": ch, "n": n, "seed": seed, "shift": shift, "loudness_db": r.loudness_db,
                      "n_windows": r.n_windows, "n_dropped": r.n_dropped, "peak": r.peak,
                      "histogram": {str(int(i)): int(hist[i]) for i in nz}})
json.dump(ka, open(os.path.join(HERE, "stage_b_known_answers.json"), "w"), indent=1, sort_keys=True)
print("ok")
