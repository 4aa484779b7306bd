#!/usr/bin/env python3
"""Differential fuzzing of the device decode path against the oracle: seeded random corruptions (bit flips in headers, side
info and payload, truncations, inserted junk) of valid synthetic streams of every MPEG version; status, frame count, PCM
and histogram must agree.  usage: fuzz_parity.py [n_cases=60] [seed=1]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("RG_SIMT_EMU") == "1":   # development aid: the interpreted kernel sources instead of a GPU (tests/dev/simt_emu)
    sys.path.insert(0, os.path.join(ROOT, "tests", "dev", "simt_emu"))
    import build_emu  # noqa: E402
    from mp3rgain_b200 import _lib  # noqa: E402
    _lib.LIB_PATH = os.environ.get("RG_SIMT_EMU_LIB") or build_emu.build()
import mp3rgain_b200 as m  # noqa: E402
import oracle_lib  # noqa: E402
import synth_corpus  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
orc = oracle_lib.load()
bases = [synth_corpus.make_track(seed=900 + i, n_frames=120, version=v, channels=c, br_idx=b, sr_idx=s, crc=cr)
         for i, (v, c, b, s, cr) in enumerate([(3, 2, 9, 0, False), (3, 1, 0, 1, False), (2, 2, 6, 0, False), (0, 2, 0, 2, False),
                                               (3, 2, 5, 2, True), (2, 1, 4, 1, True),
                                               # 320 kbps at 32 kHz (K1 walks its slice of the main data in several windows), 256 kbps
                                               (3, 2, 14, 2, False), (3, 2, 13, 0, False)])]
bad = 0
with m.Engine(device=0, exact=True) as eng:
    for case in range(n_cases):
        img = bases[case % len(bases)].copy()
        kind = rng.integers(0, 5)
        if kind == 0:      # bit flips anywhere
            for _ in range(int(rng.integers(1, 40))):
                p = int(rng.integers(0, len(img))); img[p] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:    # flips concentrated in headers / side info of random frames
            for _ in range(int(rng.integers(1, 12))):
                p = int(rng.integers(0, len(img) - 40)); img[p + int(rng.integers(0, 36))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:    # truncation
            img = img[: int(rng.integers(1, len(img)))]
        elif kind == 3:    # junk inserted
            p = int(rng.integers(0, len(img)))
            img = np.concatenate([img[:p], rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8), img[p:]])
        else:              # chunk removed
            p = int(rng.integers(0, len(img) - 10)); q = min(len(img), p + int(rng.integers(1, 2000)))
            img = np.concatenate([img[:p], img[q:]])
        rc, ores, ohist, info = orc.analyze_mp3(img)
        res, sts = eng.analyze_images([img], want_hist=True)
        ok = sts[0] == rc
        if ok and rc == 0:
            ok = res[0].n_frames == info.n_frames and (res[0].histogram == ohist).all() and res[0].peak == ores.peak
            if ok:
                rc2, opcm, _ = orc.decode(img)
                gpcm, sr = eng.decode_pcm(img)
                ok = gpcm.shape == opcm.shape and bool((gpcm == opcm).all())
        if not ok:
            bad += 1
            print("MISMATCH case", case, "kind", int(kind), "base", case % len(bases), "len", len(img), "status gpu/oracle", sts[0], rc,
                  "frames", res[0].n_frames if sts[0] == 0 else None, info.n_frames, "dropped", info.n_dropped_frames, flush=True)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.save(os.path.join(ROOT, "gpurun_out", f"fuzz_case_{seed}_{case}.npy"), img)
print("fuzz_parity: %d cases, %d mismatches" % (n_cases, bad))
sys.exit(1 if bad else 0)
