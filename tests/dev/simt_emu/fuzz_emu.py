#!/usr/bin/env python3
"""Differential fuzzing of the KERNEL SOURCES (run by the CPU SIMT interpreter of this directory) against the oracle, for use
in the GPU-less container: seeded corruptions of valid synthetic streams of every MPEG version through the analysis path
(status, frame count, histogram, peak, PCM) and through the global_gain walk (Mp3Analysis fields, patched image).
Development aid; nothing here is a product path.   usage: fuzz_emu.py [n_cases=200] [seed=1] [frames=48]"""
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
import synth_corpus  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 48
rng = np.random.default_rng(seed)
orc = oracle_lib.load()
# (version, channels, bitrate index, sample-rate index, crc): MPEG-1 128k / 32k mono, MPEG-2 48k, MPEG-2.5 8k (24-byte-class
# frames), MPEG-1 with CRC, MPEG-2 mono with CRC, MPEG-2 8 kbps at 24 kHz (the smallest Layer III frame)
CONFIGS = [(3, 2, 9, 0, False), (3, 1, 1, 1, False), (2, 2, 6, 0, False), (0, 2, 1, 2, False), (3, 2, 5, 2, True), (2, 1, 4, 1, True),
           (2, 1, 1, 1, False), (0, 1, 1, 0, False),
           # 320 kbps at 32 kHz (the largest frames: K1 walks its slice of the main data in several windows) and VBR
           (3, 2, 14, 2, False), (3, 2, 0, 0, False)]
bases = []
for i, (v, c, b, s, cr) in enumerate(CONFIGS):
    try:
        bases.append(synth_corpus.make_track(seed=700 + i + 16 * seed, n_frames=n_frames, version=v, channels=c, br_idx=b, sr_idx=s, crc=cr))
    except Exception as ex:  # a configuration the synthesiser cannot fit (payload smaller than the side info)
        print("skipping config", CONFIGS[i], ex)


def corrupt(img):
    kind = int(rng.integers(0, 9))
    img = img.copy()
    if kind == 0:
        for _ in range(int(rng.integers(1, 40))):
            img[int(rng.integers(0, len(img)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:
        for _ in range(int(rng.integers(1, 12))):
            p = int(rng.integers(0, max(1, len(img) - 40)))
            img[min(len(img) - 1, p + int(rng.integers(0, 36)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 2:
        img = img[: int(rng.integers(1, len(img)))]
    elif kind == 3:
        p = int(rng.integers(0, len(img)))
        img = np.concatenate([img[:p], rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8), img[p:]])
    elif kind == 4:
        p = int(rng.integers(0, len(img) - 10)); q = min(len(img), p + int(rng.integers(1, 2000)))
        img = np.concatenate([img[:p], img[q:]])
    elif kind == 5:    # sync-like junk: 0xFF-heavy bytes, which manufacture false syncs for both walks
        p = int(rng.integers(0, len(img)))
        junk = rng.choice(np.array([0xFF, 0xFB, 0xE3, 0xF3, 0x90, 0x00, 0x64], dtype=np.uint8), int(rng.integers(1, 600)))
        img = np.concatenate([img[:p], junk, img[p:]])
    elif kind == 7:    # an ID3v2 tag in front: honest, short, longer than the file, with the footer flag, with non-syncsafe size bytes
        body = int(rng.integers(0, 700))
        claim = body if rng.integers(0, 3) else int(rng.integers(0, 2 * len(img) + 4000))
        flags = int(rng.choice([0x00, 0x10, 0x40, 0x80]))
        size = [(claim >> 21) & 0x7F, (claim >> 14) & 0x7F, (claim >> 7) & 0x7F, claim & 0x7F]
        if rng.integers(0, 8) == 0:
            size[int(rng.integers(0, 4))] |= 0x80
        tag = np.array([0x49, 0x44, 0x33, int(rng.choice([2, 3, 4])), 0, flags] + size, dtype=np.uint8)
        fill = rng.choice(np.array([0x00, 0xFF, 0xFB, 0x54, 0x90], dtype=np.uint8), body)
        img = np.concatenate([tag, fill, np.zeros(10 if flags & 0x10 else 0, np.uint8), img[int(rng.integers(0, 3)) * int(rng.integers(0, 50)):]])
    elif kind == 8:    # trailers: ID3v1, APEv2 (+ ID3v1), with honest and dishonest sizes, possibly after a cut
        if rng.integers(0, 2):
            img = img[: int(rng.integers(len(img) // 2, len(img) + 1))]
        parts = [img]
        if rng.integers(0, 2):
            n_items = int(rng.integers(0, 300))
            sz = n_items + 32 if rng.integers(0, 4) else int(rng.integers(0, 2 * len(img)))
            foot = np.zeros(32, np.uint8)
            foot[:8] = np.frombuffer(b"APETAGEX", np.uint8)
            foot[8:12] = np.frombuffer(int(2000).to_bytes(4, "little"), np.uint8)
            foot[12:16] = np.frombuffer(int(sz).to_bytes(4, "little"), np.uint8)
            foot[20:24] = np.frombuffer(int(rng.choice([0, 0x80000000, 0xA0000000])).to_bytes(4, "little"), np.uint8)
            hdr = foot.copy()
            parts += ([hdr] if rng.integers(0, 2) else []) + [rng.choice(np.array([0x00, 0xFF, 0xFB, 0x41], dtype=np.uint8), n_items), foot]
        if rng.integers(0, 3):
            t = rng.choice(np.array([0x00, 0x20, 0xFF, 0xFB, 0x61], dtype=np.uint8), 128)
            t[:3] = np.frombuffer(b"TAG", np.uint8)
            parts.append(t)
        img = np.concatenate(parts)
    else:              # a second stream of another version spliced behind a cut
        other = bases[int(rng.integers(0, len(bases)))]
        p = int(rng.integers(1, len(img))); q = int(rng.integers(0, len(other)))
        img = np.concatenate([img[:p], other[q:]])
    return kind, np.ascontiguousarray(img)


bad = 0
recent = []      # (image, status, histogram, peak, gain stats) of the last cases, for the batch check
with m.Engine(device=0, exact=True) as eng, m.Engine(device=0, exact=False) as fast:
    fast.set_option("group_bytes_min", 4096)     # several copy / index groups even for these small corpora
    for case in range(n_cases):
        base = case % len(bases)
        kind, img = corrupt(bases[base])
        why = None
        rc, ores, ohist, info = orc.analyze_mp3(img)
        res, sts = eng.analyze_images([img], want_hist=True)
        if sts[0] != rc:
            why = f"status {sts[0]} vs {rc}"
        elif rc == 0:
            if res[0].n_frames != info.n_frames:
                why = f"frames {res[0].n_frames} vs {info.n_frames}"
            elif not (res[0].histogram == ohist).all() or res[0].peak != ores.peak:
                why = "histogram / peak"
            else:
                _, opcm, _ = orc.decode(img)
                gpcm, _sr = eng.decode_pcm(img)
                if gpcm.shape != opcm.shape or not bool((gpcm == opcm).all()):
                    why = "pcm"
                else:   # find_peak_amplitude: the peak-only pipeline (K2 in peak mode, no PCM stored) against max |sample|
                    (pk, pst), = eng.find_peak_images([img])
                    want = float(np.abs(opcm).max()) if opcm.size else 0.0
                    if pst != 0 or pk.peak != want or pk.peak != ores.peak or pk.sample_rate != info.sample_rate:
                        why = f"find_peak {pst} {pk.peak} vs {want}"
        if why is None:
            (gs,) = eng.gain_analyze_images([img])
            grc, ost = orc.gain_analyze(img)
            grc = -12 if grc == -1 else grc    # the oracle's -1 is "No valid MP3 frames found" = RG_ERR_NO_FRAMES
            if gs.status != grc:
                why = f"gain status {gs.status} vs {grc}"
            elif grc == 0 and (gs.frame_count, gs.gain_sum, gs.gain_count, gs.min_gain, gs.max_gain) != (
                    ost.frame_count, ost.gain_sum, ost.gain_count, ost.min_gain, ost.max_gain):
                why = "Mp3Analysis fields"
            else:
                steps = int(rng.choice([-60, -3, -1, 1, 2, 7, 90]))
                mode = int(rng.integers(0, 3))
                if mode == 2:
                    ch = int(rng.integers(0, 2))
                    (patched,), (nmod,) = eng.gain_apply_images([img], steps, channel=ch)
                    oimg, on = orc.gain_apply(img, steps, channel=ch)
                else:
                    (patched,), (nmod,) = eng.gain_apply_images([img], steps, wrap=bool(mode))
                    oimg, on = orc.gain_apply(img, steps, wrap=bool(mode))
                if nmod != on or not (patched == oimg).all():
                    why = f"patch steps={steps} mode={mode}: {nmod} vs {on}"
        if why is None:
            recent.append((img, sts[0], res[0].histogram.copy() if sts[0] == 0 else None, res[0].peak if sts[0] == 0 else None))
            if len(recent) == 7:
                # the same images side by side in one call of the time-blocked engine (pipelined groups, shared batches):
                # every track must come out as it did alone -- corrupt neighbours must not leak into each other
                imgs = [r[0] for r in recent]
                bres, bsts = fast.analyze_images(imgs, want_hist=True)
                for k, r in enumerate(recent):
                    if bsts[k] != r[1] or (r[1] == 0 and (not (bres[k].histogram == r[2]).all() or bres[k].peak != r[3])):
                        why = f"batch entry {k} differs from the same image analysed alone"
                        np.save(f"/tmp/fuzz_emu_batch_{seed}_{case}_{k}.npy", r[0])
                gb = fast.gain_analyze_images(imgs)
                for k, r in enumerate(recent):
                    grc1, ost1 = orc.gain_analyze(r[0])
                    if gb[k].status != (-12 if grc1 == -1 else grc1) or (grc1 == 0 and (gb[k].frame_count, gb[k].gain_sum) != (ost1.frame_count, ost1.gain_sum)):
                        why = f"batch gain entry {k}"
                recent.clear()
        if why:
            bad += 1
            print("MISMATCH case", case, "kind", kind, "base", base, "len", len(img), why, flush=True)
            if os.environ.get("FUZZ_DUMP"): img.tofile(os.path.join(os.environ["FUZZ_DUMP"], "fuzz_fail_%d_%d.mp3" % (seed, case)))
            os.makedirs("/tmp/fuzz_emu", exist_ok=True)
            np.save(f"/tmp/fuzz_emu/case_{seed}_{case}.npy", img)
print("fuzz_emu: %d cases, %d mismatches (seed %d)" % (n_cases, bad, seed))
sys.exit(1 if bad else 0)
