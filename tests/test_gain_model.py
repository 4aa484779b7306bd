"""CPU test of the ALGORITHM behind k_gain_fields.cuh at scale: tests/dev/gf_kernel_model.c (a sequential C model of the
speculative sub-segment walk, the entry-position merge and the patch pass, same arithmetic as the kernels) against
oracle/gain_fields_oracle.c on thousands of random, corrupted and adversarial images.  The bytes after each image are
garbage on purpose: the walk may load them, no result may depend on them.

    python tests/test_gain_model.py 200000      # the long campaign (seeded), run once per kernel-logic change
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_lib  # noqa: E402
import synth_corpus  # noqa: E402

SRC = os.path.join(HERE, "dev", "gf_kernel_model.c")
LIB = os.path.join(HERE, "dev", "libgf_kernel_model.so")


class Track(C.Structure):
    _fields_ = [("frame_count", C.c_uint64), ("gain_sum", C.c_uint64), ("gain_count", C.c_uint64),
                ("first_hdr", C.c_uint32), ("gminmax", C.c_uint32), ("rewalks", C.c_uint32)]


def _model():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-Wall", "-o", LIB, SRC], check=True, capture_output=True)
    lib = C.CDLL(LIB)
    lib.gfm_run.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Track)]
    return lib


def _bases():
    cfgs = [(3, 2, 9, 0, False), (3, 1, 0, 1, False), (2, 2, 6, 0, True), (0, 2, 0, 2, False), (2, 1, 1, 2, False), (3, 2, 1, 1, True), (0, 1, 1, 0, False)]
    return [synth_corpus.make_track(seed=700 + i, n_frames=90, version=v, channels=c, br_idx=b, sr_idx=s, crc=cr)
            for i, (v, c, b, s, cr) in enumerate(cfgs)]


def _mutate(rng, bases):
    kind = int(rng.integers(0, 9))
    img = bases[int(rng.integers(0, len(bases)))].copy()
    if kind == 0:      # pure noise rich in sync-like bytes
        img = rng.integers(0, 256, int(rng.integers(0, 30000)), dtype=np.uint8)
        img[rng.integers(0, max(len(img), 1), len(img) // 6)] = 0xFF if len(img) else 0
    elif kind == 1:    # bit flips
        for _ in range(int(rng.integers(1, 300))):
            img[int(rng.integers(0, len(img)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 2:    # planted headers whose frame length points at another planted sync
        for _ in range(int(rng.integers(1, 80))):
            p = int(rng.integers(0, len(img) - 4))
            img[p:p + 4] = [0xFF, int(rng.choice([0xFB, 0xFA, 0xF3, 0xF2, 0xE3, 0xE2])), int(rng.integers(0x10, 0xF0)) & 0xF6, int(rng.integers(0, 256))]
    elif kind == 3:    # truncation
        img = img[: int(rng.integers(0, len(img)))]
    elif kind == 4:    # junk inserted (crosses sub-segment boundaries)
        p = int(rng.integers(0, len(img)))
        junk = rng.integers(0, 256, int(rng.integers(1, 12000)), dtype=np.uint8)
        junk[rng.integers(0, len(junk), len(junk) // 20 + 1)] = 0xFF
        img = np.concatenate([img[:p], junk, img[p:]])
    elif kind == 5:    # chunk removed
        p = int(rng.integers(0, len(img) - 10))
        img = np.concatenate([img[:p], img[min(len(img), p + int(rng.integers(1, 6000))):]])
    elif kind == 6:    # tags: ID3v2 (sometimes lying about its size), Info frame, ID3v1, APE footer with random size / flags
        parts = []
        if rng.random() < 0.7:
            tag = bytearray(synth_corpus.id3v2_tag(int(rng.integers(0, 9000))))
            if rng.random() < 0.3:
                tag[6:10] = bytes(int(x) for x in rng.integers(0, 256, 4))
            parts.append(np.frombuffer(bytes(tag), np.uint8))
        if rng.random() < 0.5 and img[1] & 0x18 == 0x18:
            parts.append(np.frombuffer(synth_corpus.info_frame(bytes(img[:4]), b"Xing" if rng.random() < 0.5 else b"Info"), np.uint8))
        parts.append(img)
        if rng.random() < 0.5:
            ape = b"APETAGEX" + (2000).to_bytes(4, "little") + int(rng.integers(0, 70000)).to_bytes(4, "little") + bytes(4) + \
                  (int(rng.integers(0, 2)) << 31).to_bytes(4, "little") + bytes(8)
            parts.append(np.frombuffer(bytes(int(x) for x in rng.integers(0, 256, int(rng.integers(0, 600)))) + ape, np.uint8))
        if rng.random() < 0.5:
            parts.append(np.frombuffer(synth_corpus.id3v1_trailer(), np.uint8))
        img = np.concatenate(parts)
    elif kind == 7:    # all 0xFF / long 0xFF runs
        p = int(rng.integers(0, len(img)))
        img[p:p + int(rng.integers(1, 9000))] = 0xFF
    # kind 8: untouched stream
    return np.ascontiguousarray(img, np.uint8)


def run_cases(n_cases, seed):
    orc, lib = oracle_lib.load(), _model()
    rng = np.random.default_rng(seed)
    bases = _bases()
    rewalks = 0
    for case in range(n_cases):
        img = _mutate(rng, bases)
        n = len(img)
        buf = np.empty(n + 64, np.uint8)
        buf[:n] = img
        buf[n:] = rng.integers(0, 256, 64, dtype=np.uint8)      # garbage slack
        t = Track()
        lib.gfm_run(buf.ctypes.data, n, 0, 0, 0, -1, C.byref(t))
        rc, o = orc.gain_analyze(img)
        ok = t.frame_count == o.frame_count and (buf[:n] == img).all()
        if ok and rc == 0:
            ver = (t.first_hdr >> 19) & 3
            ok = (t.gain_sum, t.gain_count, t.gminmax & 0xFF, t.gminmax >> 8) == (o.gain_sum, o.gain_count, o.min_gain, o.max_gain) and \
                 (0 if ver == 3 else (1 if ver == 2 else 2), (t.first_hdr >> 6) & 3) == (o.mpeg_version, o.channel_mode)
        if ok:
            steps, wrap, ch = int(rng.integers(-300, 300)) or 5, bool(rng.integers(0, 2)), int(rng.integers(-1, 2))
            lib.gfm_run(buf.ctypes.data, n, 1, steps, 1 if wrap else 0, ch, C.byref(t))
            ref, nmod = orc.gain_apply(img, steps, wrap, ch)
            ok = t.frame_count == nmod and (buf[:n] == ref).all()
        rewalks += t.rewalks
        assert ok, ("model != oracle", seed, case, n)
    return rewalks


def test_model_equals_oracle_on_random_and_adversarial_images():
    assert run_cases(2500, 20261017) > 0      # and the campaign did exercise the re-walk path


if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    done, seed = 0, 1
    while done < total:
        rw = run_cases(min(5000, total - done), seed)
        done += min(5000, total - done)
        seed += 1
        print(f"{done} cases, model == oracle (re-walks in this chunk: {rw})", flush=True)
