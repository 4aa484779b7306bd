#!/usr/bin/env python3
"""Development aid: time rg_analyze_corpus over one resident synthetic corpus under several engine option sets.
usage: opt_sweep.py N_TRACKS "k=v,k=v" "k=v,..." ...   (every set is applied on top of the defaults listed in BASE)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mp3rgain_b200 as m  # noqa: E402
from mp3rgain_b200 import _lib  # noqa: E402
import synth_corpus  # noqa: E402

BASE = {"pipeline": 1, "k4_slim": 0, "iir_block_windows": 0, "iir_warm_windows": 1, "run_granules": 254}
n_tracks = int(sys.argv[1])
sets = sys.argv[2:] or [""]
pool = synth_corpus.GroupPool(seed=1, n_groups=256, group_len=32, intensity_prob=0.02)
imgs = [pool.track(1000 + i, 6891) for i in range(n_tracks)]
lib = _lib.load()
eng = m.Engine(device=0, exact=False, scratch_bytes=int(float(os.environ.get("RG_SCRATCH_GB", "0")) * (1 << 30)))
arr, keep = m.Engine._inputs(imgs, None)
cor = C.c_void_p()
assert lib.rg_corpus_upload(eng._h, arr, n_tracks, C.byref(cor)) == 0
out = (_lib.RgTrackResult * n_tracks)()
ref = None
for spec in sets:
    opts = dict(BASE)
    for kv in spec.split(","):
        if "=" in kv:
            opts[kv.split("=")[0]] = float(kv.split("=")[1])
    for k in BASE:
        eng.set_option(k, opts[k])
    best = None
    for it in range(4):
        t = time.time()
        rc = lib.rg_analyze_corpus(eng._h, cor, out, None)
        dt = (time.time() - t) * 1e3
        st = eng.stats()
        if best is None or dt < best[0]:
            best = (dt, st.last_ms_total, st.last_ms_index, st.last_ms_huffman, st.last_ms_hybrid, st.last_ms_iir, st.last_ms_hist, st.last_block_windows)
    res = [(out[i].loudness_db, out[i].peak) for i in range(n_tracks)]
    if ref is None:
        ref = res
    same = res == ref
    print("%-70s wall %.1f dev %.1f | idx %.1f huff %.1f hyb %.1f iir %.1f hist %.1f bw %d | same=%s" % ((spec or "(defaults)",) + best + (same,)), flush=True)
