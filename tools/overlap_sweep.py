#!/usr/bin/env python3
"""Sweep of the K4 (k_iir_fast) launch shape in the pipelined step: classic (256-thread CTAs, two per SM) against "slim" waves
of one or two 128-thread CTAs per SM, which leave register file for the decode kernels of the next batch (K1 of batch k+1 and
K4 of batch k become eligible at the same moment, when K2 of batch k ends).  usage: overlap_sweep.py [n_tracks=1000]"""
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

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pool = synth_corpus.GroupPool(seed=1, n_groups=256, group_len=32, intensity_prob=0.02)
imgs = [pool.track(1000 + i, 6891) for i in range(n_tracks)]
lib = _lib.load()
eng = m.Engine(device=0)
arr, keep = m.Engine._inputs(imgs, None)
cor = C.c_void_p()
assert lib.rg_corpus_upload(eng._h, arr, n_tracks, C.byref(cor)) == 0
out = (_lib.RgTrackResult * n_tracks)()
ref = None
for slim, pipe in [(0, 1), (1, 1), (2, 1), (0, 0), (1, 0), (2, 0), (1, 1), (0, 1)]:
    eng.set_option("k4_slim", slim)
    eng.set_option("pipeline", pipe)
    ts = []
    for it in range(4):
        t = time.time()
        assert lib.rg_analyze_corpus(eng._h, cor, out, None) == 0
        ts.append(1e3 * (time.time() - t))
    st = eng.stats()
    res = [(out[i].loudness_db, out[i].peak) for i in range(n_tracks)]
    ref = ref or res
    print("k4_slim %d pipeline %d: wall %s ms | huff %.1f hybrid %.1f iir %.1f index %.1f | block windows %d reruns %d | same results %s" % (
        slim, pipe, " ".join("%.1f" % x for x in ts), st.last_ms_huffman, st.last_ms_hybrid, st.last_ms_iir, st.last_ms_index,
        st.last_block_windows, st.exact_reruns, res == ref), flush=True)
