#!/usr/bin/env python3
"""Stage timings of BASELINE.json configs[3] (bench.py's mixed corpus: VBR 32-320 / CBR 32, 128, 320 kbps, 44.1 + 48 kHz, one
track in five mono), resident with the pipeline on and off, and end to end.  Development aid.
usage: mixed_probe.py [n_tracks=1000]   (RG_OPTS=k=v,... sets engine options, RG_SCRATCH_GIB the scratch budget)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import mp3rgain_b200 as m  # noqa: E402
from mp3rgain_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t = time.time()
imgs, secs = bench.build_mixed_corpus(n, seed=11)
print("mixed corpus: %d tracks, %.1f MB, %.1f audio-hours, built in %.1fs" % (n, sum(len(i) for i in imgs) / 1e6, secs / 3600, time.time() - t), flush=True)
lib = _lib.load()
kw = {}
if os.environ.get("RG_SCRATCH_GIB"):
    kw["scratch_bytes"] = int(float(os.environ["RG_SCRATCH_GIB"]) * (1 << 30))
eng = m.Engine(device=0, **kw)
for kv in os.environ.get("RG_OPTS", "").split(","):
    if "=" in kv:
        eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
pin, arr = bench.pin_images(imgs, _lib)
out = (_lib.RgTrackResult * n)()
cor = C.c_void_p()
assert lib.rg_corpus_upload(eng._h, arr, n, C.byref(cor)) == 0


def line(tag, dt):
    st = eng.stats()
    print("%s wall %.1f ms | index %.1f huff %.1f hybrid %.1f iir %.1f hist %.1f | audio-h/s %.1f | batches %d bw %d reruns %d" % (
        tag, dt * 1e3, st.last_ms_index, st.last_ms_huffman, st.last_ms_hybrid, st.last_ms_iir, st.last_ms_hist, secs / 3600.0 / dt,
        st.last_batches, st.last_block_windows, st.exact_reruns), flush=True)


for pipe in (1, 0, 1):
    eng.set_option("pipeline", pipe)
    for it in range(3):
        t = time.time()
        assert lib.rg_analyze_corpus(eng._h, cor, out, None) == 0
        line("resident pipeline=%d" % pipe, time.time() - t)
lib.rg_corpus_free(eng._h, cor)
for it in range(4):
    t = time.time()
    assert lib.rg_analyze_tracks(eng._h, arr, n, out, None) == 0
    line("end to end", time.time() - t)
