#!/usr/bin/env python3
"""Timing probe of the global_gain walk (k_gain_fields.cuh): N synthetic 3-min tracks resident in HBM, CUDA-event times of
k_gf_scan / k_gf_merge / k_gf_patch.  usage: gain_probe.py [n_tracks=200] [n_frames=6891]  (development aid; ncu target)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mp3rgain_b200 as m  # noqa: E402
from mp3rgain_b200 import _lib  # noqa: E402
import synth_corpus  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6891
pool = synth_corpus.GroupPool(seed=1, n_groups=256, group_len=32, intensity_prob=0.02)
imgs = [pool.track(1000 + i, n_frames) for i in range(n_tracks)]
total = sum(len(i) for i in imgs)
lib = _lib.load()
eng = m.Engine(device=0)
arr, keep = m.Engine._inputs(imgs, None)
cor = C.c_void_p()
assert lib.rg_corpus_upload(eng._h, arr, n_tracks, C.byref(cor)) == 0
gst = (_lib.RgGainStats * n_tracks)()
for it in range(4):
    t = time.time()
    assert lib.rg_gain_analyze_corpus(eng._h, cor, gst) == 0
    st = eng.stats()
    print("analyze: wall %.2f ms | scan %.3f ms (%.0f GB/s) merge %.3f ms | frames[0] %d rewalks %d" % (
        1e3 * (time.time() - t), st.last_ms_gain_scan, total / st.last_ms_gain_scan / 1e6, st.last_ms_gain_merge,
        gst[0].frame_count, sum(g.walk_rewalks for g in gst)), flush=True)
outs = [np.empty(len(i), np.uint8) for i in imgs]
ptrs = (C.c_void_p * n_tracks)(*[o.ctypes.data for o in outs])
mod = np.zeros(n_tracks, np.uint32)
for it in range(2):
    t = time.time()
    assert lib.rg_gain_apply(eng._h, arr, n_tracks, 3, 0, -1, ptrs, mod.ctypes.data) == 0
    st = eng.stats()
    print("apply (pageable host buffers): wall %.1f ms | scan %.3f merge %.3f patch %.3f ms | groups %d" % (
        1e3 * (time.time() - t), st.last_ms_gain_scan, st.last_ms_gain_merge, st.last_ms_gain_patch, st.last_gain_groups), flush=True)

# pinned in / out, and the raw bus next to it: H2D alone, D2H alone, both at once (two streams)
import torch  # noqa: E402
offs = np.zeros(n_tracks + 1, np.int64)
for i, im in enumerate(imgs):
    offs[i + 1] = offs[i] + ((len(im) + 63) & ~63)
pin_in = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=True)
pin_out = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=True)
arr2 = (_lib.RgInput * n_tracks)()
for i, im in enumerate(imgs):
    pin_in.numpy()[offs[i]: offs[i] + len(im)] = im
    arr2[i].data = pin_in.numpy().ctypes.data + int(offs[i]); arr2[i].len = len(im); arr2[i].track_index = -1
optr = (C.c_void_p * n_tracks)(*[pin_out.numpy().ctypes.data + int(offs[i]) for i in range(n_tracks)])
for it in range(3):
    t = time.time()
    assert lib.rg_gain_apply(eng._h, arr2, n_tracks, 3, 0, -1, optr, mod.ctypes.data) == 0
    print("apply (pinned): wall %.1f ms" % (1e3 * (time.time() - t)), flush=True)
dev_a = torch.empty(int(offs[-1]), dtype=torch.uint8, device="cuda")
dev_b = torch.empty(int(offs[-1]), dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, up, down in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H at once", True, True)):
    torch.cuda.synchronize()
    t = time.time()
    if up:
        with torch.cuda.stream(s1):
            dev_a.copy_(pin_in, non_blocking=True)
    if down:
        with torch.cuda.stream(s2):
            pin_out.copy_(dev_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.time() - t
    print("%s: %.1f ms (%.1f GB/s per direction)" % (name, 1e3 * dt, total / dt / 1e9), flush=True)
