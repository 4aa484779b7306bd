#!/usr/bin/env python3
"""Quick stage-timing probe (development aid): N synthetic 3-min tracks, device-resident, per-stage CUDA-event times."""
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
exact = len(sys.argv) > 3 and sys.argv[3] == "exact"
t = time.time()
pool = synth_corpus.GroupPool(seed=1, n_groups=256, group_len=32, intensity_prob=0.02)
imgs = [pool.track(1000 + i, n_frames) for i in range(n_tracks)]
print("corpus built in %.1fs, %.1f MB" % (time.time() - t, sum(len(i) for i in imgs) / 1e6), flush=True)
_lib.LIB_PATH = os.environ.get("RG_LIB", _lib.LIB_PATH)   # a variant build of the engine (kernel experiments)
lib = _lib.load()
DEV = int(os.environ.get("LOCAL_RANK", "0"))   # under torchrun: one probe per GPU at the same time (shared-host contention)
eng = m.Engine(device=DEV, exact=exact)
for kv in os.environ.get("RG_OPTS", "").split(","):   # e.g. RG_OPTS=k2_smem_pad=30000,iir_block_windows=16
    if "=" in kv:
        eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
if os.environ.get("RG_PIPELINE") == "0":
    eng.set_option("pipeline", 0)   # one batch at a time, every kernel alone on the device
arr, keep = m.Engine._inputs(imgs, None)
cor = C.c_void_p()
t = time.time()
assert lib.rg_corpus_upload(eng._h, arr, n_tracks, C.byref(cor)) == 0
print("upload %.3fs" % (time.time() - t), flush=True)
out = (_lib.RgTrackResult * n_tracks)()
audio_s = n_tracks * n_frames * 1152 / 44100.0
for it in range(3):
    t = time.time()
    rc = lib.rg_analyze_corpus(eng._h, cor, out, None)
    dt = time.time() - t
    st = eng.stats()
    print("rc", rc, "wall %.3fs  dev total %.1f ms | index %.1f huff %.1f hybrid %.1f iir %.1f hist %.1f | audio-h/s %.1f | reruns %d launches %d" % (
        dt, st.last_ms_total, st.last_ms_index, st.last_ms_huffman, st.last_ms_hybrid, st.last_ms_iir, st.last_ms_hist,
        audio_s / 3600.0 / dt, st.exact_reruns, st.kernel_launches), "batches", st.last_batches, "bw", st.last_block_windows, flush=True)
print("sample result", out[0].loudness_db, out[0].peak, out[0].n_windows, out[0].n_edge_windows, out[0].exact_rerun)

if len(sys.argv) > 4 and sys.argv[4] == "e2e":
    # host buffers (pinned) -> rg_analyze_tracks, plus the raw H2D rate of the same bytes for reference
    import torch
    offs = np.zeros(n_tracks + 1, np.int64)
    for i, im in enumerate(imgs):
        offs[i + 1] = offs[i] + ((len(im) + 63) & ~63)
    pinned = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=True)
    pnp = pinned.numpy()
    arr2 = (_lib.RgInput * n_tracks)()
    for i, im in enumerate(imgs):
        pnp[offs[i]: offs[i] + len(im)] = np.frombuffer(im, np.uint8)
        arr2[i].data = pnp.ctypes.data + int(offs[i])
        arr2[i].len = len(im)
        arr2[i].track_index = -1
    torch.cuda.set_device(DEV)
    dev = torch.empty(int(offs[-1]), dtype=torch.uint8, device="cuda")
    for it in range(3):
        torch.cuda.synchronize(); t = time.time(); dev.copy_(pinned, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t
        print("raw H2D %.1f ms  %.1f GB/s" % (dt * 1e3, offs[-1] / dt / 1e9), flush=True)
    # RG_E2E_SWEEP="k=v,k=v;k=v": option sets applied one after the other, four calls each.  Under torchrun the ranks meet at a
    # barrier before every call, so that all GPUs of the box pull their bytes through the host at the same time.
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    for opts in os.environ.get("RG_E2E_SWEEP", "").split(";"):
        for kv in opts.split(","):
            if "=" in kv:
                eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
        for it in range(4):
            if world > 1:
                dist.barrier()
            t = time.time()
            rc = lib.rg_analyze_tracks(eng._h, arr2, n_tracks, out, None)
            dt = time.time() - t
            st = eng.stats()
            print("e2e", opts, "rc", rc, "wall %.4fs | index %.1f huff %.1f hybrid %.1f iir %.1f hist %.1f | audio-h/s %.1f | batches %d bw %d" % (
                dt, st.last_ms_index, st.last_ms_huffman, st.last_ms_hybrid, st.last_ms_iir, st.last_ms_hist, audio_s / 3600.0 / dt,
                st.last_batches, st.last_block_windows), flush=True)

if len(sys.argv) > 4 and sys.argv[4] == "peak":
    pk = (_lib.RgPeakResult * n_tracks)()
    for it in range(3):
        t = time.time()
        rc = lib.rg_find_peak(eng._h, arr, n_tracks, pk)
        dt = time.time() - t
        st = eng.stats()
        print("peak-only rc", rc, "wall %.3fs (pageable host buffers) | index %.1f huff %.1f hybrid %.1f | audio-h/s %.1f | peak[0] %.6f" % (
            dt, st.last_ms_index, st.last_ms_huffman, st.last_ms_hybrid, audio_s / 3600.0 / dt, pk[0].peak), flush=True)

if len(sys.argv) > 4 and sys.argv[4] == "e2e_pageable":
    # the same call from PAGEABLE host buffers: the copies go through the driver's staging at a fraction of the pinned rate, i.e.
    # the call is bound by the host link -- a one-GPU stand-in for eight ranks sharing the host's PCIe complex
    for it in range(4):
        t = time.time()
        rc = lib.rg_analyze_tracks(eng._h, arr, n_tracks, out, None)
        dt = time.time() - t
        st = eng.stats()
        print("e2e(pageable) rc", rc, "wall %.3fs | audio-h/s %.1f | batches %d | %.1f GB/s" % (
            dt, audio_s / 3600.0 / dt, st.last_batches, sum(len(i) for i in imgs) / dt / 1e9), flush=True)

if len(sys.argv) > 4 and sys.argv[4] == "resident_with_h2d":
    # does a host->device copy running beside the kernels slow them down?  The resident corpus again, with an unrelated pinned
    # buffer streaming to the device on another stream for the whole call
    import threading
    import torch
    side = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    cs = torch.cuda.Stream()
    stop = threading.Event()

    def pump():
        with torch.cuda.stream(cs):
            while not stop.is_set():
                dst.copy_(side, non_blocking=True)
                cs.synchronize()
    for with_copy in (False, True, False, True):
        th = None
        if with_copy:
            stop.clear()
            th = threading.Thread(target=pump)
            th.start()
            time.sleep(0.05)
        t = time.time()
        rc = lib.rg_analyze_corpus(eng._h, cor, out, None)
        dt = time.time() - t
        st = eng.stats()
        if th:
            stop.set()
            th.join()
        print("resident, H2D beside it: %s | wall %.3fs | index %.1f huff %.1f hybrid %.1f iir %.1f hist %.1f" % (
            with_copy, dt, st.last_ms_index, st.last_ms_huffman, st.last_ms_hybrid, st.last_ms_iir, st.last_ms_hist), flush=True)
