#!/usr/bin/env python3
"""bench.py -- audio-hours/sec analysed (BASELINE.json metric) on N B200s of one node.

Workload (config.workload): BASELINE.json configs[1] -- 1,000 synthetic 3-minute 44.1 kHz stereo CBR-128 MP3 tracks per
GPU, track-mode ReplayGain analysis.  A "step" is one pass of the whole hot path (frame index -> Huffman -> hybrid /
polyphase -> IIR -> histogram / percentile) over that batch.  Tracks are independent, so N GPUs shard by track with no
data-path collective (weak scaling: per-GPU work is fixed).

  value  whole-job audio-hours/sec with the file images already resident in HBM (rg_analyze_corpus)
  e2e    the same metric through rg_analyze_tracks with HOST (pinned) buffers: H2D of every file image and D2H of the
         results are inside the timed region
  roofline / cpu_baseline: see DESIGN.md "Measurement".

`--impl reference` times the reference's CPU path instead.  The reference is Rust + the un-vendored symphonia crate and
cannot be built in this image (no Rust toolchain, no network), so the stand-in is the plain-C oracle (oracle/) run on all
host threads over a bounded sample of the same workload -- the one other place this file may execute oracle/.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FRAMES_3MIN = 6891          # 3 min of 44.1 kHz MPEG-1 Layer III frames
GROUP_LEN = 32              # frames per self-contained group of the synthetic pool
METRIC = "audio-hours/sec analysed"


def build_corpus(n_tracks, n_frames, seed):
    """n_tracks distinct tracks: seeded random sequences of self-contained frame groups from a seeded pool."""
    import synth_corpus
    pool = synth_corpus.GroupPool(seed=seed, n_groups=256, group_len=GROUP_LEN, intensity_prob=0.02)
    return [pool.track(seed * 100003 + i, n_frames) for i in range(n_tracks)]


def build_mixed_corpus(n_tracks, seed, seconds=180.0):
    """BASELINE.json configs[3]: 3-minute tracks of divergent frame sizes -- VBR 32-320 kbps and CBR 32 / 128 / 320, 44.1 and
    48 kHz (two IIR coefficient sets in one batch), one track in five mono.  Returns (images, audio seconds)."""
    import synth_corpus
    kinds = [(0, 0, 2), (1, 0, 2), (0, 0, 1), (1, 14, 2), (0, 1, 2), (0, 9, 2), (1, 0, 1), (1, 9, 2), (0, 14, 2), (0, 0, 2)]   # (sr_idx, br_idx (0 = VBR), channels)
    pools = {}
    imgs, secs = [], 0.0
    for i in range(n_tracks):
        k = kinds[i % len(kinds)]
        if k not in pools:
            pools[k] = synth_corpus.GroupPool(seed=seed * 31 + len(pools), n_groups=64, group_len=GROUP_LEN, sr_idx=k[0], br_idx=k[1],
                                              channels=k[2], intensity_prob=0.02)
        rate = synth_corpus.RATES[k[0]]
        n_frames = max(1, int(seconds * rate / 1152.0))
        imgs.append(pools[k].track(seed * 7 + i, n_frames))
        secs += (n_frames + GROUP_LEN - 1) // GROUP_LEN * GROUP_LEN * 1152.0 / rate
    return imgs, secs


def pin_images(imgs, _lib):
    """Pinned host copies of the file images + the rg_input array over them (the e2e paths DMA straight out of these)."""
    import torch
    n = len(imgs)
    offs = np.zeros(n + 1, np.int64)
    for i, im in enumerate(imgs):
        offs[i + 1] = offs[i] + ((len(im) + 63) & ~63)
    pinned = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=True)
    pnp = pinned.numpy()
    arr = (_lib.RgInput * n)()
    for i, im in enumerate(imgs):
        pnp[offs[i]: offs[i] + len(im)] = im
        arr[i].data = pnp.ctypes.data + int(offs[i])
        arr[i].len = len(im)
        arr[i].track_index = -1
    return pinned, arr


def audio_seconds(n_tracks, n_frames):
    return n_tracks * n_frames * 1152 / 44100.0


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_reference_run(imgs, threads):
    """Oracle (CPU restatement of the reference path) over `imgs` with `threads` host threads; returns seconds."""
    import oracle_lib
    orc = oracle_lib.load()
    idx = list(range(len(imgs)))
    lock = threading.Lock()
    out = [None] * len(imgs)

    def work():
        while True:
            with lock:
                if not idx:
                    return
                i = idx.pop()
            rc, res, hist, info = orc.analyze_mp3(imgs[i])   # ctypes releases the GIL during the call
            out[i] = (rc, res.loudness_db, int(info.n_samples))

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work) for _ in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks", type=int, default=1000, help="tracks per GPU (configs[1]: 1000)")
    ap.add_argument("--frames", type=int, default=FRAMES_3MIN, help="frames per track (6891 = 3 min)")
    ap.add_argument("--exact", action="store_true", help="bit-exact serial IIR for every track (parity mode)")
    args = ap.parse_args()

    # stdout carries exactly ONE line, the JSON record: whatever libraries print there (NCCL's version banner ...) goes to
    # stderr instead -- the real stdout is kept aside and fd 1 points at stderr until the record is written
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = os.cpu_count() or 1
    # (the same dict in both arms: the reference arm times a bounded sample of exactly this workload)
    config = {"workload": f"configs[1]: {args.tracks} synthetic 3-min 44.1 kHz stereo CBR-128 MP3 tracks per GPU, track-mode",
              "tracks_per_gpu": args.tracks, "frames_per_track": args.frames,
              "frames_decoded_per_track": (args.frames + GROUP_LEN - 1) // GROUP_LEN * GROUP_LEN,   # the pool deals whole groups
              "sample_rate": 44100, "channels": 2,
              "bitrate_kbps": 128, "sharding": f"by-track x{args.gpus}", "iir_mode": "exact" if args.exact else "time-blocked+edge-rerun",
              "l2": "inputs (2.9 GB compressed per GPU, 60 GB of intermediates) are far larger than the 126 MB L2"}

    # ------------------------------------------------------------------ reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        n_sample = max(2 * ncores, 8)
        imgs = build_corpus(n_sample, args.frames, seed=1)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_run(imgs[: max(ncores // 2, 1)], ncores)
        times = []
        for _ in range(args.steps):
            dt, out = cpu_reference_run(imgs, ncores)
            assert all(o[0] == 0 for o in out)
            times.append(dt)
        ms = 1000.0 * sum(times) / len(times)
        val = sum(o[2] for o in out) / 44100.0 / 3600.0 / (ms / 1000.0)     # audio actually decoded (whole 32-frame groups)
        sample = f"{n_sample} of the workload's tracks per step on {ncores} host threads (oracle/: C restatement of the reference path)"
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "audio-h/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "audio-h/s", "cores": ncores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "audio-h/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # N ranks pull their file images over PCIe at the same time: keep each rank's threads, and with them its pinned staging
        # buffer (first touch), on the CPUs next to its GPU.  (At N = 1 the CPU baseline needs every core, so nothing is bound.)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:  # noqa: BLE001  (no NVML / restricted cpuset: run unbound)
            pass
    import mp3rgain_b200 as m
    from mp3rgain_b200 import _lib
    lib = _lib.load()   # raises if the CUDA extension has not been built: there is no fallback

    peaks_early = {}
    try:
        peaks_early = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    imgs = build_corpus(args.tracks, args.frames, seed=1 + rank)
    n = len(imgs)
    total_bytes = int(sum(len(i) for i in imgs))
    pinned, arr = pin_images(imgs, _lib)
    imgs = imgs[: max(2 * ncores, 8)]   # only the CPU-baseline sample is needed from here on (host RAM: N ranks share one box)
    eng = m.Engine(device=local_rank, exact=args.exact)
    out = (_lib.RgTrackResult * n)()
    asec = audio_seconds(n, args.frames)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: corpus resident in HBM before the timed region
    cor = C.c_void_p()
    rc = lib.rg_corpus_upload(eng._h, arr, n, C.byref(cor))
    assert rc == 0, lib.rg_last_error()
    for _ in range(args.warmup):
        assert lib.rg_analyze_corpus(eng._h, cor, out, None) == 0
    if args.warmup > 0:
        asec = sum(out[i].n_frames for i in range(n)) * 1152 / 44100.0   # audio actually decoded: the pool deals whole 32-frame groups
        assert int(out[0].n_frames) == config["frames_decoded_per_track"]
    st0 = eng.stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, hyb_ms, huf_ms, iir_ms, idx_ms, nbat = [], [], [], [], [], []
    for _ in range(args.steps):
        assert lib.rg_analyze_corpus(eng._h, cor, out, None) == 0      # blocks until the results are on the host
        st = eng.stats()
        dev_ms.append(st.last_ms_total); hyb_ms.append(st.last_ms_hybrid); huf_ms.append(st.last_ms_huffman)
        iir_ms.append(st.last_ms_iir); idx_ms.append(st.last_ms_index); nbat.append(st.last_batches)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st1 = eng.stats()
    assert all(out[i].status == 0 for i in range(n))
    # device time (CUDA events on the engine's stream, whole call), max over ranks
    step_ms = maxr(float(np.mean(dev_ms)))
    wall_ms = maxr(1000.0 * wall / args.steps)
    launches = int(st1.kernel_launches - st0.kernel_launches)
    reruns = int(st1.exact_reruns - st0.exact_reruns)
    iir_samples = int(st1.last_iir_samples)
    # one more pass OUTSIDE the timed region with the pipeline off (one batch at a time, every kernel alone on the device):
    # the per-kernel CUDA-event times of the timed passes overlap across streams, these do not
    eng.set_option("pipeline", 0)
    assert lib.rg_analyze_corpus(eng._h, cor, out, None) == 0
    sa = eng.stats()
    alone = {"k_frame_index+k_md_gather": sa.last_ms_index, "k_huffman": sa.last_ms_huffman, "k_hybrid_synth": sa.last_ms_hybrid,
             "k_iir": sa.last_ms_iir, "k_hist": sa.last_ms_hist, "launches_per_step": int(sa.last_batches),
             "iir_samples": int(sa.last_iir_samples)}
    eng.set_option("pipeline", 1)

    # ---- e2e: host buffers in, results out, copies inside the timed region
    for _ in range(max(1, args.warmup // 3)):
        assert lib.rg_analyze_tracks(eng._h, arr, n, out, None) == 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert lib.rg_analyze_tracks(eng._h, arr, n, out, None) == 0
    barrier()
    e2e_ms = maxr(1000.0 * (time.perf_counter() - t0) / args.steps)

    value = world * asec / 3600.0 / (wall_ms / 1000.0)
    e2e = world * asec / 3600.0 / (e2e_ms / 1000.0)

    # ---- the other BASELINE.json configs beside the headline (not part of `value`), each a short run of its own:
    #   album  configs[2]: album mode over the same resident corpus, sharded over all ranks; the timed region holds the whole
    #          call -- analysis, the engine's exchange (one NCCL group: all-reduce sum u64[12000] + max u64[2]) and the percentile
    #   mixed  configs[3]: VBR 32-320 + CBR, 44.1 / 48 kHz, 20 % mono (divergent frame sizes, two coefficient sets per batch)
    #   stream configs[4]: 30-second clips out of pinned host memory, copies overlapped with the kernels; H2D GB/s beside it
    extra = {}
    xsteps = max(1, min(args.steps, 3))
    try:
        from mp3rgain_b200 import sharded
        sharded.connect(eng, rank, world)
        alb = _lib.RgAlbumResult()
        bad = C.c_int64(-1)
        ahist = np.zeros(_lib.HIST_SIZE, np.uint64)
        xch = []
        assert lib.rg_analyze_album_sharded_corpus(eng._h, cor, rank * n, out, C.byref(alb), ahist.ctypes.data, C.byref(bad)) == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(xsteps):
            assert lib.rg_analyze_album_sharded_corpus(eng._h, cor, rank * n, out, C.byref(alb), ahist.ctypes.data, C.byref(bad)) == 0
            xch.append(eng.stats().last_ms_exchange)
        barrier()
        a_ms = maxr(1000.0 * (time.perf_counter() - t0) / xsteps)
        # the merged histogram holds every window of every rank's tracks exactly once
        n_win = float(sum(out[i].n_windows - out[i].n_dropped_windows for i in range(n)))
        if world > 1:
            t = torch.tensor([n_win], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            n_win = float(t.item())
        extra["album"] = {"workload": f"configs[2]: album mode, {n} tracks per GPU x {world} GPU(s), histogram merge over NCCL inside the timed call",
                          "value": world * asec / 3600.0 / (a_ms / 1000.0), "unit": "audio-h/s", "ms_per_step": a_ms,
                          "exchange_ms": float(np.mean(xch)), "exchange": "one NCCL group on the engine's stream: all-reduce(sum) u64[12000] + all-reduce(max) u64[2]" if world > 1 else "none (one rank)",
                          "album_loudness_db": alb.album_loudness_db, "album_peak": alb.album_peak,
                          "merged_windows": int(ahist.sum()), "windows_of_all_tracks": int(n_win), "histogram_is_the_sum": bool(int(ahist.sum()) == int(n_win))}
    except Exception as ex:  # noqa: BLE001  (a secondary measurement must never take the headline line down)
        extra["album"] = {"error": repr(ex)}
    lib.rg_corpus_free(eng._h, cor)
    try:
        mimgs, msec = build_mixed_corpus(args.tracks, seed=11 + rank, seconds=180.0 * args.frames / FRAMES_3MIN)   # (3 min unless a toy run)
        mbytes = int(sum(len(i) for i in mimgs))
        mpin, marr = pin_images(mimgs, _lib)
        nm = len(mimgs)
        mout = (_lib.RgTrackResult * nm)()
        mcor = C.c_void_p()
        assert lib.rg_corpus_upload(eng._h, marr, nm, C.byref(mcor)) == 0
        assert lib.rg_analyze_corpus(eng._h, mcor, mout, None) == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(xsteps):
            assert lib.rg_analyze_corpus(eng._h, mcor, mout, None) == 0
        barrier()
        m_ms = maxr(1000.0 * (time.perf_counter() - t0) / xsteps)
        lib.rg_corpus_free(eng._h, mcor)
        assert lib.rg_analyze_tracks(eng._h, marr, nm, mout, None) == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(xsteps):
            assert lib.rg_analyze_tracks(eng._h, marr, nm, mout, None) == 0
        barrier()
        me_ms = maxr(1000.0 * (time.perf_counter() - t0) / xsteps)
        assert all(mout[i].status == 0 for i in range(nm))
        rates = sorted({int(mout[i].sample_rate) for i in range(nm)})
        extra["mixed"] = {"workload": f"configs[3]: {nm} 3-min tracks per GPU, VBR 32-320 / CBR 32, 128, 320 kbps, 44.1 + 48 kHz, 20 % mono; sharded by bytes",
                          "value": world * msec / 3600.0 / (m_ms / 1000.0), "unit": "audio-h/s", "ms_per_step": m_ms,
                          "e2e": {"value": world * msec / 3600.0 / (me_ms / 1000.0), "ms_per_step": me_ms, "h2d_bytes_per_step": mbytes * world},
                          "sample_rates": rates, "mono_tracks": int(sum(1 for i in range(nm) if mout[i].channels == 1)),
                          "compressed_bytes_per_gpu": mbytes, "exact_reruns": int(sum(mout[i].exact_rerun for i in range(nm)))}
        del mpin, marr
    except Exception as ex:  # noqa: BLE001
        extra["mixed"] = {"error": repr(ex)}
    try:
        import synth_corpus
        clip_frames = min(int(30.0 * 44100 / 1152.0), args.frames)   # (30 s unless a toy run)
        n_clips = 6 * args.tracks                      # the same 50 audio-hours per GPU, in 30-second pieces
        pool = synth_corpus.GroupPool(seed=101 + rank, n_groups=128, group_len=GROUP_LEN, intensity_prob=0.02)
        cimgs = [pool.track(977 * rank + i, clip_frames) for i in range(n_clips)]
        cbytes = int(sum(len(i) for i in cimgs))
        csec = n_clips * ((clip_frames + GROUP_LEN - 1) // GROUP_LEN * GROUP_LEN) * 1152.0 / 44100.0
        cpin, carr = pin_images(cimgs, _lib)
        del cimgs
        cout = (_lib.RgTrackResult * n_clips)()
        assert lib.rg_analyze_tracks(eng._h, carr, n_clips, cout, None) == 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(xsteps):
            assert lib.rg_analyze_tracks(eng._h, carr, n_clips, cout, None) == 0
        barrier()
        c_ms = maxr(1000.0 * (time.perf_counter() - t0) / xsteps)
        assert all(cout[i].status == 0 for i in range(n_clips))
        extra["stream"] = {"workload": f"configs[4]: {n_clips} 30-second clips per GPU out of pinned host memory, copy / index / decode / analysis overlapped",
                           "value": world * csec / 3600.0 / (c_ms / 1000.0), "unit": "audio-h/s", "ms_per_step": c_ms,
                           "clips_per_s": world * n_clips / (c_ms / 1000.0), "h2d_gbs_per_gpu": cbytes / (c_ms / 1000.0) / 1e9,
                           "h2d_bytes_per_step": cbytes * world, "hbm_peak_gbs": float(peaks_early.get("hbm_gbs", 6650.0)),
                           "note": "the compressed stream is what crosses the bus: H2D GB/s is bounded by PCIe (~55 GB/s per GPU), 1 % of the HBM peak"}
        del cpin, carr
    except Exception as ex:  # noqa: BLE001
        extra["stream"] = {"error": repr(ex)}

    # ---- row f1 beside it (not part of `value`): the global_gain walk of mp3rgain::analyze() / apply_gain() over the same
    # corpus.  Its scan kernel streams every file byte once through shared memory -- the one kernel of this repo whose
    # algorithmic bytes (the file bytes) ARE its traffic, so its HBM fraction means something.
    gain_fields, ga_ms, gp_ms = None, -1.0, -1.0
    try:   # (no collective inside: a failure on one rank must not leave the others waiting)
        gst = (_lib.RgGainStats * n)()
        cor2 = C.c_void_p()
        assert lib.rg_corpus_upload(eng._h, arr, n, C.byref(cor2)) == 0
        scan, merge = [], []
        for it in range(args.warmup + args.steps):
            assert lib.rg_gain_analyze_corpus(eng._h, cor2, gst) == 0
            if it >= args.warmup:
                sg = eng.stats()
                scan.append(sg.last_ms_gain_scan); merge.append(sg.last_ms_gain_merge)
        lib.rg_corpus_free(eng._h, cor2)
        assert all(gst[i].status == 0 for i in range(n))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            assert lib.rg_gain_analyze(eng._h, arr, n, gst) == 0
        ga_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
        # patched in place (out_images may alias the inputs): no second pinned copy of the corpus per rank.  Every timed
        # number of the ReplayGain path above is already taken; the oracle checks below use the untouched `imgs`
        optr = (C.c_void_p * n)(*[arr[i].data for i in range(n)])
        mod = np.zeros(n, np.uint32)
        assert lib.rg_gain_apply(eng._h, arr, n, 3, 0, -1, optr, mod.ctypes.data) == 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            assert lib.rg_gain_apply(eng._h, arr, n, 3, 0, -1, optr, mod.ctypes.data) == 0
        gp_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
        patch_ms = eng.stats().last_ms_gain_patch
        scan_ms = float(np.mean(scan))
        gf_traffic = None
        try:
            per_track = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_gf_scan_dram_bytes_per_track")
            gf_traffic = per_track * n if per_track is not None else None
        except Exception:  # noqa: BLE001
            pass
        hbm = float(peaks_early.get("hbm_gbs", 6650.0))
        gain_fields = {"workload": "mp3rgain::analyze() / apply_gain(+3) over the same corpus (SURVEY 8 row f1)",
                       "frames_per_track": int(gst[0].frame_count),
                       "kernel_ms": {"k_gf_scan": scan_ms, "k_gf_merge": float(np.mean(merge)), "k_gf_patch": patch_ms},
                       "roofline": {"bound": "hbm", "kernel": "k_gf_scan",
                                    "achieved": (gf_traffic / (scan_ms / 1000.0) / 1e9) if gf_traffic else None,
                                    "peak": hbm, "unit": "GB/s", "frac": (gf_traffic / (scan_ms / 1000.0) / 1e9 / hbm) if gf_traffic else None,
                                    "algorithmic_bytes_per_launch": total_bytes, "traffic": gf_traffic,
                                    "file_bytes_gbs": total_bytes / (scan_ms / 1000.0) / 1e9,
                                    "note": "achieved / frac = DRAM bytes of the launch (ncu, profiles/traffic.json) over its CUDA-event time: the "
                                            "walk reads only the sectors that hold headers and side info, about half of the file; file_bytes_gbs "
                                            "= the rate a caller sees (file bytes / time)"},
                       "frames_modified": int(mod.sum())}
    except Exception as ex:  # noqa: BLE001  (a secondary measurement must never take the headline line down)
        gain_fields = {"error": repr(ex)}
    ga_ms, gp_ms = maxr(ga_ms), maxr(gp_ms)
    if gain_fields and "error" not in gain_fields:
        gain_fields["analyze_e2e"] = {"ms_per_step": ga_ms, "files_per_s": world * n / (ga_ms / 1000.0), "h2d_gbs_per_gpu": total_bytes / (ga_ms / 1000.0) / 1e9}
        gain_fields["apply_e2e"] = {"ms_per_step": gp_ms, "files_per_s": world * n / (gp_ms / 1000.0),
                                    "h2d_plus_d2h_gbs_per_gpu": 2 * total_bytes / (gp_ms / 1000.0) / 1e9}

    # ---- roofline of the dominant kernel (k_hybrid_synth, one launch per batch of the step).  Per launch:
    #   achieved = algorithmic bytes (the compressed file bytes, 16,000 B per audio-second at CBR-128, of the
    #              tracks that launch decodes) / the launch's duration (CUDA events on the decode stream, inside the step)
    #   traffic  = dram read+write of one launch from the ncu --set full capture (profiles/traffic.json, scaled by tracks)
    # The kernels of consecutive batches overlap on separate streams, so the per-kernel times sum to more than the step.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    launches_per_step = max(1.0, float(np.mean(nbat)))
    k_ms = float(np.mean(hyb_ms)) / launches_per_step
    alg_bytes = total_bytes / launches_per_step
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    traffic = staged = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        per_track = tj.get("k_hybrid_synth_dram_bytes_per_track")
        if per_track is not None:
            traffic = per_track * n / launches_per_step
        staged = tj.get("all_kernels_dram_bytes_per_track")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": "k_hybrid_synth", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k_ms, "launches_per_step": launches_per_step,
                "kernel_ms_per_step": {"k_frame_index+k_md_gather": float(np.mean(idx_ms)), "k_huffman": float(np.mean(huf_ms)),
                                       "k_hybrid_synth": float(np.mean(hyb_ms)), "k_iir": float(np.mean(iir_ms))},
                "kernel_ms_alone": alone,
                "achieved_alone": total_bytes / (alone["k_hybrid_synth"] / 1000.0) / 1e9,
                "staged_dram_gbs_whole_step": (staged * n / (wall_ms / 1000.0) / 1e9) if staged else None,
                "note": "no kernel of this path is HBM-bound: k_hybrid_synth is bound by instruction issue and dependency latency (61 % issue-active "
                        "after the packed-f32x2 rewrite, profiles/r2_ncu_full_100tracks_summary.csv), k_iir FP64-pipe bound (88 %); the compressed stream is ~1 % of HBM peak by construction "
                        "(DESIGN.md 6). staged_dram_gbs_whole_step = ncu DRAM bytes of all kernels (spectra, PCM round trips) / step time"}

    # ---- the roofline that actually binds the second-largest kernel: k_iir is FP64-pipe bound.  28 FP64 instructions per
    # channel-sample (21 Yule + 5 Butterworth + square-accumulate + the b10 product) over every sample the
    # time-blocked chains touch (blocks + warm-up windows, counted by the engine), against the DFMA rate tools/fp64_pipe.cu
    # measured on this GPU type (profiles/fp64_peak.json), else 64 DFMA/clk/SM at the SM clock sampled during the run.
    fp64_peak = None
    try:
        fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json"))).get("dfma_per_s")
    except Exception:  # noqa: BLE001
        pass
    if not fp64_peak:
        fp64_peak = 64.0 * 148 * (clocks.get("sm_mhz") or 1965.0) * 1e6
    if not args.exact:
        ops_timed, t_timed = iir_samples * 28.0, float(np.mean(iir_ms)) / 1000.0
        ops_alone, t_alone = alone["iir_samples"] * 28.0, alone["k_iir"] / 1000.0
        roofline["fp64"] = {"kernel": "k_iir_fast", "bound": "fp64 pipe", "unit": "T FP64 instr/s (DFMA = 1)", "peak": fp64_peak / 1e12,
                            "achieved": ops_timed / t_timed / 1e12, "frac": ops_timed / t_timed / fp64_peak,
                            "achieved_alone": ops_alone / t_alone / 1e12, "frac_alone": ops_alone / t_alone / fp64_peak,
                            "ops_per_channel_sample": 28, "channel_samples_per_step": iir_samples,
                            "note": "achieved: CUDA events inside the timed (pipelined) step, where K1/K2 of the next batch share the "
                                    "SMs; achieved_alone: the same kernel with the pipeline off"}

    # ---- and the one that binds the largest: k_hybrid_synth is bound by instruction issue (one warp instruction per SM
    # sub-partition and clock).  Executed warp-instructions per track from the ncu capture (profiles/traffic.json, counted on
    # the same corpus generator), over the kernel's time with the pipeline off, against 4 x SMs x the SM clock of this run.
    try:
        wi = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["warp_instructions_per_track"]["k_hybrid_synth"]
        wi *= config["frames_decoded_per_track"] / 6912.0   # (the capture decoded 6912 frames per track)
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        try:
            n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        except Exception:  # noqa: BLE001  (the interpreted build of the CPU suite has no device)
            n_sms = 148
        issue_peak = 4.0 * n_sms * sm_hz
        roofline["issue"] = {"kernel": "k_hybrid_synth", "bound": "instruction issue", "unit": "T warp-instr/s", "peak": issue_peak / 1e12,
                             "achieved_alone": wi * n / (alone["k_hybrid_synth"] / 1000.0) / 1e12,
                             "frac_alone": wi * n / (alone["k_hybrid_synth"] / 1000.0) / issue_peak,
                             "warp_instructions_per_track": wi,
                             "note": "instruction count from ncu (3-min CBR-128 stereo tracks: 1212 per granule-channel), time from CUDA events of "
                                     "this run with the pipeline off"}
    except Exception:  # noqa: BLE001
        pass

    line = {"metric": METRIC, "value": value, "unit": "audio-h/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_ms, "device_ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": launches,
            "exact_reruns": reruns,
            "e2e": {"value": e2e, "unit": "audio-h/s", "h2d_bytes_per_step": total_bytes * world,
                    "d2h_bytes_per_step": n * C.sizeof(_lib.RgTrackResult) * world, "ms_per_step": e2e_ms},
            "roofline": roofline, "configs": extra, "gain_fields": gain_fields}

    if rank == 0 and world == 1:
        # CPU baseline beside it: the oracle (C restatement of the reference path) on a bounded sample, all host threads
        n_sample = min(max(2 * ncores, 8), n)
        dt, o = cpu_reference_run(imgs[:n_sample], ncores)
        cpu_val = sum(x[2] for x in o) / 44100.0 / 3600.0 / dt
        # parity spot check of the timed results against the oracle on the sample
        mism = sum(1 for i in range(n_sample) if o[i][1] != out[i].loudness_db)
        # BASELINE.json's "delta dB vs ref": track loudness of the timed GPU results against the CPU restatement of the reference
        line["delta_db_vs_ref"] = {"max_abs": max(abs(o[i][1] - out[i].loudness_db) for i in range(n_sample)), "tracks": n_sample,
                                   "tolerance": 0.01, "ref": "oracle/ (C restatement of src/replaygain.rs + ISO 11172-3 decode)"}
        if gain_fields and "error" not in gain_fields:
            import oracle_lib
            orc = oracle_lib.load()
            t0 = time.perf_counter()
            for i in range(n_sample):
                rc_o, st_o = orc.gain_analyze(imgs[i])
                assert rc_o == 0 and st_o.frame_count == gst[i].frame_count and st_o.gain_sum == gst[i].gain_sum
            dtg = time.perf_counter() - t0
            gain_fields["cpu_baseline"] = {"files_per_s": n_sample / dtg, "cores": 1, "kind": "port",
                                           "sample": f"oracle gfo_analyze over the first {n_sample} tracks, one thread, images in RAM"}
        dt1, o1 = cpu_reference_run(imgs[:2], 1)   # the same restatement on ONE core (the reference itself is single-threaded)
        one_core = sum(x[2] for x in o1) / 44100.0 / 3600.0 / dt1
        line["cpu_baseline"] = {"value": cpu_val, "unit": "audio-h/s", "cores": ncores, "kind": "port",
                                "sample": f"first {n_sample} tracks of the workload, one pass, {ncores} host threads, {dt:.1f} s",
                                "loudness_mismatches_vs_gpu": mism, "one_core": {"value": one_core, "unit": "audio-h/s", "sample": f"first 2 tracks, {dt1:.1f} s"}}
    if rank == 0:
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
