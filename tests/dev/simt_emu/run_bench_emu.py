#!/usr/bin/env python3
"""Runs bench.py's GPU arm in the GPU-less container: the engine is the interpreted build of the kernel sources (this
directory) and the handful of torch.cuda calls bench.py makes (set_device, synchronize, pinned allocation) are no-ops.
A check of bench.py's own code paths and of the JSON record's keys at toy sizes -- every number it prints is meaningless.
usage: run_bench_emu.py [bench.py arguments]   (development aid, not a product path)"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from mp3rgain_b200 import _lib  # noqa: E402
_lib.LIB_PATH = build_emu.build()

import torch  # noqa: E402
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
_empty = torch.empty
torch.empty = lambda *a, **k: _empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"})

sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
