"""Builds tests/dev/simt_emu/librg_engine_emu.so: the engine's CUDA sources compiled by g++ against the SIMT interpreter in
cuda_runtime.h (this directory).  TEST INFRASTRUCTURE ONLY -- see the header of cuda_runtime.h.

The sources under mp3rgain_b200/csrc are copied to a scratch directory with three mechanical rewrites, none of which touches
kernel logic:
  1. `kernel<<<grid, block, smem, stream>>>(args);`  ->  `EMU_LAUNCH(grid, block, smem, stream, kernel(args));`
  2. `extern __shared__ <attrs> T name[];`           ->  `T* name = EMU_DYN_SMEM(T);`
  3. the bodies of the five inline-PTX helpers of K2 (mbarrier init / expect_tx / try_wait, cp.async.bulk, bar.sync id, 64)
     are replaced by calls into the interpreter's rendezvous primitives of the same meaning.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, "mp3rgain_b200", "csrc")
# SIMT_EMU_ASAN=1: the same build under AddressSanitizer (global memory is heap memory and __shared__ arrays are statics here,
# so an out-of-bounds access of a kernel lands in a redzone).  The process that loads it needs
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
ASAN = os.environ.get("SIMT_EMU_ASAN") == "1"
OUT = os.path.join(HERE, "librg_engine_emu_asan.so" if ASAN else "librg_engine_emu.so")
SCRATCH = os.path.join(HERE, "_build_asan" if ASAN else "_build")

PTX_HELPERS = {
    "k2_mbar_init": "emu::mbar_init(bar, count);",
    "k2_mbar_expect": "emu::mbar_expect(bar, bytes);",
    "k2_bulk_load": "emu::bulk_load(dst, src, bytes, bar);",
    "k2_mbar_wait": "emu::mbar_wait(bar, parity);",
    "k2_pair_sync": "emu::barrier(1 + (pair < 3 ? pair : 3), 64);",
    "k2_prmt": "return emu::prmt(a, 0u, sel);",
    "k2_cp_async16": "memcpy(smem_dst, gsrc, 16);",      # (the interpreter copies at once: wait / commit have nothing to do)
    "k2_cp_async_commit": "",
    "k2_cp_async_wait1": "",
    "k2_cp_async_wait0": "",
    # packed f32x2: each half is the scalar IEEE operation (the build has -ffp-contract=off; fmaf is the fused form)
    "k2_ffma2": "return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));",
    "k2_fadd2": "return make_float2(a.x + b.x, a.y + b.y);",
    "k2_fsub2": "return make_float2(a.x - b.x, a.y - b.y);",
    "k2_fmul2": "return make_float2(a.x * b.x, a.y * b.y);",
}


def _match(src, i, open_ch, close_ch):
    """index just past the bracket that closes the one at src[i]"""
    depth = 0
    while True:
        c = src[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def rewrite(src, name):
    # 1. launches
    out = []
    pos = 0
    for m in re.finditer(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<", src):
        if m.start() < pos:
            continue
        cfg_end = src.index(">>>", m.end())
        cfg = src[m.end():cfg_end]
        j = cfg_end + 3
        while src[j].isspace():
            j += 1
        assert src[j] == "(", f"{name}: launch of {m.group(1)} without an argument list"
        args_end = _match(src, j, "(", ")")
        out.append(src[pos:m.start()])
        out.append(f"EMU_LAUNCH({cfg}, {m.group(1)}{src[j:args_end]})")
        pos = args_end
    out.append(src[pos:])
    src = "".join(out)
    assert "<<<" not in src, f"{name}: a launch was not rewritten"
    # 2. dynamic shared memory
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"\1* \2 = EMU_DYN_SMEM(\1);", src)
    # 3. inline-PTX helpers
    for fn, body in PTX_HELPERS.items():
        m = re.search(r"\b" + fn + r"\s*\([^)]*\)\s*\{", src)
        if not m:
            continue
        end = _match(src, m.end() - 1, "{", "}")
        src = src[:m.end()] + " " + body + " }" + src[end:]
    assert "asm volatile" not in src and "asm(" not in src, f"{name}: inline PTX without an interpreter counterpart"
    return src


def build(force=False):
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inl", ".h")))
    deps = [os.path.join(CSRC, f) for f in srcs] + [os.path.join(HERE, "cuda_runtime.h"), os.path.abspath(__file__),
                                                    os.path.join(ROOT, "include", "rg_engine.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    shutil.rmtree(SCRATCH, ignore_errors=True)
    os.makedirs(os.path.join(SCRATCH, "mp3rgain_b200", "csrc"))
    os.makedirs(os.path.join(SCRATCH, "include"))
    shutil.copy(os.path.join(ROOT, "include", "rg_engine.h"), os.path.join(SCRATCH, "include"))
    for f in srcs:
        with open(os.path.join(CSRC, f)) as fh:
            text = rewrite(fh.read(), f)
        dst = f[:-3] + ".cpp" if f.endswith(".cu") else f
        with open(os.path.join(SCRATCH, "mp3rgain_b200", "csrc", dst), "w") as fh:
            fh.write(text)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-w"] + (
           ["-fsanitize=address", "-fno-omit-frame-pointer"] if ASAN else []) + os.environ.get("RG_EMU_EXTRA", "").split() + [
           "-I", HERE, "-o", OUT, os.path.join(SCRATCH, "mp3rgain_b200", "csrc", "rg_engine.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr[-6000:])
        raise RuntimeError("simt_emu build failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
