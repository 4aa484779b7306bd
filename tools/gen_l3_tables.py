#!/usr/bin/env python3
"""Generate the MPEG-1/2 Layer III constant tables as a C header (two copies: product + oracle).

The reference (mp3rgain) does not contain these tables -- its decoder is the symphonia crate.  The integer tables
(Huffman code lengths/symbols, scalefactor-band widths, pretab, slen, synthesis window x 65536) are ISO/IEC 11172-3 /
13818-3 constants; the only copy in the build container is the .rodata of the bundled libavcodec
(SURVEY.md Appendix D lists the offsets).  This script reads them from there, validates them (Kraft sums, unique
symbols, band widths summing to 576) and writes them out.  The float tables (|x|^(4/3), 2^(e/4), IMDCT / DCT cosines,
windows, alias butterflies) are computed here in float64 and frozen as float32 hex literals so that the CPU oracle and
the CUDA kernels consume bit-identical constants and never call a transcendental at run time.

Run only in the build container; the generated headers are committed.
"""
import glob
import math
import pathlib
import sys

import numpy as np

root = pathlib.Path(__file__).resolve().parent.parent
cands = glob.glob("/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs/libavcodec-*.so.62.11.100")
if not cands:
    sys.exit("libavcodec 62.11.100 not found; offsets below are only valid for that build")
d = np.fromfile(cands[0], np.uint8)

# ---------------------------------------------------------------- integer tables from .rodata
HUFF_SIZES = [4, 9, 9, 16, 16, 36, 36, 36, 64, 64, 64, 256, 256, 256, 256]  # ISO tables 1,2,3,5,6,7,8,9,10,11,12,13,15,16,24
ISO_IDS = [1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 24]
NH = sum(HUFF_SIZES)
hlen = d[0xf6d8a0:0xf6d8a0 + NH].astype(int)
hsym = d[0xf6d320:0xf6d320 + NH].astype(int)
off = 0
huff_off = []
for n in HUFF_SIZES:
    l = hlen[off:off + n]
    s = hsym[off:off + n]
    assert abs(sum(2.0 ** -int(x) for x in l) - 1.0) < 1e-12, "Kraft"
    assert len(set(s.tolist())) == n, "symbols unique"
    side = int(round(math.sqrt(n)))
    assert side * side == n and all((v >> 4) < side and (v & 15) < side for v in s.tolist())
    huff_off.append(off)
    off += n
# ISO table 1 known answer: (0,0)='1', (0,1)='001', (1,0)='01', (1,1)='000'
assert hlen[:4].tolist() == [3, 3, 2, 1] and hsym[:4].tolist() == [0x11, 0x01, 0x10, 0x00]

tabsel = d[0xf6d2c0:0xf6d2c0 + 64].reshape(32, 2).astype(int)
assert tabsel[0].tolist() == [0, 0] and tabsel[16].tolist() == [14, 1] and tabsel[31].tolist() == [15, 13]
c1_bits = d[0xf6d2a0:0xf6d2a0 + 32].reshape(2, 16).astype(int)
c1_lens = d[0xf6d280:0xf6d280 + 32].reshape(2, 16).astype(int)
for t in range(2):
    assert abs(sum(2.0 ** -int(x) for x in c1_lens[t]) - 1.0) < 1e-12
    codes = {(int(c1_lens[t][v]), int(c1_bits[t][v])) for v in range(16)}
    assert len(codes) == 16
assert c1_lens[0][0] == 1 and c1_bits[0][0] == 1          # quad 0000 -> '1' in table A
assert c1_bits[1].tolist() == list(range(15, -1, -1))      # table B: 4-bit inverted index
sfb_long = d[0xf6d1a0:0xf6d1a0 + 9 * 22].reshape(9, 22).astype(int)
sfb_short = d[0xf6d120:0xf6d120 + 9 * 13].reshape(9, 13).astype(int)
assert (sfb_long.sum(1) == 576).all() and (sfb_short.sum(1) == 192).all()
assert sfb_long[0].tolist() == [4, 4, 4, 4, 4, 4, 6, 6, 8, 8, 10, 12, 16, 20, 24, 28, 34, 42, 50, 54, 76, 158]
pretab = d[0xf6d0f6:0xf6d0f6 + 22].astype(int)
assert pretab.tolist() == [0] * 11 + [1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0]
slen = d[0xf6de80:0xf6de80 + 32].reshape(2, 16).astype(int)
assert slen[0].tolist() == [0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4]
lsf_nsf = d[0xf6de20:0xf6de20 + 72].reshape(6, 3, 4).astype(int)
enwin = d[0xf6e220:0xf6e220 + 257 * 4].view(np.int32).astype(np.int64)
assert enwin[0] == 0 and enwin[1] == -1 and enwin[256] == 75038

# ---------------------------------------------------------------- float tables (float64 -> float32)
def f32(a):
    return np.asarray(a, np.float64).astype(np.float32)

POW43_N = 8207  # 15 + 8191 (linbits 13) + 1
pow43 = f32(np.arange(POW43_N, dtype=np.float64) ** (4.0 / 3.0))
GAIN_MIN, GAIN_MAX = -480, 64          # quarter-step exponent range of 2^(e/4)
gain = f32(2.0 ** (np.arange(GAIN_MIN, GAIN_MAX + 1, dtype=np.float64) / 4.0))
ci = np.array([-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037])
alias_cs = f32(1.0 / np.sqrt(1.0 + ci * ci))
alias_ca = f32(ci / np.sqrt(1.0 + ci * ci))
m = np.arange(18)[:, None]
k = np.arange(18)[None, :]
dct4_18 = f32(np.cos(np.pi / 72.0 * (2 * m + 1) * (2 * k + 1)))          # u[m] = sum_k X[k] * dct4_18[m][k]
i12 = np.arange(12)[:, None]
k6 = np.arange(6)[None, :]
imdct12 = f32(np.cos(np.pi / 24.0 * (2 * i12 + 7) * (2 * k6 + 1)))        # y[i] = sum_k X[k] * imdct12[i][k]
i36 = np.arange(36)
win = np.zeros((4, 36))
win[0] = np.sin(np.pi / 36.0 * (i36 + 0.5))
win[1, :18] = np.sin(np.pi / 36.0 * (i36[:18] + 0.5)); win[1, 18:24] = 1.0
win[1, 24:30] = np.sin(np.pi / 12.0 * (i36[24:30] - 18 + 0.5)); win[1, 30:] = 0.0
win[3, :6] = 0.0; win[3, 6:12] = np.sin(np.pi / 12.0 * (i36[6:12] - 6 + 0.5)); win[3, 12:18] = 1.0
win[3, 18:] = np.sin(np.pi / 36.0 * (i36[18:] + 0.5))
win[2, :12] = np.sin(np.pi / 12.0 * (np.arange(12) + 0.5))                # short window in the first 12 slots
# IMDCT-36 output i maps to +-u[idx]: i<9: +u[i+9]; 9..26: -u[26-i]; 27..35: -u[i-27].  Fold the sign into the window.
sgn = np.where(i36 < 9, 1.0, -1.0)
uidx = np.where(i36 < 9, i36 + 9, np.where(i36 < 27, 26 - i36, i36 - 27)).astype(int)
winsgn = f32(win * sgn[None, :])
winsgn[2] = f32(win[2])                                                   # short window keeps its own sign
# Fast 18-point DCT-IV through a 9-point complex FFT (3 x 3):
#   v[n] = (X[2n] + i X[17-2n]) e^{-i pi (4n+1)/72};  V = FFT9(v);  y[k] = V[k] e^{-i pi k/18};  u[2k] = Re y[k], u[17-2k] = -Im y[k]
n9 = np.arange(9)
d4_pre = f32(np.stack([np.cos(np.pi * (4 * n9 + 1) / 72.0), np.sin(np.pi * (4 * n9 + 1) / 72.0)]))
d4_post = f32(np.stack([np.cos(np.pi * n9 / 18.0), np.sin(np.pi * n9 / 18.0)]))
e9 = np.array([1.0, 2.0, 4.0])
d4_w9 = f32(np.stack([np.cos(2 * np.pi * e9 / 9.0), -np.sin(2 * np.pi * e9 / 9.0)]))   # W9^e = (re, im), e = 1, 2, 4
sqrt3_2 = np.float32(np.sqrt(3.0) / 2.0)
# Lee DCT-II secants: level N in {32,16,8,4,2}: 1/(2cos(pi(2n+1)/(2N))), n < N/2, concatenated
sec = []
for N in (32, 16, 8, 4, 2):
    n = np.arange(N // 2)
    sec.extend((1.0 / (2.0 * np.cos(np.pi * (2 * n + 1) / (2.0 * N)))).tolist())
sec = f32(sec)
# Synthesis window D[512] from the integer half-window (x 2^-16) and its (anti)symmetry
D = np.zeros(512)
for i in range(257):
    v = enwin[i] / 65536.0
    D[i] = v
    if i:
        D[512 - i] = -v if (i & 63) else v
# out[j] = sum_{p=0..7} ( B_{2p}[ia(j)] * sa(j) * D[64p+j]  +  B_{2p+1}[ib(j)] * sb(j) * D[64p+32+j] )
# where B_n is the 32-point DCT-II of the subband vector n slots ago (polyphase).
j = np.arange(32)
ia = np.where(j < 16, j + 16, np.where(j == 16, 0, 48 - j)).astype(int)
sa = np.where(j < 16, 1.0, np.where(j == 16, 0.0, -1.0))
ib = np.where(j < 16, 16 - j, j - 16).astype(int)
sb_ = -np.ones(32)
synth_wa = f32(np.stack([sa * D[64 * p + j] for p in range(8)]))          # [8][32]
synth_wb = f32(np.stack([sb_ * D[64 * p + 32 + j] for p in range(8)]))    # [8][32]
inv_sqrt2 = f32([1.0 / math.sqrt(2.0)])[0]
# intensity stereo (MPEG-1): is_ratio = tan(is_pos*pi/12), left = r/(1+r), right = 1/(1+r); is_pos 0..6
isr = np.tan(np.arange(7) * np.pi / 12.0)
is_l = isr / (1.0 + isr); is_r = 1.0 / (1.0 + isr)
is_l[6], is_r[6] = 1.0, 0.0
is_tab = f32(np.stack([is_l, is_r]))
# MPEG-2 intensity stereo (ISO 13818-3 2.4.3.2): is_pos 0 -> (1, 1); odd -> (i0^((p+1)/2), 1); even -> (1, i0^(p/2)),
# i0 = 2^(-1/4) (intensity_scale 0) or 2^(-1/2) (intensity_scale 1).  [scale][left/right][is_pos]
is_lsf = np.ones((2, 2, 32))
for sc in range(2):
    for pp in range(1, 32):
        v = 2.0 ** (-(sc + 1) * ((pp + 1) // 2) / 4.0)
        is_lsf[sc][0 if pp & 1 else 1][pp] = v
is_lsf = f32(is_lsf)

# ---------------------------------------------------------------- emit
def carr_int(name, ctype, a, per=24):
    flat = np.asarray(a).reshape(-1).tolist()
    shape = "".join(f"[{s}]" for s in np.asarray(a).shape)
    lines = [f"static const {ctype} {name}{shape} = {{"]
    for i in range(0, len(flat), per):
        lines.append("  " + ", ".join(str(int(v)) for v in flat[i:i + per]) + ",")
    lines.append("};")
    return "\n".join(lines)

def carr_f32(name, a, per=8):
    a = np.asarray(a, np.float32)
    flat = a.reshape(-1).tolist()
    shape = "".join(f"[{s}]" for s in a.shape)
    lines = [f"L3_FCONST float {name}{shape} = {{"]
    for i in range(0, len(flat), per):
        lines.append("  " + ", ".join(float(v).hex() + "f" for v in flat[i:i + per]) + ",")
    lines.append("};")
    return "\n".join(lines)

def emit(path, guard, note):
    o = []
    o.append("/* GENERATED by tools/gen_l3_tables.py -- do not edit.\n"
             " * MPEG-1/2 Layer III constants (ISO/IEC 11172-3, 13818-3).  Integer tables validated (Kraft sum 1, unique\n"
             " * symbols, band widths sum to 576); float tables are float64 results frozen as float32 hex literals.\n"
             f" * {note}\n */")
    o.append(f"#ifndef {guard}\n#define {guard}\n#include <stdint.h>\n")
    o.append("/* float tables are constexpr under C++ so that kernels can fold single entries into instruction immediates */\n"
             "#ifdef __cplusplus\n#define L3_FCONST static constexpr\n#else\n#define L3_FCONST static const\n#endif")
    o.append("#define L3_NUM_HUFF 15")
    o.append(f"#define L3_HUFF_TOTAL {NH}")
    o.append("/* ISO table numbers of the 15 code books, their sizes and offsets into L3_HUFF_LEN/SYM */")
    o.append(carr_int("L3_HUFF_ISO_ID", "uint8_t", ISO_IDS))
    o.append(carr_int("L3_HUFF_SIZE", "uint16_t", HUFF_SIZES))
    o.append(carr_int("L3_HUFF_OFF", "uint16_t", huff_off))
    o.append("/* per code, in left-to-right tree order (code_0 = 0, code_n = code_{n-1} + 1 at len_{n-1}, left-aligned): length and (x<<4)|y */")
    o.append(carr_int("L3_HUFF_LEN", "uint8_t", hlen))
    o.append(carr_int("L3_HUFF_SYM", "uint8_t", hsym))
    o.append("/* table_select (0..31) -> {code book slot (0 = no data, else 1-based index into the 15 books), linbits} */")
    o.append(carr_int("L3_TABSEL", "uint8_t", tabsel))
    o.append("/* count1 quad tables A,B: code value and length per quad v = (v<<3|w<<2|x<<1|y) */")
    o.append(carr_int("L3_C1_BITS", "uint8_t", c1_bits))
    o.append(carr_int("L3_C1_LENS", "uint8_t", c1_lens))
    o.append("/* scalefactor band widths; rows: 44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000 */")
    o.append(carr_int("L3_SFB_LONG_W", "uint8_t", sfb_long))
    o.append(carr_int("L3_SFB_SHORT_W", "uint8_t", sfb_short))
    o.append(carr_int("L3_PRETAB", "uint8_t", pretab))
    o.append(carr_int("L3_SLEN", "uint8_t", slen))
    o.append(carr_int("L3_LSF_NSF", "uint8_t", lsf_nsf))
    o.append(f"#define L3_POW43_N {POW43_N}")
    o.append(carr_f32("L3_POW43", pow43))
    o.append(f"#define L3_GAIN_MIN ({GAIN_MIN})\n#define L3_GAIN_MAX ({GAIN_MAX})")
    o.append("/* L3_GAIN[e - L3_GAIN_MIN] = 2^(e/4) */")
    o.append(carr_f32("L3_GAIN", gain))
    o.append(carr_f32("L3_ALIAS_CS", alias_cs))
    o.append(carr_f32("L3_ALIAS_CA", alias_ca))
    o.append("/* u[m] = sum_k X[k]*L3_DCT4_18[m][k]  (18-point DCT-IV; IMDCT-36 output i = +-u[L3_IMDCT36_IDX[i]]) */")
    o.append(carr_f32("L3_DCT4_18", dct4_18))
    o.append(carr_int("L3_IMDCT36_IDX", "uint8_t", uidx))
    o.append("/* the same u[] through a 9-point complex FFT: pre-twiddle (cos, sin)(pi(4n+1)/72), W9^{1,2,4} as (re, im), post-twiddle (cos, sin)(pi k/18) */")
    o.append(carr_f32("L3_D4_PRE", d4_pre))
    o.append(carr_f32("L3_D4_W9", d4_w9))
    o.append(carr_f32("L3_D4_POST", d4_post))
    o.append(f"#define L3_SQRT3_2 {float(sqrt3_2).hex()}f")
    o.append("/* block-type windows with the IMDCT-36 sign folded in (row 2: 12-tap short window, plain) */")
    o.append(carr_f32("L3_WINSGN", winsgn))
    o.append(carr_f32("L3_IMDCT12", imdct12))
    o.append("/* Lee DCT-II secants 1/(2cos(pi(2n+1)/(2N))) for N=32 (16 values),16,8,4,2 */")
    o.append(carr_f32("L3_DCT_SEC", sec))
    o.append("/* polyphase window with signs folded: see DESIGN.md */")
    o.append(carr_f32("L3_SYNTH_WA", synth_wa))
    o.append(carr_f32("L3_SYNTH_WB", synth_wb))
    o.append(carr_int("L3_SYNTH_IA", "uint8_t", ia))
    o.append(carr_int("L3_SYNTH_IB", "uint8_t", ib))
    o.append(f"#define L3_INV_SQRT2 {float(inv_sqrt2).hex()}f")
    o.append(carr_f32("L3_IS_RATIO", is_tab))
    o.append(carr_f32("L3_IS_LSF", is_lsf))
    o.append(f"\n#endif /* {guard} */\n")
    pathlib.Path(path).write_text("\n".join(o))

emit(root / "mp3rgain_b200/csrc/l3_tables.h", "MP3RGAIN_B200_L3_TABLES_H", "Product copy (uploaded to device memory at engine creation).")
emit(root / "oracle/l3_tables_oracle.h", "RG_ORACLE_L3_TABLES_H", "Oracle copy (test infrastructure).")
print("ok", NH)
