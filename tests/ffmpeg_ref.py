"""Independent cross-check decoder: FFmpeg's `mp3float` from the libavcodec bundled with opencv-python-headless,
driven through ctypes.  NOT the oracle and not a dependency of the product: tests that use it
skip when the libraries are absent."""
import ctypes as C
import glob
import os

import numpy as np

_LIBDIR = None
for cand in glob.glob("/opt/prime-rl/.venv/lib/python3.*/site-packages/opencv_python_headless.libs"):
    _LIBDIR = cand
_av = None


def available():
    return _LIBDIR is not None and bool(glob.glob(os.path.join(_LIBDIR, "libavcodec-*")))


def _load():
    global _av
    if _av is not None:
        return _av
    def ld(prefix):
        p = glob.glob(os.path.join(_LIBDIR, prefix + "-*"))
        if not p:
            return None
        return C.CDLL(p[0], mode=C.RTLD_GLOBAL)
    for pre in ("libdrm", "libcrypto", "libssl", "libvpx", "libaom", "libpng16"):
        ld(pre)
    avutil = ld("libavutil")
    ld("libswresample")
    avcodec = ld("libavcodec")
    _av = (avutil, avcodec)
    return _av


class AVPacket(C.Structure):
    # lavc 62: buf, pts, dts, data, size, stream_index, flags, side_data, side_data_elems, duration, pos, opaque, opaque_ref, time_base
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p), ("size", C.c_int),
                ("stream_index", C.c_int), ("flags", C.c_int), ("side_data", C.c_void_p), ("side_data_elems", C.c_int),
                ("duration", C.c_int64), ("pos", C.c_int64), ("opaque", C.c_void_p), ("opaque_ref", C.c_void_p),
                ("tb_num", C.c_int), ("tb_den", C.c_int)]


def decode_frames(frames):
    """frames: list of bytes objects, one MPEG audio frame each (Info frame excluded by the caller).
    Returns planar float32 array [nch, n]."""
    avutil, avcodec = _load()
    avcodec.avcodec_find_decoder_by_name.restype = C.c_void_p
    avcodec.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_alloc_context3.restype = C.c_void_p
    avcodec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    avcodec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    avutil.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    avcodec.av_packet_alloc.restype = C.POINTER(AVPacket)
    avutil.av_frame_alloc.restype = C.c_void_p
    avcodec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    avutil.av_malloc.restype = C.c_void_p
    avutil.av_malloc.argtypes = [C.c_size_t]
    avutil.av_free.argtypes = [C.c_void_p]
    avutil.av_frame_unref.argtypes = [C.c_void_p]
    dec = avcodec.avcodec_find_decoder_by_name(b"mp3float")
    assert dec
    ctx = avcodec.avcodec_alloc_context3(dec)
    avutil.av_opt_set(ctx, b"flags2", b"+skip_manual", 0)
    assert avcodec.avcodec_open2(ctx, dec, None) == 0
    pkt = avcodec.av_packet_alloc()
    frame = avutil.av_frame_alloc()
    out = [[], []]
    nch_seen = 0
    for fb in frames:
        n = len(fb)
        buf = avutil.av_malloc(n + 64)
        C.memset(buf, 0, n + 64)
        C.memmove(buf, fb, n)
        pkt.contents.data = buf
        pkt.contents.size = n
        rc = avcodec.avcodec_send_packet(ctx, pkt)
        while rc == 0 and avcodec.avcodec_receive_frame(ctx, frame) == 0:
            # lavu 60.8: AVFrame.data[8] at offset 0, nb_samples int32 at offset 112
            data = (C.c_void_p * 8).from_address(frame)
            nb = C.c_int.from_address(frame + 112).value
            for c in range(2):
                if data[c]:
                    out[c].append(np.ctypeslib.as_array((C.c_float * nb).from_address(data[c])).copy())
                    nch_seen = max(nch_seen, c + 1)
            avutil.av_frame_unref(frame)
        pkt.contents.data = None
        pkt.contents.size = 0
        avutil.av_free(buf)
    planes = [np.concatenate(out[c]) if out[c] else np.zeros(0, np.float32) for c in range(nch_seen)]
    return np.stack(planes) if planes else np.zeros((0, 0), np.float32)


def split_frames(img):
    """Minimal MPEG-1/2 L3 frame splitter for the harness (ID3v2 skip, Info/Xing frame dropped)."""
    d = bytes(img)
    pos = 0
    if d[:3] == b"ID3":
        pos = 10 + ((d[6] << 21) | (d[7] << 14) | (d[8] << 7) | d[9])
    br1 = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
    sr1 = [44100, 48000, 32000]
    frames = []
    first = True
    while pos + 4 <= len(d):
        ver = (d[pos + 1] >> 3) & 3   # 3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5
        if d[pos] != 0xFF or (d[pos + 1] & 0xE0) != 0xE0 or ((d[pos + 1] >> 1) & 3) != 1 or ver == 1:
            pos += 1
            continue
        bi, si = d[pos + 2] >> 4, (d[pos + 2] >> 2) & 3
        if bi in (0, 15) or si == 3:
            pos += 1
            continue
        if ver == 3:
            flen = 144 * br1[bi] * 1000 // sr1[si] + ((d[pos + 2] >> 1) & 1)
        else:
            br2 = [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]
            flen = 72 * br2[bi] * 1000 // (sr1[si] >> (1 if ver == 2 else 2)) + ((d[pos + 2] >> 1) & 1)
        if pos + flen > len(d):
            break
        fb = d[pos:pos + flen]
        mono = (d[pos + 3] >> 6) == 3
        o = 4 + (0 if d[pos + 1] & 1 else 2) + ((17 if mono else 32) if ver == 3 else (9 if mono else 17))
        if first and fb[o:o + 4] in (b"Xing", b"Info"):
            first = False
            pos += flen
            continue
        first = False
        frames.append(fb)
        pos += flen
    return frames
