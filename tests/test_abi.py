"""CPU tests: the C-ABI library loads and exports every symbol include/rg_engine.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rg_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_match_binding():
    from mp3rgain_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from mp3rgain_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_status_strings_match_reference_messages():
    from mp3rgain_b200 import _lib
    lib = _lib.load()
    assert lib.rg_status_str(_lib.RG_ERR_NO_AUDIO_TRACK) == b"No audio track found"      # src/replaygain.rs:834
    assert lib.rg_status_str(_lib.RG_ERR_UNKNOWN_SAMPLE_RATE) == b"Unknown sample rate"   # :857
    assert lib.rg_status_str(_lib.RG_ERR_UNSUPPORTED_RATE).startswith(b"Unsupported sample rate")  # :869-872
    assert lib.rg_status_str(_lib.RG_ERR_PROBE).startswith(b"Failed to probe format")    # :822


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the engine must refuse to run, not compute on the host."""
    import mp3rgain_b200 as m
    if m.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(m.ReplayGainError) as ei:
        m.Engine()
    assert ei.value.status == -2


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under mp3rgain_b200/ or include/ may reference it."""
    bad = []
    for base in ("mp3rgain_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".so", ".pyc", ".log")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"#include[^\n]*oracle|import[^\n]*oracle|oracle_lib|librg_oracle|\brgo_|\bmp3o_|oracle/", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
