"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/doc2tex_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "doc2tex_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d2t_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(built_lib):
    from doc2tex_b200 import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    for sym in header_symbols():
        assert hasattr(built_lib, sym), sym
    assert b"sm_100a" in built_lib.d2t_version()


def test_config_struct_layout(built_lib):
    from doc2tex_b200 import _lib
    src = open(os.path.join(ROOT, "include", "doc2tex_b200.h")).read()
    body = re.search(r"typedef struct d2t_config \{(.*?)\} d2t_config;", src, re.S).group(1)
    fields = re.findall(r"int32_t\s+([a-z_0-9]+);", body)
    assert fields == [n for n, _ in _lib.Config._fields_]
    assert C.sizeof(_lib.Config) == 4 * len(fields)


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from doc2tex_b200 import _lib, synth
    from doc2tex_b200.engine import Engine, EngineError, config_from_opt
    cfg = config_from_opt(synth.make_config("TFM"))
    h = C.c_void_p()
    rc = built_lib.d2t_create(C.byref(cfg), 0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in built_lib.d2t_last_error(None)
    with pytest.raises(EngineError):
        Engine(synth.make_config("TFM"), "cuda:0")
    bad = _lib.Config()
    bad.struct_size = 4
    assert built_lib.d2t_create(C.byref(bad), 0, C.byref(h)) != 0


def test_sass_has_blackwell_tensor_and_tma_instructions(built_lib):
    """The shipped library really contains tcgen05 / TMEM / TMA code (UTC*MMA, LDTM, UTMALDG)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from doc2tex_b200 import _lib
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass or "SM100" in sass.upper()


def test_every_engine_option_is_documented_in_the_header():
    """d2t_set_option is the only switchboard of the engine (nothing is read from the environment): every key it accepts must
    be described in include/doc2tex_b200.h, and the sources must not read environment variables."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "doc2tex_b200", "csrc", "engine.cu")).read()
    hdr = open(os.path.join(root, "include", "doc2tex_b200.h")).read()
    seg = src[src.index("int d2t_set_option("):]
    seg = seg[: seg.index("\nint d2t_", 10)]
    keys = re.findall(r'k == "([a-z_0-9]+)"', seg)
    assert len(keys) >= 20
    missing = [k for k in keys if f'"{k}"' not in hdr]
    assert not missing, f"options missing from the header: {missing}"
    csrc = os.path.join(root, "doc2tex_b200", "csrc")
    for f in os.listdir(csrc):
        text = open(os.path.join(csrc, f)).read()
        assert text.count("getenv") <= (1 if f == "engine.cu" else 0), f   # D2T_DBG_ACT inside the debug GEMM bench hook only
