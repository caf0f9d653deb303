"""CPU: the C-ABI shared library loads and exports every symbol include/conan_b200.h declares
(no compute calls; there is no GPU in the build container), and refuses to run without one."""
import ctypes
import os
import re

import pytest
import torch

from conan_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build(verbose=False)
    return _lib.load()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(os.path.dirname(build.HERE), "include", "conan_b200.h")).read()
    declared = set(re.findall(r"\b(conan_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"conan_engine", "conan_config", "conan_conv_params"}
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


def test_struct_sizes_match_header(lib):
    assert ctypes.sizeof(_lib.ConanConfig) == lib.conan_sizeof_config()
    assert ctypes.sizeof(_lib.ConvParams) == lib.conan_sizeof_conv_params()
    assert lib.conan_abi_version() == _lib.ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(lib):
    from conan_b200.engine import make_config
    cfg = make_config(max_slots=2, max_ref_frames=64)
    h = ctypes.c_void_p()
    rc = lib.conan_engine_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.conan_last_error()
    from conan_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine({}, {}, {}, cfg)


def test_integration_md_config_mirror_matches_header():
    """the ctypes mirror shown to maintainers in INTEGRATION.md must be the header's struct, field for field
    (a same-size struct with stale tail fields silently selects the slow engines)."""
    import ctypes as C
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "conan_b200.h")).read()
    body = hdr[hdr.index("typedef struct conan_config {"):hdr.index("} conan_config_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"int32_t\s+(\w+)(?:\[(\d+)\])?;", body)
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    snippet = md[md.index("class Cfg(C.Structure):"):md.index("assert lib.conan_sizeof_config()")]
    ns = {"C": C}
    exec(snippet, ns)
    got = [(n, (t._length_ if hasattr(t, "_length_") else 0)) for n, t in ns["Cfg"]._fields_]
    assert got == [(n, int(k) if k else 0) for n, k in fields]
    from conan_b200 import _lib
    assert [f[0] for f in _lib.ConanConfig._fields_] == [n for n, _ in fields]
    assert C.sizeof(ns["Cfg"]) == C.sizeof(_lib.ConanConfig) == _lib.load().conan_sizeof_config()
