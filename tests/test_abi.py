"""CPU: the drop-in boundary.  include/unidisc_b200.h is the contract; the built library must export every symbol it declares,
the ctypes binding (unidisc_b200/_lib.py) must describe every prototype with the right arity and C types, and the product path
must fail loudly (no CPU / eager fallback) when the library is missing or tensors are not on a GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "unidisc_b200.h")


def _prototypes():
    """name -> list of parameter C types, parsed from the header (comments stripped)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"\bint\s+(ud_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = params
    return protos


def _ctype_of(param):
    if "*" in param:
        return ctypes.c_void_p
    t = param.rsplit(" ", 1)[0].replace("const ", "").strip()
    return {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "uint64_t": ctypes.c_uint64,
            "int64_t": ctypes.c_int64}[t]


@pytest.fixture(scope="module")
def built_lib():
    from unidisc_b200 import build
    path = build.build()                      # cached by source digest; nvcc cross-compiles without a GPU
    assert os.path.exists(path)
    return path


def test_header_declares_what_the_binding_binds():
    from unidisc_b200 import _lib
    protos = _prototypes()
    assert len(protos) >= 28
    assert set(protos) == set(_lib._SIGS), (set(protos) ^ set(_lib._SIGS))
    assert set(_lib.EXPORTED_SYMBOLS) == set(protos)
    for name, params in protos.items():
        sig = _lib._SIGS[name]
        assert len(sig) == len(params), f"{name}: header has {len(params)} parameters, ctypes binding {len(sig)}"
        for k, (p, c) in enumerate(zip(params, sig)):
            assert _ctype_of(p) is c, f"{name} parameter {k} ({p!r}): binding says {c.__name__}"
    # no torch / C++ types leak into the boundary
    code = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    assert 'extern "C"' in code and "at::" not in code and "torch" not in code.lower() and "std::" not in code


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    h = ctypes.CDLL(built_lib)
    for name in _prototypes():
        assert hasattr(h, name), f"{built_lib} does not export {name}"
    h.ud_abi_version.restype = ctypes.c_int
    assert h.ud_abi_version() == 6            # bumped whenever a prototype changes (include/unidisc_b200.h)


def test_adaln_struct_layout_matches_header():
    """`ud_adaln` (time conditioning) is passed by pointer: the ctypes mirror must have the header's fields in order."""
    from unidisc_b200 import _lib
    src = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct ud_adaln \{(.*?)\} ud_adaln;", src, flags=re.S).group(1)
    fields = [f.strip().split()[-1].lstrip("*") for f in body.split(";") if f.strip()]
    assert fields == [n for n, _ in _lib.AdaLN._fields_]
    assert ctypes.sizeof(_lib.AdaLN) == 5 * 8 + 8 + 8 + 3 * 8 + 8     # 5 pointers, ld, int (+pad), 3 pointers, ld_d


def test_no_cpu_fallback(monkeypatch, built_lib):
    from unidisc_b200 import _lib
    from unidisc_b200.config import make_config
    from unidisc_b200.dit import DIT
    # a missing library is an error, not a silent eager path
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "unidisc_b200", "does_not_exist.so"))
    with pytest.raises(_lib.UnidiscB200Error):
        _lib.lib()
    monkeypatch.undo()
    # CPU tensors / CPU modules are rejected by the backbone
    cfg = make_config("small", hidden_size=128, n_blocks=1, n_heads=2, txt_length=8, img_length=16, image_vocab_size=15, text_vocab_size=17)
    m = DIT(cfg, vocab_size=32, text_vocab_size=17, mask_index=16)
    ids = torch.zeros(1, 24, dtype=torch.int64)
    with pytest.raises(_lib.UnidiscB200Error):
        m(ids, None, modality=torch.zeros_like(ids))
    # unsupported reference variants raise instead of silently computing something else
    bad = make_config("small", hidden_size=128, n_blocks=1, n_heads=2, txt_length=8, img_length=16, model__norm_type="layernorm")
    with pytest.raises(NotImplementedError):
        DIT(bad, vocab_size=32, text_vocab_size=17, mask_index=16)
