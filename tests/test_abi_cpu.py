"""CPU-side checks of the C-ABI: the library builds, loads, and exports every declared symbol."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "slime_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slime_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from slime_b200 import _lib, build

    build.build()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for dtype, code in (("bf16", 0), ("fp16", 2)):  # the two builds of the same sources export the same ABI
        lib = _lib.load(dtype)
        for name in declared:
            assert hasattr(lib, name), f"{name} declared in include/slime_b200.h but not exported ({dtype} build)"
            assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in slime_b200/_lib.py"
        assert lib.slime_version() == 1 and lib.slime_elem_dtype() == code


def test_ctx_create_fails_loudly_without_gpu():
    import ctypes as C

    import torch

    from slime_b200 import _lib

    if torch.cuda.is_available():
        return
    lib = _lib.load()
    ctx = C.c_void_p()
    desc = _lib.ModelDesc()
    rc = lib.slime_ctx_create(C.byref(ctx), 0, C.byref(desc))
    assert rc != 0
    assert len(_lib.last_error()) > 0


def test_model_desc_layout_matches_header():
    """Field order/types of the ctypes mirror follow the C struct declaration."""
    from slime_b200 import _lib

    text = (ROOT / "include" / "slime_b200.h").read_text()
    body = text.split("typedef struct slime_model_desc {")[1].split("} slime_model_desc;")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int32_t|int64_t|uint32_t|float)\s+([a-z_0-9]+)\s*;", body)
    assert fields == [f[0] for f in _lib.ModelDesc._fields_]
