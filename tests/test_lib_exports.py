"""The C-ABI library builds, loads, and exports every symbol include/feddat_b200.h declares.
No compute calls: this runs without a GPU."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def header_functions(name="feddat_b200.h"):
    text = (ROOT / "include" / name).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(feddat_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from feddat_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/feddat_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.feddat_abi_version() == _lib.ABI_VERSION
    # probes / debug switches are NOT in the product library
    for n in _lib.DEBUG_SYMBOLS:
        assert not hasattr(lib, n), f"{n} (debug-only) leaked into the product library"


def test_debug_twin_exports_product_and_debug_symbols():
    from feddat_b200 import _lib
    dbg = _lib.load_debug()
    names = header_functions("feddat_b200_debug.h")
    assert set(names) == set(_lib.DEBUG_SYMBOLS)
    for n in list(names) + header_functions():
        assert hasattr(dbg, n), n


def test_entry_points_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from feddat_b200 import _lib
    lib = _lib.load()
    rc = lib.feddat_mkd_loss(None, None, None, None, None, 1, 100, 3.0, 0.5, 0.5, 1.0, 1, None, None)
    assert rc != 0 and len(lib.feddat_last_error()) > 0       # an error code, never a silent CPU path
