"""CPU: the C-ABI library loads without a GPU and exports every symbol include/ctclip_b200.h
declares, with the parameter counts the ctypes binding assumes (no compute calls here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT / "include" / "ctclip_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(ctc_\w+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(2)] = n
    return out


def test_library_exports_every_declared_symbol():
    from ctclip_b200 import _lib
    lib = _lib.load()
    decl = _declared()
    assert len(decl) >= 25
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        if name in _lib.SIGNATURES:
            assert len(_lib.SIGNATURES[name]) == nargs, (name, len(_lib.SIGNATURES[name]), nargs)
        else:
            assert name in _lib.OTHER_SYMBOLS, f"{name} has no ctypes signature"
    for name in _lib.SIGNATURES:
        assert name in decl, f"{name} bound by ctypes but not declared in the header"
    assert lib.ctc_version() == 100
    assert isinstance(lib.ctc_last_error(), bytes)


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "ct-clip-ut_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "oracle" not in re.sub(r"#.*", "", src).replace("oracle/", ""), f"{f} references the oracle"
