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


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ types, extern "C" guarded) and a C program must
    link against the shared library and reach an entry point without a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    from ctclip_b200 import _lib
    _lib.load()
    so = ROOT / "ct-clip-ut_b200" / "ctclip_b200" / "libctclip_b200.so"
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "ctclip_b200.h"\n'
                   'int main(void) { printf("%d %s\\n", ctc_version(), ctc_last_error()); return 0; }\n')
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(src),
                        "-o", str(exe), str(so), f"-Wl,-rpath,{so.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0] == "100", (r.stdout, r.stderr)


def test_gemm_row_perm_is_the_documented_permutation():
    """ctc_gemm_row_perm (host-side, no GPU): accumulator column 8 j + 2 q + e of a 32-column group holds channel
    8 q + 2 j + e (include/ctclip_b200.h) - the packing Plan._pack applies to every GEMM weight."""
    import ctypes
    from ctclip_b200 import _lib
    perm = _lib.gemm_row_perm()
    assert sorted(perm) == list(range(32))
    for j in range(4):
        for q in range(4):
            for e in range(2):
                assert perm[8 * j + 2 * q + e] == 8 * q + 2 * j + e
    assert _lib.GEMM_BPERM == 0x100
