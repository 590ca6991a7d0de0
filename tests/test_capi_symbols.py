"""The C-ABI library loads without a GPU and exports every symbol include/fhe_b200.h declares (and the
ctypes table in fhe_study_b200/_capi.py names exactly those).  No compute calls here."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("fhe_b200.h", "fhe_b200_file.h"))
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fhe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import fhe_study_b200._capi as capi

    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(capi.lib, n), f"{n} declared in fhe_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fhe_study_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "fhe_oracle" not in txt, f


def test_rust_bindings_cover_every_symbol():
    # rust/fhe-b200-sys/src/lib.rs is generated from the header (tools/gen_rust_bindings.py); keep them in sync
    names = _header_functions()
    rs = open(os.path.join(ROOT, "rust", "fhe-b200-sys", "src", "lib.rs")).read()
    bound = sorted(set(re.findall(r"pub fn (fhe_[a-z0-9_]+)\(", rs)))
    assert bound == names
