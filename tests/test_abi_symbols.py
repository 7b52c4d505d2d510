"""CPU-side check: libpkwhir.so loads and exports every symbol include/pkwhir.h declares
(no compute calls without a GPU), and the package fails loudly without a device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pkwhir.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import provekit_b200 as pk
    pk.build()
    L = pk.lib()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pkwhir.h but not exported"
    assert b"sm_100a" in L.pk_version()


def test_no_cpu_fallback_without_device():
    import provekit_b200 as pk
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pk.PkError) as e:
        pk.Context(0)
    assert e.value.code == -2  # PK_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (tier rule 3)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "provekit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "pk_oracle.h" not in txt \
                    and "libpkoracle" not in txt, f"{f} references the oracle"
