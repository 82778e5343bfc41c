"""The C-ABI library loads without a GPU and exports every symbol include/diqt.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "diqt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(diqt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from diffusioniqt_b200 import lib
    if not os.path.isfile(lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    cdll = ctypes.CDLL(lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(cdll, name), f"{name} declared in include/diqt.h but not exported"


def test_binding_covers_header():
    from diffusioniqt_b200 import lib
    assert sorted(lib.EXPORTED_SYMBOLS) == _declared_symbols()
    handle = lib.load()
    assert handle.diqt_abi_version() == lib.ABI_VERSION
    assert lib.launch_count() == 0 or lib.launch_count() > 0


def test_bad_arguments_are_reported_without_a_gpu():
    from diffusioniqt_b200 import lib
    h = lib.load()
    d = lib.ConvDesc(mode=7, dtype=0, impl=0, n=1, d0=4, d1=4, d2=4, c_in=16, ld_in=16, c_out=16, ld_out=16, flags=0)
    nbytes = ctypes.c_size_t(0)
    rc = h.diqt_conv_packed_bytes(ctypes.byref(d), ctypes.byref(nbytes))
    assert rc == -1 and b"bad mode" in h.diqt_last_error()
    with pytest.raises(lib.DiqtError):
        lib.check(rc, "conv_packed_bytes")
    d.mode, d.impl, d.dtype = lib.CONV_K3, lib.IMPL_TC, lib.F32   # tcgen05 kernel is bf16 only
    assert h.diqt_conv_packed_bytes(ctypes.byref(d), ctypes.byref(nbytes)) == -3


def test_missing_library_fails_loudly():
    """No CPU / PyTorch fallback: without the built kernel library the product path raises instead of computing something else."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['DIQT_LIB_PATH'] = '/nonexistent/libdiqt_b200.so'\n"
            "from diffusioniqt_b200 import lib\n"
            "try:\n    lib.load()\nexcept lib.DiqtError as e:\n    print('RAISED', 'no CPU or PyTorch fallback' in str(e))\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "RAISED True" in out.stdout, out.stdout + out.stderr


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under diffusioniqt_b200/ may import it."""
    import glob
    for path in glob.glob(os.path.join(ROOT, "diffusioniqt_b200", "**", "*.py"), recursive=True):
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, path
