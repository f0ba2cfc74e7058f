"""CPU: the C-ABI library loads and exports every symbol include/nma_b200.h declares; host config mirrors."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nma_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from viforssms_b200 import build, lib
    build.build()
    L = lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert sorted(lib.EXPORTS) == names


def test_config_struct_matches_header_size():
    """sizeof(struct nma_config): 16 int32 + 2*32 int32 + double + 4 floats + 2 int32, natural alignment."""
    from viforssms_b200.config import CConfig
    assert ctypes.sizeof(CConfig) == 16 * 4 + 2 * 32 * 4 + 8 + 4 * 4 + 2 * 4


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from viforssms_b200 import lib
    from viforssms_b200.config import ar_config
    from viforssms_b200.engine import NMAEngine
    with pytest.raises(lib.NMAError):
        NMAEngine(ar_config())
    # and straight through the C-ABI: nma_create refuses without a device, it does not fall back
    L = lib.load()
    h = ctypes.c_void_p()
    c = ar_config().to_c()
    assert L.nma_create(ctypes.byref(c), ctypes.byref(h)) != 0
    assert b"no CUDA device" in L.nma_last_error() or b"CUDA" in L.nma_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "viforssms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), f
    for f in ("main.py", "AR_dat_gen.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert "import oracle" not in open(p).read() and "from oracle" not in open(p).read()
