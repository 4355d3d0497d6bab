"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol that
include/onerf.h declares; argument validation fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from object_nerf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "onerf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(onerf_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"libonerf_sm100.so does not export {n}"
    from object_nerf_b200 import _lib
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_sizes(lib):
    assert lib.onerf_abi_version() == 1
    v, p = lib.onerf_packed_weights_bytes(1), lib.onerf_packed_weights_bytes(0)
    # fp32 section + bf16 stage images; see object_nerf_b200/csrc/layout.h
    assert v > p > 4 * 704840
    assert v % 1024 == 0 and p % 1024 == 0


def test_ctx_create_fails_loudly_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.onerf_ctx_create(0, ctypes.byref(h))
    assert rc != 0
    assert len(lib.onerf_last_error()) > 0


def test_product_refuses_cpu_tensors():
    import torch
    from object_nerf_b200 import engine
    with pytest.raises(RuntimeError):
        engine.sample_coarse(torch.zeros(4, 8), 8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "object_nerf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn
