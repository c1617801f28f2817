"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mmvid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmvid_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from mmvid_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mmvid_b200.h but not exported"


def test_ctypes_signature_table_matches_header():
    from mmvid_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    lib = _lib.load()
    assert lib.mmvid_version() >= 100
    assert isinstance(lib.mmvid_last_error(), bytes)


def test_sass_contains_blackwell_tensor_core_and_tma_instructions():
    from mmvid_b200 import build
    path = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass.replace("UTCHMMA", "UTCMMA"), "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA tensor loads missing from SASS"
    assert "LDTM" in sass and "STTM" in sass, "tcgen05.ld/st missing from SASS"


def test_product_path_fails_loudly_without_gpu_or_library():
    import pytest
    import torch
    from mmvid_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.layernorm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mmvid_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
