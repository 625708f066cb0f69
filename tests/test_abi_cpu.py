"""CPU-side checks of the boundary: the library builds, loads and exports every symbol the header
declares; without a GPU it refuses loudly (no fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from realtrace_b200 import api
    return api.load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "realtrace_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(rt_\w+)\s*\(", text, re.M)))


def test_header_symbols_are_all_exported(lib):
    from realtrace_b200 import api
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/realtrace_b200.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == declared


def test_struct_sizes_match_the_header():
    import ctypes as C
    from realtrace_b200 import api, scene
    assert C.sizeof(api.RtMaterial) == 40 == scene.MATERIAL_DTYPE.itemsize
    assert C.sizeof(api.RtCamera) == 64
    assert C.sizeof(api.RtRenderParams) == 40
    assert C.sizeof(api.RtFrameStats) == 104      # 7 x u64, 6 x u32, 6 x float


def test_tile_layout_is_pure_host_code(lib):
    from realtrace_b200 import api
    total, owned, tb = api.tile_layout(3840, 2160)
    assert (total, owned, tb) == (60 * 68, 60 * 68, 64 * 32 * 3)
    parts = [api.tile_layout(1920, 1080, 0, 0, r, 8)[1] for r in range(8)]
    assert sum(parts) == api.tile_layout(1920, 1080)[0]
    assert max(parts) - min(parts) <= 1


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from realtrace_b200 import api
    with pytest.raises(api.RtError) as e:
        api.Context(0)
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import, link or execute anything
    under oracle/; nothing inside the product package does."""
    pkg = os.path.join(ROOT, "realtrace_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(oracle/)|(libserial_)|(oracle_abi\.h)|(serial_port)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                continue
            text = open(os.path.join(dirpath, f)).read()
            assert not bad.search(text), f"{os.path.join(dirpath, f)} references the oracle"
