"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU and
exports every symbol include/mrs_b200.h declares; the product package has no route to the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "mrs_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"MRS_API\s+[\w\s\*]+?\b(mrs_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import mrs_b200  # noqa: F401
    from mrs_b200 import engine
    engine.build_library()
    return ctypes.CDLL(engine.LIB_PATH)


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 28
    for must in ("mrs_ratings_from_coo", "mrs_fit", "mrs_mae", "mrs_predict", "mrs_fit_similarity", "mrs_neighbors", "mrs_recommend"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_lists_match_header():
    from mrs_b200 import engine
    assert sorted(engine.EXPORTS) == declared_symbols()


def test_engine_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mrs_b200 import engine
    with pytest.raises(engine.MrsError) as ei:
        engine.Engine(0)
    assert "no CUDA device" in str(ei.value)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "movie-recommender-system_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                for needle in ("import oracle", "from oracle", "libmrs_oracle", "oracle/", "mrs_oracle"):
                    assert needle not in src, (fn, needle)


def test_sass_is_sm100a_only(lib):
    import subprocess
    from mrs_b200 import engine
    out = subprocess.run(["cuobjdump", "--list-elf", engine.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
