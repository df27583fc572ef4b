"""The C-ABI shared library loads and exports every symbol include/dysturbd.h declares (no compute: no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dysturbd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dys_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dysturbd_b200 import _build, engine
    lib_path = _build.build()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(engine.EXPORTS) == syms
    lib.dys_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.dys_version()


def test_params_struct_matches_header():
    from dysturbd_b200.engine import _Params
    # 4*8 + 8*4 + 2*8 + 64*8 + 8 + 2*4 + 2*4 + 2*8 + 2*8
    assert ctypes.sizeof(_Params) == 32 + 32 + 16 + 512 + 8 + 8 + 8 + 16 + 16 + 48


def test_no_gpu_means_loud_failure():
    """no CUDA device -> dys_create fails with an error string; the Python engine raises (no CPU fallback)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dysturbd_b200.engine import DysError, Engine
    from dysturbd_b200.population import EpiParams
    from dysturbd_b200.synth import synth_population
    with pytest.raises(DysError) as e:
        Engine(synth_population(2048, seed=0), EpiParams())
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dysturbd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libdys_oracle" not in txt, f
