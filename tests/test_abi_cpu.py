"""CPU-side checks of the C ABI: the library builds, loads, exports every symbol include/vjf_b200.h
declares, and its host-only entry points behave (no kernel is launched here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vjf_b200 import _lib
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vjf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vjf_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from vjf_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vjf_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_layout_arithmetic(lib):
    from vjf_b200 import _lib
    cfg = _lib.make_config(200, 3, 0, 50, [64], "poisson", 4096)
    lay = _lib.get_layout(cfg)
    assert lay.lik_logvar == 0 and lay.dec_w == 32
    assert lay.mlp_w[0] >= lay.dec_b + 200
    assert lay.mlp_b[0] - lay.mlp_w[0] >= 206 * 64
    assert lay.n_train % 32 == 0 and lay.prior_mean == lay.n_train
    assert lay.w_precision - lay.w_chol >= 2500 and lay.total > lay.tr_n
    # struct sizes agree with the header (catches drift between _lib.py and vjf_b200.h)
    assert C.sizeof(_lib.Config) == 4 * (5 + 4 + 2)
    assert C.sizeof(_lib.Layout) == 8 * (3 + 8 + 4 + 12)


def test_bad_configurations_are_rejected_loudly(lib):
    from vjf_b200 import _lib
    with pytest.raises(RuntimeError, match="unsupported configuration"):
        _lib.get_layout(_lib.make_config(10, 17, 0, 5, [4], "poisson", 1))  # xdim > 16
    with pytest.raises(ValueError):
        _lib.make_config(10, 2, 0, 5, [], "poisson", 1)
    with pytest.raises(KeyError):
        _lib.make_config(10, 2, 0, 5, [4], "binomial", 1)
    assert lib.vjf_version() == 100
    assert lib.vjf_launch_count() == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse to construct a model."""
    import torch
    from vjf_b200.model import VJF
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VJF.make_model(10, 2, 0, 5, [4])


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vjf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), (dirpath, f)
