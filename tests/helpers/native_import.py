"""`nat = load_native()` for test modules.  On a CPU machine a broken C++ host is an error (the CPU suite must
fail loudly).  On a GPU box the same import failure only skips the importing module: a collection error would
abort the whole `pytest -m gpu` run, including the files that drive the C ABI without the pybind11 module."""
import pytest


def load_native():
    try:
        import pycolmap_b200 as nat
        return nat
    except ImportError:
        import torch
        if torch.cuda.is_available():
            pytest.skip("pycolmap_b200 (the pybind11 host) does not import on this box", allow_module_level=True)
        raise
