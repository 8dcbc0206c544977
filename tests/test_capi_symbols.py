"""CPU: libdvd_b200.so loads and exports every symbol include/dvd_b200.h declares
(no compute calls — there is no GPU in the authoring container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'dvd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dvd_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from dvd_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), 'missing export: %s' % n


def test_python_binding_covers_header():
    from dvd_b200 import _lib
    assert set(_declared()) == set(_lib.SIGNATURES), set(_declared()) ^ set(_lib.SIGNATURES)


def test_load_and_version():
    from dvd_b200 import _lib
    lib = _lib.load()
    assert lib.dvd_version() >= 100
    assert lib.dvd_reproject_partials_size(1, 224, 384) > 0


def test_ops_reject_cpu_tensors():
    import pytest
    import torch
    from dvd_b200 import ops
    with pytest.raises(ValueError):
        ops.unproject_fwd(torch.zeros(1, 1, 8, 8), torch.zeros(1, 48), 1)


def test_ctypes_struct_layouts_match_the_library():
    """the four structs that cross the C ABI by pointer: the ctypes mirrors of dvd_b200/_lib.py have the sizes the compiled
    library reports for the C definitions of include/dvd_b200.h"""
    from dvd_b200 import _lib
    lib = _lib.load()
    for which, cls in enumerate((_lib.LossCfg, _lib.MlpCfg, _lib.ConvDesc, _lib.PackItem)):
        assert lib.dvd_struct_size(which) == ctypes.sizeof(cls), (cls.__name__, lib.dvd_struct_size(which), ctypes.sizeof(cls))
    assert lib.dvd_struct_size(99) == -1
