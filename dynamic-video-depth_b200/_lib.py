"""ctypes binding of libdvd_b200.so (C ABI declared in include/dvd_b200.h).

The library is mandatory: importing the compute ops without it raises — there is no CPU or
PyTorch fallback on the product path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdvd_b200.so')

c_f32p = ctypes.c_void_p


class LossCfg(ctypes.Structure):
    """struct dvd_loss_cfg (include/dvd_b200.h)."""
    _fields_ = [('midas', ctypes.c_int), ('warm', ctypes.c_int), ('disp_mode', ctypes.c_int),
                ('second_is_disp', ctypes.c_int), ('flow_mul', ctypes.c_float), ('disp_mul', ctypes.c_float)]


class MlpDims(ctypes.Structure):
    """struct dvd_mlp_dims (include/dvd_b200.h)."""
    _fields_ = [('n_freq_xyz', ctypes.c_int), ('n_freq_t', ctypes.c_int), ('time_dependent', ctypes.c_int),
                ('width', ctypes.c_int), ('n_hidden', ctypes.c_int), ('out_scale', ctypes.c_float)]


_I, _F, _P = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
_L = ctypes.c_longlong

# name -> argtypes; restype is int unless listed in _RESTYPES
SIGNATURES = {
    'dvd_version': [],
    'dvd_last_error': [],
    'dvd_reproject_partials_size': [_I, _I, _I],
    'dvd_unproject_fwd': [_P, _P, _P, _I, _I, _I, _I, _P],
    'dvd_unproject_bwd': [_P, _P, _P, _I, _I, _I, _I, _P],
    'dvd_reproject_loss_fwd': [_P, _P, _P, _P, _P, _P, ctypes.POINTER(LossCfg), _P, _P, _I, _I, _I, _P],
    'dvd_reproject_loss_bwd': [_P, _P, _P, _P, _P, _P, ctypes.POINTER(LossCfg), _P, _F, _P, _P, _P, _I, _I, _I, _P],
    'dvd_reproject_materialize': [_P] * 14 + [_I, _I, _I, _P],
    'dvd_selftest_umma': [_P, _P, _P, _I, _I, _I, _I, _P],
}
_RESTYPES = {'dvd_last_error': ctypes.c_char_p}

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises RuntimeError if the .so is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libdvd_b200.so not found at %s — build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(or dynamic-video-depth_b200/csrc/build.sh). dvd_b200 has no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dvd_last_error()
        raise RuntimeError('%s failed (rc=%d): %s' % (what, rc, (msg or b'').decode()))
