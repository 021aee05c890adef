"""ctypes loader for liboracle.so (oracle_match.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle_match.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_fast_isa.restype = ctypes.c_char_p
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2 and a.shape[1] == 128, a.shape
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def match_bruteforce(d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=True):
    """Literal M1->M2->M3 (materialises the distance matrix).  Returns [m x 2] uint32."""
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros((max(1, min(len(d1), len(d2)) if cross_check else len(d1)), 2), np.uint32)
    out = np.zeros((max(1, len(d1)), 2), np.uint32)
    n = _lib().orc_match_bruteforce(_p(d1), len(d1), _p(d2), len(d2), ctypes.c_float(max_ratio),
                                    ctypes.c_float(max_distance), int(bool(cross_check)), _p(out))
    assert n >= 0
    return out[:n].copy()


def match_guided(d1, kp1, d2, kp2, kind, model, max_error, max_ratio=0.8, max_distance=0.7,
                 cross_check=True):
    """G1.  kind 0 = E/F (Sampson), 1 = H (transfer).  model: 3x3."""
    d1, d2 = _u8(d1), _u8(d2)
    kp1 = np.ascontiguousarray(kp1, np.float32)
    kp2 = np.ascontiguousarray(kp2, np.float32)
    M = np.ascontiguousarray(np.asarray(model, np.float64).reshape(9).astype(np.float32))
    out = np.zeros((max(1, len(d1)), 2), np.uint32)
    n = _lib().orc_match_guided(_p(d1), _p(kp1), len(d1), _p(d2), _p(kp2), len(d2), int(kind), _p(M),
                                ctypes.c_float(max_error), ctypes.c_float(max_ratio),
                                ctypes.c_float(max_distance), int(bool(cross_check)), _p(out))
    assert n >= 0
    return out[:n].copy()


def fast_match_pair(d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=True):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros((max(1, len(d1)), 2), np.uint32)
    n = _lib().orc_fast_match_pair(_p(d1), len(d1), _p(d2), len(d2), ctypes.c_float(max_ratio),
                                   ctypes.c_float(max_distance), int(bool(cross_check)), _p(out))
    return out[:n].copy()


def fast_match_pairs(desc_packed, n_feat, pairs, max_ratio=0.8, max_distance=0.7, cross_check=True,
                     n_threads=None):
    """Threaded batch.  desc_packed: [sum(n_feat) x 128] u8.  Returns list of [m x 2] uint32."""
    desc_packed = _u8(desc_packed)
    n_feat = np.ascontiguousarray(n_feat, np.int32)
    offsets = np.zeros(len(n_feat), np.int64)
    offsets[1:] = np.cumsum(n_feat[:-1].astype(np.int64))
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    stride = int(n_feat.max()) if len(n_feat) else 1
    out = np.zeros((len(pairs), max(1, stride), 2), np.uint32)
    counts = np.zeros(len(pairs), np.int32)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    _lib().orc_fast_match_pairs(_p(desc_packed), _p(offsets), _p(n_feat), _p(pairs),
                                ctypes.c_int64(len(pairs)), ctypes.c_float(max_ratio),
                                ctypes.c_float(max_distance), int(bool(cross_check)), int(n_threads),
                                _p(out), ctypes.c_int64(stride), _p(counts))
    return [out[k, :counts[k]].copy() for k in range(len(pairs))]


def fast_isa():
    return _lib().orc_fast_isa().decode()


def acos_lut():
    lut = np.zeros(262145, np.float32)
    _lib().orc_acos_lut(_p(lut))
    return lut
