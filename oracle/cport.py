"""ctypes loader for the C restatement (oracle/zkp_oracle.c) -- test infrastructure."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzkp_oracle.so")
_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_ntt.argtypes = [_u64p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int]
        _lib.oracle_ntt.restype = ctypes.c_int
        _lib.oracle_msm_g1.argtypes = [_u64p, _u64p, ctypes.c_size_t, _u64p, ctypes.c_int]
        _lib.oracle_msm_g1.restype = ctypes.c_int
        _lib.oracle_g1_fixed_base_mul.argtypes = [_u64p, _u64p, ctypes.c_size_t, _u64p,
                                                  ctypes.c_int]
        _lib.oracle_g1_fixed_base_mul.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_u64p)


def num_threads():
    return lib().oracle_num_threads()


def ntt(data_mont, k, inverse=False, coset=False, len_in=None, nthreads=0):
    """data_mont: (m,4) uint64 Montgomery limbs, m <= 2^k. Returns a new (2^k,4) array."""
    n = 1 << k
    data_mont = np.asarray(data_mont, dtype=np.uint64).reshape(-1, 4)
    m = data_mont.shape[0] if len_in is None else len_in
    buf = np.zeros((n, 4), dtype=np.uint64)
    buf[:m] = data_mont[:m]
    rc = lib().oracle_ntt(_p(buf), m, k, int(inverse), int(coset), nthreads)
    assert rc == 0, rc
    return buf


def msm_g1(bases, scalars_mont, nthreads=0):
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 12)
    scalars_mont = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    n = scalars_mont.shape[0]
    assert bases.shape[0] >= n
    out = np.zeros(12, dtype=np.uint64)
    rc = lib().oracle_msm_g1(_p(bases), _p(scalars_mont), n, _p(out), nthreads)
    assert rc == 0, rc
    return out


def fixed_base_mul(base_mont, scalars_raw, nthreads=0):
    base_mont = np.ascontiguousarray(base_mont, dtype=np.uint64).reshape(12)
    scalars_raw = np.ascontiguousarray(scalars_raw, dtype=np.uint64).reshape(-1, 4)
    n = scalars_raw.shape[0]
    out = np.zeros((n, 12), dtype=np.uint64)
    rc = lib().oracle_g1_fixed_base_mul(_p(base_mont), _p(scalars_raw), n, _p(out), nthreads)
    assert rc == 0, rc
    return out


def _binop(name, a, b, nl):
    f = getattr(lib(), name)
    f.argtypes = [_u64p, _u64p, _u64p]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(nl)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(nl)
    out = np.zeros(nl, dtype=np.uint64)
    f(_p(a), _p(b), _p(out))
    return out


def fr_mul(a, b):
    return _binop("oracle_fr_mul", a, b, 4)


def fr_add(a, b):
    return _binop("oracle_fr_add", a, b, 4)


def fr_sub(a, b):
    return _binop("oracle_fr_sub", a, b, 4)


def fq_mul(a, b):
    return _binop("oracle_fq_mul", a, b, 6)
